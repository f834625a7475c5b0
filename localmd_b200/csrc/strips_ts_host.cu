// Host-side table builder for pmd_project_stream_ts (K7, movie operand in tensor memory).  Pure host code.
//
// The field of view is cut into column strips that PARTITION every image row exactly (width W = 32, 64, 96 or 128 pixels,
// first column a multiple of W): every movie element is streamed once.  A strip item = (strip, contiguous range of image
// rows, n_slots accumulator slots of 4 tensor-memory columns each).  A task = (block, group of <= 4 of its kept
// components) restricted to the strip's columns, or 4 of the dense background components restricted to the strip.  A
// block that straddles a strip border gives one task to each of its (at most two: W >= block width - 1) strips; those
// partial sums are ADDED to z by the kernel's epilogue (two contributions to a zeroed element commute, so the result
// does not depend on the order).  The tasks of a strip are packed into the slots by greedy interval scheduling over the
// rows they span; what does not fit goes to extra items that cover only the rows of the left-over tasks.
// (W, N = 4 n_slots) is chosen by a cost model: rows x 32-pixel chunks x max(HBM time, tensor time ~ N).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kMaxSlots = 48;
constexpr int kSlotCols = 4;

struct STask {
    int by, bx, h, w, col, nc, kind;    // kind 0: whole block inside the strip (store), 1: background, 2: partial block (add)
};
struct SItem {
    int c0, nkc, row0, row1, part;
    std::vector<STask> slots[kMaxSlots];
};

void first_fit(const std::vector<STask>& tasks, std::vector<STask>* slots, int n_slots, std::vector<STask>& left) {
    for (const STask& tk : tasks) {
        bool placed = false;
        for (int s = 0; s < n_slots && !placed; ++s) {
            if (slots[s].empty() || slots[s].back().by + slots[s].back().h <= tk.by) {
                slots[s].push_back(tk);
                placed = true;
            }
        }
        if (!placed) left.push_back(tk);
    }
}

void push_item(std::vector<SItem>& items, int c0, int nkc, int part, std::vector<STask>* slots, int n_slots) {
    SItem it;
    it.c0 = c0;
    it.nkc = nkc;
    it.part = part;
    it.row0 = 1 << 30;
    it.row1 = 0;
    for (int s = 0; s < n_slots; ++s) {
        for (const STask& tk : slots[s]) {
            it.row0 = std::min(it.row0, tk.by);
            it.row1 = std::max(it.row1, tk.by + tk.h);
        }
        it.slots[s] = std::move(slots[s]);
    }
    if (it.row1 > it.row0) items.push_back(std::move(it));
}

void pack_strip(std::vector<STask> tasks, std::vector<SItem>& items, int c0, int nkc, int part, int n_slots) {
    std::vector<STask> left;
    {
        std::vector<STask> slots[kMaxSlots];
        first_fit(tasks, slots, n_slots, left);
        push_item(items, c0, nkc, part, slots, n_slots);
    }
    while (!left.empty()) {
        // one extra pass per cluster of left-over tasks whose row ranges touch
        std::vector<STask> cluster, rest, more;
        int end = -1;
        for (const STask& tk : left) {
            if (end < 0 || tk.by < end) {
                cluster.push_back(tk);
                end = std::max(end, tk.by + tk.h);
            } else {
                rest.push_back(tk);
            }
        }
        std::vector<STask> slots[kMaxSlots];
        first_fit(cluster, slots, n_slots, more);
        push_item(items, c0, nkc, part, slots, n_slots);
        left = more;
        left.insert(left.end(), rest.begin(), rest.end());
        std::stable_sort(left.begin(), left.end(), [](const STask& a, const STask& b) { return a.by < b.by; });
    }
}

void build(int W, int n_slots, const int32_t* rs, int nbr, const int32_t* cs, int nbc, int bh, int bw, int d1, int d2,
           const int64_t* ranks, const int64_t* col0, int n_bg, std::vector<SItem>& items, double& cost) {
    items.clear();
    int part = 0;
    for (int c0 = 0; c0 < d2; c0 += W, ++part) {
        const int wpx = std::min(W, d2 - c0), nkc = (wpx + 31) / 32;
        std::vector<STask> tasks;
        for (int k0 = 0; k0 < n_bg; k0 += kSlotCols) tasks.push_back(STask{0, 0, d1, wpx, k0, std::min(kSlotCols, n_bg - k0), 1});
        for (int a = 0; a < nbr; ++a)
            for (int c = 0; c < nbc; ++c) {
                if (cs[c] >= c0 + wpx || cs[c] + bw <= c0) continue;
                const int rk = (int)ranks[(size_t)a * nbc + c];
                const int first = (int)col0[(size_t)a * nbc + c];
                const int kind = (cs[c] >= c0 && cs[c] + bw <= c0 + wpx) ? 0 : 2;
                for (int k0 = 0; k0 < rk; k0 += kSlotCols)
                    tasks.push_back(STask{rs[a], cs[c] - c0, bh, bw, first + k0, std::min(kSlotCols, rk - k0), kind});
            }
        pack_strip(std::move(tasks), items, c0, nkc, part, n_slots);
    }
    // per (row, 32-pixel chunk, 128-frame tile): HBM ~ 16 KB at one SM's share of the bandwidth (~730 cycles), tensor pipe
    // ~ 4 N cycles (+ issue / barrier overhead), overlapped; the coefficient chunk (2 N 128 bytes) is shared by the tiles
    const int N = kSlotCols * n_slots, tiles = 384 / N;
    const double per = std::max(730.0, 4.0 * N + 100.0) + (2.0 * N * 128 / 40.0) / tiles;
    cost = 0;
    for (const SItem& it : items) cost += per * (double)(it.row1 - it.row0) * it.nkc + 4000.0;
}

}  // namespace

// Host function (all pointers are HOST pointers).  Outputs (caller allocated):
//   items    [cap_items][12] int32 = (c0, chunks per row, row0, n_rows, first B chunk, first event, n events,
//                                    bg partial index, first slot_ptr entry, 0, n_main (full-height items, first), 0)
//   slot_ptr [cap_items*49]  int32 : tasks of slot s of item i are slot_ptr[i*49+s] .. slot_ptr[i*49+s+1]-1
//   tasks    [cap_tasks][8]  int32 = (first row, first column relative to c0 (may be negative), rows, width, first output
//                                    column, n comps 1..4, kind, 0)
//   events   [cap_events][4] int32 = (row, slot, first output column, n comps | kind << 8), per item ascending in row:
//                                    after that row the slot is read and cleared; kind 0 stores finished columns, kind 1
//                                    adds a partial sum of background columns to the strip's partial buffer, kind 2 adds
//                                    (atomically) the partial sum of a block that straddles two strips
//   counts[8] = (n_items, n_tasks, n_events, total B chunks, n_parts, W, N, frame tiles per CTA); counts[0] == 0: not supported
extern "C" int pmd_make_strips_ts(const int32_t* row_starts, int64_t nbr, const int32_t* col_starts, int64_t nbc, int64_t bh,
                                  int64_t bw, int64_t d1, int64_t d2, const int64_t* ranks, const int64_t* col0, int64_t n_bg,
                                  int64_t w_fixed, int64_t n_fixed, int32_t* items_out, int64_t cap_items, int32_t* slot_ptr_out,
                                  int32_t* tasks_out, int64_t cap_tasks, int32_t* events_out, int64_t cap_events, int64_t* counts) {
    const char* fn = "pmd_make_strips_ts";
    PMD_REQUIRE(row_starts && col_starts && ranks && col0 && items_out && slot_ptr_out && tasks_out && events_out && counts, fn,
                "null pointer");
    PMD_REQUIRE(nbr > 0 && nbc > 0 && bh > 0 && bw > 0 && n_bg >= 0, fn, "bad size");
    for (int i = 0; i < 8; ++i) counts[i] = 0;
    if (bw - 1 > 128) return 0;   // a block may touch at most two strips
    std::vector<SItem> best, cur;
    double best_cost = -1, cost = 0;
    int best_w = 0, best_n = 0;
    // Candidates: the two narrowest strip widths a block fits in (wider strips multiply every 32-pixel chunk with more
    // slot columns that are zero for it), and slot counts in ascending order until one packs every task into the
    // full-height items (more slots then only cost tensor time).
    for (int N : {96, 128, 192}) {
        if (n_fixed > 0 && N != n_fixed) continue;
        if ((n_bg + kSlotCols - 1) / kSlotCols >= N / kSlotCols) continue;   // the background alone fills the slots
        if (best_cost >= 0 && n_fixed <= 0) {
            // lower bound of this N: every image row of every 32-pixel chunk exactly once, no left-over items
            const int tiles = 384 / N;
            const double per = std::max(730.0, 4.0 * N + 100.0) + (2.0 * N * 128 / 40.0) / tiles;
            if (per * (double)d1 * (double)((d2 + 31) / 32) >= best_cost) break;
        }
        int tried = 0;
        bool all_fit = false;
        for (int W : {32, 64, 96, 128}) {
            if (w_fixed > 0 && W != w_fixed) continue;
            if (bw - 1 > W) continue;
            if (w_fixed <= 0 && tried >= 2) break;
            ++tried;
            build(W, N / kSlotCols, row_starts, (int)nbr, col_starts, (int)nbc, (int)bh, (int)bw, (int)d1, (int)d2, ranks, col0,
                  (int)n_bg, cur, cost);
            if (tried == 1 && (int64_t)cur.size() * W >= d2 && (int64_t)(cur.size() - 1) * W < d2) all_fit = true;
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best_w = W;
                best_n = N;
                best.swap(cur);
            }
        }
        if (all_fit && n_fixed <= 0) break;
    }
    if (best_cost < 0) return 0;
    const int n_slots = best_n / kSlotCols;
    // items keep strip order (strip index fastest over the grid: the CTAs that run together cover whole image rows, which
    // is what lets the 128-byte-wide TMA boxes stream at full HBM bandwidth); extra items go last
    std::stable_sort(best.begin(), best.end(), [](const SItem& a, const SItem& b) {
        return (long)(a.row1 - a.row0) > (long)(b.row1 - b.row0);
    });
    int64_t ntask = 0;
    for (const SItem& it : best)
        for (int s = 0; s < n_slots; ++s) ntask += (int64_t)it.slots[s].size();
    PMD_REQUIRE((int64_t)best.size() <= cap_items && ntask <= cap_tasks, fn, "output capacity too small");
    int64_t nt = 0, nev = 0, chunks = 0, nparts = 0;
    int n_main = 0;   // items that span the tallest row range (one per strip when every strip has blocks over its full height)
    for (const SItem& it : best) n_main += (it.row1 - it.row0) == (best[0].row1 - best[0].row0);
    for (size_t i = 0; i < best.size(); ++i) {
        const SItem& it = best[i];
        int32_t* io = items_out + 12 * i;
        struct Ev { int row, slot, col, ncw; };
        std::vector<Ev> evs;
        std::vector<int> drain_rows;             // rows at which local tasks end (the MMA stream pauses there anyway)
        for (int s = 0; s <= kMaxSlots; ++s) slot_ptr_out[i * (kMaxSlots + 1) + s] = (int32_t)nt;
        for (int s = 0; s < n_slots; ++s) {
            slot_ptr_out[i * (kMaxSlots + 1) + s] = (int32_t)nt;
            for (const STask& tk : it.slots[s]) {
                int32_t* to = tasks_out + 8 * nt;
                to[0] = tk.by; to[1] = tk.bx; to[2] = tk.h; to[3] = tk.w; to[4] = tk.col; to[5] = tk.nc; to[6] = tk.kind; to[7] = 0;
                if (tk.kind != 1) {
                    evs.push_back(Ev{tk.by + tk.h - 1, s, tk.col, tk.nc | (tk.kind << 8)});
                    drain_rows.push_back(tk.by + tk.h - 1);
                }
                ++nt;
            }
        }
        for (int s = n_slots; s <= kMaxSlots; ++s) slot_ptr_out[i * (kMaxSlots + 1) + s] = (int32_t)nt;
        // Background tasks accumulate over every row of the strip.  The tensor core adds into its float32 accumulators
        // with truncation, a bias that grows with the number of accumulation steps, so the background slots are
        // drained (added to the partial sums in float32 by the epilogue) after at most kMaxChain rows -- at rows where
        // local tasks end (the MMA stream pauses there anyway) whenever the next such row would be too late.
        constexpr int kMaxChain = 20;
        drain_rows.push_back(it.row1 - 1);
        std::sort(drain_rows.begin(), drain_rows.end());
        drain_rows.erase(std::unique(drain_rows.begin(), drain_rows.end()), drain_rows.end());
        std::vector<int> bg_rows;
        int last = it.row0 - 1;
        for (size_t q = 0; q < drain_rows.size(); ++q) {
            const int r = drain_rows[q];
            while (r - last > kMaxChain) {              // no local task ends in time: a drain of the background alone
                last += kMaxChain;
                bg_rows.push_back(last);
            }
            if (q + 1 == drain_rows.size() || drain_rows[q + 1] - last > kMaxChain) {
                bg_rows.push_back(r);
                last = r;
            }
        }
        for (int s = 0; s < n_slots; ++s)
            for (const STask& tk : it.slots[s])
                if (tk.kind == 1)
                    for (int r : bg_rows)
                        if (r >= tk.by && r < tk.by + tk.h) evs.push_back(Ev{r, s, tk.col, tk.nc | (1 << 8)});
        std::stable_sort(evs.begin(), evs.end(), [](const Ev& a, const Ev& b) { return a.row < b.row; });
        io[0] = it.c0; io[1] = it.nkc; io[2] = it.row0; io[3] = it.row1 - it.row0; io[4] = (int32_t)chunks; io[5] = (int32_t)nev;
        io[6] = (int32_t)evs.size(); io[7] = it.part; io[8] = (int32_t)(i * (kMaxSlots + 1)); io[9] = io[11] = 0; io[10] = n_main;
        PMD_REQUIRE(nev + (int64_t)evs.size() <= cap_events, fn, "event capacity too small");
        for (const Ev& e : evs) {
            int32_t* eo = events_out + 4 * nev++;
            eo[0] = e.row; eo[1] = e.slot; eo[2] = e.col; eo[3] = e.ncw;
        }
        chunks += (int64_t)(it.row1 - it.row0) * it.nkc;
        PMD_REQUIRE(chunks * (2ll * best_n * 128) < (1ll << 40), fn, "coefficient image too large");
        nparts = std::max<int64_t>(nparts, it.part + 1);
    }
    counts[0] = (int64_t)best.size(); counts[1] = nt; counts[2] = nev; counts[3] = chunks; counts[4] = nparts; counts[5] = best_w;
    counts[6] = best_n; counts[7] = 384 / best_n;
    return 0;
}
