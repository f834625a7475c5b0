// Host-side table builder for pmd_project_stream_tc (K7 on tcgen05).  Pure host code.
//
// A strip item = G neighbouring block columns (<= 128 pixels wide after aligning its first column down to a
// multiple of 4 pixels and its width up to a multiple of 8), a contiguous range of image rows, and kTCSlots = 32
// accumulator slots of 4 tensor-memory columns each.  A task = (block, group of <= 4 of its kept components), or
// 4 of the dense background components restricted to the strip's own ("core") columns; the tasks of a strip are
// packed into the slots by greedy interval scheduling over the rows they span.  What does not fit goes to extra
// items (passes) that cover only the rows of the left-over tasks.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kSlots = 32;
constexpr int kSlotCols = 4;
constexpr int kMaxW = 128;

struct TTask {
    int by, bx, h, w, col, nc, kind;
};
struct TItem {
    int c0, w8, row0, row1, part;
    std::vector<TTask> slots[kSlots];
};

void first_fit(const std::vector<TTask>& tasks, std::vector<TTask> (&slots)[kSlots], std::vector<TTask>& left) {
    for (const TTask& tk : tasks) {
        bool placed = false;
        for (int s = 0; s < kSlots && !placed; ++s) {
            if (slots[s].empty() || slots[s].back().by + slots[s].back().h <= tk.by) {
                slots[s].push_back(tk);
                placed = true;
            }
        }
        if (!placed) left.push_back(tk);
    }
}

void push_item(std::vector<TItem>& items, int c0, int w8, int part, std::vector<TTask> (&slots)[kSlots]) {
    TItem it;
    it.c0 = c0;
    it.w8 = w8;
    it.part = part;
    it.row0 = 1 << 30;
    it.row1 = 0;
    for (int s = 0; s < kSlots; ++s) {
        for (const TTask& tk : slots[s]) {
            it.row0 = std::min(it.row0, tk.by);
            it.row1 = std::max(it.row1, tk.by + tk.h);
        }
        it.slots[s] = std::move(slots[s]);
    }
    if (it.row1 > it.row0) items.push_back(std::move(it));
}

void pack_strip(std::vector<TTask> tasks, std::vector<TItem>& items, int c0, int w8, int part) {
    std::vector<TTask> left;
    {
        std::vector<TTask> slots[kSlots];
        first_fit(tasks, slots, left);
        push_item(items, c0, w8, part, slots);
    }
    while (!left.empty()) {
        // one extra pass per cluster of left-over tasks whose row ranges touch
        std::vector<TTask> cluster, rest, more;
        int end = -1;
        for (const TTask& tk : left) {
            if (end < 0 || tk.by < end) {
                cluster.push_back(tk);
                end = std::max(end, tk.by + tk.h);
            } else {
                rest.push_back(tk);
            }
        }
        std::vector<TTask> slots[kSlots];
        first_fit(cluster, slots, more);
        push_item(items, c0, w8, part, slots);
        left = more;
        left.insert(left.end(), rest.begin(), rest.end());
        std::stable_sort(left.begin(), left.end(), [](const TTask& a, const TTask& b) { return a.by < b.by; });
    }
}

// returns false when a strip of g block columns is wider than the kernel supports
bool build(int g, const int32_t* rs, int nbr, const int32_t* cs, int nbc, int bh, int bw, int d1, int d2, const int64_t* ranks,
           const int64_t* col0, int n_bg, std::vector<TItem>& items, long& cost) {
    items.clear();
    cost = 0;
    int part = 0;
    for (int ca = 0; ca < nbc; ca += g, ++part) {
        const int cb = std::min(ca + g, nbc);
        // first column: aligned down to a 128-byte line of float32 pixels (32 pixels) when the strip stays within the
        // supported width -- a warp-level load of the kernel then covers whole lines --, else to 16 bytes (4 pixels)
        int a0 = cs[ca] & ~31;
        int width = (cs[cb - 1] + bw - a0 + 7) / 8 * 8;
        if (width > kMaxW) {
            a0 = cs[ca] & ~3;
            width = (cs[cb - 1] + bw - a0 + 7) / 8 * 8;
        }
        if (width > kMaxW) return false;
        if (a0 + width > d2) a0 = d2 - width;     // shift left: the extra columns get zero coefficients
        if (a0 < 0 || (a0 & 3)) return false;
        const int core_lo = cs[ca], core_hi = cb < nbc ? cs[cb] : d2;
        std::vector<TTask> tasks;
        for (int k0 = 0; k0 < n_bg; k0 += kSlotCols)
            tasks.push_back(TTask{0, core_lo - a0, d1, core_hi - core_lo, k0, std::min(kSlotCols, n_bg - k0), 1});
        for (int a = 0; a < nbr; ++a)
            for (int c = ca; c < cb; ++c) {
                const int rk = (int)ranks[(size_t)a * nbc + c];
                const int first = (int)col0[(size_t)a * nbc + c];
                for (int k0 = 0; k0 < rk; k0 += kSlotCols)
                    tasks.push_back(TTask{rs[a], cs[c] - a0, bh, bw, first + k0, std::min(kSlotCols, rk - k0), 0});
            }
        const size_t before = items.size();
        pack_strip(std::move(tasks), items, a0, width / 8, part);
        // movie bytes ~ w8, coefficient-image bytes (L2 resident) ~ 32-pixel chunks, weighted half
        for (size_t i = before; i < items.size(); ++i)
            cost += (long)(items[i].w8 * 2 + (items[i].w8 + 3) / 4 * 2) * (items[i].row1 - items[i].row0);
    }
    return true;
}

}  // namespace

// Host function (all pointers are HOST pointers).  Outputs (caller allocated):
//   items    [cap_items][12] int32 = (c0, w8, row0, n_rows, first B chunk, chunks per row, first event, n events,
//                                    bg partial index, first slot_ptr entry, 0, 0)
//   slot_ptr [cap_items*33]  int32 : tasks of slot s of item i are slot_ptr[i*33+s] .. slot_ptr[i*33+s+1]-1
//   tasks    [cap_tasks][8]  int32 = (first row, first column relative to c0, rows, width, first output column,
//                                    n comps 1..4, kind 0 local | 1 background, 0)
//   events   [cap_events][4] int32 = (row, slot, first output column, n comps | kind << 8), per item ascending in
//                                    row: after that row the slot is read and cleared; kind 0 stores the finished
//                                    local columns, kind 1 ADDS a partial sum of background columns
//   counts[8] = (n_items, n_tasks, n_events, total B chunks, n_parts, max w8, G, 0); counts[0] == 0: not supported
extern "C" int pmd_make_strips_tc(const int32_t* row_starts, int64_t nbr, const int32_t* col_starts, int64_t nbc, int64_t bh,
                                  int64_t bw, int64_t d1, int64_t d2, const int64_t* ranks, const int64_t* col0, int64_t n_bg,
                                  int64_t g_fixed, int32_t* items_out, int64_t cap_items, int32_t* slot_ptr_out,
                                  int32_t* tasks_out, int64_t cap_tasks, int32_t* events_out, int64_t cap_events, int64_t* counts) {
    const char* fn = "pmd_make_strips_tc";
    PMD_REQUIRE(row_starts && col_starts && ranks && col0 && items_out && slot_ptr_out && tasks_out && events_out && counts, fn,
                "null pointer");
    PMD_REQUIRE(nbr > 0 && nbc > 0 && bh > 0 && bw > 0 && n_bg >= 0, fn, "bad size");
    for (int i = 0; i < 8; ++i) counts[i] = 0;
    if ((d2 & 3) || bw + 3 > kMaxW) return 0;
    std::vector<TItem> best, cur;
    long best_cost = -1, cost = 0;
    int best_g = 0;
    // The cost falls with the strip width (fewer halo columns) until the live tasks overflow the slots (extra passes):
    // the optimum sits near g* = free slots / (live block rows x tasks per block); widths in [g*/2, 2 g*] are tried.
    int g_lo = 1, g_hi = (int)std::min<int64_t>(nbc, 64);
    if (g_fixed > 0) {
        g_lo = g_hi = (int)g_fixed;
    } else {
        double tasks = 0;
        for (int64_t i = 0; i < nbr * nbc; ++i) tasks += (double)((ranks[i] + kSlotCols - 1) / kSlotCols);
        const double per_block = std::max(1.0, tasks / (double)(nbr * nbc));
        const double live_rows = (double)bh / (double)std::max<int64_t>(1, nbr > 1 ? row_starts[1] - row_starts[0] : bh);
        const double free_slots = kSlots - (double)((n_bg + kSlotCols - 1) / kSlotCols);
        const int g_est = std::max(1, (int)(free_slots / (live_rows * per_block)));
        g_lo = std::max(1, g_est / 2);
        g_hi = std::min(g_hi, 2 * g_est + 1);
    }
    for (int g = g_lo; g <= g_hi; ++g) {
        if (!build(g, row_starts, (int)nbr, col_starts, (int)nbc, (int)bh, (int)bw, (int)d1, (int)d2, ranks, col0, (int)n_bg, cur,
                   cost))
            break;   // wider strips do not fit the kernel either
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best_g = g;
            best.swap(cur);
        }
    }
    if (best_cost < 0 && g_fixed <= 0) {   // nothing in the window fitted (very wide blocks): fall back to the narrowest strips
        for (int g = g_lo - 1; g >= 1 && best_cost < 0; --g)
            if (build(g, row_starts, (int)nbr, col_starts, (int)nbc, (int)bh, (int)bw, (int)d1, (int)d2, ranks, col0, (int)n_bg, cur,
                      cost)) {
                best_cost = cost;
                best_g = g;
                best.swap(cur);
            }
    }
    if (best_cost < 0) return 0;
    // largest items first: the grid is scheduled in item order, so the short row-range items fill the tail of the last wave
    std::stable_sort(best.begin(), best.end(), [](const TItem& a, const TItem& b) {
        return (long)a.w8 * (a.row1 - a.row0) > (long)b.w8 * (b.row1 - b.row0);
    });
    int64_t ntask = 0;
    for (const TItem& it : best)
        for (int s = 0; s < kSlots; ++s) ntask += (int64_t)it.slots[s].size();
    PMD_REQUIRE((int64_t)best.size() <= cap_items && ntask <= cap_tasks, fn, "output capacity too small");
    int64_t nt = 0, nev = 0, chunks = 0, nparts = 0, max_w8 = 0;
    for (size_t i = 0; i < best.size(); ++i) {
        const TItem& it = best[i];
        const int nkc = (it.w8 + 3) / 4;
        int32_t* io = items_out + 12 * i;
        struct Ev { int row, slot, col, ncw; };
        std::vector<Ev> evs;
        std::vector<int> drain_rows;             // rows at which local tasks end (the MMA stream pauses there anyway)
        for (int s = 0; s < kSlots; ++s) {
            slot_ptr_out[i * (kSlots + 1) + s] = (int32_t)nt;
            for (const TTask& tk : it.slots[s]) {
                int32_t* to = tasks_out + 8 * nt;
                to[0] = tk.by; to[1] = tk.bx; to[2] = tk.h; to[3] = tk.w; to[4] = tk.col; to[5] = tk.nc; to[6] = tk.kind; to[7] = 0;
                if (tk.kind == 0) {
                    evs.push_back(Ev{tk.by + tk.h - 1, s, tk.col, tk.nc});
                    drain_rows.push_back(tk.by + tk.h - 1);
                }
                ++nt;
            }
        }
        slot_ptr_out[i * (kSlots + 1) + kSlots] = (int32_t)nt;
        // Background tasks accumulate over every row of the strip.  The tensor core adds into its float32 accumulators
        // with truncation, a bias that grows with the number of accumulation steps, so the background slots are
        // drained (added to the partial sums in float32 by the epilogue) at every row where local tasks end, and at
        // least every kMaxChain rows.
        constexpr int kMaxChain = 16;
        drain_rows.push_back(it.row1 - 1);
        std::sort(drain_rows.begin(), drain_rows.end());
        drain_rows.erase(std::unique(drain_rows.begin(), drain_rows.end()), drain_rows.end());
        std::vector<int> bg_rows;
        int last = it.row0 - 1;
        for (int r : drain_rows) {
            while (r - last > kMaxChain) {
                last += kMaxChain;
                bg_rows.push_back(last);
            }
            bg_rows.push_back(r);
            last = r;
        }
        for (int s = 0; s < kSlots; ++s)
            for (const TTask& tk : it.slots[s])
                if (tk.kind == 1)
                    for (int r : bg_rows)
                        if (r >= tk.by && r < tk.by + tk.h) evs.push_back(Ev{r, s, tk.col, tk.nc | (1 << 8)});
        std::stable_sort(evs.begin(), evs.end(), [](const Ev& a, const Ev& b) { return a.row < b.row; });
        io[0] = it.c0; io[1] = it.w8; io[2] = it.row0; io[3] = it.row1 - it.row0; io[4] = (int32_t)chunks; io[5] = nkc;
        io[6] = (int32_t)nev; io[7] = (int32_t)evs.size(); io[8] = it.part; io[9] = (int32_t)(i * (kSlots + 1)); io[10] = io[11] = 0;
        PMD_REQUIRE(nev + (int64_t)evs.size() <= cap_events, fn, "event capacity too small");
        for (const Ev& e : evs) {
            int32_t* eo = events_out + 4 * nev++;
            eo[0] = e.row; eo[1] = e.slot; eo[2] = e.col; eo[3] = e.ncw;
        }
        chunks += (int64_t)(it.row1 - it.row0) * nkc;
        PMD_REQUIRE(chunks < (1ll << 31), fn, "coefficient image too large");
        nparts = std::max<int64_t>(nparts, it.part + 1);
        max_w8 = std::max<int64_t>(max_w8, it.w8);
    }
    counts[0] = (int64_t)best.size(); counts[1] = nt; counts[2] = nev; counts[3] = chunks; counts[4] = nparts; counts[5] = max_w8;
    counts[6] = best_g;
    return 0;
}
