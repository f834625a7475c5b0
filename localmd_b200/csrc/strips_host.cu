// Host-side table builder for pmd_project_stream (K7 v3): packs the (block, component group) and background
// tasks of every column strip into the 8 warp slots of a CTA.  Pure host code (no device work); it lives in the
// library because the orchestration calls it once per decomposition with ~10^4 tasks, where a Python loop costs
// more than the projection kernel itself.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxRW = 48;

struct Task {
    int by, bx, h, w, col, nc, ncp, kind, key;
};
struct Item {
    int c0, rw, row0, row1, part;
    std::vector<Task> slots[kWarps];
};

// greedy interval packing into one pass; returns what did not fit
void pack_slots(const std::vector<Task>& tasks, std::vector<Task> (&slots)[kWarps], std::vector<Task>& left) {
    for (const Task& tk : tasks) {
        bool placed = false;
        for (int s = 0; s < kWarps; ++s) {
            if (slots[s].empty() || slots[s].back().by + slots[s].back().h <= tk.by) {
                slots[s].push_back(tk);
                placed = true;
                break;
            }
        }
        if (!placed) left.push_back(tk);
    }
}

// spread the slots' work over the four SM sub-partitions (warp w issues on sub-partition w % 4)
void balance(std::vector<Task> (&slots)[kWarps]) {
    long load[kWarps];
    int order[kWarps];
    for (int s = 0; s < kWarps; ++s) {
        load[s] = 0;
        for (const Task& tk : slots[s]) load[s] += (long)tk.h * tk.w * tk.ncp;
        order[s] = s;
    }
    std::stable_sort(order, order + kWarps, [&](int a, int b) { return load[a] > load[b]; });
    long bl[4] = {0, 0, 0, 0};
    int used[4] = {0, 0, 0, 0};
    std::vector<Task> out[kWarps];
    for (int i = 0; i < kWarps; ++i) {
        int best = -1;
        for (int b = 0; b < 4; ++b)
            if (used[b] < kWarps / 4 && (best < 0 || bl[b] < bl[best])) best = b;
        out[best + 4 * used[best]] = std::move(slots[order[i]]);
        bl[best] += load[order[i]];
        ++used[best];
    }
    for (int s = 0; s < kWarps; ++s) slots[s] = std::move(out[s]);
}

void add_item(std::vector<Item>& items, int c0, int rw, int part, std::vector<Task> (&slots)[kWarps]) {
    Item it;
    it.c0 = c0;
    it.rw = rw;
    it.part = part;
    it.row0 = 1 << 30;
    it.row1 = 0;
    for (int s = 0; s < kWarps; ++s)
        for (const Task& tk : slots[s]) {
            it.row0 = std::min(it.row0, tk.by);
            it.row1 = std::max(it.row1, tk.by + tk.h);
        }
    balance(slots);
    for (int s = 0; s < kWarps; ++s) it.slots[s] = std::move(slots[s]);
    items.push_back(std::move(it));
}

// all passes of one strip: the first takes what fits, leftovers are clustered by contiguous row coverage
void pack_passes(std::vector<Task> tasks, std::vector<Item>& items, int c0, int rw, int part) {
    std::vector<Task> left;
    {
        std::vector<Task> slots[kWarps];
        pack_slots(tasks, slots, left);
        add_item(items, c0, rw, part, slots);
    }
    while (!left.empty()) {
        std::vector<Task> cluster, rest, more;
        int end = -1;
        for (const Task& tk : left) {
            if (end < 0 || tk.by < end) {
                cluster.push_back(tk);
                end = std::max(end, tk.by + tk.h);
            } else {
                rest.push_back(tk);
            }
        }
        std::vector<Task> slots[kWarps];
        pack_slots(cluster, slots, more);
        add_item(items, c0, rw, part, slots);
        left = more;
        left.insert(left.end(), rest.begin(), rest.end());
        std::stable_sort(left.begin(), left.end(), [](const Task& a, const Task& b) { return a.by < b.by; });
    }
}

bool build(int g, const int32_t* rs, int nbr, const int32_t* cs, int nbc, int bh, int bw, int d1, int d2, const int64_t* ranks,
           const int64_t* col0, int n_bg, std::vector<Item>& items, long& streamed) {
    items.clear();
    streamed = 0;
    int part = 0;
    for (int ca = 0; ca < nbc; ca += g, ++part) {
        const int cb = std::min(ca + g, nbc);
        const int c0 = cs[ca], c1 = cs[cb - 1] + bw;
        if (c1 - c0 > kMaxRW) return false;
        const int core_end = cb < nbc ? cs[cb] : d2;
        std::vector<Task> tasks;
        int gi = 0;
        for (int k0 = 0; k0 < n_bg; k0 += 8, ++gi) {
            const int nc = std::min(8, n_bg - k0);
            tasks.push_back(Task{0, 0, d1, core_end - c0, k0, nc, nc > 4 ? 8 : 4, 1, gi});
        }
        for (int a = 0; a < nbr; ++a)   // ascending first row: already sorted
            for (int c = ca; c < cb; ++c) {
                const int rk = (int)ranks[(size_t)a * nbc + c];
                int first = (int)col0[(size_t)a * nbc + c];
                const int n = (rk + 7) / 8;
                for (int i = 0; i < n; ++i) {
                    const int nc = rk / n + (i < rk % n ? 1 : 0);
                    tasks.push_back(Task{rs[a], cs[c] - c0, bh, bw, first, nc, nc > 4 ? 8 : 4, 0, 0});
                    first += nc;
                }
            }
        const size_t before = items.size();
        pack_passes(std::move(tasks), items, c0, c1 - c0, part);
        // cost of a pass: rows x staged lanes (the kernel stages pixels in sets of 32 lanes)
        for (size_t i = before; i < items.size(); ++i)
            streamed += (long)((items[i].rw + 31) / 32 * 32) * (items[i].row1 - items[i].row0);
    }
    return true;
}

}  // namespace

// Host function (all pointers are HOST pointers).  Outputs: items [cap_items][8], slot_ptr [cap_items*9],
// tasks [cap_tasks][12], local8 / local4 [cap_tasks][2] int64 = (first column, n comps) in U-pack order,
// counts[8] = (n_items, n_slot_ptr, n_tasks, n_local8, n_local4, n_parts, max_rw, upack floats (low 31 bits ok: int64)).
extern "C" int pmd_make_strips(const int32_t* row_starts, int64_t nbr, const int32_t* col_starts, int64_t nbc, int64_t bh,
                               int64_t bw, int64_t d1, int64_t d2, const int64_t* ranks, const int64_t* col0, int64_t n_bg,
                               int64_t g_fixed, int32_t* items_out, int64_t cap_items, int32_t* slot_ptr_out,
                               int32_t* tasks_out, int64_t cap_tasks, int64_t* local8_out, int64_t* local4_out,
                               int64_t* counts) {
    const char* fn = "pmd_make_strips";
    PMD_REQUIRE(row_starts && col_starts && ranks && col0 && items_out && slot_ptr_out && tasks_out && local8_out && local4_out &&
                    counts,
                fn, "null pointer");
    PMD_REQUIRE(nbr > 0 && nbc > 0 && bh > 0 && bw > 0 && n_bg >= 0, fn, "bad size");
    if (bw > kMaxRW) {
        counts[0] = 0;
        return 0;
    }
    std::vector<Item> best, cur;
    long best_streamed = -1, streamed = 0;
    for (int g = g_fixed > 0 ? (int)g_fixed : 1; g <= (g_fixed > 0 ? (int)g_fixed : 8); ++g) {
        if (!build(g, row_starts, (int)nbr, col_starts, (int)nbc, (int)bh, (int)bw, (int)d1, (int)d2, ranks, col0, (int)n_bg, cur,
                   streamed))
            break;
        if (best_streamed < 0 || streamed < best_streamed) {
            best_streamed = streamed;
            best.swap(cur);
        }
    }
    if (best_streamed < 0) {
        counts[0] = 0;
        return 0;
    }
    const int64_t bpix = bh * bw;
    int64_t n8 = 0, n4 = 0, ntask = 0;
    for (const Item& it : best)
        for (int s = 0; s < kWarps; ++s)
            for (const Task& tk : it.slots[s]) {
                ++ntask;
                if (tk.kind == 0) (tk.ncp == 8 ? n8 : n4) += 1;
            }
    PMD_REQUIRE((int64_t)best.size() <= cap_items && ntask <= cap_tasks, fn, "output capacity too small");
    const int64_t base4 = n8 * bpix * 8, base_bg = base4 + n4 * bpix * 4;
    std::vector<int64_t> bg_off;
    int64_t off = base_bg;
    for (int k0 = 0; k0 < n_bg; k0 += 8) {
        const int nc = std::min<int64_t>(8, n_bg - k0);
        bg_off.push_back(off);
        off += d1 * d2 * (nc > 4 ? 8 : 4);
    }
    int64_t i8 = 0, i4 = 0, nt = 0, nsp = 0, nparts = 0, max_rw = 0;
    for (size_t i = 0; i < best.size(); ++i) {
        const Item& it = best[i];
        int32_t* io = items_out + 8 * i;
        io[0] = it.c0; io[1] = it.rw; io[2] = (int32_t)nsp; io[3] = it.row1 - it.row0; io[4] = it.part; io[5] = it.row0;
        io[6] = io[7] = 0;
        nparts = std::max<int64_t>(nparts, it.part + 1);
        max_rw = std::max<int64_t>(max_rw, it.rw);
        for (int s = 0; s < kWarps; ++s) {
            slot_ptr_out[nsp++] = (int32_t)nt;
            for (const Task& tk : it.slots[s]) {
                int64_t uoff;
                int urow;
                if (tk.kind == 0) {
                    if (tk.ncp == 8) {
                        uoff = i8 * bpix * 8;
                        local8_out[2 * i8] = tk.col; local8_out[2 * i8 + 1] = tk.nc;
                        ++i8;
                    } else {
                        uoff = base4 + i4 * bpix * 4;
                        local4_out[2 * i4] = tk.col; local4_out[2 * i4 + 1] = tk.nc;
                        ++i4;
                    }
                    urow = (int)bw * tk.ncp;
                } else {
                    uoff = bg_off[tk.key] + (int64_t)it.c0 * tk.ncp;
                    urow = (int)d2 * tk.ncp;
                }
                int32_t* to = tasks_out + 12 * nt;
                to[0] = tk.by; to[1] = tk.bx; to[2] = tk.h; to[3] = tk.w; to[4] = tk.col; to[5] = tk.nc; to[6] = tk.ncp;
                to[7] = urow; to[8] = (int32_t)(uint32_t)(uoff & 0xFFFFFFFFll); to[9] = (int32_t)(uoff >> 32); to[10] = tk.kind;
                to[11] = 0;
                ++nt;
            }
        }
        slot_ptr_out[nsp++] = (int32_t)nt;
    }
    counts[0] = (int64_t)best.size(); counts[1] = nsp; counts[2] = nt; counts[3] = i8; counts[4] = i4; counts[5] = nparts;
    counts[6] = max_rw; counts[7] = off;
    return 0;
}
