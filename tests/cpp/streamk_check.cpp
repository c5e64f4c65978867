// CPU check of the stream-K schedule builder (mpstime.jl_b200/csrc/streamk.h): every (unit, chunk) is covered exactly
// once, CTAs get equal work, slots are contiguous per unit in ascending chunk order, for every walk order.
#include <cstdio>
#include <cstdlib>
#include <map>
#include "../../mpstime.jl_b200/csrc/streamk.h"

static int check(int ncta, int phases, int ncls, int nunits_per_cls, std::vector<int64_t> cb, std::vector<int64_t> ce) {
    std::vector<std::array<int, 3>> units;
    for (int c = 0; c < ncls; c++) for (int g = 0; g < nunits_per_cls; g++) units.push_back({c, g, 0});
    std::vector<GradSeg> segs; std::vector<int> hcta, hslot;
    build_streamk_table(ncta, phases, units, cb, ce, segs, hcta, hslot);
    if ((int)hcta.size() != ncta + 1 || hcta[0] != 0 || hcta[ncta] != (int)segs.size()) return 1;
    if ((int)hslot.size() != (int)units.size() + 1 || hslot.back() != (int)segs.size()) return 2;
    std::map<std::pair<int, int64_t>, int> cover;
    int64_t total = 0, wmin = 1LL << 60, wmax = 0;
    std::vector<int> seen(segs.size(), 0);
    for (int i = 0; i < ncta; i++) {
        if (hcta[i + 1] < hcta[i]) return 3;
        int64_t w = 0;
        for (int s = hcta[i]; s < hcta[i + 1]; s++) {
            const GradSeg& g = segs[s];
            if (g.chunk_end <= g.chunk_begin) return 4;
            const int u = g.cls * nunits_per_cls + g.tp;
            if (g.slot < hslot[u] || g.slot >= hslot[u + 1]) return 5;      // slot inside its unit's contiguous range
            if (seen[g.slot]++) return 6;
            for (int64_t j = g.chunk_begin; j < g.chunk_end; j++) cover[{u, j}]++;
            w += g.chunk_end - g.chunk_begin;
        }
        total += w; wmin = std::min(wmin, w); wmax = std::max(wmax, w);
    }
    int64_t expect = 0;
    for (int c = 0; c < ncls; c++) expect += (ce[c] - cb[c]) * nunits_per_cls;
    if (total != expect || (int64_t)cover.size() != expect) return 7;
    for (auto& kv : cover) if (kv.second != 1) return 8;
    if (wmax - wmin > 1) return 9;                                           // balanced to one chunk
    // slots of a unit ascend with the chunk position (fixed reduction order whatever the walk order)
    std::vector<int64_t> begin_of(segs.size());
    for (auto& g : segs) begin_of[g.slot] = g.chunk_begin;
    for (size_t u = 0; u < units.size(); u++)
        for (int s = hslot[u] + 1; s < hslot[u + 1]; s++) if (begin_of[s] <= begin_of[s - 1]) return 10;
    return 0;
}

int main() {
    int fails = 0;
    for (int phases : {0, 1, 4, 16, 37, -1, -2, -8})
        for (int ncta : {1, 7, 148}) {
            fails += check(ncta, phases, 2, 64, {0, 7813}, {7813, 15625}) != 0;       // config C per class, 64 groups
            fails += check(ncta, phases, 2, 13, {0, 782}, {782, 1563}) != 0;          // config B
            fails += check(ncta, phases, 3, 5, {0, 4, 4}, {5, 4, 6}) != 0;            // ragged, one empty class, shared boundary chunk
            fails += check(ncta, phases, 1, 1, {0}, {1}) != 0;                        // less work than CTAs
            int rc = check(ncta, phases, 2, 3, {10, 20}, {20, 33});
            if (rc) { printf("fail rc=%d ncta=%d phases=%d\n", rc, ncta, phases); fails++; }
        }
    printf(fails ? "STREAMK_FAIL %d\n" : "STREAMK_OK\n", fails);
    return fails != 0;
}
