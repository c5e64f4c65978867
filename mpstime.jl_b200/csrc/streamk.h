// Host-side stream-K schedule shared by the two gradient kernels (bond_grad_kr.cu, bond_grad.cu).
//
// Work = a list of units (class, tp, tq), each of which must visit every sample chunk [cb[cls], ce[cls]) of its class.
// The unit-major linear work space is cut into one contiguous, equally long range per CTA (classic stream-K: every SM
// gets the same number of chunks whatever the unit count); each (unit, CTA) piece is a *segment* that accumulates into
// its own partial tile (`slot`), and a second kernel sums the slots of a unit in slot order (deterministic).
//
// What this file adds is the ORDER in which a CTA walks its pieces, which decides DRAM traffic: every unit re-reads the
// same raw sample rows, so co-resident CTAs should be reading the same chunks at about the same time and let L2 serve
// all but the first of them.
//   phases == 0 : unit-major order (a CTA finishes the tail of one pass, then starts the next at chunk 0): CTAs are
//                 spread evenly over the data set, nothing is shared (ncu: 54x the algorithmic bytes at d=16, chi=64).
//   phases == 1 : ascending chunk order: all CTAs sweep the data together -- 2x the algorithmic bytes, but 128 CTAs
//                 hitting the same L2 lines in lockstep cost 5 % of the kernel (hot-spotting on single L2 slices).
//   phases == R : ascending order, rotated by (cta mod R)/R of the CTA's own work: R fronts, ncta/R CTAs per front.
//   phases == -D: ascending order with a small stagger: CTA i starts (i mod 37) * D chunks into its walk (and wraps), so
//                 the CTAs still move as one front (two, where a CTA's range spans two units) but the front is
//                 smeared over 37*D chunks -- a few MB, far inside L2 -- instead of one hot line per L2 slice.
// Slots are numbered unit-major in ascending chunk order whatever the walk order, so the reduction order is fixed.
#pragma once
#include <algorithm>
#include <array>
#include <vector>
#include "mpst_common.cuh"

inline void build_streamk_table(int ncta, int phases, const std::vector<std::array<int, 3>>& units,
                                const std::vector<int64_t>& cb, const std::vector<int64_t>& ce,
                                std::vector<GradSeg>& hsegs, std::vector<int>& hcta, std::vector<int>& hslot) {
    struct Piece { int unit, cta; int64_t a, b; int order; };
    const int nu = (int)units.size();
    std::vector<int64_t> ubeg(nu + 1, 0);
    for (int u = 0; u < nu; u++) ubeg[u + 1] = ubeg[u] + (ce[units[u][0]] - cb[units[u][0]]);
    const int64_t total = ubeg[nu];
    std::vector<Piece> pieces;
    for (int i = 0; i < ncta; i++) {
        const int64_t lo = total * i / ncta, hi = total * (i + 1) / ncta;
        std::vector<Piece> mine;
        int u = (int)(std::upper_bound(ubeg.begin(), ubeg.end(), lo) - ubeg.begin()) - 1;
        for (int64_t p = lo; p < hi && u < nu;) {
            if (ubeg[u + 1] <= p) { u++; continue; }
            const int64_t take = std::min(hi, ubeg[u + 1]) - p;
            const int64_t a = cb[units[u][0]] + (p - ubeg[u]);
            mine.push_back({u, i, a, a + take, 0});
            p += take;
        }
        if (phases != 0) {
            std::stable_sort(mine.begin(), mine.end(), [](const Piece& x, const Piece& y) { return x.a < y.a; });
            int64_t n = 0;
            for (auto& m : mine) n += m.b - m.a;
            int64_t off = phases > 1 ? (int64_t)(i % phases) * n / phases : 0;
            if (phases < 0) off = std::min<int64_t>((int64_t)(i % 37) * -phases, n > 0 ? n - 1 : 0);
            if (off > 0) {
                // split the piece that contains `off` and start the walk there, wrapping around
                int64_t acc = 0;
                size_t k = 0;
                for (; k < mine.size(); k++) {
                    const int64_t len = mine[k].b - mine[k].a;
                    if (off < acc + len) break;
                    acc += len;
                }
                if (k < mine.size() && off > acc) {
                    Piece tail = mine[k];
                    tail.a = mine[k].a + (off - acc);
                    mine[k].b = tail.a;
                    mine.insert(mine.begin() + k + 1, tail);
                    k++;
                }
                std::rotate(mine.begin(), mine.begin() + std::min(k, mine.size()), mine.end());
            }
        }
        for (size_t k = 0; k < mine.size(); k++) { mine[k].order = (int)k; pieces.push_back(mine[k]); }
    }
    // slots: unit-major, ascending chunk
    std::vector<int> idx(pieces.size());
    for (size_t k = 0; k < idx.size(); k++) idx[k] = (int)k;
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) {
        return pieces[x].unit != pieces[y].unit ? pieces[x].unit < pieces[y].unit : pieces[x].a < pieces[y].a;
    });
    std::vector<int> slot_of(pieces.size());
    hslot.assign(nu + 1, 0);
    for (size_t s = 0; s < idx.size(); s++) { slot_of[idx[s]] = (int)s; hslot[pieces[idx[s]].unit + 1] = (int)s + 1; }
    for (int u = 0; u < nu; u++) if (hslot[u + 1] < hslot[u]) hslot[u + 1] = hslot[u];       // units without work
    // segments in CTA-major walk order (pieces were appended CTA by CTA, already in walk order)
    hsegs.clear();
    hcta.assign(ncta + 1, 0);
    for (size_t k = 0; k < pieces.size(); k++) {
        const Piece& p = pieces[k];
        GradSeg sg;
        sg.cls = units[p.unit][0]; sg.tp = units[p.unit][1]; sg.tq = units[p.unit][2]; sg.slot = slot_of[k];
        sg.chunk_begin = p.a; sg.chunk_end = p.b;
        hsegs.push_back(sg);
        hcta[p.cta + 1] = (int)hsegs.size();
    }
    for (int i = 0; i < ncta; i++) if (hcta[i + 1] < hcta[i]) hcta[i + 1] = hcta[i];
}
