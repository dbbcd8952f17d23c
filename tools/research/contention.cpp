// contention.cpp — how much of the chain is forced?  A decision is "forced" when exactly one pending profile fits on the current GPU
// (no comparison of request times decides it); runs of forced decisions of one profile could be committed by a prefix sum.
//   ./contention <c3|c4>
// Result: 37-40 % of the decisions are forced but the runs are 1.2-1.5 decisions long — nothing to collapse.
#include "shoot_common.cpp"
int main(int argc, char** argv) {
    const std::string cfg = argc > 1 ? argv[1] : "c4";
    Loaded L = load_config(cfg);
    const uint32_t G = L.occ.size();
    open_batch(L, 0, L.sizes[0]);
    const int np = L.np;
    Heads h{};
    uint64_t dec = 0, forced = 0, runs = 0, hist[9] = {0};
    int lastp = -2;
    for (uint32_t g = 0; g < G; ++g) {
        uint32_t o = L.occ[g]; int nd = 0;
        while (true) {
            uint32_t best = 0xFFFFFFFFu; int bp = -1, nfeas = 0; uint32_t bm = 0;
            for (int p = 0; p < np; ++p) {
                if (h[p] >= q[p].size()) continue;
                uint32_t m = 0;
                for (uint32_t mm : profs[p].masks) if ((o & mm) == 0) { m = mm; break; }
                if (!m) continue;
                ++nfeas;
                if (q[p][h[p]] < best) { best = q[p][h[p]]; bp = p; bm = m; }
            }
            if (bp < 0) break;
            ++dec; ++nd;
            if (nfeas == 1) { ++forced; if (bp != lastp) ++runs; lastp = bp; } else lastp = -2;
            o |= bm; ++h[bp];
        }
        hist[nd]++;
    }
    printf("%s: decisions %lu, forced %lu (%.1f %%), forced runs %lu (mean length %.2f); decisions per GPU:", cfg.c_str(), dec, forced, 100.0 * forced / dec, runs, (double)forced / runs);
    for (int i = 0; i < 9; ++i) printf(" %d:%lu", i, hist[i]);
    printf("\n");
}
