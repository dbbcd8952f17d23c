// coalesce.cpp — does a chain started from slightly WRONG queue heads merge with the true one?  Perturb the true heads at GPU g0 and
// count GPUs until the perturbed trajectory carries the same heads as the true one (from there on both are identical for good).
//   ./coalesce <c3|c4>
// Result (profiles/r02_speculation_study.md): perturbations that keep the consumed slices constant AND stay inside one contention
// class (two 1g for one 2g; a 3g for a 4g) merge within tens of GPUs; every other perturbation — one request more or less of any
// profile, a 1g traded against a quad — survives for thousands of GPUs or to the end of the batch: the number of slices consumed
// before a GPU and the number of low quads taken are conserved quantities, a guess must hit them exactly.
#include "shoot_common.cpp"
int main(int argc, char** argv) {
    const std::string cfg = argc > 1 ? argv[1] : "c4";
    Loaded L = load_config(cfg);
    const uint32_t G = L.occ.size();
    open_batch(L, 0, L.sizes[0]);
    std::vector<Heads> T(G + 1); T[0] = Heads{};
    for (uint32_t g = 0; g < G; ++g) T[g + 1] = simulate(L.occ, g, g + 1, T[g]);
    struct Pert { int d[4]; const char* name; };
    const Pert perts[] = {{{1,0,0,0}, "1g +1"}, {{2,-1,0,0}, "1g +2, 2g -1 (same slices)"}, {{0,0,1,-1}, "3g +1, 4g -1 (same mask)"}, {{4,0,-1,0}, "1g +4, 3g -1 (same slices)"},
                          {{20,-10,0,0}, "1g +20, 2g -10"}, {{40,0,-10,0}, "1g +40, 3g -10"}, {{30,10,5,5}, "all ahead"}};
    const int pidx[4] = {0, 2, 3, 4};       // rows of 1g.10gb, 2g.20gb, 3g.40gb, 4g.40gb in the 80GB-class table
    const uint32_t limit = cfg == "c4" ? 28000 : 3000;
    for (const Pert& pt : perts) {
        std::vector<uint32_t> dist;
        for (uint32_t g0 = 64; g0 < limit; g0 += limit / 40) {
            Heads h = T[g0]; bool okp = true;
            for (int k = 0; k < 4; ++k) { const long v = (long)h[pidx[k]] + pt.d[k]; if (v < 0 || v > (long)q[pidx[k]].size()) okp = false; else h[pidx[k]] = (uint32_t)v; }
            if (!okp) continue;
            uint32_t g = g0;
            while (g < G && h != T[g]) { h = simulate(L.occ, g, g + 1, h); ++g; }
            dist.push_back(g - g0);
        }
        std::sort(dist.begin(), dist.end());
        printf("%s  %-28s GPUs until merged: min %u  median %u  p90 %u  max %u\n", cfg.c_str(), pt.name, dist[0], dist[dist.size() / 2], dist[dist.size() * 9 / 10], dist.back());
    }
}
