// resp.cpp — response of a segment's exit heads to small perturbations of the true entry heads (which directions heal, which pass through)
#include "shoot_common.cpp"
int main(int argc, char** argv) {
    const std::string cfg = argc > 1 ? argv[1] : "c4";
    const uint32_t seg = argc > 2 ? atoi(argv[2]) : 448;
    const int batch = argc > 3 ? atoi(argv[3]) : 15;
    Loaded L = load_config(cfg);
    const uint32_t G = L.occ.size(), S = (G + seg - 1) / seg;
    size_t off = 0;
    for (int b = 0; b <= batch; ++b) {
        const uint32_t n = L.sizes[b];
        open_batch(L, off, n);
        std::vector<Heads> truth(S + 1); Heads h0{}; truth[0] = h0;
        std::vector<uint8_t> occ_new = L.occ;
        for (uint32_t s = 0; s < S; ++s) truth[s + 1] = simulate(L.occ, s * seg, std::min(G, (s + 1) * seg), truth[s], nullptr, &occ_new);
        if (b == batch) {
            const int np = (int)profs.size();
            struct Pert { const char* name; int d[6]; };
            Pert perts[] = {{"1g-2 2g+1", {-2,0,1,0,0,0}}, {"1g+2 2g-1", {2,0,-1,0,0,0}}, {"1g-8 2g+4", {-8,0,4,0,0,0}}, {"1g+8 2g-4", {8,0,-4,0,0,0}}, {"3g+1 4g-1", {0,0,0,1,-1,0}}, {"1g+1", {1,0,0,0,0,0}}, {"3g+1", {0,0,0,1,0,0}}, {"1g-4 3g+1", {-4,0,0,1,0,0}}};
            printf("lead T1-T2 :"); for (uint32_t s = 4; s < 60; s += 6) { long t1 = truth[s][0] < q[0].size() ? q[0][truth[s][0]] : 99999, t2 = truth[s][2] < q[2].size() ? q[2][truth[s][2]] : 99999, t3 = truth[s][3] < q[3].size() ? q[3][truth[s][3]] : 99999; printf(" [%ld | %ld]", t1 - t2, t1 - t3); } printf("\n");
            for (auto& pt : perts) {
                printf("%-10s:", pt.name);
                for (uint32_t s = 4; s < 60; s += 6) {
                    Heads h = truth[s]; bool ok = true;
                    for (int p = 0; p < np; ++p) { long v = (long)h[p] + pt.d[p]; if (v < 0 || v > (long)q[p].size()) ok = false; h[p] = (uint32_t)std::max(0l, v); }
                    if (!ok) { printf(" [--]"); continue; }
                    Heads e = simulate(L.occ, s * seg, std::min(G, (s + 1) * seg), h);
                    printf(" [");
                    for (int p = 0; p < np; ++p) if (pt.d[p] != 0 || e[p] != truth[s + 1][p]) printf("%ld ", (long)e[p] - (long)truth[s + 1][p]);
                    printf("]");
                }
                printf("\n");
            }
        }
        L.occ = occ_new; off += n;
    }
}
