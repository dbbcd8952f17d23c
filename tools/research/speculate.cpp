// speculate.cpp — multiple shooting over inventory segments: every segment simulates the exact chain from GUESSED queue heads, the guesses
// are refined round by round (a round = all segments in parallel), a segment's result stands once its entry heads equal the true ones.
//   ./speculate <c3|c4> <segment GPUs> <mode>      mode 0: entry(s+1) = end(s) of the previous round; 1: prefix sums of the pops;
//                                                   2: prefix sums in round 1, then mode 0
// Reports rounds, re-simulations and the critical path in decisions (sum over rounds of the longest re-simulation) against the
// sequential chain.  Result (profiles/r02_speculation_study.md): from zero knowledge the guesses never become exact faster than
// the token travels — the rounds needed equal the busy segments.
#include "shoot_common.cpp"
int main(int argc, char** argv) {
    const std::string cfg = argc > 1 ? argv[1] : "c4";
    const uint32_t seg = argc > 2 ? atoi(argv[2]) : 512;
    const int mode = argc > 3 ? atoi(argv[3]) : 0;
    Loaded L = load_config(cfg);
    const uint32_t G = L.occ.size(), S = (G + seg - 1) / seg;
    size_t off = 0;
    for (size_t b = 0; b < L.sizes.size() && b < 4; ++b) {
        const uint32_t n = L.sizes[b];
        open_batch(L, off, n);
        std::vector<Heads> truth(S + 1); Heads h0{}; truth[0] = h0;
        std::vector<uint8_t> occ_new = L.occ; uint64_t dec = 0;
        for (uint32_t s = 0; s < S; ++s) truth[s + 1] = simulate(L.occ, s * seg, std::min(G, (s + 1) * seg), truth[s], &dec, &occ_new);
        std::vector<Heads> H(S + 1, h0), E(S + 1), Hn(S + 1), Pp(S);
        std::vector<bool> dirty(S, true);
        int rounds = 0; uint64_t recomputes = 0, crit = 0;
        while (true) {
            ++rounds;
            uint64_t maxw = 0;
            for (uint32_t s = 0; s < S; ++s) if (dirty[s]) {
                uint64_t d = 0;
                E[s + 1] = simulate(L.occ, s * seg, std::min(G, (s + 1) * seg), H[s], &d);
                for (int p = 0; p < P; ++p) Pp[s][p] = E[s + 1][p] - H[s][p];
                ++recomputes; maxw = std::max(maxw, d);
            }
            crit += maxw;
            Hn[0] = h0;
            const bool prefix = mode == 1 || (mode == 2 && rounds == 1);
            for (uint32_t s = 0; s < S; ++s) {
                if (prefix) for (int p = 0; p < P; ++p) Hn[s + 1][p] = std::min<uint32_t>(Hn[s][p] + Pp[s][p], (size_t)p < q.size() ? q[p].size() : 0);
                else Hn[s + 1] = E[s + 1];
            }
            bool any = false;
            for (uint32_t s = 0; s < S; ++s) { dirty[s] = Hn[s] != H[s]; any = any || dirty[s]; H[s] = Hn[s]; }
            if (!any || rounds > 100000) break;
        }
        bool ok = true; for (uint32_t s = 0; s < S; ++s) if (H[s] != truth[s]) ok = false;
        printf("%s batch %zu: sequential decisions %lu | rounds %d, re-simulations %lu, critical path %lu decisions (%.2fx the sequential chain) %s\n",
               cfg.c_str(), b, dec, rounds, recomputes, crit, (double)crit / dec, ok ? "exact" : "MISMATCH");
        L.occ = occ_new; off += n;
    }
}
