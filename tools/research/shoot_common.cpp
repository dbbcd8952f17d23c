// Research prototypes (CPU only) behind DESIGN.md section 4.5 "Why one batch does not parallelise": can the exact GPU-major first-fit
// chain be cut into segments that run speculatively from guessed queue heads and are kept when the guess was right?
// Inputs: binary dumps written by dump_workloads.py into $ISL_PROTO_DIR (default /tmp/isl_proto).  Nothing here is product code.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <array>
#include <algorithm>
#include <string>
#include "../../include/islplace.h"

static std::vector<uint8_t> readfile(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb"); if (!f) { perror(p.c_str()); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> v(n); if (fread(v.data(), 1, n, f) != (size_t)n) exit(1); fclose(f); return v;
}
static uint32_t cmask(uint32_t size, uint32_t v) {      // REF_EXACT
    if (v >= 8 || size == 0 || size > 8) return 0;
    if (size == 1) return 1u << v;
    if (!(size == 2 || size == 4 || size == 8)) return 0;
    if (!(v + size < 8)) return 0;
    return (((1u << size) - 1u) << v) & 0xFFu;
}
constexpr int P = 16;
typedef std::array<uint32_t, P> Heads;
struct Prof { std::vector<uint32_t> masks; };
static std::vector<Prof> profs;
static std::vector<std::vector<uint32_t>> q;   // per profile request times

// simulate GPUs [lo,hi) from entry heads; occ is the chunk-start occupancy (not modified); returns end heads; counts decisions
static Heads simulate(const std::vector<uint8_t>& occ, uint32_t lo, uint32_t hi, Heads h, uint64_t* dec = nullptr, std::vector<uint8_t>* occ_out = nullptr) {
    const int np = (int)profs.size();
    for (uint32_t g = lo; g < hi; ++g) {
        uint32_t o = occ[g];
        while (true) {
            uint32_t best = 0xFFFFFFFFu; int bp = -1; uint32_t bm = 0;
            for (int p = 0; p < np; ++p) {
                if (h[p] >= q[p].size()) continue;
                uint32_t m = 0;
                for (uint32_t mm : profs[p].masks) if ((o & mm) == 0) { m = mm; break; }
                if (!m) continue;
                const uint32_t t = q[p][h[p]];
                if (t < best) { best = t; bp = p; bm = m; }
            }
            if (bp < 0) break;
            o |= bm; ++h[bp]; if (dec) ++*dec;
        }
        if (occ_out) (*occ_out)[g] = (uint8_t)o;
    }
    return h;
}


static std::string proto_dir() { const char* d = getenv("ISL_PROTO_DIR"); return std::string(d ? d : "/tmp/isl_proto") + "/"; }

struct Loaded { std::vector<uint8_t> occ; std::vector<uint32_t> sizes; std::vector<isl_request> req; int np; };
static Loaded load_config(const std::string& cfg) {
    Loaded L;
    L.occ = readfile(proto_dir() + cfg + "_occ0.bin");
    auto sz = readfile(proto_dir() + cfg + "_sizes.bin"), rq = readfile(proto_dir() + cfg + "_req.bin"), rw = readfile(proto_dir() + "rows.bin");
    L.sizes.assign((const uint32_t*)sz.data(), (const uint32_t*)(sz.data() + sz.size()));
    L.req.assign((const isl_request*)rq.data(), (const isl_request*)(rq.data() + rq.size()));
    const isl_profile* rows = (const isl_profile*)rw.data();
    L.np = rw.size() / sizeof(isl_profile);
    profs.assign(L.np, {});
    for (int p = 0; p < L.np; ++p) for (int k = 0; k < rows[p].n_starts; ++k) { uint32_t m = cmask(rows[p].size, rows[p].starts[k]); if (m) profs[p].masks.push_back(m); }
    return L;
}
// apply the FREEs of batch [off, off+n) and build the per-profile queues of its ALLOCs
static void open_batch(Loaded& L, size_t off, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) { const isl_request& r = L.req[off + i]; if (r.op == ISL_OP_FREE) L.occ[r.handle] &= ~((((1u << r.size) - 1u) << r.start)); }
    q.assign(L.np, {});
    for (uint32_t i = 0; i < n; ++i) { const isl_request& r = L.req[off + i]; if (r.op == ISL_OP_ALLOC && r.profile < L.np && !profs[r.profile].masks.empty()) q[r.profile].push_back(i); }
}
