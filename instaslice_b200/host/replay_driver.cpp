// replay_driver.cpp — native open-loop replay harness for BASELINE config 5 (samples/vllm_dep.yaml: a Deployment whose
// replicas each ask for one nvidia.com/mig-3g.20gb).  Part of libislhost.so; decides nothing.
//
// The reference resolves one pod per Reconcile (instaslice_controller.go:188-232); what a user sees is the time from a pod
// becoming pending to its allocation being known.  This driver replays a fixed arrival / lifetime trace against the wall
// clock: every turn of its loop hands the placer every slice that has expired (as a FREE, the daemonset deleting the
// Allocations entry, instaslice_daemonset.go:261-263) and every request that has arrived since the previous call, through
// ONE call of `place` (isl_place_batch for the engine; the CPU oracle's entry point for the baseline leg of bench.py),
// and stamps submit -> result latency per request.  A Python loop around ctypes adds ~10 us per call; this one adds none.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

#include "../../include/islplace.h"

extern "C" {

typedef int (*islh_place_fn)(void* ctx, uint32_t n, const isl_request* in, isl_result* out);

// arrivals[n] (seconds, ascending) and lifetimes[n] (seconds) define the trace; every request asks for profile row `profile`.
// latency_s[n]: result available - arrival.  rec_req / rec_res (capacity `cap` records) and rec_sizes (capacity `cap_calls`)
// receive everything that was submitted, call by call, so that the caller can check the results against an oracle afterwards.
// Returns ISL_OK, the placer's error, or ISL_ERANGE when a record buffer is too small.
int islh_replay_open_loop(islh_place_fn place, void* ctx, uint32_t n, const double* arrivals, const double* lifetimes, uint8_t profile,
                          double* latency_s, isl_request* rec_req, isl_result* rec_res, uint32_t cap, uint32_t* rec_sizes, uint32_t cap_calls,
                          uint32_t* n_calls, uint32_t* n_rec, double* wall_s, uint32_t* placed) {
    struct Exp { double t; uint32_t gpu; uint8_t start, size; };
    struct Later { bool operator()(const Exp& a, const Exp& b) const { return a.t > b.t; } };
    std::priority_queue<Exp, std::vector<Exp>, Later> expiry;
    using clk = std::chrono::steady_clock;
    const clk::time_point t0 = clk::now();
    auto now_s = [&]() { return std::chrono::duration<double>(clk::now() - t0).count(); };
    uint32_t i = 0, calls = 0, rec = 0, n_placed = 0;
    while (i < n) {
        const double now = now_s();
        uint32_t j = i;
        while (j < n && arrivals[j] <= now) ++j;
        uint32_t n_free = 0;
        const uint32_t base = rec;
        while (!expiry.empty() && expiry.top().t <= now) {
            if (rec >= cap) return ISL_ERANGE;
            const Exp x = expiry.top(); expiry.pop();
            rec_req[rec++] = isl_request{x.gpu, 0, (uint8_t)ISL_OP_FREE, x.start, x.size};
            ++n_free;
        }
        if (j == i && n_free == 0) continue;
        if (rec + (j - i) > cap || calls >= cap_calls) return ISL_ERANGE;
        for (uint32_t k = i; k < j; ++k) rec_req[rec++] = isl_request{k, profile, (uint8_t)ISL_OP_ALLOC, 0, 0};
        const uint32_t m = rec - base;
        if (int rc = place(ctx, m, rec_req + base, rec_res + base)) return rc;
        const double done = now_s();
        rec_sizes[calls++] = m;
        for (uint32_t k = i; k < j; ++k) {
            latency_s[k] = done - arrivals[k];
            const isl_result& r = rec_res[base + n_free + (k - i)];
            if (r.status == ISL_ST_PLACED) { ++n_placed; expiry.push(Exp{arrivals[k] + lifetimes[k], r.gpu, r.start, r.size}); }
        }
        i = j;
    }
    *n_calls = calls; *n_rec = rec; *wall_s = now_s(); *placed = n_placed;
    return ISL_OK;
}

}  // extern "C"
