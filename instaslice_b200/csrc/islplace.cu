// islplace.cu — C ABI (include/islplace.h) and host-side orchestration of the placement engine.
//
// The entry points replace, for the allocator path only, what the Go controller does in
// internal/controller/instaslice_controller.go:188-262,303-384 (see the header for the per-symbol map).
// No Go pointer is retained after a call returns; every buffer the engine keeps is its own.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <new>
#include <vector>

#include "isl_kernels.cuh"

using namespace isl;

static constexpr uint32_t kMaxStreamChunks = 4096;

struct isl_engine {
    isl_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::mutex mu;
    char cuda_err[256] = {0};

    DevProfiles prof{};
    bool have_profiles = false, have_inventory = false;
    CandTab tab{};
    uint32_t n_cand_slots = 0;       // K of k_chain<K>
    uint32_t cand_profiles = 0;      // profiles with >= 1 valid (profile, start) candidate

    uint32_t G = 0, lo = 0, hi = 0;
    std::vector<uint32_t> node_off;
    // per-node profile tables (heterogeneous clusters): rows_all[t][p], n_starts == 0 = table t has no row of name p
    uint32_t n_tables = 1;
    isl_profile rows_all[ISL_MAX_TABLES][ISL_MAX_PROFILES] = {};
    std::vector<uint8_t> node_table;     // table of every node (empty = all 0)
    uint8_t* d_gtab = nullptr;           // table of every GPU's node, one byte per GPU
    uint8_t* d_capn = nullptr;           // [table][profile][occ]: placements of the profile the GPU takes in a row
    uint32_t* d_seq = nullptr;           // [table][profile][occ]: their starts, 4 bits each
    uint8_t* d_sizes = nullptr;          // [table][profile]: slices per placement
    uint8_t* d_score = nullptr;          // [profile][occ] of table 0: what a best-fit family policy minimises (k_bestfit)
    unsigned long long* d_cap = nullptr; // isl_capacity: per-profile counters
    uint16_t* d_cand_o16 = nullptr;      // single-chain path: occupancy + table tag of every candidate

    // device buffers
    uint8_t* d_occ = nullptr;        // one byte per GPU, padded to whole sweep blocks with 0xFF
    size_t occ_bytes = 0;
    uint8_t* d_lut = nullptr;        // [16][256]
    uint16_t* d_feas = nullptr;      // [256]
    uint2* d_req = nullptr;          // staging for the host-buffer entry point
    uint2* d_res = nullptr;
    uint16_t* d_q = nullptr;         // per-chunk queues
    uint32_t* d_tile_counts = nullptr;
    uint32_t* d_cand = nullptr;
    uint2* d_log = nullptr;          // decision log of one chunk (chain -> commit)
    uint32_t* d_bf_bitmaps = nullptr; size_t bf_words = 0;   // best-fit class bitmaps for inventories beyond the shared-memory size / several tables
    uint2* h_small_out = nullptr;     // mapped pinned results of tiny batches (k_small writes them over PCIe directly)
    uint2* d_small_out = nullptr;     // device alias of h_small_out
    uint32_t* d_sweep_counts = nullptr;
    Ctrl* d_ctrl = nullptr;
    uint8_t* d_scratch = nullptr;    // eval_starts / free_batch staging
    // stream / segment-pipeline state (grown on demand)
    ChunkDesc* d_chunks = nullptr; Ctrl* d_cctl = nullptr; uint16_t* d_qall = nullptr; uint32_t* d_tokens = nullptr;
    uint32_t* d_free_acc = nullptr; TileDesc* d_tiles = nullptr; std::vector<TileDesc> h_tiles;
    uint32_t cap_chunks = 0, cap_cctl = 0, cap_qall = 0, cap_tokens = 0, cap_free = 0, cap_tiles = 0;
    uint32_t pipe_chunk = 0;         // requests per pipeline chunk (multiple of kTile, <= kChunk)
    std::vector<ChunkDesc> h_chunks;
    uint32_t epoch = 0;
    uint32_t* d_inbox = nullptr;      // [kMaxStreamChunks][kTokStride] tokens written by the previous rank (peer store)
    uint32_t* d_outbox = nullptr;     // next rank's inbox, opened through CUDA IPC
    bool has_prev = false, outbox_local = false;
    unsigned long long* d_trace = nullptr; uint32_t cap_trace = 0, trace_chunks = 0, trace_seg = 0;
    int max_coresident = 0;          // CTAs of k_pipeline that can be resident at once (0 = not queried)
    // host-buffer streams: batches are fed on their own stream while the pipeline runs; results leave chunk by chunk
    cudaStream_t feed_stream = nullptr; cudaEvent_t ev_feed = nullptr, ev_feed_done = nullptr;
    uint32_t* d_ready = nullptr; uint32_t cap_ready = 0;      // [batch] epoch flag
    uint32_t* d_done_cnt = nullptr; uint32_t cap_done = 0;    // [chunk] committed segments
    uint8_t* d_occ_snap = nullptr; size_t snap_bytes = 0; uint32_t snap_G = 0;      // isl_snapshot_occupancy / isl_restore_occupancy
    bool delivered = false;          // the last run_stream call already put the results into the caller's host buffer
    size_t scratch_bytes = 0;
    unsigned long long wait_ns = 20000000000ull;   // a starved device-side wait traps after this long (ISL_WAIT_SECONDS overrides the 20 s)
    uint32_t window = 0;             // causal window of stream calls (isl_set_causal_window): chunk c starts after chunk c - window is committed
    uint32_t spec_mode = ISL_SPEC_AUTO;     // speculative rounds (isl_set_speculation); ISL_SPEC=0|1 in the environment overrides
    unsigned long long* d_specdbg = nullptr;
    // partitioned inventory: every rank's record memory, peer-mapped (own entry = d_spec); bounds of all ranks; the shared memory never moves
    unsigned long long* spec_peer[8] = {}; bool spec_peer_local = false, spec_shared = false;
    uint32_t spec_world = 0, spec_rank = 0, spec_bounds[9] = {};
    unsigned long long* d_spec = nullptr; uint32_t cap_spec = 0, spec_hi = 0;    // record memory of the rounds: kSpecWordsPerChunk words per chunk
    // open stream (isl_stream_open / _submit / _wait / _close): one persistent k_pipeline, batches arrive while it runs
    struct Open {
        bool active = false, launched = false;
        uint32_t max_batches = 0, submitted = 0, epoch = 0, seg = 0, n_seg = 0, sub = 0, q_stride = 0, free_stride = 0, tiles_per_batch = 0;
        bool spec = false;
        uint32_t* h_done = nullptr; uint32_t* d_done_host = nullptr; uint32_t cap_done = 0;   // mapped pinned: [batch] = epoch once its results are in host memory
        ChunkDesc* h_chunks = nullptr; TileDesc* h_tiles = nullptr; uint32_t cap_desc = 0;       // pinned staging of the per-batch descriptors
    } open;
    // partitioned inventory with the results gathered on the owner rank (rank 0): peer-mapped d_res of the owner
    uint2* d_owner_out = nullptr; bool owner_local = false;
    uint32_t ring_world = 0;         // ranks of the partitioned run (isl_set_ring_world); the causal window of a ring needs it

    // stats
    isl_stats st{};
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

inline bool bestfit_family(uint32_t policy) { return policy == ISL_POLICY_BEST_FIT || policy == ISL_POLICY_MIN_FRAG; }
inline bool reversed(const isl_engine* e) { return e->cfg.policy == ISL_POLICY_RIGHT_TO_LEFT; }

#define ISL_CUDA(e, call)                                                                    \
    do {                                                                                     \
        cudaError_t _err = (call);                                                           \
        if (_err != cudaSuccess) {                                                           \
            snprintf((e)->cuda_err, sizeof((e)->cuda_err), "%s: %s", #call, cudaGetErrorString(_err)); \
            return ISL_ECUDA;                                                                \
        }                                                                                    \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
constexpr uint32_t kFeedReserve = 16;      // SMs a fed stream keeps free for its pre-pass kernels

int check_launch(isl_engine* e, const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { snprintf(e->cuda_err, sizeof(e->cuda_err), "%s: %s", what, cudaGetErrorString(err)); return ISL_ECUDA; }
    ++e->st.kernel_launches;
    return ISL_OK;
}

int ensure_scratch(isl_engine* e, size_t bytes) {
    if (bytes <= e->scratch_bytes) return ISL_OK;
    if (e->d_scratch) cudaFree(e->d_scratch);
    e->d_scratch = nullptr; e->scratch_bytes = 0;
    ISL_CUDA(e, cudaMalloc(&e->d_scratch, bytes));
    e->scratch_bytes = bytes;
    return ISL_OK;
}

template <int K>
int launch_chain(isl_engine* e, uint2* d_out_chunk, const uint32_t* d_heads_in, uint32_t* d_heads_out) {
    const size_t smem = (size_t)kQCap * sizeof(uint16_t);      // opted in per device by isl_create
    k_chain<K><<<1, kChainThreads, smem, e->stream>>>(e->tab, e->d_ctrl, e->d_q, e->d_cand_o16, e->d_feas, e->d_log, d_heads_in, d_heads_out);
    if (int rc = check_launch(e, "k_chain")) return rc;
    k_commit<<<kChunk / 256, 256, 0, e->stream>>>(e->d_ctrl, e->d_log, e->d_cand, reinterpret_cast<uint32_t*>(e->d_occ), d_out_chunk, e->prof.flip);
    return check_launch(e, "k_commit");
}

// Resolve n requests that already sit in device memory.  Enqueues only; the caller synchronises.
// The latency path: one launch of one CTA resolves a batch of <= 1024 requests (k_small).
bool small_eligible(const isl_engine* e, uint32_t n) {
    return n > 0 && n <= kSmallMax && !bestfit_family(e->cfg.policy) && !(e->cfg.flags & (ISL_FLAG_NO_SMALL | ISL_FLAG_FORCE_PIPELINE)) &&
           e->hi > e->lo && e->hi - e->lo <= (1u << 18);
}

int run_small(isl_engine* e, uint32_t n, const uint2* d_in, const SmallReqs* inl, uint2* d_out) {
    static const SmallReqs zero{};
    const SmallReqs& params = inl ? *inl : zero;
    const bool timing = e->cfg.flags & ISL_FLAG_TIMING;
    if (timing) cudaEventRecord(e->ev[0], e->stream);
    switch (e->n_cand_slots) {
        case 1: k_small<1><<<1, kSmallThreads, 0, e->stream>>>(e->tab, e->prof, n, d_in, params, d_out, e->d_occ, e->d_gtab, e->d_feas, e->G, e->lo, e->hi,
                                                               e->cand_profiles, e->d_cand, e->d_cand_o16, e->d_ctrl); break;
        case 2: k_small<2><<<1, kSmallThreads, 0, e->stream>>>(e->tab, e->prof, n, d_in, params, d_out, e->d_occ, e->d_gtab, e->d_feas, e->G, e->lo, e->hi,
                                                               e->cand_profiles, e->d_cand, e->d_cand_o16, e->d_ctrl); break;
        default: k_small<4><<<1, kSmallThreads, 0, e->stream>>>(e->tab, e->prof, n, d_in, params, d_out, e->d_occ, e->d_gtab, e->d_feas, e->G, e->lo, e->hi,
                                                                e->cand_profiles, e->d_cand, e->d_cand_o16, e->d_ctrl); break;
    }
    if (int rc = check_launch(e, "k_small")) return rc;
    if (timing) {
        cudaEventRecord(e->ev[1], e->stream);
        cudaEventSynchronize(e->ev[1]);
        float t;
        cudaEventElapsedTime(&t, e->ev[0], e->ev[1]); e->st.ms_commit += t; e->st.ms_total += t;
    }
    ++e->st.batches; e->st.requests += n;
    return ISL_OK;
}

// The shortest path: <= kFewMax requests against <= kFewGpus GPUs (k_few).  Requests as kernel parameters, results into mapped pinned memory.
bool few_eligible(const isl_engine* e, uint32_t n) {
    return n <= kFewMax && small_eligible(e, n) && e->hi - (e->lo & ~15u) <= kFewGpus && !getenv("ISL_NO_FEW");
}
int run_few(isl_engine* e, uint32_t n, const SmallReqs& inl, uint2* d_out) {
    const bool timing = e->cfg.flags & ISL_FLAG_TIMING;
    if (timing) cudaEventRecord(e->ev[0], e->stream);
    k_few<<<1, kFewThreads, 0, e->stream>>>(e->prof, n, inl, d_out, e->d_occ, e->d_gtab, e->d_lut, e->d_sizes, e->n_tables, e->G, e->lo, e->hi, e->d_ctrl);
    if (int rc = check_launch(e, "k_few")) return rc;
    if (timing) {
        cudaEventRecord(e->ev[1], e->stream);
        cudaEventSynchronize(e->ev[1]);
        float t;
        cudaEventElapsedTime(&t, e->ev[0], e->ev[1]); e->st.ms_commit += t; e->st.ms_total += t;
    }
    ++e->st.batches; e->st.requests += n;
    return ISL_OK;
}

// ISL_POLICY_BEST_FIT: frees + defaults, then the request-major class-bitmap kernel (one CTA).
int run_bestfit(isl_engine* e, uint32_t n, const uint2* d_in, uint2* d_out) {
    if (n == 0) return ISL_OK;
    const uint32_t Gr = e->hi - e->lo;
    if (Gr == 0 || Gr > kBfMaxGpus) return ISL_ERANGE;
    const uint32_t W0 = (Gr + 31) / 32, W1 = (W0 + 31) / 32, stride = W0 + W1;
    const bool in_smem = e->n_tables == 1 && Gr <= kBfSmemGpus;
    const size_t smem = in_smem ? (size_t)256 * stride * sizeof(uint32_t) : 0;
    if (!in_smem) {         // class bitmaps in global memory: one set of 256 per table, zeroed here (HBM speed) instead of by the lone CTA
        const size_t words = (size_t)256 * e->n_tables * stride;
        if (words > e->bf_words) {
            if (e->d_bf_bitmaps) cudaFree(e->d_bf_bitmaps);
            e->d_bf_bitmaps = nullptr; e->bf_words = 0;
            ISL_CUDA(e, cudaMalloc(&e->d_bf_bitmaps, words * sizeof(uint32_t)));
            e->bf_words = words;
        }
        ISL_CUDA(e, cudaMemsetAsync(e->d_bf_bitmaps, 0, words * sizeof(uint32_t), e->stream));
    }
    const bool timing = e->cfg.flags & ISL_FLAG_TIMING;
    if (timing) cudaEventRecord(e->ev[0], e->stream);
    k_prepare<<<ceil_div(n, kTile), kTileThreads, 0, e->stream>>>(n, d_in, d_out, reinterpret_cast<uint32_t*>(e->d_occ), e->G, e->lo, e->hi,
                                                                  e->prof, e->d_tile_counts, e->d_ctrl, nullptr, nullptr, 0, 0);
    if (int rc = check_launch(e, "k_prepare")) return rc;
    if (timing) cudaEventRecord(e->ev[1], e->stream);
    if (e->n_tables == 1) k_bestfit<false><<<1, kBfThreads, smem, e->stream>>>(n, d_in, d_out, e->d_occ, e->lo, e->hi, e->d_lut, e->prof, e->d_bf_bitmaps, e->d_ctrl, e->d_score, e->d_gtab, e->d_sizes, 1);
    else k_bestfit<true><<<1, kBfThreads, 0, e->stream>>>(n, d_in, d_out, e->d_occ, e->lo, e->hi, e->d_lut, e->prof, e->d_bf_bitmaps, e->d_ctrl, e->d_score, e->d_gtab, e->d_sizes, e->n_tables);
    if (int rc = check_launch(e, "k_bestfit")) return rc;
    if (timing) {
        cudaEventRecord(e->ev[2], e->stream);
        cudaEventSynchronize(e->ev[2]);
        float t;
        cudaEventElapsedTime(&t, e->ev[0], e->ev[1]); e->st.ms_free += t;
        cudaEventElapsedTime(&t, e->ev[1], e->ev[2]); e->st.ms_commit += t;
        cudaEventElapsedTime(&t, e->ev[0], e->ev[2]); e->st.ms_total += t;
    }
    ++e->st.batches; e->st.requests += n;
    return ISL_OK;
}

int run_batch(isl_engine* e, uint32_t n, const uint2* d_in, uint2* d_out, const uint32_t* d_heads_in, uint32_t* d_heads_out) {
    if (n == 0) return ISL_OK;
    if (bestfit_family(e->cfg.policy)) return run_bestfit(e, n, d_in, d_out);
    const bool timing = e->cfg.flags & ISL_FLAG_TIMING;
    const uint32_t tiles = ceil_div(n, kTile);
    if (timing) cudaEventRecord(e->ev[0], e->stream);
    k_prepare<<<tiles, kTileThreads, 0, e->stream>>>(n, d_in, d_out, reinterpret_cast<uint32_t*>(e->d_occ), e->G, e->lo, e->hi,
                                                     e->prof, e->d_tile_counts, e->d_ctrl, nullptr, nullptr, 0, 0);
    if (int rc = check_launch(e, "k_prepare")) return rc;
    if (timing) cudaEventRecord(e->ev[1], e->stream);
    const uint32_t first_block = e->lo / kSweepBlock;
    const uint32_t sweep_blocks = e->hi > e->lo ? ceil_div(e->hi, kSweepBlock) - first_block : 0;
    float ms_part = 0, ms_sweep = 0, ms_commit = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += kChunk) {
        const uint32_t n_chunk = std::min(kChunk, n - c0);
        const uint32_t first_tile = c0 / kTile, n_tiles = ceil_div(n_chunk, kTile);
        if (timing) cudaEventRecord(e->ev[2], e->stream);
        k_partition<<<n_tiles, kTileThreads, 0, e->stream>>>(n_chunk, d_in + c0, e->prof.n, e->d_tile_counts + (size_t)first_tile * ISL_MAX_PROFILES,
                                                             n_tiles, e->cand_profiles, e->d_q, e->d_ctrl, nullptr, 0, 0);
        if (int rc = check_launch(e, "k_partition")) return rc;
        if (timing) cudaEventRecord(e->ev[3], e->stream);
        // the heads token of a chunk: first chunk of a partitioned call chains from the previous rank
        const uint32_t* h_in = d_heads_in ? d_heads_in + (size_t)(c0 / kChunk) * ISL_MAX_PROFILES : nullptr;
        uint32_t* h_out = d_heads_out ? d_heads_out + (size_t)(c0 / kChunk) * ISL_MAX_PROFILES : nullptr;
        if (h_out) {
            if (h_in) ISL_CUDA(e, cudaMemcpyAsync(h_out, h_in, ISL_MAX_PROFILES * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream));
            else ISL_CUDA(e, cudaMemsetAsync(h_out, 0, ISL_MAX_PROFILES * sizeof(uint32_t), e->stream));
        }
        if (sweep_blocks) {     // candidate compaction — or, for a single-profile chunk, the capacity scan that commits directly
            k_sweep_count<<<sweep_blocks, kSweepThreads, 0, e->stream>>>(reinterpret_cast<const uint4*>(e->d_occ), reinterpret_cast<const uint4*>(e->d_gtab), e->d_feas, first_block, e->lo, e->hi,
                                                                         e->d_ctrl, e->d_sweep_counts, e->d_capn);
            if (int rc = check_launch(e, "k_sweep_count")) return rc;
            k_sweep_scatter<<<sweep_blocks, kSweepThreads, 0, e->stream>>>(reinterpret_cast<const uint4*>(e->d_occ), reinterpret_cast<const uint4*>(e->d_gtab), e->d_feas, first_block, e->lo, e->hi,
                                                                           e->d_ctrl, e->d_sweep_counts, e->d_cand, e->d_cand_o16, e->d_capn, e->d_seq, e->d_q, e->d_occ,
                                                                           d_out + c0, h_in, h_out, e->d_sizes, e->prof.flip);
            if (int rc = check_launch(e, "k_sweep_scatter")) return rc;
        }
        if (timing) cudaEventRecord(e->ev[4], e->stream);
        int rc;
        switch (e->n_cand_slots) {
            case 1: rc = launch_chain<1>(e, d_out + c0, h_in, h_out); break;
            case 2: rc = launch_chain<2>(e, d_out + c0, h_in, h_out); break;
            default: rc = launch_chain<4>(e, d_out + c0, h_in, h_out); break;
        }
        if (rc) return rc;
        if (timing) {
            cudaEventRecord(e->ev[5], e->stream);
            cudaEventSynchronize(e->ev[5]);
            float t;
            cudaEventElapsedTime(&t, e->ev[2], e->ev[3]); ms_part += t;
            cudaEventElapsedTime(&t, e->ev[3], e->ev[4]); ms_sweep += t;
            cudaEventElapsedTime(&t, e->ev[4], e->ev[5]); ms_commit += t;
        }
    }
    if (timing) {
        float t;
        cudaEventElapsedTime(&t, e->ev[0], e->ev[1]); e->st.ms_free += t;     // prepare = frees + defaults + histogram
        e->st.ms_partition += ms_part; e->st.ms_sweep += ms_sweep; e->st.ms_commit += ms_commit;
        cudaEventElapsedTime(&t, e->ev[0], e->ev[5]); e->st.ms_total += t;
    }
    ++e->st.batches;
    e->st.requests += n;
    return ISL_OK;
}


template <int K, bool kP15>
int launch_pipeline(isl_engine* e, PipeArgs& args) {
    void* params[] = {&e->tab, &args};
    const cudaError_t err = cudaLaunchCooperativeKernel(args.spec ? (void*)k_pipeline<K, kP15, true> : (void*)k_pipeline<K, kP15, false>, dim3(args.n_seg + (args.copier ? 1u : 0u)), dim3(kPipeThreads), params, kPipeSmem, e->stream);
    if (err == cudaErrorCooperativeLaunchTooLarge || err == cudaErrorLaunchOutOfResources) {   // e.g. the GPU is shared: not all CTAs can be co-resident
        cudaGetLastError();
        return ISL_ESTATE;          // caller falls back to the chunk-by-chunk path
    }
    if (err != cudaSuccess) { snprintf(e->cuda_err, sizeof(e->cuda_err), "cudaLaunchCooperativeKernel: %s", cudaGetErrorString(err)); return ISL_ECUDA; }
    ++e->st.kernel_launches;
    return ISL_OK;
}

int query_coresident(isl_engine* e) {
    if (e->max_coresident) return ISL_OK;
    int per_sm = 0, sms = 0, coop = 0;
    ISL_CUDA(e, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device));
    ISL_CUDA(e, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
    ISL_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pipeline<4, true, true>, kPipeThreads, kPipeSmem));
    e->max_coresident = coop ? std::max(1, per_sm * sms) : -1;
    return ISL_OK;
}

template <typename T>
int grow(isl_engine* e, T** buf, uint32_t* cap, size_t need, size_t elems_per_unit) {
    if (need <= *cap) return ISL_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr; *cap = 0;
    const size_t units = need + need / 4 + 1;
    ISL_CUDA(e, cudaMalloc(buf, units * elems_per_unit * sizeof(T)));
    *cap = (uint32_t)units;
    return ISL_OK;
}

// Segment geometry of the pipeline for a stream of n_chunks chunks of ~avg_chunk requests.  ISL_ERANGE: the inventory does not fit the
// co-resident CTAs (or the tables need too many candidates per segment) — the caller takes the chunk-by-chunk path.
int plan_pipeline(isl_engine* e, uint32_t n_chunks, double avg_chunk, bool want_feed, uint32_t* seg_out, uint32_t* n_seg_out, uint32_t* sub_out, bool spec = false) {
    if (int rc = query_coresident(e)) return rc;
    const uint32_t range = e->hi - e->lo;
    uint32_t total_cand = 0;
    for (uint32_t k = 0; k < 4; ++k) for (uint32_t l = 0; l < 32; ++l) total_cand += e->tab.desc[k][l] >> 31;
    const uint32_t seg_cap = max_segment_for(total_cand);      // queue windows of a segment must fit in shared memory
    // Segments aimed for.  A batch's decisions form ONE sequential chain over the inventory, only different chunks overlap, so a
    // stream of B chunks over S stages takes about (S + B - 1) x (D x t_dec / S + t_fix), D = decisions of a chunk (tools/chain_cost.py:
    // t_dec 35-41 ns, t_fix 1.6-2.3 us per busy cell over the round: ratio ~0.018 / us).  Minimum at S = sqrt((B - 1) x D x t_dec / t_fix); D is estimated by the
    // smaller of the chunk size and ~3.5 placements per GPU.  ISL_PIPE_SEGMENTS overrides (experiments).
    uint32_t target = 148;
    {
        const double d_est = std::min(avg_chunk, 3.5 * (double)e->G);
        const double s_opt = std::sqrt(std::max(1.0, (double)n_chunks - 1.0) * d_est * (0.041 / 2.3));
        target = (uint32_t)std::min(148.0, std::max(1.0, std::floor(s_opt + 0.5)));
    }
    if (spec) target = 148;      // speculative rounds: every stage works in every round — as many (short) segments as there are SMs
    if (const char* v = getenv("ISL_PIPE_SEGMENTS")) target = std::max(1u, (uint32_t)strtoul(v, nullptr, 10));
    target = std::min(target, (uint32_t)std::max(1, e->max_coresident));
    // fed host streams run their per-batch pre-pass kernels WHILE the pipeline is resident: keep kFeedReserve SMs free of pipeline
    // CTAs (one CTA fills an SM's shared memory, and kernels with another shared-memory carve-out cannot join it there)
    if (want_feed && e->max_coresident > 2 * (int)kFeedReserve) target = std::min(target, (uint32_t)e->max_coresident - 1u - kFeedReserve);
    // segment size from the WHOLE inventory: a partitioned rank keeps the global pipeline depth (~target stages over all ranks)
    if (seg_cap < 64 || e->max_coresident <= 0) return ISL_ERANGE;
    // one sweep / chain / commit round covers `sub` GPUs; an inventory beyond target x sub gives every stage several sub-segments
    const uint32_t sub = std::min(seg_cap, std::max(64u, (ceil_div(e->G, target) + 63u) / 64u * 64u));
    // a short stream wants few stages, a large inventory needs many (a stage holds at most kSubMax sub-segments): the inventory wins
    const uint32_t stages_avail = (want_feed && e->max_coresident > 2 * (int)kFeedReserve) ? (uint32_t)e->max_coresident - 1u - kFeedReserve : (uint32_t)e->max_coresident;
    target = std::max(target, std::min(stages_avail, ceil_div(ceil_div(e->G, sub), kSubMax)));
    const uint32_t n_sub = std::max(1u, ceil_div(ceil_div(e->G, sub), target));
    if (n_sub > kSubMax) return ISL_ERANGE;
    const uint32_t seg = sub * n_sub;
    const uint32_t n_seg = std::max(1u, ceil_div(range, seg));
    if (n_seg > (uint32_t)e->max_coresident) return ISL_ERANGE;
    *seg_out = seg; *n_seg_out = n_seg; *sub_out = sub;
    return ISL_OK;
}

constexpr unsigned long long kOpenWaitNs = 600000000000ull;     // open streams may idle between batches: 10 min

// Speculative rounds (isl_kernels.cuh): wanted for this call?
bool want_spec(isl_engine* e, uint32_t n_batches, uint32_t window, bool ring, bool legacy_token) {
    if ((ring && e->spec_world < 2) || legacy_token || kPipeThreads < 208) return false;
    uint32_t mode = e->spec_mode;
    if (const char* v = getenv("ISL_SPEC")) mode = atoi(v) ? ISL_SPEC_ON : ISL_SPEC_OFF;
    if (mode == ISL_SPEC_OFF) return false;
    if (mode == ISL_SPEC_ON) return true;
    return n_batches == 1 || (window >= 1 && window <= 3);
}
// record memory for n_chunks chunks; the words carry the call epoch (24 bits) — cleared when (re)allocated and when those bits wrap
int prepare_spec(isl_engine* e, uint32_t n_chunks, uint32_t epoch, cudaStream_t st) {
    if (e->spec_shared) return n_chunks <= e->cap_spec ? ISL_OK : ISL_ERANGE;      // peers hold a mapping of it: fixed size, tags carry the stream id
    const uint32_t before = e->cap_spec;
    if (int rc = grow(e, &e->d_spec, &e->cap_spec, n_chunks, kSpecWordsPerChunk)) return rc;
    const bool fresh = e->cap_spec != before, wrapped = (epoch >> 24) != e->spec_hi;
    if (fresh || wrapped) ISL_CUDA(e, cudaMemsetAsync(e->d_spec, 0, (size_t)e->cap_spec * kSpecWordsPerChunk * sizeof(unsigned long long), st));
    e->spec_hi = epoch >> 24;
    return ISL_OK;
}

// Resolve a stream of batches (semantics: one batch after the other).  Enqueues only.
// h_in / h_out (isl_place_stream): the caller's host buffers.  With the segment pipeline the batches are copied and pre-passed one
// by one on a second stream while the pipeline already runs (it waits per batch on a device flag), and an extra CTA copies every
// finished chunk's results straight into h_out when that is mapped pinned memory (e->delivered) — H2D and D2H hide behind the kernel.
int run_stream(isl_engine* e, uint32_t n_batches, const uint32_t* sizes, const uint2* d_in, uint2* d_out,
               const uint32_t* d_heads_in, uint32_t* d_heads_out, uint32_t xepoch = 0, const uint2* h_in = nullptr, uint2* h_out = nullptr,
               bool mixed_single_chunk = false) {
    e->delivered = false;
    uint64_t total = 0;
    uint32_t n_chunks = 0;
    for (uint32_t b = 0; b < n_batches; ++b) { total += sizes[b]; n_chunks += ceil_div(sizes[b], e->pipe_chunk); }
    if (total == 0) return ISL_OK;
    if (total > e->cfg.max_batch) return ISL_ERANGE;
    const uint32_t range = e->hi - e->lo;
    const bool ring = xepoch != 0;      // partitioned inventory: tokens cross ranks through peer memory, pipeline mandatory
    if (ring && n_chunks > kMaxStreamChunks) return ISL_ERANGE;
    auto copy_in_whole = [&]() -> int {       // paths that do not feed batch by batch: one H2D copy up front
        if (h_in) ISL_CUDA(e, cudaMemcpyAsync(const_cast<uint2*>(d_in), h_in, (size_t)total * sizeof(uint2), cudaMemcpyHostToDevice, e->stream));
        return ISL_OK;
    };
    if (n_batches == 1 && !d_heads_in && !d_heads_out && !xepoch && small_eligible(e, sizes[0])) {
        if (int rc = copy_in_whole()) return rc;
        return run_small(e, sizes[0], d_in, nullptr, d_out);
    }
    if ((bestfit_family(e->cfg.policy) || reversed(e)) && (d_heads_in || d_heads_out || xepoch)) return ISL_EINVAL;   // best-fit / right-to-left do not partition
    const bool legacy_token = d_heads_in || d_heads_out || bestfit_family(e->cfg.policy);   // isl_place_batch_partitioned: host-carried token, kChunk layout
    bool pipeline = ring || (!legacy_token && !(e->cfg.flags & ISL_FLAG_NO_PIPELINE) && range > 0 && (n_chunks >= 2 || mixed_single_chunk || (e->cfg.flags & ISL_FLAG_FORCE_PIPELINE)));
    // the stream path keeps one free-mask byte per GPU and batch: very long streams over large inventories go batch by batch
    if (pipeline && (uint64_t)n_batches * e->occ_bytes > (256ull << 20)) { if (ring) return ISL_ERANGE; pipeline = false; }
    // feed mode (below): host buffers, more than one batch, no timing / tracing of the phases, no kernel-serialising tool around
    // (ncu, compute-sanitizer, CUDA_LAUNCH_BLOCKING would starve a pipeline that waits for kernels launched after it: they inject
    // through CUDA_INJECTION64_PATH; ISL_NO_FEED=1 switches feeding off by hand)
    const bool want_feed = h_in && h_out && n_batches >= 2 && !(e->cfg.flags & (ISL_FLAG_TIMING | ISL_FLAG_TRACE)) && !ring && !getenv("ISL_NO_FEED") &&
                           !getenv("CUDA_INJECTION64_PATH") && !getenv("CUDA_LAUNCH_BLOCKING") && !getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR");
    uint32_t seg = 0, n_seg = 0, sub = 0;
    bool spec = false;
    if (pipeline) {
        const uint32_t win = (ring && e->ring_world == 0) ? 0u : e->window;
        if (want_spec(e, n_batches, win, ring, legacy_token)) {      // speculative rounds need one sub-segment per stage
            if (ring) {
                // one sequence of stages over all ranks: a power-of-two stage size that every rank boundary is a multiple of (all ranks evaluate
                // the same predicate over the same bounds, so they agree)
                uint32_t sz = 64;
                while (ceil_div(e->G, sz) > 148u) sz *= 2;
                bool ok = sz <= kSegMax && e->spec_world == e->ring_world && n_chunks <= e->cap_spec && e->spec_bounds[e->spec_world] == e->G &&
                          e->spec_bounds[e->spec_rank] == e->lo && e->spec_bounds[e->spec_rank + 1] == e->hi && query_coresident(e) == ISL_OK;
                for (uint32_t r = 0; ok && r <= e->spec_world; ++r) ok = e->spec_bounds[r] % sz == 0 || e->spec_bounds[r] == e->G;
                for (uint32_t r = 0; ok && r < e->spec_world; ++r) ok = e->spec_bounds[r] < e->spec_bounds[r + 1] && ceil_div(e->spec_bounds[r + 1] - e->spec_bounds[r], sz) <= (uint32_t)std::max(1, e->max_coresident);
                uint32_t tc = 0; for (uint32_t k = 0; k < 4; ++k) for (uint32_t l = 0; l < 32; ++l) tc += e->tab.desc[k][l] >> 31;
                ok = ok && sz <= max_segment_for(tc);
                if (ok) { seg = sub = sz; n_seg = ceil_div(range, sz); spec = true; }
            } else {
                const int src = plan_pipeline(e, n_chunks, (double)total / n_chunks, want_feed, &seg, &n_seg, &sub, true);
                if (src == ISL_ECUDA) return src;
                spec = src == ISL_OK && seg == sub && n_seg <= 148 && n_seg >= 2;
            }
        }
        if (!spec) {
            const int prc = plan_pipeline(e, n_chunks, (double)total / n_chunks, want_feed, &seg, &n_seg, &sub);
            if (prc == ISL_ECUDA) return prc;
            if (prc != ISL_OK) { if (ring) return ISL_ERANGE; pipeline = false; }
        }
    }
    if (!pipeline) {       // one batch after the other through the single-chain path
        if (int rc = copy_in_whole()) return rc;
        uint64_t off = 0;
        uint32_t hoff = 0;
        for (uint32_t b = 0; b < n_batches; ++b) {
            if (int rc = run_batch(e, sizes[b], d_in + off, d_out + off, d_heads_in ? d_heads_in + hoff : nullptr, d_heads_out ? d_heads_out + hoff : nullptr)) return rc;
            off += sizes[b]; hoff += ceil_div(sizes[b], kChunk) * ISL_MAX_PROFILES;
        }
        return ISL_OK;
    }
    const bool timing = e->cfg.flags & ISL_FLAG_TIMING;
    // Tables: every batch is cut into pipeline chunks of pipe_chunk requests (the FREEs of a batch belong to its first
    // chunk) and into tiles of kTile requests for the two pre-pass launches.
    const uint32_t pc = e->pipe_chunk;
    e->h_chunks.clear(); e->h_tiles.clear();
    uint32_t off = 0;
    for (uint32_t b = 0; b < n_batches; ++b) {
        const uint32_t batch_first_tile = (uint32_t)e->h_tiles.size();
        for (uint32_t c0 = 0; c0 < sizes[b]; c0 += pc) {
            const uint32_t cn = std::min(pc, sizes[b] - c0), chunk = (uint32_t)e->h_chunks.size();
            const uint32_t chunk_first_tile = (uint32_t)e->h_tiles.size(), chunk_tiles = ceil_div(cn, kTile);
            e->h_chunks.push_back(ChunkDesc{off + c0, cn, b, c0 == 0 ? 1u : 0u});
            for (uint32_t t = 0; t < chunk_tiles; ++t)
                e->h_tiles.push_back(TileDesc{off, sizes[b], b, batch_first_tile, chunk, chunk_first_tile, chunk_tiles, off + c0, cn, 0, 0, 0});
        }
        off += sizes[b];
    }
    n_chunks = (uint32_t)e->h_chunks.size();
    const uint32_t n_tiles_total = (uint32_t)e->h_tiles.size();
    if (ring && n_chunks > kMaxStreamChunks) return ISL_ERANGE;
    const uint32_t q_stride = pc + kQPad * ISL_MAX_PROFILES;
    const uint32_t free_stride = (uint32_t)e->occ_bytes;           // bytes per batch
    if (int rc = grow(e, &e->d_chunks, &e->cap_chunks, n_chunks, 1)) return rc;
    if (int rc = grow(e, &e->d_cctl, &e->cap_cctl, n_chunks, 1)) return rc;
    if (int rc = grow(e, &e->d_qall, &e->cap_qall, (size_t)n_chunks * q_stride, 1)) return rc;
    if (int rc = grow(e, &e->d_tiles, &e->cap_tiles, n_tiles_total, 1)) return rc;
    if (int rc = grow(e, &e->d_free_acc, &e->cap_free, (size_t)n_batches * (free_stride / 4), 1)) return rc;
    {   // token flags carry the call epoch: a (re)allocated buffer must not hold stale flags of an earlier owner
        const uint32_t before = e->cap_tokens;
        if (int rc = grow(e, &e->d_tokens, &e->cap_tokens, (size_t)n_chunks * (n_seg + 1), kTokStride)) return rc;
        if (e->cap_tokens != before) ISL_CUDA(e, cudaMemsetAsync(e->d_tokens, 0, (size_t)e->cap_tokens * kTokStride * sizeof(uint32_t), e->stream));
    }
    if (n_tiles_total > ceil_div(e->cfg.max_batch, kTile) + 4096) return ISL_ERANGE;
    bool feed = want_feed;       // and room for the copier CTA plus the reserve
    uint2* h_out_dev = nullptr;
    if (feed) {
        if (n_seg + 1 + kFeedReserve > (uint32_t)e->max_coresident) feed = false;       // the pre-pass could starve behind a full house of pipeline CTAs
    }
    if (feed) {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, h_out) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) h_out_dev = static_cast<uint2*>(pa.devicePointer);
        cudaGetLastError();
        if (!e->feed_stream) {
            ISL_CUDA(e, cudaStreamCreateWithFlags(&e->feed_stream, cudaStreamNonBlocking));
            ISL_CUDA(e, cudaEventCreateWithFlags(&e->ev_feed, cudaEventDisableTiming));
            ISL_CUDA(e, cudaEventCreateWithFlags(&e->ev_feed_done, cudaEventDisableTiming));
        }
        if (int rc = grow(e, &e->d_ready, &e->cap_ready, n_batches + 1, 1)) return rc;
    }
    // causal window (device-side): chunk c waits for chunk c - window on every segment (of every rank: a ring counts ranks on the owner)
    const uint32_t window = (ring && e->ring_world == 0) ? 0u : e->window;
    const bool need_done = feed || window;
    if (need_done) if (int rc = grow(e, &e->d_done_cnt, &e->cap_done, n_chunks, 1)) return rc;
    const cudaStream_t pre = feed ? e->feed_stream : e->stream;     // the stream the tables and the pre-pass go to
    if (feed) {     // the feed stream starts behind whatever the engine's stream still holds (earlier calls, load_inventory)
        ISL_CUDA(e, cudaEventRecord(e->ev_feed, e->stream));
        ISL_CUDA(e, cudaStreamWaitEvent(e->feed_stream, e->ev_feed, 0));
    } else if (int rc = copy_in_whole()) return rc;
    if (need_done) ISL_CUDA(e, cudaMemsetAsync(e->d_done_cnt, 0, (size_t)n_chunks * sizeof(uint32_t), pre));
    ISL_CUDA(e, cudaMemcpyAsync(e->d_chunks, e->h_chunks.data(), n_chunks * sizeof(ChunkDesc), cudaMemcpyHostToDevice, pre));
    ISL_CUDA(e, cudaMemcpyAsync(e->d_tiles, e->h_tiles.data(), n_tiles_total * sizeof(TileDesc), cudaMemcpyHostToDevice, pre));
    ISL_CUDA(e, cudaMemsetAsync(e->d_free_acc, 0, (size_t)n_batches * free_stride, pre));
    uint32_t epoch = ++e->epoch;
    if ((epoch & 0x7FFFu) == 0) epoch = ++e->epoch;      // the token words carry the low 15 bits as a tag; tag 0 is what a cleared buffer holds
    if ((epoch & 0x7FFFu) == 1 && epoch != 1 && e->d_tokens)    // tag wrap-around: no stale word of 32 768 calls ago may look current
        ISL_CUDA(e, cudaMemsetAsync(e->d_tokens, 0, (size_t)e->cap_tokens * kTokStride * sizeof(uint32_t), e->stream));
    // pre-pass of the tiles [t0, t1): defaults + free masks + histograms, then the stable partition into per-profile queues
    auto prepass = [&](uint32_t t0, uint32_t t1) -> int {
        k_prepare<<<t1 - t0, kTileThreads, 0, pre>>>(0, d_in, d_out, reinterpret_cast<uint32_t*>(e->d_occ), e->G, e->lo, e->hi, e->prof,
                                                     e->d_tile_counts, e->d_ctrl, e->d_tiles, e->d_free_acc, free_stride / 4, t0);
        if (int rc = check_launch(e, "k_prepare")) return rc;
        if (timing) cudaEventRecord(e->ev[1], e->stream);
        k_partition<<<t1 - t0, kTileThreads, 0, pre>>>(0, d_in, e->prof.n, e->d_tile_counts, 0, e->cand_profiles, e->d_qall, e->d_cctl,
                                                       e->d_tiles, q_stride, t0);
        return check_launch(e, "k_partition");
    };
    // batch b of a fed stream: H2D of its requests, its pre-pass, its ready flag — all on the feed stream
    std::vector<uint32_t> batch_tile0(n_batches + 1, n_tiles_total);
    for (uint32_t t = n_tiles_total; t-- > 0;) batch_tile0[e->h_tiles[t].batch] = t;
    for (uint32_t b = n_batches; b-- > 0;) if (batch_tile0[b] == n_tiles_total) batch_tile0[b] = batch_tile0[b + 1];   // empty batch: no tiles
    uint64_t fed_off = 0;
    auto feed_batch = [&](uint32_t b) -> int {
        if (sizes[b]) {
            ISL_CUDA(e, cudaMemcpyAsync(const_cast<uint2*>(d_in) + fed_off, h_in + fed_off, (size_t)sizes[b] * sizeof(uint2), cudaMemcpyHostToDevice, pre));
            if (int rc = prepass(batch_tile0[b], batch_tile0[b + 1])) return rc;
        }
        k_set_flag<<<1, 1, 0, pre>>>(e->d_ready + b, epoch);
        fed_off += sizes[b];
        return check_launch(e, "k_set_flag");
    };
    if (timing) cudaEventRecord(e->ev[0], e->stream);
    if (feed) {
        if (int rc = feed_batch(0)) return rc;
        ISL_CUDA(e, cudaEventRecord(e->ev_feed, pre));                  // tables + first batch are on their way: the pipeline may start
        ISL_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_feed, 0));
    } else if (int rc = prepass(0, n_tiles_total)) return rc;
    if (timing) cudaEventRecord(e->ev[2], e->stream);
    PipeArgs args{};
    args.n_chunks = n_chunks; args.n_seg = n_seg; args.seg = seg; args.sub = sub; args.lo = e->lo; args.hi = e->hi; args.epoch = epoch;
    args.ready = feed ? e->d_ready : nullptr; args.done_cnt = (h_out_dev || window) ? e->d_done_cnt : nullptr; args.host_out = h_out_dev;
    if (ring && window) {       // the owner's counters sit behind its result array; the other ranks reach them through the same peer mapping
        uint2* base = e->has_prev ? e->d_owner_out : e->d_res;
        if (!base) return ISL_ESTATE;
        args.ring_done = reinterpret_cast<uint32_t*>(base + e->cfg.max_batch); args.world = e->ring_world;
        if (!e->has_prev) ISL_CUDA(e, cudaMemsetAsync(args.ring_done, 0, (size_t)n_chunks * sizeof(uint32_t), e->stream));
    }
    args.flip = e->prof.flip;
    args.copier = h_out_dev ? 1u : 0u; args.window = window; args.wait_ns = e->wait_ns; args.owner_out = ring ? e->d_owner_out : nullptr;
    args.chunks = e->d_chunks; args.cctl = e->d_cctl; args.q_all = e->d_qall; args.free_acc = reinterpret_cast<const uint8_t*>(e->d_free_acc);
    args.q_stride = q_stride; args.free_stride = free_stride; args.tokens = e->d_tokens; args.occ = e->d_occ; args.gtab = e->d_gtab; args.out = d_out; args.feas = e->d_feas; args.stats = e->d_ctrl;
    args.heads_in = d_heads_in; args.heads_out = d_heads_out;
    if (e->cfg.flags & ISL_FLAG_TRACE) {
        if (int rc = grow(e, &e->d_trace, &e->cap_trace, (size_t)n_chunks * n_seg, kTraceWords)) return rc;
        ISL_CUDA(e, cudaMemsetAsync(e->d_trace, 0, (size_t)n_chunks * n_seg * kTraceWords * sizeof(unsigned long long), e->stream));
        e->trace_chunks = n_chunks; e->trace_seg = n_seg;
    }
    args.trace = (e->cfg.flags & ISL_FLAG_TRACE) ? e->d_trace : nullptr;
    args.inbox = ring && e->has_prev ? e->d_inbox : nullptr; args.outbox = ring ? e->d_outbox : nullptr; args.xepoch = xepoch;
    if (spec) {
        if (int rc2 = prepare_spec(e, n_chunks, epoch, e->stream)) return rc2;
        args.spec = getenv("ISL_SPEC_NOREUSE") ? 3u : 1u; args.spec_mem = e->d_spec; args.spec_total = n_seg;
        if (ring) {         // no token ring: the records themselves cross the ranks
            args.inbox = nullptr; args.outbox = nullptr;
            args.spec_world = e->spec_world; args.spec_rank = e->spec_rank; args.spec_base = e->lo / seg; args.spec_total = ceil_div(e->G, seg);
            for (uint32_t r = 0; r < e->spec_world; ++r) args.spec_peer[r] = e->spec_peer[r];
        }
        if (const char* v = getenv("ISL_SPEC_DBG")) {       // per-round stamps of one (chunk, stage) cell: tools/spec_trace.py
            unsigned cchunk = 0, cstage = 0;
            if (sscanf(v, "%u,%u", &cchunk, &cstage) == 2) {
                if (!e->d_specdbg) ISL_CUDA(e, cudaMalloc(&e->d_specdbg, kSpecRounds * 8 * sizeof(unsigned long long)));
                ISL_CUDA(e, cudaMemsetAsync(e->d_specdbg, 0, kSpecRounds * 8 * sizeof(unsigned long long), e->stream));
                args.spec_dbg = e->d_specdbg; args.spec_dbg_cell = (cchunk << 16) | cstage;
            }
        }
    }
    int rc;
    const bool p15 = e->prof.n == ISL_MAX_PROFILES;      // profile index 15 in use: the pop test needs the slower, INF-safe form
    switch (e->n_cand_slots) {
        case 1: rc = p15 ? launch_pipeline<1, true>(e, args) : launch_pipeline<1, false>(e, args); break;
        case 2: rc = p15 ? launch_pipeline<2, true>(e, args) : launch_pipeline<2, false>(e, args); break;
        default: rc = p15 ? launch_pipeline<4, true>(e, args) : launch_pipeline<4, false>(e, args); break;
    }
    if (rc == ISL_ESTATE && ring) return ISL_ERANGE;    // a partitioned run cannot leave the pipeline: the token ring lives inside it
    if (rc == ISL_ESTATE) {         // the pre-pass above did not touch the occupancy (frees went to the free masks): redo batch by batch
        e->max_coresident = -1;     // and do not try the pipeline again on this engine
        if (feed) {                 // only batch 0 was fed: bring the whole stream in behind it
            ISL_CUDA(e, cudaEventRecord(e->ev_feed_done, pre));
            ISL_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_feed_done, 0));
            if (int rc2 = copy_in_whole()) return rc2;
        }
        uint64_t boff = 0;
        for (uint32_t b = 0; b < n_batches; ++b) {
            if (int rc2 = run_batch(e, sizes[b], d_in + boff, d_out + boff, nullptr, nullptr)) return rc2;
            boff += sizes[b];
        }
        return ISL_OK;
    }
    if (rc) return rc;
    if (feed) {     // the remaining batches, while the pipeline works on the first ones
        for (uint32_t b = 1; b < n_batches; ++b)
            if (int rc2 = feed_batch(b)) {
                // the resident pipeline would spin on ready[b] until its trap: publish 'closed' for every batch not fed so that it drains
                std::vector<uint32_t> closed(n_batches - b, ~epoch);
                cudaMemcpy(e->d_ready + b, closed.data(), closed.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
                cudaStreamSynchronize(e->stream);
                return rc2;
            }
        ISL_CUDA(e, cudaEventRecord(e->ev_feed_done, pre));
        ISL_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_feed_done, 0));
        e->delivered = h_out_dev != nullptr;
    }
    if (timing) {
        cudaEventRecord(e->ev[3], e->stream);
        cudaEventSynchronize(e->ev[3]);
        float t;
        cudaEventElapsedTime(&t, e->ev[0], e->ev[1]); e->st.ms_free += t;
        cudaEventElapsedTime(&t, e->ev[1], e->ev[2]); e->st.ms_partition += t;
        cudaEventElapsedTime(&t, e->ev[2], e->ev[3]); e->st.ms_commit += t;
        cudaEventElapsedTime(&t, e->ev[0], e->ev[3]); e->st.ms_total += t;
    }
    e->st.batches += n_batches;
    e->st.requests += total;
    return ISL_OK;
}

int validate_ready(isl_engine* e, uint32_t n) {
    if (!e->have_profiles || !e->have_inventory) return ISL_ESTATE;
    if (e->open.active) return ISL_ESTATE;         // an open stream owns the engine until isl_stream_close
    if (n > e->cfg.max_batch) return ISL_ERANGE;
    return ISL_OK;
}

constexpr uint32_t kSpecRingChunks = 64;         // chunks of one partitioned stream call that may speculate (17 MB of records per rank)

int spec_shared_alloc(isl_engine* e) {
    if (e->spec_shared) return ISL_OK;
    if (e->d_spec) { cudaFree(e->d_spec); e->d_spec = nullptr; e->cap_spec = 0; }
    ISL_CUDA(e, cudaMalloc(&e->d_spec, (size_t)kSpecRingChunks * kSpecWordsPerChunk * sizeof(unsigned long long)));
    ISL_CUDA(e, cudaMemset(e->d_spec, 0, (size_t)kSpecRingChunks * kSpecWordsPerChunk * sizeof(unsigned long long)));
    e->cap_spec = kSpecRingChunks; e->spec_shared = true;
    return ISL_OK;
}

void spec_disconnect(isl_engine* e) {
    for (uint32_t r = 0; r < 8; ++r) {
        if (e->spec_peer[r] && e->spec_peer[r] != e->d_spec && !e->spec_peer_local) cudaIpcCloseMemHandle(e->spec_peer[r]);
        e->spec_peer[r] = nullptr;
    }
    e->spec_world = 0;
}

}  // namespace

extern "C" {

uint32_t isl_abi_version(void) { return ISL_ABI_VERSION; }

const char* isl_strerror(int code) {
    switch (code) {
        case ISL_OK: return "ok";
        case ISL_EINVAL: return "invalid argument or malformed table";
        case ISL_ENOMEM: return "out of memory";
        case ISL_ECUDA: return "CUDA error (see isl_last_cuda_error)";
        case ISL_ESTATE: return "profiles and inventory must be loaded first";
        case ISL_ERANGE: return "batch or inventory exceeds the engine's capacity";
        default: return "unknown error";
    }
}

const char* isl_last_cuda_error(const isl_engine* e) { return e ? e->cuda_err : "null engine"; }

int isl_create(const isl_config* cfg, isl_engine** out) {
    if (!cfg || !out) return ISL_EINVAL;
    *out = nullptr;
    if (cfg->abi_version != ISL_ABI_VERSION) return ISL_EINVAL;
    if (cfg->max_gpus == 0 || cfg->max_gpus > ISL_MAX_GPUS || cfg->max_batch == 0) return ISL_EINVAL;
    if (cfg->policy > ISL_POLICY_MIN_FRAG) return ISL_EINVAL;
    if (bestfit_family(cfg->policy) && cfg->max_gpus > kBfMaxGpus) return ISL_ERANGE;
    if (cfg->quirks & ~ISL_QUIRKS_REF_EXACT) return ISL_EINVAL;
    isl_engine* e = new (std::nothrow) isl_engine;
    if (!e) return ISL_ENOMEM;
    e->cfg = *cfg;
    {   // pipeline chunk size: ISL_PIPE_CHUNK (requests, rounded to tiles) overrides the default
        uint32_t pc = kChunk;
        if (const char* v = getenv("ISL_PIPE_CHUNK")) pc = (uint32_t)strtoul(v, nullptr, 10);
        pc = std::max(kTile, std::min(kChunk, pc / kTile * kTile));
        e->pipe_chunk = pc;
    }
    if (const char* v = getenv("ISL_WAIT_SECONDS")) { const double sec = atof(v); if (sec > 0) e->wait_ns = (unsigned long long)(sec * 1e9); }
    int dev = cfg->device;
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) { delete e; return ISL_ECUDA; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || dev >= ndev) { delete e; return ISL_ECUDA; }
    e->device = dev;
    DeviceGuard guard(dev);
    auto fail = [&](int rc) { isl_destroy(e); return rc; };
#define ISL_TRY(call) do { if ((call) != cudaSuccess) { return fail(ISL_ECUDA); } } while (0)
    ISL_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->own_stream = true;
    {   // Load every kernel NOW.  With CUDA's lazy module loading the first launch of a kernel loads it, and that load can wait for the
        // device to drain — a fed host stream launches k_prepare / k_partition / k_set_flag for batch b WHILE the pipeline kernel is
        // spinning on ready[b]: a first-ever launch at that moment deadlocks (seen as the 20 s trap of a process whose first call was a
        // stream with an empty first batch).
        cudaFuncAttributes fa;
        const void* kernels[] = {(const void*)k_prepare, (const void*)k_partition, (const void*)k_set_flag, (const void*)k_few, (const void*)k_build_lut, (const void*)k_eval_starts,
                                 (const void*)k_free_spans, (const void*)k_capacity, (const void*)k_sweep_count, (const void*)k_sweep_scatter, (const void*)k_commit, (const void*)k_bestfit<false>, (const void*)k_bestfit<true>,
                                 (const void*)k_chain<1>, (const void*)k_chain<2>, (const void*)k_chain<4>, (const void*)k_small<1>, (const void*)k_small<2>, (const void*)k_small<4>,
                                 (const void*)k_pipeline<1, false, false>, (const void*)k_pipeline<1, true, false>, (const void*)k_pipeline<2, false, false>, (const void*)k_pipeline<2, true, false>,
                                 (const void*)k_pipeline<4, false, false>, (const void*)k_pipeline<4, true, false>,
                                 (const void*)k_pipeline<1, false, true>, (const void*)k_pipeline<1, true, true>, (const void*)k_pipeline<2, false, true>, (const void*)k_pipeline<2, true, true>,
                                 (const void*)k_pipeline<4, false, true>, (const void*)k_pipeline<4, true, true>};
        for (const void* k : kernels) ISL_TRY(cudaFuncGetAttributes(&fa, k));
        // dynamic shared memory opt-in, once per engine on its own device (a process-wide cache keyed by a truncated ordinal would
        // skip devices 8.. and race between threads)
        const void* chains[] = {(const void*)k_chain<1>, (const void*)k_chain<2>, (const void*)k_chain<4>};
        for (const void* k : chains) ISL_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kQCap * sizeof(uint16_t))));
        ISL_TRY(cudaFuncSetAttribute((const void*)k_bestfit<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(256 * (kBfSmemGpus / 32 + kBfSmemGpus / 1024) * sizeof(uint32_t))));
        const void* pipes[] = {(const void*)k_pipeline<1, false, false>, (const void*)k_pipeline<1, true, false>, (const void*)k_pipeline<2, false, false>, (const void*)k_pipeline<2, true, false>,
                               (const void*)k_pipeline<4, false, false>, (const void*)k_pipeline<4, true, false>,
                               (const void*)k_pipeline<1, false, true>, (const void*)k_pipeline<1, true, true>, (const void*)k_pipeline<2, false, true>, (const void*)k_pipeline<2, true, true>,
                               (const void*)k_pipeline<4, false, true>, (const void*)k_pipeline<4, true, true>};
        for (const void* k : pipes) ISL_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPipeSmem));
    }
    e->occ_bytes = ((size_t)cfg->max_gpus + kSweepBlock - 1) / kSweepBlock * kSweepBlock;
    const uint32_t max_tiles = ceil_div(cfg->max_batch, kTile) + 4096;   // + one partial tile per batch of a stream
    ISL_TRY(cudaMalloc(&e->d_occ, e->occ_bytes));
    ISL_TRY(cudaMalloc(&e->d_lut, ISL_MAX_TABLES * ISL_MAX_PROFILES * 256));
    ISL_TRY(cudaMalloc(&e->d_feas, ISL_MAX_TABLES * 256 * sizeof(uint16_t)));
    ISL_TRY(cudaMemset(e->d_feas, 0, ISL_MAX_TABLES * 256 * sizeof(uint16_t)));
    ISL_TRY(cudaMalloc(&e->d_capn, ISL_MAX_TABLES * ISL_MAX_PROFILES * 256));
    ISL_TRY(cudaMalloc(&e->d_seq, ISL_MAX_TABLES * ISL_MAX_PROFILES * 256 * sizeof(uint32_t)));
    ISL_TRY(cudaMalloc(&e->d_sizes, ISL_MAX_TABLES * ISL_MAX_PROFILES));
    ISL_TRY(cudaMalloc(&e->d_score, ISL_MAX_TABLES * ISL_MAX_PROFILES * 256));
    ISL_TRY(cudaMalloc(&e->d_cap, ISL_MAX_PROFILES * sizeof(unsigned long long)));
    ISL_TRY(cudaMalloc(&e->d_gtab, e->occ_bytes));
    ISL_TRY(cudaMemset(e->d_gtab, 0, e->occ_bytes));
    ISL_TRY(cudaMalloc(&e->d_cand_o16, e->occ_bytes * sizeof(uint16_t)));
    ISL_TRY(cudaMalloc(&e->d_req, (size_t)cfg->max_batch * sizeof(uint2)));
    ISL_TRY(cudaMalloc(&e->d_res, (size_t)cfg->max_batch * sizeof(uint2) + kMaxStreamChunks * sizeof(uint32_t)));   // + per-chunk 'ranks done' counters of a partitioned run (peer-mapped with the results)
    ISL_TRY(cudaMalloc(&e->d_q, (size_t)kQCap * sizeof(uint16_t)));
    ISL_TRY(cudaMalloc(&e->d_tile_counts, (size_t)max_tiles * ISL_MAX_PROFILES * sizeof(uint32_t)));
    ISL_TRY(cudaMalloc(&e->d_cand, e->occ_bytes * sizeof(uint32_t)));
    ISL_TRY(cudaMalloc(&e->d_log, (size_t)kChunk * sizeof(uint2)));
    ISL_TRY(cudaMalloc(&e->d_sweep_counts, (e->occ_bytes / kSweepBlock) * sizeof(uint32_t)));
    ISL_TRY(cudaMalloc(&e->d_ctrl, sizeof(Ctrl)));
    ISL_TRY(cudaMemset(e->d_ctrl, 0, sizeof(Ctrl)));
    ISL_TRY(cudaMemset(e->d_occ, 0xFF, e->occ_bytes));
    for (auto& ev : e->ev) ISL_TRY(cudaEventCreate(&ev));
    ISL_TRY(cudaHostAlloc(&e->h_small_out, kSmallInline * sizeof(uint2), cudaHostAllocMapped));
    ISL_TRY(cudaHostGetDevicePointer(&e->d_small_out, e->h_small_out, 0));
#undef ISL_TRY
    *out = e;
    return ISL_OK;
}

int isl_destroy(isl_engine* e) {
    if (!e) return ISL_EINVAL;
    if (e->open.active) isl_stream_close(e);        // a persistent pipeline would never let the stream synchronise
    {
        DeviceGuard guard(e->device);
        if (e->stream) cudaStreamSynchronize(e->stream);
        cudaFree(e->d_score); cudaFree(e->d_cap); cudaFree(e->d_gtab); cudaFree(e->d_cand_o16); cudaFree(e->d_capn); cudaFree(e->d_seq); cudaFree(e->d_sizes);
        cudaFree(e->d_occ); cudaFree(e->d_lut); cudaFree(e->d_feas); cudaFree(e->d_req); cudaFree(e->d_res);
        cudaFree(e->d_q); cudaFree(e->d_tile_counts); cudaFree(e->d_cand); cudaFree(e->d_log); cudaFree(e->d_sweep_counts);
        cudaFree(e->d_ctrl); cudaFree(e->d_scratch);
        cudaFree(e->d_chunks); cudaFree(e->d_cctl); cudaFree(e->d_qall); cudaFree(e->d_tokens); spec_disconnect(e); cudaFree(e->d_spec); cudaFree(e->d_specdbg);
        cudaFree(e->d_free_acc); cudaFree(e->d_tiles);
        if (e->d_outbox && !e->outbox_local) cudaIpcCloseMemHandle(e->d_outbox);
        cudaFree(e->d_inbox); cudaFree(e->d_trace); cudaFree(e->d_bf_bitmaps); cudaFree(e->d_ready); cudaFree(e->d_done_cnt); cudaFree(e->d_occ_snap);
        if (e->feed_stream) cudaStreamDestroy(e->feed_stream);
        if (e->ev_feed) cudaEventDestroy(e->ev_feed);
        if (e->ev_feed_done) cudaEventDestroy(e->ev_feed_done);
        if (e->h_small_out) cudaFreeHost(e->h_small_out);
        if (e->open.h_done) cudaFreeHost(e->open.h_done);
        if (e->open.h_chunks) cudaFreeHost(e->open.h_chunks);
        if (e->open.h_tiles) cudaFreeHost(e->open.h_tiles);
        if (e->d_owner_out && !e->owner_local) cudaIpcCloseMemHandle(e->d_owner_out);
        for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
        if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    }
    delete e;
    return ISL_OK;
}

int isl_set_stream(isl_engine* e, void* cuda_stream) {
    if (!e) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    if (e->stream) ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    if (e->own_stream && e->stream) { cudaStreamDestroy(e->stream); e->own_stream = false; }
    if (cuda_stream) e->stream = static_cast<cudaStream_t>(cuda_stream);
    else { ISL_CUDA(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)); e->own_stream = true; }
    return ISL_OK;
}

int isl_synchronize(isl_engine* e) {
    if (!e) return ISL_EINVAL;
    DeviceGuard guard(e->device);
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    return ISL_OK;
}

// shared by isl_load_profiles (one table, every row must have placements) and isl_load_profile_tables
static int load_tables(isl_engine* e, uint32_t n_tables, uint32_t n, const isl_profile* rows, bool allow_absent) {
    if (!e || !rows || n == 0 || n > ISL_MAX_PROFILES || n_tables == 0 || n_tables > ISL_MAX_TABLES) return ISL_EINVAL;
    // Validate what would make the reference panic: empty Placements (:334), start >= 8 (:345).
    uint32_t total_cand = 0;
    for (uint32_t r = 0; r < n_tables * n; ++r) {
        if (rows[r].n_starts == 0 && !allow_absent) return ISL_EINVAL;
        if (rows[r].n_starts > ISL_MAX_STARTS) return ISL_EINVAL;
        for (uint32_t k = 0; k < rows[r].n_starts; ++k) {
            if (rows[r].starts[k] >= ISL_SLOTS) return ISL_EINVAL;
            for (uint32_t j = 0; j < k; ++j) if (rows[r].starts[j] == rows[r].starts[k]) return ISL_EINVAL;   // shim de-duplicates
            total_cand += candidate_mask(rows[r].size, rows[r].starts[k], e->cfg.quirks) != 0;
        }
    }
    if (total_cand > kMaxCand) return ISL_EINVAL;          // more (table, profile, start) candidates than the chain's 4 x 32 lane slots
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    e->n_tables = n_tables;
    memset(e->rows_all, 0, sizeof(e->rows_all));
    for (uint32_t t = 0; t < n_tables; ++t) memcpy(e->rows_all[t], rows + (size_t)t * n, n * sizeof(isl_profile));
    e->prof.n = n; e->prof.quirks = e->cfg.quirks; e->prof.flip = reversed(e) ? e->G : 0u;
    memset(e->prof.rows, 0, sizeof(e->prof.rows));
    // default row of a name (its size is what an unplaced result reports) = the row of the first NODE in canonical order
    // that knows the name; until isl_set_node_tables every node uses table 0
    for (uint32_t p = 0; p < n; ++p) e->prof.rows[p] = e->rows_all[0][p];
    // chain candidates: every (table, profile, start) the search can ever return, in row order
    memset(&e->tab, 0, sizeof(e->tab));
    uint32_t c = 0; e->cand_profiles = 0;
    for (uint32_t t = 0; t < n_tables; ++t)
        for (uint32_t p = 0; p < n; ++p) {
            const isl_profile& row = e->rows_all[t][p];
            uint32_t ord = 0;
            for (uint32_t k = 0; k < row.n_starts; ++k) {
                const uint32_t m = candidate_mask(row.size, row.starts[k], e->cfg.quirks);
                if (!m) continue;
                e->tab.desc[c / 32][c % 32] = p | (ord << 4) | ((uint32_t)row.starts[k] << 7) | ((uint32_t)row.size << 11) | (m << 16) | (t << 24) | (1u << 31);
                ++c; ++ord;
                e->cand_profiles |= 1u << p;
            }
        }
    e->n_cand_slots = c <= 32 ? 1 : (c <= 64 ? 2 : 4);
    ISL_CUDA(e, cudaMemsetAsync(e->d_feas, 0, ISL_MAX_TABLES * 256 * sizeof(uint16_t), e->stream));
    for (uint32_t t = 0; t < n_tables; ++t) {
        DevProfiles dp{};
        dp.n = n; dp.quirks = e->cfg.quirks;
        memcpy(dp.rows, e->rows_all[t], sizeof(dp.rows));
        k_build_lut<<<1, 256, 0, e->stream>>>(dp, t, e->d_lut, e->d_feas, e->d_capn, e->d_seq);
        if (int rc = check_launch(e, "k_build_lut")) return rc;
    }
    {
        uint8_t sizes[ISL_MAX_TABLES * ISL_MAX_PROFILES] = {0};
        for (uint32_t t = 0; t < n_tables; ++t) for (uint32_t p = 0; p < n; ++p) sizes[t * ISL_MAX_PROFILES + p] = e->rows_all[t][p].size;
        ISL_CUDA(e, cudaMemcpyAsync(e->d_sizes, sizes, sizeof(sizes), cudaMemcpyHostToDevice, e->stream));
    }
    {   // what a best-fit family policy minimises, per table
        std::vector<uint8_t> score((size_t)ISL_MAX_TABLES * ISL_MAX_PROFILES * 256);      // the copy below completes before this function returns (stream sync)
        for (uint32_t t = 0; t < n_tables; ++t) {
            std::vector<uint32_t> cand;                              // every (profile, start) mask of the table the search can return
            for (uint32_t p = 0; p < n; ++p)
                for (uint32_t k = 0; k < e->rows_all[t][p].n_starts; ++k)
                    if (const uint32_t m = candidate_mask(e->rows_all[t][p].size, e->rows_all[t][p].starts[k], e->cfg.quirks)) cand.push_back(m);
            for (uint32_t p = 0; p < ISL_MAX_PROFILES; ++p)
                for (uint32_t o = 0; o < 256; ++o) {
                    uint32_t mine = 0;                                   // the mask the profile would take there: first legal start in row order
                    if (p < n)
                        for (uint32_t k = 0; k < e->rows_all[t][p].n_starts && !mine; ++k) {
                            const uint32_t m = candidate_mask(e->rows_all[t][p].size, e->rows_all[t][p].starts[k], e->cfg.quirks);
                            if (m && (o & m) == 0) mine = m;
                        }
                    // ISL_POLICY_BEST_FIT: free slices of the GPU AFTER the placement, fewest first (with several tables one name may
                    // span different sizes, so "after" is not "before minus a constant")
                    uint32_t v = 8u - (uint32_t)__builtin_popcount(o | mine);
                    if (e->cfg.policy == ISL_POLICY_MIN_FRAG) {
                        v = 0;
                        if (mine) for (uint32_t m : cand) v += ((o & m) == 0) && (((o | mine) & m) != 0);      // pairs that stop being feasible
                    }
                    score[((size_t)t * ISL_MAX_PROFILES + p) * 256 + o] = (uint8_t)v;
                }
        }
        ISL_CUDA(e, cudaMemcpyAsync(e->d_score, score.data(), score.size(), cudaMemcpyHostToDevice, e->stream));
    }
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    e->have_profiles = true;
    return ISL_OK;
}

int isl_load_profiles(isl_engine* e, uint32_t n, const isl_profile* rows) { return load_tables(e, 1, n, rows, false); }

int isl_load_profile_tables(isl_engine* e, uint32_t n_tables, uint32_t n_profiles, const isl_profile* rows) {
    return load_tables(e, n_tables, n_profiles, rows, true);
}

int isl_set_node_tables(isl_engine* e, uint32_t n_nodes, const uint8_t* table_of_node) {
    if (!e || !table_of_node) return ISL_EINVAL;
    if (!e->have_inventory || !e->have_profiles) return ISL_ESTATE;
    if (n_nodes + 1 != e->node_off.size()) return ISL_EINVAL;
    for (uint32_t n = 0; n < n_nodes; ++n) if (table_of_node[n] >= e->n_tables) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    e->node_table.assign(table_of_node, table_of_node + n_nodes);
    std::vector<uint8_t> gtab(e->G);
    for (uint32_t n = 0; n < n_nodes; ++n)
        for (uint32_t g = e->node_off[n]; g < e->node_off[n + 1]; ++g) gtab[flip_gpu(g, e->prof.flip)] = table_of_node[n];
    ISL_CUDA(e, cudaMemcpyAsync(e->d_gtab, gtab.data(), e->G, cudaMemcpyHostToDevice, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    for (uint32_t p = 0; p < e->prof.n; ++p) {              // size reported for an unplaced request: first node (canonical order) that knows the name
        e->prof.rows[p] = isl_profile{};
        for (uint32_t n = 0; n < n_nodes; ++n)
            if (e->rows_all[table_of_node[n]][p].n_starts) { e->prof.rows[p] = e->rows_all[table_of_node[n]][p]; break; }
    }
    return ISL_OK;
}

int isl_load_inventory(isl_engine* e, uint32_t n_nodes, const uint32_t* node_off, const uint8_t* occ) {
    if (!e || !node_off || n_nodes == 0) return ISL_EINVAL;
    if (node_off[0] != 0) return ISL_EINVAL;
    for (uint32_t i = 0; i < n_nodes; ++i) if (node_off[i + 1] < node_off[i]) return ISL_EINVAL;
    const uint32_t G = node_off[n_nodes];
    if (G == 0 || !occ) return ISL_EINVAL;
    if (G > e->cfg.max_gpus) return ISL_ERANGE;
    if (e->open.active) return ISL_ESTATE;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    e->node_off.assign(node_off, node_off + n_nodes + 1);
    e->G = G; e->lo = 0; e->hi = G;
    e->prof.flip = reversed(e) ? G : 0u;      // ISL_POLICY_RIGHT_TO_LEFT: the inventory is stored in reverse canonical order
    e->snap_G = 0;                  // a snapshot belongs to the inventory it was taken from
    ISL_CUDA(e, cudaMemsetAsync(e->d_occ, 0xFF, e->occ_bytes, e->stream));
    ISL_CUDA(e, cudaMemsetAsync(e->d_gtab, 0, e->occ_bytes, e->stream));        // every node uses table 0 until isl_set_node_tables
    e->node_table.clear();
    for (uint32_t p = 0; p < e->prof.n; ++p) e->prof.rows[p] = e->rows_all[0][p];
    std::vector<uint8_t> rev;
    if (e->prof.flip) { rev.assign(occ, occ + G); std::reverse(rev.begin(), rev.end()); occ = rev.data(); }
    ISL_CUDA(e, cudaMemcpyAsync(e->d_occ, occ, G, cudaMemcpyHostToDevice, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    e->have_inventory = true;
    return ISL_OK;
}

int isl_read_occupancy(isl_engine* e, uint8_t* out) {
    if (!e || !out) return ISL_EINVAL;
    if (!e->have_inventory || e->open.active) return ISL_ESTATE;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    ISL_CUDA(e, cudaMemcpyAsync(out, e->d_occ, e->G, cudaMemcpyDeviceToHost, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    if (e->prof.flip) std::reverse(out, out + e->G);       // canonical order at the boundary
    return ISL_OK;
}

int isl_write_occupancy(isl_engine* e, uint32_t first_gpu, uint32_t n, const uint8_t* occ) {
    if (!e || (n && !occ)) return ISL_EINVAL;
    if (!e->have_inventory || e->open.active) return ISL_ESTATE;
    if ((uint64_t)first_gpu + n > e->G) return ISL_ERANGE;
    if (n == 0) return ISL_OK;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    std::vector<uint8_t> rev;
    if (e->prof.flip) { rev.assign(occ, occ + n); std::reverse(rev.begin(), rev.end()); occ = rev.data(); first_gpu = e->G - first_gpu - n; }
    ISL_CUDA(e, cudaMemcpyAsync(e->d_occ + first_gpu, occ, n, cudaMemcpyHostToDevice, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    return ISL_OK;
}

int isl_snapshot_occupancy(isl_engine* e) {
    if (!e) return ISL_EINVAL;
    if (!e->have_inventory) return ISL_ESTATE;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    if (e->snap_bytes < e->occ_bytes) {
        if (e->d_occ_snap) cudaFree(e->d_occ_snap);
        e->d_occ_snap = nullptr; e->snap_bytes = 0;
        ISL_CUDA(e, cudaMalloc(&e->d_occ_snap, e->occ_bytes));
        e->snap_bytes = e->occ_bytes;
    }
    ISL_CUDA(e, cudaMemcpyAsync(e->d_occ_snap, e->d_occ, e->occ_bytes, cudaMemcpyDeviceToDevice, e->stream));
    e->snap_G = e->G;
    return ISL_OK;
}

int isl_restore_occupancy(isl_engine* e) {
    if (!e) return ISL_EINVAL;
    if (!e->have_inventory || !e->d_occ_snap || e->snap_G != e->G) return ISL_ESTATE;     // a snapshot belongs to the inventory it was taken from
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    ISL_CUDA(e, cudaMemcpyAsync(e->d_occ, e->d_occ_snap, e->occ_bytes, cudaMemcpyDeviceToDevice, e->stream));
    return ISL_OK;
}

uint32_t isl_num_gpus(const isl_engine* e) { return e ? e->G : 0; }

uint32_t isl_gpu_to_node(const isl_engine* e, uint32_t gpu) {
    if (!e || gpu >= e->G) return ISL_GPU_NONE;
    auto it = std::upper_bound(e->node_off.begin(), e->node_off.end(), gpu);
    return (uint32_t)(it - e->node_off.begin()) - 1;
}

static int place_batch_locked(isl_engine* e, uint32_t n, const isl_request* in, isl_result* out);

int isl_place_batch(isl_engine* e, uint32_t n, const isl_request* in, isl_result* out) {
    if (!e || (n && (!in || !out))) return ISL_EINVAL;
    if (int rc = validate_ready(e, n)) return rc;
    if (n == 0) return ISL_OK;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    return place_batch_locked(e, n, in, out);
}

int isl_place_batch_range(isl_engine* e, uint32_t lo, uint32_t hi, uint32_t n, const isl_request* in, isl_result* out) {
    if (!e || (n && (!in || !out))) return ISL_EINVAL;
    if (int rc = validate_ready(e, n)) return rc;
    if (lo > hi || hi > e->G) return ISL_EINVAL;
    if (n == 0) return ISL_OK;
    std::lock_guard<std::mutex> lk(e->mu);          // restriction, placement and restore under ONE lock: two callers cannot interleave
    DeviceGuard guard(e->device);
    const uint32_t lo0 = e->lo, hi0 = e->hi;
    if (e->prof.flip) { e->lo = e->G - hi; e->hi = e->G - lo; } else { e->lo = lo; e->hi = hi; }
    const int rc = place_batch_locked(e, n, in, out);
    e->lo = lo0; e->hi = hi0;
    return rc;
}

static int place_batch_plain(isl_engine* e, uint32_t n, const isl_request* in, isl_result* out);

// ISL_FLAG_ALL_NODES — the reference's literal multi-node behaviour: Reconcile's node loop (:190-227) has no `break` after a
// successful node, so a pod is allocated on EVERY node that has capacity.  Nodes do not interact (an allocation on one node never
// changes what another node can take), so "every pod over all nodes" equals "every node over all pods": one restricted pass per node,
// each consuming capacity on its node; the record reported for a pod is the first node's (what the oracle's all_nodes switch reports).
// A compatibility mode for parity studies, one engine call per node — not a fast path.
static int place_batch_locked(isl_engine* e, uint32_t n, const isl_request* in, isl_result* out) {
    if (!(e->cfg.flags & ISL_FLAG_ALL_NODES)) return place_batch_plain(e, n, in, out);
    const uint32_t lo0 = e->lo, hi0 = e->hi, n_nodes = (uint32_t)e->node_off.size() - 1;
    const uint32_t clo = e->prof.flip ? e->G - hi0 : lo0, chi = e->prof.flip ? e->G - lo0 : hi0;     // the caller's range, canonical
    std::vector<isl_result> tmp(n);
    bool first = true;
    int rc = ISL_OK;
    for (uint32_t k = 0; k < n_nodes && !rc; ++k) {
        const uint32_t node = e->prof.flip ? n_nodes - 1 - k : k;                  // nodes in policy order
        const uint32_t a = std::max(clo, e->node_off[node]), b = std::min(chi, e->node_off[node + 1]);
        if (a >= b) continue;
        if (e->prof.flip) { e->lo = e->G - b; e->hi = e->G - a; } else { e->lo = a; e->hi = b; }
        rc = place_batch_plain(e, n, in, first ? out : tmp.data());
        if (!rc && !first)
            for (uint32_t i = 0; i < n; ++i)
                if (in[i].op == ISL_OP_ALLOC && out[i].status != ISL_ST_PLACED && tmp[i].status == ISL_ST_PLACED) out[i] = tmp[i];
        first = false;
    }
    e->lo = lo0; e->hi = hi0;
    if (!rc && first) rc = place_batch_plain(e, n, in, out);       // empty range: defaults only
    return rc;
}

static int place_batch_plain(isl_engine* e, uint32_t n, const isl_request* in, isl_result* out) {
    if (n <= kSmallInline && small_eligible(e, n)) {        // requests as kernel parameters, results into mapped pinned memory: 1 launch + 1 sync
        SmallReqs inl{};
        memcpy(inl.r, in, (size_t)n * sizeof(isl_request));
        if (int rc = few_eligible(e, n) ? run_few(e, n, inl, e->d_small_out) : run_small(e, n, nullptr, &inl, e->d_small_out)) return rc;
        ISL_CUDA(e, cudaStreamSynchronize(e->stream));
        memcpy(out, e->h_small_out, (size_t)n * sizeof(isl_result));
        return ISL_OK;
    }
    ISL_CUDA(e, cudaMemcpyAsync(e->d_req, in, (size_t)n * sizeof(isl_request), cudaMemcpyHostToDevice, e->stream));
    // One large batch that mixes profiles: the segment pipeline's decision loop (35 ns per decision, one launch for the chain of all
    // segments) beats the single-chain path (~110 ns) although nothing overlaps inside one chunk.  A batch with a single placeable
    // profile stays on the single-chain path, whose scan mode commits it without any chain.  (The buffer is on the host: a look at
    // the profile bytes costs microseconds.)
    bool mixed = false;
    if (n >= 4096) {
        uint32_t seen = 0;
        for (uint32_t i = 0; i < n && !mixed; ++i)
            if (in[i].op == ISL_OP_ALLOC && in[i].profile < ISL_MAX_PROFILES) { seen |= (1u << in[i].profile) & e->cand_profiles; mixed = (seen & (seen - 1)) != 0; }
    }
    if (int rc = run_stream(e, 1, &n, e->d_req, e->d_res, nullptr, nullptr, 0, nullptr, nullptr, mixed)) return rc;
    ISL_CUDA(e, cudaMemcpyAsync(out, e->d_res, (size_t)n * sizeof(isl_result), cudaMemcpyDeviceToHost, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    return ISL_OK;
}

int isl_place_stream(isl_engine* e, uint32_t n_batches, const uint32_t* sizes, const isl_request* in, isl_result* out) {
    if (!e || !sizes || n_batches == 0 || n_batches > 4096) return ISL_EINVAL;
    uint64_t total = 0;
    for (uint32_t b = 0; b < n_batches; ++b) total += sizes[b];
    if (total && (!in || !out)) return ISL_EINVAL;
    if (total > e->cfg.max_batch) return ISL_ERANGE;
    if (int rc = validate_ready(e, (uint32_t)total)) return rc;
    if (total == 0) return ISL_OK;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    if (int rc = run_stream(e, n_batches, sizes, e->d_req, e->d_res, nullptr, nullptr, 0, reinterpret_cast<const uint2*>(in), reinterpret_cast<uint2*>(out))) {
        if (e->feed_stream) cudaStreamSynchronize(e->feed_stream);
        return rc;
    }
    if (!e->delivered) ISL_CUDA(e, cudaMemcpyAsync(out, e->d_res, (size_t)total * sizeof(isl_result), cudaMemcpyDeviceToHost, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    return ISL_OK;
}

int isl_place_stream_device(isl_engine* e, uint32_t n_batches, const uint32_t* sizes, const void* d_in, void* d_out) {
    if (!e || !sizes || n_batches == 0 || n_batches > 4096 || !d_in || !d_out) return ISL_EINVAL;
    uint64_t total = 0;
    for (uint32_t b = 0; b < n_batches; ++b) total += sizes[b];
    if (total > e->cfg.max_batch) return ISL_ERANGE;
    if (int rc = validate_ready(e, (uint32_t)total)) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    return run_stream(e, n_batches, sizes, static_cast<const uint2*>(d_in), static_cast<uint2*>(d_out), nullptr, nullptr);
}

int isl_ipc_inbox_handle(isl_engine* e, void* handle64) {
    if (!e || !handle64) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    if (!e->d_inbox) {
        ISL_CUDA(e, cudaMalloc(&e->d_inbox, (size_t)kMaxStreamChunks * kTokStride * sizeof(uint32_t)));
        ISL_CUDA(e, cudaMemset(e->d_inbox, 0, (size_t)kMaxStreamChunks * kTokStride * sizeof(uint32_t)));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    ISL_CUDA(e, cudaIpcGetMemHandle(&h, e->d_inbox));
    memcpy(handle64, &h, sizeof h);
    return ISL_OK;
}

int isl_ipc_connect(isl_engine* e, const void* next_handle64, int has_prev) {
    if (!e) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    if (e->d_outbox && !e->outbox_local) cudaIpcCloseMemHandle(e->d_outbox);
    e->d_outbox = nullptr; e->outbox_local = false;
    if (next_handle64) {
        cudaIpcMemHandle_t h;
        memcpy(&h, next_handle64, sizeof h);
        void* p = nullptr;
        ISL_CUDA(e, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        e->d_outbox = static_cast<uint32_t*>(p);
    }
    e->has_prev = has_prev != 0;
    if (e->has_prev && !e->d_inbox) return ISL_ESTATE;
    return ISL_OK;
}

int isl_connect_local(isl_engine* e, isl_engine* next, int has_prev) {
    if (!e) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    if (e->d_outbox && !e->outbox_local) cudaIpcCloseMemHandle(e->d_outbox);
    e->d_outbox = nullptr; e->outbox_local = true;
    if (next) {
        if (!next->d_inbox) return ISL_ESTATE;
        e->d_outbox = next->d_inbox;
    }
    e->has_prev = has_prev != 0;
    if (e->has_prev && !e->d_inbox) return ISL_ESTATE;
    return ISL_OK;
}

// ---- speculative rounds over a partitioned inventory: every rank's record memory mapped into every other rank ------------------------
int isl_ipc_spec_handle(isl_engine* e, void* handle64) {
    if (!e || !handle64) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    if (int rc = spec_shared_alloc(e)) return rc;
    cudaIpcMemHandle_t h;
    ISL_CUDA(e, cudaIpcGetMemHandle(&h, e->d_spec));
    memcpy(handle64, &h, sizeof h);
    return ISL_OK;
}

// handles: world x 64 bytes (isl_ipc_spec_handle of every rank, own entry ignored); bounds: world + 1 canonical GPU indices, rank r owns
// [bounds[r], bounds[r + 1]).  world = 0 disconnects.
int isl_ipc_connect_spec(isl_engine* e, uint32_t world, uint32_t rank, const void* handles, const uint32_t* bounds) {
    if (!e) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    spec_disconnect(e);
    if (world == 0) return ISL_OK;
    if (world < 2 || world > 8 || rank >= world || !handles || !bounds) return ISL_EINVAL;
    if (int rc = spec_shared_alloc(e)) return rc;
    e->spec_peer_local = false;
    for (uint32_t r = 0; r < world; ++r) {
        if (r == rank) { e->spec_peer[r] = e->d_spec; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + 64 * r, sizeof h);
        void* p = nullptr;
        ISL_CUDA(e, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        e->spec_peer[r] = static_cast<unsigned long long*>(p);
    }
    for (uint32_t r = 0; r <= world; ++r) e->spec_bounds[r] = bounds[r];
    e->spec_world = world; e->spec_rank = rank;
    return ISL_OK;
}

// same-process engines (tests: several ranks on one GPU)
int isl_connect_spec_local(isl_engine* e, uint32_t world, uint32_t rank, isl_engine* const* engines, const uint32_t* bounds) {
    if (!e) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    spec_disconnect(e);
    if (world == 0) return ISL_OK;
    if (world < 2 || world > 8 || rank >= world || !engines || !bounds) return ISL_EINVAL;
    if (int rc = spec_shared_alloc(e)) return rc;
    e->spec_peer_local = true;
    for (uint32_t r = 0; r < world; ++r) {
        if (r != rank && (!engines[r] || !engines[r]->spec_shared)) return ISL_ESTATE;     // every engine allocates first (isl_ipc_spec_handle)
        e->spec_peer[r] = r == rank ? e->d_spec : engines[r]->d_spec;
    }
    for (uint32_t r = 0; r <= world; ++r) e->spec_bounds[r] = bounds[r];
    e->spec_world = world; e->spec_rank = rank;
    return ISL_OK;
}

int isl_place_stream_partitioned(isl_engine* e, uint32_t n_batches, const uint32_t* sizes, const void* d_in, void* d_out, uint32_t stream_id) {
    if (!e || !sizes || n_batches == 0 || n_batches > 4096 || !d_in || !d_out || stream_id == 0) return ISL_EINVAL;
    uint64_t total = 0;
    for (uint32_t b = 0; b < n_batches; ++b) total += sizes[b];
    if (total > e->cfg.max_batch) return ISL_ERANGE;
    if (int rc = validate_ready(e, (uint32_t)total)) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    return run_stream(e, n_batches, sizes, static_cast<const uint2*>(d_in), static_cast<uint2*>(d_out), nullptr, nullptr, stream_id);
}

int isl_place_batch_device(isl_engine* e, uint32_t n, const void* d_in, void* d_out) {
    if (!e || (n && (!d_in || !d_out))) return ISL_EINVAL;
    if (int rc = validate_ready(e, n)) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    return run_stream(e, 1, &n, static_cast<const uint2*>(d_in), static_cast<uint2*>(d_out), nullptr, nullptr);
}

int isl_place_batch_partitioned(isl_engine* e, uint32_t n, const void* d_in, void* d_out, const void* d_heads_in, void* d_heads_out) {
    if (!e || (n && (!d_in || !d_out)) || !d_heads_out) return ISL_EINVAL;
    if (int rc = validate_ready(e, n)) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    return run_stream(e, 1, &n, static_cast<const uint2*>(d_in), static_cast<uint2*>(d_out), static_cast<const uint32_t*>(d_heads_in),
                      static_cast<uint32_t*>(d_heads_out));
}

int isl_set_partition(isl_engine* e, uint32_t lo, uint32_t hi) {
    if (!e) return ISL_EINVAL;
    if (!e->have_inventory) return ISL_ESTATE;
    if (lo > hi || hi > e->G) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    if (e->prof.flip) { e->lo = e->G - hi; e->hi = e->G - lo; } else { e->lo = lo; e->hi = hi; }
    return ISL_OK;
}

void* isl_device_occupancy(isl_engine* e) { return e ? e->d_occ : nullptr; }

int isl_free_batch(isl_engine* e, uint32_t n, const isl_span* spans) {
    if (!e || (n && !spans)) return ISL_EINVAL;
    if (!e->have_inventory || e->open.active) return ISL_ESTATE;
    if (n == 0) return ISL_OK;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    if (int rc = ensure_scratch(e, (size_t)n * sizeof(isl_span))) return rc;
    ISL_CUDA(e, cudaMemcpyAsync(e->d_scratch, spans, (size_t)n * sizeof(isl_span), cudaMemcpyHostToDevice, e->stream));
    k_free_spans<<<ceil_div(n, 256), 256, 0, e->stream>>>(n, reinterpret_cast<const isl_span*>(e->d_scratch), reinterpret_cast<uint32_t*>(e->d_occ),
                                                          e->G, e->lo, e->hi, e->d_ctrl, e->prof.flip);
    if (int rc = check_launch(e, "k_free_spans")) return rc;
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    return ISL_OK;
}

int isl_eval_starts(isl_engine* e, uint32_t profile, uint32_t n, const uint8_t* occ, uint8_t* out) {
    if (!e || (n && (!occ || !out))) return ISL_EINVAL;
    if (!e->have_profiles) return ISL_ESTATE;
    const uint32_t table = profile >> 8;
    profile &= 0xFFu;
    if (profile >= e->prof.n || table >= e->n_tables) return ISL_EINVAL;
    if (n == 0) return ISL_OK;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    if (int rc = ensure_scratch(e, (size_t)n * 2)) return rc;
    ISL_CUDA(e, cudaMemcpyAsync(e->d_scratch, occ, n, cudaMemcpyHostToDevice, e->stream));
    k_eval_starts<<<std::min(ceil_div(n, 256), 1184u), 256, 0, e->stream>>>(e->d_lut + (size_t)table * ISL_MAX_PROFILES * 256, profile, n, e->d_scratch, e->d_scratch + n);
    if (int rc = check_launch(e, "k_eval_starts")) return rc;
    ISL_CUDA(e, cudaMemcpyAsync(out, e->d_scratch + n, n, cudaMemcpyDeviceToHost, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    return ISL_OK;
}

int isl_read_trace(isl_engine* e, uint64_t* out, uint32_t max_words, uint32_t* n_chunks, uint32_t* n_seg) {
    if (!e || !n_chunks || !n_seg) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    *n_chunks = e->trace_chunks; *n_seg = e->trace_seg;
    const size_t words = (size_t)e->trace_chunks * e->trace_seg * kTraceWords;
    if (!out || words == 0) return ISL_OK;
    if (words > max_words) return ISL_ERANGE;
    ISL_CUDA(e, cudaMemcpyAsync(out, e->d_trace, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    return ISL_OK;
}

int isl_get_stats(isl_engine* e, isl_stats* out) {
    if (!e || !out) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    Ctrl c;
    ISL_CUDA(e, cudaMemcpyAsync(&c, e->d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    e->st.placed = c.placed; e->st.freed = c.freed; e->st.no_capacity = c.allocs - c.placed; e->st.chain_steps = c.steps; e->st.chain_gpus_visited = c.visited; e->st.chain_jumps = c.jumps; e->st.scan_placed = c.scanned;
    e->st.spec_chunks = c.spec_cells; e->st.spec_rounds = c.spec_rounds; e->st.spec_sims = c.spec_sims;
    *out = e->st;
    return ISL_OK;
}

int isl_reset_stats(isl_engine* e) {
    if (!e) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    const uint64_t launches = e->st.kernel_launches;
    e->st = isl_stats{};
    e->st.kernel_launches = launches;      // launches are counted since creation
    ISL_CUDA(e, cudaMemsetAsync(&e->d_ctrl->placed, 0, 8 * sizeof(unsigned long long), e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    return ISL_OK;
}

// ---- what-if / defragmentation queries -------------------------------------------------------------------
static int capacity_locked(isl_engine* e, uint64_t* cap) {
    ISL_CUDA(e, cudaMemsetAsync(e->d_cap, 0, ISL_MAX_PROFILES * sizeof(unsigned long long), e->stream));
    if (e->hi > e->lo) {
        k_capacity<<<std::min(ceil_div(e->hi - e->lo, 256), 296u), 256, 0, e->stream>>>(e->d_occ, e->d_gtab, e->d_capn, e->prof.n, e->lo, e->hi, e->d_cap);
        if (int rc = check_launch(e, "k_capacity")) return rc;
    }
    unsigned long long h[ISL_MAX_PROFILES];
    ISL_CUDA(e, cudaMemcpyAsync(h, e->d_cap, sizeof h, cudaMemcpyDeviceToHost, e->stream));
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    for (uint32_t p = 0; p < ISL_MAX_PROFILES; ++p) cap[p] = h[p];
    return ISL_OK;
}

int isl_capacity(isl_engine* e, uint64_t* cap) {
    if (!e || !cap) return ISL_EINVAL;
    if (int rc = validate_ready(e, 0)) return rc;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    return capacity_locked(e, cap);
}

int isl_what_if(isl_engine* e, uint32_t n, const isl_request* plan, isl_result* out, uint64_t* cap_before, uint64_t* cap_after) {
    if (!e || (n && (!plan || !out))) return ISL_EINVAL;
    if (int rc = validate_ready(e, n)) return rc;
    std::lock_guard<std::mutex> lk(e->mu);          // snapshot, plan, measurement and restore under ONE lock: nobody sees the hypothetical state
    DeviceGuard guard(e->device);
    if (e->snap_bytes < e->occ_bytes) {
        if (e->d_occ_snap) cudaFree(e->d_occ_snap);
        e->d_occ_snap = nullptr; e->snap_bytes = 0;
        ISL_CUDA(e, cudaMalloc(&e->d_occ_snap, e->occ_bytes));
        e->snap_bytes = e->occ_bytes;
    }
    const uint32_t keep_snap_G = e->snap_G;         // a caller's own snapshot (isl_snapshot_occupancy) does not survive a what-if: say so
    (void)keep_snap_G;
    ISL_CUDA(e, cudaMemcpyAsync(e->d_occ_snap, e->d_occ, e->occ_bytes, cudaMemcpyDeviceToDevice, e->stream));
    int rc = ISL_OK;
    if (cap_before) rc = capacity_locked(e, cap_before);
    if (!rc && n) rc = place_batch_locked(e, n, plan, out);
    if (!rc && cap_after) rc = capacity_locked(e, cap_after);
    // the live state comes back whatever happened above
    const cudaError_t err = cudaMemcpyAsync(e->d_occ, e->d_occ_snap, e->occ_bytes, cudaMemcpyDeviceToDevice, e->stream);
    const cudaError_t err2 = cudaStreamSynchronize(e->stream);
    e->snap_G = 0;
    if (!rc && (err != cudaSuccess || err2 != cudaSuccess)) { snprintf(e->cuda_err, sizeof(e->cuda_err), "isl_what_if restore: %s", cudaGetErrorString(err != cudaSuccess ? err : err2)); rc = ISL_ECUDA; }
    return rc;
}

// ---- causal window / pinned buffers / owner-gathered results --------------------------------------------------------------
int isl_set_causal_window(isl_engine* e, uint32_t window) {
    if (!e) return ISL_EINVAL;
    if (e->open.active) return ISL_ESTATE;
    std::lock_guard<std::mutex> lk(e->mu);
    e->window = window;
    return ISL_OK;
}

// debugging aid (not part of the boundary): the per-round stamps recorded under ISL_SPEC_DBG=chunk,stage; out: kSpecRounds x 8 uint64
int isl_debug_spec_rounds(isl_engine* e, uint64_t* out, uint32_t max_words) {
    if (!e || !out || !e->d_specdbg) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    ISL_CUDA(e, cudaStreamSynchronize(e->stream));
    ISL_CUDA(e, cudaMemcpy(out, e->d_specdbg, std::min<size_t>(max_words, kSpecRounds * 8) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return ISL_OK;
}

int isl_set_speculation(isl_engine* e, uint32_t mode) {
    if (!e || mode > ISL_SPEC_ON) return ISL_EINVAL;
    if (e->open.active) return ISL_ESTATE;
    std::lock_guard<std::mutex> lk(e->mu);
    e->spec_mode = mode;
    return ISL_OK;
}

int isl_set_ring_world(isl_engine* e, uint32_t world) {
    if (!e) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    e->ring_world = world;
    return ISL_OK;
}

void* isl_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void isl_host_free(void* p) { if (p) cudaFreeHost(p); }

void* isl_device_results(isl_engine* e) { return e ? e->d_res : nullptr; }

int isl_ipc_results_handle(isl_engine* e, void* handle64) {
    if (!e || !handle64) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    cudaIpcMemHandle_t h;
    ISL_CUDA(e, cudaIpcGetMemHandle(&h, e->d_res));
    memcpy(handle64, &h, sizeof h);
    return ISL_OK;
}

int isl_ipc_connect_owner(isl_engine* e, const void* owner_handle64) {
    if (!e) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    if (e->d_owner_out && !e->owner_local) cudaIpcCloseMemHandle(e->d_owner_out);
    e->d_owner_out = nullptr; e->owner_local = false;
    if (owner_handle64) {
        cudaIpcMemHandle_t h;
        memcpy(&h, owner_handle64, sizeof h);
        void* p = nullptr;
        ISL_CUDA(e, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        e->d_owner_out = static_cast<uint2*>(p);
    }
    return ISL_OK;
}

int isl_connect_owner_local(isl_engine* e, isl_engine* owner) {
    if (!e) return ISL_EINVAL;
    std::lock_guard<std::mutex> lk(e->mu);
    if (e->d_owner_out && !e->owner_local) cudaIpcCloseMemHandle(e->d_owner_out);
    e->d_owner_out = owner ? owner->d_res : nullptr; e->owner_local = true;
    return ISL_OK;
}

// ---- open streams: the causal feed ----------------------------------------------------------------------------------------
// One persistent k_pipeline resolves batches that arrive WHILE it runs: isl_stream_submit copies a batch in and pre-passes it on the
// feed stream, the kernel's extra CTA writes its results into the caller's pinned buffer and raises a host-visible word, and
// isl_stream_wait returns as soon as that word is up — the caller composes batch b+1 (or b+k) from results it has already seen.
int isl_stream_open(isl_engine* e, uint32_t max_batches) {
    if (!e || max_batches == 0 || max_batches > kMaxStreamChunks) return ISL_EINVAL;
    if (!e->have_profiles || !e->have_inventory || e->open.active) return ISL_ESTATE;
    if (bestfit_family(e->cfg.policy)) return ISL_EINVAL;
    // a tool that serialises kernels (ncu, compute-sanitizer, CUDA_LAUNCH_BLOCKING) would starve a resident kernel that waits for kernels
    // launched after it: refuse instead of hanging until the device-side trap (callers fall back to isl_place_batch per batch)
    if (getenv("ISL_NO_FEED") || getenv("CUDA_INJECTION64_PATH") || getenv("CUDA_LAUNCH_BLOCKING") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR")) {
        snprintf(e->cuda_err, sizeof(e->cuda_err), "isl_stream_open: kernel-serialising tool or ISL_NO_FEED set; open streams need concurrent kernels");
        return ISL_ESTATE;
    }
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    auto& o = e->open;
    const uint32_t pc = e->pipe_chunk;
    if ((uint64_t)max_batches * pc > e->cfg.max_batch) return ISL_ERANGE;       // every batch owns a slot of the staging buffers
    uint32_t seg = 0, n_seg = 0, sub = 0;
    // speculative rounds: the caller says (isl_set_causal_window) that it keeps at most 1..3 batches in flight, or asks for them outright
    o.spec = false;
    {
        uint32_t mode = e->spec_mode;
        if (const char* v = getenv("ISL_SPEC")) mode = atoi(v) ? ISL_SPEC_ON : ISL_SPEC_OFF;
        if (kPipeThreads >= 208 && (mode == ISL_SPEC_ON || (mode == ISL_SPEC_AUTO && e->window >= 1 && e->window <= 3)) &&
            (uint64_t)max_batches * kSpecWordsPerChunk * 8ull <= (1ull << 30)) {
            const int src = plan_pipeline(e, std::max(2u, max_batches), (double)pc, true, &seg, &n_seg, &sub, true);
            if (src == ISL_ECUDA) return src;
            o.spec = src == ISL_OK && seg == sub && n_seg <= 148 && n_seg >= 2 && n_seg + 1 + kFeedReserve <= (uint32_t)e->max_coresident;
        }
    }
    if (!o.spec) if (int rc = plan_pipeline(e, std::max(2u, max_batches), (double)pc, true, &seg, &n_seg, &sub)) return rc;
    if (n_seg + 1 + kFeedReserve > (uint32_t)e->max_coresident) return ISL_ERANGE;  // the feed kernels need SMs next to the resident pipeline
    o.seg = seg; o.n_seg = n_seg; o.sub = sub; o.max_batches = max_batches; o.submitted = 0; o.launched = false;
    o.q_stride = pc + kQPad * ISL_MAX_PROFILES; o.free_stride = (uint32_t)e->occ_bytes; o.tiles_per_batch = pc / kTile;
    if (int rc = grow(e, &e->d_chunks, &e->cap_chunks, max_batches, 1)) return rc;
    if (int rc = grow(e, &e->d_cctl, &e->cap_cctl, max_batches, 1)) return rc;
    if (int rc = grow(e, &e->d_qall, &e->cap_qall, (size_t)max_batches * o.q_stride, 1)) return rc;
    if (int rc = grow(e, &e->d_tiles, &e->cap_tiles, (size_t)max_batches * o.tiles_per_batch, 1)) return rc;
    if (int rc = grow(e, &e->d_free_acc, &e->cap_free, (size_t)max_batches * (o.free_stride / 4), 1)) return rc;
    {
        const uint32_t before = e->cap_tokens;
        if (int rc = grow(e, &e->d_tokens, &e->cap_tokens, (size_t)max_batches * (n_seg + 1), kTokStride)) return rc;
        if (e->cap_tokens != before) ISL_CUDA(e, cudaMemsetAsync(e->d_tokens, 0, (size_t)e->cap_tokens * kTokStride * sizeof(uint32_t), e->stream));
    }
    if (int rc = grow(e, &e->d_ready, &e->cap_ready, max_batches + 1, 1)) return rc;
    if (int rc = grow(e, &e->d_done_cnt, &e->cap_done, max_batches, 1)) return rc;
    if ((uint64_t)max_batches * o.tiles_per_batch > ceil_div(e->cfg.max_batch, kTile) + 4096) return ISL_ERANGE;
    if (o.cap_done < max_batches) {
        if (o.h_done) cudaFreeHost(o.h_done);
        o.h_done = nullptr; o.cap_done = 0;
        ISL_CUDA(e, cudaHostAlloc(&o.h_done, (size_t)max_batches * sizeof(uint32_t), cudaHostAllocMapped));
        ISL_CUDA(e, cudaHostGetDevicePointer(&o.d_done_host, o.h_done, 0));
        o.cap_done = max_batches;
    }
    if (o.cap_desc < max_batches) {
        if (o.h_chunks) cudaFreeHost(o.h_chunks);
        if (o.h_tiles) cudaFreeHost(o.h_tiles);
        o.h_chunks = nullptr; o.h_tiles = nullptr; o.cap_desc = 0;
        ISL_CUDA(e, cudaHostAlloc(&o.h_chunks, (size_t)max_batches * sizeof(ChunkDesc), cudaHostAllocDefault));
        ISL_CUDA(e, cudaHostAlloc(&o.h_tiles, (size_t)max_batches * o.tiles_per_batch * sizeof(TileDesc), cudaHostAllocDefault));
        o.cap_desc = max_batches;
    }
    memset(o.h_done, 0, (size_t)max_batches * sizeof(uint32_t));
    if (!e->feed_stream) {
        ISL_CUDA(e, cudaStreamCreateWithFlags(&e->feed_stream, cudaStreamNonBlocking));
        ISL_CUDA(e, cudaEventCreateWithFlags(&e->ev_feed, cudaEventDisableTiming));
        ISL_CUDA(e, cudaEventCreateWithFlags(&e->ev_feed_done, cudaEventDisableTiming));
    }
    uint32_t epoch = ++e->epoch;
    if ((epoch & 0x7FFFu) == 0) epoch = ++e->epoch;
    if ((epoch & 0x7FFFu) == 1 && epoch != 1 && e->d_tokens)
        ISL_CUDA(e, cudaMemsetAsync(e->d_tokens, 0, (size_t)e->cap_tokens * kTokStride * sizeof(uint32_t), e->stream));
    o.epoch = epoch;
    if (o.spec) if (int rc = prepare_spec(e, max_batches, epoch, e->stream)) return rc;
    // the feed stream starts behind whatever the engine's stream still holds
    ISL_CUDA(e, cudaEventRecord(e->ev_feed, e->stream));
    ISL_CUDA(e, cudaStreamWaitEvent(e->feed_stream, e->ev_feed, 0));
    ISL_CUDA(e, cudaMemsetAsync(e->d_done_cnt, 0, (size_t)max_batches * sizeof(uint32_t), e->feed_stream));
    ISL_CUDA(e, cudaMemsetAsync(e->d_ready, 0, (size_t)(max_batches + 1) * sizeof(uint32_t), e->feed_stream));
    ISL_CUDA(e, cudaMemsetAsync(e->d_free_acc, 0, (size_t)max_batches * o.free_stride, e->feed_stream));
    o.active = true;
    return ISL_OK;
}

int isl_stream_submit(isl_engine* e, uint32_t n, const isl_request* in, isl_result* out, uint32_t* ticket) {
    if (!e || n == 0 || !in || !out) return ISL_EINVAL;
    if (!e->open.active) return ISL_ESTATE;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    auto& o = e->open;
    if (o.submitted >= o.max_batches || n > e->pipe_chunk) return ISL_ERANGE;
    // the results are written by the running kernel: the destination must be mapped pinned host memory (isl_host_alloc, cudaHostAlloc,
    // cudaHostRegister)
    cudaPointerAttributes pa{};
    if (cudaPointerGetAttributes(&pa, out) != cudaSuccess || pa.type != cudaMemoryTypeHost || !pa.devicePointer) { cudaGetLastError(); return ISL_EINVAL; }
    const uint32_t b = o.submitted, pc = e->pipe_chunk, off = b * pc, tile0 = b * o.tiles_per_batch, n_tiles = ceil_div(n, kTile);
    const cudaStream_t pre = e->feed_stream;
    o.h_chunks[b] = ChunkDesc{off, n, b, 1u, static_cast<uint2*>(pa.devicePointer), 0};
    for (uint32_t t = 0; t < n_tiles; ++t) o.h_tiles[tile0 + t] = TileDesc{off, n, b, tile0, b, tile0, n_tiles, off, n, 0, 0, 0};
    ISL_CUDA(e, cudaMemcpyAsync(e->d_chunks + b, o.h_chunks + b, sizeof(ChunkDesc), cudaMemcpyHostToDevice, pre));
    ISL_CUDA(e, cudaMemcpyAsync(e->d_tiles + tile0, o.h_tiles + tile0, (size_t)n_tiles * sizeof(TileDesc), cudaMemcpyHostToDevice, pre));
    ISL_CUDA(e, cudaMemcpyAsync(e->d_req + off, in, (size_t)n * sizeof(isl_request), cudaMemcpyHostToDevice, pre));
    k_prepare<<<n_tiles, kTileThreads, 0, pre>>>(0, e->d_req, e->d_res, reinterpret_cast<uint32_t*>(e->d_occ), e->G, e->lo, e->hi, e->prof,
                                                 e->d_tile_counts, e->d_ctrl, e->d_tiles, e->d_free_acc, o.free_stride / 4, tile0);
    if (int rc = check_launch(e, "k_prepare")) return rc;
    k_partition<<<n_tiles, kTileThreads, 0, pre>>>(0, e->d_req, e->prof.n, e->d_tile_counts, 0, e->cand_profiles, e->d_qall, e->d_cctl,
                                                   e->d_tiles, o.q_stride, tile0);
    if (int rc = check_launch(e, "k_partition")) return rc;
    k_set_flag<<<1, 1, 0, pre>>>(e->d_ready + b, o.epoch);
    if (int rc = check_launch(e, "k_set_flag")) return rc;
    if (!o.launched) {          // the persistent pipeline starts behind the first batch's tables
        ISL_CUDA(e, cudaEventRecord(e->ev_feed, pre));
        ISL_CUDA(e, cudaStreamWaitEvent(e->stream, e->ev_feed, 0));
        PipeArgs args{};
        args.n_chunks = o.max_batches; args.n_seg = o.n_seg; args.seg = o.seg; args.sub = o.sub; args.lo = e->lo; args.hi = e->hi; args.epoch = o.epoch;
        args.ready = e->d_ready; args.done_cnt = e->d_done_cnt; args.host_out = nullptr; args.copier = 1; args.open = 1;
        args.host_done = o.d_done_host; args.window = 0; args.wait_ns = std::max(kOpenWaitNs, e->wait_ns); args.flip = e->prof.flip;
        args.chunks = e->d_chunks; args.cctl = e->d_cctl; args.q_all = e->d_qall; args.free_acc = reinterpret_cast<const uint8_t*>(e->d_free_acc);
        args.q_stride = o.q_stride; args.free_stride = o.free_stride; args.tokens = e->d_tokens; args.occ = e->d_occ; args.gtab = e->d_gtab;
        args.out = e->d_res; args.feas = e->d_feas; args.stats = e->d_ctrl;
        if (o.spec) { args.spec = getenv("ISL_SPEC_NOREUSE") ? 3u : 1u; args.spec_mem = e->d_spec; args.spec_total = o.n_seg; }
        int rc;
        const bool p15 = e->prof.n == ISL_MAX_PROFILES;
        switch (e->n_cand_slots) {
            case 1: rc = p15 ? launch_pipeline<1, true>(e, args) : launch_pipeline<1, false>(e, args); break;
            case 2: rc = p15 ? launch_pipeline<2, true>(e, args) : launch_pipeline<2, false>(e, args); break;
            default: rc = p15 ? launch_pipeline<4, true>(e, args) : launch_pipeline<4, false>(e, args); break;
        }
        if (rc) return rc == ISL_ESTATE ? ISL_ERANGE : rc;
        o.launched = true;
    }
    if (ticket) *ticket = b;
    ++o.submitted;
    ++e->st.batches; e->st.requests += n;
    return ISL_OK;
}

int isl_stream_wait(isl_engine* e, uint32_t ticket) {
    if (!e) return ISL_EINVAL;
    auto& o = e->open;
    if (!o.active || ticket >= o.submitted) return ISL_ESTATE;
    // no lock: only reads a word the device raises; other threads may keep submitting
    volatile uint32_t* flag = o.h_done + ticket;
    const uint32_t epoch = o.epoch;
    uint64_t spins = 0;
    while (*flag != epoch) {
        if ((++spins & 0xFFFFu) == 0) {                 // now and then: did the pipeline die (a trap, an earlier launch error)?
            DeviceGuard guard(e->device);
            const cudaError_t err = cudaStreamQuery(e->stream);
            if (err != cudaSuccess && err != cudaErrorNotReady) { snprintf(e->cuda_err, sizeof(e->cuda_err), "isl_stream_wait: %s", cudaGetErrorString(err)); return ISL_ECUDA; }
            if (err == cudaSuccess && *flag != epoch) { snprintf(e->cuda_err, sizeof(e->cuda_err), "isl_stream_wait: pipeline ended before batch %u", ticket); return ISL_ECUDA; }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return ISL_OK;
}

int isl_stream_close(isl_engine* e) {
    if (!e) return ISL_EINVAL;
    if (!e->open.active) return ISL_ESTATE;
    std::lock_guard<std::mutex> lk(e->mu);
    DeviceGuard guard(e->device);
    auto& o = e->open;
    int rc = ISL_OK;
    if (o.launched) {
        if (o.submitted < o.max_batches) {
            k_set_flag<<<1, 1, 0, e->feed_stream>>>(e->d_ready + o.submitted, ~o.epoch);     // 'closed': the kernel leaves its chunk loop
            rc = check_launch(e, "k_set_flag");
        }
        cudaError_t err = cudaStreamSynchronize(e->feed_stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
        if (err != cudaSuccess) { snprintf(e->cuda_err, sizeof(e->cuda_err), "isl_stream_close: %s", cudaGetErrorString(err)); rc = ISL_ECUDA; }
    } else {
        cudaStreamSynchronize(e->feed_stream);
    }
    o.active = false; o.launched = false;
    return rc;
}

}  // extern "C"
