// microbench_warp.cu — dependent-chain latency of the warp primitives the commit chain can be built from (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_warp.cu ; run: /tmp/mb
// Each kernel runs ONE warp with a loop-carried dependency through the primitive; cycles/iteration = latency of
// (primitive + the 1-2 ALU ops that feed it back).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 1 << 16;

__global__ void k_redux(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t x = seed + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) { uint32_t m = __reduce_min_sync(0xFFFFFFFFu, x); x = (x ^ m) + threadIdx.x; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (!threadIdx.x) cyc[0] = t1 - t0;
}
__global__ void k_ballot(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t x = seed + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) { uint32_t b = __ballot_sync(0xFFFFFFFFu, x & 1); x = (x ^ b) + threadIdx.x; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (!threadIdx.x) cyc[1] = t1 - t0;
}
__global__ void k_shfl(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t x = seed + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) { uint32_t s = __shfl_sync(0xFFFFFFFFu, x, x & 31); x = (x ^ s) + threadIdx.x; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (!threadIdx.x) cyc[2] = t1 - t0;
}
__global__ void k_lds(uint32_t* out, long long* cyc, uint32_t seed) {
    __shared__ uint32_t s[256];
    for (int i = threadIdx.x; i < 256; i += 32) s[i] = (i * 7 + seed) & 255;
    __syncwarp();
    uint32_t x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) { x = s[x & 255]; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (!threadIdx.x) cyc[3] = t1 - t0;
}
__global__ void k_alu(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t x = seed + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) { x = (x ^ (x >> 3)) + 0x9E3779B9u; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (!threadIdx.x) cyc[4] = t1 - t0;
}
__global__ void k_match(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t x = seed + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N / 16; ++i) { uint32_t b = __match_any_sync(0xFFFFFFFFu, x & 7); x = (x ^ b) + threadIdx.x; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (!threadIdx.x) cyc[5] = (t1 - t0) * 16;
}
// scalar min over 5 / 16 register values selected by a mask that depends on the carried value (single thread)
template <int P>
__global__ void k_scalar_min(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t t[P];
    for (int p = 0; p < P; ++p) t[p] = seed * (p + 3) + 17 * p;
    uint32_t x = seed;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) {
        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
        for (int p = 0; p < P; ++p) { uint32_t k = ((x >> p) & 1) ? t[p] : 0xFFFFFFFFu; best = min(best, k); }
        x = (x ^ best) * 5 + 1;
        t[i % P] += best & 3;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + t[0]; if (!threadIdx.x) cyc[P == 5 ? 6 : 7] = t1 - t0;
}
// redux followed by a dependent shared-memory load (the queue head refill on the critical path)
__global__ void k_redux_lds(uint32_t* out, long long* cyc, uint32_t seed) {
    __shared__ uint32_t s[256];
    for (int i = threadIdx.x; i < 256; i += 32) s[i] = (i * 7 + seed) & 255;
    __syncwarp();
    uint32_t x = seed + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) { uint32_t m = __reduce_min_sync(0xFFFFFFFFu, x); x = s[(m + threadIdx.x) & 255] + threadIdx.x; }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (!threadIdx.x) cyc[8] = t1 - t0;
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 1024); cudaMallocManaged(&cyc, 16 * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) {
        k_redux<<<1, 32>>>(out, cyc, 12345); k_ballot<<<1, 32>>>(out, cyc, 12345); k_shfl<<<1, 32>>>(out, cyc, 12345);
        k_lds<<<1, 32>>>(out, cyc, 3); k_alu<<<1, 32>>>(out, cyc, 12345); k_match<<<1, 32>>>(out, cyc, 12345);
        k_scalar_min<5><<<1, 1>>>(out, cyc, 12345); k_scalar_min<16><<<1, 1>>>(out, cyc, 12345); k_redux_lds<<<1, 32>>>(out, cyc, 5);
        cudaDeviceSynchronize();
    }
    const char* names[] = {"redux.min + 2 alu", "ballot + 2 alu", "shfl.idx + 2 alu", "lds (dependent)", "2 alu", "match.any + 2 alu",
                           "scalar min over 5 (1 thread)", "scalar min over 16 (1 thread)", "redux.min + dependent lds"};
    for (int i = 0; i < 9; ++i) printf("%-34s %7.1f cycles/iter\n", names[i], (double)cyc[i] / N);
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
