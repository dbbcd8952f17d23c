// microbench_pred.cu — what does a loop-carried dependency cost when it runs through a predicate, and what through the redux?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_pred tools/microbench_pred.cu ; one warp per kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int N = 1 << 14;
#define LOOP4(...) for (int i = 0; i < N; i += 4) { __VA_ARGS__ __VA_ARGS__ __VA_ARGS__ __VA_ARGS__ }

__global__ void k_lop(uint32_t* out, long long* cyc, uint32_t c) {          // 2 dependent LOP3
    uint32_t x = threadIdx.x + c;
    long long t0 = clock64();
    LOOP4(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(c), "r"(c + 7)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x) : "r"(c + 1), "r"(c + 3));)
    long long t1 = clock64(); out[threadIdx.x] = x; if (!threadIdx.x) cyc[0] = t1 - t0;
}
__global__ void k_setp_selp(uint32_t* out, long long* cyc, uint32_t c) {    // setp -> selp, twice (2 predicate hops)
    uint32_t x = threadIdx.x + c, a = c * 3 + 1, b = c * 5 + 2;
    long long t0 = clock64();
    LOOP4(asm volatile("{ .reg .pred p; setp.lt.s32 p, %0, 0; selp.b32 %0, %1, %2, p; }" : "+r"(x) : "r"(a), "r"(b));
          asm volatile("{ .reg .pred p; setp.lt.s32 p, %0, 0; selp.b32 %0, %2, %1, p; }" : "+r"(x) : "r"(a ^ 0x80000000u), "r"(b | 0x80000000u));)
    long long t1 = clock64(); out[threadIdx.x] = x; if (!threadIdx.x) cyc[1] = t1 - t0;
}
__global__ void k_lop_p_selp(uint32_t* out, long long* cyc, uint32_t c) {   // (x & m) == 0 ? a : b  twice
    uint32_t x = threadIdx.x + c, a = c * 3 + 1, b = c * 5 + 2;
    long long t0 = clock64();
    LOOP4(asm volatile("{ .reg .pred p; .reg .b32 t; and.b32 t, %0, %3; setp.eq.u32 p, t, 0; selp.b32 %0, %1, %2, p; }" : "+r"(x) : "r"(a), "r"(b), "r"(0x11u));
          asm volatile("{ .reg .pred p; .reg .b32 t; and.b32 t, %0, %3; setp.eq.u32 p, t, 0; selp.b32 %0, %2, %1, p; }" : "+r"(x) : "r"(a + 16), "r"(b + 1), "r"(0x22u));)
    long long t1 = clock64(); out[threadIdx.x] = x; if (!threadIdx.x) cyc[2] = t1 - t0;
}
__global__ void k_mask_mux(uint32_t* out, long long* cyc, uint32_t c) {     // sign mask + bitwise mux, twice (no predicate)
    uint32_t x = threadIdx.x + c, a = c * 3 + 1, b = c * 5 + 2;
    long long t0 = clock64();
    LOOP4(asm volatile("{ .reg .b32 m; shr.s32 m, %0, 31; lop3.b32 %0, %1, %2, m, 0xca; }" : "+r"(x) : "r"(a), "r"(b));       // m ? a : b bitwise... 0xca = (c&a)|(~c&b) with operand order (a,b,c)
          asm volatile("{ .reg .b32 m; shr.s32 m, %0, 31; lop3.b32 %0, %2, %1, m, 0xca; }" : "+r"(x) : "r"(a ^ 0x80000000u), "r"(b | 0x80000000u));)
    long long t1 = clock64(); out[threadIdx.x] = x; if (!threadIdx.x) cyc[3] = t1 - t0;
}
__global__ void k_redux_only(uint32_t* out, long long* cyc, uint32_t c) {   // redux -> 1 lop3 -> redux
    uint32_t x = threadIdx.x + c;
    long long t0 = clock64();
    LOOP4({ uint32_t m; asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(x)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(m), "r"(c)); })
    long long t1 = clock64(); out[threadIdx.x] = x; if (!threadIdx.x) cyc[4] = t1 - t0;
}
// the decision loop's o-path: LOP3.P -> SEL -> REDUX -> setp(sel) -> selp -> lop3
__global__ void k_opath_pred(uint32_t* out, long long* cyc, uint32_t c) {
    uint32_t o0 = threadIdx.x & 0x7f, o1 = (threadIdx.x * 7) & 0x7f, t = (threadIdx.x << 15) | 0x101, cm = 1u << (threadIdx.x & 7);
    long long t0 = clock64();
    LOOP4({ uint32_t key, m;
            asm volatile("{ .reg .pred p; .reg .b32 z; and.b32 z, %1, %2; setp.eq.u32 p, z, 0; selp.b32 %0, %3, 0xffffffff, p; }" : "=r"(key) : "r"(o0), "r"(cm), "r"(t));
            asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(key));
            asm volatile("{ .reg .pred p; .reg .b32 s; setp.lt.s32 p, %2, 0; selp.b32 s, %1, %0, p; lop3.b32 %0, s, %2, 0x3, 0xf8; }" : "+r"(o0) : "r"(o1), "r"(m));
            o1 = o1 * 5 + 1; })
    long long t1 = clock64(); out[threadIdx.x] = o0; if (!threadIdx.x) cyc[5] = t1 - t0;
}
// the same path with a sign mask instead of the sel predicate
__global__ void k_opath_mask(uint32_t* out, long long* cyc, uint32_t c) {
    uint32_t o0 = threadIdx.x & 0x7f, o1 = (threadIdx.x * 7) & 0x7f, t = (threadIdx.x << 15) | 0x101, cm = 1u << (threadIdx.x & 7);
    long long t0 = clock64();
    LOOP4({ uint32_t key, m;
            asm volatile("{ .reg .pred p; .reg .b32 z; and.b32 z, %1, %2; setp.eq.u32 p, z, 0; selp.b32 %0, %3, 0xffffffff, p; }" : "=r"(key) : "r"(o0), "r"(cm), "r"(t));
            asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(key));
            asm volatile("{ .reg .b32 s, k; shr.s32 k, %2, 31; lop3.b32 s, %1, %0, k, 0xca; lop3.b32 %0, s, %2, 0x3, 0xf8; }" : "+r"(o0) : "r"(o1), "r"(m));
            o1 = o1 * 5 + 1; })
    long long t1 = clock64(); out[threadIdx.x] = o0; if (!threadIdx.x) cyc[6] = t1 - t0;
}
// redux followed by a uniform-looking branch on its result (the rare-path test), then one lop3
__global__ void k_redux_branch(uint32_t* out, long long* cyc, uint32_t c) {
    uint32_t x = threadIdx.x + c, acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) {
        uint32_t m; asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(m) : "r"(x));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(m), "r"(c));
        if (m == 0xFFFFFFFFu) { acc += x; x = x * 3 + 1; }
    }
    long long t1 = clock64(); out[threadIdx.x] = x + acc; if (!threadIdx.x) cyc[7] = t1 - t0;
}
int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 1024); cudaMallocManaged(&cyc, 16 * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) {
        k_lop<<<1, 32>>>(out, cyc, 12345); k_setp_selp<<<1, 32>>>(out, cyc, 12345); k_lop_p_selp<<<1, 32>>>(out, cyc, 12345); k_mask_mux<<<1, 32>>>(out, cyc, 12345);
        k_redux_only<<<1, 32>>>(out, cyc, 12345); k_opath_pred<<<1, 32>>>(out, cyc, 12345); k_opath_mask<<<1, 32>>>(out, cyc, 12345); k_redux_branch<<<1, 32>>>(out, cyc, 12345);
        cudaDeviceSynchronize();
    }
    const char* names[] = {"2 x lop3", "2 x (setp -> selp)", "2 x (and -> setp -> selp)", "2 x (shr.s32 -> lop3 mux)", "redux -> lop3", "o-path with predicates (and,setp,selp | redux | setp,selp,lop3)",
                           "o-path with sign mask (and,setp,selp | redux | shr,lop3,lop3)", "redux -> lop3 -> branch on m (not unrolled)"};
    for (int i = 0; i < 8; ++i) printf("%-66s %7.1f cycles/iter\n", names[i], (double)cyc[i] / N);
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
