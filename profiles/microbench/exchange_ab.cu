// exchange_ab.cu -- the 16 x 16 transpose between the two register blocks of the contiguous NTT phase: 16 threads own a
// 256-word segment, every thread holds 16 64-bit values x[r] (word cc + 16r) and needs the 16 consecutive words 16cc + r.
// north_star names "warp-shuffle butterflies"; the kernels exchange through shared memory instead.  This measures why.
//   mode 0: shared memory, XOR-swizzled (the shipped form: word 16r+cc at 16r + (cc^r); 16 STS.64 + 16 LDS.64 + 32 LOP3)
//   mode 1: shared memory, padded rows (word 16r+cc at 17r + cc, read 17cc + r; no LOP3, conflict-free as well)
//   mode 2: warp shuffles: four rounds of pairwise register swaps with the lane cc ^ 2^s (per pair and round: 2 SHFL.BFLY
//           + selects on either side, no shared memory)
// Between exchanges every thread does WORK FP64 operations per value so that the loop is not pure exchange (WORK = 0:
// exchange only; WORK = 32: the 4 stages x 8 instructions of an FP64 register block).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_ab exchange_ab.cu ; run: ./exchange_ab
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

template <int WORK>
__device__ __forceinline__ void work(u64 (&x)[16], double m) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        double v = __longlong_as_double((long long)x[r]);
#pragma unroll
        for (int k = 0; k < WORK; ++k) v = __fma_rn(v, m, 1.0);
        x[r] = (u64)__double_as_longlong(v);
    }
}

__device__ __forceinline__ void xchg_swz(u64 (&x)[16], u64* buf, u32 cc) {
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[16 * r + (cc ^ r)] = x[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = buf[16 * cc + (r ^ cc)];
    __syncwarp();
}
__device__ __forceinline__ void xchg_pad(u64 (&x)[16], u64* buf, u32 cc) {
#pragma unroll
    for (int r = 0; r < 16; ++r) buf[17 * r + cc] = x[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = buf[17 * cc + r];
    __syncwarp();
}
// after the four rounds x[r] of lane cc holds what lane r held in x[cc]
__device__ __forceinline__ void xchg_shfl(u64 (&x)[16], u32 cc) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const u32 bit = 1u << s;
        const bool up = (cc & bit) != 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (r & bit) continue;
            // the lane with the bit clear keeps x[r] and trades x[r|bit]; the lane with the bit set keeps x[r|bit], trades x[r]
            const u64 send = up ? x[r] : x[r | bit];
            const u32 lo = __shfl_xor_sync(0xffffffffu, (u32)send, bit);
            const u32 hi = __shfl_xor_sync(0xffffffffu, (u32)(send >> 32), bit);
            const u64 got = ((u64)hi << 32) | lo;
            if (up)
                x[r] = got;
            else
                x[r | bit] = got;
        }
    }
}

template <int MODE, int WORK>
__global__ void __launch_bounds__(128) loop(u64* a, int iters) {
    __shared__ u64 sm[8 * 272];
    const u32 t = threadIdx.x, sg = t >> 4, cc = t & 15;
    u64* buf = sm + sg * 272;
    u64* p = a + (size_t)blockIdx.x * 2048 + sg * 256;
    u64 x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = p[cc + 16 * r];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        work<WORK>(x, 0.999);
        if (MODE == 0)
            xchg_swz(x, buf, cc);
        else if (MODE == 1)
            xchg_pad(x, buf, cc);
        else
            xchg_shfl(x, cc);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) p[16 * cc + r] = x[r];
}

template <int MODE, int WORK>
static double run(u64* d, int ctas, int iters, const char* name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    loop<MODE, WORK><<<ctas, 128>>>(d, 4);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    loop<MODE, WORK><<<ctas, 128>>>(d, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ex = (double)ctas * 128 * iters;  // thread-exchanges (16 values each)
    printf("%-46s WORK=%2d  %8.3f ms  %7.2f G thread-exchanges/s  %6.2f clk/SM per 128-thread exchange\n", name, WORK, ms, ex / ms * 1e-6,
           ms * 1e-3 * 1.965e9 * 148 / ((double)ctas * iters));
    return ms;
}

int main() {
    const int ctas = 148 * 16, iters = 2000;
    u64 *d, *h = (u64*)malloc((size_t)ctas * 2048 * 8), *h2 = (u64*)malloc((size_t)ctas * 2048 * 8);
    cudaMalloc(&d, (size_t)ctas * 2048 * 8);
    // correctness: one exchange (WORK = 0, iters = 1) must be the transpose in every mode
    int bad = 0;
    for (int mode = 0; mode < 3; ++mode) {
        for (size_t i = 0; i < (size_t)ctas * 2048; ++i) h[i] = i * 0x9E3779B97F4A7C15ull;
        cudaMemcpy(d, h, (size_t)ctas * 2048 * 8, cudaMemcpyHostToDevice);
        if (mode == 0) loop<0, 0><<<ctas, 128>>>(d, 1);
        if (mode == 1) loop<1, 0><<<ctas, 128>>>(d, 1);
        if (mode == 2) loop<2, 0><<<ctas, 128>>>(d, 1);
        cudaMemcpy(h2, d, (size_t)ctas * 2048 * 8, cudaMemcpyDeviceToHost);
        // load x[r] = p[cc + 16r], exchange, store p[16cc + r] = x[r]: the composition is the identity on memory
        for (size_t i = 0; i < (size_t)ctas * 2048; ++i) bad += h[i] != h2[i];
        printf("mode %d: one exchange round-trips %s\n", mode, bad ? "WRONG" : "ok");
    }
    run<0, 0>(d, ctas, iters, "shared memory, XOR swizzle (shipped)");
    run<1, 0>(d, ctas, iters, "shared memory, padded rows (stride 17)");
    run<2, 0>(d, ctas, iters, "warp shuffles (4 rounds of pair swaps)");
    run<0, 32>(d, ctas, iters, "shared memory, XOR swizzle (shipped)");
    run<1, 32>(d, ctas, iters, "shared memory, padded rows (stride 17)");
    run<2, 32>(d, ctas, iters, "warp shuffles (4 rounds of pair swaps)");
    run<0, 0>(d, ctas, iters, "shared memory, XOR swizzle again (clock check)");
    return bad != 0;
}
