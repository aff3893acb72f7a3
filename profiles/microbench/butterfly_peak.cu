// butterfly_peak.cu -- register-resident ceiling of the literal Montgomery butterfly of ring/ntt.go:32-40
// (the exact device code of csrc/modarith.cuh, no memory traffic inside the loop).  The NTT kernels are
// INT-pipe bound, so this is the roofline they are measured against: butterflies per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lattigo-fhe-by-go_b200/csrc/modarith.cuh"

template <bool FWD>
__global__ void __launch_bounds__(256) loop(u64* a, const u64* __restrict__ tw, u64 q, u64 qinv, int iters) {
    u64 x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = a[threadIdx.x + 256 * r + blockIdx.x * 4096];
    const u64 twoq = 2 * q;
    u64 w[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) w[g] = tw[g + (threadIdx.x & 7)];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) { const int u = 3 - uu;
#pragma unroll
            for (int g = 0; g < (16 >> (u + 1)); ++g) {
#pragma unroll
                for (int k = 0; k < (1 << u); ++k) {
                    const int r = (g << (u + 1)) + k;
                    if (FWD) butterfly_fwd(x[r], x[r + (1 << u)], w[g], q, qinv, twoq);
                    else butterfly_inv(x[r], x[r + (1 << u)], w[g], q, qinv, twoq);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) a[threadIdx.x + 256 * r + blockIdx.x * 4096] = x[r];
}

template <bool FWD>
void run(const char* name, u64* a, u64* tw, int ctas_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const u64 q = 0x2000000a0001ull, qinv = 0;  // any odd modulus; values are irrelevant for timing
    const int iters = 256;
    dim3 grid(sms * ctas_per_sm);
    loop<FWD><<<grid, 256>>>(a, tw, q, qinv | 1, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) loop<FWD><<<grid, 256>>>(a, tw, q, qinv | 1, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double bf = (double)grid.x * 256 * iters * 32;
    printf("%s, %d CTAs/SM: %.3e butterflies/s = %.2f butterflies/clk/SM at %d MHz nominal\n", name, ctas_per_sm,
           bf / (ms * 1e-3), bf / (ms * 1e-3) / (sms * (double)clk_khz * 1e3), clk_khz / 1000);
}

int main() {
    u64 *a, *tw;
    cudaMalloc(&a, 148 * 8 * 4096 * 8); cudaMemset(a, 1, 148 * 8 * 4096 * 8);
    cudaMalloc(&tw, 4096); cudaMemset(tw, 3, 4096);
    for (int c : {2, 3, 4, 6}) { run<true>("forward butterfly (ntt.go:32-40)", a, tw, c); }
    for (int c : {3, 4}) { run<false>("inverse butterfly (ntt.go:43-50)", a, tw, c); }
    return 0;
}
