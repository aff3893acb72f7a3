// fp64_butterfly.cu -- can the forward butterfly of the 45-bit limbs run on the FP64 pipe ALONE, and do FP64-only
// warps and integer warps overlap on one SM?  (B200 has a full-rate FP64 pipe next to the integer pipes.)
//
// FP64-only Cooley-Tukey butterfly on signed lazy values kept as doubles (exact integers, |v| < 2^51):
//     h  = RN(w*y)                      l = fma(w, y, -h)        (exact low part)
//     qf = RD(y*wd + 1.5*2^52)          qh = qf - 1.5*2^52       (wd = RD(w/q): qh = floor(y*w/q) + {-1, 0, +1})
//     d  = fma(-qh, q, h)  (exact: |d| < 2^48)                    t = d + l  in [-q, 2q)
//     X' = X + t,  Y' = X - t
// 8 FP64-pipe instructions, no integer instruction.  Values grow by at most 2q per stage: 16 stages of a modulus
// below 3*2^44 stay below 33q < 2^51.
// Modes: 0 = integer butterfly_fwd_f64 (the shipped one: 12 INT + 1 DFMA), 1 = FP64-only, 2 = per-warp mix (even
// warps FP64-only, odd warps integer), 3 = per-warp mix 3:1 (three FP64 warps per integer warp).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../lattigo-fhe-by-go_b200/csrc/modarith.cuh"

#define MAGIC 6755399441055744.0  // 1.5 * 2^52

LG_DEV void bfly_d64(double& X, double& Y, double w, double wd, double q) {
    const double y = Y;
    const double h = __dmul_rn(w, y);
    const double l = __fma_rn(w, y, -h);
    const double qf = __fma_rd(y, wd, MAGIC);
    const double qh = __dadd_rn(qf, -MAGIC);
    const double d = __fma_rn(-qh, q, h);
    const double t = __dadd_rn(d, l);
    const double x = X;
    X = __dadd_rn(x, t);
    Y = __dadd_rn(x, -t);
}

template <int MODE>
__global__ void __launch_bounds__(256) loop(u64* a, const u64* __restrict__ tw, u64 q, int iters) {
    const int warp = threadIdx.x >> 5;
    const bool fp = MODE == 1 || (MODE == 2 && (warp & 1) == 0) || (MODE == 3 && (warp & 3) != 3);
    const size_t base = threadIdx.x + blockIdx.x * 4096;
    if (fp) {
        double x[16], w[8], wd[8];
        const double qd = (double)q;
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = (double)(a[base + 256 * r] % q);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const u64 wp = tw[g + (threadIdx.x & 7)] % q;
            w[g] = (double)wp;
            wd[g] = __ddiv_rd(w[g], qd);
        }
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const int u = 3 - uu;
#pragma unroll
                for (int g = 0; g < (16 >> (u + 1)); ++g)
#pragma unroll
                    for (int k = 0; k < (1 << u); ++k) {
                        const int r = (g << (u + 1)) + k;
                        bfly_d64(x[r], x[r + (1 << u)], w[g], wd[g], qd);
                    }
            }
            if ((it & 3) == 3) {  // keep the endless loop bounded (a real transform has 16 stages)
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double c = __dadd_rn(__fma_rd(x[r], 1.0 / qd, MAGIC), -MAGIC);
                    x[r] = __fma_rn(-c, qd, x[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) a[base + 256 * r] = (u64)(long long)x[r];
    } else {
        u64 x[16], w[8], ws[8];
        const u64 fourq = 4 * q;
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = a[base + 256 * r] & 0x0003ffffffffffffull;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            w[g] = tw[g + (threadIdx.x & 7)];
            ws[g] = (u64)__double_as_longlong(__ull2double_rz(tw[64 + g + (threadIdx.x & 7)]) * 5.421010862427522e-20);
        }
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const int u = 3 - uu;
#pragma unroll
                for (int g = 0; g < (16 >> (u + 1)); ++g)
#pragma unroll
                    for (int k = 0; k < (1 << u); ++k) {
                        const int r = (g << (u + 1)) + k;
                        const double wd = __longlong_as_double((long long)ws[g]);
                        butterfly_fwd_f64(x[r], x[r + (1 << u)], w[g], wd, shoup_cw(wd), 0ull - q, fourq);
                    }
            }
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) a[base + 256 * r] = x[r];
    }
}

// correctness of the FP64-only butterfly against the literal one, modulo q, on signed lazy inputs
__global__ void check(const u64* xs, const u64* ys, const u64* wm, u64 q, u64 qinv, int n, int* bad, double* tmin, double* tmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 wmont = wm[i] % q;
    const u64 wplain = invmform(wmont, q, qinv);
    // signed lazy inputs in (-32q, 32q)
    const long long sx = (long long)(xs[i] % (64 * q)) - (long long)(32 * q);
    const long long sy = (long long)(ys[i] % (64 * q)) - (long long)(32 * q);
    u64 X1 = (u64)((sx % (long long)q + (long long)q) % (long long)q), Y1 = (u64)((sy % (long long)q + (long long)q) % (long long)q);
    butterfly_fwd(X1, Y1, wmont, q, qinv, 2 * q);
    double X = (double)sx, Y = (double)sy;
    const double qd = (double)q, w = (double)wplain, wd = __ddiv_rd(w, qd);
    bfly_d64(X, Y, w, wd, qd);
    const double t = (X - Y) * 0.5;
    if (t < -qd || t >= 2 * qd) atomicAdd(bad, 1);
    const long long rx = (long long)X, ry = (long long)Y;
    const u64 X2 = (u64)((rx % (long long)q + (long long)q) % (long long)q), Y2 = (u64)((ry % (long long)q + (long long)q) % (long long)q);
    if (X1 % q != X2 || Y1 % q != Y2) atomicAdd(bad, 1);
    (void)tmin;
    (void)tmax;
}

template <int MODE>
void run(const char* name, u64* a, u64* tw) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const u64 q = 0x2000000a0001ull;
    const int iters = 512;
    dim3 grid(sms * 4);
    loop<MODE><<<grid, 256>>>(a, tw, q, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 40;  // ~long enough for the power management to settle
    for (int r = 0; r < reps; ++r) loop<MODE><<<grid, 256>>>(a, tw, q, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double bf = (double)grid.x * 256 * iters * 32;
    printf("%-58s %.3e butterflies/s = %.2f /clk/SM at the nominal %d MHz  (%.2f ms per launch, %d launches)\n", name,
           bf / (ms * 1e-3), bf / (ms * 1e-3) / (sms * (double)clk_khz * 1e3), clk_khz / 1000, ms, reps);
    fflush(stdout);
}

int main() {
    u64 *a, *tw;
    cudaMalloc(&a, 148 * 8 * 4096 * 8);
    cudaMemset(a, 1, 148 * 8 * 4096 * 8);
    cudaMalloc(&tw, 4096);
    cudaMemset(tw, 3, 4096);
    // correctness first
    {
        const u64 qs[2] = {0x2000000a0001ull, 0x2ffffffe30001ull};  // a 45-bit CKKS prime; a value just below 3*2^44 (odd)
        const int n = 1 << 20;
        u64 *xs, *ys, *wm;
        int* bad;
        cudaMalloc(&xs, n * 8);
        cudaMalloc(&ys, n * 8);
        cudaMalloc(&wm, n * 8);
        cudaMalloc(&bad, 4);
        u64* h = (u64*)malloc(n * 8);
        u64* dst[3] = {xs, ys, wm};
        for (auto d : dst) {
            for (int i = 0; i < n; ++i) h[i] = ((u64)rand() << 43) ^ ((u64)rand() << 21) ^ rand();
            cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
        }
        for (int m = 0; m < 2; ++m) {
            const u64 qq = qs[m];
            u64 qinv = qq;
            for (int i = 0; i < 6; ++i) qinv *= 2 - qq * qinv;
            cudaMemset(bad, 0, 4);
            check<<<n / 256, 256>>>(xs, ys, wm, qq, qinv, n, bad, nullptr, nullptr);
            int hb;
            cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
            printf("FP64-only butterfly vs literal, q = %llu: %d mismatches / out-of-range products in %d\n", (unsigned long long)qq, hb, n);
        }
    }
    run<0>("integer butterfly_fwd_f64 (12 INT + 1 DFMA)", a, tw);
    run<1>("FP64-only butterfly (8 FP64 instructions)", a, tw);
    run<2>("mix by warp: 1 FP64-only : 1 integer", a, tw);
    run<3>("mix by warp: 3 FP64-only : 1 integer", a, tw);
    run<0>("integer butterfly_fwd_f64 again (clock check)", a, tw);
    return 0;
}
