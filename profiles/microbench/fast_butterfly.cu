// fast_butterfly.cu -- candidate forward butterflies that are NOT the reference's instruction sequence
// but give the same final NTT (the forward transform of ring/ntt.go never wraps, so any exact lazy NTT
// followed by a canonical reduction is bit-identical).  Measures register-resident throughput and checks
// each candidate against the literal butterfly modulo q.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../lattigo-fhe-by-go_b200/csrc/modarith.cuh"

LG_DEV u64 mulw(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
LG_DEV u64 madw(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }

// KIND 0: literal (Montgomery, ntt.go:32-40)
// KIND 1: Harvey/Shoup, exact quotient, values kept in [0,4q):  w plain, ws = floor(w*2^64/q)
// KIND 2: Shoup with the quotient from the three high partial products only (error <= 2 => T < 4q),
//         no conditional subtraction at all (moduli < 2^56: 16 stages grow a value by at most 64q)
template <int KIND>
LG_DEV void bfly(u64& X, u64& Y, u64 w, u64 ws, u64 q, u64 qinv, u64 twoq, u64 fourq) {
    if (KIND == 0) {
        butterfly_fwd(X, Y, w, q, qinv, twoq);
    } else if (KIND == 1) {
        u64 x = X;
        if (x >= twoq) x -= twoq;
        const u64 qh = mul_hi(ws, Y);
        const u64 t = mul_lo(w, Y) - mul_lo(qh, q);
        X = x + t;
        Y = x + twoq - t;
    } else if (KIND == 2) {
        const u32 a0 = (u32)ws, a1 = (u32)(ws >> 32), b0 = (u32)Y, b1 = (u32)(Y >> 32);
        const u64 m1 = mulw(a1, b0), m2 = mulw(a0, b1);
        const u64 qh = madw(a1, b1, (m1 >> 32)) + (m2 >> 32);
        const u64 t = mul_lo(w, Y) - mul_lo(qh, q);
        const u64 x = X;
        X = x + t;
        Y = x + fourq - t;
    } else if (KIND == 3) {  // library: 16-instruction chain form
        butterfly_fwd_free(X, Y, w, ws, 0ull - q, fourq);
    } else if (KIND == 4) {
        butterfly_fwd_8q(X, Y, w, ws, 0ull - q, fourq);
    } else if (KIND == 5) {
        butterfly_inv_free(X, Y, w, ws, 0ull - q, q << 17);
    } else if (KIND == 6) {
        butterfly_inv_4q(X, Y, w, ws, 0ull - q, fourq);
    } else {  // FP64-assisted quotient: ws carries the bits of wd, qinv the constant c0
        const double wd = __longlong_as_double((long long)ws);
        if (KIND == 7) butterfly_fwd_f64(X, Y, w, wd, shoup_cw(wd), 0ull - q, fourq);
        else butterfly_inv_f64(X, Y, w, wd, shoup_cw(wd), 0ull - q, fourq);
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) loop(u64* a, const u64* __restrict__ tw, u64 q, u64 qinv, int iters) {
    u64 x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = a[threadIdx.x + 256 * r + blockIdx.x * 4096];
    const u64 twoq = 2 * q, fourq = 4 * q;
    u64 w[8], ws[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) { w[g] = tw[g + (threadIdx.x & 7)]; ws[g] = tw[64 + g + (threadIdx.x & 7)]; }
    if (KIND >= 7) {
#pragma unroll
        for (int g = 0; g < 8; ++g) ws[g] = (u64)__double_as_longlong(__ull2double_rz(ws[g]) * 5.421010862427522e-20);
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] &= 0x0003ffffffffffffull;
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) { const int u = 3 - uu;
#pragma unroll
            for (int g = 0; g < (16 >> (u + 1)); ++g)
#pragma unroll
                for (int k = 0; k < (1 << u); ++k) {
                    const int r = (g << (u + 1)) + k;
                    bfly<KIND>(x[r], x[r + (1 << u)], w[g], ws[g], q, qinv, twoq, fourq);
                }
        }
        if ((KIND == 2 || KIND == 3 || KIND == 5) && (it & 3) == 3) {  // keep the values bounded in this endless loop only
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] &= 0x00ffffffffffffffull;
        }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) a[threadIdx.x + 256 * r + blockIdx.x * 4096] = x[r];
}

// correctness: one butterfly on random data, compare with the literal one modulo q
template <int KIND>
__global__ void check(const u64* xs, const u64* ys, const u64* wm, u64 q, u64 qinv, u64 u0, u64 u1, int n, int* bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 wmont = wm[i] % q;
    const u64 wplain = invmform(wmont, q, qinv);
    const u64 ws = (u64)((((unsigned __int128)wplain) << 64) / q);
    u64 X0 = xs[i], Y0 = ys[i];
    if (KIND == 1 || KIND == 6) { X0 %= 4 * q; Y0 %= 4 * q; }
    if (KIND == 4) { X0 %= 8 * q; Y0 %= 8 * q; }
    if (KIND == 2 || KIND == 3) { X0 &= (1ull << 62) - 1; }
    if (KIND == 5) { X0 %= 2 * q; Y0 %= 2 * q; }
    if (KIND == 7) { X0 &= (1ull << 51) - 1; Y0 &= (1ull << 52) - 1; }
    if (KIND == 8) { X0 %= 4 * q; Y0 %= 4 * q; }
    u64 X1 = X0, Y1 = Y0, X2 = X0, Y2 = Y0;
    if (KIND == 5 || KIND == 6 || KIND == 8) {  // Gentleman-Sande: compare with the literal InvButterfly on in-range inputs
        X1 %= 2 * q; Y1 %= 2 * q;
        butterfly_inv(X1, Y1, wmont, q, qinv, 2 * q);
    } else {
        butterfly_fwd(X1, Y1, wmont, q, qinv, 2 * q);
    }
    if (KIND >= 7)
        bfly<KIND>(X2, Y2, wplain, (u64)__double_as_longlong(__ull2double_rz(ws) * 5.421010862427522e-20), q, qinv, 2 * q,
                   4 * q);
    else
        bfly<KIND>(X2, Y2, wplain, ws, q, qinv, 2 * q, 4 * q);
    if (KIND == 7 && (X2 >= X0 + 4 * q || Y2 > X0 + 4 * q)) atomicAdd(bad, 1);
    if (KIND == 8 && (X2 >= 4 * q || Y2 >= 4 * q)) atomicAdd(bad, 1);
    if (X1 % q != X2 % q || Y1 % q != Y2 % q) atomicAdd(bad, 1);
    if ((KIND == 1 || KIND == 6) && (X2 >= 4 * q || Y2 >= 4 * q)) atomicAdd(bad, 1);
    if (KIND == 4 && (X2 >= 8 * q || Y2 >= 8 * q)) atomicAdd(bad, 1);
    (void)u0; (void)u1;
}

template <int KIND>
void run(const char* name, u64* a, u64* tw) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const u64 q = 0x2000000a0001ull;
    const int iters = 256;
    dim3 grid(sms * 4);
    loop<KIND><<<grid, 256>>>(a, tw, q, 12345 | 1, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) loop<KIND><<<grid, 256>>>(a, tw, q, 12345 | 1, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double bf = (double)grid.x * 256 * iters * 32;
    // correctness on two moduli (45-bit and 55-bit CKKS primes; 61-bit only for KIND 1)
    const u64 qs[3] = {0x2000000a0001ull, 0x80000000080001ull, 0x10000000001d0001ull};
    int total_bad = 0;
    const int n = 1 << 20;
    u64 *xs, *ys, *wm; int* bad;
    cudaMalloc(&xs, n * 8); cudaMalloc(&ys, n * 8); cudaMalloc(&wm, n * 8); cudaMalloc(&bad, 4);
    u64* h = (u64*)malloc(n * 8);
    for (int rep = 0; rep < 3; ++rep) {
        u64* dst[3] = {xs, ys, wm};
        for (auto d : dst) { for (int i = 0; i < n; ++i) h[i] = ((u64)rand() << 43) ^ ((u64)rand() << 21) ^ rand(); cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice); }
    }
    for (int m = 0; m < ((KIND == 2 || KIND == 3) ? 2 : ((KIND == 5 || KIND >= 7) ? 1 : 3)); ++m) {
        const u64 qq = qs[m];
        u64 qinv = qq; for (int i = 0; i < 6; ++i) qinv *= 2 - qq * qinv;
        cudaMemset(bad, 0, 4);
        check<KIND><<<n / 256, 256>>>(xs, ys, wm, qq, qinv, 0, 0, n, bad);
        int hb; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); total_bad += hb;
    }
    printf("%-46s %.3e butterflies/s = %.2f /clk/SM   mismatches vs literal: %d\n", name, bf / (ms * 1e-3),
           bf / (ms * 1e-3) / (sms * (double)clk_khz * 1e3), total_bad);
}

int main() {
    u64 *a, *tw;
    cudaMalloc(&a, 148 * 8 * 4096 * 8); cudaMemset(a, 1, 148 * 8 * 4096 * 8);
    cudaMalloc(&tw, 4096); cudaMemset(tw, 3, 4096);
    run<0>("literal Montgomery butterfly", a, tw);
    run<1>("Shoup exact quotient, lazy [0,4q)", a, tw);
    run<2>("Shoup 3-product quotient, no cond. subtraction", a, tw);
    run<3>("butterfly_fwd_free (16-instruction chain form)", a, tw);
    run<4>("butterfly_fwd_8q   (values in [0,8q))", a, tw);
    run<5>("butterfly_inv_free (GS, q < 2^46)", a, tw);
    run<6>("butterfly_inv_4q   (GS, values in [0,4q))", a, tw);
    run<7>("butterfly_fwd_f64  (FP64 quotient, q < 3*2^44)", a, tw);
    run<8>("butterfly_inv_f64  (GS, FP64 quotient, [0,4q))", a, tw);
    return 0;
}
