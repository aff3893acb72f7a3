// int_pipe.cu -- register-resident micro-benchmark of the integer pipes on sm_100a
// (the INT roofline is not in MEASURED_PEAKS.json and must be measured).
// Each kernel runs a long dependent-free stream of one instruction kind per thread
// (8 independent chains) and reports thread-instructions per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define ITERS 4096
#define CHAINS 8

template <int KIND>
__global__ void bench(u64* out, u32 a0, u32 b0) {
    u32 a[CHAINS], b[CHAINS], c[CHAINS]; u64 w[CHAINS]; double d[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = a0 + threadIdx.x + i; b[i] = b0 + i * 7 + 1; c[i] = i; w[i] = a[i]; d[i] = a[i]; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
            if (KIND == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
            if (KIND == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
            if (KIND == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (KIND == 4) asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));
            if (KIND == 5) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001));
            if (KIND == 6) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((u64)b[i] << 29 | 12345));
            if (KIND == 7) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((u64)b[i] << 29 | 12345));
            if (KIND == 8) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (KIND == 9) asm volatile("{.reg .u32 t; add.cc.u32 %0, %0, %1; addc.u32 %2, %2, %1;}" : "+r"(a[i]), "+r"(c[i]) : "r"(b[i]));
            if (KIND == 10) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(b[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }
            if (KIND == 11) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(b[i])); }
            if (KIND == 13) { u64 t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(b[i])); a[i] = (u32)t ^ (u32)(t >> 32); }
            if (KIND == 14) { u64 t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(b[i])); w[i] += t; a[i] += 1; }
            if (KIND == 15) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b0)); }
            if (KIND == 12) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i])); asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001)); }
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc += a[i] + c[i] + w[i] + (u64)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int KIND>
void run(const char* name, int per_iter, u64* out) {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    dim3 grid(sms * 4), block(256);
    bench<KIND><<<grid, block>>>(out, 1, 3);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) bench<KIND><<<grid, block>>>(out, 1, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double insts = (double)grid.x * block.x * ITERS * CHAINS * per_iter;
    double per_s = insts / (ms * 1e-3);
    printf("%-34s %8.1f Gthread-inst/s  = %6.1f inst/clk/SM at %d MHz nominal (%.3f ms)\n", name, per_s / 1e9,
           per_s / (sms * (double)clk_khz * 1e3), clk_khz / 1000, ms);
}

int main() {
    u64* out; cudaMalloc(&out, 148 * 4 * 256 * 8 * 2);
    run<0>("IMAD (mad.lo.u32)", 1, out);
    run<1>("IMAD.HI (mad.hi.u32)", 1, out);
    run<8>("mul.hi.u32", 1, out);
    run<2>("IMAD.WIDE.U32 (mad.wide.u32)", 1, out);
    run<3>("IADD3 (add.u32)", 1, out);
    run<4>("ISETP+SEL pair (counted as 2)", 2, out);
    run<9>("IADD3 + IADD3.X carry pair (2)", 2, out);
    run<5>("DFMA (fma.rn.f64)", 1, out);
    run<6>("mul.hi.u64 (counted as 1)", 1, out);
    run<7>("mul.lo.u64 (counted as 1)", 1, out);
    run<10>("IMAD.WIDE + 2 IADD (counted as 3)", 3, out);
    run<11>("IMAD + IADD (counted as 2)", 2, out);
    run<12>("IMAD.WIDE + DFMA (counted as 2)", 2, out);
    run<13>("mul.wide (no acc) + LOP3 (2)", 2, out);
    run<14>("mul.wide + 64-bit add + add (4)", 4, out);
    run<15>("IMAD.WIDE uniform b operand", 1, out);
    return 0;
}
