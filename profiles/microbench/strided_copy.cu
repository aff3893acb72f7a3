// strided_copy.cu -- how much HBM bandwidth does the access pattern of the strided NTT phase reach?  A CTA reads all 256 rows
// of W adjacent 8-byte columns of a 2^16-word limb (row stride 256 words = 2 KiB), 16 words per thread, and writes them back
// in place -- the strided phase without its butterflies.  W = 16 (128-byte row segments, 256 threads: the shipped tile),
// W = 32 (256-byte segments, 512 threads), W = 64 (512-byte segments, 1024 threads); tiles of a limb are walked fastest.
// Reference: a flat 128-bit copy over the same bytes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strided_copy strided_copy.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

template <int W>
__global__ void __launch_bounds__(W * 16) tile_copy(u64* p, int nlimbs_per_batch) {
    // grid: x = tile (256 / W), y = limb index (batch * limbs)
    const int t = threadIdx.x, col = t % W, g = t / W;  // g = 0..15: rows g, g+16, ...
    u64* base = p + (size_t)blockIdx.y * 65536 + blockIdx.x * W + col + (size_t)g * 256;
    u64 x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = base[(size_t)r * 16 * 256];
#pragma unroll
    for (int r = 0; r < 16; ++r) base[(size_t)r * 16 * 256] = x[r] + 1;
}
__global__ void __launch_bounds__(256) flat_copy(ulonglong2* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        ulonglong2 v = p[i];
        v.x += 1;
        p[i] = v;
    }
}
template <typename F>
static void timeit(const char* name, double bytes, F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.1f us  %7.1f GB/s (read + write)\n", name, ms * 100, bytes * 10 / ms * 1e-6);
}
int main() {
    const int limbs = 9 * 32 * 34;  // the digit launch of a 32-ciphertext step
    const size_t words = (size_t)limbs * 65536;
    u64* d;
    cudaMalloc(&d, words * 8);
    cudaMemset(d, 0, words * 8);
    const double bytes = 2.0 * words * 8;
    timeit("flat 128-bit read-modify-write", bytes, [&] { flat_copy<<<148 * 16, 256>>>((ulonglong2*)d, words / 2); });
    timeit("tiles of 16 columns (128 B segments)", bytes, [&] { tile_copy<16><<<dim3(16, limbs), 256>>>(d, 0); });
    timeit("tiles of 32 columns (256 B segments)", bytes, [&] { tile_copy<32><<<dim3(8, limbs), 512>>>(d, 0); });
    timeit("tiles of 64 columns (512 B segments)", bytes, [&] { tile_copy<64><<<dim3(4, limbs), 1024>>>(d, 0); });
    timeit("tiles of 16 columns again", bytes, [&] { tile_copy<16><<<dim3(16, limbs), 256>>>(d, 0); });
    return 0;
}
