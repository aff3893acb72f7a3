// permute.cu -- kernel family K4: Galois automorphisms and index permutations.
//
// Replaces ring/ring_galois.go:55-127 (PermuteNTT / PermuteNTTWithIndex gather,
// Context.Permute signed coefficient permutation) and ring/ring.go:663-772
// (MultByMonomial, BitReverse).  None of them is in place (ring_galois.go:54).
#include "common.cuh"
#include "kernels.h"

namespace {

// PermuteNTTWithIndex, ring_galois.go:89-101: out[l][j] = in[l][index[j]]
__global__ void __launch_bounds__(256) permute_ntt_kernel(const PermArgs a) {
    const int l = blockIdx.y, bt = blockIdx.z;
    const u64* in = a.in + bt * a.in_bs + (size_t)l * a.T.N;
    u64* out = a.out + bt * a.out_bs + (size_t)l * a.T.N;
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < a.T.N; j += gridDim.x * blockDim.x)
        out[j] = in[__ldg(a.index + j)];
}

// Context.Permute, ring_galois.go:106-127
__global__ void __launch_bounds__(256) permute_coeff_kernel(const PermArgs a) {
    const int l = blockIdx.y, bt = blockIdx.z;
    const u64 q = a.T.q[a.map(l)];
    const u64* in = a.in + bt * a.in_bs + (size_t)l * a.T.N;
    u64* out = a.out + bt * a.out_bs + (size_t)l * a.T.N;
    const u64 mask = a.T.N - 1;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < a.T.N; i += gridDim.x * blockDim.x) {
        const u64 raw = (u64)i * a.gen;
        const u64 index = raw & mask;
        const u64 tmp = (raw >> a.T.logN) & 1;
        const u64 v = in[i];
        out[index] = (v * (tmp ^ 1)) | ((q - v) * tmp);
    }
}

// MultByMonomial, ring.go:663-723 (gen = monomialDeg)
__global__ void __launch_bounds__(256) monomial_kernel(const PermArgs a) {
    const int l = blockIdx.y, bt = blockIdx.z;
    const u64 q = a.T.q[a.map(l)];
    const u64* in = a.in + bt * a.in_bs + (size_t)l * a.T.N;
    u64* out = a.out + bt * a.out_bs + (size_t)l * a.T.N;
    const u64 N = a.T.N;
    u64 shift = a.gen % (N << 1);
    const bool neg = shift >= N;  // tmpx = q - p1 when shift >= N
    const bool zero = shift == 0;
    shift %= N;
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        if (zero) {
            out[j] = in[j];
        } else if (j < shift) {
            u64 t = in[N - shift + j];
            if (neg) t = q - t;
            out[j] = q - t;
        } else {
            u64 t = in[j - shift];
            if (neg) t = q - t;
            out[j] = t;
        }
    }
}

// BitReverse, ring.go:749-772
__global__ void __launch_bounds__(256) bitrev_kernel(const PermArgs a) {
    const int l = blockIdx.y, bt = blockIdx.z;
    const u64* in = a.in + bt * a.in_bs + (size_t)l * a.T.N;
    u64* out = a.out + bt * a.out_bs + (size_t)l * a.T.N;
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < a.T.N; j += gridDim.x * blockDim.x)
        out[__brev(j) >> (32 - a.T.logN)] = in[j];
}

dim3 perm_grid(u32 N, int nlimbs, int batch) {
    u32 bx = (N + 255) / 256;
    if (bx > 64) bx = 64;
    return dim3(bx, nlimbs, batch);
}

// big-endian <-> host-endian words (ring/ring_object.go:146-156, :196-206: binary.BigEndian.PutUint64 / Uint64)
__global__ void __launch_bounds__(256) bswap64_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 x = in[i];
        const u32 lo = (u32)x, hi = (u32)(x >> 32);
        out[i] = ((u64)__byte_perm(lo, 0, 0x0123) << 32) | (u64)__byte_perm(hi, 0, 0x0123);
    }
}

}  // namespace

int lg_launch_permute_ntt(const PermArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    permute_ntt_kernel<<<perm_grid(a.T.N, nlimbs, batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    return 0;
}
int lg_launch_permute_coeff(const PermArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    permute_coeff_kernel<<<perm_grid(a.T.N, nlimbs, batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    return 0;
}
int lg_launch_mult_by_monomial(const PermArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    monomial_kernel<<<perm_grid(a.T.N, nlimbs, batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    return 0;
}
int lg_launch_bitreverse(const PermArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    bitrev_kernel<<<perm_grid(a.T.N, nlimbs, batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    return 0;
}

int lg_launch_bswap64(const u64* in, u64* out, size_t words, cudaStream_t st) {
    if (words == 0) return 0;
    size_t bx = (words + 255) / 256;
    if (bx > 148 * 8) bx = 148 * 8;
    bswap64_kernel<<<(unsigned)bx, 256, 0, st>>>(in, out, words);
    lg_g_launches += 1;
    return 0;
}
