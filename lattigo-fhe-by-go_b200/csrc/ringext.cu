// ringext.cu -- the ring.Context methods that are not on the evaluator path but belong to the type's method set:
// MulPoly / MulPolyMontgomery / MulPolyNaive / MulPolyNaiveMontgomery (ring/ring.go:357-437), Exp (:439-464),
// Shift (:574-580), Rotate (:772-800) and Equal / EqualLvl (ring/ring_context.go:423-467).  Literal restatements: the
// reference's quirks are kept (Rotate writes into p1 and never touches p2, Exp ends with InvNTT(p1) in p2, Equal reduces
// both operands in place).
#include <string.h>

#include "capi_internal.hpp"

namespace {

// n & ((1 << N) - 1) as Go evaluates it on uint64: a shift by 64 or more gives 0, so the mask is all ones for N >= 64
u64 go_mask(u64 N) { return N >= 64 ? ~0ull : ((1ull << N) - 1); }

struct ExtArgs {
    RingTables T;
    const u64* a;
    const u64* b;
    u64* c;
    size_t a_bs, b_bs, c_bs;
    u64 n;
    u64 s[LG_MAX_LIMBS];
};

// Shift, ring.go:575-580: p2[k] = p1[(k + n) mod N] (append(p1[n:], p1[:n]...))
__global__ void __launch_bounds__(256) shift_kernel(const ExtArgs a) {
    const int l = blockIdx.y, bt = blockIdx.z;
    const u32 N = a.T.N, n = (u32)a.n;
    const u64* in = a.a + bt * a.a_bs + (size_t)l * N;
    u64* out = a.c + bt * a.c_bs + (size_t)l * N;
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) {
        const u32 src = k + n;
        out[k] = in[src >= N ? src - N : src];
    }
}

// Rotate, ring.go:775-800: p1[j] = MRed(p1[j], gal_j) for j >= 1, gal_j = MRed(gal_{j-1}, root), gal_0 = MForm(1): every
// gal_j is the canonical Montgomery form of root^j, which modexpMontgomery (ring/utils.go:39-50) from the same start gives
// as well; s[l] = root of limb l (Montgomery form), coefficient 0 is left alone
__global__ void __launch_bounds__(256) rotate_kernel(const ExtArgs a) {
    const int l = blockIdx.y, bt = blockIdx.z;
    const u32 N = a.T.N;
    const LimbConst c = load_limb_const(a.T, l);
    u64* p = a.c + bt * a.c_bs + (size_t)l * N;
    const u64 one = mform(1, c.q, c.u0, c.u1);
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        if (j == 0) continue;
        u64 gal = one, x = a.s[l];
        for (u32 e = j; e > 0; e >>= 1) {
            if (e & 1) gal = mred(gal, x, c.q, c.qinv);
            x = mred(x, x, c.q, c.qinv);
        }
        p[j] = mred(p[j], gal, c.q, c.qinv);
    }
}

// Equal, ring_context.go:424-467, after both operands were reduced in place: *flag |= 1 when some word differs
__global__ void __launch_bounds__(256) differ_kernel(const ExtArgs a, u32* flag) {
    const int l = blockIdx.y, bt = blockIdx.z;
    const u32 N = a.T.N;
    const u64* x = a.a + bt * a.a_bs + (size_t)l * N;
    const u64* y = a.b + bt * a.b_bs + (size_t)l * N;
    int bad = 0;
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) bad |= x[k] != y[k];
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1u);
}

// MulPolyNaive(Montgomery), ring.go:383-437: for every i in order, p3[j] = CRed(p3[j] + (q - MRed(p1[i], p2[N-i+j]))) for
// j < i and CRed(p3[j] + MRed(p1[i], p2[j-i])) for j >= i, from p3 = 0.  One thread walks the i-sequence of its own j.
__global__ void __launch_bounds__(256) mulpoly_naive_kernel(const ExtArgs a) {
    const int l = blockIdx.y, bt = blockIdx.z;
    const u32 N = a.T.N;
    const LimbConst c = load_limb_const(a.T, l);
    const u64* p1 = a.a + bt * a.a_bs + (size_t)l * N;
    const u64* p2 = a.b + bt * a.b_bs + (size_t)l * N;
    u64* p3 = a.c + bt * a.c_bs + (size_t)l * N;
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        u64 acc = 0;
        for (u32 i = 0; i <= j; ++i) acc = cred(acc + mred(p1[i], p2[j - i], c.q, c.qinv), c.q);
        for (u32 i = j + 1; i < N; ++i) acc = cred(acc + (c.q - mred(p1[i], p2[N - i + j], c.q, c.qinv)), c.q);
        p3[j] = acc;
    }
}

__global__ void __launch_bounds__(256) fill_kernel(u64* p, size_t bs, u32 N, u64 v) {
    u64* o = p + blockIdx.z * bs + (size_t)blockIdx.y * N;
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) o[k] = v;
}

dim3 ext_grid(u32 N, int nl, int batch) {
    u32 bx = (N + 255) / 256;
    if (bx > 64) bx = 64;
    return dim3(bx, nl, batch);
}

int check3(const lg_ring* r, int nl, const lg_poly* p, const char* what) {
    LG_REQUIRE(p, "%s: null polynomial", what);
    LG_REQUIRE(p->N == r->N, "%s: polynomial degree %llu != ring degree %llu", what, (unsigned long long)p->N, (unsigned long long)r->N);
    LG_REQUIRE(nl <= p->nlimbs, "%s: %d limbs requested, polynomial has %d", what, nl, p->nlimbs);
    LG_SAME_DEVICE(what, r->device, p->device);
    return LG_OK;
}

}  // namespace

extern "C" {

int lg_ring_mul_poly(const lg_ring* r, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, int montgomery, lg_stream_t s) {
    LG_REQUIRE(r, "MulPoly: null ring");
    const int nl = r->nl;
    LG_TRY(check3(r, nl, p1, "MulPoly"));
    LG_TRY(check3(r, nl, p2, "MulPoly"));
    LG_TRY(check3(r, nl, p3, "MulPoly"));
    LG_REQUIRE(p1->batch == p3->batch && p2->batch == p3->batch, "MulPoly: batch mismatch");
    LG_ON_DEVICE(r->device);
    cudaStream_t st = (cudaStream_t)s;
    const int batch = p3->batch;
    const size_t bs = (size_t)nl * r->N;
    Scratch ab(st);  // a := context.NewPoly(), b := context.NewPoly()
    LG_TRY(ab.alloc(2 * batch * bs));
    u64 *a = ab.d, *b = ab.d + batch * bs;
    const LimbMap id = limb_map_identity();
    LG_TRY(lgi_ntt(r, id, nl, batch, p1->d, p1->bstride, a, bs, false, 0, 0, st));
    LG_TRY(lgi_ntt(r, id, nl, batch, p2->d, p2->bstride, b, bs, false, 0, 0, st));
    LG_TRY(lgi_ew(montgomery ? EW_MULMONT : EW_MUL_BARRETT, r, id, nl, batch, a, bs, b, bs, p3->d, p3->bstride, nullptr, 0, st));
    return lgi_ntt(r, id, nl, batch, p3->d, p3->bstride, p3->d, p3->bstride, true, 0, 0, st);
}

int lg_ring_mul_poly_naive(const lg_ring* r, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, int montgomery, lg_stream_t s) {
    LG_REQUIRE(r, "MulPolyNaive: null ring");
    const int nl = r->nl;
    LG_TRY(check3(r, nl, p1, "MulPolyNaive"));
    LG_TRY(check3(r, nl, p2, "MulPolyNaive"));
    LG_TRY(check3(r, nl, p3, "MulPolyNaive"));
    LG_REQUIRE(p1->batch == p3->batch && p2->batch == p3->batch, "MulPolyNaive: batch mismatch");
    LG_ON_DEVICE(r->device);
    cudaStream_t st = (cudaStream_t)s;
    const int batch = p3->batch;
    const size_t bs = (size_t)nl * r->N;
    // p1Copy, p2Copy (:385-386, :415-416): the result may alias an operand
    Scratch cp(st);
    LG_TRY(cp.alloc(2 * batch * bs));
    u64 *c1 = cp.d, *c2 = cp.d + batch * bs;
    const LimbMap id = limb_map_identity();
    LG_TRY(lgi_ew(montgomery ? EW_COPY : EW_MFORM, r, id, nl, batch, p1->d, p1->bstride, nullptr, 0, c1, bs, nullptr, 0, st));
    LG_TRY(lgi_ew(EW_COPY, r, id, nl, batch, p2->d, p2->bstride, nullptr, 0, c2, bs, nullptr, 0, st));
    ExtArgs a;
    memset(&a, 0, sizeof(a));
    a.T = r->T;
    a.a = c1;
    a.b = c2;
    a.c = p3->d;
    a.a_bs = a.b_bs = bs;
    a.c_bs = p3->bstride;
    mulpoly_naive_kernel<<<ext_grid((u32)r->N, nl, batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_ring_exp(const lg_ring* r, lg_poly* p1, uint64_t e, lg_poly* p2, lg_stream_t s) {
    LG_REQUIRE(r, "Exp: null ring");
    const int nl = r->nl;
    LG_TRY(check3(r, nl, p1, "Exp"));
    LG_TRY(check3(r, nl, p2, "Exp"));
    LG_REQUIRE(p1->batch == p2->batch, "Exp: batch mismatch");
    LG_ON_DEVICE(r->device);
    cudaStream_t st = (cudaStream_t)s;
    const int batch = p2->batch;
    const size_t bs = (size_t)nl * r->N;
    const LimbMap id = limb_map_identity();
    LG_TRY(lgi_ntt(r, id, nl, batch, p1->d, p1->bstride, p1->d, p1->bstride, false, 0, 0, st));  // :443
    Scratch tmp(st);
    LG_TRY(tmp.alloc(batch * bs));
    LG_CUDA_CHECK(cudaMemsetAsync(tmp.d, 0, batch * bs * sizeof(u64), st));  // :445
    LG_TRY(lgi_ew(EW_ADD, r, id, nl, batch, tmp.d, bs, p1->d, p1->bstride, tmp.d, bs, nullptr, 0, st));  // :446
    fill_kernel<<<ext_grid((u32)r->N, nl, batch), 256, 0, st>>>(p2->d, p2->bstride, (u32)r->N, 1ull);  // :448-453
    lg_g_launches += 1;
    LG_LAUNCH_CHECK();
    for (uint64_t i = e; i > 0; i >>= 1) {  // :455-460
        if (i & 1) LG_TRY(lgi_ew(EW_MUL_BARRETT, r, id, nl, batch, p2->d, p2->bstride, tmp.d, bs, p2->d, p2->bstride, nullptr, 0, st));
        LG_TRY(lgi_ew(EW_MUL_BARRETT, r, id, nl, batch, tmp.d, bs, p1->d, p1->bstride, tmp.d, bs, nullptr, 0, st));
    }
    LG_TRY(lgi_ntt(r, id, nl, batch, p2->d, p2->bstride, p2->d, p2->bstride, true, 0, 0, st));     // :462
    return lgi_ntt(r, id, nl, batch, p1->d, p1->bstride, p2->d, p2->bstride, true, 0, 0, st);       // :463
}

int lg_ring_shift(const lg_ring* r, const lg_poly* p1, uint64_t n, lg_poly* p2, lg_stream_t s) {
    LG_REQUIRE(r, "Shift: null ring");
    const int nl = r->nl;
    LG_TRY(check3(r, nl, p1, "Shift"));
    LG_TRY(check3(r, nl, p2, "Shift"));
    LG_REQUIRE(p1->batch == p2->batch, "Shift: batch mismatch");
    const u64 k = n & go_mask(r->N);
    LG_REQUIRE(k <= r->N, "Shift: slice bounds out of range [%llu:%llu]", (unsigned long long)k, (unsigned long long)r->N);
    LG_ON_DEVICE(r->device);
    cudaStream_t st = (cudaStream_t)s;
    const int batch = p2->batch;
    const size_t bs = (size_t)nl * r->N;
    ExtArgs a;
    memset(&a, 0, sizeof(a));
    a.T = r->T;
    a.a = p1->d;
    a.a_bs = p1->bstride;
    Scratch cp(st);
    if (p1->d == p2->d) {  // append() builds a new slice: the source is read before anything is written
        LG_TRY(cp.alloc(batch * bs));
        LG_TRY(lgi_ew(EW_COPY, r, limb_map_identity(), nl, batch, p1->d, p1->bstride, nullptr, 0, cp.d, bs, nullptr, 0, st));
        a.a = cp.d;
        a.a_bs = bs;
    }
    a.c = p2->d;
    a.c_bs = p2->bstride;
    a.n = k == r->N ? 0 : k;
    shift_kernel<<<ext_grid((u32)r->N, nl, batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_ring_rotate(const lg_ring* r, lg_poly* p1, uint64_t n, lg_stream_t s) {
    LG_REQUIRE(r, "Rotate: null ring");
    const int nl = r->nl;
    LG_TRY(check3(r, nl, p1, "Rotate"));
    LG_ON_DEVICE(r->device);
    cudaStream_t st = (cudaStream_t)s;
    n &= go_mask(r->N);  // :779
    ExtArgs a;
    memset(&a, 0, sizeof(a));
    a.T = r->T;
    a.c = p1->d;  // p1tmp, p2tmp := p1.Coeffs[i], p1.Coeffs[i] (:791)
    a.c_bs = p1->bstride;
    for (int i = 0; i < nl; ++i) {
        const u64 q = r->q[i], qinv = r->mred[i];
        const u64 psi_mont = r->psi[(size_t)i * r->N + (r->N >> 1)];  // nttPsi[bitrev(1)] = psiMont (ring_context.go:185-200)
        u64 root = lgh::mred(psi_mont, psi_mont, q, qinv);  // :785
        u64 res = lgh::mform(1, q);                         // modexpMontgomery, ring/utils.go:39-50
        for (u64 e = n; e > 0; e >>= 1) {
            if (e & 1) res = lgh::mred(res, root, q, qinv);
            root = lgh::mred(root, root, q, qinv);
        }
        a.s[i] = res;
    }
    rotate_kernel<<<ext_grid((u32)r->N, nl, p1->batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int lg_ring_equal(const lg_ring* r, int nl, lg_poly* p1, lg_poly* p2, int* equal, lg_stream_t s) {
    LG_REQUIRE(r && equal, "Equal: null argument");
    LG_REQUIRE(nl >= 1 && nl <= r->nl, "Equal: %d limbs requested, ring has %d", nl, r->nl);
    LG_TRY(check3(r, nl, p1, "Equal"));
    LG_TRY(check3(r, nl, p2, "Equal"));
    LG_REQUIRE(p1->batch == p2->batch, "Equal: batch mismatch");
    LG_ON_DEVICE(r->device);
    cudaStream_t st = (cudaStream_t)s;
    const int batch = p1->batch;
    const LimbMap id = limb_map_identity();
    // :432-433 / :455-456: both operands are reduced in place first
    LG_TRY(lgi_ew(EW_REDUCE, r, id, nl, batch, p1->d, p1->bstride, nullptr, 0, p1->d, p1->bstride, nullptr, 0, st));
    LG_TRY(lgi_ew(EW_REDUCE, r, id, nl, batch, p2->d, p2->bstride, nullptr, 0, p2->d, p2->bstride, nullptr, 0, st));
    Scratch flag(st);
    LG_TRY(flag.alloc(1));
    LG_CUDA_CHECK(cudaMemsetAsync(flag.d, 0, sizeof(u64), st));
    ExtArgs a;
    memset(&a, 0, sizeof(a));
    a.T = r->T;
    a.a = p1->d;
    a.a_bs = p1->bstride;
    a.b = p2->d;
    a.b_bs = p2->bstride;
    differ_kernel<<<ext_grid((u32)r->N, nl, batch), 256, 0, st>>>(a, (u32*)flag.d);
    lg_g_launches += 1;
    LG_LAUNCH_CHECK();
    u32 h = 0;
    LG_CUDA_CHECK(cudaMemcpyAsync(&h, flag.d, sizeof(u32), cudaMemcpyDeviceToHost, st));
    LG_CUDA_CHECK(cudaStreamSynchronize(st));  // the answer is a host bool
    *equal = h ? 0 : 1;
    return LG_OK;
}

}  // extern "C"
