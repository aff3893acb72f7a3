// scaler.cu -- ring.SimpleScaler (ring/ring_scaling.go:166-300) and the plaintext lift of the BFV encoder
// (bfv/encoder.go:121-136, bfv/utils.go:9-23) on the device, with their C-ABI entry points.
//
// Scale reconstructs round(t/Q * x) mod t from the RNS residues of x without big integers: an integer part
// sum_j MRed_t(w_j, x_j) and a fractional part accumulated in "Float128" double-double arithmetic
// (ring/float128.go, after libqd).  The result depends on the exact sequence of IEEE binary64 roundings, so the
// device code spells every operation with the round-to-nearest intrinsics (__dadd_rn / __dmul_rn / __ddiv_rn are
// never contracted into FMAs) in the order float128.go writes them; the parameters are generated on the host by
// the same sequence (x86-64 baseline code has no fused multiply-add).
#include <math.h>
#include <string.h>

#include "capi_internal.hpp"

static inline cudaStream_t cs(lg_stream_t s) { return (cudaStream_t)s; }

// ------------------------------------------------------------------------------------------------------
// double-double, written once for host and device
// ------------------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define F_ADD(a, b) __dadd_rn((a), (b))
#define F_SUB(a, b) __dsub_rn((a), (b))
#define F_MUL(a, b) __dmul_rn((a), (b))
#define F_DIV(a, b) __ddiv_rn((a), (b))
#else
#define F_ADD(a, b) ((a) + (b))
#define F_SUB(a, b) ((a) - (b))
#define F_MUL(a, b) ((a) * (b))
#define F_DIV(a, b) ((a) / (b))
#endif
#define LG_HD __host__ __device__ __forceinline__

struct f128 {
    double hi, lo;
};

// Go's uint64(float64) on amd64: CVTTSD2SQ below 2^63 (negatives wrap as two's complement), else the same conversion of
// f - 2^63 with the top bit flipped.  CVTTSD2SQ answers 0x8000000000000000 ("integer indefinite") out of range, which
// decides what out-of-contract inputs (residues far above their modulus) produce; cvt.rzi.s64.f64 would saturate instead.
LG_HD long long x86_cvttsd2sq(double f) {
    return (f >= -9223372036854775808.0 && f < 9223372036854775808.0) ? (long long)f : (long long)0x8000000000000000ull;
}
LG_HD u64 go_f64_to_u64(double f) {
    if (f < 9223372036854775808.0) return (u64)x86_cvttsd2sq(f);
    return (u64)x86_cvttsd2sq(F_SUB(f, 9223372036854775808.0)) ^ 0x8000000000000000ull;
}
LG_HD double go_round(double x) {  // math.Round, half away from zero
    if (!(x > -4503599627370496.0 && x < 4503599627370496.0)) return x;  // already integral (or NaN)
    double t = (double)(long long)x;
    const double d = F_SUB(x, t);
    if (d >= 0.5) t = F_ADD(t, 1.0);
    else if (d <= -0.5) t = F_SUB(t, 1.0);
    return t;
}
LG_HD f128 f128_set_u53(u64 i) { return f128{(double)i, 0.0}; }                              // float128.go:33-37
LG_HD f128 f128_set_u64(u64 i) { return f128{(double)(i >> 12), F_DIV((double)(i & 0xfff), 4096.0)}; }  // :44-48
LG_HD u64 f128_to_u53(f128 f) { return go_f64_to_u64(f.hi); }                                // :72-74
LG_HD u64 f128_to_u64(f128 f) {                                                               // :80-82
    const double a = F_MUL(f.hi, 4096.0);
    const u64 ai = go_f64_to_u64(a);
    return ai + go_f64_to_u64(go_round(F_ADD(F_SUB(a, (double)ai), F_MUL(f.lo, 4096.0))));
}
LG_HD void two_sum(double a, double b, double& s, double& e) {  // :85-90
    s = F_ADD(a, b);
    const double bb = F_SUB(s, a);
    e = F_ADD(F_SUB(a, F_SUB(s, bb)), F_SUB(b, bb));
}
LG_HD void quick_two_sum(double a, double b, double& s, double& e) {  // :93-97
    const double t = F_ADD(a, b);
    e = F_SUB(b, F_SUB(t, a));
    s = t;
}
LG_HD void two_diff(double a, double b, double& s, double& e) {  // :111-116
    s = F_SUB(a, b);
    const double bb = F_SUB(s, a);
    e = F_SUB(F_SUB(a, F_SUB(s, bb)), F_ADD(b, bb));
}
LG_HD f128 f128_add(f128 a, f128 b) {  // :100-108
    double s1, s2, t1, t2;
    two_sum(a.hi, b.hi, s1, s2);
    two_sum(a.lo, b.lo, t1, t2);
    s2 = F_ADD(s2, t1);
    quick_two_sum(s1, s2, s1, s2);
    s2 = F_ADD(s2, t2);
    f128 f;
    quick_two_sum(s1, s2, f.hi, f.lo);
    return f;
}
LG_HD void f_split(double a, double& hi, double& lo) {  // :132-137
    const double temp = F_MUL(134217729.0, a);
    hi = F_SUB(temp, F_SUB(temp, a));
    lo = F_SUB(a, hi);
}
LG_HD void two_prod(double a, double b, double& p, double& e) {  // :140-146
    p = F_MUL(a, b);
    double ah, al, bh, bl;
    f_split(a, ah, al);
    f_split(b, bh, bl);
    e = F_ADD(F_ADD(F_ADD(F_SUB(F_MUL(ah, bh), p), F_MUL(ah, bl)), F_MUL(al, bh)), F_MUL(al, bl));
}
LG_HD f128 f128_mul(f128 a, f128 b) {  // :149-154
    double p1, p2;
    two_prod(a.hi, b.hi, p1, p2);
    p2 = F_ADD(p2, F_ADD(F_MUL(a.hi, b.lo), F_MUL(a.lo, b.hi)));
    f128 f;
    quick_two_sum(p1, p2, f.hi, f.lo);
    return f;
}
LG_HD f128 f128_div(f128 a, f128 b) {  // :157-217, the live statements
    double p1, p2, p3, p4, v1, v2;
    const double q1 = F_DIV(a.hi, b.hi);
    two_prod(q1, b.hi, p1, p2);
    p2 = F_ADD(p2, F_MUL(q1, b.lo));
    const double t0 = F_ADD(p1, p2);
    const double t1 = F_SUB(p2, F_SUB(t0, p1));
    two_diff(a.hi, t0, p3, p4);
    two_diff(a.lo, t1, v1, v2);
    p4 = F_ADD(p4, v1);
    quick_two_sum(p3, p4, p3, p4);
    p4 = F_ADD(p4, v2);
    const double r = F_DIV(F_ADD(p3, p4), b.hi);
    f128 f;
    f.hi = F_ADD(q1, r);
    f.lo = F_SUB(r, F_SUB(f.hi, q1));
    return f;
}

// ------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------
struct ScaleArgs {
    const u64* in;
    u64* out;
    size_t in_bs, out_bs;
    const u64* wi;     // [nl]  integer parts (Montgomery form mod t unless t is a power of two)
    const double* ti;  // [nl][2] fractional parts
    u64 t, add_param, mul_param;
    u32 N;
    int nl, nl_out, pow2;
};

constexpr int kScaleMaxLimbs = 64;

// SimpleScaler.Scale, ring_scaling.go:271-300.  One coefficient per thread: nl coalesced limb reads, nl_out writes.
// All reads of a column precede its writes, so p2 may be p1 (ring_test.go:614 scales in place).
__global__ void __launch_bounds__(128) simple_scale_kernel(const ScaleArgs a) {
    __shared__ u64 s_wi[kScaleMaxLimbs];
    __shared__ double s_ti[2 * kScaleMaxLimbs];
    for (int j = threadIdx.x; j < a.nl; j += blockDim.x) {
        s_wi[j] = a.wi[j];
        s_ti[2 * j] = a.ti[2 * j];
        s_ti[2 * j + 1] = a.ti[2 * j + 1];
    }
    __syncthreads();
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const u64* in = a.in + blockIdx.y * a.in_bs + i;
    u64 acc = 0;
    f128 b{0.0, 0.0};
    for (int j = 0; j < a.nl; ++j) {
        const u64 x = in[(size_t)j * a.N];
        if (a.pow2) acc += (s_wi[j] * x) & a.add_param;                       // :206-208
        else acc += mred(s_wi[j], x, a.t, a.mul_param);                       // :219-230
        b = f128_add(b, f128_mul(f128{s_ti[2 * j], s_ti[2 * j + 1]}, f128_set_u64(x)));  // :288
    }
    acc += f128_to_u64(b);                                                    // :291
    if (a.pow2) {
        acc &= a.mul_param;                                                   // :210-212
    } else {                                                                  // :232-243
        const u64 s0 = __umul64hi(acc, a.add_param);
        u64 r = acc - s0 * a.t;
        if (r >= a.t) r -= a.t;
        acc = r;
    }
    u64* out = a.out + blockIdx.y * a.out_bs + i;
    for (int j = 0; j < a.nl_out; ++j) out[(size_t)j * a.N] = acc;           // :295-297
}

struct LiftArgs {
    const u64* m;  // one limb of values below t
    u64* out;
    size_t m_bs, out_bs;
    const u64* delta;  // [nl] deltaMont
    const u64* q;
    const u64* qinv;
    u32 N;
    int nl;
};

// encodePlaintext, bfv/encoder.go:125-135: limb i = MRed(limb 0, deltaMont[i]) for i = nl-1 .. 0 (limb 0 last,
// so the message may sit in limb 0 of the output like the reference's plaintext)
__global__ void __launch_bounds__(256) bfv_lift_kernel(const LiftArgs a) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const u64 m = a.m[blockIdx.y * a.m_bs + i];
    u64* out = a.out + blockIdx.y * a.out_bs + i;
    for (int j = a.nl - 1; j >= 0; --j) out[(size_t)j * a.N] = mred(m, __ldg(a.delta + j), __ldg(a.q + j), __ldg(a.qinv + j));
}

// ------------------------------------------------------------------------------------------------------
// host objects + C ABI
// ------------------------------------------------------------------------------------------------------
struct lg_scaler {
    const lg_ring* ring = nullptr;
    u64 t = 0, add_param = 0, mul_param = 0;
    int pow2 = 0;
    std::vector<u64> wi;
    std::vector<double> ti;
    DevArray<u64> d_wi;
    DevArray<double> d_ti;
};

struct lg_bfv_lift {
    const lg_ring* ring = nullptr;
    std::vector<u64> delta;
    DevArray<u64> d_delta;
};

extern "C" {

// NewSimpleScaler, ring_scaling.go:188-262 -- host only: the parameters for a modulus list
int lg_scaler_params_host(uint64_t t, const uint64_t* moduli, int nl, uint64_t* wi, double* ti, uint64_t* add_param,
                          uint64_t* mul_param) {
    LG_REQUIRE(moduli && wi && ti && nl >= 1, "NewSimpleScaler: bad argument");
    LG_REQUIRE(t >= 2, "NewSimpleScaler: plaintext modulus must be at least 2");
    const bool pow2 = (t & (t - 1)) == 0;
    u64 bhi = 0, blo = 0, ap, mp;
    if (pow2) {
        ap = mp = t - 1;  // :203-204
    } else {
        lgh::bred_params(t, bhi, blo);
        ap = bhi;                    // :216
        mp = lgh::mred_params(t);    // :217
    }
    if (add_param) *add_param = ap;
    if (mul_param) *mul_param = mp;
    for (int i = 0; i < nl; ++i) {
        const u64 qi = moduli[i];
        u64 star = 1 % qi;  // Q/qi mod qi
        for (int k = 0; k < nl; ++k)
            if (k != i) star = lgh::mulmod(star, moduli[k] % qi, qi);
        const u64 barre = lgh::powmod(star, qi - 2, qi);  // (Q/qi)^-1 mod qi (:241-247), qi prime
        f128 tmp = f128_div(f128_set_u53(t), f128_set_u64(qi));  // :249
        tmp = f128_mul(tmp, f128_set_u64(barre));                // :251
        u64 w = f128_to_u53(tmp);                                // :254
        if (!pow2) {                                             // :257-259 MForm(w, t, BRedParams(t))
            const u64 mhi = lgh::mulhi(w, blo);
            u64 r = (0 - (w * bhi + mhi)) * t;
            if (r >= t) r -= t;
            w = r;
        }
        wi[i] = w;
        const u64 barre_t = lgh::mulmod(barre, t % qi, qi);      // :261-262
        const f128 f = f128_div(f128_set_u64(barre_t), f128_set_u64(qi));  // :264
        ti[2 * i] = f.hi;
        ti[2 * i + 1] = f.lo;
    }
    return LG_OK;
}

int lg_scaler_create(uint64_t t, const lg_ring* ring, lg_scaler** out) {
    LG_REQUIRE(ring && out, "NewSimpleScaler: null argument");
    LG_REQUIRE(ring->nl <= kScaleMaxLimbs, "NewSimpleScaler: at most %d moduli", kScaleMaxLimbs);
    LG_ON_DEVICE(ring->device);
    std::unique_ptr<lg_scaler> s(new lg_scaler);
    s->ring = ring;
    s->t = t;
    s->pow2 = t >= 2 && (t & (t - 1)) == 0;
    s->wi.resize(ring->nl);
    s->ti.resize(2 * ring->nl);
    LG_TRY(lg_scaler_params_host(t, ring->q.data(), ring->nl, s->wi.data(), s->ti.data(), &s->add_param, &s->mul_param));
    LG_TRY(s->d_wi.upload(s->wi));
    LG_TRY(s->d_ti.upload(s->ti));
    *out = s.release();
    return LG_OK;
}

int lg_scaler_destroy(lg_scaler* s) {
    if (!s) return LG_OK;
    LG_ON_DEVICE(s->ring->device);
    delete s;
    return LG_OK;
}

int lg_scaler_get_params(const lg_scaler* s, uint64_t* wi, double* ti) {
    LG_REQUIRE(s, "SimpleScaler: null handle");
    if (wi) memcpy(wi, s->wi.data(), sizeof(u64) * s->wi.size());
    if (ti) memcpy(ti, s->ti.data(), sizeof(double) * s->ti.size());
    return LG_OK;
}

// SimpleScaler.Scale(p1, p2), ring_scaling.go:271-300
int lg_scaler_scale(const lg_scaler* s, const lg_poly* p1, lg_poly* p2, lg_stream_t stream) {
    LG_REQUIRE(s && p1 && p2, "Scale: null argument");
    const lg_ring* r = s->ring;
    LG_REQUIRE(p1->N == r->N && p2->N == r->N, "Scale: degree mismatch");
    LG_REQUIRE(p1->nlimbs >= r->nl, "Scale: input has %d limbs, the context has %d", p1->nlimbs, r->nl);
    LG_REQUIRE(p1->batch == p2->batch, "Scale: batch mismatch");
    LG_SAME_DEVICE("Scale", r->device, p1->device);
    LG_SAME_DEVICE("Scale", r->device, p2->device);
    LG_ON_DEVICE(r->device);
    ScaleArgs a;
    a.in = p1->d;
    a.out = p2->d;
    a.in_bs = p1->bstride;
    a.out_bs = p2->bstride;
    a.wi = s->d_wi.d;
    a.ti = s->d_ti.d;
    a.t = s->t;
    a.add_param = s->add_param;
    a.mul_param = s->mul_param;
    a.N = (u32)r->N;
    a.nl = r->nl;
    a.nl_out = p2->nlimbs;
    a.pow2 = s->pow2;
    simple_scale_kernel<<<dim3((unsigned)((r->N + 127) / 128), p1->batch), 128, 0, cs(stream)>>>(a);
    lg_g_launches += 1;
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// GenLiftParams, bfv/utils.go:9-23: deltaMont[i] = MForm(floor(Q/t) mod q_i) -- host only
int lg_bfv_lift_params_host(const uint64_t* moduli, int nl, uint64_t t, uint64_t* delta_mont) {
    LG_REQUIRE(moduli && delta_mont && nl >= 1 && t >= 2, "GenLiftParams: bad argument");
    // Q as a multi-word integer, then floor(Q / t) by schoolbook division (most significant word first)
    std::vector<u64> w{1};
    for (int i = 0; i < nl; ++i) {
        u64 carry = 0;
        for (auto& x : w) {
            const unsigned __int128 p = (unsigned __int128)x * moduli[i] + carry;
            x = (u64)p;
            carry = (u64)(p >> 64);
        }
        if (carry) w.push_back(carry);
    }
    unsigned __int128 rem = 0;
    for (size_t i = w.size(); i-- > 0;) {
        const unsigned __int128 cur = (rem << 64) | w[i];
        w[i] = (u64)(cur / t);
        rem = cur % t;
    }
    for (int i = 0; i < nl; ++i) {
        const u64 q = moduli[i];
        unsigned __int128 r = 0;
        for (size_t k = w.size(); k-- > 0;) r = ((r << 64) | w[k]) % q;
        delta_mont[i] = lgh::mform((u64)r, q);
    }
    return LG_OK;
}

int lg_bfv_lift_create(const lg_ring* ringQ, uint64_t t, lg_bfv_lift** out) {
    LG_REQUIRE(ringQ && out, "GenLiftParams: null argument");
    LG_ON_DEVICE(ringQ->device);
    std::unique_ptr<lg_bfv_lift> l(new lg_bfv_lift);
    l->ring = ringQ;
    l->delta.resize(ringQ->nl);
    LG_TRY(lg_bfv_lift_params_host(ringQ->q.data(), ringQ->nl, t, l->delta.data()));
    LG_TRY(l->d_delta.upload(l->delta));
    *out = l.release();
    return LG_OK;
}

int lg_bfv_lift_destroy(lg_bfv_lift* l) {
    if (!l) return LG_OK;
    LG_ON_DEVICE(l->ring->device);
    delete l;
    return LG_OK;
}

int lg_bfv_lift_get_params(const lg_bfv_lift* l, uint64_t* delta_mont) {
    LG_REQUIRE(l && delta_mont, "GenLiftParams: null argument");
    memcpy(delta_mont, l->delta.data(), sizeof(u64) * l->delta.size());
    return LG_OK;
}

// encodePlaintext, bfv/encoder.go:121-136 after the InvNTT over contextT: m (one limb, values below t) -> plaintext
int lg_bfv_lift_apply(const lg_bfv_lift* l, const lg_poly* m, lg_poly* pt, lg_stream_t stream) {
    LG_REQUIRE(l && m && pt, "encodePlaintext: null argument");
    const lg_ring* r = l->ring;
    LG_REQUIRE(m->N == r->N && pt->N == r->N, "encodePlaintext: degree mismatch");
    LG_REQUIRE(pt->nlimbs >= r->nl, "encodePlaintext: plaintext has %d limbs, %d needed", pt->nlimbs, r->nl);
    LG_REQUIRE(m->batch == pt->batch, "encodePlaintext: batch mismatch");
    LG_SAME_DEVICE("encodePlaintext", r->device, m->device);
    LG_SAME_DEVICE("encodePlaintext", r->device, pt->device);
    LG_ON_DEVICE(r->device);
    LiftArgs a;
    a.m = m->d;
    a.out = pt->d;
    a.m_bs = m->bstride;
    a.out_bs = pt->bstride;
    a.delta = l->d_delta.d;
    a.q = r->T.q;
    a.qinv = r->T.qinv;
    a.N = (u32)r->N;
    a.nl = r->nl;
    bfv_lift_kernel<<<dim3((unsigned)((r->N + 255) / 256), m->batch), 256, 0, cs(stream)>>>(a);
    lg_g_launches += 1;
    LG_LAUNCH_CHECK();
    return LG_OK;
}

}  // extern "C"
