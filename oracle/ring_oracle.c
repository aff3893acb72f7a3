/*
 * oracle/ring_oracle.c  --  TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement of the RNS-ring hot path of Lattigo v1.3.1 (the
 * reference mounted at /root/reference).  It exists so that the CUDA path can
 * be compared bit-for-bit against the reference's algorithm; it is NOT part
 * of the product and nothing under lattigo-fhe-by-go_b200/ links, imports or
 * executes it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED.  The NTT / table-generation part reproduces every
 * value of the reference's seven golden (input, NTT(input)) vector pairs in
 * ring/test_data (tests/test_oracle_golden.py); the rest is anchored by the
 * reference's own big-integer properties (ring/ring_test.go), restated in
 * tests/test_oracle_properties.py.  The reference cannot be run here (no Go
 * toolchain in the image), so there is no oracle/_ref build.
 *
 * Every function cites the reference file:line it follows.
 * All polynomials are flat, limb-major arrays: limb i = p[i*N .. i*N+N).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

#define API __attribute__((visibility("default")))

static inline u64 hi64(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

/* ---------------------------------------------------------------------- */
/* ring/modular_reduction.go                                              */
/* ---------------------------------------------------------------------- */

/* BRedParams, modular_reduction.go:97-106: floor(2^128/q) as {hi, lo}. */
API void orc_bred_params(u64 q, u64 u[2]) {
    u128 one64 = (u128)1 << 64;
    u64 hi = (u64)(one64 / q);
    u64 rem = (u64)(one64 % q);
    u64 lo = (u64)((((u128)rem) << 64) / q);
    u[0] = hi;
    u[1] = lo;
}

/* MRedParams, modular_reduction.go:53-64: q^(2^63-1) = q^-1 mod 2^64. */
API u64 orc_mred_params(u64 q) {
    u64 qinv = 1, x = q;
    for (int i = 0; i < 63; i++) {
        qinv *= x;
        x *= x;
    }
    return qinv;
}

/* MForm, modular_reduction.go:15-22 */
API u64 orc_mform(u64 a, u64 q, const u64 u[2]) {
    u64 mhi = hi64(a, u[1]);
    u64 r = (0 - (a * u[0] + mhi)) * q;
    if (r >= q) r -= q;
    return r;
}

/* MFormConstant, modular_reduction.go:26-30 */
API u64 orc_mform_constant(u64 a, u64 q, const u64 u[2]) {
    u64 mhi = hi64(a, u[1]);
    return (0 - (a * u[0] + mhi)) * q;
}

/* InvMForm, modular_reduction.go:34-41 */
API u64 orc_invmform(u64 a, u64 q, u64 qinv) {
    u64 r = hi64(a * qinv, q);
    r = q - r;
    if (r >= q) r -= q;
    return r;
}

/* InvMFormConstant, modular_reduction.go:45-49 */
API u64 orc_invmform_constant(u64 a, u64 q, u64 qinv) {
    u64 r = hi64(a * qinv, q);
    return q - r;
}

/* MRed, modular_reduction.go:70-79 */
API u64 orc_mred(u64 x, u64 y, u64 q, u64 qinv) {
    u128 a = (u128)x * y;
    u64 ahi = (u64)(a >> 64), alo = (u64)a;
    u64 R = alo * qinv;
    u64 H = hi64(R, q);
    u64 r = ahi - H + q;
    if (r >= q) r -= q;
    return r;
}

/* MRedConstant, modular_reduction.go:83-89 */
API u64 orc_mred_constant(u64 x, u64 y, u64 q, u64 qinv) {
    u128 a = (u128)x * y;
    u64 ahi = (u64)(a >> 64), alo = (u64)a;
    u64 R = alo * qinv;
    u64 H = hi64(R, q);
    return ahi - H + q;
}

/* BRedAdd, modular_reduction.go:112-119 */
API u64 orc_bred_add(u64 x, u64 q, const u64 u[2]) {
    u64 s0 = hi64(x, u[0]);
    u64 r = x - s0 * q;
    if (r >= q) r -= q;
    return r;
}

/* BRedAddConstant, modular_reduction.go:123-126 */
API u64 orc_bred_add_constant(u64 x, u64 q, const u64 u[2]) {
    u64 s0 = hi64(x, u[0]);
    return x - s0 * q;
}

/* BRedConstant, modular_reduction.go:172-207 (and the body of BRed :133-168) */
API u64 orc_bred_constant(u64 x, u64 y, u64 q, const u64 u[2]) {
    u64 lhi, mhi, mlo, s0, s1, carry;
    u128 a = (u128)x * y;
    u64 ahi = (u64)(a >> 64), alo = (u64)a;
    lhi = hi64(alo, u[1]);
    u128 m = (u128)alo * u[0];
    mhi = (u64)(m >> 64);
    mlo = (u64)m;
    s0 = mlo + lhi;
    carry = s0 < mlo;
    s1 = mhi + carry;
    m = (u128)ahi * u[1];
    mhi = (u64)(m >> 64);
    mlo = (u64)m;
    u64 t = mlo + s0;
    carry = t < mlo;
    lhi = mhi + carry;
    s0 = ahi * u[0] + s1 + lhi;
    return alo - s0 * q;
}

/* BRed, modular_reduction.go:133-168 */
API u64 orc_bred(u64 x, u64 y, u64 q, const u64 u[2]) {
    u64 r = orc_bred_constant(x, y, q, u);
    if (r >= q) r -= q;
    return r;
}

/* CRed, modular_reduction.go:211-216 */
API u64 orc_cred(u64 a, u64 q) { return a >= q ? a - q : a; }

/* ---------------------------------------------------------------------- */
/* ring/utils.go, utils/utils.go                                          */
/* ---------------------------------------------------------------------- */

/* PowerOf2, ring/utils.go:8-17 */
API u64 orc_power_of_2(u64 x, u64 n, u64 q, u64 qinv) {
    u64 ahi = n ? x >> (64 - n) : 0, alo = x << n; /* Go: x>>64 == 0 */
    u64 R = alo * qinv;
    u64 H = hi64(R, q);
    u64 r = ahi - H + q;
    if (r >= q) r -= q;
    return r;
}

/* ModExp, ring/utils.go:25-35 */
API u64 orc_modexp(u64 x, u64 e, u64 p) {
    u64 params[2];
    orc_bred_params(p, params);
    u64 result = 1;
    for (u64 i = e; i > 0; i >>= 1) {
        if (i & 1) result = orc_bred(result, x, p, params);
        x = orc_bred(x, x, p, params);
    }
    return result;
}

/* BitReverse64, utils/utils.go:58-60 */
API u64 orc_bitreverse64(u64 index, u64 bitlen) {
    u64 r = 0;
    for (int i = 0; i < 64; i++) r |= ((index >> i) & 1) << (63 - i);
    return bitlen ? r >> (64 - bitlen) : 0;
}

/* smallPrimes, ring/utils.go:290-391: exactly the first 2000 primes
 * (all primes below 17390; checked against the table by tests). */
#define N_SMALL_PRIMES 2000
static u64 small_primes[N_SMALL_PRIMES];
static int small_primes_ready = 0;
static void init_small_primes(void) {
    if (small_primes_ready) return;
    static unsigned char comp[17390];
    int k = 0;
    for (int i = 2; i < 17390; i++) {
        if (!comp[i]) {
            small_primes[k++] = (u64)i;
            for (int j = i * i; j < 17390; j += i) comp[j] = 1;
        }
    }
    small_primes_ready = 1;
}
API u64 orc_small_prime(int i) {
    init_small_primes();
    return (i >= 0 && i < N_SMALL_PRIMES) ? small_primes[i] : 0;
}

/* IsPrime, ring/utils.go:75-128.  The reference draws 50 random Miller-Rabin
 * bases (crypto/rand); any correct primality test returns the same answer, so
 * the twelve fixed bases that are deterministic for all 64-bit inputs are used. */
API int orc_is_prime(u64 num) {
    init_small_primes();
    if (num < 2) return 0;
    for (int i = 0; i < N_SMALL_PRIMES; i++)
        if (num == small_primes[i]) return 1;
    for (int i = 0; i < N_SMALL_PRIMES; i++)
        if (num % small_primes[i] == 0) return 0;
    u64 s = num - 1;
    int k = 0;
    while ((s & 1) == 0) {
        s >>= 1;
        k++;
    }
    u64 params[2];
    orc_bred_params(num, params);
    static const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (int t = 0; t < 12; t++) {
        u64 x = orc_modexp(bases[t], s, num);
        if (x != 1) {
            int i = 0;
            while (x != num - 1) {
                if (i == k - 1) return 0;
                i++;
                x = orc_bred(x, x, num, params);
            }
        }
    }
    return 1;
}

/* GenerateNTTPrimes, ring/utils.go:133-175 (the y branch is kept literally). */
API int orc_generate_ntt_primes(u64 logQ, u64 logN, u64 levels, u64 *primes) {
    if (logQ > 60) return -1;
    u64 n = 0;
    u64 Qpow2 = (u64)1 << logQ;
    u64 _2N = (u64)2 << logN;
    u64 x = Qpow2 + 1, y = Qpow2 + 1;
    if (levels == 0) return 0;
    for (;;) {
        if (orc_is_prime(x)) {
            primes[n++] = x;
            if (n == levels) return (int)n;
        }
        x += _2N;
        if (_2N > y) {
            y -= _2N;
            if (orc_is_prime(y)) {
                primes[n++] = y;
                if (n == levels) return (int)n;
            }
        }
    }
}

/* gcd, ring/utils.go:53-61 */
static u64 gcd_u64(u64 a, u64 b) {
    if (a == 0 || b == 0) return 0;
    while (b != 0) {
        u64 t = a % b;
        a = b;
        b = t;
    }
    return a;
}

/* polynomialPollardsRho, ring/utils.go:212-218 */
static u64 poly_pollard(u64 x1, u64 x2, u64 c) {
    u64 z = orc_modexp(x1, 2, x2);
    z += c;
    z %= x2;
    return z;
}

/* factorizationPollardsRho, ring/utils.go:222-248 */
static u64 pollard_rho(u64 m) {
    u64 x, y, d = 0, c;
    for (c = 1; c < 10; c++) {
        x = 2;
        y = 2;
        d = 1;
        while (d != 0) {
            x = poly_pollard(x, m, c);
            y = poly_pollard(poly_pollard(y, m, c), m, c);
            if (y > x) {
                u64 t = x;
                x = y;
                y = t;
            }
            d = gcd_u64(x - y, m);
            if (d > 1) return d;
        }
    }
    return d;
}

/* getFactors, ring/utils.go:251-288 */
static int get_factors(u64 n, u64 *factors) {
    init_small_primes();
    int nf = 0;
    u64 m = n;
    for (int i = 0; i < N_SMALL_PRIMES; i++) {
        u64 sp = small_primes[i];
        int add = 0;
        while (m % sp == 0) {
            m /= sp;
            add = 1;
        }
        if (add) factors[nf++] = sp;
    }
    if (m == 1) return nf;
    for (;;) {
        u64 factor = pollard_rho(m);
        if (factor == 0) {
            factors[nf++] = m;
            break;
        }
        m /= factor;
        if (nf > 0 && factor == factors[nf - 1]) continue;
        factors[nf++] = factor;
    }
    return nf;
}

/* primitiveRoot, ring/utils.go:182-205 */
API u64 orc_primitive_root(u64 q) {
    u64 factors[128];
    int nf = get_factors(q - 1, factors);
    u64 g = 2;
    int not_found = 1;
    while (not_found) {
        g++;
        for (int i = 0; i < nf; i++) {
            u64 tmp = (q - 1) / factors[i];
            if (orc_modexp(g, tmp, q) == 1) {
                not_found = 1;
                break;
            }
            not_found = 0;
        }
    }
    return g;
}

/* ---------------------------------------------------------------------- */
/* ring/ring_context.go                                                   */
/* ---------------------------------------------------------------------- */

typedef struct {
    u64 N;
    int nl;
    u64 *modulus;
    u64 (*bred)[2];
    u64 *mred;
    u64 **rescale; /* rescale[j-1][i], i<j   (ring_context.go:148-158) */
    u64 *psi_mont, *psi_inv_mont;
    u64 **ntt_psi, **ntt_psi_inv;
    u64 *ntt_ninv;
    int allows_ntt;
} orc_ctx;

static int log2u(u64 n) {
    int l = 0;
    while (((u64)1 << l) < n) l++;
    return l;
}

API void orc_ctx_free(orc_ctx *c) {
    if (!c) return;
    if (c->rescale) {
        for (int j = 0; j < c->nl - 1; j++) free(c->rescale[j]);
        free(c->rescale);
    }
    if (c->ntt_psi)
        for (int i = 0; i < c->nl; i++) free(c->ntt_psi[i]);
    if (c->ntt_psi_inv)
        for (int i = 0; i < c->nl; i++) free(c->ntt_psi_inv[i]);
    free(c->ntt_psi);
    free(c->ntt_psi_inv);
    free(c->modulus);
    free(c->bred);
    free(c->mred);
    free(c->psi_mont);
    free(c->psi_inv_mont);
    free(c->ntt_ninv);
    free(c);
}

/* SetParameters (ring_context.go:68-124) + GenNTTParams (:129-209).
 * Returns NULL when N is not a power of two (the reference panics, :72) or a
 * modulus does not allow the NTT (the reference returns an error, :142-145). */
API orc_ctx *orc_ctx_new(u64 N, int nl, const u64 *moduli) {
    if (N == 0 || (N & (N - 1)) != 0 || nl <= 0) return NULL;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->N = N;
    c->nl = nl;
    c->modulus = (u64 *)malloc(sizeof(u64) * nl);
    c->bred = (u64(*)[2])malloc(sizeof(u64[2]) * nl);
    c->mred = (u64 *)calloc(nl, sizeof(u64));
    for (int i = 0; i < nl; i++) {
        u64 qi = moduli[i];
        c->modulus[i] = qi;
        orc_bred_params(qi, c->bred[i]);
        if ((qi & (qi - 1)) != 0 && qi != 0) c->mred[i] = orc_mred_params(qi);
    }
    /* GenNTTParams */
    for (int i = 0; i < nl; i++) {
        u64 qi = moduli[i];
        if (!orc_is_prime(qi) || (qi & ((N << 1) - 1)) != 1) {
            orc_ctx_free(c);
            return NULL;
        }
    }
    c->rescale = (u64 **)calloc(nl > 1 ? nl - 1 : 1, sizeof(u64 *));
    for (int j = nl - 1; j > 0; j--) {
        c->rescale[j - 1] = (u64 *)malloc(sizeof(u64) * j);
        for (int i = 0; i < j; i++)
            c->rescale[j - 1][i] =
                orc_mform(orc_modexp(c->modulus[j], c->modulus[i] - 2, c->modulus[i]), c->modulus[i], c->bred[i]);
    }
    c->psi_mont = (u64 *)malloc(sizeof(u64) * nl);
    c->psi_inv_mont = (u64 *)malloc(sizeof(u64) * nl);
    c->ntt_psi = (u64 **)calloc(nl, sizeof(u64 *));
    c->ntt_psi_inv = (u64 **)calloc(nl, sizeof(u64 *));
    c->ntt_ninv = (u64 *)malloc(sizeof(u64) * nl);
    u64 bitlen = (u64)log2u(N);
    for (int i = 0; i < nl; i++) {
        u64 qi = c->modulus[i];
        c->ntt_ninv[i] = orc_mform(orc_modexp(N, qi - 2, qi), qi, c->bred[i]);
        c->ntt_psi[i] = (u64 *)malloc(sizeof(u64) * N);
        c->ntt_psi_inv[i] = (u64 *)malloc(sizeof(u64) * N);
        u64 g = orc_primitive_root(qi);
        u64 _2n = N << 1;
        u64 power = (qi - 1) / _2n;
        u64 power_inv = (qi - 1) - power;
        u64 psi = orc_mform(orc_modexp(g, power, qi), qi, c->bred[i]);
        u64 psi_inv = orc_mform(orc_modexp(g, power_inv, qi), qi, c->bred[i]);
        c->psi_mont[i] = psi;
        c->psi_inv_mont[i] = psi_inv;
        c->ntt_psi[i][0] = orc_mform(1, qi, c->bred[i]);
        c->ntt_psi_inv[i][0] = orc_mform(1, qi, c->bred[i]);
        for (u64 j = 1; j < N; j++) {
            u64 prev = orc_bitreverse64(j - 1, bitlen);
            u64 next = orc_bitreverse64(j, bitlen);
            c->ntt_psi[i][next] = orc_mred(c->ntt_psi[i][prev], psi, qi, c->mred[i]);
            c->ntt_psi_inv[i][next] = orc_mred(c->ntt_psi_inv[i][prev], psi_inv, qi, c->mred[i]);
        }
    }
    c->allows_ntt = 1;
    return c;
}

API u64 orc_ctx_n(const orc_ctx *c) { return c->N; }
API int orc_ctx_nlimbs(const orc_ctx *c) { return c->nl; }
API void orc_ctx_scalars(const orc_ctx *c, u64 *modulus, u64 *bred /*2*nl*/, u64 *mred, u64 *ninv, u64 *psi_mont,
                         u64 *psi_inv_mont) {
    for (int i = 0; i < c->nl; i++) {
        if (modulus) modulus[i] = c->modulus[i];
        if (bred) {
            bred[2 * i] = c->bred[i][0];
            bred[2 * i + 1] = c->bred[i][1];
        }
        if (mred) mred[i] = c->mred[i];
        if (ninv) ninv[i] = c->ntt_ninv[i];
        if (psi_mont) psi_mont[i] = c->psi_mont[i];
        if (psi_inv_mont) psi_inv_mont[i] = c->psi_inv_mont[i];
    }
}
API void orc_ctx_tables(const orc_ctx *c, int limb, u64 *psi, u64 *psi_inv) {
    if (psi) memcpy(psi, c->ntt_psi[limb], sizeof(u64) * c->N);
    if (psi_inv) memcpy(psi_inv, c->ntt_psi_inv[limb], sizeof(u64) * c->N);
}
/* rescaleParams[j-1][i] for i<j */
API u64 orc_ctx_rescale_param(const orc_ctx *c, int j, int i) { return c->rescale[j - 1][i]; }

/* ---------------------------------------------------------------------- */
/* ring/ntt.go                                                            */
/* ---------------------------------------------------------------------- */

/* Butterfly, ntt.go:32-40 */
static inline void butterfly(u64 U, u64 V, u64 psi, u64 Q, u64 qinv, u64 *X, u64 *Y) {
    if (U > 2 * Q) U -= 2 * Q;
    V = orc_mred_constant(V, psi, Q, qinv);
    *X = U + V;
    *Y = U + 2 * Q - V;
}

/* InvButterfly, ntt.go:43-50 */
static inline void inv_butterfly(u64 U, u64 V, u64 psi, u64 Q, u64 qinv, u64 *X, u64 *Y) {
    u64 x = U + V;
    if (x > 2 * Q) x -= 2 * Q;
    *X = x;
    *Y = orc_mred_constant(U + 2 * Q - V, psi, Q, qinv);
}

/* NTT, ntt.go:53-86 */
API void orc_ntt_limb(const u64 *in, u64 *out, u64 N, const u64 *ntt_psi, u64 Q, u64 mredp, const u64 bredp[2]) {
    u64 t = N >> 1, j2 = t - 1;
    u64 F = ntt_psi[1];
    for (u64 j = 0; j <= j2; j++) butterfly(in[j], in[j + t], F, Q, mredp, &out[j], &out[j + t]);
    for (u64 m = 2; m < N; m <<= 1) {
        t >>= 1;
        for (u64 i = 0; i < m; i++) {
            u64 j1 = (i * t) << 1;
            j2 = j1 + t - 1;
            F = ntt_psi[m + i];
            for (u64 j = j1; j <= j2; j++) butterfly(out[j], out[j + t], F, Q, mredp, &out[j], &out[j + t]);
        }
    }
    for (u64 i = 0; i < N; i++) out[i] = orc_bred_add(out[i], Q, bredp);
}

/* InvNTT, ntt.go:89-139 */
API void orc_invntt_limb(const u64 *in, u64 *out, u64 N, const u64 *ntt_psi_inv, u64 ninv, u64 Q, u64 mredp) {
    u64 t = 1, j1 = 0, h = N >> 1, j2, F;
    for (u64 i = 0; i < h; i++) {
        j2 = j1;
        F = ntt_psi_inv[h + i];
        for (u64 j = j1; j <= j2; j++) inv_butterfly(in[j], in[j + t], F, Q, mredp, &out[j], &out[j + t]);
        j1 = j1 + (t << 1);
    }
    t <<= 1;
    for (u64 m = N >> 1; m > 1; m >>= 1) {
        j1 = 0;
        h = m >> 1;
        for (u64 i = 0; i < h; i++) {
            j2 = j1 + t - 1;
            F = ntt_psi_inv[h + i];
            for (u64 j = j1; j <= j2; j++) inv_butterfly(out[j], out[j + t], F, Q, mredp, &out[j], &out[j + t]);
            j1 = j1 + (t << 1);
        }
        t <<= 1;
    }
    for (u64 j = 0; j < N; j++) out[j] = orc_mred(out[j], ninv, Q, mredp);
}

/* Context.NTTLvl / NTT, ntt.go:4-15 (nl = level+1 limbs) */
API void orc_ntt(const orc_ctx *c, int nl, const u64 *p1, u64 *p2) {
    for (int x = 0; x < nl; x++)
        orc_ntt_limb(p1 + x * c->N, p2 + x * c->N, c->N, c->ntt_psi[x], c->modulus[x], c->mred[x], c->bred[x]);
}
/* Context.InvNTTLvl / InvNTT, ntt.go:18-29 */
API void orc_invntt(const orc_ctx *c, int nl, const u64 *p1, u64 *p2) {
    for (int x = 0; x < nl; x++)
        orc_invntt_limb(p1 + x * c->N, p2 + x * c->N, c->N, c->ntt_psi_inv[x], c->ntt_ninv[x], c->modulus[x],
                        c->mred[x]);
}
/* single limb with the tables of limb `x` of the context (ring.NTT call sites
 * such as ckks/evaluator.go:1586, ring_basis_extension.go:233) */
API void orc_ntt_one(const orc_ctx *c, int x, const u64 *in, u64 *out) {
    orc_ntt_limb(in, out, c->N, c->ntt_psi[x], c->modulus[x], c->mred[x], c->bred[x]);
}
API void orc_invntt_one(const orc_ctx *c, int x, const u64 *in, u64 *out) {
    orc_invntt_limb(in, out, c->N, c->ntt_psi_inv[x], c->ntt_ninv[x], c->modulus[x], c->mred[x]);
}

/* ---------------------------------------------------------------------- */
/* ring/ring.go  (every op takes nl = level+1 active limbs)               */
/* ---------------------------------------------------------------------- */

#define LIMB_LOOP(c, nl)                 \
    for (int i = 0; i < (nl); i++) {     \
        const u64 qi = (c)->modulus[i];  \
        const u64 mp = (c)->mred[i];     \
        const u64 *bp = (c)->bred[i];    \
        (void)qi; (void)mp; (void)bp;    \
        for (u64 j = 0; j < (c)->N; j++) {  \
            const u64 k = (u64)i * (c)->N + j;
#define LIMB_END }}

/* Add/AddLvl ring.go:10-29 */
API void orc_add(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = orc_cred(p1[k] + p2[k], qi); LIMB_END
}
/* AddNoMod(Lvl) ring.go:32-51 */
API void orc_add_nomod(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = p1[k] + p2[k]; LIMB_END
}
/* Sub/SubLvl ring.go:54-73 */
API void orc_sub(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = orc_cred((p1[k] + qi) - p2[k], qi); LIMB_END
}
/* SubNoMod(Lvl) ring.go:76-97 */
API void orc_sub_nomod(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = (p1[k] + qi) - p2[k]; LIMB_END
}
/* Neg/NegLvl ring.go:100-119 */
API void orc_neg(const orc_ctx *c, int nl, const u64 *p1, u64 *p2) {
    LIMB_LOOP(c, nl) p2[k] = qi - p1[k]; LIMB_END
}
/* Reduce/ReduceLvl ring.go:122-143 */
API void orc_reduce(const orc_ctx *c, int nl, const u64 *p1, u64 *p2) {
    LIMB_LOOP(c, nl) p2[k] = orc_bred_add(p1[k], qi, bp); LIMB_END
}
/* MulCoeffs ring.go:187-195 (Barrett) */
API void orc_mulcoeffs(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = orc_bred(p1[k], p2[k], qi, bp); LIMB_END
}
/* MulCoeffsAndAdd ring.go:198-206 */
API void orc_mulcoeffs_and_add(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = orc_cred(p3[k] + orc_bred(p1[k], p2[k], qi, bp), qi); LIMB_END
}
/* MulCoeffsAndAddNoMod ring.go:209-217 */
API void orc_mulcoeffs_and_add_nomod(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] += orc_bred(p1[k], p2[k], qi, bp); LIMB_END
}
/* MulCoeffsConstant ring.go:335-343 */
API void orc_mulcoeffs_constant(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = orc_bred_constant(p1[k], p2[k], qi, bp); LIMB_END
}
/* MulCoeffsMontgomery(Lvl) ring.go:221-243 */
API void orc_mulcoeffs_montgomery(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = orc_mred(p1[k], p2[k], qi, mp); LIMB_END
}
/* MulCoeffsMontgomeryAndAdd(Lvl) ring.go:247-269 */
API void orc_mulcoeffs_montgomery_and_add(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = orc_cred(p3[k] + orc_mred(p1[k], p2[k], qi, mp), qi); LIMB_END
}
/* MulCoeffsMontgomeryAndAddNoMod(Lvl) ring.go:273-295 */
API void orc_mulcoeffs_montgomery_and_add_nomod(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] += orc_mred(p1[k], p2[k], qi, mp); LIMB_END
}
/* MulCoeffsMontgomeryConstantAndAddNoModLvl ring.go:298-308 */
API void orc_mulcoeffs_montgomery_constant_and_add_nomod(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2,
                                                        u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] += orc_mred_constant(p1[k], p2[k], qi, mp); LIMB_END
}
/* MulCoeffsMontgomeryAndSub ring.go:311-319 */
API void orc_mulcoeffs_montgomery_and_sub(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = orc_cred(p3[k] + (qi - orc_mred(p1[k], p2[k], qi, mp)), qi); LIMB_END
}
/* MulCoeffsMontgomeryAndSubNoMod ring.go:323-331 */
API void orc_mulcoeffs_montgomery_and_sub_nomod(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = p3[k] + (qi - orc_mred(p1[k], p2[k], qi, mp)); LIMB_END
}
/* MulCoeffsMontgomeryConstant ring.go:346-355 */
API void orc_mulcoeffs_montgomery_constant(const orc_ctx *c, int nl, const u64 *p1, const u64 *p2, u64 *p3) {
    LIMB_LOOP(c, nl) p3[k] = orc_mred_constant(p1[k], p2[k], qi, mp); LIMB_END
}
/* MForm/MFormLvl ring.go:583-607 */
API void orc_mform_poly(const orc_ctx *c, int nl, const u64 *p1, u64 *p2) {
    LIMB_LOOP(c, nl) p2[k] = orc_mform(p1[k], qi, bp); LIMB_END
}
/* InvMForm ring.go:610-619 */
API void orc_invmform_poly(const orc_ctx *c, int nl, const u64 *p1, u64 *p2) {
    LIMB_LOOP(c, nl) p2[k] = orc_invmform(p1[k], qi, mp); LIMB_END
}
/* AddScalar ring.go:467-474 / AddScalarBigint :477-487.  scalar[i] is the
 * per-limb value the reference adds (the same word for AddScalar, scalar mod
 * q_i for the big.Int variant).  The reference writes the result into p1 and
 * ignores p2 (p2tmp aliases p1.Coeffs[i]); that is restated. */
API void orc_add_scalar(const orc_ctx *c, int nl, u64 *p1, const u64 *scalar) {
    LIMB_LOOP(c, nl) p1[k] = orc_cred(p1[k] + scalar[i], qi); LIMB_END
}
/* SubScalar ring.go:490-497 / SubScalarBigint :500-510 (same aliasing) */
API void orc_sub_scalar(const orc_ctx *c, int nl, u64 *p1, const u64 *scalar) {
    LIMB_LOOP(c, nl) p1[k] = orc_cred(p1[k] + (qi - scalar[i]), qi); LIMB_END
}
/* MulScalar(Lvl) ring.go:513-536, MulScalarBigint(Lvl) :539-572; scalar[i] is
 * the word fed to BRedAdd (scalar itself, or scalar mod q_i). */
API void orc_mul_scalar(const orc_ctx *c, int nl, const u64 *p1, const u64 *scalar, u64 *p2) {
    for (int i = 0; i < nl; i++) {
        u64 qi = c->modulus[i];
        u64 sm = orc_mform(orc_bred_add(scalar[i], qi, c->bred[i]), qi, c->bred[i]);
        for (u64 j = 0; j < c->N; j++) p2[i * c->N + j] = orc_mred(p1[i * c->N + j], sm, qi, c->mred[i]);
    }
}
/* MulByPow2(Lvl) ring.go:629-653: MForm then PowerOf2 of the ORIGINAL p1
 * words (the loop reads p1tmp, not the MForm'd p2) -- restated literally. */
API void orc_mul_by_pow2(const orc_ctx *c, int nl, const u64 *p1, u64 pow2, u64 *p2) {
    u64 *src = (u64 *)malloc(sizeof(u64) * c->N * nl);
    memcpy(src, p1, sizeof(u64) * c->N * nl);
    orc_mform_poly(c, nl, p1, p2);
    const u64 *rd = (p1 == p2) ? p2 : src; /* if aliased the loop sees the MForm'd words */
    LIMB_LOOP(c, nl) p2[k] = orc_power_of_2(rd[k], pow2, qi, mp); LIMB_END
    free(src);
}
/* MultByMonomial ring.go:663-723 */
API void orc_mult_by_monomial(const orc_ctx *c, int nl, const u64 *p1, u64 deg, u64 *p2) {
    u64 N = c->N;
    u64 shift = deg % (N << 1);
    if (shift == 0) {
        LIMB_LOOP(c, nl) p2[k] = p1[k]; LIMB_END
        return;
    }
    u64 *tmpx = (u64 *)malloc(sizeof(u64) * N * nl);
    if (shift < N) {
        LIMB_LOOP(c, nl) tmpx[k] = p1[k]; LIMB_END
    } else {
        LIMB_LOOP(c, nl) tmpx[k] = qi - p1[k]; LIMB_END
    }
    shift %= N;
    for (int i = 0; i < nl; i++) {
        u64 qi = c->modulus[i];
        for (u64 j = 0; j < shift; j++) p2[i * N + j] = qi - tmpx[i * N + N - shift + j];
        for (u64 j = shift; j < N; j++) p2[i * N + j] = tmpx[i * N + j - shift];
    }
    free(tmpx);
}
/* MulByVectorMontgomery ring.go:726-734 */
API void orc_mul_by_vector_montgomery(const orc_ctx *c, int nl, const u64 *p1, const u64 *vec, u64 *p2) {
    LIMB_LOOP(c, nl) p2[k] = orc_mred(p1[k], vec[j], qi, mp); LIMB_END
}
/* MulByVectorMontgomeryAndAddNoMod ring.go:737-745 */
API void orc_mul_by_vector_montgomery_and_add_nomod(const orc_ctx *c, int nl, const u64 *p1, const u64 *vec,
                                                   u64 *p2) {
    LIMB_LOOP(c, nl) p2[k] += orc_mred(p1[k], vec[j], qi, mp); LIMB_END
}
/* BitReverse ring.go:749-772 (out of place form) */
API void orc_bitreverse_poly(const orc_ctx *c, int nl, const u64 *p1, u64 *p2) {
    u64 bl = (u64)log2u(c->N);
    for (int i = 0; i < nl; i++)
        for (u64 j = 0; j < c->N; j++) p2[i * c->N + orc_bitreverse64(j, bl)] = p1[i * c->N + j];
}

/* ---------------------------------------------------------------------- */
/* ring/ring_galois.go                                                    */
/* ---------------------------------------------------------------------- */

/* GenGaloisParams ring_galois.go:9-26 */
API void orc_gen_galois_params(u64 n, u64 gen, u64 *out /* n/2 */) {
    u64 mask = (n << 1) - 1;
    out[0] = 1;
    for (u64 i = 1; i < (n >> 1); i++) out[i] = (out[i - 1] * gen) & mask;
}
/* PermuteNTTIndex ring_galois.go:29-50 */
API void orc_permute_ntt_index(u64 gen, u64 power, u64 N, u64 *index) {
    u64 genpow = orc_modexp(gen, power, 2 * N);
    u64 logN = (u64)log2u(N), mask = (N << 1) - 1;
    for (u64 i = 0; i < N; i++) {
        u64 tmp1 = 2 * orc_bitreverse64(i, logN) + 1;
        u64 tmp2 = (((genpow * tmp1) & mask) - 1) >> 1;
        index[i] = orc_bitreverse64(tmp2, logN);
    }
}
/* PermuteNTTWithIndex ring_galois.go:89-101 (not in place) */
API void orc_permute_ntt_with_index(u64 N, int nl, const u64 *in, const u64 *index, u64 *out) {
    for (u64 j = 0; j < N; j++) {
        u64 tmp = index[j];
        for (int i = 0; i < nl; i++) out[i * N + j] = in[i * N + tmp];
    }
}
/* PermuteNTT ring_galois.go:55-84 (gen used as is, not exponentiated) */
API void orc_permute_ntt(u64 N, int nl, const u64 *in, u64 gen, u64 *out) {
    u64 logN = (u64)log2u(N), mask = (N << 1) - 1;
    u64 *index = (u64 *)malloc(sizeof(u64) * N);
    for (u64 i = 0; i < N; i++) {
        u64 tmp1 = 2 * orc_bitreverse64(i, logN) + 1;
        u64 tmp2 = (((gen * tmp1) & mask) - 1) >> 1;
        index[i] = orc_bitreverse64(tmp2, logN);
    }
    orc_permute_ntt_with_index(N, nl, in, index, out);
    free(index);
}
/* Context.Permute ring_galois.go:106-127 (coefficient domain, not in place) */
API void orc_permute(const orc_ctx *c, int nl, const u64 *in, u64 gen, u64 *out) {
    u64 N = c->N, mask = N - 1, logN = (u64)log2u(N);
    for (u64 i = 0; i < N; i++) {
        u64 raw = i * gen;
        u64 index = raw & mask;
        u64 tmp = (raw >> logN) & 1;
        for (int j = 0; j < nl; j++) {
            u64 qi = c->modulus[j];
            out[j * N + index] = (in[j * N + i] * (tmp ^ 1)) | ((qi - in[j * N + i]) * tmp);
        }
    }
}

/* ---------------------------------------------------------------------- */
/* ring/ring_basis_extension.go                                           */
/* ---------------------------------------------------------------------- */

typedef struct {
    int nq, np;
    u64 *Q, *P;
    u64 *qib_mont;   /* [nq] */
    u64 *qispj_mont; /* [nq][np] */
    u64 *qpj_inv;    /* [np][nq+1] */
    u64 (*bredQ)[2], *mredQ;
    u64 (*bredP)[2], *mredP;
} modup_params;

static u64 mulmod(u64 a, u64 b, u64 m) { return (u64)(((u128)a * b) % m); }
static u64 powmod(u64 a, u64 e, u64 m) {
    u64 r = 1 % m;
    a %= m;
    while (e) {
        if (e & 1) r = mulmod(r, a, m);
        a = mulmod(a, a, m);
        e >>= 1;
    }
    return r;
}

static void modup_free(modup_params *p) {
    if (!p) return;
    free(p->Q); free(p->P); free(p->qib_mont); free(p->qispj_mont); free(p->qpj_inv);
    free(p->bredQ); free(p->mredQ); free(p->bredP); free(p->mredP);
    free(p);
}

/* basisextenderparameters, ring_basis_extension.go:76-142.  The reference
 * uses math/big for Q/q_i, its inverse mod q_i and the residues mod p_j;
 * those are canonical residues, computed here with 128-bit modular products
 * (same integers). */
static modup_params *modup_new(const u64 *Q, int nq, const u64 *P, int np) {
    modup_params *p = (modup_params *)calloc(1, sizeof(modup_params));
    p->nq = nq; p->np = np;
    p->Q = (u64 *)malloc(sizeof(u64) * nq);
    p->P = (u64 *)malloc(sizeof(u64) * np);
    p->bredQ = (u64(*)[2])malloc(sizeof(u64[2]) * nq);
    p->mredQ = (u64 *)malloc(sizeof(u64) * nq);
    p->bredP = (u64(*)[2])malloc(sizeof(u64[2]) * np);
    p->mredP = (u64 *)malloc(sizeof(u64) * np);
    for (int i = 0; i < nq; i++) {
        p->Q[i] = Q[i];
        orc_bred_params(Q[i], p->bredQ[i]);
        p->mredQ[i] = orc_mred_params(Q[i]);
    }
    for (int j = 0; j < np; j++) {
        p->P[j] = P[j];
        orc_bred_params(P[j], p->bredP[j]);
        p->mredP[j] = orc_mred_params(P[j]);
    }
    p->qib_mont = (u64 *)malloc(sizeof(u64) * nq);
    p->qispj_mont = (u64 *)malloc(sizeof(u64) * nq * np);
    for (int i = 0; i < nq; i++) {
        u64 qi = Q[i];
        /* QiStar mod qi, then its inverse (big.Int ModInverse, prime modulus) */
        u64 star = 1 % qi;
        for (int k = 0; k < nq; k++)
            if (k != i) star = mulmod(star, Q[k] % qi, qi);
        u64 barre = powmod(star, qi - 2, qi);
        p->qib_mont[i] = orc_mform(barre, qi, p->bredQ[i]);
        for (int j = 0; j < np; j++) {
            u64 pj = P[j];
            u64 s = 1 % pj;
            for (int k = 0; k < nq; k++)
                if (k != i) s = mulmod(s, Q[k] % pj, pj);
            p->qispj_mont[i * np + j] = orc_mform(s, pj, p->bredP[j]);
        }
    }
    p->qpj_inv = (u64 *)malloc(sizeof(u64) * np * (nq + 1));
    for (int j = 0; j < np; j++) {
        u64 pj = P[j];
        u64 qm = 1 % pj;
        for (int k = 0; k < nq; k++) qm = mulmod(qm, Q[k] % pj, pj);
        u64 v = pj - qm;
        u64 *row = p->qpj_inv + j * (nq + 1);
        row[0] = 0;
        for (int i = 1; i < nq + 1; i++) row[i] = orc_cred(row[i - 1] + v, pj);
    }
    return p;
}

/* modUpExact, ring_basis_extension.go:352-393.  p1 = n1 source limbs,
 * p2 = n2 target limbs (n1 <= params->nq, n2 <= params->np), flat with stride N. */
static void modup_exact(const u64 *p1, int n1, u64 *p2, int n2, u64 N, const modup_params *pr) {
    u64 y[64];
    int np = pr->np, nq = pr->nq;
    for (u64 x = 0; x < N; x++) {
        double vi = 0;
        for (int i = 0; i < n1; i++) {
            y[i] = orc_mred(p1[i * N + x], pr->qib_mont[i], pr->Q[i], pr->mredQ[i]);
            vi += (double)y[i] / (double)pr->Q[i];
        }
        u64 v = (u64)vi;
        for (int j = 0; j < n2; j++) {
            u64 xpj = 0;
            for (int i = 0; i < n1; i++) {
                xpj += orc_mred(y[i], pr->qispj_mont[i * np + j], pr->P[j], pr->mredP[j]);
                if ((i & 7) == 6) xpj = orc_bred_add(xpj, pr->P[j], pr->bredP[j]);
            }
            p2[j * N + x] = orc_bred_add(xpj + pr->qpj_inv[j * (nq + 1) + v], pr->P[j], pr->bredP[j]);
        }
    }
}

/* genModDownParams, ring_basis_extension.go:39-53: for each modulus of ctx_a,
 * MForm((prod of ctx_b's moduli)^-1 mod it). */
static u64 *gen_moddown(const orc_ctx *a, const orc_ctx *b) {
    u64 *params = (u64 *)malloc(sizeof(u64) * a->nl);
    for (int i = 0; i < a->nl; i++) {
        u64 Qi = a->modulus[i];
        u64 m = 1 % Qi;
        for (int k = 0; k < b->nl; k++) m = mulmod(m, b->modulus[k] % Qi, Qi);
        m = orc_modexp(m, Qi - 2, Qi);
        params[i] = orc_mform(m, Qi, a->bred[i]);
    }
    return params;
}

/* FastBasisExtender, ring_basis_extension.go:9-74 */
typedef struct {
    const orc_ctx *ctxQ, *ctxP;
    modup_params *paramsQP, *paramsPQ;
    u64 *moddownPQ; /* per Q limb: P^-1 */
    u64 *moddownQP; /* per P limb: Q^-1 */
    u64 *poolQ, *poolP;
} orc_extender;

API orc_extender *orc_extender_new(const orc_ctx *ctxQ, const orc_ctx *ctxP) {
    orc_extender *e = (orc_extender *)calloc(1, sizeof(orc_extender));
    e->ctxQ = ctxQ; e->ctxP = ctxP;
    e->paramsQP = modup_new(ctxQ->modulus, ctxQ->nl, ctxP->modulus, ctxP->nl);
    e->paramsPQ = modup_new(ctxP->modulus, ctxP->nl, ctxQ->modulus, ctxQ->nl);
    e->moddownPQ = gen_moddown(ctxQ, ctxP);
    e->moddownQP = gen_moddown(ctxP, ctxQ);
    e->poolQ = (u64 *)calloc(ctxQ->N * ctxQ->nl, sizeof(u64));
    e->poolP = (u64 *)calloc(ctxP->N * ctxP->nl, sizeof(u64));
    return e;
}
API void orc_extender_free(orc_extender *e) {
    if (!e) return;
    modup_free(e->paramsQP); modup_free(e->paramsPQ);
    free(e->moddownPQ); free(e->moddownQP); free(e->poolQ); free(e->poolP);
    free(e);
}
API void orc_extender_params(const orc_extender *e, u64 *moddownPQ, u64 *moddownQP) {
    if (moddownPQ) memcpy(moddownPQ, e->moddownPQ, sizeof(u64) * e->ctxQ->nl);
    if (moddownQP) memcpy(moddownQP, e->moddownQP, sizeof(u64) * e->ctxP->nl);
}

/* ModUpSplitQP :147-149 : p1 over Q[:level+1] -> p2 over all of P */
API void orc_modup_split_qp(const orc_extender *e, int level, const u64 *p1, u64 *p2) {
    modup_exact(p1, level + 1, p2, e->paramsQP->np, e->ctxQ->N, e->paramsQP);
}
/* ModUpSplitPQ :154-156 : p1 over P[:level+1] -> p2 over all of Q */
API void orc_modup_split_pq(const orc_extender *e, int level, const u64 *p1, u64 *p2) {
    modup_exact(p1, level + 1, p2, e->paramsPQ->np, e->ctxQ->N, e->paramsPQ);
}

static void moddown_tail(const orc_ctx *c, int nl, const u64 *p1, const u64 *p3, const u64 *params, u64 *p2) {
    for (int i = 0; i < nl; i++) {
        u64 qi = c->modulus[i];
        for (u64 j = 0; j < c->N; j++)
            p2[i * c->N + j] = orc_mred(p1[i * c->N + j] + (qi - p3[i * c->N + j]), params[i], qi, c->mred[i]);
    }
}

/* ModDownNTTPQ :163-200.  p1 has nQ+nP limbs (all of Q then P), is clobbered
 * in its P part; p2 receives level+1 limbs. */
API void orc_moddown_ntt_pq(orc_extender *e, int level, u64 *p1, u64 *p2) {
    const orc_ctx *Q = e->ctxQ, *P = e->ctxP;
    u64 N = Q->N;
    for (int j = 0; j < P->nl; j++) orc_invntt_one(P, j, p1 + (Q->nl + j) * N, p1 + (Q->nl + j) * N);
    modup_exact(p1 + Q->nl * N, P->nl, e->poolQ, level + 1, N, e->paramsPQ);
    for (int i = 0; i < level + 1; i++) orc_ntt_one(Q, i, e->poolQ + i * N, e->poolQ + i * N);
    moddown_tail(Q, level + 1, p1, e->poolQ, e->moddownPQ, p2);
}
/* ModDownSplitedNTTPQ :207-242 (clobbers p1P) */
API void orc_moddown_splited_ntt_pq(orc_extender *e, int level, const u64 *p1Q, u64 *p1P, u64 *p2) {
    const orc_ctx *Q = e->ctxQ, *P = e->ctxP;
    u64 N = Q->N;
    orc_invntt(P, P->nl, p1P, p1P);
    modup_exact(p1P, P->nl, e->poolQ, level + 1, N, e->paramsPQ);
    for (int i = 0; i < level + 1; i++) orc_ntt_one(Q, i, e->poolQ + i * N, e->poolQ + i * N);
    moddown_tail(Q, level + 1, p1Q, e->poolQ, e->moddownPQ, p2);
}
/* ModDownPQ :248-275: p1 = level+1 Q limbs followed by nP P limbs */
API void orc_moddown_pq(orc_extender *e, int level, const u64 *p1, u64 *p2) {
    const orc_ctx *Q = e->ctxQ;
    u64 N = Q->N;
    modup_exact(p1 + (u64)(level + 1) * N, e->paramsQP->np, e->poolQ, level + 1, N, e->paramsPQ);
    moddown_tail(Q, level + 1, p1, e->poolQ, e->moddownPQ, p2);
}
/* ModDownSplitedPQ :281-308 */
API void orc_moddown_splited_pq(orc_extender *e, int level, const u64 *p1Q, const u64 *p1P, u64 *p2) {
    const orc_ctx *Q = e->ctxQ;
    modup_exact(p1P, e->ctxP->nl, e->poolQ, level + 1, Q->N, e->paramsPQ);
    moddown_tail(Q, level + 1, p1Q, e->poolQ, e->moddownPQ, p2);
}
/* ModDownSplitedQP :314-350 */
API void orc_moddown_splited_qp(orc_extender *e, int levelQ, int levelP, const u64 *p1Q, const u64 *p1P, u64 *p2) {
    const orc_ctx *P = e->ctxP;
    orc_modup_split_qp(e, levelQ, p1Q, e->poolP);
    moddown_tail(P, levelP + 1, p1P, e->poolP, e->moddownQP, p2);
}

/* Decomposer, ring_basis_extension.go:398-472 */
typedef struct {
    int nQ, nP, alpha, beta;
    int *xalpha;
    modup_params ***modup; /* [beta][xalpha-1] */
} orc_decomposer;

API orc_decomposer *orc_decomposer_new(const u64 *Q, int nQ, const u64 *P, int nP) {
    orc_decomposer *d = (orc_decomposer *)calloc(1, sizeof(orc_decomposer));
    d->nQ = nQ; d->nP = nP; d->alpha = nP;
    d->beta = (nQ + nP - 1) / nP; /* ceil(len(Q)/alpha) */
    d->xalpha = (int *)malloc(sizeof(int) * d->beta);
    for (int i = 0; i < d->beta; i++) d->xalpha[i] = d->alpha;
    if (nQ % d->alpha != 0) d->xalpha[d->beta - 1] = nQ % d->alpha;
    d->modup = (modup_params ***)calloc(d->beta, sizeof(modup_params **));
    u64 *Pi = (u64 *)malloc(sizeof(u64) * (nQ + nP));
    for (int k = 0; k < nQ; k++) Pi[k] = Q[k];
    for (int k = nQ; k < nQ + nP; k++) Pi[k] = P[k - nQ];
    for (int i = 0; i < d->beta; i++) {
        int cnt = d->xalpha[i] - 1;
        d->modup[i] = (modup_params **)calloc(cnt > 0 ? cnt : 1, sizeof(modup_params *));
        for (int j = 0; j < cnt; j++) d->modup[i][j] = modup_new(Q + i * d->alpha, j + 2, Pi, nQ + nP);
    }
    free(Pi);
    return d;
}
API void orc_decomposer_free(orc_decomposer *d) {
    if (!d) return;
    for (int i = 0; i < d->beta; i++) {
        for (int j = 0; j < d->xalpha[i] - 1; j++) modup_free(d->modup[i][j]);
        free(d->modup[i]);
    }
    free(d->modup); free(d->xalpha); free(d);
}
API int orc_decomposer_beta(const orc_decomposer *d) { return d->beta; }
API int orc_decomposer_xalpha(const orc_decomposer *d, int i) { return d->xalpha[i]; }

static inline u64 conv_one(const modup_params *pr, const u64 *y, int ny, int t, u64 v) {
    u64 xpj = 0;
    for (int i = 0; i < ny; i++) {
        xpj += orc_mred(y[i], pr->qispj_mont[i * pr->np + t], pr->P[t], pr->mredP[t]);
        if ((i & 7) == 6) xpj = orc_bred_add(xpj, pr->P[t], pr->bredP[t]);
    }
    return orc_bred_add(xpj + pr->qpj_inv[t * (pr->nq + 1) + v], pr->P[t], pr->bredP[t]);
}

/* Shared body of Decompose (:476-597) and DecomposeAndSplit (:601-713):
 * outQ receives limbs 0..level, outP the nP special-prime limbs.  For
 * Decompose outP = outQ + (level+1)*N (the P limbs follow the active Q limbs). */
static void decompose_core(const orc_decomposer *d, u64 N, int level, int crt, const u64 *p0, u64 *outQ, u64 *outP) {
    int alphai = d->xalpha[crt];
    int p0idxst = crt * d->alpha;
    int p0idxed = p0idxst + alphai;
    if ((p0idxed > level + 1 && (level + 1) % d->nP == 1) || alphai == 1) {
        for (u64 x = 0; x < N; x++) {
            for (int j = 0; j < level + 1; j++) outQ[j * N + x] = p0[p0idxst * N + x];
            for (int j = 0; j < d->nP; j++) outP[j * N + x] = p0[p0idxst * N + x];
        }
        return;
    }
    int index;
    if (level >= alphai + crt * d->alpha)
        index = d->xalpha[crt] - 2;
    else
        index = (level - 1) % d->alpha;
    const modup_params *pr = d->modup[crt][index];
    int ny = index + 2;
    u64 y[64];
    for (u64 x = 0; x < N; x++) {
        double vi = 0;
        for (int i = 0; i < ny; i++) {
            outQ[(i + p0idxst) * N + x] = p0[(i + p0idxst) * N + x];
            y[i] = orc_mred(p0[(i + p0idxst) * N + x], pr->qib_mont[i], pr->Q[i], pr->mredQ[i]);
            vi += (double)y[i] / (double)pr->Q[i];
        }
        u64 v = (u64)vi;
        for (int j = 0; j < p0idxst; j++) outQ[j * N + x] = conv_one(pr, y, ny, j, v);
        for (int j = d->alpha * crt; j < level + 1; j++) outQ[j * N + x] = conv_one(pr, y, ny, j, v);
        for (int j = 0, u = d->nQ; j < d->nP; j++, u++) outP[j * N + x] = conv_one(pr, y, ny, u, v);
    }
}
/* DecomposeAndSplit :601-713 */
API void orc_decompose_and_split(const orc_decomposer *d, u64 N, int level, int crt, const u64 *p0, u64 *p1Q,
                                 u64 *p1P) {
    decompose_core(d, N, level, crt, p0, p1Q, p1P);
}
/* Decompose :476-597: p1 has level+1 Q limbs followed by the nP P limbs */
API void orc_decompose(const orc_decomposer *d, u64 N, int level, int crt, const u64 *p0, u64 *p1) {
    decompose_core(d, N, level, crt, p0, p1, p1 + (u64)(level + 1) * N);
}

/* ---------------------------------------------------------------------- */
/* ring/ring_scaling.go:1-164  (nl = current number of limbs; the result  */
/* has nl-1 limbs, the reference re-slices p0.Coeffs[:level])             */
/* ---------------------------------------------------------------------- */

/* DivFloorByLastModulusNTT :9-34 */
API void orc_div_floor_by_last_modulus_ntt(const orc_ctx *c, int nl, u64 *p0) {
    int level = nl - 1;
    u64 N = c->N;
    u64 *ptmp = (u64 *)malloc(sizeof(u64) * N);
    orc_invntt_one(c, level, p0 + level * N, p0 + level * N);
    for (int i = 0; i < level; i++) {
        orc_ntt_one(c, i, p0 + level * N, ptmp);
        u64 qi = c->modulus[i], rp = c->rescale[level - 1][i];
        for (u64 j = 0; j < N; j++) p0[i * N + j] = orc_mred(p0[i * N + j] + (qi - ptmp[j]), rp, qi, c->mred[i]);
    }
    free(ptmp);
}
/* DivFloorByLastModulus :37-54 */
API void orc_div_floor_by_last_modulus(const orc_ctx *c, int nl, u64 *p0) {
    int level = nl - 1;
    u64 N = c->N;
    for (int i = 0; i < level; i++) {
        u64 qi = c->modulus[i], rp = c->rescale[level - 1][i];
        for (u64 j = 0; j < N; j++)
            p0[i * N + j] = orc_mred(p0[i * N + j] + (qi - orc_bred_add(p0[level * N + j], qi, c->bred[i])), rp, qi,
                                     c->mred[i]);
    }
}
/* DivRoundByLastModulusNTT :72-114 */
API void orc_div_round_by_last_modulus_ntt(const orc_ctx *c, int nl, u64 *p0) {
    int level = nl - 1;
    u64 N = c->N;
    u64 *ptmp = (u64 *)malloc(sizeof(u64) * N);
    orc_invntt_one(c, level, p0 + level * N, p0 + level * N);
    u64 phalf = (c->modulus[level] - 1) >> 1;
    u64 *p0tmp = p0 + level * N;
    u64 pj = c->modulus[level];
    for (u64 i = 0; i < N; i++) p0tmp[i] = orc_cred(p0tmp[i] + phalf, pj);
    for (int i = 0; i < level; i++) {
        u64 qi = c->modulus[i], rp = c->rescale[level - 1][i];
        u64 phalf_neg = qi - orc_bred_add(phalf, qi, c->bred[i]);
        for (u64 j = 0; j < N; j++) ptmp[j] = p0tmp[j] + phalf_neg;
        orc_ntt_one(c, i, ptmp, ptmp);
        for (u64 j = 0; j < N; j++) p0[i * N + j] = orc_mred(p0[i * N + j] + (qi - ptmp[j]), rp, qi, c->mred[i]);
    }
    free(ptmp);
}
/* DivRoundByLastModulus :117-148 */
API void orc_div_round_by_last_modulus(const orc_ctx *c, int nl, u64 *p0) {
    int level = nl - 1;
    u64 N = c->N;
    u64 phalf = (c->modulus[level] - 1) >> 1;
    u64 *p0tmp = p0 + level * N;
    u64 pj = c->modulus[level];
    for (u64 i = 0; i < N; i++) p0tmp[i] = orc_cred(p0tmp[i] + phalf, pj);
    for (int i = 0; i < level; i++) {
        u64 qi = c->modulus[i], rp = c->rescale[level - 1][i];
        u64 phalf_neg = qi - orc_bred_add(phalf, qi, c->bred[i]);
        for (u64 j = 0; j < N; j++)
            p0[i * N + j] = orc_mred(p0[i * N + j] + (qi - orc_bred_add(p0tmp[j] + phalf_neg, qi, c->bred[i])), rp, qi,
                                     c->mred[i]);
    }
}
/* DivFloorByLastModulusMany :64-69 / ManyNTT :57-61 */
API void orc_div_floor_by_last_modulus_many(const orc_ctx *c, int nl, u64 *p0, int nb) {
    for (int k = 0; k < nb; k++) orc_div_floor_by_last_modulus(c, nl - k, p0);
}
API void orc_div_floor_by_last_modulus_many_ntt(const orc_ctx *c, int nl, u64 *p0, int nb) {
    orc_invntt(c, nl, p0, p0);
    orc_div_floor_by_last_modulus_many(c, nl, p0, nb);
    orc_ntt(c, nl - nb, p0, p0);
}
/* DivRoundByLastModulusMany :159-164 / ManyNTT :152-156 */
API void orc_div_round_by_last_modulus_many(const orc_ctx *c, int nl, u64 *p0, int nb) {
    for (int k = 0; k < nb; k++) orc_div_round_by_last_modulus(c, nl - k, p0);
}
API void orc_div_round_by_last_modulus_many_ntt(const orc_ctx *c, int nl, u64 *p0, int nb) {
    orc_invntt(c, nl, p0, p0);
    orc_div_round_by_last_modulus_many(c, nl, p0, nb);
    orc_ntt(c, nl - nb, p0, p0);
}

/* ---------------------------------------------------------------------- */
/* ckks/evaluator.go hot ops                                              */
/* ---------------------------------------------------------------------- */

typedef struct {
    const orc_ctx *ctxQ, *ctxP;
    orc_extender *ext;
    orc_decomposer *dec;
    int alpha, levels; /* levels = #Q (ckks.go:59) */
    u64 *poolQ[4], *poolP[3], *ringpool[6];
} orc_ckks_eval;

/* NewEvaluator, ckks/evaluator.go:81-112 */
API orc_ckks_eval *orc_ckks_eval_new(const orc_ctx *ctxQ, const orc_ctx *ctxP) {
    orc_ckks_eval *e = (orc_ckks_eval *)calloc(1, sizeof(orc_ckks_eval));
    e->ctxQ = ctxQ; e->ctxP = ctxP;
    e->ext = orc_extender_new(ctxQ, ctxP);
    e->dec = orc_decomposer_new(ctxQ->modulus, ctxQ->nl, ctxP->modulus, ctxP->nl);
    e->alpha = ctxP->nl;
    e->levels = ctxQ->nl;
    for (int i = 0; i < 4; i++) e->poolQ[i] = (u64 *)calloc(ctxQ->N * ctxQ->nl, sizeof(u64));
    for (int i = 0; i < 3; i++) e->poolP[i] = (u64 *)calloc(ctxP->N * ctxP->nl, sizeof(u64));
    for (int i = 0; i < 6; i++) e->ringpool[i] = (u64 *)calloc(ctxQ->N * ctxQ->nl, sizeof(u64));
    return e;
}
API void orc_ckks_eval_free(orc_ckks_eval *e) {
    if (!e) return;
    orc_extender_free(e->ext); orc_decomposer_free(e->dec);
    for (int i = 0; i < 4; i++) free(e->poolQ[i]);
    for (int i = 0; i < 3; i++) free(e->poolP[i]);
    for (int i = 0; i < 6; i++) free(e->ringpool[i]);
    free(e);
}

/* decomposeAndSplitNTT, ckks/evaluator.go:1561-1591 */
static void ckks_decompose_and_split_ntt(orc_ckks_eval *e, int level, int beta, const u64 *c2ntt, const u64 *c2inv,
                                         u64 *c2QiQ, u64 *c2QiP) {
    const orc_ctx *Q = e->ctxQ, *P = e->ctxP;
    u64 N = Q->N;
    orc_decompose_and_split(e->dec, N, level, beta, c2inv, c2QiQ, c2QiP);
    int p0idxst = beta * e->alpha;
    int p0idxed = p0idxst + e->dec->xalpha[beta];
    for (int x = 0; x < level + 1; x++) {
        if (p0idxst <= x && x < p0idxed)
            memcpy(c2QiQ + x * N, c2ntt + x * N, sizeof(u64) * N);
        else
            orc_ntt_one(Q, x, c2QiQ + x * N, c2QiQ + x * N);
    }
    orc_ntt(P, P->nl, c2QiP, c2QiP);
}

/* switchKeysInPlace, ckks/evaluator.go:1475-1558.
 * evk layout: [beta_max][2][nQ+nP][N] (ckks/keygen.go:282-340).
 * p0, p1: outputs over level+1 limbs of Q (stride N). */
API void orc_ckks_switch_keys_in_place(orc_ckks_eval *e, int level, const u64 *cx, const u64 *evk, u64 *p0, u64 *p1) {
    const orc_ctx *Q = e->ctxQ, *P = e->ctxP;
    u64 N = Q->N;
    int nQP = Q->nl + P->nl;
    for (int i = 0; i < 4; i++) memset(e->poolQ[i], 0, sizeof(u64) * N * Q->nl);
    for (int i = 0; i < 3; i++) memset(e->poolP[i], 0, sizeof(u64) * N * P->nl);
    /* p0/p1 are eval.poolQ[1], poolQ[2] at every call site, i.e. zeroed */
    memset(p0, 0, sizeof(u64) * N * (level + 1));
    memset(p1, 0, sizeof(u64) * N * (level + 1));
    u64 *c2QiQ = e->poolQ[0], *c2QiP = e->poolP[0];
    u64 *pool2Q = p0, *pool2P = e->poolP[1];
    u64 *pool3Q = p1, *pool3P = e->poolP[2];
    u64 *c2 = e->poolQ[3];
    orc_invntt(Q, level + 1, cx, c2);
    u64 reduce = 0;
    int alpha = e->alpha;
    int beta = (level + 1 + alpha - 1) / alpha;
    for (int i = 0; i < beta; i++) {
        ckks_decompose_and_split_ntt(e, level, i, cx, c2, c2QiQ, c2QiP);
        const u64 *k0 = evk + ((u64)(i * 2 + 0) * nQP) * N;
        const u64 *k1 = evk + ((u64)(i * 2 + 1) * nQP) * N;
        orc_mulcoeffs_montgomery_and_add_nomod(Q, level + 1, k0, c2QiQ, pool2Q);
        orc_mulcoeffs_montgomery_and_add_nomod(Q, level + 1, k1, c2QiQ, pool3Q);
        for (int j = 0, ki = e->levels; j < P->nl; j++, ki++) {
            u64 pj = P->modulus[j], mp = P->mred[j];
            for (u64 y = 0; y < N; y++) {
                pool2P[j * N + y] += orc_mred(k0[ki * N + y], c2QiP[j * N + y], pj, mp);
                pool3P[j * N + y] += orc_mred(k1[ki * N + y], c2QiP[j * N + y], pj, mp);
            }
        }
        if ((reduce & 7) == 1) {
            orc_reduce(Q, level + 1, pool2Q, pool2Q);
            orc_reduce(Q, level + 1, pool3Q, pool3Q);
            orc_reduce(P, P->nl, pool2P, pool2P);
            orc_reduce(P, P->nl, pool3P, pool3P);
        }
        reduce++;
    }
    if (((reduce - 1) & 7) != 1) {
        orc_reduce(Q, level + 1, pool2Q, pool2Q);
        orc_reduce(Q, level + 1, pool3Q, pool3Q);
        orc_reduce(P, P->nl, pool2P, pool2P);
        orc_reduce(P, P->nl, pool3P, pool3P);
    }
    orc_moddown_splited_ntt_pq(e->ext, level, pool2Q, pool2P, pool2Q);
    orc_moddown_splited_ntt_pq(e->ext, level, pool3Q, pool3P, pool3Q);
}

/* MulRelin, ckks/evaluator.go:1016-1133, ciphertext x ciphertext case with an
 * evaluation key (degree 1 x degree 1 -> degree 1).  ct = 2 polys of
 * level+1 limbs each, flat [2][level+1][N].  `square` selects the el0==el1
 * branch (:1080-1085). */
API void orc_ckks_mul_relin(orc_ckks_eval *e, int level, const u64 *ct0, const u64 *ct1, const u64 *evk, u64 *out) {
    const orc_ctx *Q = e->ctxQ;
    u64 N = Q->N;
    int nl = level + 1;
    u64 sz = (u64)nl * N;
    u64 *c00 = e->ringpool[0], *c01 = e->ringpool[1];
    u64 *c0 = e->ringpool[2], *c1 = e->ringpool[3], *c2 = e->ringpool[4];
    orc_mform_poly(Q, nl, ct0, c00);
    orc_mform_poly(Q, nl, ct0 + sz, c01);
    if (ct0 == ct1) {
        orc_mulcoeffs_montgomery(Q, nl, c00, ct1, c0);
        orc_mulcoeffs_montgomery(Q, nl, c00, ct1 + sz, c1);
        orc_add(Q, nl, c1, c1, c1);
        orc_mulcoeffs_montgomery(Q, nl, c01, ct1 + sz, c2);
    } else {
        orc_mulcoeffs_montgomery(Q, nl, c00, ct1, c0);
        orc_mulcoeffs_montgomery(Q, nl, c00, ct1 + sz, c1);
        orc_mulcoeffs_montgomery_and_add(Q, nl, c01, ct1, c1);
        orc_mulcoeffs_montgomery(Q, nl, c01, ct1 + sz, c2);
    }
    orc_ckks_switch_keys_in_place(e, level, c2, evk, e->poolQ[1], e->poolQ[2]);
    orc_add(Q, nl, c0, e->poolQ[1], out);
    orc_add(Q, nl, c1, e->poolQ[2], out + sz);
}

/* Rescale, ckks/evaluator.go:933-968: one iteration of the loop body
 * (:955-960) per `nb`; the scale/threshold test is host metadata.  ct is
 * [2][nl][N] in, and is compacted to [2][nl-nb][N] on return. */
API void orc_ckks_rescale(orc_ckks_eval *e, int nl, u64 *ct, int nb) {
    const orc_ctx *Q = e->ctxQ;
    u64 N = Q->N;
    for (int k = 0; k < nb; k++) {
        int cur = nl - k;
        orc_div_round_by_last_modulus_ntt(Q, cur, ct);
        orc_div_round_by_last_modulus_ntt(Q, cur, ct + (u64)nl * N);
    }
    memmove(ct + (u64)(nl - nb) * N, ct + (u64)nl * N, sizeof(u64) * N * (nl - nb));
}

/* permuteNTT, ckks/evaluator.go:1452-1472 (RotateColumns with a direct key
 * :1220, Conjugate :1449).  ct/out are [2][level+1][N]. */
API void orc_ckks_permute_ntt(orc_ckks_eval *e, int level, const u64 *ct, const u64 *index, const u64 *evk, u64 *out) {
    const orc_ctx *Q = e->ctxQ;
    u64 N = Q->N;
    int nl = level + 1;
    u64 sz = (u64)nl * N;
    u64 *el0 = e->ringpool[0], *el1 = e->ringpool[1];
    orc_permute_ntt_with_index(N, nl, ct, index, el0);
    orc_permute_ntt_with_index(N, nl, ct + sz, index, el1);
    orc_ckks_switch_keys_in_place(e, level, el1, evk, e->poolQ[1], e->poolQ[2]);
    orc_add(Q, nl, el0, e->poolQ[1], out);
    memcpy(out + sz, e->poolQ[2], sizeof(u64) * sz);
}

/* RotateHoisted precomputation, ckks/evaluator.go:1252-1275: c2InvNTT = InvNTT(ct.value[1]) and, for every
 * digit, decomposeAndSplitNTT into full-size Q / P polys.  qdec: [beta][nQ][N] (limbs above `level` stay
 * zero, as in a fresh contextQ.NewPoly()), pdec: [beta][nP][N]. */
API void orc_ckks_hoist(orc_ckks_eval *e, int level, const u64 *ct, u64 *qdec, u64 *pdec) {
    const orc_ctx *Q = e->ctxQ, *P = e->ctxP;
    u64 N = Q->N;
    int nl = level + 1;
    const u64 *c2ntt = ct + (u64)nl * N;
    u64 *c2inv = e->poolQ[3];
    orc_invntt(Q, nl, c2ntt, c2inv);
    int beta = (nl + e->alpha - 1) / e->alpha;
    memset(qdec, 0, sizeof(u64) * (u64)beta * Q->nl * N);
    memset(pdec, 0, sizeof(u64) * (u64)beta * P->nl * N);
    for (int i = 0; i < beta; i++)
        ckks_decompose_and_split_ntt(e, level, i, c2ntt, c2inv, qdec + (u64)i * Q->nl * N, pdec + (u64)i * P->nl * N);
}

/* switchKeyHoisted, ckks/evaluator.go:1291-1392 (ct0 != ctOut branch).  index = permuteNTTLeftIndex[k],
 * evk = evakeyRotColLeft[k]; ct, out: [2][level+1][N]. */
API void orc_ckks_switch_key_hoisted(orc_ckks_eval *e, int level, const u64 *ct, const u64 *qdec, const u64 *pdec,
                                     const u64 *index, const u64 *evk, u64 *out) {
    const orc_ctx *Q = e->ctxQ, *P = e->ctxP;
    u64 N = Q->N;
    int nl = level + 1, nQP = Q->nl + P->nl;
    u64 sz = (u64)nl * N;
    orc_permute_ntt_with_index(N, nl, ct, index, e->ringpool[0]);
    memcpy(out, e->ringpool[0], sizeof(u64) * sz); /* CopyLvl :1320 */
    for (int i = 0; i < 4; i++) memset(e->poolQ[i], 0, sizeof(u64) * N * Q->nl);
    for (int i = 0; i < 3; i++) memset(e->poolP[i], 0, sizeof(u64) * N * P->nl);
    u64 *c2QiQPermute = e->poolQ[0], *c2QiPPermute = e->poolP[0];
    u64 *pool2Q = e->poolQ[1], *pool2P = e->poolP[1];
    u64 *pool3Q = e->poolQ[2], *pool3P = e->poolP[2];
    u64 reduce = 0;
    int beta = (nl + e->alpha - 1) / e->alpha;
    for (int i = 0; i < beta; i++) {
        orc_permute_ntt_with_index(N, Q->nl, qdec + (u64)i * Q->nl * N, index, c2QiQPermute);
        orc_permute_ntt_with_index(N, P->nl, pdec + (u64)i * P->nl * N, index, c2QiPPermute);
        const u64 *k0 = evk + ((u64)(i * 2 + 0) * nQP) * N;
        const u64 *k1 = evk + ((u64)(i * 2 + 1) * nQP) * N;
        orc_mulcoeffs_montgomery_and_add_nomod(Q, nl, k0, c2QiQPermute, pool2Q);
        orc_mulcoeffs_montgomery_and_add_nomod(Q, nl, k1, c2QiQPermute, pool3Q);
        for (int j = 0, ki = e->levels; j < P->nl; j++, ki++) {
            u64 pj = P->modulus[j], mp = P->mred[j];
            for (u64 y = 0; y < N; y++) {
                pool2P[j * N + y] += orc_mred(k0[ki * N + y], c2QiPPermute[j * N + y], pj, mp);
                pool3P[j * N + y] += orc_mred(k1[ki * N + y], c2QiPPermute[j * N + y], pj, mp);
            }
        }
        if ((reduce & 7) == 1) {
            orc_reduce(Q, nl, pool2Q, pool2Q);
            orc_reduce(Q, nl, pool3Q, pool3Q);
            orc_reduce(P, P->nl, pool2P, pool2P);
            orc_reduce(P, P->nl, pool3P, pool3P);
        }
        reduce++;
    }
    if (((reduce - 1) & 7) != 1) {
        orc_reduce(Q, nl, pool2Q, pool2Q);
        orc_reduce(Q, nl, pool3Q, pool3Q);
        orc_reduce(P, P->nl, pool2P, pool2P);
        orc_reduce(P, P->nl, pool3P, pool3P);
    }
    orc_moddown_splited_ntt_pq(e->ext, level, pool2Q, pool2P, pool2Q);
    orc_moddown_splited_ntt_pq(e->ext, level, pool3Q, pool3P, pool3Q);
    orc_add(Q, nl, out, pool2Q, out);
    memcpy(out + sz, pool3Q, sizeof(u64) * sz);
}

/* SwitchKeys, ckks/evaluator.go:1176-1189 */
API void orc_ckks_switch_keys(orc_ckks_eval *e, int level, const u64 *ct, const u64 *evk, u64 *out) {
    const orc_ctx *Q = e->ctxQ;
    int nl = level + 1;
    u64 sz = (u64)nl * Q->N;
    orc_ckks_switch_keys_in_place(e, level, ct + sz, evk, e->poolQ[1], e->poolQ[2]);
    orc_add(Q, nl, ct, e->poolQ[1], out);
    memcpy(out + sz, e->poolQ[2], sizeof(u64) * sz);
}

/* ---------------------------------------------------------------------- */
/* bfv/evaluator.go hot ops                                               */
/* ---------------------------------------------------------------------- */

typedef struct {
    const orc_ctx *Q, *QMul, *P, *QP;
    orc_extender *q1q2; /* baseconverterQ1Q2 = NewFastBasisExtender(q, qm)  bfv/evaluator.go:95 */
    orc_extender *q1p;  /* baseconverterQ1P  = NewFastBasisExtender(q, p)   :86 */
    orc_decomposer *dec;
    int alpha, beta;
    u64 t;
    u64 *phalf_qmul, *phalf_q; /* pHalf = QMul>>1 (:98) reduced mod each prime (host big.Int work) */
} orc_bfv_eval;

/* NewEvaluator, bfv/evaluator.go:62-104 */
API orc_bfv_eval *orc_bfv_eval_new(const orc_ctx *Q, const orc_ctx *QMul, const orc_ctx *P, const orc_ctx *QP, u64 t,
                                   const u64 *phalf_qmul, const u64 *phalf_q) {
    orc_bfv_eval *e = (orc_bfv_eval *)calloc(1, sizeof(orc_bfv_eval));
    e->Q = Q; e->QMul = QMul; e->P = P; e->QP = QP;
    e->q1q2 = orc_extender_new(Q, QMul);
    e->q1p = orc_extender_new(Q, P);
    e->dec = orc_decomposer_new(Q->modulus, Q->nl, P->modulus, P->nl);
    e->alpha = P->nl;
    e->beta = (Q->nl + P->nl - 1) / P->nl; /* bfv/params.go: beta = ceil(len(Qi)/alpha) */
    e->t = t;
    e->phalf_qmul = (u64 *)malloc(sizeof(u64) * QMul->nl);
    e->phalf_q = (u64 *)malloc(sizeof(u64) * Q->nl);
    memcpy(e->phalf_qmul, phalf_qmul, sizeof(u64) * QMul->nl);
    memcpy(e->phalf_q, phalf_q, sizeof(u64) * Q->nl);
    return e;
}
API void orc_bfv_eval_free(orc_bfv_eval *e) {
    if (!e) return;
    orc_extender_free(e->q1q2); orc_extender_free(e->q1p); orc_decomposer_free(e->dec);
    free(e->phalf_qmul); free(e->phalf_q); free(e);
}

/* tensorAndRescale, bfv/evaluator.go:278-464, degree 1 x degree 1 (:327-367; squaring when ct0 == ct1).
 * ct0, ct1: [2][nQ][N] coefficient domain.  out: [3][nQ][N]. */
API void orc_bfv_tensor_and_rescale(orc_bfv_eval *e, const u64 *ct0, const u64 *ct1, u64 *out) {
    const orc_ctx *Q = e->Q, *M = e->QMul;
    u64 N = Q->N;
    int nQ = Q->nl, nM = M->nl;
    int levelQ = nQ - 1, levelM = nM - 1;
    u64 szQ = (u64)nQ * N, szM = (u64)nM * N;
    u64 *c0Q1 = (u64 *)malloc(sizeof(u64) * szQ * 2), *c0Q2 = (u64 *)malloc(sizeof(u64) * szM * 2);
    u64 *c1Q1 = (u64 *)malloc(sizeof(u64) * szQ * 2), *c1Q2 = (u64 *)malloc(sizeof(u64) * szM * 2);
    u64 *c2Q1 = (u64 *)calloc(szQ * 3, sizeof(u64)), *c2Q2 = (u64 *)calloc(szM * 3, sizeof(u64));
    u64 *c00Q = (u64 *)malloc(sizeof(u64) * szQ), *c00Q2 = (u64 *)malloc(sizeof(u64) * szM);
    u64 *c01Q = (u64 *)malloc(sizeof(u64) * szQ), *c01P = (u64 *)malloc(sizeof(u64) * szM);
    int square = (ct0 == ct1);
    for (int i = 0; i < 2; i++) { /* :298-303 */
        orc_modup_split_qp(e->q1q2, levelQ, ct0 + i * szQ, c0Q2 + i * szM);
        orc_ntt(Q, nQ, ct0 + i * szQ, c0Q1 + i * szQ);
        orc_ntt(M, nM, c0Q2 + i * szM, c0Q2 + i * szM);
    }
    if (!square) { /* :305-313 */
        for (int i = 0; i < 2; i++) {
            orc_modup_split_qp(e->q1q2, levelQ, ct1 + i * szQ, c1Q2 + i * szM);
            orc_ntt(Q, nQ, ct1 + i * szQ, c1Q1 + i * szQ);
            orc_ntt(M, nM, c1Q2 + i * szM, c1Q2 + i * szM);
        }
    }
    orc_mform_poly(Q, nQ, c0Q1, c00Q);          /* :327-331 */
    orc_mform_poly(M, nM, c0Q2, c00Q2);
    orc_mform_poly(Q, nQ, c0Q1 + szQ, c01Q);
    orc_mform_poly(M, nM, c0Q2 + szM, c01P);
    if (square) { /* :334-349 */
        orc_mulcoeffs_montgomery(Q, nQ, c00Q, c0Q1, c2Q1);
        orc_mulcoeffs_montgomery(M, nM, c00Q2, c0Q2, c2Q2);
        orc_mulcoeffs_montgomery(Q, nQ, c00Q, c0Q1 + szQ, c2Q1 + szQ);
        orc_mulcoeffs_montgomery(M, nM, c00Q2, c0Q2 + szM, c2Q2 + szM);
        orc_add_nomod(Q, nQ, c2Q1 + szQ, c2Q1 + szQ, c2Q1 + szQ);
        orc_add_nomod(M, nM, c2Q2 + szM, c2Q2 + szM, c2Q2 + szM);
        orc_mulcoeffs_montgomery(Q, nQ, c01Q, c0Q1 + szQ, c2Q1 + 2 * szQ);
        orc_mulcoeffs_montgomery(M, nM, c01P, c0Q2 + szM, c2Q2 + 2 * szM);
    } else { /* :352-367 */
        orc_mulcoeffs_montgomery(Q, nQ, c00Q, c1Q1, c2Q1);
        orc_mulcoeffs_montgomery(M, nM, c00Q2, c1Q2, c2Q2);
        orc_mulcoeffs_montgomery(Q, nQ, c00Q, c1Q1 + szQ, c2Q1 + szQ);
        orc_mulcoeffs_montgomery(M, nM, c00Q2, c1Q2 + szM, c2Q2 + szM);
        orc_mulcoeffs_montgomery_and_add_nomod(Q, nQ, c01Q, c1Q1, c2Q1 + szQ);
        orc_mulcoeffs_montgomery_and_add_nomod(M, nM, c01P, c1Q2, c2Q2 + szM);
        orc_mulcoeffs_montgomery(Q, nQ, c01Q, c1Q1 + szQ, c2Q1 + 2 * szQ);
        orc_mulcoeffs_montgomery(M, nM, c01P, c1Q2 + szM, c2Q2 + 2 * szM);
    }
    u64 tvec[64];
    for (int i = 0; i < nQ; i++) tvec[i] = e->t;
    for (int i = 0; i < 3; i++) { /* :423-463 */
        orc_invntt(Q, nQ, c2Q1 + i * szQ, c2Q1 + i * szQ);
        orc_invntt(M, nM, c2Q2 + i * szM, c2Q2 + i * szM);
        orc_moddown_splited_qp(e->q1q2, levelQ, levelM, c2Q1 + i * szQ, c2Q2 + i * szM, c2Q2 + i * szM);
        orc_add_scalar(M, nM, c2Q2 + i * szM, e->phalf_qmul);
        orc_modup_split_pq(e->q1q2, levelM, c2Q2 + i * szM, out + i * szQ);
        orc_sub_scalar(Q, nQ, out + i * szQ, e->phalf_q);
        orc_mul_scalar(Q, nQ, out + i * szQ, tvec, out + i * szQ);
    }
    free(c0Q1); free(c0Q2); free(c1Q1); free(c1Q2); free(c2Q1); free(c2Q2);
    free(c00Q); free(c00Q2); free(c01Q); free(c01P);
}

/* switchKeys, bfv/evaluator.go:736-813.  cx: [nQ][N] coefficient domain.  evk: [beta][2][nQ+nP][N].
 * p0, p1: [nQ+nP][N] work polys; the results are their first nQ limbs. */
API void orc_bfv_switch_keys_core(orc_bfv_eval *e, const u64 *cx, const u64 *evk, u64 *p0, u64 *p1) {
    const orc_ctx *Q = e->Q, *K = e->QP;
    u64 N = K->N;
    int nQP = K->nl;
    int level = Q->nl - 1;
    u64 *c2Qi = (u64 *)calloc((u64)nQP * N, sizeof(u64));
    u64 *c2 = (u64 *)calloc((u64)nQP * N, sizeof(u64));
    u64 *c2QiNtt = (u64 *)malloc(sizeof(u64) * N);
    memset(p0, 0, sizeof(u64) * nQP * N);
    memset(p1, 0, sizeof(u64) * nQP * N);
    orc_ntt(Q, Q->nl, cx, c2);
    u64 reduce = 0;
    for (int i = 0; i < e->beta; i++) {
        int p0idxst = i * e->alpha;
        int p0idxed = p0idxst + e->dec->xalpha[i];
        orc_decompose(e->dec, N, level, i, cx, c2Qi);
        const u64 *k0 = evk + ((u64)(i * 2 + 0) * nQP) * N;
        const u64 *k1 = evk + ((u64)(i * 2 + 1) * nQP) * N;
        for (int x = 0; x < nQP; x++) {
            u64 qi = K->modulus[x], mp = K->mred[x];
            if (p0idxst <= x && x < p0idxed)
                memcpy(c2QiNtt, c2 + (u64)x * N, sizeof(u64) * N);
            else
                orc_ntt_limb(c2Qi + (u64)x * N, c2QiNtt, N, K->ntt_psi[x], qi, mp, K->bred[x]);
            for (u64 y = 0; y < N; y++) {
                p0[x * N + y] += orc_mred(k0[x * N + y], c2QiNtt[y], qi, mp);
                p1[x * N + y] += orc_mred(k1[x * N + y], c2QiNtt[y], qi, mp);
            }
        }
        if ((reduce & 7) == 7) {
            orc_reduce(K, nQP, p0, p0);
            orc_reduce(K, nQP, p1, p1);
        }
        reduce++;
    }
    if (((reduce - 1) & 7) != 7) {
        orc_reduce(K, nQP, p0, p0);
        orc_reduce(K, nQP, p1, p1);
    }
    orc_invntt(K, nQP, p0, p0);
    orc_invntt(K, nQP, p1, p1);
    orc_moddown_pq(e->q1p, level, p0, p0);
    orc_moddown_pq(e->q1p, level, p1, p1);
    free(c2Qi); free(c2); free(c2QiNtt);
}

/* relinearize, bfv/evaluator.go:480-500, degree 2 -> 1.  ct: [3][nQ][N], evk = evakey[0]. out: [2][nQ][N] */
API void orc_bfv_relinearize(orc_bfv_eval *e, const u64 *ct, const u64 *evk, u64 *out) {
    const orc_ctx *Q = e->Q;
    u64 N = Q->N, szQ = (u64)Q->nl * N;
    u64 *p0 = (u64 *)malloc(sizeof(u64) * e->QP->nl * N), *p1 = (u64 *)malloc(sizeof(u64) * e->QP->nl * N);
    if (out != ct) memcpy(out, ct, sizeof(u64) * 2 * szQ);
    orc_bfv_switch_keys_core(e, ct + 2 * szQ, evk, p0, p1);
    orc_add(Q, Q->nl, out, p0, out);
    orc_add(Q, Q->nl, out + szQ, p1, out + szQ);
    free(p0); free(p1);
}

/* SwitchKeys, bfv/evaluator.go:540-558 */
API void orc_bfv_switch_keys(orc_bfv_eval *e, const u64 *ct, const u64 *evk, u64 *out) {
    const orc_ctx *Q = e->Q;
    u64 N = Q->N, szQ = (u64)Q->nl * N;
    u64 *p0 = (u64 *)malloc(sizeof(u64) * e->QP->nl * N), *p1 = (u64 *)malloc(sizeof(u64) * e->QP->nl * N);
    orc_bfv_switch_keys_core(e, ct + szQ, evk, p0, p1);
    orc_add(Q, Q->nl, ct, p0, out);
    memcpy(out + szQ, p1, sizeof(u64) * szQ);
    free(p0); free(p1);
}

/* permute, bfv/evaluator.go:711-733 (RotateColumns :578 with a direct key, RotateRows :669) */
API void orc_bfv_permute(orc_bfv_eval *e, const u64 *ct, u64 gen, const u64 *evk, u64 *out) {
    const orc_ctx *Q = e->Q;
    u64 N = Q->N, szQ = (u64)Q->nl * N;
    u64 *el = (u64 *)malloc(sizeof(u64) * 2 * szQ);
    u64 *p0 = (u64 *)malloc(sizeof(u64) * e->QP->nl * N), *p1 = (u64 *)malloc(sizeof(u64) * e->QP->nl * N);
    orc_permute(Q, Q->nl, ct, gen, el);
    orc_permute(Q, Q->nl, ct + szQ, gen, el + szQ);
    orc_bfv_switch_keys_core(e, el + szQ, evk, p0, p1);
    orc_add(Q, Q->nl, el, p0, out);
    memcpy(out + szQ, p1, sizeof(u64) * szQ);
    free(el); free(p0); free(p1);
}

/* ---------------------------------------------------------------------- */
/* ring/float128.go + SimpleScaler, ring/ring_scaling.go:166-300           */
/* ---------------------------------------------------------------------- */
/* Double-double arithmetic exactly as ring/float128.go writes it: every   */
/* operation is one IEEE binary64 RN operation (the file is built with     */
/* -ffp-contract=off; Go's amd64 back end at its default GOAMD64=v1 level  */
/* does not fuse multiply-adds either).                                    */

typedef struct { double v[2]; } f128;

/* Go's uint64(float64) on amd64: CVTTSD2SQ below 2^63 (negative inputs wrap as two's complement), the upper half
 * handled by converting f - 2^63 and flipping the top bit.  CVTTSD2SQ returns 0x8000000000000000 out of range; that
 * is spelled out so that out-of-contract inputs do not depend on C's undefined conversion. */
static inline int64_t x86_cvttsd2sq(double f) {
    return (f >= -9223372036854775808.0 && f < 9223372036854775808.0) ? (int64_t)f : (int64_t)0x8000000000000000ull;
}
static inline u64 go_f64_to_u64(double f) {
    if (f < 9223372036854775808.0) return (u64)x86_cvttsd2sq(f);
    return (u64)x86_cvttsd2sq(f - 9223372036854775808.0) ^ 0x8000000000000000ull;
}
/* math.Round: half away from zero */
static inline double go_round(double x) {
    if (!(x > -4503599627370496.0 && x < 4503599627370496.0)) return x; /* already integral (or NaN) */
    double t = (double)(int64_t)x;
    double d = x - t;
    if (d >= 0.5) t += 1.0;
    else if (d <= -0.5) t -= 1.0;
    return t;
}
static inline f128 f128_set_u53(u64 i) { f128 r = {{(double)i, 0.0}}; return r; }                     /* float128.go:33-37 */
static inline f128 f128_set_u64(u64 i) { f128 r = {{(double)(i >> 12), (double)(i & 0xfff) / 4096.0}}; return r; } /* :44-48 */
static inline u64 f128_to_u53(f128 f) { return go_f64_to_u64(f.v[0]); }                                /* :72-74 */
static inline u64 f128_to_u64(f128 f) {                                                                /* :80-82 */
    double a = f.v[0] * 4096.0;
    u64 ai = go_f64_to_u64(a);
    return ai + go_f64_to_u64(go_round((a - (double)ai) + f.v[1] * 4096.0));
}
static inline void two_sum(double a, double b, double *s, double *e) {   /* :85-90 */
    *s = a + b;
    double bb = *s - a;
    *e = (a - (*s - bb)) + (b - bb);
}
static inline void quick_two_sum(double a, double b, double *s, double *e) { /* :93-97 */
    *s = a + b;
    *e = b - (*s - a);
}
static inline void two_diff(double a, double b, double *s, double *e) {  /* :111-116 */
    *s = a - b;
    double bb = *s - a;
    *e = (a - (*s - bb)) - (b + bb);
}
static inline f128 f128_add(f128 a, f128 b) {                            /* :100-108 */
    double s1, s2, t1, t2;
    two_sum(a.v[0], b.v[0], &s1, &s2);
    two_sum(a.v[1], b.v[1], &t1, &t2);
    s2 += t1;
    quick_two_sum(s1, s2, &s1, &s2);
    s2 += t2;
    f128 f;
    quick_two_sum(s1, s2, &f.v[0], &f.v[1]);
    return f;
}
static inline void f_split(double a, double *hi, double *lo) {           /* :132-137 */
    double temp = 134217729.0 * a;
    *hi = temp - (temp - a);
    *lo = a - *hi;
}
static inline void two_prod(double a, double b, double *p, double *e) {  /* :140-146 */
    *p = a * b;
    double ahi, alo, bhi, blo;
    f_split(a, &ahi, &alo);
    f_split(b, &bhi, &blo);
    *e = ((ahi * bhi - *p) + ahi * blo + alo * bhi) + alo * blo;
}
static inline f128 f128_mul(f128 a, f128 b) {                            /* :149-154 */
    double p1, p2;
    two_prod(a.v[0], b.v[0], &p1, &p2);
    p2 += a.v[0] * b.v[1] + a.v[1] * b.v[0];
    f128 f;
    quick_two_sum(p1, p2, &f.v[0], &f.v[1]);
    return f;
}
static inline f128 f128_div(f128 a, f128 b) {                            /* :157-217 (the live statements) */
    double q1, p1, p2, p3, p4, v1, v2, r, t0, t1;
    q1 = a.v[0] / b.v[0];
    two_prod(q1, b.v[0], &p1, &p2);
    p2 += q1 * b.v[1];
    t0 = p1 + p2;
    t1 = p2 - (t0 - p1);
    two_diff(a.v[0], t0, &p3, &p4);
    two_diff(a.v[1], t1, &v1, &v2);
    p4 += v1;
    quick_two_sum(p3, p4, &p3, &p4);
    p4 += v2;
    r = (p3 + p4) / b.v[0];
    f128 f;
    f.v[0] = q1 + r;
    f.v[1] = r - (f.v[0] - q1);
    return f;
}

/* exported one-op probes so the tests can compare the double-double ops against an independent restatement */
API void orc_f128_op(int op, const double a[2], const double b[2], double out[2]) {
    f128 x = {{a[0], a[1]}}, y = {{b[0], b[1]}}, r;
    r = op == 0 ? f128_add(x, y) : op == 1 ? f128_mul(x, y) : f128_div(x, y);
    out[0] = r.v[0];
    out[1] = r.v[1];
}
API u64 orc_f128_to_u64(const double a[2]) { f128 x = {{a[0], a[1]}}; return f128_to_u64(x); }

typedef struct {
    int nl;
    u64 N, t;
    int pow2;
    u64 add_param, mul_param; /* reducealgoAddParam / reducealgoMulParam (:183-184) */
    u64 *q, *wi;
    double *ti; /* [nl][2] */
} orc_scaler;

/* NewSimpleScaler, ring_scaling.go:188-262 */
API orc_scaler *orc_scaler_new(u64 t, const orc_ctx *c) {
    orc_scaler *s = (orc_scaler *)calloc(1, sizeof(orc_scaler));
    s->nl = c->nl; s->N = c->N; s->t = t;
    s->q = (u64 *)malloc(sizeof(u64) * c->nl);
    s->wi = (u64 *)malloc(sizeof(u64) * c->nl);
    s->ti = (double *)malloc(sizeof(double) * 2 * c->nl);
    s->pow2 = (t & (t - 1)) == 0 && t != 0;
    u64 bred_t[2] = {0, 0};
    if (s->pow2) {
        s->add_param = t - 1; s->mul_param = t - 1;            /* :203-204 */
    } else {
        orc_bred_params(t, bred_t);
        s->add_param = bred_t[0];                               /* :216 */
        s->mul_param = orc_mred_params(t);                      /* :217 */
    }
    for (int i = 0; i < c->nl; i++) {
        u64 qi = c->modulus[i];
        s->q[i] = qi;
        /* QiBarre = (Q/qi)^-1 mod qi (:241-247); qi prime */
        u64 star = 1;
        for (int k = 0; k < c->nl; k++) if (k != i) star = mulmod(star, c->modulus[k] % qi, qi);
        u64 barre = powmod(star, qi - 2, qi);
        f128 tmp = f128_div(f128_set_u53(t), f128_set_u64(qi)); /* :249 */
        tmp = f128_mul(tmp, f128_set_u64(barre));               /* :251 */
        s->wi[i] = f128_to_u53(tmp);                            /* :254 */
        if (!s->pow2) s->wi[i] = orc_mform(s->wi[i], t, bred_t);/* :257-259 */
        u64 barre_t = mulmod(barre, t % qi, qi);                /* :261-262 */
        f128 ti = f128_div(f128_set_u64(barre_t), f128_set_u64(qi)); /* :264 */
        s->ti[2 * i] = ti.v[0];
        s->ti[2 * i + 1] = ti.v[1];
    }
    return s;
}
API void orc_scaler_free(orc_scaler *s) {
    if (!s) return;
    free(s->q); free(s->wi); free(s->ti); free(s);
}
API void orc_scaler_params(const orc_scaler *s, u64 *wi, double *ti) {
    memcpy(wi, s->wi, sizeof(u64) * s->nl);
    memcpy(ti, s->ti, sizeof(double) * 2 * s->nl);
}

/* SimpleScaler.Scale, ring_scaling.go:271-300: p1 has the context's nl limbs; every one of p2's nl2 limbs gets the result */
API void orc_scaler_scale(const orc_scaler *s, const u64 *p1, u64 *p2, int nl2) {
    const u64 N = s->N, t = s->t;
    for (u64 i = 0; i < N; i++) {
        u64 a = 0;
        f128 b = {{0.0, 0.0}};
        for (int j = 0; j < s->nl; j++) {
            u64 x = p1[(u64)j * N + i];
            if (s->pow2) a += (s->wi[j] * x) & s->add_param;                 /* :206-208 */
            else a += orc_mred(s->wi[j], x, t, s->mul_param);                /* :219-230 */
            f128 tj = {{s->ti[2 * j], s->ti[2 * j + 1]}};
            b = f128_add(b, f128_mul(tj, f128_set_u64(x)));                  /* :288 */
        }
        a += f128_to_u64(b);                                                 /* :291 */
        if (s->pow2) a &= s->mul_param;                                      /* :210-212 */
        else {                                                               /* :232-243 */
            u64 s0 = hi64(a, s->add_param);
            u64 r = a - s0 * t;
            if (r >= t) r -= t;
            a = r;
        }
        for (int j = 0; j < nl2; j++) p2[(u64)j * N + i] = a;               /* :295-297 */
    }
}
