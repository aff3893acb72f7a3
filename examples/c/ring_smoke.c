/* ring_smoke.c -- the drop-in boundary from plain C: compiled with gcc against include/lattigpu.h and linked to
 * liblattigpu.so, no Python and no C++ in between (what a cgo binding sees).
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/c/ring_smoke.c -Llattigo-fhe-by-go_b200/lib -llattigpu \
 *       -Wl,-rpath,$PWD/lattigo-fhe-by-go_b200/lib -o examples/c/ring_smoke
 *
 * Runs, on the CUDA device given as argv[1] (default 0):
 *   1. NTT -> InvNTT on a random polynomial of a 6-limb ring, N = 2^13: must return the input (ring/ntt.go:53-139);
 *   2. switchKeysInPlace (ckks/evaluator.go:1475-1558) twice on the same input with a random key: the two results must
 *      be identical, canonical (every word below its modulus) and different from zero;
 *   3. linearity of the key switch in its input: KS(a) + KS(b) == KS(a + b) -- not an identity of the reference
 *      (the basis extensions round), so it is only REPORTED as the number of differing words, expected small;
 *   4. an operand created on another ring degree must be rejected with LG_ERR_ARG and a message.
 * Exit code 0 = all hard checks passed. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lattigpu.h"

#define CHECK(expr)                                                                 \
    do {                                                                            \
        int rc_ = (expr);                                                           \
        if (rc_ != LG_OK) {                                                         \
            fprintf(stderr, "%s:%d: %s -> %d: %s\n", __FILE__, __LINE__, #expr, rc_, lg_last_error()); \
            return 1;                                                               \
        }                                                                           \
    } while (0)

static uint64_t rng_state = 0x1A771C0ull;
static uint64_t splitmix64(void) {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void fill_uniform(uint64_t* dst, const uint64_t* moduli, int nl, uint64_t N) {
    int i;
    uint64_t x;
    for (i = 0; i < nl; ++i)
        for (x = 0; x < N; ++x) dst[(size_t)i * N + x] = splitmix64() % moduli[i];
}

int main(int argc, char** argv) {
    const uint64_t logN = 13, N = (uint64_t)1 << 13;
    enum { NQ = 6, NP = 1 };
    uint64_t Q[NQ], P[NP], primes30[NQ - 1];
    lg_ring *ringQ = NULL, *ringP = NULL, *ringSmall = NULL;
    lg_poly *a = NULL, *b = NULL, *t = NULL, *p0 = NULL, *p1 = NULL, *q0 = NULL, *q1 = NULL, *s0 = NULL, *s1 = NULL, *bad = NULL;
    lg_ckks_eval* ev = NULL;
    lg_swk* key = NULL;
    uint64_t *ha, *hb, *hk, *h0, *h1, *h2;
    const int beta = NQ; /* alpha = 1 */
    int i, ndev = 0, dev = argc > 1 ? atoi(argv[1]) : 0, rc;
    size_t words = (size_t)NQ * N, diff, nonzero;
    uint64_t x;

    printf("%s\n", lg_version());
    CHECK(lg_device_count(&ndev));
    CHECK(lg_set_device(dev));
    /* CKKS PN13QP218 (ckks/params.go:48-56): Q = 33 + 5 x 30 bits, P = 35 bits */
    CHECK(lg_generate_ntt_primes(33, logN, 1, Q));
    CHECK(lg_generate_ntt_primes(30, logN, NQ - 1, primes30));
    for (i = 1; i < NQ; ++i) Q[i] = primes30[i - 1];
    CHECK(lg_generate_ntt_primes(35, logN, 1, P));
    CHECK(lg_ring_create(N, NQ, Q, &ringQ));
    CHECK(lg_ring_create(N, NP, P, &ringP));
    ha = (uint64_t*)malloc(words * 8);
    hb = (uint64_t*)malloc(words * 8);
    h0 = (uint64_t*)malloc(words * 8);
    h1 = (uint64_t*)malloc(words * 8);
    h2 = (uint64_t*)malloc(words * 8);
    hk = (uint64_t*)malloc((size_t)beta * 2 * (NQ + NP) * N * 8);
    if (!ha || !hb || !h0 || !h1 || !h2 || !hk) return 2;
    fill_uniform(ha, Q, NQ, N);
    fill_uniform(hb, Q, NQ, N);
    {
        uint64_t QP[NQ + NP];
        int d, h;
        memcpy(QP, Q, sizeof(Q));
        memcpy(QP + NQ, P, sizeof(P));
        for (d = 0; d < beta; ++d)
            for (h = 0; h < 2; ++h) fill_uniform(hk + ((size_t)(d * 2 + h) * (NQ + NP)) * N, QP, NQ + NP, N);
    }
    CHECK(lg_poly_create(N, NQ, 1, &a));
    CHECK(lg_poly_create(N, NQ, 1, &b));
    CHECK(lg_poly_create(N, NQ, 1, &t));
    CHECK(lg_poly_upload(a, 0, 1, 0, NQ, ha, NULL));
    CHECK(lg_poly_upload(b, 0, 1, 0, NQ, hb, NULL));

    /* 1. NTT round trip */
    CHECK(lg_ring_ntt(ringQ, NQ, a, t, NULL));
    CHECK(lg_ring_invntt(ringQ, NQ, t, t, NULL));
    CHECK(lg_poly_download(t, 0, 1, 0, NQ, h0, NULL));
    if (memcmp(h0, ha, words * 8) != 0) {
        fprintf(stderr, "NTT -> InvNTT is not the identity\n");
        return 1;
    }
    printf("1. NTT -> InvNTT round trip: ok (%d limbs x %llu coefficients)\n", NQ, (unsigned long long)N);

    /* 2. key switch, twice */
    CHECK(lg_ckks_eval_create(ringQ, ringP, &ev));
    CHECK(lg_swk_create(N, beta, NQ + NP, hk, &key));
    CHECK(lg_poly_create(N, NQ, 1, &p0));
    CHECK(lg_poly_create(N, NQ, 1, &p1));
    CHECK(lg_poly_create(N, NQ, 1, &q0));
    CHECK(lg_poly_create(N, NQ, 1, &q1));
    CHECK(lg_ckks_switch_keys_in_place(ev, NQ - 1, a, key, p0, p1, NULL));
    CHECK(lg_ckks_switch_keys_in_place(ev, NQ - 1, a, key, q0, q1, NULL));
    CHECK(lg_poly_download(p0, 0, 1, 0, NQ, h0, NULL));
    CHECK(lg_poly_download(q0, 0, 1, 0, NQ, h1, NULL));
    if (memcmp(h0, h1, words * 8) != 0) {
        fprintf(stderr, "switchKeysInPlace is not deterministic\n");
        return 1;
    }
    nonzero = 0;
    for (i = 0; i < NQ; ++i)
        for (x = 0; x < N; ++x) {
            if (h0[(size_t)i * N + x] >= Q[i]) {
                fprintf(stderr, "switchKeysInPlace: word above its modulus\n");
                return 1;
            }
            nonzero += h0[(size_t)i * N + x] != 0;
        }
    if (nonzero < words / 2) {
        fprintf(stderr, "switchKeysInPlace: result is mostly zero\n");
        return 1;
    }
    printf("2. switchKeysInPlace: deterministic, canonical, %lu of %lu words non-zero: ok\n", (unsigned long)nonzero,
           (unsigned long)words);

    /* 3. KS(a) + KS(b) vs KS(a + b) */
    CHECK(lg_poly_create(N, NQ, 1, &s0));
    CHECK(lg_poly_create(N, NQ, 1, &s1));
    CHECK(lg_ckks_switch_keys_in_place(ev, NQ - 1, b, key, q0, q1, NULL));
    CHECK(lg_ring_add(ringQ, NQ, p0, q0, p0, NULL)); /* KS(a)[0] + KS(b)[0] */
    CHECK(lg_ring_add(ringQ, NQ, a, b, t, NULL));
    CHECK(lg_ckks_switch_keys_in_place(ev, NQ - 1, t, key, s0, s1, NULL));
    CHECK(lg_poly_download(p0, 0, 1, 0, NQ, h0, NULL));
    CHECK(lg_poly_download(s0, 0, 1, 0, NQ, h2, NULL));
    diff = 0;
    for (x = 0; x < words; ++x) {
        uint64_t q = Q[x / N], d = h0[x] >= h2[x] ? h0[x] - h2[x] : h2[x] - h0[x];
        if (d > q - d) d = q - d;
        if (d > 4) ++diff; /* the two differ by the rounding of the basis extensions only */
    }
    printf("3. KS(a) + KS(b) vs KS(a + b): %lu of %lu words differ by more than 4 (rounding of the basis extensions)\n",
           (unsigned long)diff, (unsigned long)words);

    /* 4. misuse is rejected, not executed */
    CHECK(lg_ring_create(N / 2, 1, P, &ringSmall));
    CHECK(lg_poly_create(N / 2, 1, 1, &bad));
    rc = lg_ring_ntt(ringQ, NQ, bad, t, NULL);
    if (rc != LG_ERR_ARG || strlen(lg_last_error()) == 0) {
        fprintf(stderr, "a polynomial of another degree was not rejected (rc = %d)\n", rc);
        return 1;
    }
    printf("4. degree mismatch rejected: \"%s\": ok\n", lg_last_error());
    CHECK(lg_stream_sync(NULL));
    printf("kernels launched: %llu\n", (unsigned long long)lg_launch_count());

    lg_poly_destroy(bad);
    lg_ring_destroy(ringSmall);
    lg_poly_destroy(s0);
    lg_poly_destroy(s1);
    lg_poly_destroy(q0);
    lg_poly_destroy(q1);
    lg_poly_destroy(p0);
    lg_poly_destroy(p1);
    lg_swk_destroy(key);
    lg_ckks_eval_destroy(ev);
    lg_poly_destroy(t);
    lg_poly_destroy(b);
    lg_poly_destroy(a);
    lg_ring_destroy(ringP);
    lg_ring_destroy(ringQ);
    free(ha);
    free(hb);
    free(h0);
    free(h1);
    free(h2);
    free(hk);
    printf("ring_smoke: ok\n");
    return 0;
}
