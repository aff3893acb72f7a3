// crp.cu -- utils.PRNG (utils/prng.go:11-72) and ring.CRPGenerator (ring/prng.go:11-103): the keyed BLAKE2b-512
// hash chain the dckks / dbfv protocols draw their common reference polynomials from, and the masked rejection
// sampling that turns the byte stream into a uniform polynomial.
//
// Host code by construction: digest k is the hash of the key block, the seed and ALL previous digests (Clock =
// Sum(nil) followed by Write(sum)), so the chain is strictly sequential (about 1.5 compressions per 64 output bytes)
// and the rejection sampling consumes it in order (the limb a word is tested against depends on how many words were
// accepted before it).  The finished polynomial is uploaded limb-major into a device handle; nothing here launches a
// kernel.  BLAKE2b follows RFC 7693 (the reference links golang.org/x/crypto/blake2b, go.mod:5, New512 = 64-byte
// digests, optional key of at most 64 bytes).
#include <string.h>

#include "capi_internal.hpp"

namespace {

struct Blake2b {
    u64 h[8];
    u64 t[2];
    uint8_t buf[128];
    size_t buflen;
    uint8_t key[64];
    size_t keylen;
};

const u64 B2B_IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                       0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
const uint8_t B2B_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

inline u64 rotr64(u64 x, int n) { return (x >> n) | (x << (64 - n)); }
inline u64 load_le64(const uint8_t* p) {
    u64 v;
    memcpy(&v, p, 8);  // x86-64 / aarch64 hosts are little-endian
    return v;
}

void b2b_compress(Blake2b& s, const uint8_t* block, bool last) {
    u64 m[16], v[16];
    for (int i = 0; i < 16; ++i) m[i] = load_le64(block + 8 * i);
    for (int i = 0; i < 8; ++i) {
        v[i] = s.h[i];
        v[8 + i] = B2B_IV[i];
    }
    v[12] ^= s.t[0];
    v[13] ^= s.t[1];
    if (last) v[14] = ~v[14];
#define B2B_G(a, b, c, d, x, y)         \
    v[a] = v[a] + v[b] + (x);           \
    v[d] = rotr64(v[d] ^ v[a], 32);     \
    v[c] = v[c] + v[d];                 \
    v[b] = rotr64(v[b] ^ v[c], 24);     \
    v[a] = v[a] + v[b] + (y);           \
    v[d] = rotr64(v[d] ^ v[a], 16);     \
    v[c] = v[c] + v[d];                 \
    v[b] = rotr64(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; ++r) {
        const uint8_t* g = B2B_SIGMA[r];
        B2B_G(0, 4, 8, 12, m[g[0]], m[g[1]])
        B2B_G(1, 5, 9, 13, m[g[2]], m[g[3]])
        B2B_G(2, 6, 10, 14, m[g[4]], m[g[5]])
        B2B_G(3, 7, 11, 15, m[g[6]], m[g[7]])
        B2B_G(0, 5, 10, 15, m[g[8]], m[g[9]])
        B2B_G(1, 6, 11, 12, m[g[10]], m[g[11]])
        B2B_G(2, 7, 8, 13, m[g[12]], m[g[13]])
        B2B_G(3, 4, 9, 14, m[g[14]], m[g[15]])
    }
#undef B2B_G
    for (int i = 0; i < 8; ++i) s.h[i] ^= v[i] ^ v[8 + i];
}

// hash.Reset(): parameter block for a 64-byte digest and the stored key; a key occupies the first block
void b2b_reset(Blake2b& s) {
    for (int i = 0; i < 8; ++i) s.h[i] = B2B_IV[i];
    s.h[0] ^= 0x01010000ull ^ ((u64)s.keylen << 8) ^ 64ull;
    s.t[0] = s.t[1] = 0;
    s.buflen = 0;
    memset(s.buf, 0, sizeof(s.buf));
    if (s.keylen > 0) {
        memcpy(s.buf, s.key, s.keylen);
        s.buflen = 128;
    }
}
// hash.Write(): a full buffer is compressed only when more input follows (the final block needs the last flag)
void b2b_update(Blake2b& s, const uint8_t* in, size_t len) {
    while (len > 0) {
        if (s.buflen == 128) {
            s.t[0] += 128;
            if (s.t[0] < 128) s.t[1]++;
            b2b_compress(s, s.buf, false);
            s.buflen = 0;
        }
        const size_t take = (128 - s.buflen) < len ? (128 - s.buflen) : len;
        memcpy(s.buf + s.buflen, in, take);
        s.buflen += take;
        in += take;
        len -= take;
    }
}
// hash.Sum(nil): finalises a COPY, the running state is untouched
void b2b_sum(const Blake2b& s, uint8_t out[64]) {
    Blake2b c = s;
    c.t[0] += c.buflen;
    if (c.t[0] < c.buflen) c.t[1]++;
    memset(c.buf + c.buflen, 0, 128 - c.buflen);
    b2b_compress(c, c.buf, true);
    memcpy(out, c.h, 64);  // little-endian words
}

}  // namespace

struct lg_prng {
    Blake2b st;
    u64 clock = 0;
    std::vector<uint8_t> seed;
    void step(uint8_t out[64]) {  // utils/prng.go:51-56
        b2b_sum(st, out);
        b2b_update(st, out, 64);
        ++clock;
    }
};

struct lg_crp {
    lg_prng prng;
    const lg_ring* ring = nullptr;
    std::vector<u64> masks;
    u64* stage = nullptr;  // pinned [nl][N]
    ~lg_crp() {
        if (stage) cudaFreeHost(stage);
    }
};

static int prng_init(lg_prng* p, const uint8_t* key, size_t keylen) {
    LG_REQUIRE(keylen <= 64, "blake2b: invalid key size");  // blake2b.New512 error
    LG_REQUIRE(key || keylen == 0, "NewPRNG: null key with non-zero length");
    p->st.keylen = keylen;
    memset(p->st.key, 0, sizeof(p->st.key));
    if (keylen) memcpy(p->st.key, key, keylen);
    b2b_reset(p->st);
    p->clock = 0;
    return LG_OK;
}
static int prng_set_clock(lg_prng* p, uint64_t n) {
    LG_REQUIRE(p->clock <= n, "error : cannot set prng clock to a previous state");  // utils/prng.go:62-64
    uint8_t tmp[64];
    while (p->clock != n) p->step(tmp);
    return LG_OK;
}

extern "C" {

int lg_prng_create(const uint8_t* key, size_t keylen, lg_prng** out) {
    LG_REQUIRE(out, "NewPRNG: null output");
    std::unique_ptr<lg_prng> p(new lg_prng());
    LG_TRY(prng_init(p.get(), key, keylen));
    *out = p.release();
    return LG_OK;
}
int lg_prng_destroy(lg_prng* p) {
    delete p;
    return LG_OK;
}
int lg_prng_seed(lg_prng* p, const uint8_t* seed, size_t len) {
    LG_REQUIRE(p && (seed || len == 0), "PRNG.Seed: null argument");
    b2b_reset(p->st);
    p->seed.assign(seed, seed + len);
    b2b_update(p->st, seed, len);
    p->clock = 0;
    return LG_OK;
}
uint64_t lg_prng_get_clock(const lg_prng* p) { return p ? p->clock : 0; }
int lg_prng_clock(lg_prng* p, uint8_t out[64]) {
    LG_REQUIRE(p && out, "PRNG.Clock: null argument");
    p->step(out);
    return LG_OK;
}
int lg_prng_set_clock(lg_prng* p, uint64_t n) {
    LG_REQUIRE(p, "PRNG.SetClock: null argument");
    return prng_set_clock(p, n);
}

int lg_crp_create(const uint8_t* key, size_t keylen, const lg_ring* ring, lg_crp** out) {
    LG_REQUIRE(ring && out, "NewCRPGenerator: null argument");
    std::unique_ptr<lg_crp> g(new lg_crp());
    LG_TRY(prng_init(&g->prng, key, keylen));
    g->ring = ring;
    g->masks.resize(ring->nl);
    for (int i = 0; i < ring->nl; ++i) {  // ring/prng.go:31-33: (1 << bits.Len64(qi)) - 1
        const int len = 64 - __builtin_clzll(ring->q[i]);
        g->masks[i] = len >= 64 ? ~0ull : ((1ull << len) - 1);
    }
    *out = g.release();
    return LG_OK;
}
int lg_crp_destroy(lg_crp* g) {
    delete g;
    return LG_OK;
}
int lg_crp_seed(lg_crp* g, const uint8_t* seed, size_t len) {
    LG_REQUIRE(g, "CRPGenerator.Seed: null argument");
    return lg_prng_seed(&g->prng, seed, len);
}
uint64_t lg_crp_get_clock(const lg_crp* g) { return g ? g->prng.clock : 0; }
int lg_crp_set_clock(lg_crp* g, uint64_t n) {
    LG_REQUIRE(g, "CRPGenerator.SetClock: null argument");
    return prng_set_clock(&g->prng, n);
}

// CRPGenerator.Clock, ring/prng.go:71-103, into host memory laid out [nl][N]
int lg_crp_clock_host(lg_crp* g, uint64_t* host) {
    LG_REQUIRE(g && host, "CRPGenerator.Clock: null argument");
    const u64 N = g->ring->N;
    const int nl = g->ring->nl;
    uint8_t bytes[64];
    size_t pos = 0;
    g->prng.step(bytes);  // :76 "starts with random bytes from the prng"
    for (u64 i = 0; i < N; ++i) {
        for (int j = 0; j < nl; ++j) {
            const u64 qi = g->ring->q[j], mask = g->masks[j];
            u64 coeff;
            for (;;) {
                if (64 - pos < 8) {  // :84-86
                    g->prng.step(bytes);
                    pos = 0;
                }
                coeff = __builtin_bswap64(load_le64(bytes + pos)) & mask;  // binary.BigEndian.Uint64, :89
                pos += 8;
                if (coeff < qi) break;
            }
            host[(size_t)j * N + i] = coeff;
        }
    }
    return LG_OK;
}

// the same into entry `batch_index` of a device polynomial (limbs 0..ring.nl-1); returns when the copy is done
int lg_crp_clock(lg_crp* g, lg_poly* out, int batch_index, lg_stream_t stream) {
    LG_REQUIRE(g && out, "CRPGenerator.Clock: null argument");
    LG_REQUIRE(out->N == g->ring->N && out->nlimbs >= g->ring->nl, "CRPGenerator.Clock: polynomial does not fit the context");
    LG_REQUIRE(batch_index >= 0 && batch_index < out->batch, "CRPGenerator.Clock: batch index out of range");
    LG_ON_DEVICE(out->device);
    const size_t words = (size_t)g->ring->nl * g->ring->N;
    if (!g->stage) LG_CUDA_CHECK(cudaMallocHost((void**)&g->stage, words * sizeof(u64)));
    LG_TRY(lg_crp_clock_host(g, g->stage));
    cudaStream_t st = (cudaStream_t)stream;
    LG_CUDA_CHECK(cudaMemcpyAsync(out->d + (size_t)batch_index * out->bstride, g->stage, words * sizeof(u64),
                                  cudaMemcpyHostToDevice, st));
    LG_CUDA_CHECK(cudaStreamSynchronize(st));
    return LG_OK;
}

}  // extern "C"
