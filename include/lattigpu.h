/*
 * lattigpu.h -- C ABI of the B200-native RNS polynomial-ring engine.
 *
 * This is the drop-in boundary for the ring hot path of Lattigo v1.3.1
 * (Eleven-Z/lattigo-FHE-by-go).  The reference has no FFI layer: ring.Context
 * is a Go struct whose methods the ckks/bfv/dckks/dbfv packages call directly.
 * Each entry point below replaces one of those methods (cited as
 * reference-file:line); the Go side keeps its types and calls these through
 * cgo.  Plain pointers and sizes only.
 *
 * Conventions
 *  - every function returns 0 (LG_OK) or a negative status; lg_last_error()
 *    gives the message for the calling thread.  The Go shim turns a non-zero
 *    status into panic(), matching the reference's misuse behaviour
 *    (ring/ring_context.go:72,136).
 *  - `nl` is the number of ACTIVE limbs an op touches (= level+1 of the
 *    reference's *Lvl variants; pass the ring's limb count for the plain ones).
 *  - a polynomial handle owns (or wraps) a device buffer laid out
 *    [batch][nlimbs][N] of uint64; every op is applied to all `batch` entries
 *    (batch = 1 reproduces the reference's one-poly call).
 *  - `stream` is a cudaStream_t (NULL = the default stream).  Ops are
 *    asynchronous; lg_stream_sync / lg_poly_download order them.
 *  - like the reference, ops do not range-check limb counts beyond what is
 *    needed for memory safety; out-of-range arguments return LG_ERR_ARG.
 */
#ifndef LATTIGPU_H
#define LATTIGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LG_OK 0
#define LG_ERR_ARG (-1)      /* invalid argument (the reference panics or indexes out of range) */
#define LG_ERR_NTT (-2)      /* moduli do not allow the NTT (ring_context.go:142-145) */
#define LG_ERR_CUDA (-3)     /* CUDA runtime error */
#define LG_ERR_NOMEM (-4)
#define LG_ERR_NODEVICE (-5) /* no CUDA device: there is no CPU fallback */

typedef struct lg_ring lg_ring;             /* ring.Context            ring/ring_context.go:18-51 */
typedef struct lg_poly lg_poly;             /* ring.Poly (device)      ring/ring_object.go:11-13 */
typedef struct lg_extender lg_extender;     /* ring.FastBasisExtender  ring/ring_basis_extension.go:9-18 */
typedef struct lg_decomposer lg_decomposer; /* ring.Decomposer         ring/ring_basis_extension.go:398-407 */
typedef struct lg_galois lg_galois;         /* []uint64 index of PermuteNTTIndex, ring/ring_galois.go:29-50 */
typedef struct lg_ckks_eval lg_ckks_eval;   /* hot ops of ckks.evaluator, ckks/evaluator.go:64-76 */
typedef struct lg_bfv_eval lg_bfv_eval;     /* hot ops of bfv.evaluator,  bfv/evaluator.go:41-60 */
typedef struct lg_swk lg_swk;               /* ckks/bfv SwitchingKey.evakey [beta][2] QP polys, ckks/keygen.go:282-340 */
typedef struct lg_hoisted lg_hoisted;       /* c2QiQDecomp/c2QiPDecomp of RotateHoisted, ckks/evaluator.go:1261-1273 */
typedef struct lg_comm lg_comm;             /* ranks of one node (NCCL over NVLink); no counterpart in the reference */
typedef void* lg_stream_t;                  /* cudaStream_t */

/* ---- library / device ---------------------------------------------------- */
const char* lg_last_error(void);
const char* lg_version(void);
int lg_device_count(int* count);
/* Sets the calling thread's current device (objects are created on it).  Every handle remembers the device it was
 * created on (wrapped pointers: the device that owns the pointer), and every entry point that touches the GPU switches
 * to its handles' device for the duration of the call and restores the caller's -- a Go host whose goroutines migrate
 * between OS threads, or one process driving several GPUs, needs no further care (SURVEY.md 8(b) Threading).
 * Operands that live on different devices return LG_ERR_ARG. */
int lg_set_device(int device);
int lg_stream_create(lg_stream_t* stream);
int lg_stream_destroy(lg_stream_t stream);
int lg_stream_sync(lg_stream_t stream);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t lg_launch_count(void);
/* Diagnostic A/B switches.  They select between kernel variants that return IDENTICAL words (every variant is
 * parity-tested); they exist for cross-checks and profiling, not for tuning by users.  Each is initialised ONCE per
 * process from the environment variable LATTIGPU_<NAME IN UPPER CASE> (no getenv on any launch path) and can then only
 * be changed through this call.  Names (value 0/1 unless noted): "literal_ntt" (literal Butterfly/InvButterfly of
 * ring/ntt.go:32-50 in every transform), "no_d64_ntt" (integer instead of FP64-only butterflies for moduli below
 * 3*2^44), "ks_acc64" (64-bit instead of 96-bit key-switch accumulators), "no_fp_mac" (integer instead of FP64 key-switch
 * accumulators),
 * "no_fp_modup" / "no_lazy_modup" / "no_wide_modup" / "modup_cpt2" (basis-extension kernel choice), "no_tail_canon" /
 * "no_fused_tail" (ModDown / rescale tail placement), "ks_scratch_words" (value = digit scratch budget of the key
 * switch in 64-bit words, 0 restores the 6 GiB default), "ntt_l2_bytes" (value = bytes of first-phase output a group of
 * batch entries of a two-phase NTT may hold, 0 = default: no grouping), "reverse_walk" (second NTT phases walk their
 * grid backwards; default off), "ntt_l2_streams" (1 / 2: the groups of "ntt_l2_bytes" alternate between two auxiliary
 * streams, by batch entries / by limbs), "tile_fastest" (default 1: strided NTT phases walk the tiles of a limb fastest),
 * "no_strided_tma" (forward strided NTT phase with per-thread loads instead of the TMA ring), "no_ks_tma" (every limb of
 * the fused digit loop on the register-key kernel instead of the TMA kernel), "ks_key_pf" (on the TMA digit loop: bit 0 =
 * suspend-time hint on the key wait, bit 1 = one key hand-off per CTA instead of one per warp pair; on the register-key
 * kernel: 1 / 2 = L1 prefetch of the key lines before the second register block / at the top of the iteration), "tail_pf"
 * (1 / 2 = L1 / L2 prefetch of the ModDown-tail operands), "no_aux_streams" (independent launches of one call stay on the
 * caller's stream instead of auxiliary streams forked from and joined to it).  Unknown names return LG_ERR_ARG. */
int lg_debug_set_switch(const char* name, uint64_t value);

/* ---- ring.Context -------------------------------------------------------- */
/* NewContextWithParams = SetParameters + GenNTTParams, ring_context.go:60-209:
 * tables are generated natively (same primitive-root search as ring/utils.go:182-288). */
int lg_ring_create(uint64_t N, int nlimbs, const uint64_t* moduli, lg_ring** out);
/* Same context from tables computed by the Go side (so that psi is literally the
 * reference's): bred = nlimbs x {hi,lo} (bredParams), mred = mredParams, psi /
 * psi_inv = nlimbs x N (nttPsi / nttPsiInv), ninv = nttNInv, rescale = the
 * triangular rescaleParams[j-1][i] flattened j-major (j=1..nlimbs-1, i<j); may be NULL. */
int lg_ring_create_from_tables(uint64_t N, int nlimbs, const uint64_t* moduli, const uint64_t* bred,
                               const uint64_t* mred, const uint64_t* psi, const uint64_t* psi_inv,
                               const uint64_t* ninv, const uint64_t* rescale, lg_ring** out);
int lg_ring_destroy(lg_ring* ring);
uint64_t lg_ring_n(const lg_ring* ring);
int lg_ring_nlimbs(const lg_ring* ring);
/* host copies of the tables (any pointer may be NULL); sizes as in create_from_tables */
int lg_ring_get_tables(const lg_ring* ring, uint64_t* moduli, uint64_t* bred, uint64_t* mred, uint64_t* psi,
                       uint64_t* psi_inv, uint64_t* ninv, uint64_t* rescale);

/* host-only number theory of the ring package (no device work) */
int lg_is_prime(uint64_t num);                                                    /* IsPrime, ring/utils.go:75-128 */
int lg_generate_ntt_primes(uint64_t logQ, uint64_t logN, uint64_t levels, uint64_t* primes); /* GenerateNTTPrimes :133-175 */
uint64_t lg_primitive_root(uint64_t q);                                           /* primitiveRoot :182-205 */

/* ---- ring.Poly ------------------------------------------------------------ */
int lg_poly_create(uint64_t N, int nlimbs, int batch, lg_poly** out);            /* ring.NewPoly, ring_object.go:16-23 */
int lg_poly_wrap(void* device_ptr, uint64_t N, int nlimbs, int batch, lg_poly** out); /* non-owning */
/* non-owning view of limbs [limb0, limb0+nlimbs) of every batch entry (Go: p.Coeffs[a:b]) */
int lg_poly_view(const lg_poly* parent, int limb0, int nlimbs, lg_poly** out);
int lg_poly_destroy(lg_poly* p);
uint64_t lg_poly_n(const lg_poly* p);
int lg_poly_nlimbs(const lg_poly* p);
int lg_poly_batch(const lg_poly* p);
void* lg_poly_device_ptr(const lg_poly* p);
size_t lg_poly_batch_stride(const lg_poly* p); /* in words */
/* host <-> device; host layout [nbatch][nl][N]; synchronous with respect to the host buffer */
int lg_poly_upload(lg_poly* p, int batch0, int nbatch, int limb0, int nl, const uint64_t* host, lg_stream_t stream);
int lg_poly_download(const lg_poly* p, int batch0, int nbatch, int limb0, int nl, uint64_t* host, lg_stream_t stream);
/* stream-ordered variants for C-allocated pinned staging buffers: the host buffer must stay valid (and
 * for uploads unchanged) until the stream reaches the copy -- not for Go-managed memory */
int lg_poly_upload_async(lg_poly* p, int batch0, int nbatch, int limb0, int nl, const uint64_t* host, lg_stream_t stream);
/* Wire format, ring/ring_object.go:146-289: WriteTo / WriteCoeffs (:161-184), GetDataLen (:186-192), DecodePolyNew /
 * UnmarshalBinary / DecodeCoeffs (:194-289).  data[0] = log2(N), data[1] = #moduli, then big-endian uint64 limb-major;
 * with_metadata = 0 is the header-less WriteCoeffs / DecodeCoeffs form.  Entry `batch_index` of the handle is written /
 * filled; the byte swap runs on the device, the call returns when `data` is complete. */
uint64_t lg_poly_get_data_len(const lg_poly* p, int nl, int with_metadata);
int lg_poly_write_to(const lg_poly* p, int batch_index, int nl, uint8_t* data, uint64_t len, int with_metadata, lg_stream_t stream);
int lg_poly_decode(lg_poly* p, int batch_index, const uint8_t* data, uint64_t len, int with_metadata, int nl, lg_stream_t stream);
int lg_poly_download_async(const lg_poly* p, int batch0, int nbatch, int limb0, int nl, uint64_t* host, lg_stream_t stream);
int lg_poly_zero(lg_poly* p, lg_stream_t stream);                                /* Poly.Zero, ring_object.go:60-67 */
int lg_poly_copy(const lg_poly* src, int nl, lg_poly* dst, lg_stream_t stream);  /* Copy/CopyLvl, ring_object.go:85-121 */

/* ---- NTT, ring/ntt.go ------------------------------------------------------ */
int lg_ring_ntt(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s);    /* NTT/NTTLvl :4-15 */
int lg_ring_invntt(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s); /* InvNTT/InvNTTLvl :18-29 */
/* free functions ring.NTT / ring.InvNTT (:53,:89) on ONE limb with the tables of `table_limb` */
int lg_ring_ntt_limb(const lg_ring* r, int table_limb, const lg_poly* p1, int limb1, lg_poly* p2, int limb2,
                     lg_stream_t s);
int lg_ring_invntt_limb(const lg_ring* r, int table_limb, const lg_poly* p1, int limb1, lg_poly* p2, int limb2,
                        lg_stream_t s);

/* ---- coefficient-wise ops, ring/ring.go ------------------------------------ */
int lg_ring_add(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);        /* :10-29 */
int lg_ring_add_nomod(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);  /* :32-51 */
int lg_ring_sub(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);        /* :54-73 */
int lg_ring_sub_nomod(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);  /* :76-97 */
int lg_ring_neg(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s);                          /* :100-119 */
int lg_ring_reduce(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s);                       /* :122-143 */
int lg_ring_mod(const lg_ring* r, int nl, const lg_poly* p1, uint64_t m, lg_poly* p2, lg_stream_t s);              /* :146-154 */
int lg_ring_and(const lg_ring* r, int nl, const lg_poly* p1, uint64_t m, lg_poly* p2, lg_stream_t s);              /* :157-164 */
int lg_ring_or(const lg_ring* r, int nl, const lg_poly* p1, uint64_t m, lg_poly* p2, lg_stream_t s);               /* :167-174 */
int lg_ring_xor(const lg_ring* r, int nl, const lg_poly* p1, uint64_t m, lg_poly* p2, lg_stream_t s);              /* :177-184 */
int lg_ring_mul_coeffs(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);               /* :187-195 */
int lg_ring_mul_coeffs_and_add(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);       /* :198-206 */
int lg_ring_mul_coeffs_and_add_nomod(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s); /* :209-217 */
int lg_ring_mul_coeffs_constant(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);      /* :335-343 */
int lg_ring_mul_coeffs_montgomery(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);               /* :221-243 */
int lg_ring_mul_coeffs_montgomery_and_add(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);       /* :247-269 */
int lg_ring_mul_coeffs_montgomery_and_add_nomod(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s); /* :273-295 */
int lg_ring_mul_coeffs_montgomery_constant_and_add_nomod(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s); /* :298-308 */
int lg_ring_mul_coeffs_montgomery_and_sub(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);       /* :311-319 */
int lg_ring_mul_coeffs_montgomery_and_sub_nomod(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s); /* :323-331 */
int lg_ring_mul_coeffs_montgomery_constant(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s);      /* :346-355 */
int lg_ring_mform(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s);    /* MForm/MFormLvl :583-607 */
int lg_ring_invmform(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s); /* InvMForm :610-619 */
/* scalar ops: `scalar` has one word per active limb -- the same word repeated for
 * AddScalar/SubScalar/MulScalar (:467,:490,:513), scalar mod q_i for the *Bigint
 * variants (:477,:500,:539,:556; the big.Int reduction stays on the Go side).
 * As in the reference (:482,:505) Add/SubScalar write into p1 itself. */
int lg_ring_add_scalar(const lg_ring* r, int nl, lg_poly* p1, const uint64_t* scalar, lg_stream_t s);
int lg_ring_sub_scalar(const lg_ring* r, int nl, lg_poly* p1, const uint64_t* scalar, lg_stream_t s);
int lg_ring_mul_scalar(const lg_ring* r, int nl, const lg_poly* p1, const uint64_t* scalar, lg_poly* p2, lg_stream_t s);
/* Inner loops of the CKKS constant ops, ckks/evaluator.go:373-833 (AddConst :433-444, MultByConstAndAdd :590-609,
 * MultByConst :700-727, MultByi :762-783, DivByi :811-832): per limb one scalar for coefficients [0, N/2) (lo[i]) and one
 * for [N/2, N) (hi[i]); the scalars (scaleUpExact, MRed by psi^2, MForm) are computed by the host as in the reference.
 *   add:      p2 = CRed(p1 + s)          mul: p2 = MRed(p1, s)          mul_and_add: p2 = CRed(p2 + MRed(p1, s)) */
int lg_ring_add_scalar_halves(const lg_ring* r, int nl, const lg_poly* p1, const uint64_t* lo, const uint64_t* hi, lg_poly* p2, lg_stream_t s);
int lg_ring_mul_scalar_montgomery_halves(const lg_ring* r, int nl, const lg_poly* p1, const uint64_t* lo, const uint64_t* hi, lg_poly* p2, lg_stream_t s);
int lg_ring_mul_scalar_montgomery_halves_and_add(const lg_ring* r, int nl, const lg_poly* p1, const uint64_t* lo, const uint64_t* hi, lg_poly* p2, lg_stream_t s);
int lg_ring_mul_by_pow2(const lg_ring* r, int nl, const lg_poly* p1, uint64_t pow2, lg_poly* p2, lg_stream_t s);          /* :629-653 */
int lg_ring_mult_by_monomial(const lg_ring* r, int nl, const lg_poly* p1, uint64_t deg, lg_poly* p2, lg_stream_t s);      /* :663-723 */
/* vector = device poly with one limb of N words */
int lg_ring_mul_by_vector_montgomery(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* vec, lg_poly* p2, lg_stream_t s);               /* :726-734 */
int lg_ring_mul_by_vector_montgomery_and_add_nomod(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* vec, lg_poly* p2, lg_stream_t s); /* :737-745 */
int lg_ring_bitreverse(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s);                          /* :749-772 (p1 != p2) */
/* The rest of ring.Context's method set (not on the evaluator path; csrc/ringext.cu).  All of them act on every limb of the
 * context, as the reference's do.  MulPoly :358-367 / MulPolyMontgomery :371-380 (montgomery != 0): p3 = InvNTT(MulCoeffs(NTT(p1),
 * NTT(p2))).  MulPolyNaive :383-410 / MulPolyNaiveMontgomery :413-437 (montgomery != 0): the N^2 negacyclic convolution.
 * Exp :441-464 keeps the reference's ending (p1 is left in the NTT domain, p2 = InvNTT(p1)).  Shift :575-580: p2[k] =
 * p1[(k + n) mod N], LG_ERR_ARG where Go's slice expression panics (n > N).  Rotate :775-800 multiplies coefficient j >= 1 of
 * p1 by root^j IN p1 (the reference never writes p2).  Equal :424-446 / EqualLvl :449-467 (nl = level + 1) reduce both
 * operands in place and synchronise the stream: *equal = 1 when every word agrees. */
int lg_ring_mul_poly(const lg_ring* r, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, int montgomery, lg_stream_t s);
int lg_ring_mul_poly_naive(const lg_ring* r, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, int montgomery, lg_stream_t s);
int lg_ring_exp(const lg_ring* r, lg_poly* p1, uint64_t e, lg_poly* p2, lg_stream_t s);
int lg_ring_shift(const lg_ring* r, const lg_poly* p1, uint64_t n, lg_poly* p2, lg_stream_t s);
int lg_ring_rotate(const lg_ring* r, lg_poly* p1, uint64_t n, lg_stream_t s);
int lg_ring_equal(const lg_ring* r, int nl, lg_poly* p1, lg_poly* p2, int* equal, lg_stream_t s);

/* ---- Galois automorphisms, ring/ring_galois.go ----------------------------- */
int lg_galois_create(uint64_t gen, uint64_t power, uint64_t N, lg_galois** out);           /* PermuteNTTIndex :29-50 */
int lg_galois_create_from_index(const uint64_t* index, uint64_t N, lg_galois** out);      /* index computed by Go */
int lg_galois_get_index(const lg_galois* g, uint64_t* index);
int lg_galois_destroy(lg_galois* g);
int lg_ring_permute_ntt_with_index(int nl, const lg_poly* in, const lg_galois* g, lg_poly* out, lg_stream_t s); /* :89-101 */
int lg_ring_permute_ntt(int nl, const lg_poly* in, uint64_t gen, lg_poly* out, lg_stream_t s);                  /* :55-84 */
int lg_ring_permute(const lg_ring* r, int nl, const lg_poly* in, uint64_t gen, lg_poly* out, lg_stream_t s);    /* Context.Permute :106-127 */

/* ---- RNS rescaling, ring/ring_scaling.go ----------------------------------- */
/* p0 has nl active limbs on entry; the result is in its first nl-1 (or nl-nb)
 * limbs -- the Go side re-slices Coeffs exactly as ring_scaling.go:33,53,113,147. */
int lg_ring_div_floor_by_last_modulus_ntt(const lg_ring* r, int nl, lg_poly* p0, lg_stream_t s);  /* :9-34 */
int lg_ring_div_floor_by_last_modulus(const lg_ring* r, int nl, lg_poly* p0, lg_stream_t s);      /* :37-54 */
int lg_ring_div_floor_by_last_modulus_many_ntt(const lg_ring* r, int nl, lg_poly* p0, int nb, lg_stream_t s); /* :57-61 */
int lg_ring_div_floor_by_last_modulus_many(const lg_ring* r, int nl, lg_poly* p0, int nb, lg_stream_t s);     /* :64-69 */
int lg_ring_div_round_by_last_modulus_ntt(const lg_ring* r, int nl, lg_poly* p0, lg_stream_t s);  /* :72-114 */
int lg_ring_div_round_by_last_modulus(const lg_ring* r, int nl, lg_poly* p0, lg_stream_t s);      /* :117-148 */
int lg_ring_div_round_by_last_modulus_many_ntt(const lg_ring* r, int nl, lg_poly* p0, int nb, lg_stream_t s); /* :152-156 */
int lg_ring_div_round_by_last_modulus_many(const lg_ring* r, int nl, lg_poly* p0, int nb, lg_stream_t s);     /* :159-164 */

/* ---- FastBasisExtender, ring/ring_basis_extension.go ------------------------ */
int lg_extender_create(const lg_ring* ringQ, const lg_ring* ringP, lg_extender** out);  /* NewFastBasisExtender :55-74 */
int lg_extender_destroy(lg_extender* e);
int lg_extender_modup_split_qp(const lg_extender* e, int level, const lg_poly* p1, lg_poly* p2, lg_stream_t s);  /* :147-149 */
int lg_extender_modup_split_pq(const lg_extender* e, int level, const lg_poly* p1, lg_poly* p2, lg_stream_t s);  /* :154-156 */
/* p1 holds all Q limbs then all P limbs; its P part is destroyed (as :172) */
int lg_extender_moddown_ntt_pq(const lg_extender* e, int level, lg_poly* p1, lg_poly* p2, lg_stream_t s);        /* :163-200 */
/* p1P is destroyed (as :215) */
int lg_extender_moddown_splited_ntt_pq(const lg_extender* e, int level, const lg_poly* p1Q, lg_poly* p1P, lg_poly* p2, lg_stream_t s); /* :207-242 */
int lg_extender_moddown_pq(const lg_extender* e, int level, const lg_poly* p1, lg_poly* p2, lg_stream_t s);      /* :248-275 */
int lg_extender_moddown_splited_pq(const lg_extender* e, int level, const lg_poly* p1Q, const lg_poly* p1P, lg_poly* p2, lg_stream_t s); /* :281-308 */
int lg_extender_moddown_splited_qp(const lg_extender* e, int levelQ, int levelP, const lg_poly* p1Q, const lg_poly* p1P, lg_poly* p2, lg_stream_t s); /* :314-350 */

/* ---- Decomposer, ring/ring_basis_extension.go:398-713 ----------------------- */
int lg_decomposer_create(uint64_t N, const uint64_t* Q, int nQ, const uint64_t* P, int nP, lg_decomposer** out); /* NewDecomposer :413-472 */
int lg_decomposer_destroy(lg_decomposer* d);
int lg_decomposer_beta(const lg_decomposer* d);
int lg_decomposer_xalpha(const lg_decomposer* d, int i);                               /* Xalpha :408-410 */
int lg_decomposer_decompose(const lg_decomposer* d, int level, int crt, const lg_poly* p0, lg_poly* p1, lg_stream_t s); /* :476-597 */
int lg_decomposer_decompose_and_split(const lg_decomposer* d, int level, int crt, const lg_poly* p0, lg_poly* p1Q, lg_poly* p1P, lg_stream_t s); /* :601-713 */

/* ---- evaluator key-switch path, ckks/evaluator.go ---------------------------- */
int lg_ckks_eval_create(const lg_ring* ringQ, const lg_ring* ringP, lg_ckks_eval** out); /* NewEvaluator :81-112 (ring part) */
int lg_ckks_eval_destroy(lg_ckks_eval* e);
/* host layout [beta][2][nQ+nP][N], NTT + Montgomery form (ckks/keygen.go:282-340) */
int lg_swk_create(uint64_t N, int beta, int nQP, const uint64_t* host, lg_swk** out);
int lg_swk_wrap(void* device_ptr, uint64_t N, int beta, int nQP, lg_swk** out);
/* marshaling support (ckks/marshaler.go:193-283, bfv/marshaler.go:202-287): an empty key of a given shape, and
 * evakey[digit][half] as a non-owning ring.Poly handle over QP for lg_poly_write_to / lg_poly_decode */
int lg_swk_alloc(uint64_t N, int beta, int nQP, lg_swk** out);
int lg_swk_poly(const lg_swk* k, int digit, int half, lg_poly** out);
/* A key is immutable once an evaluator has used it (derived device forms are built at first use).  Call this after
 * writing to the key's memory by other means than a view obtained afterwards (lg_swk_poly invalidates by itself). */
int lg_swk_invalidate(lg_swk* k);
int lg_swk_beta(const lg_swk* k);
int lg_swk_nlimbs(const lg_swk* k);
uint64_t lg_swk_n(const lg_swk* k);
int lg_swk_destroy(lg_swk* k);
/* switchKeysInPlace :1475-1558: p0,p1 receive the level+1 limb results */
int lg_ckks_switch_keys_in_place(lg_ckks_eval* e, int level, const lg_poly* cx, const lg_swk* evk, lg_poly* p0, lg_poly* p1, lg_stream_t s);
/* MulRelin :1016-1133, ciphertext x ciphertext with relinearisation key.  Passing
 * the same handles for (a0,a1) and (b0,b1) selects the squaring branch (:1080-1085). */
int lg_ckks_mul_relin(lg_ckks_eval* e, int level, const lg_poly* a0, const lg_poly* a1, const lg_poly* b0, const lg_poly* b1,
                      const lg_swk* rlk, lg_poly* out0, lg_poly* out1, lg_stream_t s);
/* Relinearize :1144-1162 */
int lg_ckks_relinearize(lg_ckks_eval* e, int level, const lg_poly* c0, const lg_poly* c1, const lg_poly* c2, const lg_swk* rlk,
                        lg_poly* out0, lg_poly* out1, lg_stream_t s);
/* Rescale loop body :955-960 applied nb times: DivRoundByLastModulusNTT on both polys */
int lg_ckks_rescale(lg_ckks_eval* e, int nl, lg_poly* c0, lg_poly* c1, int nb, lg_stream_t s);
/* SwitchKeys :1176-1189 */
int lg_ckks_switch_keys(lg_ckks_eval* e, int level, const lg_poly* c0, const lg_poly* c1, const lg_swk* k, lg_poly* out0, lg_poly* out1, lg_stream_t s);
/* permuteNTT :1452-1472 = RotateColumns with a direct key (:1220) / Conjugate (:1449) */
int lg_ckks_permute_ntt(lg_ckks_eval* e, int level, const lg_poly* c0, const lg_poly* c1, const lg_galois* g, const lg_swk* k,
                        lg_poly* out0, lg_poly* out1, lg_stream_t s);

/* RotateHoisted :1252-1289.  lg_ckks_hoist is the precomputation (:1258-1273): InvNTT of value[1] and
 * decomposeAndSplitNTT of every digit, kept on the device in the returned handle; lg_ckks_switch_key_hoisted
 * is switchKeyHoisted (:1291-1392, ct0 != ctOut branch) for one rotation: g = permuteNTTLeftIndex[k],
 * key = evakeyRotColLeft[k].  out0 must not alias c0 (PermuteNTTWithIndex is not in place).
 * The handle's device memory is allocated and released in stream order on the stream given to lg_ckks_hoist:
 * lg_hoisted_destroy may be called right after the last rotation was issued on that stream; rotations issued on
 * another stream must have completed (lg_stream_sync) before it. */
int lg_ckks_hoist(lg_ckks_eval* e, int level, const lg_poly* c1, lg_hoisted** out, lg_stream_t s);
int lg_ckks_switch_key_hoisted(lg_ckks_eval* e, const lg_hoisted* h, const lg_poly* c0, const lg_galois* g, const lg_swk* k,
                               lg_poly* out0, lg_poly* out1, lg_stream_t s);
int lg_hoisted_destroy(lg_hoisted* h);

/* ---- evaluator key-switch path, bfv/evaluator.go ----------------------------- */
/* NewEvaluator :62-104 (ring part): contexts Q, QMul, P; baseconverterQ1Q2, baseconverterQ1P, decomposer, pHalf */
int lg_bfv_eval_create(const lg_ring* ringQ, const lg_ring* ringQMul, const lg_ring* ringP, uint64_t t, lg_bfv_eval** out);
int lg_bfv_eval_destroy(lg_bfv_eval* e);
/* Mul = tensorAndRescale :278-464 for two degree-1 ciphertexts (coefficient domain) -> degree 2.
 * Passing the same handles for (a0,a1) and (b0,b1) selects the squaring branch (:334-349). */
int lg_bfv_mul(lg_bfv_eval* e, const lg_poly* a0, const lg_poly* a1, const lg_poly* b0, const lg_poly* b1,
               lg_poly* out0, lg_poly* out1, lg_poly* out2, lg_stream_t s);
/* switchKeys :736-813: p0, p1 receive the #Q-limb results */
int lg_bfv_switch_keys_core(lg_bfv_eval* e, const lg_poly* cx, const lg_swk* evk, lg_poly* p0, lg_poly* p1, lg_stream_t s);
/* relinearize :480-500 of a degree-2 ciphertext with evakey[0] */
int lg_bfv_relinearize(lg_bfv_eval* e, const lg_poly* c0, const lg_poly* c1, const lg_poly* c2, const lg_swk* rlk,
                       lg_poly* out0, lg_poly* out1, lg_stream_t s);
/* SwitchKeys :540-558 */
int lg_bfv_switch_keys(lg_bfv_eval* e, const lg_poly* c0, const lg_poly* c1, const lg_swk* k, lg_poly* out0, lg_poly* out1,
                       lg_stream_t s);
/* permute :711-733 = RotateColumns with a direct key (:595, gen = galElRotColLeft[k]) / RotateRows (:669, gen = 2N-1) */
int lg_bfv_permute(lg_bfv_eval* e, const lg_poly* c0, const lg_poly* c1, uint64_t gen, const lg_swk* k, lg_poly* out0,
                   lg_poly* out1, lg_stream_t s);

/* ---- t/Q scaling and the BFV plaintext lift: ring/ring_scaling.go:166-300, ring/float128.go, bfv/encoder.go ---- */
/* ring.SimpleScaler: NewSimpleScaler :188-262 (w_i, t_i from "Float128" double-double divisions, generated by the
 * same sequence of IEEE binary64 operations as float128.go), Scale :271-300: p2[j][x] = round(t/Q * p1[.][x]) mod t
 * for every limb j of p2; p2 may be p1.  The ring handle must outlive the scaler. */
typedef struct lg_scaler lg_scaler;
int lg_scaler_create(uint64_t t, const lg_ring* ring, lg_scaler** out);
/* the same parameters for a bare modulus list (host only, no device needed): wi[nl], ti[nl][2], BRed / MRed constants of t */
int lg_scaler_params_host(uint64_t t, const uint64_t* moduli, int nl, uint64_t* wi, double* ti, uint64_t* add_param,
                          uint64_t* mul_param);
int lg_scaler_destroy(lg_scaler* s);
int lg_scaler_get_params(const lg_scaler* s, uint64_t* wi, double* ti /* [nlimbs][2] */);
int lg_scaler_scale(const lg_scaler* s, const lg_poly* p1, lg_poly* p2, lg_stream_t stream);
/* GenLiftParams bfv/utils.go:9-23 + encodePlaintext bfv/encoder.go:121-136 (after the InvNTT over contextT):
 * pt[i][x] = MRed(m[0][x], deltaMont[i]); m may be a view of limb 0 of pt, as in the reference's plaintext. */
typedef struct lg_bfv_lift lg_bfv_lift;
int lg_bfv_lift_create(const lg_ring* ringQ, uint64_t t, lg_bfv_lift** out);
int lg_bfv_lift_params_host(const uint64_t* moduli, int nl, uint64_t t, uint64_t* delta_mont); /* host only */
int lg_bfv_lift_destroy(lg_bfv_lift* l);
int lg_bfv_lift_get_params(const lg_bfv_lift* l, uint64_t* delta_mont);
int lg_bfv_lift_apply(const lg_bfv_lift* l, const lg_poly* m, lg_poly* pt, lg_stream_t stream);

/* ---- common reference polynomials: utils/prng.go, ring/prng.go ------------------------------------- */
/* utils.PRNG (utils/prng.go:11-72): BLAKE2b-512 hash chain, optional key of at most 64 bytes (NewPRNG :22-28).
 * Seed :38-43 resets the state (keeping the key) and the clock; Clock :51-56 returns the 64-byte digest of the
 * current state and absorbs it; SetClock :61-72 clocks forward, LG_ERR_ARG when n is behind the clock.
 * Host code: the chain is sequential by construction (see csrc/crp.cu). */
typedef struct lg_prng lg_prng;
typedef struct lg_crp lg_crp;
int lg_prng_create(const uint8_t* key, size_t keylen, lg_prng** out);
int lg_prng_destroy(lg_prng* p);
int lg_prng_seed(lg_prng* p, const uint8_t* seed, size_t len);
uint64_t lg_prng_get_clock(const lg_prng* p);
int lg_prng_clock(lg_prng* p, uint8_t out[64]);
int lg_prng_set_clock(lg_prng* p, uint64_t n);
/* ring.CRPGenerator (ring/prng.go:11-69): NewCRPGenerator :21-37, Seed :45-47, GetClock :40-42, SetClock :57-61.
 * The ring handle must outlive the generator. */
int lg_crp_create(const uint8_t* key, size_t keylen, const lg_ring* ring, lg_crp** out);
int lg_crp_destroy(lg_crp* g);
int lg_crp_seed(lg_crp* g, const uint8_t* seed, size_t len);
uint64_t lg_crp_get_clock(const lg_crp* g);
int lg_crp_set_clock(lg_crp* g, uint64_t n);
/* Clock :71-103: the next uniform polynomial of the ring (masked big-endian words, rejection sampling, coefficient-major
 * consumption of the stream) into entry `batch_index` of a device handle / into host memory [nlimbs][N] */
int lg_crp_clock(lg_crp* g, lg_poly* out, int batch_index, lg_stream_t stream);
int lg_crp_clock_host(lg_crp* g, uint64_t* host);

/* ---- multi-GPU (one process per GPU) ------------------------------ */
/* The reference is single-process; these entry points add the exchange steps the path has when it is spread over the
 * GPUs of a node.  NCCL (resolved at run time, dlopen "libnccl.so.2") serves the one reduction, AggregateShares; the
 * limb axis moves its limbs through peer memory over NVLink with no library collective on the data path. */
int lg_comm_get_unique_id(uint8_t* id128);            /* rank 0; ship the 128 bytes to the other ranks */
/* after lg_set_device; id128 = NULL creates a handle for the peer-memory (limb-axis) paths only */
int lg_comm_create(int world, int rank, const uint8_t* id128, lg_comm** out);
int lg_comm_destroy(lg_comm* c);
int lg_comm_world(const lg_comm* c);
int lg_comm_rank(const lg_comm* c);
/* party axis: AggregateShares of dckks/dbfv (e.g. dckks/publickey_gen.go:45-47) over ranks =
 * all-reduce(sum,u64) + Reduce; p holds this rank's share on entry and the aggregate on return */
int lg_comm_aggregate_shares(const lg_comm* c, const lg_ring* r, int nl, lg_poly* p, lg_stream_t s);

/* Limb axis (BASELINE config 4): the RNS limbs of a ciphertext are spread cyclically over the ranks -- limb t of
 * Q || P belongs to rank t mod world (lg_comm_limb_owner), balanced at every level and stable when a limb is dropped.
 * A limb-resident polynomial is a full-size handle of which only the rank's own limbs are meaningful.
 * Exchange: every rank owns one exchange buffer that its peers map; limbs cross NVLink exactly where a basis
 * extension needs every source limb (c2 before DecomposeAndSplit ckks/evaluator.go:1503-1513, the special-prime
 * accumulators before ModDown ring_basis_extension.go:219-226, the last limb of a rescale ring_scaling.go:80-103):
 * the consuming kernels load them from the owner's buffer, ordered by a one-CTA barrier kernel in stream order.
 *   setup (once):  lg_comm_xbuf_alloc on every rank, ship handle128 to the peers, lg_comm_xbuf_open for each of them
 *                  (ranks living in ONE process attach directly: lg_comm_xbuf_attach).
 * Every rank must issue the same sequence of limb-axis calls.  A barrier that waits more than 5 s for a peer sets an
 * error that lg_comm_check reports (after synchronising the stream). */
int lg_comm_limb_owner(int limb, int world);                                    /* host only */
size_t lg_comm_xbuf_words_needed(uint64_t N, int nQ, int nP, int batch);        /* host only: MulRelin+Rescale+gather */
int lg_comm_xbuf_alloc(lg_comm* c, size_t words, uint8_t* handle128);           /* handle128 may be NULL (one process) */
int lg_comm_xbuf_open(lg_comm* c, int peer, const uint8_t* handle128);          /* CUDA IPC mapping of a peer's buffer */
int lg_comm_xbuf_attach(lg_comm* c, int peer, const lg_comm* peer_comm);        /* same process: direct pointers */
size_t lg_comm_xbuf_words(const lg_comm* c);
int lg_comm_check(lg_comm* c, lg_stream_t s);
/* replicate the first nl limbs of a limb-resident polynomial on every rank */
int lg_comm_gather_limbs(lg_comm* c, const lg_ring* r, int nl, lg_poly* p, lg_stream_t s);
/* limb-resident forms of switchKeysInPlace (ckks/evaluator.go:1475-1558), MulRelin (:1016-1133) followed by
 * nrescale Rescale steps (:933-968), and Rescale: inputs and outputs hold the rank's own limbs */
int lg_ckks_switch_keys_in_place_resident(lg_ckks_eval* e, lg_comm* c, int level, const lg_poly* cx, const lg_swk* evk,
                                          lg_poly* p0, lg_poly* p1, lg_stream_t s);
int lg_ckks_mul_relin_rescale_resident(lg_ckks_eval* e, lg_comm* c, int level, const lg_poly* a0, const lg_poly* a1,
                                       const lg_poly* b0, const lg_poly* b1, const lg_swk* rlk, lg_poly* out0, lg_poly* out1,
                                       int nrescale, lg_stream_t s);
int lg_ckks_rescale_resident(lg_ckks_eval* e, lg_comm* c, int nl, lg_poly* c0, lg_poly* c1, lg_stream_t s);
/* replicated forms: inputs are read on the rank's own limbs only, outputs are gathered onto every rank */
int lg_ckks_switch_keys_in_place_sharded(lg_ckks_eval* e, lg_comm* c, int level, const lg_poly* cx, const lg_swk* evk,
                                         lg_poly* p0, lg_poly* p1, lg_stream_t s);
int lg_ckks_mul_relin_sharded(lg_ckks_eval* e, lg_comm* c, int level, const lg_poly* a0, const lg_poly* a1,
                              const lg_poly* b0, const lg_poly* b1, const lg_swk* rlk, lg_poly* out0, lg_poly* out1,
                              lg_stream_t s);
int lg_ckks_rescale_sharded(lg_ckks_eval* e, lg_comm* c, int nl, lg_poly* c0, lg_poly* c1, lg_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* LATTIGPU_H */
