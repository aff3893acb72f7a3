// kernels.h -- internal launcher interface between the C-ABI host layer
// (capi.cu) and the kernel families.  Not part of the public boundary.
#pragma once
#include <atomic>

#include "common.cuh"

extern std::atomic<uint64_t> lg_g_launches;  // kernels launched by this library

// Diagnostic A/B switches (lattigpu.h: lg_debug_set_switch).  Read from the LATTIGPU_* environment ONCE, at first use
// (std::call_once), then only changed through lg_debug_set_switch -- no getenv on any launch path.  All default to 0.
struct LgSwitches {
    std::atomic<int> literal_ntt{0};     // LATTIGPU_LITERAL_NTT: literal Butterfly/InvButterfly in every transform
    // LATTIGPU_REVERSE_WALK: second NTT phases and the fused digit loop walk their grids backwards, reading first what the
    // previous launch wrote last.  Measured (profiles/r02_reverse_walk_ab.jsonl, ABAB): no gain (467 -> 474..499 us per
    // 1088 limb-NTTs), so it is off.
    std::atomic<int> reverse_walk{0};
    std::atomic<int> no_d64_ntt{0};      // LATTIGPU_NO_D64_NTT: integer instead of FP64-only butterflies below 3*2^44
    std::atomic<int> ks_acc64{0};        // LATTIGPU_KS_ACC64: never take the 96-bit key-switch accumulators
    std::atomic<int> no_fp_mac{0};       // LATTIGPU_NO_FP_MAC: integer key-switch accumulators on the FP64-class limbs too
    std::atomic<int> no_fp_modup{0};     // LATTIGPU_NO_FP_MODUP: integer-only basis extension
    std::atomic<int> modup_cpt2{0};      // LATTIGPU_MODUP_CPT2: two instead of four coefficients per thread in modup_fp_kernel
    std::atomic<int> no_lazy_modup{0};   // LATTIGPU_NO_LAZY_MODUP: canonical key-switch digits
    std::atomic<int> no_wide_modup{0};   // LATTIGPU_NO_WIDE_MODUP: generic kernel for 5..16 sources
    std::atomic<int> no_tail_canon{0};   // LATTIGPU_NO_TAIL_CANON: reduce the transform before the ModDown tail
    std::atomic<int> no_fused_tail{0};   // LATTIGPU_NO_FUSED_TAIL: separate ModDown / rescale tail kernels
    std::atomic<uint64_t> ks_scratch_words{(uint64_t)6 << 27};  // LATTIGPU_KS_SCRATCH_WORDS: digit scratch budget (words)
    // LATTIGPU_NTT_L2_BYTES: a two-phase transform runs over groups of batch entries whose intermediate (the first phase's
    // output) would stay in L2 until the second phase reads it; 0 (default) = one launch pair over the whole batch.
    // Measured (profiles/r02_ntt_l2_sweep.jsonl): the small grids cost more than the saved HBM pass -- 456 us for
    // 1088 limb-NTTs unsplit against 608 us at 96 MiB and 1040 us at 12..32 MiB.
    std::atomic<uint64_t> ntt_l2_bytes{0};
    std::atomic<int> ntt_l2_streams{0};  // LATTIGPU_NTT_L2_STREAMS: the groups alternate between two auxiliary streams
    // LATTIGPU_KS_KEY_PF / LATTIGPU_TAIL_PF (A/B, 0 = off; lattigpu.h lists the values): software prefetch
    // (prefetch.global.L1, one instruction per 128-byte line) of the key lines of the current digit in ks_fused_kernel and of
    // the ModDown / rescale tail operands in the last NTT phase -- measured, no gain (profiles/r02_prefetch_ab.jsonl); on
    // ks_fused_tma_kernel "ks_key_pf" selects the hand-off variants instead (suspend-time hint, one hand-off per CTA).
    std::atomic<int> ks_key_pf{0};
    // LATTIGPU_NO_KS_TMA: every limb of the fused digit loop on ks_fused_kernel (keys through registers); default: the
    // FP64-class limbs on ks_fused_tma_kernel (two batch entries per CTA, key tiles through shared memory by TMA)
    std::atomic<int> no_ks_tma{0};
    // LATTIGPU_NO_AUX_STREAMS: independent launches of one call stay on the caller's stream (the per-digit basis extensions
    // of a single ciphertext, the integer launch of the fused digit loop beside the TMA one)
    std::atomic<int> no_aux_streams{0};
    // LATTIGPU_NO_STRIDED_TMA: the forward strided NTT phase with per-thread loads and stores (ntt_fwd_strided) instead of
    // the TMA-fed ring (ntt_fwd_strided_tma)
    std::atomic<int> no_strided_tma{0};
    // LATTIGPU_TILE_FASTEST=0: the strided NTT phases walk the batch entries fastest; default 1: the tiles of a limb fastest
    // (adjacent 128-byte columns in flight together).  Measured (profiles/r02_tile_fastest_ab.jsonl, 16 interleaved rounds):
    // forward / inverse limb-NTT -1.3 % / -1.1 %, step -0.5 %.
    std::atomic<int> tile_fastest{1};
    std::atomic<int> tail_pf{0};
};
LgSwitches& lg_switches();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel, device)
template <auto Kernel>
inline void lg_ensure_dyn_smem(size_t bytes) {
    static std::atomic<uint64_t> done{0};  // bit d = set on device d (devices >= 64: set every time)
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = dev < 64 ? (1ull << dev) : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return;
    cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (bit) done.fetch_or(bit, std::memory_order_release);
}

#define LG_MAX_LIMBS 64

// Auxiliary streams for independent launches inside one call (ntt.cu): forked from the caller's stream and joined back to it
// with events, so the call stays ordered on the caller's stream (and capturable).
#define LG_AUX_STREAMS 4
struct LgAux {
    cudaStream_t s[LG_AUX_STREAMS];
    cudaEvent_t fork, join[LG_AUX_STREAMS];
};
LgAux* lg_aux_streams();  // nullptr when they cannot be created
void lg_aux_fork(LgAux* a, cudaStream_t st, int n);
void lg_aux_join(LgAux* a, cudaStream_t st, int n);

// ---- K1: NTT ----------------------------------------------------------------
// Epilogue of the forward transform's last phase (logN >= 12): instead of storing NTT(x) the kernel stores
//   out = MRed(a + (q - NTT(x)), s_j)            (the (x - y) * P^-1 / q_last^-1 tail of ModDown and of the rescaling,
//   out = CRed(out + MRed(a + (q - NTT(x)), s_j)) ring_basis_extension.go:236-238, ring_scaling.go:30,:109; + the Add that follows)
// for data limb j of batch entry b.  Entries b >= split use the second (a, out, add) set with index b - split.
struct NttTail {
    int enabled;
    int split;
    int add[2];
    const u64* a[2];
    u64* out[2];
    size_t a_bs[2], out_bs[2];
    size_t a_ls, out_ls;  // words between consecutive data limbs of a[] / out[] (0 = N)
    const u64* s;  // device array, one scalar per TABLE limb (kept out of the kernel parameters: every NTT CTA loads those)
    int a_canon;   // the words of a[] are known to be below q (key-switch accumulators): the transform is subtracted unreduced
};
// Broadcast input of the forward transform's first phase (logN >= 12): every data limb j transforms the same source
// limb plus a per-limb constant, x_j = v + add[j] (unreduced) -- the rescaling's "last limb to every other limb"
// (ring_scaling.go:17-28, :80-103, add = pHalfNegQi or nothing) without materialising the copies.
struct NttBcast {
    int enabled;
    const u64* add;  // device array per TABLE limb, or nullptr
};
struct NttArgs {
    RingTables T;
    LimbMap map;
    const u64* in;
    u64* out;
    size_t in_bstride, out_bstride;  // words between consecutive batch entries
    size_t in_ls, out_ls;            // words between consecutive data limbs (0 = N); the second phase runs in place on out
    int skip0, skip1;                // data limbs in [skip0, skip1) are left untouched
    // forward, digit-batched launches: when skip_alpha > 0 the batch index is digit*skip_div + b and the skipped limbs
    // of that entry are the digit's own ones: data limbs whose TABLE limb tl lies in
    // [digit*skip_alpha, min((digit+1)*skip_alpha, skip_nl))
    int skip_alpha, skip_div, skip_nl;
    // inverse only: per data limb flag, non-zero = some input word of that limb (in any batch entry) is above 2q, use the
    // literal butterflies for that limb (written by lg_launch_range_flags); nullptr = inputs known to be in range
    const u32* flags;
    int no_d64;                      // set by the launchers from the "no_d64_ntt" switch
    int batch0;                      // index of the launch's first batch entry in the caller's batch (tail addressing)
    int rev;                         // walk the grid backwards (second phases: read first what the first phase wrote last)
    int tfast;                       // strided phases: tile index fastest (set by the launcher from the "tile_fastest" switch)
    int pf;                          // prefetch the tail operands into L1 (set by the launcher from the "tail_pf" switch)
    NttTail tail;                    // forward only
    NttBcast bcast;                  // forward only
};
int lg_launch_ntt(const NttArgs& args, int nlimbs, int batch, bool inverse, cudaStream_t st);
// forward transform, strided phase only, in place or out of place (logN >= 12); the contiguous phase is
// then run by lg_launch_ks_fused
int lg_launch_ntt_fwd_strided(const NttArgs& args, int nlimbs, int batch, cudaStream_t st);
// flags[j] = any word of data limb j (over the whole batch) is > 2q
int lg_launch_range_flags(const NttArgs& args, int nlimbs, int batch, u32* flags, cudaStream_t st);

// Key-switch digit loop fused with the contiguous NTT phase (ckks/evaluator.go:1511-1552,
// bfv/evaluator.go:760-806): for every digit i the CTA finishes the forward NTT of its tile of
// D[i] (or takes the digit's own limbs from the NTT-domain input cx), multiplies by evk[i][0/1] and
// accumulates in registers; one canonical store of acc0/acc1 at the end.
struct KsFusedArgs {
    RingTables T;       // QP tables
    LimbMap map;        // data limb -> table limb (also the evk limb)
    const u64* D;       // digits after the strided phase: limb j of batch b of digit i at D + i*d_ds + b*d_bs + j*N
    size_t d_ds, d_bs;
    const u64* cx;      // NTT-domain key-switch input: limb j of batch b at cx + b*cx_bs + j*cx_ls
    size_t cx_bs, cx_ls;  // cx_ls = 0: N
    const u64* evk;     // evk[i][h], table limb tl at evk + i*evk_ds + h*evk_hs + tl*N (shared by the batch)
    size_t evk_ds, evk_hs;
    const u64* evk_f;   // the same key out of Montgomery form as doubles (FP64-class limbs; lg_launch_swk_prepare), or nullptr
    const u32* key_bad; // per (digit, half, table limb): a word of that key limb is not canonical -> integer accumulators
    u64* acc0;          // outputs, limb j of batch b at acc + b*acc_bs + j*N, canonical
    u64* acc1;
    size_t acc_bs;
    int beta, alpha, nl;  // digit i owns the data limbs whose TABLE limb lies in [i*alpha, min((i+1)*alpha, nl))
    int acc64;            // 1 = never take the 96-bit accumulators (LATTIGPU_KS_ACC64=1: A/B and cross-check)
    int no_d64;           // set by the launcher from the "no_d64_ntt" switch
    int rev;              // walk the grid backwards (set by the launcher)
    int pf;               // the "ks_key_pf" switch (set by the launcher): key-line prefetch / TMA hand-off variants, see lattigpu.h
    // blockIdx.z -> data limb (set by the launcher when it splits the limbs between ks_fused_kernel and ks_fused_tma_kernel)
    int use_zl;
    unsigned char zl[LG_MAX_LIMBS];
    // host-side, read by the launcher only (both or neither): the TMA descriptor of evk_f (a CUtensorMap: rows of 16 words,
    // 128-byte swizzle) and, per table limb, whether ks_fused_tma_kernel may take it (FP64-class modulus, every key word
    // canonical); lgi_swk_prepare builds both
    const void* h_keymap;
    const unsigned char* h_fp_ok;
};
int lg_launch_ks_fused(const KsFusedArgs& a, int nlimbs, int batch, cudaStream_t st);
// CUtensorMap (128 bytes at `map`) of a key in FP64 form: `words` u64 at `keyf` seen as rows of 16 words, boxes of 32 rows
// (one 2048-word tile), 128-byte swizzle.  0 = ok (needs a driver with cuTensorMapEncodeTiled)
int lg_encode_key_tensor_map(void* map, const u64* keyf, size_t words);
// keyf = double(InvMForm(key)) for the limbs with q < 3*2^44, bad[(digit*2+half)*nQP + tl] |= 1 when a word >= q
int lg_launch_swk_prepare(const RingTables& T, const u64* key, u64* keyf, u32* bad, int beta, int nQP, cudaStream_t st);

// ---- K3a: coefficient-wise ops (ring/ring.go) --------------------------------
enum EwOp {
    EW_ADD = 0,
    EW_ADD_NOMOD,
    EW_SUB,
    EW_SUB_NOMOD,
    EW_NEG,
    EW_REDUCE,
    EW_MUL_BARRETT,
    EW_MUL_BARRETT_ADD,
    EW_MUL_BARRETT_ADD_NOMOD,
    EW_MUL_BARRETT_CONSTANT,
    EW_MULMONT,
    EW_MULMONT_ADD,
    EW_MULMONT_ADD_NOMOD,
    EW_MULMONT_CONSTANT_ADD_NOMOD,
    EW_MULMONT_SUB,
    EW_MULMONT_SUB_NOMOD,
    EW_MULMONT_CONSTANT,
    EW_MFORM,
    EW_INVMFORM,
    EW_ADD_SCALAR,       // c = CRed(a + s_j)
    EW_SUB_SCALAR,       // c = CRed(a + (q - s_j))
    EW_MUL_SCALAR,       // c = MRed(a, MForm(BRedAdd(s_j)))
    EW_MUL_SCALAR_MONT,  // c = MRed(a, s_j)             (s_j already Montgomery)
    EW_MUL_POW2,         // c = PowerOf2(a, s_0)
    EW_AND,
    EW_OR,
    EW_XOR,
    EW_MOD,              // c = BRedAdd(a, m) for a foreign modulus m: s = {m, u0}
    EW_MULVEC,           // c = MRed(a, vec)   (vec = one limb of N words, shared by all limbs)
    EW_MULVEC_ADD_NOMOD, // c += MRed(a, vec)
    EW_SUB_MULMONT_SCALAR,  // c = MRed(a + (q - b), s_j)   (ModDown / rescale tail)
    EW_SUB_MULMONT_SCALAR_ADD,  // c = CRed(c + MRed(a + (q - b), s_j))  (ModDown tail fused with the AddLvl that follows)
    EW_COPY,
    // CKKS constant ops (ckks/evaluator.go:373-833): one scalar for the first N/2 coefficients (s_j) and one
    // for the last N/2 (shi_j) -- a + b*psi^2 and a - b*psi^2 in the NTT domain
    EW_ADD_SCALAR2,           // c = CRed(a + s)
    EW_MUL_SCALAR_MONT2,      // c = MRed(a, s)
    EW_MUL_SCALAR_MONT2_ADD,  // c = CRed(c + MRed(a, s))
    EW_NUM_OPS
};

struct EwArgs {
    RingTables T;
    LimbMap map;
    const u64* a;
    const u64* b;
    u64* c;
    size_t a_bs, b_bs, c_bs;  // batch strides (words); 0 broadcasts one entry over the batch
    size_t a_ls, b_ls, c_ls;  // limb strides (words); normally N, 0 broadcasts one limb
    u64 s[LG_MAX_LIMBS];      // per-data-limb scalars
    u64 shi[LG_MAX_LIMBS];    // *_SCALAR2 ops: the scalars of the upper half of the coefficients
};
int lg_launch_ew(int op, const EwArgs& args, int nlimbs, int batch, cudaStream_t st);

// ---- K4: permutations (ring/ring_galois.go, ring/ring.go) -------------------
struct PermArgs {
    RingTables T;
    LimbMap map;
    const u64* in;
    u64* out;
    size_t in_bs, out_bs;
    const u32* index;  // gather table (PermuteNTTWithIndex)
    u64 gen;           // Galois element (Permute) / monomial degree (MultByMonomial)
};
int lg_launch_permute_ntt(const PermArgs& a, int nlimbs, int batch, cudaStream_t st);
int lg_launch_permute_coeff(const PermArgs& a, int nlimbs, int batch, cudaStream_t st);
int lg_launch_mult_by_monomial(const PermArgs& a, int nlimbs, int batch, cudaStream_t st);
int lg_launch_bitreverse(const PermArgs& a, int nlimbs, int batch, cudaStream_t st);
int lg_launch_bswap64(const u64* in, u64* out, size_t words, cudaStream_t st);  // byte-reversed words (wire format)

// ---- K3b: exact basis extension (ring/ring_basis_extension.go:352-393) ------
// Device-resident modupParams.  Source basis = nsrc primes, target basis = ndst
// primes; tables are laid out for the FULL basis the params were built for
// (src_total sources, dst_total targets) so that a prefix of the sources and an
// arbitrary list of targets can be used, as the reference does with slices.
struct ModUpTables {
    const u64* srcQ;     // [src_total]
    const u64* srcQinv;  // [src_total]
    const u64* qib;      // [src_total]              qibMont
    const u64* qispj;    // [src_total][dst_total]   qispjMont
    const u64* qpjinv;   // [dst_total][src_total+1] qpjInv
    const u64* dstQ;     // [dst_total]
    const u64* dstQinv;  // [dst_total]
    const u64* dstU0;    // [dst_total]  bredParams[0]
    int src_total, dst_total;
};
struct ModUpArgs {
    ModUpTables M;
    u32 N;
    int nsrc;            // active sources: tables rows 0..nsrc-1
    const u64* in;       // source limb i at in + b*in_bs + i*N ...
    const u64* src[4];   // ... unless src[0] != nullptr (nsrc <= 4): source limb i at src[i] + b*in_bs -- limbs that live in
                         // different buffers, e.g. on peer GPUs (multi-GPU limb axis)
    size_t in_bs;
    // targets are described by up to 3 runs: run k writes ndst_k limbs starting at
    // out_k (limb stride N) using table targets tgt0_k, tgt0_k+1, ...
    int nruns;
    u64* out[3];
    size_t out_bs[3];
    int ndst[3];
    int tgt0[3];
    int tstep;           // table targets of a run are tgt0, tgt0 + tstep, ... (0 = 1); the output limbs stay N words apart
    // optional pass-through: copy the nsrc source limbs to copy_out (Decompose*'s
    // "p1.Coeffs[i+p0idxst][x] = p0.Coeffs[i+p0idxst][x]")
    u64* copy_out;
    size_t copy_bs;
    int fast;            // 1: every modulus is below 2^61 (modup_fast_kernel); 2: and the sources sum below 2^48 (modup_fp_kernel);
                         // 3: two-step FP64 quotient at granularity 2^fp_shift (modup_fp2_kernel)
    int fp_shift;
    int lazy_out;        // FP64 kernels only: leave the results in [0, 2p) (key-switch digits, read by the forward NTT alone)
};
int lg_launch_modup(const ModUpArgs& a, int batch, cudaStream_t st);

// broadcast of one limb to many (trivial decomposition case) and the rescale
// "last limb + pHalf" fan-out
struct FanoutArgs {
    u32 N;
    const u64* in;  // one limb per batch entry
    size_t in_bs;
    int nruns;
    u64* out[2];
    size_t out_bs[2];
    int ndst[2];
    // mode 0: plain copy.  mode 1: out_i = CRed(in + phalf, plast) + add[i]
    int mode;
    u64 phalf, plast;
    u64 add[LG_MAX_LIMBS];
};
int lg_launch_fanout(const FanoutArgs& a, int batch, cudaStream_t st);

// ---- K3e: key-switch multiply-accumulate -------------------------------------
struct KsMacArgs {
    RingTables T;       // QP tables
    LimbMap map;        // data limb -> table limb (also the evk limb)
    const u64* d;       // decomposed digit, NTT domain: limb j at d + b*d_bs + j*N
    size_t d_bs;
    const u64* evk0;    // evk[i][0], limb tl at evk0 + tl*N   (shared by the batch)
    const u64* evk1;
    u64* acc0;          // accumulators, limb j at acc + b*acc_bs + j*N
    u64* acc1;
    size_t acc_bs;
    int first;          // 1: acc = MRed(..) (no read), 0: acc += MRed(..)
    int reduce;         // 1: acc = BRedAdd(acc + MRed(..))
};
int lg_launch_ks_mac(const KsMacArgs& a, int nlimbs, int batch, cudaStream_t st);

// Hoisted key-switch (ckks/evaluator.go:1336-1378): acc0/acc1 = sum_i MRed(evk[i][0/1], D[i][index[.]]),
// the NTT-domain digits D gathered through the Galois index table, accumulators in registers over the
// digits, one canonical store.
struct KsHoistArgs {
    RingTables T;       // QP tables
    LimbMap map;
    const u64* D;       // NTT-domain digits: limb j of batch b of digit i at D + i*d_ds + b*d_bs + j*N
    size_t d_ds, d_bs;
    const u32* index;   // permuteNTT index table (N entries)
    const u64* evk;
    size_t evk_ds, evk_hs;
    const u64* evk_f;   // FP64 form of the key (lg_launch_swk_prepare) and its per-(digit, half, limb) flags, or nullptr
    const u32* key_bad;
    int nqp;            // limbs per key polynomial (rows of key_bad)
    u64* acc0;
    u64* acc1;
    size_t acc_bs;
    int beta;
};
int lg_launch_ks_hoisted(const KsHoistArgs& a, int nlimbs, int batch, cudaStream_t st);

// fused tensor product of MulRelin (ckks/evaluator.go:1076-1095); limb j = table limb j
struct TensorArgs {
    RingTables T;
    const u64 *a0, *a1, *b0, *b1;
    u64 *c0, *c1, *c2;
    size_t a_bs[2], b_bs[2], c_bs[3];
    int square;
    int nomod;  // 1: BFV form, c1 is accumulated without reduction (bfv/evaluator.go:344,361)
    int limb0;  // first limb processed (limb-sharded launches); blockIdx.y counts from it ...
    int lstep = 0;  // ... in steps of lstep limbs (0 = 1)
    int c2_compact = 0;  // 1: c2 limb y of the launch goes to slot y of c2 (rank-private scratch) instead of its own limb index
};
int lg_launch_tensor(const TensorArgs& a, int nlimbs, int batch, cudaStream_t st);
