// elementwise.cu -- kernel family K3a: coefficient-wise ring ops.
//
// Replaces the per-limb loops of ring/ring.go (Add/Sub/Neg/Reduce/MulCoeffs*/
// MForm/InvMForm/MulScalar*/...).  Each op transcribes the reference's uint64
// formula (non-canonical results such as Neg(0)=q are preserved).  HBM-bound:
// 128-bit vector loads/stores, grid = (chunks, limbs, batch).
#include "common.cuh"
#include "kernels.h"

namespace {

template <int OP>
LG_DEV u64 ew_apply(u64 a, u64 b, u64 c, const LimbConst& k, u64 s, u64 s2) {
    const u64 q = k.q;
    switch (OP) {
        case EW_ADD: return cred(a + b, q);                                   // ring.go:10-29
        case EW_ADD_NOMOD: return a + b;                                      // :32-51
        case EW_SUB: return cred((a + q) - b, q);                             // :54-73
        case EW_SUB_NOMOD: return (a + q) - b;                                // :76-97
        case EW_NEG: return q - a;                                            // :100-119
        case EW_REDUCE: return bred_add(a, q, k.u0);                          // :122-143
        case EW_MUL_BARRETT: return bred(a, b, q, k.u0, k.u1);                // :187-195
        case EW_MUL_BARRETT_ADD: return cred(c + bred(a, b, q, k.u0, k.u1), q);  // :198-206
        case EW_MUL_BARRETT_ADD_NOMOD: return c + bred(a, b, q, k.u0, k.u1);  // :209-217
        case EW_MUL_BARRETT_CONSTANT: return bred_constant(a, b, q, k.u0, k.u1);  // :335-343
        case EW_MULMONT: return mred(a, b, q, k.qinv);                        // :221-243
        case EW_MULMONT_ADD: return cred(c + mred(a, b, q, k.qinv), q);       // :247-269
        case EW_MULMONT_ADD_NOMOD: return c + mred(a, b, q, k.qinv);          // :273-295
        case EW_MULMONT_CONSTANT_ADD_NOMOD: return c + mred_constant(a, b, q, k.qinv);  // :298-308
        case EW_MULMONT_SUB: return cred(c + (q - mred(a, b, q, k.qinv)), q);  // :311-319
        case EW_MULMONT_SUB_NOMOD: return c + (q - mred(a, b, q, k.qinv));    // :323-331
        case EW_MULMONT_CONSTANT: return mred_constant(a, b, q, k.qinv);      // :346-355
        case EW_MFORM: return mform(a, q, k.u0, k.u1);                        // :583-607
        case EW_INVMFORM: return invmform(a, q, k.qinv);                      // :610-619
        case EW_ADD_SCALAR: return cred(a + s, q);                            // :467-487
        case EW_SUB_SCALAR: return cred(a + (q - s), q);                      // :490-510
        case EW_MUL_SCALAR:                                                   // :513-572 (s pre-converted per CTA)
        case EW_MUL_SCALAR_MONT: return mred(a, s, q, k.qinv);
        case EW_MUL_POW2: return power_of_2(a, (u32)s, q, k.qinv);            // :629-653
        case EW_AND: return a & s;                                            // :157-164
        case EW_OR: return a | s;                                             // :167-174
        case EW_XOR: return a ^ s;                                            // :177-184
        case EW_MOD: return bred_add(a, s, s2);                               // :146-154
        case EW_MULVEC: return mred(a, b, q, k.qinv);                         // :726-734
        case EW_MULVEC_ADD_NOMOD: return c + mred(a, b, q, k.qinv);           // :737-745
        case EW_SUB_MULMONT_SCALAR: return mred(a + (q - b), s, q, k.qinv);   // ring_basis_extension.go:236-238
        case EW_SUB_MULMONT_SCALAR_ADD: return cred(c + mred(a + (q - b), s, q, k.qinv), q);  // + ckks/evaluator.go:1103
        case EW_COPY: return a;
        case EW_ADD_SCALAR2: return cred(a + s, q);                           // ckks/evaluator.go:433-444
        case EW_MUL_SCALAR_MONT2: return mred(a, s, q, k.qinv);               // :700-727, :762-783, :811-832
        case EW_MUL_SCALAR_MONT2_ADD: return cred(c + mred(a, s, q, k.qinv), q);  // :590-609
    }
    return 0;
}

__host__ __device__ constexpr bool ew_reads_b(int op) {
    return op == EW_ADD || op == EW_ADD_NOMOD || op == EW_SUB || op == EW_SUB_NOMOD ||
           (op >= EW_MUL_BARRETT && op <= EW_MULMONT_CONSTANT) || op == EW_MULVEC || op == EW_MULVEC_ADD_NOMOD ||
           op == EW_SUB_MULMONT_SCALAR || op == EW_SUB_MULMONT_SCALAR_ADD;
}
__host__ __device__ constexpr bool ew_reads_c(int op) {
    return op == EW_MUL_BARRETT_ADD || op == EW_MUL_BARRETT_ADD_NOMOD || op == EW_MULMONT_ADD ||
           op == EW_MULMONT_ADD_NOMOD || op == EW_MULMONT_CONSTANT_ADD_NOMOD || op == EW_MULMONT_SUB ||
           op == EW_MULMONT_SUB_NOMOD || op == EW_MULVEC_ADD_NOMOD || op == EW_SUB_MULMONT_SCALAR_ADD ||
           op == EW_MUL_SCALAR_MONT2_ADD;
}

__host__ __device__ constexpr bool ew_two_scalars(int op) {
    return op == EW_ADD_SCALAR2 || op == EW_MUL_SCALAR_MONT2 || op == EW_MUL_SCALAR_MONT2_ADD;
}

template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(const EwArgs g) {
    const int j = blockIdx.y, bt = blockIdx.z;
    const int tl = g.map(j);
    const LimbConst k = load_limb_const(g.T, tl);
    u64 s = g.s[j < LG_MAX_LIMBS ? j : 0], s2 = 0;
    if (OP == EW_MUL_SCALAR) s = mform(bred_add(s, k.q, k.u0), k.q, k.u0, k.u1);
    if (OP == EW_MUL_POW2) s = g.s[0];
    if (OP == EW_AND || OP == EW_OR || OP == EW_XOR) s = g.s[0];
    if (OP == EW_MOD) {
        s = g.s[0];
        s2 = g.s[1];
    }
    const ulonglong2* pa = reinterpret_cast<const ulonglong2*>(g.a + bt * g.a_bs + j * g.a_ls);
    const ulonglong2* pb =
        ew_reads_b(OP) ? reinterpret_cast<const ulonglong2*>(g.b + bt * g.b_bs + j * g.b_ls) : nullptr;
    ulonglong2* pc = reinterpret_cast<ulonglong2*>(g.c + bt * g.c_bs + j * g.c_ls);
    const u32 n2 = g.T.N >> 1;
    const u64 sh = ew_two_scalars(OP) ? g.shi[j < LG_MAX_LIMBS ? j : 0] : 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const ulonglong2 a = pa[i];
        ulonglong2 b = make_ulonglong2(0, 0), c = make_ulonglong2(0, 0);
        if (ew_reads_b(OP)) b = pb[i];
        if (ew_reads_c(OP)) c = pc[i];
        ulonglong2 r;
        r.x = ew_apply<OP>(a.x, b.x, c.x, k, (ew_two_scalars(OP) && 2 * i >= n2) ? sh : s, s2);
        r.y = ew_apply<OP>(a.y, b.y, c.y, k, (ew_two_scalars(OP) && 2 * i + 1 >= n2) ? sh : s, s2);
        pc[i] = r;
    }
}

template <int OP>
void ew_launch(const EwArgs& a, dim3 grid, cudaStream_t st) {
    ew_kernel<OP><<<grid, 256, 0, st>>>(a);
}

typedef void (*ew_fn)(const EwArgs&, dim3, cudaStream_t);

template <int... I>
struct Seq {};
template <int N, int... I>
struct MakeSeq : MakeSeq<N - 1, N - 1, I...> {};
template <int... I>
struct MakeSeq<0, I...> {
    typedef Seq<I...> type;
};
template <int... I>
const ew_fn* ew_table(Seq<I...>) {
    static const ew_fn t[] = {&ew_launch<I>...};
    return t;
}

}  // namespace

int lg_launch_ew(int op, const EwArgs& args, int nlimbs, int batch, cudaStream_t st) {
    if (op < 0 || op >= EW_NUM_OPS) return 1;
    if (nlimbs <= 0 || batch <= 0) return 0;
    const u32 n2 = args.T.N >> 1;
    u32 bx = (n2 + 255) / 256;
    if (bx > 64) bx = 64;  // >= 2 vectors per thread at large N
    if (bx == 0) bx = 1;
    dim3 grid(bx, nlimbs, batch);
    ew_table(MakeSeq<EW_NUM_OPS>::type())[op](args, grid, st);
    lg_g_launches += 1;
    return 0;
}
