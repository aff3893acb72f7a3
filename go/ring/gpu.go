// +build cuda

// Package ring: cgo binding of the B200-native ring engine (include/lattigpu.h) for Lattigo v1.3.1.
//
// This file is added to the reference's `ring` package (it reads the unexported tables of Context).  It keeps the Go
// method names and argument order of ring.Context / ring.FastBasisExtender / ring.Decomposer on device-resident
// polynomials (GPUPoly), so the evaluators of ckks / bfv / dckks / dbfv drive it unchanged; see INTEGRATION.md.
// Every C entry switches to the CUDA device its handles were created on (goroutines may migrate between OS threads),
// and every op takes the stream of the GPUContext it is called on.  A non-zero status panics, as misuse of the
// reference does (ring_context.go:72,136).
//
// No Go toolchain exists in the image this engine was built in: the same entry points are exercised through ctypes by
// tests/ (lattigpu/ring.py mirrors these methods one to one) and through plain C by examples/c/ring_smoke.c.
package ring

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../lattigo-fhe-by-go_b200/lib -llattigpu
#include <stdlib.h>
#include "lattigpu.h"
*/
import "C"

import (
	"math/big"
	"runtime"
	"unsafe"
)

func must(rc C.int) {
	if rc != 0 {
		panic("lattigpu: " + C.GoString(C.lg_last_error()))
	}
}

func u64ptr(s []uint64) *C.uint64_t {
	if len(s) == 0 {
		return nil
	}
	return (*C.uint64_t)(unsafe.Pointer(&s[0]))
}

// SetDevice selects the CUDA device new handles are created on (lg_set_device).
func SetDevice(device int) { must(C.lg_set_device(C.int(device))) }

// GPUStream is a cudaStream_t; one per evaluator (the reference's "one evaluator per goroutine",
// examples/dbfv/psi/psi.go:219-233): scratch is stream-ordered, so evaluators on different streams are independent.
type GPUStream struct{ s C.lg_stream_t }

// NewGPUStream creates a non-blocking stream.
func NewGPUStream() *GPUStream {
	st := new(GPUStream)
	must(C.lg_stream_create(&st.s))
	runtime.SetFinalizer(st, func(st *GPUStream) { C.lg_stream_destroy(st.s) })
	return st
}

// Sync waits for everything issued on the stream.
func (st *GPUStream) Sync() { must(C.lg_stream_sync(st.s)) }

// GPUContext mirrors Context on the device.  The tables are the ones GenNTTParams computed in Go
// (ring_context.go:129-209), so psi is literally the reference's.
type GPUContext struct {
	*Context
	h      *C.lg_ring
	Stream *GPUStream
}

// ToGPU uploads the tables of a Context whose NTT parameters have been generated.
func (context *Context) ToGPU(stream *GPUStream) *GPUContext {
	n := len(context.Modulus)
	bred := make([]uint64, 2*n)
	for i, b := range context.bredParams {
		bred[2*i], bred[2*i+1] = b[0], b[1]
	}
	psi := make([]uint64, 0, n*int(context.N))
	psiInv := make([]uint64, 0, n*int(context.N))
	for i := range context.Modulus {
		psi = append(psi, context.nttPsi[i]...)
		psiInv = append(psiInv, context.nttPsiInv[i]...)
	}
	var rescale []uint64 // rescaleParams[j-1][i], j-major (ring_context.go:148-158)
	for j := 1; j < n; j++ {
		rescale = append(rescale, context.rescaleParams[j-1]...)
	}
	if stream == nil {
		stream = NewGPUStream()
	}
	g := &GPUContext{Context: context, Stream: stream}
	must(C.lg_ring_create_from_tables(C.uint64_t(context.N), C.int(n), u64ptr(context.Modulus), u64ptr(bred),
		u64ptr(context.mredParams), u64ptr(psi), u64ptr(psiInv), u64ptr(context.nttNInv), u64ptr(rescale), &g.h))
	runtime.SetFinalizer(g, func(g *GPUContext) { C.lg_ring_destroy(g.h) })
	return g
}

func (g *GPUContext) st() C.lg_stream_t { return g.Stream.s }
func (g *GPUContext) all() C.int         { return C.int(len(g.Modulus)) }

// GPUPoly is a device-resident Poly.  nlimbs replaces len(Coeffs): the evaluators re-slice Coeffs to drop levels
// (ring_scaling.go:113, ckks/evaluator.go:910), here the count of active limbs is metadata.
type GPUPoly struct {
	h      *C.lg_poly
	N      uint64
	nlimbs int
}

// NewPoly allocates a zero polynomial over all the moduli of the context (ring_object.go:16-23).
func (g *GPUContext) NewPoly() *GPUPoly { return g.NewPolyLvl(uint64(len(g.Modulus) - 1)) }

// NewPolyLvl allocates a zero polynomial of level+1 limbs.
func (g *GPUContext) NewPolyLvl(level uint64) *GPUPoly {
	p := &GPUPoly{N: g.N, nlimbs: int(level) + 1}
	must(C.lg_poly_create(C.uint64_t(g.N), C.int(p.nlimbs), 1, &p.h))
	runtime.SetFinalizer(p, func(p *GPUPoly) { C.lg_poly_destroy(p.h) })
	return p
}

// Level returns the index of the last active limb.
func (p *GPUPoly) Level() uint64 { return uint64(p.nlimbs - 1) }

// SetLevel re-slices the polynomial (p.Coeffs = p.Coeffs[:level+1]).
func (p *GPUPoly) SetLevel(level uint64) { p.nlimbs = int(level) + 1 }

// Upload copies a host Poly to the device.  Synchronous: cgo forbids C keeping Go pointers after return.
func (p *GPUPoly) Upload(src *Poly, stream *GPUStream) {
	for i := range src.Coeffs {
		must(C.lg_poly_upload(p.h, 0, 1, C.int(i), 1, u64ptr(src.Coeffs[i]), stream.s))
	}
	p.nlimbs = len(src.Coeffs)
}

// Download copies the active limbs back into a host Poly.
func (p *GPUPoly) Download(dst *Poly, stream *GPUStream) {
	for i := 0; i < p.nlimbs; i++ {
		must(C.lg_poly_download(p.h, 0, 1, C.int(i), 1, u64ptr(dst.Coeffs[i]), stream.s))
	}
}

// Zero sets all coefficients to zero (ring_object.go:60-67).
func (p *GPUPoly) Zero(stream *GPUStream) { must(C.lg_poly_zero(p.h, stream.s)) }

// MarshalBinary writes the reference's wire format straight from device memory (ring_object.go:161-184).
func (p *GPUPoly) MarshalBinary(stream *GPUStream) ([]byte, error) {
	n := uint64(C.lg_poly_get_data_len(p.h, C.int(p.nlimbs), 1))
	data := make([]byte, n)
	must(C.lg_poly_write_to(p.h, 0, C.int(p.nlimbs), (*C.uint8_t)(unsafe.Pointer(&data[0])), C.uint64_t(n), 1, stream.s))
	return data, nil
}

// UnmarshalBinary decodes the wire format into device memory (ring_object.go:257-274).
func (p *GPUPoly) UnmarshalBinary(data []byte, stream *GPUStream) error {
	must(C.lg_poly_decode(p.h, 0, (*C.uint8_t)(unsafe.Pointer(&data[0])), C.uint64_t(len(data)), 1, 0, stream.s))
	p.nlimbs = int(data[1])
	return nil
}

// Copy copies p0 on p1 (ring_object.go:85-101).
func (g *GPUContext) Copy(p0, p1 *GPUPoly) { must(C.lg_poly_copy(p0.h, g.all(), p1.h, g.st())) }

// CopyLvl copies the first level+1 limbs of p0 on p1 (ring_object.go:104-121).
func (g *GPUContext) CopyLvl(level uint64, p0, p1 *GPUPoly) { must(C.lg_poly_copy(p0.h, C.int(level+1), p1.h, g.st())) }

// ---- NTT (ring/ntt.go:4-29) ----

func (g *GPUContext) NTT(p1, p2 *GPUPoly) { must(C.lg_ring_ntt(g.h, g.all(), p1.h, p2.h, g.st())) }
func (g *GPUContext) NTTLvl(level uint64, p1, p2 *GPUPoly) {
	must(C.lg_ring_ntt(g.h, C.int(level+1), p1.h, p2.h, g.st()))
}
func (g *GPUContext) InvNTT(p1, p2 *GPUPoly) { must(C.lg_ring_invntt(g.h, g.all(), p1.h, p2.h, g.st())) }
func (g *GPUContext) InvNTTLvl(level uint64, p1, p2 *GPUPoly) {
	must(C.lg_ring_invntt(g.h, C.int(level+1), p1.h, p2.h, g.st()))
}

// ---- coefficient-wise ops (ring/ring.go) ----

// Add: ring.go:10-29
func (g *GPUContext) Add(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_add(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}
func (g *GPUContext) AddLvl(level uint64, p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_add(g.h, C.int(level+1), p1.h, p2.h, p3.h, g.st()))
}

// AddNoMod: ring.go:32-51
func (g *GPUContext) AddNoMod(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_add_nomod(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}
func (g *GPUContext) AddNoModLvl(level uint64, p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_add_nomod(g.h, C.int(level+1), p1.h, p2.h, p3.h, g.st()))
}

// Sub: ring.go:54-73
func (g *GPUContext) Sub(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_sub(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}
func (g *GPUContext) SubLvl(level uint64, p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_sub(g.h, C.int(level+1), p1.h, p2.h, p3.h, g.st()))
}

// SubNoMod: ring.go:76-97
func (g *GPUContext) SubNoMod(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_sub_nomod(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}
func (g *GPUContext) SubNoModLvl(level uint64, p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_sub_nomod(g.h, C.int(level+1), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffs: ring.go:187-195
func (g *GPUContext) MulCoeffs(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsAndAdd: ring.go:198-206
func (g *GPUContext) MulCoeffsAndAdd(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_and_add(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsAndAddNoMod: ring.go:209-217
func (g *GPUContext) MulCoeffsAndAddNoMod(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_and_add_nomod(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsConstant: ring.go:335-343
func (g *GPUContext) MulCoeffsConstant(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_constant(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsMontgomery: ring.go:221-243
func (g *GPUContext) MulCoeffsMontgomery(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}
func (g *GPUContext) MulCoeffsMontgomeryLvl(level uint64, p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery(g.h, C.int(level+1), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsMontgomeryAndAdd: ring.go:247-269
func (g *GPUContext) MulCoeffsMontgomeryAndAdd(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery_and_add(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}
func (g *GPUContext) MulCoeffsMontgomeryAndAddLvl(level uint64, p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery_and_add(g.h, C.int(level+1), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsMontgomeryAndAddNoMod: ring.go:273-295
func (g *GPUContext) MulCoeffsMontgomeryAndAddNoMod(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery_and_add_nomod(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}
func (g *GPUContext) MulCoeffsMontgomeryAndAddNoModLvl(level uint64, p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery_and_add_nomod(g.h, C.int(level+1), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsMontgomeryAndSub: ring.go:311-319
func (g *GPUContext) MulCoeffsMontgomeryAndSub(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery_and_sub(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsMontgomeryAndSubNoMod: ring.go:323-331
func (g *GPUContext) MulCoeffsMontgomeryAndSubNoMod(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery_and_sub_nomod(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsMontgomeryConstant: ring.go:346-355
func (g *GPUContext) MulCoeffsMontgomeryConstant(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery_constant(g.h, g.all(), p1.h, p2.h, p3.h, g.st()))
}

// MulCoeffsMontgomeryConstantAndAddNoModLvl: ring.go:298-308
func (g *GPUContext) MulCoeffsMontgomeryConstantAndAddNoModLvl(level uint64, p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_coeffs_montgomery_constant_and_add_nomod(g.h, C.int(level+1), p1.h, p2.h, p3.h, g.st()))
}

// Neg: ring.go:100-119
func (g *GPUContext) Neg(p1, p2 *GPUPoly) {
	must(C.lg_ring_neg(g.h, g.all(), p1.h, p2.h, g.st()))
}
func (g *GPUContext) NegLvl(level uint64, p1, p2 *GPUPoly) {
	must(C.lg_ring_neg(g.h, C.int(level+1), p1.h, p2.h, g.st()))
}

// Reduce: ring.go:122-143
func (g *GPUContext) Reduce(p1, p2 *GPUPoly) {
	must(C.lg_ring_reduce(g.h, g.all(), p1.h, p2.h, g.st()))
}
func (g *GPUContext) ReduceLvl(level uint64, p1, p2 *GPUPoly) {
	must(C.lg_ring_reduce(g.h, C.int(level+1), p1.h, p2.h, g.st()))
}

// MForm: ring.go:583-607
func (g *GPUContext) MForm(p1, p2 *GPUPoly) {
	must(C.lg_ring_mform(g.h, g.all(), p1.h, p2.h, g.st()))
}
func (g *GPUContext) MFormLvl(level uint64, p1, p2 *GPUPoly) {
	must(C.lg_ring_mform(g.h, C.int(level+1), p1.h, p2.h, g.st()))
}

// InvMForm: ring.go:610-619
func (g *GPUContext) InvMForm(p1, p2 *GPUPoly) {
	must(C.lg_ring_invmform(g.h, g.all(), p1.h, p2.h, g.st()))
}

// BitReverse: ring.go:749-772
func (g *GPUContext) BitReverse(p1, p2 *GPUPoly) {
	must(C.lg_ring_bitreverse(g.h, g.all(), p1.h, p2.h, g.st()))
}

// MulPoly, MulPolyMontgomery: ring.go:358-380; MulPolyNaive, MulPolyNaiveMontgomery: ring.go:383-437
func (g *GPUContext) MulPoly(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_poly(g.h, p1.h, p2.h, p3.h, 0, g.st()))
}
func (g *GPUContext) MulPolyMontgomery(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_poly(g.h, p1.h, p2.h, p3.h, 1, g.st()))
}
func (g *GPUContext) MulPolyNaive(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_poly_naive(g.h, p1.h, p2.h, p3.h, 0, g.st()))
}
func (g *GPUContext) MulPolyNaiveMontgomery(p1, p2, p3 *GPUPoly) {
	must(C.lg_ring_mul_poly_naive(g.h, p1.h, p2.h, p3.h, 1, g.st()))
}

// Exp: ring.go:441-464 (p1 is left in the NTT domain and p2 ends as InvNTT(p1), as the reference's last line does)
func (g *GPUContext) Exp(p1 *GPUPoly, e uint64, p2 *GPUPoly) {
	must(C.lg_ring_exp(g.h, p1.h, C.uint64_t(e), p2.h, g.st()))
}

// Shift: ring.go:575-580
func (g *GPUContext) Shift(p1 *GPUPoly, n uint64, p2 *GPUPoly) {
	must(C.lg_ring_shift(g.h, p1.h, C.uint64_t(n), p2.h, g.st()))
}

// Rotate: ring.go:775-800 (the reference writes p1's coefficients; p2 is never touched)
func (g *GPUContext) Rotate(p1 *GPUPoly, n uint64, p2 *GPUPoly) {
	must(C.lg_ring_rotate(g.h, p1.h, C.uint64_t(n), g.st()))
}

// Equal, EqualLvl: ring_context.go:424-467 (both operands are reduced in place; the call synchronises the stream)
func (g *GPUContext) Equal(p1, p2 *GPUPoly) bool {
	var eq C.int
	must(C.lg_ring_equal(g.h, g.all(), p1.h, p2.h, &eq, g.st()))
	return eq != 0
}
func (g *GPUContext) EqualLvl(level uint64, p1, p2 *GPUPoly) bool {
	var eq C.int
	must(C.lg_ring_equal(g.h, C.int(level+1), p1.h, p2.h, &eq, g.st()))
	return eq != 0
}

// Mod, AND, OR, XOR: ring.go:146-184
func (g *GPUContext) Mod(p1 *GPUPoly, m uint64, p2 *GPUPoly) {
	must(C.lg_ring_mod(g.h, g.all(), p1.h, C.uint64_t(m), p2.h, g.st()))
}
func (g *GPUContext) AND(p1 *GPUPoly, m uint64, p2 *GPUPoly) {
	must(C.lg_ring_and(g.h, g.all(), p1.h, C.uint64_t(m), p2.h, g.st()))
}
func (g *GPUContext) OR(p1 *GPUPoly, m uint64, p2 *GPUPoly) {
	must(C.lg_ring_or(g.h, g.all(), p1.h, C.uint64_t(m), p2.h, g.st()))
}
func (g *GPUContext) XOR(p1 *GPUPoly, m uint64, p2 *GPUPoly) {
	must(C.lg_ring_xor(g.h, g.all(), p1.h, C.uint64_t(m), p2.h, g.st()))
}

// repeat returns the word s once per limb; bigmod the residues of a big.Int (the reduction stays in Go, ring.go:539-553).
func (g *GPUContext) repeat(s uint64, nl int) []uint64 {
	out := make([]uint64, nl)
	for i := range out {
		out[i] = s
	}
	return out
}
func (g *GPUContext) bigmod(scalar *big.Int, nl int) []uint64 {
	out := make([]uint64, nl)
	tmp := new(big.Int)
	for i := 0; i < nl; i++ {
		out[i] = tmp.Mod(scalar, NewUint(g.Modulus[i])).Uint64()
	}
	return out
}

// AddScalar / SubScalar and their Bigint forms write into p1 itself, as the reference does (ring.go:467-510).
func (g *GPUContext) AddScalar(p1 *GPUPoly, scalar uint64, p2 *GPUPoly) {
	must(C.lg_ring_add_scalar(g.h, g.all(), p1.h, u64ptr(g.repeat(scalar, len(g.Modulus))), g.st()))
}
func (g *GPUContext) AddScalarBigint(p1 *GPUPoly, scalar *big.Int, p2 *GPUPoly) {
	must(C.lg_ring_add_scalar(g.h, g.all(), p1.h, u64ptr(g.bigmod(scalar, len(g.Modulus))), g.st()))
}
func (g *GPUContext) SubScalar(p1 *GPUPoly, scalar uint64, p2 *GPUPoly) {
	must(C.lg_ring_sub_scalar(g.h, g.all(), p1.h, u64ptr(g.repeat(scalar, len(g.Modulus))), g.st()))
}
func (g *GPUContext) SubScalarBigint(p1 *GPUPoly, scalar *big.Int, p2 *GPUPoly) {
	must(C.lg_ring_sub_scalar(g.h, g.all(), p1.h, u64ptr(g.bigmod(scalar, len(g.Modulus))), g.st()))
}

// MulScalar and variants: ring.go:513-572
func (g *GPUContext) MulScalar(p1 *GPUPoly, scalar uint64, p2 *GPUPoly) {
	g.MulScalarLvl(uint64(len(g.Modulus)-1), p1, scalar, p2)
}
func (g *GPUContext) MulScalarLvl(level uint64, p1 *GPUPoly, scalar uint64, p2 *GPUPoly) {
	must(C.lg_ring_mul_scalar(g.h, C.int(level+1), p1.h, u64ptr(g.repeat(scalar, int(level)+1)), p2.h, g.st()))
}
func (g *GPUContext) MulScalarBigint(p1 *GPUPoly, scalar *big.Int, p2 *GPUPoly) {
	g.MulScalarBigintLvl(uint64(len(g.Modulus)-1), p1, scalar, p2)
}
func (g *GPUContext) MulScalarBigintLvl(level uint64, p1 *GPUPoly, scalar *big.Int, p2 *GPUPoly) {
	must(C.lg_ring_mul_scalar(g.h, C.int(level+1), p1.h, u64ptr(g.bigmod(scalar, int(level)+1)), p2.h, g.st()))
}

// MulByPow2 / MulByPow2Lvl: ring.go:629-653
func (g *GPUContext) MulByPow2(p1 *GPUPoly, pow2 uint64, p2 *GPUPoly) {
	must(C.lg_ring_mul_by_pow2(g.h, g.all(), p1.h, C.uint64_t(pow2), p2.h, g.st()))
}
func (g *GPUContext) MulByPow2Lvl(level uint64, p1 *GPUPoly, pow2 uint64, p2 *GPUPoly) {
	must(C.lg_ring_mul_by_pow2(g.h, C.int(level+1), p1.h, C.uint64_t(pow2), p2.h, g.st()))
}

// MultByMonomial: ring.go:663-723
func (g *GPUContext) MultByMonomial(p1 *GPUPoly, monomialDeg uint64, p2 *GPUPoly) {
	must(C.lg_ring_mult_by_monomial(g.h, g.all(), p1.h, C.uint64_t(monomialDeg), p2.h, g.st()))
}

// MulByVectorMontgomery and MulByVectorMontgomeryAndAddNoMod: ring.go:726-745; vector is a one-limb device polynomial
func (g *GPUContext) MulByVectorMontgomery(p1 *GPUPoly, vector *GPUPoly, p2 *GPUPoly) {
	must(C.lg_ring_mul_by_vector_montgomery(g.h, g.all(), p1.h, vector.h, p2.h, g.st()))
}
func (g *GPUContext) MulByVectorMontgomeryAndAddNoMod(p1 *GPUPoly, vector *GPUPoly, p2 *GPUPoly) {
	must(C.lg_ring_mul_by_vector_montgomery_and_add_nomod(g.h, g.all(), p1.h, vector.h, p2.h, g.st()))
}

// MulPolyMontgomery: ring.go:369-384 (p1 in Montgomery form): NTT both, multiply, InvNTT
func (g *GPUContext) MulPolyMontgomery(p1, p2, p3 *GPUPoly) {
	a, b := g.NewPoly(), g.NewPoly()
	g.NTT(p1, a)
	g.NTT(p2, b)
	g.MulCoeffsMontgomery(a, b, p3)
	g.InvNTT(p3, p3)
}

// MulPoly: ring.go:358-366
func (g *GPUContext) MulPoly(p1, p2, p3 *GPUPoly) {
	a, b := g.NewPoly(), g.NewPoly()
	g.NTT(p1, a)
	g.NTT(p2, b)
	g.MulCoeffs(a, b, p3)
	g.InvNTT(p3, p3)
}

// ---- Galois automorphisms (ring/ring_galois.go) ----

// GPUGalois is the index table of PermuteNTTIndex on the device (ring_galois.go:29-50).
type GPUGalois struct{ h *C.lg_galois }

// NewGPUGalois uploads an index computed by PermuteNTTIndex.
func NewGPUGalois(index []uint64) *GPUGalois {
	x := new(GPUGalois)
	must(C.lg_galois_create_from_index(u64ptr(index), C.uint64_t(len(index)), &x.h))
	runtime.SetFinalizer(x, func(x *GPUGalois) { C.lg_galois_destroy(x.h) })
	return x
}

// PermuteNTTWithIndexGPU: ring_galois.go:89-101 (not in place)
func PermuteNTTWithIndexGPU(polIn *GPUPoly, index *GPUGalois, polOut *GPUPoly, stream *GPUStream) {
	must(C.lg_ring_permute_ntt_with_index(C.int(polIn.nlimbs), polIn.h, index.h, polOut.h, stream.s))
}

// PermuteNTTGPU: ring_galois.go:55-84 (not in place)
func PermuteNTTGPU(polIn *GPUPoly, gen uint64, polOut *GPUPoly, stream *GPUStream) {
	must(C.lg_ring_permute_ntt(C.int(polIn.nlimbs), polIn.h, C.uint64_t(gen), polOut.h, stream.s))
}

// Permute: ring_galois.go:106-127 (coefficient domain, not in place)
func (g *GPUContext) Permute(polIn *GPUPoly, gen uint64, polOut *GPUPoly) {
	must(C.lg_ring_permute(g.h, g.all(), polIn.h, C.uint64_t(gen), polOut.h, g.st()))
}

// ---- RNS rescaling (ring/ring_scaling.go:9-164); the last limb(s) are dropped as the reference re-slices Coeffs ----

func (g *GPUContext) DivFloorByLastModulusNTT(p0 *GPUPoly) {
	must(C.lg_ring_div_floor_by_last_modulus_ntt(g.h, C.int(p0.nlimbs), p0.h, g.st()))
	p0.nlimbs--
}
func (g *GPUContext) DivFloorByLastModulus(p0 *GPUPoly) {
	must(C.lg_ring_div_floor_by_last_modulus(g.h, C.int(p0.nlimbs), p0.h, g.st()))
	p0.nlimbs--
}
func (g *GPUContext) DivFloorByLastModulusManyNTT(p0 *GPUPoly, nbRescales uint64) {
	must(C.lg_ring_div_floor_by_last_modulus_many_ntt(g.h, C.int(p0.nlimbs), p0.h, C.int(nbRescales), g.st()))
	p0.nlimbs -= int(nbRescales)
}
func (g *GPUContext) DivFloorByLastModulusMany(p0 *GPUPoly, nbRescales uint64) {
	must(C.lg_ring_div_floor_by_last_modulus_many(g.h, C.int(p0.nlimbs), p0.h, C.int(nbRescales), g.st()))
	p0.nlimbs -= int(nbRescales)
}
func (g *GPUContext) DivRoundByLastModulusNTT(p0 *GPUPoly) {
	must(C.lg_ring_div_round_by_last_modulus_ntt(g.h, C.int(p0.nlimbs), p0.h, g.st()))
	p0.nlimbs--
}
func (g *GPUContext) DivRoundByLastModulus(p0 *GPUPoly) {
	must(C.lg_ring_div_round_by_last_modulus(g.h, C.int(p0.nlimbs), p0.h, g.st()))
	p0.nlimbs--
}
func (g *GPUContext) DivRoundByLastModulusManyNTT(p0 *GPUPoly, nbRescales uint64) {
	must(C.lg_ring_div_round_by_last_modulus_many_ntt(g.h, C.int(p0.nlimbs), p0.h, C.int(nbRescales), g.st()))
	p0.nlimbs -= int(nbRescales)
}
func (g *GPUContext) DivRoundByLastModulusMany(p0 *GPUPoly, nbRescales uint64) {
	must(C.lg_ring_div_round_by_last_modulus_many(g.h, C.int(p0.nlimbs), p0.h, C.int(nbRescales), g.st()))
	p0.nlimbs -= int(nbRescales)
}

// ---- FastBasisExtender (ring/ring_basis_extension.go:9-350) ----

// GPUFastBasisExtender: parameters are recomputed natively from the moduli; they are canonical residues, identical to
// the math/big values of ring_basis_extension.go:76-142.
type GPUFastBasisExtender struct {
	h      *C.lg_extender
	Q, P   *GPUContext
	stream *GPUStream
}

func NewGPUFastBasisExtender(contextQ, contextP *GPUContext) *GPUFastBasisExtender {
	be := &GPUFastBasisExtender{Q: contextQ, P: contextP, stream: contextQ.Stream}
	must(C.lg_extender_create(contextQ.h, contextP.h, &be.h))
	runtime.SetFinalizer(be, func(be *GPUFastBasisExtender) { C.lg_extender_destroy(be.h) })
	return be
}
func (be *GPUFastBasisExtender) ModUpSplitQP(level uint64, p1, p2 *GPUPoly) {
	must(C.lg_extender_modup_split_qp(be.h, C.int(level), p1.h, p2.h, be.stream.s))
}
func (be *GPUFastBasisExtender) ModUpSplitPQ(level uint64, p1, p2 *GPUPoly) {
	must(C.lg_extender_modup_split_pq(be.h, C.int(level), p1.h, p2.h, be.stream.s))
}
func (be *GPUFastBasisExtender) ModDownNTTPQ(level uint64, p1, p2 *GPUPoly) {
	must(C.lg_extender_moddown_ntt_pq(be.h, C.int(level), p1.h, p2.h, be.stream.s))
}
func (be *GPUFastBasisExtender) ModDownSplitedNTTPQ(level uint64, p1Q, p1P, p2 *GPUPoly) {
	must(C.lg_extender_moddown_splited_ntt_pq(be.h, C.int(level), p1Q.h, p1P.h, p2.h, be.stream.s))
}
func (be *GPUFastBasisExtender) ModDownPQ(level uint64, p1, p2 *GPUPoly) {
	must(C.lg_extender_moddown_pq(be.h, C.int(level), p1.h, p2.h, be.stream.s))
}
func (be *GPUFastBasisExtender) ModDownSplitedPQ(level uint64, p1Q, p1P, p2 *GPUPoly) {
	must(C.lg_extender_moddown_splited_pq(be.h, C.int(level), p1Q.h, p1P.h, p2.h, be.stream.s))
}
func (be *GPUFastBasisExtender) ModDownSplitedQP(levelQ, levelP uint64, p1Q, p1P, p2 *GPUPoly) {
	must(C.lg_extender_moddown_splited_qp(be.h, C.int(levelQ), C.int(levelP), p1Q.h, p1P.h, p2.h, be.stream.s))
}

// ---- Decomposer (ring/ring_basis_extension.go:398-713) ----

type GPUDecomposer struct {
	h      *C.lg_decomposer
	stream *GPUStream
}

func NewGPUDecomposer(N uint64, Q, P []uint64, stream *GPUStream) *GPUDecomposer {
	d := &GPUDecomposer{stream: stream}
	must(C.lg_decomposer_create(C.uint64_t(N), u64ptr(Q), C.int(len(Q)), u64ptr(P), C.int(len(P)), &d.h))
	runtime.SetFinalizer(d, func(d *GPUDecomposer) { C.lg_decomposer_destroy(d.h) })
	return d
}
func (d *GPUDecomposer) Xalpha() []uint64 {
	out := make([]uint64, int(C.lg_decomposer_beta(d.h)))
	for i := range out {
		out[i] = uint64(C.lg_decomposer_xalpha(d.h, C.int(i)))
	}
	return out
}
func (d *GPUDecomposer) Decompose(level, crtDecompLevel uint64, p0, p1 *GPUPoly) {
	must(C.lg_decomposer_decompose(d.h, C.int(level), C.int(crtDecompLevel), p0.h, p1.h, d.stream.s))
}
func (d *GPUDecomposer) DecomposeAndSplit(level, crtDecompLevel uint64, p0, p1Q, p1P *GPUPoly) {
	must(C.lg_decomposer_decompose_and_split(d.h, C.int(level), C.int(crtDecompLevel), p0.h, p1Q.h, p1P.h, d.stream.s))
}

// ---- handles other packages need ----

// Handle exposes the C handle of a polynomial / context to the evaluator shims of ckks and bfv.
func (p *GPUPoly) Handle() unsafe.Pointer      { return unsafe.Pointer(p.h) }
func (g *GPUContext) Handle() unsafe.Pointer   { return unsafe.Pointer(g.h) }
func (x *GPUGalois) Handle() unsafe.Pointer    { return unsafe.Pointer(x.h) }
func (st *GPUStream) Handle() unsafe.Pointer   { return unsafe.Pointer(st.s) }
