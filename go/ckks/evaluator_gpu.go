// +build cuda

// Package ckks: the evaluator's ring hot path on the B200-native engine (include/lattigpu.h).
//
// This file is added to the reference's `ckks` package.  It replaces the bodies of switchKeysInPlace, MulRelin (the
// ciphertext x ciphertext branch with a key), Relinearize, the Rescale / RescaleMany loops, SwitchKeys, permuteNTT and
// RotateHoisted (ckks/evaluator.go:933-1000, 1016-1189, 1252-1392, 1452-1558) by one call each; level, scale and degree
// bookkeeping stays the reference's Go code.  GPUCiphertext shadows Ciphertext with device-resident values
// (ring.GPUPoly); keys are uploaded once (GPUSwitchingKey).  See INTEGRATION.md for the raw-slice sites of the other
// methods.
package ckks

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../lattigo-fhe-by-go_b200/lib -llattigpu
#include "lattigpu.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"

	"github.com/ldsec/lattigo/ring"
	"github.com/ldsec/lattigo/utils"
)

func must(rc C.int) {
	if rc != 0 {
		panic("lattigpu: " + C.GoString(C.lg_last_error()))
	}
}

func cpoly(p *ring.GPUPoly) *C.lg_poly       { return (*C.lg_poly)(p.Handle()) }
func cring(g *ring.GPUContext) *C.lg_ring    { return (*C.lg_ring)(g.Handle()) }
func cgal(x *ring.GPUGalois) *C.lg_galois    { return (*C.lg_galois)(x.Handle()) }
func cstream(s *ring.GPUStream) C.lg_stream_t { return C.lg_stream_t(s.Handle()) }

// GPUCiphertext is a Ciphertext whose values live on the device.  IsNTT is always true for CKKS evaluator operands.
type GPUCiphertext struct {
	value []*ring.GPUPoly
	scale float64
}

func (ct *GPUCiphertext) Degree() uint64 { return uint64(len(ct.value) - 1) }
func (ct *GPUCiphertext) Level() uint64  { return ct.value[0].Level() }
func (ct *GPUCiphertext) Scale() float64 { return ct.scale }

// GPUSwitchingKey is evakey [beta][2] of a SwitchingKey (ckks/keygen.go:282-340) on the device, layout
// [beta][2][#Q+#P][N], NTT + Montgomery form.
type GPUSwitchingKey struct{ h *C.lg_swk }

// NewGPUSwitchingKey uploads a switching key once.
func NewGPUSwitchingKey(swk *SwitchingKey, N uint64) *GPUSwitchingKey {
	beta := len(swk.evakey)
	nQP := len(swk.evakey[0][0].Coeffs)
	flat := make([]uint64, 0, uint64(beta*2*nQP)*N)
	for i := 0; i < beta; i++ {
		for h := 0; h < 2; h++ {
			for j := 0; j < nQP; j++ {
				flat = append(flat, swk.evakey[i][h].Coeffs[j]...)
			}
		}
	}
	k := new(GPUSwitchingKey)
	must(C.lg_swk_create(C.uint64_t(N), C.int(beta), C.int(nQP), (*C.uint64_t)(unsafe.Pointer(&flat[0])), &k.h))
	runtime.SetFinalizer(k, func(k *GPUSwitchingKey) { C.lg_swk_destroy(k.h) })
	return k
}

// GPURotationKeys mirrors RotationKeys (ckks/keygen.go:24-33) with device keys and index tables.
type GPURotationKeys struct {
	evakeyRotColLeft         map[uint64]*GPUSwitchingKey
	evakeyRotColRight        map[uint64]*GPUSwitchingKey
	evakeyConjugate          *GPUSwitchingKey
	permuteNTTLeftIndex      map[uint64]*ring.GPUGalois
	permuteNTTRightIndex     map[uint64]*ring.GPUGalois
	permuteNTTConjugateIndex *ring.GPUGalois
}

// NewGPURotationKeys uploads every key and index table of a RotationKeys.
func NewGPURotationKeys(rk *RotationKeys, N uint64) *GPURotationKeys {
	g := &GPURotationKeys{
		evakeyRotColLeft: make(map[uint64]*GPUSwitchingKey), evakeyRotColRight: make(map[uint64]*GPUSwitchingKey),
		permuteNTTLeftIndex: make(map[uint64]*ring.GPUGalois), permuteNTTRightIndex: make(map[uint64]*ring.GPUGalois),
	}
	for k, key := range rk.evakeyRotColLeft {
		g.evakeyRotColLeft[k] = NewGPUSwitchingKey(key, N)
		g.permuteNTTLeftIndex[k] = ring.NewGPUGalois(rk.permuteNTTLeftIndex[k])
	}
	for k, key := range rk.evakeyRotColRight {
		g.evakeyRotColRight[k] = NewGPUSwitchingKey(key, N)
		g.permuteNTTRightIndex[k] = ring.NewGPUGalois(rk.permuteNTTRightIndex[k])
	}
	if rk.evakeyConjugate != nil {
		g.evakeyConjugate = NewGPUSwitchingKey(rk.evakeyConjugate, N)
		g.permuteNTTConjugateIndex = ring.NewGPUGalois(rk.permuteNTTConjugateIndex)
	}
	return g
}

// GPUEvaluator holds the device side of evaluator (ckks/evaluator.go:64-76): contexts Q and P, the basis extender
// and the decomposer live behind lg_ckks_eval; one stream per evaluator (one evaluator per goroutine).
type GPUEvaluator struct {
	h        *C.lg_ckks_eval
	contextQ *ring.GPUContext
	contextP *ring.GPUContext
	n        uint64
	stream   *ring.GPUStream
}

// NewGPUEvaluator is the ring part of NewEvaluator (ckks/evaluator.go:81-112).
func NewGPUEvaluator(contextQ, contextP *ring.Context) *GPUEvaluator {
	st := ring.NewGPUStream()
	e := &GPUEvaluator{contextQ: contextQ.ToGPU(st), contextP: contextP.ToGPU(st), n: contextQ.N, stream: st}
	must(C.lg_ckks_eval_create(cring(e.contextQ), cring(e.contextP), &e.h))
	runtime.SetFinalizer(e, func(e *GPUEvaluator) { C.lg_ckks_eval_destroy(e.h) })
	return e
}

// switchKeysInPlace: ckks/evaluator.go:1475-1558
func (eval *GPUEvaluator) switchKeysInPlace(level uint64, cx *ring.GPUPoly, evakey *GPUSwitchingKey, p0, p1 *ring.GPUPoly) {
	must(C.lg_ckks_switch_keys_in_place(eval.h, C.int(level), cpoly(cx), evakey.h, cpoly(p0), cpoly(p1), cstream(eval.stream)))
}

// MulRelin, ciphertext x ciphertext with an evaluation key: ckks/evaluator.go:1016-1133.  The same ciphertext for both
// operands selects the squaring branch (:1080-1085).
func (eval *GPUEvaluator) MulRelin(ct0, ct1 *GPUCiphertext, evakey *GPUSwitchingKey, ctOut *GPUCiphertext) {
	if ct0.Degree() != 1 || ct1.Degree() != 1 || ctOut.Degree() != 1 {
		panic("cannot MulRelin: input and output Ciphertexts must be of degree 1 on the fused path")
	}
	level := utils.MinUint64(utils.MinUint64(ct0.Level(), ct1.Level()), ctOut.Level()) // :1044
	ctOut.scale = ct0.scale * ct1.scale                                                  // :1046
	must(C.lg_ckks_mul_relin(eval.h, C.int(level), cpoly(ct0.value[0]), cpoly(ct0.value[1]), cpoly(ct1.value[0]),
		cpoly(ct1.value[1]), evakey.h, cpoly(ctOut.value[0]), cpoly(ctOut.value[1]), cstream(eval.stream)))
	ctOut.value[0].SetLevel(level)
	ctOut.value[1].SetLevel(level)
}

// Relinearize: ckks/evaluator.go:1144-1162
func (eval *GPUEvaluator) Relinearize(ct0 *GPUCiphertext, evakey *GPUSwitchingKey, ctOut *GPUCiphertext) {
	if ct0.Degree() != 2 {
		panic("cannot Relinearize: input Ciphertext is not of degree 2")
	}
	level := utils.MinUint64(ct0.Level(), ctOut.Level())
	ctOut.scale = ct0.scale
	must(C.lg_ckks_relinearize(eval.h, C.int(level), cpoly(ct0.value[0]), cpoly(ct0.value[1]), cpoly(ct0.value[2]), evakey.h,
		cpoly(ctOut.value[0]), cpoly(ctOut.value[1]), cstream(eval.stream)))
}

// Rescale: ckks/evaluator.go:933-968.  The threshold loop on the scale stays Go; the divisions run as one call.
func (eval *GPUEvaluator) Rescale(ct0 *GPUCiphertext, threshold float64, ctOut *GPUCiphertext) (err error) {
	if ct0.Level() == 0 {
		return errors.New("cannot Rescale: input Ciphertext already at level 0")
	}
	if ct0.Level() != ctOut.Level() {
		panic("cannot Rescale: degrees of receiver Ciphertext and input Ciphertext do not match")
	}
	if ct0 != ctOut {
		for i := range ct0.value {
			eval.contextQ.CopyLvl(ct0.Level(), ct0.value[i], ctOut.value[i])
		}
		ctOut.scale = ct0.scale
	}
	level := ctOut.Level()
	nb := 0
	for ctOut.scale >= (threshold*float64(eval.contextQ.Modulus[level]))/2 && level != 0 { // :955
		ctOut.scale /= float64(eval.contextQ.Modulus[level])
		level--
		nb++
	}
	if nb > 0 {
		must(C.lg_ckks_rescale(eval.h, C.int(ctOut.Level()+1), cpoly(ctOut.value[0]), cpoly(ctOut.value[1]), C.int(nb),
			cstream(eval.stream)))
		for i := range ctOut.value {
			ctOut.value[i].SetLevel(level) // Coeffs = Coeffs[:level+1], ring_scaling.go:113
		}
	}
	return nil
}

// RescaleMany: ckks/evaluator.go:971-1000
func (eval *GPUEvaluator) RescaleMany(ct0 *GPUCiphertext, nbRescales uint64, ctOut *GPUCiphertext) (err error) {
	if ct0.Level() < nbRescales {
		return errors.New("cannot RescaleMany: input Ciphertext level too low")
	}
	if ct0.Level() != ctOut.Level() {
		panic("cannot RescaleMany: degrees of receiver Ciphertext and input Ciphertext do not match")
	}
	if ct0 != ctOut {
		for i := range ct0.value {
			eval.contextQ.CopyLvl(ct0.Level(), ct0.value[i], ctOut.value[i])
		}
		ctOut.scale = ct0.scale
	}
	for i := uint64(0); i < nbRescales; i++ {
		ctOut.scale /= float64(eval.contextQ.Modulus[ctOut.Level()-i])
	}
	for i := range ctOut.value {
		eval.contextQ.DivRoundByLastModulusManyNTT(ctOut.value[i], nbRescales)
	}
	return nil
}

// SwitchKeys: ckks/evaluator.go:1176-1189
func (eval *GPUEvaluator) SwitchKeys(ct0 *GPUCiphertext, switchingKey *GPUSwitchingKey, ctOut *GPUCiphertext) {
	if ct0.Degree() != 1 || ctOut.Degree() != 1 {
		panic("cannot SwitchKeys: input and output Ciphertext must be of degree 1")
	}
	level := utils.MinUint64(ct0.Level(), ctOut.Level())
	ctOut.scale = ct0.scale
	must(C.lg_ckks_switch_keys(eval.h, C.int(level), cpoly(ct0.value[0]), cpoly(ct0.value[1]), switchingKey.h,
		cpoly(ctOut.value[0]), cpoly(ctOut.value[1]), cstream(eval.stream)))
}

// permuteNTT: ckks/evaluator.go:1452-1472 (in place allowed: the permuted values go through device scratch)
func (eval *GPUEvaluator) permuteNTT(ct0 *GPUCiphertext, index *ring.GPUGalois, evakey *GPUSwitchingKey, ctOut *GPUCiphertext) {
	level := utils.MinUint64(ct0.Level(), ctOut.Level())
	must(C.lg_ckks_permute_ntt(eval.h, C.int(level), cpoly(ct0.value[0]), cpoly(ct0.value[1]), cgal(index), evakey.h,
		cpoly(ctOut.value[0]), cpoly(ctOut.value[1]), cstream(eval.stream)))
}

// RotateColumns: ckks/evaluator.go:1201-1248
func (eval *GPUEvaluator) RotateColumns(ct0 *GPUCiphertext, k uint64, evakey *GPURotationKeys, ctOut *GPUCiphertext) {
	if ct0.Degree() != 1 || ctOut.Degree() != 1 {
		panic("cannot RotateColumns: input and output Ciphertext must be of degree 1")
	}
	k &= ((eval.n >> 1) - 1)
	if k == 0 {
		for i := range ct0.value {
			eval.contextQ.CopyLvl(ct0.Level(), ct0.value[i], ctOut.value[i])
		}
		ctOut.scale = ct0.scale
		return
	}
	ctOut.scale = ct0.scale
	if evakey.evakeyRotColLeft[k] != nil {
		eval.permuteNTT(ct0, evakey.permuteNTTLeftIndex[k], evakey.evakeyRotColLeft[k], ctOut)
		return
	}
	hasPow2Rotations := true
	for i := uint64(1); i < eval.n>>1; i <<= 1 {
		if evakey.evakeyRotColLeft[i] == nil || evakey.evakeyRotColRight[i] == nil {
			hasPow2Rotations = false
			break
		}
	}
	if !hasPow2Rotations {
		panic("cannot RotateColumns: specific rotation and pow2 rotations have not been generated")
	}
	if utils.HammingWeight64(k) <= utils.HammingWeight64((eval.n>>1)-k) {
		eval.rotateColumnsPow2(ct0, k, evakey.permuteNTTLeftIndex, evakey.evakeyRotColLeft, ctOut)
	} else {
		eval.rotateColumnsPow2(ct0, (eval.n>>1)-k, evakey.permuteNTTRightIndex, evakey.evakeyRotColRight, ctOut)
	}
}

// rotateColumnsPow2: ckks/evaluator.go:1402-1424
func (eval *GPUEvaluator) rotateColumnsPow2(ct0 *GPUCiphertext, k uint64, permuteNTTIndex map[uint64]*ring.GPUGalois,
	evakeyRotCol map[uint64]*GPUSwitchingKey, ctOut *GPUCiphertext) {
	evakeyIndex := uint64(1)
	level := utils.MinUint64(ct0.Level(), ctOut.Level())
	eval.contextQ.CopyLvl(level, ct0.value[0], ctOut.value[0])
	eval.contextQ.CopyLvl(level, ct0.value[1], ctOut.value[1])
	for k > 0 {
		if k&1 == 1 {
			eval.permuteNTT(ctOut, permuteNTTIndex[evakeyIndex], evakeyRotCol[evakeyIndex], ctOut)
		}
		evakeyIndex <<= 1
		k >>= 1
	}
}

// Conjugate: ckks/evaluator.go:1437-1450
func (eval *GPUEvaluator) Conjugate(ct0 *GPUCiphertext, evakey *GPURotationKeys, ctOut *GPUCiphertext) {
	if ct0.Degree() != 1 || ctOut.Degree() != 1 {
		panic("cannot Conjugate: input and output Ciphertext must be of degree 1")
	}
	if evakey.evakeyConjugate == nil {
		panic("cannot Conjugate: rows rotation key not generated")
	}
	ctOut.scale = ct0.scale
	eval.permuteNTT(ct0, evakey.permuteNTTConjugateIndex, evakey.evakeyConjugate, ctOut)
}

// RotateHoisted: ckks/evaluator.go:1252-1289.  The decomposition of ct0.value[1] (c2QiQDecomp / c2QiPDecomp, :1261-1273)
// is one device object shared by every rotation; switchKeyHoisted (:1291-1392) is one call per rotation.
func (eval *GPUEvaluator) RotateHoisted(ct0 *GPUCiphertext, rotations []uint64, rotkeys *GPURotationKeys,
	newCiphertext func(level uint64, scale float64) *GPUCiphertext) (cOut map[uint64]*GPUCiphertext) {
	var h *C.lg_hoisted
	must(C.lg_ckks_hoist(eval.h, C.int(ct0.Level()), cpoly(ct0.value[1]), &h, cstream(eval.stream)))
	defer C.lg_hoisted_destroy(h) // stream-ordered release, after the rotations issued below
	cOut = make(map[uint64]*GPUCiphertext)
	for _, i := range rotations {
		i &= (eval.n >> 1) - 1
		out := newCiphertext(ct0.Level(), ct0.Scale())
		if i == 0 {
			for u := range ct0.value {
				eval.contextQ.CopyLvl(ct0.Level(), ct0.value[u], out.value[u])
			}
		} else {
			must(C.lg_ckks_switch_key_hoisted(eval.h, h, cpoly(ct0.value[0]), cgal(rotkeys.permuteNTTLeftIndex[i]),
				rotkeys.evakeyRotColLeft[i].h, cpoly(out.value[0]), cpoly(out.value[1]), cstream(eval.stream)))
		}
		cOut[i] = out
	}
	return
}
