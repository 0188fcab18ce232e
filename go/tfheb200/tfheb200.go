// Package tfheb200 is the thin cgo shim between go-tfhe's Go types and the B200 engine's C ABI
// (include/tfhe_b200.h, libtfhe_b200.so).  It flattens the pointer-rich reference types
// ([]*tlwe.TLWELv0, []*trgsw.TRGSWLv1FFT, ...) into contiguous buffers, makes ONE C call per batch,
// and scatters the results into freshly allocated reference types, so gates.* / evaluator.* keep
// their signatures.  Errors become panics, matching the reference's convention.
//
// NOTE: the build image has no Go toolchain; the file is reviewed by reading and mirrored line for line by
// go-tfhe_b200/{cloudkey,evaluator,gates}.py, which the tests exercise against the same C ABI.  INTEGRATION.md lists
// what to run (go vet, the reference's own gate tests with -tags b200) the first time a toolchain is at hand.
package tfheb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../go-tfhe_b200/lib -ltfhe_b200 -Wl,-rpath,${SRCDIR}/../../go-tfhe_b200/lib
#include <stdlib.h>
#include "tfhe_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"

	"github.com/thedonutfactory/go-tfhe/cloudkey"
	"github.com/thedonutfactory/go-tfhe/params"
	"github.com/thedonutfactory/go-tfhe/tlwe"
	"github.com/thedonutfactory/go-tfhe/trgsw"
	"github.com/thedonutfactory/go-tfhe/trlwe"
)

// Op mirrors tfhe_op.
type Op uint8

const (
	NAND Op = iota
	AND
	OR
	XOR
	XNOR
	NOR
	ANDNY
	ANDYN
	ORNY
	ORYN
	MUX
	NOT
	COPY
)

// Engine owns one tfhe_ctx — by default a multi-device context over every visible GPU (tfhe_ctx_create_multi), so that
// one gates.Batch* call fans out over the node the way trgsw.BatchBlindRotate fans out over goroutines — with a cloud
// key resident on every device.
type Engine struct {
	mu     sync.Mutex // one call at a time per context
	ctx    *C.tfhe_ctx
	n      int // TLWELv0.N
	bigN   int // TRGSWLv1.N
	hasKSK bool
	kskID  **tlwe.TLWELv0 // identity of the key-switching key that is loaded (nil: none)
}

// Engines are cached per cloud key.  The maps hold strong references, so an engine (and its GPU memory) lives until
// Release / ReleaseAll is called; there is deliberately no finalizer.
var (
	enginesMu   sync.Mutex
	engines     = map[*cloudkey.CloudKey]*Engine{}
	partEngines = map[**trgsw.TRGSWLv1FFT]*Engine{}
	// Devices selects the GPUs new engines use: nil = every visible GPU.
	Devices []int
)

// For returns (creating and uploading on first use) the engine bound to ck.
// gates.* call this, so user code keeps passing *cloudkey.CloudKey exactly as before.
func For(ck *cloudkey.CloudKey) *Engine {
	enginesMu.Lock()
	defer enginesMu.Unlock()
	if e, ok := engines[ck]; ok {
		return e
	}
	e := New(ck, Devices...)
	engines[ck] = e
	return e
}

// ForParts serves the evaluator.* entry points, which receive the pieces of a CloudKey instead of the struct
// (evaluator/evaluator.go:139).  Engines are cached by the identity of the bootstrapping-key slice; the other pieces
// are compared on every call and (re)uploaded when they differ from what the engine holds — an engine first created by
// BlindRotateAssign (no key-switching key) gets its KSK the first time BootstrapAssign asks for one.
func ForParts(bsk []*trgsw.TRGSWLv1FFT, ksk []*tlwe.TLWELv0, offset params.Torus, testvec *trlwe.TRLWELv1) *Engine {
	if len(bsk) == 0 {
		panic("tfheb200: empty bootstrapping key")
	}
	enginesMu.Lock()
	defer enginesMu.Unlock()
	if testvec == nil {
		testvec = trlwe.NewTRLWELv1()
	}
	ck := &cloudkey.CloudKey{DecompositionOffset: offset, BlindRotateTestvec: testvec, KeySwitchingKey: ksk,
		BootstrappingKey: bsk}
	key := &bsk[0]
	if e, ok := partEngines[key]; ok {
		if len(ksk) > 0 && (!e.hasKSK || e.kskID != &ksk[0]) {
			e.load(ck) // same BSK, new or first KSK: load the pair again (the engine never mixes keys)
		}
		return e
	}
	e := New(ck, Devices...)
	partEngines[key] = e
	return e
}

// Release destroys the engine cached for ck (and its copies of the key on the GPUs).
func Release(ck *cloudkey.CloudKey) {
	enginesMu.Lock()
	defer enginesMu.Unlock()
	if e, ok := engines[ck]; ok {
		delete(engines, ck)
		e.Close()
	}
	if len(ck.BootstrappingKey) > 0 {
		if e, ok := partEngines[&ck.BootstrappingKey[0]]; ok {
			delete(partEngines, &ck.BootstrappingKey[0])
			e.Close()
		}
	}
}

// ReleaseAll destroys every cached engine.
func ReleaseAll() {
	enginesMu.Lock()
	defer enginesMu.Unlock()
	for k, e := range engines {
		delete(engines, k)
		e.Close()
	}
	for k, e := range partEngines {
		delete(partEngines, k)
		e.Close()
	}
}

// Close frees the context; the Engine must not be used afterwards.
func (e *Engine) Close() {
	e.mu.Lock()
	defer e.mu.Unlock()
	if e.ctx != nil {
		C.tfhe_ctx_destroy(e.ctx)
		e.ctx = nil
	}
}

func curParams() (C.tfhe_params, int, int) {
	g, l0 := params.GetTRGSWLv1(), params.GetTLWELv0()
	return C.tfhe_params{n: C.int32_t(l0.N), N: C.int32_t(g.N), L: C.int32_t(g.L), bgbit: C.int32_t(g.BGBIT),
		basebit: C.int32_t(g.BASEBIT), iks_t: C.int32_t(g.IKS_T)}, l0.N, g.N
}

func newEngine(devices []int) *Engine {
	p, n, bigN := curParams()
	var ctx *C.tfhe_ctx
	var rc C.int
	if len(devices) == 0 {
		rc = C.tfhe_ctx_create_multi(&p, 0, nil, &ctx) // every visible GPU
	} else {
		devs := make([]C.int, len(devices))
		for i, d := range devices {
			devs[i] = C.int(d)
		}
		rc = C.tfhe_ctx_create_multi(&p, C.int(len(devs)), &devs[0], &ctx)
	}
	if rc != 0 {
		panic("tfhe_ctx_create_multi: " + C.GoString(C.tfhe_last_error(nil)))
	}
	return &Engine{ctx: ctx, n: n, bigN: bigN}
}

// New flattens ck (cloudkey/cloudkey.go:16-21) and uploads it to `devices` (none given: every visible GPU).  The key
// goes to the first device once and is replicated to the others by peer copies inside the library.
func New(ck *cloudkey.CloudKey, devices ...int) *Engine {
	e := newEngine(devices)
	e.load(ck)
	return e
}

// load flattens and uploads ck; a CloudKey without a KeySwitchingKey (cloudkey.NewCloudKeyNoKSK, or the pieces handed to
// BlindRotateAssign) is uploaded with a NULL ksk: the engine then serves blind rotations only.
func (e *Engine) load(ck *cloudkey.CloudKey) {
	g, l0 := params.GetTRGSWLv1(), params.GetTLWELv0()
	// BootstrappingKey []*TRGSWLv1FFT -> [n][2L][2][N] float64, reference FourierPoly layout untouched
	bsk := make([]float64, 0, l0.N*2*g.L*2*g.N)
	for _, row := range ck.BootstrappingKey {
		for _, t := range row.TRLWEFFT {
			bsk = append(bsk, t.A.Coeffs...)
			bsk = append(bsk, t.B.Coeffs...)
		}
	}
	// KeySwitchingKey []*TLWELv0 (index base*t*i + base*j + k) -> [N*t*base][n+1] uint32
	var pksk *C.uint32_t
	var ksk []uint32
	if len(ck.KeySwitchingKey) > 0 {
		ksk = make([]uint32, 0, len(ck.KeySwitchingKey)*(l0.N+1))
		for _, row := range ck.KeySwitchingKey {
			ksk = appendTorus(ksk, row.P)
		}
		pksk = (*C.uint32_t)(&ksk[0])
	}
	tv := appendTorus(appendTorus(make([]uint32, 0, 2*g.N), ck.BlindRotateTestvec.A), ck.BlindRotateTestvec.B)
	e.check(C.tfhe_ctx_load_cloudkey(e.ctx, C.uint32_t(ck.DecompositionOffset), (*C.double)(&bsk[0]), pksk,
		(*C.uint32_t)(&tv[0])), "tfhe_ctx_load_cloudkey")
	runtime.KeepAlive(ksk)
	e.hasKSK = len(ck.KeySwitchingKey) > 0
	e.kskID = nil
	if e.hasKSK {
		e.kskID = &ck.KeySwitchingKey[0]
	}
}

// NewCloudKeyOnDevice is cloudkey.NewCloudKey (cloudkey/cloudkey.go:24-31) with the key material generated in GPU
// memory (tfhe_ctx_generate_cloudkey): the returned CloudKey holds the same fields in the same formats as the Go
// generator's, and the engine that made it is registered for it, so gates.* use it without a second upload.
// seed == 0 (what production code passes) keys the generator from the operating system's entropy source; a non-zero
// seed makes the key a reproducible function of (secret key, seed) for tests.  The secret key is read by this call only.
func NewCloudKeyOnDevice(keyLv0, keyLv1 []params.Torus, seed uint64, devices ...int) *cloudkey.CloudKey {
	g, l0 := params.GetTRGSWLv1(), params.GetTLWELv0()
	e := newEngine(devices)
	rows := g.N * g.IKS_T * (1 << g.BASEBIT)
	bsk := make([]float64, l0.N*2*g.L*2*g.N)
	ksk := make([]uint32, rows*(l0.N+1))
	tv := make([]uint32, 2*g.N)
	var off C.uint32_t
	e.check(C.tfhe_ctx_generate_cloudkey(e.ctx, (*C.uint32_t)(unsafe.Pointer(&keyLv0[0])), (*C.uint32_t)(unsafe.Pointer(&keyLv1[0])),
		C.double(params.KSKAlpha()), C.double(params.BSKAlpha()), C.uint64_t(seed), 1, &off, (*C.double)(&bsk[0]),
		(*C.uint32_t)(&ksk[0]), (*C.uint32_t)(&tv[0])), "tfhe_ctx_generate_cloudkey")
	ck := &cloudkey.CloudKey{DecompositionOffset: params.Torus(off), BlindRotateTestvec: trlwe.NewTRLWELv1()}
	for i := 0; i < g.N; i++ {
		ck.BlindRotateTestvec.A[i], ck.BlindRotateTestvec.B[i] = params.Torus(tv[i]), params.Torus(tv[g.N+i])
	}
	ck.KeySwitchingKey = make([]*tlwe.TLWELv0, rows)
	for r := range ck.KeySwitchingKey {
		t := tlwe.NewTLWELv0()
		for w := 0; w <= l0.N; w++ {
			t.P[w] = params.Torus(ksk[r*(l0.N+1)+w])
		}
		ck.KeySwitchingKey[r] = t
	}
	ck.BootstrappingKey = make([]*trgsw.TRGSWLv1FFT, l0.N)
	for i := range ck.BootstrappingKey {
		row := &trgsw.TRGSWLv1FFT{TRLWEFFT: make([]trgsw.TRLWELv1FFT, 2*g.L)}
		for r := range row.TRLWEFFT {
			o := ((i*2*g.L + r) * 2) * g.N
			row.TRLWEFFT[r].A.Coeffs = bsk[o : o+g.N : o+g.N]
			row.TRLWEFFT[r].B.Coeffs = bsk[o+g.N : o+2*g.N : o+2*g.N]
		}
		ck.BootstrappingKey[i] = row
	}
	e.hasKSK = true
	e.kskID = &ck.KeySwitchingKey[0]
	enginesMu.Lock()
	engines[ck] = e
	enginesMu.Unlock()
	return ck
}

// DeviceCount reports how many GPUs serve this engine.
func (e *Engine) DeviceCount() int { return int(C.tfhe_ctx_device_count(e.ctx)) }

func appendTorus(dst []uint32, src []params.Torus) []uint32 {
	// params.Torus is uint32 (params/params.go:27): same memory layout
	return append(dst, unsafe.Slice((*uint32)(unsafe.Pointer(&src[0])), len(src))...)
}

func (e *Engine) check(rc C.int, what string) {
	if rc != 0 {
		panic(fmt.Sprintf("%s failed (%d): %s", what, int(rc), C.GoString(C.tfhe_last_error(e.ctx))))
	}
}

func (e *Engine) flatten(cts []*tlwe.TLWELv0) []uint32 {
	out := make([]uint32, 0, len(cts)*(e.n+1))
	for _, c := range cts {
		out = appendTorus(out, c.P)
	}
	return out
}

func (e *Engine) unflatten(buf []uint32, count int) []*tlwe.TLWELv0 {
	res := make([]*tlwe.TLWELv0, count)
	for i := range res {
		c := tlwe.NewTLWELv0() // fresh allocation per result, as gates.bootstrap does (gates/gates.go:137)
		for j := range c.P {
			c.P[j] = params.Torus(buf[i*(e.n+1)+j])
		}
		res[i] = c
	}
	return res
}

// GateBatch: one opcode per gate (or a single opcode for all); c may be nil unless an op is MUX.
// Replaces the bodies of gates.{NAND..ORYN,MUX} and gates.Batch* (gates/gates.go:26-312).
func (e *Engine) GateBatch(ops []Op, a, b, c []*tlwe.TLWELv0) []*tlwe.TLWELv0 {
	e.mu.Lock()
	defer e.mu.Unlock()
	count := len(a)
	if count == 0 {
		return nil
	}
	fa, out := e.flatten(a), make([]uint32, count*(e.n+1))
	var pb, pc *C.uint32_t
	var fb, fc []uint32
	if b != nil {
		fb = e.flatten(b)
		pb = (*C.uint32_t)(&fb[0])
	}
	if c != nil {
		fc = e.flatten(c)
		pc = (*C.uint32_t)(&fc[0])
	}
	e.check(C.tfhe_gate_batch(e.ctx, C.int64_t(count), (*C.uint8_t)(unsafe.Pointer(&ops[0])), C.int64_t(len(ops)),
		(*C.uint32_t)(&fa[0]), pb, pc, (*C.uint32_t)(&out[0])), "tfhe_gate_batch")
	runtime.KeepAlive(fb)
	runtime.KeepAlive(fc)
	return e.unflatten(out, count)
}

// BootstrapBatch replaces Evaluator.BootstrapAssign / BootstrapLUTAssign applied element-wise
// (evaluator/evaluator.go:139-148, evaluator/programmable_bootstrap.go:93-115).
// luts: nil => CloudKey.BlindRotateTestvec; else one TRLWE per ciphertext or a single one.
func (e *Engine) BootstrapBatch(cts []*tlwe.TLWELv0, luts []*trlwe.TRLWELv1) []*tlwe.TLWELv0 {
	e.mu.Lock()
	defer e.mu.Unlock()
	count := len(cts)
	if count == 0 {
		return nil
	}
	in, out := e.flatten(cts), make([]uint32, count*(e.n+1))
	var pl *C.uint32_t
	var fl []uint32
	if len(luts) > 0 {
		for _, l := range luts {
			fl = appendTorus(appendTorus(fl, l.A), l.B)
		}
		pl = (*C.uint32_t)(&fl[0])
	}
	e.check(C.tfhe_bootstrap_batch(e.ctx, C.int64_t(count), (*C.uint32_t)(&in[0]), pl, C.int64_t(len(luts)),
		(*C.uint32_t)(&out[0])), "tfhe_bootstrap_batch")
	runtime.KeepAlive(fl)
	return e.unflatten(out, count)
}

// BlindRotateBatch replaces trgsw.BatchBlindRotate (trgsw/trgsw.go:234-252).
func (e *Engine) BlindRotateBatch(cts []*tlwe.TLWELv0, luts []*trlwe.TRLWELv1) []*trlwe.TRLWELv1 {
	e.mu.Lock()
	defer e.mu.Unlock()
	count := len(cts)
	if count == 0 {
		return nil
	}
	in, out := e.flatten(cts), make([]uint32, count*2*e.bigN)
	var pl *C.uint32_t
	var fl []uint32
	if len(luts) > 0 {
		for _, l := range luts {
			fl = appendTorus(appendTorus(fl, l.A), l.B)
		}
		pl = (*C.uint32_t)(&fl[0])
	}
	e.check(C.tfhe_blind_rotate_batch(e.ctx, C.int64_t(count), (*C.uint32_t)(&in[0]), pl, C.int64_t(len(luts)),
		(*C.uint32_t)(&out[0])), "tfhe_blind_rotate_batch")
	runtime.KeepAlive(fl)
	res := make([]*trlwe.TRLWELv1, count)
	for i := range res {
		t := trlwe.NewTRLWELv1()
		for j := 0; j < e.bigN; j++ {
			t.A[j] = params.Torus(out[i*2*e.bigN+j])
			t.B[j] = params.Torus(out[i*2*e.bigN+e.bigN+j])
		}
		res[i] = t
	}
	return res
}

// ---- additive entry points (no counterpart in the reference; SURVEY 8(b) "additive", 8(f) ranks 1 and 4) --------------

// Gate is one gate of a circuit over wire ids (tfhe_gate_desc): wires 0..nInputs-1 are the inputs, every gate writes a
// distinct wire >= nInputs and reads only inputs or wires written by earlier gates.  In2 is read by MUX only.
type Gate struct {
	Op                 Op
	In0, In1, In2, Out int
}

// CircuitRun evaluates the same circuit on len(inputs[w]) independent instances (tfhe_circuit_run): one batch per
// level, intermediate wires resident on the GPUs, whole instances per device.  inputs[w][k] is input wire w of instance
// k; the result is indexed [output][instance].  Every gate computes exactly what gates.X computes.
func (e *Engine) CircuitRun(gatesList []Gate, inputs [][]*tlwe.TLWELv0, outputWires []int) [][]*tlwe.TLWELv0 {
	e.mu.Lock()
	defer e.mu.Unlock()
	if len(inputs) == 0 || len(outputWires) == 0 {
		return nil
	}
	instances := len(inputs[0])
	if instances == 0 {
		return make([][]*tlwe.TLWELv0, len(outputWires))
	}
	flat := make([]uint32, 0, len(inputs)*instances*(e.n+1))
	for _, wire := range inputs {
		if len(wire) != instances {
			panic("tfheb200: every input wire needs the same number of instances")
		}
		flat = append(flat, e.flatten(wire)...)
	}
	descs := make([]C.tfhe_gate_desc, len(gatesList))
	for i, g := range gatesList {
		descs[i].op = C.uint8_t(g.Op)
		descs[i].in0, descs[i].in1, descs[i].in2, descs[i].out = C.int32_t(g.In0), C.int32_t(g.In1), C.int32_t(g.In2), C.int32_t(g.Out)
	}
	ow := make([]C.int32_t, len(outputWires))
	for i, w := range outputWires {
		ow[i] = C.int32_t(w)
	}
	out := make([]uint32, len(outputWires)*instances*(e.n+1))
	var pd *C.tfhe_gate_desc
	if len(descs) > 0 {
		pd = &descs[0]
	}
	e.check(C.tfhe_circuit_run(e.ctx, C.int64_t(instances), C.int32_t(len(inputs)), C.int32_t(len(descs)), pd,
		(*C.uint32_t)(&flat[0]), C.int32_t(len(ow)), &ow[0], (*C.uint32_t)(&out[0])), "tfhe_circuit_run")
	runtime.KeepAlive(descs)
	res := make([][]*tlwe.TLWELv0, len(outputWires))
	per := instances * (e.n + 1)
	for k := range res {
		res[k] = e.unflatten(out[k*per:(k+1)*per], instances)
	}
	return res
}

// SetCircuitGraph turns the CUDA-graph replay of repeated circuits on or off (tfhe_ctx_set_circuit_graph).
func (e *Engine) SetCircuitGraph(on bool) {
	e.mu.Lock()
	defer e.mu.Unlock()
	v := C.int(0)
	if on {
		v = 1
	}
	e.check(C.tfhe_ctx_set_circuit_graph(e.ctx, v), "tfhe_ctx_set_circuit_graph")
}

// SetMuxMode: 0 = gates.MUX exactly (three bootstraps), 1 = two blind rotations + one key switch (tfhe_ctx_set_mux_mode).
func (e *Engine) SetMuxMode(mode int) {
	e.mu.Lock()
	defer e.mu.Unlock()
	e.check(C.tfhe_ctx_set_mux_mode(e.ctx, C.int(mode)), "tfhe_ctx_set_mux_mode")
}

func flattenLUTs(luts []*trlwe.TRLWELv1) []uint32 {
	var fl []uint32
	for _, l := range luts {
		fl = appendTorus(appendTorus(fl, l.A), l.B)
	}
	return fl
}

// BootstrapBatchIndexed is BootstrapLUTAssign for a batch that draws its LUTs from a small table: ciphertext k uses
// luts[index[k]] (tfhe_bootstrap_batch_indexed) — the table crosses the bus once instead of one 8-16 KiB LUT per ciphertext.
func (e *Engine) BootstrapBatchIndexed(cts []*tlwe.TLWELv0, luts []*trlwe.TRLWELv1, index []int32) []*tlwe.TLWELv0 {
	e.mu.Lock()
	defer e.mu.Unlock()
	count := len(cts)
	if count == 0 {
		return nil
	}
	if len(index) != count || len(luts) == 0 {
		panic("tfheb200: BootstrapBatchIndexed needs one index per ciphertext and a non-empty LUT table")
	}
	in, out, fl := e.flatten(cts), make([]uint32, count*(e.n+1)), flattenLUTs(luts)
	e.check(C.tfhe_bootstrap_batch_indexed(e.ctx, C.int64_t(count), (*C.uint32_t)(&in[0]), (*C.uint32_t)(&fl[0]),
		C.int64_t(len(luts)), (*C.int32_t)(&index[0]), (*C.uint32_t)(&out[0])), "tfhe_bootstrap_batch_indexed")
	return e.unflatten(out, count)
}

// BootstrapMultiLUT evaluates 2^log2K functions of every ciphertext with ONE blind rotation each
// (tfhe_bootstrap_multi_lut_batch); packed holds one packed test vector per ciphertext or a single one for all.  The
// result is indexed [ciphertext][function].
func (e *Engine) BootstrapMultiLUT(cts []*tlwe.TLWELv0, packed []*trlwe.TRLWELv1, log2K int) [][]*tlwe.TLWELv0 {
	e.mu.Lock()
	defer e.mu.Unlock()
	count := len(cts)
	if count == 0 {
		return nil
	}
	k := 1 << log2K
	in, out, fl := e.flatten(cts), make([]uint32, count*k*(e.n+1)), flattenLUTs(packed)
	e.check(C.tfhe_bootstrap_multi_lut_batch(e.ctx, C.int64_t(count), (*C.uint32_t)(&in[0]), (*C.uint32_t)(&fl[0]),
		C.int64_t(len(packed)), C.int32_t(log2K), (*C.uint32_t)(&out[0])), "tfhe_bootstrap_multi_lut_batch")
	res := make([][]*tlwe.TLWELv0, count)
	per := k * (e.n + 1)
	for i := range res {
		res[i] = e.unflatten(out[i*per:(i+1)*per], k)
	}
	return res
}

// LoadReencryptionKey uploads proxyreenc.ProxyReencryptionKey.KeyEncryptions (proxyreenc/proxyreenc.go:249-300: row
// base*t*i + base*j + k, each a TLWELv0 under the target key) with its Base = 2^basebit and T.
func (e *Engine) LoadReencryptionKey(rows []*tlwe.TLWELv0, basebit, t int) {
	e.mu.Lock()
	defer e.mu.Unlock()
	flat := e.flatten(rows)
	e.check(C.tfhe_ctx_load_reencryption_key(e.ctx, (*C.uint32_t)(&flat[0]), C.int32_t(basebit), C.int32_t(t)),
		"tfhe_ctx_load_reencryption_key")
}

// ReencryptBatch replaces proxyreenc.ReencryptTLWELv0 (proxyreenc/proxyreenc.go:321-366) applied element-wise.
func (e *Engine) ReencryptBatch(cts []*tlwe.TLWELv0) []*tlwe.TLWELv0 {
	e.mu.Lock()
	defer e.mu.Unlock()
	count := len(cts)
	if count == 0 {
		return nil
	}
	in, out := e.flatten(cts), make([]uint32, count*(e.n+1))
	e.check(C.tfhe_reencrypt_batch(e.ctx, C.int64_t(count), (*C.uint32_t)(&in[0]), (*C.uint32_t)(&out[0])), "tfhe_reencrypt_batch")
	return e.unflatten(out, count)
}
