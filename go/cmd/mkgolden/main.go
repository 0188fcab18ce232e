// mkgolden — golden vectors from the UNMODIFIED reference, for machines that cannot run Go.
//
// The B200 engine is developed where no Go toolchain exists, so its CPU oracle (oracle/oracle.cpp, a C++ restatement of
// the reference's Go path) is pinned against the reference only at plaintext level.  This program closes the gap: run it
// once anywhere Go >= 1.21 is installed, inside a checkout of github.com/thedonutfactory/go-tfhe with this repository's
// go/tfheb200/wire.go copied to ./tfheb200/ (wire.go has no cgo dependency):
//
//	cp -r <this repo>/go/tfheb200/wire.go   <go-tfhe>/tfheb200/wire.go
//	cp -r <this repo>/go/cmd/mkgolden       <go-tfhe>/cmd/mkgolden
//	cd <go-tfhe> && go run ./cmd/mkgolden -out <this repo>/tests/golden/go [-sets 80,128,uint2,uint5] [-seed 1]
//
// For every parameter set it writes a directory with the secret key, the cloud key (both in the layouts of
// include/tfhe_b200.h) and a vector bundle holding inputs and the reference's own outputs for
//
//	gates.{NAND,AND,OR,XOR,XNOR,NOR,ANDNY,ANDYN,ORNY,ORYN,MUX,NOT,Copy}   gates/gates.go:26-130      (N = 1024 sets)
//	Evaluator.ExternalProductAssign / CMuxAssign                         evaluator/evaluator.go:50-106
//	Evaluator.BlindRotateAssign / BootstrapAssign                         evaluator/evaluator.go:110-148
//	Evaluator.BootstrapLUTAssign (identity, complement, x -> x mod m/2)   evaluator/programmable_bootstrap.go:93-115
//	poly.Evaluator.ToFourierPoly                                          poly/fourier_transform.go:18-21
//
// tests/test_golden_go.py then requires oracle == Go and GPU == Go on them (bit-exact on the 80/110/128-bit sets, stated
// tolerance on the Uint sets).  The reference's RNG cannot be seeded per call (key/key.go:17 and tlwe/tlwe.go:37 draw
// rand.Int63() from the global source, and key generation does so from concurrent goroutines), so the files record what
// this run produced; rand.Seed only makes single-goroutine parts repeatable.
package main

import (
	"flag"
	"fmt"
	"math/rand"
	"os"
	"path/filepath"
	"strings"

	"github.com/thedonutfactory/go-tfhe/cloudkey"
	"github.com/thedonutfactory/go-tfhe/evaluator"
	"github.com/thedonutfactory/go-tfhe/gates"
	"github.com/thedonutfactory/go-tfhe/key"
	"github.com/thedonutfactory/go-tfhe/lut"
	"github.com/thedonutfactory/go-tfhe/params"
	"github.com/thedonutfactory/go-tfhe/poly"
	"github.com/thedonutfactory/go-tfhe/tfheb200"
	"github.com/thedonutfactory/go-tfhe/tlwe"
	"github.com/thedonutfactory/go-tfhe/trlwe"
)

type set struct {
	name   string
	level  params.SecurityLevel
	msgMod int // 0: Boolean set (gates), else message modulus of the programmable-bootstrap vectors
}

var allSets = []set{
	{"80", params.Security80Bit, 0}, {"110", params.Security110Bit, 0}, {"128", params.Security128Bit, 0},
	{"uint1", params.SecurityUint1, 2}, {"uint2", params.SecurityUint2, 4}, {"uint3", params.SecurityUint3, 8},
	{"uint4", params.SecurityUint4, 16}, {"uint5", params.SecurityUint5, 32},
}

func must(err error) {
	if err != nil {
		fmt.Fprintln(os.Stderr, "mkgolden:", err)
		os.Exit(1)
	}
}

func u32s(t []params.Torus) []uint32 {
	out := make([]uint32, len(t))
	for i, v := range t {
		out[i] = uint32(v)
	}
	return out
}

func main() {
	out := flag.String("out", "tests/golden/go", "output directory (one sub-directory per parameter set)")
	sets := flag.String("sets", "80,128,uint2,uint5", "comma-separated parameter sets")
	seed := flag.Int64("seed", 1, "rand.Seed value")
	count := flag.Int("count", 8, "ciphertexts per vector")
	flag.Parse()
	rand.Seed(*seed)
	want := map[string]bool{}
	for _, s := range strings.Split(*sets, ",") {
		want[strings.TrimSpace(s)] = true
	}
	for _, s := range allSets {
		if !want[s.name] {
			continue
		}
		params.CurrentSecurityLevel = s.level // every Get*() below reads it (params/params.go:517-626)
		dir := filepath.Join(*out, s.name)
		must(os.MkdirAll(dir, 0o755))
		emit(dir, s, *count)
		fmt.Println("wrote", dir)
	}
}

func emit(dir string, s set, count int) {
	sk := key.NewSecretKey()
	ck := cloudkey.NewCloudKey(sk)
	must(tfheb200.WriteFile(filepath.Join(dir, "secret.tfhb"), tfheb200.PackSecretKey(sk)))
	must(tfheb200.WriteFile(filepath.Join(dir, "cloud.tfhb"), tfheb200.PackCloudKey(ck)))

	g, l0 := params.GetTRGSWLv1(), params.GetTLWELv0()
	ev := evaluator.NewEvaluator(g.N)
	var secs []tfheb200.Section
	add := func(tag string, v []uint32) { secs = append(secs, tfheb200.Section{Tag: tag, U32: v}) }

	// --- polynomial transform: ToFourierPoly of a fixed polynomial (poly/poly_test.go:10-33 uses p[i] = i * 12345) --------
	pe := poly.NewEvaluator(g.N)
	p := poly.NewPoly(g.N)
	for i := range p.Coeffs {
		p.Coeffs[i] = params.Torus(uint32(i) * 12345)
	}
	fp := pe.ToFourierPoly(p)
	add("polp", u32s(p.Coeffs))
	secs = append(secs, tfheb200.Section{Tag: "polf", F64: append([]float64(nil), fp.Coeffs...)})

	// --- external product and CMUX with bootstrapping-key rows 0 and n-1 on random TRLWEs -------------------------------
	rng := rand.New(rand.NewSource(rand.Int63()))
	var c0s, c1s, eps, cms []*trlwe.TRLWELv1
	rows := []uint32{0, uint32(l0.N - 1)}
	for _, r := range rows {
		for k := 0; k < 2; k++ {
			c0, c1 := trlwe.NewTRLWELv1(), trlwe.NewTRLWELv1()
			for i := 0; i < g.N; i++ {
				c0.A[i], c0.B[i] = params.Torus(rng.Uint32()), params.Torus(rng.Uint32())
				c1.A[i], c1.B[i] = params.Torus(rng.Uint32()), params.Torus(rng.Uint32())
			}
			ep, cm := trlwe.NewTRLWELv1(), trlwe.NewTRLWELv1()
			ev.ExternalProductAssign(ck.BootstrappingKey[r], c1, ck.DecompositionOffset, ep)
			ev.CMuxAssign(ck.BootstrappingKey[r], c0, c1, ck.DecompositionOffset, cm)
			c0s, c1s, eps, cms = append(c0s, c0), append(c1s, c1), append(eps, ep), append(cms, cm)
		}
	}
	add("erow", rows) // two vectors per listed row, in order
	add("ec0", tfheb200.FlattenTRLWE(c0s))
	add("ec1", tfheb200.FlattenTRLWE(c1s))
	add("eout", tfheb200.FlattenTRLWE(eps))
	add("cout", tfheb200.FlattenTRLWE(cms))

	if s.msgMod == 0 {
		emitBoolean(ev, sk, ck, count, add)
	} else {
		emitMessages(ev, sk, ck, s.msgMod, count, add)
	}
	must(tfheb200.WriteFile(filepath.Join(dir, "vectors.tfhb"), tfheb200.Pack(tfheb200.KindBundle, secs)))
}

// Boolean sets: fresh encryptions of all input combinations, every gate, blind rotation and bootstrap.
func emitBoolean(ev *evaluator.Evaluator, sk *key.SecretKey, ck *cloudkey.CloudKey, count int, add func(string, []uint32)) {
	alpha := params.GetTLWELv0().ALPHA
	var a, b, c []*tlwe.TLWELv0
	var bits []uint32
	for i := 0; i < count; i++ {
		x, y, z := i&1 == 1, i&2 == 2, i&4 == 4
		a = append(a, tlwe.NewTLWELv0().EncryptBool(x, alpha, sk.KeyLv0))
		b = append(b, tlwe.NewTLWELv0().EncryptBool(y, alpha, sk.KeyLv0))
		c = append(c, tlwe.NewTLWELv0().EncryptBool(z, alpha, sk.KeyLv0))
		bits = append(bits, uint32(i&7))
	}
	add("bits", bits) // bit 0 = a, bit 1 = b, bit 2 = c
	add("ina", tfheb200.FlattenLWE(a))
	add("inb", tfheb200.FlattenLWE(b))
	add("inc", tfheb200.FlattenLWE(c))
	type gate2 func(x, y *gates.Ciphertext, ck *cloudkey.CloudKey) *gates.Ciphertext
	// tags are the opcode names of include/tfhe_b200.h, cut to four characters
	for _, gt := range []struct {
		tag string
		f   gate2
	}{{"NAND", gates.NAND}, {"AND", gates.AND}, {"OR", gates.OR}, {"XOR", gates.XOR}, {"XNOR", gates.XNOR}, {"NOR", gates.NOR},
		{"ANNY", gates.ANDNY}, {"ANYN", gates.ANDYN}, {"ORNY", gates.ORNY}, {"ORYN", gates.ORYN}} {
		var outs []*tlwe.TLWELv0
		for i := range a {
			outs = append(outs, gt.f(a[i], b[i], ck))
		}
		add(gt.tag, tfheb200.FlattenLWE(outs))
	}
	var mux, not, cpy []*tlwe.TLWELv0
	for i := range a {
		mux = append(mux, gates.MUX(a[i], b[i], c[i], ck))
		not = append(not, gates.NOT(a[i]))
		cpy = append(cpy, gates.Copy(a[i]))
	}
	add("MUX", tfheb200.FlattenLWE(mux))
	add("NOT", tfheb200.FlattenLWE(not))
	add("COPY", tfheb200.FlattenLWE(cpy))
	// blind rotation and bootstrap of the raw inputs with the default test vector
	var rot []*trlwe.TRLWELv1
	var boot []*tlwe.TLWELv0
	for i := range a {
		r := trlwe.NewTRLWELv1()
		ev.BlindRotateAssign(a[i], ck.BlindRotateTestvec, ck.BootstrappingKey, ck.DecompositionOffset, r)
		rot = append(rot, r)
		o := tlwe.NewTLWELv0()
		ev.BootstrapAssign(a[i], ck.BlindRotateTestvec, ck.BootstrappingKey, ck.KeySwitchingKey, ck.DecompositionOffset, o)
		boot = append(boot, o)
	}
	add("rot", tfheb200.FlattenTRLWE(rot))
	add("boot", tfheb200.FlattenLWE(boot))
}

// Message sets: programmable bootstraps of every message with identity, complement and x mod m/2.
func emitMessages(ev *evaluator.Evaluator, sk *key.SecretKey, ck *cloudkey.CloudKey, m, count int, add func(string, []uint32)) {
	alpha := params.GetTLWELv0().ALPHA
	gen := lut.NewGenerator(m)
	fs := []func(int) int{
		func(x int) int { return x },
		func(x int) int { return (m - 1) - x },
		func(x int) int {
			if m > 2 {
				return x % (m / 2)
			}
			return x
		},
	}
	var cts []*tlwe.TLWELv0
	var msgs []uint32
	for i := 0; i < count; i++ {
		v := (i * (m - 1) / (count - 1 + 1)) % m
		if i == count-1 {
			v = m - 1
		}
		cts = append(cts, tlwe.NewTLWELv0().EncryptLWEMessage(v, m, alpha, sk.KeyLv0))
		msgs = append(msgs, uint32(v))
	}
	add("msgs", msgs)
	add("mmod", []uint32{uint32(m)})
	add("ct", tfheb200.FlattenLWE(cts))
	var luts []*trlwe.TRLWELv1
	for k, f := range fs {
		l := gen.GenLookUpTable(f)
		luts = append(luts, l.Poly)
		var rot []*trlwe.TRLWELv1
		var outs []*tlwe.TLWELv0
		for i := range cts {
			r := trlwe.NewTRLWELv1()
			ev.BlindRotateAssign(cts[i], l.Poly, ck.BootstrappingKey, ck.DecompositionOffset, r)
			rot = append(rot, r)
			o := tlwe.NewTLWELv0()
			ev.BootstrapLUTAssign(cts[i], l, ck.BootstrappingKey, ck.KeySwitchingKey, ck.DecompositionOffset, o)
			outs = append(outs, o)
		}
		add(fmt.Sprintf("rot%d", k), tfheb200.FlattenTRLWE(rot))
		add(fmt.Sprintf("pbs%d", k), tfheb200.FlattenLWE(outs))
	}
	add("luts", tfheb200.FlattenTRLWE(luts))
}
