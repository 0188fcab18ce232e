// wire.go — the TFHB flat file format (include/tfhe_b200_client.h, go-tfhe_b200/wire.py) on the Go side.
//
// The reference has no serialisation; its "format" is the Go structs key.SecretKey (key/key.go:10-13) and
// cloudkey.CloudKey (cloudkey/cloudkey.go:16-21).  This file flattens them exactly as tfheb200.New flattens them for the
// C ABI and frames the result, so that keys made by a Go process can be shipped to a GPU process, and golden vectors made
// by the unmodified reference (go/cmd/mkgolden) can be checked on a machine without a Go toolchain.
//
//	"TFHB" | version u32 = 2 | kind u32 | params 6 x i32 (n, N, L, bgbit, basebit, iks_t) | nsect u32 |
//	sections: tag [4]byte (space padded) | dtype u32 (0 = u32, 1 = f64) | count u64 | little-endian values |
//	CRC-32 (IEEE) of everything before it, as u64
//
// This file has no cgo dependency and builds with or without the b200 tag.
package tfheb200

import (
	"bytes"
	"encoding/binary"
	"errors"
	"hash/crc32"
	"io"
	"math"
	"os"

	"github.com/thedonutfactory/go-tfhe/cloudkey"
	"github.com/thedonutfactory/go-tfhe/key"
	"github.com/thedonutfactory/go-tfhe/params"
	"github.com/thedonutfactory/go-tfhe/tlwe"
	"github.com/thedonutfactory/go-tfhe/trlwe"
)

const (
	WireVersion = 2
	KindSecret  = 1
	KindCloud   = 2
	KindCT      = 3
	KindTRLWE   = 4
	KindBundle  = 5
)

// Section is one named array of a TFHB file: exactly one of U32 / F64 is set.
type Section struct {
	Tag string
	U32 []uint32
	F64 []float64
}

func wireParams() [6]int32 {
	g, l0 := params.GetTRGSWLv1(), params.GetTLWELv0()
	return [6]int32{int32(l0.N), int32(g.N), int32(g.L), int32(g.BGBIT), int32(g.BASEBIT), int32(g.IKS_T)}
}

// Pack serialises sections under the CURRENT parameter set (params.CurrentSecurityLevel).
func Pack(kind uint32, sections []Section) []byte {
	var b bytes.Buffer
	b.WriteString("TFHB")
	binary.Write(&b, binary.LittleEndian, uint32(WireVersion))
	binary.Write(&b, binary.LittleEndian, kind)
	binary.Write(&b, binary.LittleEndian, wireParams())
	binary.Write(&b, binary.LittleEndian, uint32(len(sections)))
	for _, s := range sections {
		tag := []byte(s.Tag + "    ")[:4]
		b.Write(tag)
		if s.F64 != nil {
			binary.Write(&b, binary.LittleEndian, uint32(1))
			binary.Write(&b, binary.LittleEndian, uint64(len(s.F64)))
			binary.Write(&b, binary.LittleEndian, s.F64)
		} else {
			binary.Write(&b, binary.LittleEndian, uint32(0))
			binary.Write(&b, binary.LittleEndian, uint64(len(s.U32)))
			binary.Write(&b, binary.LittleEndian, s.U32)
		}
	}
	crc := uint64(crc32.ChecksumIEEE(b.Bytes()))
	binary.Write(&b, binary.LittleEndian, crc)
	return b.Bytes()
}

// Unpack parses a TFHB blob and verifies its checksum.
func Unpack(blob []byte) (kind uint32, p [6]int32, sections map[string]Section, err error) {
	if len(blob) < 48 || string(blob[:4]) != "TFHB" {
		return 0, p, nil, errors.New("tfheb200: not a TFHB file")
	}
	if binary.LittleEndian.Uint32(blob[4:]) != WireVersion {
		return 0, p, nil, errors.New("tfheb200: unsupported TFHB version")
	}
	body := blob[:len(blob)-8]
	if uint64(crc32.ChecksumIEEE(body)) != binary.LittleEndian.Uint64(blob[len(blob)-8:]) {
		return 0, p, nil, errors.New("tfheb200: TFHB checksum mismatch")
	}
	kind = binary.LittleEndian.Uint32(blob[8:])
	for i := range p {
		p[i] = int32(binary.LittleEndian.Uint32(blob[12+4*i:]))
	}
	n := int(binary.LittleEndian.Uint32(blob[36:]))
	off := 40
	sections = map[string]Section{}
	for i := 0; i < n; i++ {
		if off+16 > len(body) {
			return 0, p, nil, io.ErrUnexpectedEOF
		}
		tag := string(bytes.TrimRight(blob[off:off+4], " "))
		dt := binary.LittleEndian.Uint32(blob[off+4:])
		cnt := int(binary.LittleEndian.Uint64(blob[off+8:]))
		off += 16
		s := Section{Tag: tag}
		if dt == 1 {
			if off+8*cnt > len(body) {
				return 0, p, nil, io.ErrUnexpectedEOF
			}
			s.F64 = make([]float64, cnt)
			for j := range s.F64 {
				s.F64[j] = math.Float64frombits(binary.LittleEndian.Uint64(blob[off+8*j:]))
			}
			off += 8 * cnt
		} else {
			if off+4*cnt > len(body) {
				return 0, p, nil, io.ErrUnexpectedEOF
			}
			s.U32 = make([]uint32, cnt)
			for j := range s.U32 {
				s.U32[j] = binary.LittleEndian.Uint32(blob[off+4*j:])
			}
			off += 4 * cnt
		}
		sections[tag] = s
	}
	return kind, p, sections, nil
}

func torusToU32(src []params.Torus) []uint32 {
	out := make([]uint32, len(src))
	for i, v := range src {
		out[i] = uint32(v)
	}
	return out
}

// FlattenLWE: []*tlwe.TLWELv0 -> [count][n+1] u32 (tlwe/tlwe.go:11-13).
func FlattenLWE(cts []*tlwe.TLWELv0) []uint32 {
	var out []uint32
	for _, c := range cts {
		out = append(out, torusToU32(c.P)...)
	}
	return out
}

// FlattenTRLWE: []*trlwe.TRLWELv1 -> [count][2][N] u32, A then B (trlwe/trlwe.go:13-16).
func FlattenTRLWE(ts []*trlwe.TRLWELv1) []uint32 {
	var out []uint32
	for _, t := range ts {
		out = append(out, torusToU32(t.A)...)
		out = append(out, torusToU32(t.B)...)
	}
	return out
}

// PackSecretKey: key.SecretKey (key/key.go:10-13) -> kind 1, sections lv0, lv1.
func PackSecretKey(sk *key.SecretKey) []byte {
	return Pack(KindSecret, []Section{{Tag: "lv0", U32: torusToU32(sk.KeyLv0)}, {Tag: "lv1", U32: torusToU32(sk.KeyLv1)}})
}

// PackCloudKey: cloudkey.CloudKey (cloudkey/cloudkey.go:16-21) -> kind 2, sections offs, tvec, bsk, ksk — the very
// buffers tfheb200.New hands to tfhe_ctx_load_cloudkey: bsk [n][2L][2][N] float64 in the reference's FourierPoly layout,
// ksk [N*t*base][n+1] in the reference's row order.
func PackCloudKey(ck *cloudkey.CloudKey) []byte {
	var bsk []float64
	for _, row := range ck.BootstrappingKey {
		for _, t := range row.TRLWEFFT {
			bsk = append(bsk, t.A.Coeffs...)
			bsk = append(bsk, t.B.Coeffs...)
		}
	}
	secs := []Section{
		{Tag: "offs", U32: []uint32{uint32(ck.DecompositionOffset)}},
		{Tag: "tvec", U32: FlattenTRLWE([]*trlwe.TRLWELv1{ck.BlindRotateTestvec})},
		{Tag: "bsk", F64: bsk},
	}
	if len(ck.KeySwitchingKey) > 0 {
		secs = append(secs, Section{Tag: "ksk", U32: FlattenLWE(ck.KeySwitchingKey)})
	}
	return Pack(KindCloud, secs)
}

// WriteFile writes a packed blob.
func WriteFile(path string, blob []byte) error { return os.WriteFile(path, blob, 0o644) }
