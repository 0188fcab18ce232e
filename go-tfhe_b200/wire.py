"""Versioned flat binary format for keys and ciphertext batches.

The reference has no serialization at all (no encoding/*, no file I/O); its "format" is the in-memory Go structs
(key/key.go:10-13, cloudkey/cloudkey.go:16-21).  This format is the flattened layout the C ABI already takes
(include/tfhe_b200.h), framed so that a key made by a Go process can be shipped to a GPU process and golden vectors
produced by a real Go run (go/cmd/mkgolden) can be brought back.  Three writers/readers hold the same format: this module,
libtfhe_b200_client (tfhe_wire_pack / tfhe_wire_unpack, include/tfhe_b200_client.h) and go/tfheb200/wire.go.

    magic  "TFHB"            4 bytes
    version u32              currently 2
    kind    u32              1 = SecretKey, 2 = CloudKey, 3 = ciphertext batch, 4 = TRLWE/LUT batch, 5 = named vector bundle
    params  6 x i32          n, N, L, bgbit, basebit, iks_t      (tfhe_params)
    nsect   u32              number of sections
    then per section: tag (4 ascii bytes, space padded), dtype (u32: 0 = u32, 1 = f64), count (u64), raw little-endian
    data; and finally the CRC-32 (IEEE, zlib.crc32 / Go hash/crc32) of everything before it, as a u64.
"""
import ctypes
import struct
import zlib

import numpy as np

MAGIC = b"TFHB"
VERSION = 2
KIND_SECRET, KIND_CLOUD, KIND_CT, KIND_TRLWE, KIND_BUNDLE = 1, 2, 3, 4, 5
_DT = {0: np.dtype("<u4"), 1: np.dtype("<f8")}


def _pvals(P):
    return (P.n, P.N, P.L, P.BGBIT, P.BASEBIT, P.IKS_T)


def _pack(kind, P, sections):
    out = [MAGIC, struct.pack("<II", VERSION, kind), struct.pack("<6i", *_pvals(P)), struct.pack("<I", len(sections))]
    for tag, arr in sections:
        arr = np.ascontiguousarray(arr)
        dt = 1 if arr.dtype == np.float64 else 0
        arr = arr.astype(_DT[dt], copy=False)
        out += [tag.encode("ascii").ljust(4)[:4], struct.pack("<IQ", dt, arr.size), arr.tobytes()]
    body = b"".join(out)
    return body + struct.pack("<Q", zlib.crc32(body) & 0xFFFFFFFF)


def _unpack(blob, want_kind):
    if blob[:4] != MAGIC:
        raise ValueError("not a TFHB file")
    body, (chk,) = blob[:-8], struct.unpack("<Q", blob[-8:])
    version, kind = struct.unpack_from("<II", blob, 4)
    if version != VERSION:
        raise ValueError("unsupported version %d" % version)
    if (zlib.crc32(body) & 0xFFFFFFFF) != chk:
        raise ValueError("checksum mismatch")
    if want_kind is not None and kind != want_kind:
        raise ValueError("wrong kind %d (wanted %d)" % (kind, want_kind))
    pvals = struct.unpack_from("<6i", blob, 12)
    (nsect,) = struct.unpack_from("<I", blob, 36)
    off, sections = 40, {}
    for _ in range(nsect):
        tag = blob[off:off + 4].decode("ascii").strip()
        dt, count = struct.unpack_from("<IQ", blob, off + 4)
        off += 16
        nbytes = count * _DT[dt].itemsize
        sections[tag] = np.frombuffer(blob, dtype=_DT[dt], count=count, offset=off).copy()
        off += nbytes
    return pvals, sections


def _match_params(pvals, P):
    if tuple(pvals) != _pvals(P):
        raise ValueError("parameter set mismatch: file has %r" % (pvals,))


def dumps_secret_key(sk):
    return _pack(KIND_SECRET, sk.P, [("lv0", sk.KeyLv0), ("lv1", sk.KeyLv1)])


def loads_secret_key(blob, P):
    from .key import SecretKey
    pvals, s = _unpack(blob, KIND_SECRET)
    _match_params(pvals, P)
    return SecretKey(P, s["lv0"].astype(np.uint32), s["lv1"].astype(np.uint32))


def dumps_cloud_key(ck):
    sections = [("offs", np.array([ck.DecompositionOffset], dtype=np.uint32)), ("tvec", ck.BlindRotateTestvec),
                ("bsk", ck.BootstrappingKey)]
    if ck.KeySwitchingKey is not None:
        sections.append(("ksk", ck.KeySwitchingKey))
    return _pack(KIND_CLOUD, ck.P, sections)


def loads_cloud_key(blob, P):
    from .cloudkey import CloudKey
    pvals, s = _unpack(blob, KIND_CLOUD)
    _match_params(pvals, P)
    ksk = s["ksk"].astype(np.uint32).reshape(P.ksk_rows, P.n + 1) if "ksk" in s else None
    return CloudKey(P, int(s["offs"][0]), s["tvec"].astype(np.uint32).reshape(2, P.N), ksk,
                    s["bsk"].reshape(P.n, 2 * P.L, 2, P.N))


def dumps_ciphertexts(P, ct):
    return _pack(KIND_CT, P, [("ct", np.asarray(ct, dtype=np.uint32).reshape(-1, P.n + 1))])


def loads_ciphertexts(blob, P):
    pvals, s = _unpack(blob, KIND_CT)
    _match_params(pvals, P)
    return s["ct"].astype(np.uint32).reshape(-1, P.n + 1)


def dumps_bundle(P, named):
    """A named vector bundle (kind 5): {tag (<= 4 ascii chars): u32 or f64 array}.  Golden-vector files use this."""
    return _pack(KIND_BUNDLE, P, list(named.items()))


def loads_bundle(blob, P=None):
    pvals, s = _unpack(blob, KIND_BUNDLE)
    if P is not None:
        _match_params(pvals, P)
    return pvals, s


def loads_any(blob):
    """(kind, params 6-tuple, {tag: array}) of any TFHB blob — parameter set taken from the file."""
    pvals, s = _unpack(blob, None)
    return struct.unpack_from("<I", blob, 8)[0], pvals, s


# ---- the same through the C library (tfhe_wire_pack / tfhe_wire_unpack): what a C or cgo caller uses -----------------
def c_pack(kind, P, sections):
    from . import _native
    lib = _native.client()
    arrs = []
    secs = (_native.WireSection * max(len(sections), 1))()
    for i, (tag, arr) in enumerate(sections):
        arr = np.ascontiguousarray(arr)
        dt = 1 if arr.dtype == np.float64 else 0
        arr = np.ascontiguousarray(arr.astype(_DT[dt], copy=False))
        arrs.append(arr)
        secs[i] = _native.WireSection(tag.encode("ascii").ljust(4)[:4], dt, arr.size, arr.ctypes.data)
    tp = _native.TfheParams(*_pvals(P))
    need = lib.tfhe_wire_pack(kind, ctypes.byref(tp), secs, len(sections), None, 0)
    if need < 0:
        raise ValueError("tfhe_wire_pack: bad arguments")
    buf = ctypes.create_string_buffer(need)
    got = lib.tfhe_wire_pack(kind, ctypes.byref(tp), secs, len(sections), buf, need)
    if got != need:
        raise ValueError("tfhe_wire_pack failed")
    return buf.raw


def c_unpack(blob):
    from . import _native
    lib = _native.client()
    kind, ns = ctypes.c_uint32(0), ctypes.c_uint32(64)
    tp = _native.TfheParams()
    secs = (_native.WireSection * 64)()
    buf = ctypes.create_string_buffer(blob, len(blob))
    rc = lib.tfhe_wire_unpack(buf, len(blob), ctypes.byref(kind), ctypes.byref(tp), secs, ctypes.byref(ns))
    if rc != 0:
        raise ValueError({-1: "malformed TFHB blob", -2: "checksum mismatch", -3: "unsupported version",
                          -4: "too many sections"}.get(rc, "tfhe_wire_unpack failed (%d)" % rc))
    out = {}
    for i in range(ns.value):
        s = secs[i]
        dt = _DT[s.dtype]
        out[s.tag.decode("ascii").strip()] = np.frombuffer(
            ctypes.string_at(s.data, s.count * dt.itemsize), dtype=dt).copy()
    return kind.value, (tp.n, tp.N, tp.L, tp.bgbit, tp.basebit, tp.iks_t), out
