"""Versioned flat binary format for keys and ciphertext batches.

The reference has no serialization at all (no encoding/*, no file I/O); its "format" is the in-memory Go structs
(key/key.go:10-13, cloudkey/cloudkey.go:16-21).  This format is the flattened layout the C ABI already takes
(include/tfhe_b200.h), framed so that a key made by a Go process can be shipped to a GPU process (and golden vectors
produced by a real Go run elsewhere can be brought back):

    magic  "TFHB"            4 bytes
    version u32              currently 1
    kind    u32              1 = SecretKey, 2 = CloudKey, 3 = ciphertext batch, 4 = TRLWE/LUT batch
    params  6 x i32          n, N, L, bgbit, basebit, iks_t      (tfhe_params)
    nsect   u32              number of sections
    then per section: tag (4 ascii bytes), dtype (u32: 0 = u32, 1 = f64), count (u64), raw little-endian data,
    and finally a u64 FNV-1a checksum of everything before it.
"""
import struct

import numpy as np

MAGIC = b"TFHB"
VERSION = 1
KIND_SECRET, KIND_CLOUD, KIND_CT, KIND_TRLWE = 1, 2, 3, 4
_DT = {0: np.dtype("<u4"), 1: np.dtype("<f8")}


def _fnv1a(data):
    h = 0xCBF29CE484222325
    # 64-bit FNV-1a over 8-byte words (tail padded with zeros): cheap and order-sensitive
    pad = (-len(data)) % 8
    words = np.frombuffer(data + b"\0" * pad, dtype="<u8")
    for chunk in np.array_split(words, max(1, len(words) // (1 << 20))):
        for w in (int(chunk.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(chunk)) if len(chunk) else 0, len(chunk)):
            h ^= w & 0xFFFFFFFFFFFFFFFF
            h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _pack(kind, P, sections):
    out = [MAGIC, struct.pack("<II", VERSION, kind), struct.pack("<6i", P.n, P.N, P.L, P.BGBIT, P.BASEBIT, P.IKS_T),
           struct.pack("<I", len(sections))]
    for tag, arr in sections:
        arr = np.ascontiguousarray(arr)
        dt = 1 if arr.dtype == np.float64 else 0
        arr = arr.astype(_DT[dt], copy=False)
        out += [tag.encode("ascii").ljust(4)[:4], struct.pack("<IQ", dt, arr.size), arr.tobytes()]
    body = b"".join(out)
    return body + struct.pack("<Q", _fnv1a(body))


def _unpack(blob, want_kind):
    if blob[:4] != MAGIC:
        raise ValueError("not a TFHB file")
    body, (chk,) = blob[:-8], struct.unpack("<Q", blob[-8:])
    if _fnv1a(body) != chk:
        raise ValueError("checksum mismatch")
    version, kind = struct.unpack_from("<II", blob, 4)
    if version != VERSION:
        raise ValueError("unsupported version %d" % version)
    if kind != want_kind:
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
    if tuple(pvals) != (P.n, P.N, P.L, P.BGBIT, P.BASEBIT, P.IKS_T):
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
