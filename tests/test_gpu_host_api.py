"""GPU tests of the host side of the C ABI added in round 2: the multi-device context (tfhe_ctx_create_multi), the
pipelined host-buffer calls, the asynchronous device-buffer gate batch (in-place use, host opcodes).  Every check is
"same words as the plain single-device, single-chunk call", which tests/test_gpu_parity.py pins to the oracle."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OPS = ["AND", "OR", "XOR", "MUX", "NOT", "COPY", "NAND", "XNOR"]


@pytest.fixture(scope="module")
def T():
    return importlib.import_module("go-tfhe_b200")


@pytest.fixture(scope="module")
def env(T, keyset):
    P, sk, ck = keyset("80")
    ctx = T.Context(T.params.get("80"), 0)
    ctx.load_cloudkey(ck.offset, ck.bsk_fft, ck.ksk, ck.testvec)
    yield P, sk, ck, ctx
    ctx.close()


def _mixed(sk, count, seed):
    rng = np.random.default_rng(seed)
    A, B, C = (rng.integers(0, 2, count).astype(np.uint8) for _ in range(3))
    ops = [OPS[k] for k in rng.integers(0, len(OPS), count)]
    a, b, c = sk.encrypt_bool(A, seed + 1), sk.encrypt_bool(B, seed + 2), sk.encrypt_bool(C, seed + 3)
    want = []
    for o, x, y, z in zip(ops, A, B, C):
        want.append({"AND": x & y, "OR": x | y, "XOR": x ^ y, "MUX": y if x else z, "NOT": 1 - x, "COPY": x,
                     "NAND": 1 - (x & y), "XNOR": 1 - (x ^ y)}[o])
    return ops, a, b, c, np.array(want, dtype=np.uint8)


def test_pipelined_host_calls_are_chunk_invariant(T, O, env):
    """tfhe_gate_batch / tfhe_bootstrap_batch / tfhe_blind_rotate_batch cut large host batches into chunks that flow
    through two staging slots on three streams.  Chunk sizes that give 1, 2, 3 and 6 chunks (incl. a ragged last one)
    must all give the words of the unchunked call; mixed opcodes exercise the per-chunk index lists."""
    P, sk, ck, ctx = env
    count = 41
    ops, a, b, c, want = _mixed(sk, count, 100)
    msgs = np.arange(count) % 2
    ct = sk.encrypt_message(msgs, 2, 7)
    luts = np.stack([O.gen_lut(P, 2, (lambda x: x) if k % 2 else (lambda x: 1 - x)) for k in range(count)])
    try:
        ctx.set_pipeline_chunk(1 << 20)
        ref_g = ctx.gate_batch(ops, a, b, c)
        ref_b = ctx.bootstrap_batch(ct, luts)
        ref_r = ctx.blind_rotate_batch(ct, luts[0])
        assert np.array_equal(sk.decrypt_bool(ref_g), want)
        for rows in (27, 20, 14, 7):
            ctx.set_pipeline_chunk(rows)
            assert np.array_equal(ctx.gate_batch(ops, a, b, c), ref_g), rows
            assert np.array_equal(ctx.bootstrap_batch(ct, luts), ref_b), rows
            assert np.array_equal(ctx.blind_rotate_batch(ct, luts[0]), ref_r), rows
            assert np.array_equal(ctx.gate_batch("NAND", a, b), ctx.gate_batch(["NAND"] * count, a, b)), rows
    finally:
        ctx.set_pipeline_chunk(16384)


def test_multi_device_context_matches_single_device(T, O, env):
    """One context over every visible GPU (tfhe_ctx_create_multi): key uploaded once and replicated by peer copies, host
    batches sharded by index (MUX-weighted), circuits by instance.  Words must equal the single-device context's —
    with one visible GPU this still runs the whole group path (sharding, threads, key replication code with no peer)."""
    P, sk, ck, ctx = env
    multi = T.Context(T.params.get("80"), devices="all")
    try:
        assert multi.device_count >= 1
        multi.load_cloudkey(ck.offset, ck.bsk_fft, ck.ksk, ck.testvec)
        count = 37
        ops, a, b, c, want = _mixed(sk, count, 200)
        got = multi.gate_batch(ops, a, b, c)
        assert np.array_equal(got, ctx.gate_batch(ops, a, b, c))
        assert np.array_equal(sk.decrypt_bool(got), want)
        ct = sk.encrypt_message(np.arange(count) % 2, 2, 9)
        lut = O.gen_lut(P, 2, lambda x: 1 - x)
        assert np.array_equal(multi.bootstrap_batch(ct, lut), ctx.bootstrap_batch(ct, lut))
        assert np.array_equal(multi.blind_rotate_batch(ct), ctx.blind_rotate_batch(ct))
        circ = T.circuit.ripple_carry_adder(2)
        inst = 11
        rng = np.random.default_rng(5)
        x, y = rng.integers(0, 4, inst), rng.integers(0, 4, inst)
        ins = np.stack([sk.encrypt_bool((x >> i) & 1, 30 + i) for i in range(2)] + [sk.encrypt_bool((y >> i) & 1, 40 + i) for i in range(2)])
        wires = np.concatenate([ins, np.broadcast_to(O.constant(P, False), (1, inst, P.n + 1))])
        g1 = multi.circuit_run(circ.gates, 5, wires, circ.out_wires)
        assert np.array_equal(g1, ctx.circuit_run(circ.gates, 5, wires, circ.out_wires))
        s = sum(sk.decrypt_bool(g1[i]).astype(np.int64) << i for i in range(2))
        assert np.array_equal(s, (x + y) % 4)
        with pytest.raises(T.TfheError):  # device pointers have no meaning on a group
            multi.bootstrap_batch_device(1, 0, 0)
    finally:
        multi.close()


def test_multi_device_generated_key(T, env):
    """tfhe_ctx_generate_cloudkey on a group: generated on the first device, replicated, every device computes with it."""
    P, sk, ck, ctx = env
    Tm = T
    sk2 = Tm.key.NewSecretKey(Tm.params.get("80"), 77)
    multi = Tm.Context(Tm.params.get("80"), devices="all")
    try:
        multi.generate_cloudkey(sk2.KeyLv0, sk2.KeyLv1, seed=5, with_ksk=True, export=False)
        bits = np.arange(64) % 2
        a, b = Tm.tlwe.EncryptBool(bits, sk2, 1), Tm.tlwe.EncryptBool(1 - bits, sk2, 2)
        assert np.array_equal(Tm.tlwe.DecryptBool(multi.gate_batch("OR", a, b), sk2), np.ones(64, dtype=np.uint8))
    finally:
        multi.close()


def test_device_gate_batch_is_async_and_in_place(T, env):
    """tfhe_gate_batch_device: opcodes are HOST memory, the call only enqueues (mixed batches included), and d_out may
    be d_a itself — prepared ciphertexts go to internal scratch, so a MUX never re-reads an overwritten input."""
    torch = pytest.importorskip("torch")
    P, sk, ck, ctx = env
    count = 29
    ops, a, b, c, want = _mixed(sk, count, 300)
    ref = ctx.gate_batch(ops, a, b, c)
    dev = torch.device("cuda", 0)
    da, db, dc = (torch.from_numpy(v.view(np.int32)).to(dev) for v in (a, b, c))
    stream = torch.cuda.current_stream()
    ctx.gate_batch_device(count, ops, da.data_ptr(), db.data_ptr(), dc.data_ptr(), da.data_ptr(), stream.cuda_stream)  # in place
    ctx.gate_batch_device(count, "COPY", da.data_ptr(), None, None, db.data_ptr(), stream.cuda_stream)              # chained, no sync between
    torch.cuda.synchronize()
    assert np.array_equal(da.cpu().numpy().view(np.uint32), ref)
    assert np.array_equal(db.cpu().numpy().view(np.uint32), ref)


def test_key_reload_without_ksk_drops_the_old_one(T, env, keyset):
    """loading a key without a key-switching key must not leave the previous one paired with the new BSK."""
    P, sk, ck, _ = env
    fresh = T.Context(T.params.get("80"), 0)
    try:
        fresh.load_cloudkey(ck.offset, ck.bsk_fft, ck.ksk, ck.testvec)
        ct = sk.encrypt_bool([1, 0], 3)
        fresh.bootstrap_batch(ct)
        fresh.load_cloudkey(ck.offset, ck.bsk_fft, None, ck.testvec)
        with pytest.raises(T.TfheError):
            fresh.bootstrap_batch(ct)
        assert fresh.blind_rotate_batch(ct).shape == (2, 2, P.N)
    finally:
        fresh.close()


def test_nonzero_k0_rows_of_an_uploaded_key_are_ignored(T, O, env):
    """the reference never reads KSK rows with digit 0 (trgsw/keyswitch.go:30); a key that carries garbage
    there must give the same words on every key-switch path (compacted gather, ordered split gather, tensor cores)."""
    P, sk, ck, ctx = env
    dirty = ck.ksk.copy().reshape(-1, P.n + 1)
    base = 1 << P.basebit
    dirty[::base] = 0xDEADBEEF
    other = T.Context(T.params.get("80"), 0)
    try:
        other.load_cloudkey(ck.offset, ck.bsk_fft, dirty, ck.testvec)
        rng = np.random.default_rng(3)
        for count in (1, 3, 400):  # split gather / gather / chunks of the contraction
            ext = rng.integers(0, 1 << 32, (count, P.N + 1), dtype=np.uint64).astype(np.uint32)
            for v in ("gather", "mma"):
                other.set_key_switch_variant(v)
                ctx.set_key_switch_variant(v)
                assert np.array_equal(other.key_switch_batch(ext), ctx.key_switch_batch(ext)), (count, v)
    finally:
        ctx.set_key_switch_variant("auto")
        other.close()


def test_c5_mixed_gates_sharded_by_index_reduced(T, O, keyset):
    """BASELINE.json configs[4] at reduced size (tools/bench_c5_multi.py and bench.py's c5 leg run it at 2^20): 2^12
    mixed AND/OR/XOR/MUX gate-ops at 128-bit whose operands are gathered from a ciphertext pool, cut into contiguous
    index ranges exactly as the ranks of a torchrun job cut them (sharding.shard_bounds), each range one host-buffer call.
    The concatenation must be the words of the single call and of the multi-device context, every output must decrypt
    to the plaintext gate, and a sample of gate-ops must equal the oracle word for word."""
    P, sk, ck = keyset("128")
    ctx = T.Context(T.params.get("128"), 0)
    multi = T.Context(T.params.get("128"), devices="all")
    try:
        ctx.load_cloudkey(ck.offset, ck.bsk_fft, ck.ksk, ck.testvec)
        multi.load_cloudkey(ck.offset, ck.bsk_fft, ck.ksk, ck.testvec)
        total, pool = 1 << 12, 256
        rng = np.random.default_rng(1)
        bits = rng.integers(0, 2, pool).astype(np.uint8)
        cts = sk.encrypt_bool(bits, 5)
        ia, ib, ic = (rng.integers(0, pool, total) for _ in range(3))
        ops = rng.integers(0, 4, total)
        names = ("AND", "OR", "XOR", "MUX")
        opcodes = np.array([T.OPCODES[o] for o in names], dtype=np.uint8)[ops]
        whole = ctx.gate_batch(opcodes, cts[ia], cts[ib], cts[ic])
        for world in (2, 3, 8):
            parts = []
            for lo, hi in T.sharding.shard_bounds(total, world):
                parts.append(ctx.gate_batch(opcodes[lo:hi], cts[ia[lo:hi]], cts[ib[lo:hi]], cts[ic[lo:hi]]))
            assert np.array_equal(np.concatenate(parts), whole), world
        assert np.array_equal(multi.gate_batch(opcodes, cts[ia], cts[ib], cts[ic]), whole)
        A, B, C = bits[ia], bits[ib], bits[ic]
        want = np.select([ops == 0, ops == 1, ops == 2], [A & B, A | B, A ^ B], np.where(A == 1, B, C))
        assert np.array_equal(sk.decrypt_bool(whole), want)
        for g in (0, 1, 2, 3, total - 1):
            if ops[g] == 3:
                ref = O.mux(ck, cts[ia[g]][None], cts[ib[g]][None], cts[ic[g]][None])[0]
            else:
                ref = O.gate_batch(ck, names[ops[g]], cts[ia[g]], cts[ib[g]])[0]
            assert np.array_equal(whole[g], ref), g
    finally:
        ctx.close()
        multi.close()


def test_circuit_graph_replay_is_identical(T, O, env):
    """tfhe_ctx_set_circuit_graph: call 1 of a circuit shape runs eagerly, call 2 records the level loop into a CUDA graph,
    calls 3+ replay it.  Every call must give the words of the plain runner on its own inputs; another shape gets its own
    graph; a reallocation of any device buffer in between forces a re-capture and still gives the same words."""
    P, sk, ck, _ = env
    ctx = T.Context(T.params.get("80"), 0)           # a fresh context: its scratch buffers start small
    ctx.load_cloudkey(ck.offset, ck.bsk_fft, ck.ksk, ck.testvec)
    circ = T.circuit.ripple_carry_adder(3)
    n_in = 7

    def wires(inst, seed):
        rng = np.random.default_rng(seed)
        x, y = rng.integers(0, 8, inst), rng.integers(0, 8, inst)
        ins = np.stack([sk.encrypt_bool((x >> i) & 1, seed + i) for i in range(3)] + [sk.encrypt_bool((y >> i) & 1, seed + 10 + i) for i in range(3)])
        return np.concatenate([ins, np.broadcast_to(O.constant(P, False), (1, inst, P.n + 1))]), (x + y) % 8

    plain = {}
    for inst, seed in ((9, 1), (9, 2), (9, 3), (9, 4), (5, 5), (5, 6), (5, 7)):
        w, _ = wires(inst, seed)
        plain[(inst, seed)] = ctx.circuit_run(circ.gates, n_in, w, circ.out_wires)
    try:
        ctx.set_circuit_graph(True)
        r0 = ctx.circuit_graph_replays
        for inst, seed in ((9, 1), (9, 2), (9, 3), (5, 5), (5, 6), (5, 7), (9, 4)):
            w, want = wires(inst, seed)
            got = ctx.circuit_run(circ.gates, n_in, w, circ.out_wires)
            assert np.array_equal(got, plain[(inst, seed)]), (inst, seed)
            s = sum(sk.decrypt_bool(got[i]).astype(np.int64) << i for i in range(3))
            assert np.array_equal(s, want)
        assert ctx.circuit_graph_replays - r0 == 3          # (9,3), (5,7), (9,4): third and later calls of a shape
        # a much larger batch grows the engine's scratch buffers: the recorded graphs hold stale pointers and must not be used
        big = sk.encrypt_bool(np.arange(3000) % 2, 77)
        ctx.gate_batch("NAND", big, big)
        r1 = ctx.circuit_graph_replays
        w, _ = wires(9, 2)
        assert np.array_equal(ctx.circuit_run(circ.gates, n_in, w, circ.out_wires), plain[(9, 2)])   # re-captured
        assert ctx.circuit_graph_replays == r1
        assert np.array_equal(ctx.circuit_run(circ.gates, n_in, w, circ.out_wires), plain[(9, 2)])   # replayed
        assert ctx.circuit_graph_replays == r1 + 1
    finally:
        ctx.set_circuit_graph(False)
        ctx.close()


def test_circuit_levels_through_the_tiled_key_switch(T, O, keyset):
    """A circuit on a large-base parameter set (Uint3: base 64) with enough instances that every level's key switch takes the
    shared-memory tile kernel, whose blocks then scatter to WIRE rows (job -> (gate of the level, instance) -> wire):
    the words must equal those of the same run with the row gather."""
    P, sk, ck = keyset("uint3")
    ctx = T.Context(T.params.get("uint3"), 0)
    try:
        ctx.load_cloudkey(ck.offset, ck.bsk_fft, ck.ksk, ck.testvec)
        circ = T.circuit.ripple_carry_adder(2)
        inst = 170
        rng = np.random.default_rng(3)
        ins = np.stack([sk.encrypt_bool(rng.integers(0, 2, inst).astype(np.uint8), 40 + i) for i in range(4)])
        wires = np.concatenate([ins, np.broadcast_to(O.constant(P, False), (1, inst, P.n + 1))])
        ctx.set_key_switch_variant("gather")
        ref = ctx.circuit_run(circ.gates, 5, wires, circ.out_wires)
        ctx.set_key_switch_variant("tile")
        got = ctx.circuit_run(circ.gates, 5, wires, circ.out_wires)
        ctx.set_key_switch_variant("auto")
        assert np.array_equal(got, ref)
        assert np.array_equal(ctx.circuit_run(circ.gates, 5, wires, circ.out_wires), ref)
    finally:
        ctx.close()
