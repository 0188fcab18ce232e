#!/usr/bin/env python
"""bench.py — bootstrapped gates/sec at 128-bit parameters (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (gate prologue -> blind rotate -> sample extract -> key switch) over
one batch of synthetic NAND gates: BASELINE.json configs[1], 4096 gates at 128-bit parameters per GPU
(weak scaling: every rank runs its own 4096-gate batch, keys replicated, no collective on the hot path).

  value     whole-job gates/s with the input ciphertexts already resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call tfhe_gate_batch: pinned HOST buffers in,
            host buffer out, H2D/D2H copies inside the timed region
  roofline  dominant kernel (blind_rotate_kernel): algorithmic bytes / measured kernel time vs measured HBM peak
  roofline_fp64  the same kernel against the FP64 roof measured in this process (tfhe_fp64_peak_probe): the binding one
  configs   BASELINE.json configs[2..4] next to the headline: c3 (8-bit adders x 1024 through the circuit runner),
            c4 (Uint5 programmable bootstraps, batch 2048), c5 (mixed AND/OR/XOR/MUX: the per-GPU share at N = 1; at
            N > 1 the whole 2^20 gate-ops STRONG-scaled over the ranks, and once more through ONE multi-device context)
  cpu_baseline  the CPU oracle (C++ restatement of the reference's Go path; no Go toolchain exists here) on the
            host cores, bounded sample.  Only this leg and --impl reference touch oracle/.
"""
import argparse
import importlib
import atexit
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bootstrapped gates/sec at 128-bit (n=700,N=1024)"
UNIT = "gates/s"
PARAMS = "128"
BATCH = 4096
OP = "NAND"
WORKLOAD = "batch %d NAND gates, 128-bit params (n=700, N=1024), per GPU"


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            atexit.register(self.proc.kill)  # never leave the sampler behind if the run aborts
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, since=0.0, until=None):
        """Statistics over the samples that arrived in [since, until] (perf_counter): the sampler is started before the
        warm-up, because nvidia-smi can take longer to start than the whole timed region lasts (8 ranks at once)."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if ts < since or (until is not None and ts > until):
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy or sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def profile_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if summarised."""
    try:
        with open(os.path.join(ROOT, "profiles", "blind_rotate_ncu_summary.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


def cpu_baseline(threads, target_s=12.0):
    """Oracle (port of the reference's Go path) on the host cores: one worker thread per core, private scratch —
    the trgsw.BatchBlindRotate goroutine-per-gate equivalent.  Returns gates/s on a bounded sample: a short probe (one
    gate per thread) sizes the sample to ~target_s seconds of work, capped at the workload's 4096 gates."""
    from oracle import oracle as O
    P = O.get_params(PARAMS)
    sk = O.SecretKey(P, 11)
    ck = O.CloudKey(sk, 12, threads=threads)
    import numpy as np
    rng = np.random.default_rng(5)
    A = rng.integers(0, 2, 4096).astype(np.uint8)
    B = rng.integers(0, 2, 4096).astype(np.uint8)
    a, b = sk.encrypt_bool(A, 1), sk.encrypt_bool(B, 2)
    O.gate_batch(ck, OP, a[:threads], b[:threads], threads=threads)  # warm
    t0 = time.perf_counter()
    O.gate_batch(ck, OP, a[:threads], b[:threads], threads=threads)
    per_round = max(time.perf_counter() - t0, 1e-3)
    count = int(min(4096, max(1, round(target_s / per_round)) * threads))
    t0 = time.perf_counter()
    out = O.gate_batch(ck, OP, a[:count], b[:count], threads=threads)
    dt = time.perf_counter() - t0
    assert np.array_equal(sk.decrypt_bool(out), 1 - (A[:count] & B[:count]))
    return count / dt, count, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path cannot run here (pure Go, no toolchain), so this times
    the oracle port of it with every host thread, on the same config / metric.  A step is the FULL 4096-gate batch when
    the whole run then fits ~4 minutes on this box's cores; otherwise the largest per-thread multiple that does."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import oracle as O
    import numpy as np
    P = O.get_params(PARAMS)
    sk = O.SecretKey(P, 11)
    ck = O.CloudKey(sk, 12, threads=threads)
    rng = np.random.default_rng(5)
    A = rng.integers(0, 2, BATCH).astype(np.uint8)
    B = rng.integers(0, 2, BATCH).astype(np.uint8)
    a, b = sk.encrypt_bool(A, 1), sk.encrypt_bool(B, 2)
    cal = 2 * threads
    O.gate_batch(ck, OP, a[:threads], b[:threads], threads=threads)
    t0 = time.perf_counter()
    O.gate_batch(ck, OP, a[:cal], b[:cal], threads=threads)
    rate = cal / (time.perf_counter() - t0)
    budget_s = 240.0
    per_step = BATCH
    if BATCH * (args.steps + args.warmup) / rate > budget_s:
        per_step = max(threads, int(rate * budget_s / (args.steps + args.warmup)) // threads * threads)
    per_step = min(per_step, BATCH)
    a, b, A, B = a[:per_step], b[:per_step], A[:per_step], B[:per_step]
    for _ in range(args.warmup):
        O.gate_batch(ck, OP, a, b, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = O.gate_batch(ck, OP, a, b, threads=threads)
    dt = time.perf_counter() - t0
    assert np.array_equal(sk.decrypt_bool(out), 1 - (A & B))
    v = per_step * args.steps / dt
    full = per_step == BATCH
    sample = ("%d NAND gates per step (%s), 128-bit, oracle C++ port of the Go path on %d host threads"
              % (per_step, "the full batch" if full else "bounded sample: the full 4096 would take %.0f s per step here" % (BATCH / rate),
                 threads))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % BATCH, "batch_per_gpu": per_step, "params": PARAMS, "op": OP, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ---- the other BASELINE.json configs (each returns a dict; all outputs are decrypted and checked) -----------------
def config_c3(T, np, ctx, sk, instances=1024, bits=8, reps=2):
    """configs[2]: 8-bit ripple-carry adder (40 bootstraps, 17 dependent levels) x `instances`, host buffers in/out
    through tfhe_circuit_run."""
    P = ctx.P
    rng = np.random.default_rng(14)
    x, y = rng.integers(0, 1 << bits, instances), rng.integers(0, 1 << bits, instances)
    circ = T.circuit.ripple_carry_adder(bits)
    ins = np.stack([T.tlwe.EncryptBool((x >> i) & 1, sk, 500 + i) for i in range(bits)] +
                   [T.tlwe.EncryptBool((y >> i) & 1, sk, 600 + i) for i in range(bits)] +
                   [np.broadcast_to(T.gates.Constant(False, P), (instances, P.n + 1))])
    run = lambda: ctx.circuit_run(circ.gates, 2 * bits + 1, ins, circ.out_wires)
    out = run()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = run()
    dt = (time.perf_counter() - t0) / reps
    # the same with the level loop replayed from a CUDA graph (tfhe_ctx_set_circuit_graph): call 1 was the eager run above,
    # call 2 records, calls 3.. replay
    graph = {}
    try:
        ctx.set_circuit_graph(True)
        run(); run()
        r0 = ctx.circuit_graph_replays
        t0 = time.perf_counter()
        for _ in range(reps):
            out_g = run()
        dt_g = (time.perf_counter() - t0) / reps
        graph = {"bootstraps_per_s_cuda_graph": instances * circ.n_bootstraps / dt_g, "cuda_graph_replays": ctx.circuit_graph_replays - r0,
                 "cuda_graph_same_words": bool(np.array_equal(out_g, out))}
    finally:
        ctx.set_circuit_graph(False)
    s = sum(T.tlwe.DecryptBool(out[i], sk).astype(np.int64) << i for i in range(bits))
    return {**graph, "workload": "%d-bit ripple-carry adder (%d bootstraps, %d levels) x %d instances, tfhe_circuit_run, host buffers"
                        % (bits, circ.n_bootstraps, circ.n_levels, instances),
            "bootstraps_per_s": instances * circ.n_bootstraps / dt, "adders_per_s": instances / dt, "seconds": dt,
            "correct": bool(np.array_equal(s, (x + y) % (1 << bits)))}


def config_c4(T, np, device, count=2048, reps=3):
    """configs[3]: programmable bootstrap, Uint5 (n=1071, N=2048, msgMod 32), batch 2048, a LUT per ciphertext
    (identity, x mod 16, x >= 16: examples/add_two_numbers/main.go:59-72), host buffers through tfhe_bootstrap_batch."""
    P = T.params.get("uint5")
    sk = T.key.NewSecretKey(P, 31)
    ctx = T.Context(P, device)
    try:
        t0 = time.perf_counter()
        ctx.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=32, with_ksk=True, export=False)
        keygen_s = time.perf_counter() - t0
        rng = np.random.default_rng(15)
        m = 32
        msgs = rng.integers(0, m, count)
        ct = T.tlwe.EncryptLWEMessage(msgs, m, sk, 77)
        fs = [lambda v: v, lambda v: v % 16, lambda v: int(v >= 16)]
        luts = np.stack([T.lut.NewGenerator(m, P).GenLookUpTable(f).Poly for f in fs]).reshape(3, -1)
        sel = rng.integers(0, 3, count)
        per = np.ascontiguousarray(luts[sel])
        out_per = ctx.bootstrap_batch(ct, per)              # a full 16 KiB LUT per ciphertext crosses the bus
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.bootstrap_batch(ct, per)
        dt_per = (time.perf_counter() - t0) / reps
        # the three LUTs once + an index per ciphertext; ciphertexts in and out through PINNED host buffers, straight
        # through the C ABI (what the cgo shim does with C.malloc'ed / registered slices)
        import torch
        ct_h = torch.from_numpy(ct).pin_memory()
        out_h = torch.empty_like(ct_h).pin_memory()
        luts_c = np.ascontiguousarray(luts, dtype=np.uint32)
        idx_c = np.ascontiguousarray(sel, dtype=np.int32)

        def call():
            rc = ctx.lib.tfhe_bootstrap_batch_indexed(ctx.h, count, ct_h.data_ptr(), luts_c.ctypes.data, len(luts_c), idx_c.ctypes.data,
                                                      out_h.data_ptr())
            if rc:
                raise RuntimeError("tfhe_bootstrap_batch_indexed failed (%d)" % rc)
        call()
        ctx.set_timing(True)
        ctx.collect_timing()
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        dt = (time.perf_counter() - t0) / reps
        tm = ctx.collect_timing()
        out = out_h.numpy().copy()
        same_api = bool(np.array_equal(ctx.bootstrap_batch_indexed(ct, luts, sel), out))   # the numpy wrapper, pageable buffers
        want = np.array([fs[k](int(v)) for k, v in zip(sel, msgs)])
        return {"workload": "programmable bootstrap, Uint5 (n=1071, N=2048, msgMod 32), batch %d, 3 LUTs chosen per ciphertext "
                            "(tfhe_bootstrap_batch_indexed), pinned host buffers" % count,
                "bootstraps_per_s": count / dt, "seconds": dt,
                "bootstraps_per_s_lut_per_ciphertext": count / dt_per, "same_words_both_ways": bool(np.array_equal(out, out_per)) and same_api,
                "blind_rotate_ms": tm["blind_rotate_ms"] / max(tm["blind_rotate_launches"], 1),
                "key_switch_ms": tm["key_switch_ms"] / max(tm["key_switch_launches"], 1),
                "device_keygen_s": keygen_s,
                "correct": bool(np.array_equal(T.tlwe.DecryptLWEMessage(out, m, sk), want))}
    finally:
        ctx.close()


def c5_inputs(T, np, sk, total, lo, hi, pool=4096):
    """The global op / index streams of configs[4] (seeded, identical on every rank), materialised for [lo, hi)."""
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2, pool).astype(np.uint8)
    cts = T.tlwe.EncryptBool(bits, sk, 5)
    ia, ib, ic = (rng.integers(0, pool, total) for _ in range(3))
    ops = rng.integers(0, 4, total)
    sl = slice(lo, hi)
    opcodes = np.array([T.OPCODES[o] for o in ("AND", "OR", "XOR", "MUX")], dtype=np.uint8)[ops[sl]]
    A, B, C = bits[ia[sl]], bits[ib[sl]], bits[ic[sl]]
    o = ops[sl]
    want = np.select([o == 0, o == 1, o == 2], [A & B, A | B, A ^ B], np.where(A == 1, B, C))
    nboot_total = int(total + 2 * (ops == 3).sum())
    return opcodes, cts[ia[sl]], cts[ib[sl]], cts[ic[sl]], want, nboot_total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the c3 / c4 / c5 block (headline config only)")
    ap.add_argument("--c5-log2", type=int, default=20, help="log2 of the gate-ops of configs[4] (whole job)")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")   # host-side waits (see the one-call multi-device leg)

    T = importlib.import_module("go-tfhe_b200")
    P = T.params.get(PARAMS)
    n1 = P.n + 1
    ctx = T.Context(P, local)
    stream = torch.cuda.current_stream()

    # --- cloud key: generated ON THE DEVICE by rank 0 (tfhe_ctx_generate_cloudkey, ~30 ms), then the one collective of
    # the job: an NCCL broadcast of the reference-layout key to the other ranks, which repack it locally ------------
    sk = T.key.NewSecretKey(P, 2024)           # the client's key; in this synthetic job every rank can derive it
    t_key = time.perf_counter()
    if world == 1:
        ctx.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=2025, with_ksk=True, export=False)
    else:
        ck = None
        if rank == 0:
            off, tv, ksk, bsk = ctx.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=2025, with_ksk=True, export=True)
            ck = T.cloudkey.CloudKey(P, off, tv, ksk, bsk)
        keys = T.sharding.broadcast_cloudkey(P, ck, dev, dist)
        if rank != 0:
            T.sharding.load_broadcast_key(ctx, keys, stream.cuda_stream)
        del keys, ck
        torch.cuda.empty_cache()
    torch.cuda.synchronize()
    key_setup_s = time.perf_counter() - t_key

    # --- synthetic inputs: fresh encryptions of uniform bits, different per rank ----------------------------
    count = args.batch
    rng = np.random.default_rng(1000 + rank)
    A = rng.integers(0, 2, count).astype(np.uint8)
    B = rng.integers(0, 2, count).astype(np.uint8)
    a_h = torch.from_numpy(T.tlwe.EncryptBool(A, sk, 10 + 2 * rank).view(np.int32)).pin_memory()
    b_h = torch.from_numpy(T.tlwe.EncryptBool(B, sk, 11 + 2 * rank).view(np.int32)).pin_memory()
    out_h = torch.empty((count, n1), dtype=torch.int32).pin_memory()
    a_d, b_d = a_h.to(dev), b_h.to(dev)
    out_d = torch.empty((count, n1), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.int8, device=dev)  # > 126 MB L2

    def step_device():
        ctx.gate_batch_device(count, OP, a_d.data_ptr(), b_d.data_ptr(), None, out_d.data_ptr(), stream.cuda_stream)

    def step_host():
        ctx.lib.tfhe_gate_batch(ctx.h, count, opv.ctypes.data, 1, a_h.data_ptr(), b_h.data_ptr(), None, out_h.data_ptr())

    opv = np.array([T.OPCODES[OP]], dtype=np.uint8)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- warm-up (also validates the result) -------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    got = out_d.cpu().numpy().view(np.uint32)
    if not np.array_equal(T.tlwe.DecryptBool(got, sk), 1 - (A & B)):
        raise SystemExit("rank %d: decrypted NAND outputs are wrong" % rank)

    # --- timed region: K steps, inputs resident in HBM, L2 flushed between steps (outside the event pairs) --
    ctx.set_timing(True)
    ctx.collect_timing()
    launches0 = ctx.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    # one more untimed step in front: the host enqueues all K timed steps (a few ms) while the GPU is busy with it, so no
    # timed kernel ever starts on a GPU that sat idle waiting for a descheduled host thread (seen once: 0.6 s stall, -0.7 %)
    step_device()
    launches0 = ctx.kernel_launches       # host-side counter: the head-start step's launches are already counted
    for s0, s1 in ev:
        flush.fill_(1)
        s0.record(stream)
        step_device()
        s1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.kernel_launches - launches0
    stage = ctx.collect_timing()
    ctx.set_timing(False)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # --- e2e: host buffers through the C ABI, copies inside the timed region --------------------------------
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_wall_end = time.perf_counter()
    got = out_h.numpy().view(np.uint32)
    if not np.array_equal(T.tlwe.DecryptBool(got, sk), 1 - (A & B)):
        raise SystemExit("rank %d: e2e outputs are wrong" % rank)

    # --- FP64 roof of this device, measured in this process ------------------------------------------------------
    fp64 = (ctypes.c_double * 3)()
    fp64_ok = ctx.lib.tfhe_fp64_peak_probe(local, ctypes.byref(fp64)) == 0

    # --- single-gate latency (BASELINE configs[0] shape: one NAND per call through the C ABI), rank 0 only ----
    single_ms = None
    if rank == 0:
        a1, b1, o1 = a_h[0:1].clone(), b_h[0:1].clone(), torch.empty((1, n1), dtype=torch.int32)
        call = lambda: ctx.lib.tfhe_gate_batch(ctx.h, 1, opv.ctypes.data, 1, a1.data_ptr(), b1.data_ptr(), None, o1.data_ptr())
        for _ in range(3):
            call()
        t1 = time.perf_counter()
        for _ in range(10):
            call()
        single_ms = (time.perf_counter() - t1) / 10 * 1e3
        if int(T.tlwe.DecryptBool(o1.numpy().view(np.uint32).reshape(1, -1), sk)[0]) != int(1 - (A[0] & B[0])):
            raise SystemExit("single-gate output is wrong")
    clocks = sampler.stop(since=t_wall0, until=t_wall_end)  # samples taken during the device-timed and the end-to-end timed regions

    # --- the other BASELINE configs ---------------------------------------------------------------------------------
    configs = None
    if not args.no_configs:
        configs = {}
        del a_d, b_d, out_d, flush
        torch.cuda.empty_cache()
        # c5: 2^20 mixed gate-ops.  N = 1: this GPU's 1/8 share (2^17).  N > 1: the whole job, sharded contiguously by
        # gate index over the ranks (strong scaling), each rank one tfhe_gate_batch call on host buffers.
        total = 1 << args.c5_log2
        if world == 1:
            total, share = total // 8, "1/8 share of 2^%d (one GPU of eight)" % args.c5_log2
        else:
            share = "all 2^%d gate-ops, sharded by index over %d ranks" % (args.c5_log2, world)
        lo, hi = T.sharding.shard_bounds(total, world)[rank]
        ops5, a5, b5, c5, want5, nboot5 = c5_inputs(T, np, sk, total, lo, hi)
        # pinned host buffers through the C ABI (as the headline e2e); the warm-up call is long enough (2.5 pipeline chunks)
        # to size both staging slots, so the timed call measures the steady state of a serving process, not cudaMalloc
        a5p, b5p, c5p = (torch.from_numpy(x).pin_memory() for x in (a5, b5, c5))
        out5p = torch.empty_like(a5p).pin_memory()
        ops5 = np.ascontiguousarray(ops5, dtype=np.uint8)

        def c5_call(cnt):
            rc = ctx.lib.tfhe_gate_batch(ctx.h, cnt, ops5.ctypes.data, cnt, a5p.data_ptr(), b5p.data_ptr(), c5p.data_ptr(), out5p.data_ptr())
            if rc:
                raise RuntimeError("tfhe_gate_batch failed (%d)" % rc)
        c5_call(min(len(ops5), 40960))
        barrier()
        t0 = time.perf_counter()
        c5_call(len(ops5))
        torch.cuda.synchronize()
        t5 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        out5 = out5p.numpy()
        ok5 = torch.tensor([int(np.array_equal(T.tlwe.DecryptBool(out5, sk), want5))], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            dist.all_reduce(ok5, op=dist.ReduceOp.MIN)
        configs["c5_mixed"] = {"workload": "mixed AND/OR/XOR/MUX (uniform), 128-bit, %s; one tfhe_gate_batch call per rank, pinned host buffers" % share,
                               "scaling": "strong" if world > 1 else "per-GPU share", "gate_ops": total, "bootstraps": nboot5,
                               "seconds_max_over_ranks": float(t5.item()), "gate_ops_per_s": total / float(t5.item()),
                               "bootstraps_per_s": nboot5 / float(t5.item()), "correct": bool(ok5.item())}
        del a5, b5, c5, out5, a5p, b5p, c5p, out5p
        # c3 / c4 on every rank (weak: per-GPU workloads), reported as the sum over ranks of rank-local rates
        c3 = config_c3(T, np, ctx, sk)
        c4 = config_c4(T, np, local)
        for name, c in (("c3_adder", c3), ("c4_pbs_uint5", c4)):
            v = torch.tensor([c["bootstraps_per_s"], float(c["correct"])], dtype=torch.float64, device=dev)
            if world > 1:
                agg = v.clone()
                dist.all_reduce(agg, op=dist.ReduceOp.SUM)
                c["bootstraps_per_s_all_gpus"] = float(agg[0])
                c["correct"] = bool(agg[1] == world)
                c["scaling"] = "weak (per-GPU workload; bootstraps_per_s is rank 0's, _all_gpus the sum)"
            configs[name] = c
        configs["c3_adder"]["vs_c2_rate"] = configs["c3_adder"]["bootstraps_per_s"] / (count / (dev_ms / args.steps * 1e-3))
        # the same c5 job once more through ONE multi-device context in ONE process (rank 0; the other ranks idle at the
        # CPU-side barrier): tfhe_ctx_create_multi shards a single tfhe_gate_batch call over every GPU behind the C ABI
        barrier()
        if world > 1 and rank == 0:
            multi = T.Context(P, devices=list(range(world)))
            try:
                multi.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=2025, with_ksk=True, export=False)
                ops5, a5, b5, c5, want5, nboot5 = c5_inputs(T, np, sk, total, 0, total)
                multi.gate_batch(ops5[:4096], a5[:4096], b5[:4096], c5[:4096])
                t0 = time.perf_counter()
                out5 = multi.gate_batch(ops5, a5, b5, c5)
                dt = time.perf_counter() - t0
                configs["c5_one_call_multi_device"] = {
                    "workload": "the same 2^%d gate-ops as ONE tfhe_gate_batch call on a %d-GPU context (tfhe_ctx_create_multi), pageable host buffers"
                                % (args.c5_log2, world),
                    "devices": multi.device_count, "gate_ops": total, "bootstraps": nboot5, "seconds": dt,
                    "gate_ops_per_s": total / dt, "bootstraps_per_s": nboot5 / dt,
                    "correct": bool(np.array_equal(T.tlwe.DecryptBool(out5, sk), want5))}
            finally:
                multi.close()
        if world > 1:   # the idle ranks wait on the CPU (gloo): a spinning NCCL barrier kernel would time-slice GPUs 1.. with the job
            dist.barrier(group=cpu_group)
        barrier()

    # --- reduce over ranks: max time --------------------------------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(times[0]), float(times[1])

    if rank == 0:
        total = count * world * args.steps
        value = total / (dev_ms_max * 1e-3)
        e2e_v = total / (e2e_ms_max * 1e-3)
        br_ms = stage["blind_rotate_ms"] / max(stage["blind_rotate_launches"], 1)
        ks_ms = stage["key_switch_ms"] / max(stage["key_switch_launches"], 1)
        # algorithmic bytes of one blind_rotate_kernel launch: per gate the n BK row-sets (n*2L*2*N*8) plus its
        # ciphertext in and extracted LWE out (DESIGN.md "Roofline accounting")
        br_bytes = count * (P.n * 2 * P.L * 2 * P.N * 8 + n1 * 4 + (P.N + 1) * 4)
        peak, peak_src = measured_peaks()
        achieved = br_bytes / (br_ms * 1e-3) / 1e9
        flops = count * P.flops_per_bootstrap
        alg_tflops = flops / (br_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % count,
                       "batch_per_gpu": count, "params": PARAMS, "op": OP,
                       "l2": "flushed between timed steps (256 MiB device fill outside the event pairs); keys (164 MiB) exceed L2; one untimed step runs in front of the K timed ones so that their launches are all queued before they execute (stage_ms averages include its kernels)",
                       "parallelism": "gates sharded by index, keys replicated (one NCCL broadcast at init)"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": 2 * count * n1 * 4 * world,
                    "d2h_bytes_per_step": count * n1 * 4 * world},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "blind_rotate_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": profile_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": br_bytes, "kernel_ms": br_ms,
                         "note": "frac > 1 is possible: co-resident gates share bootstrapping-key rows out of L2; "
                                 "the binding roof is FP64 / shared memory, see roofline_fp64"},
            "roofline_fp64": {"bound": "fp64", "kernel": "blind_rotate_kernel", "algorithmic_tflops": alg_tflops,
                              "flops_per_bootstrap": P.flops_per_bootstrap,
                              "peak_tflops": fp64[0] if fp64_ok else None,
                              "peak_thread_dfma_per_clk_per_sm": fp64[1] if fp64_ok else None,
                              "frac": (alg_tflops / fp64[0]) if fp64_ok and fp64[0] > 0 else None,
                              "peak_source": "DFMA microbenchmark in this process (tfhe_fp64_peak_probe: independent chains, "
                                             "reuse-cache operands, 2 flops per DFMA)",
                              "note": "algorithmic flops = the reference's radix-2 count (SURVEY 8d); the kernel executes "
                                      "2290 FP64 instructions per thread-step in radix-8 FMA form"},
            "stage_ms": {"blind_rotate": br_ms, "key_switch": ks_ms, "share_blind_rotate": br_ms / (br_ms + ks_ms)},
            "single_gate_ms": single_ms,  # one NAND per C-ABI call with host buffers (latency kernel), same key and parameters
            "wall_s_timed_region": t_wall,
            "key_setup_s": key_setup_s,
        }
        if configs is not None:
            line["configs"] = configs
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            v, cnt, dt = cpu_baseline(threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d NAND gates of the 4096-gate workload, 128-bit, %.1f s; oracle = C++ port of the "
                                              "reference's Go path (no Go toolchain in this image)" % (cnt, dt)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
