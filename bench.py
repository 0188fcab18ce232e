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
  cpu_baseline  the CPU oracle (C++ restatement of the reference's Go path; no Go toolchain exists here) on the
            host cores, bounded sample.  Only this leg and --impl reference touch oracle/.
"""
import argparse
import importlib
import atexit
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

    def stop(self, since=0.0):
        """Statistics over the samples that arrived after `since` (perf_counter): the sampler is started before the
        warm-up, because nvidia-smi can take longer to start than the whole timed region lasts (8 ranks at once)."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if ts < since:
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


def cpu_baseline(threads, gates_per_thread=16):
    """Oracle (port of the reference's Go path) on the host cores: one worker thread per core, private scratch —
    the trgsw.BatchBlindRotate goroutine-per-gate equivalent.  Returns gates/s on a bounded sample."""
    from oracle import oracle as O
    P = O.get_params(PARAMS)
    sk = O.SecretKey(P, 11)
    ck = O.CloudKey(sk, 12, threads=threads)
    count = threads * gates_per_thread
    import numpy as np
    rng = np.random.default_rng(5)
    A = rng.integers(0, 2, count).astype(np.uint8)
    B = rng.integers(0, 2, count).astype(np.uint8)
    a, b = sk.encrypt_bool(A, 1), sk.encrypt_bool(B, 2)
    O.gate_batch(ck, OP, a[:threads], b[:threads], threads=threads)  # warm
    t0 = time.perf_counter()
    out = O.gate_batch(ck, OP, a, b, threads=threads)
    dt = time.perf_counter() - t0
    assert np.array_equal(sk.decrypt_bool(out), 1 - (A & B))
    return count / dt, count, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU path cannot run here (pure Go, no toolchain), so this times
    the oracle port of it with every host thread, on the same config / metric."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import oracle as O
    import numpy as np
    P = O.get_params(PARAMS)
    sk = O.SecretKey(P, 11)
    ck = O.CloudKey(sk, 12, threads=threads)
    per_step = threads * 2
    rng = np.random.default_rng(5)
    A = rng.integers(0, 2, per_step).astype(np.uint8)
    B = rng.integers(0, 2, per_step).astype(np.uint8)
    a, b = sk.encrypt_bool(A, 1), sk.encrypt_bool(B, 2)
    for _ in range(args.warmup):
        O.gate_batch(ck, OP, a, b, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = O.gate_batch(ck, OP, a, b, threads=threads)
    dt = time.perf_counter() - t0
    assert np.array_equal(sk.decrypt_bool(out), 1 - (A & B))
    v = per_step * args.steps / dt
    sample = "%d NAND gates per step (2 per host thread), 128-bit, oracle C++ port of the Go path" % per_step
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batch 4096 NAND gates, 128-bit params (n=700, N=1024), per GPU", "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    T = importlib.import_module("go-tfhe_b200")
    P = T.params.get(PARAMS)
    n1 = P.n + 1
    ctx = T.Context(P, local)

    # --- cloud key: generated once on rank 0 (host client library), uploaded, and broadcast over NCCL -----------
    sk = T.key.NewSecretKey(P, 2024)           # deterministic: every rank derives the same secret key
    ck = T.cloudkey.NewCloudKey(sk, 2025) if rank == 0 else None
    keys = T.sharding.broadcast_cloudkey(P, ck, dev, dist if world > 1 else None)  # the one collective of the job
    del ck
    stream = torch.cuda.current_stream()
    T.sharding.load_broadcast_key(ctx, keys, stream.cuda_stream)
    del keys
    torch.cuda.empty_cache()

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
    clocks = sampler.stop(since=t_wall0)  # samples taken during the device-timed and the end-to-end timed regions
    got = out_h.numpy().view(np.uint32)
    if not np.array_equal(T.tlwe.DecryptBool(got, sk), 1 - (A & B)):
        raise SystemExit("rank %d: e2e outputs are wrong" % rank)

    # --- single-gate latency (BASELINE configs[0] shape: one NAND per call through the C ABI), rank 0 only, untimed part ----
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
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batch %d NAND gates, 128-bit params (n=700, N=1024), per GPU" % count,
                       "batch_per_gpu": count, "params": PARAMS, "op": OP,
                       "l2": "flushed between timed steps (256 MiB device fill outside the event pairs); keys (164 MiB) exceed L2",
                       "parallelism": "gates sharded by index, keys replicated (one NCCL broadcast at init)"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": 2 * count * n1 * 4 * world,
                    "d2h_bytes_per_step": count * n1 * 4 * world},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "blind_rotate_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": profile_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": br_bytes, "kernel_ms": br_ms,
                         "note": "frac > 1 is possible: co-resident gates share bootstrapping-key rows out of L2"},
            "roofline_fp64": {"algorithmic_tflops": flops / (br_ms * 1e-3) / 1e12, "flops_per_bootstrap": P.flops_per_bootstrap},
            "stage_ms": {"blind_rotate": br_ms, "key_switch": ks_ms, "share_blind_rotate": br_ms / (br_ms + ks_ms)},
            "single_gate_ms": single_ms,  # one NAND per C-ABI call with host buffers (latency kernel), same key and parameters
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            v, cnt, dt = cpu_baseline(threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d NAND gates (16 per host thread), 128-bit, %.1f s; oracle = C++ port of the "
                                              "reference's Go path (no Go toolchain in this image)" % (cnt, dt)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
