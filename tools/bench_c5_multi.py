"""BASELINE.json configs[4]: 2^20 mixed AND/OR/XOR/MUX gate-ops at 128-bit, sharded contiguously by gate index over the
ranks of one node (torchrun), keys replicated by one NCCL broadcast, no collective on the hot path.  Strong scaling:
the total is fixed.  Every rank draws the same global op/index streams and evaluates its own slice; every output is
decrypted and checked.  Prints one JSON line on rank 0.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/bench_c5_multi.py [--log2 20]"""
import argparse, importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--log2", type=int, default=20)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
T = importlib.import_module("go-tfhe_b200")
P = T.params.get("128")
sk = T.key.NewSecretKey(P, 2024)
ck = T.cloudkey.NewCloudKey(sk, 2025) if rank == 0 else None
ctx = T.Context(P, local)
keys = T.sharding.broadcast_cloudkey(P, ck, dev, dist if world > 1 else None)
T.sharding.load_broadcast_key(ctx, keys, torch.cuda.current_stream().cuda_stream)
del keys, ck
torch.cuda.empty_cache()

total, pool = 1 << args.log2, 4096
rng = np.random.default_rng(1)                       # the same global streams on every rank
bits = rng.integers(0, 2, pool).astype(np.uint8)
cts = T.tlwe.EncryptBool(bits, sk, 5)
ia, ib, ic = (rng.integers(0, pool, total) for _ in range(3))
ops = rng.integers(0, 4, total)
lo, hi = T.sharding.shard_bounds(total, world)[rank]
sl = slice(lo, hi)
opcodes = np.array([T.OPCODES[o] for o in ("AND", "OR", "XOR", "MUX")], dtype=np.uint8)[ops[sl]]
a, b, c = cts[ia[sl]], cts[ib[sl]], cts[ic[sl]]
ctx.gate_batch(opcodes[:2048], a[:2048], b[:2048], c[:2048])   # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
out = ctx.gate_batch(opcodes, a, b, c)                # host buffers in, host buffers out, through the C ABI
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
A, B, C = bits[ia[sl]], bits[ib[sl]], bits[ic[sl]]
o = ops[sl]
want = np.select([o == 0, o == 1, o == 2], [A & B, A | B, A ^ B], np.where(A == 1, B, C))
ok = torch.tensor([int(np.array_equal(T.tlwe.DecryptBool(out, sk), want))], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    nboot = int(total + 2 * (ops == 3).sum())
    print(json.dumps({"config": "c5: 2^%d mixed AND/OR/XOR/MUX gate-ops, 128-bit, sharded by index" % args.log2, "n_gpus": world,
                      "gate_ops": total, "bootstraps": nboot, "seconds_max_over_ranks": float(dt.item()),
                      "gate_ops_per_s_e2e": total / float(dt.item()), "bootstraps_per_s_e2e": nboot / float(dt.item()),
                      "scaling": "strong", "correct": bool(ok.item())}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
