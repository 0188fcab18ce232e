"""World-size-2 gloo test of the multi-GPU plumbing (no GPU needed): the key broadcast delivers identical keys, the
contiguous gate shards cover the batch exactly once, and sharded results gathered on rank 0 equal the unsharded
result.  The per-shard compute stands in with the CPU oracle (tests may use it); on GPUs the same code path calls
the CUDA engine (bench.py)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_exactly_once():
    T = importlib.import_module("go-tfhe_b200")
    for count in (0, 1, 5, 4096, 1 << 20):
        for world in (1, 2, 3, 8):
            b = T.sharding.shard_bounds(count, world)
            assert b[0][0] == 0 and b[-1][1] == count
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert T.sharding.shard_instances(1024, 40, 8)[3] == (3 * 128 * 40, 4 * 128 * 40)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    T = importlib.import_module("go-tfhe_b200")
    from oracle import oracle as O
    P = T.params.get("80")
    sk = T.key.NewSecretKey(P, 7)          # same seed on every rank -> same secret key (client side)
    ck = T.cloudkey.NewCloudKey(sk, 8) if rank == 0 else None
    off, bsk, ksk, tv = T.sharding.broadcast_cloudkey(P, ck, torch.device("cpu"), dist)
    # every rank now holds the same key material
    digest = torch.tensor([float(bsk.double().abs().sum()), float(ksk.long().sum()), float(off)], dtype=torch.float64)
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    assert all(torch.equal(gathered[0], g) for g in gathered)

    count = 5
    A = np.array([0, 1, 1, 0, 1], dtype=np.uint8)
    B = np.array([1, 1, 0, 0, 1], dtype=np.uint8)
    a, b = T.tlwe.EncryptBool(A, sk, 1), T.tlwe.EncryptBool(B, sk, 2)
    lo, hi = T.sharding.shard_bounds(count, world)[rank]

    class OCK:
        pass
    o = OCK()
    o.P, o.testvec, o.ksk, o.bsk_fft, o.offset = (O.get_params("80"), tv.numpy().view(np.uint32).ravel(),
                                                  ksk.numpy().view(np.uint32), bsk.numpy(), off)
    mine = O.gate_batch(o, "NAND", a[lo:hi], b[lo:hi], threads=2) if hi > lo else np.zeros((0, P.n + 1), np.uint32)
    # gather variable-size shards on rank 0 (outputs only; inputs never move between ranks)
    padded = torch.zeros((count, P.n + 1), dtype=torch.int32)
    padded[lo:hi] = torch.from_numpy(mine.view(np.int32))
    dist.reduce(padded, 0, op=dist.ReduceOp.SUM)  # shards are disjoint, so SUM == concatenation
    if rank == 0:
        full = O.gate_batch(o, "NAND", a, b, threads=2)
        ok = np.array_equal(padded.numpy().view(np.uint32), full) and \
            list(T.tlwe.DecryptBool(full, sk)) == list(1 - (A & B))
        with open(out_path, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_key_broadcast_and_sharded_gates(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"
