"""Multi-GPU plumbing: one process per GPU, gates sharded contiguously by index, cloud key replicated by ONE
broadcast at init (NCCL over NVLink on GPUs; gloo in the CPU tests).  There is no collective on the hot path."""
import numpy as np


def shard_bounds(count, world):
    """Contiguous split of [0, count) into `world` ranges whose sizes differ by at most one."""
    base, rem = divmod(int(count), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard_instances(n_instances, gates_per_instance, world):
    """Levelised circuits shard by INSTANCE so that no ciphertext ever crosses GPUs."""
    return [(lo * gates_per_instance, hi * gates_per_instance) for lo, hi in shard_bounds(n_instances, world)]


def broadcast_cloudkey(P, ck, device, dist=None):
    """rank 0 passes its CloudKey, the others pass None.  Returns (offset, bsk, ksk, testvec) as torch tensors on
    `device`, identical on every rank.  With dist=None (single process) it is a plain upload."""
    import torch
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    if rank == 0:
        bsk = torch.from_numpy(np.ascontiguousarray(ck.BootstrappingKey)).to(device)
        ksk = torch.from_numpy(np.ascontiguousarray(ck.KeySwitchingKey).view(np.int32)).to(device)
        tv = torch.from_numpy(np.ascontiguousarray(ck.BlindRotateTestvec).view(np.int32)).to(device)
        off = torch.tensor([int(ck.DecompositionOffset)], dtype=torch.int64, device=device)
    else:
        bsk = torch.empty((P.n, 2 * P.L, 2, P.N), dtype=torch.float64, device=device)
        ksk = torch.empty((P.ksk_rows, P.n + 1), dtype=torch.int32, device=device)
        tv = torch.empty((2, P.N), dtype=torch.int32, device=device)
        off = torch.zeros(1, dtype=torch.int64, device=device)
    if world > 1:
        for t in (bsk, ksk, tv, off):
            dist.broadcast(t, 0)
    return int(off.item()), bsk, ksk, tv


def load_broadcast_key(ctx, key_tensors, stream=0):
    """Hands the broadcast device buffers to tfhe_ctx_load_cloudkey_device."""
    off, bsk, ksk, tv = key_tensors
    ctx.load_cloudkey_device(off, bsk.data_ptr(), ksk.data_ptr(), tv.data_ptr(), stream)
