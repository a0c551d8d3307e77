"""Multi-GPU plumbing for the CTR path: blocks are independent, so they are sharded by
counter across ranks with no collective on the per-block path; the prepared keys (Fourier BSK and the two
tensor-layout integer keys, 1.04 GB) are replicated once with a broadcast from the rank that generated or
loaded them.  `torch.distributed` is plumbing only; every function takes the process group explicitly so
the logic is testable on CPU with gloo."""
import numpy as np


def shard_counters(first, blocks_per_rank, step, rank, world):
    """Counters of `rank` in step `step` when every rank processes `blocks_per_rank` blocks per step:
    the global stream first, first+1, ... is dealt out rank-major inside a step (main.rs:55-64 order)."""
    base = first + (step * world + rank) * blocks_per_rank
    return [base + b for b in range(blocks_per_rank)]


def shard_range(total, rank, world):
    """Contiguous split of `total` blocks (strong scaling, e.g. --number-of-outputs 1024 over 8 GPUs):
    returns (start, count); the first `total % world` ranks get one extra block."""
    q, r = divmod(total, world)
    start = rank * q + min(rank, r)
    return start, q + (1 if rank < r else 0)


class _DevPtr:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def broadcast_buffers(buffers, dist, src=0, device=None, group=None):
    """Broadcast raw device buffers [(ptr, nbytes), ...] in place (NCCL over NVLink on GPUs)."""
    import torch
    total = 0
    for ptr, nbytes in buffers:
        t = torch.as_tensor(_DevPtr(ptr, nbytes), device=device)
        dist.broadcast(t, src=src, group=group)
        total += nbytes
    return total


def replicate_keys(engine, dist, rank, src=0, device=None, group=None):
    """Rank `src` has loaded / generated keys; every other rank allocates its key buffers, receives them and
    marks them ready.  Returns the number of bytes broadcast."""
    if rank != src:
        engine.alloc_keys()
    n = broadcast_buffers(engine.key_buffers(), dist, src=src, device=device, group=group)
    if rank != src:
        engine.keys_ready()
    return n


def gather_blocks(local_blocks, dist, world, group=None):
    """All-gather of decrypted 16-byte blocks (verification harness only; ciphertexts never move)."""
    import torch
    t = torch.from_numpy(np.frombuffer(b"".join(local_blocks), dtype=np.uint8).copy())
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [bytes(o.numpy()) for o in out]
