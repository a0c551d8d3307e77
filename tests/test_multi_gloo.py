"""world_size-2 CPU test (gloo) of the multi-GPU host logic: counter sharding covers the CTR stream exactly
once, the key-replication hand-shake moves every byte, and the gathered verification agrees with FIPS-197.
The GPU path itself shards with no data-path collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from __graft_entry__ import load_package
    import orc
    pkg = load_package()
    sh = pkg.sharding
    # 1. weak-scaling shards of 3 steps x 4 blocks per rank
    mine = [c for step in range(3) for c in sh.shard_counters(1000, 4, step, rank, world)]
    t = torch.tensor(mine, dtype=torch.int64)
    allc = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allc, t)
    flat = sorted(int(x) for a in allc for x in a)
    assert flat == list(range(1000, 1000 + 3 * 4 * world)), flat
    # 2. strong-scaling split of 1024 (and a ragged 1023) blocks
    for total in (1024, 1023, 1, 0):
        s, c = sh.shard_range(total, rank, world)
        tt = torch.tensor([s, c], dtype=torch.int64)
        g = [torch.empty_like(tt) for _ in range(world)]
        dist.all_gather(g, tt)
        pos = 0
        for a in g:
            assert int(a[0]) == pos
            pos += int(a[1])
        assert pos == total
    # 3. key replication: broadcast of (stand-in) key buffers, same call sequence as on the GPUs
    bufs = [torch.full((n,), rank + 1, dtype=torch.uint8) for n in (1000, 77, 4096)]
    for b in bufs:
        dist.broadcast(b, src=0)
    assert all(int(b.min()) == 1 and int(b.max()) == 1 for b in bufs)
    # 4. each rank "encrypts" its CTR blocks in the clear and the gathered stream matches FIPS-197
    key, iv = bytes(range(16)), 2 ** 128 - 5
    local = [orc.clear_aes_encrypt(key, ((iv + c) % 2 ** 128).to_bytes(16, "big")) for c in sh.shard_counters(0, 3, 0, rank, world)]
    gathered = sh.gather_blocks(local, dist, world)
    stream = b"".join(gathered)
    for c in range(3 * world):
        assert stream[16 * c:16 * c + 16] == orc.clear_aes_encrypt(key, ((iv + c) % 2 ** 128).to_bytes(16, "big"))
    ret[rank] = 1
    dist.destroy_process_group()


def test_two_rank_sharding_and_replication():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: 1, 1: 1}
