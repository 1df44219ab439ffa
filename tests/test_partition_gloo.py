"""World-size-2 (and 3) runs of the host-side decomposition plumbing over the gloo backend, on the CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from additivefoam_b200 import partition


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank derives the same unique id bytes
        uid = partition.broadcast_unique_id(lambda: bytes(range(128)), rank, dist)
        assert uid == bytes(range(128))
        # slab of a global field (the cell index itself) gathers back to the global ordering
        k0, k1 = partition.slab_range(n[2], rank, world)
        plane = n[0] * n[1]
        local = np.arange(k0 * plane, k1 * plane, dtype=np.float64)
        full = partition.gather_field(local, n, rank, world, dist)
        assert np.array_equal(full, np.arange(n[0] * n[1] * n[2], dtype=np.float64))
        # step time is the max over ranks
        assert partition.max_over_ranks(10.0 + rank, dist) == 10.0 + world - 1
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, (5, 4, 7)), (3, (3, 2, 8))])
def test_slabs_over_gloo(tmp_path, world, n):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_slab_ranges_tile_the_block():
    for nz in (1, 7, 32, 250):
        for R in (1, 2, 3, 8):
            if R > nz:
                continue
            r = [partition.slab_range(nz, q, R) for q in range(R)]
            assert r[0][0] == 0 and r[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(b > a for a, b in r)
