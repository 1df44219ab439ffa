"""Host-side plumbing of the decomposed run: one z-slab of the block per rank (the reference's one subdomain per MPI rank,
tutorials/AMB2018-02-B/system/decomposeParDict:18-20; here the geometric equivalent of OpenFOAM's `simple` (1 1 N)).

Nothing here touches the data path: halo planes and Krylov reductions move inside the CUDA library over NCCL.  These
helpers give every rank its slab, distribute the NCCL unique id, and reassemble fields for checks.
"""
import numpy as np


def slab_range(nz, rank, nranks):
    """Global k range [k0, k1) of `rank`; identical to the library's and the oracle's split."""
    return nz * rank // nranks, nz * (rank + 1) // nranks


def local_cells(n, rank, nranks):
    k0, k1 = slab_range(n[2], rank, nranks)
    return n[0] * n[1] * (k1 - k0)


def broadcast_unique_id(make_id, rank, dist, device=None):
    """Rank 0 creates the 128-byte ncclUniqueId with `make_id()`; everybody gets it through torch.distributed."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        t = torch.tensor(list(make_id()), dtype=torch.uint8, device=device)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def gather_field(local, n, rank, nranks, dist, device=None):
    """All ranks' slabs concatenated in rank order == the global field in OpenFOAM cell order (k slowest)."""
    import torch
    sizes = [local_cells(n, r, nranks) for r in range(nranks)]
    assert local.size == sizes[rank]
    mx = max(sizes)
    buf = torch.zeros(mx, dtype=torch.float64, device=device)
    buf[: local.size] = torch.from_numpy(np.ascontiguousarray(local)).to(buf.device)
    out = [torch.zeros(mx, dtype=torch.float64, device=device) for _ in range(nranks)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(out, sizes)])


def max_over_ranks(value, dist, device=None):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def enable_peer_allreduce(ctx, dist, device=None):
    """All ranks of one node: exchange the IPC handles of the all-reduce mailboxes and switch the library's scalar
    all-reduces from NCCL to the NVLink peer-memory exchange."""
    import torch
    mine = torch.tensor(list(ctx.peer_export()), dtype=torch.uint8, device=device)
    out = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(ctx.nranks)]
    dist.all_gather(out, mine)
    ctx.peer_import(b"".join(bytes(o.cpu().tolist()) for o in out))
