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


def enable_peer(ctx, dist, device=None, halo_doubles=0):
    """All ranks of one node: exchange the IPC handles of the peer mailboxes and switch the library's scalar all-reduces -- and,
    with `halo_doubles` >= nx*ny, its halo-plane exchanges -- from NCCL to stores into peer memory over NVLink."""
    import torch
    mine = torch.tensor(list(ctx.peer_export(halo_doubles)), dtype=torch.uint8, device=device)
    out = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(ctx.nranks)]
    dist.all_gather(out, mine)
    ctx.peer_import(b"".join(bytes(o.cpu().tolist()) for o in out))


def enable_peer_allreduce(ctx, dist, device=None):
    """The all-reduce mailbox only (halo planes stay on ncclSend/ncclRecv)."""
    enable_peer(ctx, dist, device=device, halo_doubles=0)


def decompose_ldu(n_cells, lower, upper, diag, upper_coeffs, lower_coeffs, rank, nranks):
    """One rank's share of a global LDU matrix, the way decomposePar hands it to the solver: the cells
    [n*rank/nranks, n*(rank+1)/nranks) in their global order (which keeps the faces upper-triangular), the faces with both
    ends inside as the local matrix, and one processor interface per neighbouring rank for the faces that were cut --
    faceCells in global face order on both sides, so that the two sides' lists pair up, and
    interfaceBouCoeffs = -(the coefficient that multiplied the remote cell): [UPSTREAM] lduMatrix::Amul subtracts
    interfaceBouCoeffs * psi_neighbour.  Returns a dict; `lower_coeffs` may be None (symmetric)."""
    lower, upper = np.asarray(lower), np.asarray(upper)
    lc = upper_coeffs if lower_coeffs is None else lower_coeffs
    bounds = [n_cells * r // nranks for r in range(nranks + 1)]
    owner_of = lambda c: np.searchsorted(bounds, c, side="right") - 1      # noqa: E731
    c0, c1 = bounds[rank], bounds[rank + 1]
    rl, ru = owner_of(lower), owner_of(upper)
    inside = (rl == rank) & (ru == rank)
    out = dict(c0=c0, c1=c1, n=c1 - c0, lower=(lower[inside] - c0).astype(np.int32), upper=(upper[inside] - c0).astype(np.int32),
               diag=np.ascontiguousarray(diag[c0:c1]), upper_coeffs=np.ascontiguousarray(upper_coeffs[inside]),
               lower_coeffs=None if lower_coeffs is None else np.ascontiguousarray(lower_coeffs[inside]), interfaces=[], bou_coeffs=[])
    for other in range(nranks):
        if other == rank:
            continue
        mine_low = (rl == rank) & (ru == other)        # my cell is the lower address: it is multiplied by upper[f]
        mine_up = (ru == rank) & (rl == other)         # my cell is the upper address: lower[f]
        cut = mine_low | mine_up
        if not cut.any():
            continue
        face_cells = np.where(mine_low, lower, upper)[cut] - c0
        coeffs = -np.where(mine_low, upper_coeffs, lc)[cut]
        out["interfaces"].append((other, face_cells.astype(np.int32)))
        out["bou_coeffs"].append(np.ascontiguousarray(coeffs))
    return out
