"""Shared test helpers: seeded LDU problems (SURVEY.md 8d micro-bench definition) and error norms."""
import numpy as np


def rel_linf(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def block_addressing_np(n):
    """Upper-triangular LDU addressing of an nx x ny x nz block, numpy only."""
    nx, ny, nz = n
    lo, up = [], []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = i + nx * (j + ny * k)
                if i < nx - 1:
                    lo.append(c); up.append(c + 1)
                if j < ny - 1:
                    lo.append(c); up.append(c + nx)
                if k < nz - 1:
                    lo.append(c); up.append(c + nx * ny)
    return np.array(lo, dtype=np.int32), np.array(up, dtype=np.int32)


def random_ldu_addressing(n_cells, extra, seed):
    """A connected, irregular upper-triangular addressing: a chain plus random extra faces."""
    rng = np.random.default_rng(seed)
    pairs = {(c, c + 1) for c in range(n_cells - 1)}
    while len(pairs) < n_cells - 1 + extra:
        a, b = rng.integers(0, n_cells, 2)
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    pairs = sorted(pairs)
    lo = np.array([p[0] for p in pairs], dtype=np.int32)
    up = np.array([p[1] for p in pairs], dtype=np.int32)
    return lo, up


def spd_coefficients(n_cells, lo, up, seed=1234, asym=False):
    """upper = -U(0.5,1.5), diag = -sum(offdiag) + U(0.1,1), x = U(-1,1)."""
    rng = np.random.default_rng(seed)
    nf = len(lo)
    upper = -rng.uniform(0.5, 1.5, nf)
    lower = upper * rng.uniform(0.7, 1.3, nf) if asym else None
    diag = rng.uniform(0.1, 1.0, n_cells)
    np.subtract.at(diag, lo, upper)
    np.subtract.at(diag, up, lower if asym else upper)
    x = rng.uniform(-1.0, 1.0, n_cells)
    b = rng.uniform(-1.0, 1.0, n_cells)
    return diag, upper, lower, x, b


def dense_from_ldu(n_cells, lo, up, diag, upper, lower=None):
    A = np.diag(diag.astype(float))
    lw = upper if lower is None else lower
    for f in range(len(lo)):
        A[lo[f], up[f]] += upper[f]
        A[up[f], lo[f]] += lw[f]
    return A
