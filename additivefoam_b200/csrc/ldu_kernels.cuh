// Device kernels of the LDU path (see ldu.cuh for the layout and what each replaces).
#pragma once
#include "ldu.cuh"
#include "block_addr.cuh"
#include <cooperative_groups.h>

namespace b200 {
namespace cg = cooperative_groups;



// fused-reduction modes of the Amul kernels
enum { AM_NONE = 0, AM_SUMX = 1, AM_DOT_YX = 2, AM_DOT_AUXY = 3, AM_TT_TS = 4 };

__device__ __forceinline__ bool solve_done(const SolveState* st) { return st != nullptr && st->done != 0; }

template <int MODE>
__device__ __forceinline__ void amul_accumulate(double y, double x, const double* aux, int64_t c, double& a0, double& a1)
{
    if (MODE == AM_SUMX) a0 += x;
    else if (MODE == AM_DOT_YX) a0 += y * x;
    else if (MODE == AM_DOT_AUXY) a0 += aux[c] * y;
    else if (MODE == AM_TT_TS) { a0 += y * y; a1 += y * aux[c]; }
}

template <int MODE> __device__ __forceinline__ void amul_reduce(double a0, double a1, RedWork red)
{
    __shared__ double smem[32];
    if (MODE == AM_SUMX) { double v[1] = {a0}; const int ops[1] = {0}; grid_reduce<1>(v, ops, red, S_SUMPSI, smem); }
    else if (MODE == AM_DOT_YX || MODE == AM_DOT_AUXY) { double v[1] = {a0}; const int ops[1] = {0}; grid_reduce<1>(v, ops, red, S_DEN, smem); }
    else if (MODE == AM_TT_TS) { double v[2] = {a0, a1}; const int ops[2] = {0, 0}; grid_reduce<2>(v, ops, red, S_TT, smem); }
}

// ------------------------------------------------------------------ general addressing
// [UPSTREAM] lduMatrix::Amul as a gather with the face loop's per-cell order.
template <int MODE>
__global__ void __launch_bounds__(BLK) k_amul_general(int n, const int* __restrict__ rowStart, const int* __restrict__ rowMid,
                                                      const int* __restrict__ rowCol, const int* __restrict__ rowFace,
                                                      const double* __restrict__ diag, const double* __restrict__ upper,
                                                      const double* __restrict__ lower, const double* __restrict__ x,
                                                      double* __restrict__ y, const double* __restrict__ aux, RedWork red,
                                                      const SolveState* st)
{
    if (solve_done(st)) return;
    double a0 = 0, a1 = 0;
    for (int c = blockIdx.x * BLK + threadIdx.x; c < n; c += gridDim.x * BLK) {
        const double xc = x[c];
        double s = diag[c] * xc;
        int e = rowStart[c];
        const int m = rowMid[c], end = rowStart[c + 1];
        for (; e < m; e++) s += __ldg(lower + rowFace[e]) * x[rowCol[e]];
        for (; e < end; e++) s += __ldg(upper + rowFace[e]) * x[rowCol[e]];
        y[c] = s;
        amul_accumulate<MODE>(s, xc, aux, c, a0, a1);
    }
    if (MODE != AM_NONE) amul_reduce<MODE>(a0, a1, red);
}

// processorFvPatchField::updateInterfaceMatrix: result[faceCells[i]] -= coeffs[i]*pnf[i]
__global__ void k_iface_apply(int n, const int* __restrict__ faceCells, const double* __restrict__ coeffs,
                              const double* __restrict__ pnf, double* __restrict__ y, int sequential, const SolveState* st)
{
    if (solve_done(st)) return;
    if (sequential) {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (int i = 0; i < n; i++) y[faceCells[i]] -= coeffs[i] * pnf[i];
        return;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[faceCells[i]] -= coeffs[i] * pnf[i];
}
__global__ void k_iface_gather(int n, const int* __restrict__ faceCells, const double* __restrict__ x, double* __restrict__ buf)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = x[faceCells[i]];
}
__global__ void k_iface_sumA(int n, const int* __restrict__ faceCells, const double* __restrict__ coeffs, double* __restrict__ s,
                             int sequential)
{
    if (sequential) {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (int i = 0; i < n; i++) s[faceCells[i]] -= coeffs[i];
        return;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s[faceCells[i]] -= coeffs[i];
}

// [UPSTREAM] lduMatrix::sumA (without interfaces)
__global__ void __launch_bounds__(BLK) k_sumA_general(int n, const int* __restrict__ rowStart, const int* __restrict__ rowMid,
                                                      const int* __restrict__ rowFace, const double* __restrict__ diag,
                                                      const double* __restrict__ upper, const double* __restrict__ lower,
                                                      double* __restrict__ out)
{
    for (int c = blockIdx.x * BLK + threadIdx.x; c < n; c += gridDim.x * BLK) {
        double s = diag[c];
        int e = rowStart[c];
        const int m = rowMid[c], end = rowStart[c + 1];
        for (; e < m; e++) s += lower[rowFace[e]];
        for (; e < end; e++) s += upper[rowFace[e]];
        out[c] = s;
    }
}

// Level-scheduled sweeps over arbitrary addressing.  One cooperative launch; cells of a
// level are independent, levels are separated by a grid barrier.
// what: 0 = calcReciprocalD recurrence (raw rD), 1 = forward substitution, 2 = backward substitution (+ dot)
template <int WHAT, bool DILU, bool DOT>
__global__ void __launch_bounds__(BLK) k_sweep_general(int nLevels, const int* __restrict__ lvlStart, const int* __restrict__ lvlCells,
                                                       const int* __restrict__ rowStart, const int* __restrict__ rowMid,
                                                       const int* __restrict__ rowCol, const int* __restrict__ rowFace,
                                                       const double* __restrict__ diag, const double* __restrict__ upper,
                                                       const double* __restrict__ lower, double* rD, const double* __restrict__ rA,
                                                       double* wA, const double* __restrict__ dotVec, RedWork red, int slot,
                                                       const SolveState* st)
{
    if (solve_done(st)) return;
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * BLK + threadIdx.x, nthr = gridDim.x * BLK;
    double acc = 0;
    for (int q = 0; q < nLevels; q++) {
        const int L = WHAT == 2 ? nLevels - 1 - q : q;
        const int beg = lvlStart[L], end = lvlStart[L + 1];
        for (int idx = beg + tid; idx < end; idx += nthr) {
            const int c = lvlCells[idx];
            if (WHAT == 0) {
                double d = diag[c];
                for (int e = rowStart[c]; e < rowMid[c]; e++) {
                    const int f = rowFace[e];
                    const double num = DILU ? upper[f] * lower[f] : upper[f] * upper[f];
                    d -= num / __ldcg(rD + rowCol[e]);
                }
                rD[c] = d;
            } else if (WHAT == 1) {
                const double r = rD[c];
                double w = r * rA[c];
                for (int e = rowStart[c]; e < rowMid[c]; e++) {
                    const int f = rowFace[e];
                    w -= r * (DILU ? lower[f] : upper[f]) * __ldcg(wA + rowCol[e]);
                }
                wA[c] = w;
            } else {
                const double r = rD[c];
                double w = __ldcg(wA + c);
                for (int e = rowStart[c + 1] - 1; e >= rowMid[c]; e--) w -= r * upper[rowFace[e]] * __ldcg(wA + rowCol[e]);
                wA[c] = w;
                if (DOT) acc += w * dotVec[c];
            }
        }
        if (q + 1 < nLevels) grid.sync();
    }
    if (DOT) {
        __shared__ double smem[32];
        double v[1] = {acc}; const int ops[1] = {0};
        grid_reduce<1>(v, ops, red, slot, smem);
    }
}

// ------------------------------------------------------------------ block addressing (helpers in block_addr.cuh)
template <int MODE>
__global__ void __launch_bounds__(BLK) k_amul_block(BlockDims d, const double* __restrict__ diag, const double* __restrict__ upper,
                                                    const double* __restrict__ lower, const double* __restrict__ x,
                                                    double* __restrict__ y, const double* __restrict__ aux,
                                                    const double* __restrict__ bouBelow, const double* __restrict__ bouAbove,
                                                    RedWork red, const SolveState* st)
{
    if (solve_done(st)) return;
    const int nx = d.nx, nxny = d.nx * d.ny;
    const int segs = (nx + BLK - 1) / BLK, nItems = segs * d.ny * d.nz;
    double a0 = 0, a1 = 0;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int row = item / segs, seg = item - row * segs;
        const int i = seg * BLK + threadIdx.x;
        if (i >= nx) continue;
        const RowCtx r = row_ctx(d, row);
        int fzl, fyl, fxl, fx, fy, fz;
        cell_faces(d, r, i, fzl, fyl, fxl, fx, fy, fz);
        const int64_t c = (int64_t)row * nx + i;
        const double xc = x[c];
        double s = diag[c] * xc;
        if (fzl >= 0) s += lower[fzl] * x[c - nxny];
        if (fyl >= 0) s += lower[fyl] * x[c - nx];
        if (fxl >= 0) s += lower[fxl] * x[c - 1];
        if (fx >= 0) s += upper[fx] * x[c + 1];
        if (fy >= 0) s += upper[fy] * x[c + nx];
        if (fz >= 0) s += upper[fz] * x[c + nxny];
        if (bouBelow != nullptr && r.k == 0) s -= bouBelow[i + nx * r.j] * x[c - nxny];
        if (bouAbove != nullptr && r.k == d.nz - 1) s -= bouAbove[i + nx * r.j] * x[c + nxny];
        y[c] = s;
        amul_accumulate<MODE>(s, xc, aux, c, a0, a1);
    }
    if (MODE != AM_NONE) amul_reduce<MODE>(a0, a1, red);
}

// rA = source - wA, sumA gather, normFactor and initial residual sums
// ([UPSTREAM] lduMatrix::solver::normFactor + PCG.C / PBiCGStab.C set-up)
__global__ void __launch_bounds__(BLK) k_normfactor_block(BlockDims d, const double* __restrict__ diag, const double* __restrict__ upper,
                                                          const double* __restrict__ lower, const double* __restrict__ bouBelow,
                                                          const double* __restrict__ bouAbove, const double* __restrict__ source,
                                                          const double* __restrict__ wA, double* __restrict__ rA, RedWork red,
                                                          const SolveState* st)
{
    const int nx = d.nx;
    const int segs = (nx + BLK - 1) / BLK, nItems = segs * d.ny * d.nz;
    const double avg = st->avgPsi;
    double a0 = 0, a1 = 0;
    for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        const int row = item / segs, seg = item - row * segs;
        const int i = seg * BLK + threadIdx.x;
        if (i >= nx) continue;
        const RowCtx r = row_ctx(d, row);
        int fzl, fyl, fxl, fx, fy, fz;
        cell_faces(d, r, i, fzl, fyl, fxl, fx, fy, fz);
        const int64_t c = (int64_t)row * nx + i;
        double s = diag[c];
        if (fzl >= 0) s += lower[fzl];
        if (fyl >= 0) s += lower[fyl];
        if (fxl >= 0) s += lower[fxl];
        if (fx >= 0) s += upper[fx];
        if (fy >= 0) s += upper[fy];
        if (fz >= 0) s += upper[fz];
        if (bouBelow != nullptr && r.k == 0) s -= bouBelow[i + nx * r.j];
        if (bouAbove != nullptr && r.k == d.nz - 1) s -= bouAbove[i + nx * r.j];
        const double t = s * avg, w = wA[c], b = source[c];
        const double res = b - w;
        rA[c] = res;
        a0 += fabs(w - t) + fabs(b - t);
        a1 += fabs(res);
    }
    __shared__ double smem[32];
    double v[2] = {a0, a1}; const int ops[2] = {0, 0};
    grid_reduce<2>(v, ops, red, S_NORM, smem);
}

__global__ void __launch_bounds__(BLK) k_normfactor_general(int n, const double* __restrict__ sumA, const double* __restrict__ source,
                                                            const double* __restrict__ wA, double* __restrict__ rA, RedWork red,
                                                            const SolveState* st)
{
    const double avg = st->avgPsi;
    double a0 = 0, a1 = 0;
    for (int c = blockIdx.x * BLK + threadIdx.x; c < n; c += gridDim.x * BLK) {
        const double t = sumA[c] * avg, w = wA[c], b = source[c];
        const double res = b - w;
        rA[c] = res;
        a0 += fabs(w - t) + fabs(b - t);
        a1 += fabs(res);
    }
    __shared__ double smem[32];
    double v[2] = {a0, a1}; const int ops[2] = {0, 0};
    grid_reduce<2>(v, ops, red, S_NORM, smem);
}

// Pencil-marching sweeps on the block (v1): one thread per x-pencil (j,k), hyperplane i+j+k = L per step,
// grid barrier between steps.  WHAT as in k_sweep_general.
template <int WHAT, bool DILU, bool DOT>
__global__ void __launch_bounds__(BLK) k_sweep_block(BlockDims d, const double* __restrict__ diag, const double* __restrict__ upper,
                                                     const double* __restrict__ lower, double* rD, const double* __restrict__ rA,
                                                     double* wA, const double* __restrict__ dotVec, RedWork red, int slot,
                                                     const SolveState* st)
{
    if (solve_done(st)) return;
    cg::grid_group grid = cg::this_grid();
    const int nx = d.nx, nxny = d.nx * d.ny, nPencils = d.ny * d.nz;
    const int tid = blockIdx.x * BLK + threadIdx.x, nthr = gridDim.x * BLK;
    const int nLevels = d.nx + d.ny + d.nz - 2;
    double acc = 0;
    for (int q = 0; q < nLevels; q++) {
        const int L = WHAT == 2 ? nLevels - 1 - q : q;
        for (int p = tid; p < nPencils; p += nthr) {
            const int k = p / d.ny, j = p - k * d.ny;
            const int i = L - j - k;
            if (i < 0 || i >= nx) continue;
            const RowCtx r = row_ctx(d, p);
            int fzl, fyl, fxl, fx, fy, fz;
            cell_faces(d, r, i, fzl, fyl, fxl, fx, fy, fz);
            const int64_t c = (int64_t)p * nx + i;
            if (WHAT == 0) {
                double v = diag[c];
                if (fzl >= 0) v -= (DILU ? upper[fzl] * lower[fzl] : upper[fzl] * upper[fzl]) / __ldcg(rD + c - nxny);
                if (fyl >= 0) v -= (DILU ? upper[fyl] * lower[fyl] : upper[fyl] * upper[fyl]) / __ldcg(rD + c - nx);
                if (fxl >= 0) v -= (DILU ? upper[fxl] * lower[fxl] : upper[fxl] * upper[fxl]) / __ldcg(rD + c - 1);
                rD[c] = v;
            } else if (WHAT == 1) {
                const double rd = rD[c];
                const double* co = DILU ? lower : upper;
                double w = rd * rA[c];
                if (fzl >= 0) w -= rd * co[fzl] * __ldcg(wA + c - nxny);
                if (fyl >= 0) w -= rd * co[fyl] * __ldcg(wA + c - nx);
                if (fxl >= 0) w -= rd * co[fxl] * __ldcg(wA + c - 1);
                wA[c] = w;
            } else {
                const double rd = rD[c];
                double w = __ldcg(wA + c);
                if (fz >= 0) w -= rd * upper[fz] * __ldcg(wA + c + nxny);
                if (fy >= 0) w -= rd * upper[fy] * __ldcg(wA + c + nx);
                if (fx >= 0) w -= rd * upper[fx] * __ldcg(wA + c + 1);
                wA[c] = w;
                if (DOT) acc += w * dotVec[c];
            }
        }
        if (q + 1 < nLevels) grid.sync();
    }
    if (DOT) {
        __shared__ double smem[32];
        double v[1] = {acc}; const int ops[1] = {0};
        grid_reduce<1>(v, ops, red, slot, smem);
    }
}

// ------------------------------------------------------------------ vector kernels (any addressing)
__global__ void __launch_bounds__(BLK) k_reciprocal(int64_t n, double* __restrict__ v, const SolveState* st)
{
    if (solve_done(st)) return;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) v[c] = 1.0 / v[c];
}
// diagonalPreconditioner / noPreconditioner application with the fused wA.rA sum
template <bool DIAG>
__global__ void __launch_bounds__(BLK) k_precond_pointwise(int64_t n, const double* __restrict__ rD, const double* __restrict__ rA,
                                                           double* __restrict__ wA, const double* __restrict__ dotVec, RedWork red,
                                                           int slot, int doDot, const SolveState* st)
{
    if (solve_done(st)) return;
    double acc = 0;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) {
        const double w = DIAG ? rD[c] * rA[c] : rA[c];
        wA[c] = w;
        if (doDot) acc += w * dotVec[c];
    }
    if (doDot) {
        __shared__ double smem[32];
        double v[1] = {acc}; const int ops[1] = {0};
        grid_reduce<1>(v, ops, red, slot, smem);
    }
}
__global__ void __launch_bounds__(BLK) k_dot(int64_t n, const double* __restrict__ a, const double* __restrict__ b, RedWork red, int slot,
                                            const SolveState* st)
{
    if (solve_done(st)) return;
    double acc = 0;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) acc += a[c] * b[c];
    __shared__ double smem[32];
    double v[1] = {acc}; const int ops[1] = {0};
    grid_reduce<1>(v, ops, red, slot, smem);
}
__global__ void __launch_bounds__(BLK) k_sum(int64_t n, const double* __restrict__ a, RedWork red, int slot, const SolveState* st = nullptr)
{
    if (solve_done(st)) return;
    double acc = 0;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) acc += a[c];
    __shared__ double smem[32];
    double v[1] = {acc}; const int ops[1] = {0};
    grid_reduce<1>(v, ops, red, slot, smem);
}
// PCG: pA = wA (first iteration) or wA + beta*pA
__global__ void __launch_bounds__(BLK) k_pcg_p(int64_t n, const double* __restrict__ wA, double* __restrict__ pA, const SolveState* st)
{
    if (solve_done(st)) return;
    const bool first = st->nIter == 0;
    const double beta = st->beta;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK)
        pA[c] = first ? wA[c] : wA[c] + beta * pA[c];
}
// PCG: psi += alpha*pA; rA -= alpha*wA; sum|rA|
__global__ void __launch_bounds__(BLK) k_pcg_update(int64_t n, const double* __restrict__ pA, const double* __restrict__ q,
                                                   double* __restrict__ psi, double* __restrict__ rA, RedWork red, const SolveState* st)
{
    if (solve_done(st)) return;
    const double alpha = st->alpha;
    double acc = 0;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) {
        psi[c] += alpha * pA[c];
        const double r = rA[c] - alpha * q[c];
        rA[c] = r;
        acc += fabs(r);
    }
    __shared__ double smem[32];
    double v[1] = {acc}; const int ops[1] = {0};
    grid_reduce<1>(v, ops, red, S_MAGR, smem);
}
// PBiCGStab vector updates
__global__ void __launch_bounds__(BLK) k_bicg_p(int64_t n, const double* __restrict__ rA, const double* __restrict__ AyA,
                                               double* __restrict__ pA, const SolveState* st)
{
    if (solve_done(st)) return;
    const bool first = st->nIter == 0;
    const double beta = st->beta, omega = st->omega;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK)
        pA[c] = first ? rA[c] : rA[c] + beta * (pA[c] - omega * AyA[c]);
}
__global__ void __launch_bounds__(BLK) k_bicg_s(int64_t n, const double* __restrict__ rA, const double* __restrict__ AyA,
                                               double* __restrict__ sA, RedWork red, const SolveState* st)
{
    if (solve_done(st)) return;
    const double alpha = st->alpha;
    double acc = 0;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) {
        const double s = rA[c] - alpha * AyA[c];
        sA[c] = s;
        acc += fabs(s);
    }
    __shared__ double smem[32];
    double v[1] = {acc}; const int ops[1] = {0};
    grid_reduce<1>(v, ops, red, S_MAGR, smem);
}
__global__ void __launch_bounds__(BLK) k_bicg_early(int64_t n, const double* __restrict__ yA, double* __restrict__ psi, const SolveState* st,
                                                   const int* early)
{
    if (solve_done(st) || !*early) return;
    const double alpha = st->alpha;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) psi[c] += alpha * yA[c];
}
__global__ void __launch_bounds__(BLK) k_bicg_update(int64_t n, const double* __restrict__ yA, const double* __restrict__ zA,
                                                    const double* __restrict__ sA, const double* __restrict__ tA,
                                                    double* __restrict__ psi, double* __restrict__ rA, RedWork red, const SolveState* st)
{
    if (solve_done(st)) return;
    const double alpha = st->alpha, omega = st->omega;
    double acc = 0;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) {
        psi[c] += alpha * yA[c] + omega * zA[c];
        const double r = sA[c] - omega * tA[c];
        rA[c] = r;
        acc += fabs(r);
    }
    __shared__ double smem[32];
    double v[1] = {acc}; const int ops[1] = {0};
    grid_reduce<1>(v, ops, red, S_MAGR2, smem);
}
__global__ void __launch_bounds__(BLK) k_copy(int64_t n, const double* __restrict__ a, double* __restrict__ b, const SolveState* st)
{
    if (solve_done(st)) return;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) b[c] = a[c];
}
// [UPSTREAM] diagonalSolver::solve
__global__ void __launch_bounds__(BLK) k_diag_solve(int64_t n, const double* __restrict__ diag, const double* __restrict__ source,
                                                   double* __restrict__ psi)
{
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) psi[c] = source[c] / diag[c];
}

// ------------------------------------------------------------------ scalar control (one thread; solve_state.cuh)
__global__ void k_ctrl(int phase, SolveState* s, const double* __restrict__ red)
{
    ctrl_apply(phase, s, red);
}

// block addressing check / expansion used by setBlock and the probes
__global__ void k_block_addressing(BlockDims d, int* __restrict__ lower, int* __restrict__ upper)
{
    const int nItems = d.ny * d.nz;
    for (int row = blockIdx.x; row < nItems; row += gridDim.x) {
        const RowCtx r = row_ctx(d, row);
        for (int i = threadIdx.x; i < d.nx; i += blockDim.x) {
            int fzl, fyl, fxl, fx, fy, fz;
            cell_faces(d, r, i, fzl, fyl, fxl, fx, fy, fz);
            const int c = row * d.nx + i;
            if (fx >= 0) { lower[fx] = c; upper[fx] = c + 1; }
            if (fy >= 0) { lower[fy] = c; upper[fy] = c + d.nx; }
            if (fz >= 0) { lower[fz] = c; upper[fz] = c + d.nx * d.ny; }
        }
    }
}

} // namespace b200
