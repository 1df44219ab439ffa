// DIC/DILU over ARBITRARY LDU addressing without grid barriers ([UPSTREAM] DICPreconditioner / DILUPreconditioner:
// calcReciprocalD, precondition) -- what a decomposed or unstructured OpenFOAM case gets through the bPCG seam.
//
// The sweeps are recurrences over the lower-neighbour DAG.  k_sweep_general (ldu_kernels.cuh) runs them level by level
// with a grid barrier between levels: ~5 us per level, 7 ms per sweep for the 1430 levels of a 1000x400x32 block.
// Here the cells are still visited in level order, but nothing waits for a level:
//   * the result array is pre-filled with a sentinel (the all-ones NaN of the wave sweeps, which arithmetic never
//     produces) and a cell polls exactly the neighbour values it needs (ld.relaxed.gpu): the datum is its own flag, no
//     fence, no counter.  A cell consumes its neighbours in the reference's order and stops at the first one that is
//     not there yet, so the per-cell summation order -- hence every bit of the result -- is that of the face loop;
//   * chunks of 32 consecutive positions of the level-ordered cell list are handed out to warps by an atomic ticket.  A
//     cell's lower neighbours sit at earlier positions, i.e. in chunks that were claimed before -- by warps that are
//     running or done -- so every wait ends, for any grid size, without a cooperative launch;
//   * inside a warp a cell may depend on a cell of another lane (small levels).  Lanes therefore never spin on their
//     own: the warp loops `while (!all done) { every pending lane tries to advance }`, which makes progress whenever
//     any lane can;
//   * forward writes F (scratch) and re-arms the caller's result array, backward reads F, re-arms it and writes the
//     result: both arrays are all-sentinel again when the next application starts.
// The dot product fused into the backward sweep is summed per chunk (fixed tree) and then over the chunks by a fixed grid
// (k_sum), so it does not depend on which warp got which ticket.
#pragma once
#include "ldu_kernels.cuh"
#include "wave_kernels.cuh"

namespace b200 {

__device__ __forceinline__ void st_relaxed_f64(double* p, double v)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

struct FlowArgs {
    int n;                                   // cells
    const int *lvlCells, *rowStart, *rowMid, *rowCol, *rowFace;
    const double *diag, *upper, *lower;
    double* rD;                              // WHAT 0: raw recurrence values (sentinel-filled before); else the reciprocals
    const double* rA;
    double* F;                               // forward result (scratch, all sentinel between applications)
    double* wA;                              // result
    const double* dotVec;
    double* chunkPartial;                    // [nChunks] (DOT)
    int* ticket;
    int sleepNs;                             // back-off between two unsuccessful polls of a warp
    const SolveState* st;
};

// WHAT: 0 = calcReciprocalD recurrence (raw values; k_reciprocal follows), 1 = forward substitution, 2 = backward substitution
// A warp takes 32 consecutive positions per ticket.  Everything a cell needs apart from its neighbours' results -- the
// neighbour ids and the coefficients, FG faces at a time -- is in registers before the first poll, and the pending
// neighbours of a group are polled with independent loads, so that once they are there a cell costs one L2 round trip
// (polling them one after the other, each behind its own rowFace -> upper chain, cost ~4 us per level instead of ~1).
constexpr int FG = 4;

template <int WHAT, bool DILU, bool DOT>
__global__ void __launch_bounds__(BLK) k_sweep_flow(FlowArgs a)
{
    if (solve_done(a.st)) return;
    const double sent = __longlong_as_double((long long)WSENT);
    const int nChunks = (a.n + 31) / 32;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    double* out = WHAT == 0 ? a.rD : WHAT == 1 ? a.F : a.wA;
    for (;;) {
        int chunk = 0;
        if (lane == 0) chunk = atomicAdd(a.ticket, 1);
        chunk = __shfl_sync(FULL, chunk, 0);
        if (chunk >= nChunks) break;
        const int pos = chunk * 32 + lane;
        const bool act = pos < a.n;
        // forward sweeps walk the level order upwards, the backward sweep downwards
        const int c = act ? a.lvlCells[WHAT == 2 ? a.n - 1 - pos : pos] : 0;
        int e = 0, left = 0;                 // next face of the row (ascending, or descending for WHAT 2) and how many remain
        double w = 0, r = 0;
        if (act) {
            if (WHAT == 0) { w = a.diag[c]; e = a.rowStart[c]; left = a.rowMid[c] - e; }
            else if (WHAT == 1) { r = a.rD[c]; w = r * a.rA[c]; e = a.rowStart[c]; left = a.rowMid[c] - e; a.wA[c] = sent; }
            else { r = a.rD[c]; w = a.F[c]; a.F[c] = sent; e = a.rowStart[c + 1] - 1; left = e + 1 - a.rowMid[c]; }
        }
        int col[FG], gN = 0, k = 0;          // the group in registers: neighbour ids, coefficients, how many, how many consumed
        double coef[FG];
        auto load_group = [&]() {
            gN = min(FG, left); k = 0;
#pragma unroll
            for (int q = 0; q < FG; q++) {
                col[q] = 0; coef[q] = 0.0;
                if (q < gN) {
                    const int idx = WHAT == 2 ? e - q : e + q;
                    const int f = a.rowFace[idx];
                    col[q] = a.rowCol[idx];
                    if (WHAT == 0) coef[q] = DILU ? a.upper[f] * a.lower[f] : a.upper[f] * a.upper[f];
                    else coef[q] = r * ((WHAT == 1 && DILU) ? a.lower[f] : a.upper[f]);
                }
            }
        };
        if (act) load_group();
        bool done = !act;
        for (int round = 0; !__all_sync(FULL, done); round++) {
            if (round > 0 && a.sleepNs > 0) __nanosleep(a.sleepNs);
            if (!done) {
                double v[FG];
#pragma unroll
                for (int q = 0; q < FG; q++) { v[q] = sent; ld_relaxed_if(v[q], out + col[q], q >= k && q < gN); }
#pragma unroll
                for (int q = 0; q < FG; q++) {
                    if (q == k && q < gN && !is_sentinel(v[q])) {
                        if (WHAT == 0) w -= coef[q] / v[q];
                        else w -= coef[q] * v[q];
                        k++;
                    }
                }
                if (k == gN) {
                    left -= gN; e = WHAT == 2 ? e - gN : e + gN;
                    if (left == 0) { st_relaxed_f64(out + c, w); done = true; }
                    else load_group();
                }
            }
        }
        if (DOT) {
            const double part = warp_reduce<Sum>(act ? w * a.dotVec[c] : 0.0);
            if (lane == 0) a.chunkPartial[chunk] = part;
        }
    }
}

__global__ void __launch_bounds__(BLK) k_fill_sentinel(int64_t n, double* __restrict__ v, const SolveState* st)
{
    if (solve_done(st)) return;
    const double sent = __longlong_as_double((long long)WSENT);
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) v[c] = sent;
}

} // namespace b200
