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
//   * forward writes F and re-arms B, backward reads F, re-arms it and writes B (polled by the neighbours) plus the
//     caller's result: F and B are all-sentinel again when the next application starts.
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

// Everything the sweeps touch apart from the caller's vectors lives in POSITION order (position = index in the level-ordered
// cell list): the dependency tables (neighbour positions, in the order the reference consumes them), the coefficients
// gathered next to them once per solve, rD, the forward result F and the backward result B the neighbours poll.  A warp's 32
// positions therefore read and write whole lines, and the polled entries -- positions in the neighbouring level -- are close
// to each other.  Only rA[cell], the result wA[cell] and the dot-product operand are accessed through the cell number.
struct FlowArgs {
    int n;                                   // cells
    const int* perm;                         // position -> cell
    const int *fStart, *fPos;                // lower-side faces of each position: [fStart[p], fStart[p+1]) neighbour positions, ascending face order
    const int *bStart, *bPos;                // owned faces: neighbour positions, DESCENDING face order (as the backward sweep consumes them)
    const double *coefF, *coefB;             // coefficients next to fPos / bPos: (DILU ? lower : upper)[face] / upper[face]
    const double* numF;                      // factor numerators next to fPos: upper*upper (DIC) or upper*lower (DILU)
    const double* diag;                      // natural order (factor)
    double* rD;                              // position order: raw recurrence values during the factor sweep, reciprocals afterwards
    const double* rA;                        // natural order
    double* F;                               // position order, all sentinel between applications
    double* B;                               // position order, all sentinel between applications' backward sweeps
    double* wA;                              // natural order result
    const double* dotVec;                    // natural order
    double* chunkPartial;                    // [nChunks] (DOT)
    int* ticket;
    int sleepNs;                             // back-off between two unsuccessful polls of a warp
    const SolveState* st;
};

// once per solve: coefficients from LDU face order into the position-ordered dependency tables
template <bool DILU>
__global__ void __launch_bounds__(BLK) k_flow_gather(int nFaces, const int* __restrict__ fFace, const int* __restrict__ bFace,
                                                     const double* __restrict__ upper, const double* __restrict__ lower,
                                                     double* __restrict__ coefF, double* __restrict__ coefB, double* __restrict__ numF,
                                                     const SolveState* st)
{
    if (solve_done(st)) return;
    for (int e = blockIdx.x * BLK + threadIdx.x; e < nFaces; e += gridDim.x * BLK) {
        const int f = fFace[e];
        const double u = upper[f];
        coefF[e] = DILU ? lower[f] : u;
        numF[e] = DILU ? u * lower[f] : u * u;
        coefB[e] = upper[bFace[e]];
    }
}

// WHAT: 0 = calcReciprocalD recurrence (raw values; the reciprocal pass follows), 1 = forward substitution, 2 = backward substitution
// A warp takes 32 consecutive positions per ticket.  Neighbour positions and coefficients are in registers before the first
// poll (FG faces at a time) and the pending neighbours of a group are polled with independent loads, so that once they are
// there a cell costs one L2 round trip (polling them one after the other, each behind its own index -> coefficient chain,
// cost ~4 us per level instead of ~1).
constexpr int FG = 4;

template <int WHAT, bool DOT>
__global__ void __launch_bounds__(BLK) k_sweep_flow(FlowArgs a)
{
    if (solve_done(a.st)) return;
    const double sent = __longlong_as_double((long long)WSENT);
    const int nChunks = (a.n + 31) / 32;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    double* out = WHAT == 0 ? a.rD : WHAT == 1 ? a.F : a.B;
    const int* start = WHAT == 2 ? a.bStart : a.fStart;
    const int* nbr = WHAT == 2 ? a.bPos : a.fPos;
    const double* cf = WHAT == 0 ? a.numF : WHAT == 1 ? a.coefF : a.coefB;
    for (;;) {
        int chunk = 0;
        if (lane == 0) chunk = atomicAdd(a.ticket, 1);
        chunk = __shfl_sync(FULL, chunk, 0);
        if (chunk >= nChunks) break;
        const int q0 = chunk * 32 + lane;
        const bool act = q0 < a.n;
        // forward sweeps walk the level order upwards, the backward sweep downwards
        const int pos = act ? (WHAT == 2 ? a.n - 1 - q0 : q0) : 0;
        const int c = act ? a.perm[pos] : 0;
        int e = 0, left = 0;                 // next entry of the position's dependency list and how many remain
        double w = 0, r = 0;
        if (act) {
            e = start[pos]; left = start[pos + 1] - e;
            if (WHAT == 0) w = a.diag[c];
            else if (WHAT == 1) { r = a.rD[pos]; w = r * a.rA[c]; a.B[pos] = sent; }
            else { r = a.rD[pos]; w = a.F[pos]; a.F[pos] = sent; }
        }
        int col[FG], gN = 0, k = 0;          // the group in registers: neighbour positions, coefficients, how many, how many consumed
        double coef[FG];
        auto load_group = [&]() {
            gN = min(FG, left); k = 0;
#pragma unroll
            for (int q = 0; q < FG; q++) {
                col[q] = 0; coef[q] = 0.0;
                if (q < gN) {
                    col[q] = nbr[e + q];
                    coef[q] = WHAT == 0 ? cf[e + q] : r * cf[e + q];
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
                    left -= gN; e += gN;
                    if (left == 0) { st_relaxed_f64(out + pos, w); done = true; }
                    else load_group();
                }
            }
        }
        if (WHAT == 2 && act) a.wA[c] = w;
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
