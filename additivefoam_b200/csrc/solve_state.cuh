// Scalar state of a running Krylov solve, resident on the device, and the one-thread control steps that advance it
// ([UPSTREAM] PCG.C / PBiCGStab.C scalar recurrences, SolverPerformance::checkConvergence).
// The control step that follows a global sum runs in the epilogue of the kernel that completes the sum -- the last block of a
// reducing kernel on one rank (common.cuh: grid_reduce), the peer all-reduce kernel on several (ctx.cu) -- instead of in a
// launch of its own; k_ctrl (ldu_kernels.cuh) remains for the NCCL fallback.
#pragma once

namespace b200 {

struct SolveState {
    double tol, relTol, nGlobal;
    int minIter, maxIter;
    double avgPsi, normFactor, initRes, finalRes;
    double wArA, wArAold, wApA, alpha, beta, omega, rA0rA, rA0rAold, rA0AyA;
    int nIter, done, converged, singular, early;
};

// reduction slots in RedWork::result
enum { S_SUMPSI = 0, S_NORM = 1, S_MAGR0 = 2, S_RHO = 3, S_DEN = 4, S_MAGR = 5, S_TT = 6, S_TS = 7, S_MAGR2 = 8, S_NSLOTS = 16 };

enum { C_NONE = -1, C_AVG = 0, C_INIT, C_PCG_BETA, C_PCG_ALPHA, C_ITER_END, C_BICG_RHO, C_BICG_ALPHA, C_BICG_SCHECK, C_BICG_EARLY_END,
       C_BICG_OMEGA, C_ITER_END2 };

__device__ __forceinline__ bool check_convergence(SolveState* s)
{
    // [UPSTREAM] SolverPerformance::checkConvergence
    s->converged = (s->finalRes < s->tol) || (s->relTol > 1e-20 && s->finalRes < s->relTol * s->initRes);
    return s->converged != 0;
}
__device__ __forceinline__ void iter_end(SolveState* s, double sumMag)
{
    s->finalRes = sumMag / s->normFactor;
    // do { ... } while ((nIterations()++ < maxIter_ && !checkConvergence(...)) || nIterations() < minIter_)
    const bool c1 = s->nIter < s->maxIter;
    s->nIter++;
    bool cont = false;
    if (c1) cont = !check_convergence(s);
    cont = cont || s->nIter < s->minIter;
    if (!cont) s->done = 1;
}

// one thread; `red` = RedWork::result (already all-reduced over the ranks)
__device__ inline void ctrl_apply(int phase, SolveState* s, const double* red)
{
    if (s->done) return;
    int* early = &s->early;
    switch (phase) {
    case C_AVG: s->avgPsi = red[S_SUMPSI] / s->nGlobal; break;
    case C_INIT:
        s->normFactor = red[S_NORM] + 1e-20;
        s->initRes = red[S_MAGR0] / s->normFactor;
        s->finalRes = s->initRes;
        s->wArA = 1e20; s->nIter = 0; s->rA0rA = 0; s->alpha = 0; s->omega = 0;
        if (!(s->minIter > 0 || !check_convergence(s))) s->done = 1;
        break;
    case C_PCG_BETA:
        s->wArAold = s->wArA; s->wArA = red[S_RHO];
        s->beta = s->wArA / s->wArAold;
        break;
    case C_PCG_ALPHA:
        s->wApA = red[S_DEN];
        if (!(fabs(s->wApA) / s->normFactor > 1e-300)) { s->singular = 1; s->done = 1; }
        else { s->singular = 0; s->alpha = s->wArA / s->wApA; }
        break;
    case C_ITER_END: iter_end(s, red[S_MAGR]); break;
    case C_ITER_END2: iter_end(s, red[S_MAGR2]); break;
    case C_BICG_RHO:
        s->rA0rAold = s->rA0rA; s->rA0rA = red[S_RHO];
        if (!(fabs(s->rA0rA) > 1e-300)) { s->singular = 1; s->done = 1; break; }
        s->singular = 0;
        if (s->nIter > 0) {
            if (!(fabs(s->omega) > 1e-300)) { s->singular = 1; s->done = 1; break; }
            s->beta = (s->rA0rA / s->rA0rAold) * (s->alpha / s->omega);
        }
        break;
    case C_BICG_ALPHA: s->rA0AyA = red[S_DEN]; s->alpha = s->rA0rA / s->rA0AyA; break;
    case C_BICG_SCHECK:
        s->finalRes = red[S_MAGR] / s->normFactor;
        *early = check_convergence(s) ? 1 : 0;
        break;
    case C_BICG_EARLY_END:
        if (*early) { s->nIter++; s->done = 1; }
        break;
    case C_BICG_OMEGA: s->omega = red[S_TS] / red[S_TT]; break;
    }
}

} // namespace b200
