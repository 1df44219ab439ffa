// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
// CPU restatement ("oracle") of the OpenFOAM-10 lduMatrix kernels that AdditiveFOAM's
// thermal solve reaches through `solve(...)` at
// applications/solvers/additiveFoam/thermo/TEqn.H:32-42.
//
// PARITY UNPINNED: OpenFOAM-10 is a third-party dependency of the reference that is not
// vendored under /root/reference (README.md:19 pins "OpenFOAM-10"; CI image
// openfoam/openfoam10-paraview510, .github/workflows/CI.yml:24) and is not installed in
// this environment.  The algorithms below restate the published upstream sources
// (paths under OpenFOAM-10/src/ are given per function); the reference ships no golden
// vectors for them, so they are anchored on derived known answers
// and dense-solve cross-checks in tests/ instead.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may use this code.
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>

namespace afo {

using scalar = double;
using label  = int32_t;

// [UPSTREAM] OpenFOAM/primitives/Scalar/doubleScalar: small, vSmall, great, vGreat
constexpr scalar small_  = 1e-15;
constexpr scalar vSmall  = 1e-300;
constexpr scalar great   = 1e15;
constexpr scalar vGreat  = 1e300;

// [UPSTREAM] lduAddressing: lowerAddr/upperAddr given; ownerStart and losort derived
// (OpenFOAM/matrices/lduMatrix/lduAddressing/lduAddressing.C calcLosort/calcOwnerStart).
struct LduAddr {
    label nCells = 0, nFaces = 0;
    std::vector<label> lower, upper, ownerStart, losort, losortStart;
    void finalize()
    {
        ownerStart.assign(nCells + 1, 0);
        for (label f = 0; f < nFaces; f++) ownerStart[lower[f] + 1]++;
        for (label c = 0; c < nCells; c++) ownerStart[c + 1] += ownerStart[c];
        // losort: faces sorted by upper address, stable in face index
        losortStart.assign(nCells + 1, 0);
        for (label f = 0; f < nFaces; f++) losortStart[upper[f] + 1]++;
        for (label c = 0; c < nCells; c++) losortStart[c + 1] += losortStart[c];
        losort.resize(nFaces);
        std::vector<label> pos(losortStart.begin(), losortStart.end() - 1);
        for (label f = 0; f < nFaces; f++) losort[pos[upper[f]]++] = f;
    }
};

// One coupled (processor) interface as the solver sees it:
// [UPSTREAM] lduInterfaceField / processorFvPatchField::updateInterfaceMatrix.
struct Interface {
    std::vector<label> faceCells;   // lduAddr().patchAddr(patchi)
    int neighbRank = -1;
    std::vector<label> nbrCells;    // cell index on the neighbour rank facing faceCells[i]
};

struct LduMatrix {
    const LduAddr* addr = nullptr;
    std::vector<scalar> diag, upper, lower;   // lower empty => symmetric (lower aliases upper)
    std::vector<std::vector<scalar>> bouCoeffs;   // per interface
    const std::vector<Interface>* interfaces = nullptr;
    const scalar* lowerPtr() const { return lower.empty() ? upper.data() : lower.data(); }
    bool asymmetric() const { return !lower.empty(); }
};

struct SolverControls {
    int solver = 0;           // 0 PCG, 1 PBiCGStab (B200_SOLVER_*)
    int precond = 2;          // B200_PRECOND_*
    scalar tolerance = 1e-6;  // [UPSTREAM] lduMatrix::solver::readControls defaults
    scalar relTol = 0;
    label minIter = 0;
    label maxIter = 1000;
};

struct SolverPerf {
    scalar initialResidual = 0, finalResidual = 0;
    label nIterations = 0;
    bool converged = false, singular = false;
    // [UPSTREAM] OpenFOAM/matrices/LduMatrix/LduMatrix/SolverPerformance.C
    bool checkConvergence(scalar tol, scalar relTol)
    {
        converged = (finalResidual < tol)
                 || (relTol > 1e-20 && finalResidual < relTol * initialResidual);
        return converged;
    }
    bool checkSingularity(scalar residual)
    {
        singular = !(residual > 1e-300);
        return singular;
    }
};

// ---------------------------------------------------------------------------------------
// Multi-rank containers.  R ranks are emulated in one process; "reductions" sum the rank
// partials in rank order, "halo exchange" reads the neighbour rank's array directly.
// R == 1 is the serial algorithm.
// ---------------------------------------------------------------------------------------
using Field  = std::vector<scalar>;
using Fields = std::vector<Field>;       // one per rank

template <class F> inline void forRanks(int R, F&& f)
{
#pragma omp parallel for schedule(static) if (R > 1)
    for (int r = 0; r < R; r++) f(r);
}

inline scalar gSum(const std::vector<scalar>& partial)
{
    scalar s = 0;
    for (scalar p : partial) s += p;
    return s;
}

// [UPSTREAM] OpenFOAM/matrices/lduMatrix/lduMatrix/lduMatrixATmul.C  lduMatrix::Amul
// (face loop; interface terms subtracted afterwards, processorFvPatchField::updateInterfaceMatrix)
inline void Amul(const LduMatrix& A, const Fields& psiAll, int r, Field& Apsi)
{
    const LduAddr& ad = *A.addr;
    const Field& psi = psiAll[r];
    const scalar* lo = A.lowerPtr();
    const scalar* up = A.upper.data();
    for (label c = 0; c < ad.nCells; c++) Apsi[c] = A.diag[c] * psi[c];
    for (label f = 0; f < ad.nFaces; f++) {
        Apsi[ad.upper[f]] += lo[f] * psi[ad.lower[f]];
        Apsi[ad.lower[f]] += up[f] * psi[ad.upper[f]];
    }
    if (A.interfaces) {
        for (size_t p = 0; p < A.interfaces->size(); p++) {
            const Interface& itf = (*A.interfaces)[p];
            const Field& pn = psiAll[itf.neighbRank];
            for (size_t i = 0; i < itf.faceCells.size(); i++)
                Apsi[itf.faceCells[i]] -= A.bouCoeffs[p][i] * pn[itf.nbrCells[i]];
        }
    }
}

// [UPSTREAM] lduMatrixOperations.C  lduMatrix::sumA
inline void sumA(const LduMatrix& A, Field& s)
{
    const LduAddr& ad = *A.addr;
    const scalar* lo = A.lowerPtr();
    const scalar* up = A.upper.data();
    for (label c = 0; c < ad.nCells; c++) s[c] = A.diag[c];
    for (label f = 0; f < ad.nFaces; f++) {
        s[ad.upper[f]] += lo[f];
        s[ad.lower[f]] += up[f];
    }
    if (A.interfaces)
        for (size_t p = 0; p < A.interfaces->size(); p++) {
            const Interface& itf = (*A.interfaces)[p];
            for (size_t i = 0; i < itf.faceCells.size(); i++) s[itf.faceCells[i]] -= A.bouCoeffs[p][i];
        }
}

// [UPSTREAM] lduMatrix/preconditioners/{DIC,DILU,diagonal,no}Preconditioner
struct Preconditioner {
    int kind = 0;
    const LduMatrix* A = nullptr;
    Field rD;
    void build(const LduMatrix& M, int k)
    {
        A = &M; kind = k;
        const LduAddr& ad = *M.addr;
        if (kind == 0) return;
        rD = M.diag;
        if (kind == 1) {                       // diagonalPreconditioner: rD = 1/diag
            for (label c = 0; c < ad.nCells; c++) rD[c] = 1.0 / rD[c];
            return;
        }
        const scalar* up = M.upper.data();
        const scalar* lo = M.lowerPtr();
        if (kind == 2)                         // DICPreconditioner::calcReciprocalD
            for (label f = 0; f < ad.nFaces; f++) rD[ad.upper[f]] -= up[f] * up[f] / rD[ad.lower[f]];
        else                                   // DILUPreconditioner::calcReciprocalD
            for (label f = 0; f < ad.nFaces; f++) rD[ad.upper[f]] -= up[f] * lo[f] / rD[ad.lower[f]];
        for (label c = 0; c < ad.nCells; c++) rD[c] = 1.0 / rD[c];
    }
    void precondition(Field& wA, const Field& rA) const
    {
        const LduAddr& ad = *A->addr;
        if (kind == 0) { wA = rA; return; }    // noPreconditioner
        for (label c = 0; c < ad.nCells; c++) wA[c] = rD[c] * rA[c];
        if (kind == 1) return;
        const scalar* up = A->upper.data();
        const scalar* lo = A->lowerPtr();
        if (kind == 2) {                       // DICPreconditioner::precondition
            for (label f = 0; f < ad.nFaces; f++)
                wA[ad.upper[f]] -= rD[ad.upper[f]] * up[f] * wA[ad.lower[f]];
        } else {                               // DILUPreconditioner::precondition (losort order)
            for (label f = 0; f < ad.nFaces; f++) {
                const label sf = ad.losort[f];
                wA[ad.upper[sf]] -= rD[ad.upper[sf]] * lo[sf] * wA[ad.lower[sf]];
            }
        }
        for (label f = ad.nFaces - 1; f >= 0; f--)
            wA[ad.lower[f]] -= rD[ad.lower[f]] * up[f] * wA[ad.upper[f]];
    }
};

// [UPSTREAM] lduMatrix/lduMatrix/lduMatrixSolver.C  lduMatrix::solver::normFactor
inline scalar normFactor(const std::vector<LduMatrix>& A, const Fields& psi, const Fields& source,
                         const Fields& Apsi, Fields& tmp)
{
    const int R = (int)A.size();
    std::vector<scalar> ps(R), pn(R);
    forRanks(R, [&](int r) {
        sumA(A[r], tmp[r]);
        scalar s = 0;
        for (scalar v : psi[r]) s += v;
        ps[r] = s; pn[r] = (scalar)psi[r].size();
    });
    const scalar avg = gSum(ps) / gSum(pn);                 // gAverage(psi)
    forRanks(R, [&](int r) {
        scalar s = 0;
        const label n = A[r].addr->nCells;
        for (label c = 0; c < n; c++) {
            const scalar t = tmp[r][c] * avg;
            s += std::fabs(Apsi[r][c] - t) + std::fabs(source[r][c] - t);
        }
        ps[r] = s;
    });
    return gSum(ps) + 1e-20;                                // solverPerformance::small_
}

inline scalar gSumProd(const Fields& a, const Fields& b)
{
    const int R = (int)a.size();
    std::vector<scalar> ps(R);
    forRanks(R, [&](int r) {
        scalar s = 0;
        for (size_t c = 0; c < a[r].size(); c++) s += a[r][c] * b[r][c];
        ps[r] = s;
    });
    return gSum(ps);
}
inline scalar gSumMag(const Fields& a)
{
    const int R = (int)a.size();
    std::vector<scalar> ps(R);
    forRanks(R, [&](int r) {
        scalar s = 0;
        for (scalar v : a[r]) s += std::fabs(v);
        ps[r] = s;
    });
    return gSum(ps);
}

struct ResidualLog { std::vector<scalar> res; };

// [UPSTREAM] lduMatrix/solvers/PCG/PCG.C  PCG::solve
inline SolverPerf solvePCG(const std::vector<LduMatrix>& A, Fields& psi, const Fields& source,
                           const SolverControls& sc, ResidualLog* log = nullptr)
{
    const int R = (int)A.size();
    SolverPerf perf;
    Fields pA(R), wA(R), rA(R);
    for (int r = 0; r < R; r++) { const size_t n = psi[r].size(); pA[r].assign(n, 0); wA[r].assign(n, 0); rA[r].assign(n, 0); }
    scalar wArA = 1e20, wArAold = wArA;                     // solverPerf.great_
    forRanks(R, [&](int r) { Amul(A[r], psi, r, wA[r]); });
    forRanks(R, [&](int r) { for (size_t c = 0; c < psi[r].size(); c++) rA[r][c] = source[r][c] - wA[r][c]; });
    const scalar nf = normFactor(A, psi, source, wA, pA);
    perf.initialResidual = gSumMag(rA) / nf;
    perf.finalResidual = perf.initialResidual;
    if (log) log->res.push_back(perf.initialResidual);
    if (sc.minIter > 0 || !perf.checkConvergence(sc.tolerance, sc.relTol)) {
        std::vector<Preconditioner> M(R);
        forRanks(R, [&](int r) { M[r].build(A[r], sc.precond); });
        do {
            wArAold = wArA;
            forRanks(R, [&](int r) { M[r].precondition(wA[r], rA[r]); });
            wArA = gSumProd(wA, rA);
            if (perf.nIterations == 0) {
                forRanks(R, [&](int r) { pA[r] = wA[r]; });
            } else {
                const scalar beta = wArA / wArAold;
                forRanks(R, [&](int r) { for (size_t c = 0; c < pA[r].size(); c++) pA[r][c] = wA[r][c] + beta * pA[r][c]; });
            }
            forRanks(R, [&](int r) { Amul(A[r], pA, r, wA[r]); });
            const scalar wApA = gSumProd(wA, pA);
            if (perf.checkSingularity(std::fabs(wApA) / nf)) break;
            const scalar alpha = wArA / wApA;
            forRanks(R, [&](int r) {
                for (size_t c = 0; c < pA[r].size(); c++) { psi[r][c] += alpha * pA[r][c]; rA[r][c] -= alpha * wA[r][c]; }
            });
            perf.finalResidual = gSumMag(rA) / nf;
            if (log) log->res.push_back(perf.finalResidual);
        } while ((perf.nIterations++ < sc.maxIter && !perf.checkConvergence(sc.tolerance, sc.relTol))
                 || perf.nIterations < sc.minIter);
    }
    return perf;
}

// [UPSTREAM] lduMatrix/solvers/PBiCGStab/PBiCGStab.C  PBiCGStab::solve
inline SolverPerf solvePBiCGStab(const std::vector<LduMatrix>& A, Fields& psi, const Fields& source,
                                 const SolverControls& sc, ResidualLog* log = nullptr)
{
    const int R = (int)A.size();
    SolverPerf perf;
    Fields yA(R), rA(R), pA(R);
    for (int r = 0; r < R; r++) { const size_t n = psi[r].size(); yA[r].assign(n, 0); rA[r].assign(n, 0); pA[r].assign(n, 0); }
    forRanks(R, [&](int r) { Amul(A[r], psi, r, yA[r]); });
    forRanks(R, [&](int r) { for (size_t c = 0; c < psi[r].size(); c++) rA[r][c] = source[r][c] - yA[r][c]; });
    const scalar nf = normFactor(A, psi, source, yA, pA);
    perf.initialResidual = gSumMag(rA) / nf;
    perf.finalResidual = perf.initialResidual;
    if (log) log->res.push_back(perf.initialResidual);
    if (sc.minIter > 0 || !perf.checkConvergence(sc.tolerance, sc.relTol)) {
        Fields AyA(R), sA(R), zA(R), tA(R);
        for (int r = 0; r < R; r++) { const size_t n = psi[r].size(); AyA[r].assign(n, 0); sA[r].assign(n, 0); zA[r].assign(n, 0); tA[r].assign(n, 0); }
        const Fields rA0(rA);
        scalar rA0rA = 0, alpha = 0, omega = 0;
        std::vector<Preconditioner> M(R);
        forRanks(R, [&](int r) { M[r].build(A[r], sc.precond); });
        do {
            const scalar rA0rAold = rA0rA;
            rA0rA = gSumProd(rA0, rA);
            if (perf.checkSingularity(std::fabs(rA0rA))) break;
            if (perf.nIterations == 0) {
                forRanks(R, [&](int r) { pA[r] = rA[r]; });
            } else {
                if (perf.checkSingularity(std::fabs(omega))) break;
                const scalar beta = (rA0rA / rA0rAold) * (alpha / omega);
                forRanks(R, [&](int r) {
                    for (size_t c = 0; c < pA[r].size(); c++) pA[r][c] = rA[r][c] + beta * (pA[r][c] - omega * AyA[r][c]);
                });
            }
            forRanks(R, [&](int r) { M[r].precondition(yA[r], pA[r]); });
            forRanks(R, [&](int r) { Amul(A[r], yA, r, AyA[r]); });
            const scalar rA0AyA = gSumProd(rA0, AyA);
            alpha = rA0rA / rA0AyA;
            forRanks(R, [&](int r) { for (size_t c = 0; c < sA[r].size(); c++) sA[r][c] = rA[r][c] - alpha * AyA[r][c]; });
            perf.finalResidual = gSumMag(sA) / nf;
            if (perf.checkConvergence(sc.tolerance, sc.relTol)) {
                forRanks(R, [&](int r) { for (size_t c = 0; c < psi[r].size(); c++) psi[r][c] += alpha * yA[r][c]; });
                perf.nIterations++;
                if (log) log->res.push_back(perf.finalResidual);
                return perf;
            }
            forRanks(R, [&](int r) { M[r].precondition(zA[r], sA[r]); });
            forRanks(R, [&](int r) { Amul(A[r], zA, r, tA[r]); });
            const scalar tAtA = gSumProd(tA, tA);          // gSumSqr
            omega = gSumProd(tA, sA) / tAtA;
            forRanks(R, [&](int r) {
                for (size_t c = 0; c < psi[r].size(); c++) {
                    psi[r][c] += alpha * yA[r][c] + omega * zA[r][c];
                    rA[r][c] = sA[r][c] - omega * tA[r][c];
                }
            });
            perf.finalResidual = gSumMag(rA) / nf;
            if (log) log->res.push_back(perf.finalResidual);
        } while ((perf.nIterations++ < sc.maxIter && !perf.checkConvergence(sc.tolerance, sc.relTol))
                 || perf.nIterations < sc.minIter);
    }
    return perf;
}

// [UPSTREAM] lduMatrix/solvers/diagonalSolver/diagonalSolver.C : psi = source/diag, 0 iterations
inline SolverPerf solveDiagonal(const std::vector<LduMatrix>& A, Fields& psi, const Fields& source)
{
    const int R = (int)A.size();
    forRanks(R, [&](int r) { for (size_t c = 0; c < psi[r].size(); c++) psi[r][c] = source[r][c] / A[r].diag[c]; });
    SolverPerf perf; perf.converged = true;
    return perf;
}

// Level of each cell in the lower-neighbour DAG: level = 1 + max level of
// cells reached through faces where this cell is the upper address; 0 with none.
inline label levelSchedule(const LduAddr& ad, std::vector<label>& level)
{
    level.assign(ad.nCells, 0);
    label nLevels = ad.nCells ? 1 : 0;
    for (label f = 0; f < ad.nFaces; f++) {      // upper-triangular order => lower[f] is final
        const label l = level[ad.lower[f]] + 1;
        if (l > level[ad.upper[f]]) level[ad.upper[f]] = l;
        if (l + 1 > nLevels) nLevels = l + 1;
    }
    return nLevels;
}

} // namespace afo
