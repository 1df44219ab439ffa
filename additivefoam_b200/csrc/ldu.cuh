// LDU matrix kernels and Krylov solvers on the device.
// Replaces [UPSTREAM OpenFOAM-10] lduMatrix::Amul / sumA (lduMatrixATmul.C, lduMatrixOperations.C),
// PCG.C, PBiCGStab.C, diagonalSolver.C, DIC/DILU/diagonalPreconditioner.C, reached from
// applications/solvers/additiveFoam/thermo/TEqn.H:32-42.
//
// Data layout: OpenFOAM's own -- diag[nCells], upper[nFaces], lower[nFaces] in face order,
// lowerAddr/upperAddr in upper-triangular order.  Two addressing back-ends:
//   * General  : arbitrary LDU addressing.  A cell-centred row table (CSR of faces: first the
//                faces where the cell is the upper address, then the faces it owns, both
//                ascending) turns the reference's face-loop scatter into a gather with the same
//                per-cell summation order; DIC/DILU sweeps run level by level over the
//                level-renumbered cell list.
//   * Block    : a single blockMesh hex block.  Neighbour cells and face ids are closed-form,
//                so no addressing is read at all; sweeps run along x-pencils.
#pragma once
#include "common.cuh"
#include <vector>

namespace b200 {

struct Ctx;

// closed-form addressing of one hex block
struct BlockDims {
    int nx = 0, ny = 0, nz = 0;
    __host__ __device__ int64_t nCells() const { return (int64_t)nx * ny * nz; }
    __host__ __device__ int64_t nFaces() const
    {
        return (int64_t)(nx - 1) * ny * nz + (int64_t)nx * (ny - 1) * nz + (int64_t)nx * ny * (nz - 1);
    }
    // number of internal faces owned by cells that precede row (j,k)
    __host__ __device__ int rowFaceBase(int j, int k) const
    {
        return (k * ny + j) * (nx - 1) + (k * (ny - 1) * nx + j * nx) + (k * nx * ny + (k < nz - 1 ? j * nx : 0));
    }
};

// coupled interface (processor patch) as the solver sees it
struct DevInterface {
    int size = 0;
    int neighbRank = -1;
    int* faceCells = nullptr;      // device
    double* recvBuf = nullptr;     // device: neighbour psi at the interface cells
    double* sendBuf = nullptr;     // device
    bool hasDuplicates = false;    // a cell appears twice in faceCells: apply sequentially
};

struct MatrixView {
    const double* diag = nullptr;
    const double* upper = nullptr;
    const double* lower = nullptr;                 // == upper when symmetric
    const double* const* bouCoeffs = nullptr;      // host array of device pointers, one per interface
    // >= 0: the caller promises that upper/lower are unchanged since its last solve with the same value (the resident
    // stepper assembles the off-diagonals once per time step and solves once per corrector); -1: assume they changed
    int offDiagEpoch = -1;
    bool asymmetric() const { return lower != upper; }
};

struct SolveControls {
    int solver = 0, precond = 2;
    double tolerance = 1e-6, relTol = 0;
    int minIter = 0, maxIter = 1000;
};

struct SolvePerf {
    double initialResidual = 0, finalResidual = 0;
    int nIterations = 0, converged = 0, singular = 0;
};

class Ldu {
public:
    Ldu(Ctx* ctx, int nCells, int nFaces, const int* lower, const int* upper, int nInterfaces,
        const int* ifaceNeighbRank, const int* ifaceSize, const int* const* ifaceFaceCells);
    // block-only construction, no host addressing at all (used by the resident stepper)
    Ldu(Ctx* ctx, BlockDims dims);
    ~Ldu();
    void setBlock(int nx, int ny, int nz);          // verifies the addressing against the closed form

    void amul(const MatrixView& A, const double* psi, double* Apsi);
    void precondition(int kind, const MatrixView& A, const double* rA, double* wA);
    SolvePerf solve(const MatrixView& A, const double* source, double* psi, const SolveControls& sc);
    void levels(int* levelOfCell, int* nLevels);

    Ctx* ctx;
    int nCells = 0, nFaces = 0;
    bool isBlock = false, hasGeneral = false;
    BlockDims dims;
    // z-slab interfaces of the block path (resident stepper): psi ghost planes are exchanged in place
    bool ghostBelow = false, ghostAbove = false;
    // general addressing (device)
    int *d_lower = nullptr, *d_upper = nullptr;
    int *d_rowStart = nullptr, *d_rowMid = nullptr, *d_rowCol = nullptr, *d_rowFace = nullptr;
    int *d_lvlCells = nullptr, *d_lvlStart = nullptr;
    std::vector<int> h_lvlStart, h_level;
    double* d_sumA = nullptr;
    double* stage[8] = {nullptr};    // device staging of host-pointer calls (b200_ldu_solve)
    std::vector<double*> stageBou;
    int nLevels = 0, maxLevelSize = 0;
    double nGlobalCells = -1;        // cells over all ranks (cached)
    std::vector<int> h_lower, h_upper;
    std::vector<DevInterface> ifaces;
    // work vectors (device), allocated on first solve; each has one ghost plane either side in block mode
    double* work[10] = {nullptr};
    double* rD = nullptr;
    int64_t ghost = 0;               // cells of padding before element 0 of every work vector
    SolveState* d_state = nullptr;
    SolveState* h_state = nullptr;   // pinned
    RedWork red{};
    int redBlocks = 1;
    // tile-wavefront sweeps (block addressing): coefficient arrays renumbered into wave order
    bool useWave = true;
    double* wave[14] = {nullptr};    // A0 PX PY PZ QX QY QZ WF RF WB NX NY NZ DG
    int* waveFlags = nullptr;        // tile ticket
    int* waveOrder = nullptr;        // ticket -> tile
    int* waveOrderC[3] = {nullptr, nullptr, nullptr};   // ticket -> CTA tile of k_wave3 (4 / 2 / 1 tiles per CTA)
    long long* waveTrace = nullptr;  // B200_WAVE_TRACE diagnostic
    long long* waveTraceG = nullptr;
    double* wavePartial = nullptr;
    int64_t waveLen = 0;
    int waveEpoch = -1; const double *waveUpper = nullptr, *waveLower = nullptr; bool waveDilu = false;
    // barrier-free level-ordered sweeps over general addressing (flow_kernels.cuh)
    bool useFlow = true;
    int *d_fStart = nullptr, *d_fPos = nullptr, *d_fFace = nullptr;   // position-ordered dependency tables (see FlowArgs)
    int *d_bStart = nullptr, *d_bPos = nullptr, *d_bFace = nullptr;
    double *flowCoefF = nullptr, *flowCoefB = nullptr, *flowNumF = nullptr;   // coefficients gathered next to them, once per solve
    double *flowRD = nullptr, *flowF = nullptr, *flowB = nullptr;             // position-ordered rD, forward and backward results
    double* flowPartial = nullptr;   // per-chunk sums of the fused dot product
    int* flowTicket = nullptr;
    void flowSweep(int what, bool dilu, bool dot, const MatrixView& A, const double* rA, double* wA, const double* dotVec, int phase);
    void ensureWave(bool dilu);
    void waveSweep(int op, bool dilu, bool dot, const double* rA, double* wA, bool wio, int phase);
    int waveVecGrid();

    void ensureWork();
    double* vec(int i) { return work[i] + ghost; }
    void factor(int kind, const MatrixView& A);
    void applyPrecond(int kind, const MatrixView& A, const double* rA, double* wA, int dotMode, const double* dotVec, bool wio, int phase);
    void amulFused(const MatrixView& A, const double* x, double* y, int mode, const double* aux, int phase);
    void exchange(const double* x);   // interface / ghost-plane update before an Amul on x
    void reduceCtrl(int slot0, int n, int phase);
    RedWork redFor(int phase) const;
    int lastIterations = 0;          // iteration count of the previous solve (look-ahead of the host loop, rank-invariant)
};

// wave_kernels.cuh's shared-reciprocal division against the compiler's operator on n pseudo-random pairs; returns the mismatches
unsigned long long selftest_division(Ctx* ctx, unsigned long long n, unsigned long long seed, int expRange);

} // namespace b200
