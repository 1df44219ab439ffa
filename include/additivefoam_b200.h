/*
 * additivefoam_b200.h -- C ABI of the B200-native thermal hot path for AdditiveFOAM.
 *
 * This is the drop-in boundary.  Every entry point is plain C:
 * opaque handles, plain pointers and sizes, `int` status (0 = ok, nonzero = error whose
 * text is returned by b200_last_error()).  All pointers are HOST pointers unless the
 * name ends in `_dev`.  Nothing here owns caller memory; outputs are caller-allocated.
 *
 * Reference interfaces replaced (paths relative to the AdditiveFOAM tree;
 * SOLVER = applications/solvers/additiveFoam, MHS = SOLVER/movingHeatSource;
 * [UPSTREAM] = OpenFOAM-10 source, not vendored in the reference):
 *
 *   b200_ldu_*            <- [UPSTREAM] Foam::lduMatrix::solver (PCG, PBiCGStab,
 *                            diagonalSolver) + lduMatrix::preconditioner (DIC, DILU,
 *                            diagonal), reached from SOLVER/thermo/TEqn.H:32-42 through
 *                            fvMatrix<scalar>::solveSegregated; selected by
 *                            tutorials/AMB2018-02-B/system/fvSolution:34-42.
 *   b200_beam_qdot        <- MHS/heatSourceModels/heatSourceModel/heatSourceModel.H:194
 *                            (virtual qDot()), body heatSourceModel.C:221-371.
 *   b200_case_*           <- the per-time-step block of SOLVER/additiveFoam.C:66-95:
 *                            updateProperties.H, setDeltaT.H, movingHeatSourceModel::update
 *                            (MHS/movingHeatSourceModel/movingHeatSourceModel.C:100-150),
 *                            solutionControls.H, thermo/TEqn.H (+thermoScheme.H,
 *                            thermoSource.H).
 *
 * There is no CPU fallback behind this ABI: every call that computes needs a CUDA device.
 */
#ifndef ADDITIVEFOAM_B200_H
#define ADDITIVEFOAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ enums */

/* lduMatrix::solver type ([UPSTREAM] symMatrix/asymMatrix tables; fvSolution `solver`). */
enum { B200_SOLVER_PCG = 0, B200_SOLVER_PBICGSTAB = 1, B200_SOLVER_DIAGONAL = 2 };
/* lduMatrix::preconditioner ([UPSTREAM]; fvSolution `preconditioner`). */
enum { B200_PRECOND_NONE = 0, B200_PRECOND_DIAGONAL = 1, B200_PRECOND_DIC = 2, B200_PRECOND_DILU = 3 };
/* fvPatchField types on T (tutorials/AMB2018-02-B/0/T:21-41; multiLayerPBF/0/T:23-40). */
enum { B200_BC_ZERO_GRADIENT = 0, B200_BC_FIXED_VALUE = 1, B200_BC_MIXED_TEMPERATURE = 2 };
/* Sides of a blockMesh hex block in blockMesh's own boundary order [UPSTREAM block::createBoundary]. */
enum { B200_XMIN = 0, B200_XMAX = 1, B200_YMIN = 2, B200_YMAX = 3, B200_ZMIN = 4, B200_ZMAX = 5 };
/* heatSourceModel run-time types (MHS/heatSourceModels/...). */
enum { B200_SUPER_GAUSSIAN = 0, B200_MODIFIED_SUPER_GAUSSIAN = 1, B200_PROJECTED_GAUSSIAN = 2 };
/* absorptionModel run-time types (MHS/absorptionModels/...). */
enum { B200_ABSORPTION_CONSTANT = 0, B200_ABSORPTION_KELLY = 1 };
enum { B200_KELLY_CONE = 0, B200_KELLY_CYLINDER = 1 };
enum { B200_DDT_EULER = 0, B200_DDT_BACKWARD = 1, B200_DDT_CRANK_NICOLSON = 2 };
/* Fields that b200_case_get_field / set_field can address. */
enum {
    B200_FIELD_T = 0, B200_FIELD_ALPHA_SOLID = 1, B200_FIELD_ALPHA_POWDER = 2,
    B200_FIELD_KAPPA = 3, B200_FIELD_CP = 4, B200_FIELD_QDOT = 5,
    B200_FIELD_DFDT = 6, B200_FIELD_T0 = 7,
    /* old-time copies (T.oldTime(), alpha.solid.oldTime(); [UPSTREAM] GeometricField::storeOldTimes at runTime++,
     * additiveFoam.C:76): plain upload / download, for restarts and for resynchronising a run with another one */
    B200_FIELD_T_OLD = 8, B200_FIELD_ALPHA_SOLID_OLD = 9,
    /* the second old-time level (T.oldTime().oldTime(), ...), kept only with ddtScheme backward or CrankNicolson */
    B200_FIELD_T_OLD_OLD = 10, B200_FIELD_ALPHA_SOLID_OLD_OLD = 11,
    B200_FIELD_DDT0_T = 12, B200_FIELD_DDT0_ALPHA_SOLID = 13,   /* CrankNicolson: the scheme's ddt0(T), ddt0(alpha.solid) */
    /* OR-ed into the field id of b200_case_set_field: upload only, leave the old-time copy and the patch state alone
     * (a host application that owns the field between steps re-uploads the same values) */
    B200_FIELD_KEEP_OLD = 0x100
};

/* ------------------------------------------------------- case description */

/* One hex block, uniform grading: tutorials/AMB2018-02-B/system/blockMeshDict:16-38.
 * Cell index = i + nx*(j + ny*k); internal faces in upper-triangular order. */
typedef struct b200_block_mesh {
    int32_t nx, ny, nz;
    double  min[3];
    double  max[3];
} b200_block_mesh;

/* One boundary patch of T.  kappa, Cp, alpha.* are zeroGradient on every patch
 * (SOLVER/createFields.H:16-44,84-110). */
typedef struct b200_patch {
    int32_t type;        /* B200_BC_* */
    int32_t nSides;      /* block sides making up the patch, in blockMeshDict order */
    int32_t sides[6];
    double  value;       /* fixedValue */
    double  h;           /* mixedTemperature: mixedTemperatureFvPatchScalarField.C:157-188 */
    double  emissivity;
    double  Tinf;
} b200_patch;

/* constant/transportProperties (SOLVER/readTransportProperties.H:13-73) and
 * constant/thermoPath (SOLVER/createFields.H:46-71). */
typedef struct b200_material {
    double  kappa[3][3]; /* [solid, liquid, powder][c0, c1, c2] Polynomial<3> */
    double  Cp[3][3];
    double  rho;
    double  Lf;
    int32_t nThermo;     /* rows of thermoPath */
    const double* thermoT;      /* x column: temperature */
    const double* thermoAlpha;  /* y column: solid fraction */
} b200_material;

/* One entry of constant/heatSourceDict `sources` (tutorials/AMB2018-02-B/constant/heatSourceDict:18-46). */
typedef struct b200_beam {
    int32_t heatSourceModel;   /* B200_SUPER_GAUSSIAN ... */
    double  dimensions[3];
    int32_t nPoints[3];
    int32_t transient;         /* heatSourceModel.C:104-118 */
    double  isoValue;
    double  k, m, A, B;        /* shape coefficients */
    int32_t absorptionModel;   /* B200_ABSORPTION_* */
    double  eta;               /* constant */
    int32_t kellyGeometry;     /* B200_KELLY_* */
    double  eta0, etaMin;
    double  deltaT;            /* beam sub-cycle step, <=0 means GREAT (movingBeam.C:58,62) */
    int32_t hitPathIntervals;  /* movingBeam.C:63 */
    int32_t nSegments;         /* data lines of the scan path file (header excluded) */
    const double* path;        /* nSegments x 6: mode x y z power parameter (segment.C:64-75) */
} b200_beam;

/* system/controlDict + system/fvSolution entries used by the path. */
typedef struct b200_controls {
    double  startTime, endTime, deltaT;
    int32_t adjustTimeStep;
    double  maxCo, maxDi, maxAlphaCo, maxDeltaT;
    double  writeInterval;     /* adjustableRunTime write interval; <=0 disables write-time snapping */
    int32_t nThermoCorrectors; /* SOLVER/solutionControls.H:2-12 */
    double  thermoTolerance;
    double  Tmax;              /* <=0 means vGreat */
    int32_t explicitSolve;     /* SOLVER/solutionControls.H:31-49 */
    int32_t solver;            /* B200_SOLVER_PCG | B200_SOLVER_PBICGSTAB */
    int32_t preconditioner;    /* B200_PRECOND_* */
    double  tolerance, relTol;
    int32_t minIter, maxIter;
    int32_t ddtScheme;         /* system/fvSchemes ddtSchemes default: B200_DDT_EULER | B200_DDT_BACKWARD | B200_DDT_CRANK_NICOLSON
                                * (SOLVER/thermo/thermoScheme.H:2-15,41-65) */
    double  ocCoeff;           /* CrankNicolson off-centring coefficient psi of `default CrankNicolson psi;` (0..1; 1 when the entry gives
                                * none): [UPSTREAM] CrankNicolsonDdtScheme::ocCoeff(), read by thermoScheme.H:53-64.  Ignored otherwise. */
} b200_controls;

typedef struct b200_case_desc {
    b200_block_mesh mesh;
    int32_t nPatches;
    const b200_patch* patches;
    b200_material material;
    int32_t nBeams;
    const b200_beam* beams;
    b200_controls controls;
    double  Tinitial;          /* 0/T internalField uniform */
    double  powderZMin;        /* alpha.powder = 1 for cell centres with z > powderZMin; use +1e300 for none */
} b200_case_desc;

/* What one time step reports (the reference prints these: setDeltaT.H:15,33,50;
 * solutionControls.H:18; TEqn.H:55-56; [UPSTREAM] solver performance line). */
typedef struct b200_step_report {
    double  time, deltaT;
    double  alphaCoNum, DiNum;
    double  totalPower;
    int32_t explicitStep;      /* 1 when the diagonal (explicit) form was used */
    int32_t nThermoCorr;       /* corrector passes executed */
    double  thermoResidual;    /* last max|alpha1 - alpha10| */
    int32_t nSolveIterations;  /* Krylov iterations summed over correctors */
    int32_t lastSolveIterations;
    double  lastInitialResidual, lastFinalResidual;
} b200_step_report;

typedef struct b200_solver_perf {
    double  initialResidual, finalResidual;
    int32_t nIterations, converged, singular;
} b200_solver_perf;

/* ------------------------------------------------------------ the calls */

typedef struct b200_ctx  b200_ctx;
typedef struct b200_ldu  b200_ldu;
typedef struct b200_case b200_case;

const char* b200_last_error(void);
/* ABI version, bumped on any incompatible change. */
int  b200_abi_version(void);

/* One context per process = one GPU = one subdomain (the reference's one MPI rank).
 * nccl_unique_id: 128-byte ncclUniqueId shared by all ranks, or NULL when nranks == 1. */
int  b200_nccl_unique_id(void* out128);
/* Number of CUDA devices visible to this process (a shim picks its device from it and the node-local rank). */
int  b200_device_count(int* count);
int  b200_ctx_create(int device_ordinal, const void* nccl_unique_id, int rank, int nranks, b200_ctx** out);
int  b200_ctx_destroy(b200_ctx* ctx);
/* Optional, ranks of one node: replace the NCCL all-reduces of scalars by a peer-memory (NVLink) mailbox exchange.
 * Every rank exports a 64-byte cudaIpcMemHandle_t, the host gathers them in rank order, every rank imports the list. */
int  b200_ctx_peer_export(b200_ctx* ctx, void* handle64);
int  b200_ctx_peer_import(b200_ctx* ctx, const void* handles /* nranks x 64 bytes */);
/* Same export, with a halo window of 4 x haloDoubles doubles behind the mailbox: the slab neighbours then store their boundary
 * planes (processor-patch halo swaps: [UPSTREAM] processorFvPatchField::initEvaluate / initInterfaceMatrixUpdate) straight into
 * it over NVLink -- one kernel per exchange, no NCCL call.  haloDoubles >= nx*ny of the largest case run on this context; a
 * larger plane falls back to ncclSend/ncclRecv. */
int  b200_ctx_peer_export_halo(b200_ctx* ctx, int64_t haloDoubles, void* handle64);
int  b200_ctx_synchronize(b200_ctx* ctx);
/* CUDA events on the library's own stream (bench.py times the step with these): record mark i (0..3);
 * elapsed ms between two recorded marks (synchronises). */
int  b200_ctx_mark(b200_ctx* ctx, int i);
int  b200_ctx_elapsed_ms(b200_ctx* ctx, int i, int j, double* ms);
/* Built-in per-kernel-group timing (CUDA event pair around every launch group).  read() synchronises, fills
 * ms[n]/counts[n] for the first n groups and resets the accumulators; names by index. */
int  b200_profile_enable(b200_ctx* ctx, int on);
int  b200_profile_groups(void);
const char* b200_profile_name(int group);
int  b200_profile_read(b200_ctx* ctx, double* ms, int64_t* counts, int n);

/* ---- B1: lduMatrix::solver.  Addressing is cached in the handle; coefficients are per call.
 * lower/upper: [UPSTREAM] lduAddressing::lowerAddr()/upperAddr() (upper-triangular order).
 * Interfaces are processor patches: faceCells per interface, neighbour rank per interface. */
int  b200_ldu_create(b200_ctx* ctx, int32_t nCells, int32_t nFaces,
                     const int32_t* lower, const int32_t* upper,
                     int32_t nInterfaces, const int32_t* ifaceNeighbRank, const int32_t* ifaceSize,
                     const int32_t* const* ifaceFaceCells, b200_ldu** out);
int  b200_ldu_destroy(b200_ldu* ldu);
/* Declare that the addressing is a single blockMesh hex block (verified bit-exactly against
 * lower/upper); enables the structured kernels.  Returns nonzero if it does not match. */
int  b200_ldu_set_block(b200_ldu* ldu, int32_t nx, int32_t ny, int32_t nz);
/* [UPSTREAM] lduMatrix::solver::solve(psi, source).  lower == NULL means symmetric
 * (lower aliases upper).  psi is in/out. */
int  b200_ldu_solve(b200_ldu* ldu, int32_t solver, int32_t preconditioner,
                    const double* diag, const double* upper, const double* lower,
                    const double* const* ifaceBouCoeffs,
                    const double* source, double* psi,
                    double tolerance, double relTol, int32_t minIter, int32_t maxIter,
                    b200_solver_perf* perf);
/* Same with device-resident arrays (no copies); used by the resident stepper and the benchmark. */
int  b200_ldu_solve_dev(b200_ldu* ldu, int32_t solver, int32_t preconditioner,
                        const double* diag_dev, const double* upper_dev, const double* lower_dev,
                        const double* const* ifaceBouCoeffs_dev,
                        const double* source_dev, double* psi_dev,
                        double tolerance, double relTol, int32_t minIter, int32_t maxIter,
                        b200_solver_perf* perf);
/* [UPSTREAM] lduMatrix::Amul on its own (the "Amul HBM GB/s" micro-benchmark and parity probe). */
int  b200_ldu_amul(b200_ldu* ldu, const double* diag, const double* upper, const double* lower,
                   const double* const* ifaceBouCoeffs, const double* psi, double* Apsi);
int  b200_ldu_amul_dev(b200_ldu* ldu, const double* diag_dev, const double* upper_dev,
                       const double* lower_dev, const double* const* ifaceBouCoeffs_dev,
                       const double* psi_dev, double* Apsi_dev);
/* One application of the preconditioner, wA = M^-1 rA (factor computed from the given matrix). */
int  b200_ldu_precondition(b200_ldu* ldu, int32_t preconditioner,
                           const double* diag, const double* upper, const double* lower,
                           const double* rA, double* wA);
/* Level schedule of the cached addressing (renumbering parity probe): level of each cell. */
int  b200_ldu_levels(b200_ldu* ldu, int32_t* levelOfCell, int32_t* nLevels);

/* ---- B2: heatSourceModel::qDot() for one beam state on a block mesh.
 * Writes qDot (W/m^3) for all cells (host array, mesh.nx*ny*nz) and the Riemann sum of V*w. */
int  b200_beam_qdot(b200_ctx* ctx, const b200_block_mesh* mesh, const b200_beam* beam,
                    const double position[3], const double dimensions[3], double power,
                    double* qDot, double* sumWeights, double* volumeUsed, double* absorbedPower);
/* The same for ONE RANK of a decomposed run (heatSourceModel.C:287-353 loops over the rank's own cells; the normalisation
 * sum of :359 is global): the rank owns the sub-block of nLocal[] cells that starts at global cell index offset[] of `mesh`
 * (a `simple` / `hierarchical` decomposition of a blockMesh block; cell corners are those of the global block).  `reduceSum` adds
 * the ranks' local sum(V*w) -- the shim passes a function that calls Foam::reduce(..., sumOp<scalar>()); every rank must call in,
 * also those the beam does not touch.  qDot: nLocal[0]*nLocal[1]*nLocal[2] values in the rank's local cell order. */
typedef double (*b200_reduce_sum_fn)(double localValue, void* user);
int  b200_beam_qdot_sub(b200_ctx* ctx, const b200_block_mesh* mesh, const int32_t offset[3], const int32_t nLocal[3],
                        const b200_beam* beam, const double position[3], const double dimensions[3], double power,
                        b200_reduce_sum_fn reduceSum, void* user, double* qDot, double* sumWeights, double* volumeUsed,
                        double* absorbedPower);

/* The two pure virtuals of heatSourceModel (heatSourceModel.H:197-200) and absorptionModel::eta, evaluated by the library's host
 * mirror of the beam models (no device, no context): V0() for the given current dimensions -- for projectedGaussian it also
 * resolves the shape exponent k from the depth, projectedGaussian.C:96-106, returned in kResolved (may be NULL) -- and weight(d)
 * for a component-wise distance d from the beam centre.  The OpenFOAM-side heat-source shims forward their virtuals to
 * these, so the shape formulas live in one place. */
int  b200_beam_shape_v0(const b200_beam* beam, const double dimensions[3], double* V0, double* kResolved);
int  b200_beam_shape_weight(const b200_beam* beam, const double dimensions[3], const double d[3], double* weight);
int  b200_beam_absorption_eta(const b200_beam* beam, double aspectRatio, double* eta);
/* Host mirror of the top-surface boundary condition (no device, no context): the value fraction
 * mixedTemperatureFvPatchScalarField::updateCoeffs (derivedFvPatchFields/mixedTemperature/mixedTemperatureFvPatchScalarField.C:
 * 157-188) assigns to a face with patch temperature Tp, patch conductivity kappa and deltaCoeffs deltaCoeff (refValue = Tinf,
 * refGradient = 0).  The same function is what the device kernel evaluates. */
int  b200_mixed_temperature_value_fraction(double h, double emissivity, double Tinf, double Tp, double kappa, double deltaCoeff,
                                           double* valueFraction);

/* ---- The GPU-resident thermal step (everything in SURVEY section 3.2 items 1,3,4,6,8). */
int  b200_case_create(b200_ctx* ctx, const b200_case_desc* desc, b200_case** out);
int  b200_case_destroy(b200_case* c);
/* 1 while the reference's `runTime.run()` would be true. */
int  b200_case_running(b200_case* c);
int  b200_case_step(b200_case* c, b200_step_report* report);
/* Fixed-dt stepping for throughput runs: same step, adjustTimeStep bypassed. */
int  b200_case_n_cells(b200_case* c, int64_t* nLocal, int64_t* nGlobal);
int  b200_case_get_field(b200_case* c, int32_t field, double* out);
int  b200_case_set_field(b200_case* c, int32_t field, const double* in);
/* Restart (SOLVER/createFields.H:31-44,74-78): alpha.solid is never written or read by the reference -- after T has been set from a
 * time directory it is rebuilt as min(max(interpolateXY(T, thermoPath), 0), 1), its old-time level with it, and the properties are
 * updated (createFields.H:112). */
int  b200_case_reset_alpha_solid(b200_case* c);
/* Streamed host I/O for an application that feeds the step and consumes its result without waiting on either (e.g. a solver
 * application whose function objects read T every step in their `execute()`).  Host buffers must be pinned.
 *   set_field_async : host -> a device staging buffer on a copy stream of its own; returns at once.
 *   apply_staged    : everything staged becomes the field (device-to-device, upload-only semantics of B200_FIELD_KEEP_OLD).
 *   get_field_async : snapshot of the field, then device -> host on a second copy stream; returns at once.
 *   transfers_wait  : both copy streams and the compute stream have drained; the host buffers hold the snapshots.
 * Typical loop: set_field_async(inputs of step 0); for n: apply_staged; set_field_async(inputs of step n+1); step;
 * get_field_async(results of step n).  The copies of step n+1 / step n then run under the kernels of step n / n+1.
 * An upload of a field is ordered after the outstanding download of the same field of the same case, so "results to the host buffer,
 * the host buffer back as the next inputs" can be queued without a host wait in between: two cases sharing one GPU then hide each
 * other's PCIe traffic -- get_field_async(B), set_field_async(B), step(A), apply_staged(B); the same for A under step(B). */
int  b200_case_set_field_async(b200_case* c, int32_t field, const double* pinnedIn);
int  b200_case_apply_staged(b200_case* c);
int  b200_case_get_field_async(b200_case* c, int32_t field, double* pinnedOut);
int  b200_case_transfers_wait(b200_case* c);
/* functionObjects/meltPoolDimensions/meltPoolDimensions.C:123-294 evaluated on the device-resident T (no D2H of the field):
 * for each isoValue the bounding box, in the frame rotated about z by scanPathAngle (degrees), of the isotherm located linearly
 * between cell centres across internal and processor faces plus the centres of physical boundary faces whose cell or patch value
 * reaches it; dims[3*i+0..2] = max(span, 0) in x, y, z -- the three numbers the reference logs per isoValue (:283-291).
 * Collective over the ranks of the context. */
int  b200_case_melt_pool_dimensions(b200_case* c, int32_t nIso, const double* isoValues, double scanPathAngleDeg, double* dims);
/* functionObjects/solidificationData/solidificationData.C: once enabled, every b200_case_step ends with the function object's
 * execute() (:140-187) on the device: each cell of the box (setOverlapCells, :99-131) that cooled through isoValue during the
 * step appends {x, y, z, t, |R|, max(G, small), |R|/G} to a device buffer of maxEvents rows; G = mag(fvc::grad(T)) ("Gauss
 * linear"), R the cooling rate kept from the previous execute, as in the reference.  Per rank, like the reference's
 * data_<proc>.csv (:189-222). */
/* heatSourceModel::updateDimensions (heatSourceModel.C:123-219) of beam `beam` evaluated on the resident T with the beam centre at
 * `position`: the dimensions a transient beam would take (x, y static; z the deepest isoValue crossing within the search
 * radius, never less than the static depth).  The beam's own state is not changed.  Collective over the ranks of the case. */
int  b200_case_isotherm_depth(b200_case* c, int32_t beam, const double position[3], double dimensions[3]);
/* The public interface of the reference's movingHeatSourceModel (movingHeatSourceModel.H:84-94) on the resident case, for a host
 * that keeps its own time loop: adjustDeltaT(deltaT) over all beams (movingHeatSourceModel.C:92-98; *deltaT = the adjusted
 * value, as setDeltaT.H:47-49 then hands to runTime), update() over [time, time + *deltaT] (:100-150: every active beam
 * sub-cycled with its own deltaT, dt-weighted mean, summed), after which the case's clock advances by *deltaT
 * (additiveFoam.C:74-76).  qDot() is then b200_case_get_field(FIELD_QDOT).  T and the phase fields are not touched.
 * Collective over the ranks of the case. */
int  b200_case_sources_update(b200_case* c, double proposedDeltaT, double* deltaT);
int  b200_case_solidification_enable(b200_case* c, const double boxMin[3], const double boxMax[3], double isoValue, int64_t maxEvents);
/* Copies this rank's events, in the reference's append order (time step, then cell), into events[maxEvents][7] (may be NULL to
 * query).  *nRecorded = events seen so far; a value above the capacity given at enable means the excess was dropped. */
int  b200_case_solidification_events(b200_case* c, double* events, int64_t maxEvents, int64_t* nRecorded);
/* Diagnostic: the DIC/DILU factor sweep divides with a reciprocal refinement shared between the three quotients of one
 * diagonal entry ([UPSTREAM] DICPreconditioner::calcReciprocalD's `upper*upper/rD`); this runs that sequence against the
 * compiler's FP64 division on nPairs pseudo-random operand pairs with exponents in +-expRange and returns the number of
 * bit patterns that differ (0 inside the range the sweep uses it for, +-250). */
int  b200_selftest_division(b200_ctx* ctx, int64_t nPairs, int64_t seed, int32_t expRange, int64_t* mismatches);
/* Device time (ms) spent in the kernels of the last step by group, for the benchmark:
 * [0] properties+dt numbers, [1] heat source, [2] assembly, [3] corrector fold+alpha update,
 * [4] Krylov solve. */
int  b200_case_kernel_launches(b200_case* c, int64_t* launches);

/* Scan path text -> nSegments x 6 table (MHS/movingBeam/movingBeam.C:89-148: header line
 * skipped, blank lines skipped).  Returns the number of segments, or -1 if they do not fit; out == NULL only counts. */
int  b200_scan_path_parse(const char* text, double* out, int32_t maxSegments);

/* Raster scan path of one rotation angle as file text -- what applications/utilities/createScanPath writes into
 * constant/scanPath_<i> (createScanPath.C:74-101 loops over nRotations with rotation += angle; createFields.H:265-392 makes
 * and writes the lines): hatch lines `hatch` apart about the centre of the box [minPoint, maxPoint], turned by rotationDeg
 * about that centre, cropped to the box, odd lines reversed when biDirection, each preceded by a point-source row carrying
 * dwellTime (none before the first).  Writes at most `capacity` bytes (NUL-terminated when it fits) and the full length
 * (without NUL) to *length; returns nonzero on bad arguments.  Byte-identical to the reference's writer. */
int  b200_create_scan_path(const double minPoint[2], const double maxPoint[2], double hatch, double rotationDeg,
                           double power, double speed, double dwellTime, int32_t biDirection,
                           char* text, int64_t capacity, int64_t* length);

#ifdef __cplusplus
}
#endif
#endif /* ADDITIVEFOAM_B200_H */
