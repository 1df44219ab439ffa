// GPU-resident thermal step on one blockMesh hex block (one z-slab of it per rank).
// Replaces the body of the reference's time loop, applications/solvers/additiveFoam/additiveFoam.C:66-95:
//   updateProperties.H, setDeltaT.H, movingHeatSourceModel::update(), solutionControls.H, thermo/TEqn.H
//   (thermoScheme.H + thermoSource.H) and, underneath, OpenFOAM-10's Euler ddt / Gauss-harmonic laplacian
//   assembly, boundary-coefficient folding (fvMatrix::solveSegregated) and lduMatrix solvers.
// Fields stay on the device for the whole run; per step the host sees a handful of scalars.
#pragma once
#include "../../include/additivefoam_b200.h"
#include "ctx.cuh"
#include "ldu.cuh"
#include "beam.cuh"
#include <vector>
#include <memory>

namespace b200 {
// mixedTemperatureFvPatchScalarField::updateCoeffs (derivedFvPatchFields/mixedTemperature/mixedTemperatureFvPatchScalarField.C:157-188):
// the value fraction of one face; refValue = Tinf, refGrad = 0.  Same expression on the host (b200_mixed_temperature_value_fraction,
// checked against the reference's source) and in k_side_coeffs.
__host__ __device__ inline double mixed_value_fraction(double h, double emissivity, double Tinf, double Tp, double kappa, double deltaCoeff)
{
    const double sigma = 5.67e-8;
    const double hEff = h + sigma * emissivity * (Tp * Tp + Tinf * Tinf) * (Tp + Tinf);
    return 1.0 / (1.0 + kappa * deltaCoeff / fmax(hEff, 1e-15));
}


struct Geom {
    double ox, oy, oz, dx, dy, dz, V;
    double sfx, sfy, sfz;       // |Sf| of x-, y-, z-normal faces
    double dcx, dcy, dcz;       // deltaCoeffs of internal faces
    double bcx, bcy, bcz;       // deltaCoeffs of boundary faces (cell centre to face)
    int k0;                     // global k index of local plane 0
    int i0 = 0, j0 = 0;         // global i, j index of local cell (0, 0, .): a sub-block of the blockMesh block (b200_beam_qdot_sub)
};

constexpr int SIDE_PROCESSOR = 3;     // extends B200_BC_*

struct SideInfo {
    int type = B200_BC_ZERO_GRADIENT; // B200_BC_* or SIDE_PROCESSOR
    int patch = -1;
    double value = 0, h = 0, emissivity = 0, Tinf = 0;
    int n = 0;                        // faces on this side (local)
    double *valueFraction = nullptr, *Tb = nullptr, *intCoeff = nullptr, *bouCoeff = nullptr;   // device, size n
};

// what the cell kernels need to know about the six sides, by value
struct SidesDev {
    int type[6];
    int order[6];                     // sides in summation order (patch order, then processor sides)
    int nOrder;
    const double* intCoeff[6];
    const double* bouCoeff[6];
    const double* valueFraction[6];
    const double* Tb[6];
    double Tinf[6];
};

struct BoxRange { int i0, i1, j0, j1, k0, k1; };   // half-open local index box

// [UPSTREAM] backwardDdtScheme coefficients: coefft = 1 + dt/(dt + dt0), coefft00 = dt*dt/(dt0*(dt + dt0)), coefft0 = coefft + coefft00.
// Euler is {1/dt, 1, 1, 0} with `on` false (the kernels then keep the Euler expressions, operation for operation).
struct BackwardCoeffs { double rDeltaT, coefft, coefft0, coefft00; bool on; };
// [UPSTREAM] CrankNicolsonDdtScheme as one call of fvcDdt / fvmDdt needs it: rDtCoef_ = coef_/deltaT, psi = ocCoeff(), the scheme's
// ddt0 field (already brought to this time index).  offCentre_(ddt0) = psi < 1 ? psi*ddt0 : ddt0.
struct CnCoeffs { double rDtCoef, psi; const double* ddt0; bool on; };
__host__ __device__ inline double cn_off_centre(const CnCoeffs& k, double ddt0) { return k.psi < 1 ? k.psi * ddt0 : ddt0; }

class ThermalCase {
public:
    ThermalCase(Ctx* ctx, const b200_case_desc& desc);
    ~ThermalCase();
    bool running() const { return time_ < (ctl.endTime - 0.5 * deltaT_); }
    void step(b200_step_report* rep);
    int64_t nLocal() const { return dims.nCells(); }
    int64_t nGlobal() const { return (int64_t)dims.nx * dims.ny * nzGlobal; }
    void resetAlphaSolidFromT();
    void getField(int field, double* out);
    void setField(int field, const double* in);
    // streamed host I/O (b200_case_set_field_async / apply_staged / get_field_async / transfers_wait)
    void setFieldAsync(int field, const double* pinnedIn);
    void applyStaged();
    void getFieldAsync(int field, double* pinnedOut);
    void transfersWait();
    void isothermDepthProbe(int beam, const double position[3], double dimensionsOut[3]);
    void sourcesUpdate(double proposedDeltaT, double* deltaTOut);
    void meltPoolDimensions(int nIso, const double* isoValues, double scanPathAngleDeg, double* dims);
    void solidificationEnable(const double boxMin[3], const double boxMax[3], double isoValue, int64_t maxEvents);
    int64_t solidificationEvents(double* out, int64_t maxOut);          // returns the number recorded so far (may exceed capacity)

    Ctx* ctx;
private:
    // description (owned copies)
    b200_controls ctl{};
    b200_material mat{};
    std::vector<double> thermoT, thermoA;
    double Tliq = 0, Tsol = 0, Tinitial = 300;
    std::vector<b200_patch> patches;
    std::vector<heatSourceModel> sources;
    BlockDims dims; int nzGlobal = 0; Geom geo{};
    int rankBelow = -1, rankAbove = -1;
    SideInfo sides[6];
    SidesDev sidesDev{};
    std::vector<bool> patchUpdated;
    // time
    double time_ = 0, deltaT_ = 0;
    double deltaT0_ = 0, deltaTSave_ = 0;        // [UPSTREAM] Time::deltaT0_, deltaTSave_ (operator++: deltaT0_ = deltaTSave_; deltaTSave_ = deltaT_)
    // ddtScheme backward (thermoScheme.H:41-49): second old-time levels, and how many levels T / alpha.solid have so far --
    // [UPSTREAM] backwardDdtScheme::deltaT0_(vf) is `great` while vf.nOldTimes() < 2; a level that is created later starts as a copy of
    // the one above it (GeometricField::oldTime())
    int TnOld = 0, a1nOld = 0;
    bool backward() const { return ctl.ddtScheme == B200_DDT_BACKWARD; }
    BackwardCoeffs backwardCoeffs(int nOld) const;
    // ddtScheme CrankNicolson (thermoScheme.H:51-65; the scheme itself is [UPSTREAM] CrankNicolsonDdtScheme, restated in
    // oracle/af_thermal.hpp where its rules are written out): per field ddt0, created zero at its first use, the time index of that
    // moment and the time index it was last brought to
    struct Ddt0State { bool exists = false; int startTimeIndex = 0, timeIndex = 0; };
    Ddt0State ddt0StateT, ddt0StateA;
    double *ddt0T = nullptr, *ddt0A = nullptr;
    bool crankNicolson() const { return ctl.ddtScheme == B200_DDT_CRANK_NICOLSON; }
    CnCoeffs cnPrepare(bool ofT, bool fvm);       // everything one fvcDdt / fvmDdt does before it uses ddt0
    void touchOldTimes(bool ofT, int levels);
    int timeIndex = 0, writeTimeIndex = 0;
    double alphaCoNum = 0, DiNum = 0;
    // fields (device, each with one ghost plane either side; pointers address cell 0)
    size_t plane = 0, padded = 0;
    std::vector<double*> allocs;
    double *T = nullptr, *Told = nullptr, *alpha1 = nullptr, *alpha1old = nullptr, *alpha3 = nullptr, *kappa = nullptr,
           *rKappa = nullptr, *Cp = nullptr, *qDot = nullptr, *qDoti = nullptr, *dFdT = nullptr, *T0 = nullptr,
           *diag0 = nullptr, *source0 = nullptr, *diag = nullptr, *source = nullptr, *upper = nullptr, *wbuf = nullptr,
           *Toldold = nullptr, *alpha1oldold = nullptr;
    size_t wbufCap = 0;
    int upperForm = 0;                   // what `upper` holds: 0 nothing valid, 1 the implicit off-diagonals, 2 the harmonic face conductivities
    bool lastExplicit = false;           // form the previous step assembled (the guess k_dinum_faces writes)
    double* d_thermo = nullptr;          // thermoPath table on the device: x then y
    std::unique_ptr<Ldu> ldu;
    RedWork red{};
    double* h_red = nullptr;             // pinned mirror of red.result
    const double* bouPtrs[2] = {nullptr, nullptr};

    // functionObjects/solidificationData state
    bool solidOn = false;
    BoxRange solidBox{};
    double solidIso = 0;
    double *solidR = nullptr, *solidEv = nullptr;
    long long* solidKeys = nullptr;
    unsigned long long* solidCount = nullptr;
    int64_t solidMax = 0, solidExec = 0;
    void solidificationExecute();

    // streamed host I/O: two copy streams beside the compute stream, one staging buffer per field and direction
    static constexpr int NIOF = 14;
    cudaStream_t h2dStream = nullptr, d2hStream = nullptr;
    double* stageIn[NIOF] = {nullptr}; double* stageOut[NIOF] = {nullptr};
    cudaEvent_t evIn[NIOF] = {nullptr}, evInFree[NIOF] = {nullptr}, evOutReady[NIOF] = {nullptr}, evOutDone[NIOF] = {nullptr};
    bool inPending[NIOF] = {false};
    void ensureIo(int f, bool out);

    double* field(int f);
    double* allocField();
    void haloExchange(double* f);
    void updatePropertiesAndNumbers(double& alphaCoMax, double& diMax, double& minAlpha1, bool propertiesOnly = false);
    void setDeltaT(double alphaCoMax, double diMax);
    void updateSources();
    void qDotOnce(heatSourceModel& hs, double* target, double dt, double sumDt, bool accumulate);
    void updateDimensions(heatSourceModel& hs);
    void updateCoeffsT();
    void evaluateT();
    void assembleBase(bool explicitSolve, double rDeltaT);
    void readReductions(int n);
    friend void beam_qdot_standalone(Ctx*, const b200_block_mesh&, const b200_beam&, const double*, const double*, double, double*,
                                     double*, double*, double*, const int*, const int*, double (*)(double, void*), void*);
};

void beam_qdot_standalone(Ctx* ctx, const b200_block_mesh& mesh, const b200_beam& beam, const double position[3],
                          const double dimensions[3], double power, double* qDot, double* sumWeights, double* volumeUsed,
                          double* absorbedPower, const int* offset = nullptr, const int* nLocal = nullptr,
                          double (*reduceSum)(double, void*) = nullptr, void* user = nullptr);

// device entry points of beam.cu
void launch_qdot_weights(Ctx* ctx, const BlockDims& d, const Geom& g, const ShapeParams& sp, const BoxRange& box, double* wbuf,
                         RedWork red, int slot);
void launch_qdot_apply(Ctx* ctx, const BlockDims& d, const BoxRange& box, const double* wbuf, double* target, double absorbedPower,
                       double V0, const double* sumWeights, double dt, double sumDt, int directMean, double* volumeOut);
void launch_isotherm_depth(Ctx* ctx, const BlockDims& d, const Geom& g, const BoxRange& cols, const double* T, double isoValue,
                           const double pos[3], double searchRadius, double staticDepth, bool ghostBelow, bool ghostAbove,
                           RedWork red, int slot);
BoxRange beam_box(const BlockDims& d, const Geom& g, const double pos[3], const double dims[3]);

} // namespace b200
