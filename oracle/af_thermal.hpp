// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
// CPU restatement of AdditiveFOAM's per-time-step thermal path
// (applications/solvers/additiveFoam/additiveFoam.C:66-95 = "SOLVER/") on a single
// blockMesh hex block, serial or decomposed into z-slabs (the reference's MPI ranks are
// emulated in-process).  PARITY UNPINNED at the OpenFOAM-10 boundary: fvm::ddt,
// fvm::laplacian, harmonic interpolation, boundary coefficients and fvMatrix::solve are
// restated from upstream, not executed (thermoScheme.H and the
// corrector loop of TEqn.H are operator expressions over them).  See af_ldu.hpp header.
// PINNED on the reference's own text or sources built into oracle/_ref (oracle/Makefile;
// fixtures tests/golden/ref_*.json): updateProperties.H and thermoSource.H (verbatim,
// bit-exact, tests/test_ref_models.py), setDeltaT.H and solutionControls.H (verbatim,
// bit-exact per step, tests/test_ref_ctl.py), heatSourceModel.C qDot / updateDimensions
// (tests/test_ref_qdot.py), movingHeatSourceModel.C adjustDeltaT / update
// (tests/test_ref_mhs.py), mixedTemperatureFvPatchScalarField.C updateCoeffs (bit-exact,
// tests/test_ref_bc.py), and the function objects meltPoolDimensions.C and
// solidificationData.C (tests/test_ref_fo.py, tests/test_ref_solid.py).
#pragma once
#include "af_beam.hpp"
#include "../include/additivefoam_b200.h"
#include <array>
#include <cstdio>
#include <stdexcept>

namespace afo {

constexpr int PROCESSOR = 3;    // extends B200_BC_*

struct Patch {
    int type = B200_BC_ZERO_GRADIENT;
    std::vector<label> faceCells;
    std::vector<scalar> magSf, deltaCoeffs;
    scalar fixedValue = 0, h = 0, emissivity = 0, Tinf = 0;
    int ifaceIndex = -1;                        // processor: index into Mesh::interfaces
    // state of the T patch field
    std::vector<scalar> Tb, valueFraction;
    bool updated = false;
};

// One subdomain: nx x ny x nzLocal lattice, global k offset k0.
struct Mesh {
    label nx = 0, ny = 0, nz = 0, k0 = 0, nzGlobal = 0;
    scalar dx = 0, dy = 0, dz = 0, V = 0;
    Vec3 origin;
    LduAddr addr;
    std::vector<scalar> magSf, deltaCoeffs;     // per internal face
    std::vector<Patch> patches;
    std::vector<Interface> interfaces;
    label nCells() const { return addr.nCells; }
    // points of the block: x_i = xmin + i*dx  (blockMesh uniform grading; unpinned to the ulp)
    scalar px(label i) const { return origin.x + i * dx; }
    scalar py(label j) const { return origin.y + j * dy; }
    scalar pz(label k) const { return origin.z + (k + k0) * dz; }
    Vec3 C(label c) const
    {
        const label i = c % nx, j = (c / nx) % ny, k = c / (nx * ny);
        return {origin.x + (i + 0.5) * dx, origin.y + (j + 0.5) * dy, origin.z + (k + k0 + 0.5) * dz};
    }
};

// blockMesh ordering.  Rank r of R owns global k in [kBeg, kEnd).
inline void buildMesh(Mesh& m, const b200_case_desc& d, int r, int R)
{
    const b200_block_mesh& bm = d.mesh;
    m.nx = bm.nx; m.ny = bm.ny; m.nzGlobal = bm.nz;
    const label kBeg = (label)((int64_t)bm.nz * r / R), kEnd = (label)((int64_t)bm.nz * (r + 1) / R);
    m.k0 = kBeg; m.nz = kEnd - kBeg;
    m.dx = (bm.max[0] - bm.min[0]) / bm.nx;
    m.dy = (bm.max[1] - bm.min[1]) / bm.ny;
    m.dz = (bm.max[2] - bm.min[2]) / bm.nz;
    m.V = m.dx * m.dy * m.dz;
    m.origin = {bm.min[0], bm.min[1], bm.min[2]};
    const label nx = m.nx, ny = m.ny, nz = m.nz;
    LduAddr& ad = m.addr;
    ad.nCells = nx * ny * nz;
    ad.nFaces = (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1);
    ad.lower.reserve(ad.nFaces); ad.upper.reserve(ad.nFaces);
    m.magSf.reserve(ad.nFaces); m.deltaCoeffs.reserve(ad.nFaces);
    for (label k = 0; k < nz; k++)
        for (label j = 0; j < ny; j++)
            for (label i = 0; i < nx; i++) {
                const label c = i + nx * (j + ny * k);
                if (i < nx - 1) { ad.lower.push_back(c); ad.upper.push_back(c + 1); m.magSf.push_back(m.dy * m.dz); m.deltaCoeffs.push_back(1.0 / m.dx); }
                if (j < ny - 1) { ad.lower.push_back(c); ad.upper.push_back(c + nx); m.magSf.push_back(m.dx * m.dz); m.deltaCoeffs.push_back(1.0 / m.dy); }
                if (k < nz - 1) { ad.lower.push_back(c); ad.upper.push_back(c + nx * ny); m.magSf.push_back(m.dx * m.dy); m.deltaCoeffs.push_back(1.0 / m.dz); }
            }
    ad.finalize();
    // physical patches: faces of each listed block side in blockMesh order
    // ([UPSTREAM] block::createBoundary: x sides k-outer/j-inner, y sides i-outer/k-inner,
    //  z sides i-outer/j-inner), restricted to this rank.
    m.patches.clear();
    for (int p = 0; p < d.nPatches; p++) {
        const b200_patch& ps = d.patches[p];
        Patch pa;
        pa.type = ps.type; pa.fixedValue = ps.value; pa.h = ps.h; pa.emissivity = ps.emissivity; pa.Tinf = ps.Tinf;
        for (int s = 0; s < ps.nSides; s++) {
            const int side = ps.sides[s];
            auto add = [&](label c, scalar area, scalar halfSpacing) {
                pa.faceCells.push_back(c); pa.magSf.push_back(area); pa.deltaCoeffs.push_back(1.0 / halfSpacing);
            };
            if (side == B200_XMIN || side == B200_XMAX) {
                const label i = side == B200_XMIN ? 0 : nx - 1;
                for (label k = 0; k < nz; k++) for (label j = 0; j < ny; j++) add(i + nx * (j + ny * k), m.dy * m.dz, 0.5 * m.dx);
            } else if (side == B200_YMIN || side == B200_YMAX) {
                const label j = side == B200_YMIN ? 0 : ny - 1;
                for (label i = 0; i < nx; i++) for (label k = 0; k < nz; k++) add(i + nx * (j + ny * k), m.dx * m.dz, 0.5 * m.dy);
            } else {
                const bool mine = side == B200_ZMIN ? (r == 0) : (r == R - 1);
                if (!mine) continue;
                const label k = side == B200_ZMIN ? 0 : nz - 1;
                for (label i = 0; i < nx; i++) for (label j = 0; j < ny; j++) add(i + nx * (j + ny * k), m.dx * m.dy, 0.5 * m.dz);
            }
        }
        // 0/T `value` entries: a fixedValue patch is constructed from its own `value` ([UPSTREAM] fixedValueFvPatchField reads
        // it as the patch field), the others start from the entry the tutorials give them, the internal value.  (Until a case
        // with a wall temperature different from the initial one was read from a directory, this started every patch from
        // Tinitial: identical for all shipped cases, wrong for the first step otherwise -- found by tests/test_foam_case.py.)
        pa.Tb.assign(pa.faceCells.size(), pa.type == B200_BC_FIXED_VALUE ? pa.fixedValue : d.Tinitial);
        pa.valueFraction.assign(pa.faceCells.size(), 0.0);
        m.patches.push_back(std::move(pa));
    }
    // processor patches, neighbour rank ascending ([UPSTREAM] decomposePar); faces in owner-cell order
    m.interfaces.clear();
    for (int side = 0; side < 2; side++) {
        const int nb = side == 0 ? r - 1 : r + 1;
        if (nb < 0 || nb >= R) continue;
        Patch pa; pa.type = PROCESSOR;
        Interface itf; itf.neighbRank = nb;
        const label k = side == 0 ? 0 : nz - 1;
        const label nbNz = (label)((int64_t)bm.nz * (nb + 1) / R) - (label)((int64_t)bm.nz * nb / R);
        const label kn = side == 0 ? nbNz - 1 : 0;
        for (label j = 0; j < ny; j++) for (label i = 0; i < nx; i++) {
            const label c = i + nx * (j + ny * k);
            pa.faceCells.push_back(c); pa.magSf.push_back(m.dx * m.dy); pa.deltaCoeffs.push_back(1.0 / m.dz);
            itf.faceCells.push_back(c); itf.nbrCells.push_back(i + nx * (j + ny * kn));
        }
        pa.ifaceIndex = (int)m.interfaces.size();
        m.interfaces.push_back(std::move(itf));
        m.patches.push_back(std::move(pa));
    }
}

// SOLVER/interpolateXY/interpolateXY.C:33-109
inline void interpolateXYLabels(scalar x, const Field& xOld, label& lo, label& hi)
{
    const label n = (label)xOld.size();
    for (lo = 0; lo < n && xOld[lo] > x; ++lo) {}
    label low = lo;
    if (low < n) for (label i = low; i < n; ++i) if (xOld[i] > xOld[lo] && xOld[i] <= x) lo = i;
    for (hi = 0; hi < n && xOld[hi] < x; ++hi) {}
    label high = hi;
    if (high < n) for (label i = high; i < n; ++i) if (xOld[i] < xOld[hi] && xOld[i] >= x) hi = i;
}
inline scalar interpolateXY(scalar x, const Field& xOld, const Field& yOld)
{
    const label n = (label)xOld.size();
    label lo, hi;
    interpolateXYLabels(x, xOld, lo, hi);
    if (lo < n && hi < n && lo != hi) return yOld[lo] + ((x - xOld[lo]) / (xOld[hi] - xOld[lo])) * (yOld[hi] - yOld[lo]);
    else if (lo == hi) return yOld[lo];
    else if (lo == n) return yOld[hi];
    else return yOld[lo];
}

// [UPSTREAM] Polynomial<3>::value
inline scalar poly3(const scalar* v, scalar x)
{
    scalar val = v[0], powX = 1;
    for (int i = 1; i < 3; ++i) { powX *= x; val += v[i] * powX; }
    return val;
}

struct HeatSource {
    BeamShape shape;
    Absorption absorption;
    MovingBeam beam;
    label nPoints[3] = {1, 1, 1};
    bool transient = false;
    scalar isoValue = 0;
};

struct RankState {
    Mesh mesh;
    Field T, Told, alpha1, alpha1old, alpha3, kappa, Cp, qDot, dFdT, T0;
    Field Toldold, alpha1oldold;        // second old-time level (ddtScheme backward, CrankNicolson)
    Field ddt0T, ddt0A;                 // [UPSTREAM] CrankNicolsonDdtScheme: ddt0(T), ddt0(alpha.solid)
    bool alpha1HasOld = false;
    // base TEqn, built once per step (thermoScheme.H)
    LduMatrix teqn;                     // diag, upper (implicit) + bouCoeffs for processor patches
    Field source;
    std::vector<Field> internalCoeffs, boundaryCoeffs;   // per patch
};

struct Case {
    b200_case_desc d;
    std::vector<b200_patch> patchStore;
    std::vector<b200_beam> beamStore;
    std::vector<Field> pathStore;
    Field thermoX, thermoY;
    int R = 1;
    std::vector<RankState> rk;
    std::vector<HeatSource> sources;
    scalar Tliq = 0, Tsol = 0;
    // Time ([UPSTREAM] OpenFOAM/db/Time)
    scalar time = 0, deltaT = 0, deltaT0 = 0, deltaTSave = 0;          // [UPSTREAM] Time::deltaT_, deltaT0_, deltaTSave_
    // [UPSTREAM] GeometricField old-time levels exist only once asked for (oldTime() creates a level as a copy of the one above it):
    // how many T and alpha.solid have -- backwardDdtScheme::deltaT0_(vf) returns great while vf.nOldTimes() < 2
    int TnOld = 0, a1nOld = 0;
    bool backward() const { return d.controls.ddtScheme == B200_DDT_BACKWARD; }
    struct BackwardCoeffs { scalar rDeltaT, coefft, coefft0, coefft00; };
    // [UPSTREAM] backwardDdtScheme: coefft = 1 + dt/(dt + dt0), coefft00 = dt*dt/(dt0*(dt + dt0)), coefft0 = coefft + coefft00
    BackwardCoeffs backwardCoeffs(int nOld) const
    {
        const scalar dt0 = nOld < 2 ? 1.0e15 : deltaT0;
        BackwardCoeffs k;
        k.rDeltaT = 1.0 / deltaT;
        k.coefft = 1 + deltaT / (deltaT + dt0);
        k.coefft00 = deltaT * deltaT / (dt0 * (deltaT + dt0));
        k.coefft0 = k.coefft + k.coefft00;
        return k;
    }
    void touchOldTimes(bool ofT, int levels);
    // [UPSTREAM] CrankNicolsonDdtScheme (finiteVolume/finiteVolume/ddtSchemes/CrankNicolsonDdtScheme), restated.  Per field the scheme
    // keeps ddt0(name), created zero at its first use with startTimeIndex = the time index of that moment (DDt0Field, NO_READ form),
    // and with GeometricField's own timeIndex set to the same.  With psi = ocCoeff():
    //   coef_  = timeIndex > startTimeIndex ? 1 + psi : 1          rDtCoef_  = coef_/deltaT
    //   coef0_ = timeIndex > startTimeIndex + 1 ? 1 + psi : 1      rDtCoef0_ = coef0_/deltaT0          (Time::deltaT0_, unguarded)
    //   offCentre_(ddt0) = psi < 1 ? psi*ddt0 : ddt0
    //   evaluate(ddt0) = ddt0.timeIndex() != timeIndex : ddt0 = rDtCoef0_*(vf.oldTime() - vf.oldTime().oldTime()) - offCentre_(ddt0)
    //   fvcDdt = rDtCoef_*(vf - vf.oldTime()) - offCentre_(ddt0)
    //   fvmDdt: diag = rDtCoef_*V, source = (rDtCoef_*vf.oldTime() + offCentre_(ddt0))*V   (vf.oldTime().oldTime() touched in any case)
    //   and ddt0.timeIndex() = timeIndex at the end of either.
    // ddt0 is AUTO_WRITE in the reference and read back at a restart; this restatement starts the history anew at a restart.
    struct Ddt0State { bool exists = false; label startTimeIndex = 0, timeIndex = 0; };
    Ddt0State ddt0StateT, ddt0StateA;
    bool crankNicolson() const { return d.controls.ddtScheme == B200_DDT_CRANK_NICOLSON; }
    scalar offCentre(scalar ddt0) const { return d.controls.ocCoeff < 1 ? d.controls.ocCoeff * ddt0 : ddt0; }
    scalar cnPrepare(bool ofT, bool fvm);       // everything up to the use of ddt0; returns rDtCoef_
    // test probe: the scheme's state as the last step's TEqn.H found it (right after runTime++), for tests/golden/make_ref_teqn_golden.py
    struct CnSnapshot { Ddt0State T, A; int TnOld = 0, a1nOld = 0; label timeIndex = 0; Fields ddt0T, ddt0A; } cnSnap;
    label timeIndex = 0, writeTimeIndex = 0;
    scalar alphaCoNum = 0, DiNum = 0;
    bool verbose = false;
    std::vector<scalar> lastResidualLog;

    void init(const b200_case_desc& desc, int nRanks);
    bool running() const { return time < (d.controls.endTime - 0.5 * deltaT); }   // Time::run()
    void step(b200_step_report* rep);

    // pieces (public for unit parity tests)
    void updateProperties();
    void correctTBoundary(int r);
    void updateCoeffsT(int r);
    scalar surfaceSumDiMax();          // max_c S[c]/(V rho Cp[c])  (multiply by dt for DiNum)
    void setDeltaT();
    void updateSources();
    void heatSourceQDot(HeatSource& hs, Fields& out, scalar* sumWeightsOut = nullptr, scalar* volumeOut = nullptr,
                        scalar* absorbedOut = nullptr);
    void updateDimensions(HeatSource& hs);
    void assembleBase(bool explicitSolve);
    void thermoSource(int r);
    SolverPerf solveCorrector(bool explicitSolve, scalar rDeltaT, scalar rDeltaTEuler);
    void meltPoolDimensions(const std::vector<scalar>& isoValues, scalar scanPathAngleDeg, std::vector<Vec3>& dims);
    // functionObjects/solidificationData
    bool solidOn = false;
    Vec3 solidBoxMin{0, 0, 0}, solidBoxMax{0, 0, 0};
    scalar solidIso = 0;
    std::vector<Field> solidR;                          // R_ per rank
    std::vector<std::vector<label>> solidCells;         // overlapCells per rank
    std::vector<std::vector<std::array<scalar, 7>>> solidEvents;   // per rank, in append order
    void solidificationEnable(Vec3 boxMin, Vec3 boxMax, scalar isoValue);
    void solidificationExecute();
};

// ---------------------------------------------------------------------------------------
inline void Case::init(const b200_case_desc& desc, int nRanks)
{
    d = desc;
    patchStore.assign(desc.patches, desc.patches + desc.nPatches); d.patches = patchStore.data();
    beamStore.assign(desc.beams, desc.beams + desc.nBeams);
    pathStore.resize(desc.nBeams);
    for (int b = 0; b < desc.nBeams; b++) {
        pathStore[b].assign(desc.beams[b].path, desc.beams[b].path + 6 * (size_t)desc.beams[b].nSegments);
        beamStore[b].path = pathStore[b].data();
    }
    d.beams = beamStore.data();
    thermoX.assign(desc.material.thermoT, desc.material.thermoT + desc.material.nThermo);
    thermoY.assign(desc.material.thermoAlpha, desc.material.thermoAlpha + desc.material.nThermo);
    d.material.thermoT = thermoX.data(); d.material.thermoAlpha = thermoY.data();
    R = nRanks;
    if (R > d.mesh.nz) throw std::runtime_error("more ranks than z layers");
    // createFields.H:59-71
    Tliq = interpolateXY(0.0, thermoY, thermoX);
    Tsol = interpolateXY(1.0, thermoY, thermoX);
    rk.resize(R);
    forRanks(R, [&](int r) {
        RankState& s = rk[r];
        buildMesh(s.mesh, d, r, R);
        const label n = s.mesh.nCells();
        s.T.assign(n, d.Tinitial); s.Told = s.T; s.Toldold = s.T;
        s.alpha1.assign(n, 0); s.alpha3.assign(n, 0);
        for (label c = 0; c < n; c++) {
            // createFields.H:74-78
            const scalar a = interpolateXY(s.T[c], thermoX, thermoY);
            s.alpha1[c] = std::min(std::max(a, 0.0), 1.0);
            // setFields box on alpha.powder (tutorials/multiLayerPBF/system/setFieldsDict:18-29)
            if (s.mesh.C(c).z > d.powderZMin) s.alpha3[c] = 1.0;
        }
        s.alpha1old = s.alpha1; s.alpha1oldold = s.alpha1;
        s.kappa.assign(n, 0); s.Cp.assign(n, 0); s.qDot.assign(n, 0); s.dFdT.assign(n, 0); s.T0 = s.T;
        s.teqn.addr = &s.mesh.addr; s.teqn.interfaces = &s.mesh.interfaces;
    });
    time = d.controls.startTime; deltaT = d.controls.deltaT; deltaT0 = deltaT; deltaTSave = deltaT; timeIndex = 0; writeTimeIndex = 0;
    TnOld = 0; a1nOld = 0; ddt0StateT = Ddt0State(); ddt0StateA = Ddt0State();
    sources.resize(d.nBeams);
    for (int b = 0; b < d.nBeams; b++) {
        const b200_beam& bs = d.beams[b];
        HeatSource& hs = sources[b];
        hs.shape.model = bs.heatSourceModel; hs.shape.k = bs.k; hs.shape.m = bs.m; hs.shape.A = bs.A; hs.shape.B = bs.B;
        hs.shape.dimensions = {bs.dimensions[0], bs.dimensions[1], bs.dimensions[2]};
        hs.shape.staticDimensions = hs.shape.dimensions;
        if (hs.shape.model == B200_PROJECTED_GAUSSIAN) hs.shape.updateProjectedK();
        for (int q = 0; q < 3; q++) hs.nPoints[q] = bs.nPoints[q];
        hs.transient = bs.transient != 0; hs.isoValue = bs.isoValue;
        hs.absorption.model = bs.absorptionModel; hs.absorption.etaConst = bs.eta;
        hs.absorption.geometry = bs.kellyGeometry; hs.absorption.eta0 = bs.eta0; hs.absorption.etaMin = bs.etaMin;
        hs.beam.init(bs.path, bs.nSegments, bs.deltaT, bs.hitPathIntervals != 0, time, d.controls.endTime);
    }
    updateProperties();                 // createFields.H:112
}

// SOLVER/updateProperties.H:1-29
inline void Case::updateProperties()
{
    const b200_material& mt = d.material;
    forRanks(R, [&](int r) {
        RankState& s = rk[r];
        for (label c = 0; c < s.mesh.nCells(); c++) {
            const scalar a1 = s.alpha1[c], a2 = 1.0 - s.alpha1[c];
            if (a1 < 1.0) s.alpha3[c] = 0.0;
            const scalar a3 = s.alpha3[c];
            const scalar T1 = std::min(std::max(s.T[c], 300.0), Tsol);
            const scalar T2 = std::min(std::max(s.T[c], Tliq), 2.0 * Tliq);
            s.kappa[c] = (1.0 - a3) * (a1 * poly3(mt.kappa[0], T1) + a2 * poly3(mt.kappa[1], T2)) + a3 * poly3(mt.kappa[2], T1);
            s.Cp[c] = (1.0 - a3) * (a1 * poly3(mt.Cp[0], T1) + a2 * poly3(mt.Cp[1], T2)) + a3 * poly3(mt.Cp[2], T1);
        }
    });
}

// mixedTemperatureFvPatchScalarField.C:157-188 (updateCoeffs) for every T patch of rank r
inline scalar mixedValueFraction(scalar h, scalar emissivity, scalar Tinf, scalar Tp, scalar kap, scalar deltaCoeff)
{
    const scalar sigma = 5.67e-8;
    const scalar hEff = h + sigma * emissivity * (Tp * Tp + Tinf * Tinf) * (Tp + Tinf);
    return 1.0 / (1.0 + kap * deltaCoeff / std::max(hEff, 1e-15));
}
inline void Case::updateCoeffsT(int r)
{
    RankState& s = rk[r];
    for (Patch& p : s.mesh.patches) {
        if (p.type != B200_BC_MIXED_TEMPERATURE || p.updated) continue;
        const scalar sigma = 5.67e-8;
        for (size_t i = 0; i < p.faceCells.size(); i++) {
            const scalar Tp = p.Tb[i];
            const scalar kap = s.kappa[p.faceCells[i]];                 // zeroGradient kappa patch value
            (void)sigma;
            p.valueFraction[i] = mixedValueFraction(p.h, p.emissivity, p.Tinf, Tp, kap, p.deltaCoeffs[i]);
        }
        p.updated = true;
    }
}

// T.correctBoundaryConditions(): [UPSTREAM] mixedFvPatchField::evaluate, fixedValue, zeroGradient
inline void Case::correctTBoundary(int r)
{
    RankState& s = rk[r];
    updateCoeffsT(r);                   // evaluate() calls updateCoeffs() if !updated()
    for (Patch& p : s.mesh.patches) {
        if (p.type == PROCESSOR) continue;      // coupled: value comes from the neighbour at use
        for (size_t i = 0; i < p.faceCells.size(); i++) {
            const scalar Tc = s.T[p.faceCells[i]];
            if (p.type == B200_BC_MIXED_TEMPERATURE) {
                const scalar f = p.valueFraction[i];
                p.Tb[i] = f * p.Tinf + (1.0 - f) * (Tc + 0.0 / p.deltaCoeffs[i]);
            } else if (p.type == B200_BC_FIXED_VALUE) p.Tb[i] = p.fixedValue;
            else p.Tb[i] = Tc;
        }
        p.updated = false;
    }
}

// setDeltaT.H:19-29 / solutionControls.H:36-46 without the trailing "*deltaT":
// [UPSTREAM] fvc::surfaceSum (internal faces owner+neighbour, then boundary faces),
// fvc::interpolate(kappa) = linear: lambda*(P - N) + N, boundary = patch value
inline scalar Case::surfaceSumDiMax()
{
    std::vector<scalar> pm(R, 0.0);
    forRanks(R, [&](int r) {
        RankState& s = rk[r];
        const Mesh& m = s.mesh;
        Field S(m.nCells(), 0.0);
        for (label f = 0; f < m.addr.nFaces; f++) {
            const label P = m.addr.lower[f], N = m.addr.upper[f];
            const scalar kf = 0.5 * (s.kappa[P] - s.kappa[N]) + s.kappa[N];
            const scalar v = m.magSf[f] * kf * m.deltaCoeffs[f];
            S[P] += v; S[N] += v;
        }
        for (const Patch& p : m.patches)
            for (size_t i = 0; i < p.faceCells.size(); i++) {
                scalar kb = s.kappa[p.faceCells[i]];
                if (p.type == PROCESSOR) {
                    const Interface& itf = m.interfaces[p.ifaceIndex];
                    kb = 0.5 * kb + (1.0 - 0.5) * rk[itf.neighbRank].kappa[itf.nbrCells[i]];
                }
                S[p.faceCells[i]] += p.magSf[i] * kb * p.deltaCoeffs[i];
            }
        scalar mx = -vGreat;
        for (label c = 0; c < m.nCells(); c++) mx = std::max(mx, S[c] / (m.V * d.material.rho * s.Cp[c]));
        pm[r] = mx;
    });
    scalar mx = -vGreat;
    for (scalar v : pm) mx = std::max(mx, v);
    return mx;
}

// SOLVER/setDeltaT.H:1-50 + [UPSTREAM] Time::setDeltaT -> Time::adjustDeltaT (write-time snapping)
// [UPSTREAM] GeometricField::oldTime(): a level that does not exist yet is created as a copy of the level above it
inline void Case::touchOldTimes(bool ofT, int levels)
{
    int& n = ofT ? TnOld : a1nOld;
    if (n < 2 && levels >= 2) forRanks(R, [&](int r) { if (ofT) rk[r].Toldold = rk[r].Told; else rk[r].alpha1oldold = rk[r].alpha1old; });
    n = std::max(n, levels);
}

inline scalar Case::cnPrepare(bool ofT, bool fvm)
{
    Ddt0State& st = ofT ? ddt0StateT : ddt0StateA;
    if (!st.exists) {
        forRanks(R, [&](int r) { (ofT ? rk[r].ddt0T : rk[r].ddt0A).assign(rk[r].mesh.nCells(), 0.0); });
        st.exists = true; st.startTimeIndex = timeIndex; st.timeIndex = timeIndex;
    }
    const scalar psi = d.controls.ocCoeff;
    const scalar rDtCoef = (timeIndex > st.startTimeIndex ? 1 + psi : 1) / deltaT;
    touchOldTimes(ofT, fvm ? 2 : 1);
    if (st.timeIndex != timeIndex) {
        touchOldTimes(ofT, 2);
        const scalar rDtCoef0 = (timeIndex > st.startTimeIndex + 1 ? 1 + psi : 1) / deltaT0;
        forRanks(R, [&](int r) {
            RankState& s = rk[r];
            Field& d0 = ofT ? s.ddt0T : s.ddt0A;
            const Field& o = ofT ? s.Told : s.alpha1old;
            const Field& oo = ofT ? s.Toldold : s.alpha1oldold;
            for (label c = 0; c < s.mesh.nCells(); c++) d0[c] = rDtCoef0 * (o[c] - oo[c]) - offCentre(d0[c]);
        });
    }
    st.timeIndex = timeIndex;
    return rDtCoef;
}

inline void Case::setDeltaT()
{
    const b200_controls& ct = d.controls;
    if (!ct.adjustTimeStep) return;
    std::vector<scalar> pm(R, 0.0);
    const scalar rDeltaT = 1.0 / deltaT;
    // setDeltaT.H:6-11 fvc::ddt(T) with the scheme of fvSchemes: Euler rDeltaT*(T - T.oldTime()), or backward
    const BackwardCoeffs kb = backwardCoeffs(TnOld);
    const bool cn = crankNicolson();
    const scalar rDtCoef = cn ? cnPrepare(true, false) : 0.0;
    if (!cn) touchOldTimes(true, backward() ? 2 : 1);
    forRanks(R, [&](int r) {
        RankState& s = rk[r];
        scalar mx = -vGreat;
        for (label c = 0; c < s.mesh.nCells(); c++) {
            const scalar ddtT = cn ? rDtCoef * (s.T[c] - s.Told[c]) - offCentre(s.ddt0T[c])
                              : backward() ? kb.rDeltaT * (kb.coefft * s.T[c] - kb.coefft0 * s.Told[c] + kb.coefft00 * s.Toldold[c])
                                           : rDeltaT * (s.T[c] - s.Told[c]);
            mx = std::max(mx, std::fabs(ddtT) / (Tliq - Tsol) * deltaT);
        }
        pm[r] = mx;
    });
    alphaCoNum = -vGreat;
    for (scalar v : pm) alphaCoNum = std::max(alphaCoNum, v);
    DiNum = surfaceSumDiMax() * deltaT;
    const scalar CoNum = 0.0;           // [UPSTREAM] CourantNo.H with phi == 0
    const scalar maxDeltaTFact = std::min(std::min(ct.maxCo / (CoNum + small_), ct.maxAlphaCo / (alphaCoNum + small_)),
                                          ct.maxDi / (DiNum + small_));
    const scalar deltaTFact = std::min(std::min(maxDeltaTFact, 1.0 + 0.1 * maxDeltaTFact), 1.2);
    scalar dt = std::min(deltaTFact * deltaT, ct.maxDeltaT);
    for (HeatSource& hs : sources) hs.beam.adjustDeltaT(dt, time);          // movingHeatSourceModel.C:92-98
    deltaT = dt;
    if (ct.writeInterval > 0) {         // [UPSTREAM] Time::adjustDeltaT, writeControl adjustableRunTime
        const scalar timeToNextWrite = std::max(0.0, (writeTimeIndex + 1) * ct.writeInterval - (time - ct.startTime));
        const scalar nSteps = timeToNextWrite / deltaT;
        if (nSteps < (scalar)INT32_MAX) {
            const label nStepsToNextWrite = label(std::max(nSteps, 1.0) + 0.99);
            const scalar newDeltaT = timeToNextWrite / nStepsToNextWrite;
            if (newDeltaT >= deltaT) deltaT = std::min(newDeltaT, 2.0 * deltaT);
            else deltaT = std::max(newDeltaT, 0.2 * deltaT);
        }
    }
}

// MHS/heatSourceModels/heatSourceModel/heatSourceModel.C:123-219
inline void Case::updateDimensions(HeatSource& hs)
{
    if (!hs.transient) return;
    const Vec3 position = hs.beam.position;
    const Vec3 sd = hs.shape.staticDimensions;
    const scalar searchRadius = std::max(sd.x, sd.y);
    std::vector<scalar> pm(R, sd.z);
    const scalar iso = hs.isoValue;
    forRanks(R, [&](int r) {
        const RankState& s = rk[r];
        const Mesh& m = s.mesh;
        scalar maxDepth = sd.z;
        auto probe = [&](scalar To, scalar Tn, Vec3 co, Vec3 cn) {
            const scalar minFace = std::min(To, Tn), maxFace = std::max(To, Tn);
            if (minFace < iso && maxFace >= iso) {
                const Vec3 dd = cn - co;
                Vec3 p = co + dd * (iso - To) / (Tn - To);
                p = cmptMag(p - position);
                const scalar pxy = std::sqrt(p.x * p.x + p.y * p.y);
                if (pxy <= searchRadius) maxDepth = std::max(p.z, maxDepth);
            }
        };
        for (label f = 0; f < m.addr.nFaces; f++) {
            const label own = m.addr.lower[f], nei = m.addr.upper[f];
            probe(s.T[own], s.T[nei], m.C(own), m.C(nei));
        }
        for (const Patch& p : m.patches) {
            if (p.type != PROCESSOR) continue;
            const Interface& itf = m.interfaces[p.ifaceIndex];
            const RankState& sn = rk[itf.neighbRank];
            for (size_t i = 0; i < p.faceCells.size(); i++)
                probe(s.T[p.faceCells[i]], sn.T[itf.nbrCells[i]], m.C(p.faceCells[i]), sn.mesh.C(itf.nbrCells[i]));
        }
        pm[r] = maxDepth;
    });
    scalar maxDepth = sd.z;
    for (scalar v : pm) maxDepth = std::max(maxDepth, v);
    hs.shape.dimensions = {sd.x, sd.y, maxDepth};
}

// MHS/heatSourceModels/heatSourceModel/heatSourceModel.C:221-371
inline void Case::heatSourceQDot(HeatSource& hs, Fields& out, scalar* sumWeightsOut, scalar* volumeOut, scalar* absorbedOut)
{
    out.resize(R);
    for (int r = 0; r < R; r++) out[r].assign(rk[r].mesh.nCells(), 0.0);
    if (sumWeightsOut) *sumWeightsOut = 0;
    if (volumeOut) *volumeOut = 0;
    if (absorbedOut) *absorbedOut = 0;
    const scalar power = hs.beam.power;
    if (!(power > small_)) return;
    const Vec3 position = hs.beam.position;
    const Vec3 dims = hs.shape.dimensions;
    const scalar aspectRatio = dims.z / std::min(dims.x, dims.y);
    const scalar absorbedPower = hs.absorption.eta(aspectRatio) * power;
    scalar volume = hs.shape.V0();
    const Box beamBb{position - 1.5 * dims, position + 1.5 * dims};
    std::vector<scalar> ps(R, 0.0);
    forRanks(R, [&](int r) {
        const Mesh& m = rk[r].mesh;
        Field& w = out[r];
        scalar sum = 0;
        for (label c = 0; c < m.nCells(); c++) {
            const label i = c % m.nx, j = (c / m.nx) % m.ny, k = c / (m.nx * m.ny);
            const Box cellBb{{m.px(i), m.py(j), m.pz(k)}, {m.px(i + 1), m.py(j + 1), m.pz(k + 1)}};
            if (cellBb.overlaps(beamBb)) w[c] = integrateCell(hs.shape, cellBb, beamBb, position, hs.nPoints, m.V);
            sum += m.V * w[c];          // fvc::domainIntegrate(weights)
        }
        ps[r] = sum;
    });
    const scalar sumWeights = gSum(ps);
    const scalar residual = sumWeights / volume;
    if (std::fabs(1 - residual) < 0.05) volume = sumWeights;
    forRanks(R, [&](int r) { for (scalar& v : out[r]) v = absorbedPower * v / volume; });
    if (sumWeightsOut) *sumWeightsOut = sumWeights;
    if (volumeOut) *volumeOut = volume;
    if (absorbedOut) *absorbedOut = absorbedPower;
}

// MHS/movingHeatSourceModel/movingHeatSourceModel.C:100-150
inline void Case::updateSources()
{
    forRanks(R, [&](int r) { std::fill(rk[r].qDot.begin(), rk[r].qDot.end(), 0.0); });
    for (HeatSource& hs : sources) {
        if (!hs.beam.activePath(time)) continue;
        updateDimensions(hs);
        scalar pathTime = time;
        const scalar nextTime = pathTime + deltaT;
        const scalar beam_dt = hs.beam.deltaT;
        Fields qDoti(R);
        for (int r = 0; r < R; r++) qDoti[r].assign(rk[r].mesh.nCells(), 0.0);
        scalar sumWeights = 0.0;
        Fields q;
        while ((nextTime - pathTime) > small_) {
            const scalar dt = std::min(beam_dt, std::max(0.0, nextTime - pathTime));
            pathTime += dt;
            hs.beam.move(pathTime);
            heatSourceQDot(hs, q);
            forRanks(R, [&](int r) { for (size_t c = 0; c < q[r].size(); c++) qDoti[r][c] += dt * q[r][c]; });
            sumWeights += dt;
        }
        forRanks(R, [&](int r) { for (size_t c = 0; c < qDoti[r].size(); c++) rk[r].qDot[c] += qDoti[r][c] / sumWeights; });
    }
}

// SOLVER/thermo/thermoScheme.H:17-39 + SURVEY.md Appendix A.2-A.4 (Euler ddt only).
inline void Case::assembleBase(bool explicitSolve)
{
    const scalar rDeltaT = 1.0 / deltaT;
    const scalar rho = d.material.rho;
    forRanks(R, [&](int r) { updateCoeffsT(r); });     // fvMatrix(T, ...) ctor -> T.boundaryFieldRef().updateCoeffs()
    // thermoScheme.H:25-39: the explicit form always uses EulerDdtScheme; the implicit one fvm::ddt(T) with the scheme of fvSchemes
    const bool bwd = backward() && !explicitSolve;
    const BackwardCoeffs kT = backwardCoeffs(TnOld);
    const bool cn = crankNicolson() && !explicitSolve;
    const scalar rDtCoef = cn ? cnPrepare(true, true) : 0.0;
    if (!cn) touchOldTimes(true, bwd ? 2 : 1);
    forRanks(R, [&](int r) {
        RankState& s = rk[r];
        const Mesh& m = s.mesh;
        const label n = m.nCells(), nf = m.addr.nFaces;
        // rho*Cp*(EulerDdtScheme::fvmDdt(T)): diag = rDeltaT*V, source = rDeltaT*Told*V, then *= rho*Cp
        s.teqn.diag.resize(n); s.source.resize(n);
        for (label c = 0; c < n; c++) {
            const scalar rhoCp = rho * s.Cp[c];
            if (cn) {       // [UPSTREAM] CrankNicolsonDdtScheme::fvmDdt: diag = rDtCoef*V, source = (rDtCoef*old + offCentre(ddt0))*V
                s.teqn.diag[c] = (rDtCoef * m.V) * rhoCp;
                s.source[c] = ((rDtCoef * s.Told[c] + offCentre(s.ddt0T[c])) * m.V) * rhoCp;
            } else if (bwd) {      // [UPSTREAM] backwardDdtScheme::fvmDdt: diag = (coefft*rDeltaT)*V, source = rDeltaT*V*(coefft0*old - coefft00*oldOld)
                s.teqn.diag[c] = ((kT.coefft * kT.rDeltaT) * m.V) * rhoCp;
                s.source[c] = (kT.rDeltaT * m.V * (kT.coefft0 * s.Told[c] - kT.coefft00 * s.Toldold[c])) * rhoCp;
            } else {
                s.teqn.diag[c] = (rDeltaT * m.V) * rhoCp;
                s.source[c] = (rDeltaT * s.Told[c] * m.V) * rhoCp;
            }
        }
        s.internalCoeffs.resize(m.patches.size()); s.boundaryCoeffs.resize(m.patches.size());
        s.teqn.bouCoeffs.assign(m.interfaces.size(), Field());
        // boundary kappa: harmonic::interpolate = 1/(reverseLinear(1/kappa)); non-coupled patch keeps the
        // patch value of 1/kappa, coupled patch uses pLambda*pif + (1-pLambda)*pnf
        auto kappaBoundary = [&](const Patch& p, size_t i) {
            const scalar rP = 1.0 / s.kappa[p.faceCells[i]];
            if (p.type == PROCESSOR) {
                const Interface& itf = m.interfaces[p.ifaceIndex];
                const scalar rN = 1.0 / rk[itf.neighbRank].kappa[itf.nbrCells[i]];
                return 1.0 / (0.5 * rP + (1.0 - 0.5) * rN);
            }
            return 1.0 / rP;
        };
        if (!explicitSolve) {
            // gaussLaplacianScheme::fvmLaplacianUncorrected: upper = deltaCoeffs*(gamma_f*magSf); negSumDiag
            s.teqn.upper.resize(nf); s.teqn.lower.clear();
            Field lapDiag(n, 0.0);
            for (label f = 0; f < nf; f++) {
                const label P = m.addr.lower[f], N = m.addr.upper[f];
                const scalar rP = 1.0 / s.kappa[P], rN = 1.0 / s.kappa[N];
                const scalar kf = 1.0 / (0.5 * (rP - rN) + rN);
                s.teqn.upper[f] = m.deltaCoeffs[f] * (kf * m.magSf[f]);
            }
            for (label f = 0; f < nf; f++) {
                lapDiag[m.addr.lower[f]] -= s.teqn.upper[f];
                lapDiag[m.addr.upper[f]] -= s.teqn.upper[f];
            }
            // TEqn = rhoCp*ddt - laplacian
            for (label c = 0; c < n; c++) s.teqn.diag[c] = s.teqn.diag[c] - lapDiag[c];
            for (label f = 0; f < nf; f++) s.teqn.upper[f] = 0.0 - s.teqn.upper[f];
            for (size_t pi = 0; pi < m.patches.size(); pi++) {
                const Patch& p = m.patches[pi];
                const size_t np = p.faceCells.size();
                s.internalCoeffs[pi].assign(np, 0.0); s.boundaryCoeffs[pi].assign(np, 0.0);
                for (size_t i = 0; i < np; i++) {
                    const scalar pGamma = kappaBoundary(p, i) * p.magSf[i];
                    scalar gic = 0, gbc = 0;        // gradientInternalCoeffs / gradientBoundaryCoeffs
                    if (p.type == B200_BC_FIXED_VALUE) { gic = -p.deltaCoeffs[i]; gbc = p.deltaCoeffs[i] * p.Tb[i]; }
                    else if (p.type == B200_BC_MIXED_TEMPERATURE) {
                        const scalar f = p.valueFraction[i];
                        gic = -f * p.deltaCoeffs[i];
                        gbc = f * p.deltaCoeffs[i] * p.Tinf + (1.0 - f) * 0.0;
                    } else if (p.type == PROCESSOR) { gic = -p.deltaCoeffs[i]; gbc = -gic; }
                    const scalar icLap = pGamma * gic, bcLap = -pGamma * gbc;
                    s.internalCoeffs[pi][i] = 0.0 - icLap;
                    s.boundaryCoeffs[pi][i] = 0.0 - bcLap;
                }
                if (p.type == PROCESSOR) s.teqn.bouCoeffs[p.ifaceIndex] = s.boundaryCoeffs[pi];
            }
        } else {
            // - fvc::laplacian(kappa, T): source += V * ( div( (kappa_f*snGrad)*magSf ) )
            s.teqn.upper.clear(); s.teqn.lower.clear();
            Field lap(n, 0.0);
            for (label f = 0; f < nf; f++) {
                const label P = m.addr.lower[f], N = m.addr.upper[f];
                const scalar rP = 1.0 / s.kappa[P], rN = 1.0 / s.kappa[N];
                const scalar kf = 1.0 / (0.5 * (rP - rN) + rN);
                const scalar snGrad = m.deltaCoeffs[f] * (s.T[N] - s.T[P]);
                const scalar flux = kf * snGrad * m.magSf[f];
                lap[P] += flux; lap[N] -= flux;
            }
            for (size_t pi = 0; pi < m.patches.size(); pi++) {
                const Patch& p = m.patches[pi];
                s.internalCoeffs[pi].assign(p.faceCells.size(), 0.0); s.boundaryCoeffs[pi].assign(p.faceCells.size(), 0.0);
                for (size_t i = 0; i < p.faceCells.size(); i++) {
                    const scalar Tc = s.T[p.faceCells[i]];
                    scalar snGrad = 0;
                    if (p.type == B200_BC_FIXED_VALUE) snGrad = p.deltaCoeffs[i] * (p.Tb[i] - Tc);
                    else if (p.type == B200_BC_MIXED_TEMPERATURE) {
                        const scalar f = p.valueFraction[i];
                        snGrad = f * (p.Tinf - Tc) * p.deltaCoeffs[i] + (1.0 - f) * 0.0;
                    } else if (p.type == PROCESSOR) {
                        const Interface& itf = m.interfaces[p.ifaceIndex];
                        snGrad = p.deltaCoeffs[i] * (rk[itf.neighbRank].T[itf.nbrCells[i]] - Tc);
                    }
                    lap[p.faceCells[i]] += kappaBoundary(p, i) * snGrad * p.magSf[i];
                }
            }
            for (label c = 0; c < n; c++) s.source[c] += m.V * (lap[c] / m.V);
        }
        // - sources.qDot()
        for (label c = 0; c < n; c++) s.source[c] += m.V * s.qDot[c];
    });
}

// SOLVER/thermo/thermoSource.H:7-66
inline void Case::thermoSource(int r)
{
    RankState& s = rk[r];
    const Field& xvals = thermoX; const Field& yvals = thermoY;
    const label n = (label)xvals.size();
    const scalar slope = 1e10;
    const scalar thermoTol = d.controls.thermoTolerance;
    for (label c = 0; c < s.mesh.nCells(); c++) {
        const scalar x = s.T[c], y = s.alpha1[c];
        if (x < Tliq && x > Tsol) {
            label lo, hi;
            interpolateXYLabels(x, xvals, lo, hi);      // identical scans, thermoSource.H:21-51
            (void)n;
            s.dFdT[c] = (yvals[hi] - yvals[lo]) / (xvals[hi] - xvals[lo]);
            s.T0[c] = xvals[lo] + (y - yvals[lo]) / s.dFdT[c];
        } else if (x >= Tliq && y > thermoTol) { s.dFdT[c] = slope; s.T0[c] = Tliq - thermoTol; }
        else if (x <= Tsol && y < 1 - thermoTol) { s.dFdT[c] = slope; s.T0[c] = Tsol + thermoTol; }
        else { s.dFdT[c] = 0.0; s.T0[c] = x; }
    }
}

// SOLVER/thermo/TEqn.H:24-42 + [UPSTREAM] fvMatrix<scalar>::solveSegregated (Appendix A.4)
inline SolverPerf Case::solveCorrector(bool explicitSolve, scalar rDeltaT, scalar rDeltaTEuler)
{
    const scalar rho = d.material.rho, Lf = d.material.Lf;
    // TEqn.H:36-39: Euler fvcDdt(alpha1) when ddtScheme == "Euler" (always so in the explicit form, thermoScheme.H:22), else fvc::ddt
    const bool bwdA = backward() && !explicitSolve;
    const BackwardCoeffs kA = backwardCoeffs(a1nOld);
    const bool cnA = crankNicolson() && !explicitSolve;
    const scalar rDtCoefA = cnA ? cnPrepare(false, false) : 0.0;
    if (!cnA) touchOldTimes(false, bwdA ? 2 : 1);
    const scalar Tmax = d.controls.Tmax > 0 ? d.controls.Tmax : vGreat;
    std::vector<LduMatrix> A(R);
    Fields totalSource(R), psi(R);
    forRanks(R, [&](int r) {
        RankState& s = rk[r];
        const Mesh& m = s.mesh;
        const label n = m.nCells();
        LduMatrix& M = A[r];
        M = s.teqn;                                     // copy of TEqn (tmp in "TEqn + ...")
        totalSource[r] = s.source;
        const scalar s2 = rDeltaT * rho * Lf;
        for (label c = 0; c < n; c++) {
            const scalar Acoef = 1e15 * ((s.T[c] - Tmax) >= 0 ? 1.0 : 0.0);     // TEqn.H:24-29, pos()
            const scalar diag1 = (m.V * Acoef) * rDeltaT;
            const scalar source1 = (m.V * (Acoef * Tmax)) * rDeltaT;
            const scalar diag2 = (m.V * s.dFdT[c]) * s2;
            const scalar source2 = (m.V * (s.dFdT[c] * s.T0[c])) * s2;
            const scalar su = cnA ? (rho * Lf) * (rDtCoefA * (s.alpha1[c] - s.alpha1old[c]) - offCentre(s.ddt0A[c]))
                            : bwdA ? (rho * Lf) * (kA.rDeltaT * (kA.coefft * s.alpha1[c] - kA.coefft0 * s.alpha1old[c] + kA.coefft00 * s.alpha1oldold[c]))
                                   : (rho * Lf) * (rDeltaTEuler * (s.alpha1[c] - s.alpha1old[c]));
            const scalar sourceR = source2 - m.V * su;
            M.diag[c] = (M.diag[c] + diag1) - diag2;
            totalSource[r][c] = (totalSource[r][c] + source1) - sourceR;
        }
        // addBoundaryDiag (all patches), addBoundarySource (non-coupled only)
        for (size_t pi = 0; pi < m.patches.size(); pi++) {
            const Patch& p = m.patches[pi];
            for (size_t i = 0; i < p.faceCells.size(); i++) M.diag[p.faceCells[i]] += s.internalCoeffs[pi][i];
        }
        for (size_t pi = 0; pi < m.patches.size(); pi++) {
            const Patch& p = m.patches[pi];
            if (p.type == PROCESSOR) continue;
            for (size_t i = 0; i < p.faceCells.size(); i++) totalSource[r][p.faceCells[i]] += s.boundaryCoeffs[pi][i];
        }
        psi[r] = s.T;
    });
    SolverPerf perf;
    ResidualLog log;
    SolverControls sc;
    sc.solver = d.controls.solver; sc.precond = d.controls.preconditioner; sc.tolerance = d.controls.tolerance;
    sc.relTol = d.controls.relTol; sc.minIter = d.controls.minIter; sc.maxIter = d.controls.maxIter;
    if (explicitSolve) perf = solveDiagonal(A, psi, totalSource);
    else if (sc.solver == B200_SOLVER_PBICGSTAB) perf = solvePBiCGStab(A, psi, totalSource, sc, &log);
    else perf = solvePCG(A, psi, totalSource, sc, &log);
    lastResidualLog = log.res;
    forRanks(R, [&](int r) { rk[r].T = psi[r]; });
    forRanks(R, [&](int r) { correctTBoundary(r); });
    return perf;
}

// SOLVER/functionObjects/meltPoolDimensions/meltPoolDimensions.C:123-294 (a "next" row): bounding box of each
// isotherm, located linearly between cell centres across internal and processor faces, plus the face centres of physical
// boundary faces whose cell or patch value reaches the isovalue; rotated by the scan path angle.
inline void Case::meltPoolDimensions(const std::vector<scalar>& isoValues, scalar scanPathAngleDeg, std::vector<Vec3>& dims)
{
    const scalar radians = scanPathAngleDeg * (pi_ / 180.0);
    const scalar sn = std::sin(radians), cs = std::cos(radians);
    const size_t nIso = isoValues.size();
    std::vector<Vec3> bmin(nIso, Vec3{vGreat, vGreat, vGreat}), bmax(nIso, Vec3{-vGreat, -vGreat, -vGreat});
    auto add = [&](size_t i, Vec3 p) {
        const Vec3 q{p.x * cs + p.y * sn, p.y * cs - p.x * sn, p.z};
        bmin[i] = {std::min(q.x, bmin[i].x), std::min(q.y, bmin[i].y), std::min(q.z, bmin[i].z)};
        bmax[i] = {std::max(q.x, bmax[i].x), std::max(q.y, bmax[i].y), std::max(q.z, bmax[i].z)};
    };
    for (int r = 0; r < R; r++) {
        const RankState& s = rk[r];
        const Mesh& m = s.mesh;
        auto cross = [&](scalar To, scalar Tn, Vec3 co, Vec3 cn) {
            const scalar minFace = std::min(To, Tn), maxFace = std::max(To, Tn);
            for (size_t i = 0; i < nIso; i++) {
                const scalar iso = isoValues[i];
                if (minFace < iso && maxFace >= iso) add(i, co + (cn - co) * (iso - To) / (Tn - To));
            }
        };
        for (label f = 0; f < m.addr.nFaces; f++) {
            const label own = m.addr.lower[f], nei = m.addr.upper[f];
            cross(s.T[own], s.T[nei], m.C(own), m.C(nei));
        }
        for (const Patch& p : m.patches) {
            if (p.type == PROCESSOR) {
                const Interface& itf = m.interfaces[p.ifaceIndex];
                const RankState& sn_ = rk[itf.neighbRank];
                for (size_t i = 0; i < p.faceCells.size(); i++)
                    cross(s.T[p.faceCells[i]], sn_.T[itf.nbrCells[i]], m.C(p.faceCells[i]), sn_.mesh.C(itf.nbrCells[i]));
            }
        }
        // physical patches: the face centre where max(cell value, patch value) >= iso.  Faces are enumerated per block side.
        for (int pi = 0; pi < d.nPatches; pi++) {
            const Patch& p = m.patches[pi];
            size_t fi = 0;
            for (int q = 0; q < d.patches[pi].nSides; q++) {
                const int side = d.patches[pi].sides[q];
                auto visit = [&](label i, label j, label k, Vec3 cf) {
                    const label c = i + m.nx * (j + m.ny * k);
                    const scalar Tb = p.type == B200_BC_ZERO_GRADIENT ? s.T[c] : p.Tb[fi];
                    const scalar maxFace = std::max(s.T[c], Tb);
                    for (size_t is = 0; is < nIso; is++) if (maxFace >= isoValues[is]) add(is, cf);
                    fi++;
                };
                if (side == B200_XMIN || side == B200_XMAX) {
                    const label i = side == B200_XMIN ? 0 : m.nx - 1;
                    for (label k = 0; k < m.nz; k++) for (label j = 0; j < m.ny; j++) {
                        Vec3 cf = m.C(i + m.nx * (j + m.ny * k)); cf.x = side == B200_XMIN ? m.px(0) : m.px(m.nx);
                        visit(i, j, k, cf);
                    }
                } else if (side == B200_YMIN || side == B200_YMAX) {
                    const label j = side == B200_YMIN ? 0 : m.ny - 1;
                    for (label i = 0; i < m.nx; i++) for (label k = 0; k < m.nz; k++) {
                        Vec3 cf = m.C(i + m.nx * (j + m.ny * k)); cf.y = side == B200_YMIN ? m.py(0) : m.py(m.ny);
                        visit(i, j, k, cf);
                    }
                } else {
                    const bool mine = side == B200_ZMIN ? (r == 0) : (r == R - 1);
                    if (!mine) continue;
                    const label k = side == B200_ZMIN ? 0 : m.nz - 1;
                    for (label i = 0; i < m.nx; i++) for (label j = 0; j < m.ny; j++) {
                        Vec3 cf = m.C(i + m.nx * (j + m.ny * k)); cf.z = side == B200_ZMIN ? m.pz(0) : m.pz(m.nz);
                        visit(i, j, k, cf);
                    }
                }
            }
        }
    }
    dims.resize(nIso);
    for (size_t i = 0; i < nIso; i++)
        dims[i] = {std::max(bmax[i].x - bmin[i].x, 0.0), std::max(bmax[i].y - bmin[i].y, 0.0), std::max(bmax[i].z - bmin[i].z, 0.0)};
}

// SOLVER/functionObjects/solidificationData/solidificationData.C:55-83 (constructor: R_ = fvc::ddt(T) = 0 before the first step)
// and :99-131 (setOverlapCells: cells whose bounding box, grown by 1e-10, overlaps the box; inclusive comparisons as
// [UPSTREAM] boundBox::overlaps)
inline void Case::solidificationEnable(Vec3 boxMin, Vec3 boxMax, scalar isoValue)
{
    solidOn = true; solidBoxMin = boxMin; solidBoxMax = boxMax; solidIso = isoValue;
    solidR.assign(R, Field()); solidCells.assign(R, {}); solidEvents.assign(R, {});
    const scalar ext = 1e-10;
    for (int r = 0; r < R; r++) {
        const Mesh& m = rk[r].mesh;
        solidR[r].assign(m.nCells(), 0.0);
        for (label c = 0; c < m.nCells(); c++) {
            const label i = c % m.nx, j = (c / m.nx) % m.ny, k = c / (m.nx * m.ny);
            const Vec3 lo{m.px(i) - ext, m.py(j) - ext, m.pz(k) - ext}, hi{m.px(i + 1) + ext, m.py(j + 1) + ext, m.pz(k + 1) + ext};
            if (lo.x <= boxMax.x && hi.x >= boxMin.x && lo.y <= boxMax.y && hi.y >= boxMin.y && lo.z <= boxMax.z && hi.z >= boxMin.z)
                solidCells[r].push_back(c);
        }
    }
}

// solidificationData.C:140-187, run after each completed time step ([UPSTREAM] Time::run -> functionObjectList::execute):
// G = mag(fvc::grad(T)) with "Gauss linear" ([UPSTREAM] gaussGrad: sum of Sf*T_f over the faces / V; linear T_f between
// cells, patch value on boundary faces, the coupled average on processor faces); a cell that was above the isovalue at the
// old time and is at or below it now is recorded with the cooling rate R_ kept from the PREVIOUS execute; only then is R_
// refreshed to fvc::ddt(T) = (T - T.oldTime())/deltaT (Euler).
inline void Case::solidificationExecute()
{
    const scalar rDeltaT = 1.0 / deltaT;
    for (int r = 0; r < R; r++) {
        RankState& s = rk[r];
        const Mesh& m = s.mesh;
        const label nx = m.nx, ny = m.ny, nz = m.nz, nxny = nx * ny;
        // boundary value lookup: physical patch faces by side, processor neighbours by plane
        std::vector<const scalar*> sideTb(6, nullptr);
        for (int pi = 0; pi < d.nPatches; pi++) {
            size_t off = 0;
            for (int q = 0; q < d.patches[pi].nSides; q++) {
                const int side = d.patches[pi].sides[q];
                size_t cnt;
                if (side <= B200_XMAX) cnt = (size_t)ny * nz;
                else if (side <= B200_YMAX) cnt = (size_t)nx * nz;
                else { const bool mine = side == B200_ZMIN ? (r == 0) : (r == R - 1); cnt = mine ? (size_t)nx * ny : 0; }
                if (cnt) sideTb[side] = m.patches[pi].Tb.data() + off;
                off += cnt;
            }
        }
        auto faceValue = [&](int side, label i, label j, label k, scalar Tc) -> scalar {
            if (side >= B200_ZMIN) {
                const int nb = side == B200_ZMIN ? r - 1 : r + 1;
                if (nb >= 0 && nb < R) {                         // coupled face ([UPSTREAM] surfaceInterpolationScheme::interpolate): w*own + (1-w)*nbr
                    const RankState& o = rk[nb];
                    const scalar Tn = o.T[i + nx * (j + ny * (side == B200_ZMIN ? o.mesh.nz - 1 : 0))];
                    return 0.5 * Tc + (1.0 - 0.5) * Tn;
                }
            }
            if (!sideTb[side]) return Tc;                        // side not listed in any patch (treated as zeroGradient)
            const size_t fi = side <= B200_XMAX ? (size_t)j + (size_t)ny * k : side <= B200_YMAX ? (size_t)k + (size_t)nz * i : (size_t)j + (size_t)ny * i;
            return sideTb[side][fi];
        };
        for (label c : solidCells[r]) {
            if (s.Told[c] > solidIso && s.T[c] <= solidIso) {
                const label i = c % nx, j = (c / nx) % ny, k = c / nxny;
                const scalar Tc = s.T[c];
                const scalar sfx = m.dy * m.dz, sfy = m.dx * m.dz, sfz = m.dx * m.dy;
                const scalar Tw = i > 0 ? 0.5 * (s.T[c - 1] - Tc) + Tc : faceValue(B200_XMIN, i, j, k, Tc);
                const scalar Te = i < nx - 1 ? 0.5 * (Tc - s.T[c + 1]) + s.T[c + 1] : faceValue(B200_XMAX, i, j, k, Tc);
                const scalar Ts = j > 0 ? 0.5 * (s.T[c - nx] - Tc) + Tc : faceValue(B200_YMIN, i, j, k, Tc);
                const scalar Tn = j < ny - 1 ? 0.5 * (Tc - s.T[c + nx]) + s.T[c + nx] : faceValue(B200_YMAX, i, j, k, Tc);
                const scalar Tb = k > 0 ? 0.5 * (s.T[c - nxny] - Tc) + Tc : faceValue(B200_ZMIN, i, j, k, Tc);
                const scalar Tt = k < nz - 1 ? 0.5 * (Tc - s.T[c + nxny]) + s.T[c + nxny] : faceValue(B200_ZMAX, i, j, k, Tc);
                const scalar gx = (sfx * Te - sfx * Tw) / m.V, gy = (sfy * Tn - sfy * Ts) / m.V, gz = (sfz * Tt - sfz * Tb) / m.V;
                const scalar G = std::sqrt(gx * gx + gy * gy + gz * gz);
                const scalar Ri = std::fabs(solidR[r][c]), Gi = std::max(G, small_);
                const Vec3 pt = m.C(c);
                solidEvents[r].push_back({pt.x, pt.y, pt.z, time, Ri, Gi, Ri / Gi});
            }
        }
        for (label c = 0; c < m.nCells(); c++) solidR[r][c] = rDeltaT * (s.T[c] - s.Told[c]);
    }
}

// One pass of the time loop body, SOLVER/additiveFoam.C:66-95
inline void Case::step(b200_step_report* rep)
{
    const b200_controls& ct = d.controls;
    updateProperties();                                 // :68
    setDeltaT();                                        // :72
    updateSources();                                    // :74
    // runTime++ (:76): store old-time fields, advance time
    // ([UPSTREAM] GeometricField::storeOldTimes shifts the levels that exist; a level created later is a copy of the one above it at
    // that moment, see touchOldTimes)
    forRanks(R, [&](int r) {
        RankState& s = rk[r];
        if (TnOld >= 2) s.Toldold = s.Told;
        if (a1nOld >= 2) s.alpha1oldold = s.alpha1old;
        s.Told = s.T; s.alpha1old = s.alpha1;
    });
    deltaT0 = deltaTSave; deltaTSave = deltaT;          // [UPSTREAM] Time::operator++
    time += deltaT; timeIndex++;
    if (crankNicolson()) {
        cnSnap.T = ddt0StateT; cnSnap.A = ddt0StateA; cnSnap.TnOld = TnOld; cnSnap.a1nOld = a1nOld; cnSnap.timeIndex = timeIndex;
        cnSnap.ddt0T.assign(R, Field()); cnSnap.ddt0A.assign(R, Field());
        for (int r = 0; r < R; r++) { cnSnap.ddt0T[r] = rk[r].ddt0T; cnSnap.ddt0A[r] = rk[r].ddt0A; }
    }
    if (ct.writeInterval > 0) {                         // [UPSTREAM] Time::operator++ write-time bookkeeping
        const label wi = label(((time - ct.startTime) + 0.5 * deltaT) / ct.writeInterval);
        if (wi > writeTimeIndex) writeTimeIndex = wi;
    }
    // solutionControls.H
    scalar nThermoCorr = ct.nThermoCorrectors;
    const scalar thermoTol = ct.thermoTolerance;
    std::vector<scalar> pp(R), pmin(R);
    forRanks(R, [&](int r) {
        const RankState& s = rk[r];
        scalar sum = 0, mn = vGreat;
        for (label c = 0; c < s.mesh.nCells(); c++) { sum += s.mesh.V * s.qDot[c]; mn = std::min(mn, s.alpha1[c]); }
        pp[r] = sum; pmin[r] = mn;
    });
    const scalar totalPower = gSum(pp);
    scalar minAlpha = vGreat; for (scalar v : pmin) minAlpha = std::min(minAlpha, v);
    const bool beamOn = totalPower > small_;
    const bool fluidInDomain = (1 - minAlpha) > small_;
    if (!(beamOn || fluidInDomain)) nThermoCorr = 1;
    bool explicitSolve = ct.explicitSolve != 0;
    if (explicitSolve) explicitSolve = (surfaceSumDiMax() * deltaT < 1.0);
    // thermo/TEqn.H
    assembleBase(explicitSolve);
    // thermoScheme.H:41-68: rDeltaT of the Sp terms is scaled by coefft = 1 + deltaT/(deltaT + deltaT0) for backward -- with
    // runTime.deltaT0() itself, not the scheme's guarded one -- in the implicit form only
    scalar rDeltaT = 1.0 / deltaT;
    if (backward() && !explicitSolve) rDeltaT *= 1.0 + deltaT / (deltaT + deltaT0);
    if (crankNicolson() && !explicitSolve) rDeltaT *= 1.0 + ct.ocCoeff;      // thermoScheme.H:51-65: coefft = 1 + ocCoeff, at every step
    forRanks(R, [&](int r) { std::fill(rk[r].dFdT.begin(), rk[r].dFdT.end(), 0.0); rk[r].T0 = rk[r].T; });
    int nCorrDone = 0, iterSum = 0;
    scalar residual = 0;
    SolverPerf perf;
    for (int tCorr = 0; tCorr < nThermoCorr; tCorr++) {
        forRanks(R, [&](int r) { thermoSource(r); });
        perf = solveCorrector(explicitSolve, rDeltaT, 1.0 / deltaT);
        iterSum += perf.nIterations;
        // TEqn.H:44 T.correctBoundaryConditions() -- a SECOND evaluate after the one at the end of solveSegregated: the patches are
        // no longer "updated", so mixedTemperature recomputes its value fraction from the face temperatures the first evaluate
        // left and evaluates again.  (Found by including TEqn.H verbatim, oracle/_ref/refTEqn: without this the internal field of
        // the step is the same, but the face values and value fractions that enter the NEXT step's matrix are not.)
        forRanks(R, [&](int r) { correctTBoundary(r); });
        std::vector<scalar> pr(R, 0.0);
        forRanks(R, [&](int r) {
            RankState& s = rk[r];
            scalar mx = -vGreat;
            for (label c = 0; c < s.mesh.nCells(); c++) {
                const scalar a10 = s.alpha1[c];
                s.alpha1[c] = std::min(std::max(a10 + s.dFdT[c] * (s.T[c] - s.T0[c]), 0.0), 1.0);
                mx = std::max(mx, std::fabs(s.alpha1[c] - a10));
            }
            pr[r] = mx;
        });
        residual = -vGreat; for (scalar v : pr) residual = std::max(residual, v);
        nCorrDone = tCorr + 1;
        if (verbose) std::printf("Thermo: iteration %d residual: %g\n", tCorr, residual);
        if (residual < thermoTol && tCorr > 0) break;
    }
    if (rep) {
        rep->time = time; rep->deltaT = deltaT; rep->alphaCoNum = alphaCoNum; rep->DiNum = DiNum;
        rep->totalPower = totalPower; rep->explicitStep = explicitSolve ? 1 : 0; rep->nThermoCorr = nCorrDone;
        rep->thermoResidual = residual; rep->nSolveIterations = iterSum; rep->lastSolveIterations = perf.nIterations;
        rep->lastInitialResidual = perf.initialResidual; rep->lastFinalResidual = perf.finalResidual;
    }
    if (solidOn) solidificationExecute();
}

} // namespace afo
