// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
// C entry points of the CPU oracle for ctypes (tests/, smoke(), bench.py cpu_baseline only).
// Parity unpinned: see af_ldu.hpp.
#include "af_thermal.hpp"
#include <cstring>
#include <chrono>

using namespace afo;

namespace {
struct LduProblem {
    LduAddr addr;
    std::vector<LduMatrix> A;
    void set(label n, label nf, const label* lo, const label* up, const scalar* diag, const scalar* upper, const scalar* lower)
    {
        addr.nCells = n; addr.nFaces = nf;
        addr.lower.assign(lo, lo + nf); addr.upper.assign(up, up + nf);
        addr.finalize();
        A.resize(1);
        A[0].addr = &addr;
        A[0].diag.assign(diag, diag + n);
        A[0].upper.assign(upper, upper + nf);
        if (lower) A[0].lower.assign(lower, lower + nf);
    }
};
}

extern "C" {

int afo_ldu_amul(label n, label nf, const label* lo, const label* up, const scalar* diag, const scalar* upper,
                 const scalar* lower, const scalar* psi, scalar* out)
{
    LduProblem P; P.set(n, nf, lo, up, diag, upper, lower);
    Fields x(1); x[0].assign(psi, psi + n);
    Field y(n);
    Amul(P.A[0], x, 0, y);
    std::memcpy(out, y.data(), sizeof(scalar) * n);
    return 0;
}

int afo_ldu_precondition(label n, label nf, const label* lo, const label* up, const scalar* diag, const scalar* upper,
                         const scalar* lower, int kind, const scalar* rA, scalar* wA, scalar* rDout)
{
    LduProblem P; P.set(n, nf, lo, up, diag, upper, lower);
    Preconditioner M; M.build(P.A[0], kind);
    Field r(rA, rA + n), w(n);
    M.precondition(w, r);
    std::memcpy(wA, w.data(), sizeof(scalar) * n);
    if (rDout && kind != 0) std::memcpy(rDout, M.rD.data(), sizeof(scalar) * n);
    return 0;
}

int afo_ldu_solve(label n, label nf, const label* lo, const label* up, const scalar* diag, const scalar* upper,
                  const scalar* lower, const scalar* source, scalar* psi, int solver, int precond, scalar tol,
                  scalar relTol, label minIter, label maxIter, b200_solver_perf* perf, scalar* resLog, label maxLog,
                  label* nLog)
{
    LduProblem P; P.set(n, nf, lo, up, diag, upper, lower);
    Fields x(1), b(1); x[0].assign(psi, psi + n); b[0].assign(source, source + n);
    SolverControls sc; sc.solver = solver; sc.precond = precond; sc.tolerance = tol; sc.relTol = relTol;
    sc.minIter = minIter; sc.maxIter = maxIter;
    ResidualLog log;
    SolverPerf sp;
    if (solver == B200_SOLVER_DIAGONAL) sp = solveDiagonal(P.A, x, b);
    else if (solver == B200_SOLVER_PBICGSTAB) sp = solvePBiCGStab(P.A, x, b, sc, &log);
    else sp = solvePCG(P.A, x, b, sc, &log);
    std::memcpy(psi, x[0].data(), sizeof(scalar) * n);
    if (perf) {
        perf->initialResidual = sp.initialResidual; perf->finalResidual = sp.finalResidual;
        perf->nIterations = sp.nIterations; perf->converged = sp.converged; perf->singular = sp.singular;
    }
    if (resLog && nLog) {
        *nLog = (label)std::min<size_t>(log.res.size(), (size_t)maxLog);
        for (label i = 0; i < *nLog; i++) resLog[i] = log.res[i];
    }
    return 0;
}

int afo_ldu_levels(label n, label nf, const label* lo, const label* up, label* levelOfCell, label* nLevels)
{
    LduAddr ad; ad.nCells = n; ad.nFaces = nf; ad.lower.assign(lo, lo + nf); ad.upper.assign(up, up + nf);
    std::vector<label> lv;
    *nLevels = levelSchedule(ad, lv);
    std::memcpy(levelOfCell, lv.data(), sizeof(label) * n);
    return 0;
}

// block mesh addressing for rank r of R z-slabs
int afo_block_addressing(const b200_block_mesh* bm, int r, int R, label* nCells, label* nFaces, label* lower, label* upper)
{
    b200_case_desc d{}; d.mesh = *bm; d.nPatches = 0;
    Mesh m; buildMesh(m, d, r, R);
    *nCells = m.addr.nCells; *nFaces = m.addr.nFaces;
    if (lower) std::memcpy(lower, m.addr.lower.data(), sizeof(label) * m.addr.nFaces);
    if (upper) std::memcpy(upper, m.addr.upper.data(), sizeof(label) * m.addr.nFaces);
    return 0;
}

// ---- beam pieces
static void fillSource(HeatSource& hs, const b200_beam* bs, scalar start, scalar end)
{
    hs.shape.model = bs->heatSourceModel; hs.shape.k = bs->k; hs.shape.m = bs->m; hs.shape.A = bs->A; hs.shape.B = bs->B;
    hs.shape.dimensions = {bs->dimensions[0], bs->dimensions[1], bs->dimensions[2]};
    hs.shape.staticDimensions = hs.shape.dimensions;
    if (hs.shape.model == B200_PROJECTED_GAUSSIAN) hs.shape.updateProjectedK();
    for (int q = 0; q < 3; q++) hs.nPoints[q] = bs->nPoints[q];
    hs.transient = bs->transient != 0; hs.isoValue = bs->isoValue;
    hs.absorption.model = bs->absorptionModel; hs.absorption.etaConst = bs->eta;
    hs.absorption.geometry = bs->kellyGeometry; hs.absorption.eta0 = bs->eta0; hs.absorption.etaMin = bs->etaMin;
    if (bs->path && bs->nSegments > 0) hs.beam.init(bs->path, bs->nSegments, bs->deltaT, bs->hitPathIntervals != 0, start, end);
}

scalar afo_absorption_eta(const b200_beam* bs, scalar aspectRatio)
{
    HeatSource hs; fillSource(hs, bs, 0, 0);
    return hs.absorption.eta(aspectRatio);
}
scalar afo_shape_V0(const b200_beam* bs, const scalar dims[3])
{
    HeatSource hs; fillSource(hs, bs, 0, 0);
    hs.shape.dimensions = {dims[0], dims[1], dims[2]};
    return hs.shape.V0();
}
scalar afo_shape_weight(const b200_beam* bs, const scalar dims[3], const scalar dvec[3])
{
    HeatSource hs; fillSource(hs, bs, 0, 0);
    hs.shape.dimensions = {dims[0], dims[1], dims[2]};
    hs.shape.V0();
    return hs.shape.weight({dvec[0], dvec[1], dvec[2]});
}
// movingBeam: move through the listed times in order (index_ carries over as in the reference)
int afo_beam_move(const b200_beam* bs, scalar start, scalar end, label nTimes, const scalar* times, scalar* posOut,
                  scalar* powerOut, label* indexOut, scalar* endTimeOut)
{
    HeatSource hs; fillSource(hs, bs, start, end);
    for (label i = 0; i < nTimes; i++) {
        hs.beam.move(times[i]);
        posOut[3 * i] = hs.beam.position.x; posOut[3 * i + 1] = hs.beam.position.y; posOut[3 * i + 2] = hs.beam.position.z;
        powerOut[i] = hs.beam.power; indexOut[i] = hs.beam.index;
    }
    if (endTimeOut) *endTimeOut = hs.beam.endTime;
    return 0;
}
scalar afo_beam_adjust_deltaT(const b200_beam* bs, scalar start, scalar end, scalar moveTo, scalar now, scalar dt)
{
    HeatSource hs; fillSource(hs, bs, start, end);
    if (moveTo >= 0) hs.beam.move(moveTo);
    hs.beam.adjustDeltaT(dt, now);
    return dt;
}
int afo_scan_path_parse(const char* text, scalar* out, label maxSeg) { return parseScanPath(text, out, maxSeg); }
scalar afo_interpolateXY(scalar x, label n, const scalar* xs, const scalar* ys)
{
    Field X(xs, xs + n), Y(ys, ys + n);
    return interpolateXY(x, X, Y);
}

int afo_beam_qdot(const b200_block_mesh* bm, const b200_beam* bs, const scalar position[3], const scalar dims[3],
                  scalar power, scalar* qDot, scalar* sumWeights, scalar* volumeUsed, scalar* absorbedPower)
{
    Case cs;
    cs.R = 1; cs.rk.resize(1);
    b200_case_desc d{}; d.mesh = *bm; d.nPatches = 0;
    buildMesh(cs.rk[0].mesh, d, 0, 1);
    HeatSource hs; fillSource(hs, bs, 0, 0);
    hs.shape.dimensions = {dims[0], dims[1], dims[2]};
    hs.beam.position = {position[0], position[1], position[2]};
    hs.beam.power = power;
    Fields q;
    cs.heatSourceQDot(hs, q, sumWeights, volumeUsed, absorbedPower);
    std::memcpy(qDot, q[0].data(), sizeof(scalar) * q[0].size());
    return 0;
}

// ---- full case
void* afo_case_create(const b200_case_desc* d, int nRanks)
{
    try {
        Case* c = new Case();
        c->init(*d, nRanks);
        return c;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "afo_case_create: %s\n", e.what());
        return nullptr;
    }
}
void afo_case_destroy(void* h) { delete (Case*)h; }
int afo_case_running(void* h) { return ((Case*)h)->running() ? 1 : 0; }
int afo_case_step(void* h, b200_step_report* rep) { ((Case*)h)->step(rep); return 0; }
int64_t afo_case_n_cells(void* h)
{
    int64_t n = 0;
    for (auto& s : ((Case*)h)->rk) n += s.mesh.nCells();
    return n;
}
static Field* fieldOf(RankState& s, int f)
{
    switch (f) {
    case B200_FIELD_T: return &s.T;
    case B200_FIELD_ALPHA_SOLID: return &s.alpha1;
    case B200_FIELD_ALPHA_POWDER: return &s.alpha3;
    case B200_FIELD_KAPPA: return &s.kappa;
    case B200_FIELD_CP: return &s.Cp;
    case B200_FIELD_QDOT: return &s.qDot;
    case B200_FIELD_DFDT: return &s.dFdT;
    case B200_FIELD_T0: return &s.T0;
    case B200_FIELD_T_OLD: return &s.Told;
    case B200_FIELD_ALPHA_SOLID_OLD: return &s.alpha1old;
    case B200_FIELD_T_OLD_OLD: return &s.Toldold;
    case B200_FIELD_ALPHA_SOLID_OLD_OLD: return &s.alpha1oldold;
    case B200_FIELD_DDT0_T: return &s.ddt0T;
    case B200_FIELD_DDT0_ALPHA_SOLID: return &s.ddt0A;
    }
    return nullptr;
}
int afo_case_get_field(void* h, int f, scalar* out)
{
    size_t off = 0;
    for (auto& s : ((Case*)h)->rk) {
        Field* F = fieldOf(s, f);
        if (!F) return 1;
        if (F->empty()) F->assign(s.mesh.nCells(), 0.0);          // ddt0 before the scheme's first use
        std::memcpy(out + off, F->data(), sizeof(scalar) * F->size());
        off += F->size();
    }
    return 0;
}
int afo_case_set_field(void* h, int f, const scalar* in)
{
    Case* c = (Case*)h;
    size_t off = 0;
    for (int r = 0; r < c->R; r++) {
        RankState& s = c->rk[r];
        Field* F = fieldOf(s, f);
        if (!F) return 1;
        if (F->empty()) F->assign(s.mesh.nCells(), 0.0);
        std::memcpy(F->data(), in + off, sizeof(scalar) * F->size());
        off += F->size();
        if (f == B200_FIELD_T) { s.Told = s.T; c->correctTBoundary(r); }
        if (f == B200_FIELD_ALPHA_SOLID) s.alpha1old = s.alpha1;
    }
    return 0;
}
// Restart (createFields.H:31-44,74-78,112): alpha.solid and its old-time level from T through the thermoPath table, then updateProperties
int afo_case_reset_alpha_solid(void* h)
{
    Case* c = (Case*)h;
    for (RankState& s : c->rk)
        for (label i = 0; i < s.mesh.nCells(); i++) {
            s.alpha1[i] = std::min(std::max(interpolateXY(s.T[i], c->thermoX, c->thermoY), 0.0), 1.0);
            s.alpha1old[i] = s.alpha1[i];
        }
    c->updateProperties();
    return 0;
}
int afo_case_residual_log(void* h, scalar* out, label maxLog)
{
    Case* c = (Case*)h;
    const label n = (label)std::min<size_t>(c->lastResidualLog.size(), (size_t)maxLog);
    for (label i = 0; i < n; i++) out[i] = c->lastResidualLog[i];
    return n;
}
// The assembled corrector matrix of rank 0 for the current state (parity probe for the assembly kernels):
// runs updateProperties + assembleBase(implicit) + thermoSource and returns diag/upper/source with boundary folded.
int afo_case_assemble_probe(void* h, int explicitSolve, scalar* diag, scalar* upper, scalar* source)
{
    Case* c = (Case*)h;
    if (c->R != 1) return 1;
    c->updateProperties();
    c->assembleBase(explicitSolve != 0);
    RankState& s = c->rk[0];
    std::memcpy(diag, s.teqn.diag.data(), sizeof(scalar) * s.teqn.diag.size());
    if (upper && !s.teqn.upper.empty()) std::memcpy(upper, s.teqn.upper.data(), sizeof(scalar) * s.teqn.upper.size());
    std::memcpy(source, s.source.data(), sizeof(scalar) * s.source.size());
    return 0;
}
int afo_case_melt_pool_dimensions(void* h, label nIso, const scalar* iso, scalar angleDeg, scalar* dims)
{
    std::vector<scalar> v(iso, iso + nIso);
    std::vector<Vec3> out;
    ((Case*)h)->meltPoolDimensions(v, angleDeg, out);
    for (label i = 0; i < nIso; i++) { dims[3 * i] = out[i].x; dims[3 * i + 1] = out[i].y; dims[3 * i + 2] = out[i].z; }
    return 0;
}
int afo_case_solidification_enable(void* h, const scalar boxMin[3], const scalar boxMax[3], scalar isoValue)
{
    ((Case*)h)->solidificationEnable({boxMin[0], boxMin[1], boxMin[2]}, {boxMax[0], boxMax[1], boxMax[2]}, isoValue);
    return 0;
}
// events of all emulated ranks, rank by rank in append order (the order of the reference's data_<proc>.csv files)
int64_t afo_case_solidification_events(void* h, scalar* out, int64_t maxEvents)
{
    Case* c = (Case*)h;
    int64_t n = 0;
    for (auto& ev : c->solidEvents)
        for (auto& e : ev) {
            if (out && n < maxEvents) std::memcpy(out + 7 * n, e.data(), sizeof(scalar) * 7);
            n++;
        }
    return n;
}
scalar afo_case_time(void* h) { return ((Case*)h)->time; }
scalar afo_case_Tliq(void* h) { return ((Case*)h)->Tliq; }
scalar afo_case_Tsol(void* h) { return ((Case*)h)->Tsol; }
// heatSourceModel::updateDimensions (heatSourceModel.C:123-219) of beam b on the case's current T with the beam centre at `position`:
// the resulting dimensions (x, y static; z the isotherm depth).  The beam's own state is left as it was.
int afo_case_isotherm_depth(void* h, int b, const scalar position[3], scalar dimsOut[3])
{
    Case* c = (Case*)h;
    if (b < 0 || b >= (int)c->sources.size()) return 1;
    HeatSource hs = c->sources[b];
    hs.beam.position = {position[0], position[1], position[2]};
    c->updateDimensions(hs);
    dimsOut[0] = hs.shape.dimensions.x; dimsOut[1] = hs.shape.dimensions.y; dimsOut[2] = hs.shape.dimensions.z;
    return 0;
}

// Test probe: the state the control fragments of the NEXT step will see -- updateProperties() (additiveFoam.C:68; idempotent, the
// step repeats it) so that kappa and Cp are current, and T.oldTime().  Single-rank cases only.
int afo_case_pre_step_state(void* h, scalar* Told)
{
    Case* c = (Case*)h;
    if (c->R != 1) return 1;
    c->updateProperties();
    std::copy(c->rk[0].Told.begin(), c->rk[0].Told.end(), Told);
    return 0;
}
// Test probes: the patch fields of T (single-rank cases): number of physical patches, faces of patch p, its face values and value
// fractions in the oracle's face order (block sides in blockMesh order).
int afo_case_n_patches(void* h) { Case* c = (Case*)h; return c->R == 1 ? (int)c->rk[0].mesh.patches.size() : -1; }
int64_t afo_case_patch_size(void* h, int p) { return (int64_t)((Case*)h)->rk[0].mesh.patches[p].faceCells.size(); }
int afo_case_patch_state(void* h, int p, scalar* Tb, scalar* valueFraction)
{
    const Patch& pa = ((Case*)h)->rk[0].mesh.patches[p];
    if (Tb) std::copy(pa.Tb.begin(), pa.Tb.end(), Tb);
    if (valueFraction) std::copy(pa.valueFraction.begin(), pa.valueFraction.end(), valueFraction);
    return 0;
}
// Test probe: the time-scheme state the reference's TEqn sees in the NEXT step -- deltaT0 after the coming runTime++ (= the deltaT
// saved at the last one) and the number of old-time levels T and alpha.solid have ([UPSTREAM] GeometricField::nOldTimes)
int afo_case_time_scheme_state(void* h, scalar* deltaTSave, int* TnOld, int* a1nOld)
{
    Case* c = (Case*)h;
    *deltaTSave = c->deltaTSave; *TnOld = c->TnOld; *a1nOld = c->a1nOld;
    return 0;
}
// CrankNicolson: the scheme's state as the last step's TEqn.H found it.  state[9] = timeIndex, TnOld, a1nOld, then for ddt0(T) and
// ddt0(alpha.solid): exists, startTimeIndex, timeIndex.  The fields (zeros where ddt0 did not exist yet) go to ddt0T / ddt0A.
int afo_case_cn_snapshot(void* h, int* state, scalar* ddt0T, scalar* ddt0A)
{
    Case* c = (Case*)h;
    const Case::CnSnapshot& q = c->cnSnap;
    const int st[9] = {(int)q.timeIndex, q.TnOld, q.a1nOld, q.T.exists, (int)q.T.startTimeIndex, (int)q.T.timeIndex,
                       q.A.exists, (int)q.A.startTimeIndex, (int)q.A.timeIndex};
    for (int i = 0; i < 9; i++) state[i] = st[i];
    size_t off = 0;
    for (int r = 0; r < c->R; r++) {
        const size_t n = (size_t)c->rk[r].mesh.nCells();
        for (size_t i = 0; i < n; i++) {
            ddt0T[off + i] = r < (int)q.ddt0T.size() && q.ddt0T[r].size() == n ? q.ddt0T[r][i] : 0.0;
            ddt0A[off + i] = r < (int)q.ddt0A.size() && q.ddt0A[r].size() == n ? q.ddt0A[r][i] : 0.0;
        }
        off += n;
    }
    return 0;
}
scalar afo_mixed_value_fraction(scalar h, scalar emissivity, scalar Tinf, scalar Tp, scalar kap, scalar deltaCoeff)
{
    return mixedValueFraction(h, emissivity, Tinf, Tp, kap, deltaCoeff);
}
// Test probe for the multi-beam heat source: what setDeltaT.H:47-49 and additiveFoam.C:74-76 do with the sources, nothing else --
// movingHeatSourceModel::adjustDeltaT (movingHeatSourceModel.C:92-98) on the proposed deltaT, ::update (:100-150) over
// [time, time + deltaT], then time += deltaT.  *dtOut = the adjusted deltaT; qDot is read with afo_case_get_field.
int afo_case_sources_probe(void* h, scalar proposedDeltaT, scalar* dtOut)
{
    Case* c = (Case*)h;
    scalar dt = proposedDeltaT;
    for (HeatSource& hs : c->sources) hs.beam.adjustDeltaT(dt, c->time);
    c->deltaT = dt;
    c->updateSources();
    c->time += dt;
    *dtOut = dt;
    return 0;
}

} // extern "C"
