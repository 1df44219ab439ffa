#include "thermal_kernels.cuh"
#include <algorithm>
#include <cstring>
#include <climits>

namespace b200 {

static const double kSmall = 1e-15, kVGreat = 1e300;

// SOLVER/interpolateXY/interpolateXY.C:33-109 (host scalar; used for Tliq/Tsol and the initial alpha.solid)
static double interpolateXY(double x, const std::vector<double>& xOld, const std::vector<double>& yOld)
{
    const int n = (int)xOld.size();
    int lo = 0;
    for (lo = 0; lo < n && xOld[lo] > x; ++lo) {}
    int low = lo;
    if (low < n) for (int i = low; i < n; ++i) if (xOld[i] > xOld[lo] && xOld[i] <= x) lo = i;
    int hi = 0;
    for (hi = 0; hi < n && xOld[hi] < x; ++hi) {}
    int high = hi;
    if (high < n) for (int i = high; i < n; ++i) if (xOld[i] < xOld[hi] && xOld[i] >= x) hi = i;
    if (lo < n && hi < n && lo != hi) return yOld[lo] + ((x - xOld[lo]) / (xOld[hi] - xOld[lo])) * (yOld[hi] - yOld[lo]);
    if (lo == hi) return yOld[lo];
    if (lo == n) return yOld[hi];
    return yOld[lo];
}

double* ThermalCase::allocField()
{
    double* p = dev_alloc<double>(padded);
    B200_CUDA(cudaMemsetAsync(p, 0, padded * sizeof(double), ctx->stream));
    allocs.push_back(p);
    return p + plane;
}

static int cells_grid(Ctx* ctx, const BlockDims& d)
{
    const int64_t items = (int64_t)((d.nx + BLK - 1) / BLK) * d.ny * d.nz;
    return (int)std::max<int64_t>(1, std::min<int64_t>(items, ctx->persistentBlocks()));
}
static int flat_grid(Ctx* ctx, int64_t n)
{
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + BLK - 1) / BLK, ctx->persistentBlocks()));
}

ThermalCase::ThermalCase(Ctx* c, const b200_case_desc& desc) : ctx(c)
{
    ctl = desc.controls; mat = desc.material;
    B200_CHECK(mat.nThermo >= 2 && mat.nThermo <= 1024, "thermoPath needs 2..1024 rows");
    thermoT.assign(mat.thermoT, mat.thermoT + mat.nThermo);
    thermoA.assign(mat.thermoAlpha, mat.thermoAlpha + mat.nThermo);
    mat.thermoT = thermoT.data(); mat.thermoAlpha = thermoA.data();
    Tliq = interpolateXY(0.0, thermoA, thermoT);      // createFields.H:59-71
    Tsol = interpolateXY(1.0, thermoA, thermoT);
    Tinitial = desc.Tinitial;
    patches.assign(desc.patches, desc.patches + desc.nPatches);
    const b200_block_mesh& bm = desc.mesh;
    B200_CHECK(bm.nx > 0 && bm.ny > 0 && bm.nz > 0, "mesh dimensions must be positive");
    const int R = ctx->nranks, r = ctx->rank;
    B200_CHECK(R <= bm.nz, "more ranks than z layers");
    const int kBeg = (int)((int64_t)bm.nz * r / R), kEnd = (int)((int64_t)bm.nz * (r + 1) / R);
    dims.nx = bm.nx; dims.ny = bm.ny; dims.nz = kEnd - kBeg; nzGlobal = bm.nz;
    rankBelow = r > 0 ? r - 1 : -1; rankAbove = r < R - 1 ? r + 1 : -1;
    geo.ox = bm.min[0]; geo.oy = bm.min[1]; geo.oz = bm.min[2];
    geo.dx = (bm.max[0] - bm.min[0]) / bm.nx; geo.dy = (bm.max[1] - bm.min[1]) / bm.ny; geo.dz = (bm.max[2] - bm.min[2]) / bm.nz;
    geo.V = geo.dx * geo.dy * geo.dz;
    geo.sfx = geo.dy * geo.dz; geo.sfy = geo.dx * geo.dz; geo.sfz = geo.dx * geo.dy;
    geo.dcx = 1.0 / geo.dx; geo.dcy = 1.0 / geo.dy; geo.dcz = 1.0 / geo.dz;
    geo.bcx = 1.0 / (0.5 * geo.dx); geo.bcy = 1.0 / (0.5 * geo.dy); geo.bcz = 1.0 / (0.5 * geo.dz);
    geo.k0 = kBeg;
    plane = (size_t)dims.nx * dims.ny;
    padded = (size_t)dims.nCells() + 2 * plane;

    cudaStream_t s = ctx->stream;
    T = allocField(); Told = allocField(); alpha1 = allocField(); alpha1old = allocField(); alpha3 = allocField();
    kappa = allocField(); rKappa = allocField(); Cp = allocField(); qDot = allocField();
    dFdT = allocField(); T0 = allocField(); diag0 = allocField(); source0 = allocField(); diag = allocField(); source = allocField();
    upper = dev_alloc<double>((size_t)dims.nFaces() + 1);
    red.partials = dev_alloc<double>((size_t)R_NSLOTS * RED_MAX_BLOCKS);
    red.ticket = dev_alloc<unsigned>(1);
    red.result = dev_alloc<double>(R_NSLOTS);
    B200_CUDA(cudaMemsetAsync(red.ticket, 0, sizeof(unsigned), s));
    B200_CUDA(cudaMemsetAsync(red.result, 0, sizeof(double) * R_NSLOTS, s));
    B200_CUDA(cudaMallocHost(&h_red, sizeof(double) * R_NSLOTS));
    d_thermo = dev_alloc<double>(2 * (size_t)mat.nThermo);
    h2d(d_thermo, thermoT.data(), mat.nThermo, s);
    h2d(d_thermo + mat.nThermo, thermoA.data(), mat.nThermo, s);

    // the six sides: physical patches in patch order, then processor sides (below, above)
    int nOrder = 0;
    for (int q = 0; q < 6; q++) { sides[q] = SideInfo(); sidesDev.order[q] = 0; }
    bool assigned[6] = {false, false, false, false, false, false};
    patchUpdated.assign(patches.size(), false);
    auto sideCount = [&](int sd) { return sd <= B200_XMAX ? dims.ny * dims.nz : sd <= B200_YMAX ? dims.nx * dims.nz : dims.nx * dims.ny; };
    for (size_t p = 0; p < patches.size(); p++) {
        const b200_patch& ps = patches[p];
        for (int q = 0; q < ps.nSides; q++) {
            const int sd = ps.sides[q];
            B200_CHECK(sd >= 0 && sd < 6 && !assigned[sd], "block side listed twice or out of range");
            assigned[sd] = true;
            if ((sd == B200_ZMIN && rankBelow >= 0) || (sd == B200_ZMAX && rankAbove >= 0)) continue;   // not on this rank
            SideInfo& si = sides[sd];
            si.type = ps.type; si.patch = (int)p; si.value = ps.value; si.h = ps.h; si.emissivity = ps.emissivity; si.Tinf = ps.Tinf;
            sidesDev.order[nOrder++] = sd;
        }
    }
    for (int sd = 0; sd < 6; sd++) B200_CHECK(assigned[sd], "every block side must belong to a patch");
    if (rankBelow >= 0) { sides[B200_ZMIN].type = SIDE_PROCESSOR; sidesDev.order[nOrder++] = B200_ZMIN; }
    if (rankAbove >= 0) { sides[B200_ZMAX].type = SIDE_PROCESSOR; sidesDev.order[nOrder++] = B200_ZMAX; }
    sidesDev.nOrder = nOrder;
    for (int sd = 0; sd < 6; sd++) {
        SideInfo& si = sides[sd];
        si.n = sideCount(sd);
        si.valueFraction = dev_alloc<double>(si.n); si.Tb = dev_alloc<double>(si.n);
        si.intCoeff = dev_alloc<double>(si.n); si.bouCoeff = dev_alloc<double>(si.n);
        B200_CUDA(cudaMemsetAsync(si.valueFraction, 0, sizeof(double) * si.n, s));
        B200_CUDA(cudaMemsetAsync(si.intCoeff, 0, sizeof(double) * si.n, s));
        B200_CUDA(cudaMemsetAsync(si.bouCoeff, 0, sizeof(double) * si.n, s));
        k_fill<<<flat_grid(ctx, si.n), BLK, 0, s>>>(si.n, si.Tb, si.type == B200_BC_FIXED_VALUE ? si.value : Tinitial);
        B200_LAUNCH_CHECK();
        sidesDev.type[sd] = si.type; sidesDev.intCoeff[sd] = si.intCoeff; sidesDev.bouCoeff[sd] = si.bouCoeff;
        sidesDev.valueFraction[sd] = si.valueFraction; sidesDev.Tb[sd] = si.Tb; sidesDev.Tinf[sd] = si.Tinf;
    }
    // 0/T `value uniform` entries: the tutorials give the internal value; fixedValue faces hold their own value
    for (int sd = 0; sd < 6; sd++)
        if (sides[sd].type != B200_BC_FIXED_VALUE) { k_fill<<<flat_grid(ctx, sides[sd].n), BLK, 0, s>>>(sides[sd].n, sides[sd].Tb, Tinitial); B200_LAUNCH_CHECK(); }

    // fields: T uniform, alpha.solid from the thermoPath table (createFields.H:74-78), alpha.powder box
    const int64_t n = dims.nCells();
    k_fill<<<flat_grid(ctx, n), BLK, 0, s>>>(n, T, Tinitial); B200_LAUNCH_CHECK();
    k_fill<<<flat_grid(ctx, n), BLK, 0, s>>>(n, Told, Tinitial); B200_LAUNCH_CHECK();
    const double a0 = std::min(std::max(interpolateXY(Tinitial, thermoT, thermoA), 0.0), 1.0);
    k_fill<<<flat_grid(ctx, n), BLK, 0, s>>>(n, alpha1, a0); B200_LAUNCH_CHECK();
    k_fill<<<flat_grid(ctx, n), BLK, 0, s>>>(n, alpha1old, a0); B200_LAUNCH_CHECK();
    for (int k = 0; k < dims.nz; k++) {
        const double cz = geo.oz + (k + geo.k0 + 0.5) * geo.dz;
        if (cz > desc.powderZMin) { k_fill<<<flat_grid(ctx, (int64_t)plane), BLK, 0, s>>>((int64_t)plane, alpha3 + (size_t)k * plane, 1.0); B200_LAUNCH_CHECK(); }
    }
    time_ = ctl.startTime; deltaT_ = ctl.deltaT; deltaT0_ = deltaT_; deltaTSave_ = deltaT_;
    B200_CHECK(ctl.ddtScheme == B200_DDT_EULER || ctl.ddtScheme == B200_DDT_BACKWARD || ctl.ddtScheme == B200_DDT_CRANK_NICOLSON,
               "ddtScheme: Euler, backward or CrankNicolson (thermoScheme.H:4-15)");
    B200_CHECK(!crankNicolson() || (ctl.ocCoeff >= 0.0 && ctl.ocCoeff <= 1.0), "CrankNicolson: ocCoeff must lie in [0, 1]");
    if (backward() || crankNicolson()) {
        Toldold = allocField(); alpha1oldold = allocField();
        k_fill<<<flat_grid(ctx, n), BLK, 0, s>>>(n, Toldold, Tinitial); B200_LAUNCH_CHECK();
        k_fill<<<flat_grid(ctx, n), BLK, 0, s>>>(n, alpha1oldold, a0); B200_LAUNCH_CHECK();
    }
    for (int b = 0; b < desc.nBeams; b++) sources.emplace_back(desc.beams[b], time_, ctl.endTime);

    ldu.reset(new Ldu(ctx, dims));
    ldu->ghostBelow = rankBelow >= 0; ldu->ghostAbove = rankAbove >= 0;
    int nb = 0;
    if (rankBelow >= 0) bouPtrs[nb++] = sides[B200_ZMIN].bouCoeff;
    if (rankAbove >= 0) bouPtrs[nb++] = sides[B200_ZMAX].bouCoeff;

    double a, b2, c2;
    updatePropertiesAndNumbers(a, b2, c2, true);      // createFields.H:112
    ctx->sync();
}

ThermalCase::~ThermalCase()
{
    cudaStreamSynchronize(ctx->stream);
    for (double* p : allocs) cudaFree(p);
    dev_free(upper); dev_free(wbuf); dev_free(d_thermo);
    dev_free(red.partials); dev_free(red.ticket); dev_free(red.result);
    if (h_red) cudaFreeHost(h_red);
    dev_free(solidEv); dev_free(solidKeys); dev_free(solidCount);
    if (h2dStream) { cudaStreamSynchronize(h2dStream); cudaStreamSynchronize(d2hStream); cudaStreamDestroy(h2dStream); cudaStreamDestroy(d2hStream); }
    for (int f = 0; f < NIOF; f++) {
        dev_free(stageIn[f]); dev_free(stageOut[f]);
        for (cudaEvent_t e : {evIn[f], evInFree[f], evOutReady[f], evOutDone[f]}) if (e) cudaEventDestroy(e);
    }
    for (auto& si : sides) { dev_free(si.valueFraction); dev_free(si.Tb); dev_free(si.intCoeff); dev_free(si.bouCoeff); }
}

double* ThermalCase::field(int f)
{
    switch (f) {
    case B200_FIELD_T: return T;
    case B200_FIELD_ALPHA_SOLID: return alpha1;
    case B200_FIELD_ALPHA_POWDER: return alpha3;
    case B200_FIELD_KAPPA: return kappa;
    case B200_FIELD_CP: return Cp;
    case B200_FIELD_QDOT: return qDot;
    case B200_FIELD_DFDT: return dFdT;
    case B200_FIELD_T0: return T0;
    case B200_FIELD_T_OLD: return Told;
    case B200_FIELD_ALPHA_SOLID_OLD: return alpha1old;
    case B200_FIELD_T_OLD_OLD: B200_CHECK(Toldold != nullptr, "the second old-time level exists only with ddtScheme backward or CrankNicolson"); return Toldold;
    case B200_FIELD_ALPHA_SOLID_OLD_OLD: B200_CHECK(alpha1oldold != nullptr, "the second old-time level exists only with ddtScheme backward or CrankNicolson"); return alpha1oldold;
    case B200_FIELD_DDT0_T: case B200_FIELD_DDT0_ALPHA_SOLID: {
        B200_CHECK(crankNicolson(), "ddt0 exists only with ddtScheme CrankNicolson");
        double*& d0 = f == B200_FIELD_DDT0_T ? ddt0T : ddt0A;
        if (!d0) d0 = allocField();                       // zero, as the scheme would create it
        return d0;
    }
    }
    throw Error("unknown field id");
}

void ThermalCase::getField(int f, double* out)
{
    d2h(out, field(f), (size_t)dims.nCells(), ctx->stream);
    ctx->sync();
}

void ThermalCase::setField(int fIn, const double* in)
{
    const size_t n = (size_t)dims.nCells();
    const bool keepOld = (fIn & 0x100) != 0;     // B200_FIELD_KEEP_OLD: upload only (the host owns the field between steps)
    const int f = fIn & 0xff;
    h2d(field(f), in, n, ctx->stream);
    if (keepOld) { ctx->sync(); return; }
    if (f == B200_FIELD_T) {
        B200_CUDA(cudaMemcpyAsync(Told, T, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        evaluateT();
    }
    if (f == B200_FIELD_ALPHA_SOLID) B200_CUDA(cudaMemcpyAsync(alpha1old, alpha1, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->sync();
}

// Restart (createFields.H:31-44,74-78): alpha.solid is not a file of the case -- it is rebuilt from T
void ThermalCase::resetAlphaSolidFromT()
{
    const int64_t n = dims.nCells();
    k_alpha_from_T<<<flat_grid(ctx, n), BLK, 2 * (size_t)mat.nThermo * sizeof(double), ctx->stream>>>(n, T, d_thermo, mat.nThermo, alpha1, alpha1old);
    B200_LAUNCH_CHECK();
    double a, b2, c2;
    updatePropertiesAndNumbers(a, b2, c2, true);      // createFields.H:112 #include "updateProperties.H"
    ctx->sync();
}

// ---- streamed host I/O.  The host side of a step is then: inputs of the NEXT step cross PCIe on `h2dStream` while this step's
// kernels run, results of THIS step cross on `d2hStream` while the next step's kernels run; the compute stream only pays two
// device-to-device copies per field.  Host buffers must be pinned (cudaMemcpyAsync from pageable memory would serialise).
void ThermalCase::ensureIo(int f, bool out)
{
    B200_CHECK(f >= 0 && f < NIOF, "field id out of range");
    if (!h2dStream) {
        B200_CUDA(cudaStreamCreateWithFlags(&h2dStream, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&d2hStream, cudaStreamNonBlocking));
    }
    double*& buf = out ? stageOut[f] : stageIn[f];
    if (!buf) {
        buf = dev_alloc<double>((size_t)dims.nCells());
        if (out) { B200_CUDA(cudaEventCreateWithFlags(&evOutReady[f], cudaEventDisableTiming)); B200_CUDA(cudaEventCreateWithFlags(&evOutDone[f], cudaEventDisableTiming)); }
        else { B200_CUDA(cudaEventCreateWithFlags(&evIn[f], cudaEventDisableTiming)); B200_CUDA(cudaEventCreateWithFlags(&evInFree[f], cudaEventDisableTiming)); }
    }
}

// host -> staging buffer on the H2D stream; returns at once.  The field itself changes at the next applyStaged().
void ThermalCase::setFieldAsync(int f, const double* in)
{
    (void)field(f);
    ensureIo(f, false);
    B200_CUDA(cudaStreamWaitEvent(h2dStream, evInFree[f], 0));          // the previous staged value has been consumed (no-op the first time)
    if (stageOut[f]) B200_CUDA(cudaStreamWaitEvent(h2dStream, evOutDone[f], 0));   // ... and an outstanding download of this field has landed:
                                                                        // the round trip device -> host buffer -> device is ordered without the host
    B200_CUDA(cudaMemcpyAsync(stageIn[f], in, (size_t)dims.nCells() * sizeof(double), cudaMemcpyHostToDevice, h2dStream));
    B200_CUDA(cudaEventRecord(evIn[f], h2dStream));
    inPending[f] = true;
}

// staged inputs become the fields, on the compute stream (upload-only semantics of B200_FIELD_KEEP_OLD: the host owns the
// fields between steps, old-time copies follow at the step's own runTime++)
void ThermalCase::applyStaged()
{
    for (int f = 0; f < NIOF; f++) {
        if (!inPending[f]) continue;
        B200_CUDA(cudaStreamWaitEvent(ctx->stream, evIn[f], 0));
        B200_CUDA(cudaMemcpyAsync(field(f), stageIn[f], (size_t)dims.nCells() * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        B200_CUDA(cudaEventRecord(evInFree[f], ctx->stream));
        inPending[f] = false;
    }
}

// snapshot of the field on the compute stream, then staging buffer -> host on the D2H stream; returns at once
void ThermalCase::getFieldAsync(int f, double* out)
{
    double* src = field(f);
    ensureIo(f, true);
    B200_CUDA(cudaStreamWaitEvent(ctx->stream, evOutDone[f], 0));       // the previous snapshot has left the device
    B200_CUDA(cudaMemcpyAsync(stageOut[f], src, (size_t)dims.nCells() * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    B200_CUDA(cudaEventRecord(evOutReady[f], ctx->stream));
    B200_CUDA(cudaStreamWaitEvent(d2hStream, evOutReady[f], 0));
    B200_CUDA(cudaMemcpyAsync(out, stageOut[f], (size_t)dims.nCells() * sizeof(double), cudaMemcpyDeviceToHost, d2hStream));
    B200_CUDA(cudaEventRecord(evOutDone[f], d2hStream));
}

void ThermalCase::transfersWait()
{
    if (!h2dStream) return;
    B200_CUDA(cudaStreamSynchronize(h2dStream));
    B200_CUDA(cudaStreamSynchronize(d2hStream));
    ctx->sync();
}

// solidificationData.C:99-131 setOverlapCells: a cell belongs when its bounding box grown by 1e-10 overlaps the box (inclusive
// comparisons); on the block mesh that is an index box, found per axis with the same expressions.
void ThermalCase::solidificationEnable(const double boxMin[3], const double boxMax[3], double isoValue, int64_t maxEvents)
{
    B200_CHECK(!solidOn, "solidificationData is already enabled on this case");
    B200_CHECK(maxEvents > 0, "solidificationData: maxEvents must be positive");
    const double ext = 1e-10;
    auto range = [&](int n, double o, double h, int k0, double lo, double hi, int& a, int& b) {
        a = n; b = 0;
        for (int i = 0; i < n; i++) {
            const bool in = (o + (i + k0) * h) - ext <= hi && (o + (i + 1 + k0) * h) + ext >= lo;
            if (in) { a = std::min(a, i); b = std::max(b, i + 1); }
        }
        if (a > b) { a = 0; b = 0; }
    };
    range(dims.nx, geo.ox, geo.dx, 0, boxMin[0], boxMax[0], solidBox.i0, solidBox.i1);
    range(dims.ny, geo.oy, geo.dy, 0, boxMin[1], boxMax[1], solidBox.j0, solidBox.j1);
    range(dims.nz, geo.oz, geo.dz, geo.k0, boxMin[2], boxMax[2], solidBox.k0, solidBox.k1);
    solidIso = isoValue; solidMax = maxEvents; solidExec = 0;
    solidR = allocField();                               // zero: fvc::ddt(T) before the first step (:67-78)
    solidEv = dev_alloc<double>((size_t)maxEvents * 7);
    solidKeys = dev_alloc<long long>((size_t)maxEvents);
    solidCount = dev_alloc<unsigned long long>(1);
    B200_CUDA(cudaMemsetAsync(solidCount, 0, sizeof(unsigned long long), ctx->stream));
    solidOn = true;
}

void ThermalCase::solidificationExecute()
{
    const int64_t nBox = (int64_t)std::max(solidBox.i1 - solidBox.i0, 0) * std::max(solidBox.j1 - solidBox.j0, 0) *
                         std::max(solidBox.k1 - solidBox.k0, 0);
    solidExec++;
    haloExchange(T);                                     // collective: every rank calls it, with or without overlap cells
    if (nBox == 0) return;
    ProfScope ps_(ctx, PC_FUNCOBJ);
    SolidArgs a;
    a.box = solidBox; a.iso = solidIso; a.time = time_; a.rDeltaT = 1.0 / deltaT_; a.T = T; a.Told = Told; a.R = solidR;
    a.events = solidEv; a.keys = solidKeys; a.count = solidCount; a.maxEvents = solidMax; a.execIndex = solidExec;
    k_solidification<<<flat_grid(ctx, nBox), BLK, 0, ctx->stream>>>(dims, geo, sidesDev, a);
    B200_LAUNCH_CHECK();
}

// Events in the reference's order (time step, then cell).  Returns the number recorded; entries beyond the capacity given
// to solidificationEnable were dropped, which the caller sees as a return value above that capacity.
int64_t ThermalCase::solidificationEvents(double* out, int64_t maxOut)
{
    B200_CHECK(solidOn, "solidificationData is not enabled on this case");
    unsigned long long cnt = 0;
    d2h(&cnt, solidCount, 1, ctx->stream);
    ctx->sync();
    const int64_t have = std::min<int64_t>((int64_t)cnt, solidMax);
    if (out && have > 0 && maxOut > 0) {
        std::vector<double> ev((size_t)have * 7);
        std::vector<long long> keys((size_t)have);
        d2h(ev.data(), solidEv, (size_t)have * 7, ctx->stream);
        d2h(keys.data(), solidKeys, (size_t)have, ctx->stream);
        ctx->sync();
        std::vector<int64_t> order((size_t)have);
        for (int64_t i = 0; i < have; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return keys[x] < keys[y]; });
        for (int64_t i = 0; i < std::min(have, maxOut); i++) std::memcpy(out + 7 * i, ev.data() + 7 * order[i], 7 * sizeof(double));
    }
    return (int64_t)cnt;
}

// functionObjects/meltPoolDimensions: dims[3*i..] = max(span, 0) of isotherm i in the scan-path frame; the field never leaves HBM
void ThermalCase::meltPoolDimensions(int nIso, const double* isoValues, double scanPathAngleDeg, double* out)
{
    const double radians = scanPathAngleDeg * (3.14159265358979323846 / 180.0);
    const double sn = std::sin(radians), cs = std::cos(radians);
    haloExchange(T);
    for (int q = 0; q < nIso; q++) {
        k_melt_pool<<<cells_grid(ctx, dims), BLK, 0, ctx->stream>>>(dims, geo, sidesDev, T, isoValues[q], cs, sn, red);
        B200_LAUNCH_CHECK();
        ctx->allreduceMin(red.result + R_MPOOL, 3);
        ctx->allreduceMax(red.result + R_MPOOL + 3, 3);
        words_to_host(ctx, h_red + R_MPOOL, red.result + R_MPOOL, 6 * sizeof(double));
        ctx->sync();
        for (int a = 0; a < 3; a++) out[3 * q + a] = std::max(h_red[R_MPOOL + 3 + a] - h_red[R_MPOOL + a], 0.0);
    }
}

void ThermalCase::haloExchange(double* f)
{
    if (ctx->nranks <= 1) return;
    ProfScope ps_(ctx, PC_COMM);
    ctx->sendrecv(f, f - plane, rankBelow, f + (size_t)(dims.nz - 1) * plane, f + (size_t)dims.nz * plane, rankAbove, plane);
}

void ThermalCase::readReductions(int n)
{
    words_to_host(ctx, h_red, red.result, n * sizeof(double));
    ctx->sync();
}

BackwardCoeffs ThermalCase::backwardCoeffs(int nOld) const
{
    const double dt0 = nOld < 2 ? 1.0e15 : deltaT0_;            // backwardDdtScheme::deltaT0_(vf): great while vf.nOldTimes() < 2
    BackwardCoeffs k;
    k.rDeltaT = 1.0 / deltaT_;
    k.coefft = 1 + deltaT_ / (deltaT_ + dt0);
    k.coefft00 = deltaT_ * deltaT_ / (dt0 * (deltaT_ + dt0));
    k.coefft0 = k.coefft + k.coefft00;
    k.on = true;
    return k;
}

// [UPSTREAM] GeometricField::oldTime(): a level that does not exist yet is created as a copy of the level above it
void ThermalCase::touchOldTimes(bool ofT, int levels)
{
    int& nOld = ofT ? TnOld : a1nOld;
    if (nOld < 2 && levels >= 2) {
        const size_t n = (size_t)dims.nCells();
        B200_CUDA(cudaMemcpyAsync(ofT ? Toldold : alpha1oldold, ofT ? Told : alpha1old, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    nOld = std::max(nOld, levels);
}

// [UPSTREAM] CrankNicolsonDdtScheme::fvcDdt / fvmDdt up to the use of ddt0 (the rules: oracle/af_thermal.hpp, Case::cnPrepare)
CnCoeffs ThermalCase::cnPrepare(bool ofT, bool fvm)
{
    Ddt0State& st = ofT ? ddt0StateT : ddt0StateA;
    double*& d0 = ofT ? ddt0T : ddt0A;
    if (!st.exists) {
        if (!d0) d0 = allocField();
        else B200_CUDA(cudaMemsetAsync(d0, 0, (size_t)dims.nCells() * sizeof(double), ctx->stream));
        st.exists = true; st.startTimeIndex = timeIndex; st.timeIndex = timeIndex;
    }
    const double psi = ctl.ocCoeff;
    CnCoeffs k;
    k.rDtCoef = (timeIndex > st.startTimeIndex ? 1 + psi : 1) / deltaT_;
    k.psi = psi; k.ddt0 = d0; k.on = true;
    touchOldTimes(ofT, fvm ? 2 : 1);
    if (st.timeIndex != timeIndex) {
        touchOldTimes(ofT, 2);
        const double rDtCoef0 = (timeIndex > st.startTimeIndex + 1 ? 1 + psi : 1) / deltaT0_;
        const int64_t n = (int64_t)dims.nCells();
        k_ddt0_update<<<flat_grid(ctx, n), BLK, 0, ctx->stream>>>(n, d0, ofT ? Told : alpha1old, ofT ? Toldold : alpha1oldold, rDtCoef0, k);
        B200_LAUNCH_CHECK();
    }
    st.timeIndex = timeIndex;
    return k;
}

// updateProperties.H + the two numbers of setDeltaT.H (alphaCoNum without dt history issues, DiNum/dt)
void ThermalCase::updatePropertiesAndNumbers(double& alphaCoMax, double& diMax, double& minAlpha1, bool propertiesOnly)
{
    cudaStream_t s = ctx->stream;
    MatDev m;
    std::memcpy(m.kappa, mat.kappa, sizeof(m.kappa)); std::memcpy(m.Cp, mat.Cp, sizeof(m.Cp));
    m.rho = mat.rho; m.Lf = mat.Lf; m.Tliq = Tliq; m.Tsol = Tsol;
    // setDeltaT.H:6-11 evaluates fvc::ddt(T) with the scheme of fvSchemes (only when adjustTimeStep: the levels are created there)
    BackwardCoeffs kb{1.0 / deltaT_, 1, 1, 0, false};
    CnCoeffs kc{0, 1, nullptr, false};
    if (!propertiesOnly) {             // (createFields.H:112 runs updateProperties.H alone: no fvc::ddt(T), no old-time level is created)
        if (crankNicolson() && ctl.adjustTimeStep) kc = cnPrepare(true, false);
        else if (backward() && ctl.adjustTimeStep) { kb = backwardCoeffs(TnOld); touchOldTimes(true, 2); }
        else if (ctl.adjustTimeStep) touchOldTimes(true, 1);
    }
    { ProfScope ps_(ctx, PC_PROPS);
    k_props<<<cells_grid(ctx, dims), BLK, 0, s>>>(dims, m, T, Told, alpha1, alpha3, kappa, rKappa, Cp, 1.0 / deltaT_, deltaT_, red, kb, Toldold, kc);
    B200_LAUNCH_CHECK(); }
    haloExchange(kappa);
    haloExchange(rKappa);
    // DiNum and the face coefficients of the coming assembly in one pass over the kappa stencil; which form of `upper` the step will
    // want (thermoScheme.H:25-31 or :34-39) is only known after these reductions, so the form of the previous step is written and
    // assembleBase() redoes the faces alone when the guess was wrong
    { ProfScope ps_(ctx, PC_DINUM);
    if (lastExplicit) k_dinum_faces<true><<<cells_grid(ctx, dims), BLK, 0, s>>>(dims, geo, sidesDev, kappa, rKappa, Cp, mat.rho, upper, red);
    else k_dinum_faces<false><<<cells_grid(ctx, dims), BLK, 0, s>>>(dims, geo, sidesDev, kappa, rKappa, Cp, mat.rho, upper, red);
    upperForm = lastExplicit ? 2 : 1;
    B200_LAUNCH_CHECK(); }
    if (ctx->nranks > 1) {
        ctx->allreduceMax(red.result + R_ALPHACO, 1);
        ctx->allreduceMin(red.result + R_MINALPHA, 1);
        ctx->allreduceMax(red.result + R_DIMAX, 1);
    }
    readReductions(3);
    alphaCoMax = h_red[R_ALPHACO]; minAlpha1 = h_red[R_MINALPHA]; diMax = h_red[R_DIMAX];
}

// SOLVER/setDeltaT.H:1-50 and [UPSTREAM] Time::adjustDeltaT (adjustableRunTime write snapping)
void ThermalCase::setDeltaT(double alphaCoMax, double diMax)
{
    if (!ctl.adjustTimeStep) return;
    alphaCoNum = alphaCoMax;
    DiNum = diMax * deltaT_;
    const double CoNum = 0.0;          // CourantNo.H with phi == 0
    const double maxDeltaTFact = std::min(std::min(ctl.maxCo / (CoNum + kSmall), ctl.maxAlphaCo / (alphaCoNum + kSmall)),
                                          ctl.maxDi / (DiNum + kSmall));
    const double deltaTFact = std::min(std::min(maxDeltaTFact, 1.0 + 0.1 * maxDeltaTFact), 1.2);
    double dt = std::min(deltaTFact * deltaT_, ctl.maxDeltaT);
    for (auto& hs : sources) hs.beam.adjustDeltaT(dt, time_);      // movingHeatSourceModel.C:92-98
    deltaT_ = dt;
    if (ctl.writeInterval > 0) {
        const double timeToNextWrite = std::max(0.0, (writeTimeIndex + 1) * ctl.writeInterval - (time_ - ctl.startTime));
        const double nSteps = timeToNextWrite / deltaT_;
        if (nSteps < (double)INT32_MAX) {
            const int nStepsToNextWrite = int(std::max(nSteps, 1.0) + 0.99);
            const double newDeltaT = timeToNextWrite / nStepsToNextWrite;
            if (newDeltaT >= deltaT_) deltaT_ = std::min(newDeltaT, 2.0 * deltaT_);
            else deltaT_ = std::max(newDeltaT, 0.2 * deltaT_);
        }
    }
}

// heatSourceModel::updateDimensions (heatSourceModel.C:123-219)
void ThermalCase::updateDimensions(heatSourceModel& hs)
{
    if (!hs.transient) return;
    const double* pos = hs.beam.position();
    const double searchRadius = std::max(hs.staticDimensions[0], hs.staticDimensions[1]);
    haloExchange(T);
    // columns that can hold a crossing within the search radius (one extra cell each way)
    BoxRange cols;
    const double rx = searchRadius + 2.0 * geo.dx, ry = searchRadius + 2.0 * geo.dy;
    cols.i0 = std::min(std::max((int)std::floor((pos[0] - rx - geo.ox) / geo.dx), 0), dims.nx);
    cols.i1 = std::min(std::max((int)std::floor((pos[0] + rx - geo.ox) / geo.dx) + 1, 0), dims.nx);
    cols.j0 = std::min(std::max((int)std::floor((pos[1] - ry - geo.oy) / geo.dy), 0), dims.ny);
    cols.j1 = std::min(std::max((int)std::floor((pos[1] + ry - geo.oy) / geo.dy) + 1, 0), dims.ny);
    cols.k0 = 0; cols.k1 = dims.nz;
    launch_isotherm_depth(ctx, dims, geo, cols, T, hs.spec.isoValue, pos, searchRadius, hs.staticDimensions[2], rankBelow >= 0,
                          rankAbove >= 0, red, R_DEPTH);
    ctx->allreduceMax(red.result + R_DEPTH, 1);
    words_to_host(ctx, h_red + R_DEPTH, red.result + R_DEPTH, sizeof(double));
    ctx->sync();
    hs.dimensions[0] = hs.staticDimensions[0]; hs.dimensions[1] = hs.staticDimensions[1]; hs.dimensions[2] = h_red[R_DEPTH];
}

// updateDimensions of beam `beam` on the current T with the beam centre at `position`; the beam's own state is left as it was
void ThermalCase::isothermDepthProbe(int beam, const double position[3], double dimensionsOut[3])
{
    B200_CHECK(beam >= 0 && beam < (int)sources.size(), "beam index out of range");
    heatSourceModel hs = sources[beam];
    hs.beam.setPosition(position);
    updateDimensions(hs);
    for (int q = 0; q < 3; q++) dimensionsOut[q] = hs.dimensions[q];
}

// heatSourceModel::qDot() for the beam's current position/power, accumulated as movingHeatSourceModel::update does
void ThermalCase::qDotOnce(heatSourceModel& hs, double* target, double dt, double sumDt, bool directMean)
{
    const double power = hs.beam.power();
    if (!(power > kSmall)) return;
    const double aspectRatio = hs.dimensions[2] / std::min(hs.dimensions[0], hs.dimensions[1]);
    const double absorbedPower = hs.eta(aspectRatio) * power;
    const double V0 = hs.V0();
    const ShapeParams sp = hs.shapeParams();
    const BoxRange box = beam_box(dims, geo, sp.pos, sp.dims);
    const size_t nb = (size_t)std::max(box.i1 - box.i0, 0) * std::max(box.j1 - box.j0, 0) * std::max(box.k1 - box.k0, 0);
    if (nb + 1 > wbufCap) { dev_free(wbuf); wbufCap = std::max<size_t>(2 * nb + 1, 4096); wbuf = dev_alloc<double>(wbufCap); }
    launch_qdot_weights(ctx, dims, geo, sp, box, wbuf, red, R_SUMW);
    ctx->allreduceSum(red.result + R_SUMW, 1);
    launch_qdot_apply(ctx, dims, box, wbuf, target, absorbedPower, V0, red.result + R_SUMW, dt, sumDt, directMean ? 1 : 0,
                      red.result + R_VOLUME);
}

// movingHeatSourceModel::update (movingHeatSourceModel.C:100-150)
void launch_qdoti_finish(Ctx* ctx, const BlockDims& d, const BoxRange& box, double* qDoti, double* qDot, double sumDt);
void ThermalCase::updateSources()
{
    const size_t n = (size_t)dims.nCells();
    ProfScope ps_(ctx, PC_QDOT);
    B200_CUDA(cudaMemsetAsync(qDot, 0, n * sizeof(double), ctx->stream));
    for (auto& hs : sources) {
        if (!hs.beam.activePath(time_)) continue;
        updateDimensions(hs);
        double pathTime = time_;
        const double nextTime = pathTime + deltaT_;
        const double beam_dt = hs.beam.deltaT();
        const bool single = !(beam_dt < (nextTime - pathTime));       // the loop below runs exactly once
        if (!single && !qDoti) qDoti = allocField();
        double sumWeights = 0.0;
        BoxRange uni{INT_MAX, INT_MIN, INT_MAX, INT_MIN, INT_MAX, INT_MIN};
        while ((nextTime - pathTime) > kSmall) {
            const double dt = std::min(beam_dt, std::max(0.0, nextTime - pathTime));
            pathTime += dt;
            hs.beam.move(pathTime);
            if (single) qDotOnce(hs, qDot, dt, dt, true);
            else {
                qDotOnce(hs, qDoti, dt, 0.0, false);
                const ShapeParams sp = hs.shapeParams();
                const BoxRange b = beam_box(dims, geo, sp.pos, sp.dims);
                uni.i0 = std::min(uni.i0, b.i0); uni.i1 = std::max(uni.i1, b.i1); uni.j0 = std::min(uni.j0, b.j0);
                uni.j1 = std::max(uni.j1, b.j1); uni.k0 = std::min(uni.k0, b.k0); uni.k1 = std::max(uni.k1, b.k1);
            }
            sumWeights += dt;
        }
        if (!single && uni.i1 > uni.i0 && uni.j1 > uni.j0 && uni.k1 > uni.k0) launch_qdoti_finish(ctx, dims, uni, qDoti, qDot, sumWeights);
    }
}

// The sources' share of one pass of the time loop and nothing else: movingHeatSourceModel::adjustDeltaT (movingHeatSourceModel.C:92-98)
// on the proposed deltaT (setDeltaT.H:47-49), ::update (:100-150) over [time, time + deltaT] (additiveFoam.C:74), then the clock
// moves on (:76).  qDot is left in FIELD_QDOT; T and the phase fields are untouched.
void ThermalCase::sourcesUpdate(double proposedDeltaT, double* deltaTOut)
{
    double dt = proposedDeltaT;
    for (auto& hs : sources) hs.beam.adjustDeltaT(dt, time_);
    deltaT_ = dt;
    updateSources();
    time_ += dt;
    *deltaTOut = dt;
}

// mixedTemperature::updateCoeffs for every mixed side whose patch is not `updated`, plus (implicit form)
// the laplacian boundary coefficients of all sides
void ThermalCase::updateCoeffsT()
{
    for (int sd = 0; sd < 6; sd++) {
        SideInfo& si = sides[sd];
        if (si.type != B200_BC_MIXED_TEMPERATURE || si.patch < 0 || patchUpdated[si.patch]) continue;
        k_side_coeffs<<<flat_grid(ctx, si.n), BLK, 0, ctx->stream>>>(dims, geo, sd, si.type, si.value, si.h, si.emissivity, si.Tinf, kappa,
                                                                     rKappa, si.valueFraction, si.Tb, si.intCoeff, si.bouCoeff, 1, 0, si.n);
        B200_LAUNCH_CHECK();
    }
    for (size_t p = 0; p < patches.size(); p++) patchUpdated[p] = true;
}

// T.correctBoundaryConditions(): evaluate() calls updateCoeffs() when the patch is not updated, then resets the flag
void ThermalCase::evaluateT()
{
    ProfScope ps_(ctx, PC_BOUNDARY);
    updateCoeffsT();
    for (int sd = 0; sd < 6; sd++) {
        SideInfo& si = sides[sd];
        if (si.type != B200_BC_MIXED_TEMPERATURE) continue;     // fixedValue faces are constant, zeroGradient unused
        k_side_evaluate<<<flat_grid(ctx, si.n), BLK, 0, ctx->stream>>>(dims, sd, si.type, si.value, si.Tinf, T, si.valueFraction, si.Tb,
                                                                       sd <= B200_XMAX ? geo.bcx : sd <= B200_YMAX ? geo.bcy : geo.bcz, si.n);
        B200_LAUNCH_CHECK();
    }
    for (size_t p = 0; p < patches.size(); p++) patchUpdated[p] = false;
}

// thermoScheme.H:17-39
void ThermalCase::assembleBase(bool explicitSolve, double rDeltaT)
{
    cudaStream_t s = ctx->stream;
    updateCoeffsT();                   // fvScalarMatrix TEqn(T, dimPower) -> T.boundaryFieldRef().updateCoeffs()
    const int grid = cells_grid(ctx, dims);
    if (!explicitSolve) {
        for (int sd = 0; sd < 6; sd++) {
            SideInfo& si = sides[sd];
            k_side_coeffs<<<flat_grid(ctx, si.n), BLK, 0, s>>>(dims, geo, sd, si.type, si.value, si.h, si.emissivity, si.Tinf, kappa, rKappa,
                                                               si.valueFraction, si.Tb, si.intCoeff, si.bouCoeff, 0, 1, si.n);
            B200_LAUNCH_CHECK();
        }
        if (upperForm != 1) { ProfScope ps_(ctx, PC_FACES); k_faces<false><<<grid, BLK, 0, s>>>(dims, geo, rKappa, upper); B200_LAUNCH_CHECK(); }
        upperForm = 1; lastExplicit = false;
        // thermoScheme.H:34-39 fvm::ddt(T) with the scheme of fvSchemes
        BackwardCoeffs kT{rDeltaT, 1, 1, 0, false};
        CnCoeffs kc{0, 1, nullptr, false};
        if (backward()) kT = backwardCoeffs(TnOld);
        if (crankNicolson()) kc = cnPrepare(true, true);
        else touchOldTimes(true, backward() ? 2 : 1);
        { ProfScope ps_(ctx, PC_ASSEMBLE); k_assemble_implicit<<<grid, BLK, 0, s>>>(dims, geo, Cp, Told, qDot, upper, diag0, source0, rDeltaT, mat.rho, kT, Toldold, kc); B200_LAUNCH_CHECK(); }
    } else {
        touchOldTimes(true, 1);            // thermoScheme.H:25-31: the explicit form always uses EulerDdtScheme
        haloExchange(T);
        if (upperForm != 2) { ProfScope ps_(ctx, PC_FACES); k_faces<true><<<grid, BLK, 0, s>>>(dims, geo, rKappa, upper); B200_LAUNCH_CHECK(); }
        upperForm = 2; lastExplicit = true;
        { ProfScope ps_(ctx, PC_ASSEMBLE); k_assemble_explicit<<<grid, BLK, 0, s>>>(dims, geo, sidesDev, Cp, Told, T, qDot, upper, rKappa, diag0, source0, rDeltaT, mat.rho);
        B200_LAUNCH_CHECK(); }
    }
}

// One pass of the time loop body (additiveFoam.C:66-95)
void ThermalCase::step(b200_step_report* rep)
{
    cudaStream_t s = ctx->stream;
    const size_t n = (size_t)dims.nCells();
    double alphaCoMax, diMax, minAlpha1;
    updatePropertiesAndNumbers(alphaCoMax, diMax, minAlpha1);          // :68 and the reductions of :72
    setDeltaT(alphaCoMax, diMax);                                      // :72
    updateSources();                                                   // :74
    // runTime++ (:76)
    { ProfScope ps_(ctx, PC_COPY);
    // [UPSTREAM] GeometricField::storeOldTimes: the levels that exist shift (the second one by swapping buffers: its old content is dead)
    if (TnOld >= 2) std::swap(Told, Toldold);
    if (a1nOld >= 2) std::swap(alpha1old, alpha1oldold);
    B200_CUDA(cudaMemcpyAsync(Told, T, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    B200_CUDA(cudaMemcpyAsync(alpha1old, alpha1, n * sizeof(double), cudaMemcpyDeviceToDevice, s)); }
    deltaT0_ = deltaTSave_; deltaTSave_ = deltaT_;            // [UPSTREAM] Time::operator++
    time_ += deltaT_; timeIndex++;
    if (ctl.writeInterval > 0) {
        const int wi = int(((time_ - ctl.startTime) + 0.5 * deltaT_) / ctl.writeInterval);
        if (wi > writeTimeIndex) writeTimeIndex = wi;
    }
    // solutionControls.H
    { ProfScope ps_(ctx, PC_QDOT); k_power<<<flat_grid(ctx, (int64_t)n), BLK, 0, s>>>((int64_t)n, qDot, geo.V, red); B200_LAUNCH_CHECK(); }
    ctx->allreduceSum(red.result + R_POWER, 1);
    words_to_host(ctx, h_red + R_POWER, red.result + R_POWER, sizeof(double));
    ctx->sync();
    const double totalPower = h_red[R_POWER];
    double nThermoCorr = ctl.nThermoCorrectors;
    const bool beamOn = totalPower > kSmall;
    const bool fluidInDomain = (1 - minAlpha1) > kSmall;
    if (!(beamOn || fluidInDomain)) nThermoCorr = 1;
    bool explicitSolve = ctl.explicitSolve != 0;
    if (explicitSolve) explicitSolve = diMax * deltaT_ < 1.0;
    // thermo/TEqn.H
    const double rDeltaT = 1.0 / deltaT_;
    assembleBase(explicitSolve, rDeltaT);
    // thermoScheme.H:41-68: the rDeltaT of the Sp terms is scaled by coefft = 1 + deltaT/(deltaT + deltaT0) -- runTime.deltaT0() itself,
    // not the scheme's guarded one -- in the implicit form of the backward scheme
    const bool bwd = backward() && !explicitSolve;
    const bool cn = crankNicolson() && !explicitSolve;       // (:51-65: coefft = 1 + ocCoeff, at every step)
    const double rDeltaTSp = bwd ? rDeltaT * (1.0 + deltaT_ / (deltaT_ + deltaT0_)) : cn ? rDeltaT * (1.0 + ctl.ocCoeff) : rDeltaT;
    CorrectorArgs ca;
    ca.alpha1old = alpha1old; ca.diag0 = diag0; ca.source0 = source0; ca.T = T; ca.alpha1 = alpha1; ca.dFdT = dFdT; ca.T0 = T0;
    ca.diag = diag; ca.source = source; ca.thermo = d_thermo; ca.nThermo = mat.nThermo; ca.Tliq = Tliq; ca.Tsol = Tsol;
    ca.thermoTol = ctl.thermoTolerance; ca.Tmax = ctl.Tmax > 0 ? ctl.Tmax : kVGreat; ca.rDeltaT = rDeltaTSp; ca.rDeltaTEuler = rDeltaT;
    ca.kA = BackwardCoeffs{rDeltaT, 1, 1, 0, false}; ca.alpha1oldold = alpha1oldold; ca.cnA = CnCoeffs{0, 1, nullptr, false};
    ca.rho = mat.rho; ca.Lf = mat.Lf;
    const size_t shm = 2 * (size_t)mat.nThermo * sizeof(double);
    const int grid = cells_grid(ctx, dims);
    SolveControls sc;
    sc.solver = ctl.solver; sc.precond = ctl.preconditioner; sc.tolerance = ctl.tolerance; sc.relTol = ctl.relTol;
    sc.minIter = ctl.minIter; sc.maxIter = ctl.maxIter;
    MatrixView A; A.diag = diag; A.upper = upper; A.lower = upper; A.bouCoeffs = bouPtrs;
    A.offDiagEpoch = timeIndex;            // the off-diagonals are assembled once per time step (thermoScheme.H)
    SolvePerf perf;
    int nCorrDone = 0, iterSum = 0;
    double residual = 0;
    for (int tCorr = 0; tCorr < nThermoCorr; tCorr++) {
        // TEqn.H:36-39: fvc::ddt(alpha1) with the scheme (guarded by the levels alpha.solid has at THIS call), Euler in the explicit form
        if (bwd) ca.kA = backwardCoeffs(a1nOld);
        if (cn) ca.cnA = cnPrepare(false, false);
        else touchOldTimes(false, bwd ? 2 : 1);
        if (explicitSolve) {
            { ProfScope ps_(ctx, PC_CORRECTOR); k_corrector<true><<<grid, BLK, shm, s>>>(dims, geo, sidesDev, ca, red); B200_LAUNCH_CHECK(); }
            perf = SolvePerf(); perf.converged = 1;
        } else {
            { ProfScope ps_(ctx, PC_CORRECTOR); k_corrector<false><<<grid, BLK, shm, s>>>(dims, geo, sidesDev, ca, red); B200_LAUNCH_CHECK(); }
            perf = ldu->solve(A, source, T, sc);
            iterSum += perf.nIterations;
            { ProfScope ps_(ctx, PC_ALPHA); k_alpha_update<<<flat_grid(ctx, (int64_t)n), BLK, 0, s>>>((int64_t)n, T, dFdT, T0, alpha1, red); B200_LAUNCH_CHECK(); }
        }
        evaluateT();                                                   // psi.correctBoundaryConditions() at the end of solveSegregated
        evaluateT();                                                   // TEqn.H:44 T.correctBoundaryConditions(): the patches are not "updated" any more, so
                                                                       // mixedTemperature recomputes its value fraction from the new face values first
        ctx->allreduceMax(red.result + R_RESID, 1);
        words_to_host(ctx, h_red + R_RESID, red.result + R_RESID, sizeof(double));
        ctx->sync();
        residual = h_red[R_RESID];
        nCorrDone = tCorr + 1;
        if (residual < ctl.thermoTolerance && tCorr > 0) break;
    }
    if (rep) {
        rep->time = time_; rep->deltaT = deltaT_; rep->alphaCoNum = alphaCoNum; rep->DiNum = DiNum; rep->totalPower = totalPower;
        rep->explicitStep = explicitSolve ? 1 : 0; rep->nThermoCorr = nCorrDone; rep->thermoResidual = residual;
        rep->nSolveIterations = iterSum; rep->lastSolveIterations = perf.nIterations;
        rep->lastInitialResidual = perf.initialResidual; rep->lastFinalResidual = perf.finalResidual;
    }
    if (solidOn) solidificationExecute();                              // [UPSTREAM] Time::run -> functionObjectList::execute
}

// heatSourceModel::qDot() on its own (B2 of the drop-in boundary).  `offset` / `nLocal`: this process owns the sub-block of
// nLocal cells starting at global index `offset` of the blockMesh block `bm` (one rank of a decomposed run; points are those
// of the global block, min + (i0 + i)*d).  `reduceSum`: how the ranks add up sum(V*w) before the 5 % rule
// (fvc::domainIntegrate, heatSourceModel.C:359) -- the OpenFOAM-side shim passes its own Pstream reduce.
void beam_qdot_standalone(Ctx* ctx, const b200_block_mesh& bm, const b200_beam& beam, const double position[3], const double dimensions[3],
                          double power, double* qDotOut, double* sumWeights, double* volumeUsed, double* absorbedPowerOut,
                          const int* offset, const int* nLocal, double (*reduceSum)(double, void*), void* user)
{
    BlockDims d; d.nx = nLocal ? nLocal[0] : bm.nx; d.ny = nLocal ? nLocal[1] : bm.ny; d.nz = nLocal ? nLocal[2] : bm.nz;
    Geom g{};
    g.ox = bm.min[0]; g.oy = bm.min[1]; g.oz = bm.min[2];
    g.dx = (bm.max[0] - bm.min[0]) / bm.nx; g.dy = (bm.max[1] - bm.min[1]) / bm.ny; g.dz = (bm.max[2] - bm.min[2]) / bm.nz;
    g.V = g.dx * g.dy * g.dz; g.k0 = offset ? offset[2] : 0; g.i0 = offset ? offset[0] : 0; g.j0 = offset ? offset[1] : 0;
    B200_CHECK(d.nx > 0 && d.ny > 0 && d.nz > 0 && g.i0 >= 0 && g.j0 >= 0 && g.k0 >= 0 && g.i0 + d.nx <= bm.nx && g.j0 + d.ny <= bm.ny
               && g.k0 + d.nz <= bm.nz, "sub-block outside the block");
    const size_t n = (size_t)d.nCells();
    b200_beam spec = beam; spec.nSegments = 0; spec.path = nullptr;
    heatSourceModel hs(spec, 0.0, 0.0);
    for (int q = 0; q < 3; q++) hs.dimensions[q] = dimensions[q];
    if (sumWeights) *sumWeights = 0;
    if (volumeUsed) *volumeUsed = 0;
    if (absorbedPowerOut) *absorbedPowerOut = 0;
    double* q = dev_alloc<double>(n);
    B200_CUDA(cudaMemsetAsync(q, 0, n * sizeof(double), ctx->stream));
    RedWork red;
    red.partials = dev_alloc<double>((size_t)R_NSLOTS * RED_MAX_BLOCKS);
    red.ticket = dev_alloc<unsigned>(1); red.result = dev_alloc<double>(R_NSLOTS);
    B200_CUDA(cudaMemsetAsync(red.ticket, 0, sizeof(unsigned), ctx->stream));
    B200_CUDA(cudaMemsetAsync(red.result, 0, sizeof(double) * R_NSLOTS, ctx->stream));
    double* wb = nullptr;
    if (power > kSmall) {
        const double aspectRatio = hs.dimensions[2] / std::min(hs.dimensions[0], hs.dimensions[1]);
        const double absorbed = hs.eta(aspectRatio) * power;
        const double V0 = hs.V0();
        ShapeParams sp = hs.shapeParams();
        for (int c = 0; c < 3; c++) sp.pos[c] = position[c];
        const BoxRange box = beam_box(d, g, sp.pos, sp.dims);
        const size_t nb = (size_t)std::max(box.i1 - box.i0, 0) * std::max(box.j1 - box.j0, 0) * std::max(box.k1 - box.k0, 0);
        wb = dev_alloc<double>(nb + 1);
        launch_qdot_weights(ctx, d, g, sp, box, wb, red, R_SUMW);
        if (reduceSum) {          // every rank calls in, also those the beam does not touch (a collective, like gSum)
            double local = 0;
            d2h(&local, red.result + R_SUMW, 1, ctx->stream);
            ctx->sync();
            const double global = reduceSum(local, user);
            h2d(red.result + R_SUMW, &global, 1, ctx->stream);
            ctx->sync();
        }
        launch_qdot_apply(ctx, d, box, wb, q, absorbed, V0, red.result + R_SUMW, 0.0, 0.0, 2, red.result + R_VOLUME);
        double h[R_NSLOTS];
        d2h(h, red.result, R_NSLOTS, ctx->stream);
        ctx->sync();
        if (sumWeights) *sumWeights = h[R_SUMW];
        if (volumeUsed) *volumeUsed = h[R_VOLUME];
        if (absorbedPowerOut) *absorbedPowerOut = absorbed;
    }
    d2h(qDotOut, q, n, ctx->stream);
    ctx->sync();
    dev_free(q); dev_free(wb); dev_free(red.partials); dev_free(red.ticket); dev_free(red.result);
}

} // namespace b200
