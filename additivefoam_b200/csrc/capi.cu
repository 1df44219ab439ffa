// extern "C" surface declared in include/additivefoam_b200.h.  Exceptions stop here.
#include "../../include/additivefoam_b200.h"
#include "ctx.cuh"
#include "ldu.cuh"
#include "thermal.cuh"
#include <cstring>
#include <algorithm>
#include <string>
#include <sstream>

using namespace b200;

struct b200_ctx { Ctx* p; };
struct b200_ldu { Ldu* p; };
struct b200_case { ThermalCase* p; };

static thread_local std::string g_err;

#define B200_TRY try {
#define B200_CATCH                                                                               \
    }                                                                                            \
    catch (const std::exception& e) { g_err = e.what(); return 1; }                              \
    catch (...) { g_err = "unknown error"; return 1; }                                           \
    return 0;

extern "C" {

const char* b200_last_error(void) { return g_err.c_str(); }
int b200_abi_version(void) { return 3; }       // 2: b200_controls.ddtScheme; 3: b200_controls.ocCoeff (CrankNicolson)

int b200_nccl_unique_id(void* out128)
{
    B200_TRY
    ncclUniqueId id;
    B200_NCCL(ncclGetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, sizeof(id));
    B200_CATCH
}

int b200_device_count(int* count)
{
    B200_TRY
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) throw Error("no CUDA device available: the B200 path has no CPU fallback");
    *count = n;
    B200_CATCH
}
int b200_ctx_create(int device_ordinal, const void* nccl_unique_id, int rank, int nranks, b200_ctx** out)
{
    B200_TRY
    *out = new b200_ctx{new Ctx(device_ordinal, nccl_unique_id, rank, nranks)};
    B200_CATCH
}
int b200_ctx_destroy(b200_ctx* ctx)
{
    B200_TRY
    if (ctx) { cudaSetDevice(ctx->p->device); delete ctx->p; delete ctx; }
    B200_CATCH
}
int b200_ctx_peer_export(b200_ctx* ctx, void* handle64)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(ctx->p->device));
    ctx->p->peerExport(handle64);
    B200_CATCH
}
int b200_ctx_peer_export_halo(b200_ctx* ctx, int64_t haloDoubles, void* handle64)
{
    B200_TRY
    B200_CHECK(haloDoubles >= 0, "peer_export_halo: negative window size");
    B200_CUDA(cudaSetDevice(ctx->p->device));
    ctx->p->peerExport(handle64, (size_t)haloDoubles);
    B200_CATCH
}
int b200_ctx_peer_import(b200_ctx* ctx, const void* handles)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(ctx->p->device));
    ctx->p->peerImport(handles);
    B200_CATCH
}
int b200_ctx_synchronize(b200_ctx* ctx)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(ctx->p->device));
    ctx->p->sync();
    B200_CATCH
}

int b200_ctx_mark(b200_ctx* ctx, int i)
{
    B200_TRY
    B200_CHECK(i >= 0 && i < 4, "mark index");
    B200_CUDA(cudaSetDevice(ctx->p->device));
    B200_CUDA(cudaEventRecord(ctx->p->marks[i], ctx->p->stream));
    B200_CATCH
}
int b200_ctx_elapsed_ms(b200_ctx* ctx, int i, int j, double* ms)
{
    B200_TRY
    B200_CHECK(i >= 0 && i < 4 && j >= 0 && j < 4, "mark index");
    B200_CUDA(cudaSetDevice(ctx->p->device));
    B200_CUDA(cudaEventSynchronize(ctx->p->marks[j]));
    float t = 0;
    B200_CUDA(cudaEventElapsedTime(&t, ctx->p->marks[i], ctx->p->marks[j]));
    *ms = t;
    B200_CATCH
}
int b200_profile_enable(b200_ctx* ctx, int on)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(ctx->p->device));
    ctx->p->profRead(nullptr, nullptr);
    ctx->p->profOn = on != 0;
    B200_CATCH
}
int b200_profile_groups(void) { return PC_NCAT; }
const char* b200_profile_name(int group) { return prof_name(group); }
int b200_profile_read(b200_ctx* ctx, double* ms, int64_t* counts, int n)
{
    B200_TRY
    double m[PC_NCAT]; int64_t c[PC_NCAT];
    B200_CUDA(cudaSetDevice(ctx->p->device));
    ctx->p->profRead(m, c);
    for (int i = 0; i < n && i < PC_NCAT; i++) { ms[i] = m[i]; counts[i] = c[i]; }
    B200_CATCH
}

int b200_ldu_create(b200_ctx* ctx, int32_t nCells, int32_t nFaces, const int32_t* lower, const int32_t* upper, int32_t nInterfaces,
                    const int32_t* ifaceNeighbRank, const int32_t* ifaceSize, const int32_t* const* ifaceFaceCells, b200_ldu** out)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(ctx->p->device));
    *out = new b200_ldu{new Ldu(ctx->p, nCells, nFaces, lower, upper, nInterfaces, ifaceNeighbRank, ifaceSize, ifaceFaceCells)};
    B200_CATCH
}
int b200_ldu_destroy(b200_ldu* ldu)
{
    B200_TRY
    if (ldu) { cudaSetDevice(ldu->p->ctx->device); delete ldu->p; delete ldu; }
    B200_CATCH
}
int b200_ldu_set_block(b200_ldu* ldu, int32_t nx, int32_t ny, int32_t nz)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(ldu->p->ctx->device));
    ldu->p->setBlock(nx, ny, nz);
    B200_CATCH
}

// stage host arrays on the device (buffers cached in the handle)
static double* stage_in(Ldu& L, int slot, const double* host, size_t n)
{
    if (!host) return nullptr;
    if (!L.stage[slot]) {
        const size_t cap = std::max<size_t>((size_t)L.nFaces, (size_t)L.nCells) + 2 * (size_t)L.nCells + 16;
        L.stage[slot] = dev_alloc<double>(cap);
    }
    h2d(L.stage[slot], host, n, L.ctx->stream);
    return L.stage[slot];
}
static std::vector<const double*> stage_bou(Ldu& L, const double* const* hostBou)
{
    std::vector<const double*> out;
    if (L.ifaces.empty() || !hostBou) return out;
    if (L.stageBou.empty())
        for (auto& it : L.ifaces) L.stageBou.push_back(dev_alloc<double>(it.size));
    for (size_t p = 0; p < L.ifaces.size(); p++) {
        h2d(L.stageBou[p], hostBou[p], L.ifaces[p].size, L.ctx->stream);
        out.push_back(L.stageBou[p]);
    }
    return out;
}

int b200_ldu_solve_dev(b200_ldu* ldu, int32_t solver, int32_t preconditioner, const double* diag_dev, const double* upper_dev,
                       const double* lower_dev, const double* const* ifaceBouCoeffs_dev, const double* source_dev, double* psi_dev,
                       double tolerance, double relTol, int32_t minIter, int32_t maxIter, b200_solver_perf* perf)
{
    B200_TRY
    Ldu& L = *ldu->p;
    B200_CUDA(cudaSetDevice(L.ctx->device));
    MatrixView A; A.diag = diag_dev; A.upper = upper_dev; A.lower = lower_dev ? lower_dev : upper_dev; A.bouCoeffs = ifaceBouCoeffs_dev;
    SolveControls sc; sc.solver = solver; sc.precond = preconditioner; sc.tolerance = tolerance; sc.relTol = relTol;
    sc.minIter = minIter; sc.maxIter = maxIter;
    const SolvePerf sp = L.solve(A, source_dev, psi_dev, sc);
    L.ctx->sync();
    if (perf) {
        perf->initialResidual = sp.initialResidual; perf->finalResidual = sp.finalResidual;
        perf->nIterations = sp.nIterations; perf->converged = sp.converged; perf->singular = sp.singular;
    }
    B200_CATCH
}

int b200_ldu_solve(b200_ldu* ldu, int32_t solver, int32_t preconditioner, const double* diag, const double* upper, const double* lower,
                   const double* const* ifaceBouCoeffs, const double* source, double* psi, double tolerance, double relTol,
                   int32_t minIter, int32_t maxIter, b200_solver_perf* perf)
{
    B200_TRY
    Ldu& L = *ldu->p;
    B200_CUDA(cudaSetDevice(L.ctx->device));
    const double* d = stage_in(L, 0, diag, L.nCells);
    const double* u = stage_in(L, 1, upper, L.nFaces);
    const double* l = stage_in(L, 2, lower, L.nFaces);
    const double* b = stage_in(L, 3, source, L.nCells);
    double* x = stage_in(L, 4, psi, L.nCells);
    std::vector<const double*> bou = stage_bou(L, ifaceBouCoeffs);
    MatrixView A; A.diag = d; A.upper = u; A.lower = l ? l : u; A.bouCoeffs = bou.empty() ? nullptr : bou.data();
    SolveControls sc; sc.solver = solver; sc.precond = preconditioner; sc.tolerance = tolerance; sc.relTol = relTol;
    sc.minIter = minIter; sc.maxIter = maxIter;
    const SolvePerf sp = L.solve(A, b, x, sc);
    d2h(psi, x, L.nCells, L.ctx->stream);
    L.ctx->sync();
    if (perf) {
        perf->initialResidual = sp.initialResidual; perf->finalResidual = sp.finalResidual;
        perf->nIterations = sp.nIterations; perf->converged = sp.converged; perf->singular = sp.singular;
    }
    B200_CATCH
}

int b200_ldu_amul_dev(b200_ldu* ldu, const double* diag_dev, const double* upper_dev, const double* lower_dev,
                      const double* const* ifaceBouCoeffs_dev, const double* psi_dev, double* Apsi_dev)
{
    B200_TRY
    Ldu& L = *ldu->p;
    B200_CUDA(cudaSetDevice(L.ctx->device));
    MatrixView A; A.diag = diag_dev; A.upper = upper_dev; A.lower = lower_dev ? lower_dev : upper_dev; A.bouCoeffs = ifaceBouCoeffs_dev;
    L.amul(A, psi_dev, Apsi_dev);
    B200_CATCH
}

int b200_ldu_amul(b200_ldu* ldu, const double* diag, const double* upper, const double* lower, const double* const* ifaceBouCoeffs,
                  const double* psi, double* Apsi)
{
    B200_TRY
    Ldu& L = *ldu->p;
    B200_CUDA(cudaSetDevice(L.ctx->device));
    const double* d = stage_in(L, 0, diag, L.nCells);
    const double* u = stage_in(L, 1, upper, L.nFaces);
    const double* l = stage_in(L, 2, lower, L.nFaces);
    double* x = stage_in(L, 4, psi, L.nCells);
    if (!L.stage[5]) L.stage[5] = dev_alloc<double>((size_t)L.nCells + 16);
    double* y = L.stage[5];
    std::vector<const double*> bou = stage_bou(L, ifaceBouCoeffs);
    MatrixView A; A.diag = d; A.upper = u; A.lower = l ? l : u; A.bouCoeffs = bou.empty() ? nullptr : bou.data();
    L.amul(A, x, y);
    d2h(Apsi, y, L.nCells, L.ctx->stream);
    L.ctx->sync();
    B200_CATCH
}

int b200_ldu_precondition(b200_ldu* ldu, int32_t preconditioner, const double* diag, const double* upper, const double* lower,
                          const double* rA, double* wA)
{
    B200_TRY
    Ldu& L = *ldu->p;
    B200_CUDA(cudaSetDevice(L.ctx->device));
    const double* d = stage_in(L, 0, diag, L.nCells);
    const double* u = stage_in(L, 1, upper, L.nFaces);
    const double* l = stage_in(L, 2, lower, L.nFaces);
    const double* r = stage_in(L, 3, rA, L.nCells);
    // the wave sweeps never touch cells outside the block, but keep one plane of slack either side like the solver's vectors
    if (!L.stage[6]) L.stage[6] = dev_alloc<double>((size_t)3 * L.nCells + 16);
    double* w = L.stage[6] + L.nCells;
    MatrixView A; A.diag = d; A.upper = u; A.lower = l ? l : u;
    L.precondition(preconditioner, A, r, w);
    d2h(wA, w, L.nCells, L.ctx->stream);
    L.ctx->sync();
    B200_CATCH
}

int b200_ldu_levels(b200_ldu* ldu, int32_t* levelOfCell, int32_t* nLevels)
{
    B200_TRY
    ldu->p->levels(levelOfCell, nLevels);
    B200_CATCH
}

int b200_beam_qdot(b200_ctx* ctx, const b200_block_mesh* mesh, const b200_beam* beam, const double position[3],
                   const double dimensions[3], double power, double* qDot, double* sumWeights, double* volumeUsed,
                   double* absorbedPower)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(ctx->p->device));
    beam_qdot_standalone(ctx->p, *mesh, *beam, position, dimensions, power, qDot, sumWeights, volumeUsed, absorbedPower);
    B200_CATCH
}

int b200_beam_qdot_sub(b200_ctx* ctx, const b200_block_mesh* mesh, const int32_t offset[3], const int32_t nLocal[3], const b200_beam* beam,
                       const double position[3], const double dimensions[3], double power, b200_reduce_sum_fn reduceSum, void* user,
                       double* qDot, double* sumWeights, double* volumeUsed, double* absorbedPower)
{
    B200_TRY
    B200_CHECK(ctx && mesh && offset && nLocal && beam && position && dimensions && qDot, "beam_qdot_sub: null argument");
    B200_CUDA(cudaSetDevice(ctx->p->device));
    const int off[3] = {offset[0], offset[1], offset[2]}, nl[3] = {nLocal[0], nLocal[1], nLocal[2]};
    beam_qdot_standalone(ctx->p, *mesh, *beam, position, dimensions, power, qDot, sumWeights, volumeUsed, absorbedPower, off, nl, reduceSum, user);
    B200_CATCH
}

static heatSourceModel host_model(const b200_beam* beam, const double dimensions[3])
{
    B200_CHECK(beam != nullptr && dimensions != nullptr, "null beam or dimensions");
    B200_CHECK(beam->heatSourceModel >= B200_SUPER_GAUSSIAN && beam->heatSourceModel <= B200_PROJECTED_GAUSSIAN, "unknown heatSourceModel");
    b200_beam b = *beam;
    b.nSegments = 0; b.path = nullptr;                 // the shape does not depend on the path
    heatSourceModel m(b, 0.0, 0.0);
    for (int q = 0; q < 3; q++) m.dimensions[q] = dimensions[q];
    return m;
}
int b200_beam_shape_v0(const b200_beam* beam, const double dimensions[3], double* V0, double* kResolved)
{
    B200_TRY
    heatSourceModel m = host_model(beam, dimensions);
    const double v = m.V0();
    if (V0) *V0 = v;
    if (kResolved) *kResolved = m.k;
    B200_CATCH
}
int b200_beam_shape_weight(const b200_beam* beam, const double dimensions[3], const double d[3], double* weight)
{
    B200_TRY
    B200_CHECK(d != nullptr && weight != nullptr, "null argument");
    heatSourceModel m = host_model(beam, dimensions);
    m.V0();                                            // heatSourceModel.C:260: V0() precedes every weight()
    *weight = m.weight(d);
    B200_CATCH
}
int b200_beam_absorption_eta(const b200_beam* beam, double aspectRatio, double* eta)
{
    B200_TRY
    B200_CHECK(beam != nullptr && eta != nullptr, "null argument");
    b200_beam b = *beam;
    b.nSegments = 0; b.path = nullptr;
    *eta = heatSourceModel(b, 0.0, 0.0).eta(aspectRatio);
    B200_CATCH
}
int b200_mixed_temperature_value_fraction(double h, double emissivity, double Tinf, double Tp, double kappa, double deltaCoeff, double* valueFraction)
{
    B200_TRY
    B200_CHECK(valueFraction != nullptr, "null argument");
    *valueFraction = mixed_value_fraction(h, emissivity, Tinf, Tp, kappa, deltaCoeff);
    B200_CATCH
}

int b200_case_create(b200_ctx* ctx, const b200_case_desc* desc, b200_case** out)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(ctx->p->device));
    *out = new b200_case{new ThermalCase(ctx->p, *desc)};
    B200_CATCH
}
int b200_case_destroy(b200_case* c)
{
    B200_TRY
    if (c) { cudaSetDevice(c->p->ctx->device); delete c->p; delete c; }
    B200_CATCH
}
int b200_case_running(b200_case* c) { return c->p->running() ? 1 : 0; }
int b200_case_step(b200_case* c, b200_step_report* report)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->step(report);
    B200_CATCH
}
int b200_case_n_cells(b200_case* c, int64_t* nLocal, int64_t* nGlobal)
{
    B200_TRY
    if (nLocal) *nLocal = c->p->nLocal();
    if (nGlobal) *nGlobal = c->p->nGlobal();
    B200_CATCH
}
int b200_case_get_field(b200_case* c, int32_t field, double* out)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->getField(field, out);
    B200_CATCH
}
int b200_case_set_field(b200_case* c, int32_t field, const double* in)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->setField(field, in);
    B200_CATCH
}
int b200_case_reset_alpha_solid(b200_case* c)
{
    B200_TRY
    B200_CHECK(c != nullptr, "reset_alpha_solid: bad arguments");
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->resetAlphaSolidFromT();
    B200_CATCH
}
int b200_case_set_field_async(b200_case* c, int32_t field, const double* pinnedIn)
{
    B200_TRY
    B200_CHECK(c && pinnedIn, "set_field_async: bad arguments");
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->setFieldAsync(field, pinnedIn);
    B200_CATCH
}
int b200_case_apply_staged(b200_case* c)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->applyStaged();
    B200_CATCH
}
int b200_case_get_field_async(b200_case* c, int32_t field, double* pinnedOut)
{
    B200_TRY
    B200_CHECK(c && pinnedOut, "get_field_async: bad arguments");
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->getFieldAsync(field, pinnedOut);
    B200_CATCH
}
int b200_case_transfers_wait(b200_case* c)
{
    B200_TRY
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->transfersWait();
    B200_CATCH
}
int b200_case_melt_pool_dimensions(b200_case* c, int32_t nIso, const double* isoValues, double scanPathAngleDeg, double* dims)
{
    B200_TRY
    B200_CHECK(c && nIso >= 0 && (nIso == 0 || (isoValues && dims)), "melt_pool_dimensions: bad arguments");
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->meltPoolDimensions(nIso, isoValues, scanPathAngleDeg, dims);
    B200_CATCH
}
int b200_case_isotherm_depth(b200_case* c, int32_t beam, const double position[3], double dimensions[3])
{
    B200_TRY
    B200_CHECK(c && position && dimensions, "isotherm_depth: bad arguments");
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->isothermDepthProbe(beam, position, dimensions);
    B200_CATCH
}
int b200_case_sources_update(b200_case* c, double proposedDeltaT, double* deltaT)
{
    B200_TRY
    B200_CHECK(c && deltaT && proposedDeltaT > 0, "sources_update: bad arguments");
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->sourcesUpdate(proposedDeltaT, deltaT);
    B200_CATCH
}
int b200_case_solidification_enable(b200_case* c, const double boxMin[3], const double boxMax[3], double isoValue, int64_t maxEvents)
{
    B200_TRY
    B200_CHECK(c && boxMin && boxMax, "solidification_enable: bad arguments");
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    c->p->solidificationEnable(boxMin, boxMax, isoValue, maxEvents);
    B200_CATCH
}
int b200_case_solidification_events(b200_case* c, double* events, int64_t maxEvents, int64_t* nRecorded)
{
    B200_TRY
    B200_CHECK(c && nRecorded, "solidification_events: bad arguments");
    B200_CUDA(cudaSetDevice(c->p->ctx->device));
    *nRecorded = c->p->solidificationEvents(events, maxEvents);
    B200_CATCH
}
int b200_selftest_division(b200_ctx* ctx, int64_t nPairs, int64_t seed, int32_t expRange, int64_t* mismatches)
{
    B200_TRY
    B200_CHECK(ctx && mismatches && nPairs >= 0 && expRange >= 0 && expRange <= 1000, "selftest_division: bad arguments");
    B200_CUDA(cudaSetDevice(ctx->p->device));
    *mismatches = (int64_t)selftest_division(ctx->p, (unsigned long long)nPairs, (unsigned long long)seed, expRange);
    B200_CATCH
}
int b200_case_kernel_launches(b200_case*, int64_t* launches)
{
    *launches = g_launches;
    return 0;
}

int b200_create_scan_path(const double minPoint[2], const double maxPoint[2], double hatch, double rotationDeg, double power,
                          double speed, double dwellTime, int32_t biDirection, char* text, int64_t capacity, int64_t* length)
{
    B200_TRY
    B200_CHECK(minPoint != nullptr && maxPoint != nullptr && length != nullptr, "null argument");
    const std::string s = create_scan_path(minPoint, maxPoint, hatch, rotationDeg, power, speed, dwellTime, biDirection != 0);
    *length = (int64_t)s.size();
    if (text != nullptr && capacity > 0) {
        const size_t nCopy = std::min<size_t>(s.size(), (size_t)capacity);
        std::memcpy(text, s.data(), nCopy);
        if (nCopy < (size_t)capacity) text[nCopy] = 0;
    }
    B200_CATCH
}

int b200_scan_path_parse(const char* text, double* out, int32_t maxSegments)
{
    // MHS/movingBeam/movingBeam.C:89-148: first line is a header, blank lines are skipped,
    // each data line is "mode x y z power parameter" (segment.C:64-75)
    std::istringstream is(text ? text : "");
    std::string line;
    if (!std::getline(is, line)) return 0;
    int n = 0;
    while (std::getline(is, line)) {
        if (line.empty()) continue;
        std::stringstream ls(line);
        double v[6] = {0, 0, 0, 0, 0, 0};
        ls >> v[0] >> v[1] >> v[2] >> v[3] >> v[4] >> v[5];
        if (out != nullptr) {                      // out == NULL: count only
            if (n >= maxSegments) return -1;
            for (int q = 0; q < 6; q++) out[6 * n + q] = v[q];
        }
        n++;
    }
    return n;
}

} // extern "C"
