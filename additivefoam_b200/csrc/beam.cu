#include "thermal.cuh"
#include <algorithm>
#include <climits>

namespace b200 {

// =====================================================================================
// Host: scan path and kinematics
// =====================================================================================
movingBeam::movingBeam(const b200_beam& b, double startTime, double endTimeOfRun)
{
    deltaT_ = b.deltaT > 0 ? b.deltaT : 1e15;                 // GREAT, movingBeam.C:58,62
    hitPathIntervals_ = b.hitPathIntervals != 0;
    path_.assign(1, segment());                               // leading zero point source, movingBeam.C:53
    for (int i = 0; i < b.nSegments; i++) {
        const double* r = b.path + 6 * (size_t)i;
        segment s;
        s.mode = r[0]; s.x = r[1]; s.y = r[2]; s.z = r[3]; s.power = r[4]; s.parameter = r[5];
        path_.push_back(s);
    }
    // cumulative segment end times, movingBeam.C:127-145
    for (size_t i = 1; i < path_.size(); i++) {
        segment& s = path_[i];
        const segment& p = path_[i - 1];
        if (s.mode == 1) s.time = p.time + s.parameter;
        else {
            const double ex = s.x - p.x, ey = s.y - p.y, ez = s.z - p.z;
            s.time = p.time + std::sqrt(ex * ex + ey * ey + ez * ez) / s.parameter;
        }
    }
    index_ = findIndex(startTime);                            // movingBeam.C:69
    for (int i = (int)path_.size() - 1; i > 0; i--)            // movingBeam.C:74-81
        if (path_[i].power > 1e-15) { endTime_ = std::min(path_[i].time, endTimeOfRun); break; }
}

int movingBeam::findIndex(double time) const
{
    const int last = (int)path_.size() - 1;
    int i = index_;
    while (i > 0 && path_[i].time > time) --i;                // step back for safe updating
    while (i < last && path_[i].time < time) ++i;             // advance to the segment containing `time`
    while (i < last && path_[i].mode == 1 && path_[i].parameter == 0) ++i;   // skip zero-dwell point sources
    return std::min(std::max(i, 0), last);
}

void movingBeam::move(double time)
{
    index_ = findIndex(time);
    const segment& s = path_[index_];
    const segment& p = path_[index_ - 1];
    if (s.mode == 1) { position_[0] = s.x; position_[1] = s.y; position_[2] = s.z; }
    else {
        double d[3] = {0, 0, 0};
        const double dt = s.time - p.time;
        if (dt > 0) {
            const double tau = time - p.time;
            d[0] = (s.x - p.x) * tau / dt; d[1] = (s.y - p.y) * tau / dt; d[2] = (s.z - p.z) * tau / dt;
        }
        position_[0] = p.x + d[0]; position_[1] = p.y + d[1]; position_[2] = p.z + d[2];
    }
    power_ = (time - p.time) > eps ? s.power : p.power;
}

void movingBeam::adjustDeltaT(double& dt, double now) const
{
    if (!(activePath(now) && hitPathIntervals_)) return;
    double timeToNextPath = 0;
    int i = index_;
    while (timeToNextPath < eps) {
        timeToNextPath = std::max(0.0, path_[i].time - now);
        if (++i == (int)path_.size()) break;
    }
    const double nSteps = timeToNextPath / dt;
    if (nSteps < (double)INT32_MAX) {
        const int nStepsToNextPath = int(std::max(nSteps, 1.0) + 0.99);   // 1 % dilation allowed
        dt = std::min(timeToNextPath / nStepsToNextPath, dt);
    }
}

heatSourceModel::heatSourceModel(const b200_beam& b, double startTime, double endTimeOfRun) : beam(b, startTime, endTimeOfRun), spec(b)
{
    for (int q = 0; q < 3; q++) dimensions[q] = staticDimensions[q] = b.dimensions[q];
    transient = b.transient != 0;
    k = b.k;
    spec.path = nullptr;
    if (b.heatSourceModel == B200_PROJECTED_GAUSSIAN) V0();   // projectedGaussian.C:46-69 sets k_ in the ctor
}

double heatSourceModel::eta(double aspectRatio) const
{
    if (spec.absorptionModel == B200_ABSORPTION_CONSTANT) return spec.eta;      // constantAbsorption.H:86-92
    if (!(aspectRatio > 1.0)) return spec.etaMin;                                // KellyAbsorption.C:67-99
    const double theta = std::atan(1.0 / aspectRatio), e0 = spec.eta0;
    double F, G;
    if (spec.kellyGeometry == B200_KELLY_CONE) {
        F = 0.25 * (3.0 * std::sin(theta) - std::sin(3.0 * theta));
        G = 1.0 / (1.0 + std::sqrt(1.0 + std::pow(aspectRatio, 2)));
    } else {
        F = 0.5 * (1.0 - std::cos(2.0 * theta));
        G = 0.5 / (1.0 + aspectRatio);
    }
    return e0 * (1.0 + (1.0 - e0) * (G - F)) / (1.0 - (1.0 - e0) * (1.0 - G));
}

double heatSourceModel::V0()
{
    const double pi = 3.14159265358979323846;
    const double* D = dimensions;
    switch (spec.heatSourceModel) {
    case B200_SUPER_GAUSSIAN: {                                // superGaussian.C:73-85
        const double a = std::pow(2.0, 1.0 / k);
        return (2.0 / 3.0) * (D[0] / a) * (D[1] / a) * (D[2] / a) * pi * std::tgamma(1.0 + 3.0 / k);
    }
    case B200_MODIFIED_SUPER_GAUSSIAN: {                       // modifiedSuperGaussian.C:87-103
        const double a = std::pow(2.0, 1.0 / k), m = spec.m;
        return (D[0] / a) * (D[1] / a) * (D[2] / 1.0) * pi * std::tgamma(1.0 + 2.0 / k) * std::tgamma(1.0 + 1.0 / m)
             * std::tgamma(1.0 + 2.0 / m) / std::tgamma(1.0 + 3.0 / m);
    }
    default: {                                                 // projectedGaussian.C:96-118 (re-derives k_)
        const double x = std::max(D[2] / std::min(staticDimensions[0], staticDimensions[1]), 1.0);
        const double n = std::min(std::max(spec.A * std::log2(x) + spec.B, 0.0), 9.0);
        k = std::pow(2.0, n);
        return 0.5 * pi * D[0] * D[1] * D[2] * std::tgamma(1.0 / k) / (k * std::pow(3.0, 1.0 / k));
    }
    }
}

ShapeParams heatSourceModel::shapeParams() const
{
    ShapeParams sp;
    sp.model = spec.heatSourceModel;
    for (int q = 0; q < 3; q++) { sp.dims[q] = dimensions[q]; sp.pos[q] = beam.position()[q]; sp.nPoints[q] = spec.nPoints[q]; }
    sp.k = k; sp.m = spec.m;
    return sp;
}

// =====================================================================================
// Device: per-cell Riemann sum of the beam shape (heatSourceModel.C:287-356)
// =====================================================================================
__host__ __device__ __forceinline__ double shape_weight(const ShapeParams& sp, double dxv, double dyv, double dzv)
{
    if (sp.model == B200_SUPER_GAUSSIAN) {                     // superGaussian.C:62-70
        const double a = pow(2.0, 1.0 / sp.k);
        const double ux = dxv / (sp.dims[0] / a), uy = dyv / (sp.dims[1] / a), uz = dzv / (sp.dims[2] / a);
        return exp(-pow(ux * ux + uy * uy + uz * uz, sp.k / 2.0));
    }
    if (sp.model == B200_MODIFIED_SUPER_GAUSSIAN) {            // modifiedSuperGaussian.C:63-84
        const double a = pow(2.0, 1.0 / sp.k);
        double sx = sp.dims[0] / a, sy = sp.dims[1] / a;
        const double sz = sp.dims[2] / 1.0;
        if (dzv < sz) {
            const double f = pow(1.0 - pow(dzv / sz, sp.m), 1.0 / sp.m);
            sx *= f; sy *= f;
            const double ux = dxv / sx, uy = dyv / sy, uz = 0.0 / (sz * f);
            return exp(-pow(ux * ux + uy * uy + uz * uz, sp.k / 2.0));
        }
        return 0.0;
    }
    const double ux = dxv / sp.dims[0], uy = dyv / sp.dims[1];  // projectedGaussian.C:74-92
    const double f = exp(-2.0 * (ux * ux + uy * uy));
    const double s = exp(-3.0 * pow(fabs(fabs(dzv) / sp.dims[2]), sp.k));
    return f * s;
}

// heatSourceModel::weight() on the host (b200_beam_shape_weight): the same function the kernel calls
double heatSourceModel::weight(const double d[3]) const
{
    return shape_weight(shapeParams(), d[0], d[1], d[2]);
}

// One warp per cell of the index box; lanes stride over the sample points.
__global__ void __launch_bounds__(256) k_qdot_weights(BlockDims d, Geom g, ShapeParams sp, BoxRange box, double* __restrict__ wbuf,
                                                      RedWork red, int slot)
{
    const int bx = box.i1 - box.i0, by = box.j1 - box.j0, bz = box.k1 - box.k0;
    const int nBox = bx * by * bz;
    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    const double small_ = 1e-15;
    const double bmnx = sp.pos[0] - 1.5 * sp.dims[0], bmxx = sp.pos[0] + 1.5 * sp.dims[0];
    const double bmny = sp.pos[1] - 1.5 * sp.dims[1], bmxy = sp.pos[1] + 1.5 * sp.dims[1];
    const double bmnz = sp.pos[2] - 1.5 * sp.dims[2], bmxz = sp.pos[2] + 1.5 * sp.dims[2];
    double acc = 0;
    for (int b = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); b < nBox; b += gridDim.x * warpsPerBlock) {
        const int i = box.i0 + b % bx, j = box.j0 + (b / bx) % by, k = box.k0 + b / (bx * by);
        const double mnx = g.ox + (i + g.i0) * g.dx, mxx = g.ox + (i + g.i0 + 1) * g.dx;
        const double mny = g.oy + (j + g.j0) * g.dy, mxy = g.oy + (j + g.j0 + 1) * g.dy;
        const double mnz = g.oz + (k + g.k0) * g.dz, mxz = g.oz + (k + g.k0 + 1) * g.dz;
        double w = 0.0;
        // treeBoundBox::overlaps (inclusive)
        if (bmxx >= mnx && bmnx <= mxx && bmxy >= mny && bmny <= mxy && bmxz >= mnz && bmnz <= mxz) {
            const double spx = mxx - mnx, spy = mxy - mny, spz = mxz - mnz;
            double sdx = sp.dims[0] / (double)sp.nPoints[0], sdy = sp.dims[1] / (double)sp.nPoints[1], sdz = sp.dims[2] / (double)sp.nPoints[2];
            const int ncx = (int)fmax((spx + small_) / sdx, 1.0), ncy = (int)fmax((spy + small_) / sdy, 1.0),
                      ncz = (int)fmax((spz + small_) / sdz, 1.0);
            sdx = spx / (double)ncx; sdy = spy / (double)ncy; sdz = spz / (double)ncz;
            const double dVi = sdx * sdy * sdz;
            const int total = ncx * ncy * ncz;
            double wi = 0.0;
            for (int s = lane; s < total; s += 32) {
                const int si = s % ncx, sj = (s / ncx) % ncy, sk = s / (ncx * ncy);
                const double px = mxx - (si + 0.5) * sdx, py = mxy - (sj + 0.5) * sdy, pz = mxz - (sk + 0.5) * sdz;
                const double hx = 0.5 * sdx, hy = 0.5 * sdy, hz = 0.5 * sdz;
                if (px + hx >= bmnx && px - hx <= bmxx && py + hy >= bmny && py - hy <= bmxy && pz + hz >= bmnz && pz - hz <= bmxz)
                    wi += shape_weight(sp, fabs(px - sp.pos[0]), fabs(py - sp.pos[1]), fabs(pz - sp.pos[2])) * dVi;
            }
            wi = warp_reduce<Sum>(wi);
            wi = __shfl_sync(0xffffffffu, wi, 0);
            w = wi / g.V;
        }
        if (lane == 0) { wbuf[b] = w; acc += g.V * w; }
    }
    __shared__ double smem[32];
    double v[1] = {acc}; const int ops[1] = {0};
    grid_reduce<1>(v, ops, red, slot, smem);
}

// qDot = absorbedPower*weights/volume with the 5 % normalisation rule (heatSourceModel.C:359-367), then the
// time-weighted accumulation of movingHeatSourceModel::update (movingHeatSourceModel.C:128-147).
__global__ void k_qdot_apply(BlockDims d, BoxRange box, const double* __restrict__ wbuf, double* __restrict__ target, double absorbedPower,
                             double V0, const double* __restrict__ sumWeights, double dt, double sumDt, int directMean, double* volumeOut)
{
    const int bx = box.i1 - box.i0, by = box.j1 - box.j0, bz = box.k1 - box.k0;
    const int nBox = bx * by * bz;
    double volume = V0;
    const double sw = *sumWeights;
    if (fabs(1 - sw / volume) < 0.05) volume = sw;
    if (volumeOut && blockIdx.x == 0 && threadIdx.x == 0) *volumeOut = volume;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBox; b += gridDim.x * blockDim.x) {
        const int i = box.i0 + b % bx, j = box.j0 + (b / bx) % by, k = box.k0 + b / (bx * by);
        const int64_t c = i + (int64_t)d.nx * (j + (int64_t)d.ny * k);
        const double q = absorbedPower * wbuf[b] / volume;
        if (directMean == 2) target[c] = q;                      // standalone qDot()
        else if (directMean == 1) target[c] += (0.0 + dt * q) / sumDt;   // single sub-step: qDoti/sumWeights added to qDot_
        else target[c] += dt * q;                                // accumulate qDoti
    }
}

__global__ void k_qdoti_finish(BlockDims d, BoxRange box, double* __restrict__ qDoti, double* __restrict__ qDot, double sumDt)
{
    const int bx = box.i1 - box.i0, by = box.j1 - box.j0, bz = box.k1 - box.k0;
    const int nBox = bx * by * bz;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nBox; b += gridDim.x * blockDim.x) {
        const int i = box.i0 + b % bx, j = box.j0 + (b / bx) % by, k = box.k0 + b / (bx * by);
        const int64_t c = i + (int64_t)d.nx * (j + (int64_t)d.ny * k);
        qDot[c] += qDoti[c] / sumDt;
        qDoti[c] = 0.0;
    }
}

// heatSourceModel::updateDimensions (heatSourceModel.C:150-213): deepest isoValue crossing, located linearly
// between cell centres across each internal / processor face, within the xy search radius.
__global__ void __launch_bounds__(256) k_isotherm_depth(BlockDims d, Geom g, BoxRange cols, const double* __restrict__ T, double iso,
                                                        double px, double py, double pz, double searchRadius, double staticDepth,
                                                        int ghostBelow, int ghostAbove, RedWork red, int slot)
{
    const int bx = cols.i1 - cols.i0, by = cols.j1 - cols.j0;
    const int nCol = bx * by;
    const int64_t nxny = (int64_t)d.nx * d.ny;
    double maxDepth = staticDepth;
    const int64_t total = (int64_t)nCol * d.nz;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(q % nCol), k = (int)(q / nCol);
        const int i = cols.i0 + col % bx, j = cols.j0 + col / bx;
        const int64_t c = i + (int64_t)d.nx * j + nxny * k;
        const double To = T[c];
        const double cx = g.ox + (i + 0.5) * g.dx, cy = g.oy + (j + 0.5) * g.dy, cz = g.oz + (k + g.k0 + 0.5) * g.dz;
        auto probe = [&](double Town, double Tnei, double ox, double oy, double oz, double nx_, double ny_, double nz_) {
            const double mn = fmin(Town, Tnei), mx = fmax(Town, Tnei);
            if (mn < iso && mx >= iso) {
                const double fr = (iso - Town);
                const double den = (Tnei - Town);
                const double qx = fabs(ox + (nx_ - ox) * fr / den - px), qy = fabs(oy + (ny_ - oy) * fr / den - py),
                             qz = fabs(oz + (nz_ - oz) * fr / den - pz);
                if (sqrt(qx * qx + qy * qy) <= searchRadius) maxDepth = fmax(qz, maxDepth);
            }
        };
        if (i < d.nx - 1) probe(To, T[c + 1], cx, cy, cz, g.ox + (i + 1.5) * g.dx, cy, cz);
        if (j < d.ny - 1) probe(To, T[c + d.nx], cx, cy, cz, cx, g.oy + (j + 1.5) * g.dy, cz);
        if (k < d.nz - 1) probe(To, T[c + nxny], cx, cy, cz, cx, cy, g.oz + (k + g.k0 + 1.5) * g.dz);
        // processor faces: this rank is the owner side of its own patch face on both sides
        if (k == d.nz - 1 && ghostAbove) probe(To, T[c + nxny], cx, cy, cz, cx, cy, g.oz + (k + g.k0 + 1.5) * g.dz);
        if (k == 0 && ghostBelow) probe(To, T[c - nxny], cx, cy, cz, cx, cy, g.oz + (k + g.k0 - 0.5) * g.dz);
    }
    __shared__ double smem[32];
    double v[1] = {maxDepth}; const int ops[1] = {1};
    grid_reduce<1>(v, ops, red, slot, smem);
}

// ------------------------------------------------------------------ host launchers
BoxRange beam_box(const BlockDims& d, const Geom& g, const double pos[3], const double dims[3])
{
    // conservative index box around position +- 1.5*dimensions (one extra cell each side; the kernel
    // repeats the reference's inclusive overlap test per cell)
    auto lo = [](double x, double o, double h, int n) { int v = (int)std::floor((x - o) / h) - 1; return std::min(std::max(v, 0), n); };
    auto hi = [](double x, double o, double h, int n) { int v = (int)std::floor((x - o) / h) + 2; return std::min(std::max(v, 0), n); };
    BoxRange b;
    const double ox = g.ox + g.i0 * g.dx, oy = g.oy + g.j0 * g.dy;
    b.i0 = lo(pos[0] - 1.5 * dims[0], ox, g.dx, d.nx); b.i1 = hi(pos[0] + 1.5 * dims[0], ox, g.dx, d.nx);
    b.j0 = lo(pos[1] - 1.5 * dims[1], oy, g.dy, d.ny); b.j1 = hi(pos[1] + 1.5 * dims[1], oy, g.dy, d.ny);
    const double oz = g.oz + g.k0 * g.dz;
    b.k0 = lo(pos[2] - 1.5 * dims[2], oz, g.dz, d.nz); b.k1 = hi(pos[2] + 1.5 * dims[2], oz, g.dz, d.nz);
    return b;
}

static int64_t box_cells(const BoxRange& b)
{
    return (int64_t)std::max(b.i1 - b.i0, 0) * std::max(b.j1 - b.j0, 0) * std::max(b.k1 - b.k0, 0);
}

void launch_qdot_weights(Ctx* ctx, const BlockDims& d, const Geom& g, const ShapeParams& sp, const BoxRange& box, double* wbuf,
                         RedWork red, int slot)
{
    const int64_t n = box_cells(box);
    const int grid = (int)std::min<int64_t>(std::max<int64_t>((n + 7) / 8, 1), RED_MAX_BLOCKS);
    k_qdot_weights<<<grid, 256, 0, ctx->stream>>>(d, g, sp, box, wbuf, red, slot);
    B200_LAUNCH_CHECK();
}

void launch_qdot_apply(Ctx* ctx, const BlockDims& d, const BoxRange& box, const double* wbuf, double* target, double absorbedPower,
                       double V0, const double* sumWeights, double dt, double sumDt, int directMean, double* volumeOut)
{
    const int64_t n = box_cells(box);
    const int grid = (int)std::min<int64_t>(std::max<int64_t>((n + 255) / 256, 1), 4096);
    k_qdot_apply<<<grid, 256, 0, ctx->stream>>>(d, box, wbuf, target, absorbedPower, V0, sumWeights, dt, sumDt, directMean, volumeOut);
    B200_LAUNCH_CHECK();
}

void launch_qdoti_finish(Ctx* ctx, const BlockDims& d, const BoxRange& box, double* qDoti, double* qDot, double sumDt)
{
    const int64_t n = box_cells(box);
    const int grid = (int)std::min<int64_t>(std::max<int64_t>((n + 255) / 256, 1), 4096);
    k_qdoti_finish<<<grid, 256, 0, ctx->stream>>>(d, box, qDoti, qDot, sumDt);
    B200_LAUNCH_CHECK();
}

void launch_isotherm_depth(Ctx* ctx, const BlockDims& d, const Geom& g, const BoxRange& cols, const double* T, double isoValue,
                           const double pos[3], double searchRadius, double staticDepth, bool ghostBelow, bool ghostAbove,
                           RedWork red, int slot)
{
    const int64_t n = (int64_t)std::max(cols.i1 - cols.i0, 0) * std::max(cols.j1 - cols.j0, 0) * d.nz;
    const int grid = (int)std::min<int64_t>(std::max<int64_t>((n + 255) / 256, 1), RED_MAX_BLOCKS);
    k_isotherm_depth<<<grid, 256, 0, ctx->stream>>>(d, g, cols, T, isoValue, pos[0], pos[1], pos[2], searchRadius, staticDepth,
                                                    ghostBelow ? 1 : 0, ghostAbove ? 1 : 0, red, slot);
    B200_LAUNCH_CHECK();
}

} // namespace b200
