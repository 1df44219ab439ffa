// Cell- and face-stream kernels of the resident thermal step (see thermal.cuh).
// Every kernel is a persistent grid over (row, x-segment) items of the block; neighbours are
// closed-form, matrix coefficients are written in OpenFOAM's LDU face order.
#pragma once
#include "thermal.cuh"
#include "block_addr.cuh"

namespace b200 {

enum { R_ALPHACO = 0, R_MINALPHA = 1, R_DIMAX = 2, R_POWER = 3, R_RESID = 4, R_SUMW = 5, R_DEPTH = 6, R_VOLUME = 7, R_MPOOL = 8 /* 6 slots */, R_NSLOTS = 16 };

struct MatDev {
    double kappa[3][3], Cp[3][3], rho, Lf, Tliq, Tsol;
};

#define B200_FOR_CELLS(d)                                                                        \
    const int segs_ = ((d).nx + BLK - 1) / BLK, nItems_ = segs_ * (d).ny * (d).nz;               \
    for (int item_ = blockIdx.x; item_ < nItems_; item_ += gridDim.x)

#define B200_CELL_INDEX(d)                                                                       \
    const int row = item_ / segs_;                                                               \
    const int i = (item_ - row * segs_) * BLK + threadIdx.x;                                     \
    if (i >= (d).nx) continue;                                                                   \
    const int k = row / (d).ny, j = row - k * (d).ny;                                            \
    const int64_t c = (int64_t)row * (d).nx + i;

__device__ __forceinline__ bool on_side(const BlockDims& d, int s, int i, int j, int k)
{
    switch (s) {
    case B200_XMIN: return i == 0;
    case B200_XMAX: return i == d.nx - 1;
    case B200_YMIN: return j == 0;
    case B200_YMAX: return j == d.ny - 1;
    case B200_ZMIN: return k == 0;
    default: return k == d.nz - 1;
    }
}
__device__ __forceinline__ int side_index(const BlockDims& d, int s, int i, int j, int k)
{
    if (s <= B200_XMAX) return j + d.ny * k;
    if (s <= B200_YMAX) return i + d.nx * k;
    return i + d.nx * j;
}
__device__ __forceinline__ double side_area(const Geom& g, int s) { return s <= B200_XMAX ? g.sfx : s <= B200_YMAX ? g.sfy : g.sfz; }
__device__ __forceinline__ double side_delta(const Geom& g, int s) { return s <= B200_XMAX ? g.bcx : s <= B200_YMAX ? g.bcy : g.bcz; }

// [UPSTREAM] Polynomial<3>::value
__device__ __forceinline__ double poly3(const double* v, double x)
{
    double val = v[0], powX = 1;
    powX *= x; val += v[1] * powX;
    powX *= x; val += v[2] * powX;
    return val;
}

// SOLVER/updateProperties.H:1-25 fused with the thermo number of setDeltaT.H:6-13 and gMin(alpha1)
// of solutionControls.H:23
__global__ void __launch_bounds__(BLK) k_props(BlockDims d, MatDev m, const double* __restrict__ T, const double* __restrict__ Told,
                                               const double* __restrict__ alpha1, double* __restrict__ alpha3, double* __restrict__ kappa,
                                               double* __restrict__ rKappa, double* __restrict__ Cp, double rDeltaT, double deltaT,
                                               RedWork red, BackwardCoeffs kb = BackwardCoeffs{0, 1, 1, 0, false},
                                               const double* __restrict__ Toldold = nullptr, CnCoeffs kc = CnCoeffs{0, 1, nullptr, false})
{
    double mxA = -1e300, mnA1 = 1e300;
    B200_FOR_CELLS(d)
    {
        B200_CELL_INDEX(d)
        (void)j; (void)k;
        const double a1 = alpha1[c], a2 = 1.0 - a1, Tc = T[c];
        double a3 = alpha3[c];
        if (a1 < 1.0) { a3 = 0.0; alpha3[c] = 0.0; }
        const double T1 = fmin(fmax(Tc, 300.0), m.Tsol);
        const double T2 = fmin(fmax(Tc, m.Tliq), 2.0 * m.Tliq);
        const double kap = (1.0 - a3) * (a1 * poly3(m.kappa[0], T1) + a2 * poly3(m.kappa[1], T2)) + a3 * poly3(m.kappa[2], T1);
        const double cp = (1.0 - a3) * (a1 * poly3(m.Cp[0], T1) + a2 * poly3(m.Cp[1], T2)) + a3 * poly3(m.Cp[2], T1);
        kappa[c] = kap; rKappa[c] = 1.0 / kap; Cp[c] = cp;
        // setDeltaT.H:6-11 fvc::ddt(T): Euler, or backward rDeltaT*(coefft*T - coefft0*T.oldTime() + coefft00*T.oldTime().oldTime())
        // ... or CrankNicolson rDtCoef*(T - T.oldTime()) - offCentre(ddt0(T))
        const double ddtT = kc.on ? kc.rDtCoef * (Tc - Told[c]) - cn_off_centre(kc, kc.ddt0[c])
                          : kb.on ? kb.rDeltaT * (kb.coefft * Tc - kb.coefft0 * Told[c] + kb.coefft00 * Toldold[c]) : rDeltaT * (Tc - Told[c]);
        mxA = fmax(mxA, fabs(ddtT) / (m.Tliq - m.Tsol) * deltaT);
        mnA1 = fmin(mnA1, a1);
    }
    __shared__ double smem[32];
    double v[2] = {mxA, mnA1}; const int ops[2] = {1, 2};
    grid_reduce<2>(v, ops, red, R_ALPHACO, smem);
}

// setDeltaT.H:19-29: max over cells of surfaceSum(magSf*interpolate(kappa)*deltaCoeffs)/(V*rho*Cp)
__global__ void __launch_bounds__(BLK) k_dinum(BlockDims d, Geom g, SidesDev sd, const double* __restrict__ kappa,
                                               const double* __restrict__ Cp, double rho, RedWork red)
{
    const int nx = d.nx; const int64_t nxny = (int64_t)d.nx * d.ny;
    double mx = -1e300;
    B200_FOR_CELLS(d)
    {
        B200_CELL_INDEX(d)
        const double kc = kappa[c];
        double S = 0.0;
        // internal faces in face order; linear interpolate = lambda*(P - N) + N with lambda = 0.5
        if (k > 0) { const double kP = kappa[c - nxny]; S += g.sfz * (0.5 * (kP - kc) + kc) * g.dcz; }
        if (j > 0) { const double kP = kappa[c - nx]; S += g.sfy * (0.5 * (kP - kc) + kc) * g.dcy; }
        if (i > 0) { const double kP = kappa[c - 1]; S += g.sfx * (0.5 * (kP - kc) + kc) * g.dcx; }
        if (i < nx - 1) { const double kN = kappa[c + 1]; S += g.sfx * (0.5 * (kc - kN) + kN) * g.dcx; }
        if (j < d.ny - 1) { const double kN = kappa[c + nx]; S += g.sfy * (0.5 * (kc - kN) + kN) * g.dcy; }
        if (k < d.nz - 1) { const double kN = kappa[c + nxny]; S += g.sfz * (0.5 * (kc - kN) + kN) * g.dcz; }
        const bool bnd = i == 0 || i == nx - 1 || j == 0 || j == d.ny - 1 || k == 0 || k == d.nz - 1;
        for (int q = 0; bnd && q < sd.nOrder; q++) {
            const int s = sd.order[q];
            if (!on_side(d, s, i, j, k)) continue;
            if (sd.type[s] == SIDE_PROCESSOR) {
                const double kn = kappa[s == B200_ZMIN ? c - nxny : c + nxny];
                S += g.sfz * (0.5 * kc + (1.0 - 0.5) * kn) * g.dcz;
            } else S += side_area(g, s) * kc * side_delta(g, s);
        }
        mx = fmax(mx, S / (g.V * rho * Cp[c]));
    }
    __shared__ double smem[32];
    double v[1] = {mx}; const int ops[1] = {1};
    grid_reduce<1>(v, ops, red, R_DIMAX, smem);
}

// harmonic::interpolate = 1/(reverseLinear(1/kappa)): kappa_f = 1/(0.5*(rP - rN) + rN)
// IMPLICIT: upper[f] = -(deltaCoeffs*(kappa_f*magSf)) (TEqn's off-diagonal);  EXPLICIT: upper[f] = kappa_f
template <bool EXPLICIT>
__global__ void __launch_bounds__(BLK) k_faces(BlockDims d, Geom g, const double* __restrict__ rKappa, double* __restrict__ upper)
{
    const int nx = d.nx; const int64_t nxny = (int64_t)d.nx * d.ny;
    B200_FOR_CELLS(d)
    {
        B200_CELL_INDEX(d)
        (void)j; (void)k;
        const RowCtx r = row_ctx(d, row);
        int fzl, fyl, fxl, fx, fy, fz;
        cell_faces(d, r, i, fzl, fyl, fxl, fx, fy, fz);
        const double* rp = rKappa + c;
        const double rP = ld_first(rp), rX = ld_first(fx >= 0 ? rp + 1 : rp), rY = ld_first(fy >= 0 ? rp + nx : rp),
                     rZ = ld_first(fz >= 0 ? rp + nxny : rp);
        if (fx >= 0) { const double kf = 1.0 / (0.5 * (rP - rX) + rX); upper[fx] = EXPLICIT ? kf : 0.0 - g.dcx * (kf * g.sfx); }
        if (fy >= 0) { const double kf = 1.0 / (0.5 * (rP - rY) + rY); upper[fy] = EXPLICIT ? kf : 0.0 - g.dcy * (kf * g.sfy); }
        if (fz >= 0) { const double kf = 1.0 / (0.5 * (rP - rZ) + rZ); upper[fz] = EXPLICIT ? kf : 0.0 - g.dcz * (kf * g.sfz); }
    }
}

// k_dinum and k_faces in one pass over the kappa stencil (both run between updateProperties.H and the assembly, and both are
// functions of kappa alone): the six neighbour values are loaded once, unconditionally (clamped to the cell itself at the block's
// edge), the DiNum sum and the three upper coefficients follow with the very expressions of the two kernels above.
template <bool EXPLICIT>
__global__ void __launch_bounds__(BLK) k_dinum_faces(BlockDims d, Geom g, SidesDev sd, const double* __restrict__ kappa,
                                                     const double* __restrict__ rKappa, const double* __restrict__ Cp, double rho,
                                                     double* __restrict__ upper, RedWork red)
{
    const int nx = d.nx; const int64_t nxny = (int64_t)d.nx * d.ny;
    double mx = -1e300;
    B200_FOR_CELLS(d)
    {
        B200_CELL_INDEX(d)
        const RowCtx r = row_ctx(d, row);
        int fzl, fyl, fxl, fx, fy, fz;
        cell_faces(d, r, i, fzl, fyl, fxl, fx, fy, fz);
        const double* kp = kappa + c; const double* rp = rKappa + c;
        const double kc = ld_first(kp), cp = ld_first(Cp + c);
        const double kzl = ld_first(fzl >= 0 ? kp - nxny : kp), kyl = ld_first(fyl >= 0 ? kp - nx : kp), kxl = ld_first(fxl >= 0 ? kp - 1 : kp);
        const double kx = ld_first(fx >= 0 ? kp + 1 : kp), ky = ld_first(fy >= 0 ? kp + nx : kp), kz = ld_first(fz >= 0 ? kp + nxny : kp);
        const double rP = ld_first(rp), rX = ld_first(fx >= 0 ? rp + 1 : rp), rY = ld_first(fy >= 0 ? rp + nx : rp),
                     rZ = ld_first(fz >= 0 ? rp + nxny : rp);
        double S = 0.0;
        if (fzl >= 0) S += g.sfz * (0.5 * (kzl - kc) + kc) * g.dcz;
        if (fyl >= 0) S += g.sfy * (0.5 * (kyl - kc) + kc) * g.dcy;
        if (fxl >= 0) S += g.sfx * (0.5 * (kxl - kc) + kc) * g.dcx;
        if (fx >= 0) S += g.sfx * (0.5 * (kc - kx) + kx) * g.dcx;
        if (fy >= 0) S += g.sfy * (0.5 * (kc - ky) + ky) * g.dcy;
        if (fz >= 0) S += g.sfz * (0.5 * (kc - kz) + kz) * g.dcz;
        const bool bnd = i == 0 || i == nx - 1 || j == 0 || j == d.ny - 1 || k == 0 || k == d.nz - 1;
        for (int q = 0; bnd && q < sd.nOrder; q++) {
            const int s = sd.order[q];
            if (!on_side(d, s, i, j, k)) continue;
            if (sd.type[s] == SIDE_PROCESSOR) {
                const double kn = kappa[s == B200_ZMIN ? c - nxny : c + nxny];
                S += g.sfz * (0.5 * kc + (1.0 - 0.5) * kn) * g.dcz;
            } else S += side_area(g, s) * kc * side_delta(g, s);
        }
        mx = fmax(mx, S / (g.V * rho * cp));
        if (fx >= 0) { const double kf = 1.0 / (0.5 * (rP - rX) + rX); upper[fx] = EXPLICIT ? kf : 0.0 - g.dcx * (kf * g.sfx); }
        if (fy >= 0) { const double kf = 1.0 / (0.5 * (rP - rY) + rY); upper[fy] = EXPLICIT ? kf : 0.0 - g.dcy * (kf * g.sfy); }
        if (fz >= 0) { const double kf = 1.0 / (0.5 * (rP - rZ) + rZ); upper[fz] = EXPLICIT ? kf : 0.0 - g.dcz * (kf * g.sfz); }
    }
    __shared__ double smem[32];
    double v[1] = {mx}; const int ops[1] = {1};
    grid_reduce<1>(v, ops, red, R_DIMAX, smem);
}

// One thread per face of one block side: mixedTemperature updateCoeffs
// (mixedTemperatureFvPatchScalarField.C:157-188) and, for the implicit form, the boundary coefficients of
// -fvm::laplacian(kappa, T) ([UPSTREAM] gaussLaplacianScheme + mixed/fixedValue/processor gradient coeffs).
__global__ void k_side_coeffs(BlockDims d, Geom g, int s, int type, double fixedValue, double h, double emissivity, double Tinf,
                              const double* __restrict__ kappa, const double* __restrict__ rKappa, double* __restrict__ valueFraction,
                              const double* __restrict__ Tb, double* __restrict__ intCoeff, double* __restrict__ bouCoeff,
                              int doUpdateCoeffs, int implicitForm, int n)
{
    const int64_t nxny = (int64_t)d.nx * d.ny;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        int i, j, k;
        if (s <= B200_XMAX) { j = q % d.ny; k = q / d.ny; i = s == B200_XMIN ? 0 : d.nx - 1; }
        else if (s <= B200_YMAX) { i = q % d.nx; k = q / d.nx; j = s == B200_YMIN ? 0 : d.ny - 1; }
        else { i = q % d.nx; j = q / d.nx; k = s == B200_ZMIN ? 0 : d.nz - 1; }
        const int64_t c = i + (int64_t)d.nx * j + nxny * k;
        const double area = side_area(g, s);
        const double delta = type == SIDE_PROCESSOR ? g.dcz : side_delta(g, s);
        double f = 0.0;
        if (type == B200_BC_MIXED_TEMPERATURE) {
            if (doUpdateCoeffs) {
                f = mixed_value_fraction(h, emissivity, Tinf, Tb[q], kappa[c], delta);
                valueFraction[q] = f;
            } else f = valueFraction[q];
        }
        if (!implicitForm) continue;
        double kb;
        if (type == SIDE_PROCESSOR) {
            const double rP = rKappa[c], rN = rKappa[s == B200_ZMIN ? c - nxny : c + nxny];
            kb = 1.0 / (0.5 * rP + (1.0 - 0.5) * rN);
        } else kb = 1.0 / rKappa[c];
        const double pGamma = kb * area;
        double gic = 0.0, gbc = 0.0;
        if (type == B200_BC_FIXED_VALUE) { gic = -delta; gbc = delta * Tb[q]; }
        else if (type == B200_BC_MIXED_TEMPERATURE) { gic = -f * delta; gbc = f * delta * Tinf + (1.0 - f) * 0.0; }
        else if (type == SIDE_PROCESSOR) { gic = -delta; gbc = -gic; }
        const double icLap = pGamma * gic, bcLap = -pGamma * gbc;
        intCoeff[q] = 0.0 - icLap;
        bouCoeff[q] = 0.0 - bcLap;
    }
}

// T.correctBoundaryConditions(): mixed: Tb = f*Tinf + (1-f)*(Tc + refGrad/deltaCoeffs); fixedValue: Tb = value
__global__ void k_side_evaluate(BlockDims d, int s, int type, double fixedValue, double Tinf, const double* __restrict__ T,
                                const double* __restrict__ valueFraction, double* __restrict__ Tb, double delta, int n)
{
    const int64_t nxny = (int64_t)d.nx * d.ny;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        int i, j, k;
        if (s <= B200_XMAX) { j = q % d.ny; k = q / d.ny; i = s == B200_XMIN ? 0 : d.nx - 1; }
        else if (s <= B200_YMAX) { i = q % d.nx; k = q / d.nx; j = s == B200_YMIN ? 0 : d.ny - 1; }
        else { i = q % d.nx; j = q / d.nx; k = s == B200_ZMIN ? 0 : d.nz - 1; }
        const double Tc = T[i + (int64_t)d.nx * j + nxny * k];
        if (type == B200_BC_MIXED_TEMPERATURE) { const double f = valueFraction[q]; Tb[q] = f * Tinf + (1.0 - f) * (Tc + 0.0 / delta); }
        else if (type == B200_BC_FIXED_VALUE) Tb[q] = fixedValue;
        else Tb[q] = Tc;
    }
}

// thermoScheme.H:34-39, implicit: diag0 = rho*Cp*V/dt + sum(off-diagonals), source0 = rho*Cp*V/dt*Told + V*qDot
__global__ void __launch_bounds__(BLK) k_assemble_implicit(BlockDims d, Geom g, const double* __restrict__ Cp, const double* __restrict__ Told,
                                                           const double* __restrict__ qDot, const double* __restrict__ upper,
                                                           double* __restrict__ diag0, double* __restrict__ source0, double rDeltaT, double rho,
                                                           BackwardCoeffs kb = BackwardCoeffs{0, 1, 1, 0, false}, const double* __restrict__ Toldold = nullptr,
                                                           CnCoeffs kc = CnCoeffs{0, 1, nullptr, false})
{
    B200_FOR_CELLS(d)
    {
        B200_CELL_INDEX(d)
        (void)j; (void)k;
        const RowCtx r = row_ctx(d, row);
        int fzl, fyl, fxl, fx, fy, fz;
        cell_faces(d, r, i, fzl, fyl, fxl, fx, fy, fz);
        const double rhoCp = rho * Cp[c];
        // negSumDiag of the laplacian (lapDiag -= a, a = -upper) then TEqn diag = ddt - lapDiag
        double ld = 0.0;
        if (fzl >= 0) ld += upper[fzl];
        if (fyl >= 0) ld += upper[fyl];
        if (fxl >= 0) ld += upper[fxl];
        if (fx >= 0) ld += upper[fx];
        if (fy >= 0) ld += upper[fy];
        if (fz >= 0) ld += upper[fz];
        if (kc.on) {    // [UPSTREAM] CrankNicolsonDdtScheme::fvmDdt: diag = rDtCoef*V, source = (rDtCoef*old + offCentre(ddt0))*V
            diag0[c] = (kc.rDtCoef * g.V) * rhoCp - ld;
            source0[c] = ((kc.rDtCoef * Told[c] + cn_off_centre(kc, kc.ddt0[c])) * g.V) * rhoCp + g.V * qDot[c];
        } else if (kb.on) {    // [UPSTREAM] backwardDdtScheme::fvmDdt: diag = (coefft*rDeltaT)*V, source = rDeltaT*V*(coefft0*old - coefft00*oldOld)
            diag0[c] = ((kb.coefft * kb.rDeltaT) * g.V) * rhoCp - ld;
            source0[c] = (kb.rDeltaT * g.V * (kb.coefft0 * Told[c] - kb.coefft00 * Toldold[c])) * rhoCp + g.V * qDot[c];
        } else {
            diag0[c] = (rDeltaT * g.V) * rhoCp - ld;
            source0[c] = (rDeltaT * Told[c] * g.V) * rhoCp + g.V * qDot[c];
        }
    }
}

// thermoScheme.H:25-30, explicit: source0 = rho*Cp*V/dt*Told + V*fvc::laplacian(kappa,T) + V*qDot ; diag0 = rho*Cp*V/dt
// kf[] holds the harmonic face conductivities in LDU face order (k_faces<true>).
__global__ void __launch_bounds__(BLK) k_assemble_explicit(BlockDims d, Geom g, SidesDev sd, const double* __restrict__ Cp,
                                                           const double* __restrict__ Told, const double* __restrict__ T,
                                                           const double* __restrict__ qDot, const double* __restrict__ kf,
                                                           const double* __restrict__ rKappa, double* __restrict__ diag0,
                                                           double* __restrict__ source0, double rDeltaT, double rho)
{
    const int nx = d.nx; const int64_t nxny = (int64_t)d.nx * d.ny;
    B200_FOR_CELLS(d)
    {
        B200_CELL_INDEX(d)
        const RowCtx r = row_ctx(d, row);
        int fzl, fyl, fxl, fx, fy, fz;
        cell_faces(d, r, i, fzl, fyl, fxl, fx, fy, fz);
        const double Tc = T[c];
        double lap = 0.0;
        // fvc::surfaceIntegrate: owner += flux, neighbour -= flux; flux = (kappa_f*snGrad)*magSf
        if (fzl >= 0) lap -= kf[fzl] * (g.dcz * (Tc - T[c - nxny])) * g.sfz;
        if (fyl >= 0) lap -= kf[fyl] * (g.dcy * (Tc - T[c - nx])) * g.sfy;
        if (fxl >= 0) lap -= kf[fxl] * (g.dcx * (Tc - T[c - 1])) * g.sfx;
        if (fx >= 0) lap += kf[fx] * (g.dcx * (T[c + 1] - Tc)) * g.sfx;
        if (fy >= 0) lap += kf[fy] * (g.dcy * (T[c + nx] - Tc)) * g.sfy;
        if (fz >= 0) lap += kf[fz] * (g.dcz * (T[c + nxny] - Tc)) * g.sfz;
        const bool bnd = i == 0 || i == nx - 1 || j == 0 || j == d.ny - 1 || k == 0 || k == d.nz - 1;
        for (int q = 0; bnd && q < sd.nOrder; q++) {
            const int s = sd.order[q];
            if (!on_side(d, s, i, j, k)) continue;
            const int type = sd.type[s];
            const int fi = side_index(d, s, i, j, k);
            double snGrad = 0.0, kb = 1.0 / rKappa[c];
            if (type == B200_BC_FIXED_VALUE) snGrad = side_delta(g, s) * (sd.Tb[s][fi] - Tc);
            else if (type == B200_BC_MIXED_TEMPERATURE) { const double f = sd.valueFraction[s][fi]; snGrad = f * (sd.Tinf[s] - Tc) * side_delta(g, s) + (1.0 - f) * 0.0; }
            else if (type == SIDE_PROCESSOR) {
                const int64_t cn = s == B200_ZMIN ? c - nxny : c + nxny;
                snGrad = g.dcz * (T[cn] - Tc);
                kb = 1.0 / (0.5 * rKappa[c] + (1.0 - 0.5) * rKappa[cn]);
            }
            lap += kb * snGrad * side_area(g, s);
        }
        const double rhoCp = rho * Cp[c];
        diag0[c] = (rDeltaT * g.V) * rhoCp;
        double src = (rDeltaT * Told[c] * g.V) * rhoCp;
        src += g.V * (lap / g.V);
        src += g.V * qDot[c];
        source0[c] = src;
    }
}

// thermoSource.H:7-66 + TEqn.H:24-42 + fvMatrix::solveSegregated boundary folding.
// EXPLICIT additionally performs diagonalSolver::solve and the alpha update of TEqn.H:44-53 in the same pass.
struct CorrectorArgs {
    const double *alpha1old, *diag0, *source0;
    double *T, *alpha1, *dFdT, *T0, *diag, *source;
    const double* thermo; int nThermo;
    double Tliq, Tsol, thermoTol, Tmax, rDeltaT, rDeltaTEuler, rho, Lf;
    BackwardCoeffs kA;              // fvc::ddt(alpha1) of TEqn.H:36-39 when the scheme is backward (kA.on), else the Euler form
    const double* alpha1oldold;
    CnCoeffs cnA;                   // ... or CrankNicolson (cnA.on)
};

template <bool EXPLICIT>
__global__ void __launch_bounds__(BLK) k_corrector(BlockDims d, Geom g, SidesDev sd, CorrectorArgs a, RedWork red)
{
    extern __shared__ double s_thermo[];      // x[0..n) then y[0..n)
    const int n = a.nThermo;
    for (int q = threadIdx.x; q < 2 * n; q += BLK) s_thermo[q] = a.thermo[q];
    __syncthreads();
    const double* xv = s_thermo; const double* yv = s_thermo + n;
    const double slope = 1e10;
    const double s2 = a.rDeltaT * a.rho * a.Lf;
    double mxRes = -1e300;
    auto cell = [&](int i, int j, int k, int64_t c, double x, double y, double aold, double d0, double s0, double aoo) {
        double dF, t0;
        if (x < a.Tliq && x > a.Tsol) {
            int lo = 0;
            for (lo = 0; lo < n && xv[lo] > x; ++lo) {}
            int low = lo;
            if (low < n) for (int q = low; q < n; ++q) if (xv[q] > xv[lo] && xv[q] <= x) lo = q;
            int hi = 0;
            for (hi = 0; hi < n && xv[hi] < x; ++hi) {}
            int high = hi;
            if (high < n) for (int q = high; q < n; ++q) if (xv[q] < xv[hi] && xv[q] >= x) hi = q;
            dF = (yv[hi] - yv[lo]) / (xv[hi] - xv[lo]);
            t0 = xv[lo] + (y - yv[lo]) / dF;
        } else if (x >= a.Tliq && y > a.thermoTol) { dF = slope; t0 = a.Tliq - a.thermoTol; }
        else if (x <= a.Tsol && y < 1 - a.thermoTol) { dF = slope; t0 = a.Tsol + a.thermoTol; }
        else { dF = 0.0; t0 = x; }
        const double Acoef = 1e15 * ((x - a.Tmax) >= 0 ? 1.0 : 0.0);
        const double diag1 = (g.V * Acoef) * a.rDeltaT;
        const double source1 = (g.V * (Acoef * a.Tmax)) * a.rDeltaT;
        const double diag2 = (g.V * dF) * s2;
        const double source2 = (g.V * (dF * t0)) * s2;
        const double su = a.cnA.on ? (a.rho * a.Lf) * (a.cnA.rDtCoef * (y - aold) - cn_off_centre(a.cnA, a.cnA.ddt0[c]))
                        : a.kA.on ? (a.rho * a.Lf) * (a.kA.rDeltaT * (a.kA.coefft * y - a.kA.coefft0 * aold + a.kA.coefft00 * aoo))
                                  : (a.rho * a.Lf) * (a.rDeltaTEuler * (y - aold));
        const double sourceR = source2 - g.V * su;
        double dg = (d0 + diag1) - diag2;
        double sr = (s0 + source1) - sourceR;
        if (!EXPLICIT) {      // the explicit matrix has zero boundary coefficients
            const bool bnd = i == 0 || i == d.nx - 1 || j == 0 || j == d.ny - 1 || k == 0 || k == d.nz - 1;
            for (int q = 0; bnd && q < sd.nOrder; q++) {
                const int s = sd.order[q];
                if (on_side(d, s, i, j, k)) dg += sd.intCoeff[s][side_index(d, s, i, j, k)];
            }
            for (int q = 0; bnd && q < sd.nOrder; q++) {
                const int s = sd.order[q];
                if (sd.type[s] != SIDE_PROCESSOR && on_side(d, s, i, j, k)) sr += sd.bouCoeff[s][side_index(d, s, i, j, k)];
            }
            a.dFdT[c] = dF; a.T0[c] = t0; a.diag[c] = dg; a.source[c] = sr;
        } else {
            const double Tn = sr / dg;
            a.T[c] = Tn;
            const double an = fmin(fmax(y + dF * (Tn - t0), 0.0), 1.0);
            a.alpha1[c] = an;
            mxRes = fmax(mxRes, fabs(an - y));
        }
    };
    // two items per trip: the loads of both are in flight before either is worked on (one cell per thread per trip left the
    // kernel at half the memory parallelism the HBM needs)
    const int segs = (d.nx + BLK - 1) / BLK, nItems = segs * d.ny * d.nz;
    for (int itemA = blockIdx.x; itemA < nItems; itemA += 2 * gridDim.x) {
        const int itemB = itemA + gridDim.x;
        const int rowA = itemA / segs, iA = (itemA - rowA * segs) * BLK + threadIdx.x;
        const int rowB = itemB / segs, iB = (itemB - rowB * segs) * BLK + threadIdx.x;
        const bool okA = iA < d.nx, okB = itemB < nItems && iB < d.nx;
        const int64_t cA = (int64_t)rowA * d.nx + iA, cB = (int64_t)rowB * d.nx + iB;
        double xA = 0, yA = 0, oA = 0, dA = 0, sA = 0, xB = 0, yB = 0, oB = 0, dB = 0, sB = 0, ooA = 0, ooB = 0;
        if (okA) { xA = ld_first_rw(a.T + cA); yA = ld_first_rw(a.alpha1 + cA); oA = ld_first(a.alpha1old + cA); dA = ld_first(a.diag0 + cA); sA = ld_first(a.source0 + cA); }
        if (okB) { xB = ld_first_rw(a.T + cB); yB = ld_first_rw(a.alpha1 + cB); oB = ld_first(a.alpha1old + cB); dB = ld_first(a.diag0 + cB); sB = ld_first(a.source0 + cB); }
        if (a.kA.on) { if (okA) ooA = ld_first(a.alpha1oldold + cA); if (okB) ooB = ld_first(a.alpha1oldold + cB); }
        if (okA) { const int k = rowA / d.ny; cell(iA, rowA - k * d.ny, k, cA, xA, yA, oA, dA, sA, ooA); }
        if (okB) { const int k = rowB / d.ny; cell(iB, rowB - k * d.ny, k, cB, xB, yB, oB, dB, sB, ooB); }
    }
    if (EXPLICIT) {
        __shared__ double smem[32];
        double v[1] = {mxRes}; const int ops[1] = {1};
        grid_reduce<1>(v, ops, red, R_RESID, smem);
    }
}

// [UPSTREAM] CrankNicolsonDdtScheme, evaluate(ddt0): ddt0 = rDtCoef0*(vf.oldTime() - vf.oldTime().oldTime()) - offCentre(ddt0)
__global__ void __launch_bounds__(BLK) k_ddt0_update(int64_t n, double* __restrict__ ddt0, const double* __restrict__ old,
                                                     const double* __restrict__ oldold, double rDtCoef0, CnCoeffs k)
{
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK)
        ddt0[c] = rDtCoef0 * (old[c] - oldold[c]) - cn_off_centre(k, ddt0[c]);
}

// TEqn.H:46-53: alpha1 = min(max(alpha10 + dFdT*(T - T0), 0), 1); residual = max|alpha1 - alpha10|
__global__ void __launch_bounds__(BLK) k_alpha_update(int64_t n, const double* __restrict__ T, const double* __restrict__ dFdT,
                                                      const double* __restrict__ T0, double* __restrict__ alpha1, RedWork red)
{
    double mx = -1e300;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) {
        const double a10 = alpha1[c];
        const double an = fmin(fmax(a10 + dFdT[c] * (T[c] - T0[c]), 0.0), 1.0);
        alpha1[c] = an;
        mx = fmax(mx, fabs(an - a10));
    }
    __shared__ double smem[32];
    double v[1] = {mx}; const int ops[1] = {1};
    grid_reduce<1>(v, ops, red, R_RESID, smem);
}

// createFields.H:74-78 (also what a restart does, :31-44: alpha.solid is never read): alpha.solid = min(max(interpolateXY(T, thermo.x,
// thermo.y), 0), 1) with SOLVER/interpolateXY/interpolateXY.C:33-109 -- the linear scans that tolerate unsorted abscissae
__global__ void __launch_bounds__(BLK) k_alpha_from_T(int64_t n, const double* __restrict__ T, const double* __restrict__ thermo, int nT,
                                                      double* __restrict__ alpha1, double* __restrict__ alpha1old)
{
    extern __shared__ double s_tab[];           // x[0..nT) then y[0..nT)
    for (int q = threadIdx.x; q < 2 * nT; q += BLK) s_tab[q] = thermo[q];
    __syncthreads();
    const double* xOld = s_tab; const double* yOld = s_tab + nT;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) {
        const double x = T[c];
        int lo = 0;
        for (lo = 0; lo < nT && xOld[lo] > x; ++lo) {}
        const int low = lo;
        if (low < nT) for (int i = low; i < nT; ++i) if (xOld[i] > xOld[lo] && xOld[i] <= x) lo = i;
        int hi = 0;
        for (hi = 0; hi < nT && xOld[hi] < x; ++hi) {}
        const int high = hi;
        if (high < nT) for (int i = high; i < nT; ++i) if (xOld[i] < xOld[hi] && xOld[i] >= x) hi = i;
        double y;
        if (lo < nT && hi < nT && lo != hi) y = yOld[lo] + ((x - xOld[lo]) / (xOld[hi] - xOld[lo])) * (yOld[hi] - yOld[lo]);
        else if (lo == hi) y = yOld[lo];
        else if (lo == nT) y = yOld[hi];
        else y = yOld[lo];
        const double a = fmin(fmax(y, 0.0), 1.0);
        alpha1[c] = a; alpha1old[c] = a;
    }
}

// fvc::domainIntegrate(qDot) (solutionControls.H:16)
__global__ void __launch_bounds__(BLK) k_power(int64_t n, const double* __restrict__ q, double V, RedWork red)
{
    double acc = 0;
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) acc += V * q[c];
    __shared__ double smem[32];
    double v[1] = {acc}; const int ops[1] = {0};
    grid_reduce<1>(v, ops, red, R_POWER, smem);
}

__global__ void __launch_bounds__(BLK) k_fill(int64_t n, double* __restrict__ a, double v)
{
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) a[c] = v;
}
__global__ void __launch_bounds__(BLK) k_reciprocal_of(int64_t n, const double* __restrict__ a, double* __restrict__ b)
{
    for (int64_t c = blockIdx.x * (int64_t)BLK + threadIdx.x; c < n; c += (int64_t)gridDim.x * BLK) b[c] = 1.0 / a[c];
}

// functionObjects/meltPoolDimensions/meltPoolDimensions.C:123-294 for one isovalue: bounding box (in the frame rotated by the
// scan path angle) of the isotherm located linearly between cell centres across internal and processor faces, plus the face centres
// of physical boundary faces whose cell or patch value reaches the isovalue.  Six min/max reductions -> red.result[R_MPOOL..+5].
__global__ void __launch_bounds__(BLK) k_melt_pool(BlockDims d, Geom g, SidesDev sd, const double* __restrict__ T, double iso, double cs,
                                                   double sn, RedWork red)
{
    const int nx = d.nx; const int64_t nxny = (int64_t)d.nx * d.ny;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    auto add = [&](double x, double y, double z) {
        const double qx = x * cs + y * sn, qy = y * cs - x * sn;
        lo[0] = fmin(qx, lo[0]); lo[1] = fmin(qy, lo[1]); lo[2] = fmin(z, lo[2]);
        hi[0] = fmax(qx, hi[0]); hi[1] = fmax(qy, hi[1]); hi[2] = fmax(z, hi[2]);
    };
    B200_FOR_CELLS(d)
    {
        B200_CELL_INDEX(d)
        const double To = T[c];
        const double cx = g.ox + (i + 0.5) * g.dx, cy = g.oy + (j + 0.5) * g.dy, cz = g.oz + (k + g.k0 + 0.5) * g.dz;
        auto cross = [&](double Tn, double nx_, double ny_, double nz_) {
            if (fmin(To, Tn) < iso && fmax(To, Tn) >= iso) {
                const double fr = iso - To, den = Tn - To;
                add(cx + (nx_ - cx) * fr / den, cy + (ny_ - cy) * fr / den, cz + (nz_ - cz) * fr / den);
            }
        };
        if (i < nx - 1) cross(T[c + 1], g.ox + (i + 1.5) * g.dx, cy, cz);
        if (j < d.ny - 1) cross(T[c + nx], cx, g.oy + (j + 1.5) * g.dy, cz);
        if (k < d.nz - 1) cross(T[c + nxny], cx, cy, g.oz + (k + g.k0 + 1.5) * g.dz);
        const bool bnd = i == 0 || i == nx - 1 || j == 0 || j == d.ny - 1 || k == 0 || k == d.nz - 1;
        for (int s = 0; bnd && s < 6; s++) {
            if (!on_side(d, s, i, j, k)) continue;
            const int type = sd.type[s];
            if (type == SIDE_PROCESSOR) {
                // both ranks treat their own patch face with themselves as the owner side
                if (s == B200_ZMIN) cross(T[c - nxny], cx, cy, g.oz + (k + g.k0 - 0.5) * g.dz);
                else cross(T[c + nxny], cx, cy, g.oz + (k + g.k0 + 1.5) * g.dz);
                continue;
            }
            const double Tb = type == B200_BC_ZERO_GRADIENT ? To : sd.Tb[s][side_index(d, s, i, j, k)];
            if (fmax(To, Tb) >= iso) {
                double fx = cx, fy = cy, fz = cz;
                switch (s) {
                case B200_XMIN: fx = g.ox + 0 * g.dx; break;
                case B200_XMAX: fx = g.ox + nx * g.dx; break;
                case B200_YMIN: fy = g.oy + 0 * g.dy; break;
                case B200_YMAX: fy = g.oy + d.ny * g.dy; break;
                case B200_ZMIN: fz = g.oz + (0 + g.k0) * g.dz; break;
                default: fz = g.oz + (d.nz + g.k0) * g.dz; break;
                }
                add(fx, fy, fz);
            }
        }
    }
    __shared__ double smem[32];
    double v[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
    const int ops[6] = {2, 2, 2, 1, 1, 1};
    grid_reduce<6>(v, ops, red, R_MPOOL, smem);
}

// functionObjects/solidificationData/solidificationData.C:140-187 over the index box of the overlap cells.  A cell that was above
// the isovalue at the old time and is at or below it now appends {x, y, z, t, |R|, max(G, small), |R|/G}: G = mag(fvc::grad(T))
// ("Gauss linear": sum of Sf*T_f over the six faces / V, patch value on physical boundary faces, w*own + (1-w)*nbr on processor
// faces), R the cooling rate kept from the previous execute.  R is then refreshed to (T - Told)/deltaT -- only where T is above
// the isovalue, the only cells whose R a later event can read.  Events land in arrival order; key = (execute index, cell)
// lets the reader restore the reference's append order.
struct SolidArgs {
    BoxRange box;
    double iso, time, rDeltaT;
    const double *T, *Told;
    double* R;
    double* events;                 // maxEvents x 7
    long long* keys;                // maxEvents
    unsigned long long* count;
    long long maxEvents, execIndex;
};

__global__ void __launch_bounds__(BLK) k_solidification(BlockDims d, Geom g, SidesDev sd, SolidArgs a)
{
    const int bx = a.box.i1 - a.box.i0, by = a.box.j1 - a.box.j0, bz = a.box.k1 - a.box.k0;
    const int nx = d.nx; const int64_t nxny = (int64_t)d.nx * d.ny;
    const int64_t total = (int64_t)bx * by * bz;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int i = a.box.i0 + (int)(q % bx);
        const int64_t rowq = q / bx;
        const int j = a.box.j0 + (int)(rowq % by), k = a.box.k0 + (int)(rowq / by);
        const int64_t c = i + (int64_t)nx * j + nxny * k;
        const double Tc = a.T[c], Tc0 = a.Told[c];
        if (Tc0 > a.iso && Tc <= a.iso) {
            auto faceValue = [&](int s) -> double {
                const int type = sd.type[s];
                if (type == SIDE_PROCESSOR) { const double Tn = a.T[s == B200_ZMIN ? c - nxny : c + nxny]; return 0.5 * Tc + (1.0 - 0.5) * Tn; }
                if (type == B200_BC_ZERO_GRADIENT) return Tc;
                return sd.Tb[s][side_index(d, s, i, j, k)];
            };
            const double Tw = i > 0 ? 0.5 * (a.T[c - 1] - Tc) + Tc : faceValue(B200_XMIN);
            const double Te = i < nx - 1 ? 0.5 * (Tc - a.T[c + 1]) + a.T[c + 1] : faceValue(B200_XMAX);
            const double Ts = j > 0 ? 0.5 * (a.T[c - nx] - Tc) + Tc : faceValue(B200_YMIN);
            const double Tn = j < d.ny - 1 ? 0.5 * (Tc - a.T[c + nx]) + a.T[c + nx] : faceValue(B200_YMAX);
            const double Tb = k > 0 ? 0.5 * (a.T[c - nxny] - Tc) + Tc : faceValue(B200_ZMIN);
            const double Tt = k < d.nz - 1 ? 0.5 * (Tc - a.T[c + nxny]) + a.T[c + nxny] : faceValue(B200_ZMAX);
            const double gx = (g.sfx * Te - g.sfx * Tw) / g.V, gy = (g.sfy * Tn - g.sfy * Ts) / g.V, gz = (g.sfz * Tt - g.sfz * Tb) / g.V;
            const double G = sqrt(gx * gx + gy * gy + gz * gz);
            const double Ri = fabs(a.R[c]), Gi = fmax(G, 1e-15);
            const unsigned long long slot = atomicAdd(a.count, 1ull);
            if ((long long)slot < a.maxEvents) {
                double* e = a.events + 7 * slot;
                e[0] = g.ox + (i + 0.5) * g.dx; e[1] = g.oy + (j + 0.5) * g.dy; e[2] = g.oz + (k + g.k0 + 0.5) * g.dz;
                e[3] = a.time; e[4] = Ri; e[5] = Gi; e[6] = Ri / Gi;
                a.keys[slot] = (a.execIndex << 40) | (long long)c;
            }
        }
        if (Tc > a.iso) a.R[c] = a.rDeltaT * (Tc - Tc0);
    }
}

} // namespace b200
