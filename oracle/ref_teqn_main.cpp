// Driver around the REFERENCE's own temperature equation: thermo/TEqn.H -- which #includes thermo/thermoScheme.H (the equation
// for the chosen ddt scheme, explicit or implicit) and, inside its corrector loop, thermo/thermoSource.H (the liquid-fraction
// linearisation) -- included here VERBATIM where it lies, against the stand-ins of oracle/ref_stubs/teqn (what is emulated, and
// how, is listed in teqnStub.H; the lduMatrix solvers behind solve() are the oracle's restatement, af_ldu.hpp).
// Request on stdin (one per run):
//   teqn nx ny nz dx dy dz  ddtScheme explicitSolve  deltaT deltaT0  rho Lf Tliq Tsol thermoTol nThermoCorr Tmax
//        solver precond tolerance relTol minIter maxIter
//        nThermo (x y)...
//        nPatches { type nSides side... value h emissivity Tinf }...
//        nOldT nOldAlpha
//        psi timeIndex  existsT startT tiT  existsA startA tiA       (CrankNicolson: ocCoeff, the time index after runTime++, and for
//                                                                    ddt0(T) / ddt0(alpha.solid): exists, startTimeIndex, timeIndex())
//        fields: T, T.oldTime() [, T.oldTime().oldTime()], alpha1, [alpha1 old levels...], kappa, Cp, qDot, [ddt0(T)], [ddt0(alpha.solid)]
//        per patch: Tb[nFaces]
// -> "nSolves iters..." (one solve per corrector pass), then T and alpha1 (hex floats, one line each), then per patch Tb and valueFraction.
// Built by oracle/Makefile into oracle/_ref/refTEqn when /root/reference is present.  TEST INFRASTRUCTURE ONLY: it produces
// tests/golden/ref_teqn.json (tests/golden/make_ref_teqn_golden.py).
#include "teqnStub.H"

#include <cstdio>

using namespace Foam;

static int run(fvMesh& mesh, Time& runTime, volScalarField& T, volScalarField& alpha1, const volScalarField& kappa, const volScalarField& Cp,
               const surfaceScalarField& phi, movingHeatSourceModel& sources, const Graph& thermo, const dimensionedScalar& rho,
               const dimensionedScalar& Lf, const dimensionedScalar& Tliq, const dimensionedScalar& Tsol, const dimensionedScalar& Tmax,
               bool explicitSolve, scalar nThermoCorr, scalar thermoTol)
{
    // TEqn.H is one block: the equation (thermoScheme.H), then the corrector loop around thermoSource.H and solve()
    #include "thermo/TEqn.H"
    return 0;
}

static void addSide(PatchGeom& pa, int side, label nx, label ny, label nz, scalar dx, scalar dy, scalar dz)
{
    auto add = [&](label c, scalar area, scalar halfSpacing) { pa.faceCells.push_back(c); pa.magSf.push_back(area); pa.deltaCoeffs.push_back(1.0 / halfSpacing); };
    // block sides in blockMesh face order, as oracle/af_thermal.hpp buildMesh (sides 0..5 = xMin xMax yMin yMax zMin zMax)
    if (side == 0 || side == 1) { const label i = side == 0 ? 0 : nx - 1; for (label k = 0; k < nz; k++) for (label j = 0; j < ny; j++) add(i + nx * (j + ny * k), dy * dz, 0.5 * dx); }
    else if (side == 2 || side == 3) { const label j = side == 2 ? 0 : ny - 1; for (label i = 0; i < nx; i++) for (label k = 0; k < nz; k++) add(i + nx * (j + ny * k), dx * dz, 0.5 * dy); }
    else { const label k = side == 4 ? 0 : nz - 1; for (label i = 0; i < nx; i++) for (label j = 0; j < ny; j++) add(i + nx * (j + ny * k), dx * dy, 0.5 * dz); }
}

int main()
{
    std::string what;
    if (!(std::cin >> what) || what != "teqn") return 2;
    fvMesh mesh;
    Time runTime;
    mesh.time_ = &runTime;
    label nx, ny, nz;
    scalar dx, dy, dz, rho, Lf, Tliq, Tsol, thermoTol, nThermoCorr, Tmax;
    int explicitSolve;
    std::cin >> nx >> ny >> nz >> dx >> dy >> dz >> mesh.schemes_.ddt_ >> explicitSolve >> runTime.deltaT_ >> runTime.deltaT0_ >> rho >> Lf >> Tliq >> Tsol >> thermoTol
             >> nThermoCorr >> Tmax;
    afo::SolverControls& sc = mesh.solverControls;
    std::cin >> sc.solver >> sc.precond >> sc.tolerance >> sc.relTol >> sc.minIter >> sc.maxIter;
    Graph thermo;
    int nThermo; std::cin >> nThermo;
    for (int i = 0; i < nThermo; i++) { scalar x, y; std::cin >> x >> y; thermo.x_.push_back(x); thermo.y_.push_back(y); }
    const label n = nx * ny * nz;
    mesh.V_ = dx * dy * dz;
    mesh.cells_.assign(n, 0);
    afo::LduAddr& ad = mesh.addr;
    ad.nCells = n;
    for (label k = 0; k < nz; k++) for (label j = 0; j < ny; j++) for (label i = 0; i < nx; i++) {
        const label c = i + nx * (j + ny * k);
        if (i < nx - 1) { ad.lower.push_back(c); ad.upper.push_back(c + 1); mesh.magSf_.push_back(dy * dz); mesh.deltaCoeffs_.push_back(1.0 / dx); }
        if (j < ny - 1) { ad.lower.push_back(c); ad.upper.push_back(c + nx); mesh.magSf_.push_back(dx * dz); mesh.deltaCoeffs_.push_back(1.0 / dy); }
        if (k < nz - 1) { ad.lower.push_back(c); ad.upper.push_back(c + nx * ny); mesh.magSf_.push_back(dx * dy); mesh.deltaCoeffs_.push_back(1.0 / dz); }
    }
    ad.nFaces = (label)ad.lower.size();
    ad.finalize();
    int nPatches; std::cin >> nPatches;
    for (int p = 0; p < nPatches; p++) {
        PatchGeom pa; int nSides;
        std::cin >> pa.type >> nSides;
        for (int s = 0; s < nSides; s++) { int side; std::cin >> side; addSide(pa, side, nx, ny, nz, dx, dy, dz); }
        std::cin >> pa.fixedValue >> pa.h >> pa.emissivity >> pa.Tinf;
        mesh.patches.push_back(pa);
    }
    int nOldT, nOldA; std::cin >> nOldT >> nOldA;
    int cnT[3], cnA[3];
    std::cin >> mesh.schemes_.ocCoeff_ >> runTime.timeIndex_ >> cnT[0] >> cnT[1] >> cnT[2] >> cnA[0] >> cnA[1] >> cnA[2];
    auto readField = [&](volScalarField& f, const char* name) { f.mesh_ = &mesh; f.name_ = name; f.i_ = Internal(n, 0.0); for (label c = 0; c < n; c++) std::cin >> f.i_[c]; };
    volScalarField T, alpha1, kappa, Cp;
    movingHeatSourceModel sources;
    readField(T, "T");
    for (volScalarField* f = &T; nOldT > 0; nOldT--) { f->old_ = std::make_shared<volScalarField>(); readField(*f->old_, "T_0"); f = f->old_.get(); }
    readField(alpha1, "alpha.solid");
    for (volScalarField* f = &alpha1; nOldA > 0; nOldA--) { f->old_ = std::make_shared<volScalarField>(); readField(*f->old_, "alpha.solid_0"); f = f->old_.get(); }
    readField(kappa, "kappa"); readField(Cp, "Cp"); readField(sources.qDot_, "qDot");
    auto readDdt0 = [&](const char* name, const int (&st)[3]) {
        if (!st[0]) return;
        DDt0Field& f = mesh.ddt0s[name];
        f.i_.assign(n, 0.0); f.startTimeIndex_ = st[1]; f.timeIndex_ = st[2];
        for (label c = 0; c < n; c++) std::cin >> f.i_[c];
    };
    readDdt0("ddt0(T)", cnT); readDdt0("ddt0(alpha.solid)", cnA);
    std::vector<TPatchState> tp(nPatches);
    for (int p = 0; p < nPatches; p++) {
        const size_t nf = mesh.patches[p].faceCells.size();
        tp[p].Tb.assign(nf, 0.0); tp[p].valueFraction.assign(nf, 0.0);
        for (size_t i = 0; i < nf; i++) std::cin >> tp[p].Tb[i];
    }
    if (!std::cin) return 3;
    T.tPatches = &tp;
    theKappa() = &kappa;
    surfaceScalarField phi;
    try {
        run(mesh, runTime, T, alpha1, kappa, Cp, phi, sources, thermo, dimensionedScalar(rho), dimensionedScalar(Lf), dimensionedScalar(Tliq),
            dimensionedScalar(Tsol), dimensionedScalar(Tmax), explicitSolve != 0, nThermoCorr, thermoTol);
    } catch (const std::exception& e) { std::fprintf(stderr, "refTEqn: %s\n", e.what()); return 4; }
    std::printf("%d", (int)mesh.solveIterations.size());
    for (label it : mesh.solveIterations) std::printf(" %d", it);
    std::printf("\n");
    for (const volScalarField* f : {&T, &alpha1}) { for (label c = 0; c < n; c++) std::printf("%a ", f->i_[c]); std::printf("\n"); }
    for (int p = 0; p < nPatches; p++) {
        for (scalar v : tp[p].Tb) std::printf("%a ", v);
        std::printf("\n");
        for (scalar v : tp[p].valueFraction) std::printf("%a ", v);
        std::printf("\n");
    }
    return 0;
}
