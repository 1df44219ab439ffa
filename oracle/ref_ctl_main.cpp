// Driver around the REFERENCE's own time-step and solution controls, two code fragments the solver #includes into main(),
// included here VERBATIM where they lie:
//   setDeltaT.H          thermo number max|ddt(T)|/(Tliq - Tsol) dt, diffusion number, maxDeltaTFact/deltaTFact, sources.adjustDeltaT
//   solutionControls.H   nThermoCorrectors / thermoTolerance / Tmax defaults, absorbed power, beamOn / fluidInDomain (nThermoCorr = 1),
//                        the explicitSolve switch (diffusion number of the NEW deltaT < 1)
// against the stand-ins of oracle/ref_stubs/ctl (what is emulated, and how, is listed in ctlStub.H).  Request on stdin:
//   ctl nx ny nz minX minY minZ maxX maxY maxZ  rho Tliq Tsol deltaT  adjustTimeStep maxCo maxDeltaT
//       nControlDict (key value)...  nPimple (key value)...   then the fields T, T.oldTime(), kappa, Cp, alpha1, qDot (nx*ny*nz each)
// -> "alphaCoNum DiNum deltaT nThermoCorr thermoTol Tmax totalPower beamOn fluidInDomain explicitSolve" (hex floats / ints).
// Built by oracle/Makefile into oracle/_ref/refCtl when /root/reference is present.  TEST INFRASTRUCTURE ONLY: it produces
// tests/golden/ref_ctl.json (tests/golden/make_ref_ctl_golden.py).
#include "ctlStub.H"

#include <cstdio>
#include <iostream>

using namespace Foam;

struct Out { scalar alphaCoNum, DiNum, deltaT, nThermoCorr, thermoTol, Tmax, totalPower; int beamOn, fluidInDomain, explicitSolve; };

static Out run(const fvMesh& mesh, Time& runTime, pimpleControl& pimple, movingHeatSourceModel& sources, const volScalarField& T,
               const volScalarField& kappa, const volScalarField& Cp, const volScalarField& alpha1, const dimensionedScalar& rho,
               const dimensionedScalar& Tliq, const dimensionedScalar& Tsol, bool adjustTimeStep, scalar maxCo, scalar maxDeltaT)
{
    scalar DiNum = 0.0;
    scalar alphaCoNum = 0.0;
    const scalar CoNum = 0.0;           // [UPSTREAM] CourantNo.H with phi == 0
    Out o;
    {
        #include "setDeltaT.H"
    }
    o.alphaCoNum = alphaCoNum; o.DiNum = DiNum; o.deltaT = runTime.deltaTValue();
    // (additiveFoam.C:74-76: sources.update(), runTime++ -- the driver hands over this step's qDot)
    #include "solutionControls.H"
    o.nThermoCorr = nThermoCorr; o.thermoTol = thermoTol; o.Tmax = Tmax.value(); o.totalPower = totalPower;
    o.beamOn = beamOn; o.fluidInDomain = fluidInDomain; o.explicitSolve = explicitSolve ? 1 : 0;
    return o;
}

int main()
{
    std::string what;
    while (std::cin >> what) {
        if (what != "ctl") return 2;
        fvMesh mesh;
        double lo[3], hi[3], rho, Tliq, Tsol, dt, maxCo, maxDeltaT;
        int adjust, nC, nP;
        std::cin >> mesh.nx >> mesh.ny >> mesh.nz >> lo[0] >> lo[1] >> lo[2] >> hi[0] >> hi[1] >> hi[2] >> rho >> Tliq >> Tsol >> dt >> adjust >> maxCo >> maxDeltaT;
        Time runTime;
        pimpleControl pimple;
        std::cin >> nC; for (int i = 0; i < nC; i++) { std::string k, v; std::cin >> k >> v; runTime.controlDict_.e_[k] = v; }
        std::cin >> nP; for (int i = 0; i < nP; i++) { std::string k, v; std::cin >> k >> v; pimple.dict_.e_[k] = v; }
        const label nx = mesh.nx, ny = mesh.ny, nz = mesh.nz, n = nx * ny * nz;
        const scalar dx = (hi[0] - lo[0]) / nx, dy = (hi[1] - lo[1]) / ny, dz = (hi[2] - lo[2]) / nz;
        mesh.V_ = Internal(n, dx * dy * dz);
        // [UPSTREAM] blockMesh / fvMesh geometry of one uniform block, as oracle/af_thermal.hpp restates it
        for (label k = 0; k < nz; k++) for (label j = 0; j < ny; j++) for (label i = 0; i < nx; i++) {
            const label c = i + nx * (j + ny * k);
            if (i < nx - 1) { mesh.owner_.push_back(c); mesh.neighbour_.push_back(c + 1); mesh.magSf_.i_.push_back(dy * dz); mesh.deltaCoeffs_.i_.push_back(1.0 / dx); }
            if (j < ny - 1) { mesh.owner_.push_back(c); mesh.neighbour_.push_back(c + nx); mesh.magSf_.i_.push_back(dx * dz); mesh.deltaCoeffs_.i_.push_back(1.0 / dy); }
            if (k < nz - 1) { mesh.owner_.push_back(c); mesh.neighbour_.push_back(c + nx * ny); mesh.magSf_.i_.push_back(dx * dy); mesh.deltaCoeffs_.i_.push_back(1.0 / dz); }
        }
        // patches in the order of the case description the fixtures use: bottom (zMin), top (zMax), sides (xMin, xMax, yMin, yMax)
        mesh.patchFaceCells_.resize(3); mesh.magSf_.b_.resize(3); mesh.deltaCoeffs_.b_.resize(3);
        auto add = [&](int p, label c, scalar area, scalar half) { mesh.patchFaceCells_[p].push_back(c); mesh.magSf_.b_[p].push_back(area); mesh.deltaCoeffs_.b_[p].push_back(1.0 / half); };
        for (label i = 0; i < nx; i++) for (label j = 0; j < ny; j++) add(0, i + nx * j, dx * dy, 0.5 * dz);
        for (label i = 0; i < nx; i++) for (label j = 0; j < ny; j++) add(1, i + nx * (j + ny * (nz - 1)), dx * dy, 0.5 * dz);
        for (int side = 0; side < 2; side++) for (label k = 0; k < nz; k++) for (label j = 0; j < ny; j++) add(2, (side ? nx - 1 : 0) + nx * (j + ny * k), dy * dz, 0.5 * dx);
        for (int side = 0; side < 2; side++) for (label i = 0; i < nx; i++) for (label k = 0; k < nz; k++) add(2, i + nx * ((side ? ny - 1 : 0) + ny * k), dx * dz, 0.5 * dy);
        theMesh() = &mesh;
        theDeltaT() = dt;
        volScalarField T, Told, kappa, Cp, alpha1;
        movingHeatSourceModel sources;
        for (volScalarField* f : {&T, &Told, &kappa, &Cp, &alpha1, &sources.qDot_}) { f->i_ = Internal(n, 0.0); for (label c = 0; c < n; c++) std::cin >> f->i_[c]; }
        T.old_ = &Told;
        kappa.b_.resize(3);                                       // kappa patches are zeroGradient (createFields.H): face value = cell value
        for (int p = 0; p < 3; p++) for (label c : mesh.patchFaceCells_[p]) kappa.b_[p].push_back(kappa.i_[c]);
        const Out o = run(mesh, runTime, pimple, sources, T, kappa, Cp, alpha1, dimensionedScalar(rho), dimensionedScalar(Tliq), dimensionedScalar(Tsol),
                          adjust != 0, maxCo, maxDeltaT);
        std::printf("%a %a %a %a %a %a %a %d %d %d\n", o.alphaCoNum, o.DiNum, o.deltaT, o.nThermoCorr, o.thermoTol, o.Tmax, o.totalPower, o.beamOn, o.fluidInDomain, o.explicitSolve);
    }
    return 0;
}
