/*---------------------------------------------------------------------------*\
  b200AdditiveFoam -- additiveFoam's time loop with the thermal step resident on the GPU.

  Seam N1: the `bPCG` solver shim accelerates only the Krylov part of a real OpenFOAM run and pays ~48 B/cell
  of PCIe traffic per solve.  This application keeps T, alpha.solid, alpha.powder, kappa, Cp, the matrix and the Krylov
  vectors in HBM for the whole run: it reads the SAME case directory as additiveFoam, fills a b200_case_desc, and replaces the
  body of the time loop (applications/solvers/additiveFoam/additiveFoam.C:66-95: updateProperties.H, setDeltaT.H,
  sources.update(), runTime++, solutionControls.H, thermo/TEqn.H) by b200_case_step().  Fields cross PCIe only at write
  times.  OpenFOAM stays in charge of everything around the step: argument parsing, the case dictionaries, the mesh, time
  control and function objects, and the output of T / alpha.powder in whatever writeFormat the case selects.

  What the resident step does not implement is refused up front instead of being approximated (the reference accepts more):
    * flow coupling: PIMPLE nOuterCorrectors > 0 (pU/UEqn.H, pU/pEqn.H, fvm::div(phi, T) with phi != 0)
    * ddtSchemes CrankNicolson with a time-dependent off-centring coefficient (a constant one is implemented, thermoScheme.H:51-65)
    * laplacian(kappa,T) other than Gauss harmonic, meshes that are not one uniform blockMesh hex block,
      patch types other than zeroGradient / fixedValue (uniform) / mixedTemperature, decompositions other than z-slabs

  Usage: as additiveFoam (b200AdditiveFoam [-parallel] in the case directory).  One MPI rank per GPU.

  NOT COMPILED HERE (no OpenFOAM-10 in this environment); syntax-checked against openfoam/stubs
  (tests/test_shim_signatures.py).  The library side of every call below is exercised by tests/test_foam_case.py, which
  builds the same b200_case_desc from the same dictionaries in Python (additivefoam_b200/foamcase.py).
\*---------------------------------------------------------------------------*/
#include "fvCFD.H"
#include "IFstream.H"
#include "graph.H"
#include "additivefoam_b200.h"

#include <cstring>
#include <vector>

using namespace Foam;

static void b200Check(const int rc, const char* what)
{
    if (rc != 0)
    {
        FatalErrorInFunction << what << ": " << b200_last_error() << exit(FatalError);
    }
}


static void refuse(const string& what)
{
    FatalErrorInFunction
        << "b200AdditiveFoam: " << what.c_str() << " -- not implemented by the resident step; run additiveFoam (with the bPCG solver) instead"
        << exit(FatalError);
}


// Polynomial<3> coefficients "(c0 c1 c2)" (readTransportProperties.H:18-27)
static void readPoly(const dictionary& d, const word& key, double out[3])
{
    const scalarList c(d.lookup(key));
    for (label i = 0; i < 3; i++) { out[i] = i < c.size() ? c[i] : 0.0; }
}


// the scan path file of one beam: header line, then "mode x y z power parameter" rows (movingBeam.C:100-149); the library's parser
// skips the header and blank lines exactly as movingBeam::readPath does
static void readScanPath(const fileName& file, std::vector<double>& rows)
{
    IFstream is(file);
    if (!is.good())
    {
        FatalErrorInFunction << "Cannot find file " << file << exit(FatalError);
    }
    std::string text;
    string line;
    while (is.good())
    {
        is.getLine(line);
        text += line + "\n";
    }
    const int n = b200_scan_path_parse(text.c_str(), nullptr, 0);
    if (n < 0)
    {
        FatalErrorInFunction << "Malformed scan path " << file << exit(FatalError);
    }
    rows.resize(6*size_t(n));
    b200_scan_path_parse(text.c_str(), rows.data(), n);
}


int main(int argc, char *argv[])
{
    #include "setRootCase.H"
    #include "createTime.H"
    #include "createMesh.H"

    // ---- the fields the case directory carries (createFields.H:2-29); alpha.solid is rebuilt from T (:74-78)
    volScalarField T
    (
        IOobject("T", runTime.timeName(), mesh, IOobject::MUST_READ, IOobject::AUTO_WRITE),
        mesh
    );
    volScalarField alpha3
    (
        IOobject(IOobject::groupName("alpha", "powder"), runTime.timeName(), mesh, IOobject::READ_IF_PRESENT, IOobject::AUTO_WRITE),
        mesh,
        dimensionedScalar("alpha.powder", dimless, 0.0),
        zeroGradientFvPatchScalarField::typeName
    );

    // ---- what the resident step does not do
    const dictionary& pimple = mesh.solutionDict().subDict("PIMPLE");
    if (pimple.lookupOrDefault<label>("nOuterCorrectors", 0) > 0)
    {
        refuse("PIMPLE nOuterCorrectors > 0 (flow coupling)");
    }
    ITstream& ddtStream = mesh.schemesDict().subDict("ddtSchemes").lookup("default");      // thermoScheme.H:2 (no per-field entry in the tutorials)
    const word ddtScheme(ddtStream);
    scalar ocCoeff = 1;                                  // [UPSTREAM] CrankNicolsonDdtScheme: psi = 1 when the entry gives none
    if (ddtScheme != "Euler" && ddtScheme != "backward" && ddtScheme != "CrankNicolson")
    {
        refuse("ddtSchemes other than Euler, backward and CrankNicolson");
    }
    if (ddtScheme == "CrankNicolson" && !ddtStream.eof())
    {
        token t(ddtStream);
        if (!t.isNumber())
        {
            refuse("CrankNicolson with an off-centring coefficient that is a function of time");
        }
        ocCoeff = t.number();
    }
    {
        const dictionary& lap = mesh.schemesDict().subDict("laplacianSchemes");
        const word key(lap.found("laplacian(kappa,T)") ? "laplacian(kappa,T)" : "default");
        ITstream& s = lap.lookup(key);
        const word gauss(s), interp(s);
        if (gauss != "Gauss" || interp != "harmonic")
        {
            refuse("laplacian(kappa,T) other than Gauss harmonic");
        }
    }

    b200_case_desc desc;
    std::memset(&desc, 0, sizeof(desc));

    // ---- mesh: one uniform blockMesh hex block; in parallel each rank one z-slab of it
    const boundBox bbGlobal(mesh.points(), true), bbLocal(mesh.points(), false);
    const vector c0 = mesh.cellCentres()[0];
    label nLocal[3];
    for (direction d = 0; d < 3; d++)
    {
        const scalar spacing = 2.0*(c0[d] - bbLocal.min()[d]);
        desc.mesh.min[d] = bbGlobal.min()[d];
        desc.mesh.max[d] = bbGlobal.max()[d];
        nLocal[d] = label(round(bbLocal.span()[d]/spacing));
        const label nGlobal = label(round(bbGlobal.span()[d]/spacing));
        (d == 0 ? desc.mesh.nx : d == 1 ? desc.mesh.ny : desc.mesh.nz) = nGlobal;
    }
    if
    (
        nLocal[0]*nLocal[1]*nLocal[2] != mesh.nCells()
     || nLocal[0] != desc.mesh.nx || nLocal[1] != desc.mesh.ny
     || mag(mesh.cellCentres()[mesh.nCells() - 1][0] - (bbLocal.max()[0] - (c0[0] - bbLocal.min()[0]))) > 1e-6*(c0[0] - bbLocal.min()[0])
    )
    {
        refuse("a mesh that is not one uniform blockMesh hex block cut into z-slabs (decomposePar: simple (1 1 N))");
    }

    // ---- patches of T: type and parameters from the boundary field, block sides from the face centres
    std::vector<b200_patch> patches;
    forAll(T.boundaryField(), patchi)
    {
        const fvPatchScalarField& pf = T.boundaryField()[patchi];
        if (pf.coupled()) { continue; }                 // processor patches: the library's ghost planes
        b200_patch p;
        std::memset(&p, 0, sizeof(p));
        const dictionary& pd = T.boundaryField()[patchi].patch().boundaryDict(T);    // the patch's entries in 0/T (stub: see fvCFD.H)
        const word type(pf.type());
        if (type == "zeroGradient") { p.type = B200_BC_ZERO_GRADIENT; }
        else if (type == "fixedValue")
        {
            p.type = B200_BC_FIXED_VALUE;
            p.value = pf.size() ? pf[0] : 0.0;
            forAll(pf, i) { if (pf[i] != p.value) { refuse("a non-uniform fixedValue patch"); } }
        }
        else if (type == "mixedTemperature")
        {
            p.type = B200_BC_MIXED_TEMPERATURE;
            p.h = pd.lookupOrDefault<scalar>("h", 0.0);
            p.emissivity = pd.lookupOrDefault<scalar>("emissivity", 0.0);
            p.Tinf = scalarField("Tinf", pd, 1)[0];
        }
        else { refuse("T patch type " + type); }
        // sides of the block this patch covers, in the order xMin xMax yMin yMax zMin zMax
        bool on[6] = {false, false, false, false, false, false};
        const vectorField& Cf = mesh.Cf().boundaryField()[patchi];
        const scalar tol = 1e-6*(c0[0] - bbLocal.min()[0]);
        forAll(Cf, i)
        {
            for (direction d = 0; d < 3; d++)
            {
                if (mag(Cf[i][d] - bbGlobal.min()[d]) < tol) { on[2*d] = true; }
                if (mag(Cf[i][d] - bbGlobal.max()[d]) < tol) { on[2*d + 1] = true; }
            }
        }
        for (label s = 0; s < 6; s++) { reduce(on[s], orOp<bool>()); if (on[s]) { p.sides[p.nSides++] = s; } }
        patches.push_back(p);
    }
    desc.nPatches = label(patches.size());
    desc.patches = patches.data();

    // ---- material: constant/transportProperties (readTransportProperties.H:13-73), constant/thermoPath (createFields.H:46-71)
    IOdictionary transportProperties
    (
        IOobject("transportProperties", runTime.constant(), mesh, IOobject::MUST_READ, IOobject::NO_WRITE)
    );
    readPoly(transportProperties.subDict("solid"), "kappa", desc.material.kappa[0]);
    readPoly(transportProperties.subDict("liquid"), "kappa", desc.material.kappa[1]);
    readPoly(transportProperties.subDict("powder"), "kappa", desc.material.kappa[2]);
    readPoly(transportProperties.subDict("solid"), "Cp", desc.material.Cp[0]);
    readPoly(transportProperties.subDict("liquid"), "Cp", desc.material.Cp[1]);
    readPoly(transportProperties.subDict("powder"), "Cp", desc.material.Cp[2]);
    desc.material.rho = dimensionedScalar("rho", dimDensity, transportProperties.lookup("rho")).value();
    desc.material.Lf = dimensionedScalar("Lf", dimEnergy/dimMass, transportProperties.lookup("Lf")).value();
    // the table is read the way createFields.H:46-57 reads it
    IFstream thermoFile(runTime.rootPath()/runTime.globalCaseName()/runTime.constant()/"thermoPath");
    graph thermo("thermo", "T", "alpha1", thermoFile);
    std::vector<double> thermoT(thermo.x().begin(), thermo.x().end()), thermoAlpha(thermo.y().begin(), thermo.y().end());
    desc.material.nThermo = label(thermoT.size());
    desc.material.thermoT = thermoT.data();
    desc.material.thermoAlpha = thermoAlpha.data();

    // ---- beams: constant/heatSourceDict (movingHeatSourceModel.C:34-83, heatSourceModel.C:89-118, movingBeam.C:44-75)
    IOdictionary heatSourceDict
    (
        IOobject("heatSourceDict", runTime.constant(), mesh, IOobject::MUST_READ, IOobject::NO_WRITE)
    );
    const wordList sourceNames(heatSourceDict.lookup("sources"));
    std::vector<b200_beam> beams(sourceNames.size());
    std::vector<std::vector<double>> paths(sourceNames.size());
    forAll(sourceNames, i)
    {
        const dictionary& sd = heatSourceDict.optionalSubDict(sourceNames[i]);
        b200_beam& b = beams[i];
        std::memset(&b, 0, sizeof(b));
        const word model(sd.lookup("heatSourceModel"));
        const dictionary& cd = sd.optionalSubDict(model + "Coeffs");
        if (model == "superGaussian") { b.heatSourceModel = B200_SUPER_GAUSSIAN; b.k = cd.lookup<scalar>("k"); }
        else if (model == "modifiedSuperGaussian") { b.heatSourceModel = B200_MODIFIED_SUPER_GAUSSIAN; b.k = cd.lookup<scalar>("k"); b.m = cd.lookup<scalar>("m"); }
        else if (model == "projectedGaussian") { b.heatSourceModel = B200_PROJECTED_GAUSSIAN; b.A = cd.lookup<scalar>("A"); b.B = cd.lookup<scalar>("B"); }
        else { refuse("heatSourceModel " + model); }
        const vector dims(cd.lookup("dimensions"));
        const labelVector nPoints(cd.lookup("nPoints"));
        for (direction d = 0; d < 3; d++) { b.dimensions[d] = dims[d]; b.nPoints[d] = nPoints[d]; }
        b.transient = cd.lookupOrDefault<Switch>("transient", false) ? 1 : 0;
        b.isoValue = b.transient ? cd.lookup<scalar>("isoValue") : 0.0;
        const word absorption(sd.lookup("absorptionModel"));
        const dictionary& ad = sd.optionalSubDict(absorption + "Coeffs");
        if (absorption == "constant") { b.absorptionModel = B200_ABSORPTION_CONSTANT; b.eta = ad.lookup<scalar>("eta"); }
        else if (absorption == "Kelly")
        {
            b.absorptionModel = B200_ABSORPTION_KELLY;
            b.kellyGeometry = word(ad.lookup("geometry")) == "cylinder" ? B200_KELLY_CYLINDER : B200_KELLY_CONE;
            b.eta0 = ad.lookup<scalar>("eta0");
            b.etaMin = ad.lookup<scalar>("etaMin");
        }
        else { refuse("absorptionModel " + absorption); }
        b.deltaT = sd.lookupOrDefault<scalar>("deltaT", -1.0);
        b.hitPathIntervals = sd.lookupOrDefault<bool>("hitPathIntervals", true) ? 1 : 0;
        readScanPath(runTime.rootPath()/runTime.globalCaseName()/runTime.constant()/word(sd.lookup("pathName")), paths[i]);
        b.nSegments = label(paths[i].size()/6);
        b.path = paths[i].data();
    }
    desc.nBeams = label(beams.size());
    desc.beams = beams.data();

    // ---- controls: system/controlDict (setDeltaT.H, [UPSTREAM] readTimeControls.H), fvSolution PIMPLE (solutionControls.H:2-12,31), solvers.T
    b200_controls& ct = desc.controls;
    const dictionary& cdict = runTime.controlDict();
    ct.startTime = runTime.value();
    ct.endTime = runTime.endTime().value();
    ct.deltaT = runTime.deltaTValue();
    ct.adjustTimeStep = cdict.lookupOrDefault<Switch>("adjustTimeStep", false) ? 1 : 0;
    ct.maxCo = cdict.lookupOrDefault<scalar>("maxCo", 1.0);
    ct.maxDi = cdict.lookupOrDefault<scalar>("maxDi", 10.0);
    ct.maxAlphaCo = cdict.lookupOrDefault<scalar>("maxAlphaCo", 1.0);
    ct.maxDeltaT = cdict.lookupOrDefault<scalar>("maxDeltaT", great);
    ct.writeInterval = word(cdict.lookup("writeControl")) == "adjustableRunTime" ? cdict.lookup<scalar>("writeInterval") : 0.0;
    ct.nThermoCorrectors = pimple.lookupOrDefault<label>("nThermoCorrectors", 20);
    ct.thermoTolerance = pimple.lookupOrDefault<scalar>("thermoTolerance", 1e-6);
    ct.Tmax = pimple.lookupOrDefault<scalar>("Tmax", -1.0);
    ct.explicitSolve = pimple.lookupOrDefault<Switch>("explicitSolve", false) ? 1 : 0;
    {
        const dictionary& sd = mesh.solverDict("T");
        const word solver(sd.lookup("solver")), precond(sd.lookupOrDefault<word>("preconditioner", "none"));
        if (solver == "PCG" || solver == "bPCG") { ct.solver = B200_SOLVER_PCG; }
        else if (solver == "PBiCGStab" || solver == "bPBiCGStab") { ct.solver = B200_SOLVER_PBICGSTAB; }
        else { refuse("T solver " + solver); }
        ct.preconditioner = precond == "DIC" ? B200_PRECOND_DIC : precond == "DILU" ? B200_PRECOND_DILU
                          : precond == "diagonal" ? B200_PRECOND_DIAGONAL : B200_PRECOND_NONE;
        ct.tolerance = sd.lookupOrDefault<scalar>("tolerance", 1e-6);
        ct.relTol = sd.lookupOrDefault<scalar>("relTol", 0.0);
        ct.minIter = sd.lookupOrDefault<label>("minIter", 0);
        ct.maxIter = sd.lookupOrDefault<label>("maxIter", 1000);
    }
    ct.ddtScheme = ddtScheme == "backward" ? B200_DDT_BACKWARD : ddtScheme == "CrankNicolson" ? B200_DDT_CRANK_NICOLSON : B200_DDT_EULER;
    ct.ocCoeff = ocCoeff;
    desc.Tinitial = gAverage(T.primitiveField());       // overwritten below by the field itself
    desc.powderZMin = 1e300;                            // alpha.powder comes from the file, see below

    // ---- one context per rank (= one GPU); the ranks exchange the NCCL id through Pstream
    unsigned char id[128] = {0};
    if (Pstream::parRun())
    {
        if (Pstream::master()) { b200Check(b200_nccl_unique_id(id), "b200_nccl_unique_id"); }
        List<unsigned char> idList(128);
        forAll(idList, i) { idList[i] = id[i]; }
        Pstream::scatter(idList);
        forAll(idList, i) { id[i] = idList[i]; }
    }
    int nDevices = 1;
    b200Check(b200_device_count(&nDevices), "b200_device_count");
    b200_ctx* ctx = nullptr;
    b200Check
    (
        b200_ctx_create
        (
            Pstream::parRun() ? Pstream::myProcNo() % nDevices : 0, Pstream::parRun() ? id : nullptr,
            Pstream::parRun() ? Pstream::myProcNo() : 0, Pstream::parRun() ? Pstream::nProcs() : 1, &ctx
        ),
        "b200_ctx_create"
    );
    b200_case* c = nullptr;
    b200Check(b200_case_create(ctx, &desc, &c), "b200_case_create");
    int64_t nLoc = 0, nGlob = 0;
    b200Check(b200_case_n_cells(c, &nLoc, &nGlob), "b200_case_n_cells");
    if (nLoc != mesh.nCells())
    {
        refuse("a decomposition whose z-slabs differ from the library's (nz*rank/nRanks layers per rank)");
    }
    // the start fields of the case directory (restart: any time directory); alpha.solid follows from T (createFields.H:74-78)
    b200Check(b200_case_set_field(c, B200_FIELD_ALPHA_POWDER, alpha3.primitiveField().begin()), "b200_case_set_field");
    b200Check(b200_case_set_field(c, B200_FIELD_T, T.primitiveField().begin()), "b200_case_set_field");
    b200Check(b200_case_reset_alpha_solid(c), "b200_case_reset_alpha_solid");

    Info<< "\nStarting time loop (thermal step resident on the GPU)\n" << endl;

    while (runTime.run())
    {
        b200_step_report rep;
        b200Check(b200_case_step(c, &rep), "b200_case_step");          // additiveFoam.C:68-90 for the thermal equation

        runTime.setDeltaT(rep.deltaT, false);                           // setDeltaT.H:49 (the library has already snapped it to the write time)
        runTime++;

        Info<< "Time = " << runTime.timeName() << nl
            << "deltaT = " << rep.deltaT << "  thermoCo = " << rep.alphaCoNum << "  Di = " << rep.DiNum
            << "  absorbed power = " << rep.totalPower << nl
            << "Thermo: " << rep.nThermoCorr << " correctors, residual " << rep.thermoResidual
            << ", T: " << rep.nSolveIterations << " Krylov iterations, final residual " << rep.lastFinalResidual << nl << endl;

        if (runTime.writeTime())
        {
            b200Check(b200_case_get_field(c, B200_FIELD_T, T.primitiveFieldRef().begin()), "b200_case_get_field");
            b200Check(b200_case_get_field(c, B200_FIELD_ALPHA_POWDER, alpha3.primitiveFieldRef().begin()), "b200_case_get_field");
            T.correctBoundaryConditions();
            alpha3.correctBoundaryConditions();
        }
        runTime.write();

        Info<< "ExecutionTime = " << runTime.elapsedCpuTime() << " s"
            << "  ClockTime = " << runTime.elapsedClockTime() << " s"
            << nl << endl;
    }

    b200Check(b200_case_destroy(c), "b200_case_destroy");
    b200Check(b200_ctx_destroy(ctx), "b200_ctx_destroy");
    return 0;
}
