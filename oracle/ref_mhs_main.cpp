// Driver around the REFERENCE's own multi-beam heat source, compiled where it lies:
//   movingHeatSource/movingHeatSourceModel/movingHeatSourceModel.C   constructor (:34-83), adjustDeltaT (:92-98), update (:100-150:
//                                                                    sub-cycling each beam over the time step, time-weighted mean)
//   movingHeatSource/heatSourceModels/heatSourceModel/heatSourceModel.C, heatSourceModelNew.C (run-time selection by the source's
//                                                                    `heatSourceModel` entry against each class's TypeName), the three shape sources
//   movingHeatSource/movingBeam/movingBeam.C, movingHeatSource/segment/segment.C                      (ref_mhs_beam.cpp)
// against the REAL movingHeatSourceModel.H / heatSourceModel.H and the stand-ins of oracle/ref_stubs/hsm.  Stand-in written
// here: absorptionModel::New (a constant eta, or the reference's own KellyAbsorption.C through ref_mhs_kelly.cpp).  Request on stdin:
//   mhs nx ny nz minX minY minZ maxX maxY maxZ  startTime endTime  nBeams nSteps
//   per beam:  name <super|modified|projected> k m A B  dimX dimY dimZ  nPx nPy nPz  <eta | kelly:<cone|cylinder>:eta0:etaMin>  beamDeltaT hitPathIntervals pathFile
//   per step:  proposedDeltaT
// Per step it does what setDeltaT.H:47-49 and additiveFoam.C:74-76 do with the sources: adjustDeltaT(deltaT), setDeltaT,
// update(), runTime++; and prints "dt <hex>", "n <nonzero cells>", then "<cell> <hex qDot>" rows.
// Built by oracle/Makefile into oracle/_ref/refMHS when /root/reference is present.  TEST INFRASTRUCTURE ONLY: it produces
// tests/golden/ref_mhs.json (tests/golden/make_ref_mhs_golden.py).
#include "movingHeatSourceModel.C"
#include "heatSourceModel.C"
#include "heatSourceModelNew.C"
#include "superGaussian.C"
#include "modifiedSuperGaussian.C"
#include "projectedGaussian.C"
#ifdef REF_MHS_WITH_B200
// oracle/_ref/refMHSb200: the GPU shim of the drop-in seam B2 compiled into the same world, so that the reference's own run-time
// selection (heatSourceModelNew.C) constructs it by its TypeName and movingHeatSourceModel::update() drives it (kinds b200super,
// b200modified, b200projected).  Links the product library: runs on the GPU box only (tests/test_gpu_shim_exec.py).
#include "b200HeatSources.C"
#endif

#include <iostream>

extern "C" void* refbeam_new(const char* pathFile, double deltaT, int hitPathIntervals, double startTime, double endTime);
extern "C" void* refkelly_new(const char* geometry, double eta0, double etaMin);

namespace Foam
{
static std::map<std::string, movingBeam*> theBeams;
static std::map<std::string, absorptionModel*> theAbsorptions;
autoPtr<movingBeam> movingBeam::New(const word& name, const dictionary&, const Time&) { return autoPtr<movingBeam>(theBeams.at(name)); }
autoPtr<absorptionModel> absorptionModel::New(const word& name, const dictionary&, const fvMesh&) { return autoPtr<absorptionModel>(theAbsorptions.at(name)); }

}
#include "hexBlock.H"

int main()
{
    using namespace Foam;
    std::string what;
    while (std::cin >> what) {
        if (what != "mhs") return 2;
        int nx, ny, nz, nBeams, nSteps;
        double lox, loy, loz, hix, hiy, hiz, startTime, endTime;
        std::cin >> nx >> ny >> nz >> lox >> loy >> loz >> hix >> hiy >> hiz >> startTime >> endTime >> nBeams >> nSteps;
        refEntries().clear(); theBeams.clear(); theAbsorptions.clear();
        std::string names;
        for (int b = 0; b < nBeams; b++) {
            std::string name, kind, k, m, A, B, dx, dy, dz, npx, npy, npz, file;
            std::string eta; double beamDt; int hit;
            std::cin >> name >> kind >> k >> m >> A >> B >> dx >> dy >> dz >> npx >> npy >> npz >> eta >> beamDt >> hit >> file;
            const bool gpu = kind.rfind("b200", 0) == 0;
            const std::string base = gpu ? kind.substr(4) : kind;
            std::string type = base == "super" ? "superGaussian" : base == "modified" ? "modifiedSuperGaussian" : "projectedGaussian";
            if (gpu) type = "b200" + std::string(1, (char)std::toupper(type[0])) + type.substr(1);
            const std::string p = name + ".", c = name + "." + type + "Coeffs.";
            refEntries()[p + "heatSourceModel"] = type;
            refEntries()[c + "k"] = k; refEntries()[c + "m"] = m; refEntries()[c + "A"] = A; refEntries()[c + "B"] = B;
            refEntries()[c + "dimensions"] = dx + " " + dy + " " + dz;
            refEntries()[c + "nPoints"] = npx + " " + npy + " " + npz;
            theBeams[name] = new movingBeam(refbeam_new(file.c_str(), beamDt, hit, startTime, endTime));
            theAbsorptions[name] = new absorptionModel;
            if (eta.rfind("kelly:", 0) == 0) {
                const size_t a = eta.find(':', 6), b = eta.find(':', a + 1);
                theAbsorptions[name]->kelly_ = refkelly_new(eta.substr(6, a - 6).c_str(), std::atof(eta.substr(a + 1, b - a - 1).c_str()), std::atof(eta.substr(b + 1).c_str()));
            } else theAbsorptions[name]->eta_ = std::atof(eta.c_str());
            names += (b ? " " : "") + name;
        }
        refEntries()["sources"] = names;
        fvMesh mesh;
        mesh.build(nx, ny, nz, vector(lox, loy, loz), vector(hix, hiy, hiz));
        theMesh = &mesh;
        mesh.time_.now = startTime;
        movingHeatSourceModel sources(mesh);
        for (int s = 0; s < nSteps; s++) {
            double dt;
            std::cin >> dt;
            for (auto& kv : theBeams) refbeam_set_time(kv.second->handle(), mesh.time_.now);
            sources.adjustDeltaT(dt);                     // setDeltaT.H:47
            mesh.time_.dt = dt;                           // setDeltaT.H:49
            sources.update();                             // additiveFoam.C:74
            mesh.time_.now += dt;                         // additiveFoam.C:76
            const volScalarField& q = sources.qDot();
            int n = 0;
            forAll(q, i) if (q[i] != 0.0) n++;
            std::printf("dt %a\nn %d\n", dt, n);
            forAll(q, i) if (q[i] != 0.0) std::printf("%d %a\n", i, q[i]);
        }
    }
    return 0;
}
