// Driver around the REFERENCE's own meltPoolDimensions function object, compiled where it lies:
//   functionObjects/meltPoolDimensions/meltPoolDimensions.C   constructor (:56-98), read (:107-113), execute (:123-286: the isotherm
//                                                             crossing on every internal face, boundary faces at or above the
//                                                             isotherm, rotation by scanPathAngle, the bounding box)
// against its REAL header and the stand-ins of oracle/ref_stubs/fo + ref_stubs/hsm (one uniform hex block).  Request on stdin:
//   meltpool nx ny nz minX minY minZ maxX maxY maxZ  time angleDeg  nIso iso...   then per patch (xMin xMax yMin yMax zMin zMax):
//   "z" (zeroGradient: face value = cell value) or "f <value>" (fixedValue), then nx*ny*nz temperatures
// -> per isotherm the record the function object writes to its log: "time,length,width,depth" (hex floats).
//   functionObjects/solidificationData/solidificationData.C   constructor (:55-83), read, setOverlapCells (:99-131), execute (:142-183:
//                                                             cells that cooled through the isotherm during the step; cooling rate of
//                                                             the PREVIOUS execute, |grad T|), end (:189-222: the csv rows)
//   solid nx ny nz minX..maxZ  boxMin(3) boxMax(3) isoValue  <6 patch kinds as above>  nSteps, the initial T, then per step: time deltaT T
// -> the text of data_0.csv, scalars as hex floats.
// Built by oracle/Makefile into oracle/_ref/refFO when /root/reference is present.  TEST INFRASTRUCTURE ONLY: it produces
// tests/golden/ref_fo.json (tests/golden/make_ref_fo_golden.py).
#include "meltPoolDimensions.C"
#include "solidificationData.C"

#include <iostream>

namespace Foam
{
// the sibling models of hsmStub.H are declared there but not used by the function objects
autoPtr<movingBeam> movingBeam::New(const word&, const dictionary&, const Time&) { return autoPtr<movingBeam>(nullptr); }
autoPtr<absorptionModel> absorptionModel::New(const word&, const dictionary&, const fvMesh&) { return autoPtr<absorptionModel>(nullptr); }
}
#include "hexBlock.H"

int main()
{
    using namespace Foam;
    std::string what;
    while (std::cin >> what) {
        if (what == "solid") {
            int nx, ny, nz, nSteps;
            double lox, loy, loz, hix, hiy, hiz;
            std::string b[6], iso;
            std::cin >> nx >> ny >> nz >> lox >> loy >> loz >> hix >> hiy >> hiz >> b[0] >> b[1] >> b[2] >> b[3] >> b[4] >> b[5] >> iso;
            refEntries() = {{"box", b[0] + " " + b[1] + " " + b[2] + " " + b[3] + " " + b[4] + " " + b[5]}, {"isoValue", iso}};
            fvMesh mesh;
            mesh.build(nx, ny, nz, vector(lox, loy, loz), vector(hix, hiy, hiz));
            mesh.buildPatches();
            theMesh = &mesh;
            std::vector<std::pair<bool, double>> bc(6);
            for (int p = 0; p < 6; p++) {
                std::string kind; std::cin >> kind;
                bc[p].first = kind == "f";
                if (bc[p].first) std::cin >> bc[p].second;
            }
            std::cin >> nSteps;
            volScalarField T(nx * ny * nz), Told(nx * ny * nz);
            auto readT = [&]() {
                forAll(T, i) std::cin >> T[i];
                T.bf_ = volScalarField::Boundary(6, fvPatchScalarField());
                for (int p = 0; p < 6; p++) {
                    fvPatchScalarField& pf = T.bf_[p];
                    pf.p_.fc_ = mesh.patchFaceCells_[p];
                    pf.internal_ = &T;
                    forAll(pf.p_.fc_, f) pf.values_.push_back(bc[p].first ? bc[p].second : T[pf.p_.fc_[f]]);
                }
            };
            readT();
            Told = static_cast<const scalarField&>(T);
            T.old_ = &Told;
            mesh.T_ = &T;
            mesh.time_.now = 0; mesh.time_.dt = 1;                       // construction: R_ = fvc::ddt(T) = 0 whatever deltaT is
            functionObjects::fvMeshFunctionObject::theMeshOfFunctionObjects() = &mesh;
            OFstream::all().clear();
            dictionary dict;
            functionObjects::solidificationData fo("solidificationData", mesh.time(), dict);
            for (int s = 0; s < nSteps; s++) {
                Told = static_cast<const scalarField&>(T);              // runTime++: T.oldTime() = T
                std::cin >> mesh.time_.now >> mesh.time_.dt;
                readT();
                fo.execute();
            }
            fo.end();
            std::printf("%s", OFstream::finished().back().c_str());
            continue;
        }
        if (what != "meltpool") return 2;
        int nx, ny, nz, nIso;
        double lox, loy, loz, hix, hiy, hiz, time;
        std::string angle, isos;
        std::cin >> nx >> ny >> nz >> lox >> loy >> loz >> hix >> hiy >> hiz >> time >> angle >> nIso;
        for (int i = 0; i < nIso; i++) { std::string v; std::cin >> v; isos += (i ? " " : "") + v; }
        refEntries() = {{"isoValues", isos}, {"scanPathAngle", angle}};
        fvMesh mesh;
        mesh.build(nx, ny, nz, vector(lox, loy, loz), vector(hix, hiy, hiz));
        mesh.buildPatches();
        theMesh = &mesh;
        mesh.time_.now = time;
        volScalarField T(nx * ny * nz);
        std::vector<std::pair<bool, double>> bc(6);
        for (int p = 0; p < 6; p++) {
            std::string kind; std::cin >> kind;
            bc[p].first = kind == "f";
            if (bc[p].first) std::cin >> bc[p].second;
        }
        forAll(T, i) std::cin >> T[i];
        T.bf_ = volScalarField::Boundary(6, fvPatchScalarField());
        for (int p = 0; p < 6; p++) {
            fvPatchScalarField& pf = T.bf_[p];
            pf.p_.fc_ = mesh.patchFaceCells_[p];
            pf.internal_ = &T;
            forAll(pf.p_.fc_, f) pf.values_.push_back(bc[p].first ? bc[p].second : T[pf.p_.fc_[f]]);
        }
        mesh.T_ = &T;
        functionObjects::fvMeshFunctionObject::theMeshOfFunctionObjects() = &mesh;
        OFstream::all().clear();
        dictionary dict;
        functionObjects::meltPoolDimensions fo("meltPoolDimensions", mesh.time(), dict);
        fo.execute();
        for (OFstream* os : OFstream::all()) std::printf("%s\n", os->records_.back().c_str());
    }
    return 0;
}
