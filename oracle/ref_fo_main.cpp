// Driver around the REFERENCE's own meltPoolDimensions function object, compiled where it lies:
//   functionObjects/meltPoolDimensions/meltPoolDimensions.C   constructor (:56-98), read (:107-113), execute (:123-286: the isotherm
//                                                             crossing on every internal face, boundary faces at or above the
//                                                             isotherm, rotation by scanPathAngle, the bounding box)
// against its REAL header and the stand-ins of oracle/ref_stubs/fo + ref_stubs/hsm (one uniform hex block).  Request on stdin:
//   meltpool nx ny nz minX minY minZ maxX maxY maxZ  time angleDeg  nIso iso...   then per patch (xMin xMax yMin yMax zMin zMax):
//   "z" (zeroGradient: face value = cell value) or "f <value>" (fixedValue), then nx*ny*nz temperatures
// -> per isotherm the record the function object writes to its log: "time,length,width,depth" (hex floats).
// Built by oracle/Makefile into oracle/_ref/refFO when /root/reference is present.  TEST INFRASTRUCTURE ONLY: it produces
// tests/golden/ref_fo.json (tests/golden/make_ref_fo_golden.py).
#include "meltPoolDimensions.C"

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
