// Driver around the REFERENCE's own heat-source integration, compiled where it lies:
//   movingHeatSource/heatSourceModels/heatSourceModel/heatSourceModel.C     constructor (:75-118), updateDimensions (:123-219),
//                                                                           qDot (:221-371)
//   .../superGaussian/superGaussian.C, modifiedSuperGaussian.C, projectedGaussian.C   (weight, V0)
// against the REAL heatSourceModel.H and the stand-ins of oracle/ref_stubs/hsm (one uniform hex block for fvMesh; what else
// is emulated is listed in hsmStub.H).  Request on stdin:
//   qdot <super|modified|projected> k m A B  dimX dimY dimZ  depth  nPx nPy nPz  eta power  posX posY posZ  nx ny nz  minX minY minZ  maxX maxY maxZ
// -> "sum <hex>" (the Riemann sum of V*weight actually used is not exposed; printed is sum(V*qDot)), then nx*ny*nz hex floats qDot.
// Built by oracle/Makefile into oracle/_ref/refQDot when /root/reference is present.  TEST INFRASTRUCTURE ONLY: it produces
// tests/golden/ref_qdot.json (tests/golden/make_ref_qdot_golden.py).
#include "heatSourceModel.C"
#include "superGaussian.C"
#include "modifiedSuperGaussian.C"
#include "projectedGaussian.C"

#include <cstdio>
#include <iostream>

namespace Foam
{
static movingBeam theBeam;
static absorptionModel theAbsorption;
autoPtr<movingBeam> movingBeam::New(const word&, const dictionary&, const Time&) { return autoPtr<movingBeam>(&theBeam); }
autoPtr<absorptionModel> absorptionModel::New(const word&, const dictionary&, const fvMesh&) { return autoPtr<absorptionModel>(&theAbsorption); }

}
#include "hexBlock.H"

int main()
{
    using namespace Foam;
    std::string what, kind;
    while (std::cin >> what >> kind) {
        if (what == "depth") {
            // depth <modified|projected> k m A B dimX dimY dimZ isoValue posX posY posZ nx ny nz minX minY minZ maxX maxY maxZ  then nx*ny*nz temperatures
            // -> the dimensions after updateDimensions() (heatSourceModel.C:123-219)
            std::string k, m, A, B, dx, dy, dz, iso;
            double px, py, pz, lox, loy, loz, hix, hiy, hiz;
            int nx, ny, nz;
            std::cin >> k >> m >> A >> B >> dx >> dy >> dz >> iso >> px >> py >> pz >> nx >> ny >> nz >> lox >> loy >> loz >> hix >> hiy >> hiz;
            refEntries() = {{"k", k}, {"m", m}, {"A", A}, {"B", B}, {"dimensions", dx + " " + dy + " " + dz}, {"transient", "1"}, {"isoValue", iso}};
            fvMesh mesh;
            mesh.build(nx, ny, nz, vector(lox, loy, loz), vector(hix, hiy, hiz));
            theMesh = &mesh;
            volScalarField T(nx * ny * nz);
            forAll(T, i) std::cin >> T[i];
            mesh.T_ = &T;
            theBeam.position_ = vector(px, py, pz); theBeam.power_ = 1.0; theAbsorption.eta_ = 1.0;
            dictionary dict;
            heatSourceModel* model = nullptr;
            if (kind == "modified") model = new heatSourceModels::modifiedSuperGaussian("beam", dict, mesh);
            else model = new heatSourceModels::projectedGaussian("beam", dict, mesh);
            model->updateDimensions();
            const vector d = model->dimensions();
            std::printf("%a %a %a\n", d.x(), d.y(), d.z());
            delete model;
            continue;
        }
        if (what != "qdot") return 2;
        std::string k, m, A, B, dx, dy, dz, npx, npy, npz;
        double depth, eta, power, px, py, pz, lox, loy, loz, hix, hiy, hiz;
        int nx, ny, nz;
        std::cin >> k >> m >> A >> B >> dx >> dy >> dz >> depth >> npx >> npy >> npz >> eta >> power >> px >> py >> pz >> nx >> ny >> nz
                 >> lox >> loy >> loz >> hix >> hiy >> hiz;
        refEntries() = {{"k", k}, {"m", m}, {"A", A}, {"B", B}, {"dimensions", dx + " " + dy + " " + dz}, {"nPoints", npx + " " + npy + " " + npz}};
        fvMesh mesh;
        mesh.build(nx, ny, nz, vector(lox, loy, loz), vector(hix, hiy, hiz));
        theMesh = &mesh;
        theBeam.position_ = vector(px, py, pz); theBeam.power_ = power; theAbsorption.eta_ = eta;
        dictionary dict;
        heatSourceModel* model = nullptr;
        if (kind == "super") model = new heatSourceModels::superGaussian("beam", dict, mesh);
        else if (kind == "modified") model = new heatSourceModels::modifiedSuperGaussian("beam", dict, mesh);
        else model = new heatSourceModels::projectedGaussian("beam", dict, mesh);
        // a transient beam's current depth (what updateDimensions() would have set): through the public staticDimensions()/dimensions()
        // there is no setter, so the depth is given as the z of `dimensions` and the static xy stay what they are
        (void)depth;
        tmp<volScalarField> tq = model->qDot();
        const volScalarField& q = tq();
        double sum = 0;
        forAll(q, i) sum += mesh.V()[i] * q[i];
        std::printf("sum %a\n", sum);
        forAll(q, i) std::printf("%a\n", q[i]);
        delete model;
    }
    return 0;
}
