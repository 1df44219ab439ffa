// Driver around the REFERENCE's own beam-shape sources and table interpolation, compiled where they lie:
//   movingHeatSource/heatSourceModels/superGaussian/superGaussian.C            weight(), V0()   (:62-85)
//   movingHeatSource/heatSourceModels/modifiedSuperGaussian/modifiedSuperGaussian.C            (:63-103)
//   movingHeatSource/heatSourceModels/projectedGaussian/projectedGaussian.C    ctor, weight(), V0()  (:46-118)
//   interpolateXY/interpolateXY.C                                              interpolateXY<scalar>  (:33-109)
// against the stand-in headers of oracle/ref_stubs/models (what they emulate is listed there).  Reads requests on stdin,
// one per line, and answers with hex floats (exact):
//   shape <super|modified|projected> k m A B dimX dimY dimZ depth nPoints  then nPoints lines "dx dy dz"
//        -> "V0 <hex>" then one "<hex>" weight per point; V0() is called first, as heatSourceModel::qDot does (heatSourceModel.C:260)
//   table n  then n lines "x y", then m, then m lines "x"      -> one "<hex>" interpolateXY(x, X, Y) per query
// Built by oracle/Makefile into oracle/_ref/refModels when /root/reference is present.  TEST INFRASTRUCTURE ONLY: it produces
// tests/golden/ref_models.json (tests/golden/make_ref_models_golden.py).
#include "superGaussian.C"
#include "modifiedSuperGaussian.C"
#include "projectedGaussian.C"

#include <cstdio>
#include <iostream>
#include <utility>
#include <vector>

namespace Foam
{
template<class T> class Field : public std::vector<T> { public: label size() const { return label(std::vector<T>::size()); } };
typedef Field<scalar> scalarField;
class labelPair : public std::pair<label, label>
{
public:
    labelPair(label a, label b) : std::pair<label, label>(a, b) {}
    label first() const { return std::pair<label, label>::first; }
    label second() const { return std::pair<label, label>::second; }
};
}
#define NoRepository
#include "interpolateXY/interpolateXY.H"

int main()
{
    using namespace Foam;
    std::string what;
    while (std::cin >> what) {
        if (what == "shape") {
            std::string kind;
            double k, m, A, B, dx, dy, dz, depth;
            int n;
            std::cin >> kind >> k >> m >> A >> B >> dx >> dy >> dz >> depth >> n;
            refCoeffs() = {{"k", k}, {"m", m}, {"A", A}, {"B", B}, {"dimX", dx}, {"dimY", dy}, {"dimZ", dz}};
            dictionary dict;
            fvMesh mesh;
            heatSourceModel* model = nullptr;
            if (kind == "super") model = new heatSourceModels::superGaussian("beam", dict, mesh);
            else if (kind == "modified") model = new heatSourceModels::modifiedSuperGaussian("beam", dict, mesh);
            else model = new heatSourceModels::projectedGaussian("beam", dict, mesh);
            model->setDepth(depth);
            std::printf("V0 %a\n", model->V0().value());
            for (int i = 0; i < n; i++) {
                double x, y, z;
                std::cin >> x >> y >> z;
                std::printf("%a\n", model->weight(vector(x, y, z)));
            }
            delete model;
        } else if (what == "table") {
            int n, q;
            std::cin >> n;
            scalarField X, Y;
            for (int i = 0; i < n; i++) { double x, y; std::cin >> x >> y; X.push_back(x); Y.push_back(y); }
            std::cin >> q;
            for (int i = 0; i < q; i++) { double x; std::cin >> x; std::printf("%a\n", interpolateXY<scalar>(x, X, Y)); }
        } else return 2;
    }
    return 0;
}
