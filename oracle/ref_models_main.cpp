// Driver around the REFERENCE's own beam-shape sources and table interpolation, compiled where they lie:
//   movingHeatSource/heatSourceModels/superGaussian/superGaussian.C            weight(), V0()   (:62-85)
//   movingHeatSource/heatSourceModels/modifiedSuperGaussian/modifiedSuperGaussian.C            (:63-103)
//   movingHeatSource/heatSourceModels/projectedGaussian/projectedGaussian.C    ctor, weight(), V0()  (:46-118)
//   interpolateXY/interpolateXY.C                                              interpolateXY<scalar>  (:33-109)
//   updateProperties.H, thermo/thermoSource.H                                  the per-cell fragments, included verbatim
// against the stand-in headers of oracle/ref_stubs/models (what they emulate is listed there).  Reads requests on stdin,
// one per line, and answers with hex floats (exact):
//   shape <super|modified|projected> k m A B dimX dimY dimZ depth nPoints  then nPoints lines "dx dy dz"
//        -> "V0 <hex>" then one "<hex>" weight per point; V0() is called first, as heatSourceModel::qDot does (heatSourceModel.C:260)
//   table n  then n lines "x y", then m, then m lines "x"      -> one "<hex>" interpolateXY(x, X, Y) per query
//   props ... / source ...   the per-cell fragments updateProperties.H:1-25 and thermo/thermoSource.H:1-66 (formats below)
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

// ---- the two per-cell fragments of the time loop, included VERBATIM (they are code fragments the reference #includes into
// main(), additiveFoam.C:70 and thermo/TEqn.H:21): updateProperties.H and thermo/thermoSource.H.  The names they use are bound
// to stand-ins: fields are plain arrays with a no-op correctBoundaryConditions(), mesh.cells() only has a size,
// Tliq/Tsol are dimensionedScalars, thermo has the two columns of the table, kappa1..Cp3 are [UPSTREAM] Polynomial<3>
// (value(x): val = c0; powX = 1; for i >= 1: powX *= x, val += ci*powX).
#define forAll(list, i) for (Foam::label i = 0; i < (list).size(); i++)
namespace Foam
{
struct cellListStub { label n; label size() const { return n; } };
struct meshStub { cellListStub cells_; const cellListStub& cells() const { return cells_; } };
struct volScalarFieldStub : public std::vector<scalar> { void correctBoundaryConditions() {} };
struct graphStub { scalarField x_, y_; const scalarField& x() const { return x_; } const scalarField& y() const { return y_; } };
struct Polynomial3
{
    scalar v_[3];
    scalar value(const scalar x) const
    {
        scalar val = v_[0];
        scalar powX = 1;
        for (label i = 1; i < 3; ++i) { powX *= x; val += v_[i] * powX; }
        return val;
    }
};
}

static void runUpdateProperties(Foam::label n, Foam::volScalarFieldStub& T, Foam::volScalarFieldStub& alpha1, Foam::volScalarFieldStub& alpha3,
                                Foam::volScalarFieldStub& kappa, Foam::volScalarFieldStub& Cp, const Foam::dimensionedScalar& Tliq,
                                const Foam::dimensionedScalar& Tsol, const Foam::Polynomial3& kappa1, const Foam::Polynomial3& kappa2,
                                const Foam::Polynomial3& kappa3, const Foam::Polynomial3& Cp1, const Foam::Polynomial3& Cp2,
                                const Foam::Polynomial3& Cp3)
{
    using namespace Foam;
    meshStub mesh{{n}};
    #include "updateProperties.H"
}

static void runThermoSource(Foam::label nCells, Foam::volScalarFieldStub& T, Foam::volScalarFieldStub& alpha1, Foam::volScalarFieldStub& dFdT,
                            Foam::volScalarFieldStub& T0, const Foam::dimensionedScalar& Tliq, const Foam::dimensionedScalar& Tsol,
                            const Foam::scalar thermoTol, const Foam::graphStub& thermo)
{
    using namespace Foam;
    meshStub mesh{{nCells}};
    #include "thermo/thermoSource.H"
}

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
        } else if (what == "props") {
            // props n Tsol Tliq  then 18 polynomial coefficients (kappa solid/liquid/powder, Cp solid/liquid/powder), then n lines "T alpha1 alpha3"
            int n; double tsol, tliq;
            std::cin >> n >> tsol >> tliq;
            Polynomial3 p[6];
            for (auto& q : p) std::cin >> q.v_[0] >> q.v_[1] >> q.v_[2];
            volScalarFieldStub T, a1, a3, kappa, Cp;
            for (int i = 0; i < n; i++) { double t, x, y; std::cin >> t >> x >> y; T.push_back(t); a1.push_back(x); a3.push_back(y); }
            kappa.assign(n, 0.0); Cp.assign(n, 0.0);
            runUpdateProperties(n, T, a1, a3, kappa, Cp, dimensionedScalar("Tliq", dimVolume, tliq), dimensionedScalar("Tsol", dimVolume, tsol),
                                p[0], p[1], p[2], p[3], p[4], p[5]);
            for (int i = 0; i < n; i++) std::printf("%a %a %a\n", kappa[i], Cp[i], a3[i]);
        } else if (what == "source") {
            // source n Tliq Tsol thermoTol rows  then rows lines "T alpha", then n lines "T alpha1"
            int n, rows; double tliq, tsol, tol;
            std::cin >> n >> tliq >> tsol >> tol >> rows;
            graphStub thermo;
            for (int i = 0; i < rows; i++) { double x, y; std::cin >> x >> y; thermo.x_.push_back(x); thermo.y_.push_back(y); }
            volScalarFieldStub T, a1, dFdT, T0;
            for (int i = 0; i < n; i++) { double t, x; std::cin >> t >> x; T.push_back(t); a1.push_back(x); }
            dFdT.assign(n, 0.0); T0 = T;
            runThermoSource(n, T, a1, dFdT, T0, dimensionedScalar("Tliq", dimVolume, tliq), dimensionedScalar("Tsol", dimVolume, tsol), tol, thermo);
            for (int i = 0; i < n; i++) std::printf("%a %a\n", dFdT[i], T0[i]);
        } else return 2;
    }
    return 0;
}
