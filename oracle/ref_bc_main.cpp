// Driver around the REFERENCE's own top-surface boundary condition, compiled where it lies:
//   derivedFvPatchFields/mixedTemperature/mixedTemperatureFvPatchScalarField.C   dictionary constructor (:55-83), updateCoeffs (:157-188:
//                     refValue = Tinf, hEff = h + sigma eps (Tp^2 + Tinf^2)(Tp + Tinf), valueFraction = 1/(1 + kappa deltaCoeffs/max(hEff, 1e-15)))
// against its REAL header and the stand-ins of oracle/ref_stubs/bc.  Request on stdin:
//   mixed h emissivity Tinf n   then n rows "Tp kappa deltaCoeff"
// -> n rows "refValue refGrad valueFraction" (hex floats).
// Built by oracle/Makefile into oracle/_ref/refBC when /root/reference is present.  TEST INFRASTRUCTURE ONLY: it produces
// tests/golden/ref_bc.json (tests/golden/make_ref_bc_golden.py).
#include "mixedTemperatureFvPatchScalarField.C"

#include <cstdio>
#include <iostream>

int main()
{
    using namespace Foam;
    std::string what;
    while (std::cin >> what) {
        if (what != "mixed") return 2;
        std::string h, eps, Tinf;
        int n;
        std::cin >> h >> eps >> Tinf >> n;
        fvPatch patch;
        scalarField Tp(n, 0.0);
        patch.kappa_ = scalarField(n, 0.0); patch.deltaCoeffs_ = scalarField(n, 0.0);
        for (int i = 0; i < n; i++) std::cin >> Tp[i] >> patch.kappa_[i] >> patch.deltaCoeffs_[i];
        refEntries() = {{"h", h}, {"emissivity", eps}, {"Tinf", Tinf}, {"value", "300"}};
        dictionary dict;
        DimensionedField<scalar, volMesh> iF;
        mixedTemperatureFvPatchScalarField bc(patch, iF, dict);
        static_cast<fvPatchScalarField&>(bc) = Tp;                 // the patch values the condition is evaluated at
        bc.updateCoeffs();
        if (!bc.updated()) return 3;
        for (int i = 0; i < n; i++) std::printf("%a %a %a\n", bc.refValue()[i], bc.refGrad()[i], bc.valueFraction()[i]);
    }
    return 0;
}
