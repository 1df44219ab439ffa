// Driver around the REFERENCE's own Kelly absorption model, compiled where it lies:
//   movingHeatSource/absorptionModels/Kelly/KellyAbsorption.C    constructor (:42-63), eta(aspectRatio) (:67-99)
// against oracle/ref_stubs/absorption/absorptionModel.H.  Requests on stdin: "kelly <cone|cylinder> eta0 etaMin n" then n aspect
// ratios -> one hex float eta per ratio.  Built by oracle/Makefile into oracle/_ref/refKelly when /root/reference is present.
// TEST INFRASTRUCTURE ONLY: it produces the "kelly" part of tests/golden/ref_models.json.
#include "KellyAbsorption.C"

#include <cstdio>
#include <iostream>

int main()
{
    std::string what, geometry, eta0, etaMin;
    int n;
    while (std::cin >> what >> geometry >> eta0 >> etaMin >> n) {
        if (what != "kelly") return 2;
        Foam::refEntries() = {{"geometry", geometry}, {"eta0", eta0}, {"etaMin", etaMin}};
        Foam::dictionary dict;
        Foam::fvMesh mesh;
        Foam::absorptionModels::Kelly model("beam", dict, mesh);
        for (int i = 0; i < n; i++) {
            double a;
            std::cin >> a;
            std::printf("%a\n", model.eta(a));
        }
    }
    return 0;
}
