// Third translation unit of oracle/_ref/refMHS: the REFERENCE's own Kelly absorption model, compiled where it lies
//   movingHeatSource/absorptionModels/Kelly/KellyAbsorption.C
// against oracle/ref_stubs/absorption (as in ref_kelly_main.cpp), behind two C entry points; like ref_mhs_beam.cpp its other
// symbols are hidden and localised, and its namespace is renamed on the command line (-DFoam=FoamOfKellyUnit: the typeinfo of
// its polymorphic stand-ins would otherwise share COMDAT groups with the main unit), so that its stand-in world never meets that
// of ref_mhs_main.cpp (oracle/Makefile).
// TEST INFRASTRUCTURE ONLY.
#include "KellyAbsorption.C"

#include <cstdio>

#define REF_EXPORT extern "C" __attribute__((visibility("default")))

REF_EXPORT void* refkelly_new(const char* geometry, double eta0, double etaMin)
{
    char a[64], b[64];
    std::snprintf(a, sizeof a, "%.17g", eta0);
    std::snprintf(b, sizeof b, "%.17g", etaMin);
    Foam::refEntries() = {{"geometry", geometry}, {"eta0", a}, {"etaMin", b}};
    static Foam::dictionary dict;
    static Foam::fvMesh mesh;
    return new Foam::absorptionModels::Kelly("beam", dict, mesh);
}
REF_EXPORT double refkelly_eta(void* h, double aspectRatio) { return ((Foam::absorptionModels::Kelly*)h)->eta(aspectRatio); }
