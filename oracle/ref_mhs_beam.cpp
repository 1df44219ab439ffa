// Second translation unit of oracle/_ref/refMHS: the REFERENCE's own beam kinematics, compiled where they lie
//   movingHeatSource/movingBeam/movingBeam.C, movingHeatSource/segment/segment.C
// against oracle/ref_stubs/beam (as in ref_beam_main.cpp), behind a handful of C entry points.  Everything but those entry
// points is compiled with hidden visibility and then localised (objcopy --localize-hidden, oracle/Makefile), so that this
// unit's stand-in Foam::vector / Foam::Time / Foam::movingBeam never meet the differently shaped stand-ins of
// ref_stubs/hsm that ref_mhs_main.cpp is compiled against.  TEST INFRASTRUCTURE ONLY.
#include "movingBeam.C"
#include "segment.C"

const Foam::vector Foam::vector::zero(0, 0, 0);

namespace
{
struct Holder
{
    Foam::Time rt;
    Foam::dictionary dict;
    Foam::movingBeam* beam = nullptr;
};
}

#define REF_EXPORT extern "C" __attribute__((visibility("default")))

// deltaT <= 0: the dictionary has no deltaT entry (movingBeam.C takes its default)
REF_EXPORT void* refbeam_new(const char* pathFile, double deltaT, int hitPathIntervals, double startTime, double endTime)
{
    Holder* h = new Holder;
    h->rt.now = startTime; h->rt.end = endTime;
    Foam::refEntries().clear();
    if (deltaT > 0) { char b[64]; std::snprintf(b, sizeof b, "%.17g", deltaT); Foam::refEntries()["deltaT"] = b; }
    Foam::refEntries()["hitPathIntervals"] = hitPathIntervals ? "1" : "0";
    const std::string file(pathFile);
    const size_t cut = file.find_last_of('/');
    h->rt.dir = cut == std::string::npos ? "." : file.substr(0, cut);
    Foam::refEntries()["pathName"] = cut == std::string::npos ? file : file.substr(cut + 1);
    h->beam = new Foam::movingBeam("beam", h->dict, h->rt);
    return h;
}
REF_EXPORT void refbeam_set_time(void* h, double now) { ((Holder*)h)->rt.now = now; }
REF_EXPORT void refbeam_move(void* h, double t) { ((Holder*)h)->beam->move(t); }
REF_EXPORT void refbeam_position(void* h, double* p)
{
    const Foam::vector v = ((Holder*)h)->beam->position();
    p[0] = v.x(); p[1] = v.y(); p[2] = v.z();
}
REF_EXPORT double refbeam_power(void* h) { return ((Holder*)h)->beam->power(); }
REF_EXPORT double refbeam_deltaT(void* h) { return ((Holder*)h)->beam->deltaT(); }
REF_EXPORT int refbeam_active(void* h) { return ((Holder*)h)->beam->activePath() ? 1 : 0; }
REF_EXPORT void refbeam_adjust(void* h, double* dt) { ((Holder*)h)->beam->adjustDeltaT(*dt); }
