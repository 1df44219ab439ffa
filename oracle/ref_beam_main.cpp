// Driver around the REFERENCE's own beam kinematics, compiled where they lie:
//   movingHeatSource/movingBeam/movingBeam.C   constructor (:41-83), readPath (:89-148), activePath, move (:157-193),
//                                              findIndex (:196-225), adjustDeltaT (:228-256)
//   movingHeatSource/segment/segment.C         the scan-path row (:53-75)
// against the stand-in header oracle/ref_stubs/beam/fvCFD.H (what it emulates is listed there).  Requests on stdin:
//   move  <deltaT|-1> <hitPathIntervals> <startTime> <endTime> <pathFile> <n>  then n times
//        -> per time "x y z power active" (hex floats, active = activePath() with runTime at that time); the index carries over
//   adjust <deltaT|-1> <hitPathIntervals> <startTime> <endTime> <pathFile> <moveTo|-1> <now> <dt>
//        -> the Delta t after adjustDeltaT(dt) with runTime at `now` (after move(moveTo) when moveTo >= 0)
// Built by oracle/Makefile into oracle/_ref/refBeam when /root/reference is present.  TEST INFRASTRUCTURE ONLY: it produces
// tests/golden/ref_beam.json (tests/golden/make_ref_beam_golden.py).
#include "movingBeam.C"
#include "segment.C"

#include <cstdio>
#include <iostream>

const Foam::vector Foam::vector::zero(0, 0, 0);

static Foam::movingBeam* make(Foam::Time& rt, Foam::dictionary& dict, const Foam::word& name)
{
    double dT; int hit; std::string file;
    std::cin >> dT >> hit >> rt.now >> rt.end >> file;
    Foam::refEntries().clear();
    if (dT > 0) { char b[64]; std::snprintf(b, sizeof b, "%.17g", dT); Foam::refEntries()["deltaT"] = b; }
    Foam::refEntries()["hitPathIntervals"] = hit ? "1" : "0";
    const size_t cut = file.find_last_of('/');
    rt.dir = cut == std::string::npos ? "." : file.substr(0, cut);
    Foam::refEntries()["pathName"] = cut == std::string::npos ? file : file.substr(cut + 1);
    return new Foam::movingBeam(name, dict, rt);
}

int main()
{
    std::string what;
    const Foam::word name("beam");
    while (std::cin >> what) {
        Foam::Time rt;
        Foam::dictionary dict;
        if (what == "move") {
            Foam::movingBeam* b = make(rt, dict, name);
            int n; std::cin >> n;
            for (int i = 0; i < n; i++) {
                double t; std::cin >> t;
                b->move(t);
                rt.now = t;
                const Foam::vector p = b->position();
                std::printf("%a %a %a %a %d\n", p.x(), p.y(), p.z(), b->power(), b->activePath() ? 1 : 0);
            }
            delete b;
        } else if (what == "adjust") {
            Foam::movingBeam* b = make(rt, dict, name);
            double moveTo, now, dt;
            std::cin >> moveTo >> now >> dt;
            if (moveTo >= 0) b->move(moveTo);
            rt.now = now;
            b->adjustDeltaT(dt);
            std::printf("%a\n", dt);
            delete b;
        } else return 2;
    }
    return 0;
}
