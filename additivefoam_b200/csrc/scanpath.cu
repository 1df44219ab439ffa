// Raster scan-path generator (host code): b200_create_scan_path.
// Replaces applications/utilities/createScanPath (createScanPath.C:74-101 drives one file per rotation;
// createFields.H:38-392 holds the geometry and the writer) -- SURVEY.md section 8f row N4, the data format in front of the
// heat-source path: its output is what movingBeam::readPath (MHS/movingBeam/movingBeam.C:89-148) and
// b200_scan_path_parse read.  Parity is pinned on the reference itself: the geometry header compiles without OpenFOAM
// (oracle/Makefile -> oracle/_ref/createScanPath_ref), tests/golden/scanpath holds its outputs and
// tests/test_scan_path.py asserts byte equality.  To print the same digits the arithmetic keeps the reference's operand
// order: hatch lines at mid.y -+ i*hatch, rotation about the box centre, line/edge intersection by the two-parameter
// form with the 1e-10 tolerance, intersections ordered by distance from the line's start, "%.10f" fields.
#include "beam.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace b200 {

namespace {

constexpr double kTol = 1e-10;          // createFields.H:43
constexpr double kFar = 1e10;           // half length of the uncropped hatch lines, createFields.H:286

struct P2 { double x, y; };
struct Seg { P2 a, b; };

// createFields.H:60-72
P2 turned(const P2& p, const P2& about, double degrees)
{
    const double angle = degrees * (M_PI / 180.0);
    const double s = std::sin(angle), c = std::cos(angle);
    const double tx = p.x - about.x, ty = p.y - about.y;
    return P2{tx * c - ty * s + about.x, tx * s + ty * c + about.y};
}

double gap(const P2& p, const P2& q)
{
    const double dx = p.x - q.x, dy = p.y - q.y;
    return std::sqrt(dx * dx + dy * dy);
}

// createFields.H:216-262: crossing of a box edge (first) with a hatch line (second); false when parallel or outside either
bool crossing(const Seg& edge, const Seg& line, P2& out)
{
    const double x1 = edge.a.x, y1 = edge.a.y, x2 = edge.b.x, y2 = edge.b.y;
    const double x3 = line.a.x, y3 = line.a.y, x4 = line.b.x, y4 = line.b.y;
    const double den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
    if (std::abs(den) < kTol) return false;
    const double t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / den;
    const double u = -((x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)) / den;
    if (!(t >= -kTol && t <= 1.0 + kTol && u >= -kTol && u <= 1.0 + kTol)) return false;
    out = P2{x1 + t * (x2 - x1), y1 + t * (y2 - y1)};
    return !(std::isnan(out.x) || std::isnan(out.y));
}

// createFields.H:327-342: the count comes from two accumulating loops, whose rounding decides borderline cases
int hatch_count(const P2& lo, const P2& hi, double step)
{
    int nX = 0, nY = 0;
    for (double x = lo.x; x <= hi.x; x += step) nX++;
    for (double y = lo.y; y <= hi.y; y += step) nY++;
    return nX > nY ? nX : nY;
}

void put(std::string& s, double v)
{
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.10f", v);      // std::fixed << std::setprecision(ceil(-log10(1e-10)))
    s += buf;
}

} // namespace

std::string create_scan_path(const double minPoint[2], const double maxPoint[2], double hatch, double rotationDeg, double power,
                             double speed, double dwellTime, bool biDirection)
{
    B200_CHECK(hatch > 0, "hatch spacing must be positive");
    B200_CHECK(maxPoint[0] >= minPoint[0] && maxPoint[1] >= minPoint[1], "maxPoint must not lie below minPoint");
    B200_CHECK((maxPoint[0] - minPoint[0]) / hatch < 1e7 && (maxPoint[1] - minPoint[1]) / hatch < 1e7, "more than 1e7 hatch lines");
    const P2 lo{minPoint[0], minPoint[1]}, hi{maxPoint[0], maxPoint[1]};
    const P2 mid{(lo.x + hi.x) / 2.0, (lo.y + hi.y) / 2.0};
    // left, right, top, bottom (createFields.H:152-160): the order fixes which of two equidistant crossings comes first
    const Seg edges[4] = {{{lo.x, lo.y}, {lo.x, hi.y}}, {{hi.x, lo.y}, {hi.x, hi.y}}, {{lo.x, hi.y}, {hi.x, hi.y}}, {{lo.x, lo.y}, {hi.x, lo.y}}};
    const int n = hatch_count(lo, hi, hatch);
    std::vector<Seg> kept;
    auto hatch_line = [&](double height) {
        Seg line{turned(P2{-kFar, height}, mid, rotationDeg), turned(P2{kFar, height}, mid, rotationDeg)};
        std::vector<P2> hits;
        for (const Seg& e : edges) {
            P2 h;
            if (crossing(e, line, h)) hits.push_back(h);
        }
        if (hits.empty()) return;
        std::sort(hits.begin(), hits.end(), [&line](const P2& p, const P2& q) { return gap(line.a, p) < gap(line.a, q); });
        kept.push_back(Seg{hits.front(), hits.back()});
    };
    for (int i = n - 1; i > 0; --i) hatch_line(mid.y - i * hatch);       // below the centre line, outermost first
    for (int i = 0; i < n; ++i) hatch_line(mid.y + i * hatch);           // the centre line and above
    std::string s = "Mode\tX(m)\tY(m)\tZ(m)\tPower(W)\ttParam\n";          // createFields.H:352
    for (size_t i = 0; i < kept.size(); i++) {
        const bool back = biDirection && i % 2 == 1;                       // odd lines run backwards
        const P2 first = back ? kept[i].b : kept[i].a, second = back ? kept[i].a : kept[i].b;
        s += "1\t"; put(s, first.x); s += "\t"; put(s, first.y); s += "\t0\t0\t";      // point source: the jump (and dwell)
        if (i == 0) s += "0"; else put(s, dwellTime);
        s += "\n0\t"; put(s, second.x); s += "\t"; put(s, second.y); s += "\t0\t";     // the raster line
        put(s, power); s += "\t"; put(s, speed); s += "\n";
    }
    return s;
}

} // namespace b200
