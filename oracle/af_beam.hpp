// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
// CPU restatement of AdditiveFOAM's moving heat source library
// (applications/solvers/additiveFoam/movingHeatSource, "MHS/" below): scan path,
// beam kinematics, absorption models and the three beam shapes.  The reference ships no
// golden vectors, but this part of it compiles without OpenFOAM: PARITY PINNED on outputs
// of the reference's own sources built into oracle/_ref (oracle/Makefile) -- movingBeam.C,
// segment.C (refBeam, bit-exact), KellyAbsorption.C (refKelly, bit-exact), the three shape
// sources (refModels: V0 <= 2 ulp, weights 1e-13), frozen as tests/golden/ref_*.json and
// checked by tests/test_ref_beam.py and tests/test_ref_models.py; the derived values of
// SURVEY.md Appendix B stay as anchors in tests/test_oracle_anchors.py.
#pragma once
#include "af_ldu.hpp"
#include <sstream>
#include <climits>

namespace afo {

struct Vec3 {
    scalar x = 0, y = 0, z = 0;
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, scalar s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(scalar s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator/(Vec3 a, scalar s) { return {a.x / s, a.y / s, a.z / s}; }
inline scalar mag(Vec3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline scalar magSqr(Vec3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline Vec3 cmptDivide(Vec3 a, Vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline Vec3 cmptMag(Vec3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }

// MHS/segment/segment.C:53-75 ; mode is a floating scalar (segment.H:62)
struct Segment {
    scalar mode = 1, power = 0, parameter = 0, time = 0;
    Vec3 position;
};

// MHS/movingBeam/movingBeam.C
struct MovingBeam {
    static constexpr scalar eps = 1e-10;          // movingBeam.C:32
    std::vector<Segment> path;
    label index = 0;
    Vec3 position;
    scalar power = 0, endTime = 0, deltaT = great;
    bool hitPathIntervals = true;

    // movingBeam.C:42-84 (ctor) + readPath :89-148 -- table rows are the data lines of the file
    void init(const scalar* table, label nSeg, scalar beamDeltaT, bool hit, scalar runStart, scalar runEnd)
    {
        path.assign(1, Segment());                // synthetic zero point source, movingBeam.C:53
        deltaT = beamDeltaT > 0 ? beamDeltaT : great;
        hitPathIntervals = hit;
        for (label i = 0; i < nSeg; i++) {
            Segment s;
            s.mode = table[6 * i]; s.position = {table[6 * i + 1], table[6 * i + 2], table[6 * i + 3]};
            s.power = table[6 * i + 4]; s.parameter = table[6 * i + 5];
            path.push_back(s);
        }
        for (size_t i = 1; i < path.size(); i++) {             // movingBeam.C:127-145
            if (path[i].mode == 1) path[i].time = path[i - 1].time + path[i].parameter;
            else {
                const scalar d = mag(path[i].position - path[i - 1].position);
                path[i].time = path[i - 1].time + d / path[i].parameter;
            }
        }
        index = 0;
        index = findIndex(runStart);                           // movingBeam.C:69
        endTime = 0;
        for (label i = (label)path.size() - 1; i > 0; i--)      // movingBeam.C:74-81
            if (path[i].power > small_) { endTime = std::min(path[i].time, runEnd); break; }
    }
    bool activePath(scalar now) const { return (endTime - now) > eps; }   // movingBeam.C:151-154

    // movingBeam.C:196-225
    label findIndex(scalar time) const
    {
        label i = index;
        const label n = (label)path.size() - 1;
        for (; i > 0 && path[i].time > time; --i) {}
        for (; i < n && path[i].time < time; ++i) {}
        while (i < n) {
            if (path[i].mode == 1 && path[i].parameter == 0) ++i; else break;
        }
        return std::min(std::max(i, 0), n);
    }
    // movingBeam.C:157-193
    void move(scalar time)
    {
        index = findIndex(time);
        const label i = index;
        if (path[i].mode == 1) position = path[i].position;
        else {
            Vec3 displacement;
            const scalar dt = path[i].time - path[i - 1].time;
            if (dt > 0) {
                const Vec3 dx = path[i].position - path[i - 1].position;
                displacement = dx * (time - path[i - 1].time) / dt;
            }
            position = path[i - 1].position + displacement;
        }
        if ((time - path[i - 1].time) > eps) power = path[i].power;
        else power = path[i - 1].power;
    }
    // movingBeam.C:228-256
    void adjustDeltaT(scalar& dt, scalar now) const
    {
        if (activePath(now) && hitPathIntervals) {
            scalar timeToNextPath = 0;
            label i = index;
            while (timeToNextPath < eps) {
                timeToNextPath = std::max(0.0, path[i].time - now);
                i++;
                if (i == (label)path.size()) break;
            }
            const scalar nSteps = timeToNextPath / dt;
            if (nSteps < (scalar)INT32_MAX) {
                const label nStepsToNextPath = label(std::max(nSteps, 1.0) + 0.99);
                dt = std::min(timeToNextPath / nStepsToNextPath, dt);
            }
        }
    }
};

// MHS/absorptionModels/constant/constantAbsorption.H:86-92, Kelly/KellyAbsorption.C:67-99
struct Absorption {
    int model = 0;          // B200_ABSORPTION_*
    scalar etaConst = 0;
    int geometry = 0;       // B200_KELLY_*
    scalar eta0 = 0, etaMin = 0;
    scalar eta(scalar aspectRatio) const
    {
        if (model == 0) return etaConst;
        if (aspectRatio > 1.0) {
            const scalar theta = std::atan(1.0 / aspectRatio);
            scalar F = 0, G = 0;
            if (geometry == 0) {
                F = 0.25 * (3.0 * std::sin(theta) - std::sin(3.0 * theta));
                G = 1.0 / (1.0 + std::sqrt(1.0 + std::pow(aspectRatio, 2)));
            } else {
                F = 0.5 * (1.0 - std::cos(2.0 * theta));
                G = 0.5 / (1.0 + aspectRatio);
            }
            return eta0 * (1.0 + (1.0 - eta0) * (G - F)) / (1.0 - (1.0 - eta0) * (1.0 - G));
        }
        return etaMin;
    }
};

constexpr scalar pi_ = 3.14159265358979323846;

// MHS/heatSourceModels/{superGaussian,modifiedSuperGaussian,projectedGaussian}
struct BeamShape {
    int model = 0;                   // B200_SUPER_GAUSSIAN ...
    scalar k = 2, m = 2, A = 0, B = 0;
    Vec3 dimensions, staticDimensions;

    // projectedGaussian.C:46-69 (ctor) and :96-106 (inside V0): k from the current depth
    void updateProjectedK()
    {
        const scalar x = std::max(dimensions.z / std::min(staticDimensions.x, staticDimensions.y), 1.0);
        const scalar n = std::min(std::max(A * std::log2(x) + B, 0.0), 9.0);
        k = std::pow(2.0, n);
    }
    scalar weight(Vec3 d) const
    {
        if (model == 0) {            // superGaussian.C:62-70
            const Vec3 s = dimensions / std::pow(2.0, 1.0 / k);
            const scalar x = std::pow(magSqr(cmptDivide(d, s)), k / 2.0);
            return std::exp(-x);
        }
        if (model == 1) {            // modifiedSuperGaussian.C:63-84
            const scalar a = std::pow(2.0, 1.0 / k);
            Vec3 s = cmptDivide(dimensions, Vec3{a, a, 1.0});
            if (d.z < s.z) {
                s = s * std::pow(1.0 - std::pow(d.z / s.z, m), 1.0 / m);
                const Vec3 di{d.x, d.y, 0.0};
                const scalar x = std::pow(magSqr(cmptDivide(di, s)), k / 2.0);
                return std::exp(-x);
            }
            return 0.0;
        }
        // projectedGaussian.C:74-92
        const scalar sx = d.x / dimensions.x, sy = d.y / dimensions.y;
        const scalar f = std::exp(-2.0 * (sx * sx + sy * sy));
        const scalar s = std::exp(-3.0 * std::pow(std::fabs(std::fabs(d.z) / dimensions.z), k));
        return f * s;
    }
    scalar V0()
    {
        if (model == 0) {            // superGaussian.C:73-85
            const Vec3 s = dimensions / std::pow(2.0, 1.0 / k);
            return (2.0 / 3.0) * s.x * s.y * s.z * pi_ * std::tgamma(1.0 + 3.0 / k);
        }
        if (model == 1) {            // modifiedSuperGaussian.C:87-103
            const scalar a = std::pow(2.0, 1.0 / k);
            const Vec3 s = cmptDivide(dimensions, Vec3{a, a, 1.0});
            return s.x * s.y * s.z * pi_ * std::tgamma(1.0 + 2.0 / k)
                 * std::tgamma(1.0 + 1.0 / m) * std::tgamma(1.0 + 2.0 / m) / std::tgamma(1.0 + 3.0 / m);
        }
        updateProjectedK();          // projectedGaussian.C:96-106 mutates k_
        return 0.5 * pi_ * dimensions.x * dimensions.y * dimensions.z
             * std::tgamma(1.0 / k) / (k * std::pow(3.0, 1.0 / k));
    }
};

// [UPSTREAM] boundBox::overlaps (inclusive on all six planes)
struct Box {
    Vec3 mn, mx;
    bool overlaps(const Box& b) const
    {
        return b.mx.x >= mn.x && b.mn.x <= mx.x && b.mx.y >= mn.y && b.mn.y <= mx.y
            && b.mx.z >= mn.z && b.mn.z <= mx.z;
    }
};

// heatSourceModel.C:303-343 : Riemann sum of weight(d) over one hex cell
inline scalar integrateCell(const BeamShape& shape, const Box& cellBb, const Box& beamBb,
                            Vec3 position, const label nPoints[3], scalar cellVolume)
{
    const Vec3 span = cellBb.mx - cellBb.mn;
    Vec3 dx = cmptDivide(shape.dimensions, Vec3{(scalar)nPoints[0], (scalar)nPoints[1], (scalar)nPoints[2]});
    // labelVector(max(cmptDivide(span + small, dx), one)) truncates toward zero
    const label ncx = (label)std::max((span.x + small_) / dx.x, 1.0);
    const label ncy = (label)std::max((span.y + small_) / dx.y, 1.0);
    const label ncz = (label)std::max((span.z + small_) / dx.z, 1.0);
    dx = cmptDivide(span, Vec3{(scalar)ncx, (scalar)ncy, (scalar)ncz});
    const scalar dVi = dx.x * dx.y * dx.z;
    scalar wi = 0.0;
    for (label k = 0; k < ncz; ++k)
        for (label j = 0; j < ncy; ++j)
            for (label i = 0; i < ncx; ++i) {
                const Vec3 pt{cellBb.mx.x - (i + 0.5) * dx.x, cellBb.mx.y - (j + 0.5) * dx.y,
                              cellBb.mx.z - (k + 0.5) * dx.z};
                const Box ptBb{pt - 0.5 * dx, pt + 0.5 * dx};
                if (beamBb.overlaps(ptBb)) {
                    const Vec3 d = cmptMag(pt - position);
                    wi += shape.weight(d) * dVi;
                }
            }
    return wi / cellVolume;
}

// MHS/movingBeam/movingBeam.C:89-148: header line skipped, blank lines skipped
inline label parseScanPath(const char* text, scalar* out, label maxSeg)
{
    std::istringstream is(text);
    std::string line;
    if (!std::getline(is, line)) return 0;
    label n = 0;
    while (std::getline(is, line)) {
        if (line.empty()) continue;
        std::stringstream ls(line);
        scalar v[6] = {0, 0, 0, 0, 0, 0};
        ls >> v[0] >> v[1] >> v[2] >> v[3] >> v[4] >> v[5];
        if (n >= maxSeg) return -1;
        for (int q = 0; q < 6; q++) out[6 * n + q] = v[q];
        n++;
    }
    return n;
}

} // namespace afo
