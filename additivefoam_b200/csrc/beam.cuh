// Host side of the moving heat source: scan path, beam kinematics, absorption, normalisation volume.
// Mirrors the reference's classes by name and behaviour
// (applications/solvers/additiveFoam/movingHeatSource = MHS/):
//   segment          MHS/segment/segment.{H,C}
//   movingBeam       MHS/movingBeam/movingBeam.{H,C}
//   absorptionModel  MHS/absorptionModels/{constant,Kelly}
//   heatSourceModel  MHS/heatSourceModels/{heatSourceModel,superGaussian,modifiedSuperGaussian,projectedGaussian}
// The per-cell integration (heatSourceModel::qDot, heatSourceModel.C:287-367) and the isotherm scan
// (updateDimensions, :150-213) run on the device (beam.cu); everything here is O(segments) scalar work that
// the reference also does once per rank.
#pragma once
#include "../../include/additivefoam_b200.h"
#include "common.cuh"
#include <vector>
#include <string>
#include <cmath>

namespace b200 {

struct segment {                       // MHS/segment/segment.H:55-70
    double mode = 1, x = 0, y = 0, z = 0, power = 0, parameter = 0, time = 0;
};

class movingBeam {
public:
    static constexpr double eps = 1e-10;                 // movingBeam.C:32
    movingBeam() = default;
    movingBeam(const b200_beam& b, double startTime, double endTimeOfRun);
    bool activePath(double now) const { return (endTime_ - now) > eps; }     // movingBeam.C:151-154
    void move(double time);                                                   // movingBeam.C:157-193
    void adjustDeltaT(double& dt, double now) const;                          // movingBeam.C:228-256
    int findIndex(double time) const;                                         // movingBeam.C:196-225
    const double* position() const { return position_; }
    void setPosition(const double p[3]) { position_[0] = p[0]; position_[1] = p[1]; position_[2] = p[2]; }   // diagnostic probes only
    double power() const { return power_; }
    double deltaT() const { return deltaT_; }
    double endTime() const { return endTime_; }
    int index() const { return index_; }
private:
    std::vector<segment> path_;
    int index_ = 0;
    double position_[3] = {0, 0, 0};
    double power_ = 0, endTime_ = 0, deltaT_ = 1e15;
    bool hitPathIntervals_ = true;
};

// device-side description of one beam shape evaluation
struct ShapeParams {
    int model;
    double dims[3];        // current dimensions_ (z may be the transient depth)
    double k, m;           // k already resolved for projectedGaussian
    double pos[3];
    int nPoints[3];
};

class heatSourceModel {
public:
    heatSourceModel() = default;
    heatSourceModel(const b200_beam& b, double startTime, double endTimeOfRun);
    double eta(double aspectRatio) const;                 // absorptionModel::eta
    double V0();                                          // shape normalisation (mutates k for projectedGaussian)
    double weight(const double d[3]) const;               // heatSourceModel::weight for a component-wise distance
    ShapeParams shapeParams() const;
    movingBeam beam;
    b200_beam spec{};
    double dimensions[3] = {0, 0, 0}, staticDimensions[3] = {0, 0, 0};
    double k = 0;
    bool transient = false;
};

// applications/utilities/createScanPath: file text of one rotation (scanpath.cu)
std::string create_scan_path(const double minPoint[2], const double maxPoint[2], double hatch, double rotationDeg, double power,
                             double speed, double dwellTime, bool biDirection);

} // namespace b200
