// Driver around the REFERENCE's own scan-path generator, compiled where it lies:
//   #include "/root/reference/applications/utilities/createScanPath/createFields.H"   (structs Point, Line, BoundBox, Path)
// This file only replaces createScanPath.C's main(), which reads constant/createScanPathDict through OpenFOAM's IOdictionary:
// the dictionary entries arrive as command-line arguments instead, and the loop over rotations is the one of
// createScanPath.C:74-101 (rotation starts at 0, `angle` is added after each file, file i is <prefix>_<i>).
// Built by oracle/Makefile into oracle/_ref/createScanPath_ref when /root/reference is present.  TEST INFRASTRUCTURE ONLY:
// it produces the golden scan paths of tests/golden/scanpath (tests/golden/make_scanpath_golden.py).
#include "createFields.H"
#include <cstdlib>
#include <cstring>

int main(int argc, char** argv)
{
    if (argc != 13) {
        std::cerr << "usage: createScanPath_ref minX minY maxX maxY angle hatch nRotations power speed dwellTime biDirection prefix\n";
        return 2;
    }
    using namespace Foam;
    Point minPoint(std::atof(argv[1]), std::atof(argv[2]));
    Point maxPoint(std::atof(argv[3]), std::atof(argv[4]));
    const scalar angle = std::atof(argv[5]), hatch = std::atof(argv[6]);
    const label nRotations = std::atoi(argv[7]);
    const scalar power = std::atof(argv[8]), speed = std::atof(argv[9]), dwellTime = std::atof(argv[10]);
    const bool biDirection = std::atoi(argv[11]) != 0;
    BoundBox boundingBox(minPoint, maxPoint);
    scalar rotation = 0.0;
    for (label i = 0; i < nRotations; ++i) {
        Path path(boundingBox, hatch, rotation);
        path.power = power;
        path.speed = speed;
        path.dwellTime = dwellTime;
        path.write(std::string(argv[12]) + "_" + std::to_string(i), biDirection);
        rotation += angle;
    }
    return 0;
}
