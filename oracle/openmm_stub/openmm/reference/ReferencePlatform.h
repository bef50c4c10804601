#pragma once
#include <vector>
#include "openmm/Platform.h"
#include "openmm/Vec3.h"
#include "openmm/reference/ReferenceConstraints.h"
namespace OpenMM {
class ReferencePlatform : public Platform {
public:
    class PlatformData {
    public:
        explicit PlatformData(int n)
            : numParticles(n), stepCount(0), time(0.0), positions(new std::vector<Vec3>(n)),
              velocities(new std::vector<Vec3>(n)), forces(new std::vector<Vec3>(n)),
              periodicBoxSize(new Vec3()), constraints(new ReferenceConstraints()) {}
        ~PlatformData() {
            delete positions; delete velocities; delete forces; delete periodicBoxSize; delete constraints;
        }
        int numParticles, stepCount;
        double time;
        std::vector<Vec3>* positions;
        std::vector<Vec3>* velocities;
        std::vector<Vec3>* forces;
        Vec3* periodicBoxSize;
        ReferenceConstraints* constraints;
    };
};
}  // namespace OpenMM
