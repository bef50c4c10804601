#pragma once
#include <vector>
#include "openmm/Vec3.h"
#include "openmm/reference/SimTKOpenMMRealType.h"
namespace OpenMM {
class ReferenceConstraintAlgorithm {
public:
    virtual ~ReferenceConstraintAlgorithm() {}
    virtual void apply(std::vector<Vec3>&, std::vector<Vec3>&, std::vector<RealOpenMM>&, RealOpenMM) = 0;
    virtual void applyToVelocities(std::vector<Vec3>&, std::vector<Vec3>&, std::vector<RealOpenMM>&, RealOpenMM) = 0;
};
// The stand-in System has no constraints: both operations leave their arguments unchanged.
class ReferenceConstraints : public ReferenceConstraintAlgorithm {
public:
    void apply(std::vector<Vec3>&, std::vector<Vec3>&, std::vector<RealOpenMM>&, RealOpenMM) {}
    void applyToVelocities(std::vector<Vec3>&, std::vector<Vec3>&, std::vector<RealOpenMM>&, RealOpenMM) {}
};
}  // namespace OpenMM
