#pragma once
#include <vector>
#include "openmm/System.h"
#include "openmm/Vec3.h"
namespace OpenMM {
class ReferenceVirtualSites {
public:
    static void computePositions(const System&, std::vector<Vec3>&) {}   // the stand-in System has none
};
}  // namespace OpenMM
