#pragma once
#include "openmm/reference/SimTKOpenMMRealType.h"
namespace OpenMM {
// Noise comes from a queue the driver fills (zeros when it is empty), so a step is reproducible.
class SimTKOpenMMUtilities {
public:
    static void setRandomNumberSeed(unsigned int) {}
    static RealOpenMM getNormallyDistributedRandomNumber();
};
}  // namespace OpenMM
