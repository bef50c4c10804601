#pragma once
#include <vector>
#include "openmm/System.h"
#include "openmm/Vec3.h"
#include "openmm/reference/ReferenceConstraints.h"
#include "openmm/reference/SimTKOpenMMRealType.h"
namespace OpenMM {
class ReferenceDynamics {
public:
    ReferenceDynamics(int numberOfAtoms, RealOpenMM deltaT, RealOpenMM temperature)
        : numberOfAtoms(numberOfAtoms), timeStep(0), deltaT(deltaT), temperature(temperature), constraints(0) {}
    virtual ~ReferenceDynamics() {}
    int getNumberOfAtoms() const { return numberOfAtoms; }
    int getTimeStep() const { return timeStep; }
    int incrementTimeStep() { return ++timeStep; }
    RealOpenMM getDeltaT() const { return deltaT; }
    void setDeltaT(RealOpenMM dt) { deltaT = dt; }
    RealOpenMM getTemperature() const { return temperature; }
    ReferenceConstraintAlgorithm* getReferenceConstraintAlgorithm() const { return constraints; }
    void setReferenceConstraintAlgorithm(ReferenceConstraintAlgorithm* c) { constraints = c; }
private:
    int numberOfAtoms, timeStep;
    RealOpenMM deltaT, temperature;
    ReferenceConstraintAlgorithm* constraints;
};
}  // namespace OpenMM
