#pragma once
#include <vector>
#include "openmm/Force.h"
namespace OpenMM {
class System {
public:
    int getNumParticles() const { return (int)masses.size(); }
    double getParticleMass(int i) const { return masses[i]; }
    int addParticle(double mass) { masses.push_back(mass); return (int)masses.size() - 1; }
    int getNumForces() const { return (int)forces.size(); }
    const Force& getForce(int i) const { return forces[i]; }
    int addForce(const Force& f) { forces.push_back(f); return (int)forces.size() - 1; }
    int getNumConstraints() const { return 0; }
private:
    std::vector<double> masses;
    std::vector<Force> forces;
};
}  // namespace OpenMM
