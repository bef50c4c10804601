#pragma once
#include <vector>
#include "openmm/Force.h"
namespace OpenMM {
// Like OpenMM's System: forces are heap objects the System takes ownership of (addForce(Force*)),
// so a plugin's own Force subclass survives and can be found again with dynamic_cast.
class System {
public:
    System() {}
    ~System() { for (size_t i = 0; i < forces.size(); i++) delete forces[i]; }
    int getNumParticles() const { return (int)masses.size(); }
    double getParticleMass(int i) const { return masses[i]; }
    int addParticle(double mass) { masses.push_back(mass); return (int)masses.size() - 1; }
    int getNumForces() const { return (int)forces.size(); }
    const Force& getForce(int i) const { return *forces[i]; }
    int addForce(Force* f) { forces.push_back(f); return (int)forces.size() - 1; }
    int getNumConstraints() const { return (int)consDist.size(); }
    int addConstraint(int p1, int p2, double d) { consP1.push_back(p1); consP2.push_back(p2); consDist.push_back(d); return (int)consDist.size() - 1; }
    void getConstraintParameters(int i, int& p1, int& p2, double& d) const { p1 = consP1[i]; p2 = consP2[i]; d = consDist[i]; }
private:
    System(const System&);
    System& operator=(const System&);
    std::vector<double> masses;
    std::vector<Force*> forces;
    std::vector<int> consP1, consP2;
    std::vector<double> consDist;
};
}  // namespace OpenMM
