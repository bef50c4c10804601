#pragma once
#include "openmm/Context.h"
#include "openmm/Integrator.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
namespace OpenMM {
// The driver supplies the force evaluation OpenMM would do (group mask 4 = nonbonded, 2 = bonded).
typedef double (*StubForceCallback)(void* user, int groups, int n, const double* positions, double* forces);
class ContextImpl {
public:
    ContextImpl(Context& owner, const System& system, Platform& platform, void* platformData,
                StubForceCallback cb, void* user)
        : owner(owner), system(system), platform(platform), platformData(platformData), cb(cb), user(user) {}
    Context& getOwner() { return owner; }
    const System& getSystem() const { return system; }
    Platform& getPlatform() { return platform; }
    void* getPlatformData() { return platformData; }
    void updateContextState() {}
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = 0xFFFFFFFF);
    void bindIntegrator(Integrator& integ) { integ.initialize(*this); }
    void releaseIntegrator(Integrator& integ) { integ.cleanup(); }
    double kineticEnergy(Integrator& integ) { return integ.computeKineticEnergy(); }
private:
    Context& owner;
    const System& system;
    Platform& platform;
    void* platformData;
    StubForceCallback cb;
    void* user;
};
}  // namespace OpenMM
