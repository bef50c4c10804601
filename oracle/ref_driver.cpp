// ref_driver.cpp -- C entry points around the REFERENCE's own sources, compiled where they lie
// under /root/reference (oracle/Makefile target _ref/libsdmref.so):
//     openmmapi/src/LangevinIntegratorSDM.cpp            (integrator: step sequence, SoftCoreF)
//     platforms/reference/src/ReferenceSDMKernels.cpp    (Save/Make/Restore state, execute)
//     platforms/reference/src/ReferenceSDMKernelFactory.cpp
//     platforms/reference/src/ReferenceStochasticDynamicsSDM.cpp
// OpenMM itself is not available in this image, so those files are compiled against the small
// stand-in headers in oracle/openmm_stub/ (our own declarations of the handful of OpenMM classes
// the plugin touches).  What OpenMM would compute -- the force-group evaluations behind
// ContextImpl::calcForcesAndEnergy -- is delegated to a callback the test supplies (the oracle's
// restated nonbonded evaluation), and the Gaussian noise comes from a queue the test fills.
// Everything the PLUGIN owns (SURVEY.md rows a1-a4, a6, a8, a9, a11-a15 and the Langevin update)
// is therefore executed by the reference's unmodified code.
//
// TEST INFRASTRUCTURE ONLY (like the rest of oracle/): used by tests/ and by
// tools/make_ref_golden.py to pin the restated oracle; never linked into libsdmb200.
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "LangevinIntegratorSDM.h"
#include "ReferenceSDMKernelFactory.h"
#include "SDMKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"
#include "openmm/reference/SimTKOpenMMUtilities.h"

extern "C" void registerKernelFactories();   // ReferenceSDMKernelFactory.cpp:44

namespace {
std::deque<double> g_noise;
std::string g_error;
}  // namespace

namespace OpenMM {

RealOpenMM SimTKOpenMMUtilities::getNormallyDistributedRandomNumber() {
    if (g_noise.empty()) return 0.0;
    const double v = g_noise.front();
    g_noise.pop_front();
    return v;
}

// What OpenMM's ContextImpl does here: clear the force buffer, evaluate the forces whose group is
// in the mask at the current positions, return their energy.
double ContextImpl::calcForcesAndEnergy(bool, bool, int groups) {
    ReferencePlatform::PlatformData* data = static_cast<ReferencePlatform::PlatformData*>(platformData);
    std::vector<Vec3>& pos = *data->positions;
    std::vector<Vec3>& frc = *data->forces;
    const int n = (int)pos.size();
    std::vector<double> p(3 * (size_t)n), f(3 * (size_t)n, 0.0);
    for (int i = 0; i < n; i++)
        for (int d = 0; d < 3; d++) p[3 * (size_t)i + d] = pos[i][d];
    const double e = cb(user, groups, n, p.data(), f.data());
    for (int i = 0; i < n; i++) frc[i] = Vec3(f[3 * (size_t)i], f[3 * (size_t)i + 1], f[3 * (size_t)i + 2]);
    return e;
}

}  // namespace OpenMM

extern "C" {

// Scalar state of LangevinIntegratorSDM, in the order of its setters (SDMplugin.i:86-145).
struct sdmref_params {
    double temperature, friction, step_size;
    int bias_method, softcore_method;
    double lambdac, gammac, wbcoeff, w0coeff, lambda1, lambda2, alpha, u0;
    double umax, acore, ubcore;
    int nonequilibrium, pad_;
    double noneq_tmax, work_value, time;
    double m_lambda1, m_lambda2, m_u0, m_w0, b_lambda1, b_lambda2, b_u0, b_w0;
};

struct sdmref_out {
    double bind_e, pot_energy, work_value, lambdac, lambda1, lambda2, u0, w0coeff, time, kinetic_energy;
    int step_count, pad_;
};

const char* sdmref_last_error() { return g_error.c_str(); }

void sdmref_set_noise(const double* values, int n) {
    g_noise.clear();
    for (int i = 0; i < n; i++) g_noise.push_back(values[i]);
}

// LangevinIntegratorSDM::SoftCoreF (LangevinIntegratorSDM.cpp:125-149).  Returns 0, or -1 when the
// reference throws (unknown method).
int sdmref_softcore(int method, double u, double umax, double a, double ub, double* u_sc, double* fp) {
    try {
        SDMPlugin::LangevinIntegratorSDM integ(300.0, 1.0, 0.001, 1);
        integ.setSoftCoreMethod(method);
        *u_sc = integ.SoftCoreF(u, umax, a, ub, *fp);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

// Defaults of the constructor (LangevinIntegratorSDM.cpp:48-85) as the reference sets them.
int sdmref_defaults(sdmref_params* p) {
    SDMPlugin::LangevinIntegratorSDM integ(300.0, 0.5, 0.001, 3);
    std::memset(p, 0, sizeof(*p));
    p->temperature = integ.getTemperature();
    p->friction = integ.getFriction();
    p->step_size = integ.getStepSize();
    p->bias_method = integ.getBiasMethod();
    p->softcore_method = integ.getSoftCoreMethod();
    p->lambdac = integ.getLambda();
    p->gammac = integ.getGamma();
    p->wbcoeff = integ.getWBcoeff();
    p->w0coeff = integ.getW0coeff();
    p->lambda1 = integ.getLambda1();
    p->lambda2 = integ.getLambda2();
    p->alpha = integ.getAlpha();
    p->u0 = integ.getU0();
    p->umax = integ.getUmax();
    p->acore = integ.getAcore();
    p->ubcore = integ.getUbcore();
    p->nonequilibrium = integ.getNonEquilibrium();
    p->work_value = integ.getNoneqWorkvalue();
    return 0;
}

// `steps` calls of LangevinIntegratorSDM::step(1) on a stand-in Context of n particles.
//   positions / velocities [3n]: in = initial state, out = state after the last step
//   displacement [3n]: the displacement map, set atom by atom through setDisplacement
//   force_groups [n_forces]: the force groups present in the System (the integrator rejects
//       anything but 1 and 2, LangevinIntegratorSDM.cpp:92-100)
//   hybrid_force [3n]: out, the force buffer after the last execute() (= the hybrid force)
//   traj (optional) [steps][2]: BindE and PotEnergy after every step
// Returns 0, or -1 with sdmref_last_error() set when the reference throws.
int sdmref_run(int n, const double* masses, double* positions, double* velocities,
               const double* displacement, const int* force_groups, int n_forces,
               const sdmref_params* p, OpenMM::StubForceCallback cb, void* user, int steps,
               sdmref_out* out, double* hybrid_force, double* traj) {
    using namespace OpenMM;
    try {
        static ReferencePlatform* platform = 0;
        if (!platform) {
            platform = new ReferencePlatform();
            Platform::registerPlatform(platform);
            registerKernelFactories();   // the reference's own plugin entry point
        }
        System system;
        for (int i = 0; i < n; i++) system.addParticle(masses[i]);
        for (int i = 0; i < n_forces; i++) system.addForce(Force(force_groups[i]));
        ReferencePlatform::PlatformData data(n);
        data.time = p->time;
        for (int i = 0; i < n; i++) {
            (*data.positions)[i] = Vec3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
            (*data.velocities)[i] = Vec3(velocities[3 * i], velocities[3 * i + 1], velocities[3 * i + 2]);
        }
        Context owner;
        ContextImpl impl(owner, system, *platform, &data, cb, user);

        SDMPlugin::LangevinIntegratorSDM integ(p->temperature, p->friction, p->step_size, n);
        integ.setBiasMethod(p->bias_method);
        integ.setSoftCoreMethod(p->softcore_method);
        integ.setLambda(p->lambdac);
        integ.setGamma(p->gammac);
        integ.setWBcoeff(p->wbcoeff);
        integ.setW0coeff(p->w0coeff);
        integ.setLambda1(p->lambda1);
        integ.setLambda2(p->lambda2);
        integ.setAlpha(p->alpha);
        integ.setU0(p->u0);
        integ.setUmax(p->umax);
        integ.setAcore(p->acore);
        integ.setUbcore(p->ubcore);
        integ.setNonEquilibrium(p->nonequilibrium);
        integ.setNoneqtmax(p->noneq_tmax);
        integ.setNoneqWorkvalue(p->work_value);
        integ.setlambda1Slope(p->m_lambda1);
        integ.setlambda2Slope(p->m_lambda2);
        integ.setu0Slope(p->m_u0);
        integ.setw0Slope(p->m_w0);
        integ.setlambda1intercept(p->b_lambda1);
        integ.setlambda2intercept(p->b_lambda2);
        integ.setu0intercept(p->b_u0);
        integ.setw0intercept(p->b_w0);
        for (int i = 0; i < n; i++)
            integ.setDisplacement(i, displacement[3 * i], displacement[3 * i + 1], displacement[3 * i + 2]);
        for (int i = 0; i < n; i++) {   // the map reads back what was written, atom by atom
            const Vec3 d = integ.getDisplacement(i);
            if (d[0] != displacement[3 * i] || d[1] != displacement[3 * i + 1] || d[2] != displacement[3 * i + 2])
                throw OpenMMException("getDisplacement does not return what setDisplacement stored");
        }

        impl.bindIntegrator(integ);   // LangevinIntegratorSDM::initialize -> kernel initialize (snapshot of the map)
        for (int s = 0; s < steps; s++) {
            integ.step(1);
            if (traj) {
                traj[2 * s] = integ.getBindE();
                traj[2 * s + 1] = integ.getPotEnergy();
            }
        }
        out->bind_e = integ.getBindE();
        out->pot_energy = integ.getPotEnergy();
        out->work_value = integ.getNoneqWorkvalue();
        out->lambdac = integ.getLambda();
        out->lambda1 = integ.getLambda1();
        out->lambda2 = integ.getLambda2();
        out->u0 = integ.getU0();
        out->w0coeff = integ.getW0coeff();
        out->time = data.time;
        out->step_count = data.stepCount;
        out->kinetic_energy = impl.kineticEnergy(integ);
        for (int i = 0; i < n; i++)
            for (int d = 0; d < 3; d++) {
                positions[3 * i + d] = (*data.positions)[i][d];
                velocities[3 * i + d] = (*data.velocities)[i][d];
                hybrid_force[3 * i + d] = (*data.forces)[i][d];
            }
        impl.releaseIntegrator(integ);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

}  // extern "C"
