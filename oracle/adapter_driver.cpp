// adapter_driver.cpp -- C entry point that runs the REFERENCE's unmodified integrator
// (openmmapi/src/LangevinIntegratorSDM.cpp, compiled where it lies under /root/reference) with the
// B200 kernel of openmm_sdm_plugin_b200/csrc/openmm/ registered for
// IntegrateLangevinStepSDMKernel::Name() -- the same stand-in Context, the same noise queue and the
// same parameter block as ref_driver.cpp, which runs it with the reference's own Reference-platform
// kernel.  tests/test_gpu_openmm_adapter.py compares the two.
//
// TEST INFRASTRUCTURE (like the rest of oracle/).  The adapter sources themselves are product code;
// this file only drives them against oracle/openmm_stub.  level = 0: the System carries a
// B200NonbondedForce (fused path, the callback is only asked for force group 1);  level = 1: the
// System carries plain forces and the callback evaluates group 2 as OpenMM would (literal operations).
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "B200NonbondedForce.h"
#include "LangevinIntegratorSDM.h"
#include "SDMKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"
#include "openmm/reference/SimTKOpenMMUtilities.h"

extern "C" void registerKernelFactories();   // B200SDMKernelFactory.cpp

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

struct sdmref_params {   // ref_driver.cpp
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
// force group 2 as the B200NonbondedForce carries it (level 0)
struct sdmb200_nonbonded {
    int method, n_exceptions, use_dispersion_correction, pad_;
    double cutoff, eps_rf, box[3];
    const double *charge, *sigma, *epsilon;       // [n]
    const int* exception_pairs;                   // [2*n_exceptions] every addException pair
    const double* exception_params;               // [3*n_exceptions] chargeProd, sigma, epsilon
    int lj_geometric, gb_ace;                     // createSystem(OPLS=True); GBSAHCTForce SA='ACE'
    const double *gb_charge, *gb_or, *gb_sr;      // [n] GBSAHCTForce parameters, or NULL: no implicit solvent
    double gb_solute, gb_solvent;
};

const char* sdmb200_adapter_last_error() { return g_error.c_str(); }

void sdmb200_adapter_set_noise(const double* values, int n) {
    g_noise.clear();
    for (int i = 0; i < n; i++) g_noise.push_back(values[i]);
}

int sdmb200_adapter_run(int level, int n, const double* masses, double* positions, double* velocities,
                        const double* displacement, const sdmb200_nonbonded* nb, int n_constraints,
                        const int* constraint_pairs, const double* constraint_dist, const sdmref_params* p,
                        OpenMM::StubForceCallback cb, void* user, int steps, sdmref_out* out,
                        double* hybrid_force, double* traj) {
    using namespace OpenMM;
    try {
        static ReferencePlatform* platform = 0;
        if (!platform) {
            platform = new ReferencePlatform();
            Platform::registerPlatform(platform);
            registerKernelFactories();   // the B200 plugin's entry point
        }
        System system;
        for (int i = 0; i < n; i++) system.addParticle(masses[i]);
        for (int k = 0; k < n_constraints; k++)
            system.addConstraint(constraint_pairs[2 * k], constraint_pairs[2 * k + 1], constraint_dist[k]);
        if (level == 0) {
            SDMB200::B200NonbondedForce* f = new SDMB200::B200NonbondedForce();
            for (int i = 0; i < n; i++) f->addParticle(nb->charge[i], nb->sigma[i], nb->epsilon[i]);
            for (int k = 0; k < nb->n_exceptions; k++)
                f->addException(nb->exception_pairs[2 * k], nb->exception_pairs[2 * k + 1], nb->exception_params[3 * k],
                                nb->exception_params[3 * k + 1], nb->exception_params[3 * k + 2]);
            f->setNonbondedMethod((SDMB200::B200NonbondedForce::NonbondedMethod)nb->method);
            f->setCutoffDistance(nb->cutoff);
            f->setReactionFieldDielectric(nb->eps_rf);
            f->setUseDispersionCorrection(nb->use_dispersion_correction != 0);
            f->setPeriodicBox(nb->box[0], nb->box[1], nb->box[2]);
            if (nb->lj_geometric) f->setCombiningRule(SDMB200::B200NonbondedForce::Geometric);
            if (nb->gb_or) {
                for (int i = 0; i < n; i++) f->addGBParticle(nb->gb_charge[i], nb->gb_or[i], nb->gb_sr[i]);
                f->setGBDielectrics(nb->gb_solute, nb->gb_solvent);
                f->setGBSurfaceAreaACE(nb->gb_ace != 0);
            }
            system.addForce(f);               // force group 2, evaluates to nothing on the OpenMM side
        } else {
            system.addForce(new Force(2));    // OpenMM's own NonbondedForce (the callback)
        }
        system.addForce(new Force(1));        // bonded + restraints
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

        impl.bindIntegrator(integ);   // LangevinIntegratorSDM::initialize -> createKernel -> B200 kernel initialize
        for (int s = 0; s < steps; s++) {
            integ.step(1);            // the reference's own step sequence, LangevinIntegratorSDM.cpp:153-183
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
