// B200SDMKernels.cpp -- see B200SDMKernels.h.  Host glue only: every number the step produces comes
// out of libsdmb200 (CUDA, sm_100a); there is no CPU evaluation path in here.
#include "B200SDMKernels.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>

#include "B200NonbondedForce.h"
#include "LangevinIntegratorSDM.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/SimTKOpenMMUtilities.h"

using namespace OpenMM;
using namespace SDMPlugin;

namespace SDMB200 {

// ---- the data interface (ReferencePlatform::PlatformData, like ReferenceSDMKernels.cpp:78-101) ------
static std::vector<Vec3>& extractPositions(ContextImpl& context) {
    return *static_cast<ReferencePlatform::PlatformData*>(context.getPlatformData())->positions;
}
static std::vector<Vec3>& extractVelocities(ContextImpl& context) {
    return *static_cast<ReferencePlatform::PlatformData*>(context.getPlatformData())->velocities;
}
static std::vector<Vec3>& extractForces(ContextImpl& context) {
    return *static_cast<ReferencePlatform::PlatformData*>(context.getPlatformData())->forces;
}
static void flatten(const std::vector<Vec3>& v, std::vector<double>& out) {
    out.resize(3 * v.size());
    for (size_t i = 0; i < v.size(); i++)
        for (int d = 0; d < 3; d++) out[3 * i + d] = v[i][d];
}
static void unflatten(const std::vector<double>& in, std::vector<Vec3>& v) {
    for (size_t i = 0; i < v.size(); i++) v[i] = Vec3(in[3 * i], in[3 * i + 1], in[3 * i + 2]);
}

static void check(int rc, const char* what) {
    if (rc != SDM_OK)
        throw OpenMMException((std::string("libsdmb200: ") + what + ": " + sdm_last_error()).c_str());
}
static void checkCuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw OpenMMException((std::string("CUDA: ") + what + ": " + cudaGetErrorString(e)).c_str());
}

// The integrator's alchemical state as the C ABI carries it (sdm_alch mirrors the getters of
// openmmapi/include/LangevinIntegratorSDM.h:150-472 one to one).
static sdm_alch alchFromIntegrator(const LangevinIntegratorSDM& integ, double time) {
    sdm_alch a;
    sdm_default_alch(&a);
    a.bias_method = integ.getBiasMethod();
    a.softcore_method = integ.getSoftCoreMethod();
    a.lambdac = integ.getLambda();
    a.gammac = integ.getGamma();
    a.wbcoeff = integ.getWBcoeff();
    a.w0coeff = integ.getW0coeff();
    a.lambda1 = integ.getLambda1();
    a.lambda2 = integ.getLambda2();
    a.alpha = integ.getAlpha();
    a.u0 = integ.getU0();
    a.umax = integ.getUmax();
    a.acore = integ.getAcore();
    a.ubcore = integ.getUbcore();
    a.nonequilibrium = integ.getNonEquilibrium();
    a.noneq_tmax = integ.getNoneqtmax();
    a.work_value = integ.getNoneqWorkvalue();
    a.time = time;
    a.step_size = integ.getStepSize();
    a.m_lambda1 = integ.getlambda1Slope();
    a.m_lambda2 = integ.getlambda2Slope();
    a.m_u0 = integ.getu0Slope();
    a.m_w0 = integ.getw0Slope();
    a.b_lambda1 = integ.getlambda1intercept();
    a.b_lambda2 = integ.getlambda2intercept();
    a.b_u0 = integ.getu0intercept();
    a.b_w0 = integ.getw0intercept();
    return a;
}

// What ReferenceSDMKernels.cpp:232-247,283-305 writes back into the integrator.
static void alchToIntegrator(const sdm_alch& a, const sdm_scalars& sc, LangevinIntegratorSDM& integ) {
    if (a.nonequilibrium == 1) {
        integ.setLambda(a.lambdac);
        integ.setLambda1(a.lambda1);
        integ.setLambda2(a.lambda2);
        integ.setU0(a.u0);
        integ.setW0coeff(a.w0coeff);
        integ.setNoneqWorkvalue(a.work_value);
    }
    integ.setBindE(sc.bind_e);
    integ.setPotEnergy(sc.pot_energy);
}

B200IntegrateLangevinStepSDMKernel::~B200IntegrateLangevinStepSDMKernel() {
    if (ctx) sdm_destroy(ctx);
    void* bufs[] = {d_posq, d_force, d_displ, d_saveF1, d_saveX1, d_saveF2, d_velm, d_delta, d_random};
    for (void* p : bufs)
        if (p) cudaFree(p);
}

void B200IntegrateLangevinStepSDMKernel::initialize(const System& system, const LangevinIntegratorSDM& integrator) {
    n = system.getNumParticles();
    masses.resize(n);
    for (int i = 0; i < n; i++) masses[i] = system.getParticleMass(i);
    // the displacement map is snapshotted here, like ReferenceSDMKernels.cpp:150-154
    std::vector<double> displ(3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        const Vec3 d = integrator.getDisplacement(i);
        displ[3 * (size_t)i] = d[0]; displ[3 * (size_t)i + 1] = d[1]; displ[3 * (size_t)i + 2] = d[2];
    }
    SimTKOpenMMUtilities::setRandomNumberSeed((unsigned int)integrator.getRandomNumberSeed());

    const B200NonbondedForce* nb = nullptr;
    for (int i = 0; i < system.getNumForces(); i++)
        if (const B200NonbondedForce* f = dynamic_cast<const B200NonbondedForce*>(&system.getForce(i))) nb = f;

    if (nb) {
        // ---- level A: the fused path owns force group 2 ---------------------------------------------
        if (nb->getNumParticles() != n) throw OpenMMException("B200NonbondedForce must have one entry per particle");
        std::vector<double> q(n), sig(n), eps(n);
        for (int i = 0; i < n; i++) nb->getParticleParameters(i, q[i], sig[i], eps[i]);
        // every exception is an exclusion of the pair sum; those with parameters are also 1-4 terms
        std::vector<int32_t> excl, exc;
        std::vector<double> excp;
        for (int k = 0; k < nb->getNumExceptions(); k++) {
            int p1, p2;
            double qq, s, e;
            nb->getExceptionParameters(k, p1, p2, qq, s, e);
            excl.push_back(p1); excl.push_back(p2);
            if (qq != 0.0 || e != 0.0) {
                exc.push_back(p1); exc.push_back(p2);
                excp.push_back(qq); excp.push_back(s); excp.push_back(e);
            }
        }
        sdm_system s;
        std::memset(&s, 0, sizeof(s));
        s.n_atoms = n;
        s.method = (int)nb->getNonbondedMethod();
        s.cutoff = nb->getCutoffDistance();
        s.eps_rf = nb->getReactionFieldDielectric();
        for (int d = 0; d < 3; d++) s.box[d] = nb->getPeriodicBox()[d];
        s.use_dispersion_correction = nb->getUseDispersionCorrection() ? 1 : 0;
        s.n_exclusions = (int)excl.size() / 2;
        s.n_exceptions = (int)exc.size() / 2;
        s.n_replicas = 1;
        s.charge = q.data(); s.sigma = sig.data(); s.epsilon = eps.data();
        s.exclusions = excl.data(); s.exceptions = exc.data(); s.exception_params = excp.data();
        s.displacement = displ.data();
        s.ewald_tolerance = nb->getEwaldErrorTolerance();   // Ewald / PME: alpha by OpenMM's rule
        s.lj_combining = nb->getCombiningRule() == B200NonbondedForce::Geometric ? SDM_LJ_GEOMETRIC : SDM_LJ_LORENTZ_BERTHELOT;
        sdm_options opt;
        sdm_default_options(&opt);
        check(sdm_create(&s, &opt, &ctx), "sdm_create");
        // NonbondedForce::Ewald / ::PME: the complete sum, like OpenMM's own NonbondedForce kernel -- direct space
        // in the pair kernels, reciprocal space (smooth PME, both states) on the device as well
        if (s.method == SDM_EWALD || s.method == SDM_PME)
            check(sdm_enable_reciprocal_pme(ctx, nullptr), "sdm_enable_reciprocal_pme");
        // implicit solvent in the nonbonded group (GBSAHCTForce): both states on the device as well
        if (nb->getNumGBParticles() > 0) {
            if (nb->getNumGBParticles() != n) throw OpenMMException("B200NonbondedForce: one GB particle per particle");
            std::vector<double> gq(n), go(n), gs(n);
            for (int i = 0; i < n; i++) {
                const B200NonbondedForce::GBParticle& g = nb->getGBParticle(i);
                gq[i] = g.charge; go[i] = g.offsetRadius; gs[i] = g.scaledRadius;
            }
            check(sdm_enable_hct_gb(ctx, gq.data(), go.data(), gs.data(), nb->getGBSoluteDielectric(),
                                    nb->getGBSolventDielectric(), nb->getGBSurfaceAreaACE() ? 1 : 0), "sdm_enable_hct_gb");
        }
    } else {
        // ---- level B: OpenMM evaluates force group 2; the plugin's own kernels run on the device -----
        const size_t bytes = sizeof(float) * 4 * (size_t)n;
        void** bufs[] = {&d_posq, &d_force, &d_displ, &d_saveF1, &d_saveX1, &d_saveF2, &d_velm, &d_delta, &d_random};
        for (void** p : bufs) {
            checkCuda(cudaMalloc(p, bytes), "cudaMalloc");
            checkCuda(cudaMemset(*p, 0, bytes), "cudaMemset");
        }
        h4.assign(4 * (size_t)n, 0.f);
        for (int i = 0; i < n; i++)
            for (int d = 0; d < 3; d++) h4[4 * (size_t)i + d] = (float)displ[3 * (size_t)i + d];
        checkCuda(cudaMemcpy(d_displ, h4.data(), bytes, cudaMemcpyHostToDevice), "upload displacement");
        state1Positions.resize(n);
    }
}

// ---- the four state operations --------------------------------------------------------------------
// Level A: nothing to do -- the fused evaluation reads state 1 and forms state 2 for the displaced
// atoms only, inside execute().  Level B: the literal single-precision device operations.
static void upload4(void* dst, const std::vector<Vec3>& v, std::vector<float>& h4, const float* w = nullptr) {
    for (size_t i = 0; i < v.size(); i++) {
        h4[4 * i] = (float)v[i][0]; h4[4 * i + 1] = (float)v[i][1]; h4[4 * i + 2] = (float)v[i][2];
        h4[4 * i + 3] = w ? w[i] : 0.f;
    }
    checkCuda(cudaMemcpy(dst, h4.data(), sizeof(float) * h4.size(), cudaMemcpyHostToDevice), "upload");
}
static void download4(const void* src, std::vector<Vec3>& v, std::vector<float>& h4) {
    checkCuda(cudaMemcpy(h4.data(), src, sizeof(float) * h4.size(), cudaMemcpyDeviceToHost), "download");
    for (size_t i = 0; i < v.size(); i++) v[i] = Vec3(h4[4 * i], h4[4 * i + 1], h4[4 * i + 2]);
}

void B200IntegrateLangevinStepSDMKernel::SaveState1(ContextImpl& context, const LangevinIntegratorSDM&) {
    if (isFused()) return;
    state1Positions = extractPositions(context);            // the Context's doubles survive the float round trip
    upload4(d_posq, extractPositions(context), h4);
    upload4(d_force, extractForces(context), h4);
    check(sdm_k_save_state1(nullptr, n, d_posq, d_force, d_saveF1, d_saveX1), "sdm_k_save_state1");
}

void B200IntegrateLangevinStepSDMKernel::MakeState2(ContextImpl& context, const LangevinIntegratorSDM&) {
    if (isFused()) return;
    check(sdm_k_make_state2(nullptr, n, d_posq, d_displ), "sdm_k_make_state2");
    download4(d_posq, extractPositions(context), h4);        // OpenMM evaluates state 2 at these coordinates
}

void B200IntegrateLangevinStepSDMKernel::SaveState2(ContextImpl& context, const LangevinIntegratorSDM&) {
    if (isFused()) return;
    upload4(d_force, extractForces(context), h4);
    check(sdm_k_save_state2(nullptr, n, d_force, d_saveF2), "sdm_k_save_state2");
}

void B200IntegrateLangevinStepSDMKernel::RestoreState1(ContextImpl& context, const LangevinIntegratorSDM&) {
    if (isFused()) return;
    check(sdm_k_restore_state1(nullptr, n, d_posq, d_saveX1), "sdm_k_restore_state1");
    extractPositions(context) = state1Positions;
}

// The normals of one step in the order ReferenceStochasticDynamicsSDM::updatePart1 draws them
// (:156-162: atoms in order, three per atom, massless atoms draw nothing), from OpenMM's generator.
void B200IntegrateLangevinStepSDMKernel::drawNoise(std::vector<double>& xi) const {
    xi.assign(3 * (size_t)n, 0.0);
    for (int i = 0; i < n; i++)
        if (masses[i] != 0.0)
            for (int d = 0; d < 3; d++) xi[3 * (size_t)i + d] = SimTKOpenMMUtilities::getNormallyDistributedRandomNumber();
}

void B200IntegrateLangevinStepSDMKernel::execute(ContextImpl& context, LangevinIntegratorSDM& integrator,
                                                 double State1Energy, double State2Energy, double RestraintEnergy) {
    if (isFused()) executeFused(context, integrator, RestraintEnergy);
    else executeLiteral(context, integrator, State1Energy, State2Energy, RestraintEnergy);
    data.time += integrator.getStepSize();    // ReferenceSDMKernels.cpp:341-342
    data.stepCount++;
}

// Level A.  State1Energy / State2Energy of the caller are those of the no-op B200NonbondedForce
// (zero); the force buffer holds the bonded + restraint forces of group 1 (the third
// calcForcesAndEnergy of step(), LangevinIntegratorSDM.cpp:176).
void B200IntegrateLangevinStepSDMKernel::executeFused(ContextImpl& context, LangevinIntegratorSDM& integrator,
                                                      double restraintEnergy) {
    std::vector<Vec3>& pos = extractPositions(context);
    std::vector<Vec3>& vel = extractVelocities(context);
    std::vector<Vec3>& frc = extractForces(context);
    std::vector<double> x, v, fb, xi;
    flatten(pos, x);
    flatten(vel, v);
    flatten(frc, fb);
    check(sdm_set_positions(ctx, 0, x.data()), "sdm_set_positions");
    check(sdm_set_bonded_forces(ctx, 0, fb.data(), restraintEnergy), "sdm_set_bonded_forces");
    const sdm_alch a0 = alchFromIntegrator(integrator, data.time);
    sdm_scalars sc;
    for (int attempt = 0; attempt < 4; attempt++) {
        // SDM_ERR_STALE_LIST / SDM_ERR_CAPACITY heal themselves (sdmb200.h): the evaluation is repeated
        check(sdm_set_alchemical(ctx, 0, &a0), "sdm_set_alchemical");
        check(sdm_eval(ctx), "sdm_eval");
        check(sdm_get_scalars(ctx, 0, &sc), "sdm_get_scalars");
        if (sc.status != SDM_ERR_STALE_LIST && sc.status != SDM_ERR_CAPACITY) break;
    }
    if (sc.status == SDM_ERR_SOFTCORE) throw OpenMMException("Unknown soft core method");   // LangevinIntegratorSDM.cpp:147
    if (sc.status != SDM_OK) throw OpenMMException(("libsdmb200 status " + std::to_string(sc.status)).c_str());
    sdm_alch a1;
    check(sdm_get_alchemical(ctx, 0, &a1), "sdm_get_alchemical");
    alchToIntegrator(a1, sc, integrator);

    // the Langevin update of ReferenceStochasticDynamicsSDM::update on the device, with the System's
    // constraints between its two halves; the dynamics object is recreated when its parameters change
    // (ReferenceSDMKernels.cpp:320-335), velocities survive
    const double T = integrator.getTemperature(), g = integrator.getFriction(), dt = integrator.getStepSize();
    if (T != mdTemp || g != mdFriction || dt != mdStep) {
        check(sdm_md_init(ctx, masses.data(), T, g, dt, (uint64_t)integrator.getRandomNumberSeed()), "sdm_md_init");
        const System& system = context.getSystem();
        std::vector<int32_t> cp;
        std::vector<double> cd;
        for (int k = 0; k < system.getNumConstraints(); k++) {
            int p1, p2;
            double d;
            system.getConstraintParameters(k, p1, p2, d);
            cp.push_back(p1); cp.push_back(p2); cd.push_back(d);
        }
        check(sdm_md_set_constraints(ctx, (int32_t)cd.size(), cp.data(), cd.data(), integrator.getConstraintTolerance()),
              "sdm_md_set_constraints");
        mdTemp = T; mdFriction = g; mdStep = dt;
    }
    drawNoise(xi);
    check(sdm_md_set_velocities(ctx, 0, v.data()), "sdm_md_set_velocities");
    check(sdm_md_set_noise(ctx, xi.data()), "sdm_md_set_noise");
    check(sdm_md_update(ctx, nullptr), "sdm_md_update");        // integrates the hybrid force that is on the device
    check(sdm_get_positions(ctx, 0, x.data()), "sdm_get_positions");
    check(sdm_md_get_velocities(ctx, 0, v.data()), "sdm_md_get_velocities");
    check(sdm_get_forces(ctx, 0, SDM_FORCE_HYBRID, fb.data()), "sdm_get_forces");
    unflatten(x, pos);
    unflatten(v, vel);
    unflatten(fb, frc);                                          // forceData ends as the hybrid force (:309-318)
}

// Level B: scalar half on the host (sdm_execute_scalars = LangevinIntegratorSDM.cpp:125-149 +
// ReferenceSDMKernels.cpp:202-305), force mix and both integration kernels in single precision on the
// device like platforms/opencl/src/OpenCLSDMKernels.cpp:273-275,331-372.
void B200IntegrateLangevinStepSDMKernel::executeLiteral(ContextImpl& context, LangevinIntegratorSDM& integrator,
                                                        double e1, double e2, double eb) {
    sdm_alch a = alchFromIntegrator(integrator, data.time);
    sdm_scalars sc;
    const int rc = sdm_execute_scalars(&a, e1, e2, eb, &sc);
    if (rc == SDM_ERR_SOFTCORE) throw OpenMMException("Unknown soft core method");
    check(rc, "sdm_execute_scalars");
    alchToIntegrator(a, sc, integrator);

    std::vector<Vec3>& pos = extractPositions(context);
    std::vector<Vec3>& vel = extractVelocities(context);
    std::vector<Vec3>& frc = extractForces(context);
    upload4(d_force, frc, h4);                                   // bonded + restraint forces
    check(sdm_k_hybrid_force(nullptr, n, d_saveF1, d_saveF2, d_force, (float)sc.sp), "sdm_k_hybrid_force");
    std::vector<float> invm(n);
    for (int i = 0; i < n; i++) invm[i] = masses[i] == 0.0 ? 0.f : (float)(1.0 / masses[i]);
    upload4(d_velm, vel, h4, invm.data());
    upload4(d_posq, pos, h4);
    std::vector<double> xi;
    drawNoise(xi);
    for (int i = 0; i < n; i++) {
        h4[4 * (size_t)i] = (float)xi[3 * (size_t)i]; h4[4 * (size_t)i + 1] = (float)xi[3 * (size_t)i + 1];
        h4[4 * (size_t)i + 2] = (float)xi[3 * (size_t)i + 2]; h4[4 * (size_t)i + 3] = 0.f;
    }
    checkCuda(cudaMemcpy(d_random, h4.data(), sizeof(float) * h4.size(), cudaMemcpyHostToDevice), "upload noise");
    double vscale, fscale, noisescale;
    check(sdm_langevin_params(integrator.getTemperature(), integrator.getFriction(), integrator.getStepSize(), &vscale,
                              &fscale, &noisescale), "sdm_langevin_params");
    const float dt = (float)integrator.getStepSize();
    check(sdm_k_langevin_part1(nullptr, n, d_velm, d_force, d_delta, (float)vscale, (float)fscale, (float)noisescale, dt,
                               d_random, 0), "sdm_k_langevin_part1");
    // (OpenMM's constraint kernels run here on its own platforms, OpenCLSDMKernels.cpp:357-372)
    check(sdm_k_langevin_part2(nullptr, n, d_posq, d_delta, d_velm, dt), "sdm_k_langevin_part2");
    checkCuda(cudaDeviceSynchronize(), "synchronize");
    // the step in single precision is applied to the Context's double-precision coordinates
    std::vector<Vec3> delta(n), vnew(n), fh(n);
    download4(d_delta, delta, h4);
    download4(d_velm, vnew, h4);
    download4(d_force, fh, h4);
    for (int i = 0; i < n; i++) {
        if (masses[i] == 0.0) continue;
        pos[i] = pos[i] + delta[i];
        vel[i] = vnew[i];
    }
    frc = fh;
}

double B200IntegrateLangevinStepSDMKernel::computeKineticEnergy(ContextImpl& context, const LangevinIntegratorSDM&) {
    // ReferenceSDMKernels.cpp:105-137 (the time shift is ignored there too)
    const std::vector<Vec3>& vel = extractVelocities(context);
    double e = 0.0;
    for (int i = 0; i < n; i++)
        if (masses[i] > 0) e += masses[i] * vel[i].dot(vel[i]);
    return 0.5 * e;
}

#ifndef SDMB200_OPENMM_STUB
// A full OpenMM wants a ForceImpl behind every Force: this one contributes nothing (the evaluation of
// force group 2 happens in execute()).
}  // namespace SDMB200
#include "openmm/internal/ForceImpl.h"
namespace SDMB200 {
class B200NonbondedForceImpl : public OpenMM::ForceImpl {
public:
    explicit B200NonbondedForceImpl(const B200NonbondedForce& owner) : owner(owner) {}
    void initialize(ContextImpl&) {}
    const B200NonbondedForce& getOwner() const { return owner; }
    void updateContextState(ContextImpl&) {}
    double calcForcesAndEnergy(ContextImpl&, bool, bool, int) { return 0.0; }
    std::map<std::string, double> getDefaultParameters() { return std::map<std::string, double>(); }
    std::vector<std::string> getKernelNames() { return std::vector<std::string>(); }
private:
    const B200NonbondedForce& owner;
};
OpenMM::ForceImpl* B200NonbondedForce::createImpl() const { return new B200NonbondedForceImpl(*this); }
#endif

}  // namespace SDMB200
