// LangevinIntegratorSDM.h -- C++ host-side mirror of SDMPlugin::LangevinIntegratorSDM for the
// dual-state force path (header only, no OpenMM dependency; talks to libsdmb200 through the C
// ABI of include/sdmb200.h).
//
// Same class and method names, argument meaning, defaults and units (kJ/mol, nm, ps, K) as
// openmmapi/include/LangevinIntegratorSDM.h:57-520 and the constructor at
// openmmapi/src/LangevinIntegratorSDM.cpp:48-85.  What the reference's step() does around the
// force column -- integrating the Langevin equations, constraints -- is outside the hot path
// (SURVEY.md section 8(f) N2): evaluate() performs exactly LangevinIntegratorSDM.cpp:156-182 up
// to and including the hybrid force of ReferenceSDMKernels.cpp:309-318, on the GPU.
// Errors surface as SDMPlugin::SDMException with the reference's messages where one exists
// ("Unknown soft core method", LangevinIntegratorSDM.cpp:147; "This Integrator is already bound
// to a context", :90).
#pragma once

#include <array>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/sdmb200.h"

namespace SDMPlugin {

class SDMException : public std::runtime_error {
public:
    explicit SDMException(const std::string& what) : std::runtime_error(what) {}
};

class LangevinIntegratorSDM {
public:
    // LangevinIntegratorSDM.h:120-122 and :143-145
    static const int LinearMethod = 0;
    static const int QuadraticMethod = 1;
    static const int ILogisticMethod = 2;
    static const int NoSoftCoreMethod = 0;
    static const int TanhMethod = 1;
    static const int RationalMethod = 2;

    LangevinIntegratorSDM(double temperature, double frictionCoeff, double stepSize, int nParticles_t)
        : temperature(temperature), friction(frictionCoeff), stepSize(stepSize), randomNumberSeed(0),
          BindE(0.0), PotEnergy(0.0), nParticles(nParticles_t), displ((size_t)nParticles_t, {0.0, 0.0, 0.0}),
          ctx(nullptr), displDirty(false) {
        sdm_default_alch(&alch);  // umax 200, a 1/4, ub 0, no soft core, linear bias, lambda 1, ...
        alch.step_size = stepSize;
    }
    ~LangevinIntegratorSDM() { cleanup(); }
    LangevinIntegratorSDM(const LangevinIntegratorSDM&) = delete;
    LangevinIntegratorSDM& operator=(const LangevinIntegratorSDM&) = delete;

    double getTemperature() const { return temperature; }
    void setTemperature(double temp) { temperature = temp; }
    double getFriction() const { return friction; }
    void setFriction(double coeff) { friction = coeff; }
    double getStepSize() const { return stepSize; }
    void setStepSize(double size) { stepSize = size; alch.step_size = size; }
    int getRandomNumberSeed() const { return randomNumberSeed; }
    void setRandomNumberSeed(int seed) { randomNumberSeed = seed; }

    void setLambda(double lambdac) { alch.lambdac = lambdac; }
    double getLambda() const { return alch.lambdac; }
    double getBindE() const { return BindE; }
    void setBindE(double be) { BindE = be; }
    double getPotEnergy() const { return PotEnergy; }
    void setPotEnergy(double e) { PotEnergy = e; }
    double getUmax() const { return alch.umax; }
    void setUmax(double um) { alch.umax = um; }
    double getAcore() const { return alch.acore; }
    void setAcore(double a) { alch.acore = a; }
    double getUbcore() const { return alch.ubcore; }
    void setUbcore(double a) { alch.ubcore = a; }
    void setBiasMethod(int method) { alch.bias_method = method; }
    int getBiasMethod() const { return alch.bias_method; }
    void setSoftCoreMethod(int method) { alch.softcore_method = method; }
    int getSoftCoreMethod() const { return alch.softcore_method; }
    void setGamma(double gammat) { alch.gammac = gammat; }
    double getGamma() const { return alch.gammac; }
    void setWBcoeff(double wbcoeff_t) { alch.wbcoeff = wbcoeff_t; }
    double getWBcoeff() const { return alch.wbcoeff; }
    void setW0coeff(double w0coeff_t) { alch.w0coeff = w0coeff_t; }
    double getW0coeff() const { return alch.w0coeff; }
    void setLambda1(double lambda1_t) { alch.lambda1 = lambda1_t; }
    double getLambda1() const { return alch.lambda1; }
    void setLambda2(double lambda2_t) { alch.lambda2 = lambda2_t; }
    double getLambda2() const { return alch.lambda2; }
    void setAlpha(double alpha_t) { alch.alpha = alpha_t; }
    double getAlpha() const { return alch.alpha; }
    void setU0(double u0_t) { alch.u0 = u0_t; }
    double getU0() const { return alch.u0; }
    void setNonEquilibrium(int flag) { alch.nonequilibrium = flag; }
    int getNonEquilibrium() const { return alch.nonequilibrium; }
    void setNoneqtmax(double noneq_tmax) { alch.noneq_tmax = noneq_tmax; }
    double getNoneqtmax() const { return alch.noneq_tmax; }
    void setNoneqWorkvalue(double noneq_work) { alch.work_value = noneq_work; }
    double getNoneqWorkvalue() const { return alch.work_value; }
    void setlambda1Slope(double ml1) { alch.m_lambda1 = ml1; }
    double getlambda1Slope() const { return alch.m_lambda1; }
    void setlambda2Slope(double ml2) { alch.m_lambda2 = ml2; }
    double getlambda2Slope() const { return alch.m_lambda2; }
    void setu0Slope(double mu0) { alch.m_u0 = mu0; }
    double getu0Slope() const { return alch.m_u0; }
    void setw0Slope(double mw0) { alch.m_w0 = mw0; }
    double getw0Slope() const { return alch.m_w0; }
    void setlambda1intercept(double bl1) { alch.b_lambda1 = bl1; }
    double getlambda1intercept() const { return alch.b_lambda1; }
    void setlambda2intercept(double bl2) { alch.b_lambda2 = bl2; }
    double getlambda2intercept() const { return alch.b_lambda2; }
    void setu0intercept(double bu0) { alch.b_u0 = bu0; }
    double getu0intercept() const { return alch.b_u0; }
    void setw0intercept(double bw0) { alch.b_w0 = bw0; }
    double getw0intercept() const { return alch.b_w0; }

    // LangevinIntegratorSDM.h:467-472 (the reference does not check the index; this does)
    void setDisplacement(int atom, double dx, double dy, double dz) {
        checkAtom(atom);
        displ[(size_t)atom] = {dx, dy, dz};
        displDirty = true;
    }
    std::array<double, 3> getDisplacement(int atom) const {
        checkAtom(atom);
        return displ[(size_t)atom];
    }

    // What initialize() does for this path (LangevinIntegratorSDM.cpp:89-106 and the snapshot
    // of the displacement map, ReferenceSDMKernels.cpp:150-154).  `system.displacement` and
    // `system.n_replicas` are filled in here.
    void bind(sdm_system system, const sdm_options* options = nullptr) {
        if (ctx) throw SDMException("This Integrator is already bound to a context");
        if (system.n_atoms != nParticles) throw SDMException("nParticles does not match the system");
        std::vector<double> flat = flatDisplacement();
        system.displacement = flat.data();
        system.n_replicas = 1;
        if (sdm_create(&system, options, &ctx) != SDM_OK) throw SDMException(sdm_last_error());
        // NonbondedForce::Ewald / ::PME: the reference gets the complete sum from OpenMM -- direct space is in the
        // pair kernels, the reciprocal part of both states is switched on here
        if ((system.method == SDM_EWALD || system.method == SDM_PME) && sdm_enable_reciprocal_pme(ctx, nullptr) != SDM_OK) {
            const std::string msg = sdm_last_error();
            cleanup();
            throw SDMException(msg);
        }
        displDirty = false;
    }
    // The GBSAHCTForce the reference's reader adds to the nonbonded force group for implicitSolvent=HCT
    // (example/desmonddmsfile75.py:454-465): evaluated in both states from now on.  Per-particle parameters as the
    // CustomGBForce holds them (charge may be null: the NonbondedForce charges; or = radius - 0.009, sr = scale * or).
    void addImplicitSolventHCT(const double* charge, const double* offsetRadius, const double* scaledRadius,
                               double soluteDielectric = 1.0, double solventDielectric = 78.5, bool surfaceAreaACE = true) {
        if (!ctx) throw SDMException("the integrator is not bound to a context: call bind(system) first");
        if (sdm_enable_hct_gb(ctx, charge, offsetRadius, scaledRadius, soluteDielectric, solventDielectric,
                              surfaceAreaACE ? 1 : 0) != SDM_OK)
            throw SDMException(sdm_last_error());
    }
    void cleanup() {
        if (ctx) sdm_destroy(ctx);
        ctx = nullptr;
    }

    // The force column of one step(): positions [3*nParticles] (nm) in, hybrid force
    // [3*nParticles] (kJ/mol/nm) out; bondedForces may be null.
    void evaluate(const double* positions, const double* bondedForces, double restraintEnergy,
                  double* hybridForce) {
        if (!ctx) throw SDMException("the integrator is not bound to a context");
        if (displDirty) {
            std::vector<double> flat = flatDisplacement();
            check(sdm_set_displacement(ctx, flat.data()));
            displDirty = false;
        }
        check(sdm_set_positions(ctx, 0, positions));
        check(sdm_set_bonded_forces(ctx, 0, bondedForces, restraintEnergy));
        sdm_scalars sc;
        for (int attempt = 0; attempt < 4; attempt++) {
            // SDM_ERR_STALE_LIST / SDM_ERR_CAPACITY heal themselves (sdmb200.h): reading the scalars
            // made the library rebuild its list / grow its scratch; the evaluation is repeated from
            // the same alchemical state.  The reference's force path never fails this way.
            check(sdm_set_alchemical(ctx, 0, &alch));
            check(sdm_eval(ctx));
            check(sdm_get_scalars(ctx, 0, &sc));
            if (sc.status != SDM_ERR_STALE_LIST && sc.status != SDM_ERR_CAPACITY) break;
        }
        if (sc.status == SDM_ERR_SOFTCORE) throw SDMException("Unknown soft core method");
        if (sc.status != SDM_OK) throw SDMException("libsdmb200 status " + std::to_string(sc.status));
        BindE = sc.bind_e;
        PotEnergy = sc.pot_energy;
        lastScalars = sc;
        check(sdm_get_alchemical(ctx, 0, &alch));  // non-equilibrium schedule written back
        check(sdm_get_forces(ctx, 0, SDM_FORCE_HYBRID, hybridForce));
    }
    const sdm_scalars& getLastScalars() const { return lastScalars; }

    // ---- dynamics on the device (no constraints) ------------------------
    // What the reference's Context holds for the integrator: positions [3*nParticles] (nm),
    // velocities (nm/ps; null = zero) and the particle masses (amu; needed once, null = keep).
    void setState(const double* positions, const double* velocities, const double* particleMasses) {
        if (!ctx) throw SDMException("the integrator is not bound to a context");
        if (particleMasses) {
            masses.assign(particleMasses, particleMasses + nParticles);
            mdTemperature = -1;   // forces a new dynamics object
        }
        check(sdm_set_positions(ctx, 0, positions));
        check(sdm_synchronize(ctx));
        pendingVel.assign(3 * (size_t)nParticles, 0.0);
        if (velocities) pendingVel.assign(velocities, velocities + 3 * (size_t)nParticles);
        havePendingVel = true;
    }
    void getPositions(double* positions) { check(sdm_get_positions(ctx, 0, positions)); }
    void getVelocities(double* velocities) { check(sdm_md_get_velocities(ctx, 0, velocities)); }
    double computeKineticEnergy() {
        double ke = 0;
        check(sdm_md_kinetic_energy(ctx, 0, &ke));
        return ke;
    }

    // LangevinIntegratorSDM::step (LangevinIntegratorSDM.cpp:153-183) with the state on the device:
    // per step the fused dual-state evaluation and the reference's Langevin update
    // (ReferenceStochasticDynamicsSDM.cpp:131-266, FP64, no constraints).  Force group 1 is whatever
    // evaluate() / sdm_set_bonded_forces last handed over (zero by default).
    void step(int steps) {
        if (!ctx) throw SDMException("the integrator is not bound to a context");
        if (masses.empty())
            throw SDMException("LangevinIntegratorSDM::step needs the particle masses: call setState() first, "
                               "or evaluate() for the force column of a step alone");
        if (mdTemperature != temperature || mdFriction != friction || mdStepSize != stepSize) {
            // like the Reference kernel, the dynamics object is recreated when T, friction or dt
            // change (ReferenceSDMKernels.cpp:320-337); velocities survive
            std::vector<double> keep;
            if (mdTemperature >= 0) {
                keep.resize(3 * (size_t)nParticles);
                check(sdm_md_get_velocities(ctx, 0, keep.data()));
            }
            check(sdm_md_init(ctx, masses.data(), temperature, friction, stepSize, (uint64_t)randomNumberSeed));
            if (!keep.empty()) check(sdm_md_set_velocities(ctx, 0, keep.data()));
            mdTemperature = temperature; mdFriction = friction; mdStepSize = stepSize;
        }
        if (havePendingVel) {
            check(sdm_md_set_velocities(ctx, 0, pendingVel.data()));
            havePendingVel = false;
        }
        if (displDirty) {
            std::vector<double> flat = flatDisplacement();
            check(sdm_set_displacement(ctx, flat.data()));
            displDirty = false;
        }
        for (int k = 0; k < steps; k++) {
            check(sdm_set_alchemical(ctx, 0, &alch));
            check(sdm_md_step(ctx, 1));
            sdm_scalars sc;
            check(sdm_get_scalars(ctx, 0, &sc));
            if (sc.status == SDM_ERR_SOFTCORE) throw SDMException("Unknown soft core method");
            if (sc.status == SDM_ERR_STALE_LIST) check(sdm_invalidate_list(ctx));   // rebuild before the next step
            else if (sc.status != SDM_OK) throw SDMException("libsdmb200 status " + std::to_string(sc.status));
            BindE = sc.bind_e;
            PotEnergy = sc.pot_energy;
            lastScalars = sc;
            check(sdm_get_alchemical(ctx, 0, &alch));
        }
    }

private:
    void checkAtom(int atom) const {
        if (atom < 0 || atom >= nParticles) throw SDMException("particle index out of range");
    }
    static void check(int rc) {
        if (rc != SDM_OK) throw SDMException(sdm_last_error());
    }
    std::vector<double> flatDisplacement() const {
        std::vector<double> flat(3 * (size_t)nParticles);
        for (int i = 0; i < nParticles; i++)
            for (int d = 0; d < 3; d++) flat[3 * (size_t)i + d] = displ[(size_t)i][(size_t)d];
        return flat;
    }

    double temperature, friction, stepSize;
    int randomNumberSeed;
    double BindE, PotEnergy;
    int nParticles;
    std::vector<std::array<double, 3>> displ;
    sdm_alch alch;
    sdm_scalars lastScalars{};
    sdm_ctx* ctx;
    bool displDirty;
    std::vector<double> masses, pendingVel;
    bool havePendingVel = false;
    double mdTemperature = -1, mdFriction = -1, mdStepSize = -1;
};

}  // namespace SDMPlugin
