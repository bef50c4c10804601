// B200NonbondedForce.h -- the description of force group 2 (NonbondedForce as the reference's reader
// builds it, example/desmonddmsfile75.py:772-850) handed to the B200 kernel through the OpenMM
// System.
//
// In the reference, LangevinIntegratorSDM::step (openmmapi/src/LangevinIntegratorSDM.cpp:153-183)
// asks OpenMM twice per step for the energy and forces of force group 2.  With the fused path that
// work moves into IntegrateLangevinStepSDMKernel::execute on the B200, so the System carries this
// Force INSTEAD of OpenMM's NonbondedForce: same particle / exception parameters, same setters, but
// its OpenMM-side implementation evaluates to zero (the two calcForcesAndEnergy(..., 4) calls of the
// unmodified step() cost nothing) and B200IntegrateLangevinStepSDMKernel::initialize reads the
// parameters from it.  It sits in force group 2, which is what LangevinIntegratorSDM::initialize
// (:92-100) requires of a nonbonded force.
#pragma once

#include <vector>

#include "openmm/Force.h"

namespace SDMB200 {

class B200NonbondedForce : public OpenMM::Force {
public:
    // values of OpenMM::NonbondedForce::NonbondedMethod that the path supports
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2, Ewald = 3, PME = 4 };
    struct Exception { int p1, p2; double chargeProd, sigma, epsilon; };
    // Lorentz-Berthelot is NonbondedForce's rule; Geometric stands for the NonbondedForce (epsilons zeroed) +
    // CustomNonbondedForce pair that createSystem(OPLS=True) builds (example/desmonddmsfile75.py:780-810)
    enum CombiningRule { LorentzBerthelot = 0, Geometric = 1 };
    // GBSAHCTForce(SA='ACE') of desmonddmsfile75.py:460 as its CustomGBForce holds it: charge, or, sr per particle
    struct GBParticle { double charge, offsetRadius, scaledRadius; };

    B200NonbondedForce() : method(NoCutoff), cutoff(1.0), rfDielectric(78.3), ewaldTol(5e-4), dispersion(true),
                           rule(LorentzBerthelot), gbSolute(1.0), gbSolvent(78.5), gbAce(true) {
        box[0] = box[1] = box[2] = 0.0;
        setForceGroup(2);
    }
    int getNumParticles() const { return (int)charge.size(); }
    int addParticle(double q, double sig, double eps) {
        charge.push_back(q); sigma.push_back(sig); epsilon.push_back(eps);
        return (int)charge.size() - 1;
    }
    void getParticleParameters(int i, double& q, double& sig, double& eps) const { q = charge[i]; sig = sigma[i]; eps = epsilon[i]; }
    int getNumExceptions() const { return (int)exceptions.size(); }
    int addException(int p1, int p2, double chargeProd, double sig, double eps) {
        exceptions.push_back(Exception{p1, p2, chargeProd, sig, eps});
        return (int)exceptions.size() - 1;
    }
    void getExceptionParameters(int i, int& p1, int& p2, double& chargeProd, double& sig, double& eps) const {
        const Exception& e = exceptions[i];
        p1 = e.p1; p2 = e.p2; chargeProd = e.chargeProd; sig = e.sigma; eps = e.epsilon;
    }
    NonbondedMethod getNonbondedMethod() const { return method; }
    void setNonbondedMethod(NonbondedMethod m) { method = m; }
    double getCutoffDistance() const { return cutoff; }
    void setCutoffDistance(double d) { cutoff = d; }
    double getReactionFieldDielectric() const { return rfDielectric; }
    void setReactionFieldDielectric(double d) { rfDielectric = d; }
    double getEwaldErrorTolerance() const { return ewaldTol; }
    void setEwaldErrorTolerance(double t) { ewaldTol = t; }
    bool getUseDispersionCorrection() const { return dispersion; }
    void setUseDispersionCorrection(bool b) { dispersion = b; }
    CombiningRule getCombiningRule() const { return rule; }
    void setCombiningRule(CombiningRule r) { rule = r; }
    int getNumGBParticles() const { return (int)gb.size(); }
    int addGBParticle(double q, double offsetRadius, double scaledRadius) {
        gb.push_back(GBParticle{q, offsetRadius, scaledRadius});
        return (int)gb.size() - 1;
    }
    const GBParticle& getGBParticle(int i) const { return gb[i]; }
    void setGBDielectrics(double solute, double solvent) { gbSolute = solute; gbSolvent = solvent; }
    double getGBSoluteDielectric() const { return gbSolute; }
    double getGBSolventDielectric() const { return gbSolvent; }
    void setGBSurfaceAreaACE(bool on) { gbAce = on; }
    bool getGBSurfaceAreaACE() const { return gbAce; }
    // orthorhombic box edges (nm); System::getDefaultPeriodicBoxVectors in a full OpenMM
    void setPeriodicBox(double a, double b, double c) { box[0] = a; box[1] = b; box[2] = c; }
    const double* getPeriodicBox() const { return box; }
    bool usesPeriodicBoundaryConditions() const { return method == CutoffPeriodic || method == Ewald || method == PME; }

#ifndef SDMB200_OPENMM_STUB
protected:
    // a ForceImpl that contributes nothing: the evaluation happens in the B200 kernel (B200SDMKernels.cpp)
    OpenMM::ForceImpl* createImpl() const;
#endif

private:
    std::vector<double> charge, sigma, epsilon;
    std::vector<Exception> exceptions;
    NonbondedMethod method;
    double cutoff, rfDielectric, ewaldTol, box[3];
    bool dispersion;
    CombiningRule rule;
    std::vector<GBParticle> gb;
    double gbSolute, gbSolvent;
    bool gbAce;
};

}  // namespace SDMB200
