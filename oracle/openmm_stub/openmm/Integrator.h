#pragma once
#include <string>
#include <vector>
#include "openmm/internal/windowsExport.h"
namespace OpenMM {
class Context;
class ContextImpl;
class Integrator {
public:
    Integrator() : context(0), owner(0), stepSize(0), constraintTol(1e-5) {}
    virtual ~Integrator() {}
    virtual double getStepSize() const { return stepSize; }
    virtual void setStepSize(double size) { stepSize = size; }
    virtual double getConstraintTolerance() const { return constraintTol; }
    virtual void setConstraintTolerance(double tol) { constraintTol = tol; }
    virtual void step(int steps) = 0;
protected:
    friend class ContextImpl;
    ContextImpl* context;
    Context* owner;
    virtual void initialize(ContextImpl& context) = 0;
    virtual void cleanup() {}
    virtual std::vector<std::string> getKernelNames() = 0;
    virtual void stateChanged(int) {}
    virtual double computeKineticEnergy() = 0;
private:
    double stepSize, constraintTol;
};
}  // namespace OpenMM
