// B200SDMKernels.h -- the OpenMM-side adapter: IntegrateLangevinStepSDMKernel implemented on
// libsdmb200 (include/sdmb200.h).  Mirrors platforms/reference/src/ReferenceSDMKernels.h:60-101 of
// the reference: same seven virtuals (openmmapi/include/SDMKernels.h:60-111), driven by the
// reference's UNMODIFIED LangevinIntegratorSDM::step (openmmapi/src/LangevinIntegratorSDM.cpp:153-183).
//
// Two levels, chosen in initialize():
//   A (fused)   -- the System carries a B200NonbondedForce: the four state operations are no-ops
//                  (state 2 is only ever formed for the displaced atoms, on the device) and
//                  execute() runs the whole step on the B200: dual-state evaluation, soft core,
//                  bias, bookkeeping, hybrid force, Langevin update with the System's constraints.
//   B (literal) -- OpenMM keeps its own NonbondedForce: SaveState1 / MakeState2 / SaveState2 /
//                  RestoreState1 / the force mix / the Langevin update are the single-precision
//                  device operations of the reference's OpenCL kernels
//                  (platforms/opencl/src/kernels/langevin.cl) through sdm_k_*.
// The positions, velocities and forces are the Context's (ReferencePlatform::PlatformData, host
// vectors) -- the data interface that is compiled and tested here (oracle/openmm_stub stands in for
// OpenMM's headers; build with -DSDMB200_OPENMM_STUB).  Against OpenMM's CUDA platform only the five
// extract*/store* helpers at the top of B200SDMKernels.cpp change (CudaContext arrays instead of
// host vectors).
#pragma once

#include <vector>

#include "SDMKernels.h"
#include "openmm/reference/ReferencePlatform.h"
#include "sdmb200.h"

namespace SDMB200 {

class B200IntegrateLangevinStepSDMKernel : public SDMPlugin::IntegrateLangevinStepSDMKernel {
public:
    B200IntegrateLangevinStepSDMKernel(std::string name, const OpenMM::Platform& platform,
                                       OpenMM::ReferencePlatform::PlatformData& data)
        : SDMPlugin::IntegrateLangevinStepSDMKernel(name, platform), data(data) {}
    ~B200IntegrateLangevinStepSDMKernel();

    void initialize(const OpenMM::System& system, const SDMPlugin::LangevinIntegratorSDM& integrator);
    void execute(OpenMM::ContextImpl& context, SDMPlugin::LangevinIntegratorSDM& integrator,
                 double State1Energy, double State2Energy, double RestraintEnergy);
    double computeKineticEnergy(OpenMM::ContextImpl& context, const SDMPlugin::LangevinIntegratorSDM& integrator);
    void SaveState1(OpenMM::ContextImpl& context, const SDMPlugin::LangevinIntegratorSDM& integrator);
    void SaveState2(OpenMM::ContextImpl& context, const SDMPlugin::LangevinIntegratorSDM& integrator);
    void RestoreState1(OpenMM::ContextImpl& context, const SDMPlugin::LangevinIntegratorSDM& integrator);
    void MakeState2(OpenMM::ContextImpl& context, const SDMPlugin::LangevinIntegratorSDM& integrator);

    bool isFused() const { return ctx != nullptr; }

private:
    void executeFused(OpenMM::ContextImpl& context, SDMPlugin::LangevinIntegratorSDM& integrator, double restraintEnergy);
    void executeLiteral(OpenMM::ContextImpl& context, SDMPlugin::LangevinIntegratorSDM& integrator, double e1,
                        double e2, double eb);
    void drawNoise(std::vector<double>& xi) const;

    OpenMM::ReferencePlatform::PlatformData& data;
    int n = 0;
    std::vector<double> masses;
    // level A
    sdm_ctx* ctx = nullptr;
    double mdTemp = -1, mdFriction = -1, mdStep = -1;
    // level B: device float4 arrays (cudaMalloc through sdm_device_alloc) and the double-precision
    // state-1 coordinates (the Context's positions are doubles on this data interface)
    void *d_posq = nullptr, *d_force = nullptr, *d_displ = nullptr, *d_saveF1 = nullptr, *d_saveX1 = nullptr,
         *d_saveF2 = nullptr, *d_velm = nullptr, *d_delta = nullptr, *d_random = nullptr;
    std::vector<float> h4;                       // pinned-size staging [4n]
    std::vector<OpenMM::Vec3> state1Positions;
};

}  // namespace SDMB200
