// B200SDMKernelFactory.cpp -- the plugin entry points OpenMM calls when it loads the library
// (Platform::loadPluginLibrary): same shape as platforms/reference/src/ReferenceSDMKernelFactory.cpp:41-63
// of the reference, registering the B200 kernel for IntegrateLangevinStepSDMKernel::Name() instead of
// the Reference-platform one.  Load this library INSTEAD of the reference's
// libSDMPluginReference.so; the API library (LangevinIntegratorSDM, SDMplugin python module) stays.
#include "B200SDMKernelFactory.h"

#include "B200SDMKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"

#ifndef OPENMM_EXPORT
#define OPENMM_EXPORT
#endif

using namespace OpenMM;
using namespace SDMPlugin;

extern "C" OPENMM_EXPORT void registerPlatforms() {
}

extern "C" OPENMM_EXPORT void registerKernelFactories() {
    for (int i = 0; i < Platform::getNumPlatforms(); i++) {
        Platform& platform = Platform::getPlatform(i);
        if (dynamic_cast<ReferencePlatform*>(&platform) != NULL)
            platform.registerKernelFactory(IntegrateLangevinStepSDMKernel::Name(), new SDMB200::B200SDMKernelFactory());
    }
}

extern "C" OPENMM_EXPORT void registerSDMB200KernelFactories() {
    registerKernelFactories();
}

KernelImpl* SDMB200::B200SDMKernelFactory::createKernelImpl(std::string name, const Platform& platform,
                                                           ContextImpl& context) const {
    ReferencePlatform::PlatformData& data = *static_cast<ReferencePlatform::PlatformData*>(context.getPlatformData());
    if (name == IntegrateLangevinStepSDMKernel::Name())
        return new SDMB200::B200IntegrateLangevinStepSDMKernel(name, platform, data);
    throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());
}
