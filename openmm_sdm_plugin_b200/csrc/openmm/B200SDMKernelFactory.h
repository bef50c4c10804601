// B200SDMKernelFactory.h -- mirrors platforms/reference/src/ReferenceSDMKernelFactory.h of the reference.
#pragma once

#include "openmm/KernelFactory.h"

namespace SDMB200 {

class B200SDMKernelFactory : public OpenMM::KernelFactory {
public:
    OpenMM::KernelImpl* createKernelImpl(std::string name, const OpenMM::Platform& platform,
                                         OpenMM::ContextImpl& context) const;
};

}  // namespace SDMB200
