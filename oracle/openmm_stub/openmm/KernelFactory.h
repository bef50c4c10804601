#pragma once
#include <string>
namespace OpenMM {
class KernelImpl;
class Platform;
class ContextImpl;
class KernelFactory {
public:
    virtual ~KernelFactory() {}
    virtual KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const = 0;
};
}  // namespace OpenMM
