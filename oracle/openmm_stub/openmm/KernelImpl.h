#pragma once
#include <string>
namespace OpenMM {
class Platform;
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform), refs(0) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    friend class Kernel;
    std::string name;
    const Platform* platform;
    int refs;
};
}  // namespace OpenMM
