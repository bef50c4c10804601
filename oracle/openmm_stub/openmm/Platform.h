#pragma once
#include <map>
#include <string>
#include <vector>
#include "openmm/Kernel.h"
#include "openmm/KernelFactory.h"
#include "openmm/OpenMMException.h"
namespace OpenMM {
class Platform {
public:
    virtual ~Platform() {}
    void registerKernelFactory(const std::string& name, KernelFactory* factory) { factories[name] = factory; }
    Kernel createKernel(const std::string& name, ContextImpl& context) const {
        std::map<std::string, KernelFactory*>::const_iterator it = factories.find(name);
        if (it == factories.end())
            throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
        return Kernel(it->second->createKernelImpl(name, *this, context));
    }
    static std::vector<Platform*>& registry() { static std::vector<Platform*> r; return r; }
    static void registerPlatform(Platform* p) { registry().push_back(p); }
    static int getNumPlatforms() { return (int)registry().size(); }
    static Platform& getPlatform(int i) { return *registry()[i]; }
private:
    std::map<std::string, KernelFactory*> factories;
};
}  // namespace OpenMM
