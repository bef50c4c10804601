#pragma once
#include "openmm/KernelImpl.h"
namespace OpenMM {
// Reference-counted handle, like OpenMM's.
class Kernel {
public:
    Kernel() : impl(0) {}
    Kernel(KernelImpl* impl) : impl(impl) { if (impl) impl->refs++; }
    Kernel(const Kernel& o) : impl(o.impl) { if (impl) impl->refs++; }
    ~Kernel() { release(); }
    Kernel& operator=(const Kernel& o) {
        if (o.impl) o.impl->refs++;
        release();
        impl = o.impl;
        return *this;
    }
    template <class T> T& getAs() { return dynamic_cast<T&>(*impl); }
    template <class T> const T& getAs() const { return dynamic_cast<const T&>(*impl); }
private:
    void release() { if (impl && --impl->refs == 0) delete impl; impl = 0; }
    KernelImpl* impl;
};
}  // namespace OpenMM
