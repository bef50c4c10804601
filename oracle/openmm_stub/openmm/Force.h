#pragma once
namespace OpenMM {
class Force {
public:
    explicit Force(int group = 0) : group(group) {}
    virtual ~Force() {}
    int getForceGroup() const { return group; }
    void setForceGroup(int g) { group = g; }
private:
    int group;
};
}  // namespace OpenMM
