// Minimal stand-in for OpenMM's Vec3 (oracle/openmm_stub: test infrastructure, see README there).
#pragma once
#include <cassert>
namespace OpenMM {
class Vec3 {
public:
    Vec3() { data[0] = data[1] = data[2] = 0.0; }
    Vec3(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
    double operator[](int i) const { return data[i]; }
    double& operator[](int i) { return data[i]; }
    Vec3 operator+(const Vec3& r) const { return Vec3(data[0] + r[0], data[1] + r[1], data[2] + r[2]); }
    Vec3 operator-(const Vec3& r) const { return Vec3(data[0] - r[0], data[1] - r[1], data[2] - r[2]); }
    Vec3 operator*(double s) const { return Vec3(data[0] * s, data[1] * s, data[2] * s); }
    Vec3 operator/(double s) const { return Vec3(data[0] / s, data[1] / s, data[2] / s); }
    Vec3& operator+=(const Vec3& r) { data[0] += r[0]; data[1] += r[1]; data[2] += r[2]; return *this; }
    Vec3& operator-=(const Vec3& r) { data[0] -= r[0]; data[1] -= r[1]; data[2] -= r[2]; return *this; }
    double dot(const Vec3& r) const { return data[0] * r[0] + data[1] * r[1] + data[2] * r[2]; }
private:
    double data[3];
};
}  // namespace OpenMM
