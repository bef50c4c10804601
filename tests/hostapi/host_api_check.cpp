// host_api_check.cpp -- exercises the C++ mirror of SDMPlugin::LangevinIntegratorSDM without a
// GPU: constructor defaults (openmmapi/src/LangevinIntegratorSDM.cpp:48-85), setter/getter
// round trips, the displacement map, and the error behaviour of an unbound integrator and of
// sdm_create without a device.  Exit code 0 = all checks passed.  (test infrastructure)
#include <cstdio>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../openmm_sdm_plugin_b200/csrc/host/LangevinIntegratorSDM.h"

#define CHECK(x) do { if (!(x)) { std::printf("FAILED: %s (line %d)\n", #x, __LINE__); return 1; } } while (0)

int main() {
    using SDMPlugin::LangevinIntegratorSDM;
    LangevinIntegratorSDM g(300.0, 0.5, 0.001, 4);
    CHECK(g.getTemperature() == 300.0 && g.getFriction() == 0.5 && g.getStepSize() == 0.001);
    CHECK(g.getUmax() == 200.0 && g.getAcore() == 0.25 && g.getUbcore() == 0.0);
    CHECK(g.getSoftCoreMethod() == LangevinIntegratorSDM::NoSoftCoreMethod);
    CHECK(g.getBiasMethod() == LangevinIntegratorSDM::LinearMethod);
    CHECK(g.getLambda() == 1.0 && g.getGamma() == 0.0 && g.getWBcoeff() == 1.0 && g.getW0coeff() == 0.0);
    CHECK(g.getLambda1() == 1.0 && g.getLambda2() == 1.0 && g.getAlpha() == 1.0 && g.getU0() == 0.0);
    CHECK(g.getNonEquilibrium() == 0 && g.getNoneqWorkvalue() == 0.0);
    for (int i = 0; i < 4; i++) CHECK(g.getDisplacement(i)[0] == 0.0 && g.getDisplacement(i)[2] == 0.0);
    g.setDisplacement(1, -1.5559, -0.3, 0.86);  // example/test.py:181-185
    CHECK(g.getDisplacement(1)[0] == -1.5559 && g.getDisplacement(1)[1] == -0.3 && g.getDisplacement(1)[2] == 0.86);
    g.setBiasMethod(LangevinIntegratorSDM::ILogisticMethod);
    g.setSoftCoreMethod(LangevinIntegratorSDM::RationalMethod);
    g.setLambda1(0.025); g.setLambda2(0.5); g.setAlpha(0.1); g.setU0(110.0); g.setW0coeff(2.0);
    g.setUmax(418.4); g.setUbcore(209.2); g.setAcore(0.0625);
    CHECK(g.getBiasMethod() == 2 && g.getSoftCoreMethod() == 2 && g.getLambda1() == 0.025 && g.getLambda2() == 0.5);
    CHECK(g.getUmax() == 418.4 && g.getUbcore() == 209.2 && g.getAcore() == 0.0625 && g.getW0coeff() == 2.0);
    bool threw = false;
    try { g.setDisplacement(4, 0, 0, 0); } catch (const SDMPlugin::SDMException&) { threw = true; }
    CHECK(threw);
    threw = false;
    std::vector<double> x(12, 0.0), f(12, 0.0);
    try { g.evaluate(x.data(), nullptr, 0.0, f.data()); } catch (const SDMPlugin::SDMException&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { g.step(1); } catch (const SDMPlugin::SDMException&) { threw = true; }
    CHECK(threw);
    threw = false;   // implicit solvent needs a bound context as well
    try { g.addImplicitSolventHCT(nullptr, x.data(), x.data()); } catch (const SDMPlugin::SDMException&) { threw = true; }
    CHECK(threw);
    if (sdm_device_count() == 0) {
        // no CPU fallback: binding must fail loudly
        std::vector<double> q(4, 0.1), sg(4, 0.3), ep(4, 0.5);
        sdm_system s;
        std::memset(&s, 0, sizeof(s));
        s.n_atoms = 4; s.method = SDM_NOCUTOFF; s.charge = q.data(); s.sigma = sg.data(); s.epsilon = ep.data();
        threw = false;
        try { g.bind(s); } catch (const SDMPlugin::SDMException& e) { threw = std::strstr(e.what(), "no CPU fallback") != nullptr; }
        CHECK(threw);
    }
    if (sdm_device_count() > 0) {
        // with a device: bind a tiny system, one evaluation, then two steps of device dynamics
        const int n = 4;
        std::vector<double> q = {0.3, -0.3, 0.2, -0.2}, sg(n, 0.3), ep(n, 0.5), m = {12.0, 16.0, 1.0, 0.0};
        std::vector<double> pos = {0.0, 0.0, 0.0, 0.45, 0.0, 0.0, 0.0, 0.5, 0.0, 0.0, 0.0, 0.55};
        sdm_system s;
        std::memset(&s, 0, sizeof(s));
        s.n_atoms = n; s.method = SDM_NOCUTOFF; s.charge = q.data(); s.sigma = sg.data(); s.epsilon = ep.data();
        SDMPlugin::LangevinIntegratorSDM h(300.0, 1.0, 0.001, n);
        h.setDisplacement(3, 1.0, 0.0, 0.0);
        h.bind(s);
        std::vector<double> force(3 * n);
        h.evaluate(pos.data(), nullptr, 0.0, force.data());
        const double pe0 = h.getPotEnergy();
        {   // implicit solvent on a second integrator: the GB energy of both states enters PotEnergy / BindE
            SDMPlugin::LangevinIntegratorSDM gb(300.0, 1.0, 0.001, n);
            gb.setDisplacement(3, 1.0, 0.0, 0.0);
            gb.bind(s);
            std::vector<double> orad = {0.15, 0.16, 0.11, 0.14}, srad = {0.12, 0.13, 0.09, 0.11}, fgb(3 * n);
            gb.addImplicitSolventHCT(nullptr, orad.data(), srad.data());
            gb.evaluate(pos.data(), nullptr, 0.0, fgb.data());
            CHECK(std::isfinite(gb.getPotEnergy()) && std::fabs(gb.getPotEnergy() - pe0) > 1.0);
            CHECK(std::fabs(gb.getBindE() - h.getBindE()) > 1e-6);
            gb.cleanup();
        }
        threw = false;
        try { h.step(1); } catch (const SDMPlugin::SDMException&) { threw = true; }   // no masses yet
        CHECK(threw);
        h.setState(pos.data(), nullptr, m.data());
        h.step(2);
        std::vector<double> x1(3 * n), v1(3 * n);
        h.getPositions(x1.data());
        h.getVelocities(v1.data());
        double moved = 0;
        for (int i = 0; i < 9; i++) moved += std::fabs(x1[i] - pos[i]);
        CHECK(moved > 0.0);
        for (int d = 0; d < 3; d++) CHECK(x1[9 + d] == pos[9 + d] && v1[9 + d] == 0.0);   // massless particle stays
        CHECK(h.computeKineticEnergy() > 0.0);
        CHECK(std::isfinite(h.getPotEnergy()) && std::isfinite(pe0));
        h.cleanup();
    }
    std::printf("host api ok\n");
    return 0;
}
