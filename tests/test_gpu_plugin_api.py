"""The reference's user script, example/test.py:24-42,170-191, written against the mirror of the
plugin surface: same setter names, same values, same units -- checked against the oracle."""
import numpy as np
import pytest

from openmm_sdm_plugin_b200 import system as S
from openmm_sdm_plugin_b200.sdmplugin import LangevinIntegratorSDM, OpenMMException, SDMUtils
from oracle import oracle as O

pytestmark = pytest.mark.gpu
KCAL = 4.184


def make_integrator(case):
    n = case.system.n_atoms
    temperature, frictionCoeff, MDstepsize = 300.0, 0.5, 0.001          # example/test.py:168-170
    integrator = LangevinIntegratorSDM(temperature, frictionCoeff, MDstepsize, n)
    sdm_utils = SDMUtils(None)
    integrator.setBiasMethod(sdm_utils.ILogisticMethod)                   # example/test.py:177-179
    integrator.setLambda1(case.alch.lambda1)
    integrator.setLambda2(case.alch.lambda2)
    integrator.setAlpha(case.alch.alpha)
    integrator.setU0(case.alch.u0)
    integrator.setW0coeff(case.alch.w0coeff)
    integrator.setSoftCoreMethod(sdm_utils.RationalSoftCoreMethod)        # example/test.py:189-191
    integrator.setUmax(case.alch.umax)
    integrator.setUbcore(case.alch.ubcore)
    integrator.setAcore(case.alch.acore)
    for i in np.flatnonzero(np.abs(case.displacement).sum(axis=1) > 0):   # example/test.py:181-185
        integrator.setDisplacement(int(i), *case.displacement[i])
    return integrator


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2"])
def test_script_level_parity(cfg):
    case = getattr(S, cfg)()
    integrator = make_integrator(case).bind(case.system)
    ref = O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement, case.positions,
                     nthreads=O.max_threads())
    f = integrator.evaluate(case.positions)
    assert abs(integrator.getBindE() - ref["u_sc"]) <= 1e-6 * max(1.0, abs(ref["u_sc"]))
    escale = max(abs(ref["pot_energy"]), 10.0 * case.system.n_atoms)
    assert abs(integrator.getPotEnergy() - ref["pot_energy"]) <= 1e-5 * escale
    err = np.sqrt(((f - ref["forces"]) ** 2).sum() / (ref["forces"] ** 2).sum())
    assert err <= 1e-4, err
    # a later setDisplacement takes effect (the reference needs Context::reinitialize for that)
    integrator.setDisplacement(0, 0.0, 0.0, 0.0)
    integrator.evaluate(case.positions)
    with pytest.raises(OpenMMException):
        integrator.bind(case.system)         # "This Integrator is already bound to a context"
    integrator.cleanup()


def test_unknown_softcore_method_raises_like_the_reference():
    case = S.cfg1()
    integrator = make_integrator(case).bind(case.system)
    integrator.setSoftCoreMethod(7)
    integrator.setUbcore(-1e9)               # so that u > ub and the method switch is reached
    with pytest.raises(OpenMMException, match="Unknown soft core method"):
        integrator.evaluate(case.positions)
    integrator.cleanup()


def test_step_runs_the_dynamics_on_the_device_like_the_reference():
    """integrator.step(n) of the mirror = the reference's step with the state on the device: two
    steps of the 230-atom fixture with the reference's noise follow oracle/_ref (the reference's own
    integrator + kernels compiled in place) to 1e-8 nm; BindE/PotEnergy are those of the last step."""
    from oracle import oracle as O
    from oracle import reference as R
    if not R.available():
        pytest.skip("oracle/_ref is built only where /root/reference exists")
    case = S.cfg1()
    n = case.system.n_atoms
    rng = np.random.default_rng(21)
    vel = rng.normal(scale=0.3, size=(n, 3))
    xi = rng.normal(size=(2, n, 3))

    def force_fn(groups, pos):
        if groups == 4:
            r = O.nonbonded(case.system, pos, nthreads=1)
            return r["E"], r["forces"]
        return 0.0, np.zeros_like(pos)
    ref = R.run(case.masses, case.positions, vel, case.displacement,
                R.params_from_alch(case.alch, temperature=300.0, friction=0.5), force_fn, steps=2, noise=xi.ravel())

    integ = LangevinIntegratorSDM(300.0, 0.5, case.alch.step_size, n)
    integ.setBiasMethod(case.alch.bias_method)
    integ.setSoftCoreMethod(case.alch.softcore_method)
    integ.setLambda1(case.alch.lambda1); integ.setLambda2(case.alch.lambda2); integ.setAlpha(case.alch.alpha)
    integ.setU0(case.alch.u0); integ.setW0coeff(case.alch.w0coeff)
    integ.setUmax(case.alch.umax); integ.setUbcore(case.alch.ubcore); integ.setAcore(case.alch.acore)
    for i in np.nonzero(np.abs(case.displacement).sum(1))[0]:
        integ.setDisplacement(int(i), *case.displacement[i])
    integ.bind(case.system)
    try:
        with pytest.raises(Exception, match="masses"):
            integ.step(1)
        integ.setState(case.positions, vel, case.masses)
        for k in range(2):
            if k == 0:
                integ.step(0)                       # creates the dynamics object, moves nothing
            integ._ctx.md_set_noise(xi[k][None])    # test hook: the reference's normals
            integ.step(1)
        assert np.abs(integ.getPositions() - ref["positions"]).max() < 1e-8
        assert np.abs(integ.getVelocities() - ref["velocities"]).max() < 1e-5 * np.abs(ref["velocities"]).max()
        assert integ.getBindE() == pytest.approx(ref["bind_e"], abs=1e-6)
        assert integ.getPotEnergy() == pytest.approx(ref["pot_energy"], rel=1e-5)
        assert integ.computeKineticEnergy() == pytest.approx(ref["kinetic_energy"], rel=1e-5)
    finally:
        integ.cleanup()


def test_cpp_host_mirror_on_the_device(tmp_path):
    """tests/hostapi/host_api_check.cpp with a device present: the C++ mirror binds a tiny system,
    evaluates, and runs two steps of device dynamics (massless particle stays, the rest moves)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_api_check")
    libdir = os.path.join(root, "openmm_sdm_plugin_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(root, "tests", "hostapi", "host_api_check.cpp"),
                           "-o", exe, "-L" + libdir, "-lsdmb200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "host api ok" in out.stdout, out.stdout + out.stderr
