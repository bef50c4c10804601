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
