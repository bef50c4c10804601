"""The Python mirror of the plugin surface (python/SDMplugin.i:86-145, SDMUtils.py:9-15):
names, defaults and error behaviour -- no GPU needed."""
import re

import pytest

from openmm_sdm_plugin_b200.sdmplugin import LangevinIntegratorSDM, OpenMMException, SDMUtils

# the method list of python/SDMplugin.i:86-145, verbatim names
SWIG_METHODS = """setLambda getLambda getTemperature setTemperature getFriction setFriction getRandomNumberSeed
setRandomNumberSeed getBindE setBindE getPotEnergy setPotEnergy getUmax setUmax getAcore setAcore getUbcore setUbcore
setBiasMethod getBiasMethod setSoftCoreMethod getSoftCoreMethod setGamma getGamma setWBcoeff getWBcoeff setW0coeff
getW0coeff setLambda1 getLambda1 setLambda2 getLambda2 setAlpha getAlpha setU0 getU0 setNoneqtmax getNoneqtmax
getNonEquilibrium setNoneqWorkvalue getNoneqWorkvalue setlambda1Slope getlambda1Slope setlambda2Slope getlambda2Slope
setu0Slope getu0Slope setw0Slope getw0Slope setlambda1intercept getlambda1intercept setlambda2intercept
getlambda2intercept setu0intercept getu0intercept setw0intercept getw0intercept setDisplacement getDisplacement
step""".split()


def test_every_swig_method_exists():
    integ = LangevinIntegratorSDM(300.0, 0.5, 0.001, 10)
    missing = [m for m in SWIG_METHODS if not callable(getattr(integ, m, None))]
    assert not missing, missing


def test_ctor_defaults_match_the_reference():
    # openmmapi/src/LangevinIntegratorSDM.cpp:48-85
    g = LangevinIntegratorSDM(300.0, 0.5, 0.001, 4)
    assert (g.getTemperature(), g.getFriction(), g.getStepSize()) == (300.0, 0.5, 0.001)
    assert (g.getUmax(), g.getAcore(), g.getUbcore()) == (200.0, 0.25, 0.0)
    assert g.getSoftCoreMethod() == LangevinIntegratorSDM.NoSoftCoreMethod == 0
    assert g.getBiasMethod() == LangevinIntegratorSDM.LinearMethod == 0
    assert (g.getLambda(), g.getGamma(), g.getWBcoeff(), g.getW0coeff()) == (1.0, 0.0, 1.0, 0.0)
    assert (g.getLambda1(), g.getLambda2(), g.getAlpha(), g.getU0()) == (1.0, 1.0, 1.0, 0.0)
    assert g.getNonEquilibrium() == 0 and g.getNoneqWorkvalue() == 0.0
    assert all(g.getDisplacement(i) == (0.0, 0.0, 0.0) for i in range(4))


def test_constants_of_sdmutils():
    u = SDMUtils()
    assert (u.LinearMethod, u.QuadraticMethod, u.ILogisticMethod) == (0, 1, 2)
    assert (u.NoSoftCoreMethod, u.TanhSoftCoreMethod, u.RationalSoftCoreMethod) == (0, 1, 2)
    assert (LangevinIntegratorSDM.TanhMethod, LangevinIntegratorSDM.RationalMethod) == (1, 2)


def test_setters_round_trip_and_displacement_map():
    g = LangevinIntegratorSDM(300.0, 0.5, 0.001, 3)
    for name in SWIG_METHODS:
        m = re.match(r"set(.*)", name)
        if not m or name in ("setDisplacement", "setBiasMethod", "setSoftCoreMethod", "setRandomNumberSeed"):
            continue
        getattr(g, name)(0.125)
        assert getattr(g, "get" + m.group(1))() == 0.125, name
    g.setBiasMethod(2); g.setSoftCoreMethod(2); g.setRandomNumberSeed(42)
    assert (g.getBiasMethod(), g.getSoftCoreMethod(), g.getRandomNumberSeed()) == (2, 2, 42)
    g.setDisplacement(1, -1.5559, -0.3, 0.86)           # example/test.py:181-185
    assert g.getDisplacement(1) == (-1.5559, -0.3, 0.86) and g.getDisplacement(0) == (0.0, 0.0, 0.0)
    with pytest.raises(IndexError):
        g.setDisplacement(3, 0, 0, 0)


def test_unbound_integrator_and_step_report_errors():
    g = LangevinIntegratorSDM(300.0, 0.5, 0.001, 3)
    with pytest.raises(OpenMMException):
        g.evaluate([[0, 0, 0]] * 3)
    with pytest.raises(OpenMMException):
        g.step(1)


def test_reference_import_line_resolves():
    """example/test.py:11 and example/test_explicit.py:11: `from SDMplugin import *`."""
    ns = {}
    exec("from SDMplugin import *", ns)
    assert {"LangevinIntegratorSDM", "SDMUtils", "OpenMMException"} <= set(ns)
    from openmm_sdm_plugin_b200 import sdmplugin
    assert ns["LangevinIntegratorSDM"] is sdmplugin.LangevinIntegratorSDM
    integ = ns["LangevinIntegratorSDM"](300.0, 0.5, 0.001, 10)        # SDMplugin.i:86
    integ.setLambda1(0.025)
    assert integ.getLambda1() == 0.025 and integ.ILogisticMethod == 2
