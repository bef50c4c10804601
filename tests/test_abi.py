"""C-ABI surface: the product library loads and exports every symbol include/sdmb200.h
declares; without a CUDA device it refuses to create a context (no CPU fallback); host-only
entry points behave like the reference's host arithmetic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sdmb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdm_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 30
    assert sorted(_lib.SYMBOLS) == names


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert _lib.lib().sdm_abi_version() == 2


def test_struct_layouts_match_the_header():
    assert C.sizeof(_lib.SdmAlch) == 2 * 4 + 11 * 8 + 2 * 4 + 4 * 8 + 8 * 8
    assert C.sizeof(_lib.SdmScalars) == 14 * 8 + 3 * 8 + 2 * 4
    assert C.sizeof(_lib.SdmSystem) == 2 * 4 + 2 * 8 + 3 * 8 + 4 * 4 + 7 * 8 + 2 * 8 + 2 * 4
    assert C.sizeof(_lib.SdmOptions) == 2 * 4 + 8 + 2 * 4 + 8 * 4


def test_defaults_are_the_integrator_ctor_defaults():
    """openmmapi/src/LangevinIntegratorSDM.cpp:48-85"""
    a = _lib.SdmAlch()
    _lib.lib().sdm_default_alch(C.byref(a))
    assert (a.bias_method, a.softcore_method) == (0, 0)
    assert (a.lambdac, a.gammac, a.wbcoeff, a.w0coeff) == (1.0, 0.0, 1.0, 0.0)
    assert (a.lambda1, a.lambda2, a.alpha, a.u0) == (1.0, 1.0, 1.0, 0.0)
    assert (a.umax, a.acore, a.ubcore) == (200.0, 0.25, 0.0)
    assert a.nonequilibrium == 0 and a.work_value == 0.0
    d = S.AlchemicalState()
    for k in ("lambdac", "gammac", "wbcoeff", "lambda1", "alpha", "umax", "acore", "ubcore"):
        assert getattr(a, k) == getattr(d, k)


def test_execute_scalars_host_entry_matches_oracle():
    L = _lib.lib()
    rng = np.random.default_rng(0)
    for _ in range(50):
        al = S.AlchemicalState(bias_method=int(rng.integers(0, 3)), softcore_method=int(rng.integers(0, 3)),
                               lambdac=rng.uniform(0, 1), gammac=rng.uniform(0, 0.1), wbcoeff=rng.uniform(0, 1),
                               w0coeff=rng.uniform(-1, 1), lambda1=rng.uniform(0, 0.5), lambda2=rng.uniform(0, 1),
                               alpha=rng.uniform(0, 0.2), u0=rng.uniform(-5, 300), umax=418.4, acore=0.0625,
                               ubcore=209.2)
        E1, E2, Eb = rng.uniform(-1e5, 0), 0.0, rng.uniform(0, 50)
        E2 = E1 + rng.uniform(-50, 1200)
        from openmm_sdm_plugin_b200.context import alch_to_c
        c = alch_to_c(al)
        sc = _lib.SdmScalars()
        _lib.check(L.sdm_execute_scalars(C.byref(c), E1, E2, Eb, C.byref(sc)))
        usc, fp = O.softcore(al.softcore_method, E2 - E1, al.umax, al.acore, al.ubcore)
        eb, bfp = O.bias(S.AlchemicalState(**vars(al)), usc)
        assert sc.u_sc == pytest.approx(usc, rel=1e-14) and sc.fp == pytest.approx(fp, rel=1e-13)
        assert sc.ebias == pytest.approx(eb, rel=1e-13, abs=1e-12) and sc.bfp == pytest.approx(bfp, rel=1e-13)
        assert sc.sp == pytest.approx(bfp * fp, rel=1e-13)
        assert sc.pot_energy == pytest.approx(E1 + eb + Eb, rel=1e-14)
        assert sc.bind_e == sc.u_sc


def test_execute_scalars_unknown_softcore():
    c = _lib.SdmAlch()
    _lib.lib().sdm_default_alch(C.byref(c))
    c.softcore_method = 5
    sc = _lib.SdmScalars()
    assert _lib.lib().sdm_execute_scalars(C.byref(c), 0.0, -1.0, 0.0, C.byref(sc)) == 0   # u <= ub
    assert _lib.lib().sdm_execute_scalars(C.byref(c), 0.0, 10.0, 0.0, C.byref(sc)) == _lib.SDM_ERR_SOFTCORE
    assert b"soft core" in _lib.lib().sdm_last_error()


def test_no_cpu_fallback_without_a_device():
    L = _lib.lib()
    if L.sdm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from openmm_sdm_plugin_b200.context import SDMContext
    case = S.cfg1()
    with pytest.raises(_lib.SDMError) as e:
        SDMContext(case.system, case.displacement)
    assert e.value.code == _lib.SDM_ERR_NO_DEVICE


def test_argument_validation_happens_before_device_use():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.sdm_create(None, None, C.byref(h)) == _lib.SDM_ERR_INVALID
    case = S.synthetic_case(600, 30, seed=1, protein_atoms=0)
    case.system.cutoff = 1.2   # box (1.82 nm) < 2*cutoff -> the reference/OpenMM throws
    from openmm_sdm_plugin_b200.context import SDMContext
    with pytest.raises(_lib.SDMError) as e:
        SDMContext(case.system, case.displacement)
    assert e.value.code == _lib.SDM_ERR_BOX
    assert L.sdm_get_scalars(None, 0, None) == _lib.SDM_ERR_INVALID
    assert L.sdm_k_make_state2(None, 10, None, None) == _lib.SDM_ERR_INVALID


def test_cpp_host_mirror_of_the_integrator(tmp_path):
    """The C++ mirror of SDMPlugin::LangevinIntegratorSDM (header only, on top of the C ABI):
    compiled with g++ against libsdmb200.so and run here -- defaults, round trips, errors."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_api_check")
    libdir = os.path.join(root, "openmm_sdm_plugin_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(root, "tests", "hostapi", "host_api_check.cpp"),
                           "-o", exe, "-L" + libdir, "-lsdmb200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "host api ok" in out.stdout, out.stdout + out.stderr
