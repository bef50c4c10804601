"""Pins the restated oracle on the REFERENCE's own code.

oracle/_ref/libsdmref.so is the reference's integrator, Reference-platform kernels, kernel factory
and Langevin dynamics compiled in place from /root/reference against OpenMM stand-in headers
(oracle/Makefile `ref`, oracle/ref_driver.cpp).  OpenMM's force evaluation is the one thing that is
not reference code: it is a callback that returns the oracle's restated NonbondedForce (group mask 4)
and a synthetic bonded force (mask 2).  Everything the plugin itself owns -- the step sequence,
MakeState2 / Save / Restore, SoftCoreF, the bias functions, the energy bookkeeping, the
non-equilibrium schedule and work, the hybrid force and the Langevin update -- is executed by the
reference's unmodified source and compared with the oracle bit for bit (same double-precision
expressions, same libm).
"""
import dataclasses

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import system as S
from oracle import oracle as O
from oracle import reference as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref is built only where /root/reference exists")

UMAX, UB, A = 100.0 * S.KCAL, 50.0 * S.KCAL, 0.0625


def test_softcore_matches_reference_bit_for_bit():
    us = np.concatenate([np.linspace(-500, 3000, 141), [UB, UB + 1e-9, 1e5, 1e8, -1e6]])
    for method in (S.NO_SOFTCORE, S.TANH_SOFTCORE, S.RATIONAL_SOFTCORE):
        for u in us:
            assert O.softcore(method, u, UMAX, A, UB) == R.softcore(method, u, UMAX, A, UB), (method, u)
    # a different (umax, a, ub) triple, including the constructor defaults 200 / 0.25 / 0
    for u in (-3.0, 0.0, 10.0, 150.0, 900.0):
        for method in (1, 2):
            assert O.softcore(method, u, 200.0, 0.25, 0.0) == R.softcore(method, u, 200.0, 0.25, 0.0)


def test_softcore_known_answers_of_the_survey():
    # SURVEY.md appendix A.4 (computed independently during the survey): now confirmed by the reference
    for u, m, usc, fp in [(250, 1, 249.87116697612, 0.990550911516698), (250, 2, 230.602129698129, 0.268201879720804),
                          (500, 2, 253.884299107405, 0.0419471010006756), (1e8, 2, 369.594730216881, 5.38904370080525e-08)]:
        r = R.softcore(m, u, UMAX, A, UB)
        assert r[0] == pytest.approx(usc, rel=1e-13) and r[1] == pytest.approx(fp, rel=1e-12)


def test_unknown_softcore_method_throws_above_ub_only():
    with pytest.raises(ValueError, match="Unknown soft core method"):
        R.softcore(7, UB + 1.0, UMAX, A, UB)
    with pytest.raises(ValueError):
        O.softcore(7, UB + 1.0, UMAX, A, UB)
    # u <= ub is answered before the method switch (LangevinIntegratorSDM.cpp:126-129)
    assert R.softcore(7, UB - 1.0, UMAX, A, UB) == (UB - 1.0, 1.0)
    assert O.softcore(7, UB - 1.0, UMAX, A, UB) == (UB - 1.0, 1.0)


def test_constructor_defaults_match_the_mirrors():
    d = R.defaults()
    a = S.AlchemicalState()
    for k in ("bias_method", "softcore_method", "lambdac", "gammac", "wbcoeff", "w0coeff", "lambda1", "lambda2",
              "alpha", "u0", "umax", "acore", "ubcore", "nonequilibrium", "work_value"):
        assert getattr(a, k) == d[k], k
    from openmm_sdm_plugin_b200.sdmplugin import LangevinIntegratorSDM
    m = LangevinIntegratorSDM(300.0, 0.5, 0.001, 3)
    assert (m.getTemperature(), m.getFriction(), m.getStepSize()) == (d["temperature"], d["friction"], d["step_size"])
    assert (m.getLambda(), m.getGamma(), m.getWBcoeff(), m.getW0coeff()) == (1.0, 0.0, 1.0, 0.0)
    assert (m.getUmax(), m.getAcore(), m.getUbcore()) == (d["umax"], d["acore"], d["ubcore"])


def _force_fn(case, fb, eb, calls):
    def fn(groups, pos):
        calls.append((groups, pos.copy()))
        if groups == 4:
            r = O.nonbonded(case.system, pos, nthreads=1)
            return r["E"], r["forces"]
        assert groups == 2
        return eb, fb
    return fn


def _run_reference(case, alch, fb, eb, steps=1, friction=0.5, temperature=300.0, noise=None, masses=None, vel=None):
    n = case.system.n_atoms
    calls = []
    p = R.params_from_alch(alch, temperature=temperature, friction=friction)
    res = R.run(case.masses if masses is None else masses, case.positions,
                np.zeros((n, 3)) if vel is None else vel, case.displacement, p,
                _force_fn(case, fb, eb, calls), steps=steps, noise=noise)
    return res, calls


@pytest.fixture(scope="module")
def cfg1():
    case = S.cfg1()
    rng = np.random.default_rng(7)
    fb = rng.normal(scale=50.0, size=(case.system.n_atoms, 3))
    return case, fb, -123.456


def test_step_sequence_and_state_handling(cfg1):
    """a1, a4-a9: three evaluations with masks 4, 4, 2 at x, x+d (ALL atoms), x."""
    case, fb, eb = cfg1
    _, calls = _run_reference(case, dataclasses.replace(case.alch), fb, eb)
    assert [g for g, _ in calls] == [4, 4, 2]
    assert np.array_equal(calls[0][1], case.positions)
    assert np.array_equal(calls[1][1], case.positions + case.displacement)
    assert np.array_equal(calls[2][1], case.positions)


@pytest.mark.parametrize("bias,soft", [(S.ILOGISTIC, S.RATIONAL_SOFTCORE), (S.ILOGISTIC, S.TANH_SOFTCORE),
                                       (S.QUADRATIC, S.RATIONAL_SOFTCORE), (S.LINEAR, S.NO_SOFTCORE),
                                       (S.LINEAR, S.RATIONAL_SOFTCORE)])
def test_execute_matches_oracle_bit_for_bit(cfg1, bias, soft):
    """a11-a15 on the 230-atom fixture: BindE, PotEnergy and the hybrid force of the reference's
    execute() against the oracle's, for every bias / soft-core combination the scripts use."""
    case, fb, eb = cfg1
    alch = dataclasses.replace(case.alch, bias_method=bias, softcore_method=soft, lambdac=0.35, lambda1=0.2,
                               lambda2=0.6, alpha=0.05, u0=-4.0, w0coeff=1.5, gammac=0.01, wbcoeff=0.7,
                               umax=8.0, ubcore=1.0, acore=0.0625)   # u = +3.6 kJ/mol: above ub, so the soft core acts
    ref, _ = _run_reference(case, dataclasses.replace(alch), fb, eb)
    ora = O.sdm_eval(case.system, dataclasses.replace(alch), case.displacement, case.positions, fb=fb, eb=eb,
                     nthreads=1)
    assert ora["u"] > alch.ubcore
    assert ref["bind_e"] == ora["bind_e"] == ora["u_sc"]
    assert ref["pot_energy"] == ora["pot_energy"]
    assert np.array_equal(ref["hybrid_force"], ora["forces"])


def test_cfg1_shipped_settings(cfg1):
    case, fb, eb = cfg1
    ref, _ = _run_reference(case, dataclasses.replace(case.alch), None if False else fb, eb)
    ora = O.sdm_eval(case.system, dataclasses.replace(case.alch), case.displacement, case.positions, fb=fb, eb=eb,
                     nthreads=1)
    assert ref["bind_e"] == ora["bind_e"]
    assert ref["bind_e"] == pytest.approx(3.6086162625, abs=1e-8)
    assert ref["pot_energy"] == ora["pot_energy"]
    assert np.array_equal(ref["hybrid_force"], ora["forces"])


def test_nonequilibrium_schedule_and_work(cfg1):
    """a14: lambda = t/t_max, linear schedules written back into the integrator, work += dlambda*dW/dlambda."""
    case, fb, eb = cfg1
    alch = dataclasses.replace(case.alch, bias_method=S.ILOGISTIC, softcore_method=S.RATIONAL_SOFTCORE,
                               nonequilibrium=1, noneq_tmax=0.25, time=0.05, step_size=0.001, work_value=0.75,
                               alpha=0.08, m_lambda1=0.3, b_lambda1=0.05, m_lambda2=0.5, b_lambda2=0.1,
                               m_u0=20.0, b_u0=-5.0, m_w0=2.0, b_w0=0.25)
    ref, _ = _run_reference(case, dataclasses.replace(alch), fb, eb)
    a2 = dataclasses.replace(alch)
    ora = O.sdm_eval(case.system, a2, case.displacement, case.positions, fb=fb, eb=eb, nthreads=1)
    assert ref["lambdac"] == a2.lambdac == 0.05 / 0.25
    assert (ref["lambda1"], ref["lambda2"], ref["u0"], ref["w0coeff"]) == (a2.lambda1, a2.lambda2, a2.u0, a2.w0coeff)
    assert ref["work_value"] == a2.work_value and ref["work_value"] != 0.75
    assert ref["bind_e"] == ora["bind_e"] and ref["pot_energy"] == ora["pot_energy"]
    assert np.array_equal(ref["hybrid_force"], ora["forces"])


def test_force_group_check(cfg1):
    case, fb, eb = cfg1
    p = R.params_from_alch(case.alch)
    with pytest.raises(RuntimeError, match="force group 1"):
        R.run(case.masses, case.positions, np.zeros_like(case.positions), case.displacement, p,
              lambda g, x: (0.0, np.zeros_like(x)), force_groups=(1, 2, 3))


def test_langevin_update_matches_the_formulas(cfg1):
    """N2: the update the reference applies after the hybrid force (ReferenceStochasticDynamicsSDM.cpp:131-266)
    against SURVEY.md appendix A.6 evaluated in numpy, with a known noise sequence."""
    case, fb, eb = cfg1
    n = case.system.n_atoms
    rng = np.random.default_rng(3)
    noise = rng.normal(size=3 * n)
    vel = rng.normal(scale=0.3, size=(n, 3))
    masses = case.masses.copy()
    masses[5] = 0.0                                   # a massless particle does not move
    T, gamma, dt = 300.0, 2.0, case.alch.step_size
    ref, _ = _run_reference(case, dataclasses.replace(case.alch), fb, eb, friction=gamma, temperature=T,
                            noise=noise, masses=masses, vel=vel)
    F = ref["hybrid_force"]
    vscale = np.exp(-dt * gamma)
    fscale = (1 - vscale) / gamma
    kT = 1.380658e-23 * 6.0221367e23 / 1000.0 * T
    noisescale = np.sqrt(kT * (1 - vscale * vscale))
    inv = np.where(masses > 0, 1.0 / np.where(masses > 0, masses, 1.0), 0.0)
    xi = np.zeros((n, 3))
    xi[masses > 0] = noise[:3 * int((masses > 0).sum())].reshape(-1, 3)   # noise is drawn only for massive atoms
    v1 = vscale * vel + fscale * inv[:, None] * F + noisescale * np.sqrt(inv)[:, None] * xi
    x1 = case.positions + dt * v1
    mv = masses > 0
    assert np.allclose(ref["positions"][mv], x1[mv], rtol=0, atol=1e-15)
    assert np.allclose(ref["velocities"][mv], ((x1 - case.positions) / dt)[mv], rtol=1e-9, atol=1e-12)
    assert np.array_equal(ref["positions"][5], case.positions[5])
    assert ref["time"] == pytest.approx(dt) and ref["step_count"] == 1
    ke = 0.5 * np.sum(masses[:, None] * ref["velocities"] ** 2)
    assert ref["kinetic_energy"] == pytest.approx(ke, rel=1e-12)
