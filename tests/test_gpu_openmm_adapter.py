"""The OpenMM-side adapter (openmm_sdm_plugin_b200/csrc/openmm/: B200IntegrateLangevinStepSDMKernel +
factory) under the reference's UNMODIFIED integrator.

oracle/_ref/libsdmb200_openmm.so holds the reference's LangevinIntegratorSDM.cpp compiled where it lies
plus the adapter sources, against the OpenMM stand-in headers; registerKernelFactories() -- the
plugin entry point, same shape as ReferenceSDMKernelFactory.cpp:41-63 -- installs the B200 kernel and
LangevinIntegratorSDM::step (LangevinIntegratorSDM.cpp:153-183) drives it through the seven virtuals
of SDMKernels.h:60-111.  The same steps run through oracle/_ref/libsdmref.so with the reference's own
Reference-platform kernel; the two trajectories are compared.  The oracle only plays OpenMM's
NonbondedForce for the REFERENCE arm (and for level B, where OpenMM keeps that force)."""
import numpy as np
import pytest

from openmm_sdm_plugin_b200 import system as S
from oracle import oracle as O
from oracle import reference as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (R.available() and R.b200_adapter_available()),
                                 reason="oracle/_ref is built only where /root/reference exists")]

T, GAMMA, STEPS = 300.0, 2.0, 3


def _setup(seed):
    case = S.cfg1()
    n = case.system.n_atoms
    rng = np.random.default_rng(seed)
    vel = rng.normal(scale=0.4, size=(n, 3))
    xi = rng.normal(size=(STEPS, n, 3))
    fb = rng.normal(scale=30.0, size=(n, 3))       # a fixed bonded + restraint force, energy 12.5

    def nonbonded(pos):
        r = O.nonbonded(case.system, pos, nthreads=1)
        return r["E"], r["forces"]
    return case, vel, xi, fb, nonbonded


def test_fused_level_matches_the_reference_kernel_for_three_steps():
    case, vel, xi, fb, nonbonded = _setup(31)
    asked = []

    def ref_force(groups, pos):
        return nonbonded(pos) if groups == 4 else (12.5, fb)

    def b200_force(groups, pos):
        asked.append(groups)
        # the B200NonbondedForce evaluates to nothing on the OpenMM side; group 1 is OpenMM's
        return (0.0, np.zeros_like(pos)) if groups == 4 else (12.5, fb)

    p = R.params_from_alch(case.alch, temperature=T, friction=GAMMA)
    ref = R.run(case.masses, case.positions, vel, case.displacement, p, ref_force, steps=STEPS, noise=xi.ravel())
    got = R.run_b200(0, case.system, case.masses, case.positions, vel, case.displacement, p, b200_force,
                     steps=STEPS, noise=xi.ravel())
    assert asked == [4, 4, 2] * STEPS              # the reference's own step sequence drove the kernel
    assert got["step_count"] == ref["step_count"] == STEPS
    assert got["time"] == pytest.approx(ref["time"], abs=1e-15)
    # energies: north_star's 1e-5 relative; observed ~1e-7 (FP32 pair terms in the state-1 sum)
    assert np.allclose(got["traj"][:, 0], ref["traj"][:, 0], rtol=1e-5, atol=1e-5)       # BindE per step
    assert np.allclose(got["traj"][:, 1], ref["traj"][:, 1], rtol=1e-5, atol=0)          # PotEnergy per step
    rms = np.sqrt((ref["hybrid_force"] ** 2).sum(1).mean())
    assert np.sqrt(((got["hybrid_force"] - ref["hybrid_force"]) ** 2).sum(1).mean()) <= 1e-4 * rms
    assert np.abs(got["positions"] - ref["positions"]).max() < 1e-8                     # nm after three steps
    assert np.abs(got["velocities"] - ref["velocities"]).max() < 1e-5 * np.abs(ref["velocities"]).max()
    assert got["kinetic_energy"] == pytest.approx(ref["kinetic_energy"], rel=1e-5)


def test_literal_level_matches_the_reference_kernel_in_single_precision():
    """Level B: OpenMM (the callback) evaluates both states; SaveState1 / MakeState2 / SaveState2 /
    RestoreState1 / the force mix / both integration kernels are the device operations, in the single
    precision of the reference's OpenCL kernels (langevin.cl)."""
    case, vel, xi, fb, nonbonded = _setup(32)

    def force(groups, pos):
        return nonbonded(pos) if groups == 4 else (12.5, fb)

    p = R.params_from_alch(case.alch, temperature=T, friction=GAMMA)
    ref = R.run(case.masses, case.positions, vel, case.displacement, p, force, steps=STEPS, noise=xi.ravel())
    got = R.run_b200(1, case.system, case.masses, case.positions, vel, case.displacement, p, force,
                     steps=STEPS, noise=xi.ravel())
    assert got["step_count"] == STEPS
    # state 2 is formed from float32 coordinates: |du| ~ |F| * 1e-7 nm per displaced atom
    assert np.allclose(got["traj"][:, 0], ref["traj"][:, 0], rtol=1e-4, atol=2e-2)
    assert np.allclose(got["traj"][:, 1], ref["traj"][:, 1], rtol=1e-5, atol=0)
    rms = np.sqrt((ref["hybrid_force"] ** 2).sum(1).mean())
    assert np.sqrt(((got["hybrid_force"] - ref["hybrid_force"]) ** 2).sum(1).mean()) <= 1e-4 * rms
    assert np.abs(got["positions"] - ref["positions"]).max() < 5e-7
    assert np.abs(got["velocities"] - ref["velocities"]).max() < 2e-4 * np.abs(ref["velocities"]).max()


def test_non_equilibrium_schedule_is_written_back_to_the_integrator():
    case, vel, xi, fb, nonbonded = _setup(33)
    import dataclasses
    al = dataclasses.replace(case.alch, nonequilibrium=1, noneq_tmax=0.05, m_lambda1=0.3, m_lambda2=0.5,
                             b_lambda1=0.0, b_lambda2=0.1, m_u0=2.0, b_u0=1.0, alpha=0.2, work_value=0.0)

    def ref_force(groups, pos):
        return nonbonded(pos) if groups == 4 else (0.0, np.zeros_like(pos))

    def b200_force(groups, pos):
        return 0.0, np.zeros_like(pos)

    p = R.params_from_alch(al, temperature=T, friction=GAMMA)
    ref = R.run(case.masses, case.positions, vel, case.displacement, p, ref_force, steps=STEPS, noise=xi.ravel())
    got = R.run_b200(0, case.system, case.masses, case.positions, vel, case.displacement, p, b200_force,
                     steps=STEPS, noise=xi.ravel())
    for k in ("lambdac", "lambda1", "lambda2", "u0", "w0coeff"):
        assert got[k] == pytest.approx(ref[k], rel=1e-12, abs=1e-15), k
    assert got["work_value"] == pytest.approx(ref["work_value"], rel=1e-5, abs=1e-7)


def test_fused_level_with_the_opls_rule_and_hct_gb_in_the_nonbonded_group():
    """createSystem(OPLS=True, implicitSolvent=HCT): the B200NonbondedForce carries the combining rule and the
    GBSAHCTForce parameters; the reference arm gets both from the oracles playing OpenMM's force group 2."""
    import copy
    from oracle import gb as G
    case, vel, xi, fb, _ = _setup(41)
    n = case.system.n_atoms
    sysd = copy.copy(case.system)
    sysd.lj_geometric, sysd.use_dispersion_correction, sysd.eps_rf = True, False, 1.0
    rng = np.random.default_rng(3)
    gb = S.GBSAHCTForce(SA="ACE")
    for a in range(n):
        gb.addParticle([sysd.charge[a], rng.uniform(0.12, 0.2), rng.uniform(0.72, 0.88)])
    gb.finalize()
    sysd.addForce(gb)
    q, o, sr = gb.device_parameters()

    def group2(pos):
        r = O.nonbonded(sysd, pos, nthreads=1)
        e, f, _ = G.hct(pos, q, o, sr)
        return r["E"] + e, r["forces"] + f

    def ref_force(groups, pos):
        return group2(pos) if groups == 4 else (12.5, fb)

    def b200_force(groups, pos):
        return (0.0, np.zeros_like(pos)) if groups == 4 else (12.5, fb)

    p = R.params_from_alch(case.alch, temperature=T, friction=GAMMA)
    ref = R.run(case.masses, case.positions, vel, case.displacement, p, ref_force, steps=STEPS, noise=xi.ravel())
    got = R.run_b200(0, sysd, case.masses, case.positions, vel, case.displacement, p, b200_force,
                     steps=STEPS, noise=xi.ravel())
    plain = R.run_b200(0, case.system, case.masses, case.positions, vel, case.displacement, p, b200_force,
                       steps=1, noise=xi.ravel())
    assert abs(plain["traj"][0, 1] - got["traj"][0, 1]) > 1.0        # the two additions are visible in PotEnergy
    assert np.allclose(got["traj"][:, 0], ref["traj"][:, 0], rtol=1e-5, atol=1e-5)
    assert np.allclose(got["traj"][:, 1], ref["traj"][:, 1], rtol=1e-5, atol=0)
    rms = np.sqrt((ref["hybrid_force"] ** 2).sum(1).mean())
    assert np.sqrt(((got["hybrid_force"] - ref["hybrid_force"]) ** 2).sum(1).mean()) <= 1e-4 * rms
    assert np.abs(got["positions"] - ref["positions"]).max() < 1e-8
