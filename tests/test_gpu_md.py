"""Device-resident Langevin dynamics (sdm_md_*) against the reference's own update code.

oracle/_ref runs ReferenceStochasticDynamicsSDM::update (compiled from the reference's source) after
its execute(); the device kernel evaluates the same double-precision expressions in the same order,
so with the reference's hybrid force and the same normals the positions and velocities must be
bit-identical; with the CUDA path's own force they differ by the force tolerance times dt/m.
"""
import dataclasses

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import oracle as O
from oracle import reference as R

pytestmark = pytest.mark.gpu

T, GAMMA = 300.0, 2.0


def reference_steps(case, masses, vel, noise, steps):
    def force_fn(groups, pos):
        if groups == 4:
            r = O.nonbonded(case.system, pos, nthreads=1)
            return r["E"], r["forces"]
        return 0.0, np.zeros_like(pos)
    p = R.params_from_alch(case.alch, temperature=T, friction=GAMMA)
    return R.run(masses, case.positions, vel, case.displacement, p, force_fn, steps=steps, noise=noise)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref is built only where /root/reference exists")
def test_update_is_bit_identical_to_the_reference_update():
    case = S.cfg1()
    n = case.system.n_atoms
    rng = np.random.default_rng(11)
    masses = case.masses.copy()
    masses[7] = 0.0                                   # a massless particle keeps x and v
    vel = rng.normal(scale=0.4, size=(n, 3))
    xi = rng.normal(size=(n, 3))
    noise_ref = xi[masses > 0].ravel()                # the reference draws only for massive atoms
    ref = reference_steps(case, masses, vel, noise_ref, 1)
    with SDMContext(case.system, case.displacement, n_replicas=2, pair_mode=_lib.PAIR_CLUSTER) as ctx:
        ctx.md_init(masses, T, GAMMA, case.alch.step_size, seed=1)
        for r in range(2):
            ctx.set_positions(r, case.positions)
            ctx.set_alchemical(r, case.alch)
            ctx.md_set_velocities(r, vel)
        ctx.md_set_noise(np.stack([xi, xi]))
        ctx.md_update(np.stack([ref["hybrid_force"], ref["hybrid_force"]]))   # the reference's force in
        for r in range(2):
            assert np.array_equal(ctx.positions(r), ref["positions"])
            assert np.array_equal(ctx.md_velocities(r), ref["velocities"])
            assert ctx.md_kinetic_energy(r) == pytest.approx(ref["kinetic_energy"], rel=1e-13)
        assert np.array_equal(ctx.positions(0)[7], case.positions[7])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref is built only where /root/reference exists")
def test_three_md_steps_follow_the_reference_trajectory():
    """sdm_md_step = eval + update on the device: three steps of the 230-atom fixture with the
    noise sequence the reference consumes; the hybrid force agrees to ~1e-6 relative (FP32 pair
    terms), which is ~1e-9 nm per step in the positions."""
    case = S.cfg1()
    n = case.system.n_atoms
    rng = np.random.default_rng(12)
    vel = rng.normal(scale=0.4, size=(n, 3))
    xi = rng.normal(size=(3, n, 3))
    ref = reference_steps(case, case.masses, vel, xi.ravel(), 3)
    with SDMContext(case.system, case.displacement, n_replicas=1, pair_mode=_lib.PAIR_CLUSTER) as ctx:
        ctx.md_init(case.masses, T, GAMMA, case.alch.step_size, seed=1)
        ctx.set_positions(0, case.positions)
        ctx.set_alchemical(0, case.alch)
        ctx.md_set_velocities(0, vel)
        for k in range(3):
            ctx.md_set_noise(xi[k][None])
            ctx.md_step(1)
        assert ctx.scalars(0)["status"] == 0
        assert np.abs(ctx.positions(0) - ref["positions"]).max() < 1e-8      # nm, after three steps
        dv = np.abs(ctx.md_velocities(0) - ref["velocities"]).max()
        assert dv < 1e-5 * np.abs(ref["velocities"]).max()
        assert ctx.scalars(0)["bind_e"] == pytest.approx(ref["traj"][2, 0], abs=1e-6)


def test_philox_noise_has_the_right_temperature_and_is_reproducible():
    """Free particles (no forces: every atom its own far-away... here simply a cutoff system whose
    forces are swamped by a strong thermostat): after many updates with zero force the velocity
    distribution is Maxwell-Boltzmann at T; two contexts with the same seed agree bit for bit."""
    case = S.synthetic_case(3000, 30, seed=4, protein_atoms=0)
    n = case.system.n_atoms
    masses = np.full(n, 12.0)
    kT = 1.380658e-23 * 6.0221367e23 / 1000.0 * T
    outs = []
    for _ in range(2):
        with SDMContext(case.system, case.displacement, n_replicas=2, pair_mode=_lib.PAIR_CLUSTER) as ctx:
            ctx.md_init(masses, T, 50.0, 0.001, seed=77)
            for r in range(2):
                ctx.set_positions(r, case.positions)
                ctx.set_alchemical(r, case.alch)
            ctx.eval()
            zero = np.zeros((2, n, 3))
            for _k in range(200):
                ctx.md_update(zero if _k == 0 else None)    # forces stay zero: B.F is only written by eval
            outs.append((ctx.md_velocities(0), ctx.md_velocities(1)))
    v0, v1 = outs[0]
    assert np.array_equal(v0, outs[1][0]) and np.array_equal(v1, outs[1][1])
    assert not np.array_equal(v0, v1)                        # replicas draw from different streams
    t_kin = masses[0] * (v0 ** 2).mean() / kT                # <m v_x^2> = kT
    assert t_kin == pytest.approx(1.0, rel=0.05)
    assert abs(v0.mean()) < 4.0 * v0.std() / np.sqrt(v0.size)
    # consecutive steps use disjoint stretches of every atom's stream: the velocity autocorrelation of
    # a free particle is exactly vscale per step
    with SDMContext(case.system, case.displacement, n_replicas=1, pair_mode=_lib.PAIR_CLUSTER) as ctx:
        ctx.md_init(masses, T, 50.0, 0.001, seed=5)
        ctx.set_positions(0, case.positions)
        ctx.set_alchemical(0, case.alch)
        ctx.eval()
        ctx.md_update(np.zeros((1, n, 3)))
        for _k in range(300):
            ctx.md_update()
        a = ctx.md_velocities(0)
        ctx.md_update()
        b = ctx.md_velocities(0)
    corr = (a * b).mean() / (a * a).mean()
    assert corr == pytest.approx(np.exp(-50.0 * 0.001), abs=0.03)


# ---- distance constraints (sdm_md_set_constraints): SETTLE waters + SHAKE clusters -----------------
from oracle import constraints as OC   # noqa: E402


def _thermal(case, seed):
    rng = np.random.default_rng(seed)
    n = case.system.n_atoms
    kT = 1.380658e-23 * 6.0221367e23 / 1000.0 * T
    v = rng.normal(size=(n, 3)) * np.sqrt(kT / case.masses)[:, None]
    return rng, v


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_constrained_update_matches_the_constraint_oracle(name):
    """One update with given forces and normals: part 1 + 2 are the reference's expressions, the
    constraint stage must land where the tightly converged oracle lands (SHAKE clusters to the
    integrator's 1e-5 tolerance, SETTLE waters to rounding), ReferenceStochasticDynamicsSDM.cpp:250-262."""
    case = getattr(S, name)()
    n = case.system.n_atoms
    rng, vel = _thermal(case, 21)
    f = rng.normal(scale=300.0, size=(n, 3))
    xi = rng.normal(size=(n, 3))
    dt = 0.001
    x_ref, v_ref, xp = OC.langevin_step(case.positions, vel, f, case.masses, T, GAMMA, dt, xi,
                                        case.constraint_pairs, case.constraint_dist)
    p, d0 = case.constraint_pairs, case.constraint_dist
    with SDMContext(case.system, case.displacement, n_replicas=2, pair_mode=_lib.PAIR_CLUSTER) as ctx:
        ctx.md_init(case.masses, T, GAMMA, dt, seed=1)
        ctx.md_set_constraints(p, d0, 1e-5)
        for r in range(2):
            ctx.set_positions(r, case.positions)
            ctx.set_alchemical(r, case.alch)
            ctx.md_set_velocities(r, vel)
        ctx.md_set_noise(np.stack([xi, xi]))
        ctx.md_update(np.stack([f, f]))
        for r in range(2):
            x, v = ctx.positions(r), ctx.md_velocities(r)
            d = np.linalg.norm(x[p[:, 0]] - x[p[:, 1]], axis=1)
            assert np.abs(d / d0 - 1).max() < 1.01e-5
            assert np.abs(x - x_ref).max() < 3e-6                       # nm
            assert np.abs(v - v_ref).max() < 3e-6 / dt
            assert np.allclose(v, (x - case.positions) / dt, rtol=0, atol=1e-9)
            # atoms outside every cluster went through the unconstrained expressions exactly
            free = np.ones(n, bool)
            free[p.ravel()] = False
            assert np.array_equal(x[free], xp[free])
            # rigid waters: hydrogen pairs constrained to each other -> SETTLE, exact to rounding
            anum = np.load(S.GOLDEN_DIR + "/%s.npz" % ("cfg1_oa_g6_g3" if name == "cfg1" else "cfg2_temoa_g1_g4"))["anum"]
            hh = (anum[p[:, 0]] == 1) & (anum[p[:, 1]] == 1)
            if hh.any():
                assert np.abs(d[hh] / d0[hh] - 1).max() < 1e-11
                wat = np.unique(p[hh].ravel())
                assert np.abs(x[wat] - x_ref[wat]).max() < 1e-10


def test_constrained_dynamics_of_the_explicit_solvent_fixture():
    """cfg2 with its real masses, thermal velocities and all 20 278 constraints, 40 steps of 1 fs on
    the device (Philox noise): constraints hold to the tolerance, the kinetic temperature stays
    thermal, every step was taken."""
    case = S.cfg2()
    n = case.system.n_atoms
    _, vel = _thermal(case, 5)
    p, d0 = case.constraint_pairs, case.constraint_dist
    kT = 1.380658e-23 * 6.0221367e23 / 1000.0 * T
    with SDMContext(case.system, case.displacement, n_replicas=2, pair_mode=_lib.PAIR_CLUSTER) as ctx:
        ctx.md_init(case.masses, T, GAMMA, 0.001, seed=9)
        ctx.md_set_constraints(p, d0, 1e-5)
        for r in range(2):
            ctx.set_positions(r, case.positions)
            ctx.set_alchemical(r, case.alch)
            ctx.md_set_velocities(r, vel)
        ctx.md_step(40)
        taken, repeated = ctx.md_counters()
        assert taken == 40
        for r in range(2):
            assert ctx.scalars(r)["status"] == 0
            x = ctx.positions(r)
            d = np.linalg.norm(x[p[:, 0]] - x[p[:, 1]], axis=1)
            assert np.abs(d / d0 - 1).max() < 1.01e-5
            ke = ctx.md_kinetic_energy(r)
            t_kin = 2.0 * ke / ((3 * n - len(d0)) * kT) * T
            assert 200.0 < t_kin < 400.0, t_kin
        assert not np.array_equal(ctx.positions(0), ctx.positions(1))      # replicas draw different noise


def test_stale_list_steps_are_repeated_not_integrated():
    """A skin far too small for the list lifetime: the list goes stale inside sdm_md_step, the
    affected steps are not taken, the list is rebuilt and they are repeated -- the trajectory is the
    one a comfortable skin gives (the in-cutoff pair set does not depend on the list; the FP32
    partial sums inside a work unit do, at rounding level: ~1e-9 nm after 60 steps)."""
    case = S.synthetic_case(6000, 30, seed=8)
    n = case.system.n_atoms
    case.masses = case.masses * 100.0      # the lattice start is far from equilibrium: keep the motion gentle
    _, vel = _thermal(case, 6)
    out = {}
    for tag, skin, nstlist in (("wide", 0.16, 8), ("tight", 0.012, 200)):
        with SDMContext(case.system, case.displacement, n_replicas=1, pair_mode=_lib.PAIR_CLUSTER,
                        skin=skin, nstlist=nstlist) as ctx:
            ctx.md_init(case.masses, T, GAMMA, 0.001, seed=4)
            ctx.set_positions(0, case.positions)
            ctx.set_alchemical(0, case.alch)
            ctx.md_set_velocities(0, vel)
            ctx.md_step(60)
            out[tag] = (ctx.positions(0), ctx.md_counters(), ctx.info("n_list_builds"))
    assert out["wide"][1] == (60, 0)
    assert out["tight"][1][0] == 60 and out["tight"][1][1] > 0
    assert np.abs(out["tight"][0] - out["wide"][0]).max() < 1e-7
