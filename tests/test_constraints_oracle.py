"""The numpy restatement of the constrained Langevin step (oracle/constraints.py) obeys the
defining equations of the reference's constraint stage (ReferenceStochasticDynamicsSDM.cpp:250-262)
on the shipped fixtures' own constraint sets."""
import numpy as np

from openmm_sdm_plugin_b200 import system as S
from oracle import constraints as OC


def _step(case, seed):
    rng = np.random.default_rng(seed)
    n = case.system.n_atoms
    kT = 1.380658e-23 * 6.0221367e23 / 1000.0 * 300.0
    v = rng.normal(size=(n, 3)) * np.sqrt(kT / case.masses)[:, None]
    f = rng.normal(scale=300.0, size=(n, 3))
    xi = rng.normal(size=(n, 3))
    return OC.langevin_step(case.positions, v, f, case.masses, 300.0, 2.0, 0.001, xi,
                            case.constraint_pairs, case.constraint_dist), v


def test_fixture_constraint_sets():
    c1, c2 = S.cfg1(), S.cfg2()
    assert len(c1.constraint_dist) == 86               # X-H stretches of the host-guest pair
    assert len(c2.constraint_dist) == 20278            # 13 544 X-H stretches + 6 734 water H-H
    for c in (c1, c2):
        p = c.constraint_pairs
        d = np.linalg.norm(c.positions[p[:, 0]] - c.positions[p[:, 1]], axis=1)
        # the shipped coordinates sit on the constraint manifold to the writer's precision
        assert np.abs(d - c.constraint_dist).max() < 2e-3


def test_constrained_step_satisfies_the_defining_equations():
    for case, seed in ((S.cfg1(), 1), (S.cfg2(), 2)):
        (x1, v1, xp), v0 = _step(case, seed)
        p, d0 = case.constraint_pairs, case.constraint_dist
        d = np.linalg.norm(x1[p[:, 0]] - x1[p[:, 1]], axis=1)
        assert np.abs(d / d0 - 1).max() < 1e-12
        # the correction conserves the momentum of every cluster: sum_i m_i (x1 - xp)_i = 0
        corr = (x1 - xp) * case.masses[:, None]
        assert np.abs(corr.sum(0)).max() < 1e-9 * np.abs(corr).sum()
        # atoms outside constraint clusters are untouched by the constraint stage
        free = np.ones(len(x1), bool)
        free[p.ravel()] = False
        assert np.array_equal(x1[free], xp[free])
        assert np.allclose(v1, (x1 - case.positions) / 0.001, rtol=0, atol=1e-9)
