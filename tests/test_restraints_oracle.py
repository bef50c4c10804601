"""CPU tests of the restraint oracle (oracle/restraints.py: SDMUtils.py:61-85, :183-256 restated with torch
autograd) and of the SDMUtils mirror that records the terms."""
import math

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import system as S
from openmm_sdm_plugin_b200.sdmplugin import SDMUtils
from oracle import restraints as R

KCAL = 4.184


def frame(seed=0, n=40):
    return np.random.default_rng(seed).normal(size=(n, 3))


def test_centroid_distance_term_is_flat_inside_the_tolerance_and_harmonic_outside():
    pos = np.zeros((4, 3))
    pos[0] = (0.0, 0.0, 0.0); pos[1] = (0.2, 0.0, 0.0)       # ligand centroid at x = 0.1
    pos[2] = (1.0, 0.0, 0.0); pos[3] = (1.2, 0.0, 0.0)       # receptor centroid at x = 1.1: d12 = 1.0
    spec = dict(lig_cm_atoms=[0, 1], rcpt_cm_atoms=[2, 3], kfcm=100.0, tolcm=1.5)
    e, f = R.energy_and_forces(pos, [spec])
    assert e == 0.0 and np.all(f == 0.0)
    spec["tolcm"] = 0.4                                       # (kf/2)(d12 - tol)^2 = 50 * 0.36 (SDMUtils.py:61)
    e, f = R.energy_and_forces(pos, [spec])
    assert abs(e - 18.0) < 1e-12
    assert np.allclose(f[0], (0.5 * 100.0 * 0.6, 0, 0)) and np.allclose(f[2], (-0.5 * 100.0 * 0.6, 0, 0))
    e2, _ = R.energy_and_forces(pos, [spec], control=0.25)   # SDMRestraintControlParameter scales it (:61)
    assert abs(e2 - 4.5) < 1e-12
    spec["offset"] = (-1.0, 0.0, 0.0)                         # d12 = |g1 - off - g2| = 0 (:68)
    assert R.energy_and_forces(pos, [spec])[0] == 0.0


def test_mass_weighted_centroid():
    pos = np.array([[0.0, 0, 0], [1.0, 0, 0], [3.0, 0, 0]])
    spec = dict(lig_cm_atoms=[0, 1], lig_cm_weights=[3.0, 1.0], rcpt_cm_atoms=[2], kfcm=2.0, tolcm=0.0)
    e, f = R.energy_and_forces(pos, [spec])                   # centroid at 0.25: d12 = 2.75
    assert abs(e - 2.75 ** 2) < 1e-12
    assert np.allclose(f[0, 0] / f[1, 0], 3.0)


def test_forces_are_the_negative_gradient_and_sum_to_zero():
    pos = frame(3)
    cs = dict(lig_cm_atoms=[0, 1, 2, 3], rcpt_cm_atoms=list(range(10, 30)), kfcm=100.0, tolcm=0.1, offset=(0.1, 0, 0),
              lig_ref=[0, 1, 2], rcpt_ref=[10, 11, 12], kfcd=[50.0, 60.0, 70.0], a=[0.5, -0.3, 1.0], b=[0.9, 0.3, 1.4])
    al = dict(liga_ref=[0, 1, 2], ligb_ref=[5, 6, 7], kfdispl=1000.0, ktheta=200.0, kpsi=200.0, offset=(0.2, 0.1, 0))
    e, f = R.energy_and_forces(pos, [cs], [al])
    assert e > 0 and np.abs(f.sum(0)).max() < 1e-9            # translation invariance
    h = 1e-6
    for i, d in ((1, 0), (11, 2), (6, 1), (20, 0)):
        p, m = pos.copy(), pos.copy()
        p[i, d] += h; m[i, d] -= h
        fd = -(R.energy_and_forces(p, [cs], [al])[0] - R.energy_and_forces(m, [cs], [al])[0]) / (2 * h)
        assert abs(fd - f[i, d]) <= 1e-6 * max(1.0, abs(fd))


def test_alignment_energy_vanishes_for_a_translated_copy():
    """ligand b = ligand a shifted by the offset: displacement, theta and psi terms are all zero (:166-171)."""
    a = frame(5, 3)
    off = np.array([0.3, -0.2, 0.9])
    pos = np.vstack([a, a + off])
    al = dict(liga_ref=[0, 1, 2], ligb_ref=[3, 4, 5], kfdispl=1000.0, ktheta=200.0, kpsi=200.0, offset=off)
    e, f = R.energy_and_forces(pos, [], [al])
    assert abs(e) < 1e-10 and np.abs(f).max() < 1e-4
    pos[4] += (0.05, 0.0, 0.02)                               # tilt b's first axis: theta term wakes up
    assert R.energy_and_forces(pos, [], [al])[0] > 1e-3


def test_flat_bottom_angle_window():
    pos = np.zeros((8, 3))
    pos[0] = (0, 0, 5.0); pos[1] = (0, 0, -5.0)              # far-apart single-atom "centroids" (distance term off)
    pos[2] = (1.0, 0, 0); pos[3] = (1, 1, 0); pos[4] = (2, 2, 1)          # receptor refs g3..g5
    pos[5] = (0.0, 0, 0); pos[6] = (0.0, 1.0, 0); pos[7] = (0.3, 1.0, 0.7)  # ligand refs g6..g8: angle(g3,g6,g7) = 90 deg
    spec = dict(lig_cm_atoms=[0], rcpt_cm_atoms=[1], kfcm=0.0, tolcm=0.0, lig_ref=[5, 6, 7], rcpt_ref=[2, 3, 4],
                kfcd=[10.0, 0.0, 0.0], a=[math.radians(80), 0, 0], b=[math.radians(100), 0, 0])
    assert R.energy_and_forces(pos, [spec])[0] == 0.0        # inside [80, 100] degrees
    spec["a"][0], spec["b"][0] = math.radians(40), math.radians(60)
    e = R.energy_and_forces(pos, [spec])[0]                  # 30 degrees above the window: (kf/2) * (pi/6)^2 (:63)
    assert abs(e - 5.0 * (math.pi / 6) ** 2) < 1e-9


def test_sdmutils_records_the_terms_with_the_reference_argument_names():
    case = S.cfg1()
    u = SDMUtils(case.system)
    u.addRestraintForce(lig_cm_particles=[1, 2, 3], rcpt_cm_particles=[10, 11], kfcm=25.0 * KCAL * 100, tolcm=0.45,
                        offset=[0.1, 0.0, 0.0])
    u.addRestraintForce(lig_cm_particles=None, rcpt_cm_particles=[1])      # no-op like SDMUtils.py:49-53
    u.addRestraintForce(lig_cm_particles=[1], rcpt_cm_particles=[2], kfcm=1.0, tolcm=0.0, lig_ref_particles=[1, 2, 3],
                        rcpt_ref_particles=[4, 5, 6], angle_center=1.0, angletol=0.25, kfangle=3.0)
    u.addAlignmentForce(liga_ref_particles=[1, 2, 3], ligb_ref_particles=[4, 5, 6], kfdispl=10.0, ktheta=5.0, kpsi=5.0)
    rs = case.system.sdm_restraints
    assert [r["kind"] for r in rs] == ["centroid", "centroid", "alignment"]
    assert rs[1]["a"][0] == 0.75 and rs[1]["b"][0] == 1.25 and rs[1]["kfcd"][0] == 3.0
    with pytest.raises(ValueError):
        u.addRestraintForce(lig_cm_particles=[1], rcpt_cm_particles=[2], lig_ref_particles=[1, 2], rcpt_ref_particles=[4, 5, 6])
    with pytest.raises(ValueError):
        u.addAlignmentForce(liga_ref_particles=[1, 2], ligb_ref_particles=[4, 5, 6])
    assert u.getControlParameterName() == "SDMRestraintControlParameter"
