"""CPU tests of the HCT-GB oracle (oracle/gb.py: the expression strings of OpenMM's GBSAHCTForce restated with
torch, forces by autograd) and of the host-side mirror around it (GBSAHCTForce, the reader's implicitSolvent=HCT
branch, desmonddmsfile75.py:290-313 and :441-467).  The oracle has no reference-held artefact to be pinned on (no
`hct` table in either shipped fixture, no OpenMM here): these tests hold it to closed forms and to finite differences."""
import sqlite3

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import system as S
from openmm_sdm_plugin_b200.dms import DesmondDMSFile
from oracle import gb as G
from test_dms import make_dms


def gb_parameters(n, seed=0):
    """Plausible HCT parameters: radii 0.12-0.20 nm, scale 0.72-0.88 -> (or, sr)."""
    rng = np.random.default_rng(seed)
    radius = rng.uniform(0.12, 0.20, size=n)
    scale = rng.uniform(0.72, 0.88, size=n)
    o = radius - G.HCT_OFFSET
    return o, scale * o


def blob(n, seed=0, spacing=0.16):
    """n atoms on a jittered lattice (no overlaps closer than ~0.1 nm), charges summing to zero."""
    rng = np.random.default_rng(seed)
    m = int(np.ceil(n ** (1 / 3)))
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n].astype(np.float64)
    pos = g * spacing + rng.uniform(-0.02, 0.02, size=(n, 3))
    q = rng.uniform(-0.6, 0.6, size=n)
    return pos, q - q.mean()


def test_single_ion_is_the_born_energy_plus_ace():
    o, sr = np.array([0.15]), np.array([0.12])
    e, f, B = G.hct(np.zeros((1, 3)), [1.0], o, sr)
    pref = G.GB_COULOMB * (1.0 - 1.0 / 78.5)
    radius = 0.15 + 0.009
    assert B[0] == pytest.approx(0.15, rel=1e-15)
    assert e == pytest.approx(-0.5 * pref / 0.15 + G.ACE_COEFF * (radius + 0.14) ** 2 * (radius / 0.15) ** 6, rel=1e-14)
    assert np.all(f == 0.0)
    e0, _, _ = G.hct(np.zeros((1, 3)), [1.0], o, sr, sa_ace=False)
    assert e0 == pytest.approx(-0.5 * pref / 0.15, rel=1e-14)


def test_two_distant_ions_approach_coulomb_screening():
    o, sr = np.array([0.15, 0.17]), np.array([0.12, 0.13])
    pos = np.array([[0.0, 0, 0], [40.0, 0, 0]])
    e, _, B = G.hct(pos, [1.0, -1.0], o, sr, sa_ace=False)
    pref = G.GB_COULOMB * (1.0 - 1.0 / 78.5)
    assert np.allclose(B, o, rtol=1e-6)                       # descreening vanishes with distance
    assert e == pytest.approx(-0.5 * pref * (1 / B[0] + 1 / B[1]) + pref / 40.0, rel=1e-9)


def test_pair_descreening_integral_closed_form():
    """Two atoms, r > or + sr: L = r - sr, U = r + sr, the HCT integral in closed form."""
    o, sr = np.array([0.15, 0.16]), np.array([0.11, 0.12])
    r = 0.5
    _, _, B = G.hct(np.array([[0.0, 0, 0], [r, 0, 0]]), [0.3, -0.2], o, sr)
    for i, j in ((0, 1), (1, 0)):
        L, U, s = r - sr[j], r + sr[j], sr[j]
        I = 0.5 * (1 / L - 1 / U + 0.25 * (r - s * s / r) * (1 / U ** 2 - 1 / L ** 2) + 0.5 * np.log(L / U) / r)
        assert B[i] == pytest.approx(1.0 / (1.0 / o[i] - I), rel=1e-14)


def test_forces_are_the_gradient_by_central_differences():
    n = 24
    pos, q = blob(n, seed=3)
    o, sr = gb_parameters(n, seed=3)
    e, f, _ = G.hct(pos, q, o, sr)
    h = 1e-5
    rng = np.random.default_rng(0)
    for a, k in zip(rng.integers(0, n, 8), rng.integers(0, 3, 8)):
        p1, p2 = pos.copy(), pos.copy()
        p1[a, k] += h
        p2[a, k] -= h
        num = -(G.hct(p1, q, o, sr)[0] - G.hct(p2, q, o, sr)[0]) / (2 * h)
        assert f[a, k] == pytest.approx(num, rel=1e-6, abs=1e-6)
    assert np.abs(f.sum(0)).max() < 1e-9 * np.abs(f).max()    # translation invariance


def test_gbsahct_force_mirror_converts_like_finalize():
    gb = S.GBSAHCTForce(SA="ACE")
    gb.addParticle([0.4, 0.17, 0.8])
    gb.addParticle([-0.4, 0.12, 0.85])
    with pytest.raises(ValueError):
        gb.device_parameters()
    gb.finalize()
    q, o, sr = gb.device_parameters()
    assert np.allclose(q, [0.4, -0.4])
    assert np.allclose(o, [0.161, 0.111], rtol=1e-15) and np.allclose(sr, [0.8 * 0.161, 0.85 * 0.111], rtol=1e-15)
    assert gb.getNumParticles() == 2 and gb.getParticleParameters(1) == pytest.approx((-0.4, 0.111, 0.85 * 0.111))
    assert gb.solventDielectric == 78.5 and gb.soluteDielectric == 1.0
    with pytest.raises(ValueError):
        S.GBSAHCTForce(SA="LCPO")
    with pytest.raises(ValueError):
        S.GBSAHCTForce(cutoff=1.0)


def test_reader_adds_hct_gb_from_the_hct_table(tmp_path):
    p = str(tmp_path / "gb.dms")
    make_dms(p, with_cell=False)
    with DesmondDMSFile(p) as d:
        with pytest.raises(IOError):                          # desmonddmsfile75.py:467
            d.createSystem(nonbondedMethod=S.NOCUTOFF, implicitSolvent=S.HCT)
        with pytest.raises(NotImplementedError):              # external plugins, :469-526
            d.createSystem(nonbondedMethod=S.NOCUTOFF, implicitSolvent="AGBNP3")
        with pytest.raises(ValueError):                       # :443
            d.createSystem(nonbondedMethod=S.NOCUTOFF, implicitSolvent="OBC")
    conn = sqlite3.connect(p)
    conn.execute("CREATE TABLE hct (id INTEGER PRIMARY KEY, charge FLOAT, radius FLOAT, screened_radius FLOAT)")
    rows = [(i, 0.1 * (i - 2), 1.5 + 0.1 * i, 0.8 + 0.01 * i) for i in range(6)]
    conn.executemany("INSERT INTO hct VALUES (?, ?, ?, ?)", rows)
    conn.commit()
    conn.close()
    with DesmondDMSFile(p) as d:
        plain = d.createSystem(nonbondedMethod=S.NOCUTOFF)
        sysd = d.createSystem(nonbondedMethod=S.NOCUTOFF, implicitSolvent=S.HCT)
    assert plain.gb is None and plain.eps_rf == 78.3
    assert sysd.eps_rf == 1.0                                 # nb.setReactionFieldDielectric(1.0), :451
    gb = sysd.gb
    assert gb.SA == "ACE" and gb.force_group == 2 and gb.getNumParticles() == 6
    q, o, sr = gb.device_parameters()
    for i, (_, charge, radius, screen) in enumerate(rows):
        passed = radius * 0.1 - 0.009                         # what the reader hands to addParticle (:307-309)
        assert q[i] == pytest.approx(charge)
        assert o[i] == pytest.approx(passed - 0.009, rel=1e-14)           # finalize() takes the offset off again
        assert sr[i] == pytest.approx(screen * passed * (passed - 0.009), rel=1e-14)
