"""DMS reader / write-back: a synthetic SQLite file everywhere, and the two shipped
fixtures against the committed golden arrays when /root/reference is present (build container only)."""
import os
import sqlite3

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import system as S
from openmm_sdm_plugin_b200.dms import DesmondDMSFile

REF = "/root/reference/example"


def make_dms(path, n=6, with_nbtype=True, with_cell=True, seed=0):
    rng = np.random.default_rng(seed)
    conn = sqlite3.connect(path)
    nb = ", nbtype INTEGER" if with_nbtype else ""
    conn.execute("CREATE TABLE particle (id INTEGER PRIMARY KEY, anum INTEGER, name TEXT, x FLOAT, y FLOAT, "
                 "z FLOAT, vx FLOAT, vy FLOAT, vz FLOAT, resid INTEGER, mass FLOAT, charge FLOAT" + nb + ")")
    xyz = rng.uniform(0, 20, size=(n, 3))
    vel = rng.normal(size=(n, 3))
    for i in range(n):
        row = [i, 6, "C%d" % i, *xyz[i], *vel[i], i // 3, 12.0, 0.1 * (i - 2)]
        if with_nbtype:
            row.append(i % 2)
        conn.execute("INSERT INTO particle VALUES (%s)" % ",".join("?" * len(row)), row)
    conn.execute("CREATE TABLE nonbonded_param (id INTEGER PRIMARY KEY, sigma FLOAT, epsilon FLOAT)")
    conn.execute("INSERT INTO nonbonded_param VALUES (0, 3.4, 0.1)")
    conn.execute("INSERT INTO nonbonded_param VALUES (1, 2.5, 0.0)")
    conn.execute("CREATE TABLE exclusion (p0 INTEGER, p1 INTEGER)")
    conn.executemany("INSERT INTO exclusion VALUES (?, ?)", [(0, 1), (1, 2), (0, 2), (0, 3)])
    conn.execute("CREATE TABLE pair_12_6_es_param (id INTEGER PRIMARY KEY, aij FLOAT, bij FLOAT, qij FLOAT)")
    conn.execute("INSERT INTO pair_12_6_es_param VALUES (0, 1000.0, 10.0, 0.05)")
    conn.execute("INSERT INTO pair_12_6_es_param VALUES (1, 0.0, 0.0, -0.02)")
    conn.execute("CREATE TABLE pair_12_6_es_term (p0 INTEGER, p1 INTEGER, param INTEGER)")
    conn.executemany("INSERT INTO pair_12_6_es_term VALUES (?, ?, ?)", [(0, 3, 0), (0, 2, 1)])
    if with_cell:
        conn.execute("CREATE TABLE global_cell (id INTEGER PRIMARY KEY, x FLOAT, y FLOAT, z FLOAT)")
        conn.executemany("INSERT INTO global_cell VALUES (?, ?, ?, ?)",
                         [(1, 30.0, 0, 0), (2, 0, 31.0, 0), (3, 0, 0, 32.0)])
    conn.commit()
    conn.close()
    return xyz, vel


def test_read_synthetic(tmp_path):
    p = str(tmp_path / "a.dms")
    xyz, vel = make_dms(p)
    with DesmondDMSFile(p) as d:
        assert d.getNumAtoms() == 6
        assert np.allclose(d.getPositions(), xyz * 0.1)
        assert np.allclose(d.getVelocities(), vel * 0.1)
        assert np.array_equal(d.getResidueIds(), [0, 0, 0, 1, 1, 1])
        assert np.allclose(d.getBox(), [3.0, 3.1, 3.2])
        sysd = d.createSystem(nonbondedMethod=S.CUTOFF_PERIODIC, nonbondedCutoff=0.9)
        pme = d.createSystem(nonbondedMethod=S.PME, nonbondedCutoff=0.9, ewaldErrorTolerance=1e-4)   # test_explicit.py:64
    assert pme.method == S.PME and abs(pme.ewald_alpha_effective() - np.sqrt(-np.log(2e-4)) / 0.9) < 1e-14
    assert sysd.method == S.CUTOFF_PERIODIC and sysd.cutoff == 0.9 and sysd.eps_rf == 78.3
    assert np.allclose(sysd.charge, 0.1 * (np.arange(6) - 2))
    assert np.allclose(sysd.sigma, [0.34, 0.25] * 3)
    assert np.allclose(sysd.epsilon, [0.1 * 4.184, 0.0] * 3)
    assert sysd.exclusions.tolist() == [[0, 1], [1, 2], [0, 2], [0, 3]]
    assert sysd.exception_pairs.tolist() == [[0, 3], [0, 2]]
    a, b = 1000.0 * 4.184 * 0.1 ** 12, 10.0 * 4.184 * 0.1 ** 6
    assert np.allclose(sysd.exception_params[0], [0.05, (a / b) ** (1 / 6), b * b / (4 * a)])
    assert np.allclose(sysd.exception_params[1], [-0.02, 1.0, 0.0])   # a == 0: eps 0, sigma 1


def test_two_files_are_concatenated(tmp_path):
    p1, p2 = str(tmp_path / "a.dms"), str(tmp_path / "b.dms")
    xyz1, _ = make_dms(p1, n=6, seed=1)
    xyz2, _ = make_dms(p2, n=4, seed=2)
    with DesmondDMSFile([p1, p2]) as d:
        assert d.getNumAtoms() == 10
        assert np.allclose(d.getPositions(), np.concatenate([xyz1, xyz2]) * 0.1)
        sysd = d.createSystem()
    assert sysd.n_atoms == 10
    assert [6, 7] in sysd.exclusions.tolist() and [6, 9] in sysd.exception_pairs.tolist()


def test_write_back_round_trip(tmp_path):
    p = str(tmp_path / "a.dms")
    make_dms(p)
    rng = np.random.default_rng(5)
    new_x, new_v = rng.uniform(0, 3, (6, 3)), rng.normal(size=(6, 3))
    with DesmondDMSFile(p) as d:
        assert d.setPositions(new_x) == 6
        assert d.setVelocities(new_v) == 6
        d.setGlobalCell([4.0, 0, 0], [0, 4.1, 0], [0, 0, 4.2])
    with DesmondDMSFile(p) as d:
        assert np.allclose(d.getPositions(), new_x, atol=1e-14)
        assert np.allclose(d.getVelocities(), new_v, atol=1e-14)
        assert np.allclose(d.getBox(), [4.0, 4.1, 4.2])


def test_errors(tmp_path):
    with pytest.raises(IOError):
        DesmondDMSFile(str(tmp_path / "missing.dms"))
    empty = str(tmp_path / "empty.dms")
    sqlite3.connect(empty).close()
    with pytest.raises(IOError):
        DesmondDMSFile(empty)
    p = str(tmp_path / "nonb.dms")
    make_dms(p, with_nbtype=False)
    with pytest.raises(ValueError):
        DesmondDMSFile(p)
    q = str(tmp_path / "nocell.dms")
    make_dms(q, with_cell=False)
    with DesmondDMSFile(q) as d:
        with pytest.raises(ValueError):
            d.createSystem(nonbondedMethod=S.CUTOFF_PERIODIC)
        with pytest.raises(ValueError):
            d.createSystem(nonbondedMethod=7)
        with pytest.raises(ValueError):
            d.setPositions(np.zeros((5, 3)))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference fixtures only exist in the build container")
@pytest.mark.parametrize("dms,npz", [("oa-g6-g3-align-restr_0_displaced.dms", "cfg1_oa_g6_g3.npz"),
                                     ("temoa-g1-g4.dms", "cfg2_temoa_g1_g4.npz")])
def test_shipped_fixtures_match_golden(dms, npz, tmp_path):
    import shutil
    local = str(tmp_path / dms)          # /root/reference is read-only: open a copy
    shutil.copy(os.path.join(REF, dms), local)
    z = np.load(os.path.join(S.GOLDEN_DIR, npz))
    with DesmondDMSFile(local) as d:
        sysd = d.createSystem()
        assert np.array_equal(d.getPositions(), z["positions"])
        assert np.array_equal(d.getVelocities(), z["velocities"])
        assert np.array_equal(d.getMasses(), z["masses"])
        assert np.array_equal(d.getResidueIds(), z["resid"])
        assert np.array_equal(d.getBox(), z["box"])
    for k in ("charge", "sigma", "epsilon", "exclusions", "exception_pairs", "exception_params"):
        assert np.array_equal(getattr(sysd, k), z[k]), k


def test_opls_option_selects_the_geometric_rule(tmp_path):
    """createSystem(OPLS=True): desmonddmsfile75.py:780-810 and :427-438."""
    p = str(tmp_path / "a.dms")
    make_dms(p)
    with DesmondDMSFile(p) as d:
        lb = d.createSystem(nonbondedMethod=S.CUTOFF_PERIODIC, nonbondedCutoff=0.9)
        geo = d.createSystem(nonbondedMethod=S.CUTOFF_PERIODIC, nonbondedCutoff=0.9, OPLS=True)
    assert not lb.lj_geometric and lb.use_dispersion_correction
    assert geo.lj_geometric and not geo.use_dispersion_correction
    assert np.array_equal(lb.sigma, geo.sigma) and np.array_equal(lb.epsilon, geo.epsilon)
    assert np.array_equal(lb.exclusions, geo.exclusions) and np.array_equal(lb.exception_params, geo.exception_params)
