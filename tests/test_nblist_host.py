"""CPU check of the cluster-pair list logic: the per-item bodies of
openmm_sdm_plugin_b200/csrc/nblist_core.h (the same code the device kernels call) are compiled
with g++ and walked like the pair kernel walks them; every in-cutoff non-excluded pair of the
oracle must be covered exactly once (bit-exact pair sets)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import system as S
from oracle import oracle as O

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck")


@pytest.fixture(scope="module")
def chk():
    subprocess.check_call(["make", "-C", HERE, "-s"])
    L = C.CDLL(os.path.join(HERE, "libnblcheck.so"))
    L.hostcheck_pairs.restype = C.c_longlong
    L.hostcheck_pairs.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double,
                                  C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                  C.c_longlong, C.c_void_p]
    return L


def covered_pairs(L, case, R=1, skin=0.06, replica=0, chunk=32, jitter=None):
    s = case.system
    ex = np.unique(np.sort(s.exclusions, axis=1), axis=0).astype(np.int32)
    ex = np.ascontiguousarray(ex[ex[:, 0] != ex[:, 1]])
    pos = np.tile(case.positions, (R, 1, 1))
    if jitter is not None:
        pos = pos + jitter
    pos = np.ascontiguousarray(pos)
    box = np.ascontiguousarray(s.box, dtype=np.float64)
    ref = O.nonbonded(s, pos[replica], want_pairs=True, nthreads=O.max_threads())["pairs"]
    out = np.zeros((len(ref) + 4096, 2), np.int32)
    st = np.zeros(10)
    n = L.hostcheck_pairs(s.n_atoms, R, pos.ctypes.data, int(s.method == S.CUTOFF_PERIODIC),
                          box.ctypes.data, s.cutoff, skin, len(ex), ex.ctypes.data, chunk, replica,
                          out.ctypes.data, len(out), st.ctypes.data)
    assert n >= 0
    return out[:n], ref, st


def test_cfg1_nonperiodic_all_pairs_covered_once(chk):
    got, ref, st = covered_pairs(chk, S.cfg1())
    assert len(got) == 25061 and np.array_equal(got, ref)


def test_cfg1_short_cutoff(chk):
    case = S.cfg1()
    case.system.cutoff = 1.2
    got, ref, _ = covered_pairs(chk, case)
    assert np.array_equal(got, ref)


def test_small_periodic_box_with_wrapping_stencil(chk):
    """3.1 nm box: 4 cells per dimension but a 5-cell stencil, so cells are visited under two
    periodic images; each pair must still be counted once."""
    case = S.synthetic_case(3000, 30, seed=5, protein_atoms=300, displacement=(0, 0, 1.5))
    got, ref, st = covered_pairs(chk, case)
    assert st[7] > st[6] ** (1 / 3) - 1e-9
    assert np.array_equal(got, ref)


def test_positions_outside_the_box_are_wrapped(chk):
    case = S.synthetic_case(3000, 30, seed=11, protein_atoms=0)
    rng = np.random.default_rng(3)
    shift = rng.integers(-2, 3, size=(case.system.n_atoms // 3, 1, 3)) * case.system.box
    case.positions = (case.positions.reshape(-1, 3, 3) + shift).reshape(-1, 3)  # whole molecules
    got, ref, _ = covered_pairs(chk, case)
    assert np.array_equal(got, ref)


def test_cfg2_two_replicas(chk):
    case = S.cfg2()
    rng = np.random.default_rng(1)
    jit = rng.normal(scale=0.01, size=(2, case.system.n_atoms, 3))
    got, ref, st = covered_pairs(chk, case, R=2, replica=1, jitter=jit)
    assert len(ref) > 4_000_000
    assert np.array_equal(got, ref)
    # list statistics that DESIGN.md quotes: half of the lane pairs the row kernel evaluates are
    # inside the cutoff (35 % with the 8 x 8 tiles of the cluster-pair entries themselves), about
    # one row entry in twelve carries an allow word
    assert 0.48 < len(ref) * 2 / st[5] < 0.56
    assert 45.0 < st[8] / (2 * case.system.n_atoms) < 55.0
    assert st[9] / st[8] < 0.12


@pytest.mark.parametrize("skin", [0.0, 0.12])
def test_skin_does_not_change_the_in_cutoff_set(chk, skin):
    case = S.synthetic_case(6000, 60, seed=2, protein_atoms=600)
    got, ref, _ = covered_pairs(chk, case, skin=skin)
    assert np.array_equal(got, ref)


# ---- column layout (xy columns cut along z into 64-atom chunk cells): shapes that stress it -------

def test_columns_overflow_their_chunk_cells(chk):
    """All atoms squeezed into a quarter of the xy plane: those columns hold four times the
    average, more than the kz chunk cells a column has, so their last cell keeps the rest (several
    superclusters in one cell) while most columns are empty."""
    case = S.synthetic_case(6000, 30, seed=21, protein_atoms=0)
    case.positions = np.mod(case.positions, case.system.box)
    case.positions[:, :2] *= 0.5
    got, ref, st = covered_pairs(chk, case)
    assert st[1] > 6000 / 64            # more superclusters than chunk cells would give
    assert np.array_equal(got, ref)


def test_tall_and_flat_boxes(chk):
    """Same atoms in a box stretched along z (many chunks per column, the three z images are all
    visited) and in one squeezed to just over two list radii along z (every chunk sees its own
    periodic images)."""
    base = S.synthetic_case(4200, 30, seed=22, protein_atoms=0)
    L = base.system.box[0]
    for scale in ((0.8, 0.8, 1.0 / 0.64), (1.35, 1.35, 2.2 / L)):
        case = S.synthetic_case(4200, 30, seed=22, protein_atoms=0)
        sc = np.array(scale)
        case.system.box = case.system.box * sc
        case.positions = np.mod(case.positions, base.system.box) * sc
        assert np.all(case.system.box >= 2.0 * (case.system.cutoff + 0.06))
        got, ref, _ = covered_pairs(chk, case)
        assert np.array_equal(got, ref)


def test_rows_of_two_clusters_cover_the_same_pairs(chk, monkeypatch):
    """i-groups of two clusters (16 i-atoms per row, SDMB200_ROW_GROUP=2): the allow words switch
    off the half of a group that does not own a j-cluster of its own supercluster."""
    monkeypatch.setenv("SDMB200_ROW_GROUP", "2")
    for case in (S.cfg1(), S.synthetic_case(3000, 30, seed=5, protein_atoms=300, displacement=(0, 0, 1.5))):
        got, ref, st = covered_pairs(chk, case)
        assert np.array_equal(got, ref)


def test_both_layouts_cover_the_same_pairs(chk, monkeypatch):
    case = S.synthetic_case(3000, 30, seed=23, protein_atoms=300)
    got_c, ref, st_c = covered_pairs(chk, case)
    monkeypatch.setenv("SDMB200_LAYOUT", "cells")
    got_g, _, st_g = covered_pairs(chk, case)
    assert np.array_equal(got_c, ref) and np.array_equal(got_g, ref)
    assert st_c[0] < st_g[0]            # fewer padded slots with columns
