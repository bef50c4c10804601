"""Edge cases and size-independent properties of the GPU path (through the C ABI): empty and
ragged inputs, several displacement groups, stale lists, multi-replica batches, and -- at sizes
the oracle cannot reach in seconds -- Newton's third law, batch invariance and finite
differences."""
import copy

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import oracle as O
from test_gpu_parity import check_against_oracle, rms_rel, run_case, F_RMS_RTOL

pytestmark = pytest.mark.gpu
CL, AP = _lib.PAIR_CLUSTER, _lib.PAIR_ALLPAIRS


def oracle_eval(case, positions=None):
    return O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement,
                      case.positions if positions is None else positions, nthreads=O.max_threads())


@pytest.mark.parametrize("mode", [AP, CL])
def test_no_displaced_atoms(mode):
    """All-zero displacement map: no moved pairs at all (n_lig = 0), u = 0, F = F1."""
    case = S.synthetic_case(4500, 30, seed=21, protein_atoms=300)
    case.displacement[:] = 0.0
    ref = oracle_eval(case)
    with run_case(case, mode) as ctx:
        sc = check_against_oracle(ctx, case, ref)
        assert sc["u"] == 0.0 and sc["n_moved1"] == 0 and sc["n_moved2"] == 0
        assert np.array_equal(ctx.forces(0, _lib.FORCE_HYBRID), ctx.forces(0, _lib.FORCE_STATE1))


@pytest.mark.parametrize("mode", [AP, CL])
def test_every_atom_displaced_by_the_same_vector(mode):
    """A rigid shift of the whole system is one displacement group: nothing changes."""
    case = S.synthetic_case(3000, 30, seed=22, protein_atoms=0)
    case.displacement[:] = (0.3, -0.2, 0.1)
    with run_case(case, mode) as ctx:
        sc = ctx.scalars(0)
        assert sc["status"] == 0 and sc["u"] == 0.0 and sc["n_moved1"] == 0
        assert np.abs(ctx.forces(0, _lib.FORCE_DELTA)).max() == 0.0


@pytest.mark.parametrize("mode", [AP, CL])
def test_three_displacement_groups_and_displaced_displaced_pairs(mode):
    """Two ligands with opposite displacements plus a third group: pairs between differently
    displaced atoms change too and are seen from both blocks (weight 1/2 each)."""
    case = S.synthetic_case(6000, 60, seed=23, protein_atoms=300, displacement=(0.0, 0.0, 1.2))
    lig = np.flatnonzero(np.abs(case.displacement).sum(1) > 0)
    third = len(lig) // 3
    case.displacement[lig[:third]] *= -1.0                        # group 2: the opposite way
    case.displacement[lig[third:2 * third]] = (0.4, 0.0, 0.0)     # group 3
    ref = oracle_eval(case)
    with run_case(case, mode) as ctx:
        sc = check_against_oracle(ctx, case, ref)
        assert sc["n_moved1"] > 0 and sc["n_moved2"] > 0


def test_tiny_ragged_system_on_the_cluster_path():
    """11 atoms (not a multiple of the cluster size), non-periodic cutoff."""
    rng = np.random.default_rng(5)
    n = 11
    pos = rng.uniform(0.0, 1.2, size=(n, 3))
    sysd = S.NonbondedSystem(rng.uniform(-0.5, 0.5, n), rng.uniform(0.2, 0.35, n), rng.uniform(0.1, 0.8, n),
                             np.array([[0, 1], [1, 2], [5, 6]], np.int32), np.zeros((0, 2), np.int32),
                             np.zeros((0, 3)), method=S.CUTOFF_NONPERIODIC, cutoff=0.9, eps_rf=78.3,
                             box=np.zeros(3), use_dispersion_correction=False)
    disp = np.zeros((n, 3))
    disp[8:] = (0.5, 0.1, -0.2)
    case = S.SDMCase("tiny", sysd, pos, disp, S.AlchemicalState(lambdac=0.4))
    ref = oracle_eval(case)
    for mode in (AP, CL):
        with run_case(case, mode) as ctx:
            check_against_oracle(ctx, case, ref)
            assert np.array_equal(ctx.pairs(0), O.nonbonded(sysd, pos, want_pairs=True)["pairs"])


def test_positions_many_boxes_away_and_displacement_longer_than_the_box():
    case = S.synthetic_case(4500, 30, seed=24, protein_atoms=300)
    L = case.system.box[0]
    lig = np.abs(case.displacement).sum(1) > 0
    case.displacement[lig] += (2 * L, -L, 0.0)          # same state 2 modulo the box
    shift = np.random.default_rng(1).integers(-3, 4, size=(case.system.n_atoms // 3, 1, 3)) * case.system.box
    case.positions = (case.positions.reshape(-1, 3, 3) + shift).reshape(-1, 3)
    ref = oracle_eval(case)
    with run_case(case, CL) as ctx:
        check_against_oracle(ctx, case, ref)


def test_stale_list_is_reported_not_ignored():
    case = S.synthetic_case(4500, 30, seed=25, protein_atoms=300)
    with SDMContext(case.system, case.displacement, pair_mode=CL, nstlist=50, skin=0.06) as ctx:
        ctx.set_alchemical(0, case.alch)
        ctx.set_positions(0, case.positions)
        ctx.eval()
        assert ctx.scalars(0)["status"] == 0
        pos = case.positions.copy()
        pos[300:303] += 0.05                             # one molecule moves more than skin/2
        ctx.set_positions(0, pos)
        ctx.eval()
        assert ctx.scalars(0)["status"] == _lib.SDM_ERR_STALE_LIST
        ctx.eval()                                       # reading the status scheduled a rebuild
        sc = ctx.scalars(0)
        assert sc["status"] == 0 and sc["list_age"] == 1
        ref = oracle_eval(case, pos)
        check_against_oracle(ctx, case, ref)


def test_replica_batch_with_different_positions_and_states():
    """Three replicas, three coordinate sets, three lambda states in one launch sequence: each
    must equal its own single-replica evaluation bit for bit (fixed-point accumulation)."""
    case = S.synthetic_case(6000, 30, seed=26, protein_atoms=300)
    rng = np.random.default_rng(3)
    states = S.atm_lambda_schedule(6)
    pos = [case.positions + rng.normal(scale=0.01, size=case.positions.shape) for _ in range(3)]
    singles = []
    for r in range(3):
        c1 = copy.copy(case)
        c1.positions, c1.alch = pos[r], states[2 * r]
        with run_case(c1, CL) as ctx:
            singles.append((ctx.scalars(0), ctx.forces(0)))
    with SDMContext(case.system, case.displacement, n_replicas=3, pair_mode=CL) as ctx:
        for r in range(3):
            ctx.set_positions(r, pos[r])
            ctx.set_alchemical(r, states[2 * r])
        ctx.eval()
        for r in range(3):
            sc, f = ctx.scalars(r), ctx.forces(r)
            for k in ("E1_pair", "u", "u_sc", "sp", "pot_energy", "n_pairs1", "n_moved1", "n_moved2"):
                assert sc[k] == singles[r][0][k], (r, k)
            assert np.array_equal(f, singles[r][1])
    # and one of them against the oracle
    c1 = copy.copy(case)
    c1.positions, c1.alch = pos[1], states[2]
    ref = oracle_eval(c1)
    assert abs(singles[1][0]["u"] - ref["u"]) <= 1e-6 * max(1.0, abs(ref["u"]))


def test_full_size_properties_100k_atoms():
    """cfg5-size box (100 k atoms): the oracle would take minutes, so check properties instead:
    total force ~ 0 (Newton's third law), state-1 forces independent of the displacement map and
    of the batch size, u independent of replica index."""
    case = S.synthetic_case(100_000, 60, seed=27)
    n = case.system.n_atoms
    with SDMContext(case.system, case.displacement, n_replicas=2, pair_mode=CL) as ctx:
        for r in range(2):
            ctx.set_positions(r, case.positions)
            ctx.set_alchemical(r, case.alch)
        ctx.eval()
        s0, s1 = ctx.scalars(0), ctx.scalars(1)
        f1 = ctx.forces(0, _lib.FORCE_STATE1)
        df = ctx.forces(0, _lib.FORCE_DELTA)
        assert s0["status"] == 0 and s0["n_pairs1"] > 200 * n * 0.9
        assert s0["u"] == s1["u"] and s0["E1_pair"] == s1["E1_pair"]
        assert np.array_equal(f1, ctx.forces(1, _lib.FORCE_STATE1))
        scale = np.abs(f1).sum(0)
        assert np.all(np.abs(f1.sum(0)) <= 1e-6 * scale), (f1.sum(0), scale)
        assert np.all(np.abs(df.sum(0)) <= 1e-9 * max(1.0, np.abs(df).sum()))
    zero = copy.copy(case)
    zero.displacement = np.zeros_like(case.displacement)
    with run_case(zero, CL) as ctx:
        assert np.array_equal(ctx.forces(0, _lib.FORCE_STATE1), f1)


def test_forces_are_the_gradient_of_the_energies():
    """Finite differences on a small box: F1 = -dE1/dx and (F2 - F1) = -du/dx, which ties the
    force kernels to the energy kernels without the oracle."""
    case = S.synthetic_case(3000, 30, seed=28, protein_atoms=300, displacement=(0.0, 0.0, 1.5))
    lig = np.flatnonzero(np.abs(case.displacement).sum(1) > 0)
    probe_atoms = [int(lig[0]), int(lig[7]), 5, 1501]
    h = 1e-4
    with SDMContext(case.system, case.displacement, pair_mode=AP) as ctx:
        ctx.set_alchemical(0, case.alch)

        def energies(p):
            ctx.set_positions(0, p)
            ctx.eval()
            sc = ctx.scalars(0)
            return sc["E1"], sc["u"]

        ctx.set_positions(0, case.positions)
        ctx.eval()
        f1 = ctx.forces(0, _lib.FORCE_STATE1)
        df = ctx.forces(0, _lib.FORCE_DELTA)
        for a in probe_atoms:
            for d in range(3):
                p = case.positions.copy(); p[a, d] += h
                ep, up = energies(p)
                p[a, d] -= 2 * h
                em, um = energies(p)
                # E1 is an FP32 sum of ~6e5 terms: its finite difference is only good to ~1e-2 abs
                assert abs(-(ep - em) / (2 * h) - f1[a, d]) <= 0.05 * max(1.0, abs(f1[a, d])) + 60.0
                # u and dF are FP64 over moved pairs: limited by the O(h^2) truncation of the central
                # difference (u ~ 1e6 kJ/mol here: the displaced ligand overlaps solvent)
                assert abs(-(up - um) / (2 * h) - df[a, d]) <= 1e-3 * max(1.0, abs(df[a, d])), (a, d)


def test_scratch_capacity_overflow_is_reported_and_repaired():
    """A displaced atom with more neighbours than the per-hit scratch was sized for (here: an
    absurdly dense blob) must not corrupt anything: the evaluation reports SDM_ERR_CAPACITY, the
    context grows the scratch, and the repeated evaluation is correct."""
    rng = np.random.default_rng(9)
    n = 4000
    # 4000 atoms inside a 1.2 nm ball: ~550 atoms / nm^3, far above the 200 / nm^3 the scratch assumes
    v = rng.normal(size=(n, 3))
    pos = 3.0 + 0.6 * v / np.linalg.norm(v, axis=1, keepdims=True) * rng.uniform(0, 1, (n, 1)) ** (1 / 3)
    sysd = S.NonbondedSystem(np.zeros(n), np.full(n, 0.01), np.full(n, 1e-6), np.zeros((0, 2), np.int32),
                             np.zeros((0, 2), np.int32), np.zeros((0, 3)), method=S.CUTOFF_NONPERIODIC,
                             cutoff=1.0, eps_rf=78.3, box=np.zeros(3), use_dispersion_correction=False)
    disp = np.zeros((n, 3))
    disp[:4] = (0.2, 0.0, 0.0)
    case = S.SDMCase("dense", sysd, pos, disp, S.AlchemicalState(lambdac=0.5))
    ref = oracle_eval(case)
    with SDMContext(sysd, disp, pair_mode=AP) as ctx:
        ctx.set_alchemical(0, case.alch)
        ctx.set_positions(0, pos)
        ctx.eval()
        first = ctx.scalars(0)["status"]
        assert first in (0, _lib.SDM_ERR_CAPACITY)
        tries = 0
        while ctx.scalars(0)["status"] == _lib.SDM_ERR_CAPACITY and tries < 6:
            ctx.eval()
            tries += 1
        sc = ctx.scalars(0)
        assert sc["status"] == 0
        assert abs(sc["u"] - ref["u"]) <= 1e-6 * max(1.0, abs(ref["u"]))
        assert first == _lib.SDM_ERR_CAPACITY   # the scenario really overflowed the first time


def test_column_layout_with_overfull_columns():
    """The list's column layout under stress: a liquid slab that fills only a quarter of the xy plane
    of its (doubled) box, so the occupied columns hold four times the average -- more than their
    chunk cells, the last cell of a column keeps the rest -- and most columns are empty.  Pair set
    bit-exact, forces within the parity bar, and the geometric 3-D cells (SDMB200_LAYOUT=cells)
    give the same pair set."""
    import os
    case = S.synthetic_case(6000, 30, seed=31, protein_atoms=0)
    case.positions = np.mod(case.positions, case.system.box)
    case.system.box = case.system.box * np.array([2.0, 2.0, 1.0])
    ref = oracle_eval(case)
    want = O.nonbonded(case.system, case.positions, want_pairs=True, nthreads=O.max_threads())["pairs"]
    with run_case(case, CL) as ctx:
        assert ctx.info("layout_columns") == 1
        assert ctx.info("n_sci") > ctx.info("n_cells") * 0.2
        check_against_oracle(ctx, case, ref)
        assert np.array_equal(ctx.pairs(0), want)
    os.environ["SDMB200_LAYOUT"] = "cells"
    try:
        with run_case(case, CL) as ctx:
            assert ctx.info("layout_columns") == 0
            assert np.array_equal(ctx.pairs(0), want)
    finally:
        del os.environ["SDMB200_LAYOUT"]


def test_single_precision_transfers():
    """sdm_set_positions_all_f32 / sdm_enqueue_results_f32: positions in and forces out cross PCIe as
    float32 (the reference's OpenCL path holds them as float4); the device arithmetic is unchanged, so
    the result is the FP64-transfer result of the float-rounded coordinates, narrowed."""
    from openmm_sdm_plugin_b200.context import PinnedArray
    case = S.cfg2()
    n, R = case.system.n_atoms, 2
    rng = np.random.default_rng(3)
    x64 = np.stack([case.positions, case.positions + rng.normal(scale=0.003, size=(n, 3))])
    x32 = PinnedArray((R, n, 3), np.float32)
    f32 = PinnedArray((R, n, 3), np.float32)
    x32.array[...] = x64
    with SDMContext(case.system, case.displacement, n_replicas=R, pair_mode=_lib.PAIR_CLUSTER) as ctx:
        for r in range(R):
            ctx.set_alchemical(r, case.alch)
        ctx.set_positions_all(np.ascontiguousarray(x32.array.astype(np.float64)))   # the rounded coordinates, FP64 path
        ctx.eval()
        f64 = np.empty((R, n, 3))
        sc64 = ctx.read_results(f64)
        ctx.invalidate_list()
        ctx.set_positions_all(x32.array)
        ctx.eval()
        ctx.enqueue_results(f32.array)
        ctx.synchronize()
        sc32 = ctx.collect_scalars()
        for r in range(R):
            assert sc32[r]["status"] == 0 and sc32[r]["n_pairs1"] == sc64[r]["n_pairs1"]
            assert sc32[r]["u"] == sc64[r]["u"] and sc32[r]["E1"] == sc64[r]["E1"]
        assert np.array_equal(f32.array, f64.astype(np.float32))
