"""GPU parity tests of the cluster-pair-list path (SDM_PAIR_CLUSTER) through the C ABI."""
import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import oracle as O
from test_gpu_parity import check_against_oracle, rms_rel, run_case, F_RMS_RTOL

pytestmark = pytest.mark.gpu
CL = _lib.PAIR_CLUSTER


def oracle_eval(case, positions=None, nthreads=None):
    return O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement,
                      case.positions if positions is None else positions,
                      nthreads=nthreads or O.max_threads())


@pytest.fixture(scope="module")
def cfg2():
    c = S.cfg2()
    return c, oracle_eval(c)


def test_cfg2_cluster(cfg2):
    case, ref = cfg2
    with run_case(case, CL) as ctx:
        sc = check_against_oracle(ctx, case, ref)
        assert sc["n_pairs1"] == 4197871
        assert sc["n_moved1"] == 15175 and sc["n_moved2"] == 15138
        assert ctx.info("n_list_builds") == 1


def test_cfg2_cluster_pair_set_bit_exact(cfg2):
    case, _ = cfg2
    ref = O.nonbonded(case.system, case.positions, want_pairs=True, nthreads=O.max_threads())
    with run_case(case, CL) as ctx:
        got = ctx.pairs(0)
    assert got.shape == ref["pairs"].shape
    assert np.array_equal(got, ref["pairs"])


def test_cfg1_cluster_nonperiodic():
    case = S.cfg1()
    ref = oracle_eval(case)
    with run_case(case, CL) as ctx:
        check_against_oracle(ctx, case, ref)
        assert np.array_equal(ctx.pairs(0), O.nonbonded(case.system, case.positions, want_pairs=True)["pairs"])


def test_cfg1_cluster_short_cutoff_nonperiodic():
    case = S.cfg1()
    case.system.cutoff = 1.2
    ref = oracle_eval(case)
    with run_case(case, CL) as ctx:
        check_against_oracle(ctx, case, ref)
        assert np.array_equal(ctx.pairs(0), O.nonbonded(case.system, case.positions, want_pairs=True)["pairs"])


@pytest.mark.parametrize("n,seed", [(3000, 5), (12000, 9)])
def test_synthetic_cluster(n, seed):
    case = S.synthetic_case(n_atoms=n, ligand_atoms=30, seed=seed, protein_atoms=n // 10,
                            displacement=(0.0, 0.0, 1.5))
    ref = oracle_eval(case)
    with run_case(case, CL) as ctx:
        check_against_oracle(ctx, case, ref)
        pr = O.nonbonded(case.system, case.positions, want_pairs=True, nthreads=O.max_threads())["pairs"]
        assert np.array_equal(ctx.pairs(0), pr)


def test_cluster_matches_allpairs_and_is_deterministic(cfg2):
    case, _ = cfg2
    with run_case(case, CL) as a, run_case(case, CL) as b, run_case(case, _lib.PAIR_ALLPAIRS) as c:
        fa, fb, fc = a.forces(0, _lib.FORCE_STATE1), b.forces(0, _lib.FORCE_STATE1), c.forces(0, _lib.FORCE_STATE1)
        assert np.array_equal(fa, fb)                       # fixed-point accumulation: bit-reproducible
        assert a.scalars(0) == b.scalars(0)
        assert rms_rel(fa, fc) < 5e-6
        # the moved-pair path is FP64 in both modes; only its (fixed) summation order differs
        assert abs(a.scalars(0)["u"] - c.scalars(0)["u"]) <= 1e-10 * max(1.0, abs(c.scalars(0)["u"]))


def test_list_reuse_refresh_and_rebuild(cfg2):
    """MD-like use: positions drift a little every eval; the list is reused (refresh kernel keeps
    the build-time periodic image) and rebuilt every nstlist evals."""
    case, _ = cfg2
    rng = np.random.default_rng(42)
    ctx = SDMContext(case.system, case.displacement, pair_mode=CL, nstlist=4, skin=0.08)
    ctx.set_alchemical(0, case.alch)
    pos = case.positions.copy()
    for step in range(6):
        ctx.set_positions(0, pos)
        ctx.eval()
        sc = ctx.scalars(0)
        assert sc["status"] == 0, (step, sc)
        if step in (0, 3, 5):
            ref = oracle_eval(case, pos)
            check_against_oracle(ctx, case, ref)
        pos = pos + rng.normal(scale=0.002, size=pos.shape)
    assert ctx.info("n_list_builds") == 2          # evals 0 and 4
    assert ctx.scalars(0)["list_age"] == 2
    ctx.close()


def test_refresh_across_periodic_boundary():
    """An atom that crosses the box face between builds must keep its build-time image."""
    case = S.synthetic_case(6000, 30, seed=3, protein_atoms=300, displacement=(0.0, 0.0, 1.5))
    L = case.system.box[0]
    pos = case.positions.copy()
    i = int(np.argmin(pos[:, 0]))            # closest to the x = 0 face
    with SDMContext(case.system, case.displacement, pair_mode=CL, nstlist=10, skin=0.1) as ctx:
        ctx.set_alchemical(0, case.alch)
        ctx.set_positions(0, pos)
        ctx.eval()
        mol = (i // 3) * 3
        pos[mol:mol + 3, 0] -= pos[i, 0] + 0.005  # now at x = -0.005 (outside the box), moved < skin/2
        assert abs(pos[i, 0] + 0.005) < 1e-12 and (case.positions[i, 0] + 0.005) < 0.05
        ctx.set_positions(0, pos)
        ctx.eval()
        assert ctx.info("n_list_builds") == 1
        ref = oracle_eval(case, pos)
        check_against_oracle(ctx, case, ref)


def test_stale_list_is_reported_then_cured_by_rebuild(cfg2):
    case, _ = cfg2
    with SDMContext(case.system, case.displacement, pair_mode=CL, nstlist=100, skin=0.06) as ctx:
        ctx.set_alchemical(0, case.alch)
        ctx.set_positions(0, case.positions)
        ctx.eval()
        assert ctx.scalars(0)["status"] == 0
        pos = case.positions.copy()
        pos[5000] += [0.05, 0.0, 0.0]            # > skin/2
        ctx.set_positions(0, pos)
        ctx.eval()
        assert ctx.scalars(0)["status"] == _lib.SDM_ERR_STALE_LIST
        ctx.invalidate_list()
        ctx.eval()
        assert ctx.scalars(0)["status"] == 0
        check_against_oracle(ctx, case, oracle_eval(case, pos))


def test_replica_batch_cluster(cfg2):
    """Replicas share one launch; each matches its own oracle run and does not depend on what else
    is resident (bit-identical to a single-replica context)."""
    case, ref = cfg2
    rng = np.random.default_rng(8)
    pos1 = case.positions + rng.normal(scale=0.01, size=case.positions.shape)
    al1 = S.AlchemicalState(**vars(case.alch))
    al1.lambda1, al1.lambda2, al1.alpha, al1.u0 = 0.1, 0.4, 0.05, -3.0
    ctx = SDMContext(case.system, case.displacement, n_replicas=3, pair_mode=CL)
    for r, (p, al) in enumerate([(case.positions, case.alch), (pos1, al1), (case.positions, case.alch)]):
        ctx.set_positions(r, p)
        ctx.set_alchemical(r, al)
    ctx.eval()
    check_against_oracle(ctx, case, ref, replica=0)
    check_against_oracle(ctx, case, ref, replica=2)
    assert np.array_equal(ctx.forces(0), ctx.forces(2))
    c1 = S.SDMCase(case.name, case.system, pos1, case.displacement, al1)
    ref1 = oracle_eval(c1)
    check_against_oracle(ctx, c1, ref1, replica=1)
    with run_case(c1, CL) as single:
        assert np.array_equal(single.forces(0), ctx.forces(1))
        s1, sb = single.scalars(0), ctx.scalars(1)
        assert s1["E1"] == sb["E1"] and s1["u"] == sb["u"] and s1["sp"] == sb["sp"]
    ctx.close()


def test_auto_mode_picks_cluster_for_large_and_allpairs_for_small(cfg2):
    case, _ = cfg2
    with SDMContext(case.system, case.displacement) as ctx:
        assert ctx.info("pair_mode") == CL
    with SDMContext(S.cfg1().system, None) as ctx:
        assert ctx.info("pair_mode") == _lib.PAIR_ALLPAIRS


def _drift_run(case, env_async, monkeypatch, n_evals=9, nstlist=3):
    """Forces / scalars of n_evals evaluations under drifting positions, list rebuilt every nstlist."""
    monkeypatch.setenv("SDMB200_ASYNC_BUILD", env_async)
    rng = np.random.default_rng(77)
    out = []
    with SDMContext(case.system, case.displacement, pair_mode=CL, nstlist=nstlist, skin=0.08) as ctx:
        ctx.set_alchemical(0, case.alch)
        pos = case.positions.copy()
        for _ in range(n_evals):
            ctx.set_positions(0, pos)
            ctx.eval()
            sc = ctx.scalars(0)
            assert sc["status"] == 0, sc
            out.append((ctx.forces(0).copy(), sc["E1"], sc["u"], sc["n_pairs1"]))
            pos = pos + rng.normal(scale=0.002, size=pos.shape)
        counts = (ctx.info("n_list_builds"), ctx.info("n_async_builds"), ctx.info("n_sync_builds"))
    return out, counts, pos


def test_asynchronous_list_builds_match_synchronous_ones(cfg2, monkeypatch):
    """Builds after the first are enqueued without host synchronisation (counts stay on the device, grids
    sized with bounds from the previous build): same bits as builds that size every stage on the host."""
    case, _ = cfg2
    a, ca, _ = _drift_run(case, "1", monkeypatch)
    b, cb, _ = _drift_run(case, "0", monkeypatch)
    assert ca == (3, 2, 1) and cb == (3, 0, 3), (ca, cb)
    for (fa, e1a, ua, na), (fb, e1b, ub, nb) in zip(a, b):
        assert na == nb and e1a == e1b and ua == ub
        assert np.array_equal(fa, fb)


def test_asynchronous_build_that_outgrows_its_bounds_is_reported_and_repaired(monkeypatch):
    """A frame much denser than the one the bounds came from: the build notices on the device, every replica
    reports SDM_ERR_CAPACITY, and the repeated evaluation (sized on the host) is correct."""
    monkeypatch.setenv("SDMB200_ASYNC_BUILD", "1")
    case = S.synthetic_case(6000, 30, seed=11, protein_atoms=300, displacement=(0.0, 0.0, 1.5))
    dense = case.positions.copy()
    L = case.system.box
    dense[:, 0] = dense[:, 0] % L[0] * 0.5            # everything squeezed into half of the box
    with SDMContext(case.system, case.displacement, pair_mode=CL, nstlist=1, skin=0.08) as ctx:
        ctx.set_alchemical(0, case.alch)
        ctx.set_positions(0, case.positions)
        ctx.eval()
        assert ctx.scalars(0)["status"] == 0
        ctx.eval()                                     # an asynchronous build with the first frame's counts
        assert ctx.scalars(0)["status"] == 0 and ctx.info("n_async_builds") == 1
        ctx.set_positions(0, dense)
        ctx.eval()
        first = ctx.scalars(0)["status"]
        tries = 0
        while ctx.scalars(0)["status"] == _lib.SDM_ERR_CAPACITY and tries < 4:
            ctx.eval()
            tries += 1
        sc = ctx.scalars(0)
        assert sc["status"] == 0
        assert first == _lib.SDM_ERR_CAPACITY, "the scenario did not outgrow the bounds"
        ref = oracle_eval(case, dense)
        assert sc["n_pairs1"] == ref["n_pairs1"]   # (energies of a squeezed box overflow FP32: not compared)


def test_in_block_kd_refinement_gives_the_order_of_the_sort_based_rounds(cfg2, monkeypatch):
    """The two kd refinement rounds of the list build run as one kernel (ranks counted inside each cell);
    the development knob SDMB200_KD_SORTS=1 runs them as two global radix sorts: same slots, same bits."""
    case, _ = cfg2
    monkeypatch.setenv("SDMB200_KD_SORTS", "0")
    a, _, _ = _drift_run(case, "1", monkeypatch, n_evals=4, nstlist=2)
    monkeypatch.setenv("SDMB200_KD_SORTS", "1")
    b, _, _ = _drift_run(case, "1", monkeypatch, n_evals=4, nstlist=2)
    for (fa, e1a, ua, na), (fb, e1b, ub, nb) in zip(a, b):
        assert na == nb and e1a == e1b and ua == ub
        assert np.array_equal(fa, fb)
