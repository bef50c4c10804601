"""GPU tests of NonbondedForce::Ewald / ::PME direct space (SDM_EWALD / SDM_PME) and of the
external dual-state hook that carries the reciprocal-space part (sdm_set_external_dual), through the C ABI,
against the oracle's restatement of OpenMM 7.3 ReferenceLJCoulombIxn::calculateEwaldIxn (includeDirect).
Same tolerances as the reaction-field parity tests (test_gpu_parity.py)."""
import copy

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import oracle as O
from test_gpu_parity import check_against_oracle, run_case

pytestmark = pytest.mark.gpu


def as_pme(case, method=S.PME):
    c = copy.copy(case)
    c.system = copy.copy(case.system)
    c.system.method = method
    return c


def oracle_eval(case):
    return O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement, case.positions,
                      nthreads=O.max_threads())


@pytest.fixture(scope="module")
def cfg2_pme():
    """example/test_explicit.py:64 as shipped: nonbondedMethod=PME, 1 nm cutoff (alpha from the default 5e-4)."""
    c = as_pme(S.cfg2())
    return c, oracle_eval(c)


def test_cfg2_pme_direct_space_cluster_path(cfg2_pme):
    case, ref = cfg2_pme
    assert ref["E1_exc"] > 1e5           # -qq erf(alpha r)/r of 20 k intramolecular water pairs (O-H: qq < 0) dominates it
    with run_case(case, _lib.PAIR_CLUSTER) as ctx:
        sc = check_against_oracle(ctx, case, ref)
        assert sc["n_pairs1"] == 4197871         # same cutoff, same exclusions as the reaction-field run
        pairs = ctx.pairs(0)
    ref_pairs = O.nonbonded(case.system, case.positions, want_pairs=True, nthreads=O.max_threads())["pairs"]
    assert np.array_equal(pairs, ref_pairs)


def test_cfg2_pme_differs_from_reaction_field_as_it_should(cfg2_pme):
    case, ref = cfg2_pme
    rf = oracle_eval(S.cfg2())
    assert abs(ref["E1_pair"] - rf["E1_pair"]) > 1e3 and abs(ref["u"] - rf["u"]) > 1e-3


@pytest.mark.parametrize("method", [S.EWALD, S.PME])
def test_synthetic_box_pme_both_pair_paths(method):
    case = as_pme(S.synthetic_case(4500, 30, seed=21, protein_atoms=300, displacement=(0.0, 0.0, 1.5)), method)
    case.system.ewald_alpha = 3.1            # an explicit splitting parameter instead of the tolerance rule
    ref = oracle_eval(case)
    for mode in (_lib.PAIR_ALLPAIRS, _lib.PAIR_CLUSTER):
        with run_case(case, mode) as ctx:
            check_against_oracle(ctx, case, ref)


def test_replica_batch_and_list_reuse_with_pme(cfg2_pme):
    case, ref = cfg2_pme
    rng = np.random.default_rng(3)
    with SDMContext(case.system, case.displacement, n_replicas=2, pair_mode=_lib.PAIR_CLUSTER, nstlist=5, skin=0.08) as ctx:
        pos1 = case.positions + rng.normal(scale=0.003, size=case.positions.shape)
        for r, p in enumerate((case.positions, pos1)):
            ctx.set_alchemical(r, case.alch)
            ctx.set_positions(r, p)
        for _ in range(3):                       # build, capture, replay
            ctx.eval()
        check_against_oracle(ctx, case, ref, replica=0)
        c1 = copy.copy(case)
        c1.positions = pos1
        check_against_oracle(ctx, c1, oracle_eval(c1), replica=1)


def test_external_dual_state_terms_enter_like_nonbonded_ones():
    """What carries the reciprocal-space part of PME: E1 += e1, u += e2 - e1, F1 += f1, F2 - F1 += f2 - f1, and
    with them u_sc, sp and the hybrid force; removed again by passing no forces."""
    case = S.cfg1()
    rng = np.random.default_rng(12)
    n = case.system.n_atoms
    f1e, f2e = rng.normal(scale=30.0, size=(n, 3)), rng.normal(scale=30.0, size=(n, 3))
    e1e, e2e = -1234.5, -1229.25
    with run_case(case, _lib.PAIR_ALLPAIRS, replicas=2) as ctx:
        base = [(ctx.scalars(r), ctx.forces(r, _lib.FORCE_STATE1).copy(), ctx.forces(r, _lib.FORCE_DELTA).copy()) for r in range(2)]
        ctx.set_external_dual(1, f1e, f2e, e1e, e2e)
        ctx.eval()
        s0, s1 = ctx.scalars(0), ctx.scalars(1)
        assert s0["E1"] == base[0][0]["E1"] and s0["u"] == base[0][0]["u"]            # replica 0 untouched
        assert np.array_equal(ctx.forces(0), ctx.forces(0)) and np.array_equal(ctx.forces(0, _lib.FORCE_STATE1), base[0][1])
        assert abs(s1["E1"] - (base[1][0]["E1"] + e1e)) <= 1e-12 * abs(s1["E1"])
        assert abs(s1["u"] - (base[1][0]["u"] + e2e - e1e)) <= 1e-10
        f1 = ctx.forces(1, _lib.FORCE_STATE1)
        df = ctx.forces(1, _lib.FORCE_DELTA)
        assert np.abs(f1 - (base[1][1] + f1e)).max() <= 1e-9 * np.abs(f1).max()
        assert np.abs(df - (base[1][2] + f2e - f1e)).max() <= 1e-9 * max(np.abs(df).max(), 1.0)
        al = case.alch
        usc, fp = O.softcore(al.softcore_method, s1["u"], al.umax, al.acore, al.ubcore)
        _, bfp = O.bias(S.AlchemicalState(**vars(al)), usc)
        assert abs(s1["sp"] - bfp * fp) <= 1e-12 and abs(s1["u_sc"] - usc) <= 1e-9 * max(1.0, abs(usc))
        f = ctx.forces(1)
        assert np.abs(f - (f1 + s1["sp"] * df)).max() <= 1e-9 * np.abs(f).max()
        ctx.set_external_dual(1)                                                        # remove
        ctx.eval()
        assert ctx.scalars(1)["E1"] == base[1][0]["E1"] and np.array_equal(ctx.forces(1, _lib.FORCE_STATE1), base[1][1])


def test_excluded_pair_with_different_displacements_under_ewald():
    """A molecule displaced only in part: its excluded pairs change their erf(alpha r)/r correction between the
    states (the branch of the exceptions kernel that feeds dF and u), with and without the reciprocal part."""
    from test_oracle import small_ewald_case
    from oracle import pme as P
    sysd, pos = small_ewald_case(n_mol=40, seed=4)
    disp = np.zeros_like(pos)
    disp[0] = (0.05, 0.02, -0.03)            # atoms 0 and 1 of the first molecule move, atom 2 stays
    disp[1] = (0.05, 0.02, -0.03)
    disp[6:9] = (0.3, 0.1, 0.0)              # a whole molecule with another displacement vector
    case = S.SDMCase("partial", sysd, pos, disp, S.AlchemicalState(lambdac=0.5))
    ref = oracle_eval(case)
    with run_case(case, _lib.PAIR_ALLPAIRS) as ctx:
        check_against_oracle(ctx, case, ref)
        grid = [24, 27, 25]
        ctx.enable_reciprocal_pme(grid)
        ctx.eval()
        alpha = sysd.ewald_alpha_effective()
        e1r, f1r = P.reciprocal(sysd.charge, pos, sysd.box, alpha, grid)
        e2r, f2r = P.reciprocal(sysd.charge, pos + disp, sysd.box, alpha, grid)
        sc = ctx.scalars(0)
        assert abs(sc["E1"] - (ref["E1"] + e1r)) <= 1e-5 * abs(ref["E1"] + e1r)
        assert abs(sc["u"] - (ref["u"] + e2r - e1r)) <= 1e-6 * max(1.0, abs(ref["u"] + e2r - e1r))
        df = ctx.forces(0, _lib.FORCE_DELTA)
        dref = (ref["f2"] + f2r) - (ref["f1"] + f1r)
        assert np.abs(df - dref).max() <= 1e-7 * np.abs(ref["f1"]).max() + 1e-9 * np.abs(dref).max()
