"""GPU parity at BASELINE.json's full sizes: cfg4 (synthetic 50 k-atom box, two replicas) and cfg5
(100 k and 200 k atoms) against the THREADED oracle -- energies, u, forces and the sorted in-cutoff
pair set (bit-exact), through the C ABI.  The oracle's two-pass evaluation of 200 k atoms takes
about a second on the box's host cores, so nothing here needs a property-only shortcut."""
import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import oracle as O
from test_gpu_parity import check_against_oracle

pytestmark = pytest.mark.gpu
CL = _lib.PAIR_CLUSTER


def oracle_eval(case, positions):
    return O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement, positions,
                      nthreads=O.max_threads())


def oracle_pairs(case, positions):
    return O.nonbonded(case.system, positions, want_pairs=True, nthreads=O.max_threads())["pairs"]


def test_cfg4_50k_atoms_two_replicas():
    """cfg4: the synthetic ~50 k-atom protein-ligand box of the bench (seed 1234, 60 displaced atoms);
    two resident replicas at different coordinates, each held to the oracle at ITS coordinates."""
    case = S.synthetic_case(50_000, 60, seed=1234)
    rng = np.random.default_rng(44)
    pos = [case.positions, case.positions + rng.normal(scale=0.004, size=case.positions.shape)]
    with SDMContext(case.system, case.displacement, n_replicas=2, pair_mode=CL) as ctx:
        for r in range(2):
            ctx.set_positions(r, pos[r])
            ctx.set_alchemical(r, case.alch)
        ctx.eval()
        for r in range(2):
            ref = oracle_eval(case, pos[r])
            sc = check_against_oracle(ctx, case, ref, replica=r)
            assert sc["n_pairs1"] > 200 * case.system.n_atoms * 0.9
            assert np.array_equal(ctx.pairs(r), oracle_pairs(case, pos[r]))


@pytest.mark.parametrize("n_atoms,ligand", [(100_000, 60), (200_000, 100)])
def test_cfg5_large_boxes(n_atoms, ligand):
    case = S.synthetic_case(n_atoms, ligand, seed=27)
    ref = oracle_eval(case, case.positions)
    with SDMContext(case.system, case.displacement, n_replicas=1, pair_mode=CL) as ctx:
        ctx.set_positions(0, case.positions)
        ctx.set_alchemical(0, case.alch)
        ctx.eval()
        check_against_oracle(ctx, case, ref)
        got = ctx.pairs(0)
    want = oracle_pairs(case, case.positions)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
