"""GPU parity of the geometric Lennard-Jones combining rule (sdm_system.lj_combining = SDM_LJ_GEOMETRIC): what
createSystem(OPLS=True) of the reference's reader puts in the nonbonded force group (NonbondedForce with zeroed epsilons
+ the CustomNonbondedForce of desmonddmsfile75.py:780-810).  Same bars as test_gpu_parity.py, against the oracle with
the same flag; both shipped fixtures (their nonbonded_info says 'geometric'), both pair kernels."""
import copy

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import oracle as O
from test_gpu_parity import check_against_oracle, run_case

pytestmark = pytest.mark.gpu


def as_opls(case):
    c = copy.copy(case)
    c.system = copy.copy(case.system)
    c.system.lj_geometric = True
    c.system.use_dispersion_correction = False     # desmonddmsfile75.py:428,438
    return c


@pytest.mark.parametrize("name,mode", [("cfg1", _lib.PAIR_ALLPAIRS), ("cfg2", _lib.PAIR_ALLPAIRS), ("cfg2", _lib.PAIR_CLUSTER)])
def test_geometric_rule_against_the_oracle(name, mode):
    case = as_opls(getattr(S, name)())
    ref = O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement, case.positions,
                     nthreads=O.max_threads())
    lb = O.sdm_eval(getattr(S, name)().system, S.AlchemicalState(**vars(case.alch)), case.displacement, case.positions,
                    nthreads=O.max_threads())
    assert abs(ref["E1_pair"] - lb["E1_pair"]) > 1e-4 * abs(lb["E1_pair"])      # the rule matters on these fixtures
    assert ref["n_pairs1"] == lb["n_pairs1"] and ref["E1_disp"] == 0.0
    with run_case(case, mode) as ctx:
        check_against_oracle(ctx, case, ref)
        if mode == _lib.PAIR_CLUSTER:
            first = (ctx.scalars(0), ctx.forces(0).copy())
            ctx.eval()                                                          # graph capture
            ctx.eval()                                                          # replay
            check_against_oracle(ctx, case, ref)
            assert ctx.scalars(0)["E1"] == first[0]["E1"] and np.array_equal(ctx.forces(0), first[1])
            pairs = ctx.pairs(0)                                                # debug build: the pair set does not depend on the rule
            want = O.nonbonded(case.system, case.positions, want_pairs=True, nthreads=O.max_threads())["pairs"]
            assert np.array_equal(pairs, want)


def test_geometric_rule_with_pme_direct_and_reciprocal_space():
    case = as_opls(S.cfg2())
    case.system.method = S.PME                                                  # test_explicit.py:64 with OPLS=True
    ref = O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement, case.positions,
                     nthreads=O.max_threads())
    with run_case(case, _lib.PAIR_CLUSTER) as ctx:                              # direct space only: like the oracle
        check_against_oracle(ctx, case, ref)


def test_dispersion_correction_is_refused_with_the_geometric_rule():
    case = as_opls(S.cfg2())
    case.system.use_dispersion_correction = True
    with pytest.raises(_lib.SDMError):
        SDMContext(case.system, case.displacement, n_replicas=1)
    c1 = as_opls(S.cfg1())
    c1.system.use_dispersion_correction = True                                  # not periodic: nothing to refuse
    with SDMContext(c1.system, c1.displacement, n_replicas=1) as ctx:
        ctx.set_positions(0, c1.positions)
        ctx.set_alchemical(0, c1.alch)
        ctx.eval()
        assert ctx.scalars(0)["status"] == 0
