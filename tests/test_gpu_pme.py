"""GPU tests of reciprocal-space PME on the device (csrc/kernels_pme.cu, sdm_enable_reciprocal_pme) against the
numpy restatement of OpenMM's ReferencePME (oracle/pme.py) added to the direct-space oracle: the complete
NonbondedMethod=PME energies and forces of both states, through the C ABI.  FP64 on both sides: 1e-9."""
import copy

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import oracle as O
from oracle import pme as P
from test_gpu_ewald import as_pme
from test_gpu_parity import E_RTOL, F_RMS_RTOL, rms_rel
from test_oracle import small_ewald_case

pytestmark = pytest.mark.gpu


def full_pme_reference(case, grid, positions=None):
    """Direct space (C oracle, two full evaluations) + reciprocal space and self energy (numpy PME) of both states."""
    pos = case.positions if positions is None else positions
    sysd = case.system
    alpha = sysd.ewald_alpha_effective()
    d1 = O.nonbonded(sysd, pos, nthreads=O.max_threads())
    d2 = O.nonbonded(sysd, pos + case.displacement, nthreads=O.max_threads())
    e1r, f1r = P.reciprocal(sysd.charge, pos, sysd.box, alpha, grid)
    e2r, f2r = P.reciprocal(sysd.charge, pos + case.displacement, sysd.box, alpha, grid)
    return dict(E1=d1["E"] + e1r, u=(d2["E"] + e2r) - (d1["E"] + e1r), f1=d1["forces"] + f1r, f2=d2["forces"] + f2r,
                e1_rec=e1r, e2_rec=e2r, f1_rec=f1r, f2_rec=f2r)


def test_small_box_reciprocal_part_alone_and_total():
    sysd, pos = small_ewald_case(n_mol=40, seed=4)
    disp = np.zeros_like(pos)
    disp[:6] = (0.4, -0.2, 0.1)
    case = S.SDMCase("small_pme", sysd, pos, disp, S.AlchemicalState(lambdac=0.5))
    grid = [24, 27, 25]
    ref = full_pme_reference(case, grid)
    with SDMContext(sysd, disp, n_replicas=2, pair_mode=_lib.PAIR_ALLPAIRS) as ctx:
        for r in range(2):
            ctx.set_alchemical(r, case.alch)
            ctx.set_positions(r, pos)
        ctx.eval()
        dir_only = (ctx.scalars(1), ctx.forces(1, _lib.FORCE_STATE1).copy(), ctx.forces(1, _lib.FORCE_DELTA).copy())
        ctx.enable_reciprocal_pme(grid)
        assert [int(ctx.info("pme_grid_" + a)) for a in "xyz"] == grid
        ctx.eval()
        sc = ctx.scalars(1)
        # the reciprocal part by itself: what the library added on top of its direct-space result
        assert abs((sc["E1"] - dir_only[0]["E1"]) - ref["e1_rec"]) <= 1e-9 * abs(ref["e1_rec"])
        assert abs((sc["u"] - dir_only[0]["u"]) - (ref["e2_rec"] - ref["e1_rec"])) <= 1e-9 * abs(ref["e1_rec"])
        df1 = ctx.forces(1, _lib.FORCE_STATE1) - dir_only[1]
        assert np.abs(df1 - ref["f1_rec"]).max() <= 1e-9 * np.abs(ref["f1_rec"]).max()
        ddf = ctx.forces(1, _lib.FORCE_DELTA) - dir_only[2]
        assert np.abs(ddf - (ref["f2_rec"] - ref["f1_rec"])).max() <= 1e-9 * np.abs(ref["f1_rec"]).max()
        # and the total against direct-space oracle + reciprocal oracle
        assert abs(sc["E1"] - ref["E1"]) <= E_RTOL * abs(ref["E1"])
        assert abs(sc["u"] - ref["u"]) <= 1e-6 * max(1.0, abs(ref["u"]))
        assert rms_rel(ctx.forces(1, _lib.FORCE_STATE1), ref["f1"]) <= F_RMS_RTOL
        assert np.array_equal(ctx.forces(0), ctx.forces(1))       # same input, same bits in every replica
        with pytest.raises(_lib.SDMError):
            ctx.set_external_dual(0, np.zeros_like(pos), np.zeros_like(pos), 0.0, 0.0)


def test_cfg2_as_shipped_pme_complete_on_the_cluster_path():
    """example/test_explicit.py:64: nonbondedMethod=PME, 1 nm cutoff, default tolerance -- mesh 48 x 54 x 48."""
    case = as_pme(S.cfg2())
    with SDMContext(case.system, case.displacement, n_replicas=2, pair_mode=_lib.PAIR_CLUSTER) as ctx:
        rng = np.random.default_rng(2)
        pos1 = case.positions + rng.normal(scale=0.002, size=case.positions.shape)
        for r, p in enumerate((case.positions, pos1)):
            ctx.set_alchemical(r, case.alch)
            ctx.set_positions(r, p)
        ctx.enable_reciprocal_pme()
        grid = [int(ctx.info("pme_grid_" + a)) for a in "xyz"]
        assert grid == [48, 54, 48]
        out = []
        for _ in range(3):                                        # build, graph capture, replay
            ctx.eval()
            out.append((ctx.scalars(0)["E1"], ctx.scalars(1)["u"], ctx.forces(1).copy()))
        assert out[1][0] == out[2][0] and out[1][1] == out[2][1] and np.array_equal(out[1][2], out[2][2])
        for r, p in enumerate((case.positions, pos1)):
            ref = full_pme_reference(case, grid, p)
            sc = ctx.scalars(r)
            assert sc["status"] == 0
            assert abs(sc["E1"] - ref["E1"]) <= E_RTOL * abs(ref["E1"]), (sc["E1"], ref["E1"])
            assert abs(sc["u"] - ref["u"]) <= 1e-6 * max(1.0, abs(ref["u"])), (sc["u"], ref["u"])
            assert rms_rel(ctx.forces(r, _lib.FORCE_STATE1), ref["f1"]) <= F_RMS_RTOL
            dref = ref["f2"] - ref["f1"]
            df = ctx.forces(r, _lib.FORCE_DELTA)
            assert np.abs(df - dref).max() <= 1e-7 * np.abs(ref["f1"]).max() + 1e-9 * np.abs(dref).max()


def test_reciprocal_pme_needs_an_ewald_system():
    case = S.cfg1()
    with SDMContext(case.system, case.displacement) as ctx:
        with pytest.raises(_lib.SDMError):
            ctx.enable_reciprocal_pme()


def test_plugin_surface_with_a_pme_system_and_device_dynamics():
    """The integrator mirror bound to a PME system evaluates the complete sum (bind switches the reciprocal part on),
    and the device MD loop runs on it."""
    from openmm_sdm_plugin_b200.sdmplugin import LangevinIntegratorSDM
    sysd, pos = small_ewald_case(n_mol=40, seed=4)
    n = sysd.n_atoms
    integ = LangevinIntegratorSDM(300.0, 1.0, 0.0005, n)
    for i in range(6):
        integ.setDisplacement(i, 0.4, -0.2, 0.1)
    integ.setLambda1(0.3); integ.setLambda2(0.3)
    integ.bind(sysd, pair_mode=_lib.PAIR_ALLPAIRS)
    try:
        integ.evaluate(pos)
        ctx = integ._ctx
        grid = [int(ctx.info("pme_grid_" + a)) for a in "xyz"]
        disp = np.zeros_like(pos); disp[:6] = (0.4, -0.2, 0.1)
        ref = full_pme_reference(S.SDMCase("s", sysd, pos, disp, S.AlchemicalState()), grid)
        sc = ctx.scalars(0)
        assert abs(sc["E1"] - ref["E1"]) <= E_RTOL * abs(ref["E1"]) and abs(sc["u"] - ref["u"]) <= 1e-6 * max(1.0, abs(ref["u"]))
        masses = np.tile([15.999, 1.008, 1.008], n // 3)
        integ.setState(pos, np.zeros_like(pos), masses)
        integ.step(5)
        assert np.isfinite(integ.getPositions()).all() and np.abs(integ.getPositions() - pos).max() > 0.0
        assert ctx.scalars(0)["status"] == 0
    finally:
        integ.cleanup()
