"""GPU tests of the SDMUtils restraint forces (csrc/kernels_restraints.cu) against the autograd oracle
(oracle/restraints.py), through the C ABI: energies to 1e-10 relative, forces to 1e-9 of the largest force."""
import math

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from openmm_sdm_plugin_b200.sdmplugin import LangevinIntegratorSDM, SDMUtils
from oracle import restraints as R

pytestmark = pytest.mark.gpu
KCAL = 4.184


def example_terms(case):
    """The restraints of example/test.py:78-150 on the 230-atom fixture: Vsite centroid restraints of both ligands
    (kf = 25 kcal/mol/A^2, r0 = 4.5 A, the second with the displacement as offset) and the alignment force
    (25 kcal/mol/A^2, 50, 50 kcal/mol), plus -- not in the script, which passes None -- orientational windows."""
    z = np.load(S.os.path.join(S.GOLDEN_DIR, "cfg1_oa_g6_g3.npz"))
    resid = z["resid"]
    rcpt, lig1, lig2 = np.where(resid == 1)[0], np.where(resid == 3)[0], np.where(resid == 2)[0]
    displ = np.array([-15.559, -3.000, 8.600]) * 0.1
    kf, r0 = 25.0 * KCAL * 100.0, 0.45
    m = case.masses if case.masses is not None else np.ones(len(resid))
    cen = [dict(lig_cm_atoms=lig1.tolist(), rcpt_cm_atoms=rcpt.tolist(), lig_cm_weights=m[lig1], rcpt_cm_weights=m[rcpt],
                kfcm=kf, tolcm=r0, offset=(0.0, 0.0, 0.0)),
           dict(lig_cm_atoms=lig2.tolist(), rcpt_cm_atoms=rcpt.tolist(), lig_cm_weights=m[lig2], rcpt_cm_weights=m[rcpt],
                kfcm=kf, tolcm=r0, offset=displ.tolist(),
                lig_ref=[int(lig2[0]) + k for k in (5, 4, 3)], rcpt_ref=[int(rcpt[0]) + k for k in (0, 7, 19)],
                kfcd=[40.0, 30.0, 20.0], a=[0.2, -2.9, 0.4], b=[0.5, -2.5, 0.8])]
    ali = [dict(liga_ref=[int(lig1[0]) + k for k in (4, 3, 2)], ligb_ref=[int(lig2[0]) + k for k in (5, 4, 3)],
                kfdispl=25.0 * KCAL * 100.0, ktheta=50.0 * KCAL, kpsi=50.0 * KCAL, offset=displ.tolist())]
    return cen, ali


def add_terms(ctx, cen, ali):
    for s in cen:
        ctx.add_centroid_restraint(s["lig_cm_atoms"], s["rcpt_cm_atoms"], s["kfcm"], s["tolcm"], s["offset"],
                                   s.get("lig_cm_weights"), s.get("rcpt_cm_weights"), s.get("lig_ref"), s.get("rcpt_ref"),
                                   s.get("kfcd", (0, 0, 0)), s.get("a", (0, 0, 0)), s.get("b", (0, 0, 0)))
    for s in ali:
        ctx.add_alignment_restraint(s["liga_ref"], s["ligb_ref"], s["kfdispl"], s["ktheta"], s["kpsi"], s["offset"])


@pytest.mark.parametrize("replicas", [1, 3])
def test_restraint_energy_and_forces_match_the_oracle(replicas):
    case = S.cfg1()
    cen, ali = example_terms(case)
    rng = np.random.default_rng(5)
    frames = [case.positions + rng.normal(scale=0.03, size=case.positions.shape) for _ in range(replicas)]
    with SDMContext(case.system, case.displacement, n_replicas=replicas) as ctx:
        for r in range(replicas):
            ctx.set_alchemical(r, case.alch)
            ctx.set_positions(r, frames[r])
        ctx.eval()
        base = [(ctx.forces(r).copy(), ctx.forces(r, _lib.FORCE_STATE1).copy(), ctx.scalars(r)) for r in range(replicas)]
        add_terms(ctx, cen, ali)
        ctx.eval()
        for r in range(replicas):
            e_ref, f_ref = R.energy_and_forces(frames[r], cen, ali)
            assert e_ref > 1.0                                   # the jittered frame really violates the restraints
            sc = ctx.scalars(r)
            e = ctx.restraint_energy(r)
            assert abs(e - e_ref) <= 1e-10 * abs(e_ref)
            assert abs((sc["pot_energy"] - base[r][2]["pot_energy"]) - e_ref) <= 1e-9 * abs(e_ref)
            assert sc["bind_e"] == base[r][2]["bind_e"] and sc["u"] == base[r][2]["u"]
            df = ctx.forces(r) - base[r][0]
            assert np.abs(df - f_ref).max() <= 1e-9 * np.abs(f_ref).max()
            assert np.array_equal(ctx.forces(r, _lib.FORCE_STATE1), base[r][1])    # F1 stays the nonbonded force
        ctx.set_restraint_control(0.5)                           # scales the centroid terms only (SDMUtils.py:61-66)
        ctx.eval()
        e_half = R.energy_and_forces(frames[0], cen, ali, control=0.5)[0]
        assert abs(ctx.restraint_energy(0) - e_half) <= 1e-10 * abs(e_half)
        ctx.clear_restraints()
        ctx.eval()
        assert np.array_equal(ctx.forces(0), base[0][0]) and ctx.restraint_energy(0) == 0.0
        assert ctx.scalars(0)["pot_energy"] == base[0][2]["pot_energy"]


def test_restraints_on_the_cluster_path_and_under_graph_replay():
    case = S.cfg2()
    z = np.load(S.os.path.join(S.GOLDEN_DIR, "cfg2_temoa_g1_g4.npz"))
    resid = z["resid"]
    rcpt, lig1 = np.where(resid == 1)[0], np.where(resid == 2)[0]
    cen = [dict(lig_cm_atoms=lig1.tolist(), rcpt_cm_atoms=rcpt.tolist(), lig_cm_weights=case.masses[lig1],
                rcpt_cm_weights=case.masses[rcpt], kfcm=25.0 * KCAL * 100.0, tolcm=0.05, offset=(0.0, 0.0, 0.0))]
    e_ref, f_ref = R.energy_and_forces(case.positions, cen, [])
    with SDMContext(case.system, case.displacement, pair_mode=_lib.PAIR_CLUSTER) as ctx:
        ctx.set_alchemical(0, case.alch)
        ctx.set_positions(0, case.positions)
        ctx.eval()
        f0 = ctx.forces(0).copy()
        add_terms(ctx, cen, [])
        for _ in range(4):                                       # build, capture, two replays
            ctx.eval()
            assert abs(ctx.restraint_energy(0) - e_ref) <= 1e-10 * max(abs(e_ref), 1e-30)
            assert np.abs((ctx.forces(0) - f0) - f_ref).max() <= 1e-9 * max(np.abs(f_ref).max(), 1e-30)


def test_the_plugin_surface_applies_what_sdmutils_recorded():
    """example/test.py:104-150 in the mirror's terms: SDMUtils(system).addRestraintForce / addAlignmentForce, then the
    integrator bound to that system evaluates them (force group 1 of the reference)."""
    case = S.cfg1()
    cen, ali = example_terms(case)
    u = SDMUtils(case.system)
    u.addRestraintForce(lig_cm_particles=cen[0]["lig_cm_atoms"], rcpt_cm_particles=cen[0]["rcpt_cm_atoms"],
                        kfcm=cen[0]["kfcm"], tolcm=cen[0]["tolcm"], lig_cm_weights=cen[0]["lig_cm_weights"],
                        rcpt_cm_weights=cen[0]["rcpt_cm_weights"])
    u.addAlignmentForce(liga_ref_particles=ali[0]["liga_ref"], ligb_ref_particles=ali[0]["ligb_ref"],
                        kfdispl=ali[0]["kfdispl"], ktheta=ali[0]["ktheta"], kpsi=ali[0]["kpsi"], offset=ali[0]["offset"])
    n = case.system.n_atoms
    integ = LangevinIntegratorSDM(300.0, 0.5, 0.001, n)
    for i in np.nonzero(np.abs(case.displacement).sum(1))[0]:
        integ.setDisplacement(int(i), *case.displacement[i])
    integ.setBiasMethod(u.ILogisticMethod)
    integ.setSoftCoreMethod(u.RationalSoftCoreMethod)
    integ.setLambda1(0.5); integ.setLambda2(0.5); integ.setUmax(100 * KCAL); integ.setUbcore(50 * KCAL); integ.setAcore(0.0625)
    pos = case.positions + np.random.default_rng(8).normal(scale=0.03, size=case.positions.shape)
    integ.bind(case.system)
    try:
        integ.evaluate(pos)
        e_ref, _ = R.energy_and_forces(pos, [cen[0]], ali)
        assert abs(integ._ctx.restraint_energy(0) - e_ref) <= 1e-10 * abs(e_ref)
    finally:
        integ.cleanup()
        del case.system.sdm_restraints
