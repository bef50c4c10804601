"""GPU tests of the HCT-GB + ACE model on the device (csrc/kernels_gb.cu, sdm_enable_hct_gb) against the torch
restatement of OpenMM's GBSAHCTForce expressions (oracle/gb.py), through the C ABI: what the library adds to E1,
u, F1 and F2 - F1 when the implicit-solvent force of desmonddmsfile75.py:454-465 sits in the nonbonded force
group.  FP64 on both sides: 1e-9."""
import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, sdmplugin, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import gb as G
from test_gb_oracle import blob, gb_parameters

pytestmark = pytest.mark.gpu


def gb_reference(case, q, o, sr, positions=None, **kw):
    pos = case.positions if positions is None else positions
    e1, f1, b1 = G.hct(pos, q, o, sr, **kw)
    e2, f2, b2 = G.hct(pos + case.displacement, q, o, sr, **kw)
    return dict(e1=e1, e2=e2, f1=f1, f2=f2, b1=b1, b2=b2)


def check_gb_part(ctx, r, base, ref, tol=1e-9, ftol=None):
    sc = ctx.scalars(r)
    ftol = tol if ftol is None else ftol
    fs = np.abs(ref["f1"]).max()
    assert abs((sc["E1"] - base["E1"]) - ref["e1"]) <= tol * abs(ref["e1"])
    assert abs((sc["u"] - base["u"]) - (ref["e2"] - ref["e1"])) <= tol * abs(ref["e1"])
    assert np.abs(ctx.forces(r, _lib.FORCE_STATE1) - base["f1"] - ref["f1"]).max() <= ftol * fs
    assert np.abs(ctx.forces(r, _lib.FORCE_DELTA) - base["df"] - (ref["f2"] - ref["f1"])).max() <= tol * fs
    assert np.allclose(ctx.born_radii(r, 1), ref["b1"], rtol=1e-12, atol=0)
    assert np.allclose(ctx.born_radii(r, 2), ref["b2"], rtol=1e-12, atol=0)


def snapshot(ctx, r):
    sc = ctx.scalars(r)
    return dict(E1=sc["E1"], u=sc["u"], f1=ctx.forces(r, _lib.FORCE_STATE1).copy(), df=ctx.forces(r, _lib.FORCE_DELTA).copy())


def test_cfg1_host_guest_with_hct_gb_both_states():
    """The 230-atom host-guest fixture (example/test.py's system, no cutoff) with GB parameters per atom."""
    case = S.cfg1()
    n = case.system.n_atoms
    o, sr = gb_parameters(n, seed=1)
    q = case.system.charge
    rng = np.random.default_rng(5)
    pos1 = case.positions + rng.normal(scale=0.003, size=case.positions.shape)
    with SDMContext(case.system, case.displacement, n_replicas=2, pair_mode=_lib.PAIR_ALLPAIRS) as ctx:
        for r, p in enumerate((case.positions, pos1)):
            ctx.set_alchemical(r, case.alch)
            ctx.set_positions(r, p)
        ctx.eval()
        base = [snapshot(ctx, r) for r in range(2)]
        ctx.enable_hct_gb(o, sr)                                   # charges: the NonbondedForce ones
        assert ctx.info("gb_sa_ace") == 1.0
        assert ctx.info("gb_prefactor") == pytest.approx(138.935485 * (1.0 - 1.0 / 78.5), rel=1e-15)
        ctx.eval()
        for r, p in enumerate((case.positions, pos1)):
            check_gb_part(ctx, r, base[r], gb_reference(case, q, o, sr, positions=p))
        # the hybrid force carries the GB part of both states: F = F1 + sp (F2 - F1)
        sc = ctx.scalars(1)
        mix = ctx.forces(1, _lib.FORCE_STATE1) + sc["sp"] * ctx.forces(1, _lib.FORCE_DELTA)
        assert np.abs(ctx.forces(1) - mix).max() <= 1e-9 * np.abs(mix).max()
        first = (ctx.scalars(0), ctx.forces(0).copy())
        ctx.eval()                                                 # same input, same bits
        assert ctx.scalars(0)["u"] == first[0]["u"] and np.array_equal(ctx.forces(0), first[1])
        with pytest.raises(_lib.SDMError):
            ctx.set_external_dual(0, np.zeros_like(pos1), np.zeros_like(pos1), 0.0, 0.0)


def test_own_charges_dielectrics_and_no_surface_term():
    n = 150
    pos, q = blob(n, seed=7, spacing=0.28)
    o, sr = gb_parameters(n, seed=7)
    sysd = S.NonbondedSystem(q, np.full(n, 0.3), np.full(n, 0.2), np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, 3)),
                             method=S.NOCUTOFF)
    disp = np.zeros_like(pos)
    disp[:9] = (0.0, 0.25, -0.4)
    case = S.SDMCase("blob_gb", sysd, pos, disp, S.AlchemicalState(lambdac=0.5))
    q_gb = q * 0.9 + 0.01                                          # the hct table carries its own charge column
    with SDMContext(sysd, disp, n_replicas=1, pair_mode=_lib.PAIR_ALLPAIRS) as ctx:
        ctx.set_alchemical(0, case.alch)
        ctx.set_positions(0, pos)
        ctx.eval()
        base = snapshot(ctx, 0)
        ctx.enable_hct_gb(o, sr, charge=q_gb, solute_dielectric=2.0, solvent_dielectric=60.0, sa_ace=False)
        ctx.eval()
        check_gb_part(ctx, 0, base, gb_reference(case, q_gb, o, sr, solute_dielectric=2.0, solvent_dielectric=60.0,
                                                 sa_ace=False))
        ctx.enable_hct_gb(o, sr, charge=q_gb)                      # switched on again with other options
        ctx.eval()
        check_gb_part(ctx, 0, base, gb_reference(case, q_gb, o, sr))


def test_larger_nonperiodic_system_on_the_cluster_path_under_graph_replay():
    n = 3400
    pos, q = blob(n, seed=11, spacing=0.3)
    o, sr = gb_parameters(n, seed=11)
    sysd = S.NonbondedSystem(q, np.full(n, 0.25), np.full(n, 0.3), np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, 3)),
                             method=S.CUTOFF_NONPERIODIC, cutoff=0.9, eps_rf=1.0)
    disp = np.zeros_like(pos)
    disp[:20] = (0.15, 0.0, 0.45)
    case = S.SDMCase("blob_gb_cluster", sysd, pos, disp, S.AlchemicalState(lambdac=0.5))
    rng = np.random.default_rng(1)
    with SDMContext(sysd, disp, n_replicas=1, pair_mode=_lib.PAIR_CLUSTER) as ctx:
        ctx.set_alchemical(0, case.alch)
        ctx.set_positions(0, pos)
        ctx.eval()
        ctx.enable_hct_gb(o, sr)
        for it in range(3):                                        # build, graph capture, replay
            p = pos + rng.normal(scale=0.001, size=pos.shape)
            with SDMContext(sysd, disp, n_replicas=1, pair_mode=_lib.PAIR_CLUSTER) as plain:
                plain.set_alchemical(0, case.alch)
                plain.set_positions(0, p)
                plain.eval()
                base = snapshot(plain, 0)
            ctx.set_positions(0, p)
            ctx.eval()
            # the pair part to subtract comes from another context, whose list was built at other positions: its
            # FP32 cell-relative coordinates round differently (1e-7 of the force scale); FP64 parts: 1e-8
            check_gb_part(ctx, 0, base, gb_reference(case, q, o, sr, positions=p), tol=1e-8, ftol=1e-6)


def test_refused_where_the_reference_model_does_not_apply():
    case = S.cfg2()
    n = case.system.n_atoms
    o, sr = gb_parameters(n)
    with SDMContext(case.system, case.displacement, n_replicas=1) as ctx:
        with pytest.raises(_lib.SDMError):                         # periodic box: GB is a NoCutoff CustomGBForce
            ctx.enable_hct_gb(o, sr)
    c1 = S.cfg1()
    with SDMContext(c1.system, c1.displacement, n_replicas=1) as ctx:
        o1, s1 = gb_parameters(c1.system.n_atoms)
        with pytest.raises(_lib.SDMError):
            ctx.enable_hct_gb(-o1, s1)
        with pytest.raises(_lib.SDMError):
            ctx.born_radii(0)                                      # not switched on
        with pytest.raises(ValueError):
            ctx.enable_hct_gb(o1[:-1], s1[:-1])


def test_plugin_surface_binds_the_gb_force_of_the_system():
    """integrator + system as the reader builds them for implicitSolvent=HCT: evaluate() carries GB in both states."""
    case = S.cfg1()
    n = case.system.n_atoms
    rng = np.random.default_rng(2)
    radius, scale = rng.uniform(0.12, 0.2, n), rng.uniform(0.72, 0.88, n)
    gb = S.GBSAHCTForce(SA="ACE")
    for a in range(n):
        gb.addParticle([case.system.charge[a], radius[a], scale[a]])
    gb.finalize()
    import copy
    sysd = copy.copy(case.system)
    sysd.eps_rf = 1.0
    sysd.addForce(gb)
    q, o, sr = gb.device_parameters()

    def run(system):
        it = sdmplugin.LangevinIntegratorSDM(300.0, 0.5, 0.001, n)
        it.setLambda(0.5)
        for a in np.nonzero(np.abs(case.displacement).sum(1))[0]:
            it.setDisplacement(int(a), *case.displacement[a])
        it.bind(system)
        f = it.evaluate(case.positions)
        out = (it.getBindE(), it.getPotEnergy(), np.array(f))
        it.cleanup()
        return out

    plain = copy.copy(case.system)
    plain.eps_rf = 1.0
    u0, e0, _ = run(plain)
    u1, e1, _ = run(sysd)
    ref = gb_reference(case, q, o, sr)
    assert (u1 - u0) == pytest.approx(ref["e2"] - ref["e1"], rel=1e-9, abs=1e-9 * abs(ref["e1"]))


def test_device_dynamics_with_hct_gb_follow_the_reference_integrator():
    """integrator.step(n) with the GB force in the nonbonded group: two steps of the 230-atom fixture with the
    reference's noise against the reference's own integrator + kernels (oracle/_ref) whose force-group-2 callback is
    the nonbonded oracle plus the GB oracle."""
    import copy
    from oracle import oracle as O
    from oracle import reference as R
    if not R.available():
        pytest.skip("oracle/_ref is built only where /root/reference exists")
    case = S.cfg1()
    n = case.system.n_atoms
    rng = np.random.default_rng(23)
    sysd = copy.copy(case.system)
    sysd.eps_rf = 1.0
    gb = S.GBSAHCTForce(SA="ACE")
    for a in range(n):
        gb.addParticle([sysd.charge[a], rng.uniform(0.12, 0.2), rng.uniform(0.72, 0.88)])
    gb.finalize()
    sysd.addForce(gb)
    q, o, sr = gb.device_parameters()
    vel = rng.normal(scale=0.3, size=(n, 3))
    xi = rng.normal(size=(2, n, 3))

    def force_fn(groups, pos):
        if groups == 4:
            r = O.nonbonded(sysd, pos, nthreads=1)
            e, f, _ = G.hct(pos, q, o, sr)
            return r["E"] + e, r["forces"] + f
        return 0.0, np.zeros_like(pos)
    ref = R.run(case.masses, case.positions, vel, case.displacement,
                R.params_from_alch(case.alch, temperature=300.0, friction=0.5), force_fn, steps=2, noise=xi.ravel())
    integ = sdmplugin.LangevinIntegratorSDM(300.0, 0.5, case.alch.step_size, n)
    integ.setBiasMethod(case.alch.bias_method)
    integ.setSoftCoreMethod(case.alch.softcore_method)
    integ.setLambda1(case.alch.lambda1); integ.setLambda2(case.alch.lambda2); integ.setAlpha(case.alch.alpha)
    integ.setU0(case.alch.u0); integ.setW0coeff(case.alch.w0coeff)
    integ.setUmax(case.alch.umax); integ.setUbcore(case.alch.ubcore); integ.setAcore(case.alch.acore)
    for i in np.nonzero(np.abs(case.displacement).sum(1))[0]:
        integ.setDisplacement(int(i), *case.displacement[i])
    integ.bind(sysd)
    try:
        integ.setState(case.positions, vel, case.masses)
        integ.step(0)
        for k in range(2):
            integ._ctx.md_set_noise(xi[k][None])
            integ.step(1)
        assert np.abs(integ.getPositions() - ref["positions"]).max() < 1e-8
        assert integ.getBindE() == pytest.approx(ref["bind_e"], abs=1e-6)
        assert integ.getPotEnergy() == pytest.approx(ref["pot_energy"], rel=1e-5)
    finally:
        integ.cleanup()
