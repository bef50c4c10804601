"""GPU parity tests: libsdmb200 (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): energies <= 1e-5 relative, forces <= 1e-4 RMS relative,
u compared ABSOLUTELY (it is a small difference of large energies: <= 1e-6 kJ/mol * max(1,|u|)
because the moved-pair path is FP64), in-cutoff pair sets bit-exact.
"""
import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext
from oracle import oracle as O

pytestmark = pytest.mark.gpu

E_RTOL = 1e-5
F_RMS_RTOL = 1e-4
U_ATOL = 1e-6


def rms_rel(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def run_case(case, pair_mode, replicas=1, fb=None, eb=0.0, **kw):
    ctx = SDMContext(case.system, case.displacement, n_replicas=replicas, pair_mode=pair_mode, **kw)
    for r in range(replicas):
        ctx.set_positions(r, case.positions)
        ctx.set_alchemical(r, case.alch)
        if fb is not None:
            ctx.set_bonded_forces(r, fb, eb)
    ctx.eval()
    return ctx


def check_against_oracle(ctx, case, ref, replica=0, u_atol=U_ATOL):
    sc = ctx.scalars(replica)
    assert sc["status"] == 0, sc
    assert sc["n_pairs1"] == ref["n_pairs1"], (sc["n_pairs1"], ref["n_pairs1"])
    # north_star's bar, literally: <= 1e-5 RELATIVE to the energy itself (E1 and PotEnergy).  The
    # pair part alone may be a smaller remainder of the three terms, so it is held to the same
    # absolute error as the total it is part of.
    assert abs(sc["E1"] - ref["E1"]) <= E_RTOL * abs(ref["E1"]), (sc["E1"], ref["E1"])
    assert abs(sc["E1_pair"] - ref["E1_pair"]) <= E_RTOL * max(abs(ref["E1_pair"]), abs(ref["E1"]))
    assert abs(sc["E1_exc"] - ref["E1_exc"]) <= 1e-9 * max(1.0, abs(ref["E1_exc"]))
    assert abs(sc["E1_disp"] - ref["E1_disp"]) <= 1e-9 * max(1.0, abs(ref["E1_disp"]))
    tol_u = u_atol * max(1.0, abs(ref["u"]))
    assert abs(sc["u"] - ref["u"]) <= tol_u, (sc["u"], ref["u"])
    for k in ("u_sc", "fp", "ebias", "bfp", "sp"):
        assert abs(sc[k] - ref[k]) <= 1e-6 * max(1.0, abs(ref[k])), (k, sc[k], ref[k])
    assert abs(sc["pot_energy"] - ref["pot_energy"]) <= E_RTOL * abs(ref["pot_energy"])
    f1 = ctx.forces(replica, _lib.FORCE_STATE1)
    df = ctx.forces(replica, _lib.FORCE_DELTA)
    f = ctx.forces(replica, _lib.FORCE_HYBRID)
    assert rms_rel(f1, ref["f1"]) <= F_RMS_RTOL, rms_rel(f1, ref["f1"])
    dref = ref["f2"] - ref["f1"]
    # dF is FP64 over moved pairs; the oracle's F2-F1 carries cancellation noise of two sums
    tol_df = 1e-7 * max(1.0, np.abs(ref["f1"]).max()) + 1e-9 * np.abs(dref).max()
    assert np.abs(df - dref).max() <= tol_df, (np.abs(df - dref).max(), tol_df)
    assert rms_rel(f, ref["forces"]) <= F_RMS_RTOL, rms_rel(f, ref["forces"])
    return sc


@pytest.fixture(scope="module")
def cfg1():
    c = S.cfg1()
    return c, O.sdm_eval(c.system, S.AlchemicalState(**vars(c.alch)), c.displacement, c.positions)


@pytest.fixture(scope="module")
def cfg2():
    c = S.cfg2()
    return c, O.sdm_eval(c.system, S.AlchemicalState(**vars(c.alch)), c.displacement, c.positions,
                         nthreads=O.max_threads())


def test_cfg1_allpairs(cfg1):
    case, ref = cfg1
    with run_case(case, _lib.PAIR_ALLPAIRS) as ctx:
        sc = check_against_oracle(ctx, case, ref)
        assert sc["n_pairs1"] == 25061
        assert sc["n_moved1"] == 7656 and sc["n_moved2"] == 7656
        assert abs(sc["u"] - 3.6086162625) < 1e-8


def test_cfg2_allpairs(cfg2):
    case, ref = cfg2
    with run_case(case, _lib.PAIR_ALLPAIRS) as ctx:
        sc = check_against_oracle(ctx, case, ref)
        assert sc["n_pairs1"] == 4197871
        assert sc["n_moved1"] == 15175 and sc["n_moved2"] == 15138
        assert abs(sc["u"] - (-6.6429627853)) < 1e-7


def test_cfg1_pair_set_bit_exact(cfg1):
    case, _ = cfg1
    ref = O.nonbonded(case.system, case.positions, want_pairs=True)
    with run_case(case, _lib.PAIR_ALLPAIRS) as ctx:
        got = ctx.pairs(0)
    assert got.shape == ref["pairs"].shape
    assert np.array_equal(got, ref["pairs"])


def test_cfg2_pair_set_bit_exact(cfg2):
    case, _ = cfg2
    ref = O.nonbonded(case.system, case.positions, want_pairs=True, nthreads=O.max_threads())
    with run_case(case, _lib.PAIR_ALLPAIRS) as ctx:
        got = ctx.pairs(0)
    assert got.shape == ref["pairs"].shape
    assert np.array_equal(got, ref["pairs"])


def test_bonded_forces_and_eb(cfg1):
    case, _ = cfg1
    rng = np.random.default_rng(7)
    fb = rng.normal(scale=50.0, size=(case.system.n_atoms, 3))
    ref = O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement,
                     case.positions, fb=fb, eb=12.5)
    with run_case(case, _lib.PAIR_ALLPAIRS, fb=fb, eb=12.5) as ctx:
        sc = check_against_oracle(ctx, case, ref)
        assert sc["Eb"] == 12.5


def test_replica_batch_identical_and_independent(cfg1):
    """R replicas in one context: same inputs give bit-identical outputs; a replica with a
    different lambda state only differs in the alchemical scalars / hybrid force."""
    case, ref = cfg1
    ctx = SDMContext(case.system, case.displacement, n_replicas=3, pair_mode=_lib.PAIR_ALLPAIRS)
    al2 = S.AlchemicalState(**vars(case.alch))
    al2.lambda1, al2.lambda2 = 0.1, 0.4
    al2.alpha, al2.u0 = 0.05, 2.0
    for r in range(3):
        ctx.set_positions(r, case.positions)
        ctx.set_alchemical(r, al2 if r == 2 else case.alch)
    ctx.eval()
    f0, f1, f2 = (ctx.forces(r) for r in range(3))
    s0, s1, s2 = (ctx.scalars(r) for r in range(3))
    assert np.array_equal(f0, f1) and s0 == s1
    assert s2["E1"] == s0["E1"] and s2["u"] == s0["u"]
    ref2 = O.sdm_eval(case.system, S.AlchemicalState(**vars(al2)), case.displacement, case.positions)
    assert abs(s2["sp"] - ref2["sp"]) < 1e-9
    assert rms_rel(f2, ref2["forces"]) <= F_RMS_RTOL
    ctx.close()


def test_zero_displacement_gives_u_zero(cfg1):
    case, _ = cfg1
    with SDMContext(case.system, None, pair_mode=_lib.PAIR_ALLPAIRS) as ctx:
        ctx.set_positions(0, case.positions)
        ctx.set_alchemical(0, case.alch)
        ctx.eval()
        sc = ctx.scalars(0)
        assert sc["u"] == 0.0 and sc["n_moved1"] == 0
        assert np.array_equal(ctx.forces(0), ctx.forces(0, _lib.FORCE_STATE1))


def test_rigid_translation_of_all_atoms_changes_nothing(cfg1):
    """Displacing EVERY atom by the same vector is a rigid translation: u = 0."""
    case, _ = cfg1
    disp = np.tile([0.3, -0.2, 0.5], (case.system.n_atoms, 1))
    with SDMContext(case.system, disp, pair_mode=_lib.PAIR_ALLPAIRS) as ctx:
        ctx.set_positions(0, case.positions)
        ctx.eval()
        assert ctx.scalars(0)["u"] == 0.0


def test_set_displacement_later(cfg1):
    case, ref = cfg1
    with SDMContext(case.system, None, pair_mode=_lib.PAIR_ALLPAIRS) as ctx:
        ctx.set_positions(0, case.positions)
        ctx.set_alchemical(0, case.alch)
        ctx.set_displacement(case.displacement)
        ctx.eval()
        check_against_oracle(ctx, case, ref)


@pytest.mark.parametrize("method", [S.NOCUTOFF, S.CUTOFF_NONPERIODIC])
def test_small_nonperiodic_methods(method):
    case = S.cfg1()
    case.system.method = method
    case.system.cutoff = 1.2
    ref = O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement, case.positions)
    with run_case(case, _lib.PAIR_ALLPAIRS) as ctx:
        check_against_oracle(ctx, case, ref)
        if method == S.CUTOFF_NONPERIODIC:
            pr = O.nonbonded(case.system, case.positions, want_pairs=True)["pairs"]
            assert np.array_equal(ctx.pairs(0), pr)


def test_synthetic_periodic_small():
    case = S.synthetic_case(n_atoms=3000, ligand_atoms=30, seed=5, protein_atoms=300,
                            displacement=(0.0, 0.0, 1.5))
    ref = O.sdm_eval(case.system, S.AlchemicalState(**vars(case.alch)), case.displacement, case.positions)
    with run_case(case, _lib.PAIR_ALLPAIRS) as ctx:
        check_against_oracle(ctx, case, ref)


@pytest.mark.parametrize("bias,soft", [(0, 0), (1, 1), (2, 2), (2, 1)])
def test_alchemical_methods(cfg1, bias, soft):
    case, _ = cfg1
    al = S.AlchemicalState(bias_method=bias, softcore_method=soft, lambdac=0.3, gammac=0.01,
                           wbcoeff=0.2, w0coeff=1.5, lambda1=0.1, lambda2=0.45, alpha=0.3,
                           u0=2.0, umax=3.0, acore=0.0625, ubcore=1.0)
    ref = O.sdm_eval(case.system, S.AlchemicalState(**vars(al)), case.displacement, case.positions)
    c2 = S.SDMCase(case.name, case.system, case.positions, case.displacement, al)
    with run_case(c2, _lib.PAIR_ALLPAIRS) as ctx:
        check_against_oracle(ctx, c2, ref)


def test_unknown_softcore_method_is_reported(cfg1):
    """LangevinIntegratorSDM.cpp:126-129,147: the u <= ub early return comes first, so an unknown
    method only throws when u > ub.  The C ABI reports it as status SDM_ERR_SOFTCORE."""
    case, _ = cfg1
    al = S.AlchemicalState(**vars(case.alch))
    al.softcore_method = 7
    c2 = S.SDMCase(case.name, case.system, case.positions, case.displacement, al)
    with run_case(c2, _lib.PAIR_ALLPAIRS) as ctx:
        assert ctx.scalars(0)["status"] == 0          # u = 3.6 <= ub = 209.2: no throw
    al.ubcore = 1.0                                    # now u > ub
    with pytest.raises(ValueError):
        O.sdm_eval(case.system, S.AlchemicalState(**vars(al)), case.displacement, case.positions)
    with run_case(c2, _lib.PAIR_ALLPAIRS) as ctx:
        assert ctx.scalars(0)["status"] == _lib.SDM_ERR_SOFTCORE


def test_nonequilibrium_schedule(cfg1):
    case, _ = cfg1
    al = S.AlchemicalState(**vars(case.alch))
    al.nonequilibrium, al.noneq_tmax, al.step_size, al.time = 1, 0.01, 0.001, 0.002
    al.alpha = 0.2
    al.m_lambda1, al.b_lambda1, al.m_lambda2, al.b_lambda2 = 0.3, 0.0, 0.5, 0.1
    al.m_u0, al.b_u0, al.m_w0, al.b_w0 = 1.0, 0.5, 0.2, 0.0
    ora = S.AlchemicalState(**vars(al))
    ctx = SDMContext(case.system, case.displacement, pair_mode=_lib.PAIR_ALLPAIRS)
    ctx.set_positions(0, case.positions)
    ctx.set_alchemical(0, al)
    for _ in range(3):
        ref = O.sdm_eval(case.system, ora, case.displacement, case.positions)
        ctx.eval()
        sc = ctx.scalars(0)
        assert abs(sc["sp"] - ref["sp"]) < 1e-9 and abs(sc["ebias"] - ref["ebias"]) < 1e-8
    got = ctx.get_alchemical(0)
    for k in ("lambdac", "lambda1", "lambda2", "u0", "w0coeff", "work_value", "time"):
        assert abs(getattr(got, k) - getattr(ora, k)) < 1e-9, k
    ctx.close()


def test_elementwise_kernel_interface_ops():
    """(B) entry points on device float4 buffers vs numpy (langevin.cl semantics)."""
    import torch
    L = _lib.lib()
    n = 100_003
    g = torch.Generator(device="cpu").manual_seed(3)
    mk = lambda: torch.randn(n, 4, generator=g, dtype=torch.float32).cuda()
    posq, displ, force, f1, f2 = mk(), mk(), mk(), mk(), mk()
    displ[:, 3] = 0
    save_f, save_x = torch.empty_like(posq), torch.empty_like(posq)
    st = torch.cuda.current_stream().cuda_stream
    p0, fo0 = posq.clone(), force.clone()
    _lib.check(L.sdm_k_save_state1(st, n, posq.data_ptr(), force.data_ptr(), save_f.data_ptr(), save_x.data_ptr()))
    _lib.check(L.sdm_k_make_state2(st, n, posq.data_ptr(), displ.data_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(save_f, fo0) and torch.equal(save_x, p0)
    assert torch.equal(posq, p0 + displ)
    s2 = torch.empty_like(force)
    _lib.check(L.sdm_k_save_state2(st, n, force.data_ptr(), s2.data_ptr()))
    _lib.check(L.sdm_k_restore_state1(st, n, posq.data_ptr(), save_x.data_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(s2, fo0) and torch.equal(posq, p0)
    sp = np.float32(0.37)
    _lib.check(L.sdm_k_hybrid_force(st, n, f1.data_ptr(), f2.data_ptr(), force.data_ptr(), float(sp)))
    torch.cuda.synchronize()
    want = (np.float32(1.0) - sp) * f1.cpu().numpy() + sp * f2.cpu().numpy() + fo0.cpu().numpy()
    assert np.allclose(force.cpu().numpy(), want, rtol=1e-6, atol=1e-6)


def test_langevin_kernel_interface_ops():
    """(B) integrateLangevinPart1 / Part2 (langevin.cl:7-69, single-precision form) on device float4
    buffers, bit for bit against the same float32 expressions in numpy; parameters from
    sdm_langevin_params (OpenCLSDMKernels.cpp:331-336)."""
    import ctypes as C
    import torch
    L = _lib.lib()
    vs, fs, ns = C.c_double(), C.c_double(), C.c_double()
    _lib.check(L.sdm_langevin_params(300.0, 0.5, 0.001, C.byref(vs), C.byref(fs), C.byref(ns)))
    kT = 1.380658e-23 * 6.0221367e23 / 1000.0 * 300.0
    assert vs.value == np.exp(-0.001 * 0.5) and abs(fs.value - (1 - vs.value) / 0.5) < 1e-18
    assert abs(ns.value - np.sqrt(kT * (1 - vs.value ** 2))) < 1e-15
    _lib.check(L.sdm_langevin_params(300.0, 0.0, 0.002, C.byref(vs), C.byref(fs), C.byref(ns)))
    assert fs.value == 0.002 and ns.value == 0.0          # friction == 0: fscale = dt, no noise
    _lib.check(L.sdm_langevin_params(300.0, 0.5, 0.001, C.byref(vs), C.byref(fs), C.byref(ns)))

    n, off = 50_001, 17
    rng = np.random.default_rng(4)
    velm = rng.normal(size=(n, 4)).astype(np.float32)
    velm[:, 3] = 1.0 / rng.uniform(1.0, 16.0, n).astype(np.float32)
    velm[::97, 3] = 0.0                                     # massless particles are left alone
    force = rng.normal(scale=300.0, size=(n, 4)).astype(np.float32)
    posq = rng.uniform(0, 5, size=(n, 4)).astype(np.float32)
    rnd = rng.normal(size=(n + off, 4)).astype(np.float32)
    d_velm, d_force, d_posq, d_rnd = (torch.from_numpy(a).cuda() for a in (velm, force, posq, rnd))
    d_delta = torch.full((n, 4), 7.0, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    f32 = np.float32
    vscale, fscale, noise, dt = f32(vs.value), f32(fs.value), f32(ns.value), f32(0.001)
    _lib.check(L.sdm_k_langevin_part1(st, n, d_velm.data_ptr(), d_force.data_ptr(), d_delta.data_ptr(),
                                      float(vscale), float(fscale), float(noise), float(dt), d_rnd.data_ptr(), off))
    torch.cuda.synchronize()
    w = velm[:, 3:4]
    live = (w != 0)[:, 0]
    v1 = velm.copy()
    v1[:, :3] = (vscale * velm[:, :3] + (fscale * w) * force[:, :3]) + (noise * np.sqrt(w)) * rnd[off:off + n, :3]
    v1[~live] = velm[~live]
    delta = np.full((n, 4), 7.0, np.float32)
    delta[live] = dt * v1[live]
    assert np.array_equal(d_velm.cpu().numpy(), v1)
    assert np.array_equal(d_delta.cpu().numpy(), delta)

    _lib.check(L.sdm_k_langevin_part2(st, n, d_posq.data_ptr(), d_delta.data_ptr(), d_velm.data_ptr(), float(dt)))
    torch.cuda.synchronize()
    inv = f32(1.0) / dt
    corr = (f32(1.0) - inv * dt) / dt
    p2, v2 = posq.copy(), v1.copy()
    p2[live, :3] = posq[live, :3] + delta[live, :3]
    v2[live, :3] = inv * delta[live, :3] + corr * delta[live, :3]
    assert np.array_equal(d_posq.cpu().numpy(), p2)
    assert np.array_equal(d_velm.cpu().numpy(), v2)
