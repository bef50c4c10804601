"""The CPU oracle against every pinned value we have (the reference ships no tests or golden
vectors -- "parity unpinned"): the scalar known answers of SURVEY.md
Appendix A.4, the fixture energies of Appendix C, and an independent O(N^2) numpy
re-derivation of the pair arithmetic."""
import numpy as np
import pytest

from openmm_sdm_plugin_b200 import system as S
from oracle import oracle as O

K = 138.935456
UMAX, UB, A = 418.4, 209.2, 0.0625


@pytest.mark.parametrize("u,m,usc,fp", [
    (100.0, 1, 100.0, 1.0), (209.2, 2, 209.2, 1.0), (-5.0, 2, -5.0, 1.0),
    (250, 1, 249.87116697612, 0.990550911516698),
    (250, 2, 230.602129698129, 0.268201879720804),
    (500, 1, 460.743251863468, 0.638555192285094),
    (500, 2, 253.884299107405, 0.0419471010006756),
    (1000, 1, 608.930208674602, 0.0872526329072919),
    (1000, 2, 266.015525174781, 0.0151877926132863),
    (1e5, 2, 318.605914686799, 9.51826991569913e-05),
    (1e8, 2, 369.594730216881, 5.38904370080525e-08),
])
def test_softcore_known_answers(u, m, usc, fp):
    got = O.softcore(m, u, UMAX, A, UB)
    assert got[0] == pytest.approx(usc, rel=1e-13)
    assert got[1] == pytest.approx(fp, rel=1e-12)


def test_softcore_none_and_unknown():
    assert O.softcore(0, 1e6, UMAX, A, UB) == (1e6, 1.0)
    assert O.softcore(9, 1.0, UMAX, A, UB) == (1.0, 1.0)      # u <= ub returns before the switch
    with pytest.raises(ValueError):
        O.softcore(9, 1e3, UMAX, A, UB)


@pytest.mark.parametrize("B,l1,l2,alpha,u0,w0,ebias,bfp", [
    (3.60861626248879, 0.025, 0.025, 0.0, 0.0, 0.0, 0.0902154065622199, 0.025),
    (100.0, 0.0, 0.5, 0.0239005736137667, 460.24, 0.0, 230.123813041422, 9.11255723303808e-05),
    (500.0, 0.2, 0.5, 0.0239005736137667, 460.24, 0.0, 254.102952937252, 0.416351856063777),
    (-50.0, 0.1, 0.5, 0.0239005736137667, 460.24, 1.0, 180.096084609067, 0.100002022200126),
])
def test_ilogistic_known_answers(B, l1, l2, alpha, u0, w0, ebias, bfp):
    al = S.AlchemicalState(bias_method=S.ILOGISTIC, lambda1=l1, lambda2=l2, alpha=alpha, u0=u0, w0coeff=w0)
    e, b = O.bias(al, B)
    assert e == pytest.approx(ebias, rel=1e-12)
    assert b == pytest.approx(bfp, rel=1e-11)


def test_linear_and_quadratic_bias():
    al = S.AlchemicalState(bias_method=S.LINEAR, lambdac=0.3)
    assert O.bias(al, 10.0) == (3.0, 0.3)
    al = S.AlchemicalState(bias_method=S.QUADRATIC, gammac=0.02, wbcoeff=0.5, w0coeff=1.0)
    e, b = O.bias(al, 10.0)
    assert e == pytest.approx(0.5 * 0.02 * 100 + 5.0 + 1.0) and b == pytest.approx(0.7)


def test_nonequilibrium_schedule_and_work():
    """ReferenceSDMKernels.cpp:231-245,289-302 by hand."""
    al = S.AlchemicalState(bias_method=S.ILOGISTIC, nonequilibrium=1, noneq_tmax=2.0, time=0.5,
                           step_size=0.001, alpha=0.1, m_lambda1=0.2, b_lambda1=0.1, m_lambda2=0.4,
                           b_lambda2=0.3, m_u0=5.0, b_u0=1.0, m_w0=0.7, b_w0=0.2)
    B = 4.0
    e, b = O.bias(al, B)
    lam = 0.25
    l1, l2, u0, w0 = 0.2 * lam + 0.1, 0.4 * lam + 0.3, 5.0 * lam + 1.0, 0.7 * lam + 0.2
    assert (al.lambdac, al.lambda1, al.lambda2, al.u0, al.w0coeff) == pytest.approx((lam, l1, l2, u0, w0))
    ee = 1 + np.exp(-0.1 * (B - u0))
    assert e == pytest.approx((l2 - l1) / 0.1 * np.log(ee) + l2 * B + w0)
    assert b == pytest.approx((l2 - l1) / ee + l1)
    dw = (-np.log(ee) / 0.1) * 0.2 + (B + np.log(ee) / 0.1) * 0.4 + (l2 - l1) * np.exp(-0.1 * (B - u0)) / ee * 5.0 + 0.7
    assert al.work_value == pytest.approx(0.001 / 2.0 * dw)


# ---- fixtures -------------------------------------------------------------
def test_cfg1_fixture_known_answers():
    c = S.cfg1()
    r = O.sdm_eval(c.system, c.alch, c.displacement, c.positions)
    assert r["n_pairs1"] == 25061 and r["n_pairs2"] == 25061
    assert r["E1_pair"] == pytest.approx(-700.9267350534, abs=1e-8)
    assert r["E1_exc"] == pytest.approx(-161.9778775080, abs=1e-8)
    assert r["E1"] == pytest.approx(-862.9046125614, abs=1e-8)
    assert r["E2"] == pytest.approx(-859.2959962989, abs=1e-8)
    assert r["u"] == pytest.approx(3.6086162625, abs=1e-8)
    assert r["u_sc"] == r["u"] and r["fp"] == 1.0            # below ub
    assert r["ebias"] == pytest.approx(0.0902154065622199, abs=1e-9)
    assert r["sp"] == 0.025
    assert r["pot_energy"] == pytest.approx(r["E1"] + r["ebias"])
    assert np.allclose(r["forces"], 0.025 * r["f2"] + 0.975 * r["f1"], rtol=0, atol=1e-9)
    assert (np.count_nonzero(c.displacement.any(axis=1))) == 38


def test_cfg2_fixture_known_answers():
    c = S.cfg2()
    r = O.sdm_eval(c.system, c.alch, c.displacement, c.positions, nthreads=O.max_threads())
    assert r["n_pairs1"] == 4197871 and r["n_pairs2"] == 4197834
    assert r["E1_pair"] == pytest.approx(-274471.6212766, abs=2e-6)
    assert r["E2"] - r["E1"] == pytest.approx(-6.6429627853, abs=2e-7)
    assert r["sp"] == 0.5
    assert (np.count_nonzero(c.displacement.any(axis=1))) == 38
    # the position buffer is restored (RestoreState1)
    assert np.array_equal(c.positions, S.cfg2().positions)


# ---- independent numpy re-derivation ---------------------------------------------------------------
def brute_force(system, pos):
    n = system.n_atoms
    d = pos[:, None, :] - pos[None, :, :]
    if system.method == S.CUTOFF_PERIODIC:
        d -= np.floor(d / system.box + 0.5) * system.box
    r2 = (d ** 2).sum(-1)
    iu = np.triu(np.ones((n, n), bool), 1)
    for a, b in system.exclusions:
        iu[min(a, b), max(a, b)] = False
    if system.method != S.NOCUTOFF:
        iu &= r2 <= system.cutoff ** 2
    i, j = np.nonzero(iu)
    r = np.sqrt(r2[i, j])
    sig = 0.5 * (system.sigma[i] + system.sigma[j])
    eps = np.sqrt(system.epsilon[i] * system.epsilon[j])
    qq = K * system.charge[i] * system.charge[j]
    s6 = (sig / r) ** 6
    e = 4 * eps * (s6 * s6 - s6)
    dedr = 4 * eps * (12 * s6 * s6 - 6 * s6)
    if system.method == S.NOCUTOFF:
        e += qq / r
        dedr += qq / r
    else:
        rc, es = system.cutoff, system.eps_rf
        krf, crf = (es - 1) / (2 * es + 1) / rc ** 3, 3 * es / (2 * es + 1) / rc
        e += qq * (1 / r + krf * r * r - crf)
        dedr += qq * (1 / r - 2 * krf * r * r)
    f = np.zeros((n, 3))
    fv = (dedr / r ** 2)[:, None] * d[i, j]
    np.add.at(f, i, fv)
    np.add.at(f, j, -fv)
    e14 = 0.0
    for (a, b), (q14, s14, e14p) in zip(system.exception_pairs, system.exception_params):
        dd = pos[a] - pos[b]
        rr = np.sqrt((dd ** 2).sum())
        s6 = (s14 / rr) ** 6
        e14 += 4 * e14p * (s6 * s6 - s6) + K * q14 / rr
        ff = (4 * e14p * (12 * s6 * s6 - 6 * s6) + K * q14 / rr) / rr ** 2 * dd
        f[a] += ff
        f[b] -= ff
    return e.sum(), e14, f, np.stack([i, j], 1).astype(np.int32)


@pytest.mark.parametrize("method,cutoff", [(S.CUTOFF_NONPERIODIC, 15.0), (S.CUTOFF_NONPERIODIC, 1.1), (S.NOCUTOFF, 0.0)])
def test_nonbonded_matches_numpy_cfg1(method, cutoff):
    c = S.cfg1()
    c.system.method, c.system.cutoff = method, cutoff
    e, e14, f, pairs = brute_force(c.system, c.positions)
    r = O.nonbonded(c.system, c.positions, want_pairs=True)
    assert r["E_pair"] == pytest.approx(e, rel=1e-12)
    assert r["E_exc"] == pytest.approx(e14, rel=1e-12)
    assert np.allclose(r["forces"], f, rtol=1e-10, atol=1e-8)
    assert np.array_equal(r["pairs"], pairs)


def test_nonbonded_matches_numpy_periodic_with_threads():
    c = S.synthetic_case(1500, 30, seed=4, protein_atoms=150)
    e, e14, f, pairs = brute_force(c.system, c.positions)
    for nt in (1, 3):
        r = O.nonbonded(c.system, c.positions, want_pairs=True, nthreads=nt)
        assert r["E_pair"] == pytest.approx(e, rel=1e-11)
        assert r["E_exc"] == pytest.approx(e14, rel=1e-12)
        assert np.allclose(r["forces"], f, rtol=1e-9, atol=1e-7)
        assert np.array_equal(r["pairs"], pairs)


def test_dispersion_correction_formula():
    """OpenMM NonbondedForceImpl::calcDispersionCorrection for a two-class system by hand."""
    n1, n2 = 3, 2
    sysd = S.NonbondedSystem(np.zeros(5), np.array([0.3] * n1 + [0.2] * n2), np.array([0.5] * n1 + [0.1] * n2),
                             np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, 3)), method=S.CUTOFF_PERIODIC,
                             cutoff=1.0, box=np.array([3.0, 3.0, 3.0]))
    s1 = n1 * (n1 + 1) / 2 * 0.5 * 0.3 ** 12 + n2 * (n2 + 1) / 2 * 0.1 * 0.2 ** 12 + n1 * n2 * np.sqrt(0.05) * 0.25 ** 12
    s2 = n1 * (n1 + 1) / 2 * 0.5 * 0.3 ** 6 + n2 * (n2 + 1) / 2 * 0.1 * 0.2 ** 6 + n1 * n2 * np.sqrt(0.05) * 0.25 ** 6
    ni = 5 * 6 / 2
    want = 8 * 25 * np.pi * (s1 / ni / 9 - s2 / ni / 3)
    assert O.dispersion_coefficient(sysd) == pytest.approx(want, rel=1e-12)
    r = O.nonbonded(sysd, np.random.default_rng(0).uniform(0, 3, (5, 3)))
    assert r["E_disp"] == pytest.approx(want / 27.0, rel=1e-12)
    sysd.use_dispersion_correction = False
    assert O.nonbonded(sysd, np.zeros((5, 3)) + np.arange(5)[:, None] * 0.5)["E_disp"] == 0.0


def test_box_smaller_than_twice_cutoff_is_an_error():
    c = S.synthetic_case(600, 30, seed=4, protein_atoms=0)
    c.system.cutoff = 1.2
    with pytest.raises(RuntimeError):
        O.nonbonded(c.system, c.positions)


# ---- physics invariants of the restated step -----------------------------------------------------
def test_zero_map_and_lambda_limits():
    c = S.cfg1()
    zero = np.zeros_like(c.displacement)
    r = O.sdm_eval(c.system, c.alch, zero, c.positions)
    # sp*F2 + (1-sp)*F1 with F2 == F1 reproduces F1 up to rounding of the reference formula
    assert r["u"] == 0.0 and np.allclose(r["forces"], r["f1"], rtol=1e-14, atol=1e-12)
    lin0 = S.AlchemicalState(bias_method=S.LINEAR, lambdac=0.0)
    lin1 = S.AlchemicalState(bias_method=S.LINEAR, lambdac=1.0)
    fb = np.random.default_rng(1).normal(size=c.positions.shape)
    r0 = O.sdm_eval(c.system, lin0, c.displacement, c.positions, fb=fb, eb=3.0)
    r1 = O.sdm_eval(c.system, lin1, c.displacement, c.positions, fb=fb, eb=3.0)
    assert np.allclose(r0["forces"], r0["f1"] + fb, atol=1e-12)
    assert np.allclose(r1["forces"], r1["f2"] + fb, atol=1e-12)
    assert r1["pot_energy"] == pytest.approx(r1["E1"] + r1["u"] + 3.0)


def test_u_only_depends_on_moved_pairs():
    """Appendix C: u from pairs touching a displaced atom equals E2 - E1."""
    c = S.cfg1()
    r = O.sdm_eval(c.system, c.alch, c.displacement, c.positions)
    moved = c.displacement.any(axis=1)
    e1, _, _, p1 = brute_force(c.system, c.positions)
    e2, _, _, p2 = brute_force(c.system, c.positions + c.displacement)
    assert e2 - e1 == pytest.approx(r["u"], abs=1e-9)
    same = lambda p: (c.displacement[p[:, 0]] == c.displacement[p[:, 1]]).all(axis=1)
    assert (~same(p1)).sum() == 7656 and (moved[p1[:, 0]] | moved[p1[:, 1]]).sum() >= 7656


# ---- NonbondedForce::Ewald / ::PME, direct-space part (sdm_oracle.c: ewald_alpha > 0) ----------------------
def ewald_direct_numpy(system, pos, alpha):
    """Pair sum qq*erfc(alpha r)/r inside the cutoff + LJ, minus qq*erf(alpha r)/r of the excluded pairs
    (minimum image) -- ReferenceLJCoulombIxn::calculateEwaldIxn with includeDirect only."""
    from scipy.special import erf, erfc
    n = system.n_atoms
    d = pos[:, None, :] - pos[None, :, :]
    d -= np.floor(d / system.box + 0.5) * system.box
    r2 = (d ** 2).sum(-1)
    iu = np.triu(np.ones((n, n), bool), 1)
    ex = np.zeros((n, n), bool)
    for a, b in system.exclusions:
        ex[min(a, b), max(a, b)] = True
    inc = iu & ~ex & (r2 <= system.cutoff ** 2)
    i, j = np.nonzero(inc)
    r = np.sqrt(r2[i, j])
    sig = 0.5 * (system.sigma[i] + system.sigma[j])
    eps = np.sqrt(system.epsilon[i] * system.epsilon[j])
    qq = K * system.charge[i] * system.charge[j]
    s6 = (sig / r) ** 6
    e_pair = (4 * eps * (s6 * s6 - s6) + qq * erfc(alpha * r) / r).sum()
    dedr = 4 * eps * (12 * s6 * s6 - 6 * s6) + qq * (erfc(alpha * r) + 2 * alpha * r * np.exp(-(alpha * r) ** 2) / np.sqrt(np.pi)) / r
    f = np.zeros((n, 3))
    fv = (dedr / r ** 2)[:, None] * d[i, j]
    np.add.at(f, i, fv)
    np.add.at(f, j, -fv)
    i, j = np.nonzero(ex)
    r = np.sqrt(r2[i, j])
    qq = K * system.charge[i] * system.charge[j]
    e_excl = -(qq * erf(alpha * r) / r).sum()
    dedr = qq * (erf(alpha * r) - 2 * alpha * r * np.exp(-(alpha * r) ** 2) / np.sqrt(np.pi)) / r
    fv = (dedr / r ** 2)[:, None] * d[i, j]
    np.add.at(f, i, -fv)
    np.add.at(f, j, fv)
    return e_pair, e_excl, f


def ewald_reciprocal_numpy(q, pos, box, alpha, kmax):
    """Reciprocal-space Ewald energy + self energy by the plain structure-factor sum (what OpenMM books under
    includeReciprocal): (2 pi / V) K sum_k exp(-k^2 / 4 alpha^2) / k^2 |S(k)|^2 - K alpha / sqrt(pi) sum q^2."""
    V = box.prod()
    e = 0.0
    n = [np.arange(-kmax, kmax + 1)] * 3
    kx, ky, kz = np.meshgrid(*n, indexing="ij")
    kv = 2 * np.pi * np.stack([kx.ravel() / box[0], ky.ravel() / box[1], kz.ravel() / box[2]], 1)
    k2 = (kv ** 2).sum(1)
    keep = k2 > 0
    kv, k2 = kv[keep], k2[keep]
    phase = pos @ kv.T
    sre, sim = (q[:, None] * np.cos(phase)).sum(0), (q[:, None] * np.sin(phase)).sum(0)
    e = 2 * np.pi / V * K * (np.exp(-k2 / (4 * alpha ** 2)) / k2 * (sre ** 2 + sim ** 2)).sum()
    return e - K * alpha / np.sqrt(np.pi) * (q ** 2).sum()


def small_ewald_case(n_mol=40, seed=4):
    """Neutral box of rigid three-site molecules (intramolecular exclusions) small enough for numpy."""
    rng = np.random.default_rng(seed)
    box = np.array([2.3, 2.5, 2.4])
    centres = rng.uniform(0, 1, (n_mol, 3)) * box
    pos = np.concatenate([centres[:, None, :] + rng.normal(scale=0.06, size=(n_mol, 3, 3))], 0).reshape(-1, 3)
    q = np.tile([-0.834, 0.417, 0.417], n_mol)
    sig = np.tile([0.315, 0.1, 0.1], n_mol)
    eps = np.tile([0.636, 0.0, 0.0], n_mol)
    excl = np.array([[3 * m + a, 3 * m + b] for m in range(n_mol) for a, b in ((0, 1), (0, 2), (1, 2))], np.int32)
    sysd = S.NonbondedSystem(q, sig, eps, excl, np.zeros((0, 2), np.int32), np.zeros((0, 3)), method=S.PME,
                             cutoff=1.0, box=box, use_dispersion_correction=False)
    return sysd, pos


def test_ewald_direct_space_matches_numpy():
    sysd, pos = small_ewald_case()
    alpha = sysd.ewald_alpha_effective()
    assert abs(alpha - np.sqrt(-np.log(2 * 5e-4)) / 1.0) < 1e-15          # OpenMM's rule
    out = O.nonbonded(sysd, pos)
    e_pair, e_excl, f = ewald_direct_numpy(sysd, pos, alpha)
    assert abs(out["E_pair"] - e_pair) <= 1e-10 * abs(e_pair)
    assert abs(out["E_exc"] - e_excl) <= 1e-10 * abs(e_excl)
    assert np.abs(out["forces"] - f).max() <= 1e-9 * np.abs(f).max()


def test_ewald_split_is_independent_of_alpha():
    """Direct space (oracle) + reciprocal space and self energy (independent numpy structure-factor sum) is the
    Coulomb energy of the periodic system, whatever the splitting parameter: the erfc pair term and the erf
    correction of the excluded pairs are the right counterpart of the reciprocal sum."""
    sysd, pos = small_ewald_case(n_mol=24, seed=7)
    sysd.epsilon[:] = 0.0                                  # Coulomb only
    tot = []
    for alpha in (4.6, 5.4):                               # erfc(alpha * rc) < 1e-10: the cutoff truncates nothing
        sysd.ewald_alpha = alpha
        out = O.nonbonded(sysd, pos)
        rec = ewald_reciprocal_numpy(sysd.charge, pos, sysd.box, alpha, kmax=22)
        tot.append(out["E_pair"] + out["E_exc"] + rec)
    assert abs(tot[0] - tot[1]) <= 1e-7 * abs(tot[0]), tot


def test_ewald_dual_state_eval_runs_through_the_reference_sequence():
    """orc_sdm_eval with method = PME: u = E2 - E1 from two full direct-space evaluations."""
    sysd, pos = small_ewald_case()
    disp = np.zeros_like(pos)
    disp[:6] = (0.4, 0.0, 0.0)
    res = O.sdm_eval(sysd, S.AlchemicalState(lambdac=0.5), disp, pos)
    e1 = O.nonbonded(sysd, pos)["E"]
    e2 = O.nonbonded(sysd, pos + disp)["E"]
    assert abs(res["E1"] - e1) <= 1e-12 * abs(e1) and abs(res["u"] - (e2 - e1)) <= 1e-9 * max(1.0, abs(e2 - e1))


def test_geometric_combining_rule_against_the_literal_expression():
    """createSystem(OPLS=True), desmonddmsfile75.py:780-810: NonbondedForce keeps the charges, the Lennard-Jones part is
    the CustomNonbondedForce expression 4 eps12 ((s12/r)^12 - (s12/r)^6) with geometric means; no cutoff here, so the
    total is a plain double loop."""
    import copy
    case = S.cfg1()
    sysd = copy.copy(case.system)
    sysd.method = S.NOCUTOFF
    sysd.lj_geometric = True
    sysd.exception_pairs = np.zeros((0, 2), np.int32)
    sysd.exception_params = np.zeros((0, 3))
    pos = case.positions
    out = O.nonbonded(sysd, pos)
    n = sysd.n_atoms
    excl = {(min(a, b), max(a, b)) for a, b in sysd.exclusions.tolist()}
    e = 0.0
    f = np.zeros_like(pos)
    for i in range(n):
        for j in range(i + 1, n):
            if (i, j) in excl:
                continue
            d = pos[i] - pos[j]
            r = np.sqrt(d @ d)
            s12 = np.sqrt(sysd.sigma[i] * sysd.sigma[j])
            e12 = np.sqrt(sysd.epsilon[i] * sysd.epsilon[j])
            x6 = (s12 / r) ** 6
            qq = 138.935456 * sysd.charge[i] * sysd.charge[j]
            e += 4 * e12 * (x6 * x6 - x6) + qq / r
            dedr = -4 * e12 * (12 * x6 * x6 - 6 * x6) / r - qq / r ** 2
            f[i] -= dedr * d / r
            f[j] += dedr * d / r
    assert out["E"] == pytest.approx(e, rel=1e-12)
    assert np.abs(out["forces"] - f).max() <= 1e-10 * np.abs(f).max()
    lb = O.nonbonded(case.system, pos)
    assert abs(lb["E"] - out["E"]) > 1e-3                      # the rule matters on this fixture
