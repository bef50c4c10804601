"""CPU tests of the reciprocal-space PME oracle (oracle/pme.py: OpenMM 7.3 ReferencePME restated in numpy)
against the exact structure-factor sum of the same Ewald split."""
import numpy as np

from oracle import pme as P
from test_oracle import K, small_ewald_case


def exact_reciprocal(q, pos, box, alpha, kmax):
    """Energy (with self term) and forces of the plain Ewald reciprocal sum."""
    V = box.prod()
    n = [np.arange(-kmax, kmax + 1)] * 3
    kx, ky, kz = np.meshgrid(*n, indexing="ij")
    kv = 2 * np.pi * np.stack([kx.ravel() / box[0], ky.ravel() / box[1], kz.ravel() / box[2]], 1)
    k2 = (kv ** 2).sum(1)
    keep = k2 > 0
    kv, k2 = kv[keep], k2[keep]
    phase = pos @ kv.T
    c, s = np.cos(phase), np.sin(phase)
    sre, sim = (q[:, None] * c).sum(0), (q[:, None] * s).sum(0)
    a = 2 * np.pi / V * K * np.exp(-k2 / (4 * alpha ** 2)) / k2
    e = (a * (sre ** 2 + sim ** 2)).sum() - K * alpha / np.sqrt(np.pi) * (q ** 2).sum()
    # F_i = -dE/dr_i = 2 q_i sum_k a_k k (sin(k r_i) S_re - cos(k r_i) S_im)
    f = 2 * q[:, None] * ((a * (s * sre - c * sim)) @ kv)
    return e, f


def test_bspline_weights_are_a_partition_of_unity_with_zero_sum_derivatives():
    w = np.linspace(0, 1, 11, endpoint=False)
    th, dth = P.bsplines(w)
    assert np.abs(th.sum(1) - 1.0).max() < 1e-14 and np.abs(dth.sum(1)).max() < 1e-14
    assert np.all(th >= 0)
    th0, _ = P.bsplines(np.zeros(1))
    assert np.allclose(th0[0], [1 / 24, 11 / 24, 11 / 24, 1 / 24, 0.0])      # cardinal B-spline of order 5 at the knots


def test_grid_rule_and_fft_sizes():
    box = np.array([5.64590, 6.32926, 5.79653])                               # cfg2
    alpha = np.sqrt(-np.log(2 * 5e-4)) / 1.0
    g = P.grid_size_rule(alpha, box, 5e-4)
    assert g == [46, 51, 47]
    assert [P.fft_friendly(x) for x in g] == [48, 54, 48]


def test_pme_reciprocal_converges_to_the_exact_sum():
    sysd, pos = small_ewald_case(n_mol=24, seed=7)
    alpha = 3.0
    q, box = sysd.charge, sysd.box
    e_ref, f_ref = exact_reciprocal(q, pos, box, alpha, kmax=16)
    h = 1e-5                                                                  # the exact forces are the gradient
    p, m = pos.copy(), pos.copy()
    p[3, 1] += h; m[3, 1] -= h
    fd = -(exact_reciprocal(q, p, box, alpha, 16)[0] - exact_reciprocal(q, m, box, alpha, 16)[0]) / (2 * h)
    assert abs(fd - f_ref[3, 1]) <= 1e-6 * max(1.0, abs(fd))
    errs = []
    for grid in ([20, 22, 20], [40, 44, 42]):
        e, f = P.reciprocal(q, pos, box, alpha, grid)
        errs.append((abs(e - e_ref) / abs(e_ref), np.sqrt(((f - f_ref) ** 2).sum() / (f_ref ** 2).sum())))
    assert errs[0][0] < 2e-3 and errs[0][1] < 2e-2                            # coarse grid: PME's interpolation error
    assert errs[1][0] < 2e-5 and errs[1][1] < 5e-4                            # finer grid: order-5 convergence
    assert errs[1][1] < errs[0][1] / 10


def test_pme_forces_are_the_gradient_of_the_pme_energy_up_to_interpolation():
    sysd, pos = small_ewald_case(n_mol=16, seed=2)
    q, box, alpha, grid = sysd.charge, sysd.box, 3.0, [36, 40, 36]
    e, f = P.reciprocal(q, pos, box, alpha, grid)
    assert np.abs(f.sum(0)).max() < 1e-2 * np.abs(f).max()                    # momentum is conserved only approximately
    h = 1e-5
    for i, d in ((0, 0), (7, 2)):
        p, m = pos.copy(), pos.copy()
        p[i, d] += h; m[i, d] -= h
        fd = -(P.reciprocal(q, p, box, alpha, grid)[0] - P.reciprocal(q, m, box, alpha, grid)[0]) / (2 * h)
        assert abs(fd - f[i, d]) <= 2e-3 * np.abs(f).max()
