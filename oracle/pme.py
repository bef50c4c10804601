"""TEST INFRASTRUCTURE (oracle): reciprocal-space smooth particle-mesh Ewald in numpy -- the checker of
csrc/kernels_pme.cu.

What is restated: OpenMM 7.3 ReferencePME.cpp (pme_init / pme_exec: update_grid_index_and_fraction,
update_bsplines, grid_spread_charge, reciprocal_convolution, grid_interpolate_force) as called by
ReferenceLJCoulombIxn::calculateEwaldIxn for NonbondedForce::PME with includeReciprocal, plus the self
energy booked there; the grid size rule of NonbondedForceImpl::calcPMEParameters
(ceil(2 alpha L / (3 tol^(1/5)))).  B-splines of order 5, orthorhombic box, double precision.
The reference plugin calls this through Context::calcForcesAndEnergy (LangevinIntegratorSDM.cpp:160,168) when
the System was built with nonbondedMethod=PME (example/test_explicit.py:64).  OpenMM is not available here:
parity unpinned; tests/test_pme_oracle.py holds this restatement to the exact structure-factor sum.
Only tests/ may import this module."""
import numpy as np

K_COULOMB = 138.935456
ORDER = 5


def grid_size_rule(alpha, box, tol):
    """NonbondedForceImpl::calcPMEParameters: ceil(2 alpha L / (3 tol^(1/5))), at least 6."""
    return [max(int(np.ceil(2 * alpha * box[d] / (3 * tol ** 0.2))), 6) for d in range(3)]


def fft_friendly(n):
    """Next size with prime factors 2, 3, 5, 7 only (what the device path rounds the rule's sizes up to)."""
    while True:
        m = n
        for p in (2, 3, 5, 7):
            while m % p == 0:
                m //= p
        if m == 1:
            return n
        n += 1


def bsplines(w, order=ORDER):
    """update_bsplines for an array of fractions w: weights [len(w), order] and their derivatives."""
    w = np.asarray(w, dtype=np.float64)
    data = np.zeros(w.shape + (order,))
    data[..., order - 1] = 0.0
    data[..., 1] = w
    data[..., 0] = 1.0 - w
    for k in range(3, order):
        div = 1.0 / (k - 1.0)
        data[..., k - 1] = div * w * data[..., k - 2]
        for l in range(1, k - 1):
            data[..., k - l - 1] = div * ((w + l) * data[..., k - l - 2] + (k - l - w) * data[..., k - l - 1])
        data[..., 0] = div * (1.0 - w) * data[..., 0]
    ddata = np.zeros_like(data)
    ddata[..., 0] = -data[..., 0]
    for k in range(1, order):
        ddata[..., k] = data[..., k - 1] - data[..., k]
    div = 1.0 / (order - 1)
    data[..., order - 1] = div * w * data[..., order - 2]
    for l in range(1, order - 1):
        data[..., order - l - 1] = div * ((w + l) * data[..., order - l - 2] + (order - l - w) * data[..., order - l - 1])
    data[..., 0] = div * (1.0 - w) * data[..., 0]
    return data, ddata


def moduli(K, order=ORDER):
    """pme_calculate_bsplines_moduli: |sum_j M_n(j+1) exp(2 pi i m j / K)|^2, zeros patched by their neighbours."""
    data, _ = bsplines(np.zeros(1), order)
    b = np.zeros(K)
    b[1:order + 1] = data[0][:K - 1] if order + 1 > K else data[0]
    m = np.arange(K)
    arg = 2 * np.pi * np.outer(m, np.arange(K)) / K
    mod = (b * np.cos(arg)).sum(1) ** 2 + (b * np.sin(arg)).sum(1) ** 2
    for i in range(K):
        if mod[i] < 1e-7:
            mod[i] = 0.5 * (mod[(i - 1) % K] + mod[(i + 1) % K])
    return mod


def reciprocal(charge, positions, box, alpha, grid, order=ORDER, self_energy=True):
    """Reciprocal-space energy (kJ/mol; with the self energy -K alpha/sqrt(pi) sum q^2 when self_energy) and
    forces [n,3] (kJ/mol/nm)."""
    q = np.asarray(charge, dtype=np.float64)
    pos = np.asarray(positions, dtype=np.float64)
    box = np.asarray(box, dtype=np.float64)
    n = len(q)
    K = [int(g) for g in grid]
    frac = pos / box
    frac -= np.floor(frac)
    t = frac * np.array(K)
    ti = np.floor(t).astype(np.int64)
    ti = np.minimum(ti, np.array(K) - 1)        # t == K after rounding
    w = t - ti
    th, dth = [], []
    for d in range(3):
        a, b = bsplines(w[:, d], order)
        th.append(a)
        dth.append(b)
    Q = np.zeros(K)
    idx = [(ti[:, d][:, None] + np.arange(order)[None, :]) % K[d] for d in range(3)]
    for i in range(n):
        Q[np.ix_(idx[0][i], idx[1][i], idx[2][i])] += q[i] * np.einsum("a,b,c->abc", th[0][i], th[1][i], th[2][i])
    FQ = np.fft.fftn(Q)
    mods = [moduli(K[d], order) for d in range(3)]
    mm = []
    for d in range(3):
        k = np.arange(K[d])
        mm.append(np.where(k < (K[d] + 1) // 2, k, k - K[d]) / box[d])
    m2 = mm[0][:, None, None] ** 2 + mm[1][None, :, None] ** 2 + mm[2][None, None, :] ** 2
    denom = m2 * mods[0][:, None, None] * mods[1][None, :, None] * mods[2][None, None, :]
    V = box.prod()
    with np.errstate(divide="ignore", invalid="ignore"):
        eterm = K_COULOMB / (np.pi * V) * np.exp(-(np.pi ** 2 / alpha ** 2) * m2) / denom
    eterm[0, 0, 0] = 0.0
    energy = 0.5 * (eterm * (FQ.real ** 2 + FQ.imag ** 2)).sum()
    phi = np.fft.ifftn(eterm * FQ).real * np.prod(K)          # unnormalised backward transform, like FFTW / cuFFT
    f = np.zeros((n, 3))
    for i in range(n):
        g = phi[np.ix_(idx[0][i], idx[1][i], idx[2][i])]
        f[i, 0] = -q[i] * np.einsum("a,b,c,abc->", dth[0][i], th[1][i], th[2][i], g) * K[0] / box[0]
        f[i, 1] = -q[i] * np.einsum("a,b,c,abc->", th[0][i], dth[1][i], th[2][i], g) * K[1] / box[1]
        f[i, 2] = -q[i] * np.einsum("a,b,c,abc->", th[0][i], th[1][i], dth[2][i], g) * K[2] / box[2]
    if self_energy:
        energy -= K_COULOMB * alpha / np.sqrt(np.pi) * (q ** 2).sum()
    return energy, f
