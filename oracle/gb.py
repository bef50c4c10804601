"""TEST INFRASTRUCTURE (oracle): the HCT generalized-Born + ACE surface-area model the reference's DMS reader adds
for implicit solvent (example/desmonddmsfile75.py:454-465: GBSAHCTForce(SA='ACE'), force group 2, so that both
SDM states are evaluated with it), restated with torch (float64); autograd gives the forces -- the checker of
csrc/kernels_gb.cu.

What is restated: the expression strings of OpenMM 7.3's simtk/openmm/app/internal/customgbforces.py
(GBSAHCTForce + _createEnergyTerms, no cutoff, kappa = 0), a third-party dependency that is absent here:
    I_i  = sum_{j != i} step(r+sr_j-or_i) * 0.5*(1/L - 1/U + 0.25*(r - sr_j^2/r)*(1/U^2 - 1/L^2) + 0.5*log(L/U)/r),
           U = r + sr_j, L = max(or_i, |r - sr_j|)                                (ParticlePairNoExclusions)
    B_i  = 1/(1/or_i - I_i)
    E    = sum_i -0.5*138.935485*(1/soluteDielectric - 1/solventDielectric)*q_i^2/B_i
         + sum_i 28.3919551*(radius_i + 0.14)^2*(radius_i/B_i)^6, radius_i = or_i + 0.009          (SA='ACE')
         + sum_{i<j} -138.935485*(1/soluteDielectric - 1/solventDielectric)*q_i*q_j/f,
           f = sqrt(r^2 + B_i*B_j*exp(-r^2/(4*B_i*B_j)))                          (ParticlePairNoExclusions)
with CustomGBForce's semantics: NoCutoff (plain distances, no periodic image), no exclusions, step(x) = 0 for
x < 0 else 1, and the chain rule through the computed values I and B.  OpenMM is not available in this image and
neither shipped fixture carries an `hct` table: parity unpinned for these expressions; what this oracle checks is
that the device kernels evaluate them and their exact gradient.  per-particle parameters are the ones the
CustomGBForce sees: charge, or (offset radius, nm), sr (scaled offset radius, nm).
Only tests/ may import this module."""
import numpy as np
import torch

GB_COULOMB = 138.935485      # the constant customgbforces.py writes into its expressions
ACE_COEFF = 28.3919551
ACE_PROBE = 0.14
HCT_OFFSET = 0.009


def hct_energy(pos, charge, offset_radius, scaled_radius, solute_dielectric=1.0, solvent_dielectric=78.5, sa_ace=True):
    """Energy (torch scalar) of positions `pos` (torch [n,3], float64)."""
    q = torch.as_tensor(np.asarray(charge, dtype=np.float64))
    o = torch.as_tensor(np.asarray(offset_radius, dtype=np.float64))
    s = torch.as_tensor(np.asarray(scaled_radius, dtype=np.float64))
    n = pos.shape[0]
    d = pos[:, None, :] - pos[None, :, :]
    eye = torch.eye(n, dtype=torch.bool)
    r2 = (d * d).sum(-1)
    r = torch.sqrt(torch.where(eye, torch.ones_like(r2), r2))        # diagonal masked out below
    sj = s[None, :].expand(n, n)
    oi = o[:, None].expand(n, n)
    U = r + sj
    L = torch.maximum(oi, torch.abs(r - sj))
    H = 0.5 * (1.0 / L - 1.0 / U + 0.25 * (r - sj * sj / r) * (1.0 / (U * U) - 1.0 / (L * L)) + 0.5 * torch.log(L / U) / r)
    on = ((r + sj - oi) >= 0.0) & ~eye
    I = torch.where(on, H, torch.zeros_like(H)).sum(1)
    B = 1.0 / (1.0 / o - I)
    pref = GB_COULOMB * (1.0 / solute_dielectric - 1.0 / solvent_dielectric)
    e = (-0.5 * pref * q * q / B).sum()
    if sa_ace:
        radius = o + HCT_OFFSET
        e = e + (ACE_COEFF * (radius + ACE_PROBE) ** 2 * (radius / B) ** 6).sum()
    BB = B[:, None] * B[None, :]
    f = torch.sqrt(r2 + BB * torch.exp(-r2 / (4.0 * BB)))
    pair = torch.where(eye, torch.zeros_like(f), -pref * q[:, None] * q[None, :] / f)
    return e + 0.5 * pair.sum(), B


def hct(pos, charge, offset_radius, scaled_radius, solute_dielectric=1.0, solvent_dielectric=78.5, sa_ace=True):
    """(energy, forces [n,3], Born radii [n]) as numpy float64."""
    x = torch.tensor(np.asarray(pos, dtype=np.float64), requires_grad=True)
    e, B = hct_energy(x, charge, offset_radius, scaled_radius, solute_dielectric, solvent_dielectric, sa_ace)
    (g,) = torch.autograd.grad(e, x)
    return float(e.detach()), -g.numpy(), B.detach().numpy()
