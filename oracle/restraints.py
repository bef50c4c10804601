"""TEST INFRASTRUCTURE (oracle): the restraint expressions of the reference's SDMUtils restated with torch
(float64) so that autograd gives the forces -- the checker of csrc/kernels_restraints.cu.

What is restated: python/SDMUtils.py:61-85 (addRestraintForce: flat-bottom centroid distance, optional
flat-bottom angle and two dihedrals between reference atoms, all scaled by SDMRestraintControlParameter) and
:183-256 (addAlignmentForce: displacement, theta and the symmetrised psi term).  The functions the expression
strings call -- step(x) = 0 for x < 0 else 1, max, floor, angle(p1,p2,p3) at the middle point,
dihedral(p1,p2,p3,p4) = angle between (p1-p2)x(p3-p2) and (p3-p2)x(p3-p4) with the sign of
(p1-p2).((p3-p2)x(p3-p4)), centroid = weighted mean (OpenMM's default weights are the masses) -- follow
OpenMM 7.3's ReferenceCustomCentroidBondIxn / ReferenceCustomCompoundBondIxn.  OpenMM is not available here:
parity unpinned for these conventions (the derivatives are what this oracle checks).
Only tests/ may import this module."""
import math

import numpy as np
import torch


def _centroid(pos, atoms, weights=None):
    atoms = torch.as_tensor(np.asarray(atoms, dtype=np.int64))
    w = torch.ones(len(atoms), dtype=torch.float64) if weights is None else torch.as_tensor(np.asarray(weights, dtype=np.float64))
    w = w / w.sum()
    return (pos[atoms] * w[:, None]).sum(0)


def _angle(p1, p2, p3):
    a, b = p1 - p2, p3 - p2
    return torch.acos(torch.clamp(torch.dot(a, b) / torch.sqrt(torch.dot(a, a) * torch.dot(b, b)), -1.0, 1.0))


def _dihedral(p1, p2, p3, p4):
    v0, v1, v2 = p1 - p2, p3 - p2, p3 - p4
    c0, c1 = torch.linalg.cross(v0, v1), torch.linalg.cross(v1, v2)
    ang = torch.acos(torch.clamp(torch.dot(c0, c1) / torch.sqrt(torch.dot(c0, c0) * torch.dot(c1, c1)), -1.0, 1.0))
    return -ang if float(torch.dot(v0, c1).detach()) < 0.0 else ang


def _wrap(x, period):
    return x - period * math.floor(float(x.detach()) / period + 0.5)


def _flat_bottom(value, kf, a, b, period):
    db, da, dm = _wrap(value - b, period), _wrap(value - a, period), _wrap(value - 0.5 * (a + b), period)
    e = torch.zeros((), dtype=torch.float64)
    if float(dm.detach()) >= 0.0:
        e = e + torch.clamp(db, min=0.0) ** 2
    if -float(dm.detach()) >= 0.0:
        e = e + torch.clamp(-da, min=0.0) ** 2
    return 0.5 * kf * e


def centroid_restraint_energy(pos, spec, control=1.0):
    """spec: dict with lig_cm_atoms, rcpt_cm_atoms, [lig_cm_weights, rcpt_cm_weights], kfcm, tolcm, offset and,
    optionally, lig_ref, rcpt_ref, kfcd[3], a[3], b[3] (SDMUtils.py:61-85)."""
    g1 = _centroid(pos, spec["lig_cm_atoms"], spec.get("lig_cm_weights"))
    g2 = _centroid(pos, spec["rcpt_cm_atoms"], spec.get("rcpt_cm_weights"))
    off = torch.as_tensor(np.asarray(spec.get("offset", (0, 0, 0)), dtype=np.float64))
    d12 = torch.sqrt(((g1 - off - g2) ** 2).sum())
    e = torch.zeros((), dtype=torch.float64)
    if float(d12.detach()) - spec["tolcm"] >= 0.0:
        e = e + 0.5 * spec["kfcm"] * (d12 - spec["tolcm"]) ** 2
    if spec.get("lig_ref") is not None and spec.get("rcpt_ref") is not None:
        g3, g4, g5 = (pos[i] for i in spec["rcpt_ref"])
        g6, g7, g8 = (pos[i] for i in spec["lig_ref"])
        kf, a, b = spec["kfcd"], spec["a"], spec["b"]
        e = e + _flat_bottom(_angle(g3, g6, g7), kf[0], a[0], b[0], math.pi)
        e = e + _flat_bottom(_dihedral(g4, g3, g6, g7), kf[1], a[1], b[1], 2 * math.pi)
        e = e + _flat_bottom(_dihedral(g3, g6, g7, g8), kf[2], a[2], b[2], 2 * math.pi)
    return control * e


def _psi(x1, x2, x3, x4, x5, k):
    d1 = x2 - x1
    dn1 = d1 / torch.sqrt(torch.dot(d1, d1))
    d0, d3 = x3 - x1, x5 - x4
    v, w = d0 - torch.dot(d0, dn1) * dn1, d3 - torch.dot(d3, dn1) * dn1
    return 0.5 * k * (1.0 - torch.dot(v, w) / torch.sqrt(torch.dot(v, v) * torch.dot(w, w)))


def alignment_energy(pos, spec):
    """spec: dict with liga_ref[3], ligb_ref[3], kfdispl, ktheta, kpsi, offset (SDMUtils.py:183-256)."""
    a1, a2, a3 = (pos[i] for i in spec["liga_ref"])
    b1, b2, b3 = (pos[i] for i in spec["ligb_ref"])
    off = torch.as_tensor(np.asarray(spec.get("offset", (0, 0, 0)), dtype=np.float64))
    e = 0.5 * spec["kfdispl"] * ((b1 - off - a1) ** 2).sum()
    d1, d2 = b2 - b1, a2 - a1
    e = e + 0.5 * spec["ktheta"] * (1.0 - torch.dot(d1, d2) / torch.sqrt(torch.dot(d1, d1) * torch.dot(d2, d2)))
    e = e + _psi(b1, b2, b3, a1, a3, 0.5 * spec["kpsi"]) + _psi(a1, a2, a3, b1, b3, 0.5 * spec["kpsi"])
    return e


def energy_and_forces(positions, centroid_specs=(), alignment_specs=(), control=1.0):
    """Total restraint energy (kJ/mol) and forces (kJ/mol/nm, [n,3]) of one frame."""
    pos = torch.tensor(np.asarray(positions, dtype=np.float64), requires_grad=True)
    e = torch.zeros((), dtype=torch.float64)
    for s in centroid_specs:
        e = e + centroid_restraint_energy(pos, s, control)
    for s in alignment_specs:
        e = e + alignment_energy(pos, s)
    if e.requires_grad:
        e.backward()
        f = -pos.grad.numpy()
    else:
        f = np.zeros_like(np.asarray(positions, dtype=np.float64))
    return float(e.detach()), f
