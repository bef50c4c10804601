"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

CPU restatement of the constrained half of the reference's Langevin step
(/root/reference/platforms/reference/src/ReferenceStochasticDynamicsSDM.cpp:216-266):

    updatePart1  v  = vscale*v + fscale*F/m + noisescale*xi/sqrt(m)      (:131-169)
    updatePart2  x' = x + dt*v                                           (:183-201)
    constraints  x' <- apply(x, x', 1/m, tolerance)                      (:250-252)
    finish       v  = (x' - x)/dt ;  x = x'                              (:256-262)

The constraint algorithm itself is OpenMM's (ReferenceSETTLEAlgorithm / ReferenceCCMAAlgorithm, not
vendored in /root/reference, OpenMM 7.3 as pinned by the reference's CMake find module).  Both, like
SHAKE, solve the same equations: the constrained positions are x' + sum_k g_k * (1/m_i) * r_k(x),
i.e. displaced along the bond vectors r_k of the OLD positions, with the multipliers g_k chosen so
that every constrained distance has its target length.  That solution is unique for small steps;
this file iterates SHAKE to 1e-13 relative (far below the integrator's 1e-5 tolerance), so a device
SETTLE must agree with it to rounding and a device SHAKE/CCMA to within its own tolerance.

parity unpinned: no golden vector of the reference exercises constraints (its shipped examples run
with them, but the repository holds no numbers); the properties the tests check are the defining
equations above.
"""
import numpy as np


def langevin_constants(temperature, friction, dt):
    """ReferenceStochasticDynamicsSDM.cpp:144-148"""
    tau = 1.0 / friction
    boltz = 1.380658e-23 * 6.0221367e23 / 1000.0
    kT = boltz * temperature
    vscale = np.exp(-dt / tau)
    fscale = (1 - vscale) * tau
    noisescale = np.sqrt(2 * kT / tau) * np.sqrt(0.5 * (1 - vscale * vscale) * tau)
    return vscale, fscale, noisescale


def shake(x, xp, inv_mass, pairs, dist, rtol=1e-13, max_iter=100000):
    """Constrained positions reached from xp along the bond vectors of x (all clusters at once,
    Gauss-Seidel sweeps in constraint order, vectorised over independent constraints is not needed
    at fixture sizes: the loop runs over sweeps, each sweep handles every constraint with numpy
    scatter-adds in a Jacobi fashion damped for shared atoms)."""
    x = np.asarray(x, np.float64)
    p = np.array(xp, np.float64, copy=True)
    i, j = pairs[:, 0], pairs[:, 1]
    r0 = x[i] - x[j]
    d2 = dist * dist
    wi, wj = inv_mass[i], inv_mass[j]
    # number of constraints touching each atom: Jacobi updates are damped by it so that coupled
    # clusters (water: three constraints on three atoms) converge
    deg = np.zeros(len(x))
    np.add.at(deg, i, 1.0)
    np.add.at(deg, j, 1.0)
    damp = 1.0 / np.maximum(deg[i], deg[j])
    for _ in range(max_iter):
        rp = p[i] - p[j]
        rp2 = (rp * rp).sum(1)
        err = np.abs(rp2 - d2) / d2
        if err.max() < 2 * rtol:
            return p
        g = damp * (d2 - rp2) / (2.0 * (r0 * rp).sum(1) * (wi + wj))
        np.add.at(p, i, r0 * (g * wi)[:, None])
        np.add.at(p, j, -r0 * (g * wj)[:, None])
    raise RuntimeError("oracle SHAKE did not converge")


def langevin_step(x, v, force, masses, temperature, friction, dt, xi, pairs=None, dist=None):
    """One constrained step of the reference's integrator; returns (x_new, v_new, x_unconstrained)."""
    vscale, fscale, noisescale = langevin_constants(temperature, friction, dt)
    inv_m = np.where(masses > 0, 1.0 / np.where(masses > 0, masses, 1.0), 0.0)
    moving = inv_m > 0
    vn = np.array(v, np.float64, copy=True)
    vn[moving] = (vscale * v[moving] + (fscale * inv_m[moving])[:, None] * force[moving]
                  + (noisescale * np.sqrt(inv_m[moving]))[:, None] * xi[moving])
    xp = np.array(x, np.float64, copy=True)
    xp[moving] = x[moving] + dt * vn[moving]
    xc = xp
    if pairs is not None and len(pairs):
        xc = shake(x, xp, inv_m, np.asarray(pairs), np.asarray(dist))
    v_new = np.array(v, np.float64, copy=True)
    v_new[moving] = (xc[moving] - x[moving]) / dt
    x_new = np.array(x, np.float64, copy=True)
    x_new[moving] = xc[moving]
    return x_new, v_new, xp
