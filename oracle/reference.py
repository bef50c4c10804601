"""ctypes front-end of oracle/_ref/libsdmref.so: the REFERENCE's own plugin sources compiled in
place from /root/reference against OpenMM stand-in headers (oracle/Makefile target `ref`,
oracle/ref_driver.cpp).

TEST INFRASTRUCTURE ONLY -- importable from tests/ and tools/make_ref_golden.py; never from
openmm_sdm_plugin_b200/.  The library exists where /root/reference was present at build time (the
build container); it travels to the GPU box with the snapshot and needs nothing from
/root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsdmref.so")
REFERENCE_ROOT = "/root/reference"
_LIB = None

FORCE_CB = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))


class RefParams(C.Structure):
    _fields_ = [("temperature", C.c_double), ("friction", C.c_double), ("step_size", C.c_double),
                ("bias_method", C.c_int), ("softcore_method", C.c_int),
                ("lambdac", C.c_double), ("gammac", C.c_double), ("wbcoeff", C.c_double),
                ("w0coeff", C.c_double), ("lambda1", C.c_double), ("lambda2", C.c_double),
                ("alpha", C.c_double), ("u0", C.c_double),
                ("umax", C.c_double), ("acore", C.c_double), ("ubcore", C.c_double),
                ("nonequilibrium", C.c_int), ("pad_", C.c_int),
                ("noneq_tmax", C.c_double), ("work_value", C.c_double), ("time", C.c_double),
                ("m_lambda1", C.c_double), ("m_lambda2", C.c_double), ("m_u0", C.c_double),
                ("m_w0", C.c_double), ("b_lambda1", C.c_double), ("b_lambda2", C.c_double),
                ("b_u0", C.c_double), ("b_w0", C.c_double)]


class RefOut(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("bind_e", "pot_energy", "work_value", "lambdac", "lambda1",
                                          "lambda2", "u0", "w0coeff", "time", "kinetic_energy")] + \
               [("step_count", C.c_int), ("pad_", C.c_int)]


def build(force: bool = False):
    """Compile the reference where /root/reference exists; returns the library path or None."""
    if not os.path.isdir(REFERENCE_ROOT):
        return _SO if os.path.exists(_SO) else None
    if force or not os.path.exists(_SO) or \
            os.path.getmtime(os.path.join(_HERE, "ref_driver.cpp")) > os.path.getmtime(_SO):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "ref"])
    return _SO


def available() -> bool:
    return os.path.exists(_SO) or build() is not None


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libsdmref.so is not built and /root/reference is absent")
        L = C.CDLL(so)
        L.sdmref_last_error.restype = C.c_char_p
        L.sdmref_softcore.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.sdmref_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                 C.POINTER(RefParams), FORCE_CB, C.c_void_p, C.c_int, C.POINTER(RefOut),
                                 C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def softcore(method, u, umax, a, ub):
    """LangevinIntegratorSDM::SoftCoreF of the reference; raises ValueError where it throws."""
    usc, fp = C.c_double(), C.c_double()
    if lib().sdmref_softcore(int(method), float(u), float(umax), float(a), float(ub), C.byref(usc), C.byref(fp)):
        raise ValueError(lib().sdmref_last_error().decode())
    return usc.value, fp.value


def defaults() -> dict:
    p = RefParams()
    lib().sdmref_defaults(C.byref(p))
    return {k: getattr(p, k) for k, _ in RefParams._fields_ if k != "pad_"}


def params_from_alch(alch, temperature=300.0, friction=0.5) -> RefParams:
    """RefParams from an openmm_sdm_plugin_b200.system.AlchemicalState."""
    p = RefParams()
    p.temperature, p.friction, p.step_size = temperature, friction, alch.step_size
    for k in ("bias_method", "softcore_method", "lambdac", "gammac", "wbcoeff", "w0coeff", "lambda1", "lambda2",
              "alpha", "u0", "umax", "acore", "ubcore", "nonequilibrium", "noneq_tmax", "work_value", "time",
              "m_lambda1", "m_lambda2", "m_u0", "m_w0", "b_lambda1", "b_lambda2", "b_u0", "b_w0"):
        setattr(p, k, getattr(alch, k))
    return p


def run(masses, positions, velocities, displacement, params: RefParams, force_fn, steps=1,
        force_groups=(1, 2), noise=None):
    """`steps` x LangevinIntegratorSDM::step(1).  force_fn(groups, positions[n,3]) -> (energy,
    forces[n,3]) plays OpenMM's calcForcesAndEnergy (groups 4 = nonbonded, 2 = bonded).
    Returns dict(positions, velocities, hybrid_force, traj[steps,2] and the RefOut fields)."""
    L = lib()
    n = len(masses)
    m = np.ascontiguousarray(masses, np.float64)
    x = np.array(positions, np.float64, order="C").reshape(n, 3)
    v = np.array(velocities, np.float64, order="C").reshape(n, 3)
    d = np.ascontiguousarray(displacement, np.float64).reshape(n, 3)
    fg = np.ascontiguousarray(force_groups, np.int32)
    hyb = np.zeros((n, 3))
    traj = np.zeros((steps, 2))
    nz = np.ascontiguousarray(noise if noise is not None else [], np.float64).ravel()
    L.sdmref_set_noise(nz.ctypes.data_as(C.c_void_p), len(nz))
    err = []

    def cb(user, groups, nn, pos_p, f_p):
        try:
            pos = np.ctypeslib.as_array(pos_p, shape=(nn, 3))
            e, f = force_fn(int(groups), pos.copy())
            np.ctypeslib.as_array(f_p, shape=(nn, 3))[...] = f
            return float(e)
        except Exception as ex:   # never let an exception cross the C frame
            err.append(ex)
            return 0.0

    out = RefOut()
    rc = L.sdmref_run(n, m.ctypes.data, x.ctypes.data, v.ctypes.data, d.ctypes.data, fg.ctypes.data, len(fg),
                      C.byref(params), FORCE_CB(cb), None, steps, C.byref(out), hyb.ctypes.data, traj.ctypes.data)
    if err:
        raise err[0]
    if rc:
        raise RuntimeError(L.sdmref_last_error().decode())
    res = {k: getattr(out, k) for k, _ in RefOut._fields_ if k != "pad_"}
    res.update(positions=x, velocities=v, hybrid_force=hyb, traj=traj)
    return res


# ---- the reference's integrator with the PRODUCT's OpenMM-side kernel (oracle/adapter_driver.cpp) -----
_SO_B200 = os.path.join(_HERE, "_ref", "libsdmb200_openmm.so")
_LIB_B200 = None


class B200Nonbonded(C.Structure):
    _fields_ = [("method", C.c_int), ("n_exceptions", C.c_int), ("use_dispersion_correction", C.c_int), ("pad_", C.c_int),
                ("cutoff", C.c_double), ("eps_rf", C.c_double), ("box", C.c_double * 3),
                ("charge", C.c_void_p), ("sigma", C.c_void_p), ("epsilon", C.c_void_p),
                ("exception_pairs", C.c_void_p), ("exception_params", C.c_void_p),
                ("lj_geometric", C.c_int), ("gb_ace", C.c_int),
                ("gb_charge", C.c_void_p), ("gb_or", C.c_void_p), ("gb_sr", C.c_void_p),
                ("gb_solute", C.c_double), ("gb_solvent", C.c_double)]


def b200_adapter_available() -> bool:
    if os.path.isdir(REFERENCE_ROOT):
        build()
    return os.path.exists(_SO_B200)


def _lib_b200():
    global _LIB_B200
    if _LIB_B200 is None:
        from openmm_sdm_plugin_b200 import _lib as product
        product.lib()                       # the product library is built and resolvable before the adapter loads it
        L = C.CDLL(_SO_B200)
        L.sdmb200_adapter_last_error.restype = C.c_char_p
        L.sdmb200_adapter_run.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(B200Nonbonded), C.c_int, C.c_void_p, C.c_void_p,
                                          C.POINTER(RefParams), FORCE_CB, C.c_void_p, C.c_int, C.POINTER(RefOut),
                                          C.c_void_p, C.c_void_p]
        _LIB_B200 = L
    return _LIB_B200


def run_b200(level, system, masses, positions, velocities, displacement, params: RefParams, force_fn, steps=1,
             noise=None, constraint_pairs=None, constraint_dist=None):
    """The same `steps` x LangevinIntegratorSDM::step(1) as run(), with the B200 kernel registered for
    IntegrateLangevinStepSDMKernel.  level 0 = fused (the System carries a B200NonbondedForce built from
    `system`, force_fn is only asked for group mask 2 and gets zero work for mask 4), level 1 = literal
    operations (force_fn evaluates mask 4 like OpenMM would)."""
    L = _lib_b200()
    n = len(masses)
    m = np.ascontiguousarray(masses, np.float64)
    x = np.array(positions, np.float64, order="C").reshape(n, 3)
    v = np.array(velocities, np.float64, order="C").reshape(n, 3)
    d = np.ascontiguousarray(displacement, np.float64).reshape(n, 3)
    hyb = np.zeros((n, 3))
    traj = np.zeros((steps, 2))
    nz = np.ascontiguousarray(noise if noise is not None else [], np.float64).ravel()
    L.sdmb200_adapter_set_noise(nz.ctypes.data_as(C.c_void_p), len(nz))
    # every addException of the reader: zero-parameter exclusions and the parameterised 1-4 pairs
    q = np.ascontiguousarray(system.charge, np.float64)
    sg = np.ascontiguousarray(system.sigma, np.float64)
    ep = np.ascontiguousarray(system.epsilon, np.float64)
    excl = np.ascontiguousarray(system.exclusions, np.int32).reshape(-1, 2)
    par = {(min(a, b), max(a, b)): p for (a, b), p in zip(np.asarray(system.exception_pairs).reshape(-1, 2).tolist(),
                                                           np.asarray(system.exception_params).reshape(-1, 3).tolist())}
    ex_pairs = np.ascontiguousarray(excl, np.int32)
    ex_par = np.ascontiguousarray([par.get((min(a, b), max(a, b)), (0.0, 1.0, 0.0)) for a, b in excl.tolist()],
                                  np.float64).reshape(-1, 3)
    nb = B200Nonbonded()
    nb.method, nb.n_exceptions = int(system.method), len(ex_pairs)
    nb.use_dispersion_correction = int(bool(system.use_dispersion_correction))
    nb.cutoff, nb.eps_rf = float(system.cutoff), float(system.eps_rf)
    for k in range(3):
        nb.box[k] = float(system.box[k])
    nb.charge, nb.sigma, nb.epsilon = q.ctypes.data, sg.ctypes.data, ep.ctypes.data
    nb.exception_pairs, nb.exception_params = ex_pairs.ctypes.data, ex_par.ctypes.data
    nb.lj_geometric = int(bool(getattr(system, "lj_geometric", False)))
    gb = getattr(system, "gb", None)
    if gb is not None:                      # GBSAHCTForce of the nonbonded group (system.py)
        gq, go, gs = [np.ascontiguousarray(a, np.float64) for a in gb.device_parameters()]
        nb.gb_charge, nb.gb_or, nb.gb_sr = gq.ctypes.data, go.ctypes.data, gs.ctypes.data
        nb.gb_solute, nb.gb_solvent, nb.gb_ace = gb.soluteDielectric, gb.solventDielectric, int(gb.SA == "ACE")
    cp = np.ascontiguousarray(constraint_pairs if constraint_pairs is not None else np.zeros((0, 2)), np.int32)
    cd = np.ascontiguousarray(constraint_dist if constraint_dist is not None else np.zeros(0), np.float64)
    err = []

    def cb(user, groups, nn, pos_p, f_p):
        try:
            pos = np.ctypeslib.as_array(pos_p, shape=(nn, 3))
            e, f = force_fn(int(groups), pos.copy())
            np.ctypeslib.as_array(f_p, shape=(nn, 3))[...] = f
            return float(e)
        except Exception as ex:   # never let an exception cross the C frame
            err.append(ex)
            return 0.0

    out = RefOut()
    rc = L.sdmb200_adapter_run(int(level), n, m.ctypes.data, x.ctypes.data, v.ctypes.data, d.ctypes.data, C.byref(nb),
                               len(cd), cp.ctypes.data, cd.ctypes.data, C.byref(params), FORCE_CB(cb), None, steps,
                               C.byref(out), hyb.ctypes.data, traj.ctypes.data)
    if err:
        raise err[0]
    if rc:
        raise RuntimeError(L.sdmb200_adapter_last_error().decode())
    res = {k: getattr(out, k) for k, _ in RefOut._fields_ if k != "pad_"}
    res.update(positions=x, velocities=v, hybrid_force=hyb, traj=traj)
    return res
