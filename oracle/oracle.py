"""ctypes front-end of the CPU oracle (oracle/sdm_oracle.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, bench.py's cpu_baseline / --impl
reference legs and __graft_entry__.smoke(); never from openmm_sdm_plugin_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import asdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcSystem(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("method", C.c_int32), ("cutoff", C.c_double),
                ("eps_rf", C.c_double), ("box", C.c_double * 3),
                ("use_dispersion_correction", C.c_int32), ("n_exclusions", C.c_int32),
                ("n_exceptions", C.c_int32), ("pad_", C.c_int32),
                ("charge", C.c_void_p), ("sigma", C.c_void_p), ("epsilon", C.c_void_p),
                ("exclusions", C.c_void_p), ("exceptions", C.c_void_p),
                ("exception_params", C.c_void_p), ("ewald_alpha", C.c_double),
                ("lj_geometric", C.c_int32), ("pad2_", C.c_int32)]


class OrcAlch(C.Structure):
    _fields_ = [("bias_method", C.c_int32), ("softcore_method", C.c_int32),
                ("lambdac", C.c_double), ("gammac", C.c_double), ("wbcoeff", C.c_double),
                ("w0coeff", C.c_double), ("lambda1", C.c_double), ("lambda2", C.c_double),
                ("alpha", C.c_double), ("u0", C.c_double), ("umax", C.c_double),
                ("acore", C.c_double), ("ubcore", C.c_double),
                ("nonequilibrium", C.c_int32), ("pad_", C.c_int32),
                ("noneq_tmax", C.c_double), ("work_value", C.c_double), ("time", C.c_double),
                ("step_size", C.c_double),
                ("m_lambda1", C.c_double), ("m_lambda2", C.c_double), ("m_u0", C.c_double),
                ("m_w0", C.c_double), ("b_lambda1", C.c_double), ("b_lambda2", C.c_double),
                ("b_u0", C.c_double), ("b_w0", C.c_double)]


class OrcResult(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("E1", "E2", "Eb", "u", "u_sc", "fp", "ebias", "bfp", "sp", "pot_energy",
                 "bind_e", "E1_pair", "E1_exc", "E1_disp")] + \
               [("n_pairs1", C.c_int64), ("n_pairs2", C.c_int64)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "sdm_oracle.c")
    if force or not os.path.exists(so) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liborc.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_softcore.restype = C.c_double
        L.orc_softcore.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_bias.restype = None
        L.orc_bias.argtypes = [C.POINTER(OrcAlch), C.c_double, C.POINTER(C.c_double),
                               C.POINTER(C.c_double)]
        L.orc_nonbonded.restype = C.c_int
        L.orc_nonbonded.argtypes = [C.POINTER(OrcSystem), C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_int64), C.c_int]
        L.orc_sdm_eval.restype = C.c_int
        L.orc_sdm_eval.argtypes = [C.POINTER(OrcSystem), C.POINTER(OrcAlch), C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.POINTER(OrcResult), C.c_int]
        L.orc_dispersion_coefficient.restype = C.c_double
        L.orc_dispersion_coefficient.argtypes = [C.POINTER(OrcSystem)]
        L.orc_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _Sys:
    """Keeps the numpy buffers alive next to the C struct."""

    def __init__(self, s):
        self.keep = [np.ascontiguousarray(s.charge, np.float64),
                     np.ascontiguousarray(s.sigma, np.float64),
                     np.ascontiguousarray(s.epsilon, np.float64),
                     np.ascontiguousarray(s.exclusions, np.int32),
                     np.ascontiguousarray(s.exception_pairs, np.int32),
                     np.ascontiguousarray(s.exception_params, np.float64)]
        c = OrcSystem()
        c.n_atoms = s.n_atoms
        # NonbondedForce::Ewald (3) / ::PME (4): a periodic cutoff system whose Coulomb term is the direct-space
        # Ewald one (sdm_oracle.c: ewald_alpha > 0)
        ewald = int(s.method) in (3, 4)
        c.method = 2 if ewald else int(s.method)
        c.ewald_alpha = float(s.ewald_alpha_effective()) if ewald else 0.0
        c.lj_geometric = int(bool(getattr(s, "lj_geometric", False)))
        c.cutoff = float(s.cutoff)
        c.eps_rf = float(s.eps_rf)
        for d in range(3):
            c.box[d] = float(s.box[d])
        c.use_dispersion_correction = int(bool(s.use_dispersion_correction))
        c.n_exclusions = len(self.keep[3])
        c.n_exceptions = len(self.keep[4])
        (c.charge, c.sigma, c.epsilon, c.exclusions, c.exceptions,
         c.exception_params) = [_ptr(a) for a in self.keep]
        self.c = c


def _alch(a) -> OrcAlch:
    c = OrcAlch()
    for k, v in asdict(a).items():
        setattr(c, k, v)
    return c


def _alch_back(c: OrcAlch, a):
    for k in asdict(a):
        setattr(a, k, getattr(c, k))


def softcore(method, u, umax, a, ub):
    fp = C.c_double()
    err = C.c_int()
    usc = lib().orc_softcore(int(method), float(u), float(umax), float(a), float(ub),
                             C.byref(fp), C.byref(err))
    if err.value:
        raise ValueError("Unknown soft core method")
    return usc, fp.value


def bias(alch, bind_e):
    c = _alch(alch)
    e, b = C.c_double(), C.c_double()
    lib().orc_bias(C.byref(c), float(bind_e), C.byref(e), C.byref(b))
    _alch_back(c, alch)
    return e.value, b.value


def dispersion_coefficient(system) -> float:
    s = _Sys(system)
    return float(lib().orc_dispersion_coefficient(C.byref(s.c)))


def nonbonded(system, positions, want_pairs=False, nthreads=1):
    """One Reference-platform NonbondedForce evaluation.  Returns dict(E_pair, E_exc,
    E_disp, E, forces[, pairs])."""
    s = _Sys(system)
    pos = np.ascontiguousarray(positions, np.float64)
    n = system.n_atoms
    f = np.zeros((n, 3))
    ep, ee, ed = C.c_double(), C.c_double(), C.c_double()
    npairs = C.c_int64()
    pairs = None
    if want_pairs:
        rc = lib().orc_nonbonded(C.byref(s.c), _ptr(pos), _ptr(f), C.byref(ep), C.byref(ee),
                                 C.byref(ed), None, 0, C.byref(npairs), nthreads)
        if rc < 0:
            raise RuntimeError("oracle error %d" % rc)
        pairs = np.zeros((npairs.value, 2), np.int32)
    rc = lib().orc_nonbonded(C.byref(s.c), _ptr(pos), _ptr(f), C.byref(ep), C.byref(ee),
                             C.byref(ed), _ptr(pairs), len(pairs) if pairs is not None else 0,
                             C.byref(npairs), nthreads)
    if rc < 0:
        raise RuntimeError("oracle error %d (box smaller than 2*cutoff?)" % rc)
    out = dict(E_pair=ep.value, E_exc=ee.value, E_disp=ed.value,
               E=ep.value + ee.value + ed.value, forces=f, n_pairs=npairs.value)
    if want_pairs:
        out["pairs"] = pairs
    return out


def sdm_eval(system, alch, displacement, positions, fb=None, eb=0.0, nthreads=1,
             want_state_forces=True):
    """LangevinIntegratorSDM::step force path (two full evaluations + execute() up to the
    hybrid force).  `alch` is updated in place like the integrator object is."""
    s = _Sys(system)
    pos = np.array(positions, np.float64, order="C")
    disp = np.ascontiguousarray(displacement, np.float64)
    n = system.n_atoms
    fbc = np.ascontiguousarray(fb, np.float64) if fb is not None else None
    f = np.zeros((n, 3))
    f1 = np.zeros((n, 3)) if want_state_forces else None
    f2 = np.zeros((n, 3)) if want_state_forces else None
    c = _alch(alch)
    res = OrcResult()
    rc = lib().orc_sdm_eval(C.byref(s.c), C.byref(c), _ptr(disp), _ptr(pos), _ptr(fbc),
                            float(eb), _ptr(f), _ptr(f1), _ptr(f2), C.byref(res), nthreads)
    if rc == -3:
        raise ValueError("Unknown soft core method")
    if rc < 0:
        raise RuntimeError("oracle error %d" % rc)
    _alch_back(c, alch)
    out = {k: getattr(res, k) for k, _ in OrcResult._fields_}
    out.update(forces=f, f1=f1, f2=f2)
    return out
