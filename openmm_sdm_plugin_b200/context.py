"""SDMContext -- thin Python owner of an `sdm_ctx` handle (include/sdmb200.h).

All arithmetic happens in libsdmb200.so on the GPU; this class only marshals numpy buffers
through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import asdict

import numpy as np

from . import _lib
from .system import AlchemicalState, NonbondedSystem


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def alch_to_c(a: AlchemicalState) -> _lib.SdmAlch:
    c = _lib.SdmAlch()
    for k, v in asdict(a).items():
        setattr(c, k, v)
    return c


def alch_from_c(c: _lib.SdmAlch, a: AlchemicalState) -> AlchemicalState:
    for k in asdict(a):
        setattr(a, k, getattr(c, k))
    return a


class PinnedArray:
    """A numpy view of cudaMallocHost memory (for the end-to-end path)."""

    def __init__(self, shape, dtype=np.float64):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _lib.check(_lib.lib().sdm_host_alloc(C.byref(p), self.nbytes))
        self._p = p
        buf = (C.c_char * self.nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self._p is not None:
            self.array = None
            _lib.lib().sdm_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SDMContext:
    def __init__(self, system: NonbondedSystem, displacement=None, n_replicas: int = 1,
                 pair_mode: int = _lib.PAIR_AUTO, device: int = -1, skin: float = -1.0,
                 nstlist: int = 0, exact_cutoff: bool = True, use_graph: bool = True):
        L = _lib.lib()
        self._L = L
        self.system = system
        self.n = system.n_atoms
        self.R = int(n_replicas)
        keep = [np.ascontiguousarray(system.charge, np.float64),
                np.ascontiguousarray(system.sigma, np.float64),
                np.ascontiguousarray(system.epsilon, np.float64),
                np.ascontiguousarray(system.exclusions, np.int32),
                np.ascontiguousarray(system.exception_pairs, np.int32),
                np.ascontiguousarray(system.exception_params, np.float64),
                None if displacement is None else np.ascontiguousarray(displacement, np.float64)]
        if keep[6] is not None and keep[6].shape != (self.n, 3):
            raise ValueError("displacement map must be [n_atoms, 3]")
        s = _lib.SdmSystem()
        s.n_atoms = self.n
        s.method = int(system.method)
        s.cutoff = float(system.cutoff)
        s.eps_rf = float(system.eps_rf)
        for d in range(3):
            s.box[d] = float(system.box[d])
        s.use_dispersion_correction = int(bool(system.use_dispersion_correction))
        s.n_exclusions = len(keep[3])
        s.n_exceptions = len(keep[4])
        s.n_replicas = self.R
        s.ewald_alpha = float(getattr(system, "ewald_alpha", 0.0))
        s.ewald_tolerance = float(getattr(system, "ewald_tolerance", 0.0))
        s.lj_combining = 1 if getattr(system, "lj_geometric", False) else 0
        (s.charge, s.sigma, s.epsilon, s.exclusions, s.exceptions, s.exception_params,
         s.displacement) = [_ptr(a) for a in keep]
        o = _lib.SdmOptions()
        L.sdm_default_options(C.byref(o))
        o.device = device
        o.pair_mode = pair_mode
        o.skin = skin
        o.nstlist = nstlist
        o.exact_cutoff = int(exact_cutoff)
        o.use_graph = int(use_graph)
        h = C.c_void_p()
        _lib.check(L.sdm_create(C.byref(s), C.byref(o), C.byref(h)))
        self._h = h

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.sdm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- inputs ---------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        _lib.check(self._L.sdm_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_positions(self, replica: int, xyz: np.ndarray):
        a = np.ascontiguousarray(xyz, np.float64)
        if a.size != 3 * self.n:
            raise ValueError("positions must be [n_atoms, 3]")
        _lib.check(self._L.sdm_set_positions(self._h, replica, _ptr(a)))
        self._keep_pos = a  # async copy source must outlive the call

    def set_positions_all(self, xyz_all: np.ndarray):
        """[n_replicas, n_atoms, 3] host doubles -- or float32 for the single-precision transfer
        (half the PCIe bytes) -- ideally a PinnedArray view; one async copy."""
        if xyz_all.dtype not in (np.float64, np.float32) or not xyz_all.flags.c_contiguous \
                or xyz_all.size != 3 * self.n * self.R:
            raise ValueError("positions must be C-contiguous float64 / float32 [n_replicas, n_atoms, 3]")
        fn = self._L.sdm_set_positions_all if xyz_all.dtype == np.float64 else self._L.sdm_set_positions_all_f32
        _lib.check(fn(self._h, _ptr(xyz_all)))
        self._keep_pos = xyz_all

    def read_results(self, forces_out: np.ndarray | None = None, want_scalars: bool = True):
        """Hybrid forces of all replicas into forces_out ([R, n, 3] float64) and the list of
        scalar dicts; one synchronisation."""
        if forces_out is not None and (forces_out.dtype != np.float64 or not forces_out.flags.c_contiguous
                                       or forces_out.size != 3 * self.n * self.R):
            raise ValueError("forces_out must be C-contiguous float64 [n_replicas, n_atoms, 3]")
        sc = (_lib.SdmScalars * self.R)() if want_scalars else None
        _lib.check(self._L.sdm_read_results(self._h, _ptr(forces_out), C.cast(sc, C.c_void_p) if sc else None))
        if not want_scalars:
            return None
        return [{k: getattr(s, k) for k, _ in _lib.SdmScalars._fields_} for s in sc]

    def enqueue_results(self, forces_out: np.ndarray | None = None):
        """Queue the device->host copies of forces (+ scalar blocks) behind the last eval(); no
        synchronisation.  Pair with synchronize() and collect_scalars()."""
        if forces_out is not None and (forces_out.dtype not in (np.float64, np.float32) or not forces_out.flags.c_contiguous
                                       or forces_out.size != 3 * self.n * self.R):
            raise ValueError("forces_out must be C-contiguous float64 / float32 [n_replicas, n_atoms, 3]")
        f32 = forces_out is not None and forces_out.dtype == np.float32
        _lib.check((self._L.sdm_enqueue_results_f32 if f32 else self._L.sdm_enqueue_results)(self._h, _ptr(forces_out)))
        self._keep_f = forces_out

    def collect_scalars(self):
        sc = (_lib.SdmScalars * self.R)()
        _lib.check(self._L.sdm_collect_scalars(self._h, C.cast(sc, C.c_void_p)))
        return [{k: getattr(s, k) for k, _ in _lib.SdmScalars._fields_} for s in sc]

    def set_positions_ptr(self, replica: int, host_ptr: int):
        _lib.check(self._L.sdm_set_positions(self._h, replica, C.c_void_p(host_ptr)))

    def set_bonded_forces(self, replica: int, fb, eb: float = 0.0):
        a = None if fb is None else np.ascontiguousarray(fb, np.float64)
        _lib.check(self._L.sdm_set_bonded_forces(self._h, replica, _ptr(a), float(eb)))
        self._keep_fb = a

    def set_alchemical(self, replica: int, alch: AlchemicalState):
        c = alch_to_c(alch)
        _lib.check(self._L.sdm_set_alchemical(self._h, replica, C.byref(c)))

    def get_alchemical(self, replica: int, into: AlchemicalState | None = None) -> AlchemicalState:
        c = _lib.SdmAlch()
        _lib.check(self._L.sdm_get_alchemical(self._h, replica, C.byref(c)))
        return alch_from_c(c, into if into is not None else AlchemicalState())

    def set_displacement(self, displacement):
        a = np.ascontiguousarray(displacement, np.float64)
        if a.shape != (self.n, 3):
            raise ValueError("displacement map must be [n_atoms, 3]")
        _lib.check(self._L.sdm_set_displacement(self._h, _ptr(a)))

    # -- evaluation -----------------------------------------------------------------------
    def eval(self):
        _lib.check(self._L.sdm_eval(self._h))

    def synchronize(self):
        _lib.check(self._L.sdm_synchronize(self._h))

    def invalidate_list(self):
        _lib.check(self._L.sdm_invalidate_list(self._h))

    def scalars(self, replica: int = 0) -> dict:
        sc = _lib.SdmScalars()
        _lib.check(self._L.sdm_get_scalars(self._h, replica, C.byref(sc)))
        return {k: getattr(sc, k) for k, _ in _lib.SdmScalars._fields_}

    def forces(self, replica: int = 0, which: int = _lib.FORCE_HYBRID, out=None) -> np.ndarray:
        if out is None:
            out = np.empty((self.n, 3))
        _lib.check(self._L.sdm_get_forces(self._h, replica, which, _ptr(out)))
        return out

    def forces_into_ptr(self, replica: int, host_ptr: int, which: int = _lib.FORCE_HYBRID):
        _lib.check(self._L.sdm_get_forces(self._h, replica, which, C.c_void_p(host_ptr)))

    def pairs(self, replica: int = 0) -> np.ndarray:
        n = C.c_int64()
        _lib.check(self._L.sdm_get_pairs(self._h, replica, None, 0, C.byref(n)))
        out = np.zeros((n.value, 2), np.int32)
        if n.value:
            _lib.check(self._L.sdm_get_pairs(self._h, replica, _ptr(out), n.value, C.byref(n)))
        return out

    # -- introspection --------------------------------------------------------------------
    def launch_count(self) -> int:
        n = C.c_int64()
        _lib.check(self._L.sdm_get_launch_count(self._h, C.byref(n)))
        return n.value

    def set_timing(self, enabled):
        """True / 1: time the pair kernel as launched in production; 2: with every resident block it can have."""
        _lib.check(self._L.sdm_set_timing(self._h, int(enabled)))

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        _lib.check(self._L.sdm_get_last_timing(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- device-resident Langevin dynamics (sdm_md_*) ------------------------------
    def md_init(self, masses, temperature: float, friction: float, step_size: float, seed: int = 0):
        m = np.ascontiguousarray(masses, np.float64)
        if m.size != self.n:
            raise ValueError("masses must be [n_atoms]")
        _lib.check(self._L.sdm_md_init(self._h, _ptr(m), float(temperature), float(friction),
                                       float(step_size), int(seed)))

    def md_set_velocities(self, replica: int, v):
        a = np.ascontiguousarray(v, np.float64)
        if a.size != 3 * self.n:
            raise ValueError("velocities must be [n_atoms, 3]")
        _lib.check(self._L.sdm_md_set_velocities(self._h, replica, _ptr(a)))

    def md_velocities(self, replica: int) -> np.ndarray:
        out = np.empty((self.n, 3))
        _lib.check(self._L.sdm_md_get_velocities(self._h, replica, _ptr(out)))
        return out

    def positions(self, replica: int) -> np.ndarray:
        out = np.empty((self.n, 3))
        _lib.check(self._L.sdm_get_positions(self._h, replica, _ptr(out)))
        return out

    def md_set_constraints(self, pairs, distances, tolerance: float = 1e-5):
        """System.addConstraint pairs [n, 2] and distances [n] (nm); applied between the two halves
        of the update like ReferenceStochasticDynamicsSDM.cpp:250-252."""
        p = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        d = np.ascontiguousarray(distances, np.float64).reshape(-1)
        if len(p) != len(d):
            raise ValueError("pairs [n, 2] and distances [n] must have the same length")
        _lib.check(self._L.sdm_md_set_constraints(self._h, len(d), _ptr(p) if len(d) else None,
                                                  _ptr(d) if len(d) else None, float(tolerance)))

    def md_step(self, nsteps: int = 1):
        _lib.check(self._L.sdm_md_step(self._h, int(nsteps)))

    def md_counters(self):
        """(steps taken since md_init, enqueued steps that were repeated because of a stale list)"""
        a, b = C.c_uint64(), C.c_uint64()
        _lib.check(self._L.sdm_md_get_counters(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def md_update(self, forces_all=None):
        a = None if forces_all is None else np.ascontiguousarray(forces_all, np.float64)
        if a is not None and a.size != 3 * self.n * self.R:
            raise ValueError("forces must be [n_replicas, n_atoms, 3]")
        _lib.check(self._L.sdm_md_update(self._h, _ptr(a) if a is not None else None))
        if a is not None:
            self.synchronize()      # the upload source is a temporary

    def md_set_noise(self, xi_all):
        a = None if xi_all is None else np.ascontiguousarray(xi_all, np.float64)
        if a is not None and a.size != 3 * self.n * self.R:
            raise ValueError("noise must be [n_replicas, n_atoms, 3]")
        _lib.check(self._L.sdm_md_set_noise(self._h, _ptr(a) if a is not None else None))

    def md_kinetic_energy(self, replica: int) -> float:
        v = C.c_double()
        _lib.check(self._L.sdm_md_kinetic_energy(self._h, replica, C.byref(v)))
        return v.value

    def enable_reciprocal_pme(self, grid=None):
        """Reciprocal-space PME of both states on the device (SDM_PME / SDM_EWALD systems); grid = [3] mesh sizes or
        None for OpenMM's rule."""
        g = None if grid is None else np.ascontiguousarray(grid, dtype=np.int32)
        _lib.check(self._L.sdm_enable_reciprocal_pme(self._h, _ptr(g)))

    def enable_hct_gb(self, offset_radius, scaled_radius, charge=None, solute_dielectric=1.0,
                      solvent_dielectric=78.5, sa_ace=True):
        """HCT generalized Born (+ ACE surface area) of both states on the device: GBSAHCTForce(SA='ACE') of
        example/desmonddmsfile75.py:460.  offset_radius / scaled_radius: the CustomGBForce parameters "or" and "sr"
        (nm); charge None = the NonbondedForce charges."""
        o = np.ascontiguousarray(offset_radius, dtype=np.float64)
        sr = np.ascontiguousarray(scaled_radius, dtype=np.float64)
        q = None if charge is None else np.ascontiguousarray(charge, dtype=np.float64)
        if o.shape != (self.n,) or sr.shape != (self.n,) or (q is not None and q.shape != (self.n,)):
            raise ValueError("GB parameter arrays must be [n_atoms]")
        _lib.check(self._L.sdm_enable_hct_gb(self._h, _ptr(q), _ptr(o), _ptr(sr), float(solute_dielectric),
                                             float(solvent_dielectric), 1 if sa_ace else 0))

    def born_radii(self, replica: int, state: int = 1):
        """Born radii (nm) of state 1 (x) or 2 (x + d) in the last evaluation."""
        out = np.empty(self.n, np.float64)
        _lib.check(self._L.sdm_get_born_radii(self._h, replica, state, _ptr(out)))
        return out

    def set_external_dual(self, replica: int, f1_ext=None, f2_ext=None, e1_ext: float = 0.0, e2_ext: float = 0.0):
        """Energies and forces of both states computed outside the library (reciprocal-space PME, GB ...):
        E1 += e1, u += e2 - e1, F1 += f1, F2 - F1 += f2 - f1.  None removes them."""
        if f1_ext is None:
            _lib.check(self._L.sdm_set_external_dual(self._h, replica, None, None, 0.0, 0.0))
            return
        a = np.ascontiguousarray(f1_ext, dtype=np.float64)
        b = np.ascontiguousarray(f2_ext, dtype=np.float64)
        if a.shape != (self.n, 3) or b.shape != (self.n, 3):
            raise ValueError("external forces must be [n_atoms, 3]")
        _lib.check(self._L.sdm_set_external_dual(self._h, replica, _ptr(a), _ptr(b), float(e1_ext), float(e2_ext)))

    # ---- restraint forces of SDMUtils (python/SDMUtils.py:32-258); kJ/mol, nm, radians ------------------
    def add_centroid_restraint(self, lig_cm_atoms, rcpt_cm_atoms, kfcm, tolcm, offset=(0.0, 0.0, 0.0),
                               lig_cm_weights=None, rcpt_cm_weights=None, lig_ref=None, rcpt_ref=None,
                               kfcd=(0.0, 0.0, 0.0), a=(0.0, 0.0, 0.0), b=(0.0, 0.0, 0.0)):
        la = np.ascontiguousarray(lig_cm_atoms, dtype=np.int32)
        ra = np.ascontiguousarray(rcpt_cm_atoms, dtype=np.int32)
        lw = None if lig_cm_weights is None else np.ascontiguousarray(lig_cm_weights, dtype=np.float64)
        rw = None if rcpt_cm_weights is None else np.ascontiguousarray(rcpt_cm_weights, dtype=np.float64)
        r = _lib.SdmCentroidRestraint()
        r.n_lig_cm, r.n_rcpt_cm = len(la), len(ra)
        r.lig_cm_atoms = la.ctypes.data_as(C.POINTER(C.c_int32))
        r.rcpt_cm_atoms = ra.ctypes.data_as(C.POINTER(C.c_int32))
        r.lig_cm_weights = lw.ctypes.data_as(C.POINTER(C.c_double)) if lw is not None else None
        r.rcpt_cm_weights = rw.ctypes.data_as(C.POINTER(C.c_double)) if rw is not None else None
        r.kfcm, r.tolcm = float(kfcm), float(tolcm)
        r.offset = (C.c_double * 3)(*[float(x) for x in offset])
        r.do_angles = 1 if (lig_ref is not None and rcpt_ref is not None) else 0
        if r.do_angles:
            if len(lig_ref) != 3 or len(rcpt_ref) != 3:
                raise ValueError("Invalid lists of reference atoms")
            r.lig_ref = (C.c_int32 * 3)(*[int(x) for x in lig_ref])
            r.rcpt_ref = (C.c_int32 * 3)(*[int(x) for x in rcpt_ref])
            r.kfcd = (C.c_double * 3)(*[float(x) for x in kfcd])
            r.a = (C.c_double * 3)(*[float(x) for x in a])
            r.b = (C.c_double * 3)(*[float(x) for x in b])
        _lib.check(self._L.sdm_add_centroid_restraint(self._h, C.byref(r)))

    def add_alignment_restraint(self, liga_ref, ligb_ref, kfdispl, ktheta, kpsi, offset=(0.0, 0.0, 0.0)):
        if len(liga_ref) != 3 or len(ligb_ref) != 3:
            raise ValueError("Invalid lists of reference atoms")
        r = _lib.SdmAlignmentRestraint()
        r.liga_ref = (C.c_int32 * 3)(*[int(x) for x in liga_ref])
        r.ligb_ref = (C.c_int32 * 3)(*[int(x) for x in ligb_ref])
        r.kfdispl, r.ktheta, r.kpsi = float(kfdispl), float(ktheta), float(kpsi)
        r.offset = (C.c_double * 3)(*[float(x) for x in offset])
        _lib.check(self._L.sdm_add_alignment_restraint(self._h, C.byref(r)))

    def clear_restraints(self):
        _lib.check(self._L.sdm_clear_restraints(self._h))

    def set_restraint_control(self, value: float):
        _lib.check(self._L.sdm_set_restraint_control(self._h, float(value)))

    def restraint_energy(self, replica: int = 0) -> float:
        v = C.c_double()
        _lib.check(self._L.sdm_get_restraint_energy(self._h, replica, C.byref(v)))
        return v.value

    def info(self, key: str) -> float:
        v = C.c_double()
        _lib.check(self._L.sdm_get_info(self._h, key.encode(), C.byref(v)))
        return v.value
