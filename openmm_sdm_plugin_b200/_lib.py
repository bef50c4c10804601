"""ctypes binding of libsdmb200.so (include/sdmb200.h).

The library is the product: if it is missing this module raises -- there is no Python or CPU
fallback for any force arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# SDMB200_LIB: development override used to A/B kernel build variants on the GPU box
LIB_PATH = os.environ.get("SDMB200_LIB") or os.path.join(_HERE, "libsdmb200.so")
CSRC = os.path.join(_HERE, "csrc")

SDM_OK = 0
SDM_ERR_INVALID, SDM_ERR_NO_DEVICE, SDM_ERR_CUDA, SDM_ERR_BOX = -1, -2, -3, -4
SDM_ERR_SOFTCORE, SDM_ERR_STALE_LIST, SDM_ERR_CAPACITY, SDM_ERR_CONSTRAINT = -5, -6, -7, -8
FORCE_HYBRID, FORCE_STATE1, FORCE_STATE2, FORCE_DELTA = 0, 1, 2, 3
PAIR_AUTO, PAIR_ALLPAIRS, PAIR_CLUSTER = 0, 1, 2


class SdmSystem(C.Structure):
    _fields_ = [("n_atoms", C.c_int32), ("method", C.c_int32), ("cutoff", C.c_double),
                ("eps_rf", C.c_double), ("box", C.c_double * 3),
                ("use_dispersion_correction", C.c_int32), ("n_exclusions", C.c_int32),
                ("n_exceptions", C.c_int32), ("n_replicas", C.c_int32),
                ("charge", C.c_void_p), ("sigma", C.c_void_p), ("epsilon", C.c_void_p),
                ("exclusions", C.c_void_p), ("exceptions", C.c_void_p),
                ("exception_params", C.c_void_p), ("displacement", C.c_void_p),
                ("ewald_alpha", C.c_double), ("ewald_tolerance", C.c_double),
                ("lj_combining", C.c_int32), ("reserved_", C.c_int32)]


class SdmOptions(C.Structure):
    _fields_ = [("device", C.c_int32), ("pair_mode", C.c_int32), ("skin", C.c_double),
                ("nstlist", C.c_int32), ("exact_cutoff", C.c_int32), ("use_graph", C.c_int32),
                ("reserved", C.c_int32 * 7)]


class SdmAlch(C.Structure):
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


class SdmCentroidRestraint(C.Structure):
    _fields_ = [("n_lig_cm", C.c_int32), ("n_rcpt_cm", C.c_int32),
                ("lig_cm_atoms", C.POINTER(C.c_int32)), ("rcpt_cm_atoms", C.POINTER(C.c_int32)),
                ("lig_cm_weights", C.POINTER(C.c_double)), ("rcpt_cm_weights", C.POINTER(C.c_double)),
                ("kfcm", C.c_double), ("tolcm", C.c_double), ("offset", C.c_double * 3),
                ("do_angles", C.c_int32), ("rcpt_ref", C.c_int32 * 3), ("lig_ref", C.c_int32 * 3),
                ("kfcd", C.c_double * 3), ("a", C.c_double * 3), ("b", C.c_double * 3)]


class SdmAlignmentRestraint(C.Structure):
    _fields_ = [("liga_ref", C.c_int32 * 3), ("ligb_ref", C.c_int32 * 3),
                ("kfdispl", C.c_double), ("ktheta", C.c_double), ("kpsi", C.c_double), ("offset", C.c_double * 3)]


class SdmScalars(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("E1", "E2", "Eb", "u", "u_sc", "fp", "ebias", "bfp", "sp", "pot_energy",
                 "bind_e", "E1_pair", "E1_exc", "E1_disp")] + \
               [("n_pairs1", C.c_int64), ("n_moved1", C.c_int64), ("n_moved2", C.c_int64),
                ("status", C.c_int32), ("list_age", C.c_int32)]


# every symbol include/sdmb200.h declares: name -> (restype, argtypes)
_VP, _I, _D = C.c_void_p, C.c_int, C.c_double
SYMBOLS = {
    "sdm_abi_version": (_I, []),
    "sdm_last_error": (C.c_char_p, []),
    "sdm_device_count": (_I, []),
    "sdm_default_options": (None, [C.POINTER(SdmOptions)]),
    "sdm_default_alch": (None, [C.POINTER(SdmAlch)]),
    "sdm_create": (_I, [C.POINTER(SdmSystem), C.POINTER(SdmOptions), C.POINTER(_VP)]),
    "sdm_destroy": (None, [_VP]),
    "sdm_set_stream": (_I, [_VP, _VP]),
    "sdm_synchronize": (_I, [_VP]),
    "sdm_host_alloc": (_I, [C.POINTER(_VP), C.c_uint64]),
    "sdm_host_free": (_I, [_VP]),
    "sdm_set_positions": (_I, [_VP, _I, _VP]),
    "sdm_set_positions_device": (_I, [_VP, _I, _VP]),
    "sdm_positions_device_ptr": (_I, [_VP, _I, C.POINTER(_VP)]),
    "sdm_set_positions_all": (_I, [_VP, _VP]),
    "sdm_set_positions_all_f32": (_I, [_VP, _VP]),
    "sdm_read_results": (_I, [_VP, _VP, _VP]),
    "sdm_enqueue_results": (_I, [_VP, _VP]),
    "sdm_enqueue_results_f32": (_I, [_VP, _VP]),
    "sdm_collect_scalars": (_I, [_VP, _VP]),
    "sdm_set_bonded_forces": (_I, [_VP, _I, _VP, _D]),
    "sdm_set_alchemical": (_I, [_VP, _I, C.POINTER(SdmAlch)]),
    "sdm_get_alchemical": (_I, [_VP, _I, C.POINTER(SdmAlch)]),
    "sdm_set_displacement": (_I, [_VP, _VP]),
    "sdm_eval": (_I, [_VP]),
    "sdm_invalidate_list": (_I, [_VP]),
    "sdm_get_scalars": (_I, [_VP, _I, C.POINTER(SdmScalars)]),
    "sdm_get_forces": (_I, [_VP, _I, _I, _VP]),
    "sdm_forces_device_ptr": (_I, [_VP, _I, C.POINTER(_VP)]),
    "sdm_get_pairs": (_I, [_VP, _I, _VP, C.c_int64, C.POINTER(C.c_int64)]),
    "sdm_get_launch_count": (_I, [_VP, C.POINTER(C.c_int64)]),
    "sdm_set_timing": (_I, [_VP, _I]),
    "sdm_get_last_timing": (_I, [_VP, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "sdm_get_info": (_I, [_VP, C.c_char_p, C.POINTER(_D)]),
    "sdm_k_make_state2": (_I, [_VP, _I, _VP, _VP]),
    "sdm_k_save_state1": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "sdm_k_save_state2": (_I, [_VP, _I, _VP, _VP]),
    "sdm_k_restore_state1": (_I, [_VP, _I, _VP, _VP]),
    "sdm_k_hybrid_force": (_I, [_VP, _I, _VP, _VP, _VP, C.c_float]),
    "sdm_k_langevin_part1": (_I, [_VP, _I, _VP, _VP, _VP, C.c_float, C.c_float, C.c_float, C.c_float, _VP, C.c_uint32]),
    "sdm_k_langevin_part2": (_I, [_VP, _I, _VP, _VP, _VP, C.c_float]),
    "sdm_langevin_params": (_I, [_D, _D, _D, C.POINTER(_D), C.POINTER(_D), C.POINTER(_D)]),
    "sdm_execute_scalars": (_I, [C.POINTER(SdmAlch), _D, _D, _D, C.POINTER(SdmScalars)]),
    "sdm_md_init": (_I, [_VP, _VP, _D, _D, _D, C.c_uint64]),
    "sdm_md_set_velocities": (_I, [_VP, _I, _VP]),
    "sdm_md_get_velocities": (_I, [_VP, _I, _VP]),
    "sdm_get_positions": (_I, [_VP, _I, _VP]),
    "sdm_md_set_constraints": (_I, [_VP, C.c_int32, _VP, _VP, _D]),
    "sdm_md_step": (_I, [_VP, _I]),
    "sdm_md_get_counters": (_I, [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "sdm_md_update": (_I, [_VP, _VP]),
    "sdm_md_set_noise": (_I, [_VP, _VP]),
    "sdm_md_kinetic_energy": (_I, [_VP, _I, C.POINTER(_D)]),
    "sdm_enable_reciprocal_pme": (_I, [_VP, _VP]),
    "sdm_set_external_dual": (_I, [_VP, _I, _VP, _VP, _D, _D]),
    "sdm_enable_hct_gb": (_I, [_VP, _VP, _VP, _VP, _D, _D, _I]),
    "sdm_get_born_radii": (_I, [_VP, _I, _I, _VP]),
    "sdm_add_centroid_restraint": (_I, [_VP, C.POINTER(SdmCentroidRestraint)]),
    "sdm_add_alignment_restraint": (_I, [_VP, C.POINTER(SdmAlignmentRestraint)]),
    "sdm_clear_restraints": (_I, [_VP]),
    "sdm_set_restraint_control": (_I, [_VP, _D]),
    "sdm_get_restraint_energy": (_I, [_VP, _I, C.POINTER(_D)]),
}

_LIB = None


class SDMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsdmb200 error %d: %s" % (code, msg))
        self.code = code


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile libsdmb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    args = ["make", "-C", CSRC, "-j8"]
    if force:
        args.append("-B")
    out = subprocess.run(args, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libsdmb200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib():
    """The loaded C-ABI library.  Raises if it has not been built -- never falls back."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libsdmb200.so is missing (%s); run `python -c 'import __graft_entry__ as g; "
                "g.build()'`. There is no CPU fallback for the SDM force path." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if L.sdm_abi_version() != 2:
            raise ImportError("libsdmb200.so ABI version mismatch")
        _LIB = L
    return _LIB


def check(rc: int):
    if rc != SDM_OK:
        msg = lib().sdm_last_error()
        raise SDMError(rc, msg.decode() if msg else "")
    return rc
