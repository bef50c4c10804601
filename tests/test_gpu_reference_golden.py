"""The CUDA path against outputs of the REFERENCE's own step code (tests/golden/ref_step_cfg*.npz,
written by tools/make_ref_golden.py from oracle/_ref: the reference's integrator and Reference-platform
kernels compiled in place, with the oracle's restated NonbondedForce as OpenMM's part).

Tolerances (BASELINE.json north_star): BindE (= soft-core of u, from the FP64 moved-pair path)
<= 1e-6 * max(1, |.|) absolute; PotEnergy <= 1e-5 relative; hybrid force <= 1e-4 RMS relative.
"""
import dataclasses
import os

import numpy as np
import pytest

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext

GOLDEN = S.GOLDEN_DIR


def bonded(z, n):
    return np.random.default_rng(int(z["fb_seed"])).normal(scale=float(z["fb_scale"]), size=(n, 3))


def alch_of(z, k):
    fields = [f.name for f in dataclasses.fields(S.AlchemicalState)]
    vals = z["alch%d" % k]
    kw = {}
    for name, v in zip(fields, vals):
        kw[name] = int(v) if name in ("bias_method", "softcore_method", "nonequilibrium") else float(v)
    return S.AlchemicalState(**kw)


CASES = [("cfg1", "ref_step_cfg1.npz", _lib.PAIR_ALLPAIRS), ("cfg1", "ref_step_cfg1.npz", _lib.PAIR_CLUSTER),
         ("cfg2", "ref_step_cfg2.npz", _lib.PAIR_CLUSTER)]


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,npz,mode", CASES)
def test_cuda_path_matches_reference_step(cfg, npz, mode):
    case = getattr(S, cfg)()
    z = np.load(os.path.join(GOLDEN, npz))
    n = case.system.n_atoms
    fb, eb = bonded(z, n), float(z["eb"])
    rows = z["force_rows"]
    nstates = 3
    with SDMContext(case.system, case.displacement, n_replicas=nstates, pair_mode=mode) as ctx:
        for k in range(nstates):                       # the three alchemical states as three replicas
            ctx.set_positions(k, case.positions)
            ctx.set_alchemical(k, alch_of(z, k))
            ctx.set_bonded_forces(k, fb, eb)
        ctx.eval()
        for k in range(nstates):
            sc = ctx.scalars(k)
            assert sc["status"] == 0
            bind_e, pot = float(z["bind_e%d" % k]), float(z["pot_energy%d" % k])
            assert abs(sc["bind_e"] - bind_e) <= 1e-6 * max(1.0, abs(bind_e)), (k, sc["bind_e"], bind_e)
            assert abs(sc["pot_energy"] - pot) <= 1e-5 * abs(pot), (k, sc["pot_energy"], pot)
            f = ctx.forces(k, _lib.FORCE_HYBRID)[rows]
            ref = z["hybrid_force%d" % k]
            err = float(np.sqrt(((f - ref) ** 2).sum() / (ref ** 2).sum()))
            assert err <= 1e-4, (k, err)


def test_oracle_matches_reference_step_golden():
    """CPU: the restated oracle reproduces what the reference's own step code computed (cfg1: the
    generator ran the oracle's nonbonded single-threaded, so the match is bit for bit)."""
    from oracle import oracle as O
    case = S.cfg1()
    z = np.load(os.path.join(GOLDEN, "ref_step_cfg1.npz"))
    fb, eb = bonded(z, case.system.n_atoms), float(z["eb"])
    for k in range(3):
        r = O.sdm_eval(case.system, alch_of(z, k), case.displacement, case.positions, fb=fb, eb=eb, nthreads=1)
        assert r["bind_e"] == float(z["bind_e%d" % k])
        assert r["pot_energy"] == float(z["pot_energy%d" % k])
        assert np.array_equal(r["forces"][z["force_rows"]], z["hybrid_force%d" % k])
