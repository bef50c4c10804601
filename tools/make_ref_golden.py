#!/usr/bin/env python
"""Golden outputs of the REFERENCE's own step code for the two shipped fixtures.

Run in the build container (needs /root/reference to compile oracle/_ref):

    python tools/make_ref_golden.py

For cfg1 and cfg2 this drives SDMPlugin::LangevinIntegratorSDM::step(1) of the reference
(oracle/_ref/libsdmref.so: the reference's integrator + Reference-platform kernels compiled in
place, see oracle/ref_driver.cpp) with the oracle's restated NonbondedForce as the force-group-2
evaluation and a seeded synthetic bonded force as force group 1, and stores what the reference
computed: BindE, PotEnergy and the hybrid force.  tests/test_gpu_reference_golden.py compares the
CUDA path with these files on the GPU box, where /root/reference does not exist.

Writes tests/golden/ref_step_cfg1.npz and tests/golden/ref_step_cfg2.npz.
"""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openmm_sdm_plugin_b200 import system as S   # noqa: E402
from oracle import oracle as O                    # noqa: E402
from oracle import reference as R                 # noqa: E402

FB_SEED, FB_SCALE, EB = 20261017, 40.0, -321.5


def bonded(n):
    return np.random.default_rng(FB_SEED).normal(scale=FB_SCALE, size=(n, 3))


def states(case):
    """The shipped alchemical settings plus two that exercise the soft core and the other biases."""
    a = case.alch
    return [dataclasses.replace(a),
            dataclasses.replace(a, bias_method=S.ILOGISTIC, softcore_method=S.RATIONAL_SOFTCORE, lambda1=0.1,
                                lambda2=0.45, alpha=0.1 / S.KCAL, u0=-2.0, w0coeff=0.5, umax=12.0, ubcore=-20.0),
            dataclasses.replace(a, bias_method=S.QUADRATIC, softcore_method=S.TANH_SOFTCORE, gammac=0.02, wbcoeff=0.8,
                                w0coeff=-1.0, umax=30.0, ubcore=-10.0)]


def main():
    for case, name in ((S.cfg1(), "ref_step_cfg1.npz"), (S.cfg2(), "ref_step_cfg2.npz")):
        n = case.system.n_atoms
        fb = bonded(n)
        threads = 1 if n < 1000 else O.max_threads()

        def force_fn(groups, pos):
            if groups == 4:
                r = O.nonbonded(case.system, pos, nthreads=threads)
                return r["E"], r["forces"]
            return EB, fb

        out = {"fb_seed": FB_SEED, "fb_scale": FB_SCALE, "eb": EB}
        keep = np.arange(n) if n < 1000 else np.unique(np.concatenate(
            [np.arange(0, n, 16), np.nonzero(np.abs(case.displacement).sum(axis=1))[0]]))
        out["force_rows"] = keep.astype(np.int32)
        for k, alch in enumerate(states(case)):
            res = R.run(case.masses, case.positions, np.zeros((n, 3)), case.displacement,
                        R.params_from_alch(alch), force_fn)
            out["alch%d" % k] = np.array([getattr(alch, f.name) for f in dataclasses.fields(alch)], dtype=np.float64)
            out["bind_e%d" % k] = res["bind_e"]
            out["pot_energy%d" % k] = res["pot_energy"]
            out["hybrid_force%d" % k] = res["hybrid_force"][keep]
            out["force_norm%d" % k] = float(np.sqrt((res["hybrid_force"] ** 2).sum()))
            print(name, k, "BindE %.10f PotEnergy %.8f |F| %.6e" % (res["bind_e"], res["pot_energy"], out["force_norm%d" % k]))
        path = os.path.join(ROOT, "tests", "golden", name)
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())
