#!/usr/bin/env python
"""cfg2 with nonbondedMethod=PME: ms per evaluation with and without the reciprocal-space part on the device.
usage (under gpurun): python tools/pme_time.py [--only-reciprocal] [--replicas 16,1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch   # noqa: E402
from openmm_sdm_plugin_b200 import system as S   # noqa: E402
from openmm_sdm_plugin_b200.context import SDMContext   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--only-reciprocal", action="store_true")
ap.add_argument("--replicas", default="16,1")
ap.add_argument("--steps", type=int, default=100)
a = ap.parse_args()
for R in [int(x) for x in a.replicas.split(",")]:
    case = S.cfg2()
    case.system.method = S.PME
    stream = torch.cuda.current_stream()
    for recip in ((True,) if a.only_reciprocal else (False, True)):
        with SDMContext(case.system, case.displacement, n_replicas=R) as c:
            c.set_stream(stream.cuda_stream)
            if recip:
                c.enable_reciprocal_pme()
            for r in range(R):
                c.set_alchemical(r, case.alch)
                c.set_positions(r, case.positions)
            for _ in range(25):
                c.eval()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(a.steps):
                c.eval()
            e1.record(stream)
            torch.cuda.synchronize()
            assert c.scalars(0)["status"] == 0
            ms = e0.elapsed_time(e1) / a.steps
            print("R=%d reciprocal=%s ms/eval %.4f evals/s %.0f" % (R, recip, ms, R * 1e3 / ms))
