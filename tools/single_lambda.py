#!/usr/bin/env python
"""One resident replica (BASELINE.json configs[1] as worded: a single lambda on one B200): ms per
evaluation, and with --launches the per-kernel device times of a few evaluations (CUDA events).
usage (under gpurun): python tools/single_lambda.py [--replicas 1] [--steps 200]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch   # noqa: E402
from openmm_sdm_plugin_b200 import system as S   # noqa: E402
from openmm_sdm_plugin_b200.context import SDMContext   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=1)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--nstlist", type=int, default=20)
ap.add_argument("--skin", type=float, default=-1.0)
a = ap.parse_args()
case = S.cfg2()
stream = torch.cuda.current_stream()
with SDMContext(case.system, case.displacement, n_replicas=a.replicas, nstlist=a.nstlist, skin=a.skin) as c:
    c.set_stream(stream.cuda_stream)
    for r in range(a.replicas):
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
    ms = e0.elapsed_time(e1) / a.steps
    # without list builds: nstlist evaluations between two builds
    c.invalidate_list()
    c.eval()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(a.nstlist - 2):
        c.eval()
    e1.record(stream)
    torch.cuda.synchronize()
    ms_nobuild = e0.elapsed_time(e1) / (a.nstlist - 2)
    c.set_timing(True)
    c.eval()
    torch.cuda.synchronize()
    pair = c.last_timing()[0]
    assert c.scalars(0)["status"] == 0
    print("skin %.2f nstlist %d " % (a.skin, a.nstlist), end="")
    print("replicas %d  ms/eval %.4f (%.0f evals/s/replica-batch)  between builds %.4f  pair kernel %.4f  units %d  env %s" % (
        a.replicas, ms, 1e3 / ms, ms_nobuild, pair, int(c.info("n_units")),
        {k: v for k, v in os.environ.items() if k.startswith("SDMB200_")}))
