#!/usr/bin/env python
"""What a list build costs: host wall clock and device time of (a) a replayed evaluation, (b) the
evaluation that rebuilds the list, (c) the evaluation after it (graph capture + instantiation), for
cfg2 with R replicas.  usage (under gpurun): python tools/build_cost.py [--replicas 16] [--reps 20]"""
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
ap.add_argument("--replicas", type=int, default=16)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
case = S.cfg2()
stream = torch.cuda.current_stream()


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    fn()
    t1 = time.perf_counter()
    e1.record(stream)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) * 1e3, e0.elapsed_time(e1), (t2 - t0) * 1e3


with SDMContext(case.system, case.displacement, n_replicas=a.replicas, nstlist=1000) as c:
    c.set_stream(stream.cuda_stream)
    for r in range(a.replicas):
        c.set_alchemical(r, case.alch)
        c.set_positions(r, case.positions)
    for _ in range(5):
        c.eval()
    res = {"replay": [], "build": [], "after_build": []}
    for _ in range(a.reps):
        c.invalidate_list()
        res["build"].append(timed(c.eval))
        res["after_build"].append(timed(c.eval))
        res["replay"].append(timed(c.eval))
    assert c.scalars(0)["status"] == 0
    for k, v in res.items():
        m = np.median(np.array(v), axis=0)
        print("R=%d %-12s host submit %.3f ms   device %.3f ms   host until done %.3f ms" % (a.replicas, k, m[0], m[1], m[2]))
