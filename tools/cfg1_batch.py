#!/usr/bin/env python
"""cfg1 (230 atoms, all-pairs path) with a large replica batch: ms per evaluation of the batch.
usage (under gpurun): python tools/cfg1_batch.py [--replicas 512] [--steps 50]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch   # noqa: E402
from openmm_sdm_plugin_b200 import system as S   # noqa: E402
from openmm_sdm_plugin_b200.context import SDMContext   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=512)
ap.add_argument("--steps", type=int, default=50)
a = ap.parse_args()
case = S.cfg1()
stream = torch.cuda.current_stream()
rng = np.random.default_rng(3)
with SDMContext(case.system, case.displacement, n_replicas=a.replicas) as c:
    c.set_stream(stream.cuda_stream)
    for r in range(a.replicas):
        c.set_alchemical(r, case.alch)
        c.set_positions(r, case.positions + rng.normal(scale=0.002, size=case.positions.shape))
    for _ in range(5):
        c.eval()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        c.eval()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    assert c.scalars(0)["status"] == 0
    print("cfg1 R=%d  ms/step %.4f  evals/s %.0f" % (a.replicas, ms, a.replicas / ms * 1e3))
