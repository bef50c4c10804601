#!/usr/bin/env python
"""Duration of the side chain (displaced-atom rows, gather, exceptions) against the pair kernel it runs beside, per
evaluation, without graph replay (SDMB200_SIDE_TIMING=1 makes the library print both).  --flush: 160 MiB write before
every evaluation (cold L2, like the bench's resident leg).
usage (under gpurun): SDMB200_SIDE_TIMING=1 python tools/side_timing.py [--replicas 16] [--flush] [--workload cfg2]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch   # noqa: E402
import bench   # noqa: E402
from openmm_sdm_plugin_b200.context import SDMContext   # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=16)
ap.add_argument("--flush", action="store_true")
ap.add_argument("--workload", default="cfg2")
a = ap.parse_args()
case, _ = bench.load_case(a.workload)
n = case.system.n_atoms
st = torch.cuda.current_stream()
flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
rng = np.random.default_rng(9)
with SDMContext(case.system, case.displacement, n_replicas=a.replicas, use_graph=False) as c:
    c.set_stream(st.cuda_stream)
    for r in range(a.replicas):
        c.set_alchemical(r, case.alch)
        c.set_positions(r, case.positions + (rng.normal(scale=0.002, size=(n, 3)) if r else 0.0))
    for _ in range(8):
        if a.flush:
            flush.zero_()
        c.eval()
    torch.cuda.synchronize()
