#!/usr/bin/env python
"""Long constrained-dynamics runs of the explicit-solvent fixture on the device (sdm_md_step): kinetic temperature,
list builds, repeated steps and status over thousands of steps, reaction field and complete PME.
usage (under gpurun): python tools/md_soak.py [--steps 5000] [--replicas 16]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench   # noqa: E402
import torch   # noqa: E402
from openmm_sdm_plugin_b200 import system as S   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5000)
ap.add_argument("--replicas", type=int, default=16)
a = ap.parse_args()
args = argparse.Namespace(skin=0.2, nstlist=40, steps=a.steps, pair_mode=0)
stream = torch.cuda.current_stream()
flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
states = S.atm_lambda_schedule(22)
for wl, recip in (("cfg2", False), ("cfg2:pme+reciprocal", True)):
    case, name = bench.load_case(wl)
    r = bench.md_leg(case, a.replicas, args, 0, stream, flush, states, 0, skin=0.2, nstlist=40, steps=a.steps,
                     reciprocal_pme=recip)
    r.pop("note", None)
    r["workload"] = wl
    print(json.dumps(r))
