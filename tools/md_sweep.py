#!/usr/bin/env python
"""Sweep of the pair list's skin and rebuild interval under REAL dynamics (bench.py md_leg): constrained
Langevin MD of the explicit-solvent fixture, 16 replicas.  usage (under gpurun): python tools/md_sweep.py"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench   # noqa: E402
import torch   # noqa: E402
from openmm_sdm_plugin_b200 import system as S   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--replicas", type=int, default=16)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--dt", type=float, default=0.001)
ap.add_argument("--free-solute", action="store_true", help="let the solute move under nonbonded forces alone (no bonded terms)")
ap.add_argument("--grid", default="0.06:20,0.08:20,0.10:20,0.10:40,0.12:40,0.14:40,0.16:80")
a = ap.parse_args()
case, _ = bench.load_case("cfg2")
args = argparse.Namespace(skin=0.06, nstlist=20, steps=a.steps, pair_mode=0)
stream = torch.cuda.current_stream()
flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
states = S.atm_lambda_schedule(22)
for item in a.grid.split(","):
    skin, nst = item.split(":")
    r = bench.md_leg(case, a.replicas, args, 0, stream, flush, states, 0, skin=float(skin), nstlist=int(nst), steps=a.steps, dt=a.dt,
                      freeze_solute=not a.free_solute)
    print(json.dumps({k: r[k] for k in ("skin_nm", "nstlist", "dt_ps", "ms_per_step", "ns_per_day_per_replica", "list_builds",
                                        "steps_repeated_stale_list", "kinetic_temperature_K", "frozen_solute_atoms", "status_ok")}))
