#!/usr/bin/env python
"""HBM roofline of the elementwise kernel-interface ops (displace 48 B/atom,
mix 64 B/atom, ...) on arrays larger than L2.  Prints one JSON line per op.
usage (B200): python tools/bench_elementwise.py [n_atoms]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmm_sdm_plugin_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000     # 256 MB per float4 array (> 126 MB L2)
L = _lib.lib()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
mk = lambda: torch.randn(n, 4, dtype=torch.float32, device="cuda")
a, b, c, d, e = mk(), mk(), mk(), mk(), mk()
b[:, 3] = 0
e[:, 3] = 0.1
rnd = torch.randn(n + 64, 4, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ops = [
    ("make_state2 (langevin.cl:198-206)", 48, lambda: L.sdm_k_make_state2(st, n, a.data_ptr(), b.data_ptr())),
    ("save_state1 (langevin.cl:161-178)", 64, lambda: L.sdm_k_save_state1(st, n, a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr())),
    ("save_state2 / restore_state1 (langevin.cl:183-192,147-155)", 32, lambda: L.sdm_k_save_state2(st, n, a.data_ptr(), c.data_ptr())),
    ("hybrid_force (langevin.cl:72-87)", 64, lambda: L.sdm_k_hybrid_force(st, n, a.data_ptr(), b.data_ptr(), c.data_ptr(), 0.37)),
    ("langevin_part1 (langevin.cl:7-31)", 80, lambda: L.sdm_k_langevin_part1(st, n, e.data_ptr(), a.data_ptr(), d.data_ptr(), 0.9995, 0.001, 0.05, 0.001, rnd.data_ptr(), 3)),
    ("langevin_part2 (langevin.cl:37-69)", 80, lambda: L.sdm_k_langevin_part2(st, n, a.data_ptr(), d.data_ptr(), e.data_ptr(), 0.001)),
]
for name, bytes_per_atom, fn in ops:
    for _ in range(3):
        _lib.check(fn())
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for s0, s1 in ev:
        s0.record()
        _lib.check(fn())
        s1.record()
    torch.cuda.synchronize()
    ms = min(x.elapsed_time(y) for x, y in ev)
    gbs = bytes_per_atom * n / (ms * 1e-3) / 1e9
    print(json.dumps({"op": name, "n_atoms": n, "bytes_per_atom": bytes_per_atom, "ms": ms, "achieved_gbs": gbs,
                      "peak_gbs": peak, "frac": gbs / peak, "bound": "hbm"}))
