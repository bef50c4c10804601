#!/usr/bin/env python
"""Diagnostic (GPU box): where the time of one end-to-end step goes.  Drives the cfg2 replicas as
G contexts on G streams like bench.py's e2e leg and brackets every stage of every group with CUDA
events; prints, per group, when H2D / eval / D2H started and ended relative to the step start,
plus the host time spent in each call.  Usage: python tools/e2e_timeline.py [G ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from openmm_sdm_plugin_b200 import system as S
    from openmm_sdm_plugin_b200.context import PinnedArray, SDMContext

    groups = [int(a) for a in sys.argv[1:]] or [1, 2]
    R = 16
    case = S.cfg2()
    n = case.system.n_atoms
    rng = np.random.default_rng(1234)
    h_pos, h_pos_b, h_f = PinnedArray((R, n, 3)), PinnedArray((R, n, 3)), PinnedArray((R, n, 3))
    base = np.stack([case.positions + rng.normal(scale=0.002, size=(n, 3)) for _ in range(R)])
    h_pos.array[...] = base
    h_pos_b.array[...] = base + rng.normal(scale=0.0006, size=(R, n, 3))
    states = S.atm_lambda_schedule(R)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for G in groups:
        Rg = R // G
        streams = [torch.cuda.Stream() for _ in range(G)]
        ctxs = []
        for g in range(G):
            c = SDMContext(case.system, case.displacement, n_replicas=Rg, device=0)
            c.set_stream(streams[g].cuda_stream)
            for r in range(Rg):
                c.set_alchemical(r, states[(g * Rg + r) % len(states)])
            ctxs.append(c)
        nsteps, warm = 12, 6
        rows = []
        for k in range(warm + nsteps):
            src = h_pos.array if k % 2 == 0 else h_pos_b.array
            flush.zero_()
            torch.cuda.synchronize()
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(G)]
            start = torch.cuda.Event(enable_timing=True)
            start.record(torch.cuda.current_stream())
            for s in streams:
                s.wait_event(start)
            host = []
            t0 = time.perf_counter()
            for g, c in enumerate(ctxs):
                ev[g][0].record(streams[g])
                a = time.perf_counter()
                c.set_positions_all(src[g * Rg:(g + 1) * Rg])
                b = time.perf_counter()
                ev[g][1].record(streams[g])
                c.eval()
                d = time.perf_counter()
                ev[g][2].record(streams[g])
                c.enqueue_results(h_f.array[g * Rg:(g + 1) * Rg])
                e = time.perf_counter()
                ev[g][3].record(streams[g])
                host.append((b - a, d - b, e - d))
            for c in ctxs:
                c.synchronize()
                c.collect_scalars()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            if k >= warm:
                rows.append((wall, [[start.elapsed_time(x) for x in ev[g]] for g in range(G)], host))
        print("== G = %d  (R/G = %d)   wall per step: median %.3f ms" % (G, Rg, 1e3 * np.median([r[0] for r in rows])))
        for g in range(G):
            t = np.median(np.array([r[1][g] for r in rows]), axis=0)
            h = 1e3 * np.median(np.array([r[2][g] for r in rows]), axis=0)
            print("   group %d: h2d %.3f-%.3f  eval %.3f-%.3f  d2h %.3f-%.3f ms | host call ms: set_positions %.3f eval %.3f enqueue %.3f"
                  % (g, t[0], t[1], t[1], t[2], t[2], t[3], h[0], h[1], h[2]))
        for c in ctxs:
            c.close()


if __name__ == "__main__":
    main()
