#!/usr/bin/env python
"""Host model of the cluster-pair list: tiles T, entries E and the pair kernel's cost for a cluster
layout, without a GPU (numpy + scipy's cKDTree on the 20 k-atom fixture).

    python tools/list_model.py

The pair kernel's time follows  cost = 84*T + 213*E  scheduler-clocks per replica (DESIGN.md
section 8: 84 clocks per 8x8 tile, 213 per (supercluster, j-cluster) entry, from the diagnostic
builds), so a layout can be judged before any device code is written.  This is how the column
layout of nblist_core.h was chosen: predicted -11.5 %, measured -9.2 % kernel time.

Layouts modelled:
  A  geometric 3-D cells of ~52 atoms, kd split (z, y, x) inside a cell, superclusters = runs of
     <= 8 clusters of a cell (the list before v33)
  B  xy columns cut along z into chunks of exactly 64 atoms, the same kd split inside a chunk
     (the list since v33)
"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmm_sdm_plugin_b200 import system as S   # noqa: E402

TILE_CLK, ENTRY_CLK = 84.0, 213.0


def kd8(ids, pos):
    """clusters of 8 from <= 64 atoms: balanced splits along z, y, x at multiples of 8"""
    out = []

    def rec(a, dims):
        if len(a) <= 8 or not dims:
            out.extend(a[k:k + 8] for k in range(0, len(a), 8))
            return
        o = a[np.argsort(pos[a, dims[0]], kind="stable")]
        half = ((len(o) + 7) // 8 + 1) // 2 * 8
        rec(o[:half], dims[1:])
        rec(o[half:], dims[1:])
    rec(np.asarray(ids), [2, 1, 0])
    return out


def layout_cells(pos, box, side):
    nc = np.maximum(1, np.floor(box / side)).astype(int)
    c = np.minimum((pos / (box / nc)).astype(int), nc - 1)
    cid = (c[:, 2] * nc[1] + c[:, 1]) * nc[0] + c[:, 0]
    order = np.argsort(cid, kind="stable")
    bounds = np.searchsorted(cid[order], np.arange(nc.prod() + 1))
    scs = []
    for k in range(nc.prod()):
        ids = order[bounds[k]:bounds[k + 1]]
        if len(ids):
            cls = kd8(ids, pos)
            scs += [cls[s:s + 8] for s in range(0, len(cls), 8)]
    return scs


def layout_columns(pos, box, side, chunk=64):
    ncx, ncy = max(1, int(round(box[0] / side))), max(1, int(round(box[1] / side)))
    cx = np.minimum((pos[:, 0] / (box[0] / ncx)).astype(int), ncx - 1)
    cy = np.minimum((pos[:, 1] / (box[1] / ncy)).astype(int), ncy - 1)
    col = cy * ncx + cx
    scs = []
    for k in range(ncx * ncy):
        ids = np.nonzero(col == k)[0]
        ids = ids[np.argsort(pos[ids, 2], kind="stable")]
        scs += [kd8(ids[s:s + chunk], pos) for s in range(0, len(ids), chunk)]
    return scs


def stats(name, scs, n, pairs, n_in_cutoff):
    cl_of, sc_of_cl = np.full(n, -1), []
    for si, sc in enumerate(scs):
        for cl in sc:
            cl_of[cl] = len(sc_of_cl)
            sc_of_cl.append(si)
    ncl, sc_of_cl = len(sc_of_cl), np.array(sc_of_cl)
    a, b = cl_of[pairs[:, 0]], cl_of[pairs[:, 1]]
    tiles = np.unique(np.minimum(a, b).astype(np.int64) * ncl + np.maximum(a, b))   # lower index owns
    entries = np.unique(sc_of_cl[tiles // ncl].astype(np.int64) * ncl + tiles % ncl)
    T, E = len(tiles), len(entries)
    print("%-32s slots %6d (%4.1f %% dummies)  tiles %7d  entries %6d  tiles/entry %.2f  useful pairs %.4f"
          "  cost %.2f Mclk" % (name, 8 * ncl, 100 * (1 - n / (8.0 * ncl)), T, E, T / E, n_in_cutoff / (64.0 * T),
                              (TILE_CLK * T + ENTRY_CLK * E) / 1e6))
    return T, E


def main():
    case = S.cfg2()
    box = case.system.box
    pos = np.mod(case.positions, box)
    n, rc, skin = len(pos), case.system.cutoff, 0.06
    pairs = cKDTree(pos, boxsize=box).query_pairs(rc + skin, output_type="ndarray")
    d = pos[pairs[:, 0]] - pos[pairs[:, 1]]
    d -= box * np.round(d / box)
    n_in = int(((d * d).sum(1) <= rc * rc).sum())
    density = n / box.prod()
    print("atoms %d, pairs within rlist %d, within the cutoff %d (exclusions ignored)" % (n, len(pairs), n_in))
    stats("A cells (~52 atoms)", layout_cells(pos, box, max(np.cbrt(40.0 / density), 0.5 * (rc + skin))), n, pairs, n_in)
    for side in (0.75, np.cbrt(64.0 / density), 1.0):
        stats("B columns, side %.2f nm" % side, layout_columns(pos, box, side), n, pairs, n_in)


if __name__ == "__main__":
    main()
