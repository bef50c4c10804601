#!/usr/bin/env python
"""Extract the two shipped Desmond-DMS fixtures of the reference into flat .npz files.

Run in the build container (needs /root/reference, which does NOT exist on the GPU box):

    python tools/extract_dms.py

Writes tests/golden/cfg1_oa_g6_g3.npz and tests/golden/cfg2_temoa_g1_g4.npz.  The SQL and
the unit conversions restate what the reference's reader hands to OpenMM's NonbondedForce
(example/desmonddmsfile75.py:772-850: charge, sigma*angstrom, epsilon*kcal/mol, every
`exclusion` row -> a zero exception, every `pair_12_6_es_term` row -> a 1-4 exception with
epsilon=b^2/4a, sigma=(a/b)^(1/6); :393-396 box from global_cell; :206-233 positions in
angstrom, velocities in angstrom/ps).  Only python's sqlite3 + numpy are used -- OpenMM is
not installed here.

Units in the output are OpenMM's: nm, kJ/mol, e, amu, nm/ps.
"""
import os
import sqlite3
import sys

import numpy as np

REF = "/root/reference/example"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

KCAL = 4.184  # kilocalorie_per_mole -> kilojoule_per_mole (simtk.unit)
ANG = 0.1     # angstrom -> nanometer


def extract(dms_path):
    conn = sqlite3.connect("file:%s?mode=ro" % dms_path, uri=True)
    tables = [r[0] for r in conn.execute("select name from sqlite_master where type='table'")]
    rows = conn.execute(
        "SELECT id, x, y, z, vx, vy, vz, mass, resid, anum FROM particle ORDER BY id").fetchall()
    n = len(rows)
    assert [r[0] for r in rows] == list(range(n))
    pos = np.array([[r[1], r[2], r[3]] for r in rows], dtype=np.float64) * ANG
    vel = np.array([[r[4], r[5], r[6]] for r in rows], dtype=np.float64) * ANG
    mass = np.array([r[7] for r in rows], dtype=np.float64)
    resid = np.array([r[8] for r in rows], dtype=np.int32)
    anum = np.array([r[9] for r in rows], dtype=np.int32)

    q = """SELECT charge, sigma, epsilon FROM particle INNER JOIN nonbonded_param
           ON particle.nbtype=nonbonded_param.id ORDER BY particle.id"""
    nb = np.array(conn.execute(q).fetchall(), dtype=np.float64)
    charge = nb[:, 0].copy()
    sigma = nb[:, 1] * ANG
    epsilon = nb[:, 2] * KCAL

    excl = np.array(conn.execute("SELECT p0, p1 FROM exclusion").fetchall(), dtype=np.int32)
    excl = excl.reshape(-1, 2)

    exc_pairs, exc_params = [], []
    if "pair_12_6_es_term" in tables:
        q = """SELECT p0, p1, aij, bij, qij FROM pair_12_6_es_term INNER JOIN pair_12_6_es_param
               ON pair_12_6_es_term.param=pair_12_6_es_param.id"""
        for p0, p1, a_ij, b_ij, q_ij in conn.execute(q):
            a = a_ij * KCAL * ANG ** 12
            b = b_ij * KCAL * ANG ** 6
            if a == 0.0 or b == 0.0:
                eps, sig = 0.0, 1.0
            else:
                eps = b * b / (4 * a)
                sig = (a / b) ** (1.0 / 6.0)
            exc_pairs.append((p0, p1))
            exc_params.append((q_ij, sig, eps))
    exc_pairs = np.array(exc_pairs, dtype=np.int32).reshape(-1, 2)
    exc_params = np.array(exc_params, dtype=np.float64).reshape(-1, 3)
    # reference requires every 1-4 pair to also be an exclusion (desmonddmsfile75.py:838-846)
    es = {(min(a, b), max(a, b)) for a, b in excl.tolist()}
    assert all((min(a, b), max(a, b)) in es for a, b in exc_pairs.tolist())

    # distance constraints, in the order the reference's reader adds them to the System
    # (desmonddmsfile75.py:542-571 constrained stretches, :573-596 constrained angles -> a 1-3
    # distance, :603-637 constraint_a* rows that are not bonds, constraint_hoh -> the H-H distance)
    cons, bonds, angle_c = [], {}, set()
    if "stretch_harm_term" in tables:
        q = """SELECT p0, p1, r0, fc, constrained FROM stretch_harm_term INNER JOIN stretch_harm_param
               ON stretch_harm_term.param=stretch_harm_param.id"""
        for p0, p1, r0, fc, constrained in conn.execute(q):
            if constrained:
                cons.append((p0, p1, r0 * ANG))
            bonds[(p0, p1)] = bonds[(p1, p0)] = r0 * ANG
    if "angle_harm_term" in tables:
        q = """SELECT p0, p1, p2, theta0, fc, constrained FROM angle_harm_term INNER JOIN angle_harm_param
               ON angle_harm_term.param=angle_harm_param.id"""
        for p0, p1, p2, theta0, fc, constrained in conn.execute(q):
            if constrained:
                l1, l2 = bonds[(p1, p0)], bonds[(p1, p2)]
                cons.append((p0, p2, float(np.sqrt(l1 * l1 + l2 * l2 - 2 * l1 * l2 * np.cos(np.deg2rad(theta0))))))
                angle_c.add((p1, p0, p2))
    for t in sorted(n for n in tables if n.startswith("constraint_a") and n.endswith("term")):
        q = "SELECT p0, p1, r1 FROM %s INNER JOIN %s ON %s.param=%s.id" % (t, t.replace("term", "param"), t, t.replace("term", "param"))
        for p0, p1, r1 in conn.execute(q):
            if (p0, p1) not in bonds:
                cons.append((p0, p1, r1 * ANG))
                bonds[(p0, p1)] = bonds[(p1, p0)] = r1 * ANG
    if "constraint_hoh_term" in tables:
        q = """SELECT p0, p1, p2, r1, r2, theta FROM constraint_hoh_term INNER JOIN constraint_hoh_param
               ON constraint_hoh_term.param=constraint_hoh_param.id"""
        for p0, p1, p2, r1, r2, theta in conn.execute(q):
            if (p0, p1, p2) not in angle_c:
                r1, r2 = r1 * ANG, r2 * ANG
                cons.append((p1, p2, float(np.sqrt(r1 * r1 + r2 * r2 - 2 * r1 * r2 * np.cos(np.deg2rad(theta))))))
    cons_pairs = np.array([(a, b) for a, b, _ in cons], dtype=np.int32).reshape(-1, 2)
    cons_dist = np.array([d for _, _, d in cons], dtype=np.float64)

    box = np.zeros(3)
    if "global_cell" in tables:
        cell = conn.execute("SELECT x, y, z FROM global_cell ORDER BY id").fetchall()
        if len(cell) == 3:
            box = np.array([cell[0][0], cell[1][1], cell[2][2]], dtype=np.float64) * ANG
    conn.close()
    return dict(positions=pos, velocities=vel, masses=mass, resid=resid, anum=anum,
                charge=charge, sigma=sigma, epsilon=epsilon, exclusions=excl,
                exception_pairs=exc_pairs, exception_params=exc_params, box=box,
                constraint_pairs=cons_pairs, constraint_dist=cons_dist)


def main():
    os.makedirs(OUT, exist_ok=True)
    jobs = [("oa-g6-g3-align-restr_0_displaced.dms", "cfg1_oa_g6_g3.npz"),
            ("temoa-g1-g4.dms", "cfg2_temoa_g1_g4.npz")]
    for src, dst in jobs:
        d = extract(os.path.join(REF, src))
        np.savez_compressed(os.path.join(OUT, dst), **d)
        print(dst, "atoms", len(d["charge"]), "excl", len(d["exclusions"]),
              "exc", len(d["exception_pairs"]), "constraints", len(d["constraint_dist"]), "box", d["box"],
              "bytes", os.path.getsize(os.path.join(OUT, dst)))


if __name__ == "__main__":
    sys.exit(main())
