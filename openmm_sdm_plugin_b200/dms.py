"""Desmond DMS (SQLite) reader / writer for the SDM hot path.

Host-side mirror of the part of the reference's `DesmondDMSFile` (example/desmonddmsfile75.py)
that sits on the input and output side of the path: it reads coordinates, velocities, masses,
residue ids, the box and everything the reference hands to OpenMM's NonbondedForce, and it writes
coordinates / velocities / box back into the file.  Same method names and error behaviour as the
reference class; quantities are plain numpy arrays in OpenMM units (nm, nm/ps, amu, e, kJ/mol)
because `simtk.unit` is not a dependency of this package.

What is restated (reference file:line):
  * one or several files, atoms concatenated in file order           desmonddmsfile75.py:51-116
  * IOError for a missing file / a file without tables, ValueError when the particle table has
    no `nbtype` column                                                :83-98
  * positions angstrom -> nm, velocities angstrom/ps -> nm/ps         :206-233
  * box = diagonal of the three `global_cell` rows of the FIRST file  :393-396
  * charge, sigma*angstrom, epsilon*kcal/mol per particle             :772-810
  * every `exclusion` row -> an excluded pair                         :812-819
  * `pair_12_6_es_term` rows -> 1-4 exceptions with eps = b^2/4a, sigma = (a/b)^(1/6),
    (a or b == 0 -> eps 0, sigma 1); every such pair must also be an exclusion  :821-850
  * setPositions / setVelocities / setGlobalCell UPDATE statements    :236-288

Bonded terms, constraints, virtual sites, GB parameters and restraints belong to the force
group the path receives as `Fb` / `Eb` and are not read.
"""
from __future__ import annotations

import os
import sqlite3

import numpy as np

from .system import (CUTOFF_NONPERIODIC, CUTOFF_PERIODIC, EWALD, HCT, KCAL, NOCUTOFF, PME, GBSAHCTForce,
                     NonbondedSystem)

ANGSTROM = 0.1  # nm


class DesmondDMSFile(object):
    """Parses one or more Desmond DMS files into the flat description the SDM path consumes."""

    def __init__(self, file, verbose=False):
        self._verbose = verbose
        self._file = list(file) if isinstance(file, (list, tuple)) else [file]
        self._conn, self._tables, self._natoms = [], [], []
        for f in self._file:
            if not os.path.exists(str(f)):
                raise IOError("No such file or directory: %s" % str(f))
            conn = sqlite3.connect(str(f))
            tables = self._readSchemas(conn)
            if len(tables) == 0:
                conn.close()
                raise IOError("DMS file %s was not loaded sucessfully. No tables found" % str(f))
            if "nbtype" not in tables.get("particle", []):
                conn.close()
                raise ValueError("No nonbonded parameters associated with DMS file %s. You can add a "
                                 "forcefield with the viparr command line tool distributed with desmond"
                                 % str(f))
            self._conn.append(conn)
            self._tables.append(tables)
        self._read_particles()
        self._offset = np.concatenate([[0], np.cumsum(self._natoms)]).astype(int)
        self._bonded_force_group = 1
        self._nonbonded_force_group = 2

    # ---- reading ----------------------------------------------------------------------------
    @staticmethod
    def _readSchemas(conn):
        tables = {}
        for (name,) in conn.execute("SELECT name FROM sqlite_master WHERE type='table'"):
            tables[name] = [row[1] for row in conn.execute("PRAGMA table_info(%s)" % name)]
        return tables

    def _read_particles(self):
        pos, vel, mass, resid, anum, names = [], [], [], [], [], []
        for conn, tables in zip(self._conn, self._tables):
            cols = tables["particle"]
            has_v = all(c in cols for c in ("vx", "vy", "vz"))
            q = "SELECT id, x, y, z, %s, mass, %s, anum, %s FROM particle ORDER BY id" % (
                "vx, vy, vz" if has_v else "0.0, 0.0, 0.0",
                "resid" if "resid" in cols else "0", "name" if "name" in cols else "''")
            rows = conn.execute(q).fetchall()
            if [r[0] for r in rows] != list(range(len(rows))):
                raise ValueError("particle ids of a DMS file must be 0..n-1")
            self._natoms.append(len(rows))
            pos += [(r[1], r[2], r[3]) for r in rows]
            vel += [(r[4], r[5], r[6]) for r in rows]
            mass += [r[7] for r in rows]
            resid += [r[8] for r in rows]
            anum += [r[9] for r in rows]
            names += [r[10] for r in rows]
        self.positions = np.array(pos, dtype=np.float64).reshape(-1, 3) * ANGSTROM
        self.velocities = np.array(vel, dtype=np.float64).reshape(-1, 3) * ANGSTROM
        self.masses = np.array(mass, dtype=np.float64)
        self.resid = np.array(resid, dtype=np.int32)
        self.anum = np.array(anum, dtype=np.int32)
        self.atom_names = names

    def getPositions(self):
        """[n,3] nm."""
        return self.positions

    def getVelocities(self):
        """[n,3] nm/ps."""
        return self.velocities

    def getMasses(self):
        return self.masses

    def getResidueIds(self):
        return self.resid

    def getNumAtoms(self):
        return int(self._offset[-1])

    def getBondedForceGroup(self):
        return self._bonded_force_group

    def getNonBondedForceGroup(self):
        return self._nonbonded_force_group

    def getBox(self):
        """Orthorhombic box edges in nm from the first file's global_cell (zeros if absent)."""
        tables, conn = self._tables[0], self._conn[0]
        if "global_cell" not in tables:
            return np.zeros(3)
        cell = conn.execute("SELECT x, y, z FROM global_cell ORDER BY id").fetchall()
        if len(cell) != 3:
            return np.zeros(3)
        return np.array([cell[0][0], cell[1][1], cell[2][2]], dtype=np.float64) * ANGSTROM

    def _get_gb_params(self):
        """(charge, radius, screened_radius) rows of the `hct` tables as desmonddmsfile75.py:290-313 hands them
        to GBSAHCTForce.addParticle: radius in nm minus 0.009, screened_radius times that.  (finalize() subtracts
        the offset again -- the reference does both, so both are done here.)  None when a file has no table."""
        out = []
        for conn, tables in zip(self._conn, self._tables):
            if "hct" not in tables:
                return None
            for charge, radius, screened_radius in conn.execute("SELECT charge,radius,screened_radius FROM hct ORDER BY id"):
                radius_n = radius * ANGSTROM - 0.009
                out.append((charge, radius_n, screened_radius * radius_n))
        return out

    def createSystem(self, nonbondedMethod=NOCUTOFF, nonbondedCutoff=1.0, reactionFieldDielectric=78.3,
                     useDispersionCorrection=True, ewaldErrorTolerance=0.0005, implicitSolvent=None,
                     OPLS=False) -> NonbondedSystem:
        """The force-group-2 content of the reference's createSystem: the NonbondedForce
        particles, exclusions and 1-4 exceptions (desmonddmsfile75.py:772-850), with the cutoff
        method/distance of :418-426 (Ewald and PME included: direct space here, system.py) and the
        box of :393-396; ewaldErrorTolerance as :426; implicitSolvent=HCT adds GBSAHCTForce(SA='ACE') from the
        `hct` table and sets the reaction-field dielectric to 1 (:441-467).  The AGBNP / GVolSA models are external
        plugins the reference only loads if present (:469-526): not built.  OPLS=True forces the geometric combining
        rule: the reference zeroes every NonbondedForce epsilon and adds a CustomNonbondedForce with
        sigma12 = sqrt(s1 s2), eps12 = sqrt(e1 e2) on the same exclusions and cutoff, and switches both long-range
        corrections off (:780-810, :427-438) -- here one flag of the pair kernels (system.lj_geometric)."""
        if nonbondedMethod not in (NOCUTOFF, CUTOFF_NONPERIODIC, CUTOFF_PERIODIC, EWALD, PME):
            raise ValueError("Illegal value for nonbondedMethod")
        if OPLS:
            useDispersionCorrection = False        # nb.setUseDispersionCorrection(False), cnb.setUseLongRangeCorrection(False)
        gb = None
        if implicitSolvent is not None:
            if implicitSolvent in ("AGBNP", "GVolSA", "AGBNP3"):
                raise NotImplementedError("%s is not supported in this version" % implicitSolvent)
            if implicitSolvent != HCT:
                raise ValueError("Illegal implicit solvent method")
            reactionFieldDielectric = 1.0          # nb.setReactionFieldDielectric(1.0), :451
            gb_parms = self._get_gb_params()
            if not gb_parms:
                raise IOError("No HCT parameters found in DMS file")
            gb = GBSAHCTForce(SA="ACE")
            for p in gb_parms:
                gb.addParticle(list(p))
            gb.finalize()
            gb.setForceGroup(self._nonbonded_force_group)
        charge, sigma, epsilon, excl, exc_pairs, exc_params = [], [], [], [], [], []
        for conn, tables, off in zip(self._conn, self._tables, self._offset[:-1]):
            q = """SELECT charge, sigma, epsilon FROM particle INNER JOIN nonbonded_param
                   ON particle.nbtype=nonbonded_param.id ORDER BY particle.id"""
            for c, s, e in conn.execute(q):
                charge.append(c)
                sigma.append(s * ANGSTROM)
                epsilon.append(e * KCAL)
            if "exclusion" in tables:
                for p0, p1 in conn.execute("SELECT p0, p1 FROM exclusion"):
                    excl.append((p0 + off, p1 + off))
            if "pair_12_6_es_term" in tables:
                q = """SELECT p0, p1, aij, bij, qij FROM pair_12_6_es_term INNER JOIN pair_12_6_es_param
                       ON pair_12_6_es_term.param=pair_12_6_es_param.id"""
                for p0, p1, a_ij, b_ij, q_ij in conn.execute(q):
                    a = a_ij * KCAL * ANGSTROM ** 12
                    b = b_ij * KCAL * ANGSTROM ** 6
                    if a == 0.0 or b == 0.0:
                        eps, sig = 0.0, 1.0
                    else:
                        eps, sig = b * b / (4 * a), (a / b) ** (1.0 / 6.0)
                    exc_pairs.append((p0 + off, p1 + off))
                    exc_params.append((q_ij, sig, eps))
        if len(charge) != self.getNumAtoms():
            raise ValueError("every particle needs a nonbonded_param row")
        es = {(min(a, b), max(a, b)) for a, b in excl}
        for a, b in exc_pairs:
            if (min(a, b), max(a, b)) not in es:
                raise ValueError("1-4 pair (%d, %d) is not in the exclusion table" % (a, b))
        box = self.getBox()
        if nonbondedMethod in (CUTOFF_PERIODIC, EWALD, PME) and not np.all(box > 0):
            raise ValueError("a periodic cutoff needs a global_cell table")
        return NonbondedSystem(np.array(charge), np.array(sigma), np.array(epsilon),
                               np.array(excl, dtype=np.int32).reshape(-1, 2),
                               np.array(exc_pairs, dtype=np.int32).reshape(-1, 2),
                               np.array(exc_params, dtype=np.float64).reshape(-1, 3),
                               method=int(nonbondedMethod), cutoff=float(nonbondedCutoff),
                               eps_rf=float(reactionFieldDielectric), box=box,
                               use_dispersion_correction=bool(useDispersionCorrection),
                               ewald_tolerance=float(ewaldErrorTolerance), gb=gb, lj_geometric=bool(OPLS))

    # ---- write-back ---------------------------------------------------------------------------
    def _write_vec3(self, columns, values, scale):
        values = np.asarray(values, dtype=np.float64).reshape(-1, 3)
        if values.shape[0] != self.getNumAtoms():
            raise ValueError("expected %d rows" % self.getNumAtoms())
        q = "UPDATE particle SET %s = ?, %s = ?, %s = ? WHERE id == ?" % columns
        iat = 0
        for conn, natoms in zip(self._conn, self._natoms):
            rows = values[iat:iat + natoms] / scale
            conn.executemany(q, [(float(r[0]), float(r[1]), float(r[2]), i) for i, r in enumerate(rows)])
            conn.commit()
            iat += natoms
        return iat

    def setPositions(self, positions):
        """Update atomic positions (nm) in the attached DMS files; returns the atom count."""
        n = self._write_vec3(("x", "y", "z"), positions, ANGSTROM)
        self.positions = np.asarray(positions, dtype=np.float64).reshape(-1, 3).copy()
        return n

    def setVelocities(self, velocities):
        """Update atomic velocities (nm/ps) in the attached DMS files; returns the atom count."""
        n = self._write_vec3(("vx", "vy", "vz"), velocities, ANGSTROM)
        self.velocities = np.asarray(velocities, dtype=np.float64).reshape(-1, 3).copy()
        return n

    def setGlobalCell(self, a, b, c):
        """Update the box vectors (nm, each a 3-vector) in the attached DMS files."""
        q = "UPDATE global_cell SET x = ?, y = ?, z = ? WHERE id == ?"
        for conn, tables in zip(self._conn, self._tables):
            if "global_cell" in tables:
                for row, v in enumerate((a, b, c), start=1):
                    v = np.asarray(v, dtype=np.float64) / ANGSTROM
                    conn.execute(q, (float(v[0]), float(v[1]), float(v[2]), row))
                conn.commit()

    def close(self):
        for conn in self._conn:
            conn.close()
        self._conn = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
