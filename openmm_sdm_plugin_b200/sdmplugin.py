"""Host-side mirror of the reference's Python surface for the dual-state force path.

`LangevinIntegratorSDM` keeps the names, argument meaning, defaults, units (kJ/mol, nm, ps, K)
and error behaviour of `SDMPlugin::LangevinIntegratorSDM` as SWIG exposes it
(python/SDMplugin.i:86-145; ctor defaults openmmapi/src/LangevinIntegratorSDM.cpp:48-85), and
`SDMUtils` keeps the method constants of python/SDMUtils.py:9-15.  What is NOT mirrored is the
integrator proper: `step()` in the reference also integrates the Langevin equations
(not part of the hot path); here `bind()` + `evaluate()` perform
exactly the force column of `step()` -- LangevinIntegratorSDM.cpp:156-182 up to and including
the hybrid force of ReferenceSDMKernels.cpp:309-318 -- on the GPU through the C ABI.

All arithmetic happens in libsdmb200.so; this module only marshals state.
"""
from __future__ import annotations

import os

import copy

import numpy as np

from . import _lib
from .system import AlchemicalState, NonbondedSystem, apply_implicit_solvent


class OpenMMException(Exception):
    """Stands in for OpenMM::OpenMMException (the reference throws it from C++)."""


def _md_units(x):
    """A plain number (or list of numbers) in OpenMM's unit system -- kJ/mol, nm, radians.  The reference's
    scripts pass simtk.unit quantities; those are converted when simtk.unit is importable."""
    if hasattr(x, "value_in_unit_system"):
        from simtk.unit import md_unit_system   # only reached with simtk quantities in hand
        x = x.value_in_unit_system(md_unit_system)
    if isinstance(x, (list, tuple)):
        return [float(v) for v in x]
    try:
        return float(x)
    except TypeError:
        return [float(v) for v in x]


class SDMUtils(object):
    """python/SDMUtils.py: the method constants (:9-15) and the two restraint builders.  The reference adds
    OpenMM Custom*Forces (force group 1) to the System; here the same terms are recorded on the system
    description (``system.sdm_restraints``) and evaluated on the device by every context bound to it
    (LangevinIntegratorSDM.bind -> sdm_add_centroid_restraint / sdm_add_alignment_restraint).  Numbers are in
    kJ/mol, nm and radians (what the reference divides its quantities down to, SDMUtils.py:130-149)."""

    def __init__(self, system=None):
        self.system = system
        self.RestraintControlParameterName = "SDMRestraintControlParameter"
        self.force = None
        self.LinearMethod = 0
        self.QuadraticMethod = 1
        self.ILogisticMethod = 2
        self.NoSoftCoreMethod = 0
        self.TanhSoftCoreMethod = 1
        self.RationalSoftCoreMethod = 2

    def getControlParameterName(self):
        return self.RestraintControlParameterName

    def _restraints(self):
        if self.system is None:
            raise OpenMMException("SDMUtils was created without a system")
        if not hasattr(self.system, "sdm_restraints"):
            self.system.sdm_restraints = []
        return self.system.sdm_restraints

    def addRestraintForce(self, lig_cm_particles=None, rcpt_cm_particles=None, kfcm=0.0, tolcm=0.0,
                          lig_ref_particles=None, rcpt_ref_particles=None, angle_center=0.0, kfangle=0.0,
                          angletol=10.0 * 3.14159265358979323846 / 180.0, dihedral1center=0.0, kfdihedral1=0.0,
                          dihedral1tol=10.0 * 3.14159265358979323846 / 180.0, dihedral2center=0.0, kfdihedral2=0.0,
                          dihedral2tol=10.0 * 3.14159265358979323846 / 180.0, offset=(0.0, 0.0, 0.0),
                          lig_cm_weights=None, rcpt_cm_weights=None):
        """SDMUtils.py:32-162.  Centroid weights: OpenMM's CustomCentroidBondForce weighs by mass; pass the
        masses of the group atoms as lig_cm_weights / rcpt_cm_weights (None = equal weights)."""
        if not (lig_cm_particles and rcpt_cm_particles):
            return                                             # :49-53
        spec = {"kind": "centroid", "lig_cm_atoms": [int(i) for i in lig_cm_particles],
                "rcpt_cm_atoms": [int(i) for i in rcpt_cm_particles],
                "lig_cm_weights": lig_cm_weights, "rcpt_cm_weights": rcpt_cm_weights,
                "kfcm": _md_units(kfcm), "tolcm": _md_units(tolcm), "offset": _md_units(offset)}
        if lig_ref_particles and rcpt_ref_particles:
            if not (len(lig_ref_particles) == 3 and len(rcpt_ref_particles) == 3):
                raise ValueError("Invalid lists of reference atoms")      # :57-58
            c = [_md_units(angle_center), _md_units(dihedral1center), _md_units(dihedral2center)]
            t = [_md_units(angletol), _md_units(dihedral1tol), _md_units(dihedral2tol)]
            spec.update(lig_ref=[int(i) for i in lig_ref_particles], rcpt_ref=[int(i) for i in rcpt_ref_particles],
                        kfcd=[_md_units(kfangle), _md_units(kfdihedral1), _md_units(kfdihedral2)],
                        a=[c[k] - t[k] for k in range(3)], b=[c[k] + t[k] for k in range(3)])   # :137-149
        self._restraints().append(spec)
        self.force = spec

    def addAlignmentForce(self, liga_ref_particles=None, ligb_ref_particles=None, kfdispl=0.0, ktheta=0.0, kpsi=0.0,
                          offset=(0.0, 0.0, 0.0)):
        """SDMUtils.py:166-258."""
        if not (liga_ref_particles and ligb_ref_particles) or \
                not (len(liga_ref_particles) == 3 and len(ligb_ref_particles) == 3):
            raise ValueError("Invalid lists of reference atoms")          # :173-175
        self._restraints().append({"kind": "alignment", "liga_ref": [int(i) for i in liga_ref_particles],
                                   "ligb_ref": [int(i) for i in ligb_ref_particles], "kfdispl": _md_units(kfdispl),
                                   "ktheta": _md_units(ktheta), "kpsi": _md_units(kpsi), "offset": _md_units(offset)})


def apply_restraints(ctx, system):
    """Hand the restraint terms recorded on `system` by SDMUtils to a context."""
    for s in getattr(system, "sdm_restraints", []):
        if s["kind"] == "centroid":
            ctx.add_centroid_restraint(s["lig_cm_atoms"], s["rcpt_cm_atoms"], s["kfcm"], s["tolcm"], s["offset"],
                                       s.get("lig_cm_weights"), s.get("rcpt_cm_weights"), s.get("lig_ref"),
                                       s.get("rcpt_ref"), s.get("kfcd", (0, 0, 0)), s.get("a", (0, 0, 0)),
                                       s.get("b", (0, 0, 0)))
        else:
            ctx.add_alignment_restraint(s["liga_ref"], s["ligb_ref"], s["kfdispl"], s["ktheta"], s["kpsi"], s["offset"])


class LangevinIntegratorSDM(object):
    # LangevinIntegratorSDM.h:120-122 and :143-145
    LinearMethod, QuadraticMethod, ILogisticMethod = 0, 1, 2
    NoSoftCoreMethod, TanhMethod, RationalMethod = 0, 1, 2

    def __init__(self, temperature, frictionCoeff, stepSize, nParticles):
        self._temperature = float(temperature)
        self._friction = float(frictionCoeff)
        self._step_size = float(stepSize)
        self._seed = int.from_bytes(os.urandom(4), "little") & 0x7fffffff   # osrngseed()
        self._a = AlchemicalState()          # ctor defaults: umax 200, a 1/4, ub 0, lambda 1, ...
        self._a.step_size = self._step_size
        self._bind_e = 0.0
        self._pot_energy = 0.0
        self._n = int(nParticles)
        self._displ = np.zeros((self._n, 3), np.float64)   # displ.push_back(Vec3(0,0,0)) x nParticles
        self._ctx = None
        self._displ_dirty = False

    # ---- plain state (names as in SDMplugin.i) ------------------------------------------------
    def setLambda(self, lambdac): self._a.lambdac = float(lambdac)
    def getLambda(self): return self._a.lambdac
    def getTemperature(self): return self._temperature
    def setTemperature(self, temp): self._temperature = float(temp)
    def getFriction(self): return self._friction
    def setFriction(self, coeff): self._friction = float(coeff)
    def getRandomNumberSeed(self): return self._seed
    def setRandomNumberSeed(self, seed): self._seed = int(seed)
    def getStepSize(self): return self._step_size

    def setStepSize(self, size):
        self._step_size = float(size)
        self._a.step_size = self._step_size

    def getBindE(self): return self._bind_e
    def setBindE(self, be): self._bind_e = float(be)
    def getPotEnergy(self): return self._pot_energy
    def setPotEnergy(self, e): self._pot_energy = float(e)
    def getUmax(self): return self._a.umax
    def setUmax(self, um): self._a.umax = float(um)
    def getAcore(self): return self._a.acore
    def setAcore(self, a): self._a.acore = float(a)
    def getUbcore(self): return self._a.ubcore
    def setUbcore(self, a): self._a.ubcore = float(a)
    def setBiasMethod(self, method): self._a.bias_method = int(method)
    def getBiasMethod(self): return self._a.bias_method
    def setSoftCoreMethod(self, method): self._a.softcore_method = int(method)
    def getSoftCoreMethod(self): return self._a.softcore_method
    def setGamma(self, gammat): self._a.gammac = float(gammat)
    def getGamma(self): return self._a.gammac
    def setWBcoeff(self, wbcoeff_t): self._a.wbcoeff = float(wbcoeff_t)
    def getWBcoeff(self): return self._a.wbcoeff
    def setW0coeff(self, w0coeff_t): self._a.w0coeff = float(w0coeff_t)
    def getW0coeff(self): return self._a.w0coeff
    def setLambda1(self, lambda1_t): self._a.lambda1 = float(lambda1_t)
    def getLambda1(self): return self._a.lambda1
    def setLambda2(self, lambda2_t): self._a.lambda2 = float(lambda2_t)
    def getLambda2(self): return self._a.lambda2
    def setAlpha(self, alpha_t): self._a.alpha = float(alpha_t)
    def getAlpha(self): return self._a.alpha
    def setU0(self, u0_t): self._a.u0 = float(u0_t)
    def getU0(self): return self._a.u0
    def setNoneqtmax(self, noneq_tmax): self._a.noneq_tmax = float(noneq_tmax)
    def getNoneqtmax(self): return self._a.noneq_tmax
    def getNonEquilibrium(self): return self._a.nonequilibrium
    # setNonEquilibrium exists in C++ (LangevinIntegratorSDM.h) but SDMplugin.i:124 does not wrap
    # it; it is provided so that the non-equilibrium schedule can be exercised at all.
    def setNonEquilibrium(self, flag): self._a.nonequilibrium = int(flag)
    def setNoneqWorkvalue(self, noneq_work): self._a.work_value = float(noneq_work)
    def getNoneqWorkvalue(self): return self._a.work_value
    def setlambda1Slope(self, ml1): self._a.m_lambda1 = float(ml1)
    def getlambda1Slope(self): return self._a.m_lambda1
    def setlambda2Slope(self, ml2): self._a.m_lambda2 = float(ml2)
    def getlambda2Slope(self): return self._a.m_lambda2
    def setu0Slope(self, mu0): self._a.m_u0 = float(mu0)
    def getu0Slope(self): return self._a.m_u0
    def setw0Slope(self, mw0): self._a.m_w0 = float(mw0)
    def getw0Slope(self): return self._a.m_w0
    def setlambda1intercept(self, bl1): self._a.b_lambda1 = float(bl1)
    def getlambda1intercept(self): return self._a.b_lambda1
    def setlambda2intercept(self, bl2): self._a.b_lambda2 = float(bl2)
    def getlambda2intercept(self): return self._a.b_lambda2
    def setu0intercept(self, bu0): self._a.b_u0 = float(bu0)
    def getu0intercept(self): return self._a.b_u0
    def setw0intercept(self, bw0): self._a.b_w0 = float(bw0)
    def getw0intercept(self): return self._a.b_w0

    # ---- displacement map (LangevinIntegratorSDM.h:467-472) -----------------------------------
    def setDisplacement(self, atom, dx, dy, dz):
        # the reference indexes a std::vector without a bounds check; Python raises instead of
        # corrupting memory
        if not 0 <= int(atom) < self._n:
            raise IndexError("particle index %d out of range (nParticles = %d)" % (atom, self._n))
        self._displ[int(atom)] = (float(dx), float(dy), float(dz))
        self._displ_dirty = True

    def getDisplacement(self, atom):
        if not 0 <= int(atom) < self._n:
            raise IndexError("particle index %d out of range (nParticles = %d)" % (atom, self._n))
        return tuple(float(v) for v in self._displ[int(atom)])

    # ---- the force path ------------------------------------------------------------------------
    def bind(self, system: NonbondedSystem, device: int = -1, **ctx_options):
        """What LangevinIntegratorSDM::initialize does for this path
        (LangevinIntegratorSDM.cpp:89-106): bind to one context, snapshot the displacement map
        (ReferenceSDMKernels.cpp:150-154)."""
        from .context import SDMContext
        if self._ctx is not None:
            raise OpenMMException("This Integrator is already bound to a context")
        if system.n_atoms != self._n:
            raise OpenMMException("nParticles (%d) does not match the system (%d)" % (self._n, system.n_atoms))
        self._ctx = SDMContext(system, self._displ, n_replicas=1, device=device, **ctx_options)
        apply_restraints(self._ctx, system)     # what SDMUtils recorded on the system (force group 1 in the reference)
        if int(system.method) in (3, 4):        # NonbondedForce::Ewald / ::PME: the reference gets the complete sum
            self._ctx.enable_reciprocal_pme()   # from OpenMM; here direct space + reciprocal space on the device
        apply_implicit_solvent(self._ctx, system)   # GBSAHCTForce in the nonbonded force group: both states
        self._displ_dirty = False
        return self

    def cleanup(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    def evaluate(self, positions, bonded_forces=None, restraint_energy=0.0):
        """The force column of one `step()`: both states, soft-core, bias, hybrid force.
        Returns the hybrid force [nParticles, 3] (kJ/mol/nm) and updates getBindE()/getPotEnergy()
        (and, in non-equilibrium mode, lambda1/lambda2/u0/w0/work like
        ReferenceSDMKernels.cpp:231-245,289-302)."""
        if self._ctx is None:
            raise OpenMMException("the integrator is not bound to a context: call bind(system) first")
        # the reference snapshots the map at initialize(); a later setDisplacement() needs
        # Context::reinitialize there.  Here the new map is uploaded.
        if self._displ_dirty:
            self._ctx.set_displacement(self._displ)
            self._displ_dirty = False
        c = self._ctx
        c.set_positions(0, np.ascontiguousarray(positions, np.float64))
        c.set_bonded_forces(0, bonded_forces, float(restraint_energy))
        c.set_alchemical(0, self._a)
        alch0 = copy.copy(self._a)
        for attempt in range(4):
            c.eval()
            sc = c.scalars(0)
            # SDM_ERR_STALE_LIST / SDM_ERR_CAPACITY heal themselves (include/sdmb200.h): reading the
            # scalars made the library rebuild its list / grow its scratch, the evaluation is repeated
            # from the same alchemical state.  The reference's force path never fails this way.
            if sc["status"] not in (_lib.SDM_ERR_STALE_LIST, _lib.SDM_ERR_CAPACITY):
                break
            c.set_alchemical(0, alch0)
        if sc["status"] == _lib.SDM_ERR_SOFTCORE:
            raise OpenMMException("Unknown soft core method")     # LangevinIntegratorSDM.cpp:147
        if sc["status"] != 0:
            raise OpenMMException("libsdmb200 status %d" % sc["status"])
        self._bind_e = sc["bind_e"]
        self._pot_energy = sc["pot_energy"]
        c.get_alchemical(0, self._a)
        self.last_scalars = sc
        return c.forces(0, _lib.FORCE_HYBRID)

    # ---- dynamics on the device (no constraints) ----------------------------
    def setState(self, positions, velocities=None, masses=None):
        """What the reference's Context holds for the integrator: positions (nm), velocities (nm/ps,
        zero if omitted) and the particle masses (amu; needed once).  The state then lives on the
        device; `step()` advances it there."""
        if self._ctx is None:
            raise OpenMMException("the integrator is not bound to a context: call bind(system) first")
        if masses is not None:
            self._masses = np.ascontiguousarray(masses, np.float64).copy()
            self._md_params = None
        self._ctx.set_positions(0, np.ascontiguousarray(positions, np.float64))
        self._pending_vel = (np.zeros((self._n, 3)) if velocities is None
                             else np.ascontiguousarray(velocities, np.float64).copy())

    def getPositions(self):
        return self._ctx.positions(0)

    def getVelocities(self):
        return self._ctx.md_velocities(0)

    def computeKineticEnergy(self):
        """LangevinIntegratorSDM::computeKineticEnergy (ReferenceSDMKernels.cpp:105-137, no constraints)."""
        return self._ctx.md_kinetic_energy(0)

    def step(self, steps):
        """LangevinIntegratorSDM::step (LangevinIntegratorSDM.cpp:153-183) with the state on the
        device: per step the fused dual-state evaluation and the reference's Langevin update
        (ReferenceStochasticDynamicsSDM.cpp:131-266, FP64, no constraints).  Force group 1 is
        whatever was last handed over with the context's set_bonded_forces (zero by default)."""
        if self._ctx is None:
            raise OpenMMException("the integrator is not bound to a context: call bind(system) first")
        if getattr(self, "_masses", None) is None:
            raise OpenMMException("step() needs the particle masses: call setState(positions, velocities, masses) "
                                  "first, or evaluate(positions) for the force column of a step alone")
        c = self._ctx
        params = (self._temperature, self._friction, self._step_size)
        if getattr(self, "_md_params", None) != params:
            # like the Reference kernel, the dynamics object is recreated when T, friction or dt change
            # (ReferenceSDMKernels.cpp:320-337); velocities survive
            keep = None if getattr(self, "_md_params", None) is None else c.md_velocities(0)
            c.md_init(self._masses, self._temperature, self._friction, self._step_size, self._seed & (2 ** 63 - 1))
            if keep is not None:
                c.md_set_velocities(0, keep)
            self._md_params = params
        if getattr(self, "_pending_vel", None) is not None:
            c.md_set_velocities(0, self._pending_vel)
            self._pending_vel = None
        if self._displ_dirty:
            c.set_displacement(self._displ)
            self._displ_dirty = False
        for _ in range(int(steps)):
            c.set_alchemical(0, self._a)       # non-equilibrium schedules advance with the integrator object
            c.md_step(1)
            sc = c.scalars(0)
            if sc["status"] == _lib.SDM_ERR_SOFTCORE:
                raise OpenMMException("Unknown soft core method")     # LangevinIntegratorSDM.cpp:147
            if sc["status"] not in (0, _lib.SDM_ERR_STALE_LIST):
                raise OpenMMException("libsdmb200 status %d" % sc["status"])
            if sc["status"] == _lib.SDM_ERR_STALE_LIST:
                c.invalidate_list()            # an atom left the list's skin: rebuild before the next step
            c.get_alchemical(0, self._a)
            self._bind_e = sc["bind_e"]
            self._pot_energy = sc["pot_energy"]
            self.last_scalars = sc
