"""Flat description of the force-group-2 system the SDM hot path works on.

This is the host-side data format on the input side of the path: what the reference's
DMS reader hands to OpenMM's NonbondedForce (example/desmonddmsfile75.py:772-850) plus
the integrator's displacement map (openmmapi/include/LangevinIntegratorSDM.h:467-472,508)
and alchemical settings (example/test.py:24-42,169-191; example/test_explicit.py:21-43,
164-188).  Units are OpenMM's (nm, kJ/mol, e, ps).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

NOCUTOFF = 0
CUTOFF_NONPERIODIC = 1
CUTOFF_PERIODIC = 2
EWALD = 3             # NonbondedForce::Ewald / ::PME: the direct-space part (include/sdmb200.h)
PME = 4

# LangevinIntegratorSDM.h:120-122,143-145 / SDMUtils.py:9-15
LINEAR, QUADRATIC, ILOGISTIC = 0, 1, 2
NO_SOFTCORE, TANH_SOFTCORE, RATIONAL_SOFTCORE = 0, 1, 2

KCAL = 4.184
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")


HCT = "HCT"     # stands in for simtk.openmm.app.amberprmtopfile.HCT (desmonddmsfile75.py:40)


class GBSAHCTForce:
    """Host-side mirror of OpenMM 7.2/7.3 app/internal/customgbforces.py GBSAHCTForce as the reference's reader
    uses it (example/desmonddmsfile75.py:460-465): addParticle([charge, radius, scale]) records, finalize()
    turns every record into the CustomGBForce per-particle parameters charge, or = radius - OFFSET,
    sr = scale * or (CustomAmberGBForceBase._addParticles).  The energy expressions are evaluated on the device
    by csrc/kernels_gb.cu (sdm_enable_hct_gb).  Only the options the reference uses exist: no cutoff, kappa = 0,
    SA None or 'ACE'.  OpenMM is not available in this image: restated, parity unpinned."""
    OFFSET = 0.009

    def __init__(self, solventDielectric=78.5, soluteDielectric=1, SA=None, cutoff=None, kappa=0.0):
        if cutoff is not None or kappa != 0.0:
            raise ValueError("only cutoff=None, kappa=0 (what desmonddmsfile75.py:460 asks for) is built")
        if SA not in (None, "ACE"):
            raise ValueError("Unknown surface area method: " + str(SA))
        self.solventDielectric, self.soluteDielectric, self.SA = float(solventDielectric), float(soluteDielectric), SA
        self.parameters = []
        self._particles = None
        self.force_group = 0

    def addParticle(self, params):
        self.parameters.append(list(params))

    def setParticleParameters(self, idx, params):
        self.parameters[idx] = list(params)

    def finalize(self):
        out = []
        for charge, radius, scale in self.parameters:
            offset_radius = float(radius) - self.OFFSET
            out.append((float(charge), offset_radius, float(scale) * offset_radius))
        self._particles = np.array(out, dtype=np.float64).reshape(-1, 3)

    def setForceGroup(self, group):
        self.force_group = int(group)

    def getNumParticles(self):
        return len(self.parameters)

    def getParticleParameters(self, idx):
        """(charge, or, sr) of a finalized force."""
        if self._particles is None:
            raise ValueError("finalize() has not been called")
        return tuple(self._particles[idx])

    def device_parameters(self):
        """charge, or, sr arrays for sdm_enable_hct_gb."""
        if self._particles is None:
            raise ValueError("GBSAHCTForce.finalize() has not been called")
        return self._particles[:, 0].copy(), self._particles[:, 1].copy(), self._particles[:, 2].copy()


def apply_implicit_solvent(ctx, system):
    """Switch the system's GB force (if any) on in a context: both states, every replica."""
    gb = getattr(system, "gb", None)
    if gb is None:
        return
    if gb.getNumParticles() != system.n_atoms:
        raise ValueError("GBSAHCTForce has %d particles, the system %d" % (gb.getNumParticles(), system.n_atoms))
    q, o, sr = gb.device_parameters()
    ctx.enable_hct_gb(o, sr, charge=q, solute_dielectric=gb.soluteDielectric,
                      solvent_dielectric=gb.solventDielectric, sa_ace=gb.SA == "ACE")


@dataclass
class NonbondedSystem:
    charge: np.ndarray            # [n] e
    sigma: np.ndarray             # [n] nm
    epsilon: np.ndarray           # [n] kJ/mol
    exclusions: np.ndarray        # [ne,2] int32  (every addException pair)
    exception_pairs: np.ndarray   # [nx,2] int32  (1-4 pairs with non-zero parameters)
    exception_params: np.ndarray  # [nx,3] chargeProd, sigma, epsilon
    method: int = CUTOFF_PERIODIC
    cutoff: float = 1.0
    eps_rf: float = 78.3          # NonbondedForce default reaction-field dielectric
    box: np.ndarray = field(default_factory=lambda: np.zeros(3))
    use_dispersion_correction: bool = True
    ewald_alpha: float = 0.0      # EWALD / PME: 0 = OpenMM's rule sqrt(-log(2 tol)) / cutoff
    ewald_tolerance: float = 5e-4
    gb: object = None             # GBSAHCTForce (finalized) or None: implicit solvent in the nonbonded force group
    lj_geometric: bool = False    # createSystem(OPLS=True): Lennard-Jones through the CustomNonbondedForce with
                                  # sigma12 = sqrt(s1 s2), eps12 = sqrt(e1 e2) (desmonddmsfile75.py:780-810)

    def ewald_alpha_effective(self) -> float:
        """The splitting parameter the library and the oracle use (NonbondedForceImpl::calcPMEParameters)."""
        if self.method not in (EWALD, PME):
            return 0.0
        tol = self.ewald_tolerance if self.ewald_tolerance > 0 else 5e-4
        return float(self.ewald_alpha) if self.ewald_alpha > 0 else float(np.sqrt(-np.log(2.0 * tol)) / self.cutoff)

    def __post_init__(self):
        self.charge = np.ascontiguousarray(self.charge, dtype=np.float64)
        self.sigma = np.ascontiguousarray(self.sigma, dtype=np.float64)
        self.epsilon = np.ascontiguousarray(self.epsilon, dtype=np.float64)
        self.exclusions = np.ascontiguousarray(self.exclusions, dtype=np.int32).reshape(-1, 2)
        self.exception_pairs = np.ascontiguousarray(self.exception_pairs, dtype=np.int32).reshape(-1, 2)
        self.exception_params = np.ascontiguousarray(self.exception_params, dtype=np.float64).reshape(-1, 3)
        self.box = np.ascontiguousarray(self.box, dtype=np.float64)

    @property
    def n_atoms(self) -> int:
        return int(self.charge.shape[0])

    def addForce(self, force):
        """sys.addForce(gb) of desmonddmsfile75.py:464 -- the one extra force of the nonbonded group."""
        if not isinstance(force, GBSAHCTForce):
            raise TypeError("only a GBSAHCTForce can be added to the nonbonded force group")
        self.gb = force
        return 0


@dataclass
class AlchemicalState:
    """Scalar state of LangevinIntegratorSDM (defaults = ctor, LangevinIntegratorSDM.cpp:48-85)."""
    bias_method: int = LINEAR
    softcore_method: int = NO_SOFTCORE
    lambdac: float = 1.0
    gammac: float = 0.0
    wbcoeff: float = 1.0
    w0coeff: float = 0.0
    lambda1: float = 1.0
    lambda2: float = 1.0
    alpha: float = 1.0
    u0: float = 0.0
    umax: float = 200.0
    acore: float = 0.25
    ubcore: float = 0.0
    nonequilibrium: int = 0
    noneq_tmax: float = 1.0
    work_value: float = 0.0
    time: float = 0.0
    step_size: float = 0.001
    m_lambda1: float = 0.0
    m_lambda2: float = 0.0
    m_u0: float = 0.0
    m_w0: float = 0.0
    b_lambda1: float = 0.0
    b_lambda2: float = 0.0
    b_u0: float = 0.0
    b_w0: float = 0.0


@dataclass
class SDMCase:
    name: str
    system: NonbondedSystem
    positions: np.ndarray         # [n,3] nm
    displacement: np.ndarray      # [n,3] nm, the displacement map
    alch: AlchemicalState
    velocities: np.ndarray | None = None
    masses: np.ndarray | None = None
    constraint_pairs: np.ndarray | None = None   # [nc,2] System.addConstraint pairs (desmonddmsfile75.py:560-637)
    constraint_dist: np.ndarray | None = None    # [nc] nm


def _load(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


def _displacement_map(resid, lig1_resid, lig2_resid, d):
    disp = np.zeros((len(resid), 3))
    disp[resid == lig1_resid] = d
    disp[resid == lig2_resid] = -np.asarray(d)
    return disp


def cfg1() -> SDMCase:
    """example/test.py: OA-G6/G3, 230 atoms, CutoffNonPeriodic 15 nm (test.py:63), lig1 =
    resid 3 displaced +d, lig2 = resid 2 displaced -d (test.py:31-38,181-185), ILogistic
    lambda1=lambda2=0.025, alpha=0, rational soft-core umax=100 kcal, ub=50 kcal, a=1/16
    (test.py:24-42,172-191)."""
    z = _load("cfg1_oa_g6_g3.npz")
    sysd = NonbondedSystem(z["charge"], z["sigma"], z["epsilon"], z["exclusions"],
                           z["exception_pairs"], z["exception_params"],
                           method=CUTOFF_NONPERIODIC, cutoff=15.0, eps_rf=78.3,
                           box=np.zeros(3), use_dispersion_correction=True)
    d = np.array([-15.559, -3.000, 8.600]) * 0.1
    disp = _displacement_map(z["resid"], 3, 2, d)
    al = AlchemicalState(bias_method=ILOGISTIC, softcore_method=RATIONAL_SOFTCORE,
                         lambdac=0.025, lambda1=0.025, lambda2=0.025, alpha=0.0, u0=0.0,
                         w0coeff=0.0, umax=100.0 * KCAL, ubcore=50.0 * KCAL, acore=0.0625)
    return SDMCase("cfg1_oa_g6_g3", sysd, z["positions"].copy(), disp, al,
                   z["velocities"].copy(), z["masses"].copy(),
                   z["constraint_pairs"].copy(), z["constraint_dist"].copy())


def cfg2() -> SDMCase:
    """example/test_explicit.py: TEMOA-G1/G4, 20 446 atoms, explicit water, 1 nm cutoff; the
    in-scope electrostatics is CutoffPeriodic / reaction field (mismatch 3;
    the shipped script says PME at test_explicit.py:64).  lig1 = resid 2 +d, lig2 = resid 3
    -d, d = 2.2 nm each axis (test_explicit.py:28-31,178-182); lambda=lambda1=lambda2=0.5."""
    z = _load("cfg2_temoa_g1_g4.npz")
    sysd = NonbondedSystem(z["charge"], z["sigma"], z["epsilon"], z["exclusions"],
                           z["exception_pairs"], z["exception_params"],
                           method=CUTOFF_PERIODIC, cutoff=1.0, eps_rf=78.3,
                           box=z["box"].copy(), use_dispersion_correction=True)
    d = np.array([22.0, 22.0, 22.0]) * 0.1
    disp = _displacement_map(z["resid"], 2, 3, d)
    al = AlchemicalState(bias_method=ILOGISTIC, softcore_method=RATIONAL_SOFTCORE,
                         lambdac=0.5, lambda1=0.5, lambda2=0.5, alpha=0.0, u0=0.0,
                         w0coeff=0.0, umax=100.0 * KCAL, ubcore=50.0 * KCAL, acore=0.0625)
    return SDMCase("cfg2_temoa_g1_g4_rf", sysd, z["positions"].copy(), disp, al,
                   z["velocities"].copy(), z["masses"].copy(),
                   z["constraint_pairs"].copy(), z["constraint_dist"].copy())


def atm_lambda_schedule(n_windows: int = 22):
    """A lambda ladder for cfg3 (the reference ships none): ILogistic,
    first half lambda1 = 0 with lambda2 ramping 0 -> 0.5, second half lambda2 = 0.5 with
    lambda1 ramping 0 -> 0.5 (the ATM leg ends at the alchemical midpoint lambda1 = lambda2 =
    1/2), alpha = 0.1 (kcal/mol)^-1, u0 = 110 kcal/mol, w0 = 0, rational soft-core as in the
    shipped scripts (example/test.py:40-42)."""
    states = []
    half = n_windows // 2
    for k in range(n_windows):
        if k < half:
            l1, l2 = 0.0, 0.5 * k / max(half - 1, 1)
        else:
            l1, l2 = 0.5 * (k - half) / max(n_windows - half - 1, 1), 0.5
        states.append(AlchemicalState(bias_method=ILOGISTIC, softcore_method=RATIONAL_SOFTCORE,
                                      lambdac=l1 + l2,
                                      lambda1=l1, lambda2=l2,
                                      alpha=0.1 / KCAL, u0=110.0 * KCAL, w0coeff=0.0,
                                      umax=100.0 * KCAL, ubcore=50.0 * KCAL, acore=0.0625))
    return states


def synthetic_case(n_atoms: int = 50_000, ligand_atoms: int = 60, seed: int = 1234,
                   cutoff: float = 1.0, density: float = 98.7,
                   displacement=(0.0, 0.0, 3.0), protein_atoms: int = 3000) -> SDMCase:
    """Synthetic explicit-solvent ATM/ABFE box (configs 4 and 5).

    3-site rigid-water-like molecules (q = -0.834 / +0.417, sigma_O = 0.315 nm,
    eps_O = 0.636 kJ/mol, H: eps = 0) on a jittered cubic lattice at `density` atoms/nm^3,
    intramolecular exclusions; a blob of `protein_atoms` LJ + charged atoms around the box
    centre with a bonded-neighbour exclusion graph and 1-4-like exceptions; one ligand of
    `ligand_atoms` atoms at the centre displaced by `displacement` (ABFE: one group).
    Deterministic in `seed`.
    """
    rng = np.random.default_rng(seed)
    nmol = int(round(n_atoms / 3))
    n = 3 * nmol
    L = (n / density) ** (1.0 / 3.0)
    m = int(np.ceil(nmol ** (1.0 / 3.0)))
    a = L / m
    # lattice sites, shuffled deterministically so that vacancies are spread evenly
    idx = rng.permutation(m ** 3)[:nmol]
    idx.sort()
    gx, gy, gz = idx % m, (idx // m) % m, idx // (m * m)
    site = (np.stack([gx, gy, gz], axis=1) + 0.5) * a
    o = site + rng.uniform(-0.03, 0.03, size=(nmol, 3))
    # random orthonormal frames
    v1 = rng.normal(size=(nmol, 3))
    v1 /= np.linalg.norm(v1, axis=1, keepdims=True)
    v2 = rng.normal(size=(nmol, 3))
    v2 -= (v2 * v1).sum(1, keepdims=True) * v1
    v2 /= np.linalg.norm(v2, axis=1, keepdims=True)
    roh, ang = 0.09572, np.deg2rad(104.52) / 2
    h1 = o + roh * (np.cos(ang) * v1 + np.sin(ang) * v2)
    h2 = o + roh * (np.cos(ang) * v1 - np.sin(ang) * v2)
    pos = np.empty((n, 3))
    pos[0::3], pos[1::3], pos[2::3] = o, h1, h2

    charge = np.tile([-0.834, 0.417, 0.417], nmol)
    sigma = np.tile([0.315, 0.1, 0.1], nmol)
    epsilon = np.tile([0.636, 0.0, 0.0], nmol)
    excl = np.empty((nmol, 3, 2), dtype=np.int32)
    base = 3 * np.arange(nmol, dtype=np.int32)
    excl[:, 0, 0], excl[:, 0, 1] = base, base + 1
    excl[:, 1, 0], excl[:, 1, 1] = base, base + 2
    excl[:, 2, 0], excl[:, 2, 1] = base + 1, base + 2
    excl = excl.reshape(-1, 2)

    # solute: molecules nearest to the box centre, ordered by distance
    centre = np.array([L / 2, L / 2, L / 2])
    d2c = ((o - centre) ** 2).sum(1)
    order = np.argsort(d2c, kind="stable")
    nlig_mol = max(1, ligand_atoms // 3)
    nprot_mol = min(max(0, protein_atoms // 3), max(0, nmol - nlig_mol - 1))
    lig_mol = np.sort(order[:nlig_mol])
    prot_mol = np.sort(order[nlig_mol:nlig_mol + nprot_mol])
    for mols, q_scale in ((lig_mol, 0.5), (prot_mol, 0.6)):
        for k in range(3):
            ids = 3 * mols + k
            sigma[ids] = (0.34, 0.1, 0.1)[k]
            epsilon[ids] = (0.45, 0.0, 0.0)[k]     # H-like sites carry charge only, like water H
            charge[ids] = q_scale * (-0.5, 0.2, 0.3)[k]
    extra_excl, exc_pairs, exc_params = [], [], []
    for mols in (lig_mol, prot_mol):
        # chain consecutive molecules that are lattice neighbours (bonded-neighbour graph)
        for a_, b_ in zip(mols[:-1], mols[1:]):
            if ((site[a_] - site[b_]) ** 2).sum() < (1.01 * a) ** 2:
                for i in range(3):
                    for j in range(3):
                        extra_excl.append((3 * a_ + i, 3 * b_ + j))
                # one 1-4-like exception per link (its pair is also an exclusion)
                exc_pairs.append((3 * a_, 3 * b_))
                exc_params.append((0.5 * charge[3 * a_] * charge[3 * b_], 0.30, 0.20))
    if extra_excl:
        excl = np.concatenate([excl, np.array(extra_excl, dtype=np.int32)], axis=0)
    exc_pairs = np.array(exc_pairs, dtype=np.int32).reshape(-1, 2)
    exc_params = np.array(exc_params, dtype=np.float64).reshape(-1, 3)

    disp = np.zeros((n, 3))
    lig_atoms = (3 * lig_mol[:, None] + np.arange(3)[None, :]).reshape(-1)
    # land the ligand on interstitial positions of the lattice so that clashes stay finite
    dvec = np.asarray(displacement, dtype=np.float64)
    dvec = np.round(dvec / a) * a + 0.5 * a * (np.abs(dvec) > 0)
    disp[lig_atoms] = dvec

    sysd = NonbondedSystem(charge, sigma, epsilon, excl, exc_pairs, exc_params,
                           method=CUTOFF_PERIODIC, cutoff=cutoff, eps_rf=78.3,
                           box=np.array([L, L, L]), use_dispersion_correction=True)
    al = AlchemicalState(bias_method=ILOGISTIC, softcore_method=RATIONAL_SOFTCORE,
                         lambdac=0.5, lambda1=0.2, lambda2=0.5, alpha=0.1 / KCAL,
                         u0=110.0 * KCAL, w0coeff=0.0, umax=100.0 * KCAL,
                         ubcore=50.0 * KCAL, acore=0.0625)
    masses = np.tile([15.9994, 1.008, 1.008], nmol)
    return SDMCase("synthetic_%d" % n, sysd, pos, disp, al, None, masses)
