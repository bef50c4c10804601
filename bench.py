#!/usr/bin/env python
"""bench.py -- SDM dual-state force evals/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--replicas R] [--workload cfg2|cfg1|synthetic:N]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port), all host threads

One *step* = one dual-state evaluation (E1, u, u_sc, W, sp, PotEnergy, hybrid force F) of every
lambda-replica resident on a GPU.  `value` = evals/s summed over all replicas and GPUs with the
positions already in HBM; `e2e` = the same through the public C-ABI call path with HOST
buffers (H2D of positions and D2H of forces + scalars inside the timed region).

Multi-GPU: one process per GPU (torchrun), replicas sharded by rank, no data-path collective
(weak scaling: R replicas per GPU); the only collective is the (u_sc, state) all-gather that a
replica-exchange round needs, exercised once per step outside the kernel timing.

Other keys of the line (rank 0): `roofline` (pair kernel as an evaluation launches it, timed with CUDA events in this
run; `*_full_residency` = the same kernel with every resident block it can have, a diagnostic), `single_lambda`
(BASELINE.json configs[1]: one resident replica -- static evaluations with the headline list parameters and with
parameters sized for one replica, and the replica alone in the device MD loop), `md_loop` / `md_loop_pme` (constrained
dynamics on the device, measured ns/day), `cfg3` (22 windows under replica exchange, dealt over the ranks), `sweep`
(cfg1, cfg1 + HCT-GB, cfg2 with PME, cfg4, cfg5 sizes), `roofline_elementwise`, `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# periodic / non-periodic.  Ewald / PME direct space (3, 4) has no figure there: the
# reaction-field term (rinv + krf r^2 - crf and its derivative) is replaced by erfc(alpha r)/r and its
# derivative -- the exponential, a reciprocal and the five-term polynomial of Abramowitz & Stegun 7.1.26:
# 61 + 20 by our count.
FLOP_PER_PAIR = {2: 61.0, 1: 46.0, 0: 46.0, 3: 81.0, 4: 81.0}


def load_case(workload: str):
    from openmm_sdm_plugin_b200 import system as S
    if workload == "cfg2":
        return S.cfg2(), "cfg2: TEMOA-G1/G4 explicit solvent (20446 atoms, CutoffPeriodic RF, rc=1.0 nm, 38 displaced atoms)"
    if workload in ("cfg2:pme", "cfg2:pme+reciprocal"):
        c = S.cfg2()
        c.system.method = S.PME
        if workload.endswith("+reciprocal"):
            return c, "cfg2 as example/test_explicit.py:64 ships it: nonbondedMethod=PME complete -- direct space + " \
                      "reciprocal space (mesh 48x54x48, order 5, both states, cuFFT) on the device"
        return c, "cfg2 with nonbondedMethod=PME, DIRECT SPACE only (erfc pair terms + erf correction of the excluded " \
                  "pairs; the reciprocal part enters through sdm_set_external_dual)"
    if workload == "cfg1":
        return S.cfg1(), "cfg1: OA-G6/G3 (230 atoms, CutoffNonPeriodic 15 nm, 38 displaced atoms)"
    if workload == "cfg1:hct_gb":
        c = S.cfg1()
        c.system.eps_rf = 1.0       # desmonddmsfile75.py:451
        rng = np.random.default_rng(1)        # the fixture has no hct table: synthetic radii 0.12-0.20 nm, scale 0.72-0.88
        gb = S.GBSAHCTForce(SA="ACE")
        for a in range(c.system.n_atoms):
            gb.addParticle([c.system.charge[a], rng.uniform(0.12, 0.20), rng.uniform(0.72, 0.88)])
        gb.finalize()
        c.system.addForce(gb)
        return c, "cfg1 with implicitSolvent=HCT (GBSAHCTForce(SA='ACE'), desmonddmsfile75.py:454-465; synthetic GB radii): " \
                  "pair path + HCT-GB of both states on the device"
    if workload.startswith("synthetic:"):
        n = int(workload.split(":")[1])
        return S.synthetic_case(n_atoms=n, ligand_atoms=60, seed=1234), \
            "synthetic explicit-solvent ABFE box (%d atoms, CutoffPeriodic RF, rc=1.0 nm, 60 displaced atoms)" % n
    raise SystemExit("unknown workload " + workload)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


def bind_near_gpu(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index` (same NUMA node / PCIe
    root), before CUDA and the pinned host buffers are created: with one rank per GPU the
    host<->device copies of the e2e leg then do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        pick = cpus & os.sched_getaffinity(0)
        if pick:
            os.sched_setaffinity(0, pick)
            return len(pick)
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def elementwise_hbm(peak_gbs, n=16_000_000):
    """The path's two bandwidth-bound kernels as the reference's kernel interface has them --
    displace (MakeState2, langevin.cl:198-206: 48 B/atom) and hybrid-force mix (langevin.cl:72-87:
    64 B/atom) -- on arrays larger than L2 (256 MB each), timed with CUDA events; best of 10."""
    import torch
    from openmm_sdm_plugin_b200 import _lib
    L = _lib.lib()
    a, b, c = (torch.randn(n, 4, dtype=torch.float32, device="cuda") for _ in range(3))
    b[:, 3] = 0
    st = torch.cuda.current_stream().cuda_stream
    ops = [("sdm_k_make_state2", 48, lambda: L.sdm_k_make_state2(st, n, a.data_ptr(), b.data_ptr())),
           ("sdm_k_hybrid_force", 64, lambda: L.sdm_k_hybrid_force(st, n, a.data_ptr(), b.data_ptr(), c.data_ptr(), 0.37))]
    out = []
    for name, bpa, fn in ops:
        for _ in range(3):
            _lib.check(fn())
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for e0, e1 in ev:
            e0.record()
            _lib.check(fn())
            e1.record()
        torch.cuda.synchronize()
        ms = min(x.elapsed_time(y) for x, y in ev)
        gbs = bpa * n / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s",
                    "frac": gbs / peak_gbs, "bytes_per_atom": bpa, "n_atoms": n, "ms": ms})
    del a, b, c
    torch.cuda.empty_cache()
    return out


def cpu_baseline(case, seconds_budget=20.0, nthreads=1):
    """The reference's algorithm (oracle port: two full list builds + pair loops per eval,
    double precision, like LangevinIntegratorSDM::step on the Reference platform) timed on this
    host.  Bounded sample: whole evals of the same workload until the budget is used."""
    from openmm_sdm_plugin_b200 import system as S
    from oracle import oracle as O
    al = S.AlchemicalState(**vars(case.alch))
    t0 = time.perf_counter()
    n = 0
    while True:
        O.sdm_eval(case.system, al, case.displacement, case.positions, nthreads=nthreads,
                   want_state_forces=False)
        n += 1
        dt = time.perf_counter() - t0
        if dt > seconds_budget or n >= 64:
            break
    return n / dt, n


def run_reference(args):
    """--impl reference: the reference's own CPU path (its Reference platform cannot be built
    here -- OpenMM is not installed -- so this is the oracle port) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    case, wname = load_case(args.workload)
    # all host threads the box offers: torchrun exports OMP_NUM_THREADS=1, which only describes
    # its own default, so the affinity mask decides (the oracle takes an explicit thread count)
    try:
        nt = len(os.sched_getaffinity(0))
    except AttributeError:
        nt = os.cpu_count() or 1
    nt = max(nt, O.max_threads())
    from openmm_sdm_plugin_b200 import system as S
    from oracle import reference as R
    al = S.AlchemicalState(**vars(case.alch))
    use_ref = R.available()
    if use_ref:
        # the reference's OWN step code (oracle/_ref: its integrator + Reference-platform kernels
        # compiled from its sources): three force evaluations, state copies, execute(), Langevin
        # update.  OpenMM's NonbondedForce -- not available in this image -- is the oracle port,
        # called back for force group 2 with all host threads; force group 1 (bonded) is zero.
        n = case.system.n_atoms
        zeros = np.zeros((n, 3))
        params = R.params_from_alch(al)

        def force_fn(groups, pos):
            if groups == 4:
                r = O.nonbonded(case.system, pos, nthreads=nt)
                return r["E"], r["forces"]
            return 0.0, zeros

        def one_step():
            R.run(case.masses, case.positions, zeros, case.displacement, params, force_fn, steps=1)
    else:
        def one_step():
            O.sdm_eval(case.system, al, case.displacement, case.positions, nthreads=nt, want_state_forces=False)
    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = ("%d whole steps of the workload; step sequence, state handling, execute() and Langevin update by the reference's "
              "own compiled sources (oracle/_ref), NonbondedForce by the oracle port (OpenMM is not installable here)"
              if use_ref else "%d whole evals of the workload (oracle port; oracle/_ref not built)") % args.steps
    line = {"impl": "reference", "metric": "SDM dual-state force evals/s", "value": v, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "shipped fixture" if args.workload in ("cfg1", "cfg2") else "synthetic",
            "config": {"workload": wname, "replicas_per_step": 1,
                       "note": "one step = one dual-state eval of one replica on the host CPU"},
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": nt, "kind": "port",
                             "sample": sample},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def md_setup(case, freeze_solute=True):
    """Masses and constraints of the dynamics legs.  Bonded forces are external to the path (a10: they come
    from OpenMM's force group 1), so a free solute would move under nonbonded forces alone -- unphysical, and
    its runaway atoms would dictate the pair-list lifetime.  freeze_solute: atoms that have bonded terms (every
    molecule that is not a rigid three-site water or a single ion) get mass 0 (the library does not move massless
    particles, like ReferenceStochasticDynamicsSDM.cpp:150) and their constraints are dropped; the waters (SETTLE)
    and ions -- 99 % of the atoms, whose force field is complete without bonded terms -- move."""
    masses = case.masses.copy()
    pairs, dist = case.constraint_pairs, case.constraint_dist
    frozen = 0
    if freeze_solute and pairs is not None and len(pairs) > 0:
        n = len(masses)
        deg = np.bincount(pairs.ravel(), minlength=n)
        # connected components of the exclusion graph = molecules
        parent = np.arange(n)

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a
        for a, b in case.system.exclusions:
            ra, rb = find(int(a)), find(int(b))
            if ra != rb:
                parent[ra] = rb
        root = np.array([find(a) for a in range(n)])
        size = np.bincount(root, minlength=n)[root]
        rigid_water = (size == 3) & (deg == 2)
        solute = (size > 1) & ~rigid_water
        masses[solute] = 0.0
        frozen = int(solute.sum())
        keep = ~(solute[pairs[:, 0]] | solute[pairs[:, 1]])
        pairs, dist = pairs[keep], dist[keep]
    return masses, pairs, dist, frozen


def md_leg(case, R, args, device, stream, flush, states, rank, skin=None, nstlist=None, steps=None, dt=0.001,
           freeze_solute=True, reciprocal_pme=False):
    """Device-resident dynamics (sdm_md_step) as the reference's example runs it
    (example/test_explicit.py:166-169: 300 K, friction 0.1/ps, 1 fs): real masses, thermal
    velocities, the fixture's own distance constraints (SETTLE waters + X-H clusters) applied between
    the two update halves like ReferenceStochasticDynamicsSDM.cpp:250-252, Philox noise.  Positions
    and velocities never leave HBM.  Real motion ages the pair list: the stale-list rate and the
    rebuild share reported here are what the skin / nstlist choice really costs."""
    import torch
    from openmm_sdm_plugin_b200.context import SDMContext
    skin = args.skin if skin is None else skin
    nstlist = args.nstlist if nstlist is None else nstlist
    steps = max(args.steps, 100) if steps is None else steps
    n = case.system.n_atoms
    kT = 1.380658e-23 * 6.0221367e23 / 1000.0 * 300.0
    rng = np.random.default_rng(77 + rank)
    masses, cpairs, cdist, frozen = md_setup(case, freeze_solute)
    have_cons = cpairs is not None and len(cpairs) > 0
    vscale = np.where(masses > 0, np.sqrt(kT / np.maximum(masses, 1e-30)), 0.0)[:, None]
    with SDMContext(case.system, case.displacement, n_replicas=R, pair_mode=args.pair_mode, device=device,
                    skin=skin, nstlist=nstlist) as c:
        c.set_stream(stream.cuda_stream)
        if reciprocal_pme:
            c.enable_reciprocal_pme()
        c.md_init(masses, 300.0, 0.1, dt, seed=1234 + rank)
        if have_cons:
            c.md_set_constraints(cpairs, cdist, 1e-5)
        for r in range(R):
            c.set_alchemical(r, states[(rank * R + r) % len(states)])
            c.set_positions(r, case.positions)
            c.md_set_velocities(r, rng.normal(size=(n, 3)) * vscale)
        c.md_step(2 * nstlist)          # warm-up: first list, graph capture, constraint kick of the thermal start
        torch.cuda.synchronize()
        b0, (t0, r0) = c.info("n_list_builds"), c.md_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        e0.record(stream)
        c.md_step(steps)
        e1.record(stream)
        torch.cuda.synchronize()
        md_ms = e0.elapsed_time(e1) / steps
        builds, (t1, r1) = c.info("n_list_builds") - b0, c.md_counters()
        sc = c.read_results(None)
        ok = all(x["status"] == 0 for x in sc)
        ke = c.md_kinetic_energy(0)
        ncons = len(cdist) if have_cons else 0
        t_kin = 2.0 * ke / ((3 * (n - frozen) - ncons) * kT) * 300.0
    return {"value": R / (md_ms * 1e-3), "unit": "replica-steps/s", "ms_per_step": md_ms,
            "ns_per_day_per_replica": 1e3 / md_ms * dt * 1e-3 * 86400, "dt_ps": dt, "steps": steps,
            "skin_nm": skin, "nstlist": nstlist, "list_builds": int(builds),
            "steps_repeated_stale_list": int(r1 - r0), "steps_taken": int(t1 - t0), "status_ok": bool(ok),
            "kinetic_temperature_K": t_kin, "constraints": ncons, "frozen_solute_atoms": frozen,
            "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
            "note": "sdm_md_step on the device: dual-state eval + FP64 Langevin update + SETTLE/SHAKE constraints, "
                    "real masses, 300 K thermal start, friction 0.1/ps; bonded forces are external to the path, so "
                    "the atoms that have bonded terms (the host and the two ligands) are held fixed (mass 0) while the "
                    "rigid waters and the ions move; a step whose list went stale is not taken: the list is rebuilt and the step repeated "
                    "(counted in steps_repeated_stale_list, its time is inside ms_per_step)"}


def cfg3_leg(case, args, world, rank, device, stream, rounds=6, steps_per_round=10, n_windows=22):
    """BASELINE.json configs[2]: the explicit-solvent fixture as a 22-window lambda ladder under replica
    exchange, the windows dealt over the ranks ((3,3,3,3,3,3,2,2) on 8 GPUs).  Every replica runs
    constrained Langevin dynamics on its GPU (sdm_md_step); every `steps_per_round` steps the ranks
    all-gather (u_sc, state) -- the path's only collective, 16 bytes per replica -- run the same
    Metropolis sweep and APPLY the new assignment (sdm_set_alchemical).  A fixed-size job: strong
    scaling.  Device time between two events on the launching stream, max over ranks."""
    import torch
    import torch.distributed as dist
    from openmm_sdm_plugin_b200 import exchange as X, system as S
    from openmm_sdm_plugin_b200.context import SDMContext
    counts = X.split_replicas(n_windows, world)
    Rl, first = counts[rank], int(sum(counts[:rank]))
    states = S.atm_lambda_schedule(n_windows)
    n = case.system.n_atoms
    kT = 1.380658e-23 * 6.0221367e23 / 1000.0 * 300.0
    rng = np.random.default_rng(500 + rank)
    with SDMContext(case.system, case.displacement, n_replicas=Rl, pair_mode=args.pair_mode, device=device,
                    skin=args.md_skin, nstlist=args.md_nstlist) as c:
        c.set_stream(stream.cuda_stream)
        masses, cpairs, cdist, frozen = md_setup(case, True)
        vscale = np.where(masses > 0, np.sqrt(kT / np.maximum(masses, 1e-30)), 0.0)[:, None]
        c.md_init(masses, 300.0, 0.1, 0.001, seed=4321 + rank)
        c.md_set_constraints(cpairs, cdist, 1e-5)
        state_of = np.arange(first, first + Rl)
        for r in range(Rl):
            c.set_alchemical(r, states[state_of[r]])
            c.set_positions(r, case.positions)
            c.md_set_velocities(r, rng.normal(size=(n, 3)) * vscale)
        c.md_step(2 * steps_per_round)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stats, moved, same = {}, 0, True
        e0.record(stream)
        for k in range(rounds):
            c.md_step(steps_per_round)
            u_local = [x["u_sc"] for x in c.read_results(None)]   # one read-back for all replicas of the rank
            if world > 1:
                u_all, s_all = X.all_gather_replica_info(u_local, state_of, counts=counts)
            else:
                u_all, s_all = np.asarray(u_local), state_of.copy()
            new_all = X.exchange_round(u_all, s_all, states, 300.0, seed=2024, round_index=k, stats=stats)
            new_local = new_all[first:first + Rl]
            for r in range(Rl):
                if new_local[r] != state_of[r]:
                    c.set_alchemical(r, states[int(new_local[r])])
                    moved += 1
            state_of = new_local.copy()
            if world > 1:   # every rank must have reached the same assignment without a second collective
                h = torch.tensor([float((new_all * (np.arange(n_windows) + 1)).sum())], dtype=torch.float64, device="cuda")
                lo, hi = h.clone(), h.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                same = same and bool(lo.item() == hi.item())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ok = all(c.scalars(r)["status"] == 0 for r in range(Rl))
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
    nsteps = rounds * steps_per_round
    return {"workload": "cfg3: 22-window lambda ladder under replica exchange", "replicas_per_rank": counts,
            "value": n_windows * nsteps / (ms * 1e-3), "unit": "replica-steps/s (whole job)", "scaling": "strong",
            "ms_per_step": ms / nsteps, "ns_per_day_per_replica": nsteps / (ms * 1e-3) * 0.001 * 1e-3 * 86400,
            "exchange_every_steps": steps_per_round, "exchange_rounds": rounds,
            "swap_acceptance": stats.get("accepted", 0) / max(stats.get("proposed", 0), 1),
            "state_changes_applied_on_rank0": moved, "same_assignment_on_all_ranks": bool(same), "status_ok": bool(ok),
            "skin_nm": args.md_skin, "nstlist": args.md_nstlist,
            "note": "constrained dynamics per replica (real masses, 1 fs), (u_sc, state) all-gather + Metropolis sweep + "
                    "sdm_set_alchemical every %d steps; the ranks with the larger share set the pace" % steps_per_round}


def sweep_leg(args, device, stream, peak_tflops, quick=False):
    """The other configurations of BASELINE.json next to the headline workload, each measured the same
    way (resident evaluations, CUDA events on the launching stream, list rebuilds included, pair
    kernel bracketed on its own): cfg1 (230 atoms, all-pairs kernel) at growing replica batches, cfg4
    (synthetic 50 k atoms x 16 replicas), cfg5 (5 k .. 500 k atoms at 1 and 8 replicas)."""
    import torch
    from openmm_sdm_plugin_b200 import system as S
    from openmm_sdm_plugin_b200.context import SDMContext
    plan = [("cfg1", r) for r in (16, 128, 512)] + [("cfg1:hct_gb", 16), ("cfg1:hct_gb", 128), ("cfg2:pme", 16), ("cfg2:pme+reciprocal", 16), ("cfg2:pme+reciprocal", 1),
                                                      ("synthetic:50000", 16)]
    sizes = (5000, 20000, 100000) if quick else (5000, 10000, 20000, 50000, 100000, 200000, 500000)
    plan += [("synthetic:%d" % n, r) for n in sizes for r in (1, 8)]
    out = []
    for wl, R in plan:
        try:
            case, name = load_case(wl)
            n = case.system.n_atoms
            rng = np.random.default_rng(9)
            with SDMContext(case.system, case.displacement, n_replicas=R, pair_mode=args.pair_mode, device=device,
                            skin=args.skin, nstlist=args.nstlist) as c:
                c.set_stream(stream.cuda_stream)
                if wl.endswith("+reciprocal"):
                    c.enable_reciprocal_pme()
                S.apply_implicit_solvent(c, case.system)
                for r in range(R):
                    c.set_alchemical(r, case.alch)
                    c.set_positions(r, case.positions + (rng.normal(scale=0.002, size=(n, 3)) if r else 0.0))
                for _ in range(3):
                    c.eval()
                torch.cuda.synchronize()
                k = 20 if n * R < 2_000_000 else 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(k):
                    c.eval()
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / k
                c.set_timing(True)
                pk = []
                for _ in range(3):
                    c.eval()
                    pk.append(c.last_timing()[0])
                torch.cuda.synchronize()
                c.set_timing(False)
                sc = c.read_results(None)
                mode = int(c.info("pair_mode"))
            pairs = sum(x["n_pairs1"] + x["n_moved2"] for x in sc)
            flop = FLOP_PER_PAIR[int(case.system.method)] * pairs
            pm = float(np.median(pk))
            out.append({"workload": wl, "n_atoms": n, "replicas": R, "pair_mode": {1: "allpairs", 2: "cluster"}[mode],
                        "evals_per_s": R / (ms * 1e-3), "ms_per_step": ms, "pair_kernel_ms": pm,
                        "roofline_frac": flop / (pm * 1e-3) / 1e12 / peak_tflops,
                        "roofline_frac_step": flop / (ms * 1e-3) / 1e12 / peak_tflops,
                        "status_ok": all(x["status"] == 0 for x in sc)})
        except Exception as ex:
            out.append({"workload": wl, "replicas": R, "error": str(ex)[:160]})
    return out


def main():
    # NCCL prints its version banner to stdout under NCCL_DEBUG=VERSION: keep stdout to the one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--replicas", type=int, default=16, help="lambda-replicas resident per GPU")
    ap.add_argument("--nstlist", type=int, default=20)
    ap.add_argument("--skin", type=float, default=0.06)
    ap.add_argument("--pair-mode", type=int, default=0)
    ap.add_argument("--e2e-depth", type=int, default=4, help="batches in flight in the pipelined e2e leg (contexts/streams)")
    ap.add_argument("--e2e-threads", type=int, default=0,
                    help="1: one host thread per in-flight batch in the e2e leg (single GPU; measured no faster: "
                         "28.7k vs 30.0k evals/s); 0: one thread drives all batches")
    ap.add_argument("--exchange-every", type=int, default=10,
                    help="steps between replica-exchange all-gathers of (u_sc, state) in the e2e leg (N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-elementwise", action="store_true")
    ap.add_argument("--no-md-loop", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the cfg1 / cfg4 / cfg5 legs")
    ap.add_argument("--sweep-quick", action="store_true")
    ap.add_argument("--md-skin", type=float, default=0.2, help="pair-list skin of the dynamics legs (nm): real motion wants a wider one")
    ap.add_argument("--md-nstlist", type=int, default=40, help="upper limit of the list lifetime in the dynamics legs")
    ap.add_argument("--no-single-lambda", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries exactly ONE line, the JSON result: whatever libraries print there (NCCL writes its
    # version banner to stdout when NCCL_DEBUG is set in the environment) is routed to stderr by
    # pointing fd 1 at fd 2; the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    result_out = os.fdopen(result_fd, "w")
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from openmm_sdm_plugin_b200 import _lib, system as S
    from openmm_sdm_plugin_b200.context import PinnedArray, SDMContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SDM path has no CPU fallback")
    n_local_cpus = bind_near_gpu(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    case, wname = load_case(args.workload)
    n, R = case.system.n_atoms, args.replicas
    states = S.atm_lambda_schedule(max(world * R, 2))
    ctx = SDMContext(case.system, case.displacement, n_replicas=R, pair_mode=args.pair_mode,
                     device=local, skin=args.skin, nstlist=args.nstlist)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(1234 + rank)
    # per-replica positions: the fixture plus a small replica-specific thermal-like jitter
    h_pos = PinnedArray((R, n, 3))
    h_pos_b = PinnedArray((R, n, 3))
    h_f = PinnedArray((R, n, 3))
    base = np.stack([case.positions + rng.normal(scale=0.002, size=(n, 3)) for _ in range(R)])
    h_pos.array[...] = base
    for r in range(R):
        ctx.set_alchemical(r, states[(rank * R + r) % len(states)])
    ctx.set_positions_all(h_pos.array)
    # the e2e leg alternates between two host coordinate sets that differ by ~ the rms
    # displacement of a 1 fs step at 300 K, so every step uploads genuinely new positions
    h_pos_b.array[...] = base + rng.normal(scale=0.0006, size=(R, n, 3))

    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # 168 MB > 126 MB L2

    def exchange_gather(scalars):
        """(u_sc, state id) all-gather a replica-exchange round needs -- the path's only
        collective; 16 bytes per replica."""
        if world == 1:
            return
        from openmm_sdm_plugin_b200 import exchange as X
        X.all_gather_replica_info([s["u_sc"] for s in scalars], [rank * R + i for i in range(len(scalars))])

    # ---------------- resident leg: inputs already in HBM ----------------------------------
    def resident_step():
        ctx.eval()

    for _ in range(args.warmup):
        resident_step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ctx.launch_count()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                       # evict the previous step's working set from L2
        ev[k][0].record(stream)
        resident_step()
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_resident = time.perf_counter() - wall0
    launches = ctx.launch_count() - launches0
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    # the dominant kernel alone: the library brackets the pair kernel with CUDA events on the
    # launching stream (this disables the graph replay, so it is a separate pass over K steps)
    ctx.set_timing(True)
    pair_ms = []
    for k in range(args.steps):
        flush.zero_()
        resident_step()
        pair_ms.append(ctx.last_timing()[0])
    torch.cuda.synchronize()
    # diagnostic: the same kernel with every resident block it can have (the production launch leaves room on each SM
    # for the displaced-atom kernels that run beside it)
    ctx.set_timing(2)
    pair_full_ms = []
    for k in range(max(5, args.steps // 4)):
        flush.zero_()
        resident_step()
        pair_full_ms.append(ctx.last_timing()[0])
    torch.cuda.synchronize()
    ctx.set_timing(False)
    sampler.stop_flag = True
    if world > 1:
        tt = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt.item())
    sc = ctx.read_results(None)
    assert all(s["status"] == 0 for s in sc), [s["status"] for s in sc]
    value = world * R * args.steps / (t_ms * 1e-3)

    # ---------------- end-to-end leg: host buffers in, host buffers out ---------------------
    # Every step moves all positions in (pinned host -> HBM) and all forces + scalars out, through
    # the C-ABI calls a multi-replica driver makes: sdm_set_positions_all, sdm_eval,
    # sdm_enqueue_results, sdm_synchronize + sdm_collect_scalars.
    #  * serial: one batch in flight, the host waits for every step's results (latency);
    #  * pipelined (the reported e2e value): D batches of R replicas in flight, each on its own
    #    context + stream, so the H2D of batch k+1 and the D2H of batch k-1 ride the two copy
    #    engines while batch k computes.  A batch's results are consumed on the host (scalars read,
    #    exchange all-gather) before its context is reused.  Each step is still one batch of R
    #    evaluations; K steps are timed as a whole with a synchronize on both sides.
    D = max(1, args.e2e_depth)
    e_streams = [torch.cuda.Stream() for _ in range(D)]
    e_ctx, e_hf = [], []
    for d in range(D):
        cg = SDMContext(case.system, case.displacement, n_replicas=R, pair_mode=args.pair_mode,
                        device=local, skin=args.skin, nstlist=args.nstlist)
        cg.set_stream(e_streams[d].cuda_stream)
        for r in range(R):
            cg.set_alchemical(r, states[(rank * R + r) % len(states)])
        e_ctx.append(cg)
        e_hf.append(h_f if d == 0 else PinnedArray((R, n, 3)))
    flush_stream = torch.cuda.Stream()
    # single-precision transfers (sdm_set_positions_all_f32 / sdm_enqueue_results_f32): the precision
    # positions and forces have on the reference's live OpenCL path; half the PCIe bytes
    h_pos32, h_pos_b32 = PinnedArray((R, n, 3), np.float32), PinnedArray((R, n, 3), np.float32)
    h_pos32.array[...] = h_pos.array
    h_pos_b32.array[...] = h_pos_b.array
    e_hf32 = [PinnedArray((R, n, 3), np.float32) for _ in range(D)]
    io32 = [False]

    def e2e_submit(k, depth):
        d = k % depth
        cg = e_ctx[d]
        s = None
        if k >= depth:                       # consume the results of step k - depth
            cg.synchronize()
            s = cg.collect_scalars()
            if (k - depth) % args.exchange_every == 0:
                exchange_gather(s)           # one exchange period: the path's only collective
        if io32[0]:
            cg.set_positions_all(h_pos32.array if k % 2 == 0 else h_pos_b32.array)
            cg.eval()
            cg.enqueue_results(e_hf32[d].array)
        else:
            cg.set_positions_all(h_pos.array if k % 2 == 0 else h_pos_b.array)
            cg.eval()
            cg.enqueue_results(e_hf[d].array)
        return s

    def e2e_run_threaded(nsteps, depth):
        """One host thread per context (the C ABI's threading rule: distinct contexts may be driven
        from distinct threads): thread d submits steps d, d+depth, ... of its own batch, so a list
        build -- whose host synchronisations block the submitting thread for about a millisecond --
        stalls only that batch while the others keep the GPU fed.  Same work per step as the
        single-threaded loop; wall clock over all steps with a synchronize on both sides."""
        errors, lasts = [], [None] * depth

        def worker(d):
            try:
                torch.cuda.set_device(local)
                cg = e_ctx[d]
                first = True
                for k in range(d, nsteps, depth):
                    with torch.cuda.stream(flush_stream):
                        flush.zero_()
                    if not first:
                        cg.synchronize()
                        cg.collect_scalars()
                    cg.set_positions_all(h_pos.array if k % 2 == 0 else h_pos_b.array)
                    cg.eval()
                    cg.enqueue_results(e_hf[d].array)
                    first = False
                if not first:
                    cg.synchronize()
                    lasts[d] = cg.collect_scalars()
            except Exception as ex:   # reported by the caller
                errors.append(ex)

        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(d,)) for d in range(depth)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if errors:
            raise errors[0]
        return dt, [x for x in lasts if x is not None][-1]

    def e2e_run(nsteps, depth, timed, flush_each=False):
        """nsteps steps with `depth` batches in flight; returns (seconds, last scalars).  flush_each: a 160 MiB
        write runs beside every step (needed when the batches in flight fit the L2 together)."""
        if timed and depth > 1 and world == 1 and args.e2e_threads:
            return e2e_run_threaded(nsteps, depth)
        torch.cuda.synchronize()
        if world > 1 and timed:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        last = None
        for k in range(nsteps):
            if timed and flush_each:
                with torch.cuda.stream(flush_stream):
                    flush.zero_()            # keeps evicting L2 next to the pipeline, unordered
            last = e2e_submit(k, depth) or last
        for k in range(nsteps, nsteps + min(depth, nsteps)):   # drain in submission order
            cg = e_ctx[k % depth]
            cg.synchronize()
            last = cg.collect_scalars()
            if (k - depth) % args.exchange_every == 0:
                exchange_gather(last)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, last

    e2e_run(max(args.warmup, D), D, False)
    # Steady state: every context rebuilds its list every nstlist of ITS evaluations, i.e. the
    # pipeline as a whole sees one rebuild per nstlist steps.  Fresh contexts would all be young and
    # a short timed window would contain no rebuild at all, so the lists are pre-aged to evenly
    # staggered ages (untimed resident evaluations): the timed K steps then carry their K/nstlist
    # share of list builds, like the resident leg does.
    for d in range(D):
        age_now = int(e_ctx[d].info("list_age"))
        target = (args.nstlist - 2 - (d * args.nstlist) // D) % args.nstlist
        for _ in range((target - age_now) % args.nstlist):
            e_ctx[d].eval()
        e_ctx[d].synchronize()
    # L2 rule of the timed region: the D batches in flight have disjoint working sets (cfg2, R = 16: ~125 MB each --
    # rows 67 MB, positions / forces 39 MB, sorted atoms 10 MB, accumulators 8 MB) that follow each other through
    # the 126 MB L2, and every step's inputs arrive from the host: with D >= 2 no batch finds its previous step's
    # data cached, so no flush kernel is needed (it would only steal SM time from the pipeline).  With one batch
    # in flight the flush write runs beside every step.  The figure WITH the flush kernel is reported next to it.
    ws_mb = D * 125.0 * (R * n) / (16 * 20446.0)
    e2e_needs_flush = D < 2 or ws_mb <= 2 * 126.0
    e2e_t, s = e2e_run(args.steps, D, True, flush_each=e2e_needs_flush)
    assert all(x["status"] == 0 for x in s), [x["status"] for x in s]
    e2e_flush_t, s = e2e_run(args.steps, D, True, flush_each=True)
    assert all(x["status"] == 0 for x in s), [x["status"] for x in s]
    io32[0] = True
    e2e_run(max(args.warmup, D), D, False)
    e2e32_t, s = e2e_run(args.steps, D, True, flush_each=e2e_needs_flush)
    assert all(x["status"] == 0 for x in s), [x["status"] for x in s]
    f32_dev = float(np.abs(e_hf32[0].array - e_hf[0].array).max() / max(np.abs(e_hf[0].array).max(), 1e-30))
    io32[0] = False
    e2e_run(args.warmup, 1, False)
    e2e_serial_t, s = e2e_run(args.steps, 1, True, flush_each=True)
    assert all(x["status"] == 0 for x in s), [x["status"] for x in s]
    for cg in e_ctx:
        cg.close()
    if world > 1:
        tt = torch.tensor([e2e_t, e2e_serial_t, e2e32_t, e2e_flush_t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_t, e2e_serial_t, e2e32_t, e2e_flush_t = (float(tt[k].item()) for k in range(4))
    e2e_value = world * R * args.steps / e2e_t
    h2d = R * n * 3 * 8
    d2h = R * n * 3 * 8 + R * 8 * 20

    # ---------------- roofline of the dominant kernel ----------------------------------------
    pk = peaks()
    algo_pairs = sum(s["n_pairs1"] + s["n_moved2"] for s in sc)          # per launch (all replicas)
    flop = FLOP_PER_PAIR[int(case.system.method)] * algo_pairs
    pair_ms_avg = float(np.mean(pair_ms))
    achieved = flop / (pair_ms_avg * 1e-3) / 1e12
    peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    mode = int(ctx.info("pair_mode"))
    fma_measured = ctx.info("fp32_fma_tflops_measured")   # sustained packed-FMA rate of THIS device, timed in this run
    # DRAM traffic of the dominant kernel comes from the committed ncu capture of this very
    # configuration (profiles/): a number measured under a profiler is never timed here
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_pair_kernel_traffic.json")))
        if tr["workload"] == args.workload and tr["replicas_per_gpu"] == R and mode == 2:
            traffic = tr["dram_bytes_read_per_launch"] + tr["dram_bytes_write_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "fp32_simt", "bound_note": "neither hbm nor tensor: the pair kernel is bound by the FP32 SIMT pipe (SURVEY.md 8d)", "kernel": "pair_row_kernel" if mode == 2 else "allpairs_kernel",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz (%s MEASURED_PEAKS.json)" % pk["source"],
                "peak_fma_measured_in_this_run": fma_measured, "frac_of_measured_fma_peak": achieved / fma_measured if fma_measured else None,
                "algorithmic_pairs_per_launch": algo_pairs, "flop_per_pair": FLOP_PER_PAIR[int(case.system.method)],
                "kernel_ms": pair_ms_avg, "kernel_share_of_step": pair_ms_avg * args.steps / t_ms,
                "kernel_launch": "as in production: for this batch 20 of the 24 possible resident blocks per SM, the rest of the "
                                 "register file is left to the FP64 displaced-atom kernels that run beside it; timed here with "
                                 "those kernels behind it (events on the launching stream)",
                "kernel_ms_full_residency": float(np.mean(pair_full_ms)),
                "frac_full_residency": achieved * pair_ms_avg / float(np.mean(pair_full_ms)) / peak}

    line = {"metric": "SDM dual-state force evals/s", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32",
            "dtype_detail": "f32 pair terms (packed f32x2), f64 moved-pair terms / u / scalars / mix, 64-bit fixed-point force accumulation",
            "data": "shipped fixture (positions+topology), replica jitter synthetic" if args.workload in ("cfg1", "cfg2") else "synthetic",
            "config": {"workload": wname, "replicas_per_gpu": R, "lambda_schedule": "22-window ILogistic ladder",
                       "pair_mode": {1: "allpairs", 2: "cluster"}[mode], "skin_nm": args.skin, "nstlist": args.nstlist,
                       "l2": "flushed between timed steps (160 MiB write, L2 = 126 MB)", "timing": "CUDA events per step on the launching stream, max over ranks"},
            "evals_per_s_per_replica": value / (world * R),
            "ns_per_day_per_replica_upper_bound": value / (world * R) * 1e-6 * 86400,
            "roofline": roofline, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_t / args.steps, "batches_in_flight": D,
                    "pcie_gb_s_per_rank": (h2d + d2h) * args.steps / e2e_t / 1e9,
                    "host_gb_s_all_ranks": world * (h2d + d2h) * args.steps / e2e_t / 1e9,
                    "f32_io": {"value": world * R * args.steps / e2e32_t, "unit": "evals/s",
                               "ms_per_step": 1e3 * e2e32_t / args.steps,
                               "h2d_bytes_per_step": R * n * 3 * 4, "d2h_bytes_per_step": R * n * 3 * 4 + R * 8 * 20,
                               "pcie_gb_s_per_rank": (R * n * 3 * 8 + R * 8 * 20) * args.steps / e2e32_t / 1e9,
                               "max_rel_force_deviation_from_f64_io": f32_dev,
                               "note": "same pipeline through sdm_set_positions_all_f32 / sdm_enqueue_results_f32: positions and "
                                       "forces cross PCIe in single precision (the precision of the reference's OpenCL path, "
                                       "OpenCLSDMKernels.cpp:96-103); arithmetic on the device unchanged"},
                    "l2": ("%d batches in flight x ~%.0f MB working set each = %.0f MB > 126 MB L2 and fresh host inputs every "
                           "step: inputs larger than L2, no flush kernel" % (D, ws_mb / D, ws_mb)) if not e2e_needs_flush else
                          "160 MiB flush write beside every step",
                    "with_flush_kernel": {"value": world * R * args.steps / e2e_flush_t, "ms_per_step": 1e3 * e2e_flush_t / args.steps,
                                          "note": "same pipeline with a 160 MiB flush write launched beside every step (how this "
                                                  "leg ran in round 1): the fill kernel takes its ~45 us of SM time per step"},
                    "exchange_every_steps": args.exchange_every if world > 1 else None,
                    "list_builds": "lists pre-aged to staggered ages: the timed steps carry K/nstlist list builds",
                    "host_threads": D if (world == 1 and args.e2e_threads and D > 1) else 1,
                    "host_cpus_local_to_gpu": n_local_cpus,
                    "serial": {"value": world * R * args.steps / e2e_serial_t, "ms_per_step": 1e3 * e2e_serial_t / args.steps,
                               "note": "one batch in flight, host waits for each step's results"},
                    "note": "C-ABI calls with pinned HOST buffers: sdm_set_positions_all (H2D) + sdm_eval + sdm_enqueue_results (D2H of forces and scalars) "
                            "+ sdm_synchronize/sdm_collect_scalars per step inside the timed region (host wall clock over all K steps, synchronize on both "
                            "sides); D batches of R replicas in flight on D contexts/streams so copies overlap the kernels of the neighbouring batches; "
                            "see l2 for the cache rule of this leg"},
            "clocks": sampler.result(), "wall_s_resident_leg": wall_resident}

    if rank == 0 and R > 1 and not args.no_single_lambda:
        # BASELINE.json configs[1] is quoted for a single lambda on one B200: the same workload
        # with ONE resident replica (latency bound: 8 kernels of ~20k atoms per evaluation)
        with SDMContext(case.system, case.displacement, n_replicas=1, pair_mode=args.pair_mode, device=local,
                        skin=args.skin, nstlist=args.nstlist) as c1:
            c1.set_stream(stream.cuda_stream)
            c1.set_alchemical(0, case.alch)
            c1.set_positions(0, base[0])
            for _ in range(max(args.warmup, 3)):
                c1.eval()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ns = max(args.steps, 20)
            e0.record(stream)
            for _ in range(ns):
                c1.eval()
            e1.record(stream)
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / ns
            sc1 = c1.scalars(0)
            assert sc1["status"] == 0
            # the evaluations between two list builds alone (the latency a one-Context adapter sees on
            # most steps), and the pair kernel alone
            c1.invalidate_list()
            c1.eval()
            torch.cuda.synchronize()
            nb = max(args.nstlist - 2, 1)
            e0.record(stream)
            for _ in range(nb):
                c1.eval()
            e1.record(stream)
            torch.cuda.synchronize()
            ms1_nobuild = e0.elapsed_time(e1) / nb
            c1.set_timing(True)
            pk1 = []
            for _ in range(5):
                c1.eval()
                pk1.append(c1.last_timing()[0])
            torch.cuda.synchronize()
            c1.set_timing(False)
        # the same leg with list parameters suited to one replica: the pair kernel is latency bound there, so a wider
        # skin costs little and halves the list builds (the library's stale-list check guards validity either way)
        tuned = {}
        try:
            with SDMContext(case.system, case.displacement, n_replicas=1, pair_mode=args.pair_mode, device=local,
                            skin=0.12, nstlist=40) as c1:
                c1.set_stream(stream.cuda_stream)
                c1.set_alchemical(0, case.alch)
                c1.set_positions(0, base[0])
                for _ in range(45):
                    c1.eval()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                nt = 240
                e0.record(stream)
                for _ in range(nt):
                    c1.eval()
                e1.record(stream)
                torch.cuda.synchronize()
                assert c1.scalars(0)["status"] == 0
                tuned = {"ms_per_step": e0.elapsed_time(e1) / nt, "skin_nm": 0.12, "nstlist": 40, "steps": nt}
                tuned["value"] = 1e3 / tuned["ms_per_step"]
        except Exception as ex:
            tuned = {"error": str(ex)[:200]}
        # and what one replica alone gets out of the device MD loop (real motion, constraints, the list lifetime the
        # loop plans itself): the measured ns/day of a single lambda
        try:
            md1 = md_leg(case, 1, args, local, stream, flush, states, rank, skin=args.md_skin, nstlist=args.md_nstlist,
                         steps=400)
            md1 = {k: md1[k] for k in ("ms_per_step", "ns_per_day_per_replica", "skin_nm", "nstlist", "list_builds", "steps",
                                       "steps_repeated_stale_list", "status_ok", "kinetic_temperature_K")}
        except Exception as ex:
            md1 = {"error": str(ex)[:200]}
        flop1 = FLOP_PER_PAIR[int(case.system.method)] * (sc1["n_pairs1"] + sc1["n_moved2"])
        line["single_lambda"] = {"value": 1e3 / ms1, "unit": "evals/s", "ms_per_step": ms1, "replicas": 1,
                                 "ms_per_step_between_list_builds": ms1_nobuild,
                                 "pair_kernel_ms": float(np.median(pk1)),
                                 "roofline_frac_pair_kernel": flop1 / (float(np.median(pk1)) * 1e-3) / 1e12 / peak,
                                 "roofline_frac_step": flop1 / (ms1 * 1e-3) / 1e12 / peak,
                                 "ns_per_day_upper_bound": 1e3 / ms1 * 1e-6 * 86400,
                                 "list_parameters": {"skin_nm": args.skin, "nstlist": args.nstlist},
                                 "with_single_replica_list_parameters": tuned,
                                 "md_loop_single_replica": md1,
                                 "note": "BASELINE.json configs[1] as worded: ONE resident replica, positions in HBM, list rebuilds "
                                         "(every nstlist evaluations) included in ms_per_step, no L2 flush; latency bound: the "
                                         "critical path is refresh -> pair kernel (one unit per resident warp) -> scalars -> mix"}
    if rank == 0 and not args.no_md_loop and case.masses is not None:
        try:
            line["md_loop"] = md_leg(case, R, args, local, stream, flush, states, rank, skin=args.md_skin,
                                     nstlist=args.md_nstlist)
        except Exception as ex:   # an extra, never the reason for a missing bench line
            line["md_loop"] = {"error": str(ex)[:200]}
        if args.workload == "cfg2":
            # the same dynamics with the electrostatics example/test_explicit.py:64 ships: nonbondedMethod=PME,
            # direct + reciprocal space of both states on the device
            try:
                pme_case, _ = load_case("cfg2:pme+reciprocal")
                line["md_loop_pme"] = md_leg(pme_case, R, args, local, stream, flush, states, rank, skin=args.md_skin,
                                             nstlist=args.md_nstlist, reciprocal_pme=True)
                line["md_loop_pme"]["note"] = "md_loop with nonbondedMethod=PME complete (mesh 48x54x48, order 5): " + \
                    line["md_loop_pme"]["note"]
            except Exception as ex:
                line["md_loop_pme"] = {"error": str(ex)[:200]}
    if not args.no_cfg3 and args.workload == "cfg2" and case.constraint_pairs is not None:
        try:   # every rank takes part (the ladder is dealt over the ranks)
            res = cfg3_leg(case, args, world, rank, local, stream)
            if rank == 0:
                line["cfg3"] = res
        except Exception as ex:
            if rank == 0:
                line["cfg3"] = {"error": str(ex)[:200]}
    if rank == 0 and world == 1 and not args.no_sweep and args.workload == "cfg2":
        line["sweep"] = sweep_leg(args, local, stream, peak, quick=args.sweep_quick)
    if rank == 0 and not args.no_elementwise:
        # the bandwidth-bound kernels of the path against the measured HBM copy bandwidth
        line["roofline_elementwise"] = elementwise_hbm(pk["hbm_gbs"])
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ns = cpu_baseline(case, args.cpu_seconds, 1)
        line["cpu_baseline"] = {"value": v, "unit": "evals/s", "cores": 1, "kind": "port",
                                "sample": "%d whole dual-state evals of the same workload (1 replica each), "
                                          "two-pass like the reference, 1 thread like OpenMM's Reference platform" % ns}
    if rank == 0:
        result_out.write(json.dumps(line) + "\n")
        result_out.flush()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
