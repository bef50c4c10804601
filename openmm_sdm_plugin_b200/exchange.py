"""Replica exchange over lambda-states: the path's only multi-GPU exchange step (SURVEY.md
section 8(e)).  The reference has no exchange code (its users drive it from the AToM workflow
scripts); this module defines the minimal, deterministic version the benchmark exercises.

Replicas never move: every rank holds R_local replicas and only their *state index* (which
AlchemicalState of the ladder they run) is permuted.  One round =
  1. all-gather of (u_sc, state index) per replica -- 16 bytes per replica, `torch.distributed`
     (NCCL between GPUs, gloo in the CPU tests);
  2. every rank evaluates the reduced bias energies beta*W_state(u_sc) locally (O(1) per entry,
     the bias block of platforms/reference/src/ReferenceSDMKernels.cpp:247-282);
  3. every rank applies the same sequence of Metropolis pair swaps, driven by a counter-based
     RNG seeded with (seed, round), so all ranks reach the same assignment without a second
     collective.
"""
from __future__ import annotations

import math

import numpy as np

from .system import ILOGISTIC, LINEAR, QUADRATIC, AlchemicalState

BOLTZ = 1.380658e-23 * 6.0221367e23 / 1000.0   # kJ/mol/K, platforms/opencl/src/OpenCLSDMKernels.cpp:57-60


def bias_energy(state: AlchemicalState, u_sc):
    """W_state(u_sc) in kJ/mol (ReferenceSDMKernels.cpp:247-282)."""
    B = np.asarray(u_sc, np.float64)
    if state.bias_method == QUADRATIC:
        return 0.5 * state.gammac * B * B + state.wbcoeff * B + state.w0coeff
    if state.bias_method == ILOGISTIC:
        e = state.lambda2 * B + state.w0coeff
        if state.alpha > 0:
            ee = 1.0 + np.exp(-state.alpha * (B - state.u0))
            e = e + (state.lambda2 - state.lambda1) / state.alpha * np.log(ee)
        return e
    if state.bias_method == LINEAR:
        return state.lambdac * B
    raise ValueError("unknown bias method %r" % state.bias_method)


def reduced_energy_matrix(u_sc, states, temperature):
    """M[i, s] = beta * W_s(u_sc[i]) for every replica i and ladder state s."""
    beta = 1.0 / (BOLTZ * temperature)
    u = np.asarray(u_sc, np.float64)
    return beta * np.stack([bias_energy(s, u) for s in states], axis=1)


def exchange_round(u_sc, state_of_replica, states, temperature, seed, round_index, n_sweeps=None, stats=None):
    """New state index per replica after one round of Metropolis pair swaps.  Deterministic in
    (inputs, seed, round_index): every rank computes the same answer.  stats (a dict) receives the
    numbers of proposed and accepted swaps."""
    state_of = np.array(state_of_replica, dtype=np.int64, copy=True)
    n = len(state_of)
    if sorted(state_of.tolist()) != sorted(set(state_of.tolist())):
        raise ValueError("two replicas hold the same state")
    M = reduced_energy_matrix(u_sc, states, temperature).tolist()
    rng = np.random.Generator(np.random.Philox(key=[int(seed) & (2**64 - 1), int(round_index)]))
    n_sweeps = n * n if n_sweeps is None else n_sweeps
    # all random numbers of the round up front (one pair and one uniform per proposal, used or not), then a plain
    # Python loop over lists: the sweep is sequential by nature but costs ~0.4 us per proposal this way instead of
    # ~10 us with one generator call per proposal -- it sits between two dynamics segments on every rank
    pairs = rng.integers(0, n, size=(n_sweeps, 2)).tolist()
    unif = rng.random(n_sweeps).tolist()
    st = state_of.tolist()
    proposed = accepted = 0
    for (i, j), x in zip(pairs, unif):
        if i == j:
            continue
        si, sj = st[i], st[j]
        # swap the states of replicas i and j: delta = [W_sj(u_i) + W_si(u_j)] - [W_si(u_i) + W_sj(u_j)]
        delta = (M[i][sj] + M[j][si]) - (M[i][si] + M[j][sj])
        proposed += 1
        if delta <= 0.0 or (delta < 745.0 and x < math.exp(-delta)):
            st[i], st[j] = sj, si
            accepted += 1
    state_of = np.array(st, dtype=np.int64)
    if stats is not None:
        stats["proposed"] = stats.get("proposed", 0) + proposed
        stats["accepted"] = stats.get("accepted", 0) + accepted
    return state_of


def split_replicas(n_replicas: int, world: int):
    """Replicas per rank when n_replicas lambda-windows are dealt over `world` GPUs as evenly as
    possible, larger shares first: 22 over 8 -> (3, 3, 3, 3, 3, 3, 2, 2) (BASELINE.json configs[2])."""
    base, extra = divmod(int(n_replicas), int(world))
    return [base + (1 if r < extra else 0) for r in range(world)]


def all_gather_replica_info(u_sc_local, state_local, group=None, counts=None):
    """The collective: (u_sc, state index) of every replica of every rank, in rank order.
    Tensors live on the current CUDA device under NCCL and on the CPU under gloo.  counts = replicas
    per rank when the ranks hold different numbers (every rank then sends max(counts) rows, the
    padding rows are dropped on arrival)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    world = dist.get_world_size(group)
    rows = len(u_sc_local) if counts is None else max(counts)
    mine = np.zeros((rows, 2))
    mine[:len(u_sc_local), 0] = np.asarray(u_sc_local, np.float64)
    mine[:len(u_sc_local), 1] = np.asarray(state_local, np.float64)
    t = torch.tensor(mine, dtype=torch.float64, device=dev)
    out = torch.empty((world * rows, 2), dtype=torch.float64, device=dev)   # rank-major concatenation
    dist.all_gather_into_tensor(out, t, group=group)
    out = out.cpu().numpy().reshape(world, rows, 2)
    if counts is not None:
        out = np.concatenate([out[r, :counts[r]] for r in range(world)], axis=0)
    else:
        out = out.reshape(-1, 2)
    return out[:, 0].copy(), out[:, 1].astype(np.int64)


def replica_exchange_step(u_sc_local, state_local, states, temperature, seed, round_index, rank, group=None,
                          counts=None):
    """One exchange round for this rank's replicas; returns their new state indices."""
    u_all, s_all = all_gather_replica_info(u_sc_local, state_local, group, counts)
    new_all = exchange_round(u_all, s_all, states, temperature, seed, round_index)
    if counts is None:
        r_local = len(u_sc_local)
        return new_all[rank * r_local:(rank + 1) * r_local]
    first = int(sum(counts[:rank]))
    return new_all[first:first + counts[rank]]
