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


def exchange_round(u_sc, state_of_replica, states, temperature, seed, round_index, n_sweeps=None):
    """New state index per replica after one round of Metropolis pair swaps.  Deterministic in
    (inputs, seed, round_index): every rank computes the same answer."""
    state_of = np.array(state_of_replica, dtype=np.int64, copy=True)
    n = len(state_of)
    if sorted(state_of.tolist()) != sorted(set(state_of.tolist())):
        raise ValueError("two replicas hold the same state")
    M = reduced_energy_matrix(u_sc, states, temperature)
    rng = np.random.Generator(np.random.Philox(key=[int(seed) & (2**64 - 1), int(round_index)]))
    n_sweeps = n * n if n_sweeps is None else n_sweeps
    for _ in range(n_sweeps):
        i, j = rng.integers(0, n, size=2)
        if i == j:
            continue
        si, sj = state_of[i], state_of[j]
        # swap the states of replicas i and j: delta = [W_sj(u_i) + W_si(u_j)] - [W_si(u_i) + W_sj(u_j)]
        delta = (M[i, sj] + M[j, si]) - (M[i, si] + M[j, sj])
        if delta <= 0.0 or rng.random() < np.exp(-delta):
            state_of[i], state_of[j] = sj, si
    return state_of


def all_gather_replica_info(u_sc_local, state_local, group=None):
    """The collective: (u_sc, state index) of every replica of every rank, in rank order.
    Tensors live on the current CUDA device under NCCL and on the CPU under gloo."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(np.stack([np.asarray(u_sc_local, np.float64),
                               np.asarray(state_local, np.float64)], axis=1), dtype=torch.float64, device=dev)
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0], 2), dtype=torch.float64, device=dev)   # rank-major concatenation
    dist.all_gather_into_tensor(out, t, group=group)
    out = out.cpu().numpy().reshape(-1, 2)
    return out[:, 0].copy(), out[:, 1].astype(np.int64)


def replica_exchange_step(u_sc_local, state_local, states, temperature, seed, round_index, rank, group=None):
    """One exchange round for this rank's replicas; returns their new state indices."""
    u_all, s_all = all_gather_replica_info(u_sc_local, state_local, group)
    new_all = exchange_round(u_all, s_all, states, temperature, seed, round_index)
    r_local = len(u_sc_local)
    return new_all[rank * r_local:(rank + 1) * r_local]
