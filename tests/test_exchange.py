"""Host logic of the lambda-replica exchange step, including the world_size-2 gloo path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openmm_sdm_plugin_b200 import exchange as X, system as S


def test_bias_energy_matches_survey_known_answers():
    # SURVEY.md Appendix A.4, ILogistic rows
    st = S.AlchemicalState(bias_method=S.ILOGISTIC, lambda1=0.0, lambda2=0.5, alpha=0.0239005736137667,
                           u0=460.24, w0coeff=0.0)
    assert abs(X.bias_energy(st, 100.0) - 230.123813041422) < 1e-9
    st = S.AlchemicalState(bias_method=S.ILOGISTIC, lambda1=0.025, lambda2=0.025, alpha=0.0, u0=0.0)
    assert abs(X.bias_energy(st, 3.60861626248879) - 0.0902154065622199) < 1e-12
    assert X.bias_energy(S.AlchemicalState(bias_method=S.LINEAR, lambdac=0.3), 10.0) == pytest.approx(3.0)
    q = S.AlchemicalState(bias_method=S.QUADRATIC, gammac=0.1, wbcoeff=0.5, w0coeff=1.0)
    assert X.bias_energy(q, 4.0) == pytest.approx(0.5 * 0.1 * 16 + 2.0 + 1.0)


def test_exchange_round_is_a_permutation_and_deterministic():
    states = S.atm_lambda_schedule(8)
    rng = np.random.default_rng(0)
    u = rng.normal(20.0, 40.0, size=8)
    s0 = np.arange(8)
    a = X.exchange_round(u, s0, states, 300.0, seed=7, round_index=3)
    b = X.exchange_round(u, s0, states, 300.0, seed=7, round_index=3)
    assert np.array_equal(a, b) and sorted(a.tolist()) == list(range(8))
    c = X.exchange_round(u, s0, states, 300.0, seed=7, round_index=4)
    assert sorted(c.tolist()) == list(range(8))
    with pytest.raises(ValueError):
        X.exchange_round(u, np.zeros(8, int), states, 300.0, 1, 0)


def test_downhill_swaps_are_always_accepted():
    # two states with very different slopes: the replica with the larger u ends in the state with
    # the smaller slope whatever the random numbers are (delta <= 0 every time it is proposed)
    states = [S.AlchemicalState(bias_method=S.LINEAR, lambdac=0.0), S.AlchemicalState(bias_method=S.LINEAR, lambdac=1.0)]
    for seed in range(5):
        out = X.exchange_round([1000.0, -1000.0], [1, 0], states, 300.0, seed, 0, n_sweeps=200)
        assert out.tolist() == [0, 1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    states = S.atm_lambda_schedule(6)
    r_local = 3
    rng = np.random.default_rng(100 + rank)
    u_local = rng.normal(0.0, 50.0, size=r_local)
    s_local = np.arange(rank * r_local, (rank + 1) * r_local)
    u_all, s_all = X.all_gather_replica_info(u_local, s_local)
    new_local = X.replica_exchange_step(u_local, s_local, states, 300.0, seed=11, round_index=0, rank=rank)
    ret[rank] = (u_all.tolist(), s_all.tolist(), new_local.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_agree_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    (u0, s0, n0), (u1, s1, n1) = ret[0], ret[1]
    assert u0 == u1 and s0 == s1 == list(range(6))          # same gathered view on both ranks
    expected = X.exchange_round(u0, s0, S.atm_lambda_schedule(6), 300.0, 11, 0).tolist()
    assert n0 + n1 == expected                               # each rank took its own slice
    assert sorted(n0 + n1) == list(range(6))


def test_split_of_22_windows_over_8_gpus():
    assert X.split_replicas(22, 8) == [3, 3, 3, 3, 3, 3, 2, 2]        # BASELINE.json configs[2]
    assert X.split_replicas(22, 1) == [22] and X.split_replicas(22, 4) == [6, 6, 5, 5]
    assert sum(X.split_replicas(22, 3)) == 22


def _ragged_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = X.split_replicas(5, world)                      # 3 + 2: ranks hold different numbers of replicas
    states = S.atm_lambda_schedule(5)
    first = sum(counts[:rank])
    rng = np.random.default_rng(200 + rank)
    u_local = rng.normal(0.0, 50.0, size=counts[rank])
    s_local = np.arange(first, first + counts[rank])
    u_all, s_all = X.all_gather_replica_info(u_local, s_local, counts=counts)
    new_local = X.replica_exchange_step(u_local, s_local, states, 300.0, seed=5, round_index=2, rank=rank, counts=counts)
    ret[rank] = (u_all.tolist(), s_all.tolist(), new_local.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_with_unequal_replica_counts_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ragged_worker, args=(2, port, ret), nprocs=2, join=True)
    (u0, s0, n0), (u1, s1, n1) = ret[0], ret[1]
    assert u0 == u1 and s0 == s1 == list(range(5)) and len(n0) == 3 and len(n1) == 2
    expected = X.exchange_round(u0, s0, S.atm_lambda_schedule(5), 300.0, 5, 2).tolist()
    assert n0 + n1 == expected and sorted(n0 + n1) == list(range(5))
