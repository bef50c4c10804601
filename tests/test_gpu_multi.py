"""Two GPUs of one node: a replica's result does not depend on which GPU evaluates it, nor on how many
replicas share that GPU ("N replicas on k GPUs are bit-identical to 1 GPU").
One process drives contexts on both devices (every C-ABI entry point runs on its context's device and
restores the caller's)."""
import numpy as np
import pytest
import torch

from openmm_sdm_plugin_b200 import _lib, system as S
from openmm_sdm_plugin_b200.context import SDMContext

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")]


def _run(case, device, positions, states):
    R = len(positions)
    with SDMContext(case.system, case.displacement, n_replicas=R, pair_mode=_lib.PAIR_CLUSTER, device=device) as ctx:
        for r in range(R):
            ctx.set_positions(r, positions[r])
            ctx.set_alchemical(r, states[r])
        ctx.eval()
        f = np.empty((R, case.system.n_atoms, 3))
        sc = ctx.read_results(f)
        return f, sc


def test_replicas_are_bit_identical_on_one_and_two_gpus():
    case = S.cfg2()
    rng = np.random.default_rng(17)
    n = case.system.n_atoms
    pos = [case.positions + rng.normal(scale=0.003, size=(n, 3)) for _ in range(4)]
    states = S.atm_lambda_schedule(22)[3:7]
    f_all, sc_all = _run(case, 0, pos, states)                 # four replicas on GPU 0
    f_a, sc_a = _run(case, 0, pos[:2], states[:2])             # the same four, two per GPU
    f_b, sc_b = _run(case, 1, pos[2:], states[2:])
    f_split = np.concatenate([f_a, f_b])
    assert np.array_equal(f_all, f_split)                      # hybrid forces: every bit
    for r, (x, y) in enumerate(zip(sc_all, sc_a + sc_b)):
        assert x["status"] == y["status"] == 0
        for k in ("E1", "E2", "u", "u_sc", "sp", "pot_energy", "bind_e", "n_pairs1", "n_moved1", "n_moved2"):
            assert x[k] == y[k], (r, k, x[k], y[k])
    assert torch.cuda.current_device() == 0                    # the caller's device was left alone
