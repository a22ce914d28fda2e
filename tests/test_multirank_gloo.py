"""N>1 path on CPU: two gloo ranks each advance their shard of envs (through the host emulation of the device
phases -- test infrastructure, the product has no CPU path) and all-reduce the episode statistics; the result
must equal the unsharded run bit for bit (there is no cross-env coupling)."""
import os
import socket

import numpy as np
import pytest

from tests import common as T
from topotrafficrl_b200 import scenes, shard

E_TOTAL, STEPS, N = 6, 4, 20


def _run_shard(first, last, seed_actions=0):
    from tests.emu.emu import Emulator
    _, table, cfg, cfgd = T.highway_scene(N, 1.0, overrides={"duration": 2})
    emu = Emulator(cfg, table)
    st = scenes.make_highway_state(last - first, cfgd, seed=5, first_env=first)
    emu.pool, emu.autoreset = st.copy(), True
    stats = np.zeros((len(shard.STAT_FIELDS), last - first))  # per-env accumulators [field][env], like the device buffer
    acts = np.random.default_rng(seed_actions).integers(0, 5, size=(STEPS, E_TOTAL)).astype(np.int32)
    obs = []
    for k in range(STEPS):
        o = emu.step(st, acts[k, first:last], stats=stats)
        obs.append(o[0])
    return dict(zip(shard.STAT_FIELDS, stats.sum(axis=1))), np.stack(obs), st


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = shard.shard_range(E_TOTAL, rank, world)
    stats, obs, st = _run_shard(first, last)
    total = shard.all_reduce_stats(stats)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), obs=obs, veh_d=st.veh_d, first=first, last=last,
             total=np.array([total[k] for k in shard.STAT_FIELDS]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for total, world in [(6, 2), (7, 2), (4096, 8), (10, 3), (1, 4)]:
        ranges = [shard.shard_range(total, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_two_gloo_ranks_equal_unsharded_run(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    whole_stats, whole_obs, whole_st = _run_shard(0, E_TOTAL)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    np.testing.assert_array_equal(np.concatenate([p["obs"] for p in parts], axis=1), whole_obs)
    np.testing.assert_array_equal(np.concatenate([p["veh_d"] for p in parts], axis=1), whole_st.veh_d)
    want = np.array([whole_stats[k] for k in shard.STAT_FIELDS])
    assert want[0] > 0  # episodes finished
    for p in parts:
        np.testing.assert_allclose(p["total"], want, rtol=1e-12)


def _train_worker(rank, world, port, out_dir):
    """Two data-parallel ranks of the training driver's optimiser step (trainer.BatchedDQNAgent.step_optimizer): same initial
    network, different minibatches, gradients averaged over the ranks by one all-reduce."""
    import torch
    import torch.distributed as dist
    from topotrafficrl_b200.models import model_factory, size_model_config
    from topotrafficrl_b200.trainer import BatchedDQNAgent
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = size_model_config((15, 7), 3, {"type": "MultiLayerPerceptron", "layers": [32, 32]})
    agent = BatchedDQNAgent.__new__(BatchedDQNAgent)  # the learning half only: the acting half needs the CUDA rollout kernels
    agent.torch, agent.config = torch, dict(BatchedDQNAgent.default_config(), gamma=0.95)
    torch.manual_seed(0)
    agent.value_net, agent.target_net = model_factory(cfg), model_factory(cfg)
    agent.target_net.load_state_dict(agent.value_net.state_dict())
    agent.loss_function = torch.nn.functional.mse_loss
    agent.optimizer = torch.optim.SGD(agent.value_net.parameters(), lr=0.1)
    gen = torch.Generator().manual_seed(100)
    B = 16
    batches = []
    for _ in range(2):  # minibatch of rank 0, minibatch of rank 1
        batches.append((torch.rand(B, 15, 7, generator=gen), torch.randint(0, 3, (B,), generator=gen), torch.rand(B, generator=gen),
                        torch.rand(B, 15, 7, generator=gen), torch.rand(B, generator=gen) < 0.3))
    if world > 1:
        loss = agent.compute_bellman_residual(batches[rank])
    else:  # the single-process equivalent: mean of the two minibatch losses
        loss = (agent.compute_bellman_residual(batches[0]) + agent.compute_bellman_residual(batches[1])) / 2
    agent.step_optimizer(loss)
    flat = torch.cat([p.detach().reshape(-1) for p in agent.value_net.parameters()]).numpy()
    np.save(os.path.join(out_dir, f"params_w{world}_r{rank}.npy"), flat)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_two_gloo_ranks_average_gradients_in_the_training_step(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_train_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    _train_worker(0, 1, port, str(tmp_path))
    a, b = np.load(tmp_path / "params_w2_r0.npy"), np.load(tmp_path / "params_w2_r1.npy")
    single = np.load(tmp_path / "params_w1_r0.npy")
    np.testing.assert_array_equal(a, b)                      # the ranks stay in lockstep
    np.testing.assert_allclose(a, single, rtol=0, atol=1e-6)  # == one process on the mean loss (gradient clamp after averaging)
