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
    stats = np.zeros((8, last - first))  # per-env accumulators [field][env], like the device buffer
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
