"""Live pinning of the CPU oracle against the UNMODIFIED reference, imported from /root/reference through
oracle/ref_shim.py.  Runs only in the build container (the reference does not exist on the GPU box: skipped
there); uses seeds that are NOT in the committed golden vectors, so it widens their coverage each time it runs.
Bars: discrete fields bit-exact, continuous state 1e-9 per resynced sub-step / 1e-6 per env-step."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/ttrl_env"), reason="reference tree not present")

from tests import common as T  # noqa: E402
from topotrafficrl_b200 import abi, scenes  # noqa: E402


@pytest.fixture(scope="module")
def H():
    from oracle import ref_harness
    yield ref_harness
    ref_harness.restore_idm_class_constants()


def _one(st):
    return st  # extract_state returns a 1-env SimState


def test_intersection_substeps_and_steps_live(H):
    from oracle import oracle as O
    net, table, cfg, routes = T.intersection_scene()
    orc = O.Oracle(cfg, table, routes)
    env = H.IntersectionEnv()
    rng = np.random.default_rng(777)
    n_sub = n_step = 0
    for seed in (901, 902, 903):
        env.reset(seed=seed)
        proxy = H.RecordingRng(env.np_random)
        env.np_random = proxy
        env.road.np_random = proxy
        done = False
        while not done:
            a = int(rng.integers(0, 3))
            # resynced sub-steps on a deep copy of the reference env would disturb its RNG: step the oracle
            # from the reference's pre-step state and compare after the full env.step instead ...
            before = H.extract_state(env, table, 32)
            proxy.log.clear()
            obs, reward, term, trunc, info = env.step(a)
            draw = H.draws_from_log(proxy.log)
            draws = (abi.SpawnDraw * 1)()
            if draw is not None:
                draws[0] = draw
            else:
                draws[0].u_spawn = 2.0
            oo, orr, ot, ou, _ = orc.step(before, np.array([a], np.int32), draws=draws)
            T.compare_states(before, H.extract_state(env, table, 32), T.TOL_STEP, f"seed {seed} step {n_step}")
            np.testing.assert_allclose(oo.reshape(obs.shape), obs, rtol=0, atol=2e-6)
            assert abs(float(orr[0]) - reward) <= 1e-6 and bool(ot[0]) == term and bool(ou[0]) == trunc
            done = term or trunc
            n_step += 1
    assert n_step >= 15
    # ... and resynced single sub-steps on a fresh episode (no spawn draws consumed inside a sub-step)
    env.reset(seed=950)
    for k in range(30):
        a = int(rng.integers(0, 3))
        before = H.extract_state(env, table, 32)
        H.ref_substep(env, a)
        orc.substep(before, np.array([a], np.int32))
        T.compare_states(before, H.extract_state(env, table, 32), T.TOL_SUBSTEP, f"sub-step {k}")
        n_sub += 1
    assert n_sub == 30


@pytest.mark.parametrize("n,density", [(12, 1.5), (50, 2.5)])
def test_highway_substeps_live(H, n, density):
    from oracle import oracle as O
    _, table, cfg, cfgd = T.highway_scene(n, density)
    orc = O.Oracle(cfg, table)
    env = H.SyntheticHighwayEnv(config={"vehicles_count": n, "vehicles_density": density})
    rng = np.random.default_rng(5)
    for seed in (31, 32):
        env.reset(seed=seed)
        a = 1
        for k in range(45):
            if k % 15 == 0:
                a = int(rng.integers(0, 5))
            before = H.extract_state(env, table, n)
            H.ref_substep(env, a)
            orc.substep(before, np.array([a], np.int32))
            T.compare_states(before, H.extract_state(env, table, n), T.TOL_SUBSTEP, f"seed {seed} sub-step {k}")
