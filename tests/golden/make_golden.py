"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(``/root/reference`` through ``oracle/ref_shim.py``).  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors; these files are what pins the CPU
oracle (tests/test_oracle_golden.py) and, through the same inputs, the CUDA path (tests/test_gpu_*.py).
Every array is stored as float64 / int32 exactly as the reference produced it.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H  # noqa: E402
from topotrafficrl_b200 import abi, scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

GRID_DENSE = {"observation": {"type": "OccupancyGrid", "vehicles_count": 15,
                              "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                              "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]},
                              "grid_size": [[-32, 32], [-32, 32]], "grid_step": [2, 2], "absolute": False}}
GRID_ROAD = {"observation": {"type": "OccupancyGrid", "features": ["presence", "vx", "vy", "on_road"],
                             "grid_size": [[-27.5, 27.5], [-27.5, 27.5]], "grid_step": [5, 5], "absolute": False}}


class Rec:
    def __init__(self):
        self.d = {}

    def add(self, **kw):
        for k, v in kw.items():
            self.d.setdefault(k, []).append(np.asarray(v))

    def add_state(self, prefix, st):
        self.add(**{prefix + "_vd": st.veh_d[:, 0], prefix + "_vi": st.veh_i[:, 0],
                    prefix + "_ei": st.env_i[:, 0], prefix + "_ed": st.env_d[:, 0]})
        if st.lin is not None:  # LinearVehicle parameters
            self.add(**{prefix + "_lin": st.lin[:, 0]})

    def save(self, name):
        path = os.path.join(OUT, name)
        np.savez_compressed(path, **{k: np.stack(v) for k, v in self.d.items()})
        print(name, {k: np.stack(v).shape for k, v in self.d.items()}, os.path.getsize(path) // 1024, "KiB")


def draw_array(d: abi.SpawnDraw):
    return np.array([d.u_spawn, d.entry, d.exit, d.n_pos, d.n_speed, d.delta] + [d.lin_u[k] for k in range(5)], dtype=np.float64)


# --------------------------------------------------------------------------------------------------
def intersection_substeps(seeds, vcap=32):
    net = scenes.make_intersection_network()
    table = net.to_table(scenes.intersection_exit_predicate)
    env = H.IntersectionEnv()
    rec = Rec()
    rng = np.random.default_rng(1234)
    for seed in seeds:
        env.reset(seed=seed)
        done = False
        while not done:
            a = int(rng.integers(0, 3))
            for _ in range(15):
                rec.add_state("before", H.extract_state(env, table, vcap))
                rec.add(action=a)
                H.ref_substep(env, a)
                rec.add_state("after", H.extract_state(env, table, vcap))
            env.time += 1
            done = env._is_terminated() or env._is_truncated()
            env._clear_vehicles()
            env._spawn_vehicle(spawn_probability=env.config["spawn_probability"])
    rec.save("intersection_substeps.npz")


def intersection_steps(seeds, overrides, name, vcap=32):
    net = scenes.make_intersection_network()
    table = net.to_table(scenes.intersection_exit_predicate)
    env = H.IntersectionEnv(config=overrides)
    rec = Rec()
    rng = np.random.default_rng(99)
    for seed in seeds:
        env.reset(seed=seed)
        proxy = H.RecordingRng(env.np_random)
        env.np_random = proxy
        env.road.np_random = proxy
        done = False
        while not done:
            a = int(rng.integers(0, 3))
            rec.add_state("before", H.extract_state(env, table, vcap))
            proxy.log.clear()
            obs, reward, term, trunc, info = env.step(a)
            rec.add(action=a, draw=draw_array(H.draws_from_log(proxy.log)), obs=obs.astype(np.float32),
                    reward=np.float64(reward), terminated=bool(term), truncated=bool(trunc),
                    speed=np.float64(info["speed"]), crashed=bool(info["crashed"]))
            rec.add_state("after", H.extract_state(env, table, vcap))
            done = term or trunc
    rec.save(name)


def intersection_episodes(name, overrides, seeds, vcap=32):
    """Whole seeded episodes of a reference-shipped config through the public API (reset(seed) then step until done):
    the (s, a, r, s', done) stream incl. the truncated last step, the info dict, the spawn draws and the row
    permutation np_random.shuffle applied to the observation -- plus the state after reset and after every step."""
    net = scenes.make_intersection_network()
    table = net.to_table(scenes.intersection_exit_predicate)
    env = H.IntersectionEnv(config=overrides)
    m = env.config["observation"]["vehicles_count"] - 1
    keys = ("collision_reward", "high_speed_reward", "arrived_reward", "on_road_reward")
    rec = Rec()
    rng = np.random.default_rng(2718)

    def info_row(info):
        return np.array([float(info["speed"]), float(info["crashed"])] + [float(info["rewards"][k]) for k in keys], np.float64)

    for seed in seeds:
        obs, info = env.reset(seed=seed)
        proxy = H.RecordingRng(env.np_random)
        env.np_random = proxy
        env.road.np_random = proxy
        rec.add(ep_seed=seed, ep_first=len(rec.d.get("action", [])), reset_obs=np.asarray(obs, np.float32), reset_info=info_row(info))
        rec.add_state("reset", H.extract_state(env, table, vcap))
        done = False
        while not done:
            a = int(rng.integers(0, 3))
            proxy.log.clear()
            proxy.perms.clear()
            obs, reward, term, trunc, info = env.step(a)
            perm = proxy.perms[0] if proxy.perms else np.arange(m)
            rec.add(action=a, draw=draw_array(H.draws_from_log(proxy.log)), perm=np.asarray(perm, np.int32), obs=np.asarray(obs, np.float32),
                    reward=np.float64(reward), terminated=bool(term), truncated=bool(trunc), info=info_row(info),
                    agents_rewards=np.array(info["agents_rewards"], np.float64), agents_terminated=np.array(info["agents_terminated"], bool))
            rec.add_state("after", H.extract_state(env, table, vcap))
            done = term or trunc
    rec.add(n_steps=len(rec.d["action"]))
    rec.save("intersection_ep_%s.npz" % name)


def intersection_reset(seeds, vcap=32):
    """Post-reset states + first observation for seeded resets (anchors the host-driven reset)."""
    net = scenes.make_intersection_network()
    table = net.to_table(scenes.intersection_exit_predicate)
    env = H.IntersectionEnv(config={"observation": dict(scenes.INTERSECTION_CONFIG["observation"], order="sorted")})
    rec = Rec()
    for seed in seeds:
        obs, _ = env.reset(seed=seed)
        rec.add_state("state", H.extract_state(env, table, vcap))
        rec.add(seed=seed, obs=obs.astype(np.float32))
    rec.save("intersection_reset.npz")


def highway(n_vehicles, density, seeds, n_substeps, name, obs_overrides=None, with_steps=True, with_substeps=True):
    net = scenes.make_highway_network(4)
    table = net.to_table()
    cfg = {"vehicles_count": n_vehicles, "vehicles_density": density}
    if obs_overrides:
        cfg.update(obs_overrides)
    env = H.SyntheticHighwayEnv(config=cfg)
    rec = Rec()
    rng = np.random.default_rng(4321)
    for seed in seeds:
        env.reset(seed=seed)
        a = 1
        for k in range(n_substeps):
            if k % 15 == 0:
                a = int(rng.integers(0, 5))
            rec.add_state("before", H.extract_state(env, table, n_vehicles))
            rec.add(action=a)
            H.ref_substep(env, a)
            rec.add_state("after", H.extract_state(env, table, n_vehicles))
    if with_substeps:
        rec.save(name + "_substeps.npz")
    if not with_steps:
        return
    rec = Rec()
    for seed in seeds:
        env.reset(seed=seed + 100)
        for _ in range(max(2, n_substeps // 15)):
            a = int(rng.integers(0, 5))
            rec.add_state("before", H.extract_state(env, table, n_vehicles))
            obs, reward, term, trunc, info = env.step(a)
            rec.add(action=a, obs=obs.astype(np.float32), reward=np.float64(reward), terminated=bool(term),
                    truncated=bool(trunc))
            rec.add_state("after", H.extract_state(env, table, n_vehicles))
            if term or trunc:
                break
    rec.save(name + "_steps.npz")


MULTI_AGENT = {"observation": {"type": "MultiAgentObservation",
                               "observation_config": {"type": "Kinematics", "vehicles_count": 15,
                                                      "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                                                      "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]},
                                                      "absolute": True, "order": "sorted"}},
               "initial_vehicle_count": 5, "controlled_vehicles": 4}
UTURN_KIN = {"observation": {"type": "Kinematics", "vehicles_count": 6, "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                             "absolute": False, "order": "sorted"}}


def multi_agent_steps(seeds, overrides, name, vcap=32):
    """MultiAgentIntersectionEnv (intersection_env.py:372-394; scripts/configs/IntersectionEnv/env_multi_agent.json with a
    sorted observation order): post-reset states, full steps with tuple actions, per-agent rewards / terminal flags."""
    net = scenes.make_intersection_network()
    table = net.to_table(scenes.intersection_exit_predicate)
    env = H.MultiAgentIntersectionEnv(config=overrides)
    K = env.config["controlled_vehicles"]
    rec = Rec()
    rng = np.random.default_rng(77)
    for seed in seeds:
        obs, _ = env.reset(seed=seed)
        rec.add_state("reset", H.extract_state(env, table, vcap))
        rec.add(reset_seed=seed, reset_obs=np.stack(obs).astype(np.float32))
        proxy = H.RecordingRng(env.np_random)
        env.np_random = proxy
        env.road.np_random = proxy
        done = False
        while not done:
            a = tuple(int(x) for x in rng.integers(0, 3, size=K))
            rec.add_state("before", H.extract_state(env, table, vcap))
            proxy.log.clear()
            obs, reward, term, trunc, info = env.step(a)
            rec.add(action=np.array(a, np.int32), draw=draw_array(H.draws_from_log(proxy.log)), obs=np.stack(obs).astype(np.float32),
                    reward=np.float64(reward), terminated=bool(term), truncated=bool(trunc),
                    agents_rewards=np.array(info["agents_rewards"], np.float64),
                    agents_terminated=np.array(info["agents_terminated"], bool))
            rec.add_state("after", H.extract_state(env, table, vcap))
            done = term or trunc
    rec.save(name)


def scripted_scene(env_cls, net, overrides, seeds, name, n_actions, vcap=16, substeps=True):
    """RoundaboutEnv / UTurnEnv: post-reset states per seed, per-sub-step and per-step snapshots under random actions."""
    H.restore_idm_class_constants()  # IntersectionEnv mutates the IDMVehicle class constants process-wide
    table = net.to_table()
    env = env_cls(config=overrides)
    rec, sub = Rec(), Rec()
    rng = np.random.default_rng(555)
    for seed in seeds:
        obs, _ = env.reset(seed=seed)
        rec.add_state("reset", H.extract_state(env, table, vcap))
        rec.add(reset_seed=seed, reset_obs=np.asarray(obs, np.float32))
        done = False
        k = 0
        while not done:
            a = int(rng.integers(0, n_actions))
            if substeps and k % 3 == 2:  # every third env-step sub-step by sub-step (same transition; finer resync for the parity tests)
                env.time += 1 / env.config["policy_frequency"]
                for _ in range(15):
                    sub.add_state("before", H.extract_state(env, table, vcap))
                    sub.add(action=a)
                    H.ref_substep(env, a)
                    sub.add_state("after", H.extract_state(env, table, vcap))
                done = env._is_terminated() or env._is_truncated()
            else:
                rec.add_state("before", H.extract_state(env, table, vcap))
                obs, reward, term, trunc, info = env.step(a)
                rec.add(action=a, obs=np.asarray(obs, np.float32), reward=np.float64(reward), terminated=bool(term), truncated=bool(trunc))
                rec.add_state("after", H.extract_state(env, table, vcap))
                done = term or trunc
            k += 1
    rec.save(name + "_steps.npz")
    if substeps:
        sub.save(name + "_substeps.npz")


def function_kats():
    """Function-level known answers: reference function outputs on random inputs."""
    from ttrl_env import utils as U
    from ttrl_env.road.lane import SineLane
    from ttrl_env.vehicle.behavior import IDMVehicle
    from ttrl_env.vehicle.controller import ControlledVehicle, MDPVehicle
    from ttrl_env.vehicle.kinematics import Vehicle

    rng = np.random.default_rng(2024)
    out = {}
    xs = np.concatenate([rng.uniform(-50, 50, 200), [0.0, np.pi, -np.pi, 3 * np.pi, -3 * np.pi, 1e-3, -1e-3, 0.01, -0.01]])
    out["wrap_in"] = xs
    out["wrap_out"] = np.array([U.wrap_to_pi(x) for x in xs])
    out["not_zero_out"] = np.array([U.not_zero(x) for x in xs])

    env = H.IntersectionEnv()
    env.reset(seed=0)
    lanes = env.road.network.lanes_list()
    pts = rng.uniform(-120, 120, (64, 2))
    hs = rng.uniform(-4, 4, 64)
    ss = rng.uniform(-10, 110, 64)
    rs = rng.uniform(-4, 4, 64)
    out["lane_pts"], out["lane_h"], out["lane_s"], out["lane_r"] = pts, hs, ss, rs
    out["lane_local"] = np.array([[l.local_coordinates(p) for p in pts] for l in lanes], dtype=np.float64)
    out["lane_position"] = np.array([[l.position(s, r) for s, r in zip(ss, rs)] for l in lanes], dtype=np.float64)
    out["lane_heading"] = np.array([[l.heading_at(s) for s in ss] for l in lanes], dtype=np.float64)
    out["lane_dwh"] = np.array([[l.distance_with_heading(p, h) for p, h in zip(pts, hs)] for l in lanes], dtype=np.float64)
    net = env.road.network
    keys = list(net.lanes_dict().keys())
    out["closest"] = np.array([keys.index(net.get_closest_lane_index(p, h)) for p, h in zip(pts, hs)], dtype=np.int32)

    # sine lane (RoundaboutEnv geometry class; lane.py:236-308)
    sine = SineLane([0.0, 0.0], [80.0, 10.0], 3.25, 2 * np.pi / 40, np.pi / 2, speed_limit=15)
    out["sine_params"] = np.array([0.0, 0.0, 80.0, 10.0, 3.25, 2 * np.pi / 40, np.pi / 2])
    out["sine_local"] = np.array([sine.local_coordinates(p) for p in pts], dtype=np.float64)
    out["sine_position"] = np.array([sine.position(s, r) for s, r in zip(ss, rs)], dtype=np.float64)
    out["sine_heading"] = np.array([sine.heading_at(s) for s in ss], dtype=np.float64)

    # steering_control / speed_to_index on the intersection network
    road = env.road
    st_in, st_out = [], []
    for _ in range(128):
        li = int(rng.integers(0, len(lanes)))
        s = rng.uniform(0, lanes[li].length)
        p = lanes[li].position(s, rng.uniform(-3, 3))
        h = lanes[li].heading_at(s) + rng.uniform(-0.6, 0.6)
        sp = rng.choice([rng.uniform(-2, 12), rng.uniform(-0.02, 0.02)])
        v = ControlledVehicle(road, p, h, sp)
        st_in.append([p[0], p[1], h, sp, li])
        st_out.append(v.steering_control(keys[li]))
    out["steer_in"], out["steer_out"] = np.array(st_in), np.array(st_out)
    sp_in = np.concatenate([rng.uniform(-3, 12, 64), [0, 2.25, 4.5, 6.75, 9.0, 1.125, 3.375]])
    mdp = MDPVehicle(road, [2.0, 50.0], -np.pi / 2, 5.0, target_speeds=[0, 4.5, 9])
    out["s2i_in"], out["s2i_out"] = sp_in, np.array([int(mdp.speed_to_index(s)) for s in sp_in], dtype=np.int32)

    # IDM acceleration with the intersection's class constants (set by reset above)
    idm_in, idm_out = [], []
    lane0 = lanes[0]
    for _ in range(128):
        s0 = rng.uniform(5, 60)
        gap = rng.choice([rng.uniform(0.001, 0.02), rng.uniform(1, 60)])
        e = IDMVehicle(road, lane0.position(s0, rng.uniform(-0.5, 0.5)), lane0.heading_at(s0) + rng.uniform(-0.1, 0.1), rng.uniform(-1, 12))
        f = IDMVehicle(road, lane0.position(s0 + gap, rng.uniform(-0.5, 0.5)), lane0.heading_at(s0) + rng.uniform(-0.1, 0.1), rng.uniform(0, 12))
        e.target_speed = rng.choice([0.0, rng.uniform(0, 14)])
        e.DELTA = rng.uniform(3.5, 4.5)
        use_front = rng.uniform() < 0.8
        idm_in.append([e.DELTA, e.position[0], e.position[1], e.heading, e.speed, e.target_speed, keys.index(e.lane_index),
                       f.position[0], f.position[1], f.heading, f.speed, f.target_speed, keys.index(f.lane_index), float(use_front)])
        idm_out.append(e.acceleration(e, f if use_front else None))
    out["idm_in"], out["idm_out"] = np.array(idm_in), np.array(idm_out, dtype=np.float64)

    # collision SAT and regulation rectangles
    pa, po = [], []
    for _ in range(256):
        a = Vehicle(None, rng.uniform(-3, 3, 2), rng.uniform(-3.2, 3.2), rng.uniform(0, 15))
        b = Vehicle(None, a.position + rng.uniform(-6, 6, 2), rng.uniform(-3.2, 3.2), rng.uniform(0, 15))
        dt = 1 / 15
        inter, will, trans = U.are_polygons_intersecting(a.polygon(), b.polygon(), a.velocity * dt, b.velocity * dt)
        pa.append([a.position[0], a.position[1], a.heading, a.speed, b.position[0], b.position[1], b.heading, b.speed])
        po.append([float(inter), float(will), 0.0 if trans is None else trans[0], 0.0 if trans is None else trans[1]])
    out["sat_in"], out["sat_out"] = np.array(pa), np.array(po)
    ra, ro = [], []
    for _ in range(256):
        c1 = rng.uniform(-3, 3, 2)
        c2 = c1 + rng.uniform(-7, 7, 2)
        r1 = (c1, 7.5, 1.8, rng.uniform(-3.2, 3.2))
        r2 = (c2, 7.5, 1.8, rng.uniform(-3.2, 3.2))
        ra.append([c1[0], c1[1], 7.5, 1.8, r1[3], c2[0], c2[1], 7.5, 1.8, r2[3]])
        ro.append(bool(U.rotated_rectangles_intersect(r1, r2)))
    out["rect_in"], out["rect_out"] = np.array(ra), np.array(ro)
    path = os.path.join(OUT, "kat_functions.npz")
    np.savez_compressed(path, **out)
    print("kat_functions.npz", os.path.getsize(path) // 1024, "KiB")


def qnet_vectors():
    """Q-network forward known answers from the reference torch modules (models.py:431-441), CPU fp32."""
    import torch
    from ttrl_agent.agents.common.models import model_factory

    torch.manual_seed(0)
    rng = np.random.default_rng(7)
    obs = rng.uniform(-1, 1, (64, 15, 7)).astype(np.float32)
    obs[:, :, 0] = (rng.uniform(size=(64, 15)) < 0.7).astype(np.float32)
    obs[:, 0, 0] = 1.0
    obs[5, 1:, 0] = 0.0   # only the ego present
    obs[6, :, 0] = 0.0    # nothing present (all masked)
    out = {"obs": obs}
    configs = {
        "mlp": {"type": "MultiLayerPerceptron", "layers": [128, 128], "in": 105, "out": 3},
        "ego1h": {"type": "EgoAttentionNetwork", "in": 105, "out": 3,
                  "embedding_layer": {"type": "MultiLayerPerceptron", "layers": [64, 64], "reshape": False, "in": 7},
                  "others_embedding_layer": {"type": "MultiLayerPerceptron", "layers": [64, 64], "reshape": False, "in": 7},
                  "self_attention_layer": None,
                  "attention_layer": {"type": "EgoAttention", "feature_size": 64, "heads": 1},
                  "output_layer": {"type": "MultiLayerPerceptron", "layers": [64, 64], "reshape": False},
                  "presence_feature_idx": 0},
        "dueling": {"type": "DuelingNetwork", "in": 105, "out": 3,
                    "base_module": {"type": "MultiLayerPerceptron", "layers": [64, 64]},
                    "value": {"type": "MultiLayerPerceptron", "layers": []},
                    "advantage": {"type": "MultiLayerPerceptron", "layers": []}},
    }
    import copy
    configs["ego2h"] = copy.deepcopy(configs["ego1h"])
    configs["ego2h"]["attention_layer"]["heads"] = 2
    for name, cfg in configs.items():
        net = model_factory(cfg)
        net.eval()
        with torch.no_grad():
            q = net(torch.tensor(obs)).numpy()
        out[name + "_q"] = q.astype(np.float32)
        for k, v in net.state_dict().items():
            out[f"{name}/{k}"] = v.numpy().astype(np.float32)
    path = os.path.join(OUT, "qnet_vectors.npz")
    np.savez_compressed(path, **out)
    print("qnet_vectors.npz", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["kat", "int_sub", "int_steps", "int_reset", "hw", "qnet", "multi", "scripted", "ttc", "episodes"]
    if "kat" in which:
        function_kats()
    if "int_sub" in which:
        intersection_substeps(seeds=range(4))
    if "int_steps" in which:
        intersection_steps(range(100, 112), None, "intersection_steps_kin.npz")
        intersection_steps(range(200, 204), GRID_DENSE, "intersection_steps_grid_dense.npz")
        intersection_steps(range(300, 304), GRID_ROAD, "intersection_steps_grid_road.npz")
    if "episodes" in which or "linear" in which:
        from tests.common import EPISODE_CONFIGS
        for name, (over, seeds) in EPISODE_CONFIGS.items():
            if "episodes" in which or name == "linear":
                intersection_episodes(name, over, seeds)
    if "linear" in which:  # RoundaboutEnv/env.json and env_route_0.json as shipped: LinearVehicle traffic
        from tests.common import LINEAR
        scripted_scene(H.RoundaboutEnv, scenes.make_roundabout_network(), LINEAR, range(800, 806), "roundabout_linear", 5)
        scripted_scene(H.RoundaboutEnv, scenes.make_roundabout_network(), dict(LINEAR, incoming_vehicle_destination=0), range(810, 813),
                       "roundabout_linear_route0", 5, substeps=False)
    if "int_reset" in which:
        intersection_reset(range(8))
    if "hw" in which:
        highway(8, 2.0, range(2), 60, "highway_n8")
        highway(50, 2.0, range(3), 90, "highway_n50")
        highway(200, 4.0, range(1), 20, "highway_n200", with_steps=False)
        highway(40, 4.0, range(2), 30, "highway_grid_n40",
                obs_overrides={"observation": {"type": "OccupancyGrid", "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                                               "grid_size": [[-32, 32], [-32, 32]], "grid_step": [2, 2], "absolute": False}})
    if "qnet" in which:
        qnet_vectors()
    if "ttc" in which:  # TimeToCollisionObservation (observation.py:114-151): UTurnEnv's default config, and a 4-lane highway
        scripted_scene(H.UTurnEnv, scenes.make_uturn_network(), None, range(700, 708), "uturn_ttc", 5, substeps=False)
        highway(30, 2.0, range(3), 150, "highway_ttc_n30", with_substeps=False,
                obs_overrides={"observation": {"type": "TimeToCollision", "horizon": 10}})
    if "multi" in which:
        multi_agent_steps(range(400, 406), MULTI_AGENT, "multiagent_steps.npz")
    if "scripted" in which:
        scripted_scene(H.RoundaboutEnv, scenes.make_roundabout_network(), None, range(500, 508), "roundabout", 5)
        scripted_scene(H.UTurnEnv, scenes.make_uturn_network(), UTURN_KIN, range(600, 606), "uturn", 5)
