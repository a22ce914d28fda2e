#!/usr/bin/env python
"""bench.py -- throughput of the batched env step on B200 (driver contract: see README / DESIGN.md section 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload highway50|...]

One "step" = one env.step() of EVERY env of the workload (15 simulation sub-steps + observation + reward +
termination [+ clear/spawn] [+ autoreset]).  At N=1 the workload is BASELINE.json configs[1]: 4096 highway envs x
50 vehicles, Kinematics observation, random actions.  N>1 (torchrun, one rank per GPU): every rank owns its own
4096 envs (global env ids rank*E .. rank*E+E-1), no collective on the step path ("weak" scaling); only the
episode statistics are all-reduced (NCCL) after the timed region.

Numbers on the JSON line
  value      vehicle-steps/s, device-resident: actions already in HBM, obs/reward/flags written to HBM,
             CUDA events around each step (L2 flushed between steps), max over ranks.
  e2e        the same metric through the host-buffer C-ABI call (ttrl_sim_step_pinned / Sim.step_host): actions copied
             from host memory to the GPU, obs/reward/flags copied back to (page-locked) host memory every step and
             returned as numpy arrays, synchronous, wall clock around the call.
  roofline   k_step against the measured HBM copy bandwidth (MEASURED_PEAKS.json), algorithmic bytes per
             env-step from SURVEY.md section 8d / DESIGN.md section 6.
  cpu_baseline  the CPU oracle (C restatement of the reference algorithm, oracle/) on the host cores of this
             box, bounded sample of the same workload.  Rank 0, N=1 only.

The default run (headline workload, N = 1) adds an `also` block -- BASELINE configs[2] (dense200), configs[3] without and with the
DQN Q-network rollout (intersection, intersection_qnet), each >= 50 timed steps with its own roofline / e2e / clocks / CPU baseline --
and a 200-step `self_check` of the headline.  `--scaling strong` splits a fixed total (32 768 highway / 1 048 576 intersection
envs) over the GPUs instead of giving every GPU the workload's envs.  A workload whose slot capacity rejected a spawn fails the run.

`--impl reference` times the CPU oracle port alone (the Python reference itself cannot travel to the GPU box and
runs at ~1.8e3 vehicle-steps/s/core; the C port is the faster, fairer CPU arm).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene, num_envs per GPU, vehicles, scene overrides, algorithmic bytes per env-step)
    "highway50": dict(scene="highway", E=4096, n=50, over={"vehicles_count": 50, "vehicles_density": 2.0},
                      label="configs[1]: highway 4096 envs x 50 vehicles, Kinematics obs, random actions",
                      bytes_per_env_step=50 * 128 + 4 + 420 + 8, n_actions=5),
    "dense200": dict(scene="highway", E=1024, n=200,
                     over={"vehicles_count": 200, "vehicles_density": 4.0,
                           "observation": {"type": "OccupancyGrid", "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                                           "grid_size": [[-32, 32], [-32, 32]], "grid_step": [2, 2], "absolute": False}},
                     label="configs[2]: dense highway 1024 envs x 200 vehicles, OccupancyGrid 7x32x32 obs",
                     bytes_per_env_step=200 * 128 + 4 + 28672 + 8, n_actions=5),
    "intersection": dict(scene="intersection", E=8192, n=24, over=None,
                         label="configs[3]: intersection 8192 envs, regulated road, spawn, Kinematics obs",
                         bytes_per_env_step=24 * 128 + 4 + 420 + 8, n_actions=3),
    # BASELINE configs[3] proper: the DQN Q-network rollout (agent.act: forward + epsilon-greedy) in the loop
    "intersection_qnet": dict(scene="intersection", E=8192, n=24, over=None, qnet="ego_attention_2h",
                              label="configs[3]: intersection 8192 envs + DQN ego-attention (2 heads) Q-net rollout in the loop",
                              bytes_per_env_step=24 * 128 + 4 + 420 + 8 + 420 + 4, n_actions=3),
    "intersection_qnet_mlp": dict(scene="intersection", E=8192, n=24, over=None, qnet="mlp",
                                  label="configs[3]: intersection 8192 envs + DQN MLP [128,128] Q-net rollout in the loop",
                                  bytes_per_env_step=24 * 128 + 4 + 420 + 8 + 420 + 4, n_actions=3),
    # the whole DQN training iteration (CUDA Q-net act -> env step -> replay push -> double-DQN
    # update with torch autograd -> rollout weights refreshed on the device), one optimiser step of batch 64 per vector step
    "train_intersection": dict(scene="intersection", E=8192, n=24, over=None, train="mlp",
                               label="training: intersection 8192 envs, DQN MLP [128,128] (baseline.json), act + step + record per iteration",
                               bytes_per_env_step=24 * 128 + 4 + 420 + 8 + 420 + 4, n_actions=3),
    # SURVEY.md section 8 rows A29 / N3 (not BASELINE configs; parity cases with a throughput line)
    # BASELINE configs[4]: one GPU's shard of 1M envs (1 048 576 / 8), full step + observation + policy
    "c5_shard": dict(scene="intersection", E=131072, n=24, over=None, qnet="ego_attention_2h",
                     label="configs[4] shard: 131072 of 1 048 576 intersection envs (1M / 8 GPUs) + DQN ego-attention Q-net rollout in the loop",
                     bytes_per_env_step=24 * 128 + 4 + 420 + 8 + 420 + 4, n_actions=3),
    "multiagent": dict(scene="intersection", E=8192, n=24, agents=4,
                       over={"controlled_vehicles": 4, "initial_vehicle_count": 5,
                             "action": {"type": "MultiAgentAction", "action_config": {"type": "DiscreteMetaAction", "lateral": False, "longitudinal": True}},
                             "observation": {"type": "MultiAgentObservation", "observation_config": {
                                 "type": "Kinematics", "vehicles_count": 15, "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                                 "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]}, "absolute": True}}},
                       label="multi-agent intersection (env_multi_agent.json): 8192 envs x 4 controlled vehicles, 4 Kinematics obs per env",
                       bytes_per_env_step=24 * 128 + 4 * 4 + 4 * 420 + 8 + 4 * 5, n_actions=3),
    "roundabout": dict(scene="roundabout", E=8192, n=16, over=None, scripted=True,
                       label="roundabout-v0: 8192 envs x 5 vehicles, 32-lane network (arcs + sine lanes), Kinematics 5x5 obs",
                       bytes_per_env_step=5 * 128 + 4 + 100 + 8, n_actions=5),
    "uturn": dict(scene="u-turn", E=8192, n=16, scripted=True,
                  over={"observation": {"type": "Kinematics", "vehicles_count": 6, "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"]}},
                  label="u-turn-v0: 8192 envs x 7 vehicles, Kinematics 6x7 obs", bytes_per_env_step=7 * 128 + 4 + 168 + 8, n_actions=5),
}

QNET_CONFIGS = {  # scripts/configs/IntersectionEnv/agents/DQNAgent/{ego_attention_2h,baseline}.json of the reference
    "ego_attention_2h": {"type": "EgoAttentionNetwork", "embedding_layer": {"layers": [64, 64]}, "others_embedding_layer": {"layers": [64, 64]},
                         "self_attention_layer": None, "attention_layer": {"feature_size": 64, "heads": 2}, "output_layer": {"layers": [64, 64]}},
    "mlp": {"type": "MultiLayerPerceptron", "layers": [128, 128]},
}


def random_state_dict(kind: str, n_actions: int, seed: int = 0):
    """Random-init weights of the reference architectures (torch.nn.Linear default init: U(-1/sqrt(fan_in), +))."""
    rng = np.random.default_rng(seed)
    sd = {}

    def lin(name, fan_in, fan_out, bias=True):
        b = 1 / np.sqrt(fan_in)
        sd[name + ".weight"] = rng.uniform(-b, b, size=(fan_out, fan_in)).astype(np.float32)
        if bias:
            sd[name + ".bias"] = rng.uniform(-b, b, size=fan_out).astype(np.float32)

    if kind == "mlp":
        lin("layers.0", 105, 128); lin("layers.1", 128, 128); lin("predict", 128, n_actions)
    else:
        for emb in ("ego_embedding", "others_embedding"):
            lin(emb + ".layers.0", 7, 64); lin(emb + ".layers.1", 64, 64)
        for name in ("key_all", "value_all", "query_ego", "attention_combine"):
            lin("attention_layer." + name, 64, 64, bias=False)
        lin("output_layer.layers.0", 64, 64); lin("output_layer.layers.1", 64, 64); lin("output_layer.predict", 64, n_actions)
    return sd


def _dram_traffic(workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/dram_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)[workload]["bytes"])
    except (OSError, KeyError, ValueError):
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 20 ms from a thread (pynvml);
    falls back to `nvidia-smi -lms 100` (B200_PROFILING.md clocks line) when NVML is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.sm, self.mx, self.reasons = [], [], set()

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        names = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(self.handle))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def build_scene(w):
    from topotrafficrl_b200 import scenes
    if w["scene"] == "highway":
        cfgd = scenes.merged_config(scenes.HIGHWAY_CONFIG, w["over"])
        net = scenes.make_highway_network(int(cfgd["lanes_count"]), float(cfgd["road_length"]), float(cfgd["speed_limit"]))
        table = net.to_table()
        cfg = scenes.build_config(table, cfgd, "highway", ego_lanes_count=int(cfgd["lanes_count"]))
        return cfgd, net, table, cfg, None
    cfgd = scenes.merged_config(scenes.INTERSECTION_CONFIG, w["over"])
    net = scenes.make_intersection_network()
    table = net.to_table(scenes.intersection_exit_predicate)
    cfg = scenes.build_config(table, cfgd, "intersection")
    return cfgd, net, table, cfg, scenes.intersection_spawn_routes(net, table)


SPAWN_DRAW_DTYPE = np.dtype([("u_spawn", "f8"), ("entry", "i4"), ("exit", "i4"), ("n_pos", "f8"), ("n_speed", "f8"), ("delta", "f8"), ("lin_u", "f8", (5,))])


class _OracleResetBackend:
    """`reset.reset_intersection` driven by the CPU oracle (spawn attempts + warm-up sub-steps on a host SimState): the CPU
    arm generates its own initial states with the reference's procedure, without touching the GPU."""

    def __init__(self, orc, num_envs: int, vcap: int):
        from topotrafficrl_b200.state import SimState
        self.orc, self.num_envs = orc, num_envs
        self.st = SimState.zeros(num_envs, vcap)

    def spawn(self, draws, longitudinal, position_deviation, speed_deviation, spawn_probability, go_straight):
        return self.orc.spawn(self.st, draws, longitudinal, position_deviation, speed_deviation, spawn_probability, go_straight)

    def substep_none(self):
        self.orc.substep(self.st, None)

    def get_state(self):
        return self.st.copy()

    def set_state(self, st):
        self.st = st.copy()


def cpu_oracle_throughput(w, seconds: float, seed: int = 0):
    """CPU arm: the oracle port on all host cores, bounded sample of the workload.  Returns (veh-steps/s, env-steps/s, cores, sample)."""
    import ctypes as C

    from oracle import oracle as O
    from topotrafficrl_b200 import scenes
    from topotrafficrl_b200._gym import np_random as gym_np_random
    from topotrafficrl_b200.reset import reset_intersection
    if w["scene"] not in ("highway", "intersection") or w.get("agents"):
        raise SystemExit("the CPU arm is implemented for the highway and the single-agent intersection workloads")
    cfgd, net, table, cfg, routes = build_scene(w)
    cores = os.cpu_count() or 1
    E = max(cores * 8, 64)
    orc = O.Oracle(cfg, table, routes, threads=cores)
    if w["scene"] == "highway":
        st = scenes.make_highway_state(E, cfgd, seed=seed)
    else:  # IntersectionEnv._make_vehicles for E seeded envs (numpy PCG64 streams like gymnasium's), capacity as on the device
        backend = _OracleResetBackend(orc, E, w["n"])
        st = reset_intersection(backend, [gym_np_random(1000 * seed + e)[0] for e in range(E)], net, table, cfgd, cfg)
    orc.set_reset_pool(st.copy())
    orc.set_autoreset(True)
    rng = np.random.default_rng(seed)
    stats = np.zeros(8)

    def draws():
        if w["scene"] != "intersection":
            return None
        d = np.zeros(E, SPAWN_DRAW_DTYPE)  # the draws of _spawn_vehicle (intersection_env.py:328-346, behavior.py:66-69)
        d["u_spawn"] = rng.uniform(size=E)
        d["entry"] = rng.integers(0, 4, size=E)
        d["exit"] = (d["entry"] + rng.integers(1, 4, size=E)) % 4
        d["n_pos"], d["n_speed"], d["delta"] = rng.normal(size=E), rng.normal(size=E), rng.uniform(3.5, 4.5, size=E)
        draws.keep = d
        return C.cast(d.ctypes.data, C.c_void_p)

    orc.step(st, rng.integers(0, w["n_actions"], size=E).astype(np.int32), draws(), stats=stats)  # warm-up
    stats[:] = 0
    t0 = time.perf_counter()
    steps = 0
    while True:
        orc.step(st, rng.integers(0, w["n_actions"], size=E).astype(np.int32), draws(), stats=stats)
        steps += 1
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    return stats[6] / dt, stats[7] / dt, cores, f"{E} envs x {stats[6] / max(stats[7], 1) / 15:.1f} vehicles (mean) x {steps} env-steps ({dt:.1f} s) of the same scene generator"


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a "step" of this arm is a bounded sample of the workload on the host cores: 2 s each, less when many steps are asked for, so
    # that the whole run stays within ~2.5 minutes
    per_step_seconds = min(2.0, max(0.25, 150.0 / max(args.warmup + args.steps, 1)))
    vals, evals = [], []
    cores = sample = None
    for k in range(args.warmup + args.steps):
        v, ev, cores, sample = cpu_oracle_throughput(w, per_step_seconds, seed=k)
        if k >= args.warmup:
            vals.append(v); evals.append(ev)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": "vehicle_steps_per_sec", "value": value, "unit": "vehicle-steps/s",
            "env_steps_per_sec": float(np.mean(evals)), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step_seconds * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": w["label"], "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "vehicle-steps/s", "cores": cores, "kind": "port",
                             "sample": "each step: " + sample},
            "e2e": {"value": value, "unit": "vehicle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _setup_dist():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the simulator has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return rank, local_rank, world, dev


STRONG_TOTAL_ENVS = {"highway": 32768, "intersection": 1048576}  # fixed total E split over the GPUs


def measure(args, wname, K, W, with_cpu, self_check_steps=0):
    """One workload: device-resident arm, end-to-end (host buffer) arm, roofline, clocks -> the JSON line's dict (rank 0) or None."""
    import torch
    import torch.distributed as dist

    from topotrafficrl_b200.vector_env import TTRLVectorEnv
    w = WORKLOADS[wname]
    rank, local_rank, world, dev = _setup_dist()
    cfgd, net, table, cfg, routes = build_scene(w)
    if args.envs:
        E = args.envs
    elif args.scaling == "strong":
        E = STRONG_TOTAL_ENVS["highway" if w["scene"] == "highway" else "intersection"] // world
    else:
        E = w["E"]
    first_env = rank * E
    default_vcap = not args.vcap
    venv = TTRLVectorEnv(E, scene=w["scene"], config=w["over"], device=local_rank, seed=0, first_env=first_env,
                         vcap=(args.vcap or w["n"]), reset_mode=args.reset_mode, async_reset=not args.sync_reset)
    venv.reset()
    sim = venv.sim
    A = sim.num_agents  # controlled vehicles per env: actions are [E, A]

    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    n_extra = max(0, int(self_check_steps))
    actions = torch.randint(0, w["n_actions"], (W + K + n_extra, E * A), dtype=torch.int32, device=dev, generator=gen)
    obs = torch.zeros(E * sim.obs_size, dtype=torch.float32, device=dev)
    rew = torch.zeros(E, dtype=torch.float32, device=dev)
    term = torch.zeros(E, dtype=torch.uint8, device=dev)
    trunc = torch.zeros(E, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = int(torch.cuda.current_stream(dev).cuda_stream)

    qnet = None
    if w.get("qnet"):
        from topotrafficrl_b200.agent import QNetRollout
        qnet = QNetRollout(QNET_CONFIGS[w["qnet"]], random_state_dict(w["qnet"], w["n_actions"]), (15, 7), w["n_actions"], device=local_rank,
                           exploration={"method": "EpsilonGreedy", "temperature": 0.05, "final_temperature": 0.05, "tau": 15000}, seed=7 + rank,
                           mode=args.qnet_mode)
        sim.observe_ptr(obs.data_ptr(), stream)
    obs3 = obs.view(E, -1)
    trainer = None
    if w.get("train"):
        from topotrafficrl_b200.trainer import BatchedDQNAgent
        trainer = BatchedDQNAgent(venv, {"model": QNET_CONFIGS[w["train"]], "gamma": 0.95, "batch_size": 64, "memory_capacity": 15000,
                                         "target_update": 512, "exploration": {"method": "EpsilonGreedy", "tau": 15000, "temperature": 1.0,
                                                                               "final_temperature": 0.05}}, seed=rank, rollout_mode=args.qnet_mode, cuda_graph=(world == 1))
        sim.observe_ptr(obs.data_ptr(), stream)
        obs_t = obs.view((E,) + venv.obs_shape)
        final_obs = torch.zeros_like(obs_t)
        sim.set_info_outputs_ptr(None, final_obs.data_ptr())

    def step(k):
        a = actions[k]
        if trainer is not None:  # one full training iteration (trainer.BatchedDQNAgent: act / record)
            prev = obs_t.clone()
            a = trainer.act(prev)
            sim.step_ptr(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), term.data_ptr(), trunc.data_ptr(), stream)
            trainer.record(prev, a, rew, obs_t, term, trunc, {"final_observation": final_obs, "_final_observation": (term | trunc)})
            return
        if qnet is not None:  # agent.act on the observation the previous step left in HBM (no host round trip)
            a = qnet.act(obs3)
        sim.step_ptr(a.data_ptr(), obs.data_ptr(), rew.data_ptr(), term.data_ptr(), trunc.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    counted = qnet if qnet is not None else (trainer.rollout if trainer is not None else None)  # OUR kernels only (not torch's)

    def our_launches():
        return sim.launch_count + (counted._L.ttrl_qnet_launch_count(counted._h) if counted is not None else 0)

    # ---- device-resident arm -------------------------------------------------------------------------
    for k in range(W):
        step(k)
    sim.stats(reset=True)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = our_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.fill_(k & 0xFF)  # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        step(W + k)
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = our_launches() - launches0
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms))
    s = sim.stats(reset=True)  # one reduction kernel, outside the timed region
    veh_steps, env_steps = s.vehicle_steps, s.env_steps
    capacity_rejects, sync_resets = s.spawn_capacity_rejects, s.sync_resets
    red = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([veh_steps, env_steps, s.episodes, s.crashes, s.total_return, capacity_rejects, sync_resets], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)  # episode statistics: the only collective of the workload
    total_ms_max = float(red.item())
    veh_all, env_all, episodes, crashes, ret, rejects_all, sync_resets_all = (float(x) for x in tot.tolist())

    # ---- longer self-check of the same arm (the contract's K can be a 20 ms region): not part of `value` -------------
    self_check = None
    if n_extra:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for k in range(n_extra):
            step(W + K + k)
        e1.record()
        barrier()
        s3 = sim.stats(reset=True)
        sc_ms = e0.elapsed_time(e1)
        self_check = {"steps": n_extra, "ms_per_step": sc_ms / n_extra, "value": s3.vehicle_steps / (sc_ms * 1e-3),
                      "note": "back-to-back steps, no L2 flush, one event pair around all of them (this rank)"}

    # ---- end-to-end arm: host buffers through the C ABI ---------------------------------------------
    acts_host = actions.cpu().numpy()
    for k in range(min(W, 3)):
        sim.step_host(acts_host[k], copy=False)
    sim.stats(reset=True)
    barrier()
    t0 = time.perf_counter()
    o = obs3.cpu().numpy()
    for k in range(K):
        a_host = acts_host[W + k]
        if qnet is not None:  # the reference's agent.act boundary: numpy observation in, numpy actions out
            a_host = qnet.act(torch.from_numpy(o).to(dev, non_blocking=True).view(E, -1)).cpu().numpy()
        o, r, t, u = sim.step_host(a_host, copy=False)  # zero-copy views of the pinned staging buffers
    barrier()
    e2e_s = time.perf_counter() - t0
    s2 = sim.stats(reset=True)
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    e2e_v = torch.tensor([s2.vehicle_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_v, op=dist.ReduceOp.SUM)
    h2d = E * A * 4 + (E * sim.obs_size * 4 if qnet is not None else 0)
    d2h = E * (sim.obs_size * 4 + 4 + 2) + E * A * 5  # + per-agent rewards / terminal flags

    line = None
    if rank == 0:
        hbm_peak, peak_kind = _peaks()
        kernel_ms = statistics.mean(ms)
        alg_bytes = w["bytes_per_env_step"] * E
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": "vehicle_steps_per_sec", "value": veh_all / (total_ms_max * 1e-3), "unit": "vehicle-steps/s",
            "env_steps_per_sec": env_all / (total_ms_max * 1e-3),
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms_max / K, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["label"], "envs_per_gpu": E, "vehicles_per_env": w["n"], "sub_steps_per_step": 15,
                       "autoreset": ("pool of host-generated initial states" if args.reset_mode == "host" else "device-side fresh reset"), "l2": "flushed between timed steps (256 MB fill)",
                       "target": "1e8 vehicle-steps/s per B200 (BASELINE.json north_star)"},
            "e2e": {"value": float(e2e_v.item()) / float(e2e_t.item()), "unit": "vehicle-steps/s",
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": float(e2e_t.item()) / K * 1e3},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": _dram_traffic(wname) if E == w["E"] else None,
                         "traffic_source": "profiles/dram_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel, per launch",
                         "kernel": "k_step", "kernel_ms": kernel_ms, "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "clocks": clocks,
            "episode_stats": {"episodes": episodes, "crashes": crashes, "mean_return": ret / episodes if episodes else None,
                              "spawn_capacity_rejects": rejects_all, "sync_resets": sync_resets_all},
            "wall_s_timed_region": t_wall,
        }
        if self_check is not None:
            line["self_check"] = self_check
        if with_cpu and w["scene"] in ("highway", "intersection") and not w.get("agents"):
            v, evs, cores, sample = cpu_oracle_throughput(w, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "vehicle-steps/s", "cores": cores, "kind": "port", "sample": sample,
                                    "env_steps_per_sec": evs}
    if counted is not None:
        counted.close()
    venv.close()
    del flush
    torch.cuda.empty_cache()
    if default_vcap and rejects_all > 0:
        # the reference's vehicle list has no capacity: a rejected spawn makes the episodes differ from the reference's
        raise SystemExit(f"bench.py: {int(rejects_all)} spawns were rejected for lack of vehicle slots at the default capacity of workload {wname!r}")
    return line


ALSO_WORKLOADS = ("dense200", "intersection", "intersection_qnet")  # BASELINE configs[2], [3] (step only) and [3] proper


def run_ours(args, wname):
    import torch.distributed as dist
    rank, local_rank, world, dev = _setup_dist()
    headline = wname == "highway50" and not args.envs and args.scaling == "weak"
    line = measure(args, wname, args.steps, args.warmup, with_cpu=(world == 1 and not args.no_cpu_baseline),
                   self_check_steps=(200 if headline and world == 1 else 0))
    if headline and world == 1 and not args.no_also:
        # every other BASELINE config on the same box, same contract (>= 50 timed steps each), next to the headline fields
        also = {}
        for name in ALSO_WORKLOADS:
            sub = measure(args, name, max(50, min(args.steps, 100)), max(args.warmup, 5), with_cpu=(not args.no_cpu_baseline and name != "intersection_qnet"))
            also[name] = {k: sub[k] for k in ("value", "unit", "env_steps_per_sec", "steps", "warmup", "ms_per_step", "config", "e2e", "gpu_launches",
                                               "roofline", "clocks", "episode_stats") if k in sub}
            if "cpu_baseline" in sub:
                also[name]["cpu_baseline"] = sub["cpu_baseline"]
        line["also"] = also
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="highway50", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--vcap", type=int, default=0, help="vehicle slots per env (intersection workloads; default 24)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="headline run only: skip the dense200 / intersection / intersection_qnet block")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's envs on EVERY GPU; strong: a fixed total (32768 highway / 1048576 intersection envs) split over the GPUs")
    ap.add_argument("--sync-reset", action="store_true", help="device resets right after the step that finished the env (no side-stream regeneration)")
    ap.add_argument("--reset-mode", default="device", choices=["device", "host"],
                    help="device = fresh episodes generated on the GPU at every autoreset (the reference's _make_vehicles incl. its 45 warm-up sub-steps); host = replay a pool of host-generated initial states")
    ap.add_argument("--qnet-mode", default="tensor", choices=["fp32", "tensor"],
                    help="Q-net arithmetic for the *_qnet* workloads: tcgen05 tensor cores (BF16x3, default) or fp32 CUDA cores")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload])
    else:
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()
