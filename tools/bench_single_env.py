#!/usr/bin/env python
"""BASELINE config[0]: ONE env (the reference's scripts/example.py default: intersection-v0, env.json) stepped through the
gymnasium-shaped front end (numpy RNG stream of the reference, host-buffer step, info dict) -- per-step latency on the GPU, and
one DQN act per step.  The reference takes ~66 ms per step on one host core."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from topotrafficrl_b200.factory import load_environment
from topotrafficrl_b200.agent import QNetRollout
from bench import QNET_CONFIGS, random_state_dict

env = load_environment({"id": "intersection-v0", "import_module": "ttrl_env", "destination": "o1"})
q = QNetRollout(QNET_CONFIGS["ego_attention_2h"], random_state_dict("ego_attention_2h", 3), (15, 7), 3)
q.eval()
obs, _ = env.reset(seed=0)
n, t_step, t_act, episodes = 0, 0.0, 0.0, 0
for k in range(300):
    t0 = time.perf_counter()
    a = int(q.act(torch.as_tensor(obs[None], device="cuda"))[0])
    t1 = time.perf_counter()
    obs, r, term, trunc, info = env.step(a)
    t2 = time.perf_counter()
    if k >= 20:
        n += 1; t_act += t1 - t0; t_step += t2 - t1
    if term or trunc:
        obs, _ = env.reset(seed=1 + episodes); episodes += 1
print(json.dumps({"workload": "configs[0]: one intersection-v0 env through the gymnasium front end + ego-attention act", "steps": n,
                  "ms_per_env_step": 1e3 * t_step / n, "ms_per_act": 1e3 * t_act / n, "episodes": episodes,
                  "reference_ms_per_env_step": 66.0}))
