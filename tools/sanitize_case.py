#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): a few env-steps of
every kernel family on tiny batches.  usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import QNET_CONFIGS, random_state_dict
from topotrafficrl_b200.agent import QNetRollout
from topotrafficrl_b200.vector_env import TTRLVectorEnv

for scene, kw in (("highway", {}), ("highway", {"config": {"vehicles_count": 200, "vehicles_density": 4.0}}), ("intersection", {})):
    env = TTRLVectorEnv(6, scene=scene, seed=1, **kw)
    obs, _ = env.reset()
    n_act = env.single_action_space.n
    for k in range(4):
        a = torch.randint(0, n_act, (6,), dtype=torch.int32, device="cuda")
        obs, r, t, u, _ = env.step(a)
    o, r, t, u, _ = env.step_host(np.zeros(6, np.int32))
    torch.cuda.synchronize()
    print(scene, kw, "ok", env.stats()["env_steps"])
    env.close()
obs = torch.rand(70, 15, 7, device="cuda")
for kind in ("mlp", "ego_attention_2h"):
    for mode in ("fp32", "tensor"):
        net = QNetRollout(QNET_CONFIGS[kind], random_state_dict(kind, 3), (15, 7), 3, mode=mode)
        a, q = net.act(obs, return_q=True)
        torch.cuda.synchronize()
        print(kind, mode, "ok", float(q.abs().max()))
        net.close()
