#!/usr/bin/env python
"""Q-network rollout kernels: time per call (CUDA events) and agreement between the fp32 parity path and the
tcgen05 tensor-core path.  usage: python tools/bench_qnet.py [E ...]"""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bench import QNET_CONFIGS, random_state_dict
from topotrafficrl_b200.agent import QNetRollout

def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for E in [int(x) for x in sys.argv[1:]] or [8192, 131072]:
    obs = (torch.rand(E, 15, 7, device="cuda") * 2 - 1)
    obs[:, :, 0] = (torch.rand(E, 15, device="cuda") < 0.7).float(); obs[:, 0, 0] = 1
    out = {"E": E}
    for kind in ("mlp", "ego_attention_2h"):
        sd = random_state_dict(kind, 3)
        net = QNetRollout(QNET_CONFIGS[kind], sd, (15, 7), 3); net.eval()
        a32, q32 = net.act(obs, return_q=True)
        out[kind + "_fp32_ms"] = timeit(lambda: net.act(obs))
        flops = {"mlp": 2 * (105 * 128 + 128 * 128 + 128 * 3), "ego_attention_2h": 419e3}[kind] * E
        out[kind + "_fp32_tflops"] = flops / out[kind + "_fp32_ms"] / 1e9
        net.set_mode("tensor")
        at, qt = net.act(obs, return_q=True)
        out[kind + "_tensor_ms"] = timeit(lambda: net.act(obs))
        out[kind + "_tensor_tflops_useful"] = flops / out[kind + "_tensor_ms"] / 1e9
        out[kind + "_tensor_max_abs_dq"] = float((qt - q32).abs().max())
        out[kind + "_tensor_action_agreement"] = float((at == a32).float().mean())
        gap = torch.sort(q32, dim=1).values
        out[kind + "_max_top2_gap_where_disagree"] = float((gap[:, -1] - gap[:, -2])[at != a32].max()) if (at != a32).any() else None
        net.close()
    print(json.dumps(out))
