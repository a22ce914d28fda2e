#!/usr/bin/env python
"""How many vehicle slots does the intersection scene need?  Runs E envs with random actions and fresh device resets at a
generous capacity and reports the distribution of the vehicle count per env (the reference has no cap; a full env rejects a spawn)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from topotrafficrl_b200 import TTRLVectorEnv, abi
E, steps, vcap = 8192, 300, 32
env = TTRLVectorEnv(E, "intersection", vcap=vcap, seed=1)
env.reset()
gen = torch.Generator(device="cuda").manual_seed(0)
hist = np.zeros(vcap + 1, np.int64)
for k in range(steps):
    env.step(torch.randint(0, 3, (E,), dtype=torch.int32, device="cuda", generator=gen))
    if k % 5 == 4:
        n = env.get_state().env_i[abi.EI_NVEH]
        hist += np.bincount(n, minlength=vcap + 1)
tot = hist.sum()
print(json.dumps({"envs": E, "steps": steps, "max_vehicles": int(np.nonzero(hist)[0].max()), "mean": float((hist * np.arange(vcap + 1)).sum() / tot),
                  "share_ge_15": float(hist[15:].sum() / tot), "share_ge_16": float(hist[16:].sum() / tot), "share_ge_20": float(hist[20:].sum() / tot),
                  "hist": hist.tolist()}))
