#!/usr/bin/env python
"""Where a training iteration of the batched DQN driver spends its time (CUDA events around each piece, 8192 intersection envs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from topotrafficrl_b200.vector_env import TTRLVectorEnv
from topotrafficrl_b200.trainer import BatchedDQNAgent

E = 8192
env = TTRLVectorEnv(E, scene="intersection", seed=0, vcap=24)
obs, _ = env.reset()
agent = BatchedDQNAgent(env, {"model": {"type": "MultiLayerPerceptron", "layers": [128, 128]}, "gamma": 0.95, "batch_size": 64, "memory_capacity": 15000,
                              "target_update": 512, "exploration": {"method": "EpsilonGreedy", "tau": 15000, "temperature": 1.0, "final_temperature": 0.05}},
                       seed=0, rollout_mode="tensor")
def timed(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) / n * 1e3
state = {"obs": obs}
def it():
    prev = state["obs"].clone()
    a = agent.act(prev)
    o, r, t, u, info = env.step(a)
    agent.record(prev, a, r, o, t, u, info)
    state["obs"] = o
print("full iteration         gpu %.3f ms  wall %.3f ms" % timed(it))
prev = obs.clone(); a = agent.act(prev); o, r, t, u, info = env.step(a)
print("act                    gpu %.3f ms  wall %.3f ms" % timed(lambda: agent.act(prev)))
print("env.step               gpu %.3f ms  wall %.3f ms" % timed(lambda: env.step(a)))
def push():
    done = info["_final_observation"].bool().view(E, 1, 1)
    ns = torch.where(done, info["final_observation"], o)
    agent.memory.push(prev, a, r, ns, t.bool())
print("where + memory.push    gpu %.3f ms  wall %.3f ms" % timed(push))
B = 64
print("randperm               gpu %.3f ms  wall %.3f ms" % timed(lambda: torch.randperm(agent.memory.size, device="cuda", generator=agent.gen)[:B]))
idx = torch.randperm(agent.memory.size, device="cuda", generator=agent.gen)[:B]
print("kernel update          gpu %.3f ms  wall %.3f ms" % timed(lambda: agent.kernel_update.update(idx, True)))
agent.training = True
def rec():
    agent.record(prev, a, r, o, t, u, info)
print("record                 gpu %.3f ms  wall %.3f ms" % timed(rec))
