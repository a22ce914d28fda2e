"""Every registered env id x E in {1, 3, 200} through load_environment / load_agent / BatchedEvaluation (CUDA-graph update): edge sizes
of the grids, the reset lists and the regeneration queue.  Run on a B200: python tools/edge_check.py"""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from topotrafficrl_b200.factory import load_environment, load_agent
from topotrafficrl_b200.trainer import BatchedEvaluation
from scripts.train_example import AGENT_CONFIG
for env_cfg in ({"id": "intersection-v0"}, {"id": "roundabout-v0"}, {"id": "u-turn-v0"}, {"id": "intersection-multi-agent-v0"}):
    for E in (1, 3, 200):
        env = load_environment(dict(env_cfg), num_envs=E, seed=2)
        model = AGENT_CONFIG["model"] if len(env.obs_shape) == 2 and env.obs_shape[-1] == 7 or (env.num_agents > 1) else {"type": "MultiLayerPerceptron", "layers": [64, 64]}
        if env.num_agents > 1 or env.obs_shape[-1] != 7:
            model = {"type": "MultiLayerPerceptron", "layers": [64, 64]}
        agent = load_agent(dict(AGENT_CONFIG, model=model, batch_size=16), env, cuda_graph=True)
        out = BatchedEvaluation(env, agent, num_steps=40).train()
        assert np.isfinite(out["mean_return"]) and out["env_steps"] == 40 * E, out
        print(env_cfg["id"], E, env.obs_shape, {k: round(float(v), 3) for k, v in out.items()})
        agent.close(); env.close()
print("edge ok")
