"""What the reference's example script does (load an environment and a DQN agent from its JSON configuration format, train,
save, run the learned policy greedily), against the B200 backend.

    python scripts/train_example.py [--num-envs 8192] [--steps 2000] [--env-config PATH] [--agent-config PATH]

``--num-envs 1`` is BASELINE config[0] (one env, like the reference); larger values use the same configs on the vector env.
Without paths the built-in dicts below are used (the values of the reference's IntersectionEnv ``env.json`` and
``agents/DQNAgent/ego_attention_2h.json`` after ``base_config`` resolution); the reference's own files load unchanged:
``--env-config <reference>/scripts/configs/IntersectionEnv/env.json --agent-config <reference>/scripts/configs/...`` (run from
``<reference>/scripts`` so that the relative ``base_config`` paths resolve, as the reference does).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from topotrafficrl_b200.factory import load_agent, load_environment  # noqa: E402
from topotrafficrl_b200.trainer import BatchedEvaluation  # noqa: E402


ENV_CONFIG = {"id": "intersection-v0", "import_module": "ttrl_env", "destination": "o1",
              "observation": {"type": "Kinematics", "vehicles_count": 15, "features": ["presence", "x", "y", "vx", "vy", "cos_h", "sin_h"],
                              "features_range": {"x": [-100, 100], "y": [-100, 100], "vx": [-20, 20], "vy": [-20, 20]},
                              "absolute": True, "order": "shuffled"}}
_MLP = {"type": "MultiLayerPerceptron", "layers": [64, 64], "reshape": False}
AGENT_CONFIG = {"__class__": "<class 'ttrl_agent.agents.deep_q_network.pytorch.DQNAgent'>", "gamma": 0.95, "n_steps": 1, "batch_size": 64,
                "memory_capacity": 15000, "target_update": 512,
                "exploration": {"method": "EpsilonGreedy", "tau": 15000, "temperature": 1.0, "final_temperature": 0.05},
                "model": {"type": "EgoAttentionNetwork", "embedding_layer": dict(_MLP, **{"in": 7}), "others_embedding_layer": dict(_MLP, **{"in": 7}),
                          "self_attention_layer": None, "attention_layer": {"type": "EgoAttention", "feature_size": 64, "heads": 2},
                          "output_layer": dict(_MLP)}}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--test-steps", type=int, default=50)
    ap.add_argument("--env-config", default="")
    ap.add_argument("--agent-config", default="")
    ap.add_argument("--checkpoint", default="")
    args = ap.parse_args()
    env = load_environment(args.env_config or ENV_CONFIG, num_envs=args.num_envs)
    agent = load_agent(args.agent_config or AGENT_CONFIG, env, rollout_mode="tensor", cuda_graph=True)
    evaluation = BatchedEvaluation(env, agent, num_steps=args.steps)
    print(f"Ready to train a {agent.model_config['type']} DQN on {args.num_envs} x {env.scene} envs")
    t0 = time.perf_counter()
    out = evaluation.train(log_every=max(args.steps // 5, 1))
    dt = time.perf_counter() - t0
    for row in evaluation.history:
        print(row)
    print(f"trained {int(out['env_steps'])} env-steps in {dt:.2f} s ({out['env_steps'] / dt:.3e} env-steps/s); "
          f"mean return {out['mean_return']:.3f}, crash rate {out['crash_rate']:.3f}")
    if args.checkpoint:
        print("saved", agent.save(args.checkpoint))
    res = BatchedEvaluation(env, agent, num_steps=args.test_steps).test()
    print("greedy test:", {k: round(float(v), 4) for k, v in res.items()})
    agent.close()
    env.close()


if __name__ == "__main__":
    main()
