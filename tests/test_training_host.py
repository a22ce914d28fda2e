"""Host-side pieces of SURVEY.md section 8f rows N2 / N4 (CPU only): the trainable torch forms of the Q-network families
against the reference's own forward outputs (tests/golden/qnet_vectors.npz, generated from ``model_factory`` of the unmodified
reference), the replay memory, and the JSON / config loader."""
import json
import os

import numpy as np
import pytest
import torch

from topotrafficrl_b200 import factory
from topotrafficrl_b200.agent import blob_keys, pack_weights
from topotrafficrl_b200.models import model_factory, size_model_config
from topotrafficrl_b200.trainer import BatchedDQNAgent, DeviceReplayMemory
from tests import common as T

EGO = {"type": "EgoAttentionNetwork",
       "embedding_layer": {"type": "MultiLayerPerceptron", "layers": [64, 64], "reshape": False, "in": 7},
       "others_embedding_layer": {"type": "MultiLayerPerceptron", "layers": [64, 64], "reshape": False, "in": 7},
       "self_attention_layer": None,
       "attention_layer": {"type": "EgoAttention", "feature_size": 64, "heads": 1},
       "output_layer": {"type": "MultiLayerPerceptron", "layers": [64, 64], "reshape": False}}
CONFIGS = {
    "mlp": {"type": "MultiLayerPerceptron", "layers": [128, 128]},
    "ego1h": EGO,
    "ego2h": dict(EGO, attention_layer={"type": "EgoAttention", "feature_size": 64, "heads": 2}),
    "dueling": {"type": "DuelingNetwork", "base_module": {"type": "MultiLayerPerceptron", "layers": [64, 64]},
                "value": {"type": "MultiLayerPerceptron", "layers": []}, "advantage": {"type": "MultiLayerPerceptron", "layers": []}},
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_trainable_models_reproduce_reference_forward(name):
    """Same parameter names (a reference state_dict loads with strict=True) and the same Q-values as the reference modules."""
    g = T.golden("qnet_vectors.npz")
    cfg = size_model_config((15, 7), 3, CONFIGS[name])
    net = model_factory(cfg)
    sd = {k[len(name) + 1:]: torch.tensor(g[k]) for k in g.files if k.startswith(name + "/")}
    net.load_state_dict(sd, strict=True)
    with torch.no_grad():
        q = net(torch.tensor(g["obs"])).numpy()
    np.testing.assert_allclose(q, g[name + "_q"], rtol=0, atol=2e-6)
    # the weight-blob order used to refresh the rollout kernels from live parameters == pack_weights' order
    _, blob = pack_weights(cfg, {k: v.numpy() for k, v in sd.items()}, (15, 7), 3)
    parts = [(sd[k].t().contiguous() if k.endswith(".weight") else sd[k]).reshape(-1) for k in blob_keys(cfg)]
    np.testing.assert_array_equal(torch.cat(parts).numpy(), blob)
    # gradients flow to every parameter
    net(torch.tensor(g["obs"])).sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_model_factory_errors():
    with pytest.raises(ValueError):
        model_factory({"type": "Nope"})
    with pytest.raises(NotImplementedError):
        size_model_config((7, 32, 32), 3, {"type": "ConvolutionalNetwork"})


def test_replay_memory_ring_and_sampling():
    gen = torch.Generator().manual_seed(0)
    mem = DeviceReplayMemory(10, (2, 3), torch.device("cpu"), gen)
    def batch(lo, n):
        s = torch.arange(lo, lo + n, dtype=torch.float32).view(n, 1, 1).expand(n, 2, 3).contiguous()
        return s, torch.arange(lo, lo + n) % 3, torch.arange(lo, lo + n, dtype=torch.float32), s + 0.5, (torch.arange(lo, lo + n) % 4 == 0)
    mem.push(*batch(0, 6))
    assert len(mem) == 6 and mem.position == 6
    mem.push(*batch(6, 7))   # wraps: the oldest three are overwritten
    assert len(mem) == 10 and mem.position == 3
    assert sorted(mem.reward.tolist()) == [float(v) for v in range(3, 13)]
    s, a, r, ns, t = mem.sample(8)
    assert len(set(r.tolist())) == 8                       # without replacement
    assert torch.equal(s[:, 0, 0], r) and torch.equal(ns[:, 0, 0], r + 0.5) and torch.equal(a, r.long() % 3)
    assert torch.equal(t, r.long() % 4 == 0)
    mem.push(*batch(100, 25))  # larger than the capacity: the newest 10 stay
    assert sorted(mem.reward.tolist()) == [float(v) for v in range(115, 125)]


def test_bellman_residual_matches_reference_formula():
    """compute_bellman_residual without constructing the (CUDA) rollout: double-DQN target of pytorch.py:41-73."""
    cfg = size_model_config((15, 7), 3, CONFIGS["mlp"])
    agent = BatchedDQNAgent.__new__(BatchedDQNAgent)
    agent.torch, agent.config = torch, dict(BatchedDQNAgent.default_config(), gamma=0.95, double=True)
    torch.manual_seed(1)
    agent.value_net, agent.target_net = model_factory(cfg), model_factory(cfg)
    agent.loss_function = torch.nn.functional.mse_loss
    B = 32
    s, ns = torch.rand(B, 15, 7), torch.rand(B, 15, 7)
    a, r, t = torch.randint(0, 3, (B,)), torch.rand(B), torch.rand(B) < 0.3
    loss = agent.compute_bellman_residual((s, a, r, ns, t))
    with torch.no_grad():
        qv, qn, qt = agent.value_net(s), agent.value_net(ns), agent.target_net(ns)
        want = 0.0
        for k in range(B):
            boot = 0.0 if t[k] else float(qt[k, int(qn[k].argmax())])
            want += (float(qv[k, a[k]]) - (float(r[k]) + 0.95 * boot)) ** 2
    assert abs(float(loss) - want / B) < 1e-5
    agent.config["double"] = False
    loss2 = agent.compute_bellman_residual((s, a, r, ns, t))
    with torch.no_grad():
        want2 = sum((float(qv[k, a[k]]) - (float(r[k]) + 0.95 * (0.0 if t[k] else float(qt[k].max())))) ** 2 for k in range(B)) / B
    assert abs(float(loss2) - want2) < 1e-5


def test_agent_config_inheritance_and_rec_update(tmp_path):
    """load_agent_config: base_config chains + recursive override (factory.py:45-57), on copies of the reference's layout."""
    d = tmp_path / "configs" / "IntersectionEnv" / "agents" / "DQNAgent"
    d.mkdir(parents=True)
    (d / "baseline.json").write_text(json.dumps({
        "__class__": "<class 'ttrl_agent.agents.deep_q_network.pytorch.DQNAgent'>",
        "model": {"type": "MultiLayerPerceptron", "layers": [128, 128]}, "gamma": 0.95, "n_steps": 1, "batch_size": 64,
        "memory_capacity": 15000, "target_update": 512,
        "exploration": {"method": "EpsilonGreedy", "tau": 15000, "temperature": 1.0, "final_temperature": 0.05}}))
    (d / "ego_attention.json").write_text(json.dumps({
        "base_config": "configs/IntersectionEnv/agents/DQNAgent/baseline.json", "model": dict(EGO)}))
    (d / "ego_attention_2h.json").write_text(json.dumps({
        "base_config": "configs/IntersectionEnv/agents/DQNAgent/ego_attention.json", "model": {"attention_layer": {"heads": 2}}}))
    cfg = factory.load_agent_config("configs/IntersectionEnv/agents/DQNAgent/ego_attention_2h.json", search_path=str(tmp_path))
    assert "base_config" not in cfg and cfg["gamma"] == 0.95 and cfg["exploration"]["tau"] == 15000
    assert cfg["model"]["type"] == "EgoAttentionNetwork" and cfg["model"]["attention_layer"] == {"type": "EgoAttention", "feature_size": 64, "heads": 2}
    assert cfg["model"]["layers"] == [128, 128]  # rec_update keeps the parent's keys, like the reference
    cwd = os.getcwd()
    os.chdir(tmp_path)  # the reference resolves base_config against the working directory (scripts/example.py chdir)
    try:
        assert factory.load_agent_config("configs/IntersectionEnv/agents/DQNAgent/ego_attention_2h.json") == cfg
    finally:
        os.chdir(cwd)
    d0 = {"a": {"b": 1, "c": 2}, "x": 1}
    assert factory.rec_update(d0, {"a": {"b": 5}, "y": {"z": 0}}) is d0 and d0 == {"a": {"b": 5, "c": 2}, "x": 1, "y": {"z": 0}}


def test_loader_errors():
    with pytest.raises(ValueError):
        factory.load_environment({"observation": {"type": "Kinematics"}})            # no id
    with pytest.raises(ValueError):
        factory.load_environment({"id": "highway-v0"})                               # not registered
    with pytest.raises(ValueError):
        factory.agent_factory(None, {"model": {}})                                    # no __class__
    with pytest.raises(NotImplementedError):
        factory.agent_factory(None, {"__class__": "<class 'ttrl_agent.agents.tree_search.mcts.MCTSAgent'>"})
