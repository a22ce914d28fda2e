"""Trainable torch forms of the reference's Q-network families, used by the batched training driver (``trainer.py``) for
the backward pass.  The ROLLOUT never runs these: acting goes through the CUDA kernels of ``csrc/ttrl_qnet.cu``
(``agent.QNetRollout``), whose weights are refreshed from the modules below.

Parameter names equal the reference's (``ttrl_agent/agents/common/models.py``: ``layers.<k>``, ``predict``,
``ego_embedding`` / ``others_embedding``, ``attention_layer.{key_all, value_all, query_ego, attention_combine}``,
``output_layer``, ``base_module`` / ``value`` / ``advantage``), so a ``state_dict`` written by ``DQNAgent.save``
(pytorch.py:82-86) loads here and vice versa.  Initialisation is torch's ``nn.Linear`` default, as in the reference
(its XAVIER ``reset`` only runs from ``initialize_model``).
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch import nn
from torch.nn import functional as F


def _activation(name: str):
    if name == "RELU":
        return F.relu
    if name == "TANH":
        return torch.tanh
    raise ValueError(f"Unknown activation_type: {name}")


class MultiLayerPerceptron(nn.Module):
    """models.py:50-76.  ``reshape`` is truthy for the string "True" too, like the reference's default."""

    def __init__(self, n_in: int, layers: Sequence[int], n_out=None, activation: str = "RELU", reshape=True) -> None:
        super().__init__()
        sizes = [int(n_in)] + [int(v) for v in layers]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:]))
        self.predict = nn.Linear(sizes[-1], int(n_out)) if n_out else None
        self.act, self.reshape, self.width = _activation(activation), bool(reshape), sizes[-1]

    @classmethod
    def from_config(cls, cfg: dict) -> "MultiLayerPerceptron":
        return cls(cfg["in"], cfg.get("layers", [64, 64]), cfg.get("out"), cfg.get("activation", "RELU"), cfg.get("reshape", "True"))

    def forward(self, x):
        if self.reshape:
            x = x.reshape(x.shape[0], -1)
        for layer in self.layers:
            x = self.act(layer(x))
        return self.predict(x) if self.predict is not None else x


class DuelingNetwork(nn.Module):
    """models.py:79-104: V + A - mean(A) over a shared base."""

    def __init__(self, cfg: dict) -> None:
        super().__init__()
        base = dict(cfg.get("base_module") or {"type": "MultiLayerPerceptron"}, **{"in": cfg["in"], "out": None})
        self.base_module = MultiLayerPerceptron.from_config(base)
        w = self.base_module.width
        self.value = MultiLayerPerceptron.from_config(dict(cfg.get("value") or {"layers": []}, **{"in": w, "out": 1}))
        self.advantage = MultiLayerPerceptron.from_config(dict(cfg.get("advantage") or {"layers": []}, **{"in": w, "out": cfg["out"]}))

    def forward(self, x):
        h = self.base_module(x)
        adv = self.advantage(h)
        return self.value(h) + adv - adv.mean(1, keepdim=True)


class EgoAttention(nn.Module):
    """models.py:157-194: the ego's query against the keys of all entities, per head; output averaged with the ego embedding."""

    def __init__(self, feature_size: int = 64, heads: int = 4) -> None:
        super().__init__()
        self.fs, self.heads = int(feature_size), int(heads)
        self.value_all = nn.Linear(self.fs, self.fs, bias=False)
        self.key_all = nn.Linear(self.fs, self.fs, bias=False)
        self.query_ego = nn.Linear(self.fs, self.fs, bias=False)
        self.attention_combine = nn.Linear(self.fs, self.fs, bias=False)

    def forward(self, ego, others, absent):
        B, H, dk = others.shape[0], self.heads, self.fs // self.heads
        everyone = torch.cat((ego.view(B, 1, self.fs), others), dim=1)
        N = everyone.shape[1]
        k = self.key_all(everyone).view(B, N, H, dk).transpose(1, 2)
        v = self.value_all(everyone).view(B, N, H, dk).transpose(1, 2)
        q = self.query_ego(ego).view(B, 1, H, dk).transpose(1, 2)
        scores = q @ k.transpose(-2, -1) / math.sqrt(dk)                      # models.py:370-388
        scores = scores.masked_fill(absent.view(B, 1, 1, N), -1e9)
        p = F.softmax(scores, dim=-1)
        out = (p @ v).reshape(B, self.fs)
        return (self.attention_combine(out) + ego.squeeze(1)) / 2, p


class EgoAttentionNetwork(nn.Module):
    """models.py:237-312 (without the optional self-attention block, which no shipped config enables)."""

    def __init__(self, cfg: dict) -> None:
        super().__init__()
        if cfg.get("self_attention_layer"):
            raise NotImplementedError("self_attention_layer is not used by the shipped configs (ego_attention*.json: null)")
        emb = dict(cfg["embedding_layer"])
        oth = dict(cfg["others_embedding_layer"])
        emb["in"] = emb.get("in") or cfg["in"]
        oth["in"] = oth.get("in") or cfg["in"]
        att = cfg["attention_layer"]
        self.presence_idx = int(cfg.get("presence_feature_idx", 0))
        self.ego_embedding = MultiLayerPerceptron.from_config(emb)
        self.others_embedding = MultiLayerPerceptron.from_config(oth)
        self.attention_layer = EgoAttention(att.get("feature_size", 64), att.get("heads", 4))
        self.output_layer = MultiLayerPerceptron.from_config(dict(cfg["output_layer"], **{"in": self.attention_layer.fs, "out": cfg["out"]}))

    def forward(self, x):
        absent = x[:, :, self.presence_idx] < 0.5
        ego = self.ego_embedding(x[:, 0:1, :])
        others = self.others_embedding(x[:, 1:, :])
        att, _ = self.attention_layer(ego, others, absent)
        return self.output_layer(att)


def size_model_config(obs_shape, n_actions: int, model_config: dict) -> dict:
    """``size_model_config`` (models.py:399-426) for Box / Tuple-of-Box observations and Discrete actions."""
    cfg = dict(model_config)
    if cfg["type"] == "ConvolutionalNetwork":
        raise NotImplementedError("ConvolutionalNetwork (grid_convnet.json) is outside the B200 hot path (SURVEY.md section 2 row 22)")
    n_in = 1
    for v in obs_shape:
        n_in *= int(v)
    cfg["in"], cfg["out"] = n_in, int(n_actions)
    return cfg


def model_factory(config: dict) -> nn.Module:
    """models.py:431-441."""
    t = config["type"]
    if t == "MultiLayerPerceptron":
        return MultiLayerPerceptron.from_config(config)
    if t == "DuelingNetwork":
        return DuelingNetwork(config)
    if t == "EgoAttentionNetwork":
        return EgoAttentionNetwork(config)
    if t == "ConvolutionalNetwork":
        raise NotImplementedError("ConvolutionalNetwork (grid_convnet.json) is outside the B200 hot path (SURVEY.md section 2 row 22)")
    raise ValueError("Unknown model type")
