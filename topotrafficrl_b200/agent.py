"""DQN Q-network rollout on the device (forward + epsilon-greedy action selection, one fused launch).

Mirrors the act path of the reference agent: ``AbstractDQNAgent.act`` (ttrl_agent/agents/deep_q_network/
abstract.py:65-83), ``model_factory`` configs (agents/common/models.py:431-441) and ``EpsilonGreedy``
(exploration/epsilon_greedy.py:32-48).  Weights are read from a reference ``state_dict`` (same key names:
``layers.0.weight``, ``ego_embedding.layers.0.weight``, ``attention_layer.key_all.weight``,
``output_layer.predict.weight`` ...), e.g. from a checkpoint written by ``DQNAgent.save`` (pytorch.py:82-86).
Training (replay, Bellman residual, optimiser) is outside this path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import abi
from ._lib import check, lib


def _np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float32)


def _dense(sd: Dict[str, np.ndarray], prefix: str, bias: bool = True):
    out = [np.ascontiguousarray(_np(sd[prefix + ".weight"]).T).ravel()]  # [out, in] -> [in][out]
    if bias:
        out.append(_np(sd[prefix + ".bias"]).ravel())
    return out


def pack_weights(model_config: dict, state_dict: Dict[str, np.ndarray], obs_shape, n_actions: int):
    """Return (QnetDesc, float32 blob) for a reference model config + state_dict."""
    mtype = model_config["type"]
    d = abi.QnetDesc()
    d.n_entities, d.n_features = (int(obs_shape[0]), int(np.prod(obs_shape[1:]))) if len(obs_shape) > 1 else (1, int(obs_shape[0]))
    d.n_actions = int(n_actions)
    parts = []
    if mtype == "MultiLayerPerceptron":
        d.type = abi.QNET_MLP
        layers = list(model_config.get("layers", [64, 64]))
        d.n_hidden = len(layers)
        for k, v in enumerate(layers):
            d.hidden[k] = int(v)
            parts += _dense(state_dict, f"layers.{k}")
        parts += _dense(state_dict, "predict")
    elif mtype == "EgoAttentionNetwork":
        d.type = abi.QNET_EGO_ATTENTION
        if model_config.get("self_attention_layer"):
            raise NotImplementedError("self_attention_layer is not used by the shipped configs (ego_attention*.json: null)")
        emb = list(model_config["embedding_layer"].get("layers", [64, 64]))
        if list(model_config["others_embedding_layer"].get("layers", [64, 64])) != emb:
            raise NotImplementedError("ego and others embeddings must share their layer sizes")
        d.embed_layers = len(emb)
        for k, v in enumerate(emb):
            d.embed[k] = int(v)
        att = model_config["attention_layer"]
        d.feature_size, d.heads = int(att.get("feature_size", 64)), int(att.get("heads", 4))
        outl = list(model_config["output_layer"].get("layers", [64, 64]))
        d.out_layers = len(outl)
        for k, v in enumerate(outl):
            d.out_hidden[k] = int(v)
        d.presence_feature_idx = int(model_config.get("presence_feature_idx", 0))
        for k in range(len(emb)):
            parts += _dense(state_dict, f"ego_embedding.layers.{k}")
        for k in range(len(emb)):
            parts += _dense(state_dict, f"others_embedding.layers.{k}")
        for name in ("key_all", "value_all", "query_ego", "attention_combine"):
            parts += _dense(state_dict, f"attention_layer.{name}", bias=False)
        for k in range(len(outl)):
            parts += _dense(state_dict, f"output_layer.layers.{k}")
        parts += _dense(state_dict, "output_layer.predict")
    elif mtype == "DuelingNetwork":
        d.type = abi.QNET_DUELING
        base = model_config.get("base_module", {})
        layers = list(base.get("layers", [64, 64]))
        if model_config.get("value", {}).get("layers") or model_config.get("advantage", {}).get("layers"):
            raise NotImplementedError("dueling heads with hidden layers are not used by the reference defaults")
        d.n_hidden = len(layers)
        for k, v in enumerate(layers):
            d.hidden[k] = int(v)
            parts += _dense(state_dict, f"base_module.layers.{k}")
        parts += _dense(state_dict, "value.predict")
        parts += _dense(state_dict, "advantage.predict")
    elif mtype == "ConvolutionalNetwork":
        raise NotImplementedError("ConvolutionalNetwork (grid_convnet.json) is outside the B200 hot path (SURVEY.md section 2 row 22)")
    else:
        raise ValueError("Unknown model type")
    blob = np.ascontiguousarray(np.concatenate(parts).astype(np.float32))
    return d, blob


def blob_keys(model_config: dict):
    """State-dict keys in the order of the weight blob (``pack_weights``); ``*.weight`` entries are stored transposed
    ([in][out]).  Used to rebuild the blob on the device from live torch parameters (``QNetRollout.load_parameters``)."""
    mtype = model_config["type"]
    keys = []

    def dense(prefix, bias=True):
        keys.append(prefix + ".weight")
        if bias:
            keys.append(prefix + ".bias")

    if mtype == "MultiLayerPerceptron":
        for k in range(len(model_config.get("layers", [64, 64]))):
            dense(f"layers.{k}")
        dense("predict")
    elif mtype == "EgoAttentionNetwork":
        n_emb = len(model_config["embedding_layer"].get("layers", [64, 64]))
        for k in range(n_emb):
            dense(f"ego_embedding.layers.{k}")
        for k in range(n_emb):
            dense(f"others_embedding.layers.{k}")
        for name in ("key_all", "value_all", "query_ego", "attention_combine"):
            dense(f"attention_layer.{name}", bias=False)
        for k in range(len(model_config["output_layer"].get("layers", [64, 64]))):
            dense(f"output_layer.layers.{k}")
        dense("output_layer.predict")
    elif mtype == "DuelingNetwork":
        for k in range(len((model_config.get("base_module") or {}).get("layers", [64, 64]))):
            dense(f"base_module.layers.{k}")
        dense("value.predict")
        dense("advantage.predict")
    else:
        raise ValueError("Unknown model type")
    return keys


class QNetRollout:
    """Batched ``agent.act``: obs [E, V, Fe] float32 on the device -> actions int32 [E] (and Q-values)."""

    MODES = {"fp32": 0, "tensor": 1}

    def __init__(self, model_config: dict, state_dict, obs_shape, n_actions: int, device: int = 0,
                 exploration: Optional[dict] = None, seed: int = 0, mode: str = "fp32") -> None:
        import torch

        self.torch = torch
        self._L = lib()
        self.desc, self.blob = pack_weights(model_config, state_dict, obs_shape, n_actions)
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        check(self._L.ttrl_qnet_create(C.byref(self.desc), self.blob.ctypes.data_as(C.c_void_p), self.blob.size,
                                       self.device_index, C.byref(h)))
        self._h = h
        self.n_actions = int(n_actions)
        self.set_mode(mode)
        # EpsilonGreedy schedule (epsilon_greedy.py:26-30 defaults; baseline.json overrides tau/final_temperature)
        ex = dict(method="EpsilonGreedy", temperature=1.0, final_temperature=0.1, tau=5000)
        ex.update(exploration or {})
        ex["final_temperature"] = min(ex["temperature"], ex["final_temperature"])
        self.exploration = ex
        self.time = 0
        self.seed = int(seed)
        self.training = True

    def set_mode(self, mode: str) -> None:
        """``"fp32"``: CUDA-core fp32 forward, the parity path (default).  ``"tensor"``: the GEMMs on the tcgen05 tensor
        cores (BF16x3 split, FP32 accumulation in TMEM): MultiLayerPerceptron with two hidden layers and the
        EgoAttentionNetwork shapes of the shipped configs; any other network raises :class:`TTRLError` (no silent fallback)."""
        if mode not in self.MODES:
            raise ValueError(f"unknown mode {mode!r}")
        check(self._L.ttrl_qnet_set_mode(self._h, self.MODES[mode]))
        self.mode = mode

    def load_parameters(self, module, model_config: dict) -> None:
        """Refresh the rollout network from a live torch module ON THE DEVICE (no host round trip): the blob is rebuilt with
        torch ops on the current stream and copied into the library's weight buffer (``ttrl_qnet_set_weights``)."""
        torch = self.torch
        sd = dict(module.named_parameters())
        parts = []
        for key in blob_keys(model_config):
            p = sd[key].detach()
            parts.append((p.t().contiguous() if key.endswith(".weight") else p).reshape(-1))
        blob = torch.cat(parts).to(dtype=torch.float32, device=self.device).contiguous()
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        check(self._L.ttrl_qnet_set_weights(self._h, blob.data_ptr(), blob.numel(), 1, stream))
        self._blob_keepalive = blob  # until the enqueued copy has run

    def close(self):
        if getattr(self, "_h", None):
            self._L.ttrl_qnet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def eval(self):  # AbstractDQNAgent.eval abstract.py:167-170: Greedy exploration
        self.training = False
        self.exploration["method"] = "Greedy"

    @property
    def epsilon(self) -> float:
        if self.exploration["method"] == "Greedy":
            return 0.0
        e = self.exploration
        return float(e["final_temperature"] + (e["temperature"] - e["final_temperature"]) * np.exp(-self.time / e["tau"]))

    def act(self, obs, step_exploration_time: bool = True, return_q: bool = False, uniforms=None):
        torch = self.torch
        if step_exploration_time:
            self.time += 1
        E = obs.shape[0]
        obs = obs.contiguous()
        actions = torch.empty(E, dtype=torch.int32, device=self.device)
        q = torch.empty((E, self.n_actions), dtype=torch.float32, device=self.device) if return_q else None
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        if uniforms is not None:
            u = torch.as_tensor(np.asarray(uniforms, dtype=np.float64), device=self.device)
            check(self._L.ttrl_qnet_act_injected(self._h, obs.data_ptr(), E, C.c_double(self.epsilon), u.data_ptr(),
                                                 actions.data_ptr(), q.data_ptr() if q is not None else None, stream))
        else:
            check(self._L.ttrl_qnet_act(self._h, obs.data_ptr(), E, self.epsilon, self.seed, self.time, actions.data_ptr(),
                                        q.data_ptr() if q is not None else None, stream))
        return (actions, q) if return_q else actions
