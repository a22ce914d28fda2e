"""Config / JSON loader parity with ``ttrl_agent/agents/common/factory.py:30-94`` and ``ttrl_agent/configuration.py:12-44``
(SURVEY.md section 8f, row N4): the reference's env / agent JSON files load unchanged against this backend.

* :func:`load_agent_config` -- JSON with ``base_config`` inheritance (paths relative to the working directory, like the
  reference; a ``search_path`` can be given) merged with :func:`rec_update`.
* :func:`load_environment` -- ``{"id": ..., "import_module": ..., <config keys>}`` -> a configured, reset env from this
  package's registry.  ``import_module: "ttrl_env"`` is accepted and maps to this package.  With ``num_envs`` it returns
  the vector form (:class:`TTRLVectorEnv`) of the same scene.
* :func:`load_agent` -- the agent ``__class__`` strings of the reference map to this package's DQN driver.
"""
from __future__ import annotations

import copy
import json
import os
from collections.abc import Mapping
from typing import Optional, Union

from . import _gym

SCENE_OF_ID = {"intersection-v0": "intersection", "intersection-multi-agent-v0": "intersection",
               "intersection-multi-agent-v1": "intersection", "roundabout-v0": "roundabout", "u-turn-v0": "u-turn"}
_OWN_MODULES = ("ttrl_env", "topotrafficrl_b200", "topotrafficrl_b200.envs")


def rec_update(d: dict, u: Mapping) -> dict:
    """``Configurable.rec_update`` (configuration.py:30-44): recursive update of ``d`` with ``u``; returns ``d``."""
    for k, v in u.items():
        if isinstance(v, Mapping):
            d[k] = rec_update(d.get(k, {}), v)
        else:
            d[k] = v
    return d


def load_agent_config(config_path: str, search_path: Optional[str] = None) -> dict:
    """factory.py:45-57: ``base_config`` chains are resolved first, the child overrides recursively."""
    path = config_path
    if not os.path.exists(path) and search_path is not None:
        path = os.path.join(search_path, config_path)
    with open(path) as f:
        agent_config = json.loads(f.read())
    if "base_config" in agent_config:
        base = load_agent_config(agent_config["base_config"], search_path)
        del agent_config["base_config"]
        agent_config = rec_update(base, agent_config)
    return agent_config


def _read_env_config(env_config: Union[str, dict]) -> dict:
    if not isinstance(env_config, dict):
        with open(env_config) as f:
            env_config = json.loads(f.read())
    return copy.deepcopy(env_config)


def load_environment(env_config: Union[str, dict], num_envs: Optional[int] = None, **vector_kwargs):
    """factory.py:59-94.  ``num_envs=None``: the single env from the registry, configured with the whole dict (the reference
    passes ``id`` / ``import_module`` to ``configure`` too) and reset.  ``num_envs=E``: ``TTRLVectorEnv`` of that scene."""
    cfg = _read_env_config(env_config)
    module = cfg.get("import_module", None)
    if module and module not in _OWN_MODULES:
        __import__(module)
    try:
        env_id = cfg["id"]
    except KeyError:
        raise ValueError("The gym register id of the environment must be provided")
    from . import envs  # noqa: F401  (registers the ids)
    if num_envs is not None:
        if env_id not in SCENE_OF_ID:
            raise ValueError(f"Environment {env_id} not registered")
        from .vector_env import TTRLVectorEnv
        scene_cfg = {k: v for k, v in cfg.items() if k not in ("id", "import_module")}
        if env_id.startswith("intersection-multi-agent"):
            from . import scenes
            scene_cfg = dict({k: scenes.MULTI_AGENT_INTERSECTION_CONFIG[k] for k in ("action", "observation", "controlled_vehicles")},
                             **scene_cfg)
        env = TTRLVectorEnv(int(num_envs), SCENE_OF_ID[env_id], config=scene_cfg, **vector_kwargs)
        env.import_module = module
        return env
    if not _gym.HAVE_GYMNASIUM and env_id not in _gym._registry:
        raise ValueError(f"Environment {env_id} not registered. The environment module should be specified by "
                         'the "import_module" key of the environment configuration')
    env = _gym.make(env_id, config=cfg) if not _gym.HAVE_GYMNASIUM else _gym.make(env_id, render_mode="rgb_array", config=cfg)
    env.import_module = module
    return env


AGENT_CLASSES = {
    "<class 'ttrl_agent.agents.deep_q_network.pytorch.DQNAgent'>": "dqn",
    "<class 'rl_agents.agents.deep_q_network.pytorch.DQNAgent'>": "dqn",
    "<class 'topotrafficrl_b200.trainer.BatchedDQNAgent'>": "dqn",
}


def agent_factory(environment, config: dict, **kwargs):
    """factory.py:12-27: ``config["__class__"]`` selects the agent; only the DQN family is on the B200 path."""
    if "__class__" not in config:
        raise ValueError("The configuration should specify the agent __class__")
    kind = AGENT_CLASSES.get(config["__class__"])
    if kind is None:
        raise NotImplementedError(f"agent class {config['__class__']} is outside the B200 hot path (SURVEY.md section 2)")
    from .trainer import BatchedDQNAgent
    return BatchedDQNAgent(environment, config, **kwargs)


def load_agent(agent_config: Union[str, dict], env, search_path: Optional[str] = None, **kwargs):
    """factory.py:30-42."""
    if not isinstance(agent_config, dict):
        agent_config = load_agent_config(agent_config, search_path)
    return agent_factory(env, agent_config, **kwargs)
