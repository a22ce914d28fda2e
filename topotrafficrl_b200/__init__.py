"""B200-native batched simulator for the TopoTrafficRL per-step hot path.

Public surface (mirrors the reference's ``ttrl_env`` / ``ttrl_agent`` names for the hot path):

* :class:`TTRLVectorEnv` -- E envs in lockstep on one GPU (``vector_env.py``)
* :class:`IntersectionEnv`, :class:`MultiAgentIntersectionEnv`, :class:`RoundaboutEnv`, :class:`UTurnEnv` -- single-env
  gymnasium-shaped front ends (``envs.py``), ids ``intersection-v0``, ``intersection-multi-agent-v0/-v1``, ``roundabout-v0``,
  ``u-turn-v0``
* :class:`QNetRollout` -- batched DQN ``act`` (``agent.py``)
* ``factory.load_environment / load_agent_config / load_agent`` -- the reference's JSON configuration format (``factory.py``)
* ``trainer.BatchedDQNAgent / BatchedEvaluation`` -- batched DQN rollout / training driver (``trainer.py``, ``models.py``)
* :class:`Sim` -- the C ABI as an object (``sim.py``)

Importing the package does not load CUDA; the first use of any class above loads
``csrc/libttrl_b200.so`` and raises if it (or a CUDA device) is missing -- there is no CPU fallback.
"""
from . import abi, road, scenes, state  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name == "TTRLVectorEnv":
        from .vector_env import TTRLVectorEnv
        return TTRLVectorEnv
    if name in ("IntersectionEnv", "AbstractEnv", "MultiAgentIntersectionEnv", "RoundaboutEnv", "UTurnEnv", "MultiAgentWrapper"):
        from . import envs
        return getattr(envs, name)
    if name == "QNetRollout":
        from .agent import QNetRollout
        return QNetRollout
    if name == "Sim":
        from .sim import Sim
        return Sim
    raise AttributeError(name)
