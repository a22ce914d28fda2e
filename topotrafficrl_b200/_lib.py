"""ctypes binding of ``csrc/libttrl_b200.so`` (C ABI declared in ``include/ttrl_b200.h``).

There is NO CPU fallback: if the CUDA library is missing or no CUDA device is present, every entry point of
the package raises.  The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# TTRL_B200_LIB selects an alternative build of the SAME CUDA library (kernel tuning experiments)
LIB_PATH = os.environ.get("TTRL_B200_LIB") or os.path.join(_HERE, "csrc", "libttrl_b200.so")
_lib = None


class TTRLError(RuntimeError):
    pass


def build() -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess

    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TTRLError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the simulator has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    L.ttrl_last_error.restype = C.c_char_p
    L.ttrl_abi_version.restype = i32
    L.ttrl_abi_sizeof.argtypes = [i32]
    L.ttrl_sim_create.argtypes = [C.POINTER(abi.Config), vp, vp, vp, vp, i32, i32, i32, C.POINTER(vp)]
    L.ttrl_sim_destroy.argtypes = [vp]
    for name in ("ttrl_sim_num_envs", "ttrl_sim_vcap", "ttrl_sim_obs_size", "ttrl_sim_num_agents"):
        getattr(L, name).argtypes = [vp]
    L.ttrl_sim_launch_count.argtypes = [vp]
    L.ttrl_sim_launch_count.restype = i64
    L.ttrl_sim_set_spawn_routes.argtypes = [vp, vp, vp, vp]
    L.ttrl_sim_set_state.argtypes = [vp, vp, vp, vp, vp]
    L.ttrl_sim_get_state.argtypes = [vp, vp, vp, vp, vp]
    L.ttrl_sim_set_linear_params.argtypes = [vp, vp]
    L.ttrl_sim_get_linear_params.argtypes = [vp, vp]
    L.ttrl_sim_set_reset_pool_linear_params.argtypes = [vp, vp]
    L.ttrl_sim_set_reset_pool.argtypes = [vp, i32, vp, vp, vp, vp]
    L.ttrl_sim_set_autoreset.argtypes = [vp, i32]
    L.ttrl_sim_set_reset_params.argtypes = [vp, C.POINTER(abi.ResetParams)]
    L.ttrl_sim_reset.argtypes = [vp, vp, vp]
    L.ttrl_sim_seed.argtypes = [vp, u64, i64]
    L.ttrl_sim_inject_spawn.argtypes = [vp, vp]
    L.ttrl_sim_inject_shuffle.argtypes = [vp, vp]
    L.ttrl_sim_spawn_accepted.argtypes = [vp, vp]
    L.ttrl_sim_substep.argtypes = [vp, vp, vp]
    L.ttrl_sim_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.ttrl_sim_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ttrl_sim_host_buffers.argtypes = [vp] + [C.POINTER(vp)] * 5
    L.ttrl_sim_step_pinned.argtypes = [vp, i32]
    L.ttrl_sim_host_agent_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.ttrl_sim_agent_outputs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.ttrl_sim_set_agent_outputs.argtypes = [vp, vp, vp]
    L.ttrl_sim_set_info_outputs.argtypes = [vp, vp, vp]
    L.ttrl_sim_host_info_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.ttrl_sim_observe.argtypes = [vp, vp, vp]
    L.ttrl_sim_spawn.argtypes = [vp, vp, dbl, dbl, dbl, dbl, i32, vp]
    L.ttrl_sim_read_stats.argtypes = [vp, C.POINTER(abi.EpisodeStats), i32]
    L.ttrl_qnet_create.argtypes = [C.POINTER(abi.QnetDesc), vp, i64, i32, C.POINTER(vp)]
    L.ttrl_qnet_destroy.argtypes = [vp]
    L.ttrl_qnet_set_weights.argtypes = [vp, vp, i64, i32, vp]
    L.ttrl_qnet_set_mode.argtypes = [vp, i32]
    L.ttrl_qnet_act.argtypes = [vp, vp, i32, dbl, u64, u64, vp, vp, vp]
    L.ttrl_qnet_act_injected.argtypes = [vp, vp, i32, dbl, vp, vp, vp, vp]
    L.ttrl_qnet_launch_count.argtypes = [vp]
    L.ttrl_qnet_launch_count.restype = i64
    L.ttrl_qnet_weights_dev.argtypes = [vp]
    L.ttrl_qnet_weights_dev.restype = vp
    L.ttrl_dqn_create.argtypes = [C.POINTER(abi.DqnDesc), i32, C.POINTER(vp)]
    L.ttrl_dqn_destroy.argtypes = [vp]
    L.ttrl_dqn_num_params.argtypes = [vp]
    L.ttrl_dqn_num_params.restype = i64
    L.ttrl_dqn_launch_count.argtypes = [vp]
    L.ttrl_dqn_launch_count.restype = i64
    L.ttrl_dqn_grad.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ttrl_dqn_adam.argtypes = [vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, dbl, dbl, dbl, dbl, vp, vp]
    if L.ttrl_abi_version() != abi.ABI_VERSION:
        raise TTRLError(f"{LIB_PATH} has ABI version {L.ttrl_abi_version()}, this package binds version {abi.ABI_VERSION}: rebuild it")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise TTRLError(lib().ttrl_last_error().decode())
