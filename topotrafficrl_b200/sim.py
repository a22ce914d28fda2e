"""Thin object wrapper over the C ABI (``include/ttrl_b200.h``): one :class:`Sim` = one ``ttrl_sim``.

numpy in / numpy out for state resync and the host-buffer step; raw device pointers (e.g. from torch tensors)
for the resident step.  No simulation logic lives here."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import abi
from ._lib import check, lib
from .road import NetworkTable
from .state import SimState


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p)


class Sim:
    def __init__(self, cfg: abi.Config, table: NetworkTable, num_envs: int, vcap: int, device: int = 0,
                 spawn_routes=None) -> None:
        self._L = lib()
        self.cfg, self.table = cfg, table
        self.num_envs, self.vcap, self.device = int(num_envs), int(vcap), int(device)
        self.linear = cfg.vehicle_model == abi.VEHICLE_LINEAR  # LinearVehicle traffic: the state carries SimState.lin
        h = C.c_void_p()
        check(self._L.ttrl_sim_create(C.byref(cfg), C.cast(table.lanes, C.c_void_p), C.cast(table.roads, C.c_void_p),
                                      _p(table.node_first), _p(table.node_roads), self.num_envs, self.vcap, self.device,
                                      C.byref(h)))
        self._h = h
        self.obs_size = self._L.ttrl_sim_obs_size(self._h)       # floats per env: K observations
        self.num_agents = self._L.ttrl_sim_num_agents(self._h)  # K = controlled vehicles per env
        if spawn_routes is not None:
            sl, rl, rr = (np.ascontiguousarray(a, dtype=np.int32) for a in spawn_routes)
            check(self._L.ttrl_sim_set_spawn_routes(self._h, _p(sl), _p(rl), _p(rr)))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.ttrl_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ---------------------------------------------------------------------------------------
    def set_state(self, st: SimState) -> None:
        st = st.contiguous()
        assert st.num_envs == self.num_envs and st.vcap == self.vcap, (st.num_envs, st.vcap, self.num_envs, self.vcap)
        check(self._L.ttrl_sim_set_state(self._h, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d)))
        if self.linear:
            if st.lin is None:
                raise ValueError("this sim runs LinearVehicle traffic: the state needs its parameter block (SimState.lin)")
            check(self._L.ttrl_sim_set_linear_params(self._h, _p(st.lin)))

    def get_state(self) -> SimState:
        st = SimState.zeros(self.num_envs, self.vcap, linear=self.linear)
        check(self._L.ttrl_sim_get_state(self._h, _p(st.veh_d), _p(st.veh_i), _p(st.env_i), _p(st.env_d)))
        if self.linear:
            check(self._L.ttrl_sim_get_linear_params(self._h, _p(st.lin)))
        return st

    def set_reset_pool(self, pool: SimState) -> None:
        pool = pool.contiguous()
        assert pool.vcap == self.vcap
        check(self._L.ttrl_sim_set_reset_pool(self._h, pool.num_envs, _p(pool.veh_d), _p(pool.veh_i), _p(pool.env_i), _p(pool.env_d)))
        if self.linear:
            if pool.lin is None:
                raise ValueError("this sim runs LinearVehicle traffic: the pool needs its parameter block (SimState.lin)")
            check(self._L.ttrl_sim_set_reset_pool_linear_params(self._h, _p(pool.lin)))

    def set_autoreset(self, mode) -> None:
        """``False``/0: off; ``True``/1: restart finished envs from the reset pool; 2 or ``"device"``: device-side reset."""
        if mode == "device":
            mode = abi.AUTORESET_DEVICE
        elif mode == "device-async":
            mode = abi.AUTORESET_DEVICE_ASYNC
        check(self._L.ttrl_sim_set_autoreset(self._h, int(mode)))

    def set_reset_params(self, params: abi.ResetParams) -> None:
        check(self._L.ttrl_sim_set_reset_params(self._h, C.byref(params)))

    def reset_device(self, mask_ptr: Optional[int] = None, stream: int = 0) -> None:
        """Fresh episodes generated on the device for every env (or the envs of a uint8 device mask)."""
        check(self._L.ttrl_sim_reset(self._h, mask_ptr, stream))

    def seed(self, seed: int, first_global_env: int = 0) -> None:
        check(self._L.ttrl_sim_seed(self._h, int(seed), int(first_global_env)))

    # ---- parity hooks --------------------------------------------------------------------------------
    def inject_spawn(self, draws) -> None:
        check(self._L.ttrl_sim_inject_spawn(self._h, C.cast(draws, C.c_void_p) if draws is not None else None))

    def inject_shuffle(self, inv_perm: Optional[np.ndarray]) -> None:
        a = None if inv_perm is None else np.ascontiguousarray(inv_perm, dtype=np.int32)
        check(self._L.ttrl_sim_inject_shuffle(self._h, _p(a)))

    def spawn_accepted(self) -> np.ndarray:
        out = np.zeros(self.num_envs, np.int32)
        check(self._L.ttrl_sim_spawn_accepted(self._h, _p(out)))
        return out

    def spawn(self, draws, longitudinal: float, position_deviation: float = 1.0, speed_deviation: float = 1.0,
              spawn_probability: float = 0.6, go_straight: bool = False) -> np.ndarray:
        out = np.zeros(self.num_envs, np.int32)
        check(self._L.ttrl_sim_spawn(self._h, C.cast(draws, C.c_void_p), longitudinal, position_deviation, speed_deviation,
                                     spawn_probability, int(go_straight), _p(out)))
        return out

    # ---- stepping ------------------------------------------------------------------------------------
    def substep_ptr(self, actions_ptr: Optional[int], stream: int = 0) -> None:
        check(self._L.ttrl_sim_substep(self._h, actions_ptr, stream))

    def step_ptr(self, actions_ptr, obs_ptr, reward_ptr, term_ptr, trunc_ptr, stream: int = 0) -> None:
        check(self._L.ttrl_sim_step(self._h, actions_ptr, obs_ptr, reward_ptr, term_ptr, trunc_ptr, stream))

    def observe_ptr(self, obs_ptr, stream: int = 0) -> None:
        check(self._L.ttrl_sim_observe(self._h, obs_ptr, stream))

    def _pinned_views(self):
        if getattr(self, "_pinned", None) is None:
            ptrs = [C.c_void_p() for _ in range(5)]
            check(self._L.ttrl_sim_host_buffers(self._h, *[C.byref(p) for p in ptrs]))
            E, K = self.num_envs, self.num_agents

            def view(p, ctype, n, dtype):
                return np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,)).view(dtype)

            ar, at = C.c_void_p(), C.c_void_p()
            check(self._L.ttrl_sim_host_agent_buffers(self._h, C.byref(ar), C.byref(at)))
            self._pinned_agents = (view(ar, C.c_float, E * K, np.float32).reshape(E, K), view(at, C.c_uint8, E * K, np.uint8).reshape(E, K))
            self._pinned = (view(ptrs[0], C.c_int32, E * K, np.int32), view(ptrs[1], C.c_float, E * self.obs_size, np.float32).reshape(E, self.obs_size),
                            view(ptrs[2], C.c_float, E, np.float32), view(ptrs[3], C.c_uint8, E, np.uint8), view(ptrs[4], C.c_uint8, E, np.uint8))
        return self._pinned

    def step_host(self, actions: Optional[np.ndarray], copy: bool = True):
        """One env.step() for all envs from HOST buffers (H2D, kernel, D2H inside the call).

        ``copy=True`` returns fresh arrays (like the reference's ``step``).  ``copy=False`` returns views of the
        library's page-locked staging buffers: no host copy at all, but the arrays are overwritten by the next call."""
        a_pin, obs, reward, term, trunc = self._pinned_views()
        if actions is not None:
            a_pin[:] = np.asarray(actions, dtype=np.int32).reshape(-1)
        check(self._L.ttrl_sim_step_pinned(self._h, int(actions is not None)))
        if copy:
            return obs.copy(), reward.copy(), term.copy(), trunc.copy()
        return obs, reward, term, trunc

    def agent_outputs_host(self, copy: bool = True):
        """info["agents_rewards"] / info["agents_terminated"] of the last ``step_host`` call: float32 [E, K], uint8 [E, K]
        (filled by the library at every ``step_host`` call)."""
        self._pinned_views()
        r, t = self._pinned_agents
        return (r.copy(), t.copy()) if copy else (r, t)

    def agent_outputs_ptr(self):
        """Device pointers (agent_reward float32[E*K], agent_terminated uint8[E*K]) of the last ``step_ptr`` call."""
        r, t = C.c_void_p(), C.c_void_p()
        check(self._L.ttrl_sim_agent_outputs(self._h, C.byref(r), C.byref(t)))
        return r.value, t.value

    def set_agent_outputs_ptr(self, agent_reward_ptr, agent_terminated_ptr) -> None:
        """Let ``step_ptr`` write info["agents_rewards"] / ["agents_terminated"] into caller-owned device buffers
        (float32[E*K], uint8[E*K])."""
        check(self._L.ttrl_sim_set_agent_outputs(self._h, agent_reward_ptr, agent_terminated_ptr))

    def set_info_outputs_ptr(self, info_ptr, final_obs_ptr) -> None:
        """Let ``step_ptr`` write the batched ``info`` (float64 ``[abi.NINFO, E]``: speed, crashed, the four ``rewards`` entries)
        and gymnasium's ``final_observation`` (float32 ``[E, obs_size]``, rows of the envs that finished) into caller-owned
        device buffers; ``None`` switches an output off."""
        check(self._L.ttrl_sim_set_info_outputs(self._h, info_ptr, final_obs_ptr))

    def host_info(self, copy: bool = True):
        """``(info float64 [abi.NINFO, E], final_obs float32 [E, obs_size])`` of the last ``step_host`` call (page-locked
        staging buffers; the first call switches the two outputs on for the host-buffer path)."""
        if getattr(self, "_pinned_info", None) is None:
            a, b = C.c_void_p(), C.c_void_p()
            check(self._L.ttrl_sim_host_info_buffers(self._h, C.byref(a), C.byref(b)))
            E = self.num_envs
            info = np.ctypeslib.as_array(C.cast(a, C.POINTER(C.c_double)), shape=(abi.NINFO * E,)).reshape(abi.NINFO, E)
            fo = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_float)), shape=(E * self.obs_size,)).reshape(E, self.obs_size)
            self._pinned_info = (info, fo)
        info, fo = self._pinned_info
        return (info.copy(), fo.copy()) if copy else (info, fo)

    def stats(self, reset: bool = False) -> abi.EpisodeStats:
        out = abi.EpisodeStats()
        check(self._L.ttrl_sim_read_stats(self._h, C.byref(out), int(reset)))
        return out

    @property
    def launch_count(self) -> int:
        return int(self._L.ttrl_sim_launch_count(self._h))
