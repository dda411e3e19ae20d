"""Sampler-side state on the device.

`cadm/samplers/sampler.py` keeps, in NumPy, everything a model-predictive controller carries from one control step to
the next: the warm-start plan `prev_sol` (lines 49-57, 118-120), the constant `init_var`, the K-step history buffers that
feed the CaDM context encoder and their fill counters (94-97, 164-178) and the per-episode resets (190-195); every step
it feeds them to `policy.get_actions` as host arrays.  `PlannerSession` moves that state into the engine
(`cadm_session_*` in include/cadm_b200.h): a control step is one H2D of the observations, one decision, one D2H of the
first actions -- the plans and histories never leave the GPU.

    session = PlannerSession(dynamics_model, num_envs)         # replaces prev_sol / init_var / history_state / history_act
    session.reset()                                            # sampler.py:80-82, 94-97
    while sampling:
        actions = session.act(obses)                           # sampler.py:107-120  -> [m, A], clipped
        next_obses, rewards, dones, infos = vec_env.step(actions)
        session.observe(next_obses, dones)                     # sampler.py:164-195
        obses = next_obses

The arithmetic is the engine's own `cadm_plan_cem`; `tests/test_gpu_envs.py::test_session_matches_host_loop` checks that
the actions equal the host-side loop's bit for bit.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import CadmError


class PlannerSession:
    def __init__(self, dynamics_model, num_envs, state_diff=None):
        eng = dynamics_model.engine
        cfg = eng.cfg
        if num_envs < 1 or num_envs > cfg.m_max:
            raise CadmError(f"num_envs={num_envs} exceeds the model's m_max={cfg.m_max}")
        if not getattr(dynamics_model, "use_cem", True):
            raise CadmError("PlannerSession plans with CEM: build the dynamics model with use_cem=True")
        self.model, self.engine, self.m = dynamics_model, eng, int(num_envs)
        self.state_diff = bool(getattr(dynamics_model, "state_diff", False)) if state_diff is None else bool(state_diff)
        self.obs_dim, self.act_dim, self.horizon = cfg.obs_dim, cfg.act_dim, cfg.horizon
        self.hist_len = cfg.hist_len if cfg.ctx_dim > 0 else 1
        pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory()
        self._obs = pin(self.m * self.obs_dim, torch.float32)
        self._next = pin(self.m * self.obs_dim, torch.float32)
        self._act = pin(self.m * self.act_dim, torch.float32)
        self._mask = pin(self.m, torch.uint8)
        self.reset()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.engine.device).cuda_stream)

    def reset(self, idx=None):
        """Clear the warm start, the history and the counters of every environment (idx=None) or of the given ones."""
        e = self.engine
        mask = None
        if idx is not None:
            self._mask.zero_()
            self._mask.numpy()[np.atleast_1d(idx)] = 1
            mask = C.c_void_p(self._mask.data_ptr())
        with torch.cuda.device(e.device):
            e._chk(e.lib.cadm_session_reset(e._h, self.m, mask, self._stream()))
            torch.cuda.current_stream(e.device).synchronize()

    def act(self, obses, seed=None) -> np.ndarray:
        """One decision for every environment from the stored warm start / history; returns the clipped first actions [m, A]
        and shifts the warm start (sampler.py:107-120)."""
        e = self.engine
        obs = np.asarray(obses, dtype=np.float32)
        if obs.shape != (self.m, self.obs_dim):
            raise ValueError(f"obses must be [{self.m}, {self.obs_dim}], got {obs.shape}")
        self._obs.numpy()[...] = obs.reshape(-1)
        seed = self.model._next_seed() if seed is None else int(seed)
        with torch.cuda.device(e.device):
            e._chk(e.lib.cadm_session_act(e._h, self.m, C.c_void_p(self._obs.data_ptr()), C.c_uint64(seed),
                                          C.c_void_p(self._act.data_ptr()), self._stream()))
        return self._act.numpy().reshape(self.m, self.act_dim).copy()

    def observe(self, next_obses, dones=None):
        """Append the transition to the history buffers and reset finished episodes (sampler.py:164-195); asynchronous."""
        e = self.engine
        nxt = np.asarray(next_obses, dtype=np.float32)
        if nxt.shape != (self.m, self.obs_dim):
            raise ValueError(f"next_obses must be [{self.m}, {self.obs_dim}], got {nxt.shape}")
        with torch.cuda.device(e.device):
            torch.cuda.current_stream(e.device).synchronize()          # the previous observe() may still read the staging buffers
            self._next.numpy()[...] = nxt.reshape(-1)
            done_p = None
            if dones is not None:
                self._mask.numpy()[...] = np.asarray(dones, dtype=bool).astype(np.uint8)
                done_p = C.c_void_p(self._mask.data_ptr())
            e._chk(e.lib.cadm_session_observe(e._h, self.m, C.c_void_p(self._next.data_ptr()), done_p, int(self.state_diff),
                                              self._stream()))

    def state(self):
        """Host copies of (prev_sol [m, h, A], history_state [m, D*K], history_act [m, A*K], state_counts [m])."""
        e = self.engine
        K = self.hist_len
        prev = np.empty((self.m, self.horizon, self.act_dim), np.float32)
        ho = np.empty((self.m, self.obs_dim * K), np.float32)
        ha = np.empty((self.m, self.act_dim * K), np.float32)
        cnt = np.empty((self.m,), np.int32)
        with torch.cuda.device(e.device):
            e._chk(e.lib.cadm_session_state(e._h, self.m, prev.ctypes.data_as(C.c_void_p), ho.ctypes.data_as(C.c_void_p),
                                            ha.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), self._stream()))
        return prev, ho, ha, cnt
