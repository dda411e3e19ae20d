"""PlannerEngine: the python face of the C-ABI handle.

PyTorch is used for device memory and streams only; every computation happens inside libcadm_b200.so.
All tensors handed to the engine are float32 CUDA tensors (int32 for discrete action ids).
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import CadmConfig, CadmError, check


@dataclass
class PlannerConfig:
    """Python mirror of CadmConfig (include/cadm_b200.h)."""
    env: str = "halfcheetah"
    obs_dim: int = 18
    proc_obs_dim: int = 18
    act_dim: int = 6
    ctx_dim: int = 0
    hist_len: int = 10
    hidden: int = 200
    n_hidden: int = 4
    enc_hidden: Sequence[int] = (256, 128, 64)
    ensemble: int = 5
    particles: int = 20
    candidates: int = 200
    horizon: int = 30
    m_max: int = 1
    deterministic: bool = False
    discrete: bool = False
    num_elites: int = 50        # cadm/dynamics/core/utils.py:111
    cem_iters: int = 5          # :112
    alpha: float = 0.1          # :113
    precision: str = "tc3x"     # tensor cores, fp16 hi/lo split (parity mode); "fp32" = FFMA path
    rank: int = 0
    world: int = 1
    context_layout: str = "reference"
    max_torque: float = 2.0

    def to_c(self) -> CadmConfig:
        c = CadmConfig()
        c.struct_size = C.sizeof(CadmConfig)
        c.env_id = _lib.ENV_IDS[self.env]
        c.obs_dim, c.proc_obs_dim, c.act_dim = self.obs_dim, self.proc_obs_dim, self.act_dim
        c.ctx_dim, c.hist_len, c.hidden, c.n_hidden = self.ctx_dim, self.hist_len, self.hidden, self.n_hidden
        eh = list(self.enc_hidden) + [0, 0, 0]
        c.enc_hidden = (C.c_int32 * 3)(*eh[:3])
        c.ensemble, c.particles, c.candidates, c.horizon = self.ensemble, self.particles, self.candidates, self.horizon
        c.m_max = self.m_max
        c.deterministic, c.discrete = int(self.deterministic), int(self.discrete)
        c.num_elites, c.cem_iters, c.alpha = self.num_elites, self.cem_iters, self.alpha
        c.precision = _lib.PRECISIONS[self.precision]
        c.rank, c.world = self.rank, self.world
        c.context_layout = _lib.CTX_LAYOUTS[self.context_layout]
        c.max_torque = self.max_torque
        return c


class _CudaView:
    """Zero-copy torch view of engine-owned device memory via __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t, device):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32).contiguous()


class PlannerEngine:
    def __init__(self, cfg: PlannerConfig, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise CadmError("cadm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            ccfg = cfg.to_c()
            check(None, self.lib.cadm_plan_create(C.byref(ccfg), C.byref(self._h)))
        # tuning knobs of the tensor-core path (diagnostics / sweeps; the defaults are what bench.py measures)
        for opt, env in (("tc_variant", "CADM_TC_VARIANT"), ("tcs_rows", "CADM_TCS_ROWS"), ("tcs_kps", "CADM_TCS_KPS"),
                         ("tcs_skew", "CADM_TCS_SKEW")):
            if os.environ.get(env):
                self.set_option(opt, int(os.environ[env]))
        self.In = cfg.proc_obs_dim + cfg.act_dim + cfg.ctx_dim
        self.n_local = cfg.candidates // cfg.world
        self._keep = []      # tensors the engine borrowed asynchronously
        # pinned staging for the host entry point
        K = cfg.hist_len if cfg.ctx_dim > 0 else 1
        mm, hA = cfg.m_max, cfg.horizon * cfg.act_dim
        pin = lambda n: torch.empty(n, dtype=torch.float32).pin_memory()
        self._pin = dict(obs=pin(mm * cfg.obs_dim), cp_obs=pin(mm * cfg.obs_dim * K), cp_act=pin(mm * cfg.act_dim * K),
                         mean=pin(mm * hA), var=pin(mm * hA), act=pin(mm * hA))

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.cadm_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, code):
        check(self._h, code)

    # ------------------------------------------------------------------ parameters
    def set_weights(self, W: List, b: List, max_logvar, min_logvar):
        """W/b: n_hidden hidden layers then output_mu, output_logvar ([E,in,out] / [E,1,out])."""
        Wt = [_f32(w, self.device) for w in W]
        bt = [_f32(x, self.device) for x in b]
        mx, mn = _f32(np.reshape(max_logvar, -1), self.device), _f32(np.reshape(min_logvar, -1), self.device)
        n = len(Wt)
        Wp = (C.c_void_p * n)(*[w.data_ptr() for w in Wt])
        bp = (C.c_void_p * n)(*[x.data_ptr() for x in bt])
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_plan_set_weights(self._h, Wp, bp, n, _ptr(mx), _ptr(mn), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()     # sources may be freed after return

    def set_encoder(self, W: List, b: List):
        Wt = [_f32(w, self.device) for w in W]
        bt = [_f32(x, self.device) for x in b]
        n = len(Wt)
        Wp = (C.c_void_p * n)(*[w.data_ptr() for w in Wt])
        bp = (C.c_void_p * n)(*[x.data_ptr() for x in bt])
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_plan_set_encoder(self._h, Wp, bp, n, self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    def set_norm(self, obs_mean, obs_std, act_mean, act_std, delta_mean, delta_std,
                 cp_obs_mean=None, cp_obs_std=None, cp_act_mean=None, cp_act_std=None):
        ts = [_f32(None if a is None else np.asarray(a, dtype=np.float32).reshape(-1), self.device) for a in (
            obs_mean, obs_std, act_mean, act_std, delta_mean, delta_std, cp_obs_mean, cp_obs_std, cp_act_mean, cp_act_std)]
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_plan_set_norm(self._h, *[_ptr(t) for t in ts], self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------ compute
    def encode_context(self, cp_obs, cp_act) -> torch.Tensor:
        cp_obs, cp_act = _f32(cp_obs, self.device), _f32(cp_act, self.device)
        m = cp_obs.shape[0]
        out = torch.empty((self.cfg.ensemble, m, self.cfg.ctx_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_encode_context(self._h, m, _ptr(cp_obs), _ptr(cp_act), _ptr(out), self._stream()))
        return out

    def predict(self, obs, act, ctx=None, eps=None, seed=0):
        """obs [E,B,D], act [E,B,A] -> (next_obs, mu, logvar) [E,B,D]."""
        obs, act, ctx, eps = (_f32(t, self.device) for t in (obs, act, ctx, eps))
        E, B, D = obs.shape
        outs = [torch.empty((E, B, D), dtype=torch.float32, device=self.device) for _ in range(3)]
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_predict(self._h, B, _ptr(obs), _ptr(act), _ptr(ctx), _ptr(eps), C.c_uint64(seed),
                                            *[_ptr(o) for o in outs], self._stream()))
        return tuple(outs)

    def rollout(self, obs, actions, ctx_raw=None, eps=None, seed=0, it=0, trace=False):
        """obs [m,D], actions [m,n_local,h,A] -> particle returns [m,n_local,p] (+ states [h,m,n_local,p,D])."""
        obs, actions, ctx_raw, eps = (_f32(t, self.device) for t in (obs, actions, ctx_raw, eps))
        m = obs.shape[0]
        c = self.cfg
        pr = torch.empty((m, self.n_local, c.particles), dtype=torch.float32, device=self.device)
        st = torch.empty((c.horizon, m, self.n_local, c.particles, c.obs_dim), dtype=torch.float32,
                         device=self.device) if trace else None
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_rollout(self._h, m, it, _ptr(obs), _ptr(actions), _ptr(ctx_raw), _ptr(eps),
                                            C.c_uint64(seed), _ptr(pr), _ptr(st), self._stream()))
        return pr, st

    def plan_cem(self, obs, init_mean, init_var, cp_obs=None, cp_act=None, seed=0, z=None, eps=None, logs=True):
        """One CEM decision on device tensors (world == 1).  Returns dict(mean, var, returns, elites)."""
        c = self.cfg
        obs, init_mean, init_var, cp_obs, cp_act, z, eps = (
            _f32(t, self.device) for t in (obs, init_mean, init_var, cp_obs, cp_act, z, eps))
        m = obs.shape[0]
        mean = torch.empty((m, c.horizon, c.act_dim), dtype=torch.float32, device=self.device)
        var = torch.empty_like(mean)
        rets = torch.empty((c.cem_iters, m, c.candidates), dtype=torch.float32, device=self.device) if logs else None
        el = torch.empty((c.cem_iters, m, c.num_elites), dtype=torch.int32, device=self.device) if logs else None
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_plan_cem(self._h, m, _ptr(obs), _ptr(cp_obs), _ptr(cp_act), _ptr(init_mean),
                                             _ptr(init_var), C.c_uint64(seed), _ptr(z), _ptr(eps), _ptr(mean), _ptr(var),
                                             _ptr(rets), _ptr(el), self._stream()))
        self._keep = [obs, init_mean, init_var, cp_obs, cp_act, z, eps]
        return dict(mean=mean, var=var, returns=rets, elites=el)

    def plan_cem_into(self, obs, init_mean, init_var, out_mean, out_var, cp_obs=None, cp_act=None, seed=0):
        """The same decision for callers that keep everything on the device: float32 CUDA tensors in, results written into
        the caller's `out_mean` / `out_var` [m, h, A]; no conversions, no allocations, no logs (one C call per decision)."""
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_plan_cem(self._h, obs.shape[0], _ptr(obs), _ptr(cp_obs), _ptr(cp_act), _ptr(init_mean),
                                             _ptr(init_var), C.c_uint64(seed), None, None, _ptr(out_mean), _ptr(out_var), None, None,
                                             self._stream()))

    def plan_cem_host(self, obs: np.ndarray, init_mean: np.ndarray, init_var: np.ndarray, cp_obs=None, cp_act=None,
                      seed=0) -> np.ndarray:
        """The get_action() path: NumPy in, clipped plan [m,h,A] out; H2D/D2H copies inside the call."""
        c = self.cfg
        m = obs.shape[0]
        if m > c.m_max:
            raise CadmError(f"batch of {m} environments exceeds m_max={c.m_max}")
        hA = c.horizon * c.act_dim

        def stage(name, arr, n):
            buf = self._pin[name][:n]
            buf.numpy()[...] = np.asarray(arr, dtype=np.float32).reshape(-1)
            return C.c_void_p(buf.data_ptr())

        p_obs = stage("obs", obs, m * c.obs_dim)
        p_mean = stage("mean", init_mean, m * hA)
        p_var = stage("var", init_var, m * hA)
        p_co = p_ca = None
        if c.ctx_dim > 0:
            p_co = stage("cp_obs", cp_obs, m * c.obs_dim * c.hist_len)
            p_ca = stage("cp_act", cp_act, m * c.act_dim * c.hist_len)
        out = self._pin["act"][: m * hA]
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_plan_cem_host(self._h, m, p_obs, p_co, p_ca, p_mean, p_var, C.c_uint64(seed),
                                                  C.c_void_p(out.data_ptr()), self._stream()))
        return out.numpy().reshape(m, c.horizon, c.act_dim).copy()

    def plan_rs(self, obs, cp_obs=None, cp_act=None, seed=0, u=None, eps=None):
        """Random shooting.  Returns dict(action, returns, best)."""
        c = self.cfg
        obs, cp_obs, cp_act, eps = (_f32(t, self.device) for t in (obs, cp_obs, cp_act, eps))
        m = obs.shape[0]
        u_f = u_i = None
        if u is not None:
            if c.discrete:
                u_i = torch.as_tensor(np.asarray(u), dtype=torch.int32, device=self.device).contiguous()
            else:
                u_f = _f32(u, self.device)
        rets = torch.empty((m, c.candidates), dtype=torch.float32, device=self.device)
        best = torch.empty((m,), dtype=torch.int32, device=self.device)
        act_f = torch.empty((m, c.act_dim), dtype=torch.float32, device=self.device) if not c.discrete else None
        act_i = torch.empty((m,), dtype=torch.int32, device=self.device) if c.discrete else None
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_plan_rs(self._h, m, _ptr(obs), _ptr(cp_obs), _ptr(cp_act), C.c_uint64(seed), _ptr(u_f),
                                            _ptr(u_i), _ptr(eps), _ptr(act_f), _ptr(act_i), _ptr(rets), _ptr(best),
                                            self._stream()))
        self._keep = [obs, cp_obs, cp_act, eps, u_f, u_i]
        return dict(action=act_i if c.discrete else act_f, returns=rets, best=best)

    # ------------------------------------------------------------------ phase API (multi-rank)
    def cem_begin(self, obs, init_mean, init_var, cp_obs=None, cp_act=None):
        obs, init_mean, init_var, cp_obs, cp_act = (_f32(t, self.device) for t in (obs, init_mean, init_var, cp_obs, cp_act))
        self._m = obs.shape[0]
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_cem_begin(self._h, self._m, _ptr(obs), _ptr(cp_obs), _ptr(cp_act), _ptr(init_mean),
                                              _ptr(init_var), self._stream()))
        self._keep = [obs, init_mean, init_var, cp_obs, cp_act]

    def cem_rollout(self, it, seed=0, z=None, eps=None):
        z, eps = _f32(z, self.device), _f32(eps, self.device)
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_cem_rollout(self._h, it, C.c_uint64(seed), _ptr(z), _ptr(eps), self._stream()))
        self._keep += [z, eps]

    def returns_buffer(self) -> torch.Tensor:
        """[world, m, n_local] view of the engine's candidate-returns buffer (all-gather target)."""
        ptr = self.lib.cadm_cem_returns_buffer(self._h)
        return torch.as_tensor(_CudaView(ptr, (self.cfg.world, self._m, self.n_local)), device=self.device)

    def peer_export(self) -> bytes:
        """CUDA IPC handle of this rank's exchange block (cadm_peer_export)."""
        buf = C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_peer_export(self._h, buf))
        return buf.raw

    def peer_attach(self, handles: Sequence[bytes]):
        """Attach every rank's exchange block (rank order); afterwards the all-gather is fused into the rollout phase."""
        blob = b"".join(handles)
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_peer_attach(self._h, blob, len(handles)))

    @property
    def peers_enabled(self) -> bool:
        return bool(self.lib.cadm_peer_enabled(self._h))

    def cem_refit(self, it):
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_cem_refit(self._h, it, self._stream()))

    def cem_finish(self, logs=True):
        c = self.cfg
        m = self._m
        mean = torch.empty((m, c.horizon, c.act_dim), dtype=torch.float32, device=self.device)
        var = torch.empty_like(mean)
        rets = torch.empty((c.cem_iters, m, c.candidates), dtype=torch.float32, device=self.device) if logs else None
        el = torch.empty((c.cem_iters, m, c.num_elites), dtype=torch.int32, device=self.device) if logs else None
        with torch.cuda.device(self.device):
            self._chk(self.lib.cadm_cem_finish(self._h, _ptr(mean), _ptr(var), _ptr(rets), _ptr(el), self._stream()))
        return dict(mean=mean, var=var, returns=rets, elites=el)

    def set_precision(self, precision: str):
        self._chk(self.lib.cadm_set_precision(self._h, _lib.PRECISIONS[precision]))
        self.cfg.precision = precision

    def set_option(self, name: str, value: int):
        """cadm_set_option: "tc_variant" (0 auto / 1 row tiles / 2 swapped operands / 3 CTA pairs), "tcs_rows", "tcs_kps", "pdl",
        "trace", "env_offset", "peer_timeout_ms", "peer_clear_timeout" (include/cadm_b200.h)."""
        self._chk(self.lib.cadm_set_option(self._h, name.encode(), int(value)))

    # ------------------------------------------------------------------ instrumentation
    @property
    def launch_count(self) -> int:
        return int(self.lib.cadm_launch_count(self._h))

    @property
    def kernel_name(self) -> str:
        return self.lib.cadm_kernel_name(self._h).decode()

    def set_timing(self, on: bool):
        self._chk(self.lib.cadm_set_timing(self._h, int(on)))

    def debug_trace(self, steps=30) -> np.ndarray:
        """clock64 trace [steps, 32] of CTA 0 of the last tensor-core rollout (timing must be enabled)."""
        out = np.zeros((steps, 64), dtype=np.int64)
        self._chk(self.lib.cadm_debug_trace(self._h, out.ctypes.data_as(C.c_void_p), steps * 64))
        return out

    def last_rollout_ms(self) -> float:
        return float(self.lib.cadm_last_rollout_ms(self._h))


def selftest_tc_gemm(X: torch.Tensor, W: torch.Tensor, terms: int = 3) -> torch.Tensor:
    """out[128, N] = X[128, K] @ W[K, N] on the tensor-core path of the rollout kernel (device diagnostic)."""
    lib = _lib.load()
    X = X.to(dtype=torch.float32).contiguous()
    W = W.to(dtype=torch.float32).contiguous()
    assert X.is_cuda and W.is_cuda and X.shape[0] == 128 and X.shape[1] == W.shape[0]
    out = torch.empty((128, W.shape[1]), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(None, lib.cadm_selftest_tc_gemm(_ptr(X), _ptr(W), X.shape[1], W.shape[1], terms, _ptr(out),
                                              C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream)))
    return out


def selftest_tcs_gemm(X: torch.Tensor, W: torch.Tensor, kps: int = 2, terms: int = 3) -> torch.Tensor:
    """out[rows, N] = X[rows, K] @ W[K, N] on the swapped-operand tensor-core path (rollout_tcs.cu; device diagnostic)."""
    lib = _lib.load()
    X = X.to(dtype=torch.float32).contiguous()
    W = W.to(dtype=torch.float32).contiguous()
    assert X.is_cuda and W.is_cuda and X.shape[1] == W.shape[0]
    out = torch.empty((X.shape[0], W.shape[1]), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(None, lib.cadm_selftest_tcs_gemm(_ptr(X), _ptr(W), X.shape[0], X.shape[1], W.shape[1], kps, terms, _ptr(out),
                                               C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream)))
    return out
