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

The callers and data formats either side of the planner live here too, as host code with the reference's
interfaces: `BaseSampler` (base.py: the single-environment loop), `Sampler` (sampler.py: the rollout loop that owns the state above and records paths -- it plans through a
`PlannerSession` when the policy's dynamics model has an engine, and through `policy.get_actions` with NumPy state
otherwise), `IterativeEnvExecutor` (vectorized_env_executor.py:7-69), `rollout_multi` / `context_rollout_multi`
(samplers/utils.py: the trainer's evaluation rollouts, same state) and `ModelSampleProcessor`
(model_sample_processor.py: paths -> the arrays `fit()` takes).  tests/test_samplers.py replays scenarios recorded from
the unmodified reference classes (tests/golden/make_sampler_golden.py) through them and compares bit for bit.
"""
import copy
import ctypes as C
import time

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
        if cfg.world > 1 and not dynamics_model.sharded_planner().fused:
            raise CadmError("PlannerSession at world > 1 needs the fused peer-memory exchange (GPUs of one node, CUDA IPC)")
        # the session state (warm start, histories, counters) lives in the ENGINE, once: a new session takes it over and the
        # previous one must not touch it any more
        eng._session_token = self._token = object()
        self.reset()

    def _own(self):
        if getattr(self.engine, "_session_token", None) is not self._token:
            raise CadmError("this PlannerSession is stale: another session was created on the same dynamics model and owns the "
                            "engine's session state now (one live session per model)")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.engine.device).cuda_stream)

    def reset(self, idx=None):
        """Clear the warm start, the history and the counters of every environment (idx=None) or of the given ones."""
        e = self.engine
        self._own()
        mask = None
        with torch.cuda.device(e.device):
            torch.cuda.current_stream(e.device).synchronize()          # observe() may still be reading the pinned mask
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
        self._own()
        obs = np.asarray(obses, dtype=np.float32)
        if obs.shape != (self.m, self.obs_dim):
            raise ValueError(f"obses must be [{self.m}, {self.obs_dim}], got {obs.shape}")
        self._obs.numpy()[...] = obs.reshape(-1)
        seed = self.model._next_seed() if seed is None else int(seed)
        with torch.cuda.device(e.device):
            e._chk(e.lib.cadm_session_act(e._h, self.m, C.c_void_p(self._obs.data_ptr()), C.c_uint64(seed),
                                          C.c_void_p(self._act.data_ptr()), self._stream()))
        return self._act.numpy().reshape(self.m, self.act_dim).copy()

    def observe(self, next_obses, dones=None, entry=None):
        """Append the transition to the history buffers and reset finished episodes (sampler.py:164-195); asynchronous.
        The history entry is the observation acted on, or next_obs - obs with state_diff, formed on the device from the
        float32 copies.  A host that holds float64 observations and wants the reference's rounding of the difference
        (subtract in float64, then round) passes it as `entry` [m, D]; it is stored as is."""
        e = self.engine
        self._own()
        nxt = np.asarray(next_obses if entry is None else entry, dtype=np.float32)
        if nxt.shape != (self.m, self.obs_dim):
            raise ValueError(f"next_obses / entry must be [{self.m}, {self.obs_dim}], got {nxt.shape}")
        with torch.cuda.device(e.device):
            torch.cuda.current_stream(e.device).synchronize()          # the previous observe() may still read the staging buffers
            self._next.numpy()[...] = nxt.reshape(-1)
            done_p = None
            if dones is not None:
                self._mask.numpy()[...] = np.asarray(dones, dtype=bool).astype(np.uint8)
                done_p = C.c_void_p(self._mask.data_ptr())
            mode = 2 if entry is not None else int(self.state_diff)
            e._chk(e.lib.cadm_session_observe(e._h, self.m, C.c_void_p(self._next.data_ptr()), done_p, mode, self._stream()))

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


# ------------------------------------------------------------------------------------------------ host-side loop

class HostPlannerState:
    """The same state as PlannerSession, in NumPy, for all environments at once (sampler.py:49-57, 94-97, 118-120,
    164-195).  `Sampler` always keeps one: it is what the recorded paths' cp_obs / cp_act come from; without an engine it
    is also what the policy is fed."""

    def __init__(self, num_envs, obs_dim, act_dim, history_length, state_diff, use_cem, horizon=None):
        self.m, self.D, self.A, self.K = num_envs, obs_dim, act_dim, history_length
        self.state_diff, self.use_cem = bool(state_diff), bool(use_cem)
        if self.use_cem:
            self.prev_sol = np.zeros((num_envs, horizon, act_dim))
            self.init_var = np.full((num_envs, horizon, act_dim), np.square(2) / 16)
        self.history_state = np.zeros((num_envs, obs_dim * history_length))
        self.history_act = np.zeros((num_envs, act_dim * history_length))
        self.counts = np.zeros(num_envs, dtype=np.int64)

    def reset_plans(self, idx=slice(None)):
        if self.use_cem:
            self.prev_sol[idx] = 0.

    def reset_history(self, idx=slice(None)):
        self.history_state[idx] = 0.
        self.history_act[idx] = 0.
        self.counts[idx] = 0

    def shift(self, cem_solutions):
        """Warm start for the next step: the plan moved one step ahead, zero at the end; returns the first actions."""
        self.prev_sol[:, :-1] = cem_solutions[:, 1:]
        self.prev_sol[:, -1:] = 0.
        return cem_solutions[:, 0].copy()

    def observe(self, obses, actions, next_obses, dones):
        """Append one transition per environment: slot `count` while the buffer fills, then slide left by one entry.
        Finished episodes are cleared afterwards (their counters restart at 0, the others advance)."""
        D, A, K = self.D, self.A, self.K
        entry = (next_obses - obses) if self.state_diff else obses
        hs = self.history_state.reshape(self.m, K, D)
        ha = self.history_act.reshape(self.m, K, A)
        full = self.counts >= K
        rows = np.flatnonzero(full)
        hs[rows, :-1] = hs[rows, 1:]
        ha[rows, :-1] = ha[rows, 1:]
        slot = np.where(full, K - 1, self.counts)
        every = np.arange(self.m)
        hs[every, slot] = entry
        ha[every, slot] = actions
        dones = np.asarray(dones, dtype=bool)
        self.counts += 1
        self.reset_history(dones)
        return entry


class IterativeEnvExecutor:
    """num_rollouts deep copies of one environment stepped one after the other (vectorized_env_executor.py:7-69): an
    environment that reports done, or reaches max_path_length steps, is reset at once and its observation in the returned
    list is the first one of the new episode."""

    def __init__(self, env, num_rollouts, max_path_length):
        self._num_envs = num_rollouts
        self.envs = [copy.deepcopy(env) for _ in range(num_rollouts)]
        self.ts = np.zeros(num_rollouts, dtype='int')
        self.max_path_length = max_path_length

    @property
    def num_envs(self):
        return self._num_envs

    def reset(self):
        self.ts[:] = 0
        return [env.reset() for env in self.envs]

    def step(self, actions):
        assert len(actions) == self.num_envs
        obs, rewards, dones, env_infos = [], [], [], []
        for env, a in zip(self.envs, actions):
            o, r, d, info = env.step(a)
            obs.append(o)
            rewards.append(r)
            dones.append(d)
            env_infos.append(info)
        self.ts += 1
        dones = np.logical_or(self.ts >= self.max_path_length, np.asarray(dones))
        for i in np.flatnonzero(dones):
            obs[i] = self.envs[i].reset()
            self.ts[i] = 0
        return obs, rewards, dones, env_infos


def _stack_dicts(dicts):
    """List of (nested) dicts of arrays -> dict of stacked arrays (utils.py:152-174, max_path=None)."""
    out = {}
    for k in (dicts[0].keys() if dicts else ()):
        vals = [d[k] for d in dicts]
        out[k] = _stack_dicts(vals) if isinstance(vals[0], dict) else np.asarray(vals)
    return out


class BaseSampler:
    """cadm/samplers/base.py:9-112 -- the single-environment sampler: one environment, `policy.get_action(obs)` once per
    step (no warm start, no history: it serves random-shooting policies and random exploration), paths cut at `done` or
    after `max_path_length` steps, until num_rollouts * max_path_length steps of FINISHED paths are collected."""

    def __init__(self, env, policy, num_rollouts, max_path_length):
        assert hasattr(env, 'reset') and hasattr(env, 'step')
        self.env, self.policy, self.max_path_length = env, policy, max_path_length
        self.total_samples = num_rollouts * max_path_length
        self.total_timesteps_sampled = 0

    def _act(self, obs, random):
        if random:
            return self.env.action_space.sample(), {}
        action, info = self.policy.get_action(obs)
        return (action[0] if action.ndim == 2 else action), info

    def obtain_samples(self, log=False, log_prefix='', random=False):
        paths, collected = [], 0
        keys = ("observations", "actions", "rewards", "dones", "env_infos", "agent_infos")
        running = {k: [] for k in keys}
        obs, steps = np.asarray(self.env.reset()), 0
        while collected < self.total_samples:
            action, agent_info = self._act(obs, random)
            next_obs, reward, done, env_info = self.env.step(action)
            steps += 1
            done = done or steps >= self.max_path_length
            if done:                                         # the stored transition keeps the pre-reset observation only
                next_obs, steps = self.env.reset(), 0
            if isinstance(reward, np.ndarray):
                reward = reward[0]
            for k, v in zip(keys, (obs, action, reward, done, env_info, agent_info)):
                running[k].append(v)
            if done:
                path = {k: np.asarray(running[k]) for k in keys[:4]}
                path["env_infos"], path["agent_infos"] = _stack_dicts(running["env_infos"]), _stack_dicts(running["agent_infos"])
                paths.append(path)
                collected += len(running["rewards"])
                running = {k: [] for k in keys}
            obs = next_obs
        self.total_timesteps_sampled += self.total_samples
        return paths


class Sampler:
    """cadm/samplers/sampler.py: `obtain_samples()` rolls num_rollouts environments until num_rollouts*max_path_length
    steps of FINISHED paths are collected and returns the list of paths (dicts with observations, actions, rewards, dones,
    env_infos, agent_infos, cp_obs, cp_act).  Same constructor and keywords; additions: `device_state` (None = use a
    PlannerSession whenever use_cem and the policy's dynamics model has an engine; False = feed policy.get_actions from
    NumPy state as the reference does) and `vec_env` (a ready executor; n_parallel > 1 worker processes are not part of
    this package, the environments are stepped in-process)."""

    def __init__(self, env, policy, num_rollouts, max_path_length, n_parallel=1, random_flag=False, use_cem=False,
                 horizon=None, context=False, state_diff=False, history_length=10, device_state=None, vec_env=None):
        assert hasattr(env, 'reset') and hasattr(env, 'step')
        self.env, self.policy = env, policy
        self.max_path_length = max_path_length
        self.total_samples = num_rollouts * max_path_length
        self.total_timesteps_sampled = 0
        self.n_parallel, self.random_flag = n_parallel, random_flag
        self.context, self.state_diff, self.history_length = context, state_diff, history_length
        self.discrete = len(env.action_space.shape) == 0
        self.act_dim = env.action_space.n if self.discrete else env.action_space.shape[0]
        self.vec_env = vec_env if vec_env is not None else IterativeEnvExecutor(env, num_rollouts, max_path_length)
        self.use_cem, self.horizon = use_cem, horizon
        self.state = HostPlannerState(num_rollouts, env.observation_space.shape[0], self.act_dim, history_length, state_diff,
                                      use_cem, horizon)
        model = getattr(policy, "dynamics_model", None)
        has_engine = getattr(model, "engine", None) is not None
        if device_state is None:
            device_state = bool(use_cem and has_engine)
        if device_state and not (use_cem and has_engine):
            raise CadmError("device_state needs use_cem=True and a policy whose dynamics_model owns a PlannerEngine")
        self.session = PlannerSession(model, num_rollouts, state_diff=state_diff) if device_state else None
        self.last_timing = {}

    # the reference exposes these two arrays; they live in the host state
    @property
    def prev_sol(self):
        return self.state.prev_sol

    @property
    def init_var(self):
        return self.state.init_var

    def reset_cem(self, idx):
        self.state.reset_plans(idx)

    def update_tasks(self):
        pass

    def _decide(self, obses, random):
        m, st = self.vec_env.num_envs, self.state
        if random:
            return np.stack([self.env.action_space.sample() for _ in range(m)], axis=0), {}
        if self.session is not None:
            return self.session.act(np.asarray(obses)), {}
        kw = dict(cp_obs=st.history_state, cp_act=st.history_act) if self.context else {}
        if self.use_cem:
            sols, infos = self.policy.get_actions(obses, init_mean=st.prev_sol, init_var=st.init_var, **kw)
            actions = st.shift(sols)
        else:
            actions, infos = self.policy.get_actions(obses, **kw)
        if self.discrete:
            actions = actions.reshape(-1)
        return actions, infos

    def obtain_samples(self, log=False, log_prefix='', random=False):
        m = self.vec_env.num_envs
        obses = np.asarray(self.vec_env.reset())
        self.obs_dim = obses.shape[1]
        # the reference clears the warm starts and allocates fresh history buffers at every call (:80-97)
        self.state = st = HostPlannerState(m, self.obs_dim, self.act_dim, self.history_length, self.state_diff, self.use_cem,
                                           self.horizon)
        if self.session is not None:
            self.session.reset()
        fields = ("observations", "actions", "rewards", "dones", "env_infos", "agent_infos", "cp_obs", "cp_act")
        running = [{k: [] for k in fields} for _ in range(m)]
        paths, n_samples, policy_time, env_time = [], 0, 0.0, 0.0
        while n_samples < self.total_samples:
            t0 = time.time()
            actions, agent_infos = self._decide(obses, random)
            t1 = time.time()
            next_obses, rewards, dones, env_infos = self.vec_env.step(actions)
            env_time += time.time() - t1
            policy_time += t1 - t0
            env_infos = env_infos if env_infos else [dict() for _ in range(m)]
            agent_infos = agent_infos if agent_infos else [dict() for _ in range(m)]
            obs_now, obs_next = np.asarray(obses, dtype=np.float64), np.asarray(next_obses, dtype=np.float64)
            acts = np.eye(self.act_dim)[np.asarray(actions)] if self.discrete else np.asarray(actions).reshape(m, -1)
            for i in range(m):
                r = rewards[i]
                rec = running[i]
                rec["observations"].append(obses[i])
                rec["actions"].append(acts[i])
                rec["rewards"].append(r[0] if isinstance(r, np.ndarray) else r)
                rec["dones"].append(dones[i])
                rec["env_infos"].append(env_infos[i])
                rec["agent_infos"].append(agent_infos[i])
                rec["cp_obs"].append(st.history_state[i].copy())
                rec["cp_act"].append(st.history_act[i].copy())
            entry = st.observe(obs_now, acts, obs_next, dones)
            if self.session is not None:
                self.session.observe(obs_next, dones, entry=entry)       # the float64 difference, rounded once
            for i in np.flatnonzero(dones):
                rec = running[i]
                paths.append(dict(
                    observations=np.asarray(rec["observations"]), actions=np.asarray(rec["actions"]),
                    rewards=np.asarray(rec["rewards"]), dones=np.asarray(rec["dones"]),
                    env_infos=_stack_dicts(rec["env_infos"]), agent_infos=_stack_dicts(rec["agent_infos"]),
                    cp_obs=np.asarray(rec["cp_obs"]), cp_act=np.asarray(rec["cp_act"])))
                n_samples += len(rec["rewards"])
                running[i] = {k: [] for k in fields}
                if not random:
                    st.reset_plans(i)
            obses = next_obses
        self.total_timesteps_sampled += self.total_samples
        self.last_timing = {log_prefix + "PolicyExecTime": policy_time, log_prefix + "EnvExecTime": env_time}
        return paths


# ------------------------------------------------------------------------------------------------ evaluation rollouts

def _evaluate(vec_env, policy, discrete, num_rollouts, test_total, state_diff, act_dim, use_cem, horizon, history_length,
              with_context, device_state):
    """Average undiscounted return of the first `test_total` episodes that finish (cadm/samplers/utils.py:5-41 and
    :44-124; the trainer's test phase, mb_trainer.py:250-280).  Same planner state as Sampler.obtain_samples."""
    m = vec_env.num_envs
    obses = np.asarray(vec_env.reset())
    K = history_length if with_context else 1
    st = HostPlannerState(m, obses.shape[1], act_dim, K, bool(state_diff), use_cem, horizon)
    model = getattr(policy, "dynamics_model", None)
    has_engine = getattr(model, "engine", None) is not None
    if device_state is None:
        device_state = bool(use_cem and has_engine)
    if device_state and not (use_cem and has_engine):
        raise CadmError("device_state needs use_cem=True and a policy whose dynamics_model owns a PlannerEngine")
    session = PlannerSession(model, m, state_diff=bool(state_diff)) if device_state else None
    finished, running, n_test = [], np.zeros(m), 0
    while n_test < test_total:
        if session is not None:
            actions = session.act(np.asarray(obses))
        else:
            kw = dict(cp_obs=st.history_state, cp_act=st.history_act) if with_context else {}
            if use_cem:
                sols, _ = policy.get_actions(obses, init_mean=st.prev_sol, init_var=st.init_var, **kw)
                actions = st.shift(sols)
            else:
                actions, _ = policy.get_actions(obses, **kw)
        if discrete:
            actions = actions.reshape(-1)
        next_obses, rewards, dones, _ = vec_env.step(actions)
        running += np.asarray(rewards, dtype=np.float64).reshape(m)
        if with_context or session is not None:
            acts = np.eye(act_dim)[np.asarray(actions)] if discrete else np.asarray(actions).reshape(m, -1)
            entry = st.observe(np.asarray(obses, dtype=np.float64), acts, np.asarray(next_obses, dtype=np.float64), dones)
            if session is not None:
                session.observe(np.asarray(next_obses), dones, entry=entry)
        for i in np.flatnonzero(dones):
            n_test += 1
            finished.append(running[i])
            running[i] = 0.
            st.reset_plans(i)
        obses = next_obses
    return np.average(finished)


def rollout_multi(vec_env, policy, discrete, animated=False, ignore_done=False, num_rollouts=10, test_total=20,
                  adapt_batch_size=None, state_diff=False, act_dim=None, use_cem=False, horizon=None, context=None,
                  history_length=None, device_state=None):
    """cadm/samplers/utils.py:5-41 (models without a context encoder)."""
    return _evaluate(vec_env, policy, discrete, num_rollouts, test_total, state_diff, act_dim, use_cem, horizon, history_length,
                     False, device_state)


def context_rollout_multi(vec_env, policy, discrete, animated=False, ignore_done=False, num_rollouts=10, test_total=20,
                          adapt_batch_size=None, state_diff=False, act_dim=None, use_cem=False, horizon=None, context=None,
                          history_length=None, device_state=None):
    """cadm/samplers/utils.py:44-124 (the K-step history feeds the context encoder)."""
    return _evaluate(vec_env, policy, discrete, num_rollouts, test_total, state_diff, act_dim, use_cem, horizon, history_length,
                     True, device_state)


# ------------------------------------------------------------------------------------------------ paths -> fit() arrays

def discount_cumsum(x, discount):
    """y[t] = x[t] + discount * y[t+1] (tensor_utils.py:217-221, there an IIR filter over the reversed sequence)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.empty_like(x)
    acc = np.zeros(x.shape[1:])
    for t in range(x.shape[0] - 1, -1, -1):
        acc = x[t] + discount * acc
        y[t] = acc
    return y


class ModelSampleProcessor:
    """cadm/samplers/model_sample_processor.py: finished paths -> the flat arrays of transitions the dynamics model's
    fit() takes.  With context=True also the future_length-step windows (concat_obs / concat_act / concat_next_obs
    [rows, dim*future_length]), their validity mask concat_bool and the history at each row.  Kept from the reference,
    because fit() results depend on them: paths shorter than future_length+1 are zero-padded IN PLACE (observations,
    actions, cp_obs, cp_act) after the single-step arrays were taken, so `observations` can have fewer rows than
    `concat_obs`; the mask of the FIRST row of every path is cleared (:76-87, `concat_bool[-0]`); `returns` covers every
    step of a path while `rewards` drops the last."""

    def __init__(self, discount=0.99, max_path_length=200, recurrent=False, context=False, writer=None, future_length=10):
        self.discount, self.max_path_length, self.recurrent = discount, max_path_length, recurrent
        self.context, self.writer, self.future_length = context, writer, future_length
        self.last_stats = {}

    def _join(self, arrays):
        return np.array(arrays) if self.recurrent else np.concatenate(arrays, axis=0)

    def _path_stats(self, paths, log_prefix):
        """The numbers base.py:222-245 sends to the logger, kept on the object instead (there is no logger here)."""
        undiscounted = [sum(p["rewards"]) for p in paths]
        self.last_stats = {log_prefix + "AverageDiscountedReturn": float(np.mean([p["returns"][0] for p in paths])),
                           log_prefix + "AverageReturn": float(np.mean(undiscounted)), log_prefix + "NumTrajs": len(paths),
                           log_prefix + "StdReturn": float(np.std(undiscounted)), log_prefix + "MaxReturn": float(np.max(undiscounted)),
                           log_prefix + "MinReturn": float(np.min(undiscounted))}

    def _windows(self, path):
        """One path -> (obs, act, next_obs, mask) windows; pads the path in place when it is too short."""
        F = self.future_length
        short = max(F + 1 - path["observations"].shape[0], 0)
        if short:
            for k in ("observations", "actions", "cp_obs", "cp_act"):
                path[k] = np.concatenate([path[k], np.zeros((short, path[k].shape[1]))], axis=0)
        O, A = path["observations"], path["actions"]
        T = O.shape[0] - 1

        def blocks(a, lo):
            # row t, block i  ->  a[t + lo + i], zero past the end of the path
            if F == 1:
                return a[lo:lo + T]
            ext = np.concatenate([a, np.zeros((F, a.shape[1]))], axis=0)
            win = np.lib.stride_tricks.sliding_window_view(ext, F, axis=0)[lo:lo + T]        # [T, dim, F]
            return np.ascontiguousarray(win.transpose(0, 2, 1)).reshape(T, F * a.shape[1])

        real = T - short                                                # transitions that really happened
        valid = np.clip(real - np.arange(T), 0, F)                      # how many of row t's F steps exist
        mask = (np.arange(F)[None, :] < valid[:, None]).astype(np.float64)
        mask[0] = 0.
        return blocks(O, 0), blocks(A, 0), blocks(O, 1), mask

    def process_samples(self, paths, log=False, log_prefix='', itr=None):
        assert len(paths) > 0
        for path in paths:
            path["returns"] = discount_cumsum(path["rewards"], self.discount)
        if log:
            self._path_stats(paths, log_prefix)
        data = dict(
            observations=self._join([p["observations"][:-1] for p in paths]),
            next_observations=self._join([p["observations"][1:] for p in paths]),
            actions=self._join([p["actions"][:-1] for p in paths]),
            timesteps=np.concatenate([np.arange(len(p["observations"]) - 1) for p in paths], axis=0),
            rewards=self._join([p["rewards"][:-1] for p in paths]),
            returns=self._join([p["returns"] for p in paths]))
        if self.context:
            wins = [self._windows(p) for p in paths]
            data.update(
                cp_observations=self._join([p["cp_obs"][:-1] for p in paths]),
                cp_actions=self._join([p["cp_act"][:-1] for p in paths]),
                concat_next_obs=self._join([w[2] for w in wins]), concat_obs=self._join([w[0] for w in wins]),
                concat_act=self._join([w[1] for w in wins]), concat_bool=self._join([w[3] for w in wins]))
        return data
