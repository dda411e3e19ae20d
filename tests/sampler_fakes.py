"""A deterministic stand-in environment and policy for the sampler / sample-processor tests.

Shared by tests/golden/make_sampler_golden.py (which drives the UNMODIFIED reference `Sampler` and `ModelSampleProcessor`
with them) and tests/test_samplers.py (which drives cadm_b200's classes with them and compares with the recorded output).
Nothing here comes from the reference: the environment is a linear map with a scripted episode length, the policy a
closed-form function of everything it is fed, so that any difference in what the sampler feeds shows up in the paths.
"""
import numpy as np


class _Box:
    def __init__(self, dim):
        self.shape = (dim,)
        self._rng = np.random.default_rng(99)

    def sample(self):
        return self._rng.uniform(-1.0, 1.0, self.shape)


class FakeEnv:
    """obs' = 0.9 obs + B act + drift(k); episode j of copy k ends (done=True) after `lengths[k][j]` steps."""
    _copies = 0

    def __init__(self, obs_dim=3, act_dim=2, lengths=((5, 2, 30), (30, 7), (9, 30, 30)), round32=False):
        self.obs_dim, self.act_dim, self.round32 = obs_dim, act_dim, round32
        self.observation_space, self.action_space = _Box(obs_dim), _Box(act_dim)
        self.lengths = lengths
        self.k, self.episode, self.t = 0, -1, 0
        self.B = np.cos(np.arange(act_dim * obs_dim, dtype=np.float64)).reshape(act_dim, obs_dim) * 0.3
        self.obs = np.zeros(obs_dim)

    def __deepcopy__(self, memo):
        # the executors clone the environment once per rollout slot; every clone gets its own index
        new = FakeEnv(self.obs_dim, self.act_dim, self.lengths, self.round32)
        new.k = FakeEnv._copies % len(self.lengths)
        FakeEnv._copies += 1
        return new

    def reset(self):
        self.episode += 1
        self.t = 0
        self.obs = self._q(np.sin(np.arange(self.obs_dim) + 1.0 + 3.0 * self.k + 0.5 * self.episode))
        return self.obs.copy()

    def _q(self, x):
        # round32: observations that float32 represents exactly (what a float32 simulator hands out)
        return x.astype(np.float32).astype(np.float64) if self.round32 else x

    def step(self, action):
        action = np.asarray(action, dtype=np.float64)
        self.obs = self._q(0.9 * self.obs + action @ self.B + 0.01 * (self.k + 1))
        self.t += 1
        plan = self.lengths[self.k]
        done = self.t >= plan[min(self.episode, len(plan) - 1)]
        reward = float(np.sum(self.obs) - 0.1 * np.sum(action ** 2))
        return self.obs.copy(), reward, bool(done), {"t": np.float64(self.t)}


class ScriptedPolicy:
    """get_actions() is a closed-form function of ALL its inputs and records them (copies) call by call."""

    def __init__(self, horizon, act_dim, use_cem):
        self.horizon, self.act_dim, self.use_cem = horizon, act_dim, use_cem
        self.calls = []

    def get_actions(self, observations, cp_obs=None, cp_act=None, init_mean=None, init_var=None):
        obs = np.asarray(observations, dtype=np.float64)
        rec = dict(obs=obs.copy())
        m = obs.shape[0]
        s = obs.sum(axis=1)
        if cp_obs is not None:
            rec["cp_obs"], rec["cp_act"] = np.array(cp_obs, copy=True), np.array(cp_act, copy=True)
            w = np.arange(1, cp_obs.shape[1] + 1) / cp_obs.shape[1]
            s = s + np.asarray(cp_obs) @ w - 0.5 * np.asarray(cp_act).sum(axis=1)
        if init_mean is not None:
            rec["init_mean"], rec["init_var"] = np.array(init_mean, copy=True), np.array(init_var, copy=True)
        self.calls.append(rec)
        grid = np.arange(self.horizon * self.act_dim, dtype=np.float64).reshape(1, self.horizon, self.act_dim)
        if self.use_cem:
            sol = np.sin(s[:, None, None] + 0.37 * grid) * 0.8 + 0.5 * np.asarray(init_mean) + 0.1 * np.asarray(init_var)
            return np.clip(sol, -1.0, 1.0), []
        return np.clip(np.sin(s[:, None] + 0.37 * grid[0, 0][None, :]), -1.0, 1.0), []


class ScriptedSinglePolicy:
    """get_action(obs) for the single-environment sampler: closed form of the observation, returned as a [1, A] batch (what
    MPCController.get_action hands back for one observation); records what it was given."""

    def __init__(self, act_dim):
        self.act_dim, self.calls = act_dim, []

    def get_action(self, observation):
        obs = np.asarray(observation, dtype=np.float64)
        self.calls.append(obs.copy())
        return np.sin(obs.sum() + 0.37 * np.arange(self.act_dim))[None, :], {"s": np.float64(obs.sum())}


class RecordingDynamicsModel:
    """get_action() records which positional arguments it was given (by tag) and returns a constant."""

    def __init__(self):
        self.calls = []

    def get_action(self, *args, **kwargs):
        self.calls.append((tuple(str(a[0]) if isinstance(a, tuple) else repr(a) for a in args), tuple(sorted(kwargs))))
        return np.zeros((1, 1))


class RewardEnv:
    """The only thing MPCController asks of its environment: a `reward` attribute (mpc_controller.py:34)."""
    action_space = _Box(2)

    def reward(self, obs, act, next_obs):
        return 0.0


def controller_dispatch_log(controller_cls):
    """Every (context, use_cem) combination through get_actions(), and get_action() for both planners: the tags of the
    arguments that reach dynamics_model.get_action, in order."""
    log = []
    tag = lambda s: (s,)                     # a tuple so that the recorder can tell the inputs apart
    for context in (False, True):
        for use_cem in (False, True):
            dm = RecordingDynamicsModel()
            c = controller_cls(name="policy", env=RewardEnv(), dynamics_model=dm, use_cem=use_cem, context=context)
            _, info = c.get_actions(tag("obs"), cp_obs=tag("cp_obs"), cp_act=tag("cp_act"), init_mean=tag("mean"),
                                    init_var=tag("var"))
            log.append(("get_actions", context, use_cem, dm.calls[-1], type(info).__name__))
    for use_cem in (False, True):
        dm = RecordingDynamicsModel()
        c = controller_cls(name="policy", env=RewardEnv(), dynamics_model=dm, use_cem=use_cem, context=False)
        obs = np.zeros(3)
        _, info = c.get_action(obs, init_mean=tag("mean"), init_var=tag("var"))
        args = dm.calls[-1][0]
        log.append(("get_action", False, use_cem, (("obs[None]",) + args[1:], dm.calls[-1][1]), type(info).__name__))
    return repr(log)
