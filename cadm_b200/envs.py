"""Environment descriptors for the planner.

The reference bakes three closures of each gym environment into the planner graph -- `obs_preproc`,
`obs_postproc`, `tf_reward_fn` (e.g. cadm/envs/half_cheetah_env.py:46-56,82-88).  In this engine they live in
the CUDA epilogue (cadm_b200/csrc/common.cuh), selected by `env_id`.  The classes below are what the dynamics
model needs from an `env` argument: dimensions plus the env id.  A real reference env object (HalfCheetahEnv,
AntEnv, ... possibly wrapped in NormalizedEnv) is accepted too and mapped by class name, so the reference's run
scripts can pass their env unchanged.  MuJoCo simulation itself is out of scope.
"""
import numpy as np

from ._lib import ENV_IDS


class _Box:
    def __init__(self, dim, low=-1.0, high=1.0):
        self.shape = (dim,)
        self.low = np.full((dim,), low, dtype=np.float32)
        self.high = np.full((dim,), high, dtype=np.float32)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(np.float32)


class _Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()

    def sample(self):
        return np.random.randint(self.n)


class PlannerEnv:
    """Dimensions + analytic closures (NumPy, host side) of one environment family."""
    name = None
    obs_dim = proc_obs_dim = act_dim = 0
    discrete = False

    def __init__(self):
        self.observation_space = _Box(self.obs_dim, -np.inf, np.inf)
        self.action_space = _Discrete(self.act_dim) if self.discrete else _Box(self.act_dim)
        self.proc_observation_space_dims = self.proc_obs_dim
        self.env_id = ENV_IDS[self.name]

    # host-side NumPy versions (used by samplers / sample processors, never by the planner kernels)
    def obs_preproc(self, obs):
        return obs

    def obs_postproc(self, obs, pred):
        return obs + pred

    def targ_proc(self, obs, next_obs):
        return next_obs - obs

    def reward(self, obs, action, next_obs):
        raise NotImplementedError


class HalfCheetahSpec(PlannerEnv):
    name, obs_dim, proc_obs_dim, act_dim = "halfcheetah", 18, 18, 6

    def obs_preproc(self, obs):
        return np.concatenate([obs[..., 1:2], np.sin(obs[..., 2:3]), np.cos(obs[..., 2:3]), obs[..., 3:]], axis=-1)

    def obs_postproc(self, obs, pred):
        return np.concatenate([pred[..., :1], obs[..., 1:] + pred[..., 1:]], axis=-1)

    def targ_proc(self, obs, next_obs):
        return np.concatenate([next_obs[..., :1], next_obs[..., 1:] - obs[..., 1:]], axis=-1)

    def reward(self, obs, action, next_obs):
        return obs[..., 0] - 1e-1 * np.sum(np.square(action), axis=-1)


class CrippleHalfCheetahSpec(HalfCheetahSpec):
    name = "cripple_halfcheetah"


class AntSpec(PlannerEnv):
    name, obs_dim, proc_obs_dim, act_dim = "ant", 28, 27, 8

    def obs_preproc(self, obs):
        return obs[..., 1:]

    obs_postproc = HalfCheetahSpec.obs_postproc
    targ_proc = HalfCheetahSpec.targ_proc

    def reward(self, obs, act, next_obs):
        return obs[..., 0] + -0.005 * np.sum(np.square(act), axis=-1) + 0.0 + 0.05


class SlimHumanoidSpec(PlannerEnv):
    name, obs_dim, proc_obs_dim, act_dim = "slim_humanoid", 45, 45, 17

    def reward(self, obs, act, next_obs):
        alive = 5.0 * np.logical_and(obs[..., 1] > 1.0, obs[..., 1] < 2.0)
        return 0.25 / 0.015 * obs[..., 22] - 0.1 * np.sum(np.square(act), axis=-1) + alive


class CartPoleSpec(PlannerEnv):
    name, obs_dim, proc_obs_dim, act_dim, discrete = "cartpole", 4, 4, 2, True

    def reward(self, obs, act, next_obs):
        th = 12 * 2 * np.pi / 360
        cond = ((next_obs[..., 0] > 2.4) * 1.0 + (next_obs[..., 0] < -2.4) * 1.0 + (next_obs[..., 2] > th) * 1.0
                + (next_obs[..., 2] < -th) * 1.0)
        return 1 - cond


class PendulumSpec(PlannerEnv):
    name, obs_dim, proc_obs_dim, act_dim = "pendulum", 3, 3, 1
    max_torque = 2.0

    def reward(self, obs, action, next_obs):
        theta = np.arctan2(obs[..., 1], obs[..., 0])
        tn = ((theta + np.pi) % (2 * np.pi)) - np.pi
        tq = np.clip(action, -self.max_torque, self.max_torque)[..., 0]
        return -(tn ** 2 + 0.1 * obs[..., 2] ** 2 + 0.001 * tq ** 2)


SPECS = {c.name: c for c in (HalfCheetahSpec, CrippleHalfCheetahSpec, AntSpec, SlimHumanoidSpec, CartPoleSpec, PendulumSpec)}

# reference env class name -> spec (cadm/envs/*.py)
_REFERENCE_CLASSES = {
    "HalfCheetahEnv": "halfcheetah", "CrippleHalfCheetahEnv": "cripple_halfcheetah", "AntEnv": "ant",
    "SlimHumanoidEnv": "slim_humanoid", "RandomCartPole_Force_Length": "cartpole", "ModifiableCartPoleEnv": "cartpole",
    "RandomPendulumAll": "pendulum", "ModifiablePendulumEnv": "pendulum",
}


def make_env(name):
    return SPECS[name]()


def resolve_env(env):
    """Return (env_object, env_name).  Accepts a name, a PlannerEnv, or a reference env (unwrapping NormalizedEnv)."""
    if isinstance(env, str):
        e = make_env(env)
        return e, e.name
    if isinstance(env, PlannerEnv):
        return env, env.name
    inner = env
    while hasattr(inner, "wrapped_env"):
        inner = inner.wrapped_env
    for klass in type(inner).__mro__:
        if klass.__name__ in _REFERENCE_CLASSES:
            return env, _REFERENCE_CLASSES[klass.__name__]
    raise ValueError(f"cadm_b200 has no planner epilogue for environment {type(inner).__name__}")
