"""TEST INFRASTRUCTURE (oracle) -- analytic per-environment closures of the CaDM planner.

NumPy restatement of the `obs_preproc` / `obs_postproc` / `tf_reward_fn` closures that the
reference bakes into its planner graph.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package.

Reference (paths relative to /root/reference):
  HalfCheetah        cadm/envs/half_cheetah_env.py:46-56 (pre/postproc), :82-88 (tf_reward_fn)
  CrippleHalfCheetah cadm/envs/half_cheetah_cripple_env.py:60-73, :90-96   (identical closures)
  Ant                cadm/envs/ant_env.py:52-59, :89-98
  SlimHumanoid       cadm/envs/slim_humanoid_env.py:39-46, :95-111
  CartPole           cadm/envs/classic_control.py:94-101, :154-166        (discrete actions)
  Pendulum           cadm/envs/classic_control.py:209-218 (reward), :284-291

Every closure works on arrays of shape [..., dim] and keeps the array dtype (float32 or float64),
so the same code is both the fp64 "truth" and the fp32 restatement of the TF graph.
"""
import math
from dataclasses import dataclass
from typing import Callable

import numpy as np


@dataclass(frozen=True)
class EnvSpec:
    name: str
    env_id: int          # must match CADM_ENV_* in include/cadm_b200.h
    obs_dim: int         # D
    proc_obs_dim: int    # P
    act_dim: int         # A
    discrete: bool
    preproc: Callable
    postproc: Callable
    reward: Callable     # reward(obs, act, next_obs) -> [...]


def _hc_preproc(obs):
    # half_cheetah_env.py:46-50
    return np.concatenate([obs[..., 1:2], np.sin(obs[..., 2:3]), np.cos(obs[..., 2:3]), obs[..., 3:]], axis=-1)


def _vel_postproc(obs, pred):
    # half_cheetah_env.py:52-56 ; ant_env.py:55-59
    return np.concatenate([pred[..., :1], obs[..., 1:] + pred[..., 1:]], axis=-1)


def _hc_reward(obs, act, next_obs):
    # half_cheetah_env.py:82-88 -- reads the CURRENT obs (quirk Q4)
    ctrl_cost = obs.dtype.type(1e-1) * np.sum(np.square(act), axis=-1)
    return obs[..., 0] - ctrl_cost


def _ant_preproc(obs):
    return obs[..., 1:]          # ant_env.py:52-53


def _ant_reward(obs, act, next_obs):
    # ant_env.py:89-98: reward_run + reward_ctrl + 0.0 + 0.05, evaluated left to right
    t = obs.dtype.type
    reward_ctrl = t(-0.005) * np.sum(np.square(act), axis=-1)
    reward_run = obs[..., 0]
    return reward_run + reward_ctrl + t(0.0) + t(0.05)


def _id_preproc(obs):
    return obs


def _add_postproc(obs, pred):
    return obs + pred


def _humanoid_reward(obs, act, next_obs):
    # slim_humanoid_env.py:95-111
    t = obs.dtype.type
    lin_vel_cost = t(0.25 / 0.015) * obs[..., 22]
    quad_ctrl_cost = t(0.1) * np.sum(np.square(act), axis=-1)
    alive = t(5.0) * np.logical_and(obs[..., 1] > 1.0, obs[..., 1] < 2.0).astype(obs.dtype)
    return lin_vel_cost - quad_ctrl_cost - t(0.0) + alive


def _cartpole_reward(obs, act, next_obs):
    # classic_control.py:154-166 -- reads NEXT obs
    t = obs.dtype.type
    x_thr = 2.4
    th_thr = 12 * 2 * math.pi / 360
    cond = ((next_obs[..., 0] > x_thr).astype(obs.dtype) + (next_obs[..., 0] < -x_thr).astype(obs.dtype)
            + (next_obs[..., 2] > th_thr).astype(obs.dtype) + (next_obs[..., 2] < -th_thr).astype(obs.dtype))
    return t(1) - cond * t(1)


def _pendulum_reward(obs, act, next_obs, max_torque=2.0):
    # classic_control.py:209-218 ; python/TF `%` is the floored modulo
    t = obs.dtype.type
    theta = np.arctan2(obs[..., 1], obs[..., 0])
    theta_n = np.mod(theta + t(np.pi), t(2 * np.pi)) - t(np.pi)
    thetadot = obs[..., 2]
    torque = np.clip(act, -max_torque, max_torque)[..., 0]
    cost = theta_n ** 2 + t(0.1) * thetadot ** 2 + t(0.001) * torque ** 2
    return -cost


ENVS = {
    "halfcheetah": EnvSpec("halfcheetah", 0, 18, 18, 6, False, _hc_preproc, _vel_postproc, _hc_reward),
    "ant": EnvSpec("ant", 1, 28, 27, 8, False, _ant_preproc, _vel_postproc, _ant_reward),
    "slim_humanoid": EnvSpec("slim_humanoid", 2, 45, 45, 17, False, _id_preproc, _add_postproc, _humanoid_reward),
    "cartpole": EnvSpec("cartpole", 3, 4, 4, 2, True, _id_preproc, _add_postproc, _cartpole_reward),
    "pendulum": EnvSpec("pendulum", 4, 3, 3, 1, False, _id_preproc, _add_postproc, _pendulum_reward),
}
ENVS["cripple_halfcheetah"] = EnvSpec("cripple_halfcheetah", 0, 18, 18, 6, False,
                                      _hc_preproc, _vel_postproc, _hc_reward)


def get_env(name: str) -> EnvSpec:
    return ENVS[name]
