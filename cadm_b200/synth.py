"""Synthetic workloads of SURVEY.md section 8(d) / BASELINE.md: random-init weights of the reference architecture,
fixed normalisation statistics, synthetic start states.  No MuJoCo, no datasets.  NumPy only (host side)."""
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np

from .envs import make_env

# name -> (env, E, p, n, h, deterministic, context)
CONFIGS = {
    "C1": dict(env="halfcheetah", ensemble=1, particles=1, candidates=200, horizon=30, deterministic=True, context=False),
    "C2": dict(env="halfcheetah", ensemble=5, particles=20, candidates=200, horizon=30, deterministic=False, context=False),
    "C3": dict(env="halfcheetah", ensemble=5, particles=20, candidates=200, horizon=30, deterministic=False, context=True),
    "C4": dict(env="ant", ensemble=5, particles=20, candidates=1000, horizon=30, deterministic=False, context=True),
}
WORKLOAD_NAMES = {
    "C1": "HalfCheetah Vanilla-DM CEM (ens=1, part=1, cand=200, horizon=30, deterministic)",
    "C2": "HalfCheetah PE-TS CEM (ens=5, part=20, cand=200, horizon=30)",
    "C3": "HalfCheetah PE-TS + CaDM context encoder (history=10, ens=5, part=20, cand=200, horizon=30)",
    "C4": "Ant PE-TS + CaDM (ens=5, part=20, cand=1000, horizon=30)",
}
HISTORY = 10
CTX_DIM = 10


def flops_per_unit(In, H, D, deterministic, n_hidden=4):
    """Algorithmic FLOPs of one (candidate, particle, step) dynamics evaluation."""
    k = 1 if deterministic else 2
    return 2 * (In * H + (n_hidden - 1) * H * H + k * H * D)


def synthetic_normalization(env, context, trained_like=True):
    D, P, A = env.obs_dim, env.proc_obs_dim, env.act_dim
    norm = OrderedDict()
    norm["obs"] = (np.zeros(P), np.ones(P))
    norm["delta"] = (np.zeros(D), np.full(D, 0.1))
    norm["act"] = (np.zeros(A), np.full(A, 0.6))
    if context:
        norm["cp_obs"] = (np.zeros(D * HISTORY), np.ones(D * HISTORY))
        norm["cp_act"] = (np.zeros(A * HISTORY), np.ones(A * HISTORY))
        norm["back_delta"] = (np.zeros(D), np.ones(D))
    return norm


def synthetic_inputs(env, m, horizon, context, seed=0):
    """Start states obs ~ 0.1 N(0,1) (HalfCheetah angle o[2] ~ U(-0.3, 0.3)), history ~ 0.1 N(0,1),
    init_mean = 0, init_var = 0.25 (cadm/samplers/sampler.py:52-53)."""
    rng = np.random.default_rng(seed)
    D, A = env.obs_dim, env.act_dim
    obs = 0.1 * rng.standard_normal((m, D))
    if env.name in ("halfcheetah", "cripple_halfcheetah"):
        obs[:, 2] = rng.uniform(-0.3, 0.3, size=m)
    out = dict(obs=obs.astype(np.float32),
               init_mean=np.zeros((m, horizon, A), np.float32),
               init_var=np.full((m, horizon, A), 0.25, np.float32))
    if context:
        out["cp_obs"] = (0.1 * rng.standard_normal((m, D * HISTORY))).astype(np.float32)
        out["cp_act"] = (0.1 * rng.standard_normal((m, A * HISTORY))).astype(np.float32)
    return out


def build_model(config="C2", m_max=1, seed=0, trained_like=True, candidates=None, particles=None, **engine_kwargs):
    """Dynamics model for a named config with random-init weights; `trained_like` sets b_logvar = -6 so the sampled
    noise is small (random-init logvar ~ 0 makes the rollout noise-dominated)."""
    cfg = dict(CONFIGS[config])
    if candidates is not None:
        cfg["candidates"] = candidates
    if particles is not None:
        cfg["particles"] = particles
    env = make_env(cfg["env"])
    common = dict(hidden_sizes=(200, 200, 200, 200), hidden_nonlinearity="swish", n_forwards=cfg["horizon"],
                  n_candidates=cfg["candidates"], ensemble_size=cfg["ensemble"], n_particles=cfg["particles"], use_cem=True,
                  deterministic=cfg["deterministic"], normalize_input=True, seed=seed, m_max=m_max, **engine_kwargs)
    if cfg["context"]:
        from .dynamics.mlp_cadm_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as M
        model = M("dm", env, context_out_dim=CTX_DIM, history_length=HISTORY, state_diff=True, **common)
    else:
        from .dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as M
        model = M("dm", env, **common)
    if trained_like and not cfg["deterministic"]:
        model._dyn["b_lv"][...] = -6.0
        model._push_params()
    model.set_normalization(synthetic_normalization(env, cfg["context"]))
    return model, env, cfg
