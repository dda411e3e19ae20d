"""Generate the committed golden vectors of the planner hot path.

The reference (TensorFlow 1.15) cannot be executed in this environment, so these vectors are produced by the NumPy
oracle in float64 (oracle/cadm_oracle.py, pinned by the hand-derived known-answer tests in tests/test_oracle_kat.py).
They freeze the oracle (regression pin for `-m "not gpu"`) and give the GPU tests a fixture that does not depend on
re-running the oracle.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cadm_oracle as orc
from oracle import philox as ph
from oracle.envs import get_env

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, envname, E, p, n, h, H, m, context, deterministic, seed):
    env = get_env(envname)
    rng = np.random.default_rng(seed)
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    C, K = (4, 3) if context else (0, 0)
    prm = orc.init_dynamics_params(rng, E, P + A + C, H, D, dtype=np.float32)
    prm.b_lv[...] = -5.0
    for b in prm.b:
        b[...] = (0.05 * rng.standard_normal(b.shape)).astype(np.float32)
    enc = orc.init_encoder_params(rng, E, (D + A) * K, (32, 16, 8), C, dtype=np.float32) if context else None
    norm = orc.NormStats(
        (0.1 * rng.standard_normal(P)), rng.uniform(0.5, 1.5, P), (0.05 * rng.standard_normal(A)), rng.uniform(0.4, 0.8, A),
        (0.01 * rng.standard_normal(D)), rng.uniform(0.05, 0.2, D),
        np.zeros(D * K), np.ones(D * K), 0.05 * rng.standard_normal(A * K), rng.uniform(0.5, 1.0, A * K)).astype(np.float32)
    obs = (0.1 * rng.standard_normal((m, D))).astype(np.float32)
    cp_obs = (0.1 * rng.standard_normal((m, D * K))).astype(np.float32)
    cp_act = (0.1 * rng.standard_normal((m, A * K))).astype(np.float32)
    mean0 = (0.1 * rng.standard_normal((m, h, A))).astype(np.float32)
    var0 = np.full((m, h, A), 0.25, np.float32)
    z = ph.gen_z(seed, orc.NUM_CEM_ITERS, m, n, h, A)
    eps = None if deterministic else ph.gen_eps(seed, orc.NUM_CEM_ITERS, h, m, n, p, E, D)
    f8 = np.float64
    p64, n64 = prm.astype(f8), norm.astype(f8)
    ctx_raw = orc.encode_context(cp_obs.astype(f8), cp_act.astype(f8), enc.astype(f8), n64) if context else None
    res = orc.cem_plan(obs.astype(f8), mean0.astype(f8), var0.astype(f8), z.astype(f8), p64, n64, env, E, p, deterministic,
                       None if eps is None else eps.astype(f8), ctx_raw, trace=True)
    out = dict(
        meta=np.array([E, p, n, h, H, m, int(context), int(deterministic), seed, C, K]), envname=np.array(envname),
        obs=obs, cp_obs=cp_obs, cp_act=cp_act, mean0=mean0, var0=var0,
        **{f"W{i}": w for i, w in enumerate(prm.W)}, **{f"b{i}": b for i, b in enumerate(prm.b)},
        W_mu=prm.W_mu, b_mu=prm.b_mu, W_lv=prm.W_lv, b_lv=prm.b_lv, max_logvar=prm.max_logvar, min_logvar=prm.min_logvar,
        **({f"encW{i}": w for i, w in enumerate(enc.W)} if context else {}),
        **({f"encb{i}": b for i, b in enumerate(enc.b)} if context else {}),
        **{f"norm_{k}": getattr(norm, k) for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
                                                   "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std")},
        out_mean=res.mean, out_var=res.var, out_returns=res.returns, out_elites=res.elites,
        out_states_it0_last=res.states[0, -1].astype(np.float32), out_particle_returns_it0=res.particle_returns[0],
        out_ctx=np.zeros(0) if ctx_raw is None else ctx_raw)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written;", "returns[0,0,:3] =", res.returns[0, 0, :3], "elites[0,0,:5] =", res.elites[0, 0, :5])


if __name__ == "__main__":
    # noise (z, eps) is regenerated from the seed by oracle/philox.py, so the fixtures stay small
    make("hc_pets_small", "halfcheetah", E=5, p=10, n=64, h=6, H=64, m=2, context=False, deterministic=False, seed=11)
    make("hc_cadm_small", "halfcheetah", E=5, p=10, n=64, h=6, H=64, m=3, context=True, deterministic=False, seed=12)
    make("ant_vanilla_small", "ant", E=1, p=1, n=64, h=6, H=200, m=2, context=False, deterministic=True, seed=13)
