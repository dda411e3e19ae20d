"""The precision argument of the tensor-core path, checked on the CPU: the oracle run in float32 with every contraction
replaced by the kernels' split arithmetic (oracle/tc_model.py: x8 / x64 pre-scales, fp16 hi/lo halves, three products, fp32
accumulation) against the float64 oracle, on the committed golden fixtures and on a full-width (H = 200, E = 5, h = 30)
HalfCheetah PE-TS problem."""
import glob
import os

import numpy as np
import pytest

from oracle import cadm_oracle as orc
from oracle import philox as ph
from oracle import tc_model
from oracle.envs import get_env

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
TOL = 1e-4                       # BASELINE.json's parity bar


def _fixture_problem(g, dt):
    E, p, n, h, H, m, context, det, seed, C, K = [int(v) for v in g["meta"]]
    env = get_env(str(g["envname"]))
    prm = orc.DynamicsParams([g[f"W{i}"] for i in range(4)], [g[f"b{i}"] for i in range(4)], g["W_mu"], g["b_mu"], g["W_lv"],
                             g["b_lv"], g["max_logvar"], g["min_logvar"]).astype(dt)
    norm = orc.NormStats(*[g[f"norm_{k}"] for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
                                                     "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std")]).astype(dt)
    z = ph.gen_z(seed, orc.NUM_CEM_ITERS, m, n, h, env.act_dim).astype(dt)
    eps = None if det else ph.gen_eps(seed, orc.NUM_CEM_ITERS, h, m, n, p, E, env.obs_dim).astype(dt)
    return env, prm, norm, z, eps, E, p, bool(det)


def _plan(g, dt, ctx_raw=None):
    env, prm, norm, z, eps, E, p, det = _fixture_problem(g, dt)
    return orc.cem_plan(g["obs"].astype(dt), g["mean0"].astype(dt), g["var0"].astype(dt), z, prm, norm, env, E, p, det, eps,
                        None if ctx_raw is None else ctx_raw.astype(dt))


def test_split_product_is_fp32_class_and_single_pass_is_not():
    """One layer of the reference width: error of the three-term split against float64, relative to the output RMS."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 800, 200)).astype(np.float32) * 0.5
    W = (rng.standard_normal((5, 200, 200)) / (2 * np.sqrt(200))).astype(np.float32)
    want = np.matmul(x.astype(np.float64), W.astype(np.float64))
    rms = np.sqrt(np.mean(want ** 2))
    e3 = np.max(np.abs(tc_model.matmul_split(x, W, 3) - want)) / rms
    e1 = np.max(np.abs(tc_model.matmul_split(x, W, 1) - want)) / rms
    e32 = np.max(np.abs(np.matmul(x, W) - want)) / rms
    assert e3 < 5e-6 and e3 < 8 * max(e32, 1e-7)          # as good as a plain float32 product, give or take summation order
    assert e1 > 50 * e3                                   # one fp16 pass (tc1x) is a different class
    hi, lo = tc_model.split_f16(x * 8)
    assert np.max(np.abs(hi + lo - x * 8) / np.maximum(np.abs(x * 8), 2.0 ** -3)) < 2.0 ** -21


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f) for f in FIXTURES])
def test_whole_decision_in_split_arithmetic_stays_inside_the_parity_bar(path, monkeypatch):
    """5 CEM iterations with the split contraction everywhere: returns far inside 1e-4 of the float64 fixture, the same elites,
    the same plan -- the fixtures the GPU tests compare the engine with."""
    g = np.load(path)
    context = int(g["meta"][6])
    ctx = g["out_ctx"] if context else None
    monkeypatch.setattr(orc, "dense", tc_model.dense_split(3))
    res = _plan(g, np.float32, ctx)
    scale = np.max(np.abs(g["out_returns"]))
    assert np.max(np.abs(res.returns - g["out_returns"])) / scale < TOL / 20
    assert np.array_equal(res.elites, g["out_elites"])
    assert np.max(np.abs(res.mean - g["out_mean"])) < TOL / 20


def test_full_width_rollout_split_against_float64(monkeypatch):
    """Reference architecture (4 x 200, E = 5, p = 20, h = 30): every state of one 30-step rollout in split arithmetic vs
    float64 (relative to the RMS of each state dimension, the measure of the GPU parity tests), with tc1x alongside for scale."""
    env = get_env("halfcheetah")
    rng = np.random.default_rng(5)
    E, p, n, h, m = 5, 20, 8, 30, 1
    prm = orc.init_dynamics_params(rng, E, env.proc_obs_dim + env.act_dim, 200, env.obs_dim, dtype=np.float32)
    prm.b_lv[...] = -6.0
    norm = orc.NormStats(np.zeros(18), np.ones(18), np.zeros(6), np.full(6, 0.6), np.zeros(18), np.full(18, 0.1)).astype(np.float32)
    obs = (0.1 * rng.standard_normal((m, 18))).astype(np.float32)
    acts = rng.uniform(-1, 1, (m, n, h, 6)).astype(np.float32)
    eps = ph.gen_eps(3, 1, h, m, n, p, E, 18)[0]
    f8 = np.float64
    _, want = orc.rollout(obs.astype(f8), acts.astype(f8), prm.astype(f8), norm.astype(f8), env, E, p, False, eps.astype(f8),
                          trace=True)
    rms = np.sqrt(np.mean(want ** 2, axis=(0, 1, 2, 3)))               # per state dimension, states [h, m, n, p, D]
    errs = {}
    for terms in (3, 1):
        with monkeypatch.context() as mp:
            mp.setattr(orc, "dense", tc_model.dense_split(terms))
            _, got = orc.rollout(obs, acts, prm, norm, env, E, p, False, eps, trace=True)
        errs[terms] = float(np.max(np.abs(got - want) / rms))
    assert errs[3] < TOL / 10, errs
    assert errs[1] > 20 * errs[3], errs


def test_cheaper_arithmetic_is_not_robust_to_the_weight_scale(monkeypatch):
    """Two tempting shortcuts for the hidden-layer epilogue, screened with the model before spending GPU time on them:
    a one-MUFU swish (0.5x tanh.approx(0.5x) + 0.5x, tanh.approx.f32 is specified to 2^-11 relative) and fp16-only activations
    (drop X_lo: no lo split, half the operand stores).  On random-init weights both stay inside the 1e-4 bar; with the hidden
    weights scaled x2 (a less contractive, trained-like network) both leave it by a wide margin, while the three-term split with
    an fp32-accurate swish does not move.  Hence the kernels keep ex2 + rcp and the full split."""
    env = get_env("halfcheetah")
    f8 = np.float64

    def dense_xhi(x, W, b, act=None):
        xh, _ = tc_model.split_f16(np.asarray(x, np.float32) * tc_model.X_SCALE)
        wh, wl = tc_model.split_f16(np.asarray(W, np.float32) * tc_model.W_SCALE)
        out = (np.matmul(xh, wh) + np.matmul(xh, wl)) * np.float32(1 / 512) + np.asarray(b, np.float32)
        return act(out) if act is not None else out

    def swish_tanh(x):
        hx = np.float32(0.5) * x
        return hx * (np.tanh(hx) * (1 + np.float32(2.0 ** -11))) + hx

    plain_swish = orc.swish
    variants = dict(tc3x=(tc_model.dense_split(3), plain_swish), tanh=(tc_model.dense_split(3), swish_tanh),
                    xhi=(dense_xhi, plain_swish))
    errs = {}
    for wscale in (1.0, 2.0):
        rng = np.random.default_rng(7)
        E, p, n, h, m = 5, 20, 4, 30, 1
        prm = orc.init_dynamics_params(rng, E, env.proc_obs_dim + env.act_dim, 200, env.obs_dim, dtype=np.float32)
        prm.b_lv[...] = -6.0
        for W in prm.W:
            W *= np.float32(wscale)
        norm = orc.NormStats(np.zeros(18), np.ones(18), np.zeros(6), np.full(6, 0.6), np.zeros(18), np.full(18, 0.1)).astype(np.float32)
        obs = (0.1 * rng.standard_normal((m, 18))).astype(np.float32)
        acts = rng.uniform(-1, 1, (m, n, h, 6)).astype(np.float32)
        eps = ph.gen_eps(3, 1, h, m, n, p, E, 18)[0]
        _, want = orc.rollout(obs.astype(f8), acts.astype(f8), prm.astype(f8), norm.astype(f8), env, E, p, False, eps.astype(f8),
                              trace=True)
        rms = np.sqrt(np.mean(want ** 2, axis=(0, 1, 2, 3)))
        for name, (dense, act) in variants.items():
            with monkeypatch.context() as mp:
                mp.setattr(orc, "dense", dense)
                mp.setattr(orc, "swish", act)
                _, got = orc.rollout(obs, acts, prm, norm, env, E, p, False, eps, trace=True)
            errs[name, wscale] = float(np.max(np.abs(got - want) / rms))
    assert errs["tc3x", 1.0] < TOL / 20 and errs["tc3x", 2.0] < TOL / 20, errs
    assert errs["tanh", 1.0] < TOL and errs["xhi", 1.0] < TOL, errs
    assert errs["tanh", 2.0] > TOL and errs["xhi", 2.0] > TOL, errs
