"""fit() of the PE-TS ensemble (cadm_b200/dynamics/training.py) against the NumPy restatement of the reference's training
losses (oracle/train_oracle.py), finite differences, closed forms, and the reference's fit() loop semantics.  CPU only."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from cadm_b200.dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel
from cadm_b200.dynamics.mlp_cadm_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as CaDMModel
from cadm_b200.dynamics.training import CaDMTrainer, EnsembleNLLTrainer, fit_cadm_ensemble, fit_ensemble, flatten_future
from cadm_b200.envs import make_env
from oracle import cadm_oracle as orc
from oracle.train_oracle import cadm_context, cadm_losses, flatten_future_loops, pets_losses


def _dyn(rng, E, In, H, D, n_hidden=2):
    prm = orc.init_dynamics_params(rng, E, In, H, D, n_hidden=n_hidden)
    return dict(W=[np.array(w, np.float64) for w in prm.W], b=[np.array(b, np.float64) for b in prm.b],
                W_mu=np.array(prm.W_mu, np.float64), b_mu=np.array(prm.b_mu, np.float64),
                W_lv=np.array(prm.W_lv, np.float64), b_lv=np.array(prm.b_lv, np.float64),
                max_logvar=np.array(prm.max_logvar, np.float64).reshape(1, -1), min_logvar=np.array(prm.min_logvar, np.float64).reshape(1, -1))


def _batch(rng, env, E, B):
    D, A = env.obs_dim, env.act_dim
    obs = rng.standard_normal((E, B, D)) * 0.5
    act = rng.uniform(-1, 1, (E, B, A))
    delta = rng.standard_normal((E, B, D)) * 0.1
    stats = (rng.standard_normal(env.proc_obs_dim) * 0.1, rng.uniform(0.5, 1.5, env.proc_obs_dim), np.zeros(A), np.full(A, 0.6),
             rng.standard_normal(D) * 0.01, rng.uniform(0.05, 0.2, D))
    return obs, act, delta, stats


@pytest.mark.parametrize("envname", ["halfcheetah", "ant", "pendulum"])
@pytest.mark.parametrize("deterministic", [False, True])
def test_losses_match_the_numpy_restatement(envname, deterministic):
    env = make_env(envname)
    rng = np.random.default_rng(0)
    E, B = 3, 17
    dyn = _dyn(rng, E, env.proc_obs_dim + env.act_dim, 24, env.obs_dim)
    dyn["b_lv"] += rng.standard_normal(dyn["b_lv"].shape)          # exercise both soft bounds
    dyn["b_lv"][0] += 3.0
    dyn["b_lv"][1] -= 14.0
    obs, act, delta, stats = _batch(rng, env, E, B)
    wd, coeff = (1e-4, 2e-4, 3e-4), 0.5
    tr = EnsembleNLLTrainer(dyn, envname, deterministic, wd, coeff, 1e-3, dtype=torch.float64)
    got = {k: float(v.detach()) for k, v in tr.losses(obs, act, delta, stats).items()}
    want = pets_losses(dyn, envname, deterministic, wd, coeff, obs, act, delta, stats)
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-10 * max(1.0, abs(want[k])), k


def test_gradients_match_finite_differences_of_the_restatement():
    env = make_env("halfcheetah")
    rng = np.random.default_rng(1)
    E, B = 2, 9
    dyn = _dyn(rng, E, env.proc_obs_dim + env.act_dim, 16, env.obs_dim)
    obs, act, delta, stats = _batch(rng, env, E, B)
    wd, coeff = (1e-3, 1e-3, 1e-3), 1.0
    tr = EnsembleNLLTrainer(dyn, "halfcheetah", False, wd, coeff, 1e-3, dtype=torch.float64)
    tr.losses(obs, act, delta, stats)["loss"].backward()
    probes = [("W", 0, (1, 3, 5)), ("W", 1, (0, 7, 2)), ("b", 0, (1, 0, 4)), ("W_mu", None, (0, 5, 3)), ("W_lv", None, (1, 2, 9)),
              ("b_lv", None, (0, 0, 1)), ("max_logvar", None, (0, 4)), ("min_logvar", None, (0, 7))]
    for name, li, idx in probes:
        arr = dyn[name] if li is None else dyn[name][li]
        par = getattr(tr, name) if li is None else getattr(tr, name)[li]
        h = 1e-6
        old = arr[idx]
        arr[idx] = old + h
        up = pets_losses(dyn, "halfcheetah", False, wd, coeff, obs, act, delta, stats)["loss"]
        arr[idx] = old - h
        dn = pets_losses(dyn, "halfcheetah", False, wd, coeff, obs, act, delta, stats)["loss"]
        arr[idx] = old
        fd = (up - dn) / (2 * h)
        assert abs(float(par.grad[idx]) - fd) <= 1e-6 * max(1.0, abs(fd)), (name, li, idx, float(par.grad[idx]), fd)


def test_zero_weights_closed_form():
    """W = 0, b = 0: mu = 0 and logvar = softbound(0) for every sample, so every loss term has a closed form."""
    env = make_env("pendulum")
    rng = np.random.default_rng(2)
    E, B, D = 4, 11, env.obs_dim
    dyn = _dyn(rng, E, env.proc_obs_dim + env.act_dim, 8, D)
    for k in ("W_mu", "W_lv"):
        dyn[k][...] = 0
    for w in dyn["W"]:
        w[...] = 0
    obs, act, delta, stats = _batch(rng, env, E, B)
    tr = EnsembleNLLTrainer(dyn, "pendulum", False, (0., 0., 0.), 0.0, 1e-3, dtype=torch.float64)
    out = {k: float(v.detach()) for k, v in tr.losses(obs, act, delta, stats).items()}
    target = (delta - stats[4]) / (stats[5] + 1e-10)
    sp = lambda x: np.log1p(np.exp(x))
    lv = 0.5 - sp(0.5 - 0.0)
    lv = -10.0 + sp(lv + 10.0)
    assert abs(out["mse_loss"] - np.sum(np.mean(target ** 2, axis=(1, 2)))) < 1e-10
    assert abs(out["var_loss"] - E * lv) < 1e-10
    assert abs(out["mu_loss"] - np.exp(-lv) * np.sum(np.mean(target ** 2, axis=(1, 2)))) < 1e-9
    assert abs(out["reg_loss"] - (0.01 * 0.5 * D + 0.01 * 10.0 * D)) < 1e-12


class _CpuModel(MLPEnsembleCEMDynamicsModel):
    """The host mirror without an engine (no GPU in this suite): same attributes, same normalisation code."""

    def __init__(self, envname, E=3, H=32, n_hidden=2, deterministic=False, seed=0, **kw):
        self.env = make_env(envname)
        self.env_name = envname
        self.deterministic, self.ensemble_size = deterministic, E
        self.obs_space_dims, self.action_space_dims = self.env.obs_dim, self.env.act_dim
        self.proc_obs_space_dims = self.env.proc_obs_dim
        self.discrete, self.normalize_input, self.normalization, self._dataset = False, True, None, None
        self.batch_size, self.learning_rate = kw.get("batch_size", 64), kw.get("learning_rate", 3e-3)
        self.valid_split_ratio, self.rolling_average_persitency = 0.2, kw.get("rolling_average_persitency", 0.99)
        self.weight_decays, self.weight_decay_coeff = (0.,) * (n_hidden + 1), 0.0
        rng = np.random.default_rng(seed)
        d = _dyn(rng, E, self.proc_obs_space_dims + self.action_space_dims, H, self.obs_space_dims, n_hidden=n_hidden)
        self._dyn = {k: ([a.astype(np.float32) for a in v] if isinstance(v, list) else v.astype(np.float32)) for k, v in d.items()}
        self.engine = None
        self.pushed = 0

    def _push_params(self):
        self.pushed += 1

    def _push_norm(self):
        pass


def _transitions(rng, env, n):
    D, A = env.obs_dim, env.act_dim
    obs = rng.standard_normal((n, D))
    act = rng.uniform(-1, 1, (n, A))
    M = rng.standard_normal((D + A, D)) * 0.05
    nxt = obs + np.concatenate([obs, act], axis=1) @ M + 0.001 * rng.standard_normal((n, D))
    return obs, act, nxt


def test_fit_loop_learns_and_keeps_the_reference_bookkeeping():
    rng = np.random.default_rng(3)
    model = _CpuModel("pendulum", E=3, H=32)
    env = model.env
    obs, act, nxt = _transitions(rng, env, 600)
    before = EnsembleNLLTrainer(model._dyn, "pendulum", False, model.weight_decays, 0.0, 1e-3)
    logs = []
    info = fit_ensemble(model, obs, act, nxt, epochs=25, rng=np.random.default_rng(4), verbose=True, log=logs.append, device="cpu")
    assert info["epochs"] >= 1 and model.pushed == 1
    # normalisation statistics as the reference computes them (:344-352), dataset kept for the next call (:225-231)
    assert isinstance(model.normalization, OrderedDict) and list(model.normalization) == ["obs", "delta", "act"]
    np.testing.assert_allclose(model.normalization["delta"][0], np.mean(env.targ_proc(obs, nxt), axis=0))
    assert model._dataset["obs"].shape[0] == 600
    stats = model.get_normalization_stats()[:6]
    E = model.ensemble_size
    tile = lambda a: np.tile(a[None], (E, 1, 1))
    delta = env.targ_proc(obs, nxt)
    mse0, _ = before.evaluate(tile(obs), tile(act), tile(delta), stats)
    after = EnsembleNLLTrainer(model._dyn, "pendulum", False, model.weight_decays, 0.0, 1e-3)
    mse1, _ = after.evaluate(tile(obs), tile(act), tile(delta), stats)
    assert mse1 < 0.25 * mse0, (mse0, mse1)
    assert any("finished epoch 0" in l for l in logs)
    # a second call appends to the dataset and recomputes the statistics over all of it
    obs2, act2, nxt2 = _transitions(rng, env, 200)
    fit_ensemble(model, obs2, act2, nxt2, epochs=1, rng=np.random.default_rng(5), device="cpu", log=logs.append)
    assert model._dataset["obs"].shape[0] == 800 and model.pushed == 2


@pytest.mark.parametrize("script,stop_after", [([1.0, 0.9, 5.0, 0.1], 3), ([-4.0, -1.0, -1.0, -9.0], 3), ([1.0, 1.0, 1.0, 1.0], 4)])
def test_early_stopping_rule(script, stop_after, monkeypatch):
    """mlp_ensemble_cem_dynamics.py:300-311 with scripted validation losses and persistency 0.5: the first loss v0 sets
    rolling = 1.5 v0 and the (never updated) bound 2 v0 -- v0 / 1.5 and v0 / 2 when v0 < 0 -- and training stops in the first
    epoch whose rolling average 0.5 rolling + 0.5 v exceeds the bound.  [1, .9, 5]: 1.25, 1.075, 3.04 > 2 -> 3 epochs;
    [-4, -1, -1]: -3.33, -2.17, -1.58 > -2 -> 3 epochs; constant losses never stop."""
    rng = np.random.default_rng(6)
    model = _CpuModel("pendulum", E=2, H=8, rolling_average_persitency=0.5)
    obs, act, nxt = _transitions(rng, model.env, 120)
    it = iter(script)
    monkeypatch.setattr(EnsembleNLLTrainer, "evaluate", lambda self, *a: (0.0, next(it)))
    logs = []
    info = fit_ensemble(model, obs, act, nxt, epochs=len(script), rng=np.random.default_rng(7), device="cpu", log=logs.append)
    assert info["epochs"] == stop_after
    assert any("Stopping Training" in l for l in logs) == (stop_after < len(script))


# ---------------------------------------------------------------------------------------------------------- CaDM

def _enc(rng, E, In, hidden, C):
    sizes = [In] + list(hidden) + [C]
    return dict(W=[rng.standard_normal((E, sizes[i], sizes[i + 1])) / (2 * np.sqrt(sizes[i])) for i in range(len(sizes) - 1)],
                b=[rng.standard_normal((E, 1, sizes[i + 1])) * 0.1 for i in range(len(sizes) - 1)])


def _cadm_batch(rng, env, E, B, K):
    D, A = env.obs_dim, env.act_dim
    obs, act, delta, stats6 = _batch(rng, env, E, B)
    nxt = obs + delta
    back_delta = -delta + 0.01 * rng.standard_normal(delta.shape)
    cp_obs = rng.standard_normal((E, B, D * K))
    cp_act = rng.uniform(-1, 1, (E, B, A * K))
    stats = stats6 + (rng.standard_normal(D * K) * 0.1, rng.uniform(0.5, 1.5, D * K), np.zeros(A * K), np.full(A * K, 0.6),
                      rng.standard_normal(D) * 0.01, rng.uniform(0.05, 0.2, D))
    return (obs, act, delta, nxt, back_delta, cp_obs, cp_act), stats


@pytest.mark.parametrize("envname", ["halfcheetah", "pendulum"])
@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("back_coeff", [0.0, 0.5])
def test_cadm_losses_match_the_numpy_restatement(envname, deterministic, back_coeff):
    env = make_env(envname)
    rng = np.random.default_rng(10)
    E, B, K, C = 3, 13, 4, 5
    enc = _enc(rng, E, (env.obs_dim + env.act_dim) * K, (12, 8), C)
    dyn = _dyn(rng, E, env.proc_obs_dim + env.act_dim + C, 24, env.obs_dim)
    back = _dyn(rng, E, env.proc_obs_dim + env.act_dim + C, 24, env.obs_dim)
    dyn["b_lv"][0] += 3.0
    dyn["b_lv"][1] -= 14.0
    batch, stats = _cadm_batch(rng, env, E, B, K)
    wd, cwd, coeff = (1e-4, 2e-4, 3e-4), (5e-4, 6e-4, 7e-4), 0.5
    tr = CaDMTrainer(enc, dyn, back, envname, deterministic, wd, cwd, coeff, back_coeff, 1e-3, dtype=torch.float64)
    got = {k: float(v.detach()) if torch.is_tensor(v) else float(v) for k, v in tr.losses(*batch, stats).items()}
    want = cadm_losses(enc, dyn, back, envname, deterministic, wd, cwd, coeff, back_coeff, *batch, stats)
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-10 * max(1.0, abs(want[k])), k
    assert (got["back_mse_loss"] > 0) == (back_coeff > 0)
    ctx = tr.context(*[torch.as_tensor(batch[i]) for i in (5, 6)], [torch.as_tensor(np.asarray(s_)) for s_ in stats])
    np.testing.assert_allclose(ctx.detach().numpy(), cadm_context(enc, batch[5], batch[6], stats), rtol=1e-12, atol=1e-12)


def test_cadm_gradients_match_finite_differences_and_reach_the_encoder_through_both_models():
    env = make_env("halfcheetah")
    rng = np.random.default_rng(11)
    E, B, K, C = 2, 7, 3, 4
    enc = _enc(rng, E, (env.obs_dim + env.act_dim) * K, (10,), C)
    dyn = _dyn(rng, E, env.proc_obs_dim + env.act_dim + C, 16, env.obs_dim)
    back = _dyn(rng, E, env.proc_obs_dim + env.act_dim + C, 16, env.obs_dim)
    batch, stats = _cadm_batch(rng, env, E, B, K)
    wd, cwd, coeff, bc = (1e-3,) * 3, (2e-3,) * 2, 1.0, 0.7
    args = ("halfcheetah", False, wd, cwd, coeff, bc)
    tr = CaDMTrainer(enc, dyn, back, *args, 1e-3, dtype=torch.float64)
    tr.losses(*batch, stats)["loss"].backward()
    probes = [(enc["W"][0], tr.enc_W[0], (1, 5, 3)), (enc["W"][1], tr.enc_W[1], (0, 2, 1)), (enc["b"][0], tr.enc_b[0], (1, 0, 6)),
              (dyn["W"][0], tr.fwd.W[0], (0, 20, 3)),                # a context column of the forward model's first layer
              (dyn["W_lv"], tr.fwd.W_lv, (1, 2, 9)), (back["W"][1], tr.back.W[1], (0, 7, 2)),
              (back["W_mu"], tr.back.W_mu, (1, 4, 4)), (back["W_lv"], tr.back.W_lv, (0, 3, 3))]
    for arr, par, idx in probes:
        h, old = 1e-6, arr[idx]
        arr[idx] = old + h
        up = cadm_losses(enc, dyn, back, *args, *batch, stats)["loss"]
        arr[idx] = old - h
        dn = cadm_losses(enc, dyn, back, *args, *batch, stats)["loss"]
        arr[idx] = old
        fd = (up - dn) / (2 * h)
        assert abs(float(par.grad[idx]) - fd) <= 1e-6 * max(1.0, abs(fd)), (idx, float(par.grad[idx]), fd)
    # the backward model's logvar head only sees weight decay (deterministic=True, :228; its l2 term is still summed, :280)
    np.testing.assert_allclose(tr.back.W_lv.grad.numpy(), coeff * wd[-1] * back["W_lv"], rtol=1e-12)
    assert tr.back.max_logvar.grad is None and tr.back.b_lv.grad is None
    # with the forward model's loss switched off, the encoder still receives gradient through the backward model
    tr2 = CaDMTrainer(enc, dyn, back, "halfcheetah", True, (0.,) * 3, (0.,) * 2, 0.0, 1.0, 1e-3, dtype=torch.float64)
    out = tr2.losses(*batch, stats)
    (out["back_mse_loss"]).backward()
    assert float(tr2.enc_W[0].grad.abs().sum()) > 0 and tr2.fwd.W[0].grad is None


def test_flatten_future_matches_the_loop_restatement_with_ragged_masks():
    rng = np.random.default_rng(12)
    D, A, K, F, n = 3, 2, 4, 5, 23
    arrs = dict(obs=rng.standard_normal((n, D * F)), act=rng.standard_normal((n, A * F)), delta=rng.standard_normal((n, D * F)),
                cp_obs=rng.standard_normal((n, D * K)), cp_act=rng.standard_normal((n, A * K)),
                future_bool=(np.arange(F)[None, :] < rng.integers(0, F + 1, n)[:, None]).astype(np.float64),
                obs_next=rng.standard_normal((n, D * F)), back_delta=rng.standard_normal((n, D * F)))
    got = flatten_future(D, A, K, F, *arrs.values())
    want = flatten_future_loops(D, A, K, F, *arrs.values())
    assert got[0].shape[0] == int(arrs["future_bool"].sum())
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    empty = flatten_future(D, A, K, F, *[v[:0] for v in arrs.values()])
    assert [e.shape for e in empty] == [(0, D), (0, A), (0, D), (0, D), (0, D), (0, D * K), (0, A * K)]


class _CpuCadmModel(CaDMModel):
    """The CaDM host mirror without an engine: variables created by the class's own helper, normalisation code inherited."""

    def __init__(self, envname, E=3, H=32, n_hidden=2, K=3, F=2, C=4, cp_hidden=(16,), back_coeff=0.0, deterministic=False,
                 seed=0, **kw):
        self.env = make_env(envname)
        self.env_name = envname
        self.deterministic, self.ensemble_size = deterministic, E
        self.obs_space_dims, self.action_space_dims = D, A = self.env.obs_dim, self.env.act_dim
        self.proc_obs_space_dims = self.env.proc_obs_dim
        self.history_length, self.future_length, self.context_out_dim, self.back_coeff = K, F, C, back_coeff
        self.state_diff = False
        self.discrete, self.normalize_input, self.normalization, self._dataset = False, True, None, None
        self.batch_size, self.learning_rate = kw.get("batch_size", 64), kw.get("learning_rate", 3e-3)
        self.valid_split_ratio, self.rolling_average_persitency = 0.2, kw.get("rolling_average_persitency", 0.99)
        self.weight_decays, self.weight_decay_coeff = (0.,) * (n_hidden + 1), 0.0
        self.context_weight_decays = (0.,) * (len(cp_hidden) + 1)
        rng = np.random.default_rng(seed)
        e = _enc(rng, E, (D + A) * K, cp_hidden, C)
        self._enc = dict(W=[w.astype(np.float32) for w in e["W"]], b=[np.zeros_like(b, dtype=np.float32) for b in e["b"]])
        sizes = [self.proc_obs_space_dims + A + C] + [H] * n_hidden
        self._dyn = self._new_mlp(rng, E, sizes, D)
        self._back = self._new_mlp(rng, E, sizes, D) if back_coeff > 0 else None
        self.engine = None
        self.pushed = 0

    def _push_params(self):
        self.pushed += 1

    def _push_norm(self):
        pass


def _cadm_paths(rng, env, n, K, F):
    """Samples whose dynamics depend on a hidden per-sample scale that only the history reveals."""
    D, A = env.obs_dim, env.act_dim
    scale = rng.choice([0.5, 2.0], size=n)
    M = rng.standard_normal((D + A, D)) * 0.05

    def step(o, a):
        return o + scale[:, None] * (np.concatenate([o, a], axis=1) @ M)

    o = rng.standard_normal((n, D))
    hist_o, hist_a = [], []
    for _ in range(K):
        a = rng.uniform(-1, 1, (n, A))
        hist_o.append(o)
        hist_a.append(a)
        o = step(o, a)
    fut_o, fut_a, fut_n = [], [], []
    for _ in range(F):
        a = rng.uniform(-1, 1, (n, A))
        nx = step(o, a)
        fut_o.append(o)
        fut_a.append(a)
        fut_n.append(nx)
        o = nx
    cat = lambda l: np.concatenate(l, axis=1)
    future_bool = (np.arange(F)[None, :] < rng.integers(1, F + 1, n)[:, None]).astype(np.float64)
    return cat(fut_o), cat(fut_a), cat(fut_n), cat(hist_o), cat(hist_a), future_bool


@pytest.mark.parametrize("back_coeff", [0.0, 0.5])
def test_cadm_fit_loop_learns_and_keeps_the_reference_bookkeeping(back_coeff):
    rng = np.random.default_rng(13)
    K, F = 3, 2
    model = _CpuCadmModel("pendulum", E=3, H=32, K=K, F=F, back_coeff=back_coeff)
    env, D, A = model.env, model.obs_space_dims, model.action_space_dims
    data = _cadm_paths(rng, env, 500, K, F)
    n_params = len(model.params)
    assert n_params == 2 * 2 + (2 * 2 + 6) * (2 if back_coeff > 0 else 1)       # encoder, forward, [backward] -- save() order
    enc0 = [w.copy() for w in model._enc["W"]]
    back0 = None if model._back is None else model._back["W"][0].copy()
    logs = []
    info = fit_cadm_ensemble(model, *data, epochs=30, rng=np.random.default_rng(14), verbose=True, log=logs.append, device="cpu")
    assert info["epochs"] >= 1 and model.pushed == 1
    assert (info["train_back_mse"] > 0) == (back_coeff > 0)
    # statistics from the first step of every sample (:408-411, :439-444), history statistics over the raw history
    assert list(model.normalization) == ["obs", "delta", "act", "cp_obs", "cp_act", "back_delta"]
    obs, act, nxt, cp_obs, cp_act, fb = data
    np.testing.assert_allclose(model.normalization["obs"][0], np.mean(env.obs_preproc(obs[:, :D]), axis=0))
    np.testing.assert_allclose(model.normalization["delta"][1], np.std(env.targ_proc(obs[:, :D], nxt[:, :D]), axis=0))
    np.testing.assert_allclose(model.normalization["back_delta"][0], np.mean(env.targ_proc(nxt[:, :D], obs[:, :D]), axis=0))
    np.testing.assert_allclose(model.normalization["cp_act"][1], np.std(cp_act, axis=0))
    assert set(model._dataset) == {"obs", "act", "delta", "cp_obs", "cp_act", "future_bool", "obs_next", "back_delta",
                                   "single_obs", "single_act", "single_delta", "single_back_delta"}
    assert model._dataset["obs"].shape == (500, D * F)
    # every part was trained: the encoder moved, and so did the backward model when it exists
    assert all(np.abs(w - w0).max() > 0 for w, w0 in zip(model._enc["W"], enc0))
    if back_coeff > 0:
        assert np.abs(model._back["W"][0] - back0).max() > 0
    # the model learned: mse on all valid rows, before (fresh model, same seed) vs after
    stats = model.get_normalization_stats()
    rows = flatten_future(D, A, K, F, obs, act, env.targ_proc(obs.reshape(-1, D), nxt.reshape(-1, D)).reshape(-1, D * F), cp_obs,
                          cp_act, fb, nxt, env.targ_proc(nxt.reshape(-1, D), obs.reshape(-1, D)).reshape(-1, D * F))
    tile = lambda a: np.tile(a[None], (model.ensemble_size, 1, 1))
    fresh = _CpuCadmModel("pendulum", E=3, H=32, K=K, F=F, back_coeff=back_coeff)
    mk = lambda m: CaDMTrainer(m._enc, m._dyn, m._back, "pendulum", False, m.weight_decays, m.context_weight_decays, 0.0,
                               back_coeff, 1e-3)
    mse0 = mk(fresh).evaluate(*map(tile, rows), stats)[0]
    mse1 = mk(model).evaluate(*map(tile, rows), stats)[0]
    assert mse1 < 0.25 * mse0, (mse0, mse1)
    assert any("finished epoch 0" in l and "back mse loss" in l for l in logs)
    data2 = _cadm_paths(rng, env, 100, K, F)
    fit_cadm_ensemble(model, *data2, epochs=1, rng=np.random.default_rng(15), device="cpu", log=logs.append)
    assert model._dataset["single_obs"].shape == (600, D) and model.pushed == 2


@pytest.mark.parametrize("script,stop_after", [([1.0, 1.4, 1.0], 2), ([1.0, 0.9, 5.0, 0.1], 3), ([-4.0, -1.0, -1.0], 2),
                                               ([1.0, 1.0, 1.0, 1.0], 4)])
def test_cadm_early_stopping_bound_follows_the_rolling_average(script, stop_after, monkeypatch):
    """mlp_cadm_ensemble_cem_dynamics.py:540-560 with persistency 0.5: unlike the PE-TS loop, the bound is reset to the
    rolling average at the end of every epoch, so training stops as soon as the rolling average rises.  [1, 1.4]: 1.25 then
    1.325 > 1.25 -> 2 epochs (the PE-TS rule, bound 2.0, would go on); [-4, -1]: -3.33 then -2.17 > -3.33 -> 2 epochs."""
    rng = np.random.default_rng(16)
    model = _CpuCadmModel("pendulum", E=2, H=8, rolling_average_persitency=0.5)
    data = _cadm_paths(rng, model.env, 80, 3, 2)
    it = iter(script)
    monkeypatch.setattr(CaDMTrainer, "evaluate", lambda self, *a: (0.0, 0.0, next(it)))
    logs = []
    info = fit_cadm_ensemble(model, *data, epochs=len(script), rng=np.random.default_rng(17), device="cpu", log=logs.append)
    assert info["epochs"] == stop_after
    assert any("Stopping Training" in l for l in logs) == (stop_after < len(script))


def test_cadm_fit_rejects_the_shapes_the_reference_asserts_on():
    model = _CpuCadmModel("pendulum", E=2, H=8)
    data = list(_cadm_paths(np.random.default_rng(18), model.env, 10, 3, 2))
    for i in range(6):
        bad = list(data)
        bad[i] = bad[i][:, :-1]
        with pytest.raises(AssertionError):
            fit_cadm_ensemble(model, *bad, epochs=1, device="cpu")


# ---------------------------------------------------------------------------------------------------------- optimiser state
def test_tf_adam_matches_the_tf1_update_rule():
    """tf.train.AdamOptimizer (TF 1.x, the reference's optimiser): lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);
    theta -= lr_t m / (sqrt(v) + eps) -- epsilon OUTSIDE the bias correction, unlike torch.optim.Adam."""
    from cadm_b200.dynamics.training import TFAdam
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(7)
    grads = [rng.standard_normal(7) * s for s in (1.0, 1e-6, 3.0, 1e-9)]       # tiny gradients: where the eps placement matters
    p = torch.tensor(p0, dtype=torch.float64, requires_grad=True)
    opt = TFAdam([p], lr=1e-3)
    th, m, v = p0.copy(), np.zeros(7), np.zeros(7)
    for t, g in enumerate(grads, 1):
        p.grad = torch.tensor(g)
        opt.step()
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        th = th - 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
        np.testing.assert_allclose(p.detach().numpy(), th, rtol=1e-13, atol=1e-15)
    assert opt.t == 4


def test_optimizer_state_persists_across_fit_calls():
    """The reference creates its AdamOptimizer once, in the model constructor, so the moment slots and the step count carry
    over from one fit() to the next on the growing dataset (mlp_ensemble_cem_dynamics.py:169).  Here: the
    trainer is created on the first fit() and reused; replacing the parameters from outside rebuilds it."""
    rng = np.random.default_rng(8)
    model = _CpuModel("pendulum", E=2, H=8)
    obs, act, nxt = _transitions(rng, model.env, 200)
    fit_ensemble(model, obs, act, nxt, epochs=2, rng=np.random.default_rng(1), device="cpu")
    tr = model._trainer
    t1 = tr.optimizer.t
    m1 = [m.clone() for m in tr.optimizer.m]
    assert t1 > 0 and any(float(m.abs().max()) > 0 for m in m1)
    obs2, act2, nxt2 = _transitions(rng, model.env, 100)
    fit_ensemble(model, obs2, act2, nxt2, epochs=1, rng=np.random.default_rng(2), device="cpu")
    assert model._trainer is tr and tr.optimizer.t > t1                     # same optimiser, the step count went on
    # the arrays the planner reads are the trainer's tensors
    np.testing.assert_array_equal(model._dyn["W_mu"], tr.W_mu.detach().numpy())
    # parameters replaced from outside (load / set_params bump the version): a fresh trainer from the new arrays
    model._dyn["W_mu"][...] = 0.0
    model._params_version = getattr(model, "_params_version", 0) + 1
    fit_ensemble(model, obs2, act2, nxt2, epochs=1, rng=np.random.default_rng(3), device="cpu")
    assert model._trainer is not tr and model._trainer.optimizer.t < t1
