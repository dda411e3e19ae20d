"""fit() of the PE-TS ensemble (cadm_b200/dynamics/training.py) against the NumPy restatement of the reference's training
losses (oracle/train_oracle.py), finite differences, closed forms, and the reference's fit() loop semantics.  CPU only."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from cadm_b200.dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel
from cadm_b200.dynamics.training import EnsembleNLLTrainer, fit_ensemble
from cadm_b200.envs import make_env
from oracle import cadm_oracle as orc
from oracle.train_oracle import pets_losses


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
