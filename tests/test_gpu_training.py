"""fit() on the B200 (csrc/trainer.cu through cadm_b200/dynamics/native_trainer.py) against the autograd restatement of the
reference's training graph in float64 (cadm_b200/dynamics/training.py, itself pinned to oracle/train_oracle.py and to the
reference's own model constructors in tests/test_training.py and tests/test_reference_pinned.py): losses, gradients, Adam
updates over several steps, the whole fit() loop, and the optimiser state across fit() calls.

Tolerances: the kernels compute in fp32 with fixed summation orders; against float64 the losses agree to 2e-5 relative, a
gradient tensor to 2e-5 of its largest entry, parameters after 4 Adam steps to 2e-5 absolute (one step moves a parameter by
about lr = 1e-3)."""
import numpy as np
import pytest
import torch

from test_training import _batch, _cadm_batch, _cadm_paths, _dyn, _enc
from cadm_b200.dynamics.native_trainer import NativeTrainer
from cadm_b200.dynamics.training import CaDMTrainer, EnsembleNLLTrainer
from cadm_b200.envs import make_env

pytestmark = pytest.mark.gpu


def _flat_from_torch(tr, cadm):
    """The torch trainer's parameters (or their gradients) in the native flat layout."""
    def mlp(t, grad):
        g = (lambda p: (torch.zeros_like(p) if p.grad is None else p.grad)) if grad else (lambda p: p)
        out = []
        for W, b in zip(t.W, t.b):
            out += [g(W), g(b).reshape(b.shape[0], -1)]
        out.append(torch.cat([g(t.W_mu), g(t.W_lv)], dim=2))
        out.append(torch.cat([g(t.b_mu).reshape(t.b_mu.shape[0], -1), g(t.b_lv).reshape(t.b_lv.shape[0], -1)], dim=1))
        return out

    def flat(grad):
        g = (lambda p: (torch.zeros_like(p) if p.grad is None else p.grad)) if grad else (lambda p: p)
        parts = []
        fwd = tr.fwd if cadm else tr
        if cadm:
            for W, b in zip(tr.enc_W, tr.enc_b):
                parts += [g(W), g(b).reshape(b.shape[0], -1)]
        parts += mlp(fwd, grad)
        parts += [g(fwd.max_logvar).reshape(-1), g(fwd.min_logvar).reshape(-1)]
        if cadm and tr.back is not None:
            parts += mlp(tr.back, grad)
        return torch.cat([p.detach().reshape(-1) for p in parts]).double().cpu().numpy()
    return flat


def _round32(d):
    """Start both trainers from the same fp32-representable values."""
    if d is None:
        return None
    for k, v in d.items():
        for a in (v if isinstance(v, list) else [v]):
            a[...] = a.astype(np.float32)
    return d


def _segments(nt):
    """(name, start, stop) of every tensor of the flat vector, for per-tensor error reports."""
    segs, pos = [], 0
    c = nt.cfg
    E = c.ensemble

    def add(name, n):
        nonlocal pos
        segs.append((name, pos, pos + n))
        pos += n

    def mlp(prefix, In):
        i_ = In
        for i in range(c.n_hidden):
            add(f"{prefix}.W{i}", E * i_ * c.hidden)
            add(f"{prefix}.b{i}", E * c.hidden)
            i_ = c.hidden
        add(f"{prefix}.W_heads", E * c.hidden * 2 * c.obs_dim)
        add(f"{prefix}.b_heads", E * 2 * c.obs_dim)

    if nt.has_enc:
        i_ = (c.obs_dim + c.act_dim) * c.hist_len
        hs = [h for h in c.enc_hidden if h > 0] + [c.ctx_dim]
        for i, h in enumerate(hs):
            add(f"enc.W{i}", E * i_ * h)
            add(f"enc.b{i}", E * h)
            i_ = h
    In = c.proc_obs_dim + c.act_dim + c.ctx_dim
    mlp("fwd", In)
    add("max_logvar", c.obs_dim)
    add("min_logvar", c.obs_dim)
    if nt.has_back:
        mlp("back", In)
    assert pos == nt.n
    return segs


def _assert_params_close(got, want):
    """Parameters after a few Adam steps.  Adam divides by sqrt(v): where a gradient is within fp32 rounding of zero (dead relu
    units of the encoder, saturated log-variance bounds) the update direction is rounding noise at a fraction of lr = 1e-3, so
    a handful of entries may differ by more than the bulk; no entry may be off by anything like a whole step."""
    d = np.abs(got - want)
    assert d.max() <= 2e-4, d.max()
    assert (d > 2e-5).mean() <= 1e-4, (d > 2e-5).mean()


def _assert_flat_close(nt, got, want, rel, what):
    for name, a, b in _segments(nt):
        scale = max(np.abs(want[a:b]).max(), 1e-12)
        err = np.abs(got[a:b] - want[a:b]).max()
        assert err <= rel * scale, f"{what} {name}: max error {err:.3e} vs scale {scale:.3e}"


PETS_CASES = [("halfcheetah", False, 4), ("ant", True, 2), ("pendulum", False, 1)]


@pytest.mark.parametrize("envname,deterministic,n_hidden", PETS_CASES)
def test_pets_steps_match_autograd(envname, deterministic, n_hidden):
    env = make_env(envname)
    rng = np.random.default_rng(5)
    E, H, N, B = 5, 200, 150, 37                                    # reference width, ragged batch
    dyn = _dyn(rng, E, env.proc_obs_dim + env.act_dim, H, env.obs_dim, n_hidden=n_hidden)
    for b in dyn["b"]:
        b += rng.standard_normal(b.shape) * 0.1
    dyn["b_lv"] += rng.standard_normal(dyn["b_lv"].shape)
    dyn["b_lv"][0] += 3.0                                            # both soft bounds active
    dyn["b_lv"][1] -= 14.0
    _round32(dyn)
    obs, act, delta, stats = _batch(rng, env, 1, N)
    data = tuple(a[0].astype(np.float32) for a in (obs, act, delta))
    stats = tuple(np.asarray(s_, np.float32) for s_ in stats)
    wd, coeff = tuple(1e-4 * (i + 1) for i in range(n_hidden + 1)), 0.5
    ref = EnsembleNLLTrainer(dyn, envname, deterministic, wd, coeff, 1e-3, dtype=torch.float64)
    nt = NativeTrainer(None, dyn, None, envname, env.obs_dim, env.proc_obs_dim, env.act_dim, 0, deterministic, wd, (0.0,), coeff, 0.0, 1e-3)
    nt.begin_fit(data, data, stats)
    ref.begin_fit(data, data, stats)
    flat = _flat_from_torch(ref, cadm=False)
    np.testing.assert_allclose(nt.flat_params(), flat(False), rtol=0, atol=1e-7)
    for step in range(4):
        idx = rng.integers(0, N, size=(E, B))
        want = ref.train_step_idx(idx)
        got = nt.train_step_idx(idx)
        for g_, w_ in zip(got, want):
            assert abs(g_ - w_) <= 2e-5 * max(1.0, abs(w_)), (step, got, want)
        if step == 0:
            _assert_flat_close(nt, nt.flat_grads().astype(np.float64), flat(True), 2e-5, "gradient")
    _assert_params_close(nt.flat_params(), flat(False))
    # losses only: nothing moves, the validation batch may be larger than any training batch
    before = nt.flat_params()
    idx = np.tile(np.arange(N), (E, 1))
    want, got = ref.evaluate_idx(idx), nt.evaluate_idx(idx)
    for g_, w_ in zip(got, want):
        assert abs(g_ - w_) <= 2e-5 * max(1.0, abs(w_))
    assert np.array_equal(before, nt.flat_params())
    m, v, t = nt.adam_state()
    assert t == 4 and np.abs(m).max() > 0 and v.min() >= 0
    out = {k: (np.zeros_like(a) if not isinstance(a, list) else [np.zeros_like(x) for x in a]) for k, a in dyn.items()}
    nt.export(out)
    ref_out = {k: (np.zeros_like(a) if not isinstance(a, list) else [np.zeros_like(x) for x in a]) for k, a in dyn.items()}
    ref.export(ref_out)
    for k in out:
        pairs = zip(out[k], ref_out[k]) if isinstance(out[k], list) else [(out[k], ref_out[k])]
        for a, b in pairs:
            assert a.shape == b.shape and np.abs(a - b).max() <= 2e-4, k
    nt.close()


CADM_CASES = [("halfcheetah", False, 0.0), ("ant", False, 0.5), ("halfcheetah", True, 0.5)]


@pytest.mark.parametrize("envname,deterministic,back_coeff", CADM_CASES)
def test_cadm_steps_match_autograd(envname, deterministic, back_coeff):
    env = make_env(envname)
    rng = np.random.default_rng(6)
    E, H, N, B, K, C = 5, 200, 120, 33, 10, 10                      # reference sizes: encoder 256-128-64 -> 10
    enc = _enc(rng, E, (env.obs_dim + env.act_dim) * K, (256, 128, 64), C)
    In = env.proc_obs_dim + env.act_dim + C
    dyn = _dyn(rng, E, In, H, env.obs_dim, n_hidden=2)
    back = _dyn(rng, E, In, H, env.obs_dim, n_hidden=2) if back_coeff > 0 else None
    dyn["b_lv"][0] += 3.0
    dyn["b_lv"][1] -= 14.0
    _round32(enc), _round32(dyn), _round32(back)
    batch, stats = _cadm_batch(rng, env, 1, N, K)
    data = tuple(a[0].astype(np.float32) for a in batch)
    stats = tuple(np.asarray(s_, np.float32) for s_ in stats)
    wd, cwd, coeff = (1e-4, 2e-4, 3e-4), (5e-4, 6e-4, 7e-4, 8e-4), 0.5
    ref = CaDMTrainer(enc, dyn, back, envname, deterministic, wd, cwd, coeff, back_coeff, 1e-3, dtype=torch.float64)
    nt = NativeTrainer(enc, dyn, back, envname, env.obs_dim, env.proc_obs_dim, env.act_dim, K, deterministic, wd, cwd, coeff, back_coeff, 1e-3)
    nt.begin_fit(data, data, stats)
    ref.begin_fit(data, data, stats)
    flat = _flat_from_torch(ref, cadm=True)
    np.testing.assert_allclose(nt.flat_params(), flat(False), rtol=0, atol=1e-7)
    for step in range(4):
        idx = rng.integers(0, N, size=(E, B))
        want = ref.train_step_idx(idx)
        got = nt.train_step_idx(idx)
        assert len(got) == len(want) == 3
        for g_, w_ in zip(got, want):
            assert abs(g_ - w_) <= 2e-5 * max(1.0, abs(w_)), (step, got, want)
        assert (got[1] > 0) == (back_coeff > 0)
        if step == 0:
            _assert_flat_close(nt, nt.flat_grads().astype(np.float64), flat(True), 2e-5, "gradient")
    _assert_params_close(nt.flat_params(), flat(False))
    nt.close()


def test_fit_on_the_device_matches_the_autograd_fit():
    """The whole PE-TS fit() loop twice from the same seed: hand-written kernels (the default on a CUDA engine) and the autograd
    restatement on the CPU -- same bootstrap, same epochs until the early stop, the same losses and weights to fp32."""
    from cadm_b200.synth import build_model
    rng = np.random.default_rng(0)

    def data(env):
        D, A, N = env.obs_dim, env.act_dim, 900
        obs = rng.standard_normal((N, D)) * 0.5
        act = rng.uniform(-1, 1, (N, A))
        M = np.random.default_rng(3).standard_normal((D + A, D)) * 0.05
        nxt = obs + np.concatenate([obs, act], axis=1) @ M
        return obs, act, nxt

    a, env, _ = build_model("C2", m_max=1, seed=4, candidates=64)
    b, _, _ = build_model("C2", m_max=1, seed=4, candidates=64)
    obs, act, nxt = data(env)
    from cadm_b200.dynamics.training import fit_ensemble
    ia = a.fit(obs, act, nxt, epochs=6, rng=np.random.default_rng(1))
    ib = fit_ensemble(b, obs, act, nxt, epochs=6, rng=np.random.default_rng(1), device="cpu")
    assert type(a._trainer).__name__ == "NativeTrainer" and type(b._trainer).__name__ == "EnsembleNLLTrainer"
    assert a._trainer.launches > 0
    assert ia["epochs"] == ib["epochs"]
    assert abs(ia["train_recon"] - ib["train_recon"]) <= 1e-3 * max(1.0, abs(ib["train_recon"]))
    for x, y in zip(a.params, b.params):
        assert np.abs(x - y).max() <= 5e-4                           # ~40 Adam steps of fp32 against fp32-on-CPU
    # a second fit() continues with the optimiser state of the first (the reference builds its optimiser once)
    t0 = a._trainer.adam_state()[2]
    a.fit(obs[:200], act[:200], nxt[:200], epochs=1, rng=np.random.default_rng(2))
    assert a._trainer.adam_state()[2] > t0 > 0


def test_cadm_fit_on_the_device_learns():
    from cadm_b200.synth import build_model
    model, env, _ = build_model("C3", m_max=2, seed=3, candidates=64, back_coeff=0.5, future_length=2)
    K, F = model.history_length, 2
    obs, act, nxt, cp_obs, cp_act, fb = _cadm_paths(np.random.default_rng(0), env, 600, K, F)
    first = model.fit(obs, act, nxt, cp_obs, cp_act, fb, epochs=1, rng=np.random.default_rng(1))
    assert type(model._trainer).__name__ == "NativeTrainer"
    later = model.fit(obs[:1], act[:1], nxt[:1], cp_obs[:1], cp_act[:1], fb[:1], epochs=12, rng=np.random.default_rng(2))
    assert np.isfinite(later["train_recon"]) and later["train_mse"] < first["train_mse"]
    assert later["train_back_mse"] > 0
