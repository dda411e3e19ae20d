"""The remaining environment epilogues, random shooting with discrete actions, the drop-in loop of
the reference sampler, and the joblib checkpoint format -- on the GPU, against the oracle."""
import numpy as np
import pytest
import torch

from oracle import cadm_oracle as orc
from oracle import philox as ph

from helpers import oracle_pack, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _pets(envname, precision, n=64, h=8, E=5, p=10, use_cem=True, deterministic=False, m_max=2, **kw):
    from cadm_b200.dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel
    from cadm_b200.envs import make_env
    from cadm_b200.synth import synthetic_normalization
    env = make_env(envname)
    model = MLPEnsembleCEMDynamicsModel("dm", env, hidden_nonlinearity="swish", n_forwards=h, n_candidates=n, ensemble_size=E,
                                        n_particles=p, use_cem=use_cem, deterministic=deterministic, m_max=m_max,
                                        precision=precision, **kw)
    if not deterministic:
        model._dyn["b_lv"][...] = -6.0
        model._push_params()
    model.set_normalization(synthetic_normalization(env, False))
    return model, env


@pytest.mark.parametrize("precision", ["fp32", "tc3x"])
@pytest.mark.parametrize("envname", ["slim_humanoid", "pendulum", "cripple_halfcheetah"])
def test_other_env_epilogues(envname, precision):
    model, env = _pets(envname, precision)
    prm, enc, norm, oenv = oracle_pack(model)
    rng = np.random.default_rng(1)
    m, n, h, E, p = 2, 64, 8, 5, 10
    D, A = env.obs_dim, env.act_dim
    obs = (0.3 * rng.standard_normal((m, D))).astype(np.float32)
    if envname == "slim_humanoid":
        obs[:, 1] = [1.5, 0.5]                              # alive bonus on / off
    actions = rng.uniform(-1, 1, (m, n, h, A)).astype(np.float32)
    if envname == "pendulum":
        actions *= 3.0                                      # exercises the torque clip at +-2
    eps = ph.gen_eps(2, 1, h, m, n, p, E, D)[0]
    o_ret, o_st = orc.rollout(obs.astype(np.float64), actions.astype(np.float64), prm, norm, oenv, E, p, False,
                              eps.astype(np.float64), trace=True)
    pr, st = model.engine.rollout(obs, actions, None, eps, trace=True)
    assert rel_err(st.cpu().numpy(), o_st, axis=(1, 2, 3)) < TOL
    assert np.max(np.abs(pr.cpu().numpy() - o_ret)) / np.max(np.abs(o_ret)) < TOL


@pytest.mark.parametrize("precision", ["fp32", "tc3x"])
def test_cartpole_discrete_random_shooting(precision):
    model, env = _pets("cartpole", precision, n=128, h=10, use_cem=False)
    assert model.discrete
    prm, enc, norm, oenv = oracle_pack(model)
    rng = np.random.default_rng(3)
    m, n, h, E, p = 2, 128, 10, 5, 10
    obs = (0.05 * rng.standard_normal((m, 4))).astype(np.float32)
    obs[1, 0] = 2.35                                        # close to the x threshold -> rewards differ
    u = ph.gen_discrete_actions(5, m, n, h, 2)
    eps = ph.gen_eps(5, 1, h, m, n, p, E, 4)[0]
    ref = orc.rs_plan(obs.astype(np.float64), u, prm, norm, oenv, E, p, False, eps.astype(np.float64), discrete=True)
    out = model.engine.plan_rs(obs, u=u, eps=eps)
    np.testing.assert_allclose(out["returns"].cpu().numpy(), ref["returns"], atol=1e-4)
    # returns are counts / p: ties are common, argmax must pick the first maximum like tf.argmax
    assert np.array_equal(out["best"].cpu().numpy(), ref["best"])
    assert np.array_equal(out["action"].cpu().numpy(), ref["action"])
    # seed-only path uses the same integer stream as the NumPy specification
    out2 = model.engine.plan_rs(obs, seed=5, eps=eps)
    assert np.array_equal(out2["best"].cpu().numpy(), ref["best"])
    act = model.get_action(obs)
    assert act.shape == (m,) and set(np.unique(act)) <= {0, 1}


def test_drop_in_sampler_loop():
    """The calling pattern of cadm/samplers/sampler.py:107-120 and :164-178 with a stub vectorised env: MPCController
    -> get_actions(obses, init_mean=prev_sol, init_var, cp_obs, cp_act); warm-start shift; K-step history buffers."""
    from cadm_b200.policies.mpc_controller import MPCController
    from cadm_b200.synth import build_model
    m, horizon, K = 4, 30, 10
    model, env, cfg = build_model("C3", m_max=m, seed=1)
    policy = MPCController(name="policy", env=env, dynamics_model=model, use_cem=True, n_candidates=200, horizon=horizon,
                           num_rollouts=m, context=True)
    A, D = env.act_dim, env.obs_dim
    prev_sol = np.tile(0., [m, horizon, A])
    init_var = np.tile(np.square(2) / 16, [m, horizon, A])
    history_state = np.zeros((m, D * K))
    history_act = np.zeros((m, A * K))
    rng = np.random.default_rng(0)
    obses = 0.1 * rng.standard_normal((m, D))                # float64, like env observations
    for step in range(3):
        sols, infos = policy.get_actions(obses, init_mean=prev_sol, init_var=init_var, cp_obs=history_state, cp_act=history_act)
        assert infos == {} and sols.shape == (m, horizon, A) and sols.dtype == np.float32
        assert np.abs(sols).max() <= 1.0 and np.isfinite(sols).all()
        prev_sol[:, :-1] = sols[:, 1:].copy()
        prev_sol[:, -1:] = 0.
        actions = sols[:, 0].copy()
        next_obses = obses + 0.01 * rng.standard_normal((m, D))          # stub env step
        history_state[:, step * D:(step + 1) * D] = next_obses - obses
        history_act[:, step * A:(step + 1) * A] = actions
        obses = next_obses
    # the plan moved away from the zero initialisation and differs between environments
    assert np.abs(sols).mean() > 1e-3 and not np.allclose(sols[0], sols[1])
    # single-observation entry point (mpc_controller.py:43-53) on a non-context model
    model2, env2, _ = build_model("C2", m_max=1)
    pol2 = MPCController(name="p", env=env2, dynamics_model=model2, use_cem=True, horizon=30)
    a, _ = pol2.get_action(obses[0], init_mean=prev_sol[:1], init_var=init_var[:1])
    assert a.shape == (1, horizon, A)
    with pytest.raises(ValueError):
        model2.get_action(obses[:1])                       # use_cem=True needs the CEM feeds (TF: placeholder not fed)
    with pytest.raises(ValueError):
        model2.get_action(obses[:1, :5], prev_sol[:1], init_var[:1])


def test_checkpoint_roundtrip_reference_format(tmp_path):
    """save()/load() use the reference's joblib layout (list of arrays in trainable_variables() order + _norm_stats,
    mlp_cadm_ensemble_cem_dynamics.py:571-588); extra backward-model variables at the tail are ignored."""
    import joblib
    from cadm_b200.synth import build_model, synthetic_inputs
    a, env, cfg = build_model("C3", m_max=1, seed=3)
    path = str(tmp_path / "params_epoch_7")
    a.save(path)
    saved = joblib.load(path)
    assert isinstance(saved, list) and len(saved) == 8 + 8 + 4 + 2           # encoder W,b x4 ; hidden W,b x4 ; heads ; logvar bounds
    assert saved[0].shape == (5, (18 + 6) * 10, 256) and saved[8].shape == (5, 18 + 6 + 10, 200)
    joblib.dump(saved + [np.zeros((5, 34, 200), np.float32)] * 3, path)      # pretend a backward model follows
    b, _, _ = build_model("C3", m_max=1, seed=99)
    b.load(path)
    inp = synthetic_inputs(env, 1, 30, True, seed=2)
    seed = 123
    pa = a.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], inp["cp_obs"], inp["cp_act"], seed=seed)
    pb = b.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], inp["cp_obs"], inp["cp_act"], seed=seed)
    assert torch.equal(pa["mean"], pb["mean"]) and torch.equal(pa["elites"], pb["elites"])
    assert list(b.normalization.keys()) == list(a.normalization.keys())


@pytest.mark.parametrize("config", ["C3", "C2"])
def test_session_matches_host_loop(config):
    """Sampler-side state on the device (cadm_session_*): warm-start shift, constant init_var, K-step history buffers
    filling left to right and then sliding, per-episode resets -- against the NumPy loop of cadm/samplers/sampler.py:107-195
    driving get_actions() with host arrays.  Same engine arithmetic, same seeds: the actions must agree bit for bit."""
    from cadm_b200.policies.mpc_controller import MPCController
    from cadm_b200.samplers import PlannerSession
    from cadm_b200.synth import build_model
    m, horizon = 3, 30
    ctx = config == "C3"
    host_model, env, cfg = build_model(config, m_max=m, seed=1, candidates=64)
    dev_model, _, _ = build_model(config, m_max=m, seed=1, candidates=64)
    policy = MPCController(name="policy", env=env, dynamics_model=host_model, use_cem=True, n_candidates=64, horizon=horizon,
                           num_rollouts=m, context=ctx)
    A, D = env.act_dim, env.obs_dim
    K = 10 if ctx else 1
    session = PlannerSession(dev_model, m, state_diff=True)
    prev_sol = np.tile(0., [m, horizon, A])
    init_var = np.tile(np.square(2) / 16, [m, horizon, A])
    history_state = np.zeros((m, D * K))
    history_act = np.zeros((m, A * K))
    counts = [0] * m
    rng = np.random.default_rng(0)
    obses = (0.1 * rng.standard_normal((m, D))).astype(np.float32).astype(np.float64)     # fp32-representable, like env output
    for step in range(K + 4):
        if ctx:
            sols, _ = policy.get_actions(obses, init_mean=prev_sol, init_var=init_var, cp_obs=history_state, cp_act=history_act)
        else:
            sols, _ = policy.get_actions(obses, init_mean=prev_sol, init_var=init_var)
        prev_sol[:, :-1] = sols[:, 1:].copy()
        prev_sol[:, -1:] = 0.
        actions = sols[:, 0].copy()
        got = session.act(obses)
        assert np.array_equal(got, actions), step
        next_obses = (obses + 0.01 * rng.standard_normal((m, D))).astype(np.float32).astype(np.float64)   # stub env step
        dones = np.zeros(m, bool)
        if step == K + 1:
            dones[1] = True                                   # one episode ends after the buffers have started sliding
        for idx in range(m):                                  # sampler.py:164-195
            if counts[idx] < K:
                history_state[idx][counts[idx] * D:(counts[idx] + 1) * D] = next_obses[idx] - obses[idx]
                history_act[idx][counts[idx] * A:(counts[idx] + 1) * A] = actions[idx]
            else:
                history_state[idx][:-D] = history_state[idx][D:]
                history_state[idx][-D:] = next_obses[idx] - obses[idx]
                history_act[idx][:-A] = history_act[idx][A:]
                history_act[idx][-A:] = actions[idx]
            if dones[idx]:
                prev_sol[idx] = 0.
                counts[idx] = 0
                history_state[idx] = 0.
                history_act[idx] = 0.
            else:
                counts[idx] += 1
        session.observe(next_obses, dones)
        p, ho, ha, cnt = session.state()
        assert np.array_equal(p, prev_sol.astype(np.float32)), step
        if ctx:
            assert np.array_equal(ho, history_state.astype(np.float32)), step
            assert np.array_equal(ha, history_act.astype(np.float32)), step
            assert list(cnt) == counts
        obses = next_obses
    session.reset(idx=[0])
    p, ho, ha, cnt = session.state()
    assert not p[0].any() and not ho[0].any() and cnt[0] == 0 and p[2].any()


@pytest.mark.parametrize("variant", ["1", "2"])
def test_tc1x_fast_mode_runs_and_is_fp16_class(variant, monkeypatch):
    """`precision="tc1x"` (single fp16 pass, does NOT claim the 1e-4 bar) on both tensor-core kernels: finite, and within
    the fp16-class error one expects from 10-bit operands over an 8-step rollout."""
    from oracle import cadm_oracle as orc
    from oracle import philox as ph
    from helpers import oracle_pack
    monkeypatch.setenv("CADM_TC_VARIANT", variant)
    model, env = _pets("halfcheetah", "tc1x", n=64, h=8)
    prm, enc, norm, oenv = oracle_pack(model)
    rng = np.random.default_rng(0)
    m, n, h, p, E = 2, 64, 8, 10, 5
    obs = (0.1 * rng.standard_normal((m, env.obs_dim))).astype(np.float32)
    actions = rng.uniform(-1, 1, (m, n, h, env.act_dim)).astype(np.float32)
    eps = ph.gen_eps(5, 1, h, m, n, p, E, env.obs_dim)[0]
    o_ret, _ = orc.rollout(obs.astype(np.float64), actions.astype(np.float64), prm, norm, oenv, E, p, False, eps.astype(np.float64), None)
    pr, _ = model.engine.rollout(obs, actions, None, eps)
    pr = pr.cpu().numpy()
    assert np.isfinite(pr).all()
    err = np.max(np.abs(pr - o_ret)) / np.max(np.abs(o_ret))
    assert err < 2e-2, err


@pytest.mark.parametrize("precision,variant", [("fp32", "0"), ("tc3x", "1"), ("tc3x", "2")])
@pytest.mark.parametrize("hidden", [(64, 64), (128, 128, 128), (96, 96, 96, 96, 96), (200, 200), (208,)])
def test_other_architectures(hidden, precision, variant, monkeypatch):
    """Hidden widths / depths other than the reference's 4 x 200 (one M tile instead of two, widths that are not a multiple of
    16, 1..5 hidden layers): the table-driven schedules of both tensor-core kernels and the FFMA path against the oracle."""
    monkeypatch.setenv("CADM_TC_VARIANT", variant)
    model, env = _pets("halfcheetah", precision, n=64, h=6, hidden_sizes=hidden)
    prm, enc, norm, oenv = oracle_pack(model)
    rng = np.random.default_rng(1)
    m, n, h, p, E = 2, 64, 6, 10, 5
    obs = (0.1 * rng.standard_normal((m, env.obs_dim))).astype(np.float32)
    actions = rng.uniform(-1, 1, (m, n, h, env.act_dim)).astype(np.float32)
    eps = ph.gen_eps(7, 1, h, m, n, p, E, env.obs_dim)[0]
    o_ret, o_st = orc.rollout(obs.astype(np.float64), actions.astype(np.float64), prm, norm, oenv, E, p, False, eps.astype(np.float64),
                              None, trace=True)
    pr, st = model.engine.rollout(obs, actions, None, eps, trace=True)
    assert rel_err(st.cpu().numpy(), o_st, axis=(1, 2, 3)) < TOL
    assert np.max(np.abs(pr.cpu().numpy() - o_ret)) / np.max(np.abs(o_ret)) < TOL


def test_fit_then_plan():
    """fit() (PyTorch autograd on the device, cadm_b200/dynamics/training.py) hands its weights to the engine: after training on
    synthetic transitions the hand-written kernels reproduce the trainer's own forward pass, the model has learned the
    dynamics, and planning runs on the new weights."""
    from cadm_b200.dynamics.training import EnsembleNLLTrainer
    model, env = _pets("halfcheetah", "tc3x", n=64, h=8, E=5, p=10)
    rng = np.random.default_rng(0)
    D, A, N = env.obs_dim, env.act_dim, 2000
    obs = rng.standard_normal((N, D)) * 0.5
    act = rng.uniform(-1, 1, (N, A))
    M = rng.standard_normal((D + A, D)) * 0.05
    nxt = obs + np.concatenate([obs, act], axis=1) @ M
    nxt[:, 0] = (np.concatenate([obs, act], axis=1) @ M)[:, 0]        # HalfCheetah postproc: o[0] is predicted directly
    info = model.fit(obs, act, nxt, epochs=80, rng=np.random.default_rng(1))
    assert info["epochs"] >= 1
    stats = model.get_normalization_stats()[:6]
    tr = EnsembleNLLTrainer(model._dyn, "halfcheetah", False, model.weight_decays, 0.0, 1e-3, device="cuda")
    E, B = 5, 64
    bo = np.tile(obs[None, :B], (E, 1, 1)).astype(np.float32)
    ba = np.tile(act[None, :B], (E, 1, 1)).astype(np.float32)
    with torch.no_grad():
        mu_t, lv_t = tr.forward(torch.from_numpy(bo).cuda(), torch.from_numpy(ba).cuda(), tr._norm(stats))
    nxt_k, mu_k, lv_k = model.predict(bo, ba, eps=np.zeros((E, B, D), np.float32))
    assert rel_err(mu_k, mu_t.cpu().numpy()) < 1e-4 and rel_err(lv_k, lv_t.cpu().numpy()) < 1e-4
    # the trained model predicts the synthetic dynamics far better than chance (targets have unit variance after normalisation)
    delta = env.targ_proc(obs[:B], nxt[:B])
    pred = mu_k[0] * (stats[5] + 1e-10) + stats[4]
    assert np.mean((pred - delta) ** 2) < 0.5 * np.mean((delta - delta.mean(0)) ** 2)
    plan = model.get_action(obs[:2].astype(np.float32), np.zeros((2, 8, A), np.float32), np.full((2, 8, A), 0.25, np.float32))
    assert plan.shape == (2, 8, A) and np.isfinite(plan).all()
