"""GPU checks of the rows either side of the planner: the Sampler planning through the device session, and the CaDM
model's fit() handing encoder + forward model to the engine."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("round32", [True, False])
def test_sampler_with_device_session_records_the_same_paths_as_the_host_loop(round32):
    """cadm_b200.samplers.Sampler with device_state=True (PlannerSession: plans and histories stay on the GPU) against
    device_state=False (the reference's loop: NumPy state fed to policy.get_actions every step).  Same engine arithmetic,
    same seeds: the recorded paths must agree bit for bit -- also with float64 observations that float32 cannot hold
    exactly, because the Sampler hands the session the float64 state difference (cadm_session_observe, state_diff = 2)."""
    from sampler_fakes import FakeEnv
    from cadm_b200.policies.mpc_controller import MPCController
    from cadm_b200.samplers import Sampler
    from cadm_b200.synth import build_model
    m, horizon = 2, 30

    def run(device_state):
        model, env, _ = build_model("C3", m_max=m, seed=1, candidates=64)
        policy = MPCController(name="policy", env=env, dynamics_model=model, use_cem=True, n_candidates=64, horizon=horizon,
                               num_rollouts=m, context=True)
        FakeEnv._copies = 0
        fake = FakeEnv(obs_dim=env.obs_dim, act_dim=env.act_dim, lengths=((4, 30), (30,)), round32=round32)
        s = Sampler(fake, policy, num_rollouts=m, max_path_length=13, use_cem=True, horizon=horizon, context=True,
                    state_diff=True, history_length=10, device_state=device_state)
        assert (s.session is not None) == device_state
        return s, s.obtain_samples()

    s_dev, dev = run(True)
    s_host, host = run(False)
    assert len(dev) == len(host) >= 3
    assert max(p["observations"].shape[0] for p in host) == 13          # long enough for the 10-entry history to slide
    for a, b in zip(dev, host):
        for k in ("observations", "actions", "rewards", "dones", "cp_obs", "cp_act"):
            assert np.array_equal(np.asarray(a[k], np.float64), np.asarray(b[k], np.float64)), k
    prev, ho, ha, cnt = s_dev.session.state()
    assert np.array_equal(prev, s_host.prev_sol.astype(np.float32))
    assert np.array_equal(ho, s_host.state.history_state.astype(np.float32))
    assert np.array_equal(ha, s_host.state.history_act.astype(np.float32))
    assert list(cnt) == list(s_host.state.counts)


def test_cadm_fit_hands_encoder_and_forward_model_to_the_engine(tmp_path):
    """fit() of the CaDM model (training.py, PyTorch autograd on the device) followed by the hand-written kernels: the
    engine's context encoder and one-step prediction reproduce the trainer's own forward pass on the trained weights,
    planning runs, and save() holds encoder + forward + backward model."""
    import joblib
    from test_training import _cadm_paths
    from cadm_b200.dynamics.training import CaDMTrainer
    from cadm_b200.synth import build_model
    model, env, _ = build_model("C3", m_max=2, seed=3, candidates=64, back_coeff=0.5, future_length=2)
    D, A, K, F, E = env.obs_dim, env.act_dim, model.history_length, 2, model.ensemble_size
    obs, act, nxt, cp_obs, cp_act, fb = _cadm_paths(np.random.default_rng(0), env, 600, K, F)
    before = [w.copy() for w in model._enc["W"]]
    info = model.fit(obs, act, nxt, cp_obs, cp_act, fb, epochs=15, rng=np.random.default_rng(1))
    assert info["epochs"] >= 1 and np.isfinite(info["train_recon"]) and info["train_back_mse"] > 0
    assert all(np.abs(w - w0).max() > 0 for w, w0 in zip(model._enc["W"], before))
    stats = model.get_normalization_stats()
    tr = CaDMTrainer(model._enc, model._dyn, model._back, model.env_name, False, model.weight_decays,
                     model.context_weight_decays, 0.0, 0.5, 1e-3, device="cuda")
    B = 48
    t = lambda a: torch.as_tensor(np.asarray(a, np.float32)).cuda()
    tile = lambda a: np.tile(np.asarray(a, np.float32)[None, :B], (E, 1, 1))
    with torch.no_grad():
        st = [t(s) for s in stats]
        ctx_t = tr.context(t(tile(cp_obs)), t(tile(cp_act)), st)
        mu_t, lv_t = tr.fwd.forward(t(tile(obs[:, :D])), t(tile(act[:, :A])), st, ctx_t)
    ctx_k = model.get_context_pred(cp_obs[:B].astype(np.float32), cp_act[:B].astype(np.float32))
    assert ctx_k.shape == (E, B, model.context_out_dim)
    assert rel_err(ctx_k, ctx_t.cpu().numpy()) < 1e-4
    _, mu_k, lv_k = model.predict(tile(obs[:, :D]), tile(act[:, :A]), ctx=ctx_t.cpu().numpy(), eps=np.zeros((E, B, D), np.float32))
    assert rel_err(mu_k, mu_t.cpu().numpy()) < 1e-4 and rel_err(lv_k, lv_t.cpu().numpy()) < 1e-4
    plan = model.get_action(obs[:2, :D].astype(np.float32), cp_obs[:2].astype(np.float32), cp_act[:2].astype(np.float32),
                            np.zeros((2, 30, A), np.float32), np.full((2, 30, A), 0.25, np.float32))
    assert plan.shape == (2, 30, A) and np.isfinite(plan).all()
    path = str(tmp_path / "params.pkl")
    model.save(path)
    saved = joblib.load(path)
    assert len(saved) == 8 + 14 + 14                                  # encoder ; forward model ; backward model
    twin, _, _ = build_model("C3", m_max=2, seed=9, candidates=64, back_coeff=0.5, future_length=2)
    twin.load(path)
    assert all(np.array_equal(a, b) for a, b in zip(twin.params, model.params))
