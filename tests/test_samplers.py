"""Host-side sampler loop and sample processor (cadm_b200/samplers.py) against scenarios recorded from the UNMODIFIED
reference classes (tests/golden/recorded/sampler_golden.npz, generator tests/golden/make_sampler_golden.py): every argument of every
policy.get_actions() call, every finished path and every array of process_samples() must agree bit for bit.  CPU only."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sampler_fakes import FakeEnv, ScriptedPolicy, controller_dispatch_log

from cadm_b200.samplers import (HostPlannerState, IterativeEnvExecutor, ModelSampleProcessor, Sampler, context_rollout_multi,
                                discount_cumsum, rollout_multi)
from oracle.sampler_oracle import future_windows_loops

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "recorded", "sampler_golden.npz"))
SCENARIOS = dict(                      # the generator's table
    cem_ctx_diff=(True, True, True, 3, 4, 3, 12, 5),
    cem_ctx_abs=(True, False, True, 3, 4, 3, 12, 5),
    cem_plain=(False, False, True, 1, 1, 3, 12, 5),
    rs_ctx=(True, True, False, 4, 2, 2, 10, 5),
)


def _run(name):
    context, state_diff, use_cem, K, F, m, T, h = SCENARIOS[name]
    FakeEnv._copies = 0
    env = FakeEnv()
    policy = ScriptedPolicy(h, env.act_dim, use_cem)
    sampler = Sampler(env=env, policy=policy, num_rollouts=m, max_path_length=T, n_parallel=1, use_cem=use_cem, horizon=h,
                      context=context, state_diff=state_diff, history_length=K)
    assert sampler.session is None                                  # no engine behind this policy: NumPy state
    return sampler, policy, sampler.obtain_samples(log=False)


def _same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    assert np.array_equal(a, b), what


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_sampler_feeds_the_policy_and_records_paths_like_the_reference(name):
    g = lambda k: GOLDEN[f"{name}/{k}"]
    sampler, policy, paths = _run(name)
    assert len(policy.calls) == int(g("n_calls")) and len(paths) == int(g("n_paths"))
    for i, call in enumerate(policy.calls):
        want = {k[len(f"{name}/call{i}_"):] for k in GOLDEN.files if k.startswith(f"{name}/call{i}_")}
        assert set(call) == want, i
        for k, v in call.items():
            _same(v, g(f"call{i}_{k}"), (i, k))
    for i, p in enumerate(paths):
        for k in ("observations", "actions", "rewards", "dones", "cp_obs", "cp_act"):
            _same(p[k], g(f"path{i}_{k}"), (i, k))
        _same(p["env_infos"]["t"], g(f"path{i}_env_t"), i)
        assert p["agent_infos"] == {}
    if SCENARIOS[name][2]:
        _same(sampler.prev_sol, g("final_prev_sol"), "prev_sol")
        assert float(sampler.init_var.min()) == float(sampler.init_var.max()) == 0.25
    assert sampler.total_timesteps_sampled == SCENARIOS[name][5] * SCENARIOS[name][6]


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_sample_processor_arrays_match_the_reference(name):
    context, _, _, K, F, m, T, h = SCENARIOS[name]
    g = lambda k: GOLDEN[f"{name}/{k}"]
    _, _, paths = _run(name)
    proc = ModelSampleProcessor(discount=0.99, max_path_length=T, context=context, future_length=F)
    data = proc.process_samples(paths, log=True, log_prefix="x-")
    want = {k.split("/samples_")[1] for k in GOLDEN.files if k.startswith(f"{name}/samples_")}
    assert set(data) == want
    for k in want:
        _same(data[k], g(f"samples_{k}"), k)
    for i, p in enumerate(paths):                                    # short paths were padded in place, returns attached
        assert p["observations"].shape[0] == p["actions"].shape[0] == p["cp_obs"].shape[0] == int(g(f"path{i}_len_after"))
        _same(p["returns"], g(f"path{i}_returns"), i)
    assert proc.last_stats["x-NumTrajs"] == len(paths)
    assert proc.last_stats["x-MaxReturn"] >= proc.last_stats["x-AverageReturn"] >= proc.last_stats["x-MinReturn"]
    if context:
        # the arrays go straight into fit() (mb_trainer.py:196-204): shapes the CaDM model asserts on
        D, A = 3, 2
        n = data["concat_obs"].shape[0]
        assert data["concat_obs"].shape == (n, D * F) and data["concat_act"].shape == (n, A * F)
        assert data["cp_observations"].shape == (n, D * K) and data["concat_bool"].shape == (n, F)
        assert data["observations"].shape[0] <= n                    # padding happens after the single-step arrays are taken


@pytest.mark.parametrize("F", [1, 2, 5])
def test_future_windows_match_the_loop_restatement_on_ragged_paths(F):
    rng = np.random.default_rng(F)
    D, A, K = 4, 3, 2
    lengths = [1, 2, F, F + 1, F + 2, 3 * F + 1, 17]
    mk = lambda: [dict(observations=rng.standard_normal((L, D)), actions=rng.standard_normal((L, A)),
                       rewards=rng.standard_normal(L), cp_obs=rng.standard_normal((L, D * K)),
                       cp_act=rng.standard_normal((L, A * K))) for L in lengths]
    paths = mk()
    originals = [(p["observations"].copy(), p["actions"].copy()) for p in paths]
    data = ModelSampleProcessor(context=True, future_length=F).process_samples(paths)
    want = [future_windows_loops(o, a, F) for o, a in originals]
    for j, k in enumerate(("concat_obs", "concat_act", "concat_next_obs", "concat_bool")):
        np.testing.assert_array_equal(data[k], np.concatenate([w[j] for w in want], axis=0), err_msg=k)
    assert data["concat_obs"].shape[0] == sum(max(L, F + 1) - 1 for L in lengths) == data["cp_observations"].shape[0]
    assert data["observations"].shape[0] == sum(L - 1 for L in lengths)
    # rows whose mask is set describe transitions that happened: next_obs block i == obs block i+1
    o, n, b = data["concat_obs"], data["concat_next_obs"], data["concat_bool"]
    for i in range(F - 1):
        rows = b[:, i + 1] > 0
        np.testing.assert_array_equal(n[rows, i * D:(i + 1) * D], o[rows, (i + 1) * D:(i + 2) * D])


def test_single_step_windows_keep_the_dtype_and_recurrent_stacks():
    rng = np.random.default_rng(0)
    paths = [dict(observations=rng.standard_normal((6, 3)).astype(np.float32), actions=rng.standard_normal((6, 2)).astype(np.float32),
                  rewards=rng.standard_normal(6), cp_obs=np.zeros((6, 3)), cp_act=np.zeros((6, 2))) for _ in range(2)]
    data = ModelSampleProcessor(context=True, future_length=1, recurrent=True).process_samples(paths)
    assert data["concat_obs"].dtype == np.float32 and data["concat_obs"].shape == (2, 5, 3)
    np.testing.assert_array_equal(data["concat_obs"], data["observations"])
    np.testing.assert_array_equal(data["concat_next_obs"], data["next_observations"])
    assert data["concat_bool"][:, 0].sum() == 0 and data["concat_bool"][:, 1:].all()


def test_discount_cumsum_is_the_backward_recursion():
    x = np.array([1.0, 2.0, 3.0])
    np.testing.assert_array_equal(discount_cumsum(x, 0.5), [1 + 0.5 * (2 + 0.5 * 3), 2 + 0.5 * 3, 3.0])
    assert discount_cumsum(np.zeros(0), 0.9).shape == (0,)


def test_host_planner_state_fills_then_slides_and_resets_per_environment():
    st = HostPlannerState(2, obs_dim=2, act_dim=1, history_length=2, state_diff=True, use_cem=True, horizon=3)
    sols = np.arange(6, dtype=np.float64).reshape(2, 3, 1)
    first = st.shift(sols)
    np.testing.assert_array_equal(first, [[0.], [3.]])
    np.testing.assert_array_equal(st.prev_sol[:, :, 0], [[1., 2., 0.], [4., 5., 0.]])
    o = np.zeros((2, 2))
    for step in range(3):
        st.observe(o, np.full((2, 1), step + 1.0), o + (step + 1), np.array([False, step == 1]))
    # env 0: three transitions through a 2-slot buffer -> slots hold the last two; env 1 was cleared after the second
    np.testing.assert_array_equal(st.history_state, [[2., 2., 3., 3.], [3., 3., 0., 0.]])
    np.testing.assert_array_equal(st.history_act, [[2., 3.], [3., 0.]])
    assert list(st.counts) == [3, 1]
    st.reset_plans(1)
    assert not st.prev_sol[1].any() and st.prev_sol[0].any()


def test_executor_resets_on_done_and_on_max_path_length():
    FakeEnv._copies = 0
    ex = IterativeEnvExecutor(FakeEnv(lengths=((2, 50), (50,))), 2, max_path_length=3)
    first = ex.reset()
    acts = np.zeros((2, 2))
    _, _, d1, _ = ex.step(acts)
    obs2, _, d2, _ = ex.step(acts)
    assert list(d1) == [False, False] and list(d2) == [True, False]         # env 0 ends by itself after 2 steps
    assert not np.array_equal(obs2[0], first[0]) and ex.ts[0] == 0           # ... and already shows its next episode
    _, _, d3, _ = ex.step(acts)
    assert list(d3) == [False, True] and list(ex.ts) == [1, 0]                # env 1 hits max_path_length


def test_device_state_needs_an_engine():
    from cadm_b200._lib import CadmError
    FakeEnv._copies = 0
    env = FakeEnv()
    with pytest.raises(CadmError):
        Sampler(env, ScriptedPolicy(5, 2, True), 2, 10, use_cem=True, horizon=5, device_state=True)


def test_paths_flow_from_the_sampler_through_the_processor_into_fit():
    """The trainer's glue (mb_trainer.py:179-211): obtain_samples -> process_samples -> fit(concat_obs, concat_act,
    concat_next_obs, cp_observations, cp_actions, concat_bool).  Dimensions of the pendulum task, stand-in environment."""
    from test_training import _CpuCadmModel
    from cadm_b200.dynamics.training import fit_cadm_ensemble
    K, F = 3, 4
    FakeEnv._copies = 0
    env = FakeEnv(obs_dim=3, act_dim=1)
    policy = ScriptedPolicy(5, 1, True)
    sampler = Sampler(env, policy, num_rollouts=3, max_path_length=12, use_cem=True, horizon=5, context=True, state_diff=True,
                      history_length=K)
    data = ModelSampleProcessor(context=True, future_length=F).process_samples(sampler.obtain_samples())
    model = _CpuCadmModel("pendulum", E=2, H=16, K=K, F=F, back_coeff=0.5)
    info = fit_cadm_ensemble(model, data['concat_obs'], data['concat_act'], data['concat_next_obs'], data['cp_observations'],
                             data['cp_actions'], data['concat_bool'], epochs=3, rng=np.random.default_rng(0), device="cpu",
                             log=lambda *_: None)
    assert info["epochs"] >= 1 and np.isfinite(info["train_recon"]) and model.pushed == 1
    assert model._dataset["future_bool"].shape == data["concat_bool"].shape


EVAL_SCENARIOS = dict(                 # the generator's table
    eval_plain_cem=(False, False, True, 1, 3, 12, 5, 5),
    eval_ctx_cem_diff=(True, True, True, 3, 3, 12, 5, 5),
    eval_ctx_rs_abs=(True, False, False, 4, 2, 10, 5, 4),
)


@pytest.mark.parametrize("name", list(EVAL_SCENARIOS))
def test_evaluation_rollouts_match_the_reference(name):
    """rollout_multi / context_rollout_multi (cadm/samplers/utils.py): every policy call and the returned average."""
    context, state_diff, use_cem, K, m, T, h, total = EVAL_SCENARIOS[name]
    g = lambda k: GOLDEN[f"{name}/{k}"]
    FakeEnv._copies = 0
    env = FakeEnv()
    policy = ScriptedPolicy(h, env.act_dim, use_cem)
    fn = context_rollout_multi if context else rollout_multi
    avg = fn(IterativeEnvExecutor(env, m, T), policy, False, num_rollouts=m, test_total=total, state_diff=state_diff,
             act_dim=env.act_dim, use_cem=use_cem, horizon=h, context=context, history_length=K)
    assert len(policy.calls) == int(g("n_calls"))
    for i, call in enumerate(policy.calls):
        assert set(call) == {k[len(f"{name}/call{i}_"):] for k in GOLDEN.files if k.startswith(f"{name}/call{i}_")}
        for k, v in call.items():
            _same(v, g(f"call{i}_{k}"), (i, k))
    assert float(avg) == float(g("average"))


def test_mpc_controller_dispatches_like_the_reference():
    """cadm/policies/mpc_controller.py: which of (obs, cp_obs, cp_act, init_mean, init_var) reach dynamics_model.get_action,
    and in which order, for every (context, use_cem) combination -- recorded from the reference class."""
    from cadm_b200.policies.mpc_controller import MPCController
    got = controller_dispatch_log(MPCController)
    assert got == str(GOLDEN["mpc_controller/dispatch"])
    assert "('obs', 'cp_obs', 'cp_act', 'mean', 'var')" in got and "('obs', 'mean', 'var')" in got


BASE_SCENARIOS = dict(base_policy=(False, 3, 6), base_random=(True, 2, 5))       # the generator's table


@pytest.mark.parametrize("name", list(BASE_SCENARIOS))
def test_single_environment_sampler_matches_the_reference(name):
    """cadm/samplers/base.py BaseSampler.obtain_samples (the single-environment caller of policy.get_action): what
    the policy is fed, where paths are cut (done, or max_path_length steps), the [1, A] -> [A] action, the stacked env / agent
    infos and the sample accounting, against a recording of the unmodified class."""
    from sampler_fakes import ScriptedSinglePolicy
    from cadm_b200.samplers import BaseSampler
    random, m, T = BASE_SCENARIOS[name]
    g = lambda k: GOLDEN[f"{name}/{k}"]
    FakeEnv._copies = 0
    env = FakeEnv(lengths=((4, 2, 30, 3, 30),))
    policy = ScriptedSinglePolicy(env.act_dim)
    sampler = BaseSampler(env, policy, m, T)
    paths = sampler.obtain_samples(log=False, random=random)
    assert len(policy.calls) == int(g("n_calls")) and len(paths) == int(g("n_paths"))
    assert sampler.total_timesteps_sampled == int(g("total_timesteps_sampled")) == m * T
    for i, obs in enumerate(policy.calls):
        _same(obs, g(f"call{i}_obs"), i)
    for i, p in enumerate(paths):
        for k in ("observations", "actions", "rewards", "dones"):
            _same(p[k], g(f"path{i}_{k}"), (i, k))
        _same(p["env_infos"]["t"], g(f"path{i}_env_t"), i)
        if random:
            assert p["agent_infos"] == {}
        else:
            _same(p["agent_infos"]["s"], g(f"path{i}_agent_s"), i)
    lengths = [len(p["rewards"]) for p in paths]
    assert max(lengths) == T and min(lengths) < T                    # both kinds of path end occur
