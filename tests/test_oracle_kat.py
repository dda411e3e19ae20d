"""Known-answer pins for the oracle.  Each expected value is derived by hand
from the cited reference lines, not from running the oracle."""
import numpy as np
import pytest

from oracle import cadm_oracle as orc
from oracle import philox as ph
from oracle.envs import get_env


def _norm(env, dt=np.float64, delta_mean=None, delta_std=None, K=0):
    P, A, D = env.proc_obs_dim, env.act_dim, env.obs_dim
    return orc.NormStats(np.zeros(P), np.ones(P), np.zeros(A), np.full(A, 0.6),
                         np.zeros(D) if delta_mean is None else delta_mean,
                         np.full(D, 0.1) if delta_std is None else delta_std,
                         np.zeros(D * K), np.ones(D * K), np.zeros(A * K), np.ones(A * K)).astype(dt)


def _zero_params(E, In, H, D, dt=np.float64):
    prm = orc.init_dynamics_params(np.random.default_rng(0), E, In, H, D, dtype=dt)
    for w in prm.W + [prm.W_mu, prm.W_lv]:
        w[...] = 0
    return prm


@pytest.mark.parametrize("envname,coef,bonus", [("halfcheetah", 0.1, 0.0), ("ant", 0.005, 0.05)])
def test_kat1_zero_weights_closed_form(envname, coef, bonus):
    """KAT 1: zero weights, deterministic -> mu = 0 -> delta = mu_delta (utils.py:79); the velocity slot is
    REPLACED by delta[0] (half_cheetah_env.py:54), the reward reads the CURRENT obs[0] (:85), so
    return = o0[0] + (h-1) mu_delta[0] - coef sum_t |a_t|^2 + h*bonus and the elites are the 50 candidates
    with the smallest action energy."""
    env = get_env(envname)
    rng = np.random.default_rng(1)
    m, n, h, E, p = 2, 64, 7, 1, 1
    D, A = env.obs_dim, env.act_dim
    dmean = rng.normal(size=D) * 0.3
    norm = _norm(env, delta_mean=dmean)
    prm = _zero_params(E, env.proc_obs_dim + A, 16, D)
    obs = rng.normal(size=(m, D))
    z = ph.gen_z(3, orc.NUM_CEM_ITERS, m, n, h, A, dtype=np.float64)
    res = orc.cem_plan(obs, np.zeros((m, h, A)), np.full((m, h, A), 0.25), z, prm, norm, env, E, p, True)
    a0 = 0.5 * z[0]                                            # mean 0, cvar = min(.25,.25,.25) -> sigma 0.5
    np.testing.assert_allclose(res.actions[0], a0, rtol=0, atol=1e-15)
    energy = np.sum(a0 ** 2, axis=(2, 3))
    expect = obs[:, :1] + (h - 1) * dmean[0] - coef * energy + h * bonus
    np.testing.assert_allclose(res.returns[0], expect, rtol=1e-12, atol=1e-12)
    want = np.argsort(energy, axis=1, kind="stable")[:, :50]
    assert np.array_equal(res.elites[0], want)


def test_kat2_member_probe_row_map():
    """KAT 2 (quirk Q1): with b_mu of member e equal to e and zero weights, next_obs[..., 0] = e * sigma_delta
    reveals which member served each particle: e = pi // (p/E) (utils.py:144-160)."""
    env = get_env("halfcheetah")
    m, n, h, E, p = 3, 4, 1, 5, 20
    D, A = env.obs_dim, env.act_dim
    norm = _norm(env, delta_std=np.ones(D))
    prm = _zero_params(E, env.proc_obs_dim + A, 8, D)
    for e in range(E):
        prm.b_mu[e] = e
    obs = np.zeros((m, D))
    acts = np.zeros((m, n, h, A))
    _, st = orc.rollout(obs, acts, prm, norm, env, E, p, True, trace=True)
    got = st[0][..., 0]                                        # [m, n, p]
    e_map, r_map = orc.row_maps(m, n, p, E)
    np.testing.assert_allclose(got, e_map * (1.0 + 1e-10), rtol=1e-12)
    assert np.array_equal(e_map[0, 0], np.arange(p) // (p // E))
    # row order j*m*n + mi*n + ni: make the delta depend on the row through the eps input
    eps = np.arange(E * (p // E) * m * n, dtype=np.float64).reshape(1, E, -1, 1) * np.ones(D)
    prm2 = _zero_params(E, env.proc_obs_dim + A, 8, D)
    prm2.b_lv[...] = 0.0
    _, st2 = orc.rollout(obs, acts, prm2, norm, env, E, p, False, eps_it=eps, trace=True)
    lv = 0.5 - np.log1p(np.exp(0.5))
    lv = -10 + np.log1p(np.exp(lv + 10))
    std = np.exp(lv / 2)
    R = (p // E) * m * n
    np.testing.assert_allclose(st2[0][..., 0], (e_map * R + r_map) * std, rtol=1e-12)


@pytest.mark.parametrize("m", [1, 3])
def test_kat3_context_probe(m):
    """KAT 3 (quirks Q2, Q3): a dynamics net that copies the first context feature into delta[0] shows that
    particle pi sees ENCODER member pi % E (utils.py:435), and that on odd CEM iterations the [E, m, C]
    tensor is re-interpreted as [m, E, C] (the transpose at :433-434 sits inside the iteration loop)."""
    env = get_env("halfcheetah")
    n, h, E, p, C = 50, 1, 5, 20, 2
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    norm = _norm(env, delta_std=np.ones(D))
    ctx_raw = (np.arange(E)[:, None, None] * 10.0 + np.arange(m)[None, :, None] + np.zeros((E, m, C)))
    e_map, _ = orc.row_maps(m, n, p, E)
    for it in range(4):
        rows = None
        carried = ctx_raw
        for k in range(it + 1):
            rows, carried = orc._context_rows(carried, k, m, n, p, E, C)
        got = orc._from_rows(rows, p, m, n, C)[..., 0]           # [m, n, p] context value each particle sees
        want = orc.context_map(ctx_raw, it, m, p, E)[..., 0]     # [m, p]
        np.testing.assert_array_equal(got, np.broadcast_to(want[:, None, :], got.shape))
        if it % 2 == 0:
            # proper transpose: value = 10 * (pi % E) + mi
            np.testing.assert_array_equal(want, 10.0 * (np.arange(p) % E)[None, :] + np.arange(m)[:, None])
        elif m > 1:
            flat = ctx_raw[..., 0].reshape(-1)                  # memory order of [E, m]
            exp = np.array([[flat[mi * E + (pi % E)] for pi in range(p)] for mi in range(m)])
            np.testing.assert_array_equal(want, exp)
            assert not np.array_equal(want, 10.0 * (np.arange(p) % E)[None, :] + np.arange(m)[:, None])


def test_kat4_ties_lower_index_first():
    """KAT 4: tf.nn.top_k on equal values returns the lower indices first (utils.py:171)."""
    r = np.zeros((2, 200))
    assert np.array_equal(orc.top_k_desc(r, 50), np.tile(np.arange(50, dtype=np.int32), (2, 1)))
    r[0, 150] = 1.0
    r[0, 7] = 1.0
    assert list(orc.top_k_desc(r, 50)[0][:4]) == [7, 150, 0, 1]


def test_kat5_bounds():
    """KAT 5: |z| <= 2; constrained var (utils.py:131-132); sigma = 0.5 for mean 0 / var 0.25; the clipped
    plan lies in [-1, 1] (mlp_ensemble_cem_dynamics.py:205-206)."""
    z = ph.gen_z(11, 2, 2, 300, 30, 6, dtype=np.float64)
    assert np.abs(z).max() <= 2.0
    assert abs(z.std() - 0.8796) < 5e-3 and abs(z.mean()) < 5e-3
    mean = np.array([[[0.0, 0.9, -0.8]]])
    var = np.full((1, 1, 3), 0.25)
    acts, cvar = orc.sample_actions(mean, var, np.full((1, 4, 1, 3), 2.0))
    np.testing.assert_allclose(cvar[0, 0], [0.25, 0.05 ** 2, 0.1 ** 2])
    assert np.all(np.abs(acts) <= 1.0 + 1e-12)


def test_kat6_logvar_clamp_and_zero_std():
    """KAT 6: huge +-logvar is squashed into (min_logvar, max_logvar) (utils.py:84-85); sigma_delta = 0 gives
    exp((lv + 2 log 0)/2) = 0 so the sample equals the mean (utils.py:87-90)."""
    env = get_env("halfcheetah")
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    E = 2
    prm = _zero_params(E, P + A, 4, D)
    prm.b_lv[0] = 1e4
    prm.b_lv[1] = -1e4
    x = np.zeros((E, 3, P + A))
    norm = _norm(env, delta_std=np.ones(D))
    _, _, lv = orc.forward(x, prm, norm, False, eps=np.ones((E, 3, D)))
    # upper: first squash gives 0.5 - softplus(-1e4) = 0.5, second gives -10 + softplus(10.5)
    np.testing.assert_allclose(lv[0], -10.0 + np.log1p(np.exp(10.5)), rtol=1e-12)
    # lower: 0.5 - softplus(1e4 + 0.5) = -1e4, then -10 + softplus(-1e4 + 10) = -10 (+ e^-9990)
    np.testing.assert_allclose(lv[1], -10.0, rtol=1e-12)
    with np.errstate(divide="ignore"):
        out, _, _ = orc.forward(x, prm, _norm(env, delta_std=np.zeros(D), delta_mean=np.full(D, 2.0)), False,
                                eps=np.ones((E, 3, D)))
    np.testing.assert_allclose(out, 2.0)


def test_kat7_ema_closed_form():
    """KAT 7: one refit by hand (utils.py:171-182): population variance, alpha = 0.1, unconstrained var carried."""
    rng = np.random.default_rng(5)
    m, n, h, A = 1, 60, 2, 3
    actions = rng.normal(size=(m, n, h, A))
    ret = rng.normal(size=(m, n))
    mean0, var0 = rng.normal(size=(m, h, A)), rng.uniform(0.1, 1, size=(m, h, A))
    mean1, var1, idx = orc.refit(mean0, var0, actions, ret)
    order = sorted(range(n), key=lambda i: (-ret[0, i], i))[:50]
    assert list(idx[0]) == order
    el = actions[0, order]
    nm = el.sum(0) / 50
    nv = ((el - nm) ** 2).sum(0) / 50
    np.testing.assert_allclose(mean1[0], 0.1 * mean0[0] + 0.9 * nm, rtol=1e-12)
    np.testing.assert_allclose(var1[0], 0.1 * var0[0] + 0.9 * nv, rtol=1e-12)


def test_kat8_reward_reads_current_obs():
    """KAT 8 (quirk Q4): the step-0 reward is o0[0] - 0.1 |a_0|^2 whatever the model predicts."""
    env = get_env("halfcheetah")
    rng = np.random.default_rng(2)
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    prm = orc.init_dynamics_params(rng, 1, P + A, 32, D, dtype=np.float64)
    norm = _norm(env)
    obs = rng.normal(size=(2, D))
    acts = rng.uniform(-1, 1, size=(2, 5, 1, A))
    pr, _ = orc.rollout(obs, acts, prm, norm, env, 1, 1, True)
    np.testing.assert_allclose(pr[..., 0], obs[:, :1] - 0.1 * np.sum(acts[:, :, 0] ** 2, -1), rtol=1e-12)


def test_kat9_fp32_vs_fp64_budget():
    """KAT 9: the fp32 restatement stays within 1e-5 of fp64 on C2-shaped inputs (so a 1e-4 bar is meaningful)."""
    env = get_env("halfcheetah")
    rng = np.random.default_rng(3)
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    E, p, m, n, h = 5, 20, 1, 50, 30
    prm = orc.init_dynamics_params(rng, E, P + A, 200, D, dtype=np.float64)
    prm.b_lv[...] = -6.0
    norm = _norm(env)
    obs = rng.normal(size=(m, D)) * 0.1
    z = ph.gen_z(0, 1, m, n, h, A, dtype=np.float64)
    eps = ph.gen_eps(0, 1, h, m, n, p, E, D, dtype=np.float64)
    acts, _ = orc.sample_actions(np.zeros((m, h, A)), np.full((m, h, A), 0.25), z[0])
    r64, s64 = orc.rollout(obs, acts, prm, norm, env, E, p, False, eps[0], trace=True)
    f = np.float32
    r32, s32 = orc.rollout(obs.astype(f), acts.astype(f), prm.astype(f), norm.astype(f), env, E, p, False,
                           eps[0].astype(f), trace=True)
    assert r32.dtype == np.float32 and s32.dtype == np.float32
    rms = np.sqrt(np.mean(s64 ** 2, axis=(1, 2, 3)))            # [h, D]
    err = np.max(np.abs(s32 - s64), axis=(1, 2, 3)) / rms
    assert err.max() < 1e-5, err.max()
    assert np.max(np.abs(r32 - r64)) / np.max(np.abs(r64)) < 1e-5


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def run(c, k):
        seed = k[0] | (k[1] << 32)
        return [int(x[0]) for x in ph.philox4x32_10(*[np.array([v]) for v in c], seed)]
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_sharding_invariance():
    """KAT 10 (host side): draws are keyed by the GLOBAL candidate id, so a shard sees the same values."""
    full = ph.gen_z(9, 2, 2, 16, 5, 6)
    for G in (2, 4, 8):
        parts = [ph.gen_z(9, 2, 2, 16 // G, 5, 6, n_offset=g * (16 // G)) for g in range(G)]
        assert np.array_equal(np.concatenate(parts, axis=2), full)


def test_eps_layout_matches_row_map():
    m, n, p, E, D, h = 2, 3, 10, 5, 18, 2
    eps = ph.gen_eps(4, 1, h, m, n, p, E, D, dtype=np.float64)
    e_map, r_map = orc.row_maps(m, n, p, E)
    for (mi, ni, pi) in [(0, 0, 0), (1, 2, 9), (1, 0, 4)]:
        rid = (mi * n + ni) * p + pi
        want = ph.normals_for_rows(4, 0, 1, [rid], D)[0]
        np.testing.assert_array_equal(eps[0, 1, e_map[mi, ni, pi], r_map[mi, ni, pi]], want)
    assert abs(eps.std() - 1.0) < 0.05


def test_rs_plan_argmax_first():
    env = get_env("halfcheetah")
    rng = np.random.default_rng(8)
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    prm = _zero_params(1, P + A, 8, D)
    norm = _norm(env)
    obs = rng.normal(size=(2, D))
    u = ph.gen_uniform_actions(1, 2, 40, 3, A, dtype=np.float64)
    u[0, 5] = 0.0
    u[0, 9] = 0.0                                             # two zero-energy candidates tie -> first wins
    out = orc.rs_plan(obs, u, prm, norm, env, 1, 1, True)
    assert out["best"][0] == 5
    np.testing.assert_array_equal(out["action"][0], u[0, 5, 0])
