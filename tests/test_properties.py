"""Property tests (hypothesis) of the invariants SURVEY section 4 lists for the planner path, on the CPU: the oracle (the checker
of the GPU tests), the noise specification and the host-side logic.  The same invariants are checked on the device by the
parity suite at fixed sizes (sharding bit-identity, tile sizes, m batching); here hypothesis picks the shapes.

  - sampled actions stay inside [-1, 1] and truncated normals inside [-2, 2], whatever mean / variance / counters
  - candidate sharding: rolling out the shards separately and concatenating equals the unsharded rollout; the Philox draws of
    a shard are the unsharded draws of its candidates
  - m batching: an environment's returns do not depend on which other environments are planned with it
  - tf.nn.top_k's order (descending, ties -> lower index), the EMA refit's bounds
  - environment shard bounds partition [0, m); the flattening of future steps keeps exactly the rows whose mask is positive
  - the flat parameter vector of the native trainer round-trips through pack / unpack
"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import cadm_oracle as orc
from oracle import philox as ph
from oracle.envs import get_env

SET = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


def _norm(env, rng):
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    return orc.NormStats(obs_mean=rng.normal(size=P) * 0.1, obs_std=rng.uniform(0.5, 1.5, P), act_mean=np.zeros(A), act_std=np.full(A, 0.6),
                         delta_mean=rng.normal(size=D) * 0.01, delta_std=rng.uniform(0.05, 0.2, D))


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), m=st.integers(1, 3), n=st.integers(1, 9), h=st.integers(1, 5), A=st.integers(1, 6),
       var_scale=st.floats(1e-6, 50.0))
def test_sampled_actions_stay_in_the_action_box(seed, m, n, h, A, var_scale):
    """core/utils.py:131-135: with the constrained variance and |z| <= 2 every candidate lies in [-1, 1] for any mean in the
    box and any variance -- the reason the reference never clips inside the planner."""
    rng = np.random.default_rng(seed)
    mean = rng.uniform(-1, 1, (m, h, A))
    var = rng.uniform(0, var_scale, (m, h, A))
    z = ph.gen_z(seed, 1, m, n, h, A, dtype=np.float64)[0]
    assert np.abs(z).max() <= 2.0
    actions, cvar = orc.sample_actions(mean, var, z)
    assert actions.shape == (m, n, h, A)
    assert np.all(actions >= -1 - 1e-12) and np.all(actions <= 1 + 1e-12)
    assert np.all(cvar <= var + 1e-15) and np.all(cvar >= 0)


@settings(**SET)
@given(seed=st.integers(0, 2 ** 63 - 1), it=st.integers(0, 4), m=st.integers(1, 3), G=st.sampled_from([1, 2, 4, 8]), k=st.integers(1, 4),
       h=st.integers(1, 4), A=st.integers(1, 7), m_offset=st.integers(0, 5))
def test_philox_draws_are_keyed_by_global_ids(seed, it, m, G, k, h, A, m_offset):
    """A shard's draws are the unsharded draws of its candidates (n_offset) and of its environments (m_offset)."""
    n = G * k
    full = ph.gen_z(seed, it + 1, m + m_offset, n, h, A)
    assert np.abs(full).max() <= 2.0 and np.isfinite(full).all()
    parts = [ph.gen_z(seed, it + 1, m, k, h, A, n_offset=g * k, m_offset=m_offset) for g in range(G)]
    assert np.array_equal(np.concatenate(parts, axis=2), full[:, m_offset:])


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), envname=st.sampled_from(["halfcheetah", "ant", "pendulum"]), m=st.integers(1, 3), G=st.sampled_from([1, 2, 4]),
       k=st.integers(1, 3), q=st.integers(1, 2), E=st.sampled_from([1, 2, 5]), h=st.integers(1, 3), deterministic=st.booleans())
def test_candidate_sharding_and_env_batching_do_not_change_returns(seed, envname, m, G, k, q, E, h, deterministic):
    """core/utils.py:137-170 couples candidates only through the elite selection: the particle returns of a shard, computed alone,
    are those rows of the unsharded rollout; and an environment planned alone gets the returns it gets inside a batch."""
    env = get_env(envname)
    rng = np.random.default_rng(seed)
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    n, p = G * k, E * q
    prm = orc.init_dynamics_params(rng, E, P + A, 12, D, n_hidden=2, dtype=np.float64)
    norm = _norm(env, rng)
    obs = rng.normal(size=(m, D)) * 0.3
    actions = rng.uniform(-1, 1, (m, n, h, A))
    eps = ph.gen_eps(seed, 1, h, m, n, p, E, D, dtype=np.float64)[0]                 # [h, E, q m n, D], rows (j, mi, ni)
    full, _ = orc.rollout(obs, actions, prm, norm, env, E, p, deterministic, eps, None)
    assert full.shape == (m, n, p) and np.isfinite(full).all()
    e5 = eps.reshape(h, E, q, m, n, D)
    for g in range(G):
        sl = slice(g * k, (g + 1) * k)
        part, _ = orc.rollout(obs, actions[:, sl], prm, norm, env, E, p, deterministic, e5[:, :, :, :, sl].reshape(h, E, q * m * k, D), None)
        assert np.array_equal(part, full[:, sl])
    for mi in range(m):
        alone, _ = orc.rollout(obs[mi:mi + 1], actions[mi:mi + 1], prm, norm, env, E, p, deterministic,
                               e5[:, :, :, mi:mi + 1].reshape(h, E, q * n, D), None)
        assert np.array_equal(alone[0], full[mi])


@settings(**SET)
@given(data=st.data(), n=st.integers(1, 40), k=st.integers(1, 12))
def test_top_k_is_descending_with_ties_to_the_lower_index(data, n, k):
    k = min(k, n)
    vals = np.array(data.draw(st.lists(st.integers(-3, 3), min_size=n, max_size=n)), dtype=np.float64)      # many ties
    idx = orc.top_k_desc(vals[None, :], k)[0]
    assert len(set(idx.tolist())) == k
    for a, b in zip(idx[:-1], idx[1:]):
        assert vals[a] > vals[b] or (vals[a] == vals[b] and a < b)
    rest = np.setdiff1d(np.arange(n), idx)
    if len(rest):
        assert vals[rest].max() <= vals[idx[-1]]
        assert not np.any((vals[rest] == vals[idx[-1]]) & (rest < idx[-1]))


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), m=st.integers(1, 3), n=st.integers(2, 30), h=st.integers(1, 4), A=st.integers(1, 4),
       alpha=st.floats(0.0, 1.0))
def test_refit_is_a_convex_step_towards_the_elites(seed, m, n, h, A, alpha):
    rng = np.random.default_rng(seed)
    k = min(5, n)
    mean, var = rng.uniform(-1, 1, (m, h, A)), rng.uniform(0, 1, (m, h, A))
    actions = rng.uniform(-1, 1, (m, n, h, A))
    returns = rng.normal(size=(m, n))
    new_mean, new_var, idx = orc.refit(mean, var, actions, returns, num_elites=k, alpha=alpha)
    assert idx.shape == (m, k)
    el = np.stack([actions[i, idx[i]] for i in range(m)])
    lo, hi = np.minimum(mean, el.min(1)), np.maximum(mean, el.max(1))
    assert np.all(new_mean >= lo - 1e-12) and np.all(new_mean <= hi + 1e-12)
    assert np.all(new_var >= -1e-15)
    if alpha == 1.0:
        assert np.array_equal(new_mean, mean) and np.array_equal(new_var, var)


@settings(**SET)
@given(m=st.integers(0, 50), world=st.integers(1, 9))
def test_environment_shard_bounds_partition_the_environments(m, world):
    from cadm_b200.parallel import env_shard_bounds
    b = env_shard_bounds(m, world)
    assert len(b) == world + 1 and b[0] == 0 and b[-1] == m
    sizes = np.diff(b)
    assert np.all(sizes >= 0) and sizes.max() - sizes.min() <= 1


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 12), F=st.integers(1, 4), K=st.integers(1, 3))
def test_flatten_future_keeps_exactly_the_unmasked_steps_in_order(seed, n, F, K):
    from cadm_b200.dynamics.training import flatten_future
    rng = np.random.default_rng(seed)
    D, A = 3, 2
    obs, nxt, delta, back = (rng.normal(size=(n, F * D)) for _ in range(4))
    act = rng.normal(size=(n, F * A))
    cp_obs, cp_act = rng.normal(size=(n, K * D)), rng.normal(size=(n, K * A))
    fb = (rng.uniform(size=(n, F)) < 0.6).astype(np.float64)
    o, a, d, on, bd, co, ca = flatten_future(D, A, K, F, obs, act, delta, cp_obs, cp_act, fb, nxt, back)
    rows = [(i, f) for i in range(n) for f in range(F) if fb[i, f] > 0]
    assert o.shape == (len(rows), D) and ca.shape == (len(rows), K * A)
    for r, (i, f) in enumerate(rows):
        assert np.array_equal(o[r], obs[i, f * D:(f + 1) * D]) and np.array_equal(a[r], act[i, f * A:(f + 1) * A])
        assert np.array_equal(on[r], nxt[i, f * D:(f + 1) * D]) and np.array_equal(bd[r], back[i, f * D:(f + 1) * D])
        assert np.array_equal(co[r], cp_obs[i]) and np.array_equal(ca[r], cp_act[i])


@settings(**SET)
@given(seed=st.integers(0, 2 ** 31), E=st.integers(1, 4), H=st.integers(1, 9), D=st.integers(1, 5), In=st.integers(1, 7), n_hidden=st.integers(1, 3),
       with_enc=st.booleans(), with_back=st.booleans())
def test_native_trainer_flat_vector_round_trips(seed, E, H, D, In, n_hidden, with_enc, with_back):
    """cadm_train_set_params / get_params move ONE flat vector; pack and unpack (pure NumPy) must be inverse to each other and
    put [output_mu | output_logvar] side by side per hidden unit (layout in include/cadm_b200.h)."""
    from cadm_b200.dynamics.native_trainer import NativeTrainer
    rng = np.random.default_rng(seed)

    def mlp():
        sizes = [In] + [H] * n_hidden
        return dict(W=[rng.normal(size=(E, sizes[i], sizes[i + 1])).astype(np.float32) for i in range(n_hidden)],
                    b=[rng.normal(size=(E, 1, H)).astype(np.float32) for _ in range(n_hidden)],
                    W_mu=rng.normal(size=(E, H, D)).astype(np.float32), b_mu=rng.normal(size=(E, 1, D)).astype(np.float32),
                    W_lv=rng.normal(size=(E, H, D)).astype(np.float32), b_lv=rng.normal(size=(E, 1, D)).astype(np.float32),
                    max_logvar=rng.normal(size=(1, D)).astype(np.float32), min_logvar=rng.normal(size=(1, D)).astype(np.float32))

    dyn, back = mlp(), (mlp() if with_back else None)
    enc = dict(W=[rng.normal(size=(E, 4, 3)).astype(np.float32), rng.normal(size=(E, 3, 2)).astype(np.float32)],
               b=[rng.normal(size=(E, 1, 3)).astype(np.float32), rng.normal(size=(E, 1, 2)).astype(np.float32)]) if with_enc else None
    tr = NativeTrainer.__new__(NativeTrainer)                       # the layout code only: no handle, no GPU
    flat = tr._pack(enc, dyn, back)
    n_mlp = sum(w.size for w in dyn["W"]) + sum(b.size for b in dyn["b"]) + 2 * E * H * D + 2 * E * D
    assert flat.dtype == np.float32 and flat.size == (sum(w.size + b.size for w, b in zip(enc["W"], enc["b"])) if with_enc else 0) + \
        n_mlp + 2 * D + (n_mlp if with_back else 0)
    zero = lambda d: None if d is None else {k: ([np.zeros_like(x) for x in v] if isinstance(v, list) else np.zeros_like(v)) for k, v in d.items()}
    enc2, dyn2, back2 = zero(enc), zero(dyn), zero(back)
    tr._unpack(flat, enc2, dyn2, back2)
    for a, b in ((enc, enc2), (dyn, dyn2), (back, back2)):
        if a is None:
            continue
        for key, v in a.items():
            if a is back and key in ("max_logvar", "min_logvar"):
                continue                          # built deterministic (:228): its bounds never enter the loss and are not trained
            for x, y in zip(v if isinstance(v, list) else [v], b[key] if isinstance(v, list) else [b[key]]):
                assert np.array_equal(x, y), key
    # heads: unit j of member e holds [W_mu[e, j, :], W_lv[e, j, :]] contiguously
    off = (sum(w.size + b.size for w, b in zip(enc["W"], enc["b"])) if with_enc else 0) + sum(w.size for w in dyn["W"]) + sum(b.size for b in dyn["b"])
    heads = flat[off:off + E * H * 2 * D].reshape(E, H, 2 * D)
    assert np.array_equal(heads[:, :, :D], dyn["W_mu"]) and np.array_equal(heads[:, :, D:], dyn["W_lv"])
