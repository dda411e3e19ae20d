"""GPU parity: the CUDA engine (through the C ABI) against the oracle on identical inputs.

Bars (BASELINE.json north_star): next states and candidate returns within 1e-4 relative in fp32; elite indices
bit-exact.  "Relative" is measured against the RMS of the fp64 oracle's tensor (per state dimension for states).
"""
import numpy as np
import pytest
import torch

from oracle import cadm_oracle as orc
from oracle import philox as ph

from helpers import elite_margin_ok, oracle_pack, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _model(config, m_max=4, **kw):
    from cadm_b200.synth import build_model
    return build_model(config, m_max=m_max, **kw)


@pytest.fixture(scope="module", params=["fp32", "tc3x-tiles", "tc3x-swapped", "tc3x-pairs"])
def precision(request):
    """fp32 FFMA path and the tensor-core kernels: 128-row tiles (rollout_tc.cu), swapped operands (rollout_tcs.cu) and the
    CTA-pair variant of the latter (rollout_tcp.cu);
    the variant is forced through the engine's CADM_TC_VARIANT knob (cadm_set_option "tc_variant")."""
    import os
    name, _, variant = request.param.partition("-")
    old = os.environ.get("CADM_TC_VARIANT")
    if variant:
        os.environ["CADM_TC_VARIANT"] = {"tiles": "1", "swapped": "2", "pairs": "3"}[variant]
    yield name
    if old is None:
        os.environ.pop("CADM_TC_VARIANT", None)
    else:
        os.environ["CADM_TC_VARIANT"] = old


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("config", ["C1", "C2", "C3", "C4"])
def test_predict_one_step(config, precision):
    """predict(): mu / logvar / sampled next state of ONE model evaluation, training layout [E, B, .]."""
    model, env, cfg = _model(config, candidates=64, precision=precision)
    prm, enc, norm, oenv = oracle_pack(model)
    rng = np.random.default_rng(0)
    E, B, D, A = cfg["ensemble"], 77, env.obs_dim, env.act_dim
    obs = rng.standard_normal((E, B, D)) * 0.5
    act = rng.uniform(-1, 1, (E, B, A))
    eps = rng.standard_normal((E, B, D))
    ctx = rng.standard_normal((E, B, 10)) * 0.3 if cfg["context"] else None
    f = np.float32
    args = (obs.astype(f), act.astype(f)) + ((None, None, ctx.astype(f)) if cfg["context"] else ())
    nxt, mu, lv = model.predict(*args, eps=eps.astype(f))
    o_nxt, o_mu, o_lv = orc.predict(obs.astype(f).astype(np.float64), act.astype(f).astype(np.float64), prm, norm, oenv,
                                    cfg["deterministic"], eps.astype(f).astype(np.float64),
                                    None if ctx is None else ctx.astype(f).astype(np.float64))
    assert rel_err(mu, o_mu) < TOL
    assert rel_err(lv, o_lv) < TOL
    assert rel_err(nxt, o_nxt, axis=(0, 1)) < TOL


@pytest.mark.parametrize("config,m,n", [("C1", 1, 200), ("C2", 1, 40), ("C2", 3, 16), ("C3", 3, 16), ("C4", 2, 24)])
def test_rollout_states_and_returns(config, m, n, precision):
    """30-step rollouts with injected noise: every intermediate state and the particle returns."""
    model, env, cfg = _model(config, candidates=n, precision=precision, num_elites=min(50, n))
    prm, enc, norm, oenv = oracle_pack(model)
    from cadm_b200.synth import synthetic_inputs
    inp = synthetic_inputs(env, m, cfg["horizon"], cfg["context"], seed=1)
    E, p, h, D, A = cfg["ensemble"], cfg["particles"], cfg["horizon"], env.obs_dim, env.act_dim
    rng = np.random.default_rng(2)
    actions = rng.uniform(-1, 1, (m, n, h, A)).astype(np.float32)
    eps = None if cfg["deterministic"] else ph.gen_eps(5, 1, h, m, n, p, E, D)[0]
    ctx_raw = None
    for it in (0, 1):                                     # odd iteration exercises quirk Q3
        ctx_rows = None
        if cfg["context"]:
            ctx_raw = orc.encode_context(inp["cp_obs"].astype(np.float64), inp["cp_act"].astype(np.float64), enc, norm)
            carried = ctx_raw
            for k in range(it + 1):
                ctx_rows, carried = orc._context_rows(carried, k, m, n, p, E, ctx_raw.shape[-1])
        o_ret, o_st = orc.rollout(inp["obs"].astype(np.float64), actions.astype(np.float64), prm, norm, oenv, E, p,
                                  cfg["deterministic"], None if eps is None else eps.astype(np.float64), ctx_rows, trace=True)
        g_ctx = model.engine.encode_context(inp["cp_obs"], inp["cp_act"]) if cfg["context"] else None
        if cfg["context"]:
            assert rel_err(g_ctx.cpu().numpy(), ctx_raw) < TOL
        pr, st = model.engine.rollout(inp["obs"], actions, g_ctx, eps, it=it, trace=True)
        pr, st = pr.cpu().numpy(), st.cpu().numpy()
        assert np.isfinite(st).all()
        e_state = rel_err(st, o_st, axis=(1, 2, 3))
        e_ret = np.max(np.abs(pr - o_ret)) / np.max(np.abs(o_ret))
        assert e_state < TOL, (it, e_state)
        assert e_ret < TOL, (it, e_ret)
        if not cfg["context"]:
            break


@pytest.mark.parametrize("config,m", [("C1", 1), ("C2", 1), ("C2", 2), ("C3", 2)])
def test_cem_decision_injected_noise(config, m, precision):
    """A whole CEM decision (5 iterations) with injected z / eps: returns, elite indices, final mean / var."""
    model, env, cfg = _model(config, precision=precision)
    prm, enc, norm, oenv = oracle_pack(model)
    from cadm_b200.synth import synthetic_inputs
    inp = synthetic_inputs(env, m, cfg["horizon"], cfg["context"], seed=3)
    E, p, n, h, D, A = cfg["ensemble"], cfg["particles"], cfg["candidates"], cfg["horizon"], env.obs_dim, env.act_dim
    z = ph.gen_z(7, orc.NUM_CEM_ITERS, m, n, h, A)
    eps = None if cfg["deterministic"] else ph.gen_eps(7, orc.NUM_CEM_ITERS, h, m, n, p, E, D)
    ctx_raw = orc.encode_context(inp["cp_obs"].astype(np.float64), inp["cp_act"].astype(np.float64), enc, norm) \
        if cfg["context"] else None
    ref = orc.cem_plan(inp["obs"].astype(np.float64), inp["init_mean"].astype(np.float64), inp["init_var"].astype(np.float64),
                       z.astype(np.float64), prm, norm, oenv, E, p, cfg["deterministic"],
                       None if eps is None else eps.astype(np.float64), ctx_raw)
    out = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], inp.get("cp_obs"), inp.get("cp_act"),
                                seed=0, z=z, eps=eps)
    rets, el = out["returns"].cpu().numpy(), out["elites"].cpu().numpy()
    mean, var = out["mean"].cpu().numpy(), out["var"].cpu().numpy()
    for it in range(orc.NUM_CEM_ITERS):
        err = np.max(np.abs(rets[it] - ref.returns[it])) / np.max(np.abs(ref.returns[it]))
        assert err < TOL, (it, err)
        ok, gap, e = elite_margin_ok(ref.returns[it], ref.elites[it], rets[it], orc.NUM_ELITES)
        if ok:
            assert np.array_equal(el[it], ref.elites[it]), it
        else:   # margin rule: the same SET must still be selected up to boundary ties
            assert len(set(el[it].ravel()) ^ set(ref.elites[it].ravel())) <= 2 * m, (it, gap, e)
    assert np.max(np.abs(mean - ref.mean)) < TOL
    assert np.max(np.abs(var - ref.var)) < TOL


def test_cem_decision_philox_seed(precision):
    """Seed-only mode: the engine's own Philox noise equals the NumPy specification (oracle/philox.py)."""
    model, env, cfg = _model("C2", precision=precision)
    prm, enc, norm, oenv = oracle_pack(model)
    from cadm_b200.synth import synthetic_inputs
    m = 1
    inp = synthetic_inputs(env, m, cfg["horizon"], False, seed=4)
    E, p, n, h, D, A = cfg["ensemble"], cfg["particles"], cfg["candidates"], cfg["horizon"], env.obs_dim, env.act_dim
    seed = (12345 << 32) | 17
    z = ph.gen_z(seed, orc.NUM_CEM_ITERS, m, n, h, A, dtype=np.float64)
    eps = ph.gen_eps(seed, orc.NUM_CEM_ITERS, h, m, n, p, E, D, dtype=np.float64)
    ref = orc.cem_plan(inp["obs"].astype(np.float64), inp["init_mean"].astype(np.float64), inp["init_var"].astype(np.float64),
                       z, prm, norm, oenv, E, p, False, eps)
    out = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=seed)
    rets, el = out["returns"].cpu().numpy(), out["elites"].cpu().numpy()
    for it in range(orc.NUM_CEM_ITERS):
        err = np.max(np.abs(rets[it] - ref.returns[it])) / np.max(np.abs(ref.returns[it]))
        assert err < TOL, (it, err)
        ok, gap, e = elite_margin_ok(ref.returns[it], ref.elites[it], rets[it], orc.NUM_ELITES)
        if ok:
            assert np.array_equal(el[it], ref.elites[it]), it
    assert np.max(np.abs(out["mean"].cpu().numpy() - ref.mean)) < TOL


def test_kat_zero_weights_on_gpu(precision):
    """KAT 1 + 4 on the device: zero weights -> closed-form returns; elites = smallest action energy."""
    model, env, cfg = _model("C1", precision=precision)
    for k in ("W_mu", "W_lv"):
        model._dyn[k][...] = 0
    for w in model._dyn["W"]:
        w[...] = 0
    model._push_params()
    from cadm_b200.synth import synthetic_inputs
    inp = synthetic_inputs(env, 2, 30, False, seed=5)
    n, h, A = 200, 30, env.act_dim
    z = ph.gen_z(1, 5, 2, n, h, A)
    out = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], z=z)
    a0 = 0.5 * z[0].astype(np.float64)
    energy = np.sum(a0 ** 2, axis=(2, 3))
    expect = inp["obs"][:, :1].astype(np.float64) + 0.0 - 0.1 * energy     # mu_delta = 0
    rets = out["returns"].cpu().numpy()
    np.testing.assert_allclose(rets[0], expect, rtol=0, atol=2e-5)
    want = np.argsort(energy, axis=1, kind="stable")[:, :50]
    assert np.array_equal(out["elites"].cpu().numpy()[0], want)


def test_ties_pick_lower_index(precision):
    """All candidates identical -> elites 0..49 (tf.nn.top_k tie rule)."""
    model, env, cfg = _model("C1", precision=precision)
    from cadm_b200.synth import synthetic_inputs
    inp = synthetic_inputs(env, 1, 30, False, seed=6)
    z = np.zeros((5, 1, 200, 30, env.act_dim), np.float32)
    out = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], z=z)
    assert np.array_equal(out["elites"].cpu().numpy()[0, 0], np.arange(50))


@pytest.mark.parametrize("G", [2, 4, 8])
def test_virtual_rank_sharding_bit_identical(G, precision):
    """Candidate sharding over G virtual ranks on one device gives the SAME bits as a single rank
    (returns, elites, plan); the all-gather is emulated by copying slices between the engines' buffers."""
    from cadm_b200.synth import build_model, synthetic_inputs
    single, env, cfg = build_model("C2", m_max=2, precision=precision)
    m = 2
    inp = synthetic_inputs(env, m, cfg["horizon"], False, seed=8)
    seed = 99
    ref = single.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=seed)
    ranks = [build_model("C2", m_max=2, precision=precision, rank=r, world=G)[0] for r in range(G)]
    for r in ranks:
        r.engine.cem_begin(inp["obs"], inp["init_mean"], inp["init_var"])
    for it in range(5):
        for r in ranks:
            r.engine.cem_rollout(it, seed=seed)
        bufs = [r.engine.returns_buffer() for r in ranks]
        for i, bi in enumerate(bufs):
            for j, bj in enumerate(bufs):
                if i != j:
                    bi[j].copy_(bj[j])
        for r in ranks:
            r.engine.cem_refit(it)
    outs = [r.engine.cem_finish() for r in ranks]
    for o in outs:
        assert torch.equal(o["returns"], ref["returns"])
        assert torch.equal(o["elites"], ref["elites"])
        assert torch.equal(o["mean"], ref["mean"])
        assert torch.equal(o["var"], ref["var"])


def test_random_shooting_parity(precision):
    from cadm_b200.synth import synthetic_inputs
    from cadm_b200.dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel
    from cadm_b200.synth import synthetic_normalization
    from cadm_b200.envs import make_env
    env = make_env("halfcheetah")
    model = MLPEnsembleCEMDynamicsModel("dm", env, hidden_nonlinearity="swish", n_forwards=10, n_candidates=128,
                                        ensemble_size=5, n_particles=10, use_cem=False, m_max=2, precision=precision)
    model._dyn["b_lv"][...] = -6.0
    model._push_params()
    model.set_normalization(synthetic_normalization(env, False))
    prm, enc, norm, oenv = oracle_pack(model)
    inp = synthetic_inputs(env, 2, 10, False, seed=9)
    u = ph.gen_uniform_actions(3, 2, 128, 10, env.act_dim)
    eps = ph.gen_eps(3, 1, 10, 2, 128, 10, 5, env.obs_dim)[0]
    ref = orc.rs_plan(inp["obs"].astype(np.float64), u.astype(np.float64), prm, norm, oenv, 5, 10, False, eps.astype(np.float64))
    out = model.engine.plan_rs(inp["obs"], u=u, eps=eps)
    assert np.max(np.abs(out["returns"].cpu().numpy() - ref["returns"])) / np.max(np.abs(ref["returns"])) < TOL
    assert np.array_equal(out["best"].cpu().numpy(), ref["best"])
    np.testing.assert_allclose(out["action"].cpu().numpy(), ref["action"], atol=1e-6)
    # seed-only path agrees with the Philox specification
    out2 = model.engine.plan_rs(inp["obs"], seed=3, eps=eps)
    assert np.array_equal(out2["best"].cpu().numpy(), ref["best"])


def test_size_independent_properties_full_size(precision):
    """BASELINE-size C2 (n=200, p=20, h=30), m=4: properties that need no oracle run --
    plan within [-1, 1], elites sorted by return, distinct and valid, deterministic replay, returns finite."""
    model, env, cfg = _model("C2", m_max=4, precision=precision)
    from cadm_b200.synth import synthetic_inputs
    inp = synthetic_inputs(env, 4, 30, False, seed=10)
    a = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=5)
    b = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=5)
    c = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=6)
    assert torch.equal(a["mean"], b["mean"]) and torch.equal(a["elites"], b["elites"])
    assert not torch.equal(a["mean"], c["mean"])
    rets, el = a["returns"].cpu().numpy(), a["elites"].cpu().numpy()
    assert np.isfinite(rets).all()
    for it in range(5):
        for mi in range(4):
            idx = el[it, mi]
            assert len(set(idx)) == 50 and idx.min() >= 0 and idx.max() < 200
            r = rets[it, mi, idx]
            assert np.all(r[:-1] >= r[1:])
            assert r[-1] >= np.sort(rets[it, mi])[::-1][49] - 0.0
    act = model.get_action(inp["obs"], inp["init_mean"], inp["init_var"])
    assert act.shape == (4, 30, 6) and act.dtype == np.float32 and np.abs(act).max() <= 1.0


def test_many_candidates_sort_path_and_sharding():
    """n = 1600 candidates (8 x 200, the 8-GPU weak-scaling size): the elite selection takes the bitonic-sort path (n above the
    CTA size) and must still return the oracle's indices; eight virtual ranks of 200 give the single-rank bits."""
    from cadm_b200.dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel
    from cadm_b200.envs import make_env
    from cadm_b200.synth import synthetic_inputs, synthetic_normalization
    import os
    old = os.environ.get("CADM_TC_VARIANT")
    os.environ["CADM_TC_VARIANT"] = "2"
    try:
        env = make_env("halfcheetah")
        n, h, p, E, m = 1600, 4, 5, 5, 1

        def mk(rank=0, world=1):
            mdl = MLPEnsembleCEMDynamicsModel("dm", env, hidden_nonlinearity="swish", n_forwards=h, n_candidates=n, ensemble_size=E,
                                              n_particles=p, use_cem=True, m_max=m, precision="tc3x", rank=rank, world=world, seed=3)
            mdl._dyn["b_lv"][...] = -6.0
            mdl._push_params()
            mdl.set_normalization(synthetic_normalization(env, False))
            return mdl
        single = mk()
        prm, enc, norm, oenv = oracle_pack(single)
        inp = synthetic_inputs(env, m, h, False, seed=11)
        D, A = env.obs_dim, env.act_dim
        z = ph.gen_z(5, orc.NUM_CEM_ITERS, m, n, h, A)
        eps = ph.gen_eps(5, orc.NUM_CEM_ITERS, h, m, n, p, E, D)
        ref = orc.cem_plan(inp["obs"].astype(np.float64), inp["init_mean"].astype(np.float64), inp["init_var"].astype(np.float64),
                           z.astype(np.float64), prm, norm, oenv, E, p, False, eps.astype(np.float64))
        out = single.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=0, z=z, eps=eps)
        rets, el = out["returns"].cpu().numpy(), out["elites"].cpu().numpy()
        for it in range(orc.NUM_CEM_ITERS):
            assert np.max(np.abs(rets[it] - ref.returns[it])) / np.max(np.abs(ref.returns[it])) < TOL
            ok, gap, e = elite_margin_ok(ref.returns[it], ref.elites[it], rets[it], orc.NUM_ELITES)
            if ok:
                assert np.array_equal(el[it], ref.elites[it]), it
        assert np.max(np.abs(out["mean"].cpu().numpy() - ref.mean)) < TOL
        # eight virtual ranks (seed-only noise: every rank regenerates the elites it does not own)
        G = 8
        ref2 = single.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=21)
        ranks = [mk(r, G) for r in range(G)]
        for r in ranks:
            r.engine.cem_begin(inp["obs"], inp["init_mean"], inp["init_var"])
        for it in range(orc.NUM_CEM_ITERS):
            for r in ranks:
                r.engine.cem_rollout(it, seed=21)
            bufs = [r.engine.returns_buffer() for r in ranks]
            for i, bi in enumerate(bufs):
                for j, bj in enumerate(bufs):
                    if i != j:
                        bi[j].copy_(bj[j])
            for r in ranks:
                r.engine.cem_refit(it)
        for o in [r.engine.cem_finish() for r in ranks]:
            assert torch.equal(o["elites"], ref2["elites"]) and torch.equal(o["mean"], ref2["mean"]) and torch.equal(o["returns"], ref2["returns"])
    finally:
        if old is None:
            os.environ.pop("CADM_TC_VARIANT", None)
        else:
            os.environ["CADM_TC_VARIANT"] = old


def test_env_sharding_virtual_ranks(precision):
    """SURVEY 8(e), the alternative for m >= G: blocks of environments planned independently (EnvShardedPlanner, two
    virtual ranks on one device, ragged 2 + 1, injected z / eps in the global layouts) give the rows of the unsharded
    decision.  The rollout kernel may differ with the local batch size, so the comparison is to fp32 rounding (TOL), with
    the elite margin rule of the other CEM tests."""
    from cadm_b200.parallel import EnvShardedPlanner
    from cadm_b200.synth import synthetic_inputs
    model, env, cfg = _model("C2", precision=precision, candidates=64)
    m = 3
    inp = synthetic_inputs(env, m, cfg["horizon"], False, seed=4)
    E, p, n, h, D, A = cfg["ensemble"], cfg["particles"], 64, cfg["horizon"], env.obs_dim, env.act_dim
    z = ph.gen_z(9, orc.NUM_CEM_ITERS, m, n, h, A)
    eps = ph.gen_eps(9, orc.NUM_CEM_ITERS, h, m, n, p, E, D)
    full = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=0, z=z, eps=eps)
    full = {k: v.cpu().numpy() for k, v in full.items()}
    for rank in range(2):
        planner = EnvShardedPlanner(model.engine, rank=rank, world=2, gather=False)
        out = planner.plan(inp["obs"], inp["init_mean"], inp["init_var"], seed=0, z=z, eps=eps, logs=True)
        assert out["bounds"] == [0, 2, 3] and planner.collectives == 0
        lo, hi = out["bounds"][rank], out["bounds"][rank + 1]
        rets, el = out["returns"].cpu().numpy(), out["elites"].cpu().numpy()
        same = True
        for it in range(orc.NUM_CEM_ITERS):
            want = full["returns"][it, lo:hi]
            assert np.max(np.abs(rets[it] - want)) / np.max(np.abs(want)) < TOL, (rank, it)
            if not np.array_equal(el[it], full["elites"][it, lo:hi]):
                ok, gap, e = elite_margin_ok(want.astype(np.float64), None, rets[it], orc.NUM_ELITES)
                assert not ok, (rank, it, gap, e)          # only a boundary tie may change the selection
                same = False
                break
        if same:
            assert np.max(np.abs(out["mean"].cpu().numpy() - full["mean"][lo:hi])) < TOL
            assert np.max(np.abs(out["var"].cpu().numpy() - full["var"][lo:hi])) < TOL
