"""GPU parity at the sizes the throughput numbers are quoted on ("parity gap"):

* batches with MORE TILES THAN SMs, so that every rollout kernel runs its `for (tile = blockIdx.x; ...; tile += gridDim.x)`
  loop more than once per CTA (mbarrier phases, ring positions and accumulator parities carried across tiles):
  C2 with m = 10 environments = 8000 rows per member = 315 tiles of 128 rows (rollout_tc.cu), 625 tiles of 64 rows /
  1250 tiles of 32 rows (rollout_tcs.cu), and the FFMA kernel's tiles -- states, particle returns and whole CEM decisions;
* BASELINE.json configs[3] as a DECISION: Ant PE-TS + CaDM, n = 1000 candidates, full horizon;
* two cells of the configs[4] sweep that never met the oracle: p = 100 particles and n = 5000 candidates.

The fp64 oracle (cadm/dynamics/core/utils.py:137-182 restated in oracle/cadm_oracle.py) runs ONCE per case and every
kernel variant is compared with it; horizons are shortened where the oracle would take minutes, n / p / m are kept.
"""
import contextlib
import os

import numpy as np
import pytest
import torch

from oracle import cadm_oracle as orc

from helpers import elite_margin_ok, oracle_pack, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4
NUM_SMS = 148

# name -> (precision, CADM_TC_VARIANT, CADM_TCS_ROWS)
KERNELS = {
    "fp32": ("fp32", None, None),
    "tiles128": ("tc3x", "1", None),
    "swapped64": ("tc3x", "2", "64"),
    "swapped32": ("tc3x", "2", "32"),
    "pairs32": ("tc3x", "3", "32"),
    "pairs64": ("tc3x", "3", "64"),
    "auto": ("tc3x", None, None),            # the engine's choice: for these batches full waves of 128-row tiles + swapped small tiles
}


@contextlib.contextmanager
def kernel_env(name):
    prec, variant, rows = KERNELS[name]
    keys = {"CADM_TC_VARIANT": variant, "CADM_TCS_ROWS": rows}
    old = {k: os.environ.get(k) for k in keys}
    for k, v in keys.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    try:
        yield prec
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _build(config, kernel, m_max, horizon=None, **kw):
    """build_model with a shortened horizon (the synthetic configs fix h = 30)."""
    from cadm_b200 import synth
    with kernel_env(kernel) as prec:
        if horizon is None:
            return synth.build_model(config, m_max=m_max, precision=prec, **kw)
        saved = dict(synth.CONFIGS[config])
        synth.CONFIGS[config]["horizon"] = horizon
        try:
            return synth.build_model(config, m_max=m_max, precision=prec, **kw)
        finally:
            synth.CONFIGS[config].clear()
            synth.CONFIGS[config].update(saved)


def _tiles(cfg, m, n, rows):
    q = cfg["particles"] // cfg["ensemble"]
    return cfg["ensemble"] * -(-(q * m * n) // rows)


def _noise(seed, cfg, m, n, h, D, A):
    """Injected draws in the global layouts of PlannerEngine.plan_cem (any values are legal inputs; NumPy's generator is
    faster than the Philox specification at these sizes)."""
    rng = np.random.default_rng(seed)
    E, p = cfg["ensemble"], cfg["particles"]
    z = np.clip(rng.standard_normal((orc.NUM_CEM_ITERS, m, n, h, A)), -2.0, 2.0).astype(np.float32)
    eps = None
    if not cfg["deterministic"]:
        eps = rng.standard_normal((orc.NUM_CEM_ITERS, h, E, (p // E) * m * n, D)).astype(np.float32)
    return z, eps


def _check_decision(out, ref, m, tag):
    rets, el = out["returns"].cpu().numpy(), out["elites"].cpu().numpy()
    assert np.isfinite(rets).all(), tag
    for it in range(orc.NUM_CEM_ITERS):
        err = np.max(np.abs(rets[it] - ref.returns[it])) / np.max(np.abs(ref.returns[it]))
        assert err < TOL, (tag, it, err)
        ok, gap, e = elite_margin_ok(ref.returns[it], ref.elites[it], rets[it], orc.NUM_ELITES)
        if ok:
            assert np.array_equal(el[it], ref.elites[it]), (tag, it)
        else:   # the same SET up to boundary ties (margin rule)
            assert len(set(map(int, el[it].ravel())) ^ set(map(int, ref.elites[it].ravel()))) <= 2 * m, (tag, it, gap, e)
    assert np.max(np.abs(out["mean"].cpu().numpy() - ref.mean)) < TOL, tag
    assert np.max(np.abs(out["var"].cpu().numpy() - ref.var)) < TOL, tag


def _oracle_decision(model, env, cfg, inp, z, eps):
    prm, enc, norm, oenv = oracle_pack(model)
    f64 = lambda a: None if a is None else a.astype(np.float64)
    ctx_raw = orc.encode_context(f64(inp["cp_obs"]), f64(inp["cp_act"]), enc, norm) if cfg["context"] else None
    return orc.cem_plan(f64(inp["obs"]), f64(inp["init_mean"]), f64(inp["init_var"]), f64(z), prm, norm, oenv, cfg["ensemble"],
                        cfg["particles"], cfg["deterministic"], f64(eps), ctx_raw)


# ---------------------------------------------------------------------------------------------------
def test_multi_tile_rollout_states_and_returns():
    """C2, m = 10, n = 200: every intermediate state and the particle returns of a 4-step rollout, on every kernel, with
    2.1 .. 8.4 tiles per CTA."""
    from cadm_b200.synth import synthetic_inputs
    m, n, h = 10, 200, 4
    ref_ret = ref_st = None
    for kernel in KERNELS:
        model, env, cfg = _build("C2", kernel, m_max=m, horizon=h)
        rows = {"tiles128": 128, "swapped64": 64, "swapped32": 32}.get(kernel)
        if rows:
            assert _tiles(cfg, m, n, rows) > 2 * NUM_SMS
        inp = synthetic_inputs(env, m, h, False, seed=21)
        rng = np.random.default_rng(22)
        E, p, D, A = cfg["ensemble"], cfg["particles"], env.obs_dim, env.act_dim
        actions = rng.uniform(-1, 1, (m, n, h, A)).astype(np.float32)
        eps = rng.standard_normal((h, E, (p // E) * m * n, D)).astype(np.float32)
        if ref_ret is None:
            prm, enc, norm, oenv = oracle_pack(model)
            ref_ret, ref_st = orc.rollout(inp["obs"].astype(np.float64), actions.astype(np.float64), prm, norm, oenv, E, p, False,
                                          eps.astype(np.float64), None, trace=True)
        pr, st = model.engine.rollout(inp["obs"], actions, None, eps, it=0, trace=True)
        pr, st = pr.cpu().numpy(), st.cpu().numpy()
        if kernel == "auto":                              # 40 000 rows: two launches over disjoint rows of every member
            assert "+" in model.engine.kernel_name, model.engine.kernel_name
        assert np.isfinite(st).all(), kernel
        assert rel_err(st, ref_st, axis=(1, 2, 3)) < TOL, kernel
        assert np.max(np.abs(pr - ref_ret)) / np.max(np.abs(ref_ret)) < TOL, kernel
        model.engine.close()


def test_multi_tile_cem_decision():
    """C2, m = 10 (the reference plans for 10-20 environments per call, cadm/samplers/sampler.py:107-120): a whole
    5-iteration decision with injected noise on every kernel -- returns, elite indices, final mean / variance."""
    from cadm_b200.synth import synthetic_inputs
    m, h = 10, 5
    ref = None
    for kernel in KERNELS:
        model, env, cfg = _build("C2", kernel, m_max=m, horizon=h)
        inp = synthetic_inputs(env, m, h, False, seed=23)
        z, eps = _noise(24, cfg, m, cfg["candidates"], h, env.obs_dim, env.act_dim)
        if ref is None:
            ref = _oracle_decision(model, env, cfg, inp, z, eps)
        out = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=0, z=z, eps=eps)
        _check_decision(out, ref, m, kernel)
        # the seed-only decision on the same engine: deterministic replay across the multi-tile loop
        a = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=77)
        b = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=77)
        assert torch.equal(a["returns"], b["returns"]) and torch.equal(a["elites"], b["elites"]), kernel
        model.engine.close()


def test_c4_full_decision():
    """BASELINE.json configs[3]: Ant PE-TS + CaDM context encoder, n = 1000, p = 20, E = 5, h = 30, m = 1 -- the whole
    decision (context encoder, odd-iteration context pairing, 5 x 30 steps, top-50 of 1000, refit) on every kernel."""
    from cadm_b200.synth import synthetic_inputs
    ref = None
    for kernel in ("tiles128", "swapped64", "pairs64", "auto", "fp32"):
        model, env, cfg = _build("C4", kernel, m_max=1)
        assert cfg["candidates"] == 1000 and cfg["horizon"] == 30
        if kernel in ("tiles128", "swapped64"):
            assert _tiles(cfg, 1, 1000, 128 if kernel == "tiles128" else 64) > NUM_SMS
        inp = synthetic_inputs(env, 1, 30, True, seed=25)
        z, eps = _noise(26, cfg, 1, 1000, 30, env.obs_dim, env.act_dim)
        if ref is None:
            ref = _oracle_decision(model, env, cfg, inp, z, eps)
        out = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], inp["cp_obs"], inp["cp_act"], seed=0, z=z, eps=eps)
        _check_decision(out, ref, 1, kernel)
        model.engine.close()


@pytest.mark.parametrize("n,p,h", [(200, 100, 6), (5000, 20, 3)])
def test_c5_sweep_cells(n, p, h):
    """BASELINE.json configs[4]: the sweep's p = 100 and n = 5000 corners as decisions (short horizon, n / p kept)."""
    from cadm_b200.synth import synthetic_inputs
    ref = None
    for kernel in ("tiles128", "swapped64", "auto"):
        model, env, cfg = _build("C2", kernel, m_max=1, horizon=h, candidates=n, particles=p)
        assert _tiles(cfg, 1, n, 128 if kernel == "tiles128" else 64) > NUM_SMS
        inp = synthetic_inputs(env, 1, h, False, seed=27)
        z, eps = _noise(28, cfg, 1, n, h, env.obs_dim, env.act_dim)
        if ref is None:
            ref = _oracle_decision(model, env, cfg, inp, z, eps)
        out = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=0, z=z, eps=eps)
        _check_decision(out, ref, 1, kernel)
        model.engine.close()
