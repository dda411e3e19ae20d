"""ref_cpu (the timed CPU baseline) is the verified thing: cross-check against the NumPy oracle."""
import numpy as np
import pytest

from oracle import cadm_oracle as orc
from oracle import philox as ph
from oracle.envs import get_env
from oracle.ref_cpu import RefCpuPlanner


@pytest.mark.parametrize("envname,ctx,m", [("halfcheetah", False, 1), ("halfcheetah", True, 2), ("ant", True, 2)])
def test_ref_cpu_matches_numpy_oracle(envname, ctx, m):
    env = get_env(envname)
    rng = np.random.default_rng(0)
    E, p, n, h, H, C, K = 5, 10, 60, 6, 64, 4, 3
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    prm = orc.init_dynamics_params(rng, E, P + A + (C if ctx else 0), H, D)
    prm.b_lv[...] = -5.0
    enc = orc.init_encoder_params(rng, E, (D + A) * K, (32, 16, 8), C) if ctx else None
    norm = orc.NormStats(np.zeros(P), np.ones(P), np.zeros(A), np.full(A, 0.6), np.zeros(D), np.full(D, 0.1),
                         np.zeros(D * K), np.ones(D * K), np.zeros(A * K), np.ones(A * K)).astype(np.float32)
    obs = (rng.standard_normal((m, D)) * 0.1).astype(np.float32)
    cp_obs = (rng.standard_normal((m, D * K)) * 0.1).astype(np.float32)
    cp_act = (rng.standard_normal((m, A * K)) * 0.1).astype(np.float32)
    z = ph.gen_z(1, 5, m, n, h, A)
    eps = ph.gen_eps(1, 5, h, m, n, p, E, D)
    mean0, var0 = np.zeros((m, h, A), np.float32), np.full((m, h, A), 0.25, np.float32)
    ctx_raw = orc.encode_context(cp_obs, cp_act, enc, norm) if ctx else None
    ref = orc.cem_plan(obs, mean0, var0, z, prm, norm, env, E, p, False, eps, ctx_raw)
    pl = RefCpuPlanner(prm, norm, envname, E, p, False, enc)
    out = pl.cem(obs, mean0, var0, n, cp_obs if ctx else None, cp_act if ctx else None, z=z, eps=eps)
    np.testing.assert_allclose(out["returns"], ref.returns, rtol=2e-4, atol=2e-5)
    assert np.array_equal(out["elites"], ref.elites)
    np.testing.assert_allclose(out["mean"], ref.mean, atol=2e-5)
    np.testing.assert_allclose(out["var"], ref.var, atol=2e-5)
    # unseeded mode (own RNG, like the reference) runs and respects the bounds
    out2 = pl.cem(obs, mean0, var0, n, cp_obs if ctx else None, cp_act if ctx else None)
    assert np.abs(out2["action"]).max() <= 1.0 and np.isfinite(out2["returns"]).all()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs without a GPU and prints ONE JSON line with the keys the driver reads."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root,
                         env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "actions/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
