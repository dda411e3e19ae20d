"""Live check (build container only): the reference's OWN planner graph code against the oracle on a seeded sweep of cases.

The recordings under tests/golden/recorded/ pin ten hand-picked cases.  This script draws further cases -- environment,
planner (CEM / random shooting / discrete), ensemble size, particles, candidates, horizon, number of environments, context,
deterministic -- from a seed, runs each through the UNMODIFIED builders of /root/reference over the NumPy TensorFlow
stand-in (exactly as make_reference_golden.py does) and through oracle/cadm_oracle.py with the same weights and noise, and
prints one JSON line per case with the differences.  tests/test_reference_live.py runs it in a subprocess (the stand-in
registers a fake `tensorflow` module, which should not leak into the test process) and skips where /root/reference does
not exist (the GPU box).

usage: python tests/golden/live_reference_check.py SEED COUNT
"""
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, "tests")]
import make_reference_golden as G                       # noqa: E402  (run_fixture: the wiring of the reference builders)
import tf_numpy_shim as shim                            # noqa: E402
from reference_cases import make_case                   # noqa: E402
from oracle import cadm_oracle as orc                   # noqa: E402
from oracle import philox as ph                         # noqa: E402
from oracle.envs import get_env                         # noqa: E402

ENVS = ["halfcheetah", "cripple_halfcheetah", "ant", "slim_humanoid", "pendulum", "cartpole"]


def draw_case(rng):
    envname = ENVS[int(rng.integers(len(ENVS)))]
    discrete = envname == "cartpole"                    # the only discrete action space (classic_control.py:94-101)
    mode = "rs_discrete" if discrete else ("cem" if rng.random() < 0.65 else "rs")
    det = bool(rng.random() < 0.25)
    E = 1 if det and rng.random() < 0.5 else int(rng.choice([1, 2, 3, 5]))
    p = E * int(rng.integers(1, 4))
    context = bool(rng.random() < 0.4)
    return dict(envname=envname, mode=mode, E=E, p=p, n=int(rng.choice([50, 56, 64, 80])), h=int(rng.integers(2, 6)),
                H=int(rng.choice([16, 24, 32])), m=int(rng.integers(1, 5)), context=context, det=det,
                seed=int(rng.integers(1, 1 << 20)))


def oracle_run(g):
    E, p, n, h, H, m, context, det, seed, C, K = [int(v) for v in g["meta"]]
    f8 = np.float64
    mode = str(g["mode"])
    env = get_env(str(g["envname"]))
    prm = orc.DynamicsParams([g[f"W{i}"] for i in range(4)], [g[f"b{i}"] for i in range(4)], g["W_mu"], g["b_mu"], g["W_lv"],
                             g["b_lv"], g["max_logvar"], g["min_logvar"]).astype(f8)
    norm = orc.NormStats(*[g[f"norm_{k}"] for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
                                                     "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std")]).astype(f8)
    ctx_raw = None
    if context:
        enc = orc.EncoderParams([g[f"encW{i}"] for i in range(4)], [g[f"encb{i}"] for i in range(4)]).astype(f8)
        ctx_raw = orc.encode_context(g["cp_obs"].astype(f8), g["cp_act"].astype(f8), enc, norm)
    iters = orc.NUM_CEM_ITERS if mode == "cem" else 1
    eps = None if det else ph.gen_eps(seed, iters, h, m, n, p, E, env.obs_dim).astype(f8)
    if mode == "cem":
        z = ph.gen_z(seed, iters, m, n, h, env.act_dim).astype(f8)
        r = orc.cem_plan(g["obs"].astype(f8), g["mean0"].astype(f8), g["var0"].astype(f8), z, prm, norm, env, E, p, bool(det),
                         eps, ctx_raw)
        return dict(returns=r.returns, index=r.elites, plan=r.mean)
    if mode == "rs_discrete":
        u = ph.gen_discrete_actions(seed, m, n, h, env.act_dim)
    else:
        u = ph.gen_uniform_actions(seed, m, n, h, env.act_dim).astype(f8)
    r = orc.rs_plan(g["obs"].astype(f8), u, prm, norm, env, E, p, bool(det), None if eps is None else eps[0], ctx_raw,
                    discrete=mode == "rs_discrete")
    return dict(returns=r["returns"], index=r["best"], plan=np.asarray(r["action"], f8))


def main():
    seed, count = int(sys.argv[1]), int(sys.argv[2])
    tf = shim.install(np.float64)
    sys.path.insert(0, "/root/reference")
    from cadm.dynamics.core import utils as U
    rng = np.random.default_rng(seed)
    for _ in range(count):
        spec = draw_case(rng)
        g = make_case(**spec)
        ref = G.run_fixture(g, U, tf)
        got = oracle_run(g)
        ref_index = ref["elites"] if spec["mode"] == "cem" else ref["best"]
        scale = float(np.max(np.abs(ref["returns"])))
        print(json.dumps(dict(spec=spec, returns_rel=float(np.max(np.abs(got["returns"] - ref["returns"]))) / max(scale, 1e-300),
                              plan_abs=float(np.max(np.abs(got["plan"] - np.asarray(ref["plan"], np.float64).reshape(got["plan"].shape)))),
                              index_equal=bool(np.array_equal(got["index"], ref_index)),
                              spread=float(np.ptp(ref["returns"])))), flush=True)


if __name__ == "__main__":
    main()
