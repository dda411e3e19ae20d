"""Debug aid for the CTA-pair kernel: one-step predictions of the pair kernel against the single-CTA swapped kernel, error per
row block and per ensemble member."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from cadm_b200.synth import build_model
config = sys.argv[1] if len(sys.argv) > 1 else "C2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
outs = {}
for variant in ("2", "3"):
    os.environ["CADM_TC_VARIANT"] = variant
    model, env, cfg = build_model(config, m_max=1, candidates=64, precision="tc3x")
    rng = np.random.default_rng(0)
    E, D, A = cfg["ensemble"], env.obs_dim, env.act_dim
    obs = (rng.standard_normal((E, B, D)) * 0.5).astype(np.float32)
    act = rng.uniform(-1, 1, (E, B, A)).astype(np.float32)
    ctx = (rng.standard_normal((E, B, 10)) * 0.3).astype(np.float32) if cfg["context"] else None
    args = (obs, act) + ((None, None, ctx) if cfg["context"] else ())
    nxt, mu, lv = model.predict(*args, eps=np.zeros((E, B, D), np.float32))
    outs[variant] = (nxt, mu, lv)
    print(variant, model.engine.kernel_name)
    model.engine.close()
a, b = outs["2"], outs["3"]
for name, x, y in zip(("next", "mu", "lv"), a, b):
    err = np.abs(x - y)
    print(name, "max abs diff", err.max(), "scale", np.abs(x).max(), "identical" if np.array_equal(x, y) else "")
    for e in range(x.shape[0]):
        print("  member", e, " ".join(f"rows[{r0}:{r0+8}]={err[e, r0:r0+8].max():.2e}" for r0 in range(0, B, 8)))
