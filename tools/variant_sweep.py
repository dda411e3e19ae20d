import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from cadm_b200.synth import build_model, synthetic_inputs
def run(config, m, variant, rows):
    os.environ["CADM_TC_VARIANT"] = str(variant); os.environ["CADM_TCS_ROWS"] = str(rows)
    model, env, cfg = build_model(config, m_max=m, precision="tc3x")
    inp = synthetic_inputs(env, m, 30, cfg["context"])
    eng = model.engine
    args = (inp["obs"], inp["init_mean"], inp["init_var"], inp.get("cp_obs"), inp.get("cp_act"))
    for _ in range(2): eng.plan_cem(*args, seed=1, logs=False)
    eng.set_timing(True); ms = []
    for i in range(4):
        eng.plan_cem(*args, seed=2 + i, logs=False); ms.append(eng.last_rollout_ms() / 5)
    print(f"{config} m={m} variant={variant} rows={rows}: {np.median(ms)*1e3:8.1f} us per rollout launch [{eng.kernel_name[:18]}]", flush=True)
    eng.close()
for config, m in (("C2", 1), ("C4", 1), ("C2", 10), ("C2", 4), ("C3", 2)):
    run(config, m, 1, 0)
    for rows in (32, 48, 64): run(config, m, 2, rows)
