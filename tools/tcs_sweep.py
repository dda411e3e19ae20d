"""Sweep the swapped-operand rollout kernel's tile rows / stage size on the C2 workload: ms per rollout launch (CUDA events
recorded by the engine around each launch), and the clock64 phase trace of CTA 0 for the default choice."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cadm_b200.synth import build_model, synthetic_inputs

config = sys.argv[1] if len(sys.argv) > 1 else "C2"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1


def run(variant, rows, kps, trace=False, skew=0):
    os.environ["CADM_TC_VARIANT"] = str(variant)
    os.environ["CADM_TCS_ROWS"] = str(rows)
    os.environ["CADM_TCS_KPS"] = str(kps)
    os.environ["CADM_TCS_SKEW"] = str(skew)
    model, env, cfg = build_model(config, m_max=m, precision="tc3x")
    inp = synthetic_inputs(env, m, 30, cfg["context"])
    eng = model.engine
    args = (inp["obs"], inp["init_mean"], inp["init_var"], inp.get("cp_obs"), inp.get("cp_act"))
    for _ in range(3):
        eng.plan_cem(*args, seed=1, logs=False)
    eng.set_timing(True)
    ms = []
    for i in range(5):
        eng.plan_cem(*args, seed=2 + i, logs=False)
        ms.append(eng.last_rollout_ms() / 5)
    if trace:
        eng.set_option("trace", 1)
        eng.plan_cem(*args, seed=9, logs=False)
    print(f"variant={variant} rows={rows} kps={kps} skew={skew}: {np.median(ms) * 1e3:8.1f} us per rollout launch  [{eng.kernel_name}]", flush=True)
    if trace:
        tr = eng.debug_trace(30)
        names = ["start", "prologue_done", "L0_acc", "L0_epi", "L1_acc", "L1_epi", "L2_acc", "L2_epi", "L3_acc", "L3_epi",
                 "head_acc", "head_bar", "final_done"]
        for t in (15,):
            base = tr[t, 0]
            print(f"  step {t}: total {tr[t + 1, 0] - base} cycles")
            print("    epi :", " ".join(f"{n}={tr[t, i] - base}" for i, n in enumerate(names)))
            print("    fin : items_done=%d published=%d bar=%d" % (tr[t, 13] - base, tr[t, 14] - base, tr[t, 15] - base))
            print("    mma :", " ".join(f"g{g}:first_ready={tr[t, 32 + 4 * g] - base},issued={tr[t, 33 + 4 * g] - base},xr1={tr[t, 34 + 4 * g] - base},t0_commit={tr[t, 35 + 4 * g] - base}" for g in range(5)))
    model.engine.close()


run(1, 0, 4)
for rows in (32, 48, 64):
    run(2, rows, 4, trace=(rows == 32))
