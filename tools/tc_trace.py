"""Print the clock64 phase trace of the 128-row-tile tensor-core rollout kernel (rollout_tc.cu, CTA 0) for the C2 workload
(run with CADM_TC_VARIANT=1; tools/tcs_sweep.py traces the swapped-operand kernel)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cadm_b200.synth import build_model, synthetic_inputs

prec = sys.argv[1] if len(sys.argv) > 1 else "tc3x"
model, env, cfg = build_model("C2", m_max=1, precision=prec)
inp = synthetic_inputs(env, 1, 30, False)
eng = model.engine
for _ in range(3):
    eng.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=1, logs=False)
eng.set_timing(True)
eng.set_option("trace", 1)
eng.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=2, logs=False)
tr = eng.debug_trace(30)
print("rollout ms (5 launches):", eng.last_rollout_ms())
names = ["start", "prologue_done", "L0_acc", "L0_epi", "L1_acc", "L1_epi", "L2_acc", "L2_epi", "L3_acc", "L3_epi", "head_acc",
         "head_bar", "final_done"]
for t in (1, 2, 15):
    base = tr[t, 0]
    print(f"step {t}: total {tr[t + 1, 0] - base if t + 1 < 30 else -1} cycles")
    print("  epi :", " ".join(f"{n}={tr[t, i] - base}" for i, n in enumerate(names)))
    print("  pro : prefetch=%d reward=%d built=%d" % (tr[t, 13] - base, tr[t, 14] - base, tr[t, 15] - base))
    print("  mma :", " ".join(f"g{g}:first_ready={tr[t, 32 + 4 * g] - base},issued={tr[t, 33 + 4 * g] - base}" for g in range(5)))
