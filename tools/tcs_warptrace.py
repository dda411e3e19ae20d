"""Per-warp clock64 detail of one horizon step of the swapped-operand kernel (CTA 0, DIAG build path): for every epilogue warp
and every (layer, M tile) the cycles at which it reached the accumulator wait, passed it, had its first TMEM chunk in registers,
finished its stores, passed the fences and arrived on the layer-input barrier; plus the MMA thread's view of each GEMM."""
import os
import sys
import numpy as np
sys.path.insert(0, ".")
from cadm_b200.synth import build_model, synthetic_inputs

config = sys.argv[1] if len(sys.argv) > 1 else "C2"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 32
os.environ["CADM_TC_VARIANT"] = "2"
os.environ["CADM_TCS_ROWS"] = str(rows)
model, env, cfg = build_model(config, m_max=m, precision="tc3x")
inp = synthetic_inputs(env, m, 30, cfg["context"])
eng = model.engine
args = (inp["obs"], inp["init_mean"], inp["init_var"], inp.get("cp_obs"), inp.get("cp_act"))
for _ in range(3):
    eng.plan_cem(*args, seed=1, logs=False)
eng.set_timing(True)
eng.set_option("trace", 1)
eng.plan_cem(*args, seed=9, logs=False)
print(f"{eng.kernel_name}: {eng.last_rollout_ms() / 5 * 1e3:.1f} us per launch (traced)")
tr = eng.debug_trace(64)
t = 15
base = tr[t, 0]
print(f"step {t}: {tr[t + 1, 0] - base} cycles")
print("MMA thread, per GEMM: xr0_wait_done / xr1_wait_done / tile0_committed / all_issued")
for g in range(5):
    print(f"  g{g}: " + " / ".join(str(tr[t, 32 + 4 * g + i] - base) for i in (0, 2, 3, 1)))
print("epilogue warps (ew: quarter = (ew + 2) & 3, column slice = ew >> 2); per (layer, tile): at_wait, wait_done, first_ld, stores_done, fenced, arrived")
for l in range(4):
    for mt in range(2):
        print(f" layer {l} tile {mt}")
        for ew in range(16):
            w = tr[32 + ew]
            v = [w[(l * 2 + mt) * 6 + i] - base for i in range(6)]
            print(f"   ew{ew:2d} q{(ew + 2) & 3}: " + " ".join(f"{x:6d}" for x in v))
print(" heads: at_wait, wait_done, loop_done, after_bar | final: items_done, published, after_bar")
for ew in range(16):
    w = tr[32 + ew]
    print(f"   ew{ew:2d} q{(ew + 2) & 3}: " + " ".join(f"{w[i] - base:6d}" if w[i] else "     -" for i in range(48, 55)))
eng.close()
