"""clock64 trace of pair 0 of the CTA-pair kernel at step h / 2 (DIAG instantiation): the MMA thread and epilogue warp 0 of both CTAs."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from cadm_b200.synth import build_model, synthetic_inputs
config = sys.argv[1] if len(sys.argv) > 1 else "C2"
rows = sys.argv[2] if len(sys.argv) > 2 else "32"
os.environ["CADM_TC_VARIANT"] = "3"; os.environ["CADM_TCS_ROWS"] = rows
model, env, cfg = build_model(config, m_max=1, precision="tc3x")
inp = synthetic_inputs(env, 1, 30, cfg["context"])
eng = model.engine
args = (inp["obs"], inp["init_mean"], inp["init_var"], inp.get("cp_obs"), inp.get("cp_act"))
for _ in range(3): eng.plan_cem(*args, seed=1, logs=False)
eng.set_timing(True)
ms = []
for i in range(4):
    eng.plan_cem(*args, seed=2 + i, logs=False); ms.append(eng.last_rollout_ms() / 5)
print(f"{eng.kernel_name}: {np.median(ms)*1e3:.1f} us per launch")
eng.set_option("trace", 1)
eng.plan_cem(*args, seed=9, logs=False)
tr = eng.debug_trace(64).reshape(-1)
base = tr[0]
for rank in range(2):
    e = tr[rank * 128: rank * 128 + 64]; m = tr[rank * 128 + 64: rank * 128 + 128]
    print(f"rank {rank}: epilogue warp 0 (start {e[0]-base}); per (layer, subtile): at_wait, wait_done, stores_done, fenced, arrived")
    for l in range(4):
        for s in range(2):
            v = [e[1 + (l * 2 + s) * 5 + i] - base for i in range(5)]
            print(f"   L{l} {'AB'[s]}: " + " ".join(f"{x:6d}" for x in v))
    print(f"   head: at_wait {e[52]-base} wait_done {e[53]-base}  final_items_done {e[54]-base} published {e[55]-base} bar {e[56]-base}")
    print(f"rank {rank}: MMA thread (start {m[0]-base}); per (GEMM, subtile): at_wait, wait_done, committed")
    for g in range(4):
        for s in range(2):
            v = [m[1 + (g * 2 + s) * 3 + i] - base for i in range(3)]
            print(f"   G{g} {'AB'[s]}: " + " ".join(f"{x:6d}" for x in v))
    print(f"   heads: at_wait {m[40]-base} wait_done {m[41]-base} committed {m[42]-base}")
eng.close()
