"""A few CEM decisions of one workload: the command ncu wraps (B200_PROFILING.md recipe).
usage: python tools/prof_one.py [config] [m] [n_decisions]     (CADM_TC_VARIANT / CADM_TCS_ROWS / CADM_TCS_KPS honoured)"""
import sys
import torch
sys.path.insert(0, ".")
from cadm_b200.synth import build_model, synthetic_inputs

config = sys.argv[1] if len(sys.argv) > 1 else "C2"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nd = int(sys.argv[3]) if len(sys.argv) > 3 else 3
model, env, cfg = build_model(config, m_max=m, precision="tc3x")
inp = synthetic_inputs(env, m, 30, cfg["context"])
for i in range(nd):
    model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], inp.get("cp_obs"), inp.get("cp_act"), seed=i, logs=False)
torch.cuda.synchronize()
print("kernel:", model.engine.kernel_name)
