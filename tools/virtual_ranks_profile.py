"""G virtual ranks on ONE device (the all-gather emulated by copies): the command ncu wraps to see what the per-rank kernels
of a sharded decision cost at the weak-scaling sizes (n = 200 G).  usage: python tools/virtual_ranks_profile.py [G]"""
import sys
import torch
sys.path.insert(0, ".")
from cadm_b200.synth import build_model, synthetic_inputs

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
config = sys.argv[2] if len(sys.argv) > 2 else "C2"                 # C2: weak scaling (200 candidates per rank); C4: the named n = 1000 split
ranks = [build_model(config, m_max=1, candidates=200 * G if config == "C2" else 1000, rank=r, world=G)[0] for r in range(G)]
env = ranks[0].env
ctx = config in ("C3", "C4")
inp = synthetic_inputs(env, 1, 30, ctx, seed=8)
for dec in range(2):
    for r in ranks:
        r.engine.cem_begin(inp["obs"], inp["init_mean"], inp["init_var"], inp.get("cp_obs"), inp.get("cp_act"))
    for it in range(5):
        for r in ranks:
            r.engine.cem_rollout(it, seed=99 + dec)
        bufs = [r.engine.returns_buffer() for r in ranks]
        for i, bi in enumerate(bufs):
            for j, bj in enumerate(bufs):
                if i != j:
                    bi[j].copy_(bj[j])
        for r in ranks:
            r.engine.cem_refit(it)
    outs = [r.engine.cem_finish() for r in ranks]
torch.cuda.synchronize()
print("ok")
