"""Run under torchrun on G GPUs: candidate-sharded CEM must give, on every rank, exactly the plan a single rank computes
for the same global problem (n = 200 G candidates) -- with the fused peer-memory all-gather and with NCCL; and
environment-sharded CEM (EnvShardedPlanner) must reproduce the single-process decision for 2 G environments."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# bit-identity needs the same rollout kernel on the sharded ranks and on the single-rank reference (in auto mode the
# engine picks the kernel from the LOCAL batch size; the two tensor-core kernels agree to fp32 rounding, not to the bit)
os.environ.setdefault("CADM_TC_VARIANT", "2")
from cadm_b200.parallel import ShardedCEMPlanner
from cadm_b200.synth import build_model, synthetic_inputs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 200 * world
m = 2
ok = True
ref = None
for fused in (True, False):
    model, env, cfg = build_model("C2", m_max=m, candidates=n, rank=rank, world=world, device=f"cuda:{local}")
    inp = synthetic_inputs(env, m, 30, False, seed=3)
    planner = ShardedCEMPlanner(model.engine, fused=fused)
    assert planner.fused == fused
    for rep in range(3):                                   # several decisions: epochs / parities wrap
        out = planner.plan(inp["obs"], inp["init_mean"], inp["init_var"], seed=77)
    torch.cuda.synchronize()
    if rank == 0:
        if ref is None:
            single, _, _ = build_model("C2", m_max=m, candidates=n, device=f"cuda:{local}")
            ref = single.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=77)
        for k in ("mean", "var", "returns", "elites"):
            same = torch.equal(out[k], ref[k])
            ok &= same
            print(f"[{'fused' if fused else 'nccl '}] rank0 vs single-rank {k}: {'bit-identical' if same else 'DIFFERENT'}")
    gathered = [torch.empty_like(out["mean"]) for _ in range(world)]
    dist.all_gather(gathered, out["mean"])
    same_all = all(torch.equal(g, gathered[0]) for g in gathered)
    ok &= same_all
    if rank == 0:
        print(f"[{'fused' if fused else 'nccl '}] plans identical on all {world} ranks: {same_all}; NCCL all-gathers per decision: "
              f"{planner.collectives // 3}")
# ---- environment sharding (m >= G): every rank plans a block of environments on a world = 1 engine, injected
# noise in the global layouts; the gathered plan must equal the single-process decision to fp32 rounding (the kernel is
# picked from the local batch) and be the same on every rank
from cadm_b200.parallel import EnvShardedPlanner
import numpy as np
m_env, n_env = 2 * world, 64
model, env, cfg = build_model("C2", m_max=m_env, candidates=n_env, device=f"cuda:{local}")
inp = synthetic_inputs(env, m_env, 30, False, seed=5)
noise = np.random.default_rng(13)                         # same seed on every rank: identical injected noise
z = np.clip(noise.standard_normal((5, m_env, n_env, 30, env.act_dim)), -2, 2).astype(np.float32)
eps = noise.standard_normal((5, 30, cfg["ensemble"], cfg["particles"] // cfg["ensemble"] * m_env * n_env, env.obs_dim)).astype(np.float32)
planner = EnvShardedPlanner(model.engine, gather=True)
out = planner.plan(inp["obs"], inp["init_mean"], inp["init_var"], seed=0, z=z, eps=eps)
full = model.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=0, z=z, eps=eps)
torch.cuda.synchronize()
err = float((out["mean"] - full["mean"]).abs().max())
gathered = [torch.empty_like(out["mean"]) for _ in range(world)]
dist.all_gather(gathered, out["mean"])
same_all = all(torch.equal(g, gathered[0]) for g in gathered)
ok &= same_all and err < 1e-4 and planner.collectives == 1
if rank == 0:
    print(f"[envs ] {m_env} environments over {world} ranks: max |plan - single-process plan| = {err:.2e}; identical on all ranks: "
          f"{same_all}; collectives per decision: {planner.collectives}")
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()
