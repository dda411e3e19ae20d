"""Run under torchrun on G GPUs: candidate-sharded CEM over NCCL must give, on every rank, exactly the plan a single
rank computes for the same global problem (n = 200 G candidates)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadm_b200.parallel import ShardedCEMPlanner
from cadm_b200.synth import build_model, synthetic_inputs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 200 * world
m = 2
model, env, cfg = build_model("C2", m_max=m, candidates=n, rank=rank, world=world, device=f"cuda:{local}")
inp = synthetic_inputs(env, m, 30, False, seed=3)
planner = ShardedCEMPlanner(model.engine)
out = planner.plan(inp["obs"], inp["init_mean"], inp["init_var"], seed=77)
torch.cuda.synchronize()
ok = True
if rank == 0:
    single, _, _ = build_model("C2", m_max=m, candidates=n, device=f"cuda:{local}")
    ref = single.engine.plan_cem(inp["obs"], inp["init_mean"], inp["init_var"], seed=77)
    for k in ("mean", "var", "returns", "elites"):
        same = torch.equal(out[k], ref[k])
        ok &= same
        print(f"rank0 vs single-rank {k}: {'bit-identical' if same else 'DIFFERENT'}")
# every rank holds the same plan
gathered = [torch.empty_like(out["mean"]) for _ in range(world)]
dist.all_gather(gathered, out["mean"])
same_all = all(torch.equal(g, gathered[0]) for g in gathered)
if rank == 0:
    print(f"plans identical on all {world} ranks: {same_all}; all-gathers per decision: {planner.collectives}")
    print("MULTI_GPU_CHECK", "PASS" if ok and same_all else "FAIL")
dist.destroy_process_group()
