"""Run under torchrun on 2 GPUs: a rank that stops taking part must not hang the others on the device -- the refit kernel's
wait for its returns slice gives up after ~2 s and the next host call raises."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadm_b200._lib import CadmError
from cadm_b200.parallel import ShardedCEMPlanner
from cadm_b200.synth import build_model, synthetic_inputs

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
model, env, cfg = build_model("C2", m_max=1, candidates=200 * world, rank=rank, world=world, device=f"cuda:{local}")
inp = synthetic_inputs(env, 1, 30, False, seed=3)
planner = ShardedCEMPlanner(model.engine, fused=True)
planner.plan(inp["obs"], inp["init_mean"], inp["init_var"], seed=1)          # a healthy decision first
torch.cuda.synchronize()
dist.barrier()
eng = model.engine
eng.cem_begin(inp["obs"], inp["init_mean"], inp["init_var"])
t0 = time.time()
raised = False
try:
    for it in range(cfg_it := eng.cfg.cem_iters):
        if not (rank == 1 and it == 2):
            eng.cem_rollout(it, seed=2)                                      # rank 1 "dies" in iteration 2: no slice, no flag
        eng.cem_refit(it)
    eng.cem_finish()
    torch.cuda.synchronize()                 # the launches are asynchronous: the report surfaces on the next call
    eng.cem_begin(inp["obs"], inp["init_mean"], inp["init_var"])
except CadmError as e:
    raised = True
    print(f"rank {rank}: raised after {time.time() - t0:.1f} s: {e}")
torch.cuda.synchronize()
if rank == 0:
    print(f"PEER_TIMEOUT_CHECK {'PASS' if raised else 'FAIL'} ({time.time() - t0:.1f} s, no hang)")
dist.destroy_process_group()
