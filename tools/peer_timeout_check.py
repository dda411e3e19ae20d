"""Run under torchrun on 2 GPUs: a rank that stops taking part must not hang the others on the device -- the refit kernel's
wait for its returns slice gives up after the configured bound ("peer_timeout_ms", here 2 s; default 30 s) and
cadm_cem_finish of THAT decision raises; "peer_clear_timeout" makes the engine usable again once the ranks are back in step."""
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
eng.set_option("peer_timeout_ms", 2000)
eng.cem_begin(inp["obs"], inp["init_mean"], inp["init_var"])
t0 = time.time()
raised = False
try:
    for it in range(cfg_it := eng.cfg.cem_iters):
        if not (rank == 1 and it == 2):
            eng.cem_rollout(it, seed=2)                                      # rank 1 "dies" in iteration 2: no slice, no flag
        eng.cem_refit(it)
    eng.cem_finish()                         # waits for the refit kernels and checks their report: this decision fails
except CadmError as e:
    raised = True
    print(f"rank {rank}: raised after {time.time() - t0:.1f} s: {e}")
torch.cuda.synchronize()
# both ranks re-synchronise, forget the report and plan again (rank 1 never saw a timeout: rank 0's slices all arrived)
dist.barrier()
eng.set_option("peer_clear_timeout", 1)
recovered = True
try:
    model2, _, _ = build_model("C2", m_max=1, candidates=200 * world, rank=rank, world=world, device=f"cuda:{local}")
    planner2 = ShardedCEMPlanner(model2.engine, fused=True)
    planner2.plan(inp["obs"], inp["init_mean"], inp["init_var"], seed=3)
    torch.cuda.synchronize()
except CadmError as e:
    recovered = False
    print(f"rank {rank}: did not recover: {e}")
flags = [None] * world
dist.all_gather_object(flags, (raised, recovered))
if rank == 0:
    good = flags[0][0] and all(f[1] for f in flags)
    print(f"PEER_TIMEOUT_CHECK {'PASS' if good else 'FAIL'} (rank 0 raised: {flags[0][0]}, recovered: {[f[1] for f in flags]}, "
          f"{time.time() - t0:.1f} s, no hang)")
dist.destroy_process_group()
