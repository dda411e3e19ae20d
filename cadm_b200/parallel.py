"""Candidate-sharded CEM over the GPUs of one box (one process per GPU, torch.distributed).

The planner shards embarrassingly over candidates: rank g rolls out the global candidates
[g n/G, (g+1) n/G) for every environment, particle and ensemble member (weights are replicated, 2.7 MB).  The only
coupling is tf.nn.top_k over all n candidates (cadm/dynamics/core/utils.py:171), so ONE all-gather of the
per-candidate returns [m, n/G] (fp32, 100 B .. 25 kB per rank) per CEM iteration is the whole exchange; every rank
then does the same top-k + refit on identical inputs and arrives at the identical plan without a broadcast.
Action sequences are drawn from a counter-based RNG keyed by the GLOBAL candidate index, so results do not depend
on G and elite sequences owned by other ranks are regenerated locally instead of gathered.

Two transports for that exchange:
  fused (default on GPUs of one node)  the engines' exchange blocks are shared over CUDA IPC once; from then on the kernel
         that averages the particle returns also stores the rank's slice into EVERY rank's buffer (peer stores over NVLink /
         NVSwitch) and publishes an epoch flag, and the refit kernel waits on the flags on the device -- no collective call,
         no host synchronisation between the phases (include/cadm_b200.h cadm_peer_*; csrc/cem_kernels.cu)
  NCCL   one dist.all_gather_into_tensor per iteration, in place on the engine's returns buffer (fused=False, or the
         fallback when IPC is unavailable; gloo on CPU for the tests)

The reference has no multi-GPU mode; this is a new capability behind the same plan.
`backend` is any object with the PlannerEngine phase API (cem_begin / cem_rollout / returns_buffer / cem_refit /
cem_finish and cfg.rank / cfg.world / cfg.cem_iters); the product passes a PlannerEngine.
"""
import os

import torch
import torch.distributed as dist


class ShardedCEMPlanner:
    def __init__(self, backend, group=None, fused=None):
        self.backend = backend
        self.group = group
        self.world = backend.cfg.world
        self.rank = backend.cfg.rank
        if self.world > 1:
            if not dist.is_initialized():
                raise RuntimeError("torch.distributed must be initialised for world > 1")
            if dist.get_world_size(group) != self.world or dist.get_rank(group) != self.rank:
                raise ValueError("engine rank/world do not match the process group")
        self.collectives = 0
        self.fused = False
        if fused is None and os.environ.get("CADM_FUSED_ALLGATHER") == "0":
            fused = False
        if self.world > 1 and fused is not False and hasattr(backend, "peer_export"):
            self.fused = self._attach_peers(required=fused is True)

    def _attach_peers(self, required):
        """Exchange the CUDA IPC handles of the engines' exchange blocks and attach them: the per-iteration all-gather then
        happens inside the rollout phase (peer stores over NVLink + device-side flags) instead of through NCCL."""
        be = self.backend
        ok, handles = 1, [None] * self.world
        try:
            mine = be.peer_export()
        except Exception:
            if required:
                raise
            ok, mine = 0, b""
        dist.all_gather_object(handles, mine, group=self.group)
        if ok and all(h is not None and len(h) == 64 for h in handles):
            try:
                be.peer_attach(handles)
            except Exception:
                if required:
                    raise
                ok = 0
        else:
            ok = 0
        # fused only if EVERY rank attached (the protocol has no mixed mode)
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=self.group)
        if not all(flags):
            if ok:
                raise RuntimeError("peer attach succeeded on some ranks only; restart with fused=False")
            return False
        dist.barrier(group=self.group)
        return True

    def _all_gather(self, buf: torch.Tensor):
        """In-place all-gather of buf [world, m, n_local]: rank r contributes slice r."""
        flat = buf.view(self.world, -1)
        mine = flat[self.rank]
        if not buf.is_cuda:
            mine = mine.clone()            # gloo: no in-place aliasing
        dist.all_gather_into_tensor(flat.view(-1), mine, group=self.group)
        self.collectives += 1

    def plan(self, obs, init_mean, init_var, cp_obs=None, cp_act=None, seed=0, z=None, eps=None, logs=True):
        be = self.backend
        be.cem_begin(obs, init_mean, init_var, cp_obs, cp_act)
        for it in range(be.cfg.cem_iters):
            be.cem_rollout(it, seed=seed, z=z, eps=eps)
            if self.world > 1 and not self.fused:
                self._all_gather(be.returns_buffer())
            be.cem_refit(it)
        return be.cem_finish(logs=logs)
