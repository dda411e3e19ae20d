"""Candidate-sharded CEM over the GPUs of one box (one process per GPU, torch.distributed).

The planner shards embarrassingly over candidates: rank g rolls out the global candidates
[g n/G, (g+1) n/G) for every environment, particle and ensemble member (weights are replicated, 2.7 MB).  The only
coupling is tf.nn.top_k over all n candidates (cadm/dynamics/core/utils.py:171), so ONE all-gather of the
per-candidate returns [m, n/G] (fp32, 100 B .. 25 kB per rank) per CEM iteration is the whole exchange; every rank
then does the same top-k + refit on identical inputs and arrives at the identical plan without a broadcast.
Action sequences are drawn from a counter-based RNG keyed by the GLOBAL candidate index, so results do not depend
on G and elite sequences owned by other ranks are regenerated locally instead of gathered.

The reference has no multi-GPU mode; this is a new capability behind the same plan.
`backend` is any object with the PlannerEngine phase API (cem_begin / cem_rollout / returns_buffer / cem_refit /
cem_finish and cfg.rank / cfg.world / cfg.cem_iters); the product passes a PlannerEngine.
"""
import torch
import torch.distributed as dist


class ShardedCEMPlanner:
    def __init__(self, backend, group=None):
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
            if self.world > 1:
                self._all_gather(be.returns_buffer())
            be.cem_refit(it)
        return be.cem_finish(logs=logs)
