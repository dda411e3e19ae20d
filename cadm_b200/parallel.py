"""CEM planning over the GPUs of one box (one process per GPU, torch.distributed): candidates sharded with one small
exchange per iteration (ShardedCEMPlanner), or environments sharded with none (EnvShardedPlanner, end of file).

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

import numpy as np
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


def env_shard_bounds(m, world):
    """Contiguous, balanced blocks of environments: rank r plans for [lo[r], lo[r+1]).  Ragged m is allowed (the first
    m % world ranks take one more); ranks beyond m get an empty block."""
    base, extra = divmod(int(m), int(world))
    lo = [0]
    for r in range(world):
        lo.append(lo[-1] + base + (1 if r < extra else 0))
    return lo


_SEED_STRIDE = 0x9E3779B97F4A7C15      # odd 64-bit constant: per-rank Philox keys that never collide for world < 2^63


class EnvShardedPlanner:
    """The zero-exchange alternative of SURVEY.md section 8(e) for m >= G: shard ENVIRONMENTS instead of candidates.

    The reference always plans for 10-20 environments at once (`num_rollouts`, cadm/samplers/sampler.py:107-120) and
    the environments of one decision never interact: every (env, candidate, particle) trajectory, the top-k and the
    refit are per environment (cadm/dynamics/core/utils.py:137-182).  So rank r runs the WHOLE decision -- all n
    candidates, all iterations, on a world=1 engine -- for its block of environments, and nothing is exchanged during
    planning.  What remains is optional: one all-gather of the finished plans [m/G, h, A] when every rank wants the
    complete plan (`gather=True`; a rank that also steps its own environments passes gather=False and there is no
    collective at all).

    Exactness.  With injected noise (`z`, `eps` in the global layouts of PlannerEngine.plan_cem) the blocks reproduce
    the single-process decision: element for element under the float64 oracle (tests/test_sharded_gloo.py) and, on
    the engine, whenever both runs select the same rollout kernel.  Two things depend on m
    and are handled explicitly rather than silently:
      * the reference's odd-iteration context pairing (quirk Q3, core/utils.py:433-434) reinterprets the [E, m, C]
        context tensor as [m, E, C] and so mixes environments when m > 1 and E > 1; a block of m/G environments
        cannot reproduce that, therefore a backend with a context encoder must use context_layout="matched" (or
        E == 1) unless `allow_local_context=True` accepts the pairing of the local block;
      * seed-only noise is keyed by the GLOBAL environment index: the block's first environment is handed to the
        engine as its "env_offset" option (added to the local index in every Philox counter, oracle/philox.py
        `m_offset`), every rank uses the same key, and the block draws exactly the numbers it draws inside the
        unsharded decision.  A backend without `set_option` (the test doubles) falls back to per-rank keys
        seed + rank * 0x9E3779B97F4A7C15 (mod 2^64): same distribution, different numbers.
    `backend` is a PlannerEngine built with world=1 (or any object with `plan_cem` and `cfg`)."""

    def __init__(self, backend, rank=None, world=None, group=None, gather=True, allow_local_context=False):
        self.backend, self.group, self.gather = backend, group, gather
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world = int(rank), int(world)
        cfg = backend.cfg
        if getattr(cfg, "world", 1) != 1:
            raise ValueError("environment sharding needs an engine that owns all candidates (world=1)")
        if self.world > 1 and gather and not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised to gather the plans of world > 1")
        mixes = getattr(cfg, "ctx_dim", 0) > 0 and getattr(cfg, "ensemble", 1) > 1 and \
            getattr(cfg, "context_layout", "reference") == "reference"
        if mixes and self.world > 1 and not allow_local_context:
            raise ValueError("context_layout='reference' pairs contexts across environments on odd CEM iterations "
                             "(core/utils.py:433-434); shard candidates instead, build the engine with "
                             "context_layout='matched', or pass allow_local_context=True")
        self.collectives = 0

    def bounds(self, m):
        return env_shard_bounds(m, self.world)

    def rank_seed(self, seed):
        return (int(seed) + self.rank * _SEED_STRIDE) & 0xFFFFFFFFFFFFFFFF

    @staticmethod
    def _rows(x, lo, hi):
        return None if x is None else x[lo:hi]

    def _slice_eps(self, eps, m, lo, hi):
        """eps [iters, h, E, R, D] with R = (p/E) m n and row = j m n + mi n + ni (core/utils.py:144-154) -> the rows of
        environments lo..hi-1 in the same order for the local problem."""
        if eps is None:
            return None
        c = self.backend.cfg
        q, n = c.particles // c.ensemble, c.candidates
        it, h, E, R, D = eps.shape
        if R != q * m * n:
            raise ValueError(f"eps has {R} rows per member, expected (p/E) m n = {q * m * n}")
        return eps.reshape(it, h, E, q, m, n, D)[:, :, :, :, lo:hi].reshape(it, h, E, q * (hi - lo) * n, D)

    def plan(self, obs, init_mean, init_var, cp_obs=None, cp_act=None, seed=0, z=None, eps=None, logs=False):
        """Global inputs on every rank ([m, ...]; only the rank's block is read) -> dict(mean, var): the complete
        [m, h, A] plan on every rank (gather=True) or the rank's block (gather=False), plus `bounds` and, with
        logs=True, the block's `returns` / `elites`."""
        m = obs.shape[0]
        lo_all = self.bounds(m)
        lo, hi = lo_all[self.rank], lo_all[self.rank + 1]
        out = dict(bounds=lo_all, mean=None, var=None)
        if hi > lo:
            zz = None if z is None else z[:, lo:hi]
            if hasattr(self.backend, "set_option"):
                self.backend.set_option("env_offset", lo)          # Philox counters see the global environment index
                block_seed = int(seed)
            else:
                block_seed = self.rank_seed(seed)
            local = self.backend.plan_cem(obs[lo:hi], init_mean[lo:hi], init_var[lo:hi], self._rows(cp_obs, lo, hi),
                                          self._rows(cp_act, lo, hi), seed=block_seed, z=zz,
                                          eps=self._slice_eps(eps, m, lo, hi), logs=logs)
            out.update(local)
        if not self.gather or self.world == 1:
            return out
        # one padded all-gather of [mean | var] blocks; ragged blocks are cut back afterwards
        c = self.backend.cfg
        hA = c.horizon * c.act_dim
        width = max(lo_all[r + 1] - lo_all[r] for r in range(self.world))
        # ONE dtype and ONE device on every rank, whether or not its block is empty: the engine's float32 on its CUDA device
        # (a test double on the CPU may declare another `plan_dtype`); never taken from the caller's arrays
        dev = getattr(self.backend, "device", None)
        dev = torch.device(dev) if dev is not None else torch.device("cpu")
        dt = getattr(self.backend, "plan_dtype", torch.float32)
        as_t = lambda a: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))).to(device=dev, dtype=dt)
        mine = torch.zeros((2, width, hA), dtype=dt, device=dev)
        if hi > lo:
            mine[0, :hi - lo] = as_t(out["mean"]).reshape(hi - lo, hA)
            mine[1, :hi - lo] = as_t(out["var"]).reshape(hi - lo, hA)
        full = torch.empty((self.world, 2, width, hA), dtype=dt, device=dev)
        dist.all_gather_into_tensor(full.view(-1), mine.view(-1), group=self.group)
        self.collectives += 1
        parts = [full[r, :, :lo_all[r + 1] - lo_all[r]] for r in range(self.world)]
        both = torch.cat(parts, dim=1).reshape(2, m, c.horizon, c.act_dim)
        conv = (lambda t: t) if torch.is_tensor(init_mean) or dev.type != "cpu" else (lambda t: t.numpy())
        out["mean"], out["var"] = conv(both[0]), conv(both[1])
        return out
