"""Host logic of the candidate-sharded planner on CPU: 2 processes, gloo.  The engine is replaced by a NumPy
backend built from the oracle (restricted to the rank's candidate shard) that speaks the same phase API, so what
is tested is the sharding protocol of cadm_b200.parallel: slices, the in-place all-gather layout, identical plans."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardBackend:
    """Phase API of PlannerEngine implemented with the oracle for the candidates of one rank."""

    class Cfg:
        pass

    def __init__(self, rank, world, n, prm, norm, env, E, p):
        from oracle import cadm_oracle as orc
        self.orc = orc
        self.cfg = self.Cfg()
        self.cfg.rank, self.cfg.world, self.cfg.cem_iters = rank, world, orc.NUM_CEM_ITERS
        self.n, self.nl = n, n // world
        self.prm, self.norm, self.env, self.E, self.p = prm, norm, env, E, p
        self.logs = []

    def cem_begin(self, obs, init_mean, init_var, cp_obs=None, cp_act=None):
        self.obs, self.mean, self.var = obs, init_mean.copy(), init_var.copy()
        self.m = obs.shape[0]
        self.buf = torch.zeros(self.cfg.world, self.m, self.nl, dtype=torch.float64)
        self.logs = []

    def cem_rollout(self, it, seed=0, z=None, eps=None):
        from oracle import philox as ph
        m, h, A = self.mean.shape
        self.z_full = ph.gen_z(seed, it + 1, m, self.n, h, A, dtype=np.float64)[it]       # any rank can regenerate
        lo = self.cfg.rank * self.nl
        acts, _ = self.orc.sample_actions(self.mean, self.var, self.z_full[:, lo:lo + self.nl])
        pr, _ = self.orc.rollout(self.obs, acts, self.prm, self.norm, self.env, self.E, self.p, True)
        self.buf[self.cfg.rank] = torch.from_numpy(pr.mean(axis=2))

    def returns_buffer(self):
        return self.buf

    def cem_refit(self, it):
        r = self.buf.numpy().transpose(1, 0, 2).reshape(self.m, self.n)                  # [m, world*n_local]
        acts, _ = self.orc.sample_actions(self.mean, self.var, self.z_full)
        self.mean, self.var, idx = self.orc.refit(self.mean, self.var, acts, r)
        self.logs.append((r.copy(), idx))

    def cem_finish(self, logs=True):
        return dict(mean=self.mean, var=self.var, returns=np.stack([l[0] for l in self.logs]),
                    elites=np.stack([l[1] for l in self.logs]))


def _problem():
    from oracle import cadm_oracle as orc
    from oracle.envs import get_env
    env = get_env("halfcheetah")
    rng = np.random.default_rng(0)
    prm = orc.init_dynamics_params(rng, 1, env.proc_obs_dim + env.act_dim, 32, env.obs_dim, dtype=np.float64)
    norm = orc.NormStats(np.zeros(18), np.ones(18), np.zeros(6), np.full(6, 0.6), np.zeros(18), np.full(18, 0.1)).astype(np.float64)
    obs = rng.standard_normal((2, 18)) * 0.1
    return env, prm, norm, obs, np.zeros((2, 5, 6)), np.full((2, 5, 6), 0.25)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cadm_b200.parallel import ShardedCEMPlanner
        env, prm, norm, obs, m0, v0 = _problem()
        be = OracleShardBackend(rank, world, 64, prm, norm, env, 1, 1)
        planner = ShardedCEMPlanner(be)
        out = planner.plan(obs, m0, v0, seed=5)
        q.put((rank, out["mean"], out["returns"], out["elites"], planner.collectives))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_gloo_matches_single_rank():
    from oracle import cadm_oracle as orc
    from oracle import philox as ph
    env, prm, norm, obs, m0, v0 = _problem()
    z = np.stack([ph.gen_z(5, it + 1, 2, 64, 5, 6, dtype=np.float64)[it] for it in range(5)])
    ref = orc.cem_plan(obs, m0, v0, z, prm, norm, env, 1, 1, True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mean, rets, el, ncoll in res:
        assert ncoll == 5                                         # one all-gather per CEM iteration
        np.testing.assert_allclose(rets, ref.returns, rtol=1e-12, atol=1e-12)
        assert np.array_equal(el, ref.elites)
        np.testing.assert_allclose(mean, ref.mean, rtol=1e-12, atol=1e-12)
    assert np.array_equal(res[0][1], res[1][1])                   # identical plan on every rank, no broadcast


def test_single_rank_needs_no_process_group():
    sys.path.insert(0, ROOT)
    from cadm_b200.parallel import ShardedCEMPlanner
    from oracle import cadm_oracle as orc
    env, prm, norm, obs, m0, v0 = _problem()
    be = OracleShardBackend(0, 1, 64, prm, norm, env, 1, 1)
    pl = ShardedCEMPlanner(be)
    out = pl.plan(obs, m0, v0, seed=5)
    assert pl.collectives == 0 and out["elites"].shape == (5, 2, 50)


# ------------------------------------------------------------------------------------------ environment sharding (8e)

class OracleEnvBackend:
    """`plan_cem` of a world=1 PlannerEngine implemented with the oracle (probabilistic ensemble, injected noise)."""

    class Cfg:
        world, rank, ctx_dim, context_layout = 1, 0, 0, "reference"

    def __init__(self, n, h, A, prm, norm, env, E, p):
        from oracle import cadm_oracle as orc
        self.orc = orc
        self.cfg = self.Cfg()
        self.cfg.ensemble, self.cfg.particles, self.cfg.candidates = E, p, n
        self.cfg.horizon, self.cfg.act_dim, self.cfg.cem_iters = h, A, orc.NUM_CEM_ITERS
        self.prm, self.norm, self.env = prm, norm, env
        self.seeds = []

    def plan_cem(self, obs, init_mean, init_var, cp_obs=None, cp_act=None, seed=0, z=None, eps=None, logs=True):
        self.seeds.append(seed)
        c = self.cfg
        r = self.orc.cem_plan(obs, init_mean, init_var, z, self.prm, self.norm, self.env, c.ensemble, c.particles, False,
                              eps, num_elites=8)
        return dict(mean=r.mean, var=r.var, returns=r.returns, elites=r.elites)


ENV_SHARD = dict(m=3, n=16, h=4, A=6, E=2, p=4)


def _env_problem():
    from oracle import cadm_oracle as orc
    from oracle import philox as ph
    from oracle.envs import get_env
    s = ENV_SHARD
    env = get_env("halfcheetah")
    rng = np.random.default_rng(3)
    prm = orc.init_dynamics_params(rng, s["E"], env.proc_obs_dim + env.act_dim, 24, env.obs_dim, dtype=np.float64)
    norm = orc.NormStats(np.zeros(18), np.ones(18), np.zeros(6), np.full(6, 0.6), np.zeros(18), np.full(18, 0.1)).astype(np.float64)
    obs = rng.standard_normal((s["m"], 18)) * 0.1
    z = ph.gen_z(11, 5, s["m"], s["n"], s["h"], s["A"], dtype=np.float64)
    eps = ph.gen_eps(11, 5, s["h"], s["m"], s["n"], s["p"], s["E"], env.obs_dim, dtype=np.float64)
    m0 = rng.uniform(-0.2, 0.2, (s["m"], s["h"], s["A"]))
    return env, prm, norm, obs, m0, np.full((s["m"], s["h"], s["A"]), 0.25), z, eps


def _env_worker(rank, world, port, q, gather):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cadm_b200.parallel import EnvShardedPlanner
        s = ENV_SHARD
        env, prm, norm, obs, m0, v0, z, eps = _env_problem()
        be = OracleEnvBackend(s["n"], s["h"], s["A"], prm, norm, env, s["E"], s["p"])
        planner = EnvShardedPlanner(be, gather=gather)
        out = planner.plan(obs, m0, v0, seed=7, z=z, eps=eps, logs=True)
        q.put((rank, out["mean"], out["var"], out["returns"], out["elites"], out["bounds"], planner.collectives, be.seeds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gather", [True, False])
def test_env_sharded_two_rank_gloo_matches_single_process(gather):
    """m = 3 environments over 2 ranks (ragged: 2 + 1), probabilistic ensemble with injected z / eps in the GLOBAL
    layouts: every block equals the corresponding rows of the single-process decision exactly; no collective during
    planning, one all-gather of the finished plans only when asked for."""
    from oracle import cadm_oracle as orc
    s = ENV_SHARD
    env, prm, norm, obs, m0, v0, z, eps = _env_problem()
    ref = orc.cem_plan(obs, m0, v0, z, prm, norm, env, s["E"], s["p"], False, eps, num_elites=8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_env_worker, args=(r, 2, port, q, gather)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mean, var, rets, el, bounds, ncoll, seeds in res:
        assert bounds == [0, 2, 3]
        lo, hi = bounds[rank], bounds[rank + 1]
        assert np.array_equal(rets, ref.returns[:, lo:hi]) and np.array_equal(el, ref.elites[:, lo:hi])
        if gather:
            assert ncoll == 1 and np.array_equal(mean, ref.mean) and np.array_equal(var, ref.var)
        else:
            assert ncoll == 0 and np.array_equal(mean, ref.mean[lo:hi]) and np.array_equal(var, ref.var[lo:hi])
    assert res[0][7] != res[1][7]                                 # per-rank Philox keys for seed-only noise


def test_env_shard_bounds_and_guards():
    sys.path.insert(0, ROOT)
    from cadm_b200.parallel import EnvShardedPlanner, env_shard_bounds
    assert env_shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert env_shard_bounds(2, 4) == [0, 1, 2, 2, 2]              # more ranks than environments: empty blocks
    assert env_shard_bounds(16, 8) == list(range(0, 17, 2))

    class Cfg:
        world, ctx_dim, ensemble, context_layout = 1, 10, 5, "reference"

    class Be:
        cfg = Cfg()

    with pytest.raises(ValueError, match="context_layout"):      # quirk Q3 mixes environments: refuse, do not "fix"
        EnvShardedPlanner(Be(), rank=0, world=2, gather=False)
    EnvShardedPlanner(Be(), rank=0, world=2, gather=False, allow_local_context=True)
    EnvShardedPlanner(Be(), rank=0, world=1)                      # one rank: the pairing is the reference's own
    Cfg.context_layout = "matched"
    pl = EnvShardedPlanner(Be(), rank=1, world=2, gather=False)
    assert pl.rank_seed(5) != 5 and EnvShardedPlanner(Be(), rank=0, world=2, gather=False).rank_seed(5) == 5
    Cfg.world = 2
    with pytest.raises(ValueError, match="world=1"):
        EnvShardedPlanner(Be(), rank=0, world=2, gather=False)
