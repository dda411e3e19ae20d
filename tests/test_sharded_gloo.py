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
    """`plan_cem` of a world=1 PlannerEngine implemented with the oracle (probabilistic ensemble; injected noise, or the
    Philox specification with the engine's "env_offset" option when only a seed is given)."""
    plan_dtype = torch.float64

    class Cfg:
        world, rank, ctx_dim, context_layout = 1, 0, 0, "reference"

    def __init__(self, n, h, A, prm, norm, env, E, p, with_options=True):
        from oracle import cadm_oracle as orc
        self.orc = orc
        self.cfg = self.Cfg()
        self.cfg.ensemble, self.cfg.particles, self.cfg.candidates = E, p, n
        self.cfg.horizon, self.cfg.act_dim, self.cfg.cem_iters = h, A, orc.NUM_CEM_ITERS
        self.prm, self.norm, self.env = prm, norm, env
        self.seeds, self.env_offset = [], 0
        if with_options:
            self.set_option = self._set_option

    def _set_option(self, name, value):
        assert name == "env_offset"
        self.env_offset = int(value)

    def plan_cem(self, obs, init_mean, init_var, cp_obs=None, cp_act=None, seed=0, z=None, eps=None, logs=True):
        from oracle import philox as ph
        self.seeds.append(seed)
        c = self.cfg
        if z is None:            # seed-only: the counter-based streams, keyed by the GLOBAL environment index
            m = obs.shape[0]
            z = ph.gen_z(seed, c.cem_iters, m, c.candidates, c.horizon, c.act_dim, dtype=np.float64, m_offset=self.env_offset)
            eps = ph.gen_eps(seed, c.cem_iters, c.horizon, m, c.candidates, c.particles, c.ensemble, obs.shape[1], dtype=np.float64,
                             m_offset=self.env_offset)
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


def _env_worker(rank, world, port, q, gather, seed_only=False, m_use=None):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cadm_b200.parallel import EnvShardedPlanner
        s = ENV_SHARD
        env, prm, norm, obs, m0, v0, z, eps = _env_problem()
        if m_use is not None:
            obs, m0, v0, z, eps = obs[:m_use], m0[:m_use], v0[:m_use], None, None
        if seed_only:
            z = eps = None
        be = OracleEnvBackend(s["n"], s["h"], s["A"], prm, norm, env, s["E"], s["p"])
        planner = EnvShardedPlanner(be, gather=gather)
        out = planner.plan(obs, m0, v0, seed=7, z=z, eps=eps, logs=True)
        q.put((rank, out["mean"], out["var"], out.get("returns"), out.get("elites"), out["bounds"], planner.collectives, be.seeds,
               be.env_offset))
    finally:
        dist.destroy_process_group()


def _run_env_workers(gather, **kw):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_env_worker, args=(r, 2, port, q, gather), kwargs=kw) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("gather", [True, False])
def test_env_sharded_two_rank_gloo_matches_single_process(gather):
    """m = 3 environments over 2 ranks (ragged: 2 + 1), probabilistic ensemble with injected z / eps in the GLOBAL
    layouts: every block equals the corresponding rows of the single-process decision exactly; no collective during
    planning, one all-gather of the finished plans only when asked for."""
    from oracle import cadm_oracle as orc
    s = ENV_SHARD
    env, prm, norm, obs, m0, v0, z, eps = _env_problem()
    ref = orc.cem_plan(obs, m0, v0, z, prm, norm, env, s["E"], s["p"], False, eps, num_elites=8)
    res = _run_env_workers(gather)
    for rank, mean, var, rets, el, bounds, ncoll, seeds, off in res:
        assert bounds == [0, 2, 3]
        lo, hi = bounds[rank], bounds[rank + 1]
        assert np.array_equal(rets, ref.returns[:, lo:hi]) and np.array_equal(el, ref.elites[:, lo:hi])
        if gather:
            assert ncoll == 1 and np.array_equal(mean, ref.mean) and np.array_equal(var, ref.var)
        else:
            assert ncoll == 0 and np.array_equal(mean, ref.mean[lo:hi]) and np.array_equal(var, ref.var[lo:hi])
    assert res[0][7] == res[1][7] == [7] and [r[8] for r in res] == [0, 2]      # one key, the block's first environment as offset


def test_env_sharded_seed_only_is_bit_identical():
    """Seed-only noise: with the block's first environment as "env_offset" in the Philox counters, both blocks draw the numbers
    they draw inside the unsharded decision -- the gathered plan equals the single-process plan bit for bit."""
    from oracle import cadm_oracle as orc
    from oracle import philox as ph
    s = ENV_SHARD
    env, prm, norm, obs, m0, v0, _, _ = _env_problem()
    z = ph.gen_z(7, orc.NUM_CEM_ITERS, s["m"], s["n"], s["h"], s["A"], dtype=np.float64)
    eps = ph.gen_eps(7, orc.NUM_CEM_ITERS, s["h"], s["m"], s["n"], s["p"], s["E"], obs.shape[1], dtype=np.float64)
    ref = orc.cem_plan(obs, m0, v0, z, prm, norm, env, s["E"], s["p"], False, eps, num_elites=8)
    for rank, mean, var, rets, el, bounds, ncoll, seeds, off in _run_env_workers(True, seed_only=True):
        lo, hi = bounds[rank], bounds[rank + 1]
        assert np.array_equal(rets, ref.returns[:, lo:hi]) and np.array_equal(el, ref.elites[:, lo:hi])
        assert np.array_equal(mean, ref.mean) and np.array_equal(var, ref.var)


def test_env_sharded_empty_block_gathers():
    """More ranks than environments (m = 1, 2 ranks): the rank with the empty block joins the all-gather with buffers of
    the same dtype and device as the rank that planned and receives the complete plan."""
    res = _run_env_workers(True, seed_only=True, m_use=1)
    assert res[0][5] == [0, 1, 1] and res[1][3] is None           # rank 1 planned nothing
    assert res[0][1].shape == (1, ENV_SHARD["h"], ENV_SHARD["A"])
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2]) and res[1][6] == 1


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
    assert pl.rank_seed(5) != 5 and EnvShardedPlanner(Be(), rank=0, world=2, gather=False).rank_seed(5) == 5   # fallback keys
    Cfg.world = 2
    with pytest.raises(ValueError, match="world=1"):
        EnvShardedPlanner(Be(), rank=0, world=2, gather=False)


# ------------------------------------------------------------------------------------------ get_action() at world > 1
def _get_action_worker(rank, world, port, q):
    """DynamicsModel.get_action() -> PlannerModelBase._plan -> the candidate-sharded planner (missing 2): the
    model object is the product's class with a phase-API test double in place of the engine (no GPU here)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cadm_b200.dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel
        env, prm, norm, obs, m0, v0 = _problem()
        be = OracleShardBackend(rank, world, 64, prm, norm, env, 1, 1)
        model = object.__new__(MLPEnsembleCEMDynamicsModel)          # the planning surface only: no engine, no device
        model.engine, model.use_cem, model._seed, model._calls = be, True, 5, 0
        model.obs_space_dims, model.n_forwards, model.action_space_dims, model.discrete = 18, 5, 6, False
        model.normalize_input = False
        a0 = model.get_action(obs.astype(np.float32), m0.astype(np.float32), v0.astype(np.float32))
        a1 = model.get_action(obs.astype(np.float32), m0.astype(np.float32), v0.astype(np.float32))    # next seed of the model's counter
        q.put((rank, a0, a1, model.sharded_planner().collectives, model.sharded_planner().fused))
    finally:
        dist.destroy_process_group()


def test_get_action_two_rank_gloo():
    from oracle import cadm_oracle as orc
    from oracle import philox as ph
    env, prm, norm, obs, m0, v0 = _problem()
    f = lambda a: a.astype(np.float32).astype(np.float64)
    want = []
    for call in range(2):
        seed = (5 << 32) | call                                       # PlannerModelBase._next_seed
        z = np.stack([ph.gen_z(seed, it + 1, 2, 64, 5, 6, dtype=np.float64)[it] for it in range(5)])
        ref = orc.cem_plan(f(obs), f(m0), f(v0), z, prm, norm, env, 1, 1, True)
        want.append(np.clip(ref.mean, -1.0, 1.0).astype(np.float32))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_get_action_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a0, a1, ncoll, fused in res:
        assert not fused and ncoll == 10                              # gloo: one all-gather per CEM iteration, two decisions
        assert a0.dtype == np.float32 and a0.shape == (2, 5, 6) and np.abs(a0).max() <= 1.0
        np.testing.assert_allclose(a0, want[0], rtol=0, atol=1e-6)
        np.testing.assert_allclose(a1, want[1], rtol=0, atol=1e-6)
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])      # every rank returns the same plan
