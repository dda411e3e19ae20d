"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the CaDM CEM/MPC planner hot path.

PARITY STATUS: pinned by the reference's own graph code.  /root/reference holds no tests, golden
vectors or fixtures for this path (its only test is the failing stub cadm/__init__.py:1-9) and
TensorFlow 1.15 cannot be imported or installed here; but the planner is Python that composes ~35 TF
functions, so tests/golden/make_reference_golden.py runs the UNMODIFIED builder functions of
cadm/dynamics/core/utils.py and the unmodified environment classes over a NumPy stand-in for those
functions (tests/golden/tf_numpy_shim.py), in float64, with the same weights and injected noise, and
records what they compute (tests/golden/recorded/planner_reference*.npz).  tests/test_reference_pinned.py
requires this oracle and the committed fixtures to reproduce those recordings (returns, elite indices,
plans; CEM and random shooting; all six environments).  Not exercised by that: TensorFlow's kernels and
random generators (noise is injected).  The hand-derived known-answer tests in tests/test_oracle_kat.py
 remain as independent pins.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's `cpu_baseline` / `--impl reference` legs may
import this package; the product (`cadm_b200/`) never does.

What is restated (all paths relative to /root/reference):
  cadm/dynamics/core/utils.py:635-647   create_dense_layer   -> dense()
  cadm/dynamics/core/utils.py:73-92     forward (PE-TS)      -> forward()
  cadm/dynamics/core/utils.py:341-370   forward (CaDM)       -> forward()      (same math)
  cadm/dynamics/core/utils.py:569-624   context encoder      -> encode_context()
  cadm/dynamics/core/utils.py:118-184   CEM planner (PE-TS)  -> cem_plan()
  cadm/dynamics/core/utils.py:400-488   CEM planner (CaDM)   -> cem_plan(ctx=...)
  cadm/dynamics/core/utils.py:186-246, 490-561 random shooting -> rs_plan()
  cadm/dynamics/core/utils.py:627-632   normalize/denormalize
  cadm/dynamics/mlp_ensemble_cem_dynamics.py:191-207, mlp_cadm_...:344-367  get_action clip

The planner is written with the SAME tile / transpose / reshape sequence as the TF graph (so the
reference's layout quirks Q1-Q3 of SURVEY.md appendix A are reproduced by construction, not by
formula); `row_maps()` gives the closed-form index maps, and tests check the two agree.

Randomness is an explicit input: `z` are the truncated-normal draws of `tf.random.truncated_normal`
(utils.py:135) and `eps` the draws of `tf.random.normal` (utils.py:90).  oracle/philox.py defines the
counter-based generator both this oracle and the CUDA engine use when a seed is given instead.

dtype: every function computes in the dtype of its inputs.  float32 restates the TF graph (float32
placeholders/variables); float64 is the "truth" used to measure how much of the 1e-4 budget fp32 itself
consumes.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .envs import EnvSpec

NUM_ELITES = 50      # utils.py:111
NUM_CEM_ITERS = 5    # utils.py:112
ALPHA = 0.1          # utils.py:113
LOWER_BOUND = -1.0   # utils.py:115
UPPER_BOUND = 1.0    # utils.py:116


# --------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------
@dataclass
class DynamicsParams:
    """Dynamics-MLP variables in `tf.trainable_variables()` order (utils.py:44-71)."""
    W: List[np.ndarray]          # n_hidden x [E, in, out]
    b: List[np.ndarray]          # n_hidden x [E, 1, out]
    W_mu: np.ndarray             # [E, H, D]
    b_mu: np.ndarray             # [E, 1, D]
    W_lv: np.ndarray             # [E, H, D]
    b_lv: np.ndarray             # [E, 1, D]
    max_logvar: np.ndarray       # [1, D]  (shared by all members, utils.py:70)
    min_logvar: np.ndarray       # [1, D]

    def astype(self, dt):
        c = lambda a: np.asarray(a, dtype=dt)
        return DynamicsParams([c(w) for w in self.W], [c(x) for x in self.b], c(self.W_mu), c(self.b_mu),
                              c(self.W_lv), c(self.b_lv), c(self.max_logvar), c(self.min_logvar))


@dataclass
class EncoderParams:
    """Context-encoder variables cp_hidden_0..2, cp_output (utils.py:594-612)."""
    W: List[np.ndarray]          # 4 x [E, in, out]
    b: List[np.ndarray]          # 4 x [E, 1, out]

    def astype(self, dt):
        return EncoderParams([np.asarray(w, dtype=dt) for w in self.W], [np.asarray(x, dtype=dt) for x in self.b])


@dataclass
class NormStats:
    """The normalisation vectors fed as placeholders (mlp_ensemble_cem_dynamics.py:354-373)."""
    obs_mean: np.ndarray         # [P]
    obs_std: np.ndarray
    act_mean: np.ndarray         # [A]
    act_std: np.ndarray
    delta_mean: np.ndarray       # [D]
    delta_std: np.ndarray
    cp_obs_mean: Optional[np.ndarray] = None   # [D*K]
    cp_obs_std: Optional[np.ndarray] = None
    cp_act_mean: Optional[np.ndarray] = None   # [A*K]
    cp_act_std: Optional[np.ndarray] = None

    def astype(self, dt):
        c = lambda a: None if a is None else np.asarray(a, dtype=dt)
        return NormStats(*[c(getattr(self, f)) for f in (
            "obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
            "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std")])


def _trunc_normal(rng, shape, std):
    """tf.truncated_normal_initializer: N(0, std) re-drawn until |x| <= 2 std (utils.py:638)."""
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return x * std


def init_dynamics_params(rng, E, in_dim, hidden, D, n_hidden=4, dtype=np.float32) -> DynamicsParams:
    """Reference initialisation: W ~ TruncN(0, 1/(2 sqrt(in))), b = 0 (utils.py:636-641),
    max_logvar = 0.5, min_logvar = -10 (utils.py:70-71)."""
    sizes = [in_dim] + [hidden] * n_hidden
    W = [_trunc_normal(rng, (E, sizes[i], sizes[i + 1]), 1.0 / (2.0 * np.sqrt(sizes[i]))) for i in range(n_hidden)]
    b = [np.zeros((E, 1, sizes[i + 1])) for i in range(n_hidden)]
    W_mu = _trunc_normal(rng, (E, hidden, D), 1.0 / (2.0 * np.sqrt(hidden)))
    W_lv = _trunc_normal(rng, (E, hidden, D), 1.0 / (2.0 * np.sqrt(hidden)))
    return DynamicsParams(W, b, W_mu, np.zeros((E, 1, D)), W_lv, np.zeros((E, 1, D)),
                          np.ones((1, D)) / 2.0, -np.ones((1, D)) * 10.0).astype(dtype)


def init_encoder_params(rng, E, in_dim, hidden=(256, 128, 64), C=10, dtype=np.float32) -> EncoderParams:
    sizes = [in_dim] + list(hidden) + [C]
    W = [_trunc_normal(rng, (E, sizes[i], sizes[i + 1]), 1.0 / (2.0 * np.sqrt(sizes[i]))) for i in range(len(sizes) - 1)]
    b = [np.zeros((E, 1, sizes[i + 1])) for i in range(len(sizes) - 1)]
    return EncoderParams(W, b).astype(dtype)


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def normalize(x, mean, std):
    return (x - mean) / (std + x.dtype.type(1e-10))            # utils.py:627-628


def denormalize(x, mean, std):
    return x * (std + x.dtype.type(1e-10)) + mean              # utils.py:631-632


def swish(x):
    return x * (x.dtype.type(1) / (x.dtype.type(1) + np.exp(-x)))   # mlp_ensemble_cem_dynamics.py:22


def relu(x):
    return np.maximum(x, x.dtype.type(0))


def softplus(x):
    # tf.nn.softplus = log(1 + exp(x)); logaddexp is the overflow-safe spelling of the same function
    return np.logaddexp(x.dtype.type(0), x)


def dense(x, W, b, act=None):
    """create_dense_layer._thunk (utils.py:643-646): act(x @ W + b), batched over the ensemble axis."""
    out = np.matmul(x, W) + b
    return act(out) if act is not None else out


def forward(x, prm: DynamicsParams, norm: NormStats, deterministic: bool, eps=None):
    """One evaluation of the ensemble dynamics MLP (utils.py:73-92 / :341-370).

    x [E, R, In] -> (out [E, R, D] in delta units, mu, logvar) ; eps [E, R, D] are the N(0,1) draws.
    """
    hdn = x
    for W, b in zip(prm.W, prm.b):
        hdn = dense(hdn, W, b, swish)
    mu = dense(hdn, prm.W_mu, prm.b_mu)
    logvar = dense(hdn, prm.W_lv, prm.b_lv)
    dmu = denormalize(mu, norm.delta_mean, norm.delta_std)
    if deterministic:
        return dmu, mu, logvar
    logvar = prm.max_logvar - softplus(prm.max_logvar - logvar)          # utils.py:84
    logvar = prm.min_logvar + softplus(logvar - prm.min_logvar)          # utils.py:85
    dlogvar = logvar + x.dtype.type(2) * np.log(norm.delta_std)          # utils.py:87 (no 1e-10 here)
    dstd = np.exp(dlogvar / x.dtype.type(2.0))                           # utils.py:88
    return dmu + eps * dstd, mu, logvar                                  # utils.py:90


def encode_context(cp_obs, cp_act, enc: EncoderParams, norm: NormStats):
    """Context encoder at inference (utils.py:400-407, 614-617): [m, D*K], [m, A*K] -> [E, m, C].
    relu hidden layers (quirk Q9), linear output; the obs block precedes the act block."""
    E = enc.W[0].shape[0]
    bo = np.tile(cp_obs[None, :, :], (E, 1, 1))
    ba = np.tile(cp_act[None, :, :], (E, 1, 1))
    x = np.concatenate([normalize(bo, norm.cp_obs_mean, norm.cp_obs_std),
                        normalize(ba, norm.cp_act_mean, norm.cp_act_std)], axis=-1)
    for W, b in zip(enc.W[:-1], enc.b[:-1]):
        x = dense(x, W, b, relu)
    return dense(x, enc.W[-1], enc.b[-1])


def predict(obs, act, prm, norm, env: EnvSpec, deterministic=True, eps=None, ctx=None):
    """One-step model evaluation in the training-graph layout (utils.py:94-99 / :372-380; the
    reference's compiled-but-unused `_get_pred`, mlp_ensemble_cem_dynamics.py:185-189).

    obs [E, B, D], act [E, B, A] (+ ctx [E, B, C]) -> (next_obs [E, B, D], mu, logvar)."""
    x = np.concatenate([normalize(env.preproc(obs), norm.obs_mean, norm.obs_std),
                        normalize(act, norm.act_mean, norm.act_std)] + ([ctx] if ctx is not None else []), axis=2)
    delta, mu, logvar = forward(x, prm, norm, deterministic, eps)
    return env.postproc(obs, delta), mu, logvar


# --------------------------------------------------------------------------------------------
# planner
# --------------------------------------------------------------------------------------------
def top_k_desc(r, k):
    """tf.nn.top_k(sorted=True): descending values, ties -> lower index first (utils.py:171)."""
    idx = np.argsort(-r, axis=-1, kind="stable")[..., :k]
    return idx.astype(np.int32)


def _to_rows(a, p, m, n, E, width):
    """[m, n, p, w] -> [E, (p/E) m n, w] exactly as utils.py:144-154 (transpose then reshape)."""
    return np.reshape(np.transpose(a, (2, 0, 1, 3)), (E, (p // E) * m * n, width))


def _from_rows(a, p, m, n, width):
    """[E, R, w] -> [m, n, p, w] as utils.py:159-160."""
    return np.transpose(np.reshape(a, (p, m, n, width)), (1, 2, 0, 3))


@dataclass
class PlanResult:
    mean: np.ndarray                 # [m, h, A]  un-clipped CEM mean (= optimal_action_var)
    var: np.ndarray                  # [m, h, A]
    action: np.ndarray               # clip(mean, -1, 1)  (get_action, mlp_ensemble...:205-206)
    returns: np.ndarray              # [iters, m, n]   particle-mean returns per candidate
    elites: np.ndarray               # [iters, m, 50]  int32
    actions: np.ndarray              # [iters, m, n, h, A]
    particle_returns: Optional[np.ndarray] = None   # [iters, m, n, p]
    states: Optional[np.ndarray] = None             # [iters, h, m, n, p, D] state AFTER each step
    extra: dict = field(default_factory=dict)


def _context_rows(ctx_raw, it_or_none, m, n, p, E, C):
    """Reproduce utils.py:433-439 literally.  `ctx_raw` is the tensor as it arrives at the top of the
    CEM-iteration body; returns (reshaped_context [E, R, C], tensor to carry into the next iteration).
    The transpose is re-applied every iteration (quirk Q3)."""
    carried = np.transpose(ctx_raw, (1, 0, 2))                                   # :434
    context = np.tile(np.reshape(carried, (m, 1, E, C)), (1, n, p // E, 1))      # :435 (quirk Q2)
    rows = _to_rows(context, p, m, n, E, C)                                      # :436-439
    return rows, carried


def rollout(obs0, actions, prm, norm, env, E, p, deterministic, eps_it=None, ctx_rows=None, trace=False):
    """The horizon loop of one CEM iteration (utils.py:137-168 / :431-472).

    obs0 [m, D]; actions [m, n, h, A]; eps_it [h, E, R, D]; ctx_rows [E, R, C] or None.
    Returns (returns [m, n, p], states [h, m, n, p, D] or None)."""
    m, n, h, A = actions.shape
    D = obs0.shape[-1]
    dt = obs0.dtype
    returns = np.zeros((m, n, p), dtype=dt)
    observation = np.tile(np.reshape(obs0, (m, 1, 1, D)), (1, n, p, 1))          # :138
    states = np.zeros((h, m, n, p, D), dtype=dt) if trace else None
    for t in range(h):
        action = actions[:, :, t]                                                 # :141
        nact = normalize(action, norm.act_mean, norm.act_std)
        nact = np.tile(nact[:, :, None, :], (1, 1, p, 1))
        nact = _to_rows(nact, p, m, n, E, A)
        pobs = env.preproc(observation)
        nobs = normalize(pobs, norm.obs_mean, norm.obs_std)
        nobs = _to_rows(nobs, p, m, n, E, nobs.shape[-1])
        parts = [nobs, nact] + ([ctx_rows] if ctx_rows is not None else [])       # :156 / :458
        x = np.concatenate(parts, axis=2)
        delta, _, _ = forward(x, prm, norm, deterministic, None if eps_it is None else eps_it[t])
        delta = _from_rows(delta, p, m, n, D)
        next_observation = env.postproc(observation, delta)                       # :162
        repeated_action = np.tile(action[:, :, None, :], (1, 1, p, 1))
        reward = env.reward(observation, repeated_action, next_observation)       # :165 (quirk Q4)
        returns = returns + reward
        observation = next_observation
        if trace:
            states[t] = observation
    return returns, states


def sample_actions(mean, var, z_it):
    """utils.py:131-135: constrained variance, then mean + sqrt(cvar) * TN(0,1,+-2)."""
    t = mean.dtype.type
    lb_dist, ub_dist = mean - t(LOWER_BOUND), t(UPPER_BOUND) - mean
    cvar = np.minimum(np.minimum(np.square(lb_dist / t(2)), np.square(ub_dist / t(2))), var)
    return mean[:, None, :, :] + np.sqrt(cvar)[:, None, :, :] * z_it, cvar


def refit(mean, var, actions, returns_mean, num_elites=NUM_ELITES, alpha=ALPHA):
    """utils.py:171-182: top-k, gather, population mean/var, EMA (the unconstrained var is carried, Q5)."""
    t = mean.dtype.type
    m = actions.shape[0]
    idx = top_k_desc(returns_mean, num_elites)                                   # [m, k]
    elites = np.stack([actions[i, idx[i]] for i in range(m)], axis=0)            # [m, k, h, A]
    new_mean = np.mean(elites, axis=1)
    new_var = np.mean(np.square(elites - new_mean[:, None, :, :]), axis=1)
    mean = mean * t(alpha) + t(1 - alpha) * new_mean
    var = var * t(alpha) + t(1 - alpha) * new_var
    return mean, var, idx


def cem_plan(obs, init_mean, init_var, z, prm, norm, env, E, p, deterministic, eps=None, ctx_raw=None,
             num_elites=NUM_ELITES, iters=NUM_CEM_ITERS, alpha=ALPHA, trace=False) -> PlanResult:
    """The whole CEM decision (utils.py:121-184 / :414-488).

    obs [m, D]; init_mean, init_var [m, h, A]; z [iters, m, n, h, A];
    eps [iters, h, E, R, D] (None when deterministic); ctx_raw [E, m, C] = encode_context(...) or None.
    """
    dt = obs.dtype
    m, D = obs.shape
    _, _, n, h, A = z.shape
    assert p % E == 0 and n >= num_elites
    mean, var = init_mean.astype(dt), init_var.astype(dt)
    rets, elites_all, acts_all, prets, states = [], [], [], [], []
    carried = ctx_raw
    for it in range(iters):
        actions, _ = sample_actions(mean, var, z[it])
        ctx_rows = None
        if carried is not None:
            C = carried.shape[-1]
            ctx_rows, carried = _context_rows(carried, it, m, n, p, E, C)
        pr, st = rollout(obs, actions, prm, norm, env, E, p, deterministic,
                         None if eps is None else eps[it], ctx_rows, trace)
        r = np.mean(pr, axis=2)                                                   # :170
        mean, var, idx = refit(mean, var, actions, r, num_elites, alpha)
        rets.append(r); elites_all.append(idx); acts_all.append(actions); prets.append(pr)
        if trace:
            states.append(st)
    t = dt.type
    return PlanResult(mean=mean, var=var, action=np.minimum(np.maximum(mean, t(-1.0)), t(1.0)),
                      returns=np.stack(rets), elites=np.stack(elites_all), actions=np.stack(acts_all),
                      particle_returns=np.stack(prets), states=np.stack(states) if trace else None)


def rs_plan(obs, u, prm, norm, env, E, p, deterministic, eps=None, ctx_raw=None, discrete=False, trace=False):
    """Random shooting (utils.py:186-246 / :490-561).

    continuous: u [m, n, h, A] are the uniform(-1, 1) draws (= the actions);
    discrete:   u [m, n, h] int32 action ids, the MLP sees the un-normalised one-hot (utils.py:199).
    Returns dict(action [m, A] or [m], returns [m, n], best [m] int32, particle_returns)."""
    dt = obs.dtype
    m, D = obs.shape
    n, h = u.shape[1], u.shape[2]
    A = env.act_dim
    if discrete:
        # one-hot goes in WITHOUT normalisation; emulate by pre-multiplying so normalize() is undone
        onehot = np.eye(A, dtype=dt)[u]                                           # [m, n, h, A]
        actions_in = onehot * (norm.act_std + dt.type(1e-10)) + norm.act_mean
        # the reward sees the integer action tiled over particles
    else:
        actions_in = u.astype(dt)
    ctx_rows = None
    if ctx_raw is not None:
        ctx_rows, _ = _context_rows(ctx_raw, 0, m, n, p, E, ctx_raw.shape[-1])   # transposed once (:512)
    if discrete:
        pr, st = _rollout_discrete(obs, u, onehot, prm, norm, env, E, p, deterministic, eps, ctx_rows, trace)
    else:
        pr, st = rollout(obs, actions_in, prm, norm, env, E, p, deterministic, eps, ctx_rows, trace)
    r = np.mean(pr, axis=2)
    best = np.argmax(r, axis=1).astype(np.int32)                                  # first max wins
    action = np.stack([u[i, best[i], 0] for i in range(m)], axis=0)
    return dict(action=action, returns=r, best=best, particle_returns=pr, states=st)


def _rollout_discrete(obs0, u, onehot, prm, norm, env, E, p, deterministic, eps_it, ctx_rows, trace):
    m, n, h = u.shape
    A = onehot.shape[-1]
    D = obs0.shape[-1]
    dt = obs0.dtype
    returns = np.zeros((m, n, p), dtype=dt)
    observation = np.tile(np.reshape(obs0, (m, 1, 1, D)), (1, n, p, 1))
    states = np.zeros((h, m, n, p, D), dtype=dt) if trace else None
    for t in range(h):
        nact = np.tile(onehot[:, :, t][:, :, None, :], (1, 1, p, 1))
        nact = _to_rows(nact, p, m, n, E, A)
        nobs = normalize(env.preproc(observation), norm.obs_mean, norm.obs_std)
        nobs = _to_rows(nobs, p, m, n, E, nobs.shape[-1])
        x = np.concatenate([nobs, nact] + ([ctx_rows] if ctx_rows is not None else []), axis=2)
        delta, _, _ = forward(x, prm, norm, deterministic, None if eps_it is None else eps_it[t])
        delta = _from_rows(delta, p, m, n, D)
        nxt = env.postproc(observation, delta)
        rep = np.tile(u[:, :, t][:, :, None], (1, 1, p))
        returns = returns + env.reward(observation, rep, nxt)
        observation = nxt
        if trace:
            states[t] = observation
    return returns, states


# --------------------------------------------------------------------------------------------
# closed-form index maps -- checked against the literal code above
# --------------------------------------------------------------------------------------------
def row_maps(m, n, p, E):
    """For every (mi, ni, pi): dynamics member e and row r such that x[e, r] is that particle's row."""
    q = p // E
    mi, ni, pi = np.meshgrid(np.arange(m), np.arange(n), np.arange(p), indexing="ij")
    e = pi // q
    r = (pi % q) * m * n + mi * n + ni
    return e, r


def context_map(ctx_raw, it, m, p, E):
    """Context vector each (mi, pi) sees in CEM iteration `it`: [m, p, C] (quirks Q2, Q3)."""
    C = ctx_raw.shape[-1]
    if it % 2 == 0:
        cx = np.transpose(ctx_raw, (1, 0, 2))                   # [m, E, C]
    else:
        cx = np.reshape(ctx_raw, (m, E, C))                     # memory reinterpretation of [E, m, C]
    pi = np.arange(p)
    return cx[:, pi % E, :]
