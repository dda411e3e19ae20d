"""Inputs of the extra reference-pinned planner cases (other environments, random shooting), drawn from a seed.

Shared by tests/golden/make_reference_golden.py (which feeds them to the reference's own graph code) and
tests/test_reference_pinned.py (which feeds them to the oracle), so the recorded file holds only the reference's outputs."""
import numpy as np

from oracle import cadm_oracle as orc
from oracle.envs import get_env


def make_case(envname, mode, E, p, n, h, H, m, context, det, seed, tweak=None):
    """Inputs in the layout of a golden fixture, drawn like tests/golden/make_golden.py draws them."""
    env = get_env(envname)
    rng = np.random.default_rng(seed)
    D, A, P = env.obs_dim, env.act_dim, env.proc_obs_dim
    C, K = (4, 3) if context else (0, 0)
    prm = orc.init_dynamics_params(rng, E, P + A + C, H, D, dtype=np.float32)
    prm.b_lv[...] = -5.0
    for b in prm.b:
        b[...] = (0.05 * rng.standard_normal(b.shape)).astype(np.float32)
    g = dict(meta=np.array([E, p, n, h, H, m, int(context), int(det), seed, C, K]), envname=np.array(envname), mode=np.array(mode))
    if context:
        enc = orc.init_encoder_params(rng, E, (D + A) * K, (32, 16, 8), C, dtype=np.float32)
        g.update({f"encW{i}": w for i, w in enumerate(enc.W)}, **{f"encb{i}": b for i, b in enumerate(enc.b)})
    f4 = np.float32
    g.update(norm_obs_mean=(0.1 * rng.standard_normal(P)).astype(f4), norm_obs_std=rng.uniform(0.5, 1.5, P).astype(f4),
             norm_act_mean=(0.05 * rng.standard_normal(A)).astype(f4), norm_act_std=rng.uniform(0.4, 0.8, A).astype(f4),
             norm_delta_mean=(0.01 * rng.standard_normal(D)).astype(f4), norm_delta_std=rng.uniform(0.05, 0.2, D).astype(f4),
             norm_cp_obs_mean=np.zeros(D * K, f4), norm_cp_obs_std=np.ones(D * K, f4),
             norm_cp_act_mean=(0.05 * rng.standard_normal(A * K)).astype(f4), norm_cp_act_std=rng.uniform(0.5, 1.0, A * K).astype(f4))
    g.update(obs=(0.1 * rng.standard_normal((m, D))).astype(f4), cp_obs=(0.1 * rng.standard_normal((m, D * K))).astype(f4),
             cp_act=(0.1 * rng.standard_normal((m, A * K))).astype(f4), mean0=(0.1 * rng.standard_normal((m, h, A))).astype(f4),
             var0=np.full((m, h, A), 0.25, f4))
    g.update({f"W{i}": w for i, w in enumerate(prm.W)}, **{f"b{i}": b for i, b in enumerate(prm.b)})
    g.update(W_mu=prm.W_mu, b_mu=prm.b_mu, W_lv=prm.W_lv, b_lv=prm.b_lv, max_logvar=prm.max_logvar, min_logvar=prm.min_logvar)
    if tweak is not None:
        tweak(g)
    return g


def _humanoid_heights(g):            # one environment inside the alive band (1, 2), one outside (slim_humanoid_env.py:104-107)
    g["obs"][0, 1], g["obs"][-1, 1] = 1.5, 0.4
    g["norm_delta_std"][1] = 0.6      # ... and steps large enough for particles to cross its edges


def _cartpole_edges(g):              # start close to the position / angle limits so that some steps cross them (:156-165)
    g["obs"][0, 0], g["obs"][-1, 2] = 2.35, -0.2
    g["norm_delta_std"][...] = 0.3


def _pendulum_angles(g):             # (cos, sin) on both sides of the branch cut of atan2 / the floored modulo (:211-212)
    g["obs"][0, :2], g["obs"][-1, :2] = (-0.99, 0.05), (-0.99, -0.05)
    g["mean0"][...] *= 8.0            # a warm start beyond the action bounds (the sampler never clips it): the reward's torque
    #                                   clip at +-max_torque becomes active (:214)


CASES = dict(
    pendulum_cem=dict(envname="pendulum", mode="cem", E=5, p=10, n=64, h=5, H=32, m=2, context=False, det=False, seed=21,
                      tweak=_pendulum_angles),
    humanoid_cem=dict(envname="slim_humanoid", mode="cem", E=5, p=5, n=64, h=3, H=32, m=2, context=False, det=False, seed=22,
                      tweak=_humanoid_heights),
    cripple_cem_det=dict(envname="cripple_halfcheetah", mode="cem", E=1, p=1, n=64, h=4, H=32, m=2, context=False, det=True, seed=23),
    ant_cadm_cem=dict(envname="ant", mode="cem", E=5, p=10, n=64, h=3, H=32, m=2, context=True, det=False, seed=24),
    cartpole_rs_discrete=dict(envname="cartpole", mode="rs_discrete", E=5, p=10, n=64, h=5, H=32, m=2, context=False, det=False,
                              seed=25, tweak=_cartpole_edges),
    hc_rs=dict(envname="halfcheetah", mode="rs", E=5, p=10, n=64, h=4, H=32, m=2, context=False, det=False, seed=26),
    hc_cadm_rs=dict(envname="halfcheetah", mode="rs", E=5, p=10, n=64, h=4, H=32, m=3, context=True, det=False, seed=27),
)


def make_train_batch(seed=31, E=3, B=7, H=24):
    """A bootstrap batch [E, B, .] for the training-graph forward passes: half-cheetah with context, plus a second set of
    dynamics weights standing for the backward model and the statistics of the backward target."""
    g = make_case("halfcheetah", "rs", E=E, p=E, n=2, h=1, H=H, m=1, context=True, det=False, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    env = get_env("halfcheetah")
    D, A, K = env.obs_dim, env.act_dim, int(g["meta"][10])
    back = orc.init_dynamics_params(rng, E, env.proc_obs_dim + A + int(g["meta"][9]), H, D, dtype=np.float32)
    for b in back.b:
        b[...] = (0.05 * rng.standard_normal(b.shape)).astype(np.float32)
    g.update({f"backW{i}": w for i, w in enumerate(back.W)}, **{f"backb{i}": b for i, b in enumerate(back.b)})
    g.update(backW_mu=back.W_mu, backb_mu=back.b_mu, backW_lv=back.W_lv, backb_lv=back.b_lv)
    g["b_lv"][0] += 9.0                      # reach both soft bounds of the log-variance
    g["b_lv"][1] -= 8.0
    g.update(bs_obs=rng.standard_normal((E, B, D)) * 0.5, bs_act=rng.uniform(-1, 1, (E, B, A)),
             bs_next=rng.standard_normal((E, B, D)) * 0.5, bs_cp_obs=rng.standard_normal((E, B, D * K)) * 0.3,
             bs_cp_act=rng.uniform(-1, 1, (E, B, A * K)),
             norm_back_delta_mean=(0.01 * rng.standard_normal(D)).astype(np.float32),
             norm_back_delta_std=rng.uniform(0.05, 0.2, D).astype(np.float32))
    return g


FIT_REPLAY = dict(E=3, H=16, K=3, F=2, C=4, batch_size=16, n=45, epochs=6, persistency=0.5, np_seed=4321,
                  valid_script=[1.0, 0.9, 5.0, 0.1, 0.1, 0.1])


def make_fit_data(seed=41):
    """Transitions for the fit-loop replay (tests/golden/make_reference_golden.py run_fit_replay and
    tests/test_reference_pinned.py): n samples of F consecutive steps with a K-step history, ragged future masks."""
    c = FIT_REPLAY
    env = get_env("halfcheetah")
    D, A, n, F, K = env.obs_dim, env.act_dim, c["n"], c["F"], c["K"]
    rng = np.random.default_rng(seed)
    obs = rng.standard_normal((n, F * D)) * 0.5
    act = rng.uniform(-1, 1, (n, F * A))
    obs_next = obs + 0.1 * rng.standard_normal((n, F * D))
    cp_obs = rng.standard_normal((n, K * D)) * 0.3
    cp_act = rng.uniform(-1, 1, (n, K * A))
    future_bool = np.ones((n, F))
    future_bool[rng.random(n) < 0.3, 1:] = 0.0          # paths that end before the second future step
    return dict(obs=obs, act=act, obs_next=obs_next, cp_obs=cp_obs, cp_act=cp_act, future_bool=future_bool)


GET_ACTION_CASES = dict(
    cadm_norm=dict(kind="cadm", envname="halfcheetah", normalize_input=True, state_diff=False),
    cadm_state_diff=dict(kind="cadm", envname="halfcheetah", normalize_input=True, state_diff=True),
    cadm_raw=dict(kind="cadm", envname="halfcheetah", normalize_input=False, state_diff=False),
    pets_norm=dict(kind="pets", envname="halfcheetah", normalize_input=True),
    pets_raw=dict(kind="pets", envname="halfcheetah", normalize_input=False),
    pets_discrete=dict(kind="pets", envname="cartpole", normalize_input=True),
)


def get_action_inputs(name, spec):
    """Statistics, observations, histories and the scripted planner answer of one get_action case (shared with the test)."""
    env = get_env(spec["envname"])
    D, A, P, K, h, m = env.obs_dim, env.act_dim, env.proc_obs_dim, 3, 4, 2
    rng = np.random.default_rng(sum(map(ord, name)))
    pair = lambda n: (rng.standard_normal(n) * 0.1, rng.uniform(0.5, 1.5, n))
    normalization = dict(obs=pair(P), delta=pair(D), act=pair(A), cp_obs=pair(D * K), cp_act=pair(A * K), back_delta=pair(D))
    discrete = spec["envname"] == "cartpole"
    return dict(D=D, A=A, P=P, K=K, h=h, m=m, discrete=discrete, normalization=normalization,
                obs=rng.standard_normal((m, D)), cp_obs=rng.standard_normal((m, D * K)), cp_act=rng.standard_normal((m, A * K)),
                init_mean=rng.standard_normal((m, h, A)) * 0.3, init_var=np.full((m, h, A), 0.25),
                cem_answer=rng.standard_normal((m, h, A)) * 2.0,               # beyond [-1, 1]: the clip must show
                rs_answer=rng.integers(0, A, size=m) if discrete else rng.standard_normal((m, A)) * 2.0)
