"""Run the reference's OWN planner graph code on the inputs of the committed golden fixtures and record what it computes.

The planner of younggyoseo/CaDM is the Python in cadm/dynamics/core/utils.py (create_plus_ensemble_cem_mlp :5-250,
create_plus_cadm_ensemble_cem_mlp :251-565, create_ensemble_pure_context_predictor :569-624, create_dense_layer :635-647)
written against TensorFlow 1.15, which cannot be installed here.  tests/golden/tf_numpy_shim.py stands in for the ~35 TF
functions that code uses, with NumPy semantics and eager evaluation; this script imports the UNMODIFIED builder functions
and environment classes from /root/reference and calls them with arrays where the model classes pass placeholders
(argument wiring as in core/layers.py:244-272, :300-345 and mlp_cadm_ensemble_cem_dynamics.py:140-210).  The builders run
the whole CEM loop -- 5 iterations x h steps of the ensemble, rewards, top-k, refit -- and return the planned mean.

For every fixture tests/golden/*.npz (inputs, weights and the oracle's float64 outputs) it injects the same weights and the
same noise (z per iteration into tf.random.truncated_normal, eps per (iteration, step) into tf.random.normal, in call
order), runs the reference in float64 and writes tests/golden/recorded/planner_reference.npz: per fixture the final mean
returned by the reference, and the candidate returns and elite indices of every iteration as seen by tf.nn.top_k.
A second file, planner_reference_cases.npz, covers what the three fixtures do not: the other environments' reward /
pre- / post-processing code (pendulum, slim humanoid, crippled half-cheetah, cart-pole, ant with context) and the
random-shooting branches (continuous, discrete one-hot, with context); their inputs are drawn from a seed by
tests/reference_cases.py, which the tests use too, so only the reference's outputs are stored.
tests/test_reference_pinned.py compares the committed fixtures and the oracle with these recordings.

Run in the build container (needs /root/reference):  python tests/golden/make_reference_golden.py
"""
import glob
import os
import sys

import numpy as np

sys.dont_write_bytecode = True                              # /root/reference is read-only: leave no __pycache__ there
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import tf_numpy_shim as shim                                # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from reference_cases import (CASES, FIT_REPLAY, GET_ACTION_CASES, get_action_inputs, make_case, make_fit_data,   # noqa: E402
                             make_train_batch)   # inputs of the extra cases, shared with the tests
from oracle import philox as ph                             # noqa: E402  (noise specification only: gen_z / gen_eps / ...)

ITERS = 5                                                   # num_cem_iters, core/utils.py:112


def reference_env(name):
    """An instance of the reference's environment class without its simulator: only obs_preproc / obs_postproc /
    tf_reward_fn are used by the planner graph, and they touch no simulator state."""
    import cadm.envs as E
    cls = dict(halfcheetah=E.HalfCheetahEnv, cripple_halfcheetah=E.CrippleHalfCheetahEnv, ant=E.AntEnv,
               slim_humanoid=E.SlimHumanoidEnv, cartpole=E.RandomCartPole_Force_Length, pendulum=E.RandomPendulumAll)[name]
    env = cls.__new__(cls)
    if name == "pendulum":
        env.max_torque = 2.0                                # gym's PendulumEnv.__init__, which is not run
    return env


def run_fixture(g, U, tf, f8=np.float64):
    """g: a mapping with the keys of a golden fixture (+ optional "mode": cem | rs | rs_discrete); f8: the float type the
    whole graph runs in."""
    E, p, n, h, H, m, context, det, seed, C, K = [int(v) for v in g["meta"]]
    mode = str(g["mode"]) if "mode" in g else "cem"
    env = reference_env(str(g["envname"]))
    D, A = g["obs"].shape[1], g["mean0"].shape[2]
    T = shim.TAPE
    T.__init__()
    T.dtype = tf.float32 = f8
    for i in range(4):
        T.variables[f"hidden_{i}_weight"], T.variables[f"hidden_{i}_bias"] = g[f"W{i}"], g[f"b{i}"]
    T.variables.update(output_mu_weight=g["W_mu"], output_mu_bias=g["b_mu"], output_logvar_weight=g["W_lv"],
                       output_logvar_bias=g["b_lv"])
    for nm in ("max_log_var", "max_logvar"):                # the two builders spell the names differently
        T.variables[nm] = g["max_logvar"].reshape(1, D)
    for nm in ("min_log_var", "min_logvar"):
        T.variables[nm] = g["min_logvar"].reshape(1, D)
    iters = ITERS if mode == "cem" else 1
    if mode == "cem":
        z = ph.gen_z(seed, ITERS, m, n, h, A).astype(f8)
        T.truncated = [z[it] for it in range(ITERS)]
    elif mode == "rs":
        T.uniform = [ph.gen_uniform_actions(seed, m, n, h, A).astype(f8)]
    else:
        T.uniform = [ph.gen_discrete_actions(seed, m, n, h, A).astype(np.int32)]
    if not det:
        eps = ph.gen_eps(seed, iters, h, m, n, p, E, D).astype(f8)
        T.normal = [None] + [eps[it, t] for it in range(iters) for t in range(h)]      # first draw: the training-batch forward
    norm = {k: g[f"norm_{k}"].astype(f8) for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
                                                     "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std")}
    swish = lambda x: x * tf.sigmoid(x)                     # mlp_ensemble_cem_dynamics.py:22
    common = dict(output_dim=D, hidden_sizes=(H,) * 4, hidden_nonlinearity=swish, output_nonlinearity=tf.identity,
                  input_obs_dim=D, input_act_dim=A, input_obs_var=g["obs"].astype(f8), input_act_var=np.zeros((m, A), f8),
                  n_forwards=h, reward_fn=env.tf_reward_fn(), n_candidates=n, norm_obs_mean_var=norm["obs_mean"],
                  norm_obs_std_var=norm["obs_std"], norm_act_mean_var=norm["act_mean"], norm_act_std_var=norm["act_std"],
                  norm_delta_mean_var=norm["delta_mean"], norm_delta_std_var=norm["delta_std"], discrete=mode == "rs_discrete",
                  ensemble_size=E, bs_input_obs_var=np.zeros((E, 1, D), f8), bs_input_act_var=np.zeros((E, 1, A), f8), n_particles=p,
                  cem_init_mean_var=g["mean0"].astype(f8) if mode == "cem" else None,         # None selects random shooting
                  cem_init_var_var=g["var0"].astype(f8),
                  obs_preproc_fn=env.obs_preproc, obs_postproc_fn=env.obs_postproc, deterministic=bool(det),
                  weight_decays=(0.,) * 5)
    out = {}
    if context:
        for i in range(3):
            T.variables[f"cp_hidden_{i}_weight"], T.variables[f"cp_hidden_{i}_bias"] = g[f"encW{i}"], g[f"encb{i}"]
        T.variables["cp_output_weight"], T.variables["cp_output_bias"] = g["encW3"], g["encb3"]
        hidden = tuple(int(g[f"encW{i}"].shape[2]) for i in range(3))
        cp_kw = dict(norm_cp_obs_mean_var=norm["cp_obs_mean"], norm_cp_obs_std_var=norm["cp_obs_std"],
                     norm_cp_act_mean_var=norm["cp_act_mean"], norm_cp_act_std_var=norm["cp_act_std"])
        bs_ctx, _, cp_forward = U.create_ensemble_pure_context_predictor(
            context_hidden_sizes=hidden, context_hidden_nonlinearity=tf.nn.relu, output_nonlinearity=tf.identity,
            ensemble_size=E, cp_input_dim=(D + A) * K, context_weight_decays=(0.,) * 4,
            bs_input_cp_obs_var=np.zeros((E, 1, D * K), f8), bs_input_cp_act_var=np.zeros((E, 1, A * K), f8), cp_output_dim=C, **cp_kw)
        res = U.create_plus_cadm_ensemble_cem_mlp(
            input_cp_obs_var=g["cp_obs"].astype(f8), input_cp_act_var=g["cp_act"].astype(f8), bs_input_cp_var=bs_ctx,
            cp_output_dim=C, cp_forward=cp_forward, build_policy_graph=True, norm_back_delta_mean_var=None,
            norm_back_delta_std_var=None, **cp_kw, **common)
        out["plan"] = np.asarray(res[3])
        # the context of every (member, environment) as the graph computes it for inference (:398-405)
        out["ctx"] = np.asarray(cp_forward(np.concatenate(
            [(np.tile(g["cp_obs"].astype(f8)[None], (E, 1, 1)) - norm["cp_obs_mean"]) / (norm["cp_obs_std"] + 1e-10),
             (np.tile(g["cp_act"].astype(f8)[None], (E, 1, 1)) - norm["cp_act_mean"]) / (norm["cp_act_std"] + 1e-10)], axis=-1),
            inference=True))
    else:
        res = U.create_plus_ensemble_cem_mlp(**common)
        out["plan"] = np.asarray(res[3])                     # CEM: final mean [m, h, A]; RS: first action of the best candidate
    assert not T.truncated and not T.normal and not T.uniform, "the graph consumed a different number of draws"
    if mode == "cem":
        assert len(T.top_k) == ITERS
        out["returns"] = np.stack([r for r, _ in T.top_k])  # [iters, m, n]: what tf.nn.top_k was handed
        out["elites"] = np.stack([i for _, i in T.top_k])   # [iters, m, 50]
    else:
        assert len(T.argmax) == 1
        out["returns"], out["best"] = T.argmax[0]           # [m, n], [m]
    out["served"] = np.array(T.served)
    return out


def run_train_forward(U, tf):
    """The forward passes of the TRAINING graph on a bootstrap batch [E, B, .] (what fit() differentiates): context encoder
    (core/utils.py:605-622), forward model with the context appended (:365-372, outputs mu and the soft-bounded logvar),
    backward model (deterministic, fed the next observation; mlp_cadm_ensemble_cem_dynamics.py:212-264), and the PE-TS
    model without context (:73-97).  The planner part of each builder is given a one-step, two-candidate problem."""
    g = make_train_batch()
    E, p, n, h, H, m, context, det, seed, C, K = [int(v) for v in g["meta"]]
    f8 = np.float64
    env = reference_env("halfcheetah")
    D, A = g["obs"].shape[1], g["mean0"].shape[2]
    T = shim.TAPE
    norm = {k: g[f"norm_{k}"].astype(f8) for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
                                                     "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std", "back_delta_mean",
                                                     "back_delta_std")}
    swish = lambda x: x * tf.sigmoid(x)

    def variables(prefix):
        T.__init__()
        T.dtype = f8
        for i in range(4):
            T.variables[f"hidden_{i}_weight"], T.variables[f"hidden_{i}_bias"] = g[f"{prefix}W{i}"], g[f"{prefix}b{i}"]
        T.variables.update(output_mu_weight=g[prefix + "W_mu"], output_mu_bias=g[prefix + "b_mu"],
                           output_logvar_weight=g[prefix + "W_lv"], output_logvar_bias=g[prefix + "b_lv"])
        for nm, key in (("max_log_var", "max_logvar"), ("max_logvar", "max_logvar"), ("min_log_var", "min_logvar"),
                        ("min_logvar", "min_logvar")):
            T.variables[nm] = g[key].reshape(1, D)
        for i in range(3):
            T.variables[f"cp_hidden_{i}_weight"], T.variables[f"cp_hidden_{i}_bias"] = g[f"encW{i}"], g[f"encb{i}"]
        T.variables["cp_output_weight"], T.variables["cp_output_bias"] = g["encW3"], g["encb3"]
        T.uniform = [ph.gen_uniform_actions(seed, m, n, h, A).astype(f8)]
        T.normal = [None] * (1 + h)                         # the batch forward and the one planning step: no noise

    def common(obs_batch, deterministic):
        return dict(output_dim=D, hidden_sizes=(H,) * 4, hidden_nonlinearity=swish, output_nonlinearity=tf.identity,
                    input_obs_dim=D, input_act_dim=A, input_obs_var=g["obs"].astype(f8), input_act_var=np.zeros((m, A)),
                    n_forwards=h, reward_fn=env.tf_reward_fn(), n_candidates=n, norm_obs_mean_var=norm["obs_mean"],
                    norm_obs_std_var=norm["obs_std"], norm_act_mean_var=norm["act_mean"], norm_act_std_var=norm["act_std"],
                    discrete=False, ensemble_size=E, bs_input_obs_var=obs_batch, bs_input_act_var=g["bs_act"].astype(f8),
                    n_particles=p, cem_init_mean_var=None, cem_init_var_var=None, obs_preproc_fn=env.obs_preproc,
                    obs_postproc_fn=env.obs_postproc, deterministic=deterministic, weight_decays=(0.,) * 5)

    cp_kw = dict(norm_cp_obs_mean_var=norm["cp_obs_mean"], norm_cp_obs_std_var=norm["cp_obs_std"],
                 norm_cp_act_mean_var=norm["cp_act_mean"], norm_cp_act_std_var=norm["cp_act_std"])
    out = {}
    variables("")
    hidden = tuple(int(g[f"encW{i}"].shape[2]) for i in range(3))
    bs_ctx, _, cp_forward = U.create_ensemble_pure_context_predictor(
        context_hidden_sizes=hidden, context_hidden_nonlinearity=tf.nn.relu, output_nonlinearity=tf.identity, ensemble_size=E,
        cp_input_dim=(D + A) * K, context_weight_decays=(0.,) * 4, bs_input_cp_obs_var=g["bs_cp_obs"].astype(f8),
        bs_input_cp_act_var=g["bs_cp_act"].astype(f8), cp_output_dim=C, **cp_kw)
    out["ctx"] = np.asarray(bs_ctx)
    res = U.create_plus_cadm_ensemble_cem_mlp(
        input_cp_obs_var=g["cp_obs"].astype(f8), input_cp_act_var=g["cp_act"].astype(f8), bs_input_cp_var=bs_ctx, cp_output_dim=C,
        cp_forward=cp_forward, build_policy_graph=True, norm_delta_mean_var=norm["delta_mean"],
        norm_delta_std_var=norm["delta_std"], norm_back_delta_mean_var=None, norm_back_delta_std_var=None, **cp_kw,
        **common(g["bs_obs"].astype(f8), False))
    out["fwd_mu"], out["fwd_logvar"] = np.asarray(res[4]), np.asarray(res[5])
    variables("back")
    res = U.create_plus_cadm_ensemble_cem_mlp(
        input_cp_obs_var=g["cp_obs"].astype(f8), input_cp_act_var=g["cp_act"].astype(f8), bs_input_cp_var=bs_ctx, cp_output_dim=C,
        cp_forward=None, build_policy_graph=False, norm_delta_mean_var=None, norm_delta_std_var=None,
        norm_back_delta_mean_var=norm["back_delta_mean"], norm_back_delta_std_var=norm["back_delta_std"], **cp_kw,
        **common(g["bs_next"].astype(f8), True))
    out["back_mu"] = np.asarray(res[4])
    # PE-TS model: same first-layer shapes are not possible (no context columns), so it gets the leading rows of W0
    variables("")
    In = int(g["W0"].shape[1]) - C
    T.variables["hidden_0_weight"] = g["W0"][:, :In]
    res = U.create_plus_ensemble_cem_mlp(norm_delta_mean_var=norm["delta_mean"], norm_delta_std_var=norm["delta_std"],
                                         **common(g["bs_obs"].astype(f8), False))
    out["pets_mu"], out["pets_logvar"] = np.asarray(res[4]), np.asarray(res[5])
    return out


class _Optimizer:
    """Stands for tf.compat.v1.train.AdamOptimizer in the model constructors (an `optimizer=` argument of theirs)."""

    def __init__(self, learning_rate):
        self.learning_rate = learning_rate

    def minimize(self, loss):
        return None


LOSS_CASES = [dict(deterministic=False, back_coeff=0.5), dict(deterministic=True, back_coeff=0.5),
              dict(deterministic=False, back_coeff=0.0)]
LOSS_DECAYS = dict(weight_decays=(1e-3, 2e-3, 3e-3, 4e-3, 5e-3), context_weight_decays=(6e-3, 7e-3, 8e-3, 9e-3),
                   weight_decay_coeff=0.7)


def run_model_losses(tf):
    """The scalar training losses as the reference's MODEL CONSTRUCTORS define them
    (mlp_ensemble_cem_dynamics.py:140-167; mlp_cadm_ensemble_cem_dynamics.py:266-314), on the bootstrap batch of
    reference_cases.make_train_batch: the unmodified classes are instantiated with every placeholder served its value in
    creation order (the graph is evaluated while it is built), an inert optimizer, non-zero weight decays, and a one-step
    random-shooting planner so that the policy part of the graph stays small.  Returns {case/loss_name: float}."""
    from cadm.dynamics.mlp_cadm_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as CaDMModel
    from cadm.dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as PETSModel
    import types
    g = make_train_batch()
    E, p, n, h, H, m, context, det, seed, C, K = [int(v) for v in g["meta"]]
    f8 = np.float64
    D, A = g["obs"].shape[1], g["mean0"].shape[2]
    env = reference_env("halfcheetah")
    env.observation_space = types.SimpleNamespace(shape=(D,))
    env.action_space = types.SimpleNamespace(shape=(A,))
    env.proc_observation_space_dims = int(g["norm_obs_mean"].shape[0])
    T = shim.TAPE
    norm = {k: g[f"norm_{k}"].astype(f8) for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
                                                     "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std", "back_delta_mean",
                                                     "back_delta_std")}
    rng = np.random.default_rng(77)
    bs_delta = rng.standard_normal((E, g["bs_obs"].shape[1], D)) * 0.1          # targets: any numbers do
    bs_back_delta = rng.standard_normal((E, g["bs_obs"].shape[1], D)) * 0.1

    def variables(C_cols):
        T.__init__()
        T.dtype = tf.float32 = f8
        for scope, pre in (("ff_model", ""), ("backward_model", "back")):
            for i in range(4):
                W = g[f"{pre}W{i}"]
                if i == 0 and C_cols == 0:
                    W = W[:, :W.shape[1] - C]                                    # the PE-TS model has no context columns
                T.variables[f"{scope}/hidden_{i}_weight"], T.variables[f"{scope}/hidden_{i}_bias"] = W, g[f"{pre}b{i}"]
            T.variables.update({f"{scope}/output_mu_weight": g[pre + "W_mu"], f"{scope}/output_mu_bias": g[pre + "b_mu"],
                                f"{scope}/output_logvar_weight": g[pre + "W_lv"], f"{scope}/output_logvar_bias": g[pre + "b_lv"]})
        for nm, key in (("max_log_var", "max_logvar"), ("max_logvar", "max_logvar"), ("min_log_var", "min_logvar"),
                        ("min_logvar", "min_logvar")):
            T.variables[nm] = g[key].reshape(1, D)
        for i in range(3):
            T.variables[f"cp_hidden_{i}_weight"], T.variables[f"cp_hidden_{i}_bias"] = g[f"encW{i}"], g[f"encb{i}"]
        T.variables["cp_output_weight"], T.variables["cp_output_bias"] = g["encW3"], g["encb3"]

    def pick(model, names):
        return {k: float(np.asarray(getattr(model, k))) for k in names if hasattr(model, k)}

    names = ("mse_loss", "back_mse_loss", "l2_reg_loss", "context_l2_reg_loss", "back_l2_reg_loss", "l2_loss", "mu_loss", "var_loss",
             "reg_loss", "recon_loss", "loss")
    hidden = tuple(int(g[f"encW{i}"].shape[2]) for i in range(3))
    common = dict(hidden_sizes=(H,) * 4, hidden_nonlinearity="swish", optimizer=_Optimizer, n_forwards=h, n_candidates=n,
                  ensemble_size=E, n_particles=p, use_cem=False, weight_decays=LOSS_DECAYS["weight_decays"],
                  weight_decay_coeff=LOSS_DECAYS["weight_decay_coeff"])
    out = {}
    for case in LOSS_CASES:
        det_, bc = case["deterministic"], case["back_coeff"]
        variables(C)
        # placeholder creation order of mlp_cadm_ensemble_cem_dynamics.py:109-137
        T.placeholders = [g["obs"], g["obs"], np.zeros((m, A)), g["cp_obs"], g["cp_act"],
                          g["bs_obs"], g["bs_next"], g["bs_act"], bs_delta, bs_back_delta, g["bs_cp_obs"], g["bs_cp_act"],
                          norm["obs_mean"], norm["obs_std"], norm["act_mean"], norm["act_std"], norm["delta_mean"], norm["delta_std"],
                          norm["cp_obs_mean"], norm["cp_obs_std"], norm["cp_act_mean"], norm["cp_act_std"],
                          norm["back_delta_mean"], norm["back_delta_std"],
                          np.zeros((m, h, A)), np.full((m, h, A), 0.25)]
        T.uniform = [ph.gen_uniform_actions(seed, m, n, h, A).astype(f8)]
        T.normal = [None] * (0 if det_ else 1 + h)
        model = CaDMModel("dm", env, cp_hidden_sizes=hidden, context_weight_decays=LOSS_DECAYS["context_weight_decays"],
                          context_out_dim=C, history_length=K, future_length=1, state_diff=True, back_coeff=bc,
                          deterministic=det_, **common)
        assert not T.placeholders and not T.uniform and not T.normal
        tag = f"cadm_{'det' if det_ else 'prob'}_back{bc}"
        for k, v in pick(model, names).items():
            out[f"{tag}/{k}"] = v
    for det_ in (False, True):
        variables(0)
        # placeholder creation order of mlp_ensemble_cem_dynamics.py:90-107
        T.placeholders = [g["obs"], np.zeros((m, A)), np.zeros((m, D)), g["bs_obs"], g["bs_act"], bs_delta,
                          norm["obs_mean"], norm["obs_std"], norm["act_mean"], norm["act_std"], norm["delta_mean"], norm["delta_std"],
                          np.zeros((m, h, A)), np.full((m, h, A), 0.25)]
        T.uniform = [ph.gen_uniform_actions(seed, m, n, h, A).astype(f8)]
        T.normal = [None] * (0 if det_ else 1 + h)
        model = PETSModel("dm", env, deterministic=det_, **common)
        assert not T.placeholders and not T.uniform and not T.normal
        for k, v in pick(model, names).items():
            out[f"pets_{'det' if det_ else 'prob'}/{k}"] = v
    out["targets/bs_delta"], out["targets/bs_back_delta"] = bs_delta, bs_back_delta
    return out


class _Session:
    """Stands for the TF session inside fit(): records what every sess.run is fed (by placeholder attribute name) and answers
    with scripted losses -- the training step's are constants, the validation recon loss follows FIT_REPLAY['valid_script']."""

    def __init__(self, model, script):
        self.names = {id(v): k[:-3] for k, v in vars(model).items() if k.endswith("_ph")}
        self.script, self.calls, self.n_valid = list(script), [], 0

    def run(self, fetches, feed_dict=None):
        train = fetches[-1] is None                      # the train_op of the inert optimizer
        self.calls.append(("train" if train else "valid", {self.names[id(k)]: np.array(v, np.float64) for k, v in feed_dict.items()}))
        n = len(fetches) - (1 if train else 0)           # mse, [back mse,] recon
        if train:
            return [0.25] * n + [None]
        self.n_valid += 1
        return [0.5] * (n - 1) + [self.script[self.n_valid - 1]]


def run_fit_replay(tf):
    """fit() of both UNMODIFIED reference model classes (mlp_ensemble_cem_dynamics.py:209-323,
    mlp_cadm_ensemble_cem_dynamics.py:382-569) on reference_cases.make_fit_data, with NumPy's global generator seeded and the
    session replaced by a recorder: everything fit() does in NumPy -- targets, dataset bookkeeping, normalisation statistics,
    validation split, flattening of the future steps, per-member bootstrap indices, per-epoch reshuffling, minibatches, the
    early-stopping rule -- is recorded as the sequence of feeds and the number of epochs run."""
    from cadm.dynamics.mlp_cadm_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as CaDMModel
    from cadm.dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as PETSModel
    import types
    c = FIT_REPLAY
    data = make_fit_data()
    f8 = np.float64
    env = reference_env("halfcheetah")
    D, A = 18, 6
    env.observation_space = types.SimpleNamespace(shape=(D,))
    env.action_space = types.SimpleNamespace(shape=(A,))
    P = env.proc_observation_space_dims = D
    E, H, K, F, C = c["E"], c["H"], c["K"], c["F"], c["C"]
    T = shim.TAPE
    rng = np.random.default_rng(5)
    out = {}

    def build(cls, context):
        T.__init__()
        T.dtype = tf.float32 = f8
        In = P + A + (C if context else 0)
        sizes = [In, H, H]
        for scope in ("ff_model", "backward_model"):
            for i in range(2):
                T.variables[f"{scope}/hidden_{i}_weight"] = rng.standard_normal((E, sizes[i], sizes[i + 1])) * 0.1
                T.variables[f"{scope}/hidden_{i}_bias"] = np.zeros((E, 1, sizes[i + 1]))
            for head in ("mu", "logvar"):
                T.variables[f"{scope}/output_{head}_weight"] = rng.standard_normal((E, H, D)) * 0.1
                T.variables[f"{scope}/output_{head}_bias"] = np.zeros((E, 1, D))
        for nm in ("max_log_var", "max_logvar"):
            T.variables[nm] = np.full((1, D), 0.5)
        for nm in ("min_log_var", "min_logvar"):
            T.variables[nm] = np.full((1, D), -10.0)
        enc_sizes = [(D + A) * K, 8, 8, 8, C]
        for i in range(3):
            T.variables[f"cp_hidden_{i}_weight"] = rng.standard_normal((E, enc_sizes[i], enc_sizes[i + 1])) * 0.1
            T.variables[f"cp_hidden_{i}_bias"] = np.zeros((E, 1, enc_sizes[i + 1]))
        T.variables["cp_output_weight"], T.variables["cp_output_bias"] = rng.standard_normal((E, 8, C)) * 0.1, np.zeros((E, 1, C))
        m, h, n, B = 1, 1, 2, 2
        z = lambda *shape: np.zeros(shape)
        o = lambda *shape: np.ones(shape)
        if context:
            T.placeholders = [z(m, D), z(m, D), z(m, A), z(m, D * K), z(m, A * K),
                              z(E, B, D), z(E, B, D), z(E, B, A), z(E, B, D), z(E, B, D), z(E, B, D * K), z(E, B, A * K),
                              z(P), o(P), z(A), o(A), z(D), o(D), z(D * K), o(D * K), z(A * K), o(A * K), z(D), o(D),
                              z(m, h, A), o(m, h, A)]
        else:
            T.placeholders = [z(m, D), z(m, A), z(m, D), z(E, B, D), z(E, B, A), z(E, B, D),
                              z(P), o(P), z(A), o(A), z(D), o(D), z(m, h, A), o(m, h, A)]
        T.uniform = [ph.gen_uniform_actions(1, m, n, h, A).astype(f8)]
        T.normal = [None] * (1 + h)
        kw = dict(hidden_sizes=(H, H), hidden_nonlinearity="swish", optimizer=_Optimizer, n_forwards=h, n_candidates=n,
                  ensemble_size=E, n_particles=E, use_cem=False, batch_size=c["batch_size"], weight_decays=(0.,) * 3)
        if context:
            kw.update(cp_hidden_sizes=(8, 8, 8), context_weight_decays=(0.,) * 4, context_out_dim=C, history_length=K,
                      future_length=F, state_diff=False, back_coeff=0.5)
        return cls("dm", env, **kw)

    for tag, cls, context in (("pets", PETSModel, False), ("cadm", CaDMModel, True)):
        model = build(cls, context)
        sess = _Session(model, c["valid_script"])
        tf.compat.v1.get_default_session = lambda sess=sess: sess
        np.random.seed(c["np_seed"])
        if context:
            model.fit(data["obs"], data["act"], data["obs_next"], data["cp_obs"], data["cp_act"], data["future_bool"],
                      epochs=c["epochs"], rolling_average_persitency=c["persistency"])
        else:                                            # the PE-TS model trains on single steps: the first of each sample
            model.fit(data["obs"][:, :D], data["act"][:, :A], data["obs_next"][:, :D], epochs=c["epochs"],
                      rolling_average_persitency=c["persistency"])
        out[f"{tag}/kinds"] = np.array([k for k, _ in sess.calls])
        for i, (_, feed) in enumerate(sess.calls):
            for k, v in feed.items():
                if not k.startswith("norm_") or i == 0:                        # the statistics do not change within one fit()
                    out[f"{tag}/call{i:03d}/{k}"] = v
        out[f"{tag}/valid_runs"] = np.array(sess.n_valid)
    return out


class _ActionSession:
    """The session inside get_action(): records the feed of the compiled planner function and answers with a scripted,
    deliberately out-of-range action."""

    def __init__(self, model, answer):
        self.names = {id(v): k[:-3] for k, v in vars(model).items() if k.endswith("_ph")}
        self.answer, self.feeds = answer, []

    def run(self, fetches, feed_dict=None):
        self.feeds.append({self.names[id(k)]: np.array(v, np.float64) for k, v in feed_dict.items()})
        return np.array(self.answer)


def run_get_action(tf):
    """get_action() of both UNMODIFIED model classes (mlp_ensemble_cem_dynamics.py:191-207,
    mlp_cadm_ensemble_cem_dynamics.py:344-367): which statistics reach the compiled planner function in which slot
    (normalize_input on / off, state_diff, discrete actions), what else is fed, and what comes back (clip for continuous
    actions only)."""
    from cadm.dynamics.mlp_cadm_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as CaDMModel
    from cadm.dynamics.mlp_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as PETSModel
    import types
    from collections import OrderedDict
    f8 = np.float64
    T = shim.TAPE
    rng = np.random.default_rng(9)
    out = {}
    for name, spec in GET_ACTION_CASES.items():
        x = get_action_inputs(name, spec)
        D, A, P, K, h, m = (x[k] for k in ("D", "A", "P", "K", "h", "m"))
        context = spec["kind"] == "cadm"
        env = reference_env(spec["envname"])
        env.observation_space = types.SimpleNamespace(shape=(D,))
        env.action_space = types.SimpleNamespace(shape=(), n=A) if x["discrete"] else types.SimpleNamespace(shape=(A,))
        env.proc_observation_space_dims = P
        E, H, C = 2, 8, 4
        for use_cem in ([False] if x["discrete"] else [True, False]):
            T.__init__()
            T.dtype = tf.float32 = f8
            In = P + A + (C if context else 0)
            for i, (a, b) in enumerate(((In, H), (H, H))):
                T.variables[f"hidden_{i}_weight"], T.variables[f"hidden_{i}_bias"] = rng.standard_normal((E, a, b)) * 0.1, np.zeros((E, 1, b))
            for head in ("mu", "logvar"):
                T.variables[f"output_{head}_weight"], T.variables[f"output_{head}_bias"] = rng.standard_normal((E, H, D)) * 0.1, np.zeros((E, 1, D))
            for nm in ("max_log_var", "max_logvar"):
                T.variables[nm] = np.full((1, D), 0.5)
            for nm in ("min_log_var", "min_logvar"):
                T.variables[nm] = np.full((1, D), -10.0)
            sizes = [(D + A) * K, 8, 8, 8, C]
            for i in range(3):
                T.variables[f"cp_hidden_{i}_weight"] = rng.standard_normal((E, sizes[i], sizes[i + 1])) * 0.1
                T.variables[f"cp_hidden_{i}_bias"] = np.zeros((E, 1, sizes[i + 1]))
            T.variables["cp_output_weight"], T.variables["cp_output_bias"] = rng.standard_normal((E, 8, C)) * 0.1, np.zeros((E, 1, C))
            z, o = (lambda *sh: np.zeros(sh)), (lambda *sh: np.ones(sh))
            n, B = 50, 2
            if context:
                T.placeholders = [z(m, D), z(m, D), z(m, A), z(m, D * K), z(m, A * K),
                                  z(E, B, D), z(E, B, D), z(E, B, A), z(E, B, D), z(E, B, D), z(E, B, D * K), z(E, B, A * K),
                                  z(P), o(P), z(A), o(A), z(D), o(D), z(D * K), o(D * K), z(A * K), o(A * K), z(D), o(D),
                                  z(m, h, A), o(m, h, A)]
            else:
                T.placeholders = [z(m, D), z(m, A), z(m, D), z(E, B, D), z(E, B, A), z(E, B, D),
                                  z(P), o(P), z(A), o(A), z(D), o(D), z(m, h, A), o(m, h, A)]
            if use_cem:
                T.truncated = [None] * ITERS
                T.normal = [None] * (1 + ITERS * h)
            else:
                T.uniform = [ph.gen_discrete_actions(1, m, n, h, A) if x["discrete"] else ph.gen_uniform_actions(1, m, n, h, A).astype(f8)]
                T.normal = [None] * (1 + h)
            kw = dict(hidden_sizes=(H, H), hidden_nonlinearity="swish", optimizer=_Optimizer, n_forwards=h, n_candidates=n,
                      ensemble_size=E, n_particles=E, use_cem=use_cem, normalize_input=spec["normalize_input"], weight_decays=(0.,) * 3)
            if context:
                kw.update(cp_hidden_sizes=(8, 8, 8), context_weight_decays=(0.,) * 4, context_out_dim=C, history_length=K,
                          future_length=1, state_diff=spec["state_diff"], back_coeff=0.0)
            model = (CaDMModel if context else PETSModel)("dm", env, **kw)
            assert not T.placeholders
            model.normalization = OrderedDict(x["normalization"]) if spec["normalize_input"] else None
            sess = _ActionSession(model, x["cem_answer"] if use_cem else x["rs_answer"])
            tf.compat.v1.get_default_session = lambda sess=sess: sess
            hist = (x["cp_obs"], x["cp_act"]) if context else ()
            warm = (x["init_mean"], x["init_var"]) if use_cem else ()
            action = model.get_action(x["obs"], *hist, *warm)
            tag = f"{name}/{'cem' if use_cem else 'rs'}"
            out[f"{tag}/action"] = np.asarray(action)
            assert len(sess.feeds) == 1
            for k, v in sess.feeds[0].items():
                out[f"{tag}/feed_{k}"] = v
    return out


class _ParamSession:
    """The session inside save(): `self.params` is the list the stand-in's trainable_variables() returned (scoped names in
    creation order); running it yields the injected values."""

    def __init__(self, values):
        self.values = values

    def run(self, fetches, feed_dict=None):
        return [np.asarray(self.values[name], np.float32) for name in fetches]


def run_checkpoint(tf, dst_dir):
    """save() of the UNMODIFIED CaDM model class (mlp_cadm_ensemble_cem_dynamics.py:571-577) with a backward model: writes
    tests/golden/recorded/reference_checkpoint.joblib (+ _norm_stats) through the reference's own code, and returns the
    value of every variable by scoped name so that the test can tell which array must land in which slot after load()."""
    from cadm.dynamics.mlp_cadm_ensemble_cem_dynamics import MLPEnsembleCEMDynamicsModel as CaDMModel
    import types
    from collections import OrderedDict
    f8 = np.float64
    T = shim.TAPE
    T.__init__()
    T.dtype = tf.float32 = f8
    env = reference_env("halfcheetah")
    D, A, P, K, C, E, H, m, h, n, B = 18, 6, 18, 3, 4, 2, 8, 1, 1, 2, 2
    env.observation_space = types.SimpleNamespace(shape=(D,))
    env.action_space = types.SimpleNamespace(shape=(A,))
    env.proc_observation_space_dims = P
    rng = np.random.default_rng(123)
    In = P + A + C
    for scope in ("ff_model", "backward_model"):
        for i, (a, b) in enumerate(((In, H), (H, H))):
            T.variables[f"{scope}/hidden_{i}_weight"] = rng.standard_normal((E, a, b))
            T.variables[f"{scope}/hidden_{i}_bias"] = rng.standard_normal((E, 1, b))
        for head in ("mu", "logvar"):
            T.variables[f"{scope}/output_{head}_weight"] = rng.standard_normal((E, H, D))
            T.variables[f"{scope}/output_{head}_bias"] = rng.standard_normal((E, 1, D))
        for nm in ("max_log_var", "max_logvar", "min_log_var", "min_logvar"):
            T.variables[f"{scope}/{nm}"] = rng.standard_normal((1, D))
    sizes = [(D + A) * K, 8, 8, 8, C]
    for i in range(3):
        T.variables[f"cp_hidden_{i}_weight"] = rng.standard_normal((E, sizes[i], sizes[i + 1]))
        T.variables[f"cp_hidden_{i}_bias"] = rng.standard_normal((E, 1, sizes[i + 1]))
    T.variables["cp_output_weight"], T.variables["cp_output_bias"] = rng.standard_normal((E, 8, C)), rng.standard_normal((E, 1, C))
    z, o = (lambda *sh: np.zeros(sh)), (lambda *sh: np.ones(sh))
    T.placeholders = [z(m, D), z(m, D), z(m, A), z(m, D * K), z(m, A * K),
                      z(E, B, D), z(E, B, D), z(E, B, A), z(E, B, D), z(E, B, D), z(E, B, D * K), z(E, B, A * K),
                      z(P), o(P), z(A), o(A), z(D), o(D), z(D * K), o(D * K), z(A * K), o(A * K), z(D), o(D), z(m, h, A), o(m, h, A)]
    T.uniform = [ph.gen_uniform_actions(1, m, n, h, A).astype(f8)]
    T.normal = [None] * (1 + h)
    model = CaDMModel("dm", env, hidden_sizes=(H, H), hidden_nonlinearity="swish", optimizer=_Optimizer, n_forwards=h,
                      n_candidates=n, ensemble_size=E, n_particles=E, use_cem=False, weight_decays=(0.,) * 3,
                      cp_hidden_sizes=(8, 8, 8), context_weight_decays=(0.,) * 4, context_out_dim=C, history_length=K,
                      future_length=1, state_diff=False, back_coeff=0.5)
    # value of every variable under the scoped name the constructor created it with
    values = {}
    for scoped in model.params:
        parts = scoped.split("/")
        keys = [f"{sc}/{parts[-1]}" for sc in reversed(parts[:-1]) if f"{sc}/{parts[-1]}" in T.variables] + [parts[-1]]
        values[scoped] = T.variables[keys[0]]
    pair = lambda k: (rng.standard_normal(k), rng.uniform(0.5, 1.5, k))
    model.normalization = OrderedDict(obs=pair(P), delta=pair(D), act=pair(A), cp_obs=pair(D * K), cp_act=pair(A * K),
                                      back_delta=pair(D))
    tf.compat.v1.get_default_session = lambda: _ParamSession(values)
    path = os.path.join(dst_dir, "reference_checkpoint.joblib")
    model.save(path)
    out = {f"var{idx:02d}:{scoped}": np.asarray(values[scoped], np.float32) for idx, scoped in enumerate(model.params)}
    for k, (mu, sd) in model.normalization.items():
        out[f"norm/{k}/mean"], out[f"norm/{k}/std"] = mu, sd
    return out


def main():
    tf = shim.install(np.float64)
    sys.path.insert(0, "/root/reference")
    from cadm.dynamics.core import utils as U
    blob = {}
    for path in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        name = os.path.splitext(os.path.basename(path))[0]
        g = np.load(path)
        out = run_fixture(g, U, tf)
        scale = np.max(np.abs(g["out_returns"]))
        print(f"{name}: reference vs committed fixture: returns {np.max(np.abs(out['returns'] - g['out_returns'])) / scale:.2e} "
              f"(relative), mean {np.max(np.abs(out['plan'] - g['out_mean'])):.2e}, elites equal: "
              f"{np.array_equal(out['elites'], g['out_elites'])}")
        for k, v in out.items():
            blob[f"{name}/{k}"] = v
        # the same graph in float32 (NumPy's float32 kernels, not TensorFlow's): pins the oracle's float32 mode and shows
        # what single precision costs the reference's own formulation
        o32 = run_fixture(g, U, tf, np.float32)
        assert o32["returns"].dtype == np.float32 and o32["plan"].dtype == np.float32
        print(f"{name}: float32 graph vs float64 graph: returns {np.max(np.abs(o32['returns'] - out['returns'])) / scale:.2e} "
              f"(relative), mean {np.max(np.abs(o32['plan'] - out['plan'])):.2e}, elites equal: "
              f"{np.array_equal(o32['elites'], out['elites'])}")
        for k in ("plan", "returns", "elites"):
            blob[f"{name}/f32_{k}"] = o32[k]
    dst = os.path.join(HERE, "recorded", "planner_reference.npz")
    np.savez_compressed(dst, **blob)
    print(dst, os.path.getsize(dst), "bytes")
    blob = {}
    for name, spec in CASES.items():
        g = make_case(**spec)
        out = run_fixture(g, U, tf)
        print(f"{name}: plan {out['plan'].shape}, returns {out['returns'].shape}, spread of returns "
              f"{float(out['returns'].min()):.3f} .. {float(out['returns'].max()):.3f}")
        for k, v in out.items():
            blob[f"{name}/{k}"] = v
    for k, v in run_train_forward(U, tf).items():
        blob[f"train_forward/{k}"] = v
    print("train_forward:", {k: v.shape for k, v in blob.items() if k.startswith("train_forward/")})
    for k, v in run_checkpoint(tf, os.path.join(HERE, "recorded")).items():
        blob[f"checkpoint/{k}"] = v
    print("checkpoint:", [k.split(":", 1)[1] for k in sorted(blob) if k.startswith("checkpoint/var")][:4], "...")
    for k, v in run_get_action(tf).items():
        blob[f"get_action/{k}"] = v
    print("get_action:", sorted({k.split("/feed_")[0] for k in blob if k.startswith("get_action/") and "/feed_" in k}))
    replay = run_fit_replay(tf)
    for k, v in replay.items():
        blob[f"fit_replay/{k}"] = v
    print("fit_replay:", {t: (list(replay[f"{t}/kinds"]).count("train"), int(replay[f"{t}/valid_runs"])) for t in ("pets", "cadm")},
          "(training steps, epochs)")
    losses = run_model_losses(tf)
    for k, v in losses.items():
        blob[f"model_losses/{k}"] = np.asarray(v)
    print("model_losses:", {k: round(float(v), 6) for k, v in losses.items() if np.ndim(v) == 0})
    dst = os.path.join(HERE, "recorded", "planner_reference_cases.npz")
    np.savez_compressed(dst, **blob)
    print(dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
