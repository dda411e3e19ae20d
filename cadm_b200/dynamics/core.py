"""Shared implementation of the two dynamics-model classes.

The reference builds its whole TF graph in the constructor (mlp_ensemble_cem_dynamics.py:29-189,
mlp_cadm_ensemble_cem_dynamics.py:26-342); here the constructor allocates the variables (NumPy, reference
initialisation) and a PlannerEngine.  Only the planning surface is implemented:
get_action / get_context_pred / get_normalization_stats / save / load, plus the additive predict().  `fit()` is the
widening row of SURVEY.md section 8(f): it lives in training.py (PyTorch autograd) and hands the arrays back here.
"""
from collections import OrderedDict

import joblib
import numpy as np

from ..engine import PlannerConfig, PlannerEngine
from ..envs import resolve_env

_ACTIVATIONS = (None, "relu", "tanh", "sigmoid", "softmax", "swish")


def _trunc_normal(rng, shape, std):
    # tf.truncated_normal_initializer(stddev): resample beyond 2 std (core/utils.py:636-638)
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return (x * std).astype(np.float32)


class PlannerModelBase:
    _has_context = False

    def _init_common(self, name, env, hidden_sizes, hidden_nonlinearity, output_nonlinearity, normalize_input,
                     n_forwards, n_candidates, ensemble_size, n_particles, use_cem, deterministic,
                     cp_hidden_sizes=(256, 128, 64), context_out_dim=10, history_length=10, state_diff=False,
                     seed=0, m_max=32, precision="tc3x", device=None, rank=0, world=1, context_layout="reference",
                     num_elites=50, cem_iters=5, alpha=0.1):
        if hidden_nonlinearity not in _ACTIVATIONS or output_nonlinearity not in _ACTIVATIONS:
            raise KeyError(hidden_nonlinearity)                      # the reference indexes _activations[...]
        if hidden_nonlinearity != "swish" or output_nonlinearity is not None:
            raise NotImplementedError(
                "the fused planner implements the configuration every reference run script uses "
                "(hidden_nonlinearity='swish', output_nonlinearity=None; run_pets.py:189-190)")
        hidden_sizes = tuple(hidden_sizes)
        if len(set(hidden_sizes)) != 1:
            raise NotImplementedError("hidden layers must share one width (reference default (200,)*4)")
        self.name = name
        self.env, self.env_name = resolve_env(env)
        self._dataset = None
        self.deterministic = deterministic
        self.n_forwards = n_forwards
        self.n_candidates = n_candidates
        self.use_cem = use_cem
        self.normalization = None
        self.normalize_input = normalize_input
        self.ensemble_size = ensemble_size
        self.n_particles = n_particles
        self.hidden_sizes = hidden_sizes
        self.obs_space_dims = D = self.env.observation_space.shape[0]
        self.proc_obs_space_dims = P = self.env.proc_observation_space_dims
        if len(self.env.action_space.shape) == 0:
            self.action_space_dims = A = self.env.action_space.n
            self.discrete = True
        else:
            self.action_space_dims = A = self.env.action_space.shape[0]
            self.discrete = False
        if self.discrete and use_cem:
            raise NotImplementedError("CEM needs continuous actions")
        self.cp_hidden_sizes = tuple(cp_hidden_sizes)
        self.context_out_dim = context_out_dim if self._has_context else 0
        self.history_length = history_length
        self.state_diff = state_diff
        self._seed = int(seed)
        self._calls = 0

        # ---- variables, in tf.trainable_variables() creation order
        rng = np.random.default_rng(seed)
        E, H = ensemble_size, hidden_sizes[0]
        self._enc = None
        self._back = None                                            # backward model, only built when back_coeff > 0
        if self._has_context:
            sizes = [(D + A) * history_length] + list(self.cp_hidden_sizes) + [context_out_dim]
            self._enc = dict(
                W=[_trunc_normal(rng, (E, sizes[i], sizes[i + 1]), 1 / (2 * np.sqrt(sizes[i]))) for i in range(len(sizes) - 1)],
                b=[np.zeros((E, 1, sizes[i + 1]), np.float32) for i in range(len(sizes) - 1)])
        In = P + A + self.context_out_dim
        sizes = [In] + list(hidden_sizes)
        self._dyn = self._new_mlp(rng, E, sizes, D)
        if self._has_context and getattr(self, "back_coeff", 0.0) > 0.0:
            # 'backward_model' scope (mlp_cadm_ensemble_cem_dynamics.py:212-264): same shapes, created after the forward model
            self._back = self._new_mlp(rng, E, sizes, D)

        cfg = PlannerConfig(env=self.env_name, obs_dim=D, proc_obs_dim=P, act_dim=A, ctx_dim=self.context_out_dim,
                            hist_len=history_length, hidden=H, n_hidden=len(hidden_sizes),
                            enc_hidden=tuple(self.cp_hidden_sizes), ensemble=E, particles=n_particles,
                            candidates=n_candidates, horizon=n_forwards, m_max=m_max, deterministic=deterministic,
                            discrete=self.discrete, precision=precision, rank=rank, world=world,
                            num_elites=num_elites, cem_iters=cem_iters, alpha=alpha,
                            context_layout=context_layout, max_torque=getattr(self.env, "max_torque", 2.0))
        self.engine = PlannerEngine(cfg, device=device)
        self._push_params()
        self._push_norm()

    @staticmethod
    def _new_mlp(rng, E, sizes, D):
        """Variables of one create_*_ensemble_cem_mlp call (core/utils.py:306-336): hidden layers, mu / logvar heads,
        max / min logvar."""
        H, n = sizes[-1], len(sizes) - 1
        return dict(
            W=[_trunc_normal(rng, (E, sizes[i], sizes[i + 1]), 1 / (2 * np.sqrt(sizes[i]))) for i in range(n)],
            b=[np.zeros((E, 1, sizes[i + 1]), np.float32) for i in range(n)],
            W_mu=_trunc_normal(rng, (E, H, D), 1 / (2 * np.sqrt(H))), b_mu=np.zeros((E, 1, D), np.float32),
            W_lv=_trunc_normal(rng, (E, H, D), 1 / (2 * np.sqrt(H))), b_lv=np.zeros((E, 1, D), np.float32),
            max_logvar=np.ones((1, D), np.float32) / 2.0, min_logvar=-np.ones((1, D), np.float32) * 10.0)

    # ------------------------------------------------------------------ parameters
    @property
    def params(self):
        """List of arrays in the reference's trainable_variables() order (what save() dumps)."""
        out = []
        if self._enc is not None:
            for W, b in zip(self._enc["W"], self._enc["b"]):
                out += [W, b]
        for d in (self._dyn, self._back):
            if d is None:
                continue
            for W, b in zip(d["W"], d["b"]):
                out += [W, b]
            out += [d["W_mu"], d["b_mu"], d["W_lv"], d["b_lv"], d["max_logvar"], d["min_logvar"]]
        return out

    def set_params(self, arrays):
        arrays = [np.asarray(a, dtype=np.float32) for a in arrays]
        cur = self.params
        if len(arrays) < len(cur):
            raise ValueError(f"expected at least {len(cur)} arrays, got {len(arrays)}")
        for i, (dst, src) in enumerate(zip(cur, arrays)):       # extra arrays (a backward model this one lacks) are ignored
            if dst.shape != src.shape:
                raise ValueError(f"variable {i}: shape {src.shape} does not match {dst.shape}")
            dst[...] = src
        self._push_params()

    def _push_params(self):
        # every repack counts as "the parameters may have changed": a trainer kept from an earlier fit() is rebuilt from the
        # arrays unless fit() itself did the repack (training.trainer_done)
        self._params_version = getattr(self, "_params_version", 0) + 1
        d = self._dyn
        self.engine.set_weights(d["W"] + [d["W_mu"], d["W_lv"]], d["b"] + [d["b_mu"], d["b_lv"]],
                                d["max_logvar"], d["min_logvar"])
        if self._enc is not None:
            self.engine.set_encoder(self._enc["W"], self._enc["b"])

    def _push_norm(self):
        self.engine.set_norm(*self.get_normalization_stats()[:10] if self._has_context else
                             self.get_normalization_stats()[:6])

    def set_normalization(self, normalization):
        self.normalization = normalization
        self._push_norm()

    # ------------------------------------------------------------------ persistence (joblib, reference layout)
    def save(self, save_path):
        joblib.dump([np.array(p) for p in self.params], save_path)
        if self.normalization is not None:
            joblib.dump(self.normalization, save_path + "_norm_stats")

    def load(self, load_path):
        self.set_params(joblib.load(load_path))
        if self.normalize_input:
            self.normalization = joblib.load(load_path + "_norm_stats")
        self._push_norm()

    # ------------------------------------------------------------------ planning
    def _next_seed(self):
        s = ((self._seed & 0xFFFFFFFF) << 32) | (self._calls & 0xFFFFFFFF)
        self._calls += 1
        return s

    def _plan(self, obs, cp_obs, cp_act, cem_init_mean, cem_init_var):
        obs = np.asarray(obs, dtype=np.float32)
        if obs.ndim != 2 or obs.shape[1] != self.obs_space_dims:
            raise ValueError(f"obs must be [m, {self.obs_space_dims}], got {obs.shape}")
        if cem_init_mean is not None:
            if not self.use_cem:
                # the reference builds the RS graph when use_cem=False; extra feeds are ignored
                return self._rs(obs, cp_obs, cp_act)
            shp = (obs.shape[0], self.n_forwards, self.action_space_dims)
            if tuple(np.shape(cem_init_mean)) != shp or tuple(np.shape(cem_init_var)) != shp:
                raise ValueError(f"cem_init_mean / cem_init_var must be {shp}")
            if getattr(getattr(self.engine, "cfg", None), "world", 1) > 1:
                return self._plan_sharded(obs, cp_obs, cp_act, cem_init_mean, cem_init_var)
            return self.engine.plan_cem_host(obs, cem_init_mean, cem_init_var, cp_obs, cp_act, seed=self._next_seed())
        if self.use_cem:
            raise ValueError("model was built with use_cem=True: cem_init_mean / cem_init_var must be fed")
        return self._rs(obs, cp_obs, cp_act)

    # ------------------------------------------------------------------ planning over the GPUs of one box
    def sharded_planner(self, group=None, fused=None):
        """The candidate-sharded planner of this model's engine (world > 1; created on first use, then kept).  Every rank
        builds the model with its own `rank` / the common `world` and the same `seed`, and calls get_action() with the
        same inputs: the decision is SPMD and every rank returns the same plan (cadm_b200/parallel.py)."""
        if getattr(self, "_sharded", None) is None:
            from ..parallel import ShardedCEMPlanner
            self._sharded = ShardedCEMPlanner(self.engine, group=group, fused=fused)
        return self._sharded

    def _plan_sharded(self, obs, cp_obs, cp_act, cem_init_mean, cem_init_var):
        planner = self.sharded_planner()
        seed = self._next_seed()
        if planner.fused:
            # the exchange happens inside the library (peer memory): the single host call works at world > 1 as it is
            return self.engine.plan_cem_host(obs, cem_init_mean, cem_init_var, cp_obs, cp_act, seed=seed)
        out = planner.plan(obs, cem_init_mean, cem_init_var, cp_obs, cp_act, seed=seed, logs=False)
        mean = out["mean"]
        mean = mean.detach().cpu().numpy() if hasattr(mean, "detach") else np.asarray(mean)
        return np.minimum(np.maximum(mean.astype(np.float32), -1.0), 1.0)      # get_action's clip (reference :205-206)

    def _rs(self, obs, cp_obs, cp_act):
        out = self.engine.plan_rs(obs, cp_obs, cp_act, seed=self._next_seed())
        action = out["action"].cpu().numpy()
        if not self.discrete:
            action = np.minimum(np.maximum(action, -1.0), 1.0)
        return action
