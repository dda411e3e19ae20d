"""CaDM dynamics model (context encoder + PE-TS) with the planner on B200.

Mirror of cadm/dynamics/mlp_cadm_ensemble_cem_dynamics.py (class MLPEnsembleCEMDynamicsModel, :12): same
constructor keywords, same get_action / get_context_pred / get_normalization_stats / save / load behaviour.
The backward-dynamics model exists only for training (:212-264): its variables are created when back_coeff > 0 (after the
forward model's, as in the reference's variable list), trained by fit() and saved; the planner never reads them.  A model
built with back_coeff = 0 ignores them when load() finds them at the tail of a checkpoint.
"""
from collections import OrderedDict

import numpy as np

from .core import PlannerModelBase


class MLPEnsembleCEMDynamicsModel(PlannerModelBase):
    _has_context = True

    def __init__(self, name, env, hidden_sizes=(200, 200, 200, 200), hidden_nonlinearity="swish",
                 output_nonlinearity=None, batch_size=128, learning_rate=0.001, normalize_input=True, optimizer=None,
                 valid_split_ratio=0.2, rolling_average_persitency=0.99, n_forwards=30, n_candidates=2500,
                 ensemble_size=5, n_particles=20, use_cem=False, deterministic=False, weight_decays=(0., 0., 0., 0., 0.),
                 weight_decay_coeff=0.0, cp_hidden_sizes=(256, 128, 64), context_weight_decays=(0., 0., 0., 0.),
                 context_out_dim=10, context_hidden_nonlinearity="relu", history_length=10, future_length=10,
                 state_diff=False, back_coeff=0.0, **engine_kwargs):
        self.batch_size, self.learning_rate = batch_size, learning_rate
        self.valid_split_ratio, self.rolling_average_persitency = valid_split_ratio, rolling_average_persitency
        self.weight_decays, self.weight_decay_coeff = weight_decays, weight_decay_coeff
        self.context_weight_decays, self.future_length, self.back_coeff = context_weight_decays, future_length, back_coeff
        # context_hidden_nonlinearity is accepted and ignored, as in the reference (quirk Q9: it is never forwarded
        # to PureEnsembleContextPredictor, so relu applies; mlp_cadm_ensemble_cem_dynamics.py:141-156)
        self._init_common(name, env, hidden_sizes, hidden_nonlinearity, output_nonlinearity, normalize_input, n_forwards,
                          n_candidates, ensemble_size, n_particles, use_cem, deterministic,
                          cp_hidden_sizes=cp_hidden_sizes, context_out_dim=context_out_dim,
                          history_length=history_length, state_diff=state_diff, **engine_kwargs)

    def get_action(self, obs, cp_obs, cp_act, cem_init_mean=None, cem_init_var=None):
        """mlp_cadm_ensemble_cem_dynamics.py:344-367."""
        cp_obs = np.asarray(cp_obs, dtype=np.float32)
        cp_act = np.asarray(cp_act, dtype=np.float32)
        m = np.shape(obs)[0]
        if cp_obs.shape != (m, self.obs_space_dims * self.history_length) or \
                cp_act.shape != (m, self.action_space_dims * self.history_length):
            raise ValueError("cp_obs / cp_act must be [m, obs_dim*history_length] / [m, act_dim*history_length]")
        return self._plan(obs, cp_obs, cp_act, cem_init_mean, cem_init_var)

    def get_context_pred(self, cp_obs, cp_act):
        """mlp_cadm_ensemble_cem_dynamics.py:369-380: [m, D*K], [m, A*K] -> context [E, m, C]."""
        return self.engine.encode_context(np.asarray(cp_obs, np.float32), np.asarray(cp_act, np.float32)).cpu().numpy()

    def predict(self, obs, act, cp_obs=None, cp_act=None, ctx=None, eps=None, seed=0):
        """Additive API: one model step in the training layout [E, B, .]; the context is either given ([E, B, C]) or
        encoded from cp_obs / cp_act [B, .] with member e paired with encoder member e (training pairing, :377)."""
        if ctx is None:
            ctx = self.engine.encode_context(np.asarray(cp_obs, np.float32), np.asarray(cp_act, np.float32))
        return tuple(t.cpu().numpy() for t in self.engine.predict(obs, act, ctx, eps, seed))

    def fit(self, obs, act, obs_next, cp_obs, cp_act, future_bool, epochs=1000, compute_normalization=True,
            valid_split_ratio=None, rolling_average_persitency=None, verbose=False, log_tabular=False, max_logging=5000,
            rng=None):
        """mlp_cadm_ensemble_cem_dynamics.py:382-569: obs / act / obs_next carry future_length steps per sample
        ([n, dim*future_length]), cp_obs / cp_act the history, future_bool [n, future_length] masks steps past the end
        of a path.  Runs the reference's loss and loop through PyTorch autograd on the engine's device
        (cadm_b200/dynamics/training.py), then repacks encoder and forward model for the planner."""
        from .training import fit_cadm_ensemble
        return fit_cadm_ensemble(self, np.asarray(obs), np.asarray(act), np.asarray(obs_next), np.asarray(cp_obs),
                                 np.asarray(cp_act), np.asarray(future_bool), epochs=epochs,
                                 valid_split_ratio=valid_split_ratio, rolling_average_persitency=rolling_average_persitency,
                                 verbose=verbose, max_logging=max_logging, rng=rng)

    def compute_normalization(self, obs, act, delta, cp_obs, cp_act, back_delta):
        """mlp_cadm_ensemble_cem_dynamics.py:590-602."""
        assert obs.shape[0] == delta.shape[0] == act.shape[0]
        proc_obs = self.env.obs_preproc(obs)
        self.normalization = OrderedDict()
        self.normalization['obs'] = (np.mean(proc_obs, axis=0), np.std(proc_obs, axis=0))
        self.normalization['delta'] = (np.mean(delta, axis=0), np.std(delta, axis=0))
        self.normalization['act'] = (np.mean(act, axis=0), np.std(act, axis=0))
        self.normalization['cp_obs'] = (np.mean(cp_obs, axis=0), np.std(cp_obs, axis=0))
        self.normalization['cp_act'] = (np.mean(cp_act, axis=0), np.std(cp_act, axis=0))
        self.normalization['back_delta'] = (np.mean(back_delta, axis=0), np.std(back_delta, axis=0))
        self._push_norm()

    def get_normalization_stats(self):
        """mlp_cadm_ensemble_cem_dynamics.py:604-645 (12 vectors)."""
        D, A, K = self.obs_space_dims, self.action_space_dims, self.history_length
        if self.normalize_input and self.normalization is not None:
            n = self.normalization
            norm_obs_mean, norm_obs_std = n['obs']
            norm_delta_mean, norm_delta_std = n['delta']
            if self.discrete:
                norm_act_mean, norm_act_std = np.zeros((A,)), np.ones((A,))
            else:
                norm_act_mean, norm_act_std = n['act']
            if self.state_diff:
                norm_cp_obs_mean, norm_cp_obs_std = np.zeros((D * K,)), np.ones((D * K,))
            else:
                norm_cp_obs_mean, norm_cp_obs_std = n['cp_obs']
            if self.discrete:
                norm_cp_act_mean, norm_cp_act_std = np.zeros((A * K,)), np.ones((A * K,))
            else:
                norm_cp_act_mean, norm_cp_act_std = n['cp_act']
            norm_back_delta_mean, norm_back_delta_std = n.get('back_delta', (np.zeros((D,)), np.ones((D,))))
        else:
            norm_obs_mean, norm_obs_std = np.zeros((self.proc_obs_space_dims,)), np.ones((self.proc_obs_space_dims,))
            norm_act_mean, norm_act_std = np.zeros((A,)), np.ones((A,))
            norm_delta_mean, norm_delta_std = np.zeros((D,)), np.ones((D,))
            norm_cp_obs_mean, norm_cp_obs_std = np.zeros((D * K,)), np.ones((D * K,))
            norm_cp_act_mean, norm_cp_act_std = np.zeros((A * K,)), np.ones((A * K,))
            norm_back_delta_mean, norm_back_delta_std = np.zeros((D,)), np.ones((D,))
        return (norm_obs_mean, norm_obs_std, norm_act_mean, norm_act_std, norm_delta_mean, norm_delta_std,
                norm_cp_obs_mean, norm_cp_obs_std, norm_cp_act_mean, norm_cp_act_std, norm_back_delta_mean,
                norm_back_delta_std)
