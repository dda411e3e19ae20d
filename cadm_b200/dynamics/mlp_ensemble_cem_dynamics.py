"""PE-TS / vanilla dynamics model with the planner on B200.

Mirror of cadm/dynamics/mlp_ensemble_cem_dynamics.py (class MLPEnsembleCEMDynamicsModel, :11): same constructor
keywords, same get_action / get_normalization_stats / save / load behaviour.
"""
from collections import OrderedDict

import numpy as np

from .core import PlannerModelBase


class MLPEnsembleCEMDynamicsModel(PlannerModelBase):
    _has_context = False

    def __init__(self, name, env, hidden_sizes=(200, 200, 200, 200), hidden_nonlinearity="swish",
                 output_nonlinearity=None, batch_size=128, learning_rate=0.001, normalize_input=True, optimizer=None,
                 valid_split_ratio=0.2, rolling_average_persitency=0.99, n_forwards=30, n_candidates=2500,
                 ensemble_size=5, n_particles=20, use_cem=False, deterministic=False, weight_decays=(0., 0., 0., 0., 0.),
                 weight_decay_coeff=0.0, **engine_kwargs):
        self.batch_size, self.learning_rate = batch_size, learning_rate
        self.valid_split_ratio, self.rolling_average_persitency = valid_split_ratio, rolling_average_persitency
        self.weight_decays, self.weight_decay_coeff = weight_decays, weight_decay_coeff
        self._init_common(name, env, hidden_sizes, hidden_nonlinearity, output_nonlinearity, normalize_input, n_forwards,
                          n_candidates, ensemble_size, n_particles, use_cem, deterministic, **engine_kwargs)

    def get_action(self, obs, cem_init_mean=None, cem_init_var=None):
        """mlp_ensemble_cem_dynamics.py:191-207: obs [m, D] (+ init mean/var [m, h, A]) -> plan [m, h, A] (CEM) or
        first action [m, A] (random shooting), clipped to [-1, 1] for continuous actions."""
        return self._plan(obs, None, None, cem_init_mean, cem_init_var)

    def predict(self, obs, act, eps=None, seed=0):
        """Additive API: one model step in the training layout.  obs [E, B, D], act [E, B, A] ->
        (next_obs, mu, logvar) as NumPy; mu / logvar are the outputs of the reference's `_get_pred` (:185-189)."""
        return tuple(t.cpu().numpy() for t in self.engine.predict(obs, act, None, eps, seed))

    def fit(self, obs, act, obs_next, epochs=1000, compute_normalization=True, valid_split_ratio=None,
            rolling_average_persitency=None, verbose=False, log_tabular=False, max_logging=5000, rng=None):
        """mlp_ensemble_cem_dynamics.py:209-323.  Training is not the hot path this package accelerates: it runs the
        reference's loss and loop through PyTorch autograd on the engine's device (cadm_b200/dynamics/training.py), then
        repacks the weights for the planner.  Returns a small dict (epochs run, last training losses)."""
        from .training import fit_ensemble
        return fit_ensemble(self, np.asarray(obs), np.asarray(act), np.asarray(obs_next), epochs=epochs,
                            valid_split_ratio=valid_split_ratio, rolling_average_persitency=rolling_average_persitency,
                            verbose=verbose, max_logging=max_logging, rng=rng)

    def compute_normalization(self, obs, act, delta):
        """mlp_ensemble_cem_dynamics.py:344-352."""
        assert obs.shape[0] == delta.shape[0] == act.shape[0]
        proc_obs = self.env.obs_preproc(obs)
        self.normalization = OrderedDict()
        self.normalization['obs'] = (np.mean(proc_obs, axis=0), np.std(proc_obs, axis=0))
        self.normalization['delta'] = (np.mean(delta, axis=0), np.std(delta, axis=0))
        self.normalization['act'] = (np.mean(act, axis=0), np.std(act, axis=0))
        self._push_norm()

    def get_normalization_stats(self):
        """mlp_ensemble_cem_dynamics.py:354-373."""
        if self.normalize_input and self.normalization is not None:
            norm_obs_mean, norm_obs_std = self.normalization['obs']
            norm_delta_mean, norm_delta_std = self.normalization['delta']
            if self.discrete:
                norm_act_mean = np.zeros((self.action_space_dims,))
                norm_act_std = np.ones((self.action_space_dims,))
            else:
                norm_act_mean, norm_act_std = self.normalization['act']
        else:
            norm_obs_mean = np.zeros((self.proc_obs_space_dims,))
            norm_obs_std = np.ones((self.proc_obs_space_dims,))
            norm_act_mean = np.zeros((self.action_space_dims,))
            norm_act_std = np.ones((self.action_space_dims,))
            norm_delta_mean = np.zeros((self.obs_space_dims,))
            norm_delta_std = np.ones((self.obs_space_dims,))
        return norm_obs_mean, norm_obs_std, norm_act_mean, norm_act_std, norm_delta_mean, norm_delta_std
