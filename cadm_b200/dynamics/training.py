"""fit() of the PE-TS ensemble -- NOT part of the planner hot path.

The training graph of cadm/dynamics/mlp_ensemble_cem_dynamics.py:86-170 (losses) and core/utils.py:43-97 (forward on the
bootstrap batch) restated with PyTorch autograd: the ensemble MLP on [E, B, .] batches, Gaussian NLL with the soft-bounded
log-variance, the max/min-logvar regulariser, per-layer L2 terms, Adam.  PyTorch is used here the way the reference uses
TensorFlow -- library GEMMs and autograd on whatever device the tensors live on (the B200 in production, the CPU in the
tests); the hand-written CUDA of this package is the planner.  After training the arrays are handed back to the model,
which repacks them for the engine (cadm_plan_set_weights).
"""
import numpy as np
import torch


def preproc_torch(env_name, obs):
    """obs_preproc of the environment (cadm/envs/*.py) on torch tensors [..., D]."""
    if env_name in ("halfcheetah", "cripple_halfcheetah"):          # half_cheetah_env.py:46-50
        return torch.cat([obs[..., 1:2], torch.sin(obs[..., 2:3]), torch.cos(obs[..., 2:3]), obs[..., 3:]], dim=-1)
    if env_name == "ant":                                           # ant_env.py:52-53
        return obs[..., 1:]
    return obs


class EnsembleNLLTrainer:
    """Parameters of the dynamics ensemble as torch leaves + the reference's loss and optimiser."""

    def __init__(self, dyn, env_name, deterministic, weight_decays, weight_decay_coeff, learning_rate, device="cpu",
                 dtype=torch.float32):
        self.env_name, self.deterministic = env_name, bool(deterministic)
        self.weight_decay_coeff = float(weight_decay_coeff)
        self.device, self.dtype = torch.device(device), dtype
        t = lambda a: torch.tensor(np.asarray(a), dtype=dtype, device=self.device, requires_grad=True)
        self.W = [t(w) for w in dyn["W"]]
        self.b = [t(b) for b in dyn["b"]]
        self.W_mu, self.b_mu = t(dyn["W_mu"]), t(dyn["b_mu"])
        self.W_lv, self.b_lv = t(dyn["W_lv"]), t(dyn["b_lv"])
        self.max_logvar, self.min_logvar = t(dyn["max_logvar"]), t(dyn["min_logvar"])
        wd = list(weight_decays)
        # create_dense_layer(weight_decay=weight_decays[idx]) for the hidden layers, weight_decays[-1] for both heads
        # (core/utils.py:46-69)
        self.layer_decays = [float(wd[min(i, len(wd) - 1)]) for i in range(len(self.W))] + [float(wd[-1]), float(wd[-1])]
        self.optimizer = torch.optim.Adam(self.parameters(), lr=learning_rate, betas=(0.9, 0.999), eps=1e-8)   # tf AdamOptimizer defaults

    def parameters(self):
        return self.W + self.b + [self.W_mu, self.b_mu, self.W_lv, self.b_lv, self.max_logvar, self.min_logvar]

    def _norm(self, stats):
        return [torch.as_tensor(np.asarray(s), dtype=self.dtype, device=self.device) for s in stats]

    def forward(self, bs_obs, bs_act, stats):
        """mu, bounded logvar [E, B, D] of the normalised delta (core/utils.py:73-97)."""
        om, os_, am, as_, dm, ds = stats
        x = torch.cat([(preproc_torch(self.env_name, bs_obs) - om) / (os_ + 1e-10), (bs_act - am) / (as_ + 1e-10)], dim=2)
        for W, b in zip(self.W, self.b):
            x = torch.baddbmm(b, x, W)
            x = x * torch.sigmoid(x)                                 # swish (mlp_ensemble_cem_dynamics.py:22)
        mu = torch.baddbmm(self.b_mu, x, self.W_mu)
        logvar = torch.baddbmm(self.b_lv, x, self.W_lv)
        if not self.deterministic:
            logvar = self.max_logvar - torch.nn.functional.softplus(self.max_logvar - logvar)
            logvar = self.min_logvar + torch.nn.functional.softplus(logvar - self.min_logvar)
        return mu, logvar

    def losses(self, bs_obs, bs_act, bs_delta, stats):
        """The scalars of mlp_ensemble_cem_dynamics.py:150-167 as a dict of torch scalars."""
        stats = self._norm(stats)
        f = lambda a: torch.as_tensor(np.asarray(a), dtype=self.dtype, device=self.device)
        bs_obs, bs_act, bs_delta = f(bs_obs), f(bs_act), f(bs_delta)
        mu, logvar = self.forward(bs_obs, bs_act, stats)
        target = (bs_delta - stats[4]) / (stats[5] + 1e-10)
        sq = (mu - target) ** 2
        mse = sq.mean(-1).mean(-1).sum()
        l2 = sum(d * 0.5 * (w ** 2).sum() for d, w in zip(self.layer_decays, self.W + [self.W_mu, self.W_lv]))   # tf.nn.l2_loss
        out = dict(mse_loss=mse, l2_reg_loss=l2)
        if self.deterministic:
            out["recon_loss"] = mse
            out["loss"] = mse + l2 * self.weight_decay_coeff
        else:
            mu_loss = (sq * torch.exp(-logvar)).mean(-1).mean(-1).sum()
            var_loss = logvar.mean(-1).mean(-1).sum()
            reg = 0.01 * self.max_logvar.sum() - 0.01 * self.min_logvar.sum()
            out.update(mu_loss=mu_loss, var_loss=var_loss, reg_loss=reg, recon_loss=mu_loss + var_loss)
            out["loss"] = out["recon_loss"] + reg + l2 * self.weight_decay_coeff
        return out

    def train_step(self, bs_obs, bs_act, bs_delta, stats):
        self.optimizer.zero_grad(set_to_none=True)
        out = self.losses(bs_obs, bs_act, bs_delta, stats)
        out["loss"].backward()
        self.optimizer.step()
        return float(out["mse_loss"].detach()), float(out["recon_loss"].detach())

    @torch.no_grad()
    def evaluate(self, bs_obs, bs_act, bs_delta, stats):
        out = self.losses(bs_obs, bs_act, bs_delta, stats)
        return float(out["mse_loss"]), float(out["recon_loss"])

    def export(self, dyn):
        """Write the trained values back into the model's arrays (in place, float32)."""
        g = lambda p: p.detach().to("cpu", torch.float32).numpy()
        for dst, src in zip(dyn["W"], self.W):
            dst[...] = g(src)
        for dst, src in zip(dyn["b"], self.b):
            dst[...] = g(src)
        for k, src in (("W_mu", self.W_mu), ("b_mu", self.b_mu), ("W_lv", self.W_lv), ("b_lv", self.b_lv),
                       ("max_logvar", self.max_logvar), ("min_logvar", self.min_logvar)):
            dyn[k][...] = g(src)


def fit_ensemble(model, obs, act, obs_next, epochs=1000, valid_split_ratio=None, rolling_average_persitency=None,
                 verbose=False, max_logging=5000, rng=None, device=None, log=print):
    """MLPEnsembleCEMDynamicsModel.fit (mlp_ensemble_cem_dynamics.py:209-323), line for line: dataset accumulation,
    normalisation, validation split, per-member bootstrap indices reshuffled every epoch, minibatches, early stopping on
    the rolling average of the validation loss.  `rng` (numpy Generator) replaces the reference's global np.random."""
    rng = np.random.default_rng() if rng is None else rng
    assert obs.ndim == 2 and obs.shape[1] == model.obs_space_dims
    assert obs_next.ndim == 2 and obs_next.shape[1] == model.obs_space_dims
    assert act.ndim == 2 and act.shape[1] == model.action_space_dims
    if valid_split_ratio is None:
        valid_split_ratio = model.valid_split_ratio
    if rolling_average_persitency is None:
        rolling_average_persitency = model.rolling_average_persitency
    assert 1 > valid_split_ratio >= 0

    delta = model.env.targ_proc(obs, obs_next)
    if model._dataset is None:
        model._dataset = dict(obs=obs, act=act, delta=delta)
    else:
        model._dataset['obs'] = np.concatenate([model._dataset['obs'], obs])
        model._dataset['act'] = np.concatenate([model._dataset['act'], act])
        model._dataset['delta'] = np.concatenate([model._dataset['delta'], delta])
    model.compute_normalization(model._dataset['obs'], model._dataset['act'], model._dataset['delta'])
    stats = model.get_normalization_stats()[:6]

    ds = model._dataset
    dataset_size = ds['obs'].shape[0]
    n_valid_split = min(int(dataset_size * valid_split_ratio), max_logging)
    permutation = rng.permutation(dataset_size)
    tr, va = permutation[n_valid_split:], permutation[:n_valid_split]
    train_obs, valid_obs = ds['obs'][tr], ds['obs'][va]
    train_act, valid_act = ds['act'][tr], ds['act'][va]
    train_delta, valid_delta = ds['delta'][tr], ds['delta'][va]

    E = model.ensemble_size
    train_size = train_obs.shape[0]
    if E > 1:
        bootstrap_idx = rng.integers(0, train_size, size=(E, train_size))
    else:
        bootstrap_idx = np.tile(np.arange(train_size, dtype='int32'), (E, 1))
    valid_idx = np.tile(np.arange(valid_obs.shape[0], dtype='int32'), (E, 1))

    def shuffle_rows(arr):
        idxs = np.argsort(rng.uniform(size=arr.shape), axis=-1)
        return arr[np.arange(arr.shape[0])[:, None], idxs]

    if device is None:
        device = model.engine.device if getattr(model, "engine", None) is not None else "cpu"
    trainer = EnsembleNLLTrainer(model._dyn, model.env_name, model.deterministic, model.weight_decays, model.weight_decay_coeff,
                                 model.learning_rate, device=device)
    rolling, rolling_prev = None, None
    epoch = -1
    for epoch in range(epochs):
        mse_losses, recon_losses = [], []
        bootstrap_idx = shuffle_rows(bootstrap_idx)
        for batch_num in range(int(np.ceil(bootstrap_idx.shape[-1] / model.batch_size))):
            idx = bootstrap_idx[:, batch_num * model.batch_size:(batch_num + 1) * model.batch_size]
            m_, r_ = trainer.train_step(train_obs[idx], train_act[idx], train_delta[idx], stats)
            mse_losses.append(m_)
            recon_losses.append(r_)
        if n_valid_split > 0:
            v_mse, v_recon = trainer.evaluate(valid_obs[valid_idx], valid_act[valid_idx], valid_delta[valid_idx], stats)
            if verbose:
                log("Training DynamicsModel - finished epoch %i --[Training] mse loss: %.4f  recon loss:  %.4f "
                    "[Validation] mse loss: %.4f  recon loss:  %.4f" % (epoch, np.mean(mse_losses), np.mean(recon_losses), v_mse, v_recon))
            if rolling is None:                                     # :300-305
                rolling, rolling_prev = 1.5 * v_recon, 2 * v_recon
                if v_recon < 0:
                    rolling, rolling_prev = v_recon / 1.5, v_recon / 2
            rolling = rolling_average_persitency * rolling + (1.0 - rolling_average_persitency) * v_recon
            if rolling_prev < rolling:
                log('Stopping Training of Model since its valid_loss_rolling_average decreased')
                break
        elif verbose:
            log("Training DynamicsModel - finished epoch %i --[Training] mse loss: %.4f  recon loss: %.4f"
                % (epoch, np.mean(mse_losses), np.mean(recon_losses)))
    trainer.export(model._dyn)
    model._push_params()                                            # repack for the engine (cadm_plan_set_weights)
    return dict(epochs=epoch + 1, train_mse=float(np.mean(mse_losses)) if epoch >= 0 else None,
                train_recon=float(np.mean(recon_losses)) if epoch >= 0 else None)
