"""fit() of the PE-TS ensemble and of the CaDM model -- NOT part of the planner hot path.

The training graphs of cadm/dynamics/mlp_ensemble_cem_dynamics.py:86-170 and mlp_cadm_ensemble_cem_dynamics.py:108-317
(losses) and core/utils.py:43-97, 251-372, 569-624 (forward on the bootstrap batch) restated with PyTorch autograd: the
ensemble MLP on [E, B, .] batches, Gaussian NLL with the soft-bounded log-variance, the max/min-logvar regulariser,
per-layer L2 terms, Adam; for CaDM also the context encoder trained end to end through the forward (and backward) model,
the deterministic backward model weighted by back_coeff, and the flattening of the future_length-step samples.

Two trainers implement the step: on a CUDA device fit() runs the hand-written kernels of csrc/trainer.cu through
native_trainer.NativeTrainer (dataset resident in HBM, minibatches as index matrices, no torch on the path); the autograd
classes below are the restatement the kernels are checked against (and what a CPU-only caller gets).  The loops around them
are shared.  After training the arrays are handed back to the model, which repacks them for the engine
(cadm_plan_set_weights).
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


class TFAdam:
    """tf.train.AdamOptimizer of TensorFlow 1.x (the reference's optimiser, mlp_ensemble_cem_dynamics.py:169,
    mlp_cadm_ensemble_cem_dynamics.py:316), "epsilon hat" form:
        m <- b1 m + (1 - b1) g ;  v <- b2 v + (1 - b2) g^2 ;  theta <- theta - lr sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)
    (torch.optim.Adam puts eps inside the bias correction).  The slots m, v and the step count t belong to the MODEL: the
    reference creates the optimiser once, in the constructor, so they persist across fit() calls on the growing dataset."""

    def __init__(self, params, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps, self.t = float(lr), b1, b2, eps, 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        self.t += 1
        lr_t = self.lr * (1.0 - self.b2 ** self.t) ** 0.5 / (1.0 - self.b1 ** self.t)
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is None:
                continue
            g = p.grad
            m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            p.addcdiv_(m, v.sqrt().add_(self.eps), value=-lr_t)


def model_trainer(model, make):
    """The model's trainer, created on the first fit() and REUSED afterwards (optimiser slots and step count persist, as in
    the reference); rebuilt only when the parameters were replaced from outside (load / set_params / direct edits followed by
    _push_params), which the model counts in `_params_version`."""
    ver = getattr(model, "_params_version", 0)
    tr = getattr(model, "_trainer", None)
    if tr is None or getattr(model, "_trainer_version", None) != ver:
        tr = make()
        model._trainer = tr
    return tr


def _is_cuda(device):
    return device is not None and torch.device(device).type == "cuda"


def make_trainer(model, device):
    """The trainer of `model` on `device`: the hand-written kernels on a CUDA device, the autograd restatement otherwise."""
    cadm = getattr(model, "_enc", None) is not None
    if _is_cuda(device):
        from .native_trainer import NativeTrainer
        return NativeTrainer(model._enc if cadm else None, model._dyn, model._back if cadm else None, model.env_name,
                             model.obs_space_dims, model.proc_obs_space_dims, model.action_space_dims,
                             getattr(model, "history_length", 0), model.deterministic, model.weight_decays,
                             getattr(model, "context_weight_decays", (0.0,)), model.weight_decay_coeff,
                             getattr(model, "back_coeff", 0.0), model.learning_rate, device=device)
    if cadm:
        return CaDMTrainer(model._enc, model._dyn, model._back, model.env_name, model.deterministic, model.weight_decays,
                           model.context_weight_decays, model.weight_decay_coeff, model.back_coeff, model.learning_rate, device=device)
    return EnsembleNLLTrainer(model._dyn, model.env_name, model.deterministic, model.weight_decays, model.weight_decay_coeff,
                              model.learning_rate, device=device)


def trainer_done(model):
    """After export + repack: the trainer's tensors and the model's arrays agree again."""
    model._trainer_version = getattr(model, "_params_version", 0)


class IndexedFeed:
    """The interface the fit loops drive (shared with native_trainer.NativeTrainer): the training and validation arrays are
    handed over once per fit(), a minibatch is an index matrix [E, B].  A trainer that wants gathered batches -- the autograd
    classes below, the recording doubles of the tests -- gets them through train_step() / evaluate() exactly as the reference's
    session was fed (mlp_ensemble_cem_dynamics.py:262-283)."""

    def begin_fit(self, train, valid, stats):
        self._data, self._stats = (train, valid), stats

    def train_step_idx(self, idx):
        return self.train_step(*[a[idx] for a in self._data[0]], self._stats)

    def evaluate_idx(self, idx, which=1):
        return self.evaluate(*[a[idx] for a in self._data[which]], self._stats)


class EnsembleNLLTrainer(IndexedFeed):
    """Parameters of the dynamics ensemble as torch leaves + the reference's loss and optimiser."""

    def __init__(self, dyn, env_name, deterministic, weight_decays, weight_decay_coeff, learning_rate, device="cpu",
                 dtype=torch.float32, make_optimizer=True):
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
        if make_optimizer:
            self.optimizer = TFAdam(self.parameters(), lr=learning_rate)          # tf.train.AdamOptimizer defaults

    def parameters(self):
        return self.W + self.b + [self.W_mu, self.b_mu, self.W_lv, self.b_lv, self.max_logvar, self.min_logvar]

    def _norm(self, stats):
        return [torch.as_tensor(np.asarray(s), dtype=self.dtype, device=self.device) for s in stats]

    def forward(self, bs_obs, bs_act, stats, ctx=None):
        """mu, bounded logvar [E, B, D] of the normalised delta (core/utils.py:73-97; with a context [E, B, C] appended to
        the input, core/utils.py:365-372)."""
        om, os_, am, as_ = stats[:4]
        x = [(preproc_torch(self.env_name, bs_obs) - om) / (os_ + 1e-10), (bs_act - am) / (as_ + 1e-10)]
        x = torch.cat(x if ctx is None else x + [ctx], dim=2)
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
        l2 = self.l2()
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

    def l2(self):
        """sum of weight_decay * tf.nn.l2_loss(weight) over the hidden layers and BOTH heads (core/utils.py:642)."""
        return sum(d * 0.5 * (w ** 2).sum() for d, w in zip(self.layer_decays, self.W + [self.W_mu, self.W_lv]))

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
    trainer = model_trainer(model, lambda: make_trainer(model, device))
    trainer.begin_fit((train_obs, train_act, train_delta), (valid_obs, valid_act, valid_delta), stats)
    rolling, rolling_prev = None, None
    epoch = -1
    for epoch in range(epochs):
        mse_losses, recon_losses = [], []
        bootstrap_idx = shuffle_rows(bootstrap_idx)
        for batch_num in range(int(np.ceil(bootstrap_idx.shape[-1] / model.batch_size))):
            idx = bootstrap_idx[:, batch_num * model.batch_size:(batch_num + 1) * model.batch_size]
            m_, r_ = trainer.train_step_idx(idx)
            mse_losses.append(m_)
            recon_losses.append(r_)
        if n_valid_split > 0:
            v_mse, v_recon = trainer.evaluate_idx(valid_idx)
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
    trainer_done(model)
    return dict(epochs=epoch + 1, train_mse=float(np.mean(mse_losses)) if epoch >= 0 else None,
                train_recon=float(np.mean(recon_losses)) if epoch >= 0 else None)


# ---------------------------------------------------------------------------------------------------------- CaDM

class CaDMTrainer(IndexedFeed):
    """Context encoder + forward model (+ backward model when back_coeff > 0) with the joint loss of
    mlp_cadm_ensemble_cem_dynamics.py:266-317.  Member e of the encoder feeds member e of both models (:113-127, the
    bootstrap batch [E, B, .] goes through the encoder's batched matmul), and both models receive the SAME context
    tensor, so the backward loss also trains the encoder."""

    def __init__(self, enc, dyn, back, env_name, deterministic, weight_decays, context_weight_decays, weight_decay_coeff,
                 back_coeff, learning_rate, device="cpu", dtype=torch.float32):
        self.device, self.dtype = torch.device(device), dtype
        self.deterministic, self.back_coeff = bool(deterministic), float(back_coeff)
        self.weight_decay_coeff = float(weight_decay_coeff)
        kw = dict(device=device, dtype=dtype, make_optimizer=False)
        self.fwd = EnsembleNLLTrainer(dyn, env_name, deterministic, weight_decays, weight_decay_coeff, learning_rate, **kw)
        # the backward model is built with deterministic=True whatever the forward model is (:228)
        self.back = EnsembleNLLTrainer(back, env_name, True, weight_decays, weight_decay_coeff, learning_rate, **kw) \
            if self.back_coeff > 0.0 else None
        t = lambda a: torch.tensor(np.asarray(a), dtype=dtype, device=self.device, requires_grad=True)
        self.enc_W = [t(w) for w in enc["W"]]
        self.enc_b = [t(b) for b in enc["b"]]
        cwd = list(context_weight_decays)
        n = len(self.enc_W)                                          # hidden layers take [idx], the output layer [-1]
        self.enc_decays = [float(cwd[i]) for i in range(n - 1)] + [float(cwd[-1])]
        self.optimizer = TFAdam(self.parameters(), lr=learning_rate)

    def parameters(self):
        return self.enc_W + self.enc_b + self.fwd.parameters() + (self.back.parameters() if self.back is not None else [])

    def context(self, bs_cp_obs, bs_cp_act, stats):
        """core/utils.py:605-622: relu hidden layers, linear output, on [norm(cp_obs), norm(cp_act)]."""
        x = torch.cat([(bs_cp_obs - stats[6]) / (stats[7] + 1e-10), (bs_cp_act - stats[8]) / (stats[9] + 1e-10)], dim=-1)
        n = len(self.enc_W)
        for i, (W, b) in enumerate(zip(self.enc_W, self.enc_b)):
            x = torch.baddbmm(b, x, W)
            if i < n - 1:
                x = torch.relu(x)
        return x

    def losses(self, bs_obs, bs_act, bs_delta, bs_obs_next, bs_back_delta, bs_cp_obs, bs_cp_act, stats):
        """stats: the 12 vectors of get_normalization_stats().  Returns the scalars of :266-314."""
        f = lambda a: torch.as_tensor(np.asarray(a), dtype=self.dtype, device=self.device)
        stats = [f(s) for s in stats]
        bs_obs, bs_act, bs_delta, bs_obs_next, bs_back_delta, bs_cp_obs, bs_cp_act = map(
            f, (bs_obs, bs_act, bs_delta, bs_obs_next, bs_back_delta, bs_cp_obs, bs_cp_act))
        ctx = self.context(bs_cp_obs, bs_cp_act, stats)
        mu, logvar = self.fwd.forward(bs_obs, bs_act, stats, ctx)
        sq = (mu - (bs_delta - stats[4]) / (stats[5] + 1e-10)) ** 2
        mse = sq.mean(-1).mean(-1).sum()
        l2_fwd = self.fwd.l2()
        l2_ctx = sum(d * 0.5 * (w ** 2).sum() for d, w in zip(self.enc_decays, self.enc_W))
        l2 = l2_fwd + l2_ctx
        out = dict(mse_loss=mse, l2_reg_loss=l2_fwd, context_l2_reg_loss=l2_ctx)
        if self.back is not None:
            # input: the NEXT observation, normalised with the forward model's observation statistics (:222, :241-242);
            # target: normalize(obs - obs_next) with the back_delta statistics (:276-279)
            back_mu, _ = self.back.forward(bs_obs_next, bs_act, stats, ctx)
            back_mse = ((back_mu - (bs_back_delta - stats[10]) / (stats[11] + 1e-10)) ** 2).mean(-1).mean(-1).sum()
            out["back_l2_reg_loss"] = self.back.l2()
            l2 = l2 + out["back_l2_reg_loss"]
        else:
            back_mse = torch.zeros((), dtype=self.dtype, device=self.device)
        out["back_mse_loss"], out["l2_loss"] = back_mse, l2
        if self.deterministic:
            recon = mse
        else:
            out["mu_loss"] = (sq * torch.exp(-logvar)).mean(-1).mean(-1).sum()
            out["var_loss"] = logvar.mean(-1).mean(-1).sum()
            out["reg_loss"] = 0.01 * self.fwd.max_logvar.sum() - 0.01 * self.fwd.min_logvar.sum()
            recon = out["mu_loss"] + out["var_loss"]
        if self.back is not None:
            recon = recon + self.back_coeff * back_mse
        out["recon_loss"] = recon
        out["loss"] = recon + (0.0 if self.deterministic else out["reg_loss"]) + l2 * self.weight_decay_coeff
        return out

    def train_step(self, *batch_and_stats):
        self.optimizer.zero_grad(set_to_none=True)
        out = self.losses(*batch_and_stats)
        out["loss"].backward()
        self.optimizer.step()
        return float(out["mse_loss"].detach()), float(out["back_mse_loss"].detach()), float(out["recon_loss"].detach())

    @torch.no_grad()
    def evaluate(self, *batch_and_stats):
        out = self.losses(*batch_and_stats)
        return float(out["mse_loss"]), float(out["back_mse_loss"]), float(out["recon_loss"])

    def export(self, enc, dyn, back):
        g = lambda p: p.detach().to("cpu", torch.float32).numpy()
        for dst, src in zip(enc["W"], self.enc_W):
            dst[...] = g(src)
        for dst, src in zip(enc["b"], self.enc_b):
            dst[...] = g(src)
        self.fwd.export(dyn)
        if self.back is not None:
            self.back.export(back)


def flatten_future(D, A, K, F, obs, act, delta, cp_obs, cp_act, future_bool, obs_next, back_delta):
    """_preprocess_inputs (mlp_cadm_ensemble_cem_dynamics.py:676-696): every [n, F*dim] sample becomes F rows, the history
    is repeated for each of them, and rows whose future_bool is not positive (steps past the end of a path) are dropped."""
    keep = future_bool.reshape(-1) > 0
    rows = lambda a, d: a.reshape((-1, d))[keep, :]
    _cp_obs = np.tile(cp_obs, (1, F)).reshape((-1, D * K))[keep, :]
    _cp_act = np.tile(cp_act, (1, F)).reshape((-1, A * K))[keep, :]
    return rows(obs, D), rows(act, A), rows(delta, D), rows(obs_next, D), rows(back_delta, D), _cp_obs, _cp_act


def fit_cadm_ensemble(model, obs, act, obs_next, cp_obs, cp_act, future_bool, epochs=1000, valid_split_ratio=None,
                      rolling_average_persitency=None, verbose=False, max_logging=5000, rng=None, device=None, log=print):
    """MLPEnsembleCEMDynamicsModel.fit of the CaDM model (mlp_cadm_ensemble_cem_dynamics.py:382-569).  Differences from the
    PE-TS loop that are kept: samples carry future_length steps; the statistics come from the FIRST step of each sample
    (single_*, :408-411, :439-444); the early-stopping bound follows the rolling average every epoch (:560)."""
    rng = np.random.default_rng() if rng is None else rng
    D, A, K, F = model.obs_space_dims, model.action_space_dims, model.history_length, model.future_length
    assert obs.ndim == 2 and obs.shape[1] == D * F
    assert obs_next.ndim == 2 and obs_next.shape[1] == D * F
    assert act.ndim == 2 and act.shape[1] == A * F
    assert cp_obs.ndim == 2 and cp_obs.shape[1] == D * K
    assert cp_act.ndim == 2 and cp_act.shape[1] == A * K
    assert future_bool.ndim == 2 and future_bool.shape[1] == F
    if valid_split_ratio is None:
        valid_split_ratio = model.valid_split_ratio
    if rolling_average_persitency is None:
        rolling_average_persitency = model.rolling_average_persitency
    assert 1 > valid_split_ratio >= 0

    obs, obs_next = obs.reshape(-1, D), obs_next.reshape(-1, D)
    delta = model.env.targ_proc(obs, obs_next)
    back_delta = model.env.targ_proc(obs_next, obs)
    obs, obs_next = obs.reshape(-1, F * D), obs_next.reshape(-1, F * D)
    delta, back_delta = delta.reshape(-1, F * D), back_delta.reshape(-1, F * D)
    new = dict(obs=obs, act=act, delta=delta, cp_obs=cp_obs, cp_act=cp_act, future_bool=future_bool, obs_next=obs_next,
               back_delta=back_delta, single_obs=obs[:, :D], single_act=act[:, :A], single_delta=delta[:, :D],
               single_back_delta=back_delta[:, :D])
    if model._dataset is None:
        model._dataset = new
    else:
        for k, v in new.items():
            model._dataset[k] = np.concatenate([model._dataset[k], v])
    ds = model._dataset
    model.compute_normalization(ds['single_obs'], ds['single_act'], ds['single_delta'], ds['cp_obs'], ds['cp_act'],
                                ds['single_back_delta'])
    stats = model.get_normalization_stats()

    dataset_size = ds['obs'].shape[0]
    n_valid_split = min(int(dataset_size * valid_split_ratio), max_logging)
    permutation = rng.permutation(dataset_size)
    keys = ('obs', 'act', 'delta', 'cp_obs', 'cp_act', 'future_bool', 'obs_next', 'back_delta')
    train = flatten_future(D, A, K, F, *[ds[k][permutation[n_valid_split:]] for k in keys])
    valid = flatten_future(D, A, K, F, *[ds[k][permutation[:n_valid_split]] for k in keys])

    E = model.ensemble_size
    train_size = train[0].shape[0]
    if E > 1:
        bootstrap_idx = rng.integers(0, train_size, size=(E, train_size))
    else:
        bootstrap_idx = np.tile(np.arange(train_size, dtype='int32'), (E, 1))
    valid_idx = np.tile(np.arange(valid[0].shape[0], dtype='int32'), (E, 1))

    def shuffle_rows(arr):
        idxs = np.argsort(rng.uniform(size=arr.shape), axis=-1)
        return arr[np.arange(arr.shape[0])[:, None], idxs]

    if device is None:
        device = model.engine.device if getattr(model, "engine", None) is not None else "cpu"
    trainer = model_trainer(model, lambda: make_trainer(model, device))
    trainer.begin_fit(train, valid, stats)
    rolling, rolling_prev = None, None
    epoch = -1
    mse_losses, back_mse_losses, recon_losses = [], [], []
    for epoch in range(epochs):
        mse_losses, back_mse_losses, recon_losses = [], [], []
        bootstrap_idx = shuffle_rows(bootstrap_idx)
        for batch_num in range(int(np.ceil(bootstrap_idx.shape[-1] / model.batch_size))):
            idx = bootstrap_idx[:, batch_num * model.batch_size:(batch_num + 1) * model.batch_size]
            m_, b_, r_ = trainer.train_step_idx(idx)
            mse_losses.append(m_)
            back_mse_losses.append(b_)
            recon_losses.append(r_)
        if n_valid_split > 0:
            v_mse, v_back, v_recon = trainer.evaluate_idx(valid_idx)
            if verbose:
                log("Training DynamicsModel - finished epoch %i --[Training] mse loss: %.4f  back mse loss: %.4f  recon loss:  %.4f "
                    "[Validation] mse loss: %.4f  back mse loss: %.4f  recon loss:  %.4f"
                    % (epoch, np.mean(mse_losses), np.mean(back_mse_losses), np.mean(recon_losses), v_mse, v_back, v_recon))
            if rolling is None:                                     # :540-545
                rolling, rolling_prev = 1.5 * v_recon, 2 * v_recon
                if v_recon < 0:
                    rolling, rolling_prev = v_recon / 1.5, v_recon / 2
            rolling = rolling_average_persitency * rolling + (1.0 - rolling_average_persitency) * v_recon
            if rolling_prev < rolling:
                log('Stopping Training of Model since its valid_loss_rolling_average decreased')
                break
        elif verbose:
            log("Training DynamicsModel - finished epoch %i --[Training] mse loss: %.4f  back mse loss: %.4f  recon loss: %.4f"
                % (epoch, np.mean(mse_losses), np.mean(back_mse_losses), np.mean(recon_losses)))
        rolling_prev = rolling                                      # :560 (the PE-TS loop has no such line)
    trainer.export(model._enc, model._dyn, model._back)
    model._push_params()                                            # repack for the engine (weights + encoder)
    trainer_done(model)
    mean = lambda l: float(np.mean(l)) if l else None
    return dict(epochs=epoch + 1, train_mse=mean(mse_losses), train_back_mse=mean(back_mse_losses),
                train_recon=mean(recon_losses))
