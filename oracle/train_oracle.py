"""TEST INFRASTRUCTURE ONLY: NumPy (fp64) restatement of the training losses, to check cadm_b200/dynamics/training.py.

PE-TS: cadm/dynamics/core/utils.py:73-97 (forward on the bootstrap batch, soft-bounded logvar) and
cadm/dynamics/mlp_ensemble_cem_dynamics.py:150-167 (mse / mu / var / reg / l2 losses).  CaDM: core/utils.py:605-622
(encoder), :365-372 (input with the context appended), mlp_cadm_ensemble_cem_dynamics.py:266-314 (joint loss with the
backward model) and :676-696 (flattening of the future_length-step samples, written here as explicit loops).  The forward passes (encoder,
forward model with context, backward model, PE-TS model) of cadm_b200/dynamics/training.py and every scalar loss the
reference's model constructors define (mse, backward mse, the l2 terms with their per-layer decay indices, mu / var / reg /
recon / total) are pinned by the reference's own code run over the NumPy TensorFlow stand-in: the unmodified builder
functions and the unmodified MLPEnsembleCEMDynamicsModel classes of both modules, instantiated with their placeholders served
as values (tests/golden/make_reference_golden.py run_train_forward / run_model_losses, tests/test_reference_pinned.py).  The
functions below must reproduce those recordings to 1e-12; closed-form cases and finite differences in tests/test_training.py
remain as independent checks."""
import numpy as np


def _softplus(x):
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def preproc(env_name, obs):
    if env_name in ("halfcheetah", "cripple_halfcheetah"):
        return np.concatenate([obs[..., 1:2], np.sin(obs[..., 2:3]), np.cos(obs[..., 2:3]), obs[..., 3:]], axis=-1)
    if env_name == "ant":
        return obs[..., 1:]
    return obs


def pets_losses(dyn, env_name, deterministic, weight_decays, weight_decay_coeff, bs_obs, bs_act, bs_delta, stats):
    f = np.float64
    om, os_, am, as_, dm, ds = [np.asarray(s, f) for s in stats]
    x = np.concatenate([(preproc(env_name, bs_obs.astype(f)) - om) / (os_ + 1e-10), (bs_act.astype(f) - am) / (as_ + 1e-10)], axis=2)
    for W, b in zip(dyn["W"], dyn["b"]):
        x = np.einsum("ebi,eio->ebo", x, W.astype(f)) + b.astype(f)
        x = x / (1.0 + np.exp(-x))
    mu = np.einsum("ebi,eio->ebo", x, dyn["W_mu"].astype(f)) + dyn["b_mu"].astype(f)
    lv = np.einsum("ebi,eio->ebo", x, dyn["W_lv"].astype(f)) + dyn["b_lv"].astype(f)
    mx, mn = dyn["max_logvar"].astype(f), dyn["min_logvar"].astype(f)
    if not deterministic:
        lv = mx - _softplus(mx - lv)
        lv = mn + _softplus(lv - mn)
    target = (bs_delta.astype(f) - dm) / (ds + 1e-10)
    sq = (mu - target) ** 2
    mse = sq.mean(-1).mean(-1).sum()
    wd = list(weight_decays)
    decays = [wd[min(i, len(wd) - 1)] for i in range(len(dyn["W"]))] + [wd[-1], wd[-1]]
    l2 = sum(d * 0.5 * np.sum(w.astype(f) ** 2) for d, w in zip(decays, list(dyn["W"]) + [dyn["W_mu"], dyn["W_lv"]]))
    out = dict(mse_loss=mse, l2_reg_loss=l2)
    if deterministic:
        out["recon_loss"] = mse
        out["loss"] = mse + l2 * weight_decay_coeff
    else:
        mu_loss = (sq * np.exp(-lv)).mean(-1).mean(-1).sum()
        var_loss = lv.mean(-1).mean(-1).sum()
        reg = 0.01 * mx.sum() - 0.01 * mn.sum()
        out.update(mu_loss=mu_loss, var_loss=var_loss, reg_loss=reg, recon_loss=mu_loss + var_loss)
        out["loss"] = out["recon_loss"] + reg + l2 * weight_decay_coeff
    return out


def _mlp(dyn, x, bounded):
    f = np.float64
    for W, b in zip(dyn["W"], dyn["b"]):
        x = np.einsum("ebi,eio->ebo", x, W.astype(f)) + b.astype(f)
        x = x / (1.0 + np.exp(-x))
    mu = np.einsum("ebi,eio->ebo", x, dyn["W_mu"].astype(f)) + dyn["b_mu"].astype(f)
    lv = np.einsum("ebi,eio->ebo", x, dyn["W_lv"].astype(f)) + dyn["b_lv"].astype(f)
    if bounded:
        mx, mn = dyn["max_logvar"].astype(f), dyn["min_logvar"].astype(f)
        lv = mx - _softplus(mx - lv)
        lv = mn + _softplus(lv - mn)
    return mu, lv


def _l2(dyn, weight_decays):
    wd = list(weight_decays)
    decays = [wd[i] for i in range(len(dyn["W"]))] + [wd[-1], wd[-1]]
    return sum(d * 0.5 * np.sum(w.astype(np.float64) ** 2) for d, w in zip(decays, list(dyn["W"]) + [dyn["W_mu"], dyn["W_lv"]]))


def cadm_context(enc, bs_cp_obs, bs_cp_act, stats):
    f = np.float64
    s = [np.asarray(v, f) for v in stats]
    x = np.concatenate([(bs_cp_obs.astype(f) - s[6]) / (s[7] + 1e-10), (bs_cp_act.astype(f) - s[8]) / (s[9] + 1e-10)], axis=-1)
    n = len(enc["W"])
    for i in range(n):
        x = np.einsum("ebi,eio->ebo", x, enc["W"][i].astype(f)) + enc["b"][i].astype(f)
        if i < n - 1:
            x = np.maximum(x, 0.0)
    return x


def cadm_losses(enc, dyn, back, env_name, deterministic, weight_decays, context_weight_decays, weight_decay_coeff, back_coeff,
                bs_obs, bs_act, bs_delta, bs_obs_next, bs_back_delta, bs_cp_obs, bs_cp_act, stats):
    f = np.float64
    s = [np.asarray(v, f) for v in stats]
    ctx = cadm_context(enc, bs_cp_obs, bs_cp_act, stats)
    nact = (bs_act.astype(f) - s[2]) / (s[3] + 1e-10)
    x = np.concatenate([(preproc(env_name, bs_obs.astype(f)) - s[0]) / (s[1] + 1e-10), nact, ctx], axis=2)
    mu, lv = _mlp(dyn, x, not deterministic)
    sq = (mu - (bs_delta.astype(f) - s[4]) / (s[5] + 1e-10)) ** 2
    mse = sq.mean(-1).mean(-1).sum()
    cwd = list(context_weight_decays)
    n = len(enc["W"])
    cdec = [cwd[i] for i in range(n - 1)] + [cwd[-1]]
    l2_fwd = _l2(dyn, weight_decays)
    l2_ctx = sum(d * 0.5 * np.sum(w.astype(f) ** 2) for d, w in zip(cdec, enc["W"]))
    l2 = l2_fwd + l2_ctx
    out = dict(mse_loss=mse, l2_reg_loss=l2_fwd, context_l2_reg_loss=l2_ctx)
    back_mse = 0.0
    if back_coeff > 0.0:
        xb = np.concatenate([(preproc(env_name, bs_obs_next.astype(f)) - s[0]) / (s[1] + 1e-10), nact, ctx], axis=2)
        bmu, _ = _mlp(back, xb, False)
        back_mse = ((bmu - (bs_back_delta.astype(f) - s[10]) / (s[11] + 1e-10)) ** 2).mean(-1).mean(-1).sum()
        out["back_l2_reg_loss"] = _l2(back, weight_decays)
        l2 = l2 + out["back_l2_reg_loss"]
    out["back_mse_loss"], out["l2_loss"] = back_mse, l2
    if deterministic:
        recon = mse + (back_coeff * back_mse if back_coeff > 0.0 else 0.0)
        out["recon_loss"] = recon
        out["loss"] = recon + l2 * weight_decay_coeff
    else:
        out["mu_loss"] = (sq * np.exp(-lv)).mean(-1).mean(-1).sum()
        out["var_loss"] = lv.mean(-1).mean(-1).sum()
        out["reg_loss"] = 0.01 * dyn["max_logvar"].astype(f).sum() - 0.01 * dyn["min_logvar"].astype(f).sum()
        recon = out["mu_loss"] + out["var_loss"] + (back_coeff * back_mse if back_coeff > 0.0 else 0.0)
        out["recon_loss"] = recon
        out["loss"] = recon + out["reg_loss"] + l2 * weight_decay_coeff
    return out


def flatten_future_loops(D, A, K, F, obs, act, delta, cp_obs, cp_act, future_bool, obs_next, back_delta):
    """_preprocess_inputs as loops: sample i, step j -> one row, if future_bool[i, j] > 0; the history of sample i is
    attached to each of its rows."""
    rows = [[] for _ in range(7)]
    for i in range(obs.shape[0]):
        for j in range(F):
            if not future_bool[i, j] > 0:
                continue
            rows[0].append(obs[i, j * D:(j + 1) * D])
            rows[1].append(act[i, j * A:(j + 1) * A])
            rows[2].append(delta[i, j * D:(j + 1) * D])
            rows[3].append(obs_next[i, j * D:(j + 1) * D])
            rows[4].append(back_delta[i, j * D:(j + 1) * D])
            rows[5].append(cp_obs[i])
            rows[6].append(cp_act[i])
    widths = (D, A, D, D, D, D * K, A * K)
    return tuple(np.array(r).reshape(-1, w) for r, w in zip(rows, widths))
