"""TEST INFRASTRUCTURE ONLY: NumPy (fp64) restatement of the PE-TS training losses, to check cadm_b200/dynamics/training.py.

Follows cadm/dynamics/core/utils.py:73-97 (forward on the bootstrap batch, soft-bounded logvar) and
cadm/dynamics/mlp_ensemble_cem_dynamics.py:150-167 (mse / mu / var / reg / l2 losses).  Parity unpinned by the reference
(TensorFlow 1.15 cannot run here); pinned by closed-form cases and by finite differences in tests/test_training.py."""
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
