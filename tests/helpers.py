"""Test-side glue between the product (cadm_b200) and the oracle."""
import numpy as np

from oracle import cadm_oracle as orc
from oracle.envs import get_env


def oracle_pack(model, dtype=np.float64):
    """(DynamicsParams, EncoderParams|None, NormStats, EnvSpec) of a cadm_b200 dynamics model."""
    d = model._dyn
    prm = orc.DynamicsParams(list(d["W"]), list(d["b"]), d["W_mu"], d["b_mu"], d["W_lv"], d["b_lv"],
                             d["max_logvar"], d["min_logvar"]).astype(dtype)
    enc = None
    if model._enc is not None:
        enc = orc.EncoderParams(list(model._enc["W"]), list(model._enc["b"])).astype(dtype)
    st = model.get_normalization_stats()
    # the engine receives float32 statistics; the oracle must see the same numbers
    f = lambda a: np.asarray(a, dtype=np.float32)
    norm = orc.NormStats(*[f(s) for s in st[:6]], *([f(s) for s in st[6:10]] if len(st) > 6 else [None] * 4)).astype(dtype)
    return prm, enc, norm, get_env(model.env_name)


def rel_err(got, want, axis=None):
    """max |got - want| / RMS(want) -- the 'relative to the state scale' error of SURVEY section 7."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    rms = np.sqrt(np.mean(want ** 2, axis=axis, keepdims=axis is not None))
    rms = np.maximum(rms, 1e-12)
    return float(np.max(np.abs(got - want) / rms))


def elite_margin_ok(returns_oracle, elites_oracle, returns_got, k):
    """True when the gap around every position of the oracle's sorted top-k exceeds the observed return error,
    i.e. when bit-exact elite indices are a meaningful requirement."""
    err = np.max(np.abs(np.asarray(returns_got, np.float64) - returns_oracle))
    srt = -np.sort(-returns_oracle, axis=-1)[..., : k + 1]
    gaps = srt[..., :-1] - srt[..., 1:]
    return bool(np.min(gaps) > 4 * err), float(np.min(gaps)), float(err)
