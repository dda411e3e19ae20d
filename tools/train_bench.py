"""Time of one training step (minibatch of 128 rows per ensemble member, the reference's batch size): the hand-written kernels of
csrc/trainer.cu against the autograd restatement on the same GPU (library GEMMs, what round 1 shipped), and the launches per
step.  C2 = PE-TS HalfCheetah (E = 5, 4 x 200), C3 = CaDM HalfCheetah (encoder 256-128-64 -> 10), C4 = CaDM Ant with the
backward model."""
import sys
import time
sys.path.insert(0, ".")
import numpy as np
import torch
from cadm_b200.synth import build_model
from cadm_b200.dynamics.training import make_trainer
sys.path.insert(0, "tests")

def data(model, env, n, rng):
    D, A, K = env.obs_dim, env.act_dim, getattr(model, "history_length", 0)
    obs = rng.standard_normal((n, D)).astype(np.float32) * 0.5
    act = rng.uniform(-1, 1, (n, A)).astype(np.float32)
    delta = (rng.standard_normal((n, D)) * 0.1).astype(np.float32)
    if getattr(model, "_enc", None) is None:
        return (obs, act, delta)
    return (obs, act, delta, obs + delta, -delta, rng.standard_normal((n, D * K)).astype(np.float32), rng.uniform(-1, 1, (n, A * K)).astype(np.float32))

for config, kw in (("C2", {}), ("C3", {}), ("C4", dict(back_coeff=0.5))):
    model, env, cfg = build_model(config, m_max=1, candidates=64, **kw)
    rng = np.random.default_rng(0)
    n, B, E = 20000, 128, model.ensemble_size
    d = data(model, env, n, rng)
    stats = [np.asarray(s, np.float32) for s in model.get_normalization_stats()]
    stats = [np.where(np.abs(s) < 1e-6, s + (i % 2), s).astype(np.float32) for i, s in enumerate(stats)]      # std = 1 where unset
    for name, dev in (("native kernels", "cuda"), ("torch autograd (cuBLAS)", None)):
        if dev is None:
            from cadm_b200.dynamics import training as T
            tr = (T.CaDMTrainer(model._enc, model._dyn, model._back, model.env_name, model.deterministic, model.weight_decays,
                                model.context_weight_decays, model.weight_decay_coeff, model.back_coeff, model.learning_rate, device="cuda")
                  if model._enc is not None else
                  T.EnsembleNLLTrainer(model._dyn, model.env_name, model.deterministic, model.weight_decays, model.weight_decay_coeff,
                                       model.learning_rate, device="cuda"))
        else:
            tr = make_trainer(model, "cuda")
        tr.begin_fit(d, d, stats)
        idx = [rng.integers(0, n, size=(E, B)) for _ in range(60)]
        for i in range(10):
            tr.train_step_idx(idx[i])
        torch.cuda.synchronize()
        l0 = tr.launches if hasattr(tr, "launches") else 0
        t0 = time.perf_counter()
        for i in range(10, 60):
            out = tr.train_step_idx(idx[i])
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
        extra = f", {(tr.launches - l0) / 50:.0f} launches per step" if hasattr(tr, "launches") else ""
        print(f"{config} {name:26s}: {dt * 1e3:7.3f} ms per step (host clock, losses read back every step){extra}; last losses {tuple(round(x, 4) for x in out)}", flush=True)
    model.engine.close()
