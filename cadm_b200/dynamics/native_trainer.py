"""fit() on the B200: ctypes front end of the hand-written training kernels (csrc/trainer.cu, cadm_train_* in
include/cadm_b200.h).  It is what `fit()` uses whenever the model's engine lives on a CUDA device; the PyTorch trainers of
training.py restate the same graph with autograd and serve as its checker on the CPU (tests/test_training.py) and on the GPU
(tests/test_gpu_training.py).  No torch on this path: NumPy arrays in, NumPy arrays out, the dataset resident on the device.

Reference: the loss / optimiser construction of cadm/dynamics/mlp_ensemble_cem_dynamics.py:150-170 and
mlp_cadm_ensemble_cem_dynamics.py:266-317, fed per minibatch by the fit loops (:262-283 / :478-527).
"""
import ctypes as C

import numpy as np

from .. import _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class NativeTrainer:
    """One handle = parameters, gradients, Adam slots and the datasets on the device.  `enc` / `back` are None for the PE-TS
    model; `back` is None unless back_coeff > 0.  The handle (and so the optimiser state) lives as long as this object: the
    model keeps it across fit() calls, as the reference keeps its tf.train.AdamOptimizer."""

    def __init__(self, enc, dyn, back, env_name, obs_dim, proc_obs_dim, act_dim, hist_len, deterministic, weight_decays,
                 context_weight_decays, weight_decay_coeff, back_coeff, learning_rate, device=None):
        self.lib = _lib.load()
        self._h = C.c_void_p()
        self.has_enc, self.has_back = enc is not None, back is not None
        E = dyn["W"][0].shape[0]
        n_hidden = len(dyn["W"])
        wd = list(weight_decays)
        cfg = _lib.CadmTrainConfig()
        cfg.struct_size = C.sizeof(_lib.CadmTrainConfig)
        cfg.env_id = _lib.ENV_IDS[env_name]
        cfg.obs_dim, cfg.proc_obs_dim, cfg.act_dim = obs_dim, proc_obs_dim, act_dim
        cfg.ctx_dim = enc["W"][-1].shape[2] if self.has_enc else 0
        cfg.hist_len = hist_len if self.has_enc else 0
        cfg.hidden, cfg.n_hidden, cfg.ensemble = dyn["W"][0].shape[2], n_hidden, E
        if any(w.shape[2] != cfg.hidden for w in dyn["W"]):
            raise ValueError("the native trainer expects hidden layers of one width")
        if self.has_enc:
            hs = [w.shape[2] for w in enc["W"][:-1]]
            if len(hs) > 3:
                raise ValueError("at most three encoder hidden layers")
            for i, h in enumerate(hs):
                cfg.enc_hidden[i] = h
            cwd = list(context_weight_decays)
            for i in range(len(hs)):
                cfg.context_weight_decays[i] = float(cwd[i])
            cfg.context_weight_decays[len(hs)] = float(cwd[-1])       # the output layer takes [-1] (core/utils.py:597-603)
        cfg.deterministic = int(bool(deterministic))
        cfg.has_back = int(self.has_back)
        cfg.back_coeff = float(back_coeff) if self.has_back else 0.0
        cfg.weight_decay_coeff, cfg.learning_rate = float(weight_decay_coeff), float(learning_rate)
        if n_hidden + 1 > 8:
            raise ValueError("too many hidden layers")
        for i in range(n_hidden):                                     # create_dense_layer(weight_decay=weight_decays[idx])
            cfg.weight_decays[i] = float(wd[min(i, len(wd) - 1)])
        cfg.weight_decays[n_hidden] = float(wd[-1])                   # both heads: weight_decays[-1] (core/utils.py:46-69)
        cfg.adam_beta1, cfg.adam_beta2, cfg.adam_eps = 0.9, 0.999, 1e-8
        self._device = device
        with self._on_device():
            rc = self.lib.cadm_train_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise _lib.CadmError(f"cadm_train_create failed ({rc}): {self.lib.cadm_train_last_error(None).decode()}")
        self.cfg = cfg
        self.D, self.E = obs_dim, E
        self.n = int(self.lib.cadm_train_param_count(self._h))
        flat = self._pack(enc, dyn, back)
        assert flat.size == self.n, (flat.size, self.n)
        self._chk(self.lib.cadm_train_set_params(self._h, flat.ctypes.data_as(C.c_void_p), self.n))
        self._rows = [0, 0]

    # ------------------------------------------------------------------ plumbing
    def _on_device(self):
        import contextlib
        if self._device is None:
            return contextlib.nullcontext()
        import torch
        return torch.cuda.device(self._device)

    def _chk(self, rc):
        if rc != 0:
            raise _lib.CadmError(f"cadm_b200 trainer error {rc}: {self.lib.cadm_train_last_error(self._h).decode()}")

    def close(self):
        if self._h:
            self.lib.cadm_train_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ flat parameter vector (layout: include/cadm_b200.h)
    @staticmethod
    def _mlp_parts(d):
        out = []
        for W, b in zip(d["W"], d["b"]):
            out += [W, np.reshape(b, (b.shape[0], -1))]
        out.append(np.concatenate([d["W_mu"], d["W_lv"]], axis=2))
        out.append(np.concatenate([np.reshape(d["b_mu"], (d["b_mu"].shape[0], -1)), np.reshape(d["b_lv"], (d["b_lv"].shape[0], -1))], axis=1))
        return out

    def _pack(self, enc, dyn, back):
        parts = []
        if enc is not None:
            for W, b in zip(enc["W"], enc["b"]):
                parts += [W, np.reshape(b, (b.shape[0], -1))]
        parts += self._mlp_parts(dyn)
        parts += [np.reshape(dyn["max_logvar"], -1), np.reshape(dyn["min_logvar"], -1)]
        if back is not None:
            parts += self._mlp_parts(back)
        return np.concatenate([_f32(p).reshape(-1) for p in parts])

    def _unpack(self, flat, enc, dyn, back):
        pos = 0

        def take(dst):
            nonlocal pos
            n = dst.size
            dst[...] = flat[pos:pos + n].reshape(dst.shape)
            pos += n

        def take_mlp(d):
            nonlocal pos
            for W, b in zip(d["W"], d["b"]):
                take(W)
                take(b)
            E, H, D = d["W_mu"].shape
            heads = flat[pos:pos + E * H * 2 * D].reshape(E, H, 2 * D)
            pos += heads.size
            d["W_mu"][...] = heads[:, :, :D]
            d["W_lv"][...] = heads[:, :, D:]
            hb = flat[pos:pos + E * 2 * D].reshape(E, 2 * D)
            pos += hb.size
            d["b_mu"][...] = hb[:, :D].reshape(d["b_mu"].shape)
            d["b_lv"][...] = hb[:, D:].reshape(d["b_lv"].shape)

        if enc is not None:
            for W, b in zip(enc["W"], enc["b"]):
                take(W)
                take(b)
        take_mlp(dyn)
        take(dyn["max_logvar"])
        take(dyn["min_logvar"])
        if back is not None:
            take_mlp(back)
        assert pos == flat.size

    def flat_params(self):
        out = np.empty(self.n, np.float32)
        self._chk(self.lib.cadm_train_get_params(self._h, out.ctypes.data_as(C.c_void_p), self.n))
        return out

    def flat_grads(self):
        out = np.empty(self.n, np.float32)
        self._chk(self.lib.cadm_train_get_grads(self._h, out.ctypes.data_as(C.c_void_p), self.n))
        return out

    def adam_state(self):
        m, v, t = np.empty(self.n, np.float32), np.empty(self.n, np.float32), C.c_int64(0)
        self._chk(self.lib.cadm_train_adam_state(self._h, 0, m.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), self.n, C.byref(t)))
        return m, v, int(t.value)

    def export(self, *dicts):
        """Write the trained values back into the model's arrays: export(dyn) or export(enc, dyn, back)."""
        enc, dyn, back = (None, dicts[0], None) if len(dicts) == 1 else dicts
        self._unpack(self.flat_params(), enc if self.has_enc else None, dyn, back if self.has_back else None)

    @property
    def launches(self):
        return int(self.lib.cadm_train_launch_count(self._h))

    # ------------------------------------------------------------------ what the fit loops call
    def begin_fit(self, train, valid, stats):
        """train / valid: tuples of arrays, (obs, act, delta) for PE-TS, (obs, act, delta, obs_next, back_delta, cp_obs, cp_act)
        for CaDM; stats: get_normalization_stats().  Uploads everything once; minibatches are index matrices afterwards."""
        stats = [_f32(s).reshape(-1) for s in stats]
        need = 12 if self.has_back else (10 if self.has_enc else 6)
        if len(stats) < need:
            raise ValueError(f"need {need} normalisation vectors, got {len(stats)}")
        arr = (C.c_void_p * need)(*[s.ctypes.data for s in stats[:need]])
        self._chk(self.lib.cadm_train_set_norm(self._h, arr, need))
        for which, data in ((0, train), (1, valid)):
            data = [_f32(a) for a in data]
            rows = data[0].shape[0]
            ptr = lambda i: data[i].ctypes.data_as(C.c_void_p) if i < len(data) and rows > 0 else None
            order = (0, 1, 2, 3, 4, 5, 6)                         # obs, act, delta, obs_next, back_delta, cp_obs, cp_act
            self._chk(self.lib.cadm_train_set_dataset(self._h, which, rows, *[ptr(i) for i in order]))
            self._rows[which] = rows

    def _step(self, which, idx, train):
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int32))
        if idx.ndim != 2 or idx.shape[0] != self.E:
            raise ValueError("idx must be [E, B]")
        out = np.zeros(4, np.float32)
        with self._on_device():
            self._chk(self.lib.cadm_train_step(self._h, which, idx.ctypes.data_as(C.c_void_p), idx.shape[1], int(train),
                                               out.ctypes.data_as(C.c_void_p)))
        mse, recon, back_mse = float(out[0]), float(out[1]), float(out[2])
        return (mse, back_mse, recon) if self.has_enc else (mse, recon)

    def train_step_idx(self, idx):
        return self._step(0, idx, True)

    def evaluate_idx(self, idx, which=1):
        return self._step(which, idx, False)
