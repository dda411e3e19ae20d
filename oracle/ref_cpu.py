"""TEST INFRASTRUCTURE (oracle) -- `ref_cpu`: op-for-op PyTorch-CPU restatement of the reference's planner graph.

This is the CPU baseline timed beside the GPU numbers ("CPU restatement of the TF1.15 graph; TF is not installable
here"): the same tile / transpose / reshape / concat -> batched matmul x6 -> sigmoid / softplus /
exp -> randn -> postproc / reward -> reduce_mean / top_k / gather / refit sequence as
cadm/dynamics/core/utils.py:130-184 (PE-TS) and :400-488 (CaDM), in float32, multi-threaded through torch's intra-op
pool (MKL/oneDNN sgemm, like TF's Eigen/MKL CPU kernels).  tests/test_ref_cpu.py cross-checks it against the NumPy
oracle with injected noise, so the timed thing is the verified thing.  Never imported by cadm_b200/.
"""
import math

import torch

NUM_ELITES, NUM_CEM_ITERS, ALPHA = 50, 5, 0.1


def _normalize(x, mean, std):
    return (x - mean) / (std + 1e-10)


def _denormalize(x, mean, std):
    return x * (std + 1e-10) + mean


class RefCpuPlanner:
    def __init__(self, prm, norm, env_name, E, p, deterministic, enc=None, threads=None):
        """prm / enc / norm: the oracle dataclasses (NumPy); converted to float32 torch tensors."""
        if threads:
            torch.set_num_threads(threads)
        t = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32)
        self.W = [t(w) for w in prm.W]
        self.b = [t(x) for x in prm.b]
        self.W_mu, self.b_mu, self.W_lv, self.b_lv = t(prm.W_mu), t(prm.b_mu), t(prm.W_lv), t(prm.b_lv)
        self.max_lv, self.min_lv = t(prm.max_logvar), t(prm.min_logvar)
        self.n = {k: t(getattr(norm, k)) for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
                                                    "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std")}
        self.enc = None if enc is None else ([t(w) for w in enc.W], [t(x) for x in enc.b])
        self.env, self.E, self.p, self.det = env_name, E, p, deterministic

    # ---- env closures (cadm/envs/half_cheetah_env.py:46-56,82-88 ; ant_env.py:52-59,89-98)
    def preproc(self, o):
        if self.env in ("halfcheetah", "cripple_halfcheetah"):
            return torch.cat([o[..., 1:2], torch.sin(o[..., 2:3]), torch.cos(o[..., 2:3]), o[..., 3:]], dim=-1)
        if self.env == "ant":
            return o[..., 1:]
        return o

    def postproc(self, o, pred):
        if self.env in ("halfcheetah", "cripple_halfcheetah", "ant"):
            return torch.cat([pred[..., :1], o[..., 1:] + pred[..., 1:]], dim=-1)
        return o + pred

    def reward(self, o, a, nxt):
        if self.env in ("halfcheetah", "cripple_halfcheetah"):
            return o[..., 0] - 1e-1 * torch.sum(torch.square(a), dim=-1)
        if self.env == "ant":
            return o[..., 0] + (-0.005) * torch.sum(torch.square(a), dim=-1) + 0.0 + 0.05
        raise NotImplementedError(self.env)

    # ---- forward (core/utils.py:73-92)
    def forward(self, x, eps=None):
        for W, b in zip(self.W, self.b):
            x = torch.matmul(x, W) + b
            x = x * torch.sigmoid(x)
        mu = torch.matmul(x, self.W_mu) + self.b_mu
        dmu = _denormalize(mu, self.n["delta_mean"], self.n["delta_std"])
        if self.det:
            return dmu
        logvar = torch.matmul(x, self.W_lv) + self.b_lv
        logvar = self.max_lv - torch.nn.functional.softplus(self.max_lv - logvar)
        logvar = self.min_lv + torch.nn.functional.softplus(logvar - self.min_lv)
        dstd = torch.exp((logvar + 2 * torch.log(self.n["delta_std"])) / 2.0)
        noise = torch.randn(dmu.shape) if eps is None else eps
        return dmu + noise * dstd

    def encode(self, cp_obs, cp_act):
        E = self.E
        bo = cp_obs[None].repeat(E, 1, 1)
        ba = cp_act[None].repeat(E, 1, 1)
        x = torch.cat([_normalize(bo, self.n["cp_obs_mean"], self.n["cp_obs_std"]),
                       _normalize(ba, self.n["cp_act_mean"], self.n["cp_act_std"])], dim=-1)
        W, b = self.enc
        for i in range(len(W)):
            x = torch.matmul(x, W[i]) + b[i]
            if i < len(W) - 1:
                x = torch.relu(x)
        return x

    # ---- planner (core/utils.py:130-184 / 424-488)
    @torch.no_grad()
    def cem(self, obs, init_mean, init_var, n, cp_obs=None, cp_act=None, z=None, eps=None, iters=NUM_CEM_ITERS):
        f = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32)
        obs, mean, var, cp_obs, cp_act, z, eps = map(f, (obs, init_mean, init_var, cp_obs, cp_act, z, eps))
        E, p = self.E, self.p
        m, D = obs.shape
        h, A = mean.shape[1], mean.shape[2]
        q = p // E
        bs_cp = self.encode(cp_obs, cp_act) if self.enc is not None else None
        all_ret, all_el = [], []
        for it in range(iters):
            lb, ub = mean - (-1.0), 1.0 - mean
            cvar = torch.minimum(torch.minimum(torch.square(lb / 2), torch.square(ub / 2)), var)
            rmean = mean[:, None].repeat(1, n, 1, 1)
            rvar = cvar[:, None].repeat(1, n, 1, 1)
            if z is None:
                zz = torch.empty(m, n, h, A)
                torch.nn.init.trunc_normal_(zz, 0.0, 1.0, -2.0, 2.0)
            else:
                zz = z[it]
            actions = rmean + torch.sqrt(rvar) * zz
            returns = 0
            observation = obs.reshape(m, 1, 1, D).repeat(1, n, p, 1)
            ctx_rows = None
            if bs_cp is not None:
                bs_cp = bs_cp.permute(1, 0, 2)                                   # inside the loop (quirk Q3)
                C = bs_cp.shape[-1]
                context = bs_cp.reshape(m, 1, E, C).repeat(1, n, q, 1)
                ctx_rows = context.permute(2, 0, 1, 3).reshape(E, q * m * n, C)
            for t in range(h):
                action = actions[:, :, t]
                nact = _normalize(action, self.n["act_mean"], self.n["act_std"])
                nact = nact[:, :, None, :].repeat(1, 1, p, 1).permute(2, 0, 1, 3).reshape(E, q * m * n, A)
                pobs = self.preproc(observation)
                nobs = _normalize(pobs, self.n["obs_mean"], self.n["obs_std"])
                nobs = nobs.permute(2, 0, 1, 3).reshape(E, q * m * n, nobs.shape[-1])
                x = torch.cat([nobs, nact] + ([ctx_rows] if ctx_rows is not None else []), dim=2)
                delta = self.forward(x, None if eps is None else eps[it, t])
                delta = delta.reshape(p, m, n, D).permute(1, 2, 0, 3)
                nxt = self.postproc(observation, delta)
                rep = action[:, :, None, :].repeat(1, 1, p, 1)
                returns = returns + self.reward(observation, rep, nxt)
                observation = nxt
            returns = returns.mean(dim=2)
            _, idx = torch.topk(returns, NUM_ELITES, dim=1, sorted=True)
            flat = (idx + torch.arange(0, m * n, n)[:, None]).reshape(-1)
            elites = actions.reshape(m * n, h, A)[flat].reshape(m, NUM_ELITES, h, A)
            new_mean = elites.mean(dim=1)
            new_var = torch.square(elites - new_mean[:, None]).mean(dim=1)
            mean = mean * ALPHA + (1 - ALPHA) * new_mean
            var = var * ALPHA + (1 - ALPHA) * new_var
            all_ret.append(returns)
            all_el.append(idx)
        return dict(mean=mean.numpy(), var=var.numpy(), returns=torch.stack(all_ret).numpy(),
                    elites=torch.stack(all_el).numpy(), action=mean.clamp(-1.0, 1.0).numpy())


def cpu_info():
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    import os
    return dict(model=model, cores=os.cpu_count(), torch_threads=torch.get_num_threads())
