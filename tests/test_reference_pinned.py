"""The planner oracle and the committed golden fixtures against recordings of the reference's OWN graph code.

tests/golden/make_reference_golden.py imports the unmodified builder functions of cadm/dynamics/core/utils.py and the
unmodified environment classes from /root/reference, runs them in float64 over a NumPy stand-in for the TensorFlow
functions they use (tests/golden/tf_numpy_shim.py) and records what they compute (tests/golden/recorded/planner_*.npz).
Here: the committed fixtures (which the CUDA engine is tested against on the GPU) and the oracle must reproduce those
recordings -- candidate returns, elite indices and final plan of every CEM iteration; returns, best candidate and action
of random shooting -- and the variable creation order of the reference must be the order save() / load() assume.
CPU only; nothing here reads /root/reference."""
import glob
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
from reference_cases import CASES, make_case, make_train_batch

from oracle import cadm_oracle as orc
from oracle import philox as ph
from oracle.envs import get_env

REF = np.load(os.path.join(HERE, "golden", "recorded", "planner_reference.npz"))
REF_CASES = np.load(os.path.join(HERE, "golden", "recorded", "planner_reference_cases.npz"))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
TOL = 1e-12


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f) for f in FIXTURES])
def test_committed_fixtures_equal_what_the_reference_graph_computes(path):
    """The fixtures were written by the oracle; the reference's graph code, fed the same weights and noise, gives the same
    numbers -- so the GPU test that compares the engine with the fixtures compares it with the reference."""
    name = os.path.splitext(os.path.basename(path))[0]
    g = np.load(path)
    scale = np.max(np.abs(g["out_returns"]))
    assert np.max(np.abs(REF[f"{name}/returns"] - g["out_returns"])) <= TOL * scale
    assert np.array_equal(REF[f"{name}/elites"], g["out_elites"])
    assert np.max(np.abs(REF[f"{name}/plan"] - g["out_mean"])) <= TOL
    if int(g["meta"][6]):
        assert np.max(np.abs(REF[f"{name}/ctx"] - g["out_ctx"])) <= TOL * np.max(np.abs(g["out_ctx"]))


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f) for f in FIXTURES])
def test_oracle_float32_mode_equals_the_reference_graph_run_in_float32(path):
    """The oracle is dtype-generic; run in float32 it follows the reference graph run in float32 (NumPy float32 kernels under
    both) step for step, and both stay within ~1e-7 of the float64 results on these problems: the 1e-4 parity tolerance is
    room for a different float32 summation order (tensor cores, split operands), not for a different formulation."""
    name = os.path.splitext(os.path.basename(path))[0]
    g = np.load(path)
    E, p, n, h, H, m, context, det, seed, C, K = [int(v) for v in g["meta"]]
    t = np.float32
    env = get_env(str(g["envname"]))
    prm = orc.DynamicsParams([g[f"W{i}"] for i in range(4)], [g[f"b{i}"] for i in range(4)], g["W_mu"], g["b_mu"], g["W_lv"],
                             g["b_lv"], g["max_logvar"], g["min_logvar"]).astype(t)
    norm = orc.NormStats(*[g[f"norm_{k}"] for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
                                                     "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std")]).astype(t)
    ctx_raw = None
    if context:
        enc = orc.EncoderParams([g[f"encW{i}"] for i in range(4)], [g[f"encb{i}"] for i in range(4)]).astype(t)
        ctx_raw = orc.encode_context(g["cp_obs"].astype(t), g["cp_act"].astype(t), enc, norm)
    z = ph.gen_z(seed, orc.NUM_CEM_ITERS, m, n, h, env.act_dim).astype(t)
    eps = None if det else ph.gen_eps(seed, orc.NUM_CEM_ITERS, h, m, n, p, E, env.obs_dim).astype(t)
    res = orc.cem_plan(g["obs"].astype(t), g["mean0"].astype(t), g["var0"].astype(t), z, prm, norm, env, E, p, bool(det), eps, ctx_raw)
    assert res.returns.dtype == np.float32
    scale = np.max(np.abs(g["out_returns"]))
    # the same bits on two of the three fixtures; on the third, 0.2 % of the last iteration's returns differ by one float32
    # ulp (NumPy's float32 reductions depend on the memory layout of their input, which the two programs do not share)
    assert np.max(np.abs(res.returns - REF[f"{name}/f32_returns"])) <= 2.5e-7 * scale
    assert np.array_equal(res.elites, REF[f"{name}/f32_elites"])
    assert np.max(np.abs(res.mean - REF[f"{name}/f32_plan"])) <= 2.5e-7
    assert np.max(np.abs(REF[f"{name}/f32_returns"] - REF[f"{name}/returns"])) / scale < 1e-6
    assert np.array_equal(REF[f"{name}/f32_elites"], REF[f"{name}/elites"])


def _oracle_inputs(g):
    E, p, n, h, H, m, context, det, seed, C, K = [int(v) for v in g["meta"]]
    f8 = np.float64
    env = get_env(str(g["envname"]))
    prm = orc.DynamicsParams([g[f"W{i}"] for i in range(4)], [g[f"b{i}"] for i in range(4)], g["W_mu"], g["b_mu"], g["W_lv"],
                             g["b_lv"], g["max_logvar"], g["min_logvar"]).astype(f8)
    norm = orc.NormStats(*[g[f"norm_{k}"] for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std",
                                                     "cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std")]).astype(f8)
    ctx_raw = None
    if context:
        enc = orc.EncoderParams([g[f"encW{i}"] for i in range(4)], [g[f"encb{i}"] for i in range(4)]).astype(f8)
        ctx_raw = orc.encode_context(g["cp_obs"].astype(f8), g["cp_act"].astype(f8), enc, norm)
    return dict(E=E, p=p, n=n, h=h, m=m, det=bool(det), seed=seed), env, prm, norm, ctx_raw


@pytest.mark.parametrize("name", [k for k, v in CASES.items() if v["mode"] == "cem"])
def test_oracle_cem_equals_the_reference_graph_on_the_other_environments(name):
    g = make_case(**CASES[name])
    c, env, prm, norm, ctx_raw = _oracle_inputs(g)
    f8 = np.float64
    z = ph.gen_z(c["seed"], orc.NUM_CEM_ITERS, c["m"], c["n"], c["h"], env.act_dim).astype(f8)
    eps = None if c["det"] else ph.gen_eps(c["seed"], orc.NUM_CEM_ITERS, c["h"], c["m"], c["n"], c["p"], c["E"], env.obs_dim).astype(f8)
    res = orc.cem_plan(g["obs"].astype(f8), g["mean0"].astype(f8), g["var0"].astype(f8), z, prm, norm, env, c["E"], c["p"],
                       c["det"], eps, ctx_raw)
    want = REF_CASES[f"{name}/returns"]
    assert np.max(np.abs(res.returns - want)) <= TOL * np.max(np.abs(want))
    assert np.array_equal(res.elites, REF_CASES[f"{name}/elites"])
    assert np.max(np.abs(res.mean - REF_CASES[f"{name}/plan"])) <= TOL
    assert np.ptp(want) > 1e-3                                       # the case separates candidates at all


@pytest.mark.parametrize("name", [k for k, v in CASES.items() if v["mode"] != "cem"])
def test_oracle_random_shooting_equals_the_reference_graph(name):
    g = make_case(**CASES[name])
    c, env, prm, norm, ctx_raw = _oracle_inputs(g)
    f8 = np.float64
    discrete = CASES[name]["mode"] == "rs_discrete"
    if discrete:
        u = ph.gen_discrete_actions(c["seed"], c["m"], c["n"], c["h"], env.act_dim)
    else:
        u = ph.gen_uniform_actions(c["seed"], c["m"], c["n"], c["h"], env.act_dim).astype(f8)
    eps = ph.gen_eps(c["seed"], 1, c["h"], c["m"], c["n"], c["p"], c["E"], env.obs_dim).astype(f8)[0]
    res = orc.rs_plan(g["obs"].astype(f8), u, prm, norm, env, c["E"], c["p"], c["det"], eps, ctx_raw, discrete=discrete)
    want = REF_CASES[f"{name}/returns"]
    assert np.max(np.abs(res["returns"] - want)) <= TOL * np.max(np.abs(want))
    assert np.array_equal(res["best"], REF_CASES[f"{name}/best"])
    assert np.array_equal(np.asarray(res["action"], f8), np.asarray(REF_CASES[f"{name}/plan"], f8))
    if discrete:
        assert len(np.unique(want)) > 2                              # some particles crossed the cart-pole limits, some did not


def test_the_cases_reach_the_branches_they_are_there_for():
    """Humanoid: one environment inside the alive band and one outside; pendulum: both sides of the atan2 branch cut."""
    hum = make_case(**CASES["humanoid_cem"])
    assert 1.0 < hum["obs"][0, 1] < 2.0 and not 1.0 < hum["obs"][-1, 1] < 2.0
    r = REF_CASES["humanoid_cem/returns"][0]
    assert r[0].mean() - r[-1].mean() > 5.0                          # the alive bonus shows in the returns
    pen = make_case(**CASES["pendulum_cem"])
    th = np.arctan2(pen["obs"][:, 1], pen["obs"][:, 0])
    assert th[0] > 3.0 and th[-1] < -3.0


def test_variable_creation_order_of_the_reference_is_the_checkpoint_order():
    """save() dumps tf.trainable_variables() in creation order (mlp_cadm_ensemble_cem_dynamics.py:571-577); the names below
    are what the reference's builders asked for, in order, when run by the generator.  cadm_b200's `params` list -- encoder
    (W, b) per layer, hidden (W, b) per layer, mu head, logvar head, max_logvar, min_logvar -- must follow it."""
    served = [str(s) for s in REF["hc_cadm_small/served"]]
    enc = [f"cp_hidden_{i}_{w}" for i in range(3) for w in ("weight", "bias")] + ["cp_output_weight", "cp_output_bias"]
    dyn = [f"hidden_{i}_{w}" for i in range(4) for w in ("weight", "bias")] + \
          ["output_mu_weight", "output_mu_bias", "output_logvar_weight", "output_logvar_bias"]
    assert served == enc + dyn + ["max_logvar", "min_logvar"]
    assert [str(s) for s in REF["hc_pets_small/served"]] == dyn + ["max_log_var", "min_log_var"]


def test_stand_in_ops_follow_the_tensorflow_rules_the_planner_depends_on():
    """The few places where the NumPy stand-in has to choose a convention: top_k and argmax tie-breaking (lower index first),
    the floored modulo, softplus for large arguments, gather along axis 0."""
    import tf_numpy_shim as shim
    shim.TAPE.__init__()
    vals, idx = shim.top_k(np.array([[1.0, 3.0, 3.0, 2.0, 3.0]]), k=3)
    assert idx.tolist() == [[1, 2, 4]] and idx.dtype == np.int32 and vals.tolist() == [[3.0, 3.0, 3.0]]
    assert shim.argmax(np.array([[0.0, 2.0, 2.0]]), 1, output_type=np.int32).tolist() == [1]
    assert np.isclose(shim.softplus(np.array([800.0, -800.0, 0.0])), [800.0, 0.0, np.log(2.0)]).all()
    assert shim.gather(np.arange(12).reshape(4, 3), np.array([3, 0])).tolist() == [[9, 10, 11], [0, 1, 2]]
    assert shim.reshape(np.arange(6), np.array([2, 3], np.int32)).shape == (2, 3)
    shim.TAPE.truncated = [np.full((2,), 0.5)]
    assert shim.random_truncated_normal([2], 1.0, 2.0).tolist() == [2.0, 2.0]
    with pytest.raises(RuntimeError):
        shim.random_normal([2])


def test_training_forward_passes_equal_the_reference_graph():
    """What fit() differentiates (cadm_b200/dynamics/training.py, PyTorch) against the reference's training graph on a
    bootstrap batch: the context encoder, the forward model's mu and soft-bounded logvar with the context appended, the
    deterministic backward model fed the next observation, and the PE-TS model without context."""
    import torch
    from cadm_b200.dynamics.training import CaDMTrainer, EnsembleNLLTrainer
    g = make_train_batch()
    C = int(g["meta"][9])
    D = g["obs"].shape[1]
    f8 = lambda a: np.asarray(a, np.float64)
    mlp = lambda pre: dict(W=[f8(g[f"{pre}W{i}"]) for i in range(4)], b=[f8(g[f"{pre}b{i}"]) for i in range(4)],
                           W_mu=f8(g[pre + "W_mu"]), b_mu=f8(g[pre + "b_mu"]), W_lv=f8(g[pre + "W_lv"]), b_lv=f8(g[pre + "b_lv"]),
                           max_logvar=f8(g["max_logvar"]).reshape(1, D), min_logvar=f8(g["min_logvar"]).reshape(1, D))
    enc = dict(W=[f8(g[f"encW{i}"]) for i in range(4)], b=[f8(g[f"encb{i}"]) for i in range(4)])
    stats = [f8(g[f"norm_{k}"]) for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std", "cp_obs_mean",
                                          "cp_obs_std", "cp_act_mean", "cp_act_std", "back_delta_mean", "back_delta_std")]
    tr = CaDMTrainer(enc, mlp(""), mlp("back"), "halfcheetah", False, (0.,) * 5, (0.,) * 4, 0.0, 0.5, 1e-3, dtype=torch.float64)
    t = lambda a: torch.as_tensor(f8(a))
    st = [t(s_) for s_ in stats]
    want = lambda k: REF_CASES[f"train_forward/{k}"]
    close = lambda got, ref: np.max(np.abs(got.detach().numpy() - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    with torch.no_grad():
        ctx = tr.context(t(g["bs_cp_obs"]), t(g["bs_cp_act"]), st)
        assert close(ctx, want("ctx"))
        mu, lv = tr.fwd.forward(t(g["bs_obs"]), t(g["bs_act"]), st, ctx)
        assert close(mu, want("fwd_mu")) and close(lv, want("fwd_logvar"))
        back_mu, _ = tr.back.forward(t(g["bs_next"]), t(g["bs_act"]), st, ctx)
        assert close(back_mu, want("back_mu"))
        pets = mlp("")
        pets["W"][0] = pets["W"][0][:, :pets["W"][0].shape[1] - C]
        mu, lv = EnsembleNLLTrainer(pets, "halfcheetah", False, (0.,) * 5, 0.0, 1e-3, dtype=torch.float64).forward(
            t(g["bs_obs"]), t(g["bs_act"]), st[:6])
        assert close(mu, want("pets_mu")) and close(lv, want("pets_logvar"))
    lv = want("fwd_logvar")
    assert 0.4 < lv.max() < 0.5 and -10.0 < lv.min() < -9.0          # both soft bounds were reached


MODEL_LOSS_CASES = [("cadm_prob_back0.5", False, 0.5), ("cadm_det_back0.5", True, 0.5), ("cadm_prob_back0.0", False, 0.0)]
LOSS_DECAYS = dict(weight_decays=(1e-3, 2e-3, 3e-3, 4e-3, 5e-3), context_weight_decays=(6e-3, 7e-3, 8e-3, 9e-3),
                   weight_decay_coeff=0.7)           # as in tests/golden/make_reference_golden.py


def _loss_inputs():
    g = make_train_batch()
    C, D = int(g["meta"][9]), g["obs"].shape[1]
    f8 = lambda a: np.asarray(a, np.float64)
    mlp = lambda pre: dict(W=[f8(g[f"{pre}W{i}"]) for i in range(4)], b=[f8(g[f"{pre}b{i}"]) for i in range(4)],
                           W_mu=f8(g[pre + "W_mu"]), b_mu=f8(g[pre + "b_mu"]), W_lv=f8(g[pre + "W_lv"]), b_lv=f8(g[pre + "b_lv"]),
                           max_logvar=f8(g["max_logvar"]).reshape(1, D), min_logvar=f8(g["min_logvar"]).reshape(1, D))
    enc = dict(W=[f8(g[f"encW{i}"]) for i in range(4)], b=[f8(g[f"encb{i}"]) for i in range(4)])
    stats = [f8(g[f"norm_{k}"]) for k in ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std", "cp_obs_mean",
                                          "cp_obs_std", "cp_act_mean", "cp_act_std", "back_delta_mean", "back_delta_std")]
    batch = (f8(g["bs_obs"]), f8(g["bs_act"]), REF_CASES["model_losses/targets/bs_delta"], f8(g["bs_next"]),
             REF_CASES["model_losses/targets/bs_back_delta"], f8(g["bs_cp_obs"]), f8(g["bs_cp_act"]))
    return g, C, mlp, enc, stats, batch


def _check_losses(tag, got):
    keys = [k.split("/")[-1] for k in REF_CASES.files if k.startswith(f"model_losses/{tag}/")]
    assert {"mse_loss", "recon_loss", "loss", "l2_reg_loss"} <= set(keys)
    for k in keys:
        want = float(REF_CASES[f"model_losses/{tag}/{k}"])
        assert abs(float(got[k]) - want) <= 1e-12 * max(1.0, abs(want)), (tag, k, float(got[k]), want)


@pytest.mark.parametrize("tag,det,back_coeff", MODEL_LOSS_CASES)
def test_cadm_training_losses_equal_the_reference_model_constructor(tag, det, back_coeff):
    """Every scalar the reference's CaDM model constructor defines (mlp_cadm_ensemble_cem_dynamics.py:266-314: mse, backward
    mse, the three l2 terms, mu / var / reg / recon / total loss), recorded from the UNMODIFIED class instantiated over the
    TensorFlow stand-in, against what fit() minimises (CaDMTrainer.losses) and against the NumPy restatement."""
    import torch
    from cadm_b200.dynamics.training import CaDMTrainer
    from oracle import train_oracle
    g, C, mlp, enc, stats, batch = _loss_inputs()
    d = LOSS_DECAYS
    tr = CaDMTrainer(enc, mlp(""), mlp("back"), "halfcheetah", det, d["weight_decays"], d["context_weight_decays"],
                     d["weight_decay_coeff"], back_coeff, 1e-3, dtype=torch.float64)
    with torch.no_grad():
        _check_losses(tag, tr.losses(*batch, stats))
    _check_losses(tag, train_oracle.cadm_losses(enc, mlp(""), mlp("back"), "halfcheetah", det, d["weight_decays"],
                                                d["context_weight_decays"], d["weight_decay_coeff"], back_coeff, *batch, stats))


@pytest.mark.parametrize("tag,det", [("pets_prob", False), ("pets_det", True)])
def test_pets_training_losses_equal_the_reference_model_constructor(tag, det):
    """The same for the PE-TS / vanilla model (mlp_ensemble_cem_dynamics.py:140-167)."""
    import torch
    from cadm_b200.dynamics.training import EnsembleNLLTrainer
    from oracle import train_oracle
    g, C, mlp, enc, stats, batch = _loss_inputs()
    d = LOSS_DECAYS
    dyn = mlp("")
    dyn["W"][0] = dyn["W"][0][:, :dyn["W"][0].shape[1] - C]
    tr = EnsembleNLLTrainer(dyn, "halfcheetah", det, d["weight_decays"], d["weight_decay_coeff"], 1e-3, dtype=torch.float64)
    with torch.no_grad():
        _check_losses(tag, tr.losses(batch[0], batch[1], batch[2], stats[:6]))
    _check_losses(tag, train_oracle.pets_losses(dyn, "halfcheetah", det, d["weight_decays"], d["weight_decay_coeff"], batch[0],
                                                batch[1], batch[2], stats[:6]))


class _LegacyRng:
    """The Generator interface fit_ensemble / fit_cadm_ensemble draw from, forwarded to NumPy's global legacy generator, which
    is what the reference's fit() uses (np.random.permutation / randint / uniform): the same seed gives the same draws."""

    @staticmethod
    def permutation(n):
        return np.random.permutation(n)

    @staticmethod
    def integers(lo, hi, size):
        return np.random.randint(lo, hi, size=size)

    @staticmethod
    def uniform(size):
        return np.random.uniform(size=size)


from cadm_b200.dynamics.training import IndexedFeed


class _RecordingTrainer(IndexedFeed):
    """Stands where the reference has its session: records what every training / validation step is fed and answers with the
    scripted losses of the recording."""

    calls, script = [], []

    def __init__(self, *a, **k):
        pass

    def train_step(self, *batch_and_stats):
        type(self).calls.append(("train", batch_and_stats))
        return (0.25,) * (2 if len(batch_and_stats) == 4 else 3)

    def evaluate(self, *batch_and_stats):
        type(self).calls.append(("valid", batch_and_stats))
        n_valid = sum(k == "valid" for k, _ in type(self).calls)
        return (0.5,) * (1 if len(batch_and_stats) == 4 else 2) + (type(self).script[n_valid - 1],)

    def export(self, *a):
        pass


@pytest.mark.parametrize("tag", ["pets", "cadm"])
def test_fit_loop_feeds_what_the_reference_fit_feeds(tag, monkeypatch):
    """fit() of the UNMODIFIED reference classes was run with NumPy's global generator seeded and its session replaced by a
    recorder (make_reference_golden.py run_fit_replay).  Here fit_ensemble / fit_cadm_ensemble run on the same data with the same
    generator and a recording trainer: every minibatch of every epoch (per-member bootstrap rows after the per-epoch reshuffle,
    future steps flattened and masked), the validation batches, the normalisation statistics and the epoch at which the
    early-stopping rule fires must be the reference's, bit for bit."""
    from reference_cases import FIT_REPLAY, make_fit_data
    from test_training import _CpuCadmModel, _CpuModel
    from cadm_b200.dynamics import training
    c, data = FIT_REPLAY, make_fit_data()
    D, A = 18, 6
    _RecordingTrainer.calls, _RecordingTrainer.script = [], list(c["valid_script"])
    monkeypatch.setattr(training, "EnsembleNLLTrainer", _RecordingTrainer)
    monkeypatch.setattr(training, "CaDMTrainer", _RecordingTrainer)
    np.random.seed(c["np_seed"])
    kw = dict(epochs=c["epochs"], rolling_average_persitency=c["persistency"], rng=_LegacyRng(), device="cpu", log=lambda *_: None)
    if tag == "pets":
        model = _CpuModel("halfcheetah", E=c["E"], H=c["H"], batch_size=c["batch_size"])
        info = training.fit_ensemble(model, data["obs"][:, :D], data["act"][:, :A], data["obs_next"][:, :D], **kw)
        names = ("bs_obs", "bs_act", "bs_delta")
        stat_names = ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std")
    else:
        model = _CpuCadmModel("halfcheetah", E=c["E"], H=c["H"], K=c["K"], F=c["F"], C=c["C"], back_coeff=0.5, batch_size=c["batch_size"])
        info = training.fit_cadm_ensemble(model, data["obs"], data["act"], data["obs_next"], data["cp_obs"], data["cp_act"],
                                          data["future_bool"], **kw)
        names = ("bs_obs", "bs_act", "bs_delta", "bs_obs_next", "bs_back_delta", "bs_cp_obs", "bs_cp_act")
        stat_names = ("obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std", "cp_obs_mean", "cp_obs_std",
                      "cp_act_mean", "cp_act_std", "back_delta_mean", "back_delta_std")
    rec = lambda k: REF_CASES[f"fit_replay/{tag}/{k}"]
    kinds = [k for k, _ in _RecordingTrainer.calls]
    assert kinds == list(rec("kinds")) and info["epochs"] == int(rec("valid_runs")) == 3
    assert kinds.count("train") > kinds.count("valid") == 3                       # several minibatches per epoch
    for i, (_, fed) in enumerate(_RecordingTrainer.calls):
        *batch, stats = fed
        assert len(batch) == len(names)
        for name, got in zip(names, batch):
            want = rec(f"call{i:03d}/{name}")
            assert got.shape == want.shape and np.array_equal(np.asarray(got, np.float64), want), (tag, i, name)
        if i == 0:
            for name, got in zip(stat_names, stats):
                assert np.array_equal(np.asarray(got, np.float64), rec(f"call000/norm_{name}")), (tag, name)
    if tag == "cadm":                                                              # the ragged future masks removed rows
        n_rows = sum(b[0].shape[1] for k, b in _RecordingTrainer.calls[:4] if k == "train")
        assert n_rows < int(0.8 * c["n"]) * c["F"] + 1


class _FakeEngine:
    """Records what the host model hands to the native boundary; answers like the library would (the CEM entry point clips
    inside the C call, cadm_plan_cem_host; random shooting returns the raw first action)."""

    def __init__(self, cem_answer, rs_answer):
        self.cem_answer, self.rs_answer, self.norm, self.cem_args, self.rs_args = cem_answer, rs_answer, None, None, None

    def set_norm(self, *stats):
        self.norm = [np.asarray(s, np.float64) for s in stats]

    def plan_cem_host(self, obs, init_mean, init_var, cp_obs=None, cp_act=None, seed=0):
        self.cem_args = dict(obs=obs, cem_init_mean=init_mean, cem_init_var=init_var, cp_obs=cp_obs, cp_act=cp_act)
        return np.clip(self.cem_answer, -1, 1).astype(np.float32)

    def plan_rs(self, obs, cp_obs=None, cp_act=None, seed=0):
        import torch
        self.rs_args = dict(obs=obs, cp_obs=cp_obs, cp_act=cp_act)
        return dict(action=torch.as_tensor(np.asarray(self.rs_answer)))


@pytest.mark.parametrize("name", ["cadm_norm", "cadm_state_diff", "cadm_raw", "pets_norm", "pets_raw", "pets_discrete"])
def test_get_action_hands_the_planner_what_the_reference_get_action_feeds(name):
    """get_action() of the UNMODIFIED reference model classes was run with its session replaced by a recorder.
    The host mirrors, with a recording engine in place of the library, must hand over the same observation / history / warm
    start, the same normalisation vectors in the same slots (zeros / ones when normalize_input is off, for the observation
    history under state_diff, for discrete actions), and return the same action: clipped to [-1, 1] for continuous actions,
    untouched for discrete ones."""
    from collections import OrderedDict
    from test_training import _CpuCadmModel, _CpuModel
    from cadm_b200.dynamics.core import PlannerModelBase
    from reference_cases import GET_ACTION_CASES, get_action_inputs
    spec = GET_ACTION_CASES[name]
    x = get_action_inputs(name, spec)
    context = spec["kind"] == "cadm"
    stat_names = ["obs_mean", "obs_std", "act_mean", "act_std", "delta_mean", "delta_std"] + \
        (["cp_obs_mean", "cp_obs_std", "cp_act_mean", "cp_act_std"] if context else [])
    for mode in (["rs"] if x["discrete"] else ["cem", "rs"]):
        rec = lambda k: REF_CASES[f"get_action/{name}/{mode}/{k}"]
        model = (_CpuCadmModel(spec["envname"], E=2, H=8, K=x["K"], F=1, C=4) if context else _CpuModel(spec["envname"], E=2, H=8))
        model.use_cem, model.n_forwards, model._seed, model._calls = mode == "cem", x["h"], 0, 0
        model.normalize_input, model.discrete = spec["normalize_input"], x["discrete"]
        if context:
            model.state_diff = spec["state_diff"]
        model.engine = _FakeEngine(x["cem_answer"], x["rs_answer"])
        model.normalization = OrderedDict(x["normalization"]) if spec["normalize_input"] else None
        PlannerModelBase._push_norm(model)                               # what set_normalization / load / fit do
        assert len(model.engine.norm) == len(stat_names)
        for k, got in zip(stat_names, model.engine.norm):
            assert np.array_equal(got, rec(f"feed_norm_{k}")), (name, mode, k)
        hist = (x["cp_obs"], x["cp_act"]) if context else ()
        warm = (x["init_mean"], x["init_var"]) if mode == "cem" else ()
        action = model.get_action(x["obs"], *hist, *warm)
        fed = model.engine.cem_args if mode == "cem" else model.engine.rs_args
        want_keys = ["obs"] + (["cp_obs", "cp_act"] if context else []) + (["cem_init_mean", "cem_init_var"] if mode == "cem" else [])
        for k in want_keys:
            assert np.array_equal(np.asarray(fed[k], np.float32), rec(f"feed_{k}").astype(np.float32)), (name, mode, k)
        want = rec("action")
        assert action.shape == want.shape
        if x["discrete"]:
            assert np.array_equal(action, want) and np.array_equal(want, x["rs_answer"])      # no clip
        else:
            assert np.max(np.abs(action - want)) < 1e-6 and np.abs(want).max() == 1.0          # the clip was active


def test_load_reads_a_checkpoint_written_by_the_reference_save():
    """SURVEY 8(f) rank 1.  tests/golden/recorded/reference_checkpoint.joblib (+ _norm_stats) was written by save() of the
    UNMODIFIED CaDM model class with a backward model (mlp_cadm_ensemble_cem_dynamics.py:571-577).  load() of the host mirror
    must put every array into the slot the reference created it in -- encoder, forward model, backward model, each with its
    own logvar bounds -- and a model built without a backward model must accept the same file and ignore the tail."""
    from test_training import _CpuCadmModel
    path = os.path.join(HERE, "golden", "recorded", "reference_checkpoint.joblib")
    ref = {k.split("/", 1)[1]: REF_CASES[k] for k in REF_CASES.files if k.startswith("checkpoint/var")}
    by_name = {k.split(":", 1)[1]: v for k, v in ref.items()}
    order = [k.split(":", 1)[1] for k in sorted(ref)]
    assert len(order) == 8 + 2 * (4 + 4 + 2)                          # encoder 4 x (W, b); each MLP: 2 x (W, b), heads, bounds
    pick = lambda scope, name: next(v for k, v in by_name.items() if f"/{scope}/" in k and k.endswith("/" + name))

    def check_mlp(d, scope):
        for i in range(2):
            assert np.array_equal(d["W"][i], pick(scope, f"hidden_{i}_weight")) and np.array_equal(d["b"][i], pick(scope, f"hidden_{i}_bias"))
        assert np.array_equal(d["W_mu"], pick(scope, "output_mu_weight")) and np.array_equal(d["b_mu"], pick(scope, "output_mu_bias"))
        assert np.array_equal(d["W_lv"], pick(scope, "output_logvar_weight")) and np.array_equal(d["b_lv"], pick(scope, "output_logvar_bias"))
        mx = [v for k, v in by_name.items() if f"/{scope}/" in k and k.rsplit("/", 1)[1] in ("max_logvar", "max_log_var")]
        mn = [v for k, v in by_name.items() if f"/{scope}/" in k and k.rsplit("/", 1)[1] in ("min_logvar", "min_log_var")]
        assert len(mx) == 1 and len(mn) == 1
        assert np.array_equal(d["max_logvar"], mx[0]) and np.array_equal(d["min_logvar"], mn[0])

    for back_coeff in (0.5, 0.0):
        model = _CpuCadmModel("halfcheetah", E=2, H=8, n_hidden=2, K=3, F=1, C=4, cp_hidden=(8, 8, 8), back_coeff=back_coeff)
        model.load(path)
        assert model.pushed == 1
        for i in range(3):
            assert np.array_equal(model._enc["W"][i], pick("context_model", f"cp_hidden_{i}_weight"))
            assert np.array_equal(model._enc["b"][i], pick("context_model", f"cp_hidden_{i}_bias"))
        assert np.array_equal(model._enc["W"][3], pick("context_model", "cp_output_weight"))
        assert np.array_equal(model._enc["b"][3], pick("context_model", "cp_output_bias"))
        check_mlp(model._dyn, "ff_model")
        if back_coeff > 0:
            check_mlp(model._back, "backward_model")
            assert not np.array_equal(model._back["max_logvar"], model._dyn["max_logvar"])       # distinct values per scope
        else:
            assert model._back is None
        assert list(model.normalization) == ["obs", "delta", "act", "cp_obs", "cp_act", "back_delta"]
        for k, (mu, sd) in model.normalization.items():
            assert np.array_equal(mu, REF_CASES[f"checkpoint/norm/{k}/mean"]) and np.array_equal(sd, REF_CASES[f"checkpoint/norm/{k}/std"])
    # and the way back: what save() of the mirror writes is, array for array, what the reference wrote
    import joblib
    import tempfile
    model = _CpuCadmModel("halfcheetah", E=2, H=8, n_hidden=2, K=3, F=1, C=4, cp_hidden=(8, 8, 8), back_coeff=0.5)
    model.load(path)
    with tempfile.TemporaryDirectory() as tmp:
        model.save(os.path.join(tmp, "ck"))
        ours, theirs = joblib.load(os.path.join(tmp, "ck")), joblib.load(path)
        assert len(ours) == len(theirs) and all(np.array_equal(a, b) and a.shape == b.shape for a, b in zip(ours, theirs))
        n_ours, n_theirs = joblib.load(os.path.join(tmp, "ck_norm_stats")), joblib.load(path + "_norm_stats")
        assert list(n_ours) == list(n_theirs) and all(np.array_equal(n_ours[k][j], n_theirs[k][j]) for k in n_ours for j in (0, 1))
