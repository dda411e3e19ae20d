"""Generate tests/golden/recorded/sampler_golden.npz by running the UNMODIFIED reference sampler and sample processor.

`cadm/samplers/sampler.py` (Sampler.obtain_samples) and `cadm/samplers/model_sample_processor.py`
(ModelSampleProcessor.process_samples) are pure NumPy at run time; only their import chain pulls in TensorFlow 1.15 and
pyprind, which this container lacks.  Both are replaced by inert stand-ins in sys.modules (no reference file is modified
or copied), the reference classes are imported from /root/reference and driven with the stand-in environment and policy
of tests/sampler_fakes.py.  Recorded per scenario: every argument of every policy.get_actions() call, the finished paths,
and the arrays process_samples() returns; for the evaluation rollouts of `cadm/samplers/utils.py` (rollout_multi,
context_rollout_multi) the calls and the returned average; for `cadm/policies/mpc_controller.py` which arguments reach
dynamics_model.get_action(), in which order, for every (context, use_cem) combination.  tests/test_samplers.py replays the same scenarios through cadm_b200's classes.

Run in the build container (needs /root/reference):  python tests/golden/make_sampler_golden.py
"""
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True                              # /root/reference is read-only: leave no __pycache__ there

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sampler_fakes import FakeEnv, ScriptedPolicy, ScriptedSinglePolicy, controller_dispatch_log          # noqa: E402

SCENARIOS = dict(
    # name: (context, state_diff, use_cem, history_length, future_length, num_rollouts, max_path_length, horizon)
    cem_ctx_diff=(True, True, True, 3, 4, 3, 12, 5),
    cem_ctx_abs=(True, False, True, 3, 4, 3, 12, 5),
    cem_plain=(False, False, True, 1, 1, 3, 12, 5),
    rs_ctx=(True, True, False, 4, 2, 2, 10, 5),
)


EVAL_SCENARIOS = dict(
    # name: (context, state_diff, use_cem, history_length, num_rollouts, max_path_length, horizon, test_total)
    eval_plain_cem=(False, False, True, 1, 3, 12, 5, 5),
    eval_ctx_cem_diff=(True, True, True, 3, 3, 12, 5, 5),
    eval_ctx_rs_abs=(True, False, False, 4, 2, 10, 5, 4),
)


class _Inert(types.ModuleType):
    """Stands in for a module the reference imports but never touches on this path."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return self


class _ProgBar:
    def __init__(self, *a, **k):
        pass

    def update(self, *a, **k):
        pass

    def stop(self):
        pass


def import_reference():
    sys.modules.setdefault("tensorflow", _Inert("tensorflow"))
    pp = types.ModuleType("pyprind")
    pp.ProgBar = _ProgBar
    sys.modules.setdefault("pyprind", pp)
    sys.path.insert(0, "/root/reference")
    from cadm.samplers.model_sample_processor import ModelSampleProcessor
    from cadm.samplers.sampler import Sampler
    return Sampler, ModelSampleProcessor


def run_eval(name):
    from cadm.samplers.utils import context_rollout_multi, rollout_multi
    from cadm.samplers.vectorized_env_executor import IterativeEnvExecutor
    context, state_diff, use_cem, K, m, T, h, total = EVAL_SCENARIOS[name]
    FakeEnv._copies = 0
    env = FakeEnv()
    policy = ScriptedPolicy(h, env.act_dim, use_cem)
    vec_env = IterativeEnvExecutor(env, m, T)
    fn = context_rollout_multi if context else rollout_multi
    avg = fn(vec_env, policy, False, num_rollouts=m, test_total=total, state_diff=state_diff, act_dim=env.act_dim,
             use_cem=use_cem, horizon=h, context=context, history_length=K)
    out = {"n_calls": np.int64(len(policy.calls)), "average": np.float64(avg)}
    for i, c in enumerate(policy.calls):
        for k, v in c.items():
            out[f"call{i}_{k}"] = v
    return out


def run(Sampler, Processor, name):
    context, state_diff, use_cem, K, F, m, T, h = SCENARIOS[name]
    FakeEnv._copies = 0
    env = FakeEnv()
    policy = ScriptedPolicy(h, env.act_dim, use_cem)
    sampler = Sampler(env=env, policy=policy, num_rollouts=m, max_path_length=T, n_parallel=1, use_cem=use_cem, horizon=h,
                      context=context, state_diff=state_diff, history_length=K)
    paths = sampler.obtain_samples(log=False)
    out = {"n_calls": np.int64(len(policy.calls)), "n_paths": np.int64(len(paths))}
    for i, c in enumerate(policy.calls):
        for k, v in c.items():
            out[f"call{i}_{k}"] = v
    for i, p in enumerate(paths):
        for k in ("observations", "actions", "rewards", "dones", "cp_obs", "cp_act"):
            out[f"path{i}_{k}"] = np.array(p[k], copy=True)
        out[f"path{i}_env_t"] = p["env_infos"]["t"]
    if use_cem:
        out["final_prev_sol"] = sampler.prev_sol.copy()
    samples = Processor(discount=0.99, max_path_length=T, context=context, future_length=F).process_samples(paths, log=False)
    for k, v in samples.items():
        out[f"samples_{k}"] = v
    for i, p in enumerate(paths):                           # process_samples pads short paths in place
        out[f"path{i}_len_after"] = np.int64(p["observations"].shape[0])
        out[f"path{i}_returns"] = p["returns"]
    return out


BASE_SCENARIOS = dict(base_policy=(False, 3, 6), base_random=(True, 2, 5))       # name: (random, num_rollouts, max_path_length)


def run_base(name):
    """cadm/samplers/base.py BaseSampler.obtain_samples: one environment, policy.get_action per step."""
    from cadm.samplers.base import BaseSampler
    random, m, T = BASE_SCENARIOS[name]
    FakeEnv._copies = 0
    env = FakeEnv(lengths=((4, 2, 30, 3, 30),))
    policy = ScriptedSinglePolicy(env.act_dim)
    sampler = BaseSampler(env, policy, m, T)
    paths = sampler.obtain_samples(log=False, random=random)
    out = {"n_calls": np.int64(len(policy.calls)), "n_paths": np.int64(len(paths)),
           "total_timesteps_sampled": np.int64(sampler.total_timesteps_sampled)}
    for i, c in enumerate(policy.calls):
        out[f"call{i}_obs"] = c
    for i, p in enumerate(paths):
        for k in ("observations", "actions", "rewards", "dones"):
            out[f"path{i}_{k}"] = np.array(p[k], copy=True)
        out[f"path{i}_env_t"] = p["env_infos"]["t"]
        if not random:
            out[f"path{i}_agent_s"] = p["agent_infos"]["s"]
    return out


def main():
    Sampler, Processor = import_reference()
    blob = {}
    for name in SCENARIOS:
        for k, v in run(Sampler, Processor, name).items():
            blob[f"{name}/{k}"] = v
    for name in EVAL_SCENARIOS:
        for k, v in run_eval(name).items():
            blob[f"{name}/{k}"] = v
    for name in BASE_SCENARIOS:
        for k, v in run_base(name).items():
            blob[f"{name}/{k}"] = v
    from cadm.policies.mpc_controller import MPCController
    blob["mpc_controller/dispatch"] = np.array(controller_dispatch_log(MPCController))
    path = os.path.join(HERE, "recorded", "sampler_golden.npz")
    np.savez_compressed(path, **blob)
    print(path, len(blob), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
