"""MPCController -- the public planning surface (same method names, arguments and return shapes as
cadm/policies/mpc_controller.py:7-90).

The controller holds no planning logic of its own: it forwards a request to `dynamics_model.get_action`, whose
positional signature depends on two switches fixed at construction -- whether a context encoder is attached
(`context`: cp_obs, cp_act travel with the observations) and which planner the engine was built for (`use_cem`:
init_mean, init_var are the warm start; random shooting takes none).  `_plan` assembles that argument list in one place;
every public method is a thin view of it.  The reference's own controller works unchanged on the cadm_b200 dynamics
models too; this one exists so the package runs without the reference (and its TensorFlow imports) on the path.
tests/test_samplers.py checks the dispatch against a recording of the reference class.
"""
import numpy as np


def _innermost(env):
    """Strip `wrapped_env` layers (normalisation wrappers etc.)."""
    while hasattr(env, "wrapped_env"):
        env = env.wrapped_env
    return env


class MPCController:
    vectorized = True            # a property in the reference (:40-42); plans for all environments in one call

    def __init__(self, name, env, dynamics_model, reward_model=None, discount=1, use_cem=False, n_candidates=1024,
                 horizon=10, num_rollouts=10, context=False):
        base = _innermost(env)
        if not hasattr(base, "reward"):
            raise AssertionError("env must have a reward function")        # same failure type as mpc_controller.py:34
        self.name, self.env, self.unwrapped_env = name, env, base
        self.action_space = getattr(env, "action_space", None)
        self.dynamics_model, self.reward_model = dynamics_model, reward_model
        self.use_cem, self.context = use_cem, context
        self.n_candidates, self.horizon, self.discount = n_candidates, horizon, discount

    # ---------------------------------------------------------------- the one place that talks to the dynamics model
    def _plan(self, observations, history=(), warm_start=()):
        """history = (cp_obs, cp_act) or (); warm_start = (init_mean, init_var) or ().  Each group is passed only if the
        controller was configured for it, in the order get_action of the dynamics models expects (context first)."""
        extra = (tuple(history) if self.context else ()) + (tuple(warm_start) if self.use_cem else ())
        return self.dynamics_model.get_action(observations, *extra)

    def get_cem_gpu_action(self, observations, init_mean, init_var, cp_obs=None, cp_act=None):
        """:84-90 -- a CEM plan [m, h, A]; the history is ignored unless the controller has a context."""
        history = (cp_obs, cp_act) if self.context else ()
        return self.dynamics_model.get_action(observations, *history, init_mean, init_var)

    def get_rs_gpu_action(self, observations, cp_obs=None, cp_act=None):
        """:77-82 -- the first action of the best random-shooting candidate [m, A]."""
        history = (cp_obs, cp_act) if self.context else ()
        return self.dynamics_model.get_action(observations, *history)

    # ---------------------------------------------------------------- public surface
    def get_actions(self, observations, cp_obs=None, cp_act=None, init_mean=None, init_var=None):
        """:55-69 -- what Sampler.obtain_samples and the evaluation rollouts call once per environment step."""
        return self._plan(observations, (cp_obs, cp_act), (init_mean, init_var)), {}

    def get_action(self, observation, init_mean=None, init_var=None):
        """:43-53 -- single-observation form; never carries a history (the reference calls the planner without
        cp_obs / cp_act here, so a context model fails in get_action exactly as it does there)."""
        batch = observation[None] if observation.ndim == 1 else observation
        planner = self.get_cem_gpu_action if self.use_cem else self.get_rs_gpu_action
        args = (init_mean, init_var) if self.use_cem else ()
        return planner(batch, *args), {}

    def get_random_action(self, n):
        """:71-76 -- uniform exploration actions for the initial random data collection."""
        space = self.unwrapped_env.action_space
        if len(space.shape) == 0:
            return np.random.randint(space.n, size=n)
        box = self.action_space
        return np.random.uniform(low=box.low, high=box.high, size=(n,) + box.low.shape)

    def reset(self, dones=None):
        """Stateless: warm starts and histories live in the sampler (or in a PlannerSession)."""
