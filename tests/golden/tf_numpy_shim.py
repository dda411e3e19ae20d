"""A NumPy stand-in for the ~35 TensorFlow-1.15 functions the reference's planner graph is written in.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_reference_golden.py in the build container).  The reference builds
its planner as a TF1 graph inside plain Python loops (cadm/dynamics/core/utils.py:99-184, :373-561).  TensorFlow cannot be
installed here, but the graph-building code only composes a small set of shape / arithmetic ops.  This module gives each
of them its documented TF semantics on NumPy arrays, evaluated eagerly, so that calling the reference's UNMODIFIED builder
functions with arrays in place of placeholders executes the planner and returns its result.  What is the reference's own
is everything that matters for parity: the tile / transpose / reshape sequence that maps particles to ensemble members,
the way the context is attached, which observation the reward reads, the elite refit.  What is NOT exercised is TF's
kernels and TF's random generators: randomness is injected (the queues below), variables come from a table by name.

The whole graph runs in one float type chosen at install() -- float64 for the recordings, so that they are the exact
reference against which float32 implementations are measured; `tf.float32` inside the graph (casts of comparison
results in the reward functions) means that type.

install(dtype) registers the stand-in as `tensorflow` (plus inert stand-ins for baselines / gym / pyprind, which the import
chain touches but the planner does not use) in sys.modules.
"""
import sys
import types

import numpy as np


class Tape:
    """Everything the caller injects or wants recorded, reset per run."""

    def __init__(self):
        self.dtype = np.float64
        self.variables = {}          # name -> array, served by get_variable / Variable
        self.normal, self.truncated, self.uniform = [], [], []     # FIFO queues of arrays (or None = zeros)
        self.top_k = []              # (input, indices) of every tf.nn.top_k call
        self.argmax = []             # (input, indices) of every tf.argmax call
        self.served = []             # names of the variables handed out, in creation order
        self.scope = []              # variable_scope stack (names)
        self.placeholders = []       # FIFO queue of arrays served by tf.compat.v1.placeholder, in creation order
        self.served_scoped = []      # the same names with their scope path


TAPE = Tape()


def _ints(shape):
    return [int(s) for s in (shape.tolist() if isinstance(shape, np.ndarray) else shape)]


def _pop(queue, shape, what):
    if not queue:
        raise RuntimeError(f"the graph asked for more {what} draws than were injected")
    v = queue.pop(0)
    if v is None:
        return np.zeros(_ints(shape), TAPE.dtype)
    v = np.asarray(v)
    assert list(v.shape) == _ints(shape), (what, v.shape, _ints(shape))
    return v


# ---- shape ops
def reshape(x, shape, name=None):
    return np.reshape(x, _ints(shape))


def tile(x, multiples, name=None):
    return np.tile(x, _ints(multiples))


def transpose(x, perm=None, name=None):
    return np.transpose(x, perm)


def concat(values, axis, name=None):
    return np.concatenate(values, axis=axis)


def shape(x, name=None):
    return np.array(np.shape(x), dtype=np.int32)


def range_(start, limit=None, delta=1, dtype=None, name=None):
    return np.arange(start, limit, delta, dtype=np.int32) if limit is not None else np.arange(start, dtype=np.int32)


def gather(params, indices, axis=0, name=None):
    return np.take(params, indices, axis=axis)


def one_hot(indices, depth, dtype=None, name=None):
    return np.eye(int(depth), dtype=TAPE.dtype)[np.asarray(indices)]


# ---- arithmetic
def reduce_mean(x, axis=None, keepdims=False, name=None):
    return np.mean(x, axis=axis, keepdims=keepdims)


def reduce_sum(x, axis=None, keepdims=False, name=None):
    return np.sum(x, axis=axis, keepdims=keepdims)


def matmul(a, b, name=None):
    return np.matmul(a, b)


def softplus(x, name=None):
    x = np.asarray(x)
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def sigmoid(x, name=None):
    return 1.0 / (1.0 + np.exp(-x))


def relu(x, name=None):
    return np.maximum(x, 0)


def l2_loss(x, name=None):
    return 0.5 * np.sum(np.square(x))


def top_k(x, k=1, sorted=True, name=None):
    """Largest k along the last axis, descending; equal values keep the lower index first (TF's documented tie rule)."""
    x = np.asarray(x)
    idx = np.argsort(-x, axis=-1, kind="stable")[..., :k].astype(np.int32)
    TAPE.top_k.append((x.copy(), idx.copy()))
    return np.take_along_axis(x, idx, axis=-1), idx


def argmax(x, axis=None, output_type=np.int64, name=None):
    idx = np.argmax(x, axis=axis).astype(output_type)          # first maximum, as TF
    TAPE.argmax.append((np.array(x, copy=True), idx.copy()))
    return idx


# ---- randomness (injected)
def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    return mean + stddev * _pop(TAPE.normal, shape, "normal")


def random_truncated_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    """mean + stddev * z with z a standard normal truncated to [-2, 2] (TF redraws samples beyond two standard deviations)."""
    return mean + stddev * _pop(TAPE.truncated, shape, "truncated normal")


def random_uniform(shape, minval=0, maxval=None, dtype=None, seed=None, name=None):
    """The injected array is the draw itself: floats in [minval, maxval) or, for an integer dtype, the integers."""
    return _pop(TAPE.uniform, shape, "uniform")


# ---- variables (from a table, by name)
def _serve(name, initial=None, want_shape=None):
    TAPE.served.append(name)
    TAPE.served_scoped.append("/".join(TAPE.scope + [name]))
    scoped = [f"{sc}/{name}" for sc in reversed(TAPE.scope) if f"{sc}/{name}" in TAPE.variables]
    if scoped:                       # "<scope>/<name>" wins over the bare name (two models with equal variable names)
        v = np.asarray(TAPE.variables[scoped[0]], dtype=TAPE.dtype)
    elif name in TAPE.variables:
        v = np.asarray(TAPE.variables[name], dtype=TAPE.dtype)
    elif initial is not None:
        v = np.asarray(initial, dtype=TAPE.dtype)
    else:
        raise KeyError(f"no value injected for variable {name!r}")
    if want_shape is not None:
        assert list(v.shape) == _ints(want_shape), (name, v.shape, want_shape)
    return v


def Variable(initial_value=None, dtype=None, name=None, trainable=True):
    return _serve(name, initial=initial_value)


def get_variable(name, shape=None, initializer=None, dtype=None, trainable=True):
    return _serve(name, want_shape=shape)


class variable_scope:
    """tf.compat.v1.variable_scope as a plain name stack (no reuse rules: variables come from a table)."""

    def __init__(self, name, reuse=None, **k):
        self.name = name

    def __enter__(self):
        TAPE.scope.append(self.name)
        return self

    def __exit__(self, *a):
        TAPE.scope.pop()
        return False


def placeholder(dtype=None, shape=None, name=None):
    """The model constructors create placeholders and build the graph on them; here the graph is evaluated eagerly, so a
    placeholder IS its value: the next array of the injected queue (creation order), checked against the declared shape."""
    if not TAPE.placeholders:
        raise RuntimeError("the constructor created more placeholders than values were injected")
    v = np.asarray(TAPE.placeholders.pop(0), dtype=TAPE.dtype)
    if shape is not None:
        assert v.ndim == len(shape) and all(d is None or int(d) == n for d, n in zip(shape, v.shape)), (shape, v.shape)
    return v.view(_Placeholder)                      # hashable by identity: the models use placeholders as feed_dict keys


class _Placeholder(np.ndarray):
    """Identity-hashable view.  Arithmetic on it yields plain arrays / NumPy scalars (as on the plain arrays the builders are
    otherwise handed), so that `x += y` on a derived 0-d result rebinds instead of writing through -- TF tensors are
    immutable, and the model constructors rely on it (`recon_loss = self.mse_loss; recon_loss += ...`)."""
    __hash__ = object.__hash__

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        plain = tuple(np.asarray(i) if isinstance(i, _Placeholder) else i for i in inputs)
        return getattr(ufunc, method)(*plain, **kwargs)


class _Graph:
    def get_name_scope(self):
        return "/".join(TAPE.scope)


class _GraphKeys:
    TRAINABLE_VARIABLES = "trainable_variables"


def _initializer(*a, **k):
    return None


def _n(f):
    """A NumPy function under TF's calling convention (every TF op takes an optional name=)."""
    return lambda *a, name=None, **k: f(*a, **k)


class _Stub(types.ModuleType):
    """A module whose every attribute is an empty class (usable as a base class or called for an inert object)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _ProgBar:
    def __init__(self, *a, **k):
        pass

    def update(self, *a, **k):
        pass

    def stop(self):
        pass


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install(dtype=np.float64):
    """Register the stand-ins; returns the fake `tensorflow` module."""
    TAPE.dtype = dtype
    tf = _module(
        "tensorflow", float32=dtype, int32=np.int32, float64=np.float64, int64=np.int64,     # "float32" = the run's float type
        reshape=reshape, tile=tile, transpose=transpose, concat=concat, shape=shape, range=range_, gather=gather,
        one_hot=one_hot, reduce_mean=reduce_mean, reduce_sum=reduce_sum, square=_n(np.square), sqrt=_n(np.sqrt), exp=_n(np.exp),
        minimum=_n(np.minimum), maximum=_n(np.maximum), multiply=_n(np.multiply), matmul=matmul, sin=_n(np.sin), cos=_n(np.cos),
        abs=_n(np.abs),
        argmax=argmax, Variable=Variable, constant=lambda v, dtype=None, name=None: np.asarray(v, TAPE.dtype),
        cast=lambda x, dtype, name=None: np.asarray(x).astype(dtype), zeros_initializer=_initializer,
        constant_initializer=_initializer, truncated_normal_initializer=_initializer, identity=lambda x, name=None: x,
        where=_n(np.where), logical_and=_n(np.logical_and), logical_or=_n(np.logical_or), less=_n(np.less),
        greater=_n(np.greater), atan2=_n(np.arctan2), ones_like=_n(np.ones_like), zeros_like=_n(np.zeros_like),
        clip_by_value=_n(np.clip), sign=_n(np.sign), pow=_n(np.power))
    tf.math = _module("tensorflow.math", log=_n(np.log), exp=_n(np.exp), atan2=_n(np.arctan2), floormod=_n(np.mod),
                      mod=_n(np.mod), sin=_n(np.sin), cos=_n(np.cos), square=_n(np.square), abs=_n(np.abs))
    tf.nn = _module("tensorflow.nn", softplus=softplus, top_k=top_k, l2_loss=l2_loss, relu=relu, tanh=_n(np.tanh), sigmoid=sigmoid,
                    softmax=None, swish=lambda x: x * sigmoid(x))
    tf.tanh, tf.sigmoid = _n(np.tanh), sigmoid
    tf.random = _module("tensorflow.random", normal=random_normal, truncated_normal=random_truncated_normal,
                        uniform=random_uniform)
    tf.contrib = _module("tensorflow.contrib", layers=_module("tensorflow.contrib.layers", xavier_initializer=_initializer))
    v1 = _module("tensorflow.compat.v1", get_variable=get_variable, Variable=Variable, placeholder=placeholder,
                 variable_scope=variable_scope, AUTO_REUSE=object(), trainable_variables=lambda: list(TAPE.served_scoped),
                 get_default_graph=lambda: _Graph(), get_collection=lambda key, scope=None: [], GraphKeys=_GraphKeys)
    v1.train = _Stub("tensorflow.compat.v1.train")
    tf.compat = _module("tensorflow.compat", v1=v1)
    tf.train = _Stub("tensorflow.train")
    sys.modules["tensorflow"] = tf
    for name in ("baselines", "baselines.common", "baselines.common.distributions", "gym", "gym.utils", "gym.spaces", "gym.error",
                 "gym.envs", "gym.envs.mujoco", "gym.envs.mujoco.mujoco_env", "gym.envs.classic_control",
                 "gym.envs.classic_control.cartpole", "gym.envs.classic_control.pendulum"):
        sys.modules[name] = _Stub(name)
    for name in list(sys.modules):                               # wire submodules as attributes of their parents
        if "." in name and isinstance(sys.modules[name], _Stub):
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])
    sys.modules["pyprind"] = _module("pyprind", ProgBar=_ProgBar)
    return tf
