"""Stand-ins that let the reference's own source files run WITHOUT jax / flax (neither is installed in this image).

TEST-FIXTURE TOOLING ONLY: used by the golden generators in this directory, inside the build container, to execute the
reference's files where they lie under /root/reference (loaded by path, nothing copied).  What is replaced:

  jax.numpy            numpy, with every float computed in float64 (jnp.float32 -> float64)
  jax.jit / jax.vmap / jax.ops.segment_sum / jax.nn.{silu, softplus} / jax.nn.initializers.constant
                       identity / Python loop / np.add.at / closed forms
  jax.value_and_grad   central finite differences in float64 (used for the reference's stress function only)
  flax.linen           an APPLY-ONLY re-implementation of the few pieces the SO3krates path uses:
                       Module (dataclass fields, setup, compact, flax's naming rules: attribute names for sub-modules
                       assigned in setup or passed as fields -- `layers_0`, `filter_fn` -- and `ClassName_<k>` in call
                       order for sub-modules created inside a compact method), Dense, Embed, LayerNorm (flax defaults:
                       eps 1e-6, fast variance), vmap over the leading parameter axis, sow (ignored), param (lookup).
                       Parameters are never initialised here: `Module.apply({'params': tree}, ...)` looks every leaf up
                       by the path flax would use, so a tree in the reference's checkpoint layout is required -- which
                       also checks the flat-name <-> flax-path map of mlff_b200/params.py.
The arithmetic of the model itself -- every line of embed.py, so3krates_layer.py, mlp.py, observable.py, stacknet.py,
spherical.py, radial.py, contract.py, mask.py -- is the reference's."""
import dataclasses
import importlib.util
import io
import os
import pickle
import sys
import types
from typing import Optional

import numpy as np

REF = '/root/reference'
_stack = []                                    # modules whose __call__ is running (innermost last)


class _F64(np.float64):
    def __new__(cls, x=0.0):
        return np.float64(np.nan if x is None else x)      # jnp.float32(None) is nan in jax


def _segment_sum(data, segment_ids, num_segments=None):
    data, ids = np.asarray(data), np.asarray(segment_ids)
    out = np.zeros((int(num_segments if num_segments is not None else ids.max() + 1),) + data.shape[1:], data.dtype)
    np.add.at(out, ids, data)                  # negative (padding) ids wrap like jax's; their data is masked to zero
    return out


def _vmap(fn, *a, **k):
    return lambda *xs: np.stack([fn(*row) for row in zip(*xs)])


def _value_and_grad(fn, argnums=0, has_aux=False, allow_int=False, h=1e-4):
    """jax.value_and_grad by central finite differences in float64 (error ~h^2): the gradient w.r.t. every float array
    among the selected arguments (dict arguments leaf by leaf; integer leaves get None)."""
    single = isinstance(argnums, int)
    nums = (argnums,) if single else tuple(argnums)

    def wrapped(*args):
        value = fn(*args)

        def fd(k, get, put):
            x0 = np.asarray(get(args[k]), np.float64)
            g = np.zeros_like(x0)
            for idx in np.ndindex(x0.shape):
                vals = []
                for sgn in (1.0, -1.0):
                    xp = x0.copy()
                    xp[idx] += sgn * h
                    a = list(args)
                    a[k] = put(args[k], xp)
                    vals.append(float(np.asarray(fn(*a)).reshape(())))
                g[idx] = (vals[0] - vals[1]) / (2 * h)
            return g

        grads = []
        for k in nums:
            if isinstance(args[k], dict):
                gk = {}
                for key, leaf in args[k].items():
                    if isinstance(leaf, np.ndarray) and np.issubdtype(leaf.dtype, np.floating):
                        gk[key] = fd(k, lambda d, key=key: d[key], lambda d, v, key=key: {**d, key: v})
                    else:
                        gk[key] = None
                grads.append(gk)
            else:
                grads.append(fd(k, lambda a: a, lambda a, v: v))
        return value, (grads[0] if single else tuple(grads))

    return wrapped


def _silu(x):
    return x / (1.0 + np.exp(-x))


def _softplus(x):
    return np.logaddexp(x, 0.0)


class _JArray(np.ndarray):
    """numpy array that accepts jax's argument-less `.reshape()` (-> scalar shape) on the outputs of apply()."""

    def reshape(self, *shape, **kw):
        return np.asarray(self).reshape(*(shape if shape else ((),)), **kw)


# ---- flax.linen, apply-only ---------------------------------------------------------------------------------------
class Module:
    name: Optional[str] = dataclasses.field(default=None, kw_only=True)

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        dataclasses.dataclass(cls, eq=False, repr=False)
        call = cls.__dict__.get('__call__')
        if call is not None:
            def bound_call(self, *args, __call=call, **kwargs):
                self._bind()
                self.__dict__['_counters'] = {}          # a compact method re-creates the same sub-module names per call
                _stack.append(self)
                try:
                    return __call(self, *args, **kwargs)
                finally:
                    _stack.pop()
            cls.__call__ = bound_call

    def __post_init__(self):
        d = self.__dict__
        d['_params'], d['_parent'], d['_given_name'], d['_counters'], d['_setup_done'] = None, None, None, {}, False
        for f in dataclasses.fields(self):     # sub-modules passed as fields are named after the field (flax)
            v = getattr(self, f.name)
            if isinstance(v, Module):
                v._adopt(self, f.name)
            elif isinstance(v, (list, tuple)):
                for k, m in enumerate(v):
                    if isinstance(m, Module):
                        m._adopt(self, f'{f.name}_{k}')

    def __setattr__(self, key, value):         # sub-modules assigned in setup() are named after the attribute (flax)
        if isinstance(value, Module) and '_params' in self.__dict__:
            value._adopt(self, key)
        object.__setattr__(self, key, value)

    def _adopt(self, parent, name):
        self.__dict__['_parent'], self.__dict__['_given_name'] = parent, name

    def _bind(self):
        d = self.__dict__
        if d['_params'] is None:
            parent = d['_parent'] if d['_parent'] is not None else (_stack[-1] if _stack else None)
            if parent is None:
                raise RuntimeError('module called outside apply()')
            name = self.name or d['_given_name']
            if name is None:                   # created inside a compact method: ClassName_<k> in creation order
                k = parent._counters.get(type(self).__name__, 0)
                parent._counters[type(self).__name__] = k + 1
                name = f'{type(self).__name__}_{k}'
            parent._bind()
            d['_params'] = parent._params.get(name, {}) if isinstance(parent._params, dict) else {}
            d['_path'] = getattr(parent, '_path', '') + '/' + name
        if not d['_setup_done']:
            d['_setup_done'] = True
            setup = getattr(self, 'setup', None)
            if setup is not None:
                setup()

    def param(self, name, init_fn=None, *init_args):
        self._bind()
        if name not in self._params:
            raise KeyError(f'parameter {self._path}/{name} is missing from the tree')
        return np.asarray(self._params[name], np.float64)

    def sow(self, *a, **k):
        return None

    def apply(self, variables, *args, **kwargs):
        d = self.__dict__
        d['_params'], d['_path'] = variables['params'] if 'params' in variables else variables, ''
        out = self(*args, **kwargs)
        if isinstance(out, dict):
            out = {k: np.asarray(v).view(_JArray) if isinstance(v, np.ndarray) else v for k, v in out.items()}
        return out


Module = dataclasses.dataclass(eq=False, repr=False)(Module)


class Dense(Module):
    features: int
    use_bias: bool = True

    def __call__(self, x):
        y = x @ self.param('kernel')
        assert y.shape[-1] == self.features
        return y + self.param('bias') if self.use_bias else y


class Embed(Module):
    num_embeddings: int
    features: int

    def __call__(self, z):
        emb = self.param('embedding')
        assert emb.shape == (self.num_embeddings, self.features)
        return emb[z]


class LayerNorm(Module):
    epsilon: float = 1e-6

    def __call__(self, x):
        mean = x.mean(-1, keepdims=True)
        var = np.maximum(0.0, (x * x).mean(-1, keepdims=True) - mean * mean)       # use_fast_variance=True
        return (x - mean) / np.sqrt(var + self.epsilon) * self.param('scale') + self.param('bias')


def vmap(cls, in_axes=0, out_axes=0, axis_size=None, variable_axes=None, split_rngs=None):
    """nn.vmap(cls, ...) with variable_axes={'params': 0}: the returned class holds the parameters of `axis_size` copies
    of `cls` stacked along a leading axis and maps the call over it (flax names the class 'Vmap' + cls.__name__)."""
    assert variable_axes == {'params': 0}

    def take(tree, h):
        return {k: take(v, h) for k, v in tree.items()} if isinstance(tree, dict) else np.asarray(tree)[h]

    class Vmapped(Module):
        def __init__(self, *args, name=None, **kwargs):
            object.__setattr__(self, 'name', name)
            self.__post_init__()
            self.__dict__['_ctor'] = (args, kwargs)

        def __call__(self, *inputs):
            axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(inputs)
            outs = []
            for h in range(axis_size):
                inner = cls(*self._ctor[0], **self._ctor[1])
                inner.__dict__['_params'], inner.__dict__['_path'] = take(self._params, h), f'{self._path}[{h}]'
                outs.append(inner(*[x if a is None else np.take(x, h, axis=a) for x, a in zip(inputs, axes)]))
            return np.stack(outs, axis=out_axes)

    Vmapped.__name__ = Vmapped.__qualname__ = 'Vmap' + cls.__name__
    return Vmapped


def install():
    """Put the stand-ins into sys.modules and declare the reference's packages WITHOUT running their __init__ files."""
    jnp = types.ModuleType('jax.numpy')
    jnp.__dict__.update({k: v for k, v in np.__dict__.items() if not k.startswith('__')})
    jnp.float32, jnp.ndarray = _F64, np.ndarray
    jnp.repeat = lambda a, repeats, axis=None, total_repeat_length=None: np.repeat(a, repeats, axis=axis)   # jax-only kwarg
    jax = types.ModuleType('jax')
    jax.numpy, jax.jit, jax.vmap = jnp, (lambda f=None, **k: f), _vmap
    jax.value_and_grad = _value_and_grad
    jax.ops = types.ModuleType('jax.ops')
    jax.ops.segment_sum = _segment_sum
    jax.nn = types.ModuleType('jax.nn')
    jax.nn.silu, jax.nn.softplus = _silu, _softplus
    jax.nn.initializers = types.ModuleType('jax.nn.initializers')
    jax.nn.initializers.constant = lambda v: (lambda *a, **k: v)
    linen = types.ModuleType('flax.linen')
    linen.Module, linen.compact, linen.Dense, linen.Embed, linen.LayerNorm, linen.vmap = \
        Module, (lambda f: f), Dense, Embed, LayerNorm, vmap
    flax = types.ModuleType('flax')
    flax.linen = linen
    flax.core = types.ModuleType('flax.core')
    flax.core.frozen_dict = types.ModuleType('flax.core.frozen_dict')
    flax.core.frozen_dict.FrozenDict = dict
    sys.modules.update({'flax.core': flax.core, 'flax.core.frozen_dict': flax.core.frozen_dict})
    pkg = types.ModuleType('pkg_resources')

    def resource_stream(name, fn):
        path = os.path.join(REF, *name.split('.')[:-1], fn)
        if not os.path.exists(path):      # u_matrix.pickle (SymmetricContraction, another model family) is not in the checkout
            return io.BytesIO(pickle.dumps({}))
        return open(path, 'rb')

    pkg.resource_stream = resource_stream
    sys.modules.update({'jax': jax, 'jax.numpy': jnp, 'jax.ops': jax.ops, 'jax.nn': jax.nn,
                        'jax.nn.initializers': jax.nn.initializers, 'flax': flax, 'flax.linen': linen,
                        'pkg_resources': pkg})
    for name in ('mlff', 'mlff.masking', 'mlff.basis_function', 'mlff.cutoff_function', 'mlff.sph_ops', 'mlff.nn',
                 'mlff.nn.base', 'mlff.nn.embed', 'mlff.nn.mlp', 'mlff.nn.layer', 'mlff.nn.observable',
                 'mlff.nn.stacknet', 'mlff.nn.representation', 'mlff.nn.activation_function', 'mlff.properties',
                 'mlff.io'):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m


def load(name):
    """Execute one reference source file, by path, as module `name`."""
    path = os.path.join(REF, *name.split('.')) + '.py'
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    parent, _, leaf = name.rpartition('.')
    setattr(sys.modules[parent], leaf, mod)
    return mod


def export(package, module, *names):
    """What the package's __init__ re-exports (`from .x import a, b`)."""
    for n in names:
        setattr(sys.modules[package], n, getattr(module, n))
