"""A minimal *eager* stand-in for the Keras 2.2 API, used ONLY by make_golden.py.

Keras/TensorFlow cannot be installed in the build container (no network), so the
reference's model-building code (models/*.py, utils.build_network,
learn_image_embeddings.cls_model, sgdr_callback.SGDR, utils.nn_accuracy, ...)
cannot run as-is.  This stub lets that *unmodified reference code* execute: every
`Layer.__call__` computes its result immediately with the per-layer semantics of
`oracle/nn.py`, pulling weights by Keras layer name from `CTX.weights`.  What the
resulting fixtures pin is therefore the reference's own graph topology, layer
order, layer names, hyper-parameters and formula composition -- NOT the numerics
of a Keras layer (those remain "parity unpinned", see oracle/__init__.py).

This file is test tooling that runs in the build container only; it is never
imported by the product or by the GPU-side tests.
"""
import importlib.abc
import importlib.machinery
import re
import sys
import types

import numpy as np
import torch

from oracle import nn as onn


class _Ctx:
    def __init__(self):
        self.reset()

    def reset(self, x=None, weights=None, training=True):
        self.x_input = x
        self.weights = weights or {}
        self.training = training
        self.counters = {}
        self.trace = []
        self.used = set()

    def auto_name(self, cls_name):
        s = re.sub('(.)([A-Z][a-z0-9]+)', r'\1_\2', cls_name)
        s = re.sub('([a-z])([A-Z])', r'\1_\2', s).lower()
        self.counters[s] = self.counters.get(s, 0) + 1
        return '%s_%d' % (s, self.counters[s])

    def w(self, name):
        self.used.add(name)
        return self.weights[name]


CTX = _Ctx()


class _Any:
    """Permissive placeholder for every Keras symbol the traced code never really uses."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, name):
        return _Any()


class L2:
    def __init__(self, l2):
        self.l2 = float(l2)


def _l2_of(reg):
    return reg.l2 if isinstance(reg, L2) else 0.0


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _act(name, x):
    if name is None or name == 'linear':
        return x
    if name == 'relu':
        return torch.relu(x)
    if name == 'softmax':
        return torch.softmax(x, dim=-1)
    raise NotImplementedError(name)


class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name if name is not None else CTX.auto_name(type(self).__name__)
        self.trainable = True

    def config(self):
        return {}

    def __call__(self, inputs):
        out = self.call(inputs)
        ins = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        rec = {'class': type(self).__name__, 'name': self.name,
               'in_shapes': [list(t.shape[1:]) for t in ins],
               'out_shape': list(out.shape[1:])}
        rec.update(self.config())
        CTX.trace.append(rec)
        self.output = out
        return out


class InputSpec:
    def __init__(self, *a, **k):
        pass


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', activation=None,
                 use_bias=True, kernel_initializer='glorot_uniform', kernel_regularizer=None,
                 input_shape=None, name=None, **kw):
        super().__init__(name=name)
        self.filters, self.kernel_size, self.strides = filters, _pair(kernel_size), _pair(strides)
        self.padding, self.activation, self.use_bias = padding, activation, use_bias
        self.kernel_initializer, self.kernel_regularizer = kernel_initializer, kernel_regularizer

    def config(self):
        return {'filters': self.filters, 'kernel_size': list(self.kernel_size), 'strides': list(self.strides),
                'padding': self.padding, 'activation': self.activation, 'use_bias': self.use_bias,
                'kernel_initializer': self.kernel_initializer, 'l2': _l2_of(self.kernel_regularizer)}

    def call(self, x):
        k = CTX.w(self.name + '/kernel')
        assert tuple(k.shape) == (self.kernel_size[0], self.kernel_size[1], x.shape[-1], self.filters), \
            (self.name, tuple(k.shape), x.shape, self.filters)
        b = CTX.w(self.name + '/bias') if self.use_bias else None
        assert self.strides[0] == self.strides[1]
        return _act(self.activation, onn.conv2d(x, k, b, self.strides[0], self.padding))


Convolution2D = Conv2D


class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, gamma_initializer='ones', name=None, **kw):
        super().__init__(name=name)
        self.axis, self.momentum, self.epsilon, self.gamma_initializer = axis, momentum, epsilon, gamma_initializer

    def config(self):
        return {'axis': self.axis, 'momentum': self.momentum, 'epsilon': self.epsilon,
                'gamma_initializer': self.gamma_initializer}

    def call(self, x):
        assert self.axis in (-1, x.dim() - 1)
        g, b = CTX.w(self.name + '/gamma'), CTX.w(self.name + '/beta')
        if CTX.training:
            return onn.batchnorm_train(x, g, b, self.epsilon)[0]
        return onn.batchnorm_infer(x, g, b, CTX.w(self.name + '/moving_mean'),
                                   CTX.w(self.name + '/moving_variance'), self.epsilon)


class Activation(Layer):
    def __init__(self, activation, name=None, **kw):
        super().__init__(name=name)
        self.activation = activation

    def config(self):
        return {'activation': self.activation}

    def call(self, x):
        return _act(self.activation, x)


class AveragePooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding='valid', name=None, **kw):
        super().__init__(name=name)
        self.pool_size = _pair(pool_size)
        self.strides = self.pool_size if strides is None else _pair(strides)
        assert padding == 'valid' and self.strides == self.pool_size

    def config(self):
        return {'pool_size': list(self.pool_size)}

    def call(self, x):
        return onn.avgpool2(x, self.pool_size[0])


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding='valid', name=None, **kw):
        super().__init__(name=name)
        self.pool_size = _pair(pool_size)
        self.strides = self.pool_size if strides is None else _pair(strides)

    def config(self):
        return {'pool_size': list(self.pool_size), 'strides': list(self.strides)}

    def call(self, x):
        return onn.maxpool(x, self.pool_size[0], self.strides[0])


class GlobalAveragePooling2D(Layer):
    def call(self, x):
        return onn.gap(x)


GlobalAvgPool2D = GlobalAveragePooling2D


class Flatten(Layer):
    def call(self, x):
        return x.reshape(x.shape[0], -1)


class Dropout(Layer):
    def __init__(self, rate, name=None, **kw):
        super().__init__(name=name)

    def call(self, x):
        raise NotImplementedError('dropout is off in every traced configuration')


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_regularizer=None, name=None, **kw):
        super().__init__(name=name)
        self.units, self.activation, self.use_bias, self.kernel_regularizer = units, activation, use_bias, kernel_regularizer

    def config(self):
        return {'units': self.units, 'activation': self.activation, 'use_bias': self.use_bias,
                'l2': _l2_of(self.kernel_regularizer)}

    def call(self, x):
        k = CTX.w(self.name + '/kernel')
        assert tuple(k.shape) == (x.shape[-1], self.units), (self.name, tuple(k.shape), x.shape)
        return _act(self.activation, onn.dense(x, k, CTX.w(self.name + '/bias') if self.use_bias else None))


class Add(Layer):
    def call(self, xs):
        out = xs[0]
        for t in xs[1:]:
            out = out + t
        return out


def add(xs, **kw):
    return Add(**kw)(xs)


class Lambda(Layer):
    def __init__(self, function, name=None, **kw):
        super().__init__(name=name)
        self.function = function

    def config(self):
        return {'function': getattr(self.function, '__name__', '?')}

    def call(self, x):
        return self.function(x)


def Input(shape=None, tensor=None, **kw):
    CTX.trace.append({'class': 'InputLayer', 'name': 'input', 'shape': list(shape) if shape else None})
    return CTX.x_input if tensor is None else tensor


class Model:
    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = outputs if isinstance(outputs, (list, tuple)) else [outputs]
        self.output = outputs
        self.name = name
        self.optimizer = None

    def load_weights(self, *a, **k):
        raise NotImplementedError


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        x = CTX.x_input
        CTX.trace.append({'class': 'InputLayer', 'name': 'input', 'shape': None})
        for l in layers:
            x = l(x)
        super().__init__(CTX.x_input, x, name)


class Callback:
    def __init__(self):
        self.model = None


class _Var:
    def __init__(self, v):
        self.value = v


# ---- keras.backend -------------------------------------------------------------------------------
class _TFnn:
    @staticmethod
    def l2_normalize(x, axis):
        assert axis == -1
        return onn.l2norm(x)

    @staticmethod
    def top_k(x, k, sorted=False):
        return torch.topk(x, k, dim=-1)


class _TF:
    nn = _TFnn()

    @staticmethod
    def pad(x, pattern):
        flat = []
        for lo, hi in reversed(pattern):
            flat += [int(lo), int(hi)]
        return torch.nn.functional.pad(x, flat)


def _t(x):
    return x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x), dtype=torch.float64)


def _make_backend():
    K = types.ModuleType('keras.backend')
    K.tf = _TF()
    K.image_data_format = lambda: 'channels_last'
    K.backend = lambda: 'tensorflow'
    K.floatx = lambda: 'float64'
    K.is_keras_tensor = lambda t: True
    K.normalize_data_format = lambda v: 'channels_last' if v is None else v
    K.sum = lambda x, axis=None, keepdims=False: torch.sum(x, dim=axis, keepdim=keepdims)
    K.square = lambda x: x * x
    K.sqrt = torch.sqrt
    K.abs = torch.abs
    K.dot = lambda a, b: a @ b
    K.max = lambda x, axis=None: torch.max(x, dim=axis).values
    K.min = lambda x, axis=None: torch.min(x, dim=axis).values
    K.less = lambda a, b: a < b
    K.any = lambda x, axis=None: torch.any(x, dim=axis)
    K.cast = lambda x, dtype: x.to(torch.float64)
    K.constant = lambda v: _t(v)
    K.relu = torch.relu
    K.set_value = lambda var, v: setattr(var, 'value', float(v))
    K.get_value = lambda var: var.value
    K.set_session = lambda *a, **k: None
    return K


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return type(name, (_Any,), {})


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ('keras', 'tensorflow', 'keras_applications', 'keras_preprocessing', 'keras_resnet',
                'numexpr', 'h5py', 'matplotlib', 'subpixel', 'tensorflow_backend')

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        if spec.name in sys.modules:
            return sys.modules[spec.name]
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def install():
    """Registers the stub `keras` package (and permissive placeholders for every other absent
    dependency of the reference) in sys.modules."""
    if 'keras' in sys.modules and getattr(sys.modules['keras'], '_se_stub', False):
        return
    sys.meta_path.insert(0, _Finder())
    keras = _StubModule('keras')
    keras.__path__ = []
    keras._se_stub = True
    K = _make_backend()
    layers = _StubModule('keras.layers')
    layers.__path__ = []
    for cls in (Conv2D, BatchNormalization, Activation, AveragePooling2D, MaxPooling2D,
                GlobalAveragePooling2D, Flatten, Dropout, Dense, Add, Lambda, Layer):
        setattr(layers, cls.__name__, cls)
    layers.Convolution2D = Conv2D
    layers.GlobalAvgPool2D = GlobalAveragePooling2D
    layers.Input = Input
    layers.add = add
    layers.InputSpec = InputSpec
    conv_mod = _StubModule('keras.layers.convolutional'); conv_mod.Convolution2D = Conv2D; conv_mod.Conv2D = Conv2D
    norm_mod = _StubModule('keras.layers.normalization'); norm_mod.BatchNormalization = BatchNormalization
    regs = _StubModule('keras.regularizers'); regs.l2 = L2
    models = _StubModule('keras.models'); models.Model = Model; models.Sequential = Sequential
    engine = _StubModule('keras.engine'); engine.__path__ = []; engine.Layer = Layer; engine.InputSpec = InputSpec
    topo = _StubModule('keras.engine.topology'); topo.get_source_inputs = lambda t: t
    utils = _StubModule('keras.utils'); utils.__path__ = []
    conv_utils = _StubModule('keras.utils.conv_utils')
    conv_utils.normalize_tuple = lambda v, n, name: (v,) * n if isinstance(v, int) else tuple(v)
    conv_utils.normalize_data_format = K.normalize_data_format
    utils.conv_utils = conv_utils
    utils.to_categorical = lambda y, n: np.eye(n)[np.asarray(y)]
    callbacks = _StubModule('keras.callbacks'); callbacks.Callback = Callback
    keras.backend, keras.layers, keras.regularizers, keras.models = K, layers, regs, models
    keras.engine, keras.utils, keras.callbacks = engine, utils, callbacks
    mods = {'keras': keras, 'keras.backend': K, 'keras.layers': layers, 'keras.layers.convolutional': conv_mod,
            'keras.layers.normalization': norm_mod, 'keras.regularizers': regs, 'keras.models': models,
            'keras.engine': engine, 'keras.engine.topology': topo, 'keras.utils': utils,
            'keras.utils.conv_utils': conv_utils, 'keras.callbacks': callbacks}
    sys.modules.update(mods)
    # numexpr: evaluate_retrieval.py:62 uses ne.evaluate('A + B - 2 * C', {...}) -- elementwise fp32
    ne = _StubModule('numexpr')
    ne.evaluate = lambda expr, local_dict: eval(expr, {}, local_dict)
    sys.modules['numexpr'] = ne
    # datasets: host input pipeline (Keras ImageDataGenerator), out of scope; only the import must succeed
    ds = _StubModule('datasets'); ds.get_data_generator = lambda *a, **k: None
    sys.modules['datasets'] = ds
