"""Architectures of the reference restated as plain functions over a parameter dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each `build_*` returns an `OracleModel` whose `params` maps Keras-style weight
names ('<layer>/kernel', '/bias', '/gamma', '/beta', '/moving_mean',
'/moving_variance') to float64 torch tensors in Keras layouts (HWIO conv
kernels, (in,out) dense kernels), and whose `forward(x, training)` follows the
layer order of the cited reference file line by line.

  plainnet      -> models/plainnet.py:5-78
  small_resnet  -> models/cifar_resnet.py:69-257
  wrn           -> models/wide_residual_network.py:8-101
  resnet50      -> keras.applications.ResNet50 (un-vendored) + utils.py:237-243
  dispatch      -> utils.py:130-276 (`build_network`)
"""
from collections import OrderedDict

import torch

from . import nn


class OracleModel:
    def __init__(self, name):
        self.name = name
        self.params = OrderedDict()      # every weight, trainable or not
        self.trainable = []              # names, in creation order
        self.l2 = {}                     # name -> lambda (kernel_regularizer=l2(lambda))
        self.bn_cfg = {}                 # layer -> (momentum, eps)
        self._fwd = None
        self.out_layer = None

    # ---- parameter creation helpers (Keras defaults: SURVEY.md Appendix A.1/A.2/A.4)
    def add_conv(self, layer, kh, kw, cin, cout, gen, use_bias=True, l2=0.0, init='glorot_uniform'):
        shape = (kh, kw, cin, cout)
        k = nn.glorot_uniform(shape, gen) if init == 'glorot_uniform' else nn.he_normal(shape, gen)
        self.params[layer + '/kernel'] = k
        self.trainable.append(layer + '/kernel')
        if l2:
            self.l2[layer + '/kernel'] = l2
        if use_bias:
            self.params[layer + '/bias'] = torch.zeros(cout, dtype=torch.float64)
            self.trainable.append(layer + '/bias')

    def add_dense(self, layer, cin, cout, gen, l2=0.0):
        self.params[layer + '/kernel'] = nn.glorot_uniform((cin, cout), gen)
        self.trainable.append(layer + '/kernel')
        if l2:
            self.l2[layer + '/kernel'] = l2
        self.params[layer + '/bias'] = torch.zeros(cout, dtype=torch.float64)
        self.trainable.append(layer + '/bias')

    def add_bn(self, layer, c, gen, momentum=0.99, eps=1e-3, gamma_init='ones'):
        if gamma_init == 'ones':
            g = torch.ones(c, dtype=torch.float64)
        else:  # Keras 'uniform' = RandomUniform(-0.05, 0.05)  (wide_residual_network.py:14)
            g = (torch.rand(c, generator=gen, dtype=torch.float64) * 2 - 1) * 0.05
        self.params[layer + '/gamma'] = g
        self.params[layer + '/beta'] = torch.zeros(c, dtype=torch.float64)
        self.params[layer + '/moving_mean'] = torch.zeros(c, dtype=torch.float64)
        self.params[layer + '/moving_variance'] = torch.ones(c, dtype=torch.float64)
        self.trainable += [layer + '/gamma', layer + '/beta']
        self.bn_cfg[layer] = (momentum, eps)

    # ---- layer application helpers
    def conv(self, p, layer, x, stride=1, padding='same'):
        return nn.conv2d(x, p[layer + '/kernel'], p.get(layer + '/bias'), stride, padding)

    def dense(self, p, layer, x):
        return nn.dense(x, p[layer + '/kernel'], p.get(layer + '/bias'))

    def bn(self, p, layer, x, training, updates):
        momentum, eps = self.bn_cfg[layer]
        if training:
            y, mean, var = nn.batchnorm_train(x, p[layer + '/gamma'], p[layer + '/beta'], eps)
            if updates is not None:
                count = x.numel() // x.shape[-1]
                updates[layer + '/moving_mean'] = nn.moving_update(
                    p[layer + '/moving_mean'], mean.detach(), momentum)
                updates[layer + '/moving_variance'] = nn.moving_update(
                    p[layer + '/moving_variance'], nn.unbiased_var(var.detach(), count, eps), momentum)
            return y
        return nn.batchnorm_infer(x, p[layer + '/gamma'], p[layer + '/beta'],
                                  p[layer + '/moving_mean'], p[layer + '/moving_variance'], eps)

    def forward(self, x, training=True, params=None, updates=None, taps=None):
        """x: (N,H,W,C) -> raw network output (before l2norm).  `updates` (dict)
        receives new moving statistics; `taps` (dict) receives named intermediates."""
        return self._fwd(self.params if params is None else params, x, training, updates, taps)


def _tap(taps, name, t):
    if taps is not None:
        taps[name] = t
    return t


# ------------------------------------------------------------------------------------------ Plain-11
PLAIN11_FILTERS = [64, 64, 'ap', 128, 128, 128, 'ap', 256, 256, 256, 'ap', 512, 'gap', 'fc512']


def build_plainnet(output_dim, input_channels=3, seed=0, filters=PLAIN11_FILTERS, l2=0.0005):
    """models/plainnet.py:5-78: Conv(+bias, ReLU) -> BN; 'ap'; 'gap'; fc(+ReLU) -> BN; Dense(D)."""
    m = OracleModel('plain11')
    gen = torch.Generator().manual_seed(seed)
    plan = []
    cin = input_channels
    m.add_conv('conv1', 3, 3, cin, filters[0], gen, l2=l2)
    m.add_bn('bn1', filters[0], gen)
    plan.append(('conv', 'conv1', 'bn1'))
    cin = filters[0]
    for i, f in enumerate(filters[1:], start=2):
        if f == 'ap':
            plan.append(('ap', 'ap%d' % i, None))
        elif f == 'gap':
            plan.append(('gap', 'avg_pool', None))
        elif isinstance(f, str) and f.startswith('fc'):
            units = int(f[2:])
            m.add_dense('fc%d' % i, cin, units, gen, l2=l2)
            m.add_bn('bn%d' % i, units, gen)
            plan.append(('fc', 'fc%d' % i, 'bn%d' % i))
            cin = units
        else:
            m.add_conv('conv%d' % i, 3, 3, cin, f, gen, l2=l2)
            m.add_bn('bn%d' % i, f, gen)
            plan.append(('conv', 'conv%d' % i, 'bn%d' % i))
            cin = f
    m.add_dense('embedding', cin, output_dim, gen)      # plainnet.py:76: no regulariser
    m.out_layer = 'embedding'

    def fwd(p, x, training, updates, taps):
        for kind, layer, bnl in plan:
            if kind == 'conv':
                x = torch.relu(m.conv(p, layer, x))
                x = _tap(taps, bnl, m.bn(p, bnl, x, training, updates))
            elif kind == 'ap':
                x = nn.avgpool2(x)
            elif kind == 'gap':
                x = nn.gap(x)
            elif kind == 'fc':
                x = torch.relu(m.dense(p, layer, x))
                x = _tap(taps, bnl, m.bn(p, bnl, x, training, updates))
        return _tap(taps, 'embedding', m.dense(p, 'embedding', x))

    m._fwd = fwd
    return m


# ------------------------------------------------------------------------------------------ CIFAR ResNet
def build_small_resnet(n, filters, include_top, classes, input_channels=3, seed=0, l2=0.0002):
    """models/cifar_resnet.py:149-257 (`SmallResNet`), blocks :69-125, units :128-146."""
    m = OracleModel('cifar-resnet%d' % (2 * len(filters) * n))
    gen = torch.Generator().manual_seed(seed)
    m.add_conv('conv0', 3, 3, input_channels, filters[0], gen, l2=l2)
    m.add_bn('bn0', filters[0], gen)
    blocks = []
    for u in range(len(filters)):
        cin = filters[0] if u == 0 else filters[u - 1]
        cout = filters[u]
        for b in range(n):
            prefix = '%d-%d' % (u + 1, b + 1)
            stride = 2 if (u > 0 and b == 0) else 1
            bc_in = cin if b == 0 else cout
            m.add_conv('res' + prefix + 'x', 3, 3, bc_in, cout, gen, l2=l2)
            m.add_bn('bn' + prefix + 'x', cout, gen)
            m.add_conv('res' + prefix + 'y', 3, 3, cout, cout, gen, l2=l2)
            m.add_bn('bn' + prefix + 'y', cout, gen)
            blocks.append((prefix, stride, bc_in, cout))
    if include_top:
        m.add_dense('embedding', filters[-1], classes, gen, l2=l2)   # cifar_resnet.py:233
        m.out_layer = 'embedding'
    else:
        m.out_layer = 'avg_pool'

    def fwd(p, x, training, updates, taps):
        x = m.conv(p, 'conv0', x)
        x = torch.relu(m.bn(p, 'bn0', x, training, updates))
        for prefix, stride, bc_in, cout in blocks:
            inp = x
            y = m.conv(p, 'res' + prefix + 'x', inp, stride=stride)
            y = torch.relu(m.bn(p, 'bn' + prefix + 'x', y, training, updates))
            y = m.conv(p, 'res' + prefix + 'y', y)
            y = m.bn(p, 'bn' + prefix + 'y', y, training, updates)
            sc = inp
            if stride > 1:
                sc = nn.avgpool2(sc, stride)                     # :117-118
            if bc_in < cout:
                d = cout - bc_in
                sc = nn.channel_pad(sc, d // 2, d - d // 2)      # :119-121
            x = _tap(taps, 'block' + prefix, torch.relu(y + sc))
        x = _tap(taps, 'avg_pool', nn.gap(x))
        if include_top:
            x = _tap(taps, 'embedding', m.dense(p, 'embedding', x))
        return x

    m._fwd = fwd
    return m


# ------------------------------------------------------------------------------------------ WRN
def build_wrn(input_channels, nb_classes, N=4, k=10, seed=0):
    """models/wide_residual_network.py:60-101: no bias, no L2, he_normal, BN(momentum .1, eps 1e-5,
    gamma 'uniform'); Keras auto layer names conv2d_i / batch_normalization_i in creation order."""
    m = OracleModel('wrn-%d-%d' % (6 * N + 4, k))
    gen = torch.Generator().manual_seed(seed)
    cnt = {'c': 0, 'b': 0}

    def new_conv(kh, cin, cout):
        cnt['c'] += 1
        name = 'conv2d_%d' % cnt['c']
        m.add_conv(name, kh, kh, cin, cout, gen, use_bias=False, init='he_normal')
        return name

    def new_bn(c):
        cnt['b'] += 1
        name = 'batch_normalization_%d' % cnt['b']
        m.add_bn(name, c, gen, momentum=0.1, eps=1e-5, gamma_init='uniform')
        return name

    prog = []
    c0 = new_conv(3, input_channels, 16)
    b0 = new_bn(16)
    prog.append(('initial', c0, b0))
    cin = 16
    for gi, base in enumerate([16, 32, 64]):
        cout = base * k
        stride = 2 if gi > 0 else 1
        ca = new_conv(3, cin, cout)
        ba = new_bn(cout)
        cb = new_conv(3, cout, cout)
        cs = new_conv(1, cin, cout)
        prog.append(('expand', ca, ba, cb, cs, stride))
        for _ in range(N - 1):
            b1 = new_bn(cout)
            c1 = new_conv(3, cout, cout)
            b2 = new_bn(cout)
            c2 = new_conv(3, cout, cout)
            prog.append(('block', b1, c1, b2, c2))
        bl = new_bn(cout)
        prog.append(('bnrelu', bl))
        cin = cout
    m.add_dense('embedding', cin, nb_classes, gen)           # :96 Dense with bias, no L2
    m.out_layer = 'embedding'

    def fwd(p, x, training, updates, taps):
        for op in prog:
            if op[0] == 'initial':
                x = torch.relu(m.bn(p, op[2], m.conv(p, op[1], x), training, updates))
            elif op[0] == 'expand':
                _, ca, ba, cb, cs, stride = op
                y = m.conv(p, ca, x, stride=stride)
                y = torch.relu(m.bn(p, ba, y, training, updates))
                y = m.conv(p, cb, y)
                skip = m.conv(p, cs, x, stride=stride)
                x = y + skip
            elif op[0] == 'block':
                _, b1, c1, b2, c2 = op
                y = torch.relu(m.bn(p, b1, x, training, updates))
                y = m.conv(p, c1, y)
                y = torch.relu(m.bn(p, b2, y, training, updates))
                y = m.conv(p, c2, y)
                x = x + y
            elif op[0] == 'bnrelu':
                x = torch.relu(m.bn(p, op[1], x, training, updates))
        x = _tap(taps, 'avg_pool', nn.gap(x))
        return _tap(taps, 'embedding', m.dense(p, 'embedding', x))

    m._fwd = fwd
    return m


# ------------------------------------------------------------------------------------------ ResNet-50 v1
def build_resnet50(num_outputs, input_channels=3, seed=0):
    """keras.applications.ResNet50(include_top=False) + GAP + Dense('embedding') (utils.py:237-243).

    The backbone is third-party, un-vendored and unpinned ("parity unpinned"):
    this restates keras_applications/resnet50.py of the Keras 2.2 era -- ZeroPadding(3) + 7x7/2
    'valid' conv, BN (eps 1.001e-5... in later versions; Keras-2.2.0 ships default eps 1e-3), ReLU,
    3x3/2 max-pool ('valid' after the 112x112 map => 55x55), stages [3,4,6,3] with the stride on
    the first 1x1 conv; all convs with bias; BN after every conv.
    """
    m = OracleModel('resnet50')
    gen = torch.Generator().manual_seed(seed)
    m.add_conv('conv1', 7, 7, input_channels, 64, gen)
    m.add_bn('bn_conv1', 64, gen)
    stages = [(2, [64, 64, 256], 3, 1), (3, [128, 128, 512], 4, 2),
              (4, [256, 256, 1024], 6, 2), (5, [512, 512, 2048], 3, 2)]
    prog = []
    cin = 64
    for stage, (f1, f2, f3), nblocks, stride in stages:
        for bi in range(nblocks):
            blk = chr(ord('a') + bi)
            base = 'res%d%s_branch' % (stage, blk)
            bnb = 'bn%d%s_branch' % (stage, blk)
            s = stride if bi == 0 else 1
            m.add_conv(base + '2a', 1, 1, cin, f1, gen)
            m.add_bn(bnb + '2a', f1, gen)
            m.add_conv(base + '2b', 3, 3, f1, f2, gen)
            m.add_bn(bnb + '2b', f2, gen)
            m.add_conv(base + '2c', 1, 1, f2, f3, gen)
            m.add_bn(bnb + '2c', f3, gen)
            if bi == 0:
                m.add_conv(base + '1', 1, 1, cin, f3, gen)
                m.add_bn(bnb + '1', f3, gen)
            prog.append((base, bnb, s, bi == 0))
            cin = f3
    m.add_dense('embedding', 2048, num_outputs, gen)
    m.out_layer = 'embedding'

    def fwd(p, x, training, updates, taps):
        x = m.conv(p, 'conv1', x, stride=2, padding=(3, 3, 3, 3))
        x = torch.relu(m.bn(p, 'bn_conv1', x, training, updates))
        x = _tap(taps, 'pool1', nn.maxpool(x, 3, 2))
        for base, bnb, s, proj in prog:
            y = m.conv(p, base + '2a', x, stride=s, padding='valid')
            y = torch.relu(m.bn(p, bnb + '2a', y, training, updates))
            y = m.conv(p, base + '2b', y)
            y = torch.relu(m.bn(p, bnb + '2b', y, training, updates))
            y = m.conv(p, base + '2c', y, padding='valid')
            y = m.bn(p, bnb + '2c', y, training, updates)
            if proj:
                sc = m.conv(p, base + '1', x, stride=s, padding='valid')
                sc = m.bn(p, bnb + '1', sc, training, updates)
            else:
                sc = x
            x = torch.relu(y + sc)
        x = _tap(taps, 'avg_pool', nn.gap(x))
        return _tap(taps, 'embedding', m.dense(p, 'embedding', x))

    m._fwd = fwd
    return m


# ------------------------------------------------------------------------------------------ dispatch
ARCHITECTURES = ['simple', 'resnet-32', 'resnet-110', 'resnet-110-fc', 'resnet-110-wfc', 'wrn-28-10', 'resnet-50']


def build_network(num_outputs, architecture, input_channels=None, seed=0):
    """utils.py:130-276 with classification=False (the embedding-learning call site,
    learn_image_embeddings.py:125)."""
    ic = 3 if input_channels is None else input_channels
    if architecture == 'simple':
        return build_plainnet(num_outputs, ic, seed)
    if architecture == 'resnet-32':
        return build_small_resnet(5, [16, 32, 64], False, num_outputs, ic, seed)
    if architecture == 'resnet-110':
        return build_small_resnet(18, [16, 32, 64], False, num_outputs, ic, seed)     # utils.py:168-172
    if architecture == 'resnet-110-fc':
        return build_small_resnet(18, [16, 32, 64], True, num_outputs, ic, seed)      # utils.py:174-178
    if architecture == 'resnet-110-wfc':
        return build_small_resnet(18, [32, 64, 128], True, num_outputs, ic, seed)     # utils.py:180-184
    if architecture == 'wrn-28-10':
        return build_wrn(ic, num_outputs, N=4, k=10, seed=seed)                       # utils.py:186-191
    if architecture == 'resnet-50':
        return build_resnet50(num_outputs, ic, seed)
    raise ValueError('Unknown network architecture: {}'.format(architecture))


def randomize(model_or_params, seed=123):
    """Fills biases / BN parameters / moving statistics (all-zero or all-one after a Keras init)
    with seeded non-trivial values so that parity tests exercise every term."""
    params = model_or_params.params if hasattr(model_or_params, 'params') else model_or_params
    gen = torch.Generator().manual_seed(seed)
    for name, t in params.items():
        if name.endswith('/bias') or name.endswith('/beta') or name.endswith('/moving_mean'):
            t.copy_(torch.randn(t.shape, generator=gen, dtype=torch.float64) * 0.1)
        elif name.endswith('/gamma') or name.endswith('/moving_variance'):
            t.copy_(torch.rand(t.shape, generator=gen, dtype=torch.float64) + 0.5)
    return model_or_params
