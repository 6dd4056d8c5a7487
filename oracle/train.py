"""Training-step semantics of learn_image_embeddings.py restated on the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  head variants          learn_image_embeddings.py:127-132,164-175
  cls_model              learn_image_embeddings.py:16-45
  transform_inputs       learn_image_embeddings.py:48-50
  total loss             Keras: mean over batch of each output loss, weighted sum, + regularisers (A.5)
  SGD(clipnorm, decay)   learn_image_embeddings.py:224-236; keras.optimizers.SGD (A.6)
  SGDR                   sgdr_callback.py:6-87; utils.py:357-368 (A.7)
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import nn


class ClsHead:
    """cls_model (learn_image_embeddings.py:16-45): relu -> BatchNormalization() -> Dense(C, softmax, l2 5e-4)."""

    def __init__(self, dim, num_classes, seed=1):
        gen = torch.Generator().manual_seed(seed)
        self.params = OrderedDict()
        self.params['cls_bn/gamma'] = torch.ones(dim, dtype=torch.float64)
        self.params['cls_bn/beta'] = torch.zeros(dim, dtype=torch.float64)
        self.params['cls_bn/moving_mean'] = torch.zeros(dim, dtype=torch.float64)
        self.params['cls_bn/moving_variance'] = torch.ones(dim, dtype=torch.float64)
        self.params['prob/kernel'] = nn.glorot_uniform((dim, num_classes), gen)
        self.params['prob/bias'] = torch.zeros(num_classes, dtype=torch.float64)
        self.trainable = ['cls_bn/gamma', 'cls_bn/beta', 'prob/kernel', 'prob/bias']
        self.l2 = {'prob/kernel': 5e-4}
        self.momentum, self.eps = 0.99, 1e-3

    def forward(self, p, base, training=True, updates=None):
        x = torch.relu(base)
        if training:
            x, mean, var = nn.batchnorm_train(x, p['cls_bn/gamma'], p['cls_bn/beta'], self.eps)
            if updates is not None:
                n = base.shape[0]
                updates['cls_bn/moving_mean'] = nn.moving_update(p['cls_bn/moving_mean'], mean.detach(), self.momentum)
                updates['cls_bn/moving_variance'] = nn.moving_update(
                    p['cls_bn/moving_variance'], nn.unbiased_var(var.detach(), n, self.eps), self.momentum)
        else:
            x = nn.batchnorm_infer(x, p['cls_bn/gamma'], p['cls_bn/beta'],
                                   p['cls_bn/moving_mean'], p['cls_bn/moving_variance'], self.eps)
        logits = nn.dense(x, p['prob/kernel'], p['prob/bias'])
        return torch.softmax(logits, dim=-1)


def head_forward(z, loss_kind):
    """Output wrapper chosen by --loss (learn_image_embeddings.py:127-130)."""
    if loss_kind == 'inv_corr':
        return nn.l2norm(z)
    if loss_kind == 'softmax_corr':
        return torch.softmax(z, dim=-1)
    return z                                            # 'unnorm_corr', 'mse'


def per_sample_loss(y_true, y_pred, loss_kind):
    if loss_kind.endswith('_corr'):
        return nn.inv_correlation(y_true, y_pred)       # :164-165
    return nn.squared_distance(y_true, y_pred)          # :171


def train_objective(model, x, labels, embedding, loss_kind='inv_corr', cls=None, cls_weight=0.0,
                    params=None, updates=None, cls_base=None):
    """Total Keras training loss for one batch and the metric tensors.

    params: dict of ALL weights (model + cls head); defaults to the objects' own.
    Returns dict(total, embed_loss, cls_loss, reg, acc, emb (wrapped output), prob)."""
    p = dict(model.params) if params is None else params
    if cls is not None and params is None:
        p.update(cls.params)
    taps = {} if cls_base is not None else None
    z = model.forward(x, training=True, params=p, updates=updates, taps=taps)
    emb = head_forward(z, loss_kind)
    y_true = embedding[labels]                          # transform_inputs, :48-50
    ls = per_sample_loss(y_true, emb, loss_kind)
    embed_loss = ls.mean()
    if loss_kind == 'mse':
        acc = nn.nn_accuracy(embedding, y_true, emb)
    elif loss_kind == 'softmax_corr':
        acc = nn.categorical_accuracy(y_true, emb)          # metrics = ['accuracy'], learn_image_embeddings.py:166
    else:
        acc = nn.max_sim_acc(embedding, y_true, emb)
    total = embed_loss
    cls_loss = None
    prob = None
    if cls is not None and cls_weight > 0:
        # cls_model(embed_model, num_classes, cls_base) (learn_image_embeddings.py:34-40): the classifier reads the wrapped
        # embedding output, or the output of the named inner layer
        prob = cls.forward(p, emb if cls_base is None else taps[cls_base], True, updates)
        onehot = torch.nn.functional.one_hot(labels, prob.shape[-1]).to(prob.dtype)
        cls_loss = nn.categorical_crossentropy(onehot, prob).mean()
        total = total + cls_weight * cls_loss
    reg = torch.zeros((), dtype=z.dtype)
    l2 = dict(model.l2)
    if cls is not None and cls_weight > 0:
        l2.update(cls.l2)
    for name, lam in l2.items():
        reg = reg + lam * (p[name] ** 2).sum()           # regularizers.l2: lambda * sum(W^2)
    total = total + reg
    return dict(total=total, embed_loss=embed_loss, cls_loss=cls_loss, reg=reg, acc=acc,
                emb=emb, prob=prob, per_sample=ls, z=z)


def sgd_step(params, grads, velocity, lr, momentum=0.9, nesterov=False, clipnorm=10.0,
             decay=0.0, iterations=0):
    """keras.optimizers.SGD.get_updates (Keras 2.2): global-norm clip, lr decay, momentum.

    params/grads/velocity: dict name -> tensor (updated in place).  Returns the global grad norm."""
    names = list(grads.keys())
    norm = math.sqrt(sum(float((grads[n] ** 2).sum()) for n in names))
    scale = 1.0
    if clipnorm and clipnorm > 0 and norm >= clipnorm:
        scale = clipnorm / norm
    lr_t = lr * (1.0 / (1.0 + decay * iterations)) if decay > 0 else lr
    for n in names:
        g = grads[n] * scale
        v = momentum * velocity[n] - lr_t * g
        velocity[n].copy_(v)
        if nesterov:
            params[n].add_(momentum * v - lr_t * g)
        else:
            params[n].add_(v)
    return norm


def train_step(model, x, labels, embedding, velocity, lr, loss_kind='inv_corr', cls=None,
               cls_weight=0.0, nesterov=False, clipnorm=10.0, decay=0.0, iterations=0, cls_base=None):
    """One full reference training step (fwd, autodiff bwd, clip, momentum SGD, BN moving stats).
    Returns (objective dict, grads dict, grad norm)."""
    allp = OrderedDict(model.params)
    trainable = list(model.trainable)
    if cls is not None and cls_weight > 0:
        allp.update(cls.params)
        trainable += cls.trainable
    leaf = {n: allp[n].detach().clone().requires_grad_(n in trainable) for n in allp}
    updates = {}
    obj = train_objective(model, x, labels, embedding, loss_kind, cls, cls_weight, params=leaf, updates=updates,
                          cls_base=cls_base)
    tl = [leaf[n] for n in trainable]
    gl = torch.autograd.grad(obj['total'], tl, allow_unused=True)
    grads = OrderedDict((n, (g if g is not None else torch.zeros_like(leaf[n])).detach())
                        for n, g in zip(trainable, gl))
    pview = {n: allp[n] for n in trainable}
    norm = sgd_step(pview, grads, velocity, lr, 0.9, nesterov, clipnorm, decay, iterations)
    for n, v in updates.items():
        allp[n].copy_(v)
    return obj, grads, norm


def make_velocity(model, cls=None):
    v = OrderedDict((n, torch.zeros_like(model.params[n])) for n in model.trainable)
    if cls is not None:
        for n in cls.trainable:
            v[n] = torch.zeros_like(cls.params[n])
    return v


# ------------------------------------------------------------------------------------------ SGDR
class SGDR:
    """sgdr_callback.py:6-87, restated without Keras: `lr` is the optimizer learning rate."""

    def __init__(self, min_lr=0.0, max_lr=0.05, base_epochs=10, mul_epochs=2):
        self.min_lr, self.max_lr = min_lr, max_lr
        self.base_epochs, self.mul_epochs = base_epochs, mul_epochs
        self.cycles = 0.
        self.cycle_iterations = 0.
        self.trn_iterations = 0.
        self.lr = None

    def sgdr(self):                                      # :63-66
        cycle_epochs = self.base_epochs * (self.mul_epochs ** self.cycles)
        return self.min_lr + 0.5 * (self.max_lr - self.min_lr) * \
            (1 + np.cos(np.pi * (self.cycle_iterations + 1) / cycle_epochs))

    def on_train_begin(self):                            # :68-73
        self.lr = self.max_lr if self.cycle_iterations == 0 else self.sgdr()

    def on_epoch_end(self):                              # :75-87
        self.trn_iterations += 1
        self.cycle_iterations += 1
        if self.cycle_iterations >= self.base_epochs * (self.mul_epochs ** self.cycles):
            self.cycles += 1
            self.cycle_iterations = 0
            self.lr = self.max_lr
        else:
            self.lr = self.sgdr()


def sgdr_lr_sequence(num_epochs, min_lr=1e-6, max_lr=0.1, base=12, mul=2):
    """LR used during each epoch under utils.get_lr_schedule('SGDR') defaults (utils.py:357-368)."""
    s = SGDR(min_lr, max_lr, base, mul)
    s.on_train_begin()
    out = []
    for _ in range(num_epochs):
        out.append(float(s.lr))
        s.on_epoch_end()
    return out


def sgdr_default_epochs(base=12, mul=2):
    return sum(base * (mul ** i) for i in range(5))      # utils.py:367


def cast_model(model, dtype, cls=None):
    """In-place dtype change of every weight (float32 for the timed CPU baseline of bench.py)."""
    for k in list(model.params.keys()):
        model.params[k] = model.params[k].to(dtype)
    if cls is not None:
        for k in list(cls.params.keys()):
            cls.params[k] = cls.params[k].to(dtype)
    return model
