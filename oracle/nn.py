"""Layer semantics of Keras 2.2 / TF 1.x (channels_last) restated with torch-CPU ops.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  All tensors are NHWC at the
interface, like the reference graph (`K.image_data_format() == 'channels_last'`,
models/cifar_resnet.py:87-90).  Autograd provides the backward pass, exactly as
TF autodiff does for the reference (learn_image_embeddings.py:238).
"""
import math

import torch
import torch.nn.functional as F


def same_pad(in_size, k, stride):
    """TF 'SAME' padding: returns (before, after, out)."""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    before = total // 2
    return before, total - before, out


def conv2d(x, kernel, bias=None, stride=1, padding='same'):
    """Keras Conv2D (cross-correlation, HWIO kernel), models/cifar_resnet.py:96-105.

    x: (N,H,W,Cin); kernel: (kh,kw,Cin,Cout); padding 'same' | 'valid' | (pt,pb,pl,pr).
    """
    kh, kw = kernel.shape[0], kernel.shape[1]
    n, h, w, c = x.shape
    if padding == 'same':
        pt, pb, _ = same_pad(h, kh, stride)
        pl, pr, _ = same_pad(w, kw, stride)
    elif padding == 'valid':
        pt = pb = pl = pr = 0
    else:
        pt, pb, pl, pr = padding
    xn = x.permute(0, 3, 1, 2)
    xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, kernel.permute(3, 2, 0, 1), bias, stride=stride)
    return y.permute(0, 2, 3, 1)


def batchnorm_train(x, gamma, beta, eps):
    """Keras BatchNormalization(axis=-1) in training mode: per-channel mean and
    BIASED variance over every axis but the last.
    Returns (y, batch_mean, batch_var_biased)."""
    red = tuple(range(x.dim() - 1))
    mean = x.mean(dim=red)
    var = ((x - mean) ** 2).mean(dim=red)
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    return y, mean, var


def batchnorm_infer(x, gamma, beta, moving_mean, moving_var, eps):
    return (x - moving_mean) * torch.rsqrt(moving_var + eps) * gamma + beta


def moving_update(moving, batch, momentum):
    """m <- m*momentum + batch*(1-momentum)  (Keras BatchNormalization)."""
    return moving * momentum + batch * (1.0 - momentum)


def unbiased_var(var_biased, count, eps):
    """Keras 2.2 feeds `var * n / (n - (1 + eps))` to the moving average
    ((K); only matters for inference-mode parity)."""
    return var_biased * (count / (count - (1.0 + eps)))


def avgpool2(x, pool=2):
    """AveragePooling2D(pool) default strides=pool, 'valid' (models/plainnet.py:59)."""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), pool).permute(0, 2, 3, 1)


def maxpool(x, k=3, stride=2, pad=(0, 0, 0, 0)):
    """MaxPooling2D; pad=(pt,pb,pl,pr) applied with -inf first ('same' style)."""
    xn = x.permute(0, 3, 1, 2)
    if any(pad):
        xn = F.pad(xn, (pad[2], pad[3], pad[0], pad[1]), value=float('-inf'))
    return F.max_pool2d(xn, k, stride).permute(0, 2, 3, 1)


def gap(x):
    """GlobalAveragePooling2D (models/cifar_resnet.py:228)."""
    return x.mean(dim=(1, 2))


def channel_pad(x, lo, hi):
    """models/cifar_resnet.py:57-61 ChannelPadding: zeros on the channel axis."""
    return F.pad(x, (lo, hi))


def dense(x, kernel, bias=None):
    y = x @ kernel
    return y if bias is None else y + bias


def l2norm(x):
    """utils.py:125-127 -> tf.nn.l2_normalize(x, -1) = x * rsqrt(max(sum x^2, 1e-12))."""
    ss = (x * x).sum(dim=-1, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=1e-12))


def inv_correlation(y_true, y_pred):
    """utils.py:44-46."""
    return 1.0 - (y_true * y_pred).sum(dim=-1)


def squared_distance(y_true, y_pred):
    """utils.py:34-36."""
    return ((y_pred - y_true) ** 2).sum(dim=-1)


def max_sim_acc(embedding, y_true, y_pred):
    """utils.py:87-93 (k<=1 branch)."""
    sim = y_pred @ embedding.t()
    true_sim = (y_pred * y_true).sum(dim=-1)
    return ((sim.max(dim=-1).values - true_sim).abs() < 1e-6).to(y_pred.dtype)


def nn_accuracy(embedding, y_true, y_pred):
    """utils.py:73-83 (Euclidean variant, k<=1 branch)."""
    cn = (embedding.t() ** 2).sum(dim=0, keepdim=True)
    pn = (y_pred ** 2).sum(dim=1, keepdim=True)
    dist = pn + cn - 2 * (y_pred @ embedding.t())
    true_dist = ((y_pred - y_true) ** 2).sum(dim=-1)
    return ((true_dist - dist.min(dim=-1).values).abs() < 1e-6).to(y_pred.dtype)


def max_sim_acc_k(embedding, y_true, y_pred, k):
    """utils.py:95 (k > 1 branch): any of the k largest similarities within 1e-6 of the true one."""
    sim = y_pred @ embedding.t()
    true_sim = (y_pred * y_true).sum(dim=-1)
    top = torch.topk(sim, k, dim=-1).values
    return ((top - true_sim[:, None]).abs() < 1e-6).any(dim=-1).to(y_pred.dtype)


def nn_accuracy_k(embedding, y_true, y_pred, k):
    """utils.py:85 (k > 1 branch, Euclidean variant)."""
    cn = (embedding.t() ** 2).sum(dim=0, keepdim=True)
    pn = (y_pred ** 2).sum(dim=1, keepdim=True)
    dist = pn + cn - 2 * (y_pred @ embedding.t())
    true_dist = ((y_pred - y_true) ** 2).sum(dim=-1)
    top = -torch.topk(-dist, k, dim=-1).values
    return ((top - true_dist[:, None]).abs() < 1e-6).any(dim=-1).to(y_pred.dtype)


def categorical_accuracy(y_true, y_pred):
    """Keras metric 'accuracy' for 2-d targets (learn_image_embeddings.py:166 with --loss softmax_corr)."""
    return (y_true.argmax(dim=-1) == y_pred.argmax(dim=-1)).to(y_pred.dtype)


def top_k_categorical_accuracy(y_true, y_pred, k):
    """utils.top_k_acc (utils.py:49-54) = K.in_top_k(y_pred, argmax(y_true), k): the target is in the top k when fewer
    than k entries are strictly larger than its own."""
    tgt = y_true.argmax(dim=-1)
    own = y_pred.gather(1, tgt[:, None])
    return ((y_pred > own).sum(dim=-1) < k).to(y_pred.dtype)


def categorical_crossentropy(onehot, prob):
    """Keras categorical_crossentropy on probabilities:
    renormalise, clip to [1e-7, 1-1e-7], -sum t log p."""
    p = prob / prob.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, 1e-7, 1.0 - 1e-7)
    return -(onehot * torch.log(p)).sum(dim=-1)


# ----------------------------------------------------------------------------- initialisers
def glorot_uniform(shape, gen, dtype=torch.float64):
    """Keras default kernel init: U(-l, l), l = sqrt(6/(fan_in+fan_out))."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = shape[0] * shape[1]
        fan_in, fan_out = rf * shape[2], rf * shape[3]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).mul_(lim).to(dtype)


def he_normal(shape, gen, dtype=torch.float64):
    """Keras he_normal: truncated normal, stddev sqrt(2/fan_in) (truncation at 2 sigma)."""
    rf = shape[0] * shape[1] if len(shape) == 4 else 1
    fan_in = rf * shape[-2]
    std = math.sqrt(2.0 / fan_in)
    t = torch.randn(shape, generator=gen, dtype=torch.float64).clamp_(-2, 2)
    return t.mul_(std).to(dtype)
