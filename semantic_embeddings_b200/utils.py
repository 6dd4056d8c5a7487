"""Host-side mirror of the reference's utils.py for the hot path: same function names, argument
meaning and error behaviour (reference file:line cited per function), returning layer-list graphs
(graph.Graph) instead of Keras models and plain schedule objects instead of Keras callbacks.

The tensor functions (l2norm, inv_correlation, squared_distance, nn_accuracy) are the stand-alone
forms of what the fused head kernel computes inside a training step; they run the same CUDA kernel
(se_embed_head_fwd_bwd) on device tensors -- there is no CPU implementation in the product."""
import numpy as np

from . import _lib
from .models import cifar_resnet, plainnet, resnet50, wide_residual_network as wrn
from .sgdr_callback import SGDR

# utils.py:26-28 lists more names; the ones below are the architectures of the hot path (BASELINE.json configs)
ARCHITECTURES = ['simple', 'resnet-32', 'resnet-110', 'resnet-110-fc', 'resnet-110-wfc', 'wrn-28-10', 'resnet-50']
REFERENCE_ARCHITECTURES = ['simple', 'resnet-32', 'resnet-110', 'resnet-110-fc', 'resnet-110-wfc', 'wrn-28-10',
                           'densenet-100-12', 'densenet-100-24', 'densenet-bc-190-40', 'pyramidnet-272-200',
                           'pyramidnet-110-270', 'resnet-50', 'resnet-101', 'resnet-152', 'rn18', 'rn34', 'rn50', 'rn101',
                           'rn152', 'rn200', 'nasnet-a']
LR_SCHEDULES = ['SGD', 'SGDR', 'CLR', 'ResNet-Schedule']


def build_network(num_outputs, architecture, classification=False, no_softmax=False, input_channels=None, name=None,
                  input_size=None):
    """utils.py:130-276.  Returns a graph.Graph whose output is the raw embedding (no l2norm wrapper; the
    wrapper and the loss are the fused head of the engine, learn_image_embeddings.py:127-128).

    `classification=True` (softmax classifier training) belongs to learn_classifier.py, which is outside the
    hot path: it raises NotImplementedError.  Unknown names raise ValueError like utils.py:276."""
    if architecture.lower().endswith('-selu'):
        raise NotImplementedError('SELU variants (utils.py:152-156) are outside the accelerated hot path')
    if classification and not no_softmax:
        raise NotImplementedError('softmax classification heads belong to learn_classifier.py (not on the hot path)')
    ic = 3 if input_channels is None else input_channels
    hw = input_size
    if architecture == 'resnet-32':
        return cifar_resnet.SmallResNet(5, [16, 32, 64], include_top=classification, input_shape=(hw or 32, hw or 32, ic),
                                        classes=num_outputs, name=name)
    elif architecture == 'resnet-110':
        # utils.py:168-172: include_top=classification -> the 64-d pooled output, num_outputs ignored
        return cifar_resnet.SmallResNet(18, [16, 32, 64], include_top=classification, input_shape=(hw or 32, hw or 32, ic),
                                        classes=num_outputs, name=name)
    elif architecture == 'resnet-110-fc':
        return cifar_resnet.SmallResNet(18, [16, 32, 64], include_top=True, input_shape=(hw or 32, hw or 32, ic),
                                        classes=num_outputs, name=name)
    elif architecture == 'resnet-110-wfc':
        return cifar_resnet.SmallResNet(18, [32, 64, 128], include_top=True, input_shape=(hw or 32, hw or 32, ic),
                                        classes=num_outputs, name=name)
    elif architecture == 'wrn-28-10':
        return wrn.create_wide_residual_network((hw or 32, hw or 32, ic), nb_classes=num_outputs, N=4, k=10, name=name)
    elif architecture == 'simple':
        return plainnet.PlainNet(num_outputs, input_shape=(hw or 32, hw or 32, ic), name=name)
    elif architecture == 'resnet-50':
        return resnet50.ResNet50(num_outputs, input_shape=(hw or 224, hw or 224, ic), name=name)
    elif architecture in REFERENCE_ARCHITECTURES:
        raise NotImplementedError('architecture {} exists in the reference but is outside the accelerated hot path '
                                  '(SURVEY.md section 2)'.format(architecture))
    else:
        raise ValueError('Unknown network architecture: {}'.format(architecture))


# ------------------------------------------------------------------------------------------------ head functions
def _head(z, labels, embedding, loss_kind, want):
    import torch
    z = z.contiguous().float()
    B, D = z.shape
    E = torch.as_tensor(np.ascontiguousarray(np.asarray(embedding, dtype=np.float32))).to(z.device) \
        if not torch.is_tensor(embedding) else embedding.contiguous().float()
    C = E.shape[0]
    lab = torch.zeros(B, dtype=torch.int32, device=z.device) if labels is None else labels.to(torch.int32).contiguous()
    x = torch.empty_like(z)
    loss = torch.empty(B, device=z.device)
    acc = torch.empty(B, device=z.device)
    _lib.call('se_embed_head_fwd_bwd', z.data_ptr(), D, lab.data_ptr(), E.data_ptr(), D, B, D, C, loss_kind, 1.0, None,
              x.data_ptr(), loss.data_ptr(), acc.data_ptr(), None, _lib.stream_ptr())
    return {'x': x, 'loss': loss, 'acc': acc}[want]


def l2norm(x):
    """utils.py:125-127: L2-normalises a (B,D) CUDA tensor along the last axis (x * rsqrt(max(sum x^2, 1e-12)))."""
    import torch
    eye = torch.zeros(1, x.shape[1], device=x.device)
    return _head(x, None, eye, _lib.SE_LOSS_INV_CORR, 'x')


def inv_correlation(embedding, labels, y_pred):
    """utils.py:44-46 with the target gather of learn_image_embeddings.py:48-50 folded in:
    1 - <embedding[labels], y_pred> per sample."""
    return _head(y_pred, labels, embedding, _lib.SE_LOSS_UNNORM_CORR, 'loss')


def squared_distance(embedding, labels, y_pred):
    """utils.py:34-36."""
    return _head(y_pred, labels, embedding, _lib.SE_LOSS_MSE, 'loss')


def nn_accuracy(embedding, dot_prod_sim=False, k=1):
    """utils.py:57-100: returns metric(labels, y_pred) -> per-sample 0/1 tensor.  k > 1 (top-k) is not part of the
    fused kernel."""
    if k > 1:
        raise NotImplementedError('top-k nn_accuracy (utils.py:85,95) is outside the fused head kernel')

    def nn_accuracy(labels, y_pred):
        return _head(y_pred, labels, embedding, _lib.SE_LOSS_MSE, 'acc')

    def max_sim_acc(labels, y_pred):
        return _head(y_pred, labels, embedding, _lib.SE_LOSS_UNNORM_CORR, 'acc')

    return max_sim_acc if dot_prod_sim else nn_accuracy


# ------------------------------------------------------------------------------------------------ LR schedules
def get_lr_schedule(schedule, num_samples, batch_size, schedule_args={}):
    """utils.py:288-399.  Returns ([schedule objects], suggested number of epochs).  Only the SGDR branch
    (utils.py:357-368, the default of learn_image_embeddings.py:67) is on the hot path."""
    if schedule.lower() == 'sgdr':
        if 'sgdr_base_len' not in schedule_args:
            schedule_args['sgdr_base_len'] = 12
        if 'sgdr_mul' not in schedule_args:
            schedule_args['sgdr_mul'] = 2
        if 'sgdr_max_lr' not in schedule_args:
            schedule_args['sgdr_max_lr'] = 0.1
        return (
            [SGDR(1e-6, schedule_args['sgdr_max_lr'], schedule_args['sgdr_base_len'], schedule_args['sgdr_mul'])],
            sum(schedule_args['sgdr_base_len'] * (schedule_args['sgdr_mul'] ** i) for i in range(5))
        )
    elif schedule.lower() in ('sgd', 'clr', 'resnet-schedule'):
        raise NotImplementedError('LR schedule {} exists in the reference (utils.py:326-391) but is outside the '
                                  'accelerated hot path'.format(schedule))
    else:
        raise ValueError('Unknown learning rate schedule: {}'.format(schedule))


def add_lr_schedule_arguments(parser):
    """utils.py:402-418 -- same flags, types and defaults."""
    arggroup = parser.add_argument_group('Parameters for --lr_schedule=SGD')
    arggroup.add_argument('--sgd_patience', type=int, default=None, help='Patience of learning rate reduction in epochs.')
    arggroup.add_argument('--sgd_lr', type=float, default=0.1, help='Initial learning rate.')
    arggroup.add_argument('--sgd_min_lr', type=float, default=None, help='Minimum learning rate.')
    arggroup.add_argument('--sgd_schedule', type=str, default=None,
                          help='Comma-separated list of `epoch:lr` pairs, defining a learning rate schedule.')
    arggroup = parser.add_argument_group('Parameters for --lr_schedule=SGDR')
    arggroup.add_argument('--sgdr_base_len', type=int, default=None, help='Length of first cycle in epochs.')
    arggroup.add_argument('--sgdr_mul', type=int, default=None, help='Multiplier for cycle length after each cycle.')
    arggroup.add_argument('--sgdr_max_lr', type=float, default=None, help='Maximum learning rate.')
    arggroup = parser.add_argument_group('Parameters for --lr_schedule=CLR')
    arggroup.add_argument('--clr_step_len', type=int, default=None, help='Length of each step in epochs.')
    arggroup.add_argument('--clr_min_lr', type=float, default=None, help='Minimum learning rate.')
    arggroup.add_argument('--clr_max_lr', type=float, default=None, help='Maximum learning rate.')
