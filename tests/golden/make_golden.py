"""Generates the committed fixtures under tests/golden/ by executing the REFERENCE's own Python.

Runs ONLY in the build container (needs /root/reference); the fixtures it writes are what
travels.  Usage:  python tests/golden/make_golden.py

What is produced and what each fixture pins:
  class_matrices.npz       the fixed class-embedding matrices shipped by the reference
                           (embeddings/cifar100.unitsphere.pickle, embeddings/nab.unitsphere.pickle),
                           re-derived with compute_class_embedding.unitsphere_embedding from the taxonomy
                           files to check they are what the reference's code generates.
  cifar_hierarchy.npz      parent/child id pairs of Cifar-Hierarchy/cifar.parent-child.txt
  retrieval_ref.npz        evaluate_retrieval.pairwise_retrieval (reference code, numexpr stub) on seeded
                           features: full rankings for ndarray / dict / {'feat':..} inputs, normalize on/off,
                           + ClassHierarchy.hierarchical_precision metrics on them
  sgdr_ref.json            sgdr_callback.SGDR driven for 372 epochs (reference code, Keras stub)
  formulas_ref.npz         utils.l2norm / inv_correlation / squared_distance / nn_accuracy (reference code,
                           backend stub) on seeded inputs
  arch_<name>.json/.npz    layer trace and forward output of the reference's model builders
                           (utils.build_network + the Lambda(l2norm) wrap + cls_model) executed eagerly
                           under tests/golden/keras_stub.py with oracle-seeded weights
"""
import io
import json
import os
import pickle
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import keras_stub  # noqa: E402

keras_stub.install()
sys.path.insert(0, REF)
warnings.filterwarnings('ignore')

from oracle import models as omodels  # noqa: E402
from oracle import train as otrain  # noqa: E402


def class_matrices():
    import compute_class_embedding as cce
    from class_hierarchy import ClassHierarchy
    out = {}
    for key, pk, hier, is_a in (('cifar100', 'cifar100.unitsphere.pickle', 'Cifar-Hierarchy/cifar.parent-child.txt', False),
                                ('nab', 'nab.unitsphere.pickle', 'NAB-Hierarchy/hierarchy.txt', True)):
        with open(os.path.join(REF, 'embeddings', pk), 'rb') as f:
            d = pickle.load(f)
        emb = np.asarray(d['embedding'], dtype=np.float64)
        ind2label = np.asarray(d['ind2label'])
        # re-derive with the reference's code (compute_class_embedding.py:201-223)
        h = ClassHierarchy.from_file(os.path.join(REF, hier), is_a_relations=is_a, id_type=int)
        labels = list(ind2label.tolist())
        sem = np.zeros((len(labels), len(labels)))
        for i in range(len(labels)):
            for j in range(i + 1, len(labels)):
                sem[i, j] = sem[j, i] = h.lcs_height(labels[i], labels[j])
        re = cce.unitsphere_embedding(1. - sem)
        dev = float(np.abs(re - emb).max())
        print('%s: shipped pickle vs re-derived embedding: max abs dev %.3g' % (key, dev))
        assert dev < 1e-12
        out[key + '_embedding'] = emb
        out[key + '_ind2label'] = ind2label
    np.savez_compressed(os.path.join(HERE, 'class_matrices.npz'), **out)


def cifar_hierarchy():
    pairs = np.loadtxt(os.path.join(REF, 'Cifar-Hierarchy/cifar.parent-child.txt'), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, 'cifar_hierarchy.npz'), parent_child=pairs)


def retrieval():
    import evaluate_retrieval as er
    from class_hierarchy import ClassHierarchy
    rng = np.random.RandomState(0)
    n, d = 256, 100
    feat = rng.randn(n, d).astype(np.float32)
    feat_unit = feat / np.linalg.norm(feat, axis=-1, keepdims=True)
    labels = rng.randint(0, 100, n)
    ids = rng.permutation(10000)[:n]
    out = {'feat': feat, 'feat_unit': feat_unit, 'labels': labels, 'ids': ids}
    # 2-d array input, both modes (normalize=True mutates its argument: evaluate_retrieval.py:58)
    out['rank_sq'] = np.array([r for _, r in er.pairwise_retrieval(feat.copy(), normalize=False)], dtype=np.int32)
    out['rank_cos'] = np.array([r for _, r in er.pairwise_retrieval(feat.copy(), normalize=True)], dtype=np.int32)
    out['rank_sq_unit'] = np.array([r for _, r in er.pairwise_retrieval(feat_unit.copy(), normalize=False)], dtype=np.int32)
    # dict / {'feat': dict} inputs
    fd = {int(i): f.copy() for i, f in zip(ids, feat)}
    r_dict = er.pairwise_retrieval({'feat': fd}, normalize=False, return_generator=False)
    out['rank_dict_keys'] = np.array(list(r_dict.keys()), dtype=np.int64)
    out['rank_dict_vals'] = np.array(list(r_dict.values()), dtype=np.int64)
    # hierarchical precision with the reference's ClassHierarchy on the reference's ranking
    h = ClassHierarchy.from_file(os.path.join(REF, 'Cifar-Hierarchy/cifar.parent-child.txt'), id_type=int)
    lab = dict(enumerate(labels.tolist()))
    ret = dict(enumerate(out['rank_sq_unit'].tolist()))
    avg, per = h.hierarchical_precision(ret, lab, ks=[1, 10, 50, 100], compute_ahp=True, compute_ap=True,
                                        all_ids=list(range(n)))
    avg250, per250 = h.hierarchical_precision(ret, lab, ks=[1, 10, 50, 100], compute_ahp=250, compute_ap=False,
                                              all_ids=list(range(n)))
    out['prec_names'] = np.array(sorted(avg.keys()))
    out['prec_avg'] = np.array([avg[k] for k in sorted(avg.keys())])
    out['prec_per_query'] = np.array([[per[k][q] for q in range(n)] for k in sorted(avg.keys())])
    out['prec250_names'] = np.array(sorted(avg250.keys()))
    out['prec250_avg'] = np.array([avg250[k] for k in sorted(avg250.keys())])
    out['prec250_per_query'] = np.array([[per250[k][q] for q in range(n)] for k in sorted(avg250.keys())])
    # class-similarity look-up tables of the reference hierarchy (used by the GPU-side metric path)
    wup = np.array([[h.wup_similarity(a, b) for b in range(100)] for a in range(100)])
    lcsh = np.array([[h.lcs_height(a, b) for b in range(100)] for a in range(100)])
    out['wup_lut'] = wup
    out['lcs_height_lut'] = lcsh
    np.savez_compressed(os.path.join(HERE, 'retrieval_ref.npz'), **out)
    print('retrieval fixtures: n=%d' % n, {k: round(float(v), 6) for k, v in avg.items()})


def sgdr():
    from sgdr_callback import SGDR
    import utils as rutils

    class _Opt:
        lr = keras_stub._Var(0.1)

    class _M:
        optimizer = _Opt()

    res = {}
    for tag, kw in (('default', {}), ('short', {'sgdr_base_len': 3, 'sgdr_mul': 2, 'sgdr_max_lr': 0.05})):
        cbs, num_epochs = rutils.get_lr_schedule('SGDR', 50000, 128, dict(kw))
        cb = cbs[0]
        assert isinstance(cb, SGDR)
        cb.model = _M()
        cb.model.optimizer.lr = keras_stub._Var(0.1)          # --sgd_lr default; overwritten at train begin
        cb.on_train_begin()
        seq = []
        for ep in range(num_epochs):
            seq.append(cb.model.optimizer.lr.value)             # lr used during epoch `ep`
            cb.on_epoch_end(ep, {})
        res[tag] = {'num_epochs': int(num_epochs), 'lr': seq, 'args': kw}
    with open(os.path.join(HERE, 'sgdr_ref.json'), 'w') as f:
        json.dump(res, f)
    print('sgdr: default epochs', res['default']['num_epochs'])


def formulas():
    import utils as rutils
    g = torch.Generator().manual_seed(7)
    emb = torch.as_tensor(np.load(os.path.join(HERE, 'class_matrices.npz'))['cifar100_embedding'])
    b = 64
    z = torch.randn(b, 100, generator=g, dtype=torch.float64)
    z[3] = 0.0                                              # the max(sum z^2, 1e-12) edge
    z[4] = 1e-8 * z[4]
    y = torch.randint(0, 100, (b,), generator=g)
    # make a few samples exactly correct so that the accuracy metric has both outcomes
    z[10:20] = emb[y[10:20]] * 3.0 + 0.01 * torch.randn(10, 100, generator=g, dtype=torch.float64)
    x = rutils.l2norm(z)
    t = emb[y]
    out = {
        'z': z.numpy(), 'labels': y.numpy(),
        'l2norm': x.numpy(),
        'inv_correlation': rutils.inv_correlation(t, x).numpy(),
        'squared_distance': rutils.squared_distance(t, z).numpy(),
        'max_sim_acc': rutils.nn_accuracy(emb.numpy(), dot_prod_sim=True)(t, x).numpy(),
        'nn_accuracy': rutils.nn_accuracy(emb.numpy(), dot_prod_sim=False)(t, z).numpy(),
    }
    np.savez_compressed(os.path.join(HERE, 'formulas_ref.npz'), **out)
    print('formulas: max_sim_acc mean %.3f, nn_accuracy mean %.3f' % (out['max_sim_acc'].mean(), out['nn_accuracy'].mean()))


ARCH_CASES = [
    # (tag, architecture, D, batch, hw, loss, cls head?)
    ('simple', 'simple', 100, 4, 32, 'inv_corr', False),
    ('resnet-110-fc', 'resnet-110-fc', 100, 2, 32, 'inv_corr', False),
    ('resnet-110', 'resnet-110', 64, 2, 32, 'inv_corr', False),
    ('resnet-32', 'resnet-32', 64, 2, 32, 'inv_corr', False),
    ('wrn-28-10', 'wrn-28-10', 100, 2, 32, 'inv_corr', True),
]


def architectures():
    import utils as rutils
    import learn_image_embeddings as lie
    import keras
    for tag, arch, dim, batch, hw, loss, with_cls in ARCH_CASES:
        om = omodels.build_network(dim, arch, input_channels=3, seed=11)
        omodels.randomize(om, seed=12)
        weights = dict(om.params)
        cls = None
        if with_cls:
            cls = otrain.ClsHead(dim, 100, seed=13)
            omodels.randomize(cls.params, seed=14)
            # cls_model's BatchNormalization() is the (n+1)-th auto-named BN of the Keras graph
            weights.update({'prob/kernel': cls.params['prob/kernel'], 'prob/bias': cls.params['prob/bias']})
        g = torch.Generator().manual_seed(99)
        x = torch.randn(batch, hw, hw, 3, generator=g, dtype=torch.float64)
        keras_stub.CTX.reset(x=x, weights=weights, training=True)
        if with_cls:
            nbn = sum(1 for k in om.params if k.endswith('/gamma'))
            for s in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                weights['batch_normalization_%d/%s' % (nbn + 1, s)] = cls.params['cls_bn/' + s]
        embed_model = rutils.build_network(dim, arch, input_channels=3)             # learn_image_embeddings.py:125
        model = keras.models.Model(embed_model.inputs,
                                   keras.layers.Lambda(rutils.l2norm, name='l2norm')(embed_model.output))  # :127-128
        outs = {'x': x.numpy(), 'z': embed_model.output.numpy(), 'emb': model.output.numpy()}
        if with_cls:
            model = lie.cls_model(model, 100, None)                                  # :131-132
            outs['prob'] = model.outputs[1].numpy()
        unused = [k for k in weights if k not in keras_stub.CTX.used and not k.endswith('moving_mean')
                  and not k.endswith('moving_variance')]
        assert not unused, unused[:5]
        trace = list(keras_stub.CTX.trace)
        with open(os.path.join(HERE, 'arch_%s.json' % tag), 'w') as f:
            json.dump({'architecture': arch, 'dim': dim, 'seeds': {'build': 11, 'randomize': 12, 'cls': 13, 'cls_rand': 14, 'x': 99},
                       'batch': batch, 'hw': hw, 'with_cls': with_cls, 'trace': trace}, f)
        np.savez_compressed(os.path.join(HERE, 'arch_%s.npz' % tag), **outs)
        nconv = sum(1 for t in trace if t['class'] == 'Conv2D')
        print('%s: %d layers traced (%d convs), |z| mean %.4f' % (tag, len(trace), nconv, np.abs(outs['z']).mean()))


if __name__ == '__main__':
    torch.set_num_threads(8)
    which = sys.argv[1:] or ['class_matrices', 'cifar_hierarchy', 'retrieval', 'sgdr', 'formulas', 'architectures']
    for w in which:
        globals()[w]()
