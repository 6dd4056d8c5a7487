"""CPU tests: the oracle (oracle/) against fixtures produced by executing the reference's own
Python in the build container (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import models as omodels
from oracle import nn as onn
from oracle import retrieval as oret
from oracle import train as otrain

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_class_matrices_are_unit_norm_and_hierarchical():
    d = np.load(os.path.join(G, 'class_matrices.npz'))
    e = d['cifar100_embedding']
    assert e.shape == (100, 100) and d['nab_embedding'].shape == (555, 555)
    np.testing.assert_allclose(np.linalg.norm(e, axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(d['nab_embedding'], axis=1), 1.0, atol=1e-12)
    sims = np.unique(np.round(e @ e.T * 8).astype(int))
    assert set(sims.tolist()) <= set(range(9))            # E E^T in {0, 1/8, ..., 1}
    np.testing.assert_allclose(e @ e.T * 8, np.round(e @ e.T * 8), atol=1e-9)


def test_retrieval_rankings_match_reference():
    d = np.load(os.path.join(G, 'retrieval_ref.npz'))
    feat, fu = d['feat'], d['feat_unit']
    for key, f, norm in (('rank_sq', feat, False), ('rank_cos', feat, True), ('rank_sq_unit', fu, False)):
        pd = oret.pairwise_dist_ref32(f, norm)
        ours = oret.rank_stable(pd)
        ref = d[key]
        # positions may differ only inside exact float32 ties (np.argsort is not stable)
        diff = ours != ref
        if diff.any():
            dv = np.take_along_axis(pd, ours, -1)
            dr = np.take_along_axis(pd, ref.astype(np.int64), -1)
            np.testing.assert_array_equal(dv, dr)
        assert diff.mean() < 1e-3
    # dict input: ids are mapped through ind2id (evaluate_retrieval.py:43-50,69-70)
    fd = {int(i): f for i, f in zip(d['ids'], feat)}
    r = oret.pairwise_retrieval({'feat': fd}, normalize=False)
    assert list(r.keys()) == d['rank_dict_keys'].tolist()
    got = np.array(list(r.values()))
    assert (got != d['rank_dict_vals']).mean() < 1e-3


def test_sgdr_schedule_matches_reference():
    with open(os.path.join(G, 'sgdr_ref.json')) as f:
        ref = json.load(f)
    assert otrain.sgdr_default_epochs() == ref['default']['num_epochs'] == 372
    np.testing.assert_allclose(otrain.sgdr_lr_sequence(372), ref['default']['lr'], rtol=1e-14)
    s = ref['short']
    np.testing.assert_allclose(otrain.sgdr_lr_sequence(s['num_epochs'], 1e-6, 0.05, 3, 2), s['lr'], rtol=1e-14)


def test_head_formulas_match_reference():
    d = np.load(os.path.join(G, 'formulas_ref.npz'))
    emb = torch.as_tensor(np.load(os.path.join(G, 'class_matrices.npz'))['cifar100_embedding'])
    z = torch.as_tensor(d['z'])
    y = torch.as_tensor(d['labels'])
    x = onn.l2norm(z)
    t = emb[y]
    np.testing.assert_allclose(x.numpy(), d['l2norm'], rtol=0, atol=1e-15)
    np.testing.assert_allclose(onn.inv_correlation(t, x).numpy(), d['inv_correlation'], atol=1e-15)
    np.testing.assert_allclose(onn.squared_distance(t, z).numpy(), d['squared_distance'], rtol=1e-15)
    np.testing.assert_array_equal(onn.max_sim_acc(emb, t, x).numpy(), d['max_sim_acc'])
    np.testing.assert_array_equal(onn.nn_accuracy(emb, t, z).numpy(), d['nn_accuracy'])
    assert 0 < d['max_sim_acc'].mean() < 1


@pytest.mark.parametrize('tag', ['simple', 'resnet-110-fc', 'resnet-110', 'resnet-32', 'wrn-28-10'])
def test_oracle_architecture_matches_reference_graph(tag):
    """The reference's own model-building code (run under the eager Keras stub) and oracle/models.py
    must give the same output for the same weights: pins topology, layer order, names, padding, strides."""
    with open(os.path.join(G, 'arch_%s.json' % tag)) as f:
        meta = json.load(f)
    d = np.load(os.path.join(G, 'arch_%s.npz' % tag))
    s = meta['seeds']
    m = omodels.build_network(meta['dim'], meta['architecture'], input_channels=3, seed=s['build'])
    omodels.randomize(m, seed=s['randomize'])
    x = torch.as_tensor(d['x'])
    with torch.no_grad():
        z = m.forward(x, training=True)
        np.testing.assert_allclose(z.numpy(), d['z'], rtol=1e-9, atol=1e-10)
        emb = otrain.head_forward(z, 'inv_corr')
        np.testing.assert_allclose(emb.numpy(), d['emb'], rtol=1e-9, atol=1e-11)
        if meta['with_cls']:
            cls = otrain.ClsHead(meta['dim'], 100, seed=s['cls'])
            omodels.randomize(cls.params, seed=s['cls_rand'])
            prob = cls.forward(cls.params, emb, training=True)
            np.testing.assert_allclose(prob.numpy(), d['prob'], rtol=1e-9, atol=1e-12)
    # hyper-parameters recorded from the reference's layer constructors
    convs = [t for t in meta['trace'] if t['class'] == 'Conv2D']
    for t in convs:
        k = m.params[t['name'] + '/kernel']
        assert list(k.shape[:2]) == t['kernel_size'] and k.shape[3] == t['filters']
        assert (t['name'] + '/bias' in m.params) == t['use_bias']
        assert m.l2.get(t['name'] + '/kernel', 0.0) == t['l2']
    for t in meta['trace']:
        if t['class'] == 'BatchNormalization' and t['name'] in m.bn_cfg:
            assert m.bn_cfg[t['name']] == (t['momentum'], t['epsilon'])
        if t['class'] == 'Dense' and t['name'] + '/kernel' in m.params:
            assert m.l2.get(t['name'] + '/kernel', 0.0) == t['l2']
    nparams = sum(p.numel() for n, p in m.params.items() if n in m.trainable)
    expected = {'simple': None, 'resnet-110-fc': None}
    assert nparams > 0


def test_same_padding_is_asymmetric_for_stride2():
    assert onn.same_pad(32, 3, 1) == (1, 1, 32)
    assert onn.same_pad(32, 3, 2) == (0, 1, 16)
    assert onn.same_pad(32, 1, 2) == (0, 0, 16)
    assert onn.same_pad(224, 7, 2) == (2, 3, 112)


def test_sgd_clip_and_momentum():
    p = {'w': torch.tensor([1.0, 2.0], dtype=torch.float64)}
    g = {'w': torch.tensor([30.0, 40.0], dtype=torch.float64)}      # norm 50 >= 10 -> scaled by 0.2
    v = {'w': torch.zeros(2, dtype=torch.float64)}
    n = otrain.sgd_step(p, g, v, lr=0.1, clipnorm=10.0)
    assert n == 50.0
    np.testing.assert_allclose(v['w'].numpy(), [-0.6, -0.8])
    np.testing.assert_allclose(p['w'].numpy(), [0.4, 1.2])
    otrain.sgd_step(p, g, v, lr=0.1, clipnorm=10.0, nesterov=True)
    np.testing.assert_allclose(v['w'].numpy(), [-1.14, -1.52])
    np.testing.assert_allclose(p['w'].numpy(), [0.4 + 0.9 * -1.14 - 0.6, 1.2 + 0.9 * -1.52 - 0.8])


def test_hierarchical_precision_oracle_matches_reference_metrics():
    """oracle/hierarchy.py (the oracle of SURVEY.md section 8(f) rank 2) against ClassHierarchy.hierarchical_precision
    of the reference itself (class_hierarchy.py:211-316), run by tests/golden/make_golden.py on the reference's own
    rankings of the 256-item fixture: every metric of every query, clipped (compute_ahp=250) and unclipped + AP."""
    from oracle import hierarchy as ohier
    d = np.load(os.path.join(G, 'retrieval_ref.npz'))
    rank, labels = d['rank_sq_unit'].astype(np.int64), d['labels']
    avg, per = ohier.hierarchical_precision(rank, labels, d['wup_lut'], d['lcs_height_lut'], compute_ahp=True, compute_ap=True)
    for k, name in enumerate(d['prec_names']):
        np.testing.assert_allclose(per[str(name)], d['prec_per_query'][k], rtol=0, atol=1e-12, err_msg=str(name))
        assert abs(avg[str(name)] - d['prec_avg'][k]) < 1e-12
    avg, per = ohier.hierarchical_precision(rank, labels, d['wup_lut'], d['lcs_height_lut'], compute_ahp=250)
    assert sorted(per.keys()) == [str(n) for n in d['prec250_names']]
    for k, name in enumerate(d['prec250_names']):
        np.testing.assert_allclose(per[str(name)], d['prec250_per_query'][k], rtol=0, atol=1e-12, err_msg=str(name))
        assert abs(avg[str(name)] - d['prec250_avg'][k]) < 1e-12
    # a ranking truncated to clip+1 items gives the same clipped metrics except through the ideal gain, which needs the
    # label histogram of the whole database: the reason pairwise_retrieval(topk=...) relies on `all_ids` completion
    assert rank.shape[1] > 251
