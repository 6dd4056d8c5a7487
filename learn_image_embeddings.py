#!/usr/bin/env python
"""Drop-in for the reference's learn_image_embeddings.py (same flags, same pickles) on the B200-native engine.

Reference: learn_image_embeddings.py:54-275.  Differences, all outside the accelerated hot path:
  * model / weight dumps are pickles of {Keras weight name: array} instead of Keras HDF5 (h5py is not a
    dependency; names and layouts are Keras', so they convert 1:1);
  * datasets: 'CIFAR-100' / 'CIFAR-10' (python pickles, datasets/cifar.py:43-81) and 'synthetic' (N(0,1) images
    for machines without data).  The other dataset parsers are host-side file readers;
  * --loss softmax_corr, --finetune, --log_dir are accepted but rejected / ignored with a message;
  * --gpus N > 1: launch with `python -m torch.distributed.run --nproc-per-node N learn_image_embeddings.py ...`
    (one process per GPU, NCCL all-reduce) instead of in-graph towers.
"""
import argparse
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from semantic_embeddings_b200 import utils  # noqa: E402


class TinyData:
    """In-memory image dataset with the augmentation of datasets/common.py:638-670,771-796 (TinyDatasetGenerator):
    featurewise mean/std from the training set, random horizontal flip, +-15% shifts with nearest fill."""

    def __init__(self, X_train, y_train, X_test, y_test, classes=None):
        if classes is not None:                                  # datasets/cifar.py:59-77: subset + re-enumeration
            lut = {c: i for i, c in enumerate(classes)}
            keep = np.array([c in lut for c in y_train])
            X_train, y_train = X_train[keep], np.array([lut[c] for c in y_train[keep]])
            keep = np.array([c in lut for c in y_test])
            X_test, y_test = X_test[keep], np.array([lut[c] for c in y_test[keep]])
        self.X_train, self.y_train = X_train.astype(np.float32), np.asarray(y_train)
        self.X_test, self.y_test = X_test.astype(np.float32), np.asarray(y_test)
        self.mean = self.X_train.mean(axis=(0, 1, 2))
        self.std = self.X_train.std(axis=(0, 1, 2))
        self.num_classes = int(max(self.y_train.max(), self.y_test.max())) + 1
        self.num_train, self.num_test = len(self.X_train), len(self.X_test)
        self.num_channels = self.X_train.shape[-1]
        self.labels_test = self.y_test

    def standardize(self, X):
        return (X - self.mean) / (self.std + 1e-6)

    def train_batches(self, batch_size, rng):
        perm = rng.permutation(self.num_train)
        n, h, w = self.num_train, self.X_train.shape[1], self.X_train.shape[2]
        for i in range(0, n - batch_size + 1, batch_size):
            idx = perm[i:i + batch_size]
            X = self.standardize(self.X_train[idx])
            flip = rng.rand(len(idx)) < 0.5
            X[flip] = X[flip, :, ::-1]
            dy = rng.uniform(-0.15, 0.15, len(idx)) * h
            dx = rng.uniform(-0.15, 0.15, len(idx)) * w
            for k in range(len(idx)):                           # nearest-fill shift
                ry = np.clip(np.arange(h) + int(round(dy[k])), 0, h - 1)
                rx = np.clip(np.arange(w) + int(round(dx[k])), 0, w - 1)
                X[k] = X[k][ry][:, rx]
            yield np.ascontiguousarray(X, dtype=np.float32), self.y_train[idx]

    def test_batches(self, batch_size):
        for i in range(0, self.num_test, batch_size):
            yield self.standardize(self.X_test[i:i + batch_size]).astype(np.float32), self.y_test[i:i + batch_size]


def get_data_generator(dataset, data_root, classes=None):
    """datasets/__init__.py:21-166, CIFAR branch (:85-87) + a synthetic stand-in."""
    name = dataset.lower()
    if name in ('cifar-100', 'cifar-10'):
        def load(fn):
            with open(os.path.join(data_root, fn), 'rb') as f:
                d = pickle.load(f, encoding='bytes')
            X = d[b'data'].reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1)       # datasets/cifar.py:80-81
            y = np.asarray(d[b'fine_labels'] if b'fine_labels' in d else d[b'labels'])
            return X, y
        if name == 'cifar-100':
            Xtr, ytr = load('train')
            Xte, yte = load('test')
        else:
            parts = [load('data_batch_%d' % i) for i in range(1, 6)]
            Xtr, ytr = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
            Xte, yte = load('test_batch')
        return TinyData(Xtr, ytr, Xte, yte, classes)
    if name.startswith('synthetic'):
        rng = np.random.RandomState(0)
        ncls = len(classes) if classes is not None else 100
        n = int(name.split(':')[1]) if ':' in name else 2048
        return TinyData(rng.randn(n, 32, 32, 3) * 60 + 120, rng.randint(0, ncls, n),
                        rng.randn(n // 4, 32, 32, 3) * 60 + 120, rng.randint(0, ncls, n // 4), None)
    raise ValueError('Unknown dataset: {}'.format(dataset))


def main():
    parser = argparse.ArgumentParser(description='Learns to map images onto class embeddings.',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    g = parser.add_argument_group('Data parameters')
    g.add_argument('--dataset', type=str, required=True)
    g.add_argument('--data_root', type=str, required=True)
    g.add_argument('--embedding', type=str, required=True)
    g = parser.add_argument_group('Training parameters')
    g.add_argument('--architecture', type=str, default='simple', choices=utils.REFERENCE_ARCHITECTURES)
    g.add_argument('--loss', type=str, default='inv_corr', choices=['mse', 'inv_corr', 'unnorm_corr', 'softmax_corr'])
    g.add_argument('--cls_weight', type=float, default=0.0)
    g.add_argument('--cls_base', type=str, default=None)
    g.add_argument('--lr_schedule', type=str, default='SGDR', choices=utils.LR_SCHEDULES)
    g.add_argument('--clipgrad', type=float, default=10.0)
    g.add_argument('--max_decay', type=float, default=0.0)
    g.add_argument('--nesterov', action='store_true', default=False)
    g.add_argument('--epochs', type=int, default=None)
    g.add_argument('--batch_size', type=int, default=100)
    g.add_argument('--val_batch_size', type=int, default=None)
    g.add_argument('--snapshot', type=str, default=None)
    g.add_argument('--snapshot_best', type=str, nargs='?', default=None, const='val_loss')
    g.add_argument('--initial_epoch', type=int, default=0)
    g.add_argument('--finetune', type=str, default=None)
    g.add_argument('--finetune_init', type=int, default=8)
    g.add_argument('--gpus', type=int, default=1)
    g.add_argument('--read_workers', type=int, default=8)
    g.add_argument('--queue_size', type=int, default=100)
    g.add_argument('--gpu_merge', action='store_true', default=False)
    g = parser.add_argument_group('Output parameters')
    g.add_argument('--model_dump', type=str, default=None)
    g.add_argument('--weight_dump', type=str, default=None)
    g.add_argument('--feature_dump', type=str, default=None)
    g.add_argument('--log_dir', type=str, default=None)
    g.add_argument('--no_progress', action='store_true', default=False)
    g.add_argument('--top_k_acc', type=int, nargs='+', default=[])
    g.add_argument('--arith', type=str, default='tf32', choices=['tf32', 'f32'],
                   help='(new) tensor-core fast mode or fp32 parity mode')
    utils.add_lr_schedule_arguments(parser)
    args = parser.parse_args()
    if args.val_batch_size is None:
        args.val_batch_size = args.batch_size
    if args.cls_base is not None or args.max_decay > 0 or args.finetune or args.loss == 'softmax_corr' or args.top_k_acc:
        raise NotImplementedError('--cls_base / --max_decay / --finetune / --loss softmax_corr / --top_k_acc are outside '
                                  'the accelerated hot path (SURVEY.md section 8)')

    import torch
    from semantic_embeddings_b200 import _lib
    from semantic_embeddings_b200.engine import Engine
    from semantic_embeddings_b200.parallel import broadcast_parameters, init_process_group, shard_batch

    # class embeddings (learn_image_embeddings.py:104-117)
    if args.embedding == 'onehot':
        embed_labels, embedding = None, None
    else:
        with open(args.embedding, 'rb') as pf:
            emb = pickle.load(pf)
        embed_labels, embedding = emb['ind2label'], emb['embedding']
    data = get_data_generator(args.dataset, args.data_root, classes=embed_labels)
    if embedding is None:
        embedding = np.eye(data.num_classes)

    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    rank, world = init_process_group(device=torch.device('cuda', local))
    if world != max(1, args.gpus) and rank == 0:
        print('note: --gpus {} but {} process(es) were launched; using {}'.format(args.gpus, world, world))
    start, per_gpu = shard_batch(args.batch_size, world, rank)

    graph = utils.build_network(embedding.shape[1], args.architecture, input_channels=data.num_channels)
    mode = _lib.SE_MODE_TF32 if args.arith == 'tf32' else _lib.SE_MODE_F32
    eng = Engine(graph, args.batch_size // world, embedding, loss=args.loss, cls_weight=args.cls_weight,
                 num_classes=data.num_classes, mode=mode, device='cuda:%d' % local, nesterov=args.nesterov,
                 clipnorm=args.clipgrad, world_size=world)
    if args.snapshot and os.path.exists(args.snapshot):
        print('Resuming from snapshot {}'.format(args.snapshot))
        with open(args.snapshot, 'rb') as f:
            snap = pickle.load(f)
        eng.set_weights(snap['weights'])
        eng.set_velocity(snap['velocity'])
    broadcast_parameters([eng.P, eng.S, eng.V])

    callbacks, num_epochs = utils.get_lr_schedule(args.lr_schedule, data.num_train, args.batch_size,
                                                  schedule_args={k: v for k, v in vars(args).items() if v is not None})
    sched = callbacks[0]
    sched.on_train_begin()
    for _ in range(args.initial_epoch):
        sched.on_epoch_end()
    epochs = args.epochs if args.epochs else num_epochs
    rng = np.random.RandomState(1234)
    pb = args.batch_size // world
    for epoch in range(args.initial_epoch, epochs):
        eng.set_lr(sched.lr)
        tot_loss = tot_acc = nb = 0
        for X, y in data.train_batches(args.batch_size, rng):
            eng.train_step(torch.from_numpy(X[rank * pb:(rank + 1) * pb]), torch.from_numpy(y[rank * pb:(rank + 1) * pb]))
            if not args.no_progress and nb % 50 == 0:
                m = eng.metrics()
                tot_loss += m['loss']; tot_acc += m['acc']
            nb += 1
        if rank == 0:
            m = eng.metrics()
            print('Epoch {}/{} - lr {:.6f} - loss {:.4f} - acc {:.4f}'.format(epoch + 1, epochs, sched.lr, m['loss'], m['acc']))
            if args.snapshot:
                with open(args.snapshot, 'wb') as f:
                    pickle.dump({'weights': eng.get_weights(), 'velocity': eng.get_velocity(), 'epoch': epoch + 1}, f)
        sched.on_epoch_end(epoch)

    if rank == 0:
        if args.weight_dump or args.model_dump:
            for fn in (args.weight_dump, args.model_dump):
                if fn:
                    with open(fn, 'wb') as f:
                        pickle.dump({'architecture': args.architecture, 'weights': eng.get_weights()}, f)
        if args.feature_dump:                                  # learn_image_embeddings.py:270-275
            feats = []
            for X, _ in data.test_batches(pb):
                n = len(X)
                if n < pb:
                    X = np.concatenate([X, np.zeros((pb - n,) + X.shape[1:], np.float32)])
                feats.append(eng.predict(torch.from_numpy(X))[:n])
            feats = np.concatenate(feats)
            with open(args.feature_dump, 'wb') as dump_file:
                pickle.dump({'feat': dict(enumerate(feats))}, dump_file)


if __name__ == '__main__':
    main()
