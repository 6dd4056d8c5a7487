"""Host-side mirror of the reference's in-memory dataset interface (datasets/common.py:635-844 TinyDatasetGenerator,
datasets/cifar.py:43-81 CifarGenerator, datasets/__init__.py:21-166 get_data_generator) for the hot path.

The images stay resident in device memory (CIFAR-100: 150 MB as uint8); a training batch is gathered, augmented
(horizontal flip, +-15 % shifts with linear resampling and nearest fill) and standardised by ONE CUDA launch
(se_augment_batch, csrc/augment.cu) straight into the engine's input tensor -- the replacement of
TinyDatasetGenerator.compose_batch's per-image Python loop around Keras' ImageDataGenerator (datasets/common.py:771-796).
Only the random draws (three numbers per image) are made on the host.  There is no CPU implementation of the transform
in the product; tests compare the kernel with oracle/augment.py.
"""
import os
import pickle

import numpy as np

from . import _lib


class TinyDatasetGenerator:
    """datasets/common.py:635-844.  X_*: (n, H, W, C) uint8 or float arrays of raw pixel values, y_*: integer labels."""

    def __init__(self, X_train, X_test, y_train, y_test, device='cuda'):
        import torch
        self.dev = torch.device(device)
        self.y_train, self.y_test = np.asarray(y_train), np.asarray(y_test)
        self.shape = tuple(X_train.shape[1:])
        self.is_u8 = X_train.dtype == np.uint8 and X_test.dtype == np.uint8
        dt = np.uint8 if self.is_u8 else np.float32
        self.X_train = torch.from_numpy(np.ascontiguousarray(X_train, dtype=dt)).to(self.dev)
        self.X_test = torch.from_numpy(np.ascontiguousarray(X_test, dtype=dt)).to(self.dev)
        # ImageDataGenerator.fit (datasets/common.py:666-670): per-channel mean / std of the training images; reduced on the
        # device in float64 (a one-off reduction: torch is the container library here, not the hot path)
        xt = self.X_train.to(torch.float64)
        self.mean = xt.mean(dim=(0, 1, 2)).to(torch.float32)
        self.std = xt.std(dim=(0, 1, 2), unbiased=False).to(torch.float32)
        del xt
        self.inv_std = (1.0 / (self.std + 1e-7)).contiguous()          # standardize: x /= (std + K.epsilon())
        self.mean = self.mean.contiguous()
        self._buf = {}

    # ---- properties of the reference interface
    @property
    def labels_train(self):
        return self.y_train

    @property
    def labels_test(self):
        return self.y_test

    @property
    def num_classes(self):
        return int(max(self.y_train.max(), self.y_test.max())) + 1

    @property
    def num_train(self):
        return int(self.X_train.shape[0])

    @property
    def num_test(self):
        return int(self.X_test.shape[0])

    @property
    def num_channels(self):
        return int(self.shape[-1])

    # ---- batches
    def _params(self, n):
        import torch
        if n not in self._buf:
            self._buf[n] = (torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory(),
                            torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(),
                            [torch.empty(n, dtype=t, device=self.dev) for t in (torch.int32, torch.float32, torch.float32, torch.uint8)])
        return self._buf[n]

    def compose_batch(self, indices, train, out, augment=False, rng=None, params=None):
        """datasets/common.py:771-796 on the device: writes the standardised (and, for augment=True, randomly flipped /
        shifted) images `indices` of the training or test set into `out` (a (B, H, W, C) float32 CUDA tensor).
        params: optional (tx, ty, flip) arrays instead of fresh random draws (tests)."""
        import torch
        n = len(indices)
        H, W, C = self.shape
        hi, htx, hty, hfl, (di, dtx, dty, dfl) = self._params(n)
        hi.numpy()[:] = np.asarray(indices, dtype=np.int32)
        di.copy_(hi, non_blocking=True)
        if augment:
            if params is None:
                rng = rng or np.random
                # the order ImageDataGenerator.get_random_transform draws in: row shift, column shift, flip -- per image
                draws = rng.random_sample((n, 3))
                params = ((draws[:, 0] * 0.3 - 0.15) * H, (draws[:, 1] * 0.3 - 0.15) * W, draws[:, 2] < 0.5)
            htx.numpy()[:] = params[0]
            hty.numpy()[:] = params[1]
            hfl.numpy()[:] = np.asarray(params[2], dtype=np.uint8)
            dtx.copy_(htx, non_blocking=True)
            dty.copy_(hty, non_blocking=True)
            dfl.copy_(hfl, non_blocking=True)
        src = self.X_train if train else self.X_test
        with torch.cuda.device(self.dev):
            _lib.call('se_augment_batch', src.data_ptr(), 1 if self.is_u8 else 0, di.data_ptr(),
                      dtx.data_ptr() if augment else None, dty.data_ptr() if augment else None,
                      dfl.data_ptr() if augment else None, self.mean.data_ptr(), self.inv_std.data_ptr(), out.data_ptr(),
                      n, H, W, C, _lib.stream_ptr())
        return out

    def train_batches(self, batch_size, rng, rank=0, world=1):
        """One epoch of shuffled, augmented training batches (DataSequence with shuffle, datasets/common.py:26-122; a
        trailing partial batch is dropped so that the per-GPU batch of the launch plans stays fixed).  Yields
        (indices of this rank's slice, labels of the slice); the images go to the tensor given to `compose_batch`."""
        perm = rng.permutation(self.num_train)
        per = batch_size // world
        for i in range(0, self.num_train - batch_size + 1, batch_size):
            idx = perm[i + rank * per:i + (rank + 1) * per]
            yield idx, self.y_train[idx]

    def test_batches(self, batch_size):
        for i in range(0, self.num_test, batch_size):
            idx = np.arange(i, min(i + batch_size, self.num_test))
            yield idx, self.y_test[idx]


def _load_cifar(data_root, name, classes):
    def load(fn):
        with open(os.path.join(data_root, fn), 'rb') as f:
            d = pickle.load(f, encoding='bytes')
        X = d[b'data'].reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1)           # datasets/cifar.py:80-81
        y = np.asarray(d[b'fine_labels'] if b'fine_labels' in d else d[b'labels'])
        return np.ascontiguousarray(X), y
    if name == 'cifar-100':
        Xtr, ytr = load('train')
        Xte, yte = load('test')
    else:
        parts = [load('data_batch_%d' % i) for i in range(1, 6)]
        Xtr, ytr = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
        Xte, yte = load('test_batch')
    if classes is not None:                                                   # datasets/cifar.py:59-77: subset + re-enumeration
        lut = {c: i for i, c in enumerate(classes)}
        keep = np.array([c in lut for c in ytr])
        Xtr, ytr = Xtr[keep], np.array([lut[c] for c in ytr[keep]])
        keep = np.array([c in lut for c in yte])
        Xte, yte = Xte[keep], np.array([lut[c] for c in yte[keep]])
    return Xtr, Xte, ytr, yte


def get_data_generator(dataset, data_root, classes=None, device='cuda'):
    """datasets/__init__.py:21-166, CIFAR branch (:85-87), plus 'synthetic[:n]' (uint8 images = a fixed random colour
    template per class blended with pixel noise, for machines without data).  The other dataset parsers of the reference are host-side file readers outside the
    hot path."""
    name = dataset.lower()
    if name in ('cifar-100', 'cifar-10'):
        return TinyDatasetGenerator(*_load_cifar(data_root, name, classes), device=device)
    if name.startswith('synthetic'):
        rng = np.random.RandomState(0)
        ncls = len(classes) if classes is not None else 100
        n = int(name.split(':')[1]) if ':' in name else 2048
        # every class has a fixed random colour template; an image is its class template blended with pixel noise, so
        # that embeddings of different images are distinct and a short run has something to learn
        templates = rng.randint(0, 256, (ncls, 4, 4, 3)).repeat(8, axis=1).repeat(8, axis=2).astype(np.float32)
        ytr, yte = rng.randint(0, ncls, n), rng.randint(0, ncls, max(n // 4, 1))
        make = lambda y: np.clip(0.6 * templates[y] + 0.4 * rng.randint(0, 256, (len(y), 32, 32, 3)), 0, 255).astype(np.uint8)
        return TinyDatasetGenerator(make(ytr), make(yte), ytr, yte, device=device)
    raise ValueError('Unknown dataset: {}'.format(dataset))
