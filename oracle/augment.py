"""Training-input pipeline of the reference restated on the CPU (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

TinyDatasetGenerator.compose_batch (datasets/common.py:771-796): per image
  keras.preprocessing.image.ImageDataGenerator.random_transform  (horizontal_flip, width/height_shift_range = 0.15,
                                                                   fill_mode 'nearest'; datasets/common.py:640)
  keras.preprocessing.image.ImageDataGenerator.standardize       (featurewise_center / _std_normalization; :639)
Keras 2.2 ships keras_preprocessing 1.0.x, whose apply_affine_transform calls
scipy.ndimage.affine_transform(channel, matrix, offset, order=1, mode='nearest') per channel and flips afterwards;
standardize is x -= mean; x /= (std + K.epsilon()) with K.epsilon() = 1e-7.  Keras itself is not installable here
(SURVEY.md section 8c), so this restatement is pinned to scipy -- the library Keras delegates the resampling to.
"""
import numpy as np
from scipy import ndimage


def fit_statistics(X_train):
    """ImageDataGenerator.fit with featurewise_center / featurewise_std_normalization: per-channel mean and std over
    (N, H, W) of the float32 training images (datasets/common.py:666-670)."""
    x = np.asarray(X_train, dtype=np.float32)
    mean = x.mean(axis=(0, 1, 2), dtype=np.float64)
    std = np.sqrt(((x.astype(np.float64) - mean) ** 2).mean(axis=(0, 1, 2)))
    return mean.astype(np.float32), std.astype(np.float32)


def draw_transform(rng, h, w, shift=0.15):
    """One image's random parameters in the order ImageDataGenerator.get_random_transform draws them."""
    tx = rng.uniform(-shift, shift) * h          # height_shift_range: rows
    ty = rng.uniform(-shift, shift) * w          # width_shift_range: columns
    flip = rng.random_sample() < 0.5
    return tx, ty, bool(flip)


def random_transform(x, tx, ty, flip):
    """apply_affine_transform(x, tx=tx, ty=ty, fill_mode='nearest') then flip_axis(x, column axis)."""
    x = np.asarray(x, dtype=np.float32)
    matrix = np.eye(2)
    out = np.stack([ndimage.affine_transform(x[..., c], matrix, offset=(tx, ty), order=1, mode='nearest')
                    for c in range(x.shape[-1])], axis=-1)
    if flip:
        out = out[:, ::-1]
    return out.astype(np.float32)


def standardize(x, mean, std):
    return (x - mean) / (std + 1e-7)


def compose_batch(X, indices, params, mean, std):
    """datasets/common.py:788-794 for the given per-image (tx, ty, flip)."""
    return np.stack([standardize(random_transform(X[j], *params[i]), mean, std) for i, j in enumerate(indices)]).astype(np.float32)
