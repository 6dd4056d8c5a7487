"""SGDR learning-rate schedule -- same class name, constructor and per-epoch behaviour as the reference's
Keras callback (sgdr_callback.py:6-87), without Keras: the learning rate lives in `self.lr` and the
training loop pushes it to the device scalar the optimizer kernel reads (Engine.set_lr)."""
import numpy as np


class SGDR(object):
    """lr(i) = min_lr + 0.5 * (max_lr - min_lr) * (1 + cos(pi * (i+1)/num_epochs)) inside a cycle of
    `base_epochs * mul_epochs**cycle` epochs; every cycle starts at max_lr (sgdr_callback.py:63-87)."""

    def __init__(self, min_lr=0.0, max_lr=0.05, base_epochs=10, mul_epochs=2):
        self.min_lr = min_lr
        self.max_lr = max_lr
        self.base_epochs = base_epochs
        self.mul_epochs = mul_epochs
        self.cycles = 0.
        self.cycle_iterations = 0.
        self.trn_iterations = 0.
        self.lr = None

    def _cycle_len(self):
        return self.base_epochs * (self.mul_epochs ** self.cycles)

    def sgdr(self):
        return self.min_lr + 0.5 * (self.max_lr - self.min_lr) * (1 + np.cos(np.pi * (self.cycle_iterations + 1) / self._cycle_len()))

    def on_train_begin(self, logs=None):
        self.lr = self.max_lr if self.cycle_iterations == 0 else self.sgdr()

    def on_epoch_end(self, epoch=None, logs=None):
        if logs is not None:
            logs['lr'] = self.lr
        self.trn_iterations += 1
        self.cycle_iterations += 1
        if self.cycle_iterations >= self._cycle_len():
            self.cycles += 1
            self.cycle_iterations = 0
            self.lr = self.max_lr
        else:
            self.lr = self.sgdr()
