"""The training-loop caller of the hot path: what conv_cINN.py (C) and
conv_pre_training_cINN_on_noise.py (P) do around `model.fit`, with the dataset resident in HBM.

  EarlyStopping(monitor, patience)      C:140-141   tf.keras.callbacks.EarlyStopping (min mode, min_delta 0)
  CSVLogger(filename, separator, append) C:529-536  one row per epoch: epoch, then the metric names sorted (keras order)
  fit(model, xy_train, ...)             C:617-636   epochs / initial_epoch / validation_data / callbacks; shuffles the
                                                    training examples every epoch (C:478-486) on the device
  anneal_and_fit(model, ...)            C:593-636   alpha = i / num_annealing_epochs instance-noise epochs, then clean
                                                    epochs; the same callbacks and epoch counter across both stages
  pretrain_on_noise(model, ...)         P:100-145   every epoch trains (and validates) on fresh N(0,1) batches

Noise is drawn on the device by `instance_noise` / `renew_noise` (Philox, csrc/data_kernels.cu); nothing in these loops
touches the host except the four metric scalars per step that the reference's trackers also read.
"""
import csv
import math
import os

import torch

from .conv_cINN_base_functions import instance_noise, renew_noise

METRICS = ('loss', 'z_loss', 'y_loss', 'detJ_loss')


class EarlyStopping:
    """tf.keras.callbacks.EarlyStopping(monitor='val_loss', patience=p) as the reference configures it (C:140-141):
    stop when `monitor` has not improved (strictly decreased) for `patience` consecutive epochs."""

    def __init__(self, monitor='val_loss', patience=0):
        self.monitor = monitor
        self.patience = int(patience)
        self.best = math.inf
        self.wait = 0
        self.stopped_epoch = None

    def on_train_begin(self):
        # keras resets the wait counter and the best value at the start of every fit() call
        self.best = math.inf
        self.wait = 0
        self.stopped_epoch = None

    def on_epoch_end(self, epoch, logs):
        if self.monitor not in logs:
            return False                       # keras warns and carries on
        cur = logs[self.monitor]
        if cur < self.best:
            self.best = cur
            self.wait = 0
            return False
        self.wait += 1
        if self.wait >= self.patience:
            self.stopped_epoch = epoch
            return True
        return False


class CSVLogger:
    """tf.keras.callbacks.CSVLogger(filename, separator=',', append=True) (C:529-536): header `epoch` + the sorted log
    keys, one row per epoch; with append=True an existing file is continued without a second header."""

    def __init__(self, filename, separator=',', append=False):
        self.filename = filename
        self.sep = separator
        self.append = append
        self.keys = None
        self.append_header = True

    def on_train_begin(self):
        if self.append:
            self.append_header = not (os.path.exists(self.filename) and os.path.getsize(self.filename) > 0)
        else:
            open(self.filename, 'w').close()
            self.append_header = True

    def on_epoch_end(self, epoch, logs):
        if self.keys is None:
            self.keys = sorted(logs.keys())
        with open(self.filename, 'a', newline='') as f:
            w = csv.writer(f, delimiter=self.sep)
            if self.append_header:
                w.writerow(['epoch'] + self.keys)
                self.append_header = False
            w.writerow([epoch] + [logs.get(k, 'NA') for k in self.keys])
        return False


def _batches(xy, batch_size, shuffle, generator=None):
    n = xy.shape[0]
    if shuffle:
        perm = torch.randperm(n, device=xy.device, generator=generator)
        xy = xy[perm]
    for i in range(0, n, batch_size):
        yield xy[i:i + batch_size]


def _run_epoch(model, xy, batch_size, train, transform=None, shuffle=False):
    for m in model.metrics:
        m.reset_state()
    logs = None
    for xb in _batches(xy, batch_size, shuffle):
        if transform is not None:
            xb = transform(xb)
        logs = model.train_step(xb) if train else model.test_step(xb)
    if logs is None:
        raise ValueError("empty dataset")
    return dict(logs)


def fit(model, xy_train, batch_size=32, epochs=1, initial_epoch=0, validation_data=None, callbacks=None, verbose=0,
        shuffle=True, train_transform=None, val_transform=None):
    """keras `model.fit(xy_train, epochs=..., initial_epoch=..., validation_data=..., callbacks=...)` (C:617-636) over a
    device-resident example tensor `xy_train` [N,H,W,D].  `epochs` is the index of the LAST epoch + 1, like keras.
    Returns the History.history dict: {'loss': [...], ..., 'val_loss': [...], ...}."""
    callbacks = list(callbacks or [])
    for cb in callbacks:
        cb.on_train_begin()
    history = {}
    for epoch in range(initial_epoch, epochs):
        logs = _run_epoch(model, xy_train, batch_size, True, train_transform, shuffle)
        if validation_data is not None:
            vlogs = _run_epoch(model, validation_data, batch_size, False, val_transform)
            logs.update({'val_' + k: v for k, v in vlogs.items()})
        for k, v in logs.items():
            history.setdefault(k, []).append(v)
        if verbose:
            print(f"Epoch {epoch + 1}/{epochs} - " + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()))
        stop = False
        for cb in callbacks:
            stop = bool(cb.on_epoch_end(epoch, logs)) or stop
        if stop:
            break
    return history


def anneal_and_fit(model, xy_train, xy_val=None, batch_size=32, num_annealing_epochs=None, num_epochs=500,
                   callbacks=None, verbose=0):
    """C:593-636: `num_annealing_epochs` single-epoch fits on alpha x + (1 - alpha) N(0,1) with alpha = i / N (fresh noise
    for every batch of every epoch, for training AND validation data), then clean training up to epoch `num_epochs`.
    Every annealing epoch is its own fit() call in the reference, so EarlyStopping restarts there (C:617-622).
    Returns the concatenated history."""
    history = {}

    def merge(h):
        for k, v in h.items():
            history.setdefault(k, []).extend(v)

    completed = 0
    for i in range(num_annealing_epochs or 0):
        alpha = i / num_annealing_epochs
        if verbose:
            print(f'Annealing instance noise, alpha={alpha}, annealing epoch {i} of {num_annealing_epochs}.')
        noisy = lambda xb, a=alpha: instance_noise(xb, a)
        merge(fit(model, xy_train, batch_size, epochs=completed + 1, initial_epoch=completed, validation_data=xy_val,
                  callbacks=callbacks, verbose=verbose, train_transform=noisy, val_transform=noisy))
        completed += 1
    merge(fit(model, xy_train, batch_size, epochs=num_epochs, initial_epoch=completed, validation_data=xy_val,
              callbacks=callbacks, verbose=verbose))
    return history


def pretrain_on_noise(model, io_shape, num_train_examples, num_val_examples=0, batch_size=512, num_epochs=1,
                      callbacks=None, verbose=0, device=None):
    """P:100-145: the dataset is N(0,1) noise of shape io_shape, re-drawn every time it is iterated (renew_noise,
    F:660-676).  One buffer per split lives in HBM and is overwritten by the generator kernel each epoch."""
    device = device or model.params.device
    H, W, D = (int(s) for s in io_shape)
    train = torch.empty((num_train_examples, H, W, D), dtype=torch.float32, device=device)
    val = torch.empty((num_val_examples, H, W, D), dtype=torch.float32, device=device) if num_val_examples else None
    return fit(model, train, batch_size, epochs=num_epochs, validation_data=val, callbacks=callbacks, verbose=verbose,
               shuffle=False, train_transform=renew_noise, val_transform=renew_noise if val is not None else None)
