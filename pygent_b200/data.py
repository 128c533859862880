"""Transition store (reference: pygent/data.py `DataSet`): FIFO list of samples with the reference's methods, plus a
columnar view of transition dicts (`x_, u, x, o_, o, c, t`, as stored by pygent/algorithms/mbrl.py:160-167) for the
batched training path."""
import pickle

import numpy as np


class DataSet(object):
    def __init__(self, length, seed=0):
        self.data = list()
        self.length = int(length)
        self.rng = np.random.RandomState(seed)  # same seed on every rank => same shuffles, no broadcast needed

    def add_sample(self, sample):
        exists = sample in self.data
        if not exists:
            self.force_add_sample(sample)
        return exists

    def force_add_sample(self, sample):
        self.data.append(sample)
        if self.length < len(self.data):
            del self.data[0]
        return True

    def rm_sample(self, sample):
        exists = sample in self.data
        if exists:
            self.data.remove(sample)
        return exists

    def random_sample(self):
        return self.data[self.rng.randint(len(self.data))]

    def random_batch(self, n):
        return [self.data[i] for i in self.rng.randint(len(self.data), size=min(n, len(self.data)))]

    def _split(self, idxs, n):
        if len(idxs) == 0:
            return []
        full = max(1, len(idxs) // n)
        parts = [idxs[i * n:(i + 1) * n] for i in range(full)]
        if len(idxs) % n != 0 and len(idxs) > n:
            parts.append(idxs[full * n:])
        return parts

    def batch_indices(self, n, shuffle=True):
        """Index arrays of the batches of one epoch (the last one may be shorter)."""
        idxs = np.arange(len(self.data))
        if shuffle:
            self.rng.shuffle(idxs)
        return self._split(idxs, n)

    def shuffled_batches(self, n):
        return [[self.data[i] for i in idx] for idx in self.batch_indices(n, True)]

    def batches(self, n):
        return [[self.data[i] for i in idx] for idx in self.batch_indices(n, False)]

    def columns(self, keys=("x_", "u", "x", "o_")):
        """{key: float32 array [len, dim]} of transition dicts."""
        return {k: np.asarray([np.asarray(s[k], dtype=np.float64).ravel() for s in self.data]).astype(np.float32) for k in keys}

    def save(self, path):
        with open(path, "wb") as fh:
            pickle.dump(self.data, fh)

    def load(self, path):
        with open(path, "rb") as fh:
            self.data += pickle.load(fh)
