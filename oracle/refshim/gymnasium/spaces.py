import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float64, **k):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape if shape is None else tuple(shape)
        self.dtype = np.dtype(dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class Discrete:
    def __init__(self, n, **k):
        self.n = n


class MultiDiscrete:
    def __init__(self, nvec, **k):
        self.nvec = np.asarray(nvec)
