import hashlib
import numpy as np


class _Key:
    def __init__(self, material):
        self.material = str(material)

    def rng(self):
        h = hashlib.sha256(self.material.encode()).digest()
        return np.random.default_rng(int.from_bytes(h[:8], "little"))

    def fold(self, tag):
        return _Key(self.material + "/" + str(tag))


def PRNGKey(seed):
    return _Key(seed)


def split(key, num=2):
    return [key.fold(i) for i in range(num)]


def normal(key, shape=(), dtype=float):
    return key.rng().standard_normal(shape)


def uniform(key, shape=(), dtype=float, minval=0.0, maxval=1.0):
    return key.rng().uniform(minval, maxval, shape)


def permutation(key, x, independent=False):
    return key.rng().permutation(x)
