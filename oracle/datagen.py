"""TEST INFRASTRUCTURE - seeded synthetic data in the style of the reference's generator.

The reference generator (reference: test/datagen/gen.py:24-103) is unseeded (``random.random()``) and needs
xarray; this is a seeded numpy/scipy restatement of its "circle" and "curve" patterns:
``arr = 288 + 5*U() + pattern * U() * 5`` on the lcm grid, block-mean coarsened to the input and output
sizes, float32, dims (n, chan, y, x).
"""

import math

import numpy as np
from scipy import ndimage


def _lcm(a, b):
    return a * b // math.gcd(a, b)


def _pattern(name, height, width, mu=1.0):
    if name == "circle":
        y, x = np.meshgrid(np.linspace(-2, 2, width), np.linspace(-3, 3, height))
        d = np.sqrt(y * y + x * x)
        g = np.exp(-((d - mu) ** 2 / (2.0 * 0.2 ** 2)))
        return ndimage.rotate(g, 15)[0:height, 0:width]
    if name == "curve":
        y, x = np.meshgrid(np.linspace(0, 100, width), np.linspace(0, 100, height))
        return np.sqrt((y - 50) ** 2 + (x - 50) ** 2) / math.sqrt(50 ** 2 + 50 ** 2)
    raise ValueError(name)


def _block_mean(arr, out_h, out_w):
    h, w = arr.shape
    return arr.reshape(out_h, h // out_h, out_w, w // out_w).mean(axis=(1, 3))


def generate(n, input_size, output_size, pattern="circle", seed=0):
    """-> (lowres f32[n,1,ih,iw], hires f32[n,1,oh,ow])"""
    rng = np.random.RandomState(seed)
    sh, sw = _lcm(output_size[0], input_size[0]), _lcm(output_size[1], input_size[1])
    base = _pattern(pattern, sh, sw)
    lo = np.zeros((n, 1, input_size[0], input_size[1]), dtype=np.float32)
    hi = np.zeros((n, 1, output_size[0], output_size[1]), dtype=np.float32)
    for i in range(n):
        arr = 288 + 5 * rng.rand() + base * rng.rand() * 5
        lo[i, 0] = _block_mean(arr, *input_size)
        hi[i, 0] = _block_mean(arr, *output_size)
    return lo, hi


class ArrayDataset(dict):
    """the duck-typed 'xarray Dataset' the model API consumes: name -> object with shape/dims/values/data"""

    class _DA:
        def __init__(self, data, dims):
            self.data = data
            self.values = data
            self.shape = data.shape
            self.dims = dims
            self.size = data.size

        def __getitem__(self, idx):
            return ArrayDataset._DA(self.data[idx], self.dims[-np.ndim(self.data[idx]):] if np.ndim(self.data[idx]) else ())

    def add(self, name, data, dims=("n", "chan", "y", "x")):
        self[name] = ArrayDataset._DA(data, dims)
        return self

    def __setitem__(self, name, value):
        if not hasattr(value, "dims"):
            value = ArrayDataset._DA(np.asarray(value), tuple(f"d{i}" for i in range(np.ndim(value))))
        dict.__setitem__(self, name, value)


def circle_datasets(n_train=100, n_test=100, input_size=(16, 16), output_size=(256, 256), pattern="circle"):
    """config 1 of BASELINE.json / SURVEY section 8(d): seed 0 train, seed 1 test"""
    lo, hi = generate(n_train, input_size, output_size, pattern, seed=0)
    tr = ArrayDataset().add("lowres", lo).add("hires", hi)
    lo, hi = generate(n_test, input_size, output_size, pattern, seed=1)
    te = ArrayDataset().add("lowres", lo).add("hires", hi)
    return tr, te
