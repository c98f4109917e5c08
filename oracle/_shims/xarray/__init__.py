"""TEST INFRASTRUCTURE - file-based stand-in so the reference's modules (which `import xarray as xr` at
module top) import in an image without xarray.  Only DataArray construction is ever reached."""
import numpy as np


class DataArray:
    def __init__(self, data=None, dims=None, **kw):
        self.data = np.asarray(data)
        self.values = self.data
        self.dims = tuple(dims) if dims is not None else ()
        self.shape = self.data.shape
        self.size = self.data.size

    def __getitem__(self, idx):
        return DataArray(self.data[idx])


class Dataset(dict):
    pass


def open_dataset(*a, **k):
    raise NotImplementedError("xarray shim: no file I/O")


open_mfdataset = open_dataset
