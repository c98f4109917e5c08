"""Minimal stand-in for the few xarray features the hot path touches.

The reference duck-types its data sets: only ``ds[name]``, ``ds[name] = DataArray(...)``,
``.shape``, ``.dims``, ``.values``, ``.data``, ``.size`` and integer indexing are used
(reference: models/ds_dataset.py:31-67,145-154; models/base_model.py:118-119,151-152).
When xarray is installed the real thing works unchanged; this module exists because the
build image has no xarray.  NetCDF-3 I/O goes through scipy.io.netcdf_file.
"""

from __future__ import annotations

import numpy as np


class DataArray:
    def __init__(self, data, dims=None, attrs=None):
        self.data = np.asarray(data)
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(self.data.ndim))
        self.attrs = dict(attrs or {})

    @property
    def values(self):
        return self.data

    @property
    def shape(self):
        return self.data.shape

    @property
    def size(self):
        return self.data.size

    @property
    def dtype(self):
        return self.data.dtype

    def __getitem__(self, idx):
        sub = self.data[idx]
        return DataArray(sub, dims=tuple(f"dim_{i}" for i in range(np.ndim(sub))))

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


class Dataset:
    def __init__(self, data_vars=None):
        self._vars = {}
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __getitem__(self, name):
        return self._vars[name]

    def __setitem__(self, name, value):
        if not isinstance(value, DataArray):
            value = DataArray(value)
        self._vars[name] = value

    def __contains__(self, name):
        return name in self._vars

    def keys(self):
        return self._vars.keys()

    @property
    def data_vars(self):
        return self._vars

    # ---- NetCDF-3 (classic / 64-bit offset) via scipy
    def to_netcdf(self, path):
        from scipy.io import netcdf_file
        with netcdf_file(path, "w", version=2) as f:
            sizes = {}
            for name, da in self._vars.items():
                for d, n in zip(da.dims, da.shape):
                    if d in sizes and sizes[d] != n:
                        raise ValueError(f"dimension {d} has conflicting sizes {sizes[d]} and {n}")
                    if d not in sizes:
                        sizes[d] = n
                        f.createDimension(d, n)
            for name, da in self._vars.items():
                arr = da.data
                if arr.dtype == np.float64 or arr.dtype == np.float32 or arr.dtype.kind in "iu":
                    pass
                else:
                    arr = arr.astype(np.float32)
                if arr.dtype.kind == "i" and arr.dtype.itemsize == 8:
                    arr = arr.astype(np.int32)
                v = f.createVariable(name, arr.dtype, da.dims)
                v[:] = arr
                for k, a in da.attrs.items():
                    setattr(v, k, a)


def open_dataset(path):
    from scipy.io import netcdf_file
    ds = Dataset()
    with netcdf_file(path, "r", mmap=False) as f:
        for name, var in f.variables.items():
            ds[name] = DataArray(np.array(var[:]), dims=var.dimensions)
    return ds


def open_mfdataset(paths, concat_dim=None, combine=None, **_):
    """concatenate along the first dimension of every variable (the 'case' axis)"""
    if isinstance(paths, str):
        paths = [paths]
    parts = [open_dataset(p) for p in paths]
    if len(parts) == 1:
        return parts[0]
    out = Dataset()
    for name in parts[0].keys():
        out[name] = DataArray(np.concatenate([p[name].data for p in parts], axis=0), dims=parts[0][name].dims)
    return out
