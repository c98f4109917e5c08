"""Data adaptor: named 4-D arrays (N,C,Y,X) -> normalised fp32 batches.

Semantics follow the reference ``DSDataset`` (reference: src/cae_tools/models/ds_dataset.py:20-159):
global min/max per input variable and for the output, NaNs rejected, inputs concatenated along the
channel axis, items are ``(input, output, mask, label)``.  In addition to the per-item protocol this
class assembles whole (optionally permuted) arrays in one vectorised pass - that is what feeds the
device-resident batches of the engines (the reference collates item by item).
"""

from __future__ import annotations

import numpy as np
import torch


class DSDataset(torch.utils.data.Dataset):

    def __init__(self, ds, input_variable_names, output_variable_name=None, normalise_in=True, normalise_out=True,
                 mask_variable_name=None):
        self.ds = ds
        self.input_variable_names = input_variable_names
        self.output_variable_name = output_variable_name
        self.normalise_in = normalise_in
        self.normalise_out = normalise_out
        self.input_spec = []
        self.output_spec = None
        self.input_das = [ds[name] for name in input_variable_names]
        first = self.input_das[0]
        self.n = first.shape[0]
        self.input_chan = sum(da.shape[1] for da in self.input_das)
        self.input_y, self.input_x = first.shape[2], first.shape[3]
        self.mask_da = ds[mask_variable_name] if mask_variable_name is not None else None

        self._raw = {}      # variable name -> raw fp32 CUDA tensor (device ingest path)
        out_values = np.asarray(ds[output_variable_name].values)
        out_lo, out_hi, bad = self._scan(output_variable_name, out_values)
        if bad > 0:
            raise ValueError(f"output variable contains {bad} NaN values")

        self.min_inputs, self.max_inputs = {}, {}
        for name, da in zip(input_variable_names, self.input_das):
            values = np.asarray(da.values)
            self.min_inputs[name], self.max_inputs[name], bad = self._scan(name, values)
            if bad > 0:
                raise ValueError(f"input variable {name} contains {bad} NaN values")
            self.input_spec.append({"name": name, "shape": list(da.shape[1:])})

        if output_variable_name:
            self.output_da = ds[output_variable_name]
            self.output_chan, self.output_y, self.output_x = self.output_da.shape[1:4]
            self.min_output, self.max_output = out_lo, out_hi
            self.output_spec = {"name": output_variable_name, "shape": list(self.output_da.shape[1:])}
        else:
            self.output_da = None
            self.output_chan = self.output_y = self.output_x = None
            self.min_output = self.max_output = None

    # ---- device ingest: raw fp32 arrays are uploaded once; the min / max / NaN scan, the min-max
    # normalisation and the assembly of the shuffled batches run as HBM-bound kernels (csrc/ingest.cu) instead of numpy
    # passes on the host.  Falls back to numpy without a CUDA device, for non-fp32 data and for arrays beyond DEVICE_BUDGET.
    DEVICE_BUDGET = 16 << 30

    def _device_ok(self, values):
        return torch.cuda.is_available() and values.dtype == np.float32 and values.ndim == 4 and \
            0 < values.nbytes <= self.DEVICE_BUDGET

    def _scan(self, name, values):
        """(min, max, NaN count) of one variable: reference ds_dataset.py:49-75 (np.nanmin / np.nanmax / isnan().sum())"""
        if self._device_ok(values):
            from ..engine import ops
            if name not in self._raw:
                self._raw[name] = torch.from_numpy(np.ascontiguousarray(values)).cuda()
            return ops.minmax(self._raw[name])
        return float(np.nanmin(values)), float(np.nanmax(values)), int(np.isnan(values).sum())

    def device_arrays(self, order=None, device=None, with_mask=False):
        """(X, Y, M) normalised fp32 CUDA tensors in `order` - the device-side equivalent of input_array / output_array /
        mask_array followed by the H2D copy; bit-identical to them (fp32 subtract + IEEE divide).  None when the device
        path does not apply (the caller then uses the numpy methods)."""
        names = list(self.input_variable_names) + ([self.output_variable_name] if self.output_da is not None else [])
        if not torch.cuda.is_available() or any(n not in self._raw for n in names):
            return None
        from ..engine import ops
        dev = self._raw[names[0]].device
        n_out = self.n if order is None else len(order)
        idx = None if order is None else torch.as_tensor(np.asarray(order, dtype=np.int32), device=dev)
        X = torch.empty(n_out, self.input_chan, self.input_y, self.input_x, dtype=torch.float32, device=dev)
        c = 0
        for name, da in zip(self.input_variable_names, self.input_das):
            ops.normalise_gather(self._raw[name], idx, self.min_inputs[name], self.max_inputs[name], self.normalise_in, X, c)
            c += da.shape[1]
        Y = None
        if self.output_da is not None:
            Y = torch.empty(n_out, self.output_chan, self.output_y, self.output_x, dtype=torch.float32, device=dev)
            ops.normalise_gather(self._raw[self.output_variable_name], idx, self.min_output, self.max_output,
                                 self.normalise_out, Y, 0)
        M = torch.from_numpy(self.mask_array(order)).to(dev) if with_mask else None
        return X, Y, M

    def release_device(self):
        self._raw = {}

    # ---- configuration
    def set_normalise_output(self, normalise_out):
        self.normalise_out = normalise_out

    def get_normalisation_parameters(self):
        return [self.min_inputs, self.max_inputs, self.min_output, self.max_output]

    def set_normalisation_parameters(self, parameters):
        (self.min_inputs, self.max_inputs, self.min_output, self.max_output) = tuple(parameters)

    def get_input_shape(self):
        return (self.input_chan, self.input_y, self.input_x)

    def get_input_spec(self):
        return self.input_spec

    def get_output_shape(self):
        return (self.output_chan, self.output_y, self.output_x)

    def get_output_spec(self):
        return self.output_spec

    # ---- normalisation (min-max to [0,1])
    def normalise_input(self, arr, input_name):
        if not self.normalise_in:
            return arr
        lo, hi = self.min_inputs[input_name], self.max_inputs[input_name]
        if hi - lo == 0:
            return 0.0
        return (arr - lo) / (hi - lo)

    def normalise_output(self, arr):
        if not self.normalise_out:
            return arr
        return (arr - self.min_output) / (self.max_output - self.min_output)

    def denormalise_output(self, arr, force=False):
        if force or self.normalise_out:
            return self.min_output + (arr * (self.max_output - self.min_output))
        return arr

    # ---- whole-array assembly (vectorised equivalent of iterating __getitem__ + default collate)
    def input_array(self, order=None):
        out = np.zeros((self.n if order is None else len(order), self.input_chan, self.input_y, self.input_x),
                       dtype=np.float32)
        c = 0
        for name, da in zip(self.input_variable_names, self.input_das):
            vals = np.asarray(da.data)
            if order is not None:
                vals = vals[order]
            nchan = da.shape[1]
            out[:, c:c + nchan] = self.normalise_input(vals, name)
            c += nchan
        return out

    def output_array(self, order=None):
        vals = np.asarray(self.output_da.values)
        if order is not None:
            vals = vals[order]
        return np.asarray(self.normalise_output(vals), dtype=np.float32)

    def mask_array(self, order=None, like_output=True):
        if self.mask_da is not None and self.mask_da.size > 0:
            vals = np.asarray(self.mask_da.values)
            if order is not None:
                vals = vals[order]
            return vals.astype(np.float32)
        n = self.n if order is None else len(order)
        if like_output and self.output_da is not None:
            return np.ones((n, self.output_chan, self.output_y, self.output_x), dtype=np.float32)
        return np.ones((n, self.input_chan, self.input_y, self.input_x), dtype=np.float32)

    # ---- per-item protocol (kept for API compatibility)
    def __getitem__(self, index):
        in_arr = self.input_array(order=[index])[0]
        out_arr = self.output_array(order=[index])[0] if self.output_da is not None else None
        mask = self.mask_array(order=[index], like_output=False)[0]
        return (in_arr, out_arr, mask, f"image{index}")

    def __len__(self):
        return self.n
