"""Behaviour shared by the model classes: ids, input/output specs on disk, ``apply`` and ``evaluate``.

API and on-disk files follow the reference ``BaseModel`` (reference: src/cae_tools/models/base_model.py:28-203).
``apply`` keeps the reference contract (normalise -> batches on device -> ``score`` -> denormalise ->
new variable on the data set) but assembles the batches in one vectorised pass and lets the engine
stream predictions back through pinned memory.
"""

from __future__ import annotations

import json
import os
import uuid

import numpy as np
import torch

from .ds_dataset import DSDataset
from .model_metric import ModelMetric

try:  # pragma: no cover - depends on the image
    import xarray as _xr
except ImportError:  # the build image has no xarray
    from ..utils import xr_lite as _xr


class BaseModel:

    def __init__(self):
        self.input_spec = None
        self.output_spec = None
        self.model_id = str(uuid.uuid4())

    # ---- specs / ids
    def set_input_spec(self, input_spec):
        self.input_spec = input_spec

    def get_input_spec(self):
        return self.input_spec

    def set_output_spec(self, output_spec):
        self.output_spec = output_spec

    def get_output_spec(self):
        return self.output_spec

    def get_input_variable_names(self):
        return None if self.input_spec is None else [item["name"] for item in self.input_spec]

    def get_output_variable_name(self):
        return None if self.output_spec is None else self.output_spec["name"]

    def set_model_id(self, model_id):
        self.model_id = model_id

    def get_model_id(self):
        return self.model_id

    def torch_load(self, from_path):
        return torch.load(from_path, map_location=None if torch.cuda.is_available() else torch.device("cpu"))

    # ---- inference over a data set
    def predict_array(self, inputs: np.ndarray) -> np.ndarray:
        """normalised fp32 inputs [n,C,y,x] -> model outputs [n,C',y',x'] (sub-classes implement)"""
        raise NotImplementedError

    def apply(self, score_ds, input_variables, prediction_variable="model_output",
              channel_dimension="model_output_channel", y_dimension="model_output_y", x_dimension="model_output_x",
              mask_variable_name=None):
        """Add the model's (de-normalised) estimate to `score_ds` as `prediction_variable`."""
        n_dimension = score_ds[input_variables[0]].dims[0]
        ds = DSDataset(score_ds, input_variables, input_variables[0], normalise_in=self.normalise_input,
                       mask_variable_name=mask_variable_name)
        ds.set_normalisation_parameters(self.normalisation_parameters)
        inputs = ds.input_array()
        n = inputs.shape[0]
        # one process per GPU (torch.distributed initialised): every rank predicts its contiguous share of the cases -
        # the compute has no collective; the shares are then exchanged so that every rank's data set carries the whole
        # variable, as in the single-process call.  apply_shard() below skips the exchange (sweeps too large to gather).
        from ..engine.dp import DPContext, shard_bounds
        dp = DPContext.from_env()
        if dp is None:
            scores = self.predict_array(inputs)
        else:
            lo, hi, mine = self.predict_shard(inputs, dp)
            scores = np.empty((n,) + mine.shape[1:], dtype=np.float32)
            scores[lo:hi] = mine
            import torch.distributed as dist
            for r in range(dp.world):
                rlo, rhi = shard_bounds(n, r, dp.world)
                if rhi > rlo:
                    t = torch.from_numpy(scores[rlo:rhi])
                    if dist.get_backend(dp.group) == "nccl":
                        t = t.cuda()
                    dist.broadcast(t, src=r, group=dp.group)
                    if r != dp.rank:
                        scores[rlo:rhi] = t.cpu().numpy()
        out = ds.denormalise_output(scores.astype(np.float64))
        score_ds[prediction_variable] = _xr.DataArray(out, dims=(n_dimension, channel_dimension, y_dimension,
                                                                  x_dimension))

    def apply_shard(self, score_ds, input_variables, mask_variable_name=None):
        """Sharded apply() for sweeps that must not be gathered (BASELINE configs[4]: N = 1 M cases over 8 GPUs): returns
        (lo, hi, de-normalised float64 estimates of cases [lo, hi)) of THIS rank; no collective, no full-N allocation
        (the reference allocates all N outputs as float64 on the host, base_model.py:123)."""
        from ..engine.dp import DPContext, shard_bounds
        dp = DPContext.from_env()
        n = score_ds[input_variables[0]].shape[0]
        lo, hi = (0, n) if dp is None else shard_bounds(n, dp.rank, dp.world)
        ds = DSDataset(score_ds, input_variables, input_variables[0], normalise_in=self.normalise_input,
                       mask_variable_name=mask_variable_name)
        ds.set_normalisation_parameters(self.normalisation_parameters)
        scores = self.predict_array(ds.input_array(list(range(lo, hi))))
        return lo, hi, ds.denormalise_output(scores.astype(np.float64))

    def evaluate(self, dataset, device=None):
        """mse / rmse / mae / mean Pearson of de-normalised predictions against the data set's output (reference:
        base_model.py:116-125 + model_metric.py).  With an engine on a CUDA device the per-case sums are formed on the device
        next to the predictions (`cae_case_metrics`, float64) and only eight numbers per case come back; the host numpy path
        below is what runs otherwise and what the device path is tested against."""
        dataset.set_normalise_output(False)
        got = self._evaluate_device(dataset)
        if got is not None:
            return got
        scores = dataset.denormalise_output(self.predict_array(dataset.input_array()).astype(np.float64), force=True)
        actual = np.asarray(dataset.output_da.values)
        mask = dataset.mask_array()
        mm = ModelMetric()
        for i in range(actual.shape[0]):
            mm.accumulate(actual[i], scores[i], mask[i])
        return mm.get_metrics()

    device_metrics = True

    def _evaluate_device(self, dataset):
        if not (self.device_metrics and torch.cuda.is_available() and hasattr(self, "_ensure_engine")):
            return None
        actual = np.asarray(dataset.output_da.values)
        if actual.dtype != np.float32 or actual.ndim != 4:
            return None
        from ..engine import ops
        eng = self._ensure_engine()
        dev = eng.device
        n = actual.shape[0]
        bs = max(1, min(getattr(self, "apply_batch_size", n), n))
        data = eng.bind(torch.from_numpy(np.ascontiguousarray(dataset.input_array(), dtype=np.float32)), None, bs)
        raw = getattr(dataset, "_raw", {}).get(dataset.output_variable_name)          # uploaded by the ingest scan
        act = raw if raw is not None and raw.device == dev else torch.from_numpy(np.ascontiguousarray(actual)).to(dev)
        mask = None
        if dataset.mask_da is not None and dataset.mask_da.size > 0:
            mask = torch.from_numpy(np.ascontiguousarray(dataset.mask_array(like_output=False), dtype=np.float32)).to(dev)
        rows = torch.zeros(n, 8, dtype=torch.float64, device=dev)
        lo, hi = float(dataset.min_output), float(dataset.max_output)

        def sink(i, yhat):
            a, b = i * bs, i * bs + yhat.shape[0]
            ops.case_metrics(yhat.contiguous(), act[a:b], None if mask is None else mask[a:b], lo, hi - lo, rows[a:b])

        eng.score_batches(data, sink)
        r = rows.cpu().numpy()
        cnt, sa, se, saa, see, sae, ab, sq = (r[:, k] for k in range(8))
        if cnt.sum() == 0:
            raise ValueError("No data accumulated to calculate metrics.")
        live = cnt > 0
        with np.errstate(invalid="ignore", divide="ignore"):
            cov = sae[live] - sa[live] * se[live] / cnt[live]
            den = np.sqrt((saa[live] - sa[live] ** 2 / cnt[live]) * (see[live] - se[live] ** 2 / cnt[live]))
            corr = np.where(den > 0, cov / den, np.nan)
        mse = float(sq.sum() / cnt.sum())
        return {"mse": mse, "rmse": float(np.sqrt(mse)), "mae": float(ab.sum() / cnt.sum()),
                "mean_pearson_correlation": float(np.mean(corr)) if corr.size else 0.0}

    def dump_metrics(self, title, metrics):
        print("\n" + title)
        for key in metrics:
            print(f"\t{key:30s}:{metrics[key]}")

    def score(self, batches, save_arr):
        raise NotImplementedError

    # ---- persistence of the input/output specs
    def save(self, to_folder):
        for name, spec in (("input_spec.json", self.input_spec), ("output_spec.json", self.output_spec)):
            if spec is not None:
                with open(os.path.join(to_folder, name), "w") as f:
                    f.write(json.dumps(spec))

    def load(self, from_folder):
        path = os.path.join(from_folder, "input_spec.json")
        if os.path.exists(path):
            with open(path) as f:
                self.input_spec = json.loads(f.read())
        path = os.path.join(from_folder, "output_spec.json")
        if os.path.exists(path):
            with open(path) as f:
                self.output_spec = json.loads(f.read())

    def train(self, input_variables, output_variable, training_ds, testing_ds, model_path="", training_paths="",
              testing_paths="", mask_variable_name=None):
        raise NotImplementedError

    def summary(self):
        raise NotImplementedError

    def get_parameters(self):
        raise NotImplementedError
