"""ConvAEModel: the reference's train / apply / save / load API on the B200 kernels.

Constructor arguments, method signatures, the saved model-folder layout and the bookkeeping of the
training loop follow the reference (reference: src/cae_tools/models/conv_ae_model.py:33-380):

* ``train``: data set -> min/max normalisation -> layer spec -> Encoder, Decoder (constructed in that
  order so the global torch RNG yields the reference's initial weights) -> ONE shuffle of the train and
  test sets (drawn from the same RNG stream the reference's DataLoaders use) -> all batches moved to
  the device once -> ``nr_epochs`` x (train epoch [+ test epoch every ``test_interval``]) -> history,
  save, metrics.
* per-epoch train loss = unweighted mean of the per-batch MSE values; the last ragged batch is kept.

What differs is where the arithmetic runs: every forward/backward/optimiser op is a hand-written
sm_100a kernel driven by ``engine.convae.ConvAEEngine`` (CUDA graph per batch geometry, no host sync
inside an epoch).  There is no CPU path: without a CUDA device (or with ``use_gpu=False``) the
compute entry points raise.
"""

from __future__ import annotations

import json
import os
import time

import numpy as np
import torch

from .base_model import BaseModel
from .ds_dataset import DSDataset
from .decoder import Decoder
from .encoder import Encoder
from .model_sizer import ModelSpec, create_model_spec
from .._lib import CaeError


def shuffled_order(n, batch_size):
    """The sample order a ``DataLoader(ds, batch_size, shuffle=True)`` pass produces, consuming the global
    torch RNG exactly as the reference's loaders do (reference: conv_ae_model.py:291-292,316,322)."""
    loader = torch.utils.data.DataLoader(range(n), batch_size=batch_size, shuffle=True)
    order = []
    for idx in loader:
        order.extend(int(i) for i in idx)
    return order


class ConvAEModel(BaseModel):

    MODEL_TYPE = "ConvAEModel"
    DB_TYPE = "ConvAE"

    def __init__(self, normalise_input=True, normalise_output=True, batch_size=10,
                 nr_epochs=500, test_interval=10, encoded_dim_size=32, fc_size=128,
                 lr=0.001, weight_decay=1e-5, use_gpu=True, conv_kernel_size=3, conv_stride=2,
                 conv_input_layer_count=None, conv_output_layer_count=None, database_path=None):
        super().__init__()
        self.normalise_input = normalise_input
        self.normalise_output = normalise_output
        self.normalisation_parameters = None
        self.input_shape = self.output_shape = None
        self.encoder = self.decoder = None
        self.batch_size = batch_size
        self.nr_epochs = nr_epochs
        self.test_interval = test_interval
        self.encoded_dim_size = encoded_dim_size
        self.fc_size = fc_size
        self.lr = lr
        self.weight_decay = weight_decay
        self.use_gpu = use_gpu
        self.conv_kernel_size = conv_kernel_size
        self.conv_stride = conv_stride
        self.conv_input_layer_count = conv_input_layer_count
        self.conv_output_layer_count = conv_output_layer_count
        self.spec = None
        self.history = {'train_loss': [], 'test_loss': [], 'nr_epochs': 0}
        self.optim = None
        self.db = None
        if database_path:
            from ..utils.model_database import ModelDatabase
            self.db = ModelDatabase(database_path)
        self.engine = None
        self.apply_batch_size = 4096   # eval-mode outputs do not depend on the batch split (kernels tuned at 4096)
        self.verbose = True

    # ------------------------------------------------------------------ parameters / persistence
    def get_parameters(self):
        return {
            "type": self.MODEL_TYPE,
            "input_shape": list(self.input_shape),
            "output_shape": list(self.output_shape),
            "batch_size": self.batch_size,
            "test_interval": self.test_interval,
            "encoded_dim_size": self.encoded_dim_size,
            "fc_size": self.fc_size,
            "lr": self.lr,
            "weight_decay": self.weight_decay,
            "normalise_input": self.normalise_input,
            "normalise_output": self.normalise_output,
            "conv_kernel_size": self.conv_kernel_size,
            "conv_stride": self.conv_stride,
            "conv_input_layer_count": self.conv_input_layer_count,
            "conv_output_layer_count": self.conv_output_layer_count,
            "model_id": self.get_model_id()
        }

    @staticmethod
    def _cpu_state(module):
        return {k: v.detach().to("cpu").clone() for k, v in module.state_dict().items()}

    def save(self, to_folder):
        os.makedirs(to_folder, exist_ok=True)
        torch.save(self._cpu_state(self.encoder), os.path.join(to_folder, "encoder.weights"))
        torch.save(self._cpu_state(self.decoder), os.path.join(to_folder, "decoder.weights"))
        payload = {
            "normalisation.weights": json.dumps(self.normalisation_parameters),
            "parameters.json": json.dumps(self.get_parameters()),
            "spec.json": json.dumps(self.spec.save()),
            "history.json": json.dumps(self.history),
            "summary.txt": self.summary(),
        }
        for name, text in payload.items():
            with open(os.path.join(to_folder, name), "w") as f:
                f.write(text)
        super().save(to_folder)

    def _load_parameters(self, parameters):
        if "model_id" in parameters:
            self.set_model_id(parameters["model_id"])
        self.input_shape = tuple(parameters["input_shape"])
        self.output_shape = tuple(parameters["output_shape"])
        for key in ("batch_size", "test_interval", "encoded_dim_size", "fc_size", "lr", "weight_decay",
                    "normalise_input", "normalise_output"):
            setattr(self, key, parameters[key])
        for key in ("conv_kernel_size", "conv_stride", "conv_input_layer_count", "conv_output_layer_count"):
            setattr(self, key, parameters.get(key, None))

    def _build_modules(self):
        self.encoder = Encoder(self.spec.get_input_layers(), encoded_space_dim=self.encoded_dim_size,
                               fc_size=self.fc_size)
        self.decoder = Decoder(self.spec.get_output_layers(), encoded_space_dim=self.encoded_dim_size,
                               fc_size=self.fc_size)

    def load(self, from_folder):
        with open(os.path.join(from_folder, "normalisation.weights"), "r") as f:
            self.normalisation_parameters = json.loads(f.read())
        with open(os.path.join(from_folder, "parameters.json")) as f:
            self._load_parameters(json.loads(f.read()))
        with open(os.path.join(from_folder, "history.json")) as f:
            self.history = json.loads(f.read())
        with open(os.path.join(from_folder, "spec.json")) as f:
            self.spec = ModelSpec()
            self.spec.load(json.loads(f.read()))
        self._build_modules()
        self.encoder.load_state_dict(self.torch_load(os.path.join(from_folder, "encoder.weights")))
        self.encoder.eval()
        self.decoder.load_state_dict(self.torch_load(os.path.join(from_folder, "decoder.weights")))
        self.decoder.eval()
        self.engine = None
        super().load(from_folder)

    # ------------------------------------------------------------------ engine plumbing
    def _device(self):
        if not self.use_gpu:
            raise CaeError("use_gpu=False: cae_tools_b200 has no CPU execution path (kernels are sm_100a only)")
        if not torch.cuda.is_available():
            raise CaeError("no CUDA device: cae_tools_b200 has no CPU execution path")
        return torch.device("cuda", torch.cuda.current_device())

    def _make_engine(self, device, dp=None):
        from ..engine.convae import ConvAEEngine
        if dp is None:
            return ConvAEEngine(self.encoder, self.decoder, lr=self.lr, weight_decay=self.weight_decay, device=device)
        return ConvAEEngine(self.encoder, self.decoder, lr=self.lr, weight_decay=self.weight_decay, device=device,
                            grad_hook=dp.allreduce_grads, grad_hook_async=dp.allreduce_grads_async,
                            count_scale=1.0 / dp.world, dp=dp)

    def _ensure_engine(self):
        if self.engine is None:
            self.engine = self._make_engine(self._device())
        return self.engine

    # ------------------------------------------------------------------ inference
    def predict_array(self, inputs, out=None):
        """eval-mode forward of `inputs` [n, C, y, x] (normalised fp32) -> [n, C', y', x'].  Predictions stream back
        through two pinned staging buffers on a copy stream: the device -> host copy of batch i overlaps the kernels of
        batch i + 1 and the host-side scatter of batch i - 1 (reference: one blocking .cpu() per batch,
        conv_ae_model.py:238).  `out` may be a preallocated array (any float dtype) that receives the result."""
        eng = self._ensure_engine()
        n = inputs.shape[0]
        bs = max(1, min(self.apply_batch_size, n))
        data = eng.bind(torch.from_numpy(np.ascontiguousarray(inputs, dtype=np.float32)), None, bs)
        if out is None:
            out = np.empty((n,) + tuple(self.output_shape), dtype=np.float32)
        dev = eng.device
        copy = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        stage = [torch.empty((bs,) + tuple(self.output_shape), dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        taken = [torch.cuda.Event(), torch.cuda.Event()]      # the copy of slot k has read the engine's output buffer
        pending = [None, None]                                # (first row, rows) waiting in the staging slot

        def flush(k):
            if pending[k] is not None:
                done[k].synchronize()
                lo, cnt = pending[k]
                out[lo:lo + cnt] = stage[k][:cnt].numpy()
                pending[k] = None

        def sink(i, yhat):
            k = i & 1
            flush(k)                                          # host scatter of batch i - 2 (its copy finished long ago)
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(copy):
                copy.wait_event(ready)
                stage[k][:yhat.shape[0]].copy_(yhat, non_blocking=True)
                done[k].record(copy)
                taken[k].record(copy)
            main.wait_event(taken[k])                         # the next batch overwrites the engine's output buffer
            pending[k] = (i * bs, yhat.shape[0])

        eng.score_batches(data, sink)
        flush(0)
        flush(1)
        return out

    def predict_shard(self, inputs, dp=None):
        """apply() sharding (N split into contiguous ranges over the ranks, no collective): this rank's
        (lo, hi, predictions[lo:hi]).  `dp` defaults to the initialised torch.distributed group (None / world 1: all)."""
        from ..engine.dp import DPContext, shard_bounds
        dp = dp if dp is not None else DPContext.from_env()
        n = inputs.shape[0]
        lo, hi = (0, n) if dp is None else shard_bounds(n, dp.rank, dp.world)
        return lo, hi, self.predict_array(inputs[lo:hi])

    def score(self, batches, save_arr):
        """reference signature (conv_ae_model.py:223-239): `batches` is a list of input tensors, predictions are
        written to save_arr[ctr:ctr+batch_size] with ctr advancing by self.batch_size per batch"""
        ctr = 0
        for input_data in batches:
            pred = self.predict_array(input_data.detach().cpu().numpy())
            save_arr[ctr:ctr + self.batch_size, :, :, :] = pred
            ctr += self.batch_size

    # ------------------------------------------------------------------ training
    def train(self, input_variables, output_variable, training_ds, testing_ds, model_path="", training_paths="",
              testing_paths="", mask_variable_name=None):
        train_ds = DSDataset(training_ds, input_variables, output_variable,
                             normalise_in=self.normalise_input, normalise_out=self.normalise_output)
        self.normalisation_parameters = train_ds.get_normalisation_parameters()
        self.set_input_spec(train_ds.get_input_spec())
        self.set_output_spec(train_ds.get_output_spec())
        test_ds = DSDataset(testing_ds, input_variables, output_variable,
                            normalise_in=self.normalise_input, normalise_out=self.normalise_output)
        test_ds.set_normalisation_parameters(self.normalisation_parameters)
        self.input_shape = tuple(train_ds.get_input_shape())
        self.output_shape = tuple(train_ds.get_output_shape())

        if not self.spec:
            (ic, iy, ix), (oc, oy, ox) = self.input_shape, self.output_shape
            self.spec = create_model_spec(input_size=(iy, ix), input_channels=ic, output_size=(oy, ox),
                                          output_channels=oc, kernel_size=self.conv_kernel_size,
                                          stride=self.conv_stride, input_layer_count=self.conv_input_layer_count,
                                          output_layer_count=self.conv_output_layer_count)
        if not self.encoder or not self.decoder:
            self._build_modules()

        device = self._device()
        if self.verbose:
            print(f'Running on device: {device}')
        start = time.time()

        # one shuffle per data set, frozen for all epochs (train first, then test: RNG order matters)
        train_order = shuffled_order(len(train_ds), self.batch_size)
        test_order = shuffled_order(len(test_ds), self.batch_size)

        # data parallel (torch.distributed initialised with >1 rank): every rank holds the same weights (same
        # seed => same init; rank 0's are broadcast to be safe) and its contiguous share of every batch
        from ..engine.dp import DPContext, shard_batches
        dp = DPContext.from_env()
        local_batch = self.batch_size
        if dp is not None:
            train_order, local_batch, _ = shard_batches(train_order, self.batch_size, dp.rank, dp.world)
            test_order, _, _ = shard_batches(test_order, self.batch_size, dp.rank, dp.world)
        self.engine = eng = self._make_engine(device, dp)
        if dp is not None:
            dp.broadcast_([eng.arena] + [b for m in (self.encoder, self.decoder) for b in m.buffers()])
        def bind(ds, order):
            dev_arrays = ds.device_arrays(order)        # normalise + shuffle + batch assembly on the device (csrc/ingest.cu)
            if dev_arrays is not None:
                ds.release_device()
                return eng.bind(dev_arrays[0], dev_arrays[1], local_batch)
            return eng.bind(torch.from_numpy(ds.input_array(order)), torch.from_numpy(ds.output_array(order)), local_batch)

        train_data, test_data = bind(train_ds, train_order), bind(test_ds, test_order)
        gather = (lambda t: dp.reduce_losses(t)) if dp is not None else (lambda t: t)

        train_loss = test_loss = 0.0
        last = self.nr_epochs - 1
        for epoch in range(self.nr_epochs):
            losses = eng.train_epoch(train_data)
            report = (epoch % self.test_interval == 0)
            if report or epoch == last:
                train_loss = float(np.mean(gather(losses).cpu().numpy()))   # the only host sync of the epoch
            if report:
                test_loss = float(np.mean(gather(eng.test_epoch(test_data)).cpu().numpy()))
                self.history["train_loss"].append(train_loss)
                self.history["test_loss"].append(test_loss)
                if self.verbose:
                    print("%5d %.6f %.6f" % (epoch, train_loss, test_loss))
        torch.cuda.synchronize()
        elapsed = time.time() - start
        self.history['nr_epochs'] = self.history['nr_epochs'] + self.nr_epochs
        if self.verbose:
            print("elapsed:" + str(elapsed))
        self.encoder.eval()
        self.decoder.eval()

        # data parallel: weights are identical on every rank; rank 0 alone writes the model folder / tracking DB and runs
        # the post-training evaluation (no duplicate sqlite rows, no concurrent writers)
        lead = dp is None or dp.rank == 0
        if self.db and lead:
            self.db.add_training_result(self.get_model_id(), self.DB_TYPE, output_variable, input_variables,
                                        self.summary(), model_path, training_paths, train_loss, testing_paths,
                                        test_loss, self.get_parameters(), self.spec.save())
        if model_path and lead:
            self.save(model_path)
        if not lead:
            return

        metrics = {"test": self.evaluate(test_ds, device), "train": self.evaluate(train_ds, device)}
        if self.verbose:
            self.dump_metrics("Test Metrics", metrics["test"])
            self.dump_metrics("Train Metrics", metrics["train"])
        if self.db:
            self.db.add_evaluation_result(self.get_model_id(), training_paths, testing_paths, metrics)

    def summary(self):
        if not self.spec:
            return "Model has not been trained - no layers assigned yet"
        s = "Model Summary:\n"
        s += "".join(str(layer) for layer in self.spec.input_layers)
        s += f"\tFully Connected Layer:\n\t\tsize={self.fc_size}\n"
        s += f"\tLatent Vector:\n\t\tsize={self.encoded_dim_size}\n"
        s += f"\tFully Connected Layer:\n\t\tsize={self.fc_size}\n"
        s += "".join(str(layer) for layer in self.spec.output_layers)
        return s
