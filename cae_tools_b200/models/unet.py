"""UNET: encoder-decoder with skip connections and channel-attention gates, the reference's `--method unet`
(reference: src/cae_tools/models/unet.py:200-633 - constructor arguments, train / apply / save / load, parameters.json
"type": "UNET", AdamW, loss = masked MSE + lambda_pearson * (1 - mean Pearson); history records the masked MSE).

Differences that are deliberate:
* the reference's constructor builds a VGG perceptual loss (network download, never used in training) and a BCE
  loss (unused): not built here; WGAN-GP / TV-loss helpers are dead code in the reference and are left out;
* the reference's random flip / rotation / crop transforms are assigned to an attribute nothing reads
  (unet.py:430-442): no augmentation happens there, none here;
* the cosine schedule with eta_min == lr is the identity (unet.py:458-459): constant learning rate;
* with no mask variable the reference builds a mask shaped like the INPUT, which cannot broadcast against the
  output unless both sizes agree; here the default mask is all ones of the OUTPUT shape;
* dropout (the reference default is 0.1): masks come from a counter-based hash inside the fused training stem
  (csrc/unet_stem_train.cu), not from torch's generator, so a run is reproducible from (seed, step) but not
  bit-comparable with the reference's random stream; at dropout_rate = 0 the two paths are comparable to 1e-4.
The layer spec has to be symmetric (decoder layer j's output = encoder skip j) and is normally supplied through
`--layer-definitions-path`, exactly as for the reference.
"""

from __future__ import annotations

import time

import numpy as np
import torch

from .conv_ae_model import ConvAEModel, shuffled_order
from .ds_dataset import DSDataset
from .model_sizer import create_model_spec
from .unet_modules import UNetDecoder, UNetEncoder


class UNET(ConvAEModel):

    MODEL_TYPE = "UNET"
    DB_TYPE = "UNET"

    def __init__(self, normalise_input=True, normalise_output=True, batch_size=10,
                 nr_epochs=500, test_interval=10, encoded_dim_size=32, fc_size=128,
                 lr=0.001, weight_decay=1e-5, dropout_rate=0.1, use_gpu=True, conv_kernel_size=3, conv_stride=2,
                 conv_input_layer_count=None, conv_output_layer_count=None, database_path=None, lambda_l1=0.001,
                 lambda_pearson=1):
        super().__init__(normalise_input=normalise_input, normalise_output=normalise_output, batch_size=batch_size,
                         nr_epochs=nr_epochs, test_interval=test_interval, encoded_dim_size=encoded_dim_size,
                         fc_size=fc_size, lr=lr, weight_decay=weight_decay, use_gpu=use_gpu,
                         conv_kernel_size=conv_kernel_size, conv_stride=conv_stride,
                         conv_input_layer_count=conv_input_layer_count,
                         conv_output_layer_count=conv_output_layer_count, database_path=database_path)
        self.dropout_rate = dropout_rate
        self.lambda_l1 = lambda_l1
        self.lambda_pearson = lambda_pearson
        self.history_pearson = {"train": [], "test": []}

    def get_parameters(self):
        p = super().get_parameters()
        p["lambda_pearson"] = self.lambda_pearson
        p["dropout_rate"] = self.dropout_rate
        return p

    def _load_parameters(self, parameters):
        super()._load_parameters(parameters)
        self.lambda_pearson = parameters.get("lambda_pearson", self.lambda_pearson)
        self.dropout_rate = parameters.get("dropout_rate", self.dropout_rate)

    def _build_modules(self):
        self.encoder = UNetEncoder(self.spec.get_input_layers(), encoded_space_dim=self.encoded_dim_size,
                                   fc_size=self.fc_size, dropout_rate=self.dropout_rate)
        self.decoder = UNetDecoder(self.spec.get_output_layers(), encoded_space_dim=self.encoded_dim_size,
                                   fc_size=self.fc_size, dropout_rate=self.dropout_rate)

    def _make_engine(self, device, dp=None):
        from ..engine.unet import UNetEngine
        kw = dict(lr=self.lr, weight_decay=self.weight_decay, device=device)
        seed = torch.initial_seed()
        if dp is not None:
            kw.update(grad_hook=dp.allreduce_grads, grad_hook_async=dp.allreduce_grads_async,
                      count_scale=1.0 / dp.world, dp=dp)
            seed ^= (dp.rank + 1) * 0x9E3779B97F4A7C15           # independent dropout masks on every rank's shard
        return UNetEngine(self.encoder, self.decoder, lambda_pearson=self.lambda_pearson,
                          dropout_rate=self.dropout_rate, seed=seed & 0xFFFFFFFFFFFFFFFF, **kw)

    def train(self, input_variables, output_variable, training_ds, testing_ds, model_path="", training_paths="",
              testing_paths="", mask_variable_name=None):
        train_ds = DSDataset(training_ds, input_variables, output_variable, normalise_in=self.normalise_input,
                             normalise_out=self.normalise_output, mask_variable_name=mask_variable_name)
        self.set_input_spec(train_ds.get_input_spec())
        self.set_output_spec(train_ds.get_output_spec())
        self.normalisation_parameters = train_ds.get_normalisation_parameters()
        test_ds = DSDataset(testing_ds, input_variables, output_variable, normalise_in=self.normalise_input,
                            normalise_out=self.normalise_output, mask_variable_name=mask_variable_name)
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
        train_order = shuffled_order(len(train_ds), self.batch_size)
        test_order = shuffled_order(len(test_ds), self.batch_size)
        from ..engine.dp import DPContext, shard_batches
        dp = DPContext.from_env()
        local_batch = self.batch_size
        if dp is not None:
            train_order, local_batch, _ = shard_batches(train_order, self.batch_size, dp.rank, dp.world)
            test_order, _, _ = shard_batches(test_order, self.batch_size, dp.rank, dp.world)
        self.engine = eng = self._make_engine(device, dp)
        if dp is not None:
            dp.broadcast_([eng.arena] + [b for m in (self.encoder, self.decoder) for b in m.buffers()])
        has_mask = mask_variable_name is not None

        def bind(ds, order):
            dev_arrays = ds.device_arrays(order, with_mask=has_mask)    # device ingest (csrc/ingest.cu) when possible
            if dev_arrays is not None:
                ds.release_device()
                return eng.bind(dev_arrays[0], dev_arrays[1], local_batch, mask=dev_arrays[2])
            mask = torch.from_numpy(ds.mask_array(order)) if has_mask else None
            return eng.bind(torch.from_numpy(ds.input_array(order)), torch.from_numpy(ds.output_array(order)),
                            local_batch, mask=mask)

        train_data, test_data = bind(train_ds, train_order), bind(test_ds, test_order)
        gather = (lambda t: dp.reduce_losses(t)) if dp is not None else (lambda t: t)
        train_loss = test_loss = 0.0
        last = self.nr_epochs - 1
        for epoch in range(self.nr_epochs):
            losses = eng.train_epoch(train_data)
            report = (epoch % self.test_interval == 0)
            if report or epoch == last:
                train_loss = float(np.mean(gather(losses).cpu().numpy()))
                train_pearson = float(np.mean(gather(train_data.pearson).cpu().numpy()))
            if report:
                test_loss = float(np.mean(gather(eng.test_epoch(test_data)).cpu().numpy()))
                test_pearson = float(np.mean(gather(test_data.pearson).cpu().numpy()))
                self.history["train_loss"].append(train_loss)
                self.history["test_loss"].append(test_loss)
                self.history_pearson["train"].append(train_pearson)
                self.history_pearson["test"].append(test_pearson)
                if self.verbose:
                    print(f"epoch: {epoch}, train_mse: {train_loss:.6f}, train_pearson_loss: {train_pearson:.4f}, "
                          f"test_mse: {test_loss:.6f}, test_pearson_loss: {test_pearson:.4f}")
                    print(f"learn rate: {self.lr:.6f}")
        torch.cuda.synchronize()
        elapsed = time.time() - start
        self.history['nr_epochs'] += self.nr_epochs
        if self.verbose:
            print("elapsed:" + str(elapsed))
        self.encoder.eval()
        self.decoder.eval()
        lead = dp is None or dp.rank == 0      # rank 0 alone writes the folder / tracking DB and evaluates
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
