"""LinearModel: the reference's `--method linear` (reference: src/cae_tools/models/linear_model.py:30-290 - constructor
arguments, train / apply / save / load, model-folder files `weights`, `normalisation.weights`, `parameters.json` with
"type": "LinearModel", `history.json`, `summary.txt`).  The arithmetic runs in libcae_b200 through engine/linear.py."""

from __future__ import annotations

import json
import os
import time

import numpy as np
import torch

from .base_model import BaseModel
from .conv_ae_model import ConvAEModel, shuffled_order
from .ds_dataset import DSDataset
from .linear import Linear


class LinearModel(ConvAEModel):

    MODEL_TYPE = "LinearModel"
    DB_TYPE = "Linear"

    def __init__(self, normalise_input=True, normalise_output=True, batch_size=10, nr_epochs=500, test_interval=10,
                 lr=0.001, weight_decay=1e-5, use_gpu=True, database_path=None):
        super().__init__(normalise_input=normalise_input, normalise_output=normalise_output, batch_size=batch_size,
                         nr_epochs=nr_epochs, test_interval=test_interval, lr=lr, weight_decay=weight_decay,
                         use_gpu=use_gpu, database_path=database_path)
        self.weights = None

    def get_parameters(self):
        return {
            "model_id": self.get_model_id(),
            "type": "LinearModel",
            "input_shape": list(self.input_shape),
            "output_shape": list(self.output_shape),
            "batch_size": self.batch_size,
            "test_interval": self.test_interval,
            "lr": self.lr,
            "weight_decay": self.weight_decay,
            "normalise_input": self.normalise_input,
            "normalise_output": self.normalise_output,
        }

    def save(self, to_folder):
        os.makedirs(to_folder, exist_ok=True)
        torch.save(self._cpu_state(self.weights), os.path.join(to_folder, "weights"))
        payload = {
            "normalisation.weights": json.dumps(self.normalisation_parameters),
            "parameters.json": json.dumps(self.get_parameters()),
            "history.json": json.dumps(self.history),
            "summary.txt": self.summary(),
        }
        for name, text in payload.items():
            with open(os.path.join(to_folder, name), "w") as f:
                f.write(text)
        BaseModel.save(self, to_folder)

    def load(self, from_folder):
        with open(os.path.join(from_folder, "normalisation.weights"), "r") as f:
            self.normalisation_parameters = json.loads(f.read())
        with open(os.path.join(from_folder, "parameters.json")) as f:
            parameters = json.loads(f.read())
        if "model_id" in parameters:
            self.set_model_id(parameters["model_id"])
        self.input_shape = tuple(parameters["input_shape"])
        self.output_shape = tuple(parameters["output_shape"])
        for key in ("batch_size", "test_interval", "lr", "weight_decay", "normalise_input", "normalise_output"):
            setattr(self, key, parameters[key])
        with open(os.path.join(from_folder, "history.json")) as f:
            self.history = json.loads(f.read())
        self.weights = Linear(self.input_shape, self.output_shape)
        self.weights.load_state_dict(self.torch_load(os.path.join(from_folder, "weights")))
        self.weights.eval()
        self.engine = None
        BaseModel.load(self, from_folder)

    def _make_engine(self, device, dp=None):
        from ..engine.linear import LinearEngine
        if dp is not None:
            return LinearEngine(self.weights, lr=self.lr, weight_decay=self.weight_decay, device=device,
                                grad_hook=dp.allreduce_grads, count_scale=1.0 / dp.world)
        return LinearEngine(self.weights, lr=self.lr, weight_decay=self.weight_decay, device=device)

    def train(self, input_variables, output_variable, training_ds, testing_ds, model_path="", training_paths="",
              testing_paths="", mask_variable_name=None):
        train_ds = DSDataset(training_ds, input_variables, output_variable,
                             normalise_in=self.normalise_input, normalise_out=self.normalise_output)
        self.set_input_spec(train_ds.get_input_spec())
        self.set_output_spec(train_ds.get_output_spec())
        self.normalisation_parameters = train_ds.get_normalisation_parameters()
        test_ds = DSDataset(testing_ds, input_variables, output_variable,
                            normalise_in=self.normalise_input, normalise_out=self.normalise_output)
        test_ds.set_normalisation_parameters(self.normalisation_parameters)
        self.input_shape = tuple(train_ds.get_input_shape())
        self.output_shape = tuple(train_ds.get_output_shape())
        if not self.weights:
            self.weights = Linear(self.input_shape, self.output_shape)
        device = self._device()
        if self.verbose:
            print(f'Running on device: {device}')
        start = time.time()
        train_order = shuffled_order(len(train_ds), self.batch_size)
        test_order = shuffled_order(len(test_ds), self.batch_size)
        # data parallel (torch.distributed initialised with > 1 rank): as ConvAEModel.train - every rank takes its contiguous
        # share of every batch, one SUM all-reduce of the gradients per step, rank 0 saves / evaluates
        from ..engine.dp import DPContext, shard_batches
        dp = DPContext.from_env()
        local_batch = self.batch_size
        if dp is not None:
            train_order, local_batch, _ = shard_batches(train_order, self.batch_size, dp.rank, dp.world)
            test_order, _, _ = shard_batches(test_order, self.batch_size, dp.rank, dp.world)
        self.engine = eng = self._make_engine(device, dp)
        if dp is not None:
            dp.broadcast_([eng.arena])
        gather = (lambda t: dp.reduce_losses(t)) if dp is not None else (lambda t: t)
        train_data = eng.bind(torch.from_numpy(train_ds.input_array(train_order)),
                              torch.from_numpy(train_ds.output_array(train_order)), local_batch)
        test_data = eng.bind(torch.from_numpy(test_ds.input_array(test_order)),
                             torch.from_numpy(test_ds.output_array(test_order)), local_batch)
        train_loss = test_loss = 0.0
        last = self.nr_epochs - 1
        for epoch in range(self.nr_epochs):
            losses = eng.train_epoch(train_data)
            report = (epoch % self.test_interval == 0)
            if report or epoch == last:
                train_loss = float(np.mean(gather(losses).cpu().numpy()))
            if report:
                test_loss = float(np.mean(gather(eng.test_epoch(test_data)).cpu().numpy()))
                self.history["train_loss"].append(train_loss)
                self.history["test_loss"].append(test_loss)
                if self.verbose:
                    print("%5d %.6f %.6f" % (epoch, train_loss, test_loss))
        torch.cuda.synchronize()
        self.history['nr_epochs'] = self.history['nr_epochs'] + self.nr_epochs
        if self.verbose:
            print("elapsed:" + str(time.time() - start))
        self.weights.eval()
        lead = dp is None or dp.rank == 0
        if self.db and lead:
            self.db.add_training_result(self.get_model_id(), self.DB_TYPE, output_variable, input_variables,
                                        self.summary(), model_path, training_paths, train_loss, testing_paths,
                                        test_loss, self.get_parameters(), {})
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
        if self.input_shape:
            s = "Model Summary:\n"
            s += "\tInput shape:\n"
            s += f"\t\tsize={self.input_shape}\n"
            s += "\tOutput shape:\n"
            s += f"\t\tsize={self.output_shape}\n"
            return s
        return "Model has not been trained"
