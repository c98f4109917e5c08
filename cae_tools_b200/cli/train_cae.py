"""train_cae: command-line training, same flags as the reference (reference: src/cae_tools/cli/train_cae.py:15-158).

Differences, all deliberate: `--stride`, `--kernel-size`, `--input-layer-count`, `--output-layer-count`,
`--weight-decay` and `--database-path` are forwarded to ConvAEModel / VarAEModel as their help text promises
(the reference parses but drops them for `--method conv`); `--method var|vae` constructs a VarAEModel (absent
from the reference snapshot).  NetCDF files are opened with xarray when it is installed, otherwise with the
NetCDF-3 reader in utils/xr_lite.py.
"""

import argparse
import json
import os
import time

import numpy as np

try:  # pragma: no cover
    import xarray as xr
except ImportError:
    from ..utils import xr_lite as xr

from ..models.model_sizer import ModelSpec

METHODS = ["conv", "unet", "unet_res", "srcnn_res", "resunet_gan", "var", "vae", "linear"]


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--train-inputs", nargs="+", help="path(s) to netcdf4 file containing training data", required=True)
    p.add_argument("--test-inputs", nargs="+", help="path(s) to netcdf4 file containing test data", required=True)
    p.add_argument("--model-folder", help="folder to save the trained model to", required=True)
    p.add_argument("--continue-training", action="store_true", help="continue training model")
    p.add_argument("--input-variables", nargs="+", help="name of the input variable(s) in training/test data",
                   required=True)
    p.add_argument("--output-variable", help="name of the output variable in training/test data", required=True)
    p.add_argument("--nr-epochs", type=int, help="number of training epochs", default=500)
    p.add_argument("--latent-size", type=int, help="size of the latent space", default=4)
    p.add_argument("--fc-size", type=int, help="size of the fully-connected layers", default=16)
    p.add_argument("--batch-size", type=int, help="number of images to process in one batch", default=10)
    p.add_argument("--learning-rate", type=float, help="the learning rate", default=0.001)
    p.add_argument("--lr-step-size", type=int, help="the schedular steps for the learning rate", default=500)
    p.add_argument("--lr-gamma", type=float, help="decay factor of the scheduled learning rate", default=0.5)
    p.add_argument("--lambda-mse", type=float, help="controls the strength of the mse loss in vae", default=1)
    p.add_argument("--lambda-kl", type=float, help="controls the strength of the kl loss in vae", default=1)
    p.add_argument("--lambda-l1", type=float, help="controls the strength of l1 regularization", default=0.001)
    p.add_argument("--lambda-pearson", type=float, help="controls the strength of the pearson loss", default=1)
    p.add_argument("--lambda-ssim", type=float, help="controls the strength of the ssim loss", default=1)
    p.add_argument("--lambda-additional", type=float, help="controls the strength of additional regularization",
                   default=1)
    p.add_argument("--weight-decay", type=float, help="weight decay coefficient", default=1e-5)
    p.add_argument("--dropout-rate", type=float, help="dropout rate", default=1e-1)
    p.add_argument("--additional-loss", type=str, help="additional loss types", default=None)
    p.add_argument("--scheduler-type", type=str, help="scheduler type", default=None)
    p.add_argument("--method", choices=METHODS, default="var", help="methods")
    p.add_argument("--layer-definitions-path", help="specify path of a JSON file with layer definitions", default=None)
    p.add_argument("--stride", type=int, help="stride to use in convolutional layers", default=2)
    p.add_argument("--kernel-size", type=int, help="kernel size to use in convolutional layers", default=3)
    p.add_argument("--input-layer-count", type=int, help="number of input convolutional layers", default=None)
    p.add_argument("--output-layer-count", type=int, help="number of output convolutional layers", default=None)
    p.add_argument("--model-id", type=str, help="specify the model id when creating a model", default=None)
    p.add_argument("--database-path", type=str, help="path to a database to store evaluation results", default=None)
    p.add_argument("--chunk-size", type=int, help="chunk size for xarray", default=1000)
    p.add_argument("--include-coasts", help="include coastal areas", default=False)
    p.add_argument("--mask-variable", type=str, help="name of the mask variable", default=None)
    # not in the reference (it has no multi-GPU path): batch-sharded data parallelism, one process per GPU
    p.add_argument("--gpus", type=int, default=1, help="data-parallel training on N GPUs of this node (one process per "
                   "GPU, gradient all-reduce; BatchNorm uses each rank's local statistics); under torchrun the environment "
                   "decides and this flag is ignored")
    return p


def model_class_for_type(type_name):
    """model class from the "type" field of parameters.json (reference: cli/train_cae.py:116-121, cli/apply_cae.py:46-51)"""
    if type_name == "ConvAEModel":
        from ..models.conv_ae_model import ConvAEModel
        return ConvAEModel
    if type_name == "VarAEModel":
        from ..models.var_ae_model import VarAEModel
        return VarAEModel
    if type_name == "UNET":
        from ..models.unet import UNET
        return UNET
    if type_name == "LinearModel":
        from ..models.linear_model import LinearModel
        return LinearModel
    raise ValueError(f"unknown model type {type_name!r} in parameters.json")


def expand_scalar_inputs(ds, names, case_dimension):
    """1-D per-case variables are broadcast to (case, 1, y, x) (reference: cli/train_cae.py:73-109)"""
    ref = None
    for var in names:
        if len(ds[var].dims) == 4:
            ref = ds[var]
            break
    for var in names:
        if tuple(ds[var].dims) == (case_dimension,):
            if ref is None:
                raise ValueError("cannot expand 1-D input variables without a 4-D input to take the grid size from")
            vals = np.asarray(ds[var].values)
            full = np.broadcast_to(vals[:, None, None, None], (vals.shape[0], 1, ref.shape[2], ref.shape[3])).copy()
            ds[var] = xr.DataArray(full, dims=(case_dimension, "channel", ref.dims[2], ref.dims[3]))


def main(argv=None):
    import sys
    args = build_parser().parse_args(argv)
    from ..engine import dp as _dp
    if args.gpus > 1 and "RANK" not in os.environ:
        _dp.respawn_under_torchrun(args.gpus, sys.argv[1:] if argv is None else argv, "cae_tools_b200.cli.train_cae")
    ctx = _dp.init_from_env()
    lead = ctx is None or ctx.rank == 0
    train_ds = xr.open_mfdataset(args.train_inputs, concat_dim="box", combine="nested")
    test_ds = xr.open_mfdataset(args.test_inputs, concat_dim="box", combine="nested")
    case_dimension = train_ds[args.output_variable].dims[0]
    if lead:
        print("Training cases: %d, Test cases: %d" % (train_ds[args.output_variable].shape[0],
                                                      test_ds[args.output_variable].shape[0]))
        if ctx is not None:
            print(f"Data parallel over {ctx.world} GPUs (batch {args.batch_size} split {args.batch_size // ctx.world} per GPU)")
    expand_scalar_inputs(train_ds, args.input_variables, case_dimension)
    expand_scalar_inputs(test_ds, args.input_variables, case_dimension)

    if args.continue_training:
        with open(os.path.join(args.model_folder, "parameters.json")) as f:
            parameters = json.loads(f.read())
        mt = model_class_for_type(parameters["type"])()
        mt.load(args.model_folder)
        mt.nr_epochs = args.nr_epochs
        mt.lr = args.learning_rate
        mt.batch_size = args.batch_size
    else:
        common = dict(fc_size=args.fc_size, encoded_dim_size=args.latent_size, nr_epochs=args.nr_epochs,
                      batch_size=args.batch_size, lr=args.learning_rate, weight_decay=args.weight_decay,
                      conv_kernel_size=args.kernel_size, conv_stride=args.stride,
                      conv_input_layer_count=args.input_layer_count, conv_output_layer_count=args.output_layer_count,
                      database_path=args.database_path)
        if args.method == "conv":
            mt = model_class_for_type("ConvAEModel")(**common)
        elif args.method in ("var", "vae"):
            mt = model_class_for_type("VarAEModel")(lambda_mse=args.lambda_mse, lambda_kl=args.lambda_kl, **common)
        elif args.method == "unet":
            mt = model_class_for_type("UNET")(lambda_l1=args.lambda_l1, lambda_pearson=args.lambda_pearson,
                                              dropout_rate=args.dropout_rate, **common)
        elif args.method == "linear":
            mt = model_class_for_type("LinearModel")(batch_size=args.batch_size, nr_epochs=args.nr_epochs,
                                                     lr=args.learning_rate)
        else:
            raise SystemExit(f"--method {args.method} has no implementation (the reference snapshot has none either)")
        if args.model_id:
            mt.set_model_id(args.model_id)
        if args.layer_definitions_path:
            with open(args.layer_definitions_path) as f:
                spec = ModelSpec()
                spec.load(json.loads(f.read()))
                mt.spec = spec

    if ctx is not None and not lead:
        mt.verbose = False
    start_time = time.time()
    if lead:
        print("Ready for training process")
    mt.train(args.input_variables, args.output_variable, training_ds=train_ds, testing_ds=test_ds,
             model_path=args.model_folder, training_paths=";".join(args.train_inputs),
             testing_paths=";".join(args.test_inputs), mask_variable_name=args.mask_variable)
    if lead:
        print(f"Time taken to train: {time.time() - start_time:.2f} seconds")
    if ctx is not None:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
