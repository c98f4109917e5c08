"""apply_cae: apply a saved model to NetCDF data (reference: src/cae_tools/cli/apply_cae.py:28-90, same arguments)."""

import argparse
import json
import os

try:  # pragma: no cover
    import xarray as xr
except ImportError:
    from ..utils import xr_lite as xr

from .train_cae import expand_scalar_inputs, model_class_for_type


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("data_paths", nargs="+",
                        help="path to netcdf4 file(s) containing data to which model is applied")
    parser.add_argument("output_path", help="path to write the netcdf4 file containing input data plus model outputs")
    parser.add_argument("--model-folder", help="folder to save the trained model to", required=True)
    parser.add_argument("--input-variables", nargs="+", help="name of the input variable(s) in training/test data",
                        required=False)
    parser.add_argument("--prediction-variable", help="name of the prediction variable to create in output data",
                        default="model_output")
    parser.add_argument("--mask-variable", type=str, help="name of the mask variable", default=None)
    # not in the reference: the cases are sharded over N GPUs of this node (one process per GPU, no collective on the
    # compute path); rank 0 writes the output file
    parser.add_argument("--gpus", type=int, default=1, help="shard the cases over N GPUs (ignored under torchrun)")
    args = parser.parse_args(argv)
    import sys
    from ..engine import dp as _dp
    if args.gpus > 1 and "RANK" not in os.environ:
        _dp.respawn_under_torchrun(args.gpus, sys.argv[1:] if argv is None else argv, "cae_tools_b200.cli.apply_cae")
    ctx = _dp.init_from_env()
    lead = ctx is None or ctx.rank == 0

    with open(os.path.join(args.model_folder, "parameters.json")) as f:
        parameters = json.loads(f.read())
    mt = model_class_for_type(parameters["type"])()
    mt.load(args.model_folder)

    model_names = mt.get_input_variable_names()
    names = args.input_variables
    if not names:
        if model_names is None:
            raise Exception("Please specify the input variable names using --input-variables")
        names = model_names
    elif model_names is not None and names != model_names:
        raise Exception(f"input_variables [{','.join(names)}] inconsistent with those used to train the model "
                        f"[{','.join(model_names)}]")

    score_ds = xr.open_mfdataset(args.data_paths, concat_dim="box", combine="nested")
    case_dimension = score_ds[names[0]].dims[0]
    expand_scalar_inputs(score_ds, names, case_dimension)
    if lead:
        print("Applying model for %d cases" % score_ds[names[0]].shape[0])
    mt.apply(score_ds, names, args.prediction_variable, mask_variable_name=args.mask_variable)
    if lead:
        score_ds.to_netcdf(args.output_path)
    if ctx is not None:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
