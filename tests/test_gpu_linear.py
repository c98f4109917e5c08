"""`--method linear` on the CUDA path against the oracle port of the reference's Linear / MSELoss / Adam loop
(reference: linear.py:33-49, linear_model.py:142-184,236-242)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def test_linear_engine_vs_oracle():
    from cae_tools_b200.engine.linear import LinearEngine
    from cae_tools_b200.models.linear import Linear
    from oracle.torch_port import OracleLinear
    torch.manual_seed(5)
    mod = Linear((1, 16, 16), (1, 64, 64))
    oracle = OracleLinear(mod.state_dict(), (1, 64, 64), lr=1e-3, weight_decay=1e-5)
    g = torch.Generator().manual_seed(6)
    X, Y = torch.rand(22, 1, 16, 16, generator=g), torch.rand(22, 1, 64, 64, generator=g)
    eng = LinearEngine(mod, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(X, Y, 8)                                   # 8 + 8 + ragged 6
    for epoch in range(3):
        got = eng.train_epoch(data).cpu().numpy()
        want = [oracle.train_step(X[i:i + 8], Y[i:i + 8]) for i in range(0, 22, 8)]
        np.testing.assert_allclose(got, want, rtol=2e-5)
    assert rel_err(mod.linear[1].weight.detach().cpu().numpy(), oracle.w.detach().numpy()) < 1e-4
    assert rel_err(mod.linear[1].bias.detach().cpu().numpy(), oracle.b.detach().numpy()) < 1e-4
    np.testing.assert_allclose(eng.test_epoch(data).cpu().numpy(),
                               [oracle.test_loss(X[i:i + 8], Y[i:i + 8]) for i in range(0, 22, 8)], rtol=2e-5)
    out = []
    eng.score_batches(eng.bind(X, None, 16), lambda i, yh: out.append(yh.cpu().numpy().copy()))
    assert rel_err(np.concatenate(out), oracle.score(X).numpy()) < 1e-4


def test_linear_model_api_and_cli(tmp_path):
    """LinearModel.train / save / load / apply and `train_cae --method linear` + `apply_cae` (parameters.json type)"""
    from cae_tools_b200.cli import apply_cae, train_cae
    from cae_tools_b200.models.linear_model import LinearModel
    from cae_tools_b200.utils import xr_lite
    from oracle import datagen
    tr, te = datagen.circle_datasets(20, 10, input_size=(16, 16), output_size=(32, 32))
    torch.manual_seed(3)
    m = LinearModel(batch_size=5, nr_epochs=6, test_interval=2, lr=1e-3)
    m.verbose = False
    m.train(["lowres"], "hires", tr, te)
    assert len(m.history["train_loss"]) == 3 and m.history["train_loss"][-1] < m.history["train_loss"][0]
    folder = str(tmp_path / "lin")
    m.save(folder)
    params = json.load(open(os.path.join(folder, "parameters.json")))
    assert params["type"] == "LinearModel" and params["output_shape"] == [1, 32, 32]
    assert {"history.json", "normalisation.weights", "parameters.json", "summary.txt", "weights"} <= set(os.listdir(folder))
    m2 = LinearModel()
    m2.load(folder)
    m.apply(te, ["lowres"], "a")
    m2.apply(te, ["lowres"], "b")
    np.testing.assert_allclose(np.asarray(te["a"].data), np.asarray(te["b"].data), rtol=1e-6)
    paths = {}
    for name, seed in (("train", 0), ("test", 1)):
        lo, hi = datagen.generate(12, (16, 16), (32, 32), "circle", seed=seed)
        ds = xr_lite.Dataset()
        ds["lowres"] = xr_lite.DataArray(lo, dims=("n", "chan", "y1", "x1"))
        ds["hires"] = xr_lite.DataArray(hi, dims=("n", "chan", "y2", "x2"))
        paths[name] = str(tmp_path / f"{name}.nc")
        ds.to_netcdf(paths[name])
    cli_folder = str(tmp_path / "cli_model")
    train_cae.main(["--train-inputs", paths["train"], "--test-inputs", paths["test"], "--model-folder", cli_folder,
                    "--input-variables", "lowres", "--output-variable", "hires", "--nr-epochs", "3", "--batch-size", "4",
                    "--method", "linear"])
    assert json.load(open(os.path.join(cli_folder, "parameters.json")))["type"] == "LinearModel"
    out = str(tmp_path / "scores.nc")
    apply_cae.main([paths["test"], out, "--model-folder", cli_folder, "--prediction-variable", "est"])
    assert xr_lite.open_dataset(out)["est"].shape == (12, 1, 32, 32)


def test_linear_engine_vs_reference_fixture():
    """the CUDA LinearEngine against values the reference's Linear module + MSELoss + Adam produced (linear_mini.npz)"""
    from helpers import load_npz
    from cae_tools_b200.engine.linear import LinearEngine
    from cae_tools_b200.models.linear import Linear
    g = load_npz("linear_mini.npz")
    torch.manual_seed(int(g["seed"]))
    mod = Linear((1, 16, 16), (1, 64, 64))
    assert np.array_equal(mod.linear[1].weight.detach().numpy()[::16], g["init.weight_sub"])
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    eng = LinearEngine(mod, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(x, y, 8)
    losses = [float(eng.train_epoch(data).cpu()[0]) for _ in range(3)]
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-5)
    assert rel_err(mod.linear[1].weight.detach().cpu().numpy()[::16], g["after.weight_sub"]) < 1e-4
    assert rel_err(mod.linear[1].bias.detach().cpu().numpy(), g["after.bias"]) < 1e-4
    out = []
    eng.score_batches(eng.bind(x, None, 8), lambda i, yh: out.append(yh.cpu().numpy().copy()))
    assert rel_err(np.concatenate(out), g["pred"]) < 1e-4
