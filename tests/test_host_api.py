"""Host-side behaviour that needs no GPU: data adaptor semantics, shuffle order, model-folder layout."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import load_npz, split_sd
from oracle import datagen


def test_dsdataset_normalisation_matches_reference_semantics():
    """reference: models/ds_dataset.py:49-75,99-135 (global min/max per variable, NaN rejection)"""
    from cae_tools_b200.models.ds_dataset import DSDataset
    rng = np.random.RandomState(0)
    a = (rng.rand(7, 1, 4, 5) * 10 + 280).astype(np.float32)
    b = np.full((7, 2, 4, 5), 3.0, dtype=np.float32)       # zero range -> normalises to 0
    y = (rng.rand(7, 1, 8, 10) * 5 + 285).astype(np.float32)
    ds = datagen.ArrayDataset().add("a", a).add("b", b).add("y", y)
    d = DSDataset(ds, ["a", "b"], "y")
    assert d.get_input_shape() == (3, 4, 5) and d.get_output_shape() == (1, 8, 10)
    X = d.input_array()
    np.testing.assert_allclose(X[:, 0], ((a - a.min()) / (a.max() - a.min()))[:, 0], rtol=1e-6)
    assert np.all(X[:, 1:] == 0)
    Y = d.output_array()
    assert Y.dtype == np.float32 and Y.min() == 0.0 and Y.max() == 1.0
    np.testing.assert_allclose(d.denormalise_output(Y), y, rtol=1e-6)
    item = d[3]
    np.testing.assert_array_equal(item[0], X[3])
    np.testing.assert_array_equal(item[1], Y[3])
    assert item[3] == "image3" and item[2].shape == (3, 4, 5)
    params = d.get_normalisation_parameters()
    assert json.loads(json.dumps(params)) == params          # JSON-serialisable, as saved to normalisation.weights
    order = [4, 0, 6]
    np.testing.assert_array_equal(d.input_array(order), X[order])
    bad = y.copy()
    bad[0, 0, 0, 0] = np.nan
    try:
        DSDataset(datagen.ArrayDataset().add("a", a).add("y", bad), ["a"], "y")
        assert False, "NaN output must be rejected"
    except ValueError:
        pass


def test_shuffle_order_consumes_rng_like_dataloader():
    from cae_tools_b200.models.conv_ae_model import shuffled_order
    torch.manual_seed(5)
    mine = [shuffled_order(23, 5), shuffled_order(11, 4)]
    torch.manual_seed(5)
    ref = []
    for n, bs in ((23, 5), (11, 4)):
        loader = torch.utils.data.DataLoader(torch.arange(n), batch_size=bs, shuffle=True)
        ref.append([int(i) for batch in loader for i in batch])
    assert mine == ref and sorted(mine[0]) == list(range(23))


def test_model_folder_layout_and_state_dict_keys(tmp_path):
    """save()/load() keep the reference's folder layout and state_dict keys (conv_ae_model.py:101-183)"""
    from cae_tools_b200.models.conv_ae_model import ConvAEModel
    from cae_tools_b200.models.model_sizer import create_model_spec
    g = load_npz("curve_conv_b64_e5.npz")
    m = ConvAEModel(encoded_dim_size=4, fc_size=16, batch_size=64)
    m.input_shape, m.output_shape = (1, 16, 16), (1, 256, 256)
    m.spec = create_model_spec(input_size=(16, 16), input_channels=1, output_size=(256, 256), output_channels=1)
    m._build_modules()
    ref_enc, ref_dec = split_sd(g, "final.enc."), split_sd(g, "final.dec.")
    assert list(m.encoder.state_dict().keys()) == list(ref_enc.keys())
    assert list(m.decoder.state_dict().keys()) == list(ref_dec.keys())
    m.encoder.load_state_dict(ref_enc)
    m.decoder.load_state_dict(ref_dec)
    m.normalisation_parameters = json.loads(str(g["norm"]))
    m.set_input_spec([{"name": "lowres", "shape": [1, 16, 16]}])
    m.set_output_spec({"name": "hires", "shape": [1, 256, 256]})
    folder = str(tmp_path / "model")
    m.save(folder)
    assert sorted(os.listdir(folder)) == sorted(["encoder.weights", "decoder.weights", "normalisation.weights",
                                                 "parameters.json", "spec.json", "history.json", "summary.txt",
                                                 "input_spec.json", "output_spec.json"])
    params = json.load(open(os.path.join(folder, "parameters.json")))
    ref_params = json.loads(str(g["params_json"]))
    assert params["type"] == "ConvAEModel"
    assert {k: v for k, v in params.items() if k != "model_id"}.keys() == ref_params.keys()
    m2 = ConvAEModel()
    m2.load(folder)
    assert m2.get_model_id() == m.get_model_id() and m2.spec.save() == m.spec.save()
    for k, v in m2.decoder.state_dict().items():
        assert torch.equal(v, ref_dec[k]), k
    assert m2.get_input_variable_names() == ["lowres"] and m2.get_output_variable_name() == "hires"
    assert m2.summary() == open(os.path.join(folder, "summary.txt")).read()
    assert "Latent Vector" in m2.summary()


def test_xr_lite_netcdf_roundtrip(tmp_path):
    from cae_tools_b200.utils import xr_lite
    ds = xr_lite.Dataset()
    ds["lowres"] = xr_lite.DataArray(np.arange(24, dtype=np.float32).reshape(2, 1, 3, 4), dims=("n", "chan", "y1", "x1"))
    ds["hires"] = xr_lite.DataArray(np.ones((2, 1, 6, 8), dtype=np.float64), dims=("n", "chan", "y2", "x2"))
    path = str(tmp_path / "t.nc")
    ds.to_netcdf(path)
    back = xr_lite.open_dataset(path)
    np.testing.assert_array_equal(back["lowres"].values, ds["lowres"].values)
    assert back["hires"].dims == ("n", "chan", "y2", "x2")
    both = xr_lite.open_mfdataset([path, path], concat_dim="box", combine="nested")
    assert both["lowres"].shape == (4, 1, 3, 4)


def test_bench_algorithmic_bytes_convention():
    """B_train = 2 in + 5 inter + 5 out, B_apply = in + 2 inter + out (fp32) - the roofline numerators"""
    import bench
    spec, enc, dec = bench.build_modules("conv")
    bt, ba = bench.bytes_per_sample(spec, bench.FC, bench.LATENT)
    assert (bt, ba) == (2547488, 757056)                       # the figures SURVEY 8(d) states for config 1
    uspec, _, _ = bench.build_modules("unet")
    bt, ba = bench.bytes_per_sample(uspec, bench.FC, bench.LATENT)
    inter = 4 * (8 * 8 * 8 + 16 * 4 * 4 + 32 * 2 * 2 + 16 * 4 * 4 + 8 * 8 * 8 + 16 + 4 + 16 + 32 * 2 * 2)
    assert bt == 2 * 1024 + 5 * inter + 5 * 262144 and ba == 1024 + 2 * inter + 262144
    # per-launch bytes of the fused head: input + one pass over the target (forward), + act / gradient of the input (backward)
    assert bench.op_bytes("fwd.head2+sigmoid+loss", uspec, 64) == 4 * 64 * (16 * 8 * 8 + 256 * 256)
    assert bench.op_bytes("bwd.head2", uspec, 64) == 4 * 64 * (3 * 16 * 8 * 8 + 256 * 256)
    assert bench.op_bytes("bwd.convT0.db", uspec, 64) is None


def test_linear_container_matches_reference_module_tree():
    """`--method linear`: same state_dict keys and the same initial values for the same seed as the reference's module
    (skipped where the reference tree is not mounted)"""
    import importlib.util
    import torch
    path = "/root/reference/src/cae_tools/models/linear.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not available")
    spec = importlib.util.spec_from_file_location("_ref_linear", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from cae_tools_b200.models.linear import Linear
    torch.manual_seed(1)
    a = Linear((1, 4, 4), (2, 3, 3)).state_dict()
    torch.manual_seed(1)
    b = ref.Linear((1, 4, 4), (2, 3, 3)).state_dict()
    assert list(a.keys()) == list(b.keys()) == ["linear.1.weight", "linear.1.bias"]
    assert all(torch.equal(a[k], b[k]) for k in a)
    from cae_tools_b200.models.linear_model import LinearModel
    m = LinearModel(batch_size=7, lr=0.01)
    assert (m.batch_size, m.lr, m.weight_decay, m.test_interval) == (7, 0.01, 1e-5, 10)
    assert m.summary() == "Model has not been trained"
