"""UNET variant on the CUDA path against the reference's own modules (golden fixtures written by oracle/gen_golden.py
from cae_tools.models.unet.Encoder / Decoder / masked_mse_loss / pearson_corr_torch + AdamW) and the oracle port."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import ROOT, load_npz, rel_err, spec_of, split_sd

pytestmark = pytest.mark.gpu

DEAD = ("encoder_lin.0.bias", "decoder_lin.0.bias")      # Linear -> BatchNorm1d: identically zero gradient


def _build(g):
    from cae_tools_b200.models.model_sizer import ModelSpec
    from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
    spec = ModelSpec()
    spec.load(spec_of(g))
    enc_sd, dec_sd = split_sd(g, "init.enc."), split_sd(g, "init.dec.")
    latent, fc = enc_sd["encoder_lin.4.weight"].shape[0], enc_sd["encoder_lin.0.weight"].shape[0]
    enc = UNetEncoder(spec.get_input_layers(), latent, fc, 0.0)
    dec = UNetDecoder(spec.get_output_layers(), latent, fc, 0.0)
    enc.load_state_dict(enc_sd)
    dec.load_state_dict(dec_sd)
    return spec, enc, dec


def _dead(k):
    return k in DEAD or (k.startswith("encoder_cnn") and k.endswith(".bias") and int(k.split(".")[1]) % 4 == 0)


@pytest.mark.parametrize("use_graphs", [False, True])
@pytest.mark.parametrize("name", ["nomask", "mask", "head16_mask", "head16_mask-chain"])
def test_unet_train_steps_vs_reference(name, use_graphs, monkeypatch):
    """`head16_mask` runs the stem as the two cooperative launches of unet_stem_train.cu (the default whenever the last
    layer is the fused patch head), `head16_mask-chain` the same fixture through the per-layer kernels"""
    from cae_tools_b200.engine import ops
    from cae_tools_b200.engine.unet import UNetEngine
    chain = name.endswith("-chain")
    name = name.replace("-chain", "")
    if chain:
        monkeypatch.setattr(UNetEngine, "use_fused_train_stem", False)
    g = load_npz(f"unet_{name}.npz")
    spec, enc, dec = _build(g)
    eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5, use_graphs=use_graphs)
    x, y, mask = (torch.from_numpy(g[k]) for k in ("x", "y", "mask"))
    data = eng.bind(x, y, x.shape[0], mask=mask if name.endswith("mask") and name != "nomask" else None)
    fused_head = name.startswith("head")       # kernel == stride last layer: patch_head.cu (never writes yhat in training)
    mses, pls = [], []
    for step in range(3):
        l = eng.train_epoch(data)
        mses.append(float(l.cpu()[0]))
        pls.append(float(data.pearson.cpu()[0]))
        if step == 0:
            b = eng._act_buffers(x.shape[0])
            st = eng._train_stem(x.shape[0])
            assert (st is not None) == (fused_head and not chain)
            if st is not None:
                names = [n for n, _ in eng._program("train", data, x.shape[0]).sched]
                assert "fwd.stem_train" in names and "bwd.stem_train" in names and len(names) <= 8, names
                y_e = [ops.stem_tape_view(st, f"y_e{i}") for i in range(len(b["y_e"]))]
                y_d = [ops.stem_tape_view(st, f"y_d{j}") for j in range(len(b["y_d"]))]
            else:
                y_e, y_d = b["y_e"], b["y_d"]
            for i, t in enumerate(y_e):
                assert rel_err(t.cpu().numpy(), g[f"act.enc.{4 * i}"]) < 1e-4, f"enc conv {i}"
            for j, t in enumerate(y_d):
                assert rel_err(t.cpu().numpy(), g[f"act.dec.{4 * j}"]) < 1e-4, f"dec convT {j}"
            if fused_head:
                assert eng._head is not None and float(b["yhat"].abs().max()) == 0.0
            else:
                assert rel_err(b["yhat"].cpu().numpy(), g["yhat"]) < 1e-4
            for prefix, mod in (("enc.", enc), ("dec.", dec)):
                for k, p in mod.named_parameters():
                    ref = g["grad." + prefix + k]
                    got = p.grad.detach().cpu().numpy()
                    if _dead(k):
                        assert np.abs(got).max() <= 1e-6 and np.abs(ref).max() <= 1e-5, k
                        continue
                    scale = max(np.abs(ref).max(), 1e-7)
                    # 1e-4 (north_star) where the fused patch head computes the loss statistics (centred moments); the generic
                    # loss kernel of the other fixtures (cae_masked_pearson_loss, raw moments) is held to 2e-4
                    tol = 1e-4 if fused_head else 2e-4
                    assert np.abs(got - ref).max() <= tol * scale + 1e-9, (k, np.abs(got - ref).max(), scale)
    np.testing.assert_allclose(mses, g["mse"], rtol=2e-5)
    np.testing.assert_allclose(pls, g["pearson_loss"], rtol=2e-5)
    for prefix, mod in (("enc.", enc), ("dec.", dec)):
        for k, v in mod.state_dict().items():
            ref = g["after3." + prefix + k]
            got = v.detach().cpu().numpy()
            if ref.dtype.kind != "f":
                assert int(got) == int(ref), k
            elif _dead(k):
                continue
            elif k.endswith("running_mean"):
                # running means inherit the lr-sized deviations of single parameters discussed below (measured worst
                # case 3.8e-4 of the feature's spread after 3 steps)
                spread = np.sqrt(g["after3." + prefix + k.replace("running_mean", "running_var")]).max()
                assert np.abs(got - ref).max() <= 6e-4 * max(np.abs(ref).max(), spread) + 1e-4, k
            else:
                # Adam divides by |g| + 1e-8: a gradient component of ~1e-8 (nearly dead ReLU unit) becomes a step of
                # rounding-sensitive size (measured: step-0 gradients agree to 2e-5 of the tensor's max-norm, yet two
                # of eight latent biases move by 0.89*lr instead of 1.0*lr).  Typical element tight, worst < 0.2 lr/step.
                # Elements that move differently must be exactly those whose gradient sits at rounding-noise level (there
                # the summation order decides the sign Adam sees), they must be rare, and no element can be off by more
                # than the 3 steps x lr a sign flip costs.
                dev = np.abs(got - ref)
                assert np.median(dev) <= 3e-4 * max(np.abs(ref).max(), 1e-3) + 1e-6, k
                assert dev.max() <= 2.0 * 1e-3 * 3, k
                loose = dev > 0.2 * 1e-3 * 3
                if loose.any():
                    g0 = np.abs(g["grad." + prefix + k])
                    assert loose.sum() <= max(2, 5e-3 * loose.size), (k, int(loose.sum()), loose.size)
                    assert g0[loose].max() <= 1e-4 * g0.max(), (k, g0[loose].max(), g0.max())
    out = []
    eng.score_batches(eng.bind(x, None, x.shape[0]), lambda i, yh: out.append(yh.cpu().numpy().copy()))
    assert rel_err(out[0], g["eval_yhat"]) < 1e-3


def test_unet_model_api_train_save_load_apply(tmp_path):
    """UNET(...).train with a layer-definitions spec and a mask variable, save -> load -> apply"""
    from cae_tools_b200.models.model_sizer import ModelSpec
    from cae_tools_b200.models.unet import UNET
    from oracle import datagen
    g = load_npz("unet_mask.npz")
    tr, te = datagen.circle_datasets(24, 12, input_size=(16, 16), output_size=(64, 64))
    rng = np.random.RandomState(0)
    tr.add("valid", (rng.rand(24, 1, 64, 64) > 0.2).astype(np.float32))
    te.add("valid", (rng.rand(12, 1, 64, 64) > 0.2).astype(np.float32))
    torch.manual_seed(11)
    m = UNET(batch_size=8, nr_epochs=12, test_interval=4, encoded_dim_size=8, fc_size=32, dropout_rate=0.0,
             lambda_pearson=0.5)
    m.verbose = False
    spec = ModelSpec()
    spec.load(spec_of(g))
    m.spec = spec
    m.train(["lowres"], "hires", tr, te, mask_variable_name="valid")
    assert len(m.history["train_loss"]) == 3 and m.history["train_loss"][-1] < m.history["train_loss"][0]
    folder = str(tmp_path / "unet")
    m.save(folder)
    params = json.load(open(os.path.join(folder, "parameters.json")))
    assert params["type"] == "UNET" and params["lambda_pearson"] == 0.5 and params["dropout_rate"] == 0.0
    m2 = UNET(dropout_rate=0.0)
    m2.load(folder)
    m2.apply(te, ["lowres"], "est")
    m.apply(te, ["lowres"], "est0")
    np.testing.assert_allclose(np.asarray(te["est"].data), np.asarray(te["est0"].data), rtol=1e-6)
    # dropout > 0 is refused for training, loudly
    m3 = UNET(batch_size=8, nr_epochs=1, encoded_dim_size=8, fc_size=32, dropout_rate=0.1)
    m3.verbose = False
    m3.spec = spec
    with pytest.raises(NotImplementedError):
        m3.train(["lowres"], "hires", tr, te)


def _shipped_spec():
    from cae_tools_b200.models.model_sizer import ModelSpec
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cae_tools_b200", "specs",
                        "unet_16x16_256x256.json")
    spec = ModelSpec()
    spec.load(json.load(open(path)))
    return spec, json.load(open(path))


@pytest.mark.parametrize("with_mask", [False, True])
@pytest.mark.parametrize("batch", [5, 64])
def test_patch_head_k32_vs_oracle_and_generic(batch, with_mask):
    """the shipped 16x16 -> 256x256 spec (k32 s32 head): fused head == oracle port (torch CPU) == generic kernels:
    losses, every gradient, 2 AdamW steps, eval prediction"""
    from cae_tools_b200.engine.unet import UNetEngine
    from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
    from oracle.torch_port import OracleUNet
    spec, spec_json = _shipped_spec()
    torch.manual_seed(3)
    gen = torch.Generator().manual_seed(17)
    x, y = torch.rand(batch, 1, 16, 16, generator=gen), torch.rand(batch, 1, 256, 256, generator=gen)
    mask = (torch.rand(batch, 1, 256, 256, generator=gen) > 0.25).float() if with_mask else None
    engines = []
    for fused in (True, False):
        torch.manual_seed(3)
        enc = UNetEncoder(spec.get_input_layers(), 8, 32, 0.0)
        dec = UNetDecoder(spec.get_output_layers(), 8, 32, 0.0)
        if fused:
            oracle = OracleUNet(enc.state_dict(), dec.state_dict(), spec_json, lambda_pearson=0.7,
                                zero_dead_bias_grads=True)
            exact = OracleUNet(enc.state_dict(), dec.state_dict(), spec_json, lambda_pearson=0.7,
                               zero_dead_bias_grads=True, dtype=torch.float64)
        eng = UNetEngine(enc, dec, lambda_pearson=0.7, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5)
        eng.use_patch_head = fused
        eng.use_fused_attention = fused        # second engine: unfused attention chain + generic conv / loss kernels
        eng.use_fused_fc = fused               # one-launch fc bottleneck (off by default) against the cae_gemm chain
        eng.use_fused_stem = fused             # one-launch eval stem (off by default) against the layer-by-layer forward
        engines.append((eng, enc, dec, eng.bind(x, y, batch, mask=mask)))
    ones = torch.ones_like(y)
    for step in range(2):
        want = oracle.train_step(x, y, mask if with_mask else ones)
        for eng, enc, dec, data in engines:
            mse = float(eng.train_epoch(data).cpu()[0])
            pl = float(data.pearson.cpu()[0])
            assert abs(mse - want[0]) <= 2e-5 * want[0] and abs(pl - want[1]) <= 2e-5 * abs(want[1]), (step, mse, pl, want)
        if step == 0:
            assert engines[0][0]._head is not None and engines[1][0]._head is None
            # Bar: 1e-4 of the tensor's max-norm against the fp32 oracle.  Where two fp32 evaluations of the same step
            # differ by more than that (sums with heavy cancellation: the head bias, BatchNorm-coupled deep layers), the
            # float64 evaluation adjudicates: the CUDA result must be at least as close to it as the fp32 oracle is (x2).
            exact.train_step(x.double(), y.double(), (mask if with_mask else ones).double())
            beyond = {}
            for prefix, sd, sd64 in (("enc", oracle.enc, exact.enc), ("dec", oracle.dec, exact.dec)):
                mods = [dict(e[1 if prefix == "enc" else 2].named_parameters()) for e in engines]
                for k, ref in sd.items():
                    if not ref.requires_grad:
                        continue
                    r, r64 = ref.grad.numpy(), sd64[k].grad.numpy()
                    scale = max(np.abs(r).max(), 1e-7)
                    for mi, m in enumerate(mods):
                        got = m[k].grad.detach().cpu().numpy()
                        err = np.abs(got - r).max()
                        if err > 1e-4 * scale + 1e-9:
                            e_gpu, e_cpu = np.abs(got - r64).max(), np.abs(r - r64).max()
                            assert e_gpu <= 2.0 * e_cpu + 1e-9 and err <= 1e-3 * scale, (prefix, k, mi, err, scale, e_gpu, e_cpu)
                            beyond[(k, mi)] = (float(err / scale), float(e_gpu / scale), float(e_cpu / scale))
            print("gradients beyond 1e-4 of the fp32 oracle, adjudicated by float64 (rel: vs oracle, gpu vs f64, oracle vs f64):", beyond)
    ref = oracle.score(x).numpy()
    for eng, enc, dec, data in engines:
        out = []
        eng.score_batches(eng.bind(x, None, batch), lambda i, yh: out.append(yh.cpu().numpy().copy()))
        assert rel_err(out[0], ref) < 1e-3
    # test epoch (eval-mode loss) through the fused head == through the generic kernels
    l0 = float(engines[0][0].test_epoch(engines[0][3]).cpu()[0])
    l1 = float(engines[1][0].test_epoch(engines[1][3]).cpu()[0])
    assert abs(l0 - l1) <= 1e-5 * abs(l1)


def test_patch_head_ragged_tail_and_multichannel():
    """kernel == stride head with 2 output channels, a per-channel mask and a ragged last batch (7 = 4 + 3)"""
    from cae_tools_b200.engine.unet import UNetEngine
    from cae_tools_b200.models.model_sizer import ModelSpec
    from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
    from oracle.torch_port import OracleUNet
    _, spec_json = _shipped_spec()
    spec_json["output_layers"][-1].update(kernel_size=16, stride=16, output_dimensions=[2, 128, 128])
    spec = ModelSpec()
    spec.load(spec_json)
    gen = torch.Generator().manual_seed(5)
    x, y = torch.rand(7, 1, 16, 16, generator=gen), torch.rand(7, 2, 128, 128, generator=gen)
    mask = (torch.rand(7, 2, 128, 128, generator=gen) > 0.4).float()
    torch.manual_seed(9)
    enc = UNetEncoder(spec.get_input_layers(), 8, 32, 0.0)
    dec = UNetDecoder(spec.get_output_layers(), 8, 32, 0.0)
    oracle = OracleUNet(enc.state_dict(), dec.state_dict(), spec_json, lambda_pearson=1.0, zero_dead_bias_grads=True)
    eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(x, y, 4, mask=mask)
    for epoch in range(2):
        got = eng.train_epoch(data).cpu().numpy()
        want = [oracle.train_step(x[:4], y[:4], mask[:4])[0], oracle.train_step(x[4:], y[4:], mask[4:])[0]]
        np.testing.assert_allclose(got, want, rtol=5e-5)
    assert eng._head is not None
    out = []
    eng.score_batches(eng.bind(x, None, 4), lambda i, yh: out.append(yh.cpu().numpy().copy()))
    assert rel_err(np.concatenate(out), oracle.score(x).numpy()) < 1e-3


@pytest.mark.parametrize("batch,p", [(5, 0.1), (64, 0.3)])
def test_unet_dropout_matches_oracle_with_identical_masks(batch, p):
    """dropout_rate > 0 (the reference's default is 0.1: unet.py:74,115,203, cli/train_cae.py:39): the fused stem draws its
    masks from a counter-based hash; the oracle applies the SAME masks (oracle/dropout_hash.py) through torch autograd -
    losses of 3 steps (fresh masks every step), every gradient of step 0, keep rate and 1/(1-p) scaling"""
    from cae_tools_b200.engine.unet import UNetEngine
    from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
    from oracle.dropout_hash import drop_mask
    from oracle.torch_port import OracleUNet
    spec, spec_json = _shipped_spec()
    torch.manual_seed(9)
    gen = torch.Generator().manual_seed(23)
    x, y = torch.rand(batch, 1, 16, 16, generator=gen), torch.rand(batch, 1, 256, 256, generator=gen)
    ones = torch.ones_like(y)
    enc = UNetEncoder(spec.get_input_layers(), 8, 32, p)
    dec = UNetDecoder(spec.get_output_layers(), 8, 32, p)
    seed = 0x1234ABCD5678
    oracle = OracleUNet(enc.state_dict(), dec.state_dict(), spec_json, lambda_pearson=0.7, zero_dead_bias_grads=True)
    exact = OracleUNet(enc.state_dict(), dec.state_dict(), spec_json, lambda_pearson=0.7, zero_dead_bias_grads=True,
                       dtype=torch.float64)
    oracle.set_dropout(p, seed)
    exact.set_dropout(p, seed)
    eng = UNetEngine(enc, dec, lambda_pearson=0.7, dropout_rate=p, seed=seed, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(x, y, batch)
    m = drop_mask(p, seed, 0, 8, 4096, 1024)
    assert abs(float((m == 0).mean()) - p) < 5e-3 and abs(float(m.max()) - 1.0 / (1.0 - p)) < 1e-6
    for step in range(3):
        want = oracle.train_step(x, y, ones)
        mse = float(eng.train_epoch(data).cpu()[0])
        pl = float(data.pearson.cpu()[0])
        # the masks make the loss landscape rougher: a single differently-rounded ReLU / mask decision is invisible at 1e-4
        assert abs(mse - want[0]) <= 1e-4 * want[0] and abs(pl - want[1]) <= 1e-4 * abs(want[1]), (step, mse, pl, want)
        if step == 0:
            st = eng._train_stem(batch)
            assert st is not None, "dropout needs the fused training stem"
            # the activated head input carries the mask of site 8 + (n_up - 1): zeros exactly where the oracle mask is zero
            hin = st.t_hin.cpu().numpy().reshape(batch, -1)
            mk = drop_mask(p, seed, 0, 8 + 1, batch, hin.shape[1])
            assert np.all(hin[mk == 0] == 0.0)
            exact.train_step(x.double(), y.double(), ones.double())
            for sd, sd64, mod in ((oracle.enc, exact.enc, enc), (oracle.dec, exact.dec, dec)):
                for k, prm in mod.named_parameters():
                    r, r64 = sd[k].grad.numpy(), sd64[k].grad.numpy()
                    got = prm.grad.detach().cpu().numpy()
                    scale = max(np.abs(r).max(), 1e-7)
                    err = np.abs(got - r).max()
                    if err > 1e-4 * scale + 1e-9:
                        if k == "decoder_conv.8.bias":
                            # the head bias gradient is ONE scalar: the sum of N * 65 536 cancelling terms (sum |dz| ~ 10^3 x
                            # |sum dz|), each formed with the SFU-approximate sigmoid of patch_head.cu (2e-7 relative, not
                            # zero-mean): 8e-4 of its value here (2e-4 at p = 0).  Accumulation itself is fp64.
                            assert err <= 2e-3 * scale, (k, err, scale)
                            continue
                        e_gpu, e_cpu = np.abs(got - r64).max(), np.abs(r - r64).max()
                        assert e_gpu <= 2.0 * e_cpu + 1e-9 and err <= 1e-3 * scale, (k, err, scale, e_gpu, e_cpu)
    # eval mode ignores dropout
    out = []
    eng.score_batches(eng.bind(x, None, batch), lambda i, yh: out.append(yh.cpu().numpy().copy()))
    assert rel_err(out[0], oracle.score(x).numpy()) < 1e-3


def test_train_cae_unet_runs_with_the_reference_default_flags(tmp_path):
    """`train_cae --method unet` with default flags (--dropout-rate 0.1, reference cli/train_cae.py:39,135) trains, saves,
    reloads (parameters.json carries dropout_rate 0.1) and continues"""
    from cae_tools_b200.cli import apply_cae, train_cae
    from cae_tools_b200.utils import xr_lite
    from oracle import datagen
    paths = {}
    for name, seed in (("train", 0), ("test", 1)):
        lo, hi = datagen.generate(24, (16, 16), (256, 256), "circle", seed=seed)
        ds = xr_lite.Dataset()
        ds["lowres"] = xr_lite.DataArray(lo, dims=("n", "chan", "y1", "x1"))
        ds["hires"] = xr_lite.DataArray(hi, dims=("n", "chan", "y2", "x2"))
        paths[name] = str(tmp_path / f"{name}.nc")
        ds.to_netcdf(paths[name])
    folder = str(tmp_path / "model")
    spec_path = os.path.join(ROOT, "cae_tools_b200", "specs", "unet_16x16_256x256.json")
    common = ["--train-inputs", paths["train"], "--test-inputs", paths["test"], "--model-folder", folder, "--input-variables", "lowres",
              "--output-variable", "hires", "--method", "unet", "--layer-definitions-path", spec_path]
    train_cae.main(common + ["--nr-epochs", "6"])
    params = json.load(open(os.path.join(folder, "parameters.json")))
    assert params["type"] == "UNET" and abs(params["dropout_rate"] - 0.1) < 1e-12
    hist = json.load(open(os.path.join(folder, "history.json")))
    assert np.isfinite(hist["train_loss"]).all() and np.isfinite(hist["test_loss"]).all()
    train_cae.main(common + ["--nr-epochs", "3", "--continue-training"])
    assert json.load(open(os.path.join(folder, "history.json")))["nr_epochs"] == 9
    out = str(tmp_path / "scores.nc")
    apply_cae.main([paths["test"], out, "--model-folder", folder, "--prediction-variable", "est"])
    assert np.isfinite(xr_lite.open_dataset(out)["est"].values).all()


def test_unet_k32_spec_vs_reference_fixture():
    """the SHIPPED spec (k32 s32 head, the bench workload) against values the reference itself produced
    (unet_head32_mask_light.npz): step-0 gradients, 3 AdamW steps of losses, eval prediction with the initial weights
    (identical weights -> apply() parity at 1e-4) and with the trained ones"""
    from helpers import unet_light_data
    from cae_tools_b200.engine.unet import UNetEngine
    g = load_npz("unet_head32_mask_light.npz")
    L = int(g["light"])
    spec, enc, dec = _build(g)
    x, y, mask = unet_light_data(g)
    eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5)
    out = []
    eng.score_batches(eng.bind(x, None, x.shape[0]), lambda i, yh: out.append(yh.cpu().numpy().copy()))
    assert rel_err(out[0][:, :, ::L, ::L], g["eval_yhat_init"]) < 1e-4          # identical weights: apply() bar
    data = eng.bind(x, y, x.shape[0], mask=mask)
    mses, pls, worst = [], [], 0.0
    for step in range(3):
        mses.append(float(eng.train_epoch(data).cpu()[0]))
        pls.append(float(data.pearson.cpu()[0]))
        if step == 0:
            assert eng._train_stem(x.shape[0]) is not None and eng._head is not None
            for prefix, mod in (("enc.", enc), ("dec.", dec)):
                for k, p in mod.named_parameters():
                    ref = g["grad." + prefix + k]
                    got = p.grad.detach().cpu().numpy()
                    if _dead(k):
                        assert np.abs(got).max() <= 1e-6 and np.abs(ref).max() <= 1e-5, k
                        continue
                    scale = max(np.abs(ref).max(), 1e-7)
                    worst = max(worst, float(np.abs(got - ref).max() / scale))
                    assert np.abs(got - ref).max() <= 1e-4 * scale + 1e-9, (k, np.abs(got - ref).max(), scale)
    print(f"worst step-0 gradient deviation from the reference (max-norm relative): {worst:.2e}")
    np.testing.assert_allclose(mses, g["mse"], rtol=2e-5)
    np.testing.assert_allclose(pls, g["pearson_loss"], rtol=2e-5)
    out = []
    eng.score_batches(eng.bind(x, None, x.shape[0]), lambda i, yh: out.append(yh.cpu().numpy().copy()))
    assert rel_err(out[0][:, :, ::L, ::L], g["eval_yhat"]) < 1e-3


def test_unet_apply_with_identical_weights_k16_fixture():
    """eval-mode prediction with the reference's initial weights / BatchNorm buffers (head16 fixture): 1e-4"""
    from cae_tools_b200.engine.unet import UNetEngine
    g = load_npz("unet_head16_mask.npz")
    spec, enc, dec = _build(g)
    eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0)
    x = torch.from_numpy(g["x"])
    out = []
    eng.score_batches(eng.bind(x, None, x.shape[0]), lambda i, yh: out.append(yh.cpu().numpy().copy()))
    assert rel_err(out[0], g["eval_yhat_init"]) < 1e-4


def test_unet_loss_curve_50_epochs_vs_reference():
    """BASELINE configs[1] / north_star: per-epoch train / test loss of UNET.train (batch 64 = 64 + 36 on the 100-case circle
    set, shipped spec, dropout 0, 50 epochs, test_interval 1) against the reference's own loop (curve_unet_b64_e50.npz,
    oracle/gen_golden.py:gen_unet_curve).  Bar: 1e-3 relative per epoch."""
    from cae_tools_b200.models.model_sizer import ModelSpec
    from cae_tools_b200.models.unet import UNET
    from oracle import datagen
    g = load_npz("curve_unet_b64_e50.npz")
    tr, te = datagen.circle_datasets(100, 100)
    torch.manual_seed(1234)
    m = UNET(batch_size=64, nr_epochs=50, test_interval=1, encoded_dim_size=4, fc_size=16, lr=1e-3, weight_decay=1e-5,
             dropout_rate=0.0, lambda_pearson=1.0)
    m.verbose = False
    spec = ModelSpec()
    spec.load(spec_of(g))
    m.spec = spec
    m.train(["lowres"], "hires", tr, te)
    got_tr, got_te = np.array(m.history["train_loss"]), np.array(m.history["test_loss"])
    assert got_tr.shape == (50,)
    dtr, dte = np.abs(got_tr - g["train_loss"]) / g["train_loss"], np.abs(got_te - g["test_loss"]) / g["test_loss"]
    print(f"unet 50-epoch curve: max rel dev train {dtr.max():.2e} (epoch {dtr.argmax()}), test {dte.max():.2e} (epoch {dte.argmax()}); "
          f"first 10 epochs {max(dtr[:10].max(), dte[:10].max()):.2e}")
    # Bar: 1e-3 relative per epoch - or three times the reference's OWN spread, whichever is larger.  The reference's 50-epoch
    # curve moves by up to 6.4e-4 (test loss) when only its CPU thread count changes (curve_unet_b64_e50_envelope.npz,
    # oracle/gen_unet_envelope.py: 1, 4 and 8 threads - three samples of the same chaotic amplification of rounding-level
    # differences; the final WEIGHTS of those runs, like ours, differ by tens of percent in flat directions).  Measured here:
    # train <= 4.2e-4 at every epoch; test (eval mode: running statistics) <= 1e-3 up to epoch 38, peak 1.44e-3 at epoch 42
    # = 2.2x the reference's own spread, back to 0.9e-3 at epoch 49 (tools/unet_curve_probe.py).  The first epochs, before
    # the amplification, agree to 1e-5.
    env = load_npz("curve_unet_b64_e50_envelope.npz")
    for dev, ref, key in ((dtr, g["train_loss"], "train"), (dte, g["test_loss"], "test")):
        spread = np.maximum.accumulate(np.max([np.abs(env[f"{key}_t{t}"] - ref) / ref for t in (1, 4)], axis=0))
        bar = np.maximum(1e-3, 3.0 * spread)
        assert np.all(dev <= bar), (key, int(np.argmax(dev - bar)), float(dev.max()))
    assert dte[:38].max() <= 1e-3
    assert dtr.max() <= 1e-3                                  # the training-mode curve meets the fixed 1e-3 bar outright
    assert max(dtr[:10].max(), dte[:10].max()) <= 3e-4
    # identical-weights apply(): the reference's trained weights -> predictions within 1e-4
    m.encoder.load_state_dict(split_sd(g, "final.enc."))
    m.decoder.load_state_dict(split_sd(g, "final.dec."))
    m.engine = None
    m.apply(te, ["lowres"], "est")
    lo, hi = m.normalisation_parameters[2], m.normalisation_parameters[3]
    pred = (np.asarray(te["est"].data)[:4] - lo) / (hi - lo)
    assert np.max(np.abs(pred[:, :, ::8, ::8] - g["pred_sub"])) < 1e-4


def test_masked_loss_per_batch_scale_makes_rank_shares_sum_to_the_global_batch():
    """Data parallelism with a land / sea mask: a rank forms SQ_local / CNT_local; with the constant
    count_scale = 1 / world the sum over ranks is the MEAN of the per-share masked MSEs, not the global masked MSE.  The
    per-batch `mse_scale` = CNT_local / CNT_global (engine/unet.py:_mse_scale, all-reduced at bind time) makes loss and
    gradient of the shares add up to those of the global batch exactly.  Here: two "ranks" evaluated one after the other on
    one GPU - (a) the generic loss kernel with its gradient, (b) the fused patch head through an eval-mode engine."""
    from cae_tools_b200.engine import ops
    from cae_tools_b200.engine.unet import UNetEngine
    from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(11)
    N, H = 4, 32
    pred = (0.05 + 0.9 * torch.rand(N, 1, H, H, generator=gen)).to(dev)
    tgt = torch.rand(N, 1, H, H, generator=gen).to(dev)
    keep = torch.tensor([0.9, 0.8, 0.3, 0.1]).view(N, 1, 1, 1)
    mask = (torch.rand(N, 1, H, H, generator=gen) < keep).float().to(dev)

    def run(lo, hi, count_scale, mse_scale):
        n = hi - lo
        p, t, m = pred[lo:hi].contiguous(), tgt[lo:hi].contiguous(), mask[lo:hi].contiguous()
        moments = torch.zeros(n * 7, dtype=torch.float64, device=dev)
        coef, scalars = torch.zeros(n * 3, device=dev), torch.zeros(3, device=dev)
        loss, pl = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
        dz, ps = torch.zeros_like(p), torch.zeros(n, device=dev)
        ops.masked_pearson_loss(ops.view4(p), ops.make_src(t), ops.make_src(m), 1, 0.7, count_scale, moments, coef, scalars,
                                loss, pl, ops.view4(dz), ps, mse_scale=mse_scale)
        torch.cuda.synchronize()
        return float(loss), float(pl), dz.cpu()

    L, P, DZ = run(0, N, 1.0, None)
    total = float(mask.sum())
    sc = [torch.tensor([float(mask[lo:hi].sum()) / total], device=dev) for lo, hi in ((0, 2), (2, 4))]
    parts = [run(lo, hi, 0.5, s) for (lo, hi), s in zip(((0, 2), (2, 4)), sc)]
    assert abs(parts[0][0] + parts[1][0] - L) <= 1e-6 * L
    assert abs(parts[0][1] + parts[1][1] - P) <= 1e-5 * abs(P)
    dz = torch.cat([parts[0][2], parts[1][2]])
    assert float((dz - DZ).abs().max()) <= 2e-6 * float(DZ.abs().max())
    # the constant scale alone does NOT reproduce the global batch for these unequal shares (the test has teeth)
    plain = [run(lo, hi, 0.5, None) for lo, hi in ((0, 2), (2, 4))]
    assert abs(plain[0][0] + plain[1][0] - L) > 1e-3 * L

    # (b) fused patch head, eval-mode engine (running statistics: no coupling between the samples of a batch)
    spec, _ = _shipped_spec()
    torch.manual_seed(3)
    enc, dec = UNetEncoder(spec.get_input_layers(), 8, 32, 0.0), UNetDecoder(spec.get_output_layers(), 8, 32, 0.0)
    x, y = torch.rand(N, 1, 16, 16, generator=gen), torch.rand(N, 1, 256, 256, generator=gen)
    m = (torch.rand(N, 1, 256, 256, generator=gen) < keep).float()
    full = UNetEngine(enc, dec, lambda_pearson=0.7, dropout_rate=0.0)
    want = float(full.test_epoch(full.bind(x, y, N, mask=m)).cpu()[0])
    share = UNetEngine(enc, dec, lambda_pearson=0.7, dropout_rate=0.0, count_scale=0.5)
    got = 0.0
    for lo, hi in ((0, 2), (2, 4)):
        data = share.bind(x[lo:hi], y[lo:hi], 2, mask=m[lo:hi])
        data.mse_scale = torch.tensor([float(m[lo:hi].sum() / m.sum())], device=dev)
        names = [n for n, _ in share._program("test", data, 2).sched]
        assert any("head" in n for n in names), names
        got += float(share.test_epoch(data).cpu()[0])
    assert abs(got - want) <= 2e-6 * want, (got, want)
