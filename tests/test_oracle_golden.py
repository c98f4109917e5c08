"""Pin the oracle: the torch restatement (oracle/torch_port.py) and the numpy definitions
(oracle/numpy_ops.py) must reproduce what the LIVE reference produced (tests/golden/*, written by
oracle/gen_golden.py in the build container)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import load_npz, rel_err, spec_of, split_sd
from oracle import datagen, numpy_ops
from oracle.torch_port import OracleModel, make_batches, shuffled_order

torch.set_num_threads(min(8, torch.get_num_threads()))


@pytest.mark.parametrize("name", ["mini", "nonsquare", "multich"])
def test_port_matches_reference_layers(name):
    g = load_npz(f"layers_{name}.npz")
    spec = spec_of(g)
    m = OracleModel(split_sd(g, "init.enc."), split_sd(g, "init.dec."), spec)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    trace = []
    yhat = m.forward(x, True, trace)
    loss = F.mse_loss(yhat, y)
    loss.backward()
    acts = sorted((k for k in g if k.startswith("act.enc.")), key=lambda s: int(s.split(".")[-1])) + \
        sorted((k for k in g if k.startswith("act.dec.")), key=lambda s: int(s.split(".")[-1]))
    assert len(acts) == len(trace)
    for k, t in zip(acts, trace):
        assert rel_err(t.detach().numpy(), g[k]) < 1e-5, k
    assert rel_err(yhat.detach().numpy(), g["yhat"]) < 1e-5
    assert abs(float(loss) - g["losses"][0]) < 1e-6 * g["losses"][0]
    for prefix, sd in (("enc.", m.enc), ("dec.", m.dec)):
        for k, v in sd.items():
            gk = "grad." + prefix + k
            if gk in g:
                scale = max(np.abs(g[gk]).max(), 1e-6)
                assert np.abs(v.grad.numpy() - g[gk]).max() <= 2e-4 * scale + 1e-8, gk


@pytest.mark.parametrize("name", ["mini", "nonsquare"])
def test_port_matches_reference_adam_steps(name):
    g = load_npz(f"layers_{name}.npz")
    m = OracleModel(split_sd(g, "init.enc."), split_sd(g, "init.dec."), spec_of(g))
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    losses = [float(m.train_step(x, y)) for _ in range(3)]
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-5)
    for prefix, sd in (("enc.", m.enc), ("dec.", m.dec)):
        for k, v in sd.items():
            ref = g["after3." + prefix + k]
            if ref.dtype.kind == "f":
                assert np.abs(v.detach().numpy() - ref).max() <= 1e-4 * max(np.abs(ref).max(), 1e-3), k
            else:
                assert int(v) == int(ref)
    assert rel_err(m.score(x).numpy(), g["eval_yhat"]) < 1e-4


def test_port_matches_reference_training_loop_ragged():
    """reference ConvAEModel.train, batch 64 on 100 samples (64 + ragged 36), 5 epochs"""
    g = load_npz("curve_conv_b64_e5.npz")
    tr, te = datagen.circle_datasets(100, 100)
    norm = lambda a, lo, hi: ((a - lo) / (hi - lo)).astype(np.float32)
    lo_min, lo_max = float(tr["lowres"].data.min()), float(tr["lowres"].data.max())
    hi_min, hi_max = float(tr["hires"].data.min()), float(tr["hires"].data.max())
    torch.manual_seed(1234)
    from cae_tools_b200.models.model_sizer import create_model_spec
    from cae_tools_b200.models.encoder import Encoder
    from cae_tools_b200.models.decoder import Decoder
    spec = create_model_spec(input_size=(16, 16), input_channels=1, output_size=(256, 256), output_channels=1)
    assert spec.save() == spec_of(g)
    enc = Encoder(spec.get_input_layers(), 4, 16)      # containers only supply the reference's init stream
    dec = Decoder(spec.get_output_layers(), 4, 16)
    order_tr = shuffled_order(100, 64)
    order_te = shuffled_order(100, 64)
    m = OracleModel(enc.state_dict(), dec.state_dict(), spec.save())
    btr = make_batches(norm(tr["lowres"].data, lo_min, lo_max), norm(tr["hires"].data, hi_min, hi_max), order_tr, 64)
    bte = make_batches(norm(te["lowres"].data, lo_min, lo_max), norm(te["hires"].data, hi_min, hi_max), order_te, 64)
    assert [b[0].shape[0] for b in btr] == [64, 36]
    train, test = [], []
    for _ in range(5):
        train.append(m.train_epoch(btr))
        test.append(m.test_epoch(bte))
    np.testing.assert_allclose(train, g["train_loss"], rtol=2e-4)
    np.testing.assert_allclose(test, g["test_loss"], rtol=2e-4)


def test_numpy_definitions_match_torch():
    rng = np.random.RandomState(0)
    for (ci, co, k, s, p, op, h, w) in [(3, 4, 3, 2, 0, 0, 7, 9), (2, 5, (4, 3), 2, 1, 1, 6, 5), (1, 2, 5, 3, 2, 0, 8, 8)]:
        kh, kw = (k, k) if isinstance(k, int) else k
        x = rng.randn(2, ci, h, w)
        wc = rng.randn(co, ci, kh, kw)
        b = rng.randn(co)
        ref = F.conv2d(torch.tensor(x), torch.tensor(wc), torch.tensor(b), stride=s, padding=p).numpy()
        np.testing.assert_allclose(numpy_ops.conv2d(x, wc, b, s, p), ref, atol=1e-10)
        wt = rng.randn(ci, co, kh, kw)
        ref = F.conv_transpose2d(torch.tensor(x), torch.tensor(wt), torch.tensor(b), stride=s, padding=p,
                                 output_padding=op).numpy()
        np.testing.assert_allclose(numpy_ops.conv_transpose2d(x, wt, b, s, p, op), ref, atol=1e-10)
    x = rng.randn(4, 3, 5, 6) * 2 + 1
    gam, bet, rm, rv = rng.rand(3) + 0.5, rng.randn(3), rng.randn(3), rng.rand(3) + 0.5
    trm, trv = torch.tensor(rm.copy()), torch.tensor(rv.copy())
    ref = F.batch_norm(torch.tensor(x), trm, trv, torch.tensor(gam), torch.tensor(bet), True, 0.1, 1e-5).numpy()
    y, nrm, nrv = numpy_ops.batch_norm_train(x, gam, bet, rm, rv)
    np.testing.assert_allclose(y, ref, atol=1e-10)
    np.testing.assert_allclose(nrm, trm.numpy(), atol=1e-12)
    np.testing.assert_allclose(nrv, trv.numpy(), atol=1e-12)
    ref = F.batch_norm(torch.tensor(x), trm, trv, torch.tensor(gam), torch.tensor(bet), False, 0.1, 1e-5).numpy()
    np.testing.assert_allclose(numpy_ops.batch_norm_eval(x, gam, bet, trm.numpy(), trv.numpy()), ref, atol=1e-10)
    # Adam / AdamW, three steps
    for decoupled in (False, True):
        p0 = rng.randn(50)
        tp = torch.tensor(p0.copy(), requires_grad=True)
        opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([tp], lr=1e-2, weight_decay=0.1)
        p, m, v = p0.copy(), np.zeros(50), np.zeros(50)
        for t in range(1, 4):
            gnp = rng.randn(50)
            tp.grad = torch.tensor(gnp.copy())
            opt.step()
            p, m, v = numpy_ops.adam_step(p, gnp, m, v, t, lr=1e-2, wd=0.1, decoupled=decoupled)
            np.testing.assert_allclose(p, tp.detach().numpy(), atol=1e-12)


@pytest.mark.parametrize("name", ["nomask", "mask"])
def test_unet_port_matches_reference(name):
    """oracle/torch_port.OracleUNet vs the reference's unet Encoder / Decoder / losses + AdamW (3 steps)"""
    from oracle.torch_port import OracleUNet, unet_forward
    g = load_npz(f"unet_{name}.npz")
    spec = spec_of(g)
    m = OracleUNet(split_sd(g, "init.enc."), split_sd(g, "init.dec."), spec, lambda_pearson=1.0)
    x, y, mask = (torch.from_numpy(g[k]) for k in ("x", "y", "mask"))
    trace = {"enc": [], "dec": []}
    yhat = unet_forward(m.enc, m.dec, spec, x, True, trace)
    for i, t in enumerate(trace["enc"]):
        assert rel_err(t.detach().numpy(), g[f"act.enc.{4 * i}"]) < 1e-5
    for j, t in enumerate(trace["dec"]):
        assert rel_err(t.detach().numpy(), g[f"act.dec.{4 * j}"]) < 1e-5
    assert rel_err(yhat.detach().numpy(), g["yhat"]) < 1e-5
    # fresh model (the trace pass above advanced the BatchNorm buffers)
    m = OracleUNet(split_sd(g, "init.enc."), split_sd(g, "init.dec."), spec, lambda_pearson=1.0)
    mses, pls = [], []
    for step in range(3):
        a, b = m.train_step(x, y, mask)
        mses.append(a)
        pls.append(b)
        if step == 0:
            pass
    np.testing.assert_allclose(mses, g["mse"], rtol=1e-5)
    np.testing.assert_allclose(pls, g["pearson_loss"], rtol=1e-5)
    for prefix, sd in (("enc.", m.enc), ("dec.", m.dec)):
        for k, v in sd.items():
            ref = g["after3." + prefix + k]
            if ref.dtype.kind == "f":
                assert np.abs(v.detach().numpy() - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-3), k
    assert rel_err(m.score(x).numpy(), g["eval_yhat"]) < 2e-4


@pytest.mark.parametrize("name", ["nomask", "mask", "head16_mask"])
def test_numpy_unet_definitions_match_reference_values(name):
    """oracle/numpy_ops restatements of the UNET pieces against numbers the REFERENCE itself produced (golden fixtures
    written by oracle/gen_golden.py from cae_tools.models.unet): the loss terms of step 0 from its yhat, the first
    decoder block (transposed conv -> ChannelAttention gate -> concat -> BatchNorm -> ReLU -> next transposed conv),
    the last layer + sigmoid."""
    from oracle import numpy_ops as npo
    g = load_npz(f"unet_{name}.npz")
    spec = spec_of(g)
    mask = g["mask"]
    assert abs(npo.masked_mse(g["yhat"], g["y"], mask) - g["mse"][0]) <= 1e-6 * g["mse"][0]
    pl = 1.0 - npo.pearson_corr(g["yhat"], g["y"], mask).mean()
    assert abs(pl - g["pearson_loss"][0]) <= 2e-6 * abs(g["pearson_loss"][0])
    # encoder skip activations: relu(bn_train(conv)) of the golden raw conv outputs
    enc = split_sd(g, "init.enc.")
    dec = split_sd(g, "init.dec.")
    skips = []
    for i in range(len(spec["input_layers"])):
        raw = g[f"act.enc.{4 * i}"]
        a, _, _ = npo.batch_norm_train(raw, enc[f"encoder_cnn.{4 * i + 1}.weight"].numpy(), enc[f"encoder_cnn.{4 * i + 1}.bias"].numpy(),
                                       np.zeros(raw.shape[1]), np.ones(raw.shape[1]))
        skips.append(np.maximum(a, 0.0))
    # decoder block 0: gate the golden raw output, concat the deepest-but-one skip, BN, ReLU, then the next transposed conv
    y0 = g["act.dec.0"]
    att = npo.channel_attention(y0, dec["attention_layers.0.fc1.weight"].numpy(), dec["attention_layers.0.fc2.weight"].numpy())
    cat = np.concatenate([y0 * att, skips[-2]], axis=1)
    a0, _, _ = npo.batch_norm_train(cat, dec["decoder_conv.1.weight"].numpy(), dec["decoder_conv.1.bias"].numpy(),
                                    np.zeros(cat.shape[1]), np.ones(cat.shape[1]))
    sp1 = spec["output_layers"][1]
    y1 = npo.conv_transpose2d(np.maximum(a0, 0.0), dec["decoder_conv.4.weight"].numpy(), dec["decoder_conv.4.bias"].numpy(),
                              sp1["stride"], pad=sp1["output_padding"])
    assert rel_err(y1, g["act.dec.4"]) < 1e-5
    # last layer: the golden raw output of the last transposed conv through the sigmoid is the golden prediction
    last = 4 * (len(spec["output_layers"]) - 1)
    assert rel_err(npo.sigmoid(g[f"act.dec.{last}"]), g["yhat"]) < 1e-6


def test_numpy_linear_and_adamw_match_torch():
    from oracle import numpy_ops as npo
    rng = np.random.RandomState(3)
    x, w, b = rng.randn(5, 7), rng.randn(4, 7), rng.randn(4)
    ref = torch.nn.functional.linear(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b)).numpy()
    np.testing.assert_allclose(npo.linear(x, w, b), ref, atol=1e-12)


def test_unet_port_matches_reference_k32_fixture():
    """the shipped 16x16 -> 256x256 spec (k32 s32 head) at reference-produced values: losses of 3 AdamW steps, gradients of
    step 0, eval prediction at the initial and the trained weights (fixture stores the 256x256 tensors subsampled by 8)"""
    from helpers import unet_light_data
    from oracle.torch_port import OracleUNet
    g = load_npz("unet_head32_mask_light.npz")
    spec, L = spec_of(g), int(g["light"])
    x, y, mask = unet_light_data(g)
    m = OracleUNet(split_sd(g, "init.enc."), split_sd(g, "init.dec."), spec, lambda_pearson=1.0)
    assert rel_err(m.score(x).numpy()[:, :, ::L, ::L], g["eval_yhat_init"]) < 1e-5
    mses, pls = [], []
    for step in range(3):
        a, b = m.train_step(x, y, mask)
        mses.append(a)
        pls.append(b)
        if step == 0:
            for prefix, sd in (("enc.", m.enc), ("dec.", m.dec)):
                for k, v in sd.items():
                    gk = "grad." + prefix + k
                    if gk in g:
                        scale = max(np.abs(g[gk]).max(), 1e-7)
                        assert np.abs(v.grad.numpy() - g[gk]).max() <= 2e-4 * scale + 1e-9, gk
    np.testing.assert_allclose(mses, g["mse"], rtol=1e-5)
    np.testing.assert_allclose(pls, g["pearson_loss"], rtol=1e-5)
    assert rel_err(m.score(x).numpy()[:, :, ::L, ::L], g["eval_yhat"]) < 2e-4


def test_unet_port_reproduces_reference_50_epoch_curve():
    """BASELINE configs[1]: the port's loop against the reference's UNET training loop on the circle data, batch 64 (64 + 36),
    50 epochs (fixture curve_unet_b64_e50.npz; init = same seed and module tree, shuffles = same RNG stream)"""
    import json
    import os
    from cae_tools_b200.models.model_sizer import ModelSpec
    from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
    from oracle.torch_port import OracleUNet
    g = load_npz("curve_unet_b64_e50.npz")
    spec_json = spec_of(g)
    spec = ModelSpec()
    spec.load(spec_json)
    tr, te = datagen.circle_datasets(100, 100)
    lo_min, lo_max = float(tr["lowres"].data.min()), float(tr["lowres"].data.max())
    hi_min, hi_max = float(tr["hires"].data.min()), float(tr["hires"].data.max())
    norm = lambda a, lo, hi: ((a - lo) / (hi - lo)).astype(np.float32)
    torch.manual_seed(1234)
    enc, dec = UNetEncoder(spec.get_input_layers(), 4, 16, 0.0), UNetDecoder(spec.get_output_layers(), 4, 16, 0.0)
    otr, ote = shuffled_order(100, 64), shuffled_order(100, 64)
    m = OracleUNet(enc.state_dict(), dec.state_dict(), spec_json, lambda_pearson=1.0)
    btr = make_batches(norm(tr["lowres"].data, lo_min, lo_max), norm(tr["hires"].data, hi_min, hi_max), otr, 64)
    bte = make_batches(norm(te["lowres"].data, lo_min, lo_max), norm(te["hires"].data, hi_min, hi_max), ote, 64)
    tl, el = [], []
    for epoch in range(12):                    # 12 of the 50 epochs keep the CPU suite short
        tl.append(float(np.mean([m.train_step(x, y, torch.ones_like(y))[0] for x, y in btr])))
        with torch.no_grad():
            el.append(float(np.mean([float(m.losses(x, y, torch.ones_like(y), False)[0]) for x, y in bte])))
    np.testing.assert_allclose(tl, g["train_loss"][:12], rtol=2e-5)
    np.testing.assert_allclose(el, g["test_loss"][:12], rtol=2e-5)


def test_linear_port_matches_reference():
    """oracle/torch_port.OracleLinear pinned to the reference's Linear module + MSELoss + Adam (fixture linear_mini.npz)"""
    from cae_tools_b200.models.linear import Linear
    from oracle.torch_port import OracleLinear
    g = load_npz("linear_mini.npz")
    torch.manual_seed(int(g["seed"]))
    mod = Linear((1, 16, 16), (1, 64, 64))
    assert np.array_equal(mod.linear[1].weight.detach().numpy()[::16], g["init.weight_sub"])     # same init stream
    assert np.array_equal(mod.linear[1].bias.detach().numpy(), g["init.bias"])
    m = OracleLinear(mod.state_dict(), (1, 64, 64), lr=1e-3, weight_decay=1e-5)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    losses = []
    for step in range(3):
        losses.append(m.train_step(x, y))
        if step == 0:
            assert rel_err(m.w.grad.numpy()[::16], g["grad.weight_sub"]) < 1e-5
            assert rel_err(m.b.grad.numpy(), g["grad.bias"]) < 1e-5
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-6)
    assert rel_err(m.w.detach().numpy()[::16], g["after.weight_sub"]) < 1e-5
    assert rel_err(m.score(x).numpy(), g["pred"]) < 1e-5
