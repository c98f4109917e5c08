"""TEST INFRASTRUCTURE - writes tests/golden/* by running the UNMODIFIED reference in the build container.

    python oracle/gen_golden.py            (needs /root/reference; not runnable on the GPU box)

The reference's modules are imported from /root/reference/src (with the file-based xarray stand-in of
oracle/_shims on sys.path because the image has no xarray).  Two run-time patches are applied to the
*loaded module objects* - no reference source is copied or edited:
  * DSDataset.__getitem__ is wrapped to return the 3-tuple (input, output, label) that
    ConvAEModel.train unpacks (reference snapshot inconsistency: ds_dataset.py:159 returns 4 values,
    conv_ae_model.py:316,322 unpack 3);
  * BaseModel.evaluate / dump_metrics are stubbed (reference bug: default mask shaped like the INPUT,
    base_model.py:93-98 - IndexError whenever input and output sizes differ).
Everything written here is small (<1 MB per file) and committed.
"""

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("CAE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, ROOT)

from oracle import datagen  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_modules():
    from cae_tools.models import encoder, decoder, model_sizer
    return encoder, decoder, model_sizer


def sd_np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def gen_specs():
    _, _, sizer = ref_modules()
    cases = {
        "circle_16x16_256x256": dict(input_size=(16, 16), input_channels=1, output_size=(256, 256), output_channels=1),
        "tidal_6x6_256x256": dict(input_size=(6, 6), input_channels=2, output_size=(256, 256), output_channels=1),
        "circle2_24x20_280x256": dict(input_size=(24, 20), input_channels=1, output_size=(280, 256), output_channels=1),
        "config4_64x64_1024x1024": dict(input_size=(64, 64), input_channels=4, output_size=(1024, 1024),
                                        output_channels=4),
        "mini_16x16_64x64": dict(input_size=(16, 16), input_channels=1, output_size=(64, 64), output_channels=1),
        "nonsquare_12x10_40x36": dict(input_size=(12, 10), input_channels=1, output_size=(40, 36), output_channels=1),
        "layers_2_3": dict(input_size=(32, 32), input_channels=3, output_size=(128, 128), output_channels=2,
                           input_layer_count=2, output_layer_count=3),
    }
    out = {}
    for name, kw in cases.items():
        spec = sizer.create_model_spec(kernel_size=3, stride=2, **kw)
        out[name] = {"args": {k: list(v) if isinstance(v, tuple) else v for k, v in kw.items()}, "spec": spec.save()}
    with open(os.path.join(GOLD, "specs.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("specs.json", len(out))


def gen_layers(name, in_size, out_size, in_ch, out_ch, batch, latent=4, fc=16, seed=1234, steps=3):
    """reference Encoder/Decoder + MSELoss + Adam: per-layer activations, grads, 3 optimiser steps"""
    enc_m, dec_m, sizer = ref_modules()
    torch.manual_seed(seed)
    spec = sizer.create_model_spec(input_size=in_size, input_channels=in_ch, output_size=out_size,
                                   output_channels=out_ch, kernel_size=3, stride=2)
    enc = enc_m.Encoder(spec.get_input_layers(), encoded_space_dim=latent, fc_size=fc)
    dec = dec_m.Decoder(spec.get_output_layers(), encoded_space_dim=latent, fc_size=fc)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(batch, in_ch, *in_size, generator=g)
    y = torch.rand(batch, out_ch, *out_size, generator=g)
    out = {"x": x.numpy(), "y": y.numpy()}
    out.update(sd_np(enc.state_dict(), "init.enc."))
    out.update(sd_np(dec.state_dict(), "init.dec."))

    # per-layer raw conv outputs (hooks on the conv modules; clone - the following ReLU is in-place)
    acts = {}
    hooks = []
    for prefix, seq in (("enc", enc.encoder_cnn), ("dec", dec.decoder_conv)):
        for idx, mod in enumerate(seq):
            if isinstance(mod, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                hooks.append(mod.register_forward_hook(
                    lambda m, i, o, key=f"act.{prefix}.{idx}": acts.__setitem__(key, o.detach().clone().numpy())))
    loss_fn = torch.nn.MSELoss()
    optim = torch.optim.Adam([{'params': enc.parameters()}, {'params': dec.parameters()}], lr=1e-3, weight_decay=1e-5)
    enc.train(); dec.train()
    losses = []
    for step in range(steps):
        z = enc(x)
        yhat = dec(z)
        loss = loss_fn(yhat, y)
        optim.zero_grad()
        loss.backward()
        if step == 0:
            out.update({k: (v[:, :, ::light, ::light] if light and v.shape[-1] >= 128 else v) for k, v in acts.items()})
            out["z"] = z.detach().numpy().copy()
            yh = yhat.detach().numpy().copy()
            out["yhat"] = yh[:, :, ::light, ::light] if light else yh
            for k, p in enc.named_parameters():
                out["grad.enc." + k] = p.grad.detach().numpy().copy()
            for k, p in dec.named_parameters():
                out["grad.dec." + k] = p.grad.detach().numpy().copy()
        optim.step()
        losses.append(float(loss.detach()))
    for h in hooks:
        h.remove()
    out["losses"] = np.array(losses, dtype=np.float64)
    out.update(sd_np(enc.state_dict(), f"after{steps}.enc."))
    out.update(sd_np(dec.state_dict(), f"after{steps}.dec."))
    enc.eval(); dec.eval()
    with torch.no_grad():
        out["eval_yhat"] = dec(enc(x)).numpy().copy()
    out["spec_json"] = np.array(json.dumps(spec.save()))
    path = os.path.join(GOLD, f"layers_{name}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KB", "losses", losses)


def gen_loss_curve(name, batch_size, nr_epochs, seed=1234, n=100):
    """the reference ConvAEModel.train itself (patched as described in the module docstring)"""
    from cae_tools.models import conv_ae_model as cam, ds_dataset, base_model
    orig_getitem = ds_dataset.DSDataset.__getitem__
    try:
        def getitem3(self, index):
            a, b, _m, lbl = orig_getitem(self, index)
            return a, b, lbl
        ds_dataset.DSDataset.__getitem__ = getitem3
        ds_dataset.DSDataset.set_normalisation_parameters = \
            lambda self, p: [setattr(self, k, v) for k, v in zip(("min_inputs", "max_inputs", "min_output", "max_output"), p)]
        base_model.BaseModel.evaluate = lambda self, dataset, device: {}
        base_model.BaseModel.dump_metrics = lambda self, title, metrics: None
        train_ds, test_ds = datagen.circle_datasets(n, n)
        torch.manual_seed(seed)
        m = cam.ConvAEModel(batch_size=batch_size, nr_epochs=nr_epochs, test_interval=1, encoded_dim_size=4,
                            fc_size=16, lr=1e-3, weight_decay=1e-5, use_gpu=False)
        m.train(["lowres"], "hires", train_ds, test_ds)
        # predictions of the trained reference on the first 4 test samples (eval mode)
        lo = np.asarray(test_ds["lowres"].data[:4])
        mn, mx = m.normalisation_parameters[0]["lowres"], m.normalisation_parameters[1]["lowres"]
        xin = torch.from_numpy(((lo - mn) / (mx - mn)).astype(np.float32))
        m.encoder.eval(); m.decoder.eval()
        with torch.no_grad():
            pred = m.decoder(m.encoder(xin)).numpy()
        out = {
            "train_loss": np.array(m.history["train_loss"], dtype=np.float64),
            "test_loss": np.array(m.history["test_loss"], dtype=np.float64),
            "pred_sub": pred[:, :, ::8, ::8].copy(),
            "pred_mean": pred.mean(axis=(1, 2, 3)),
            "norm": np.array(json.dumps(m.normalisation_parameters)),
            "spec_json": np.array(json.dumps(m.spec.save())),
            "params_json": np.array(json.dumps({k: v for k, v in m.get_parameters().items() if k != "model_id"})),
        }
        out.update(sd_np(m.encoder.state_dict(), "final.enc."))
        out.update(sd_np(m.decoder.state_dict(), "final.dec."))
        path = os.path.join(GOLD, f"curve_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KB", out["train_loss"][[0, -1]], out["test_loss"][[0, -1]])
    finally:
        ds_dataset.DSDataset.__getitem__ = orig_getitem


UNET_SPEC = {
    "input_layers": [
        {"is_input": True, "kernel_size": 3, "stride": 2, "output_padding": 1, "input_dimensions": [1, 16, 16], "output_dimensions": [8, 8, 8]},
        {"is_input": True, "kernel_size": 3, "stride": 2, "output_padding": 1, "input_dimensions": [8, 8, 8], "output_dimensions": [16, 4, 4]},
        {"is_input": True, "kernel_size": 3, "stride": 2, "output_padding": 1, "input_dimensions": [16, 4, 4], "output_dimensions": [32, 2, 2]}],
    "output_layers": [
        {"is_input": False, "kernel_size": 4, "stride": 2, "output_padding": 1, "input_dimensions": [32, 2, 2], "output_dimensions": [16, 4, 4]},
        {"is_input": False, "kernel_size": 4, "stride": 2, "output_padding": 1, "input_dimensions": [32, 4, 4], "output_dimensions": [8, 8, 8]},
        {"is_input": False, "kernel_size": 8, "stride": 8, "output_padding": 0, "input_dimensions": [16, 8, 8], "output_dimensions": [1, 64, 64]}],
}


def unet_spec_with_head(k):
    """UNET_SPEC with the last transposed conv replaced by a kernel == stride == k head (16x8x8 -> 1 x 8k x 8k);
    k = 32 is the shipped 16x16 -> 256x256 spec (cae_tools_b200/specs/unet_16x16_256x256.json)"""
    spec = json.loads(json.dumps(UNET_SPEC))
    spec["output_layers"][-1].update(kernel_size=k, stride=k, output_dimensions=[1, 8 * k, 8 * k])
    return spec


def gen_unet(name, with_mask, batch=6, latent=8, fc=32, seed=4321, steps=3, lambda_pearson=1.0, spec_dict=None, light=0):
    # light > 0: the 256x256 fixtures of the shipped k32 spec - y / mask are regenerated by the test from `data_seed`
    # (torch's CPU generator is platform-independent) and the full-resolution tensors are stored subsampled by `light`
    """the reference's unet Encoder / Decoder / masked_mse_loss / pearson_corr_torch + AdamW, dropout 0
    (UNET() itself cannot be constructed offline: its constructor downloads VGG weights - SURVEY section 0)"""
    from cae_tools.models import unet as ru
    from cae_tools.models.model_sizer import ModelSpec
    spec_dict = spec_dict or UNET_SPEC
    oh, ow = spec_dict["output_layers"][-1]["output_dimensions"][1:]
    spec = ModelSpec()
    spec.load(spec_dict)
    torch.manual_seed(seed)
    enc = ru.Encoder(spec.get_input_layers(), encoded_space_dim=latent, fc_size=fc, dropout_rate=0.0)
    dec = ru.Decoder(spec.get_output_layers(), encoded_space_dim=latent, fc_size=fc, dropout_rate=0.0)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(batch, 1, 16, 16, generator=g)
    y = torch.rand(batch, 1, oh, ow, generator=g)
    mask = (torch.rand(batch, 1, oh, ow, generator=g) > 0.3).float() if with_mask else torch.ones(batch, 1, oh, ow)
    out = {"x": x.numpy(), "data_seed": np.array(seed + 1), "batch": np.array(batch), "with_mask": np.array(int(with_mask))}
    if not light:
        out.update({"y": y.numpy(), "mask": mask.numpy()})
    out.update(sd_np(enc.state_dict(), "init.enc."))
    out.update(sd_np(dec.state_dict(), "init.dec."))
    enc.eval(); dec.eval()
    with torch.no_grad():                       # eval-mode prediction with the INITIAL weights / BatchNorm buffers
        z0, skip0 = enc(x)
        e0 = dec(z0, skip0).numpy().copy()
        out["eval_yhat_init"] = e0[:, :, ::light, ::light] if light else e0
    acts, hooks = {}, []
    for prefix, seq in (("enc", enc.encoder_cnn), ("dec", dec.decoder_conv)):
        for idx, mod in enumerate(seq):
            if isinstance(mod, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                hooks.append(mod.register_forward_hook(
                    lambda m, i, o, key=f"act.{prefix}.{idx}": acts.__setitem__(key, o.detach().clone().numpy())))
    optim = torch.optim.AdamW(list(enc.parameters()) + list(dec.parameters()), lr=1e-3, weight_decay=1e-5)
    enc.train(); dec.train()
    mses, pls = [], []
    for step in range(steps):
        optim.zero_grad()
        z, skip = enc(x)
        yhat = dec(z, skip)
        mse = ru.UNET.masked_mse_loss(None, yhat, y, mask)
        pl = 1 - torch.mean(ru.UNET.pearson_corr_torch(None, yhat, y, mask))
        (mse + lambda_pearson * pl).backward()
        if step == 0:
            out.update({k: (v[:, :, ::light, ::light] if light and v.shape[-1] >= 128 else v) for k, v in acts.items()})
            out["z"] = z.detach().numpy().copy()
            yh = yhat.detach().numpy().copy()
            out["yhat"] = yh[:, :, ::light, ::light] if light else yh
            for k, p in enc.named_parameters():
                out["grad.enc." + k] = p.grad.detach().numpy().copy()
            for k, p in dec.named_parameters():
                out["grad.dec." + k] = p.grad.detach().numpy().copy()
        optim.step()
        mses.append(float(mse.detach())); pls.append(float(pl.detach()))
    for h in hooks:
        h.remove()
    out["mse"] = np.array(mses); out["pearson_loss"] = np.array(pls)
    out.update(sd_np(enc.state_dict(), f"after{steps}.enc."))
    out.update(sd_np(dec.state_dict(), f"after{steps}.dec."))
    enc.eval(); dec.eval()
    with torch.no_grad():
        z, skip = enc(x)
        ey = dec(z, skip).numpy().copy()
        out["eval_yhat"] = ey[:, :, ::light, ::light] if light else ey
    out["light"] = np.array(light)
    out["spec_json"] = np.array(json.dumps(spec_dict))
    path = os.path.join(GOLD, f"unet_{name}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KB", "mse", mses, "pearson", pls)


def gen_unet_curve(name="unet_b64_e50", batch_size=64, nr_epochs=50, seed=1234, n=100, latent=4, fc=16, lambda_pearson=1.0):
    """BASELINE configs[1]: the reference's UNET training loop (unet.py:388-529 setup, :295-337 train epoch, :339-372 test
    epoch) on the seeded circle data with the shipped 16x16 -> 256x256 spec, dropout 0, batch 64 (64 + 36), 50 epochs,
    test_interval 1.  UNET() itself cannot be constructed offline (VGG download in its constructor), so the loop is run on
    the reference's own unet.Encoder / unet.Decoder / masked_mse_loss / pearson_corr_torch / AdamW with the reference's
    DSDataset and torch DataLoader(shuffle=True) - module construction first, then the two loaders, like UNET.train."""
    from cae_tools.models import unet as ru
    from cae_tools.models.ds_dataset import DSDataset
    from cae_tools.models.model_sizer import ModelSpec
    spec_dict = unet_spec_with_head(32)
    spec = ModelSpec()
    spec.load(spec_dict)
    tr, te = datagen.circle_datasets(n, n)
    train_ds = DSDataset(tr, ["lowres"], "hires", normalise_in=True, normalise_out=True)
    test_ds = DSDataset(te, ["lowres"], "hires", normalise_in=True, normalise_out=True)
    test_ds.set_normalisation_parameters(train_ds.get_normalisation_parameters())
    torch.manual_seed(seed)
    enc = ru.Encoder(spec.get_input_layers(), encoded_space_dim=latent, fc_size=fc, dropout_rate=0.0)
    dec = ru.Decoder(spec.get_output_layers(), encoded_space_dim=latent, fc_size=fc, dropout_rate=0.0)
    train_loader = torch.utils.data.DataLoader(train_ds, batch_size=batch_size, shuffle=True)
    test_loader = torch.utils.data.DataLoader(test_ds, batch_size=batch_size, shuffle=True)
    optim = torch.optim.AdamW(list(enc.parameters()) + list(dec.parameters()), lr=1e-3, weight_decay=1e-5)
    # the reference's default mask has the INPUT's shape (ds_dataset.py:157) and cannot broadcast against the output: the
    # all-ones mask of the OUTPUT's shape is what "no mask" means (documented deviation, cae_tools_b200/models/unet.py)
    fix = lambda b: (b[0], b[1], torch.ones_like(b[1]))
    train_batches = [fix(b) for b in train_loader]
    test_batches = [fix(b) for b in test_loader]
    hist = {"train_loss": [], "test_loss": [], "train_pearson": [], "test_pearson": []}
    for epoch in range(nr_epochs):
        enc.train(); dec.train()
        tl, tp = [], []
        for x, y, m in train_batches:
            optim.zero_grad()
            z, skip = enc(x)
            yhat = dec(z, skip)
            mse = ru.UNET.masked_mse_loss(None, yhat, y, m)
            pl = 1 - torch.mean(ru.UNET.pearson_corr_torch(None, yhat, y, m))
            (mse + lambda_pearson * pl).backward()
            optim.step()
            tl.append(mse.item()); tp.append(pl.item())
        enc.eval(); dec.eval()
        el, ep = [], []
        with torch.no_grad():
            for x, y, m in test_batches:
                z, skip = enc(x)
                yhat = dec(z, skip)
                el.append(ru.UNET.masked_mse_loss(None, yhat, y, m).numpy())
                ep.append((1 - torch.mean(ru.UNET.pearson_corr_torch(None, yhat, y, m))).numpy())
        hist["train_loss"].append(float(np.mean(tl))); hist["train_pearson"].append(float(np.mean(tp)))
        hist["test_loss"].append(float(np.mean(el))); hist["test_pearson"].append(float(np.mean(ep)))
    out = {k: np.array(v) for k, v in hist.items()}
    enc.eval(); dec.eval()
    with torch.no_grad():
        x4 = torch.stack([torch.as_tensor(np.asarray(test_ds[i][0])) for i in range(4)])
        z, skip = enc(x4)
        pred = dec(z, skip).numpy()
    out["pred_sub"] = pred[:, :, ::8, ::8].copy()
    out.update(sd_np(enc.state_dict(), "final.enc."))
    out.update(sd_np(dec.state_dict(), "final.dec."))
    out["spec_json"] = np.array(json.dumps(spec_dict))
    path = os.path.join(GOLD, f"curve_{name}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KB", "train", hist["train_loss"][:2], hist["train_loss"][-1], "test", hist["test_loss"][-1])


def gen_linear(name="linear_mini", seed=77, steps=3):
    """the reference's Linear module (linear.py:33-49) + MSELoss + Adam(lr, weight_decay) (linear_model.py:236-247): losses,
    gradients of step 0, weights after `steps`, prediction - pins oracle/torch_port.OracleLinear and the CUDA engine"""
    from cae_tools.models.linear import Linear
    torch.manual_seed(seed)
    mod = Linear((1, 16, 16), (1, 64, 64))
    g = torch.Generator().manual_seed(seed + 1)
    x, y = torch.rand(8, 1, 16, 16, generator=g), torch.rand(8, 1, 64, 64, generator=g)
    # (the 4 MB initial weight is not stored: torch.manual_seed(seed) + the same module tree regenerates it; 256 of its
    #  4096 rows are kept to verify that)
    out = {"x": x.numpy(), "y": y.numpy(), "seed": np.array(seed)}
    out["init.weight_sub"] = mod.linear[1].weight.detach().numpy()[::16].copy()
    out["init.bias"] = mod.linear[1].bias.detach().numpy().copy()
    optim = torch.optim.Adam([{"params": mod.parameters()}], lr=1e-3, weight_decay=1e-5)
    loss_fn = torch.nn.MSELoss()
    losses = []
    for step in range(steps):
        loss = loss_fn(mod(x), y)
        optim.zero_grad()
        loss.backward()
        if step == 0:
            w = mod.linear[1].weight.grad.numpy()
            out["grad.weight_sub"] = w[::16].copy()               # 256 of the 4096 rows
            out["grad.bias"] = mod.linear[1].bias.grad.numpy().copy()
        optim.step()
        losses.append(float(loss.detach()))
    out["losses"] = np.array(losses)
    out["after.weight_sub"] = mod.linear[1].weight.detach().numpy()[::16].copy()
    out["after.bias"] = mod.linear[1].bias.detach().numpy().copy()
    with torch.no_grad():
        out["pred"] = mod(x).numpy().copy()
    path = os.path.join(GOLD, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KB", losses)


def gen_chaos_envelope(seed=1234, nr_epochs=50, batch_size=10):
    """How far do the REFERENCE's own 50-epoch loss curves move when only the CPU thread count changes?
    (Adam amplifies fp32 summation-order differences; SURVEY section 7 'loss-curve chaos'.)  Uses the oracle port,
    which reproduces the reference's 8-thread curve bit for bit (asserted below), at 1 and 4 threads."""
    from oracle.torch_port import OracleModel, make_batches, shuffled_order
    from cae_tools.models import encoder, decoder, model_sizer
    ref = dict(np.load(os.path.join(GOLD, "curve_conv_b10_e50.npz")))
    tr, te = datagen.circle_datasets(100, 100)
    norm = lambda a, lo, hi: ((a - lo) / (hi - lo)).astype(np.float32)
    lo_min, lo_max = float(tr["lowres"].data.min()), float(tr["lowres"].data.max())
    hi_min, hi_max = float(tr["hires"].data.min()), float(tr["hires"].data.max())
    out = {}
    for threads in (8, 1, 4):
        torch.set_num_threads(threads)
        torch.manual_seed(seed)
        spec = model_sizer.create_model_spec(input_size=(16, 16), input_channels=1, output_size=(256, 256),
                                             output_channels=1)
        enc = encoder.Encoder(spec.get_input_layers(), 4, 16)
        dec = decoder.Decoder(spec.get_output_layers(), 4, 16)
        otr, ote = shuffled_order(100, batch_size), shuffled_order(100, batch_size)
        m = OracleModel(enc.state_dict(), dec.state_dict(), spec.save())
        btr = make_batches(norm(tr["lowres"].data, lo_min, lo_max), norm(tr["hires"].data, hi_min, hi_max), otr, batch_size)
        bte = make_batches(norm(te["lowres"].data, lo_min, lo_max), norm(te["hires"].data, hi_min, hi_max), ote, batch_size)
        a, b = [], []
        for _ in range(nr_epochs):
            a.append(m.train_epoch(btr))
            b.append(m.test_epoch(bte))
        a, b = np.array(a), np.array(b)
        if threads == 8:
            assert np.array_equal(a, ref["train_loss"]) and np.array_equal(b, ref["test_loss"]), \
                "port no longer reproduces the reference bit for bit at 8 threads"
        else:
            out[f"train_t{threads}"] = a
            out[f"test_t{threads}"] = b
            print(f"threads={threads}: max rel dev train {np.max(np.abs(a - ref['train_loss']) / ref['train_loss']):.2e}"
                  f" test {np.max(np.abs(b - ref['test_loss']) / ref['test_loss']):.2e}")
    torch.set_num_threads(8)
    np.savez_compressed(os.path.join(GOLD, "curve_conv_b10_e50_envelope.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        # fixtures added in round 2 (the round-1 files are reproduced bit for bit by the full run below)
        gen_unet_curve()
        gen_unet("head32_mask_light", with_mask=True, batch=4, latent=4, fc=16, spec_dict=unet_spec_with_head(32), light=8)
        gen_unet("head16_mask", with_mask=True, batch=3, spec_dict=unet_spec_with_head(16))
        gen_linear()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "unet_head":
        # the fused kernel == stride head (patch_head.cu): k16 -> 128x128 output, masked, batch 3
        gen_unet("head16_mask", with_mask=True, batch=3, spec_dict=unet_spec_with_head(16))
        sys.exit(0)
    gen_specs()
    gen_layers("mini", (16, 16), (64, 64), 1, 1, batch=6)
    gen_layers("nonsquare", (12, 10), (40, 36), 1, 1, batch=5)
    gen_layers("multich", (16, 16), (64, 64), 2, 3, batch=4, latent=6, fc=24)
    gen_loss_curve("conv_b10_e50", batch_size=10, nr_epochs=50)
    gen_loss_curve("conv_b64_e5", batch_size=64, nr_epochs=5)
    gen_chaos_envelope()
    gen_unet("nomask", with_mask=False)
    gen_unet("mask", with_mask=True)
    gen_unet("head16_mask", with_mask=True, batch=3, spec_dict=unet_spec_with_head(16))
    gen_unet_curve()
    gen_unet("head32_mask_light", with_mask=True, batch=4, latent=4, fc=16, spec_dict=unet_spec_with_head(32), light=8)
    gen_linear()
