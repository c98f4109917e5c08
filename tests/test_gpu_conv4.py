"""BASELINE configs[3] (4x64x64 -> 4x1024x1024 ConvAE) and the tensor-core layer path inside the engine.

* the reference-produced fixture layers_multich.npz (48 -> 24 first decoder layer) with the tcgen05 path forced on:
  gradients / losses held to the same 1e-4 bar as the SIMT path;
* the config-4 geometry itself (spec = golden config4_64x64_1024x1024, from the live reference's create_model_spec) at a
  small batch against the oracle port: loss and every gradient after one step, parameters after two."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_npz, pre_bn_bias_keys, spec_of, split_sd

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_tc(monkeypatch):
    from cae_tools_b200.engine.convae import ConvAEEngine
    monkeypatch.setattr(ConvAEEngine, "TC_MIN_CIN", 32)
    monkeypatch.setattr(ConvAEEngine, "TC_MIN_FLOPS", 0.0)


def test_tc_layer_vs_reference_fixture(force_tc):
    from test_gpu_model import _build
    from cae_tools_b200.engine.convae import ConvAEEngine
    g = load_npz("layers_multich.npz")
    spec, enc, dec = _build(g)
    eng = ConvAEEngine(enc, dec, lr=1e-3, weight_decay=1e-5)
    x, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    data = eng.bind(x, y, x.shape[0])
    losses = []
    for step in range(3):
        losses.append(float(eng.train_epoch(data).cpu()[0]))
        if step == 0:
            names = [n for n, _ in eng._program("train", data, x.shape[0]).sched]
            assert "fwd.convT0.tc" in names and "bwd.convT0.tc.wgrad_gemm" in names, names
            for prefix, mod in (("enc.", enc), ("dec.", dec)):
                zero_bias = set(pre_bn_bias_keys(list(mod.state_dict().keys()), prefix))
                for k, p in mod.named_parameters():
                    ref = g["grad." + prefix + k]
                    got = p.grad.detach().cpu().numpy()
                    if k in zero_bias:
                        assert np.abs(got).max() <= 1e-6, k
                        continue
                    scale = max(np.abs(ref).max(), 1e-7)
                    assert np.abs(got - ref).max() <= 1e-4 * scale + 1e-9, (k, np.abs(got - ref).max(), scale)
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-5)


@pytest.mark.parametrize("tc", [True, False])
def test_config4_geometry_vs_oracle(tc, monkeypatch):
    from cae_tools_b200.engine.convae import ConvAEEngine
    from cae_tools_b200.models.decoder import Decoder
    from cae_tools_b200.models.encoder import Encoder
    from cae_tools_b200.models.model_sizer import ModelSpec, create_model_spec
    from oracle.torch_port import OracleModel
    if tc:
        monkeypatch.setattr(ConvAEEngine, "TC_MIN_FLOPS", 0.0)      # batch 2: below the production FLOP threshold
    else:
        monkeypatch.setattr(ConvAEEngine, "TC_MIN_CIN", 1 << 30)
    spec = create_model_spec(input_size=(64, 64), input_channels=4, output_size=(1024, 1024), output_channels=4)
    golden = json.load(open(os.path.join(GOLDEN, "specs.json")))["config4_64x64_1024x1024"]
    assert spec.save() == golden["spec"]
    torch.manual_seed(3)
    enc, dec = Encoder(spec.get_input_layers(), 4, 16), Decoder(spec.get_output_layers(), 4, 16)
    oracle = OracleModel(enc.state_dict(), dec.state_dict(), spec.save(), zero_dead_bias_grads=True)
    exact = OracleModel(enc.state_dict(), dec.state_dict(), spec.save(), zero_dead_bias_grads=True, dtype=torch.float64)
    B = 2
    x, y = torch.rand(B, 4, 64, 64), torch.rand(B, 4, 1024, 1024)
    eng = ConvAEEngine(enc, dec, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(x, y, B)
    names = [n for n, _ in eng._program("train", data, B).sched]
    assert ("fwd.convT0.tc" in names) == tc and ("fwd.convT4.tc" in names) == tc
    for step in range(2):
        got = float(eng.train_epoch(data).cpu()[0])
        want = float(oracle.train_step(x, y))
        assert abs(got - want) <= 2e-5 * want, (step, got, want)
        if step == 0:
            # Bar: 1e-4 of the max-norm against the fp32 oracle.  With batch-2 BatchNorm over 3x3 ... 511x511 planes many
            # gradients of this 17-layer chain are small differences of large terms (max |g| ~ 1e-5 ... 1e-3) and two fp32
            # evaluations of the same step differ by more than that - the fp32 oracle itself is 1e-3 away from the float64
            # evaluation on decoder_conv.0.weight.  Tensors beyond the bar are therefore adjudicated by the float64
            # evaluation in the L2 norm (single elements are pure noise): the CUDA gradient must be as close to it as the
            # fp32 oracle is, within 4x - 8x with the tensor-core layers, whose fp32 accumulation truncates instead of rounding
            # (measured worst case: the BatchNorm gamma behind the 64 -> 32 layer, a cancelling sum of dz * xhat, 6x) - or within
            # 1e-4 of the tensor's norm.  Per-layer accuracy of the tensor-core kernels themselves: tests/test_gpu_tc_conv.py
            # (2e-5) and test_tc_layer_vs_reference_fixture above (1e-4 against the reference's own gradients).
            exact.train_step(x.double(), y.double())
            worst = {}
            for sd, sd64, mod in ((oracle.enc, exact.enc, enc), (oracle.dec, exact.dec, dec)):
                for k, p in mod.named_parameters():
                    ref = sd[k].grad.numpy()
                    ref64 = sd64[k].grad.numpy()
                    gotg = p.grad.detach().cpu().numpy()
                    scale = max(np.abs(ref).max(), 1e-7)
                    err = np.abs(gotg - ref).max()
                    if err > 1e-4 * scale + 1e-9:
                        d_gpu = float(np.linalg.norm((gotg - ref64).ravel()))
                        d_cpu = float(np.linalg.norm((ref - ref64).ravel()))
                        nrm = float(np.linalg.norm(ref64.ravel()))
                        assert d_gpu <= (8.0 if tc else 4.0) * d_cpu + 1e-12 or d_gpu <= 1e-4 * nrm, (k, err / scale, d_gpu / nrm, d_cpu / nrm)
                        worst[k] = (float(err / scale), d_gpu / max(nrm, 1e-30), d_cpu / max(nrm, 1e-30))
            print("beyond 1e-4 (max-norm) of the fp32 oracle; (that ratio, L2 gpu vs f64, L2 oracle vs f64):", worst)
    # parameters after the two steps: Adam turns a rounding-level gradient difference into a full lr-sized step difference
    # (see tests/test_gpu_unet.py), so the typical element is held tight and no element may differ by more than 2 steps x lr
    for sd, mod in ((oracle.enc, enc), (oracle.dec, dec)):
        for k, v in mod.state_dict().items():
            ref = sd[k].detach().numpy()
            gv = v.detach().cpu().numpy()
            if ref.dtype.kind == "f" and not k.endswith(("running_mean", "running_var")):
                dev = np.abs(gv - ref)
                # typical element: 1e-4 of the tensor's scale plus 2.5 % of the two lr-sized Adam steps.  Batch-2 BatchNorm over
                # 3x3 ... 511x511 planes makes the SECOND step chaotic: the fp32 CPU oracle's own sums move with the host's
                # thread count, and the measured median deviation of e.g. decoder_lin.2.bias varies between 5e-6 and 1.2e-5
                # from one box to the next with identical GPU results.  (Step-0 gradients are held to the float64 evaluation
                # above; the loss of both steps to 2e-5.)
                assert np.median(dev) <= 1e-4 * max(np.abs(ref).max(), 1e-3) + 5e-5, k
                assert dev.max() <= 2.0 * 1e-3 * 2 + 1e-4 * np.abs(ref).max(), k


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("batch", [128])
def test_config4_full_batch_step_vs_float64_on_device(batch, tc, monkeypatch):
    """BASELINE configs[3] at its FULL size (4x64x64 -> 4x1024x1024, batch 128): one optimiser step of the production schedule -
    tcgen05 layers, tile-resident weight gradients, cp.async tile pipelines - against the oracle port evaluated in float64
    on the same GPU (torch ops; test infrastructure only).  At this batch the BatchNorm planes hold 10^3 ... 10^8 elements and
    the float64 evaluation is the reference: loss to 1e-5; every gradient to 1e-4 of its tensor's max-norm (north_star's bar) or
    to twice the distance at which torch's own fp32 evaluation (cuDNN / cuBLAS, TF32 off) of the same step sits from float64."""
    from cae_tools_b200.engine.convae import ConvAEEngine
    from cae_tools_b200.models.decoder import Decoder
    from cae_tools_b200.models.encoder import Encoder
    from cae_tools_b200.models.model_sizer import create_model_spec
    from oracle.torch_port import OracleModel
    if not tc:
        monkeypatch.setattr(ConvAEEngine, "TC_MIN_CIN", 1 << 30)
    free, _ = torch.cuda.mem_get_info()
    if free < 100 * 2 ** 30:
        pytest.skip("needs ~100 GB of free device memory for the float64 evaluation")
    dev = torch.device("cuda")
    spec = create_model_spec(input_size=(64, 64), input_channels=4, output_size=(1024, 1024), output_channels=4)
    torch.manual_seed(3)
    enc, dec = Encoder(spec.get_input_layers(), 4, 16), Decoder(spec.get_output_layers(), 4, 16)
    esd = {k: v.detach().clone().to(dev) for k, v in enc.state_dict().items()}
    dsd = {k: v.detach().clone().to(dev) for k, v in dec.state_dict().items()}
    gen = torch.Generator(device=dev).manual_seed(5)
    x = torch.rand(batch, 4, 64, 64, device=dev, generator=gen)
    y = torch.rand(batch, 4, 1024, 1024, device=dev, generator=gen)
    eng = ConvAEEngine(enc, dec, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(x, y, batch)
    names = [n for n, _ in eng._program("train", data, batch).sched]
    assert ("fwd.convT0.tc" in names) == tc and ("fwd.convT4.tc" in names) == tc and "fwd.convT5" in names
    got_loss = float(eng.train_epoch(data).cpu()[0])
    grads = {("enc." + k): p.grad.detach().double().cpu() for k, p in enc.named_parameters()}
    grads.update({("dec." + k): p.grad.detach().double().cpu() for k, p in dec.named_parameters()})
    del eng, data
    torch.cuda.empty_cache()
    exact = OracleModel(esd, dsd, spec.save(), zero_dead_bias_grads=True, dtype=torch.float64)
    want_loss = float(exact.train_step(x.double(), y.double()))
    ref64 = {(pre + k): v.grad.detach().cpu() for pre, sd in (("enc.", exact.enc), ("dec.", exact.dec)) for k, v in sd.items()
             if v.grad is not None}
    del exact
    torch.cuda.empty_cache()
    # the same step in plain fp32 torch (cuDNN / cuBLAS with TF32 off): how far does an fp32 evaluation sit from float64?
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        plain = OracleModel(esd, dsd, spec.save(), zero_dead_bias_grads=True)
        plain_loss = float(plain.train_step(x, y))
        ref32 = {(pre + k): v.grad.detach().double().cpu() for pre, sd in (("enc.", plain.enc), ("dec.", plain.dec))
                 for k, v in sd.items() if v.grad is not None}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert abs(got_loss - want_loss) <= 1e-5 * want_loss, (got_loss, want_loss, plain_loss)
    table = []
    for k, ref in ref64.items():
        scale = max(float(ref.abs().max()), 1e-12)
        table.append((float((grads[k] - ref).abs().max()) / scale, float((ref32[k] - ref).abs().max()) / scale, k))
    table.sort(reverse=True)
    print("config 4, batch", batch, ": loss", got_loss, "float64", want_loss, "torch fp32", plain_loss)
    for e_gpu, e_32, k in table[:12]:
        print(f"   {k:34s} this path vs float64 {e_gpu:.2e}   torch fp32 vs float64 {e_32:.2e}")
    for e_gpu, e_32, k in table:
        assert e_gpu <= max(1e-4, 2.0 * e_32), (k, e_gpu, e_32)
