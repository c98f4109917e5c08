"""cae_gemm routed to the tensor cores (tc_dense.cu) - the nn.Linear contractions of the large-fc regimes (SURVEY 8f rows 2
and 3): op level against torch float64 in every operand layout the engines use, and one training step of the UNET with
fc 3200 / latent 800 at batch 128 against the oracle port and against the SIMT route."""
import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64)


def _maxrel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


@pytest.mark.parametrize("N,fin,fout", [(256, 800, 3200), (128, 3200, 800), (200, 1000, 1000), (384, 36, 4100)])
def test_linear_layer_three_gemms_on_tensor_cores(N, fin, fout):
    """forward (on-load BatchNorm + ReLU, bias, ReLU), input gradient (mask), weight gradient (+ bias row sums) of one
    nn.Linear, each through ops.gemm with the strides the engines pass; eligible -> tcgen05 route, checked against float64
    and against the SIMT kernel"""
    from cae_tools_b200.engine import ops
    from cae_tools_b200 import _lib
    dev = torch.device("cuda")
    hw = 4 if fin % 4 == 0 else 1
    x, W, b = rnd(N, fin, seed=1), rnd(fout, fin, seed=2) * 0.05, rnd(fout, seed=3)
    k0, k2 = rnd(fin // hw, seed=4).abs() + 0.5, rnd(fin // hw, seed=5)
    a = torch.relu(x * k0.repeat_interleave(hw) + k2.repeat_interleave(hw))
    y = torch.relu(a @ W.T + b)
    dy = rnd(N, fout, seed=6) * (y > 0)
    dW, db, dx = dy.T @ a, dy.sum(0), (dy @ W) * (a > 0)
    xd, Wd, bd, k0d, k2d, dyd = (t.float().to(dev) for t in (x, W, b, k0, k2, dy))
    ad = a.float().to(dev)
    res = {}
    for tc in (True, False):
        ops.USE_TC_DENSE = tc
        try:
            yo = torch.full((N, fout), float("nan"), device=dev)
            ops.gemm(N, fout, fin, xd, fin, 1, Wd, 1, fin, yo, fout, 1, a_k0=k0d, a_k2=k2d, a_hw=hw, a_relu=True, bias=bd,
                     relu_out=True)
            dWo, dbo = torch.full((fout, fin), float("nan"), device=dev), torch.full((fout,), float("nan"), device=dev)
            ops.gemm(fout, fin, N, dyd, 1, fout, xd, fin, 1, dWo, fin, 1, b_k0=k0d, b_k2=k2d, b_hw=hw, b_relu=True, rowsum_A=dbo)
            dxo = torch.full((N, fin), float("nan"), device=dev)
            ops.gemm(N, fin, fout, dyd, fout, 1, Wd, fin, 1, dxo, fin, 1, mask=ad)
            torch.cuda.synchronize()
        finally:
            ops.USE_TC_DENSE = True
        res[tc] = (yo, dWo, dbo, dxo)
        for got, want, what in zip(res[tc], (y, dW, db, dx), ("forward", "weight gradient", "bias gradient", "input gradient")):
            # 3xTF32: ~3e-6 at K = 1000 (fp32 accumulation inside the MMA), see tests/test_gpu_tc_gemm.py
            assert _maxrel(got, want) < 2e-5, (what, tc, _maxrel(got, want))
    # the route was really taken for the eligible shapes
    g = _lib.CaeGemm(N, fout, fin, xd.data_ptr(), fin, 1, Wd.data_ptr(), 1, fin, res[True][0].data_ptr(), fout, 1)
    import ctypes
    need = _lib.lib().cae_gemm_tc_workspace(ctypes.byref(g))
    assert (need > 0) == (min(N, fout) >= 128 and fin >= 32 and N * fout * fin >= 5e7)


def test_small_or_skinny_problems_stay_on_the_simt_kernel():
    from cae_tools_b200 import _lib
    import ctypes
    x = torch.zeros(4, device="cuda")
    for M, N, K in ((64, 65536, 256), (256, 64, 4096), (128, 128, 16), (128, 128, 2048)):
        g = _lib.CaeGemm(M, N, K, x.data_ptr(), K, 1, x.data_ptr(), 1, K, x.data_ptr(), N, 1)
        assert _lib.lib().cae_gemm_tc_workspace(ctypes.byref(g)) == 0, (M, N, K)


def test_unet_large_fc_training_step_vs_oracle():
    """SURVEY 8f row 2 regime (fc 3200, latent 800) on the shipped 16x16 -> 256x256 spec at batch 128: the fc GEMMs with
    >= 0.1 GFLOP run on tcgen05; losses and every gradient of one step against the oracle port (fp32 CPU) with float64
    adjudication, and against the same engine on the SIMT route"""
    from cae_tools_b200.engine import ops
    from cae_tools_b200.engine.unet import UNetEngine
    from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
    from oracle.torch_port import OracleUNet
    from test_gpu_unet import _shipped_spec
    spec, spec_json = _shipped_spec()
    batch = 128
    gen = torch.Generator().manual_seed(5)
    x, y = torch.rand(batch, 1, 16, 16, generator=gen), torch.rand(batch, 1, 256, 256, generator=gen)
    ones = torch.ones_like(y)
    torch.manual_seed(11)
    enc = UNetEncoder(spec.get_input_layers(), 800, 3200, 0.0)
    dec = UNetDecoder(spec.get_output_layers(), 800, 3200, 0.0)
    esd, dsd = {k: v.clone() for k, v in enc.state_dict().items()}, {k: v.clone() for k, v in dec.state_dict().items()}
    oracle = OracleUNet(esd, dsd, spec_json, lambda_pearson=1.0, zero_dead_bias_grads=True)
    exact = OracleUNet(esd, dsd, spec_json, lambda_pearson=1.0, zero_dead_bias_grads=True, dtype=torch.float64)
    want = oracle.train_step(x, y, ones)
    exact.train_step(x.double(), y.double(), ones.double())
    grads = {}
    for tc in (True, False):
        ops.USE_TC_DENSE = tc
        try:
            e2, d2 = UNetEncoder(spec.get_input_layers(), 800, 3200, 0.0), UNetDecoder(spec.get_output_layers(), 800, 3200, 0.0)
            e2.load_state_dict(esd)
            d2.load_state_dict(dsd)
            eng = UNetEngine(e2, d2, lambda_pearson=1.0, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5)
            data = eng.bind(x, y, batch)
            mse = float(eng.train_epoch(data).cpu()[0])
            pl = float(data.pearson.cpu()[0])
            assert eng._train_stem(batch) is None            # this geometry runs the per-layer chain
            names = [n for n, _ in eng._program("train", data, batch).sched]
            assert "fwd.fc2" in names
        finally:
            ops.USE_TC_DENSE = True
        assert abs(mse - want[0]) <= 2e-5 * want[0] and abs(pl - want[1]) <= 2e-5 * abs(want[1]), (tc, mse, pl, want)
        grads[tc] = {("enc." + k): p.grad.detach().cpu().numpy() for k, p in e2.named_parameters()}
        grads[tc].update({("dec." + k): p.grad.detach().cpu().numpy() for k, p in d2.named_parameters()})
    assert any(k[1:] == (batch, 800, 3200) for k in ops._gemm_ws), list(ops._gemm_ws)[:4]
    for pre, sd, sd64 in (("enc.", oracle.enc, exact.enc), ("dec.", oracle.dec, exact.dec)):
        for k in sd:
            if sd[k].grad is None:
                continue
            r, r64 = sd[k].grad.numpy(), sd64[k].grad.numpy()
            scale = max(np.abs(r).max(), 1e-7)
            for tc in (True, False):
                got = grads[tc][pre + k]
                err = np.abs(got - r).max()
                if err > 1e-4 * scale + 1e-9:
                    e_gpu, e_cpu = np.abs(got - r64).max(), np.abs(r - r64).max()
                    assert e_gpu <= 2.0 * e_cpu + 1e-9 and err <= 1e-3 * scale, (k, tc, err, scale, e_gpu, e_cpu)


def test_linear_model_large_batch_on_tensor_cores_vs_oracle():
    """LinearModel at a batch that makes its 256 -> 4096 contraction dense (both GEMMs eligible for
    tcgen05): three optimiser steps against the oracle port"""
    from cae_tools_b200.engine import ops
    from cae_tools_b200.engine.linear import LinearEngine
    from cae_tools_b200.models.linear import Linear
    from oracle.torch_port import OracleLinear
    torch.manual_seed(5)
    mod = Linear((1, 16, 16), (1, 64, 64))
    oracle = OracleLinear(mod.state_dict(), (1, 64, 64), lr=1e-3, weight_decay=1e-5)
    g = torch.Generator().manual_seed(6)
    B = 384
    X, Y = torch.rand(B, 1, 16, 16, generator=g), torch.rand(B, 1, 64, 64, generator=g)
    eng = LinearEngine(mod, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(X, Y, B)
    before = len(ops._gemm_ws)
    for step in range(3):
        got = float(eng.train_epoch(data).cpu()[0])
        want = oracle.train_step(X, Y)
        assert abs(got - want) <= 2e-5 * want, (step, got, want)
        if step == 0:
            for got_g, ref_g, what in ((mod.linear[1].weight.grad, oracle.w.grad, "dW"), (mod.linear[1].bias.grad, oracle.b.grad, "db")):
                assert _maxrel(got_g, ref_g) < 2e-5, (what, _maxrel(got_g, ref_g))
    assert len(ops._gemm_ws) >= before + 2          # forward and weight-gradient GEMMs took the tensor-core route
    # Weights after three Adam steps.  The gradients carry the 1 / (B * 4096) of the mean: 1e-6 ... 1e-5, and the few elements
    # whose gradient cancels to below Adam's eps = 1e-8 turn a 1e-10 summation difference (3xTF32 vs MKL, both ~1e-6 of the sum
    # of magnitudes) into a visible fraction of the lr-sized step.  Typical element tight, 99.9 % within 1e-4, none beyond 5 %
    # of the three steps.
    for got_p, ref_p in ((mod.linear[1].weight, oracle.w), (mod.linear[1].bias, oracle.b)):
        dev = (got_p.detach().cpu() - ref_p.detach()).abs().numpy()
        scale = float(ref_p.detach().abs().max())
        assert np.median(dev) <= 1e-5 * scale and np.quantile(dev, 0.999) <= 1e-4 * scale and dev.max() <= 0.05 * 3e-3, \
            (float(np.median(dev)), float(np.quantile(dev, 0.999)), float(dev.max()), scale)
