"""Op-level parity of the C-ABI kernels (through ctypes) against plain torch fp64 on CPU.

Tolerance: 2e-5 of the max-norm for fp32 kernels (north_star asks 1e-4 per layer)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 2e-5


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda")


def close(a, b, tol=TOL, what=""):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(float(b.abs().max()), 1e-12)
    err = float((a - b).abs().max()) / scale
    assert err <= tol, f"{what}: rel err {err:.3e} > {tol}"


def rnd(*shape, seed=0, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale + shift)


CONV_CASES = [
    # N, Cin, Cout, k, s, p, H, W
    (3, 1, 2, 3, 2, 0, 16, 16),
    (2, 2, 4, 3, 2, 0, 7, 7),
    (2, 5, 3, (4, 3), 2, 0, 12, 11),
    (2, 3, 9, (3, 4), 2, 1, 9, 10),
    (2, 4, 16, 4, 2, 1, 10, 10),
    (1, 70, 20, 3, 2, 0, 9, 9),
    (2, 3, 4, 5, 3, 2, 14, 13),       # generic fallback
    (2, 2, 3, 3, 1, 1, 6, 7),         # generic fallback, stride 1
    (2, 2, 2, 2, 2, 0, 8, 8),         # generic fallback
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("mode", ["plain", "affine_relu", "affine2"])
def test_conv_down(case, mode):
    from cae_tools_b200.engine import ops
    dev = _dev()
    N, Ci, Co, k, s, p, H, W = case
    x = rnd(N, Ci, H, W, seed=1)
    w = rnd(Co, Ci, *((k, k) if isinstance(k, int) else k), seed=2, scale=0.3)
    b = rnd(Co, seed=3)
    k0, k1, k2 = rnd(Ci, seed=4), rnd(Ci, seed=5), rnd(Ci, seed=6)
    x1 = rnd(N, Ci, H, W, seed=7)
    sh = (1, Ci, 1, 1)
    if mode == "plain":
        xin = x
        src = ops.make_src(x.float().to(dev))
    elif mode == "affine_relu":
        xin = F.relu(x.float().double() * k0.float().double().view(sh) + k2.float().double().view(sh))
        src = ops.make_src(x.float().to(dev), k0=k0.float().to(dev), k2=k2.float().to(dev), relu=True)
    else:
        xin = x.float().double() * k0.float().double().view(sh) + x1.float().double() * k1.float().double().view(sh) \
            + k2.float().double().view(sh)
        src = ops.make_src(x.float().to(dev), t1=x1.float().to(dev), k0=k0.float().to(dev), k1=k1.float().to(dev),
                           k2=k2.float().to(dev))
    if mode == "plain":
        xin = x.float().double()
    ref = F.conv2d(xin, w.float().double(), b.float().double(), stride=s, padding=p)
    out = torch.full(ref.shape, float("nan"), dtype=torch.float32, device=dev)
    wd, bd = w.float().to(dev), b.float().to(dev)
    ops.conv_down(src, wd, ops.geom(k, s, p), ops.view4(out), ops.make_epilogue(ops.EPI_PLAIN, bias=bd))
    torch.cuda.synchronize()
    close(out, ref, what=f"conv_down {case} {mode}")


UP_CASES = [
    # N, Cin, Cout, k, s, p, op, H, W
    (3, 2, 1, 4, 2, 0, 0, 15, 15),
    (2, 4, 2, 3, 2, 0, 0, 7, 7),
    (2, 64, 32, 3, 2, 0, 0, 3, 3),
    (2, 5, 3, (4, 3), 2, 0, 0, 6, 5),
    (2, 3, 9, (3, 4), 2, 1, 1, 5, 6),
    (2, 16, 8, 4, 2, 1, 0, 4, 4),
    (2, 3, 2, 3, 2, 0, 1, 5, 5),      # output_padding
    (1, 4, 2, 8, 8, 0, 0, 3, 3),      # generic: k == s (unet last layer style)
    (2, 3, 4, 5, 3, 2, 1, 6, 5),      # generic
    (2, 2, 3, 3, 1, 1, 0, 6, 7),      # generic stride 1
]


@pytest.mark.parametrize("case", UP_CASES)
@pytest.mark.parametrize("mode", ["plain", "affine_relu"])
def test_conv_up(case, mode):
    from cae_tools_b200.engine import ops
    dev = _dev()
    N, Ci, Co, k, s, p, op, H, W = case
    x = rnd(N, Ci, H, W, seed=11).float()
    w = rnd(Ci, Co, *((k, k) if isinstance(k, int) else k), seed=12, scale=0.3).float()
    b = rnd(Co, seed=13).float()
    k0, k2 = rnd(Ci, seed=14).float(), rnd(Ci, seed=16).float()
    sh = (1, Ci, 1, 1)
    if mode == "plain":
        xin = x.double()
        src = ops.make_src(x.to(dev))
    else:
        xin = F.relu(x.double() * k0.double().view(sh) + k2.double().view(sh))
        src = ops.make_src(x.to(dev), k0=k0.to(dev), k2=k2.to(dev), relu=True)
    ref = F.conv_transpose2d(xin, w.double(), b.double(), stride=s, padding=p, output_padding=op)
    out = torch.full(ref.shape, float("nan"), dtype=torch.float32, device=dev)
    ops.conv_up(src, w.to(dev), ops.geom(k, s, p), ops.view4(out), ops.make_epilogue(ops.EPI_PLAIN, bias=b.to(dev)))
    torch.cuda.synchronize()
    close(out, ref, what=f"conv_up {case} {mode}")
    # sigmoid epilogue
    out2 = torch.empty_like(out)
    ops.conv_up(src, w.to(dev), ops.geom(k, s, p), ops.view4(out2), ops.make_epilogue(ops.EPI_SIGMOID, bias=b.to(dev)))
    torch.cuda.synchronize()
    close(out2, torch.sigmoid(ref), what=f"conv_up sigmoid {case}")


def test_conv_up_channel_offset_view_and_cursor():
    """writes into channels [0,C) of a 2C-channel buffer (skip-concat layout); input picked by a device cursor"""
    from cae_tools_b200.engine import ops
    dev = _dev()
    nb, B, Ci, Co = 3, 2, 3, 4
    X = rnd(nb * B, Ci, 5, 5, seed=21).float()
    w = rnd(Ci, Co, 3, 3, seed=22, scale=0.3).float()
    Xd = X.to(dev)
    cursor = torch.tensor([2], dtype=torch.int32, device=dev)
    big = torch.zeros(B, 2 * Co, 11, 11, dtype=torch.float32, device=dev)
    src = ops.make_src(Xd[:B], cursor=cursor, cursor_stride=B * Ci * 25)
    ops.conv_up(src, w.to(dev), ops.geom(3, 2, 0), ops.view4(big[:, :Co]), ops.make_epilogue(ops.EPI_PLAIN))
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(X[2 * B:3 * B].double(), w.double(), None, stride=2)
    close(big[:, :Co], ref, what="offset view")
    assert float(big[:, Co:].abs().max()) == 0.0


@pytest.mark.parametrize("up", [False, True])
@pytest.mark.parametrize("geo", [(3, 2, 0), (4, 2, 1), (5, 3, 2)])
def test_bn_stats_epilogue(up, geo):
    """conv + bias with the STATS epilogue == conv -> BatchNorm2d(training): scale/shift, running stats"""
    from cae_tools_b200.engine import ops
    dev = _dev()
    k, s, p = geo
    N, Ci, Co, H, W = 4, 3, 5, 9, 8
    x = rnd(N, Ci, H, W, seed=31).float()
    b = rnd(Co, seed=33).float()
    gamma, beta = (rnd(Co, seed=34).abs() + 0.5).float(), rnd(Co, seed=35).float()
    rm0, rv0 = rnd(Co, seed=36).float(), (rnd(Co, seed=37).abs() + 0.5).float()
    if up:
        w = rnd(Ci, Co, k, k, seed=32, scale=0.3).float()
        y_ref = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=s, padding=p)
    else:
        w = rnd(Co, Ci, k, k, seed=32, scale=0.3).float()
        y_ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p)
    rm, rv = rm0.double().clone(), rv0.double().clone()
    z_ref = F.batch_norm(y_ref, rm, rv, gamma.double(), beta.double(), True, 0.1, 1e-5)
    y = torch.empty(y_ref.shape, dtype=torch.float32, device=dev)
    scr = torch.zeros(7, Co, dtype=torch.float32, device=dev)
    rmd, rvd = rm0.to(dev), rv0.to(dev)
    nbt = torch.zeros(1, dtype=torch.int64, device=dev)
    gd, bd = gamma.to(dev), beta.to(dev)
    bn = ops.make_bn(Co, 1e-5, 0.1, gd, bd, rmd, rvd, nbt, scale=scr[0], shift=scr[1], mean=scr[2], invstd=scr[3])
    part = torch.zeros(ops.partials_len(Co), dtype=torch.float64, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    epi = ops.make_epilogue(ops.EPI_STATS, bias=b.to(dev), partials=part, ticket=ticket, bn=bn)
    fn = ops.conv_up if up else ops.conv_down
    for rep in range(2):  # twice: the ticket must reset itself
        fn(ops.make_src(x.to(dev)), w.to(dev), ops.geom(k, s, p), ops.view4(y), epi)
    torch.cuda.synchronize()
    close(y, y_ref, what="raw conv output")
    z = y.double().cpu() * scr[0].double().cpu().view(1, -1, 1, 1) + scr[1].double().cpu().view(1, -1, 1, 1)
    close(z, z_ref, what="normalised")
    assert int(nbt.item()) == 2 and int(ticket.item()) == 0
    rm2, rv2 = rm0.double().clone(), rv0.double().clone()
    for rep in range(2):
        F.batch_norm(y_ref, rm2, rv2, gamma.double(), beta.double(), True, 0.1, 1e-5)
    close(rmd, rm2, what="running_mean")
    close(rvd, rv2, what="running_var")


@pytest.mark.parametrize("geo", [(3, 2, 0, 0), (4, 2, 0, 0), ((4, 3), 2, 0, 0), (4, 2, 1, 0), (3, 2, 0, 1), (5, 3, 1, 0)])
def test_convT_layer_backward(geo):
    """ConvT -> BN(train) -> ReLU sandwich: dgrad with MASKSTATS epilogue, on-load BN-backward affine, wgrad,
    against autograd.  Layer l-1 (conv_prev -> BN -> ReLU) feeds layer l (convT)."""
    from cae_tools_b200.engine import ops
    dev = _dev()
    k, s, p, op = geo
    kh, kw = (k, k) if isinstance(k, int) else k
    N, C0, C1, C2, H, W = 3, 4, 6, 3, 5, 6
    # reference graph in fp64
    a0 = rnd(N, C0, H, W, seed=41).float().double().requires_grad_(True)           # input of layer l-1 (already activated)
    w1 = rnd(C0, C1, 3, 3, seed=42, scale=0.3).float().double().requires_grad_(True)
    g1, b1 = (rnd(C1, seed=43).abs() + 0.5).float().double().requires_grad_(True), rnd(C1, seed=44).float().double().requires_grad_(True)
    w2 = rnd(C1, C2, kh, kw, seed=45, scale=0.3).float().double().requires_grad_(True)
    g2, b2 = (rnd(C2, seed=46).abs() + 0.5).float().double().requires_grad_(True), rnd(C2, seed=47).float().double().requires_grad_(True)
    y1 = F.conv_transpose2d(a0, w1, None, stride=2)
    y1.retain_grad()
    a1 = F.relu(F.batch_norm(y1, None, None, g1, b1, True, 0.1, 1e-5))
    y2 = F.conv_transpose2d(a1, w2, None, stride=s, padding=p, output_padding=op)
    y2.retain_grad()
    a2 = F.relu(F.batch_norm(y2, None, None, g2, b2, True, 0.1, 1e-5))
    up = rnd(*a2.shape, seed=48).float().double()           # upstream gradient wrt a2
    (a2 * up).sum().backward()

    # device side: forward of both layers (STATS), then the backward kernels
    f32 = lambda t: t.detach().float().to(dev)
    a0d, w1d, w2d = f32(a0), f32(w1), f32(w2)
    y1d = torch.empty(y1.shape, dtype=torch.float32, device=dev)
    y2d = torch.empty(y2.shape, dtype=torch.float32, device=dev)
    s1, s2 = torch.zeros(7, C1, device=dev), torch.zeros(7, C2, device=dev)
    dg1, db1, dg2, db2 = (torch.zeros(c, device=dev) for c in (C1, C1, C2, C2))
    mk = lambda Cn, g, b, sc, dg, db: ops.make_bn(Cn, 1e-5, 0.1, f32(g), f32(b), scale=sc[0], shift=sc[1], mean=sc[2],
                                                  invstd=sc[3], dgamma=dg, dbeta=db, bwdA=sc[4], bwdB=sc[5], bwdC=sc[6])
    keep = []

    def P(Cn):
        t = torch.zeros(ops.partials_len(Cn), dtype=torch.float64, device=dev)
        keep.append(t)
        return t

    def T():
        t = torch.zeros(1, dtype=torch.int32, device=dev)
        keep.append(t)
        return t

    bn1, bn2 = mk(C1, g1, b1, s1, dg1, db1), mk(C2, g2, b2, s2, dg2, db2)
    ops.conv_up(ops.make_src(a0d), w1d, ops.geom(3, 2, 0), ops.view4(y1d),
                ops.make_epilogue(ops.EPI_STATS, partials=P(C1), ticket=T(), bn=bn1))
    src1 = ops.make_src(y1d, k0=s1[0], k2=s1[1], relu=True)
    ops.conv_up(src1, w2d, ops.geom(k, s, p), ops.view4(y2d),
                ops.make_epilogue(ops.EPI_STATS, partials=P(C2), ticket=T(), bn=bn2))
    # gradient wrt a2 arrives -> mask + BN sums of layer 2 (elementwise member)
    dz2 = torch.empty_like(y2d)
    ops.ew_epilogue(ops.make_src(f32(up)), ops.view4(dz2),
                    ops.make_epilogue(ops.EPI_MASKSTATS, partials=P(C2), ticket=T(), bn=bn2, act=y2d))
    dy2 = ops.make_src(dz2, t1=y2d, k0=s2[4], k1=s2[5], k2=s2[6])
    # weight gradient of layer 2 and its dgrad (conv_down) with mask + BN sums of layer 1
    gw2 = torch.empty_like(w2d)
    g = ops.geom(k, s, p)
    part = torch.zeros(ops.wgrad_partials_len(src1, dy2, g), dtype=torch.float32, device=dev)
    ops.conv_wgrad(src1, dy2, g, gw2, part, T())
    dz1 = torch.empty_like(y1d)
    ops.conv_down(dy2, w2d, g, ops.view4(dz1),
                  ops.make_epilogue(ops.EPI_MASKSTATS, partials=P(C1), ticket=T(), bn=bn1, act=y1d))
    torch.cuda.synchronize()
    close(y2d, y2, what="y2")
    close(dg2, g2.grad, 1e-4, "dgamma2")
    close(db2, b2.grad, 1e-4, "dbeta2")
    dy2_dev = dz2.double() * s2[4].double().view(1, -1, 1, 1) + y2d.double() * s2[5].double().view(1, -1, 1, 1) + \
        s2[6].double().view(1, -1, 1, 1)
    close(dy2_dev, y2.grad, 1e-4, "dL/dy2")
    close(gw2, w2.grad, 1e-4, "wgrad2")
    close(dg1, g1.grad, 1e-4, "dgamma1")
    close(db1, b1.grad, 1e-4, "dbeta1")
    dy1_dev = dz1.double() * s1[4].double().view(1, -1, 1, 1) + y1d.double() * s1[5].double().view(1, -1, 1, 1) + \
        s1[6].double().view(1, -1, 1, 1)
    close(dy1_dev, y1.grad, 1e-4, "dL/dy1")


@pytest.mark.parametrize("geo", [(3, 2, 0), (4, 2, 1), ((3, 4), 2, 0), (5, 3, 2)])
def test_conv_layer_backward(geo):
    """Conv2d: wgrad (small operand = dL/dy) and dgrad through conv_up, against autograd"""
    from cae_tools_b200.engine import ops
    dev = _dev()
    k, s, p = geo
    kh, kw = (k, k) if isinstance(k, int) else k
    N, Ci, Co, H, W = 3, 3, 5, 12, 11
    x = rnd(N, Ci, H, W, seed=51).float().double().requires_grad_(True)
    w = rnd(Co, Ci, kh, kw, seed=52, scale=0.3).float().double().requires_grad_(True)
    y = F.conv2d(x, w, None, stride=s, padding=p)
    up = rnd(*y.shape, seed=53).float().double()
    (y * up).sum().backward()
    f32 = lambda t: t.detach().float().to(dev)
    xd, wd, upd = f32(x), f32(w), f32(up)
    g = ops.geom(k, s, p)
    gw = torch.empty_like(wd)
    ssrc, bsrc = ops.make_src(upd), ops.make_src(xd)
    part = torch.zeros(ops.wgrad_partials_len(ssrc, bsrc, g), dtype=torch.float32, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.conv_wgrad(ssrc, bsrc, g, gw, part, ticket)
    dx = torch.empty_like(xd)
    ops.conv_up(ssrc, wd, g, ops.view4(dx), ops.make_epilogue(ops.EPI_PLAIN))
    torch.cuda.synchronize()
    close(gw, w.grad, 1e-4, "conv wgrad")
    close(dx, x.grad, 1e-4, "conv dgrad")


def test_sigmoid_mse_epilogue_with_cursor():
    from cae_tools_b200.engine import ops
    dev = _dev()
    nb, B, Ci, Co = 3, 4, 2, 2
    x = rnd(B, Ci, 7, 7, seed=61).float()
    w = rnd(Ci, Co, 4, 4, seed=62, scale=0.3).float()
    b = rnd(Co, seed=63).float()
    Y = torch.rand(nb * B, Co, 16, 16, generator=torch.Generator().manual_seed(64))
    cur = 1
    v = F.conv_transpose2d(x.double(), w.double(), b.double(), stride=2).requires_grad_(True)
    yh = torch.sigmoid(v)
    loss = F.mse_loss(yh, Y[cur * B:(cur + 1) * B].double())
    loss.backward()
    Yd = Y.to(dev)
    cursor = torch.tensor([cur], dtype=torch.int32, device=dev)
    losses = torch.zeros(nb, device=dev)
    dbias = torch.zeros(Co, device=dev)
    part = torch.zeros(ops.partials_len(Co), dtype=torch.float64, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    tgt = ops.make_src(Yd[:B], cursor=cursor, cursor_stride=B * Co * 256)
    out = torch.empty(B, Co, 16, 16, device=dev)
    for wm in (0, 1):
        epi = ops.make_epilogue(ops.EPI_SIGMOID_MSE, bias=b.to(dev), partials=part, ticket=ticket, target=tgt,
                                loss_out=losses, dbias=dbias, write_mode=wm)
        ops.conv_up(ops.make_src(x.to(dev)), w.to(dev), ops.geom(4, 2, 0), ops.view4(out), epi)
        torch.cuda.synchronize()
        if wm == 0:
            close(out, v.grad, 1e-4, "dL/dz")
        else:
            close(out, yh, what="yhat")
        assert abs(float(losses[cur]) - float(loss)) <= 1e-5 * float(loss)
        assert float(losses[0]) == 0.0 and float(losses[2]) == 0.0
        close(dbias, v.grad.sum(dim=(0, 2, 3)), 1e-4, "dbias")


def test_gemm_variants():
    from cae_tools_b200.engine import ops
    dev = _dev()
    N, K, O, hw = 37, 36, 19, 9
    x = rnd(N, K, seed=71).float()
    W = rnd(O, K, seed=72, scale=0.3).float()
    b = rnd(O, seed=73).float()
    k0, k2 = rnd(K // hw, seed=74).float(), rnd(K // hw, seed=75).float()
    xd, Wd, bd = x.to(dev), W.to(dev), b.to(dev)
    # forward with BN+ReLU+flatten on load, bias, relu
    a = F.relu(x.double() * k0.double().repeat_interleave(hw) + k2.double().repeat_interleave(hw))
    ref = F.relu(a @ W.double().t() + b.double())
    out = torch.empty(N, O, device=dev)
    ops.gemm(N, O, K, xd, K, 1, Wd, 1, K, out, O, 1, a_k0=k0.to(dev), a_k2=k2.to(dev), a_hw=hw, a_relu=True, bias=bd,
             relu_out=True)
    torch.cuda.synchronize()
    close(out, ref, what="gemm fwd")
    # dW = dy^T a (B operand transformed), db = rowsum
    dy = rnd(N, O, seed=76).float()
    dW = torch.empty(O, K, device=dev)
    db = torch.empty(O, device=dev)
    ops.gemm(O, K, N, dy.to(dev), 1, O, xd, K, 1, dW, K, 1, b_k0=k0.to(dev), b_k2=k2.to(dev), b_hw=hw, b_relu=True,
             rowsum_A=db)
    torch.cuda.synchronize()
    close(dW, dy.double().t() @ a, what="gemm dW")
    close(db, dy.double().sum(0), what="gemm db")
    # dx = (dy W) * mask
    mask = rnd(N, K, seed=77).float()
    dx = torch.empty(N, K, device=dev)
    ops.gemm(N, K, O, dy.to(dev), O, 1, Wd, K, 1, dx, K, 1, mask=mask.to(dev))
    torch.cuda.synchronize()
    close(dx, (dy.double() @ W.double()) * (mask.double() > 0), what="gemm dx")


@pytest.mark.parametrize("K", [256, 576, 1000])
def test_gemm_long_k_few_outputs(K):
    """the warp-per-output path (cae_gemm with K >= 256 and M*N <= 16384): bias, relu, mask, strided operands"""
    from cae_tools_b200.engine import ops
    dev = _dev()
    N, O = 64, 16
    dy = rnd(N, K, seed=171).float()
    W = rnd(K, O, seed=172, scale=0.2).float()              # nn.Linear weight [out = K][in = O]: dx = dy W
    mask = rnd(N, O, seed=173).float()
    b = rnd(O, seed=174).float()
    dx = torch.empty(N, O, device=dev)
    ops.gemm(N, O, K, dy.to(dev), K, 1, W.to(dev), O, 1, dx, O, 1, mask=mask.to(dev))
    torch.cuda.synchronize()
    close(dx, (dy.double() @ W.double()) * (mask.double() > 0), what="dx long K")
    out = torch.empty(N, O, device=dev)
    ops.gemm(N, O, K, dy.to(dev), K, 1, W.to(dev), O, 1, out, O, 1, bias=b.to(dev), relu_out=True)
    torch.cuda.synchronize()
    close(out, F.relu(dy.double() @ W.double() + b.double()), what="fwd long K")


@pytest.mark.parametrize("decoupled", [False, True])
def test_adam_matches_torch(decoupled):
    from cae_tools_b200.engine import ops
    dev = _dev()
    n = 1000 + 3
    p0 = rnd(n, seed=81).float()
    tp = p0.clone().requires_grad_(True)
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([tp], lr=1e-3, weight_decay=1e-2)
    p = p0.to(dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    cursor = torch.zeros(1, dtype=torch.int32, device=dev)
    for t in range(5):
        g = rnd(n, seed=90 + t).float()
        tp.grad = g.clone()
        opt.step()
        ops.adam(p, g.to(dev), m, v, n, 1e-3, 0.9, 0.999, 1e-8, 1e-2, decoupled, 1.0, step)
        ops.step_advance(step, cursor, 3)
    torch.cuda.synchronize()
    assert int(step.item()) == 5 and int(cursor.item()) == 5 % 3
    close(p, tp, 1e-6, "adam params")


def test_mse_kernel():
    from cae_tools_b200.engine import ops
    dev = _dev()
    n = 123457
    a, b = rnd(n, seed=101).float(), rnd(n, seed=102).float()
    part = torch.zeros(ops.partials_len(1), dtype=torch.float64, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.zeros(2, device=dev)
    cursor = torch.tensor([1], dtype=torch.int32, device=dev)
    ops.mse(a.to(dev), b.to(dev), n, part, ticket, out, cursor)
    torch.cuda.synchronize()
    ref = float(((a.double() - b.double()) ** 2).mean())
    assert abs(float(out[1]) - ref) <= 1e-6 * ref and float(out[0]) == 0.0


def test_bad_arguments_raise():
    from cae_tools_b200.engine import ops
    from cae_tools_b200._lib import CaeError
    dev = _dev()
    x = torch.zeros(1, 1, 4, 4, device=dev)
    w = torch.zeros(1, 1, 3, 3, device=dev)
    out = torch.zeros(1, 1, 5, 5, device=dev)   # wrong size for k3 s2 (expects 9x9)
    with pytest.raises(CaeError):
        ops.conv_up(ops.make_src(x), w, ops.geom(3, 2, 0), ops.view4(out), ops.make_epilogue(ops.EPI_PLAIN))
