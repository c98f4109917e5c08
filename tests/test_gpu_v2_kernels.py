"""The tiled (v2) kernels against the generic (v1) kernels and against torch fp64, at geometries that exercise
multi-tile grids, ragged edges, channel chunking and both weight-gradient regimes."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64) * scale


def close(a, b, tol, what):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)
    assert err <= tol, f"{what}: rel err {err:.3e} > {tol}"


@pytest.fixture(autouse=True)
def _gen2():
    from cae_tools_b200.engine import ops
    ops.set_kernel_generation(2)
    yield
    ops.set_kernel_generation(2)


BIG = [
    # N, Cin, Cout, k, p, H, W  (transposed conv input size)
    (3, 2, 1, 4, 0, 127, 127),
    (2, 4, 2, 3, 0, 63, 63),
    (5, 8, 4, 3, 0, 31, 31),
    (9, 16, 8, 3, 0, 15, 15),
    (17, 32, 16, 3, 0, 7, 7),
    (33, 64, 32, 3, 0, 3, 3),
    (2, 70, 9, 3, 0, 20, 37),
    (3, 16, 8, 4, 1, 8, 8),
    (2, 3, 5, 4, 1, 33, 17),
]


@pytest.mark.parametrize("case", BIG)
def test_up_down_wgrad_v2(case):
    from cae_tools_b200.engine import ops
    dev = torch.device("cuda")
    N, Ci, Co, k, p, H, W = case
    x = rnd(N, Ci, H, W, seed=1).float()
    w = rnd(Ci, Co, k, k, seed=2, scale=0.2).float()
    b = rnd(Co, seed=3).float()
    k0, k2 = (rnd(Ci, seed=4).abs() + 0.5).float(), rnd(Ci, seed=5).float()
    xin = F.relu(x.double() * k0.double().view(1, -1, 1, 1) + k2.double().view(1, -1, 1, 1)).requires_grad_(True)
    wd = w.double().requires_grad_(True)
    ref = F.conv_transpose2d(xin, wd, b.double(), stride=2, padding=p)
    up = rnd(*ref.shape, seed=6).float()
    (ref * up.double()).sum().backward()
    xd, wdev, bd, k0d, k2d, upd = (t.to(dev) for t in (x, w, b, k0, k2, up))
    src = ops.make_src(xd, k0=k0d, k2=k2d, relu=True)
    g = ops.geom(k, 2, p)
    for gen in (2, 1):
        ops.set_kernel_generation(gen)
        out = torch.full(ref.shape, float("nan"), dtype=torch.float32, device=dev)
        ops.conv_up(src, wdev, g, ops.view4(out), ops.make_epilogue(ops.EPI_PLAIN, bias=bd))
        # dgrad: strided conv of the upstream gradient
        dx = torch.full(x.shape, float("nan"), dtype=torch.float32, device=dev)
        ops.conv_down(ops.make_src(upd), wdev, g, ops.view4(dx), ops.make_epilogue(ops.EPI_PLAIN))
        gw = torch.full(w.shape, float("nan"), dtype=torch.float32, device=dev)
        dy = ops.make_src(upd)
        part = torch.zeros(ops.wgrad_partials_len(src, dy, g), dtype=torch.float32, device=dev)
        ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        for _ in range(2):
            ops.conv_wgrad(src, dy, g, gw, part, ticket)
        torch.cuda.synchronize()
        close(out, ref, 2e-5, f"up gen{gen} {case}")
        close(dx, xin.grad, 5e-5, f"down gen{gen} {case}")
        close(gw, wd.grad, 1e-4, f"wgrad gen{gen} {case}")
        assert int(ticket.item()) == 0


@pytest.mark.parametrize("k", [3, 4])
def test_v2_stats_and_maskstats_match_v1_bitwise_inputs(k):
    """STATS / MASKSTATS / SIGMOID_MSE epilogues through the tiled kernels vs the generic ones"""
    from cae_tools_b200.engine import ops
    dev = torch.device("cuda")
    N, Ci, Co, H, W = 6, 4, 3, 21, 19
    x = rnd(N, Ci, H, W, seed=11).float().to(dev)
    w = rnd(Ci, Co, k, k, seed=12, scale=0.3).float().to(dev)
    b = rnd(Co, seed=13).float().to(dev)
    Ho, Wo = (H - 1) * 2 + k, (W - 1) * 2 + k
    tgt = torch.rand(N, Co, Ho, Wo, generator=torch.Generator().manual_seed(14)).to(dev)
    res = {}
    for gen in (1, 2):
        ops.set_kernel_generation(gen)
        scr = torch.zeros(7, Co, device=dev)
        gam, bet = torch.ones(Co, device=dev) * 1.3, torch.ones(Co, device=dev) * 0.1
        rm, rv = torch.zeros(Co, device=dev), torch.ones(Co, device=dev)
        bn = ops.make_bn(Co, 1e-5, 0.1, gam, bet, rm, rv, None, scale=scr[0], shift=scr[1], mean=scr[2], invstd=scr[3],
                         bwdA=scr[4], bwdB=scr[5], bwdC=scr[6], dgamma=torch.zeros(Co, device=dev),
                         dbeta=torch.zeros(Co, device=dev))
        P = lambda: torch.zeros(ops.partials_len(max(Co, Ci)), dtype=torch.float64, device=dev)
        T = lambda: torch.zeros(1, dtype=torch.int32, device=dev)
        y = torch.empty(N, Co, Ho, Wo, device=dev)
        ops.conv_up(ops.make_src(x), w, ops.geom(k, 2, 0), ops.view4(y),
                    ops.make_epilogue(ops.EPI_STATS, bias=b, partials=P(), ticket=T(), bn=bn))
        # loss epilogue on the same conv
        losses, dbias = torch.zeros(1, device=dev), torch.zeros(Co, device=dev)
        dz = torch.empty_like(y)
        ops.conv_up(ops.make_src(x), w, ops.geom(k, 2, 0), ops.view4(dz),
                    ops.make_epilogue(ops.EPI_SIGMOID_MSE, bias=b, partials=P(), ticket=T(), target=ops.make_src(tgt),
                                      loss_out=losses, dbias=dbias))
        # dgrad with mask + BN sums of a (fictitious) producer layer whose raw output is x itself
        scr_in = torch.zeros(7, Ci, device=dev)
        scr_in[0] = 0.7; scr_in[1] = 0.05; scr_in[2] = 0.1; scr_in[3] = 1.2
        bn_in = ops.make_bn(Ci, 1e-5, 0.1, torch.ones(Ci, device=dev), torch.zeros(Ci, device=dev), scale=scr_in[0],
                            shift=scr_in[1], mean=scr_in[2], invstd=scr_in[3], bwdA=scr_in[4], bwdB=scr_in[5],
                            bwdC=scr_in[6], dgamma=torch.zeros(Ci, device=dev), dbeta=torch.zeros(Ci, device=dev))
        dzx = torch.empty_like(x)
        ops.conv_down(ops.make_src(dz), w, ops.geom(k, 2, 0), ops.view4(dzx),
                      ops.make_epilogue(ops.EPI_MASKSTATS, partials=P(), ticket=T(), bn=bn_in, act=x))
        torch.cuda.synchronize()
        res[gen] = [t.clone() for t in (y, scr, rm, rv, losses, dbias, dz, dzx, scr_in)]
    names = ["y", "bn scratch", "running_mean", "running_var", "loss", "dbias", "dz", "dz_prev", "bn bwd coefficients"]
    for name, a, b_ in zip(names, res[2], res[1]):
        close(a, b_, 2e-5, name)


def padded(t):
    """same values, rows padded to a multiple of 4 floats (what the engine allocates for wide layers)"""
    W = t.shape[-1]
    ld = (W + 3) // 4 * 4
    buf = torch.full(t.shape[:-1] + (ld,), 7.0, dtype=t.dtype, device=t.device)   # poison in the padding
    v = buf[..., :W]
    v.copy_(t)
    return v


@pytest.mark.parametrize("case", [(3, 2, 1, 4, 127), (2, 4, 2, 3, 63), (3, 8, 4, 3, 31), (2, 3, 5, 3, 45), (2, 6, 3, 4, 26)])
def test_direct_v3_kernels_on_padded_buffers(case):
    """wide thin layers: the float4 'direct' kernels (mask bit 16) vs the generic ones, every epilogue"""
    from cae_tools_b200.engine import ops
    dev = torch.device("cuda")
    N, Ci, Co, k, H = case
    W = H + 2
    Ho, Wo = (H - 1) * 2 + k, (W - 1) * 2 + k
    x = padded(rnd(N, Ci, H, W, seed=21).float().to(dev))
    x1 = padded(rnd(N, Ci, H, W, seed=22).float().to(dev))
    w = rnd(Ci, Co, k, k, seed=23, scale=0.3).float().to(dev)
    b = rnd(Co, seed=24).float().to(dev)
    k0, k1, k2 = (rnd(Ci, seed=25).abs() + 0.5).float().to(dev), rnd(Ci, seed=26).float().to(dev), rnd(Ci, seed=27).float().to(dev)
    tgt = torch.rand(3 * N, Co, Ho, Wo, generator=torch.Generator().manual_seed(28)).to(dev) if Wo % 4 == 0 else None
    tgt_p = padded(torch.rand(N, Co, Ho, Wo, generator=torch.Generator().manual_seed(28)).to(dev))
    cursor = torch.tensor([2], dtype=torch.int32, device=dev)
    res = {}
    for mask in (0, 16):
        ops.set_kernel_generation((mask << 4) | 3)
        g = ops.geom(k, 2, 0)
        src = ops.make_src(x, k0=k0, k2=k2, relu=True)
        scr = torch.zeros(7, Co, device=dev)
        gam, bet = torch.full((Co,), 1.3, device=dev), torch.full((Co,), 0.1, device=dev)
        bn = ops.make_bn(Co, 1e-5, 0.1, gam, bet, scale=scr[0], shift=scr[1], mean=scr[2], invstd=scr[3], bwdA=scr[4],
                         bwdB=scr[5], bwdC=scr[6], dgamma=torch.zeros(Co, device=dev), dbeta=torch.zeros(Co, device=dev))
        P = lambda: torch.zeros(ops.partials_len(max(Co, Ci)), dtype=torch.float64, device=dev)
        T = lambda: torch.zeros(1, dtype=torch.int32, device=dev)
        y = padded(torch.zeros(N, Co, Ho, Wo, device=dev))
        ops.conv_up(src, w, g, ops.view4(y), ops.make_epilogue(ops.EPI_STATS, bias=b, partials=P(), ticket=T(), bn=bn))
        ysig = padded(torch.zeros(N, Co, Ho, Wo, device=dev))
        ops.conv_up(src, w, g, ops.view4(ysig), ops.make_epilogue(ops.EPI_SIGMOID, bias=b))
        losses, dbias = torch.zeros(3, device=dev), torch.zeros(Co, device=dev)
        dz = padded(torch.zeros(N, Co, Ho, Wo, device=dev))
        if tgt is not None:
            tsrc = ops.make_src(tgt[:N], cursor=cursor, cursor_stride=N * Co * Ho * Wo)
        else:
            tsrc = ops.make_src(tgt_p)
        ops.conv_up(src, w, g, ops.view4(dz),
                    ops.make_epilogue(ops.EPI_SIGMOID_MSE, bias=b, partials=P(), ticket=T(), target=tsrc,
                                      loss_out=losses, dbias=dbias))
        # dgrad of the layer with the BatchNorm-backward affine on load (two tensors) and mask + sums in the epilogue
        scr_in = torch.zeros(7, Ci, device=dev)
        scr_in[0] = 0.7; scr_in[1] = 0.05; scr_in[2] = 0.1; scr_in[3] = 1.2
        bn_in = ops.make_bn(Ci, 1e-5, 0.1, torch.ones(Ci, device=dev), torch.zeros(Ci, device=dev), scale=scr_in[0],
                            shift=scr_in[1], mean=scr_in[2], invstd=scr_in[3], bwdA=scr_in[4], bwdB=scr_in[5],
                            bwdC=scr_in[6], dgamma=torch.zeros(Ci, device=dev), dbeta=torch.zeros(Ci, device=dev))
        kA, kB, kC = (rnd(Co, seed=31).abs() + 0.5).float().to(dev), rnd(Co, seed=32).float().to(dev), rnd(Co, seed=33).float().to(dev)
        dy = ops.make_src(dz, t1=y, k0=kA, k1=kB, k2=kC)
        dzx = padded(torch.zeros(N, Ci, H, W, device=dev))
        ops.conv_down(dy, w, g, ops.view4(dzx),
                      ops.make_epilogue(ops.EPI_MASKSTATS, partials=P(), ticket=T(), bn=bn_in, act=x))
        dplain = padded(torch.zeros(N, Ci, H, W, device=dev))
        ops.conv_down(dy, w, g, ops.view4(dplain), ops.make_epilogue(ops.EPI_PLAIN))
        # weight gradient: small operand = activated input, big operand = dL/dy (two-tensor affine)
        gw = torch.zeros_like(w)
        part = torch.zeros(ops.wgrad_partials_len(src, dy, g), dtype=torch.float32, device=dev)
        tkt = T()
        for _ in range(2):
            ops.conv_wgrad(src, dy, g, gw, part, tkt)
        # conv-style weight gradient: small operand = dL/dy of a strided conv whose input is `y`
        wc = torch.zeros(Ci, Co, k, k, device=dev)
        sm2 = ops.make_src(dzx, t1=x1, k0=k0, k1=k1, k2=k2)
        bg2 = ops.make_src(y, k0=kA, k2=kC, relu=True)
        part2 = torch.zeros(ops.wgrad_partials_len(sm2, bg2, g), dtype=torch.float32, device=dev)
        ops.conv_wgrad(sm2, bg2, g, wc, part2, tkt)
        torch.cuda.synchronize()
        assert int(tkt.item()) == 0
        res[mask] = [t.clone() for t in (y, scr, ysig, losses, dbias, dz, dzx, scr_in, dplain, gw, wc)]
    names = ["y", "bn scratch", "sigmoid", "loss", "dbias", "dz", "dz_prev", "bn bwd coefficients", "dgrad plain",
             "wgrad (convT)", "wgrad (conv)"]
    for name, a, b_ in zip(names, res[16], res[0]):
        close(a, b_, 3e-5, name)
    # and against torch for the plain path
    xin = F.relu(x.double().cpu() * k0.double().cpu().view(1, -1, 1, 1) + k2.double().cpu().view(1, -1, 1, 1))
    ref = F.conv_transpose2d(xin, w.double().cpu(), b.double().cpu(), stride=2)
    close(res[16][0], ref, 2e-5, "y vs torch")
    close(res[16][2], torch.sigmoid(ref), 2e-5, "sigmoid vs torch")


def pitched(t, dev):
    """copy of a 4-D tensor whose rows are padded to a multiple of 4 floats (what the engines allocate)"""
    N, C, H, W = t.shape
    ld = (W + 3) // 4 * 4
    buf = torch.zeros(N, C, H, ld, dtype=torch.float32, device=dev)
    buf[..., :W] = t.to(dev)
    return buf[..., :W]


TILE = [
    # N, Cin, Cout, k, H, W: >= 2^20 positions, wide planes, few channels (BASELINE configs[3]'s last three layers, cropped)
    (9, 8, 4, 4, 350, 341),
    (5, 16, 8, 3, 470, 450),
    (3, 32, 16, 3, 600, 590),
]


@pytest.mark.parametrize("case", TILE)
def test_wgrad_tile_resident(case):
    """k_wgrad_tile (both operands staged once per CTA tile) against torch fp64 autograd and against the direct kernel it
    replaces, with the on-load transforms of the training step on both operands (BN+ReLU on the small one, the
    two-tensor BN-backward affine on the big one)."""
    from cae_tools_b200.engine import ops
    dev = torch.device("cuda")
    N, Ci, Co, k, H, W = case
    x = rnd(N, Ci, H, W, seed=1).float()
    k0, k2 = (rnd(Ci, seed=4).abs() + 0.5).float(), rnd(Ci, seed=5).float()
    Hb, Wb = 2 * H + k - 2, 2 * W + k - 2
    up, t1 = rnd(N, Co, Hb, Wb, seed=6).float(), rnd(N, Co, Hb, Wb, seed=7).float()
    b0, b1, b2 = (rnd(Co, seed=8).abs() + 0.5).float(), rnd(Co, seed=9).float(), (rnd(Co, seed=10) * 0.1).float()
    xin = F.relu(x.double() * k0.double().view(1, -1, 1, 1) + k2.double().view(1, -1, 1, 1))
    big = up.double() * b0.double().view(1, -1, 1, 1) + t1.double() * b1.double().view(1, -1, 1, 1) + \
        b2.double().view(1, -1, 1, 1)
    wd = torch.zeros(Ci, Co, k, k, dtype=torch.float64, requires_grad=True)
    (F.conv_transpose2d(xin, wd, None, stride=2) * big).sum().backward()
    xd, upd, t1d = pitched(x, dev), pitched(up, dev), pitched(t1, dev)
    src = ops.make_src(xd, k0=k0.to(dev), k2=k2.to(dev), relu=True)
    dy = ops.make_src(upd, t1=t1d, k0=b0.to(dev), k1=b1.to(dev), k2=b2.to(dev))
    g = ops.geom(k, 2, 0)
    got = {}
    for mask in (1 | 2 | 16 | 64, 1 | 2 | 16):          # with / without CAE_WGRAD_TILE (capi_host.h)
        ops.set_kernel_generation(mask << 4)
        gw = torch.full(wd.shape, float("nan"), dtype=torch.float32, device=dev)
        part = torch.zeros(ops.wgrad_partials_len(src, dy, g), dtype=torch.float32, device=dev)
        ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        for _ in range(2):
            ops.conv_wgrad(src, dy, g, gw, part, ticket)
        torch.cuda.synchronize()
        close(gw, wd.grad, 2e-5, f"wgrad mask {mask} {case}")
        assert int(ticket.item()) == 0
        got[mask] = gw.clone()
    # same launch twice gives the same bits (fixed-order reduction)
    gw2 = torch.empty_like(got[83])
    ops.set_kernel_generation(83 << 4)
    part = torch.zeros(ops.wgrad_partials_len(src, dy, g), dtype=torch.float32, device=dev)
    ops.conv_wgrad(src, dy, g, gw2, part, torch.zeros(1, dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    assert torch.equal(gw2, got[83])


DOWN_TILE = [
    # N, C_small (conv output channels), C_big (conv input channels), k, H, W of the small plane; >= 2^19 output positions
    (9, 8, 4, 4, 240, 250),
    (5, 16, 8, 3, 330, 322),
    (3, 12, 8, 3, 420, 430),
]


@pytest.mark.parametrize("mode", ["plain", "maskstats"])
@pytest.mark.parametrize("case", DOWN_TILE)
def test_down_tile_pipeline(case, mode):
    """k_down_tile (cp.async stages of raw rows, on-load affine at the shared -> register move) as the input gradient of a wide
    thin ConvTranspose2d: against torch fp64 and against the direct kernel, two-tensor BN-backward affine on the operand,
    plain and ReLU-mask + BatchNorm-backward-sums epilogues"""
    from cae_tools_b200.engine import ops
    dev = torch.device("cuda")
    N, Cs, Cb, k, H, W = case
    Hb, Wb = 2 * H + k - 2, 2 * W + k - 2
    w = (rnd(Cs, Cb, k, k, seed=2) * 0.2).float()
    up, t1 = rnd(N, Cb, Hb, Wb, seed=6).float(), rnd(N, Cb, Hb, Wb, seed=7).float()
    b0, b1, b2 = (rnd(Cb, seed=8).abs() + 0.5).float(), rnd(Cb, seed=9).float(), (rnd(Cb, seed=10) * 0.1).float()
    big = up.double() * b0.double().view(1, -1, 1, 1) + t1.double() * b1.double().view(1, -1, 1, 1) + b2.double().view(1, -1, 1, 1)
    ref = F.conv2d(big, w.double(), None, stride=2)                        # [N, Cs, H, W]
    assert ref.shape == (N, Cs, H, W)
    ysm = rnd(N, Cs, H, W, seed=11).float()
    sc, sh = (rnd(Cs, seed=12).abs() + 0.5).float(), rnd(Cs, seed=13).float() * 0.3
    mean, invstd = rnd(Cs, seed=14).float() * 0.1, (rnd(Cs, seed=15).abs() + 0.5).float()
    if mode == "maskstats":
        keepm = (ysm.double() * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)) > 0
        ref = ref * keepm
        want_db = ref.sum((0, 2, 3))
        want_dg = (ref * (ysm.double() - mean.double().view(1, -1, 1, 1)) * invstd.double().view(1, -1, 1, 1)).sum((0, 2, 3))
    upd, t1d, wd, ysd = pitched(up, dev), pitched(t1, dev), w.to(dev), pitched(ysm, dev)
    src = ops.make_src(upd, t1=t1d, k0=b0.to(dev), k1=b1.to(dev), k2=b2.to(dev))
    g = ops.geom(k, 2, 0)
    outs = {}
    for mask in (1 | 2 | 16 | 64 | 128, 1 | 2 | 16 | 64):               # with / without CAE_DOWN_TILE (capi_host.h)
        ops.set_kernel_generation(mask << 4)
        out = pitched(torch.full((N, Cs, H, W), float("nan")), dev)
        if mode == "plain":
            epi = ops.make_epilogue(ops.EPI_PLAIN)
        else:
            st = torch.zeros(7, Cs, device=dev)
            st[0], st[1], st[2], st[3] = sc.to(dev), sh.to(dev), mean.to(dev), invstd.to(dev)
            dg, db = torch.zeros(Cs, device=dev), torch.zeros(Cs, device=dev)
            bn = ops.make_bn(Cs, 1e-5, 0.1, torch.ones(Cs, device=dev), torch.zeros(Cs, device=dev), scale=st[0], shift=st[1],
                             mean=st[2], invstd=st[3], dgamma=dg, dbeta=db, bwdA=st[4], bwdB=st[5], bwdC=st[6])
            part = torch.zeros(ops.partials_len(Cs), dtype=torch.float64, device=dev)
            tick = torch.zeros(1, dtype=torch.int32, device=dev)
            epi = ops.make_epilogue(ops.EPI_MASKSTATS, partials=part, ticket=tick, bn=bn, act=ysd)
        for _ in range(2):
            ops.conv_down(src, wd, g, ops.view4(out), epi)
        torch.cuda.synchronize()
        close(out, ref, 2e-5, f"down mask {mask} {case} {mode}")
        if mode == "maskstats":
            close(db, want_db, 2e-5, "dbeta")
            close(dg, want_dg, 2e-5, "dgamma")
            assert int(tick.item()) == 0
        outs[mask] = out.clone()
    assert _maxdiff(outs[211], outs[83]) <= 2e-5


def _maxdiff(a, b):
    return float((a.double() - b.double()).abs().max()) / max(float(b.double().abs().max()), 1e-12)


UP_TILE = [
    # N, Cin, Cout, k, H, W of the input plane; >= 2^21 output pixels
    (4, 8, 4, 4, 250, 243),
    (3, 16, 8, 3, 300, 290),
    (2, 32, 16, 3, 270, 260),
    (3, 8, 3, 3, 330, 341),
]


@pytest.mark.parametrize("mode", ["stats", "sigmoid_mse"])
@pytest.mark.parametrize("case", UP_TILE)
def test_up_tile_pipeline(case, mode):
    """k_up_tile (cp.async stages of raw rows; affine + ReLU + bounds mask at the shared -> register move) as the forward pass of
    a wide thin ConvTranspose2d: against torch fp64 and against the direct kernel, with the BatchNorm-statistics epilogue
    and with the fused sigmoid + MSE (+ dL/dz) epilogue"""
    from cae_tools_b200.engine import ops
    dev = torch.device("cuda")
    N, Ci, Co, k, H, W = case
    x = rnd(N, Ci, H, W, seed=1).float()
    w = (rnd(Ci, Co, k, k, seed=2) * 0.2).float()
    b = rnd(Co, seed=3).float()
    k0, k2 = (rnd(Ci, seed=4).abs() + 0.5).float(), rnd(Ci, seed=5).float()
    xin = F.relu(x.double() * k0.double().view(1, -1, 1, 1) + k2.double().view(1, -1, 1, 1))
    v = F.conv_transpose2d(xin, w.double(), b.double(), stride=2).requires_grad_(True)
    Ho, Wo = v.shape[2], v.shape[3]
    if mode == "sigmoid_mse":
        Y = torch.rand(N, Co, Ho, Wo, generator=torch.Generator().manual_seed(64))
        loss = F.mse_loss(torch.sigmoid(v), Y.double())
        loss.backward()
    xd, wd, bd = pitched(x, dev), w.to(dev), b.to(dev)
    src = ops.make_src(xd, k0=k0.to(dev), k2=k2.to(dev), relu=True)
    g = ops.geom(k, 2, 0)
    outs = {}
    for mask in (1 | 2 | 16 | 64 | 128 | 256, 1 | 2 | 16 | 64 | 128):        # with / without CAE_UP_TILE (capi_host.h)
        ops.set_kernel_generation(mask << 4)
        out = pitched(torch.full((N, Co, Ho, Wo), float("nan")), dev)
        part = torch.zeros(ops.partials_len(Co), dtype=torch.float64, device=dev)
        tick = torch.zeros(1, dtype=torch.int32, device=dev)
        if mode == "stats":
            st = torch.zeros(7, Co, device=dev)
            bn = ops.make_bn(Co, 1e-5, 0.1, torch.ones(Co, device=dev), torch.zeros(Co, device=dev), scale=st[0], shift=st[1],
                             mean=st[2], invstd=st[3], bwdA=st[4], bwdB=st[5], bwdC=st[6])
            epi = ops.make_epilogue(ops.EPI_STATS, bias=bd, partials=part, ticket=tick, bn=bn)
        else:
            Yd = pitched(Y, dev)
            losses, dbias = torch.zeros(1, device=dev), torch.zeros(Co, device=dev)
            epi = ops.make_epilogue(ops.EPI_SIGMOID_MSE, bias=bd, partials=part, ticket=tick, target=ops.make_src(Yd),
                                    loss_out=losses, dbias=dbias, write_mode=0)
        for _ in range(2):
            ops.conv_up(src, wd, g, ops.view4(out), epi)
        torch.cuda.synchronize()
        assert int(tick.item()) == 0
        if mode == "stats":
            close(out, v, 2e-5, f"up mask {mask} {case}")
            mean = v.detach().mean((0, 2, 3))
            var = v.detach().var((0, 2, 3), unbiased=False)
            close(st[2], mean, 2e-5, "batch mean")
            close(st[3], 1.0 / torch.sqrt(var + 1e-5), 2e-5, "invstd")
        else:
            close(out, v.grad, 1e-4, f"dL/dz mask {mask} {case}")
            assert abs(float(losses[0]) - float(loss)) <= 1e-5 * float(loss)
            close(dbias, v.grad.sum(dim=(0, 2, 3)), 1e-4, "dbias")
        outs[mask] = out.clone()
    assert _maxdiff(outs[467], outs[211]) <= 2e-5
