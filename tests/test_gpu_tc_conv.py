"""ConvTranspose2d on the tensor cores (csrc/tc_conv.cu: pack -> tcgen05 3xTF32 GEMM -> col2im / unpack) against torch fp64
(F.conv_transpose2d and its autograd gradients) and against the SIMT kernels it replaces.  Tolerance 2e-5 of the max-norm
(north_star: 1e-4 per layer)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-5

# N, Cin, Cout, k, s, op, H, W
CASES = [
    (2, 64, 32, 3, 2, 0, 3, 3),
    (3, 128, 64, 3, 2, 0, 7, 7),
    (2, 36, 12, 3, 2, 0, 5, 6),        # ragged channel counts (K and N tails inside one tile)
    (2, 256, 40, (4, 3), 2, 0, 5, 4),  # tuple kernel
    (2, 64, 8, 4, 2, 0, 6, 6),
    (2, 48, 16, 3, 2, 1, 5, 5),        # output_padding
    (5, 160, 136, 3, 2, 0, 9, 9),      # several M / N tiles, split-K weight gradient
    (128, 1024, 512, 3, 2, 0, 3, 3),   # BASELINE configs[3], first decoder layer at the full batch
    (128, 64, 32, 3, 2, 0, 63, 63),    # ... and the last tensor-core layer (508 k positions: 248 split-K slices)
    # wide planes -> the tile kernels of the pack passes (k_tc_im2col_tile: Win >= 24, k_tc_col2im_tile: Wout >= 48):
    (3, 64, 40, 3, 2, 1, 9, 30),       # partial channel chunk (32 + 8), output_padding, one ragged 30-position tile per row
    (2, 32, 8, 4, 2, 0, 6, 40),        # 4-wide kernel (every output pixel has 2 x 2 taps), two position tiles per row (32 + 8)
    (2, 96, 36, (3, 4), 2, 0, 5, 33),  # tuple kernel, channel chunks 32 + 4, tiles 32 + 1
]


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64) * scale


def close(a, b, tol=TOL, what=""):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)
    assert err <= tol, f"{what}: rel err {err:.3e} > {tol}"


@pytest.mark.parametrize("case", CASES)
def test_tc_convt_forward_backward(case):
    from cae_tools_b200.engine import ops
    dev = torch.device("cuda")
    N, Ci, Co, k, s, op, H, W = case
    kh, kw = (k, k) if isinstance(k, int) else k
    assert ops.tc_convT_supported(Ci, Co, k, s, 0)
    x = rnd(N, Ci, H, W, seed=1).float()
    w = rnd(Ci, Co, kh, kw, seed=2, scale=0.2).float()
    b = rnd(Co, seed=3).float()
    k0, k2 = rnd(Ci, seed=4).float(), rnd(Ci, seed=5).float()
    sh = (1, Ci, 1, 1)
    # forward: the operand is relu(k0*x + k2) (BatchNorm + ReLU applied on load), bias added, BN statistics in the epilogue
    xin = F.relu(x.double() * k0.double().view(sh) + k2.double().view(sh)).requires_grad_(True)
    wd64 = w.double().requires_grad_(True)
    ref = F.conv_transpose2d(xin, wd64, b.double(), stride=s, output_padding=op)
    Ho, Wo = ref.shape[2], ref.shape[3]
    desc = ops.make_tc_conv(Ci, Co, k, s, N, H, W, Ho, Wo, dev)
    out = torch.full(ref.shape, float("nan"), dtype=torch.float32, device=dev)
    src = ops.make_src(x.to(dev), k0=k0.to(dev), k2=k2.to(dev), relu=True)
    bn_s = torch.zeros(7, Co, device=dev)
    gamma, beta = torch.ones(Co, device=dev), torch.zeros(Co, device=dev)
    blk = ops.make_bn(Co, gamma=gamma, beta=beta, scale=bn_s[0], shift=bn_s[1], mean=bn_s[2], invstd=bn_s[3])
    part = torch.zeros(ops.partials_len(Co), dtype=torch.float64, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    epi = ops.make_epilogue(ops.EPI_STATS, bias=b.to(dev), partials=part, ticket=ticket, bn=blk)
    ops.tc_convT_fwd(desc, src, w.to(dev), ops.view4(out), epi)
    torch.cuda.synchronize()
    close(out, ref, what=f"tc forward {case}")
    close(bn_s[2], ref.mean(dim=(0, 2, 3)), tol=1e-4, what="BN mean from the epilogue")
    close(bn_s[3], 1.0 / torch.sqrt(ref.var(dim=(0, 2, 3), unbiased=False) + 1e-5), tol=1e-4, what="BN invstd")
    # ... and the SIMT kernel it replaces gives the same tensor
    out_simt = torch.empty_like(out)
    ops.conv_up(src, w.to(dev), ops.geom(k, s, 0), ops.view4(out_simt), ops.make_epilogue(ops.EPI_PLAIN, bias=b.to(dev)))
    torch.cuda.synchronize()
    close(out, out_simt, what="tc vs SIMT forward")

    # backward: dy arrives as the affine a*dz + bb*y + cc (BatchNorm-backward applied on load)
    dz, yy = rnd(N, Co, Ho, Wo, seed=6).float(), rnd(N, Co, Ho, Wo, seed=7).float()
    ka, kb, kc = rnd(Co, seed=8).float(), rnd(Co, seed=9, scale=0.1).float(), rnd(Co, seed=10, scale=0.1).float()
    so = (1, Co, 1, 1)
    dy = dz.double() * ka.double().view(so) + yy.double() * kb.double().view(so) + kc.double().view(so)
    gx, gw = torch.autograd.grad(ref, (xin, wd64), dy)
    dsrc = ops.make_src(dz.to(dev), t1=yy.to(dev), k0=ka.to(dev), k1=kb.to(dev), k2=kc.to(dev))
    ops.tc_convT_im2col(desc, dsrc)
    dx = torch.full((N, Ci, H, W), float("nan"), dtype=torch.float32, device=dev)
    ops.tc_convT_dgrad(desc, w.to(dev), ops.view4(dx), ops.make_epilogue(ops.EPI_PLAIN))
    grad = torch.full((Ci, Co, kh, kw), float("nan"), dtype=torch.float32, device=dev)
    ops.tc_convT_wgrad(desc, grad)
    torch.cuda.synchronize()
    close(dx, gx, what=f"tc input gradient {case}")
    close(grad, gw, what=f"tc weight gradient {case}")

    # input gradient with the ReLU-mask + BatchNorm-backward-sums epilogue against the SIMT path
    act = rnd(N, Ci, H, W, seed=11).float().to(dev)
    sc = torch.zeros(7, Ci, device=dev)
    sc[0], sc[1], sc[2], sc[3] = 1.3, 0.1, 0.05, 0.9

    def masked(fn):
        o = torch.full((N, Ci, H, W), float("nan"), dtype=torch.float32, device=dev)
        s7 = sc.clone()
        g2, b2, dg, db = torch.ones(Ci, device=dev), torch.zeros(Ci, device=dev), torch.zeros(Ci, device=dev), torch.zeros(Ci, device=dev)
        bl = ops.make_bn(Ci, gamma=g2, beta=b2, scale=s7[0], shift=s7[1], mean=s7[2], invstd=s7[3], dgamma=dg, dbeta=db,
                         bwdA=s7[4], bwdB=s7[5], bwdC=s7[6])
        p2 = torch.zeros(ops.partials_len(Ci), dtype=torch.float64, device=dev)
        t2 = torch.zeros(1, dtype=torch.int32, device=dev)
        e = ops.make_epilogue(ops.EPI_MASKSTATS, partials=p2, ticket=t2, bn=bl, act=act)
        fn(ops.view4(o), e)
        torch.cuda.synchronize()
        return o, s7, dg, db

    o_tc, s_tc, dg_tc, db_tc = masked(lambda o, e: ops.tc_convT_dgrad(desc, w.to(dev), o, e))
    o_si, s_si, dg_si, db_si = masked(lambda o, e: ops.conv_down(dsrc, w.to(dev), ops.geom(k, s, 0), o, e))
    close(o_tc, o_si, what="masked input gradient tc vs SIMT")
    close(dg_tc, dg_si, tol=1e-4, what="dgamma")
    close(db_tc, db_si, tol=1e-4, what="dbeta")
    close(s_tc[4:7], s_si[4:7], tol=1e-4, what="BN-backward coefficients")
