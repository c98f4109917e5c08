"""Thin Python wrappers over the C ABI: build the POD descriptors from torch tensors and enqueue
kernels on torch's current CUDA stream (so everything is capturable in a CUDA graph).

Nothing here computes; if the library is missing, `lib()` raises.
"""

from __future__ import annotations

import ctypes as C

import torch

from .._lib import (CaeDpPeers, CaeStemTrain, CaeStemTrainConv, CaeStemTrainFc, CaeStemTrainUp, CaeTcConv, CaeTcGemm, CaeBN, CaeConvGeom, CaeEpilogue, CaeFcStack, CaeGemm, CaePatchHead, CaeSrc, CaeStemConv, CaeStemFc, CaeStemUp,
                    CaeUnetStem, CaeView, STEM_MAX, EPI_MASK, EPI_MASKSTATS, EPI_PLAIN,
                    EPI_SIGMOID, EPI_SIGMOID_MSE, EPI_STATS, check, lib)

__all__ = ["view4", "make_src", "make_bn", "make_epilogue", "geom", "conv_up", "conv_down", "conv_wgrad",
           "wgrad_partials_len", "ew_epilogue", "gemm", "bn_eval_prepare", "mse", "adam", "step_advance",
           "partials_len", "EPI_PLAIN", "EPI_STATS", "EPI_MASKSTATS", "EPI_SIGMOID", "EPI_SIGMOID_MSE"]


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "device tensor expected"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def view4(t: torch.Tensor, n=None) -> CaeView:
    """CaeView of a 4-D fp32 CUDA tensor (innermost stride must be 1); `n` restricts the batch."""
    assert t.dim() == 4 and t.dtype == torch.float32 and t.is_cuda, (t.shape, t.dtype, t.device)
    sn, sc, sh, sw = t.stride()
    N, Cc, H, W = t.shape
    assert sw == 1 or W == 1
    if H == 1:
        sh = max(W, 1)
    v = CaeView(t.data_ptr(), int(N if n is None else n), int(Cc), int(H), int(W), int(sh), int(sc), int(sn))
    v._keep = t   # descriptors hold raw pointers: keep the tensor alive as long as the descriptor
    return v


def make_src(t0, t1=None, k0=None, k1=None, k2=None, relu=False, cursor=None, cursor_stride=0, n=None, kn=None) -> CaeSrc:
    if t1 is not None:
        assert t1.shape == t0.shape and t1.stride() == t0.stride(), "t1 must share t0's geometry"
    s = CaeSrc(view4(t0, n), _ptr(t1), _ptr(k0), _ptr(k1), _ptr(k2), int(bool(relu)), _ptr(cursor),
               int(cursor_stride), _ptr(kn))
    s._keep = (t0, t1, k0, k1, k2, cursor, kn)
    return s


def make_bn(Cn, eps=1e-5, momentum=0.1, gamma=None, beta=None, running_mean=None, running_var=None, nbt=None,
            scale=None, shift=None, mean=None, invstd=None, dgamma=None, dbeta=None, dbias=None, bwdA=None,
            bwdB=None, bwdC=None) -> CaeBN:
    b = CaeBN(int(Cn), float(eps), float(momentum), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
              _ptr(nbt), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), _ptr(dgamma), _ptr(dbeta),
              _ptr(dbias), _ptr(bwdA), _ptr(bwdB), _ptr(bwdC))
    b._keep = (gamma, beta, running_mean, running_var, nbt, scale, shift, mean, invstd, dgamma, dbeta, dbias, bwdA,
               bwdB, bwdC)
    return b


_NULL_VIEW = CaeView(None, 0, 0, 0, 0, 0, 0, 0)
_NULL_BN = CaeBN()


def make_epilogue(mode, bias=None, partials=None, ticket=None, bn=None, act=None, target=None, loss_out=None,
                  dbias=None, write_mode=0, n=None, count_scale=1.0, addend=None) -> CaeEpilogue:
    e = CaeEpilogue()
    e.mode = int(mode)
    e.bias = _ptr(bias)
    e.partials = _ptr(partials)
    e.ticket = _ptr(ticket)
    if bn is not None:
        e.bn = bn
    if act is not None:
        e.act = view4(act, n)
    if target is not None:
        e.target = target
    e.loss_out = _ptr(loss_out)
    e.dbias = _ptr(dbias)
    e.write_mode = int(write_mode)
    e.count_scale = float(count_scale)
    if addend is not None:
        e.addend = addend
    e._keep = (bias, partials, ticket, bn, act, target, loss_out, dbias, addend)
    return e


def geom(kernel, stride, pad=0) -> CaeConvGeom:
    if isinstance(kernel, (tuple, list)):
        kh, kw = int(kernel[0]), int(kernel[1])
    else:
        kh = kw = int(kernel)
    return CaeConvGeom(kh, kw, int(stride), int(pad))


def partials_len(Cn: int) -> int:
    return int(lib().cae_partials_len(int(Cn)))


def conv_up(src: CaeSrc, weight, g: CaeConvGeom, out: CaeView, epi: CaeEpilogue):
    check(lib().cae_conv_up(C.byref(src), _ptr(weight), C.byref(g), C.byref(out), C.byref(epi), _stream()),
          "cae_conv_up")


def conv_down(src: CaeSrc, weight, g: CaeConvGeom, out: CaeView, epi: CaeEpilogue):
    check(lib().cae_conv_down(C.byref(src), _ptr(weight), C.byref(g), C.byref(out), C.byref(epi), _stream()),
          "cae_conv_down")


def wgrad_partials_len(small: CaeSrc, big: CaeSrc, g: CaeConvGeom) -> int:
    n = int(lib().cae_wgrad_partials_len(C.byref(small), C.byref(big), C.byref(g)))
    if n < 0:
        check(-1, "cae_wgrad_partials_len")
    return n


def conv_wgrad(small: CaeSrc, big: CaeSrc, g: CaeConvGeom, grad, partials, ticket):
    check(lib().cae_conv_wgrad(C.byref(small), C.byref(big), C.byref(g), _ptr(grad), _ptr(partials), _ptr(ticket),
                               _stream()), "cae_conv_wgrad")


def ew_epilogue(src: CaeSrc, out: CaeView, epi: CaeEpilogue):
    check(lib().cae_ew_epilogue(C.byref(src), C.byref(out), C.byref(epi), _stream()), "cae_ew_epilogue")


USE_TC_DENSE = True       # tests switch it off to compare against the SIMT kernel
_gemm_ws = {}


def gemm(M, N, K, A, sAm, sAk, B, sBk, sBn, Cout, sCm, sCn, a_k0=None, a_k2=None, a_hw=1, a_relu=False, b_k0=None,
         b_k2=None, b_hw=1, b_relu=False, bias=None, relu_out=False, mask=None, rowsum_A=None):
    g = CaeGemm(int(M), int(N), int(K), _ptr(A), int(sAm), int(sAk), _ptr(B), int(sBk), int(sBn), _ptr(Cout),
                int(sCm), int(sCn), _ptr(a_k0), _ptr(a_k2), int(a_hw), int(bool(a_relu)), _ptr(b_k0), _ptr(b_k2),
                int(b_hw), int(bool(b_relu)), _ptr(bias), int(bool(relu_out)), _ptr(mask), _ptr(rowsum_A))
    # genuinely dense contractions (both free dimensions >= 128: the large-fc regimes) go to the tensor cores; the scratch
    # for the split operands and the split-K slices is owned per call site (keyed by the output: weight-gradient GEMMs run
    # on side streams beside the input-gradient ones) and allocated on the eager warm-up pass that precedes graph capture
    need = lib().cae_gemm_tc_workspace(C.byref(g)) if USE_TC_DENSE else 0
    if need > 0:
        key = (_ptr(Cout), int(M), int(N), int(K))
        ws = _gemm_ws.get(key)
        if ws is None or ws.numel() < need:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("cae_gemm_tc: workspace must exist before graph capture (run the op once eagerly)")
            ws = _gemm_ws[key] = torch.empty(need, dtype=torch.float32, device=Cout.device)
        check(lib().cae_gemm_tc(C.byref(g), _ptr(ws), int(ws.numel()), _stream()), "cae_gemm_tc")
        return
    check(lib().cae_gemm(C.byref(g), _stream()), "cae_gemm")


def bn_eval_prepare(device_table: torch.Tensor, count: int):
    check(lib().cae_bn_eval_prepare(_ptr(device_table), int(count), _stream()), "cae_bn_eval_prepare")


def bn_table(blocks, device) -> torch.Tensor:
    """Pack CaeBN PODs into a device byte tensor for cae_bn_eval_prepare."""
    raw = b"".join(bytes(b) for b in blocks)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
    return t


def mse(a, b, n, partials, ticket, loss_out, cursor=None):
    check(lib().cae_mse(_ptr(a), _ptr(b), int(n), _ptr(partials), _ptr(ticket), _ptr(loss_out), _ptr(cursor),
                        _stream()), "cae_mse")


def adam(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled, grad_scale, step_count):
    check(lib().cae_adam(_ptr(p), _ptr(g), _ptr(m), _ptr(v), int(n), float(lr), float(beta1), float(beta2),
                         float(eps), float(weight_decay), int(bool(decoupled)), float(grad_scale), _ptr(step_count),
                         _stream()), "cae_adam")


def adam_advance(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled, grad_scale, step_count, cursor, n_batches,
                 ticket):
    check(lib().cae_adam_advance(_ptr(p), _ptr(g), _ptr(m), _ptr(v), int(n), float(lr), float(beta1), float(beta2),
                                 float(eps), float(weight_decay), int(bool(decoupled)), float(grad_scale), _ptr(step_count),
                                 _ptr(cursor), int(n_batches), _ptr(ticket), _stream()), "cae_adam_advance")


def step_advance(step_count, cursor, n_batches):
    check(lib().cae_step_advance(_ptr(step_count), _ptr(cursor), int(n_batches), _stream()), "cae_step_advance")


def vae_reparam_fwd(mu, logvar, eps, eps_stride, cursor, z, n_samples, latent, sample, kl_scale, kl_out):
    check(lib().cae_vae_reparam_fwd(_ptr(mu), _ptr(logvar), _ptr(eps), int(eps_stride), _ptr(cursor), _ptr(z),
                                    int(n_samples), int(latent), int(bool(sample)), float(kl_scale), _ptr(kl_out),
                                    _stream()), "cae_vae_reparam_fwd")


def vae_reparam_bwd(dz, mu, logvar, eps, eps_stride, cursor, dmu, dlogvar, n_samples, latent, kl_weight):
    check(lib().cae_vae_reparam_bwd(_ptr(dz), _ptr(mu), _ptr(logvar), _ptr(eps), int(eps_stride), _ptr(cursor),
                                    _ptr(dmu), _ptr(dlogvar), int(n_samples), int(latent), float(kl_weight),
                                    _stream()), "cae_vae_reparam_bwd")


def add2(a, b, out, n):
    check(lib().cae_add2(_ptr(a), _ptr(b), _ptr(out), int(n), _stream()), "cae_add2")


def randn(out, n, seed, step_count):
    check(lib().cae_randn(_ptr(out), int(n), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(step_count), _stream()), "cae_randn")


def set_kernel_generation(gen: int):
    """1: generic direct kernels only; 2 (default): tiled shared-memory kernels where they apply"""
    lib().cae_set_kernel_generation(int(gen))


# ---- UNET pieces -------------------------------------------------------------------------------------
def plane_stats(y_view: CaeView, stats):
    check(lib().cae_plane_stats(C.byref(y_view), _ptr(stats), _stream()), "cae_plane_stats")


def channel_attention_fwd(stats, W1, W2, N, Cn, Cr, HW, att, hid):
    check(lib().cae_channel_attention_fwd(_ptr(stats), _ptr(W1), _ptr(W2), int(N), int(Cn), int(Cr), int(HW), _ptr(att),
                                          _ptr(hid), _stream()), "cae_channel_attention_fwd")


def channel_attention_bwd(datt, att, hid, stats, W1, W2, N, Cn, Cr, HW, dW1, dW2, davg, dmax):
    check(lib().cae_channel_attention_bwd(_ptr(datt), _ptr(att), _ptr(hid), _ptr(stats), _ptr(W1), _ptr(W2), int(N),
                                          int(Cn), int(Cr), int(HW), _ptr(dW1), _ptr(dW2), _ptr(davg), _ptr(dmax),
                                          _stream()), "cae_channel_attention_bwd")


def plane_dot(g: CaeSrc, y_view: CaeView, out):
    check(lib().cae_plane_dot(C.byref(g), C.byref(y_view), _ptr(out), _stream()), "cae_plane_dot")


def gate_bwd(g: CaeSrc, att, davg, dmax, stats, dy_view: CaeView, plane_sum):
    check(lib().cae_gate_bwd(C.byref(g), _ptr(att), _ptr(davg), _ptr(dmax), _ptr(stats), C.byref(dy_view),
                             _ptr(plane_sum), _stream()), "cae_gate_bwd")


def sum_over_n(inp, N, Cn, out):
    check(lib().cae_sum_over_n(_ptr(inp), int(N), int(Cn), _ptr(out), _stream()), "cae_sum_over_n")


def masked_pearson_loss(pred: CaeView, target: CaeSrc, mask, mask_channels, lambda_pearson, count_scale, moments, coef,
                        scalars, loss_out, pearson_out, dz=None, plane_sum=None, mse_scale=None):
    check(lib().cae_masked_pearson_loss(C.byref(pred), C.byref(target), C.byref(mask) if mask is not None else None,
                                        int(mask_channels), float(lambda_pearson), float(count_scale), _ptr(moments),
                                        _ptr(coef), _ptr(scalars), _ptr(loss_out), _ptr(pearson_out),
                                        C.byref(dz) if dz is not None else None, _ptr(plane_sum), _ptr(mse_scale), _stream()),
          "cae_masked_pearson_loss")


# ---- patch head (kernel == stride transposed conv + sigmoid + masked MSE / Pearson loss) -------------------------
def patch_head_supported(K, stride, pad, Cin, Win) -> bool:
    return bool(lib().cae_patch_head_supported(int(K), int(stride), int(pad), int(Cin), int(Win)))


def make_patch_head(src: CaeSrc, weight, bias, K, Cout, target=None, mask=None, mask_channels=0, lambda_pearson=0.0,
                    count_scale=1.0, moments=None, coef=None, scalars=None, loss_out=None, pearson_out=None,
                    ticket=None, mse_scale=None) -> CaePatchHead:
    h = CaePatchHead()
    h.inp = src
    h.weight = _ptr(weight)
    h.bias = _ptr(bias)
    h.K = int(K)
    h.Cout = int(Cout)
    if target is not None:
        h.target = target
    if mask is not None:
        h.mask = mask
    h.mask_channels = int(mask_channels)
    h.lambda_pearson = float(lambda_pearson)
    h.count_scale = float(count_scale)
    h.moments, h.coef, h.scalars = _ptr(moments), _ptr(coef), _ptr(scalars)
    h.loss_out, h.pearson_out = _ptr(loss_out), _ptr(pearson_out)
    h.ticket = _ptr(ticket)
    h.mse_scale = _ptr(mse_scale)
    h._keep = (src, weight, bias, target, mask, moments, coef, scalars, loss_out, pearson_out, ticket, mse_scale)
    return h


def patch_head_fwd(h: CaePatchHead, yhat: CaeView = None):
    check(lib().cae_patch_head_fwd(C.byref(h), C.byref(yhat) if yhat is not None else None, _stream()),
          "cae_patch_head_fwd")


def patch_head_partials_len(h: CaePatchHead) -> int:
    n = int(lib().cae_patch_head_partials_len(C.byref(h)))
    if n < 0:
        check(-1, "cae_patch_head_partials_len")
    return n


def patch_head_bwd(h: CaePatchHead, din: CaeView, epi: CaeEpilogue, partials):
    check(lib().cae_patch_head_bwd(C.byref(h), C.byref(din), C.byref(epi), _ptr(partials), _stream()),
          "cae_patch_head_bwd")


def patch_head_wgrad_reduce(h: CaePatchHead, grad_w, grad_b, partials):
    check(lib().cae_patch_head_wgrad_reduce(C.byref(h), _ptr(grad_w), _ptr(grad_b), _ptr(partials), _stream()),
          "cae_patch_head_wgrad_reduce")


# ---- fused attention block ---------------------------------------------------------------------------------------
def attention_block_supported(Cn, H, W, Cr) -> bool:
    return bool(lib().cae_attention_block_supported(int(Cn), int(H), int(W), int(Cr)))


def attention_block_partials_len(Cn, Cr) -> int:
    return int(lib().cae_attention_block_partials_len(int(Cn), int(Cr)))


def attention_block_fwd(y: CaeView, skip: CaeSrc, W1, W2, Cr, cat: CaeView, epi: CaeEpilogue, stats, att, hid):
    check(lib().cae_attention_block_fwd(C.byref(y), C.byref(skip), _ptr(W1), _ptr(W2), int(Cr), C.byref(cat),
                                        C.byref(epi), _ptr(stats), _ptr(att), _ptr(hid), _stream()),
          "cae_attention_block_fwd")


def attention_block_bwd(g: CaeSrc, y: CaeView, att, hid, stats, W1, W2, Cr, dy: CaeView, dW1, dW2, dbias, partials,
                        ticket):
    check(lib().cae_attention_block_bwd(C.byref(g), C.byref(y), _ptr(att), _ptr(hid), _ptr(stats), _ptr(W1), _ptr(W2),
                                        int(Cr), C.byref(dy), _ptr(dW1), _ptr(dW2), _ptr(dbias), _ptr(partials),
                                        _ptr(ticket), _stream()), "cae_attention_block_bwd")


# ---- eval-mode UNET stem ---------------------------------------------------------------------------------------------
def make_unet_stem(convs, fcs, ups) -> CaeUnetStem:
    """convs: [(Cin,Hin,Win,Cout,Hout,Wout,k,stride,pad, w,b,scale,shift)], fcs: [(in,out,relu, w,b,scale,shift)],
    ups: [(Cin,Hin,Win,Cout,Hout,Wout,k,stride,pad,Cr,skip, w,b,W1,W2,scale,shift)]; tensors or None"""
    if max(len(convs), len(fcs), len(ups)) > STEM_MAX:
        return None
    st = CaeUnetStem()
    st.n_conv, st.n_fc, st.n_up = len(convs), len(fcs), len(ups)
    keep = []
    for i, c in enumerate(convs):
        st.conv[i] = CaeStemConv(*[int(v) for v in c[:9]], *[_ptr(t) for t in c[9:]])
        keep += list(c[9:])
    for i, c in enumerate(fcs):
        st.fc[i] = CaeStemFc(*[int(v) for v in c[:3]], *[_ptr(t) for t in c[3:]])
        keep += list(c[3:])
    for i, c in enumerate(ups):
        st.up[i] = CaeStemUp(*[int(v) for v in c[:11]], *[_ptr(t) for t in c[11:]])
        keep += list(c[11:])
    st._keep = keep
    return st


def unet_stem_supported(stem: CaeUnetStem) -> bool:
    return stem is not None and bool(lib().cae_unet_stem_supported(C.byref(stem)))


def unet_stem_eval(stem: CaeUnetStem, x: CaeSrc, out: CaeView):
    check(lib().cae_unet_stem_eval(C.byref(stem), C.byref(x), C.byref(out), _stream()), "cae_unet_stem_eval")


# ---- fc bottleneck in one launch -----------------------------------------------------------------------------------------
def fc_stack_supported(N, in1, fc1, lat, fc2, out4) -> bool:
    return bool(lib().cae_fc_stack_supported(int(N), int(in1), int(fc1), int(lat), int(fc2), int(out4)))


def make_fc_stack(N, A, lins, t1, z, t3, u, a_k0=None, a_k2=None, a_hw=1, a_relu=False, bn1=None, bn3=None, train=False,
                  relu_mid=False, du=None, grads=None, dA=None) -> CaeFcStack:
    """lins: the four nn.Linear modules; grads: callable parameter -> gradient view (backward descriptor only)"""
    p = CaeFcStack()
    l1, l2, l3, l4 = lins
    p.N, p.in1, p.fc1, p.lat, p.fc2, p.out4 = int(N), l1.in_features, l1.out_features, l2.out_features, \
        l3.out_features, l4.out_features
    p.A, p.a_k0, p.a_k2, p.a_hw, p.a_relu = _ptr(A), _ptr(a_k0), _ptr(a_k2), int(a_hw), int(bool(a_relu))
    for i, l in enumerate(lins, 1):
        setattr(p, f"W{i}", _ptr(l.weight))
        setattr(p, f"b{i}", _ptr(l.bias))
    if bn1 is not None:
        p.bn1, p.bn3 = bn1, bn3
    p.train, p.relu_mid = int(bool(train)), int(bool(relu_mid))
    p.t1, p.z, p.t3, p.u = _ptr(t1), _ptr(z), _ptr(t3), _ptr(u)
    if grads is not None:
        p.du, p.dA = _ptr(du), _ptr(dA)
        for i, l in enumerate(lins, 1):
            setattr(p, f"dW{i}", _ptr(grads(l.weight)))
            setattr(p, f"db{i}", _ptr(grads(l.bias)))
    p._keep = (A, a_k0, a_k2, lins, bn1, bn3, t1, z, t3, u, du, dA)
    return p


def fc_stack_fwd(p: CaeFcStack):
    check(lib().cae_fc_stack_fwd(C.byref(p), _stream()), "cae_fc_stack_fwd")


def fc_stack_bwd(p: CaeFcStack):
    check(lib().cae_fc_stack_bwd(C.byref(p), _stream()), "cae_fc_stack_bwd")


# ---- tensor-core GEMM (tcgen05 / TMEM / TMA) ---------------------------------------------------------------------------
def tc_split(x, hi, lo):
    """hi = x rounded to TF32, lo = x - hi rounded to TF32 (operands of the 3xTF32 GEMM; the pair drops <= 2^-22 |x|)"""
    check(lib().cae_tc_split(_ptr(x), _ptr(hi), _ptr(lo), int(x.numel()), _stream()), "cae_tc_split")


def tc_gemm(M, N, K, a_hi, a_lo, lda, a_mn, b_hi, b_lo, ldb, b_mn, Cout, ldc, splits=1, split_stride=0, tile_n=128):
    """C[m,n] = sum_k A[m,k] B[n,k]; operands K-major (x[row*ld + k]) or MN-major (x[k*ld + row]); lo=None: 1xTF32"""
    g = CaeTcGemm(int(M), int(N), int(K), _ptr(a_hi), _ptr(a_lo), int(lda), int(bool(a_mn)), _ptr(b_hi), _ptr(b_lo),
                  int(ldb), int(bool(b_mn)), _ptr(Cout), int(ldc), int(splits), int(split_stride), int(tile_n))
    check(lib().cae_tc_gemm(C.byref(g), _stream()), "cae_tc_gemm")


# ---- ConvTranspose2d on the tensor cores ---------------------------------------------------------------------------------
def tc_convT_supported(Cin, Cout, kernel, stride, pad) -> bool:
    kh, kw = (kernel if isinstance(kernel, (tuple, list)) else (kernel, kernel))
    return bool(lib().cae_tc_convt_supported(int(Cin), int(Cout), int(kh), int(kw), int(stride), int(pad)))


def make_tc_conv(Cin, Cout, kernel, stride, N, Hin, Win, Hout, Wout, device, shared=None) -> CaeTcConv:
    """descriptor + workspaces of one layer; `shared` (dict) carries the scratch buffers layers can share
    (weight operand, GEMM output, im2col operand): they are grown to the largest request"""
    kh, kw = (kernel if isinstance(kernel, (tuple, list)) else (kernel, kernel))
    T = kh * kw
    lda = (Cin + 3) // 4 * 4
    ldn = (T * Cout + 3) // 4 * 4
    P = N * Hin * Win
    c = CaeTcConv()
    c.Cin, c.Cout, c.kh, c.kw, c.stride = int(Cin), int(Cout), int(kh), int(kw), int(stride)
    c.N, c.Hin, c.Win, c.Hout, c.Wout = int(N), int(Hin), int(Win), int(Hout), int(Wout)
    c.lda, c.ldn = lda, ldn
    a = torch.zeros(2, P, lda, dtype=torch.float32, device=device)
    c.a_hi, c.a_lo = a[0].data_ptr(), a[1].data_ptr()
    shared = shared if shared is not None else {}
    splits = int(lib().cae_tc_convt_wgrad_splits(C.byref(c)))
    need = {"w": 2 * max(T * Cout * lda, Cin * ldn), "cols": max(P * max(ldn, lda), splits * Cin * ldn), "dcols": 2 * P * ldn}
    for k, n in need.items():
        if k not in shared or shared[k].numel() < n:
            shared[k] = torch.zeros(n, dtype=torch.float32, device=device)
    c._keep = (a, shared)
    c._shared = shared
    return c


def _tc_bind(c: CaeTcConv):
    """(re)point the descriptor at the current shared scratch (it may have grown after this layer was described)"""
    sh = c._shared
    w, cols, dcols = sh["w"], sh["cols"], sh["dcols"]
    c.w_hi, c.w_lo = w.data_ptr(), w.data_ptr() + (w.numel() // 2) * 4
    c.cols, c.cols_len = cols.data_ptr(), cols.numel()
    c.dcols_hi, c.dcols_lo = dcols.data_ptr(), dcols.data_ptr() + (dcols.numel() // 2) * 4
    return c


def tc_convT_fwd(c: CaeTcConv, src: CaeSrc, weight, out: CaeView, epi: CaeEpilogue):
    check(lib().cae_tc_convt_fwd(C.byref(_tc_bind(c)), C.byref(src), _ptr(weight), C.byref(out), C.byref(epi), _stream()),
          "cae_tc_convt_fwd")


def tc_convT_im2col(c: CaeTcConv, dy: CaeSrc):
    check(lib().cae_tc_convt_im2col(C.byref(_tc_bind(c)), C.byref(dy), _stream()), "cae_tc_convt_im2col")


def tc_convT_dgrad(c: CaeTcConv, weight, dx: CaeView, epi: CaeEpilogue):
    check(lib().cae_tc_convt_dgrad(C.byref(_tc_bind(c)), _ptr(weight), C.byref(dx), C.byref(epi), _stream()),
          "cae_tc_convt_dgrad")


def tc_convT_wgrad(c: CaeTcConv, grad):
    check(lib().cae_tc_convt_wgrad(C.byref(_tc_bind(c)), _ptr(grad), _stream()), "cae_tc_convt_wgrad")


# ---- training-mode UNET stem (two cooperative launches) ---------------------------------------------------------------------
def make_stem_train(N, convs, fcs, ups, dropout_p, seed, step_count, device, params) -> CaeStemTrain:
    """convs: [(Cin,Hin,Win,Cout,Hout,Wout,k,stride,pad, w,b,dw,db, CaeBN)], fcs: [(in,out,has_bn, w,b,dw,db, CaeBN|None)],
    ups: [(Cin,Hin,Win,Cout,Hout,Wout,k,stride,pad,Cr,skip, w,b,W1,W2,dw,db,dW1,dW2, CaeBN)].  Returns None when the geometry
    or the batch does not fit the fused kernels; otherwise the descriptor with its tape / hin / dhin / workspaces allocated
    (attributes .t_tape, .t_hin, .t_dhin)."""
    if max(len(convs), len(fcs), len(ups)) > STEM_MAX:
        return None
    st = CaeStemTrain()
    st.n_conv, st.n_fc, st.n_up, st.N = len(convs), len(fcs), len(ups), int(N)
    keep = []
    for i, c in enumerate(convs):
        e = CaeStemTrainConv(*[int(v) for v in c[:9]], *[_ptr(t) for t in c[9:13]])
        e.bn = c[13]
        st.conv[i] = e
        keep += list(c[9:])
    for i, c in enumerate(fcs):
        e = CaeStemTrainFc(*[int(v) for v in c[:3]], *[_ptr(t) for t in c[3:7]])
        if c[7] is not None:
            e.bn = c[7]
        st.fc[i] = e
        keep += list(c[3:])
    for i, c in enumerate(ups):
        e = CaeStemTrainUp(*[int(v) for v in c[:11]], *[_ptr(t) for t in c[11:19]])
        e.bn = c[19]
        st.up[i] = e
        keep += list(c[11:])
    st.params, st.params_len = params.data_ptr(), int(params.numel())     # contiguous block holding every stem parameter
    st.dropout_p = float(dropout_p)
    st.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    st.step_count = _ptr(step_count)
    # the support check wants non-null workspaces: size them first with placeholders
    c_last, h_last, w_last = ups[-1][3], ups[-1][4], ups[-1][5]
    hin = torch.zeros(N, 2 * c_last, h_last, w_last, dtype=torch.float32, device=device)
    dhin = torch.zeros_like(hin)
    st.hin, st.dhin = hin.data_ptr(), dhin.data_ptr()
    if not lib().cae_unet_stem_train_supported(C.byref(st)):
        return None
    tape = torch.zeros(N, int(lib().cae_unet_stem_train_tape_elems(C.byref(st))), dtype=torch.float32, device=device)
    bnpart = torch.zeros(int(lib().cae_unet_stem_train_workspace(C.byref(st), 0)), dtype=torch.float64, device=device)
    wpart = torch.zeros(int(lib().cae_unet_stem_train_workspace(C.byref(st), 1)), dtype=torch.float32, device=device)
    st.tape, st.bnpart, st.wpart = tape.data_ptr(), bnpart.data_ptr(), wpart.data_ptr()
    st.t_tape, st.t_hin, st.t_dhin = tape, hin, dhin
    # per-sample tape layout (mirror of st_plan in unet_stem_train.cu; every slot rounded up to 4 floats): for inspection
    off, lay = 0, {}

    def take(name, shape):
        nonlocal off
        n = 1
        for d in shape:
            n *= int(d)
        lay[name] = (off, tuple(int(d) for d in shape))
        off += (n + 3) // 4 * 4
    take("x", convs[0][0:3])
    for i, c in enumerate(convs):
        take(f"y_e{i}", c[3:6])
    for i, c in enumerate(fcs):
        take(f"t{i}", (c[1],))
    for j, c in enumerate(ups):
        take(f"y_d{j}", c[3:6])
        take(f"cat{j}", (2 * c[3], c[4], c[5]))
        take(f"att{j}", (c[3],))
        take(f"hid{j}", (2 * c[9],))
        take(f"pool{j}", (3, c[3]))
    assert off == tape.shape[1], (off, tape.shape)
    st.layout = lay
    st._keep = (keep, step_count, tape, hin, dhin, bnpart, wpart, params)
    return st


def stem_train_fwd(st: CaeStemTrain, x: CaeSrc):
    check(lib().cae_unet_stem_train_fwd(C.byref(st), C.byref(x), _stream()), "cae_unet_stem_train_fwd")


def stem_train_bwd(st: CaeStemTrain, x: CaeSrc):
    check(lib().cae_unet_stem_train_bwd(C.byref(st), C.byref(x), _stream()), "cae_unet_stem_train_bwd")


def stem_tape_view(st: CaeStemTrain, name):
    """tensor view [N, ...] of one tape slot (raw layer output) after a fused forward launch"""
    off, shape = st.layout[name]
    n = 1
    for d in shape:
        n *= d
    return st.t_tape[:, off:off + n].reshape(st.t_tape.shape[0], *shape)


# ---- data ingest on the device ----------------------------------------------------------------------------------------------
def minmax(x):
    """(min, max, nan_count) of a fp32 CUDA tensor in one pass"""
    x = x.contiguous()
    part = torch.empty(int(lib().cae_minmax_partials_len()), dtype=torch.float32, device=x.device)
    ticket = torch.zeros(1, dtype=torch.int32, device=x.device)
    out = torch.empty(3, dtype=torch.float32, device=x.device)
    check(lib().cae_minmax(_ptr(x), int(x.numel()), _ptr(part), _ptr(ticket), _ptr(out), _stream()), "cae_minmax")
    lo, hi, nan = out.cpu().tolist()
    return float(lo), float(hi), int(nan)


def normalise_gather(src, order, lo, hi, normalise, dst, chan_offset=0):
    """dst[i, chan_offset:chan_offset + C] = (src[order[i]] - lo) / (hi - lo); src [n, C, y, x], dst [n_out, Ctot, y, x]"""
    assert src.is_contiguous() and dst.is_contiguous() and src.dtype == torch.float32 and dst.dtype == torch.float32
    elems = src[0].numel()
    off = chan_offset * src.shape[2] * src.shape[3]
    check(lib().cae_normalise_gather(_ptr(src), int(elems), _ptr(order), int(dst.shape[0]), float(lo), float(hi),
                                     int(bool(normalise)), dst.data_ptr() + 4 * off, int(dst[0].numel()), _stream()),
          "cae_normalise_gather")


# ---- data-parallel exchange fused into the optimiser ------------------------------------------------------------------------
def case_metrics(yhat, actual, mask, lo, scale, out):
    """rows of eight float64 sums per case (count, sum a, sum e, sum aa, sum ee, sum ae, sum |a-e|, sum (a-e)^2) for the
    post-training metrics; yhat / actual [n, ...] fp32, mask [n, 1 or C, H, W] fp32 or None, out [n, 8] float64"""
    n = yhat.shape[0]
    per_case = yhat[0].numel()
    assert actual.shape == yhat.shape and yhat.is_contiguous() and actual.is_contiguous() and out.shape == (n, 8)
    mpc = mask[0].numel() if mask is not None else per_case
    check(lib().cae_case_metrics(_ptr(yhat), _ptr(actual), _ptr(mask), int(n), int(per_case), int(mpc), float(lo), float(scale),
                                 _ptr(out), _stream()), "cae_case_metrics")


def make_dp_peers(world, rank, grad_ptrs, flag_ptrs) -> CaeDpPeers:
    p = CaeDpPeers()
    p.world, p.rank = int(world), int(rank)
    for r in range(world):
        p.grads[r] = int(grad_ptrs[r])
        p.flags[r] = int(flag_ptrs[r])
    return p


def adam_allreduce(p, peers: CaeDpPeers, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled, grad_scale, step_count, cursor,
                   n_batches, epoch, ticket):
    check(lib().cae_adam_allreduce(_ptr(p), C.byref(peers), _ptr(m), _ptr(v), int(n), float(lr), float(beta1), float(beta2),
                                   float(eps), float(weight_decay), int(bool(decoupled)), float(grad_scale), _ptr(step_count),
                                   _ptr(cursor), int(n_batches), _ptr(epoch), _ptr(ticket), _stream()), "cae_adam_allreduce")


def dp_wait_done(peers: CaeDpPeers, epoch):
    check(lib().cae_dp_wait_done(C.byref(peers), _ptr(epoch), _stream()), "cae_dp_wait_done")
