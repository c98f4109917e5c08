"""Layer-at-a-time execution of the containers' forward pass on the sm_100a kernels (no autograd).

Used by ``Encoder.forward`` / ``Decoder.forward`` and by the per-layer parity tests; it returns every
intermediate so tests can compare against the oracle layer by layer.  Training uses
``engine.convae.ConvAEEngine`` instead, which runs the same kernels from a captured CUDA graph.
"""

from __future__ import annotations

import torch

from . import ops
from .._lib import require_cuda


def _scratch(C, dev):
    return torch.zeros(7, C, dtype=torch.float32, device=dev)


def _bn_block(bn, s):
    return ops.make_bn(bn.num_features, bn.eps, bn.momentum, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                       bn.num_batches_tracked, scale=s[0], shift=s[1], mean=s[2], invstd=s[3], bwdA=s[4], bwdB=s[5],
                       bwdC=s[6])


def _prepare_eval(blocks, dev):
    table = ops.bn_table(blocks, dev)
    ops.bn_eval_prepare(table, len(blocks))
    return table


def conv_stack_forward(x, layers, specs, up, training, final_sigmoid=False, trace=None):
    """Run [(conv, bn|None)] over x. Returns (raw output of last conv, src descriptor pieces of the activation)."""
    require_cuda()
    dev = x.device
    N = x.shape[0]
    keep = []
    src = ops.make_src(x)
    cur_scratch = None
    y = x
    if not training:
        blocks, scr = [], []
        for conv, bn in layers:
            if bn is not None:
                s = _scratch(bn.num_features, dev)
                scr.append(s)
                blocks.append(_bn_block(bn, s))
        if blocks:
            keep.append(_prepare_eval(blocks, dev))
        scr_iter = iter(scr)
    for idx, ((conv, bn), sp) in enumerate(zip(layers, specs)):
        co, ho, wo = sp.get_output_dimensions()
        y = torch.empty(N, co, ho, wo, dtype=torch.float32, device=dev)
        g = ops.geom(sp.get_kernel_size(), sp.get_stride(), 0)
        if bn is not None and training:
            s = _scratch(bn.num_features, dev)
            part = torch.zeros(ops.partials_len(co), dtype=torch.float64, device=dev)
            ticket = torch.zeros(1, dtype=torch.int32, device=dev)
            keep += [s, part, ticket]
            epi = ops.make_epilogue(ops.EPI_STATS, bias=conv.bias, partials=part, ticket=ticket, bn=_bn_block(bn, s))
        elif bn is not None:
            s = next(scr_iter)
            epi = ops.make_epilogue(ops.EPI_PLAIN, bias=conv.bias)
        else:
            s = None
            epi = ops.make_epilogue(ops.EPI_SIGMOID if final_sigmoid else ops.EPI_PLAIN, bias=conv.bias)
        (ops.conv_up if up else ops.conv_down)(src, conv.weight, g, ops.view4(y), epi)
        if trace is not None:
            trace.append((y, s))
        if s is not None:
            src = ops.make_src(y, k0=s[0], k2=s[1], relu=True)
            cur_scratch = s
        else:
            src = ops.make_src(y)
            cur_scratch = None
        keep.append(y)
    torch.cuda.current_stream().synchronize()  # temporaries die with this frame
    return y, cur_scratch


def linear(x, lin, relu_out=False, a_k0=None, a_k2=None, a_hw=1, a_relu=False):
    N, K = x.shape[0], lin.in_features
    out = torch.empty(N, lin.out_features, dtype=torch.float32, device=x.device)
    ops.gemm(N, lin.out_features, K, x, K, 1, lin.weight, 1, K, out, lin.out_features, 1, a_k0=a_k0, a_k2=a_k2,
             a_hw=a_hw, a_relu=a_relu, bias=lin.bias, relu_out=relu_out)
    return out


def encoder_forward(enc, x, trace=None):
    x = x.contiguous().float()
    y, s = conv_stack_forward(x, enc.conv_layers(), enc.layer_specs, up=False, training=enc.training, trace=trace)
    c, h, w = enc.layer_specs[-1].get_output_dimensions()
    h1 = linear(y.view(y.shape[0], -1), enc.encoder_lin[0], relu_out=True, a_k0=s[0], a_k2=s[1], a_hw=h * w,
                a_relu=True)
    z = linear(h1, enc.encoder_lin[2])
    torch.cuda.current_stream().synchronize()
    return z


def decoder_forward(dec, z, trace=None):
    z = z.contiguous().float()
    h3 = linear(z, dec.decoder_lin[0], relu_out=True)
    u = linear(h3, dec.decoder_lin[2])
    c, h, w = dec.layer_specs[0].get_input_dimensions()
    y, _ = conv_stack_forward(u.view(-1, c, h, w), dec.conv_layers(), dec.layer_specs, up=True,
                              training=dec.training, final_sigmoid=True, trace=trace)
    return y
