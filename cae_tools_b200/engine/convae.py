"""Step program for the ConvAEModel hot path (reference: conv_ae_model.py:185-239 train/test/score loops,
encoder.py:60-64, decoder.py:73-78, torch.nn.MSELoss :303, torch.optim.Adam :310).

The engine owns
  * one flat fp32 arena for all trainable parameters (+ grads, Adam m and v of the same layout); the
    nn.Parameters of the Encoder/Decoder containers are re-pointed at views of it so state_dict(),
    save() and load() keep working unchanged;
  * per-layer activation / gradient buffers, BatchNorm scratch, reduction workspaces;
  * an explicit forward + backward kernel schedule (no autograd) that is captured once per batch
    geometry into a CUDA graph and replayed; the batch to use is selected on the device through a
    cursor, so an epoch is `n_batches` graph launches with no host<->device traffic in between.

Kernel schedule for one optimiser step (L_e encoder convs, L_d transposed convs):
  forward   conv_down x L_e (bias + BN statistics in the epilogue; BN+ReLU applied by the *consumer* on load)
            gemm x 4       (fc stack; first one applies BN+ReLU+Flatten on load)
            conv_up x L_d  (same; the last one fuses sigmoid + MSE loss + dL/dz)
  backward  per transposed conv: wgrad, then conv_down as dgrad with ReLU-mask + BN-backward sums in the epilogue
            gemm x 7 (fc), ew_epilogue (mask + BN sums), per conv: wgrad (+ conv_up as dgrad)
  update    one fused Adam launch over the arena, one bookkeeping launch (step counter, batch cursor)
"""

from __future__ import annotations

import os

import torch

from . import ops
from .._lib import require_cuda


class DataBinding:
    """A pre-batched, device-resident data set (reference keeps every batch on the device,
    conv_ae_model.py:315-325): X [n,C,H,W], optional Y, a device cursor and a per-batch loss array."""

    def __init__(self, X, Y, batch_size):
        self.X = X.contiguous()
        self.Y = Y.contiguous() if Y is not None else None
        self.n = int(X.shape[0])
        self.batch_size = int(batch_size)
        self.n_batches = (self.n + self.batch_size - 1) // self.batch_size
        self.tail = self.n - (self.n_batches - 1) * self.batch_size
        dev = X.device
        self.cursor = torch.zeros(1, dtype=torch.int32, device=dev)
        self.losses = torch.zeros(self.n_batches, dtype=torch.float32, device=dev)
        self.graphs = {}

    def batch_sizes(self):
        """[(N, count)] in replay order: full batches then the ragged tail"""
        if self.tail == self.batch_size:
            return [(self.batch_size, self.n_batches)]
        out = []
        if self.n_batches > 1:
            out.append((self.batch_size, self.n_batches - 1))
        out.append((self.tail, 1))
        return out


def _align(n, a=4):
    return (n + a - 1) // a * a


class ConvAEEngine:

    # Data-parallel exchange in two buckets, the decoder + fc bucket overlapped with the encoder backward
    # (_BucketedProgram).  Measured on 2 x B200 (unet, batch 64 per GPU): 0.359 ms/step against 0.346 ms for ONE
    # all-reduce of the whole 141 KB arena after the backward pass - the third graph launch and the second NCCL call
    # cost more than a 141 KB all-reduce over NVLink (~15 us) can hide.  Off by default; worth it for config-4-sized
    # arenas (26 MB).
    overlap_allreduce = False

    # Data-parallel exchange captured inside the step graph (CAE_CAPTURE_ALLREDUCE=1).  OFF by default - measured on 2 x B200
    # (unet, batch 64 per GPU): 0.239 ms/step captured against 0.222 ms single-GPU, no better than the eager NCCL call
    # between two graphs (the 21 us latency of a 157 KB all-reduce is on the critical path either way), and a process
    # group destroyed while graphs that captured NCCL kernels are alive hung at exit.
    capture_allreduce = os.environ.get("CAE_CAPTURE_ALLREDUCE", "0") == "1"

    # fc bottleneck as one launch per direction (fc_stack.cu).  Correct and tested, but measured no faster than the
    # cae_gemm chain on B200 (unet batch 64: 25 + 40 us fused against 31 + 35 us; conv: 30 + 41 against 14 + 20): a
    # dependent launch costs only ~1 us inside a graph, while a single CTA pays every phase's latency serially.
    use_fused_fc = False

    # data parallel, small arenas: the gradient all-reduce runs INSIDE the optimiser launch (peer reads over NVLink,
    # csrc/dp_fused.cu) instead of an NCCL call between the backward pass and Adam; larger arenas keep NCCL
    DP_FUSED_MAX_BYTES = 1 << 20

    def __init__(self, encoder, decoder, lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8, decoupled=False,
                 device="cuda", use_graphs=True, grad_hook=None, grad_scale=1.0, count_scale=1.0, grad_hook_async=None,
                 dp=None):
        require_cuda()
        self.device = torch.device(device)
        self.encoder = encoder.to(self.device)
        self.decoder = decoder.to(self.device)
        self.lr, self.weight_decay, self.betas, self.eps, self.decoupled = lr, weight_decay, betas, eps, decoupled
        self.use_graphs = use_graphs
        self.grad_hook = grad_hook      # callable(flat_grads) between backward and Adam (data-parallel all-reduce)
        # callable(tensor) -> handle with .wait(): asynchronous all-reduce of one gradient bucket.  With it the exchange
        # is bucketed and overlapped: the decoder + fc gradients are reduced while the encoder backward still runs.
        self.grad_hook_async = grad_hook_async
        self.grad_scale = grad_scale
        self.count_scale = count_scale  # n_local / n_global when a batch is sharded over data-parallel ranks
        self.mse_weight = 1.0           # weight of the MSE term in the reported loss / gradient (VarAE: lambda_mse)
        self._keep = []                 # descriptors' tensors must outlive the graphs
        self.dp = dp                    # engine/dp.py:DPContext (enables the fused exchange when the arena is small)
        self._dp_peers = None
        self._build_arena()
        self.enc_layers = self.encoder.conv_layers()
        self.dec_layers = self.decoder.conv_layers()
        self.enc_specs = self.encoder.layer_specs
        self.dec_specs = self.decoder.layer_specs
        self.step_count = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._bufs = {}
        self._bn_scratch = {}
        self._progs = {}
        self._tickets = torch.zeros(4096, dtype=torch.int32, device=self.device)
        self._next_ticket = 0
        self._bn_table = None
        self._tc, self._tc_shared = {}, {}

    # ------------------------------------------------------------------ parameters
    def _build_arena(self):
        params = list(self.encoder.parameters()) + list(self.decoder.parameters())
        n_enc = len(list(self.encoder.parameters()))
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += _align(p.numel())
        dev = self.device
        self.arena = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grads = None
        if self.dp is not None and total * 4 <= self.DP_FUSED_MAX_BYTES:
            sym = self.dp.symmetric_arena(total)
            if sym is not None:
                self.grads, self._dp_peers, keep = sym
                self._keep.append(keep)
                self._dp_epoch = torch.zeros(1, dtype=torch.int32, device=dev)
                self.grad_hook = self.grad_hook_async = None        # the exchange happens inside the optimiser launch
        if self.grads is None:
            self.grads = torch.zeros(total, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self._gview = {}
        with torch.no_grad():
            for p, off in zip(params, offs):
                n = p.numel()
                self.arena[off:off + n].copy_(p.detach().reshape(-1).to(dev, torch.float32))
                p.data = self.arena[off:off + n].view(p.shape)
                g = self.grads[off:off + n].view(p.shape)
                p.grad = g
                self._gview[id(p)] = g
        self.n_params = total
        self.enc_param_end = offs[n_enc] if n_enc < len(offs) else total     # [0, end): encoder bucket, [end, total): decoder

    def g(self, p):
        """gradient view (inside the flat grad arena) of parameter p"""
        return self._gview[id(p)]

    # ------------------------------------------------------------------ buffers
    def _f32(self, *shape):
        t = torch.zeros(*shape, dtype=torch.float32, device=self.device)
        self._keep.append(t)
        return t

    def _conv_buf(self, B, C, H, W, dense=False):
        """activation buffer [B,C,H,W]; rows padded to a multiple of 4 floats (16-byte aligned rows let the wide-layer
        kernels use float4 loads/stores) unless the fc stack reads it as a flat matrix (dense)"""
        if dense or W % 4 == 0:
            return self._f32(B, C, H, W)
        ld = (W + 3) // 4 * 4
        return self._f32(B, C, H, ld)[:, :, :, :W]

    def _ticket(self):
        i = self._next_ticket
        self._next_ticket += 1
        assert i < self._tickets.numel()
        return self._tickets[i:i + 1]

    def _partials(self, C):
        t = torch.zeros(ops.partials_len(C), dtype=torch.float64, device=self.device)
        self._keep.append(t)
        return t

    def _bn(self, key, bn_mod, conv_bias_grad=None):
        """CaeBN block for a BatchNorm2d module (scratch allocated once per module)"""
        if key not in self._bn_scratch:
            Cn = bn_mod.num_features
            self._bn_scratch[key] = self._f32(7, Cn)
        s = self._bn_scratch[key]
        return ops.make_bn(bn_mod.num_features, bn_mod.eps, bn_mod.momentum, bn_mod.weight, bn_mod.bias,
                           bn_mod.running_mean, bn_mod.running_var, bn_mod.num_batches_tracked,
                           scale=s[0], shift=s[1], mean=s[2], invstd=s[3],
                           dgamma=self.g(bn_mod.weight), dbeta=self.g(bn_mod.bias), dbias=conv_bias_grad,
                           bwdA=s[4], bwdB=s[5], bwdC=s[6]), s

    def _act_buffers(self, B):
        """activation / gradient buffers for batch capacity B"""
        if B in self._bufs:
            return self._bufs[B]
        b = {}
        ne = len(self.enc_specs)
        b["y_e"] = [self._conv_buf(B, *sp.get_output_dimensions(), dense=(i == ne - 1))
                    for i, sp in enumerate(self.enc_specs)]
        b["dz_e"] = [self._conv_buf(B, *sp.get_output_dimensions(), dense=(i == ne - 1))
                     for i, sp in enumerate(self.enc_specs)]
        c0, h0, w0 = self.dec_specs[0].get_input_dimensions()
        b["u"] = self._f32(B, c0, h0, w0)
        b["du"] = self._f32(B, c0, h0, w0)
        ce, he, we = self.enc_specs[-1].get_output_dimensions()
        b["da"] = self._f32(B, ce, he, we)
        self._fc_buffers(b, B)
        b["y_d"] = [self._conv_buf(B, *sp.get_output_dimensions()) for sp in self.dec_specs]   # last: dL/dz or yhat
        b["dz_d"] = [self._conv_buf(B, *sp.get_output_dimensions()) for sp in self.dec_specs[:-1]]
        self._bufs[B] = b
        return b

    def _fc_buffers(self, b, B):
        lin, dlin = self.encoder.encoder_lin, self.decoder.decoder_lin
        b["h1"] = self._f32(B, lin[0].out_features)
        b["z"] = self._f32(B, lin[2].out_features)
        b["h3"] = self._f32(B, dlin[0].out_features)
        b["dh3"] = self._f32(B, dlin[0].out_features)
        b["dzl"] = self._f32(B, lin[2].out_features)
        b["dh1"] = self._f32(B, lin[0].out_features)

    # ------------------------------------------------------------------ tensor-core layers
    # A transposed conv goes to the tcgen05 path (tc_conv.cu: pack -> 3xTF32 GEMM -> col2im) when it is a genuinely dense
    # contraction: >= TC_MIN_CIN input channels and >= TC_MIN_FLOPS per launch.  BASELINE configs[3] (4x64x64 ->
    # 4x1024x1024, batch 128): the layers 1024->512 ... 64->32; nothing in the 16x16 -> 256x256 configs qualifies.
    TC_MIN_CIN = int(os.environ.get("CAE_TC_MIN_CIN", "64"))
    TC_MIN_FLOPS = float(os.environ.get("CAE_TC_MIN_FLOPS", "2e9"))

    def _tc_desc(self, j, N):
        """CaeTcConv descriptor of decoder layer j at batch N, or None when the layer stays on the SIMT kernels"""
        key = (j, N)
        if key in self._tc:
            return self._tc[key]
        sp = self.dec_specs[j]
        cin, hin, win = sp.get_input_dimensions()
        cout, hout, wout = sp.get_output_dimensions()
        k = sp.get_kernel_size()
        kh, kw = (k if isinstance(k, (tuple, list)) else (k, k))
        flops = 2.0 * N * hin * win * cin * cout * kh * kw
        desc = None
        if j < len(self.dec_specs) - 1 and cin >= self.TC_MIN_CIN and flops >= self.TC_MIN_FLOPS and \
                ops.tc_convT_supported(cin, cout, k, sp.get_stride(), 0):
            desc = ops.make_tc_conv(cin, cout, k, sp.get_stride(), N, hin, win, hout, wout, self.device, self._tc_shared)
        self._tc[key] = desc
        return desc

    # ------------------------------------------------------------------ schedules
    def _forward_ops(self, b, N, data, train, final):
        """final: 'loss_grad' (train), 'loss' (test epoch), 'yhat' (score: writes sigmoid output into y_d[-1])"""
        sched = []
        X = data.X
        src = ops.make_src(X[:data.batch_size] if X.shape[0] >= data.batch_size else X, cursor=data.cursor,
                           cursor_stride=data.batch_size * X[0].numel(), n=N)
        # the view above must describe one batch worth of samples starting at X[0]; N limits the count
        for i, ((conv, bn), sp) in enumerate(zip(self.enc_layers, self.enc_specs)):
            y = b["y_e"][i]
            blk, s = self._bn(("e", i), bn, self.g(conv.bias))
            if train:
                epi = ops.make_epilogue(ops.EPI_STATS, bias=conv.bias, partials=self._partials(conv.out_channels),
                                        ticket=self._ticket(), bn=blk)
            else:
                epi = ops.make_epilogue(ops.EPI_PLAIN, bias=conv.bias)
            g = ops.geom(sp.get_kernel_size(), sp.get_stride(), 0)
            sched.append((f"fwd.conv{i}", lambda src=src, w=conv.weight, g=g, o=ops.view4(y, N), e=epi: ops.conv_down(src, w, g, o, e)))
            src = ops.make_src(y, k0=s[0], k2=s[1], relu=True, n=N)
        sched += self._fc_forward_ops(b, N, data, train)
        src = ops.make_src(b["u"], n=N)
        nd = len(self.dec_layers)
        for j, ((conv, bn), sp) in enumerate(zip(self.dec_layers, self.dec_specs)):
            y = b["y_d"][j]
            g = ops.geom(sp.get_kernel_size(), sp.get_stride(), 0)
            if j < nd - 1:
                blk, s = self._bn(("d", j), bn, self.g(conv.bias))
                if train:
                    epi = ops.make_epilogue(ops.EPI_STATS, bias=conv.bias,
                                            partials=self._partials(conv.out_channels), ticket=self._ticket(), bn=blk)
                else:
                    epi = ops.make_epilogue(ops.EPI_PLAIN, bias=conv.bias)
                tc = self._tc_desc(j, N)
                if tc is not None:
                    sched.append((f"fwd.convT{j}.tc", lambda tc=tc, src=src, w=conv.weight, o=ops.view4(y, N), e=epi:
                                  ops.tc_convT_fwd(tc, src, w, o, e)))
                else:
                    sched.append((f"fwd.convT{j}", lambda src=src, w=conv.weight, g=g, o=ops.view4(y, N), e=epi:
                                  ops.conv_up(src, w, g, o, e)))
                src = ops.make_src(y, k0=s[0], k2=s[1], relu=True, n=N)
            else:
                if final == "yhat":
                    epi = ops.make_epilogue(ops.EPI_SIGMOID, bias=conv.bias)
                else:
                    Y = data.Y
                    tgt = ops.make_src(Y[:data.batch_size] if Y.shape[0] >= data.batch_size else Y, cursor=data.cursor,
                                       cursor_stride=data.batch_size * Y[0].numel(), n=N)
                    epi = ops.make_epilogue(ops.EPI_SIGMOID_MSE, bias=conv.bias,
                                            partials=self._partials(conv.out_channels), ticket=self._ticket(),
                                            target=tgt, loss_out=data.losses,
                                            dbias=self.g(conv.bias) if train else None,
                                            write_mode=0 if final == "loss_grad" else 2,
                                            count_scale=self.count_scale * self.mse_weight)
                sched.append((f"fwd.convT{j}+sigmoid" + ("" if final == "yhat" else "+mse"),
                              lambda src=src, w=conv.weight, g=g, o=ops.view4(y, N), e=epi: ops.conv_up(src, w, g, o, e)))
        return sched

    def _fc_forward_ops(self, b, N, data, train):
        """encoder_lin + decoder_lin: y_e[-1] (BN+ReLU+Flatten on load) -> h1 -> z -> h3 -> u"""
        sched = []
        lin, dlin = self.encoder.encoder_lin, self.decoder.decoder_lin
        ylast = b["y_e"][-1]
        ce, he, we = self.enc_specs[-1].get_output_dimensions()
        flat = ce * he * we
        s_last = self._bn_scratch[("e", len(self.enc_layers) - 1)]
        fc, lat = lin[0].out_features, lin[2].out_features
        fc2 = dlin[0].out_features
        out4 = dlin[2].out_features
        if self._fc_fused(N, flat, fc, lat, fc2, out4):
            # h1 / h3 hold the PRE-activations here (the fused backward differentiates the ReLU from them)
            p = ops.make_fc_stack(N, ylast, (lin[0], lin[2], dlin[0], dlin[2]), b["h1"], b["z"], b["h3"], b["u"],
                                  a_k0=s_last[0], a_k2=s_last[1], a_hw=he * we, a_relu=True)
            return [("fwd.fcstack", lambda p=p: ops.fc_stack_fwd(p))]
        sched.append(("fwd.fc1", lambda: ops.gemm(N, fc, flat, ylast, flat, 1, lin[0].weight, 1, flat, b["h1"], fc, 1,
                                                  a_k0=s_last[0], a_k2=s_last[1], a_hw=he * we, a_relu=True,
                                                  bias=lin[0].bias, relu_out=True)))
        sched.append(("fwd.fc2", lambda: ops.gemm(N, lat, fc, b["h1"], fc, 1, lin[2].weight, 1, fc, b["z"], lat, 1,
                                                  bias=lin[2].bias)))
        sched.append(("fwd.fc3", lambda: ops.gemm(N, fc2, lat, b["z"], lat, 1, dlin[0].weight, 1, lat, b["h3"], fc2, 1,
                                                  bias=dlin[0].bias, relu_out=True)))
        sched.append(("fwd.fc4", lambda: ops.gemm(N, out4, fc2, b["h3"], fc2, 1, dlin[2].weight, 1, fc2, b["u"], out4,
                                                  1, bias=dlin[2].bias)))
        return sched

    def _fc_fused(self, N, *dims):
        return self.use_fused_fc and ops.fc_stack_supported(N, *dims)

    def _fc_backward_ops(self, b, N, data):
        """du -> gradients of the four Linear layers -> da (wrt the last encoder activation)"""
        sched = []
        lin, dlin = self.encoder.encoder_lin, self.decoder.decoder_lin
        ce, he, we = self.enc_specs[-1].get_output_dimensions()
        flat = ce * he * we
        fc, lat = lin[0].out_features, lin[2].out_features
        fc2, out4 = dlin[0].out_features, dlin[2].out_features
        G = self.g
        le = len(self.enc_layers) - 1
        s_last = self._bn_scratch[("e", le)]
        ylast = b["y_e"][-1]
        if self._fc_fused(N, flat, fc, lat, fc2, out4):
            p = ops.make_fc_stack(N, ylast, (lin[0], lin[2], dlin[0], dlin[2]), b["h1"], b["z"], b["h3"], b["u"],
                                  a_k0=s_last[0], a_k2=s_last[1], a_hw=he * we, a_relu=True, du=b["du"], grads=G,
                                  dA=b["da"])
            return [("bwd.fcstack", lambda p=p: ops.fc_stack_bwd(p))]
        # Linear 4: u = h3 W4^T + b4
        sched.append(("bwd.fc4.dW", lambda: ops.gemm(out4, fc2, N, b["du"], 1, out4, b["h3"], fc2, 1, G(dlin[2].weight), fc2, 1,
                                      rowsum_A=G(dlin[2].bias))))
        sched.append(("bwd.fc4.dx", lambda: ops.gemm(N, fc2, out4, b["du"], out4, 1, dlin[2].weight, fc2, 1, b["dh3"], fc2, 1,
                                      mask=b["h3"])))
        # Linear 3: h3 = relu(z W3^T + b3)
        sched.append(("bwd.fc3.dW", lambda: ops.gemm(fc2, lat, N, b["dh3"], 1, fc2, b["z"], lat, 1, G(dlin[0].weight), lat, 1,
                                      rowsum_A=G(dlin[0].bias))))
        sched.append(("bwd.fc3.dx", lambda: ops.gemm(N, lat, fc2, b["dh3"], fc2, 1, dlin[0].weight, lat, 1, b["dzl"], lat, 1)))
        # Linear 2: z = h1 W2^T + b2
        sched.append(("bwd.fc2.dW", lambda: ops.gemm(lat, fc, N, b["dzl"], 1, lat, b["h1"], fc, 1, G(lin[2].weight), fc, 1,
                                      rowsum_A=G(lin[2].bias))))
        sched.append(("bwd.fc2.dx", lambda: ops.gemm(N, fc, lat, b["dzl"], lat, 1, lin[2].weight, fc, 1, b["dh1"], fc, 1,
                                      mask=b["h1"])))
        # Linear 1: h1 = relu(a W1^T + b1), a = relu(bn(y_last)) flattened
        sched.append(("bwd.fc1.dW", lambda: ops.gemm(fc, flat, N, b["dh1"], 1, fc, ylast, flat, 1, G(lin[0].weight), flat, 1,
                                      b_k0=s_last[0], b_k2=s_last[1], b_hw=he * we, b_relu=True,
                                      rowsum_A=G(lin[0].bias))))
        sched.append(("bwd.fc1.dx", lambda: ops.gemm(N, flat, fc, b["dh1"], fc, 1, lin[0].weight, flat, 1, b["da"], flat, 1)))
        return sched

    def _wgrad_op(self, small, big, g, grad):
        part = torch.zeros(ops.wgrad_partials_len(small, big, g), dtype=torch.float32, device=self.device)
        self._keep.append(part)
        t = self._ticket()
        return lambda: ops.conv_wgrad(small, big, g, grad, part, t)

    def _backward_ops(self, b, N, data):
        sched = []
        nd = len(self.dec_layers)
        # ---- decoder
        for j in range(nd - 1, -1, -1):
            conv, bn = self.dec_layers[j]
            sp = self.dec_specs[j]
            g = ops.geom(sp.get_kernel_size(), sp.get_stride(), 0)
            if j == nd - 1:
                dy = ops.make_src(b["y_d"][j], n=N)            # holds dL/dz of the fused sigmoid+MSE epilogue
            else:
                s = self._bn_scratch[("d", j)]
                dy = ops.make_src(b["dz_d"][j], t1=b["y_d"][j], k0=s[4], k1=s[5], k2=s[6], n=N)
            if j > 0:
                sp_prev = self._bn_scratch[("d", j - 1)]
                x_in = ops.make_src(b["y_d"][j - 1], k0=sp_prev[0], k2=sp_prev[1], relu=True, n=N)
            else:
                x_in = ops.make_src(b["u"], n=N)
            tc = self._tc_desc(j, N)
            if tc is None:
                sched.append((f"bwd.convT{j}.wgrad", self._wgrad_op(x_in, dy, g, self.g(conv.weight))))
            if j > 0:
                pconv, pbn = self.dec_layers[j - 1]
                blk, _ = self._bn(("d", j - 1), pbn, self.g(pconv.bias))
                epi = ops.make_epilogue(ops.EPI_MASKSTATS, partials=self._partials(pconv.out_channels),
                                        ticket=self._ticket(), bn=blk, act=b["y_d"][j - 1], n=N)
                out = ops.view4(b["dz_d"][j - 1], N)
            else:
                epi = ops.make_epilogue(ops.EPI_PLAIN)
                out = ops.view4(b["du"], N)
            if tc is not None:
                # tensor-core layer: one im2col of dL/dy feeds both GEMMs; they fill the machine, so everything stays on
                # the main stream (the scratch operands are shared between layers)
                sched.append((f"bwd.convT{j}.tc.im2col", lambda tc=tc, dy=dy: ops.tc_convT_im2col(tc, dy)))
                sched.append((f"bwd.convT{j}.tc.dgrad", lambda tc=tc, w=conv.weight, o=out, e=epi:
                              ops.tc_convT_dgrad(tc, w, o, e)))
                sched.append((f"bwd.convT{j}.tc.wgrad_gemm", lambda tc=tc, gw=self.g(conv.weight): ops.tc_convT_wgrad(tc, gw)))
                continue
            sched.append((f"bwd.convT{j}.dgrad", lambda dy=dy, w=conv.weight, g=g, o=out, e=epi:
                          ops.conv_down(dy, w, g, o, e)))
        sched += self._fc_backward_ops(b, N, data)
        le = len(self.enc_layers) - 1
        ylast = b["y_e"][-1]
        # ReLU mask + BN-backward sums of the last encoder layer
        conv, bn = self.enc_layers[le]
        blk, _ = self._bn(("e", le), bn, self.g(conv.bias))
        epi = ops.make_epilogue(ops.EPI_MASKSTATS, partials=self._partials(conv.out_channels), ticket=self._ticket(),
                                bn=blk, act=ylast, n=N)
        sched.append(("bwd.enc_last.mask+bnsums", lambda s=ops.make_src(b["da"], n=N), o=ops.view4(b["dz_e"][le], N), e=epi:
                      ops.ew_epilogue(s, o, e)))
        # ---- encoder
        X = data.X
        for i in range(le, -1, -1):
            conv, bn = self.enc_layers[i]
            sp = self.enc_specs[i]
            g = ops.geom(sp.get_kernel_size(), sp.get_stride(), 0)
            s = self._bn_scratch[("e", i)]
            dy = ops.make_src(b["dz_e"][i], t1=b["y_e"][i], k0=s[4], k1=s[5], k2=s[6], n=N)
            if i > 0:
                sp_prev = self._bn_scratch[("e", i - 1)]
                x_in = ops.make_src(b["y_e"][i - 1], k0=sp_prev[0], k2=sp_prev[1], relu=True, n=N)
            else:
                x_in = ops.make_src(X[:data.batch_size] if X.shape[0] >= data.batch_size else X, cursor=data.cursor,
                                    cursor_stride=data.batch_size * X[0].numel(), n=N)
            sched.append((f"bwd.conv{i}.wgrad", self._wgrad_op(dy, x_in, g, self.g(conv.weight))))
            if i > 0:
                pconv, pbn = self.enc_layers[i - 1]
                blk, _ = self._bn(("e", i - 1), pbn, self.g(pconv.bias))
                epi = ops.make_epilogue(ops.EPI_MASKSTATS, partials=self._partials(pconv.out_channels),
                                        ticket=self._ticket(), bn=blk, act=b["y_e"][i - 1], n=N)
                sched.append((f"bwd.conv{i}.dgrad", lambda dy=dy, w=conv.weight, g=g, o=ops.view4(b["dz_e"][i - 1], N),
                              e=epi: ops.conv_up(dy, w, g, o, e)))
        return sched

    def _update_ops(self, data):
        out = []
        if self.grad_hook is not None:
            out.append(("grad_allreduce", lambda: self.grad_hook(self.grads)))
        if self._dp_peers is not None:
            # gradient all-reduce (peer reads over NVLink) + optimiser + bookkeeping in one launch
            out.append(("adam+allreduce", lambda t=self._ticket(): ops.adam_allreduce(
                self.arena, self._dp_peers, self.adam_m, self.adam_v, self.n_params, self.lr, self.betas[0], self.betas[1],
                self.eps, self.weight_decay, self.decoupled, self.grad_scale, self.step_count, data.cursor, data.n_batches,
                self._dp_epoch, t)))
            return out
        # optimiser + bookkeeping (step counter, batch cursor) in one launch
        out.append(("adam", lambda t=self._ticket(): ops.adam_advance(
            self.arena, self.grads, self.adam_m, self.adam_v, self.n_params, self.lr, self.betas[0], self.betas[1], self.eps,
            self.weight_decay, self.decoupled, self.grad_scale, self.step_count, data.cursor, data.n_batches, t)))
        return out

    def _eval_prepare_op(self):
        if self._bn_table is None:
            blocks = []
            for i, (conv, bn) in enumerate(self.enc_layers):
                blocks.append(self._bn(("e", i), bn)[0])
            for j, (conv, bn) in enumerate(self.dec_layers):
                if bn is not None:
                    blocks.append(self._bn(("d", j), bn)[0])
            self._bn_count = len(blocks)
            self._bn_table = ops.bn_table(blocks, self.device)
        return lambda: ops.bn_eval_prepare(self._bn_table, self._bn_count)

    # ------------------------------------------------------------------ execution
    def program(self, kind, data, N):
        """the step program (op list captured into a CUDA graph on first run) for one batch geometry: kind = "train" |
        "test" | "score"; .run() executes one step, .n_launches / .sched describe it, .profile() / .timeline() time it"""
        return self._program(kind, data, N)

    def _program(self, kind, data, N):
        """build (and cache on the binding) the op list / CUDA graph for one batch geometry"""
        key = (kind, N)
        if key in data.graphs:
            return data.graphs[key]
        b = self._act_buffers(data.batch_size)
        if kind == "train":
            sched = self._forward_ops(b, N, data, True, "loss_grad") + self._backward_ops(b, N, data) + \
                self._update_ops(data)
            if self._dp_peers is not None:
                # fused exchange: before anything of this step can overwrite a gradient, every peer must have finished
                # reading the previous step's (flags published by its cae_adam_allreduce ~one forward pass ago: a free wait)
                sched = [("dp.wait_peers", lambda: ops.dp_wait_done(self._dp_peers, self._dp_epoch))] + sched
        elif kind == "test":
            sched = self._forward_ops(b, N, data, False, "loss") + \
                [("advance", lambda: ops.step_advance(None, data.cursor, data.n_batches))]
        elif kind == "score":
            sched = self._forward_ops(b, N, data, False, "yhat") + \
                [("advance", lambda: ops.step_advance(None, data.cursor, data.n_batches))]
        else:
            raise ValueError(kind)
        state = [data.cursor, data.losses] + list(getattr(data, "extra_state", []))
        if kind == "train":
            state += [self.arena, self.adam_m, self.adam_v, self.grads, self.step_count]
            for mod in list(self.encoder.modules()) + list(self.decoder.modules()):
                if isinstance(mod, torch.nn.modules.batchnorm._BatchNorm):
                    state += [mod.running_mean, mod.running_var, mod.num_batches_tracked]
        names = [n for n, _ in sched]
        SPLIT = "bwd.enc_last.mask+bnsums"
        if "grad_allreduce" in names and self.grad_hook_async is not None and self.overlap_allreduce and SPLIT in names:
            # bucketed + overlapped exchange: [forward, decoder + fc backward] -> all-reduce(decoder bucket) in flight
            # while [encoder backward] runs -> all-reduce(encoder bucket) -> wait both -> [Adam]
            i, j = names.index(SPLIT), names.index("grad_allreduce")
            e = self.enc_param_end
            prog = _BucketedProgram(_Program(sched[:i], self.use_graphs, state), _Program(sched[i:j], self.use_graphs, state),
                                    _Program(sched[j + 1:], self.use_graphs, state), self.grad_hook_async,
                                    self.grads[e:], self.grads[:e])
        elif "grad_allreduce" in names and self.capture_allreduce and self.use_graphs:
            # opt-in: the gradient all-reduce captured INSIDE the step graph (see capture_allreduce above)
            prog = _Program(sched, self.use_graphs, state)
        elif "grad_allreduce" in names:
            i = names.index("grad_allreduce")
            prog = _SplitProgram(_Program(sched[:i], self.use_graphs, state), sched[i][1],
                                 _Program(sched[i + 1:], self.use_graphs, state))
        else:
            prog = _Program(sched, self.use_graphs, state)
        data.graphs[key] = prog
        return prog

    def bind(self, X, Y, batch_size):
        X = X.to(self.device, torch.float32)
        Y = Y.to(self.device, torch.float32) if Y is not None else None
        return DataBinding(X, Y, batch_size)

    def batch_losses(self, data):
        """per-batch loss values of the last epoch run on this binding (device tensor)"""
        return data.losses

    def train_epoch(self, data):
        """one pass over all batches; returns the per-batch losses (device tensor, no sync)"""
        data.cursor.zero_()
        for N, count in data.batch_sizes():
            prog = self._program("train", data, N)
            for _ in range(count):
                prog.run()
        return self.batch_losses(data)

    def train_steps(self, data, steps):
        """run `steps` optimiser steps cycling through the batches (bench helper; full batches only)"""
        assert data.tail == data.batch_size, "train_steps needs n % batch_size == 0"
        prog = self._program("train", data, data.batch_size)
        for _ in range(steps):
            prog.run()

    def train_stream(self, host_batches, batch_size):
        """Optimiser steps fed from HOST memory: `host_batches` yields (x, y) or (x, y, mask) CPU tensors of one batch
        each (pinned memory makes the copies asynchronous).  Two device slots: while step i computes on slot i % 2 the
        copy stream fills slot (i + 1) % 2, so the PCIe transfer of the next batch hides behind the current step; the
        4-byte loss of every step is copied back asynchronously into pinned memory.  Returns the per-step losses
        (host tensor) after one final synchronisation.  Every batch must hold exactly `batch_size` samples."""
        st = getattr(self, "_stream_state", None)
        if st is None or st["B"] != batch_size:
            st = {"B": batch_size, "data": None, "copy": torch.cuda.Stream(device=self.device),
                  "filled": [torch.cuda.Event(), torch.cuda.Event()], "freed": [torch.cuda.Event(), torch.cuda.Event()]}
            self._stream_state = st
        main = torch.cuda.current_stream(self.device)
        B = batch_size
        losses_host, n = None, 0
        for i, hb in enumerate(host_batches):
            xh, yh = hb[0], hb[1]
            mh = hb[2] if len(hb) > 2 else None
            if st["data"] is None:
                X2 = torch.empty(2 * B, *xh.shape[1:], dtype=torch.float32, device=self.device)
                Y2 = torch.empty(2 * B, *yh.shape[1:], dtype=torch.float32, device=self.device)
                if mh is not None:
                    st["data"] = self.bind(X2, Y2, B, mask=torch.zeros(2 * B, *mh.shape[1:], dtype=torch.float32, device=self.device))
                    st["data"].mse_scale = None      # streamed masks are not known at bind time: constant count_scale
                else:
                    st["data"] = self.bind(X2, Y2, B)
                st["prog"] = self._program("train", st["data"], B)
                st["host_losses"] = torch.empty(1 << 16, dtype=torch.float32).pin_memory()
            data = st["data"]
            if i == 0:
                data.cursor.zero_()
                st["copy"].wait_stream(main)
            slot = i % 2
            lo, hi = slot * B, (slot + 1) * B
            with torch.cuda.stream(st["copy"]):
                if i >= 2:
                    st["copy"].wait_event(st["freed"][slot])        # step i-2 no longer reads this slot
                data.X[lo:hi].copy_(xh, non_blocking=True)
                data.Y[lo:hi].copy_(yh, non_blocking=True)
                if mh is not None:
                    data.M[lo:hi].copy_(mh, non_blocking=True)
                st["filled"][slot].record(st["copy"])
            main.wait_event(st["filled"][slot])
            st["prog"].run()
            st["freed"][slot].record(main)
            st["host_losses"][i % (1 << 16):i % (1 << 16) + 1].copy_(self.batch_losses(data)[slot:slot + 1], non_blocking=True)
            n = i + 1
        main.synchronize()
        return st["host_losses"][:min(n, 1 << 16)].clone() if n else torch.empty(0)

    def test_epoch(self, data):
        self._eval_prepare_op()()
        data.cursor.zero_()
        for N, count in data.batch_sizes():
            prog = self._program("test", data, N)
            for _ in range(count):
                prog.run()
        return self.batch_losses(data)

    def output_buffer(self, b):
        """buffer that holds the prediction after a "score" program ran"""
        return b["y_d"][-1]

    def score_batches(self, data, sink):
        """eval-mode forward of every batch; sink(batch_index, yhat[N,C,H,W] device view) consumes each output"""
        self._eval_prepare_op()()
        data.cursor.zero_()
        b = self._act_buffers(data.batch_size)
        idx = 0
        for N, count in data.batch_sizes():
            prog = self._program("score", data, N)
            for _ in range(count):
                prog.run()
                sink(idx, self.output_buffer(b)[:N])
                idx += 1


class _BucketedProgram:
    """graph A -> async all-reduce of the first bucket, overlapped with graph B -> second bucket -> optimiser graph"""

    def __init__(self, a, b, c, hook_async, bucket_first, bucket_second):
        self.a, self.b, self.c, self.hook = a, b, c, hook_async
        self.bucket_first, self.bucket_second = bucket_first, bucket_second

    @property
    def n_launches(self):
        return self.a.n_launches + self.b.n_launches + self.c.n_launches + 2

    @property
    def sched(self):
        return self.a.sched + self.b.sched + self.c.sched

    def run(self):
        if any(p.use_graph and p.graph is None for p in (self.a, self.b, self.c)):
            # first step: every sub-program does an eager warm-up and restores the state it touched (the gradient arena
            # included) before it is captured - nothing may be in flight on the communication stream meanwhile
            self.a.run()
            self.b.run()
            self.hook(self.bucket_first).wait()
            if self.bucket_second.numel():
                self.hook(self.bucket_second).wait()
            self.c.run()
            return
        self.a.run()
        h1 = self.hook(self.bucket_first)
        self.b.run()
        h2 = self.hook(self.bucket_second) if self.bucket_second.numel() else None
        h1.wait()
        if h2 is not None:
            h2.wait()
        self.c.run()

    def profile(self, reps=5):
        return self.a.profile(reps) + self.b.profile(reps) + self.c.profile(reps)


class _SplitProgram:
    """backward graph -> eager collective -> optimiser graph"""

    def __init__(self, before, exchange, after):
        self.before, self.exchange, self.after = before, exchange, after

    @property
    def n_launches(self):
        return self.before.n_launches + 1 + self.after.n_launches

    def run(self):
        self.before.run()
        self.exchange()
        self.after.run()

    def profile(self, reps=5):
        return self.before.profile(reps) + self.after.profile(reps)


class _Program:
    """An op list; captured into a CUDA graph on first use when graphs are enabled."""

    def __init__(self, sched, use_graph, state=()):
        self.sched = sched
        self.use_graph = use_graph
        self.state = list(state)
        self.graph = None

    def run_eager(self):
        for _, op in self.sched:
            op()

    @property
    def n_launches(self):
        return len(self.sched)

    def profile(self, reps=5):
        """eager run with a CUDA-event pair around every op: [(name, mean ms)] (state is restored afterwards)"""
        saved = [t.clone() for t in self.state]
        st = torch.cuda.current_stream()
        acc = [0.0] * len(self.sched)
        for _ in range(reps):
            evs = []
            for _, op in self.sched:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                op()
                b.record(st)
                evs.append((a, b))
            st.synchronize()
            for i, (a, b) in enumerate(evs):
                acc[i] += a.elapsed_time(b)
        with torch.no_grad():
            for t, c in zip(self.state, saved):
                t.copy_(c)
        return [(name, acc[i] / reps) for i, (name, _) in enumerate(self.sched)]

    def timeline(self, reps=50, stride=1):
        """in-graph cost of every op: the prefixes sched[:k] are captured as graphs and their replays timed; the
        difference between consecutive prefixes is what op k adds to the critical path of the captured step (warm,
        overlapped with the side streams - unlike eager per-op events or ncu's cold serialised durations).
        -> [(name, microseconds added)]; state is restored."""
        saved = [t.clone() for t in self.state]
        out, prev = [], 0.0
        ks = [k for k in range(1, len(self.sched) + 1) if k % stride == 0 or k == len(self.sched)]
        for k in ks:
            p = _Program(self.sched[:k], True, self.state)
            for _ in range(3):
                p.run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                p.run()
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / reps * 1e3
            out.append((self.sched[k - 1][0], us - prev))
            prev = us
        with torch.no_grad():
            for t, c in zip(self.state, saved):
                t.copy_(c)
        return out

    def run(self):
        if not self.use_graph:
            self.run_eager()
            return
        if self.graph is None:
            # one eager warm-up (loads every kernel outside of capture and surfaces launch errors with
            # a readable message); the state it touches is snapshotted and restored so it is invisible.
            saved = [t.clone() for t in self.state]
            self.run_eager()
            torch.cuda.current_stream().synchronize()
            with torch.no_grad():
                for t, c in zip(self.state, saved):
                    t.copy_(c)
            # Nothing may free CUDA objects while the capture is open: a garbage-collection pass that destroys the
            # CUDAGraph / events of an engine that went out of scope earlier invalidates it (seen as a rare
            # cudaErrorStreamCaptureInvalidated in long test sessions).  Collect now, keep the collector off during
            # the capture, and do not let calls of other threads count against this capture.
            import gc
            gc.collect()
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                        self.run_forked()
            finally:
                if gc_was_on:
                    gc.enable()
            torch.cuda.current_stream().wait_stream(s)
            self.graph = g
        self.graph.replay()

    # weight-gradient kernels only feed the optimiser: inside the captured graph they run on side streams,
    # concurrently with the input-gradient chain (most launches of this network fill a fraction of the GPU)
    SIDE_SUFFIXES = (".wgrad", ".dW")
    JOIN_BEFORE = ("adam", "adam+allreduce", "grad_allreduce")
    N_SIDE = int(os.environ.get("CAE_SIDE_STREAMS", "3"))     # 0: everything on one stream

    def run_forked(self):
        if self.N_SIDE <= 0:
            self.run_eager()
            return
        main = torch.cuda.current_stream()
        if not hasattr(self, "_side"):
            self._side = [torch.cuda.Stream() for _ in range(self.N_SIDE)]
        used, k = [], 0
        for name, op in self.sched:
            if name.endswith(self.SIDE_SUFFIXES):
                side = self._side[k % self.N_SIDE]
                k += 1
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    op()
                if side not in used:
                    used.append(side)
            else:
                if name in self.JOIN_BEFORE:
                    for side in used:
                        ev = torch.cuda.Event()
                        ev.record(side)
                        main.wait_event(ev)
                    used = []
                op()
        for side in used:
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)
