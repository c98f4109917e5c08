"""Step program of the UNET variant (reference: src/cae_tools/models/unet.py:73-163 modules, :295-337 train step,
:635-678 loss, :457 AdamW).

Per decoder block (all but the last transposed conv):

    y = ConvT(in) + b          cae_conv_up (PLAIN)
    att = CA(y)                cae_plane_stats -> cae_channel_attention_fwd        (avg / max pool, 1x1 MLP, sigmoid)
    cat = [att*y ; skip]       two cae_ew_epilogue launches that write the halves of ONE 2C-channel buffer in place
                               (gate through the per-(n,c) multiplier; skip = BatchNorm+ReLU of the encoder output
                               applied on load) and accumulate the BatchNorm(2C) statistics on the way
    next input                 relu(BN_2C(cat)) is never materialised: applied on load by the consumer

Backward mirrors it: the consumer's dgrad launch masks by the ReLU and reduces the BatchNorm-backward sums of the
2C channels; the skip half flows into the encoder as an `addend` of the matching encoder dgrad launch, the gate
half goes through cae_plane_dot / cae_channel_attention_bwd / cae_gate_bwd.  BatchNorm1d of the fc stacks is the
same BatchNorm kernel on a [N, F, 1, 1] view.  Loss = masked MSE + lambda * (1 - mean Pearson).

Dropout: the reference places nn.Dropout after every ReLU.  p = 0 is implemented (identity); p > 0 raises in
training (eval / apply are unaffected: dropout is the identity there).
"""

from __future__ import annotations

import torch

from . import ops
from .convae import ConvAEEngine, DataBinding


class UNetEngine(ConvAEEngine):

    def __init__(self, encoder, decoder, lambda_pearson=1.0, dropout_rate=0.0, seed=None, **kw):
        kw.setdefault("decoupled", True)         # torch.optim.AdamW (unet.py:457)
        super().__init__(encoder, decoder, **kw)
        self.lambda_pearson = float(lambda_pearson)
        self.dropout_rate = float(dropout_rate)
        # dropout masks are a counter-based hash of (seed, optimiser step, site, sample, element): nothing is stored and a
        # replayed CUDA graph draws fresh masks every step.  Data-parallel ranks pass different seeds.
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self._stem_train = {}
        self.dec3 = self.decoder.conv_layers()   # [(convT, bn2c|None, attention|None)]
        if len(self.dec3) != len(self.enc_layers):
            raise ValueError("UNET needs as many decoder as encoder layers (skip connections pair them up)")
        for j, (conv, bn, att) in enumerate(self.dec3[:-1]):
            skip = self.enc_specs[len(self.enc_specs) - 2 - j].get_output_dimensions()
            if tuple(self.dec_specs[j].get_output_dimensions()) != tuple(skip):
                raise ValueError(f"decoder layer {j} output {self.dec_specs[j].get_output_dimensions()} does not match "
                                 f"the skip connection {skip}")

    # ------------------------------------------------------------------ data
    def bind(self, X, Y, batch_size, mask=None):
        data = super().bind(X, Y, batch_size)
        data.M = mask.to(self.device, torch.float32).contiguous() if mask is not None else None
        data.mse_scale = self._mse_scale(data)
        data.pearson = torch.zeros(data.n_batches, dtype=torch.float32, device=self.device)
        data.extra_state = [data.pearson]
        return data

    def _mse_scale(self, data):
        """Data parallelism with a mask: per batch, the valid pixels of this rank's share over those of the global batch (one
        all-reduce at bind time - masks are data).  With it the masked-MSE term each rank forms, SQ_local / CNT_local, is
        weighted so that the SUM over ranks is SQ_global / CNT_global exactly, whatever the land / sea split between the
        shares; without a mask (or without ranks) the constant count_scale = 1 / world is already exact -> None."""
        dp = getattr(self, "dp", None)
        if dp is None or data.M is None:
            return None
        return dp.mask_scales(data.M, data.batch_size)

    def batch_losses(self, data):
        return data.losses                        # the reference's history records the masked MSE term only

    # ------------------------------------------------------------------ buffers
    def _act_buffers(self, B):
        if B in self._bufs:
            return self._bufs[B]
        b = {}
        f = self._f32
        b["y_e"] = [f(B, *sp.get_output_dimensions()) for sp in self.enc_specs]
        b["dz_e"] = [f(B, *sp.get_output_dimensions()) for sp in self.enc_specs]
        lin, dlin = self.encoder.encoder_lin, self.decoder.decoder_lin
        fc_e, lat, fc_d = lin[0].out_features, lin[4].out_features, dlin[0].out_features
        c0, h0, w0 = self.dec_specs[0].get_input_dimensions()
        ce, he, we = self.enc_specs[-1].get_output_dimensions()
        for k, shape in (("t1", (B, fc_e)), ("dz1", (B, fc_e)), ("dt1", (B, fc_e)), ("da1", (B, fc_e)),
                         ("z", (B, lat)), ("dzl", (B, lat)),
                         ("t3", (B, fc_d)), ("dz3", (B, fc_d)), ("dt3", (B, fc_d)), ("da3", (B, fc_d)),
                         ("u", (B, c0, h0, w0)), ("du", (B, c0, h0, w0)), ("da", (B, ce, he, we))):
            b[k] = f(*shape)
        b["y_d"], b["catg"], b["dz_catg"], b["dy_d"] = [], [], [], []
        b["st"], b["att"], b["hid"], b["datt"], b["davg"], b["dmax"], b["psum"] = [], [], [], [], [], [], []
        for j, sp in enumerate(self.dec_specs[:-1]):
            C, H, W = sp.get_output_dimensions()
            cr = self.dec3[j][2].fc1.out_channels
            b["y_d"].append(f(B, C, H, W))
            b["dy_d"].append(f(B, C, H, W))
            b["catg"].append(f(B, 2 * C, H, W))
            b["dz_catg"].append(f(B, 2 * C, H, W))
            b["st"].append(f(B * C * 4))
            b["hid"].append(f(B * 2 * cr))
            for k in ("att", "datt", "davg", "dmax", "psum"):
                b[k].append(f(B, C))
        co, ho, wo = self.dec_specs[-1].get_output_dimensions()
        b["yhat"] = f(B, co, ho, wo)
        b["dzL"] = f(B, co, ho, wo)
        b["psL"] = f(B, co)
        b["moments"] = torch.zeros(B * co * 7, dtype=torch.float64, device=self.device)
        b["coef"] = f(B * co * 3)
        b["scalars"] = f(4)
        self._keep.append(b["moments"])
        self._bufs[B] = b
        return b

    def output_buffer(self, b):
        return b["yhat"]

    # ------------------------------------------------------------------ helpers
    def _bn_half(self, key, bn_mod, lo, hi, with_nbt):
        """CaeBN block for channels [lo, hi) of a BatchNorm2d(2C); scratch shared with the full block"""
        _, s = self._bn(key, bn_mod)
        G = self.g
        return ops.make_bn(hi - lo, bn_mod.eps, bn_mod.momentum, bn_mod.weight[lo:hi], bn_mod.bias[lo:hi],
                           bn_mod.running_mean[lo:hi], bn_mod.running_var[lo:hi],
                           bn_mod.num_batches_tracked if with_nbt else None,
                           scale=s[0][lo:hi], shift=s[1][lo:hi], mean=s[2][lo:hi], invstd=s[3][lo:hi],
                           dgamma=G(bn_mod.weight)[lo:hi], dbeta=G(bn_mod.bias)[lo:hi],
                           bwdA=s[4][lo:hi], bwdB=s[5][lo:hi], bwdC=s[6][lo:hi])

    def _x_src(self, data, N):
        X = data.X
        return ops.make_src(X[:data.batch_size] if X.shape[0] >= data.batch_size else X, cursor=data.cursor,
                            cursor_stride=data.batch_size * X[0].numel(), n=N)

    def _cursor_src(self, T, data, N):
        return ops.make_src(T[:data.batch_size] if T.shape[0] >= data.batch_size else T, cursor=data.cursor,
                            cursor_stride=data.batch_size * T[0].numel(), n=N)

    def _geom(self, sp):
        return ops.geom(sp.get_kernel_size(), sp.get_stride(), sp.get_output_padding())

    def _stats_epi(self, train, blk, Cn, bias=None):
        if train:
            return ops.make_epilogue(ops.EPI_STATS, bias=bias, partials=self._partials(Cn), ticket=self._ticket(), bn=blk)
        return ops.make_epilogue(ops.EPI_PLAIN, bias=bias)

    # eval mode: every layer before the last one in ONE launch (unet_stem_eval.cu).  Measured (B200, shipped spec, whole score
    # batch incl. the head, tools/eval_stem_probe.py): batch 256 - 80 us against 155 us for the 11-launch chain; 1024 - 135
    # against 294; 4096 - 492 against 820 (stem 190 + head 304).  (Round 1's version of the kernel lost to the chain at 4096 - 577 us for the stem
    # alone - and was capped at 2048 samples; its run-time tap loops were replaced by compile-time-K ones.)
    use_fused_stem = True
    fused_stem_max_batch = 1 << 30
    use_fused_attention = True  # one launch per decoder block and direction (attention_block.cu); False = unfused chain
    use_patch_head = True       # fused kernel==stride last layer (patch_head.cu); False = generic conv + loss kernels

    def _patch_head(self, b, N, data, src, conv, sp, final):
        """descriptor of the fused last layer when its geometry allows it (kernel == stride, pad 0), else None"""
        k, st, pad = sp.get_kernel_size(), sp.get_stride(), sp.get_output_padding()
        cin, hin, win = sp.get_input_dimensions()
        if not self.use_patch_head or isinstance(k, (tuple, list)) or \
                not ops.patch_head_supported(k, st, pad, cin, win):
            return None
        co = sp.get_output_dimensions()[0]
        if final == "yhat":
            return ops.make_patch_head(src, conv.weight, conv.bias, k, co)
        if "ph_moments" not in b:
            b["ph_moments"] = torch.zeros(b["yhat"].shape[0] * co * hin * (k * k // 128) * 7, dtype=torch.float64,
                                          device=self.device)
            self._keep.append(b["ph_moments"])
        tgt = self._cursor_src(data.Y, data, N)
        msk = self._cursor_src(data.M, data, N) if data.M is not None else None
        mch = data.M.shape[1] if data.M is not None else co
        return ops.make_patch_head(src, conv.weight, conv.bias, k, co, target=tgt, mask=msk, mask_channels=mch,
                                   lambda_pearson=self.lambda_pearson, count_scale=self.count_scale,
                                   moments=b["ph_moments"], coef=b["coef"], scalars=b["scalars"], loss_out=data.losses,
                                   pearson_out=data.pearson, ticket=self._ticket(), mse_scale=data.mse_scale)

    def _last_layer_ops(self, S, b, N, data, src, conv, sp, j, final):
        """last transposed conv + sigmoid (+ loss): the fused patch head when the geometry allows it, else generic"""
        g = self._geom(sp)
        head = self._patch_head(b, N, data, src, conv, sp, final)
        if head is not None:
            if final == "yhat":
                S.append((f"fwd.head{j}+sigmoid", lambda h=head, o=ops.view4(b["yhat"], N): ops.patch_head_fwd(h, o)))
            else:
                S.append((f"fwd.head{j}+sigmoid+loss", lambda h=head: ops.patch_head_fwd(h)))
            self._head = head if final == "loss_grad" else None
            return
        self._head = None
        S.append((f"fwd.convT{j}+sigmoid", lambda src=src, w=conv.weight, g=g, o=ops.view4(b["yhat"], N),
                  e=ops.make_epilogue(ops.EPI_SIGMOID, bias=conv.bias): ops.conv_up(src, w, g, o, e)))
        if final != "yhat":
            tgt = self._cursor_src(data.Y, data, N)
            msk = self._cursor_src(data.M, data, N) if data.M is not None else None
            mch = data.M.shape[1] if data.M is not None else b["yhat"].shape[1]
            dz = ops.view4(b["dzL"], N) if final == "loss_grad" else None
            S.append(("loss.masked_mse+pearson", lambda tgt=tgt, msk=msk, mch=mch, dz=dz: ops.masked_pearson_loss(
                ops.view4(b["yhat"], N), tgt, msk, mch, self.lambda_pearson, self.count_scale, b["moments"], b["coef"],
                b["scalars"], data.losses, data.pearson, dz, b["psL"] if dz is not None else None,
                mse_scale=data.mse_scale)))

    def _eval_stem(self):
        """descriptor of the fused eval-mode stem (cached), or None when the geometry does not fit the kernel"""
        if hasattr(self, "_stem"):
            return self._stem
        self._stem = None
        ks = lambda sp: sp.get_kernel_size()
        if any(isinstance(ks(sp), (tuple, list)) for sp in list(self.enc_specs) + list(self.dec_specs)):
            return None
        convs, fcs, ups = [], [], []
        for i, ((conv, bn), sp) in enumerate(zip(self.enc_layers, self.enc_specs)):
            _, s = self._bn(("e", i), bn)
            convs.append((*sp.get_input_dimensions(), *sp.get_output_dimensions(), ks(sp), sp.get_stride(),
                          sp.get_output_padding(), conv.weight, conv.bias, s[0], s[1]))
        lin, dlin = self.encoder.encoder_lin, self.decoder.decoder_lin
        _, s1 = self._bn(("l", 0), lin[1])
        _, s3 = self._bn(("l", 1), dlin[1])
        fcs.append((lin[0].in_features, lin[0].out_features, 1, lin[0].weight, lin[0].bias, s1[0], s1[1]))
        fcs.append((lin[4].in_features, lin[4].out_features, 1, lin[4].weight, lin[4].bias, None, None))
        fcs.append((dlin[0].in_features, dlin[0].out_features, 1, dlin[0].weight, dlin[0].bias, s3[0], s3[1]))
        fcs.append((dlin[4].in_features, dlin[4].out_features, 1, dlin[4].weight, dlin[4].bias, None, None))
        ne = len(self.enc_layers)
        for j, ((conv, bn, att), sp) in enumerate(zip(self.dec3[:-1], self.dec_specs[:-1])):
            _, s2 = self._bn(("d", j), bn)
            ups.append((*sp.get_input_dimensions(), *sp.get_output_dimensions(), ks(sp), sp.get_stride(),
                        sp.get_output_padding(), att.fc1.out_channels, ne - 2 - j, conv.weight, conv.bias,
                        att.fc1.weight, att.fc2.weight, s2[0], s2[1]))
        if not ups:
            return None
        stem = ops.make_unet_stem(convs, fcs, ups)
        if ops.unet_stem_supported(stem):
            self._stem = stem
        return self._stem

    # training: the whole stem in one forward and one backward cooperative launch (unet_stem_train.cu) whenever the
    # geometry fits (the last layer must be the fused patch head; <= 4 samples per CTA on 148 CTAs: batch <= 592)
    use_fused_train_stem = True

    def _train_stem(self, N):
        """CaeStemTrain descriptor for batch N (cached), or None when the stem stays on the per-layer kernels"""
        if N in self._stem_train:
            return self._stem_train[N]
        st = None
        ks = lambda sp: sp.get_kernel_size()
        tuple_k = any(isinstance(ks(sp), (tuple, list)) for sp in list(self.enc_specs) + list(self.dec_specs))
        sp_last = self.dec_specs[-1]
        cin_l, hin_l, win_l = sp_last.get_input_dimensions()
        if self.use_fused_train_stem and self.use_patch_head and not tuple_k and len(self.dec3) >= 2 and \
                ops.patch_head_supported(ks(sp_last), sp_last.get_stride(), sp_last.get_output_padding(), cin_l, win_l):
            G = self.g
            convs, fcs, ups = [], [], []
            for i, ((conv, bn), sp) in enumerate(zip(self.enc_layers, self.enc_specs)):
                blk, _ = self._bn(("e", i), bn, G(conv.bias))
                convs.append((*sp.get_input_dimensions(), *sp.get_output_dimensions(), ks(sp), sp.get_stride(),
                              sp.get_output_padding(), conv.weight, conv.bias, G(conv.weight), G(conv.bias), blk))
            lin, dlin = self.encoder.encoder_lin, self.decoder.decoder_lin
            blk1, _ = self._bn(("l", 0), lin[1])
            blk3, _ = self._bn(("l", 1), dlin[1])
            for L, blk in ((lin[0], blk1), (lin[4], None), (dlin[0], blk3), (dlin[4], None)):
                fcs.append((L.in_features, L.out_features, int(blk is not None), L.weight, L.bias, G(L.weight), G(L.bias), blk))
            ne = len(self.enc_layers)
            for j, ((conv, bn, att), sp) in enumerate(zip(self.dec3[:-1], self.dec_specs[:-1])):
                blk, _ = self._bn(("d", j), bn)
                ups.append((*sp.get_input_dimensions(), *sp.get_output_dimensions(), ks(sp), sp.get_stride(),
                            sp.get_output_padding(), att.fc1.out_channels, ne - 2 - j, conv.weight, conv.bias, att.fc1.weight,
                            att.fc2.weight, G(conv.weight), G(conv.bias), G(att.fc1.weight), G(att.fc2.weight), blk))
            # every stem parameter lives in the flat arena in front of the head's weight (parameter order: encoder,
            # decoder_lin, attention layers, decoder_conv with the head last)
            head_w = self.dec3[-1][0].weight
            end = (head_w.data_ptr() - self.arena.data_ptr()) // 4
            st = ops.make_stem_train(N, convs, fcs, ups, self.dropout_rate, self.seed, self.step_count, self.device,
                                     self.arena[:end])
        self._stem_train[N] = st
        return st

    # ------------------------------------------------------------------ forward
    def _forward_ops(self, b, N, data, train, final):
        S = []
        src = self._x_src(data, N)
        self._stem_active = None
        if train:
            st = self._train_stem(N)
            if st is not None:
                nd = len(self.dec3)
                S.append(("fwd.stem_train", lambda st=st, x=src: ops.stem_train_fwd(st, x)))
                self._last_layer_ops(S, b, N, data, ops.make_src(st.t_hin, n=N), self.dec3[-1][0], self.dec_specs[-1], nd - 1, final)
                assert self._head is not None
                self._stem_active = (st, src)
                return S
            if self.dropout_rate > 0:
                raise NotImplementedError("UNET training with dropout_rate > 0 needs the fused training stem (unet_stem_train.cu), "
                                          "which does not cover this geometry / batch size; use dropout_rate=0")
        if not train and self.use_fused_stem and N <= self.fused_stem_max_batch:
            stem = self._eval_stem()
            if stem is not None:
                nd = len(self.dec3)
                C2, H, W = self.dec_specs[nd - 1].get_input_dimensions()
                if "stem_out" not in b:
                    b["stem_out"] = self._f32(b["yhat"].shape[0], C2, H, W)
                S.append(("fwd.stem", lambda st=stem, x=src, o=ops.view4(b["stem_out"], N): ops.unet_stem_eval(st, x, o)))
                conv, sp = self.dec3[-1][0], self.dec_specs[-1]
                self._last_layer_ops(S, b, N, data, ops.make_src(b["stem_out"], n=N), conv, sp, nd - 1, final)
                return S
        for i, ((conv, bn), sp) in enumerate(zip(self.enc_layers, self.enc_specs)):
            y = b["y_e"][i]
            blk, s = self._bn(("e", i), bn, self.g(conv.bias))
            epi = self._stats_epi(train, blk, conv.out_channels, conv.bias)
            S.append((f"fwd.conv{i}", lambda src=src, w=conv.weight, g=self._geom(sp), o=ops.view4(y, N), e=epi:
                      ops.conv_down(src, w, g, o, e)))
            src = ops.make_src(y, k0=s[0], k2=s[1], relu=True, n=N)
        # ---- fc stacks: Linear - BatchNorm1d - ReLU - Linear - ReLU (twice)
        lin, dlin = self.encoder.encoder_lin, self.decoder.decoder_lin
        ce, he, we = self.enc_specs[-1].get_output_dimensions()
        flat = ce * he * we
        fc_e, lat, fc_d, flat0 = lin[0].out_features, lin[4].out_features, dlin[0].out_features, dlin[4].out_features
        s_last = self._bn_scratch[("e", len(self.enc_layers) - 1)]
        ylast = b["y_e"][-1]
        blk1, s1 = self._bn(("l", 0), lin[1])
        blk3, s3 = self._bn(("l", 1), dlin[1])
        t1v, t3v = b["t1"].view(-1, fc_e, 1, 1), b["t3"].view(-1, fc_d, 1, 1)
        fused_fc = self._fc_fused(N, flat, fc_e, lat, fc_d, flat0)
        if fused_fc:
            p = ops.make_fc_stack(N, ylast, (lin[0], lin[4], dlin[0], dlin[4]), b["t1"], b["z"], b["t3"], b["u"],
                                  a_k0=s_last[0], a_k2=s_last[1], a_hw=he * we, a_relu=True, bn1=blk1, bn3=blk3,
                                  train=train, relu_mid=True)
            S.append(("fwd.fcstack", lambda p=p: ops.fc_stack_fwd(p)))
        S_main, S = S, []          # the cae_gemm chain is collected on its own and dropped when the fused kernel runs
        S.append(("fwd.fc1", lambda: ops.gemm(N, fc_e, flat, ylast, flat, 1, lin[0].weight, 1, flat, b["t1"], fc_e, 1,
                                              a_k0=s_last[0], a_k2=s_last[1], a_hw=he * we, a_relu=True,
                                              bias=lin[0].bias)))
        if train:
            e1 = self._stats_epi(True, blk1, fc_e)
            S.append(("fwd.bn1d_e", lambda e=e1: ops.ew_epilogue(ops.make_src(t1v, n=N), ops.view4(t1v, N), e)))
        S.append(("fwd.fc2", lambda: ops.gemm(N, lat, fc_e, b["t1"], fc_e, 1, lin[4].weight, 1, fc_e, b["z"], lat, 1,
                                              a_k0=s1[0], a_k2=s1[1], a_hw=1, a_relu=True, bias=lin[4].bias,
                                              relu_out=True)))
        S.append(("fwd.fc3", lambda: ops.gemm(N, fc_d, lat, b["z"], lat, 1, dlin[0].weight, 1, lat, b["t3"], fc_d, 1,
                                              bias=dlin[0].bias)))
        if train:
            e3 = self._stats_epi(True, blk3, fc_d)
            S.append(("fwd.bn1d_d", lambda e=e3: ops.ew_epilogue(ops.make_src(t3v, n=N), ops.view4(t3v, N), e)))
        S.append(("fwd.fc4", lambda: ops.gemm(N, flat0, fc_d, b["t3"], fc_d, 1, dlin[4].weight, 1, fc_d, b["u"], flat0,
                                              1, a_k0=s3[0], a_k2=s3[1], a_hw=1, a_relu=True, bias=dlin[4].bias,
                                              relu_out=True)))
        S = S_main + ([] if fused_fc else S)
        # ---- decoder
        src = ops.make_src(b["u"], n=N)
        nd, ne = len(self.dec3), len(self.enc_layers)
        for j, ((conv, bn, att), sp) in enumerate(zip(self.dec3, self.dec_specs)):
            g = self._geom(sp)
            if j < nd - 1:
                C, H, W = sp.get_output_dimensions()
                y, cat = b["y_d"][j], b["catg"][j]
                S.append((f"fwd.convT{j}", lambda src=src, w=conv.weight, g=g, o=ops.view4(y, N),
                          e=ops.make_epilogue(ops.EPI_PLAIN, bias=conv.bias): ops.conv_up(src, w, g, o, e)))
                cr = att.fc1.out_channels
                i_skip = ne - 2 - j
                ss = self._bn_scratch[("e", i_skip)]
                skip_src = ops.make_src(b["y_e"][i_skip], k0=ss[0], k2=ss[1], relu=True, n=N)
                if self.use_fused_attention and ops.attention_block_supported(C, H, W, cr):
                    blk_full, _ = self._bn(("d", j), bn)
                    S.append((f"fwd.attblock{j}", lambda y=y, sk=skip_src, a=att, cr=cr, o=ops.view4(cat, N),
                              e=self._stats_epi(train, blk_full, 2 * C), st=b["st"][j], at=b["att"][j], hd=b["hid"][j]:
                              ops.attention_block_fwd(ops.view4(y, N), sk, a.fc1.weight, a.fc2.weight, cr, o, e, st, at, hd)))
                else:
                    S.append((f"fwd.planestats{j}", lambda y=y, st=b["st"][j]: ops.plane_stats(ops.view4(y, N), st)))
                    S.append((f"fwd.attention{j}", lambda st=b["st"][j], a=att, C=C, cr=cr, hw=H * W, at=b["att"][j],
                              hd=b["hid"][j]: ops.channel_attention_fwd(st, a.fc1.weight, a.fc2.weight, N, C, cr, hw, at, hd)))
                    half0 = self._bn_half(("d", j), bn, 0, C, True)
                    half1 = self._bn_half(("d", j), bn, C, 2 * C, False)
                    S.append((f"fwd.gate{j}", lambda y=y, at=b["att"][j], o=ops.view4(cat[:, :C], N),
                              e=self._stats_epi(train, half0, C): ops.ew_epilogue(ops.make_src(y, kn=at, n=N), o, e)))
                    S.append((f"fwd.skip{j}", lambda sk=skip_src, o=ops.view4(cat[:, C:], N),
                              e=self._stats_epi(train, half1, C): ops.ew_epilogue(sk, o, e)))
                s2 = self._bn_scratch[("d", j)]
                src = ops.make_src(cat, k0=s2[0], k2=s2[1], relu=True, n=N)
            else:
                self._last_layer_ops(S, b, N, data, src, conv, sp, j, final)
        return S

    # ------------------------------------------------------------------ backward
    def _backward_ops(self, b, N, data):
        S = []
        G = self.g
        nd, ne = len(self.dec3), len(self.enc_layers)
        skip_grad = {}      # encoder layer index -> CaeSrc of the gradient arriving through the skip connection
        co = self.dec_specs[-1].get_output_dimensions()[0]
        last_conv = self.dec3[-1][0]
        head = getattr(self, "_head", None)
        if getattr(self, "_stem_active", None) is not None:
            # fused stem: the head hands over the raw gradient of its (activated) input; masks, BatchNorm backward and
            # every other gradient happen inside the one backward launch
            st, xsrc = self._stem_active
            part = torch.zeros(ops.patch_head_partials_len(head), dtype=torch.float32, device=self.device)
            self._keep.append(part)
            S.append((f"bwd.head{nd - 1}", lambda h=head, o=ops.view4(st.t_dhin, N), e=ops.make_epilogue(ops.EPI_PLAIN), part=part:
                      ops.patch_head_bwd(h, o, e, part)))
            S.append((f"bwd.head{nd - 1}.wgrad", lambda h=head, cv=last_conv, part=part:
                      ops.patch_head_wgrad_reduce(h, G(cv.weight), G(cv.bias), part)))
            S.append(("bwd.stem_train", lambda st=st, x=xsrc: ops.stem_train_bwd(st, x)))
            return S
        if head is None:
            S.append(("bwd.convT_last.db", lambda: ops.sum_over_n(b["psL"], N, co, G(last_conv.bias))))
        for j in range(nd - 1, -1, -1):
            conv, bn, att = self.dec3[j]
            sp = self.dec_specs[j]
            g = self._geom(sp)
            if j == nd - 1 and head is not None:
                # fused: recompute yhat, dL/dz in registers -> weight, bias and input gradients in one pass
                if j > 0:
                    pbn = self.dec3[j - 1][1]
                    blk, _ = self._bn(("d", j - 1), pbn)
                    epi = ops.make_epilogue(ops.EPI_MASKSTATS, partials=self._partials(pbn.num_features),
                                            ticket=self._ticket(), bn=blk, act=b["catg"][j - 1], n=N)
                    out = ops.view4(b["dz_catg"][j - 1], N)
                else:
                    epi = ops.make_epilogue(ops.EPI_MASK, act=b["u"], n=N)
                    out = ops.view4(b["du"], N)
                part = torch.zeros(ops.patch_head_partials_len(head), dtype=torch.float32, device=self.device)
                self._keep.append(part)
                S.append((f"bwd.head{j}", lambda h=head, o=out, e=epi, part=part: ops.patch_head_bwd(h, o, e, part)))
                S.append((f"bwd.head{j}.wgrad", lambda h=head, cv=conv, part=part:
                          ops.patch_head_wgrad_reduce(h, G(cv.weight), G(cv.bias), part)))
                continue
            if j == nd - 1:
                dy = ops.make_src(b["dzL"], n=N)
            else:
                C, H, W = sp.get_output_dimensions()
                s2 = self._bn_scratch[("d", j)]
                dzc, cat = b["dz_catg"][j], b["catg"][j]
                gsrc = ops.make_src(dzc[:, :C], t1=cat[:, :C], k0=s2[4][:C], k1=s2[5][:C], k2=s2[6][:C], n=N)
                skip_grad[ne - 2 - j] = ops.make_src(dzc[:, C:], t1=cat[:, C:], k0=s2[4][C:], k1=s2[5][C:], k2=s2[6][C:],
                                                     n=N)
                cr = att.fc1.out_channels
                if self.use_fused_attention and ops.attention_block_supported(C, H, W, cr):
                    part = torch.zeros(ops.attention_block_partials_len(C, cr), dtype=torch.float32, device=self.device)
                    self._keep.append(part)
                    S.append((f"bwd.attblock{j}", lambda gs=gsrc, a=att, j=j, cr=cr, cv=conv, part=part, t=self._ticket():
                              ops.attention_block_bwd(gs, ops.view4(b["y_d"][j], N), b["att"][j], b["hid"][j], b["st"][j],
                                                      a.fc1.weight, a.fc2.weight, cr, ops.view4(b["dy_d"][j], N),
                                                      G(a.fc1.weight), G(a.fc2.weight), G(cv.bias), part, t)))
                else:
                    S.append((f"bwd.gate{j}.datt", lambda gs=gsrc, y=b["y_d"][j], o=b["datt"][j]:
                              ops.plane_dot(gs, ops.view4(y, N), o)))
                    S.append((f"bwd.attention{j}", lambda a=att, j=j, C=C, cr=cr, hw=H * W: ops.channel_attention_bwd(
                        b["datt"][j], b["att"][j], b["hid"][j], b["st"][j], a.fc1.weight, a.fc2.weight, N, C, cr, hw,
                        G(a.fc1.weight), G(a.fc2.weight), b["davg"][j], b["dmax"][j])))
                    S.append((f"bwd.gate{j}.dy", lambda gs=gsrc, j=j: ops.gate_bwd(
                        gs, b["att"][j], b["davg"][j], b["dmax"][j], b["st"][j], ops.view4(b["dy_d"][j], N), b["psum"][j])))
                    S.append((f"bwd.convT{j}.db", lambda j=j, C=C, cv=conv: ops.sum_over_n(b["psum"][j], N, C, G(cv.bias))))
                dy = ops.make_src(b["dy_d"][j], n=N)
            if j > 0:
                sp2 = self._bn_scratch[("d", j - 1)]
                x_in = ops.make_src(b["catg"][j - 1], k0=sp2[0], k2=sp2[1], relu=True, n=N)
            else:
                x_in = ops.make_src(b["u"], n=N)
            S.append((f"bwd.convT{j}.wgrad", self._wgrad_op(x_in, dy, g, G(conv.weight))))
            if j > 0:
                pbn = self.dec3[j - 1][1]
                blk, _ = self._bn(("d", j - 1), pbn)
                epi = ops.make_epilogue(ops.EPI_MASKSTATS, partials=self._partials(pbn.num_features),
                                        ticket=self._ticket(), bn=blk, act=b["catg"][j - 1], n=N)
                out = ops.view4(b["dz_catg"][j - 1], N)
            else:
                epi = ops.make_epilogue(ops.EPI_MASK, act=b["u"], n=N)      # ReLU after decoder_lin's last Linear
                out = ops.view4(b["du"], N)
            S.append((f"bwd.convT{j}.dgrad", lambda dy=dy, w=conv.weight, g=g, o=out, e=epi: ops.conv_down(dy, w, g, o, e)))
        # ---- fc stacks
        lin, dlin = self.encoder.encoder_lin, self.decoder.decoder_lin
        ce, he, we = self.enc_specs[-1].get_output_dimensions()
        flat = ce * he * we
        fc_e, lat, fc_d, flat0 = lin[0].out_features, lin[4].out_features, dlin[0].out_features, dlin[4].out_features
        s_last = self._bn_scratch[("e", ne - 1)]
        ylast = b["y_e"][-1]
        s1, s3 = self._bn_scratch[("l", 0)], self._bn_scratch[("l", 1)]
        blk1, _ = self._bn(("l", 0), lin[1])
        blk3, _ = self._bn(("l", 1), dlin[1])
        v4 = lambda t, F: t.view(-1, F, 1, 1)
        fused_fc = self._fc_fused(N, flat, fc_e, lat, fc_d, flat0)
        if fused_fc:
            p = ops.make_fc_stack(N, ylast, (lin[0], lin[4], dlin[0], dlin[4]), b["t1"], b["z"], b["t3"], b["u"],
                                  a_k0=s_last[0], a_k2=s_last[1], a_hw=he * we, a_relu=True, bn1=blk1, bn3=blk3,
                                  train=True, relu_mid=True, du=b["du"], grads=G, dA=b["da"])
            S.append(("bwd.fcstack", lambda p=p: ops.fc_stack_bwd(p)))
        S_main, S = S, []          # (same pattern as in the forward schedule)
        # decoder_lin.4: u = relu(a3 W^T + b), a3 = relu(bn1d(t3))
        S.append(("bwd.fc4.dW", lambda: ops.gemm(flat0, fc_d, N, b["du"], 1, flat0, b["t3"], fc_d, 1, G(dlin[4].weight),
                                                 fc_d, 1, b_k0=s3[0], b_k2=s3[1], b_hw=1, b_relu=True,
                                                 rowsum_A=G(dlin[4].bias))))
        S.append(("bwd.fc4.dx", lambda: ops.gemm(N, fc_d, flat0, b["du"], flat0, 1, dlin[4].weight, fc_d, 1, b["da3"],
                                                 fc_d, 1)))
        e = ops.make_epilogue(ops.EPI_MASKSTATS, partials=self._partials(fc_d), ticket=self._ticket(), bn=blk3,
                              act=v4(b["t3"], fc_d), n=N)
        S.append(("bwd.bn1d_d.reduce", lambda e=e: ops.ew_epilogue(ops.make_src(v4(b["da3"], fc_d), n=N),
                                                                   ops.view4(v4(b["dz3"], fc_d), N), e)))
        S.append(("bwd.bn1d_d.apply", lambda: ops.ew_epilogue(
            ops.make_src(v4(b["dz3"], fc_d), t1=v4(b["t3"], fc_d), k0=s3[4], k1=s3[5], k2=s3[6], n=N),
            ops.view4(v4(b["dt3"], fc_d), N), ops.make_epilogue(ops.EPI_PLAIN))))
        # decoder_lin.0: t3 = z W^T + b, z = relu(...)
        # (the biases of the Linear layers that feed a BatchNorm1d have an identically zero gradient: never written,
        #  the gradient arena is zero there)
        S.append(("bwd.fc3.dW", lambda: ops.gemm(fc_d, lat, N, b["dt3"], 1, fc_d, b["z"], lat, 1, G(dlin[0].weight), lat,
                                                 1)))
        S.append(("bwd.fc3.dx", lambda: ops.gemm(N, lat, fc_d, b["dt3"], fc_d, 1, dlin[0].weight, lat, 1, b["dzl"], lat,
                                                 1, mask=b["z"])))
        # encoder_lin.4: z = relu(a1 W^T + b), a1 = relu(bn1d(t1))
        S.append(("bwd.fc2.dW", lambda: ops.gemm(lat, fc_e, N, b["dzl"], 1, lat, b["t1"], fc_e, 1, G(lin[4].weight), fc_e,
                                                 1, b_k0=s1[0], b_k2=s1[1], b_hw=1, b_relu=True,
                                                 rowsum_A=G(lin[4].bias))))
        S.append(("bwd.fc2.dx", lambda: ops.gemm(N, fc_e, lat, b["dzl"], lat, 1, lin[4].weight, fc_e, 1, b["da1"], fc_e,
                                                 1)))
        e = ops.make_epilogue(ops.EPI_MASKSTATS, partials=self._partials(fc_e), ticket=self._ticket(), bn=blk1,
                              act=v4(b["t1"], fc_e), n=N)
        S.append(("bwd.bn1d_e.reduce", lambda e=e: ops.ew_epilogue(ops.make_src(v4(b["da1"], fc_e), n=N),
                                                                   ops.view4(v4(b["dz1"], fc_e), N), e)))
        S.append(("bwd.bn1d_e.apply", lambda: ops.ew_epilogue(
            ops.make_src(v4(b["dz1"], fc_e), t1=v4(b["t1"], fc_e), k0=s1[4], k1=s1[5], k2=s1[6], n=N),
            ops.view4(v4(b["dt1"], fc_e), N), ops.make_epilogue(ops.EPI_PLAIN))))
        # encoder_lin.0: t1 = a W^T + b, a = relu(bn(y_last)) flattened
        S.append(("bwd.fc1.dW", lambda: ops.gemm(fc_e, flat, N, b["dt1"], 1, fc_e, ylast, flat, 1, G(lin[0].weight), flat,
                                                 1, b_k0=s_last[0], b_k2=s_last[1], b_hw=he * we, b_relu=True)))
        S.append(("bwd.fc1.dx", lambda: ops.gemm(N, flat, fc_e, b["dt1"], fc_e, 1, lin[0].weight, flat, 1, b["da"], flat,
                                                 1)))
        S = S_main + ([] if fused_fc else S)
        # ---- encoder
        conv, bn = self.enc_layers[ne - 1]
        blk, _ = self._bn(("e", ne - 1), bn, G(conv.bias))
        epi = ops.make_epilogue(ops.EPI_MASKSTATS, partials=self._partials(conv.out_channels), ticket=self._ticket(),
                                bn=blk, act=ylast, n=N)
        S.append(("bwd.enc_last.mask+bnsums", lambda s=ops.make_src(b["da"], n=N), o=ops.view4(b["dz_e"][ne - 1], N),
                  e=epi: ops.ew_epilogue(s, o, e)))
        for i in range(ne - 1, -1, -1):
            conv, bn = self.enc_layers[i]
            sp = self.enc_specs[i]
            g = self._geom(sp)
            s = self._bn_scratch[("e", i)]
            dy = ops.make_src(b["dz_e"][i], t1=b["y_e"][i], k0=s[4], k1=s[5], k2=s[6], n=N)
            if i > 0:
                sprev = self._bn_scratch[("e", i - 1)]
                x_in = ops.make_src(b["y_e"][i - 1], k0=sprev[0], k2=sprev[1], relu=True, n=N)
            else:
                x_in = self._x_src(data, N)
            S.append((f"bwd.conv{i}.wgrad", self._wgrad_op(dy, x_in, g, G(conv.weight))))
            if i > 0:
                pconv, pbn = self.enc_layers[i - 1]
                blk, _ = self._bn(("e", i - 1), pbn, G(pconv.bias))
                epi = ops.make_epilogue(ops.EPI_MASKSTATS, partials=self._partials(pconv.out_channels),
                                        ticket=self._ticket(), bn=blk, act=b["y_e"][i - 1], n=N,
                                        addend=skip_grad.get(i - 1))
                S.append((f"bwd.conv{i}.dgrad", lambda dy=dy, w=conv.weight, g=g, o=ops.view4(b["dz_e"][i - 1], N),
                          e=epi: ops.conv_up(dy, w, g, o, e)))
        return S

    def _eval_prepare_op(self):
        if self._bn_table is None:
            blocks = [self._bn(("e", i), bn)[0] for i, (conv, bn) in enumerate(self.enc_layers)]
            blocks.append(self._bn(("l", 0), self.encoder.encoder_lin[1])[0])
            blocks.append(self._bn(("l", 1), self.decoder.decoder_lin[1])[0])
            for j, (conv, bn, att) in enumerate(self.dec3):
                if bn is not None:
                    blocks.append(self._bn(("d", j), bn)[0])
            self._bn_count = len(blocks)
            self._bn_table = ops.bn_table(blocks, self.device)
        return lambda: ops.bn_eval_prepare(self._bn_table, self._bn_count)
