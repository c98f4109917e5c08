"""Step program of the variational variant: same conv stacks as ConvAEEngine, a different bottleneck.

    h1 = relu(W1 a + b1);  mu = Wm h1 + bm;  logvar = Wv h1 + bv;  z = mu + eps * exp(logvar / 2)
    loss = lambda_mse * MSE(yhat, y) + lambda_kl * KL,   KL = -1/2 mean_n sum_l (1 + logvar - mu^2 - exp(logvar))

The reference names this variant (`--method var`, `--lambda-mse`, `--lambda-kl`: cli/train_cae.py:32-33,42;
`VarAEModel`: model_evaluator.py:35) but ships no implementation, so this is the textbook VAE on the
reference's encoder / decoder stacks (parity unpinned).  eps is drawn on the device by a
counter-based generator keyed on the step counter (fresh noise on every graph replay), or supplied
(`fixed_eps`) for oracle comparisons.  Eval mode uses z = mu.
"""

from __future__ import annotations

import torch

from . import ops
from .convae import ConvAEEngine


class VarAEEngine(ConvAEEngine):

    def __init__(self, encoder, decoder, lambda_mse=1.0, lambda_kl=1.0, seed=0, **kw):
        super().__init__(encoder, decoder, **kw)
        self.mse_weight = float(lambda_mse)
        self.lambda_kl = float(lambda_kl)
        self.seed = int(seed)

    def bind(self, X, Y, batch_size, fixed_eps=None):
        data = super().bind(X, Y, batch_size)
        data.kl = torch.zeros(data.n_batches, dtype=torch.float32, device=self.device)
        data.extra_state = [data.kl]
        data.fixed_eps = fixed_eps.to(self.device, torch.float32).contiguous() if fixed_eps is not None else None
        return data

    def batch_losses(self, data):
        return data.losses + data.kl        # lambda_mse * mse + lambda_kl * kl (both already weighted)

    def _fc_buffers(self, b, B):
        enc, dlin = self.encoder, self.decoder.decoder_lin
        fc, lat = enc.encoder_lin[0].out_features, enc.fc_mu.out_features
        for k, n in (("h1", fc), ("mu", lat), ("lv", lat), ("z", lat), ("eps", lat), ("dzl", lat), ("dmu", lat),
                     ("dlv", lat), ("dh1", fc), ("dh1a", fc), ("dh1b", fc), ("h3", dlin[0].out_features),
                     ("dh3", dlin[0].out_features)):
            b[k] = self._f32(B, n)

    def _fc_forward_ops(self, b, N, data, train):
        enc, dlin = self.encoder, self.decoder.decoder_lin
        lin1, fmu, flv = enc.encoder_lin[0], enc.fc_mu, enc.fc_logvar
        ylast = b["y_e"][-1]
        ce, he, we = self.enc_specs[-1].get_output_dimensions()
        flat = ce * he * we
        s_last = self._bn_scratch[("e", len(self.enc_layers) - 1)]
        fc, lat, fc2, out4 = lin1.out_features, fmu.out_features, dlin[0].out_features, dlin[2].out_features
        sched = [
            ("fwd.fc1", lambda: ops.gemm(N, fc, flat, ylast, flat, 1, lin1.weight, 1, flat, b["h1"], fc, 1,
                                         a_k0=s_last[0], a_k2=s_last[1], a_hw=he * we, a_relu=True, bias=lin1.bias,
                                         relu_out=True)),
            ("fwd.fc_mu", lambda: ops.gemm(N, lat, fc, b["h1"], fc, 1, fmu.weight, 1, fc, b["mu"], lat, 1,
                                           bias=fmu.bias)),
            ("fwd.fc_logvar", lambda: ops.gemm(N, lat, fc, b["h1"], fc, 1, flv.weight, 1, fc, b["lv"], lat, 1,
                                               bias=flv.bias)),
        ]
        fixed = getattr(data, "fixed_eps", None)
        kl_scale = self.lambda_kl * self.count_scale
        if train and fixed is None:
            sched.append(("fwd.randn", lambda: ops.randn(b["eps"], N * lat, self.seed, self.step_count)))
            # (the cursor selects the KL slot of this batch; the noise buffer is per step: stride 0)
            sched.append(("fwd.reparam+kl", lambda: ops.vae_reparam_fwd(b["mu"], b["lv"], b["eps"], 0, data.cursor, b["z"], N,
                                                                        lat, True, kl_scale, data.kl)))
        elif train:
            sched.append(("fwd.reparam+kl", lambda: ops.vae_reparam_fwd(b["mu"], b["lv"], fixed,
                                                                        data.batch_size * lat, data.cursor, b["z"], N,
                                                                        lat, True, kl_scale, data.kl)))
        else:
            sched.append(("fwd.reparam+kl", lambda: ops.vae_reparam_fwd(b["mu"], b["lv"], None, 0, data.cursor, b["z"],
                                                                        N, lat, False, kl_scale, data.kl)))
        sched += [
            ("fwd.fc3", lambda: ops.gemm(N, fc2, lat, b["z"], lat, 1, dlin[0].weight, 1, lat, b["h3"], fc2, 1,
                                         bias=dlin[0].bias, relu_out=True)),
            ("fwd.fc4", lambda: ops.gemm(N, out4, fc2, b["h3"], fc2, 1, dlin[2].weight, 1, fc2, b["u"], out4, 1,
                                         bias=dlin[2].bias)),
        ]
        return sched

    def _fc_backward_ops(self, b, N, data):
        enc, dlin = self.encoder, self.decoder.decoder_lin
        lin1, fmu, flv = enc.encoder_lin[0], enc.fc_mu, enc.fc_logvar
        ce, he, we = self.enc_specs[-1].get_output_dimensions()
        flat = ce * he * we
        fc, lat, fc2, out4 = lin1.out_features, fmu.out_features, dlin[0].out_features, dlin[2].out_features
        G = self.g
        s_last = self._bn_scratch[("e", len(self.enc_layers) - 1)]
        ylast = b["y_e"][-1]
        fixed = getattr(data, "fixed_eps", None)
        eps, stride, cur = (b["eps"], 0, None) if fixed is None else (fixed, data.batch_size * lat, data.cursor)
        klw = self.lambda_kl * self.count_scale
        return [
            ("bwd.fc4.dW", lambda: ops.gemm(out4, fc2, N, b["du"], 1, out4, b["h3"], fc2, 1, G(dlin[2].weight), fc2, 1,
                                            rowsum_A=G(dlin[2].bias))),
            ("bwd.fc4.dx", lambda: ops.gemm(N, fc2, out4, b["du"], out4, 1, dlin[2].weight, fc2, 1, b["dh3"], fc2, 1,
                                            mask=b["h3"])),
            ("bwd.fc3.dW", lambda: ops.gemm(fc2, lat, N, b["dh3"], 1, fc2, b["z"], lat, 1, G(dlin[0].weight), lat, 1,
                                            rowsum_A=G(dlin[0].bias))),
            ("bwd.fc3.dx", lambda: ops.gemm(N, lat, fc2, b["dh3"], fc2, 1, dlin[0].weight, lat, 1, b["dzl"], lat, 1)),
            ("bwd.reparam", lambda: ops.vae_reparam_bwd(b["dzl"], b["mu"], b["lv"], eps, stride, cur, b["dmu"],
                                                        b["dlv"], N, lat, klw)),
            ("bwd.fc_mu.dW", lambda: ops.gemm(lat, fc, N, b["dmu"], 1, lat, b["h1"], fc, 1, G(fmu.weight), fc, 1,
                                              rowsum_A=G(fmu.bias))),
            ("bwd.fc_logvar.dW", lambda: ops.gemm(lat, fc, N, b["dlv"], 1, lat, b["h1"], fc, 1, G(flv.weight), fc, 1,
                                                  rowsum_A=G(flv.bias))),
            ("bwd.fc_mu.dx", lambda: ops.gemm(N, fc, lat, b["dmu"], lat, 1, fmu.weight, fc, 1, b["dh1a"], fc, 1,
                                              mask=b["h1"])),
            ("bwd.fc_logvar.dx", lambda: ops.gemm(N, fc, lat, b["dlv"], lat, 1, flv.weight, fc, 1, b["dh1b"], fc, 1,
                                                  mask=b["h1"])),
            ("bwd.fc_heads.sum", lambda: ops.add2(b["dh1a"], b["dh1b"], b["dh1"], N * fc)),
            ("bwd.fc1.dW", lambda: ops.gemm(fc, flat, N, b["dh1"], 1, fc, ylast, flat, 1, G(lin1.weight), flat, 1,
                                            b_k0=s_last[0], b_k2=s_last[1], b_hw=he * we, b_relu=True,
                                            rowsum_A=G(lin1.bias))),
            ("bwd.fc1.dx", lambda: ops.gemm(N, flat, fc, b["dh1"], fc, 1, lin1.weight, flat, 1, b["da"], flat, 1)),
        ]
