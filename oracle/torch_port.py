"""TEST INFRASTRUCTURE - plain-PyTorch (CPU, fp32) restatement of the reference hot path.

Written functionally (torch.nn.functional on explicit weight dictionaries) so that it shares no code
with either the reference's nn.Module classes or the product's containers/kernels.  Each function names
the reference lines it follows.  Checked against the live reference by tests/test_oracle_golden.py via
the fixtures that oracle/gen_golden.py wrote.

State-dict key layout (reference, probed - SURVEY section 8b):
  encoder: encoder_cnn.{3i}.{weight,bias}  encoder_cnn.{3i+1}.{weight,bias,running_mean,running_var,num_batches_tracked}
           encoder_lin.{0,2}.{weight,bias}
  decoder: decoder_lin.{0,2}.{weight,bias} decoder_conv.{3j}.{weight,bias} decoder_conv.{3j+1}.{...BN...} (none after the last)
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _bn(x, sd, prefix, training):
    """nn.BatchNorm2d forward (encoder.py:45, decoder.py:47): batch stats in training (and running-stat
    update, unbiased variance, momentum 0.1), running stats in eval."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    out = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if training:
        sd[prefix + ".num_batches_tracked"] += 1
    return out


def encoder_forward(sd, specs, x, training, trace=None):
    """Encoder.forward (encoder.py:60-64): [Conv2d -> BN -> ReLU] per layer, flatten, Linear-ReLU-Linear"""
    for i, sp in enumerate(specs):
        x = F.conv2d(x, sd[f"encoder_cnn.{3 * i}.weight"], sd[f"encoder_cnn.{3 * i}.bias"], stride=sp["stride"])
        if trace is not None:
            trace.append(x)
        x = F.relu(_bn(x, sd, f"encoder_cnn.{3 * i + 1}", training))
    x = x.flatten(1)
    x = F.relu(F.linear(x, sd["encoder_lin.0.weight"], sd["encoder_lin.0.bias"]))
    return F.linear(x, sd["encoder_lin.2.weight"], sd["encoder_lin.2.bias"])


def decoder_forward(sd, specs, z, training, trace=None):
    """Decoder.forward (decoder.py:73-78): Linear-ReLU-Linear, unflatten, [ConvT -> BN -> ReLU]..., ConvT, sigmoid"""
    x = F.relu(F.linear(z, sd["decoder_lin.0.weight"], sd["decoder_lin.0.bias"]))
    x = F.linear(x, sd["decoder_lin.2.weight"], sd["decoder_lin.2.bias"])
    c, h, w = specs[0]["input_dimensions"]
    x = x.view(-1, c, h, w)
    last = len(specs) - 1
    for j, sp in enumerate(specs):
        x = F.conv_transpose2d(x, sd[f"decoder_conv.{3 * j}.weight"], sd[f"decoder_conv.{3 * j}.bias"],
                               stride=sp["stride"], output_padding=sp["output_padding"])
        if trace is not None:
            trace.append(x)
        if j != last:
            x = F.relu(_bn(x, sd, f"decoder_conv.{3 * j + 1}", training))
    return torch.sigmoid(x)


def trainable_keys(sd):
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


class OracleModel:
    """weights + Adam state + the train/test/score loops of ConvAEModel (conv_ae_model.py:185-239,303-334)"""

    def __init__(self, enc_sd, dec_sd, spec, lr=1e-3, weight_decay=1e-5, decoupled=False, zero_dead_bias_grads=False,
                 dtype=torch.float32):
        """zero_dead_bias_grads: the bias of a conv that feeds a training-mode BatchNorm has an identically zero
        gradient; autograd returns rounding noise (~1e-9) there which Adam then amplifies into a random walk of
        that (output-irrelevant) bias.  The CUDA path writes the exact zero; set this flag to compare tightly."""
        self.zero_dead_bias_grads = zero_dead_bias_grads
        # dtype=torch.float64 gives the adjudicator for "which fp32 result is closer to the exact one" (tests only)
        self.enc = {k: v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone()
                    for k, v in enc_sd.items()}
        self.dec = {k: v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone()
                    for k, v in dec_sd.items()}
        self.spec = spec  # dict as written to spec.json
        self.params = []
        for sd in (self.enc, self.dec):
            for k in trainable_keys(sd):
                sd[k].requires_grad_(True)
                self.params.append(sd[k])
        opt = torch.optim.AdamW if decoupled else torch.optim.Adam
        self.optim = opt(self.params, lr=lr, weight_decay=weight_decay)  # conv_ae_model.py:310 / unet.py:457

    def forward(self, x, training, trace=None):
        z = encoder_forward(self.enc, self.spec["input_layers"], x, training, trace)
        return decoder_forward(self.dec, self.spec["output_layers"], z, training, trace)

    def train_step(self, x, y):
        """conv_ae_model.py:191-197"""
        loss = F.mse_loss(self.forward(x, True), y)
        self.optim.zero_grad()
        loss.backward()
        if self.zero_dead_bias_grads:
            for sd in (self.enc, self.dec):
                for k in sd:
                    if k.endswith(".bias") and k.split(".")[0] in ("encoder_cnn", "decoder_conv"):
                        stem, idx = k.split(".")[0], int(k.split(".")[1])
                        if f"{stem}.{idx + 1}.running_mean" in sd and idx % 3 == 0:
                            sd[k].grad.zero_()
        self.optim.step()
        return loss.detach()

    def train_epoch(self, batches):
        """conv_ae_model.py:185-203: mean of the per-batch mean losses"""
        return float(np.mean([self.train_step(x, y).numpy() for x, y in batches]))

    def test_epoch(self, batches):
        """conv_ae_model.py:205-221 (eval mode: running statistics, no grad)"""
        with torch.no_grad():
            return float(np.mean([F.mse_loss(self.forward(x, False), y).numpy() for x, y in batches]))

    def score(self, x):
        with torch.no_grad():
            return self.forward(x, False)


def make_batches(X, Y, order, batch_size):
    """the frozen, once-shuffled batch list of conv_ae_model.py:291-325 (last ragged batch kept)"""
    X, Y = torch.as_tensor(X)[order], torch.as_tensor(Y)[order]
    return [(X[i:i + batch_size], Y[i:i + batch_size]) for i in range(0, X.shape[0], batch_size)]


def minmax(arr, lo, hi):
    """DSDataset.normalise_* (ds_dataset.py:99-113)"""
    return (arr - lo) / (hi - lo)


def shuffled_order(n, batch_size):
    """sample order of one pass over DataLoader(ds, batch_size, shuffle=True) drawing from torch's global RNG
    (conv_ae_model.py:291-292; consumed at :316 and :322)"""
    loader = torch.utils.data.DataLoader(range(n), batch_size=batch_size, shuffle=True)
    return [int(i) for idx in loader for i in idx]


class OracleVarModel(OracleModel):
    """Plain-PyTorch VAE on the reference's stacks (the reference ships no VarAEModel - parity unpinned; this is
    the definition the CUDA path is held to): heads `fc_mu` / `fc_logvar` on relu(encoder_lin.0(.)),
    z = mu + eps * exp(logvar / 2), loss = lambda_mse * MSE + lambda_kl * KL,
    KL = -1/2 * mean_n sum_l (1 + logvar - mu^2 - exp(logvar))."""

    def __init__(self, enc_sd, dec_sd, spec, lambda_mse=1.0, lambda_kl=1.0, **kw):
        super().__init__(enc_sd, dec_sd, spec, **kw)
        self.lambda_mse, self.lambda_kl = lambda_mse, lambda_kl

    def encode(self, x, training):
        sd = self.enc
        for i, sp in enumerate(self.spec["input_layers"]):
            x = F.conv2d(x, sd[f"encoder_cnn.{3 * i}.weight"], sd[f"encoder_cnn.{3 * i}.bias"], stride=sp["stride"])
            x = F.relu(_bn(x, sd, f"encoder_cnn.{3 * i + 1}", training))
        h = F.relu(F.linear(x.flatten(1), sd["encoder_lin.0.weight"], sd["encoder_lin.0.bias"]))
        return (F.linear(h, sd["fc_mu.weight"], sd["fc_mu.bias"]),
                F.linear(h, sd["fc_logvar.weight"], sd["fc_logvar.bias"]))

    def loss(self, x, y, eps, training):
        mu, lv = self.encode(x, training)
        z = mu + eps * torch.exp(0.5 * lv) if eps is not None else mu
        yhat = decoder_forward(self.dec, self.spec["output_layers"], z, training)
        kl = -0.5 * torch.sum(1 + lv - mu * mu - torch.exp(lv)) / x.shape[0]
        return self.lambda_mse * F.mse_loss(yhat, y) + self.lambda_kl * kl, yhat

    def train_step(self, x, y, eps):
        loss, _ = self.loss(x, y, eps, True)
        self.optim.zero_grad()
        loss.backward()
        if self.zero_dead_bias_grads:
            for sd in (self.enc, self.dec):
                for k in sd:
                    if k.endswith(".bias") and k.split(".")[0] in ("encoder_cnn", "decoder_conv"):
                        stem, idx = k.split(".")[0], int(k.split(".")[1])
                        if f"{stem}.{idx + 1}.running_mean" in sd and idx % 3 == 0:
                            sd[k].grad.zero_()
        self.optim.step()
        return loss.detach()

    def score(self, x):
        with torch.no_grad():
            mu, _ = self.encode(x, False)
            return decoder_forward(self.dec, self.spec["output_layers"], mu, False)


# ---------------------------------------------------------------------------------------------------------------
# UNET variant (reference: src/cae_tools/models/unet.py)
# ---------------------------------------------------------------------------------------------------------------
def masked_mse_loss(pred, target, mask):
    """unet.py:635-639"""
    diff = (pred - target) * mask
    return torch.sum(diff ** 2) / torch.sum(mask)


def pearson_corr(pred, target, mask):
    """unet.py:641-678: masked Pearson correlation per (sample, channel)"""
    d = pred.reshape(pred.size(0), pred.size(1), -1)
    t = target.reshape(target.size(0), target.size(1), -1)
    m = mask.reshape(mask.size(0), mask.size(1), -1).float()
    msum = torch.sum(m, dim=2, keepdim=True)
    mean_d = torch.sum(d * m, dim=2, keepdim=True) / (msum + 1e-8)
    mean_t = torch.sum(t * m, dim=2, keepdim=True) / (msum + 1e-8)
    std_d = torch.sqrt(torch.sum(m * (d - mean_d) ** 2, dim=2, keepdim=True) / (msum + 1e-8) + 1e-8)
    std_t = torch.sqrt(torch.sum(m * (t - mean_t) ** 2, dim=2, keepdim=True) / (msum + 1e-8) + 1e-8)
    num = torch.sum(m * ((d - mean_d) / std_d) * ((t - mean_t) / std_t), dim=2)
    return num / torch.sum(m, dim=2)


def unet_forward(enc, dec, spec, x, training, trace=None, drop=None):
    """Encoder.forward unet.py:102-112 + Decoder.forward unet.py:149-163.  drop(site, x) applies the dropout of the
    reference's nn.Dropout layers (unet.py:85,96,99,125,128,145) with an explicit mask; None = p 0.  The skip list holds
    the in-place ReLU outputs, i.e. the activations BEFORE dropout (unet.py:84-85,108-109)."""
    dr = (lambda site, t: t) if (drop is None or not training) else drop
    skips = []
    for i, sp in enumerate(spec["input_layers"]):
        x = F.conv2d(x, enc[f"encoder_cnn.{4 * i}.weight"], enc[f"encoder_cnn.{4 * i}.bias"], stride=sp["stride"],
                     padding=sp["output_padding"])
        if trace is not None:
            trace["enc"].append(x)
        x = F.relu(_bn(x, enc, f"encoder_cnn.{4 * i + 1}", training))
        skips.append(x)
        x = dr(i, x)
    skips.pop()
    x = x.flatten(1)
    x = F.linear(x, enc["encoder_lin.0.weight"], enc["encoder_lin.0.bias"])
    x = dr(4, F.relu(_bn(x, enc, "encoder_lin.1", training)))
    x = F.relu(F.linear(x, enc["encoder_lin.4.weight"], enc["encoder_lin.4.bias"]))
    if trace is not None:
        trace["z"] = x
    x = dr(5, x)
    x = F.linear(x, dec["decoder_lin.0.weight"], dec["decoder_lin.0.bias"])
    x = dr(6, F.relu(_bn(x, dec, "decoder_lin.1", training)))
    x = dr(7, F.relu(F.linear(x, dec["decoder_lin.4.weight"], dec["decoder_lin.4.bias"])))
    c, h, w = spec["output_layers"][0]["input_dimensions"]
    x = x.view(-1, c, h, w)
    skips = skips[::-1]
    for j, sp in enumerate(spec["output_layers"]):
        x = F.conv_transpose2d(x, dec[f"decoder_conv.{4 * j}.weight"], dec[f"decoder_conv.{4 * j}.bias"],
                               stride=sp["stride"], padding=sp["output_padding"])
        if trace is not None:
            trace["dec"].append(x)
        if j < len(skips):
            w1, w2 = dec[f"attention_layers.{j}.fc1.weight"], dec[f"attention_layers.{j}.fc2.weight"]
            avg, mx = x.mean(dim=(2, 3), keepdim=True), x.amax(dim=(2, 3), keepdim=True)
            att = torch.sigmoid(F.conv2d(F.relu(F.conv2d(avg, w1)), w2) + F.conv2d(F.relu(F.conv2d(mx, w1)), w2))
            x = torch.cat((x * att, skips[j]), 1)
            x = dr(8 + j, F.relu(_bn(x, dec, f"decoder_conv.{4 * j + 1}", training)))
    return torch.sigmoid(x)


class OracleUNet:
    """UNET.__train_epoch / __test_epoch / score (unet.py:295-380) with AdamW (unet.py:457), dropout 0"""

    def __init__(self, enc_sd, dec_sd, spec, lr=1e-3, weight_decay=1e-5, lambda_pearson=1.0, zero_dead_bias_grads=False,
                 dtype=torch.float32):
        # dtype=torch.float64: the adjudicator for "which fp32 result is closer to the exact one" (tests only)
        clone = lambda sd: {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone())
                            for k, v in sd.items()}
        self.enc, self.dec, self.spec = clone(enc_sd), clone(dec_sd), spec
        self.lambda_pearson = lambda_pearson
        self.zero_dead_bias_grads = zero_dead_bias_grads
        self.params = []
        for sd in (self.enc, self.dec):
            for k in trainable_keys(sd):
                sd[k].requires_grad_(True)
                self.params.append(sd[k])
        self.optim = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)

    def set_dropout(self, p, seed):
        """dropout with the CUDA path's counter-based masks (oracle/dropout_hash.py); the step counter advances with
        every train_step like the device-side one"""
        self.dropout_p, self.dropout_seed, self.steps_done = float(p), int(seed), 0

    def _drop(self):
        p = getattr(self, "dropout_p", 0.0)
        if p <= 0:
            return None
        from .dropout_hash import drop_mask

        def apply(site, t):
            n = t.shape[0]
            m = drop_mask(p, self.dropout_seed, self.steps_done, site, n, t[0].numel())
            return t * torch.from_numpy(m).to(t.dtype).view(t.shape)
        return apply

    def losses(self, x, y, mask, training):
        yhat = unet_forward(self.enc, self.dec, self.spec, x, training, drop=self._drop())
        mse = masked_mse_loss(yhat, y, mask)
        pl = 1 - torch.mean(pearson_corr(yhat, y, mask))
        return mse, pl, yhat

    def train_step(self, x, y, mask):
        self.optim.zero_grad()
        mse, pl, _ = self.losses(x, y, mask, True)
        (mse + self.lambda_pearson * pl).backward()
        if self.zero_dead_bias_grads:
            for k in self.enc:      # encoder convs feed a BatchNorm directly: identically zero bias gradient
                if k.startswith("encoder_cnn") and k.endswith(".bias") and int(k.split(".")[1]) % 4 == 0:
                    self.enc[k].grad.zero_()
            for k in ("encoder_lin.0.bias",):
                self.enc[k].grad.zero_()
            self.dec["decoder_lin.0.bias"].grad.zero_()
        self.optim.step()
        self.steps_done = getattr(self, "steps_done", 0) + 1
        return float(mse.detach()), float(pl.detach())

    def score(self, x):
        with torch.no_grad():
            return unet_forward(self.enc, self.dec, self.spec, x, False)


# ---------------------------------------------------------------------------------------------------------------
# LinearModel (reference: src/cae_tools/models/linear.py:33-49, linear_model.py:142-184,236-242)
# ---------------------------------------------------------------------------------------------------------------
class OracleLinear:
    """Flatten - nn.Linear - Unflatten with MSELoss and Adam(lr, weight_decay): the reference's `--method linear`"""

    def __init__(self, sd, out_shape, lr=1e-3, weight_decay=1e-5):
        self.w = sd["linear.1.weight"].detach().clone().float().requires_grad_(True)
        self.b = sd["linear.1.bias"].detach().clone().float().requires_grad_(True)
        self.out_shape = tuple(out_shape)
        self.optim = torch.optim.Adam([self.w, self.b], lr=lr, weight_decay=weight_decay)

    def forward(self, x):
        return F.linear(x.flatten(1), self.w, self.b).view(-1, *self.out_shape)

    def train_step(self, x, y):
        loss = F.mse_loss(self.forward(x), y)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return float(loss.detach())

    def test_loss(self, x, y):
        with torch.no_grad():
            return float(F.mse_loss(self.forward(x), y))

    def score(self, x):
        with torch.no_grad():
            return self.forward(x)
