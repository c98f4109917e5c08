"""TEST INFRASTRUCTURE - the arithmetic of the hot path restated from its published definitions in numpy.

The reference delegates its arithmetic to PyTorch (not vendored in /root/reference; unpinned there, torch
2.11.0 in this image): nn.Conv2d / nn.ConvTranspose2d / nn.BatchNorm2d / nn.Linear / MSELoss / Adam.  These
are the textbook definitions (SURVEY.md section 8(a')), written with explicit loops over kernel taps so they
share nothing with torch or with the CUDA kernels.  float64 accumulation; small inputs only.
"""

import numpy as np


def conv2d(x, w, b, stride, pad=0):
    """y[n,co,oy,ox] = b[co] + sum_{ci,ky,kx} x[n,ci,oy*s+ky-p,ox*s+kx-p] * w[co,ci,ky,kx]  (encoder.py:43-44)"""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    N, Ci, H, W = x.shape
    Co, _, kh, kw = w.shape
    xp = np.zeros((N, Ci, H + 2 * pad, W + 2 * pad))
    xp[:, :, pad:pad + H, pad:pad + W] = x
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    y = np.zeros((N, Co, Ho, Wo))
    for ky in range(kh):
        for kx in range(kw):
            patch = xp[:, :, ky:ky + (Ho - 1) * stride + 1:stride, kx:kx + (Wo - 1) * stride + 1:stride]
            y += np.einsum("nchw,oc->nohw", patch, w[:, :, ky, kx])
    if b is not None:
        y += np.asarray(b, np.float64)[None, :, None, None]
    return y


def conv_transpose2d(x, w, b, stride, pad=0, output_padding=0):
    """scatter form: y[n,co,iy*s+ky-p,ix*s+kx-p] += x[n,ci,iy,ix] * w[ci,co,ky,kx]  (decoder.py:44-45)
    Hout = (Hin-1)*s - 2p + kh + output_padding"""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    N, Ci, H, W = x.shape
    _, Co, kh, kw = w.shape
    Hf, Wf = (H - 1) * stride + kh + output_padding, (W - 1) * stride + kw + output_padding
    y = np.zeros((N, Co, Hf, Wf))
    for ky in range(kh):
        for kx in range(kw):
            y[:, :, ky:ky + (H - 1) * stride + 1:stride, kx:kx + (W - 1) * stride + 1:stride] += \
                np.einsum("nchw,co->nohw", x, w[:, :, ky, kx])
    y = y[:, :, pad:Hf - pad, pad:Wf - pad]
    if b is not None:
        y = y + np.asarray(b, np.float64)[None, :, None, None]
    return y


def batch_norm_train(x, gamma, beta, running_mean, running_var, eps=1e-5, momentum=0.1):
    """biased variance for normalisation, unbiased for the running estimate (nn.BatchNorm2d)"""
    x = np.asarray(x, np.float64)
    M = x.shape[0] * x.shape[2] * x.shape[3]
    mean = x.mean(axis=(0, 2, 3))
    var = x.var(axis=(0, 2, 3))
    y = (x - mean[None, :, None, None]) / np.sqrt(var + eps)[None, :, None, None]
    y = y * np.asarray(gamma, np.float64)[None, :, None, None] + np.asarray(beta, np.float64)[None, :, None, None]
    new_rm = (1 - momentum) * np.asarray(running_mean, np.float64) + momentum * mean
    new_rv = (1 - momentum) * np.asarray(running_var, np.float64) + momentum * var * M / (M - 1)
    return y, new_rm, new_rv


def batch_norm_eval(x, gamma, beta, running_mean, running_var, eps=1e-5):
    sh = (1, -1, 1, 1)
    x = np.asarray(x, np.float64)
    return (x - np.reshape(running_mean, sh)) / np.sqrt(np.reshape(running_var, sh) + eps) * np.reshape(gamma, sh) \
        + np.reshape(beta, sh)


def mse(a, b):
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    return float(np.mean(d * d))


def adam_step(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=0.0, decoupled=False):
    """Adam with coupled L2 (torch.optim.Adam, conv_ae_model.py:310) or decoupled decay (AdamW, unet.py:457); t >= 1"""
    p, g, m, v = (np.asarray(a, np.float64) for a in (p, g, m, v))
    if decoupled:
        p = p * (1 - lr * wd)
    else:
        g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    p = p - (lr / (1 - b1 ** t)) * m / (np.sqrt(v) / np.sqrt(1 - b2 ** t) + eps)
    return p, m, v


# ---- UNET pieces (reference: src/cae_tools/models/unet.py) --------------------------------------------------------------
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-np.asarray(x, np.float64)))


def channel_attention(y, w1, w2):
    """att[n,c] = sigmoid(W2 relu(W1 avg) + W2 relu(W1 max)), avg / max over (H, W); W1 [Cr,C], W2 [C,Cr] are the 1x1
    convolutions without bias of ChannelAttention (unet.py:23-39).  Returns att with shape [N, C, 1, 1]."""
    y = np.asarray(y, np.float64)
    w1 = np.asarray(w1, np.float64).reshape(w1.shape[0], -1)
    w2 = np.asarray(w2, np.float64).reshape(w2.shape[0], -1)
    avg, mx = y.mean(axis=(2, 3)), y.max(axis=(2, 3))
    branch = lambda v: np.maximum(v @ w1.T, 0.0) @ w2.T
    return sigmoid(branch(avg) + branch(mx))[:, :, None, None]


def masked_mse(pred, target, mask):
    """sum(((pred - target) * mask)^2) / sum(mask)   (unet.py:635-639)"""
    pred, target, mask = (np.asarray(a, np.float64) for a in (pred, target, mask))
    d = (pred - target) * mask
    return float(np.sum(d * d) / np.sum(mask))


def pearson_corr(pred, target, mask):
    """masked Pearson correlation per (sample, channel) plane, the +1e-8 terms exactly where the reference puts them
    (unet.py:641-678); loops over planes, no broadcasting tricks.  Returns [N, C]."""
    pred, target, mask = (np.asarray(a, np.float64) for a in (pred, target, mask))
    N, C = pred.shape[:2]
    mask = np.broadcast_to(mask, pred.shape)
    out = np.zeros((N, C))
    for n in range(N):
        for c in range(C):
            d, t, m = pred[n, c].ravel(), target[n, c].ravel(), mask[n, c].ravel()
            M = m.sum()
            mu_d, mu_t = (d * m).sum() / (M + 1e-8), (t * m).sum() / (M + 1e-8)
            sd = np.sqrt((m * (d - mu_d) ** 2).sum() / (M + 1e-8) + 1e-8)
            st = np.sqrt((m * (t - mu_t) ** 2).sum() / (M + 1e-8) + 1e-8)
            out[n, c] = (m * ((d - mu_d) / sd) * ((t - mu_t) / st)).sum() / M
    return out


def linear(x, w, b=None):
    """y = x W^T + b  (nn.Linear)"""
    y = np.asarray(x, np.float64) @ np.asarray(w, np.float64).T
    return y if b is None else y + np.asarray(b, np.float64)
