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
