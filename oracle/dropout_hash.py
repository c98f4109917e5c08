"""TEST INFRASTRUCTURE (oracle): numpy restatement of the counter-based dropout mask of the fused UNET training stem
(cae_tools_b200/csrc/unet_stem_train.cu: st_mix / st_drop_init / st_keep).  The reference uses torch.nn.Dropout
(unet.py:85,96,99,125,128,145), whose CPU / CUDA random streams no kernel can reproduce; the CUDA path therefore defines
its own stateless generator, and this file lets the oracle apply EXACTLY the same masks so that forward values, losses and
every gradient with p > 0 are checked bit-for-bit in structure (1e-4 in value) against torch autograd.

    mask(site, n, e) = 1/(1-p)  if  (mix32(n*elems + e + mix32(key ^ site*0x85EBCA6B)) >> 8) >= floor(p * 2^24)  else 0
    key = mix32(seed_lo ^ mix32(seed_hi + step*0x9E3779B9))        step = optimiser steps completed before this one
    sites: encoder conv l -> l, fc layer i -> 4 + i, decoder block j -> 8 + j
"""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def mix32(x):
    x = np.asarray(x, dtype=np.uint64) & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7feb352d)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846ca68b)) & M32
    x ^= x >> np.uint64(16)
    return x


def drop_key(seed, step):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    lo, hi = seed & 0xFFFFFFFF, seed >> 32
    inner = mix32(np.uint64((hi + (int(step) * 0x9E3779B9)) & 0xFFFFFFFF))
    return int(mix32(np.uint64(lo) ^ inner))


def drop_mask(p, seed, step, site, n_samples, elems):
    """float32 [n_samples, elems]: 0 where dropped, 1/(1-p) where kept (all ones for p == 0)"""
    if p <= 0:
        return np.ones((n_samples, elems), dtype=np.float32)
    thresh = np.uint64(int(np.float32(p) * np.float32(16777216.0)))
    key = np.uint64(drop_key(seed, step))
    salt = mix32(key ^ np.uint64((site * 0x85EBCA6B) & 0xFFFFFFFF))
    idx = np.arange(n_samples * elems, dtype=np.uint64)
    h = mix32((idx + salt) & M32)
    keep = (h >> np.uint64(8)) >= thresh
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return (keep.astype(np.float32) * scale).reshape(n_samples, elems)
