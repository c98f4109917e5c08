"""shared helpers for the test-suite (oracle side)"""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def split_sd(g, prefix):
    """{'init.enc.foo': arr} -> {'foo': tensor}"""
    return {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith(prefix)}


def spec_of(g):
    return json.loads(str(g["spec_json"]))


def rel_err(a, b):
    """max-norm relative error  ||a-b||inf / max(||b||inf, tiny)  """
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def pre_bn_bias_keys(sd_keys, prefix):
    """biases of convs that feed a training-mode BatchNorm: their true gradient is identically zero"""
    out = []
    for k in sd_keys:
        if k.endswith(".bias") and (k.startswith("encoder_cnn") or k.startswith("decoder_conv")):
            idx = int(k.split(".")[1])
            if f"{k.split('.')[0]}.{idx + 1}.running_mean" in sd_keys:
                out.append(k)
    return out


def unet_light_data(g):
    """x, y, mask of a `light` unet fixture: y / mask are regenerated from the recorded seed with torch's CPU generator
    (platform-independent) exactly as oracle/gen_golden.py:gen_unet drew them"""
    spec = spec_of(g)
    oh, ow = spec["output_layers"][-1]["output_dimensions"][1:]
    batch = int(g["batch"])
    gen = torch.Generator().manual_seed(int(g["data_seed"]))
    x = torch.rand(batch, 1, 16, 16, generator=gen)
    y = torch.rand(batch, 1, oh, ow, generator=gen)
    mask = (torch.rand(batch, 1, oh, ow, generator=gen) > 0.3).float() if int(g["with_mask"]) else torch.ones(batch, 1, oh, ow)
    assert np.array_equal(x.numpy(), g["x"]), "the generator no longer reproduces the recorded inputs"
    return x, y, mask
