#!/usr/bin/env python
"""SURVEY 8f row 2 regime: UNET with fc 3200 / latent 800 on the shipped 16x16 -> 256x256 spec at batch 256 - the training
step with the fc contractions on the tensor cores (tc_dense.cu) against the same step on the SIMT GEMM kernel.

  python tools/largefc_probe.py [--batch 256] [--fc 3200] [--latent 800] [--profile]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cae_tools_b200.engine import ops  # noqa: E402
from cae_tools_b200.engine.unet import UNetEngine  # noqa: E402
from cae_tools_b200.models.model_sizer import ModelSpec  # noqa: E402
from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--fc", type=int, default=3200)
    ap.add_argument("--latent", type=int, default=800)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--linear", action="store_true", help="LinearModel 16x16 -> 256x256 (256 -> 65 536) instead of the unet")
    a = ap.parse_args()
    if a.linear:
        return linear(a)
    spec = ModelSpec()
    with open(os.path.join(ROOT, "cae_tools_b200", "specs", "unet_16x16_256x256.json")) as f:
        spec.load(json.load(f))
    dev = torch.device("cuda")
    B = a.batch
    gen = torch.Generator(device=dev).manual_seed(1)
    X = torch.rand(4 * B, 1, 16, 16, device=dev, generator=gen)
    Y = torch.rand(4 * B, 1, 256, 256, device=dev, generator=gen)
    for tc in (True, False):
        ops.USE_TC_DENSE = tc
        torch.manual_seed(0)
        enc, dec = UNetEncoder(spec.get_input_layers(), a.latent, a.fc, 0.0), UNetDecoder(spec.get_output_layers(), a.latent, a.fc, 0.0)
        eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5, device=dev)
        data = eng.bind(X, Y, B)
        prog = eng.program("train", data, B)
        for _ in range(3):
            prog.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            prog.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print(f"unet fc {a.fc} latent {a.latent} batch {B}, fc GEMMs on {'tcgen05 (tc_dense)' if tc else 'SIMT k_gemm'}: "
              f"{ms:.3f} ms/step, {B / ms * 1e3:.0f} samples/s, launches {prog.n_launches}, loss {float(data.losses[0]):.6f}", flush=True)
        if a.profile:
            table = prog.profile(reps=2)
            tot = sum(t for _, t in table)
            for name, t in table:
                if "fc" in name or t / tot > 0.03:
                    print(f"  {name:30s} {t * 1e3:10.1f} us {100 * t / tot:5.1f}%")
            print(f"  sum {tot:.3f} ms")
    ops.USE_TC_DENSE = True


def linear(a):
    from cae_tools_b200.engine.linear import LinearEngine
    from cae_tools_b200.models.linear import Linear
    dev = torch.device("cuda")
    B = a.batch
    gen = torch.Generator(device=dev).manual_seed(1)
    X = torch.rand(2 * B, 1, 16, 16, device=dev, generator=gen)
    Y = torch.rand(2 * B, 1, 256, 256, device=dev, generator=gen)
    for tc in (True, False):
        ops.USE_TC_DENSE = tc
        torch.manual_seed(0)
        eng = LinearEngine(Linear((1, 16, 16), (1, 256, 256)), lr=1e-3, weight_decay=1e-5, device=dev)
        data = eng.bind(X, Y, B)
        for _ in range(2):
            eng.train_epoch(data)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            eng.train_epoch(data)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (a.steps * 2)
        print(f"linear 256 -> 65536 batch {B}, GEMMs on {'tcgen05 (tc_dense)' if tc else 'SIMT k_gemm'}: {ms:.3f} ms/step, "
              f"{B / ms * 1e3:.0f} samples/s, loss {float(data.losses[0]):.6f}", flush=True)
    ops.USE_TC_DENSE = True


if __name__ == "__main__":
    main()
