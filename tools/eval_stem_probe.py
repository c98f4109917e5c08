#!/usr/bin/env python
"""apply() / score path of the unet: one batch through the fused eval stem (unet_stem_eval.cu) against the per-layer chain,
resident in HBM, CUDA-event timed.   python tools/eval_stem_probe.py [--batches 1024 4096]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cae_tools_b200.engine.unet import UNetEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=int, nargs="+", default=[1024, 4096])
ap.add_argument("--reps", type=int, default=30)
a = ap.parse_args()
dev = torch.device("cuda", 0)
for B in a.batches:
    ref = None
    for fused in (True, False):
        spec, enc, dec = bench.build_modules("unet")
        eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, device=dev)
        eng.use_fused_stem = fused
        eng.fused_stem_max_batch = 1 << 30
        X = torch.rand(2 * B, *bench.IN_SHAPE, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
        data = eng.bind(X, None, B)
        out = []
        eng.score_batches(data, lambda i, y: out.append(y[:8].clone()) if i == 0 else None)
        if ref is None:
            ref = out[0]
        else:
            print(f"   fused vs chain max abs diff of yhat: {float((ref - out[0]).abs().max()):.2e}")
        prog = eng.program("score", data, B)
        for _ in range(3):
            prog.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            prog.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print(f"batch {B} {'fused eval stem' if fused else 'per-layer chain'}: {ms * 1e3:.1f} us/batch, {B / ms * 1e3 / 1e6:.2f} M images/s, "
              f"launches {prog.n_launches}", flush=True)
        if fused:
            print("   " + ", ".join(f"{n} {t * 1e3:.1f} us" for n, t in prog.profile(reps=5)))
