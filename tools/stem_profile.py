#!/usr/bin/env python
"""Phase timeline of the fused UNET training stem (clock64 of CTA 0) + CUDA-event time of each launch, unet batch 64."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cae_tools_b200._lib import lib  # noqa: E402
from cae_tools_b200.engine.unet import UNetEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
spec, enc, dec = bench.build_modules("unet")
eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=float(os.environ.get("DROPOUT", "0")), lr=1e-3, weight_decay=1e-5)
X, Y = torch.rand(8 * B, 1, 16, 16, device="cuda"), torch.rand(8 * B, 1, 256, 256, device="cuda")
data = eng.bind(X, Y, B)
prog = eng._program("train", data, B)
for _ in range(5):
    prog.run()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 64)()
lib().cae_unet_stem_train_profile(buf)
t = list(buf)
for name, lo in (("forward", 0), ("backward", 32)):
    n = int(t[lo + 31]) - lo
    ts = t[lo:lo + n]
    print(f"{name}: total {(ts[-1] - ts[0]) / 1965.0:.1f} us over {n} stamps")
    print("  deltas (us):", " ".join(f"{(b - a) / 1965.0:.1f}" for a, b in zip(ts, ts[1:])))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    prog.run()
e1.record()
torch.cuda.synchronize()
print(f"step (graph): {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
for name, ms in prog.profile(reps=5):
    print(f"  {name:28s} {ms * 1e3:8.1f} us")
