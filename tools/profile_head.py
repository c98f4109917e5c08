#!/usr/bin/env python
"""Short driver for ncu: a few training steps at batch 64 and a few score batches at batch 1024 of the unet workload
(small data set: ncu's kernel replay saves / restores all device memory).  Usage (on the GPU box):
    python tools/profile_head.py && ncu --set full --import-source on -k regex:k_ph_ -c 8 -o rep python tools/profile_head.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cae_tools_b200.engine.unet import UNetEngine  # noqa: E402

dev = torch.device("cuda", 0)
spec, enc, dec = bench.build_modules("unet")
eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, device=dev, use_graphs=False)
X, Y = torch.rand(128, *bench.IN_SHAPE, device=dev), torch.rand(128, *bench.OUT_SHAPE, device=dev)
data = eng.bind(X, Y, 64)
for _ in range(2):
    eng.train_epoch(data)
XA = torch.rand(1024, *bench.IN_SHAPE, device=dev)
eng.score_batches(eng.bind(XA, None, 1024), lambda i, y: None)
eng.score_batches(eng.bind(XA, None, 1024), lambda i, y: None)
torch.cuda.synchronize()
print("ok")
