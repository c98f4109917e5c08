#!/usr/bin/env python
"""ncu driver: a few eager training steps of the conv and unet workloads at batch 64 (for -k regex:k_fc_stack)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from cae_tools_b200.engine.convae import ConvAEEngine  # noqa: E402
from cae_tools_b200.engine.unet import UNetEngine  # noqa: E402

dev = torch.device("cuda", 0)
for method in ("unet", "conv"):
    spec, enc, dec = bench.build_modules(method)
    eng = (UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, device=dev, use_graphs=False) if method == "unet"
           else ConvAEEngine(enc, dec, device=dev, use_graphs=False))
    eng.use_fused_fc = True        # the one-launch fc bottleneck is off by default
    X, Y = torch.rand(64, *bench.IN_SHAPE, device=dev), torch.rand(64, *bench.OUT_SHAPE, device=dev)
    data = eng.bind(X, Y, 64)
    for _ in range(2):
        eng.train_epoch(data)
torch.cuda.synchronize()
print("ok")
