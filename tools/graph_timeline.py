#!/usr/bin/env python
"""In-graph cost of every op of a step program: captures the prefixes ops[:k] as CUDA graphs and times their replays;
the difference between consecutive prefixes is what op k adds to the critical path inside the captured step
(per-kernel durations from ncu are cold-cache and serialised; this is the warm, overlapped picture).

    python tools/graph_timeline.py [--method unet|conv] [--batch 64] [--reps 200]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from cae_tools_b200.engine.convae import ConvAEEngine, _Program  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="unet")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--stride", type=int, default=1, help="time every stride-th prefix")
    ap.add_argument("--kind", default="train", choices=["train", "score", "test"])
    ap.add_argument("--main-only", action="store_true", help="drop the side-stream ops (weight gradients): the critical chain alone")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    spec, enc, dec = bench.build_modules(args.method)
    if args.method == "unet":
        from cae_tools_b200.engine.unet import UNetEngine
        eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, device=dev)
    else:
        eng = ConvAEEngine(enc, dec, device=dev)
    B = args.batch
    nb = 8 if args.kind == "train" else 2
    X = torch.rand(nb * B, *bench.IN_SHAPE, device=dev)
    Y = torch.rand(nb * B, *bench.OUT_SHAPE, device=dev)
    data = eng.bind(X, Y, B)
    if args.kind != "train":
        eng._eval_prepare_op()()
    full = eng._program(args.kind, data, B)
    sched = full.sched
    if args.main_only:
        sched = [(n, op) for n, op in sched if not n.endswith(_Program.SIDE_SUFFIXES)]
    prev = 0.0
    print(f"{'op':34s} {'cumulative us':>14s} {'adds us':>10s}")
    for k in range(1, len(sched) + 1):
        if k % args.stride and k != len(sched):
            continue
        p = _Program(sched[:k], True, full.state)
        for _ in range(5):
            p.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            p.run()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / args.reps * 1e3
        print(f"{sched[k - 1][0]:34s} {us:14.1f} {us - prev:10.1f}")
        prev = us


if __name__ == "__main__":
    main()
