#!/usr/bin/env python
"""BASELINE configs[3] probe: synthetic 4x64x64 -> 4x1024x1024 ConvAE (spec from create_model_spec).

  python tools/conv4_probe.py --batch 2 --check      one optimiser step against the oracle port (CPU, slow: small batch)
  python tools/conv4_probe.py --batch 128 --profile  per-launch table of one training step (CUDA events, eager)
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from cae_tools_b200.engine.convae import ConvAEEngine  # noqa: E402
from cae_tools_b200.models.decoder import Decoder  # noqa: E402
from cae_tools_b200.models.encoder import Encoder  # noqa: E402
from cae_tools_b200.models.model_sizer import create_model_spec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--n-batches", type=int, default=2)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--out-channels", type=int, default=4)
    a = ap.parse_args()
    torch.manual_seed(0)
    spec = create_model_spec(input_size=(64, 64), input_channels=4, output_size=(1024, 1024), output_channels=a.out_channels)
    enc, dec = Encoder(spec.get_input_layers(), 4, 16), Decoder(spec.get_output_layers(), 4, 16)
    B = a.batch
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(1000)
    X = torch.rand(a.n_batches * B, 4, 64, 64, device=dev, generator=gen)
    Y = torch.rand(a.n_batches * B, a.out_channels, 1024, 1024, device=dev, generator=gen)
    if a.check:
        from oracle.torch_port import OracleModel
        oracle = OracleModel(enc.state_dict(), dec.state_dict(), spec.save(), zero_dead_bias_grads=True)
    eng = ConvAEEngine(enc, dec, lr=1e-3, weight_decay=1e-5, device=dev)
    data = eng.bind(X, Y, B)
    if a.check:
        for s in range(2):
            got = float(eng.train_epoch(eng.bind(X[:B], Y[:B], B)).cpu()[0])
            want = float(oracle.train_step(X[:B].cpu(), Y[:B].cpu()))
            print(f"step {s}: loss {got:.7f} oracle {want:.7f} rel {abs(got - want) / abs(want):.2e}", flush=True)
    prog = eng.program("train", data, B) if hasattr(eng, "program") else eng._program("train", data, B)
    for _ in range(2):
        prog.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        prog.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"conv4 batch {B}: {ms:.3f} ms/step, {B / ms * 1e3:.0f} samples/s, launches {prog.n_launches}, "
          f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    if a.profile:
        table = prog.profile(reps=2)
        tot = sum(t for _, t in table)
        for name, t in table:
            print(f"  {name:30s} {t * 1e3:10.1f} us {100 * t / tot:5.1f}%")
        print(f"  sum {tot:.3f} ms")


if __name__ == "__main__":
    main()
