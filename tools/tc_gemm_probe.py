#!/usr/bin/env python
"""tcgen05 GEMM probe: error of every operand-major combination + throughput at the config-4 fat-layer GEMM shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cae_tools_b200.engine import ops  # noqa: E402


def operand(rows, K, mn, gen):
    pad = lambda v: (v + 3) // 4 * 4
    if mn:
        st = torch.randn(K, pad(rows), device="cuda", generator=gen)
        return st, st.shape[1], st[:, :rows].t().double()
    st = torch.randn(rows, pad(K), device="cuda", generator=gen)
    return st, st.shape[1], st[:, :K].double()


def split(x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    ops.tc_split(x, hi, lo)
    return hi, lo


def check():
    gen = torch.Generator(device="cuda").manual_seed(1)
    for (M, N, K, splits, tn) in [(128, 128, 32, 1, 128), (128, 128, 64, 1, 128), (200, 136, 100, 1, 128), (384, 520, 1000, 3, 128),
                                  (130, 300, 264, 2, 256), (256, 512, 4608, 1, 256), (256, 256, 4608, 1, 128), (256, 512, 8192, 4, 256)]:
        for a_mn, b_mn in [(0, 0), (1, 1), (0, 1), (1, 0)]:
            A, lda, Ad = operand(M, K, a_mn, gen)
            B, ldb, Bd = operand(N, K, b_mn, gen)
            ah, al = split(A)
            bh, bl = split(B)
            ldc = (N + 3) // 4 * 4
            Cb = torch.zeros(splits, M, ldc, device="cuda")
            ops.tc_gemm(M, N, K, ah, al, lda, a_mn, bh, bl, ldb, b_mn, Cb, ldc, splits=splits, split_stride=M * ldc, tile_n=tn)
            torch.cuda.synchronize()
            got = Cb[:, :, :N].double().sum(0)
            want = Ad @ Bd.t()
            err = float((got - want).abs().max() / want.abs().max())
            # the same product as a plain fp32 GEMM (cuBLAS, TF32 off): the accuracy the tensor-core route has to match
            tf = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            plain = (Ad.float() @ Bd.float().t()).double()
            torch.backends.cuda.matmul.allow_tf32 = tf
            e32 = float((plain - want).abs().max() / want.abs().max())
            rms = float((got - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt())
            rms32 = float((plain - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt())
            print(f"M{M} N{N} K{K} splits{splits} tile_n{tn} a_mn{a_mn} b_mn{b_mn}: err {err:.2e} (fp32 GEMM {e32:.2e})  "
                  f"rms {rms:.2e} (fp32 GEMM {rms32:.2e})", flush=True)


def bench():
    gen = torch.Generator(device="cuda").manual_seed(2)
    # forward GEMMs of the config-4 fat layers at batch 128: M = B*Hin*Win, N = 9*Cout, K = Cin; + wgrad (MN-major) shapes
    shapes = [("fwd T0", 1152, 4608, 1024, 0, 0, 1), ("fwd T1", 6272, 2304, 512, 0, 0, 1), ("fwd T2", 28800, 1152, 256, 0, 0, 1),
              ("fwd T3", 123008, 576, 128, 0, 0, 1), ("dgrad T3", 123008, 128, 576, 0, 0, 1),
              ("wgrad T3", 128, 576, 123008, 1, 1, 64), ("wgrad T0", 1024, 4608, 1152, 1, 1, 1)]
    for name, M, N, K, a_mn, b_mn, splits in shapes:
        for tn in (128, 256):
            for three in (True, False):
                A, lda, _ = operand(M, K, a_mn, gen)
                B, ldb, _ = operand(N, K, b_mn, gen)
                al, bl = (torch.empty_like(A), torch.empty_like(B)) if three else (None, None)
                ldc = (N + 3) // 4 * 4
                Cb = torch.empty(splits, M, ldc, device="cuda")
                run = lambda: ops.tc_gemm(M, N, K, A, al, lda, a_mn, B, bl, ldb, b_mn, Cb, ldc, splits=splits,
                                          split_stride=M * ldc, tile_n=tn)
                for _ in range(3):
                    run()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                fl = 2.0 * M * N * K
                print(f"{name:9s} M{M} N{N} K{K} tile_n{tn} {'3xTF32' if three else '1xTF32'}: {ms * 1e3:8.1f} us  "
                      f"{fl / ms / 1e9:7.1f} TFLOP/s fp32-equivalent ({fl * (3 if three else 1) / ms / 1e9:7.1f} TF32)", flush=True)


if __name__ == "__main__":
    check()
    if len(sys.argv) > 1 and sys.argv[1] == "bench":
        bench()
