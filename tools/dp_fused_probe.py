#!/usr/bin/env python
"""Cost of the data-parallel exchange alone (run under torchrun, one rank per GPU): Adam launch, Adam + peer-memory
all-reduce in one launch (cae_adam_allreduce), NCCL all-reduce of the same arena - back to back, CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cae_tools_b200.engine import dp as dpm, ops  # noqa: E402

ctx = dpm.init_from_env()
dev = torch.device("cuda", torch.cuda.current_device())
for n in (39328, 39328 * 4):
    sym = ctx.symmetric_arena(n)
    assert sym is not None, getattr(ctx, "_symm_error", "")
    grads, peers, keep = sym
    grads.normal_()
    p, m, v = (torch.zeros(n, device=dev) for _ in range(3))
    step, cur = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    epoch, t1, t2 = (torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(3))

    def timed(fn, reps=500):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3

    def graphed(fn, k=50):
        g = torch.cuda.CUDAGraph()
        fn()
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(k):
                    fn()
        return lambda: g.replay(), k

    adam = lambda: ops.adam_advance(p, grads, m, v, n, 1e-3, 0.9, 0.999, 1e-8, 1e-5, True, 1.0, step, cur, 4, t1)
    def fused():
        ops.dp_wait_done(peers, epoch)
        ops.adam_allreduce(p, peers, m, v, n, 1e-3, 0.9, 0.999, 1e-8, 1e-5, True, 0.5, step, cur, 4, epoch, t2)
    nccl = lambda: dist.all_reduce(grads)
    res = {}
    for name, fn in (("adam", adam), ("adam+allreduce fused", fused)):
        r, k = graphed(fn)
        res[name] = timed(r, 40) / k
    res["nccl all_reduce (eager)"] = timed(nccl, 300)
    if ctx.rank == 0:
        print(f"arena {n * 4 / 1024:.0f} KB, world {ctx.world}: " + ", ".join(f"{k} {v:.1f} us" for k, v in res.items()), flush=True)
    grads.zero_()
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
