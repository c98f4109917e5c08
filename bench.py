#!/usr/bin/env python
"""Benchmark of the cae_tools hot path on B200 (BASELINE.json metric: train samples/s and apply images/s,
16x16 -> 256x256 CAE).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one optimiser step (forward + loss + backward + Adam/AdamW) on one batch of synthetic data of the named
shape.  Default workload: BASELINE.json configs[1] - method=unet, 16x16 -> 256x256 with skip connections, batch 64,
fp32 (the shipped layer spec cae_tools_b200/specs/unet_16x16_256x256.json; loss = masked MSE + Pearson term, AdamW).
`--method conv` runs the ConvAEModel geometry of configs[0] at the same batch instead; the default line also carries a
short conv run under "conv".  N>1 is launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N` (if it is
not, this script re-executes itself that way): one rank per GPU, batch-sharded data parallel, the flat
gradient arena all-reduced by NCCL inside the captured step, weak scaling (per-GPU batch fixed).

Rank 0 prints ONE JSON line.  `value` is device-timed with inputs resident in HBM; `e2e` is the same metric with
every step's inputs copied from pinned host memory and its loss read back inside the timed region.
`--impl reference` times the CPU oracle port of the reference's PyTorch path (the reference itself is
pure Python over PyTorch; see oracle/) on the host cores for the same config.
"""

import argparse
import json
import os
import socket
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 64                 # per-GPU batch (BASELINE config 2 names batch 64)
IN_SHAPE = (1, 16, 16)
OUT_SHAPE = (1, 256, 256)
LATENT, FC = 4, 16         # train_cae defaults (--latent-size 4 --fc-size 16)
UNET_SPEC = os.path.join(ROOT, "cae_tools_b200", "specs", "unet_16x16_256x256.json")
N_BATCHES = 64             # device-resident batches that the steps cycle through (1.1 GB > 126 MB L2)
APPLY_BATCH = 4096         # micro-batch >= 4096 for the apply sweep
APPLY_BATCHES = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: 64; conv4: 128)")
    ap.add_argument("--method", default="unet", choices=["unet", "conv", "var", "conv4"],
                    help="conv4 = BASELINE configs[3]: synthetic 4x64x64 -> 4x1024x1024 ConvAE, batch 128")
    ap.add_argument("--n-batches", type=int, default=None,
                    help="device-resident batches the steps cycle through (small values only for runs under ncu, whose "
                         "kernel replay saves / restores all device memory)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-ops", action="store_true", help="print the per-kernel time table to stderr")
    ap.add_argument("--train-only", action="store_true", help="skip the e2e / apply legs (short runs under ncu)")
    ap.add_argument("--apply-sweep", type=int, default=0, metavar="N",
                    help="BASELINE configs[4]: apply() inference sweep over N images (16x16 -> 256x256 unet), sharded over the "
                         "ranks with no collective; prints one apply_images_per_sec line instead of the training line")
    ap.add_argument("--no-api-leg", action="store_true", help="skip the UNET.train / UNET.apply model-class leg")
    a = ap.parse_args()
    set_workload(a.method)
    if a.batch is None:
        a.batch = 128 if a.method == "conv4" else BATCH
    if a.n_batches is None:
        a.n_batches = 2 if a.method == "conv4" else N_BATCHES      # conv4: one batch of targets alone is 2.1 GB
    return a


def set_workload(method):
    """shapes of the synthetic data per method (module globals read by every leg)"""
    global IN_SHAPE, OUT_SHAPE, APPLY_BATCH, APPLY_BATCHES
    if method == "conv4":
        IN_SHAPE, OUT_SHAPE = (4, 64, 64), (4, 1024, 1024)
        APPLY_BATCH, APPLY_BATCHES = 128, 2


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# ----------------------------------------------------------------------------------------------------
# geometry / algorithmic bytes (SURVEY section 8(d) convention; stated in DESIGN.md)
# ----------------------------------------------------------------------------------------------------
def layer_table(spec):
    enc = [(l.get_input_dimensions(), l.get_output_dimensions()) for l in spec.get_input_layers()]
    dec = [(l.get_input_dimensions(), l.get_output_dimensions()) for l in spec.get_output_layers()]
    return enc, dec


def numel(d):
    return d[0] * d[1] * d[2]


def bytes_per_sample(spec, fc, latent):
    """B_train = 2*in + 5*inter + 5*out ; B_apply = in + 2*inter + out   (fp32)"""
    enc, dec = layer_table(spec)
    inp = numel(enc[0][0]) * 4
    out = numel(dec[-1][1]) * 4
    inter = sum(numel(o) for _, o in enc) + sum(numel(o) for _, o in dec[:-1])
    inter += fc + latent + fc + numel(dec[0][0])
    inter *= 4
    return 2 * inp + 5 * inter + 5 * out, inp + 2 * inter + out


def op_bytes(name, spec, B):
    """algorithmic (compulsory) HBM bytes of one launch of a conv-family op: tensors read once + written once"""
    enc, dec = layer_table(spec)
    nd = len(dec)
    f = 4 * B
    if name.startswith("fwd.head"):            # fused last layer: input + (target | yhat)
        return f * (numel(dec[-1][0]) + numel(dec[-1][1]))
    if name.startswith("bwd.head"):            # input, target, mask source (act) of the input gradient, its write
        return f * (3 * numel(dec[-1][0]) + numel(dec[-1][1]))
    if name.startswith("bwd.") and not name.endswith((".wgrad", ".dgrad")):
        return None
    if ".tc" in name:
        return None                          # tensor-core layers are compute-bound: see the "tensor" roofline entry
    if name.startswith("fwd.convT"):
        j = int(name[len("fwd.convT"):].split("+")[0])
        b = f * (numel(dec[j][0]) + numel(dec[j][1]))
        if "mse" in name:
            b += f * numel(dec[j][1])          # target
        return b
    if name.startswith("fwd.conv"):
        i = int(name[len("fwd.conv"):])
        return f * (numel(enc[i][0]) + numel(enc[i][1]))
    if name.startswith("bwd.convT"):
        j = int(name[len("bwd.convT"):].split(".")[0])
        dy = numel(dec[j][1]) * (1 if j == nd - 1 else 2)      # dz (+ y for the BatchNorm-backward affine)
        if name.endswith("wgrad"):
            return f * (numel(dec[j][0]) + dy)
        return f * (dy + (2 * numel(dec[j][0]) if j > 0 else numel(dec[j][0])))   # + mask read + dz write
    if name.startswith("bwd.conv"):
        i = int(name[len("bwd.conv"):].split(".")[0])
        dy = 2 * numel(enc[i][1])
        if name.endswith("wgrad"):
            return f * (numel(enc[i][0]) + dy)
        return f * (dy + 2 * numel(enc[i][0]))
    return None


def tc_flops(name, spec, B):
    """fp32-equivalent FLOPs (2 * MACs) of one tensor-core launch group of decoder layer j"""
    if ".tc" not in name or "im2col" in name:
        return None
    enc, dec = layer_table(spec)
    j = int(name.split("convT")[1].split(".")[0])
    lay = spec.get_output_layers()[j]
    k = lay.get_kernel_size()
    kh, kw = (k if isinstance(k, (tuple, list)) else (k, k))
    (ci, hi, wi), (co, _, _) = dec[j]
    return 2.0 * B * hi * wi * ci * co * kh * kw


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference's PyTorch path on the host cores
# ----------------------------------------------------------------------------------------------------
def build_modules(method):
    """(spec, encoder, decoder) of the benchmark workload; identical construction on every rank / leg"""
    import torch
    from cae_tools_b200.models.model_sizer import ModelSpec, create_model_spec
    torch.manual_seed(0)
    if method == "unet":
        from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
        spec = ModelSpec()
        with open(UNET_SPEC) as f:
            spec.load(json.load(f))
        return spec, UNetEncoder(spec.get_input_layers(), LATENT, FC, 0.0), UNetDecoder(spec.get_output_layers(), LATENT, FC, 0.0)
    from cae_tools_b200.models.decoder import Decoder
    from cae_tools_b200.models.encoder import Encoder
    if method == "conv4":
        spec = create_model_spec(input_size=(64, 64), input_channels=4, output_size=(1024, 1024), output_channels=4)
        return spec, Encoder(spec.get_input_layers(), LATENT, FC), Decoder(spec.get_output_layers(), LATENT, FC)
    spec = create_model_spec(input_size=(16, 16), input_channels=1, output_size=(256, 256), output_channels=1)
    if method == "var":
        from cae_tools_b200.models.var_encoder import VarEncoder
        return spec, VarEncoder(spec.get_input_layers(), LATENT, FC), Decoder(spec.get_output_layers(), LATENT, FC)
    return spec, Encoder(spec.get_input_layers(), LATENT, FC), Decoder(spec.get_output_layers(), LATENT, FC)


def cpu_train_rate(method, batch, steps, warmup):
    import torch
    from oracle.torch_port import OracleModel, OracleUNet, OracleVarModel
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    spec, enc, dec = build_modules(method)
    ishape, oshape = ((4, 64, 64), (4, 1024, 1024)) if method == "conv4" else ((1, 16, 16), (1, 256, 256))
    x, y = torch.rand(batch, *ishape), torch.rand(batch, *oshape)
    if method == "unet":
        m = OracleUNet(enc.state_dict(), dec.state_dict(), spec.save(), lambda_pearson=1.0)
        ones = torch.ones_like(y)
        step = lambda: m.train_step(x, y, ones)
    elif method == "var":
        m = OracleVarModel(enc.state_dict(), dec.state_dict(), spec.save(), lambda_mse=1.0, lambda_kl=1.0)
        step = lambda: m.train_step(x, y, torch.randn(batch, LATENT))
    else:
        m = OracleModel(enc.state_dict(), dec.state_dict(), spec.save())
        step = lambda: m.train_step(x, y)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    ta = time.perf_counter()
    reps = max(1, steps // 4)
    for _ in range(reps):
        m.score(x)
    apply_rate = batch * reps / (time.perf_counter() - ta)
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads(), apply_rate


def run_reference(args):
    """--impl reference: rank 0 only; the CPU restatement of the reference path with all host threads"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 60))      # bounded sample: ~0.15 s per step on 8 cores
    cpu_batch = args.batch
    if args.method == "conv4":               # 3.6 GFLOP per sample and step: a few batch-8 steps are the bounded sample
        steps, cpu_batch = min(steps, 3), min(args.batch, 8)
    rate, ms, cores, apply_rate = cpu_train_rate(args.method, cpu_batch, steps, max(1, min(args.warmup, 3 if args.method != "conv4" else 1)))
    line = {
        "impl": "reference", "metric": "train_samples_per_sec", "value": rate, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.method, args.batch, 1),
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} optimiser steps at batch {cpu_batch} (oracle/torch_port.py, torch CPU)"},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "apply": {"value": apply_rate, "unit": "images/s"},
    }
    print(json.dumps(line), flush=True)


def workload_config(method, batch, world):
    if method == "unet":
        w = ("UNET method=unet 1x16x16->1x256x256 with skip connections + channel attention (BASELINE configs[1]): shipped "
             "spec enc 8x8x8/16x4x4/32x2x2 (k3 s2 p1), dec k4 s2 p1 x2 + k32 s32 head, latent 4, fc 16, dropout 0, "
             "loss masked-MSE + 1.0*(1 - Pearson), AdamW, fp32")
    elif method == "var":
        w = ("VarAEModel method=var 1x16x16->1x256x256 (BASELINE configs[2] geometry; the reference ships no implementation - "
             "parity unpinned), latent 4, fc 16, k3 s2, loss MSE + KL, reparameterisation noise drawn on the device, Adam, fp32")
    elif method == "conv4":
        w = ("ConvAEModel method=conv synthetic 4x64x64->4x1024x1024 (BASELINE configs[3]): spec from create_model_spec "
             "(4 encoder convs, 8 transposed convs 1024->512->...->8->4), latent 4, fc 16, k3 s2, MSE, Adam, fp32; decoder "
             "layers with >= 64 input channels run as tcgen05 3xTF32 GEMMs (tc_conv.cu), the rest on the SIMT kernels")
    else:
        w = "ConvAEModel method=conv 1x16x16->1x256x256, latent 4, fc 16, k3 s2 (BASELINE configs[0] geometry), MSE, Adam, fp32"
    nb = 2 if method == "conv4" else N_BATCHES
    return {"workload": w, "per_gpu_batch": batch, "global_batch": batch * world, "parallelism": f"dp{world}",
            "l2": f"steps cycle through {nb} device-resident batches "
                  f"({nb * batch * (numel(IN_SHAPE) + numel(OUT_SHAPE)) * 4 / 1e6:.0f} MB > 126 MB L2); "
                  "no explicit flush"}


# ----------------------------------------------------------------------------------------------------
# B200 leg
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from cae_tools_b200.engine.convae import ConvAEEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"        # keep stdout to the one JSON line (NCCL prints its version there)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = None
    if world > 1:
        from cae_tools_b200.engine.dp import bind_to_local_numa
        numa_node = bind_to_local_numa(local)     # pinned staging buffers of this rank on the GPU's own NUMA node
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)

    method = args.method
    spec, enc, dec = build_modules(method)          # identical initial weights on every rank (seed 0)
    hook = (lambda g: dist.all_reduce(g)) if world > 1 else None
    hook_async = (lambda g: dist.all_reduce(g, async_op=True)) if world > 1 else None
    dpc = None
    if world > 1:
        from cae_tools_b200.engine.dp import DPContext
        dpc = DPContext()          # small arenas: all-reduce fused into the optimiser launch (csrc/dp_fused.cu)
    if method == "unet":
        from cae_tools_b200.engine.unet import UNetEngine
        eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5, device=dev,
                         grad_hook=hook, grad_hook_async=hook_async, grad_scale=1.0 / world, dp=dpc)
    elif method == "var":
        from cae_tools_b200.engine.varae import VarAEEngine
        eng = VarAEEngine(enc, dec, lambda_mse=1.0, lambda_kl=1.0, lr=1e-3, weight_decay=1e-5, device=dev,
                          grad_hook=hook, grad_scale=1.0 / world, dp=dpc)
    else:
        eng = ConvAEEngine(enc, dec, lr=1e-3, weight_decay=1e-5, device=dev, grad_hook=hook, grad_hook_async=hook_async,
                           grad_scale=1.0 / world, dp=dpc)

    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    NB = args.n_batches
    X = torch.rand(NB * B, *IN_SHAPE, device=dev, generator=gen)
    Y = torch.rand(NB * B, *OUT_SHAPE, device=dev, generator=gen)
    data = eng.bind(X, Y, B)
    prog = eng.program("train", data, B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    # ---- train, inputs resident
    for _ in range(W):
        prog.run()
    clocks = ClockSampler(local)
    clocks.start()
    ms = timed(prog.run, K)
    value = world * B * K / (ms / 1e3)
    final_loss = float(data.losses.mean().item())

    if args.train_only:
        clocks.stop()
        if rank == 0:
            print(json.dumps({"metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
                              "steps": K, "warmup": W, "ms_per_step": ms / K, "launches_per_step": prog.n_launches,
                              "note": "--train-only run (profiling aid, not a bench line)"}), flush=True)
        finish(world)
        return

    # ---- train end to end through the engine's host-fed entry point: every step's inputs come from pinned host
    # memory (double-buffered: the copy of batch i+1 overlaps step i) and every step's loss goes back to the host
    nh = 2 if method == "conv4" else 8
    xh = torch.rand(nh, B, *IN_SHAPE).pin_memory()
    yh = torch.rand(nh, B, *OUT_SHAPE).pin_memory()
    state = {"i": 0}

    def host_batches(count):
        for i in range(count):
            yield xh[i % nh], yh[i % nh]

    eng.train_stream(host_batches(W + 2), B)
    Ke = max(20, K // 2)
    ms_e2e = timed(lambda: eng.train_stream(host_batches(Ke), B), 1)
    e2e_value = world * B * Ke / (ms_e2e / 1e3)
    h2d = B * (numel(IN_SHAPE) + numel(OUT_SHAPE)) * 4

    # ---- apply (eval forward), inputs resident / end to end
    AB = APPLY_BATCH
    XA = torch.rand(APPLY_BATCHES * AB, *IN_SHAPE, device=dev, generator=gen)
    adata = eng.bind(XA, None, AB)
    eng._eval_prepare_op()()
    aprog = eng.program("score", adata, AB)
    for _ in range(W):
        aprog.run()
    Ka = max(10, K // 4)
    ms_apply = timed(aprog.run, Ka)
    apply_value = world * AB * Ka / (ms_apply / 1e3)
    xah = torch.rand(4, AB, *IN_SHAPE).pin_memory()
    yah = torch.empty(AB, *OUT_SHAPE).pin_memory()
    xas = torch.empty(AB, *IN_SHAPE, device=dev)
    sadata = eng.bind(xas, None, AB)
    saprog = eng.program("score", sadata, AB)
    yout = eng.output_buffer(eng._act_buffers(AB))

    def apply_e2e():
        sadata.X.copy_(xah[state["i"] % 4], non_blocking=True)
        state["i"] += 1
        saprog.run()
        yah.copy_(yout, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        apply_e2e()
    Kae = max(5, K // 10)
    ms_apply_e2e = timed(apply_e2e, Kae)
    apply_e2e_value = world * AB * Kae / (ms_apply_e2e / 1e3)

    clocks.stop()       # sampled over every timed region above (train, host-fed train, apply, host-fed apply)

    # ---- per-kernel table: what every launch adds to the critical path INSIDE the captured step (prefix graphs, warm;
    # engine/convae.py:_Program.timeline) - the step is bound by its longest chain, not by an eager per-op sum.
    # Programs with an eager collective between two graphs (CAE_CAPTURE_ALLREDUCE=0) fall back to eager per-op events.
    if hasattr(prog, "timeline") and world == 1:
        table = [(n, us / 1e3) for n, us in prog.timeline(reps=30 if method == "conv4" else 100)]
        table_kind = "in-graph prefix timeline"
    else:
        table = prog.profile(reps=5)
        table_kind = "eager CUDA events per launch"
    step_ms = ms / K
    if args.profile_ops and rank == 0:
        for name, t in sorted(table, key=lambda r: -r[1]):
            ob = op_bytes(name, spec, B)
            gbs = f"{ob / t / 1e6:8.0f} GB/s" if ob and t > 0 else ""
            print(f"  {name:28s} {t * 1e3:8.1f} us {100 * t / step_ms:5.1f}% of the step  {gbs}", file=sys.stderr)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    conv_rows = [(n, t, op_bytes(n, spec, B)) for n, t in table if op_bytes(n, spec, B) and t > 0]
    # dominant kernel = the launch that carries the most algorithmic bytes of the step (what bounds the step at the HBM
    # roofline); "longest" = the launch that adds the most time to the step, whatever it moves
    top = max(conv_rows, key=lambda r: (r[2], r[1]))
    achieved = top[2] / (top[1] / 1e3) / 1e9
    b_train, b_apply = bytes_per_sample(spec, FC, LATENT)
    longest = max(table, key=lambda r: r[1])
    roofline = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src, "timing": table_kind,
                "kernel_us": top[1] * 1e3, "kernel_share_of_step": top[1] / step_ms,
                "algorithmic_bytes_per_launch": top[2],
                "longest": {"kernel": longest[0], "us": longest[1] * 1e3, "share_of_step": longest[1] / step_ms,
                            "algorithmic_bytes": op_bytes(longest[0], spec, B)},
                "launches": [{"kernel": n, "us": t * 1e3, "share_of_step": t / step_ms} for n, t in table] if len(table) <= 12 else None,
                "kernels": [{"kernel": n, "us": t * 1e3, "algorithmic_bytes": ob, "achieved_gbs": ob / (t / 1e3) / 1e9}
                            for n, t, ob in sorted(conv_rows, key=lambda r: -r[1])[:5]],
                "step": {"bytes_per_sample": b_train, "achieved": b_train * B / (ms / K / 1e3) / 1e9,
                         "frac": b_train * B / (ms / K / 1e3) / 1e9 / hbm_peak},
                "apply_step": {"bytes_per_image": b_apply, "achieved": b_apply * AB / (ms_apply / Ka / 1e3) / 1e9,
                               "frac": b_apply * AB / (ms_apply / Ka / 1e3) / 1e9 / hbm_peak}}

    # tensor-core layers (conv4): fp32-equivalent FLOPs of the three GEMMs of every tcgen05 layer against the TF32 peak
    # (= half the measured bf16 GEMM throughput); 3xTF32 issues three MMAs per product, so the pipe runs at 3x this
    tc_rows = [(n, t, tc_flops(n, spec, B)) for n, t in table if tc_flops(n, spec, B)]
    if tc_rows:
        tf32_peak = 0.5 * float(peaks.get("bf16_tflops_sustained", 1404.4))
        fl, tt = sum(r[2] for r in tc_rows), sum(r[1] for r in tc_rows)
        roofline["tensor"] = {"bound": "tensor", "unit": "TFLOP/s", "peak": tf32_peak,
                              "peak_source": "0.5 x bf16_tflops_sustained (MEASURED_PEAKS.json)" if "bf16_tflops_sustained" in peaks else "fallback",
                              "achieved": fl / (tt / 1e3) / 1e12, "tf32_issue_rate": 3 * fl / (tt / 1e3) / 1e12,
                              "frac": fl / (tt / 1e3) / 1e12 / tf32_peak, "frac_of_issue_rate": 3 * fl / (tt / 1e3) / 1e12 / tf32_peak,
                              "launch_groups_us": {n: t * 1e3 for n, t, _ in tc_rows},
                              "note": "time includes the pack / im2col / col2im passes of each layer, not the GEMM alone"}

    # DRAM traffic of the dominant kernel from the committed ncu capture (profiles/ncu_traffic.json: op name, batch and the
    # hash of the kernel's source file at capture time; a stale hash is reported instead of a stale number)
    try:
        import hashlib
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ncu = json.load(f)
        ent = ncu.get(f"{method}:{top[0]}:{B}")
        if ent:
            with open(os.path.join(ROOT, ent["source"]), "rb") as f:
                cur = hashlib.sha256(f.read()).hexdigest()[:16]
            if cur == ent["source_sha256_16"]:
                roofline["traffic"] = ent["dram_bytes"]
                roofline["traffic_source"] = ent["capture"]
            else:
                roofline["traffic_source"] = f"stale: {ent['source']} changed since {ent['capture']}"
    except Exception:
        pass

    # ---- the ConvAEModel geometry of BASELINE configs[0] at the same batch, short run (secondary numbers)
    also = None
    if method == "unet" and world == 1:
        spec_c, enc_c, dec_c = build_modules("conv")
        eng_c = ConvAEEngine(enc_c, dec_c, lr=1e-3, weight_decay=1e-5, device=dev)
        data_c = eng_c.bind(X[:16 * B], Y[:16 * B], B)
        prog_c = eng_c.program("train", data_c, B)
        for _ in range(W):
            prog_c.run()
        kc = max(20, K // 4)
        ms_c = timed(prog_c.run, kc)
        adata_c = eng_c.bind(XA, None, AB)
        eng_c._eval_prepare_op()()
        aprog_c = eng_c.program("score", adata_c, AB)
        for _ in range(3):
            aprog_c.run()
        ms_ca = timed(aprog_c.run, 10)
        bt_c, ba_c = bytes_per_sample(spec_c, FC, LATENT)
        # ... and at configs[0]'s own batch size (10)
        data_c10 = eng_c.bind(X[:64 * 10], Y[:64 * 10], 10)
        prog_c10 = eng_c.program("train", data_c10, 10)
        for _ in range(W):
            prog_c10.run()
        ms_c10 = timed(prog_c10.run, kc)
        also = {"workload": workload_config("conv", B, 1)["workload"],
                "train_samples_per_sec": B * kc / (ms_c / 1e3), "ms_per_step": ms_c / kc,
                "batch10": {"train_samples_per_sec": 10 * kc / (ms_c10 / 1e3), "ms_per_step": ms_c10 / kc,
                            "step_frac_of_hbm_roofline": bt_c * 10 / (ms_c10 / kc / 1e3) / 1e9 / hbm_peak},
                "apply_images_per_sec": AB * 10 / (ms_ca / 1e3), "launches_per_step": prog_c.n_launches,
                "step_frac_of_hbm_roofline": bt_c * B / (ms_c / kc / 1e3) / 1e9 / hbm_peak,
                "apply_frac_of_hbm_roofline": ba_c * AB / (ms_ca / 10 / 1e3) / 1e9 / hbm_peak}

    # ---- the model classes themselves (what train_cae / apply_cae call): UNET(...).train on an in-memory data set
    # (normalisation + one H2D of everything + epochs, as the reference does) and UNET.apply (host arrays in, host
    # float64 array out).  Wall clock, everything included.
    api = None
    if method == "unet" and world == 1 and not args.no_api_leg:
        api = api_leg(B)

    # ---- data-parallel exchange: the flat gradient arena, all-reduced alone (what one step's exchange costs)
    comm = None
    if world > 1:
        g = eng.grads
        for _ in range(5):
            dist.all_reduce(g)
        ms_ar = timed(lambda: dist.all_reduce(g), 50)
        fused = getattr(eng, "_dp_peers", None) is not None
        comm = {"nccl_allreduce_us": ms_ar / 50 * 1e3, "arena_bytes": g.numel() * 4, "fused_into_optimiser": fused,
                "in_step_graph": fused or bool(getattr(eng, "capture_allreduce", False)),
                "note": ("the gradient all-reduce runs inside the optimiser launch (peer reads over NVLink, cae_adam_allreduce); "
                         "nccl_allreduce_us is what a standalone NCCL all-reduce of the same arena costs") if fused else
                        "one NCCL all-reduce of the flat fp32 gradient arena per optimiser step",
                "rank0_numa_node": numa_node}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        csteps, cb = (3, min(B, 8)) if method == "conv4" else (40, B)
        rate, cms, cores, arate = cpu_train_rate(method, cb, csteps, 1 if method == "conv4" else 3)
        cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "apply_images_per_sec": arate,
               "sample": f"{csteps} optimiser steps at batch {cb} (+ eval batches) of the same workload, oracle port on torch CPU"}

    if rank == 0:
        line = {
            "metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(method, B, world),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / Ke},
            "gpu_launches": prog.n_launches * K,
            "launches_per_step": prog.n_launches,
            "apply": {"value": apply_value, "unit": "images/s", "batch": AB, "ms_per_batch": ms_apply / Ka,
                      "e2e": {"value": apply_e2e_value, "unit": "images/s",
                              "h2d_bytes_per_step": AB * numel(IN_SHAPE) * 4,
                              "d2h_bytes_per_step": AB * numel(OUT_SHAPE) * 4}},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "api": api,
            "comm": comm,
            "conv": also,
            "final_loss": final_loss,
        }
        print(json.dumps(line), flush=True)
    finish(world)


def api_leg(B):
    """UNET model class end to end on host data: train (E epochs over n cases) and apply; wall-clock rates"""
    import numpy as np
    import torch
    from cae_tools_b200.models.model_sizer import ModelSpec
    from cae_tools_b200.models.unet import UNET
    from cae_tools_b200.utils import xr_lite
    n, E = 8 * B, 6
    rng = np.random.RandomState(0)
    ds, dt = xr_lite.Dataset(), xr_lite.Dataset()
    for d, m in ((ds, n), (dt, B)):
        d["lowres"] = xr_lite.DataArray(rng.rand(m, *IN_SHAPE).astype(np.float32), dims=("n", "chan", "y1", "x1"))
        d["hires"] = xr_lite.DataArray(rng.rand(m, *OUT_SHAPE).astype(np.float32), dims=("n", "chan", "y2", "x2"))
    spec = ModelSpec()
    with open(UNET_SPEC) as f:
        spec.load(json.load(f))
    torch.manual_seed(0)
    mdl = UNET(batch_size=B, nr_epochs=E, test_interval=E, encoded_dim_size=LATENT, fc_size=FC, dropout_rate=0.1)
    mdl.verbose = False
    mdl.spec = spec
    t0 = time.perf_counter()
    mdl.train(["lowres"], "hires", ds, dt)
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    t0 = time.perf_counter()
    mdl.apply(ds, ["lowres"], "est")
    t_apply = time.perf_counter() - t0
    return {"call": "UNET(batch_size=%d, dropout_rate=0.1).train(...) / .apply(...) on in-memory data sets" % B,
            "train_samples_per_sec": n * E / t_train, "train_seconds": t_train, "cases": n, "epochs": E,
            "apply_images_per_sec": n / t_apply, "apply_seconds": t_apply,
            "note": "wall clock incl. min/max normalisation on the host, one H2D of all batches, the two post-training "
                    "evaluate() passes (train) and the float64 de-normalised host array (apply)"}


def run_apply_sweep(args):
    """BASELINE configs[4]: N images sharded contiguously over the ranks (engine/dp.py:shard_bounds), no collective; every
    rank streams its shard through the eval-mode program in micro-batches of 4096 with inputs device-resident; outputs stay
    on the device (the engine's output buffer is the ring) - and, second number, are copied to pinned host memory."""
    import torch
    import torch.distributed as dist
    from cae_tools_b200.engine.dp import shard_bounds
    from cae_tools_b200.engine.unet import UNetEngine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec, enc, dec = build_modules("unet")
    eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, device=dev)
    lo, hi = shard_bounds(args.apply_sweep, rank, world)
    n = hi - lo
    gen = torch.Generator(device=dev).manual_seed(2000 + rank)
    X = torch.rand(n, *IN_SHAPE, device=dev, generator=gen)
    AB = APPLY_BATCH
    data = eng.bind(X, None, AB)
    stage = [torch.empty(AB, *OUT_SHAPE).pin_memory() for _ in range(2)]
    copy = torch.cuda.Stream(device=dev)

    def sweep(d2h):
        main = torch.cuda.current_stream(dev)
        k = [0]

        def sink(i, yh):
            if d2h:
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(copy):
                    copy.wait_event(ev)
                    stage[i & 1][:yh.shape[0]].copy_(yh, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy)
                main.wait_event(done)
            k[0] += yh.shape[0]
        eng.score_batches(data, sink)
        assert k[0] == n

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sweep(False)                                   # warm-up: graphs captured, eval BatchNorm folded
    clocks = ClockSampler(local)
    clocks.start()
    ms = min(timed(lambda: sweep(False)) for _ in range(3))
    ms_d2h = timed(lambda: sweep(True)) if args.apply_sweep <= 200000 * world else None
    clocks.stop()
    _, b_apply = bytes_per_sample(spec, FC, LATENT)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    if rank == 0:
        value = args.apply_sweep / (ms / 1e3)
        line = {"metric": "apply_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": 3, "warmup": 1,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"apply() sweep over N={args.apply_sweep} images 1x16x16->1x256x256 unet (BASELINE configs[4]), "
                                       f"sharded contiguously over {world} GPU(s), micro-batch {AB}, no collective; inputs device-resident, "
                                       "outputs left in the engine's device buffer", "parallelism": f"shard{world}",
                           "l2": f"{n * 1024 / 1e6:.0f} MB of inputs and 1 GB of outputs per micro-batch per GPU: nothing is reused from L2"},
                "clocks": clocks.summary(),
                "e2e": None if ms_d2h is None else {"value": args.apply_sweep / (ms_d2h / 1e3), "unit": "images/s",
                                                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": n * numel(OUT_SHAPE) * 4,
                                                    "note": "outputs copied to pinned host memory (double-buffered); inputs resident"},
                "gpu_launches": eng.program("score", data, AB).n_launches * ((n + AB - 1) // AB) * 3,
                "roofline": {"bound": "hbm", "bytes_per_image": b_apply, "achieved": b_apply * args.apply_sweep / world / (ms / 1e3) / 1e9,
                             "peak": hbm, "unit": "GB/s", "frac": b_apply * args.apply_sweep / world / (ms / 1e3) / 1e9 / hbm,
                             "traffic": None, "note": "SURVEY 8(d) convention B_apply = in + 2*inter + out per image, per GPU"}}
        print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """leave without tearing down NCCL communicators / CUDA graphs one by one (a teardown that hangs after the line was
    printed would cost the whole run): every rank reaches the barrier, flushes, exits 0"""
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.gpus > 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if args.impl == "reference":
        run_reference(args)
    elif args.apply_sweep > 0:
        run_apply_sweep(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
