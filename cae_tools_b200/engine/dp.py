"""Batch-sharded data parallelism: one process per GPU, torch.distributed for the plumbing.

The reference has no multi-GPU path (single `device`, conv_ae_model.py:294-297); this is the
data-parallel wrapper BASELINE.json's north_star asks for.  Every optimiser step has exactly one
exchange of the flat fp32 gradient arena (NCCL over NVLink on GPUs; gloo in the CPU tests), in two buckets: the
decoder + fc gradients are SUM-all-reduced asynchronously while the encoder backward still runs, the encoder bucket
right after it; the optimiser waits for both (engine/convae.py:_BucketedProgram).  Each rank computes on its contiguous share of
every global batch; the loss epilogue divides by the GLOBAL element count (count_scale =
n_local/n_global), so the summed gradients - and the summed per-batch losses - are exactly those of
the global batch for the plain MSE.  The UNET's masked MSE divides by the mask count: there the constant is replaced by
a per-batch factor, valid pixels of the share / valid pixels of the global batch (`DPContext.mask_scales`, all-reduced
once at bind time, read by the loss kernels at the batch cursor: `mse_scale` in cae_b200.h), so the sum over ranks is the
global sum((d-t)^2 m^2) / sum(m) for any land / sea split (host-streamed batches, `train_stream`, keep the constant).  BatchNorm uses the statistics of the local share ("local BN").  `apply` shards the
samples over ranks with no collective at all.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """contiguous split of n items: the first n % world ranks get one extra"""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batches(order, batch_size, rank, world):
    """Split every batch of the (already shuffled) sample order over the ranks.
    Returns (local_order, local_batch_size, count_scale).  Batches must divide evenly so that every
    rank runs the same kernel schedule: batch_size % world == 0 and (n % batch_size) % world == 0."""
    n = len(order)
    if batch_size % world != 0:
        raise ValueError(f"batch_size {batch_size} is not a multiple of the {world} data-parallel ranks")
    tail = n % batch_size
    if tail % world != 0:
        raise ValueError(f"the last batch ({tail} samples) does not split evenly over {world} ranks")
    local = []
    for b0 in range(0, n, batch_size):
        chunk = order[b0:b0 + batch_size]
        lo, hi = shard_bounds(len(chunk), rank, world)
        local.extend(chunk[lo:hi])
    return local, batch_size // world, 1.0 / world


class DPContext:
    """rank / world bookkeeping + the two collectives the hot path uses"""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    @staticmethod
    def from_env():
        """a context if torch.distributed is initialised with more than one rank, else None"""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return DPContext()
        return None

    def allreduce_grads(self, flat):
        """the per-step exchange: in-place SUM over ranks of the flat gradient arena"""
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat

    def mask_scales(self, mask, batch_size):
        """mask [n_local, c, H, W] of this rank's shares in batch order -> float32 [n_batches]: valid pixels of the share /
        valid pixels of the global batch (one SUM all-reduce; every rank holds the same number of batches).  The factor the
        masked-MSE term of a share is multiplied with so that the shares' losses and gradients ADD UP to the global
        batch's sum((d-t)^2 m^2) / sum(m), whatever the split of valid pixels between the ranks."""
        n = int(mask.shape[0])
        per_sample = mask.sum(dim=(1, 2, 3), dtype=torch.float64)
        local = torch.stack([per_sample[lo:lo + batch_size].sum() for lo in range(0, n, batch_size)])
        total = local.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return (local / total.clamp_min(1e-300)).to(torch.float32).contiguous()

    def allreduce_grads_async(self, bucket):
        """one gradient bucket (a contiguous slice of the flat arena): SUM over ranks, asynchronous; the caller keeps
        computing (the encoder backward) and calls .wait() on the handle before the optimiser"""
        return dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def symmetric_arena(self, numel):
        """A gradient arena every rank can read directly over NVLink (torch symmetric memory: a CUDA VMM allocation whose
        handles are exchanged once, `buffer_ptrs[r]` = rank r's copy mapped into this process) + a flag array for the
        in-kernel handshake of cae_adam_allreduce.  Returns (grads, peers descriptor, keep-alive) or None when symmetric
        memory is unavailable (every rank then falls back to the NCCL all-reduce: the decision is taken collectively)."""
        import os
        ok = torch.tensor([1], dtype=torch.int32, device="cuda")
        res = None
        try:
            if os.environ.get("CAE_DP_FUSED", "1") == "0" or dist.get_backend(self.group) != "nccl" or self.world > 8:
                raise RuntimeError("fused exchange disabled / unsupported")
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.group if self.group is not None else dist.group.WORLD
            enable = getattr(symm_mem, "enable_symm_mem_for_group", None)
            if enable is not None:
                try:
                    enable(grp.group_name)
                except Exception:
                    pass
            grads = symm_mem.empty(int(numel), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
            flags = symm_mem.empty(64, dtype=torch.int32, device=grads.device)
            grads.zero_()
            flags.zero_()
            hg = symm_mem.rendezvous(grads, grp)
            hf = symm_mem.rendezvous(flags, grp)
            torch.cuda.synchronize()
            from . import ops
            peers = ops.make_dp_peers(self.world, self.rank, list(hg.buffer_ptrs), list(hf.buffer_ptrs))
            res = (grads, peers, (flags, hg, hf))
        except Exception as exc:  # noqa: BLE001 - any failure means "use NCCL"
            ok.zero_()
            self._symm_error = repr(exc)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)      # all ranks or none; also orders the flag zeroing
        if int(ok.item()) == 0:
            return None
        return res

    def reduce_losses(self, losses):
        """per-batch losses were divided by the global count on every rank: SUM gives the global batch MSE"""
        out = losses.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def broadcast_(self, tensors, src=0):
        for t in tensors:
            dist.broadcast(t, src=src, group=self.group)


def bind_to_local_numa(device_index):
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated
    (cudaHostAlloc places the pages on the node of the calling thread).  With one process per GPU and every rank's staging
    buffers on node 0, the host-fed step of 8 ranks shared one node's memory controllers (round 1: e2e scaling 0.51 at 8
    GPUs).  Best effort: returns the node, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:       # no sysfs entry / no pci ids in this torch build / not permitted
        return None


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR /
    MASTER_PORT): one process per GPU, NCCL when CUDA is present (gloo otherwise: CPU tests).  No-op outside torchrun or
    when a group already exists.  Returns the DPContext (None for a single process)."""
    import os
    if not dist.is_available():
        return None
    if not dist.is_initialized():
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1 or "RANK" not in os.environ:
            return None
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            bind_to_local_numa(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return DPContext.from_env()


def respawn_under_torchrun(gpus, argv, module):
    """`--gpus N` given to a CLI outside torchrun: re-execute `python -m torch.distributed.run --nproc-per-node N -m
    <module> <argv>` on this node (the way bench.py launches itself); never returns."""
    import os
    import socket
    import sys
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), "-m", module] + list(argv)
    os.execv(sys.executable, cmd)
