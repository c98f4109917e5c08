"""Data-parallel host logic on CPU: world_size 2, gloo.  The compute leg is the oracle port (the CUDA kernels need
a GPU); what is under test is the sharding, the count_scale convention and the two collectives of engine/dp.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from cae_tools_b200.engine.dp import shard_batches, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 10, 64):
        for world in (1, 2, 3, 8):
            parts = [shard_bounds(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batches_covers_every_batch():
    order = list(np.random.RandomState(0).permutation(100))
    world, bs = 4, 32       # 3 full batches + tail of 4
    shards = [shard_batches(order, bs, r, world) for r in range(world)]
    assert all(s[1] == 8 and s[2] == 0.25 for s in shards)
    for b in range(4):
        got = []
        for r in range(world):
            loc = shards[r][0]
            n_loc = 8 if b < 3 else 1
            start = b * 8
            got += loc[start:start + n_loc]
        assert sorted(got) == sorted(order[b * bs:(b + 1) * bs])
    try:
        shard_batches(order, 30, 0, 4)
        assert False
    except ValueError:
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from cae_tools_b200.engine.dp import DPContext
    from cae_tools_b200.models.decoder import Decoder
    from cae_tools_b200.models.encoder import Encoder
    from cae_tools_b200.models.model_sizer import create_model_spec
    from oracle.torch_port import OracleModel
    dp = DPContext.from_env()
    assert dp is not None and dp.rank == rank and dp.world == world
    torch.manual_seed(3)
    spec = create_model_spec(input_size=(16, 16), input_channels=1, output_size=(64, 64), output_channels=1)
    enc, dec = Encoder(spec.get_input_layers(), 4, 16), Decoder(spec.get_output_layers(), 4, 16)
    g = torch.Generator().manual_seed(11)
    xs, ys = torch.rand(4, 1, 16, 16, generator=g), torch.rand(4, 1, 64, 64, generator=g)
    # global batch = the local shard repeated on every rank -> local BN statistics == global ones
    x_glob, y_glob = xs.repeat(world, 1, 1, 1), ys.repeat(world, 1, 1, 1)
    order = list(range(4 * world))
    local, lb, cscale = shard_batches(order, 4 * world, rank, world)
    assert lb == 4 and cscale == 1.0 / world
    m = OracleModel(enc.state_dict(), dec.state_dict(), spec.save())
    losses = []
    for step in range(3):
        yhat = m.forward(x_glob[local], True)
        loss = F.mse_loss(yhat, y_glob[local]) * cscale        # what the kernels' count_scale does
        m.optim.zero_grad()
        loss.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in m.params])
        if step == 1:
            # the bucketed variant (engine/convae.py:_BucketedProgram): two contiguous slices of the arena, each reduced
            # asynchronously, both waited for before the optimiser
            cut = flat.numel() // 3
            handles = [dp.allreduce_grads_async(flat[cut:]), dp.allreduce_grads_async(flat[:cut])]
            for h in handles:
                h.wait()
        else:
            dp.allreduce_grads(flat)                             # the per-step exchange
        off = 0
        for p in m.params:
            p.grad.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        m.optim.step()
        losses.append(float(dp.reduce_losses(loss.detach().reshape(1))[0]))
    ref = OracleModel(enc.state_dict(), dec.state_dict(), spec.save())
    ref_losses = [float(ref.train_step(x_glob, y_glob)) for _ in range(3)]
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-5)
    from oracle.torch_port import trainable_keys
    names = [k for sd in (m.enc, m.dec) for k in trainable_keys(sd)]
    for k, a, b in zip(names, m.params, ref.params):
        dead = k.endswith(".bias") and k.split(".")[0] in ("encoder_cnn", "decoder_conv") and \
            any(f"{k.split('.')[0]}.{int(k.split('.')[1]) + 1}.running_mean" in sd for sd in (m.enc, m.dec))
        if dead:
            continue    # zero true gradient; autograd's rounding noise is amplified differently by Adam
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5), k
    # masked MSE under data parallelism (DPContext.mask_scales): with the per-batch factor CNT_share / CNT_global the shares'
    # masked-MSE terms add up to the global batch's sum((d-t)^2 m^2) / sum(m) although the ranks see very different masks
    gm = torch.Generator().manual_seed(9)
    d, t = torch.rand(2 * world, 1, 8, 8, generator=gm), torch.rand(2 * world, 1, 8, 8, generator=gm)
    keep = torch.linspace(0.9, 0.1, 2 * world).view(-1, 1, 1, 1)
    msk = (torch.rand(2 * world, 1, 8, 8, generator=gm) < keep).float()
    mine = slice(2 * rank, 2 * rank + 2)                      # two batches of one sample per rank: global batches (0,2), (1,3)
    scales = dp.mask_scales(msk[mine], 1)
    assert scales.dtype == torch.float32 and scales.shape == (2,)
    for b in range(2):
        idx = [b + 2 * r for r in range(world)]
        want = float(((d[idx] - t[idx]) ** 2 * msk[idx] ** 2).sum() / msk[idx].sum())
        i = 2 * rank + b
        part = ((d[i] - t[i]) ** 2 * msk[i] ** 2).sum() / msk[i].sum() * scales[b]
        got = float(dp.reduce_losses(part.reshape(1))[0])
        assert abs(got - want) <= 1e-6 * want, (b, got, want)
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.destroy_process_group()


def test_dp_two_ranks_gloo_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok"))
