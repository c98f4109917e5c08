"""Data-parallel training on 2 GPUs over NCCL (skipped on a single-GPU box): each rank trains on the same shard,
so local BatchNorm statistics equal the global ones and the 2-rank run must reproduce a single-GPU run on the
duplicated batch (all-reduce inside the captured step graph, count_scale = 1/world)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _models():
    from cae_tools_b200.models.decoder import Decoder
    from cae_tools_b200.models.encoder import Encoder
    from cae_tools_b200.models.model_sizer import create_model_spec
    torch.manual_seed(21)
    spec = create_model_spec(input_size=(16, 16), input_channels=1, output_size=(64, 64), output_channels=1)
    return Encoder(spec.get_input_layers(), 4, 16), Decoder(spec.get_output_layers(), 4, 16)


def _data():
    g = torch.Generator().manual_seed(22)
    return torch.rand(8, 1, 16, 16, generator=g), torch.rand(8, 1, 64, 64, generator=g)


def _worker(rank, world, port, out_dir, overlap, fused=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cae_tools_b200.engine.convae import ConvAEEngine
    from cae_tools_b200.engine.dp import DPContext
    dp = DPContext.from_env()
    enc, dec = _models()
    x, y = _data()
    eng = ConvAEEngine(enc, dec, lr=1e-3, weight_decay=1e-5, device=torch.device("cuda", rank),
                       grad_hook=dp.allreduce_grads, grad_hook_async=dp.allreduce_grads_async, count_scale=1.0 / world,
                       dp=dp if fused else None)
    eng.overlap_allreduce = overlap          # two buckets, the first one reduced while the encoder backward runs
    data = eng.bind(x, y, 8)
    losses = []
    for _ in range(4):
        l = eng.train_epoch(data)
        losses.append(float(dp.reduce_losses(l).cpu()[0]))
    if rank == 0:
        sd = {k: v.detach().cpu().numpy() for k, v in list(enc.state_dict().items()) + list(dec.state_dict().items())}
        np.savez(os.path.join(out_dir, "dp.npz"), losses=np.array(losses), **{k.replace(".", "_"): v for k, v in sd.items()})
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True, "fused"])
def test_two_rank_nccl_equals_single_gpu(tmp_path, overlap):
    """overlap False / True: NCCL all-reduce (one call / two overlapped buckets); "fused": the all-reduce inside the optimiser
    launch over peer memory (csrc/dp_fused.cu)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    fused = overlap == "fused"
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), overlap is True, fused), nprocs=2, join=True)
    got = dict(np.load(os.path.join(str(tmp_path), "dp.npz")))
    from cae_tools_b200.engine.convae import ConvAEEngine
    enc, dec = _models()
    x, y = _data()
    eng = ConvAEEngine(enc, dec, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(x.repeat(2, 1, 1, 1), y.repeat(2, 1, 1, 1), 16)
    ref = [float(eng.train_epoch(data).cpu()[0]) for _ in range(4)]
    np.testing.assert_allclose(got["losses"], ref, rtol=2e-5)
    for k, v in list(enc.state_dict().items()) + list(dec.state_dict().items()):
        a, b = got[k.replace(".", "_")], v.detach().cpu().numpy()
        if k.endswith("running_var"):
            continue   # local BN: the unbiased correction M/(M-1) uses the per-rank element count
        if a.dtype.kind == "f":
            assert np.abs(a - b).max() <= 1e-4 * max(np.abs(b).max(), 1e-3), k


def _unet_models():
    import json
    from cae_tools_b200.models.model_sizer import ModelSpec
    from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = ModelSpec()
    spec.load(json.load(open(os.path.join(root, "cae_tools_b200", "specs", "unet_16x16_256x256.json"))))
    torch.manual_seed(31)
    return UNetEncoder(spec.get_input_layers(), 4, 16, 0.0), UNetDecoder(spec.get_output_layers(), 4, 16, 0.0)


def _unet_data():
    g = torch.Generator().manual_seed(32)
    return torch.rand(6, 1, 16, 16, generator=g), torch.rand(6, 1, 256, 256, generator=g)


def _unet_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cae_tools_b200.engine.dp import DPContext
    from cae_tools_b200.engine.unet import UNetEngine
    dp = DPContext.from_env()
    enc, dec = _unet_models()
    x, y = _unet_data()
    eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5, device=torch.device("cuda", rank),
                     grad_hook=dp.allreduce_grads, count_scale=1.0 / world, dp=dp)
    data = eng.bind(x, y, 6)
    losses = [float(dp.reduce_losses(eng.train_epoch(data)).cpu()[0]) for _ in range(3)]
    assert eng._train_stem(6) is not None
    fused = eng._dp_peers is not None
    if rank == 0:
        print("fused exchange:", fused, getattr(dp, "_symm_error", ""), flush=True)
        sd = {k: v.detach().cpu().numpy() for k, v in list(enc.state_dict().items()) + list(dec.state_dict().items())}
        np.savez(os.path.join(out_dir, "dpu.npz"), losses=np.array(losses), **{k.replace(".", "_"): v for k, v in sd.items()})
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_unet_fused_stem_equals_single_gpu(tmp_path):
    """the fused (cooperative) training stem under data parallelism: both ranks hold the same shard, so the 2-rank run must
    reproduce the single-GPU run on the duplicated batch"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_unet_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = dict(np.load(os.path.join(str(tmp_path), "dpu.npz")))
    from cae_tools_b200.engine.unet import UNetEngine
    enc, dec = _unet_models()
    x, y = _unet_data()
    eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(x.repeat(2, 1, 1, 1), y.repeat(2, 1, 1, 1), 12)
    ref = [float(eng.train_epoch(data).cpu()[0]) for _ in range(3)]
    np.testing.assert_allclose(got["losses"], ref, rtol=5e-5)
    for k, v in list(enc.state_dict().items()) + list(dec.state_dict().items()):
        a, b = got[k.replace(".", "_")], v.detach().cpu().numpy()
        if k.endswith("running_var") or a.dtype.kind != "f":
            continue
        assert np.abs(a - b).max() <= 2e-4 * max(np.abs(b).max(), 1e-3), k


def _linear_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cae_tools_b200.models.linear_model import LinearModel
    from oracle import datagen
    tr, te = datagen.circle_datasets(48, 16, output_size=(64, 64))
    torch.manual_seed(77)
    m = LinearModel(batch_size=16, nr_epochs=3, test_interval=1, lr=1e-3, weight_decay=1e-5)
    m.verbose = False
    m.train(["lowres"], "hires", tr, te)          # DPContext.from_env(): every batch of 16 is split 8 + 8
    if rank == 0:
        np.savez(os.path.join(out_dir, "dpl.npz"), train=np.array(m.history["train_loss"]), test=np.array(m.history["test_loss"]),
                 w=m.weights.linear[1].weight.detach().cpu().numpy(), b=m.weights.linear[1].bias.detach().cpu().numpy())
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_linear_model_equals_single_gpu(tmp_path):
    """LinearModel.train under data parallelism: batches of 16 split 8 + 8 over two ranks, gradients
    SUM-all-reduced - same loss history and weights as the single-process run"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_linear_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = dict(np.load(os.path.join(str(tmp_path), "dpl.npz")))
    from cae_tools_b200.models.linear_model import LinearModel
    from oracle import datagen
    tr, te = datagen.circle_datasets(48, 16, output_size=(64, 64))
    torch.manual_seed(77)
    m = LinearModel(batch_size=16, nr_epochs=3, test_interval=1, lr=1e-3, weight_decay=1e-5)
    m.verbose = False
    m.train(["lowres"], "hires", tr, te)
    np.testing.assert_allclose(got["train"], m.history["train_loss"], rtol=2e-5)
    np.testing.assert_allclose(got["test"], m.history["test_loss"], rtol=2e-5)
    w = m.weights.linear[1].weight.detach().cpu().numpy()
    assert np.abs(got["w"] - w).max() <= 1e-4 * np.abs(w).max()
    assert np.abs(got["b"] - m.weights.linear[1].bias.detach().cpu().numpy()).max() <= 1e-4 * np.abs(w).max()
