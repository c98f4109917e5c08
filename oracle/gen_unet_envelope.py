"""TEST INFRASTRUCTURE - how far does the REFERENCE's own 50-epoch UNET loss curve (tests/golden/curve_unet_b64_e50.npz) move
when only the CPU thread count changes, or when the dead (pre-BatchNorm) bias gradients are the exact zeros the CUDA path
writes?  Uses oracle/torch_port.OracleUNet, which reproduces the reference's 8-thread curve bit for bit (first line of the
output).  Writes tests/golden/curve_unet_b64_e50_envelope.npz.  Run in the build container: python oracle/gen_unet_envelope.py"""
import sys, json, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import load_npz, spec_of
from oracle import datagen
from oracle.torch_port import OracleUNet, make_batches, shuffled_order
from cae_tools_b200.models.model_sizer import ModelSpec
from cae_tools_b200.models.unet_modules import UNetDecoder, UNetEncoder
g = load_npz("curve_unet_b64_e50.npz")
spec_json = spec_of(g); spec = ModelSpec(); spec.load(spec_json)
tr, te = datagen.circle_datasets(100, 100)
lo_min, lo_max = float(tr["lowres"].data.min()), float(tr["lowres"].data.max())
hi_min, hi_max = float(tr["hires"].data.min()), float(tr["hires"].data.max())
norm = lambda a, lo, hi: ((a - lo) / (hi - lo)).astype(np.float32)
out={}
for threads, zero in ((8,False),(1,False),(4,False),(8,True)):
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    enc, dec = UNetEncoder(spec.get_input_layers(), 4, 16, 0.0), UNetDecoder(spec.get_output_layers(), 4, 16, 0.0)
    otr, ote = shuffled_order(100, 64), shuffled_order(100, 64)
    m = OracleUNet(enc.state_dict(), dec.state_dict(), spec_json, lambda_pearson=1.0, zero_dead_bias_grads=zero)
    btr = make_batches(norm(tr["lowres"].data, lo_min, lo_max), norm(tr["hires"].data, hi_min, hi_max), otr, 64)
    bte = make_batches(norm(te["lowres"].data, lo_min, lo_max), norm(te["hires"].data, hi_min, hi_max), ote, 64)
    tl, el = [], []
    for epoch in range(50):
        tl.append(float(np.mean([m.train_step(x, y, torch.ones_like(y))[0] for x, y in btr])))
        with torch.no_grad():
            el.append(float(np.mean([float(m.losses(x, y, torch.ones_like(y), False)[0]) for x, y in bte])))
    tl, el = np.array(tl), np.array(el)
    dt, de = np.abs(tl-g["train_loss"])/g["train_loss"], np.abs(el-g["test_loss"])/g["test_loss"]
    print(f"threads {threads} zero_dead {zero}: train max {dt.max():.2e} (first10 {dt[:10].max():.2e}) test max {de.max():.2e} (first10 {de[:10].max():.2e})", flush=True)
    out[f"train_t{threads}_z{int(zero)}"]=tl; out[f"test_t{threads}_z{int(zero)}"]=el
res = {"train_t1": out["train_t1_z0"], "test_t1": out["test_t1_z0"], "train_t4": out["train_t4_z0"], "test_t4": out["test_t4_z0"],
       "train_t8_zero_dead_bias": out["train_t8_z1"], "test_t8_zero_dead_bias": out["test_t8_z1"]}
np.savez_compressed('/root/repo/tests/golden/curve_unet_b64_e50_envelope.npz', **res)
