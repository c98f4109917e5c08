import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import load_npz, unet_light_data, spec_of, split_sd
from test_gpu_unet import _build, _dead
from cae_tools_b200.engine.unet import UNetEngine
from oracle.torch_port import OracleUNet
g = load_npz("unet_head32_mask_light.npz")
spec, enc, dec = _build(g)
x, y, mask = unet_light_data(g)
ex = OracleUNet(split_sd(g, "init.enc."), split_sd(g, "init.dec."), spec_of(g), lambda_pearson=1.0, dtype=torch.float64)
ex.train_step(x.double(), y.double(), mask.double())
for fused in (True, False):
    spec, enc, dec = _build(g)
    UNetEngine.use_fused_train_stem = fused
    eng = UNetEngine(enc, dec, lambda_pearson=1.0, dropout_rate=0.0, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(x, y, x.shape[0], mask=mask)
    eng.train_epoch(data)
    rows=[]
    for prefix, mod, sd64 in (("enc.", enc, ex.enc), ("dec.", dec, ex.dec)):
        for k, p in mod.named_parameters():
            if _dead(k): continue
            ref = g["grad." + prefix + k]; r64 = sd64[k].grad.numpy(); got = p.grad.detach().cpu().numpy()
            sc = max(np.abs(r64).max(), 1e-12)
            rows.append((np.abs(got-r64).max()/sc, np.abs(ref-r64).max()/sc, prefix+k))
    rows.sort(reverse=True)
    print("fused stem" if fused else "chain", "(gpu vs f64, reference vs f64)")
    for r in rows[:8]: print(f"   {r[0]:.2e} {r[1]:.2e} {r[2]}")
