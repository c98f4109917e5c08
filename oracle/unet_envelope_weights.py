"""TEST INFRASTRUCTURE - distance between the FINAL WEIGHTS of two 50-epoch runs of the reference unet loop that differ only in the
CPU thread count (1 thread here vs the 8-thread run stored in tests/golden/curve_unet_b64_e50.npz).  Measured in the build
container: up to 61 % max-norm relative (encoder_cnn.9.bias), 50 % on decoder_conv.0.weight - the same size as the distance of
the CUDA run from the stored one (tools/unet_curve_probe.py: 43 % / 45 %), while all loss curves agree to ~1e-3.
Run: python oracle/unet_envelope_weights.py"""
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
torch.set_num_threads(1)
torch.manual_seed(1234)
enc, dec = UNetEncoder(spec.get_input_layers(), 4, 16, 0.0), UNetDecoder(spec.get_output_layers(), 4, 16, 0.0)
otr, ote = shuffled_order(100, 64), shuffled_order(100, 64)
m = OracleUNet(enc.state_dict(), dec.state_dict(), spec_json, lambda_pearson=1.0)
btr = make_batches(norm(tr["lowres"].data, lo_min, lo_max), norm(tr["hires"].data, hi_min, hi_max), otr, 64)
for epoch in range(50):
    for x, y in btr: m.train_step(x, y, torch.ones_like(y))
sd = {}
for pref, mod in (("final.enc.", m.enc), ("final.dec.", m.dec)):
    items = mod.state_dict().items() if hasattr(mod, 'state_dict') else mod.items()
    for k, v in items: sd[pref + k] = v
worst = []
for k in sorted(sd):
    if k in g and "num_batches" not in k:
        a, b = sd[k].detach().double().numpy(), np.asarray(g[k], dtype=np.float64)
        worst.append((np.abs(a-b).max()/max(np.abs(b).max(),1e-12), k))
for w, k in sorted(worst, reverse=True)[:12]: print(f"{k:50s} {w:.2e}")
