#!/usr/bin/env python
"""Where does the 50-epoch unet run end up relative to the reference's own run (tests/golden/curve_unet_b64_e50.npz)?
Prints the per-tensor distance of the final state (weights, BatchNorm running statistics) and the loss-curve deviations."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cae_tools_b200.models.model_sizer import ModelSpec  # noqa: E402
from cae_tools_b200.models.unet import UNET  # noqa: E402
from oracle import datagen  # noqa: E402

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 50
g = np.load(os.path.join(ROOT, "tests", "golden", "curve_unet_b64_e50.npz"), allow_pickle=True)
tr, te = datagen.circle_datasets(100, 100)
torch.manual_seed(1234)
m = UNET(batch_size=64, nr_epochs=epochs, test_interval=1, encoded_dim_size=4, fc_size=16, lr=1e-3, weight_decay=1e-5,
         dropout_rate=0.0, lambda_pearson=1.0)
m.verbose = False
spec = ModelSpec()
spec.load(json.loads(str(g["spec_json"])))
m.spec = spec
m.train(["lowres"], "hires", tr, te)
got_tr, got_te = np.array(m.history["train_loss"]), np.array(m.history["test_loss"])
dtr = np.abs(got_tr - g["train_loss"][:epochs]) / g["train_loss"][:epochs]
dte = np.abs(got_te - g["test_loss"][:epochs]) / g["test_loss"][:epochs]
print("train dev per epoch (1e-4):", np.round(dtr * 1e4, 1))
print("test  dev per epoch (1e-4):", np.round(dte * 1e4, 1))
if epochs == 50:
    sd = {("final.enc." + k): v for k, v in m.encoder.state_dict().items()}
    sd.update({("final.dec." + k): v for k, v in m.decoder.state_dict().items()})
    for k in sorted(sd):
        if k in g.files and "num_batches" not in k:
            a, b = sd[k].detach().cpu().double().numpy(), np.asarray(g[k], dtype=np.float64)
            print(f"{k:55s} max-norm rel {np.abs(a - b).max() / max(np.abs(b).max(), 1e-12):.2e}")
