"""VarAEModel (no reference implementation -> parity unpinned): the CUDA path against the
plain-PyTorch definition in oracle/torch_port.OracleVarModel with the SAME noise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _models(lambda_mse=1.0, lambda_kl=0.5):
    from cae_tools_b200.models.decoder import Decoder
    from cae_tools_b200.models.model_sizer import create_model_spec
    from cae_tools_b200.models.var_encoder import VarEncoder
    from oracle.torch_port import OracleVarModel
    torch.manual_seed(5)
    spec = create_model_spec(input_size=(16, 16), input_channels=1, output_size=(64, 64), output_channels=1)
    enc, dec = VarEncoder(spec.get_input_layers(), 6, 16), Decoder(spec.get_output_layers(), 6, 16)
    oracle = OracleVarModel(enc.state_dict(), dec.state_dict(), spec.save(), lambda_mse=lambda_mse,
                            lambda_kl=lambda_kl, zero_dead_bias_grads=True)
    return spec, enc, dec, oracle


def test_var_train_steps_match_oracle_with_shared_noise():
    from cae_tools_b200.engine.varae import VarAEEngine
    spec, enc, dec, oracle = _models()
    g = torch.Generator().manual_seed(9)
    x, y = torch.rand(12, 1, 16, 16, generator=g), torch.rand(12, 1, 64, 64, generator=g)
    eps = torch.randn(12, 6, generator=g)
    eng = VarAEEngine(enc, dec, lambda_mse=1.0, lambda_kl=0.5, lr=1e-3, weight_decay=1e-5)
    data = eng.bind(x, y, 6, fixed_eps=eps)          # two batches of 6
    for epoch in range(3):
        got = eng.train_epoch(data).cpu().numpy()
        want = [float(oracle.train_step(x[i:i + 6], y[i:i + 6], eps[i:i + 6])) for i in (0, 6)]
        np.testing.assert_allclose(got, want, rtol=5e-5)
    for sd, mod in ((oracle.enc, enc), (oracle.dec, dec)):
        for k, v in mod.state_dict().items():
            ref = sd[k].detach().numpy()
            got_v = v.detach().cpu().numpy()
            if ref.dtype.kind == "f":
                assert np.abs(got_v - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-3), k
    out = []
    eng.score_batches(eng.bind(x, None, 12), lambda i, yh: out.append(yh.cpu().numpy().copy()))
    ref = oracle.score(x).numpy()
    assert np.abs(out[0] - ref).max() / np.abs(ref).max() < 1e-4
    # eval-mode loss (z = mu)
    test_losses = eng.test_epoch(eng.bind(x, y, 12)).cpu().numpy()
    with torch.no_grad():
        want, _ = oracle.loss(x, y, None, False)
    assert abs(test_losses[0] - float(want)) <= 1e-4 * float(want)


def test_device_noise_is_standard_normal_and_fresh_every_step():
    from cae_tools_b200.engine import ops
    n = 1 << 20
    out = torch.empty(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.randn(out, n, 1234, step)
    a = out.cpu().double()
    step.fill_(1)
    ops.randn(out, n, 1234, step)
    b = out.cpu().double()
    for t in (a, b):
        assert abs(float(t.mean())) < 5e-3 and abs(float(t.var()) - 1.0) < 1e-2
        assert abs(float((t ** 3).mean())) < 2e-2 and abs(float((t ** 4).mean()) - 3.0) < 5e-2
        assert torch.isfinite(t).all()
    assert abs(float((a * b).mean())) < 5e-3            # different steps are uncorrelated
    assert float((a[:-1] * a[1:]).mean()) < 5e-3        # neighbours too


def test_var_model_api_trains_and_applies(tmp_path):
    from cae_tools_b200.models.var_ae_model import VarAEModel
    from oracle import datagen
    tr, te = datagen.circle_datasets(40, 20, input_size=(16, 16), output_size=(64, 64))
    torch.manual_seed(3)
    m = VarAEModel(lambda_mse=1.0, lambda_kl=1e-3, batch_size=8, nr_epochs=30, test_interval=5, encoded_dim_size=4,
                   fc_size=16)
    m.verbose = False
    m.train(["lowres"], "hires", tr, te)
    assert len(m.history["train_loss"]) == 6 and m.history["train_loss"][-1] < m.history["train_loss"][0]
    folder = str(tmp_path / "var")
    m.save(folder)
    m2 = VarAEModel()
    m2.load(folder)
    assert m2.get_parameters()["type"] == "VarAEModel" and m2.lambda_kl == 1e-3
    m2.apply(te, ["lowres"], "est")
    m.apply(te, ["lowres"], "est0")
    np.testing.assert_allclose(np.asarray(te["est"].data), np.asarray(te["est0"].data), rtol=1e-6)


def test_var_kl_recorded_for_every_batch_with_device_noise():
    """the KL term of every batch lands in its own slot when the noise is drawn on the device (the path VarAEModel.train
    uses): the reported epoch loss is mean(mse + kl), not mean(mse) + kl_last / n_batches"""
    from cae_tools_b200.engine.varae import VarAEEngine
    from cae_tools_b200.models.decoder import Decoder
    from cae_tools_b200.models.model_sizer import create_model_spec
    from cae_tools_b200.models.var_encoder import VarEncoder
    torch.manual_seed(5)
    spec = create_model_spec(input_size=(16, 16), input_channels=1, output_size=(64, 64), output_channels=1)
    enc, dec = VarEncoder(spec.get_input_layers(), 4, 16), Decoder(spec.get_output_layers(), 4, 16)
    eng = VarAEEngine(enc, dec, lambda_mse=1.0, lambda_kl=0.5, seed=7)
    x, y = torch.rand(24, 1, 16, 16), torch.rand(24, 1, 64, 64)
    data = eng.bind(x, y, 8)                       # 3 batches, noise drawn by cae_randn
    losses = eng.train_epoch(data).cpu()
    kl = data.kl.cpu()
    assert kl.shape == (3,) and bool((kl > 0).all()), kl
    assert torch.allclose(losses, data.losses.cpu() + kl)
