"""VarAEModel: variational auto-encoder with the ConvAEModel API (`--method var|vae`).

The reference refers to this class (`cae_tools.models.var_ae_model.VarAEModel`: model_evaluator.py:35; type string
"VarAEModel": model_evaluator.py:74-75; CLI flags `--lambda-mse`, `--lambda-kl`: cli/train_cae.py:32-33) but the
module is absent from the reference snapshot, so the model is defined here: the reference's encoder / decoder
stacks with mu / log-variance heads, the reparameterisation trick and
``loss = lambda_mse * MSE + lambda_kl * KL``.  history / save / load / apply behave like ConvAEModel.
Parity: unpinned by the reference (no code to compare with); pinned against oracle/torch_port.OracleVarModel.
"""

from __future__ import annotations

from .conv_ae_model import ConvAEModel
from .decoder import Decoder
from .var_encoder import VarEncoder


class VarAEModel(ConvAEModel):

    MODEL_TYPE = "VarAEModel"
    DB_TYPE = "VarAE"

    def __init__(self, lambda_mse=1.0, lambda_kl=1.0, seed=0, **kwargs):
        super().__init__(**kwargs)
        self.lambda_mse = lambda_mse
        self.lambda_kl = lambda_kl
        self.seed = seed

    def get_parameters(self):
        p = super().get_parameters()
        p.update({"lambda_mse": self.lambda_mse, "lambda_kl": self.lambda_kl})
        return p

    def _load_parameters(self, parameters):
        super()._load_parameters(parameters)
        self.lambda_mse = parameters.get("lambda_mse", 1.0)
        self.lambda_kl = parameters.get("lambda_kl", 1.0)

    def _build_modules(self):
        self.encoder = VarEncoder(self.spec.get_input_layers(), encoded_space_dim=self.encoded_dim_size,
                                  fc_size=self.fc_size)
        self.decoder = Decoder(self.spec.get_output_layers(), encoded_space_dim=self.encoded_dim_size,
                               fc_size=self.fc_size)

    def _make_engine(self, device, dp=None):
        from ..engine.varae import VarAEEngine
        kw = dict(lr=self.lr, weight_decay=self.weight_decay, device=device)
        seed = self.seed
        if dp is not None:
            kw.update(grad_hook=dp.allreduce_grads, count_scale=1.0 / dp.world,
                      grad_hook_async=dp.allreduce_grads_async, dp=dp)
            # every rank holds a different shard: its reparameterisation noise must be independent too
            seed = (seed ^ ((dp.rank + 1) * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF
        return VarAEEngine(self.encoder, self.decoder, lambda_mse=self.lambda_mse, lambda_kl=self.lambda_kl,
                           seed=seed, **kw)
