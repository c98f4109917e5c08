"""Encoder parameter container.

Same module tree, construction order (=> same RNG stream and default PyTorch init) and state_dict
keys as the reference ``Encoder`` (reference: src/cae_tools/models/encoder.py:34-64):
``encoder_cnn.{3i}`` Conv2d / ``{3i+1}`` BatchNorm2d / ``{3i+2}`` ReLU per layer, then
``encoder_lin`` = Linear(C*H*W, fc) - ReLU - Linear(fc, latent).

The module only *holds* parameters; arithmetic runs in the sm_100a kernels
(cae_tools_b200.engine).  ``forward`` is an inference-mode convenience that drives those kernels.
"""

import torch
from torch import nn


class Encoder(nn.Module):

    def __init__(self, layers, encoded_space_dim, fc_size):
        super().__init__()
        self.layer_specs = list(layers)
        stack = []
        for spec in layers:
            cin = spec.get_input_dimensions()[0]
            cout = spec.get_output_dimensions()[0]
            stack += [nn.Conv2d(cin, cout, kernel_size=spec.get_kernel_size(), stride=spec.get_stride()),
                      nn.BatchNorm2d(cout),
                      nn.ReLU(True)]
        self.encoder_cnn = nn.Sequential(*stack)
        self.flatten = nn.Flatten(start_dim=1)
        c, h, w = layers[-1].get_output_dimensions()
        self.encoder_lin = nn.Sequential(nn.Linear(c * h * w, fc_size), nn.ReLU(True),
                                         nn.Linear(fc_size, encoded_space_dim))

    def conv_layers(self):
        """[(conv, bn)] in forward order"""
        mods = list(self.encoder_cnn)
        return [(mods[i], mods[i + 1]) for i in range(0, len(mods), 3)]

    def forward(self, x):
        from ..engine.eager import encoder_forward
        return encoder_forward(self, x)
