"""Encoder container of the variational variant: the reference's conv stack (encoder.py:39-48) followed by
Linear(C*H*W, fc) - ReLU and two heads, ``fc_mu`` and ``fc_logvar`` (Linear(fc, latent) each).
The reference ships no VarAEModel; key names here are this repository's choice."""

from torch import nn


class VarEncoder(nn.Module):

    def __init__(self, layers, encoded_space_dim, fc_size):
        super().__init__()
        self.layer_specs = list(layers)
        stack = []
        for spec in layers:
            cin, cout = spec.get_input_dimensions()[0], spec.get_output_dimensions()[0]
            stack += [nn.Conv2d(cin, cout, kernel_size=spec.get_kernel_size(), stride=spec.get_stride()),
                      nn.BatchNorm2d(cout), nn.ReLU(True)]
        self.encoder_cnn = nn.Sequential(*stack)
        self.flatten = nn.Flatten(start_dim=1)
        c, h, w = layers[-1].get_output_dimensions()
        self.encoder_lin = nn.Sequential(nn.Linear(c * h * w, fc_size), nn.ReLU(True))
        self.fc_mu = nn.Linear(fc_size, encoded_space_dim)
        self.fc_logvar = nn.Linear(fc_size, encoded_space_dim)

    def conv_layers(self):
        mods = list(self.encoder_cnn)
        return [(mods[i], mods[i + 1]) for i in range(0, len(mods), 3)]
