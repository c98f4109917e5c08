"""Decoder parameter container.

Same module tree, construction order, custom initialisation and state_dict keys as the reference
``Decoder`` (reference: src/cae_tools/models/decoder.py:22-78): ``decoder_lin`` =
Linear(latent, fc) - ReLU - Linear(fc, C*H*W); ``decoder_conv`` = ConvTranspose2d (+ BatchNorm2d + ReLU
on all but the last layer); sigmoid on the output.  Initialisation (decoder.py:55-71): transposed
convs and the first Linear use Kaiming-normal(fan_out, relu), the Linear that feeds the conv stack
uses Xavier-normal, biases are zero, BatchNorm is (1, 0).

The module only *holds* parameters; arithmetic runs in the sm_100a kernels.
"""

import torch.nn.init as init
from torch import nn


class Decoder(nn.Module):

    def __init__(self, layers, encoded_space_dim, fc_size):
        super().__init__()
        self.layer_specs = list(layers)
        self.chan, self.y, self.x = layers[0].get_input_dimensions()
        flat = self.chan * self.y * self.x
        self.decoder_lin = nn.Sequential(nn.Linear(encoded_space_dim, fc_size), nn.ReLU(True),
                                         nn.Linear(fc_size, flat))
        self.unflatten = nn.Unflatten(dim=1, unflattened_size=(self.chan, self.y, self.x))
        stack = []
        for idx, spec in enumerate(layers):
            cin = spec.get_input_dimensions()[0]
            cout = spec.get_output_dimensions()[0]
            stack.append(nn.ConvTranspose2d(cin, cout, kernel_size=spec.get_kernel_size(), stride=spec.get_stride(),
                                            output_padding=spec.get_output_padding()))
            if idx != len(layers) - 1:
                stack += [nn.BatchNorm2d(cout), nn.ReLU(True)]
        self.decoder_conv = nn.Sequential(*stack)
        self._initialize_weights()

    def _initialize_weights(self):
        flat = self.chan * self.y * self.x
        for m in self.modules():
            if isinstance(m, nn.ConvTranspose2d):
                init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                if m.out_features == flat:
                    init.xavier_normal_(m.weight)
                else:
                    init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                init.constant_(m.weight, 1)
                init.constant_(m.bias, 0)

    def conv_layers(self):
        """[(convT, bn or None)] in forward order"""
        out, mods, i = [], list(self.decoder_conv), 0
        while i < len(mods):
            conv = mods[i]
            bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d) else None
            out.append((conv, bn))
            i += 3 if bn is not None else 1
        return out

    def forward(self, x):
        from ..engine.eager import decoder_forward
        return decoder_forward(self, x)
