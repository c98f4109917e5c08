"""Parameter containers of the UNET variant: same module tree, construction order (=> RNG stream, default PyTorch
initialisation) and state_dict keys as the reference (reference: src/cae_tools/models/unet.py:23-39 ChannelAttention,
:73-112 Encoder, :114-163 Decoder).  The modules only hold parameters; arithmetic runs in the sm_100a kernels
(engine/unet.py).

Layer specs are used the way the reference uses them: ``spec.output_padding`` is the *padding* of both the Conv2d
(unet.py:82) and the ConvTranspose2d (unet.py:140) layers."""

from torch import nn


class ChannelAttention(nn.Module):
    def __init__(self, in_planes, ratio=8):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.max_pool = nn.AdaptiveMaxPool2d(1)
        self.fc1 = nn.Conv2d(in_planes, in_planes // ratio, 1, bias=False)
        self.relu1 = nn.ReLU()
        self.fc2 = nn.Conv2d(in_planes // ratio, in_planes, 1, bias=False)
        self.sigmoid = nn.Sigmoid()


class UNetEncoder(nn.Module):
    def __init__(self, layers, encoded_space_dim, fc_size, dropout_rate=0.1):
        super().__init__()
        self.layer_specs = list(layers)
        mods = []
        for spec in layers:
            cin, cout = spec.get_input_dimensions()[0], spec.get_output_dimensions()[0]
            mods += [nn.Conv2d(cin, cout, kernel_size=spec.get_kernel_size(), stride=spec.get_stride(),
                               padding=spec.get_output_padding()),
                     nn.BatchNorm2d(cout), nn.ReLU(True), nn.Dropout(dropout_rate)]
        self.encoder_cnn = nn.ModuleList(mods)
        self.flatten = nn.Flatten(start_dim=1)
        c, h, w = layers[-1].get_output_dimensions()
        self.encoder_lin = nn.Sequential(nn.Linear(c * h * w, fc_size), nn.BatchNorm1d(fc_size), nn.ReLU(True),
                                         nn.Dropout(dropout_rate), nn.Linear(fc_size, encoded_space_dim), nn.ReLU(True),
                                         nn.Dropout(dropout_rate))

    def conv_layers(self):
        m = list(self.encoder_cnn)
        return [(m[i], m[i + 1]) for i in range(0, len(m), 4)]


class UNetDecoder(nn.Module):
    def __init__(self, layers, encoded_space_dim, fc_size, dropout_rate=0.1):
        super().__init__()
        self.layer_specs = list(layers)
        self.chan, self.y, self.x = layers[0].get_input_dimensions()
        flat = self.chan * self.y * self.x
        self.decoder_lin = nn.Sequential(nn.Linear(encoded_space_dim, fc_size), nn.BatchNorm1d(fc_size), nn.ReLU(True),
                                         nn.Dropout(dropout_rate), nn.Linear(fc_size, flat), nn.ReLU(True),
                                         nn.Dropout(dropout_rate))
        self.unflatten = nn.Unflatten(dim=1, unflattened_size=(self.chan, self.y, self.x))
        mods = []
        self.attention_layers = nn.ModuleList()
        for idx, spec in enumerate(layers):
            cin, cout = spec.get_input_dimensions()[0], spec.get_output_dimensions()[0]
            mods.append(nn.ConvTranspose2d(cin, cout, kernel_size=spec.get_kernel_size(), stride=spec.get_stride(),
                                           padding=spec.get_output_padding()))
            if idx != len(layers) - 1:
                self.attention_layers.append(ChannelAttention(cout))
                mods += [nn.BatchNorm2d(cout * 2), nn.ReLU(True), nn.Dropout(dropout_rate)]
        self.decoder_conv = nn.ModuleList(mods)

    def conv_layers(self):
        """[(convT, bn2c or None, attention or None)] in forward order"""
        out, m, i, j = [], list(self.decoder_conv), 0, 0
        while i < len(m):
            if i + 1 < len(m) and isinstance(m[i + 1], nn.BatchNorm2d):
                out.append((m[i], m[i + 1], self.attention_layers[j]))
                j += 1
                i += 4
            else:
                out.append((m[i], None, None))
                i += 1
        return out
