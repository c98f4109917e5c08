"""Parameter container of the `--method linear` model: a single fully connected map from all input pixels to all output
pixels.  It mirrors the module tree of the reference's model (src/cae_tools/models/linear.py:33-49: a three-stage
`linear` Sequential whose middle stage is the nn.Linear), so checkpoints carry the same keys - `linear.1.weight`,
`linear.1.bias` - and construction consumes the torch RNG identically.  The arithmetic does not run through
forward(): engine/linear.py drives libcae_b200."""

import math

from torch import nn


class Linear(nn.Module):

    def __init__(self, input_shape, output_shape):
        super().__init__()
        self.input_shape = tuple(int(v) for v in input_shape)       # (channels, y, x)
        self.output_shape = tuple(int(v) for v in output_shape)
        n_in, n_out = math.prod(self.input_shape), math.prod(self.output_shape)
        stages = [nn.Flatten(1), nn.Linear(n_in, n_out), nn.Unflatten(1, self.output_shape)]
        self.linear = nn.Sequential(*stages)

    def forward(self, x):
        raise RuntimeError("cae_tools_b200: the Linear container holds parameters only; LinearEngine runs the kernels")
