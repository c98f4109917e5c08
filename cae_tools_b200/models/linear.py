"""Container of the `--method linear` model: one nn.Linear from all input pixels to all output pixels
(reference: src/cae_tools/models/linear.py:33-49 - same module tree, so state_dict keys `linear.1.{weight,bias}` and the
default nn.Linear initialisation / RNG consumption are the reference's).  forward() is not the product path: the
arithmetic runs in libcae_b200 (engine/linear.py)."""

from torch import nn


class Linear(nn.Module):

    def __init__(self, input_shape, output_shape):
        super().__init__()
        (chan1, y1, x1) = input_shape
        (chan2, y2, x2) = output_shape
        self.input_shape, self.output_shape = tuple(input_shape), tuple(output_shape)
        self.linear = nn.Sequential(
            nn.Flatten(start_dim=1),
            nn.Linear(chan1 * y1 * x1, chan2 * y2 * x2),
            nn.Unflatten(dim=1, unflattened_size=(chan2, y2, x2)),
        )

    def forward(self, x):
        raise RuntimeError("cae_tools_b200: the Linear container holds parameters only; LinearEngine runs the kernels")
