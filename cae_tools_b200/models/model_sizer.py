"""Layer geometry for the convolutional auto-encoder.

Host-side only.  Mirrors the behaviour of the reference sizer
(reference: src/cae_tools/models/model_sizer.py:16-67 ``LayerSpec``,
:70-109 ``ModelSpec``, :112-162 ``create_model_spec``) because every kernel
shape and the on-disk ``spec.json`` derive from it:

* encoder layers: no padding, ``out = (in - k) // stride + 1``, channels double,
  stop once the next layer would be smaller than ``limit`` (or the requested
  layer count is reached) - always at least one layer;
* decoder layers are derived backwards from the output size: the kernel grows
  (per axis) until ``(out - k) % stride == 0`` so that a transposed convolution
  with no output padding lands exactly on ``out``; channels double going inwards.
"""

from __future__ import annotations


def _pair(v):
    """kernel sizes may be an int or an (h, w) pair"""
    if isinstance(v, (tuple, list)):
        return int(v[0]), int(v[1])
    return int(v), int(v)


class LayerSpec:
    """One (transposed) convolution layer: kernel, stride, (C,H,W) in and out."""

    def __init__(self, is_input=True, kernel_size=3, stride=2, input_dimensions=None,
                 output_dimensions=None, output_padding=0):
        self.is_input = is_input
        self.kernel_size = kernel_size  # int or (h, w)
        self.stride = stride
        self.input_dimensions = input_dimensions
        self.output_dimensions = output_dimensions
        self.output_padding = output_padding

    # accessors used by the nn.Module containers (same names as the reference)
    def get_kernel_size(self):
        return self.kernel_size

    def get_stride(self):
        return self.stride

    def get_input_dimensions(self):
        return self.input_dimensions

    def get_output_dimensions(self):
        return self.output_dimensions

    def get_output_padding(self):
        return self.output_padding

    def kernel_hw(self):
        return _pair(self.kernel_size)

    def __repr__(self):
        head = "\tInput Convolutional Layer:\n" if self.is_input else "\tOutput Convolutional Layer:\n"
        text = head + f"\t\tkernel_size={self.kernel_size}  stride={self.stride}\n"
        if self.output_padding:
            text += f"\t\toutput_padding=({self.output_padding})\n"
        text += f"\t\t{self.input_dimensions} => {self.output_dimensions}\n"
        return text

    def save(self):
        k = self.kernel_size
        return {
            "is_input": self.is_input,
            "kernel_size": list(k) if isinstance(k, tuple) else k,
            "stride": self.stride,
            "output_padding": self.output_padding,
            "input_dimensions": list(self.input_dimensions),
            "output_dimensions": list(self.output_dimensions),
        }

    def load(self, from_obj):
        self.is_input = from_obj["is_input"]
        k = from_obj["kernel_size"]
        self.kernel_size = tuple(k) if isinstance(k, list) else k
        self.stride = from_obj["stride"]
        self.output_padding = from_obj["output_padding"]
        self.input_dimensions = tuple(from_obj["input_dimensions"])
        self.output_dimensions = tuple(from_obj["output_dimensions"])


class ModelSpec:
    """Ordered encoder ("input") and decoder ("output") layer lists."""

    def __init__(self, input_layer_specs=None, output_layer_specs=None):
        self.input_layers = list(input_layer_specs) if input_layer_specs is not None else []
        self.output_layers = list(output_layer_specs) if output_layer_specs is not None else []

    def get_input_layers(self):
        return self.input_layers

    def get_output_layers(self):
        return self.output_layers

    def save(self):
        return {
            "input_layers": [layer.save() for layer in self.input_layers],
            "output_layers": [layer.save() for layer in self.output_layers],
        }

    def load(self, from_obj):
        self.input_layers = []
        self.output_layers = []
        for key, dest in (("input_layers", self.input_layers), ("output_layers", self.output_layers)):
            for obj in from_obj[key]:
                layer = LayerSpec()
                layer.load(obj)
                dest.append(layer)

    def __repr__(self):
        text = "Input Layers:\n" + "".join(str(layer) for layer in self.input_layers)
        text += "Output Layers:\n" + "".join(str(layer) for layer in self.output_layers)
        return text


def _down(size, k, stride):
    return (size - k) // stride + 1


def create_model_spec(input_size=(7, 7), input_channels=1, output_size=(28, 28), output_channels=1, stride=2,
                      kernel_size=3, limit=3, input_layer_count=None, output_layer_count=None):
    """Derive encoder and decoder layer lists (reference: model_sizer.py:112-162)."""
    # ---- encoder: shrink until the next layer would fall below `limit`
    enc = []
    (h, w) = input_size
    c = input_channels
    while True:
        nh, nw = _down(h, kernel_size, stride), _down(w, kernel_size, stride)
        if enc:
            enough = input_layer_count is not None and len(enc) >= input_layer_count
            if enough or min(nh, nw) < limit:
                break
        enc.append(LayerSpec(True, kernel_size, stride, (int(c), int(h), int(w)), (int(2 * c), int(nh), int(nw))))
        c, h, w = 2 * c, nh, nw

    # ---- decoder: walk from the output size inwards until we reach the encoder's final size
    (small_h, small_w) = (h, w)
    dec = []
    (h, w) = output_size
    c = output_channels
    while True:
        if dec:
            enough = output_layer_count is not None and len(dec) >= output_layer_count
            if enough or w <= small_w or h <= small_h:
                break
        kh = kw = kernel_size
        while (w - kw) % stride != 0:
            kw += 1
        while (h - kh) % stride != 0:
            kh += 1
        k = kw if kh == kw else (kh, kw)
        ih, iw = _down(h, kh, stride), _down(w, kw, stride)
        dec.insert(0, LayerSpec(False, k, stride, (int(2 * c), int(ih), int(iw)), (int(c), int(h), int(w))))
        c, h, w = 2 * c, ih, iw

    return ModelSpec(enc, dec)
