"""host geometry == reference create_model_spec (reference: models/model_sizer.py:112-162), via golden specs
written by oracle/gen_golden.py from the live reference."""
import json
import os

from cae_tools_b200.models.model_sizer import ModelSpec, create_model_spec


def test_specs_match_reference(golden_dir):
    with open(os.path.join(golden_dir, "specs.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 6
    for name, case in cases.items():
        args = {k: tuple(v) if isinstance(v, list) else v for k, v in case["args"].items()}
        spec = create_model_spec(kernel_size=3, stride=2, **args)
        assert spec.save() == case["spec"], name


def test_spec_roundtrip_and_tuple_kernels(golden_dir):
    with open(os.path.join(golden_dir, "specs.json")) as f:
        cases = json.load(f)
    obj = cases["circle2_24x20_280x256"]["spec"]
    spec = ModelSpec()
    spec.load(obj)
    assert spec.save() == obj
    kernels = [layer.get_kernel_size() for layer in spec.get_output_layers()]
    assert (4, 3) in kernels  # non-square geometry produces tuple kernels
    assert "kernel_size=(4, 3)" in repr(spec)


def test_config1_geometry():
    spec = create_model_spec(input_size=(16, 16), input_channels=1, output_size=(256, 256), output_channels=1)
    assert [l.get_output_dimensions() for l in spec.get_input_layers()] == [(2, 7, 7), (4, 3, 3)]
    outs = [l.get_output_dimensions() for l in spec.get_output_layers()]
    assert outs == [(32, 7, 7), (16, 15, 15), (8, 31, 31), (4, 63, 63), (2, 127, 127), (1, 256, 256)]
    assert spec.get_output_layers()[-1].get_kernel_size() == 4
