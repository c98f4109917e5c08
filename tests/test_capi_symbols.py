"""The C-ABI library loads on a machine without a GPU and exports every symbol include/cae_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "cae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(cae_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from cae_tools_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    names = declared_functions()
    assert len(names) >= 12
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in cae_b200.h but not exported"
    assert sorted(_lib.EXPORTS.keys()) == names, "ctypes table and header disagree"
    lib = _lib.lib()
    assert lib.cae_version() >= 100
    assert lib.cae_partials_len(3) > 0
    assert lib.cae_last_error() is not None


def test_struct_sizes_match_c_layout():
    """ctypes mirrors of the PODs: sizes follow the natural C layout the header implies"""
    from cae_tools_b200 import _lib
    assert ctypes.sizeof(_lib.CaeView) == 8 + 5 * 4 + 4 + 16          # ptr, 5 ints (+pad), 2 long long
    assert ctypes.sizeof(_lib.CaeConvGeom) == 16
    assert ctypes.sizeof(_lib.CaeSrc) == ctypes.sizeof(_lib.CaeView) + 4 * 8 + 8 + 8 + 8 + 8
    assert ctypes.sizeof(_lib.CaeBN) == 16 + 15 * 8
    # ... and every mirror has exactly the size the compiled library sees
    lib = _lib.lib()
    for which, cls in enumerate([_lib.CaeView, _lib.CaeSrc, _lib.CaeConvGeom, _lib.CaeBN, _lib.CaeEpilogue, _lib.CaeGemm,
                                 _lib.CaePatchHead, _lib.CaeFcStack, _lib.CaeUnetStem, _lib.CaeTcGemm, _lib.CaeTcConv, _lib.CaeStemTrain, _lib.CaeDpPeers]):
        assert lib.cae_struct_size(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.cae_struct_size(99) == -1


def test_product_path_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cae_tools_b200._lib import CaeError
    from cae_tools_b200.models.conv_ae_model import ConvAEModel
    from oracle import datagen
    tr, te = datagen.circle_datasets(4, 4, input_size=(16, 16), output_size=(64, 64))
    m = ConvAEModel(batch_size=2, nr_epochs=1)
    m.verbose = False
    with pytest.raises(CaeError):
        m.train(["lowres"], "hires", tr, te)
    m2 = ConvAEModel(use_gpu=False)
    with pytest.raises(CaeError):
        m2.train(["lowres"], "hires", tr, te)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under cae_tools_b200/ may import or execute it"""
    pkg = os.path.join(ROOT, "cae_tools_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "import_module(\"oracle" in text:
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
