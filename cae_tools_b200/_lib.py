"""ctypes binding of libcae_b200.so (see include/cae_b200.h).

There is deliberately no fallback: if the shared library is missing or was built for the
wrong architecture, importing :func:`lib` raises and every product path that needs
arithmetic fails loudly.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcae_b200.so")

c_float_p = C.POINTER(C.c_float)


class CaeView(C.Structure):
    _fields_ = [("p", C.c_void_p), ("N", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("ld", C.c_int), ("sC", C.c_longlong), ("sN", C.c_longlong)]


class CaeSrc(C.Structure):
    _fields_ = [("t0", CaeView), ("t1", C.c_void_p), ("k0", C.c_void_p), ("k1", C.c_void_p), ("k2", C.c_void_p),
                ("relu", C.c_int), ("cursor", C.c_void_p), ("cursor_stride", C.c_longlong), ("kn", C.c_void_p)]


class CaeConvGeom(C.Structure):
    _fields_ = [("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("pad", C.c_int)]


class CaeBN(C.Structure):
    _fields_ = [("C", C.c_int), ("eps", C.c_float), ("momentum", C.c_float),
                ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p), ("invstd", C.c_void_p),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dbias", C.c_void_p),
                ("bwdA", C.c_void_p), ("bwdB", C.c_void_p), ("bwdC", C.c_void_p)]


class CaeEpilogue(C.Structure):
    _fields_ = [("mode", C.c_int), ("bias", C.c_void_p), ("partials", C.c_void_p), ("ticket", C.c_void_p),
                ("bn", CaeBN), ("act", CaeView), ("target", CaeSrc), ("loss_out", C.c_void_p),
                ("dbias", C.c_void_p), ("write_mode", C.c_int), ("count_scale", C.c_float), ("addend", CaeSrc)]


class CaeGemm(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("A", C.c_void_p), ("sAm", C.c_longlong), ("sAk", C.c_longlong),
                ("B", C.c_void_p), ("sBk", C.c_longlong), ("sBn", C.c_longlong),
                ("C", C.c_void_p), ("sCm", C.c_longlong), ("sCn", C.c_longlong),
                ("a_k0", C.c_void_p), ("a_k2", C.c_void_p), ("a_hw", C.c_int), ("a_relu", C.c_int),
                ("b_k0", C.c_void_p), ("b_k2", C.c_void_p), ("b_hw", C.c_int), ("b_relu", C.c_int),
                ("bias", C.c_void_p), ("relu_out", C.c_int), ("mask", C.c_void_p), ("rowsum_A", C.c_void_p)]


class CaePatchHead(C.Structure):
    _fields_ = [("inp", CaeSrc), ("weight", C.c_void_p), ("bias", C.c_void_p), ("K", C.c_int), ("Cout", C.c_int),
                ("target", CaeSrc), ("mask", CaeSrc), ("mask_channels", C.c_int), ("lambda_pearson", C.c_float),
                ("count_scale", C.c_float), ("moments", C.c_void_p), ("coef", C.c_void_p), ("scalars", C.c_void_p),
                ("loss_out", C.c_void_p), ("pearson_out", C.c_void_p), ("ticket", C.c_void_p), ("mse_scale", C.c_void_p)]


class CaeFcStack(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("N", "in1", "fc1", "lat", "fc2", "out4")] + \
               [("A", C.c_void_p), ("a_k0", C.c_void_p), ("a_k2", C.c_void_p), ("a_hw", C.c_int), ("a_relu", C.c_int)] + \
               [(k, C.c_void_p) for k in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4")] + \
               [("bn1", CaeBN), ("bn3", CaeBN), ("train", C.c_int), ("relu_mid", C.c_int)] + \
               [(k, C.c_void_p) for k in ("t1", "z", "t3", "u", "du", "dW1", "db1", "dW2", "db2", "dW3", "db3", "dW4",
                                          "db4", "dA")]


STEM_MAX = 4


class CaeStemConv(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("Cin", "Hin", "Win", "Cout", "Hout", "Wout", "k", "stride", "pad")] + \
               [(k, C.c_void_p) for k in ("w", "b", "scale", "shift")]


class CaeStemFc(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("inp", "out", "relu")] + [(k, C.c_void_p) for k in ("w", "b", "scale", "shift")]


class CaeStemUp(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("Cin", "Hin", "Win", "Cout", "Hout", "Wout", "k", "stride", "pad", "Cr", "skip")] + \
               [(k, C.c_void_p) for k in ("w", "b", "W1", "W2", "scale", "shift")]


class CaeUnetStem(C.Structure):
    _fields_ = [("n_conv", C.c_int), ("n_fc", C.c_int), ("n_up", C.c_int), ("conv", CaeStemConv * STEM_MAX),
                ("fc", CaeStemFc * STEM_MAX), ("up", CaeStemUp * STEM_MAX)]


class CaeTcGemm(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("a_hi", C.c_void_p), ("a_lo", C.c_void_p),
                ("lda", C.c_longlong), ("a_mn_major", C.c_int), ("b_hi", C.c_void_p), ("b_lo", C.c_void_p),
                ("ldb", C.c_longlong), ("b_mn_major", C.c_int), ("C", C.c_void_p), ("ldc", C.c_longlong),
                ("splits", C.c_int), ("split_stride", C.c_longlong), ("tile_n", C.c_int)]


class CaeStemTrainConv(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("Cin", "Hin", "Win", "Cout", "Hout", "Wout", "k", "stride", "pad")] + \
               [(k, C.c_void_p) for k in ("w", "b", "dw", "db")] + [("bn", CaeBN)]


class CaeStemTrainFc(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("inp", "out", "has_bn")] + [(k, C.c_void_p) for k in ("w", "b", "dw", "db")] + \
               [("bn", CaeBN)]


class CaeStemTrainUp(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("Cin", "Hin", "Win", "Cout", "Hout", "Wout", "k", "stride", "pad", "Cr", "skip")] + \
               [(k, C.c_void_p) for k in ("w", "b", "W1", "W2", "dw", "db", "dW1", "dW2")] + [("bn", CaeBN)]


class CaeStemTrain(C.Structure):
    _fields_ = [("n_conv", C.c_int), ("n_fc", C.c_int), ("n_up", C.c_int), ("N", C.c_int),
                ("conv", CaeStemTrainConv * STEM_MAX), ("fc", CaeStemTrainFc * STEM_MAX), ("up", CaeStemTrainUp * STEM_MAX),
                ("params", C.c_void_p), ("params_len", C.c_longlong), ("tape", C.c_void_p), ("hin", C.c_void_p), ("dhin", C.c_void_p), ("dropout_p", C.c_float),
                ("seed", C.c_ulonglong), ("step_count", C.c_void_p), ("bnpart", C.c_void_p), ("wpart", C.c_void_p)]


class CaeDpPeers(C.Structure):
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("grads", C.c_void_p * 8), ("flags", C.c_void_p * 8)]


class CaeTcConv(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("Cin", "Cout", "kh", "kw", "stride", "N", "Hin", "Win", "Hout", "Wout")] + \
               [("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("lda", C.c_longlong), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p),
                ("cols", C.c_void_p), ("cols_len", C.c_longlong), ("dcols_hi", C.c_void_p), ("dcols_lo", C.c_void_p),
                ("ldn", C.c_longlong)]


EPI_PLAIN, EPI_STATS, EPI_MASKSTATS, EPI_SIGMOID, EPI_SIGMOID_MSE, EPI_MASK = 0, 1, 2, 3, 4, 5

# every symbol include/cae_b200.h declares
EXPORTS = {
    "cae_last_error": (C.c_char_p, []),
    "cae_version": (C.c_int, []),
    "cae_struct_size": (C.c_longlong, [C.c_int]),
    "cae_set_kernel_generation": (None, [C.c_int]),
    "cae_partials_len": (C.c_longlong, [C.c_int]),
    "cae_conv_down": (C.c_int, [C.POINTER(CaeSrc), C.c_void_p, C.POINTER(CaeConvGeom), C.POINTER(CaeView),
                                C.POINTER(CaeEpilogue), C.c_void_p]),
    "cae_conv_up": (C.c_int, [C.POINTER(CaeSrc), C.c_void_p, C.POINTER(CaeConvGeom), C.POINTER(CaeView),
                              C.POINTER(CaeEpilogue), C.c_void_p]),
    "cae_conv_wgrad": (C.c_int, [C.POINTER(CaeSrc), C.POINTER(CaeSrc), C.POINTER(CaeConvGeom), C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "cae_wgrad_partials_len": (C.c_longlong, [C.POINTER(CaeSrc), C.POINTER(CaeSrc), C.POINTER(CaeConvGeom)]),
    "cae_ew_epilogue": (C.c_int, [C.POINTER(CaeSrc), C.POINTER(CaeView), C.POINTER(CaeEpilogue), C.c_void_p]),
    "cae_gemm": (C.c_int, [C.POINTER(CaeGemm), C.c_void_p]),
    "cae_gemm_tc_workspace": (C.c_longlong, [C.POINTER(CaeGemm)]),
    "cae_gemm_tc": (C.c_int, [C.POINTER(CaeGemm), C.c_void_p, C.c_longlong, C.c_void_p]),
    "cae_bn_eval_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "cae_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                          C.c_void_p]),
    "cae_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float,
                           C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "cae_step_advance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cae_adam_advance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float,
                                   C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p]),
    "cae_vae_reparam_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "cae_vae_reparam_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "cae_add2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cae_plane_stats": (C.c_int, [C.POINTER(CaeView), C.c_void_p, C.c_void_p]),
    "cae_channel_attention_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p]),
    "cae_channel_attention_bwd": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p] * 5),
    "cae_plane_dot": (C.c_int, [C.POINTER(CaeSrc), C.POINTER(CaeView), C.c_void_p, C.c_void_p]),
    "cae_gate_bwd": (C.c_int, [C.POINTER(CaeSrc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(CaeView),
                               C.c_void_p, C.c_void_p]),
    "cae_sum_over_n": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cae_masked_pearson_loss": (C.c_int, [C.POINTER(CaeView), C.POINTER(CaeSrc), C.POINTER(CaeSrc), C.c_int, C.c_float,
                                          C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(CaeView), C.c_void_p, C.c_void_p, C.c_void_p]),
    "cae_attention_block_supported": (C.c_int, [C.c_int] * 4),
    "cae_attention_block_partials_len": (C.c_longlong, [C.c_int, C.c_int]),
    "cae_attention_block_fwd": (C.c_int, [C.POINTER(CaeView), C.POINTER(CaeSrc), C.c_void_p, C.c_void_p, C.c_int,
                                          C.POINTER(CaeView), C.POINTER(CaeEpilogue), C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "cae_attention_block_bwd": (C.c_int, [C.POINTER(CaeSrc), C.POINTER(CaeView)] + [C.c_void_p] * 5 + [C.c_int] +
                                [C.POINTER(CaeView)] + [C.c_void_p] * 6),
    "cae_fc_stack_supported": (C.c_int, [C.c_int] * 6),
    "cae_fc_stack_fwd": (C.c_int, [C.POINTER(CaeFcStack), C.c_void_p]),
    "cae_fc_stack_bwd": (C.c_int, [C.POINTER(CaeFcStack), C.c_void_p]),
    "cae_unet_stem_supported": (C.c_int, [C.POINTER(CaeUnetStem)]),
    "cae_unet_stem_eval": (C.c_int, [C.POINTER(CaeUnetStem), C.POINTER(CaeSrc), C.POINTER(CaeView), C.c_void_p]),
    "cae_patch_head_supported": (C.c_int, [C.c_int] * 5),
    "cae_patch_head_fwd": (C.c_int, [C.POINTER(CaePatchHead), C.POINTER(CaeView), C.c_void_p]),
    "cae_patch_head_bwd": (C.c_int, [C.POINTER(CaePatchHead), C.POINTER(CaeView), C.POINTER(CaeEpilogue), C.c_void_p,
                                     C.c_void_p]),
    "cae_patch_head_wgrad_reduce": (C.c_int, [C.POINTER(CaePatchHead), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cae_patch_head_partials_len": (C.c_longlong, [C.POINTER(CaePatchHead)]),
    "cae_tc_gemm": (C.c_int, [C.POINTER(CaeTcGemm), C.c_void_p]),
    "cae_tc_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cae_tc_convt_supported": (C.c_int, [C.c_int] * 6),
    "cae_tc_convt_wgrad_splits": (C.c_longlong, [C.POINTER(CaeTcConv)]),
    "cae_tc_convt_fwd": (C.c_int, [C.POINTER(CaeTcConv), C.POINTER(CaeSrc), C.c_void_p, C.POINTER(CaeView),
                                   C.POINTER(CaeEpilogue), C.c_void_p]),
    "cae_tc_convt_im2col": (C.c_int, [C.POINTER(CaeTcConv), C.POINTER(CaeSrc), C.c_void_p]),
    "cae_tc_convt_dgrad": (C.c_int, [C.POINTER(CaeTcConv), C.c_void_p, C.POINTER(CaeView), C.POINTER(CaeEpilogue),
                                     C.c_void_p]),
    "cae_tc_convt_wgrad": (C.c_int, [C.POINTER(CaeTcConv), C.c_void_p, C.c_void_p]),
    "cae_unet_stem_train_supported": (C.c_int, [C.POINTER(CaeStemTrain)]),
    "cae_unet_stem_train_tape_elems": (C.c_longlong, [C.POINTER(CaeStemTrain)]),
    "cae_unet_stem_train_workspace": (C.c_longlong, [C.POINTER(CaeStemTrain), C.c_int]),
    "cae_unet_stem_train_fwd": (C.c_int, [C.POINTER(CaeStemTrain), C.POINTER(CaeSrc), C.c_void_p]),
    "cae_unet_stem_train_profile": (C.c_int, [C.c_void_p]),
    "cae_unet_stem_train_bwd": (C.c_int, [C.POINTER(CaeStemTrain), C.POINTER(CaeSrc), C.c_void_p]),
    "cae_minmax_partials_len": (C.c_longlong, []),
    "cae_minmax": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cae_case_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_double,
                                  C.c_double, C.c_void_p, C.c_void_p]),
    "cae_normalise_gather": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int,
                                       C.c_void_p, C.c_longlong, C.c_void_p]),
    "cae_dp_wait_done": (C.c_int, [C.POINTER(CaeDpPeers), C.c_void_p, C.c_void_p]),
    "cae_adam_allreduce": (C.c_int, [C.c_void_p, C.POINTER(CaeDpPeers), C.c_void_p, C.c_void_p, C.c_longlong, C.c_float,
                                     C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cae_randn": (C.c_int, [C.c_void_p, C.c_longlong, C.c_ulonglong, C.c_void_p, C.c_void_p]),
}

_lib = None


class CaeError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CaeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in EXPORTS.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().cae_last_error().decode("utf-8", "replace")
        raise CaeError(f"{what or 'libcae_b200'} failed (code {rc}): {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise CaeError("cae_tools_b200 needs a CUDA device (sm_100a); there is no CPU fallback for the hot path")
    lib()
