"""ctypes binding of ``csrc/libou_b200.so`` (C ABI declared in ``include/ou_b200.h``).

Error codes are turned into Python exceptions (``ValueError`` for OU_ERR_INVALID,
``NotImplementedError`` for OU_ERR_UNSUPPORTED, ``RuntimeError`` otherwise).  There is no
fallback: if the library is missing it is built with nvcc, and if that is impossible every
device call raises.
"""
import ctypes
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_void_p)
from pathlib import Path

from ..build import LIB as LIB_PATH
OU_ABI_VERSION = 4


class ConvParams(Structure):
    _fields_ = [
        ("x", c_void_p), ("w", c_void_p), ("w_tc", c_void_p), ("bias", c_void_p), ("add1", c_void_p),
        ("add2", c_void_p), ("gamma", c_void_p), ("beta", c_void_p), ("out", c_void_p),
        ("out_f32_blk", c_void_p),
        ("batch", c_int32), ("cin", c_int32), ("t_in", c_int32),
        ("s", c_int32), ("taps", c_int32), ("tap_off", c_int32),
        ("n", c_int32), ("cout", c_int32), ("up", c_int32),
        ("kpad", c_int32), ("npad", c_int32),
        ("rows", c_int32), ("t_out", c_int32),
        ("film_bstride", c_int32),
        ("has_prelu_in", c_int32), ("has_prelu_out", c_int32), ("has_prelu_out2", c_int32),
        ("prelu_in", c_float), ("prelu_out", c_float), ("prelu_out2", c_float),
        ("scale1", c_float), ("scale2", c_float),
        ("max_ctas", c_int32),
    ]


class TrunkParams(Structure):
    _fields_ = [
        ("x", c_void_p), ("w1", c_void_p), ("w2", c_void_p), ("w3", c_void_p),
        ("b1", c_void_p), ("b2", c_void_p), ("b3", c_void_p), ("sc", c_void_p),
        ("gamma", c_void_p), ("beta", c_void_p), ("out", c_void_p),
        ("batch", c_int32), ("channels", c_int32), ("t", c_int32),
        ("taps1", c_int32), ("taps2", c_int32), ("taps3", c_int32),
        ("film_bstride", c_int32),
        ("has_prelu_out", c_int32), ("has_prelu_out2", c_int32),
        ("prelu_in", c_float), ("prelu_mid1", c_float), ("prelu_mid2", c_float),
        ("prelu_out", c_float), ("prelu_out2", c_float),
        ("scale1", c_float), ("scale3", c_float),
        ("max_ctas", c_int32),
        ("up_w", c_void_p), ("up_bias", c_void_p), ("up_skip", c_void_p), ("up_out", c_void_p),
        ("up_t_out", c_int32), ("up_scale", c_float), ("up_prelu_in", c_float),
        ("out_w", c_void_p), ("out_bias", c_float), ("out_coef", c_void_p), ("out_x", c_void_p),
        ("out_noise", c_void_p), ("out_xout", c_void_p), ("out_net", c_void_p),
        ("dn_w", c_void_p), ("dn_bias", c_void_p), ("dn_out", c_void_p), ("dn_t_out", c_int32),
        ("dn_prelu_in", c_float), ("up_taps", c_int32), ("dn_taps", c_int32),
    ]


class StepArgs(Structure):
    _fields_ = [("x", c_void_p), ("film", c_void_p), ("film_bstride", c_int32), ("in_scale", c_void_p),
                ("coef", c_void_p), ("noise", c_void_p), ("xout", c_void_p), ("net_out", c_void_p),
                ("x_wav", c_void_p)]


# name -> (restype, argtypes); every symbol of include/ou_b200.h
SIGNATURES = {
    "ou_abi_version": (c_int, []),
    "ou_last_error": (c_int, [c_char_p, c_size_t]),
    "ou_launch_count": (c_int64, []),
    "ou_act_dtype": (c_int, []),
    "ou_conv_fallback_count": (c_int64, []),
    "ou_conv1d": (c_int, [POINTER(ConvParams), c_void_p]),
    "ou_conv1d_naive": (c_int, [POINTER(ConvParams), c_void_p]),
    "ou_conv_trunk": (c_int, [POINTER(TrunkParams), c_void_p]),
    "ou_input_conv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                              c_int, c_int, c_void_p]),
    "ou_output_sde": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ou_gru_bidir": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int,
                             c_int, c_int, c_void_p]),
    "ou_gru_bidir_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int,
                                c_int, c_int, c_int, c_void_p]),
    "ou_gru_ctas": (c_int, [c_int, c_int, c_int]),
    "ou_mel_power": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ou_mel_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_void_p]),
    "ou_sigma_embed_simple": (c_int, [c_void_p, c_float, c_float, c_void_p, c_int, c_int, c_void_p]),
    "ou_sigma_embed_rff": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ou_linear_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                              c_int, c_float, c_void_p]),
    "ou_pad_normalize": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                 c_float, c_void_p]),
    "ou_unpad_limit": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_void_p]),
    "ou_pack_blocked": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ou_unpack_blocked": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ou_film_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ou_alias_free_snake": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                    c_void_p, c_int, c_void_p, c_float, c_int, c_void_p, c_int,
                                    c_int, c_int, c_void_p]),
    "ou_resample_poly": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_void_p]),
    "ou_lsd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                       c_int, c_int, c_float, c_int, c_float, c_float, c_int, c_void_p]),
    "ou_plan_create": (c_int, [POINTER(c_void_p)]),
    "ou_plan_destroy": (c_int, [c_void_p]),
    "ou_plan_size": (c_int, [c_void_p]),
    "ou_plan_add_conv": (c_int, [c_void_p, POINTER(ConvParams), c_int32]),
    "ou_plan_add_trunk": (c_int, [c_void_p, POINTER(TrunkParams), c_int32]),   # out tail: fed from ou_step_args
    "ou_plan_add_input_conv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_int]),
    "ou_plan_add_output_sde": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int,
                                       c_int]),
    "ou_plan_add_mel": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                c_int, c_int, c_int, c_int, c_int, c_int]),
    "ou_plan_add_gru": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int,
                                c_int, c_int, c_int]),
    "ou_plan_run": (c_int, [c_void_p, POINTER(StepArgs), c_int, c_int, c_void_p]),
    "ou_debug_set_trace": (c_int, [c_void_p]),
}

_lib = None


class OuError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Load (building first if necessary) the shared library and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from ..build import build, needs_build
        if needs_build():
            build()
    if not LIB_PATH.exists():
        raise OuError(f"{LIB_PATH} is missing: build it with `python -m open_universe_b200.build` "
                      "(there is no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ou_abi_version() != OU_ABI_VERSION:
        raise OuError(f"ABI mismatch: library {lib.ou_abi_version()} vs binding {OU_ABI_VERSION}")
    _lib = lib
    return lib


def last_error():
    buf = ctypes.create_string_buffer(512)
    load().ou_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc):
    if rc == 0:
        return
    msg = last_error()
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise NotImplementedError(msg)
    raise OuError(f"libou_b200 error {rc}: {msg}")


def launch_count():
    return int(load().ou_launch_count())


def conv_fallback_count():
    """ou_conv1d calls served by the mma.sync kernel because the tcgen05 kernel rejected the geometry."""
    return int(load().ou_conv_fallback_count())


def act_dtype():
    """torch dtype of blocked activations and packed conv weights in this build of the library."""
    import torch
    return torch.bfloat16 if load().ou_act_dtype() == 1 else torch.float16


def act_name():
    return "bf16" if load().ou_act_dtype() == 1 else "f16"
