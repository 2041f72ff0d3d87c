"""ctypes binding of the C ABI declared in include/gww.h.

There is deliberately no fallback: if libgww_b200.so is missing or a call fails, a RuntimeError is
raised.  PyTorch is used only for device memory, streams and torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
# GWW_OPERAND=bf16 selects the bf16-operand build (default: fp16 operands, see include/gww.h gww_operand_dtype);
# GWW_LIB overrides the library path (A/B builds of kernel variants for tuning runs)
_VARIANT = "_bf16" if os.environ.get("GWW_OPERAND", "").lower() == "bf16" else ""
LIB_PATH = os.environ.get("GWW_LIB") or os.path.join(HERE, f"libgww_b200{_VARIANT}.so")

GWW_MAX_HEAD_LAYERS = 6
c_float_p = C.POINTER(C.c_float)


class EncoderConfig(C.Structure):
    _fields_ = [("d_model", C.c_int), ("n_layers", C.c_int), ("n_heads", C.c_int), ("ffn_dim", C.c_int)]


class Dora(C.Structure):
    _fields_ = [("lora_A", c_float_p), ("lora_B", c_float_p), ("magnitude", c_float_p),
                ("r", C.c_int), ("scale", C.c_float)]


class LayerWeights(C.Structure):
    _fields_ = [(n, c_float_p) for n in (
        "ln1_g", "ln1_b", "q_w", "q_b", "k_w", "v_w", "v_b", "o_w", "o_b", "ln2_g", "ln2_b",
        "fc1_w", "fc1_b", "fc2_w", "fc2_b")] + [
        ("dora_q", Dora), ("dora_k", Dora), ("dora_v", Dora), ("dora_o", Dora)]


class EncoderWeights(C.Structure):
    _fields_ = [(n, c_float_p) for n in (
        "conv1_w", "conv1_b", "conv2_w", "conv2_b", "pos_emb", "ln_post_g", "ln_post_b")] + [
        ("layers", C.POINTER(LayerWeights))]


class HeadWeights(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("dims", C.c_int * (GWW_MAX_HEAD_LAYERS + 1)),
                ("w", c_float_p * GWW_MAX_HEAD_LAYERS), ("b", c_float_p * GWW_MAX_HEAD_LAYERS),
                ("softmax", C.c_int)]


class QAdapterWeights(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("conv1_w", "conv1_b", "conv2_w", "conv2_b", "conv3_w", "conv3_b",
                                          "conv4_w", "conv4_b")] + [
        ("scale", C.c_float), ("bias", C.c_float), ("n_detectors", C.c_int),
        ("film_gamma", c_float_p), ("film_beta", c_float_p), ("c1", C.c_int), ("c2", C.c_int), ("c3", C.c_int)]


# every symbol include/gww.h declares: (restype, argtypes)
_vp, _l, _i, _f, _sz = C.c_void_p, C.c_long, C.c_int, C.c_float, C.c_size_t
SYMBOLS = {
    "gww_last_error": (C.c_char_p, []),
    "gww_version": (C.c_char_p, []),
    "gww_operand_dtype": (C.c_char_p, []),
    "gww_device_ok": (_i, []),
    "gww_model_create": (_i, [C.POINTER(EncoderConfig), C.POINTER(EncoderWeights), C.POINTER(_vp)]),
    "gww_model_set_head": (_i, [_vp, C.POINTER(HeadWeights)]),
    "gww_model_destroy": (None, [_vp]),
    "gww_model_ln_fold_state": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "gww_workspace_bytes": (_sz, [_vp, _i]),
    "gww_logmel_frontend": (_i, [_vp, _l, _vp, _vp]),
    "gww_resample_16k": (_i, [_vp, _l, _vp, _vp]),
    "gww_logmel_from_16k": (_i, [_vp, _l, _vp, _vp]),
    "gww_encoder_forward": (_i, [_vp, _vp, _l, _vp, _vp, _i, _vp, _sz, _i, _vp]),
    "gww_head_forward": (_i, [_vp, _vp, _l, _vp, _vp]),
    "gww_forward_windows_logmel": (_i, [_vp, _vp, _l, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "gww_stream_search_logmel": (_i, [_vp, _vp, _i, _l, _i, _l, _l, _f, _vp, _vp, _vp, _vp, _i, _vp, _sz, _i, _vp]),
    "gww_threshold_compact": (_i, [_vp, _i, _l, _f, _l, _vp, _vp, _vp, _i, _vp]),
    "gww_qfront_create": (_i, [C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _i, _i, C.POINTER(_vp)]),
    "gww_qfront_set_adapter": (_i, [_vp, C.POINTER(QAdapterWeights)]),
    "gww_qfront_destroy": (None, [_vp]),
    "gww_qfront_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "gww_qfront_plan": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gww_qfront_workspace_bytes": (_sz, [_vp, _l]),
    "gww_qscan": (_i, [_vp, _vp, _l, _l, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gww_qadapter": (_i, [_vp, _vp, _l, _i, _vp, _vp, _sz, _vp]),
    "gww_forward_windows_qscan": (_i, [_vp, _vp, _vp, _l, _i, _i, _vp, _vp, _sz, _vp, _sz, _vp]),
    "gww_stream_search_qscan": (_i, [_vp, _vp, _vp, _i, _l, _i, _l, _l, _i, _f, _vp, _vp, _vp, _vp, _i,
                                     _vp, _sz, _vp, _sz, _vp]),
    "gww_whiten_workspace_bytes": (_sz, [_l, _i, _i, _i, _i]),
    "gww_whiten": (_i, [_vp, _l, C.c_double, _i, _i, _i, C.c_double, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gww_gemm_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _l, _i, _i, _i, _i, _vp]),
    "gww_attention": (_i, [_vp, _vp, _l, _i, _i, _vp]),
    "gww_layernorm": (_i, [_vp, _vp, _vp, _vp, _l, _i, _i, _vp]),
    "gww_set_last_layer_pruning": (_i, [_i]),
    "gww_launch_count": (_l, []),
    "gww_profile_num_kinds": (_i, []),
    "gww_profile_kind_name": (C.c_char_p, [_i]),
    "gww_profile_begin": (_i, []),
    "gww_profile_end": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_long)]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads libgww_b200.so (building nothing: use gw_whisper_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Run `python -m gw_whisper_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def operand_dtype():
    """torch dtype of the library's 16-bit tensor-core operands (torch.float16 or torch.bfloat16)."""
    import torch
    return torch.bfloat16 if load().gww_operand_dtype() == b"bf16" else torch.float16


def check(rc: int) -> None:
    if rc != 0:
        msg = load().gww_last_error().decode()
        raise RuntimeError(f"gww error {rc}: {msg}")


def ptr(t) -> Optional[int]:
    """Raw device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
