"""ctypes binding of libb200whisper.so (include/b200w.h).  There is no fallback: if the library is
missing or CUDA is unavailable the product raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200W_LIB: another build of the same library (A/B timing of kernel variants on one GPU box)
LIB_PATH = os.environ.get("B200W_LIB") or os.path.join(_HERE, "lib", "libb200whisper.so")

_lib = None


class B200WError(RuntimeError):
    pass


class ModelConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_mels", "d_model", "n_heads", "n_enc_layers", "n_dec_layers", "n_vocab",
        "n_audio_ctx", "n_text_ctx", "max_batch", "max_beam")]


class DecodeOpts(C.Structure):
    _fields_ = [
        ("beam_size", C.c_int32), ("patience", C.c_float), ("length_penalty", C.c_float),
        ("max_new_tokens", C.c_int32), ("suppress_blank", C.c_int32),
        ("without_timestamps", C.c_int32), ("max_initial_timestamp_index", C.c_int32),
        ("eot", C.c_int32), ("sot", C.c_int32), ("no_speech", C.c_int32),
        ("no_timestamps", C.c_int32), ("timestamp_begin", C.c_int32),
        ("suppress_ids", C.POINTER(C.c_int32)), ("n_suppress", C.c_int32),
        ("blank_ids", C.POINTER(C.c_int32)), ("n_blank", C.c_int32),
        ("temperature", C.c_float), ("best_of", C.c_int32), ("seed", C.c_uint32),
    ]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_pi32 = C.POINTER(C.c_int32)
_pf = C.POINTER(C.c_float)

# name -> (restype, argtypes).  Must list every symbol include/b200w.h declares.
SIGNATURES = {
    "b200w_last_error": (C.c_char_p, []),
    "b200w_abi_version": (_i, []),
    "b200w_launch_count": (C.c_uint64, []),
    "b200w_gemm_set_trace": (_i, [_vp]),
    "b200w_op_profile_enable": (_i, [_vp, _i]),
    "b200w_op_profile_read": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "b200w_step_probe_reset": (_i, [_vp]),
    "b200w_step_probe_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "b200w_gemm_profile_begin": (_i, [_i]),
    "b200w_gemm_profile_end": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "b200w_attn_profile_begin": (_i, []),
    "b200w_attn_profile_end": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "b200w_logmel": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "b200w_pack_windows": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b200w_dedup_features": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "b200w_bandpass_filtfilt": (_i, [_vp, _vp, _i, C.POINTER(C.c_double), _vp, _vp, _vp]),
    "b200w_vad_window_energy": (_i, [_vp, _i, _f, _vp, _vp]),
    "b200w_vad_gate": (_i, [_vp, _vp, _i, _vp, _vp, _f, _i64, _vp, _vp]),
    "b200w_gemm_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _i, _i, _vp]),
    "b200w_skinny_gemm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i, _i, _vp]),
    "b200w_step_self_attention": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "b200w_sample_topk": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _f, C.c_uint32, _vp]),
    "b200w_cross_attention_tc_scratch_bytes": (C.c_int64, []),
    "b200w_cross_attention_tc": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _f, _vp, _i, _vp]),
    "b200w_cross_attention_flash": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "b200w_conv1d_k3": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "b200w_enc_attention": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "b200w_enc_attention_trace_read": (_i, [_vp, _i]),
    "b200w_engine_create": (_i, [C.POINTER(ModelConfig), C.POINTER(_vp)]),
    "b200w_engine_destroy": (_i, [_vp]),
    "b200w_engine_set_weight": (_i, [_vp, C.c_char_p, _vp, _i64, _i]),
    "b200w_engine_finalize": (_i, [_vp]),
    "b200w_encode": (_i, [_vp, _vp, _i, _vp, _vp]),
    "b200w_cross_kv": (_i, [_vp, _vp, _i, _vp, _vp]),
    "b200w_decode": (_i, [_vp, _vp, _i, _pi32, _pi32, C.POINTER(DecodeOpts), _pi32, C.c_int32, _pi32,
                          _pf, _pf, _vp]),
    "b200w_align": (_i, [_vp, _vp, _i, _pi32, _pi32, _pi32, C.c_int32, _pi32, C.c_int32, C.c_int32,
                         _pi32, C.c_int32, _pi32, _pf, C.c_int32, _vp]),
    "b200w_detect_language": (_i, [_vp, _vp, _i, C.c_int32, _pi32, C.c_int32, _pf, _vp]),
    "b200w_decoder_logits": (_i, [_vp, _vp, _i, _pi32, _pi32, _vp, _vp]),
}


def load(path: str | None = None):
    """Load the shared library (once) and attach prototypes.  Raises B200WError if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise B200WError(
            f"{path} is missing: build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.b200w_abi_version() != 3:
        raise B200WError("libb200whisper.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200w_last_error()
        raise B200WError(msg.decode() if msg else "libb200whisper call failed")


def launch_count() -> int:
    return int(load().b200w_launch_count())
