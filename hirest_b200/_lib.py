"""ctypes binding of libhirest_b200.so (C ABI declared in include/hirest_b200.h).

There is deliberately NO fallback: if the shared library is missing or the device is not sm_100 every call
raises.  The library is built in-tree by ``build.sh`` / ``__graft_entry__.build()``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

# HIREST_B200_LIB selects another build of the same library (A/B experiments of kernel variants); never a different backend.
_LIB_PATH = os.environ.get("HIREST_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhirest_b200.so")
_lib = None
_lock = threading.Lock()
_inited_device = None

c_float_p = C.POINTER(C.c_float)
c_float_pp = C.POINTER(c_float_p)


class HbVitConfig(C.Structure):
    _fields_ = [("image_size", C.c_int), ("patch_size", C.c_int), ("width", C.c_int), ("layers", C.c_int),
                ("heads", C.c_int), ("mlp_hidden", C.c_int), ("embed_dim", C.c_int), ("ln_eps", C.c_float)]


class HbVitWeights(C.Structure):
    _fields_ = [("cls_token", C.c_void_p), ("pos_embed", C.c_void_p), ("patch_w", C.c_void_p), ("patch_b", C.c_void_p),
                ("norm1_w", C.c_void_p), ("norm1_b", C.c_void_p), ("q_bias", C.c_void_p), ("v_bias", C.c_void_p),
                ("qkv_w", C.c_void_p), ("proj_w", C.c_void_p), ("proj_b", C.c_void_p), ("norm2_w", C.c_void_p),
                ("norm2_b", C.c_void_p), ("fc1_w", C.c_void_p), ("fc1_b", C.c_void_p), ("fc2_w", C.c_void_p),
                ("fc2_b", C.c_void_p), ("norm_w", C.c_void_p), ("norm_b", C.c_void_p), ("head_w", C.c_void_p),
                ("head_b", C.c_void_p)]


class HbTextConfig(C.Structure):
    _fields_ = [("context_length", C.c_int), ("vocab_size", C.c_int), ("width", C.c_int), ("heads", C.c_int),
                ("layers", C.c_int), ("embed_dim", C.c_int), ("ln_eps", C.c_float), ("precise", C.c_int)]


class HbTextWeights(C.Structure):
    _fields_ = [("token_embedding", C.c_void_p), ("positional_embedding", C.c_void_p), ("ln1_w", C.c_void_p),
                ("ln1_b", C.c_void_p), ("in_proj_w", C.c_void_p), ("in_proj_b", C.c_void_p), ("out_proj_w", C.c_void_p),
                ("out_proj_b", C.c_void_p), ("ln2_w", C.c_void_p), ("ln2_b", C.c_void_p), ("fc_w", C.c_void_p),
                ("fc_b", C.c_void_p), ("cproj_w", C.c_void_p), ("cproj_b", C.c_void_p), ("ln_final_w", C.c_void_p),
                ("ln_final_b", C.c_void_p), ("text_projection", C.c_void_p)]


class HbMomentConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("embed_dim", "hidden", "heads", "ffn", "layers", "asr_dim", "clip_dim", "max_pos")]


MOMENT_WEIGHT_FIELDS = ("asr_ln_w", "asr_ln_b", "asr_w", "asr_b", "temp_w1", "temp_b1", "temp_w2", "temp_b2", "mask_embed",
                        "boundary_embed", "head_w", "head_b", "vis_norm_w", "vis_norm_b", "clip_g_map_w", "clip_g_map_b",
                        "clip_g_map_text_w", "clip_g_map_text_b", "emb_w", "emb_b", "pos_emb", "emb_ln_w", "emb_ln_b",
                        "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "ao_w", "ao_b", "ao_ln_w", "ao_ln_b", "i_w", "i_b", "o_w", "o_b",
                        "o_ln_w", "o_ln_b")


class HbMomentWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in MOMENT_WEIGHT_FIELDS]


class HbDecoderConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("hidden", "heads", "ffn", "layers", "vocab", "max_pos", "max_words", "bos", "eos")]


DECODER_WEIGHT_FIELDS = ("word_emb", "pos_emb", "emb_ln_w", "emb_ln_b", "sq_w", "sq_b", "sk_w", "sk_b", "sv_w", "sv_b", "so_w", "so_b",
                         "so_ln_w", "so_ln_b", "eq_w", "eq_b", "ek_w", "ek_b", "ev_w", "ev_b", "eo_w", "eo_b", "eo_ln_w", "eo_ln_b",
                         "i_w", "i_b", "o_w", "o_b", "o_ln_w", "o_ln_b", "cls_dense_w", "cls_dense_b", "cls_ln_w", "cls_ln_b", "cls_bias")


class HbDecoderWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in DECODER_WEIGHT_FIELDS]


class HbProfileSummary(C.Structure):
    _fields_ = [("ms", C.c_double * 6), ("flops", C.c_double * 6), ("launches", C.c_int64 * 6)]


PROF_CATEGORIES = ("gemm_bf16", "gemm_gelu", "gemm_f32", "vit_attention", "layernorm", "other")

# name -> (restype, argtypes); must list every symbol include/hirest_b200.h declares (tests check this).
SIGNATURES = {
    "hb_init": (C.c_int, [C.c_int]),
    "hb_last_error": (C.c_char_p, []),
    "hb_strerror": (C.c_char_p, [C.c_int]),
    "hb_launch_count": (C.c_int64, []),
    "hb_debug_set": (C.c_int, [C.c_char_p, C.c_int]),
    "hb_gemm_n_tiling": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "hb_profile_start": (C.c_int, []),
    "hb_profile_stop": (C.c_int, [C.POINTER(HbProfileSummary)]),
    "hb_vit_create": (C.c_int, [C.POINTER(HbVitConfig), C.POINTER(HbVitWeights), C.c_int, C.c_void_p,
                                C.POINTER(C.c_void_p)]),
    "hb_vit_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "hb_vit_encode_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    "hb_vit_set_tap": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "hb_vit_destroy": (None, [C.c_void_p]),
    "hb_text_create": (C.c_int, [C.POINTER(HbTextConfig), C.POINTER(HbTextWeights), C.c_int, C.c_void_p,
                                 C.POINTER(C.c_void_p)]),
    "hb_text_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "hb_text_destroy": (None, [C.c_void_p]),
    "hb_moment_create": (C.c_int, [C.POINTER(HbMomentConfig), C.POINTER(HbMomentWeights), C.c_int64, C.c_int, C.c_void_p,
                                   C.POINTER(C.c_void_p)]),
    "hb_moment_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_moment_destroy": (None, [C.c_void_p]),
    "hb_moment_mr_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "hb_moment_ms_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_double, C.c_void_p, C.c_void_p]),
    "hb_trim_feats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hb_decoder_create": (C.c_int, [C.POINTER(HbDecoderConfig), C.POINTER(HbDecoderWeights), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.POINTER(C.c_void_p)]),
    "hb_decoder_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hb_decoder_step": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hb_decoder_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_decoder_destroy": (None, [C.c_void_p]),
    "hb_pool_normalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hb_resize_crop_u8": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hb_resize_geometry": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int)]),
    "hb_resize_tables": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_int64]),
    "hb_subsample_pool_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hb_subsample_pool_normalize_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hb_resample_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hb_wordpiece_create": (C.c_int, [C.c_char_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "hb_wordpiece_destroy": (None, [C.c_void_p]),
    "hb_wordpiece_encode_captions": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p]),
    "hb_bpe_create": (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "hb_bpe_destroy": (None, [C.c_void_p]),
    "hb_bpe_tokenize": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "hb_asr_warp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                              C.c_void_p]),
    "hb_similarity": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                C.c_void_p]),
    "hb_linear": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                            C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "hb_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int, C.c_void_p, C.c_int,
                               C.c_void_p]),
    "hb_vit_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "hb_small_attention_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                         C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]),
    "hb_small_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                     C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]),
}


def lib_path() -> str:
    return _LIB_PATH


def load():
    """dlopen the library (no device needed) and attach signatures."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                raise RuntimeError(
                    f"{_LIB_PATH} not found: build it with ./build.sh (nvcc, sm_100a). hirest_b200 has no CPU fallback.")
            lib = C.CDLL(_LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def debug_set(key: str, value: int) -> None:
    """Kernel-variant switch for A/B measurements / cross-check tests (include/hirest_b200_debug.h); not part of the boundary."""
    check(load().hb_debug_set(key.encode(), int(value)), f"hb_debug_set({key})")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        lib = load()
        msg = lib.hb_last_error().decode(errors="replace")
        raise RuntimeError(f"hirest_b200: {what} failed with {rc} ({lib.hb_strerror(rc).decode()}): {msg}")


def init(device: int = 0):
    """hb_init once per process/device. Raises if the device is not a B200-class (sm_100) GPU."""
    global _inited_device
    lib = load()
    if _inited_device != device:
        check(lib.hb_init(int(device)), "hb_init")
        _inited_device = device
    return lib


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def ptr_array(tensors):
    """Host array of device pointers (HbVitWeights per-layer members). Keeps no reference to the tensors."""
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr
