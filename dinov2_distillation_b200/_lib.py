"""ctypes binding of libb200distill.so (see include/b200_distill.h). No torch types cross this boundary: callers pass
raw device pointers (tensor.data_ptr()), sizes and the current cudaStream_t.

There is deliberately NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libb200distill.so"

c_ll = C.c_longlong
c_vp = C.c_void_p
c_fp = C.c_void_p  # float* passed as an integer address


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", c_vp), ("lda", c_ll), ("a_mn_major", C.c_int),
        ("B", c_vp), ("ldb", c_ll), ("b_mn_major", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("split_k", C.c_int),
        ("bias", c_fp),
        ("act", C.c_int),
        ("aux", c_vp), ("ldaux", c_ll), ("aux_mode", C.c_int),
        ("col_scale", c_fp),
        ("residual", c_fp), ("ldres", c_ll), ("res_row_period", C.c_int),
        ("out_f32", c_fp), ("ldo32", c_ll), ("atomic_add", C.c_int),
        ("out_bf16", c_vp), ("ldo16", c_ll),
        ("out_bf16_pre", c_vp), ("ldo16_pre", c_ll),
        ("out_row_period", C.c_int), ("out_row_pad", C.c_int),
        ("a_is_fp16", C.c_int), ("b_is_fp16", C.c_int), ("out16_is_fp16", C.c_int), ("aux_is_fp16", C.c_int),
        ("algo_flops_scale", C.c_float),
        ("out_batch_period", C.c_int), ("out_batch_stride", c_ll),
        ("out16_pre_alt", C.c_int),
        ("out16_colsum", c_fp),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", c_vp), ("q_bs", c_ll), ("q_ts", c_ll),
        ("k", c_vp), ("k_bs", c_ll), ("k_ts", c_ll),
        ("v", c_vp), ("v_bs", c_ll), ("v_ts", c_ll),
        ("o", c_vp), ("o_bs", c_ll), ("o_ts", c_ll),
        ("lse", c_fp),
        ("B", C.c_int), ("heads", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("hd", C.c_int),
        ("scale", C.c_float),
        ("d_o", c_vp), ("do_bs", c_ll), ("do_ts", c_ll),
        ("delta", c_fp),
        ("dq", c_vp), ("dq_bs", c_ll), ("dq_ts", c_ll),
        ("dk", c_vp), ("dk_bs", c_ll), ("dk_ts", c_ll),
        ("dv", c_vp), ("dv_bs", c_ll), ("dv_ts", c_ll),
        ("qkvo_is_fp16", C.c_int),
        ("dq_colsum", c_fp), ("dk_colsum", c_fp), ("dv_colsum", c_fp),
        ("o_alt", c_vp),
        ("q_alt", c_vp), ("k_alt", c_vp), ("v_alt", c_vp),
        ("dq_accum", c_fp),
    ]


_VIT_BLOCK_FIELDS = [
    "ln1_w", "ln1_b", "ln2_w", "ln2_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ls1", "ls2",
    "fc1_w", "fc1_b", "fc2_w", "fc2_b", "qkv_wT", "proj_wT", "fc1_wT", "fc2_wT",
]


class VitBlock(C.Structure):
    _fields_ = [(n, c_vp) for n in _VIT_BLOCK_FIELDS]


class VitConfig(C.Structure):
    _fields_ = [("D", C.c_int), ("L", C.c_int), ("heads", C.c_int), ("F", C.c_int), ("swiglu", C.c_int),
                ("ln_eps", C.c_float)]


PROJ_PARAM_FIELDS = [
    "conv_w", "conv_b", "bn_w", "bn_b", "bn_running_mean", "bn_running_var", "pos_embed",
    "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "p_w", "p_b",
    "ffn1_w", "ffn1_b", "ffn2_w", "ffn2_b", "ln1_w", "ln1_b", "ln2_w", "ln2_b", "query_w",
    "bn_num_batches_tracked",
]
PROJ_GRAD_FIELDS = [
    "conv_w", "conv_b", "bn_w", "bn_b", "pos_embed",
    "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "p_w", "p_b",
    "ffn1_w", "ffn1_b", "ffn2_w", "ffn2_b", "ln1_w", "ln1_b", "ln2_w", "ln2_b", "query_w",
]


class ProjectorParams(C.Structure):
    _fields_ = [(n, c_vp) for n in PROJ_PARAM_FIELDS]


class ProjectorGrads(C.Structure):
    _fields_ = [(n, c_vp) for n in PROJ_GRAD_FIELDS]


class ProjectorConfig(C.Structure):
    _fields_ = [("Cs", C.c_int), ("D", C.c_int), ("HW", C.c_int), ("heads", C.c_int),
                ("softmax_scale", C.c_float), ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
                ("ln_eps", C.c_float), ("training", C.c_int),
                ("raw_h", C.c_int), ("raw_w", C.c_int), ("grid_h", C.c_int), ("grid_w", C.c_int),
                ("win_h", C.c_int), ("win_w", C.c_int)]


# name -> (restype, argtypes); every symbol include/b200_distill.h declares
_i, _f, _sz = C.c_int, C.c_float, C.c_size_t
SIGNATURES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_abi_version": (_i, []),
    "b200_launch_count": (c_ll, []),
    "b200_reset_launch_count": (None, []),
    "b200_profile_enable": (None, [_i]),
    "b200_set_option": (_i, [C.c_char_p, _i]),
    "b200_get_option": (_i, [C.c_char_p]),
    "b200_profile_read": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(c_ll)]),
    "b200_gemm_bf16": (_i, [C.POINTER(GemmDesc), c_vp]),
    "b200_ln_gemm_bf16": (_i, [c_fp, c_fp, c_fp, C.c_float, c_fp, c_fp, c_vp, C.POINTER(GemmDesc), c_vp]),
    "b200_cast_f32_bf16": (_i, [c_fp, c_vp, c_ll, c_vp]),
    "b200_split3_16": (_i, [c_fp, c_vp, c_ll, _i, _i, _i, c_vp]),
    "b200_cast_f32_f16": (_i, [c_fp, c_vp, c_ll, c_vp]),
    "b200_cast_f16_bf16": (_i, [c_vp, c_vp, c_ll, c_vp]),
    "b200_transpose_f32_bf16": (_i, [c_fp, c_vp, _i, _i, c_fp, c_vp]),
    "b200_transpose_f32_bf16_ld": (_i, [c_fp, c_vp, _i, _i, c_ll, c_fp, c_vp]),
    "b200_nchw_to_tokens": (_i, [c_fp, c_vp, c_fp, _i, _i, _i, _i, c_vp]),
    "b200_tokens_to_nchw": (_i, [c_fp, c_fp, _i, _i, _i, _i, c_vp]),
    "b200_window_rows16": (_i, [c_vp, c_vp, c_ll, _i, _i, _i, _i, _i, c_ll, _i, c_vp]),
    "b200_bilinear_tokens_fwd": (_i, [c_fp, c_fp, _i, _i, _i, _i, _i, _i, c_vp]),
    "b200_bilinear_tokens_bwd": (_i, [c_vp, c_vp, _i, _i, _i, _i, _i, _i, c_vp]),
    "b200_patch_im2col": (_i, [c_fp, c_vp, _i, _i, _i, _i, c_vp]),
    "b200_profile_event_overhead": (_i, [_i, C.POINTER(C.c_float), C.POINTER(C.c_float), c_vp]),
    "b200_profile_read_records": (_i, [_i, c_vp, c_vp, c_vp, c_vp]),
    "b200_augment_ws_bytes": (_sz, [_i, _i, _i]),
    "b200_augment_batch": (_i, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_fp, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), c_vp, _sz, c_vp]),
    "b200_feature_cache_store": (_i, [c_fp, C.c_longlong, C.c_longlong, c_vp, c_vp, _i, _i, _i, _i, c_vp]),
    "b200_feature_cache_load": (_i, [c_vp, _i, c_vp, c_fp, _i, _i, _i, c_vp]),
    "b200_write_cls_rows": (_i, [c_fp, c_fp, c_fp, _i, _i, _i, c_vp]),
    "b200_layernorm_fwd": (_i, [c_fp, c_fp, c_fp, _f, c_fp, c_vp, c_fp, c_fp, _i, _i, _i, _i, _i, c_vp]),
    "b200_layernorm_bwd": (_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_vp, c_fp, c_fp, _i, _i, c_vp]),
    "b200_layernorm_bwd_colsum": (_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_vp, c_fp, c_fp, c_fp, _i, _i, c_vp]),
    "b200_bn_stats": (_i, [c_fp, c_fp, _i, _i, c_vp]),
    "b200_bn_finalize": (_i, [c_fp, c_fp, c_fp, c_fp, c_fp, _f, _f, _i, _i, c_vp]),
    "b200_bn_relu_pos_fwd": (_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_vp, _i, _i, _i, _i, c_vp]),
    "b200_bn_relu_pos_bwd_reduce": (_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, _i, _i, _i, c_vp]),
    "b200_bn_relu_pos_bwd_apply": (_i, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_vp, _i, _i, _i, c_vp]),
    "b200_colsum": (_i, [c_vp, _i, c_ll, c_fp, _i, _i, c_vp]),
    "b200_batch_sum": (_i, [c_fp, c_fp, _i, c_ll, c_vp]),
    "b200_batch_sum_bf16": (_i, [c_vp, c_fp, c_vp, _i, c_ll, c_vp]),
    "b200_axpy": (_i, [c_fp, c_fp, _f, c_ll, c_vp]),
    "b200_swiglu": (_i, [c_vp, c_vp, _i, _i, c_vp]),
    "b200_swiglu_bwd": (_i, [c_vp, c_vp, c_vp, _i, _i, c_vp]),
    "b200_attention_fwd": (_i, [C.POINTER(AttnDesc), c_vp]),
    "b200_attention_bwd": (_i, [C.POINTER(AttnDesc), c_vp]),
    "b200_kd_loss_ws_floats": (c_ll, [_i, _i, _i]),
    "b200_kd_loss_fwd": (_i, [c_fp, c_fp, _i, _i, _i, _i, _i, _i, _f, c_fp, c_fp, c_vp]),
    "b200_kd_loss_bwd": (_i, [c_fp, c_fp, _i, _i, _i, _i, _i, _i, _f, c_fp, c_fp, _i, c_fp, c_vp]),
    "b200_kd_loss_bwd_split": (_i, [c_fp, c_fp, _i, _i, _i, _i, _i, _i, _f, c_fp, c_fp, c_fp, _i, c_fp, c_vp]),
    "b200_dct_zero_dc_idct": (_i, [c_fp, c_fp, _i, _i, _i, c_ll, c_ll, c_vp]),
    "b200_vit_forward_ws_bytes": (_sz, [C.POINTER(VitConfig), _i, _i, _i]),
    "b200_vit_forward": (_i, [C.POINTER(VitConfig), C.POINTER(VitBlock), c_vp, _i, c_fp, c_fp, c_fp, c_fp, c_fp,
                              c_fp, _i, _i, _i, c_fp, c_vp, _sz, c_vp]),
    "b200_vit_block_ws_bytes": (_sz, [C.POINTER(VitConfig), _i, _i]),
    "b200_vit_block_save_bytes": (_sz, [C.POINTER(VitConfig), _i, _i]),
    "b200_vit_block_fwd": (_i, [C.POINTER(VitConfig), C.POINTER(VitBlock), c_fp, c_fp, _i, _i, c_vp, c_vp, _sz, c_vp]),
    "b200_vit_block_bwd_input": (_i, [C.POINTER(VitConfig), C.POINTER(VitBlock), c_fp, c_fp, c_fp, _i, _i, c_vp, c_vp,
                                      _sz, c_vp]),
    "b200_projector_ws_bytes": (_sz, [C.POINTER(ProjectorConfig), _i]),
    "b200_projector_save_bytes": (_sz, [C.POINTER(ProjectorConfig), _i]),
    "b200_projector_fwd": (_i, [C.POINTER(ProjectorConfig), C.POINTER(ProjectorParams), c_fp, c_fp, _i, c_fp, c_vp,
                                c_vp, _sz, c_vp]),
    "b200_projector_bwd": (_i, [C.POINTER(ProjectorConfig), C.POINTER(ProjectorParams), C.POINTER(ProjectorGrads),
                                c_fp, c_fp, c_fp, _i, c_fp, _i, c_fp, c_vp, c_vp, _sz, c_vp]),
    "b200_sqnorm_f32": (_i, [c_fp, c_ll, c_fp, c_vp]),
    "b200_adamw_step": (_i, [c_fp, c_fp, c_fp, c_fp, c_ll, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _i, c_fp,
                            c_fp, C.c_float, c_vp]),
    "b200_projector_tokens_bytes": (_sz, [C.POINTER(ProjectorConfig), _i]),
    "b200_projector_tokenize": (_i, [C.POINTER(ProjectorConfig), c_fp, _i, c_vp, c_vp, _sz, c_vp]),
    "b200_projector_fwd_tok": (_i, [C.POINTER(ProjectorConfig), C.POINTER(ProjectorParams), c_fp, c_fp, _i, c_fp, c_vp,
                                    c_vp, _sz, c_vp, c_vp]),
    "b200_projector_bwd_tok": (_i, [C.POINTER(ProjectorConfig), C.POINTER(ProjectorParams), C.POINTER(ProjectorGrads),
                                    c_fp, c_fp, c_fp, _i, c_fp, _i, c_fp, c_vp, c_vp, _sz, c_vp, c_vp]),
}

ABI_VERSION = 3   # include/b200_distill.h: B200_ABI_VERSION

_lib = None


class B200Error(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree library (building it with nvcc when absent). Raises when impossible -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise B200Error(f"{LIB_PATH} is missing; run `python -m dinov2_distillation_b200.build`")
        from . import build as _build
        _build.build(verbose=False)
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
        fn.restype = res
        fn.argtypes = args
    if lib.b200_abi_version() != ABI_VERSION:
        raise B200Error("libb200distill.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().b200_last_error().decode(errors="replace")
        raise B200Error(f"{what or 'b200 call'} failed ({rc}): {msg}")
