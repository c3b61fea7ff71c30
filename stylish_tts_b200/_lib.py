"""ctypes binding of libstylish_b200.so (C ABI in include/stylish_b200.h).

The library is built in-tree by ``python -m stylish_tts_b200.csrc.build``
(``__graft_entry__.build()``).  There is deliberately NO fallback: if the
library is missing, or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libstylish_b200.so")

ACT_NONE, ACT_RELU, ACT_LEAKY02, ACT_SNAKE, ACT_SWISH, ACT_GELU = 0, 1, 2, 3, 4, 5

_f32p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f32 = C.c_float


class ConvArgs(C.Structure):
    """Mirror of ``sty_conv1d_args``."""
    _fields_ = [
        ("x", _f32p), ("x_bs", _i64), ("x_cs", _i64),
        ("w", _f32p), ("w_bs", _i64),
        ("bias", _f32p),
        ("y", _f32p), ("y_bs", _i64), ("y_cs", _i64),
        ("res", _f32p), ("r_bs", _i64), ("r_cs", _i64),
        ("in_scale", _f32p), ("in_shift", _f32p), ("in_alpha", _f32p),
        ("in_mask", _f32p), ("out_mask", _f32p), ("out_alpha", _f32p),
        ("out_sumsq", _f32p),
        ("B", _i32), ("CI", _i32), ("CO", _i32), ("T", _i32), ("K", _i32), ("dil", _i32),
        ("pad", _i32),
        ("in_act", _i32), ("out_act", _i32), ("shuffle", _i32),
        ("out_scale", _f32), ("res_scale", _f32),
        ("w_split", _f32p),
        ("dw_w", _f32p), ("dw_b", _f32p), ("dw_gb", _f32p), ("dw_gb_bs", _i64), ("dw_eps", _f32),
        ("out_sum", _f32p),
    ]


class WgradArgs(C.Structure):
    """Mirror of ``sty_conv1d_wgrad_args``."""
    _fields_ = [
        ("x", _f32p), ("x_bs", _i64), ("x_cs", _i64),
        ("dy", _f32p), ("dy_bs", _i64), ("dy_cs", _i64),
        ("in_scale", _f32p), ("in_shift", _f32p), ("in_alpha", _f32p), ("in_mask", _f32p),
        ("out_mask", _f32p), ("dw", _f32p),
        ("B", _i32), ("CI", _i32), ("CO", _i32), ("T", _i32), ("K", _i32), ("dil", _i32),
        ("pad", _i32), ("in_act", _i32), ("out_scale", _f32), ("tensor_cores", _i32),
    ]


class Dropout(C.Structure):
    """Mirror of ``sty_dropout``."""
    _fields_ = [("seed", _f32p), ("site", C.c_uint32), ("p", _f32)]


# name -> argtypes (restype is always int unless listed in _SPECIAL)
_SIGNATURES = {
    "sty_embed_fwd": [_f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32, _f32p],
    "sty_sequence_mask_fwd": [_f32p, _f32p, _i32, _i32, _f32p],
    "sty_conv1d_fwd": [C.POINTER(ConvArgs), _f32p],
    "sty_dwconv1d_fwd": [_f32p, _i64, _i64, _f32p, _f32p, _f32p, _f32p, _f32p, _i64, _i64, _i32,
                         _i32, _i32, _i32, _i32, _i32, _f32p],
    "sty_dwconv_ln_fwd": [_f32p, _i64, _f32p, _f32p, _f32p, _i64, _f32p, _i64, _i32, _i32, _i32,
                          _f32, _f32p],
    "sty_chan_layernorm_fwd": [_f32p, _f32p, _i64, _f32p, _f32p, _i64, _i32, _f32p, _i64, _f32p,
                               _i32, _i32, _i32, _f32, _i32, _f32p],
    "sty_convnext_fused_fwd": [_f32p, _i64, _i64, _f32p, _i64, _i64, _f32p, _f32p, _f32p, _i64, _f32, _f32p, _f32p,
                               _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_chan_layernorm_pitched_fwd": [_f32p, _f32p, _i64, _i64, _f32p, _f32p, _i64, _i32, _f32p, _i64, _i64,
                                       _f32p, _i32, _i32, _i32, _f32, _i32, _f32p],
    "sty_instnorm_affine_fwd": [_f32p, _i64, _i64, _f32p, _i64, _f32p, _f32p, _i32, _i32, _i32,
                                _f32, _f32p],
    "sty_moments_affine_fwd": [_f32p, _f32p, _f32p, _i64, _f32p, _f32p, _i32, _i32, _i32, _f32, _f32p],
    "sty_linear_rows_fwd": [_f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_grn_scale_fwd": [_f32p, _f32p, _f32p, _i32, _i32, _f32p],
    "sty_rope_table": [_f32p, _f32p, _i32, _i32, _f32, _f32p],
    "sty_attention_fwd": [_f32p, _f32p, _f32p, _i64, _f32p, _i64, _f32p, _f32p, _f32p, _i32, _i32,
                          _i32, _i32, _i32, _f32, _f32p],
    "sty_attention_generic_fwd": [_f32p, _i64, _f32p, _f32p, _i64, _f32p, _i64, _f32p, _f32p, _f32p,
                                  _i32, _i32, _i32, _i32, _i32, _f32, _f32p],
    "sty_duration_head_fwd": [_f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_soft_duration_fwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_alignment_fwd": [_f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_scale_mask_fwd": [_f32p, _f32p, _f32p, _i32, _i32, _i32, _f32, _f32p],
    "sty_broadcast_rows_fwd": [_f32p, _f32p, _i64, _i32, _i32, _i32, _f32p],
    "sty_bmm_fwd": [_f32p, _i64, _f32p, _i64, _f32p, _i64, _i32, _i32, _i32, _i32, _f32p],
    "sty_glu_fwd": [_f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_source_fwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32,
                       _f32, _f32, _f32, _f32, _f32p],
    "sty_stft_fwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _i32, _f32p],
    "sty_istft_head_fwd": [_f32p, _i64, _f32p, _f32p, _i64, _f32p, _f32p, _f32p, _i32, _i32, _i32,
                           _i32, _i32, _f32p],
    "sty_stft_pitched_fwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _f32p],
    "sty_istft_head_pitched_fwd": [_f32p, _i64, _i64, _f32p, _f32p, _i64, _i64, _f32p, _f32p, _f32p, _i32, _i32,
                                   _i32, _i32, _i32, _f32p],
    # spectral front-end + STFT losses (struct mirror: spectral.SpecArgs)
    "sty_spectrogram_fwd": [_f32p, _f32p],
    "sty_spectrogram_bwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _i64, _f32p],
    "sty_mel_energy_fwd": [_f32p, _f32p, _i32, _i32, _i32, _f32, _f32, _f32p],
    "sty_l1_sums_fwd": [_f32p, _f32p, _i64, _f32p, _f32p],
    "sty_l1_sums_bwd": [_f32p, _f32p, _i64, _f32p, _f32p, _f32p],
    "sty_phase_loss_fwd": [_f32p, _f32p, _i32, _i32, _i32, _f32p, _f32p],
    "sty_phase_loss_bwd": [_f32p, _f32p, _i32, _i32, _i32, _f32p, _f32p, _f32p],
    # training path
    "sty_attention_lse_fwd": [_f32p, _f32p, _f32p, _i64, _f32p, _i64, _f32p, _f32p, _f32p, _i32, _i32,
                              _i32, _i32, _i32, _f32, _f32p, _f32p],
    "sty_attention_bwd": [_f32p, _f32p, _f32p, _i64, _f32p, _f32p, _i64, _f32p, _f32p, _f32p, _f32p, _i32,
                          _f32p, _f32p, _f32p, _i64, _f32p, _i32, _i32, _i32, _i32, _f32, _f32p],
    "sty_conv1d_wgrad": [_f32p, _f32p],
    "sty_channel_sum": [_f32p, _i64, _i64, _f32p, _f32p, _i32, _i32, _i32, _f32, _f32p],
    "sty_row_dot": [_f32p, _f32p, _f32p, _i64, _i32, _f32p],
    "sty_row_moments": [_f32p, _i64, _i64, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_prologue_bwd_reduce": [_f32p, _f32p, _i64, _i64, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i32,
                                _i32, _i32, _i32, _f32p],
    "sty_prologue_bwd_apply": [_f32p, _f32p, _i64, _i64, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
                               _i64, _i64, _f32p, _i64, _i64, _i32, _i32, _i32, _i32, _f32p],
    "sty_grn_snake_bwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_grn_act_bwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_heads_to_rows": [_f32p, _i64, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _f32p],
    "sty_attn_probs": [_f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_softmax_bwd": [_f32p, _f32p, _i64, _i32, _f32p],
    "sty_bmm_tn_fwd": [_f32p, _i64, _f32p, _i64, _f32p, _i64, _i32, _i32, _i32, _i32, _f32p],
    "sty_chan_layernorm_bwd": [_f32p, _f32p, _i64, _f32p, _f32p, _i64, _i32, _f32p, _f32p, _f32p, _f32p,
                               _i64, _i32, _i32, _i32, _f32, _i32, _f32p],
    "sty_dwconv1d_bwd": [_f32p, _f32p, _i64, _i64, _f32p, _f32p, _i64, _i64, _f32p, _f32p, _i32, _i32, _i32,
                         _i32, _i32, _f32p],
    "sty_glu_bwd": [_f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_unshuffle": [_f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_embed_bwd": [_f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32, _f32p],
    "sty_bmm_nt_fwd": [_f32p, _i64, _f32p, _i64, _f32p, _i64, _i32, _i32, _i32, _i32, _f32p],
    "sty_linear_rows_bwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_istft_head_bwd": [_f32p, _f32p, _f32p, _i64, _f32p, _f32p, _i64, _f32p, _f32p, _f32p, _f32p, _f32p,
                           _i64, _i32, _i32, _i32, _i32, _i32, _f32p],
    "sty_adamw_step": [_f32p, _f32p, _f32p, _f32p, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _f32, _f32p],
    # mel style encoder (row-channel images)
    "sty_fold_rows": [_f32p, _f32p, _i32, _i32, _i32, _i32, _i32, _f32p],
    "sty_dwconv3x3s2_fwd": [_f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_dwconv3x3s2_bwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_avgpool2_fwd": [_f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_avgpool2_bwd": [_f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_region_mean_fwd": [_f32p, _f32p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32p],
    "sty_region_mean_bwd": [_f32p, _f32p, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32p],
    "sty_adamw_step_dev": [_f32p, _f32p, _f32p, _f32p, _i64, _f32p, _f32, _f32, _f32, _f32, _f32, _f32p],
    "sty_dropout_fwd": [_f32p, _f32p, _f32p, _i64, _i64, _i32, _f32, C.POINTER(Dropout), _f32p],
    "sty_dropout_bwd": [_f32p, _f32p, _f32p, _i64, _i64, _i32, _f32, C.POINTER(Dropout), _f32p],
    "sty_affine_act_dropout_fwd": [_f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, C.POINTER(Dropout), _f32p],
    "sty_attention_drop_fwd": [_f32p, _f32p, _f32p, _i64, _f32p, _i64, _f32p, _f32p, _f32p, _i32, _i32,
                               _i32, _i32, _i32, _f32, _f32p, C.POINTER(Dropout), _f32p],
    "sty_attention_drop_bwd": [_f32p, _f32p, _f32p, _i64, _f32p, _f32p, _i64, _f32p, _f32p, _f32p, _f32p, _i32,
                               _f32p, _f32p, _f32p, _i64, _f32p, _i32, _i32, _i32, _i32, _f32,
                               C.POINTER(Dropout), _f32p],
    "sty_stft_loss_finalize": [_f32p, _f32p, _f32p, _i32, _f32, _f32, _i32, _f32p, _f32p],
    # style-diffusion denoiser (token-major TMA-fed GEMMs)
    "sty_split_planes_fwd": [_f32p, _f32p, _i64, _f32p],
    "sty_gemm_split_fwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_build_tokens_fwd": [_f32p, _f32p, _f32, _f32p, _i32, _i32, _i32, _i32, _i32, _f32p],
    "sty_row_ln_split_fwd": [_f32p, _f32p, _f32p, _f32p, _f32, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_token_mean_fwd": [_f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_attention_tokens_fwd": [_f32p, _i64, _f32p, _i64, _i32, _i32, _i32, _f32, _f32p],
    # adversarial losses (spectrogram discriminators, LSGAN / TPRLS)
    "sty_leaky_s2d_fwd": [_f32p, _f32p, _i64, _i32, _i32, _i32, _f32, _f32p],
    "sty_leaky_s2d_bwd": [_f32p, _f32p, _f32p, _i64, _i32, _i32, _i32, _f32, _f32p],
    "sty_row_scale_fwd": [_f32p, _f32p, _f32p, _i64, _i32, _f32, _f32p],
    "sty_segment_sum_fwd": [_f32p, _f32p, _i64, _i32, _f32, _f32p],
    "sty_sqdiff_sum_fwd": [_f32p, _i64, _f32, _f32p, _f32p],
    "sty_tprls_fwd": [_f32p, _f32p, _i64, _f32p, _f32p, _f32p, _f32p],
    "sty_tprls_bwd": [_f32p, _f32p, _i64, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p],
    "sty_attention64_fwd": [_f32p, _f32p, _f32p, _i64, _f32p, _i64, _i32, _i32, _i32, _f32, _f32p, _f32p, _f32p],
    "sty_attention64_bwd": [_f32p, _f32p, _f32p, _i64, _f32p, _f32p, _i64, _f32p, _f32p, _f32p, _f32p, _i64, _i32, _i32,
                            _i32, _f32, _f32p, _f32p],
    "sty_attention64_tokens_fwd": [_f32p, _i64, _f32p, _i64, _i32, _i32, _i32, _f32, _f32p, _f32p],
    "sty_disc_first_fwd": [_f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_disc_first_dgrad": [_f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_disc_first_wgrad": [_f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
    "sty_disc_tail_fwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_disc_tail_bwd": [_f32p, _f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _i32, _f32p],
    "sty_disc_score_wgrad": [_f32p, _f32p, _f32p, _f32p, _i32, _i32, _i32, _f32p],
}
_SPECIAL = {
    "sty_version": ([], C.c_int),
    "sty_last_error": ([], C.c_char_p),
    "sty_device_sm_count": ([], C.c_int),
    "sty_tprls_workspace_bytes": ([], C.c_int64),
    "sty_attention64_workspace_bytes": ([_i32, _i32, _i32], C.c_int64),
    "sty_attention64_bwd_workspace_bytes": ([_i32, _i32, _i32], C.c_int64),
}
EXPORTED = tuple(_SIGNATURES) + tuple(_SPECIAL)

_lib: Optional[C.CDLL] = None
launches = 0  # number of kernel-launching C-ABI calls made (diagnostics / bench)
param_epoch = 0  # bumped by optim.FlatAdamW.step: parameters changed in place, repack cached weights


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"stylish_tts_b200: CUDA library not built ({LIB_PATH} missing). Run "
                "`python -m stylish_tts_b200.csrc.build`; there is no CPU/PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, argt in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argt
            fn.restype = C.c_int
        for name, (argt, rest) in _SPECIAL.items():
            fn = getattr(lib, name)
            fn.argtypes = argt
            fn.restype = rest
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sty_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"stylish_tts_b200.{what} failed (code {rc}): {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# Host-logic dry run (tests only): with DRY_RUN set AND `call` replaced by a recorder, the modules accept CPU tensors
# so that shapes / strides / call order / the autograd wiring can be exercised without a GPU — no arithmetic happens.
# The product never sets it: a missing library or a CPU tensor fails loudly.
DRY_RUN = False


def require_cuda(obj, what: str) -> None:
    """obj: tensor or torch.device"""
    dev = obj.device if isinstance(obj, torch.Tensor) else obj
    if dev.type != "cuda" and not DRY_RUN:
        raise RuntimeError(f"stylish_tts_b200: {what} must live on a CUDA device; there is no CPU fallback")


def _req(t: torch.Tensor, name: str, dtype=torch.float32) -> None:
    if not t.is_cuda and not DRY_RUN:
        raise RuntimeError(f"stylish_tts_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"stylish_tts_b200: `{name}` must be {dtype}, got {t.dtype}")


def _bct(t: torch.Tensor, name: str):
    """(B,C,T) view with unit time stride -> (batch stride, channel stride)."""
    _req(t, name)
    if t.dim() != 3 or (t.shape[2] > 1 and t.stride(2) != 1):
        raise ValueError(f"stylish_tts_b200: `{name}` must be (B,C,T) with contiguous T")
    return t.stride(0), t.stride(1)


profile_log = None  # set to a list to time every call with CUDA events (bench.py roofline leg)


def _signature(name: str, args) -> str:
    if name == "sty_conv1d_wgrad":
        a = args[0]._obj
        return f"conv1d_wgrad[ci={a.CI},co={a.CO},k={a.K},d={a.dil},B={a.B},T={a.T}{'+umma' if a.tensor_cores else ''}]"
    if name == "sty_conv1d_fwd":
        a = args[0]._obj
        extra = ""
        if a.in_act or a.in_scale:
            extra += "+pro"
        if a.out_sumsq:
            extra += "+ssq"
        if a.shuffle > 1:
            extra += f"+shuf{a.shuffle}"
        if a.w_split:
            extra += "+umma"
        if a.dw_w:
            extra += "+dwln"
        return f"conv1d[ci={a.CI},co={a.CO},k={a.K},d={a.dil},B={a.B},T={a.T}{extra}]"
    pos = {"sty_dwconv_ln_fwd": (9, 10), "sty_chan_layernorm_fwd": (11, 12),
           "sty_chan_layernorm_pitched_fwd": (13, 14), "sty_convnext_fused_fwd": (20, 22),
           "sty_instnorm_affine_fwd": (8, 9), "sty_attention_fwd": (12, 13)}.get(name)
    if name == "sty_attention64_fwd":
        return f"attention64_fwd[c=64,T={args[8]}]"
    if pos:
        return f"{name[4:]}[c={args[pos[0]]},T={args[pos[1]]}]"
    return name[4:]


def call(name: str, *args) -> None:
    global launches
    lib = load()
    # kernels per call: source = phase + wave; fused ConvNeXt block = pass 1 + GRN scale + pass 2
    launches += (2 if name in ("sty_source_fwd", "sty_attention64_fwd", "sty_attention64_tokens_fwd")
                 else 3 if name == "sty_attention64_bwd"
                 else 3 if name == "sty_convnext_fused_fwd" else 1)
    if profile_log is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        info = None
        if name == "sty_conv1d_fwd":
            a = args[0]._obj
            info = dict(B=a.B, CI=a.CI, CO=a.CO, K=a.K, T=a.T, res=bool(a.res))
        elif name == "sty_convnext_fused_fwd":
            info = dict(kind="convnext_fused", B=args[19], C=args[20], J=args[21], T=args[22])
        profile_log.append((_signature(name, args), e0, e1, info))
        check(rc, name)
        return
    check(getattr(lib, name)(*args), name)
