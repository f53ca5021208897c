"""ctypes binding of libb200fe.so (include/b200_fe.h).

Host code stays Python over PyTorch tensors; this module is the only place that crosses into the
native library: raw ``data_ptr()`` pointers, explicit sizes, the current CUDA stream, integer
return codes.  There is NO fallback: if the library is missing or a call fails, a B200Error is
raised - a CPU / eager-PyTorch path silently standing in for the kernels would void every
parity and performance claim made for this repository.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / 'libb200fe.so'
_lib = None

EPI_STORE, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_PARTIAL, EPI_GELU_Q8, EPI_DGELU_Q8 = range(7)
OPT_SGD, OPT_ADAMW = 0, 1

c_ll, c_int, c_float, c_vp = C.c_longlong, C.c_int, C.c_float, C.c_void_p


class B200Error(RuntimeError):
    pass


class OptTensor(C.Structure):
    _fields_ = [('param', c_vp), ('grad', c_vp), ('state1', c_vp), ('state2', c_vp), ('param_bf16', c_vp),
                ('numel', c_ll), ('lr', c_float), ('weight_decay', c_float), ('beta1', c_float), ('beta2', c_float),
                ('eps', c_float), ('step', c_int)]


class AugParams(C.Structure):
    _fields_ = [('sharpen', c_int), ('autocontrast', c_int), ('crop_y', c_int), ('crop_x', c_int), ('rot', c_int * 6)]


# name -> (restype, argtypes); mirrors include/b200_fe.h line by line
_PROTOS = {
    'b200_last_error': (C.c_char_p, []),
    'b200_device_check': (c_int, []),
    'b200_prof_begin': (c_int, [c_int]),
    'b200_prof_end': (c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(c_ll), C.POINTER(c_ll)]),
    'b200_prof_gemm_bytes': (c_int, [C.POINTER(C.c_double)]),
    'b200_prof_kernels': (c_int, [c_int, c_vp, c_vp, c_vp, c_vp]),
    'b200_gemm_tn': (c_int, [c_vp, c_ll, c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_vp, c_ll, c_int, c_vp, c_ll, c_vp,
                             c_vp, c_ll, c_int, c_ll, c_int, c_vp]),
    'b200_gemm_wgrad': (c_int, [c_vp, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_vp, c_int, c_int, c_vp]),
    'b200_gemm_wgrad_bias': (c_int, [c_vp, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_vp, c_vp, c_int, c_int, C.POINTER(c_int), c_vp]),
    'b200_gemm_splits': (c_int, [c_int, c_int]),
    'b200_splitk_reduce': (c_int, [c_vp, c_vp, c_ll, c_int, c_int, c_vp]),
    'b200_reduce_defer_begin': (c_int, []),
    'b200_reduce_pending': (c_int, []),
    'b200_reduce_flush': (c_int, [c_vp, c_int]),
    'b200_layernorm_fwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_float, c_vp]),
    'b200_layernorm_bwd_blocks': (c_int, [c_ll, c_int]),
    'b200_layernorm_bwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_vp]),
    'b200_layernorm_fwd_windows': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_float, c_vp]),
    'b200_layernorm_bwd_windows': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b200_window_rows': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b200_patch_gather_image': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_ll, c_vp]),
    'b200_patch_gather_image_u8': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_ll, c_vp]),
    'b200_patch_gather_nhwc': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b200_augment_train': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'b200_gemm_taps': (c_int, [c_vp, c_ll, c_vp, c_ll, c_int, c_int, c_int, c_int, C.POINTER(c_int), c_int, c_vp, c_ll, c_vp, c_ll, c_vp]),
    'b200_gemm_wgrad_taps': (c_int, [c_vp, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_int, C.POINTER(c_int), c_vp, c_int, c_vp]),
    'b200_gemm_conv_bn': (c_int, [c_vp, c_ll, c_vp, c_ll, c_int, c_int, c_int, c_int, C.POINTER(c_int), c_vp, c_int, c_int, c_int, c_vp, c_ll,
                                  c_vp, c_ll, c_vp]),
    'b200_bn_stats_blocks': (c_int, [c_ll]),
    'b200_bn_stats': (c_int, [c_vp, c_ll, c_int, c_int, c_int, C.c_double, c_vp, c_vp, c_vp, c_vp, c_float, c_float, c_vp, c_vp, c_vp]),
    'b200_bn_apply': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_vp, c_vp]),
    'b200_bn_backward': (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, C.c_double, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'b200_stem_im2col': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'b200_stem_pool_fwd': (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'b200_stem_pool_bwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'b200_grid_sample2': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'b200_grid_patches_s2': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'b200_grid_avgpool': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'b200_mean_pool': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
    'b200_transpose16': (c_int, [c_vp, c_vp, c_ll, c_int, c_ll, c_ll, c_vp]),
    'b200_cast_transpose': (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    'b200_cast_f32_bf16': (c_int, [c_vp, c_vp, c_ll, c_vp]),
    'b200_colsum_blocks': (c_int, [c_ll]),
    'b200_colsum': (c_int, [c_vp, c_ll, c_ll, c_int, c_vp, c_vp, c_int, c_vp]),
    'b200_window_attn_fwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'b200_window_attn_bwd_blocks': (c_int, [c_int, c_int, c_int, c_int]),
    'b200_window_attn_bwd_scratch_floats': (c_ll, [c_int]),
    'b200_window_attn_bwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_vp]),
    'b200_unit_rows': (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_ll, c_float, c_int, c_vp]),
    'b200_margin_logits': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_float, c_float, c_int, c_int, c_vp, c_ll, c_vp, c_vp]),
    'b200_margin_ce': (c_int, [c_vp, c_ll, c_vp, c_vp, c_int, c_int, c_float, c_float, c_int, c_int, c_float, c_vp, c_vp, c_vp,
                               c_ll, c_vp, c_vp, c_vp]),
    'b200_focal_loss': (c_int, [c_vp, c_ll, c_vp, c_int, c_int, c_float, c_vp, c_vp, c_vp, c_ll, c_vp]),
    'b200_unit_rows_bwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_vp]),
    'b200_opt_chunk_elems': (c_int, []),
    'b200_optimizer_step': (c_int, [c_int, c_vp, c_vp, c_int, c_float, c_int, c_vp]),
    'b200_optimizer_step_sum': (c_int, [c_int, c_vp, c_vp, c_int, c_float, c_int, c_int, c_ll, c_ll, c_vp]),
    'b200_peer_alloc': (c_int, [c_ll, C.POINTER(c_vp), C.c_char_p]),
    'b200_peer_open': (c_int, [C.c_char_p, C.POINTER(c_vp)]),
    'b200_peer_close': (c_int, [c_vp]),
    'b200_peer_free': (c_int, [c_vp]),
    'b200_peer_copy': (c_int, [c_vp, c_vp, c_ll, c_vp]),
    'b200_swin_create': (c_vp, [c_int, c_int, c_int, c_int, C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), c_int, c_int,
                                c_int, c_int]),
    'b200_swin_destroy': (None, [c_vp]),
    'b200_swin_param_elems': (c_ll, [c_vp]),
    'b200_swin_param_count': (c_int, [c_vp]),
    'b200_swin_param_offsets': (c_int, [c_vp, C.POINTER(c_ll), C.POINTER(c_ll), c_int]),
    'b200_swin_wcache_bytes': (c_ll, [c_vp]),
    'b200_swin_workspace_bytes': (c_ll, [c_vp]),
    'b200_swin_sync_weights': (c_int, [c_vp, c_vp, c_vp, c_vp]),
    'b200_swin_forward': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_ll, c_vp]),
    'b200_swin_backward': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_vp]),
    'b200_gallery_prepare': (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_vp]),
    'b200_gallery_prepare_ex': (c_int, [c_vp, c_vp, c_int, c_float, c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_vp]),
    'b200_gallery_frame': (c_int, [c_vp, c_int, c_vp, c_vp]),
    'b200_unit_row_mean_blocks': (c_int, [c_ll]),
    'b200_unit_row_mean': (c_int, [c_vp, c_ll, c_int, c_vp, c_vp, c_vp]),
    'b200_cosine_topk_workspace_bytes': (c_ll, [c_ll, c_ll, c_int, c_int]),
    'b200_cosine_topk': (c_int, [c_vp, c_vp, c_vp, c_ll, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_ll, c_ll, c_vp, c_vp, c_vp, c_ll,
                                 c_vp]),
    'b200_cosine_topk_certified': (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_vp, c_vp, c_vp, c_vp, c_float, c_vp, c_ll, c_int, c_int, c_ll, c_ll,
                                           c_vp, c_vp, c_vp, c_vp, c_ll, c_vp]),
    'b200_topk_merge': (c_int, [c_vp, c_vp, c_ll, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'b200_pair_similarity': (c_int, [c_vp, c_ll, c_int, c_vp, c_vp, c_ll, c_vp, c_vp]),
    'b200_recall_hits': (c_int, [c_vp, c_ll, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
}


_PENDING = set()
NO_EXCLUDE = -(1 << 62)


def lib_path() -> Path:
    return _LIB_PATH


def exported_names():
    return list(_PROTOS)


def lib():
    """Load the native library (once).  Raises B200Error if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise B200Error(f'{_LIB_PATH} is missing - run `python pets-face-recognition_b200/build.py` '
                            f'(or __graft_entry__.build()); there is no CPU fallback for this path')
        handle = C.CDLL(str(_LIB_PATH), mode=os.RTLD_LOCAL if hasattr(os, 'RTLD_LOCAL') else 0)
        for name, (res, args) in _PROTOS.items():
            if name in _PENDING:
                continue
            fn = getattr(handle, name)       # AttributeError here == header/library mismatch: fail loudly
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().b200_last_error().decode(errors='replace')


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        raise B200Error(f'{what or "b200 call"} failed ({rc}): {last_error()}')


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    """data_ptr of a CUDA tensor (None -> NULL)."""
    if t is None:
        return 0
    if not t.is_cuda:
        raise B200Error('b200 kernels take CUDA tensors only (no CPU fallback)')
    return t.data_ptr()


def require_device() -> None:
    if not torch.cuda.is_available():
        raise B200Error('no CUDA device: the B200 path has no CPU fallback')
    check(lib().b200_device_check(), 'device_check')
