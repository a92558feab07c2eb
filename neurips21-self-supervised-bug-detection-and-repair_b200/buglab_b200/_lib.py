"""ctypes binding of the C ABI declared in ``include/buglab_b200.h``.

The shared library is built in-tree (``csrc/build.sh`` / ``__graft_entry__.build()``) and loaded lazily
by path, so modules that use it stay picklable (no ctypes handle is ever stored on an ``nn.Module``;
reference requirement: ``buglab/data/modelsync/server.py:34`` pickles ``(model, nn)``).

There is NO CPU fallback: if the library is missing, or a tensor is not a contiguous CUDA tensor of the
expected dtype, the call raises.
"""
import ctypes
import os
import threading
from typing import Optional

import torch

_LIB_NAME = "libbuglab_b200.so"
_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()

c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_f32 = ctypes.c_float
c_u64 = ctypes.c_uint64
c_ptr = ctypes.c_void_p
c_size = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/buglab_b200.h one to one.
_SIGNATURES = {
    "bl_version": (c_i32, []),
    "bl_error_string": (ctypes.c_char_p, [c_i32]),
    "bl_plan_workspace_bytes": (c_size, [c_i64, c_i64, c_i32]),
    "bl_plan_build": (c_i32, [c_ptr] * 3 + [c_i64, c_i64, c_i32] + [c_ptr] * 19 + [c_i32, c_ptr, c_size, c_ptr]),
    "bl_rows_gather": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    "bl_rows_segment_sum": (c_i32, [c_ptr] * 6 + [c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    "bl_rows_split3_f16": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr, c_ptr]),
    "bl_unscale_pow2": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr]),
    "bl_weights_split3_f16": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    "bl_pair_project_fwd": (c_i32, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_ptr, c_ptr]),
    "bl_pair_project_bwd_input": (c_i32, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_ptr, c_ptr]),
    "bl_rows_split2_f16": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr, c_ptr]),
    "bl_pair_project_bwd_weight": (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_ptr]),
    "bl_grouped_colsum": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_ptr, c_ptr]),
    "bl_absmax": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr]),
    "bl_weight_parts_f16": (c_i32, [c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    "bl_pair_project_tc_supported": (c_i32, [c_i32, c_i32]),
    "bl_pair_project_tc": (c_i32, [c_ptr] * 6 + [c_i32, c_i64, c_i32, c_i32, c_ptr, c_ptr]),
    "bl_pair_weight_grad_tc_supported": (c_i32, [c_i32, c_i32]),
    "bl_pair_weight_grad_tc": (c_i32, [c_ptr] * 5 + [c_i32, c_i64, c_i32, c_i32, c_ptr, c_i32, c_i32, c_ptr]),
    "bl_rows_split_f16": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr, c_ptr]),
    "bl_segment_unit_prefix": (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_ptr]),
    "bl_segment_units": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "bl_tma_tile_rows": (c_i32, []),
    "bl_tma_slab_rows": (c_i32, []),
    "bl_tma_gemm_supported": (c_i32, [c_i32, c_i32]),
    "bl_tma_project": (c_i32, [c_ptr, c_i64] + [c_ptr] * 7 + [c_i32, c_i64, c_i64, c_i32, c_i32, c_ptr, c_ptr]),
    "bl_tma_project_stationary_supported": (c_i32, [c_i32, c_i32]),
    "bl_tma_project_stationary": (c_i32, [c_ptr, c_i64] + [c_ptr] * 7 + [c_i32, c_i64, c_i64, c_i32, c_i32, c_ptr, c_ptr]),
    "bl_tma_weight_grad_supported": (c_i32, [c_i32, c_i32]),
    "bl_tma_weight_grad": (c_i32, [c_ptr, c_i64, c_ptr, c_i64] + [c_ptr] * 5 + [c_i32, c_i64, c_i64, c_i32, c_i32,
                                   c_ptr, c_i32, c_i32, c_ptr]),
    "bl_edge_segmax_fwd": (c_i32, [c_ptr] * 5 + [c_i64, c_i32] + [c_ptr] * 3 + [c_ptr]),
    "bl_edge_segmax_bwd": (c_i32, [c_ptr] * 6 + [c_i64, c_i32, c_i64, c_i64] + [c_ptr] * 3 + [c_ptr]),
    "bl_edge_bwd_targets": (c_i32, [c_ptr] * 6 + [c_i64, c_i32, c_i32, c_i64] + [c_ptr] * 5 + [c_ptr]),
    "bl_edge_bwd_sources": (c_i32, [c_ptr] * 5 + [c_i64, c_i32, c_ptr, c_ptr, c_ptr]),
    "bl_layernorm_fwd": (c_i32, [c_ptr] * 3 + [c_i64, c_i32, c_f32] + [c_ptr] * 3 + [c_ptr]),
    "bl_layernorm_bwd": (c_i32, [c_ptr] * 5 + [c_i64, c_i32] + [c_ptr] * 4 + [c_ptr]),
    "bl_tanh_dropout_fwd": (c_i32, [c_ptr, c_i64, c_f32, c_u64, c_ptr, c_ptr, c_ptr]),
    "bl_tanh_dropout_bwd": (c_i32, [c_ptr, c_ptr, c_i64, c_f32, c_u64, c_ptr, c_ptr]),
    "bl_segment_minmax": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_i64, c_i32, c_ptr, c_ptr, c_ptr]),
    "bl_segment_minmax_bwd": (c_i32, [c_ptr, c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    "bl_segment_sum": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_i64, c_ptr, c_ptr]),
    "bl_segment_log_softmax_fwd": (c_i32, [c_ptr, c_ptr, c_i64, c_i64, c_f32] + [c_ptr] * 3 + [c_ptr]),
    "bl_segment_log_softmax_bwd": (c_i32, [c_ptr] * 3 + [c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "bl_subtoken_maxpool_fwd": (c_i32, [c_ptr] * 3 + [c_i64, c_i32, c_i32, c_f32, c_u64, c_ptr, c_ptr, c_ptr]),
    "bl_subtoken_maxpool_bwd": (c_i32, [c_ptr] * 3 + [c_i64, c_i32, c_i32, c_f32, c_u64, c_ptr, c_ptr]),
    "bl_grad_sqnorm": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "bl_adam_step": (c_i32, [c_ptr] * 4 + [c_i64, c_f32, c_f32, c_f32, c_f32, c_i64, c_f32, c_ptr, c_f32, c_ptr]),
    "bl_seq_attention_supported": (c_i32, [c_i32]),
    "bl_seq_attention_fwd": (c_i32, [c_ptr] * 9 + [c_i32] * 5 + [c_f32, c_u64] + [c_ptr, c_ptr, c_ptr]),
    "bl_seq_attention_bwd": (c_i32, [c_ptr] * 12 + [c_i32] * 5 + [c_f32, c_u64] + [c_ptr] * 9 + [c_ptr]),
    "bl_seq_attention_tc_supported": (c_i32, [c_i32, c_i32]),
    "bl_seq_softmax_fwd": (c_i32, [c_ptr] * 8 + [c_i32] * 5 + [c_f32, c_u64] + [c_ptr] * 4 + [c_ptr]),
    "bl_seq_softmax_bwd": (c_i32, [c_ptr] * 9 + [c_i32] * 5 + [c_f32, c_u64] + [c_ptr] * 8 + [c_i32, c_ptr]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            path = library_path()
            if not os.path.exists(path):
                raise RuntimeError(
                    f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or csrc/build.sh). buglab_b200 has no CPU fallback."
                )
            lib = ctypes.CDLL(path)
            for name, (restype, argtypes) in _SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


class BuglabB200Error(RuntimeError):
    pass


# hand-written kernels launched per successful C-ABI call (CUB sorts/scans, memsets and cuBLAS GEMMs not counted)
KERNELS_PER_CALL = {
    "bl_plan_build": 19, "bl_rows_gather": 1, "bl_rows_segment_sum": 1, "bl_edge_segmax_fwd": 1,
    "bl_edge_segmax_bwd": 1, "bl_edge_bwd_targets": 1, "bl_edge_bwd_sources": 1, "bl_layernorm_fwd": 1, "bl_layernorm_bwd": 2, "bl_tanh_dropout_fwd": 1,
    "bl_tanh_dropout_bwd": 1, "bl_segment_minmax": 5, "bl_segment_minmax_bwd": 1, "bl_segment_sum": 1,
    "bl_segment_log_softmax_fwd": 5, "bl_segment_log_softmax_bwd": 2, "bl_subtoken_maxpool_fwd": 1,
    "bl_subtoken_maxpool_bwd": 1, "bl_grad_sqnorm": 2, "bl_adam_step": 1,
    "bl_rows_split3_f16": 1, "bl_unscale_pow2": 1, "bl_weights_split3_f16": 2, "bl_pair_project_fwd": 0, "bl_pair_project_bwd_input": 0,
    "bl_pair_project_bwd_weight": 1, "bl_rows_split2_f16": 1, "bl_grouped_colsum": 1, "bl_absmax": 1, "bl_weight_parts_f16": 1,
    "bl_pair_project_tc": 1, "bl_pair_weight_grad_tc": 1, "bl_tma_project": 1, "bl_tma_project_stationary": 1, "bl_tma_weight_grad": 1,
    "bl_segment_unit_prefix": 1, "bl_segment_units": 2, "bl_seq_attention_fwd": 1, "bl_seq_attention_bwd": 2,
    "bl_seq_softmax_fwd": 1, "bl_seq_softmax_bwd": 1,
}
launch_counter = {"kernels": 0, "calls": 0}


def check(code: int, what: str) -> None:
    launch_counter["calls"] += 1
    launch_counter["kernels"] += KERNELS_PER_CALL.get(what, 1)
    if code != 0:
        msg = load().bl_error_string(code)
        raise BuglabB200Error(f"{what} failed with code {code}: {msg.decode() if msg else '?'}")


def ptr(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = None) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise BuglabB200Error("buglab_b200 kernels need CUDA tensors; there is no CPU fallback")
    if not t.is_contiguous():
        raise BuglabB200Error("buglab_b200 kernels need contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise BuglabB200Error(f"expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


def f32(t):
    return ptr(t, torch.float32)


def i32(t):
    return ptr(t, torch.int32)


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream
