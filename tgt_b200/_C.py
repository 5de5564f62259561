"""ctypes binding of libtgt_b200.so (the C ABI declared in include/tgt_b200.h).

There is deliberately no fallback: if the shared library is missing the first kernel call
raises, and every non-zero return code becomes a RuntimeError carrying tgt_last_error().
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtgt_b200.so")

F32, BF16, F16 = 0, 1, 2
_DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPE_CODE[dt]
    except KeyError:
        raise RuntimeError(f"tgt_b200: unsupported dtype {dt}") from None


class TripletAttnDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("d", C.c_int32),
                ("ld", C.c_int64),
                ("off_q", C.c_int32 * 2), ("off_k", C.c_int32 * 2), ("off_v", C.c_int32 * 2),
                ("off_e", C.c_int32 * 2), ("off_g", C.c_int32 * 2),
                ("scale", C.c_float), ("dtype", C.c_int32)]


class TripletAggrDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("d", C.c_int32),
                ("ld", C.c_int64),
                ("off_v", C.c_int32 * 2), ("off_e", C.c_int32 * 2), ("off_g", C.c_int32 * 2),
                ("mask_dir", C.c_int32 * 2), ("dtype", C.c_int32)]


class TriangularDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("ld", C.c_int64),
                ("off_v", C.c_int32), ("off_e", C.c_int32), ("dtype", C.c_int32)]


class EgtDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("d", C.c_int32),
                ("ld_qkv", C.c_int64), ("ld_eg", C.c_int64),
                ("scale", C.c_float), ("attend", C.c_int32), ("scale_degree", C.c_int32),
                ("dtype", C.c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int32), ("K", C.c_int32),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64),
                ("dtype", C.c_int32), ("flags", C.c_int32),
                ("row_mean", C.c_void_p), ("row_rstd", C.c_void_p), ("col_sum", C.c_void_p),
                ("bias", C.c_void_p), ("row_scale", C.c_void_p),
                ("res", C.c_void_p), ("ldres", C.c_int64), ("res_dtype", C.c_int32),
                ("rows_per_scale", C.c_int32),
                ("U", C.c_void_p), ("ldu", C.c_int64),
                ("p_drop", C.c_float), ("seed", C.c_uint64),
                ("stat_mean", C.c_void_p), ("stat_rstd", C.c_void_p), ("stat_eps", C.c_float),
                ("res2", C.c_void_p), ("ldres2", C.c_int64), ("stat_partial", C.c_void_p),
                ("seed_ptr", C.c_void_p)]


EPI_LN, EPI_BIAS, EPI_GELU, EPI_RES, EPI_STORE_U, EPI_ROWSCALE, EPI_GELU_BWD, EPI_STATS, EPI_LN_BWD = 1, 2, 4, 8, 16, 64, 128, 256, 512
EPI_STORE_GP, EPI_MULRES = 1024, 2048

_P = C.c_void_p
_SIGNATURES = {
    "tgt_version": (C.c_int, []),
    "tgt_last_error": (C.c_char_p, []),
    "tgt_launch_count": (C.c_uint64, []),
    "tgt_set_kernel_policy": (None, [C.c_int]),
    "tgt_kernel_timer_enable": (None, [C.c_int]),
    "tgt_kernel_timer_read": (C.c_int, [C.c_char_p, C.c_size_t]),
    "tgt_layernorm_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int64, C.c_float, C.c_int, C.c_int, _P]),
    "tgt_layernorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _P]),
    "tgt_layernorm_bwd_y": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      _P, C.c_int64, _P, _P]),
    "tgt_triplet_attn_workspace_bytes": (C.c_size_t, [C.POINTER(TripletAttnDesc), C.c_int]),
    "tgt_triplet_attn_fwd": (C.c_int, [C.POINTER(TripletAttnDesc), _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "tgt_triplet_attn_fwd_f32out": (C.c_int, [C.POINTER(TripletAttnDesc), _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "tgt_triplet_attn_bwd": (C.c_int, [C.POINTER(TripletAttnDesc), _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "tgt_triplet_attn_bwd_tiles": (C.c_int, [C.POINTER(TripletAttnDesc), _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P, _P]),
    "tgt_triplet_attn_bwd_bias_supported": (C.c_int, [C.POINTER(TripletAttnDesc)]),
    "tgt_triplet_attn_bwd_bias": (C.c_int, [C.POINTER(TripletAttnDesc), _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P, _P, _P]),
    "tgt_triplet_attn_fused_supported": (C.c_int, [C.POINTER(TripletAttnDesc), C.c_int]),
    "tgt_triplet_attn_fused_fwd": (C.c_int, [C.POINTER(TripletAttnDesc), _P, C.c_int64, C.c_int, _P, _P, _P, _P, _P, _P, _P,
                                             _P, _P, _P, C.c_size_t, _P]),
    "tgt_triplet_aggr_fwd": (C.c_int, [C.POINTER(TripletAggrDesc), _P, _P, _P, _P, _P]),
    "tgt_triplet_aggr_bwd": (C.c_int, [C.POINTER(TripletAggrDesc), _P, _P, _P, _P, _P, _P, _P]),
    "tgt_triangular_fwd": (C.c_int, [C.POINTER(TriangularDesc), _P, _P, _P, _P]),
    "tgt_triangular_bwd": (C.c_int, [C.POINTER(TriangularDesc), _P, _P, _P, _P, _P]),
    "tgt_siglin_fwd": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "tgt_siglin_bwd": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "tgt_egt_attn_fwd": (C.c_int, [C.POINTER(EgtDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "tgt_egt_attn_workspace_bytes": (C.c_size_t, [C.POINTER(EgtDesc)]),
    "tgt_egt_attn_bwd": (C.c_int, [C.POINTER(EgtDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "tgt_gelu_dropout_fwd": (C.c_int, [_P, _P, C.c_int64, C.c_float, C.c_uint64, C.c_int, _P]),
    "tgt_gelu_dropout_fwd_dseed": (C.c_int, [_P, _P, C.c_int64, C.c_float, _P, C.c_int, _P]),
    "tgt_gelu_dropout_bwd": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, C.c_uint64, C.c_int, _P]),
    "tgt_gemm_tc": (C.c_int, [C.POINTER(GemmDesc), _P, _P, _P, _P]),
    "tgt_gemm_tc_slices": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "tgt_row_stats": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_int64, C.c_float, C.c_int, _P]),
    "tgt_gaussian_basis_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "tgt_gaussian_basis_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "tgt_bins_decode": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _P]),
    "tgt_xent_rows_fwd": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "tgt_xent_rows_bwd": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, _P]),
    "tgt_scaled_residual": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"tgt_b200: CUDA library {LIB_PATH} is missing -- run "
                        "`python -c 'import __graft_entry__ as g; g.build()'` (there is no fallback path)")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"tgt_b200.{what} failed: {lib().tgt_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(lib().tgt_launch_count())


def kernel_timer(enable: bool) -> None:
    lib().tgt_kernel_timer_enable(1 if enable else 0)


def kernel_timer_read() -> dict:
    """{kernel name: (launches, total_ms)} measured on the device around each main kernel since kernel_timer(True)."""
    buf = C.create_string_buffer(1 << 16)
    lib().tgt_kernel_timer_read(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split()
        out[name] = (int(n), float(ms))
    return out


_policy = 0


def set_kernel_policy(policy: int) -> None:
    """0 = fastest supported kernel (default), 1 = generic SIMT kernels, 2 = cp.async-staged tensor-core triplet kernels,
    3 = fused projection + attention forward, 4 = tcgen05 / TMEM triplet core, 5 = TMA-staged mma.sync core
    (see include/tgt_b200.h)."""
    global _policy
    _policy = int(policy)
    lib().tgt_set_kernel_policy(int(policy))


def kernel_policy() -> int:
    return _policy
