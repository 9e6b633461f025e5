"""ctypes binding of libqgemm_sm100.so (include/qgemm.h).  Fails loudly when the
library is missing: there is no fallback implementation."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "lib", "libqgemm_sm100.so")

TYPE_Q4_0, TYPE_Q4_1, TYPE_Q5_0, TYPE_Q5_1, TYPE_Q8_0, TYPE_Q8_1 = 2, 3, 6, 7, 8, 9
Q81_ROUND_AWAY, Q81_ROUND_EVEN, Q81_S_FROM_QSUM, Q81_CLAMP127 = 0, 1, 2, 4
Q81_TREE_SUM, Q81_ID_FROM_HALF_D, Q81_ZERO_D1 = 8, 16, 32
Q81_FUSED_F16 = Q81_TREE_SUM | Q81_ID_FROM_HALF_D | Q81_CLAMP127
GEMM_MS_EXACT, GEMM_SEQUENTIAL, GEMM_WEIGHTS_STATIC, GEMM_INPUTS_READY = 0x1, 0x8, 0x10, 0x20
GEMM_WEIGHTS_PREPACKED, GEMM_STREAM_ALLOC, GEMM_FOLD_REFSEQ = 0x40, 0x80, 0x1000
PATH_AUTO, PATH_GENERIC, PATH_GEMV, PATH_MMA, PATH_TCGEN05 = 0x000, 0x100, 0x200, 0x300, 0x400

# name -> (restype, argtypes): every symbol include/qgemm.h declares
_p, _i, _i64, _u32, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_size_t
SYMBOLS = {
    "qgemm_version": (_i, []),
    "qgemm_strerror": (C.c_char_p, [_i]),
    "qgemm_block_bytes": (_sz, [_i]),
    "qgemm_launch_count": (_i64, []),
    "qgemm_reset_launch_count": (None, []),
    "qgemm_last_path": (_u32, []),
    "qgemm_last_error_detail": (C.c_char_p, []),
    "qgemm_quantize_q8_1": (_i, [_p, _p, _i64, _i64, _u32, _p]),
    "qgemm_quantize_q8_1_f16": (_i, [_p, _p, _i64, _i64, _u32, _p]),
    "qgemm_quantize_q8_1_silu_mul": (_i, [_p, _p, _p, _i64, _i64, _u32, _p]),
    "qgemm_quantize_q8_1_rms_norm": (_i, [_p, _p, _p, _i64, _i64, C.c_float, _u32, _p, _p]),
    "qgemm_quantize_weight": (_i, [_i, _p, _p, _i64, _i64, _u32, _p]),
    "qgemm_dequantize": (_i, [_i, _p, _p, _i64, _i64, _p]),
    "qgemm_set_default_workspace": (_i, [_p, _sz]),
    "qgemm_hint_next_weights": (_i, [_p, _sz]),
    "qgemm_prepack_bytes": (_sz, [_i, _i, _i]),
    "qgemm_prepack_weights": (_i, [_i, _p, _i, _i, _p, _p]),
    "qgemm_gemm_group": (_i, [_i, _p, _i, C.POINTER(_p), C.POINTER(_p), C.POINTER(_i), _i, _i, _i64, _i64, _u32, _p]),
    "qgemm_gemm_hinted": (_i, [_i, _p, _p, _p, _i, _i, _i, _i64, _i64, _u32, _p, _sz, _p, _p, _sz]),
    "qgemm_gemm_group_hinted": (_i, [_i, _p, _i, C.POINTER(_p), C.POINTER(_p), C.POINTER(_i), _i, _i, _i64, _i64, _u32, _p, _p, _sz]),
    "qgemm_workspace_bytes": (_sz, [_i, _i, _i, _i, _u32]),
    "qgemm_gemm": (_i, [_i, _p, _p, _p, _i, _i, _i, _i64, _i64, _u32, _p, _sz, _p]),
    "qgemm_gemm_f16act": (_i, [_i, _p, _p, _p, _i, _i, _i, _i64, _i64, _u32, _p, _sz, _p]),
    "qgemm_gemm_f32act": (_i, [_i, _p, _p, _p, _i, _i, _i, _i64, _i64, _u32, _p, _sz, _p]),
    "qgemm_gemm_f32act_silu_mul": (_i, [_i, _p, _p, _p, _p, _i, _i, _i, _i64, _i64, _u32, _p, _sz, _p]),
    "qgemm_gemm_a16": (_i, [_i, _p, _p, _p, _i, _i, _i, _i64, _i64, _u32, _p]),
    "qgemm_sumi": (_i, [_i, _p, _p, _p, _i, _i, _i, _u32, _p, _sz, _p]),
    "qgemm_shard_range": (_i, [_i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
}



class QgemmPeers(C.Structure):
    """struct qgemm_peers of include/qgemm.h"""
    _fields_ = [("world", _i), ("rank", _i), ("C", _p * 8), ("flag", _p * 8), ("done", _p), ("step", _p),
                ("launches_per_step", _u32), ("launch_index", _u32), ("wait_index", _u32), ("C_multicast", _p)]


class QgemmChainStep(C.Structure):
    """struct qgemm_chain_step of include/qgemm.h"""
    _fields_ = [("act_q8_1", _p), ("act_f32", _p), ("gate_f32", _p), ("nmat", _i), ("weights", _p * 3), ("C", _p * 3),
                ("F", _i * 3), ("K", _i), ("ldc_f", _i64), ("flags", _u32)]


PATH_CHAINED = 0x1000000

SYMBOLS.update({
    "qgemm_gemv_chain_sync_bytes": (_sz, [_i]),
    "qgemm_gemv_chain_max_steps": (_i, []),
    "qgemm_gemv_chain": (_i, [_i, C.POINTER(QgemmChainStep), _i, _u32, _p, _sz, _p]),
    "qgemm_gemm_peers": (_i, [_i, _p, _p, C.POINTER(QgemmPeers), _i, _i, _i, _i64, _i64, _u32, _p]),
    "qgemm_gemm_group_peers": (_i, [_i, _p, _i, C.POINTER(_p), C.POINTER(_i), C.POINTER(_i64), C.POINTER(QgemmPeers), _i, _i,
                               _i64, _i64, _u32, _p]),
    "qgemm_peer_step_advance": (_i, [_p, _p]),
    "qgemm_peer_wait": (_i, [C.POINTER(QgemmPeers), _p]),
})

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(or `make -C llama.cpp-quant-gemm_b200/csrc`). quant_gemm has no fallback path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def strerror(code: int) -> str:
    return lib().qgemm_strerror(code).decode()


def raise_on_error(rc: int, what: str) -> None:
    if rc != 0:
        detail = lib().qgemm_last_error_detail().decode() if rc == -4 else ""
        raise RuntimeError(f"qgemm {what} failed: {strerror(rc)} (code {rc}) {detail}")


def shard_range(F: int, world: int, rank: int, align: int = 1) -> tuple[int, int]:
    f0, f1 = _i(), _i()
    raise_on_error(lib().qgemm_shard_range(F, world, rank, align, C.byref(f0), C.byref(f1)), "shard_range")
    return f0.value, f1.value
