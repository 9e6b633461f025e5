"""ctypes front-end for the CPU oracle and the reference doorway library.

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(llama.cpp-quant-gemm_b200/) must never import this module.

Two libraries:
  * oracle/libqgemm_oracle.so   -- qgemm_oracle.c, the plain-C restatement
                                   (built with gcc on demand; travels with the repo)
  * oracle/_ref/libqgemm_ref.so -- the reference's own sources compiled from
                                   /root/reference (built by `make -C oracle ref`
                                   in the dev container only; the prebuilt file
                                   travels to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libqgemm_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libqgemm_ref.so")

Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, Q8_1 = 2, 3, 6, 7, 8, 9
WEIGHT_TYPES = (Q4_0, Q4_1, Q5_0, Q5_1, Q8_0)
BLOCK_BYTES = {Q4_0: 18, Q4_1: 20, Q5_0: 22, Q5_1: 24, Q8_0: 34, Q8_1: 36}
TYPE_NAMES = {Q4_0: "q4_0", Q4_1: "q4_1", Q5_0: "q5_0", Q5_1: "q5_1", Q8_0: "q8_0", Q8_1: "q8_1"}

# quantize_q8_1 flags / gemm flags (same bits as include/qgemm.h)
Q81_ROUND_AWAY, Q81_ROUND_EVEN, Q81_S_FROM_QSUM, Q81_CLAMP127 = 0, 1, 2, 4
Q81_TREE_SUM, Q81_ID_FROM_HALF_D, Q81_ZERO_D1 = 8, 16, 32
Q81_FUSED_F16 = Q81_TREE_SUM | Q81_ID_FROM_HALF_D | Q81_CLAMP127   # kernels/gemm/gemm_fused.cuh:76-143
GEMM_MS_EXACT, GEMM_Q80_ASSOC_UNIT, GEMM_FMA = 1, 2, 4


def build_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "qgemm_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _ORACLE_SO


def build_ref(reference_root: str = "/root/reference") -> str | None:
    """Build oracle/_ref/libqgemm_ref.so if the reference tree is present (dev container)."""
    if os.path.isdir(reference_root):
        subprocess.check_call(["make", "-C", _HERE, "ref", f"REF={reference_root}"], stdout=subprocess.DEVNULL)
    return _REF_SO if os.path.exists(_REF_SO) else None


_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_u = C.c_uint
_f = C.c_float


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_p)


class Oracle:
    """qgemm_oracle.c through ctypes; numpy in, numpy out."""

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.qo_fp32_to_fp16.restype = C.c_uint16
        L.qo_fp32_to_fp16.argtypes = [_f]
        L.qo_fp16_to_fp32.restype = _f
        L.qo_fp16_to_fp32.argtypes = [C.c_uint16]
        L.qo_quantize_q8_1.argtypes = [_p, _p, _i64, _u]
        L.qo_quantize_q8_1_f16.argtypes = [_p, _p, _i64, _u]
        L.qo_silu_mul.argtypes = [_p, _p, _p, _i64]
        L.qo_rms_norm.argtypes = [_p, _p, _p, _i64, _i64, _f]
        for n in ("qo_quantize_q4_0_ref", "qo_quantize_q8_0_ref", "qo_to_q4_0", "qo_to_q4_1",
                  "qo_to_q5_0", "qo_to_q5_1", "qo_to_q8_0"):
            getattr(L, n).argtypes = [_p, _p, _i64]
        L.qo_dequantize.argtypes = [_i, _p, _p, _i64]
        L.qo_block_sumi.restype = C.c_int32
        L.qo_block_sumi.argtypes = [_i, _p, _p]
        L.qo_block_dot.restype = _f
        L.qo_block_dot.argtypes = [_i, _p, _p, _u]
        L.qo_gemm.argtypes = [_i, _p, _p, _p, _i, _i, _i, _i64, _i64, _u, _i, _i]
        L.qo_gemm_sumi.argtypes = [_i, _p, _p, _p, _i, _i, _i, _i, _i]
        L.qo_gemm_f32act_dequant.argtypes = [_i, _p, _p, _p, _i, _i, _i, _i64, _i64, _u]

    # -- scalars
    def f2h(self, x: float) -> int:
        return int(self.lib.qo_fp32_to_fp16(x))

    def h2f(self, h: int) -> float:
        return float(self.lib.qo_fp16_to_fp32(h))

    # -- quantizers: x float32 [..., K] -> uint8 [..., K/32, bytes]
    def _quant(self, fn, x: np.ndarray, bs: int, *extra) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.shape[-1] % 32 == 0
        out = np.empty(x.shape[:-1] + (x.shape[-1] // 32, bs), dtype=np.uint8)
        fn(_ptr(x), _ptr(out), x.size, *extra)
        return out

    def quantize_q8_1(self, x, flags: int = Q81_ROUND_AWAY):
        return self._quant(self.lib.qo_quantize_q8_1, x, 36, flags)

    def quantize_q8_1_f16(self, x_f16, flags: int = Q81_FUSED_F16):
        """fp16 input [..., K] -> q8_1, the reference's in-kernel quantizer (kernels/gemm/gemm_fused.cuh:76-143) by default."""
        x = np.ascontiguousarray(x_f16, dtype=np.float16)
        assert x.shape[-1] % 32 == 0
        out = np.empty(x.shape[:-1] + (x.shape[-1] // 32, 36), dtype=np.uint8)
        self.lib.qo_quantize_q8_1_f16(_ptr(x), _ptr(out), x.size, flags)
        return out

    def silu_mul(self, x, gate) -> np.ndarray:
        """silu(x) * gate in fp32, operation order of kernels/activation/silu.cuh:97-108."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        gate = np.ascontiguousarray(gate, dtype=np.float32)
        assert x.shape == gate.shape
        out = np.empty_like(x)
        self.lib.qo_silu_mul(_ptr(x), _ptr(gate), _ptr(out), x.size)
        return out

    def rms_norm(self, x, weight, eps: float = 1e-5) -> np.ndarray:
        """rms_norm(x) * weight over the last axis, arithmetic of kernels/normalization/rms_norm.cuh:32-58."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        weight = np.ascontiguousarray(weight, dtype=np.float32)
        assert weight.size == x.shape[-1]
        out = np.empty_like(x)
        self.lib.qo_rms_norm(_ptr(x), _ptr(weight), _ptr(out), x.size // x.shape[-1], x.shape[-1], eps)
        return out

    def quantize_weight(self, wtype: int, x, flavour: str = "framework"):
        """flavour 'include' = include/quantize.h (q4_0/q8_0 only), 'framework' = tests/framework."""
        L = self.lib
        if flavour == "include":
            fn = {Q4_0: L.qo_quantize_q4_0_ref, Q8_0: L.qo_quantize_q8_0_ref}[wtype]
        else:
            fn = {Q4_0: L.qo_to_q4_0, Q4_1: L.qo_to_q4_1, Q5_0: L.qo_to_q5_0, Q5_1: L.qo_to_q5_1,
                  Q8_0: L.qo_to_q8_0}[wtype]
        return self._quant(fn, x, BLOCK_BYTES[wtype])

    def dequantize(self, qtype: int, q: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.uint8)
        assert q.shape[-1] == BLOCK_BYTES[qtype]
        out = np.empty(q.shape[:-2] + (q.shape[-2] * 32,), dtype=np.float32)
        self.lib.qo_dequantize(qtype, _ptr(q), _ptr(out), out.size)
        return out

    def block_sumi(self, wtype, wblock: np.ndarray, ablock: np.ndarray) -> int:
        return int(self.lib.qo_block_sumi(wtype, _ptr(np.ascontiguousarray(wblock)), _ptr(np.ascontiguousarray(ablock))))

    def block_dot(self, wtype, wblock, ablock, flags: int = 0) -> float:
        return float(self.lib.qo_block_dot(wtype, _ptr(np.ascontiguousarray(wblock)),
                                           _ptr(np.ascontiguousarray(ablock)), flags))

    # -- GEMM
    def gemm(self, wtype: int, act: np.ndarray, weight: np.ndarray, *, layout: str = "TF",
             flags: int = 0, threads: int = 0) -> np.ndarray:
        """act uint8 [T, nb, 36], weight uint8 [F, nb, bs] -> C float32.

        layout 'TF': C[T,F] (include/ convention), 'FT': C[F,T] (kernels/gemm + python).
        """
        act = np.ascontiguousarray(act, dtype=np.uint8)
        weight = np.ascontiguousarray(weight, dtype=np.uint8)
        T, nb, _ = act.shape
        F = weight.shape[0]
        assert weight.shape[1] == nb and weight.shape[2] == BLOCK_BYTES[wtype]
        K = nb * 32
        if layout == "TF":
            out = np.empty((T, F), dtype=np.float32)
            ldc_t, ldc_f = F, 1
        else:
            out = np.empty((F, T), dtype=np.float32)
            ldc_t, ldc_f = 1, T
        nthreads = threads or min(os.cpu_count() or 1, max(1, T))
        if nthreads <= 1 or T == 1:
            self.lib.qo_gemm(wtype, _ptr(act), _ptr(weight), _ptr(out), T, F, K, ldc_t, ldc_f, flags, 0, T)
        else:
            slab = (T + nthreads - 1) // nthreads
            ths = []
            for i in range(nthreads):
                t0, t1 = i * slab, min(T, (i + 1) * slab)
                if t0 >= t1:
                    break
                th = threading.Thread(target=self.lib.qo_gemm, args=(
                    wtype, _ptr(act), _ptr(weight), _ptr(out), T, F, K, ldc_t, ldc_f, flags, t0, t1))
                th.start()
                ths.append(th)
            for th in ths:
                th.join()
        return out

    def gemm_f32act_dequant(self, wtype: int, act_f32: np.ndarray, weight: np.ndarray, *, layout: str = "TF",
                            flags: int = 0) -> np.ndarray:
        """W4A16 / W8A16 (oracle half): act float32 [T, K], weight uint8 [F, nb, bs] (Q4_0 / Q8_0)."""
        assert wtype in (Q4_0, Q8_0)
        act_f32 = np.ascontiguousarray(act_f32, dtype=np.float32)
        weight = np.ascontiguousarray(weight, dtype=np.uint8)
        T, K = act_f32.shape
        F = weight.shape[0]
        assert weight.shape[1] * 32 == K and weight.shape[2] == BLOCK_BYTES[wtype]
        if layout == "TF":
            out, ldc_t, ldc_f = np.empty((T, F), dtype=np.float32), F, 1
        else:
            out, ldc_t, ldc_f = np.empty((F, T), dtype=np.float32), 1, T
        self.lib.qo_gemm_f32act_dequant(wtype, _ptr(act_f32), _ptr(weight), _ptr(out), T, F, K, ldc_t, ldc_f, flags)
        return out

    def gemm_sumi(self, wtype: int, act: np.ndarray, weight: np.ndarray) -> np.ndarray:
        act = np.ascontiguousarray(act, dtype=np.uint8)
        weight = np.ascontiguousarray(weight, dtype=np.uint8)
        T, nb, _ = act.shape
        F = weight.shape[0]
        out = np.empty((T, F, nb), dtype=np.int32)
        self.lib.qo_gemm_sumi(wtype, _ptr(act), _ptr(weight), _ptr(out), T, F, nb * 32, 0, T)
        return out


class Reference:
    """The reference's own compiled code (oracle/_ref/libqgemm_ref.so)."""

    def __init__(self, so: str | None = None):
        so = so or _REF_SO
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        self.lib = C.CDLL(so)
        L = self.lib
        for n in ("ref_quantize_row_q8_1_ref", "ref_quantize_row_q4_0_ref", "ref_quantize_row_q8_0_ref"):
            getattr(L, n).argtypes = [_p, _p, _i64]
        for n in ("ref_dequantize_row_q4_0", "ref_dequantize_row_q8_0", "ref_dequantize_row_q8_1"):
            getattr(L, n).argtypes = [_p, _p, _i64]
        for n in ("ref_to_q4_0", "ref_to_q4_1", "ref_to_q5_0", "ref_to_q5_1", "ref_to_q8_0", "ref_to_q8_1"):
            getattr(L, n).argtypes = [_p, _p, _i]
        L.ref_gemm_w4a8_reference.argtypes = [_p, _p, _p, _i, _i, _i]
        L.ref_gemm_w8a8_reference.argtypes = [_p, _p, _p, _i, _i, _i]
        for n in ("ref_gemm_w4a16_reference", "ref_gemm_w8a16_reference"):
            if hasattr(L, n):   # a prebuilt library from before the fp32-activation references were added lacks them
                getattr(L, n).argtypes = [_p, _p, _p, _i, _i, _i]
        for n in ("ref_gpu_gemm_w4a16_naive", "ref_gpu_gemm_w8a16_naive"):
            if hasattr(L, n):
                getattr(L, n).argtypes = [_p, _p, _p, _i, _i, _i, _p]
        if hasattr(L, "ref_gpu_silu_mul_f32"):   # kernels/activation/silu.cuh (added with the fused SwiGLU quantizer)
            L.ref_gpu_silu_mul_f32.argtypes = [_p, _p, _p, _i, _p]
            L.ref_cpu_silu_f32.argtypes = [_p, _p, _i]
        if hasattr(L, "ref_gpu_quantize_fp16_to_q8_1_smem"):   # kernels/gemm/gemm_fused.cuh:76-143 (ref_shim_fused.cu)
            L.ref_gpu_quantize_fp16_to_q8_1_smem.argtypes = [_p, _p, _i, _p]
        if hasattr(L, "ref_cpu_rms_norm_f32"):   # kernels/normalization/rms_norm.cuh
            L.ref_cpu_rms_norm_f32.argtypes = [_p, _p, _p, _i, _i, _f]
            L.ref_gpu_rms_norm_f32.argtypes = [_p, _p, _p, _i, _i, _f, _p]
        L.ref_vec_dot_q4_0_q8_1.restype = _f
        L.ref_vec_dot_q4_0_q8_1.argtypes = [_i, _p, _p]
        L.ref_vec_dot_q8_0_q8_1.restype = _f
        L.ref_vec_dot_q8_0_q8_1.argtypes = [_i, _p, _p]
        L.ref_float2half_bits.restype = C.c_uint16
        L.ref_float2half_bits.argtypes = [_f]
        L.ref_half_bits2float.restype = _f
        L.ref_half_bits2float.argtypes = [C.c_uint16]
        L.ref_cpu_gemm.argtypes = [_i, _p, _p, _p, _i, _i, _i]
        L.ref_cpu_gemm_threaded.argtypes = [_i, _p, _p, _p, _i, _i, _i, _i]
        # GPU launchers (device pointers as integers)
        L.ref_gpu_quantize_q8_1.argtypes = [_p, _p, _i64, _p]
        for n in ("ref_gpu_gemm_w4a8_naive", "ref_gpu_gemm_w8a8_naive", "ref_gpu_gemm_w4a8_tiled_dp4a",
                  "ref_gpu_gemm_w8a8_dp4a"):
            getattr(L, n).argtypes = [_p, _p, _p, _i, _i, _i, _p]
        L.ref_gpu_gemm_quant.argtypes = [_i, _p, _p, _p, _i, _i, _i, _p]
        L.ref_gpu_gemm_q4_0_tile2d.argtypes = [_p, _p, _p, _i, _i, _i, _p]
        L.ref_gpu_gemm_q4_0_warp_multirow.argtypes = [_p, _p, _p, _i, _i, _i, _p]

    def _quant(self, fn, x, bs):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty(x.shape[:-1] + (x.shape[-1] // 32, bs), dtype=np.uint8)
        fn(_ptr(x), _ptr(out), x.size)
        return out

    def quantize_row_q8_1_ref(self, x):
        return self._quant(self.lib.ref_quantize_row_q8_1_ref, x, 36)

    def quantize_row_q4_0_ref(self, x):
        return self._quant(self.lib.ref_quantize_row_q4_0_ref, x, 18)

    def quantize_row_q8_0_ref(self, x):
        return self._quant(self.lib.ref_quantize_row_q8_0_ref, x, 34)

    def to_q(self, qtype: int, x):
        fn = getattr(self.lib, "ref_to_" + TYPE_NAMES[qtype])
        return self._quant(fn, x, BLOCK_BYTES[qtype])

    def dequantize(self, qtype: int, q):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        out = np.empty(q.shape[:-2] + (q.shape[-2] * 32,), dtype=np.float32)
        fn = {Q4_0: self.lib.ref_dequantize_row_q4_0, Q8_0: self.lib.ref_dequantize_row_q8_0,
              Q8_1: self.lib.ref_dequantize_row_q8_1}[qtype]
        fn(_ptr(q), _ptr(out), out.size)
        return out

    def gemm_include(self, wtype: int, act, weight) -> np.ndarray:
        """include/gemm_reference.h: C[M=T, N=F]."""
        act = np.ascontiguousarray(act, dtype=np.uint8)
        weight = np.ascontiguousarray(weight, dtype=np.uint8)
        T, nb, _ = act.shape
        F = weight.shape[0]
        out = np.empty((T, F), dtype=np.float32)
        fn = {Q4_0: self.lib.ref_gemm_w4a8_reference, Q8_0: self.lib.ref_gemm_w8a8_reference}[wtype]
        fn(_ptr(act), _ptr(weight), _ptr(out), T, F, nb * 32)
        return out

    def gemm_f32act_include(self, wtype: int, act_f32, weight) -> np.ndarray:
        """include/gemm_reference.h:73-147 gemm_w4a16_reference / gemm_w8a16_reference: C[M=T, N=F]."""
        act_f32 = np.ascontiguousarray(act_f32, dtype=np.float32)
        weight = np.ascontiguousarray(weight, dtype=np.uint8)
        T, K = act_f32.shape
        F = weight.shape[0]
        out = np.empty((T, F), dtype=np.float32)
        fn = {Q4_0: self.lib.ref_gemm_w4a16_reference, Q8_0: self.lib.ref_gemm_w8a16_reference}[wtype]
        fn(_ptr(act_f32), _ptr(weight), _ptr(out), T, F, K)
        return out

    def cpu_gemm(self, wtype: int, weight, act, threads: int = 1) -> np.ndarray:
        """tests/unit/test_gemm_all_quants.cu: output[M=F, N=T]."""
        act = np.ascontiguousarray(act, dtype=np.uint8)
        weight = np.ascontiguousarray(weight, dtype=np.uint8)
        T, nb, _ = act.shape
        F = weight.shape[0]
        out = np.empty((F, T), dtype=np.float32)
        if threads <= 1:
            rc = self.lib.ref_cpu_gemm(wtype, _ptr(weight), _ptr(act), _ptr(out), F, T, nb * 32)
        else:
            rc = self.lib.ref_cpu_gemm_threaded(wtype, _ptr(weight), _ptr(act), _ptr(out), F, T, nb * 32, threads)
        assert rc == 0
        return out

    def vec_dot(self, wtype: int, wrow, arow) -> float:
        wrow = np.ascontiguousarray(wrow, dtype=np.uint8)
        arow = np.ascontiguousarray(arow, dtype=np.uint8)
        n = arow.shape[0] * 32
        fn = {Q4_0: self.lib.ref_vec_dot_q4_0_q8_1, Q8_0: self.lib.ref_vec_dot_q8_0_q8_1}[wtype]
        return float(fn(n, _ptr(wrow), _ptr(arow)))


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


# ---------------------------------------------------------------------------
# error metrics (tests/framework/test_framework.cuh:41-64 definition of NMSE)
# ---------------------------------------------------------------------------
def nmse(actual: np.ndarray, expected: np.ndarray) -> float:
    a = actual.astype(np.float64).ravel()
    e = expected.astype(np.float64).ravel()
    den = float(np.sum(e * e))
    return float(np.sum((a - e) ** 2) / den) if den > 0 else 0.0


def max_norm_err(actual: np.ndarray, expected: np.ndarray) -> float:
    """max|dC| / max|C| -- the normalised tolerance of SURVEY.md section 8d."""
    a = actual.astype(np.float64).ravel()
    e = expected.astype(np.float64).ravel()
    den = float(np.max(np.abs(e)))
    return float(np.max(np.abs(a - e)) / den) if den > 0 else float(np.max(np.abs(a - e)))
