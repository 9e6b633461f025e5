"""quant_gemm -- host-side mirror of the reference's PyTorch extension.

Same names, argument meaning, shapes, dtypes and error behaviour as
python/quant_gemm/__init__.py:33-89 + csrc/bindings.cpp:19-91 of
qhy991/llama.cpp-quant-gemm, reaching the sm_100a kernels through the C ABI of
libqgemm_sm100.so (include/qgemm.h) with ctypes.  torch is plumbing only:
device memory, the current stream, and torch.distributed for the sharded path.

    import quant_gemm
    wq = quant_gemm.quantize_q4_0(weight)            # [M, K] f32 -> [M, K/32, 18] u8
    aq = quant_gemm.quantize_q8_1(activation)        # [N, K] f32 -> [N, K/32, 36] u8
    out = quant_gemm.gemm_q4_0_q8_1(wq, aq, M, N, K) # [M, N] f32   (M weight rows, N tokens)

Differences from the reference, all supersets: the launch goes on torch's
current stream (the reference uses stream 0, gemm_ops.cu:195,250); q4_1/q5_0/
q5_1/q8_0 entry points exist; gemm_w4a8() fuses quantize_q8_1 of fp32
activations with the GEMM.  There is no CPU path: without the CUDA library or
a B200 the calls raise.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import (  # noqa: F401  (re-exported constants)
    GEMM_FOLD_REFSEQ, GEMM_INPUTS_READY, GEMM_MS_EXACT, GEMM_WEIGHTS_PREPACKED, GEMM_SEQUENTIAL, GEMM_WEIGHTS_STATIC, PATH_AUTO, PATH_CHAINED, PATH_GEMV, PATH_GENERIC, PATH_MMA, PATH_TCGEN05,
    Q81_CLAMP127, Q81_FUSED_F16, Q81_ID_FROM_HALF_D, Q81_ROUND_AWAY, Q81_ROUND_EVEN, Q81_S_FROM_QSUM, Q81_TREE_SUM, Q81_ZERO_D1,
    TYPE_Q4_0, TYPE_Q4_1, TYPE_Q5_0, TYPE_Q5_1, TYPE_Q8_0, TYPE_Q8_1,
)

__version__ = "0.1.0"

# Block sizes (python/quant_gemm/__init__.py:26-30)
QK4_0 = 32
QK8_1 = 32
BLOCK_Q4_0_BYTES = 18
BLOCK_Q8_1_BYTES = 36
BLOCK_BYTES = {TYPE_Q4_0: 18, TYPE_Q4_1: 20, TYPE_Q5_0: 22, TYPE_Q5_1: 24, TYPE_Q8_0: 34, TYPE_Q8_1: 36}

_workspaces: dict[tuple[int, int], torch.Tensor] = {}


def _check(cond: bool, msg: str) -> None:
    # TORCH_CHECK -> RuntimeError (bindings.cpp:20-67)
    if not cond:
        raise RuntimeError(msg)


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor | None:
    """Per-(device, stream) scratch owned by the python shim (the library itself never allocates)."""
    if nbytes == 0:
        return None
    key = (device.index or 0, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _quantize(x: torch.Tensor, qtype: int, flags: int, out: torch.Tensor | None = None) -> torch.Tensor:
    _check(x.is_cuda, "Input must be a CUDA tensor")
    if x.dtype == torch.float16 and qtype == TYPE_Q8_1:
        # superset: fp16 activations (the input type of the reference's gemm_q4_0_fp16_fused, kernels/gemm/gemm_fused.cuh:157)
        _check(x.dim() >= 1 and x.shape[-1] % 32 == 0, f"Last dimension must be divisible by 32, got {x.shape[-1]}")
        x = x.contiguous()
        K = x.shape[-1]
        rows = x.numel() // K if K else 0
        if out is None:
            out = torch.empty(*x.shape[:-1], K // 32, 36, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().qgemm_quantize_q8_1_f16(x.data_ptr(), out.data_ptr(), rows, K, flags, _stream(x))
        _lib.raise_on_error(rc, "quantize")
        return out
    _check(x.dtype == torch.float32, "Input must be float32")
    _check(x.dim() >= 1, "Input must have at least 1 dimension")
    K = x.shape[-1]
    _check(K % 32 == 0, f"Last dimension must be divisible by 32, got {K}")
    x = x.contiguous()
    rows = x.numel() // K if K else 0
    if out is None:
        out = torch.empty(*x.shape[:-1], K // 32, BLOCK_BYTES[qtype], dtype=torch.uint8, device=x.device)
    else:   # superset: write into a caller-owned buffer (a decode runtime keeps its activation buffers)
        _check(out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous()
               and out.numel() == rows * (K // 32) * BLOCK_BYTES[qtype], "out must be a contiguous uint8 tensor of the quantized size")
    with torch.cuda.device(x.device):
        if qtype == TYPE_Q8_1:
            rc = _lib.lib().qgemm_quantize_q8_1(x.data_ptr(), out.data_ptr(), rows, K, flags, _stream(x))
        else:
            rc = _lib.lib().qgemm_quantize_weight(qtype, x.data_ptr(), out.data_ptr(), rows, K, flags, _stream(x))
    _lib.raise_on_error(rc, "quantize")
    return out


def quantize_q4_0(x: torch.Tensor, flags: int = Q81_ROUND_AWAY) -> torch.Tensor:
    """FP32 [..., K] -> Q4_0 bytes [..., K//32, 18] (python/quant_gemm/__init__.py:33-43)."""
    return _quantize(x, TYPE_Q4_0, flags)


def quantize_q8_1(x: torch.Tensor, flags: int = Q81_ROUND_AWAY, out: torch.Tensor | None = None) -> torch.Tensor:
    """FP32 [..., K] -> Q8_1 bytes [..., K//32, 36] (python/quant_gemm/__init__.py:46-56).

    Default flags reproduce include/quantize.h:165-193 byte for byte; pass
    Q81_CLAMP127 for the reference python extension's clamp (gemm_ops.cu:106-108).
    """
    return _quantize(x, TYPE_Q8_1, flags, out)


def quantize_q8_1_silu_mul(x: torch.Tensor, gate: torch.Tensor, flags: int = Q81_ROUND_AWAY) -> torch.Tensor:
    """quantize_q8_1(silu(x) * gate) in one pass: the reference's silu_mul_forward_f32 (kernels/activation/silu.cuh:97-108,
    162-175) folded into the quantizer that feeds the FFN down projection.  x, gate: FP32 [..., K] -> Q8_1 [..., K//32, 36]."""
    _check(x.is_cuda and gate.is_cuda, "Inputs must be CUDA tensors")
    _check(x.dtype == torch.float32 and gate.dtype == torch.float32, "Inputs must be float32")
    _check(x.shape == gate.shape and x.dim() >= 1, "x and gate must have the same shape")
    K = x.shape[-1]
    _check(K % 32 == 0, f"Last dimension must be divisible by 32, got {K}")
    x, gate = x.contiguous(), gate.contiguous()
    rows = x.numel() // K if K else 0
    out = torch.empty(*x.shape[:-1], K // 32, BLOCK_BYTES[TYPE_Q8_1], dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().qgemm_quantize_q8_1_silu_mul(x.data_ptr(), gate.data_ptr(), out.data_ptr(), rows, K, flags, _stream(x))
    _lib.raise_on_error(rc, "quantize_q8_1_silu_mul")
    return out


def quantize_q8_1_rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5, flags: int = Q81_ROUND_AWAY) -> torch.Tensor:
    """quantize_q8_1(rms_norm(x) * weight): the reference's rms_norm_forward_f32 (kernels/normalization/rms_norm.cuh:249-272)
    folded into the quantizer in front of the q/k/v and gate/up projections.  x: FP32 [..., K], weight: FP32 [K]."""
    _check(x.is_cuda and weight.is_cuda, "Inputs must be CUDA tensors")
    _check(x.dtype == torch.float32 and weight.dtype == torch.float32, "Inputs must be float32")
    K = x.shape[-1]
    _check(K % 32 == 0, f"Last dimension must be divisible by 32, got {K}")
    _check(weight.numel() == K, "Weight must have K elements")
    x, weight = x.contiguous(), weight.contiguous()
    rows = x.numel() // K if K else 0
    out = torch.empty(*x.shape[:-1], K // 32, BLOCK_BYTES[TYPE_Q8_1], dtype=torch.uint8, device=x.device)
    scratch = torch.empty(max(rows, 1), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().qgemm_quantize_q8_1_rms_norm(x.data_ptr(), weight.data_ptr(), out.data_ptr(), rows, K, float(eps), flags,
                                                     scratch.data_ptr(), _stream(x))
    _lib.raise_on_error(rc, "quantize_q8_1_rms_norm")
    return out


def quantize_q4_1(x, flags: int = 0):
    return _quantize(x, TYPE_Q4_1, flags)


def quantize_q5_0(x, flags: int = 0):
    return _quantize(x, TYPE_Q5_0, flags)


def quantize_q5_1(x, flags: int = 0):
    return _quantize(x, TYPE_Q5_1, flags)


def quantize_q8_0(x, flags: int = 0):
    return _quantize(x, TYPE_Q8_0, flags)


def dequantize(x_q: torch.Tensor, K: int, qtype: int) -> torch.Tensor:
    _check(x_q.is_cuda, "Input must be a CUDA tensor")
    _check(x_q.dtype == torch.uint8, "Input must be uint8")
    _check(K % 32 == 0, f"K must be divisible by 32, got {K}")
    x_q = x_q.contiguous()
    bs = BLOCK_BYTES[qtype]
    nblocks = x_q.numel() // bs
    _check(nblocks * bs == x_q.numel() and (K == 0 or nblocks % (K // 32) == 0), "Input shape mismatch")
    lead = x_q.shape[:-2]
    out = torch.empty(*lead, K, dtype=torch.float32, device=x_q.device)
    rows = nblocks // (K // 32) if K else 0
    with torch.cuda.device(x_q.device):
        rc = _lib.lib().qgemm_dequantize(qtype, x_q.data_ptr(), out.data_ptr(), rows, K, _stream(x_q))
    _lib.raise_on_error(rc, "dequantize")
    return out


def dequantize_q4_0(x_q: torch.Tensor, K: int) -> torch.Tensor:
    """Q4_0 bytes [..., K//32, 18] -> FP32 [..., K] (python/quant_gemm/__init__.py:78-89)."""
    return dequantize(x_q, K, TYPE_Q4_0)


def prepack_weights(weight_q: torch.Tensor, M: int, K: int, wtype: int) -> torch.Tensor:
    """Static weights -> the prefill kernel's operand-tile layout, once.  Use the result as `weight_q` of
    gemm(..., flags=GEMM_WEIGHTS_PREPACKED) for calls with >= 96 tokens; keep the native blocks for decode."""
    weight_q = weight_q.contiguous()
    L = _lib.lib()
    n = L.qgemm_prepack_bytes(wtype, M, K)
    _check(n > 0, "unsupported type / shape for prepack")
    packed = torch.empty(n, dtype=torch.uint8, device=weight_q.device)
    with torch.cuda.device(weight_q.device):
        rc = L.qgemm_prepack_weights(wtype, weight_q.data_ptr(), M, K, packed.data_ptr(), _stream(weight_q))
    _lib.raise_on_error(rc, "prepack_weights")
    return packed


def hint_next_weights(next_weight_q: torch.Tensor | None, nbytes: int | None = None) -> None:
    """Decode hint: the next gemm() call also prefetches `next_weight_q` (the weights of the GEMV after
    it) into L2 while it runs.  `nbytes` overrides the length, for weights that lie back to back in one
    allocation (the window may then run on into the matrices that follow).  No effect on results."""
    if next_weight_q is None:
        _lib.lib().qgemm_hint_next_weights(None, 0)
    else:
        n = next_weight_q.numel() * next_weight_q.element_size() if nbytes is None else int(nbytes)
        _lib.lib().qgemm_hint_next_weights(next_weight_q.data_ptr(), n)


def gemm(weight_q: torch.Tensor, activation_q: torch.Tensor, M: int, N: int, K: int, wtype: int,
         flags: int = 0, out: torch.Tensor | None = None, next_weight_q: torch.Tensor | None = None,
         next_nbytes: int | None = None) -> torch.Tensor:
    """C[M,N] = W[M,K] @ A[N,K]^T, M = weight rows, N = tokens (ggml convention).

    Checks mirror bindings.cpp:49-70.
    """
    _check(weight_q.is_cuda, "Weight must be a CUDA tensor")
    _check(activation_q.is_cuda, "Activation must be a CUDA tensor")
    _check(weight_q.dtype == torch.uint8, "Weight must be uint8")
    _check(activation_q.dtype == torch.uint8, "Activation must be uint8")
    _check(K % 32 == 0, f"K must be divisible by 32, got {K}")
    nb = K // 32
    bs = BLOCK_BYTES[wtype]
    if flags & GEMM_WEIGHTS_PREPACKED:
        _check(weight_q.numel() == _lib.lib().qgemm_prepack_bytes(wtype, M, K), "Prepacked weight size mismatch")
    else:
        _check(weight_q.numel() == M * nb * bs,
               f"Weight shape mismatch: expected {M * nb * bs} elements, got {weight_q.numel()}")
    _check(activation_q.numel() == N * nb * 36,
           f"Activation shape mismatch: expected {N * nb * 36} elements, got {activation_q.numel()}")
    weight_q = weight_q.contiguous()
    activation_q = activation_q.contiguous()
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=weight_q.device)
    else:
        _check(out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == M * N,
               "out must be a contiguous CUDA float32 tensor of M*N elements")
    L = _lib.lib()
    with torch.cuda.device(weight_q.device):
        ws_bytes = L.qgemm_workspace_bytes(wtype, N, M, K, flags)
        ws = _workspace(weight_q.device, ws_bytes)
        if next_weight_q is not None:   # the L2 hint as an argument of this call (qgemm_gemm_hinted): no state between calls
            nbytes = next_weight_q.numel() * next_weight_q.element_size() if next_nbytes is None else int(next_nbytes)
            rc = L.qgemm_gemm_hinted(wtype, activation_q.data_ptr(), weight_q.data_ptr(), out.data_ptr(), N, M, K,
                                     1, N, flags, ws.data_ptr() if ws is not None else None,
                                     ws.numel() if ws is not None else 0, _stream(weight_q), next_weight_q.data_ptr(), nbytes)
        else:
            rc = L.qgemm_gemm(wtype, activation_q.data_ptr(), weight_q.data_ptr(), out.data_ptr(), N, M, K,
                              1, N, flags, ws.data_ptr() if ws is not None else None,
                              ws.numel() if ws is not None else 0, _stream(weight_q))
    _lib.raise_on_error(rc, "gemm")
    return out


def gemm_group(weights_q: list, activation_q: torch.Tensor, Ms: list, N: int, K: int, wtype: int, flags: int = 0,
               outs: list | None = None) -> list:
    """Several weight matrices (same type, same K) against the same activations in ONE launch: fused
    q/k/v or gate/up projections.  Returns [M_i, N] outputs identical to separate gemm() calls."""
    import ctypes as C
    n = len(weights_q)
    _check(1 <= n <= 8 and len(Ms) == n, "gemm_group takes 1..8 matrices")
    nb = K // 32
    _check(K % 32 == 0, f"K must be divisible by 32, got {K}")
    _check(activation_q.is_cuda and activation_q.dtype == torch.uint8 and activation_q.numel() == N * nb * 36,
           "Activation shape mismatch")
    activation_q = activation_q.contiguous()
    ws = [w.contiguous() for w in weights_q]
    for w, M in zip(ws, Ms):
        _check(w.is_cuda and w.dtype == torch.uint8 and w.numel() == M * nb * BLOCK_BYTES[wtype], "Weight shape mismatch")
    if outs is None:
        outs = [torch.empty((M, N), dtype=torch.float32, device=activation_q.device) for M in Ms]
    wp = (C.c_void_p * n)(*[w.data_ptr() for w in ws])
    cp = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    fs = (C.c_int * n)(*Ms)
    with torch.cuda.device(activation_q.device):
        rc = _lib.lib().qgemm_gemm_group(wtype, activation_q.data_ptr(), n, wp, cp, fs, N, K, 1, N, flags, _stream(activation_q))
    _lib.raise_on_error(rc, "gemm_group")
    return outs


_chain_sync: dict[tuple[int, int], torch.Tensor] = {}


class GemvChain:
    """A list of one-token GEMV steps that ONE persistent launch executes in order (qgemm_gemv_chain): the successor of
    the reference's launch-per-projection decode loop.  Each step is a dict:

        weights : list of 1..3 quantized matrices that share the step's activations (fused q/k/v, gate/up)
        Ms      : their row counts
        K       : row length
        act_q   : ready-made q8_1 activations [1, K/32, 36] u8          -- or --
        act     : fp32 [K] (e.g. the `out` of an earlier step; quantized inside the kernel), optional `gate`: fp32 [K],
                  the value quantized is then silu(act) * gate
        outs    : optional list of [M, 1] fp32 outputs (allocated when absent)
        ready   : True = the activations do not come from an earlier step: the step need not wait for them

    The descriptor array is built once; calling the object launches it on the current stream."""

    def __init__(self, steps: list, wtype: int, flags: int = 0):
        _check(len(steps) >= 1, "GemvChain needs at least one step")
        self.wtype, self.flags, self.n = wtype, flags, len(steps)
        self.arr = (_lib.QgemmChainStep * self.n)()
        self.outs, self._keep = [], []
        dev = None
        for k, st in enumerate(steps):
            ws, Ms, K = [w.contiguous() for w in st["weights"]], list(st["Ms"]), int(st["K"])
            _check(1 <= len(ws) <= 3 and len(Ms) == len(ws), "a chain step takes 1..3 matrices")
            _check(K % 32 == 0, f"K must be divisible by 32, got {K}")
            nb = K // 32
            dev = ws[0].device
            d = self.arr[k]
            for m, (w, M) in enumerate(zip(ws, Ms)):
                _check(w.is_cuda and w.dtype == torch.uint8 and w.numel() == M * nb * BLOCK_BYTES[wtype], "Weight shape mismatch")
            outs = st.get("outs") or [torch.empty((M, 1), dtype=torch.float32, device=dev) for M in Ms]
            for o, M in zip(outs, Ms):
                _check(o.is_cuda and o.dtype == torch.float32 and o.is_contiguous() and o.numel() == M, "out must be M fp32 values")
            if st.get("act_q") is not None:
                a = st["act_q"].contiguous()
                _check(a.is_cuda and a.dtype == torch.uint8 and a.numel() == nb * 36, "Activation shape mismatch")
                d.act_q8_1 = a.data_ptr()
                self._keep.append(a)
            else:
                a = st["act"]
                _check(a.is_cuda and a.dtype == torch.float32 and a.is_contiguous() and a.numel() == K, "act must be K fp32 values")
                d.act_f32 = a.data_ptr()
                self._keep.append(a)
                g = st.get("gate")
                if g is not None:
                    _check(g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.numel() == K, "gate must be K fp32 values")
                    d.gate_f32 = g.data_ptr()
                    self._keep.append(g)
            d.nmat = len(ws)
            for m, (w, o, M) in enumerate(zip(ws, outs, Ms)):
                d.weights[m], d.C[m], d.F[m] = w.data_ptr(), o.data_ptr(), M
            d.K, d.ldc_f = K, 1
            d.flags = GEMM_INPUTS_READY if st.get("ready") else 0
            self._keep += ws
            self.outs.append(outs)
        self.device = dev
        self.sync_bytes = int(_lib.lib().qgemm_gemv_chain_sync_bytes(self.n))

    def _sync(self) -> torch.Tensor:
        key = (self.device.index or 0, torch.cuda.current_stream(self.device).cuda_stream)
        t = _chain_sync.get(key)
        if t is None or t.numel() < self.sync_bytes:
            t = torch.zeros(max(self.sync_bytes, 4096), dtype=torch.uint8, device=self.device)
            torch.cuda.synchronize(self.device)   # the zeros are in place whatever stream the first launch uses
            _chain_sync[key] = t
        return t

    def __call__(self) -> list:
        t = self._sync()
        with torch.cuda.device(self.device):
            rc = _lib.lib().qgemm_gemv_chain(self.wtype, self.arr, self.n, self.flags, t.data_ptr(), t.numel(),
                                             torch.cuda.current_stream(self.device).cuda_stream)
        _lib.raise_on_error(rc, "gemv_chain")
        return self.outs


def gemv_chain(steps: list, wtype: int, flags: int = 0) -> list:
    """One-shot form of GemvChain: returns the list of per-step output lists."""
    return GemvChain(steps, wtype, flags)()


def gemm_q4_0_q8_1(weight_q, activation_q, M, N, K, flags: int = 0):
    """Q4_0 x Q8_1 GEMM -> [M, N] float32 (python/quant_gemm/__init__.py:59-75)."""
    return gemm(weight_q, activation_q, M, N, K, TYPE_Q4_0, flags)


def gemm_q4_1_q8_1(weight_q, activation_q, M, N, K, flags: int = 0):
    return gemm(weight_q, activation_q, M, N, K, TYPE_Q4_1, flags)


def gemm_q5_0_q8_1(weight_q, activation_q, M, N, K, flags: int = 0):
    return gemm(weight_q, activation_q, M, N, K, TYPE_Q5_0, flags)


def gemm_q5_1_q8_1(weight_q, activation_q, M, N, K, flags: int = 0):
    return gemm(weight_q, activation_q, M, N, K, TYPE_Q5_1, flags)


def gemm_q8_0_q8_1(weight_q, activation_q, M, N, K, flags: int = 0):
    return gemm(weight_q, activation_q, M, N, K, TYPE_Q8_0, flags)


def gemm_w4a8(weight_q: torch.Tensor, activation: torch.Tensor, M: int, N: int, K: int,
              wtype: int = TYPE_Q4_0, flags: int = 0, q81_flags: int = Q81_ROUND_AWAY,
              gate: torch.Tensor | None = None) -> torch.Tensor:
    """One call: quantize_q8_1(activation fp32 [N,K]) then the GEMM -> [M, N].

    The reference designs this entry (docs/analysis/W4A8_DATAFLOW_ANALYSIS.md:93-160)
    but never wrote it; its FP16 precursor is kernels/gemm/gemm_fused.cuh:311-338.
    With `gate` (same shape as `activation`) what is quantized is silu(activation) * gate: the FFN down projection
    with its SwiGLU neighbour (kernels/activation/silu.cuh:97-108) folded in.
    """
    _check(weight_q.is_cuda and activation.is_cuda, "Inputs must be CUDA tensors")
    _check(weight_q.dtype == torch.uint8, "Weight must be uint8")
    _check(activation.dtype == torch.float32, "Activation must be float32")
    _check(K % 32 == 0, f"K must be divisible by 32, got {K}")
    nb = K // 32
    _check(weight_q.numel() == M * nb * BLOCK_BYTES[wtype], "Weight shape mismatch")
    _check(activation.numel() == N * K, "Activation shape mismatch")
    weight_q = weight_q.contiguous()
    activation = activation.contiguous()
    if gate is not None:
        _check(gate.is_cuda and gate.dtype == torch.float32 and gate.numel() == N * K, "Gate must match the activation")
        gate = gate.contiguous()
    out = torch.empty((M, N), dtype=torch.float32, device=weight_q.device)
    L = _lib.lib()
    with torch.cuda.device(weight_q.device):
        ws_bytes = L.qgemm_workspace_bytes(wtype, N, M, K, flags)
        ws = _workspace(weight_q.device, ws_bytes)
        tail = (N, M, K, 1, N, (flags & 0xFFFF) | (q81_flags << 16), ws.data_ptr() if ws is not None else None,
                ws.numel() if ws is not None else 0, _stream(weight_q))
        if gate is None:
            rc = L.qgemm_gemm_f32act(wtype, activation.data_ptr(), weight_q.data_ptr(), out.data_ptr(), *tail)
        else:
            rc = L.qgemm_gemm_f32act_silu_mul(wtype, activation.data_ptr(), gate.data_ptr(), weight_q.data_ptr(), out.data_ptr(), *tail)
    _lib.raise_on_error(rc, "gemm_w4a8")
    return out


def gemm_a16(weight_q: torch.Tensor, activation: torch.Tensor, M: int, N: int, K: int, wtype: int = TYPE_Q4_0,
             flags: int = 0) -> torch.Tensor:
    """fp32 activations [M tokens, K] against Q4_0 / Q8_0 weights [N rows, K/32, bytes] with NO activation quantization
    -> [M, N] (the include/ convention, like gemm_w4a16_naive and the python extension's gemm_q4_0_fp32).
    flags=GEMM_SEQUENTIAL reproduces the reference GPU kernel bit for bit."""
    _check(weight_q.is_cuda and activation.is_cuda, "Inputs must be CUDA tensors")
    _check(weight_q.dtype == torch.uint8, "Weight must be uint8")
    _check(activation.dtype == torch.float32, "Activation must be float32")
    _check(wtype in (TYPE_Q4_0, TYPE_Q8_0), "W4A16 / W8A16 exist for Q4_0 and Q8_0 weights")
    _check(K % 32 == 0, f"K must be divisible by 32, got {K}")
    _check(weight_q.numel() == N * (K // 32) * BLOCK_BYTES[wtype], "Weight shape mismatch")
    _check(activation.numel() == M * K, "Activation shape mismatch")
    weight_q = weight_q.contiguous()
    activation = activation.contiguous()
    out = torch.empty((M, N), dtype=torch.float32, device=weight_q.device)
    with torch.cuda.device(weight_q.device):
        rc = _lib.lib().qgemm_gemm_a16(wtype, activation.data_ptr(), weight_q.data_ptr(), out.data_ptr(), M, N, K, N, 1, flags,
                                       _stream(weight_q))
    _lib.raise_on_error(rc, "gemm_a16")
    return out


def gemm_q4_0_fp32(weight_q: torch.Tensor, activation: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    """Q4_0 weights [N, K/32, 18] x fp32 activations [M, K] -> [M, N]: the entry the reference's extension implements
    (python/quant_gemm/csrc/gemm_ops.cu:271-463, gemm_q4_0_fp32_cuda) but never binds."""
    return gemm_a16(weight_q, activation, M, N, K, TYPE_Q4_0)


def gemm_q4_0_fp16_fused(weight_q: torch.Tensor, activation_f16: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    """Q4_0 weights [M, K/32, 18] x FP16 activations [N, K] -> [M, N]: the reference's gemm_q4_0_fp16_fused
    (kernels/gemm/gemm_fused.cuh:311-338, never bound by its extension).  The activations are quantized with the arithmetic of
    that kernel's in-kernel quantizer (Q81_FUSED_F16), then take the q8_1 GEMM."""
    _check(activation_f16.is_cuda and activation_f16.dtype == torch.float16, "Activation must be a CUDA float16 tensor")
    _check(activation_f16.numel() == N * K, f"Activation shape mismatch: expected {N * K} elements, got {activation_f16.numel()}")
    aq = _quantize(activation_f16.reshape(N, K), TYPE_Q8_1, Q81_FUSED_F16)
    return gemm(weight_q, aq, M, N, K, TYPE_Q4_0)


def block_sumi(weight_q: torch.Tensor, activation_q: torch.Tensor, M: int, N: int, K: int, wtype: int,
               flags: int = 0) -> torch.Tensor:
    """Test hook: int32 sumi[N tokens, M rows, K/32] exactly as the selected path computes it."""
    weight_q = weight_q.contiguous()
    activation_q = activation_q.contiguous()
    out = torch.empty((N, M, K // 32), dtype=torch.int32, device=weight_q.device)
    L = _lib.lib()
    with torch.cuda.device(weight_q.device):
        ws_bytes = L.qgemm_workspace_bytes(wtype, N, M, K, flags)
        ws = _workspace(weight_q.device, ws_bytes)
        rc = L.qgemm_sumi(wtype, activation_q.data_ptr(), weight_q.data_ptr(), out.data_ptr(), N, M, K, flags,
                          ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                          _stream(weight_q))
    _lib.raise_on_error(rc, "sumi")
    return out


def launch_count() -> int:
    return int(_lib.lib().qgemm_launch_count())


def reset_launch_count() -> None:
    _lib.lib().qgemm_reset_launch_count()


def last_path() -> int:
    return int(_lib.lib().qgemm_last_path())


__all__ = [
    "quantize_q4_0", "quantize_q8_1", "gemm_q4_0_q8_1", "dequantize_q4_0",
    "QK4_0", "QK8_1", "BLOCK_Q4_0_BYTES", "BLOCK_Q8_1_BYTES",
    # supersets
    "quantize_q4_1", "quantize_q5_0", "quantize_q5_1", "quantize_q8_0", "dequantize",
    "gemm", "gemm_q4_1_q8_1", "gemm_q5_0_q8_1", "gemm_q5_1_q8_1", "gemm_q8_0_q8_1", "gemm_w4a8",
    "gemm_a16", "gemm_q4_0_fp32", "gemm_q4_0_fp16_fused", "quantize_q8_1_silu_mul", "quantize_q8_1_rms_norm",
    "gemm_group", "GemvChain", "gemv_chain", "prepack_weights", "block_sumi", "launch_count", "reset_launch_count", "last_path", "hint_next_weights",
]
