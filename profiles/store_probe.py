"""How much of a large prefill call is the C store?  plain tcgen05 call, [F,T] vs [T,F] strides, stores on/off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
from quant_gemm import _lib
L = _lib.lib()
dev = torch.device("cuda")
T, F, K = 4096, 14336, 8192
w = bench_detail.make_weights(torch, 2, F, K, 1, dev)[0]
aq = quant_gemm.quantize_q8_1(torch.randn((T, K), device=dev))
out = torch.empty((F, T), device=dev)
wsb = L.qgemm_workspace_bytes(2, T, F, K, 0x400)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for name, (lt, lf) in (("FT", (1, T)), ("TF", (F, 1))):
    best = 1e9
    for i in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert L.qgemm_gemm(2, aq.data_ptr(), w.data_ptr(), out.data_ptr(), T, F, K, lt, lf, 0x410, ws.data_ptr(), wsb, st) == 0
        e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    print(name, "dbg", os.environ.get("QGEMM_MMQ_DBG"), "ms", round(best, 3), flush=True)
