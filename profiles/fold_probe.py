"""Timing experiment behind DESIGN.md 4.3: the tcgen05 call with the int->float conversion removed and/or two MMAs per block."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
from quant_gemm import _lib
L = _lib.lib()
dev = torch.device("cuda")
for wt, T, F, K in ((2, 512, 4096, 4096), (2, 4096, 14336, 8192), (8, 2048, 4096, 4096)):
    w = bench_detail.make_weights(torch, wt, F, K, 1, dev)[0]
    aq = quant_gemm.quantize_q8_1(torch.randn((T, K), device=dev))
    out = torch.empty((F, T), device=dev)
    wsb = L.qgemm_workspace_bytes(wt, T, F, K, 0x400)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    best = 1e9
    for i in range(5):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert L.qgemm_gemm(wt, aq.data_ptr(), w.data_ptr(), out.data_ptr(), T, F, K, 1, T, 0x410, ws.data_ptr(), wsb, st) == 0
        e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    print(os.environ.get("TAG"), wt, T, F, K, "ms", round(best, 4), "tops", round(2.0 * T * F * K / best / 1e9, 1), flush=True)
