"""Tiny driver for ncu: a few prefill GEMM calls.  python profiles/prof_prefill.py <wt> <T> <F> <K> [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
wt, T, F, K = (int(v) for v in sys.argv[1:5])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda")
w = bench_detail.make_weights(torch, wt, F, K, 1, dev)[0]
aq = quant_gemm.quantize_q8_1(torch.randn((T, K), device=dev))
for i in range(n):
    out = quant_gemm.gemm(w, aq, F, T, K, wt)
torch.cuda.synchronize()
print("ok", hex(quant_gemm.last_path()))
