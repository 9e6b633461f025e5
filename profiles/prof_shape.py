"""Tiny driver for ncu: a few launches of one (type, T, F, K) shape over distinct weight matrices.
    python profiles/prof_shape.py 2 1 11008 4096 [flags] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
wt, T, F, K = (int(v) for v in sys.argv[1:5])
flags = int(sys.argv[5], 0) if len(sys.argv) > 5 else 0
n = int(sys.argv[6]) if len(sys.argv) > 6 else 12
dev = torch.device("cuda")
w = bench_detail.make_weights(torch, wt, F, K, n, dev)
aq = quant_gemm.quantize_q8_1(torch.randn((T, K), device=dev))
out = torch.empty((F, T), device=dev)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(n):
    quant_gemm.gemm(w[i], aq, F, T, K, wt, flags, out=out)
torch.cuda.synchronize()
print("ok", quant_gemm.launch_count(), hex(quant_gemm.last_path()))
