"""T=1 timing of the K=11008 down projection (and the 11008x4096 up projection for reference) under tuning env vars."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
out = []
for F, K in ((4096, 11008), (11008, 4096)):
    r = bench_detail.time_shape(torch, quant_gemm, 2, 1, F, K, 0x10, reps=3, pool_bytes=512 << 20)
    out.append((F, K, round(r["us"], 2)))
print(os.environ.get("TAG", ""), out, flush=True)
