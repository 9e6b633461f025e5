"""Skinny-path timing, a few (format, T) points at 11008x4096 and 4096x4096 (graph rotation as bench_detail)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
out = []
for wt in (2, 7, 8):
    for F, K in ((11008, 4096), (4096, 4096)):
        for T in (2, 8):
            r = bench_detail.time_shape(torch, quant_gemm, wt, T, F, K, 0x10, reps=3, pool_bytes=512 << 20)
            out.append((r["type"], T, F, K, round(r["us"], 2)))
print(os.environ.get("TAG", ""), out, flush=True)
