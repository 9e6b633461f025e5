"""Small-batch timing (9 <= T < 128) at Llama shapes, graph rotation as bench_detail."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
out = []
for wt, F, K in ((2, 11008, 4096), (8, 11008, 4096), (2, 4096, 4096), (2, 8192, 8192)):
    for T in (8, 16, 32, 64, 96, 128):
        r = bench_detail.time_shape(torch, quant_gemm, wt, T, F, K, 0x10, reps=3, pool_bytes=384 << 20)
        out.append((r["type"], T, F, K, round(r["us"], 1), r["path"]))
        print(os.environ.get("TAG", ""), out[-1], flush=True)
