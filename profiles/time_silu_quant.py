"""time_silu_quant.py -- quantize_q8_1(silu(x) * gate): one fused pass vs the reference's structure (an fp32 silu*gate pass
written to HBM, then quantize_q8_1), both on tensors larger than L2.  Algorithmic bytes: fused 8 + 1.125 per element."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llama.cpp-quant-gemm_b200"))
import torch  # noqa: E402
import quant_gemm  # noqa: E402


def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


rows = []
for (T, K) in ((2048, 14336), (4096, 28672), (512, 11008)):
    x = torch.randn((T, K), device="cuda")
    g = torch.randn((T, K), device="cuda")
    fused = t(lambda: quant_gemm.quantize_q8_1_silu_mul(x, g))
    two = t(lambda: quant_gemm.quantize_q8_1(torch.nn.functional.silu(x) * g))
    plain = t(lambda: quant_gemm.quantize_q8_1(x))
    n = T * K
    r = {"T": T, "K": K, "fused_us": fused, "fused_gbs": n * 9.125 / fused / 1e3, "silu_mul_then_quantize_us": two,
         "plain_quantize_us": plain, "plain_quantize_gbs": n * 5.125 / plain / 1e3}
    print(r, flush=True)
    rows.append(r)
json.dump({"rows": rows}, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_silu_quant.json", "w"), indent=1)
