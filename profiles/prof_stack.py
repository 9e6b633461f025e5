"""ncu driver: one decode GEMV launch per distinct shape of the bench workload (q4_0, M=1), cold weights."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
dev = torch.device("cuda")
shapes = [(4096, 4096), (11008, 4096), (4096, 11008)]
ws = {s: bench_detail.make_weights(torch, 2, s[0], s[1], 4, dev) for s in shapes}
aq = {K: quant_gemm.quantize_q8_1(torch.randn((1, K), device=dev)) for K in (4096, 11008)}
for rep in range(4):
    for (F, K) in shapes:
        quant_gemm.gemm(ws[(F, K)][rep], aq[K], F, 1, K, 2, 0x10)
torch.cuda.synchronize()
print("ok")
