"""ncu driver: the four launches of one layer of the bench workload (q4_0, M=1, grouped like bench.py:
[wq wk wv], wo, [gate up], down; each hinting the next one's weights), cold weights, 3 layers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm, bench_detail
dev = torch.device("cuda")
LAYERS = 8
sq = bench_detail.make_weights(torch, 2, 4096, 4096, 4 * LAYERS, dev)
up = bench_detail.make_weights(torch, 2, 11008, 4096, 2 * LAYERS, dev)
dn = bench_detail.make_weights(torch, 2, 4096, 11008, LAYERS, dev)
aq = {K: quant_gemm.quantize_q8_1(torch.randn((1, K), device=dev)) for K in (4096, 11008)}
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
flush.zero_()
torch.cuda.synchronize()
for l in range(LAYERS):
    q = sq[4 * l:4 * l + 4]
    quant_gemm.hint_next_weights(q[3])
    quant_gemm.gemm_group([q[0], q[1], q[2]], aq[4096], [4096] * 3, 1, 4096, 2, 0x10)
    quant_gemm.hint_next_weights(up[2 * l])
    quant_gemm.gemm(q[3], aq[4096], 4096, 1, 4096, 2, 0x10)
    quant_gemm.hint_next_weights(dn[l])
    quant_gemm.gemm_group([up[2 * l], up[2 * l + 1]], aq[4096], [11008] * 2, 1, 4096, 2, 0x10)
    quant_gemm.hint_next_weights(sq[(4 * l + 4) % (4 * LAYERS)])
    quant_gemm.gemm(dn[l], aq[11008], 4096, 1, 11008, 2, 0x10)
torch.cuda.synchronize()
print("ok")
