"""ncu driver: the bench workload's launch sequence (q4_0, M=1, grouped like bench.py: [wq wk wv], wo, [gate up], down; the
weights back to back in one allocation, every launch hinting the next 12 MB of it into L2), cold weights, 8 layers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, quant_gemm
dev = torch.device("cuda")
LAYERS = 8
WINDOW = 12 << 20
SHAPES = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
g = torch.Generator(device=dev)
g.manual_seed(1)
sizes = [F * (K // 32) * 18 for _ in range(LAYERS) for F, K in SHAPES]
arena = torch.empty(sum((n + 255) // 256 * 256 for n in sizes), dtype=torch.uint8, device=dev)
mats, off = [], 0
for _ in range(LAYERS):
    for F, K in SHAPES:
        nb = K // 32
        w = arena[off:off + F * nb * 18].view(F, nb, 18)
        off += (F * nb * 18 + 255) // 256 * 256
        w.copy_(torch.randint(0, 256, (F, nb, 18), dtype=torch.uint8, device=dev, generator=g))
        d = (torch.rand((F, nb), device=dev, generator=g) * 0.02 + 0.001).to(torch.float16)
        w[:, :, 0:2] = d.view(torch.uint8).view(F, nb, 2)
        mats.append(w)
end = arena.data_ptr() + arena.numel()
aq = {K: quant_gemm.quantize_q8_1(torch.randn((1, K), device=dev)) for K in (4096, 11008)}
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
flush.zero_()
torch.cuda.synchronize()


def hint(w):
    quant_gemm.hint_next_weights(w, min(WINDOW, end - w.data_ptr()))


CHAIN = int(os.environ.get("PROF_CHAIN", "0"))   # layers per chained launch (qgemm_gemv_chain); 0: one launch per grouped projection
if CHAIN:
    chains = []
    for c0 in range(0, LAYERS, CHAIN):
        steps = []
        for l in range(c0, min(c0 + CHAIN, LAYERS)):
            m = mats[7 * l:7 * l + 7]
            steps += [{"weights": [m[0], m[1], m[2]], "Ms": [4096] * 3, "K": 4096, "act_q": aq[4096]},
                      {"weights": [m[3]], "Ms": [4096], "K": 4096, "act_q": aq[4096]},
                      {"weights": [m[4], m[5]], "Ms": [11008] * 2, "K": 4096, "act_q": aq[4096]},
                      {"weights": [m[6]], "Ms": [4096], "K": 11008, "act_q": aq[11008]}]
        chains.append(quant_gemm.GemvChain(steps, 2, 0x10))
    for ci, ch in enumerate(chains):
        hint(mats[(7 * CHAIN * (ci + 1)) % len(mats)])
        ch()
    torch.cuda.synchronize()
    print("ok (chained)")
    sys.exit(0)

for l in range(LAYERS):
    m = mats[7 * l:7 * l + 7]
    nxt = mats[(7 * l + 7) % len(mats)]
    hint(m[3])
    quant_gemm.gemm_group([m[0], m[1], m[2]], aq[4096], [4096] * 3, 1, 4096, 2, 0x10)
    hint(m[4])
    quant_gemm.gemm(m[3], aq[4096], 4096, 1, 4096, 2, 0x10)
    hint(m[6])
    quant_gemm.gemm_group([m[4], m[5]], aq[4096], [11008] * 2, 1, 4096, 2, 0x10)
    hint(nxt)
    quant_gemm.gemm(m[6], aq[11008], 4096, 1, 11008, 2, 0x10)
torch.cuda.synchronize()
print("ok")
