"""quantize_q8_1 throughput (HBM-bound: 4 B read + 1.125 B written per element) vs the reference GPU kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), os.path.join(ROOT, "oracle"), ROOT]
import torch, quant_gemm
import qgemm_oracle as qo
dev = torch.device("cuda")
R = qo.Reference() if qo.have_ref() else None
for (T, K) in [(1, 4096), (512, 4096), (2048, 4096), (4096, 8192), (16384, 8192)]:
    n = max(2, min(64, (1 << 30) // (T * K * 4)))
    xs = torch.randn((n, T, K), device=dev)
    ys = torch.empty((n, T, K // 32, 36), dtype=torch.uint8, device=dev)
    L = quant_gemm._lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    def ours():
        for i in range(n):
            L.qgemm_quantize_q8_1(xs[i].data_ptr(), ys[i].data_ptr(), T, K, 0, st)
    def ref():
        for i in range(n):
            R.lib.ref_gpu_quantize_q8_1(xs[i].data_ptr(), ys[i].data_ptr(), T * K, None)
    res = {}
    for name, fn in (("ours", ours),) + ((("reference_gpu", ref),) if R else ()):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        res[name] = (us, T * K * 5.125 / us / 1e3)
    print(f"T={T} K={K}: " + "  ".join(f"{k}: {v[0]:.2f} us {v[1]:.0f} GB/s" for k, v in res.items()), flush=True)
