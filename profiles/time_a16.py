"""time_a16.py -- W4A16 / W8A16 (fp32 activations, no quantization) entry timed per launch: rotation over >= 768 MB of
distinct weights as one CUDA graph (bench_detail.make_weights), algorithmic bytes = weights + fp32 activations + fp32 C.

    python profiles/time_a16.py gpurun_out/r02_a16.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "llama.cpp-quant-gemm_b200"))
import torch  # noqa: E402
import bench_detail as bd  # noqa: E402
import quant_gemm  # noqa: E402


def time_a16(wt, T, F, K, reps=3):
    dev = torch.device("cuda")
    wbytes = F * (K // 32) * bd.BS[wt]
    n = max(2, min(128, bd.POOL_BYTES // wbytes))
    w = bd.make_weights(torch, wt, F, K, n, dev)
    x = torch.randn((T, K), device=dev)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        quant_gemm.gemm_a16(w[0], x, T, F, K, wt)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(n):
                quant_gemm.gemm_a16(w[i], x, T, F, K, wt)
        g.replay()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.replay()
            e1.record(stream)
            stream.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    abytes = wbytes + 4 * T * K + 4 * T * F
    return {"type": bd.NAMES[wt], "T": T, "F": F, "K": K, "us": best, "gbs": abytes / best / 1e3,
            "tflops": 2.0 * T * F * K / best / 1e6}


if __name__ == "__main__":
    rows = []
    for wt in (2, 8):
        for (T, F, K) in ((1, 4096, 4096), (1, 11008, 4096), (1, 4096, 11008), (4, 11008, 4096), (8, 11008, 4096),
                          (1, 32768, 4096), (64, 4096, 4096), (512, 4096, 4096)):
            r = time_a16(wt, T, F, K)
            print(r, flush=True)
            rows.append(r)
    with open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_a16.json", "w") as f:
        json.dump({"rows": rows}, f, indent=1)
