#!/usr/bin/env python
"""bench_detail.py -- per-shape sweep behind the headline number (not a bench line).

For every (format, T, F, K) it rotates over a pool of distinct weight matrices >= 4x the L2 so each
launch streams from HBM, replays the rotation as one CUDA graph, and reports us per launch, the
algorithmic GB/s (SURVEY.md section 8d byte model) and, for T >= 64, TOPS = 2*T*F*K / t.

    python bench_detail.py --out profiles/detail_r01.json [--quick]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

BS = {2: 18, 3: 20, 6: 22, 7: 24, 8: 34}
NAMES = {2: "q4_0", 3: "q4_1", 6: "q5_0", 7: "q5_1", 8: "q8_0"}
POOL_BYTES = 768 << 20


def make_weights(torch, wtype, F, K, n, dev, seed=0):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    nb = K // 32
    w = torch.randint(0, 256, (n, F, nb, BS[wtype]), dtype=torch.uint8, device=dev, generator=g)
    d = (torch.rand((n, F, nb), device=dev, generator=g) * 0.02 + 0.001).to(torch.float16)
    w[..., 0:2] = d.view(torch.uint8).view(n, F, nb, 2)
    if wtype in (3, 7):
        m = (torch.rand((n, F, nb), device=dev, generator=g) - 0.5).to(torch.float16)
        w[..., 2:4] = m.view(torch.uint8).view(n, F, nb, 2)
    return w


def time_shape(torch, quant_gemm, wtype, T, F, K, flags=0x10, reps=3, pool_bytes=POOL_BYTES):
    dev = torch.device("cuda")
    wbytes = F * (K // 32) * BS[wtype]
    n = max(2, min(256, pool_bytes // wbytes))
    w = make_weights(torch, wtype, F, K, n, dev)
    x = torch.randn((T, K), device=dev)
    aq = quant_gemm.quantize_q8_1(x)
    out = torch.empty((F, T), device=dev)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        quant_gemm.gemm(w[0], aq, F, T, K, wtype, flags, out=out)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(n):
                quant_gemm.gemm(w[i], aq, F, T, K, wtype, flags, out=out)
        for _ in range(2):
            g.replay()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.replay()
            e1.record(stream)
            stream.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / n)
        # single hot launch (weights in L2): the latency floor
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            quant_gemm.gemm(w[0], aq, F, T, K, wtype, flags, out=out)
        e0.record(stream)
        for _ in range(20):
            quant_gemm.gemm(w[0], aq, F, T, K, wtype, flags, out=out)
        e1.record(stream)
        stream.synchronize()
        hot = e0.elapsed_time(e1) * 1e3 / 20
    nb = K // 32
    abytes = F * nb * BS[wtype] + T * nb * 36 + 4 * T * F
    return {"type": NAMES[wtype], "T": T, "F": F, "K": K, "us": best, "us_hot_l2": hot,
            "gbs": abytes / best / 1e3, "tops": 2.0 * T * F * K / best / 1e6, "pool": n,
            "path": hex(quant_gemm.last_path())}


def time_prefill(torch, quant_gemm, wtype, T, F, K, flags=0, reps=5, fused_f32=False, prepacked=False):
    """Whole-call time (prepass + tensor-core kernel [+ quantize_q8_1 when fused_f32]), L2 flushed between reps."""
    dev = torch.device("cuda")
    w = make_weights(torch, wtype, F, K, 1, dev)[0]
    x = torch.randn((T, K), device=dev)
    aq = quant_gemm.quantize_q8_1(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if prepacked:
        w = quant_gemm.prepack_weights(w, F, K, wtype)
        flags |= quant_gemm.GEMM_WEIGHTS_PREPACKED
    call = (lambda: quant_gemm.gemm_w4a8(w, x, F, T, K, wtype, flags)) if fused_f32 else \
           (lambda: quant_gemm.gemm(w, aq, F, T, K, wtype, flags))
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    best = min(times)
    return {"type": NAMES[wtype], "T": T, "F": F, "K": K, "us": best, "us_median": sorted(times)[len(times) // 2],
            "tops": 2.0 * T * F * K / best / 1e6, "fused_quantize": fused_f32, "prepacked_weights": prepacked,
            "path": hex(quant_gemm.last_path())}


def time_reference_gpu(torch, wtype, T, F, K, reps=3):
    """The reference's own GPU kernels (oracle/_ref, compiled in place from /root/reference for sm_100a) on the same
    device and the same shapes, as a reported column: gemm_q4_0_q8_1_tile2d (its best decode-style kernel),
    gemm_w4a8_tiled_dp4a (include/) and gemm_quant_kernel (kernels/gemm, one thread per output).  Q4_0 only, like those
    kernels.  Returns {} when oracle/_ref is absent."""
    import qgemm_oracle as qo
    if not qo.have_ref() or wtype != 2:
        return {}
    R = qo.Reference()
    dev = torch.device("cuda")
    w = make_weights(torch, wtype, F, K, 1, dev)[0]
    import quant_gemm
    aq = quant_gemm.quantize_q8_1(torch.randn((T, K), device=dev))
    out = torch.empty((F, T), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    L = R.lib
    cands = {
        "ref_tile2d_us": lambda: L.ref_gpu_gemm_q4_0_tile2d(w.data_ptr(), aq.data_ptr(), out.data_ptr(), F, T, K, None),
        "ref_tiled_dp4a_us": lambda: L.ref_gpu_gemm_w4a8_tiled_dp4a(aq.data_ptr(), w.data_ptr(), out.data_ptr(), T, F, K, None),
        "ref_quant_kernel_us": lambda: L.ref_gpu_gemm_quant(wtype, w.data_ptr(), aq.data_ptr(), out.data_ptr(), F, T, K, None),
    }
    for name, fn in cands.items():
        if name == "ref_tile2d_us" and T > 64:
            continue   # a decode-style kernel: minutes at prefill sizes
        torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        res[name] = best
    return res


def run_vs_reference(out_path):
    """BASELINE configs 1-3 shapes: ours next to the reference's GPU kernels on the same B200 (L2 flushed before every call)."""
    import torch
    import quant_gemm
    rows = []
    for (T, F, K) in [(1, 4096, 4096), (1, 11008, 4096), (8, 11008, 4096), (512, 4096, 4096)]:
        dev = torch.device("cuda")
        w = make_weights(torch, 2, F, K, 1, dev)[0]
        aq = quant_gemm.quantize_q8_1(torch.randn((T, K), device=dev))
        out = torch.empty((F, T), device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        quant_gemm.gemm(w, aq, F, T, K, 2, 0x10, out=out)
        best = 1e30
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            quant_gemm.gemm(w, aq, F, T, K, 2, 0x10, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        r = {"type": "q4_0", "T": T, "F": F, "K": K, "ours_us_cold_l2": best, "path": hex(quant_gemm.last_path())}
        r.update(time_reference_gpu(torch, 2, T, F, K))
        for k in list(r):
            if k.startswith("ref_") and k.endswith("_us"):
                r[k.replace("_us", "_speedup")] = r[k] / best
        rows.append(r)
        print(json.dumps(r), flush=True)
    with open(out_path, "w") as f:
        json.dump({"note": "single cold-L2 calls (CUDA events around one call, L2 flushed before it): launch latency included on both sides",
                   "rows": rows}, f, indent=1)


def run_prefill(out_path):
    import torch
    import quant_gemm
    rows = []
    for (wt, T, F, K, fused) in [(2, 512, 4096, 4096, False), (2, 2048, 4096, 4096, False), (7, 2048, 14336, 4096, True),
                                 (2, 4096, 28672, 8192, False), (8, 512, 4096, 4096, False), (2, 128, 4096, 4096, False),
                                 (2, 256, 4096, 4096, False), (3, 512, 4096, 4096, False), (6, 512, 4096, 4096, False)]:
        r = time_prefill(torch, quant_gemm, wt, T, F, K, fused_f32=fused)
        rows.append(r)
        print(json.dumps(r), flush=True)
        if not fused:
            r = time_prefill(torch, quant_gemm, wt, T, F, K, prepacked=True)
            rows.append(r)
            print(json.dumps(r), flush=True)
    with open(out_path, "w") as f:
        json.dump({"rows": rows}, f, indent=1)


def run(out_path, quick=False, flags=0x10):
    import torch
    import quant_gemm
    rows = []
    types = [2] if quick else [2, 3, 6, 7, 8]
    shapes = [(4096, 4096), (11008, 4096)] + ([] if quick else [(4096, 11008)])
    for wt in types:
        for F, K in shapes:
            for T in ([1, 8] if quick else [1, 2, 4, 8]):
                r = time_shape(torch, quant_gemm, wt, T, F, K, flags)
                rows.append(r)
                print(json.dumps(r), flush=True)
    if not quick:   # beyond BASELINE configs[1]: the skinny path up to the tcgen05 crossover, and a matrix large
        for wt, T, F, K in [(2, 16, 11008, 4096), (2, 32, 11008, 4096), (2, 64, 11008, 4096), (2, 96, 11008, 4096),
                            (2, 1, 32768, 4096), (8, 1, 32768, 4096), (2, 1, 28672, 8192)]:   # enough to hide the launch chain
            r = time_shape(torch, quant_gemm, wt, T, F, K, flags)
            rows.append(r)
            print(json.dumps(r), flush=True)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    with open(out_path, "w") as f:
        json.dump({"peaks": peaks, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/detail.json")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--flags", type=lambda v: int(v, 0), default=0x10)
    ap.add_argument("--prefill", action="store_true")
    ap.add_argument("--vs-reference", action="store_true")
    a = ap.parse_args()
    if a.vs_reference:
        run_vs_reference(a.out)
    elif a.prefill:
        run_prefill(a.out)
    else:
        run(a.out, a.quick, a.flags)
