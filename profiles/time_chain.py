#!/usr/bin/env python
"""Decode stack of bench.py (Llama-7B shapes, Q4_0 x Q8_1, M = 1, 3.64 GB of weights) timed in its launch forms:

  grouped      one launch per grouped projection (4 per layer, 128 per step) -- the form of rounds 1 / early 2
  chain L      qgemm_gemv_chain over L layers per launch, ready-made q8_1 activations, every step waits for its predecessor
  dataflow L   the same chains as a REAL dependent dataflow: every projection's activations are quantized inside the kernel
               from the fp32 output of the step before it (o <- q, gate/up <- o, down <- silu(gate) * up, next q/k/v <- down);
               compared with the same dataflow as separate launches (GEMV + quantize kernels in between)

CUDA-graph replays, CUDA events on the launch stream.  Usage: python profiles/time_chain.py [--layers 32] [--reps 30]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llama.cpp-quant-gemm_b200"))
import torch  # noqa: E402

import quant_gemm as qg  # noqa: E402

LLAMA7B = [("wq", 4096, 4096), ("wk", 4096, 4096), ("wv", 4096, 4096), ("wo", 4096, 4096),
           ("gate", 11008, 4096), ("up", 11008, 4096), ("down", 4096, 11008)]
WT, FLAGS = 2, 0x10


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--chains", default="1,2,4,8,32")
    ap.add_argument("--dataflow", type=int, default=1)
    ap.add_argument("--prefetch-mb", type=int, default=12)
    ap.add_argument("--out", default=None)
    ap.add_argument("--ready", type=int, default=0, help="diagnostic: mark every chain step QGEMM_INPUTS_READY (no device-wide waits)")
    ap.add_argument("--l2-resident", type=int, default=0, help="diagnostic: every layer uses layer 0's weights (L2-resident after the first pass)")
    ap.add_argument("--grouped", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1234)
    sizes = [F * (K // 32) * 18 for _ in range(args.layers) for _, F, K in LLAMA7B]
    arena = torch.empty(sum((n + 255) // 256 * 256 for n in sizes), dtype=torch.uint8, device=dev)
    mats, off = [], 0
    for _ in range(args.layers):
        for _, F, K in LLAMA7B:
            nb = K // 32
            w = arena[off:off + F * nb * 18].view(F, nb, 18)
            off += (F * nb * 18 + 255) // 256 * 256
            w.copy_(torch.randint(0, 256, (F, nb, 18), dtype=torch.uint8, device=dev, generator=g))
            d = (torch.rand((F, nb), device=dev, generator=g) * 0.02 + 0.001).to(torch.float16)
            w[:, :, 0:2] = d.view(torch.uint8).view(F, nb, 2)
            mats.append((F, K, w))
    arena_end = arena.data_ptr() + arena.numel()
    step_bytes = sum(F * (K // 32) * 18 + (K // 32) * 36 + 4 * F for F, K, _ in mats)
    acts_q = {K: qg.quantize_q8_1(torch.randn((1, K), device=dev, generator=g)) for K in (4096, 11008)}
    outs = [torch.empty((F, 1), device=dev) for F, K, _ in mats]
    groups = []
    for l in range(args.layers):
        b = 7 * l
        groups += [[b, b + 1, b + 2], [b + 3], [b + 4, b + 5], [b + 6]]
    if args.l2_resident:   # 128 steps over the same 9.4 MB matrix: what the consumers can do when HBM is not in the way
        groups = [[0] for _ in groups]
        step_bytes = len(groups) * (4096 * 128 * 18 + 128 * 36 + 4 * 4096)

    def hint(w):
        if args.prefetch_mb > 0:
            qg.hint_next_weights(w, min(args.prefetch_mb << 20, arena_end - w.data_ptr()))

    def grouped():
        for gi, gr in enumerate(groups):
            hint(mats[groups[(gi + 1) % len(groups)][0]][2])
            K = mats[gr[0]][1]
            if len(gr) == 1:
                qg.gemm(mats[gr[0]][2], acts_q[K], mats[gr[0]][0], 1, K, WT, FLAGS, out=outs[gr[0]])
            else:
                qg.gemm_group([mats[i][2] for i in gr], acts_q[K], [mats[i][0] for i in gr], 1, K, WT, FLAGS, outs=[outs[i] for i in gr])

    def make_chains(L, dataflow):
        per, chains = 4 * L, []
        for c0 in range(0, len(groups), per):
            steps = []
            for gi in range(c0, min(c0 + per, len(groups))):
                gr = groups[gi]
                K = mats[gr[0]][1]
                st = {"weights": [mats[i][2] for i in gr], "Ms": [mats[i][0] for i in gr], "K": K, "outs": [outs[i] for i in gr]}
                j = gi % 4
                if not dataflow or gi == 0:
                    st["act_q"] = acts_q[K]
                    st["ready"] = bool(args.ready)
                elif j == 0:
                    st["act"] = outs[gr[0] - 1].view(-1)            # q/k/v <- previous layer's down
                elif j == 1:
                    st["act"] = outs[gr[0] - 3].view(-1)            # o <- q (stands in for the attention output)
                elif j == 2:
                    st["act"] = outs[gr[0] - 1].view(-1)            # gate/up <- o
                else:
                    st["act"], st["gate"] = outs[gr[0] - 2].view(-1), outs[gr[0] - 1].view(-1)   # down <- silu(gate) * up
                steps.append(st)
            chains.append((qg.GemvChain(steps, WT, FLAGS), mats[groups[c0][0]][2]))
        return chains

    def run_chains(chains):
        for ci, (ch, _) in enumerate(chains):
            hint(chains[(ci + 1) % len(chains)][1])
            ch()

    def dataflow_separate():
        """the dataflow of make_chains(dataflow=True) as separate launches: quantize kernels between the GEMVs"""
        for gi, gr in enumerate(groups):
            K = mats[gr[0]][1]
            j = gi % 4
            if gi == 0:
                a = acts_q[K]
            elif j == 0:
                a = qg.quantize_q8_1(outs[gr[0] - 1].view(1, K))
            elif j == 1:
                a = qg.quantize_q8_1(outs[gr[0] - 3].view(1, K))
            elif j == 2:
                a = qg.quantize_q8_1(outs[gr[0] - 1].view(1, K))
            else:
                a = qg.quantize_q8_1_silu_mul(outs[gr[0] - 2].view(1, K), outs[gr[0] - 1].view(1, K))
            hint(mats[groups[(gi + 1) % len(groups)][0]][2])
            if len(gr) == 1:
                qg.gemm(mats[gr[0]][2], a, mats[gr[0]][0], 1, K, WT, FLAGS, out=outs[gr[0]])
            else:
                qg.gemm_group([mats[i][2] for i in gr], a, [mats[i][0] for i in gr], 1, K, WT, FLAGS, outs=[outs[i] for i in gr])

    def calibrate():
        """rescale every matrix's block scales so that the dependent dataflow keeps O(1) magnitudes through all layers"""
        for gi, gr in enumerate(groups):
            K = mats[gr[0]][1]
            j = gi % 4
            src = None if gi == 0 else (outs[gr[0] - 1] if j in (0, 2) else outs[gr[0] - 3] if j == 1 else None)
            if gi == 0:
                a = acts_q[K]
            elif j == 3:
                a = qg.quantize_q8_1_silu_mul(outs[gr[0] - 2].view(1, K), outs[gr[0] - 1].view(1, K))
            else:
                a = qg.quantize_q8_1(src.view(1, K))
            for i in gr:
                F, _, w = mats[i]
                for _ in range(2):
                    qg.gemm(w, a, F, 1, K, WT, 0, out=outs[i])
                    rms = float(outs[i].square().mean().sqrt())
                    d = w[:, :, 0:2].contiguous().view(torch.float16).float() * (1.0 / max(rms, 1e-20))
                    w[:, :, 0:2] = d.clamp(1e-6, 6e4).to(torch.float16).view(torch.uint8).view(F, K // 32, 2)
                qg.gemm(w, a, F, 1, K, WT, 0, out=outs[i])
        torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)

    def timed(fn, name, extra=None):
        with torch.cuda.stream(stream):
            fn()
            stream.synchronize()
            qg.reset_launch_count()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=stream):
                fn()
            launches = qg.launch_count()
            for _ in range(5):
                gr.replay()
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.reps):
                gr.replay()
            e1.record(stream)
            stream.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        r = {"form": name, "us_per_step": round(us, 1), "gbs": round(step_bytes / us / 1e3, 1), "launches_per_step": launches}
        if extra:
            r.update(extra)
        print(json.dumps(r), flush=True)
        return r

    res = [timed(grouped, "grouped launches")] if args.grouped else [timed(lambda: None, "empty")]
    ref = [o.clone() for o in outs]
    for L in [int(x) for x in args.chains.split(",") if x]:
        if L > args.layers:
            continue
        ch = make_chains(L, False)
        r = timed(lambda: run_chains(ch), f"chain, {L} layer(s) per launch")
        torch.cuda.synchronize()
        r["bit_equal_to_grouped"] = all(torch.equal(a, b) for a, b in zip(outs, ref))
        print(json.dumps({"check": r["form"], "bit_equal_to_grouped": r["bit_equal_to_grouped"]}), flush=True)
        res.append(r)
    if args.dataflow:
        calibrate()
        res.append(timed(dataflow_separate, "dependent dataflow, separate launches (GEMV + quantize kernels)"))
        torch.cuda.synchronize()
        ref = [o.clone() for o in outs]
        for L in (1, args.layers):
            ch = make_chains(L, True)
            r = timed(lambda: run_chains(ch), f"dependent dataflow, chain of {L} layer(s), quantize inside the kernel")
            torch.cuda.synchronize()
            r["bit_equal_to_separate"] = all(torch.equal(a, b) for a, b in zip(outs, ref))
            r["out_rms_last_layer"] = float(outs[-1].square().mean().sqrt())
            print(json.dumps({"check": r["form"], "bit_equal_to_separate": r["bit_equal_to_separate"], "rms": r["out_rms_last_layer"]}), flush=True)
            res.append(r)
    if args.out:
        json.dump({"step_bytes": step_bytes, "layers": args.layers, "results": res}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
