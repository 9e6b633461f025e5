"""GPU parity of the chained decode launch (qgemm_gemv_chain, include/qgemm.h): one persistent kernel walking a list
of one-token GEMV steps must give, bit for bit, what the same steps give as separate launches (the reference's form:
one launch per projection, kernels/gemm/gemm_warp_optimized.cuh:377-1210), and both must match the CPU oracle.
Steps that bring fp32 activations are quantized inside the kernel: bytes-equal to quantize_q8_1
(include/quantize.h:165-193) and to the SwiGLU quantizer (kernels/activation/silu.cuh:97-108)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import datagen
import qgemm_oracle as qo

pytestmark = pytest.mark.gpu

TOL_MAXNORM = 1e-5
TOL_NMSE = 1e-10


@pytest.fixture(scope="module")
def qg():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import quant_gemm
    quant_gemm._lib.lib()
    return quant_gemm


@pytest.fixture(scope="module")
def O():
    return qo.Oracle()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def check_c(c_gpu, c_ref, what=""):
    assert np.isfinite(c_gpu).all(), what
    e_max, e_nmse = qo.max_norm_err(c_gpu, c_ref), qo.nmse(c_gpu, c_ref)
    assert e_max <= TOL_MAXNORM and e_nmse <= TOL_NMSE, f"{what}: max-norm {e_max:.3e} nmse {e_nmse:.3e}"


# K -> consumer variant of the chain kernel: 1024 (1 pair/lane, ragged), 2048 (1, full), 3072 (2, ragged), 4096 (2, full),
# 8192 (2 warps per row, full), 11008 (3 pairs/lane, ragged), 12288 (3, full)
CHAIN_SHAPES = [(1024, [300]), (4096, [400, 300, 296]), (2048, [512]), (11008, [320, 310]), (3072, [333]), (8192, [301]),
                (12288, [299]), (4096, [1000])]


def _make_steps(O, wt, shapes, seed):
    steps = []
    for i, (K, Fs) in enumerate(shapes):
        x, _ = datagen.model_like(1, 8, K, seed=seed + i)
        aq = O.quantize_q8_1(x)
        wqs = [O.quantize_weight(wt, datagen.model_like(1, F, K, seed=seed + 31 * i + m)[1]) for m, F in enumerate(Fs)]
        steps.append((K, Fs, aq, wqs))
    return steps


@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
def test_chain_equals_separate_launches_and_oracle(qg, O, wt):
    hs = _make_steps(O, wt, CHAIN_SHAPES, seed=100 + wt)
    dsteps, keep = [], []
    for si, (K, Fs, aq, wqs) in enumerate(hs):
        da, dws = dev(aq), [dev(w) for w in wqs]
        keep.append((da, dws))
        dsteps.append({"weights": dws, "Ms": Fs, "K": K, "act_q": da, "ready": si % 3 == 2})
    chain = qg.GemvChain(dsteps, wt)
    for rep in range(3):   # the arrival counters must return to zero by themselves
        for outs in chain.outs:
            for o in outs:
                o.fill_(float("nan"))
        res = chain()
        assert qg.last_path() == (qg.PATH_GEMV | qg.PATH_CHAINED), "the persistent kernel must have taken this list"
        for (K, Fs, aq, wqs), (da, dws), outs in zip(hs, keep, res):
            for F, wq, dw, o in zip(Fs, wqs, dws, outs):
                c = host(o)
                sep = host(qg.gemm(dw, da, F, 1, K, wt, flags=qg.PATH_GEMV))
                assert (bits(c) == bits(sep)).all(), f"rep {rep} K={K} F={F}: chained != separate launch"
                if rep == 0:
                    check_c(c, O.gemm(wt, aq, wq, layout="FT"), f"chain vs oracle K={K} F={F}")
    sync = next(iter(qg._chain_sync.values()))
    assert int(host(sync).view(np.uint32)[: 2 * chain.n + 1].sum()) == 0, "arrival counters not cleared"


@pytest.mark.parametrize("wt", [qo.Q4_0, qo.Q5_1, qo.Q8_0])
def test_chain_quantizes_fp32_activations_in_kernel(qg, O, wt):
    """A Llama FFN as a real dependent chain: x_q -> [gate, up] -> quantize(silu(gate) * up) -> down -> quantize -> next."""
    K, Fi = 4096, 11008
    x, _ = datagen.model_like(1, 8, K, seed=3)
    aq = O.quantize_q8_1(x)
    mk = lambda F, Kk, s: O.quantize_weight(wt, datagen.model_like(1, F, Kk, seed=s)[1] * (50.0 / np.sqrt(Kk)))   # rows of unit norm: O(1) values through the chain
    w_gate, w_up, w_down, w_next = mk(Fi, K, 11), mk(Fi, K, 12), mk(K, Fi, 13), mk(600, K, 14)
    da, dg, du, dd, dn = dev(aq), dev(w_gate), dev(w_up), dev(w_down), dev(w_next)
    og, ou = torch.empty((Fi, 1), device="cuda"), torch.empty((Fi, 1), device="cuda")
    od, on = torch.empty((K, 1), device="cuda"), torch.empty((600, 1), device="cuda")
    chain = qg.GemvChain([
        {"weights": [dg, du], "Ms": [Fi, Fi], "K": K, "act_q": da, "outs": [og, ou]},
        {"weights": [dd], "Ms": [K], "K": Fi, "act": og.view(-1), "gate": ou.view(-1), "outs": [od]},
        {"weights": [dn], "Ms": [600], "K": K, "act": od.view(-1), "outs": [on]},
    ], wt)
    for rep in range(2):
        for o in (og, ou, od, on):
            o.fill_(float("nan"))
        chain()
        assert qg.last_path() == (qg.PATH_GEMV | qg.PATH_CHAINED)
        # the same dataflow as separate launches of this library: bit-equal
        rg, ru = qg.gemm_group([dg, du], da, [Fi, Fi], 1, K, wt, flags=qg.PATH_GEMV)
        assert (bits(host(og)) == bits(host(rg))).all() and (bits(host(ou)) == bits(host(ru))).all()
        hq = qg.quantize_q8_1_silu_mul(rg.view(1, Fi), ru.view(1, Fi))
        rd = qg.gemm(dd, hq, K, 1, Fi, wt, flags=qg.PATH_GEMV)
        assert (bits(host(od)) == bits(host(rd))).all(), "in-kernel SwiGLU quantizer differs from quantize_q8_1_silu_mul"
        rn = qg.gemm(dn, qg.quantize_q8_1(rd.view(1, K)), 600, 1, K, wt, flags=qg.PATH_GEMV)
        assert (bits(host(on)) == bits(host(rn))).all(), "in-kernel quantizer differs from quantize_q8_1"
    sync = next(iter(qg._chain_sync.values()))
    assert int(host(sync).view(np.uint32)[: 2 * chain.n + 1].sum()) == 0, "arrival counters not cleared"
    # and against the CPU oracle, stage by stage on the GPU's own intermediate values
    cg, cu, cd = host(og).reshape(1, Fi), host(ou).reshape(1, Fi), host(od).reshape(1, K)
    check_c(host(og), O.gemm(wt, aq, w_gate, layout="FT"), "gate")
    hq = host(qg.quantize_q8_1_silu_mul(og.view(1, Fi), ou.view(1, Fi)))   # within one step of the oracle (libm expf), tested elsewhere
    check_c(host(od), O.gemm(wt, hq, w_down, layout="FT"), "down")
    check_c(host(on), O.gemm(wt, O.quantize_q8_1(cd), w_next, layout="FT"), "next")
    assert np.abs(cd).max() > 1e-3 and np.abs(host(on)).max() > 1e-4, "degenerate data"


def test_chains_of_different_lengths_share_one_sync_buffer(qg, O):
    """A short chain with in-kernel quantization followed by a long one on the same stream (same sync buffer): the short
    chain's scratch must not land on the long chain's counters."""
    wt = qo.Q8_0
    K, F = 2048, 320
    x, w = datagen.model_like(1, F, K, seed=8)
    wq = O.quantize_weight(wt, w * (50.0 / np.sqrt(K)))
    dw, dx = dev(wq), dev(x.reshape(-1))
    ref = O.gemm(wt, O.quantize_q8_1(x), wq, layout="FT")
    long_n = 100
    for n in (long_n, 2, long_n, 3, long_n):     # the long one first, so that the shared buffer has its size from the start
        res = qg.gemv_chain([{"weights": [dw], "Ms": [F], "K": K, "act": dx} for _ in range(n)], wt)
        assert qg.last_path() & qg.PATH_CHAINED
        for outs in (res[0], res[-1]):
            check_c(host(outs[0]), ref, f"chain of {n}")


def test_chain_fallbacks_give_the_same_results(qg, O):
    """Lists the persistent kernel declines (fewer rows than CTAs, QGEMM_MS_EXACT, an unaligned row length) run as one
    launch per step -- with a quantize launch in front of fp32 activations -- and give the same numbers."""
    wt = qo.Q4_1
    hs = _make_steps(O, wt, [(4096, [100]), (1056, [400])], seed=9)   # 100 rows < 296 CTAs; K = 1056: 33 blocks (odd)
    for flags in (0, qg.GEMM_MS_EXACT):
        dsteps, keep = [], []
        for K, Fs, aq, wqs in hs:
            da, dws = dev(aq), [dev(w) for w in wqs]
            keep.append((da, dws))
            dsteps.append({"weights": dws, "Ms": Fs, "K": K, "act_q": da})
        res = qg.gemv_chain(dsteps, wt, flags)
        assert not (qg.last_path() & qg.PATH_CHAINED)
        for (K, Fs, aq, wqs), outs in zip(hs, res):
            check_c(host(outs[0]), O.gemm(wt, aq, wqs[0], layout="FT", flags=qo.GEMM_MS_EXACT if flags else 0), "fallback")
    # fp32 activations through the fallback: quantize launch + GEMV
    K, F = 4096, 64
    x, w = datagen.model_like(1, F, K, seed=21)
    wq = O.quantize_weight(wt, w)
    res = qg.gemv_chain([{"weights": [dev(wq)], "Ms": [F], "K": K, "act": dev(x.reshape(-1))}], wt)
    assert not (qg.last_path() & qg.PATH_CHAINED)
    check_c(host(res[0][0]), O.gemm(wt, O.quantize_q8_1(x), wq, layout="FT"), "fallback f32")


def test_chain_llama_layer_in_a_cuda_graph(qg, O):
    """BASELINE configs[1] shapes: one Llama-7B layer (fused q/k/v, o, gate/up, down) as one chained launch, captured in a
    CUDA graph and replayed: bit-equal to the grouped launches, sampled rows equal to the oracle."""
    wt = qo.Q4_0
    g = torch.Generator(device="cuda").manual_seed(77)

    def rand_w(F, K):
        nb = K // 32
        w = torch.randint(0, 256, (F, nb, 18), dtype=torch.uint8, device="cuda", generator=g)
        d = (torch.rand((F, nb), device="cuda", generator=g) * 0.02 + 0.001).to(torch.float16)
        w[:, :, 0:2] = d.view(torch.uint8).view(F, nb, 2)
        return w
    acts = {K: qg.quantize_q8_1(torch.randn((1, K), device="cuda", generator=g)) for K in (4096, 11008)}
    layer = [([rand_w(4096, 4096) for _ in range(3)], 4096), ([rand_w(4096, 4096)], 4096),
             ([rand_w(11008, 4096) for _ in range(2)], 4096), ([rand_w(4096, 11008)], 11008)]
    steps = [{"weights": ws, "Ms": [w.shape[0] for w in ws], "K": K, "act_q": acts[K]} for ws, K in layer]
    chain = qg.GemvChain(steps, wt, qg.GEMM_WEIGHTS_STATIC)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        chain()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            chain()
            chain()
        for outs in chain.outs:
            for o in outs:
                o.fill_(float("nan"))
        for _ in range(3):
            graph.replay()
        stream.synchronize()
    for (ws, K), outs in zip(layer, chain.outs):
        ref = qg.gemm_group(ws, acts[K], [w.shape[0] for w in ws], 1, K, wt, flags=qg.PATH_GEMV)
        for w, o, r in zip(ws, outs, ref):
            assert (bits(host(o)) == bits(host(r))).all()
            rows = np.r_[0:3, w.shape[0] - 3:w.shape[0]]
            check_c(host(o)[rows], O.gemm(wt, host(acts[K]), host(w)[rows], layout="FT"), "layer vs oracle")


def test_chain_longest_list(qg, O):
    wt = qo.Q8_0
    n = qg._lib.lib().qgemm_gemv_chain_max_steps()
    K, F = 1024, 296
    x, w = datagen.model_like(1, F, K, seed=2)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
    da, dw = dev(aq), dev(wq)
    res = qg.gemv_chain([{"weights": [dw], "Ms": [F], "K": K, "act_q": da} for _ in range(n)], wt)
    assert qg.last_path() & qg.PATH_CHAINED
    ref = O.gemm(wt, aq, wq, layout="FT")
    for outs in (res[0], res[n // 2], res[-1]):
        check_c(host(outs[0]), ref, "longest chain")
    # one step more than the kernel takes: one launch per step, same numbers
    res = qg.gemv_chain([{"weights": [dw], "Ms": [F], "K": K, "act_q": da} for _ in range(n + 1)], wt)
    assert not (qg.last_path() & qg.PATH_CHAINED)
    check_c(host(res[-1][0]), ref, "over-long chain")
