"""GPU suite (-m gpu): the CUDA path, called through the C ABI (ctypes on libqgemm_sm100.so via
quant_gemm._lib, device memory from torch), against the CPU oracle on identical seeded inputs.

Gates:
  * integer sumi: bit-exact
  * quantize_q8_1 / weight quantizers: bytes identical to the oracle
  * C: max|dC|/max|C| <= 1e-5 and NMSE <= 1e-10 vs the oracle's CPU-order result; the
    QGEMM_SEQUENTIAL path is additionally bit-identical to the oracle's FMA-order result
    (= what nvcc builds from the reference's GPU kernel) and, when oracle/_ref is present,
    to the reference's GPU kernel itself run on the same device.
"""
import os

import numpy as np
import pytest
import torch

import datagen
import qgemm_oracle as qo

pytestmark = pytest.mark.gpu

TOL_MAXNORM = 1e-5   # north_star: <= 1e-5 relative (normalised) on identical quantized inputs
TOL_NMSE = 1e-10

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.fixture(scope="module")
def qg():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import quant_gemm
    quant_gemm._lib.lib()  # raises if the CUDA library is missing: no silent fallback
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


# ------------------------------------------------------------------------------------------
# quantize_q8_1
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [0, qo.Q81_ROUND_EVEN, qo.Q81_S_FROM_QSUM | qo.Q81_CLAMP127, qo.Q81_CLAMP127])
@pytest.mark.parametrize("shape", [(1, 32), (3, 4096), (7, 11008), (33, 1056), (2, 5, 256)])
def test_quantize_q8_1_bytes(qg, O, flags, shape):
    rng = np.random.default_rng(sum(shape) + flags)
    x = (rng.standard_normal(shape) * rng.choice([1e-3, 1.0, 50.0], size=shape[:-1] + (1,))).astype(np.float32)
    x.reshape(-1)[:32] = 0.0                      # an all-zero block
    x.reshape(-1, 32)[-1] = np.float32(0.5) * np.arange(-16, 16)  # exact .5 ties after scaling by 127/amax=...
    tie = np.zeros(32, np.float32)
    tie[:6] = [127.0, 0.5, 1.5, -0.5, -2.5, 126.5]
    x.reshape(-1, 32)[x.size // 64] = tie         # d = 1 exactly: genuine round-half cases
    got = host(qg.quantize_q8_1(dev(x), flags))
    want = O.quantize_q8_1(x, flags)
    assert got.shape == want.shape == shape[:-1] + (shape[-1] // 32, 36)
    assert (got == want).all()


def test_quantize_q8_1_golden(qg):
    for name in ("g1", "g3"):
        assert (host(qg.quantize_q8_1(dev(G[f"{name}_x"]))) == G[f"{name}_a_q8_1_ref"]).all()
        fw = qo.Q81_S_FROM_QSUM | qo.Q81_CLAMP127
        assert (host(qg.quantize_q8_1(dev(G[f"{name}_x"]), fw)) == G[f"{name}_a_q8_1_fw"]).all()
    assert (host(qg.quantize_q8_1(dev(G["step4_a"]))) == G["step4_a_q8_1"]).all()


def test_quantize_q8_1_unaligned_and_empty(qg, O):
    x = np.random.default_rng(1).standard_normal((5, 96)).astype(np.float32)
    base = dev(np.concatenate([np.zeros(1, np.float32), x.ravel()]))
    view = base[1:].view(5, 96)  # 4-byte aligned only
    assert (host(qg.quantize_q8_1(view)) == O.quantize_q8_1(x)).all()
    assert qg.quantize_q8_1(torch.zeros((0, 64), device="cuda")).shape == (0, 2, 36)


@pytest.mark.parametrize("shape", [(3, 256), (1, 11008), (130, 4096), (2, 5, 64)])
def test_quantize_q8_1_silu_mul_fused(qg, O, shape):
    """quantize_q8_1(silu(x) * gate) in one pass.  Against the oracle (libm expf on the host, CUDA expf on
    the device: the fp32 products may differ in the last bit, so a quantized value may differ by one step and d / s by one
    fp16 unit in a few blocks); and bit for bit against the reference's own silu_mul_forward_f32 run on this GPU followed
    by the plain quantizer, when oracle/_ref is present."""
    rng = np.random.default_rng(sum(shape))
    x = (rng.standard_normal(shape) * 3).astype(np.float32)
    g = rng.standard_normal(shape).astype(np.float32)
    x.reshape(-1)[:32] = 0.0
    got = host(qg.quantize_q8_1_silu_mul(dev(x), dev(g)))
    want = O.quantize_q8_1(O.silu_mul(x, g))
    assert got.shape == want.shape
    gq, wq_ = got[..., 4:].view(np.int8).astype(np.int32), want[..., 4:].view(np.int8).astype(np.int32)
    assert np.abs(gq - wq_).max() <= 1 and (gq != wq_).mean() < 2e-3
    gd, wd = got[..., :4].copy().view(np.float16).astype(np.float32), want[..., :4].copy().view(np.float16).astype(np.float32)
    assert np.allclose(gd, wd, rtol=2e-3, atol=1e-6)
    # dequantized values agree with the fp32 product to within the quantization step
    y = O.silu_mul(x, g).reshape(-1, 32)
    deq = (gq.reshape(-1, 32) * gd.reshape(-1, 2)[:, :1])
    assert np.abs(deq - y).max() <= 0.51 * np.abs(y).max(axis=1, keepdims=True).max() / 127 + 1e-6
    if qo.have_ref():
        R = qo.Reference()
        fn = getattr(R.lib, "ref_gpu_silu_mul_f32", None)
        if fn is not None:
            dx, dg = dev(x), dev(g)
            y_ref = torch.empty_like(dx)
            torch.cuda.synchronize()
            fn(dx.data_ptr(), dg.data_ptr(), y_ref.data_ptr(), dx.numel(), None)
            torch.cuda.synchronize()
            assert (host(qg.quantize_q8_1(y_ref)) == got).all()
    with pytest.raises(RuntimeError):
        qg.quantize_q8_1_silu_mul(dev(x), dev(g[..., :32]))


@pytest.mark.parametrize("shape", [(1, 4096), (5, 11008), (300, 1024), (2, 3, 64)])
def test_quantize_q8_1_rms_norm_fused(qg, O, shape):
    """quantize_q8_1(rms_norm(x) * weight) without the fp32 intermediate: against the oracle's restatement of
    rms_norm_cpu_f32 followed by the quantizer.  1 / rms comes from a double-precision sum of squares on both sides (other
    summation order on the GPU, same value after rounding to fp32 except in rare ties), so bytes agree up to one
    quantization step in a few elements; and within that of the reference's GPU kernel followed by the quantizer."""
    rng = np.random.default_rng(sum(shape))
    x = (rng.standard_normal(shape) * rng.choice([1e-2, 1.0, 20.0], size=shape[:-1] + (1,))).astype(np.float32)
    w = (1.0 + 0.2 * rng.standard_normal(shape[-1])).astype(np.float32)
    got = host(qg.quantize_q8_1_rms_norm(dev(x), dev(w), 1e-5))
    want = O.quantize_q8_1(O.rms_norm(x, w, 1e-5))
    assert got.shape == want.shape
    gq, wq_ = got[..., 4:].view(np.int8).astype(np.int32), want[..., 4:].view(np.int8).astype(np.int32)
    assert np.abs(gq - wq_).max() <= 1 and (gq != wq_).mean() < 2e-3
    gd, wd = got[..., :4].copy().view(np.float16).astype(np.float32), want[..., :4].copy().view(np.float16).astype(np.float32)
    assert np.allclose(gd, wd, rtol=2e-3, atol=1e-6)
    if qo.have_ref():
        R = qo.Reference()
        fn = getattr(R.lib, "ref_gpu_rms_norm_f32", None)
        if fn is not None:
            dx, dw = dev(x.reshape(-1, shape[-1])), dev(w)
            y_ref = torch.empty_like(dx)
            torch.cuda.synchronize()
            fn(dx.data_ptr(), dw.data_ptr(), y_ref.data_ptr(), dx.shape[0], dx.shape[1], 1e-5, None)
            torch.cuda.synchronize()
            rq = host(qg.quantize_q8_1(y_ref)).reshape(got.shape)[..., 4:].view(np.int8).astype(np.int32)
            assert np.abs(gq - rq).max() <= 1 and (gq != rq).mean() < 5e-3   # its sum of squares is fp32, block-reduced
    with pytest.raises(RuntimeError):
        qg.quantize_q8_1_rms_norm(dev(x), dev(w[:32]))


@pytest.mark.parametrize("wt,T,F,K", [(qo.Q4_0, 1, 300, 1024), (qo.Q5_1, 6, 129, 512), (qo.Q4_0, 256, 1000, 2048), (qo.Q8_0, 40, 384, 1024)])
def test_swiglu_down_projection_one_call(qg, O, wt, T, F, K):
    """gemm_w4a8(..., gate=g) = W . quantize_q8_1(silu(x) * g): bit-equal to the fused quantizer followed by the GEMM on
    every path (decode, small batch, tensor cores: there the quantizer writes the operand tiles itself, two launches)."""
    rng = np.random.default_rng(T + F)
    x = (rng.standard_normal((T, K)) * 2).astype(np.float32)
    g = rng.standard_normal((T, K)).astype(np.float32)
    _, w = datagen.model_like(1, F, K, seed=F)
    wq = dev(O.quantize_weight(wt, w))
    dx, dg = dev(x), dev(g)
    qg.reset_launch_count()
    c1 = qg.gemm_w4a8(wq, dx, F, T, K, wtype=wt, gate=dg)
    n1 = qg.launch_count()
    c2 = qg.gemm(wq, qg.quantize_q8_1_silu_mul(dx, dg), F, T, K, wt)
    assert torch.equal(c1, c2)
    assert n1 <= 3      # quantizer (+ the tile padding sliver when T % 128 != 0) + GEMM
    aq = host(qg.quantize_q8_1_silu_mul(dx, dg))
    check_c(host(c1), O.gemm(wt, aq, host(wq), layout="FT"), "swiglu down projection vs oracle on the GPU-quantized input")


@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
def test_weight_quantizers_and_dequantize(qg, O, wt):
    _, w = datagen.model_like(1, 9, 512, seed=wt)
    w[3, 64:96] = 0
    fn = {2: qg.quantize_q4_0, 3: qg.quantize_q4_1, 6: qg.quantize_q5_0, 7: qg.quantize_q5_1, 8: qg.quantize_q8_0}[wt]
    got = host(fn(dev(w)))
    want = O.quantize_weight(wt, w)
    assert (got == want).all()
    deq = host(qg.dequantize(dev(want), 512, wt))
    assert (bits(deq) == bits(O.dequantize(wt, want))).all()
    if wt in (2, 8):
        assert (got == O.quantize_weight(wt, w, "include")).all()


def test_dequantize_q4_0_reference_name(qg, O):
    wq = G["g1_w_q4_0_inc"]
    out = host(qg.dequantize_q4_0(dev(wq), 256))
    assert (bits(out) == bits(G["g1_deq_q4_0_inc"])).all()


# ------------------------------------------------------------------------------------------
# sumi: bit-exact
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
def test_sumi_bit_exact_fuzz(qg, O, wt):
    T, F, nb = 5, 67, 24
    wq = datagen.fuzz_weight_blocks(wt, F, nb, seed=wt)
    aq = datagen.fuzz_act_blocks(T, nb, seed=wt)
    got = host(qg.block_sumi(dev(wq), dev(aq), F, T, nb * 32, wt))
    assert (got == O.gemm_sumi(wt, aq, wq)).all()
    a, w = datagen.extreme_blocks(wt)
    got = host(qg.block_sumi(dev(w), dev(a), 2, 2, 4 * 32, wt))
    assert (got == O.gemm_sumi(wt, a, w)).all()


# ------------------------------------------------------------------------------------------
# GEMM parity
# ------------------------------------------------------------------------------------------
PATHS = {"auto": 0, "gemv": 0x200, "generic": 0x100, "mma": 0x300}


def run_gemm(qg, wt, aq, wq, path, flags=0):
    T, nb, _ = aq.shape
    F = wq.shape[0]
    out = qg.gemm(dev(wq), dev(aq), F, T, nb * 32, wt, flags | PATHS[path])
    return host(out)  # [F, T]


@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
def test_gemm_golden_fixtures(qg, wt):
    n = qo.TYPE_NAMES[wt]
    for name in ("g1", "g3"):
        c = run_gemm(qg, wt, G[f"{name}_a_q8_1_ref"], G[f"{name}_w_{n}_fw"], "auto")
        check_c(c, G[f"{name}_c_{n}_FT"], f"{name} {n}")
    c = run_gemm(qg, wt, G[f"fuzz_a_{n}"], G[f"fuzz_w_{n}"], "auto")
    check_c(c, G[f"fuzz_c_{n}_FT"], f"fuzz {n}")


def test_step4_block_through_gpu(qg):
    c = run_gemm(qg, qo.Q4_0, G["step4_a_q8_1"].reshape(1, 1, 36), G["step4_w_q4_0"].reshape(1, 1, 18), "auto")
    assert abs(float(c[0, 0]) - (-0.335609674)) < 1e-7


@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
@pytest.mark.parametrize("path", ["gemv", "generic"])
@pytest.mark.parametrize("T,F,K", [(1, 256, 4096), (2, 193, 1024), (4, 512, 1024), (8, 130, 2048), (3, 64, 11008),
                                   (11, 96, 512)])
def test_gemm_vs_oracle_model_like(qg, O, wt, path, T, F, K):
    x, w = datagen.model_like(T, F, K, seed=T * 7 + F)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
    c = run_gemm(qg, wt, aq, wq, path)
    check_c(c, O.gemm(wt, aq, wq, layout="FT"), f"{path} {qo.TYPE_NAMES[wt]} {T}x{F}x{K}")
    if path == "generic":  # sequential path == nvcc's FMA order of the reference GPU kernel, bit for bit
        assert (bits(c) == bits(O.gemm(wt, aq, wq, layout="FT", flags=qo.GEMM_FMA))).all()


@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
def test_gemm_fuzz_blocks_and_ms_exact(qg, O, wt):
    T, F, nb = 4, 200, 64
    wq = datagen.fuzz_weight_blocks(wt, F, nb, seed=10 + wt)
    aq = datagen.fuzz_act_blocks(T, nb, seed=10 + wt, const_ds=False)
    for path in ("gemv", "generic"):
        c = run_gemm(qg, wt, aq, wq, path)
        check_c(c, O.gemm(wt, aq, wq, layout="FT"), f"fuzz {path}")
        c = run_gemm(qg, wt, aq, wq, path, flags=qo.GEMM_MS_EXACT)
        check_c(c, O.gemm(wt, aq, wq, layout="FT", flags=qo.GEMM_MS_EXACT), f"fuzz ms_exact {path}")


def test_gemm_include_convention_and_strides(qg, O):
    """C[T,F] (include/ order) through the raw C ABI with explicit strides."""
    from quant_gemm import _lib
    T, F, K = 3, 160, 1024
    x, w = datagen.uniform(T, F, K, seed=3)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(qo.Q4_0, w, "include")
    da, dw = dev(aq), dev(wq)
    out = torch.full((T, F + 5), -7.0, device="cuda")  # padded rows: ldc_t = F + 5
    rc = _lib.lib().qgemm_gemm(2, da.data_ptr(), dw.data_ptr(), out.data_ptr(), T, F, K, F + 5, 1, 0, None, 0,
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    o = host(out)
    check_c(o[:, :F], O.gemm(qo.Q4_0, aq, wq, layout="TF"))
    assert (o[:, F:] == -7.0).all()


def test_gemm_edge_shapes(qg, O):
    # K = 32 (one block), F = 1, odd block counts (generic path), T = 0
    for (T, F, K) in [(1, 1, 32), (2, 3, 96), (5, 7, 160), (1, 1000, 32)]:
        x, w = datagen.model_like(T, F, K, seed=K)
        for wt in qo.WEIGHT_TYPES:
            aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
            check_c(run_gemm(qg, wt, aq, wq, "auto"), O.gemm(wt, aq, wq, layout="FT"), f"edge {T},{F},{K}")
    e = qg.gemm(torch.zeros((4, 2, 18), dtype=torch.uint8, device="cuda"),
                torch.zeros((0, 2, 36), dtype=torch.uint8, device="cuda"), 4, 0, 64, 2)
    assert e.shape == (4, 0)
    with pytest.raises(RuntimeError, match="divisible by 32"):
        qg.gemm_q4_0_q8_1(torch.zeros(18, dtype=torch.uint8, device="cuda"),
                          torch.zeros(36, dtype=torch.uint8, device="cuda"), 1, 1, 33)
    with pytest.raises(RuntimeError, match="shape mismatch"):
        qg.gemm_q4_0_q8_1(torch.zeros(17, dtype=torch.uint8, device="cuda"),
                          torch.zeros(36, dtype=torch.uint8, device="cuda"), 1, 1, 32)


def test_fused_f32_activation_entry(qg, O):
    T, F, K = 6, 256, 2048
    x, w = datagen.model_like(T, F, K, seed=21)
    for wt in (qo.Q4_0, qo.Q5_1):
        wq = O.quantize_weight(wt, w)
        c = host(qg.gemm_w4a8(dev(wq), dev(x), F, T, K, wt))
        check_c(c, O.gemm(wt, O.quantize_q8_1(x), wq, layout="FT"))


@pytest.mark.skipif(not qo.have_ref(), reason="oracle/_ref/libqgemm_ref.so not shipped")
@pytest.mark.parametrize("wt", [qo.Q4_0, qo.Q4_1, qo.Q5_0, qo.Q5_1])
def test_against_reference_gpu_kernels(qg, O, wt):
    """Same bytes through the reference's own GPU kernel (kernels/gemm/gemm_quant_formats.cuh built
    for sm_100a) and ours.  q8_0 is excluded: the reference kernel does a misaligned 4-byte load
    there; include/gemm_w8a8_naive covers it below."""
    R = qo.Reference()
    T, F, K = 4, 300, 1024
    x, w = datagen.model_like(T, F, K, seed=30 + wt)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
    da, dw = dev(aq), dev(wq)
    ref_out = torch.empty((F, T), device="cuda")
    torch.cuda.synchronize()
    rc = R.lib.ref_gpu_gemm_quant(wt, dw.data_ptr(), da.data_ptr(), ref_out.data_ptr(), F, T, K, None)
    assert rc == 0
    r = host(ref_out)
    assert (bits(run_gemm(qg, wt, aq, wq, "generic")) == bits(r)).all()
    check_c(run_gemm(qg, wt, aq, wq, "gemv"), r, "vs reference GPU")


@pytest.mark.skipif(not qo.have_ref(), reason="oracle/_ref/libqgemm_ref.so not shipped")
def test_against_reference_gpu_include_kernels(qg, O):
    R = qo.Reference()
    from quant_gemm import _lib
    T, F, K = 5, 128, 512
    x, w = datagen.uniform(T, F, K, seed=40)
    aq = O.quantize_q8_1(x)
    for wt, fn in ((qo.Q4_0, R.lib.ref_gpu_gemm_w4a8_naive), (qo.Q8_0, R.lib.ref_gpu_gemm_w8a8_naive)):
        wq = O.quantize_weight(wt, w, "include")
        da, dw = dev(aq), dev(wq)
        r = torch.empty((T, F), device="cuda")
        torch.cuda.synchronize()
        fn(da.data_ptr(), dw.data_ptr(), r.data_ptr(), T, F, K, None)
        ours = torch.empty((T, F), device="cuda")
        assert _lib.lib().qgemm_gemm(wt, da.data_ptr(), dw.data_ptr(), ours.data_ptr(), T, F, K, F, 1, 0, None, 0,
                                     torch.cuda.current_stream().cuda_stream) == 0
        check_c(host(ours), host(r), "include naive")
    # reference GPU quantize_q8_1 kernel == our ROUND_EVEN mode
    dx = dev(x)
    y = torch.empty((T, K // 32, 36), dtype=torch.uint8, device="cuda")
    R.lib.ref_gpu_quantize_q8_1(dx.data_ptr(), y.data_ptr(), T * K, None)
    assert (host(y) == host(qg.quantize_q8_1(dx, qo.Q81_ROUND_EVEN))).all()


# ------------------------------------------------------------------------------------------
# full-size configs: size-independent properties (oracle would take minutes)
# ------------------------------------------------------------------------------------------
def test_full_size_decode_properties(qg, O):
    """BASELINE config 1/2 sizes: spot-check rows against the oracle, plus linearity in the
    activation scale and row-permutation equivariance."""
    T, F, K = 8, 11008, 4096
    x, w = datagen.model_like(T, F, K, seed=50)
    aq = O.quantize_q8_1(x)
    dw32 = dev(w)
    for wt in qo.WEIGHT_TYPES:
        fn = {2: qg.quantize_q4_0, 3: qg.quantize_q4_1, 6: qg.quantize_q5_0, 7: qg.quantize_q5_1, 8: qg.quantize_q8_0}[wt]
        dwq = fn(dw32)
        wq = host(dwq)
        c = host(qg.gemm(dwq, dev(aq), F, T, K, wt))
        rows = np.r_[0:8, 5000:5008, F - 8:F]
        check_c(c[rows], O.gemm(wt, aq, wq[rows], layout="FT"), f"full decode {qo.TYPE_NAMES[wt]}")
        perm = np.random.default_rng(wt).permutation(F)
        c2 = host(qg.gemm(dev(wq[perm]), dev(aq), F, T, K, wt))
        assert (bits(c2) == bits(c[perm])).all()
        # T=1 slice equals column 0 of the T=8 run up to accumulation-order noise; sumi identical
        c1 = host(qg.gemm(dwq, dev(aq[:1]), F, 1, K, wt))
        check_c(c1[:, 0], c[:, 0], "T=1 vs T=8")


# ------------------------------------------------------------------------------------------
# prefill path: tcgen05.mma kind::i8 with TMEM accumulators (QGEMM_PATH_TCGEN05)
# ------------------------------------------------------------------------------------------
PATHS["tcgen05"] = 0x400
FOLD_REFSEQ = 0x1000   # QGEMM_FOLD_REFSEQ
MMQ_SHAPES = [(128, 128, 128), (64, 256, 4096), (200, 300, 1056), (130, 129, 32), (512, 384, 2048),
              # small batches on the native kernel run weight-major (tokens on the N side of the MMA: 32 or 64 columns)
              (17, 200, 1024), (32, 129, 512), (48, 384, 2048), (9, 130, 256), (33, 1000, 4096)]


@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
@pytest.mark.parametrize("T,F,K", MMQ_SHAPES[:4])
def test_mmq_sumi_bit_exact(qg, O, wt, T, F, K):
    """The s32 tile each UTCIMMA leaves in TMEM is the reference's integer block sum, bit for bit."""
    nb = K // 32
    wq = datagen.fuzz_weight_blocks(wt, F, nb, seed=wt + T)
    aq = datagen.fuzz_act_blocks(T, nb, seed=wt + F)
    got = host(qg.block_sumi(dev(wq), dev(aq), F, T, K, wt, flags=0x400))
    assert (got == O.gemm_sumi(wt, aq, wq)).all()


@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
@pytest.mark.parametrize("T,F,K", MMQ_SHAPES)
def test_mmq_gemm_vs_oracle(qg, O, wt, T, F, K):
    x, w = datagen.model_like(T, F, K, seed=T + F + K)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
    c = run_gemm(qg, wt, aq, wq, "tcgen05")
    assert qg.last_path() == 0x400
    check_c(c, O.gemm(wt, aq, wq, layout="FT"), f"tcgen05 {qo.TYPE_NAMES[wt]} {T}x{F}x{K}")
    # blocks are folded in order b = 0..nb-1 with the reference GPU kernel's FMA sequence: bit-identical
    # (q4_1 / q5_1: with QGEMM_FOLD_REFSEQ; their default fold associates the same terms differently)
    c = run_gemm(qg, wt, aq, wq, "tcgen05", flags=FOLD_REFSEQ)
    assert (bits(c) == bits(O.gemm(wt, aq, wq, layout="FT", flags=qo.GEMM_FMA))).all()


@pytest.mark.parametrize("wt", [qo.Q4_0, qo.Q5_1, qo.Q8_0])
def test_mmq_fuzz_ms_exact_and_layouts(qg, O, wt):
    from quant_gemm import _lib
    T, F, nb = 96, 200, 24
    wq = datagen.fuzz_weight_blocks(wt, F, nb, seed=77 + wt)
    aq = datagen.fuzz_act_blocks(T, nb, seed=77 + wt, const_ds=False)
    for fl in (0, qo.GEMM_MS_EXACT):
        c = run_gemm(qg, wt, aq, wq, "tcgen05", flags=fl)
        check_c(c, O.gemm(wt, aq, wq, layout="FT", flags=fl), "mmq fuzz")
    # AUTO takes the tensor-core path from 96 tokens up, and the include/ [T,F] layout works too
    c = run_gemm(qg, wt, aq, wq, "auto")
    assert qg.last_path() == 0x400
    da, dw = dev(aq), dev(wq)
    L = _lib.lib()
    nws = L.qgemm_workspace_bytes(wt, T, F, nb * 32, 0)
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    out = torch.empty((T, F), device="cuda")
    assert L.qgemm_gemm(wt, da.data_ptr(), dw.data_ptr(), out.data_ptr(), T, F, nb * 32, F, 1, 0, ws.data_ptr(), nws,
                        torch.cuda.current_stream().cuda_stream) == 0
    assert (bits(host(out).T) == bits(c)).all()
    # a prefill-sized call without scratch is refused (QGEMM_E_WORKSPACE) rather than demoted silently; the streaming
    # path remains available on request, and QGEMM_STREAM_ALLOC lends the scratch
    st = torch.cuda.current_stream().cuda_stream
    assert L.qgemm_gemm(wt, da.data_ptr(), dw.data_ptr(), out.data_ptr(), T, F, nb * 32, F, 1, 0, None, 0, st) == -5
    assert L.qgemm_gemm(wt, da.data_ptr(), dw.data_ptr(), out.data_ptr(), T, F, nb * 32, F, 1, 0x400, None, 0, st) == -5
    assert L.qgemm_gemm(wt, da.data_ptr(), dw.data_ptr(), out.data_ptr(), T, F, nb * 32, F, 1, 0x300, None, 0, st) == 0
    check_c(host(out).T, c, "streaming path without workspace")
    assert L.qgemm_gemm(wt, da.data_ptr(), dw.data_ptr(), out.data_ptr(), T, F, nb * 32, F, 1, 0x80, None, 0, st) == 0
    assert qg.last_path() == 0x400
    check_c(host(out).T, c, "auto with stream-pool scratch")


def test_mmq_full_size_prefill_properties(qg, O):
    """BASELINE config 3 (Q4_0, M=512 N=4096 K=4096): sampled rows against the oracle, bit-identical
    to the sequential path on a slab, and token-permutation equivariance."""
    T, F, K = 512, 4096, 4096
    x, w = datagen.model_like(T, F, K, seed=60)
    dwq = qg.quantize_q4_0(dev(w))
    daq = qg.quantize_q8_1(dev(x))
    aq, wq = host(daq), host(dwq)
    c = host(qg.gemm(dwq, daq, F, T, K, qo.Q4_0))
    assert qg.last_path() == 0x400
    rows = np.r_[0:4, 2047:2051, F - 4:F]
    ref = O.gemm(qo.Q4_0, aq, wq[rows], layout="FT", flags=qo.GEMM_FMA)
    assert (bits(c[rows]) == bits(ref)).all()
    check_c(c[rows], O.gemm(qo.Q4_0, aq, wq[rows], layout="FT"), "prefill vs CPU-order oracle")
    perm = np.random.default_rng(1).permutation(T)
    c2 = host(qg.gemm(dwq, dev(aq[perm]), F, T, K, qo.Q4_0))
    assert (bits(c2) == bits(c[:, perm])).all()


def test_baseline_config0_m1_4096x4096_q4_0_all_rows(qg, O):
    """BASELINE configs[0], exactly: Q4_0 x Q8_1, M=1 N=4096 K=4096, every output against the reference's own CPU
    gemm_w4a8_reference (include/gemm_reference.h:175-222, compiled in place as oracle/_ref) on operands quantized by
    the reference's own quantizers; where oracle/_ref is absent the restated oracle (pinned against it) stands in."""
    T, F, K = 1, 4096, 4096
    x, w = datagen.uniform(T, F, K, seed=7)      # the reference's step tests draw uniform[-1, 1]
    if qo.have_ref():
        R = qo.Reference()
        aq, wq = R.quantize_row_q8_1_ref(x), R.quantize_row_q4_0_ref(w)
        ref = R.gemm_include(qo.Q4_0, aq, wq)          # [T, F]
        assert (aq == O.quantize_q8_1(x)).all() and (wq == O.quantize_weight(qo.Q4_0, w, "include")).all()
    else:
        aq, wq = O.quantize_q8_1(x), O.quantize_weight(qo.Q4_0, w, "include")
        ref = O.gemm(qo.Q4_0, aq, wq, layout="TF")
    assert (host(qg.quantize_q8_1(dev(x))) == aq).all()
    c = host(qg.gemm(dev(wq), dev(aq), F, T, K, qo.Q4_0))     # [F, T]
    assert qg.last_path() == 0x200
    check_c(c.T, ref, "configs[0] vs gemm_w4a8_reference")
    s = host(qg.block_sumi(dev(wq), dev(aq), F, T, K, qo.Q4_0))
    assert (s == O.gemm_sumi(qo.Q4_0, aq, wq)).all()


@pytest.mark.parametrize("wt", [qo.Q4_1, qo.Q5_0, qo.Q8_0])
def test_mmq_full_size_prefill_other_formats(qg, O, wt):
    """BASELINE config 3's shape (M=512 N=4096 K=4096) for the formats the other full-size tests do not cover:
    sampled rows against the oracle (bit-identical in the reference GPU kernel's operation order with
    QGEMM_FOLD_REFSEQ), token-permutation equivariance."""
    T, F, K = 512, 4096, 4096
    x, w = datagen.model_like(T, F, K, seed=61 + wt)
    dwq = {3: qg.quantize_q4_1, 6: qg.quantize_q5_0, 8: qg.quantize_q8_0}[wt](dev(w))
    daq = qg.quantize_q8_1(dev(x))
    aq, wq = host(daq), host(dwq)
    c = host(qg.gemm(dwq, daq, F, T, K, wt))
    assert qg.last_path() == 0x400
    rows = np.r_[0:4, 1000:1004, F - 4:F]
    check_c(c[rows], O.gemm(wt, aq, wq[rows], layout="FT"), f"prefill {qo.TYPE_NAMES[wt]} vs CPU-order oracle")
    cr = host(qg.gemm(dwq, daq, F, T, K, wt, FOLD_REFSEQ))
    assert (bits(cr[rows]) == bits(O.gemm(wt, aq, wq[rows], layout="FT", flags=qo.GEMM_FMA))).all()
    perm = np.random.default_rng(2).permutation(T)
    c2 = host(qg.gemm(dwq, dev(aq[perm]), F, T, K, wt))
    assert (bits(c2) == bits(c[:, perm])).all()


@pytest.mark.parametrize("wt,T,F,K", [(qo.Q4_0, 128, 4096, 4096),    # 32 tiles: every tile cut into 4 segments
                                      (qo.Q5_0, 200, 1000, 11008),   # 16 tiles of 43 raw stages in 9 ragged segments
                                      (qo.Q8_0, 256, 9600, 2048),    # 150 tiles: one whole wave + 2 tiles cut in 4
                                      (qo.Q4_1, 96, 2500, 4096),     # 20 tiles x 7 segments, ragged
                                      (qo.Q5_1, 300, 19072, 1024),   # 447 tiles: three waves + 3 tiles cut in 2
                                      (qo.Q4_0, 640, 4096, 11008),   # 160 tiles: one wave + 12 tiles cut in 12
                                      (qo.Q5_0, 384, 4096, 4096),    # 96 tiles: shared ranges, two units per CTA
                                      (qo.Q4_1, 128, 11008, 2048)])  # 86 tiles of 8 raw stages: shared ranges
def test_mmq_split_k_plans(qg, O, wt, T, F, K):
    """K-split of the last (or only) wave of tiles: the segments of a tile are added in K order by whichever CTA
    arrives last, so the result is deterministic, within the tolerance of the unsplit evaluation (reference
    summation order, QGEMM_FOLD_REFSEQ, which the tests above pin to the oracle bit for bit), the arrival counters
    return to zero for the next call, and sampled rows meet the oracle bar."""
    x, w = _gpu_model_like(T, F, K, seed=T + F + K + wt)
    quant = {2: qg.quantize_q4_0, 3: qg.quantize_q4_1, 6: qg.quantize_q5_0, 7: qg.quantize_q5_1, 8: qg.quantize_q8_0}[wt]
    dwq, daq = quant(w), qg.quantize_q8_1(x)
    cr = qg.gemm(dwq, daq, F, T, K, wt, 0x400 | FOLD_REFSEQ)
    c1 = qg.gemm(dwq, daq, F, T, K, wt, 0x400)
    c2 = qg.gemm(dwq, daq, F, T, K, wt, 0x400)
    c3 = qg.gemm(dwq, daq, F, T, K, wt, 0x400)
    assert torch.equal(c1, c2) and torch.equal(c1, c3)
    assert float((c1 - cr).abs().max()) <= 1e-5 * float(cr.abs().max())
    if wt in (qo.Q4_0, qo.Q5_0, qo.Q8_0):
        assert not torch.equal(c1, cr)      # the plan did cut tiles (their default fold is the reference's otherwise)
    rows = np.r_[0:2, F // 2:F // 2 + 2, F - 2:F]
    ri = torch.from_numpy(rows).cuda()
    check_c(host(c1[ri]), O.gemm(wt, host(daq), host(dwq[ri]), layout="FT"), f"split-K {qo.TYPE_NAMES[wt]} {T}x{F}x{K}")


def test_reference_python_test_file_runs_unchanged(qg):
    """The reference's own python/tests/test_gemm_q4_0.py, staged unmodified under baseline/_ref/ by
    __graft_entry__.build() (git-ignored, travels to the GPU box), run with `quant_gemm` resolving to this package."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    test = os.path.join(root, "baseline", "_ref", "python", "tests", "test_gemm_q4_0.py")
    if not os.path.exists(test):
        pytest.skip("reference test file not staged (needs /root/reference at build time)")
    env = dict(os.environ, PYTHONPATH=os.path.join(root, "llama.cpp-quant-gemm_b200") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", test], capture_output=True, text=True,
                         env=env, cwd=os.path.join(root, "baseline", "_ref", "python"), timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


@pytest.mark.parametrize("wt", [qo.Q4_0, qo.Q5_1, qo.Q8_0])
@pytest.mark.parametrize("T", [256, 200])
def test_fused_quantize_gemm_is_two_launches_and_bit_equal(qg, O, wt, T):
    """The one-call fp32-activation GEMM at tensor-core sizes: quantize_q8_1 writes the operand tiles itself (no
    block_q8_1 round trip), then the GEMM kernel -- two launches (three when T is not a multiple of 128: a sliver
    zeroes the padding rows).  d, s and q must be the values of quantize_row_q8_1_ref: the result equals
    quantize + GEMM bit for bit for every quantizer flavour, and the oracle's on sampled rows."""
    F, K = 384, 1024
    x, w = datagen.model_like(T, F, K, seed=900 + wt + T)
    x[3, 64:96] = 0.0                      # an all-zero block (d = 0)
    x[5, :32] = np.float32(0.5) * np.arange(-16, 16, dtype=np.float32)   # exact .5 ties after scaling
    wq = O.quantize_weight(wt, w)
    dx, dwq = dev(x), dev(wq)
    for qf in (0, qo.Q81_ROUND_EVEN, qo.Q81_S_FROM_QSUM | qo.Q81_CLAMP127):
        qg.reset_launch_count()
        c1 = qg.gemm_w4a8(dwq, dx, F, T, K, wtype=wt, q81_flags=qf)
        assert qg.last_path() == 0x400
        assert qg.launch_count() == (2 if T % 128 == 0 else 3)
        c2 = qg.gemm(dwq, qg.quantize_q8_1(dx, qf), F, T, K, wt)
        assert torch.equal(c1, c2), f"flags {qf}"
    aq = O.quantize_q8_1(x)
    rows = np.r_[0:8, F - 8:F]
    check_c(host(qg.gemm_w4a8(dwq, dx, F, T, K, wtype=wt))[rows], O.gemm(wt, aq, wq[rows], layout="FT"), "fused quantize + GEMM vs oracle")


# ------------------------------------------------------------------------------------------
# W4A16 / W8A16: fp32 activations, no activation quantization (qgemm_gemm_a16)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wt", [qo.Q4_0, qo.Q8_0])
@pytest.mark.parametrize("T,F,K", [(1, 300, 4096), (3, 129, 1024), (8, 64, 11008), (9, 70, 256), (100, 130, 1024), (64, 64, 96)])
def test_gemm_a16_vs_oracle_and_reference_gpu(qg, O, wt, T, F, K):
    """Against the oracle's restatement of gemm_w4a16_reference / gemm_w8a16_reference (include/gemm_reference.h:73-147,
    pinned bit for bit against the reference in test_oracle_golden.py): <= 1e-5 on the fast kernels; the sequential
    kernel is bit-identical to the FMA-order result = what nvcc builds from gemm_w4a16_naive_kernel, and to that
    kernel itself run on this device when oracle/_ref is present."""
    x, w = datagen.model_like(T, F, K, seed=K + T)
    wq = O.quantize_weight(wt, w, "include")
    ref = O.gemm_f32act_dequant(wt, x, wq, layout="TF")
    c = host(qg.gemm_a16(dev(wq), dev(x), T, F, K, wt))
    check_c(c, ref, f"a16 {qo.TYPE_NAMES[wt]} {T}x{F}x{K}")
    cs = host(qg.gemm_a16(dev(wq), dev(x), T, F, K, wt, qg.GEMM_SEQUENTIAL))
    assert (bits(cs) == bits(O.gemm_f32act_dequant(wt, x, wq, layout="TF", flags=qo.GEMM_FMA))).all()
    if qo.have_ref():
        R = qo.Reference()
        fn = getattr(R.lib, "ref_gpu_gemm_w4a16_naive" if wt == qo.Q4_0 else "ref_gpu_gemm_w8a16_naive", None)
        if fn is not None:
            r = torch.empty((T, F), device="cuda")
            dx, dw = dev(x), dev(wq)
            torch.cuda.synchronize()
            fn(dx.data_ptr(), dw.data_ptr(), r.data_ptr(), T, F, K, None)
            assert (bits(host(r)) == bits(cs)).all()


def test_gemm_q4_0_fp32_python_entry(qg, O):
    """python gemm_q4_0_fp32(weight_q [N,K/32,18], activation [M,K]) -> [M,N] (gemm_ops.cu:431-463)."""
    M, N, K = 5, 96, 512
    x, w = datagen.uniform(M, N, K, seed=3)
    wq = O.quantize_weight(qo.Q4_0, w)
    out = qg.gemm_q4_0_fp32(dev(wq), dev(x), M, N, K)
    assert out.shape == (M, N) and out.dtype == torch.float32
    check_c(host(out), O.gemm_f32act_dequant(qo.Q4_0, x, wq, layout="TF"), "gemm_q4_0_fp32")
    with pytest.raises(RuntimeError):
        qg.gemm_q4_0_fp32(dev(wq), dev(x), M, N, K + 32)


def _gpu_model_like(T, F, K, seed):
    """Model-like fp32 operands generated on the device (the full-size shapes are too slow to draw on the host)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    x = torch.randn((T, K), device="cuda", generator=g)
    w = torch.randn((F, K), device="cuda", generator=g) * 0.02
    w[:, ::128] *= 8.0   # an outlier column every 128 elements, like datagen.model_like
    return x, w


def test_full_size_config4_q5_1_with_fused_quantize(qg, O):
    """BASELINE config 4 (Llama-3-8B FFN, Q5_1, M=2048 N=14336 K=4096, quantize_q8_1 of A inside the call):
    the one-call fp32-activation entry equals quantize + GEMM bit for bit; sampled rows against the oracle
    (which quantizes A itself: bytes must agree too)."""
    T, F, K = 2048, 14336, 4096
    x, w = _gpu_model_like(T, F, K, seed=404)
    dwq = qg.quantize_q5_1(w)
    del w
    c1 = qg.gemm_w4a8(dwq, x, F, T, K, wtype=qo.Q5_1)
    assert qg.last_path() == 0x400
    daq = qg.quantize_q8_1(x)
    c2 = qg.gemm(dwq, daq, F, T, K, qo.Q5_1)
    assert torch.equal(c1, c2)
    aq = host(daq)
    assert (aq == O.quantize_q8_1(host(x))).all()
    rows = np.r_[0:3, 7000:7003, F - 3:F]
    ri = torch.from_numpy(rows).cuda()
    wq = host(dwq[ri])
    check_c(host(c1[ri]), O.gemm(qo.Q5_1, aq, wq, layout="FT"), "config 4 vs CPU-order oracle")
    # with the reference's per-block operation sequence the result is the reference GPU kernel's, bit for bit
    c3 = qg.gemm(dwq, daq, F, T, K, qo.Q5_1, FOLD_REFSEQ)
    assert (bits(host(c3[ri])) == bits(O.gemm(qo.Q5_1, aq, wq, layout="FT", flags=qo.GEMM_FMA))).all()


def test_full_size_config5_q4_0_properties(qg, O):
    """BASELINE config 5 (Llama-3-70B FFN, Q4_0, M=4096 N=28672 K=8192) on one GPU: sampled rows x sampled tokens
    against the oracle, and the row-sharded evaluation (what every rank of the N-sharded run computes) equals
    the unsharded one -- bit for bit with the reference's summation order (QGEMM_FOLD_REFSEQ), within the
    tolerance by default (the tiles of the last, ragged wave are summed in K segments, and a shard's last wave
    holds other tiles than the whole matrix's)."""
    T, F, K = 4096, 28672, 8192
    x, w = _gpu_model_like(T, F, K, seed=505)
    dwq = qg.quantize_q4_0(w)
    del w
    daq = qg.quantize_q8_1(x)
    del x
    c = qg.gemm(dwq, daq, F, T, K, qo.Q4_0)
    assert qg.last_path() == 0x400
    cr = qg.gemm(dwq, daq, F, T, K, qo.Q4_0, FOLD_REFSEQ)
    rows = np.r_[0:2, 14335:14337, F - 2:F]
    toks = np.r_[0:8, 2047:2055, T - 8:T]
    ri, ti = torch.from_numpy(rows).cuda(), torch.from_numpy(toks).cuda()
    aq, wq = host(daq[ti]), host(dwq[ri])
    assert (bits(host(cr[ri][:, ti])) == bits(O.gemm(qo.Q4_0, aq, wq, layout="FT", flags=qo.GEMM_FMA))).all()
    check_c(host(c[ri][:, ti]), O.gemm(qo.Q4_0, aq, wq, layout="FT"), "config 5 vs CPU-order oracle")
    scale = float(cr.abs().max())
    assert float((c - cr).abs().max()) <= 1e-5 * scale
    assert not torch.equal(c, cr)          # the default did cut its last wave (64 of 7168 tiles)
    from quant_gemm import sharded
    for world in (2, 8):
        for rank in (0, world - 1):
            f0, f1 = sharded.shard_rows(F, world, rank)
            part = qg.gemm(dwq[f0:f1], daq, f1 - f0, T, K, qo.Q4_0, FOLD_REFSEQ)
            assert torch.equal(part, cr[f0:f1])
            part = qg.gemm(dwq[f0:f1], daq, f1 - f0, T, K, qo.Q4_0)
            assert float((part - cr[f0:f1]).abs().max()) <= 1e-5 * scale


# ------------------------------------------------------------------------------------------
# skinny path: mma.sync m16n8k32 with tokens on N (QGEMM_PATH_MMA), 3 <= T < 64
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
@pytest.mark.parametrize("T,F,K", [(3, 256, 4096), (8, 130, 2048), (5, 33, 256), (8, 512, 8192), (20, 96, 1024), (1, 64, 512),
                                   (4, 100, 11008), (8, 48, 14336)])
def test_mma_skinny_vs_oracle(qg, O, wt, T, F, K):
    x, w = datagen.model_like(T, F, K, seed=T * 3 + F)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
    try:
        c = run_gemm(qg, wt, aq, wq, "mma")
    except RuntimeError as e:  # a forced path refuses shapes whose 16-row tile does not fit in shared memory
        if K >= 8192 and "cannot take this layout" in str(e):
            pytest.skip("tile too large for the skinny path at this K / format")
        raise
    assert qg.last_path() == 0x300
    check_c(c, O.gemm(wt, aq, wq, layout="FT"), f"mma {qo.TYPE_NAMES[wt]} {T}x{F}x{K}")


@pytest.mark.parametrize("wt", qo.WEIGHT_TYPES)
@pytest.mark.parametrize("T,F,K", [(9, 130, 4096), (16, 256, 4096), (31, 100, 4096), (33, 64, 4096), (64, 48, 4096),
                                   (95, 40, 4096), (12, 64, 8192), (40, 33, 8192)])
def test_mma_wide_small_batch_vs_oracle(qg, O, wt, T, F, K):
    """9 <= T < 96 at K = 4096 / 8192: several 8-token tiles per pass (gemv_mma_wide_kernel).  Same answers as the
    passes of 8 it replaces, bit for bit is not required (different K split), the oracle bar is."""
    x, w = datagen.model_like(T, F, K, seed=T * 7 + F)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
    c = run_gemm(qg, wt, aq, wq, "mma")
    assert qg.last_path() == 0x300
    check_c(c, O.gemm(wt, aq, wq, layout="FT"), f"mma wide {qo.TYPE_NAMES[wt]} {T}x{F}x{K}")
    c_auto = run_gemm(qg, wt, aq, wq, "auto")
    if T < 48:
        assert qg.last_path() == 0x300 and (bits(c_auto) == bits(c)).all()
    else:   # with scratch at hand AUTO takes the tcgen05 path from the measured crossover (T = 48 for these shapes)
        assert qg.last_path() == 0x400
        check_c(c_auto, O.gemm(wt, aq, wq, layout="FT"), f"auto small batch {qo.TYPE_NAMES[wt]} {T}x{F}x{K}")
    # include/ convention (C[T,F]) through the same kernel
    from quant_gemm import _lib
    L = _lib.lib()
    da, dw = dev(aq), dev(wq)
    out = torch.empty((T, F), device="cuda")
    assert L.qgemm_gemm(wt, da.data_ptr(), dw.data_ptr(), out.data_ptr(), T, F, K, F, 1, 0x300, None, 0,
                        torch.cuda.current_stream().cuda_stream) == 0
    assert (bits(host(out).T) == bits(c)).all()


@pytest.mark.parametrize("wt", [qo.Q4_0, qo.Q5_1, qo.Q8_0])
def test_mma_skinny_fuzz_and_auto(qg, O, wt):
    T, F, nb = 7, 200, 64
    wq = datagen.fuzz_weight_blocks(wt, F, nb, seed=31 + wt)
    aq = datagen.fuzz_act_blocks(T, nb, seed=31 + wt, const_ds=False)
    for fl in (0, qo.GEMM_MS_EXACT):
        check_c(run_gemm(qg, wt, aq, wq, "mma", flags=fl), O.gemm(wt, aq, wq, layout="FT", flags=fl), "mma fuzz")
    run_gemm(qg, wt, aq, wq, "auto")
    assert qg.last_path() == 0x300        # AUTO: dp4a GEMV for T = 1, mma.sync from 2, tcgen05 from 16 .. 96 (shape dependent)
    run_gemm(qg, wt, aq[:1], wq, "auto")
    assert qg.last_path() == 0x200


# ------------------------------------------------------------------------------------------
# fused all-gather (qgemm_gemm_peers): one GPU stands in for two ranks' memory
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wt,T", [(qo.Q4_0, 1), (qo.Q5_1, 2), (qo.Q8_0, 6)])
def test_peer_store_protocol_single_gpu(qg, O, wt, T):
    """Both 'peers' are buffers on this GPU and both arrival counters are the same word, so every
    launch delivers the `world` arrivals the next launch waits for: exercises the peer stores, the
    done/step/flag accounting and the prologue wait without a second device."""
    from quant_gemm import _lib
    L = _lib.lib()
    F, K, world, steps = 300, 1024, 2, 5
    x, w = datagen.model_like(T, F, K, seed=70 + wt)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
    da, dw = dev(aq), dev(wq)
    bufs = [torch.full((F, T), -1.0, device="cuda") for _ in range(world)]
    flag = torch.zeros(32, dtype=torch.int32, device="cuda")
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ps = _lib.QgemmPeers()
    ps.world, ps.rank = world, 0
    for r in range(world):
        ps.C[r] = bufs[r].data_ptr()
        ps.flag[r] = flag.data_ptr()
    ps.done, ps.step, ps.launches_per_step, ps.launch_index = done.data_ptr(), step.data_ptr(), 1, 0
    ps.wait_index = 0
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(steps):
        assert L.qgemm_gemm_peers(wt, da.data_ptr(), dw.data_ptr(), ps, T, F, K, 1, T, 0, st) == 0
        assert L.qgemm_peer_wait(ps, st) == 0
        assert L.qgemm_peer_step_advance(step.data_ptr(), st) == 0
    ref = O.gemm(wt, aq, wq, layout="FT")
    for b in bufs:
        check_c(host(b), ref, "peer store")
    assert (bits(host(bufs[0])) == bits(host(bufs[1]))).all()
    assert int(flag[0]) == steps * world and int(done[0]) == 0 and int(step[0]) == steps
    # argument checks
    ps.launch_index = 1   # outside launches_per_step
    assert L.qgemm_gemm_peers(wt, da.data_ptr(), dw.data_ptr(), ps, T, F, K, 1, T, 0, st) == -1
    ps.launch_index = 0
    # T > 8 is the tensor-core path: without registered scratch or QGEMM_STREAM_ALLOC it asks for a workspace
    assert L.qgemm_gemm_peers(wt, da.data_ptr(), dw.data_ptr(), ps, 64, F, K, 1, 64, 0, st) == -5


@pytest.mark.parametrize("wt,T,K", [(qo.Q4_0, 200, 1056), (qo.Q5_1, 130, 1056), (qo.Q8_0, 96, 1056), (qo.Q4_0, 256, 1024),
                                    (qo.Q5_1, 130, 2048)])
def test_peer_store_protocol_prefill_single_gpu(qg, O, wt, T, K):
    """Same stand-in as above for the tcgen05 epilogue: the tile goes to both 'ranks' and is bit-identical
    to the plain tensor-core call; launch 1 of each step waits for launch 0 through the wait kernel."""
    from quant_gemm import _lib
    L = _lib.lib()
    F, world, steps = 300, 2, 3   # K = 1056: the prepass kernel; K % 256 == 0: the native-layout kernel
    x, w = datagen.model_like(T, F, K, seed=170 + wt)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
    da, dw = dev(aq), dev(wq)
    bufs = [torch.full((2, F, T), -1.0, device="cuda") for _ in range(world)]
    flag = torch.zeros(32, dtype=torch.int32, device="cuda")
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    pss = []
    for li in range(2):
        ps = _lib.QgemmPeers()
        ps.world, ps.rank = world, 0
        for r in range(world):
            ps.C[r] = bufs[r][li].data_ptr()
            ps.flag[r] = flag.data_ptr()
        ps.done, ps.step, ps.launches_per_step, ps.launch_index = done.data_ptr(), step.data_ptr(), 2, li
        ps.wait_index = li
        pss.append(ps)
    for _ in range(steps):
        for ps in pss:
            assert L.qgemm_gemm_peers(wt, da.data_ptr(), dw.data_ptr(), ps, T, F, K, 1, T, _lib.GEMM_STREAM_ALLOC, st) == 0
            assert qg.last_path() == 0x400
        assert L.qgemm_peer_wait(pss[0], st) == 0
        assert L.qgemm_peer_step_advance(step.data_ptr(), st) == 0
    plain = host(qg.gemm(dw, da, F, T, K, wt, flags=0x400))
    check_c(plain, O.gemm(wt, aq, wq, layout="FT"), "tcgen05 vs oracle")
    for b in bufs:
        for li in range(2):
            if K % 256:
                assert (bits(host(b[li])) == bits(plain)).all()
            else:   # the plain call may split K over the idle SMs (another summation order); the two peer launches share one plan
                check_c(host(b[li]), plain, "peer stores vs the plain call")
                assert (bits(host(b[li])) == bits(host(bufs[0][0]))).all()
    assert int(flag[0]) == steps * 2 * world and int(done[0]) == 0 and int(step[0]) == steps


@pytest.mark.parametrize("wt,T", [(qo.Q4_0, 1), (qo.Q8_0, 5), (qo.Q5_1, 128)])
def test_peer_multicast_destination_single_gpu(qg, O, wt, T):
    """qgemm_peers.C_multicast: one destination stands for every rank (on a real NVLS mapping the switch replicates
    each store).  Here it is an ordinary buffer: it must receive exactly what the plain call computes, the per-rank
    buffers must stay untouched, and the arrival accounting must not change."""
    from quant_gemm import _lib
    L = _lib.lib()
    F, K, world = 300, 1024, 2
    x, w = datagen.model_like(T, F, K, seed=270 + wt)
    aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
    da, dw = dev(aq), dev(wq)
    bufs = [torch.full((F, T), -2.0, device="cuda") for _ in range(world)]
    mc = torch.full((F, T), -1.0, device="cuda")
    flag = torch.zeros(32, dtype=torch.int32, device="cuda")
    done = torch.zeros(1, dtype=torch.int32, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ps = _lib.QgemmPeers()
    ps.world, ps.rank = world, 0
    for r in range(world):
        ps.C[r] = bufs[r].data_ptr()
        ps.flag[r] = flag.data_ptr()
    ps.done, ps.step, ps.launches_per_step, ps.launch_index, ps.wait_index = done.data_ptr(), step.data_ptr(), 1, 0, 0
    ps.C_multicast = mc.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        assert L.qgemm_gemm_peers(wt, da.data_ptr(), dw.data_ptr(), ps, T, F, K, 1, T, _lib.GEMM_STREAM_ALLOC, st) == 0
        path = qg.last_path()
        assert L.qgemm_peer_wait(ps, st) == 0
        assert L.qgemm_peer_step_advance(step.data_ptr(), st) == 0
    plain = host(qg.gemm(dw, da, F, T, K, wt, flags=path))
    if T <= 8:
        assert (bits(host(mc)) == bits(plain)).all()
    else:   # the plain tensor-core call may split K over the idle SMs (another summation order); the two peer launches share one plan
        check_c(host(mc), plain, "multicast destination vs the plain call")
    check_c(host(mc), O.gemm(wt, aq, wq, layout="FT"), "multicast destination vs oracle")
    for b in bufs:
        assert (host(b) == -2.0).all()
    assert int(flag[0]) == 2 * world and int(done[0]) == 0


# ------------------------------------------------------------------------------------------
# grouped decode launch (fused q/k/v, gate/up): identical to separate calls
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wt,T,K", [(qo.Q4_0, 1, 4096), (qo.Q5_1, 2, 1024), (qo.Q8_0, 1, 11008), (qo.Q4_1, 8, 2048)])
def test_gemm_group_equals_separate_calls(qg, O, wt, T, K):
    Fs = [256, 100, 37]          # sizes that are not tile multiples: tiles must break at matrix boundaries
    x, _ = datagen.model_like(T, 8, K, seed=5)
    aq = O.quantize_q8_1(x)
    wqs = [O.quantize_weight(wt, datagen.model_like(1, F, K, seed=10 + i)[1]) for i, F in enumerate(Fs)]
    da = dev(aq)
    dws = [dev(w) for w in wqs]
    outs = qg.gemm_group(dws, da, Fs, T, K, wt)
    for F, wq, dw, o in zip(Fs, wqs, dws, outs):
        c = host(o)
        check_c(c, O.gemm(wt, aq, wq, layout="FT"), "group vs oracle")
        sep = host(qg.gemm(dw, da, F, T, K, wt, flags=0x200))
        assert (bits(c) == bits(sep)).all()
    # the L2 hint as an argument of the call: same numbers (it is a pure performance hint), and nothing is left behind
    hinted = qg.gemm(dws[0], da, Fs[0], T, K, wt, flags=0x200, next_weight_q=dws[1])
    assert (bits(host(hinted)) == bits(host(qg.gemm(dws[0], da, Fs[0], T, K, wt, flags=0x200)))).all()
    # a group of one matrix is a plain GEMV
    one = qg.gemm_group(dws[:1], da, Fs[:1], T, K, wt)
    assert (bits(host(one[0])) == bits(host(qg.gemm(dws[0], da, Fs[0], T, K, wt, flags=0x200)))).all()
    # T > 8 falls back to one launch per matrix, same answers
    x2, _ = datagen.model_like(12, 8, K, seed=6)
    aq2 = O.quantize_q8_1(x2)
    outs = qg.gemm_group(dws, dev(aq2), Fs, 12, K, wt)
    for wq, o in zip(wqs, outs):
        check_c(host(o), O.gemm(wt, aq2, wq, layout="FT"), "group fallback")


def test_prepacked_weights_give_identical_results(qg, O):
    """qgemm_prepack_weights + QGEMM_WEIGHTS_PREPACKED: same bits as the per-call path, no weight prepass."""
    T, F, K = 200, 300, 1056
    x, w = datagen.model_like(T, F, K, seed=88)
    aq = O.quantize_q8_1(x)
    for wt in (qo.Q4_0, qo.Q5_1, qo.Q8_0):
        wq = O.quantize_weight(wt, w)
        dw, da = dev(wq), dev(aq)
        packed = qg.prepack_weights(dw, F, K, wt)
        ref = host(qg.gemm(dw, da, F, T, K, wt, flags=0x400))
        qg.reset_launch_count()
        got = host(qg.gemm(packed, da, F, T, K, wt, flags=qg.GEMM_WEIGHTS_PREPACKED))
        assert qg.last_path() == 0x400 and qg.launch_count() == 2   # activation repack + the MMA kernel only
        assert (bits(got) == bits(ref)).all()
        check_c(got, O.gemm(wt, aq, wq, layout="FT"), "prepacked")
    with pytest.raises(RuntimeError):
        qg.gemm(packed, da, F, T, K, qo.Q8_0, flags=qg.GEMM_WEIGHTS_PREPACKED | 0x200)   # decode path cannot read it
