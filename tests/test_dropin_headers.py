"""The C++ drop-in surface: a program written against the reference's header names and
signatures (tests/cpp/dropin_check.cu) builds against this repo's headers + libqgemm_sm100.so
(CPU test), and on a B200 its results match the oracle (GPU test)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "llama.cpp-quant-gemm_b200", "lib")
EXE = os.path.join(ROOT, "tests", "cpp", "dropin_check")


def build_exe():
    src = os.path.join(ROOT, "tests", "cpp", "dropin_check.cu")
    lib = os.path.join(LIBDIR, "libqgemm_sm100.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", ROOT,
                               "-I", os.path.join(ROOT, "include"), src, "-o", EXE, "-L", LIBDIR, "-lqgemm_sm100",
                               "-Xlinker", "-rpath", "-Xlinker", LIBDIR])
    return EXE


def test_dropin_program_builds_and_static_checks_pass():
    exe = build_exe()
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_reference_entry_point_names_are_all_present():
    """Every hot-path launcher name of the reference's headers exists in ours."""
    names = {
        "include/gemm_cuda_naive.cuh": ["gemm_w4a8_naive", "gemm_w8a8_naive", "gemm_w4a16_naive", "gemm_w8a16_naive"],
        "include/gemm_cuda_tiled.cuh": ["gemm_w4a8_tiled"],
        "include/gemm_cuda_dp4a.cuh": ["gemm_w4a8_dp4a", "gemm_w8a8_dp4a", "gemm_w4a8_tiled_dp4a", "gemm_w4a8_vectorized_dp4a"],
        "include/quantize.h": ["quantize_q4_0_cuda", "quantize_q8_0_cuda", "quantize_q8_1_cuda"],
        "include/llama_adapter.h": ["gemm_w4a8_from_ggml", "gemm_w4a16_from_ggml", "validate_tensor_types", "extract_dims_from_tensor"],
        "kernels/gemm/gemm_quant_formats.cuh": ["gemm_q4_0_q8_1", "gemm_q4_1_q8_1", "gemm_q5_0_q8_1", "gemm_q5_1_q8_1", "gemm_q8_0_q8_1"],
        "kernels/gemm/gemm_warp_optimized.cuh": ["gemm_q4_0_q8_1_warp", "gemm_q4_0_q8_1_warp_v2", "gemm_q4_0_q8_1_warp_prefetch",
                                                 "gemm_q4_0_q8_1_warp_multirow", "gemm_q4_0_q8_1_warp_multirow8", "gemm_q4_0_q8_1_smem",
                                                 "gemm_q4_0_q8_1_smem_large", "gemm_q4_0_q8_1_vec", "gemm_q4_0_q8_1_tile2d",
                                                 "gemm_q4_0_q8_1_tile2d_n8", "gemm_q4_0_q8_1_tile2d_k256", "gemm_q4_0_q8_1_tile2d_r8",
                                                 "gemm_q4_0_q8_1_tile2d_large"],
        "kernels/gemm/gemm_async_copy.cuh": ["gemm_q4_0_q8_1_async"],
        "kernels/gemm/gemm_vectorized.cuh": ["gemm_q4_0_q8_1_vec_safe", "gemm_q4_0_q8_1_vec_float4"],
        "kernels/gemm/gemm_fused.cuh": ["gemm_q4_0_fp16_fused"],
    }
    for path, fns in names.items():
        text = open(os.path.join(ROOT, path)).read()
        for fn in fns:
            assert fn in text, f"{fn} missing from {path}"


@pytest.mark.gpu
def test_dropin_program_matches_oracle(tmp_path):
    import datagen
    import qgemm_oracle as qo
    exe = build_exe()
    O = qo.Oracle()
    T, F, K = 96, 160, 1024
    x, w = datagen.model_like(T, F, K, seed=99)
    wq = {n: O.quantize_weight(t, w) for n, t in (("q4_0", qo.Q4_0), ("q5_1", qo.Q5_1), ("q8_0", qo.Q8_0))}
    x.tofile(tmp_path / "x.f32")
    x.astype(np.float16).tofile(tmp_path / "x.f16")
    for n, q in wq.items():
        q.tofile(tmp_path / f"w_{n}.bin")
    (tmp_path / "dims.txt").write_text(f"{T} {F} {K}\n")
    out = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr
    assert "launcher path 0x400" in out.stdout  # T = 96: the plain launcher took the tcgen05 path (pool scratch)
    assert "last path 0x400" in out.stdout      # and again with scratch registered by the caller
    aq = np.fromfile(tmp_path / "a_q8_1.bin", dtype=np.uint8).reshape(T, K // 32, 36)
    assert (aq == O.quantize_q8_1(x, qo.Q81_ROUND_EVEN)).all()       # include/quantize.h GPU semantics
    ref = {n: O.gemm(t, aq, wq[n], layout="FT") for n, t in (("q4_0", qo.Q4_0), ("q5_1", qo.Q5_1), ("q8_0", qo.Q8_0))}

    def load(name, shape):
        return np.fromfile(tmp_path / name, dtype=np.float32).reshape(shape)

    # fp32-activation entries against the oracle's restatement of gemm_w4a16_reference / gemm_w8a16_reference
    for name, wt, key in [("c_w4a16.f32", qo.Q4_0, "q4_0"), ("c_w8a16.f32", qo.Q8_0, "q8_0"), ("c_adapter16.f32", qo.Q4_0, "q4_0"),
                          ("c_hook16.f32", qo.Q4_0, "q4_0")]:
        r16 = O.gemm_f32act_dequant(wt, x, wq[key], layout="TF")
        assert qo.max_norm_err(load(name, (T, F)), r16) <= 1e-5, name
    assert qo.max_norm_err(load("c_hook.f32", (T, F)).T, ref["q4_0"]) <= 1e-5
    # gemm_q4_0_fp16_fused: half activations quantized with the arithmetic of the reference's in-kernel quantizer
    a16 = O.quantize_q8_1_f16(x.astype(np.float16), qo.Q81_FUSED_F16)
    assert qo.max_norm_err(load("c_fused16.f32", (F, T)), O.gemm(qo.Q4_0, a16, wq["q4_0"], layout="FT")) <= 1e-5
    for name, key, transposed in [("c_inc_q4_0.f32", "q4_0", True), ("c_inc_q4_0_b.f32", "q4_0", True),
                                  ("c_inc_q8_0.f32", "q8_0", True), ("c_ggml_q4_0.f32", "q4_0", False),
                                  ("c_ggml_q5_1.f32", "q5_1", False), ("c_tile2d.f32", "q4_0", False),
                                  ("c_async.f32", "q4_0", False), ("c_adapter.f32", "q4_0", True),
                                  ("c_big.f32", "q4_0", False)]:
        c = load(name, (T, F) if transposed else (F, T))
        c = c.T if transposed else c
        assert qo.max_norm_err(c, ref[key]) <= 1e-5, name
