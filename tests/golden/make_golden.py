"""Generate tests/golden/*.npz by running the REFERENCE's own code.

Run in the dev container only (needs /root/reference to build oracle/_ref/libqgemm_ref.so):
    python tests/golden/make_golden.py
The fixtures are committed; the GPU box and CI read them without the reference.
Every array comes from the unmodified reference functions named in the key.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import datagen  # noqa: E402
import qgemm_oracle as qo  # noqa: E402

STEP4_A = np.array([0.5, 0.3, -0.2, 0.1, 0.4, -0.5, 0.2, 0.3, -0.1, 0.6, 0.2, -0.3, 0.1, 0.4, -0.2, 0.5,
                    0.3, -0.4, 0.2, 0.1, -0.3, 0.5, 0.2, -0.1, 0.4, 0.3, -0.2, 0.1, 0.5, -0.4, 0.3, 0.2], np.float32)
STEP4_W = np.array([0.1, -0.2, 0.3, 0.4, -0.1, 0.2, -0.3, 0.1, 0.2, -0.1, 0.4, -0.2, 0.1, 0.3, -0.4, 0.2,
                    -0.2, 0.3, 0.1, -0.3, 0.2, 0.1, -0.2, 0.4, 0.1, -0.3, 0.2, 0.3, -0.1, 0.2, 0.1, -0.2], np.float32)


def main():
    qo.build_ref()
    R = qo.Reference()
    out = {}
    # 1. the reference's fixed block, tests/step4_w4a8_gemm.cu:51-63
    out["step4_a"], out["step4_w"] = STEP4_A, STEP4_W
    out["step4_a_q8_1"] = R.quantize_row_q8_1_ref(STEP4_A)
    out["step4_w_q4_0"] = R.quantize_row_q4_0_ref(STEP4_W)
    out["step4_dot"] = np.float32(R.vec_dot(qo.Q4_0, out["step4_w_q4_0"], out["step4_a_q8_1"]))
    # 2. test_cpu_ref.cpp:7-66 inputs, quantized by the framework quantizers
    w = ((np.arange(32) % 16) - 8).astype(np.float32)
    a = (np.arange(32) - 16).astype(np.float32)
    out["cpuref_w"], out["cpuref_a"] = w, a
    out["cpuref_w_q4_0"] = R.to_q(qo.Q4_0, w)
    out["cpuref_a_q8_1"] = R.to_q(qo.Q8_1, a)
    out["cpuref_out"] = R.cpu_gemm(qo.Q4_0, out["cpuref_w_q4_0"].reshape(1, 1, 18), out["cpuref_a_q8_1"].reshape(1, 1, 36))
    # 3. seeded multi-block cases, every format, both quantizer families
    T, F, K = 3, 16, 256
    for name, (x, wf) in {"g1": datagen.model_like(T, F, K, seed=7), "g3": datagen.uniform(T, F, K, seed=11)}.items():
        x[1, 32:64] = 0.0   # an all-zero activation block (d = 0, id = 0)
        wf[2, 0:32] = 0.0   # an all-zero weight block
        out[f"{name}_x"], out[f"{name}_w"] = x, wf
        out[f"{name}_a_q8_1_ref"] = R.quantize_row_q8_1_ref(x)
        out[f"{name}_a_q8_1_fw"] = R.to_q(qo.Q8_1, x)
        for wt in qo.WEIGHT_TYPES:
            n = qo.TYPE_NAMES[wt]
            wq = R.to_q(wt, wf)
            out[f"{name}_w_{n}_fw"] = wq
            out[f"{name}_c_{n}_FT"] = R.cpu_gemm(wt, wq, out[f"{name}_a_q8_1_ref"])
        for wt, fn in ((qo.Q4_0, R.quantize_row_q4_0_ref), (qo.Q8_0, R.quantize_row_q8_0_ref)):
            n = qo.TYPE_NAMES[wt]
            wq = fn(wf)
            out[f"{name}_w_{n}_inc"] = wq
            out[f"{name}_c_{n}_TF_inc"] = R.gemm_include(wt, out[f"{name}_a_q8_1_ref"], wq)
            out[f"{name}_deq_{n}_inc"] = R.dequantize(wt, wq)
            # fp32-activation references: gemm_w4a16_reference / gemm_w8a16_reference
            out[f"{name}_c_{n}_TF_f32act"] = R.gemm_f32act_include(wt, x, wq)
    # 4. raw-block fuzz (benchmark_best.cu:31-55 style), every format
    nb = 8
    for wt in qo.WEIGHT_TYPES:
        n = qo.TYPE_NAMES[wt]
        wq = datagen.fuzz_weight_blocks(wt, 8, nb, seed=wt)
        aq = datagen.fuzz_act_blocks(4, nb, seed=wt)
        out[f"fuzz_w_{n}"], out[f"fuzz_a_{n}"] = wq, aq
        out[f"fuzz_c_{n}_FT"] = R.cpu_gemm(wt, wq, aq)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
