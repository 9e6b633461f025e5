"""The remaining quantize_q8_1 flavours of the reference (SURVEY 8 row A2'):

  * quantize_fp16_to_q8_1_smem, the in-kernel quantizer of gemm_q4_0_fp16_fused (kernels/gemm/gemm_fused.cuh:76-143): fp16
    input, pairwise tree sum for s, 1/d from the fp16-rounded d, int8 narrowing before a +-127 clamp;
  * the JSON spec's all-zero block (schemas/definitions/quantization/quantize_q8_1.json: d = 1.0 when amax == 0).

CPU part: the oracle's restatement against an independent numpy computation.  GPU part: the product kernel against the
oracle, byte for byte, and both against the reference's own device function run on the same GPU (oracle/_ref)."""
from __future__ import annotations

import numpy as np
import pytest

import qgemm_oracle as qo


def _blocks(seed, nblocks=257):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((nblocks, 32)).astype(np.float32)
    x[0] = 0.0                                   # all-zero block
    x[1] *= 1e-4                                 # fp16-subnormal d: the reference's int8 narrowing wraps here
    x[2] *= 3e-4
    x[3] *= 1e-6                                 # d underflows to fp16 zero
    x[4] = np.linspace(-4, 4, 32)                # exact .5 ties after scaling are likely
    x[5] *= 1e3
    x[6, :] = 0.25
    return x


def _numpy_fused(x16):
    """Independent restatement of gemm_fused.cuh:76-143 on fp16 input, vectorised numpy (fp32 arithmetic)."""
    v = x16.astype(np.float32)
    t = v.copy()
    for w in (16, 8, 4, 2, 1):
        t[:, :w] = t[:, :w] + t[:, w:2 * w]
    s = t[:, 0]
    amax = np.abs(v).max(axis=1)
    d = (amax / np.float32(127.0)).astype(np.float32)
    dh = d.astype(np.float16).astype(np.float32)
    with np.errstate(divide="ignore"):
        idv = np.where(dh != 0, np.float32(1.0) / dh, np.float32(0)).astype(np.float32)
    sv = (v * idv[:, None]).astype(np.float32)
    r = np.where(sv >= 0, np.floor(sv + np.float32(0.5)), np.ceil(sv - np.float32(0.5)))      # roundf: ties away from zero
    # floor(x + 0.5) can be off by one when x + 0.5 rounds up in fp32; roundf itself never is: redo those exactly
    frac = np.abs(sv) - np.floor(np.abs(sv))
    r = np.where(frac < 0.5, np.sign(sv) * np.floor(np.abs(sv)), np.sign(sv) * (np.floor(np.abs(sv)) + 1))
    q = r.astype(np.int64).astype(np.int32)
    q = ((q & 0xff) ^ 0x80) - 0x80               # (int8_t) narrowing
    q = np.clip(q, -127, 127).astype(np.int8)
    out = np.zeros((v.shape[0], 36), dtype=np.uint8)
    out[:, 0:2] = d.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 2:4] = s.astype(np.float16).view(np.uint8).reshape(-1, 2)
    out[:, 4:] = q.view(np.uint8)
    return out


def test_oracle_fused_f16_flavour_matches_numpy():
    O = qo.Oracle()
    x16 = _blocks(1).astype(np.float16)
    got = O.quantize_q8_1_f16(x16, qo.Q81_FUSED_F16)
    want = _numpy_fused(x16)
    assert got.shape == (x16.shape[0], 1, 36)
    assert (got.reshape(-1, 36) == want).all()
    # the narrowing really fires in the tiny-amplitude blocks (otherwise this test pins nothing about it)
    plain = O.quantize_q8_1(x16.astype(np.float32), qo.Q81_CLAMP127 | qo.Q81_TREE_SUM).reshape(-1, 36)
    assert (plain[1:3, 4:] != want[1:3, 4:]).any()
    # the tree sum differs from the sequential sum in the last bits for some blocks, never by more than rounding
    seq = O.quantize_q8_1(x16.astype(np.float32), qo.Q81_CLAMP127).reshape(-1, 36)
    s_tree = want[:, 2:4].copy().view(np.float16).astype(np.float32)
    s_seq = seq[:, 2:4].copy().view(np.float16).astype(np.float32)
    assert np.allclose(s_tree, s_seq, rtol=2e-3, atol=1e-3)


def test_oracle_zero_block_d1_and_f16_entry():
    O = qo.Oracle()
    z = np.zeros((1, 64), dtype=np.float32)
    z[0, 32:] = np.arange(32) - 10
    got = O.quantize_q8_1(z, qo.Q81_ZERO_D1).reshape(-1, 36)
    assert got[0, 0:2].copy().view(np.float16)[0] == np.float16(1.0) and (got[0, 2:] == 0).all()
    assert (got[1] == O.quantize_q8_1(z).reshape(-1, 36)[1]).all()          # non-zero blocks unchanged
    assert (O.quantize_q8_1(z).reshape(-1, 36)[0] == 0).all()               # default: d = 0 like the reference code
    x16 = _blocks(3, 40).astype(np.float16)
    for flags in (0, qo.Q81_ROUND_EVEN, qo.Q81_CLAMP127 | qo.Q81_S_FROM_QSUM, qo.Q81_TREE_SUM):
        assert (O.quantize_q8_1_f16(x16, flags) == O.quantize_q8_1(x16.astype(np.float32), flags)).all()


FLAVOURS = [qo.Q81_FUSED_F16, qo.Q81_TREE_SUM, qo.Q81_ID_FROM_HALF_D, qo.Q81_ZERO_D1, qo.Q81_ZERO_D1 | qo.Q81_ROUND_EVEN,
            qo.Q81_FUSED_F16 | qo.Q81_ZERO_D1, qo.Q81_TREE_SUM | qo.Q81_S_FROM_QSUM | qo.Q81_CLAMP127]


@pytest.mark.gpu
@pytest.mark.parametrize("flags", FLAVOURS)
def test_gpu_flavours_match_oracle_bytes(flags):
    import torch
    import quant_gemm as qg
    O = qo.Oracle()
    x = _blocks(7 + flags, 1031).reshape(1, -1)
    got = qg.quantize_q8_1(torch.from_numpy(x).cuda(), flags)
    torch.cuda.synchronize()
    assert (got.cpu().numpy() == O.quantize_q8_1(x, flags)).all()
    x16 = x.astype(np.float16)
    got = qg.quantize_q8_1(torch.from_numpy(x16).cuda(), flags)
    torch.cuda.synchronize()
    assert (got.cpu().numpy() == O.quantize_q8_1_f16(x16, flags)).all()


@pytest.mark.gpu
def test_gpu_f16_input_default_flags_and_reference_device_function():
    """fp16 input with the default arithmetic == the fp32 entry on the widened values; and with QGEMM_Q81_FUSED_F16 the bytes
    are those the reference's quantize_fp16_to_q8_1_smem writes on this GPU (which also pins the oracle's restatement)."""
    import torch
    import quant_gemm as qg
    O = qo.Oracle()
    x16 = _blocks(11, 4099).astype(np.float16).reshape(1, -1)
    d16 = torch.from_numpy(x16).cuda()
    a = qg.quantize_q8_1(d16)
    b = qg.quantize_q8_1(d16.float())
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    if not qo.have_ref():
        pytest.skip("oracle/_ref not built")
    R = qo.Reference()
    fn = getattr(R.lib, "ref_gpu_quantize_fp16_to_q8_1_smem", None)
    if fn is None:
        pytest.skip("oracle/_ref predates the fused doorway")
    nblocks = x16.size // 32
    y_ref = torch.zeros((nblocks, 36), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert fn(d16.data_ptr(), y_ref.data_ptr(), nblocks, None) == 0
    torch.cuda.synchronize()
    ref = y_ref.cpu().numpy()
    ours = qg.quantize_q8_1(d16, qg.Q81_FUSED_F16).cpu().numpy().reshape(-1, 36)
    assert (ours == ref).all(), "product kernel differs from the reference's device function"
    assert (O.quantize_q8_1_f16(x16, qo.Q81_FUSED_F16).reshape(-1, 36) == ref).all(), "oracle restatement differs from the reference"


@pytest.mark.gpu
def test_gemm_q4_0_fp16_fused_python_entry():
    import torch
    import datagen
    import quant_gemm as qg
    O = qo.Oracle()
    N, M, K = 5, 300, 2048
    x, w = datagen.model_like(N, M, K, seed=4)
    wq = O.quantize_weight(qo.Q4_0, w)
    x16 = x.astype(np.float16)
    out = qg.gemm_q4_0_fp16_fused(torch.from_numpy(wq).cuda(), torch.from_numpy(x16).cuda(), M, N, K)
    torch.cuda.synchronize()
    ref = O.gemm(qo.Q4_0, O.quantize_q8_1_f16(x16, qo.Q81_FUSED_F16), wq, layout="FT")
    assert qo.max_norm_err(out.cpu().numpy(), ref) <= 1e-5
