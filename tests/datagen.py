"""Seeded synthetic inputs for the parity tests and the bench.

G1 "model-like": weights ~ N(0, 0.02) with an 8x outlier every 128th element, activations ~ N(0, 1).
G2 "raw-block fuzz": every nibble / qh bit / int8 value including -128, d_w in {0.1..1.0},
   ds = (0.1, 1.0) -- the generator of the reference's tests/benchmark_best.cu:31-55, reaching
   sumi extremes real quantizers never emit.
G3 uniform[-1, 1]: the reference's step tests (tests/step4_w4a8_gemm.cu:130-150).
"""
import numpy as np

BLOCK_BYTES = {2: 18, 3: 20, 6: 22, 7: 24, 8: 34, 9: 36}


def model_like(T, F, K, seed=0):
    rng = np.random.default_rng(seed)
    w = (rng.standard_normal((F, K)) * 0.02).astype(np.float32)
    w.reshape(-1)[::128] *= 8.0
    x = rng.standard_normal((T, K)).astype(np.float32)
    return x, w


def uniform(T, F, K, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1, 1, (T, K)).astype(np.float32), rng.uniform(-1, 1, (F, K)).astype(np.float32))


def _half_bytes(vals):
    return np.asarray(vals, dtype=np.float16).view(np.uint8)


def fuzz_weight_blocks(wtype, F, nb, seed=0):
    rng = np.random.default_rng(seed)
    bs = BLOCK_BYTES[wtype]
    blk = rng.integers(0, 256, size=(F, nb, bs), dtype=np.uint8)
    d = (0.1 * rng.integers(1, 11, size=(F, nb))).astype(np.float16)
    blk[:, :, 0:2] = d.view(np.uint8).reshape(F, nb, 2)
    if wtype in (3, 7):  # m
        m = (rng.uniform(-1, 1, size=(F, nb))).astype(np.float16)
        blk[:, :, 2:4] = m.view(np.uint8).reshape(F, nb, 2)
    return blk


def fuzz_act_blocks(T, nb, seed=0, const_ds=True):
    rng = np.random.default_rng(seed + 1000)
    blk = rng.integers(0, 256, size=(T, nb, 36), dtype=np.uint8)
    if const_ds:
        blk[:, :, 0:2] = _half_bytes([0.1])
        blk[:, :, 2:4] = _half_bytes([1.0])
    else:
        d = rng.uniform(0.001, 0.1, size=(T, nb)).astype(np.float16)
        s = rng.uniform(-4, 4, size=(T, nb)).astype(np.float16)
        blk[:, :, 0:2] = d.view(np.uint8).reshape(T, nb, 2)
        blk[:, :, 2:4] = s.view(np.uint8).reshape(T, nb, 2)
    return blk


def extreme_blocks(wtype, nb=4):
    """All-max weights against all -128 / +127 activations: |sumi| at its bound."""
    bs = BLOCK_BYTES[wtype]
    w = np.full((2, nb, bs), 0xFF, dtype=np.uint8)
    if wtype == 8:
        w[0, :, 2:] = 0x80  # -128
        w[1, :, 2:] = 0x7F
    w[:, :, 0:2] = _half_bytes([1.0])
    if wtype in (3, 7):
        w[:, :, 2:4] = _half_bytes([-0.5])
    a = np.zeros((2, nb, 36), dtype=np.uint8)
    a[0, :, 4:] = 0x80
    a[1, :, 4:] = 0x7F
    a[:, :, 0:2] = _half_bytes([0.5])
    a[:, :, 2:4] = _half_bytes([2.0])
    return a, w
