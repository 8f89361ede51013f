"""Arithmetic identities the CUDA path relies on to stay bit-identical while doing less work (CPU-only checks)."""
import numpy as np


def test_float_of_double_sqrt_equals_float_sqrt():
    """pgm::sqrt_f64_as_f32 (crux_b200/csrc/pg_math.cuh): (float)sqrt((double)x) -- what distance.c:91 computes -- equals the
    correctly rounded binary32 root for every non-negative float (double rounding is innocuous for sqrt when the wide
    format has >= 2p + 2 bits).  tools/check_sqrt_double_rounding.py walks all 2,139,095,041 bit patterns (~20 s); here:
    every pattern of 64 exponents' worth around the scales the scenes use, the subnormals' edge, and 8 M random ones."""
    rng = np.random.default_rng(1)
    chunks = [np.arange(0x3a000000, 0x3a000000 + (1 << 22), dtype=np.uint32),          # ~5e-4 ...
              np.arange(0x3f000000, 0x3f000000 + (1 << 22), dtype=np.uint32),          # 0.5 ... 1
              np.arange(0x42000000, 0x42000000 + (1 << 22), dtype=np.uint32),          # 32 ...
              np.arange(0, 1 << 22, dtype=np.uint32),                                  # zero and subnormals
              np.arange(0x7f800000 - (1 << 20), 0x7f800001, dtype=np.uint32),          # up to +inf
              rng.integers(0, 0x7f800001, size=1 << 23, dtype=np.uint32)]
    for bits in chunks:
        x = bits.view(np.float32)
        a = np.sqrt(x)
        b = np.sqrt(x.astype(np.float64)).astype(np.float32)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
