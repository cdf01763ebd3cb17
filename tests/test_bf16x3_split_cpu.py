"""The tensor-core kernel's claim of FP32-grade gradients rests on two numerical facts, checked here on the CPU with a
numpy restatement of csrc/random_tc.cu:split3 (round-to-nearest-even bf16 conversion of the running residual):

  1. the three-part split is EXACT: x == b1 + b2 + b3 for every float32 x away from the subnormal range and from overflow
     (|x| < 2^127; positions are O(1..1e3));
  2. the six part products the kernel keeps -- (1,3) (3,1) (2,2) (1,2) (2,1) (1,1) -- reproduce a float64 dot product
     of float32 operands to a relative error of a few 2^-24 of sum |x_k f_k|, i.e. at the level of an FP32 FMA chain
     (what the FFMA kernels and the reference's float64 arithmetic rounded to float32 deliver).
"""
import numpy as np


def bf16_rn(x):
    """float32 -> nearest bfloat16 (ties to even), returned as float32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return r.astype(np.uint32).view(np.float32)


def split3(x):
    x = np.asarray(x, dtype=np.float32)
    b1 = bf16_rn(x)
    r = (x - b1).astype(np.float32)
    b2 = bf16_rn(r)
    r2 = (r - b2).astype(np.float32)
    b3 = bf16_rn(r2)
    return b1, b2, b3


def test_split_is_exact():
    rng = np.random.RandomState(0)
    x = np.concatenate([
        rng.standard_normal(200000).astype(np.float32) * 3,
        (rng.standard_normal(200000) * 10.0 ** rng.uniform(-20, 20, 200000)).astype(np.float32),
        np.array([0.0, -0.0, 1.0, -1.0, 1 + 2.0 ** -23, 1 - 2.0 ** -24, 255.99998, 1e38, 1e-30], dtype=np.float32),
    ])
    b1, b2, b3 = split3(x)
    total = b1.astype(np.float64) + b2.astype(np.float64) + b3.astype(np.float64)
    assert np.array_equal(total, x.astype(np.float64))
    # every part really is a bfloat16 (low 16 bits clear)
    for b in (b1, b2, b3):
        assert not np.any(b.view(np.uint32) & 0xFFFF)


def test_six_part_products_are_fp32_grade():
    rng = np.random.RandomState(1)
    D = 100
    rho = 0.95
    cov = (1 - rho) * np.eye(D) + rho
    F = np.linalg.inv(cov).astype(np.float32)                     # Case 3c force matrix (cond ~ 1900)
    d = (rng.standard_normal((512, D)) * 1.4).astype(np.float32)  # shifted positions
    a1, a2, a3 = (p.astype(np.float64) for p in split3(d))
    f1, f2, f3 = (p.astype(np.float64) for p in split3(F))
    # accumulation order of the kernel: small terms first; tensor-core accumulation is emulated in float64 and rounded
    # to float32 once per part product (fp32 accumulators hold each partial result)
    acc = np.zeros((512, D), dtype=np.float32)
    for pa, pb in ((a1, f3), (a3, f1), (a2, f2), (a1, f2), (a2, f1), (a1, f1)):
        acc = (acc.astype(np.float64) + pa @ pb.T).astype(np.float32)
    exact = d.astype(np.float64) @ F.astype(np.float64).T
    scale = np.abs(d).astype(np.float64) @ np.abs(F).astype(np.float64).T
    err = np.abs(acc.astype(np.float64) - exact) / scale
    assert err.max() < 4 * 2.0 ** -24
    # dropping the three O(2^-16) terms would not do: the error grows by two orders of magnitude
    acc3 = (a1 @ f2.T + a2 @ f1.T + a1 @ f1.T)
    assert (np.abs(acc3 - exact) / scale).max() > 50 * err.max()
