"""CPU checks of the integer / float32 identities the tensor-core backend relies on (oracle/fixed_point.py restates
the kernels' number representations in numpy).  They hold exactly -- no tolerances:
  * the balanced base-256 limbs of a lattice point recombine to the point, and the limb-wise contraction equals the
    integer contraction (energy GEMM),
  * the bytes of float_as_int(fma(g, +-scale, magic)) are the digits of round(g * +-scale) + BIAS (epilogue),
  * sum_k digits(k) S[k,f] - BIAS colsum[f] = sum_k q_k S[k,f] (gradient GEMM with unsigned digits)."""
import numpy as np
import pytest

import fixed_point as fx


@pytest.mark.parametrize("xl,lattice,xmax", [(4, fx.X_LATTICE_FINE, 7.9), (3, fx.X_LATTICE_COARSE, 1.95), (2, fx.X_LATTICE_ROUGH, 1.95)])
def test_limbs_recombine_and_contract_exactly(xl, lattice, xmax):
    rng = np.random.default_rng(xl)
    F, K = 257, 96
    x = rng.uniform(-xmax, xmax, size=F) * (rng.random(F) < 0.5)
    x[:4] = [xmax, -xmax, lattice, -lattice]
    q = np.rint(x / lattice).astype(np.int64)
    limbs = fx.balanced_limbs(q, xl)
    assert limbs.dtype == np.int8
    assert np.array_equal(fx.recombine_limb_sums(limbs.astype(np.int32)), q)
    S = rng.choice(np.array([-1, 1], dtype=np.int8), size=(K, F))
    acc = np.stack([S.astype(np.int32) @ limbs[j].astype(np.int32) for j in range(xl)])      # what the MMAs accumulate
    assert np.all(np.abs(acc) <= 128 * F)                  # int32 accumulators: far inside their range
    assert np.array_equal(fx.recombine_limb_sums(acc), S.astype(np.int64) @ q)


@pytest.mark.parametrize("nr,qmax", [(1, 120), (2, 32000), (3, 4000000)])
def test_residual_digits_are_bytes_of_the_rounding_word(nr, qmax):
    rng = np.random.default_rng(nr)
    n = 200_000
    scale = np.float32(qmax / 3.7)                        # 1/deltaR of some node
    g = (rng.random(n) * 3.7).astype(np.float32)          # w psi in [0, top)
    g[:3] = [0.0, np.float32(3.7 * 0.999999), np.float32(1e-12)]
    sgn = rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=n)
    sscale = sgn * scale
    digits = fx.residual_digits(g, sscale, nr)
    # reference: round-to-nearest-even of the fp32 product chain done as ONE fma rounding to the integer grid
    want = np.rint(g.astype(np.float64) * sscale.astype(np.float64)).astype(np.int64)
    got = fx.digits_to_q(digits)
    exact_half = np.abs(np.abs(g.astype(np.float64) * sscale.astype(np.float64)) % 1.0 - 0.5) < 1e-9
    assert np.array_equal(got[~exact_half], want[~exact_half])
    assert np.all(np.abs(got) <= qmax + 1) and digits.dtype == np.uint8


@pytest.mark.parametrize("nr,qmax", [(1, 120), (2, 32000), (3, 4000000)])
def test_bias_is_removed_exactly_by_the_column_sums(nr, qmax):
    rng = np.random.default_rng(10 + nr)
    K, F = 4096, 33
    S = rng.choice(np.array([-1, 1], dtype=np.int8), size=(K, F))
    S[:, -1] = 1                                           # the constant feature (field)
    g = (rng.random(K) * 2.0).astype(np.float32)
    sscale = (rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=K) * np.float32(qmax / 2.0))
    digits = fx.residual_digits(g, sscale, nr)
    q = fx.digits_to_q(digits)
    assert np.array_equal(fx.gradient_from_digits(digits, S), q @ S.astype(np.int64))
    # padded samples (weight 0 -> q = 0) store exactly the bias and therefore cancel
    pad = fx.residual_digits(np.zeros(4, np.float32), np.full(4, qmax, np.float32), nr)
    assert np.all(fx.digits_to_q(pad) == 0)
