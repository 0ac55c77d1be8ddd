"""numpy restatement of the fixed-point representations used by the tensor-core backend
(graphicalmodellearning.jl_b200/csrc/eval_tc.cu).  TEST INFRASTRUCTURE ONLY, like the rest of oracle/: the product
never imports it.  It states, in plain integer / float32 arithmetic, what the kernels compute, so that the algebraic
identities the CUDA path relies on (exact limb recombination, bias removal through column sums, digits as byte extracts
of the rounding word) are checked on the CPU; the GPU parity tests then compare the kernels with the float64 oracle.

The objective being evaluated is the reference's (src/GraphicalModelLearning.jl:169-172): nothing here changes it, these
are only number representations."""
from __future__ import annotations

import numpy as np

X_LATTICE_FINE, X_LATTICE_COARSE, X_LATTICE_ROUGH = 2.0 ** -24, 2.0 ** -22, 2.0 ** -14


def r_bias(nr: int) -> int:
    return {1: 0x80, 2: 0x8000, 3: 0x400000, 4: 0x80000000}[nr]


def r_magic(nr: int) -> np.float32:
    return np.float32(12582912.0 + {1: 128.0, 2: 32768.0}.get(nr, 0.0))


def balanced_limbs(q: np.ndarray, xl: int) -> np.ndarray:
    """q (int64, |q| within the range of xl digits) -> [xl, ...] int8 limbs, most significant first:
    digits 1..xl-1 are balanced base-256 digits in [-128, 127] (the full int8 range), digit 0 takes what is left
    (tc_quantize_x_kernel)."""
    q = q.astype(np.int64).copy()
    d = np.zeros((xl,) + q.shape, dtype=np.int64)
    for j in range(xl - 1, 0, -1):
        dj = ((q + 128) & 255) - 128
        q = (q - dj) >> 8
        d[j] = dj
    d[0] = q
    assert np.all(d[0] <= 127) and np.all(d[0] >= -128), "iterate outside the lattice range"
    return d.astype(np.int8)


def recombine_limb_sums(acc: np.ndarray) -> np.ndarray:
    """[xl, ...] int32 limb sums (most significant first) -> exact integer energy in lattice units (epilogue).
    The kernel joins them in 32-bit wrap-around arithmetic (join2): partial products may leave the int32 range, the
    energy itself does not -- restated here with explicit wrapping."""
    def join2(hi, lo):
        return ((hi.astype(np.int64) * 256 + lo.astype(np.int64) + 2 ** 31) % 2 ** 32) - 2 ** 31
    xl = acc.shape[0]
    if xl == 2:
        return join2(acc[0], acc[1])
    if xl == 3:
        return join2(join2(acc[0], acc[1]), acc[2])
    # 4 limbs: the two 16-bit halves are joined separately and combined in floating point (fmaf(hi, 65536, lo)); the
    # integer they represent:
    return join2(acc[0], acc[1]) * 65536 + join2(acc[2], acc[3])


def fma_f32(a: np.ndarray, b: np.ndarray, c) -> np.ndarray:
    """fp32 fused multiply-add, round to nearest even, emulated exactly: a 24x24-bit product plus a 25-bit addend of
    comparable magnitude fits float64, so one rounding float64 -> float32 is the fma's rounding."""
    return (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(np.float32)


def residual_digits(g: np.ndarray, sscale: np.ndarray, nr: int) -> np.ndarray:
    """Epilogue quantisation for nr <= 3: word = float_as_int(fma(g, sscale, magic)); the nr low bytes of the word are
    the stored digits (most significant first).  g = w psi >= 0, sscale = +-1/deltaR, |g*sscale| < 2^22."""
    word = fma_f32(g, sscale, r_magic(nr)).view(np.uint32)
    return np.stack([(word >> (8 * (nr - 1 - j))) & 0xFF for j in range(nr)]).astype(np.uint8)


def digits_to_q(digits: np.ndarray) -> np.ndarray:
    """stored digits -> signed residual integer q (what the gradient contraction effectively sums)."""
    nr = digits.shape[0]
    v = np.zeros(digits.shape[1:], dtype=np.int64)
    for j in range(nr):
        v = v * 256 + digits[j].astype(np.int64)
    return v - r_bias(nr)


def gradient_from_digits(digits: np.ndarray, S: np.ndarray) -> np.ndarray:
    """tc_grad_kernel + tc_grad_init_kernel for one node: digits [nr, K] (u8), S [K, F] (+-1).
    Per-plane int32 accumulation, planes recombined base 256, accumulators started at -BIAS * colsum."""
    nr = digits.shape[0]
    colsum = S.astype(np.int64).sum(axis=0)
    acc = np.stack([digits[j].astype(np.int64) @ S.astype(np.int64) for j in range(nr)])
    assert np.all(np.abs(acc) < 2 ** 31), "int32 accumulator range"
    v = np.zeros(S.shape[1], dtype=np.int64)
    for j in range(nr):
        v = v * 256 + acc[j]
    return v - r_bias(nr) * colsum
