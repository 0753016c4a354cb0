"""The two erfc forms the GEMM epilogues evaluate (passl_b200/csrc/gemm.cuh: gelu_phi8 = Abramowitz & Stegun 7.1.26 for the exact-erf
GELU of `vision_transformer.py:84-113` / nn.GELU, gate_gelu8 = A&S 7.1.25 for its derivative), restated in float32 numpy with the
SAME constants and checked against erf in double: the approximation error must vanish under the bf16 rounding of the outputs
(2^-9 relative).  Keeps the constants honest without a GPU; the kernels themselves are checked in tests/test_gemm_gpu.py."""
import math

import numpy as np
from scipy.special import erf


def _phi_7_1_26(x):
    x = x.astype(np.float32)
    z = np.abs(x) * np.float32(0.70710678118654752)
    t = np.float32(1.0) / (np.float32(0.3275911) * z + np.float32(1.0))
    y = np.float32(1.061405429) * t + np.float32(-1.453152027)
    y = y * t + np.float32(1.421413741)
    y = y * t + np.float32(-0.284496736)
    y = y * t + np.float32(0.254829592)
    e = np.exp2(x * x * np.float32(-0.72134752044448170)).astype(np.float32)
    h = np.float32(0.5) * (y * t) * e
    return np.where(x < 0, h, np.float32(1.0) - h), e


def _gate_7_1_25(a):
    a = a.astype(np.float32)
    u = np.abs(a) * np.float32(0.8493218002880191)
    t = np.float32(1.0) / (np.float32(0.39169196791136207) * u + np.float32(1.0))
    e = np.exp2(-(u * u)).astype(np.float32)
    y = np.float32(0.5 * 0.7478556) * t + np.float32(0.5 * -0.0958798)
    y = y * t + np.float32(0.5 * 0.3480242)
    w = np.float32(-0.46971863934982566) * u + y * t
    s = w * e
    return np.where(a < 0, s, np.float32(1.0) - s)


def test_gelu_forward_form_is_exact_to_float32():
    x = np.linspace(-9, 9, 400001)
    phi, _ = _phi_7_1_26(x)
    ref = 0.5 * (1.0 + erf(x / math.sqrt(2.0)))
    assert np.abs(phi - ref).max() < 4e-7
    gelu, refg = x.astype(np.float32) * phi, x * ref
    assert np.abs(gelu - refg).max() < 2e-6
    big = np.abs(refg) > 1e-3                                    # relative error where bf16 has something to round
    assert (np.abs(gelu - refg)[big] / np.abs(refg)[big]).max() < 2.0 ** -11


def test_gelu_gate_form_error_is_far_below_bf16():
    a = np.linspace(-9, 9, 400001)
    g = _gate_7_1_25(a)
    ref = 0.5 * (1.0 + erf(a / math.sqrt(2.0))) + a * np.exp(-a * a / 2.0) / math.sqrt(2.0 * math.pi)
    assert np.abs(g - ref).max() < 1.5e-5                        # A&S 7.1.25: |erfc error| < 2.5e-5, halved by Phi = erfc / 2
    assert abs(float(_gate_7_1_25(np.array([0.0]))[0]) - 0.5) < 1e-5
    # gelu'(a) + gelu'(-a) = 1 is built into the form
    assert np.abs(_gate_7_1_25(a) + _gate_7_1_25(-a) - 1.0).max() < 1e-6


def test_constants_are_the_ones_they_claim_to_be():
    k = math.sqrt(0.5 * math.log2(math.e))
    assert abs(k - 0.8493218002880191) < 1e-15
    assert abs(0.47047 * math.sqrt(0.5) / k - 0.39169196791136207) < 1e-15
    assert abs(1.0 / math.sqrt(2.0 * math.pi) / k - 0.46971863934982566) < 1e-15
    assert abs(0.5 * math.log2(math.e) - 0.72134752044448170) < 1e-15
