"""CPU: host-side logic of the int8 digit-plane variance path (cfg.variance_slices) and the validity of the numpy
digit restatement the GPU tests compare the tcgen05 kernel with (tests/test_gpu_sliced.py)."""
import ctypes

import numpy as np
import pytest

from test_gpu_sliced import _digit_restatement, _split_rows


def test_digit_planes_are_exact_and_bounded():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(40, 300)) * np.exp(4 * rng.normal(size=(40, 1)))
    a[3] = 0.0
    planes, e = _split_rows(a, 6)
    for q in planes:
        assert np.array_equal(q, np.rint(q)) and np.max(np.abs(q)) <= 64          # signed 7-bit digits: int8
    rec = sum(q * 2.0 ** (-7 * p) for p, q in enumerate(planes)) * np.ldexp(1.0, e - 6)[:, None]
    amax = np.max(np.abs(a), axis=1, keepdims=True)
    assert np.all(np.abs(rec - a) <= amax * 2.0 ** (-7 * 6 + 1) + 1e-300)        # 6 planes: 2^-41 of the row maximum
    assert e[3] == 0 and not np.any([q[3].any() for q in planes])                  # all-zero row


@pytest.mark.parametrize("s", [5, 6, 7, 8, 9])
def test_digit_restatement_converges_like_two_to_the_minus_seven_s(s):
    rng = np.random.default_rng(s)
    a = rng.normal(size=(30, 500))
    b = np.tril(rng.normal(size=(500, 500)))
    ref = a @ b.T
    scale = np.max(np.abs(a), axis=1)[:, None] * np.max(np.abs(b), axis=1)[None, :] * 500
    err = np.max(np.abs(_digit_restatement(a, b, s, True) - ref) / scale)
    assert err < 2.0 ** (-7 * s + 8)
    assert s >= 8 or err > 2.0 ** (-7 * s - 12)           # and not by accident better than the planes allow


def test_int32_plane_sums_cannot_overflow_within_the_documented_bound():
    # (g + 1) pairs of |digit| <= 64 over K columns: (g + 1) * 4096 * K < 2^31  <=>  K < 2^19 / s for the widest group
    for s in range(5, 10):
        k_max = (1 << 19) // s
        assert s * 4096 * (k_max - 1) < 2 ** 31


def test_runtime_knob_reaches_the_handle_config(monkeypatch):
    from nngp_b200 import _lib, runtime, stax
    seen = {}

    class Probe:
        def __init__(self, **kw):
            seen.update(kw)

    monkeypatch.setattr(_lib, "Handle", Probe)
    monkeypatch.setattr(runtime, "_variance_slices", 0)
    _, _, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))
    runtime.new_handle(kernel_fn.spec, diag_reg=1e-3)
    assert seen["variance_slices"] == 0
    runtime.set_variance_slices(7)
    runtime.new_handle(kernel_fn.spec, diag_reg=1e-3)
    assert seen["variance_slices"] == 7
    runtime.new_handle(kernel_fn.spec, diag_reg=1e-3, kernel_type="ntk")           # 'ntk' never takes the planes
    assert seen["variance_slices"] == 0
    runtime.set_variance_slices(0)


def test_config_struct_carries_variance_slices():
    from nngp_b200 import _lib
    cfg = _lib.NngpConfig()
    cfg.variance_slices = 7
    assert ctypes.sizeof(cfg) == 368 and _lib.NngpConfig.variance_slices.offset == 360
    assert "nngp_sliced_product" in _lib.EXPORTS
