"""GPU test of the FP64-pipe approximations behind every spectra kernel (csrc/common.cuh fast_exp / fast_rcp,
csrc/spectra_feqmod.cuh fast_sqrt) against numpy/libm, through the C ABI's is3d_probe_math.

The continuous paths promise 1e-10 relative per bin; these primitives are held to 4e-15 for exp at x <= 60 (1024-entry
table + degree-3 polynomial, truncation 5.5e-16, plus |x| 1.1e-16 from the one-fma argument reduction), 4e-14 up to the
clamp at x = 680, and 4.5e-16 (2 ulp) for rcp and sqrt."""
import numpy as np
import pytest

import cases
import harness

pytestmark = pytest.mark.gpu


def test_fast_math_against_libm(libs, tmp_path):
    name = "s3d_m1"
    surf, _ = harness.load_golden(name)
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(-20.0, 60.0, 400_000), np.array([0.0, 1e-300, -1e-300, 1e-9, -1e-9])])
    xl = np.concatenate([rng.uniform(60.0, 679.9, 50_000), rng.uniform(-700.0, -20.0, 50_000), np.array([679.99])])
    pos = np.concatenate([np.exp(rng.uniform(-40.0, 700.0, 300_000)), rng.uniform(0.1, 10.0, 200_000), np.array([1.0, 2.0, 1e300, 0.1])])
    with harness.open_session(str(tmp_path), cases.SPECTRA_CASES[name], surf) as h:
        e, _, _ = h.abi_probe_math(x)
        el, _, _ = h.abi_probe_math(xl)
        _, r, s = h.abi_probe_math(pos)
        far, _, _ = h.abi_probe_math(np.array([680.5, 708.5, 1e4, 1e8, 1e300, np.inf, np.nan]))
    for got, arg, tol in ((e, x, 4e-15), (el, xl, 4e-14)):
        ref = np.exp(arg.astype(np.longdouble))
        err = np.abs((got - ref) / ref).astype(np.float64)
        assert err.max() < tol, (err.max(), arg[np.argmax(err)])
    assert np.abs(r * pos - 1.0).max() < 4.5e-16
    assert np.abs(s / np.sqrt(pos) - 1.0).max() < 4.5e-16
    # x is clamped at 680 (and NaN behaves like it): "huge" and finite, so that 1/(e^x + s) is the reference's 0 to more
    # than 200 decades and e^x times an energy ratio still fits a double
    assert np.all(far >= 2e295) and np.all(far <= 3e295)


def test_aniso_angular_primitives_against_libm(libs, tmp_path):
    """df_mode 5 (PTMA) term sums: atan(s)/s for z = s^2 > 0 and atanh(s)/s for z = -s^2 in (-1, 0), and ln (csrc/aniso.cuh
    fast_atan / fast_atanh_over_s / fast_log) against long-double libm.  The closed forms of the angular functions divide by
    z up to twice, so t = atan(s)/s resp. atanh(s)/s must be good to a few ulp RELATIVE down to |z| = 0.01."""
    name = "s3d_m1"
    surf, _ = harness.load_golden(name)
    rng = np.random.default_rng(11)
    z = np.concatenate([rng.uniform(0.01, 0.999999, 300_000), np.exp(rng.uniform(np.log(0.01), 0.0, 100_000)) * 0.999999,
                        np.array([0.01, 0.0100001, 0.25, 0.5, 0.9, 0.99, 0.999999])])
    big = np.concatenate([np.exp(rng.uniform(-30.0, 30.0, 200_000)), rng.uniform(0.5, 2.0, 100_000), np.array([1.0, 1.5, 0.75, 2.0, 1e-300, 1e300])])
    with harness.open_session(str(tmp_path), cases.SPECTRA_CASES[name], surf) as h:
        a, b, _ = h.abi_probe_aniso_math(z)
        a2, _, lg = h.abi_probe_aniso_math(big)
    zl = z.astype(np.longdouble)
    s = np.sqrt(zl)
    ref_atan = np.arctan(s) / s
    ref_atanh = np.arctanh(s) / s
    e_atan = np.abs((a - ref_atan) / ref_atan).astype(np.float64)
    e_atanh = np.abs((b - ref_atanh) / ref_atanh).astype(np.float64)
    # atanh is ill-conditioned towards s -> 1: the (correctly rounded or not) sqrt's last bit is amplified by
    # kappa = s / ((1 - s^2) atanh s); the reference's own sqrt + atanh carries the same factor
    kappa = (s / ((1 - zl) * np.arctanh(s))).astype(np.float64)
    print(f"atan(s)/s max rel err {e_atan.max():.2e}; atanh(s)/s max rel err / (1 + kappa) {(e_atanh / (1 + kappa)).max():.2e}")
    assert e_atan.max() < 1e-15, (e_atan.max(), z[np.argmax(e_atan)])
    assert np.all(e_atanh <= 6e-16 * (1.0 + kappa)), (float((e_atanh / (1 + kappa)).max()), z[np.argmax(e_atanh / (1 + kappa))])
    sb = np.sqrt(big.astype(np.longdouble))
    e2 = np.abs((a2 - np.arctan(sb) / sb) / (np.arctan(sb) / sb)).astype(np.float64)
    assert e2.max() < 1e-15, (e2.max(), big[np.argmax(e2)])
    ref_log = np.log(big.astype(np.longdouble))
    err = np.abs(lg - ref_log).astype(np.float64)
    # relative 4e-16 away from ln 1 = 0; near v = 1 the absolute error is bounded by the rounding of v - 1 itself
    bound = 8e-16 * np.maximum(np.abs(ref_log.astype(np.float64)), 1e-3)   # k ln2 + ln m cancels by up to 2.4 at v -> 0.75
    print(f"ln max err / bound {(err / bound).max():.2e}")
    assert np.all(err <= bound), (float((err / bound).max()), big[np.argmax(err / bound)])
