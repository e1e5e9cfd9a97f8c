"""CPU checks of arithmetic identities the device code relies on (no GPU, no library calls)."""
import numpy as np


def _py_mod2(a):
    """common.cuh py_mod2: python's `a % 2` as a - 2 trunc(a / 2) + the sign fix-up of float_rem (no fmod)."""
    with np.errstate(invalid="ignore"):  # inf - inf -> nan, like fmod(inf, 2)
        m = a - a.dtype.type(2) * np.trunc(a * a.dtype.type(0.5))
        m = np.where(m < 0, m + a.dtype.type(2), m)
    return np.where(m == 0, a.dtype.type(0), m)


def _py_mod_fmod(a):
    """common.cuh py_mod(a, 2): fmod + the same fix-up (what CPython's float % does)."""
    with np.errstate(invalid="ignore"):
        m = np.fmod(a, a.dtype.type(2))
        m = np.where(m < 0, m + a.dtype.type(2), m)
    return np.where(m == 0, a.dtype.type(0), m)


def test_mod2_without_fmod_is_bit_exact():
    """inverted_double_pendulum.py:59 `(theta + pi) % 2`: every operation of a - 2 trunc(a / 2) is exact for finite a,
    so it equals fmod(a, 2) bit for bit (i2p.cuh uses it instead of fmod's remainder loop)."""
    rng = np.random.default_rng(0)
    for dt, ut, n in ((np.float32, np.uint32, 4_000_000), (np.float64, np.uint64, 4_000_000)):
        bits = rng.integers(0, np.iinfo(ut).max, size=n, dtype=ut, endpoint=True)
        a = bits.view(dt)
        edge = np.array([0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 3.0, -3.0, 1e-30, -1e-30, 2.0 ** 23, 2.0 ** 24 + 2, -(2.0 ** 52) - 1, 2.0 ** 53,
                         np.finfo(dt).max, -np.finfo(dt).max, np.finfo(dt).tiny, -np.finfo(dt).tiny, np.inf, -np.inf, np.nan, np.pi, -np.pi,
                         6.283185307179586, 1.9999999, -1.9999999], dtype=dt)
        a = np.concatenate([a, edge, (rng.uniform(-50, 50, size=n // 4)).astype(dt)])
        x, y = _py_mod2(a), _py_mod_fmod(a)
        both_nan = np.isnan(x) & np.isnan(y)
        assert np.array_equal(x.view(ut)[~both_nan], y.view(ut)[~both_nan])
        fin = np.isfinite(a)
        assert ((x[fin] >= 0) & (x[fin] <= 2)).all()  # == 2 only by the rounding of m + 2 for tiny negative a, as in python
        py = np.array([float(v) % 2.0 for v in a[fin][:20000].astype(np.float64)]) if dt is np.float64 else None
        if py is not None:  # CPython's own float % on the same values
            assert np.array_equal(x[fin][:20000].view(ut), py.view(ut))
