"""Accuracy of the branch-free exp the streaming kernels use (abn_b200/csrc/fastmath.cuh), host build."""
import ctypes as C
import math

import numpy as np

import helpers as H


def _fast_exp(t):
    lib = H.sim_lib()
    lib.sim_fast_exp.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long]
    t = np.ascontiguousarray(t, dtype=np.float64)
    out = np.empty_like(t)
    lib.sim_fast_exp(t.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)), len(t))
    return out


def test_fast_exp_about_one_ulp_over_the_working_range():
    rng = np.random.default_rng(0)
    t = np.concatenate([rng.uniform(-700, 700, 200_000), rng.uniform(-40, 40, 200_000), rng.normal(0, 2, 200_000),
                        np.array([0.0, -0.0, 1e-300, -1e-300, 700.0, -700.0, 36.7368005696771, -354.891356446692])])
    got = _fast_exp(t)
    ref = np.exp(t.astype(np.longdouble))
    ulp = np.spacing(np.exp(t))
    err = np.abs((got.astype(np.longdouble) - ref) / ulp.astype(np.longdouble)).astype(np.float64)
    assert err.max() < 1.35, err.max()      # table entry rounding + final fma + truncation and 21-bit c4, c5 (0.25 ulp)
    assert np.mean(err) < 0.35


def test_fast_exp256_about_one_ulp_over_the_working_range():
    """fast_exp256 (lean passes): 256-entry pre-scaled table, quartic polynomial."""
    lib = H.sim_lib()
    lib.sim_fast_exp256.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long]
    rng = np.random.default_rng(2)
    t = np.concatenate([rng.uniform(-700, 700, 200_000), rng.uniform(-40, 40, 200_000), rng.normal(0, 2, 200_000),
                        np.array([0.0, -0.0, 1e-300, -1e-300, 700.0, -700.0, 36.7368005696771, -354.891356446692])])
    got = np.empty_like(t)
    lib.sim_fast_exp256(t.ctypes.data_as(C.POINTER(C.c_double)), got.ctypes.data_as(C.POINTER(C.c_double)), len(t))
    ref = np.exp(t.astype(np.longdouble))
    ulp = np.spacing(np.exp(t))
    err = np.abs((got.astype(np.longdouble) - ref) / ulp.astype(np.longdouble)).astype(np.float64)
    assert err.max() < 1.5, err.max()
    assert np.mean(err) < 0.4
    assert got[-8] == 1.0 and got[-7] == 1.0


def test_fast_exp64r_about_one_ulp_over_the_working_range():
    """fast_exp64r (lean passes): 64 entries in 16 bank-interleaved replicas, quintic polynomial."""
    lib = H.sim_lib()
    lib.sim_fast_exp64r.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long]
    rng = np.random.default_rng(3)
    t = np.concatenate([rng.uniform(-700, 700, 200_000), rng.uniform(-40, 40, 200_000), rng.normal(0, 2, 200_000),
                        np.array([0.0, -0.0, 1e-300, -1e-300, 700.0, -700.0, 36.7368005696771, -354.891356446692])])
    got = np.empty_like(t)
    lib.sim_fast_exp64r(t.ctypes.data_as(C.POINTER(C.c_double)), got.ctypes.data_as(C.POINTER(C.c_double)), len(t))
    ref = np.exp(t.astype(np.longdouble))
    ulp = np.spacing(np.exp(t))
    err = np.abs((got.astype(np.longdouble) - ref) / ulp.astype(np.longdouble)).astype(np.float64)
    assert err.max() < 1.5, err.max()
    assert np.mean(err) < 0.4
    assert got[-8] == 1.0 and got[-7] == 1.0


def test_fast_exp256s_is_exp_of_a_uniformly_rescaled_argument():
    """fast_exp256s (lean passes, one reduction step): the error is n * delta with n = round(256 t / ln 2) and
    |delta| <= 2.2e-19, i.e. it computes exp((1 + kappa) t), |kappa| <= 8e-17: <= 6e-14 relative at |t| = 700, a few
    ulp for |t| <= 20, and once the common factor is taken out the usual ~1 ulp."""
    lib = H.sim_lib()
    lib.sim_fast_exp256s.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long]
    rng = np.random.default_rng(4)
    t = np.concatenate([rng.uniform(-700, 700, 200_000), rng.uniform(-20, 20, 200_000), np.array([0.0, -0.0])])
    got = np.empty_like(t)
    lib.sim_fast_exp256s(t.ctypes.data_as(C.POINTER(C.c_double)), got.ctypes.data_as(C.POINTER(C.c_double)), len(t))
    ref = np.exp(t.astype(np.longdouble))
    rel = np.abs((got.astype(np.longdouble) - ref) / ref).astype(np.float64)
    assert rel.max() < 7e-14
    assert rel[np.abs(t) <= 20].max() < 3e-15
    # the structured part: rel error / t is (nearly) one constant kappa
    big = np.abs(t) > 100
    kappa = ((got.astype(np.longdouble) - ref) / ref / t.astype(np.longdouble))[big].astype(np.float64)
    assert np.abs(kappa).max() < 1e-16 and kappa.std() < 2e-18
    assert got[-2] == 1.0 and got[-1] == 1.0


def test_fast_exp_exact_points():
    got = _fast_exp(np.array([0.0, math.log(2.0) * 0]))
    assert got[0] == 1.0


def test_fast_log_for_arguments_at_least_one():
    lib = H.sim_lib()
    lib.sim_fast_log.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long]
    rng = np.random.default_rng(1)
    x = np.concatenate([1.0 + np.exp(rng.uniform(-40, 40, 300_000)), rng.uniform(1, 6, 200_000),
                        np.array([1.0, 1.0 + 2 ** -52, 2.0, 4.0, 1e300, 1.0 + 1e-10])])
    out = np.empty_like(x)
    lib.sim_fast_log(x.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)), len(x))
    ref = np.log(x.astype(np.longdouble))
    abs_err = np.abs(out.astype(np.longdouble) - ref).astype(np.float64)
    rel_err = abs_err / np.maximum(np.abs(ref.astype(np.float64)), 1e-300)
    # sums of log terms need absolute accuracy near x = 1 and ~ulp relative accuracy elsewhere
    assert (abs_err / np.maximum(1.0, np.abs(ref.astype(np.float64)))).max() < 4e-16
    assert rel_err[x > 1.5].max() < 5e-16
    assert out[-6] == 0.0 and abs_err[-1] < 1e-25  # near 1: r = m - 1 exactly, log1p series
