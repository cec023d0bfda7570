"""hhd_exp / Philox / step-rounding helpers (csrc/hhd_math.h), exercised through the oracle's plain exports."""
import ctypes as C
import math

import numpy as np

from oracle_binding import oracle_lib


def _lib():
    lib = oracle_lib()
    lib.hhdo_exp.restype = C.c_double
    lib.hhdo_exp.argtypes = [C.c_double]
    lib.hhdo_philox_u01.restype = C.c_double
    lib.hhdo_philox_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
    lib.hhdo_delay_steps.restype = C.c_int64
    lib.hhdo_delay_steps.argtypes = [C.c_double, C.c_double]
    lib.hhdo_timestep.restype = C.c_int64
    lib.hhdo_timestep.argtypes = [C.c_double, C.c_double]
    return lib


def test_exp_within_one_ulp_of_libm():
    lib = _lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-745, 709, 20000), rng.uniform(-40, 40, 20000), rng.uniform(-1e-3, 1e-3, 2000),
                         [0.0, 1.0, -1.0, 709.78, -708.4, -744.0, 1e-300, -1e-300, 0.34657359, -0.34657359]])
    worst = 0.0
    for x in xs:
        got, ref = lib.hhdo_exp(float(x)), math.exp(float(x))
        if ref == 0.0 or math.isinf(ref):
            assert got == ref
            continue
        ulp = math.ulp(ref)
        worst = max(worst, abs(got - ref) / ulp)
    assert worst <= 1.0, worst


def test_exp_special_values():
    lib = _lib()
    assert lib.hhdo_exp(0.0) == 1.0
    assert lib.hhdo_exp(710.0) == math.inf
    assert lib.hhdo_exp(float("inf")) == math.inf
    assert lib.hhdo_exp(-746.0) == 0.0
    assert lib.hhdo_exp(float("-inf")) == 0.0
    assert math.isnan(lib.hhdo_exp(float("nan")))
    assert 0.0 < lib.hhdo_exp(-740.0) < 1e-320          # subnormal result, rounded once


def _expected_u01(w0, w1):
    return float((((w0 >> 5) << 26) | (w1 >> 6))) / 9007199254740992.0


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors: philox4x32-10 of counter/key all 0 and all ff."""
    lib = _lib()
    assert lib.hhdo_philox_u01(0, 0, 0, 0) == _expected_u01(0x6627e8d5, 0xe169c58d)
    allf = 0xFFFFFFFFFFFFFFFF
    assert lib.hhdo_philox_u01(allf, allf, 0x00FFFFFFFFFFFFFF, 0xFF) == _expected_u01(0x408f276d, 0x41c83b0e)


def test_philox_uniformity_and_independence():
    lib = _lib()
    u = np.array([lib.hhdo_philox_u01(42, i, 7, 0) for i in range(20000)])
    assert 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs((u < 0.3).mean() - 0.3) < 0.01
    v = np.array([lib.hhdo_philox_u01(42, i, 8, 0) for i in range(20000)])
    assert abs(np.corrcoef(u, v)[0, 1]) < 0.03
    assert lib.hhdo_philox_u01(42, 5, 7, 0) != lib.hhdo_philox_u01(42, 5, 7, 1)


def test_step_rounding_rules():
    lib = _lib()
    # Brian's C++ SpikeQueue: int(delay/dt + 0.5)
    assert lib.hhdo_delay_steps(0.0, 0.05) == 0
    assert lib.hhdo_delay_steps(0.0249, 0.05) == 0
    assert lib.hhdo_delay_steps(0.0251, 0.05) == 1
    assert lib.hhdo_delay_steps(1.57, 0.05) == 31
    assert lib.hhdo_delay_steps(3.6, 0.01) == 360
    # Brian's timestep(): int((t + 1e-3*dt)/dt): refractory 5 ms -> 100 steps at 0.05, 500 at 0.01
    assert lib.hhdo_timestep(5.0, 0.05) == 100
    assert lib.hhdo_timestep(5.0, 0.01) == 500
    assert lib.hhdo_timestep(0.3, 0.1) == 3          # 0.3/0.1 = 2.9999999999999996 without the epsilon


def test_cbrt_within_two_ulp():
    """hhd_cbrt (growth factor of the adaptive integrator's step-size controller): same code on host and device."""
    lib = _lib()
    lib.hhdo_cbrt.restype = C.c_double
    lib.hhdo_cbrt.argtypes = [C.c_double]
    rng = np.random.default_rng(1)
    xs = np.concatenate([10.0 ** rng.uniform(-9, 9, 5000), rng.uniform(9.765625e-4, 0.5, 5000), [1.0, 8.0, 27.0, 0.125, 0.5]])
    for x in xs:
        got, ref = lib.hhdo_cbrt(float(x)), float(np.cbrt(np.longdouble(x)))
        assert abs(got - ref) <= 2 * math.ulp(ref), x
    assert lib.hhdo_cbrt(27.0) == 3.0 and lib.hhdo_cbrt(0.125) == 0.5


def test_log_and_normals():
    """hhd_log (<= 2 ulp away from 1) and the polar-method Gaussian pair built on it (conductance-noise variant)."""
    lib = _lib()
    lib.hhdo_log.restype = C.c_double
    lib.hhdo_log.argtypes = [C.c_double]
    lib.hhdo_normal2.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_double)]
    rng = np.random.default_rng(2)
    for x in np.concatenate([rng.uniform(1e-12, 0.9, 4000), rng.uniform(1.1, 1e6, 2000), [0.5, 0.25, 2.0]]):
        got, ref = lib.hhdo_log(float(x)), float(np.log(np.longdouble(x)))
        assert abs(got - ref) <= 2 * math.ulp(ref), x
    assert lib.hhdo_log(1.0) == 0.0
    z = (C.c_double * 2)()
    zs = []
    for i in range(20000):
        lib.hhdo_normal2(7, i, 123, 2, z)
        zs += [z[0], z[1]]
    zs = np.asarray(zs)
    assert abs(zs.mean()) < 0.02 and abs(zs.std() - 1) < 0.02
    assert abs(((zs - zs.mean()) ** 4).mean() / zs.var() ** 2 - 3) < 0.1
    assert abs(np.corrcoef(zs[0::2], zs[1::2])[0, 1]) < 0.03
    lib.hhdo_normal2(7, 5, 123, 2, z)
    a = (z[0], z[1])
    lib.hhdo_normal2(7, 5, 123, 2, z)
    assert a == (z[0], z[1])                                  # a pure function of (seed, index, step, stream)
