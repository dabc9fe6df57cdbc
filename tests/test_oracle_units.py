"""not gpu: unit checks of the oracle's geometry / filters / id parsing / polynomial roots, modelled on the
reference's test/geometry_test.cpp (1e-4 round trips) and test/avg_filter_test.cpp."""
import numpy as np
import pytest

from tests import orc


def _rand_quats(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def test_rot_quat_round_trip():   # geometry_test.cpp:25-53
    L = orc.lib()
    rng = np.random.default_rng(1)
    for q in _rand_quats(rng, 100):
        R = np.zeros(9); q2 = np.zeros(4)
        L.orc_quat_to_rot(orc.ptr(q), orc.ptr(R))
        L.orc_rot_to_quat(orc.ptr(R), orc.ptr(q2))
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-4
        Rm = R.reshape(3, 3)
        assert np.allclose(Rm @ Rm.T, np.eye(3), atol=1e-12)


def test_quat_rpy_round_trip():   # geometry_test.cpp:55-66,131-151
    L = orc.lib()
    rng = np.random.default_rng(2)
    for _ in range(100):
        rpy = rng.uniform([-3.1, -1.5, -3.1], [3.1, 1.5, 3.1])
        q = np.zeros(4); back = np.zeros(3)
        L.orc_rpy_to_quat(orc.ptr(rpy), orc.ptr(q))
        L.orc_quat_to_rpy(orc.ptr(q), orc.ptr(back))
        assert np.abs(back - rpy).max() < 1e-4
        R = np.zeros(9); r2 = np.zeros(3)
        L.orc_quat_to_rot(orc.ptr(q), orc.ptr(R))
        L.orc_rot_to_rpy(orc.ptr(R), orc.ptr(r2))          # geometry_test.cpp:173 (quatToRpy vs rot->rpy)
        assert np.abs(r2 - rpy).max() < 1e-4


def test_angle_helpers():
    L = orc.lib()
    assert np.isclose(L.orc_constrain_angle(3 * np.pi / 2), -np.pi / 2)
    assert np.isclose(L.orc_angle_diff(0.1, 0.3), 0.2)
    assert np.isclose(L.orc_angle_diff(3.0, -3.0), 2 * np.pi - 6.0)
    assert np.isclose(L.orc_wrap_min_max(4.0, -np.pi, np.pi), 4.0 - 2 * np.pi)
    prev = np.array([3.1, 0.0, -3.1]); new = np.array([-3.1, 0.1, 3.1]); out = np.zeros(3)
    L.orc_unwrap3(orc.ptr(prev), orc.ptr(new), orc.ptr(out))
    assert np.allclose(out, [3.1 + (2 * np.pi - 6.2), 0.1, -3.1 - (2 * np.pi - 6.2)])


def test_to_sec_is_uncontracted():
    L = orc.lib()
    for sec, nsec in [(1000, 4000000), (1697600000, 999999999), (5, 1)]:
        assert L.orc_to_sec(sec, nsec) == float(sec) + 1e-9 * float(nsec)


def test_inverse_vs_numpy():
    L = orc.lib()
    rng = np.random.default_rng(3)
    for n in (3, 6):
        A = rng.normal(size=(n, n)); S = A @ A.T + np.eye(n)
        out = np.zeros((n, n))
        L.orc_inverse(orc.ptr(np.ascontiguousarray(S)), n, orc.ptr(out))
        assert np.allclose(out, np.linalg.inv(S), rtol=1e-11, atol=1e-13)


def test_avg_filters():   # avg_filter_test.cpp:14-44
    L = orc.lib()
    rng = np.random.default_rng(4)
    v = rng.normal(5.0, 1.0, 10000)
    a = L.orc_avg_new(1000); m = L.orc_mavg_new(1000)
    for x in v:
        ra = L.orc_avg_update(a, x); rm = L.orc_mavg_update(m, x)
    assert abs(ra - 5.0) < 0.1 and abs(rm - 5.0) < 0.1
    assert abs(L.orc_mavg_variance(m) - 1.0) < 0.1
    assert np.isclose(rm, v[-1000:].mean(), rtol=1e-10)
    L.orc_avg_delete(a); L.orc_mavg_delete(m)
    # partially filled window: mean over the filled part, variance over ALL slots (utils.hpp:222-251)
    m = L.orc_mavg_new(4)
    assert L.orc_mavg_update(m, 2.0) == 2.0
    assert L.orc_mavg_update(m, 4.0) == 3.0
    assert np.isclose(L.orc_mavg_variance(m), ((2 - 3) ** 2 + (4 - 3) ** 2 + 9 + 9) / 2)
    L.orc_mavg_delete(m)


def test_get_id():   # utils.hpp:302-313, SURVEY.md H10
    import ctypes as C
    L = orc.lib()
    out = C.c_uint()
    assert L.orc_get_id(b"target_12", C.byref(out)) == 1 and out.value == 12
    assert L.orc_get_id(b"target_filt_12", C.byref(out)) == 0
    assert L.orc_get_id(b"target", C.byref(out)) == 0
    assert L.orc_get_id(b"target_x", C.byref(out)) == -1      # std::stoi throws


def test_poly_roots_vs_numpy():
    L = orc.lib()
    rng = np.random.default_rng(5)
    for _ in range(50):
        c = rng.normal(size=5)
        re = np.zeros(4); im = np.zeros(4)
        n = L.orc_poly_roots(orc.ptr(c), 5, orc.ptr(re), orc.ptr(im))
        assert n == 4
        got = np.sort_complex(re + 1j * im)
        ref = np.sort_complex(np.roots(c[::-1]))
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-10)
    # lowestRealRoot rules (src/intersection_solver.cpp:4-17): leading coefficient 0 -> -1; complex only -> -1
    assert L.orc_lowest_real_root(orc.ptr(np.array([1.0, 2.0, 1.0, 0.0, 0.0])), 5) == -1
    assert L.orc_lowest_real_root(orc.ptr(np.array([1.0, 0.0, 0.0, 0.0, 1.0])), 5) == -1
    r = L.orc_lowest_real_root(orc.ptr(np.array([-6.0, 1.0, 7.0, -1.0, -1.0]) * -1), 5)   # (t-1)(t+1)(t-2)(t+3)
    assert np.isclose(r, -3.0)
