"""Eigen-free host algebra (slam_b200/csrc/small_math.hpp) against numpy / scipy.

The reference does this algebra with Eigen (a system package that is not vendored): K.inverse(),
resultRt.inverse(), ldlt().solve(), lu().inverse(), rodrigues, computeUpdateSE3.  No golden vectors exist for it
in the reference, so it is pinned against independent numpy implementations of the same textbook operations.
"""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def hm():
    from slam_b200 import build
    return C.CDLL(str(build.build_hostmath()))


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def random_spd(rng, n, cond=1e4):
    q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    d = np.logspace(0, np.log10(cond), n)
    return (q * d) @ q.T


def test_inverses(hm):
    rng = np.random.default_rng(0)
    for _ in range(50):
        m = rng.normal(size=(3, 3))
        out = np.zeros((3, 3))
        hm.hm_mat3_inverse_d(P(m), P(out))
        assert np.allclose(out, np.linalg.inv(m), rtol=1e-10, atol=1e-10)
        mf = m.astype(np.float32)
        of = np.zeros((3, 3), np.float32)
        hm.hm_mat3_inverse_f(P(mf), P(of))
        assert np.allclose(of, np.linalg.inv(mf.astype(np.float64)), rtol=2e-4, atol=2e-4 * np.abs(np.linalg.inv(mf.astype(np.float64))).max())
        m4 = rng.normal(size=(4, 4))
        o4 = np.zeros((4, 4))
        hm.hm_mat4_inverse_d(P(m4), P(o4))
        assert np.allclose(o4, np.linalg.inv(m4), rtol=1e-9, atol=1e-9)
        aff = np.eye(4)
        aff[:3, :4] = rng.normal(size=(3, 4))
        oa = np.zeros((4, 4))
        hm.hm_mat4_affine_inverse_d(P(aff), P(oa))
        assert np.allclose(oa, np.linalg.inv(aff), rtol=1e-9, atol=1e-9)


def test_ldlt_solves_match_numpy(hm):
    rng = np.random.default_rng(1)
    for _ in range(100):
        A = random_spd(rng, 6, cond=10 ** rng.uniform(0, 8))
        b = rng.normal(size=6)
        x = np.zeros(6)
        xp = np.zeros(6)
        hm.hm_ldlt6_d(P(A), P(b), P(x))
        hm.hm_ldlt6_pivoted_d(P(A), P(b), P(xp))
        ref = np.linalg.solve(A, b)
        assert np.allclose(x, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
        assert np.allclose(xp, ref, rtol=1e-7, atol=1e-9 * np.abs(ref).max())
        A3 = random_spd(rng, 3, cond=100).astype(np.float32)
        b3 = rng.normal(size=3).astype(np.float32)
        x3 = np.zeros(3, np.float32)
        hm.hm_ldlt3_f(P(A3), P(b3), P(x3))
        assert np.allclose(x3, np.linalg.solve(A3.astype(np.float64), b3.astype(np.float64)), rtol=1e-3, atol=1e-4)


def test_ldlt_degenerate_systems(hm):
    """Eigen's ldlt().solve() returns the pseudo-inverse solution for singular PSD systems; A = 0 (no correspondences) must give 0."""
    A = np.zeros((6, 6))
    b = np.arange(1.0, 7.0)
    x = np.ones(6)
    hm.hm_ldlt6_d(P(A), P(b), P(x))
    assert np.all(x == 0)
    # rank-deficient: only the first three unknowns are observable
    rng = np.random.default_rng(2)
    A = np.zeros((6, 6))
    A[:3, :3] = random_spd(rng, 3, 10)
    b = np.zeros(6)
    b[:3] = rng.normal(size=3)
    hm.hm_ldlt6_d(P(A), P(b), P(x))
    assert np.allclose(x[:3], np.linalg.solve(A[:3, :3], b[:3]), rtol=1e-9)
    assert np.all(x[3:] == 0)


def test_rodrigues_and_se3_update(hm):
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(3)
    for _ in range(50):
        r = rng.normal(size=3) * 10 ** rng.uniform(-6, 0)
        R = np.zeros((3, 3))
        hm.hm_rodrigues(P(r), P(R))
        assert np.allclose(R, Rotation.from_rotvec(r).as_matrix(), atol=1e-12)
    R = np.zeros((3, 3))
    hm.hm_rodrigues(P(np.zeros(3)), P(R))
    assert np.array_equal(R, np.eye(3))          # theta < DBL_EPSILON => identity (odom/utils.h:28)

    T = np.eye(4)
    for _ in range(5):
        x = rng.normal(size=6) * 0.01
        expected = np.eye(4)
        expected[:3, :3] = Rotation.from_rotvec(x[3:]).as_matrix()
        expected[:3, 3] = x[:3]
        expected = expected @ T                    # left-multiply, odom/utils.h:66
        Tc = np.ascontiguousarray(T)
        hm.hm_update_se3(P(Tc), P(x))
        assert np.allclose(Tc, expected, atol=1e-12)
        T = Tc
    # currentT = [Rprev|tprev] * rgbOdom^-1   (RGBDOdometryef.cpp:563-575)
    Rprev = Rotation.from_rotvec(rng.normal(size=3)).as_matrix().astype(np.float32)
    tprev = rng.normal(size=3).astype(np.float32)
    Rc = np.zeros((3, 3), np.float32)
    tc = np.zeros(3, np.float32)
    hm.hm_compose_current_pose(P(Rprev), P(tprev), P(T), P(Rc), P(tc))
    Tp = np.eye(4)
    Tp[:3, :3], Tp[:3, 3] = Rprev, tprev
    ref = Tp @ np.linalg.inv(T)
    assert np.allclose(Rc, ref[:3, :3], atol=1e-6) and np.allclose(tc, ref[:3, 3], atol=1e-6)


def test_covariance_inverse(hm):
    rng = np.random.default_rng(4)
    A = random_spd(rng, 6, 1e5)
    out = np.zeros((6, 6))
    assert hm.hm_lu_inverse6(P(A), P(out)) == 1
    assert np.allclose(out @ A, np.eye(6), atol=1e-8)
