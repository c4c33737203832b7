"""The LAPACK stand-in (streamparanumal_b200/shim/minilapack.cpp) that lets the reference's libs/core matrix wrappers
link without a system LAPACK (libs/core/matrixEig.cpp:29-34, matrixInverse.cpp:29-35, matrixRightSolve.cpp:31-45):
dgeev_ on the symmetric Jacobi matrices of the Gauss quadrature, dgetrf_/dgetri_/dgesv_ and the s* variants against
numpy on the matrix sizes of the reference element (N <= 15 -> at most 16 x 16)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = ctypes.POINTER


@pytest.fixture(scope="module")
def lapack(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("minilapack") / "libminilapack.so")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                        os.path.join(ROOT, "streamparanumal_b200", "shim", "minilapack.cpp")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return ctypes.CDLL(so)


def _i(v):
    return ctypes.byref(ctypes.c_int(v))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _jacobi_matrix(n):
    """symmetric tridiagonal Jacobi matrix of the Gauss-Legendre rule (libs/mesh/meshBasis1D.cpp:195-231)"""
    J = np.zeros((n, n))
    for i in range(1, n):
        J[i, i - 1] = J[i - 1, i] = i / np.sqrt(4.0 * i * i - 1.0)
    return J


@pytest.mark.parametrize("n", [1, 2, 3, 8, 16])
@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_geev_symmetric(lapack, n, real):
    fn = lapack.dgeev_ if real == np.float64 else lapack.sgeev_
    A = np.asfortranarray(_jacobi_matrix(n).astype(real))
    wr, wi = np.zeros(n, real), np.ones(n, real)
    vr = np.asfortranarray(np.zeros((n, n), real))
    work = np.zeros(max(4 * n, 1), real)
    info = ctypes.c_int(99)
    fn(ctypes.c_char_p(b"N"), ctypes.c_char_p(b"V"), _i(n), _ptr(A.copy(order="F")), _i(n), _ptr(wr), _ptr(wi), None, _i(1), _ptr(vr), _i(n),
       _ptr(work), _i(work.size), ctypes.byref(info))
    assert info.value == 0
    tol = 1e-13 if real == np.float64 else 2e-6
    assert np.allclose(np.sort(wr), np.linalg.eigvalsh(A.astype(np.float64)), atol=tol)   # the Gauss nodes
    assert np.all(wi == 0)
    # eigenvectors: A v = w v, unit length
    Ad, V, w = A.astype(np.float64), vr.astype(np.float64), wr.astype(np.float64)
    assert np.allclose(Ad @ V, V * w, atol=10 * tol)
    assert np.allclose(np.linalg.norm(V, axis=0), 1.0, atol=10 * tol)


def test_geev_refuses_a_non_symmetric_matrix(lapack):
    n = 3
    A = np.asfortranarray(np.arange(9, dtype=np.float64).reshape(3, 3))
    wr, wi, vr, work = np.zeros(n), np.zeros(n), np.asfortranarray(np.zeros((n, n))), np.zeros(12)
    info = ctypes.c_int(0)
    lapack.dgeev_(ctypes.c_char_p(b"N"), ctypes.c_char_p(b"V"), _i(n), _ptr(A), _i(n), _ptr(wr), _ptr(wi), None, _i(1), _ptr(vr), _i(n),
                  _ptr(work), _i(12), ctypes.byref(info))
    assert info.value == -4


@pytest.mark.parametrize("n", [1, 2, 5, 16])
@pytest.mark.parametrize("real", [np.float64, np.float32])
def test_lu_inverse_and_solve(lapack, n, real):
    rng = np.random.default_rng(100 + n)
    A = np.asfortranarray((rng.uniform(-1, 1, (n, n)) + 0.1 * np.eye(n)).astype(real))   # general: pivoting matters
    tol = 1e-10 if real == np.float64 else 5e-3
    getrf, getri, gesv = (lapack.dgetrf_, lapack.dgetri_, lapack.dgesv_) if real == np.float64 else (lapack.sgetrf_, lapack.sgetri_, lapack.sgesv_)
    # inverse = getrf + getri (libs/core/matrixInverse.cpp:47-75)
    LU = A.copy(order="F")
    ipiv = np.zeros(n, np.int32)
    info = ctypes.c_int(99)
    getrf(_i(n), _i(n), _ptr(LU), _i(n), _ptr(ipiv), ctypes.byref(info))
    assert info.value == 0 and ipiv.min() >= 1 and ipiv.max() <= n   # Fortran (1-based) pivots
    work = np.zeros(max(n * n, 1), real)
    getri(_i(n), _ptr(LU), _i(n), _ptr(ipiv), _ptr(work), _i(work.size), ctypes.byref(info))
    assert info.value == 0
    assert np.allclose(LU.astype(np.float64) @ A.astype(np.float64), np.eye(n), atol=tol * max(1.0, np.linalg.cond(A.astype(np.float64))))
    # solve with several right-hand sides (libs/core/matrixRightSolve.cpp:47-110)
    nrhs = 3
    B = np.asfortranarray(rng.uniform(-1, 1, (n, nrhs)).astype(real))
    X, A2 = B.copy(order="F"), A.copy(order="F")
    gesv(_i(n), _i(nrhs), _ptr(A2), _i(n), _ptr(ipiv), _ptr(X), _i(n), ctypes.byref(info))
    assert info.value == 0
    ref = np.linalg.solve(A.astype(np.float64), B.astype(np.float64))
    assert np.allclose(X.astype(np.float64), ref, atol=tol * max(1.0, np.linalg.cond(A.astype(np.float64))))


def test_singular_matrix_is_reported(lapack):
    n = 3
    A = np.asfortranarray(np.ones((n, n)))
    ipiv = np.zeros(n, np.int32)
    info = ctypes.c_int(0)
    lapack.dgetrf_(_i(n), _i(n), _ptr(A), _i(n), _ptr(ipiv), ctypes.byref(info))
    assert info.value > 0   # U(i,i) == 0, as LAPACK reports it
