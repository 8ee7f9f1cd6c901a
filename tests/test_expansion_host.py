"""CPU tests of the Cartesian expansion operators (nbody_b200/csrc/expansion.cuh compiled
for the host with g++): P2M -> M2M -> M2L -> L2L -> L2P against FP64 direct summation
and against the oracle's multipoles, for every supported order."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
f32p = np.ctypeslib.ndpointer(np.float32, flags="C")


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(HERE, "host", "expansion_host.cpp")
    so = os.path.join(HERE, "host", "libexpansion_host.so")
    hdr = os.path.join(HERE, "..", "nbody_b200", "csrc", "expansion.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", src, "-o", so])
    L = C.CDLL(so)
    L.exp_chain.argtypes = [C.c_int, f32p, C.c_int, f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_float, f32p, f32p, f32p]
    L.exp_derivatives.argtypes = [C.c_int] + [C.c_float] * 4 + [f32p]
    return L


def multi_indices(p):
    return [(i, j, o - i - j) for o in range(p + 1) for i in range(o, -1, -1) for j in range(o - i, -1, -1)]


def test_index_order_matches_oracle(hostlib):
    for p in (2, 3, 4, 5):
        mi = multi_indices(p)
        assert hostlib.exp_ncoef(p) == len(mi)
        for a, (i, j, k) in enumerate(mi):
            assert hostlib.exp_index(i, j, k) == a


@pytest.mark.parametrize("p", [2, 3, 4, 5])
def test_derivative_tensor_against_finite_differences(hostlib, p):
    x = np.array([0.31, -0.22, 0.47]); eps2 = 1e-4
    D = np.zeros(hostlib.exp_ncoef(p), np.float32)
    hostlib.exp_derivatives(p, *[float(v) for v in x], eps2, D)
    phi = lambda y: 1.0 / np.sqrt((y ** 2).sum() + eps2)

    def deriv(n, y, h=2e-3):
        n = list(n)
        for d in range(3):
            if n[d] > 0:
                n2 = n.copy(); n2[d] -= 1
                e = np.zeros(3); e[d] = h
                return (deriv(n2, y + e, h) - deriv(n2, y - e, h)) / (2 * h)
        return phi(y)
    for a, n in enumerate(multi_indices(p)):
        ref = deriv(n, x)
        assert abs(D[a] - ref) <= 2e-3 * max(1.0, abs(ref)) * (1 + sum(n)), (n, D[a], ref)


@pytest.mark.parametrize("p,tol", [(2, 6e-2), (3, 1.5e-2), (4, 4e-3), (5, 1.5e-3)])
def test_operator_chain_converges(hostlib, p, tol):
    rng = np.random.default_rng(1)
    s = 0.125
    cB = np.array([0.25, 0.25, 0.25], np.float32); cBc = cB + np.array([s / 4, -s / 4, s / 4], np.float32)
    cA = np.array([0.25 + 4 * s, 0.25 + s, 0.25 - 2 * s], np.float32); cAc = cA + np.array([-s / 4, s / 4, s / 4], np.float32)
    ns, nt = 20, 10
    src = np.zeros((ns, 4), np.float32); src[:, :3] = cBc + (rng.random((ns, 3)) - 0.5) * s / 2; src[:, 3] = rng.random(ns) + 0.5
    tgt = (cAc + (rng.random((nt, 3)) - 0.5) * s / 2).astype(np.float32)
    eps = 0.01
    d = src[None, :, :3].astype(np.float64) - tgt[:, None, :].astype(np.float64)
    r2 = (d ** 2).sum(-1) + eps * eps
    g = (src[None, :, 3:4] * d / r2[..., None] ** 1.5).sum(1)
    ph = (src[None, :, 3] / np.sqrt(r2)).sum(1)
    nc = hostlib.exp_ncoef(p)
    out = np.zeros((nt, 4), np.float32); M = np.zeros(nc, np.float32); Lo = np.zeros(nc, np.float32)
    hostlib.exp_chain(p, src, ns, cBc, cB, cA, cAc, tgt, nt, eps, out, M, Lo)
    err = np.sqrt(((out[:, :3] - g) ** 2).sum() / (g ** 2).sum())
    assert err < tol
    assert np.abs(out[:, 3] - ph).max() / ph.max() < tol
    # multipoles about cB against their definition M_m = sum q (y - c)^m / m!
    from math import factorial as f
    r = src[:, :3].astype(np.float64) - cB.astype(np.float64)
    for a, (i, j, k) in enumerate(multi_indices(p)):
        ref = (src[:, 3] * r[:, 0] ** i * r[:, 1] ** j * r[:, 2] ** k).sum() / (f(i) * f(j) * f(k))
        assert abs(M[a] - ref) <= 1e-5 * max(abs(ref), 1e-3), (i, j, k)
