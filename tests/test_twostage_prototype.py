"""CPU checks of the executable specification of the two-stage tridiagonalisation (tests/twostage_prototype.py):
the numpy algorithm the CUDA kernels sy2sb.cu / sb2st.cu / q2_blocked.cu were written from, and the block
ordering rule of the GEMM-based Q2 back-transformation."""
import numpy as np
import pytest
from scipy.linalg import eigh_tridiagonal

import krls_oracle as o
import twostage_prototype as tp


def kernel_matrix(n, p, seed=5):
    X, y = o.synthetic(n, p, seed)
    Xs, *_ = o.standardize(X, y)
    return o.gauss_kernel(Xs, p)


@pytest.mark.parametrize("n,p,b", [(90, 3, 8), (131, 4, 16), (65, 3, 64)])
def test_prototype_reduction_and_backtransform(n, p, b):
    A = kernel_matrix(n, p)
    B, r1 = tp.sy2sb(A, b)
    assert np.max(np.abs(np.tril(B, -(b + 1)))) == 0.0                       # banded
    d, e, r2, Bt = tp.sb2st(B, b)
    assert np.max(np.abs(np.tril(Bt, -2))) < 1e-13                           # tridiagonal
    assert max(v.size for _, v, _ in r2) <= b                                # reflector length <= bandwidth
    lam, S = eigh_tridiagonal(d, e)
    ref = np.linalg.eigvalsh(A)
    assert np.max(np.abs(lam - ref)) < 1e-12 * ref.max()
    Q = tp.apply_q1(r1, tp.apply_q2(r2, S))
    assert np.max(np.abs(A @ Q - Q * lam)) < 1e-12 * ref.max()
    assert np.max(np.abs(Q.T @ Q - np.eye(n))) < 1e-12


def label_reflectors(r2, n, b):
    """(sweep j, hop t) labels of the stage-2 reflectors, which sb2st returns in generation order."""
    refl, idx = {}, 0
    for j in range(n - 2):
        if min(n, j + 1 + b) - (j + 1) < 2:
            continue
        t = 0
        while idx < len(r2) and r2[idx][0] == j + 1 + t * b:
            refl[(j, t)] = r2[idx]
            idx += 1
            t += 1
    assert idx == len(r2)
    return refl


@pytest.mark.parametrize("n,p,b", [(100, 3, 8), (150, 3, 16)])
def test_blocked_q2_order(n, p, b):
    """q2_blocked.cu applies blocks (J, t) = 64 consecutive sweeps of one hop index in steps
    sigma = 3 (Jmax - J) + t; blocks of one step must act on disjoint rows and the stepped order must give the
    sequential result bit for bit (reflectors on disjoint rows commute exactly)."""
    A = kernel_matrix(n, p)
    B, _ = tp.sy2sb(A, b)
    _, _, r2, _ = tp.sb2st(B, b)
    refl = label_reflectors(r2, n, b)
    Z = np.random.default_rng(0).standard_normal((n, 5))
    ref = tp.apply_q2(r2, Z)
    Jmax = max(j for j, _ in refl) // b
    blocks = {}
    for (j, t) in refl:
        blocks.setdefault((j // b, t), []).append(j)
    steps = {}
    for (J, t) in blocks:
        steps.setdefault(3 * (Jmax - J) + t, []).append((J, t))
    Zb = Z.copy()
    for sigma in sorted(steps):
        rows = [set(range(b * (J + t), min(n, b * (J + t) + 2 * b))) for (J, t) in steps[sigma]]
        assert sum(len(r) for r in rows) == len(set().union(*rows))          # disjoint row ranges in a step
        for (J, t) in reversed(steps[sigma]):                                 # any order inside a step
            for j in sorted(blocks[(J, t)], reverse=True):
                lo, v, tau = refl[(j, t)]
                blk = Zb[lo:lo + v.size]
                blk -= np.outer(v, tau * (v @ blk))
    assert np.array_equal(Zb, ref)
