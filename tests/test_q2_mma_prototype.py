"""CPU checks of the executable specification of the DMMA back-transformation (tests/q2_mma_prototype.py): the
block-reflector passes on transposed fragments reproduce the one-reflector-at-a-time product at every window shape
the kernel meets (orders below one window, windows clipped by the end of the matrix, zero-padded last passes, column
counts that do not fill a tile)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import q2_mma_prototype as qp  # noqa: E402


@pytest.mark.parametrize("n,kc", [(3, 3), (5, 2), (10, 5), (66, 4), (67, 9), (70, 8), (131, 16), (200, 7)])
def test_block_reflector_passes_match_sequential_reflectors(n, kc):
    rng = np.random.default_rng(100 * n + kc)
    VV, TAU = qp.random_reflectors(n, rng)
    Z = rng.standard_normal((n, kc))
    ref = qp.apply_sequential(VV, TAU, Z)
    got = qp.apply_mma(VV, TAU, Z)
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref))
    # orthogonality is preserved (the product of reflectors with tau = 2 / v'v)
    assert abs(np.linalg.norm(got) - np.linalg.norm(Z)) <= 1e-12 * np.linalg.norm(Z)


def test_triangular_factor_of_a_pass_is_the_compact_wy_factor():
    n, t = 150, 0
    rng = np.random.default_rng(5)
    VV, TAU = qp.random_reflectors(n, rng)
    jp = n - 3 - 8  # second pass of hop index 0
    buf = qp.stage_pass(VV, TAU, n, t, jp)
    T = np.zeros((8, 8))
    for ee in range(2):
        for l in range(32):
            T[2 * (l % 4) + ee][l // 4] = buf[1152 + 32 * ee + l]
    lo = jp + 1
    V = np.zeros((72, 8))
    for s in range(8):
        for w in range(72):
            vi, row = w - (7 - s), lo - 7 + w
            if 0 <= vi < 64 and row < n:
                V[w, s] = VV[row, jp - s]
    H = np.eye(72)
    for s in range(8):  # H = H_7 ... H_1 H_0
        H = (np.eye(72) - TAU[jp - s, t] * np.outer(V[:, s], V[:, s])) @ H
    assert np.allclose(H, np.eye(72) - V @ T.T @ V.T, atol=1e-13)
    assert np.allclose(np.tril(T, -1), 0.0)
