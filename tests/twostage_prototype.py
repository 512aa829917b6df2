"""numpy prototype of the two-stage tridiagonalisation (executable specification of
bigkrls_b200/csrc/sy2sb.cu / sb2st.cu):

  stage 1  dense -> band (bandwidth b): panel Householder QR + two-sided block update
           A22 <- Q' A22 Q,  Q = I - V T V',  Z = A22 V T,  W = Z - 1/2 V (T' V' Z),  A22 -= V W' + W V'
  stage 2  band -> tridiagonal by bulge chasing (one column of the bulge per hop; reflectors of length <= b)
  back     eigenvectors of A = Q1 Q2 (eigenvectors of T)

Written from the published algorithms (Bischof/Lang/Sun successive band reduction).
"""
import numpy as np


def house(x):
    """Householder: (I - tau v v') x = beta e1, v[0] = 1."""
    alpha = x[0]
    xn2 = float(np.dot(x[1:], x[1:]))
    if xn2 == 0.0:
        return np.concatenate([[1.0], np.zeros(x.size - 1)]), 0.0, alpha
    beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
    tau = (beta - alpha) / beta
    v = x / (alpha - beta)
    v[0] = 1.0
    return v, tau, beta


def panel_qr(P):
    """Householder QR of m x bb panel -> V (m x nr unit lower trapezoidal), taus, R (in place, upper)."""
    P = P.copy()
    m, bb = P.shape
    nr = min(bb, m - 1)
    V = np.zeros((m, nr))
    taus = np.zeros(nr)
    for j in range(nr):
        v, tau, beta = house(P[j:, j].copy())
        V[j:, j] = v
        taus[j] = tau
        P[j, j] = beta
        P[j + 1:, j] = 0.0
        if j + 1 < bb:
            w = tau * (v @ P[j:, j + 1:])
            P[j:, j + 1:] -= np.outer(v, w)
    return V, taus, P


def larft(V, taus):
    nr = V.shape[1]
    T = np.zeros((nr, nr))
    S = V.T @ V
    for i in range(nr):
        T[i, i] = taus[i]
        if i > 0:
            T[:i, i] = -taus[i] * (T[:i, :i] @ S[:i, i])
    return T


def sy2sb(A, b):
    """Returns band matrix B (dense storage, bandwidth b) and the list of (r0, V, T) block reflectors."""
    A = A.copy()
    n = A.shape[0]
    refl = []
    for c0 in range(0, n, b):
        r0 = c0 + b
        m = n - r0
        if m < 2:
            break
        bb = min(b, n - c0)
        V, taus, R = panel_qr(A[r0:, c0:c0 + bb])
        T = larft(V, taus)
        A[r0:, c0:c0 + bb] = R
        A[c0:c0 + bb, r0:] = R.T
        A22 = A[r0:, r0:]
        Z = A22 @ (V @ T)
        W = Z - 0.5 * V @ (T.T @ (V.T @ Z))
        A22 -= V @ W.T + W @ V.T
        refl.append((r0, V, T))
    return A, refl


def sb2st(B, b):
    """Bulge chasing on the dense-stored band matrix.  Returns d, e and the reflectors
    [(row_lo, v, tau)] in the order they were generated."""
    B = B.copy()
    n = B.shape[0]
    refl = []
    for j in range(n - 2):
        lo, hi = j + 1, min(n, j + 1 + b)
        if hi - lo < 2:
            continue
        v, tau, beta = house(B[lo:hi, j].copy())
        B[lo:hi, j] = 0.0
        B[lo, j] = beta
        B[j, lo:hi] = B[lo:hi, j]
        refl.append((lo, v, tau))
        # two-sided update of the diagonal block
        D = B[lo:hi, lo:hi]
        w = tau * (D @ v)
        w -= 0.5 * tau * (w @ v) * v
        D -= np.outer(v, w) + np.outer(w, v)
        while True:
            nlo, nhi = hi, min(n, hi + b)
            if nlo >= n:
                break
            Bk = B[nlo:nhi, lo:hi]
            # right-apply the previous reflector: Bk <- Bk (I - tau v v')
            Bk -= np.outer(tau * (Bk @ v), v)
            B[lo:hi, nlo:nhi] = Bk.T
            if nhi - nlo < 2:
                break
            v2, tau2, beta2 = house(Bk[:, 0].copy())
            refl.append((nlo, v2, tau2))
            # left-apply the new reflector to the rest of the block
            Bk[:, 0] = 0.0
            Bk[0, 0] = beta2
            if Bk.shape[1] > 1:
                Bk[:, 1:] -= np.outer(v2, tau2 * (v2 @ Bk[:, 1:]))
            B[lo:hi, nlo:nhi] = Bk.T
            D = B[nlo:nhi, nlo:nhi]
            w = tau2 * (D @ v2)
            w -= 0.5 * tau2 * (w @ v2) * v2
            D -= np.outer(v2, w) + np.outer(w, v2)
            lo, hi, v, tau = nlo, nhi, v2, tau2
    return np.diag(B).copy(), np.diag(B, -1).copy(), refl, B


def apply_q2(refl, Z):
    """Z <- Q2 Z: reflectors in reverse order of generation."""
    Z = Z.copy()
    for (lo, v, tau) in reversed(refl):
        blk = Z[lo:lo + v.size]
        blk -= np.outer(v, tau * (v @ blk))
    return Z


def apply_q1(refl, Z):
    Z = Z.copy()
    for (r0, V, T) in reversed(refl):
        blk = Z[r0:]
        blk -= V @ (T @ (V.T @ blk))
    return Z


if __name__ == "__main__":
    import sys
    from scipy.linalg import eigh_tridiagonal
    sys.path.insert(0, "oracle")
    import krls_oracle as o
    for (n, p, b) in [(300, 4, 16), (517, 6, 32), (130, 3, 64), (65, 2, 64)]:
        X, y = o.synthetic(n, p, 5)
        Xs, *_ = o.standardize(X, y)
        A = o.gauss_kernel(Xs, p)
        B, r1 = sy2sb(A, b)
        offband = np.max(np.abs(np.tril(B, -(b + 1))))
        d, e, r2, Bt = sb2st(B, b)
        offtri = np.max(np.abs(np.tril(Bt, -2)))
        lam, S = eigh_tridiagonal(d, e)
        ref = np.linalg.eigvalsh(A)
        Q = apply_q1(r1, apply_q2(r2, S))
        print(f"n={n} b={b}: band off {offband:.1e} tri off {offtri:.1e} eig err {np.max(np.abs(lam-ref))/ref.max():.1e} "
              f"resid {np.max(np.abs(A@Q-Q*lam))/ref.max():.1e} orth {np.max(np.abs(Q.T@Q-np.eye(n))):.1e} "
              f"maxlen {max(v.size for _,v,_ in r2)} nrefl2 {len(r2)} band-bulge width {max((np.nonzero(np.abs(Bt[:,0])>-1)[0]).max(),0)}")

