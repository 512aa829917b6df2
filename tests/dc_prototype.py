"""numpy prototype of the divide-and-conquer symmetric tridiagonal eigensolver.

Development aid + executable specification for bigkrls_b200/csrc/stedc.cu: the same split
between HOST logic (tree, sorting, deflation bookkeeping: O(n) per merge) and DEVICE kernels
(secular equation roots, Loewner re-computation of z, eigenvector matrix, GEMMs).  The C++
host logic is unit-tested against `host_deflate` of this file (tests/test_host_logic.py).

Algorithm: Cuppen's rank-one tearing with the Gu/Eisenstat stable eigenvector formula (the
method behind LAPACK dstedc/dlaed0-4, which arma::eig_sym -> dsyevd uses; reference
src/eigen.cpp:24).  Written from the published algorithm, not from LAPACK source.
"""
import numpy as np

EPS = np.finfo(np.float64).eps / 2  # unit roundoff (LAPACK dlamch('E'))


def build_tree(n, leaf):
    """Returns (leaves [(lo,hi)], merges grouped by height [[(lo,mid,hi)]])."""
    leaves, by_height = [], {}

    def rec(lo, hi):
        if hi - lo <= leaf:
            leaves.append((lo, hi))
            return 0
        mid = (lo + hi) // 2
        h = 1 + max(rec(lo, mid), rec(mid, hi))
        by_height.setdefault(h, []).append((lo, mid, hi))
        return h

    rec(0, n)
    return leaves, [by_height[h] for h in sorted(by_height)]


def host_deflate(d, z, rho, n1):
    """Deflation bookkeeping for one merge (host side).

    d: eigenvalues of the two children (any order), z: [last row of Q1 ; s * first row of Q2]
    (un-normalised), rho: |beta|.  Returns a dict:
      K, dlam[K], w[K]            non-deflated poles (ascending) and weights
      rho                         2*|beta|
      nd_cols[K]                  source column of each non-deflated pole
      nd_type[K]                  1 (top rows only), 2 (dense), 3 (bottom rows only)
      defl_cols[n-K], defl_vals   deflated source columns and their eigenvalues
      rots [(pj, nj, c, s)]       Givens rotations to apply to the columns of Q, in order
    """
    n = d.size
    d = d.copy()
    z = z / np.sqrt(2.0)
    rho = abs(2.0 * rho)
    order = np.argsort(d, kind="stable")
    coltyp = np.where(np.arange(n) < n1, 1, 3)
    tol = 8.0 * EPS * max(np.max(np.abs(d)), np.max(np.abs(z)))
    rots, nd, defl = [], [], []
    if rho * np.max(np.abs(z)) <= tol:
        return dict(K=0, dlam=np.zeros(0), w=np.zeros(0), rho=rho, nd_cols=np.zeros(0, int),
                    nd_type=np.zeros(0, int), defl_cols=order.copy(), defl_vals=d[order], rots=[])
    pj = -1
    for j in range(n):
        nj = order[j]
        if rho * abs(z[nj]) <= tol:
            defl.append(nj)
            continue
        if pj < 0:
            pj = nj
            continue
        s, c = z[pj], z[nj]
        tau = np.hypot(c, s)
        t = d[nj] - d[pj]
        c /= tau
        s = -s / tau
        if abs(t * c * s) <= tol:
            z[nj] = tau
            z[pj] = 0.0
            if coltyp[nj] != coltyp[pj]:
                coltyp[nj] = 2
            rots.append((pj, nj, c, s))
            t = d[pj] * c * c + d[nj] * s * s
            d[nj] = d[pj] * s * s + d[nj] * c * c
            d[pj] = t
            defl.append(pj)
            pj = nj
        else:
            nd.append(pj)
            pj = nj
    nd.append(pj)
    nd = np.array(nd, int)
    defl = np.array(defl, int)
    return dict(K=nd.size, dlam=d[nd], w=z[nd], rho=rho, nd_cols=nd, nd_type=coltyp[nd],
                defl_cols=defl, defl_vals=d[defl], rots=rots)


def secular_root(j, dlam, w2, rho):
    """Root j of 1 + rho * sum w2_i/(dlam_i - x) in (dlam_j, dlam_{j+1}) (last: (dlam_K-1, +rho]).
    Returns (origin index, mu): lambda = dlam[origin] + mu, mu computed to high relative accuracy."""
    K = dlam.size
    if K == 1:
        return 0, rho * w2[0]
    last = (j == K - 1)
    if last:
        org = K - 1
        lo, hi = 0.0, rho * np.sum(w2)       # g(lo+) = -inf, g(hi) >= 0
    else:
        gap = dlam[j + 1] - dlam[j]
        half = 0.5 * gap
        dl = dlam - dlam[j]
        gmid = 1.0 + rho * np.sum(w2 / (dl - half))
        if gmid >= 0.0:
            org, lo, hi = j, 0.0, half
        else:
            org, lo, hi = j + 1, -half, 0.0
    dl = dlam - dlam[org]                     # shifted poles, dl[org] == 0

    def g_parts(mu):
        t = w2 / (dl - mu)
        psi = rho * np.sum(t[: j + 1])                 # poles <= j  (negative terms)
        phi = rho * np.sum(t[j + 1:])                  # poles >  j  (positive terms)
        dpsi = rho * np.sum(t[: j + 1] / (dl[: j + 1] - mu))
        dphi = rho * np.sum(t[j + 1:] / (dl[j + 1:] - mu))
        return psi, phi, dpsi, dphi

    # start: midpoint of the bracket (the pole end of the bracket is excluded by construction)
    mu = 0.5 * (lo + hi)
    for it in range(200):
        psi, phi, dpsi, dphi = g_parts(mu)
        gval = 1.0 + psi + phi
        err = 8.0 * EPS * (1.0 + abs(psi) + abs(phi))
        if abs(gval) <= err:
            break
        if gval < 0.0:
            lo = mu
        else:
            hi = mu
        if (hi - lo) <= 2.0 * EPS * max(abs(lo), abs(hi)):
            mu = 0.5 * (lo + hi)
            break
        # rational interpolation ("middle way"): psi ~ a + s/(dj - x), phi ~ b + S/(dj1 - x)
        if last:
            dj = dl[K - 1] - mu
            s = dpsi * dj * dj
            cst = 1.0 + psi - dpsi * dj + phi
            new = dl[K - 1] + s / cst if cst > 0 else np.inf
        else:
            dj, dj1 = dl[j] - mu, dl[j + 1] - mu
            s = dpsi * dj * dj
            S = dphi * dj1 * dj1
            cst = 1.0 + (psi - dpsi * dj) + (phi - dphi * dj1)
            # cst + s/(Dj - x) + S/(Dj1 - x) = 0  with Dj = dl[j], Dj1 = dl[j+1]
            Dj, Dj1 = dl[j], dl[j + 1]
            # cst (Dj-x)(Dj1-x) + s (Dj1-x) + S (Dj-x) = 0
            qa = cst
            qb = -(cst * (Dj + Dj1) + s + S)
            qc = cst * Dj * Dj1 + s * Dj1 + S * Dj
            if qa == 0.0:
                new = qc / -qb if qb != 0 else np.inf
            else:
                disc = qb * qb - 4 * qa * qc
                if disc < 0:
                    new = np.inf
                else:
                    sq = np.sqrt(disc)
                    # numerically stable pair of roots
                    q = -0.5 * (qb + np.copysign(sq, qb))
                    r1 = q / qa
                    r2 = qc / q if q != 0 else np.inf
                    new = r1 if (lo < r1 < hi) else r2
        if not (lo < new < hi) or not np.isfinite(new):
            new = 0.5 * (lo + hi)
        mu = new
    return org, mu


def merge(Q, D, lo, mid, hi, beta):
    """Merge children [lo,mid) and [mid,hi) in place."""
    n, n1 = hi - lo, mid - lo
    Qb = Q[lo:hi, lo:hi]
    z = np.concatenate([Qb[n1 - 1, :n1], np.sign(beta) * Qb[n1, n1:]])
    plan = host_deflate(D[lo:hi], z, abs(beta), n1)
    for (pj, nj, c, s) in plan["rots"]:                     # device: rotation kernel
        x, y = Qb[:, pj].copy(), Qb[:, nj].copy()
        Qb[:, pj] = c * x + s * y
        Qb[:, nj] = c * y - s * x
    K = plan["K"]
    newQ = np.zeros((n, n))
    newD = np.empty(n)
    if K > 0:
        dlam, w, rho = plan["dlam"], plan["w"], plan["rho"]
        w2 = w * w
        org = np.empty(K, int)
        mu = np.empty(K)
        for j in range(K):                                  # device: one thread per root
            org[j], mu[j] = secular_root(j, dlam, w2, rho)
        lam = dlam[org] + mu
        # delta[i, j] = dlam_i - lambda_j, accurate
        delta = (dlam[:, None] - dlam[org][None, :]) - mu[None, :]
        # Loewner / Gu-Eisenstat re-computed z (device: one thread per i)
        zhat = np.empty(K)
        for i in range(K):
            p = delta[i, i]
            for jj in range(K):
                if jj != i:
                    p *= delta[i, jj] / (dlam[i] - dlam[jj])
            zhat[i] = np.copysign(np.sqrt(-p), w[i])
        U = zhat[:, None] / delta
        U /= np.linalg.norm(U, axis=0)[None, :]
        # grouped GEMMs
        typ = plan["nd_type"]
        g = np.concatenate([np.nonzero(typ == 1)[0], np.nonzero(typ == 2)[0], np.nonzero(typ == 3)[0]])
        c1, c2 = int(np.sum(typ == 1)), int(np.sum(typ == 2))
        cols = plan["nd_cols"][g]
        Ug = U[g, :]
        newQ[:n1, :K] = Qb[:n1, cols[: c1 + c2]] @ Ug[: c1 + c2, :]
        newQ[n1:, :K] = Qb[n1:, cols[c1:]] @ Ug[c1:, :]
        newD[:K] = lam
    newQ[:, K:] = Qb[:, plan["defl_cols"]]
    newD[K:] = plan["defl_vals"]
    Q[lo:hi, lo:hi] = newQ
    D[lo:hi] = newD
    return plan


def stedc(d, e, leaf=32):
    from scipy.linalg import eigh_tridiagonal
    d = np.array(d, dtype=np.float64)
    e = np.array(e, dtype=np.float64)
    n = d.size
    leaves, levels = build_tree(n, leaf)
    for lv in levels:
        for (lo, mid, hi) in lv:
            b = abs(e[mid - 1])
            d[mid - 1] -= b
            d[mid] -= b
    Q = np.zeros((n, n))
    D = np.empty(n)
    for (lo, hi) in leaves:
        if hi - lo == 1:
            D[lo], Q[lo, lo] = d[lo], 1.0
        else:
            D[lo:hi], Q[lo:hi, lo:hi] = eigh_tridiagonal(d[lo:hi], e[lo:hi - 1])
    stats = []
    for lv in levels:
        for (lo, mid, hi) in lv:
            plan = merge(Q, D, lo, mid, hi, e[mid - 1])
            stats.append((hi - lo, plan["K"]))
    order = np.argsort(D, kind="stable")
    return D[order], Q[:, order], stats


if __name__ == "__main__":
    import sys, time
    from scipy.linalg import eigh_tridiagonal, hessenberg
    rng = np.random.default_rng(0)
    for name, n in [("random", 300), ("kernel", 400), ("wilkinson", 201), ("glued", 256)]:
        if name == "random":
            d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
        elif name == "kernel":
            X = rng.standard_normal((n, 3))
            Kmat = np.exp(-((X[:, None, :] - X[None, :, :]) ** 2).sum(2) / 3)
            H = hessenberg(Kmat)
            d, e = np.diag(H).copy(), np.diag(H, -1).copy()
        elif name == "wilkinson":
            m = (n - 1) // 2
            d, e = np.abs(np.arange(-m, m + 1)).astype(float), np.ones(n - 1)
        else:
            d = np.tile(np.arange(1, 17, dtype=float), n // 16)
            e = np.ones(n - 1); e[15::16] = 1e-9
        t0 = time.time()
        lam, Q, stats = stedc(d, e, leaf=16)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        ref = eigh_tridiagonal(d, e, eigvals_only=True)
        print(f"{name:10s} n={n} |lam-ref|/|T|={np.max(np.abs(lam-ref))/np.max(np.abs(ref)):.2e} "
              f"orth={np.max(np.abs(Q.T@Q-np.eye(n))):.2e} resid={np.max(np.abs(T@Q-Q*lam))/np.max(np.abs(ref)):.2e} "
              f"top-merge K={stats[-1]} t={time.time()-t0:.1f}s")
