"""numpy prototype of the restarted block-Krylov top-k symmetric eigensolver (executable specification of
bigkrls_b200/csrc/eigen_topk.cu; replaces arma::eigs_sym of reference src/eigen.cpp:18-22)."""
import numpy as np


def orth_gram(W, drop=1e-14):
    """Orthonormal basis of span(W) through the eigen-decomposition of the Gram matrix (twice).
    Robust to rank deficiency: directions with tiny Gram eigenvalues are dropped."""
    for _ in range(2):
        G = W.T @ W
        lam, S = np.linalg.eigh(G)
        keep = lam > drop * lam.max()
        W = W @ (S[:, keep] / np.sqrt(lam[keep]))
    return W


def topk(K, k, b=32, extra=8, tol=1e-13, seed=0, max_outer=200, verbose=False):
    n = K.shape[0]
    rng = np.random.default_rng(seed)
    m_max = min(n, k + extra * b)
    X = orth_gram(rng.standard_normal((n, b)))
    V = np.zeros((n, 0))
    KV = np.zeros((n, 0))
    nmv = 0
    for outer in range(max_outer):
        while V.shape[1] + X.shape[1] <= m_max and X.shape[1] > 0:
            W = K @ X
            nmv += X.shape[1]
            V = np.hstack([V, X])
            KV = np.hstack([KV, W])
            for _ in range(2):
                W = W - V @ (V.T @ W)
            X = orth_gram(W)
            if V.shape[1] + X.shape[1] > m_max:
                break
        H = V.T @ KV
        H = (H + H.T) / 2
        th, S = np.linalg.eigh(H)
        idx = np.argsort(th)[::-1][:k]
        th, S = th[idx], S[:, idx]
        Q, KQ = V @ S, KV @ S
        R = KQ - Q * th
        res = np.linalg.norm(R, axis=0)
        if verbose:
            print(outer, V.shape[1], "max res", res.max() / th[0], "nconv", int(np.sum(res <= tol * th[0])), "matvecs", nmv)
        if res.max() <= tol * th[0]:
            return th, Q, dict(outer=outer + 1, matvecs=nmv)
        # thick restart: keep the Ritz pairs, continue with the worst residual directions
        worst = np.argsort(res)[::-1][:b]
        Rb = R[:, worst]
        for _ in range(2):
            Rb = Rb - Q @ (Q.T @ Rb)
        X = orth_gram(Rb)
        V, KV = Q, KQ
    raise RuntimeError("no convergence")


if __name__ == "__main__":
    import sys, time
    sys.path.insert(0, "oracle")
    import krls_oracle as o
    for (N, P, k) in [(2000, 5, 50), (3000, 10, 150), (4000, 20, 200)]:
        X, y = o.synthetic(N, P, 1004)
        Xs, *_ = o.standardize(X, y)
        K = o.gauss_kernel(Xs, P)
        ref = np.linalg.eigvalsh(K)[::-1]
        t = time.time()
        th, Q, st = topk(K, k, b=32, extra=8, verbose=False)
        print(N, P, k, "stats", st, "val err", np.max(np.abs(th - ref[:k])) / ref[0], "orth", np.max(np.abs(Q.T @ Q - np.eye(k))),
              "resid", np.max(np.abs(K @ Q - Q * th)) / ref[0], "lam_k/lam_1", ref[k - 1] / ref[0], round(time.time() - t, 1), "s")
