"""numpy prototype of the band -> tridiagonal back-transformation on the FP64 tensor path (executable specification of
q2_mma_kernel / q2_tfac8_kernel in bigkrls_b200/csrc/sb2st.cu): the lane-level fragment layout of mma.sync.m8n8k4,
the permuted k index that lets the window tiles serve as A and as C fragments, the staged pass buffers (dots / update /
triangular factor, in fragment order), the lattice of passes shared by all hop indices, the zero-padded last pass and
the rows a hop index may load and store.  Hop indices are run one after the other (a valid schedule of the kernel's
flag protocol: hop index t waits for hop index t-1 pass by pass)."""
import numpy as np

CB = 64          # bandwidth = reflector length
PB = 18 * 32 + 9 * 64 + 64


def dmma884(c0, c1, a, b):
    """D(8x8) += A(8x4) B(4x8) with the register layout of the instruction: lane l holds A[l/4][l%4], B[l%4][l/4],
    C[l/4][2(l%4) + {0,1}] (common.cuh: dmma884)."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    C = np.zeros((8, 8))
    for l in range(32):
        A[l // 4][l % 4] = a[l]
        B[l % 4][l // 4] = b[l]
        C[l // 4][2 * (l % 4)] = c0[l]
        C[l // 4][2 * (l % 4) + 1] = c1[l]
    C = C + A @ B
    for l in range(32):
        c0[l] = C[l // 4][2 * (l % 4)]
        c1[l] = C[l // 4][2 * (l % 4) + 1]


def random_reflectors(n, rng):
    """VV[row, sweep], TAU[sweep, hop] as the chasing kernel leaves them: reflector (j, t) on rows j+1+64t .."""
    nhop = (n - 3) // CB + 1
    VV = np.zeros((n, n))
    TAU = np.zeros((n, nhop + 1))
    for j in range(n - 2):
        for t in range(nhop):
            if j > n - 3 - CB * t:
                continue
            lo = j + 1 + CB * t
            L = min(CB, n - lo)
            v = rng.standard_normal(L)
            v[0] = 1.0
            VV[lo:lo + L, j] = v
            TAU[j, t] = 2.0 / (v @ v) if L > 1 else 0.0
    return VV, TAU


def apply_sequential(VV, TAU, Z):
    """One reflector after the other, hop index by hop index, sweeps descending (the order q2_apply_kernel realises)."""
    n = Z.shape[0]
    Zr = Z.copy()
    for t in range((n - 3) // CB + 1):
        for j in range(n - 3 - CB * t, -1, -1):
            lo = j + 1 + CB * t
            L = min(CB, n - lo)
            v = VV[lo:lo + L, j]
            Zr[lo:lo + L] -= TAU[j, t] * np.outer(v, v @ Zr[lo:lo + L])
    return Zr


def stage_pass(VV, TAU, n, t, jp):
    """The 1216 doubles of a pass in the order the kernel reads them (q2_mma_kernel::stage + q2_tfac8_kernel)."""
    lo = jp + 1 + CB * t
    buf = np.zeros(PB)

    def entry(s, w):
        vi, row, sw = w - (7 - s), lo - 7 + w, jp - s
        return VV[row, sw] if (sw >= 0 and 0 <= vi < CB and row < n) else 0.0

    for e in range(1152):
        if e < 576:       # dots: k-step q = 2j+ee, lane l: V(w = 8j + 2(l%4) + ee, s = l/4)
            q, l = e // 32, e % 32
            s, w = l // 4, 8 * (q // 2) + 2 * (l % 4) + (q % 2)
        else:             # update: row tile j, k-step ee, lane l: V(w = 8j + l/4, s = 2(l%4) + ee)
            e2 = e - 576
            l = e2 % 32
            s, w = 2 * (l % 4) + ((e2 // 32) % 2), 8 * (e2 // 64) + l // 4
        buf[e] = entry(s, w)
    Vp = np.array([[entry(s, w) for w in range(72)] for s in range(8)])
    tau = np.array([TAU[jp - s, t] if jp - s >= 0 else 0.0 for s in range(8)])
    G = Vp @ Vp.T
    T = np.zeros((8, 8))
    for s in range(8):
        T[s, s] = tau[s]
        if s > 0:
            T[:s, s] = -tau[s] * (T[:s, :s] @ G[:s, s])
    for ee in range(2):
        for l in range(32):
            buf[1152 + 32 * ee + l] = T[2 * (l % 4) + ee][l // 4]
    return buf


def apply_mma(VV, TAU, Z):
    n, kc = Z.shape
    Zt = Z.copy()
    ntile = (kc + 7) // 8
    for t in range((n - 3) // CB + 1):
        jmax = n - 3 - CB * t
        npass = jmax // 8 + 1
        rmin = 1 + CB * t
        lo = jmax + 1 + CB * t
        z = np.zeros((ntile, 18, 32))
        for c in range(ntile):
            for l in range(32):
                col = 8 * c + l // 4
                for j in range(9):
                    for e in range(2):
                        row = lo - 7 + 8 * j + 2 * (l % 4) + e
                        if col < kc and rmin <= row < n:
                            z[c, 2 * j + e, l] = Zt[row, col]
        for i in range(npass):
            jp = jmax - 8 * i
            lo = jp + 1 + CB * t
            last = i == npass - 1
            buf = stage_pass(VV, TAU, n, t, jp)
            nr = np.zeros((ntile, 2, 32))
            if not last:
                for c in range(ntile):
                    for l in range(32):
                        col = 8 * c + l // 4
                        for e in range(2):
                            row = lo - 15 + 2 * (l % 4) + e
                            if col < kc and row >= rmin:
                                nr[c, e, l] = Zt[row, col]
            for c in range(ntile):
                d = np.zeros((2, 2, 32))
                for q in range(18):
                    dmma884(d[q & 1, 0], d[q & 1, 1], z[c, q], buf[32 * q:32 * q + 32])
                w0, w1 = np.zeros(32), np.zeros(32)
                dmma884(w0, w1, d[0, 0] + d[1, 0], buf[1152:1184])
                dmma884(w0, w1, d[0, 1] + d[1, 1], buf[1184:1216])
                for j in range(9):
                    dmma884(z[c, 2 * j], z[c, 2 * j + 1], -w0, buf[576 + 64 * j:608 + 64 * j])
                    dmma884(z[c, 2 * j], z[c, 2 * j + 1], -w1, buf[608 + 64 * j:640 + 64 * j])
                for l in range(32):
                    col = 8 * c + l // 4
                    if col >= kc:
                        continue
                    for e in range(2):
                        row = lo + 57 + 2 * (l % 4) + e
                        if row < n:
                            Zt[row, col] = z[c, 16 + e, l]
                    if last:
                        for j in range(8):
                            for e in range(2):
                                row = lo - 7 + 8 * j + 2 * (l % 4) + e
                                if rmin <= row < n:
                                    Zt[row, col] = z[c, 2 * j + e, l]
                if not last:
                    for j in range(8, 0, -1):
                        z[c, 2 * j] = z[c, 2 * j - 2].copy()
                        z[c, 2 * j + 1] = z[c, 2 * j - 1].copy()
                    z[c, 0], z[c, 1] = nr[c, 0], nr[c, 1]
    return Zt
