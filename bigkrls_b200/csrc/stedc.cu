// Divide-and-conquer eigensolver for the symmetric tridiagonal matrix produced by sytrd.cu.
// Together with sytrd.cu / ormtr.cu this replaces LAPACK dsyevd behind arma::eig_sym
// (reference src/eigen.cpp:24).  Cuppen's rank-one tearing + Gu/Eisenstat stable eigenvectors,
// written from the published algorithm; executable specification: tests/dc_prototype.py.
//
// Split of work:
//   HOST   (O(n) per level): the merge tree, sorting of the children's eigenvalues, the
//          deflation decisions (tiny z_i / close poles -> Givens rotation) and the index
//          bookkeeping.  One D2H of (D, z) and one H2D of the plans per tree level.
//   DEVICE (everything O(n^2) and above), batched over all merges of a level:
//          leaf QL iterations (one warp per <=32 leaf), Givens rotations on Q, column gather,
//          secular-equation roots (one thread per root, origin shifted to the nearer pole,
//          safeguarded rational interpolation), Loewner recomputation of z, the eigenvector
//          matrix U, and the merge GEMMs  Q_new = [Q1 0; 0 Q2]_nondeflated * U  on the DMMA GEMM
//          (two per merge: top rows x (type 1+2 columns), bottom rows x (type 2+3 columns)).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"

namespace bk {

static const double DC_EPS = 1.1102230246251565e-16;  // unit roundoff
static constexpr int LEAF = 32;

// =============================================================================================
// host logic
// =============================================================================================
void host_deflate(const double* d_in, const double* z_in, int n, int n1, double beta,
                  MergePlan* plan) {
  std::vector<double> d(d_in, d_in + n), z(n);
  const double sqrt2 = std::sqrt(2.0);
  for (int i = 0; i < n; ++i) z[i] = z_in[i] / sqrt2;
  const double rho = std::fabs(2.0 * beta);
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return d[a] < d[b]; });
  std::vector<int> coltyp(n);
  for (int i = 0; i < n; ++i) coltyp[i] = (i < n1) ? 1 : 3;
  double dmax = 0, zmax = 0;
  for (int i = 0; i < n; ++i) {
    dmax = std::max(dmax, std::fabs(d[i]));
    zmax = std::max(zmax, std::fabs(z[i]));
  }
  const double tol = 8.0 * DC_EPS * std::max(dmax, zmax);
  plan->K = 0;
  plan->rho = rho;
  plan->dlam.clear();
  plan->w.clear();
  plan->nd_cols.clear();
  plan->nd_type.clear();
  plan->defl_cols.clear();
  plan->defl_vals.clear();
  plan->rots.clear();
  if (rho * zmax <= tol) {
    for (int i = 0; i < n; ++i) {
      plan->defl_cols.push_back(order[i]);
      plan->defl_vals.push_back(d[order[i]]);
    }
    return;
  }
  std::vector<int> nd, defl;
  int pj = -1;
  for (int j = 0; j < n; ++j) {
    const int nj = order[j];
    if (rho * std::fabs(z[nj]) <= tol) {
      defl.push_back(nj);
      continue;
    }
    if (pj < 0) {
      pj = nj;
      continue;
    }
    double s = z[pj], c = z[nj];
    const double tau = std::hypot(c, s);
    double t = d[nj] - d[pj];
    c /= tau;
    s = -s / tau;
    if (std::fabs(t * c * s) <= tol) {
      z[nj] = tau;
      z[pj] = 0.0;
      if (coltyp[nj] != coltyp[pj]) coltyp[nj] = 2;
      plan->rots.push_back({pj, nj, c, s});
      t = d[pj] * c * c + d[nj] * s * s;
      d[nj] = d[pj] * s * s + d[nj] * c * c;
      d[pj] = t;
      defl.push_back(pj);
      pj = nj;
    } else {
      nd.push_back(pj);
      pj = nj;
    }
  }
  nd.push_back(pj);
  plan->K = (int)nd.size();
  for (int c : nd) {
    plan->dlam.push_back(d[c]);
    plan->w.push_back(z[c]);
    plan->nd_cols.push_back(c);
    plan->nd_type.push_back(coltyp[c]);
  }
  for (int c : defl) {
    plan->defl_cols.push_back(c);
    plan->defl_vals.push_back(d[c]);
  }
}

struct Node {
  int lo, mid, hi, height;
};

static int build_tree(int lo, int hi, std::vector<Node>& merges, std::vector<std::pair<int, int>>& leaves) {
  if (hi - lo <= LEAF) {
    leaves.push_back({lo, hi});
    return 0;
  }
  const int mid = (lo + hi) / 2;
  const int h = 1 + std::max(build_tree(lo, mid, merges, leaves), build_tree(mid, hi, merges, leaves));
  merges.push_back({lo, mid, hi, h});
  return h;
}

// =============================================================================================
// device kernels
// =============================================================================================
struct MergeDesc {
  int lo, mid, hi, K, c1, c2, c3, rot_beg, rot_end;
  double rho, sgn;
};

// ---- leaves: implicit QL with eigenvectors, one warp per leaf (size <= 32) --------------------
__global__ void __launch_bounds__(128)
    dc_leaf_kernel(const int2* __restrict__ leaves, int nleaves, const double* __restrict__ dadj,
                   const double* __restrict__ e, double* __restrict__ Q, long long ldq,
                   double* __restrict__ D, int* __restrict__ fail) {
  __shared__ double zs[4][LEAF][LEAF + 1];
  __shared__ double ds[4][LEAF];
  __shared__ double es[4][LEAF];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int leaf = blockIdx.x * 4 + wl;
  if (leaf >= nleaves) return;
  const int lo = leaves[leaf].x, m = leaves[leaf].y - lo;
  double(*z)[LEAF + 1] = zs[wl];
  double* d = ds[wl];
  double* ee = es[wl];
  for (int c = 0; c < m; ++c) z[lane][c] = (lane == c) ? 1.0 : 0.0;  // lane = row
  if (lane < m) {
    d[lane] = dadj[lo + lane];
    ee[lane] = (lane < m - 1) ? e[lo + lane] : 0.0;
  }
  __syncwarp();
  // All lanes run the scalar recurrences redundantly on the shared d/e (identical values),
  // lane 0 writes; each lane owns row `lane` of Z for the plane rotations.
  for (int l = 0; l < m; ++l) {
    int iter = 0;
    while (true) {
      int mm = l;
      for (; mm < m - 1; ++mm) {
        const double dd = fabs(d[mm]) + fabs(d[mm + 1]);
        if (fabs(ee[mm]) <= DBL_EPSILON * dd) break;
      }
      if (mm == l) break;
      if (++iter > 60) {
        if (lane == 0) atomicExch(fail, 1);
        break;
      }
      double g = (d[l + 1] - d[l]) / (2.0 * ee[l]);
      double r = hypot(g, 1.0);
      g = d[mm] - d[l] + ee[l] / (g + copysign(r, g));
      double s = 1.0, c = 1.0, p = 0.0;
      int i = mm - 1;
      bool early = false;
      for (; i >= l; --i) {
        double f = s * ee[i];
        const double b = c * ee[i];
        r = hypot(f, g);
        __syncwarp();
        if (lane == 0) ee[i + 1] = r;
        if (r == 0.0) {
          __syncwarp();
          if (lane == 0) {
            d[i + 1] -= p;
            ee[mm] = 0.0;
          }
          __syncwarp();
          early = true;
          break;
        }
        s = f / r;
        c = g / r;
        g = d[i + 1] - p;
        r = (d[i] - g) * s + 2.0 * c * b;
        p = s * r;
        __syncwarp();
        if (lane == 0) d[i + 1] = g + p;
        g = c * r - b;
        if (lane < m) {
          f = z[lane][i + 1];
          z[lane][i + 1] = s * z[lane][i] + c * f;
          z[lane][i] = c * z[lane][i] - s * f;
        }
        __syncwarp();
      }
      if (early) continue;
      __syncwarp();
      if (lane == 0) {
        d[l] -= p;
        ee[l] = g;
        ee[mm] = 0.0;
      }
      __syncwarp();
    }
  }
  __syncwarp();
  // selection sort ascending (eigenvalues + columns)
  for (int i = 0; i < m - 1; ++i) {
    int k = i;
    double p = d[i];
    for (int jx = i + 1; jx < m; ++jx)
      if (d[jx] < p) {
        k = jx;
        p = d[jx];
      }
    __syncwarp();
    if (k != i) {
      if (lane == 0) {
        d[k] = d[i];
        d[i] = p;
      }
      if (lane < m) {
        const double t = z[lane][i];
        z[lane][i] = z[lane][k];
        z[lane][k] = t;
      }
    }
    __syncwarp();
  }
  if (lane < m) {
    D[lo + lane] = d[lane];
    for (int c = 0; c < m; ++c) Q[(lo + lane) + (long long)(lo + c) * ldq] = z[lane][c];
  }
}

// ---- z = [last row of Q1 ; sgn * first row of Q2] ---------------------------------------------
__global__ void dc_gather_z_kernel(const MergeDesc* __restrict__ descs, const double* __restrict__ Q,
                                   long long ldq, double* __restrict__ z) {
  const MergeDesc m = descs[blockIdx.y];
  for (int i = m.lo + blockIdx.x * blockDim.x + threadIdx.x; i < m.hi; i += gridDim.x * blockDim.x)
    z[i] = (i < m.mid) ? Q[(m.mid - 1) + (long long)i * ldq] : m.sgn * Q[m.mid + (long long)i * ldq];
}

// ---- Givens rotations of the deflation step, applied in order; one thread per row --------------
__global__ void dc_rot_kernel(const MergeDesc* __restrict__ descs, const int* __restrict__ rp,
                              const int* __restrict__ rn, const double* __restrict__ rc,
                              const double* __restrict__ rs, double* __restrict__ Q, long long ldq) {
  const MergeDesc m = descs[blockIdx.y];
  if (m.rot_beg >= m.rot_end) return;
  for (int r = m.lo + blockIdx.x * blockDim.x + threadIdx.x; r < m.hi; r += gridDim.x * blockDim.x) {
    int cur = -1;
    double x = 0.0;
    for (int t = m.rot_beg; t < m.rot_end; ++t) {
      const int pj = rp[t], nj = rn[t];
      if (pj != cur) {
        if (cur >= 0) Q[r + (long long)cur * ldq] = x;
        x = Q[r + (long long)pj * ldq];
      }
      const double y = Q[r + (long long)nj * ldq];
      const double c = rc[t], s = rs[t];
      Q[r + (long long)pj * ldq] = c * x + s * y;  // pj is deflated: final value
      x = c * y - s * x;                           // nj continues (it is the next pj if chained)
      cur = nj;
    }
    if (cur >= 0) Q[r + (long long)cur * ldq] = x;
  }
}

// ---- G[:, lo + r] = Q[:, gsrc[lo + r]] over the merge's rows ----------------------------------
__global__ void dc_gather_cols_kernel(const MergeDesc* __restrict__ descs,
                                      const int* __restrict__ gsrc, const double* __restrict__ Q,
                                      double* __restrict__ Gm, long long ld) {
  const MergeDesc m = descs[blockIdx.y];
  const int nm = m.hi - m.lo;
  const long long total = (long long)nm * nm;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % nm), c = (int)(idx / nm);
    Gm[(m.lo + r) + (long long)(m.lo + c) * ld] = Q[(m.lo + r) + (long long)gsrc[m.lo + c] * ld];
  }
}
// ---- Q[:, lo+K : hi) = G[:, lo+K : hi)  (deflated eigenvectors go back unchanged) ----------------
__global__ void dc_copy_defl_kernel(const MergeDesc* __restrict__ descs, const double* __restrict__ Gm,
                                    double* __restrict__ Q, long long ld) {
  const MergeDesc m = descs[blockIdx.y];
  const int nm = m.hi - m.lo, nd = nm - m.K;
  const long long total = (long long)nm * nd;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % nm), c = m.K + (int)(idx / nm);
    Q[(m.lo + r) + (long long)(m.lo + c) * ld] = Gm[(m.lo + r) + (long long)(m.lo + c) * ld];
  }
}

// ---- secular equation: DC_LPR lanes per root (the poles are split over the lanes, butterfly sums give every
// lane the same bits, so the lanes of a root take identical decisions) ------------------------------------------
static constexpr int DC_LPR = 8;
__device__ __forceinline__ double lpr_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = 1; o < DC_LPR; o <<= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
__global__ void dc_secular_kernel(const MergeDesc* __restrict__ descs, const double* __restrict__ dlam,
                                  const double* __restrict__ w, int* __restrict__ org,
                                  double* __restrict__ mu, double* __restrict__ Dnew,
                                  int* __restrict__ fail) {
  const MergeDesc m = descs[blockIdx.y];
  const int K = m.K;
  const int jraw = (blockIdx.x * blockDim.x + threadIdx.x) / DC_LPR, sub = threadIdx.x % DC_LPR;
  if (K <= 0) return;
  const unsigned gmask = ((1u << DC_LPR) - 1u) << ((threadIdx.x & 31) & ~(DC_LPR - 1));  // the lanes of this root
  const bool active = jraw < K;            // inactive lane groups keep shuffling along with a clamped root
  const int j = active ? jraw : K - 1;
  const double* dl = dlam + m.lo;
  const double* ww = w + m.lo;
  const double rho = m.rho;
  if (K == 1) {
    if (active && sub == 0) {
      org[m.lo] = 0;
      mu[m.lo] = rho * ww[0] * ww[0];
      Dnew[m.lo] = dl[0] + mu[m.lo];
    }
    return;
  }
  const bool last = (j == K - 1);
  int o;
  double lo, hi;
  if (last) {
    o = K - 1;
    double s = 0.0;
    for (int i = sub; i < K; i += DC_LPR) s += ww[i] * ww[i];
    s = lpr_sum(s, gmask);
    lo = 0.0;
    hi = rho * s;
  } else {
    const double dj = dl[j];
    const double half = 0.5 * (dl[j + 1] - dj);
    double s = 0.0;
    for (int i = sub; i < K; i += DC_LPR) s += ww[i] * ww[i] / ((dl[i] - dj) - half);
    s = lpr_sum(s, gmask);
    const double gmid = 1.0 + rho * s;
    if (gmid >= 0.0) {
      o = j;
      lo = 0.0;
      hi = half;
    } else {
      o = j + 1;
      lo = -half;
      hi = 0.0;
    }
  }
  const double dorg = dl[o];
  const double Dj = dl[j] - dorg;
  const double Dj1 = last ? 0.0 : (dl[j + 1] - dorg);
  double x = 0.5 * (lo + hi);
  bool converged = false;
  for (int it = 0; it < 120; ++it) {
    double psi = 0.0, phi = 0.0, dpsi = 0.0, dphi = 0.0;
    for (int i = sub; i <= j; i += DC_LPR) {
      const double inv = 1.0 / ((dl[i] - dorg) - x);
      const double t = ww[i] * ww[i] * inv;
      psi += t;
      dpsi = fma(t, inv, dpsi);
    }
    for (int i = j + 1 + sub; i < K; i += DC_LPR) {
      const double inv = 1.0 / ((dl[i] - dorg) - x);
      const double t = ww[i] * ww[i] * inv;
      phi += t;
      dphi = fma(t, inv, dphi);
    }
    psi = lpr_sum(psi, gmask);
    phi = lpr_sum(phi, gmask);
    dpsi = lpr_sum(dpsi, gmask);
    dphi = lpr_sum(dphi, gmask);
    psi *= rho;
    phi *= rho;
    dpsi *= rho;
    dphi *= rho;
    const double g = 1.0 + psi + phi;
    const double err = 8.0 * DC_EPS * (1.0 + fabs(psi) + fabs(phi));
    if (fabs(g) <= err) {
      converged = true;
      break;
    }
    if (g < 0.0)
      lo = x;
    else
      hi = x;
    if ((hi - lo) <= 2.0 * DC_EPS * fmax(fabs(lo), fabs(hi))) {
      x = 0.5 * (lo + hi);
      converged = true;
      break;
    }
    double nx;
    if (last) {
      const double dj = Dj - x;
      const double s = dpsi * dj * dj;
      const double cst = 1.0 + psi - dpsi * dj + phi;
      nx = (cst > 0.0) ? (Dj + s / cst) : INFINITY;
    } else {
      const double dj = Dj - x, dj1 = Dj1 - x;
      const double s = dpsi * dj * dj, S = dphi * dj1 * dj1;
      const double cst = 1.0 + (psi - dpsi * dj) + (phi - dphi * dj1);
      const double qa = cst;
      const double qb = -(cst * (Dj + Dj1) + s + S);
      const double qc = cst * Dj * Dj1 + s * Dj1 + S * Dj;
      if (qa == 0.0) {
        nx = (qb != 0.0) ? (qc / -qb) : INFINITY;
      } else {
        const double disc = qb * qb - 4.0 * qa * qc;
        if (disc < 0.0) {
          nx = INFINITY;
        } else {
          const double q = -0.5 * (qb + copysign(sqrt(disc), qb));
          const double r1 = q / qa;
          const double r2 = (q != 0.0) ? (qc / q) : INFINITY;
          nx = (r1 > lo && r1 < hi) ? r1 : r2;
        }
      }
    }
    if (!(nx > lo && nx < hi)) nx = 0.5 * (lo + hi);  // also catches NaN / inf
    x = nx;
  }
  if (active && sub == 0) {
    if (!converged) atomicExch(fail, 2);
    org[m.lo + j] = o;
    mu[m.lo + j] = x;
    Dnew[m.lo + j] = dorg + x;
  }
}

__device__ __forceinline__ double dc_delta(const double* dl, const int* org, const double* mu, int i,
                                           int j) {
  return (dl[i] - dl[org[j]]) - mu[j];  // d_i - lambda_j without cancellation
}

// ---- Loewner / Gu-Eisenstat:  zhat_i^2 = prod_j (lambda_j - d_i) / prod_{j != i} (d_j - d_i) ------
__global__ void dc_zhat_kernel(const MergeDesc* __restrict__ descs, const double* __restrict__ dlam,
                               const double* __restrict__ w, const int* __restrict__ org,
                               const double* __restrict__ mu, double* __restrict__ zhat) {
  const MergeDesc m = descs[blockIdx.y];
  const int K = m.K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const double* dl = dlam + m.lo;
  const int* og = org + m.lo;
  const double* mm = mu + m.lo;
  const double di = dl[i];
  double p = dc_delta(dl, og, mm, i, i);
  for (int j = 0; j < K; ++j) {
    if (j == i) continue;
    p *= dc_delta(dl, og, mm, i, j) / (di - dl[j]);
  }
  zhat[m.lo + i] = copysign(sqrt(fabs(p)), w[m.lo + i]);
}

// ---- eigenvector matrix of the rank-one update, rows in grouped (type 1,2,3) order ------------------
__global__ void __launch_bounds__(256)
    dc_u_kernel(const MergeDesc* __restrict__ descs, const double* __restrict__ dlam,
                const int* __restrict__ org, const double* __restrict__ mu,
                const double* __restrict__ zhat, const int* __restrict__ grow,
                double* __restrict__ U, long long ldu, const int* __restrict__ colmap) {
  __shared__ double red[32];
  const MergeDesc m = descs[blockIdx.y];
  const int K = m.K;
  // colmap (root merge with truncation): only the wanted roots are formed, packed left
  const int j = colmap ? colmap[blockIdx.x] : (int)blockIdx.x;
  if (j >= K) return;
  const double* dl = dlam + m.lo;
  const int* og = org + m.lo;
  const double* mm = mu + m.lo;
  const double* zh = zhat + m.lo;
  const double dorg = dl[og[j]], muj = mm[j];
  double s = 0.0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const double u = zh[i] / ((dl[i] - dorg) - muj);
    s = fma(u, u, s);
  }
  s = block_sum(s, red);
  const double inv = 1.0 / sqrt(s);
  double* ucol = U + m.lo + (long long)(m.lo + (int)blockIdx.x) * ldu;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const double u = zh[i] / ((dl[i] - dorg) - muj);
    ucol[grow[m.lo + i]] = u * inv;
  }
}

// ---- lazy children of the root (see stedc): R = [U 0; 0 I] per merge, on a zeroed buffer that dc_u_kernel has filled
__global__ void dc_unit_defl_kernel(const MergeDesc* __restrict__ descs, double* __restrict__ R, long long ld) {
  const MergeDesc m = descs[blockIdx.y];
  const int nm = m.hi - m.lo;
  for (int c = m.K + blockIdx.x * blockDim.x + threadIdx.x; c < nm; c += gridDim.x * blockDim.x)
    R[(m.lo + c) + (long long)(m.lo + c) * ld] = 1.0;
}
// row `row` of G over the columns [c0, c0 + cnt) as a contiguous vector
__global__ void dc_row_kernel(const double* __restrict__ G, long long ld, int row, int c0, int cnt, double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x)
    out[i] = G[row + (long long)(c0 + i) * ld];
}
__global__ void dc_lazy_z_kernel(const double* __restrict__ zl, const double* __restrict__ zr, int mid, int n, double sgn,
                                 double* __restrict__ z) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    z[i] = (i < mid) ? zl[i] : sgn * zr[i - mid];
}

// ---- dst[:, dstcol[i]] = src[:, srccol[i]] --------------------------------------------------------------
__global__ void dc_assemble_kernel(const double* __restrict__ src, long long lds, const int* __restrict__ srccol,
                                   double* __restrict__ dst, long long ldd, const int* __restrict__ dstcol,
                                   int rows, int count) {
  const long long total = (long long)rows * count;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % rows), i = (int)(idx / rows);
    dst[r + (long long)dstcol[i] * ldd] = src[r + (long long)srccol[i] * lds];
  }
}

// =============================================================================================
// driver
// =============================================================================================
template <typename T>
static int upload(bk_ctx* ctx, DevBuf<T>& buf, const std::vector<T>& v) {
  BK_TRY(buf.ensure(std::max<size_t>(1, v.size())));
  if (!v.empty())
    BK_CUDA(cudaMemcpyAsync(buf.p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, ctx->stream));
  return BK_OK;
}

int stedc(bk_ctx* ctx, int n, const double* d_host, const double* e_host, double* evals_host,
          int max_want, double rel_thresh, int* n_want, double* Z, long long ldz,
          StedcStats* stats) {
  BK_REQUIRE(n >= 1, "stedc: n must be positive");
  max_want = std::max(0, std::min(max_want, n));
  std::vector<Node> merges;
  std::vector<std::pair<int, int>> leaves;
  const int height = build_tree(0, n, merges, leaves);
  std::vector<double> dadj(d_host, d_host + n), e(n, 0.0);
  for (int i = 0; i < n - 1; ++i) e[i] = e_host[i];
  for (const Node& m : merges) {
    const double b = std::fabs(e[m.mid - 1]);
    dadj[m.mid - 1] -= b;
    dadj[m.mid] -= b;
  }
  const long long ld = n;
  DevBuf<double> Q, Gm, U, Dcur, Dnew, zv, dadj_d, e_d, dlam_d, w_d, mu_d, zhat_d, rc_d, rs_d;
  DevBuf<int> org_d, gsrc_d, grow_d, rp_d, rn_d, fail_d;
  DevBuf<int2> leaves_d;
  DevBuf<MergeDesc> desc_d;
  DevBuf<GemmProb> probs_d;
  BK_TRY(Q.borrow(ctx->ws[2], (size_t)n * n));
  BK_TRY(Dcur.alloc(n));
  BK_TRY(Dnew.alloc(n));
  BK_TRY(zv.alloc(n));
  BK_TRY(fail_d.alloc(1));
  BK_CUDA(cudaMemsetAsync(Q.p, 0, sizeof(double) * (size_t)n * n, ctx->stream));
  BK_CUDA(cudaMemsetAsync(fail_d.p, 0, sizeof(int), ctx->stream));
  BK_TRY(upload(ctx, dadj_d, dadj));
  BK_TRY(upload(ctx, e_d, e));
  {
    std::vector<int2> lv(leaves.size());
    for (size_t i = 0; i < leaves.size(); ++i) lv[i] = make_int2(leaves[i].first, leaves[i].second);
    BK_TRY(upload(ctx, leaves_d, lv));
    dc_leaf_kernel<<<(unsigned)ceil_div((int64_t)lv.size(), 4), 128, 0, ctx->stream>>>(
        leaves_d.p, (int)lv.size(), dadj_d.p, e_d.p, Q.p, ld, Dcur.p, fail_d.p);
    BK_LAUNCHED(ctx);
    BK_CUDA(cudaGetLastError());
  }
  if (height > 0) {
    BK_TRY(Gm.borrow(ctx->ws[3], (size_t)n * n));
    BK_TRY(U.borrow(ctx->ws[4], (size_t)n * n));
  }
  // Lazy children of the root.  When only a few eigenvectors are wanted the two largest merge GEMMs below the root
  // (each (n/2) x K x K: 40 % of the whole D&C at n = 20 000) form eigenvector matrices Q1, Q2 of which the root then
  // uses a thin slice.  Instead the children stay FACTORED, Qc = diag(G1, G2) R with G = the gathered eigenvectors of
  // their own children and R = [U 0; 0 I] (n x n): the root takes its z from two rows G[row, :] R, applies its
  // deflation rotations and its column gather to R, multiplies R's gathered columns by the wanted columns of its own U
  // (the same batched GEMM as before, same zero structure) and only then goes through G: two (n/2) x want x (n/2)
  // products.
  bool lazy = false;
  {
    int cnt = 0;
    bool halves = false;
    for (const Node& m : merges)
      if (m.height == height - 1) ++cnt;
    for (const Node& m : merges)
      if (m.height == height && m.lo == 0 && m.hi == n) halves = true;
    int covered = 0;
    for (const Node& m : merges)
      if (m.height == height - 1) covered += m.hi - m.lo;
    const char* lz = getenv("BK_DC_LAZY");
    lazy = Z && height >= 2 && cnt == 2 && halves && covered == n && n >= 1024 && !(lz && atoi(lz) == 0);
  }
  // structurally possible; whether it pays is decided at the children's level, when their eigenvalues bound the number
  // of eigenvectors the root will be asked for (interlacing: at most one more than the children have above the threshold)
  const bool lazy_cand = lazy;
  lazy = false;
  DevBuf<double> GL, RL, zrow, Ycat;
  if (lazy_cand) {
    BK_TRY(GL.borrow(ctx->ws[5], (size_t)n * n));
    BK_TRY(RL.borrow(ctx->ws[6], (size_t)n * n));
    BK_TRY(zrow.alloc((size_t)3 * n));
  }
  std::vector<double> Dh(n), zh(n);
  if (stats) *stats = StedcStats();
  std::vector<int> order(n);
  int want = 0;
  bool root_done = false;
  // ascending eigenvalues + the reference's truncation rule on the leading max_want values
  // (R/bigKRLS_Rcpp_functions.R:190: lastkeeper = max(which(values >= eigtrunc*values[1])))
  auto finalize_values = [&](const std::vector<double>& D) -> int {
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return D[a] < D[b]; });
    for (int i = 0; i < n; ++i) {
      if (!std::isfinite(D[order[i]])) {
        set_error("stedc: non-finite eigenvalue");
        return BK_ERR_NUMERIC;
      }
      evals_host[i] = D[order[i]];
    }
    want = 0;
    for (int c = 0; c < max_want; ++c)
      if (evals_host[n - 1 - c] >= rel_thresh * evals_host[n - 1]) want = c + 1;
    return BK_OK;
  };

  for (int h = 1; h <= height; ++h) {
    const bool cand_lvl = lazy_cand && h == height - 1;  // the root's children: may stay factored
    const bool lazy_root = lazy && h == height;          // the root over factored children
    double* const Gp = cand_lvl ? GL.p : Gm.p;           // gathered columns of this level
    double* const Up = cand_lvl ? RL.p : U.p;            // rank-one eigenvector blocks of this level
    double* const Qsrc = lazy_root ? RL.p : Q.p;    // what the root rotates / gathers
    std::vector<Node> lvl;
    for (const Node& m : merges)
      if (m.height == h) lvl.push_back(m);
    const int nm = (int)lvl.size();
    // z vectors
    std::vector<MergeDesc> descs(nm);
    for (int i = 0; i < nm; ++i) {
      MergeDesc& md = descs[i];
      md.lo = lvl[i].lo;
      md.mid = lvl[i].mid;
      md.hi = lvl[i].hi;
      md.K = md.c1 = md.c2 = md.c3 = md.rot_beg = md.rot_end = 0;
      md.rho = 0.0;
      md.sgn = (e[lvl[i].mid - 1] < 0.0) ? -1.0 : 1.0;
    }
    BK_TRY(upload(ctx, desc_d, descs));
    int maxn = 0;
    for (const Node& m : lvl) maxn = std::max(maxn, m.hi - m.lo);
    if (lazy_root) {
      // z = [last row of Q1 ; sgn * first row of Q2] with Q = diag(G1, G2) R: two row-vector x matrix products
      const int mid = lvl[0].mid;
      dc_row_kernel<<<64, 256, 0, ctx->stream>>>(GL.p, ld, mid - 1, 0, mid, zrow.p + 2 * (size_t)n);
      BK_LAUNCHED(ctx);
      BK_TRY(gemm(ctx, true, false, mid, 1, mid, 1.0, RL.p, ld, zrow.p + 2 * (size_t)n, mid, 0.0, zrow.p, mid));
      dc_row_kernel<<<64, 256, 0, ctx->stream>>>(GL.p, ld, mid, mid, n - mid, zrow.p + 2 * (size_t)n);
      BK_LAUNCHED(ctx);
      BK_TRY(gemm(ctx, true, false, n - mid, 1, n - mid, 1.0, RL.p + mid + (long long)mid * ld, ld, zrow.p + 2 * (size_t)n,
                  n - mid, 0.0, zrow.p + n, n - mid));
      dc_lazy_z_kernel<<<64, 256, 0, ctx->stream>>>(zrow.p, zrow.p + n, mid, n, descs[0].sgn, zv.p);
      BK_LAUNCHED(ctx);
    } else {
      dc_gather_z_kernel<<<dim3((unsigned)ceil_div(maxn, 256), nm), 256, 0, ctx->stream>>>(desc_d.p, Q.p,
                                                                                           ld, zv.p);
      BK_LAUNCHED(ctx);
    }
    BK_CUDA(cudaMemcpyAsync(Dh.data(), Dcur.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaMemcpyAsync(zh.data(), zv.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));

    // host: deflation plans
    std::vector<double> dlam(n, 0.0), wv(n, 0.0), Dn(Dh), rc, rs;
    std::vector<int> gsrc(n), grow(n, 0), rp, rn;
    for (int i = 0; i < n; ++i) gsrc[i] = i;
    std::vector<GemmProb> probs;
    int maxK = 0, gm_max_m = 0, gm_max_n = 0;
    bool vec = true;
    MergePlan plan;
    for (int i = 0; i < nm; ++i) {
      MergeDesc& md = descs[i];
      const int lo = md.lo, nmm = md.hi - md.lo, n1 = md.mid - md.lo;
      for (int t = lo; t < md.hi; ++t) {
        if (!std::isfinite(Dh[t]) || !std::isfinite(zh[t])) {
          set_error("stedc: non-finite intermediate at level %d", h);
          return BK_ERR_NUMERIC;
        }
      }
      host_deflate(Dh.data() + lo, zh.data() + lo, nmm, n1, e[md.mid - 1], &plan);
      // zh was gathered with the sign already applied: host_deflate only needs |beta| -> rho
      const int K = plan.K;
      md.K = K;
      md.rho = plan.rho;
      md.rot_beg = (int)rp.size();
      for (const auto& r : plan.rots) {
        rp.push_back(lo + r.pj);
        rn.push_back(lo + r.nj);
        rc.push_back(r.c);
        rs.push_back(r.s);
      }
      md.rot_end = (int)rp.size();
      // grouped order of the non-deflated columns: type 1, 2, 3 (stable)
      int pos = 0;
      for (int typ = 1; typ <= 3; ++typ) {
        int cnt = 0;
        for (int t = 0; t < K; ++t)
          if (plan.nd_type[t] == typ) {
            gsrc[lo + pos] = lo + plan.nd_cols[t];
            grow[lo + t] = pos;
            ++pos;
            ++cnt;
          }
        if (typ == 1) md.c1 = cnt;
        if (typ == 2) md.c2 = cnt;
        if (typ == 3) md.c3 = cnt;
      }
      for (int t = 0; t < K; ++t) {
        dlam[lo + t] = plan.dlam[t];
        wv[lo + t] = plan.w[t];
      }
      for (int t = 0; t < nmm - K; ++t) {
        gsrc[lo + K + t] = lo + plan.defl_cols[t];
        Dn[lo + K + t] = plan.defl_vals[t];
      }
      maxK = std::max(maxK, K);
      if (K > 0) {
        const int n2 = nmm - n1;
        GemmProb p;
        p.alpha = 1.0;
        p.beta = 0.0;
        p.lower = 0;
        p.lda = p.ldb = p.ldc = ld;
        // top rows: Q[lo:mid, lo:lo+K] = G[lo:mid, lo:lo+c1+c2] * U[lo:lo+c1+c2, lo:lo+K]
        p.m = n1;
        p.n = K;
        p.k = md.c1 + md.c2;
        p.A = Gp + lo + (long long)lo * ld;
        p.B = Up + lo + (long long)lo * ld;
        p.C = Q.p + lo + (long long)lo * ld;
        probs.push_back(p);
        vec = vec && gemm_operands_vec_ok(p.A, ld, p.B, ld);
        // bottom rows: Q[mid:hi, lo:lo+K] = G[mid:hi, lo+c1:lo+K] * U[lo+c1:lo+K, lo:lo+K].
        // When lo + c1 is odd the B operand would start at an odd row and the whole batch would fall back to the
        // 8-byte operand loads (measured: 7 launches, 50 of the 70 ms of the D&C at N = 20 000).  The contraction may
        // start one index earlier instead: column lo+c1-1 of G is a type-1 column, exactly zero in the bottom rows.
        const int back = (((lo + md.c1) & 1) && md.c1 > 0) ? 1 : 0;
        p.m = n2;
        p.k = md.c2 + md.c3 + back;
        p.A = Gp + md.mid + (long long)(lo + md.c1 - back) * ld;
        p.B = Up + (lo + md.c1 - back) + (long long)lo * ld;
        p.C = Q.p + md.mid + (long long)lo * ld;
        probs.push_back(p);
        vec = vec && gemm_operands_vec_ok(p.A, ld, p.B, ld);
        gm_max_m = std::max(gm_max_m, std::max(n1, n2));
        gm_max_n = std::max(gm_max_n, K);
        if (stats) stats->merge_flops += 2LL * n1 * K * (md.c1 + md.c2) + 2LL * n2 * K * (md.c2 + md.c3);
      }
      if (stats && h == height) {
        stats->top_n = nmm;
        stats->top_k = K;
      }
    }
    // device: apply
    BK_TRY(upload(ctx, desc_d, descs));
    BK_TRY(upload(ctx, dlam_d, dlam));
    BK_TRY(upload(ctx, w_d, wv));
    BK_TRY(upload(ctx, gsrc_d, gsrc));
    BK_TRY(upload(ctx, grow_d, grow));
    BK_TRY(upload(ctx, rp_d, rp));
    BK_TRY(upload(ctx, rn_d, rn));
    BK_TRY(upload(ctx, rc_d, rc));
    BK_TRY(upload(ctx, rs_d, rs));
    BK_CUDA(cudaMemcpyAsync(Dnew.p, Dn.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    BK_TRY(org_d.ensure(n));
    BK_TRY(mu_d.ensure(n));
    BK_TRY(zhat_d.ensure(n));
    if (!rp.empty()) {
      dc_rot_kernel<<<dim3((unsigned)ceil_div(maxn, 128), nm), 128, 0, ctx->stream>>>(
          desc_d.p, rp_d.p, rn_d.p, rc_d.p, rs_d.p, Qsrc, ld);
      BK_LAUNCHED(ctx);
    }
    {
      const long long per = (long long)maxn * maxn;
      const unsigned gx = (unsigned)std::min<long long>(ceil_div(per, 256), 8LL * ctx->sm_count);
      dc_gather_cols_kernel<<<dim3(gx, nm), 256, 0, ctx->stream>>>(desc_d.p, gsrc_d.p, Qsrc, Gp, ld);
      BK_LAUNCHED(ctx);
    }
    if (maxK > 0) {
      dc_secular_kernel<<<dim3((unsigned)ceil_div((long long)maxK * DC_LPR, 128), nm), 128, 0, ctx->stream>>>(
          desc_d.p, dlam_d.p, w_d.p, org_d.p, mu_d.p, Dnew.p, fail_d.p);
      BK_LAUNCHED(ctx);
      if (h == height && nm == 1) {
        // Root merge: all eigenvalues are known now.  When only part of the eigenvectors is wanted
        // (eigtrunc / Neig, or values only) form just those columns of U: the root GEMM shrinks from
        // n x K x K to n x K x (wanted roots).
        BK_CUDA(cudaMemcpyAsync(Dh.data(), Dnew.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        BK_CUDA(cudaStreamSynchronize(ctx->stream));
        BK_TRY(finalize_values(Dh));
        if (lazy_root) BK_TRY(Ycat.alloc((size_t)n * std::max(1, want)));
        if (!Z || want < n) {
          const MergeDesc& md = descs[0];
          const int K = md.K;
          std::vector<int> colmap, rsrc, rdst, dsrc, ddst;
          for (int c = 0; Z && c < want; ++c) {
            const int pos = order[n - 1 - c];  // position in D: [0,K) roots, [K,n) deflated (columns of G)
            if (pos < K) {
              rsrc.push_back((int)colmap.size());
              colmap.push_back(pos);
              rdst.push_back(c);
            } else {
              dsrc.push_back(pos);
              ddst.push_back(c);
            }
          }
          const int nwr = (int)colmap.size();
          DevBuf<int> colmap_d, rsrc_d, rdst_d, dsrc_d, ddst_d;
          if (nwr > 0) {
            BK_TRY(upload(ctx, colmap_d, colmap));
            BK_TRY(upload(ctx, rsrc_d, rsrc));
            BK_TRY(upload(ctx, rdst_d, rdst));
            dc_zhat_kernel<<<dim3((unsigned)ceil_div(maxK, 128), nm), 128, 0, ctx->stream>>>(
                desc_d.p, dlam_d.p, w_d.p, org_d.p, mu_d.p, zhat_d.p);
            BK_LAUNCHED(ctx);
            dc_u_kernel<<<dim3((unsigned)nwr, 1), 256, 0, ctx->stream>>>(desc_d.p, dlam_d.p, org_d.p, mu_d.p,
                                                                         zhat_d.p, grow_d.p, U.p, ld, colmap_d.p);
            BK_LAUNCHED(ctx);
            for (auto& pr : probs) pr.n = nwr;
            BK_TRY(upload(ctx, probs_d, probs));
            BK_TRY(gemm_batched(ctx, false, false, probs_d.p, (int)probs.size(), gm_max_m, nwr, vec));
            if (stats)
              stats->merge_flops -= 2LL * (K - nwr) * ((long long)(md.mid - md.lo) * (md.c1 + md.c2) +
                                                     (long long)(md.hi - md.mid) * (md.c2 + md.c3));
            dc_assemble_kernel<<<(unsigned)std::min<long long>(ceil_div((long long)n * nwr, 256), 8LL * ctx->sm_count),
                                 256, 0, ctx->stream>>>(Q.p, ld, rsrc_d.p, lazy_root ? Ycat.p : Z, lazy_root ? (long long)n : ldz,
                                                        rdst_d.p, n, nwr);
            BK_LAUNCHED(ctx);
          }
          if (!dsrc.empty()) {
            BK_TRY(upload(ctx, dsrc_d, dsrc));
            BK_TRY(upload(ctx, ddst_d, ddst));
            dc_assemble_kernel<<<(unsigned)std::min<long long>(ceil_div((long long)n * (long long)dsrc.size(), 256),
                                                               8LL * ctx->sm_count),
                                 256, 0, ctx->stream>>>(Gm.p, ld, dsrc_d.p, lazy_root ? Ycat.p : Z, lazy_root ? (long long)n : ldz,
                                                        ddst_d.p, n, (int)dsrc.size());
            BK_LAUNCHED(ctx);
          }
          if (lazy_root && want > 0) {
            // the wanted columns are coordinates in the children's gathered bases: through G1, G2 now
            const int mid = descs[0].mid;
            BK_TRY(gemm(ctx, false, false, mid, want, mid, 1.0, GL.p, ld, Ycat.p, n, 0.0, Z, ldz));
            BK_TRY(gemm(ctx, false, false, n - mid, want, n - mid, 1.0, GL.p + mid + (long long)mid * ld, ld, Ycat.p + mid, n, 0.0,
                        Z + mid, ldz));
          }
          BK_CUDA(cudaGetLastError());
          BK_CUDA(cudaStreamSynchronize(ctx->stream));
          if (stats) stats->levels = h;
          root_done = true;
          break;
        }
      }
      dc_zhat_kernel<<<dim3((unsigned)ceil_div(maxK, 128), nm), 128, 0, ctx->stream>>>(
          desc_d.p, dlam_d.p, w_d.p, org_d.p, mu_d.p, zhat_d.p);
      BK_LAUNCHED(ctx);
      if (cand_lvl) {
        // the children's eigenvalues are known: bound the root's demand
        BK_CUDA(cudaMemcpyAsync(Dh.data(), Dnew.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
        BK_CUDA(cudaStreamSynchronize(ctx->stream));
        double dmax = -INFINITY;
        bool finite = true;
        for (int i = 0; i < n; ++i) {
          finite = finite && std::isfinite(Dh[i]);
          dmax = std::max(dmax, Dh[i]);
        }
        long long above = 0;
        for (int i = 0; i < n; ++i)
          if (Dh[i] >= rel_thresh * dmax) ++above;
        const long long bound = std::min<long long>(max_want, above + 1);
        lazy = finite && dmax > 0.0 && bound <= n / 4;
        if (lazy) BK_CUDA(cudaMemsetAsync(RL.p, 0, sizeof(double) * (size_t)n * n, ctx->stream));
      }
      dc_u_kernel<<<dim3((unsigned)maxK, nm), 256, 0, ctx->stream>>>(desc_d.p, dlam_d.p, org_d.p, mu_d.p,
                                                                     zhat_d.p, grow_d.p, Up, ld, nullptr);
      BK_LAUNCHED(ctx);
      if (!(cand_lvl && lazy)) {
        BK_TRY(upload(ctx, probs_d, probs));
        BK_TRY(gemm_batched(ctx, false, false, probs_d.p, (int)probs.size(), gm_max_m, gm_max_n, vec));
      }
    }
    if (cand_lvl && lazy) {
      dc_unit_defl_kernel<<<dim3((unsigned)ceil_div(maxn, 256), nm), 256, 0, ctx->stream>>>(desc_d.p, RL.p, ld);
      BK_LAUNCHED(ctx);
    } else {
      const long long per = (long long)maxn * maxn;
      const unsigned gx = (unsigned)std::min<long long>(ceil_div(per, 256), 8LL * ctx->sm_count);
      dc_copy_defl_kernel<<<dim3(gx, nm), 256, 0, ctx->stream>>>(desc_d.p, Gp, Q.p, ld);
      BK_LAUNCHED(ctx);
    }
    BK_CUDA(cudaGetLastError());
    std::swap(Dcur.p, Dnew.p);
    std::swap(Dcur.n, Dnew.n);
    if (stats) stats->levels = h;
    // the host vectors of this level must outlive the async copies
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
  }

  // final ordering
  int fail = 0;
  BK_CUDA(cudaMemcpyAsync(&fail, fail_d.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (!root_done)
    BK_CUDA(cudaMemcpyAsync(Dh.data(), Dcur.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (fail) {
    set_error("stedc: %s did not converge", fail == 1 ? "leaf QL iteration" : "secular equation");
    return BK_ERR_NUMERIC;
  }
  if (!root_done) {
    BK_TRY(finalize_values(Dh));
    if (want > 0 && Z) {
      std::vector<int> perm(want);
      for (int c = 0; c < want; ++c) perm[c] = order[n - 1 - c];  // descending
      DevBuf<int> perm_d;
      BK_TRY(upload(ctx, perm_d, perm));
      BK_TRY(gather_columns(ctx, Q.p, ld, n, want, perm_d.p, Z, ldz));
      BK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
  }
  if (n_want) *n_want = want;
  return BK_OK;
}

}  // namespace bk
