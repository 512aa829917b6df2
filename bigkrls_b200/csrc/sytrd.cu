// Blocked Householder tridiagonalisation of a dense symmetric matrix (lower triangle),
// A = Q T Q'.  Stage one of the full eigensolver that replaces arma::eig_sym -> LAPACK dsyevd
// (reference src/eigen.cpp:24).  Written from the published algorithm (Dongarra/Sorensen/
// Hammarling blocked reduction: panel of nb reflectors with deferred rank-2nb update).
//
// One PERSISTENT COOPERATIVE kernel per panel (one CTA per SM: 16 consumer warps + 1 TMA producer
// warp, software grid barrier, 3 barriers per column).  Per column j = j0 + c of the panel:
//
//   PB  Householder (beta, tau, v) computed redundantly by every CTA from the norm partials
//       -- barrier --
//   PC  y = A22 v with A22 the UN-updated trailing matrix.  The lower triangle is cut into
//       512 x 64 tiles, streamed in 512 x 8 slices by TMA bulk copies (cp.async.bulk, 4 KB
//       contiguous per column) into a 5-stage mbarrier ring (160 KB in flight per SM) - register
//       loads could not keep enough bytes in flight to reach the HBM roofline (profiles/);
//       every element is read from HBM ONCE and used for both y_i += a_ij v_j and
//       y_j += a_ij v_i (4 m^2 bytes per column instead of 8 m^2 - this read is the HBM roofline
//       of the whole eigensolver).  Work items are S row tiles x 1 column tile: the transposed
//       sums stay in registers across the S tile-rows and are reduced over the 512 rows once per
//       8-column slice (warp reduce-scatter + shared memory); every partial result is written
//       to an owned slot (no atomics => bitwise reproducible).  Also u1 = W'v, u2 = V'v and the
//       partials of v'A22v.
//       -- barrier --
//   PD  8 threads per row: y = sum of partials, w = tau (y - V u1 - W u2) + alpha v with
//       alpha = -tau/2 (w.v) obtained from v'A22v and u1.u2 (no extra reduction), then the
//       SAME threads immediately update the next column a = A[:,j+1] - V W[j+1,:]' - W V[j+1,:]'
//       and accumulate its norm partials.
//       -- barrier --
//
// After the panel: A22 -= [V W][W V]'  (lower tiles only) on the DMMA GEMM.
//
// Layout: A column-major n x n; panel workspace P = [V | W | V] (n x 3nb) so that the rank-2nb
// update is ONE GEMM  A22 -= P[:,0:2nb] * P[:,nb:3nb]'.  Reflector j is stored LAPACK-style in
// A[j+2:, j] (v[j+1] = 1 implicit; A[j+1, j] is left holding alpha), e[j] = beta, d[j] = A[j,j].
//
// Data written by other CTAs earlier in the same launch is always read through L2 (__ldcg).
#include <cstdlib>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"

namespace bk {

static constexpr int TR = 512;    // tile rows  = consumer threads per CTA (one row per thread)
static constexpr int TC = 64;     // tile columns (8 slices of 8 columns)
static constexpr int NCONS = 512; // consumer threads (16 warps)
static constexpr int NTH = 544;   // + 1 producer warp that drives the TMA ring
static constexpr int NBMAX = 64;  // panel width supported by the shared-memory staging
static constexpr int SEGMAX = 256; // max row segments per column (n <= 131072)
static constexpr int NS = 5;      // TMA ring stages
static constexpr int SLICE = 8;   // columns per stage
static constexpr int STAGE_DOUBLES = TR * SLICE;  // 512 x 8 doubles = 32 KB per stage
static constexpr int VCS_MAX = TC;                // v entries of one column tile
// dynamic shared memory: ring | st[8][16][8] | vcs[256] | full[NS] | empty[NS]
static constexpr size_t SMEM_BYTES =
    sizeof(double) * ((size_t)NS * STAGE_DOUBLES + 8 * 16 * 8 + VCS_MAX) + sizeof(uint64_t) * 2 * NS;

struct SytrdArgs {
  double* A;
  long long lda;
  int n, j0, nb;
  double* P;       // n x 3nb  [V | W | V]
  double* d;
  double* e;
  double* tau;
  double* part;    // per-CTA partial sums: [0,G) column norm, [G,2G) v'A22v
  double* dots;    // 2 * nb: u1 = W'v, u2 = V'v
  double* ypart;   // [G][n]    direct partials, owned by (CTA, row)
  double* ytpart;  // [NSB][n]  transposed partials, owned by (row segment sb)
  unsigned* barrier;
  long long* prof;  // optional per-phase cycle counters of CTA 0 (debug)
};

// sum over the 8 lanes of an octet (lanes 8q .. 8q+7); only those lanes need to be converged
__device__ __forceinline__ double oct_sum(double v) {
  const unsigned mask = 0xFFu << (threadIdx.x & 24);
  v += __shfl_xor_sync(mask, v, 1);
  v += __shfl_xor_sync(mask, v, 2);
  v += __shfl_xor_sync(mask, v, 4);
  return v;
}

// Tiling of the trailing matrix of column j.  Tiles are ABSOLUTE (row tiles of 512, column tiles
// of 64) so that loads stay aligned; rows/columns <= j are neutralised by the zeros of v.
// A tile is 512 rows x 64 columns: one thread per row, so each column of a tile is a 4 KB
// contiguous run in memory (DRAM-page friendly) and the direct product needs no cross-thread
// reduction at all.
struct ColGeom {
  int t0r, t0c, Tr, Tc, S, NSB;
};

__device__ __forceinline__ ColGeom col_geom(int n, int j, int G) {
  ColGeom g;
  g.t0r = (j + 1) / TR;
  g.t0c = (j + 1) / TC;
  g.Tr = (n + TR - 1) / TR - g.t0r;
  g.Tc = (n + TC - 1) / TC - g.t0c;
  // work item = S row tiles x 1 column tile; about half of the Tc x ceil(Tr/S) items carry work and
  // we want >= ~6 of them per CTA so that the static round-robin assignment stays balanced
  const long long budget = ((long long)g.Tr * g.Tc) / (12LL * G);
  g.S = (budget >= 4) ? 4 : (budget >= 2 ? 2 : 1);
  g.NSB = (g.Tr + g.S - 1) / g.S;
  return g;
}

// (mbarrier / TMA bulk-copy primitives: common.cuh)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(NCONS) : "memory"); }

// Work items: (row segment sb of S row tiles) x (one column tile bc), only those that reach below the
// diagonal.  They are enumerated sb-major / bc-minor and CTA b takes the CONTIGUOUS range
// [b M / G, (b+1) M / G): perfectly balanced, and consecutive items of a CTA share the row segment, so
// the direct sums stay in registers across them and a row receives one partial per CTA that touched
// its segment (~M/G column tiles each) instead of one per column tile.
__device__ __forceinline__ int seg_count(int n, const ColGeom& g, int sb) {
  const int Tr_abs = (n + TR - 1) / TR, Tc_abs = (n + TC - 1) / TC;
  const int seg_hi = min(Tr_abs, g.t0r + sb * g.S + g.S);
  const int max_bc = min(Tc_abs - 1, (seg_hi * TR - 1) / TC);
  return max_bc - g.t0c + 1;  // >= 1
}
__device__ __forceinline__ long long range_start(long long M, int G, int b) { return (M * b) / G; }

struct ItemGeom {
  int bc, sb, seg_lo, seg_hi;
};
__device__ __forceinline__ ItemGeom item_geom(int n, const ColGeom& g, int sb, int bc) {
  ItemGeom it;
  const int Tr_abs = (n + TR - 1) / TR;
  it.bc = bc;
  it.sb = sb;
  it.seg_lo = g.t0r + sb * g.S;
  it.seg_hi = min(Tr_abs, it.seg_lo + g.S);
  return it;
}
// iterator over the items of this CTA (identical in the producer and in every consumer)
struct ItemIter {
  int sb, bc, left_in_seg;
  long long left;
  __device__ __forceinline__ void init(int n, const ColGeom& g, long long M, int G, int b) {
    const long long lo = range_start(M, G, b), hi = range_start(M, G, b + 1);
    left = hi - lo;
    sb = 0;
    long long off = 0;
    int cnt = seg_count(n, g, 0);
    while (left > 0 && off + cnt <= lo) {
      off += cnt;
      ++sb;
      cnt = seg_count(n, g, sb);
    }
    bc = g.t0c + (int)(lo - off);
    left_in_seg = cnt - (int)(lo - off);
  }
  __device__ __forceinline__ void advance(int n, const ColGeom& g) {
    --left;
    ++bc;
    if (--left_in_seg == 0 && left > 0) {
      ++sb;
      bc = g.t0c;
      left_in_seg = seg_count(n, g, sb);
    }
  }
};
// A (row tile, 8-column slice) step is streamed iff the tile reaches down to the slice's first column.
__device__ __forceinline__ bool step_active(const ItemGeom& it, int s, int col0) {
  return (it.seg_lo + s < it.seg_hi) && ((it.seg_lo + s) * TR + TR - 1 >= col0);
}

// ---- PC producer: one thread walks the same (item, slice, row tile) sequence as the consumers and
// keeps the ring full: per step 8 bulk copies of one 4 KB column run each. ------------------------------
__device__ __forceinline__ void symv_producer(const SytrdArgs& a, const ColGeom& g, long long M, int G,
                                              double* ring, uint64_t* full, uint64_t* empty,
                                              unsigned& stage, unsigned& parity) {
  const int n = a.n;
  const long long lda = a.lda;
  ItemIter iter;
  iter.init(n, g, M, G, blockIdx.x);
  for (; iter.left > 0; iter.advance(n, g)) {
    const ItemGeom it = item_geom(n, g, iter.sb, iter.bc);
    for (int sl = 0; sl < 8; ++sl) {
      const int col0 = it.bc * TC + sl * SLICE;
      const int ncol = max(0, min(SLICE, n - col0));  // slices past the last column carry no bytes
      for (int s = 0; s < g.S; ++s) {
        if (!step_active(it, s, col0)) continue;
        mbar_wait(&empty[stage], parity ^ 1u);
        const int row0 = (it.seg_lo + s) * TR;
        const int rows = min(TR, n - row0);
        const unsigned bytes = (unsigned)(((rows + 1) & ~1) * sizeof(double));  // 16-byte multiple
        mbar_expect_tx(&full[stage], bytes * (unsigned)ncol);
        double* dst = ring + (size_t)stage * STAGE_DOUBLES;
        const double* src = a.A + row0 + (long long)col0 * lda;
        for (int k = 0; k < ncol; ++k) tma_bulk_g2s(dst + k * TR, src + (long long)k * lda, bytes, &full[stage]);
        if (++stage == NS) {
          stage = 0;
          parity ^= 1u;
        }
      }
    }
  }
}

// sum of 8 values over the 32 lanes of a warp by recursive halving (9 exchanges instead of 40):
// on return lane L holds the warp total of value ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1).
__device__ __forceinline__ double warp_reduce_scatter8(const double (&v)[8], int lane) {
  const bool h1 = lane & 16, h2 = lane & 8, h3 = lane & 4;
  double w4[4], w2[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = h1 ? v[i] : v[i + 4], keep = h1 ? v[i + 4] : v[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = h2 ? w4[i] : w4[i + 2], keep = h2 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const double send = h3 ? w2[0] : w2[1], keep = h3 ? w2[1] : w2[0];
  double w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
  w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
  return w1;
}

// ---- PC consumers: 512 threads, thread = row of the tile ---------------------------------------------------
template <int S>
__device__ __forceinline__ void symv_consumer(const SytrdArgs& a, const ColGeom& g, long long M, int G,
                                              const double* __restrict__ vcol, double& vav,
                                              const double* ring, uint64_t* full, uint64_t* empty,
                                              double (*st)[16][8], double* vcs, unsigned& stage,
                                              unsigned& parity) {
  const int n = a.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ridx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  ItemIter iter;
  iter.init(n, g, M, G, blockIdx.x);
  int cur_sb = -1;
  double dsum[S], vr[S];
  int row[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    dsum[s] = 0.0;
    vr[s] = 0.0;
    row[s] = 0;
  }
  // direct sums of the finished row segment: one thread per row, nothing to reduce; one owned slot
  // per (CTA, row)
  auto flush = [&]() {
    if (cur_sb < 0) return;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (row[s] < n && row[s] / TR - g.t0r < (cur_sb + 1) * g.S)
        a.ypart[(size_t)blockIdx.x * n + row[s]] = dsum[s];
      dsum[s] = 0.0;
    }
  };
  for (; iter.left > 0; iter.advance(n, g)) {
    const ItemGeom it = item_geom(n, g, iter.sb, iter.bc);
    consumer_sync();  // previous item's readers of vcs / st are done
    if (threadIdx.x < TC) {
      const int col = it.bc * TC + threadIdx.x;
      vcs[threadIdx.x] = (col < n) ? __ldcg(vcol + col) : 0.0;
    }
    if (it.sb != cur_sb) {
      flush();
      cur_sb = it.sb;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        row[s] = (it.seg_lo + s) * TR + threadIdx.x;
        vr[s] = (it.seg_lo + s < it.seg_hi && row[s] < n) ? __ldcg(vcol + row[s]) : 0.0;
        if (it.seg_lo + s >= it.seg_hi) row[s] = n;  // tile row beyond the matrix: never stored
      }
    }
    consumer_sync();
#pragma unroll 1
    for (int sl = 0; sl < 8; ++sl) {
      const int col0 = it.bc * TC + sl * SLICE;
      double vc[8], accT[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        vc[k] = vcs[sl * SLICE + k];
        accT[k] = 0.0;
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        if (!step_active(it, s, col0)) continue;  // uniform
        mbar_wait(&full[stage], parity);
        const double* sm = ring + (size_t)stage * STAGE_DOUBLES + threadIdx.x;
        double av[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) av[k] = sm[k * TR];
        double t = 0.0;
        if ((it.seg_lo + s) * TR > col0 + 7) {
          // strictly below the diagonal
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            t = fma(av[k], vc[k], t);
            accT[k] = fma(av[k], vr[s], accT[k]);
          }
        } else {
          // tile touches the diagonal: lower triangle only
          const int r = (it.seg_lo + s) * TR + threadIdx.x;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (r >= col0 + k) t = fma(av[k], vc[k], t);                 // includes the diagonal once
            if (r > col0 + k) accT[k] = fma(av[k], vr[s], accT[k]);      // strictly lower only
          }
        }
        dsum[s] += t;
        vav = fma(vr[s], t, vav);
        // the values have been consumed (so the shared-memory reads are complete): hand the
        // stage back to the TMA producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == NS) {
          stage = 0;
          parity ^= 1u;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) vav = fma(vc[k], accT[k], vav);
      const double tot = warp_reduce_scatter8(accT, lane);
      if ((lane & 3) == 0) st[sl][warp][ridx] = tot;
    }
    // transposed sums of this column tile: fixed-order sum over the 16 warps, one owned slot
    consumer_sync();
    if (threadIdx.x < TC) {
      const int col = it.bc * TC + threadIdx.x;
      if (col < n) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < 16; ++w) acc += st[threadIdx.x >> 3][w][threadIdx.x & 7];
        a.ytpart[(size_t)it.sb * n + col] = acc;
      }
    }
  }
  flush();
}

// ---- PD helper: final w of row r, computed by the 8 lanes of an octet (all return the same value) ---------
__device__ __forceinline__ double final_w(const SytrdArgs& a, const ColGeom& g, long long M, int G,
                                          const int* s_off, int r, int sub, int c, const double* V,
                                          const double* W, const double* vcol, const double* u1,
                                          const double* u2, double tau, double alpha) {
  const int n = a.n;
  double acc = 0.0;
  // CTAs whose item range intersects this row's segment (s_off: prefix sums of the segment sizes)
  const int sbr = (r / TR - g.t0r) / g.S;
  const long long o1 = s_off[sbr], o2 = s_off[sbr + 1] - 1;
  const int b_lo = (int)(((o1 + 1) * G + M - 1) / M) - 1, b_hi = (int)(((o2 + 1) * G + M - 1) / M) - 1;
  for (int b = b_lo + sub; b <= b_hi; b += 8)
    if (range_start(M, G, b + 1) > range_start(M, G, b)) acc += __ldcg(a.ypart + (size_t)b * n + r);
  const int sb_min = ((r / TC) / 8 - g.t0r) / g.S;
  for (int i = sb_min + sub; i < g.NSB; i += 8) acc += __ldcg(a.ytpart + (size_t)i * n + r);
  for (int q = sub; q < c; q += 8)
    acc -= __ldcg(V + (size_t)q * n + r) * u1[q] + __ldcg(W + (size_t)q * n + r) * u2[q];
  acc = oct_sum(acc);
  return tau * acc + alpha * __ldcg(vcol + r);
}

__global__ void __launch_bounds__(NTH, 1) sytrd_panel_kernel(SytrdArgs a) {
  __shared__ double red[32];
  __shared__ double s_u1[NBMAX], s_u2[NBMAX], s_wrow[NBMAX + 1], s_vrow[NBMAX + 1];
  __shared__ double s_scal[2];
  __shared__ int s_off[SEGMAX + 1];
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  double* ring = reinterpret_cast<double*>(dyn_smem);                       // NS x 8 x 512
  double(*st)[16][8] = reinterpret_cast<double(*)[16][8]>(ring + (size_t)NS * STAGE_DOUBLES);
  double* vcs = ring + (size_t)NS * STAGE_DOUBLES + 8 * 16 * 8;
  uint64_t* full = reinterpret_cast<uint64_t*>(vcs + VCS_MAX);
  uint64_t* empty = full + NS;
  unsigned stage = 0, parity = 0;  // ring position (same sequence in the producer and every consumer)
  const int n = a.n, nb = a.nb;
  double* A = a.A;
  const long long lda = a.lda;
  double* V = a.P;
  double* W = a.P + (size_t)nb * n;
  double* V2 = a.P + (size_t)2 * nb * n;
  const int G = gridDim.x;
  const int gid = blockIdx.x * NTH + threadIdx.x;
  const int gthreads = G * NTH;
  const int oct = gid >> 3, sub = gid & 7, nocts = gthreads >> 3;  // PD: 8 lanes per row
  unsigned epoch = 0;
  // ring: zero (stale-but-finite data is multiplied by v = 0 at the matrix edge), barriers, fences
  for (int i = threadIdx.x; i < NS * STAGE_DOUBLES; i += NTH) ring[i] = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NCONS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  long long tprev = clock64();
#define PROF(slot)                                                   \
  do {                                                               \
    if (a.prof && gid == 0) {                                        \
      const long long tnow = clock64();                              \
      a.prof[slot] += tnow - tprev;                                  \
      tprev = tnow;                                                  \
    }                                                                \
  } while (0)

  // ---- prologue: column j0 needs no update inside this panel: d and the norm partials -------------------
  {
    const int j = a.j0;
    double ss = 0.0;
    for (int r = j + gid; r < n; r += gthreads) {
      const double v = A[r + (long long)j * lda];
      if (r == j) a.d[j] = v;
      if (r > j + 1) ss += v * v;
    }
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) a.part[blockIdx.x] = ss;
    grid_barrier(a.barrier, epoch);
  }

  for (int c = 0; c < nb; ++c) {
    const int j = a.j0 + c;
    if (j >= n - 1) break;  // the last diagonal element has no reflector
    // ---------------- PB: Householder vector ------------------------------------------------
    double xnorm2 = 0.0;  // same fixed-shape reduction in every CTA => identical value everywhere
    for (int b = threadIdx.x; b < G; b += NTH) xnorm2 += __ldcg(a.part + b);
    xnorm2 = block_sum(xnorm2, red);
    const double alpha0 = __ldcg(A + (j + 1) + (long long)j * lda);
    double beta, tau, scale;
    if (xnorm2 == 0.0) {
      beta = alpha0;
      tau = 0.0;
      scale = 0.0;
    } else {
      beta = -copysign(sqrt(alpha0 * alpha0 + xnorm2), alpha0);
      tau = (beta - alpha0) / beta;
      scale = 1.0 / (alpha0 - beta);
    }
    if (gid == 0) {
      a.e[j] = beta;
      a.tau[j] = tau;
    }
    for (int r = j + 1 + gid; r < n; r += gthreads) {
      double v;
      if (r == j + 1) {
        // A[j+1, j] keeps alpha: every CTA reads it at the top of this phase, so it must not be
        // overwritten here (beta lives in e[j]; the back-transform uses an explicit 1).
        v = 1.0;
      } else {
        v = __ldcg(A + r + (long long)j * lda) * scale;
        A[r + (long long)j * lda] = v;
      }
      V[(size_t)c * n + r] = v;
      V2[(size_t)c * n + r] = v;
    }
    PROF(0);
    grid_barrier(a.barrier, epoch);
    PROF(1);

    const double* vcol = V + (size_t)c * n;
    const ColGeom g = col_geom(n, j, G);
    // prefix sums of the per-segment item counts (needed by PD; M = total number of items)
    if (threadIdx.x == 0) {
      int acc = 0;
      for (int sb = 0; sb < g.NSB; ++sb) {
        s_off[sb] = acc;
        acc += seg_count(n, g, sb);
      }
      s_off[g.NSB] = acc;
    }
    __syncthreads();
    const long long M = s_off[g.NSB];
    // ---------------- PC: u = [W V]'v,  y = A22 v,  v'A22v ---------------------------------------------
    const long long tpc0 = clock64();
    for (int q = blockIdx.x; q < 2 * c; q += G) {
      const double* col = (q < c) ? (W + (size_t)q * n) : (V + (size_t)(q - c) * n);
      double s = 0.0;
      for (int r = j + 1 + threadIdx.x; r < n; r += NTH) s += __ldcg(col + r) * __ldcg(vcol + r);
      s = block_sum(s, red);
      if (threadIdx.x == 0) a.dots[(q < c) ? q : (nb + q - c)] = s;
    }
    PROF(2);
    {
      double vav = 0.0;
      if (threadIdx.x < NCONS) {
        if (g.S == 4)
          symv_consumer<4>(a, g, M, G, vcol, vav, ring, full, empty, st, vcs, stage, parity);
        else if (g.S == 2)
          symv_consumer<2>(a, g, M, G, vcol, vav, ring, full, empty, st, vcs, stage, parity);
        else
          symv_consumer<1>(a, g, M, G, vcol, vav, ring, full, empty, st, vcs, stage, parity);
      } else if (threadIdx.x == NCONS) {
        symv_producer(a, g, M, G, ring, full, empty, stage, parity);
      }
      stage = __shfl_sync(0xffffffffu, stage, 0);  // producer warp: lanes follow lane 0
      parity = __shfl_sync(0xffffffffu, parity, 0);
      vav = block_sum(vav, red);
      if (threadIdx.x == 0) a.part[G + blockIdx.x] = vav;
    }
    if (a.prof && threadIdx.x == 0) a.prof[16 + blockIdx.x] += clock64() - tpc0;
    PROF(3);
    grid_barrier(a.barrier, epoch);
    PROF(4);

    // ---------------- PD: w (final) for own rows, then the next column's update -------------------------
    {
      double vav = 0.0;
      for (int b = threadIdx.x; b < G; b += NTH) vav += __ldcg(a.part + G + b);
      vav = block_sum(vav, red);
      for (int q = threadIdx.x; q < c; q += NTH) {
        s_u1[q] = __ldcg(a.dots + q);
        s_u2[q] = __ldcg(a.dots + nb + q);
        s_wrow[q] = __ldcg(W + (size_t)q * n + (j + 1));
        s_vrow[q] = __ldcg(V + (size_t)q * n + (j + 1));
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double uu = 0.0;
        for (int q = 0; q < c; ++q) uu = fma(s_u1[q], s_u2[q], uu);
        // w.v = tau (v'A22v - 2 u1.u2)  =>  alpha = -tau/2 (w.v)
        s_scal[0] = -0.5 * tau * (tau * (vav - 2.0 * uu));
      }
      __syncthreads();
      PROF(5);
      const double alpha = s_scal[0];
      // W[j+1, c] is needed by every thread for the next column: each CTA recomputes it with the
      // owner's exact procedure (same octet split, same order) so the value is bitwise identical.
      if (threadIdx.x < 8) {
        const double wj1 = final_w(a, g, M, G, s_off, j + 1, threadIdx.x, c, V, W, vcol, s_u1, s_u2, tau, alpha);
        if (threadIdx.x == 0) {
          s_wrow[c] = wj1;
          s_vrow[c] = 1.0;
        }
      }
      __syncthreads();
      PROF(7);
      const bool next = (c + 1 < nb);  // the next panel's prologue handles its own first column
      const int jn = j + 1;
      double ss = 0.0;
      for (int r = j + 1 + oct; r < n; r += nocts) {
        const double w = final_w(a, g, M, G, s_off, r, sub, c, V, W, vcol, s_u1, s_u2, tau, alpha);
        if (sub == 0) W[(size_t)c * n + r] = w;
        if (next) {
          double acc = (sub == 0) ? A[r + (long long)jn * lda] : 0.0;
          for (int q = sub; q < c; q += 8)
            acc -= __ldcg(V + (size_t)q * n + r) * s_wrow[q] + __ldcg(W + (size_t)q * n + r) * s_vrow[q];
          if (sub == (c & 7)) acc -= __ldcg(vcol + r) * s_wrow[c] + w * s_vrow[c];
          acc = oct_sum(acc);
          if (sub == 0) {
            A[r + (long long)jn * lda] = acc;
            if (r == jn) a.d[jn] = acc;
            if (r > jn + 1) ss += acc * acc;
          }
        }
      }
      PROF(8);
      if (next) {
        ss = block_sum(ss, red);
        if (threadIdx.x == 0) a.part[blockIdx.x] = ss;
        grid_barrier(a.barrier, epoch);
      }
      PROF(6);
    }
  }
}

int sytrd_lower(bk_ctx* ctx, double* A, long long lda, int n, double* d, double* e, double* tau,
                int nb, SytrdStats* stats) {
  if (n <= 0) return BK_OK;
  if (n == 1) {
    BK_CUDA(cudaMemcpyAsync(d, A, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return BK_OK;
  }
  BK_REQUIRE(nb >= 1 && nb <= NBMAX, "sytrd: panel width must be in 1..%d", NBMAX);
  BK_REQUIRE(ceil_div(n, TR) <= SEGMAX, "sytrd: n too large (max %d)", SEGMAX * TR);
  int occ = 0;
  BK_REQUIRE(lda % 2 == 0, "sytrd: leading dimension must be even (16-byte aligned columns for TMA)");
  BK_CUDA(cudaFuncSetAttribute(sytrd_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)SMEM_BYTES));
  BK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sytrd_panel_kernel, NTH, SMEM_BYTES));
  BK_REQUIRE(occ >= 1, "sytrd: kernel does not fit on an SM");
  const int G = ctx->sm_count;  // one CTA per SM
  const int T = (int)ceil_div(n, TC);
  DevBuf<double> P, part, dots, ypart, ytpart;
  BK_TRY(P.alloc((size_t)3 * nb * n));
  BK_TRY(part.alloc((size_t)2 * G));
  BK_TRY(dots.alloc((size_t)2 * nb));
  BK_TRY(ypart.alloc((size_t)G * n));
  BK_TRY(ytpart.alloc((size_t)T * n));  // NSB <= T
  BK_TRY(ctx->barrier.ensure(4));
  DevBuf<long long> prof;
  const bool do_prof = getenv("BK_SYTRD_PROF") != nullptr;
  if (do_prof) {
    BK_TRY(prof.alloc(16 + 2 * G));
    BK_CUDA(cudaMemsetAsync(prof.p, 0, (16 + 2 * G) * sizeof(long long), ctx->stream));
  }

  // per-panel CUDA events on the launching stream: the roofline numbers of bench.py are the
  // sum of these kernel durations against the algorithmic bytes 4 m_j^2 per column
  const int npanels = (int)ceil_div(n, nb);
  std::vector<cudaEvent_t> ev0(stats ? npanels : 0), ev1(stats ? npanels : 0), ev2(stats ? npanels : 0);
  for (auto& x : ev0) BK_CUDA(cudaEventCreate(&x));
  for (auto& x : ev1) BK_CUDA(cudaEventCreate(&x));
  for (auto& x : ev2) BK_CUDA(cudaEventCreate(&x));
  int pi = 0;
  for (int j0 = 0; j0 < n; j0 += nb, ++pi) {
    BK_CUDA(cudaMemsetAsync(P.p, 0, sizeof(double) * (size_t)3 * nb * n, ctx->stream));
    BK_CUDA(cudaMemsetAsync(ctx->barrier.p, 0, sizeof(unsigned) * 4, ctx->stream));
    SytrdArgs args;
    args.A = A;
    args.lda = lda;
    args.n = n;
    args.j0 = j0;
    args.nb = nb;
    args.P = P.p;
    args.d = d;
    args.e = e;
    args.tau = tau;
    args.part = part.p;
    args.dots = dots.p;
    args.ypart = ypart.p;
    args.ytpart = ytpart.p;
    args.barrier = ctx->barrier.p;
    args.prof = prof.p;
    void* kargs[] = {&args};
    if (stats) BK_CUDA(cudaEventRecord(ev0[pi], ctx->stream));
    BK_CUDA(cudaLaunchCooperativeKernel((void*)sytrd_panel_kernel, dim3(G), dim3(NTH), kargs, SMEM_BYTES,
                                        ctx->stream));
    if (stats) BK_CUDA(cudaEventRecord(ev1[pi], ctx->stream));
    BK_LAUNCHED(ctx);
    const int jn = j0 + nb;
    if (jn < n) {
      // A22 -= [V W] [W V]'   (lower tiles only)
      const int m = n - jn;
      BK_TRY(gemm(ctx, false, true, m, m, 2 * nb, -1.0, P.p + jn, n, P.p + (size_t)nb * n + jn, n,
                  1.0, A + jn + (long long)jn * lda, lda, true));
    }
    if (stats) BK_CUDA(cudaEventRecord(ev2[pi], ctx->stream));
  }
  BK_CUDA(cudaStreamSynchronize(ctx->stream));  // workspaces are freed on return
  if (do_prof) {
    std::vector<long long> h(16 + 2 * G);
    BK_CUDA(cudaMemcpy(h.data(), prof.p, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    const char* names[9] = {"PB", "wait(B2)", "PC dots", "PC symv", "wait(B3)", "PD setup", "wait(B1)", "PD wj1", "PD rows"};
    long long tot = 0;
    for (int i = 0; i < 9; ++i) tot += h[i];
    fprintf(stderr, "[sytrd prof n=%d] CTA0 cycles:", n);
    for (int i = 0; i < 9; ++i) fprintf(stderr, " %s=%.1f%%", names[i], 100.0 * h[i] / (double)tot);
    fprintf(stderr, " total=%.3f Gcyc\n", tot * 1e-9);
    long long mn = 1LL << 62, mx = 0, sm = 0;
    for (int b = 0; b < G; ++b) {
      mn = std::min(mn, h[16 + b]);
      mx = std::max(mx, h[16 + b]);
      sm += h[16 + b];
    }
    fprintf(stderr, "[sytrd prof] per-CTA PC (dots+symv) cycles: min %.3f mean %.3f max %.3f Gcyc\n", mn * 1e-9,
            sm * 1e-9 / G, mx * 1e-9);
  }
  if (stats) {
    stats->launches = npanels;
    stats->kernel_seconds = 0.0;
    stats->update_seconds = 0.0;
    stats->algorithmic_bytes = 0.0;
    for (int i = 0; i < npanels; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev0[i], ev1[i]);
      stats->kernel_seconds += 1e-3 * ms;
      cudaEventElapsedTime(&ms, ev1[i], ev2[i]);
      stats->update_seconds += 1e-3 * ms;
      cudaEventDestroy(ev0[i]);
      cudaEventDestroy(ev1[i]);
      cudaEventDestroy(ev2[i]);
    }
    for (int j = 0; j < n - 1; ++j) {
      const double m = (double)(n - 1 - j);
      stats->algorithmic_bytes += 4.0 * m * m;  // lower triangle (incl. diagonal ~ m^2/2 doubles)
    }
  }
  return BK_OK;
}

}  // namespace bk
