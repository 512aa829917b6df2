// Leave-one-out loss of the lambda search, many candidate lambdas per pass over Q.
//
// Reference: src/solveforc.cpp:36-58 computes, for ONE lambda, row i of the lower triangle of
// Ginv = Q diag(1/(ev+lambda)) Q' with N gemv's, then c = Ginv y, Le = sum (c_i/Ginv_ii)^2.
// Closed form:  c = Q ((Q'y) o w),  diag(Ginv) = (Q o Q) w,  w = 1/(ev+lambda).
// For L candidates this is the contraction  [Q , QoQ] x [ (z o W) ; W ]  (n x k x 2L) - one
// read of Q (8nk bytes) serves every candidate, which is what makes the speculative
// golden-section tree (fit.cu) cost one pass per 3 iterations.
//
// Kernel 1: thread = two rows, CTA = 512 rows x one k-slice; the 2L weight columns of the slice
//           are staged in shared memory (broadcast 16-byte reads: one load feeds 4 FMAs - with one row per
//           thread and 8-byte reads the kernel was bound by the shared-memory pipe at 80 us per pass), Q is
//           read coalesced along rows, 2L FP64 FMAs per loaded element.  Partials go to a [ksplit][2L][rows]
//           buffer.  The FMA sequence per (row, candidate) is unchanged, so are the bits.
// Kernel 2: fixed-order reduction over the k-slices, (c/d)^2, per-CTA partial sums; the CTA that
//           finishes last adds them in a fixed order -> Le[L].  Deterministic (the only atomic is
//           the arrival counter), so the discrete decisions of the golden section are
//           reproducible run to run.
// Roofline: HBM, 8 n k bytes per pass.
#include "common.cuh"
#include "kernels.cuh"

namespace bk {

struct LamPack {
  double lam[16];
};

static constexpr int LOO_JC = 32;  // eigen-columns staged per chunk

template <int L>
__global__ void __launch_bounds__(256)
    loo_partial_kernel(const double* __restrict__ Q, long long ldq, int n_rows, int k,
                       const double* __restrict__ ev, const double* __restrict__ z, LamPack lp,
                       int ksplit, double* __restrict__ part) {
  constexpr int LP = (L + 1) & ~1;                   // padded to pairs for the 16-byte reads
  __shared__ __align__(16) double g1[LOO_JC][LP];    // z_j * w_jl
  __shared__ __align__(16) double g2[LOO_JC][LP];    // w_jl
  const int row0 = blockIdx.x * 512 + threadIdx.x, row1 = row0 + 256;
  const int per = (k + ksplit - 1) / ksplit;
  const int kb = min(k, (int)blockIdx.y * per), ke = min(k, kb + per);
  double c0[LP], d0[LP], c1[LP], d1[LP];
#pragma unroll
  for (int l = 0; l < LP; ++l) c0[l] = d0[l] = c1[l] = d1[l] = 0.0;
  const double* q0 = Q + (row0 < n_rows ? row0 : 0);
  const double* q1 = Q + (row1 < n_rows ? row1 : 0);

  for (int j0 = kb; j0 < ke; j0 += LOO_JC) {
    const int jc = min(LOO_JC, ke - j0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < LOO_JC * LP; idx += 256) {
      const int jj = idx / LP, l = idx % LP;
      double w = 0.0, zz = 0.0;
      if (jj < jc && l < L) {
        w = 1.0 / (ev[j0 + jj] + lp.lam[l]);
        zz = z[j0 + jj];
      }
      g1[jj][l] = zz * w;
      g2[jj][l] = w;
    }
    __syncthreads();
    // the loads of 8 eigen-columns are issued before their FMAs (the kernel is otherwise bound by the latency of two
    // dependent loads per column)
    for (int j8 = 0; j8 < jc; j8 += 8) {
      double va[8], vb[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ok = j8 + u < jc;
        va[u] = ok ? q0[(long long)(j0 + j8 + u) * ldq] : 0.0;
        vb[u] = ok ? q1[(long long)(j0 + j8 + u) * ldq] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int jj = j8 + u;
        if (jj >= jc) break;
        const double va2 = va[u] * va[u], vb2 = vb[u] * vb[u];
        const double2* p1 = reinterpret_cast<const double2*>(&g1[jj][0]);
        const double2* p2 = reinterpret_cast<const double2*>(&g2[jj][0]);
#pragma unroll
        for (int l = 0; l < LP; l += 2) {
          const double2 a1 = p1[l / 2], a2 = p2[l / 2];
          c0[l] = fma(va[u], a1.x, c0[l]);
          c0[l + 1] = fma(va[u], a1.y, c0[l + 1]);
          d0[l] = fma(va2, a2.x, d0[l]);
          d0[l + 1] = fma(va2, a2.y, d0[l + 1]);
          c1[l] = fma(vb[u], a1.x, c1[l]);
          c1[l + 1] = fma(vb[u], a1.y, c1[l + 1]);
          d1[l] = fma(vb2, a2.x, d1[l]);
          d1[l + 1] = fma(vb2, a2.y, d1[l + 1]);
        }
      }
    }
  }
  if (row0 < n_rows) {
    double* o = part + (size_t)blockIdx.y * (2 * L) * (size_t)n_rows + row0;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      o[(size_t)(2 * l) * n_rows] = c0[l];
      o[(size_t)(2 * l + 1) * n_rows] = d0[l];
    }
  }
  if (row1 < n_rows) {
    double* o = part + (size_t)blockIdx.y * (2 * L) * (size_t)n_rows + row1;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      o[(size_t)(2 * l) * n_rows] = c1[l];
      o[(size_t)(2 * l + 1) * n_rows] = d1[l];
    }
  }
}

template <int L>
__global__ void __launch_bounds__(256)
    loo_finish_kernel(const double* __restrict__ part, int n_rows, int ksplit,
                      double* __restrict__ blockpart, double* __restrict__ coeffs,
                      unsigned* __restrict__ done, double* __restrict__ Le) {
  __shared__ bool last;
  const int row = blockIdx.x * 256 + threadIdx.x;
  double e[L];
#pragma unroll
  for (int l = 0; l < L; ++l) e[l] = 0.0;
  if (row < n_rows) {
#pragma unroll
    for (int l = 0; l < L; ++l) {
      double c = 0.0, d = 0.0;
      for (int s = 0; s < ksplit; ++s) {
        const double* o = part + (size_t)s * (2 * L) * (size_t)n_rows + row;
        c += o[(size_t)(2 * l) * n_rows];
        d += o[(size_t)(2 * l + 1) * n_rows];
      }
      const double r = c / d;  // src/solveforc.cpp:57
      e[l] = r * r;
      if (l == 0 && coeffs) coeffs[row] = c;
    }
  }
  // per-CTA sums of all L candidates with ONE barrier (a block_sum per candidate cost 3 barriers each): warp sums by
  // shuffles, then thread l adds the 8 warp sums in warp order - fixed order, deterministic
  {
    __shared__ double wsum[8][L];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const double s = warp_sum(e[l]);
      if (lane == 0) wsum[wid][l] = s;
    }
    __syncthreads();
    if (threadIdx.x < L) {
      double s = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += wsum[w8][threadIdx.x];
      blockpart[(size_t)blockIdx.x * L + threadIdx.x] = s;
    }
  }
  // the CTA that finishes last adds the per-CTA sums in a FIXED order (deterministic: the golden section is a
  // discrete decision path) - this used to be a third launch
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < L) {
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(blockpart + (size_t)b * L + threadIdx.x);
    Le[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) *done = 0u;  // ready for the next pass
}

template <int L>
static int loo_run(bk_ctx* ctx, const double* Q, long long ldq, int n_rows, int k,
                   const double* ev, const double* z, const LamPack& lp, double* Le_dev,
                   double* coeffs) {
  const int rb = (int)ceil_div(n_rows, 256);
  int ksplit = (int)ceil_div(2LL * ctx->sm_count * 1024, (long long)rb * 256);
  ksplit = std::max(1, std::min(ksplit, std::min(64, (k + 63) / 64)));
  const size_t part_elems = (size_t)ksplit * 2 * L * (size_t)n_rows;
  const size_t bp_elems = (size_t)rb * L;
  BK_TRY(ctx->gemm_ws.ensure(part_elems + bp_elems));
  double* part = ctx->gemm_ws.p;
  double* bp = part + part_elems;
  dim3 grid((unsigned)ceil_div(n_rows, 512), ksplit);
  loo_partial_kernel<L><<<grid, 256, 0, ctx->stream>>>(Q, ldq, n_rows, k, ev, z, lp, ksplit, part);
  BK_LAUNCHED(ctx);
  loo_finish_kernel<L><<<rb, 256, 0, ctx->stream>>>(part, n_rows, ksplit, bp, coeffs, ctx->counters.p, Le_dev);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

int loo_batch(bk_ctx* ctx, const double* Q, long long ldq, int n_rows, int k, const double* ev,
              const double* z, const double* lambdas_host, int nlam, double* Le_dev,
              double* coeffs) {
  BK_REQUIRE(nlam >= 1 && nlam <= 16, "loo_batch: nlam must be in 1..16 (got %d)", nlam);
  BK_REQUIRE(coeffs == nullptr || nlam == 1, "loo_batch: coefficients need a single lambda");
  LamPack lp;
  for (int i = 0; i < 16; ++i) lp.lam[i] = lambdas_host[std::min(i, nlam - 1)];
  if (nlam == 1) return loo_run<1>(ctx, Q, ldq, n_rows, k, ev, z, lp, Le_dev, coeffs);
  if (nlam == 2) return loo_run<2>(ctx, Q, ldq, n_rows, k, ev, z, lp, Le_dev, coeffs);
  if (nlam <= 4) return loo_run<4>(ctx, Q, ldq, n_rows, k, ev, z, lp, Le_dev, coeffs);
  if (nlam <= 8) return loo_run<8>(ctx, Q, ldq, n_rows, k, ev, z, lp, Le_dev, coeffs);
  return loo_run<16>(ctx, Q, ldq, n_rows, k, ev, z, lp, Le_dev, coeffs);
}

}  // namespace bk
