// Q2 back-transformation for MANY eigenvectors: the stage-2 reflectors grouped into compact-WY blocks and
// applied with the DMMA GEMM (the sliding-window kernel of sb2st.cu moves every window row through shared
// memory once per 4 reflectors, which is the better trade for a few hundred columns only).
//
// Block (J, t) = reflectors (sweep j, hop t) for the 64 sweeps j = 64J .. 64J+63.  Reflector (j, t) acts on rows
// [j+1+64t, j+65+64t), so the block touches the 128 rows starting at R0 = 64 (J + t) (row R0 itself is never
// touched: it is included so that every operand stays 16-byte aligned).  V_blk is the 128 x 64 parallelogram of
// the reflectors padded with zeros, Q_blk = H_{64J} ... H_{64J+63} = I - V T V' (larft, forward columnwise),
// Z[R0:R0+128, :] <- Q_blk Z[R0:R0+128, :].
// Order: reflectors are generated sweep by sweep; applying Q2 needs the reverse.  Blocks (J, t) and (J', t')
// commute unless their rows overlap; a valid order is J descending, t ascending, and blocks with equal
// 3 (Jmax - J) + t have disjoint rows, which gives ~4 n / 64 batched steps (verified against the sequential
// application in the numpy prototype).
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"

namespace bk {

static constexpr int QB = 64;         // bandwidth = reflector length = sweeps per block
static constexpr int QR = 2 * QB;     // rows per block

struct Q2Block {
  int J, t, R0, rows;  // rows = min(QR, n - R0)
};

__global__ void q2b_build_kernel(const double* __restrict__ VV, const double* __restrict__ TAU, int maxhops, int n,
                                 const Q2Block* __restrict__ blocks, double* __restrict__ Vb,
                                 double* __restrict__ taub) {
  const Q2Block bl = blocks[blockIdx.x];
  double* V = Vb + (size_t)blockIdx.x * QR * QB;
  for (int idx = threadIdx.x; idx < QR * QB; idx += blockDim.x) {
    const int r = idx % QR, c = idx / QR;
    const int j = QB * bl.J + c;
    const int i = r - 1 - c;  // index inside reflector (j, t): its first row is R0 + c + 1
    double v = 0.0;
    if (i >= 0 && i < QB && j <= n - 3 - QB * bl.t && bl.R0 + r < n) v = VV[(size_t)(bl.R0 + r) + (size_t)j * n];
    V[idx] = v;
  }
  if (threadIdx.x < QB) {
    const int j = QB * bl.J + threadIdx.x;
    taub[(size_t)blockIdx.x * QB + threadIdx.x] = (j <= n - 3 - QB * bl.t) ? TAU[bl.t + (size_t)j * maxhops] : 0.0;
  }
}

// T (64 x 64 upper triangular) per block from S = V'V and tau; 1024 threads, 16 per row of T (as sb_larft_kernel)
__global__ void q2b_larft_kernel(const double* __restrict__ Sb, const double* __restrict__ taub,
                                 double* __restrict__ Tb) {
  extern __shared__ double ts[];  // T row-major (64 x 64) then S (64 x 64)
  __shared__ double s_tau[QB];
  double* ss = ts + QB * QB;
  const double* S = Sb + (size_t)blockIdx.x * QB * QB;
  double* T = Tb + (size_t)blockIdx.x * QB * QB;
  const int r = threadIdx.x >> 4, part = threadIdx.x & 15;
  for (int idx = threadIdx.x; idx < QB * QB; idx += blockDim.x) {
    ts[idx] = 0.0;
    ss[idx] = S[idx];
  }
  if (threadIdx.x < QB) s_tau[threadIdx.x] = taub[(size_t)blockIdx.x * QB + threadIdx.x];
  __syncthreads();
  for (int i = 0; i < QB; ++i) {
    const double ti = s_tau[i];
    double acc = 0.0;
    if (r < i)
      for (int q = r + part; q < i; q += 16) acc = fma(ts[r * QB + q], ss[q + i * QB], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    __syncthreads();
    if (part == 0) {
      if (r < i) ts[r * QB + i] = -ti * acc;
      if (r == i) ts[i * QB + i] = ti;
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < QB * QB; idx += blockDim.x) T[idx] = ts[(idx % QB) * QB + idx / QB];
}

template <typename T>
static int upload_vec(bk_ctx* ctx, DevBuf<T>& buf, const std::vector<T>& v) {
  BK_TRY(buf.alloc(std::max<size_t>(1, v.size())));
  if (!v.empty())
    BK_CUDA(cudaMemcpyAsync(buf.p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, ctx->stream));
  return BK_OK;
}

// Z (n x k, ld ldz) <- Q2 Z
int q2_apply_blocked(bk_ctx* ctx, const double* VV, const double* TAU, int maxhops, int n, double* Z, long long ldz,
                     int k) {
  if (n < 3 || k <= 0) return BK_OK;
  // ---- enumerate the blocks, grouped into steps of mutually independent blocks --------------------------
  const int Jmax = (n - 3) / QB;
  std::vector<std::vector<Q2Block>> steps;
  for (int J = Jmax; J >= 0; --J) {
    for (int t = 0; QB * J <= n - 3 - QB * t; ++t) {
      const size_t sigma = (size_t)3 * (Jmax - J) + t;
      if (steps.size() <= sigma) steps.resize(sigma + 1);
      Q2Block b;
      b.J = J;
      b.t = t;
      b.R0 = QB * (J + t);
      b.rows = std::min(QR, n - b.R0);
      steps[sigma].push_back(b);
    }
  }
  std::vector<Q2Block> blocks;
  std::vector<int> step_off{0};
  size_t max_step = 0;
  for (const auto& s : steps) {
    blocks.insert(blocks.end(), s.begin(), s.end());
    step_off.push_back((int)blocks.size());
    max_step = std::max(max_step, s.size());
  }
  const int nb = (int)blocks.size();
  if (nb == 0) return BK_OK;
  DevBuf<Q2Block> blocks_d;
  DevBuf<double> Vb, Yb, taub, Sb, Tb, W2;
  DevBuf<GemmProb> pS_d, pY_d, p1_d, p3_d;
  BK_TRY(upload_vec(ctx, blocks_d, blocks));
  BK_TRY(Vb.alloc((size_t)nb * QR * QB));
  BK_TRY(taub.alloc((size_t)nb * QB));
  BK_TRY(Sb.alloc((size_t)nb * QB * QB));
  BK_TRY(Tb.alloc((size_t)nb * QB * QB));
  BK_TRY(Yb.alloc((size_t)nb * QR * QB));
  BK_TRY(W2.alloc(max_step * QB * (size_t)k));
  q2b_build_kernel<<<nb, 256, 0, ctx->stream>>>(VV, TAU, maxhops, n, blocks_d.p, Vb.p, taub.p);
  BK_LAUNCHED(ctx);
  // ---- T factors: S = V'V (batched GEMM), then the larft recurrence ---------------------------------------
  std::vector<GemmProb> pS(nb), pY(nb), p1(nb), p3(nb);
  for (int s = 0; s + 1 < (int)step_off.size(); ++s) {
    for (int i = step_off[s]; i < step_off[s + 1]; ++i) {
      const Q2Block& b = blocks[i];
      const int slot = i - step_off[s];
      double* V = Vb.p + (size_t)i * QR * QB;
      double* T = Tb.p + (size_t)i * QB * QB;
      double* Y = Yb.p + (size_t)i * QR * QB;
      double* w2 = W2.p + (size_t)slot * QB * k;
      double* Zr = Z + b.R0;
      GemmProb g{};
      g.lower = 0;
      // S = V'V (64 x 64 x 128)
      g.A = V; g.lda = QR; g.B = V; g.ldb = QR; g.C = Sb.p + (size_t)i * QB * QB; g.ldc = QB;
      g.m = QB; g.n = QB; g.k = QR; g.alpha = 1.0; g.beta = 0.0;
      pS[i] = g;
      // Y = V T'  (128 x 64 x 64): Q Z = Z - V (T V' Z) = Z - V (Y' Z)
      g.A = V; g.lda = QR; g.B = T; g.ldb = QB; g.C = Y; g.ldc = QR;
      g.m = QR; g.n = QB; g.k = QB; g.alpha = 1.0; g.beta = 0.0;
      pY[i] = g;
      // W2 = Y' Z[R0:R0+rows, :]   (64 x k x rows)
      g.A = Y; g.lda = QR; g.B = Zr; g.ldb = ldz; g.C = w2; g.ldc = QB;
      g.m = QB; g.n = k; g.k = b.rows; g.alpha = 1.0; g.beta = 0.0;
      p1[i] = g;
      // Z[R0:R0+rows, :] -= V W2   (rows x k x 64)
      g.A = V; g.lda = QR; g.B = w2; g.ldb = QB; g.C = Zr; g.ldc = ldz;
      g.m = b.rows; g.n = k; g.k = QB; g.alpha = -1.0; g.beta = 1.0;
      p3[i] = g;
    }
  }
  BK_TRY(upload_vec(ctx, pS_d, pS));
  BK_TRY(upload_vec(ctx, p1_d, p1));
  BK_TRY(upload_vec(ctx, pY_d, pY));
  BK_TRY(upload_vec(ctx, p3_d, p3));
  const bool vecZ = gemm_operands_vec_ok(Z, ldz, Z, ldz);  // R0 is even: block rows keep the alignment of Z
  BK_TRY(gemm_batched(ctx, true, false, pS_d.p, nb, QB, QB, true));
  BK_CUDA(cudaFuncSetAttribute(q2b_larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sizeof(double) * 2 * QB * QB)));
  q2b_larft_kernel<<<nb, 16 * QB, sizeof(double) * 2 * QB * QB, ctx->stream>>>(Sb.p, taub.p, Tb.p);
  BK_LAUNCHED(ctx);
  BK_TRY(gemm_batched(ctx, false, true, pY_d.p, nb, QR, QB, true));
  // ---- apply, one batched launch triple per step ------------------------------------------------------------
  for (int s = 0; s + 1 < (int)step_off.size(); ++s) {
    const int o = step_off[s], cnt = step_off[s + 1] - o;
    if (cnt == 0) continue;
    BK_TRY(gemm_batched(ctx, true, false, p1_d.p + o, cnt, QB, k, vecZ));
    BK_TRY(gemm_batched(ctx, false, false, p3_d.p + o, cnt, QR, k, true));
  }
  BK_CUDA(cudaGetLastError());
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}

}  // namespace bk
