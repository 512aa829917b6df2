// FP64 GEMM for sm_100a on the DMMA (mma.sync m8n8k4 f64) tensor path.
//
// C (m x n, col-major) = alpha * op(A) * op(B) + beta * C
//
// This is the workhorse behind every dense contraction on the hot path: vcov = M Q'
// (reference src/crossprod.cpp:53), the tall-skinny K-pass K [1 c X X.c] (replaces the
// N x N temporaries of src/bigderiv_v3.cpp:90-106), Q'y / Q'R, the rank-2b trailing update
// and the block-reflector back-transform of the eigensolver, and the divide-and-conquer
// merge GEMMs (replacing LAPACK dsyevd behind src/eigen.cpp:24).
//
// Design: CTA tile BM x BN x 16, 3-stage cp.async pipeline into padded shared memory
// (row strides = 4 mod 16 doubles so that the 8-byte fragment loads of a half-warp hit 16
// distinct bank pairs), 8 warps, each warp holds (BM/WM) x (BN/WN) accumulators in
// registers as 8x8 DMMA tiles.  FP64 has no tcgen05 kind, so TMEM/UMMA do not apply; TMA is
// not used because operands are arbitrary sub-matrix views (odd offsets / leading
// dimensions break the 16-byte global alignment TMA needs) - cp.async handles both the
// 16-byte fast path and the 8-byte general path.
//
// Variants: transposed operands, batched (array of problem descriptors, blockIdx.z),
// lower-triangle-only tile skipping (SYRK/SYR2K-like updates), deterministic split-K for
// short-and-wide reductions (partials to a workspace, fixed-order reduce).
#include <cmath>
#include <cstdlib>
#include "common.cuh"
#include "dgemm.cuh"

namespace bk {

static constexpr int BK = 16;
static constexpr int kStages = 3;  // default pipeline depth (template parameter STAGES of the kernel)
static constexpr int KPAD = BK + 4;  // K-major row stride (doubles), 20 = 4 mod 16

template <int BMN>
struct TileSize {
  static constexpr int mn_major = BK * (BMN + 4);  // [k][mn], stride BMN+4
  static constexpr int k_major = BMN * KPAD;       // [mn][k], stride 20
  static constexpr int max = (mn_major > k_major) ? mn_major : k_major;
};

// ---- global -> shared tile loaders ---------------------------------------------------------
// MN-major: logical element (mn, kk) lives at src[(mn0+mn) + (k0+kk)*ld] -> dst[kk*(BMN+4)+mn]
template <int BMN, int NT, bool VEC>
__device__ __forceinline__ void load_mn_major(double* dst, const double* __restrict__ src,
                                              long long ld, int mn0, int k0, int mn_max,
                                              int k_max) {
  constexpr int S = BMN + 4;
  if (VEC) {
    constexpr int PAIRS = BMN / 2;
#pragma unroll
    for (int idx = threadIdx.x; idx < PAIRS * BK; idx += NT) {
      const int mn = (idx % PAIRS) * 2, kk = idx / PAIRS;
      const int gm = mn0 + mn, gk = k0 + kk;
      int valid = (gk < k_max) ? max(0, min(2, mn_max - gm)) : 0;
      const double* g = valid ? (src + gm + (long long)gk * ld) : src;
      cp_async16(dst + kk * S + mn, g, valid * 8);
    }
  } else {
#pragma unroll
    for (int idx = threadIdx.x; idx < BMN * BK; idx += NT) {
      const int mn = idx % BMN, kk = idx / BMN;
      const int gm = mn0 + mn, gk = k0 + kk;
      const bool valid = (gk < k_max) && (gm < mn_max);
      const double* g = valid ? (src + gm + (long long)gk * ld) : src;
      cp_async8(dst + kk * S + mn, g, valid ? 8 : 0);
    }
  }
}
// K-major: logical element (mn, kk) lives at src[(k0+kk) + (mn0+mn)*ld] -> dst[mn*20 + kk]
template <int BMN, int NT, bool VEC>
__device__ __forceinline__ void load_k_major(double* dst, const double* __restrict__ src,
                                             long long ld, int mn0, int k0, int mn_max,
                                             int k_max) {
  if (VEC) {
    constexpr int PAIRS = BK / 2;
#pragma unroll
    for (int idx = threadIdx.x; idx < BMN * PAIRS; idx += NT) {
      const int kk = (idx % PAIRS) * 2, mn = idx / PAIRS;
      const int gm = mn0 + mn, gk = k0 + kk;
      int valid = (gm < mn_max) ? max(0, min(2, k_max - gk)) : 0;
      const double* g = valid ? (src + gk + (long long)gm * ld) : src;
      cp_async16(dst + mn * KPAD + kk, g, valid * 8);
    }
  } else {
#pragma unroll
    for (int idx = threadIdx.x; idx < BMN * BK; idx += NT) {
      const int kk = idx % BK, mn = idx / BK;
      const int gm = mn0 + mn, gk = k0 + kk;
      const bool valid = (gk < k_max) && (gm < mn_max);
      const double* g = valid ? (src + gk + (long long)gm * ld) : src;
      cp_async8(dst + mn * KPAD + kk, g, valid ? 8 : 0);
    }
  }
}

template <int BM, int BN, int WM, int WN, bool TA, bool TB, bool VEC, int STAGES = 3, bool PFC = false>
__global__ void __launch_bounds__(WM* WN * 32, PFC ? 3 : 0)
    dgemm_kernel(const GemmProb single, const GemmProb* __restrict__ probs, int splits,
                 double* __restrict__ ws) {
  constexpr int NT = WM * WN * 32;
  constexpr int WTM = BM / WM, WTN = BN / WN;  // warp tile
  constexpr int MT = WTM / 8, NTL = WTN / 8;   // 8x8 DMMA tiles per warp
  constexpr int SA = BM + 4, SB = BN + 4;
  constexpr int A_STAGE = TileSize<BM>::max, B_STAGE = TileSize<BN>::max;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;

  const GemmProb pr = probs ? probs[blockIdx.z] : single;
  const int tiles_m = (pr.m + BM - 1) / BM, tiles_n = (pr.n + BN - 1) / BN;
  if ((int)blockIdx.x >= tiles_m * tiles_n) return;
  const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
  const int m0 = tm * BM, n0 = tn * BN;
  if (pr.lower && n0 > m0 + BM - 1) return;  // tile entirely above the diagonal
  // Phase stagger: the CTAs that share an SM start together and, all tiles being equal, stay in lock-step - they
  // reach the epilogue (no DMMA) at the same time and the tensor pipe idles.  Delaying the 2nd / 3rd resident CTA of
  // the first wave by a fraction of a tile time de-phases them once; their successors inherit the offset.
  if (pr.stagger > 0) {
    const unsigned wave = (blockIdx.x + gridDim.x * blockIdx.y) / (unsigned)pr.stagger_sms;
    if (wave > 0 && wave < (unsigned)pr.stagger_slots) {
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)wave * pr.stagger) {
      }
    }
  }

  // split-K range (multiples of BK)
  int kbeg = 0, kend = pr.k;
  if (splits > 1) {
    const int ktiles = (pr.k + BK - 1) / BK;
    const int per = (ktiles + splits - 1) / splits;
    kbeg = min(pr.k, (int)blockIdx.y * per * BK);
    kend = min(pr.k, kbeg + per * BK);
  }
  const int KT = (kend - kbeg + BK - 1) / BK;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % WM) * WTM, wn0 = (warp / WM) * WTN;

  double acc[MT][NTL][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // Read-modify-write updates (beta != 0) on 64 x 64 tiles: the C tile is fetched into registers NOW, so that its HBM
  // latency is hidden under the main loop instead of being exposed in the epilogue of every tile (the short-k
  // rank-128 / rank-256 updates of the dense->band stage spend a quarter of a tile's time there).
  constexpr bool kPrefetchC = PFC && (BM == 64 && BN == 64);
  constexpr int NPRE = kPrefetchC ? (BM / 2) * BN / NT : 1;
  double2 cpre[NPRE];
  const bool vecC0 = ((((uintptr_t)pr.C) & 15u) == 0) && (pr.ldc % 2 == 0);
  const bool pre_ok = kPrefetchC && pr.prefetch_c && !pr.Cin && splits == 1 && pr.beta != 0.0 && vecC0 && (m0 + BM <= pr.m) && (n0 + BN <= pr.n);
  if (pre_ok) {
#pragma unroll
    for (int it = 0; it < NPRE; ++it) {
      const int idx = threadIdx.x + it * NT;
      const int rp = idx % (BM / 2), cc = idx / (BM / 2);
      cpre[it] = *reinterpret_cast<const double2*>(pr.C + (m0 + 2 * rp) + (long long)(n0 + cc) * pr.ldc);
    }
  }

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kbeg + kt * BK;
    double* a = As + stage * A_STAGE;
    double* b = Bs + stage * B_STAGE;
    if (TA)
      load_k_major<BM, NT, VEC>(a, pr.A, pr.lda, m0, k0, pr.m, kend);
    else
      load_mn_major<BM, NT, VEC>(a, pr.A, pr.lda, m0, k0, pr.m, kend);
    if (TB)
      load_mn_major<BN, NT, VEC>(b, pr.B, pr.ldb, n0, k0, pr.n, kend);
    else
      load_k_major<BN, NT, VEC>(b, pr.B, pr.ldb, n0, k0, pr.n, kend);
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < KT) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const double* a = As + (kt % STAGES) * A_STAGE;
    const double* b = Bs + (kt % STAGES) * B_STAGE;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      double af[MT], bf[NTL];
      const int kk = ks * 4 + t;
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int row = wm0 + i * 8 + g;
        af[i] = TA ? a[row * KPAD + kk] : a[kk * SA + row];
      }
#pragma unroll
      for (int j = 0; j < NTL; ++j) {
        const int col = wn0 + j * 8 + g;
        bf[j] = TB ? b[kk * SB + col] : b[col * KPAD + kk];
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: the accumulators go through shared memory (the pipeline buffers are free now) so that
  // C is read and written with coalesced 16-byte row-contiguous accesses - the fragment layout itself
  // would touch C in 64-byte pieces, one dependent load/store per accumulator.
  __syncthreads();
  double* C = pr.C;
  long long ldc = pr.ldc;
  double alpha = pr.alpha, beta = pr.beta;
  if (splits > 1) {
    C = ws + (size_t)blockIdx.y * (size_t)pr.m * (size_t)pr.n;
    ldc = pr.m;
    alpha = 1.0;
    beta = 0.0;
  }
  constexpr int CW = (BN < 64) ? BN : 64;  // columns staged per pass
  constexpr int LDS = BM + 2;              // 2*LDS = 4 (mod 16): conflict-free fragment stores
  static_assert((size_t)CW * LDS <= (size_t)STAGES * (A_STAGE + B_STAGE), "staging tile must fit");
  double* Cs = smem;
  // source of the read-modify-write (the same matrix unless an out-of-place update was asked for)
  const double* Ci = (splits > 1 || !pr.Cin) ? C : pr.Cin;
  const long long ldci = (splits > 1 || !pr.Cin) ? ldc : pr.ldcin;
  const bool vecC = ((((uintptr_t)C) & 15u) == 0) && (ldc % 2 == 0) && ((((uintptr_t)Ci) & 15u) == 0) && (ldci % 2 == 0);
  const bool mirror = (pr.lower == 2) && (n0 + BN - 1 < m0);  // tile strictly below the diagonal
#pragma unroll
  for (int chunk = 0; chunk < BN / CW; ++chunk) {
    const int c_lo = chunk * CW;
    if (wn0 >= c_lo && wn0 < c_lo + CW) {
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            Cs[(wn0 - c_lo + j * 8 + 2 * t + e) * LDS + wm0 + i * 8 + g] = acc[i][j][e];
    }
    __syncthreads();
    int pre_it = 0;
    for (int idx = threadIdx.x; idx < (BM / 2) * CW; idx += NT, ++pre_it) {
      const int rp = idx % (BM / 2), cc = idx / (BM / 2);
      const int row = m0 + 2 * rp, col = n0 + c_lo + cc;
      if (row >= pr.m || col >= pr.n) continue;
      const double v0 = alpha * Cs[cc * LDS + 2 * rp], v1 = alpha * Cs[cc * LDS + 2 * rp + 1];
      double* cp = C + row + (long long)col * ldc;
      const double* cip = Ci + row + (long long)col * ldci;
      if (vecC && row + 1 < pr.m) {
        double2 o = make_double2(v0, v1);
        if (beta != 0.0) {
          const double2 old = pre_ok ? cpre[kPrefetchC ? pre_it : 0] : *reinterpret_cast<const double2*>(cip);
          o.x += beta * old.x;
          o.y += beta * old.y;
        }
        *reinterpret_cast<double2*>(cp) = o;
        if (mirror) {
          Cs[cc * LDS + 2 * rp] = o.x;
          Cs[cc * LDS + 2 * rp + 1] = o.y;
        }
      } else {
        cp[0] = (beta != 0.0) ? v0 + beta * cip[0] : v0;
        if (mirror) Cs[cc * LDS + 2 * rp] = cp[0];
        if (row + 1 < pr.m) {
          cp[1] = (beta != 0.0) ? v1 + beta * cip[1] : v1;
          if (mirror) Cs[cc * LDS + 2 * rp + 1] = cp[1];
        }
      }
    }
    if (mirror) {
      // transposed copy of the finished tile: C[col, row], contiguous along the tile's columns
      __syncthreads();
      // (16-byte mirror stores - two tile columns per thread - were measured slower: 27.2 against 28.0 TF at k = 256;
      //  the strided shared-memory reads they need conflict two-way)
      for (int idx = threadIdx.x; idx < BM * CW; idx += NT) {
        const int cc = idx % CW, rr = idx / CW;
        const int row = m0 + rr, col = n0 + c_lo + cc;
        if (row < pr.m && col < pr.n) C[col + (long long)row * ldc] = Cs[cc * LDS + rr];
      }
    }
    if (chunk + 1 < BN / CW) __syncthreads();
  }
}

__global__ void splitk_reduce_kernel(GemmProb pr, int splits, const double* __restrict__ ws) {
  const long long total = (long long)pr.m * pr.n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx % pr.m), col = (int)(idx / pr.m);
    double s = 0.0;
    for (int sp = 0; sp < splits; ++sp) s += ws[(size_t)sp * total + idx];
    double* c = pr.C + row + (long long)col * pr.ldc;
    double v = pr.alpha * s;
    if (pr.beta != 0.0) v += pr.beta * (pr.Cin ? pr.Cin[row + (long long)col * pr.ldcin] : *c);
    *c = v;
  }
}

template <int BM, int BN, int WM, int WN, int ST = kStages>
static size_t smem_bytes() {
  return (size_t)ST * (TileSize<BM>::max + TileSize<BN>::max) * sizeof(double);
}

template <int BM, int BN, int WM, int WN, bool TA, bool TB, bool VEC, int ST = kStages, bool PFC = false>
static int launch_one(bk_ctx* ctx, const GemmProb& single, const GemmProb* dprobs, int nprob,
                      int max_tiles, int splits, double* ws) {
  auto kern = dgemm_kernel<BM, BN, WM, WN, TA, TB, VEC, ST, PFC>;
  const size_t smem = smem_bytes<BM, BN, WM, WN, ST>();
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    BK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(max_tiles, splits, nprob);
  kern<<<grid, WM * WN * 32, smem, ctx->stream>>>(single, dprobs, splits, ws);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

template <int BM, int BN, int WM, int WN>
static int dispatch(bk_ctx* ctx, bool ta, bool tb, bool vec, const GemmProb& single,
                    const GemmProb* dprobs, int nprob, int max_tiles, int splits, double* ws) {
#define BK_GEMM_CASE(TA_, TB_, V_)                                                              \
  if (ta == TA_ && tb == TB_ && vec == V_)                                                      \
    return launch_one<BM, BN, WM, WN, TA_, TB_, V_>(ctx, single, dprobs, nprob, max_tiles, splits, \
                                                     ws);
  BK_GEMM_CASE(false, false, false)
  BK_GEMM_CASE(false, false, true)
  BK_GEMM_CASE(false, true, false)
  BK_GEMM_CASE(false, true, true)
  BK_GEMM_CASE(true, false, false)
  BK_GEMM_CASE(true, false, true)
  BK_GEMM_CASE(true, true, false)
  BK_GEMM_CASE(true, true, true)
#undef BK_GEMM_CASE
  return BK_ERR_ARG;
}

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

bool gemm_operands_vec_ok(const void* A, long long lda, const void* B, long long ldb) {
  static const bool novec = getenv("BK_GEMM_NOVEC") != nullptr;
  if (novec) return false;
  return aligned16(A) && aligned16(B) && (lda % 2 == 0) && (ldb % 2 == 0);
}

int gemm_batched(bk_ctx* ctx, bool ta, bool tb, const GemmProb* dprobs, int nprob, int max_m,
                 int max_n, bool vec) {
  if (nprob <= 0 || max_m <= 0 || max_n <= 0) return BK_OK;
  GemmProb none{};
  if (max_n <= 32) {
    const int tiles = (int)(ceil_div(max_m, 128) * ceil_div(max_n, 32));
    return dispatch<128, 32, 8, 1>(ctx, ta, tb, vec, none, dprobs, nprob, tiles, 1, nullptr);
  } else if (max_m <= 64 || max_n <= 64) {
    const int tiles = (int)(ceil_div(max_m, 64) * ceil_div(max_n, 64));
    return dispatch<64, 64, 2, 4>(ctx, ta, tb, vec, none, dprobs, nprob, tiles, 1, nullptr);
  }
  const int tiles = (int)(ceil_div(max_m, 128) * ceil_div(max_n, 64));
  return dispatch<128, 64, 4, 2>(ctx, ta, tb, vec, none, dprobs, nprob, tiles, 1, nullptr);
}

int gemm(bk_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A,
         long long lda, const double* B, long long ldb, double beta, double* C, long long ldc,
         int lower) {
  return gemm_oop(ctx, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, nullptr, 0, C, ldc, lower);
}

int gemm_oop(bk_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, long long lda,
             const double* B, long long ldb, double beta, const double* Cin, long long ldcin, double* C, long long ldc,
             int lower) {
  if (m <= 0 || n <= 0) return BK_OK;
  GemmProb p;
  p.Cin = (Cin && Cin != C) ? Cin : nullptr;
  p.ldcin = ldcin;
  p.A = A;
  p.B = B;
  p.C = C;
  p.m = m;
  p.n = n;
  p.k = k;
  p.lda = lda;
  p.ldb = ldb;
  p.ldc = ldc;
  p.alpha = alpha;
  p.beta = beta;
  p.lower = lower;
  const bool vec = gemm_operands_vec_ok(A, lda, B, ldb);
  p.stagger = 0;
  static const bool no_prefetch = getenv("BK_GEMM_NOPREFETCH") != nullptr;
  p.prefetch_c = no_prefetch ? 0 : 1;
  p.stagger_sms = ctx->sm_count;
  p.stagger_slots = 0;

  int bm, bn;
  if (n <= 24 && n > 16) {
    // the K-pass of the marginal effects (n = 2P + 2 = 22 at P = 10): three 8-wide DMMA tiles instead of four, a
    // quarter of the tensor work of the 32-wide tile was padding
    bm = 128;
    bn = 24;
  } else if (n <= 32) {
    bm = 128;
    bn = 32;
  } else if (m <= 64) {
    bm = 64;
    bn = 64;
  } else if (n <= 64) {
    bm = 128;  // tall-and-64-wide (block-reflector products of the two-stage reduction)
    bn = 64;
  } else {
    // Measured on B200 (tools/gemm_shape_bench.py): residency beats tile size for FP64 DMMA.  128x64 tiles run
    // two CTAs per SM (32.8 TF at 8192^3 against 29.1 TF for 128x128, one CTA per SM); for short k, where the
    // prologue and the C traffic of the epilogue weigh most, 64x64 tiles with three CTAs per SM are best
    // (rank-128 read-modify-write update: 22.8 TF against 18.9 / 13.2).
    bm = 128;
    bn = 64;
    if (k <= 256) bm = 64;
  }
  if (const char* f = getenv("BK_GEMM_FORCE")) {  // tuning experiments only
    int fm = 0, fn = 0;
    if (sscanf(f, "%dx%d", &fm, &fn) == 2 && ((fm == 128 && (fn == 128 || fn == 64 || fn == 32)) || (fm == 64 && fn == 64)) &&
        !(fn < n && fn <= 32)) {
      bm = fm;
      bn = fn;
    }
  }
  const int tiles = (int)(ceil_div(m, bm) * ceil_div(n, bn));
  // deterministic split-K when the tile grid cannot fill the machine and k is long.  The number of splits is
  // chosen against wave quantisation: tiles * splits CTAs run in waves of (resident CTAs per SM) * SMs, and a
  // last wave that is 10 % full costs as much as a full one (e.g. 157 tiles x 4 splits on 296 slots = 2.12 waves).
  int splits = 1;
  const int resident = (bm == 128 && bn == 128) ? 1 : (bm == 64 ? 3 : 2);  // (the narrow 128 x 24 / 32 tiles: >= 2)
  const int slots = resident * ctx->sm_count;
  if (!lower && tiles < 8 * slots && k >= 1024) {
    const int smax = (int)std::max<int64_t>(1, std::min<int64_t>(64, k / 256));
    if ((long long)tiles * smax <= slots) {
      splits = smax;  // cannot even fill one wave: as many splits as the k extent allows
    } else {
      double best = -1.0;
      for (int sct = 1; sct <= smax; ++sct) {
        const double waves = (double)tiles * sct / slots;
        const double score = waves / std::ceil(waves) - 0.004 * sct;  // small penalty: partials are extra traffic
        if (score > best) {
          best = score;
          splits = sct;
        }
      }
    }
  }
  double* ws = nullptr;
  if (splits > 1) {
    BK_TRY(ctx->gemm_ws.ensure((size_t)splits * (size_t)m * (size_t)n));
    ws = ctx->gemm_ws.p;
  }
  {
    // start stagger of the first wave (see the kernel): only for grids of many equal tiles
    static const int stagger_cycles = getenv("BK_GEMM_STAGGER") ? atoi(getenv("BK_GEMM_STAGGER")) : 0;
    if (stagger_cycles > 0 && (long long)tiles * splits > 4LL * slots) {
      p.stagger = stagger_cycles;
      p.stagger_slots = resident;
    }
  }
  int rc;
  // Short read-modify-write updates (k <= 128, the rank-128 updates of the distributed / un-paired dense->band stage):
  // the variant that fetches the C tile into registers before the main loop (24.1 against 22.8 TF at k = 128; at
  // k = 256 the extra 16 registers cost more than the hidden latency gains: 27.4 against 28.0 TF).
  // (Also measured and dropped: a 2-stage pipeline with 4 CTAs per SM, 26.6 against 28.0 TF; a start stagger of the
  // first wave, no effect - the CTAs of an SM are not phase-locked.)
  if (bm == 64 && bn == 64 && !ta && tb && vec && splits == 1 && beta != 0.0 && k <= 128 && p.prefetch_c)
    rc = launch_one<64, 64, 2, 4, false, true, true, 3, true>(ctx, p, nullptr, 1, tiles, splits, ws);
  else if (bn == 24)
    rc = dispatch<128, 24, 8, 1>(ctx, ta, tb, vec, p, nullptr, 1, tiles, splits, ws);
  else if (bn == 32)
    rc = dispatch<128, 32, 8, 1>(ctx, ta, tb, vec, p, nullptr, 1, tiles, splits, ws);
  else if (bm == 128 && bn == 64)
    rc = dispatch<128, 64, 4, 2>(ctx, ta, tb, vec, p, nullptr, 1, tiles, splits, ws);
  else if (bm == 64)
    rc = dispatch<64, 64, 2, 4>(ctx, ta, tb, vec, p, nullptr, 1, tiles, splits, ws);
  else
    rc = dispatch<128, 128, 2, 4>(ctx, ta, tb, vec, p, nullptr, 1, tiles, splits, ws);
  BK_TRY(rc);
  if (splits > 1) {
    // (The reduction fused into the GEMM - the CTA that finishes a tile last adds the partials - was measured and
    //  dropped: one CTA reading splits x 64 KB serialises what this kernel spreads over the machine; the m x 64 x m
    //  products of the dense->band stage went from 0.39 to 0.42 s.)
    const long long total = (long long)m * n;
    const int blocks = (int)std::min<long long>(ceil_div(total, 256), 4 * ctx->sm_count);
    splitk_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(p, splits, ws);
    BK_LAUNCHED(ctx);
    BK_CUDA(cudaGetLastError());
  }
  return BK_OK;
}

}  // namespace bk
