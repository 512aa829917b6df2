// Fused, device-resident bigKRLS fit: the five stages of reference R/bigKRLS.R:262-329 behind
// one C call, with X, K, Q, Lambda, c never leaving HBM between stages.
//
//   1/5 kernel            gauss_kernel.cu                       (src/gauss_kernel.cpp)
//   2/5 eigen             sy2sb.cu + sb2st.cu (two-stage) or sytrd.cu (one-stage), stedc.cu, ormtr.cu;
//                         eigen_topk.cu for Neig << N           (src/eigen.cpp, bEigen R:173-199)
//   3/5 lambda            golden section on the host, bit-for-bit the arithmetic of
//                         R/bigKRLS_Rcpp_functions.R:5-82, with the LOO loss of up to 15
//                         speculative candidates per pass over Q (loo.cu)
//   4/5 coefficients      c (loo.cu), yhat = K c, sigma^2, vcov(c) = s2 Q (L+lam)^-2 Q',
//                         vcov(yhat) = s2 Q (L/(L+lam))^2 Q'    (R/bigKRLS.R:286-307)
//   5/5 marginal effects  one tall-skinny pass K [1 c X X.c], epilogue, spectral variances
//                         (src/bigderiv_v3.cpp)
//
// Multi-GPU (one process per GPU): N x N matrices are partitioned by COLUMN BLOCKS - for the
// symmetric K, V this is the transposed row panel, and it is contiguous in column-major
// storage so blocks can be all-gathered in place.  Native path (peer-memory communicator, peer.cu): the dense->band
// stage of the eigensolver and the back-transformation are distributed, the band->tridiagonal stage and the divide &
// conquer run on every rank (same bits, nothing to broadcast); small problems and the generic bk_comm callback
// transport (NCCL / gloo via torch.distributed in the Python host) run the eigensolver on rank 0 and broadcast Q.
#include <algorithm>
#include <cmath>
#include <functional>
#include <map>
#include <cstring>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"
#include "peer.cuh"

using namespace bk;

struct bk_fit {
  bk_ctx* ctx = nullptr;
  int n = 0, p = 0, neig = 0, k = 0, pd = 0;
  int rank = 0, world = 1;
  int c0 = 0, c1 = 0;  // owned column block of the N x N fields
  bk_fit_opts opts;
  std::vector<int> which;      // derivative columns (0-based)
  std::vector<double> evals;   // all neig eigenvalues, descending (host)
  double lambda = 0, Le = 0, sigmasq = 0, neffective = 0;
  int n_probes = 0, n_passes = 0;
  bk_fit_info info;
  DevBuf<double> X, y, K, Q, ev, w, c, yhat, sig2, Vc, Vf, D, var, binfo;
  bool have_vcov = false, have_vf = false, have_deriv = false;
  const double* K_host_done = nullptr;  // host buffer that already holds K (early D2H under the eigensolver)
  bk::CopyTicket k_copy;
  bool k_block_only = false;  // K holds only the own column block (native multi-GPU fits), not the whole matrix
  const double* kblock() const { return k_block_only ? K.p : K.p + (long long)c0 * n; }
  ~bk_fit() { bk::copier_wait(&k_copy); }  // K must outlive the queued copy
};

namespace {

#define COMM_CALL(expr, what)                                   \
  do {                                                          \
    BK_CUDA(cudaStreamSynchronize(ctx->stream));                \
    if ((expr) != 0) {                                          \
      set_error("communicator callback failed: %s", what);      \
      return BK_ERR_COMM;                                       \
    }                                                           \
  } while (0)

// ---- lambda bounds: R/bigKRLS_Rcpp_functions.R:16-36 (all Neig eigenvalues; R's sum()
// accumulates in long double) -------------------------------------------------------------------
long double ratio_sum(const std::vector<double>& ev, double x) {
  long double s = 0.0L;
  for (double e : ev) s += (long double)(e / (e + x));
  return s;
}
// The same sum in plain double with eight independent accumulators (vectorisable; ~15 us at Neig = 20 000
// against ~60 us for the dependent long-double chain).  Its error is bounded by (Neig/8) eps sum|terms|.
double ratio_sum_fast(const std::vector<double>& ev, double x, double* abs_terms) {
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const size_t n = ev.size(), n8 = n & ~(size_t)7;
  for (size_t i = 0; i < n8; i += 8)
    for (int j = 0; j < 8; ++j) {
      const double t = ev[i + j] / (ev[i + j] + x);
      s[j] += t;
      a[j] += std::fabs(t);
    }
  for (size_t i = n8; i < n; ++i) {
    const double t = ev[i] / (ev[i] + x);
    s[0] += t;
    a[0] += std::fabs(t);
  }
  *abs_terms = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}
// sign of (R's sum(ev/(ev+x)) - thr): -1 below, 0 equal, +1 above.  Decided by the fast sum whenever it is
// further from the threshold than its own error bound (plus the long-double sum's), otherwise by the reference's
// arithmetic itself - so every comparison of the bounds loops has exactly the outcome R would get.
// Most of a kernel matrix's spectrum is rounding noise around zero: for |e| <= 1e-7 x the term e/(e+x) equals e/x up
// to a relative 2e-7 (of a term that is itself below 1e-7), so the tail of the (descending) spectrum enters the fast sums through suffix sums of e and |e|
// and only the leading part is divided out - a few hundred terms instead of Neig.  The tail's truncation error joins
// the error band of the comparison, so the decisions stay exactly R's.
static int g_cmp_calls = 0, g_cmp_exact = 0;  // BK_LAMBDA_PROF
struct RatioTail {
  const std::vector<double>* ev = nullptr;
  bool sorted = false;
  std::vector<double> s1, a1;  // suffix sums of e and |e|
  void build(const std::vector<double>& e) {
    ev = &e;
    const size_t n = e.size();
    sorted = true;
    for (size_t i = 1; i < n && sorted; ++i) sorted = e[i] <= e[i - 1];
    if (!sorted) return;
    s1.assign(n + 1, 0.0);
    a1.assign(n + 1, 0.0);
    for (size_t i = n; i-- > 0;) {
      s1[i] = s1[i + 1] + e[i];
      a1[i] = a1[i + 1] + std::fabs(e[i]);
    }
  }
  // first index whose value (and, the spectrum being descending, every later POSITIVE value) is below 1e-7 x; the
  // negative noise behind it must be as small in magnitude for the expansion to hold: checked on the last element
  size_t cut(double x) const {
    const std::vector<double>& e = *ev;
    if (!sorted || !(x > 0.0) || e.empty() || std::fabs(e.back()) > 1e-7 * x) return e.size();
    size_t lo = 0, hi = e.size();
    while (lo < hi) {
      const size_t mid = lo + (hi - lo) / 2;
      if (e[mid] < 1e-7 * x) hi = mid; else lo = mid + 1;
    }
    return lo;
  }
};
thread_local RatioTail g_tail;

int ratio_cmp(const std::vector<double>& ev, double x, double thr) {
  if (g_tail.ev == &ev && g_tail.sorted) {
    const size_t c = g_tail.cut(x);
    if (c < ev.size()) {
      double s[4] = {0, 0, 0, 0}, a[4] = {0, 0, 0, 0};
      for (size_t i = 0; i < c; ++i) {
        const double t = ev[i] / (ev[i] + x);
        s[i & 3] += t;
        a[i & 3] += std::fabs(t);
      }
      const double tail = g_tail.s1[c] / x, tail_abs = g_tail.a1[c] / x;
      const double f = ((s[0] + s[1]) + (s[2] + s[3])) + tail;
      const double at = ((a[0] + a[1]) + (a[2] + a[3])) + tail_abs;
      const double band = 4.0 * 2.220446049250313e-16 * (double)(ev.size() / 4 + 8) * at + 4e-7 * tail_abs +
                          1e-13 * std::fabs(thr);
      ++g_cmp_calls;
      if (std::isfinite(f) && std::fabs(f - thr) > band) return f < thr ? -1 : 1;
      ++g_cmp_exact;
      const double sx = (double)ratio_sum(ev, x);
      return sx < thr ? -1 : (sx > thr ? 1 : 0);
    }
  }
  double at = 0.0;
  const double f = ratio_sum_fast(ev, x, &at);
  const double band = 4.0 * 2.220446049250313e-16 * (double)(ev.size() / 8 + 8) * at + 1e-13 * std::fabs(thr);
  if (std::isfinite(f) && std::fabs(f - thr) > band) return f < thr ? -1 : 1;
  const double s = (double)ratio_sum(ev, x);
  return s < thr ? -1 : (s > thr ? 1 : 0);
}

// Approximate real root of sum(ev/(ev+x)) = thr on x > 0 (the function is convex and decreasing there) by a few
// Newton steps in plain double - only a STARTING GUESS for the exact integer searches below, which verify it with
// the reference's own comparisons.
double ratio_root_guess(const std::vector<double>& ev, double thr, double x0) {
  double x = x0;
  const bool tail_ok = g_tail.ev == &ev && g_tail.sorted;
  for (int it = 0; it < 12; ++it) {
    double f = 0.0, g = 0.0;
    const size_t c = tail_ok ? g_tail.cut(x) : ev.size();
    {
      // four independent accumulator pairs: the divisions pipeline instead of waiting for one dependent add chain
      double fa[4] = {0, 0, 0, 0}, ga[4] = {0, 0, 0, 0};
      size_t i = 0;
      for (; i + 4 <= c; i += 4)
        for (int u = 0; u < 4; ++u) {
          const double e = ev[i + u], d = e + x, r = e / d;
          fa[u] += r;
          ga[u] -= r / d;
        }
      for (; i < c; ++i) {
        const double e = ev[i], d = e + x, r = e / d;
        fa[0] += r;
        ga[0] -= r / d;
      }
      f = (fa[0] + fa[1]) + (fa[2] + fa[3]);
      g = (ga[0] + ga[1]) + (ga[2] + ga[3]);
    }
    if (c < ev.size()) {  // the noise tail: e/(e+x) ~ e/x (a starting guess only)
      f += g_tail.s1[c] / x;
      g -= g_tail.s1[c] / (x * x);
    }
    if (!(g < 0.0) || !std::isfinite(f)) break;
    const double xn = x - (f - thr) / g;
    const double nx = (xn > 0.0) ? xn : 0.5 * x;  // Newton from the left of a convex decreasing function stays left
    if (std::fabs(nx - x) <= 5e-3) {  // the candidates are 1 (U) and 0.05 (L) apart: a guess this close is enough
      x = nx;
      break;
    }
    x = nx;
  }
  return x;
}

int lambda_bounds(const std::vector<double>& ev, int n, double* L, double* U) {
  const bool bprof = getenv("BK_LAMBDA_PROF") != nullptr;
  auto now_us = []() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
  };
  const double tb0 = bprof ? now_us() : 0.0;
  g_cmp_calls = g_cmp_exact = 0;
  struct ProfEnd {
    bool on;
    double t0;
    std::function<double()> now;
    ~ProfEnd() {
      if (on) fprintf(stderr, "[lambda bounds prof] %.0f us, %d comparisons of which %d by the exact long-double sum\n", now() - t0, g_cmp_calls, g_cmp_exact);
    }
  } prof_end{bprof, tb0, now_us};
  g_tail.build(ev);
  struct TailReset {
    ~TailReset() { g_tail.ev = nullptr; }
  } tail_reset;
  if (!(*U > 0.0)) {
    // Reference (R/bigKRLS_Rcpp_functions.R:16-20): U = n; while (sum(ev/(ev+U)) < 1) U = U - 1 - hundreds to
    // thousands of O(Neig) sums.  f(U) = sum(ev/(ev+U)) is strictly decreasing with f(U) - f(U+1) ~ 1/U, many
    // orders above the rounding of the sum, so the first integer step m with f(n-m) >= 1 is found by bisection
    // and is exactly where the scan stops (f(n-m) >= 1 and f(n-m+1) < 1).
    auto ok = [&](long long m) { return ratio_cmp(ev, (double)n - (double)m, 1.0) >= 0; };
    long long lo = 0, hi = -1;
    if (ok(0)) {
      hi = 0;
    } else if ([&]() {
                 // Newton guess for the crossing, then the scan's own stopping rule on the two neighbouring integers
                 // sum e/(e+U) = 1 with U far above most of the spectrum: U ~ sum e - sum e^2 / U ~ tr - sum e^2 / tr
                 double tr = 0.0, tr2 = 0.0;
                 for (double e : ev) {
                   tr += e;
                   tr2 += e * e;
                 }
                 double x0 = (tr > 0.0) ? tr - tr2 / tr : 0.5 * (double)n;
                 if (!(x0 > 1.0) || !(x0 < (double)n)) x0 = 0.5 * (double)n;
                 const double xr = ratio_root_guess(ev, 1.0, x0);
                 if (!(xr > 1.0) || !(xr < (double)n)) return false;
                 for (long long m = (long long)n - (long long)std::floor(xr) - 1, t = 0; t < 3; ++m, ++t) {
                   if (m < 1 || m > (long long)n - 1) continue;
                   if (ok(m) && !ok(m - 1)) {
                     hi = m;
                     return true;
                   }
                 }
                 return false;
               }()) {
      // hi = first passing offset, verified: hi - 1 fails
    } else {
      long long step = 1;
      while (hi < 0) {  // gallop: lo is always a failing offset
        const long long m = std::min<long long>(lo + step, (long long)n - 1);
        if (ok(m)) {
          hi = m;
        } else {
          lo = m;
          if (m == (long long)n - 1) {
            set_error("lambda search: upper bound loop did not terminate (degenerate spectrum)");
            return BK_ERR_NUMERIC;
          }
          step *= 2;
        }
      }
      while (hi - lo > 1) {
        const long long mid = lo + (hi - lo) / 2;
        if (ok(mid)) hi = mid; else lo = mid;
      }
      // lo fails, hi = lo + 1 passes: exactly where the linear scan stops
    }
    *U = (double)n - (double)hi;
  }
  if (!(*L >= 0.0)) {  // negative or NaN = not supplied (a user L = 0 is legal, R/bigKRLS.R:225-228)
    // Reference (:26-36): L = eps; q = which.min(abs(ev - max(ev)/1000)); while (sum(ev/(ev+L)) > q) L = L + 0.05.
    // The candidates are the partial sums l_0 = eps, l_i = fl(l_{i-1} + 0.05) exactly as the scan forms them;
    // sum(ev/(ev+l)) decreases along them, so the first index that fails the loop test is found by gallop + bisection.
    const double emax = *std::max_element(ev.begin(), ev.end());
    int q = 0;
    double best = INFINITY;
    for (size_t i = 0; i < ev.size(); ++i) {
      const double v = std::fabs(ev[i] - emax / 1000.0);
      if (v < best) {
        best = v;
        q = (int)i + 1;  // which.min: 1-based, first minimum
      }
    }
    std::vector<double> ls{2.220446049250313e-16};  // .Machine$double.eps (:28)
    auto cand = [&](size_t i) {
      while (ls.size() <= i) ls.push_back(ls.back() + 0.05);
      return ls[i];
    };
    auto stop = [&](size_t i) { return ratio_cmp(ev, cand(i), (double)q) <= 0; };  // loop test false
    size_t lo = 0, hi = 0;
    bool found = stop(0);
    if (!found) {
      // about q eigenvalues lie above the root: start at the q-th largest (the spectrum is sorted when the tail is)
      const double xs = (g_tail.sorted && q >= 1 && (size_t)q <= ev.size() && ev[q - 1] > 0.0) ? ev[q - 1] : 1.0;
      const double xr = ratio_root_guess(ev, (double)q, xs);
      if (xr > 0.0 && xr < 4e6) {
        const long long i0 = (long long)std::ceil((xr - ls[0]) / 0.05);
        for (long long i = std::max(1LL, i0 - 1), t = 0; t < 3 && !found; ++i, ++t)
          if (stop((size_t)i) && !stop((size_t)i - 1)) {
            hi = (size_t)i;
            found = true;
          }
      }
    }
    if (!found) {
      size_t step = 1;
      for (;;) {
        const size_t m = lo + step;
        if (m > 100000000u) {
          set_error("lambda search: lower bound loop did not terminate");
          return BK_ERR_NUMERIC;
        }
        if (stop(m)) {
          hi = m;
          break;
        }
        lo = m;
        step *= 2;
      }
      while (hi - lo > 1) {
        const size_t mid = lo + (hi - lo) / 2;
        if (stop(mid)) hi = mid; else lo = mid;
      }
    }
    *L = cand(hi);
  }
  return BK_OK;
}

struct GState {
  double L, U, X1, X2;
};
const double GOLD = 0.381966;

// one golden-section step given the outcome of `S1 < S2` (R/bigKRLS_Rcpp_functions.R:56-69);
// returns the newly required evaluation point.
double gs_step(GState& s, bool s1_less) {
  if (s1_less) {
    s.U = s.X2;
    s.X2 = s.X1;
    s.X1 = s.L + GOLD * (s.U - s.L);
    return s.X1;
  }
  s.L = s.X1;
  s.X1 = s.X2;
  s.X2 = s.U - GOLD * (s.U - s.L);
  return s.X2;
}

struct LooEvaluator {
  // evaluates Le for a batch of candidate lambdas (<= 16) into out[]
  std::function<int(const std::vector<double>&, double*)> eval_batch;
  int batch = 7;
  std::map<uint64_t, double> cache;
  int passes = 0;

  static uint64_t key(double v) {
    uint64_t u;
    memcpy(&u, &v, 8);
    return u;
  }
  bool has(double lam) const { return cache.count(key(lam)) != 0; }

  int run(const std::vector<double>& lams) {
    double host[16];
    BK_TRY(eval_batch(lams, host));
    for (size_t i = 0; i < lams.size(); ++i) cache[key(lams[i])] = host[i];
    ++passes;
    return BK_OK;
  }

  // Le(lam); on a miss evaluates lam together with the speculative continuation of the search
  // tree rooted at `st` (breadth first, so the shallow levels are kept when the batch is full).
  // Candidate points are produced by the same floating-point expressions the sequential search
  // would use, so a later lookup hits the cache bit for bit.
  int get(double lam, const GState& st, const std::vector<double>& also, double* out) {
    if (!has(lam)) {
      std::vector<double> lams;
      auto push = [&](double v) {
        if ((int)lams.size() >= batch || has(v)) return;
        for (double x : lams)
          if (key(x) == key(v)) return;
        lams.push_back(v);
      };
      push(lam);
      for (double v : also) push(v);
      std::vector<GState> frontier{st};
      while ((int)lams.size() < batch && !frontier.empty() && frontier.size() < 64) {
        std::vector<GState> next;
        for (const GState& s : frontier) {
          for (int o = 0; o < 2; ++o) {
            GState c = s;
            push(gs_step(c, o == 0));
            next.push_back(c);
          }
        }
        frontier.swap(next);
      }
      BK_TRY(run(lams));
    }
    *out = cache[key(lam)];
    return BK_OK;
  }
};

int lambda_search(LooEvaluator& ev, double L, double U, double tol, double* lam_out, int* probes) {
  GState s;
  s.L = L;
  s.U = U;
  s.X1 = L + GOLD * (U - L);  // :38
  s.X2 = U - GOLD * (U - L);  // :39
  double S1, S2;
  BK_TRY(ev.get(s.X1, s, {s.X2}, &S1));
  BK_TRY(ev.get(s.X2, s, {}, &S2));
  int np = 2;
  auto finite = [&]() {
    if (std::isfinite(S1) && std::isfinite(S2)) return true;
    // R: `while (abs(S1 - S2) > tol)` stops with "missing value where TRUE/FALSE needed"
    set_error("lambda search: leave-one-out loss is not finite (NaN in y, or a zero diagonal of the inverse)");
    return false;
  };
  if (!finite()) return BK_ERR_NUMERIC;
  while (std::fabs(S1 - S2) > tol) {  // :54
    if (S1 < S2) {
      const double x = gs_step(s, true);
      S2 = S1;
      BK_TRY(ev.get(x, s, {}, &S1));
    } else {
      const double x = gs_step(s, false);
      S1 = S2;
      BK_TRY(ev.get(x, s, {}, &S2));
    }
    if (!finite()) return BK_ERR_NUMERIC;
    if (++np > 10000) {
      set_error("lambda search: no convergence after 10000 probes");
      return BK_ERR_NUMERIC;
    }
  }
  *lam_out = (S1 < S2) ? s.X1 : s.X2;  // :71
  *probes = np;
  return BK_OK;
}

// sizes of the per-rank segments of an n-long (times `width`) field partitioned like the column blocks
void block_partition(int n, int world, long long width, std::vector<long long>& counts, std::vector<long long>& displs) {
  counts.resize(world);
  displs.resize(world);
  for (int r = 0; r < world; ++r) {
    const long long a = (long long)n * r / world, b = (long long)n * (r + 1) / world;
    counts[r] = (b - a) * width;
    displs[r] = a * width;
  }
}

int run_fit(bk_fit* f, const bk_comm* comm) {
  bk_ctx* ctx = f->ctx;
  const int n = f->n, p = f->p;
  const bk_fit_opts& o = f->opts;
  const long long ld = n;
  const bool multi = comm && comm->world > 1;
  // native: every exchange is a kernel of this library storing into the peers' HBM (peer.cu); otherwise the
  // host-provided callbacks (generic path: gloo on CPU boxes, any other transport)
  bk_peer* peer = (multi && comm->peer) ? comm->peer : nullptr;
  const bool native = peer != nullptr;
  const int world = multi ? comm->world : 1, rank = multi ? comm->rank : 0;
  Timer tm, total;
  BK_TRY(tm.init(ctx->stream));
  BK_TRY(total.init(ctx->stream));
  memset(&f->info, 0, sizeof(f->info));
  const uint64_t launches0 = ctx->n_launches;
  const int nloc = f->c1 - f->c0;
  const int pd_all = o.derivative ? f->pd : 0;
  const bool topk = use_topk(n, f->neig);
  const bool dist_eig = native && !topk && use_twostage(n, f->neig, o.eigtrunc);

  size_t off_pack = 0, off_Q = 0, off_yhat = 0, off_D = 0;
  if (native) {
    BK_REQUIRE(peer->world == world && peer->rank == rank, "bk_fit_run: bk_comm and its peer communicator disagree");
    // symmetric heap for this fit (collective; grows once, then reused): eigenvalue pack, Q for the broadcast from
    // rank 0, gathered yhat and derivatives, and the buffers of the distributed eigensolver stages
    size_t need = sizeof(double) * ((size_t)f->neig + 16 + (size_t)n * (topk ? 0 : f->neig) + (size_t)n +
                                    (size_t)n * std::max(1, pd_all)) + 8192;
    need += dist_eig ? sy2sb_dist_heap_bytes(n) : 0;
    need += topk ? eigen_topk_heap_bytes(n, f->neig) : 0;
    BK_TRY(peer_ensure_heap(peer, need));
    BK_TRY(peer_alloc(peer, sizeof(double) * ((size_t)f->neig + 16), &off_pack));
    if (!topk) BK_TRY(peer_alloc(peer, sizeof(double) * (size_t)n * f->neig, &off_Q));
    BK_TRY(peer_alloc(peer, sizeof(double) * (size_t)n, &off_yhat));
    BK_TRY(peer_alloc(peer, sizeof(double) * (size_t)n * std::max(1, pd_all), &off_D));
    // the previous fit is over on every rank before anybody stores into these buffers again
    BK_TRY(peer_barrier(peer, ctx->stream));
  }
  total.start();

  // ---- 1/5 kernel -------------------------------------------------------------------------
  tm.start();
  if (!multi) {
    BK_TRY(f->K.alloc((size_t)n * n));
    BK_TRY(gauss_kernel_sym(ctx, f->X.p, ld, n, p, o.sigma, f->K.p, ld));
  } else if (native) {
    // own column block only: nobody needs the whole kernel matrix (the eigensolver stages build their own columns
    // straight from X, the K-pass and the D2H of K use the own block) - no exchange at all
    BK_TRY(f->K.alloc((size_t)n * std::max(1, nloc)));
    f->k_block_only = true;
    BK_TRY(gauss_kernel_rect(ctx, f->X.p, ld, n, f->X.p + f->c0, ld, nloc, p, o.sigma, f->K.p, ld));
  } else {
    // own column block K[:, c0:c1] = kernel(X, X[c0:c1, :]) then in-place all-gather
    BK_TRY(f->K.alloc((size_t)n * n));
    BK_TRY(gauss_kernel_rect(ctx, f->X.p, ld, n, f->X.p + f->c0, ld, nloc, p, o.sigma,
                             f->K.p + (long long)f->c0 * ld, ld));
    std::vector<int64_t> counts(world), displs(world);
    for (int r = 0; r < world; ++r) {
      const int64_t a = (int64_t)n * r / world, b = (int64_t)n * (r + 1) / world;
      counts[r] = (b - a) * n;
      displs[r] = a * n;
    }
    COMM_CALL(comm->allgatherv(comm->user, f->K.p, counts.data(), displs.data()), "allgatherv(K)");
  }
  const double* Kblk = f->kblock();  // K[:, c0:c1], ld n
  f->info.t_kernel = tm.stop();
  if (o.K_host) {
    // K is final: send this rank's column block to the host now, under the eigensolver (pinned destination: one
    // DMA on the copy stream; pageable big.matrix memory: the copy engine's bounce lanes, hostcopy.cu)
    BK_TRY(copier_submit(ctx, o.K_host, Kblk, sizeof(double) * (size_t)n * nloc, ctx->stream, &f->k_copy));
  }

  // ---- 2/5 eigen ----------------------------------------------------------------------------
  tm.start();
  f->evals.assign(f->neig, 0.0);
  int k = 0;
  {
    DevBuf<double> Zfull;
    BK_TRY(Zfull.alloc((size_t)n * f->neig));
    std::vector<double> ev(n);
    EigenTimes et;
    int eig_rc = BK_OK;
    bool replicated = false;  // every rank already holds the result (no broadcast)
    if (native && topk) {
      // Neig << N: block Krylov with K partitioned over the ranks (one all-gather of n x b per K X)
      TopkStats ts;
      BK_TRY(eigen_topk(ctx, Kblk, ld, n, f->neig, ev.data(), Zfull.p, ld, &ts, peer, f->c0, nloc));
      for (int i = 0; i < f->neig; ++i)
        if (ev[i] >= o.eigtrunc * ev[0]) k = i + 1;
      f->info.krylov_matvecs = ts.matvecs;
      f->info.krylov_restarts = ts.restarts;
      replicated = true;
    } else if (dist_eig) {
      // dense -> band distributed over all ranks; band -> tridiagonal and D&C on every rank (same bits); the
      // back-transformation split by columns and all-gathered - every rank returns with the complete result
      BK_TRY(eigen_full_dist(ctx, peer, f->X.p, ld, p, o.sigma, n, ev.data(), f->neig, o.eigtrunc, &k, Zfull.p, ld, off_Q,
                             &et));
      replicated = true;
    } else if (!multi || rank == 0) {
      DevBuf<double> Kfull;
      const double* Kmat = f->K.p;
      if (native) {
        // single-GPU eigensolver path (small n, or all eigenvectors of a large matrix): rank 0 builds the whole
        // kernel matrix itself - cheaper than gathering it
        BK_TRY(Kfull.alloc((size_t)n * n));
        BK_TRY(gauss_kernel_sym(ctx, f->X.p, ld, n, p, o.sigma, Kfull.p, ld));
        Kmat = Kfull.p;
      }
      if (topk) {
        // Neig << N (reference: sp_mat + eigs_sym, src/eigen.cpp:18-22): restarted block Krylov
        TopkStats ts;
        eig_rc = eigen_topk(ctx, Kmat, ld, n, f->neig, ev.data(), Zfull.p, ld, &ts);
        k = 0;  // lastkeeper over the Neig values (R/bigKRLS_Rcpp_functions.R:190)
        for (int i = 0; i < f->neig && eig_rc == BK_OK; ++i)
          if (ev[i] >= o.eigtrunc * ev[0]) k = i + 1;
        f->info.krylov_matvecs = ts.matvecs;
        f->info.krylov_restarts = ts.restarts;
      } else {
        eig_rc = eigen_full(ctx, Kmat, ld, n, ev.data(), f->neig, o.eigtrunc, &k, Zfull.p, ld, &et);
      }
      if (!multi) BK_TRY(eig_rc);
      if (eig_rc == BK_ERR_CUDA) return eig_rc;
    }
    for (int i = 0; i < f->neig; ++i) f->evals[i] = ev[i];
    f->info.t_tridiag = et.tridiag;
    f->info.t_dc = et.dc;
    f->info.t_backtransform = et.backtransform;
    f->info.sytrd_launches = et.sytrd.launches;
    f->info.sytrd_kernel_seconds = et.sytrd.kernel_seconds;
    f->info.sytrd_bytes = et.sytrd.algorithmic_bytes;
    f->info.twostage = et.twostage;
    f->info.t_sy2sb = et.t_sy2sb;
    f->info.t_sb2st = et.t_sb2st;
    f->info.t_q2 = et.t_q2;
    f->info.t_q1 = et.t_q1;
    f->info.band_gemm_launches = et.band.gemm_launches;
    f->info.band_gemm_seconds = et.band.gemm_seconds;
    f->info.band_gemm_flops = et.band.gemm_flops;
    f->info.dc_levels = et.dc_stats.levels;
    f->info.dc_merge_flops = (double)et.dc_stats.merge_flops;
    f->info.dc_top_n = et.dc_stats.top_n;
    f->info.dc_top_k = et.dc_stats.top_k;
    BK_TRY(f->ev.alloc(f->neig + 1));
    if (multi && !replicated) {
      // broadcast [status | k, evals] then Q[:, :k].  Rank 0's status travels with the data: a numerical failure of
      // the eigensolver there must not leave the other ranks waiting in a collective (they all return the error).
      std::vector<double> pack(f->neig + 1);
      pack[0] = (eig_rc != BK_OK) ? (double)eig_rc : (double)k;
      for (int i = 0; i < f->neig; ++i) pack[1 + i] = f->evals[i];
      double* dpack = native ? peer_ptr(peer, off_pack) : f->ev.p;
      if (rank == 0)
        BK_CUDA(cudaMemcpyAsync(dpack, pack.data(), sizeof(double) * (f->neig + 1), cudaMemcpyHostToDevice, ctx->stream));
      if (native)
        BK_TRY(peer_broadcast_sym(peer, off_pack, f->neig + 1, 0, ctx->stream));
      else
        COMM_CALL(comm->broadcast(comm->user, dpack, f->neig + 1, 0), "broadcast(evals)");
      BK_CUDA(cudaMemcpyAsync(pack.data(), dpack, sizeof(double) * (f->neig + 1), cudaMemcpyDeviceToHost, ctx->stream));
      BK_CUDA(cudaStreamSynchronize(ctx->stream));
      if (native) BK_TRY(peer_check(peer, ctx->stream));
      if (pack[0] < 0.0) {
        if (rank != 0) set_error("the eigensolver failed on rank 0 (status %d)", (int)pack[0]);
        return (int)pack[0];
      }
      k = (int)pack[0];
      for (int i = 0; i < f->neig; ++i) f->evals[i] = pack[1 + i];
      if (native) {
        double* Qs = peer_ptr(peer, off_Q);
        if (rank == 0)
          BK_CUDA(cudaMemcpyAsync(Qs, Zfull.p, sizeof(double) * (size_t)n * k, cudaMemcpyDeviceToDevice, ctx->stream));
        BK_TRY(peer_broadcast_sym(peer, off_Q, (long long)n * k, 0, ctx->stream));
        if (rank != 0)
          BK_CUDA(cudaMemcpyAsync(Zfull.p, Qs, sizeof(double) * (size_t)n * k, cudaMemcpyDeviceToDevice, ctx->stream));
      } else {
        COMM_CALL(comm->broadcast(comm->user, Zfull.p, (int64_t)n * k, 0), "broadcast(Q)");
      }
    }
    BK_REQUIRE(k >= 1, "eigen: no eigenpair retained");
    BK_CUDA(cudaMemcpyAsync(f->ev.p, f->evals.data(), sizeof(double) * f->neig, cudaMemcpyHostToDevice,
                            ctx->stream));
    // keep a compact n x k copy when truncation dropped most columns
    if ((size_t)k * 2 < (size_t)f->neig) {
      BK_TRY(f->Q.alloc((size_t)n * k));
      BK_CUDA(cudaMemcpyAsync(f->Q.p, Zfull.p, sizeof(double) * (size_t)n * k, cudaMemcpyDeviceToDevice,
                              ctx->stream));
      BK_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
      f->Q.swap_block(Zfull);
    }
  }
  f->k = k;
  for (double e : f->evals)
    if (!std::isfinite(e)) {
      // R/bigKRLS_Rcpp_functions.R:8-9
      set_error("Missing eigenvalues prevent bigKRLS from obtaining the regularization parameter "
                "lambda. Check for repeated observations (or other perfect linear combinations in X).");
      return BK_ERR_NUMERIC;
    }
  f->info.t_eigen = tm.stop();

  // ---- 3/5 lambda ---------------------------------------------------------------------------
  tm.start();
  DevBuf<double> z, Le_dev;
  BK_TRY(z.alloc(k));
  BK_TRY(Le_dev.alloc(16));
  const bool lprof = getenv("BK_LAMBDA_PROF") != nullptr;
  auto now_us = []() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
  };
  double tp[6] = {0, 0, 0, 0, 0, 0};
  if (lprof) {
    cudaStreamSynchronize(ctx->stream);
    tp[0] = now_us();
  }
  BK_TRY(gemm(ctx, true, false, k, 1, n, 1.0, f->Q.p, ld, f->y.p, ld, 0.0, z.p, k));
  if (lprof) {
    cudaStreamSynchronize(ctx->stream);
    tp[1] = now_us();
  }
  double lam = o.lambda;
  if (!(lam > 0.0)) {
    double L = o.L, U = o.U;
    BK_TRY(lambda_bounds(f->evals, n, &L, &U));
    if (lprof) tp[2] = now_us();
    const double tol = (o.tol > 0.0) ? o.tol : 1e-3 * n;  // R:10-12 (bigKRLS() never forwards tol)
    LooEvaluator le;
    const double* Qpanel = f->Q.p + f->c0;
    le.eval_batch = [&](const std::vector<double>& lams, double* out) -> int {
      BK_TRY(loo_batch(ctx, Qpanel, ld, nloc, k, f->ev.p, z.p, lams.data(), (int)lams.size(),
                       Le_dev.p, nullptr));
      if (native)
        BK_TRY(peer_allreduce_sum(peer, Le_dev.p, 16, ctx->stream));
      else if (multi)
        COMM_CALL(comm->allreduce_sum(comm->user, Le_dev.p, 16), "allreduce(Le)");
      // pinned landing buffer: a pageable destination makes the copy a staged, blocking one (20-30 us per pass)
      BK_CUDA(cudaMemcpyAsync(ctx->host_scratch, Le_dev.p, sizeof(double) * 16, cudaMemcpyDeviceToHost, ctx->stream));
      BK_CUDA(cudaStreamSynchronize(ctx->stream));
      memcpy(out, ctx->host_scratch, sizeof(double) * 16);
      return BK_OK;
    };
    le.batch = std::max(1, std::min(15, o.loo_batch > 0 ? o.loo_batch : 15));
    BK_TRY(lambda_search(le, L, U, tol, &lam, &f->n_probes));
    f->n_passes = le.passes;
    if (lprof) tp[3] = now_us();
  }
  f->lambda = lam;
  {
    // R/bigKRLS.R:280 (all Neig eigenvalues, R's long-double sum).  The noise tail of the spectrum (|e| <= 1e-10 lam)
    // enters through its plain sum: e/(e+lam) = e/lam to a relative 1e-10, far below the 1e-8 the value is compared at
    long double s = 0.0L;
    size_t cut = f->evals.size();
    bool sorted = true;
    for (size_t i = 1; i < f->evals.size() && sorted; ++i) sorted = f->evals[i] <= f->evals[i - 1];
    if (sorted && !f->evals.empty() && std::fabs(f->evals.back()) <= 1e-10 * lam) {
      cut = (size_t)(std::partition_point(f->evals.begin(), f->evals.end(), [&](double e) { return e >= 1e-10 * lam; }) -
                     f->evals.begin());
    }
    for (size_t i = 0; i < cut; ++i) s += (long double)(f->evals[i] / (f->evals[i] + lam));
    if (cut < f->evals.size()) {
      double tail = 0.0;
      for (size_t i = cut; i < f->evals.size(); ++i) tail += f->evals[i];
      s += (long double)(tail / lam);
    }
    f->neffective = (double)((long double)n - s);
  }
  if (lprof) {
    tp[4] = now_us();
    fprintf(stderr, "[lambda prof] Q'y %.0f us, bounds %.0f us, search %.0f us (%d passes), neffective %.0f us\n", tp[1] - tp[0],
            tp[2] - tp[1], tp[3] - tp[2], f->n_passes, tp[4] - tp[3]);
  }
  f->info.t_lambda = tm.stop();

  // ---- 4/5 coefficients, fitted values ----------------------------------------------------------
  tm.start();
  BK_TRY(f->c.alloc(n));
  std::vector<long long> cnt1, dsp1;
  block_partition(n, world, 1, cnt1, dsp1);
  if (native) {
    // every rank holds all of Q: the n coefficients and Le are formed redundantly, nothing to exchange
    BK_TRY(loo_batch(ctx, f->Q.p, ld, n, k, f->ev.p, z.p, &lam, 1, Le_dev.p, f->c.p));
  } else {
    BK_TRY(loo_batch(ctx, f->Q.p + f->c0, ld, nloc, k, f->ev.p, z.p, &lam, 1, Le_dev.p, f->c.p + f->c0));
    if (multi) {
      COMM_CALL(comm->allreduce_sum(comm->user, Le_dev.p, 16), "allreduce(Le)");
      std::vector<int64_t> counts(cnt1.begin(), cnt1.end()), displs(dsp1.begin(), dsp1.end());
      COMM_CALL(comm->allgatherv(comm->user, f->c.p, counts.data(), displs.data()), "allgatherv(coeffs)");
    }
  }
  BK_CUDA(cudaMemcpyAsync(&f->Le, Le_dev.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));

  // K-pass: KW = K [1 c {x_j|b1_j} {x_j c|b1_j c} {b0} {b0 c}] ; column 1 is yhat = K c  (R/bigKRLS.R:291)
  const int pd = pd_all;
  int nbin = 0;
  DevBuf<double> Xd, W, KW;
  BK_TRY(f->binfo.alloc(4 * std::max(1, pd) + 1));
  if (pd > 0) {
    BK_TRY(Xd.alloc((size_t)n * pd));
    std::vector<int> wh(f->which);
    DevBuf<int> whd;
    BK_TRY(whd.alloc(pd));
    BK_CUDA(cudaMemcpyAsync(whd.p, wh.data(), sizeof(int) * pd, cudaMemcpyHostToDevice, ctx->stream));
    BK_TRY(gather_columns(ctx, f->X.p, ld, n, pd, whd.p, Xd.p, ld));
    BK_TRY(column_binary_info(ctx, Xd.p, ld, n, pd, f->binfo.p, &nbin));  // synchronises (wh, whd are temporaries)
  }
  const int mw = 2 * pd + 2 + 2 * nbin;
  BK_TRY(W.alloc((size_t)n * mw));
  BK_TRY(KW.alloc((size_t)n * mw));
  BK_TRY(build_kpass_rhs(ctx, Xd.p, ld, n, pd, nbin, f->c.p, f->binfo.p, W.p, ld));
  // rows [c0, c1) of K W = (K[:, c0:c1])' W     (K symmetric)
  if (!multi)
    BK_TRY(gemm(ctx, false, false, n, mw, n, 1.0, f->K.p, ld, W.p, ld, 0.0, KW.p, ld));
  else
    BK_TRY(gemm(ctx, true, false, nloc, mw, n, 1.0, Kblk, ld, W.p, ld, 0.0, KW.p + f->c0, ld));
  BK_TRY(f->yhat.alloc(n));
  if (native) {
    double* ys = peer_ptr(peer, off_yhat);
    BK_TRY(copy_matrix(ctx, KW.p + f->c0 + ld, ld, nloc, 1, 1.0, ys + f->c0, ld));
    BK_TRY(peer_allgatherv_sym(peer, off_yhat, cnt1.data(), dsp1.data(), ctx->stream));
    BK_CUDA(cudaMemcpyAsync(f->yhat.p, ys, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    BK_TRY(copy_matrix(ctx, KW.p + f->c0 + ld, ld, nloc, 1, 1.0, f->yhat.p + f->c0, ld));
    if (multi) {
      std::vector<int64_t> counts(cnt1.begin(), cnt1.end()), displs(dsp1.begin(), dsp1.end());
      COMM_CALL(comm->allgatherv(comm->user, f->yhat.p, counts.data(), displs.data()), "allgatherv(yhat)");
    }
  }
  BK_TRY(f->sig2.alloc(1));
  BK_TRY(residual_sigmasq(ctx, f->y.p, f->yhat.p, n, f->sig2.p));  // R/bigKRLS.R:294
  f->info.t_coef = tm.stop();

  // ---- vcov(c), vcov(yhat) in y_sd^2 units (R/bigKRLS.R:299-307, 439, 446) ----------------------------
  tm.start();
  DevBuf<double> w2;
  BK_TRY(w2.alloc(k));
  BK_TRY(spectral_weights(ctx, f->ev.p, k, lam, 1, w2.p));
  if (o.vcov) {
    const double ys2 = o.y_sd * o.y_sd;
    DevBuf<double> M;
    BK_TRY(M.alloc((size_t)n * k));
    BK_TRY(f->Vc.alloc((size_t)n * nloc));
    // m = Q diag(sigmasq (ev+lam)^-2)   (bMultDiag, R:299);  vcovmatc = m Q' (bTCrossProd, R:301)
    BK_TRY(col_scale(ctx, f->Q.p, ld, n, k, w2.p, f->sig2.p, M.p, ld));
    if (!multi) {
      BK_TRY(gemm(ctx, false, true, n, n, k, ys2, M.p, ld, f->Q.p, ld, 0.0, f->Vc.p, ld, true));
      BK_TRY(symmetrize_from_lower(ctx, f->Vc.p, ld, n));
    } else {
      BK_TRY(gemm(ctx, false, true, n, nloc, k, ys2, M.p, ld, f->Q.p + f->c0, ld, 0.0, f->Vc.p, ld));
    }
    f->have_vcov = true;
    if (o.keep_vcov_fitted) {
      // K'(V K) = sigmasq Q diag((ev/(ev+lam))^2) Q'  for exact eigenpairs of K (R:307)
      DevBuf<double> w3;
      BK_TRY(w3.alloc(k));
      BK_TRY(spectral_weights(ctx, f->ev.p, k, lam, 2, w3.p));
      BK_TRY(col_scale(ctx, f->Q.p, ld, n, k, w3.p, f->sig2.p, M.p, ld));
      BK_TRY(f->Vf.alloc((size_t)n * nloc));
      if (!multi) {
        BK_TRY(gemm(ctx, false, true, n, n, k, ys2, M.p, ld, f->Q.p, ld, 0.0, f->Vf.p, ld, true));
        BK_TRY(symmetrize_from_lower(ctx, f->Vf.p, ld, n));
      } else {
        BK_TRY(gemm(ctx, false, true, n, nloc, k, ys2, M.p, ld, f->Q.p + f->c0, ld, 0.0, f->Vf.p, ld));
      }
      f->have_vf = true;
    }
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  f->info.t_vcov = tm.stop();

  // ---- 5/5 marginal effects ------------------------------------------------------------------------
  tm.start();
  if (pd > 0) {
    DevBuf<double> R, G;
    BK_TRY(f->D.alloc((size_t)n * pd));
    BK_TRY(R.alloc((size_t)n * pd));
    BK_TRY(G.alloc((size_t)k * pd));
    BK_TRY(f->var.alloc(pd));
    BK_TRY(deriv_epilogue(ctx, Xd.p + f->c0, ld, nloc, pd, nbin, KW.p + f->c0, ld, f->binfo.p, o.sigma,
                          f->D.p + f->c0, ld, R.p + f->c0, ld));
    // G = Q' R (k x pd); r'V r = sigmasq * sum_i w2_i G_i^2
    BK_TRY(gemm(ctx, true, false, k, pd, nloc, 1.0, f->Q.p + f->c0, ld, R.p + f->c0, ld, 0.0, G.p, k));
    if (multi) {
      std::vector<long long> cntD, dspD;
      block_partition(n, world, pd, cntD, dspD);
      if (native) {
        BK_TRY(peer_allreduce_sum(peer, G.p, (long long)k * pd, ctx->stream));
        // derivatives: row panels through a packed symmetric buffer (n x pd is column-major)
        double* pack = peer_ptr(peer, off_D);
        BK_TRY(copy_matrix(ctx, f->D.p + f->c0, ld, nloc, pd, 1.0, pack + dspD[rank], nloc));
        BK_TRY(peer_allgatherv_sym(peer, off_D, cntD.data(), dspD.data(), ctx->stream));
        for (int r = 0; r < world; ++r) {
          const int a = (int)((int64_t)n * r / world), b = (int)((int64_t)n * (r + 1) / world);
          if (r != rank) BK_TRY(copy_matrix(ctx, pack + dspD[r], b - a, b - a, pd, 1.0, f->D.p + a, ld));
        }
      } else {
        COMM_CALL(comm->allreduce_sum(comm->user, G.p, (int64_t)k * pd), "allreduce(Q'R)");
        DevBuf<double> pack;
        BK_TRY(pack.alloc((size_t)n * pd));
        std::vector<int64_t> counts(cntD.begin(), cntD.end()), displs(dspD.begin(), dspD.end());
        BK_TRY(copy_matrix(ctx, f->D.p + f->c0, ld, nloc, pd, 1.0, pack.p + displs[rank], nloc));
        COMM_CALL(comm->allgatherv(comm->user, pack.p, counts.data(), displs.data()), "allgatherv(D)");
        for (int r = 0; r < world; ++r) {
          const int a = (int)((int64_t)n * r / world), b = (int)((int64_t)n * (r + 1) / world);
          BK_TRY(copy_matrix(ctx, pack.p + displs[r], b - a, b - a, pd, 1.0, f->D.p + a, ld));
        }
        BK_CUDA(cudaStreamSynchronize(ctx->stream));
      }
    }
    BK_TRY(deriv_variance_spectral(ctx, G.p, k, k, pd, w2.p, f->sig2.p, f->binfo.p, o.sigma, n, f->var.p));
    f->have_deriv = true;
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  BK_CUDA(cudaMemcpyAsync(&f->sigmasq, f->sig2.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (native) {
    // every rank is through with the symmetric buffers of this fit; a timed-out wait surfaces here
    BK_TRY(peer_barrier(peer, ctx->stream));
    BK_TRY(peer_check(peer, ctx->stream));
  }
  f->info.t_deriv = tm.stop();
  if (o.K_host) {
    BK_TRY(copier_wait(&f->k_copy));
    f->K_host_done = o.K_host;
  }
  f->info.t_total = total.stop();

  f->info.n = n;
  f->info.p = p;
  f->info.neig = f->neig;
  f->info.lastkeeper = k;
  f->info.n_deriv = pd;
  f->info.lambda = f->lambda;
  f->info.Le = f->Le;
  f->info.sigmasq = f->sigmasq;
  f->info.neffective = f->neffective;
  f->info.n_probes = f->n_probes;
  f->info.n_passes = f->n_passes;
  f->info.gpu_launches = (double)(ctx->n_launches - launches0);
  return BK_OK;
}

int create_fit(bk_ctx* ctx, const double* Xs, const double* ys, bool on_device, int64_t n, int64_t p,
               const bk_fit_opts* opts, const bk_comm* comm, bk_fit** out) {
  BK_REQUIRE(ctx && Xs && ys && opts && out, "bk_fit_run: NULL argument");
  BK_REQUIRE(n >= 2 && p >= 1 && n < 2147483647LL && p < 100000, "bk_fit_run: bad dimensions");
  BK_REQUIRE(opts->sigma > 0.0, "bk_fit_run: sigma must be positive");
  BK_REQUIRE(opts->eigtrunc >= 0.0 && opts->eigtrunc <= 1.0,
             "eigtrunc must be between 0 (no truncation) and 1 (keep largest only).");  // R/bigKRLS.R:203
  BK_REQUIRE(opts->neig >= 1 && opts->neig <= n, "bk_fit_run: neig must be in 1..n");
  BK_REQUIRE(!(opts->derivative && !opts->vcov),
             "vcov.est is needed to get derivatives (derivative==TRUE requires vcov.est=TRUE).");  // R:238
  BK_REQUIRE(opts->y_sd > 0.0, "bk_fit_run: y_sd must be positive");
  for (int i = 0; i < opts->n_which; ++i)
    BK_REQUIRE(opts->which && opts->which[i] >= 0 && opts->which[i] < p,
               "which.derivatives out of range");
  if (comm && comm->world > 1 && !comm->peer)
    BK_REQUIRE(comm->allreduce_sum && comm->allgatherv && comm->broadcast,
               "bk_fit_run: communicator callbacks missing");
  BK_CUDA(bk::bind_ctx(ctx));
  bk_fit* f = new bk_fit();
  f->ctx = ctx;
  f->n = (int)n;
  f->p = (int)p;
  f->neig = (int)opts->neig;
  f->opts = *opts;
  f->opts.which = nullptr;
  if (opts->n_which > 0)
    f->which.assign(opts->which, opts->which + opts->n_which);
  else
    for (int i = 0; i < (int)p; ++i) f->which.push_back(i);
  f->pd = (int)f->which.size();
  f->rank = comm ? comm->rank : 0;
  f->world = comm ? comm->world : 1;
  f->c0 = (int)((int64_t)n * f->rank / f->world);
  f->c1 = (int)((int64_t)n * (f->rank + 1) / f->world);
  int rc = f->X.alloc((size_t)n * p);
  if (rc == BK_OK) rc = f->y.alloc(n);
  if (rc == BK_OK) {
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    cudaError_t e = cudaMemcpyAsync(f->X.p, Xs, sizeof(double) * (size_t)n * p, kind, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(f->y.p, ys, sizeof(double) * (size_t)n, kind, ctx->stream);
    if (e != cudaSuccess) {
      set_error("bk_fit_run: input copy failed: %s", cudaGetErrorString(e));
      rc = BK_ERR_CUDA;
    }
  }
  if (rc == BK_OK) rc = run_fit(f, comm);
  if (rc != BK_OK) {
    cudaStreamSynchronize(ctx->stream);
    delete f;
    return rc;
  }
  *out = f;
  return BK_OK;
}

int get_block(const bk_fit* f, const DevBuf<double>& buf, bool have, double* host, const char* what) {
  BK_REQUIRE(f && host, "getter: NULL argument");
  if (!have || !buf.p) {
    set_error("%s was not computed for this fit", what);
    return BK_ERR_STATE;
  }
  bk_ctx* ctx = f->ctx;
  BK_CUDA(bk::bind_ctx(ctx));
  const size_t cnt = (size_t)f->n * (size_t)(f->c1 - f->c0);
  // single-GPU fits hold the full matrix; multi-GPU fits hold only the owned block (Vc, Vf) or
  // the full K (offset to the owned block)
  return copy_to_host(ctx, host, buf.p, sizeof(double) * cnt, ctx->stream);
}

int get_vec(const bk_fit* f, const double* dev, size_t cnt, double* host) {
  BK_REQUIRE(f && host && dev, "getter: NULL argument or field not computed");
  bk_ctx* ctx = f->ctx;
  BK_CUDA(bk::bind_ctx(ctx));
  return copy_to_host(ctx, host, dev, sizeof(double) * cnt, ctx->stream);
}

}  // namespace

extern "C" {

void bk_fit_default_opts(bk_fit_opts* o, int64_t n, int64_t p) {
  memset(o, 0, sizeof(*o));
  o->sigma = (double)p;                      // R/bigKRLS.R:230
  o->eigtrunc = (n > 3000) ? 0.001 : 0.0;    // R/bigKRLS.R:195-201
  o->neig = n;                               // R/bigKRLS.R:194
  o->lambda = 0.0;
  o->L = -1.0;  /* not supplied */
  o->U = 0.0;
  o->tol = 0.0;
  o->derivative = 1;
  o->vcov = 1;
  o->n_which = 0;
  o->which = nullptr;
  o->y_sd = 1.0;
  o->loo_batch = 15;
  o->keep_vcov_fitted = 1;
  o->K_host = nullptr;
}

int bk_fit_run(bk_ctx* ctx, const double* Xs, const double* ys, int64_t n, int64_t p,
               const bk_fit_opts* opts, const bk_comm* comm, bk_fit** out) {
  return create_fit(ctx, Xs, ys, false, n, p, opts, comm, out);
}
int bk_fit_run_device(bk_ctx* ctx, const double* dXs, const double* dys, int64_t n, int64_t p,
                      const bk_fit_opts* opts, const bk_comm* comm, bk_fit** out) {
  return create_fit(ctx, dXs, dys, true, n, p, opts, comm, out);
}
void bk_fit_free(bk_fit* f) {
  if (!f) return;
  bk::bind_ctx(f->ctx);
  cudaStreamSynchronize(f->ctx->stream);
  delete f;
}
int bk_fit_get_info(const bk_fit* f, bk_fit_info* info) {
  BK_REQUIRE(f && info, "bk_fit_get_info: NULL argument");
  *info = f->info;
  return BK_OK;
}
int bk_fit_col_range(const bk_fit* f, int64_t* c0, int64_t* c1) {
  BK_REQUIRE(f && c0 && c1, "bk_fit_col_range: NULL argument");
  *c0 = f->c0;
  *c1 = f->c1;
  return BK_OK;
}
int bk_fit_get_K(const bk_fit* f, double* host) {
  BK_REQUIRE(f && host, "bk_fit_get_K: NULL argument");
  if (host == f->K_host_done) return BK_OK;  // delivered during the fit (bk_fit_opts.K_host)
  return get_vec(f, f->kblock(), (size_t)f->n * (size_t)(f->c1 - f->c0), host);
}
int bk_fit_get_eigenvalues(const bk_fit* f, double* host) {
  BK_REQUIRE(f && host, "bk_fit_get_eigenvalues: NULL argument");
  memcpy(host, f->evals.data(), sizeof(double) * f->evals.size());
  return BK_OK;
}
int bk_fit_get_eigenvectors(const bk_fit* f, double* host) {
  BK_REQUIRE(f, "bk_fit_get_eigenvectors: NULL argument");
  return get_vec(f, f->Q.p, (size_t)f->n * f->k, host);
}
int bk_fit_get_coeffs(const bk_fit* f, double* host) {
  BK_REQUIRE(f, "bk_fit_get_coeffs: NULL argument");
  return get_vec(f, f->c.p, f->n, host);
}
int bk_fit_get_yfitted(const bk_fit* f, double* host) {
  BK_REQUIRE(f, "bk_fit_get_yfitted: NULL argument");
  return get_vec(f, f->yhat.p, f->n, host);
}
int bk_fit_get_vcov_c(const bk_fit* f, double* host) {
  return get_block(f, f->Vc, f->have_vcov, host, "vcov.est.c");
}
int bk_fit_get_vcov_fitted(const bk_fit* f, double* host) {
  return get_block(f, f->Vf, f->have_vf, host, "vcov.est.fitted");
}
int bk_fit_get_derivatives(const bk_fit* f, double* host) {
  BK_REQUIRE(f, "bk_fit_get_derivatives: NULL argument");
  if (!f->have_deriv) {
    set_error("derivatives were not computed for this fit");
    return BK_ERR_STATE;
  }
  return get_vec(f, f->D.p, (size_t)f->n * f->pd, host);
}
int bk_fit_get_var_avgderiv(const bk_fit* f, double* host) {
  BK_REQUIRE(f, "bk_fit_get_var_avgderiv: NULL argument");
  if (!f->have_deriv) {
    set_error("derivatives were not computed for this fit");
    return BK_ERR_STATE;
  }
  return get_vec(f, f->var.p, f->pd, host);
}
int bk_fit_get_binary(const bk_fit* f, int32_t* host) {
  BK_REQUIRE(f && host, "bk_fit_get_binary: NULL argument");
  if (!f->have_deriv) {
    set_error("derivatives were not computed for this fit");
    return BK_ERR_STATE;
  }
  std::vector<double> info(4 * f->pd);
  BK_TRY(get_vec(f, f->binfo.p, 4 * f->pd, info.data()));
  for (int j = 0; j < f->pd; ++j) host[j] = info[4 * j + 2] != 0.0;
  return BK_OK;
}

int bk_fit_predict_full(const bk_fit* f, const double* newXs, int64_t m, double* pred_std, double* Knew,
                        double* se2, double* vcov_pred) {
  BK_REQUIRE(f && newXs && pred_std && m > 0 && m < 2147483647LL, "bk_fit_predict: bad arguments");
  bk_ctx* ctx = f->ctx;
  BK_CUDA(bk::bind_ctx(ctx));
  const int n = f->n, p = f->p, k = f->k, mi = (int)m;
  DevBuf<double> dN, dK, dp;
  BK_TRY(dN.alloc((size_t)m * p));
  BK_TRY(dK.alloc((size_t)m * n));
  BK_TRY(dp.alloc(m));
  BK_CUDA(cudaMemcpyAsync(dN.p, newXs, sizeof(double) * (size_t)m * p, cudaMemcpyHostToDevice, ctx->stream));
  // newdataK = bTempKernel(newdata, X, sigma)  (R/bigKRLS.R:599) ; ypred = newdataK %*% coeffs (:601)
  BK_TRY(gauss_kernel_rect(ctx, dN.p, m, mi, f->X.p, n, n, p, f->opts.sigma, dK.p, m));
  BK_TRY(gemm(ctx, false, false, mi, 1, n, 1.0, dK.p, m, f->c.p, n, 0.0, dp.p, m));
  BK_CUDA(cudaMemcpyAsync(pred_std, dp.p, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
  if (Knew) BK_TRY(copy_to_host(ctx, Knew, dK.p, sizeof(double) * (size_t)m * n, ctx->stream));
  if (se2 || vcov_pred) {
    BK_REQUIRE(f->sig2.p != nullptr, "recompute bigKRLS object with bigKRLS(,vcov.est=TRUE) to compute standard errors");
    // Knew V Knew' with V = y_sd^2 sigmasq Q (ev+lam)^-2 Q'   (R/bigKRLS.R:605-613): G = Knew Q (m x k),
    // vcov.est.pred = y_sd^2 sigmasq G diag(w2) G', se.pred^2 = its diagonal
    DevBuf<double> G, w2, s2;
    const double ys2 = f->opts.y_sd * f->opts.y_sd;
    BK_TRY(G.alloc((size_t)m * k));
    BK_TRY(w2.alloc(k));
    BK_TRY(spectral_weights(ctx, f->ev.p, k, f->lambda, 1, w2.p));
    BK_TRY(gemm(ctx, false, false, mi, k, n, 1.0, dK.p, m, f->Q.p, n, 0.0, G.p, m));
    if (se2) {
      BK_TRY(s2.alloc(m));
      BK_TRY(row_quadform(ctx, G.p, m, mi, k, w2.p, f->sig2.p, ys2, s2.p));
      BK_CUDA(cudaMemcpyAsync(se2, s2.p, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (vcov_pred) {
      DevBuf<double> M, VP;
      BK_TRY(M.alloc((size_t)m * k));
      BK_TRY(VP.alloc((size_t)m * m));
      BK_TRY(col_scale(ctx, G.p, m, mi, k, w2.p, f->sig2.p, M.p, m));
      BK_TRY(gemm(ctx, false, true, mi, mi, k, ys2, M.p, m, G.p, m, 0.0, VP.p, m, true));
      BK_TRY(symmetrize_from_lower(ctx, VP.p, m, mi));
      BK_TRY(copy_to_host(ctx, vcov_pred, VP.p, sizeof(double) * (size_t)m * m, ctx->stream));
    }
  }
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}

int bk_fit_predict(const bk_fit* f, const double* newXs, int64_t m, double* pred_std, double* Knew,
                   double* se2) {
  return bk_fit_predict_full(f, newXs, m, pred_std, Knew, se2, nullptr);
}


// ---- host-logic hooks (no GPU needed): used by the CPU test-suite ---------------------------------
int bk_host_lambda_search(const double* evals, int64_t neig, int64_t n, double L, double U, double tol,
                          int batch, bk_le_callback cb, void* user, double* lambda, double* L_out,
                          double* U_out, int* probes, int* passes) {
  BK_REQUIRE(evals && cb && lambda && neig > 0 && n > 0, "bk_host_lambda_search: bad arguments");
  std::vector<double> ev(evals, evals + neig);
  BK_TRY(lambda_bounds(ev, (int)n, &L, &U));
  if (L_out) *L_out = L;
  if (U_out) *U_out = U;
  LooEvaluator le;
  le.batch = std::max(1, std::min(15, batch > 0 ? batch : 15));
  le.eval_batch = [&](const std::vector<double>& lams, double* out) -> int {
    if (cb(user, lams.data(), (int)lams.size(), out) != 0) {
      set_error("bk_host_lambda_search: callback failed");
      return BK_ERR_ARG;
    }
    return BK_OK;
  };
  int np = 0;
  BK_TRY(lambda_search(le, L, U, tol > 0.0 ? tol : 1e-3 * (double)n, lambda, &np));
  if (probes) *probes = np;
  if (passes) *passes = le.passes;
  return BK_OK;
}

int bk_host_deflate_test(const double* d, const double* z, int n, int n1, double beta, int* K,
                         double* dlam, double* w, int32_t* nd_cols, int32_t* nd_type,
                         int32_t* defl_cols, double* defl_vals, int* nrot, int32_t* rot_idx,
                         double* rot_cs) {
  BK_REQUIRE(d && z && K && n > 0 && n1 > 0 && n1 < n, "bk_host_deflate_test: bad arguments");
  MergePlan plan;
  host_deflate(d, z, n, n1, beta, &plan);
  *K = plan.K;
  for (int i = 0; i < plan.K; ++i) {
    dlam[i] = plan.dlam[i];
    w[i] = plan.w[i];
    nd_cols[i] = plan.nd_cols[i];
    nd_type[i] = plan.nd_type[i];
  }
  for (int i = 0; i < n - plan.K; ++i) {
    defl_cols[i] = plan.defl_cols[i];
    defl_vals[i] = plan.defl_vals[i];
  }
  *nrot = (int)plan.rots.size();
  for (size_t i = 0; i < plan.rots.size(); ++i) {
    rot_idx[2 * i] = plan.rots[i].pj;
    rot_idx[2 * i + 1] = plan.rots[i].nj;
    rot_cs[2 * i] = plan.rots[i].c;
    rot_cs[2 * i + 1] = plan.rots[i].s;
  }
  return BK_OK;
}

}  // extern "C"
