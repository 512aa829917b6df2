// Two-stage tridiagonalisation, stage 2: symmetric band (bandwidth b = 64) -> tridiagonal by bulge chasing,
// and the back-transformation of eigenvectors through both stages.
// Executable specification: tests/twostage_prototype.py (sb2st, apply_q2, apply_q1).
//
// Bulge chasing (chase_kernel).  Sweep j annihilates column j below the sub-diagonal with a Householder
// reflector of length <= b and chases the bulge down the band in hops of b rows: per hop a two-sided update
// of a b x b diagonal block, a right-update of the b x b block below it, and a new reflector that removes the
// first column of the bulge.  Sweeps are pipelined: one CTA per sweep, sweep j may execute hop t once
// sweep j-1 has finished hop t+1 (progress counters in global memory, acquire/release) - about n/(2b)
// sweeps are in flight, which matches the SM count at the target size.  The band lives in L2 (n x 2b doubles).
//
// Q2 back-transformation (q2_apply_kernel).  The reflector (sweep j, hop t) acts on rows
// [j+1+tb, j+1+(t+1)b).  For a fixed hop index t the row window slides up by ONE row per sweep, so a CTA that
// owns hop index t keeps the 64 x k window of Z in SHARED MEMORY for all sweeps, loads one new row and
// retires one finished row per sweep.  CTA t consumes the rows retired by CTA t-1 (flag per hop index), so the
// whole back-transformation is a software pipeline over hop indices with no grid barrier.
#include <climits>
#include <type_traits>
#include <cstdlib>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"
#include "peer.cuh"

namespace bk {

static constexpr int CB = 64;        // bandwidth
static constexpr int LDAB = 2 * CB;  // working band storage (bulges reach 2b-1 below the diagonal)
static constexpr int CH_NT = 256;

__device__ __forceinline__ int ld_acquire_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_i32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

struct ChaseArgs {
  double* AB;     // LDAB x n
  int n;
  double* VV;     // n x n (ld n): column j holds the reflectors of sweep j at their row positions
  double* TAU;    // maxhops x n
  int maxhops;
  int* prog;      // n: hops completed by sweep j (INT_MAX when finished)
  double* d;
  double* e;
  long long* prof;  // optional (BK_CHASE_PROF): [0] hops, [1] wait, [2] group A busy, [3] group B busy, [4] hop total
  long long* trace;        // optional (BK_CHASE_TRACE=j0): SM clock at 12 points of hops 0..TR_HOPS-1 of sweeps j0..j0+TR_SWEEPS-1
  int trace_j0, trace_t0;
  double* mb;              // mailboxes: 2 x maxhops x MB_STRIDE doubles (ws kernel)
  unsigned long long* ll;  // early hand-off slots: LL_RING x maxhops x LL_PER_HOP x 2 tagged words (LLP kernel)
  int* err;                // set when a hand-off wait gives up (cannot happen with all CTAs co-resident)
};

// ---- early hand-off between neighbouring sweeps ("late column") -----------------------------------------------
// Hop (j, t) reads what sweep j-1 left behind.  All of it is final once sweep j-1 has finished hop t, EXCEPT the
// last column of the block below the diagonal block and the two corners: those 65 numbers are the first column of
// the diagonal block of hop (j-1, t+1) and the head of its new bulge column - known to sweep j-1 about 40 % into
// that hop.  Waiting for the completion flag of (j-1, t+1) (store fence, flag flight, poll, then 64 KB of loads)
// made the lag between neighbouring sweeps two full hops plus ~4500 cycles of flag latency; the lag is what bounds
// the kernel (n sweeps, one after the other).  The LLP kernel therefore
//   * hands the 65 late numbers over through self-validating slots (each 8-byte word = 32 bits of payload + a
//     32-bit tag naming the producing sweep, as in NCCL's LL protocol: no fence, no separate flag), written as
//     soon as they exist and polled by exactly the threads that need them; the producer does not store them into
//     the band at all (the consumer owns those positions from then on), so the early start creates no race;
//   * gives every CTA a ninth warp that does nothing but move data (chase_ws_kernel): it watches the completion flag
//     of sweep j-1, stages everything else of the NEXT hop in shared memory with 16-byte cp.async while the eight
//     compute warps work on the current one, and publishes this sweep's completion flags - the store fence of a
//     release (1200-2500 cycles: it drains the whole SM's outstanding stores) and the 50 KB fetch of a hop
//     (~2000 cycles) no longer sit on any compute thread's path.
static constexpr int LL_RING = 4;           // sweeps whose slots are live at once (>= 2 by the dependency order)
static constexpr int LL_PER_HOP = CB + 1;   // rows 0..63 of the column, then the corner of the block below
static constexpr int TR_SWEEPS = 4, TR_HOPS = 48, TR_EV = 12;
__device__ __forceinline__ long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return (long long)t;
}
// global-timer stamps (kind 0: hop starts, 1: stores issued, 2: flag released, 3: staging of the hop commanded)
#define CH_GSTAMP(kind, jj, tt_)                                                                         \
  if (TRACE && (jj) >= a.trace_j0 && (jj) < a.trace_j0 + 16 && (tt_) >= a.trace_t0 && (tt_) < a.trace_t0 + 64)       \
    a.trace[(size_t)TR_SWEEPS * TR_HOPS * TR_EV + ((size_t)(kind) * 16 + ((jj) - a.trace_j0)) * 64 + (tt_) - a.trace_t0] = global_ns();
static constexpr int kSpinLimit = 1 << 22;  // polls before a wait gives up (seconds)
static constexpr int SBN = CB + 2, SDN = CB + 1;  // column strides of the staging buffers (see chase_kernel)
// Mailboxes.  A full hop (j, t) whose successor hop (j+1, t) is full as well does not write its two blocks back into
// the band: it writes them, shifted by one row and one column, into the mailbox (j & 1, t) in exactly the layout of
// sweep j+1's staging buffers - block below: (r, c) at [c * 64 + r], diagonal block: (row, col) at [4096 + row + 65 col]
// (its last row is row 0 of the block below).  Staging such a hop is two contiguous TMA bulk copies issued by one
// thread instead of ~100 cp.async per transfer thread gathering 128 band columns.  The band keeps what the mailboxes
// do not carry: sweep 0's input, the clipped hops at the end of every sweep, column j and d[j] for hop 0.
static constexpr int MB_B = CB * CB, MB_D = CB * SDN, MB_STRIDE = MB_B + MB_D;

__device__ __forceinline__ void ll_store(unsigned long long* slot, double val, unsigned tag) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(val);
  const unsigned long long hi = (unsigned long long)tag << 32;
  const unsigned long long w0 = hi | (u & 0xffffffffull), w1 = hi | (u >> 32);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};\n" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_try(const unsigned long long* slot, unsigned tag, double& val) {
  unsigned long long w0, w1;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  val = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
  return (unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag;
}
__device__ __forceinline__ bool spin_giveup(int& spins, int* err) {
  if (((++spins) & 1023) != 0) return false;
  if (spins > kSpinLimit || *(volatile int*)err != 0) {
    *(volatile int*)err = 1;
    return true;
  }
  return false;
}
__device__ __forceinline__ double ll_wait(const unsigned long long* slot, unsigned tag, int* err) {
  double v;
  int spins = 0;
  while (!ll_try(slot, tag, v))
    if (spin_giveup(spins, err)) break;
  return v;
}
__device__ __forceinline__ void wait_prog(const int* p, int target, int* err) {
  int spins = 0;
  while (ld_acquire_i32(p) < target)
    if (spin_giveup(spins, err)) break;
}

__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, 128;\n" ::"r"(id) : "memory"); }

// Householder scalars from alpha = x[0] and s = |x[1:]|^2: (I - tau v v') x = beta e1, v = [1, x[1:] * scale]
__device__ __forceinline__ void house_scalars(double alpha, double s, double& beta, double& tau, double& scale) {
  if (s == 0.0) {
    beta = alpha;
    tau = 0.0;
    scale = 0.0;
  } else {
    beta = -copysign(sqrt(alpha * alpha + s), alpha);
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
}

// Group B's balanced, coalesced map of the lower triangle of a 64 x 64 block onto 128 threads x 17 slots: columns p
// and 63-p hold 65 elements together; warp w takes the pairs p = w, w+4, ..., w+28, slots 0..63 of a pair go to
// lanes (two rounds), slot 64 (row 63 of column 63-p) to lane (p - w)/4 in round 16.  Returns false for an empty slot.
__device__ __forceinline__ bool tri_slot(int q, int gw, int lane, int& row, int& col) {
  if (q < 16) {
    const int p = gw + 4 * (q >> 1), slot = lane + 32 * (q & 1);
    if (slot < CB - p) {
      col = p;
      row = p + slot;
    } else {
      col = CB - 1 - p;
      row = slot - 1;  // (63 - p) + (slot - (64 - p))
    }
    return true;
  }
  col = CB - 1 - (gw + 4 * lane);
  row = CB - 1;
  return lane < 8;
}

// The CTA is split into two groups of 128 threads that work concurrently inside a hop: group A owns the
// reflector chain (block below the diagonal block: right-update, new reflector, left-update), group B the
// two-sided update of the diagonal block.  Both hold their 64 x 64 block in registers (32 doubles per
// thread: row r = thread % 64, columns cq, cq+2, ...) and use shared memory only for the reductions.
__global__ void __launch_bounds__(CH_NT, 1) chase_kernel(ChaseArgs a) {
  extern __shared__ __align__(16) double chase_sm[];
  double (*Ds)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(chase_sm);
  double (*Bs)[CB + 1] = Ds + CB;
  __shared__ double v[CB], v2[CB], w[CB], w2[CB], wa[CB];
  __shared__ double pA[2][CB], pD[2][CB], redA[4], redB[4];
  __shared__ double s_alpha, s_tau0, s_tau2;
  const int n = a.n;
  double* AB = a.AB;
  const int tid = threadIdx.x;
  const int grp = tid >> 7, gt = tid & 127, gw = gt >> 5, lane = tid & 31;
  const int r = gt & (CB - 1), cq = gt >> 6;
  for (int j = blockIdx.x; j < n - 2; j += gridDim.x) {
    double tau = 0.0;
    for (int t = 0;; ++t) {
      const int lo = j + 1 + t * CB;
      if (lo >= n) break;
      const int hi = min(n, lo + CB), L = hi - lo;
      if (t == 0 && L < 2) break;
      const int hi2 = min(n, hi + CB), L2 = hi2 - hi;
      long long tp0 = 0, tp1 = 0;
      long long tpa[6] = {0, 0, 0, 0, 0, 0};
      if (a.prof) tp0 = clock64();
      // ---- wait until sweep j-1 is two hops ahead (also orders the v <- v2 copy of the previous hop) -------
      if (j > 0 && tid == 0) wait_prog(a.prog + j - 1, t + 2, a.err);
      __syncthreads();
      if (a.prof) tp1 = clock64();
      const bool has_b = hi < n;
      const bool more = has_b && (L2 >= 2);
      const int Lrt = L, L2rt = L2;
      // The hop body is instantiated twice: FULL (64 x 64 blocks, the common case) has no bounds predicates and
      // constant address strides; the generic version handles the clipped blocks at the end of the band.
      auto hop = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int L = FULL ? CB : Lrt, L2 = FULL ? CB : L2rt;
        double* const pb = AB + (size_t)lo * LDAB + (L + r);  // block below the diagonal block, row hi + r
        double* const pd0 = AB + (size_t)lo * LDAB;           // diagonal block: (row, col) at pd0[row + col (LDAB-1)]
        // ---- every global load of the hop is issued up front ------------------------------------------------
        double x[32];
        double colv = 0.0;
        if (grp == 0) {
  #pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = cq + 2 * i;
            const int dd = L + r - c;  // row hi+r, column lo+c: AB[dd + (lo+c) LDAB] = pb[c (LDAB-1)]
            x[i] = (FULL || (r < L2 && c < L && dd < LDAB)) ? __ldcg(pb + c * (LDAB - 1)) : 0.0;
          }
          if (t == 0 && gt < L) colv = __ldcg(AB + (1 + gt) + (size_t)j * LDAB);
        } else {
  #pragma unroll
          for (int q = 0; q < 17; ++q) {
            int row, col;
            const bool ok = tri_slot(q, gw, lane, row, col) && (FULL || row < L);
            x[q] = ok ? __ldcg(pd0 + row + col * (LDAB - 1)) : 0.0;  // AB[(row-col) + (lo+col) LDAB]
          }
        }
        if (t == 0) {
          // first reflector of the sweep (group A): annihilate column j below the sub-diagonal
          if (grp == 0) {
            double s = (gt >= 1 && gt < L) ? colv * colv : 0.0;
            s = warp_sum(s);
            if (lane == 0) redA[gw] = s;
            if (gt == 0) s_alpha = colv;
            group_bar(1);
            double beta, tau0, scale;
            house_scalars(s_alpha, redA[0] + redA[1] + redA[2] + redA[3], beta, tau0, scale);
            if (gt < L) v[gt] = (gt == 0) ? 1.0 : colv * scale;
            if (gt >= 1 && gt < L) AB[(1 + gt) + (size_t)j * LDAB] = 0.0;
            if (gt == 0) {
              AB[1 + (size_t)j * LDAB] = beta;
              a.e[j] = beta;
              a.d[j] = __ldcg(AB + (size_t)j * LDAB);
              s_tau0 = tau0;
            }
          }
          __syncthreads();
          tau = s_tau0;
        }
        if (grp == 1) {
          // ================= group B: two-sided update of D = A[lo:hi, lo:hi] =================================
          if (gt < L) a.VV[(size_t)(lo + gt) + (size_t)j * n] = v[gt];
          if (gt == 0) a.TAU[t + (size_t)j * a.maxhops] = tau;
  #pragma unroll
          for (int q = 0; q < 17; ++q) {
            int row, col;
            if (tri_slot(q, gw, lane, row, col) && (FULL || row < L)) {
              Ds[row][col] = x[q];
              Ds[col][row] = x[q];
            }
          }
          group_bar(2);
          {
            // w = tau D v: half of the columns per thread, row r
            double sa[4] = {0.0, 0.0, 0.0, 0.0};
            if (r < L) {
              const int c0 = cq * 32;
  #pragma unroll
              for (int c = 0; c < 32; ++c)
                if (c0 + c < L) sa[c & 3] = fma(Ds[r][c0 + c], v[c0 + c], sa[c & 3]);
            }
            pD[cq][r] = (sa[0] + sa[1]) + (sa[2] + sa[3]);
          }
          group_bar(2);
          {
            const double wr0 = (gt < L) ? tau * (pD[0][gt] + pD[1][gt]) : 0.0;
            if (gt < CB) w[gt] = wr0;
            double s = (gt < L) ? wr0 * v[gt] : 0.0;
            s = warp_sum(s);
            if (lane == 0) redB[gw] = s;
          }
          group_bar(2);
          {
            const double al = -0.5 * tau * (redB[0] + redB[1] + redB[2] + redB[3]);
            if (gt < CB) w2[gt] = (gt < L) ? fma(al, v[gt], w[gt]) : 0.0;  // w + al v
            group_bar(2);
  #pragma unroll
            for (int q = 0; q < 17; ++q) {
              int row, col;
              if (tri_slot(q, gw, lane, row, col) && (FULL || row < L))
                pd0[row + col * (LDAB - 1)] = x[q] - v[row] * w2[col] - w2[row] * v[col];
            }
          }
        } else if (has_b) {
          // ================= group A: Bk = A[hi:hi2, lo:hi] <- H2 (Bk H) and the next reflector ====================
          {
            double sa[4] = {0.0, 0.0, 0.0, 0.0};
  #pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = cq + 2 * i;
              if (c < L) sa[i & 3] = fma(x[i], v[c], sa[i & 3]);
            }
            pA[cq][r] = (sa[0] + sa[1]) + (sa[2] + sa[3]);
          }
          group_bar(1);
          if (a.prof && tid == 0) tpa[1] = clock64();
          {
            const double ur = tau * (pA[0][r] + pA[1][r]);
  #pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = cq + 2 * i;
              if (c < L) x[i] = fma(-ur, v[c], x[i]);
            }
          }
          if (more) {
            // reflector from the first column of the bulge (held by the cq == 0 half: x[0] = Bk[r][0])
            double s = (cq == 0 && r >= 1 && r < L2) ? x[0] * x[0] : 0.0;
            s = warp_sum(s);
            if (lane == 0) redA[gw] = s;
            if (gt == 0) s_alpha = x[0];
            group_bar(1);
            if (a.prof && tid == 0) tpa[2] = clock64();
            double beta, tau2, scale;
            house_scalars(s_alpha, redA[0] + redA[1] + redA[2] + redA[3], beta, tau2, scale);
            if (cq == 0) {
              if (r < L2) v2[r] = (r == 0) ? 1.0 : x[0] * scale;
              x[0] = (r == 0) ? beta : 0.0;
            }
            if (gt == 0) s_tau2 = tau2;
  #pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = cq + 2 * i;
              if (r < L2 && c < L) Bs[r][c] = x[i];
            }
            group_bar(1);
            if (a.prof && tid == 0) tpa[3] = clock64();
            {
              // wa[c] = tau2 v2' Bk[:, c]: half of the rows per thread, column r (used as the column index here)
              double sa[4] = {0.0, 0.0, 0.0, 0.0};
              if (r < L) {
                const int q0 = cq * 32;
  #pragma unroll
                for (int q = 0; q < 32; ++q)
                  if (q0 + q < L2) sa[q & 3] = fma(v2[q0 + q], Bs[q0 + q][r], sa[q & 3]);
              }
              pA[cq][r] = (sa[0] + sa[1]) + (sa[2] + sa[3]);
            }
            group_bar(1);
            if (a.prof && tid == 0) tpa[4] = clock64();
            if (gt < CB) wa[gt] = tau2 * (pA[0][gt] + pA[1][gt]);
            group_bar(1);
            if (a.prof && tid == 0) tpa[5] = clock64();
            {
              const double v2r = (r < L2) ? v2[r] : 0.0;
  #pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int c = cq + 2 * i;
                if (c >= 1 && c < L) x[i] = fma(-v2r, wa[c], x[i]);
              }
            }
          }
  #pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = cq + 2 * i;
            const int dd = L + r - c;
            if (FULL || (r < L2 && c < L && dd < LDAB)) pb[c * (LDAB - 1)] = x[i];
          }
        }
      };
      if (L == CB && has_b && L2 == CB)
        hop(std::true_type{});
      else
        hop(std::false_type{});
      long long tp2 = 0;
      if (a.prof) tp2 = clock64();
      __syncthreads();
      if (a.prof && (tid == 0 || tid == 128) && blockIdx.x == 0) {
        const long long tp3 = clock64();
        if (tid == 0) {
          a.prof[0] += 1;
          a.prof[1] += tp1 - tp0;
          a.prof[2] += tp2 - tp1;
          a.prof[4] += tp3 - tp0;
          if (tpa[5] != 0) {  // a hop with a full group-A chain
            a.prof[5] += 1;
            a.prof[6] += tpa[1] - tp1;     // loads + mat-vec partials
            a.prof[7] += tpa[2] - tpa[1];  // rank-1 + norm
            a.prof[8] += tpa[3] - tpa[2];  // scalars + staging
            a.prof[9] += tpa[5] - tpa[3];  // column dots
            a.prof[10] += tp2 - tpa[5];    // apply + stores
          }
        } else {
          a.prof[3] += tp2 - tp1;
        }
      }
      if (more) {
        if (tid < L2) v[tid] = v2[tid];
        tau = s_tau2;
      }
      if (tid == 128) st_release_i32(a.prog + j, t + 1);  // release is cumulative over the CTA barrier
      if (!more) break;
    }
    __syncthreads();
    if (tid == 128) st_release_i32(a.prog + j, INT_MAX);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Warp-specialised bulge chasing: 8 compute warps (the two groups of chase_kernel, same arithmetic in the same
// order - d, e and the reflectors are bit-identical) + 1 transfer warp.
//   transfer warp, per hop:  [flag of sweep j-1] -> stage the next hop's blocks (cp.async, completion on the mbarrier
//                            s_full) ; [named barrier: this hop's stores are issued] -> release this sweep's flag
//   compute warps, per hop:  wait s_full -> blocks from the staging buffers into registers -> arrive on s_empty
//                            -> late numbers from the hand-off slots -> hop -> stores -> arrive on the named barrier
// ---------------------------------------------------------------------------------------------------------
// 8 compute warps + a warp group of 4 transfer warps (registers are handed out per warp group; setmaxnreg moves most
// of the transfer group's share to the compute warps: 232 against 168 registers)
static constexpr int WS_NT = CH_NT + 128;

__device__ __forceinline__ void cta_sync256() { asm volatile("bar.sync 0, 256;\n" ::: "memory"); }
// non-blocking phase test (try_wait may sleep in hardware; the transfer warp polls several things in turn)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded like every other wait of this kernel (try_wait sleeps in hardware for a while before it returns false)
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, unsigned parity, int* err) {
  int spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > (1 << 16)) {
      if (spins > (1 << 20) || *(volatile int*)err != 0) {
        *(volatile int*)err = 1;
        break;
      }
    }
  }
}
// the mbarrier gets one arrival from this thread once all of its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// Transfer warps (tw = warp 0..3, cl = lane): stage the blocks of the hop at `lo` from the (zero-padded) band.
//   block below: element (r, c) lives at pdn[64 + r + 127 c]; pair k of column c = rows 2k - s, 2k - s + 1 with
//   s = c & 1 (16-byte aligned in the band) -> Bn[66 c + 2 k] (k < 32, and k = 32 for odd c: row 63 and a padding row)
//   diagonal block: column col, rows col + 2k, col + 2k + 1 at pdn[128 col + 2k] -> Dn[66 col + 2k]; columns p and
//   63-p hold 33 pairs together
__device__ __forceinline__ void ws_stage_band(const double* __restrict__ AB, int lo, bool with_b, double* Bn, double* Dn,
                                              int tw, int cl) {
  const double* const pdn = AB + (size_t)lo * LDAB;
  if (with_b) {
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int c = tw + 4 * m;
      cp_async16(Bn + c * SBN + 2 * cl, pdn + CB + c * (LDAB - 1) - (c & 1) + 2 * cl, 16);
    }
    if (tw == 0) {
      const int c = 2 * cl + 1;
      cp_async16(Bn + c * SBN + CB, pdn + CB + c * (LDAB - 1) - 1 + CB, 16);
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int p = tw + 4 * m;
    const int kp = (CB - p + 1) >> 1;  // pairs of column p; column 63-p has 33 - kp
    const int col = (cl < kp) ? p : CB - 1 - p, k = (cl < kp) ? cl : cl - kp;
    cp_async16(Dn + col * (SDN + 1) + 2 * k, pdn + (size_t)col * LDAB + 2 * k, 16);
  }
  if (tw == 1) {  // the 33rd pair of each column pair: the last pair of column 63-p
    const int p = cl, kp = (CB - p + 1) >> 1;
    const int col = CB - 1 - p, k = 32 - kp;
    cp_async16(Dn + col * (SDN + 1) + 2 * k, pdn + (size_t)col * LDAB + 2 * k, 16);
  }
}

// Geometry of hop (j, t) in an n x n band (the band storage is zero-padded by 3 x 64 columns, so every hop computes
// on full 64 x 64 blocks; the clipped sizes only decide which hops exist and what is stored outside the band).
struct HopGeom {
  int lo, L, L2;
  bool has_b, more;
};
__device__ __forceinline__ bool hop_geom(int n, int j, int t, HopGeom& g) {
  g.lo = j + 1 + t * CB;
  if (g.lo >= n) return false;
  g.L = min(CB, n - g.lo);
  if (t == 0 && g.L < 2) return false;
  g.has_b = g.lo + CB < n;
  g.L2 = g.has_b ? min(CB, n - g.lo - CB) : 0;
  g.more = g.has_b && g.L2 >= 2;
  return true;
}

template <bool TRACE>
__global__ void __launch_bounds__(WS_NT, 1) chase_ws_kernel(ChaseArgs a) {
  extern __shared__ __align__(16) double chase_sm[];
  double (*Ds)[CB + 1] = reinterpret_cast<double (*)[CB + 1]>(chase_sm);
  double (*Bs)[CB + 1] = Ds + CB;
  // staging of the next hop's blocks.  From a mailbox: block below (r, c) at [c * 64 + r], diagonal block (row, col) at
  // [row + col * 65].  From the band (sweep 0, last hop of a sweep): block below at [c * 66 + r + (c & 1)] - the stride
  // and the one-element shift of the odd columns make every pair of rows a 16-byte aligned copy on both sides.
  double* const Bn = chase_sm + 2 * CB * (CB + 1);
  double* const Dn = Bn + CB * SBN;
  double* const colv_s = Dn + CB * SDN;  // hop 0: column j below the diagonal
  __shared__ double v[CB], v2[CB], w[CB], w2[CB], wa[CB];
  __shared__ double pA[2][CB], pD[2][CB], redA[4], redB[4];
  __shared__ double s_alpha, s_tau0, s_tau2;
  __shared__ __align__(8) uint64_t s_full, s_empty, s_done;
  __shared__ int s_cmd[2];
  const int n = a.n;
  double* AB = a.AB;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&s_full, 128);    // one arrival per transfer thread (deferred until its copies have landed)
    mbar_init(&s_empty, CH_NT); // every compute thread, once its staged numbers are in registers
    mbar_init(&s_done, CH_NT);  // every compute thread, once its stores of the hop are issued
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (tid >= CH_NT) {
    // =============================== transfer warps =====================================================
    // Thread 0 of the group watches three things in turn - the completion flag of sweep j-1, the staging buffers
    // (s_empty), this hop's stores (s_done) - and hands the group one command per round through shared memory.
    // (Measured alternatives: one transfer warp - its 100 copies per hop take 8000 cycles to issue, a warp has only
    // a few cp.async in flight; three staging warps + one publishing warp - 1750..5000 cycles per staging.)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;\n");
    const int tt = tid - CH_NT, tw = tt >> 5, cl = tt & 31;
    unsigned h = 0;       // hops of this CTA so far: parity of s_done
    unsigned staged = 0;  // stagings issued so far: parity of s_full / s_empty
    unsigned round = 0;
    // Hop (j, t) reads what sweep j-1 had finished after ITS hop t (the late numbers travel by the slots).  The last
    // hop of a sweep (no block below) reads the band, where the corner was left by sweep j-2's last hop.
    auto stage_ready = [&](int j, int t, bool has_b) -> bool {
      if (j > 0 && ld_acquire_i32(a.prog + j - 1) < t + 1) return false;
      if (!has_b && j > 1 && ld_acquire_i32(a.prog + j - 2) < t + 2) return false;
      if (staged > 0 && !mbar_test(&s_empty, (staged - 1) & 1u)) return false;  // previous staging still in use
      return true;
    };
    // mailbox staging: thread 0 alone, at once (the other threads only add their arrivals a round later)
    auto stage_mb = [&](int j, int t) {
      const double* mb = a.mb + ((size_t)((j - 1) & 1) * a.maxhops + t) * MB_STRIDE;
      fence_proxy_async();  // the staging buffers were read, the mailbox acquired, in the generic proxy
      mbar_expect_tx(&s_full, (unsigned)(sizeof(double) * MB_STRIDE));
      tma_bulk_g2s(Bn, mb, (unsigned)(sizeof(double) * MB_B), &s_full);
      tma_bulk_g2s(Dn, mb + MB_B, (unsigned)(sizeof(double) * MB_D), &s_full);
    };
    auto stage = [&](int j, int t, int lo, bool has_b) {
      const bool from_mb = j > 0 && has_b;
      if (from_mb) {
        // thread 0 has issued the two bulk copies before it gave the command
      } else {
        ws_stage_band(AB, lo, has_b, Bn, Dn, tw, cl);
      }
      if (t == 0 && tt >= 32 && tt < 32 + CB) {  // column j (not thread 0: its arrival may have gone with expect_tx)
        const int g = tt - 32;
        cp_async8(colv_s + g, AB + (1 + g) + (size_t)j * LDAB, 8);  // zero beyond row n (padding)
      }
      if (!(from_mb && tt == 0)) cp_async_arrive_noinc(&s_full);
      ++staged;
    };
    for (int j = blockIdx.x; j < n - 2; j += gridDim.x) {
      bool any = false;
      for (int t = 0;; ++t) {
        HopGeom g;
        if (!hop_geom(n, j, t, g)) break;
        any = true;
        long long* const trp = (TRACE && j >= a.trace_j0 && (j - a.trace_j0) % (int)gridDim.x == 0 &&
                                    (j - a.trace_j0) / (int)gridDim.x < TR_SWEEPS && t >= a.trace_t0 && t < a.trace_t0 + TR_HOPS)
                                   ? a.trace + ((size_t)((j - a.trace_j0) / (int)gridDim.x) * TR_HOPS + t - a.trace_t0) * TR_EV
                                   : nullptr;
        HopGeom gn;
        const bool have_next = g.more && hop_geom(n, j, t + 1, gn);
        int spins = 0;
        // commands: 1 publish this hop (its stores are issued), 2 stage hop 0, 3 stage the next hop, 4 give up
        bool need_first = (t == 0), need_stage = have_next, need_rel = true;
        while (need_first || need_stage || need_rel) {
          if (tt == 0) {
            // poll until there is something to do (the other transfer threads wait at the barrier below)
            int cmd = 0;
            while (cmd == 0) {
              if (need_first) {
                if (stage_ready(j, 0, g.has_b)) cmd = 2;
              } else if (need_rel && mbar_test(&s_done, h & 1u)) {
                cmd = 1;
              } else if (need_stage && stage_ready(j, t + 1, gn.has_b)) {
                cmd = 3;
              }
              if (cmd == 0 && spin_giveup(spins, a.err)) cmd = 4;
            }
            if (cmd == 2 && j > 0 && g.has_b) stage_mb(j, 0);
            if (cmd == 3 && j > 0 && gn.has_b) stage_mb(j, t + 1);
            s_cmd[round & 1u] = cmd;
          }
          asm volatile("bar.sync 3, 128;\n" ::: "memory");
          const int cmd = s_cmd[round & 1u];
          ++round;
          if (cmd == 1) {
            if (tt == 0) {
              st_release_i32(a.prog + j, g.more ? t + 1 : INT_MAX);
              if (TRACE && trp) trp[11] = clock64();
              CH_GSTAMP(2, j, t)
            }
            need_rel = false;
          } else if (cmd == 2) {
            if (tt == 0) { CH_GSTAMP(3, j, 0) }
            stage(j, 0, g.lo, g.has_b);
            need_first = false;
          } else if (cmd == 3) {
            if (tt == 0) { CH_GSTAMP(3, j, t + 1) }
            if (TRACE && trp && tt == 0) trp[8] = clock64();
            stage(j, t + 1, gn.lo, gn.has_b);
            if (TRACE && trp && tt == 0) trp[9] = clock64();
            need_stage = false;
          } else if (cmd == 4) {
            need_first = need_stage = need_rel = false;
          }
        }
        ++h;
        if (!g.more) break;
      }
      if (!any && tt == 0) st_release_i32(a.prog + j, INT_MAX);
    }
    return;
  }

  // ================================= compute warps ========================================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 224;\n");
  const int grp = tid >> 7, gt = tid & 127, gw = gt >> 5, lane = tid & 31;
  const int r = gt & (CB - 1), cq = gt >> 6;
  unsigned h = 0;
  for (int j = blockIdx.x; j < n - 2; j += gridDim.x) {
    double tau = 0.0;
    for (int t = 0;; ++t) {
      HopGeom g;
      if (!hop_geom(n, j, t, g)) break;
      const int lo = g.lo, L = g.L;
      const bool has_b = g.has_b, more = g.more;
      long long* const trp = (TRACE && j >= a.trace_j0 && (j - a.trace_j0) % (int)gridDim.x == 0 &&
                                  (j - a.trace_j0) / (int)gridDim.x < TR_SWEEPS && t >= a.trace_t0 && t < a.trace_t0 + TR_HOPS)
                                 ? a.trace + ((size_t)((j - a.trace_j0) / (int)gridDim.x) * TR_HOPS + t - a.trace_t0) * TR_EV
                                 : nullptr;
#define CH_TRACE(ev, who) \
  if (TRACE && trp && tid == (who)) trp[ev] = clock64();
      CH_TRACE(0, 0)
      // Who hands what to whom:
      //   late    hop (j-1, t+1) exists exactly when this hop has a block below its diagonal block; it hands over, through
      //           the slots, the first column of its diagonal block (our corner D[63][63] and B[0..62][63]) and the head of
      //           its bulge column (our B[63][63]).  Conversely every hop t >= 1 is such a hop for sweep j+1: it hands
      //           those 65 numbers over and does not store them.
      //   blocks  a hop with a block below takes its blocks from the mailbox of hop (j-1, t); the last hop of a sweep
      //           (and all of sweep 0) takes them from the band.  So a hop writes into its mailbox when hop (j+1, t)
      //           has a block below (lo + 65 < n), into the band otherwise.
      const bool late = j > 0 && has_b;
      const bool emit = t >= 1;
      const bool from_mb = j > 0 && has_b;
      const bool to_mb = lo + CB + 1 < n;
      double* const mbw = a.mb + ((size_t)(j & 1) * a.maxhops + t) * MB_STRIDE;
      const unsigned long long* const ll_in =
          a.ll + ((size_t)((j + LL_RING - 1) % LL_RING) * a.maxhops + (t + 1)) * (2 * LL_PER_HOP);
      unsigned long long* const ll_out = a.ll + ((size_t)(j % LL_RING) * a.maxhops + t) * (2 * LL_PER_HOP);
      const unsigned tag_in = (unsigned)j, tag_out = (unsigned)j + 1u;
      double* const pb = AB + (size_t)lo * LDAB + (CB + r);  // block below the diagonal block, row lo + 64 + r
      double* const pd0 = AB + (size_t)lo * LDAB;            // diagonal block: (row, col) at pd0[row + col (LDAB-1)]
      // ---- the staged blocks of this hop have landed (also orders the v <- v2 copy of the previous hop) ------
      mbar_wait_b(&s_full, h & 1u, a.err);
      cta_sync256();
      CH_TRACE(1, 0)
      if (tid == 0) { CH_GSTAMP(0, j, t) }
      double x[32];
      double colv = 0.0;
      if (grp == 0) {
        if (from_mb) {
          // mailbox layout; its last row is beyond the bulge (zero), its last column is late
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = cq + 2 * i;
            x[i] = (r < CB - 1) ? Bn[c * CB + r] : 0.0;
          }
        } else if (has_b) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = cq + 2 * i;
            x[i] = Bn[c * SBN + r + (c & 1)];
          }
        }
        if (t == 0 && gt < CB) colv = colv_s[gt];
      } else {
#pragma unroll
        for (int q = 0; q < 17; ++q) {
          int row, col;
          x[q] = tri_slot(q, gw, lane, row, col) ? Dn[row + col * SDN] : 0.0;
        }
      }
      mbar_arrive(&s_empty);  // this thread's staged numbers are in registers
      if (t == 0) {
        // first reflector of the sweep (group A): annihilate column j below the sub-diagonal
        if (grp == 0) {
          double s = (gt >= 1 && gt < L) ? colv * colv : 0.0;
          s = warp_sum(s);
          if (lane == 0) redA[gw] = s;
          if (gt == 0) s_alpha = colv;
          group_bar(1);
          double beta, tau0, scale;
          house_scalars(s_alpha, redA[0] + redA[1] + redA[2] + redA[3], beta, tau0, scale);
          if (gt < CB) v[gt] = (gt == 0) ? 1.0 : (gt < L ? colv * scale : 0.0);
          if (gt >= 1 && gt < L) AB[(1 + gt) + (size_t)j * LDAB] = 0.0;
          if (gt == 0) {
            AB[1 + (size_t)j * LDAB] = beta;
            a.e[j] = beta;
            a.d[j] = __ldcg(AB + (size_t)j * LDAB);
            s_tau0 = tau0;
          }
        }
        cta_sync256();
        tau = s_tau0;
      }
      if (grp == 1) {
        // ================= group B: two-sided update of D = A[lo:lo+64, lo:lo+64] ===========================
        if (gt < L) a.VV[(size_t)(lo + gt) + (size_t)j * n] = v[gt];
        if (gt == 0) a.TAU[t + (size_t)j * a.maxhops] = tau;
#pragma unroll
        for (int q = 0; q < 17; ++q) {
          int row, col;
          if (tri_slot(q, gw, lane, row, col)) {
            Ds[row][col] = x[q];
            Ds[col][row] = x[q];
          }
        }
        group_bar(2);
        {
          // w = tau D v: half of the columns per thread, row r
          double sa[4] = {0.0, 0.0, 0.0, 0.0};
          const int c0 = cq * 32;
#pragma unroll
          for (int c = 0; c < 31; ++c) sa[c & 3] = fma(Ds[r][c0 + c], v[c0 + c], sa[c & 3]);
          // late: the corner D[63][63] is the last term of row 63, so everything above runs before the hand-off
          // has to be there
          double dl = Ds[r][c0 + 31];
          if (late && gt == 2 * CB - 1) {
            dl = ll_wait(ll_in, tag_in, a.err);
            Ds[CB - 1][CB - 1] = dl;
          }
          sa[3] = fma(dl, v[c0 + 31], sa[3]);
          pD[cq][r] = (sa[0] + sa[1]) + (sa[2] + sa[3]);
        }
        CH_TRACE(5, 255)
        group_bar(2);
        if (late && gt == 0) x[16] = Ds[CB - 1][CB - 1];  // slot 16 of thread 0 is that corner
        {
          const double wr0 = (gt < CB) ? tau * (pD[0][gt] + pD[1][gt]) : 0.0;
          if (gt < CB) w[gt] = wr0;
          double s = (gt < CB) ? wr0 * v[gt] : 0.0;
          s = warp_sum(s);
          if (lane == 0) redB[gw] = s;
        }
        group_bar(2);
        {
          const double al = -0.5 * tau * (redB[0] + redB[1] + redB[2] + redB[3]);
          if (gt < CB) w2[gt] = fma(al, v[gt], w[gt]);  // w + al v
          group_bar(2);
#pragma unroll
          for (int q = 0; q < 17; ++q) {
            int row, col;
            if (tri_slot(q, gw, lane, row, col)) {
              const double val = x[q] - v[row] * w2[col] - w2[row] * v[col];
              if (q < 2 && emit && col == 0)
                ll_store(ll_out + 2 * row, val, tag_out);  // first column: handed to sweep j+1, not stored
              else if (to_mb && col > 0)
                mbw[MB_B + (row - 1) + (col - 1) * SDN] = val;
              else
                pd0[row + col * (LDAB - 1)] = val;
            }
          }
        }
        CH_TRACE(6, 128)
      } else if (has_b) {
        // ================= group A: Bk = A[lo+64:lo+128, lo:lo+64] <- H2 (Bk H) and the next reflector ==========
        {
          double sa[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int i = 0; i < 31; ++i) sa[i & 3] = fma(x[i], v[cq + 2 * i], sa[i & 3]);
          // late: column 63 of the block = first column of the diagonal block of hop (j-1, t+1), rows 1..63, and the
          // head of that hop's bulge column; it enters the sums last, so the wait sits as deep in the hop as it can
          if (late && cq == 1) x[31] = ll_wait(ll_in + 2 * (r + 1), tag_in, a.err);
          CH_TRACE(2, 64)
          sa[3] = fma(x[31], v[cq + 62], sa[3]);
          pA[cq][r] = (sa[0] + sa[1]) + (sa[2] + sa[3]);
        }
        group_bar(1);
        {
          const double ur = tau * (pA[0][r] + pA[1][r]);
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = fma(-ur, v[cq + 2 * i], x[i]);
        }
        if (more) {
          // reflector from the first column of the bulge (held by the cq == 0 half: x[0] = Bk[r][0])
          double s = (cq == 0 && r >= 1) ? x[0] * x[0] : 0.0;
          s = warp_sum(s);
          if (lane == 0) redA[gw] = s;
          if (gt == 0) s_alpha = x[0];
          group_bar(1);
          double beta, tau2, scale;
          house_scalars(s_alpha, redA[0] + redA[1] + redA[2] + redA[3], beta, tau2, scale);
          if (cq == 0) {
            v2[r] = (r == 0) ? 1.0 : x[0] * scale;
            x[0] = (r == 0) ? beta : 0.0;
          }
          if (gt == 0) s_tau2 = tau2;
          if (emit && gt == 0) ll_store(ll_out + 2 * CB, beta, tag_out);  // head of the new bulge column
          CH_TRACE(3, 0)
#pragma unroll
          for (int i = 0; i < 32; ++i) Bs[r][cq + 2 * i] = x[i];
          group_bar(1);
          {
            // wa[c] = tau2 v2' Bk[:, c]: half of the rows per thread, column r (used as the column index here)
            double sa[4] = {0.0, 0.0, 0.0, 0.0};
            const int q0 = cq * 32;
#pragma unroll
            for (int q = 0; q < 32; ++q) sa[q & 3] = fma(v2[q0 + q], Bs[q0 + q][r], sa[q & 3]);
            pA[cq][r] = (sa[0] + sa[1]) + (sa[2] + sa[3]);
          }
          group_bar(1);
          if (gt < CB) wa[gt] = tau2 * (pA[0][gt] + pA[1][gt]);
          group_bar(1);
          {
            const double v2r = v2[r];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = cq + 2 * i;
              if (c >= 1) x[i] = fma(-v2r, wa[c], x[i]);
            }
          }
        } else if (emit && gt == 0) {
          ll_store(ll_out + 2 * CB, x[0], tag_out);  // no further reflector: the element as the right-update left it
        }
        if (to_mb) {
          // row 0 is the last row of the next sweep's diagonal block; column 0 (the annihilated bulge column) is
          // never read again - except its head at hop 0, which is the last entry of the next sweep's column
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = cq + 2 * i;
            if (c == 0) continue;
            if (r > 0)
              mbw[(c - 1) * CB + (r - 1)] = x[i];
            else
              mbw[MB_B + (CB - 1) + (c - 1) * SDN] = x[i];
          }
          if (gt == 0 && !emit) pb[0] = x[0];
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i == 0 && emit && gt == 0) continue;  // handed over above
            pb[(cq + 2 * i) * (LDAB - 1)] = x[i];
          }
        }
        CH_TRACE(4, 0)
      } else if (emit && gt == 0) {
        ll_store(ll_out + 2 * CB, 0.0, tag_out);  // no block below: the head sweep j+1 waits for is a padding zero
      }
      // this thread's stores of the hop are issued: the transfer warps publish the hop once everybody is here
      mbar_arrive(&s_done);
      if (tid == 0) { CH_GSTAMP(1, j, t) }
      ++h;
      cta_sync256();
      CH_TRACE(10, 0)
      if (more) {
        if (tid < CB) v[tid] = v2[tid];
        tau = s_tau2;
      }
      if (!more) break;
    }
    cta_sync256();
  }
#undef CH_TRACE
}

// d[n-2], d[n-1], e[n-2] are never touched by a sweep with a reflector: read them off the band at the end
__global__ void chase_tail_kernel(const double* __restrict__ AB, int n, double* d, double* e) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int j = max(0, n - 2); j < n; ++j) {
      d[j] = AB[(size_t)j * LDAB];
      if (j + 1 < n) e[j] = AB[1 + (size_t)j * LDAB];
    }
  }
}

int sb2st(bk_ctx* ctx, double* AB, int n, double* d, double* e, double* VV, double* TAU, int maxhops) {
  DevBuf<int> prog, err;
  BK_TRY(prog.alloc(n));
  BK_TRY(err.alloc(1));
  BK_CUDA(cudaMemsetAsync(prog.p, 0, sizeof(int) * n, ctx->stream));
  BK_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), ctx->stream));
  ChaseArgs a;
  a.AB = AB;
  a.n = n;
  a.VV = VV;
  a.TAU = TAU;
  a.maxhops = maxhops;
  a.prog = prog.p;
  a.d = d;
  a.e = e;
  a.err = err.p;
  a.ll = nullptr;
  a.trace = nullptr;
  a.trace_j0 = 0;
  a.trace_t0 = 0;
  DevBuf<long long> prof, trace;
  a.prof = nullptr;
  // warp-specialised kernel with the early hand-off (default); BK_CHASE_LL=0 selects the completion-flag kernel
  // (same bits; read per call: the tests compare both in one process)
  const char* ll_env = getenv("BK_CHASE_LL");
  const bool use_ws = !(ll_env && atoi(ll_env) == 0);
  if (!use_ws && getenv("BK_CHASE_PROF")) {
    BK_TRY(prof.alloc(16));
    BK_CUDA(cudaMemsetAsync(prof.p, 0, 16 * sizeof(long long), ctx->stream));
    a.prof = prof.p;
  }
  const size_t trace_n = (size_t)TR_SWEEPS * TR_HOPS * TR_EV + 6 * 16 * 64;  // + global-timer stamps of 16 sweeps
  if (const char* tj = getenv("BK_CHASE_TRACE")) {
    if (use_ws) {
      BK_TRY(trace.alloc(trace_n));
      BK_CUDA(cudaMemsetAsync(trace.p, 0, trace_n * sizeof(long long), ctx->stream));
      a.trace = trace.p;
      a.trace_j0 = atoi(tj);
      a.trace_t0 = getenv("BK_CHASE_TRACE_T0") ? atoi(getenv("BK_CHASE_TRACE_T0")) : 0;
    }
  }
  DevBuf<unsigned long long> ll;
  DevBuf<double> mbox;
  a.mb = nullptr;
  if (use_ws) {
    BK_TRY(mbox.alloc((size_t)2 * maxhops * MB_STRIDE));
    a.mb = mbox.p;
    const size_t words = (size_t)LL_RING * maxhops * LL_PER_HOP * 2;
    BK_TRY(ll.alloc(words));
    BK_CUDA(cudaMemsetAsync(ll.p, 0, sizeof(unsigned long long) * words, ctx->stream));  // tag 0 = never written
    a.ll = ll.p;
  }
  if (n > 2) {
    void* kargs[] = {&a};
    if (use_ws) {
      const size_t smem = sizeof(double) * (2 * CB * (CB + 1) + CB * (SBN + SDN) + CB + 2);
      void* kern = a.trace ? (void*)chase_ws_kernel<true> : (void*)chase_ws_kernel<false>;
      BK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      BK_CUDA(cudaLaunchCooperativeKernel(kern, dim3(ctx->sm_count), dim3(WS_NT), kargs, smem, ctx->stream));
    } else {
      const size_t smem = sizeof(double) * 2 * CB * (CB + 1);
      BK_CUDA(cudaFuncSetAttribute(chase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      BK_CUDA(cudaLaunchCooperativeKernel((void*)chase_kernel, dim3(ctx->sm_count), dim3(CH_NT), kargs, smem,
                                          ctx->stream));
    }
    BK_LAUNCHED(ctx);
  }
  chase_tail_kernel<<<1, 32, 0, ctx->stream>>>(AB, n, d, e);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  int herr = 0;
  BK_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (herr != 0) {
    set_error("sb2st: a hand-off between sweeps of the bulge chasing never arrived");
    return BK_ERR_NUMERIC;
  }
  if (a.trace) {
    // one line per traced hop: sweep, hop, then the clock readings relative to the first hop of the sweep
    std::vector<long long> h(trace_n);
    BK_CUDA(cudaMemcpyAsync(h.data(), trace.p, trace_n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int sj = 0; sj < TR_SWEEPS; ++sj)
      for (int th = 0; th < TR_HOPS; ++th) {
        const long long* e0 = h.data() + ((size_t)sj * TR_HOPS) * TR_EV;
        const long long* ev = h.data() + ((size_t)sj * TR_HOPS + th) * TR_EV;
        if (ev[0] == 0) continue;
        // sweeps j0, j0 + CTAs, ... run on the same CTA: one clock; the second number is the start of the sweep
        // relative to the first traced sweep (CTAs x the lag between neighbouring sweeps)
        fprintf(stderr, "[chase trace] %d %d | %lld |", a.trace_j0 + sj * ctx->sm_count, th + a.trace_t0, e0[0] - h[0]);
        for (int q = 0; q < TR_EV; ++q) fprintf(stderr, " %lld", ev[q] ? ev[q] - e0[0] : -1LL);
        fprintf(stderr, "\n");
      }
  }
  if (a.trace) {
    // global timer (ns) at the start of hops 0..63 of 16 consecutive sweeps: the lag between neighbours, hop by hop
    std::vector<long long> h(trace_n);
    BK_CUDA(cudaMemcpyAsync(h.data(), trace.p, trace_n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    const long long* g = h.data() + (size_t)TR_SWEEPS * TR_HOPS * TR_EV;
    auto G = [&](int kind, int sj, int th) { return g[((size_t)kind * 16 + sj) * 64 + th]; };
    for (int th : {0, 1, 2, 8, 32}) {
      fprintf(stderr, "[chase lag ns] hop %d: start(j)-start(j-1):", th);
      for (int sj = 1; sj < 16; ++sj) fprintf(stderr, " %lld", G(0, sj, th) - G(0, sj - 1, th));
      fprintf(stderr, "\n[chase lag ns] hop %d: per sweep j: start -> stores issued -> flag released | sweep j+1: staging "
                      "commanded -> hop starts:", th);
      for (int sj = 1; sj < 8; ++sj)
        fprintf(stderr, " [%lld %lld | %lld %lld]", G(1, sj, th) - G(0, sj, th), G(2, sj, th) - G(1, sj, th),
                G(3, sj + 1, th) - G(2, sj, th), G(0, sj + 1, th) - G(3, sj + 1, th));
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "[chase lag ns] hop 0 of sweep j+1: flag of (j,0) released -> first poll that sees it (polls so far) ; "
                    "transfer warps waiting since (relative to the release):");
    for (int sj = 1; sj < 8; ++sj)
      fprintf(stderr, " [%lld (%lld) ; %lld]", G(5, sj + 1, 0) - G(2, sj, 0), G(5, sj + 1, 1), G(4, sj + 1, 0) - G(2, sj, 0));
    fprintf(stderr, "\n");
  }
  if (a.prof) {
    long long h[16];
    BK_CUDA(cudaMemcpyAsync(h, prof.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    const double hops = (double)std::max(1LL, h[0]), fa = (double)std::max(1LL, h[5]);
    fprintf(stderr, "[chase prof, CTA 0] hops %lld: cycles per hop: wait %.0f, group A %.0f, group B %.0f, total %.0f | "
            "group A chain: loads+matvec %.0f, rank-1+norm %.0f, scalars+staging %.0f, column dots %.0f, apply+stores %.0f\n",
            h[0], h[1] / hops, h[2] / hops, h[3] / hops, h[4] / hops, h[6] / fa, h[7] / fa, h[8] / fa, h[9] / fa, h[10] / fa);
  }
  return BK_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Q2 back-transformation on Zt (k x n, row r of Z = column r of Zt, contiguous)
// ---------------------------------------------------------------------------------------------------------
static constexpr int Q2_KC = 160;           // columns of Z per launch (widest variant of the kernel)
// per variant: 2 WC compute threads, two per column (upper / lower half of the window), + one warp that only polls /
// publishes the pipeline flags
static constexpr int Q2_R = 4;              // sweeps applied per pass over the window
static constexpr int Q2_HALF = 34;          // window rows per thread: 2 * 34 >= 64 + Q2_R - 1
static constexpr int Q2_ROWS = 112;         // buffer rows: the window slides through the buffer, no wrap-around
static constexpr int Q2_PTOP = Q2_ROWS - 2 * Q2_HALF - 1;  // highest buffer row for the top of a window (43)

struct Q2Args {
  double* Zt;     // kc x n (ld = ldzt)
  long long ldzt;
  int n, kc;
  const double* VV;
  const double* TAU;
  int maxhops;
  int* done;      // per hop index: lowest sweep already applied (INT_MAX initially)
  long long* prof;  // optional (BK_Q2_PROF), CTA 0: [0] passes, [1] compute cycles, [2] barrier wait, [3] flag wait
};

// One CTA owns hop index t for all sweeps (j = jmax(t) .. 0).  The reflector (j, t) acts on rows
// [j+1+64t, j+65+64t): the window slides up one row per sweep.  The window lives in a shared-memory buffer
// (row lo of the current sweep at buffer row p, p decreasing; when p reaches the top the window is stored
// back at the bottom, so every access is a compile-time offset from one moving base).
//   fast path (full 64-row window): per pass the 63 resident rows plus 4 new rows go to registers, 4 consecutive
//   reflectors are applied and 4 rows retire.  TWO threads share a column (rows 0..33 / 34..67 of the 68-row
//   pass window, lanes l and l+16 of a warp): half the registers per thread, twice the warps to hide latency;
//   both halves run the same instruction stream on zero-padded copies of the reflectors and combine their dot
//   products with one shuffle.
//   slow path (window still growing near the bottom of the matrix, last < 4 sweeps): one sweep per step from
//   shared memory, one thread per column.
// Rows retired by hop index t-1 are the rows entering the window of hop index t: done[t-1] is the only
// synchronisation between CTAs (acquire/release, one warp of each CTA does nothing else).
// WC = columns per launch (row stride of the window buffer).  160 for wide blocks (one CTA per SM, one launch per
// 160 columns); 80 / 48 for the narrow blocks a rank gets when the back-transformation is split over 4 / 8 GPUs: less
// shared memory and fewer threads per CTA let 2 / 3 CTAs share an SM, so that every hop index has its own CTA and
// the pipeline over hop indices is not run twice per CTA.
template <int WC>
__global__ void __launch_bounds__(2 * WC + 32, (WC <= 48) ? 3 : (WC <= 80 ? 2 : 1)) q2_apply_kernel(Q2Args a) {
  constexpr int Q2_KC = WC, Q2_CT = 2 * WC, Q2_NT = Q2_CT + 32;
  extern __shared__ __align__(16) double win[];  // Q2_ROWS x Q2_KC
  __shared__ __align__(16) double vsx[2][Q2_R][2][Q2_HALF];  // zero-padded reflectors of a pass, per half
  __shared__ double taus[2][Q2_R];
  __shared__ double vs0[CB];  // slow path
  __shared__ double tau0;
  const int n = a.n, kc = a.kc, tid = threadIdx.x;
  const bool flagger = (tid == Q2_CT);  // lane 0 of the flag warp
  const bool comp = tid < Q2_CT;
  const int col = comp ? ((tid >> 5) * 16 + (tid & 15)) : 0;
  const int h = (tid >> 4) & 1;
  const bool act = comp && col < kc;  // compute thread with a real column (global accesses)
  const int nhop = (n - 3) / CB + 1;  // hop indices t with jmax(t) = n-3-t*CB >= 0
  const double* VV = a.VV;
  double* colp = win + col;  // this thread's column: buffer row q at colp[q * Q2_KC]
  long long tk0 = 0;
  if (a.prof) tk0 = clock64();
  for (int t = blockIdx.x; t < nhop; t += gridDim.x) {
    const int jmax = n - 3 - t * CB;
    const int* pdone = a.done + t - 1;
    int* mydone = a.done + t;
    int j = jmax;
    int p = Q2_PTOP;  // buffer row of the top row (lo) of the window of sweep j
    // rows past the end of the matrix (windows clipped at the bottom) are zero rows of the buffer
    for (int i = tid; i < Q2_ROWS * Q2_KC; i += Q2_NT) win[i] = 0.0;
    __syncthreads();
    // move the resident rows buf[p+1 .. p+cnt] to buf[PTOP+1 .. PTOP+cnt] (descending: the ranges may overlap)
    auto rebase_slow = [&](int cnt) {
      if (comp && h == 0)
        for (int i = cnt; i >= 1; --i) colp[(Q2_PTOP + i) * Q2_KC] = colp[(p + i) * Q2_KC];
      p = Q2_PTOP;
    };
    // One sweep, window read from shared memory with run-time length (growing window / tail sweeps).
    auto slow_sweep = [&](int js, bool first) {
      long long ts0 = 0;
      if (a.prof) ts0 = clock64();
      const int lo = js + 1 + t * CB, hi = min(n, lo + CB), L = hi - lo;
      if (p < 0) rebase_slow(L - 1);
      if (t > 0 && flagger) {
        while (ld_acquire_i32(pdone) > js + 1) {
        }
      }
      __syncthreads();
      if (act && h == 0) {
        if (first) {
          for (int i = 0; i < L; ++i) colp[(p + i) * Q2_KC] = __ldcg(a.Zt + col + (size_t)(lo + i) * a.ldzt);
        } else {
          colp[p * Q2_KC] = __ldcg(a.Zt + col + (size_t)lo * a.ldzt);
        }
      }
      if (tid < L) vs0[tid] = VV[(size_t)(lo + tid) + (size_t)js * n];
      if (tid == 0) tau0 = a.TAU[t + (size_t)js * a.maxhops];
      __syncthreads();
      if (act && h == 0) {
        double s0 = 0.0, s1 = 0.0;
        int i = 0;
        for (; i + 1 < L; i += 2) {
          s0 = fma(vs0[i], colp[(p + i) * Q2_KC], s0);
          s1 = fma(vs0[i + 1], colp[(p + i + 1) * Q2_KC], s1);
        }
        if (i < L) s0 = fma(vs0[i], colp[(p + i) * Q2_KC], s0);
        const double wv = tau0 * (s0 + s1);
        for (i = 0; i < L; ++i) colp[(p + i) * Q2_KC] -= vs0[i] * wv;
        if (lo + CB <= n) a.Zt[col + (size_t)(lo + CB - 1) * a.ldzt] = colp[(p + CB - 1) * Q2_KC];
      }
      __syncthreads();
      if (flagger) st_release_i32(mydone, js);
      --p;
      if (a.prof && blockIdx.x == 0 && tid == 0) {
        a.prof[4] += 1;
        a.prof[5] += clock64() - ts0;
      }
    };
    // ---- main phase: Q2_R sweeps per pass (windows clipped by the end of the matrix run on zero rows) ------
    const bool fast_phase = (j >= Q2_R - 1);
    if (fast_phase) {
      // loads for a pass starting at sweep jp: its 4 new rows (two per half-thread) and its 4 reflectors
      constexpr int NVE = Q2_R * 2 * Q2_HALF;             // reflector entries of a pass
      constexpr int NV = (NVE + Q2_NT - 1) / Q2_NT;        // per thread (1 for the wide variant)
      double pr[2], vreg[NV], treg = 0.0;
#pragma unroll
      for (int q = 0; q < NV; ++q) vreg[q] = 0.0;
      auto issue_pass_loads = [&](int jp) {
        const int lo = jp + 1 + t * CB;
        pr[0] = act ? __ldcg(a.Zt + col + (size_t)(lo - 2 * h) * a.ldzt) : 0.0;
        pr[1] = act ? __ldcg(a.Zt + col + (size_t)(lo - 2 * h - 1) * a.ldzt) : 0.0;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          const int e = tid + q * Q2_NT;
          if (e < NVE) {
            // zero-padded reflector s on the pass window: window index i <-> row lo-3+i, reflector rows lo-s ..
            const int s = e / (2 * Q2_HALF), i = e % (2 * Q2_HALF);
            const int vi = i - (Q2_R - 1 - s);
            vreg[q] = (vi >= 0 && vi < CB && lo - s + vi < n) ? VV[(size_t)(lo - s + vi) + (size_t)(jp - s) * n] : 0.0;
          }
        }
        if (tid < Q2_R) treg = a.TAU[t + (size_t)(jp - tid) * a.maxhops];
      };
      auto stage_pass = [&](int buf, int pn) {  // prefetched rows -> window buffer, reflectors -> vsx[buf]
        if (comp) {
          colp[(pn - 2 * h) * Q2_KC] = pr[0];
          colp[(pn - 2 * h - 1) * Q2_KC] = pr[1];
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          const int e = tid + q * Q2_NT;
          if (e < NVE) (&vsx[buf][0][0][0])[e] = vreg[q];
        }
        if (tid < Q2_R) taus[buf][tid] = treg;
      };
      // a pass at sweep jp needs hop index t-1 to have applied sweep jp-2
      auto wait_pass = [&](int jp) {
        if (t > 0 && flagger) {
          while (ld_acquire_i32(pdone) > jp - 2) {
          }
        }
      };
      wait_pass(j);
      if (j - Q2_R >= Q2_R - 1) wait_pass(j - Q2_R);
      __syncthreads();
      {
        // rows already below the top of the first window (the last rows of the matrix)
        const int lo = j + 1 + t * CB;
        if (act && h == 0)
          for (int r = lo + 1; r < min(n, lo + CB); ++r)
            colp[(p + r - lo) * Q2_KC] = __ldcg(a.Zt + col + (size_t)r * a.ldzt);
      }
      issue_pass_loads(j);
      int cur = 0;
      stage_pass(cur, p);
      __syncthreads();
      while (j >= Q2_R - 1) {
        long long tq0 = 0, tq1 = 0, tq2 = 0;
        if (a.prof) tq0 = clock64();
        const int lo = j + 1 + t * CB;
        const bool have_next = (j - Q2_R >= Q2_R - 1);
        if (have_next) issue_pass_loads(j - Q2_R);  // verified before the previous barrier
        const int newp = (p - Q2_R >= 2 * Q2_R - 1) ? p - Q2_R : Q2_PTOP;
        if (comp) {
          // z[li] <-> pass-window index 34 h + li <-> row lo-3+34h+li <-> buffer row p-3+34h+li
          const double* rb = colp + (p - (Q2_R - 1) + Q2_HALF * h) * Q2_KC;
          double z[Q2_HALF];
#pragma unroll
          for (int i = 0; i < Q2_HALF; ++i) z[i] = rb[i * Q2_KC];
#pragma unroll 1
          for (int s = 0; s < Q2_R; ++s) {  // the register indices do not depend on s: keep one copy of the body
            const double2* v2p = reinterpret_cast<const double2*>(&vsx[cur][s][h][0]);
            double sa[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int i = 0; i < Q2_HALF; i += 2) {
              const double2 va = v2p[i / 2];
              sa[i & 2] = fma(va.x, z[i], sa[i & 2]);
              sa[(i & 2) + 1] = fma(va.y, z[i + 1], sa[(i & 2) + 1]);
            }
            double dot = (sa[0] + sa[1]) + (sa[2] + sa[3]);
            dot += __shfl_xor_sync(0xffffffffu, dot, 16);
            const double wv = taus[cur][s] * dot;
#pragma unroll
            for (int i = 0; i < Q2_HALF; i += 2) {
              const double2 va = v2p[i / 2];
              z[i] = fma(-va.x, wv, z[i]);
              z[i + 1] = fma(-va.y, wv, z[i + 1]);
            }
          }
          // the 4 bottom rows (window indices 63..66 = local 29..32 of the lower half) retire
          if (act && h == 1) {
            double* zt = a.Zt + col + (size_t)(lo - (Q2_R - 1) + Q2_HALF) * a.ldzt;
#pragma unroll
            for (int i = CB - 1 - Q2_HALF; i < CB + Q2_R - 1 - Q2_HALF; ++i)
              if (lo - (Q2_R - 1) + Q2_HALF + i < n) zt[(size_t)i * a.ldzt] = z[i];
          }
          // window index i goes to buffer row newp+1+i (row lo-3 becomes row lo'+1 of the next pass)
          double* wb = colp + (newp + 1 + Q2_HALF * h) * Q2_KC;
#pragma unroll
          for (int i = 0; i < Q2_HALF; ++i) wb[i * Q2_KC] = z[i];
        }
        p = newp;
        if (have_next) {
          stage_pass(cur ^ 1, p);
          if (a.prof) tq1 = clock64();
          if (j - 2 * Q2_R >= Q2_R - 1) wait_pass(j - 2 * Q2_R);
        }
        if (a.prof) tq2 = clock64();
        __syncthreads();
        if (a.prof && blockIdx.x == 0) {
          const long long tq3 = clock64();
          if (tid == 0) {
            a.prof[0] += 1;
            a.prof[1] += tq2 - tq0;
            a.prof[2] += tq3 - tq2;
          }
          if (flagger && have_next) a.prof[3] += tq2 - tq1;
        }
        if (flagger) st_release_i32(mydone, j - (Q2_R - 1));
        cur ^= 1;
        j -= Q2_R;
      }
    }
    // ---- the last (< Q2_R) sweeps, or every sweep of a hop index with fewer than Q2_R sweeps ------------------
    for (; j >= 0; --j) slow_sweep(j, !fast_phase && j == jmax);
    // flush what is left of the window (nobody inside this kernel waits for these rows): after the last
    // sweep p points one above the row of lo0 = 1 + 64 t
    {
      const int lo0 = 1 + t * CB;
      const int cnt = ((lo0 + CB <= n) ? lo0 + CB - 1 : n) - lo0;
      if (act && h == 0)
        for (int i = 0; i < cnt; ++i) a.Zt[col + (size_t)(lo0 + i) * a.ldzt] = colp[(p + 1 + i) * Q2_KC];
    }
    __syncthreads();
  }
  if (a.prof && blockIdx.x == 0 && tid == 0) a.prof[6] = clock64() - tk0;
}

// ---------------------------------------------------------------------------------------------------------
// Q2 back-transformation on the FP64 tensor path (q2_mma_kernel): 8 sweeps of a hop index as one block reflector
// Executable specification: tests/q2_mma_prototype.py (lane-level fragments, staged pass buffers, pass lattice).
// ---------------------------------------------------------------------------------------------------------
// The kernel above applies one reflector after the other with one (half) column per thread: every thread reads
// every reflector entry from shared memory, and that broadcast traffic bounds it (ncu, profiles/r02g_q2_ncu_key.txt:
// 55 % of the warp samples at barriers, shared-memory data pipe 70 % busy in a CTA that computes, FP64 pipe 14 %).
// Here a pass of hop index t covers the 8 sweeps jp .. jp-7 (jp = jmax(t) - 8 i; sweeps below 0 are identity
// reflectors): their union of rows is the 72-row window  row(w) = lo - 7 + w,  lo = jp + 1 + 64 t,  reflector s on
// w = 7-s .. 70-s.  With V (72 x 8, zero-padded) and the compact-WY factor T (8 x 8 upper triangular,
// q2_tfac8_kernel) the pass is  Z_w <- Z_w - V T' (V' Z_w),  three small products that run as DMMA.8x8x4 on
// TRANSPOSED tiles so that the window never changes registers:
//   a warp owns 16 columns (2 tiles of 8); lane l of tile c holds z[c][2j+e] = Z(row(8j + 2(l%4) + e), col 8c + l/4),
//   j = 0..8 - this is at once the A fragment of k-step q = 2j+e of  D' = Z_w' V  (the k index runs over the rows in
//   the order 8j + 2k + e, the B fragments are staged in the same order) and the C fragment of row tile j of
//   Z_w' -= W' V'  (W' = D' T, whose C fragment is again the A fragment of the next product under the same
//   permutation of the reflector index).  No shuffles, no transposition; the reflectors are read from shared memory
//   once per 16 columns and 8 sweeps (40 x 256 B per warp) instead of once per column and sweep.
//   Between passes the window moves down by exactly one row tile (registers renamed), tile 8 (rows lo+57 .. lo+64)
//   retires to global memory - it is tile 0 of hop index t+1 at the pass with the same jp (jmax(t+1) = jmax(t) - 64:
//   the passes of all hop indices lie on one lattice) - and tile 0 of the next pass comes in from global memory.
//   done[t] = jp after the pass: hop index t+1 may run the pass with that jp.  Rows below 1 + 64 t never belong to hop
//   index t (only the zero-padded last pass sees them): neither loaded nor stored.
// The fragments of a pass (B operands of the three products, in fragment order: 18 + 18 + 2 warp-wide 8-byte loads)
// arrive by cp.async, zero-filled where a reflector has no entry, three passes deep.
static constexpr int Q2M_PB = 18 * 32 + 9 * 64 + 64;  // doubles per staged pass: dots, update, triangular factor
static constexpr int Q2M_NB = 3;                      // staged passes in flight

struct Q2MArgs {
  double* Zt;     // kc x n (ld = ldzt)
  long long ldzt;
  int n, kc;
  const double* VV;
  const double* TAU;
  int maxhops;
  int* done;            // per hop index: jp of the last finished pass (INT_MAX initially)
  const double* TF;     // per (hop index, pass): T in B-fragment order (64 doubles)
  const long long* offs;  // first pass of hop index t in TF
};

// T of the 8 reflectors of (hop index t, pass i), forward column-wise recurrence of the compact-WY form:
// T[s][s] = tau_s, T[0:s, s] = -tau_s T[0:s, 0:s] (V[:, 0:s]' v_s).  One warp per pass; output in the order the
// kernel reads it as B fragments of W' = D' T: entry 32 e + l = T[2 (l%4) + e][l/4].
__global__ void __launch_bounds__(256) q2_tfac8_kernel(const double* __restrict__ VV, const double* __restrict__ TAU,
                                                       int maxhops, int n, const long long* __restrict__ offs,
                                                       double* __restrict__ TF) {
  __shared__ double tsm[8][64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t = blockIdx.y, jmax = n - 3 - t * CB, np = jmax / 8 + 1;
  const int i = blockIdx.x * 8 + wid;
  if (i >= np) return;
  const int jp = jmax - 8 * i, lo = jp + 1 + t * CB;
  double v[8][3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int w = lane + 32 * q, row = lo - 7 + w;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int vi = w - (7 - s);
      v[s][q] = (w < 72 && jp - s >= 0 && vi >= 0 && vi < CB && row < n) ? VV[(size_t)row + (size_t)(jp - s) * n] : 0.0;
    }
  }
  double tau[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) tau[s] = (jp - s >= 0) ? TAU[t + (size_t)(jp - s) * maxhops] : 0.0;
  double T[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int s = 0; s < 8; ++s) T[r][s] = 0.0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    double g[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < s) {
        double x = fma(v[r][0], v[s][0], fma(v[r][1], v[s][1], v[r][2] * v[s][2]));
        g[r] = warp_sum(x);
      }
    }
    T[s][s] = tau[s];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < s) {
        double x = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c >= r && c < s) x = fma(T[r][c], g[c], x);
        T[r][s] = -tau[s] * x;
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int s = 0; s < 8; ++s) tsm[wid][r * 8 + s] = T[r][s];
  }
  __syncwarp();
  double* o = TF + (size_t)(offs[t] + i) * 64;
#pragma unroll
  for (int e = 0; e < 2; ++e) o[32 * e + lane] = tsm[wid][(2 * (lane & 3) + e) * 8 + (lane >> 2)];
}

template <int NW>  // compute warps of 16 columns each, plus one warp for the flags
__global__ void __launch_bounds__(32 * NW + 32, (NW <= 3) ? 3 : (NW <= 5 ? 2 : 1)) q2_mma_kernel(Q2MArgs a) {
  constexpr int NT = 2;
  constexpr int NCT = 32 * NW;  // compute threads
  extern __shared__ __align__(16) double q2m_sm[];  // Q2M_NB staged passes
  const int n = a.n, kc = a.kc, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, l4 = lane >> 2, lm = lane & 3;
  const bool comp = warp < NW;
  const bool flagger = (tid == NCT);
  const int nhop = (n - 3) / CB + 1;
  const double* const VV = a.VV;
  int colg[NT];
  bool act[NT];
#pragma unroll
  for (int c = 0; c < NT; ++c) {
    colg[c] = warp * (8 * NT) + 8 * c + l4;
    act[c] = comp && colg[c] < kc;
  }
  for (int t = blockIdx.x; t < nhop; t += gridDim.x) {
    const int jmax = n - 3 - t * CB, np = jmax / 8 + 1;
    const int rmin = 1 + t * CB;  // first row that ever belongs to this hop index
    const int* pdone = a.done + t - 1;
    int* mydone = a.done + t;
    const double* const tf = a.TF + (size_t)a.offs[t] * 64;
    // cp.async of the fragments of pass i into its ring slot (compute threads; everybody commits)
    auto stage = [&](int i) {
      if (comp && i < np) {
        const int jp = jmax - 8 * i, lo = jp + 1 + t * CB;
        double* buf = q2m_sm + (i % Q2M_NB) * Q2M_PB;
        for (int e = tid; e < Q2M_PB; e += NCT) {
          const double* src = VV;
          int bytes = 0;
          if (e < 1152) {
            int s, w;
            if (e < 576) {  // dots: k-step q = 2j+ee, lane l: V(w = 8j + 2(l%4) + ee, s = l/4)
              const int q = e >> 5, l = e & 31;
              s = l >> 2;
              w = 8 * (q >> 1) + 2 * (l & 3) + (q & 1);
            } else {  // update: row tile j, k-step ee, lane l: V(w = 8j + l/4, s = 2(l%4) + ee)
              const int e2 = e - 576, l = e2 & 31;
              s = 2 * (l & 3) + ((e2 >> 5) & 1);
              w = 8 * (e2 >> 6) + (l >> 2);
            }
            const int vi = w - (7 - s), row = lo - 7 + w, sw = jp - s;
            if (sw >= 0 && vi >= 0 && vi < CB && row < n) {
              src = VV + (size_t)row + (size_t)sw * n;
              bytes = 8;
            }
          } else {
            src = tf + (size_t)i * 64 + (e - 1152);
            bytes = 8;
          }
          cp_async8(buf + e, src, bytes);
        }
      }
      cp_async_commit();
    };
    auto wait_pass = [&](int i) {  // hop index t-1 has finished the pass with the same jp
      if (t > 0 && flagger && i < np) {
        const int jp = jmax - 8 * i;
        while (ld_acquire_i32(pdone) > jp) {
        }
      }
    };
    wait_pass(0);
    wait_pass(1);
    stage(0);
    stage(1);
    __syncthreads();  // passes 0 and 1 of hop index t-1 are finished: their retired rows may be read
    double z[NT][18];
    {
      const int lo = jmax + 1 + t * CB;
#pragma unroll
      for (int c = 0; c < NT; ++c)
#pragma unroll
        for (int j = 0; j < 9; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int row = lo - 7 + 8 * j + 2 * lm + e;
            z[c][2 * j + e] = (act[c] && row >= rmin && row < n) ? __ldcg(a.Zt + colg[c] + (size_t)row * a.ldzt) : 0.0;
          }
    }
    cp_async_wait<1>();
    __syncthreads();  // the fragments of pass 0 are visible
    for (int i = 0; i < np; ++i) {
      const int jp = jmax - 8 * i, lo = jp + 1 + t * CB;
      const bool last = (i == np - 1);
      stage(i + 2);  // its slot was read in pass i-1
      // tile 0 of the next pass: rows lo-15 .. lo-8 (retired by hop index t-1 in its pass jp-8, verified before the
      // barrier that ended pass i-1)
      double nr[NT][2];
#pragma unroll
      for (int c = 0; c < NT; ++c)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int row = lo - 15 + 2 * lm + e;
          nr[c][e] = (!last && act[c] && row >= rmin) ? __ldcg(a.Zt + colg[c] + (size_t)row * a.ldzt) : 0.0;
        }
      if (comp) {
        const double* buf = q2m_sm + (i % Q2M_NB) * Q2M_PB;
        // D' = Z_w' V: 18 k-steps, two accumulator pairs per tile (even / odd k-steps)
        double d[NT][2][2];
#pragma unroll
        for (int c = 0; c < NT; ++c) d[c][0][0] = d[c][0][1] = d[c][1][0] = d[c][1][1] = 0.0;
#pragma unroll
        for (int q = 0; q < 18; ++q) {
          const double bv = buf[q * 32 + lane];
#pragma unroll
          for (int c = 0; c < NT; ++c) dmma884(d[c][q & 1][0], d[c][q & 1][1], z[c][q], bv);
        }
        // W' = D' T, negated
        const double tb0 = buf[1152 + lane], tb1 = buf[1184 + lane];
        double wn[NT][2];
#pragma unroll
        for (int c = 0; c < NT; ++c) {
          double w0 = 0.0, w1 = 0.0;
          dmma884(w0, w1, d[c][0][0] + d[c][1][0], tb0);
          dmma884(w0, w1, d[c][0][1] + d[c][1][1], tb1);
          wn[c][0] = -w0;
          wn[c][1] = -w1;
        }
        // Z_w' -= W' V', row tile by row tile
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const double bu0 = buf[576 + j * 64 + lane], bu1 = buf[576 + j * 64 + 32 + lane];
#pragma unroll
          for (int c = 0; c < NT; ++c) {
            dmma884(z[c][2 * j], z[c][2 * j + 1], wn[c][0], bu0);
            dmma884(z[c][2 * j], z[c][2 * j + 1], wn[c][1], bu1);
          }
        }
        // tile 8 retires (tile 0 of hop index t+1 in its pass jp)
#pragma unroll
        for (int c = 0; c < NT; ++c)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int row = lo + 57 + 2 * lm + e;
            if (act[c] && row < n) a.Zt[colg[c] + (size_t)row * a.ldzt] = z[c][16 + e];
          }
        if (last) {
          // the rest of the window is final as well
#pragma unroll
          for (int c = 0; c < NT; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int row = lo - 7 + 8 * j + 2 * lm + e;
                if (act[c] && row >= rmin && row < n) a.Zt[colg[c] + (size_t)row * a.ldzt] = z[c][2 * j + e];
              }
        } else {
#pragma unroll
          for (int c = 0; c < NT; ++c) {
#pragma unroll
            for (int j = 8; j >= 1; --j) {
              z[c][2 * j] = z[c][2 * j - 2];
              z[c][2 * j + 1] = z[c][2 * j - 1];
            }
            z[c][0] = nr[c][0];
            z[c][1] = nr[c][1];
          }
        }
      }
      cp_async_wait<1>();  // this thread's copies of pass i+1 have landed
      wait_pass(i + 2);
      __syncthreads();
      if (flagger) st_release_i32(mydone, jp);  // cumulative over the barrier: the retired rows of the pass are visible
    }
  }
}

__global__ void transpose_kernel(const double* __restrict__ src, long long lds, int rows, int cols,
                                 double* __restrict__ dst, long long ldd) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int cc = ty; cc < 32; cc += 8) {
    const int r = r0 + tx, c = c0 + cc;
    tile[cc][tx] = (r < rows && c < cols) ? src[r + (long long)c * lds] : 0.0;
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int c = c0 + tx, r = r0 + rr;  // dst(c, r) = src(r, c)
    if (r < rows && c < cols) dst[c + (long long)r * ldd] = tile[tx][rr];
  }
}

// Z (n x k, ld ldz) <- Q2 Z
int q2_apply(bk_ctx* ctx, const double* VV, const double* TAU, int maxhops, int n, double* Z, long long ldz,
             int k) {
  if (n < 3 || k <= 0) return BK_OK;
  const int nhop = (n - 3) / CB + 1;
  DevBuf<double> Zt;
  DevBuf<int> done;
  const int nchunk = (int)ceil_div(k, Q2_KC);
  const int KC = (int)ceil_div(k, nchunk);  // balanced column chunks, each <= Q2_KC
  BK_TRY(Zt.alloc((size_t)KC * n));
  BK_TRY(done.alloc(nhop));
  // kernel variant by chunk width; CTAs that can be resident at once (the pipeline over hop indices needs them all)
  const int wc = (KC <= 48) ? 48 : (KC <= 80 ? 80 : Q2_KC);
  static const bool use_mma = !(getenv("BK_Q2_MMA") && atoi(getenv("BK_Q2_MMA")) == 0);  // 0: the one-reflector-at-a-time kernel
  void* kern;
  size_t smem;
  int nthreads;
  DevBuf<double> TF;
  DevBuf<long long> offs;
  if (use_mma) {
    kern = (wc == 48) ? (void*)q2_mma_kernel<3> : (wc == 80 ? (void*)q2_mma_kernel<5> : (void*)q2_mma_kernel<10>);
    smem = sizeof(double) * Q2M_NB * Q2M_PB;
    nthreads = 2 * wc + 32;
    // triangular factors of the 8-sweep block reflectors (one table for all column chunks)
    std::vector<long long> hoffs(nhop + 1, 0);
    for (int t = 0; t < nhop; ++t) hoffs[t + 1] = hoffs[t] + (n - 3 - t * CB) / 8 + 1;
    BK_TRY(offs.alloc(nhop + 1));
    BK_TRY(TF.alloc((size_t)hoffs[nhop] * 64));
    BK_CUDA(cudaMemcpyAsync(offs.p, hoffs.data(), sizeof(long long) * (nhop + 1), cudaMemcpyHostToDevice, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));  // hoffs is a stack object
    dim3 tg((unsigned)ceil_div((n - 3) / 8 + 1, 8), (unsigned)nhop);
    q2_tfac8_kernel<<<tg, 256, 0, ctx->stream>>>(VV, TAU, maxhops, n, offs.p, TF.p);
    BK_LAUNCHED(ctx);
    BK_CUDA(cudaGetLastError());
  } else {
    kern = (wc == 48) ? (void*)q2_apply_kernel<48> : (wc == 80 ? (void*)q2_apply_kernel<80> : (void*)q2_apply_kernel<Q2_KC>);
    smem = sizeof(double) * Q2_ROWS * wc;
    nthreads = 2 * wc + 32;
  }
  BK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  BK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthreads, smem));
  per_sm = std::max(1, per_sm);
  for (int k0 = 0; k0 < k; k0 += KC) {
    const int kc = std::min(KC, k - k0);
    dim3 tg((unsigned)ceil_div(n, 32), (unsigned)ceil_div(kc, 32));
    transpose_kernel<<<tg, 256, 0, ctx->stream>>>(Z + (size_t)k0 * ldz, ldz, n, kc, Zt.p, kc);
    BK_LAUNCHED(ctx);
    std::vector<int> init(nhop, INT_MAX);
    BK_CUDA(cudaMemcpyAsync(done.p, init.data(), sizeof(int) * nhop, cudaMemcpyHostToDevice, ctx->stream));
    Q2Args a;
    a.Zt = Zt.p;
    a.ldzt = kc;
    a.n = n;
    a.kc = kc;
    a.VV = VV;
    a.TAU = TAU;
    a.maxhops = maxhops;
    a.done = done.p;
    DevBuf<long long> prof;
    a.prof = nullptr;
    if (getenv("BK_Q2_PROF") && !use_mma) {
      BK_TRY(prof.alloc(8));
      BK_CUDA(cudaMemsetAsync(prof.p, 0, 8 * sizeof(long long), ctx->stream));
      a.prof = prof.p;
    }
    Q2MArgs am;
    am.Zt = Zt.p;
    am.ldzt = kc;
    am.n = n;
    am.kc = kc;
    am.VV = VV;
    am.TAU = TAU;
    am.maxhops = maxhops;
    am.done = done.p;
    am.TF = TF.p;
    am.offs = offs.p;
    void* kargs[] = {use_mma ? (void*)&am : (void*)&a};
    const int G = std::min(ctx->sm_count * per_sm, nhop);
    BK_CUDA(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(nthreads), kargs, smem, ctx->stream));
    BK_LAUNCHED(ctx);
    dim3 tg2((unsigned)ceil_div(kc, 32), (unsigned)ceil_div(n, 32));
    transpose_kernel<<<tg2, 256, 0, ctx->stream>>>(Zt.p, kc, kc, n, Z + (size_t)k0 * ldz, ldz);
    BK_LAUNCHED(ctx);
    BK_CUDA(cudaGetLastError());
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (a.prof) {
      long long h[8];
      BK_CUDA(cudaMemcpyAsync(h, prof.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
      BK_CUDA(cudaStreamSynchronize(ctx->stream));
      const double np = (double)std::max(1LL, h[0]);
      fprintf(stderr, "[q2 prof, CTA 0, kc=%d] passes %lld: cycles per pass: compute %.0f, barrier wait %.0f, flag wait %.0f; "
              "slow sweeps %lld at %.0f cycles; CTA total %.3f Mcycles\n",
              kc, h[0], h[1] / np, h[2] / np, h[3] / np, h[4], h[5] / (double)std::max(1LL, h[4]), h[6] * 1e-6);
    }
  }
  return BK_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Q1 back-transformation: block reflectors of stage 1, last panel first
// ---------------------------------------------------------------------------------------------------------
__global__ void unpack_panel_v_kernel(const double* __restrict__ A, long long lda, int r0, int c0, int b, int m,
                                      double* __restrict__ Vb) {
  const long long total = (long long)m * b;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % m), c = (int)(idx / m);
    double v;
    if (r < c || c >= m - 1)
      v = 0.0;  // above the diagonal, or no reflector for this column
    else if (r == c)
      v = 1.0;
    else
      v = A[(long long)(r0 + r) + (long long)(c0 + c) * lda];
    Vb[r + (long long)c * m] = v;
  }
}

int q1_apply(bk_ctx* ctx, const double* A, long long lda, int n, const double* Tstore, double* Z, long long ldz,
             int k) {
  const int b = CB;
  DevBuf<double> Vb, W1, W2;
  BK_TRY(Vb.alloc((size_t)n * b));
  BK_TRY(W1.alloc((size_t)b * k));
  BK_TRY(W2.alloc((size_t)b * k));
  int npan = 0;
  for (int c0 = 0; c0 < n; c0 += b) {
    if (n - (c0 + b) < 2) break;
    ++npan;
  }
  for (int p = npan - 1; p >= 0; --p) {
    const int c0 = p * b, r0 = c0 + b, m = n - r0;
    const long long tot = (long long)m * b;
    unpack_panel_v_kernel<<<(unsigned)std::min<long long>(ceil_div(tot, 256), 8LL * ctx->sm_count), 256, 0, ctx->stream>>>(
        A, lda, r0, c0, b, m, Vb.p);
    BK_LAUNCHED(ctx);
    const double* Tk = Tstore + (size_t)p * b * b;
    double* Zr = Z + r0;
    BK_TRY(gemm(ctx, true, false, b, k, m, 1.0, Vb.p, m, Zr, ldz, 0.0, W1.p, b));
    BK_TRY(gemm(ctx, false, false, b, k, b, 1.0, Tk, b, W1.p, b, 0.0, W2.p, b));
    BK_TRY(gemm(ctx, false, false, m, k, b, -1.0, Vb.p, m, W2.p, b, 1.0, Zr, ldz));
  }
  BK_CUDA(cudaGetLastError());
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}

int twostage_reduce(bk_ctx* ctx, const double* K, long long ldk, int n, TwoStage* ts, double* d, double* e) {
  const int b = CB;
  Timer tm;
  BK_TRY(tm.init(ctx->stream));
  ts->n = n;
  ts->maxhops = n / b + 2;
  const int npan = (int)ceil_div(n, b);
  BK_TRY(ts->work.borrow(ctx->ws[0], (size_t)n * n));
  // the band is followed by 3 x 64 zero columns: the chasing kernel computes every hop on full 64 x 64 blocks
  BK_TRY(ts->AB.alloc((size_t)LDAB * (n + 3 * CB)));
  BK_CUDA(cudaMemsetAsync(ts->AB.p + (size_t)LDAB * n, 0, sizeof(double) * LDAB * 3 * CB, ctx->stream));
  BK_TRY(ts->Tstore.alloc((size_t)npan * b * b));
  BK_TRY(ts->TAU.alloc((size_t)ts->maxhops * n));
  BK_CUDA(cudaMemsetAsync(ts->TAU.p, 0, sizeof(double) * (size_t)ts->maxhops * n, ctx->stream));
  tm.start();
  BK_TRY(sy2sb(ctx, ts->work.p, n, n, ts->Tstore.p, ts->AB.p, LDAB, &ts->band, K, ldk));
  ts->t_sy2sb = tm.stop();
  BK_TRY(ts->VV.borrow(ctx->ws[1], (size_t)n * n));
  tm.start();
  BK_TRY(sb2st(ctx, ts->AB.p, n, d, e, ts->VV.p, ts->TAU.p, ts->maxhops));
  ts->t_sb2st = tm.stop();
  return BK_OK;
}

// Multi-GPU variant: stage 1 distributed over the ranks of `peer` (sy2sb_dist, straight from X - the kernel matrix is
// never gathered); every rank ends with the complete factored matrix and runs the band -> tridiagonal stage itself.
int twostage_reduce_dist(bk_ctx* ctx, bk_peer* peer, const double* X, long long ldx, int p, double sigma, int n,
                         TwoStage* ts, double* d, double* e) {
  const int b = CB;
  Timer tm;
  BK_TRY(tm.init(ctx->stream));
  ts->n = n;
  ts->maxhops = n / b + 2;
  const int npan = (int)ceil_div(n, b);
  BK_TRY(ts->work.borrow(ctx->ws[0], (size_t)n * n));
  // the band is followed by 3 x 64 zero columns: the chasing kernel computes every hop on full 64 x 64 blocks
  BK_TRY(ts->AB.alloc((size_t)LDAB * (n + 3 * CB)));
  BK_CUDA(cudaMemsetAsync(ts->AB.p + (size_t)LDAB * n, 0, sizeof(double) * LDAB * 3 * CB, ctx->stream));
  BK_TRY(ts->Tstore.alloc((size_t)npan * b * b));
  tm.start();
  BK_TRY(sy2sb_dist(ctx, peer, X, ldx, p, sigma, n, ts->work.p, ts->Tstore.p, ts->AB.p, LDAB, ctx->ws[2], &ts->band));
  ts->t_sy2sb = tm.stop();
  // every rank holds the complete band: all of them chase it (same bits everywhere, nobody waits for a broadcast of
  // the reflectors) so that each can back-transform its own share of the eigenvectors
  BK_TRY(ts->TAU.alloc((size_t)ts->maxhops * n));
  BK_CUDA(cudaMemsetAsync(ts->TAU.p, 0, sizeof(double) * (size_t)ts->maxhops * n, ctx->stream));
  BK_TRY(ts->VV.borrow(ctx->ws[1], (size_t)n * n));
  tm.start();
  BK_TRY(sb2st(ctx, ts->AB.p, n, d, e, ts->VV.p, ts->TAU.p, ts->maxhops));
  ts->t_sb2st = tm.stop();
  return BK_OK;
}

int twostage_back(bk_ctx* ctx, TwoStage* ts, double* Z, long long ldz, int k) {
  Timer tm;
  BK_TRY(tm.init(ctx->stream));
  tm.start();
  // few columns: sliding-window kernel (each window row crosses shared memory once per 4 reflectors);
  // many columns: compact-WY blocks on the DMMA GEMM
  bool blocked = k > 1536;
  if (const char* f = getenv("BK_Q2_BLOCKED")) blocked = atoi(f) != 0;
  if (blocked)
    BK_TRY(q2_apply_blocked(ctx, ts->VV.p, ts->TAU.p, ts->maxhops, ts->n, Z, ldz, k));
  else
    BK_TRY(q2_apply(ctx, ts->VV.p, ts->TAU.p, ts->maxhops, ts->n, Z, ldz, k));
  ts->t_q2 = tm.stop();
  tm.start();
  BK_TRY(q1_apply(ctx, ts->work.p, ts->n, ts->n, ts->Tstore.p, Z, ldz, k));
  ts->t_q1 = tm.stop();
  return BK_OK;
}

// Two-stage pays off when the back-transformation is narrow (the Q2 pass costs ~2 n^2 k flops at low
// arithmetic intensity) and the matrix is large enough for the GEMM-bound stage 1 to beat the SYMV-bound
// one-stage panels.
bool use_twostage(int n, int max_want, double rel_thresh) {
  if (const char* f = getenv("BK_EIG_TWOSTAGE")) return atoi(f) != 0;
  if (n < 4096 || n > 56000) return false;
  // few eigenvectors, or an eigenvalue threshold (eigtrunc: the count is only known after the tridiagonal
  // eigenvalues; eigen_full falls back if the threshold keeps too many at a large n)
  if (max_want <= n / 8 || rel_thresh > 0.0) return true;
  // all eigenvectors: with the GEMM-based Q2 the two-stage path still wins up to n ~ 16k (measured at n = 10k:
  // 0.62 s against 0.79 s); beyond, the 6 n^3 flops of the two back-transformations catch up with the
  // HBM-bound one-stage panels
  return n <= kTwoStageFullMax;
}

}  // namespace bk
