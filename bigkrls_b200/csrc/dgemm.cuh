// FP64 GEMM (DMMA) - host-side launchers.  See dgemm.cu.
#pragma once
#include "common.cuh"

namespace bk {

struct GemmProb {
  const double* A;
  const double* B;
  double* C;
  int m, n, k;
  long long lda, ldb, ldc;
  double alpha, beta;
  int stagger = 0;        // cycles of start delay per resident-slot index for the CTAs of the FIRST wave (0 = none)
  int stagger_sms = 1;    // SMs of the device (CTA id / stagger_sms = slot index of a first-wave CTA)
  int stagger_slots = 0;  // resident CTAs per SM
  int prefetch_c = 1;     // 64 x 64 read-modify-write tiles: fetch the C tile into registers before the main loop
  const double* Cin = nullptr;  // read-modify-write from a different matrix: C = alpha op(A) op(B) + beta Cin (nullptr: C)
  long long ldcin = 0;
  int lower;  // 1: tiles strictly above the diagonal are skipped (symmetric / SYR2K-like updates);
              // 2: as 1, and every tile strictly below the diagonal is also written transposed (full symmetric C)
};

// C = alpha*op(A)*op(B) + beta*C on the context stream; all pointers are device pointers.
// op(A) is m x k, op(B) is k x n; ta/tb select the transposed operand (A stored k x m / B n x k).
int gemm(bk_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A,
         long long lda, const double* B, long long ldb, double beta, double* C, long long ldc,
         int lower = 0);
// out of place: C = alpha op(A) op(B) + beta Cin (Cin may be read-only; with lower = 2 the mirrored tile is written to C)
int gemm_oop(bk_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, long long lda,
             const double* B, long long ldb, double beta, const double* Cin, long long ldcin, double* C, long long ldc,
             int lower = 0);

// Batched: `dprobs` is a DEVICE array of nprob descriptors (caller uploads it); max_m/max_n
// bound the problem sizes, `vec` says every operand is 16-byte aligned with even ld.
int gemm_batched(bk_ctx* ctx, bool ta, bool tb, const GemmProb* dprobs, int nprob, int max_m,
                 int max_n, bool vec);

bool gemm_operands_vec_ok(const void* A, long long lda, const void* B, long long ldb);

}  // namespace bk
