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
  int lower;  // 1: tiles strictly above the diagonal are skipped (symmetric / SYR2K-like updates);
              // 2: as 1, and every tile strictly below the diagonal is also written transposed (full symmetric C)
};

// C = alpha*op(A)*op(B) + beta*C on the context stream; all pointers are device pointers.
// op(A) is m x k, op(B) is k x n; ta/tb select the transposed operand (A stored k x m / B n x k).
int gemm(bk_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A,
         long long lda, const double* B, long long ldb, double beta, double* C, long long ldc,
         int lower = 0);

// Batched: `dprobs` is a DEVICE array of nprob descriptors (caller uploads it); max_m/max_n
// bound the problem sizes, `vec` says every operand is 16-byte aligned with even ld.
int gemm_batched(bk_ctx* ctx, bool ta, bool tb, const GemmProb* dprobs, int nprob, int max_m,
                 int max_n, bool vec);

bool gemm_operands_vec_ok(const void* A, long long lda, const void* B, long long ldb);

}  // namespace bk
