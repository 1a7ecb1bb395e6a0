/* Test-only entry point of libmopoe_b200_selftest.so (built beside the product library by csrc/Makefile; not part of
 * the drop-in boundary and not linked into libmopoe_b200.so).  Used by tests/test_gpu_umma.py. */
#ifndef MOPOE_B200_SELFTEST_H
#define MOPOE_B200_SELFTEST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Self-test of the tcgen05 building blocks: one CTA computes D[128][N] = A[128][K] * B[N][K]^T on
 * the tensor cores with the 3xFP16 split the DAA kernel uses (N%16==0, K%16==0). */
int mopoe_umma_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t variant,
                        int32_t* err_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOPOE_B200_SELFTEST_H */
