// Self-test of the tcgen05 building blocks (mopoe_umma.cuh): one CTA computes
// D[128][N] = A[128][K] * B[N][K]^T with the 3xFP16 split on the tensor cores.
#include "mopoe_common.cuh"
#include "mopoe_umma.cuh"

namespace mopoe {

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* A, const float* B, float* D, int N, int K,
                                                               int variant, int* err) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  using namespace umma;
  const int t = threadIdx.x, warp = t >> 5;
  unsigned char* a_hi = smraw;
  unsigned char* a_lo = a_hi + 128 * K * 2;
  unsigned char* b_hi = a_lo + 128 * K * 2;
  unsigned char* b_lo = b_hi + N * K * 2;
  const int k8n = K / 8;
  if (variant >= 3) {
    // MN-major operands (variants 3, 4): element (m, k) of A at (m/8)*(K/8)*128 + (k/8)*128 + (k%8)*16 + (m%8)*2 --
    // the layout the K-major z tile of the pipelined DAA kernel has when it is read TRANSPOSED (M = latent, K = row)
    for (int i = t; i < 128 * K; i += 128) {
      const int m = i % 128, k = i / 128;
      __half h, l;
      split_f16(A[m * K + k], h, l);
      const uint32_t off = (m >> 3) * k8n * 128 + (k >> 3) * 128 + (k & 7) * 16 + (m & 7) * 2;
      *reinterpret_cast<__half*>(a_hi + off) = h; *reinterpret_cast<__half*>(a_lo + off) = l;
    }
    for (int i = t; i < N * K; i += 128) {
      const int n = i % N, k = i / N;
      __half h, l;
      split_f16(B[n * K + k], h, l);
      const uint32_t off = (n >> 3) * k8n * 128 + (k >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2;
      *reinterpret_cast<__half*>(b_hi + off) = h; *reinterpret_cast<__half*>(b_lo + off) = l;
    }
  } else {
  if (variant != 2)
  for (int i = t; i < 128 * k8n; i += 128) {
    const int row = i % 128, k8 = i / 128;
    float x[8];
    for (int q = 0; q < 8; ++q) x[q] = A[row * K + k8 * 8 + q];
    store_split8(a_hi, a_lo, core_off(row, k8, 128), x);
  }
  for (int i = t; i < N * k8n; i += 128) {
    const int row = i % N, k8 = i / N;
    float x[8];
    for (int q = 0; q < 8; ++q) x[q] = B[row * K + k8 * 8 + q];
    store_split8(b_hi, b_lo, core_off(row, k8, N), x);
  }
  }
  uint32_t ncols = 32;
  const int a_col = (N + 31) & ~31;   // variant 2: A planes in TMEM behind the accumulator
  while ((int)ncols < (variant == 2 ? a_col + K : N)) ncols <<= 1;
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (variant == 2) {   // thread = row: split the row and park both planes in TMEM
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < K; k0 += 32) {
      uint32_t hi[16], lo[16];
      for (int q = 0; q < 16; ++q) split_pack2(A[t * K + k0 + 2 * q], A[t * K + k0 + 2 * q + 1], hi[q], lo[q]);
      tmem_st16(lane_addr + a_col + k0 / 2, hi);
      tmem_st16(lane_addr + a_col + K / 2 + k0 / 2, lo);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (t == 0 && variant == 2) {
    const uint32_t lbo_b = (N / 8) * 128, sbo = 128;
    const uint32_t idesc = idesc_f16(128, N);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t offb = ks * 2 * lbo_b;
      const uint64_t dbh = smem_desc(smem_u32(b_hi) + offb, lbo_b, sbo), dbl = smem_desc(smem_u32(b_lo) + offb, lbo_b, sbo);
      const uint32_t ah = tmem + a_col + ks * 8, al = tmem + a_col + K / 2 + ks * 8;
      mma_f16_ts(tmem, ah, dbh, idesc, acc); acc = 1;
      mma_f16_ts(tmem, ah, dbl, idesc, 1);
      mma_f16_ts(tmem, al, dbh, idesc, 1);
    }
    mma_commit(&bar);
  }
  if (t == 0 && variant >= 3) {
    // both operands MN-major (instruction descriptor bits 15 / 16); variant 3: LBO = K-group stride (128),
    // SBO = MN-chunk stride; variant 4: swapped
    const uint32_t mn_stride = k8n * 128, kg = 128;
    const uint32_t idesc = idesc_f16(128, N) | (1u << 15) | (1u << 16);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t off = ks * 2 * kg;
      const uint32_t l_ = variant == 3 ? kg : mn_stride, s_ = variant == 3 ? mn_stride : kg;
      const uint64_t dah = smem_desc(smem_u32(a_hi) + off, l_, s_), dal = smem_desc(smem_u32(a_lo) + off, l_, s_);
      const uint64_t dbh = smem_desc(smem_u32(b_hi) + off, l_, s_), dbl = smem_desc(smem_u32(b_lo) + off, l_, s_);
      mma_f16(tmem, dah, dbh, idesc, acc); acc = 1;
      mma_f16(tmem, dah, dbl, idesc, 1);
      mma_f16(tmem, dal, dbh, idesc, 1);
    }
    mma_commit(&bar);
  }
  if (t == 0 && variant != 2 && variant < 3) {
    const uint32_t lbo_a = (128 / 8) * 128, lbo_b = (N / 8) * 128, sbo = 128;
    const uint32_t idesc = idesc_f16(128, N);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t offa = ks * 2 * lbo_a, offb = ks * 2 * lbo_b;
      uint64_t dah, dal, dbh, dbl;
      if (variant == 0) {
        dah = smem_desc(smem_u32(a_hi) + offa, lbo_a, sbo); dal = smem_desc(smem_u32(a_lo) + offa, lbo_a, sbo);
        dbh = smem_desc(smem_u32(b_hi) + offb, lbo_b, sbo); dbl = smem_desc(smem_u32(b_lo) + offb, lbo_b, sbo);
      } else {
        dah = smem_desc(smem_u32(a_hi) + offa, sbo, lbo_a); dal = smem_desc(smem_u32(a_lo) + offa, sbo, lbo_a);
        dbh = smem_desc(smem_u32(b_hi) + offb, sbo, lbo_b); dbl = smem_desc(smem_u32(b_lo) + offb, sbo, lbo_b);
      }
      mma_f16(tmem, dah, dbh, idesc, acc); acc = 1;
      mma_f16(tmem, dah, dbl, idesc, 1);
      mma_f16(tmem, dal, dbh, idesc, 1);
    }
    mma_commit(&bar);
  }
  if (!mbar_wait(&bar, 0)) { if (t == 0) *err = 1; }
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int q = 0; q < 16; ++q) D[(warp * 32 + (t & 31)) * N + c0 + q] = v[q];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

}  // namespace mopoe

using namespace mopoe;

extern "C" int mopoe_umma_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t variant,
                                   int32_t* err_flag, void* stream) {
  if (mopoe_device_count() == 0) { set_error("no CUDA device"); return MOPOE_ENODEV; }
  if (variant == 2 && (K % 32 || ((N + 31) & ~31) + K > 512)) { set_error("selftest variant 2 needs K%%32==0 and N+K<=512 TMEM columns"); return MOPOE_EINVAL; }
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16) { set_error("selftest needs N%%16==0 (16..256), K%%16==0"); return MOPOE_EINVAL; }
  const int smem = (128 + N) * K * 2 * 2;
  if (smem > 200 * 1024) { set_error("selftest operands too large"); return MOPOE_EINVAL; }
  MOPOE_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, variant, err_flag);
  MOPOE_CUDA(cudaGetLastError());
  return MOPOE_OK;
}
