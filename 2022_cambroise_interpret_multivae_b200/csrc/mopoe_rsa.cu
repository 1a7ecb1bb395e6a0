// Representational similarity analysis on the GPU: the arithmetic of
// experiments/stat_utils.py:25-53 (data2cmat / cmat2triu / vec2cmat) and :81-95 (fit_rsa = scipy.stats.kendalltau of
// the upper triangles) that experiments/workflow.py:741-789 (rsa_exp) runs per (model, latent, validation, score).
//
//   mopoe_rsa_cmat      pairwise (dis)similarity matrix of n rows: Euclidean distance in fp64, accumulated in
//                       column order with separately rounded products and sums (what scipy's pdist loop does),
//                       or the categorical 0 / 1 "differs" matrix
//   mopoe_rsa_kendall   Kendall tau-b sufficient statistics of ONE matrix against n_ref reference matrices: the
//                       upper triangles (P = n (n - 1) / 2 entries each) are compared entry pair by entry pair --
//                       P^2 sign products per reference (2e9 for n = 301), integer counts only, so the result is
//                       exact and independent of the launch geometry.  tau and the p-value follow from seven
//                       integers per reference (host side, rsa.py).
//
// HBM traffic is negligible (the triangles are 0.36 MB each); the kernel is bound by fp64 compares: thread = entry
// i, the j entries stream through shared memory in tiles, the j range is split over blockIdx.y for occupancy.
#include "mopoe_common.cuh"

namespace mopoe {

__global__ void __launch_bounds__(256) rsa_cmat_kernel(int n, int d, const float* __restrict__ data, int categorical, double* __restrict__ cmat) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
  if (b >= n) return;
  double out;
  if (categorical) {
    out = data[(int64_t)a * d] != data[(int64_t)b * d] ? 1.0 : 0.0;
  } else {
    double s = 0.0;
    for (int k = 0; k < d; ++k) {
      const double df = (double)data[(int64_t)a * d + k] - (double)data[(int64_t)b * d + k];
      s = __dadd_rn(s, __dmul_rn(df, df));       // no fused multiply-add: same rounding as the sequential C loop
    }
    out = sqrt(s);
  }
  cmat[(int64_t)a * n + b] = out;
}

// upper triangle (k = 1) in np.triu_indices order
__global__ void __launch_bounds__(256) rsa_triu_kernel(int n, int n_mat, const double* __restrict__ cmat, const double* __restrict__ refs, double* __restrict__ vec) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y, m = blockIdx.z;
  if (b >= n || b <= a) return;
  const int64_t P = (int64_t)n * (n - 1) / 2;
  const int64_t p = (int64_t)a * n - (int64_t)a * (a + 1) / 2 + (b - a - 1);
  const double* src = m == 0 ? cmat : refs + (int64_t)(m - 1) * n * n;
  vec[m * P + p] = src[(int64_t)a * n + b];
}

constexpr int KT_THREADS = 256, KT_TILE = 2048;

// counts[ref][0] = sum_{i != j} sign(x_i - x_j) sign(y_i - y_j)   (= 2 (concordant - discordant))
// counts[ref][1..3] = sum_i cx_i, sum_i cx_i (cx_i - 1), sum_i cx_i (2 cx_i + 7)     cx_i = #{j != i: x_j == x_i}
// counts[ref][4..6] = the same for y.  The tie sums need the full count of an entry, so each CTA row block first
// accumulates per-entry counts over its j chunk into global int arrays; a second kernel folds them.
__global__ void __launch_bounds__(KT_THREADS) rsa_kendall_pairs_kernel(int64_t P, const double* __restrict__ vec, int* __restrict__ per_entry) {
  __shared__ double sx[KT_TILE], sy[KT_TILE];
  const int ref = blockIdx.z;
  const double* x = vec;
  const double* y = vec + (int64_t)(ref + 1) * P;
  const int64_t i = (int64_t)blockIdx.x * KT_THREADS + threadIdx.x;
  const double xi = i < P ? x[i] : 0.0, yi = i < P ? y[i] : 0.0;
  const int64_t chunk = (P + gridDim.y - 1) / gridDim.y;
  const int64_t j0 = blockIdx.y * chunk, j1 = j0 + chunk < P ? j0 + chunk : P;
  int s = 0, cx = 0, cy = 0;
  for (int64_t t0 = j0; t0 < j1; t0 += KT_TILE) {
    const int nt = (int)(j1 - t0 < KT_TILE ? j1 - t0 : KT_TILE);
    __syncthreads();
    for (int k = threadIdx.x; k < nt; k += KT_THREADS) { sx[k] = x[t0 + k]; sy[k] = y[t0 + k]; }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < nt; ++k) {
      const double xj = sx[k], yj = sy[k];
      const int gx = (xi > xj) - (xi < xj), gy = (yi > yj) - (yi < yj);
      s += gx * gy;
      cx += xi == xj;
      cy += yi == yj;
    }
  }
  if (i < P) {
    if (i >= j0 && i < j1) { --cx; --cy; }        // the entry itself
    int* o = per_entry + ((int64_t)ref * 3) * P;
    atomicAdd(o + i, s);
    atomicAdd(o + P + i, cx);
    atomicAdd(o + 2 * P + i, cy);
  }
}

__global__ void __launch_bounds__(256) rsa_kendall_fold_kernel(int64_t P, const int* __restrict__ per_entry, long long* __restrict__ counts) {
  const int ref = blockIdx.y;
  const int* o = per_entry + ((int64_t)ref * 3) * P;
  long long a[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    const long long s = o[i], cx = o[P + i], cy = o[2 * P + i];
    a[0] += s;
    a[1] += cx; a[2] += cx * (cx - 1); a[3] += cx * (2 * cx + 7);
    a[4] += cy; a[5] += cy * (cy - 1); a[6] += cy * (2 * cy + 7);
  }
  __shared__ long long red[7][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    long long v = a[k];
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    long long v = 0;
    for (int q = 0; q < 8; ++q) v += red[threadIdx.x][q];
    atomicAdd(reinterpret_cast<unsigned long long*>(counts + ref * 7 + threadIdx.x), (unsigned long long)v);   // integers: order-free
  }
}

}  // namespace mopoe

using namespace mopoe;

extern "C" {

int mopoe_rsa_cmat(int32_t n, int32_t d, const float* data, int32_t categorical, double* cmat, void* stream_) {
  if (mopoe_device_count() == 0) { set_error("no CUDA device: the RSA path has no CPU fallback"); return MOPOE_ENODEV; }
  if (n < 2 || d < 1 || !data || !cmat) { set_error("rsa_cmat: n=%d d=%d or NULL argument", n, d); return MOPOE_EINVAL; }
  if (categorical && d != 1) { set_error("rsa_cmat: a categorical characteristic is one column"); return MOPOE_EINVAL; }
  if (n > 65535) { set_error("rsa_cmat: n=%d > 65535", n); return MOPOE_EINVAL; }
  rsa_cmat_kernel<<<dim3((n + 255) / 256, n), 256, 0, (cudaStream_t)stream_>>>(n, d, data, categorical, cmat);
  MOPOE_CUDA(cudaGetLastError());
  return MOPOE_OK;
}

int64_t mopoe_rsa_kendall_workspace_bytes(int32_t n, int32_t n_ref) {
  if (n < 2 || n_ref < 1) return -1;
  const int64_t P = (int64_t)n * (n - 1) / 2;
  return (int64_t)(n_ref + 1) * P * 8 + (int64_t)n_ref * 3 * P * 4 + 256;
}

int mopoe_rsa_kendall(int32_t n, int32_t n_ref, const double* cmat, const double* ref_cmats, int64_t* counts,
                      void* workspace, int64_t workspace_bytes, void* stream_) {
  if (mopoe_device_count() == 0) { set_error("no CUDA device: the RSA path has no CPU fallback"); return MOPOE_ENODEV; }
  if (n < 2 || n_ref < 1 || n_ref > 65535 || n > 65535 || !cmat || !ref_cmats || !counts || !workspace) { set_error("rsa_kendall: invalid argument"); return MOPOE_EINVAL; }
  const int64_t need = mopoe_rsa_kendall_workspace_bytes(n, n_ref);
  if (workspace_bytes < need) { set_error("rsa_kendall: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need); return MOPOE_ENOSPC; }
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t P = (int64_t)n * (n - 1) / 2;
  if (P > ((int64_t)1 << 30)) { set_error("rsa_kendall: %lld entries per triangle is too many", (long long)P); return MOPOE_EINVAL; }
  double* vec = reinterpret_cast<double*>(workspace);
  int* per_entry = reinterpret_cast<int*>(vec + (int64_t)(n_ref + 1) * P);
  MOPOE_CUDA(cudaMemsetAsync(per_entry, 0, (size_t)n_ref * 3 * P * 4, stream));
  MOPOE_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_ref * 7 * 8, stream));
  rsa_triu_kernel<<<dim3((n + 255) / 256, n, n_ref + 1), 256, 0, stream>>>(n, n_ref + 1, cmat, ref_cmats, vec);
  MOPOE_CUDA(cudaGetLastError());
  const int gx = (int)((P + KT_THREADS - 1) / KT_THREADS);
  // enough CTAs for a few waves over the SMs: the j range of every entry block is split
  int split = (int)((8LL * num_sms() + (int64_t)gx * n_ref - 1) / ((int64_t)gx * n_ref));
  const int max_split = (int)((P + KT_TILE - 1) / KT_TILE);
  split = split < 1 ? 1 : (split > max_split ? max_split : split);
  rsa_kendall_pairs_kernel<<<dim3(gx, split, n_ref), KT_THREADS, 0, stream>>>(P, vec, per_entry);
  MOPOE_CUDA(cudaGetLastError());
  rsa_kendall_fold_kernel<<<dim3(64, n_ref), 256, 0, stream>>>(P, per_entry, reinterpret_cast<long long*>(counts));
  MOPOE_CUDA(cudaGetLastError());
  return MOPOE_OK;
}

}  // extern "C"
