// Shared device/host helpers of the MoPoE-VAE B200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mopoe_b200.h"

#define MOPOE_THREADS 256
#define MOPOE_POE_EPS 1e-8f  // mm_div.py:13

namespace mopoe {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_desc(const mopoe_model_desc* d);
int cuda_fail(cudaError_t e, const char* what);
int num_sms();

#define MOPOE_CUDA(call)                                        \
  do {                                                          \
    cudaError_t _e = (call);                                    \
    if (_e != cudaSuccess) return ::mopoe::cuda_fail(_e, #call); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// device view of the model (passed by value as a kernel argument)
// ---------------------------------------------------------------------------------------------
struct ModView {
  float* w1;   // (256, D)
  float* b1;   // (256)
  float* wh;   // (HC, 256)   HC = 2L + 2S
  float* bh;   // (HC)
  float* wd;   // (D, ZD)     ZD = S + L
  float* bd;   // (D)
  float* lv;   // (D)
  int D, S, HC, ZD;
  int eps_off;  // column of this modality's style block inside an eps row
  int peps_off; // the same inside a block-padded noise row (see ModelView::EP)
};

struct SubsetTable {
  int n_subsets;                       // 2^M - 1, BaseExperiment.set_subsets order
  int mask[MOPOE_MAX_SUBSETS];         // member bitmask
  int n_members[MOPOE_MAX_SUBSETS];
  int members[MOPOE_MAX_SUBSETS][MOPOE_MAX_MODS];  // fusion order (sorted by modality NAME)
};

struct ModelView {
  ModView mod[MOPOE_MAX_MODS];
  SubsetTable sub;
  int M, L, E;      // E = eps row width = L + sum S
  int EP;           // block-padded noise row: content and every style section rounded up to 4 (one
                    // Philox4x32 block = 4 normals), so a thread that owns a row draws whole blocks
  int method;
  int learn_scale;
  float beta, beta_style, beta_content;
};

void build_subsets(const mopoe_model_desc* d, SubsetTable* t);
void build_view(const mopoe_model_desc* d, const mopoe_param_layout* lay, float* base, ModelView* v);

// ---------------------------------------------------------------------------------------------
// device utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Grid-wide barrier for cooperative (co-resident) launches.  `bar` is a zero-initialised counter;
// every CTA keeps its own running `target`.  Same fence/atomic/spin pattern as cooperative groups'
// grid.sync(): bar.sync orders the CTA's writes before thread 0's gpu-scope fence, the fence after
// the spin invalidates L1 so later plain loads observe other SMs' writes.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    while (*((volatile unsigned int*)bar) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// counter-based N(0,1): Philox4x32-10 + Box-Muller (restated in oracle/philox.py)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                                       uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// The ten round keys of a (seed) are launch constants: the host expands them once (Noise::rk, kernel
// parameter = constant bank), which removes the 20 per-thread key-schedule adds from every block.
__host__ __device__ __forceinline__ void philox_round_keys(uint64_t seed, uint32_t* rk /* [20] */) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}
__device__ __forceinline__ void philox4x32_10_rk(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, const uint32_t* rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk[2 * r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
}

__device__ __forceinline__ float u01(uint32_t x) {
  return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// -ln(u) for u in (0,1): MUFU.LG2 away from 1, the log1p series next to 1 (where lg2's absolute error
// would dominate); both evaluated, selected branch-free
__device__ __forceinline__ float neg_log_u(float u) {
  const float t = 1.0f - u;   // exact
  const float series = t * (1.0f + t * (0.5f + t * (0.33333334f + t * 0.25f)));
  return u > 0.99f ? series : -0.69314718f * __log2f(u);
}

// the four normals of block `blk` (elements 4*blk .. 4*blk+3) of (seed, stream):
// Box-Muller with sin/cos evaluated on [-pi, pi) (MUFU range of full accuracy) and negated
__device__ __forceinline__ void box_muller4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, float out[4]) {
  float r0, r1;   // MUFU.SQRT: the argument is a positive normal number, no special cases to patch up
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r0) : "f"(2.0f * neg_log_u(u01(c0))));
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r1) : "f"(2.0f * neg_log_u(u01(c2))));
  float s0, q0, s1, q1;
  __sincosf(6.28318530718f * (u01(c1) - 0.5f), &s0, &q0);
  __sincosf(6.28318530718f * (u01(c3) - 0.5f), &s1, &q1);
  out[0] = -r0 * q0; out[1] = -r0 * s0; out[2] = -r1 * q1; out[3] = -r1 * s1;
}
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t stream, uint64_t blk, float out[4]) {
  uint32_t c0 = (uint32_t)blk, c1 = (uint32_t)(blk >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  box_muller4(c0, c1, c2, c3, out);
}

// element `idx` of (seed, stream)
__device__ __forceinline__ float philox_normal1(uint64_t seed, uint64_t stream, uint64_t idx) {
  float v[4];
  philox_normal4(seed, stream, idx >> 2, v);
  return v[idx & 3];
}

// noise source: injected tensor or the generator, addressed by the same flat index
struct Noise {
  const float* eps;  // NULL => philox
  uint64_t seed;
  uint64_t stream;
  uint32_t rk[20];   // round keys of `seed` (make_noise)
  __device__ __forceinline__ float at(uint64_t idx) const;
};
inline Noise make_noise(const float* eps, uint64_t seed, uint64_t stream) {
  Noise n;
  n.eps = eps; n.seed = seed; n.stream = stream;
  philox_round_keys(seed, n.rk);
  return n;
}
// the four normals of block `blk` of a noise source whose round keys sit in the constant bank
__device__ __forceinline__ void philox_normal4(const Noise& nz, uint64_t blk, float out[4]) {
  uint32_t c0 = (uint32_t)blk, c1 = (uint32_t)(blk >> 32), c2 = (uint32_t)nz.stream, c3 = (uint32_t)(nz.stream >> 32);
  philox4x32_10_rk(c0, c1, c2, c3, nz.rk);
  box_muller4(c0, c1, c2, c3, out);
}
__device__ __forceinline__ float Noise::at(uint64_t idx) const {
  if (eps) return eps[idx];
  float v[4];
  philox_normal4(*this, idx >> 2, v);
  return v[idx & 3];
}

}  // namespace mopoe
