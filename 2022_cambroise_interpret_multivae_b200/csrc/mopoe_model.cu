// MoPoE-VAE forward / ELBO / backward / Adam kernels for sm_100a.
//
// One batch of the path is processed in three phases (DESIGN.md section "kernels"):
//   P1  h_m = relu(x_m W1_m^T + b1_m)                     32x32 register-tiled NT GEMM tiles
//   P2  per row tile: heads -> subset PoE / MoE -> mixture selection -> reparameterise -> decoders
//       -> NLL / KL reductions -> (backward) d z, d heads, d pre-activations      (no cross-row deps)
//   P3  weight gradients as output-stationary TN GEMM tiles (deterministic, no atomics) with the
//       Adam update applied in the tile epilogue
// `mopoe_forward` runs P1 + P2(forward only) as two launches; `mopoe_train_steps` runs
// P1 | grid barrier | P2 | grid barrier | P3 for n_steps batches inside ONE cooperative persistent
// launch (weights never leave L2, no host round trips, logging scalars written per step).
//
// Reference semantics restated here (paths relative to <reference>/experiments):
//   networks.py:30-36,66-77 (Encoder/Decoder), BaseMMVae.py:137-239 (forward/inference),
//   mm_div.py:13-20 (poe), utils.py:63-85 (mixture_component_selection), kl_div.py:7-14,
//   modality.py:42-45 (calc_log_prob), run_epochs.py:73-135 (basic_routine_epoch),
//   utils.py:88-112 (calc_elbo), experiment.py:268-271 (Adam).
#include <cooperative_groups.h>
#include <algorithm>
#include <vector>

#include "mopoe_common.cuh"
#include "mopoe_latent.cuh"
#include "mopoe_umma.cuh"

namespace mopoe {

constexpr int TILE = 32;      // GEMM tile edge (P1/P3)
constexpr int TLD = 34;       // padded leading dimension of a staged chunk
constexpr float HALF_LOG_2PI = 0.91893853320467274178f;

// -------------------------------------------------------------------------------------------
// workspace
// -------------------------------------------------------------------------------------------
// Split-K of the encoder's first layer: small batches have too few 32x32 output tiles to fill the GPU
// and a tile's K loop (up to 444 wide) is their longest serial chain, so wide inputs are cut into
// partial sums (ws.hp) that the P2 tile adds up (+ bias, ReLU) when it stages its hidden rows.
constexpr int64_t P1_SPLIT_MAX_ROWS = 4096;
__host__ __device__ inline int p1_ksplit(int D, int64_t max_rows) {
  if (max_rows > P1_SPLIT_MAX_ROWS) return 1;
  return D >= 256 ? 3 : (D >= 128 ? 2 : 1);
}

struct Workspace {
  float* h[MOPOE_MAX_MODS];    // (N, 256)      post-ReLU hidden
  float* hp[MOPOE_MAX_MODS];   // (ks, N, 256)  split-K partial pre-activations (ks > 1 only)
  int ks[MOPOE_MAX_MODS];      // K splits of the first layer of this modality
  float* dA[MOPOE_MAX_MODS];   // (N, 256)      d loss / d pre-activation
  float* de[MOPOE_MAX_MODS];   // (N, HC)       d loss / d heads
  float* zz[MOPOE_MAX_MODS];   // (2, N, ZD)    decoder inputs [style | content], pass 0 / unimodal
  float* dx[MOPOE_MAX_MODS];   // (2, N, D)     d loss / d x_hat
  double* acc;                 // (MOPOE_N_SCALARS) scalar accumulators of the current step
  unsigned int* bar;           // grid barrier counter
  int64_t max_rows;
};

static int64_t carve(const mopoe_model_desc* d, int64_t N, char* base, Workspace* w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += (bytes + 255) & ~(int64_t)255; return base ? base + o : (char*)nullptr; };
  const int L = d->latent_dim;
  Workspace tmp;
  memset(&tmp, 0, sizeof(tmp));
  tmp.acc = (double*)take(MOPOE_N_SCALARS * sizeof(double));
  tmp.bar = (unsigned int*)take(256);
  for (int m = 0; m < d->n_mods; ++m) {
    const int S = d->style_dims[m], D = d->dims[m];
    tmp.h[m] = (float*)take(N * MOPOE_HIDDEN * 4);
    tmp.ks[m] = p1_ksplit(D, N);
    tmp.hp[m] = tmp.ks[m] > 1 ? (float*)take(tmp.ks[m] * N * MOPOE_HIDDEN * 4) : nullptr;
    tmp.dA[m] = (float*)take(N * MOPOE_HIDDEN * 4);
    tmp.de[m] = (float*)take(N * (2 * L + 2 * S) * 4);
    tmp.zz[m] = (float*)take(2 * N * (S + L) * 4);
    tmp.dx[m] = (float*)take(2 * N * D * 4);
  }
  tmp.max_rows = N;
  if (w) *w = tmp;
  return off;
}

// -------------------------------------------------------------------------------------------
// per-step context (kernel argument)
// -------------------------------------------------------------------------------------------
struct StepCtx {
  const float* x[MOPOE_MAX_MODS];          // data blocks
  const int32_t* row_index[MOPOE_MAX_MODS];  // optional gather lists
  Noise noise;
  int64_t eps_step_stride;                 // n_pass * max_rows * E   (0 for the forward API)
  int64_t eps_pass_stride;                 // max_rows * E
  int sample_latents, use_expert, with_nll, uni_pass, mode;
  int heads_only;                          // forward API: only the encoder heads were asked for (DAA sweep)
  mopoe_forward_out out;                   // forward API outputs (NULLs in training)
  float lr, b1, b2, adam_eps;
  float* adam_m; float* adam_v; int32_t* adam_t; float* grads; float* params;
  mopoe_param_layout lay;
};

__device__ __forceinline__ int64_t src_row(const StepCtx& cx, const mopoe_batch_desc& b, int m, int n) {
  return cx.row_index[m] ? (int64_t)cx.row_index[m][b.row_offset + n] : (int64_t)n;
}

// -------------------------------------------------------------------------------------------
// 32x32 register-tiled GEMM tile:  acc[i][j] = sum_k A(i,k) * B(j,k), 256 threads, 2x2 per thread,
// K streamed in chunks of 32 through shared memory with register double buffering.
// FA/FB: (row-in-tile, k) -> float (0 outside the matrix).  *_KC: operand is contiguous along k
// in memory (lanes walk k) else contiguous along the tile row index (lanes walk i).
// -------------------------------------------------------------------------------------------
template <bool A_KC, bool B_KC, class FA, class FB>
__device__ __forceinline__ void tile_gemm(FA fa, FB fb, int K, float acc[2][2], float* sm) {
  float* As = sm;                   // [2][TILE][TLD]
  float* Bs = sm + 2 * TILE * TLD;  // [2][TILE][TLD]
  const int t = threadIdx.x;
  const int lo = t & 31, hi = t >> 5;
  const int ty = t >> 4, tx = t & 15;
  float ra[4], rb[4];
  acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.f;
  const int nchunk = (K + TILE - 1) / TILE;
  auto fetch = [&](int c) {
    const int k0 = c * TILE;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kk_a = A_KC ? lo : hi + 8 * q, ii_a = A_KC ? hi + 8 * q : lo;
      const int kk_b = B_KC ? lo : hi + 8 * q, ii_b = B_KC ? hi + 8 * q : lo;
      ra[q] = fa(ii_a, k0 + kk_a);
      rb[q] = fb(ii_b, k0 + kk_b);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kk_a = A_KC ? lo : hi + 8 * q, ii_a = A_KC ? hi + 8 * q : lo;
      const int kk_b = B_KC ? lo : hi + 8 * q, ii_b = B_KC ? hi + 8 * q : lo;
      As[(buf * TILE + kk_a) * TLD + ii_a] = ra[q];
      Bs[(buf * TILE + kk_b) * TLD + ii_b] = rb[q];
    }
  };
  __syncthreads();  // previous users of `sm` are done
  fetch(0);
  stash(0);
  __syncthreads();
  // long reductions (weight gradients over tens of thousands of rows) are summed in two levels: a running
  // fp32 sum of 1 024 products is folded into a second accumulator, which keeps the rounding error of the
  // sum at the level of the library GEMM the reference calls (a single fp32 chain of 65 536 terms does not)
  float tot[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int c = 0; c < nchunk; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunk) fetch(c + 1);
    const float* a = As + buf * TILE * TLD + 2 * ty;
    const float* b = Bs + buf * TILE * TLD + 2 * tx;
#pragma unroll 8
    for (int kk = 0; kk < TILE; ++kk) {
      const float2 av = *reinterpret_cast<const float2*>(a + kk * TLD);
      const float2 bv = *reinterpret_cast<const float2*>(b + kk * TLD);
      acc[0][0] = fmaf(av.x, bv.x, acc[0][0]);
      acc[0][1] = fmaf(av.x, bv.y, acc[0][1]);
      acc[1][0] = fmaf(av.y, bv.x, acc[1][0]);
      acc[1][1] = fmaf(av.y, bv.y, acc[1][1]);
    }
    if ((c & 31) == 31) {
      tot[0][0] += acc[0][0]; tot[0][1] += acc[0][1]; tot[1][0] += acc[1][0]; tot[1][1] += acc[1][1];
      acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.f;
    }
    if (c + 1 < nchunk) stash(buf ^ 1);
    __syncthreads();
  }
  acc[0][0] += tot[0][0]; acc[0][1] += tot[0][1]; acc[1][0] += tot[1][0]; acc[1][1] += tot[1][1];
}

// -------------------------------------------------------------------------------------------
// P1: encoder first layer.  Work unit = (present modality, 32-row tile, 32-column tile).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ int p1_units(const ModelView& mv, const mopoe_batch_desc& b, const Workspace& ws) {
  const int tn = (b.n_rows + TILE - 1) / TILE;
  int n = 0;
  for (int m = 0; m < mv.M; ++m)
    if (b.present_mask >> m & 1) n += tn * (MOPOE_HIDDEN / TILE) * ws.ks[m];
  return n;
}

__device__ void p1_unit(const ModelView& mv, const StepCtx& cx, const mopoe_batch_desc& b,
                        const Workspace& ws, int u, float* sm) {
  const int tn = (b.n_rows + TILE - 1) / TILE;
  const int per_split = tn * (MOPOE_HIDDEN / TILE);
  int m = 0, rem = u;
  for (int mm = 0; mm < mv.M; ++mm) {
    if (!(b.present_mask >> mm & 1)) continue;
    const int cnt = per_split * ws.ks[mm];
    if (rem < cnt) { m = mm; break; }
    rem -= cnt;
  }
  const int ks = ws.ks[m], ksi = rem / per_split;
  rem -= ksi * per_split;
  const int n0 = (rem / (MOPOE_HIDDEN / TILE)) * TILE, j0 = (rem % (MOPOE_HIDDEN / TILE)) * TILE;
  const ModView& md = mv.mod[m];
  const int D = md.D, N = b.n_rows;
  const int nch = (D + TILE - 1) / TILE;
  const int k_lo = (ksi * nch / ks) * TILE, k_hi = min(D, ((ksi + 1) * nch / ks) * TILE);   // this split's K range
  const float* x = cx.x[m];
  const float* w1 = md.w1;
  // this thread always fetches the same rows: resolve the gather once
  const int t = threadIdx.x, hi = t >> 5;
  int64_t rows[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int n = n0 + hi + 8 * q;
    rows[q] = n < N ? src_row(cx, b, m, n) : -1;
  }
  auto fa = [&](int i, int k) -> float {
    const int64_t r = rows[(i - hi) >> 3];
    return (r >= 0 && k_lo + k < k_hi) ? x[r * D + k_lo + k] : 0.f;
  };
  auto fb = [&](int j, int k) -> float { return k_lo + k < k_hi ? w1[(int64_t)(j0 + j) * D + k_lo + k] : 0.f; };
  float acc[2][2];
  tile_gemm<true, true>(fa, fb, k_hi - k_lo, acc, sm);
  const int ty = t >> 4, tx = t & 15;
  const int j = j0 + 2 * tx;
  if (ks == 1) {
    const float bj0 = md.b1[j], bj1 = md.b1[j + 1];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int n = n0 + 2 * ty + a;
      if (n < N) {
        float2 o = make_float2(fmaxf(acc[a][0] + bj0, 0.f), fmaxf(acc[a][1] + bj1, 0.f));
        *reinterpret_cast<float2*>(ws.h[m] + (int64_t)n * MOPOE_HIDDEN + j) = o;
      }
    }
  } else {          // partial sums; bias + ReLU when P2 adds the splits up
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int n = n0 + 2 * ty + a;
      if (n < N)
        *reinterpret_cast<float2*>(ws.hp[m] + ((int64_t)ksi * ws.max_rows + n) * MOPOE_HIDDEN + j) = make_float2(acc[a][0], acc[a][1]);
    }
  }
}

// -------------------------------------------------------------------------------------------
// P2 shared-memory plan
// -------------------------------------------------------------------------------------------
struct P2Smem {
  int h, e, de, zz, dzz, rp, rps, dx, part, red, srow, total;  // offsets in floats
  int hc_max, zd_max, s_max;
};

__host__ __device__ inline P2Smem p2_plan(const ModelView& mv, int R) {
  P2Smem p;
  int hc = 0, zd = 0, s = 1;
  for (int m = 0; m < mv.M; ++m) {
    hc = mv.mod[m].HC > hc ? mv.mod[m].HC : hc;
    zd = mv.mod[m].ZD > zd ? mv.mod[m].ZD : zd;
    s = mv.mod[m].S > s ? mv.mod[m].S : s;
  }
  p.hc_max = hc; p.zd_max = zd; p.s_max = s;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 3) & ~3; return o; };
  p.h = take(mv.M * R * MOPOE_HIDDEN);
  p.e = take(mv.M * R * hc);
  p.de = take(mv.M * R * hc);
  p.zz = take(mv.M * 2 * R * zd);
  p.dzz = take(mv.M * 2 * R * zd);
  p.rp = take((1 + mv.M) * R * mv.L);
  p.rps = take(mv.M * 2 * R * s);
  p.dx = take(R * MOPOE_THREADS);
  p.part = take(R * zd > MOPOE_THREADS ? R * zd : MOPOE_THREADS);
  p.red = take(MOPOE_N_SCALARS);
  p.srow = take(2 * mv.M * R);     // int64 source row of every (modality, tile row)
  p.total = off;
  return p;
}

__device__ __forceinline__ void block_add(float* red, int slot, float v) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(red + slot, v);
}

// -------------------------------------------------------------------------------------------
// Latent stage of a row tile, shared by the CUDA-core tile (p2_tile) and the tensor-core tile
// (mopoe_train_tc.cuh): everything between the encoder heads (sh.e, fp32 [modality][row][head column]) and the
// decoder inputs (sh.zz [modality][pass][row][style | content]), and its backward.  256 threads, thread per
// (row, latent dim); the caller synchronises before and after.
// -------------------------------------------------------------------------------------------
struct LatSh {
  float *e, *de, *zz, *dzz, *rp, *rps, *red;
  int R, HCM, ZDM, SM_;
  int NP;   // decoder passes held per modality in zz / dzz / rps (2 in the CUDA-core tile; 1 or 2 in the tensor-core tile)
};

// posterior of subset s at one (row, latent) element, from per-modality precisions T_m = 1 / (exp(lv_m) + eps)
// and mu_m T_m held in registers.  The subset loop is a RUNTIME loop (compact code: an unrolled 15-subset
// version was 7 500 instructions and ran out of the instruction cache); only the modality loops are unrolled, so
// every register array is indexed by compile-time constants and predicated on the subset's member mask.
// poe_fusion + poe: mm_div.py:13-20, BaseMMVae.py:109-122; moe_fusion: BaseMMVae.py:96-106, utils.py:63-85.
struct SubPost { float mu, lv, var, inv; int sel_m; };   // var = exp(lv); inv = 1 / sum T (PoE); sel_m: chosen member (MoE)

__device__ __forceinline__ SubPost sub_post(const ModelView& mv, const mopoe_batch_desc& b, int s, int n, const float* mu_e,
                                            const float* lv_e, const float* ex, const float* T, const float* muT) {
  SubPost r;
  const int mask = mv.sub.mask[s], nm = mv.sub.n_members[s];
  r.sel_m = 0;
  if (moe_like(mv)) {
    int sel = 0;
    for (int i = 0; i < nm; ++i)
      if (n >= b.moe_bounds[nm][i] && n < b.moe_bounds[nm][i + 1]) sel = i;
    const int mem = mv.sub.members[s][sel];
    r.mu = 0.f; r.lv = 0.f; r.var = 1.f;
#pragma unroll
    for (int m = 0; m < MOPOE_MAX_MODS; ++m)
      if (m == mem) { r.mu = mu_e[m]; r.lv = lv_e[m]; r.var = ex[m]; }
    r.sel_m = mem; r.inv = 1.f;
  } else {
    float sT = 0.f, sMT = 0.f;
#pragma unroll
    for (int m = 0; m < MOPOE_MAX_MODS; ++m)
      if (mask >> m & 1) { sT += T[m]; sMT += muT[m]; }
    if (mv.method == MOPOE_METHOD_POE || nm == mv.M) sT += 1.f / (1.f + MOPOE_POE_EPS);   // prior expert N(0, I)
    r.inv = 1.f / sT;
    r.mu = sMT * r.inv;
    r.var = r.inv;
    r.lv = logf(r.inv);
  }
  return r;
}

#ifdef TC_PROF
__device__ float g_latprof[16];
#define LTP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long _n = clock64(); g_latprof[i] += (float)(_n - _lt); _lt = _n; } } while (0)
#else
#define LTP(i) do { } while (0)
#endif

__device__ void lat_forward(const ModelView& mv, const StepCtx& cx, const mopoe_batch_desc& b, int64_t eps_base,
                            int r0, int nr, const LatSh& sh) {
  const int t = threadIdx.x;
#ifdef TC_PROF
  long long _lt = clock64();
#endif
  const int N = b.n_rows, L = mv.L, M = mv.M, present = b.present_mask;
  const bool uni = cx.uni_pass != 0;
  // ---- latent element-wise forward: thread per (row, latent dim); the scalar sums are kept in registers over the
  // thread's items and reduced once (one shuffle tree + one shared atomic per scalar) ----
  float mh_acc[2 * MOPOE_MAX_MODS];
#pragma unroll
  for (int i = 0; i < 2 * MOPOE_MAX_MODS; ++i) mh_acc[i] = 0.f;
  for (int base = 0; base < sh.R * L; base += MOPOE_THREADS) {
    const int idx = base + t;
    const int r = idx / L, l = idx % L;
    const bool valid = idx < sh.R * L && r < nr;
    const int n = r0 + r;
    float mu_e[MOPOE_MAX_MODS], lv_e[MOPOE_MAX_MODS], ex[MOPOE_MAX_MODS], T[MOPOE_MAX_MODS], muT[MOPOE_MAX_MODS];
#pragma unroll
    for (int m = 0; m < MOPOE_MAX_MODS; ++m) {
      const bool on = valid && m < M && (present >> m & 1);
      mu_e[m] = on ? sh.e[(m * sh.R + r) * sh.HCM + l] : 0.f;
      lv_e[m] = on ? sh.e[(m * sh.R + r) * sh.HCM + L + l] : 0.f;
      ex[m] = expf(lv_e[m]);
      T[m] = 1.f / (ex[m] + MOPOE_POE_EPS);
      muT[m] = mu_e[m] * T[m];
      mh_acc[2 * m] += mu_e[m];
      mh_acc[2 * m + 1] += lv_e[m];
    }
    LTP(0);
    float jmu = 0.f, jlv = 0.f;
    float smu[MOPOE_MAX_MODS], slv[MOPOE_MAX_MODS];  // singleton posteriors (poe unimodal passes)
#pragma unroll
    for (int m = 0; m < MOPOE_MAX_MODS; ++m) smu[m] = slv[m] = 0.f;
    const int no = b.owner_mod ? (n / b.owner_div) % b.owner_mod : n;   // row of the selection pattern (mopoe_batch_desc)
    int kidx = 0, owner = 0;
    for (int k = 0; k < b.n_mix; ++k)
      if (no >= b.joint_bounds[k] && no < b.joint_bounds[k + 1]) owner = k;
    LTP(1);
#pragma unroll 1
    for (int s = 0; s < mv.sub.n_subsets; ++s) {
      const int mask = mv.sub.mask[s];
      if ((mask & present) != mask) continue;
      const SubPost ev = sub_post(mv, b, s, no, mu_e, lv_e, ex, T, muT);
      if (valid) {
        if (cx.out.subset_mu) cx.out.subset_mu[((int64_t)s * N + n) * L + l] = ev.mu;
        if (cx.out.subset_logvar) cx.out.subset_logvar[((int64_t)s * N + n) * L + l] = ev.lv;
      }
      // KL of the subset posterior, summed over the CTA's items: one shuffle tree + one shared atomic per warp
      block_add(sh.red, MOPOE_S_KLD_SUBSET + s, valid ? -0.5f * (1.f - ev.var - ev.mu * ev.mu + ev.lv) : 0.f);
      if (mv.sub.n_members[s] == 1) {
#pragma unroll
        for (int m = 0; m < MOPOE_MAX_MODS; ++m)
          if (mask == (1 << m)) { smu[m] = ev.mu; slv[m] = ev.lv; }
      }
      if (in_mixture(mv, b, s)) {
        if (cx.use_expert < 0) {
          if (cx.sample_latents) { if (kidx == owner) { jmu = ev.mu; jlv = ev.lv; } }
          else { jmu += ev.mu; jlv += ev.lv; }
        }
        ++kidx;
      }
      if (cx.use_expert == s) { jmu = ev.mu; jlv = ev.lv; }
    }
    LTP(2);
    if (mv.method == MOPOE_METHOD_JSD) {
      // divergence to the dynamic prior (BaseMMVae.py:81-93, mm_div.py:23-35,69-89): weighted PoE of the mixture
      // components (present unimodal experts, then the prior N(0, I); weights 1 / n_mix), KL of each component to
      // it.  (Rows owned by the prior component keep jmu = jlv = 0: no subset matches their owner index.)
      const float wj = 1.f / (float)b.n_mix;
      float A = 1.f / (1.f + MOPOE_POE_EPS), B = 0.f;
#pragma unroll
      for (int m = 0; m < MOPOE_MAX_MODS; ++m)
        if (m < M && (present >> m & 1)) { A += T[m]; B += muT[m]; }
      const float P = wj * A, mu_a = B / A, lv_a = -logf(P);
      int kc = 0;
#pragma unroll
      for (int m = 0; m < MOPOE_MAX_MODS; ++m)
        if (m < M && (present >> m & 1)) {
          const float dm = mu_e[m] - mu_a;
          block_add(sh.red, MOPOE_S_JSD_DIV + kc, valid ? -0.5f * (1.f - ex[m] * P - dm * dm * P + lv_e[m] - lv_a) : 0.f);
          ++kc;
        }
      block_add(sh.red, MOPOE_S_JSD_DIV + kc, valid ? -0.5f * (1.f - P - mu_a * mu_a * P - lv_a) : 0.f);
    }
    if (cx.use_expert < 0 && !cx.sample_latents) { jmu /= (float)b.n_mix; jlv /= (float)b.n_mix; }
    if (valid) {
      float z = jmu, rp = 0.f;
      if (cx.sample_latents) {
        const float e0 = cx.noise.at(eps_base + (int64_t)n * mv.E + l);
        const float sd = expf(0.5f * jlv);
        z = e0 * sd + jmu;
        rp = 0.5f * e0 * sd;
      }
      sh.rp[r * L + l] = rp;
      for (int m = 0; m < M; ++m)
        if (present >> m & 1) sh.zz[((m * sh.NP + 0) * sh.R + r) * sh.ZDM + mv.mod[m].S + l] = z;
      if (cx.out.joint_mu) cx.out.joint_mu[(int64_t)n * L + l] = jmu;
      if (cx.out.joint_logvar) cx.out.joint_logvar[(int64_t)n * L + l] = jlv;
      if (cx.out.z) cx.out.z[(int64_t)n * L + l] = z;
      if (uni) {
#pragma unroll
        for (int m = 0; m < MOPOE_MAX_MODS; ++m) {
          if (m >= M || !(present >> m & 1)) continue;
          const float e1 = cx.noise.at(eps_base + (1 + m) * cx.eps_pass_stride + (int64_t)n * mv.E + l);
          const float sd = expf(0.5f * slv[m]);
          sh.zz[((m * sh.NP + 1) * sh.R + r) * sh.ZDM + mv.mod[m].S + l] = e1 * sd + smu[m];
          sh.rp[((1 + m) * sh.R + r) * L + l] = 0.5f * e1 * sd;
        }
      }
    }
    LTP(3);
  }
#pragma unroll
  for (int m = 0; m < MOPOE_MAX_MODS; ++m)
    if (m < M && (present >> m & 1)) {
      block_add(sh.red, MOPOE_S_MEAN_HEAD + 4 * m + 0, mh_acc[2 * m]);
      block_add(sh.red, MOPOE_S_MEAN_HEAD + 4 * m + 1, mh_acc[2 * m + 1]);
    }
  LTP(4);
  // ---- style element-wise forward ----
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const ModView& md = mv.mod[m];
    const int S = md.S;
    if (S == 0) continue;
    float a_kl = 0.f, a_mu = 0.f, a_lv = 0.f;
    // (thread index rotated by one warp per modality: for small tiles the content items above occupy warp 0 only, so
    // the style items of modality m go to warp 1 + m and the three sections run side by side)
    const int ts = (t + MOPOE_THREADS - 32 * (1 + m)) & (MOPOE_THREADS - 1);
    for (int base = 0; base < sh.R * S; base += MOPOE_THREADS) {
      const int idx = base + ts;
      const int r = idx / S, s = idx % S;
      const bool valid = idx < sh.R * S && r < nr;
      const int n = r0 + r;
      if (valid) {
        const float mu = sh.e[(m * sh.R + r) * sh.HCM + 2 * L + s];
        const float lv = sh.e[(m * sh.R + r) * sh.HCM + 2 * L + S + s];
        a_kl += -0.5f * (1.f - expf(lv) - mu * mu + lv);
        a_mu += mu; a_lv += lv;
        const float sd = expf(0.5f * lv);
        float zs = mu, rp = 0.f;
        if (cx.sample_latents) {
          const float e0 = cx.noise.at(eps_base + (int64_t)n * mv.E + md.eps_off + s);
          zs = e0 * sd + mu; rp = 0.5f * e0 * sd;
        }
        sh.zz[((m * sh.NP + 0) * sh.R + r) * sh.ZDM + s] = zs;
        sh.rps[((m * sh.NP + 0) * sh.R + r) * sh.SM_ + s] = rp;
        if (cx.out.z_style[m]) cx.out.z_style[m][(int64_t)n * S + s] = zs;
        if (uni) {
          const float e1 = cx.noise.at(eps_base + (1 + m) * cx.eps_pass_stride + (int64_t)n * mv.E + md.eps_off + s);
          sh.zz[((m * sh.NP + 1) * sh.R + r) * sh.ZDM + s] = e1 * sd + mu;
          sh.rps[((m * sh.NP + 1) * sh.R + r) * sh.SM_ + s] = 0.5f * e1 * sd;
        }
      }
    }
    LTP(5);
    block_add(sh.red, MOPOE_S_KLD_STYLE + m, a_kl);
    block_add(sh.red, MOPOE_S_MEAN_HEAD + 4 * m + 2, a_mu);
    block_add(sh.red, MOPOE_S_MEAN_HEAD + 4 * m + 3, a_lv);
    LTP(6);
  }
}

__device__ void lat_backward(const ModelView& mv, const StepCtx& cx, const mopoe_batch_desc& b, int r0, int nr,
                             const LatSh& sh) {
  const int t = threadIdx.x;
  const int N = b.n_rows, L = mv.L, M = mv.M, present = b.present_mask;
  const bool uni = cx.uni_pass != 0;
  const float invN = 1.f / (float)N;
  const float wmix = 1.f / (float)b.n_mix;
  // ---- latent element-wise backward ----
  const float ckl = mv.beta * mv.beta_content * invN;
  for (int base = 0; base < sh.R * L; base += MOPOE_THREADS) {
    const int idx = base + t;
    const int r = idx / L, l = idx % L;
    if (idx < sh.R * L && r < nr) {
      const int n = r0 + r;
      float mu_e[MOPOE_MAX_MODS], lv_e[MOPOE_MAX_MODS], ex[MOPOE_MAX_MODS], T[MOPOE_MAX_MODS], muT[MOPOE_MAX_MODS];
      float dmu[MOPOE_MAX_MODS], dlv[MOPOE_MAX_MODS];
      float gz = 0.f;
#pragma unroll
      for (int m = 0; m < MOPOE_MAX_MODS; ++m) {
        const bool on = m < M && (present >> m & 1);
        mu_e[m] = on ? sh.e[(m * sh.R + r) * sh.HCM + l] : 0.f;
        lv_e[m] = on ? sh.e[(m * sh.R + r) * sh.HCM + L + l] : 0.f;
        ex[m] = expf(lv_e[m]);
        T[m] = 1.f / (ex[m] + MOPOE_POE_EPS);
        muT[m] = mu_e[m] * T[m];
        dmu[m] = dlv[m] = 0.f;
        if (on) gz += sh.dzz[((m * sh.NP + 0) * sh.R + r) * sh.ZDM + mv.mod[m].S + l];
      }
      const int no = b.owner_mod ? (n / b.owner_div) % b.owner_mod : n;
      int kidx = 0, owner = 0;
      for (int k = 0; k < b.n_mix; ++k)
        if (no >= b.joint_bounds[k] && no < b.joint_bounds[k + 1]) owner = k;
      // upstream gradients of each subset posterior (mixture KL share, reparameterised z of the owner rows,
      // unimodal ELBO in poe mode) distributed to the experts (hand-derived)
#pragma unroll 1
      for (int s = 0; s < mv.sub.n_subsets; ++s) {
        const int mask = mv.sub.mask[s];
        if ((mask & present) != mask) continue;
        const SubPost ev = sub_post(mv, b, s, no, mu_e, lv_e, ex, T, muT);
        const int nm = mv.sub.n_members[s];
        float umu = 0.f, ulv = 0.f;
        const float dkl_lv = 0.5f * (ev.var - 1.f);
        if (in_mixture(mv, b, s)) {
          if (mv.method != MOPOE_METHOD_JSD) {   // static prior; the dynamic-prior divergence of jsd follows the loop
            umu += ckl * wmix * ev.mu;
            ulv += ckl * wmix * dkl_lv;
          }
          if (kidx == owner) { umu += gz; ulv += gz * sh.rp[r * L + l]; }
          ++kidx;
        }
        if (mv.method == MOPOE_METHOD_POE && nm == 1) {  // unimodal ELBO of modality m (run_epochs.py:115-125)
          umu += ckl * ev.mu;
          ulv += ckl * dkl_lv;
          if (uni) {
            const int m = mv.sub.members[s][0];
            const float g1 = sh.dzz[((m * sh.NP + 1) * sh.R + r) * sh.ZDM + mv.mod[m].S + l];
            umu += g1; ulv += g1 * sh.rp[((1 + m) * sh.R + r) * L + l];
          }
        }
        if (umu == 0.f && ulv == 0.f) continue;
        if (moe_like(mv)) {
#pragma unroll
          for (int m = 0; m < MOPOE_MAX_MODS; ++m)
            if (m == ev.sel_m) { dmu[m] += umu; dlv[m] += ulv; }
        } else {
          const float invP = ev.inv;
#pragma unroll
          for (int m = 0; m < MOPOE_MAX_MODS; ++m)
            if (mask >> m & 1) {
              dmu[m] += umu * T[m] * invP;
              const float dT = umu * (mu_e[m] - ev.mu) * invP - ulv * invP;
              dlv[m] += dT * (-T[m] * T[m] * ex[m]);
            }
        }
      }
      if (mv.method == MOPOE_METHOD_JSD) {
        // d/d(expert) of D = sum_k w KL(component k || dynamic prior), components = present experts + prior:
        // with A = sum T_k, mu_a = sum mu_k T_k / A, P = w A = 1 / var_a, Q = sum var_k, Sd = sum (mu_k - mu_a),
        // S2 = sum (mu_k - mu_a)^2 and a_m = dT_m / dlv_m = -T_m^2 var_m:
        //   dD/dmu_m = w P ((mu_m - mu_a) - T_m Sd / A)
        //   dD/dlv_m = -w/2 (1 - P var_m - w a_m (Q + S2) + a_m / P + 2 P a_m (mu_m - mu_a) Sd / A)
        float A = 1.f / (1.f + MOPOE_POE_EPS), B = 0.f, Q = 1.f;
#pragma unroll
        for (int m = 0; m < MOPOE_MAX_MODS; ++m)
          if (m < M && (present >> m & 1)) { A += T[m]; B += muT[m]; Q += ex[m]; }
        const float P = wmix * A, mu_a = B / A;
        float Sd = -mu_a, S2 = mu_a * mu_a;
#pragma unroll
        for (int m = 0; m < MOPOE_MAX_MODS; ++m)
          if (m < M && (present >> m & 1)) { const float dm = mu_e[m] - mu_a; Sd += dm; S2 += dm * dm; }
#pragma unroll
        for (int m = 0; m < MOPOE_MAX_MODS; ++m)
          if (m < M && (present >> m & 1)) {
            const float dm = mu_e[m] - mu_a, am = -T[m] * T[m] * ex[m];
            dmu[m] += ckl * wmix * P * (dm - T[m] * Sd / A);
            dlv[m] += ckl * -0.5f * wmix * (1.f - P * ex[m] - wmix * am * (Q + S2) + am / P + 2.f * P * am * dm * Sd / A);
          }
      }
#pragma unroll
      for (int m = 0; m < MOPOE_MAX_MODS; ++m)
        if (m < M && (present >> m & 1)) {
          sh.de[(m * sh.R + r) * sh.HCM + l] = dmu[m];
          sh.de[(m * sh.R + r) * sh.HCM + L + l] = dlv[m];
        }
    }
  }
  // ---- style element-wise backward ----
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const int S = mv.mod[m].S;
    // style KL enters the joint ELBO and (poe) the unimodal ELBO of m, each with beta*beta_style^2
    const float cks = mv.beta * mv.beta_style * mv.beta_style * invN * (mv.method == MOPOE_METHOD_POE ? 2.f : 1.f);
    const int ts = (t + MOPOE_THREADS - 32 * (1 + m)) & (MOPOE_THREADS - 1);   // as in lat_forward: one warp per modality
    for (int idx = ts; idx < sh.R * S; idx += MOPOE_THREADS) {
      const int r = idx / S, s = idx % S;
      if (r >= nr) continue;
      const float mu = sh.e[(m * sh.R + r) * sh.HCM + 2 * L + s];
      const float lv = sh.e[(m * sh.R + r) * sh.HCM + 2 * L + S + s];
      const float g0 = sh.dzz[((m * sh.NP + 0) * sh.R + r) * sh.ZDM + s];
      float gmu = g0 + cks * mu;
      float glv = g0 * sh.rps[((m * sh.NP + 0) * sh.R + r) * sh.SM_ + s] + cks * 0.5f * (expf(lv) - 1.f);
      if (uni) {
        const float g1 = sh.dzz[((m * sh.NP + 1) * sh.R + r) * sh.ZDM + s];
        gmu += g1; glv += g1 * sh.rps[((m * sh.NP + 1) * sh.R + r) * sh.SM_ + s];
      }
      sh.de[(m * sh.R + r) * sh.HCM + 2 * L + s] = gmu;
      sh.de[(m * sh.R + r) * sh.HCM + 2 * L + S + s] = glv;
    }
  }
}

// -------------------------------------------------------------------------------------------
// P2: one tile of R rows, everything between the hidden layer and the per-row gradients.
// -------------------------------------------------------------------------------------------
#ifdef TRAIN_PROF
__device__ float g_p2prof[16];
#define P2T(i) do { if (BWD && blockIdx.x == 0 && threadIdx.x == 0) { const long long _n = clock64(); g_p2prof[i] += (float)(_n - _tp); _tp = _n; } } while (0)
#else
#define P2T(i) do { } while (0)
#endif

// head / decoder weights of every modality resident in shared memory (small models, small batches: the tile's
// phases are chains of weight reads, a round trip to L2 each; from shared memory they cost a tenth)
struct SmemWeights { const float* wh[MOPOE_MAX_MODS]; const float* wd[MOPOE_MAX_MODS]; };

template <int R, bool BWD>
__device__ void p2_tile(const ModelView& mv, const StepCtx& cx, const mopoe_batch_desc& b,
                        const Workspace& ws, int64_t eps_base, int r0, float* sm, const SmemWeights* sw = nullptr) {
  const P2Smem pl = p2_plan(mv, R);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int N = b.n_rows, L = mv.L, M = mv.M, present = b.present_mask;
  const int nr = min(R, N - r0);
  const float invN = 1.f / (float)N;
  float* sh_h = sm + pl.h;
  float* sh_e = sm + pl.e;
  float* sh_de = sm + pl.de;
  float* sh_zz = sm + pl.zz;
  float* sh_dzz = sm + pl.dzz;
  float* sh_rp = sm + pl.rp;
  float* sh_rps = sm + pl.rps;
  float* sh_dx = sm + pl.dx;
  float* sh_part = sm + pl.part;
  float* sh_red = sm + pl.red;
  long long* sh_srow = reinterpret_cast<long long*>(sm + pl.srow);
  const int HCM = pl.hc_max, ZDM = pl.zd_max, SM_ = pl.s_max;
  const bool uni = cx.uni_pass != 0;

#ifdef TRAIN_PROF
  long long _tp = clock64();
#endif
  __syncthreads();
  if (t < MOPOE_N_SCALARS) sh_red[t] = 0.f;
  for (int i = t; i < M * 2 * R * ZDM; i += MOPOE_THREADS) sh_dzz[i] = 0.f;
  for (int i = t; i < M * R; i += MOPOE_THREADS) {      // gather indices of the tile's rows, resolved once
    const int m = i / R, r = i - m * R;
    sh_srow[i] = ((present >> m & 1) && r < nr && cx.with_nll) ? src_row(cx, b, m, r0 + r) : 0;
  }
  // ---- hidden rows -> smem ----
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    for (int i = t; i < R * (MOPOE_HIDDEN / 4); i += MOPOE_THREADS) {
      const int r = i / (MOPOE_HIDDEN / 4), c = i % (MOPOE_HIDDEN / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nr) {
        if (ws.ks[m] == 1) v = *reinterpret_cast<const float4*>(ws.h[m] + (int64_t)(r0 + r) * MOPOE_HIDDEN + 4 * c);
        else {      // add the split-K partial sums of P1, bias, ReLU; P3 reads the finished row from ws.h
          v = *reinterpret_cast<const float4*>(mv.mod[m].b1 + 4 * c);
          for (int ksi = 0; ksi < ws.ks[m]; ++ksi) {
            const float4 q = *reinterpret_cast<const float4*>(ws.hp[m] + ((int64_t)ksi * ws.max_rows + r0 + r) * MOPOE_HIDDEN + 4 * c);
            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
          }
          v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
          *reinterpret_cast<float4*>(ws.h[m] + (int64_t)(r0 + r) * MOPOE_HIDDEN + 4 * c) = v;
        }
      }
      *reinterpret_cast<float4*>(sh_h + (m * R + r) * MOPOE_HIDDEN + 4 * c) = v;
    }
  }
  __syncthreads();
  P2T(0);
  // ---- heads: e[r][j] = bh[j] + h[r] . wh[j]   (warp per output, lanes split k) ----
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const ModView& md = mv.mod[m];
    // every weight read is an L2 round trip (weights change each step, L1 is invalidated by the grid
    // barrier): keep four outputs per warp (8 x LDG.128 per lane) in flight
    for (int jb = warp; jb < md.HC; jb += 4 * (MOPOE_THREADS / 32)) {
      float4 wa[4], wb[4];
      float bjs[4];          // (the bias travels with the weights: loaded after the dot product it is one more L2 round trip per output)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = jb + u * (MOPOE_THREADS / 32);
        const float4* wrow = reinterpret_cast<const float4*>((sw ? sw->wh[m] : md.wh) + (int64_t)(j < md.HC ? j : jb) * MOPOE_HIDDEN);
        wa[u] = wrow[lane]; wb[u] = wrow[lane + 32];
        bjs[u] = md.bh[j < md.HC ? j : jb];
      }
      if constexpr (R <= 4) {
        // small tiles: the 4 R dot products of the warp are reduced TOGETHER, round by round (independent shuffles back
        // to back instead of 4 R dependent five-shuffle chains)
        float accs[4][R];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4* hr = reinterpret_cast<const float4*>(sh_h + (m * R + r) * MOPOE_HIDDEN);
            const float4 ha = hr[lane], hb = hr[lane + 32];
            accs[u][r] = wa[u].x * ha.x + wa[u].y * ha.y + wa[u].z * ha.z + wa[u].w * ha.w + wb[u].x * hb.x + wb[u].y * hb.y +
                         wb[u].z * hb.z + wb[u].w * hb.w;
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int r = 0; r < R; ++r) accs[u][r] += __shfl_xor_sync(0xffffffffu, accs[u][r], o);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = jb + u * (MOPOE_THREADS / 32);
          if (j < md.HC) {
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (lane == r % 32) sh_e[(m * R + r) * HCM + j] = accs[u][r] + bjs[u];
          }
        }
      } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = jb + u * (MOPOE_THREADS / 32);
        if (j < md.HC) {
          float acc[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4* hr = reinterpret_cast<const float4*>(sh_h + (m * R + r) * MOPOE_HIDDEN);
            const float4 ha = hr[lane], hb = hr[lane + 32];
            acc[r] = wa[u].x * ha.x + wa[u].y * ha.y + wa[u].z * ha.z + wa[u].w * ha.w + wb[u].x * hb.x + wb[u].y * hb.y +
                     wb[u].z * hb.z + wb[u].w * hb.w;
          }
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
          const float bj = bjs[u];
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (lane == r % 32) sh_e[(m * R + r) * HCM + j] = acc[r] + bj;
        }
      }
      }
    }
  }
  __syncthreads();
  P2T(1);
  if (cx.out.enc_heads[0] || cx.out.enc_heads[1] || cx.out.enc_heads[2] || cx.out.enc_heads[3]) {
    for (int m = 0; m < M; ++m) {
      if (!(present >> m & 1) || !cx.out.enc_heads[m]) continue;
      const int HC = mv.mod[m].HC;
      for (int i = t; i < nr * HC; i += MOPOE_THREADS)
        cx.out.enc_heads[m][(int64_t)(r0 + i / HC) * HC + i % HC] = sh_e[(m * R + i / HC) * HCM + i % HC];
    }
  }
  if (!BWD && cx.heads_only) return;   // uniform over the CTA; the caller's next tile starts with a barrier
  LatSh lsh;
  lsh.e = sh_e; lsh.de = sh_de; lsh.zz = sh_zz; lsh.dzz = sh_dzz; lsh.rp = sh_rp; lsh.rps = sh_rps; lsh.red = sh_red;
  lsh.R = R; lsh.HCM = HCM; lsh.ZDM = ZDM; lsh.SM_ = SM_; lsh.NP = 2;
  lat_forward(mv, cx, b, eps_base, r0, nr, lsh);
  P2T(2);
  __syncthreads();
  P2T(3);
  // ---- decoders (+ NLL, d x_hat, d z) : thread per output feature, 256 features at a time ----
  const int npass = uni ? 2 : 1;
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const ModView& md = mv.mod[m];
    const int D = md.D, ZD = md.ZD;
    const float* wd_m = sw ? sw->wd[m] : md.wd;
    if (BWD) {  // decoder inputs -> workspace for the weight-gradient phase
      for (int p = 0; p < npass; ++p)
        for (int i = t; i < nr * ZD; i += MOPOE_THREADS)
          ws.zz[m][((int64_t)p * ws.max_rows + r0 + i / ZD) * ZD + i % ZD] = sh_zz[((m * 2 + p) * R + i / ZD) * ZDM + i % ZD];
    }
    for (int p = 0; p < npass; ++p) {
      const float* zrow = sh_zz + (m * 2 + p) * R * ZDM;
      for (int d0 = 0; d0 < D; d0 += MOPOE_THREADS) {
        const int d = d0 + t;
        const int dlen = min(MOPOE_THREADS, D - d0);
        float nll = 0.f;
        if (d < D) {
          float acc[R], xv[R];
          const float bd = md.bd[d];
          const float lam = md.lv[d];        // log-variance and targets: in flight with the weight rows (not behind them)
#pragma unroll
          for (int r = 0; r < R; ++r) {
            acc[r] = bd;
            xv[r] = (cx.with_nll && r < nr) ? cx.x[m][sh_srow[m * R + r] * D + d] : 0.f;
          }
          const float* w = (sw ? sw->wd[m] : md.wd) + (int64_t)d * ZD;
          if ((ZD & 3) == 0) {       // the row is contiguous and 16-byte aligned: 8 x LDG.128 in flight
            for (int k0 = 0; k0 < ZD; k0 += 32) {
              float4 wv[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) wv[q] = k0 + 4 * q < ZD ? *reinterpret_cast<const float4*>(w + k0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (k0 + 4 * q < ZD) {
                  const int k = k0 + 4 * q;
#pragma unroll
                  for (int r = 0; r < R; ++r) {
                    const float* z = zrow + r * ZDM + k;
                    acc[r] = fmaf(z[0], wv[q].x, acc[r]); acc[r] = fmaf(z[1], wv[q].y, acc[r]);
                    acc[r] = fmaf(z[2], wv[q].z, acc[r]); acc[r] = fmaf(z[3], wv[q].w, acc[r]);
                  }
                }
              }
            }
          } else {
#pragma unroll 8
            for (int k = 0; k < ZD; ++k) {
              const float wk = w[k];
#pragma unroll
              for (int r = 0; r < R; ++r) acc[r] = fmaf(zrow[r * ZDM + k], wk, acc[r]);
            }
          }
          const float iv = expf(-lam);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (r < nr) {
              const int n = r0 + r;
              if (p == 0 && cx.out.rec_loc[m]) cx.out.rec_loc[m][(int64_t)n * D + d] = acc[r];
              if (cx.with_nll) {
                const float diff = xv[r] - acc[r];
                nll += 0.5f * diff * diff * iv + 0.5f * lam + HALF_LOG_2PI;
                if (BWD) {
                  const float g = -diff * iv * invN;
                  sh_dx[r * MOPOE_THREADS + t] = g;
                  ws.dx[m][((int64_t)p * ws.max_rows + n) * D + d] = g;
                }
              }
            } else if (BWD) {
              sh_dx[r * MOPOE_THREADS + t] = 0.f;
            }
          }
        }
        if (cx.with_nll) block_add(sh_red, (p == 0 ? MOPOE_S_NLL : MOPOE_S_NLL_UNI) + m, nll);
        if (BWD) {
          __syncthreads();
          // d zz[r][k] += sum_d dx[r][d] * wd[d][k]
          const int P = R * ZD;
          const int nsplit = P >= MOPOE_THREADS ? 1 : MOPOE_THREADS / P;
          for (int q0 = 0; q0 < P * nsplit; q0 += MOPOE_THREADS) {
            const int q = q0 + t;
            if (q < P * nsplit) {
              const int pr = q % P, sp = q / P;
              const int r = pr / ZD, k = pr % ZD;
              float a = 0.f;
#pragma unroll 8
              for (int dd = sp; dd < dlen; dd += nsplit)
                a = fmaf(sh_dx[r * MOPOE_THREADS + dd], wd_m[(int64_t)(d0 + dd) * ZD + k], a);
              if (nsplit == 1) sh_dzz[((m * 2 + p) * R + r) * ZDM + k] += a;
              else sh_part[sp * P + pr] = a;
            }
          }
          if (nsplit > 1) {
            __syncthreads();
            if (t < P) {
              float a = 0.f;
              for (int sp = 0; sp < nsplit; ++sp) a += sh_part[sp * P + t];
              sh_dzz[((m * 2 + p) * R + t / ZD) * ZDM + t % ZD] += a;
            }
          }
          __syncthreads();
        }
      }
    }
  }
  __syncthreads();
  P2T(4);
  if (BWD) {
    lat_backward(mv, cx, b, r0, nr, lsh);
    P2T(5);
    __syncthreads();
    P2T(6);
    // ---- d heads -> workspace; d pre-activation = (W_h^T d heads) * relu' ----
    for (int m = 0; m < M; ++m) {
      if (!(present >> m & 1)) continue;
      const ModView& md = mv.mod[m];
      const int HC = md.HC;
      const float* wh_m = sw ? sw->wh[m] : md.wh;
      for (int i = t; i < nr * HC; i += MOPOE_THREADS)
        ws.de[m][(int64_t)(r0 + i / HC) * HC + i % HC] = sh_de[(m * R + i / HC) * HCM + i % HC];
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll 16
      for (int j = 0; j < HC; ++j) {
        const float w = wh_m[(int64_t)j * MOPOE_HIDDEN + t];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(sh_de[(m * R + r) * HCM + j], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r < nr) ws.dA[m][(int64_t)(r0 + r) * MOPOE_HIDDEN + t] = sh_h[(m * R + r) * MOPOE_HIDDEN + t] > 0.f ? acc[r] : 0.f;
    }
  }
  __syncthreads();
  P2T(7);
  if (t < MOPOE_N_SCALARS && sh_red[t] != 0.f) atomicAdd(ws.acc + t, (double)sh_red[t]);
}

// -------------------------------------------------------------------------------------------
// scalars of one step from the accumulated sums (one thread)
// -------------------------------------------------------------------------------------------
__device__ void finalize_scalars(const ModelView& mv, const StepCtx& cx, const mopoe_batch_desc& b,
                                 const double* acc, float* out) {
  const int M = mv.M, L = mv.L, present = b.present_mask;
  const double N = (double)b.n_rows;
  float s[MOPOE_N_SCALARS];
  for (int i = 0; i < MOPOE_N_SCALARS; ++i) s[i] = 0.f;
  double jd = 0.0, nll = 0.0, kls = 0.0, uni = 0.0;
  for (int q = 0; q < mv.sub.n_subsets; ++q) {
    if ((mv.sub.mask[q] & present) != mv.sub.mask[q]) continue;
    const double kl = acc[MOPOE_S_KLD_SUBSET + q] / N;
    s[MOPOE_S_KLD_SUBSET + q] = (float)kl;
    if (in_mixture(mv, b, q)) jd += kl / (double)b.n_mix;
  }
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const int S = mv.mod[m].S;
    const double nm = acc[MOPOE_S_NLL + m] / N, nu = acc[MOPOE_S_NLL_UNI + m] / N;
    const double ks = acc[MOPOE_S_KLD_STYLE + m] / N;
    s[MOPOE_S_NLL + m] = (float)nm;
    s[MOPOE_S_NLL_UNI + m] = (float)nu;
    s[MOPOE_S_KLD_STYLE + m] = (float)ks;
    s[MOPOE_S_MEAN_HEAD + 4 * m + 0] = (float)(acc[MOPOE_S_MEAN_HEAD + 4 * m + 0] / (N * L));
    s[MOPOE_S_MEAN_HEAD + 4 * m + 1] = (float)(acc[MOPOE_S_MEAN_HEAD + 4 * m + 1] / (N * L));
    if (S > 0) {
      s[MOPOE_S_MEAN_HEAD + 4 * m + 2] = (float)(acc[MOPOE_S_MEAN_HEAD + 4 * m + 2] / (N * S));
      s[MOPOE_S_MEAN_HEAD + 4 * m + 3] = (float)(acc[MOPOE_S_MEAN_HEAD + 4 * m + 3] / (N * S));
    }
    nll += nm;
    kls += mv.beta_style * ks;  // calc_style_kld: style_weights[m] = beta_style
    if (mv.method == MOPOE_METHOD_POE)  // unimodal ELBO (utils.calc_elbo, modality != 'joint')
      uni += nu + mv.beta * (mv.beta_content * (acc[MOPOE_S_KLD_SUBSET + m] / N) + mv.beta_style * (mv.beta_style * ks));
  }
  if (mv.method == MOPOE_METHOD_JSD) {      // divergence_dynamic_prior: sum_k (1 / n_mix) KL(component k || dynamic prior) / N
    jd = 0.0;
    for (int k = 0; k < b.n_mix; ++k) {
      const double kl = acc[MOPOE_S_JSD_DIV + k] / N;
      s[MOPOE_S_JSD_DIV + k] = (float)kl;
      jd += kl / (double)b.n_mix;
    }
  }
  s[MOPOE_S_JOINT_DIV] = (float)jd;
  s[MOPOE_S_TOTAL_LOSS] = (float)(nll + mv.beta * (mv.beta_style * kls + mv.beta_content * jd) + uni);
  s[MOPOE_S_N_ROWS] = (float)b.n_rows;
  s[MOPOE_S_PRESENT] = (float)present;
  for (int i = 0; i < MOPOE_N_SCALARS; ++i) out[i] = s[i];
}

// -------------------------------------------------------------------------------------------
// P3: weight gradients (+ Adam).  Work units:
//   kind 0  dW1_m[j][d]  = sum_n dA_m[n][j] x_m[n][d]            (256 x D)
//   kind 1  dWh_m[j][k]  = sum_n de_m[n][j] h_m[n][k]            (HC x 256)
//   kind 2  dWd_m[d][z]  = sum_{p,n} dx_m[p][n][d] zz_m[p][n][z] (D x ZD)
//   kind 3  column sums: db1 (256), dbh (HC), dbd (D), d logvar (D), 32 columns per unit
// -------------------------------------------------------------------------------------------
struct P3Unit { int m, kind, ti, tj; };

__device__ __forceinline__ int p3_units_of(const ModView& md, int kind) {
  auto c = [](int v) { return (v + TILE - 1) / TILE; };
  switch (kind) {
    case 0: return (MOPOE_HIDDEN / TILE) * c(md.D);
    case 1: return c(md.HC) * (MOPOE_HIDDEN / TILE);
    case 2: return c(md.D) * c(md.ZD);
    default: return (MOPOE_HIDDEN / TILE) + c(md.HC) + 2 * c(md.D);
  }
}

__device__ int p3_total(const ModelView& mv, int present) {
  int n = 0;
  for (int m = 0; m < mv.M; ++m)
    if (present >> m & 1)
      for (int k = 0; k < 4; ++k) n += p3_units_of(mv.mod[m], k);
  return n;
}

__device__ P3Unit p3_decode(const ModelView& mv, int present, int u) {
  P3Unit r = {0, 0, 0, 0};
  auto c = [](int v) { return (v + TILE - 1) / TILE; };
  for (int m = 0; m < mv.M; ++m) {
    if (!(present >> m & 1)) continue;
    for (int k = 0; k < 4; ++k) {
      const int cnt = p3_units_of(mv.mod[m], k);
      if (u < cnt) {
        r.m = m; r.kind = k;
        const int tjn = k == 0 ? c(mv.mod[m].D) : k == 1 ? MOPOE_HIDDEN / TILE : k == 2 ? c(mv.mod[m].ZD) : 1;
        r.ti = u / tjn; r.tj = u % tjn;
        return r;
      }
      u -= cnt;
    }
  }
  return r;
}

// Adam update / gradient store of one scalar parameter (torch.optim.Adam, no amsgrad / decay)
__device__ __forceinline__ void apply_grad(const StepCtx& cx, int64_t idx, float g, float bc1, float bc2s) {
  if (cx.mode == 1) { cx.grads[idx] = g; return; }
  const float m = cx.b1 * cx.adam_m[idx] + (1.f - cx.b1) * g;
  const float v = cx.b2 * cx.adam_v[idx] + (1.f - cx.b2) * g * g;
  cx.adam_m[idx] = m;
  cx.adam_v[idx] = v;
  const float denom = sqrtf(v) / bc2s + cx.adam_eps;
  cx.params[idx] -= (cx.lr / bc1) * (m / denom);
}

__device__ void p3_unit(const ModelView& mv, const StepCtx& cx, const mopoe_batch_desc& b,
                        const Workspace& ws, int u, float* sm) {
  const P3Unit pu = p3_decode(mv, b.present_mask, u);
  const int m = pu.m;
  const ModView& md = mv.mod[m];
  const int N = b.n_rows, D = md.D, HC = md.HC, ZD = md.ZD, t = threadIdx.x;
  const int npass = cx.uni_pass ? 2 : 1;
  // bias corrections of this modality's parameter group (step count t_m incremented by the caller)
  float bc1 = 1.f, bc2s = 1.f;
  if (cx.mode == 2) {
    const float tt = (float)(cx.adam_t[m] + 1);
    bc1 = 1.f - powf(cx.b1, tt);
    bc2s = sqrtf(1.f - powf(cx.b2, tt));
  }
  const int ty = t >> 4, tx = t & 15;
  float acc[2][2];
  if (pu.kind == 0) {
    const int i0 = pu.ti * TILE, j0 = pu.tj * TILE;
    const float* dA = ws.dA[m];
    const float* x = cx.x[m];
    // B operand walks j (feature) fastest: each thread resolves its 4 gathered rows per chunk
    auto fa = [&](int i, int k) -> float { return k < N ? dA[(int64_t)k * MOPOE_HIDDEN + i0 + i] : 0.f; };
    auto fb = [&](int j, int k) -> float {
      return (k < N && j0 + j < D) ? x[src_row(cx, b, m, k) * D + j0 + j] : 0.f;
    };
    tile_gemm<false, false>(fa, fb, N, acc, sm);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int i = i0 + 2 * ty + a, j = j0 + 2 * tx + c;
        if (j < D) apply_grad(cx, cx.lay.enc_w1[m] + (int64_t)i * D + j, acc[a][c], bc1, bc2s);
      }
  } else if (pu.kind == 1) {
    const int i0 = pu.ti * TILE, j0 = pu.tj * TILE;
    const float* de = ws.de[m];
    const float* h = ws.h[m];
    auto fa = [&](int i, int k) -> float { return (k < N && i0 + i < HC) ? de[(int64_t)k * HC + i0 + i] : 0.f; };
    auto fb = [&](int j, int k) -> float { return k < N ? h[(int64_t)k * MOPOE_HIDDEN + j0 + j] : 0.f; };
    tile_gemm<false, false>(fa, fb, N, acc, sm);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int i = i0 + 2 * ty + a, j = j0 + 2 * tx + c;
        if (i < HC) apply_grad(cx, cx.lay.enc_wh[m] + (int64_t)i * MOPOE_HIDDEN + j, acc[a][c], bc1, bc2s);
      }
  } else if (pu.kind == 2) {
    const int i0 = pu.ti * TILE, j0 = pu.tj * TILE;
    const float* dx = ws.dx[m];
    const float* zz = ws.zz[m];
    const int64_t mr = ws.max_rows;
    const int K = npass * N;
    auto fa = [&](int i, int k) -> float {
      if (k >= K || i0 + i >= D) return 0.f;
      const int p = k / N, n = k - p * N;
      return dx[((int64_t)p * mr + n) * D + i0 + i];
    };
    auto fb = [&](int j, int k) -> float {
      if (k >= K || j0 + j >= ZD) return 0.f;
      const int p = k / N, n = k - p * N;
      return zz[((int64_t)p * mr + n) * ZD + j0 + j];
    };
    tile_gemm<false, false>(fa, fb, K, acc, sm);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int i = i0 + 2 * ty + a, j = j0 + 2 * tx + c;
        if (i < D && j < ZD) apply_grad(cx, cx.lay.dec_w[m] + (int64_t)i * ZD + j, acc[a][c], bc1, bc2s);
      }
  } else {
    // column sums: which vector does this unit belong to?
    auto c32 = [](int v) { return (v + TILE - 1) / TILE; };
    int q = pu.ti, which = 0;
    if (q >= MOPOE_HIDDEN / TILE) { q -= MOPOE_HIDDEN / TILE; which = 1;
      if (q >= c32(HC)) { q -= c32(HC); which = 2; if (q >= c32(D)) { q -= c32(D); which = 3; } } }
    const int c0 = q * TILE, col = c0 + (t & 31), grp = t >> 5;
    const int width = which == 0 ? MOPOE_HIDDEN : which == 1 ? HC : D;
    float sum = 0.f;
    if (col < width) {
      if (which == 0) {
        for (int n = grp; n < N; n += 8) sum += ws.dA[m][(int64_t)n * MOPOE_HIDDEN + col];
      } else if (which == 1) {
        for (int n = grp; n < N; n += 8) sum += ws.de[m][(int64_t)n * HC + col];
      } else {
        const float var = which == 3 ? expf(md.lv[col]) : 0.f;
        const float fN = (float)N;
        for (int p = 0; p < npass; ++p)
          for (int n = grp; n < N; n += 8) {
            const float g = ws.dx[m][((int64_t)p * ws.max_rows + n) * D + col];
            // d nll / d logvar_d = (1/N) sum_n 0.5 (1 - diff^2/var),  diff = -g var N
            sum += which == 2 ? g : 0.5f / fN - 0.5f * fN * var * g * g;
          }
      }
    }
    __syncthreads();
    sm[grp * 32 + (t & 31)] = sum;
    __syncthreads();
    if (t < 32 && c0 + t < width) {
      float tot = 0.f;
      for (int g = 0; g < 8; ++g) tot += sm[g * 32 + t];
      const int64_t base = which == 0 ? cx.lay.enc_b1[m] : which == 1 ? cx.lay.enc_bh[m]
                           : which == 2 ? cx.lay.dec_b[m] : cx.lay.dec_lv[m];
      if (which != 3 || mv.learn_scale) apply_grad(cx, base + c0 + t, tot, bc1, bc2s);
      else if (cx.mode == 1) cx.grads[base + c0 + t] = 0.f;
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------
// kernels
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MOPOE_THREADS) p1_kernel(ModelView mv, StepCtx cx, mopoe_batch_desc b, Workspace ws) {
  extern __shared__ __align__(16) float sm[];
  const int nu = p1_units(mv, b, ws);
  for (int u = blockIdx.x; u < nu; u += gridDim.x) p1_unit(mv, cx, b, ws, u, sm);
}

template <int R>
__global__ void __launch_bounds__(MOPOE_THREADS) p2_forward_kernel(ModelView mv, StepCtx cx, mopoe_batch_desc b, Workspace ws) {
  extern __shared__ __align__(16) float sm[];
  const int nt = (b.n_rows + R - 1) / R;
  for (int tile = blockIdx.x; tile < nt; tile += gridDim.x) p2_tile<R, false>(mv, cx, b, ws, 0, tile * R, sm);
}

__global__ void finalize_kernel(ModelView mv, StepCtx cx, mopoe_batch_desc b, Workspace ws) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && cx.out.scalars) finalize_scalars(mv, cx, b, ws.acc, cx.out.scalars);
}

// the fused persistent training kernel: n_steps x (P1 | barrier | P2 | barrier | P3 | barrier)
// WS: the head and decoder weights of every modality are copied into shared memory at the start of each step (one
// cp.async.bulk per matrix, in flight while P1 runs) and the row tiles of P2 read them from there (one CTA per SM).
template <int R, bool WS>
__global__ void __launch_bounds__(MOPOE_THREADS) train_kernel(ModelView mv, StepCtx cx, const mopoe_batch_desc* batches,
                                                              int n_steps, float* scalars, Workspace ws, int ws_smem_off) {
  extern __shared__ __align__(16) float sm[];
  __shared__ mopoe_batch_desc sb;
  __shared__ __align__(8) uint64_t wbar;
  SmemWeights sw;
  uint32_t wbytes = 0;
  if (WS) {
    float* wbase = sm + ws_smem_off;
    int off = 0;
    for (int m = 0; m < MOPOE_MAX_MODS; ++m) {
      sw.wh[m] = sw.wd[m] = nullptr;
      if (m >= mv.M) continue;
      const int nh = (mv.mod[m].HC * MOPOE_HIDDEN + 3) & ~3, nd = (mv.mod[m].D * mv.mod[m].ZD + 3) & ~3;
      sw.wh[m] = wbase + off; off += nh;
      sw.wd[m] = wbase + off; off += nd;
      wbytes += (uint32_t)(nh + nd) * 4;
    }
    if (threadIdx.x == 0) { umma::mbar_init(&wbar, 1); umma::fence_mbar_init(); }
    __syncthreads();
  }
  uint32_t wphase = 0;
  unsigned int target = 0;
  for (int step = 0; step < n_steps; ++step) {
    __syncthreads();
    if (threadIdx.x == 0) sb = batches[step];
    __syncthreads();
    const mopoe_batch_desc& b = sb;
    if (blockIdx.x == 0 && threadIdx.x < MOPOE_N_SCALARS) ws.acc[threadIdx.x] = 0.0;
    if (WS && threadIdx.x == 0) {
      // the weights were updated by other CTAs (generic-proxy stores, ordered by the grid barrier): make them
      // visible to the async proxy, then one bulk copy per matrix (sizes rounded up to 16 bytes: every block of the
      // parameter buffer is 128-byte aligned, the tail is padding)
      umma::fence_async_all();
      umma::mbar_expect_tx(&wbar, wbytes);
      for (int m = 0; m < mv.M; ++m) {
        const int nh = (mv.mod[m].HC * MOPOE_HIDDEN + 3) & ~3, nd = (mv.mod[m].D * mv.mod[m].ZD + 3) & ~3;
        umma::bulk_g2s(const_cast<float*>(sw.wh[m]), mv.mod[m].wh, nh * 4, &wbar);
        umma::bulk_g2s(const_cast<float*>(sw.wd[m]), mv.mod[m].wd, nd * 4, &wbar);
      }
    }
#ifdef TRAIN_PROF
    const long long tp0 = clock64();
#endif
    const int nu1 = p1_units(mv, b, ws);
    for (int u = blockIdx.x; u < nu1; u += gridDim.x) p1_unit(mv, cx, b, ws, u, sm);
#ifdef TRAIN_PROF
    const long long tp0b = clock64();
#endif
    grid_barrier(ws.bar, target);
#ifdef TRAIN_PROF
    const long long tp1 = clock64();
#endif
    if (WS) { umma::mbar_wait(&wbar, wphase); wphase ^= 1; }
    const int nt = (b.n_rows + R - 1) / R;
    const int64_t eps_base = (int64_t)step * cx.eps_step_stride;
    for (int tile = blockIdx.x; tile < nt; tile += gridDim.x) {
      if (cx.mode == 0) p2_tile<R, false>(mv, cx, b, ws, eps_base, tile * R, sm, WS ? &sw : nullptr);
      else p2_tile<R, true>(mv, cx, b, ws, eps_base, tile * R, sm, WS ? &sw : nullptr);
    }
#ifdef TRAIN_PROF
    const long long tp1b = clock64();
#endif
    grid_barrier(ws.bar, target);
#ifdef TRAIN_PROF
    const long long tp2 = clock64();
#endif
    if (blockIdx.x == 0 && threadIdx.x == 0) finalize_scalars(mv, cx, b, ws.acc, scalars + (int64_t)step * MOPOE_N_SCALARS);
    if (cx.mode != 0) {
      const int nu3 = p3_total(mv, b.present_mask);
      for (int u = blockIdx.x; u < nu3; u += gridDim.x) p3_unit(mv, cx, b, ws, u, sm);
#ifdef TRAIN_PROF
      const long long tp2b = clock64();
#endif
      grid_barrier(ws.bar, target);
#ifdef TRAIN_PROF
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        float* o = scalars + (int64_t)step * MOPOE_N_SCALARS + 56;
        const long long tp3 = clock64();
        o[0] = (float)(tp0b - tp0); o[1] = (float)(tp1 - tp0b); o[2] = (float)(tp1b - tp1); o[3] = (float)(tp2 - tp1b);
        o[4] = (float)(tp2b - tp2); o[5] = (float)(tp3 - tp2b);
      }
#endif
      if (cx.mode == 2 && blockIdx.x == 0 && threadIdx.x < mv.M && (b.present_mask >> threadIdx.x & 1))
        cx.adam_t[threadIdx.x] += 1;
      // adam_t is next read in P3 of the following step, two barriers away
    }
  }
}

// -------------------------------------------------------------------------------------------
// host launchers
// -------------------------------------------------------------------------------------------
static int pick_rows(const ModelView& mv, int64_t n_rows, int* smem_bytes, int ctas_per_sm = 2) {
  // small batches: one round of row tiles over the co-resident CTAs (a tile's latency is dominated by
  // weight streaming, not by R)
  const int64_t ctas = (int64_t)ctas_per_sm * num_sms();
  int R = n_rows > 2048 ? 16 : (n_rows <= ctas ? 1 : (n_rows <= 2 * ctas ? 2 : 4));
  int bytes = p2_plan(mv, R).total * 4;
  if (bytes > 220 * 1024) { R = 4; bytes = p2_plan(mv, R).total * 4; }
  const int gemm = 4 * TILE * TLD * 4;
  *smem_bytes = bytes > gemm ? bytes : gemm;
  return R;
}

static int validate_batch(const mopoe_model_desc* d, const mopoe_batch_desc* b) {
  if (b->n_rows < 1) { set_error("n_rows=%d", b->n_rows); return MOPOE_EINVAL; }
  if (b->present_mask <= 0 || b->present_mask >= (1 << d->n_mods)) { set_error("present_mask=%d invalid", b->present_mask); return MOPOE_EINVAL; }
  if (b->n_mix < 1 || b->n_mix > MOPOE_MAX_SUBSETS) { set_error("n_mix=%d invalid", b->n_mix); return MOPOE_EINVAL; }
  if (b->owner_mod < 0 || (b->owner_mod > 0 && b->owner_div < 1)) { set_error("owner_div=%d owner_mod=%d invalid", b->owner_div, b->owner_mod); return MOPOE_EINVAL; }
  if (b->joint_bounds[0] != 0 || b->joint_bounds[b->n_mix] != (b->owner_mod ? b->owner_mod : b->n_rows)) { set_error("joint_bounds do not span the batch"); return MOPOE_EINVAL; }
  return MOPOE_OK;
}

#include "mopoe_train_tc.cuh"

// 227 KB of shared memory per CTA minus the kernel's static shared variables (1 KB with alignment)
constexpr int TC_SMEM_LIMIT = 227 * 1024 - 2048;
static float* g_tc_prof = nullptr;   // device address of the TC_PROF counters of the last tensor-core launch
static int g_train_impl = 0;   // implementation of the last mopoe_train_steps call: 0 CUDA cores, 1 tensor cores

// MOPOE_TRAIN_IMPL=tc|ffma forces one implementation (the tests cross-check both).  Default: the tensor-core
// kernel for batches of TC_MIN_ROWS rows and more (measured on B200: 4 096 rows 0.61 vs 0.82 ms, 65 536 rows
// 2.2 vs 17 ms, stress shape 4.3 vs 27 ms), the CUDA-core kernel below that (a 256-row step is bound by the
// latency of its dependent stages, not by arithmetic: 74 vs 136 us)
constexpr int64_t TC_MIN_ROWS = 1024;
static int pick_train_impl(const mopoe_model_desc* d, int64_t max_rows, tc::TcPlan* plan) {
  const char* force = getenv("MOPOE_TRAIN_IMPL");
  if (force && !strcmp(force, "ffma")) return 0;
  const bool forced = force && !strcmp(force, "tc");
  if (!forced && max_rows < TC_MIN_ROWS) return 0;
  const bool ok = tc::make_plan(d, max_rows, TC_SMEM_LIMIT, plan);
  if (forced) return ok ? 1 : -1;
  return ok ? 1 : 0;
}

#include "mopoe_generic.cuh"

// architectures outside the train_exp defaults run on the layered path (mopoe_generic.cuh)
static bool layered_path(const mopoe_model_desc* d) {
  return d->n_hidden_enc != 1 || d->n_hidden_dec != 0 || d->scale_mode != 0 || d->likelihood != 0;
}

// build_view for the layered path: dims, subset table and loss weights only (the weight pointers of ModView belong
// to the fused kernels; the layered path addresses the buffer through GModel offsets)
static void build_view_layered(const mopoe_model_desc* d, const mopoe_param_layout* lay, float* base, ModelView* v) {
  mopoe_param_layout tmp = *lay;
  for (int m = 0; m < MOPOE_MAX_MODS; ++m) {
    if (tmp.enc_w1[m] < 0) tmp.enc_w1[m] = 0;
    if (tmp.enc_b1[m] < 0) tmp.enc_b1[m] = 0;
    if (tmp.dec_lv[m] < 0) tmp.dec_lv[m] = 0;
  }
  build_view(d, &tmp, base, v);
}

}  // namespace mopoe

using namespace mopoe;

extern "C" {

#ifdef TRAIN_PROF
int mopoe_debug_p2prof(float* out16_host) {
  MOPOE_CUDA(cudaMemcpyFromSymbol(out16_host, g_p2prof, 16 * sizeof(float)));
  float z[16] = {0};
  MOPOE_CUDA(cudaMemcpyToSymbol(g_p2prof, z, sizeof(z)));
  return 0;
}
#endif

int64_t mopoe_workspace_bytes(const mopoe_model_desc* desc, int64_t max_rows) {
  if (check_desc(desc)) return MOPOE_EINVAL;
  if (max_rows < 1) { set_error("max_rows=%lld", (long long)max_rows); return MOPOE_EINVAL; }
  if (layered_path(desc)) {
    mopoe_param_layout lay;
    mopoe_param_layout_of(desc, &lay);
    return gen::gen_carve(desc, &lay, max_rows, nullptr, nullptr);
  }
  int64_t bytes = (carve(desc, max_rows, nullptr, nullptr) + 1023) & ~(int64_t)1023;
  tc::TcPlan plan;
  if (tc::make_plan(desc, max_rows, TC_SMEM_LIMIT, &plan)) bytes += plan.total;   // operand blobs of the tensor-core training step
  return bytes;
}

int mopoe_train_last_impl(void) { return g_train_impl; }

#ifdef TC_PROF
// profiling builds only: copy out and clear the 64 per-stage cycle counters of the tensor-core training kernel
int mopoe_debug_tcprof(float* out64_host) {
  if (!g_tc_prof) { set_error("no tensor-core launch yet"); return MOPOE_EINVAL; }
  MOPOE_CUDA(cudaMemcpy(out64_host, g_tc_prof, 64 * sizeof(float), cudaMemcpyDeviceToHost));
  MOPOE_CUDA(cudaMemset(g_tc_prof, 0, 64 * sizeof(float)));
  MOPOE_CUDA(cudaMemcpyFromSymbol(out64_host + 40, g_latprof, 16 * sizeof(float)));
  float z16[16] = {0};
  MOPOE_CUDA(cudaMemcpyToSymbol(g_latprof, z16, sizeof(z16)));
  return MOPOE_OK;
}
#endif

int mopoe_forward(const mopoe_model_desc* desc, const float* params, const mopoe_batch_desc* batch,
                  const float* const* x, const float* eps, uint64_t seed, int sample_latents, int use_expert,
                  int with_nll, const mopoe_forward_out* out, void* workspace, int64_t workspace_bytes, void* stream_) {
  int rc = check_desc(desc);
  if (rc) return rc;
  if (mopoe_device_count() == 0) { set_error("no CUDA device: the MoPoE path has no CPU fallback"); return MOPOE_ENODEV; }
  if (!params || !batch || !x || !out || !workspace) { set_error("NULL argument"); return MOPOE_EINVAL; }
  if ((rc = validate_batch(desc, batch))) return rc;
  mopoe_param_layout lay;
  mopoe_param_layout_of(desc, &lay);
  ModelView mv;
  const bool layered = layered_path(desc);
  if (layered) build_view_layered(desc, &lay, const_cast<float*>(params), &mv);
  else build_view(desc, &lay, const_cast<float*>(params), &mv);
  if (use_expert >= mv.sub.n_subsets) { set_error("use_expert=%d out of range", use_expert); return MOPOE_EINVAL; }
  if (use_expert >= 0 && (mv.sub.mask[use_expert] & batch->present_mask) != mv.sub.mask[use_expert]) {
    set_error("use_expert subset %d is not available in this batch", use_expert); return MOPOE_EINVAL; }
  if (layered) {
    const int64_t gneed = gen::gen_carve(desc, &lay, batch->n_rows, nullptr, nullptr);
    if (workspace_bytes < gneed) { set_error("workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)gneed); return MOPOE_ENOSPC; }
    gen::GWs gws;
    gen::gen_carve(desc, &lay, batch->n_rows, (char*)workspace, &gws);
    gen::GModel gm;
    gen::build_gmodel(desc, &lay, &gm);
    StepCtx cx;
    memset(&cx, 0, sizeof(cx));
    for (int m = 0; m < desc->n_mods; ++m) {
      cx.x[m] = (batch->present_mask >> m & 1) ? x[m] : nullptr;
      if ((batch->present_mask >> m & 1) && !x[m]) { set_error("x[%d] is NULL but modality is present", m); return MOPOE_EINVAL; }
    }
    cx.noise = make_noise(eps, seed, MOPOE_STREAM_FORWARD);
    cx.eps_pass_stride = (int64_t)batch->n_rows * mv.E;
    cx.sample_latents = sample_latents; cx.use_expert = use_expert < 0 ? -1 : use_expert;
    cx.with_nll = with_nll; cx.uni_pass = 0; cx.mode = 0;
    cx.out = *out;
    {
      bool only = !with_nll && !out->scalars && !out->subset_mu && !out->subset_logvar && !out->joint_mu && !out->joint_logvar && !out->z;
      for (int m = 0; m < MOPOE_MAX_MODS; ++m) only = only && !out->z_style[m] && !out->rec_loc[m] && !out->rec_logvar[m];
      cx.heads_only = only ? 1 : 0;
    }
    cx.lay = lay;
    mopoe_batch_desc b = *batch;
    b.row_offset = 0;
    return gen::gen_step(desc, gm, mv, cx, b, gws, const_cast<float*>(params), 0, out->scalars, (cudaStream_t)stream_);
  }
  const int64_t need = carve(desc, batch->n_rows, nullptr, nullptr);
  if (workspace_bytes < need) { set_error("workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need); return MOPOE_ENOSPC; }
  Workspace ws;
  carve(desc, batch->n_rows, (char*)workspace, &ws);
  cudaStream_t stream = (cudaStream_t)stream_;
  StepCtx cx;
  memset(&cx, 0, sizeof(cx));
  for (int m = 0; m < desc->n_mods; ++m) {
    cx.x[m] = (batch->present_mask >> m & 1) ? x[m] : nullptr;
    if ((batch->present_mask >> m & 1) && !x[m]) { set_error("x[%d] is NULL but modality is present", m); return MOPOE_EINVAL; }
  }
  cx.noise = make_noise(eps, seed, MOPOE_STREAM_FORWARD);
  cx.eps_pass_stride = (int64_t)batch->n_rows * mv.E;
  cx.sample_latents = sample_latents; cx.use_expert = use_expert < 0 ? -1 : use_expert;
  cx.with_nll = with_nll; cx.uni_pass = 0; cx.mode = 0;
  cx.out = *out;
  {
    bool only = !with_nll && !out->scalars && !out->subset_mu && !out->subset_logvar && !out->joint_mu && !out->joint_logvar && !out->z;
    for (int m = 0; m < MOPOE_MAX_MODS; ++m) only = only && !out->z_style[m] && !out->rec_loc[m];
    cx.heads_only = only ? 1 : 0;
  }
  cx.lay = lay;
  mopoe_batch_desc b = *batch;
  b.row_offset = 0;
  if (!cx.heads_only) MOPOE_CUDA(cudaMemsetAsync(ws.acc, 0, MOPOE_N_SCALARS * sizeof(double), stream));   // (the heads-only pass of the DAA sweep accumulates no scalars)
  const int tn = (b.n_rows + TILE - 1) / TILE;
  int nu1 = 0;
  for (int m = 0; m < desc->n_mods; ++m)
    if (b.present_mask >> m & 1) nu1 += tn * (MOPOE_HIDDEN / TILE) * ws.ks[m];
  const int sms = num_sms();
  p1_kernel<<<nu1 < 8 * sms ? nu1 : 8 * sms, MOPOE_THREADS, 4 * TILE * TLD * 4, stream>>>(mv, cx, b, ws);
  MOPOE_CUDA(cudaGetLastError());
  int smem = 0;
  int R = pick_rows(mv, b.n_rows, &smem, 4);
  // heads-only pass (the encoder sweep of the DAA): the tile streams the head weights and does little else, so 4 rows
  // per tile quarter the L2 traffic of the 1- and 2-row tilings (1 000 rows: 29 -> ~12 us)
  if (cx.heads_only && R < 4 && b.n_rows >= 2 * sms) { R = 4; smem = p2_plan(mv, 4).total * 4; const int gemm = 4 * TILE * TLD * 4; if (smem < gemm) smem = gemm; }
  const int nt = (b.n_rows + R - 1) / R;
  if (R == 1) {
    MOPOE_CUDA(cudaFuncSetAttribute(p2_forward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    p2_forward_kernel<1><<<nt < 4 * sms ? nt : 4 * sms, MOPOE_THREADS, smem, stream>>>(mv, cx, b, ws);
  } else if (R == 2) {
    MOPOE_CUDA(cudaFuncSetAttribute(p2_forward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    p2_forward_kernel<2><<<nt < 4 * sms ? nt : 4 * sms, MOPOE_THREADS, smem, stream>>>(mv, cx, b, ws);
  } else if (R == 4) {
    MOPOE_CUDA(cudaFuncSetAttribute(p2_forward_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    p2_forward_kernel<4><<<nt < 4 * sms ? nt : 4 * sms, MOPOE_THREADS, smem, stream>>>(mv, cx, b, ws);
  } else {
    MOPOE_CUDA(cudaFuncSetAttribute(p2_forward_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    p2_forward_kernel<16><<<nt < 2 * sms ? nt : 2 * sms, MOPOE_THREADS, smem, stream>>>(mv, cx, b, ws);
  }
  MOPOE_CUDA(cudaGetLastError());
  if (out->scalars) {
    finalize_kernel<<<1, 32, 0, stream>>>(mv, cx, b, ws);
    MOPOE_CUDA(cudaGetLastError());
  }
  return MOPOE_OK;
}

int mopoe_train_steps(const mopoe_model_desc* desc, float* params, float* adam_m, float* adam_v, int32_t* adam_t,
                      float* grads, const float* const* data, const int32_t* const* row_index,
                      const mopoe_batch_desc* batches, int32_t n_steps, int64_t max_rows, const float* eps,
                      uint64_t seed, int mode, float lr, float b1, float b2, float adam_eps, float* scalars,
                      const mopoe_forward_out* out, void* workspace, int64_t workspace_bytes, void* stream_) {
  int rc = check_desc(desc);
  if (rc) return rc;
  if (mopoe_device_count() == 0) { set_error("no CUDA device: the MoPoE path has no CPU fallback"); return MOPOE_ENODEV; }
  if (!params || !data || !batches || !scalars || !workspace) { set_error("NULL argument"); return MOPOE_EINVAL; }
  if (mode < 0 || mode > 2) { set_error("mode=%d", mode); return MOPOE_EINVAL; }
  if (mode == 1 && (!grads || n_steps != 1)) { set_error("mode 1 needs grads and n_steps == 1"); return MOPOE_EINVAL; }
  if (mode == 2 && (!adam_m || !adam_v || !adam_t)) { set_error("mode 2 needs Adam state"); return MOPOE_EINVAL; }
  if (out && n_steps != 1) { set_error("forward outputs need n_steps == 1"); return MOPOE_EINVAL; }
  if (n_steps < 1 || max_rows < 1) { set_error("n_steps=%d max_rows=%lld", n_steps, (long long)max_rows); return MOPOE_EINVAL; }
  if (layered_path(desc)) {
    // host-driven: the batch descriptors come back to the host once (one synchronisation per call; this path is not
    // capturable in a CUDA graph), then every step is a sequence of launches
    mopoe_param_layout glay;
    mopoe_param_layout_of(desc, &glay);
    const int64_t gneed = gen::gen_carve(desc, &glay, max_rows, nullptr, nullptr);
    if (workspace_bytes < gneed) { set_error("workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)gneed); return MOPOE_ENOSPC; }
    gen::GWs gws;
    gen::gen_carve(desc, &glay, max_rows, (char*)workspace, &gws);
    gen::GModel gm;
    gen::build_gmodel(desc, &glay, &gm);
    ModelView gmv;
    build_view_layered(desc, &glay, params, &gmv);
    cudaStream_t gstream = (cudaStream_t)stream_;
    std::vector<mopoe_batch_desc> hb(n_steps);
    MOPOE_CUDA(cudaMemcpyAsync(hb.data(), batches, sizeof(mopoe_batch_desc) * n_steps, cudaMemcpyDeviceToHost, gstream));
    MOPOE_CUDA(cudaStreamSynchronize(gstream));
    StepCtx gcx;
    memset(&gcx, 0, sizeof(gcx));
    for (int m = 0; m < desc->n_mods; ++m) { gcx.x[m] = data[m]; gcx.row_index[m] = row_index ? row_index[m] : nullptr; }
    const int g_pass = desc->method == MOPOE_METHOD_POE ? 1 + desc->n_mods : 1;
    gcx.noise = make_noise(eps, seed, MOPOE_STREAM_TRAIN);
    gcx.eps_pass_stride = max_rows * gmv.E;
    gcx.eps_step_stride = (int64_t)g_pass * max_rows * gmv.E;
    gcx.sample_latents = 1; gcx.use_expert = -1; gcx.with_nll = 1;
    gcx.uni_pass = desc->method == MOPOE_METHOD_POE; gcx.mode = mode;
    gcx.lr = lr; gcx.b1 = b1; gcx.b2 = b2; gcx.adam_eps = adam_eps;
    gcx.adam_m = adam_m; gcx.adam_v = adam_v; gcx.adam_t = adam_t; gcx.grads = grads; gcx.params = params;
    gcx.lay = glay;
    if (out) { gcx.out = *out; gcx.out.scalars = nullptr; }
    if (mode == 1) MOPOE_CUDA(cudaMemsetAsync(grads, 0, glay.total * sizeof(float), gstream));
    g_train_impl = 2;
    for (int step = 0; step < n_steps; ++step) {
      if ((rc = validate_batch(desc, &hb[step]))) return rc;
      if (hb[step].n_rows > max_rows) { set_error("step %d: n_rows=%d > max_rows", step, hb[step].n_rows); return MOPOE_EINVAL; }
      for (int m = 0; m < desc->n_mods; ++m)
        if ((hb[step].present_mask >> m & 1) && !data[m]) { set_error("data[%d] is NULL but modality is present", m); return MOPOE_EINVAL; }
      if ((rc = gen::gen_step(desc, gm, gmv, gcx, hb[step], gws, params, (int64_t)step * gcx.eps_step_stride,
                              scalars + (int64_t)step * MOPOE_N_SCALARS, gstream))) return rc;
    }
    return MOPOE_OK;
  }
  const int64_t need = carve(desc, max_rows, nullptr, nullptr);
  if (workspace_bytes < need) { set_error("workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need); return MOPOE_ENOSPC; }
  mopoe_param_layout lay;
  mopoe_param_layout_of(desc, &lay);
  ModelView mv;
  build_view(desc, &lay, params, &mv);
  Workspace ws;
  carve(desc, max_rows, (char*)workspace, &ws);
  cudaStream_t stream = (cudaStream_t)stream_;
  StepCtx cx;
  memset(&cx, 0, sizeof(cx));
  for (int m = 0; m < desc->n_mods; ++m) {
    cx.x[m] = data[m];
    cx.row_index[m] = row_index ? row_index[m] : nullptr;
  }
  const int n_pass = desc->method == MOPOE_METHOD_POE ? 1 + desc->n_mods : 1;
  cx.noise = make_noise(eps, seed, MOPOE_STREAM_TRAIN);
  cx.eps_pass_stride = max_rows * mv.E;
  cx.eps_step_stride = (int64_t)n_pass * max_rows * mv.E;
  cx.sample_latents = 1; cx.use_expert = -1; cx.with_nll = 1;
  cx.uni_pass = desc->method == MOPOE_METHOD_POE; cx.mode = mode;
  cx.lr = lr; cx.b1 = b1; cx.b2 = b2; cx.adam_eps = adam_eps;
  cx.adam_m = adam_m; cx.adam_v = adam_v; cx.adam_t = adam_t; cx.grads = grads; cx.params = params;
  cx.lay = lay;
  if (out) { cx.out = *out; cx.out.scalars = nullptr; }
  if (mode == 1) MOPOE_CUDA(cudaMemsetAsync(grads, 0, lay.total * sizeof(float), stream));
  MOPOE_CUDA(cudaMemsetAsync(ws.bar, 0, sizeof(unsigned int), stream));
  tc::TcPlan plan;
  const int impl = pick_train_impl(desc, max_rows, &plan);
  if (impl < 0) { set_error("MOPOE_TRAIN_IMPL=tc but the configuration does not fit the tensor-core tiling"); return MOPOE_EINVAL; }
  g_train_impl = impl;
  if (impl == 1) {
    const int64_t base_off = (need + 1023) & ~(int64_t)1023;
    if (workspace_bytes < base_off + plan.total) { set_error("workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)(base_off + plan.total)); return MOPOE_ENOSPC; }
    plan.base = (unsigned char*)workspace + base_off;
    MOPOE_CUDA(cudaMemsetAsync(plan.base + plan.err, 0, 256, stream));
    g_tc_prof = reinterpret_cast<float*>(plan.base + plan.err) + 64;
    MOPOE_CUDA(cudaMemsetAsync(plan.base + plan.p3cnt, 0, (size_t)(tc::MAX_UNITS + 64) * 4, stream));
    void* fn = plan.R == 32 ? (void*)tc::train_tc_kernel<32> : (void*)tc::train_tc_kernel<16>;
    MOPOE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.s_total));
    int per_sm = 0;
    MOPOE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, tc::THREADS, plan.s_total));
    if (per_sm < 1) { set_error("tensor-core train kernel does not fit on an SM (smem %d)", plan.s_total); return MOPOE_EINVAL; }
    const mopoe_batch_desc* bptr = batches;
    float* sptr = scalars;
    void* args[] = {&mv, &cx, &bptr, &n_steps, &sptr, &ws, &plan};
    MOPOE_CUDA(cudaLaunchCooperativeKernel(fn, dim3(num_sms()), dim3(tc::THREADS), args, plan.s_total, stream));
    return MOPOE_OK;
  }
  int smem = 0;
  int R = pick_rows(mv, max_rows, &smem);
  // small model + small batch: head / decoder weights resident in shared memory, one CTA per SM.  OPT-IN
  // (MOPOE_TRAIN_SMEM_WEIGHTS=1): measured on B200 at 256 rows it is SLOWER (82 vs 65 us per step) -- the tile's phases
  // are bound by their other round trips and by cold code, not by the weight reads, and with one CTA per SM the
  // first-layer and weight-gradient phases need two rounds of work units (P3 18 -> 35 us)
  int wfloats = 0;
  for (int m = 0; m < desc->n_mods; ++m)
    wfloats += ((mv.mod[m].HC * MOPOE_HIDDEN + 3) & ~3) + ((mv.mod[m].D * mv.mod[m].ZD + 3) & ~3);
  bool use_ws = false;
  int ws_smem_off = 0;
  {
    const char* e = getenv("MOPOE_TRAIN_SMEM_WEIGHTS");
    int smem1 = 0;
    const int R1 = pick_rows(mv, max_rows, &smem1, 1);
    const int total = ((smem1 + 15) & ~15) + wfloats * 4;
    if (e && e[0] == '1' && R1 <= 2 && total <= 227 * 1024 - 1024) {
      use_ws = true; R = R1; ws_smem_off = ((smem1 + 15) & ~15) / 4; smem = total;
    }
  }
  void* fn = use_ws ? (R == 1 ? (void*)train_kernel<1, true> : (void*)train_kernel<2, true>)
                    : (R == 1 ? (void*)train_kernel<1, false> : R == 2 ? (void*)train_kernel<2, false>
                       : R == 4 ? (void*)train_kernel<4, false> : (void*)train_kernel<16, false>);
  MOPOE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  MOPOE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, MOPOE_THREADS, smem));
  if (per_sm < 1) { set_error("train kernel does not fit on an SM (smem %d)", smem); return MOPOE_EINVAL; }
  // small batches are latency bound (serial weight-streaming chains per tile): two co-resident CTAs per SM
  // hide each other's latency and give every phase one round of work units; large batches keep one CTA
  // per SM (the phases are tile loops, extra CTAs only lengthen the barriers)
  const int grid = (!use_ws && R <= 2 && per_sm >= 2) ? 2 * num_sms() : num_sms();   // measured: 130 / 86 / 90 / 99 us per step at 1 / 2 / 3 / 4 CTAs per SM
  const mopoe_batch_desc* bptr = batches;
  float* sptr = scalars;
  void* args[] = {&mv, &cx, &bptr, &n_steps, &sptr, &ws, &ws_smem_off};
  MOPOE_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(MOPOE_THREADS), args, smem, stream));
  return MOPOE_OK;
}

}  // extern "C"
