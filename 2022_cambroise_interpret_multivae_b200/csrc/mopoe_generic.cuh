// Layered path for the NON-DEFAULT architectures of the reference, included by mopoe_model.cu
// inside namespace mopoe (it uses ModelView, StepCtx, LatSh, lat_forward / lat_backward, tile_gemm,
// finalize_scalars defined there):
//   num_hidden_layer_encoder != 1 / num_hidden_layer_decoder != 0   networks.py:16-20,51-55
//   learn_output_sample_scale (per-sample output log-variance)      networks.py:58-59,73-74
//   likelihood = laplace                                            modalities/modality.py:18-30
// The fused persistent kernels (mopoe_model.cu, mopoe_train_tc.cuh) are built around the train_exp defaults -- one
// hidden encoder layer, affine decoder, per-feature scale, normal likelihood; every other combination runs here as a
// sequence of launches per step: one tiled GEMM kernel per layer (the 32 x 32 register tile of the fused kernel, with
// bias / ReLU / ReLU-mask epilogues), the SAME latent stage as the fused kernels (lat_forward / lat_backward on row
// tiles), an element-wise likelihood kernel, column sums, and one Adam / gradient-store kernel.  Same C-ABI entry
// points, same parameter-buffer / state-dict layout rules, same scalars; deterministic (no atomics on gradients).
#pragma once

namespace gen {

constexpr int LAT_R = 8;                // rows per CTA of the latent kernels

struct GMod {                           // offsets (floats) into the parameter / gradient / Adam buffers
  int64_t ew[MOPOE_MAX_LAYERS], eb[MOPOE_MAX_LAYERS];
  int64_t wh, bh;
  int64_t dw[MOPOE_MAX_LAYERS], db[MOPOE_MAX_LAYERS];
  int64_t wo, bo, lv, lvw, lvb;
  int D, S, HC, ZD, in_h, in_o;
};
struct GModel { GMod mod[MOPOE_MAX_MODS]; int He, Hd, scale_mode, likelihood; int64_t total; };

static void build_gmodel(const mopoe_model_desc* d, const mopoe_param_layout* lay, GModel* g) {
  memset(g, 0, sizeof(*g));
  g->He = d->n_hidden_enc; g->Hd = d->n_hidden_dec; g->scale_mode = d->scale_mode; g->likelihood = d->likelihood;
  g->total = lay->total;
  for (int m = 0; m < d->n_mods; ++m) {
    GMod& q = g->mod[m];
    q.D = d->dims[m]; q.S = d->style_dims[m]; q.HC = 2 * d->latent_dim + 2 * q.S; q.ZD = q.S + d->latent_dim;
    q.in_h = g->He >= 1 ? MOPOE_HIDDEN : q.D;
    q.in_o = g->Hd >= 1 ? MOPOE_HIDDEN : q.ZD;
    for (int l = 0; l < MOPOE_MAX_LAYERS; ++l) {
      q.ew[l] = l == 0 ? lay->enc_w1[m] : lay->enc_wx[m][l - 1];
      q.eb[l] = l == 0 ? lay->enc_b1[m] : lay->enc_bx[m][l - 1];
      q.dw[l] = lay->dec_hw[m][l]; q.db[l] = lay->dec_hb[m][l];
    }
    q.wh = lay->enc_wh[m]; q.bh = lay->enc_bh[m];
    q.wo = lay->dec_w[m]; q.bo = lay->dec_b[m];
    q.lv = lay->dec_lv[m]; q.lvw = lay->dec_lvw[m]; q.lvb = lay->dec_lvb[m];
  }
}

struct GWs {
  float* xg[MOPOE_MAX_MODS];                       // (N, D)     gathered input rows
  float* ha[MOPOE_MAX_MODS][MOPOE_MAX_LAYERS];     // (N, 256)   encoder hidden activations (post-ReLU)
  float* heads[MOPOE_MAX_MODS];                    // (N, HC)
  float* de[MOPOE_MAX_MODS];                       // (N, HC)
  float* zz[MOPOE_MAX_MODS];                       // (2, N, ZD) decoder inputs, pass 0 / unimodal pass
  float* dzz[MOPOE_MAX_MODS];                      // (2, N, ZD)
  float* hd[MOPOE_MAX_MODS][2][MOPOE_MAX_LAYERS];  // (N, 256)   decoder hidden activations
  float* loc[MOPOE_MAX_MODS][2];                   // (N, D)
  float* lvs[MOPOE_MAX_MODS][2];                   // (N, D)     per-sample log-variance (scale_mode 1)
  float* dx[MOPOE_MAX_MODS][2];                    // (N, D)     d loss / d loc
  float* dlv[MOPOE_MAX_MODS][2];                   // (N, D)     d loss / d log-variance (element-wise)
  float* ga; float* gb;                            // (N, max(256, ZD))  activation gradients, ping-pong
  float* rp;                                       // (1 + M, N, L)
  float* rps;                                      // (M, 2, N, Smax)
  float* G;                                        // flat gradient buffer (parameter layout)
  double* acc;                                     // MOPOE_N_SCALARS
  int64_t N; int smax;
};

static int64_t gen_carve(const mopoe_model_desc* d, const mopoe_param_layout* lay, int64_t N, char* base, GWs* w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += (bytes + 255) & ~(int64_t)255; return base ? base + o : (char*)nullptr; };
  GWs t;
  memset(&t, 0, sizeof(t));
  const int L = d->latent_dim, H = MOPOE_HIDDEN;
  int smax = 1, gw = H;
  t.acc = (double*)take(MOPOE_N_SCALARS * sizeof(double));
  for (int m = 0; m < d->n_mods; ++m) {
    const int S = d->style_dims[m], D = d->dims[m], HC = 2 * L + 2 * S, ZD = S + L;
    smax = S > smax ? S : smax;
    gw = ZD > gw ? ZD : gw; gw = D > gw ? D : gw;
    t.xg[m] = (float*)take(N * D * 4);
    for (int l = 0; l < d->n_hidden_enc; ++l) t.ha[m][l] = (float*)take(N * H * 4);
    t.heads[m] = (float*)take(N * HC * 4);
    t.de[m] = (float*)take(N * HC * 4);
    t.zz[m] = (float*)take(2 * N * ZD * 4);
    t.dzz[m] = (float*)take(2 * N * ZD * 4);
    for (int p = 0; p < 2; ++p) {
      for (int l = 0; l < d->n_hidden_dec; ++l) t.hd[m][p][l] = (float*)take(N * H * 4);
      t.loc[m][p] = (float*)take(N * D * 4);
      t.lvs[m][p] = d->scale_mode ? (float*)take(N * D * 4) : nullptr;
      t.dx[m][p] = (float*)take(N * D * 4);
      t.dlv[m][p] = (float*)take(N * D * 4);
    }
  }
  t.ga = (float*)take(N * gw * 4);
  t.gb = (float*)take(N * gw * 4);
  t.rp = (float*)take((int64_t)(1 + d->n_mods) * N * L * 4);
  t.rps = (float*)take((int64_t)d->n_mods * 2 * N * smax * 4);
  t.G = (float*)take(lay->total * 4);
  t.N = N; t.smax = smax;
  if (w) *w = t;
  return off;
}

// ---- kernels -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_index, int64_t row_offset,
                                                      int N, int D, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)N * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / D), d = (int)(i - (int64_t)n * D);
    const int64_t r = row_index ? (int64_t)row_index[row_offset + n] : (int64_t)n;
    out[i] = x[r * D + d];
  }
}

// C[n][o] = act(sum_k A[n][k] W[o][k] + bias[o])                               (kind 0: layer forward)
// C[n][k] = (acc ? C : 0) + sum_o A[n][o] W[o][k], then * (mask[n][k] > 0)      (kind 1: activation gradient)
// C[o][k] = (acc ? C : 0) + sum_n A[n][o] X[n][k]                               (kind 2: weight gradient)
struct GemmArgs {
  const float* A; const float* B; const float* bias; const float* mask; float* C;
  int N, O, K;        // rows, layer outputs, layer inputs
  int kind, relu, acc;
};
__global__ void __launch_bounds__(MOPOE_THREADS) gemm_kernel(GemmArgs g) {
  extern __shared__ __align__(16) float gsm[];
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  float acc[2][2];
  if (g.kind == 0) {
    const int n0 = blockIdx.y * TILE, o0 = blockIdx.x * TILE;
    auto fa = [&](int i, int k) -> float { return (n0 + i < g.N && k < g.K) ? g.A[(int64_t)(n0 + i) * g.K + k] : 0.f; };
    auto fb = [&](int j, int k) -> float { return (o0 + j < g.O && k < g.K) ? g.B[(int64_t)(o0 + j) * g.K + k] : 0.f; };
    tile_gemm<true, true>(fa, fb, g.K, acc, gsm);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n = n0 + 2 * ty + a, o = o0 + 2 * tx + c;
        if (n < g.N && o < g.O) {
          float v = acc[a][c] + (g.bias ? g.bias[o] : 0.f);
          if (g.relu) v = fmaxf(v, 0.f);
          g.C[(int64_t)n * g.O + o] = v;
        }
      }
  } else if (g.kind == 1) {
    const int n0 = blockIdx.y * TILE, k0 = blockIdx.x * TILE;
    auto fa = [&](int i, int o) -> float { return (n0 + i < g.N && o < g.O) ? g.A[(int64_t)(n0 + i) * g.O + o] : 0.f; };
    auto fb = [&](int j, int o) -> float { return (k0 + j < g.K && o < g.O) ? g.B[(int64_t)o * g.K + k0 + j] : 0.f; };
    tile_gemm<true, false>(fa, fb, g.O, acc, gsm);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int n = n0 + 2 * ty + a, k = k0 + 2 * tx + c;
        if (n < g.N && k < g.K) {
          const int64_t idx = (int64_t)n * g.K + k;
          float v = acc[a][c] + (g.acc ? g.C[idx] : 0.f);
          if (g.mask && !(g.mask[idx] > 0.f)) v = 0.f;
          g.C[idx] = v;
        }
      }
  } else {
    const int o0 = blockIdx.y * TILE, k0 = blockIdx.x * TILE;
    auto fa = [&](int i, int n) -> float { return (o0 + i < g.O && n < g.N) ? g.A[(int64_t)n * g.O + o0 + i] : 0.f; };
    auto fb = [&](int j, int n) -> float { return (k0 + j < g.K && n < g.N) ? g.B[(int64_t)n * g.K + k0 + j] : 0.f; };
    tile_gemm<false, false>(fa, fb, g.N, acc, gsm);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int o = o0 + 2 * ty + a, k = k0 + 2 * tx + c;
        if (o < g.O && k < g.K) {
          const int64_t idx = (int64_t)o * g.K + k;
          g.C[idx] = acc[a][c] + (g.acc ? g.C[idx] : 0.f);
        }
      }
  }
}

// out[o] = (acc ? out : 0) + sum_n A[n][o]: 32 columns per CTA, 8 row groups, two-level sums
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int N, int O, float* __restrict__ out, int acc) {
  __shared__ float sm[8][33];
  const int t = threadIdx.x, col = blockIdx.x * 32 + (t & 31), grp = t >> 5;
  float tot = 0.f, run = 0.f;
  if (col < O)
    for (int n = grp, c = 0; n < N; n += 8, ++c) {
      run += A[(int64_t)n * O + col];
      if ((c & 127) == 127) { tot += run; run = 0.f; }
    }
  sm[grp][t & 31] = tot + run;
  __syncthreads();
  if (t < 32 && col < O) {
    float s = 0.f;
    for (int q = 0; q < 8; ++q) s += sm[q][t];
    out[col] = s + (acc ? out[col] : 0.f);
  }
}

struct LatPlan { int e, de, zz, dzz, rp, rps, red, total, hcm, zdm, sm_; };
__host__ __device__ inline LatPlan lat_plan(const ModelView& mv) {
  LatPlan p;
  int hc = 0, zd = 0, s = 1;
  for (int m = 0; m < mv.M; ++m) {
    hc = mv.mod[m].HC > hc ? mv.mod[m].HC : hc;
    zd = mv.mod[m].ZD > zd ? mv.mod[m].ZD : zd;
    s = mv.mod[m].S > s ? mv.mod[m].S : s;
  }
  p.hcm = hc; p.zdm = zd; p.sm_ = s;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 3) & ~3; return o; };
  p.e = take(mv.M * LAT_R * hc); p.de = take(mv.M * LAT_R * hc);
  p.zz = take(mv.M * 2 * LAT_R * zd); p.dzz = take(mv.M * 2 * LAT_R * zd);
  p.rp = take((1 + mv.M) * LAT_R * mv.L); p.rps = take(mv.M * 2 * LAT_R * s);
  p.red = take(MOPOE_N_SCALARS);
  p.total = off;
  return p;
}

// the latent stage of the fused kernels on tiles of LAT_R rows: heads -> subset posteriors, KL sums, mixture owner,
// reparameterised z / style -> decoder inputs (BWD: d decoder inputs -> d heads)
template <bool BWD>
__global__ void __launch_bounds__(MOPOE_THREADS) latent_kernel(ModelView mv, StepCtx cx, mopoe_batch_desc b, GWs ws, int64_t eps_base) {
  extern __shared__ __align__(16) float lsm[];
  const LatPlan pl = lat_plan(mv);
  const int t = threadIdx.x, M = mv.M, L = mv.L, N = b.n_rows, present = b.present_mask;
  const int r0 = blockIdx.x * LAT_R, nr = min(LAT_R, N - r0);
  LatSh sh;
  sh.e = lsm + pl.e; sh.de = lsm + pl.de; sh.zz = lsm + pl.zz; sh.dzz = lsm + pl.dzz; sh.rp = lsm + pl.rp; sh.rps = lsm + pl.rps;
  sh.red = lsm + pl.red; sh.R = LAT_R; sh.HCM = pl.hcm; sh.ZDM = pl.zdm; sh.SM_ = pl.sm_; sh.NP = 2;
  if (t < MOPOE_N_SCALARS) sh.red[t] = 0.f;
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const int HC = mv.mod[m].HC, ZD = mv.mod[m].ZD, S = mv.mod[m].S;
    for (int i = t; i < nr * HC; i += MOPOE_THREADS) sh.e[(m * LAT_R + i / HC) * pl.hcm + i % HC] = ws.heads[m][(int64_t)(r0 + i / HC) * HC + i % HC];
    if (BWD) {
      for (int p = 0; p < 2; ++p) {
        for (int i = t; i < nr * ZD; i += MOPOE_THREADS)
          sh.dzz[((m * 2 + p) * LAT_R + i / ZD) * pl.zdm + i % ZD] = (p == 0 || cx.uni_pass) ? ws.dzz[m][((int64_t)p * ws.N + r0 + i / ZD) * ZD + i % ZD] : 0.f;
        for (int i = t; i < nr * S; i += MOPOE_THREADS)
          sh.rps[((m * 2 + p) * LAT_R + i / S) * pl.sm_ + i % S] = ws.rps[(((int64_t)m * 2 + p) * ws.N + r0 + i / S) * ws.smax + i % S];
      }
    }
  }
  if (BWD)
    for (int q = 0; q < 1 + M; ++q)
      for (int i = t; i < nr * L; i += MOPOE_THREADS) sh.rp[(q * LAT_R + i / L) * L + i % L] = ws.rp[((int64_t)q * ws.N + r0 + i / L) * L + i % L];
  __syncthreads();
  if (!BWD) {
    lat_forward(mv, cx, b, eps_base, r0, nr, sh);
    __syncthreads();
    for (int m = 0; m < M; ++m) {
      if (!(present >> m & 1)) continue;
      const int ZD = mv.mod[m].ZD, S = mv.mod[m].S;
      for (int p = 0; p < (cx.uni_pass ? 2 : 1); ++p) {
        for (int i = t; i < nr * ZD; i += MOPOE_THREADS)
          ws.zz[m][((int64_t)p * ws.N + r0 + i / ZD) * ZD + i % ZD] = sh.zz[((m * 2 + p) * LAT_R + i / ZD) * pl.zdm + i % ZD];
        for (int i = t; i < nr * S; i += MOPOE_THREADS)
          ws.rps[(((int64_t)m * 2 + p) * ws.N + r0 + i / S) * ws.smax + i % S] = sh.rps[((m * 2 + p) * LAT_R + i / S) * pl.sm_ + i % S];
      }
    }
    for (int q = 0; q < 1 + M; ++q)
      for (int i = t; i < nr * L; i += MOPOE_THREADS) ws.rp[((int64_t)q * ws.N + r0 + i / L) * L + i % L] = sh.rp[(q * LAT_R + i / L) * L + i % L];
    if (t < MOPOE_N_SCALARS && sh.red[t] != 0.f) atomicAdd(ws.acc + t, (double)sh.red[t]);
  } else {
    lat_backward(mv, cx, b, r0, nr, sh);
    __syncthreads();
    for (int m = 0; m < M; ++m) {
      if (!(present >> m & 1)) continue;
      const int HC = mv.mod[m].HC;
      for (int i = t; i < nr * HC; i += MOPOE_THREADS) ws.de[m][(int64_t)(r0 + i / HC) * HC + i % HC] = sh.de[(m * LAT_R + i / HC) * pl.hcm + i % HC];
    }
  }
}

// element-wise likelihood: -log p summed into the scalar accumulators, d/d loc and d/d log-variance (per element;
// the per-feature Parameter sums them over rows).  normal: Normal(loc, exp(lv / 2)); laplace: Laplace(loc, exp(lv / 2))
__global__ void __launch_bounds__(256) nll_kernel(const float* __restrict__ x, const float* __restrict__ loc, const float* __restrict__ lvs,
                                                  const float* __restrict__ lvp, int N, int D, int likelihood, int slot, float invN,
                                                  float* __restrict__ dx, float* __restrict__ dlv, double* acc) {
  float nll = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)N * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const float lam = lvs ? lvs[i] : lvp[d];
    const float diff = x[i] - loc[i];
    if (likelihood == 0) {
      const float iv = expf(-lam);
      nll += 0.5f * diff * diff * iv + 0.5f * lam + HALF_LOG_2PI;
      dx[i] = -diff * iv * invN;
      dlv[i] = (0.5f - 0.5f * diff * diff * iv) * invN;
    } else {
      const float ib = expf(-0.5f * lam);
      const float a = fabsf(diff) * ib;
      nll += 0.69314718055994531f + 0.5f * lam + a;
      dx[i] = (diff > 0.f ? -ib : (diff < 0.f ? ib : 0.f)) * invN;
      dlv[i] = (0.5f - 0.5f * a) * invN;
    }
  }
  nll = warp_sum(nll);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nll;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int q = 0; q < 8; ++q) s += red[q];
    atomicAdd(acc + slot, (double)s);
  }
}

// gradient store (mode 1) or Adam (mode 2) over the parameter blocks of the PRESENT modalities; torch.optim.Adam
// skips parameters whose gradient is None, so absent modalities keep their moments and step counts
struct ApplyBlocks { int n; int64_t off[96]; int64_t len[96]; int mod[96]; };
__global__ void __launch_bounds__(256) apply_kernel(StepCtx cx, ApplyBlocks blk, const float* __restrict__ G) {
  const int q = blockIdx.y;
  float bc1 = 1.f, bc2s = 1.f;
  if (cx.mode == 2) {
    const float tt = (float)(cx.adam_t[blk.mod[q]] + 1);
    bc1 = 1.f - powf(cx.b1, tt);
    bc2s = sqrtf(1.f - powf(cx.b2, tt));
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < blk.len[q]; i += (int64_t)gridDim.x * blockDim.x)
    apply_grad(cx, blk.off[q] + i, G[blk.off[q] + i], bc1, bc2s);
}

__global__ void gen_finalize_kernel(ModelView mv, StepCtx cx, mopoe_batch_desc b, double* acc, float* out) {
  if (threadIdx.x == 0) {
    if (out) finalize_scalars(mv, cx, b, acc, out);
    if (cx.mode == 2)
      for (int m = 0; m < mv.M; ++m)
        if (b.present_mask >> m & 1) cx.adam_t[m] += 1;
  }
}

// ---- host orchestration --------------------------------------------------------------------------------------
static int launch_gemm(cudaStream_t s, int kind, const float* A, const float* B, const float* bias, const float* mask, float* C,
                       int N, int O, int K, int relu, int acc) {
  GemmArgs g;
  g.A = A; g.B = B; g.bias = bias; g.mask = mask; g.C = C; g.N = N; g.O = O; g.K = K; g.kind = kind; g.relu = relu; g.acc = acc;
  dim3 grid = kind == 0 ? dim3((O + TILE - 1) / TILE, (N + TILE - 1) / TILE)
            : kind == 1 ? dim3((K + TILE - 1) / TILE, (N + TILE - 1) / TILE)
                        : dim3((K + TILE - 1) / TILE, (O + TILE - 1) / TILE);
  gemm_kernel<<<grid, MOPOE_THREADS, 4 * TILE * TLD * 4, s>>>(g);
  MOPOE_CUDA(cudaGetLastError());
  return MOPOE_OK;
}
static int launch_colsum(cudaStream_t s, const float* A, int N, int O, float* out, int acc) {
  colsum_kernel<<<(O + 31) / 32, 256, 0, s>>>(A, N, O, out, acc);
  MOPOE_CUDA(cudaGetLastError());
  return MOPOE_OK;
}
#define GEN_TRY(call) do { int _rc = (call); if (_rc) return _rc; } while (0)

// one step (forward; + backward and gradient application when cx.mode >= 1) of batch `b` (host copy)
static int gen_step(const mopoe_model_desc* desc, const GModel& gm, const ModelView& mv, const StepCtx& cx, const mopoe_batch_desc& b,
                    const GWs& ws, float* params, int64_t eps_base, float* scalars_out, cudaStream_t s) {
  const int M = desc->n_mods, N = b.n_rows, H = MOPOE_HIDDEN, present = b.present_mask;
  const int npass = cx.uni_pass ? 2 : 1;
  const float invN = 1.f / (float)N;
  const bool bwd = cx.mode >= 1;
  MOPOE_CUDA(cudaMemsetAsync(ws.acc, 0, MOPOE_N_SCALARS * sizeof(double), s));
  const LatPlan lp = lat_plan(mv);
  MOPOE_CUDA(cudaFuncSetAttribute(latent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lp.total * 4));
  MOPOE_CUDA(cudaFuncSetAttribute(latent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lp.total * 4));
  // ---- encoders ----
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const GMod& q = gm.mod[m];
    gather_kernel<<<(int)std::min<int64_t>(((int64_t)N * q.D + 255) / 256, 4096), 256, 0, s>>>(cx.x[m], cx.row_index[m], b.row_offset, N, q.D, ws.xg[m]);
    MOPOE_CUDA(cudaGetLastError());
    const float* a = ws.xg[m];
    int in = q.D;
    for (int l = 0; l < gm.He; ++l) {
      GEN_TRY(launch_gemm(s, 0, a, params + q.ew[l], params + q.eb[l], nullptr, ws.ha[m][l], N, H, in, 1, 0));
      a = ws.ha[m][l]; in = H;
    }
    GEN_TRY(launch_gemm(s, 0, a, params + q.wh, params + q.bh, nullptr, ws.heads[m], N, q.HC, in, 0, 0));
    if (cx.out.enc_heads[m]) MOPOE_CUDA(cudaMemcpyAsync(cx.out.enc_heads[m], ws.heads[m], (size_t)N * q.HC * 4, cudaMemcpyDeviceToDevice, s));
  }
  if (cx.heads_only) return MOPOE_OK;
  // ---- latent stage ----
  const int nt = (N + LAT_R - 1) / LAT_R;
  latent_kernel<false><<<nt, MOPOE_THREADS, lp.total * 4, s>>>(mv, cx, b, ws, eps_base);
  MOPOE_CUDA(cudaGetLastError());
  // ---- decoders + likelihood ----
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const GMod& q = gm.mod[m];
    for (int p = 0; p < npass; ++p) {
      const float* a = ws.zz[m] + (int64_t)p * ws.N * q.ZD;
      int in = q.ZD;
      for (int l = 0; l < gm.Hd; ++l) {
        GEN_TRY(launch_gemm(s, 0, a, params + q.dw[l], params + q.db[l], nullptr, ws.hd[m][p][l], N, H, in, 1, 0));
        a = ws.hd[m][p][l]; in = H;
      }
      GEN_TRY(launch_gemm(s, 0, a, params + q.wo, params + q.bo, nullptr, ws.loc[m][p], N, q.D, in, 0, 0));
      if (gm.scale_mode) GEN_TRY(launch_gemm(s, 0, a, params + q.lvw, params + q.lvb, nullptr, ws.lvs[m][p], N, q.D, in, 0, 0));
      if (p == 0 && cx.out.rec_loc[m]) MOPOE_CUDA(cudaMemcpyAsync(cx.out.rec_loc[m], ws.loc[m][0], (size_t)N * q.D * 4, cudaMemcpyDeviceToDevice, s));
      if (p == 0 && gm.scale_mode && cx.out.rec_logvar[m])
        MOPOE_CUDA(cudaMemcpyAsync(cx.out.rec_logvar[m], ws.lvs[m][0], (size_t)N * q.D * 4, cudaMemcpyDeviceToDevice, s));
      if (cx.with_nll) {
        nll_kernel<<<(int)std::min<int64_t>(((int64_t)N * q.D + 255) / 256, 2048), 256, 0, s>>>(
            ws.xg[m], ws.loc[m][p], gm.scale_mode ? ws.lvs[m][p] : nullptr, gm.scale_mode ? nullptr : params + q.lv, N, q.D, gm.likelihood,
            (p == 0 ? MOPOE_S_NLL : MOPOE_S_NLL_UNI) + m, invN, ws.dx[m][p], ws.dlv[m][p], ws.acc);
        MOPOE_CUDA(cudaGetLastError());
      }
    }
  }
  if (bwd) {
    float* G = ws.G;
    // ---- decoders backward ----
    for (int m = 0; m < M; ++m) {
      if (!(present >> m & 1)) continue;
      const GMod& q = gm.mod[m];
      for (int p = 0; p < npass; ++p) {
        const int accp = p > 0;                                   // the unimodal pass adds to the joint pass
        const float* feat = gm.Hd ? ws.hd[m][p][gm.Hd - 1] : ws.zz[m] + (int64_t)p * ws.N * q.ZD;
        GEN_TRY(launch_gemm(s, 2, ws.dx[m][p], feat, nullptr, nullptr, G + q.wo, N, q.D, q.in_o, 0, accp));
        GEN_TRY(launch_colsum(s, ws.dx[m][p], N, q.D, G + q.bo, accp));
        if (gm.scale_mode) {
          GEN_TRY(launch_gemm(s, 2, ws.dlv[m][p], feat, nullptr, nullptr, G + q.lvw, N, q.D, q.in_o, 0, accp));
          GEN_TRY(launch_colsum(s, ws.dlv[m][p], N, q.D, G + q.lvb, accp));
        } else {
          GEN_TRY(launch_colsum(s, ws.dlv[m][p], N, q.D, G + q.lv, accp));
        }
        // gradient of the decoder features (d pre-activation of the last hidden layer when there is one: ReLU mask in
        // the epilogue of the last contribution), then down the hidden layers to the decoder input
        float* dzz = ws.dzz[m] + (int64_t)p * ws.N * q.ZD;
        const float* mask_top = gm.Hd ? ws.hd[m][p][gm.Hd - 1] : nullptr;
        float* gcur = gm.Hd ? ws.ga : dzz;
        float* gnext = ws.gb;
        GEN_TRY(launch_gemm(s, 1, ws.dx[m][p], params + q.wo, nullptr, gm.scale_mode ? nullptr : mask_top, gcur, N, q.D, q.in_o, 0, 0));
        if (gm.scale_mode) GEN_TRY(launch_gemm(s, 1, ws.dlv[m][p], params + q.lvw, nullptr, mask_top, gcur, N, q.D, q.in_o, 0, 1));
        for (int l = gm.Hd - 1; l >= 0; --l) {
          const float* inp = l ? ws.hd[m][p][l - 1] : ws.zz[m] + (int64_t)p * ws.N * q.ZD;
          const int in = l ? H : q.ZD;
          GEN_TRY(launch_gemm(s, 2, gcur, inp, nullptr, nullptr, G + q.dw[l], N, H, in, 0, accp));
          GEN_TRY(launch_colsum(s, gcur, N, H, G + q.db[l], accp));
          float* dst = l ? gnext : dzz;
          GEN_TRY(launch_gemm(s, 1, gcur, params + q.dw[l], nullptr, l ? ws.hd[m][p][l - 1] : nullptr, dst, N, H, in, 0, 0));
          if (l) { gnext = gcur; gcur = dst; }
        }
      }
    }
    // ---- latent stage backward ----
    latent_kernel<true><<<nt, MOPOE_THREADS, lp.total * 4, s>>>(mv, cx, b, ws, eps_base);
    MOPOE_CUDA(cudaGetLastError());
    // ---- encoders backward ----
    for (int m = 0; m < M; ++m) {
      if (!(present >> m & 1)) continue;
      const GMod& q = gm.mod[m];
      const float* feat = gm.He ? ws.ha[m][gm.He - 1] : ws.xg[m];
      GEN_TRY(launch_gemm(s, 2, ws.de[m], feat, nullptr, nullptr, G + q.wh, N, q.HC, q.in_h, 0, 0));
      GEN_TRY(launch_colsum(s, ws.de[m], N, q.HC, G + q.bh, 0));
      if (gm.He) {
        float* gcur = ws.ga;
        float* gnext = ws.gb;
        GEN_TRY(launch_gemm(s, 1, ws.de[m], params + q.wh, nullptr, ws.ha[m][gm.He - 1], gcur, N, q.HC, H, 0, 0));   // masked: d pre-activation
        for (int l = gm.He - 1; l >= 0; --l) {
          const float* inp = l ? ws.ha[m][l - 1] : ws.xg[m];
          const int in = l ? H : q.D;
          GEN_TRY(launch_gemm(s, 2, gcur, inp, nullptr, nullptr, G + q.ew[l], N, H, in, 0, 0));
          GEN_TRY(launch_colsum(s, gcur, N, H, G + q.eb[l], 0));
          if (l) {
            GEN_TRY(launch_gemm(s, 1, gcur, params + q.ew[l], nullptr, ws.ha[m][l - 1], gnext, N, H, H, 0, 0));
            float* tmp = gcur; gcur = gnext; gnext = tmp;
          }
        }
      }
    }
    // ---- gradient store / Adam over the blocks of the present modalities ----
    ApplyBlocks blk;
    blk.n = 0;
    auto add = [&](int64_t off, int64_t len, int m) { if (off >= 0 && len > 0 && blk.n < 96) { blk.off[blk.n] = off; blk.len[blk.n] = len; blk.mod[blk.n] = m; ++blk.n; } };
    for (int m = 0; m < M; ++m) {
      if (!(present >> m & 1)) continue;
      const GMod& q = gm.mod[m];
      for (int l = 0; l < gm.He; ++l) { add(q.ew[l], (int64_t)H * (l ? H : q.D), m); add(q.eb[l], H, m); }
      add(q.wh, (int64_t)q.HC * q.in_h, m); add(q.bh, q.HC, m);
      for (int l = 0; l < gm.Hd; ++l) { add(q.dw[l], (int64_t)H * (l ? H : q.ZD), m); add(q.db[l], H, m); }
      add(q.wo, (int64_t)q.D * q.in_o, m); add(q.bo, q.D, m);
      if (gm.scale_mode) { add(q.lvw, (int64_t)q.D * q.in_o, m); add(q.lvb, q.D, m); }
      else if (desc->learn_output_scale) add(q.lv, q.D, m);
    }
    apply_kernel<<<dim3(64, blk.n), 256, 0, s>>>(cx, blk, G);
    MOPOE_CUDA(cudaGetLastError());
  }
  gen_finalize_kernel<<<1, 32, 0, s>>>(mv, cx, b, ws.acc, scalars_out);
  MOPOE_CUDA(cudaGetLastError());
  return MOPOE_OK;
}

}  // namespace gen
