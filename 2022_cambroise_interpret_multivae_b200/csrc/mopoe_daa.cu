// Digital Avatars Analysis sweep for sm_100a  (workflow.daa_exp, workflow.py:361-537).
//
// Launch plan for one shard of validations (all launches on the caller's stream):
//   1. mopoe_forward(P1+P2) over all n_val*N rows              -> encoder heads (computed ONCE: the
//      reference re-encodes the unchanged "rois" block in every one of its 41 000 forwards)
//   2. daa_base_kernel   one CTA per (validation, subject): the M stochastic reconstructions
//      (workflow.py:388-400; the decoders are affine, so the mean over M passes is taken on z),
//      loc_hat / scale_hat, the sampled scores (workflow.py:401-405) and rois_reconstructions
//   3. daa_avatar_kernel persistent, one CTA per SM, work unit = (validation, subject, score):
//      all n_samples avatars of that unit: perturbed src encoder (rank-1 update of the hidden
//      pre-activation), class heads, subset PoE / mixture owner, reparameterisation, dst decoder,
//      optional write of the avatar tile, and the per-subject OLS slope accumulated in fp64 in the
//      epilogue (stat_utils.py:66-68) -- the avatar tensor never has to be re-read
//   4. daa_stats_kernel  one thread per (validation, score, roi): second-level t-test
//      (stat_utils.py:73-75) or the pooled "fixed" regression (stat_utils.py:62-63)
#include <stdlib.h>
#include <cuda.h>   // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through the runtime, no libcuda link)

#include "mopoe_common.cuh"
#include <mutex>

#include "mopoe_latent.cuh"
#include "mopoe_umma.cuh"

namespace mopoe {

constexpr int RC = 32;          // avatar rows (samples) per chunk
constexpr int CB = 448;         // decoder output columns per block (2 float4 x 56 lanes)
constexpr int DEC_THREADS = 224;  // threads that own decoder micro-tiles (4 row groups x 56)

struct DaaWs {
  float* enc[MOPOE_MAX_MODS];  // (n_val*N, HC_m) encoder heads
  float* scores;               // (n_val, N, C, J)   sampled scores, J contiguous
  float* loc_hat;              // (n_val, N, C)
  double* betas;               // (n_val, C, N, R)  (used when the caller does not want them)
  double* ybar;                // (n_val, C, N, R)  fixed: mean_j y
  double* syy;                 // (n_val, C, N, R)  fixed: sum_j (y - ybar)^2
  double* xstat;               // (n_val, C, N, 2)  xbar, Sxx
  double* sacc;                // (n_val*N*C, tiles per series, 64)  pipelined kernel: sum_j (x - xbar) * z[k]
  int* counter;                // work-unit counter of the persistent kernel
  int* err;                    // device error flag (tcgen05 barrier time-out)
  long long* phase;            // [grid][32] per-role cycle counters of the tcgen05 kernels (profiling builds)
  unsigned char* bsplit;       // fp16 hi/lo operand planes of the decoder / class-head weights (UMMA layout)
  float* mean_eps;             // (n_val*N, 176)  mean noise row of the base passes (two-phase daa_base_kernel)
  float* srec;                 // (n_val*N*C, DAA_SERIES_REC_F)  pipelined kernel: per-series records (daa_series_rec_kernel)
  void* fwd_ws;                // workspace of the encoder forward
  int64_t fwd_ws_bytes;
};

constexpr int DAA_BASE_SC_MAX = 8192;     // sampled scores of one subject kept in shared memory by daa_base_kernel (floats)
constexpr int DAA_SERIES_REC_F = 2 * MOPOE_HIDDEN + 128 + 4;   // == PK_REC_F (mopoe_daa_pipe.cuh)

static int64_t daa_carve(const mopoe_model_desc* d, const mopoe_daa_desc* q, char* base, DaaWs* w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += (bytes + 255) & ~(int64_t)255; return base ? base + o : (char*)nullptr; };
  const int64_t rows = (int64_t)q->n_val * q->n_subjects;
  const int C = d->dims[q->src_mod], R = d->dims[q->dst_mod];
  DaaWs t;
  memset(&t, 0, sizeof(t));
  for (int m = 0; m < d->n_mods; ++m) t.enc[m] = (float*)take(rows * (2 * d->latent_dim + 2 * d->style_dims[m]) * 4);
  t.scores = (float*)take(rows * C * q->n_samples * 4);
  t.loc_hat = (float*)take(rows * C * 4);
  t.betas = (double*)take(rows * C * R * 8);
  if (q->reg_method == 1) {
    t.ybar = (double*)take(rows * C * R * 8);
    t.syy = (double*)take(rows * C * R * 8);
  }
  t.xstat = (double*)take(rows * C * 2 * 8);
  t.sacc = (double*)take(rows * C * ((q->n_samples + 126) / 128 + 1) * 64 * 8);
  t.counter = (int*)take(256);
  t.err = (int*)take(256);
  t.phase = (long long*)take(256 * 32 * 8);
  t.bsplit = (unsigned char*)take(2 * (480 * 64 * 2) + 2 * (64 * 256 * 2));
  t.srec = (float*)take(rows * C * DAA_SERIES_REC_F * 4);
  t.mean_eps = (float*)take(rows * 176 * 4);
  t.fwd_ws_bytes = mopoe_workspace_bytes(d, rows);
  t.fwd_ws = take(t.fwd_ws_bytes);
  if (w) *w = t;
  return off;
}

struct DaaCtx {
  mopoe_daa_desc q;
  mopoe_batch_desc b;   // n_rows = n_subjects, all modalities present
  const float* x[MOPOE_MAX_MODS];
  Noise nz_base, nz_score, nz_av;
  int v_base_off, v_score_off, v_av_off;  // validation index offset inside each noise index space
  float* avatars; float* sampled_scores; float* recon;
  double* betas;
  int C, R, J, N;
  int make_rec;        // daa_base_kernel also writes the per-series records of the pipelined kernel
};

// Latent noise of the DAA streams is addressed by ROW: injected tensors hold E columns per row; the
// generator draws row `ridx` as EP/4 whole Philox blocks, block b of the row = counter ridx*(EP/4)+b,
// laid out [content | style_0 | style_1 ...] with every section padded to a multiple of 4
// (oracle/philox.py: philox_rows).  fill dst[0..E) with row `ridx` using `nthr` cooperating threads.
__device__ __forceinline__ void fill_noise_row(const ModelView& mv, const Noise& nz, int64_t ridx, float* dst, int tid, int nthr) {
  if (nz.eps) {
    for (int e = tid; e < mv.E; e += nthr) dst[e] = nz.eps[ridx * mv.E + e];
  } else {
    const int nb = mv.EP >> 2;
    for (int b = tid; b < nb; b += nthr) {
      float v[4];
      philox_normal4(nz, (uint64_t)(ridx * nb + b), v);
      const int pe = 4 * b;
      int e0 = pe, lim = mv.L;               // content section
      for (int m = 0; m < mv.M; ++m)
        if (pe >= mv.mod[m].peps_off) { e0 = mv.mod[m].eps_off + (pe - mv.mod[m].peps_off); lim = mv.mod[m].eps_off + mv.mod[m].S; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (e0 + i < lim) dst[e0 + i] = v[i];
    }
  }
}

// -------------------------------------------------------------------------------------------
// 2. base passes + score sampling: one CTA per (validation, subject)
// -------------------------------------------------------------------------------------------
constexpr int BASE_THREADS = 128;   // 8+ CTAs per SM: the (validation, subject) grid fits in one wave

// Per-series records of the pipelined avatar kernel (mopoe_daa_pipe.cuh), computed by the CTA that owns the
// (validation, subject) row at the end of daa_base_kernel: hidden pre-activation of the src encoder WITHOUT the
// perturbed column (a0 = b1 + sum_{k != c} W1[:,k] x[k]) and that column of W1 (the rank-1 direction), the
// posterior partial sums of the row's mixture owner without the src expert (mm_div.py:13-20), the dst style
// posterior.  (xbar | need_src are appended where xbar is computed.)  Built inside the persistent kernel by
// its aux warp this was a chain of L2 round trips per tile on ONE warp -- busy 84 % of the launch, every
// producer waiting for it at the tile barrier; the aux warp now only copies the 2.5 KB record.
__device__ bool series_owner(const ModelView& mv, const DaaCtx& cx, int g, int& s_own) {
  int owner = 0, kidx = 0;
  s_own = 0;
  for (int k = 0; k < cx.b.n_mix; ++k)
    if (g >= cx.b.joint_bounds[k] && g < cx.b.joint_bounds[k + 1]) owner = k;
  if (prior_component(mv, cx.b, owner)) { s_own = -1; return false; }   // jsd: the row samples z from the prior N(0, I)
  for (int s = 0; s < mv.sub.n_subsets; ++s) {
    if (!in_mixture(mv, cx.b, s)) continue;
    if (kidx == owner) s_own = s;
    ++kidx;
  }
  return ((mv.sub.mask[s_own] >> cx.q.src_mod) & 1) || (moe_like(mv) && mv.sub.n_members[s_own] > 1);
}

__device__ void series_records(const ModelView& mv, const DaaCtx& cx, const DaaWs& ws, int64_t row, int g, int t, int nthreads,
                               int parts /* 1: hidden pre-activations (inputs only) | 2: posterior parts (encoder heads) */) {
  const int src = cx.q.src_mod, dst = cx.q.dst_mod;
  const ModView& ms = mv.mod[src];
  const ModView& mdst = mv.mod[dst];
  const int L = mv.L, M = mv.M, Sd = mdst.S, C = cx.C;
  int so;
  series_owner(mv, cx, g, so);
  float* rec0 = ws.srec + row * C * DAA_SERIES_REC_F;
  if (parts & 1) {
  const float* xs = cx.x[src] + row * C;
  constexpr int MAXC = 16;                              // the pipelined kernel is selected for C <= UM_MAXC = 16 only
  float xr[MAXC];
#pragma unroll
  for (int k = 0; k < MAXC; ++k) xr[k] = k < C ? xs[k] : 0.f;
  for (int h = t; h < MOPOE_HIDDEN; h += nthreads) {
    const float* w = ms.w1 + (int64_t)h * C;
    float wk[MAXC];
#pragma unroll
    for (int k = 0; k < MAXC; ++k) wk[k] = k < C ? w[k] : 0.f;
    const float b = ms.b1[h];
    for (int uc = 0; uc < C; ++uc) {
      float a = b, wc = 0.f;
#pragma unroll
      for (int k = 0; k < MAXC; ++k) {
        if (k < C) a = (k == uc) ? a : fmaf(wk[k], xr[k], a);
        wc = (k == uc) ? wk[k] : wc;
      }
      rec0[uc * DAA_SERIES_REC_F + h] = a;
      rec0[uc * DAA_SERIES_REC_F + MOPOE_HIDDEN + h] = wc;
    }
  }
  }
  if (!(parts & 2)) return;
  for (int i = t; i < 128; i += nthreads) {
    float val = 0.f;
    const int sec = i >> 5, k = i & 31;                 // sections: other precisions | other mu*T | style mu | style sd
    if (sec < 2 && k < L) {
      float A = 0.f, B = 0.f;
      if (so < 0) {                                     // owned by the prior component (jsd): finished posterior (0, 1)
        A = 0.f; B = 1.f;
      } else if (moe_like(mv)) {
        const int m = mv.sub.members[so][0];
        A = ws.enc[m][row * mv.mod[m].HC + k];
        B = expf(0.5f * ws.enc[m][row * mv.mod[m].HC + L + k]);
      } else {
        const int nm = mv.sub.n_members[so];
        for (int q = 0; q < nm; ++q) {
          const int m = mv.sub.members[so][q];
          if (m == src) continue;
          const float T = 1.f / (expf(ws.enc[m][row * mv.mod[m].HC + L + k]) + MOPOE_POE_EPS);
          A += T; B += ws.enc[m][row * mv.mod[m].HC + k] * T;
        }
        if (mv.method == MOPOE_METHOD_POE || nm == M) A += 1.f / (1.f + MOPOE_POE_EPS);
        if (!((mv.sub.mask[so] >> src) & 1)) {         // owner without src: finished posterior (mu, sd)
          const float mu = B / A, lv = logf(1.f / A);
          A = mu; B = expf(0.5f * lv);
        }
      }
      val = sec == 0 ? A : B;
    } else if (sec >= 2 && k < Sd) {
      val = sec == 2 ? ws.enc[dst][row * mdst.HC + 2 * L + k] : expf(0.5f * ws.enc[dst][row * mdst.HC + 2 * L + Sd + k]);
    }
    for (int uc = 0; uc < C; ++uc) rec0[uc * DAA_SERIES_REC_F + 2 * MOPOE_HIDDEN + i] = val;
  }
}

// phase 0: whole kernel; phase 1: only the mean noise row of the n_base passes (needs no encoder output: runs
// concurrently with the encoder kernels on a second stream) -> ws.mean_eps; phase 2: everything after it
__global__ void __launch_bounds__(BASE_THREADS) daa_base_kernel(ModelView mv, DaaCtx cx, DaaWs ws, int phase) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int v = blockIdx.x / cx.N, g = blockIdx.x % cx.N;
  const int M = mv.M, L = mv.L, E = mv.E;
  const int64_t row = (int64_t)v * cx.N + g;
  float* s_mean = sm;                 // [E] mean eps
  float* s_acc = s_mean + 176;        // [threads / nb][EP] per-group sums (block-padded columns)
  float* s_zz = s_acc + 4 * BASE_THREADS;         // [M][64] decoder inputs
  float* s_loc = s_zz + MOPOE_MAX_MODS * 64;  // [C]
  float* s_sc = s_loc + 64;                   // [C][J] sampled scores (when they fit: DAA_BASE_SC_MAX floats)
  const bool sc_smem = cx.J * cx.C <= DAA_BASE_SC_MAX;
  // records of the pipelined avatar kernel first: they depend on the encoder heads and the inputs only, and
  // their load -> store chains then overlap the Philox loop of the co-resident CTAs
  // (the error flag and the tile counter of the sweep are reset here rather than by memset nodes: each node on the
  // stream is a couple of microseconds of the sweep's critical path)
  if (blockIdx.x == 0 && t == 0 && phase != 1) *ws.err = 0;
  if (cx.make_rec) {
    if (blockIdx.x == 0 && t == 0) *ws.counter = 0;    // tile counter of the pipelined kernel's dynamic schedule
    series_records(mv, cx, ws, row, g, t, BASE_THREADS, phase == 0 ? 3 : phase);
  }
  if (phase == 2) {
    for (int e = t; e < E; e += BASE_THREADS) s_mean[e] = ws.mean_eps[row * 176 + e];
  } else  // mean over the n_base passes of the noise row of this subject.  Thread = (Philox block b of the row,
  // pass group): every lane draws whole blocks, four independent passes in flight per thread.
  {
    const int nb = mv.EP >> 2, ngrp = BASE_THREADS / nb;
    const int b = t % nb, grp = t / nb;
    int e0 = 4 * b, lim = mv.L;           // unpadded column of the block's first element, end of its section
    for (int m = 0; m < M; ++m)
      if (4 * b >= mv.mod[m].peps_off) { e0 = mv.mod[m].eps_off + (4 * b - mv.mod[m].peps_off); lim = mv.mod[m].eps_off + mv.mod[m].S; }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    // base_mode 1: the mean row is drawn directly (one row per subject, scaled by 1/sqrt(M) below)
    const int n_draw = cx.q.base_mode == 1 ? 1 : cx.q.n_base;
    if (grp < ngrp) {
      const int64_t r0 = (int64_t)(cx.v_base_off + v) * n_draw * cx.N + g;   // noise row of pass 0
      if (cx.nz_base.eps) {
        for (int p = grp; p < n_draw; p += ngrp) {
          const float* ep = cx.nz_base.eps + (r0 + (int64_t)p * cx.N) * E;
          if (e0 + 0 < lim) a0 += ep[e0 + 0];
          if (e0 + 1 < lim) a1 += ep[e0 + 1];
          if (e0 + 2 < lim) a2 += ep[e0 + 2];
          if (e0 + 3 < lim) a3 += ep[e0 + 3];
        }
      } else {
#pragma unroll 4
        for (int p = grp; p < n_draw; p += ngrp) {
          float x[4];
          philox_normal4(cx.nz_base, (uint64_t)((r0 + (int64_t)p * cx.N) * nb + b), x);
          a0 += x[0]; a1 += x[1]; a2 += x[2]; a3 += x[3];
        }
      }
      float* o = s_acc + grp * mv.EP + 4 * b;
      o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
    }
    __syncthreads();
    if (t < mv.EP) {
      const int bb = t >> 2, i = t & 3;
      int ee = 4 * bb, ll = mv.L;
      for (int m = 0; m < M; ++m)
        if (4 * bb >= mv.mod[m].peps_off) { ee = mv.mod[m].eps_off + (4 * bb - mv.mod[m].peps_off); ll = mv.mod[m].eps_off + mv.mod[m].S; }
      if (ee + i < ll) {
        float a = 0.f;
        for (int q = 0; q < ngrp; ++q) a += s_acc[q * mv.EP + t];
        s_mean[ee + i] = cx.q.base_mode == 1 ? a * rsqrtf((float)cx.q.n_base) : a / (float)cx.q.n_base;
        if (phase == 1) ws.mean_eps[row * 176 + ee + i] = s_mean[ee + i];
      }
    }
  }
  if (phase == 1) return;
  __syncthreads();
  // joint posterior of this row (sampling semantics: the row's mixture owner), mean latent
  if (t < L) {
    const int l = t;
    float mu_e[MOPOE_MAX_MODS], lv_e[MOPOE_MAX_MODS];
#pragma unroll
    for (int m = 0; m < MOPOE_MAX_MODS; ++m) {
      mu_e[m] = m < M ? ws.enc[m][row * mv.mod[m].HC + l] : 0.f;
      lv_e[m] = m < M ? ws.enc[m][row * mv.mod[m].HC + L + l] : 0.f;
    }
    int owner = 0, kidx = 0;
    for (int k = 0; k < cx.b.n_mix; ++k)
      if (g >= cx.b.joint_bounds[k] && g < cx.b.joint_bounds[k + 1]) owner = k;
    float jmu = 0.f, jlv = 0.f;
    for (int s = 0; s < mv.sub.n_subsets; ++s) {
      if (!in_mixture(mv, cx.b, s)) continue;
      if (kidx == owner) { const SubsetEval ev = eval_subset(mv, cx.b, s, g, mu_e, lv_e); jmu = ev.mu; jlv = ev.lv; }
      ++kidx;
    }
    const float z = jmu + expf(0.5f * jlv) * s_mean[l];
    for (int m = 0; m < M; ++m) s_zz[m * 64 + mv.mod[m].S + l] = z;
  }
  for (int m = 0; m < M; ++m) {
    const ModView& md = mv.mod[m];
    if (t < md.S) {
      const float mu = ws.enc[m][row * md.HC + 2 * L + t], lv = ws.enc[m][row * md.HC + 2 * L + md.S + t];
      s_zz[m * 64 + t] = mu + expf(0.5f * lv) * s_mean[md.eps_off + t];
    }
  }
  __syncthreads();
  // decode src -> loc_hat, dst -> reconstruction   (affine decoders: mean over passes == decode of mean z);
  // thread per output row, its ZD weights read as float4 (rows are contiguous)
  {
    const ModView ms = mv.mod[cx.q.src_mod];
    const ModView mdst = mv.mod[cx.q.dst_mod];
    auto decode = [&](const ModView& md, const float* z, int r) -> float {
      const float* w = md.wd + (int64_t)r * md.ZD;
      float a = md.bd[r];
      if ((md.ZD & 3) == 0) {
        for (int k = 0; k < md.ZD; k += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(w + k);
          a = fmaf(z[k], wv.x, a); a = fmaf(z[k + 1], wv.y, a); a = fmaf(z[k + 2], wv.z, a); a = fmaf(z[k + 3], wv.w, a);
        }
      } else {
        for (int k = 0; k < md.ZD; ++k) a = fmaf(z[k], w[k], a);
      }
      return a;
    };
    for (int c = t; c < cx.C; c += BASE_THREADS) {
      const float a = decode(ms, s_zz + cx.q.src_mod * 64, c);
      s_loc[c] = a;
      ws.loc_hat[row * cx.C + c] = a;
    }
    if (cx.recon)
      for (int r = t; r < cx.R; r += BASE_THREADS) cx.recon[row * cx.R + r] = decode(mdst, s_zz + cx.q.dst_mod * 64, r);
  }
  __syncthreads();
  // scores[j][c] = loc_hat[c] + scale_hat[c] * eps   (Normal(loc_hat, scale_hat).sample, workflow.py:401-405)
  const ModView& ms = mv.mod[cx.q.src_mod];
  for (int i = t; i < cx.J * cx.C; i += BASE_THREADS) {
    const int j = i / cx.C, c = i % cx.C;
    const int64_t idx = (((int64_t)(cx.v_score_off + v) * cx.J + j) * cx.N + g) * cx.C + c;
    const float e = cx.nz_score.at(idx);
    const float s = cx.q.score_mode ? e : s_loc[c] + expf(0.5f * ms.lv[c]) * e;      // given values | Normal sample
    ws.scores[(row * cx.C + c) * cx.J + j] = s;
    if (sc_smem) s_sc[c * cx.J + j] = s;
    if (cx.sampled_scores) cx.sampled_scores[(row * cx.J + j) * cx.C + c] = s;
  }
  __syncthreads();
  // xbar, Sxx of every (subject, score) series in fp64 (centred OLS: slope = Sxy / Sxx)
  for (int c = warp; c < cx.C; c += BASE_THREADS / 32) {
    const float* sx = sc_smem ? s_sc + c * cx.J : ws.scores + (row * cx.C + c) * cx.J;   // (an L2 round trip per load otherwise)
    double a = 0.0;
    for (int j = lane; j < cx.J; j += 32) a += (double)sx[j];
    a = warp_sum(a);
    const double xb = a / (double)cx.J;
    double q = 0.0;
    for (int j = lane; j < cx.J; j += 32) { const double d = (double)sx[j] - xb; q += d * d; }
    q = warp_sum(q);
    if (lane == 0) {
      const int64_t o = (((int64_t)v * cx.C + c) * cx.N + g) * 2;
      ws.xstat[o] = xb; ws.xstat[o + 1] = q;
      if (cx.make_rec) {
        int so;
        float* tail = ws.srec + (row * cx.C + c) * DAA_SERIES_REC_F + 2 * MOPOE_HIDDEN + 128;
        *reinterpret_cast<double*>(tail) = xb;
        reinterpret_cast<int*>(tail)[2] = series_owner(mv, cx, g, so) ? 1 : 0;
        tail[3] = 0.f;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// 3. avatars + per-subject regression: persistent kernel, unit = (validation, subject, score)
// -------------------------------------------------------------------------------------------
struct AvSmem {
  int wdT, bd, whT, bh, w1c, a0, hT, part, zzT, e, eps, sx, oth, misc, total;  // float offsets
};

__host__ __device__ inline AvSmem av_plan(const ModelView& mv, int src, int dst, int J) {
  AvSmem p;
  const int ZD = mv.mod[dst].ZD, HO = 2 * mv.L;
  int off = 0;
  auto take = [&](int n) { int o = off; off += (n + 3) & ~3; return o; };
  p.wdT = take(ZD * CB);
  p.bd = take(CB);
  p.whT = take(MOPOE_HIDDEN * HO);
  p.bh = take(HO);
  p.w1c = take(MOPOE_HIDDEN);
  p.a0 = take(MOPOE_HIDDEN);
  p.hT = take(MOPOE_HIDDEN * RC);
  const int part = 8 * RC * HO, stat = 2 * 4 * CB * 3;  // partial head sums | fp64 stats reduction
  p.part = take(part > stat ? part : stat);
  p.zzT = take(ZD * RC);
  p.e = take(RC * HO);
  p.eps = take(RC * mv.E);
  p.sx = take(J);
  p.oth = take(MOPOE_MAX_MODS * 2 * 32 + 2 * 32);  // other experts (mu, lv)[L], dst style (mu, lv)[S]
  p.misc = take(64);
  p.total = off;
  return p;
}

template <bool FIXED>
__global__ void __launch_bounds__(MOPOE_THREADS, 1) daa_avatar_kernel(ModelView mv, DaaCtx cx, DaaWs ws, int col0) {
  extern __shared__ __align__(16) float sm[];
  const AvSmem pl = av_plan(mv, cx.q.src_mod, cx.q.dst_mod, cx.J);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int src = cx.q.src_mod, dst = cx.q.dst_mod;
  const ModView& ms = mv.mod[src];
  const ModView& mdst = mv.mod[dst];
  const int L = mv.L, M = mv.M, E = mv.E, HO = 2 * L, ZD = mdst.ZD, Sd = mdst.S;
  const int C = cx.C, R = cx.R, J = cx.J, N = cx.N;
  const int ncol = min(CB, R - col0);
  float* s_wdT = sm + pl.wdT;
  float* s_bd = sm + pl.bd;
  float* s_whT = sm + pl.whT;
  float* s_bh = sm + pl.bh;
  float* s_w1c = sm + pl.w1c;
  float* s_a0 = sm + pl.a0;
  float* s_hT = sm + pl.hT;
  float* s_part = sm + pl.part;
  float* s_zzT = sm + pl.zzT;
  float* s_e = sm + pl.e;
  float* s_eps = sm + pl.eps;
  float* s_sx = sm + pl.sx;
  float* s_oth = sm + pl.oth;
  float* s_dsty = s_oth + MOPOE_MAX_MODS * 2 * 32;
  double* s_d = reinterpret_cast<double*>(sm + pl.misc);  // [0] xbar [1] Sxx
  int* s_i = reinterpret_cast<int*>(sm + pl.misc + 8);    // [0] unit
  // ---- weights resident in shared memory for the whole launch ----
  for (int i = t; i < ZD * CB; i += MOPOE_THREADS) {
    const int k = i / CB, c = i % CB;
    s_wdT[i] = c < ncol ? mdst.wd[(int64_t)(col0 + c) * ZD + k] : 0.f;
  }
  for (int i = t; i < CB; i += MOPOE_THREADS) s_bd[i] = i < ncol ? mdst.bd[col0 + i] : 0.f;
  for (int i = t; i < MOPOE_HIDDEN * HO; i += MOPOE_THREADS) {
    const int o = i / MOPOE_HIDDEN, k = i % MOPOE_HIDDEN;  // coalesced read of wh rows
    s_whT[k * HO + o] = ms.wh[(int64_t)o * MOPOE_HIDDEN + k];
  }
  for (int i = t; i < HO; i += MOPOE_THREADS) s_bh[i] = ms.bh[i];
  const int n_units = cx.q.n_val * N * C;
  const int rg = t / 56, cg = t % 56;  // decoder micro-tile owner (t < 224)
  while (true) {
    __syncthreads();
    if (t == 0) s_i[0] = atomicAdd(ws.counter, 1);
    __syncthreads();
    const int unit = s_i[0];
    if (unit >= n_units) break;
    const int c = unit % C, g = (unit / C) % N, v = unit / (C * N);
    const int64_t row = (int64_t)v * N + g;
    // ---- per-unit setup ----
    int owner = 0;
    for (int k = 0; k < cx.b.n_mix; ++k)
      if (g >= cx.b.joint_bounds[k] && g < cx.b.joint_bounds[k + 1]) owner = k;
    int s_own = 0, kidx = 0;
    bool need_src = !cx.q.sample_latents;
    for (int s = 0; s < mv.sub.n_subsets; ++s) {
      if (!in_mixture(mv, cx.b, s)) continue;
      if (kidx == owner) s_own = s;
      ++kidx;
    }
    const bool own_prior = prior_component(mv, cx.b, owner);   // jsd: the row samples z from the prior N(0, I)
    if (cx.q.sample_latents) {
      need_src = (mv.sub.mask[s_own] >> src) & 1;
      if (moe_like(mv) && mv.sub.n_members[s_own] > 1) need_src = true;
      if (own_prior) need_src = false;
    }
    {  // hidden pre-activation without the perturbed column, and that column of W1
      const float* w = ms.w1 + (int64_t)t * C;
      const float* xr = cx.x[src] + row * C;
      float a = ms.b1[t];
      for (int i = 0; i < C; ++i) a = (i == c) ? a : fmaf(w[i], xr[i], a);
      s_a0[t] = a;
      s_w1c[t] = w[c];
    }
    for (int j = t; j < J; j += MOPOE_THREADS) s_sx[j] = ws.scores[(row * C + c) * J + j];
    for (int i = t; i < M * L; i += MOPOE_THREADS) {
      const int m = i / L, l = i % L;
      s_oth[(m * 2 + 0) * 32 + l] = ws.enc[m][row * mv.mod[m].HC + l];
      s_oth[(m * 2 + 1) * 32 + l] = ws.enc[m][row * mv.mod[m].HC + L + l];
    }
    if (t < Sd) {
      s_dsty[t] = ws.enc[dst][row * mdst.HC + 2 * L + t];
      s_dsty[32 + t] = expf(0.5f * ws.enc[dst][row * mdst.HC + 2 * L + Sd + t]);
    }
    __syncthreads();
    if (warp == 0) {  // xbar, Sxx in fp64 (centred: stat_utils OLS slope = Sxy / Sxx)
      double sx = 0.0;
      for (int j = lane; j < J; j += 32) sx += (double)s_sx[j];
      sx = warp_sum(sx);
      const double xb = sx / (double)J;
      double sxx = 0.0;
      for (int j = lane; j < J; j += 32) { const double d = (double)s_sx[j] - xb; sxx += d * d; }
      sxx = warp_sum(sxx);
      if (lane == 0) { s_d[0] = xb; s_d[1] = sxx; ws.xstat[(((int64_t)v * C + c) * N + g) * 2] = xb; ws.xstat[(((int64_t)v * C + c) * N + g) * 2 + 1] = sxx; }
    }
    __syncthreads();
    const double xbar = s_d[0], sxx_d = s_d[1];
    double st_xy[8], st_y[8], st_yy[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) st_xy[i] = st_y[i] = st_yy[i] = 0.0;

    for (int j0 = 0; j0 < J; j0 += RC) {
      const int nrow = min(RC, J - j0);
      // noise rows of this chunk
      if (cx.q.sample_latents) {
        const int jj = t >> 3, sub = t & 7;  // 8 threads per row
        if (jj < nrow) {
          const int64_t ridx = (((int64_t)(cx.v_av_off + v) * J + j0 + jj) * C + c) * N + g;
          fill_noise_row(mv, cx.nz_av, ridx, s_eps + jj * E, sub, 8);
        }
      }
      if (need_src) {
        // B1: hidden layer of the perturbed src row:  h = relu(a0 + W1[:,c] * score)
        {  // warp w owns hidden units 32w..32w+31, lanes walk the rows (conflict-free stores)
          const float sc = lane < nrow ? s_sx[j0 + lane] : 0.f;
#pragma unroll 8
          for (int kk = 0; kk < 32; ++kk) {
            const int k = warp * 32 + kk;
            s_hT[k * RC + lane] = fmaxf(fmaf(s_w1c[k], sc, s_a0[k]), 0.f);
          }
        }
        __syncthreads();
        // B2: class heads, K split over the 8 warps (32 hidden units each); lane = 4 row groups x 8 col groups
        {
          const int hrg = lane >> 3, hcg = lane & 7;
          const int cpl = (HO + 7) >> 3;  // columns per lane (<= 8)
          float acc[8][8];
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int bq = 0; bq < 8; ++bq) acc[a][bq] = 0.f;
          for (int kk = 0; kk < 32; ++kk) {
            const int k = warp * 32 + kk;
            const float4 h0 = *reinterpret_cast<const float4*>(s_hT + k * RC + hrg * 8);
            const float4 h1 = *reinterpret_cast<const float4*>(s_hT + k * RC + hrg * 8 + 4);
            const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int bq = 0; bq < 8; ++bq) {
              if (bq < cpl) {
                const int o = hcg * cpl + bq;
                const float w = o < HO ? s_whT[k * HO + o] : 0.f;
#pragma unroll
                for (int a = 0; a < 8; ++a) acc[a][bq] = fmaf(hv[a], w, acc[a][bq]);
              }
            }
          }
#pragma unroll
          for (int bq = 0; bq < 8; ++bq) {
            if (bq < cpl) {
              const int o = hcg * cpl + bq;
              if (o < HO)
#pragma unroll
                for (int a = 0; a < 8; ++a) s_part[(warp * RC + hrg * 8 + a) * HO + o] = acc[a][bq];
            }
          }
        }
        __syncthreads();
        for (int i = t; i < RC * HO; i += MOPOE_THREADS) {
          float a = s_bh[i % HO];
#pragma unroll
          for (int w = 0; w < 8; ++w) a += s_part[w * RC * HO + i];
          s_e[i] = a;
        }
      }
      __syncthreads();
      // B3: subset posterior of the row's mixture owner (or the mixture mean), reparameterise
      for (int i = t; i < RC * L; i += MOPOE_THREADS) {
        const int jj = i / L, l = i % L;
        float mu_e[MOPOE_MAX_MODS], lv_e[MOPOE_MAX_MODS];
#pragma unroll
        for (int m = 0; m < MOPOE_MAX_MODS; ++m) {
          mu_e[m] = m < M ? s_oth[(m * 2 + 0) * 32 + l] : 0.f;
          lv_e[m] = m < M ? s_oth[(m * 2 + 1) * 32 + l] : 0.f;
        }
        if (need_src) { mu_e[src] = s_e[jj * HO + l]; lv_e[src] = s_e[jj * HO + L + l]; }
        float z;
        if (cx.q.sample_latents && own_prior) {
          z = s_eps[jj * E + l];
        } else if (cx.q.sample_latents) {
          const SubsetEval ev = eval_subset(mv, cx.b, s_own, g, mu_e, lv_e);
          z = s_eps[jj * E + l] * expf(0.5f * ev.lv) + ev.mu;
        } else {
          float jmu = 0.f;
          for (int s = 0; s < mv.sub.n_subsets; ++s)
            if (in_mixture(mv, cx.b, s)) jmu += eval_subset(mv, cx.b, s, g, mu_e, lv_e).mu;
          z = jmu / (float)cx.b.n_mix;
        }
        s_zzT[(Sd + l) * RC + jj] = z;
      }
      for (int i = t; i < RC * Sd; i += MOPOE_THREADS) {
        const int jj = i / Sd, s = i % Sd;
        const float e0 = cx.q.sample_latents ? s_eps[jj * E + mdst.eps_off + s] : 0.f;
        s_zzT[s * RC + jj] = s_dsty[s] + s_dsty[32 + s] * e0;
      }
      __syncthreads();
      // B4: dst decoder, 8 rows x 8 columns per thread (224 threads), fp64 regression epilogue
      if (t < DEC_THREADS) {
        float acc[8][8];
        const float4 b0 = *reinterpret_cast<const float4*>(s_bd + cg * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(s_bd + 224 + cg * 4);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          acc[a][0] = b0.x; acc[a][1] = b0.y; acc[a][2] = b0.z; acc[a][3] = b0.w;
          acc[a][4] = b1.x; acc[a][5] = b1.y; acc[a][6] = b1.z; acc[a][7] = b1.w;
        }
#pragma unroll 4
        for (int k = 0; k < ZD; ++k) {
          const float4 z0 = *reinterpret_cast<const float4*>(s_zzT + k * RC + rg * 8);
          const float4 z1 = *reinterpret_cast<const float4*>(s_zzT + k * RC + rg * 8 + 4);
          const float4 w0 = *reinterpret_cast<const float4*>(s_wdT + k * CB + cg * 4);
          const float4 w1 = *reinterpret_cast<const float4*>(s_wdT + k * CB + 224 + cg * 4);
          const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int bq = 0; bq < 8; ++bq) acc[a][bq] = fmaf(zv[a], wv[bq], acc[a][bq]);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const int jj = rg * 8 + a;
          if (jj < nrow) {
            const int j = j0 + jj;
            const double xc = (double)s_sx[j] - xbar;
#pragma unroll
            for (int bq = 0; bq < 8; ++bq) {
              const double y = (double)acc[a][bq];
              st_xy[bq] = fma(xc, y, st_xy[bq]);
              if (FIXED) { st_y[bq] += y; st_yy[bq] = fma(y, y, st_yy[bq]); }
            }
            if (cx.avatars) {
              float* o = cx.avatars + ((((int64_t)v * N + g) * C + c) * J + j) * (int64_t)R + col0;
              const int ca = cg * 4, cb = 224 + cg * 4;
              if ((R & 3) == 0 && ca + 4 <= ncol) *reinterpret_cast<float4*>(o + ca) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
              else for (int q = 0; q < 4; ++q) if (ca + q < ncol) o[ca + q] = acc[a][q];
              if ((R & 3) == 0 && cb + 4 <= ncol) *reinterpret_cast<float4*>(o + cb) = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
              else for (int q = 0; q < 4; ++q) if (cb + q < ncol) o[cb + q] = acc[a][4 + q];
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- unit epilogue: combine the 4 row groups, slope per ROI ----
    double* red = reinterpret_cast<double*>(s_part);  // [3][4][CB]
    if (t < DEC_THREADS) {
#pragma unroll
      for (int bq = 0; bq < 8; ++bq) {
        const int col = (bq < 4 ? cg * 4 + bq : 224 + cg * 4 + bq - 4);
        red[(0 * 4 + rg) * CB + col] = st_xy[bq];
        if (FIXED) { red[(1 * 4 + rg) * CB + col] = st_y[bq]; red[(2 * 4 + rg) * CB + col] = st_yy[bq]; }
      }
    }
    __syncthreads();
    for (int col = t; col < ncol; col += MOPOE_THREADS) {
      const double sxy = red[0 * CB + col] + red[1 * CB + col] + red[2 * CB + col] + red[3 * CB + col];
      const int64_t o = (((int64_t)v * C + c) * N + g) * R + col0 + col;
      const double beta = sxy / sxx_d;
      ws.betas[o] = beta;
      if (FIXED) {
        const double sy = red[(4 + 0) * CB + col] + red[(4 + 1) * CB + col] + red[(4 + 2) * CB + col] + red[(4 + 3) * CB + col];
        const double syy = red[(8 + 0) * CB + col] + red[(8 + 1) * CB + col] + red[(8 + 2) * CB + col] + red[(8 + 3) * CB + col];
        const double yb = sy / (double)J;
        ws.ybar[o] = yb;
        ws.syy[o] = syy - (double)J * yb * yb;
      }
    }
  }
}

}  // namespace mopoe
#include "mopoe_daa_umma.cuh"
#include "mopoe_daa_pipe.cuh"
namespace mopoe {

// -------------------------------------------------------------------------------------------
// statistics on a materialised avatar tensor (mopoe_daa_regression): CTA per (v, g, c)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MOPOE_THREADS) daa_moments_kernel(int N, int C, int J, int R, const float* avatars,
                                                                    const float* scores /*(v,N,J,C)*/, double* betas,
                                                                    double* ybar, double* syy, double* xstat) {
  const int unit = blockIdx.x;
  const int c = unit % C, g = (unit / C) % N, v = unit / (C * N);
  const int t = threadIdx.x, lane = t & 31;
  __shared__ double s_d[2];
  extern __shared__ __align__(16) float s_sx[];
  const int64_t row = (int64_t)v * N + g;
  for (int j = t; j < J; j += MOPOE_THREADS) s_sx[j] = scores[(row * J + j) * C + c];
  __syncthreads();
  if (t < 32) {
    double sx = 0.0;
    for (int j = lane; j < J; j += 32) sx += (double)s_sx[j];
    sx = warp_sum(sx);
    const double xb = sx / (double)J;
    double sxx = 0.0;
    for (int j = lane; j < J; j += 32) { const double d = (double)s_sx[j] - xb; sxx += d * d; }
    sxx = warp_sum(sxx);
    if (lane == 0) { s_d[0] = xb; s_d[1] = sxx; xstat[(((int64_t)v * C + c) * N + g) * 2] = xb; xstat[(((int64_t)v * C + c) * N + g) * 2 + 1] = sxx; }
  }
  __syncthreads();
  const double xb = s_d[0], sxx = s_d[1];
  const float* a = avatars + ((row * C + c) * (int64_t)J) * R;
  for (int r = t; r < R; r += MOPOE_THREADS) {
    double sxy = 0.0, sy = 0.0, syy_ = 0.0;
    for (int j = 0; j < J; ++j) {
      const double y = (double)a[(int64_t)j * R + r];
      sxy = fma((double)s_sx[j] - xb, y, sxy);
      sy += y; syy_ = fma(y, y, syy_);
    }
    const int64_t o = (((int64_t)v * C + c) * N + g) * R + r;
    betas[o] = sxy / sxx;
    if (ybar) { const double yb = sy / (double)J; ybar[o] = yb; syy[o] = syy_ - (double)J * yb * yb; }
  }
}

// -------------------------------------------------------------------------------------------
// Student-t survival function in fp64: 2*sf(|t|, nu) = I_x(nu/2, 1/2), x = nu / (nu + t^2)
// (regularised incomplete beta by Lentz's continued fraction)
// -------------------------------------------------------------------------------------------
__device__ double betacf(double a, double b, double x) {
  const double FPMIN = 1e-300, EPS = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < FPMIN) d = FPMIN;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 500; ++m) {
    const int m2 = 2 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d; h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return h;
}

__device__ double betai(double a, double b, double x, double xc /* 1 - x, accurate */) {
  if (x <= 0.0) return 0.0;
  if (xc <= 0.0) return 1.0;
  const double bt = exp(lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log(xc));
  if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
  return 1.0 - bt * betacf(b, a, xc) / b;
}

// Same continued fraction with ONE fp64 division per half step instead of three (aa = num / den, d = 1 / (1 + aa d)
// and c = 1 + aa / c share the reciprocal of (den + num d)(c den)); the log-beta prefactor depends on nu only
// and is passed in (three lgamma evaluations per statistic otherwise).
// the FPMIN guard of the modified Lentz iteration: a vanishing denominator is replaced by a tiny number
// (the next step corrects it) instead of producing inf / NaN
__device__ __forceinline__ double lentz_guard(double v) { return fabs(v) < 1e-150 ? 1e-150 : v; }
__device__ double betacf_fast(double a, double b, double x) {
  const double EPS = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 / (1.0 - qab * x / qap);
  double h = d;
  for (int m = 1; m <= 500; ++m) {
    const double dm = (double)m, m2 = 2.0 * dm;
    {
      const double num = dm * (b - dm) * x, den = (qam + m2) * (a + m2);
      const double e = lentz_guard(fma(num, d, den)), cd = c * den;
      const double r = 1.0 / (e * cd);
      d = den * cd * r;
      c = lentz_guard(cd + num) * e * r;
      h *= d * c;
    }
    const double num = -(a + dm) * (qab + dm) * x, den = (a + m2) * (qap + m2);
    const double e = lentz_guard(fma(num, d, den)), cd = c * den;
    const double r = 1.0 / (e * cd);
    d = den * cd * r;
    c = lentz_guard(cd + num) * e * r;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < EPS) break;
  }
  return h;
}

__device__ double two_sided_t_pvalue_fast(double tval, double nu, double lg_pref /* lgamma(a+b) - lgamma(a) - lgamma(b), a = nu/2, b = 1/2 */) {
  if (isnan(tval)) return tval;
  if (isinf(tval)) return 0.0;
  const double t2 = tval * tval;
  const double x = nu / (nu + t2), xc = t2 / (nu + t2);
  const double a = 0.5 * nu, b = 0.5;
  if (x <= 0.0) return 0.0;
  if (xc <= 0.0) return 1.0;
  const double bt = exp(lg_pref + a * log(x) + b * log(xc));
  if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf_fast(a, b, x) / a;
  return 1.0 - bt * betacf_fast(b, a, xc) / b;
}

__device__ double two_sided_t_pvalue(double tval, double nu) {
  if (isnan(tval)) return tval;
  if (isinf(tval)) return 0.0;
  const double t2 = tval * tval;
  const double x = nu / (nu + t2), xc = t2 / (nu + t2);
  return betai(0.5 * nu, 0.5, x, xc);
}

// 4. second-level statistics: one thread per (validation, score, roi)
__global__ void daa_stats_kernel(int n_val, int N, int C, int J, int R, int reg_method, const double* betas,
                                 const double* ybar, const double* syy, const double* xstat,
                                 const float* recon, double* coefs, double* pvalues) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_val * C * R) return;
  const int r = (int)(i % R), c = (int)((i / R) % C), v = (int)(i / ((int64_t)R * C));
  const int64_t base = ((int64_t)v * C + c) * N;
  if (reg_method == 0) {
    // hierarchical (stat_utils.py:66-75): OLS "beta ~ 1" == one-sample t-test of the subject slopes
    double mean = 0.0;
    for (int g = 0; g < N; ++g) mean += betas[(base + g) * R + r];
    mean /= (double)N;
    double ss = 0.0;
    for (int g = 0; g < N; ++g) { const double d = betas[(base + g) * R + r] - mean; ss += d * d; }
    const double sd = sqrt(ss / (double)(N - 1));
    const double tval = mean / (sd / sqrt((double)N));
    coefs[i] = mean;
    pvalues[i] = two_sided_t_pvalue(tval, (double)(N - 1));
  } else {
    // fixed (stat_utils.py:62-63, workflow.py:489-492): pooled OLS of (avatar - reconstruction) ~ score
    double xm = 0.0, ym = 0.0;
    for (int g = 0; g < N; ++g) {
      xm += xstat[(base + g) * 2];
      ym += ybar[(base + g) * R + r] - (double)recon[((int64_t)v * N + g) * R + r];
    }
    xm /= (double)N; ym /= (double)N;
    double sxx = 0.0, sxy = 0.0, syy_ = 0.0;
    for (int g = 0; g < N; ++g) {
      const double xb = xstat[(base + g) * 2], sxx_g = xstat[(base + g) * 2 + 1];
      const double yb = ybar[(base + g) * R + r] - (double)recon[((int64_t)v * N + g) * R + r];
      const double sxy_g = betas[(base + g) * R + r] * sxx_g;
      sxx += sxx_g + (double)J * (xb - xm) * (xb - xm);
      sxy += sxy_g + (double)J * (xb - xm) * (yb - ym);
      syy_ += syy[(base + g) * R + r] + (double)J * (yb - ym) * (yb - ym);
    }
    const double n = (double)N * (double)J;
    const double beta = sxy / sxx;
    const double rss = syy_ - beta * sxy;
    const double se = sqrt(rss / (n - 2.0) / sxx);
    coefs[i] = beta;
    pvalues[i] = two_sided_t_pvalue(beta / se, n - 2.0);
  }
}

}  // namespace mopoe

using namespace mopoe;

// 3-D tensor map of the avatar tensor for the TMA epilogue of the pipelined kernel: (column, row inside the
// validation, validation), fp32, 32 x 32 boxes, 128-byte swizzle.  Returns false when the map cannot be used
// (driver entry point missing, unaligned buffer, R not a multiple of 4): the kernel then stores through the LSU.
static bool make_avatar_tmap(float* avatars, int64_t n_val, int64_t rows_per_val, int64_t R, CUtensorMap* out) {
  memset(out, 0, sizeof(*out));
  if (!avatars || (R & 3) || (reinterpret_cast<uintptr_t>(avatars) & 15) || getenv("MOPOE_DAA_NO_TMA")) return false;
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (encode_fn)p;
    else
      cudaGetLastError();
  }
  if (!fn) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)R, (cuuint64_t)rows_per_val, (cuuint64_t)n_val};
  const cuuint64_t strides[2] = {(cuuint64_t)R * 4, (cuuint64_t)rows_per_val * R * 4};
  const cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, avatars, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_profile = 0;
static int g_last_impl = 0;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
// fork/join of the sweep: the encoder kernels run on g_side while the caller's stream draws the base-pass noise
// (one set per device, created under a lock: a process may drive several GPUs, from several host threads)
constexpr int MAX_DEVICES = 64;
struct SideStream { cudaStream_t side = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static SideStream g_sides[MAX_DEVICES];
static std::mutex g_side_mutex;
static int side_stream_of_current_device(SideStream** out) {
  int dev = 0;
  MOPOE_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEVICES) { set_error("device ordinal %d out of range", dev); return MOPOE_EINVAL; }
  std::lock_guard<std::mutex> lock(g_side_mutex);
  SideStream& s = g_sides[dev];
  if (!s.side) {
    MOPOE_CUDA(cudaStreamCreateWithFlags(&s.side, cudaStreamNonBlocking));
    MOPOE_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    MOPOE_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return MOPOE_OK;
}

extern "C" {

int mopoe_profile_enable(int on) {
  g_profile = on;
  if (on && !g_ev0) {
    MOPOE_CUDA(cudaEventCreate(&g_ev0));
    MOPOE_CUDA(cudaEventCreate(&g_ev1));
  }
  return MOPOE_OK;
}

int mopoe_daa_last_kernel_ms(float* ms_out) {
  if (!g_ev0 || !ms_out) { set_error("profiling not enabled"); return MOPOE_EINVAL; }
  MOPOE_CUDA(cudaEventSynchronize(g_ev1));
  MOPOE_CUDA(cudaEventElapsedTime(ms_out, g_ev0, g_ev1));
  return MOPOE_OK;
}

// The sweep reads reconstruction locations and scales only: the likelihood family does not enter it
// (workflow.py:388-419 samples the scores from a Normal whatever the training likelihood was).
#define DAA_DESC_WITHOUT_LIKELIHOOD(desc)   \
  mopoe_model_desc desc##_local;            \
  if (desc) { desc##_local = *desc; desc##_local.likelihood = 0; desc = &desc##_local; }

int64_t mopoe_daa_workspace_bytes(const mopoe_model_desc* desc, const mopoe_daa_desc* daa) {
  DAA_DESC_WITHOUT_LIKELIHOOD(desc)
  if (check_desc(desc)) return MOPOE_EINVAL;
  if (!daa) { set_error("daa is NULL"); return MOPOE_EINVAL; }
  return daa_carve(desc, daa, nullptr, nullptr);
}

static int check_daa(const mopoe_model_desc* d, const mopoe_daa_desc* q) {
  if (q->n_val < 1 || q->n_subjects < 2 || q->n_samples < 3 || q->n_base < 1) {
    set_error("daa sizes invalid (n_val=%d n_subjects=%d n_samples=%d n_base=%d)", q->n_val, q->n_subjects, q->n_samples, q->n_base);
    return MOPOE_EINVAL; }
  if (q->src_mod < 0 || q->src_mod >= d->n_mods || q->dst_mod < 0 || q->dst_mod >= d->n_mods || q->src_mod == q->dst_mod) {
    set_error("src_mod=%d dst_mod=%d invalid", q->src_mod, q->dst_mod); return MOPOE_EINVAL; }
  if (d->n_hidden_enc != 1 || d->n_hidden_dec != 0 || d->scale_mode != 0) {
    set_error("the fused DAA sweep covers num_hidden_layer_encoder=1, num_hidden_layer_decoder=0, per-feature output scale "
              "(its base passes use the affine decoder); other architectures go through daa.daa_sweep_layered"); return MOPOE_EINVAL; }
  if (q->base_mode < 0 || q->base_mode > 1) { set_error("base_mode=%d invalid", q->base_mode); return MOPOE_EINVAL; }
  if (q->reg_method < 0 || q->reg_method > 1) { set_error("reg_method=%d unsupported (hierarchical, fixed; mixed is not on this path)", q->reg_method); return MOPOE_EINVAL; }
  if (d->dims[q->src_mod] > 64) { set_error("src modality wider than 64 columns is unsupported in the DAA kernel"); return MOPOE_EINVAL; }
  if ((q->unit_begin || q->unit_end) &&
      (q->unit_begin < 0 || q->unit_begin >= q->unit_end || (int64_t)q->unit_end > (int64_t)q->n_val * d->dims[q->src_mod])) {
    set_error("unit range [%d, %d) outside the %d x %d (validation, score) units of this call", q->unit_begin, q->unit_end, q->n_val, d->dims[q->src_mod]);
    return MOPOE_EINVAL; }
  int E = d->latent_dim;
  for (int m = 0; m < d->n_mods; ++m) E += d->style_dims[m];
  if (E > 160) { set_error("noise row wider than 160"); return MOPOE_EINVAL; }
  return MOPOE_OK;
}

int mopoe_daa_sweep(const mopoe_model_desc* desc, const float* params, const mopoe_daa_desc* daa,
                    const mopoe_batch_desc* batch, const float* const* x, const float* eps_base,
                    const float* eps_score, const float* eps_av, uint64_t seed, float* avatars,
                    float* sampled_scores, float* reconstructions, double* betas, double* coefs, double* pvalues,
                    void* workspace, int64_t workspace_bytes, void* stream_) {
  DAA_DESC_WITHOUT_LIKELIHOOD(desc)
  int rc = check_desc(desc);
  if (rc) return rc;
  if (mopoe_device_count() == 0) { set_error("no CUDA device: the DAA path has no CPU fallback"); return MOPOE_ENODEV; }
  if (!params || !daa || !batch || !x || !coefs || !pvalues || !workspace) { set_error("NULL argument"); return MOPOE_EINVAL; }
  if ((rc = check_daa(desc, daa))) return rc;
  const int M = desc->n_mods, N = daa->n_subjects;
  if (batch->n_rows != N || batch->present_mask != (1 << M) - 1) { set_error("batch desc must describe n_subjects rows with every modality present"); return MOPOE_EINVAL; }
  if (daa->reg_method == 1 && !reconstructions) { set_error("reg_method fixed needs the reconstructions buffer"); return MOPOE_EINVAL; }
  if (daa->score_mode < 0 || daa->score_mode > 1 || (daa->score_mode == 1 && !eps_score)) { set_error("score_mode=%d invalid (1 needs the score values in eps_score)", daa->score_mode); return MOPOE_EINVAL; }
  if (daa->base_mode == 1 && eps_base) { set_error("base_mode 1 (mean noise row drawn directly) needs the in-kernel generator: eps_base must be NULL"); return MOPOE_EINVAL; }
  const int64_t need = daa_carve(desc, daa, nullptr, nullptr);
  if (workspace_bytes < need) { set_error("workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need); return MOPOE_ENOSPC; }
  DaaWs ws;
  daa_carve(desc, daa, (char*)workspace, &ws);
  if (betas) ws.betas = betas;  // caller wants the per-subject slopes: write them in place
  cudaStream_t stream = (cudaStream_t)stream_;
  mopoe_param_layout lay;
  mopoe_param_layout_of(desc, &lay);
  ModelView mv;
  build_view(desc, &lay, const_cast<float*>(params), &mv);
  DaaCtx cx;
  memset(&cx, 0, sizeof(cx));
  cx.q = *daa; cx.b = *batch;
  for (int m = 0; m < M; ++m) cx.x[m] = x[m];
  cx.nz_base = make_noise(eps_base, seed, MOPOE_STREAM_DAA_BASE);
  cx.nz_score = make_noise(eps_score, seed, MOPOE_STREAM_DAA_SCORE);
  cx.nz_av = make_noise(eps_av, seed, MOPOE_STREAM_DAA_AVATAR);
  cx.v_base_off = eps_base ? 0 : daa->val_begin;
  cx.v_score_off = eps_score ? 0 : daa->val_begin;
  cx.v_av_off = eps_av ? 0 : daa->val_begin;
  cx.avatars = avatars; cx.sampled_scores = sampled_scores; cx.recon = reconstructions; cx.betas = betas;
  cx.C = desc->dims[daa->src_mod]; cx.R = desc->dims[daa->dst_mod]; cx.J = daa->n_samples; cx.N = N;
  // 3. avatars + first-level regression
  const AvSmem pl = av_plan(mv, daa->src_mod, daa->dst_mod, cx.J);
  const int av_smem = pl.total * 4;
  if (av_smem > 227 * 1024) { set_error("DAA avatar kernel needs %d bytes of shared memory (> 227 KB): latent/style dims too large", av_smem); return MOPOE_EINVAL; }
  const int unit_begin = daa->unit_end ? daa->unit_begin : 0, unit_end = daa->unit_end ? daa->unit_end : daa->n_val * cx.C;
  cx.q.unit_begin = unit_begin; cx.q.unit_end = unit_end;
  const int n_units = daa->n_val * N * cx.C;
  const int grid = n_units < num_sms() ? n_units : num_sms();
  // implementation choice (MOPOE_DAA_IMPL=pipe|umma|ffma forces one; the tests cross-check all three):
  //   pipe  warp-specialised tcgen05 pipeline (hierarchical regression, sampled latents)  -> 2
  //   umma  phase-serial tcgen05 kernel (also "fixed" regression / mean latents)            -> 1
  //   ffma  CUDA-core kernel for shapes outside the tcgen05 tilings                          -> 0
  const UmmaDims ud0 = umma_dims(mv, daa->src_mod, daa->dst_mod, CB);
  const int um_smem = umma_plan(mv, daa->src_mod, ud0).total;
  const bool umma_ok = cx.J >= UM_ROWS && cx.C <= UM_MAXC && ud0.NH <= 48 && ud0.KZ <= 64 && um_smem <= 227 * 1024;
  const int pk_smem = pipe_plan(ud0).total;
  const bool pipe_ok = umma_ok && daa->reg_method == 0 && daa->sample_latents && ud0.bias_slot >= 0 && ud0.KZ - ud0.KC <= 32 &&
                       pk_smem <= 227 * 1024 && (int64_t)n_units * cx.J < ((int64_t)1 << 31);
  int impl = pipe_ok ? 2 : (umma_ok ? 1 : 0);
  if (desc->method == MOPOE_METHOD_JSD) impl = pipe_ok ? 2 : 0;   // prior-owned rows: pipelined and CUDA-core kernels only
  const char* force = getenv("MOPOE_DAA_IMPL");
  if (force && !strcmp(force, "ffma")) impl = 0;
  if (force && desc->method == MOPOE_METHOD_JSD && !strcmp(force, "umma")) { set_error("MOPOE_DAA_IMPL=umma: method jsd runs on the pipelined and the CUDA-core avatar kernels"); return MOPOE_EINVAL; }
  if (force && !strcmp(force, "umma")) { if (!umma_ok) { set_error("MOPOE_DAA_IMPL=umma but the shapes do not fit the tcgen05 tiling"); return MOPOE_EINVAL; } impl = 1; }
  if (force && !strcmp(force, "pipe")) { if (!pipe_ok) { set_error("MOPOE_DAA_IMPL=pipe but the configuration does not fit the pipelined kernel"); return MOPOE_EINVAL; } impl = 2; }
  g_last_impl = impl;
  cx.make_rec = impl == 2 ? 1 : 0;
  // 1. encoder heads of every (validation, subject) row, on a second stream: the noise phase of the base
  // passes (most of daa_base_kernel's time) does not depend on them
  SideStream* ss = nullptr;
  if ((rc = side_stream_of_current_device(&ss))) return rc;
  cudaStream_t g_side = ss->side;
  cudaEvent_t g_fork = ss->fork, g_join = ss->join;
  const bool forked = getenv("MOPOE_DAA_NO_FORK") == nullptr;
  if (forked) {
    MOPOE_CUDA(cudaEventRecord(g_fork, stream));
    MOPOE_CUDA(cudaStreamWaitEvent(g_side, g_fork, 0));
  }
  {
    mopoe_batch_desc fb;
    memset(&fb, 0, sizeof(fb));
    fb.n_rows = daa->n_val * N; fb.present_mask = (1 << M) - 1; fb.n_mix = 1;
    fb.joint_bounds[0] = 0; fb.joint_bounds[1] = fb.n_rows;
    for (int k = 1; k <= M; ++k) { fb.moe_bounds[k][0] = 0; for (int i = 1; i <= k; ++i) fb.moe_bounds[k][i] = fb.n_rows; }
    mopoe_forward_out fo;
    memset(&fo, 0, sizeof(fo));
    for (int m = 0; m < M; ++m) fo.enc_heads[m] = ws.enc[m];
    rc = mopoe_forward(desc, params, &fb, x, nullptr, seed, 0, -1, 0, &fo, ws.fwd_ws, ws.fwd_ws_bytes, forked ? (void*)g_side : stream_);
    if (forked && impl == 2 && rc == 0) {   // operand planes of the first column block: weights only, also off the critical path
      const UmmaDims ud = umma_dims(mv, daa->src_mod, daa->dst_mod, cx.R < PK_CBP ? cx.R : PK_CBP);
      daa_umma_prep_kernel<<<64, 256, 0, g_side>>>(mv, daa->src_mod, daa->dst_mod, 0, ud, PK_CBP, ws.bsplit);
      if (cudaGetLastError() != cudaSuccess) rc = MOPOE_ECUDA;
    }
    if (forked) MOPOE_CUDA(cudaEventRecord(g_join, g_side));   // (recorded even on error: the side stream must rejoin a capture)
    if (rc) { if (forked) cudaStreamWaitEvent(stream, g_join, 0); return rc; }
  }
  // 2. base passes
  const int base_smem = (176 + 4 * BASE_THREADS + MOPOE_MAX_MODS * 64 + 64 + (cx.J * cx.C <= DAA_BASE_SC_MAX ? cx.J * cx.C : 0)) * 4;
  if (forked && daa->base_mode == 1) {      // no noise phase worth a launch of its own: one launch behind the encoder heads
                                            // (measured: records in a launch of their own beside the encoder, 0.528 vs 0.526 ms)
    MOPOE_CUDA(cudaStreamWaitEvent(stream, g_join, 0));
    daa_base_kernel<<<daa->n_val * N, BASE_THREADS, base_smem, stream>>>(mv, cx, ws, 0);
  } else if (forked) {
    daa_base_kernel<<<daa->n_val * N, BASE_THREADS, base_smem, stream>>>(mv, cx, ws, 1);
    MOPOE_CUDA(cudaGetLastError());
    MOPOE_CUDA(cudaStreamWaitEvent(stream, g_join, 0));
    daa_base_kernel<<<daa->n_val * N, BASE_THREADS, base_smem, stream>>>(mv, cx, ws, 2);
  } else {
    daa_base_kernel<<<daa->n_val * N, BASE_THREADS, base_smem, stream>>>(mv, cx, ws, 0);
  }
  MOPOE_CUDA(cudaGetLastError());
  if (impl == 2) {
    MOPOE_CUDA(cudaFuncSetAttribute((void*)daa_avatar_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pk_smem));
    CUtensorMap tmap;
    const bool tma_ok = make_avatar_tmap(avatars, daa->n_val, (int64_t)N * cx.C * cx.J, cx.R, &tmap);
    for (int col0 = 0; col0 < cx.R; col0 += PK_CBP) {
      const UmmaDims ud = umma_dims(mv, daa->src_mod, daa->dst_mod, cx.R - col0 < PK_CBP ? cx.R - col0 : PK_CBP);
      if (col0 > 0 || !forked) daa_umma_prep_kernel<<<64, 256, 0, stream>>>(mv, daa->src_mod, daa->dst_mod, col0, ud, PK_CBP, ws.bsplit);
      MOPOE_CUDA(cudaGetLastError());
      if (col0 > 0) MOPOE_CUDA(cudaMemsetAsync(ws.counter, 0, sizeof(int), stream));   // (the first launch's counter is zeroed by daa_base_kernel)
      if (g_profile && col0 == 0) MOPOE_CUDA(cudaEventRecord(g_ev0, stream));
      daa_avatar_pipe_kernel<<<num_sms(), PK_THREADS, pk_smem, stream>>>(mv, cx, ws, col0, tma_ok ? 1 : 0, tmap);
      MOPOE_CUDA(cudaGetLastError());
      if (g_profile && col0 == 0) MOPOE_CUDA(cudaEventRecord(g_ev1, stream));
    }
    // slopes + second-level test of the pipelined path (replaces step 4 below)
    const int bs_nc = cx.R < BS_ROIS ? cx.R : BS_ROIS;
    const int bs_threads = ((bs_nc + 1) / 2 + 31) / 32 * 32 < 256 ? ((bs_nc + 1) / 2 + 31) / 32 * 32 : 256;
    const int bsm = (((ud0.KZ * (bs_nc | 1) + 3) & ~3) + 2 * ud0.KZ * BS_GP) * 4 + BS_GB * 8 + 16;
    MOPOE_CUDA(cudaFuncSetAttribute((void*)daa_beta_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bsm));
    daa_beta_stats_kernel<<<dim3(unit_end - unit_begin, (cx.R + BS_ROIS - 1) / BS_ROIS), bs_threads, bsm, stream>>>(
        mv, daa->dst_mod, cx.R, cx.C, N, cx.J, ud0, unit_begin, ws.sacc, ws.xstat, ws.betas, coefs, pvalues);
    MOPOE_CUDA(cudaGetLastError());
  } else if (impl == 1) {
    void* ufn = daa->reg_method == 1 ? (void*)daa_avatar_umma_kernel<true> : (void*)daa_avatar_umma_kernel<false>;
    MOPOE_CUDA(cudaFuncSetAttribute(ufn, cudaFuncAttributeMaxDynamicSharedMemorySize, um_smem));
    for (int col0 = 0; col0 < cx.R; col0 += CB) {
      const UmmaDims ud = umma_dims(mv, daa->src_mod, daa->dst_mod, cx.R - col0 < CB ? cx.R - col0 : CB);
      daa_umma_prep_kernel<<<64, 256, 0, stream>>>(mv, daa->src_mod, daa->dst_mod, col0, ud, CB, ws.bsplit);
      MOPOE_CUDA(cudaGetLastError());
      if (g_profile && col0 == 0) MOPOE_CUDA(cudaEventRecord(g_ev0, stream));
      if (daa->reg_method == 1) daa_avatar_umma_kernel<true><<<grid, MOPOE_THREADS, um_smem, stream>>>(mv, cx, ws, col0);
      else daa_avatar_umma_kernel<false><<<grid, MOPOE_THREADS, um_smem, stream>>>(mv, cx, ws, col0);
      MOPOE_CUDA(cudaGetLastError());
    }
  } else {
    void* fn = daa->reg_method == 1 ? (void*)daa_avatar_kernel<true> : (void*)daa_avatar_kernel<false>;
    MOPOE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, av_smem));
    for (int col0 = 0; col0 < cx.R; col0 += CB) {
      MOPOE_CUDA(cudaMemsetAsync(ws.counter, 0, sizeof(int), stream));
      if (g_profile && col0 == 0) MOPOE_CUDA(cudaEventRecord(g_ev0, stream));
      if (daa->reg_method == 1) daa_avatar_kernel<true><<<grid, MOPOE_THREADS, av_smem, stream>>>(mv, cx, ws, col0);
      else daa_avatar_kernel<false><<<grid, MOPOE_THREADS, av_smem, stream>>>(mv, cx, ws, col0);
      MOPOE_CUDA(cudaGetLastError());
    }
  }
  if (g_profile && impl != 2) MOPOE_CUDA(cudaEventRecord(g_ev1, stream));
  // 4. second level
  // (the pipelined path honours the unit range; the other two compute every unit of the call)
  const int64_t stat0 = impl == 2 ? (int64_t)unit_begin * cx.R : 0;
  const int64_t nstat = impl == 2 ? (int64_t)(unit_end - unit_begin) * cx.R : (int64_t)daa->n_val * cx.C * cx.R;
  if (impl != 2) {
    daa_stats_kernel<<<(unsigned)((nstat + 127) / 128), 128, 0, stream>>>(daa->n_val, N, cx.C, cx.J, cx.R, daa->reg_method, ws.betas,
                                                                         ws.ybar, ws.syy, ws.xstat, reconstructions, coefs, pvalues);
    MOPOE_CUDA(cudaGetLastError());
  }
  if (impl != 0) {
    // p-values of the pipelined path (its slopes kernel leaves the t statistics in `pvalues`) + error poisoning
    daa_pvalue_kernel<<<(unsigned)((nstat + 127) / 128), 128, 0, stream>>>(ws.err, coefs + stat0, pvalues + stat0, nstat, (double)(N - 1),
                                                                            lgamma(0.5 * (N - 1) + 0.5) - lgamma(0.5 * (N - 1)) - lgamma(0.5), impl == 2 ? 1 : 0);
    MOPOE_CUDA(cudaGetLastError());
  }
  return MOPOE_OK;
}

// ---- multi-GPU: push-style exchange of the association tables over peer memory ---------------------
// Every rank owns a slice (its validations) of the (n_val_total, C, R) fp64 coefs / p-value tables.  Instead of
// an NCCL all_gather launched behind the sweep, the slice is stored straight into EVERY rank's full table through
// peer-mapped pointers (NVLink P2P stores into symmetric memory), followed by a system-scope release of a
// per-source sequence flag; a one-warp kernel on each rank acquires the flags of all sources.  Both kernels are
// ordinary launches on the caller's stream (capturable: the sequence number lives in device memory).
// Exchange buffer layout (bytes): [0] uint64 seq, [8] uint32 block counter, [64 + 8 r] uint64 flag of source rank r,
// [256 ...) two slots (seq parity) x {coefs, pvalues} x elems_total doubles.
constexpr int64_t EX_HDR = 256;
struct ExPeers { unsigned char* base[MOPOE_MAX_PEERS]; };

// root < 0: every rank receives every slice (all-gather); root >= 0: only `root` receives (gather: what daa_exp needs,
// 1/world of the bytes, and no rank but the root ever waits).  In gather mode the senders are kept at most two
// sequence numbers ahead of the root by an acknowledgement flag ([64 + 8 * MOPOE_MAX_PEERS]) the root releases to every
// rank once a step's tables are complete, so a slot is never overwritten before the root has finished with it.
__global__ void __launch_bounds__(256) daa_table_push_kernel(ExPeers peers, int world, int rank, int root, int64_t elems_local,
                                                             int64_t elem_offset, int64_t elems_total, const double* coefs,
                                                             const double* pvalues) {
  unsigned char* own = peers.base[rank];
  const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(own) + 1ull;
  if (root >= 0 && rank != root && seq > 2 && threadIdx.x == 0) {
    const unsigned long long* ack = reinterpret_cast<const unsigned long long*>(own + 64 + 8 * MOPOE_MAX_PEERS);
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ack) : "memory");
      if (v + 2 >= seq || clock64() - t0 > 8000000000LL) break;
      __nanosleep(64);
    }
  }
  __syncthreads();
  const int64_t slot_off = EX_HDR + (int64_t)(seq & 1ull) * 2 * elems_total * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems_local; i += (int64_t)gridDim.x * blockDim.x) {
    const double c = coefs[i], p = pvalues[i];
    for (int r = 0; r < world; ++r) {
      if (root >= 0 && r != root) continue;
      double* dst = reinterpret_cast<double*>(peers.base[r] + slot_off);
      dst[elem_offset + i] = c;
      dst[elems_total + elem_offset + i] = p;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(own + 8);
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {      // last block: every block's stores are fenced
      *counter = 0;
      *reinterpret_cast<volatile unsigned long long*>(own) = seq;
      __threadfence_system();
      for (int r = 0; r < world; ++r) {
        if (root >= 0 && r != root) continue;
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(peers.base[r] + 64 + 8 * rank);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
      }
    }
  }
}

__global__ void daa_table_wait_kernel(ExPeers peers, int world, int rank, int root, int64_t elems_total) {
  unsigned char* own = peers.base[rank];
  const int r = threadIdx.x;
  const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(own);
  if (r < world) {
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(own + 64 + 8 * r);
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
      if (v >= seq) break;
      if (clock64() - t0 > 8000000000LL) {   // a peer never arrived: poison this rank's view instead of hanging
        double* dst = reinterpret_cast<double*>(own + EX_HDR + (int64_t)(seq & 1ull) * 2 * elems_total * 8);
        dst[0] = __longlong_as_double(0x7ff8000000000000LL);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncwarp();
  if (root >= 0 && r < world && r != rank) {     // gather mode: tell the senders that this step's tables are complete
    unsigned long long* ack = reinterpret_cast<unsigned long long*>(peers.base[r] + 64 + 8 * MOPOE_MAX_PEERS);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ack), "l"(seq) : "memory");
  }
}

int mopoe_daa_last_impl(void) { return g_last_impl; }

int64_t mopoe_table_exchange_bytes(int64_t elems_total) {
  if (elems_total < 1) { set_error("elems_total=%lld", (long long)elems_total); return MOPOE_EINVAL; }
  return EX_HDR + 2 * 2 * elems_total * 8;
}

int mopoe_daa_exchange_tables(const mopoe_table_exchange* ex, const double* coefs_local, const double* pvalues_local, void* stream_) {
  if (mopoe_device_count() == 0) { set_error("no CUDA device"); return MOPOE_ENODEV; }
  if (!ex || !coefs_local || !pvalues_local) { set_error("NULL argument"); return MOPOE_EINVAL; }
  if (ex->world < 1 || ex->world > MOPOE_MAX_PEERS || ex->rank < 0 || ex->rank >= ex->world) { set_error("world=%d rank=%d invalid", ex->world, ex->rank); return MOPOE_EINVAL; }
  if (ex->elems_local < 0 || ex->elem_offset < 0 || ex->elem_offset + ex->elems_local > ex->elems_total) { set_error("table slice out of range"); return MOPOE_EINVAL; }
  ExPeers peers;
  for (int r = 0; r < MOPOE_MAX_PEERS; ++r) peers.base[r] = r < ex->world ? (unsigned char*)ex->peer_base[r] : nullptr;
  for (int r = 0; r < ex->world; ++r) if (!peers.base[r]) { set_error("peer_base[%d] is NULL", r); return MOPOE_EINVAL; }
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t n = ex->elems_local;
  const int grid = (int)((n + 255) / 256 < 1 ? 1 : ((n + 255) / 256 > 4 * num_sms() ? 4 * num_sms() : (n + 255) / 256));
  if (ex->root >= ex->world) { set_error("root=%d invalid", ex->root); return MOPOE_EINVAL; }
  const int root = ex->root < 0 ? -1 : ex->root;
  daa_table_push_kernel<<<grid, 256, 0, stream>>>(peers, ex->world, ex->rank, root, n, ex->elem_offset, ex->elems_total, coefs_local, pvalues_local);
  MOPOE_CUDA(cudaGetLastError());
  if (root < 0 || root == ex->rank) {       // receivers only: in gather mode no other rank ever waits
    daa_table_wait_kernel<<<1, 32, 0, stream>>>(peers, ex->world, ex->rank, root, ex->elems_total);
    MOPOE_CUDA(cudaGetLastError());
  }
  return MOPOE_OK;
}

int mopoe_daa_status(const mopoe_model_desc* desc, const mopoe_daa_desc* daa, void* workspace, void* stream_) {
  DAA_DESC_WITHOUT_LIKELIHOOD(desc)
  if (check_desc(desc)) return MOPOE_EINVAL;
  if (!daa || !workspace) { set_error("NULL argument"); return MOPOE_EINVAL; }
  DaaWs ws;
  daa_carve(desc, daa, (char*)workspace, &ws);
  int flag = 0;
  MOPOE_CUDA(cudaMemcpyAsync(&flag, ws.err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
  MOPOE_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
  if (flag) {
    set_error("DAA sweep: a tcgen05 pipeline wait timed out on the device (flag %d); the result tables were poisoned with NaN", flag);
    return MOPOE_EDEVICE;
  }
  return MOPOE_OK;
}

int mopoe_daa_read_phases(const mopoe_model_desc* desc, const mopoe_daa_desc* daa, void* workspace, int64_t* out32_host) {
  DAA_DESC_WITHOUT_LIKELIHOOD(desc)
  if (check_desc(desc)) return MOPOE_EINVAL;
  DaaWs ws;
  daa_carve(desc, daa, (char*)workspace, &ws);
  static long long h[256 * 32];
  MOPOE_CUDA(cudaMemcpy(h, ws.phase, sizeof(h), cudaMemcpyDeviceToHost));
  const int grid = num_sms();
  const int stride = g_last_impl == 2 ? 32 : 8, n = g_last_impl == 2 ? 32 : 8;
  for (int i = 0; i < 32; ++i) out32_host[i] = 0;
  const char* one = getenv("MOPOE_PHASE_CTA");   // one CTA's counters instead of the max over CTAs
  for (int i = 0; i < n; ++i) {
    long long mx = 0;
    for (int b = 0; b < grid && b < 256; ++b) mx = h[b * stride + i] > mx ? h[b * stride + i] : mx;
    out32_host[i] = one ? h[(atoi(one) % grid) * stride + i] : mx;
  }
  return MOPOE_OK;
}

int mopoe_daa_regression(int32_t n_val, int32_t n_subjects, int32_t n_scores, int32_t n_samples, int32_t n_rois,
                         int32_t reg_method, const float* avatars, const float* sampled_scores,
                         const float* reconstructions, double* betas, double* coefs, double* pvalues, void* stream_) {
  if (mopoe_device_count() == 0) { set_error("no CUDA device: the DAA path has no CPU fallback"); return MOPOE_ENODEV; }
  if (!avatars || !sampled_scores || !betas || !coefs || !pvalues) { set_error("NULL argument"); return MOPOE_EINVAL; }
  if (reg_method < 0 || reg_method > 1) { set_error("reg_method=%d unsupported", reg_method); return MOPOE_EINVAL; }
  if (reg_method == 1 && !reconstructions) { set_error("reg_method fixed needs reconstructions"); return MOPOE_EINVAL; }
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t nb = (int64_t)n_val * n_scores * n_subjects * n_rois;
  double* scratch = nullptr;
  // ybar | syy | xstat live in one stream-ordered allocation
  const int64_t bytes = (reg_method == 1 ? 2 * nb : 0) * 8 + (int64_t)n_val * n_scores * n_subjects * 2 * 8;
  MOPOE_CUDA(cudaMallocAsync((void**)&scratch, bytes, stream));
  double* ybar = reg_method == 1 ? scratch : nullptr;
  double* syy = reg_method == 1 ? scratch + nb : nullptr;
  double* xstat = scratch + (reg_method == 1 ? 2 * nb : 0);
  daa_moments_kernel<<<n_val * n_subjects * n_scores, MOPOE_THREADS, n_samples * 4, stream>>>(
      n_subjects, n_scores, n_samples, n_rois, avatars, sampled_scores, betas, ybar, syy, xstat);
  MOPOE_CUDA(cudaGetLastError());
  const int64_t nstat = (int64_t)n_val * n_scores * n_rois;
  daa_stats_kernel<<<(unsigned)((nstat + 127) / 128), 128, 0, stream>>>(n_val, n_subjects, n_scores, n_samples, n_rois, reg_method,
                                                                       betas, ybar, syy, xstat, reconstructions, coefs, pvalues);
  MOPOE_CUDA(cudaGetLastError());
  MOPOE_CUDA(cudaFreeAsync(scratch, stream));
  return MOPOE_OK;
}

}  // extern "C"
