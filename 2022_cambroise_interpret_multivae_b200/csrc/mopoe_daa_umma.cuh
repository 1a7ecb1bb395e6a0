// Tensor-core version of the DAA avatar kernel (included by mopoe_daa.cu).
//
// Work unit of the GEMMs: a tile of 128 consecutive avatar rows (row = (validation, subject, score,
// sample) in the order of the output tensor, so a tile is one contiguous block of the avatar file).
// Each CTA owns a contiguous range of (validation, subject, score) series and walks it tile by tile:
//   T0  row metadata (score, centred score, mixture owner), cached experts of the <=2 subjects in the tile
//   T1  class heads of the perturbed src encoder:  D_h[128 x 2L] = relu(x W1^T + b1) Wh^T
//       hidden activations are produced 64 columns at a time by the CUDA cores straight into the
//       UMMA operand layout (fp16 hi/lo planes); 16 x 3 tcgen05.mma (kind::f16, M=128, N=48) per tile
//   T2  subset posterior / mixture owner / reparameterisation per row (thread = row, heads read
//       back from TMEM with tcgen05.ld), z written as the next A operand
//   T3  dst decoder: D[128 x 448] = z Wd^T, 3 x 3 x 2 tcgen05.mma (M=128, N=224)
//   T4  epilogue: tcgen05.ld 32x32b -> +bias -> smem transpose -> coalesced float4 stores of the
//       avatar tile AND the fp64 first-level regression sums (lane = ROI column)
// Accumulators live in TMEM (448 + 48 of the 512 columns); both weight matrices stay resident in
// shared memory as fp16 hi/lo planes for the whole launch (3xFP16 split, see mopoe_umma.cuh).
#pragma once

namespace mopoe {

constexpr int UM_ROWS = 128;
constexpr int UM_HCHUNK = 64;      // hidden columns produced per T1 step
constexpr int UM_STAGE_LD = 36;    // floats per staged row (16-byte aligned, conflict-free for .128 access)
constexpr int UM_MAXC = 16;        // max src width handled in registers
constexpr int UM_HEAD_LD = 49;     // staged heads row stride (odd: conflict-free thread-per-row access)

__device__ __noinline__ float4 philox_normal4_call(uint64_t seed, uint64_t stream, uint64_t blk) {
  float v[4];
  philox_normal4(seed, stream, blk, v);
  return make_float4(v[0], v[1], v[2], v[3]);
}

struct UmmaDims {
  int KC, KS, KZ;     // content / style / total K of the decoder GEMM (multiples of 8 / 8 / 16)
  int NH;             // N of the heads GEMM (2L rounded to 16)
  int ncol;           // valid decoder columns of this block (<= 448)
  int bias_slot;      // K slot carrying a constant 1 in z and the decoder bias in B (-1: no free pad slot)
};

struct UmmaSmem {
  int bd_hi, bd_lo, bh_hi, bh_lo, az_hi, az_lo, ah, a0, biasd, biash, score, xc, need, cache, spart, sacc, bars, total;
};

__host__ __device__ inline UmmaDims umma_dims(const ModelView& mv, int src, int dst, int ncol) {
  UmmaDims d;
  d.KC = (mv.L + 7) & ~7;
  d.KS = (mv.mod[dst].S + 7) & ~7;
  d.KZ = (d.KC + d.KS + 15) & ~15;
  d.NH = (2 * mv.L + 15) & ~15;
  d.ncol = ncol;
  d.bias_slot = d.KC > mv.L ? mv.L : (d.KS > mv.mod[dst].S ? d.KC + mv.mod[dst].S : (d.KZ > d.KC + d.KS ? d.KC + d.KS : -1));
  return d;
}

__host__ __device__ inline UmmaSmem umma_plan(const ModelView& mv, int src, const UmmaDims& d) {
  UmmaSmem p;
  int off = 0;
  auto take = [&](int bytes) { int o = off; off += (bytes + 127) & ~127; return o; };
  p.bd_hi = take(CB * d.KZ * 2); p.bd_lo = take(CB * d.KZ * 2);
  p.bh_hi = take(d.NH * MOPOE_HIDDEN * 2); p.bh_lo = take(d.NH * MOPOE_HIDDEN * 2);
  p.az_hi = take(UM_ROWS * d.KZ * 2); p.az_lo = take(UM_ROWS * d.KZ * 2);   // also: fp64 reduction scratch
  const int ah = 2 * UM_ROWS * UM_HCHUNK * 2, stage = 8 * 32 * UM_STAGE_LD * 4;
  p.ah = take(ah > stage ? ah : stage);                                        // A_h chunk | epilogue staging
  p.a0 = take(2 * 2 * MOPOE_HIDDEN * 4);   // per slot: hidden pre-activation without the perturbed column | that W1 column
  p.biasd = take(CB * 4);
  p.biash = take(d.NH * 4);
  p.score = take(UM_ROWS * 4);
  p.xc = take(UM_ROWS * 8);
  p.need = take(UM_ROWS * 4);
  p.cache = take(2 * 128 * 4);
  p.spart = take(4 * 2 * 64 * 8);   // [row group][slot][k] partial sums of xc * z
  p.sacc = take(2 * 64 * 8);        // [slot][k] running sums of the current / next series
  p.bars = take(64);
  p.total = off;
  return p;
}

// fp32 weights -> fp16 hi/lo planes in UMMA layout (global scratch), once per sweep
__global__ void daa_umma_prep_kernel(ModelView mv, int src, int dst, int col0, UmmaDims d, int cb /* decoder rows of the operand */, unsigned char* out) {
  using namespace umma;
  const ModView& ms = mv.mod[src];
  const ModView& md = mv.mod[dst];
  unsigned char* bd_hi = out;
  unsigned char* bd_lo = bd_hi + cb * d.KZ * 2;
  unsigned char* bh_hi = bd_lo + cb * d.KZ * 2;
  unsigned char* bh_lo = bh_hi + d.NH * MOPOE_HIDDEN * 2;
  const int nd = cb * (d.KZ / 8), nh = d.NH * (MOPOE_HIDDEN / 8);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nd + nh; i += gridDim.x * blockDim.x) {
    float x[8];
    if (i < nd) {
      const int n = i % cb, k8 = i / cb;       // decoder row (ROI) n, K permuted to [content | style | pad]
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int kz = k8 * 8 + q;
        float w = 0.f;
        if (n < d.ncol) {
          if (kz < d.KC) { if (kz < mv.L) w = md.wd[(int64_t)(col0 + n) * md.ZD + md.S + kz]; }
          else if (kz - d.KC < md.S) w = md.wd[(int64_t)(col0 + n) * md.ZD + (kz - d.KC)];
          if (kz == d.bias_slot) w = md.bd[col0 + n];
        }
        x[q] = w;
      }
      store_split8(bd_hi, bd_lo, core_off(n, k8, cb), x);
    } else {
      const int ii = i - nd, n = ii % d.NH, k8 = ii / d.NH;   // class-head output n (mu | logvar), K = hidden
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = n < 2 * mv.L ? ms.wh[(int64_t)n * MOPOE_HIDDEN + k8 * 8 + q] : 0.f;
      store_split8(bh_hi, bh_lo, core_off(n, k8, d.NH), x);
    }
  }
}

template <bool FIXED>
__global__ void __launch_bounds__(MOPOE_THREADS, 1) daa_avatar_umma_kernel(ModelView mv, DaaCtx cx, DaaWs ws, int col0) {
  using namespace umma;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int src = cx.q.src_mod, dst = cx.q.dst_mod;
  const ModView& ms = mv.mod[src];
  const ModView& mdst = mv.mod[dst];
  const int L = mv.L, M = mv.M, E = mv.E, Sd = mdst.S;
  const int C = cx.C, R = cx.R, J = cx.J, N = cx.N;
  const UmmaDims dm = umma_dims(mv, src, dst, min(CB, R - col0));
  const UmmaSmem pl = umma_plan(mv, src, dm);
  const int ncol = dm.ncol, KZ = dm.KZ, NH = dm.NH;
  unsigned char* s_bd_hi = smem + pl.bd_hi;
  unsigned char* s_bd_lo = smem + pl.bd_lo;
  unsigned char* s_bh_hi = smem + pl.bh_hi;
  unsigned char* s_bh_lo = smem + pl.bh_lo;
  unsigned char* s_az_hi = smem + pl.az_hi;
  unsigned char* s_az_lo = smem + pl.az_lo;
  unsigned char* s_ah_hi = smem + pl.ah;
  unsigned char* s_ah_lo = s_ah_hi + UM_ROWS * UM_HCHUNK * 2;
  float* s_stage = reinterpret_cast<float*>(smem + pl.ah) + warp * 32 * UM_STAGE_LD;
  float* s_heads = reinterpret_cast<float*>(smem + pl.ah);      // [128][UM_HEAD_LD], free between T1 and T4
  // fp64 reduction scratch [3][4][CB]: spans az_hi, az_lo and (FIXED) the head of the ah region, all idle then
  double* s_red = reinterpret_cast<double*>(smem + pl.az_hi);
  float* s_a0 = reinterpret_cast<float*>(smem + pl.a0);         // [2 slots][256] pre-activation w/o perturbed column
  float* s_w1c = s_a0 + 2 * MOPOE_HIDDEN;                       // [2 slots][256] perturbed column of W1
  float* s_biasd = reinterpret_cast<float*>(smem + pl.biasd);
  float* s_biash = reinterpret_cast<float*>(smem + pl.biash);
  float* s_score = reinterpret_cast<float*>(smem + pl.score);
  double* s_xc = reinterpret_cast<double*>(smem + pl.xc);
  int* s_need = reinterpret_cast<int*>(smem + pl.need);
  // per slot: [0..31] A, [32..63] B (posterior partials, see T0), [64..95] dst style mu, [96..127] dst style sd
  float* s_cache = reinterpret_cast<float*>(smem + pl.cache);
  double* s_spart = reinterpret_cast<double*>(smem + pl.spart);
  double* s_sacc = reinterpret_cast<double*>(smem + pl.sacc);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + pl.bars);   // [0] heads, [1] decoder
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + pl.bars + 32);
  constexpr int CSLOT = 128;

  // ---- launch-lifetime state: operands, biases, barriers, TMEM ----
  {
    const uint4* g = reinterpret_cast<const uint4*>(ws.bsplit);
    const int nbd = CB * KZ * 2 / 16, nbh = NH * MOPOE_HIDDEN * 2 / 16;
    for (int i = t; i < nbd; i += MOPOE_THREADS) {
      reinterpret_cast<uint4*>(s_bd_hi)[i] = g[i];
      reinterpret_cast<uint4*>(s_bd_lo)[i] = g[nbd + i];
    }
    for (int i = t; i < nbh; i += MOPOE_THREADS) {
      reinterpret_cast<uint4*>(s_bh_hi)[i] = g[2 * nbd + i];
      reinterpret_cast<uint4*>(s_bh_lo)[i] = g[2 * nbd + nbh + i];
    }
    for (int i = t; i < CB; i += MOPOE_THREADS) s_biasd[i] = i < ncol ? mdst.bd[col0 + i] : 0.f;
    for (int i = t; i < NH; i += MOPOE_THREADS) s_biash[i] = i < 2 * L ? ms.bh[i] : 0.f;
  }
  if (t < 128) s_sacc[t] = 0.0;
  if (t == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(s_tmem, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t tmem_heads = tmem + CB;                      // columns [448, 448 + NH)
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t idesc_h = idesc_f16(UM_ROWS, NH), idesc_d = idesc_f16(UM_ROWS, CB / 2);
  const uint32_t LBO_A = (UM_ROWS / 8) * 128, LBO_BH = (NH / 8) * 128, LBO_BD = (CB / 8) * 128;
  uint32_t ph_h = 0, ph_d = 0;
  bool timed_out = false;
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#define MOPOE_PHASE(i) do { if (t == 0) { const long long _n = clock64(); pc[i] += _n - tprev; tprev = _n; } } while (0)
  const bool fast_post = cx.q.sample_latents != 0;    // posterior of the row's mixture owner from cached partial sums

  // ---- this CTA's contiguous range of (validation, subject, score) series ----
  const int n_units = cx.q.n_val * N * C;
  const int per = (n_units + gridDim.x - 1) / gridDim.x;
  const int u0 = min(n_units, (int)blockIdx.x * per), u1 = min(n_units, u0 + per);
  const int64_t row_begin = (int64_t)u0 * J, row_end = (int64_t)u1 * J;
  int cur_unit = u0;
  constexpr int NCH = 7;   // column chunks of 32 per warp (2 warps per TMEM lane quarter)
  double accA[NCH], accB[NCH], syA[FIXED ? NCH : 1], syB[FIXED ? NCH : 1], syyA[FIXED ? NCH : 1], syyB[FIXED ? NCH : 1];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    accA[i] = accB[i] = 0.0;
    if (FIXED) { syA[i] = syB[i] = syyA[i] = syyB[i] = 0.0; }
  }
  // mixture owner of subject row g (launch-invariant table lookups)
  auto owner_subset = [&](int g, bool& need_src) -> int {
    int owner = 0, kidx = 0, s_own = 0;
    for (int k = 0; k < cx.b.n_mix; ++k)
      if (g >= cx.b.joint_bounds[k] && g < cx.b.joint_bounds[k + 1]) owner = k;
    for (int s = 0; s < mv.sub.n_subsets; ++s) {
      if (!in_mixture(mv, cx.b, s)) continue;
      if (kidx == owner) s_own = s;
      ++kidx;
    }
    need_src = !cx.q.sample_latents || ((mv.sub.mask[s_own] >> src) & 1) ||
               (mv.method == MOPOE_METHOD_MOE && mv.sub.n_members[s_own] > 1);
    return s_own;
  };

  for (int64_t tile_row = row_begin; tile_row < row_end; tile_row += UM_ROWS) {
    // ================= T0: row metadata, per-series caches =================
    const int r = t & 127;
    const int64_t rho = tile_row + r;
    const bool valid = rho < row_end;
    const int u = valid ? (int)(rho / J) : cur_unit;
    const int j = valid ? (int)(rho - (int64_t)u * J) : 0;
    const int c = u % C, g = (u / C) % N, v = u / (C * N);
    bool need = false;
    const int s_own = owner_subset(g, need);
    need = need && valid;
    const float score = valid ? ws.scores[(int64_t)u * J + j] : 0.f;
    if (t < UM_ROWS) {
      s_score[r] = score;
      s_xc[r] = (double)score - ws.xstat[(((int64_t)v * C + c) * N + g) * 2];
      s_need[r] = need;
    }
    // at most one series boundary per tile (J >= 128, enforced by the host dispatch)
    const int64_t bnd = (int64_t)(cur_unit + 1) * J - tile_row;
    const int rb = bnd < UM_ROWS ? (int)bnd : UM_ROWS;
    // slot 0 = current series, slot 1 = next series.  Thread k owns hidden unit k.
#pragma unroll 1
    for (int slot = 0; slot < 2; ++slot) {
      const int uu = min(cur_unit + slot, n_units - 1);
      const int uc = uu % C;
      const float* xs = cx.x[src] + ((int64_t)(uu / (C * N)) * N + (uu / C) % N) * C;
      const float* w = ms.w1 + (int64_t)t * C;
      float a = ms.b1[t];
      for (int i = 0; i < C; ++i) a = (i == uc) ? a : fmaf(w[i], xs[i], a);
      s_a0[slot * MOPOE_HIDDEN + t] = a;
      s_w1c[slot * MOPOE_HIDDEN + t] = w[uc];
    }
    // posterior partial sums of the mixture owner and dst style of both series
    for (int i = t; i < 2 * (L + Sd); i += MOPOE_THREADS) {
      const int slot = i / (L + Sd), k = i % (L + Sd);
      const int uu = min(cur_unit + slot, n_units - 1);
      const int ug = (uu / C) % N;
      const int64_t row = (int64_t)(uu / (C * N)) * N + ug;
      float* cs = s_cache + slot * CSLOT;
      if (k < L) {
        bool nd;
        const int so = owner_subset(ug, nd);
        float A = 0.f, B = 0.f;
        if (mv.method == MOPOE_METHOD_MOE) {
          // singleton owner: the member's own posterior (copy); multi-member subsets never own rows
          const int m = mv.sub.members[so][0];
          A = ws.enc[m][row * mv.mod[m].HC + k];
          B = expf(0.5f * ws.enc[m][row * mv.mod[m].HC + L + k]);
        } else {
          // sum of precisions / precision-weighted means over every expert of the owner but src
          const int nm = mv.sub.n_members[so];
          for (int q = 0; q < nm; ++q) {
            const int m = mv.sub.members[so][q];
            if (m == src) continue;
            const float T = 1.f / (expf(ws.enc[m][row * mv.mod[m].HC + L + k]) + MOPOE_POE_EPS);
            A += T; B += ws.enc[m][row * mv.mod[m].HC + k] * T;
          }
          if (mv.method == MOPOE_METHOD_POE || nm == M) A += 1.f / (1.f + MOPOE_POE_EPS);
          if (!((mv.sub.mask[so] >> src) & 1)) {   // owner without src: finished posterior (mu, sd)
            const float mu = B / A, lv = logf(1.f / A);
            A = mu; B = expf(0.5f * lv);
          }
        }
        cs[k] = A; cs[32 + k] = B;
      } else {
        const int s = k - L;
        cs[64 + s] = ws.enc[dst][row * mdst.HC + 2 * L + s];
        cs[96 + s] = expf(0.5f * ws.enc[dst][row * mdst.HC + 2 * L + Sd + s]);
      }
    }
    const int tile_need = __syncthreads_or(need ? 1 : 0);
    const int slot = (r < rb) ? 0 : 1;
    MOPOE_PHASE(0);
    // ================= T1: class heads of the perturbed src rows =================
    if (tile_need) {
      const int half = t >> 7;
      const float4* a0p = reinterpret_cast<const float4*>(s_a0 + slot * MOPOE_HIDDEN);
      const float4* wcp = reinterpret_cast<const float4*>(s_w1c + slot * MOPOE_HIDDEN);
#pragma unroll 1
      for (int kc = 0; kc < MOPOE_HIDDEN / UM_HCHUNK; ++kc) {
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) {
          const int k4 = (kc * UM_HCHUNK + half * 32 + i8 * 8) >> 2;
          const float4 a0 = a0p[k4], a1 = a0p[k4 + 1], w0 = wcp[k4], w1 = wcp[k4 + 1];
          float hv[8];
          hv[0] = fmaxf(fmaf(w0.x, score, a0.x), 0.f); hv[1] = fmaxf(fmaf(w0.y, score, a0.y), 0.f);
          hv[2] = fmaxf(fmaf(w0.z, score, a0.z), 0.f); hv[3] = fmaxf(fmaf(w0.w, score, a0.w), 0.f);
          hv[4] = fmaxf(fmaf(w1.x, score, a1.x), 0.f); hv[5] = fmaxf(fmaf(w1.y, score, a1.y), 0.f);
          hv[6] = fmaxf(fmaf(w1.z, score, a1.z), 0.f); hv[7] = fmaxf(fmaf(w1.w, score, a1.w), 0.f);
          store_split8(s_ah_hi, s_ah_lo, core_off(r, half * 4 + i8, UM_ROWS), hv);
        }
        fence_proxy_async();
        __syncthreads();
        if (t == 0) {
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < UM_HCHUNK / 16; ++ks) {
            const uint32_t oa = ks * 2 * LBO_A, ob = (kc * (UM_HCHUNK / 16) + ks) * 2 * LBO_BH;
            const uint64_t dah = smem_desc(smem_u32(s_ah_hi) + oa, LBO_A, 128), dal = smem_desc(smem_u32(s_ah_lo) + oa, LBO_A, 128);
            const uint64_t dbh = smem_desc(smem_u32(s_bh_hi) + ob, LBO_BH, 128), dbl = smem_desc(smem_u32(s_bh_lo) + ob, LBO_BH, 128);
            mma_f16(tmem_heads, dah, dbh, idesc_h, (kc | ks) ? 1u : 0u);
            mma_f16(tmem_heads, dah, dbl, idesc_h, 1u);
            mma_f16(tmem_heads, dal, dbh, idesc_h, 1u);
          }
          mma_commit(&s_bar[0]);
        }
        if (!mbar_wait(&s_bar[0], ph_h)) timed_out = true;
        ph_h ^= 1;
      }
      tc_fence_after();
      if (t < UM_ROWS) {   // heads (+bias) of this thread's row -> smem, so both thread halves can use them
        float hd[48];
        tmem_ld32(tmem_heads + lane_base, hd);
        if (NH > 32) tmem_ld16(tmem_heads + lane_base + 32, hd + 32);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 48; ++i)
          if (i < 2 * L) s_heads[r * UM_HEAD_LD + i] = hd[i] + s_biash[i];
      }
      __syncthreads();
    }
    MOPOE_PHASE(1);
    // ================= T2: posterior, reparameterisation, z -> A operand =================
    {
      const float* cs = s_cache + slot * CSLOT;
      const int64_t ridx = (((int64_t)(cx.v_av_off + v) * J + j) * C + c) * N + g;   // noise row (see fill_noise_row)
      const bool my_need = need;
      const int nq = KZ / 8, nqc = dm.KC / 8;
      const int nq0 = nqc < 2 ? nqc : 2;                 // thread half 0: first content chunks, half 1: the rest
      const int qbeg = (t < UM_ROWS) ? 0 : nq0, qend = (t < UM_ROWS) ? nq0 : nq;
      float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);
      int64_t blk_have = -1;
#pragma unroll 1
      for (int q = qbeg; q < qend; ++q) {
        float zb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kz = q * 8 + i;
          float z = 0.f;
          const bool content = kz < dm.KC;
          const int l = content ? kz : kz - dm.KC;
          if (valid && kz == dm.bias_slot) z = 1.0f;
          if (valid && l < (content ? L : Sd)) {
            float e0 = 0.f;
            if (cx.q.sample_latents) {
              if (cx.nz_av.eps) e0 = cx.nz_av.eps[ridx * E + (content ? 0 : mdst.eps_off) + l];
              else {
                const int64_t idx = ridx * mv.EP + (content ? 0 : mdst.peps_off) + l;
                if ((idx >> 2) != blk_have) { blk_have = idx >> 2; nv = philox_normal4_call(cx.nz_av.seed, cx.nz_av.stream, (uint64_t)blk_have); }
                const int w = (int)(idx & 3);
                e0 = w == 0 ? nv.x : w == 1 ? nv.y : w == 2 ? nv.z : nv.w;
              }
            }
            if (!content) {
              z = cs[64 + l] + cs[96 + l] * e0;
            } else if (fast_post) {
              float mu = cs[l], sd = cs[32 + l];
              if (my_need) {
                const float hm = s_heads[r * UM_HEAD_LD + l], hl = s_heads[r * UM_HEAD_LD + L + l];
                if (mv.method == MOPOE_METHOD_MOE) { mu = hm; sd = expf(0.5f * hl); }
                else {
                  const float T = 1.f / (expf(hl) + MOPOE_POE_EPS);
                  const float sT = mu + T;                 // cs[l] = sum of the other precisions
                  mu = (sd + hm * T) / sT;                 // cs[32+l] = sum of the other mu*T
                  sd = expf(0.5f * logf(1.f / sT));
                }
              }
              z = e0 * sd + mu;
            } else {
              // sample_latents = False: mean of the mixture components' means (BaseMMVae.py:228-229)
              float mu_e[MOPOE_MAX_MODS], lv_e[MOPOE_MAX_MODS];
              const int64_t row = (int64_t)v * N + g;
#pragma unroll
              for (int m = 0; m < MOPOE_MAX_MODS; ++m) {
                mu_e[m] = m < M ? ws.enc[m][row * mv.mod[m].HC + l] : 0.f;
                lv_e[m] = m < M ? ws.enc[m][row * mv.mod[m].HC + L + l] : 0.f;
              }
              mu_e[src] = s_heads[r * UM_HEAD_LD + l]; lv_e[src] = s_heads[r * UM_HEAD_LD + L + l];
              float jmu = 0.f;
#pragma unroll 1
              for (int s = 0; s < mv.sub.n_subsets; ++s)
                if (in_mixture(mv, cx.b, s)) jmu += eval_subset(mv, cx.b, s, g, mu_e, lv_e).mu;
              z = jmu / (float)cx.b.n_mix;
            }
          }
          zb[i] = z;
        }
        store_split8(s_az_hi, s_az_lo, core_off(r, q, UM_ROWS), zb);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    MOPOE_PHASE(2);
    // ================= T3: dst decoder GEMM =================
    if (t == 0) {
      tc_fence_after();
      for (int ks = 0; ks < KZ / 16; ++ks) {
        const uint32_t oa = ks * 2 * LBO_A;
        const uint64_t dah = smem_desc(smem_u32(s_az_hi) + oa, LBO_A, 128), dal = smem_desc(smem_u32(s_az_lo) + oa, LBO_A, 128);
#pragma unroll
        for (int hN = 0; hN < 2; ++hN) {
          const uint32_t ob = ks * 2 * LBO_BD + hN * (CB / 2 / 8) * 128;
          const uint64_t dbh = smem_desc(smem_u32(s_bd_hi) + ob, LBO_BD, 128), dbl = smem_desc(smem_u32(s_bd_lo) + ob, LBO_BD, 128);
          const uint32_t dcol = tmem + hN * (CB / 2);
          mma_f16(dcol, dah, dbh, idesc_d, ks ? 1u : 0u);
          mma_f16(dcol, dah, dbl, idesc_d, 1u);
          mma_f16(dcol, dal, dbh, idesc_d, 1u);
        }
      }
      mma_commit(&s_bar[1]);
    }
    if (!FIXED) {
      // First-level regression by linearity, while the tensor cores run:  sum_rows xc * y[:, col] =
      // sum_k Wd[col][k] * (sum_rows xc * z[:, k])  (+ bias * sum xc through the constant-1 slot), so only
      // KZ running sums per series are needed instead of one fp64 FMA per avatar element.
      // z is read back exactly as the tensor cores see it (fp16 hi + lo planes).
      if (t < 4 * KZ) {
        const int k = t % KZ, rg = t / KZ;
        const unsigned char* ph = s_az_hi + (k >> 3) * LBO_A + (k & 7) * 2;
        const unsigned char* plo = s_az_lo + (k >> 3) * LBO_A + (k & 7) * 2;
        double a = 0.0, b = 0.0;
        const int rlo = rg * 32, rhi = min(rg * 32 + 32, (int)min((int64_t)UM_ROWS, row_end - tile_row));
        for (int rr = rlo; rr < rhi; ++rr) {
          const uint32_t o = (rr >> 3) * 128u + (rr & 7) * 16u;
          const float zv = __half2float(*reinterpret_cast<const __half*>(ph + o)) + __half2float(*reinterpret_cast<const __half*>(plo + o));
          const double p = s_xc[rr] * (double)zv;
          if (rr < rb) a += p; else b += p;
        }
        s_spart[(rg * 2 + 0) * 64 + k] = a;
        s_spart[(rg * 2 + 1) * 64 + k] = b;
      }
    }
    if (!mbar_wait(&s_bar[1], ph_d)) timed_out = true;
    ph_d ^= 1;
    tc_fence_after();
    MOPOE_PHASE(3);
    // ================= T4: epilogue =================
    {
      const int q4 = warp & 3, hh = warp >> 2;
      const int rows_left = (int)max((int64_t)0, min((int64_t)32, row_end - (tile_row + q4 * 32)));   // valid rows of this quarter
      const int rowsA = max(0, min(rows_left, rb - q4 * 32));   // rows of the current series; the rest belong to the next
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int cb0 = (hh * NCH + ci) * 32;
        float vv[32];
        tmem_ld32(tmem + lane_base + cb0, vv);
        tmem_ld_wait();
        if (dm.bias_slot < 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) vv[i] += s_biasd[cb0 + i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(s_stage + lane * UM_STAGE_LD + 4 * i) = make_float4(vv[4 * i], vv[4 * i + 1], vv[4 * i + 2], vv[4 * i + 3]);
        __syncwarp();
        if (cx.avatars && cb0 < ncol) {
          const int rr0 = lane >> 3, cg = (lane & 7) * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rr0;
            if (rr < rows_left) {
              const float4 o = *reinterpret_cast<const float4*>(s_stage + rr * UM_STAGE_LD + cg);
              float* dstp = cx.avatars + (tile_row + q4 * 32 + rr) * (int64_t)R + col0 + cb0 + cg;
              if ((R & 3) == 0 && cb0 + cg + 4 <= ncol) *reinterpret_cast<float4*>(dstp) = o;
              else {
                if (cb0 + cg + 0 < ncol) dstp[0] = o.x;
                if (cb0 + cg + 1 < ncol) dstp[1] = o.y;
                if (cb0 + cg + 2 < ncol) dstp[2] = o.z;
                if (cb0 + cg + 3 < ncol) dstp[3] = o.w;
              }
            }
          }
        }
        // pooled ("fixed") regression needs sum y and sum y^2 per element: lane = ROI column cb0 + lane
        if (FIXED) {
          const float* sp = s_stage + lane;
          const double* xp = s_xc + q4 * 32;
          double a = accA[ci], sy = FIXED ? syA[ci] : 0.0, syy = FIXED ? syyA[ci] : 0.0;
#pragma unroll 4
          for (int rr = 0; rr < rowsA; ++rr) {
            const double y = (double)sp[rr * UM_STAGE_LD];
            a = fma(xp[rr], y, a);
            if (FIXED) { sy += y; syy = fma(y, y, syy); }
          }
          accA[ci] = a;
          if (FIXED) { syA[ci] = sy; syyA[ci] = syy; }
          if (rowsA < rows_left) {
            double b = accB[ci], ty = FIXED ? syB[ci] : 0.0, tyy = FIXED ? syyB[ci] : 0.0;
#pragma unroll 4
            for (int rr = rowsA; rr < rows_left; ++rr) {
              const double y = (double)sp[rr * UM_STAGE_LD];
              b = fma(xp[rr], y, b);
              if (FIXED) { ty += y; tyy = fma(y, y, tyy); }
            }
            accB[ci] = b;
            if (FIXED) { syB[ci] = ty; syyB[ci] = tyy; }
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    __syncthreads();
    MOPOE_PHASE(4);
    // ================= series finished inside this tile: reduce the 4 row quarters, emit slopes =================
    if (!FIXED && t < 2 * KZ) {   // fold this tile's partial sums into the running sums of both series
      const int sl = t / KZ, k = t % KZ;
      s_sacc[sl * 64 + k] += s_spart[(0 * 2 + sl) * 64 + k] + s_spart[(1 * 2 + sl) * 64 + k] +
                             s_spart[(2 * 2 + sl) * 64 + k] + s_spart[(3 * 2 + sl) * 64 + k];
    }
    if (!FIXED) __syncthreads();
    if (!FIXED && (int64_t)(cur_unit + 1) * J <= min(tile_row + UM_ROWS, row_end)) {
      const int uc = cur_unit % C, ug = (cur_unit / C) % N, uv = cur_unit / (C * N);
      const int64_t obase = (((int64_t)uv * C + uc) * N + ug);
      const double sxx = ws.xstat[obase * 2 + 1];
      for (int col = t; col < ncol; col += MOPOE_THREADS) {
        const float* wrow = mdst.wd + (int64_t)(col0 + col) * mdst.ZD;
        double sxy = 0.0;
        for (int k = 0; k < L; ++k) sxy = fma(s_sacc[k], (double)wrow[Sd + k], sxy);
        for (int k = 0; k < Sd; ++k) sxy = fma(s_sacc[dm.KC + k], (double)wrow[k], sxy);
        if (dm.bias_slot >= 0) sxy = fma(s_sacc[dm.bias_slot], (double)mdst.bd[col0 + col], sxy);
        ws.betas[obase * R + col0 + col] = sxy / sxx;
      }
      __syncthreads();
      if (t < KZ) { s_sacc[t] = s_sacc[64 + t]; s_sacc[64 + t] = 0.0; }
      ++cur_unit;
      __syncthreads();
    }
    if (FIXED && (int64_t)(cur_unit + 1) * J <= min(tile_row + UM_ROWS, row_end)) {
      const int q4 = warp & 3, hh = warp >> 2;
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int col = (hh * NCH + ci) * 32 + lane;
        s_red[(0 * 4 + q4) * CB + col] = accA[ci];
        if (FIXED) { s_red[(1 * 4 + q4) * CB + col] = syA[ci]; s_red[(2 * 4 + q4) * CB + col] = syyA[ci]; }
      }
      __syncthreads();
      const int uc = cur_unit % C, ug = (cur_unit / C) % N, uv = cur_unit / (C * N);
      const int64_t obase = (((int64_t)uv * C + uc) * N + ug);
      const double sxx = ws.xstat[obase * 2 + 1];
      for (int col = t; col < ncol; col += MOPOE_THREADS) {
        const double sxy = s_red[0 * CB + col] + s_red[1 * CB + col] + s_red[2 * CB + col] + s_red[3 * CB + col];
        const int64_t o = obase * R + col0 + col;
        ws.betas[o] = sxy / sxx;
        if (FIXED) {
          const double sy = s_red[4 * CB + col] + s_red[5 * CB + col] + s_red[6 * CB + col] + s_red[7 * CB + col];
          const double syy = s_red[8 * CB + col] + s_red[9 * CB + col] + s_red[10 * CB + col] + s_red[11 * CB + col];
          const double yb = sy / (double)J;
          ws.ybar[o] = yb;
          ws.syy[o] = syy - (double)J * yb * yb;
        }
      }
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        accA[ci] = accB[ci]; accB[ci] = 0.0;
        if (FIXED) { syA[ci] = syB[ci]; syB[ci] = 0.0; syyA[ci] = syyB[ci]; syyB[ci] = 0.0; }
      }
      ++cur_unit;
      __syncthreads();
    }

    MOPOE_PHASE(5);
  }
  if (t == 0 && ws.phase) for (int i = 0; i < 8; ++i) ws.phase[blockIdx.x * 8 + i] = pc[i];
  if (timed_out && t == 0) atomicExch(ws.err, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace mopoe
