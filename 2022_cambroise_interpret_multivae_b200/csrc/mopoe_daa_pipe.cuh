// Pipelined tensor-core DAA avatar kernel (included by mopoe_daa.cu): the production kernel for
// reg_method = hierarchical, sample_latents = True (workflow.py:406-419 + stat_utils.py:66-68).
//
// One persistent CTA per SM walks a contiguous range of 128-row tiles of the avatar tensor (row =
// (validation, subject, score, sample) in output order, so a tile is one contiguous 128 x R block of
// rois_digital_avatars.npy).  The tile pipeline is warp specialised and runs through mbarriers only:
//
//   producers (12 warps, thread = avatar row, three warps per TMEM lane quarter which split the
//   hidden units of P1 and the 8-latent K chunks of P2/P3 round-robin)
//       P1  hidden layer of the perturbed src encoder, rank-1 in the score:
//           h = relu(a0 + W1[:,c] * score), split into fp16 hi/lo and written STRAIGHT INTO TMEM
//           (tcgen05.st) as the A operand of the class-head GEMM            -> bar h_full
//       P2  Philox4x32-10 + Box-Muller noise of the row (whole blocks, see fill_noise_row)
//       P3  class heads read back from TMEM (tcgen05.ld), posterior of the row's mixture owner from
//           cached per-series partial sums, reparameterisation, z written as the A operand of the
//           decoder GEMM (shared memory, fp16 hi/lo, double buffered)          -> bar z_full
//           + the first-level regression sums  sum_rows (x - xbar) * z[k]  (fp32 warp butterfly over 32 rows,
//           fp64 across warps and tiles):
//           by linearity  sum_rows (x - xbar) * y[:, roi] = Wd[roi,:] . that vector, so the slope
//           of every ROI (stat_utils.py:66-68) needs KZ numbers per series, not one FMA per element
//   heads issuer   (1 thread)  16 x 3 tcgen05.mma kind::f16, A from TMEM, N = 48  -> bar heads_done
//   decoder issuer (1 thread)  per 96-column chunk 3 x 3 tcgen05.mma, A/B from shared memory,
//                              accumulators ping-pong between two TMEM buffers      -> bar acc_full
//   epilogue (4 warps)         tcgen05.ld -> shared-memory transpose -> coalesced streaming float4
//                              stores of the avatar tile                             -> bar acc_empty
//   aux (1 warp)               per-series caches one tile ahead (hidden pre-activation without the
//                              perturbed column, posterior partial sums of the other experts, dst
//                              style), folds the regression partials, flushes finished series
//
// TMEM (512 columns): [0,48) class-head accumulator | [64,192) h hi | [192,320) h lo |
//                     [320,416) decoder accumulator 0 | [416,512) decoder accumulator 1
// Shared memory: both weight matrices as fp16 hi/lo UMMA operands for the whole launch (3xFP16
// split: a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulation in TMEM).
#pragma once

namespace mopoe {

constexpr int PK_ROWS = 128;
#ifndef PK_NCH_
#define PK_NCH_ 96
#define PK_MAXCH_ 5
#define PK_ACC_BUFS_ 2
#endif
constexpr int PK_NCH = PK_NCH_;        // decoder columns per accumulator buffer (a multiple of 32, <= 256)
constexpr int PK_MAXCH = PK_MAXCH_;    // chunks per launch
constexpr int PK_ACC_BUFS = PK_ACC_BUFS_;   // accumulator buffers in TMEM: 2 = ping-pong between the decoder MMAs and the epilogue
#ifndef PK_Z_TMEM
#define PK_Z_TMEM 0                    // 1: the hi plane of z (A operand of the decoder GEMM) lives in TMEM, double buffered
#endif                                 //    (needs 48 free columns behind the accumulators: 64-column chunks)
constexpr int PK_TM_AZ = 320 + PK_ACC_BUFS * PK_NCH;   // [2][KZ / 2] packed fp16 pairs of z_hi (PK_Z_TMEM)
static_assert(PK_NCH % 32 == 0 && PK_NCH <= 256 && 320 + PK_ACC_BUFS * PK_NCH + (PK_Z_TMEM ? 64 : 0) <= 512,
              "accumulators (and the z operand) do not fit the TMEM plan");
constexpr int PK_CBP = PK_NCH * PK_MAXCH;   // 480 decoder columns per launch
#ifndef PK_NPW_
#define PK_NPW_ 3
#endif
constexpr int PK_NPW = PK_NPW_;        // producer warps per TMEM lane quarter
constexpr int PK_PROD = 4 * PK_NPW, PK_EPI = 4;
constexpr int PK_W_EPI = PK_PROD, PK_W_HMMA = PK_PROD + PK_EPI, PK_W_DMMA = PK_W_HMMA + 1, PK_W_AUX = PK_W_HMMA + 2;
constexpr int PK_THREADS = (PK_PROD + PK_EPI + 3) * 32;   // 608
constexpr int PK_STAGE_LD = 36;
constexpr int PK_STAGE_BYTES = 5 * 1024;   // per epilogue warp: 32 x 36 floats (padded rows) or one 4 KB swizzled TMA box
constexpr int PK_SLOTS = 4;            // per-series caches: [tile parity][first | second series of the tile]
constexpr int PK_TM_HEADS = 0, PK_TM_AH_HI = 64, PK_TM_AH_LO = 192, PK_TM_ACC = 320;
constexpr int PK_MAXSUBJ = 1024;       // subjects per validation batch (owner table in shared memory)
constexpr int PK_CACHE_F = 2 * MOPOE_HIDDEN + 128;   // floats per series cache: a0 | w1c | cs
constexpr int PK_REC_F = PK_CACHE_F + 4;             // floats per series record in HBM: cache | xbar (fp64) | need_src | pad
static_assert(PK_REC_F == DAA_SERIES_REC_F, "workspace carve and record layout disagree");

struct PipeSmem {
  int bd_hi, bd_lo, bh_hi, bh_lo, az, stage, cache, meta, xbar, biash, part, gmeta, tinfo, score, bars, total;
};

__host__ __device__ inline PipeSmem pipe_plan(const UmmaDims& d) {
  PipeSmem p;
  int off = 0;
  auto take = [&](int bytes) { int o = off; off += (bytes + 127) & ~127; return o; };
  p.bd_hi = take(PK_CBP * d.KZ * 2); p.bd_lo = take(PK_CBP * d.KZ * 2);
  p.bh_hi = take(d.NH * MOPOE_HIDDEN * 2); p.bh_lo = take(d.NH * MOPOE_HIDDEN * 2);
  p.az = take(2 * 2 * PK_ROWS * d.KZ * 2);             // [buffer][hi|lo]
  off = (off + 1023) & ~1023;                           // TMA 128-byte swizzle: the pattern repeats every 1024 bytes
  p.stage = take(PK_EPI * PK_STAGE_BYTES);
  p.cache = take(PK_SLOTS * PK_CACHE_F * 4);
  p.meta = take(PK_SLOTS * 4 * 4);
  p.xbar = take(PK_SLOTS * 2 * 8);                       // per series: xbar | noise row index of sample 0 (int64)
  p.tinfo = take(2 * 4 * 4);                             // [tile parity]: first, last series | first, end row of the tile
  p.biash = take(d.NH * 4);
  p.part = take(2 * 4 * 2 * 64 * 8);                    // [tile parity][lane quarter][series 0|1][k] fp64
  p.gmeta = take(PK_MAXSUBJ);                           // per subject row: bit 7 need_src | owner subset
  p.score = take(2 * PK_ROWS * 4);                      // [tile parity][row] perturbed scores of the tile
  p.bars = take(256);                                   // mbarriers | tmem base | abort | tile ring | published count
  p.total = off;
  return p;
}

// upper bound of the 128-row tiles one J-row series can touch
__host__ __device__ inline int pipe_tiles_per_unit(int J) { return (J + PK_ROWS - 2) / PK_ROWS + 1; }

// bounded mbarrier wait with a CTA-wide sticky abort flag: a protocol bug ends the launch with the
// error flag set (results are poisoned by the host wrapper) instead of hanging the GPU
__device__ __forceinline__ bool pk_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
  const uint32_t addr = umma::smem_u32(bar);
  // waiting roles must not steal issue slots from the working ones: poll with an exponential
  // __nanosleep backoff (32 .. 256 ns) instead of spinning on try_wait
  uint32_t ns = 32;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 21); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return true;
    __nanosleep(ns);
    ns = ns < 256 ? ns * 2 : 256;
    if ((spin & 63u) == 63u && *abort_flag) return false;
  }
  *abort_flag = 1;
  return false;
}
__device__ __forceinline__ void pk_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pk_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// 8 per-lane values xc * zq[k] -> their sum over the 32 lanes, returned in every lane with (lane & 7) == k
// (transposing butterfly inside each group of 8 lanes, then two exchanges between the groups).
// fp32: FP64 arithmetic on the CUDA cores of this part is slow enough that the fp64 version of this
// reduction was ~45 % of the producers' per-tile time; the 32-row partial sums are exact to ~3e-7 of their
// absolute sum, everything downstream (accumulation over warps and tiles, slopes, t-test) stays fp64.
__device__ __forceinline__ float pk_lane_transpose_sum8(const float* zq, float xc, int lane) {
  float a[4];
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lo = xc * zq[i], hi = xc * zq[i + 4];
      const float keep = up ? hi : lo, send = up ? lo : hi;
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
#pragma unroll
  for (int w = 2; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float keep = up ? a[i + w] : a[i], send = up ? a[i] : a[i + w];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  float r = a[0];
  r += __shfl_xor_sync(0xffffffffu, r, 8);
  r += __shfl_xor_sync(0xffffffffu, r, 16);
  return r;
}

__device__ __forceinline__ float pk_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float pk_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// mixture owner of subject row g and whether its posterior needs the perturbed src expert
__device__ __forceinline__ int pk_owner_subset(const ModelView& mv, const DaaCtx& cx, int g, bool& need_src) {
  int owner = 0, kidx = 0, s_own = 0;
  for (int k = 0; k < cx.b.n_mix; ++k)
    if (g >= cx.b.joint_bounds[k] && g < cx.b.joint_bounds[k + 1]) owner = k;
  if (prior_component(mv, cx.b, owner)) { need_src = false; return 0x7f; }   // jsd: z from the prior, cached posterior (0, 1)
  for (int s = 0; s < mv.sub.n_subsets; ++s) {
    if (!in_mixture(mv, cx.b, s)) continue;
    if (kidx == owner) s_own = s;
    ++kidx;
  }
  need_src = ((mv.sub.mask[s_own] >> cx.q.src_mod) & 1) || (moe_like(mv) && mv.sub.n_members[s_own] > 1);
  return s_own;
}

#ifdef PK_PROF
#define PK_T(i) do { if (lane == 0) { const long long _n = clock64(); pc[i] += _n - tprev; tprev = _n; } } while (0)
#else
#define PK_T(i) do { } while (0)
#endif

__global__ void __launch_bounds__(PK_THREADS, 1) daa_avatar_pipe_kernel(ModelView mv, DaaCtx cx, DaaWs ws, int col0, int tma_ok,
                                                                         const __grid_constant__ CUtensorMap tmap) {
  using namespace umma;
  extern __shared__ __align__(1024) unsigned char smem[];
#ifdef PK_PROF
  const long long t_entry = clock64();
  unsigned long long g_entry;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_entry));
#endif
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int src = cx.q.src_mod, dst = cx.q.dst_mod;
  const ModView& ms = mv.mod[src];
  const ModView& mdst = mv.mod[dst];
  const int L = mv.L, E = mv.E, Sd = mdst.S;
  const int C = cx.C, R = cx.R, J = cx.J, N = cx.N;
  const UmmaDims dm = umma_dims(mv, src, dst, min(PK_CBP, R - col0));
  const PipeSmem pl = pipe_plan(dm);
  const int ncol = dm.ncol, KZ = dm.KZ, KC = dm.KC, NH = dm.NH;
  const int n_chunks = (ncol + PK_NCH - 1) / PK_NCH;
  unsigned char* s_bd_hi = smem + pl.bd_hi;
  unsigned char* s_bd_lo = smem + pl.bd_lo;
  unsigned char* s_bh_hi = smem + pl.bh_hi;
  unsigned char* s_bh_lo = smem + pl.bh_lo;
  unsigned char* s_az = smem + pl.az;
  const int AZ_PLANE = PK_ROWS * KZ * 2;          // bytes of one fp16 plane of one z buffer
  float* s_cache = reinterpret_cast<float*>(smem + pl.cache);
  int* s_meta = reinterpret_cast<int*>(smem + pl.meta);         // [slot][0] need_src
  double* s_xbar = reinterpret_cast<double*>(smem + pl.xbar);
  const int64_t* s_rbase = reinterpret_cast<const int64_t*>(smem + pl.xbar) + PK_SLOTS;
  int* s_tinfo = reinterpret_cast<int*>(smem + pl.tinfo);
  float* s_biash = reinterpret_cast<float*>(smem + pl.biash);
  double* s_part = reinterpret_cast<double*>(smem + pl.part);
  unsigned char* s_gmeta = smem + pl.gmeta;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + pl.bars);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + pl.bars + 96);
  volatile int* s_abort = reinterpret_cast<volatile int*>(smem + pl.bars + 100);
  volatile int* s_ring = reinterpret_cast<volatile int*>(smem + pl.bars + 128);   // [8] global tile id of iteration i (or -1: no more tiles)
  volatile int* s_npub = reinterpret_cast<volatile int*>(smem + pl.bars + 160);   // iterations published so far
  float* s_score = reinterpret_cast<float*>(smem + pl.score);
  uint64_t* bar_h_full = s_bar + 0;       // producers -> heads issuer           (8 arrivals)
  uint64_t* bar_heads_done = s_bar + 1;   // heads MMAs complete                 (tcgen05.commit)
  uint64_t* bar_z_full = s_bar + 2;       // [2] producers -> decoder issuer     (8 arrivals)
  uint64_t* bar_z_free = s_bar + 4;       // [2] decoder MMAs of the tile done   (tcgen05.commit)
  uint64_t* bar_acc_full = s_bar + 6;     // [2] chunk accumulated               (tcgen05.commit)
  uint64_t* bar_acc_empty = s_bar + 8;    // [2] epilogue drained the buffer     (4 arrivals)
  uint64_t* bar_cache = s_bar + 10;       // [2] aux -> producers: series caches + tile info of the tile (1 arrival)

  // ---- launch-lifetime state ----
  {
    const uint4* g = reinterpret_cast<const uint4*>(ws.bsplit);
    const int nbd = PK_CBP * KZ * 2 / 16, nbh = NH * MOPOE_HIDDEN * 2 / 16;
    for (int i = t; i < nbd; i += PK_THREADS) {
      reinterpret_cast<uint4*>(s_bd_hi)[i] = g[i];
      reinterpret_cast<uint4*>(s_bd_lo)[i] = g[nbd + i];
    }
    for (int i = t; i < nbh; i += PK_THREADS) {
      reinterpret_cast<uint4*>(s_bh_hi)[i] = g[2 * nbd + i];
      reinterpret_cast<uint4*>(s_bh_lo)[i] = g[2 * nbd + nbh + i];
    }
    for (int i = t; i < NH; i += PK_THREADS) s_biash[i] = i < 2 * L ? ms.bh[i] : 0.f;
    for (int g = t; g < N; g += PK_THREADS) {
      bool nd;
      const int so = pk_owner_subset(mv, cx, g, nd);
      s_gmeta[g] = (unsigned char)(so | (nd ? 0x80 : 0));
    }
  }
  if (t == 0) {
    mbar_init(bar_h_full, PK_PROD); mbar_init(bar_heads_done, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_z_full + b, PK_PROD); mbar_init(bar_z_free + b, 1);
      mbar_init(bar_acc_full + b, 1); mbar_init(bar_acc_empty + b, PK_EPI);
      mbar_init(bar_cache + b, 1);
    }
    *s_abort = 0;
    *s_npub = 0;
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(s_tmem, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  // ---- tiles are handed out DYNAMICALLY: iteration 0 of CTA b is tile b, every further tile comes from a global
  // counter (the aux warp draws it one tile ahead and publishes it through a small ring in shared memory).
  // CTAs differ by up to 25 % in speed (the store path of some SMs is slower), a static split waits for the
  // slowest.  The tile grid restarts at every validation (the last tile of a validation is partial), and every
  // tile writes its regression sums to its own (series, tile) slot, so neither the grouping of the fp64 sums
  // nor the tables depend on which CTA ran which tile or on how validations are sharded.
  const int rpv = N * C * J;                          // rows per validation; n_val * rpv < 2^31 (host check)
  const int tpv = (rpv + PK_ROWS - 1) / PK_ROWS;      // tiles per validation
  const int total_tiles = cx.q.n_val * tpv;
  const int tpu = pipe_tiles_per_unit(J);
  auto unit_need = [&](int u) -> bool { return (s_gmeta[(u / C) % N] & 0x80) != 0; };
  auto tile_rows = [&](int gt, int& r0, int& r1) {    // rows [r0, r1) of global tile gt
    const int v = gt / tpv;
    r0 = v * rpv + (gt - v * tpv) * PK_ROWS;
    r1 = min(r0 + PK_ROWS, (v + 1) * rpv);
  };
  auto tile_units = [&](int gt, int& uA, int& uB) {
    int r0, r1;
    tile_rows(gt, r0, r1);
    uA = r0 / J; uB = (r1 - 1) / J;
  };
  // global tile id of this CTA's iteration i, once the aux warp has published it (-1: no more tiles)
  auto tile_of = [&](int i) -> int {
    uint32_t ns = 32;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 21) && *s_npub <= i; ++spin) {
      __nanosleep(ns);
      ns = ns < 256 ? ns * 2 : 256;
      if ((spin & 63u) == 63u && *s_abort) return -1;
    }
    if (*s_npub <= i) { *s_abort = 1; return -1; }
    __threadfence_block();
    return s_ring[i & 7];
  };

#ifdef PK_PROF
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
  const long long tstart = tprev;
#endif
  if (warp < PK_PROD) {
    // =============================== producers ===============================
    const int q4 = warp & 3, wq = warp >> 2;             // TMEM lane quarter, rank among its producer warps
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
    const int nbrow = mv.EP >> 2;
    const int nqc = KC >> 3, nqt = KZ >> 3;              // content chunks, all chunks of 8 latents
    uint32_t heads_waits = 0;
#pragma unroll 1
    for (int i = 0;; ++i) {
      if (!pk_wait(bar_cache + (i & 1), (i >> 1) & 1, s_abort)) break;   // tile info, series caches, scores are ready
      const int rot = (wq + i) % PK_NPW;                 // the warp with one more P1 step rotates over the tiles
      PK_T(0);
      const int* ti = s_tinfo + (i & 1) * 4;                  // series and rows of the tile (aux warp)
      const int uA = ti[0], uB = ti[1], tile_row = ti[2], tile_end = ti[3];
      if (tile_row < 0) break;                           // no more tiles
      const int rho = tile_row + r;
      const bool valid = rho < tile_end;
      const int u = (rho < (uA + 1) * J) ? uA : uB;
      const int j = valid ? rho - u * J : 0;
      const int slot = (i & 1) * 2 + (u - uA);           // the tile's (<= 2) series sit in the slots of its parity
      const float* cache = s_cache + slot * PK_CACHE_F;
      const float* cs = cache + 2 * MOPOE_HIDDEN;
      const bool tile_need = s_meta[(i & 1) * 8] || (uB != uA && s_meta[(i & 1) * 8 + 4]);
      const bool need = valid && s_meta[slot * 4];
      const float score = s_score[(i & 1) * PK_ROWS + r];
      // ---- P1: hidden layer -> TMEM (A operand of the class-head GEMM), 16 hidden units per step ----
      if (tile_need) {
#pragma unroll 2
        for (int kc = rot; kc < MOPOE_HIDDEN / 16; kc += PK_NPW) {
          const float4* a0p = reinterpret_cast<const float4*>(cache + kc * 16);
          const float4* wcp = reinterpret_cast<const float4*>(cache + MOPOE_HIDDEN + kc * 16);
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 a = a0p[q], w = wcp[q];
            relu_split_pack2(fmaf(w.x, score, a.x), fmaf(w.y, score, a.y), hi[2 * q], lo[2 * q]);
            relu_split_pack2(fmaf(w.z, score, a.z), fmaf(w.w, score, a.w), hi[2 * q + 1], lo[2 * q + 1]);
          }
          tmem_st8(lane_addr + PK_TM_AH_HI + kc * 8, hi);    // 32-bit column = 2 hidden units
          tmem_st8(lane_addr + PK_TM_AH_LO + kc * 8, lo);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) pk_arrive(bar_h_full);
      }
      PK_T(1);
      pk_wait(bar_z_free + (i & 1), ((i >> 1) & 1) ^ 1, s_abort);
      PK_T(5);
      // ---- P2 + P3, one K chunk of 8 latents per pass (content chunks, then dst style chunks): noise,
      // class heads from TMEM, posterior of the row's mixture owner, reparameterisation, z -> A operand of
      // the decoder GEMM, first-level regression sums ----
      {
        const int64_t ridx = s_rbase[slot] + (int64_t)j * (C * N);   // (((v_off + v) J + j) C + c) N + g
        unsigned char* az_hi = s_az + (i & 1) * 2 * AZ_PLANE;
        unsigned char* az_lo = az_hi + AZ_PLANE;
        const float xc = valid ? (float)((double)score - s_xbar[slot]) : 0.f;
        const bool inA = valid && u == uA, inB = valid && u != uA;
        const bool anyA = __any_sync(0xffffffffu, inA), anyB = __any_sync(0xffffffffu, inB);
        double* part = s_part + (((i & 1) * 4 + q4) * 2) * 64;
        bool waited = false;
        // chunks in DESCENDING order: the style chunks need no class heads, so their whole pass (and the noise
        // of the content chunk) runs while the class-head MMAs of this tile are still in flight -- in
        // ascending order every producer warp sat ~2.4 K of its ~15 K cycles per tile in the heads_done wait
        const int ci_last = rot < nqt ? rot + ((nqt - 1 - rot) / PK_NPW) * PK_NPW : -1;
#pragma unroll 1
        for (int ci = ci_last; ci >= 0; ci -= PK_NPW) {
          const bool content = ci < nqc;
          const int l0 = (content ? ci : ci - nqc) * 8;        // first latent of the chunk inside its section
          const int cnt = content ? L : Sd;
          const float* csb = cs + (content ? 0 : 64) + l0;     // [0..32) mean part, [32..64) scale part
          float e[8];
          if (cx.nz_av.eps) {
            const float* ep = cx.nz_av.eps + ridx * E + (content ? 0 : mdst.eps_off) + l0;
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = (valid && l0 + k < cnt) ? ep[k] : 0.f;
          } else {
            // two independent Philox chains (draws past the end of the section are unused)
            const uint64_t blk = (uint64_t)(ridx * nbrow + ((content ? 0 : mdst.peps_off) + l0) / 4);
#ifdef PK_EXP_NONOISE    // experiment (WRONG results): what the generator costs on the kernel's critical path
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = __uint_as_float(0x3e000000u | (((uint32_t)blk + k) & 0xffffu));
#else
            philox_normal4(cx.nz_av, blk, e);
            philox_normal4(cx.nz_av, blk + 1, e + 4);
#endif
          }
          const bool heads = content && tile_need;
          float hm[8], hl[8];
          if (heads) {
            if (!waited) {
              PK_T(2);
              pk_wait(bar_heads_done, heads_waits & 1, s_abort);
              ++heads_waits;
              waited = true;
              tc_fence_after();
              PK_T(3);
            }
            tmem_ld8(lane_addr + PK_TM_HEADS + l0, hm);
            tmem_ld8(lane_addr + PK_TM_HEADS + L + l0, hl);
            tmem_ld_wait();
          }
          const bool upd = heads && need;
          const int kb = dm.bias_slot - ci * 8;          // decoder bias rides in a free pad slot of K
          float zq[8];      // ends up holding z exactly as the tensor cores see it (hi + lo)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float mu = csb[k], sd = csb[32 + k];
            if (heads) {        // warp-uniform; rows that do not need the src expert keep the cached posterior
              const float m_ = hm[k] + s_biash[l0 + k], lv = hl[k] + s_biash[L + l0 + k];
              float mu2, sd2;
              if (moe_like(mv)) { mu2 = m_; sd2 = pk_ex2(0.5f * 1.4426950408889634f * lv); }
              else {
                // poe (mm_div.py:13-20) in MUFU arithmetic: T = 1/(exp(lv)+eps), var = 1/sum T
                const float T = pk_rcp(pk_ex2(1.4426950408889634f * lv) + MOPOE_POE_EPS);
                const float sT = mu + T;                 // cs[l] = sum of the other precisions
                mu2 = (sd + m_ * T) * pk_rcp(sT);        // cs[32+l] = sum of the other mu*T
                sd2 = rsqrtf(sT);                        // exp(0.5 * log(1 / sT))
              }
              mu = upd ? mu2 : mu; sd = upd ? sd2 : sd;
            }
            float z = fmaf(e[k], sd, mu);
            z = (valid && l0 + k < cnt) ? z : 0.f;
            zq[k] = (k == kb && valid) ? 1.0f : z;
          }
          {
            __half h8[8], l8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              split_f16(zq[k], h8[k], l8[k]);
              zq[k] = __half2float(h8[k]) + __half2float(l8[k]);
            }
            const uint32_t off = core_off(r, ci, PK_ROWS);
#if PK_Z_TMEM
            // hi plane -> TMEM (lane = row, 32-bit column = 2 consecutive latents): read by two of the three MMA passes
            // of every chunk without touching shared memory
            tmem_st4(lane_addr + PK_TM_AZ + (i & 1) * 32 + ci * 4, reinterpret_cast<const uint32_t*>(h8));
#else
            *reinterpret_cast<uint4*>(az_hi + off) = *reinterpret_cast<const uint4*>(h8);
#endif
            *reinterpret_cast<uint4*>(az_lo + off) = *reinterpret_cast<const uint4*>(l8);
          }
          // regression sums of this warp's 32 rows (series uA and, past a boundary, uA + 1)
          if (col0 == 0) {
#pragma unroll 1
            for (int sidx = 0; sidx < 2; ++sidx) {
              float sum = 0.f;
              if (sidx ? anyB : anyA) sum = pk_lane_transpose_sum8(zq, (sidx ? inB : inA) ? xc : 0.f, lane);
              if (lane < 8) part[sidx * 64 + ci * 8 + lane] = (double)sum;
            }
          }
        }
        // a warp without a content chunk (class_dim <= 16) has not seen the class-head MMAs of this tile
        // complete: it must, before its P1 of the next tile overwrites their A operand in TMEM
        if (tile_need && !waited) { pk_wait(bar_heads_done, heads_waits & 1, s_abort); ++heads_waits; }
      }
      PK_T(4);
#if PK_Z_TMEM
      tmem_st_wait();
      tc_fence_before();
#endif
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) pk_arrive(bar_z_full + (i & 1));
      PK_T(6);
    }
#ifdef PK_PROF
    if ((warp == 0 || warp == 4) && lane == 0) for (int k = 0; k < 8; ++k) ws.phase[(blockIdx.x * 4 + (warp >> 2)) * 8 + k] = pc[k];
#endif
  } else if (warp < PK_W_HMMA) {
    // =============================== epilogue ===============================
    const int q4 = warp & 3;
    const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
    float* s_stage = reinterpret_cast<float*>(smem + pl.stage + q4 * PK_STAGE_BYTES);   // 1024-byte aligned (swizzle atom)
    int q = 0;
#pragma unroll 1
    for (int i = 0;; ++i) {
      const int gt = tile_of(i);
      if (gt < 0) break;
      int tile_row, tile_end;
      tile_rows(gt, tile_row, tile_end);
      const int rows_left = max(0, min(32, tile_end - (tile_row + q4 * 32)));
      const int tile_v = tile_row / rpv, row_in_val = tile_row - tile_v * rpv + q4 * 32;
#pragma unroll 1
      for (int ch = 0; ch < n_chunks; ++ch, ++q) {
        const int b = PK_ACC_BUFS == 2 ? (q & 1) : 0;
        PK_T(1);
        pk_wait(bar_acc_full + b, (PK_ACC_BUFS == 2 ? (q >> 1) : q) & 1, s_abort);
        PK_T(0);
        tc_fence_after();
        const int nsub = min(PK_NCH / 32, (ncol - ch * PK_NCH + 31) >> 5);
#pragma unroll 1
        for (int sub = 0; sub < nsub; ++sub) {
          const int cb0 = ch * PK_NCH + sub * 32;
          float vv[32];
          tmem_ld32(lane_base + PK_TM_ACC + b * PK_NCH + sub * 32, vv);
          tmem_ld_wait();
          if (sub == nsub - 1) {            // accumulator fully read: hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) pk_arrive(bar_acc_empty + b);
          }
          if (cx.avatars && tma_ok) {
            // TMA epilogue: the warp stages its 32 x 32 block in the 128-byte-swizzled box layout (lane = row,
            // 16-byte chunk k of the row at position k ^ (row & 7): conflict-free) and ONE lane hands the 4 KB box
            // to the copy engine (3-D map (column, row inside the validation, validation): rows past the end of a
            // validation and columns past R are clipped by the hardware).  No LDS / STG on the LSU path that the
            // producers' shared-memory traffic uses.
#ifndef PK_EXP_NOWAIT   // experiment (WRONG results): the gain an unbounded stage ring could give at most
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous box has left the stage
#endif
            __syncwarp();
            {
              unsigned char* rowp = reinterpret_cast<unsigned char*>(s_stage) + lane * 128;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                *reinterpret_cast<float4*>(rowp + ((k ^ (lane & 7)) << 4)) = make_float4(vv[4 * k], vv[4 * k + 1], vv[4 * k + 2], vv[4 * k + 3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && rows_left > 0) {
              asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmap)), "r"(col0 + cb0), "r"(row_in_val), "r"(tile_v), "r"(smem_u32(s_stage)) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else if (cx.avatars) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              *reinterpret_cast<float4*>(s_stage + lane * PK_STAGE_LD + 4 * k) = make_float4(vv[4 * k], vv[4 * k + 1], vv[4 * k + 2], vv[4 * k + 3]);
            __syncwarp();
            const int rr0 = lane >> 3, cg = (lane & 7) * 4;
            float* gbase = cx.avatars + (int64_t)(tile_row + q4 * 32 + rr0) * R + col0 + cb0 + cg;
            const float* sbase = s_stage + rr0 * PK_STAGE_LD + cg;
            if (rows_left == 32 && (R & 3) == 0 && cb0 + 32 <= ncol) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
#ifdef PK_EXP_NOSTG   // experiment: staging without the global stores (the loaded value is kept alive through a never-true store)
                { const float4 o_ = *reinterpret_cast<const float4*>(sbase + it * 4 * PK_STAGE_LD); if (o_.x == 1.2345e-30f) __stcs(reinterpret_cast<float4*>(gbase), o_); }
#elif defined(PK_EXP_WRAP)   // experiment: same stores into a 32 MB (L2-resident) window of the tensor
                __stcs(reinterpret_cast<float4*>(cx.avatars + (((gbase - cx.avatars) + (int64_t)it * 4 * R) & 0x7FFFFF)), *reinterpret_cast<const float4*>(sbase + it * 4 * PK_STAGE_LD));
#else
                __stcs(reinterpret_cast<float4*>(gbase + (int64_t)it * 4 * R), *reinterpret_cast<const float4*>(sbase + it * 4 * PK_STAGE_LD));
#endif
            } else {
#pragma unroll 1
              for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + rr0;
                if (rr < rows_left) {
                  const float4 o = *reinterpret_cast<const float4*>(sbase + it * 4 * PK_STAGE_LD);
                  float* dstp = gbase + (int64_t)it * 4 * R;
                  if ((R & 3) == 0 && cb0 + cg + 4 <= ncol) __stcs(reinterpret_cast<float4*>(dstp), o);
                  else {
                    if (cb0 + cg + 0 < ncol) dstp[0] = o.x;
                    if (cb0 + cg + 1 < ncol) dstp[1] = o.y;
                    if (cb0 + cg + 2 < ncol) dstp[2] = o.z;
                    if (cb0 + cg + 3 < ncol) dstp[3] = o.w;
                  }
                }
              }
            }
            __syncwarp();
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef PK_PROF
    if (warp == PK_W_EPI && lane == 0) { pc[2] = clock64() - tstart; for (int k = 0; k < 4; ++k) ws.phase[(blockIdx.x * 4 + 2) * 8 + k] = pc[k]; }
#endif
  } else if (warp == PK_W_HMMA) {
    // =============================== class-head MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc_h = idesc_f16(PK_ROWS, NH);
      const uint32_t LBO_BH = (NH / 8) * 128;
      uint32_t hc = 0;
#pragma unroll 1
      for (int i = 0;; ++i) {
        const int gt = tile_of(i);
        if (gt < 0) break;
        int uA, uB;
        tile_units(gt, uA, uB);
        if (!(unit_need(uA) || unit_need(uB))) continue;
        pk_wait(bar_h_full, hc & 1, s_abort);
        ++hc;
        tc_fence_after();
#pragma unroll 4
        for (int ks = 0; ks < MOPOE_HIDDEN / 16; ++ks) {
          const uint32_t ob = ks * 2 * LBO_BH;
          const uint64_t dbh = smem_desc(smem_u32(s_bh_hi) + ob, LBO_BH, 128), dbl = smem_desc(smem_u32(s_bh_lo) + ob, LBO_BH, 128);
          const uint32_t ah = tmem + PK_TM_AH_HI + ks * 8, al = tmem + PK_TM_AH_LO + ks * 8;
          mma_f16_ts(tmem + PK_TM_HEADS, ah, dbh, idesc_h, ks ? 1u : 0u);
          mma_f16_ts(tmem + PK_TM_HEADS, ah, dbl, idesc_h, 1u);
          mma_f16_ts(tmem + PK_TM_HEADS, al, dbh, idesc_h, 1u);
        }
        mma_commit(bar_heads_done);
      }
    }
  } else if (warp == PK_W_DMMA) {
    // =============================== decoder MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc_d = idesc_f16(PK_ROWS, PK_NCH);
      const uint32_t LBO_A = (PK_ROWS / 8) * 128, LBO_BD = (PK_CBP / 8) * 128;
      int q = 0, n_tiles = 0;
#pragma unroll 1
      for (int i = 0;; ++i) {
        if (tile_of(i) < 0) break;
        n_tiles = i + 1;
        PK_T(2);
        pk_wait(bar_z_full + (i & 1), (i >> 1) & 1, s_abort);
        PK_T(0);
        tc_fence_after();
        const uint32_t az_hi = smem_u32(s_az + (i & 1) * 2 * AZ_PLANE), az_lo = az_hi + AZ_PLANE;
#pragma unroll 1
        for (int ch = 0; ch < n_chunks; ++ch, ++q) {
          const int b = PK_ACC_BUFS == 2 ? (q & 1) : 0;
          PK_T(2);
          pk_wait(bar_acc_empty + b, ((PK_ACC_BUFS == 2 ? (q >> 1) : q) & 1) ^ 1, s_abort);
          PK_T(1);
          tc_fence_after();
          const uint32_t dcol = tmem + PK_TM_ACC + b * PK_NCH;
          for (int ks = 0; ks < KZ / 16; ++ks) {
            const uint32_t oa = ks * 2 * LBO_A, ob = ks * 2 * LBO_BD + ch * (PK_NCH / 8) * 128;
            const uint64_t dah = smem_desc(az_hi + oa, LBO_A, 128), dal = smem_desc(az_lo + oa, LBO_A, 128);
            const uint64_t dbh = smem_desc(smem_u32(s_bd_hi) + ob, LBO_BD, 128), dbl = smem_desc(smem_u32(s_bd_lo) + ob, LBO_BD, 128);
#if PK_Z_TMEM
            const uint32_t tah = tmem + PK_TM_AZ + (i & 1) * 32 + ks * 8;
            mma_f16_ts(dcol, tah, dbh, idesc_d, ks ? 1u : 0u);
            mma_f16_ts(dcol, tah, dbl, idesc_d, 1u);
            (void)dah;
#else
            mma_f16(dcol, dah, dbh, idesc_d, ks ? 1u : 0u);
            mma_f16(dcol, dah, dbl, idesc_d, 1u);
#endif
            mma_f16(dcol, dal, dbh, idesc_d, 1u);
          }
          mma_commit(bar_acc_full + b);
        }
        mma_commit(bar_z_free + (i & 1));
      }
      // every MMA has completed before the CTA tears down TMEM / shared memory
      if (n_tiles >= 1) pk_wait(bar_z_free + ((n_tiles - 1) & 1), ((n_tiles - 1) >> 1) & 1, s_abort);
#ifdef PK_PROF
      for (int k = 0; k < 4; ++k) ws.phase[(blockIdx.x * 4 + 2) * 8 + 4 + k] = pc[k];
#endif
    }
  } else {
    // =============================== aux: series caches, regression sums ===============================
    auto build = [&](int u, int slot) {
      const int uc = u % C, ug = (u / C) % N, uv = u / (C * N);
      // the record was written by daa_base_kernel (series_records): 161 16-byte loads in flight at once (one
      // L2 round trip), then the shared-memory stores
      const float4* rec = reinterpret_cast<const float4*>(ws.srec + (int64_t)u * PK_REC_F);
      float4* cache4 = reinterpret_cast<float4*>(s_cache + slot * PK_CACHE_F);
      constexpr int NV = PK_REC_F / 4, NIT = (NV + 31) / 32;
      float4 vbuf[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int q = it * 32 + lane;
        vbuf[it] = q < NV ? __ldg(rec + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int q = it * 32 + lane;
        if (q < PK_CACHE_F / 4) cache4[q] = vbuf[it];
        else if (q == PK_CACHE_F / 4) {
          s_meta[slot * 4] = __float_as_int(vbuf[it].z);
          s_xbar[slot] = __hiloint2double(__float_as_int(vbuf[it].y), __float_as_int(vbuf[it].x));
          const_cast<int64_t*>(s_rbase)[slot] = ((int64_t)(cx.v_av_off + uv) * J * C + uc) * N + ug;
        }
      }
    };
    // reduce the 4 row quarters of a finished tile (global id gt, parity par of its iteration) and store its
    // contribution to each of its (<= 2) series: slot (series, tile index inside the series), summed in fixed
    // order by daa_beta_stats_kernel, so the tables do not depend on which CTA ran the tile
    auto fold = [&](int gt, int par) {
      int uA, uB;
      tile_units(gt, uA, uB);
      for (int s = 0; s <= uB - uA; ++s) {
        const int u = uA + s, uv = u / (N * C);
        const int tp = gt - (uv * tpv + (int)(((int64_t)(u - uv * N * C) * J) / PK_ROWS));   // tile index inside the series
        double* o = ws.sacc + ((int64_t)u * tpu + tp) * 64;
        for (int k = lane; k < KZ; k += 32) {
          double a = 0.0;
          for (int w = 0; w < 4; ++w) a += s_part[((par * 4 + w) * 2 + s) * 64 + k];
          o[k] = a;
        }
      }
    };
    // iteration i of this CTA is global tile gt (or -1: stop): tile info, the records of its (<= 2) series into
    // the cache slots of its parity, its 128 scores; then the producers (bar_cache) and the lagging roles (ring +
    // published count) are released
    auto publish = [&](int i, int gt) {
      int* ti = s_tinfo + (i & 1) * 4;
      if (gt >= 0) {
        int r0, r1;
        tile_rows(gt, r0, r1);
        const int uA = r0 / J, uB = (r1 - 1) / J;
        if (lane == 0) { ti[0] = uA; ti[1] = uB; ti[2] = r0; ti[3] = r1; }
        float sc[PK_ROWS / 32];
#pragma unroll
        for (int k = 0; k < PK_ROWS / 32; ++k) sc[k] = r0 + k * 32 + lane < r1 ? ws.scores[r0 + k * 32 + lane] : 0.f;
#pragma unroll 1
        for (int u = uA; u <= uB; ++u) build(u, (i & 1) * 2 + (u - uA));
#pragma unroll
        for (int k = 0; k < PK_ROWS / 32; ++k) s_score[(i & 1) * PK_ROWS + k * 32 + lane] = sc[k];
      } else if (lane == 0) {
        ti[2] = -1;
      }
      if (lane == 0) s_ring[i & 7] = gt;
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        *s_npub = i + 1;
        pk_arrive(bar_cache + (i & 1));
      }
    };
    // shard units (mopoe_daa_desc.unit_begin / unit_end): a tile is run iff one of its (<= 2) series belongs to an
    // owned (validation, score) unit; tiles of the other scores of a shared validation are skipped (another rank
    // runs them).  Lane 0 only: draws from the global counter until it holds an owned tile.
    const int own0 = cx.q.unit_begin, own1 = cx.q.unit_end;
    const bool own_all = own0 == 0 && own1 == cx.q.n_val * C;
    auto claim = [&](int gt) -> int {
      while (gt < total_tiles && !own_all) {
        int uA, uB;
        tile_units(gt, uA, uB);
        bool mine = false;
        for (int u = uA; u <= uB; ++u) {
          const int unit = (u / (N * C)) * C + u % C;
          mine = mine || (unit >= own0 && unit < own1);
        }
        if (mine) break;
        gt = (int)gridDim.x + atomicAdd(ws.counter, 1);
      }
      return gt < total_tiles ? gt : -1;
    };
    int gt_cur = 0, gt_prev = -1, n_done = 0;
    if (lane == 0) gt_cur = claim((int)blockIdx.x);
    gt_cur = __shfl_sync(0xffffffffu, gt_cur, 0);
    publish(0, gt_cur);
#pragma unroll 1
    for (int i = 0; gt_cur >= 0; ++i) {
      // the next tile of this CTA: the round trip of the atomic overlaps the wait below
      int gt_next = 0;
      if (lane == 0) {
        gt_next = claim((int)gridDim.x + atomicAdd(ws.counter, 1));
      }
      gt_next = __shfl_sync(0xffffffffu, gt_next, 0);
      PK_T(1);
      // every producer has finished tile i-1 (its z_full arrivals): its partial sums are complete, and the
      // cache slots / tile info / scores of its parity, which iteration i+1 reuses, are no longer read
      if (i > 0) {
        pk_wait(bar_z_full + ((i - 1) & 1), ((i - 1) >> 1) & 1, s_abort);
        if (col0 == 0) fold(gt_prev, (i - 1) & 1);
      }
      PK_T(0);
      publish(i + 1, gt_next);
      gt_prev = gt_cur; gt_cur = gt_next; n_done = i + 1;
    }
    if (n_done > 0) {
      pk_wait(bar_z_full + ((n_done - 1) & 1), ((n_done - 1) >> 1) & 1, s_abort);
      if (col0 == 0) fold(gt_prev, (n_done - 1) & 1);
    }
#ifdef PK_PROF
    if (lane == 0) for (int k = 0; k < 2; ++k) ws.phase[(blockIdx.x * 4 + 3) * 8 + k] = pc[k];
#endif
  }
  tc_fence_before();
  __syncthreads();
#ifdef PK_PROF
  if (t == 0) {
    unsigned long long g_exit;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_exit));
    ws.phase[(blockIdx.x * 4 + 3) * 8 + 2] = tstart - t_entry;          // prologue cycles
    ws.phase[(blockIdx.x * 4 + 3) * 8 + 3] = clock64() - t_entry;       // whole CTA cycles
    ws.phase[(blockIdx.x * 4 + 3) * 8 + 4] = (long long)g_entry;        // ns
    ws.phase[(blockIdx.x * 4 + 3) * 8 + 5] = (long long)g_exit;
  }
#endif
  if (t == 0 && *s_abort) atomicExch(ws.err, 1);
  if (warp == 0) tmem_dealloc(tmem, 512);
}

#ifndef BS_ROIS_
#define BS_ROIS_ 512
#endif
// Per-subject slopes and the second-level t statistic of the pipelined path:
//   beta[v,c,g,roi] = (Wd[roi,:] . sum_tiles sacc[u,tile,:]) / Sxx[u]            (stat_utils.py:66-68)
//   coef = mean_g beta, t = coef / (sd_g(beta) / sqrt(N))                         (stat_utils.py:73-75)
// (the p-value 2 sf(|t|, N-1) is evaluated by daa_pvalue_kernel, one thread per statistic).
// CTA = (validation, score) x block of <= 512 ROIs.  The decoder weights of the block sit in shared memory as
// fp32, transposed to the K order of z (+ bias slot), row stride odd (conflict-free transpose and reads);
// thread = 2 ROIs (t, t + blockDim), subjects in batches of BS_GB = 25: per k step 2 weight loads + 13 broadcast
// float4 loads feed 100 FMAs.
// Arithmetic: the regression sums carry fp32-level rounding from their 32-row partial sums anyway, so the
// 149 M-term contraction is done with fp32 FMAs on the sums split into hi + lo fp32 parts (48 bits of each sum
// enter the product; accumulation error ~1e-7 of the absolute term sum, the same order as the rounding already
// in the sums); the sum over tiles, the division by Sxx, the subject statistics (shifted sums) and the t
// statistic are fp64.
#ifndef BS_SPLIT_LO
#define BS_SPLIT_LO 0
#endif
constexpr int BS_ROIS = BS_ROIS_, BS_GB = 25, BS_GP = 28;    // ROIs per CTA, subjects per batch, padded batch stride

__device__ double two_sided_t_pvalue(double tval, double nu);

__global__ void __launch_bounds__(256) daa_beta_stats_kernel(ModelView mv, int dst, int R, int C, int N, int J, UmmaDims dm, int unit0,
                                                              const double* sacc, const double* xstat, double* betas,
                                                              double* coefs, double* tvals) {
  extern __shared__ __align__(16) float s_dyn[];
  const ModView& md = mv.mod[dst];
  const int t = threadIdx.x, nt = blockDim.x;
  const int vc = unit0 + blockIdx.x;                       // (validation, score) unit of this CTA
  const int v = vc / C, c = vc % C;
  const int c0 = blockIdx.y * BS_ROIS, nc = min(BS_ROIS, R - c0);
  const int tpu = pipe_tiles_per_unit(J), KZ = dm.KZ;
  const int RP = nc | 1;                                   // odd row stride
  float* s_wf = s_dyn;                                     // [KZ][RP]
  float* s_hi = s_wf + ((KZ * RP + 3) & ~3);               // [KZ][BS_GP] regression sums of the batch, k-major: fp32 hi part
  float* s_lo = s_hi + KZ * BS_GP;                         //                                                     fp32 lo part
  double* s_sxx = reinterpret_cast<double*>(s_lo + KZ * BS_GP);   // [BS_GB] 1 / Sxx
  for (int i = t; i < KZ * RP; i += nt) s_wf[i] = 0.f;
  __syncthreads();
  {
    // coalesced: the block's weight rows are contiguous.  Many independent loads in flight per thread (every
    // iteration is an L2 round trip otherwise: 80 of them for 444 ROIs)
    const float* wsrc = md.wd + (int64_t)c0 * md.ZD;
    const int ntot = nc * md.ZD;
    auto put = [&](int i, float w) {
      const int col = i / md.ZD, zd = i % md.ZD;
      const int kz = zd < md.S ? dm.KC + zd : zd - md.S;   // decoder input = [style | content], z = [content | style]
      s_wf[kz * RP + col] = w;
    };
    if ((ntot & 3) == 0 && (reinterpret_cast<uintptr_t>(wsrc) & 15) == 0) {
      const float4* w4 = reinterpret_cast<const float4*>(wsrc);
      const int n4 = ntot >> 2;
      for (int i0 = 0; i0 < n4; i0 += nt * 8) {
        float4 buf[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int i = i0 + q * nt + t; buf[q] = i < n4 ? __ldg(w4 + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = i0 + q * nt + t;
          if (i < n4) {      // one division per float4: (column, latent) of its first element, then stepped
            int col = (4 * i) / md.ZD, zd = 4 * i - col * md.ZD;
            const float wv[4] = {buf[q].x, buf[q].y, buf[q].z, buf[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int kz = zd < md.S ? dm.KC + zd : zd - md.S;
              s_wf[kz * RP + col] = wv[e];
              if (++zd == md.ZD) { zd = 0; ++col; }
            }
          }
        }
      }
    } else {
      for (int i = t; i < ntot; i += nt) put(i, wsrc[i]);
    }
  }
  if (dm.bias_slot >= 0) for (int i = t; i < nc; i += nt) s_wf[dm.bias_slot * RP + i] = md.bd[c0 + i];
  const int r0 = t, r1 = t + nt;                           // this thread's ROIs inside the block
  const bool act0 = r0 < nc, act1 = r1 < nc;
  const int q0 = act0 ? r0 : 0, q1 = act1 ? r1 : 0;
  double b0[2] = {0.0, 0.0}, sd1[2] = {0.0, 0.0}, sd2[2] = {0.0, 0.0};
  for (int g0 = 0; g0 < N; g0 += BS_GB) {
    const int ng = min(BS_GB, N - g0);
    __syncthreads();
    // sums over the tiles of each series: 4 items x up to 4 tile slots in flight per thread (independent loads)
    for (int i0 = 0; i0 < KZ * BS_GB; i0 += nt * 4) {
      double part[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + q * nt + t;
        const int k = i % KZ, gb = i / KZ;
        const bool on = i < KZ * BS_GB && gb < ng;
        const int ul = ((g0 + gb) * C + c), u = v * N * C + ul;   // series index inside the validation / global
        const int first = (int)(((int64_t)ul * J) / PK_ROWS), last = (int)(((int64_t)(ul + 1) * J - 1) / PK_ROWS);
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) part[q][tp] = (on && tp <= last - first) ? sacc[((int64_t)u * tpu + tp) * 64 + k] : 0.0;
        if (on && last - first >= 4) for (int tp = 4; tp <= last - first; ++tp) part[q][0] += sacc[((int64_t)u * tpu + tp) * 64 + k];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + q * nt + t;
        if (i < KZ * BS_GB) {
          const int k = i % KZ, gb = i / KZ;
          const double a = ((part[q][0] + part[q][1]) + part[q][2]) + part[q][3];
          const float hi = (float)a;
          s_hi[k * BS_GP + gb] = hi;
#if BS_SPLIT_LO
          s_lo[k * BS_GP + gb] = (float)(a - (double)hi);
#endif
        }
      }
    }
    // 1 / Sxx once per subject: an fp64 division per slope would cost as much as the whole contraction
    if (t < BS_GB) s_sxx[t] = t < ng ? 1.0 / xstat[(((int64_t)v * C + c) * N + g0 + t) * 2 + 1] : 1.0;
    __syncthreads();
    // fp32 contraction of the fp32-rounded sums: the 48-term fp32 accumulation rounds at ~3e-7 of the slope, which a lo
    // part of the sums (BS_SPLIT_LO: 6e-8 each) cannot improve -- it only doubled the FMAs and the accumulator registers
    float ah0[BS_GB], ah1[BS_GB];
#if BS_SPLIT_LO
    float al0[BS_GB], al1[BS_GB];
#pragma unroll
    for (int gb = 0; gb < BS_GB; ++gb) al0[gb] = al1[gb] = 0.f;
#endif
#pragma unroll
    for (int gb = 0; gb < BS_GB; ++gb) ah0[gb] = ah1[gb] = 0.f;
#pragma unroll 2
    for (int k = 0; k < KZ; ++k) {
      const float w0 = s_wf[k * RP + q0], w1 = s_wf[k * RP + q1];
      const float4* hp = reinterpret_cast<const float4*>(s_hi + k * BS_GP);
#if BS_SPLIT_LO
      const float4* lp = reinterpret_cast<const float4*>(s_lo + k * BS_GP);
#endif
#pragma unroll
      for (int q = 0; q < BS_GP / 4; ++q) {
        const float4 h = hp[q];
        const float hv[4] = {h.x, h.y, h.z, h.w};
#if BS_SPLIT_LO
        const float4 l = lp[q];
        const float lv[4] = {l.x, l.y, l.z, l.w};
#endif
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int gb = 4 * q + e;
          if (gb < BS_GB) {
            ah0[gb] = fmaf(hv[e], w0, ah0[gb]);
            ah1[gb] = fmaf(hv[e], w1, ah1[gb]);
#if BS_SPLIT_LO
            al0[gb] = fmaf(lv[e], w0, al0[gb]);
            al1[gb] = fmaf(lv[e], w1, al1[gb]);
#endif
          }
        }
      }
    }
#pragma unroll
    for (int gb = 0; gb < BS_GB; ++gb) {
      if (gb < ng) {
        const double isx = s_sxx[gb];
        double* brow = betas + (((int64_t)v * C + c) * N + g0 + gb) * R + c0;
        if (act0) {
#if BS_SPLIT_LO
          const double beta = ((double)ah0[gb] + (double)al0[gb]) * isx;
#else
          const double beta = (double)ah0[gb] * isx;
#endif
          brow[r0] = beta;
          if (g0 + gb == 0) b0[0] = beta;
          const double d = beta - b0[0];                  // shifted sums: no cancellation in the variance
          sd1[0] += d; sd2[0] = fma(d, d, sd2[0]);
        }
        if (act1) {
#if BS_SPLIT_LO
          const double beta = ((double)ah1[gb] + (double)al1[gb]) * isx;
#else
          const double beta = (double)ah1[gb] * isx;
#endif
          brow[r1] = beta;
          if (g0 + gb == 0) b0[1] = beta;
          const double d = beta - b0[1];
          sd1[1] += d; sd2[1] = fma(d, d, sd2[1]);
        }
      }
    }
  }
  const double n = (double)N;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    if (e == 0 ? act0 : act1) {
      const double mean = b0[e] + sd1[e] / n;
      const double ss = sd2[e] - sd1[e] * sd1[e] / n;
      const double sd = sqrt(ss / (n - 1.0));
      const int64_t o = ((int64_t)v * C + c) * R + c0 + (e == 0 ? r0 : r1);
      coefs[o] = mean;
      tvals[o] = mean / (sd / sqrt(n));
    }
  }
}

// p = 2 sf(|t|, nu) of every statistic, one thread each (62 160 independent continued fractions: run them
// all at once instead of one after the other inside the 140 slope CTAs); also poisons the tables when a
// tcgen05 kernel flagged a protocol error.
__device__ double two_sided_t_pvalue_fast(double tval, double nu, double lg_pref);

__global__ void __launch_bounds__(128) daa_pvalue_kernel(const int* err, double* coefs, double* pvalues, int64_t n, double nu, double lg_pref, int from_t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (*err) {
    coefs[i] = __longlong_as_double(0x7ff8000000000000LL);
    pvalues[i] = __longlong_as_double(0x7ff8000000000000LL);
  } else if (from_t) {
    pvalues[i] = two_sided_t_pvalue_fast(pvalues[i], nu, lg_pref);
  }
}

}  // namespace mopoe
