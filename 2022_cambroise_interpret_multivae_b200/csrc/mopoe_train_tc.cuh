// Tensor-core (tcgen05 / TMEM / TMA bulk copy) training step of the MoPoE-VAE for sm_100a.
// Included by mopoe_model.cu inside namespace mopoe (it uses ModelView, StepCtx, Workspace, LatSh, lat_forward,
// lat_backward, apply_grad, finalize_scalars, grid_barrier defined there).
//
// Same three phases as the CUDA-core kernel (P1 hidden layer, P2 per-row-tile forward + backward, P3 weight
// gradients + Adam), same reference semantics (run_epochs.py:73-135,180-182, BaseMMVae.py:137-239,
// networks.py:30-36,66-77), but every contraction is a tcgen05.mma:
//
//   formulation   D^T[feature, row] = W[feature, k] * act[row, k]^T : the WEIGHTS are the M side (128-lane M
//                 tiles, zero padded), the rows of the batch tile are the N side (R = 16 or 32 rows per tile),
//                 so a 256-row batch still gives 8-16 independent tiles and a 65 536-row batch 2 048.
//   operands      fp16 hi/lo planes ("3xFP16 split", mopoe_umma.cuh), fp32 accumulation in TMEM.  Weight planes
//                 are pre-split once per step into 16 KB chunks (128 rows x 32 K x {hi, lo}, UMMA canonical
//                 K-major layout) by tc_prep; a loader warp streams the chunks through a shared-memory ring
//                 with cp.async.bulk + mbarrier complete_tx, an MMA warp consumes them, 8 compute warps run the
//                 epilogues (thread = feature lane, registers = rows) and the latent stage (lat_forward /
//                 lat_backward, shared with the CUDA-core tile).
//   activations   every epilogue writes its result as the NEXT contraction's operand: layout T (16 bytes = 8
//                 consecutive rows of one feature), which is an MN-major B operand for the next GEMM of the tile
//                 (N = rows, K = features) and, stored to HBM, a K-major operand of the weight-gradient GEMMs of
//                 P3 (K = rows).  The input rows x are converted once (layout F: 16 bytes = 8 features of one
//                 row; K-major B of the first layer, MN-major B of dW1).
//   gradients     operands that carry a 1/N factor in the reference (d x_hat, d heads, d pre-activation) are kept
//                 N-times larger (fp16 range) and the factor is applied in fp32 in the consuming epilogue.
//   P3            output-stationary 128 x Nw tiles of dW1 / dWheads / dWdec with the batch as K, split over row
//                 ranges for large batches (partials summed in fixed order by the last CTA to arrive:
//                 deterministic), Adam in the epilogue; bias / output-log-variance gradients are per-tile column
//                 sums taken from the epilogue registers of P2 and reduced in fixed order.
#pragma once

namespace tc {

using namespace umma;

// -DTC_PROF: CTA 0 / thread 0 accumulates cycle counts per stage into the 64 floats behind the error flag
// (read back with mopoe_debug_tcprof); no effect on results
#ifdef TC_PROF
#define TCW(i, stmt) do { const long long _w0 = clock64(); stmt; if (blockIdx.x == 0) (reinterpret_cast<float*>(pl.base + pl.err) + 64)[i] += (float)(clock64() - _w0); } while (0)
#define TCP_DECL long long _tp = clock64()
#define TCP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long _n = clock64(); prof[i] += (float)(_n - _tp); _tp = _n; } } while (0)
#else
#define TCW(i, stmt) stmt
#define TCP_DECL do { } while (0)
#define TCP(i) do { } while (0)
#endif

constexpr int THREADS = 320;       // warps 0-7 compute, warp 8 loader, warp 9 MMA issuer
constexpr int CHUNK = 16384;       // one streamed weight chunk: 128 rows x 32 K x {hi, lo} fp16
constexpr int MAX_SLOTS = 10;
constexpr int MAX_UNITS = 96;

struct TcMod {
  int NCx, NCz, NCh, MtD;                      // K chunks (of 32) of W1 / Wdec / Wheads^T; decoder M tiles
  int Dk16, ZDk16, HCk16;                      // K extents rounded up to the MMA K (16)
  int64_t w1, wh, wht, wd, wdt;                // weight chunk blobs (byte offsets from TcPlan::base)
  int64_t xop, hop, daop, deop, dxop, zzop;    // activation operand blobs of row tile 0 (byte offsets)
  int xop_t, dxop_t;                           // bytes per row tile of xop / of one pass of dxop
  int col0;                                    // first column of this modality in a column-partial row
};

struct TcUnit { unsigned char m, g, mt, nb; };  // weight-gradient output tile: modality, GEMM, M tile, N block

struct TcPlan {
  TcMod mod[MOPOE_MAX_MODS];
  TcUnit unit[MAX_UNITS];
  int n_units, ksplit, nw[3], slot3;           // P3: units of the full model, row-range splits, N widths, ring slot bytes
  int R, nslot, np;                            // rows per tile, ring slots, decoder passes (2 for poe)
  int ntiles_max, ccols;
  int64_t colpart, p3part, p3cnt, err;         // byte offsets
  int64_t total;
  int s_ring, s_u, s_e, s_zz, s_dzz, s_rp, s_rps, s_mask, s_srow, s_red, s_bar, s_total;   // shared memory plan (bytes)
  int hcm, zdm, sm_;
  unsigned char* base;
};

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- host: plan ----------------------------------------------------------------------------
// returns false if the configuration does not fit the tensor-core tiling (caller falls back / errors)
static bool make_plan(const mopoe_model_desc* d, int64_t max_rows, int smem_limit, TcPlan* out) {
  TcPlan p;
  memset(&p, 0, sizeof(p));
  const int M = d->n_mods, L = d->latent_dim;
  if (L > 32) return false;
  const bool uni = d->method == MOPOE_METHOD_POE;
  p.np = uni ? 2 : 1;
  int hcm = 0, zdm = 0, sm = 1;
  for (int m = 0; m < M; ++m) {
    const int S = d->style_dims[m];
    if (S > 32) return false;
    hcm = max(hcm, 2 * L + 2 * S); zdm = max(zdm, S + L); sm = max(sm, S);
  }
  p.hcm = hcm; p.zdm = zdm; p.sm_ = sm;
  // shared memory: largest R in {32, 16} whose plan fits with at least 4 ring slots
  int R = 0;
  for (int cand : {32, 16}) {
    if (cand == 32 && max_rows <= 2048) continue;   // small batches: more, smaller tiles (each tile re-streams the weights at the per-SM rate)
    int off = 0;
    auto take = [&](int n) { int o = off; off += (n + 127) & ~127; return o; };
    p.s_bar = take(640);
    p.s_red = take(MOPOE_N_SCALARS * 4);
    p.s_srow = take(M * cand * 8);
    p.s_mask = take(M * 256 * 4);
    p.s_e = take(M * cand * hcm * 4);
    p.s_zz = take(M * p.np * cand * zdm * 4);
    p.s_dzz = take(M * p.np * cand * zdm * 4);
    p.s_rp = take((uni ? 1 + M : 1) * cand * L * 4);
    p.s_rps = take(M * p.np * cand * sm * 4);
    p.s_u = take(1280 * cand);
    off = (off + 1023) & ~1023;
    p.s_ring = off;
    const int slots = (smem_limit - off) / CHUNK;
    if (slots >= 4) { R = cand; p.nslot = slots > MAX_SLOTS ? MAX_SLOTS : slots; p.s_total = off + p.nslot * CHUNK; break; }
  }
  if (!R) return false;
  p.R = R;
  p.ntiles_max = (int)((max_rows + R - 1) / R);
  int64_t off = 0;
  auto takeg = [&](int64_t n) { int64_t o = off; off += (n + 255) & ~(int64_t)255; return o; };
  p.err = takeg(1024);   // [0] error flag, [64..127] profiling counters (floats)
  int col = 0;
  for (int m = 0; m < M; ++m) {
    TcMod& t = p.mod[m];
    const int D = d->dims[m], S = d->style_dims[m], HC = 2 * L + 2 * S, ZD = S + L;
    if (HC > 128 || ZD > 64) return false;
    t.NCx = cdiv(D, 32); t.NCz = cdiv(ZD, 32); t.NCh = cdiv(HC, 32); t.MtD = cdiv(D, 128);
    t.Dk16 = cdiv(D, 16) * 16; t.ZDk16 = cdiv(ZD, 16) * 16; t.HCk16 = cdiv(HC, 16) * 16;
    t.w1 = takeg((int64_t)2 * t.NCx * CHUNK);
    t.wh = takeg((int64_t)8 * CHUNK);
    t.wht = takeg((int64_t)2 * t.NCh * CHUNK);
    t.wd = takeg((int64_t)t.MtD * t.NCz * CHUNK);
    t.wdt = takeg((int64_t)t.NCx * CHUNK);
    t.xop_t = t.NCx * 32 * R * 4;
    t.dxop_t = t.MtD * 128 * R * 4;
    t.xop = takeg((int64_t)p.ntiles_max * t.xop_t);
    t.hop = takeg((int64_t)p.ntiles_max * 256 * R * 4);
    t.daop = takeg((int64_t)p.ntiles_max * 256 * R * 4);
    t.deop = takeg((int64_t)p.ntiles_max * 128 * R * 4);
    t.dxop = takeg((int64_t)p.ntiles_max * p.np * t.dxop_t);
    t.zzop = takeg((int64_t)p.ntiles_max * p.np * 64 * R * 4);
    t.col0 = col;
    col += 256 + 128 + 2 * t.MtD * 128;
  }
  p.ccols = col;
  p.colpart = takeg((int64_t)2 * p.ntiles_max * col * 4);
  // P3 units (full model; absent modalities are skipped per step)
  const int nwmax = max_rows <= 1024 ? 64 : 256;
  p.nw[2] = 64;
  int nu = 0;
  int nw0max = 16, nw1 = 256 / cdiv(256, nwmax);
  for (int m = 0; m < M; ++m) {
    const TcMod& t = p.mod[m];
    const int nb0 = cdiv(t.Dk16, nwmax);
    const int nw0 = cdiv(cdiv(t.Dk16, nb0), 16) * 16;
    nw0max = max(nw0max, nw0);
    for (int mt = 0; mt < 2; ++mt)
      for (int nb = 0; nb * nw0 < t.Dk16; ++nb) { if (nu >= MAX_UNITS) return false; p.unit[nu++] = {(unsigned char)m, 0, (unsigned char)mt, (unsigned char)nb}; }
    for (int nb = 0; nb * nw1 < 256; ++nb) { if (nu >= MAX_UNITS) return false; p.unit[nu++] = {(unsigned char)m, 1, 0, (unsigned char)nb}; }
    for (int mt = 0; mt < t.MtD; ++mt) { if (nu >= MAX_UNITS) return false; p.unit[nu++] = {(unsigned char)m, 2, (unsigned char)mt, 0}; }
  }
  p.n_units = nu;
  {   // heaviest units first (cost ~ operand bytes per row): longest-processing-time order for the dynamic queue
    auto cost = [&](const TcUnit& u) {
      const TcMod& t = p.mod[u.m];
      if (u.g == 0) { const int nb0 = cdiv(t.Dk16, nwmax); return 128 + cdiv(cdiv(t.Dk16, nb0), 16) * 16; }
      if (u.g == 1) return 128 + nw1;
      return (128 + t.ZDk16) * p.np;
    };
    std::stable_sort(p.unit, p.unit + nu, [&](const TcUnit& a, const TcUnit& b) { return cost(a) > cost(b); });
  }
  p.nw[0] = nwmax; p.nw[1] = nw1;   // nw[0] is the CAP of the per-modality width of GEMM 0 (recomputed on the device)
  p.slot3 = (128 + max(max(nw0max, nw1), 64)) * R * 4;
  const int sms = num_sms();
  p.ksplit = max_rows <= 512 ? 1 : max(1, min(p.ntiles_max, (2 * sms) / max(1, nu)));   // ~2 items per CTA (4 measured slower: more partial tiles to reduce)
  p.p3part = takeg((int64_t)nu * p.ksplit * 128 * 256 * 4);
  p.p3cnt = takeg((int64_t)(MAX_UNITS + 64) * 4);   // per-unit arrival counters, then the P3 work-queue counter
  p.total = off;
  *out = p;
  return true;
}

// ---- device helpers ------------------------------------------------------------------------
struct Bars {
  uint64_t ring_full[MAX_SLOTS], ring_empty[MAX_SLOTS], xfull[2], xempty[2], b_ready, acc_done;
  uint64_t ring3_full[MAX_SLOTS], ring3_empty[MAX_SLOTS], acc_free;   // P3: its own ring (other slot size), accumulator hand-back
  uint32_t tmem_slot;
  int dead;
};

__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// bounded wait: a protocol error marks the CTA dead (all later waits fall through) and raises the device error
// flag; the step's loss is poisoned with NaN at the end of the launch instead of hanging the GPU
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity, Bars* bars, int* gerr) {
  const uint32_t addr = smem_u32(bar);
  if (*(volatile int*)&bars->dead) return;
  const long long t0 = clock64();
#pragma unroll 1
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 2000000000LL || *(volatile int*)&bars->dead) {
      *(volatile int*)&bars->dead = 1;
      atomicExch(gerr, 1);
      return;
    }
  }
}

// instruction descriptor: kind::f16, fp16 operands, fp32 accumulator, M = 128, N = n; a/b major bits (1 = MN-major)
__device__ __forceinline__ uint32_t idesc(int n, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// 8 values -> packed fp16 hi / lo words of the 3xFP16 split
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  split_pack2(x[0], x[1], hi.x, lo.x); split_pack2(x[2], x[3], hi.y, lo.y);
  split_pack2(x[4], x[5], hi.z, lo.z); split_pack2(x[6], x[7], hi.w, lo.w);
}

template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
  if (NC == 32) tmem_ld32(taddr, v);
  else if (NC == 16) tmem_ld16(taddr, v);
  else tmem_ld8(taddr, v);
}

// ---- weight chunk blobs --------------------------------------------------------------------
// matrix W (mrows x kcols), element (r, c) = src[r * ld_r + c * ld_c], as Mt x NC chunks of 16 KB:
// chunk (t, cc) at (t * NC + cc) * CHUNK; inside: half * 8192 + plane * 2048 + (rr / 8) * 128 + (rr % 8) * 16 + (c % 8) * 2
__device__ void prep_blob(unsigned char* blob, const float* src, int mrows, int kcols, int64_t ld_r, int64_t ld_c,
                          int Mt, int NC, int gtid, int gthreads) {
  const int groups = Mt * NC * 4 * 128;   // 16-byte groups per half
  for (int g = gtid; g < groups; g += gthreads) {
    const int rr = g & 127, pl = (g >> 7) & 3, ch = g >> 9;
    const int t = ch / NC, cc = ch - t * NC;
    const int r = t * 128 + rr, c0 = cc * 32 + pl * 8;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (r < mrows && c0 + i < kcols) ? src[r * ld_r + (c0 + i) * ld_c] : 0.f;
    uint4 hi, lo;
    split8(x, hi, lo);
    unsigned char* dst = blob + (int64_t)ch * CHUNK + pl * 2048 + (rr >> 3) * 128 + (rr & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + 8192) = lo;
  }
}

// all weight blobs of the model as ONE index space of 16-byte groups (a thread converts one or two groups per step,
// all its loads independent); the blob table is built in shared memory by the first threads of the CTA
struct PrepBlob { unsigned char* blob; const float* src; int mrows, kcols, ld_r, ld_c, NC, first; };

__device__ void tc_prep(const ModelView& mv, const TcPlan& pl, PrepBlob* tab) {
  const int t = threadIdx.x;
  if (t == 0) {
    int first = 0, nb = 0;
    for (int m = 0; m < mv.M; ++m) {
      const ModView& md = mv.mod[m];
      const TcMod& tm = pl.mod[m];
      auto add = [&](int64_t off, const float* src, int mrows, int kcols, int ld_r, int ld_c, int Mt, int NC) {
        tab[nb].blob = pl.base + off; tab[nb].src = src; tab[nb].mrows = mrows; tab[nb].kcols = kcols;
        tab[nb].ld_r = ld_r; tab[nb].ld_c = ld_c; tab[nb].NC = NC; tab[nb].first = first;
        first += Mt * NC * 512; ++nb;
      };
      add(tm.w1, md.w1, MOPOE_HIDDEN, md.D, md.D, 1, 2, tm.NCx);
      add(tm.wh, md.wh, md.HC, MOPOE_HIDDEN, MOPOE_HIDDEN, 1, 1, 8);
      add(tm.wht, md.wh, MOPOE_HIDDEN, md.HC, 1, MOPOE_HIDDEN, 2, tm.NCh);
      add(tm.wd, md.wd, md.D, md.ZD, md.ZD, 1, tm.MtD, tm.NCz);
      add(tm.wdt, md.wd, md.ZD, md.D, 1, md.ZD, 1, tm.NCx);
    }
    tab[nb].first = first;   // sentinel: total number of groups
    tab[nb].blob = nullptr;
  }
  __syncthreads();
  const int nb = 5 * mv.M, total = tab[nb].first;
  for (int g0 = blockIdx.x * blockDim.x + t; g0 < total; g0 += gridDim.x * blockDim.x) {
    int bi = 0;
    for (int k = 1; k < nb; ++k) if (g0 >= tab[k].first) bi = k;
    const PrepBlob& pb = tab[bi];
    const int g = g0 - pb.first;
    const int rr = g & 127, plane = (g >> 7) & 3, ch = g >> 9;
    const int tt = ch / pb.NC, cc = ch - tt * pb.NC;
    const int r = tt * 128 + rr, c0 = cc * 32 + plane * 8;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = (r < pb.mrows && c0 + i < pb.kcols) ? pb.src[(int64_t)r * pb.ld_r + (int64_t)(c0 + i) * pb.ld_c] : 0.f;
    uint4 hi, lo;
    split8(x, hi, lo);
    unsigned char* dst = pb.blob + (int64_t)ch * CHUNK + plane * 2048 + (rr >> 3) * 128 + (rr & 7) * 16;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + 8192) = lo;
  }
  fence_async_all();
}

// ---- streaming state of the loader / MMA roles -----------------------------------------------
struct Ring {
  int slot; uint32_t phase;     // P2: one phase bit for the whole ring; P3: bit s = parity of the uses of slot s
  __device__ __forceinline__ void next(int nslot) { if (++slot == nslot) { slot = 0; phase ^= 1; } }
};

// loader: one 16 KB weight chunk into the next ring slot
__device__ __forceinline__ void load_chunk(Ring& rg, const TcPlan& pl, Bars* bars, unsigned char* ring, const unsigned char* src, int* gerr) {
  TCW(30, tc_wait(&bars->ring_empty[rg.slot], rg.phase ^ 1, bars, gerr));
  mbar_expect_tx(&bars->ring_full[rg.slot], CHUNK);
  bulk_g2s(ring + rg.slot * CHUNK, src, CHUNK, &bars->ring_full[rg.slot]);
  rg.next(pl.nslot);
}

// MMA issuer: consume the next ring slot (A = 128 x 32 chunk, K-major) against a B operand whose K groups of 8 are
// `b_kg` bytes apart (b_lo = byte offset of the lo plane), `ksteps` (1 or 2) MMA K steps, accumulate into tmem_d
__device__ __forceinline__ void mma_chunk(Ring& rg, const TcPlan& pl, Bars* bars, unsigned char* ring, int* gerr, uint32_t tmem_d,
                                          uint32_t b_addr, uint32_t b_lo, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                                          uint32_t id, int ksteps, uint32_t& accum) {
  TCW(31, tc_wait(&bars->ring_full[rg.slot], rg.phase, bars, gerr));
  tc_fence_after();
  const uint32_t a = smem_u32(ring + rg.slot * CHUNK);
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t ah = smem_desc(a + ks * 4096, 2048, 128), al = smem_desc(a + 8192 + ks * 4096, 2048, 128);
    const uint64_t bh = smem_desc(b_addr + ks * b_kstep, b_lbo, b_sbo), bl = smem_desc(b_addr + b_lo + ks * b_kstep, b_lbo, b_sbo);
    mma_f16(tmem_d, ah, bh, id, accum); accum = 1;
    mma_f16(tmem_d, ah, bl, id, 1);
    mma_f16(tmem_d, al, bh, id, 1);
  }
  mma_commit(&bars->ring_empty[rg.slot]);
  rg.next(pl.nslot);
}

__device__ __forceinline__ int ksteps_of(int k16, int c) { return min(2, (k16 - 32 * c) / 16); }

// -------------------------------------------------------------------------------------------
// P2 (+P1): one tile of R rows
// -------------------------------------------------------------------------------------------
struct TileCtx {
  const ModelView* mv; const StepCtx* cx; const mopoe_batch_desc* b; const Workspace* ws; const TcPlan* pl;
  unsigned char* sm; Bars* bars; int* gerr;
  int tile, r0, nr; int64_t eps_base; bool bwd;
};

template <int R>
__device__ void tile_loader(const TileCtx& c, Ring& rg) {
  const ModelView& mv = *c.mv; const TcPlan& pl = *c.pl;
  unsigned char* ring = c.sm + pl.s_ring;
  const int present = c.b->present_mask;
  for (int m = 0; m < mv.M; ++m) {
    if (!(present >> m & 1)) continue;
    const TcMod& t = pl.mod[m];
    for (int cc = 0; cc < t.NCx; ++cc)
      for (int mt = 0; mt < 2; ++mt) load_chunk(rg, pl, c.bars, ring, pl.base + t.w1 + (int64_t)(mt * t.NCx + cc) * CHUNK, c.gerr);
    for (int cc = 0; cc < 8; ++cc) load_chunk(rg, pl, c.bars, ring, pl.base + t.wh + (int64_t)cc * CHUNK, c.gerr);
  }
  for (int m = 0; m < mv.M; ++m) {
    if (!(present >> m & 1)) continue;
    const TcMod& t = pl.mod[m];
    for (int p = 0; p < pl.np; ++p) {
      for (int cc = 0; cc < t.NCz; ++cc) load_chunk(rg, pl, c.bars, ring, pl.base + t.wd + (int64_t)cc * CHUNK, c.gerr);
      for (int mt = 0; mt < t.MtD; ++mt) {
        if (c.bwd)
          for (int c4 = 0; c4 < 4 && 4 * mt + c4 < t.NCx; ++c4) load_chunk(rg, pl, c.bars, ring, pl.base + t.wdt + (int64_t)(4 * mt + c4) * CHUNK, c.gerr);
        if (mt + 1 < t.MtD)
          for (int cc = 0; cc < t.NCz; ++cc) load_chunk(rg, pl, c.bars, ring, pl.base + t.wd + (int64_t)((mt + 1) * t.NCz + cc) * CHUNK, c.gerr);
      }
    }
  }
  if (c.bwd)
    for (int m = 0; m < mv.M; ++m) {
      if (!(present >> m & 1)) continue;
      const TcMod& t = pl.mod[m];
      for (int mt = 0; mt < 2; ++mt)
        for (int cc = 0; cc < t.NCh; ++cc) load_chunk(rg, pl, c.bars, ring, pl.base + t.wht + (int64_t)(mt * t.NCh + cc) * CHUNK, c.gerr);
    }
}

// phase counters of the role hand-overs (kept in registers by each role, advanced identically)
struct Sync { uint32_t nb, na, nx, nf; };

template <int R>
__device__ void tile_mma(const TileCtx& c, Ring& rg, Sync& sy, uint32_t tmem) {
  const ModelView& mv = *c.mv; const TcPlan& pl = *c.pl;
  unsigned char* ring = c.sm + pl.s_ring;
  Bars* bars = c.bars;
  const int present = c.b->present_mask;
  const uint32_t SF = R * 16;                         // bytes between feature groups of 8 in a tile operand
  const uint32_t u = smem_u32(c.sm + pl.s_u);
  const uint32_t hbuf = u + 256u * R, xbuf[2] = {hbuf, hbuf + 512u * R}, zop = u, dxbuf = u + 256u * R, deop = u;
  const uint32_t id_k = idesc(R, 0, 0), id_mn = idesc(R, 0, 1);
  const uint32_t acc0 = tmem, acc1 = tmem + R, acc2 = tmem + 2 * R;
  for (int m = 0; m < mv.M; ++m) {
    if (!(present >> m & 1)) continue;
    const TcMod& t = pl.mod[m];
    // P1: h^T = W1 x^T; the compute warps convert x in blocks of 128 features (layout F: K-major B), double
    // buffered inside the (not yet used) h operand buffer
    uint32_t a0 = 0, a1 = 0;
    for (int xi = 0; 4 * xi < t.NCx; ++xi) {
      const int xb = sy.nx & 1;
      TCW(32, tc_wait(&bars->xfull[xb], (sy.nx >> 1) & 1, bars, c.gerr));
      tc_fence_after();
      for (int c4 = 0; c4 < 4 && 4 * xi + c4 < t.NCx; ++c4) {
        const int ks = ksteps_of(t.Dk16, 4 * xi + c4);
        mma_chunk(rg, pl, bars, ring, c.gerr, acc0, xbuf[xb] + c4 * 4 * SF, 256u * R, SF, 128, 2 * SF, id_k, ks, a0);
        mma_chunk(rg, pl, bars, ring, c.gerr, acc1, xbuf[xb] + c4 * 4 * SF, 256u * R, SF, 128, 2 * SF, id_k, ks, a1);
      }
      mma_commit(&bars->xempty[xb]);
      ++sy.nx;
    }
    mma_commit(&bars->acc_done); ++sy.na;
    // S1: heads^T = Wh h^T (B = hbuf, layout T: MN-major, K groups SF apart)
    TCW(33, tc_wait(&bars->b_ready, sy.nb & 1, bars, c.gerr)); ++sy.nb;
    tc_fence_after();
    uint32_t a2 = 0;
    for (int cc = 0; cc < 8; ++cc)
      mma_chunk(rg, pl, bars, ring, c.gerr, acc2, hbuf + cc * 4 * SF, 512u * R, SF, 128, 2 * SF, id_mn, 2, a2);
    mma_commit(&bars->acc_done); ++sy.na;
  }
  for (int m = 0; m < mv.M; ++m) {
    if (!(present >> m & 1)) continue;
    const TcMod& t = pl.mod[m];
    for (int p = 0; p < pl.np; ++p) {
      TCW(33, tc_wait(&bars->b_ready, sy.nb & 1, bars, c.gerr)); ++sy.nb;     // zop ready
      tc_fence_after();
      {
        uint32_t a = 0;
        for (int cc = 0; cc < t.NCz; ++cc)
          mma_chunk(rg, pl, bars, ring, c.gerr, acc0, zop + cc * 4 * SF, 128u * R, SF, 128, 2 * SF, id_mn, ksteps_of(t.ZDk16, cc), a);
        mma_commit(&bars->acc_done); ++sy.na;
      }
      uint32_t adz = 0;
      for (int mt = 0; mt < t.MtD; ++mt) {
        if (!c.bwd && mt + 1 >= t.MtD) break;
        TCW(33, tc_wait(&bars->b_ready, sy.nb & 1, bars, c.gerr)); ++sy.nb;   // epilogue of decoder tile mt done (dxbuf ready)
        tc_fence_after();
        if (c.bwd)
          for (int c4 = 0; c4 < 4 && 4 * mt + c4 < t.NCx; ++c4)
            mma_chunk(rg, pl, bars, ring, c.gerr, acc2, dxbuf + c4 * 4 * SF, 256u * R, SF, 128, 2 * SF, id_mn, ksteps_of(t.Dk16, 4 * mt + c4), adz);
        if (mt + 1 < t.MtD) {
          uint32_t a = 0;
          for (int cc = 0; cc < t.NCz; ++cc)
            mma_chunk(rg, pl, bars, ring, c.gerr, ((mt + 1) & 1) ? acc1 : acc0, zop + cc * 4 * SF, 128u * R, SF, 128, 2 * SF, id_mn, ksteps_of(t.ZDk16, cc), a);
        }
        mma_commit(&bars->acc_done); ++sy.na;
      }
    }
  }
  if (c.bwd)
    for (int m = 0; m < mv.M; ++m) {
      if (!(present >> m & 1)) continue;
      const TcMod& t = pl.mod[m];
      TCW(33, tc_wait(&bars->b_ready, sy.nb & 1, bars, c.gerr)); ++sy.nb;     // deop ready
      tc_fence_after();
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t a = 0;
        for (int cc = 0; cc < t.NCh; ++cc)
          mma_chunk(rg, pl, bars, ring, c.gerr, mt ? acc1 : acc0, deop + cc * 4 * SF, 256u * R, SF, 128, 2 * SF, id_mn, ksteps_of(t.HCk16, cc), a);
      }
      mma_commit(&bars->acc_done); ++sy.na;
    }
}

// compute warps: hand a finished B operand to the MMA warp
__device__ __forceinline__ void publish_b(Bars* bars) {
  tc_fence_before();
  fence_proxy_async();
  bar_compute();
  if (threadIdx.x == 0) mbar_arrive(&bars->b_ready);
}
__device__ __forceinline__ void await_acc(Bars* bars, Sync& sy, int* gerr) {
  tc_wait(&bars->acc_done, sy.na & 1, bars, gerr); ++sy.na;
  tc_fence_after();
}

template <int R>
__device__ void tile_compute(const TileCtx& c, Sync& sy, uint32_t tmem) {
  const ModelView& mv = *c.mv; const StepCtx& cx = *c.cx; const mopoe_batch_desc& b = *c.b; const TcPlan& pl = *c.pl;
  Bars* bars = c.bars;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, q = warp & 3, hf = warp >> 2;
  const int M = mv.M, L = mv.L, present = b.present_mask, N = b.n_rows, nr = c.nr, r0 = c.r0;
  const bool bwd = c.bwd;
  constexpr uint32_t SF = R * 16;
  constexpr int RH = R / 2;                                // rows per column half
  unsigned char* U = c.sm + pl.s_u;
  unsigned char* hbuf = U + 256 * R;
  unsigned char* xbuf[2] = {hbuf, hbuf + 512 * R};
  unsigned char* zop = U;
  unsigned char* dxbuf = U + 256 * R;
  unsigned char* deop = U;
  float* sh_red = reinterpret_cast<float*>(c.sm + pl.s_red);
  long long* srow = reinterpret_cast<long long*>(c.sm + pl.s_srow);
  uint32_t* mask = reinterpret_cast<uint32_t*>(c.sm + pl.s_mask);
  LatSh sh;
  sh.e = reinterpret_cast<float*>(c.sm + pl.s_e); sh.de = sh.e;
  sh.zz = reinterpret_cast<float*>(c.sm + pl.s_zz); sh.dzz = reinterpret_cast<float*>(c.sm + pl.s_dzz);
  sh.rp = reinterpret_cast<float*>(c.sm + pl.s_rp); sh.rps = reinterpret_cast<float*>(c.sm + pl.s_rps);
  sh.red = sh_red; sh.R = R; sh.HCM = pl.hcm; sh.ZDM = pl.zdm; sh.SM_ = pl.sm_; sh.NP = pl.np;
  const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
  const float invN = 1.f / (float)N, fN = (float)N;
  float* colpart = reinterpret_cast<float*>(pl.base + pl.colpart) + (int64_t)(2 * c.tile) * pl.ccols;

#ifdef TC_PROF
  float* prof = reinterpret_cast<float*>(pl.base + pl.err) + 64;
#endif
  TCP_DECL;
  bar_compute();                                            // previous tile of this CTA is finished with smem
  if (t < MOPOE_N_SCALARS) sh_red[t] = 0.f;
  for (int i = t; i < M * R; i += 256) {
    const int m = i / R, n = i - m * R;
    srow[i] = (n < nr && (present >> m & 1)) ? src_row(cx, b, m, r0 + n) : -1;
  }
  bar_compute();
  TCP(0);
  // ================= P1 + S1 per modality =================
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const ModView& md = mv.mod[m];
    const TcMod& tm = pl.mod[m];
    const int D = md.D;
    unsigned char* g_x = pl.base + tm.xop + (int64_t)c.tile * tm.xop_t;
    // x rows -> fp16 hi/lo operand blocks of 128 features (layout F).  Software pipelined: the global loads of
    // block xi + 1 are in flight while block xi is converted, handed over and multiplied.
    constexpr int NI = (R * 16) / 256;                      // (row, 8-feature group) items per thread and block
    float xr[NI][8];
    auto x_load = [&](int xi) {
#pragma unroll
      for (int it = 0; it < NI; ++it) {
        const int i = t + 256 * it, n = i >> 4, g = i & 15;
        const long long sr = srow[m * R + n];
        const int d0 = xi * 128 + g * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) xr[it][k] = (sr >= 0 && d0 + k < D) ? __ldg(cx.x[m] + sr * D + d0 + k) : 0.f;
      }
    };
    x_load(0);
    for (int xi = 0; 4 * xi < tm.NCx; ++xi) {
      const int xb = sy.nx & 1;
      uint4 hi[NI], lo[NI];
#pragma unroll
      for (int it = 0; it < NI; ++it) split8(xr[it], hi[it], lo[it]);
      if (4 * (xi + 1) < tm.NCx) x_load(xi + 1);
      tc_wait(&bars->xempty[xb], ((sy.nx >> 1) & 1) ^ 1, bars, c.gerr);
#pragma unroll
      for (int it = 0; it < NI; ++it) {
        const int i = t + 256 * it, n = i >> 4, g = i & 15;
        const uint32_t off = g * SF + (n >> 3) * 128 + (n & 7) * 16;
        *reinterpret_cast<uint4*>(xbuf[xb] + off) = hi[it];
        *reinterpret_cast<uint4*>(xbuf[xb] + 256 * R + off) = lo[it];
        if (bwd && xi * 16 + g < tm.NCx * 4) {
          const uint32_t goff = (xi * 16 + g) * SF + (n >> 3) * 128 + (n & 7) * 16;
          *reinterpret_cast<uint4*>(g_x + goff) = hi[it];
          *reinterpret_cast<uint4*>(g_x + tm.xop_t / 2 + goff) = lo[it];
        }
      }
      fence_proxy_async();
      bar_compute();
      if (t == 0) mbar_arrive(&bars->xfull[xb]);
      ++sy.nx;
    }
    TCP(1);
    // ---- P1 epilogue: h = relu(acc + b1) -> hbuf (layout T), HBM copy for dWheads, relu mask ----
    await_acc(bars, sy, c.gerr);
    TCP(2);
    {
      const int mt = hf, j = 128 * mt + 32 * q + lane;      // warp = (M tile, lane quarter), all R rows
      float v[R];
      tmem_ld_cols<R>(lane_base + mt * R, v);
      tmem_ld_wait();
      const float bj = md.b1[j];
      uint32_t mk = 0;
      unsigned char* g_h = pl.base + tm.hop + (int64_t)c.tile * (256 * R * 4);
#pragma unroll
      for (int g8 = 0; g8 < R / 8; ++g8) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = g8 * 8 + i;
          const float hv = n < nr ? fmaxf(v[n] + bj, 0.f) : 0.f;
          x[i] = hv;
          mk |= (hv > 0.f ? 1u : 0u) << n;
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = (j >> 3) * SF + g8 * 128 + (j & 7) * 16;
        *reinterpret_cast<uint4*>(hbuf + off) = hi;
        *reinterpret_cast<uint4*>(hbuf + 512 * R + off) = lo;
        if (bwd) {
          *reinterpret_cast<uint4*>(g_h + off) = hi;
          *reinterpret_cast<uint4*>(g_h + 512 * R + off) = lo;
        }
      }
      mask[m * 256 + j] = mk;
    }
    publish_b(bars);
    TCP(3);
    // ---- S1 epilogue: heads -> sh.e[m][row][j] ----
    await_acc(bars, sy, c.gerr);
    TCP(4);
    {
      const int j = 32 * q + lane;
      float v[RH];
      tmem_ld_cols<RH>(lane_base + 2 * R + hf * RH, v);
      tmem_ld_wait();
      if (j < md.HC) {
        const float bj = md.bh[j];
#pragma unroll
        for (int i = 0; i < RH; ++i) {
          const int n = hf * RH + i;
          const float ev = v[i] + bj;
          sh.e[(m * R + n) * pl.hcm + j] = ev;
          if (cx.out.enc_heads[m] && n < nr) cx.out.enc_heads[m][(int64_t)(r0 + n) * md.HC + j] = ev;
        }
      }
    }
    tc_fence_before();
    bar_compute();
    TCP(5);
  }
  // ================= latent forward =================
  for (int i = t; i < M * pl.np * R * pl.zdm; i += 256) sh.dzz[i] = 0.f;
  lat_forward(mv, cx, b, c.eps_base, r0, nr, sh);
  bar_compute();
  TCP(6);
  // ================= decoders (+ NLL, d x_hat, d z) =================
  for (int m = 0; m < M; ++m) {
    if (!(present >> m & 1)) continue;
    const ModView& md = mv.mod[m];
    const TcMod& tm = pl.mod[m];
    const int D = md.D, ZD = md.ZD;
    for (int p = 0; p < pl.np; ++p) {
      // decoder input -> operand (layout T, 64 feature slots, zero padded; invalid rows zero)
      unsigned char* g_z = pl.base + tm.zzop + ((int64_t)c.tile * pl.np + p) * (64 * R * 4);
      for (int i = t; i < 64 * (R / 8); i += 256) {
        const int z = i / (R / 8), g8 = i - z * (R / 8);
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int n = g8 * 8 + k;
          x[k] = (z < ZD && n < nr) ? sh.zz[((m * pl.np + p) * R + n) * pl.zdm + z] : 0.f;
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = (z >> 3) * SF + g8 * 128 + (z & 7) * 16;
        *reinterpret_cast<uint4*>(zop + off) = hi;
        *reinterpret_cast<uint4*>(zop + 128 * R + off) = lo;
        if (bwd) {
          *reinterpret_cast<uint4*>(g_z + off) = hi;
          *reinterpret_cast<uint4*>(g_z + 128 * R + off) = lo;
        }
      }
      publish_b(bars);
      TCP(7);
      unsigned char* g_dx = pl.base + tm.dxop + ((int64_t)c.tile * pl.np + p) * tm.dxop_t;
      float nll = 0.f;
      for (int mt = 0; mt < tm.MtD; ++mt) {
        await_acc(bars, sy, c.gerr);
        TCP(8);
        const int f = 32 * q + lane, d = 128 * mt + f;
        float v[RH];
        tmem_ld_cols<RH>(lane_base + (mt & 1) * R + hf * RH, v);
        tmem_ld_wait();
        float gs = 0.f, ls = 0.f;
        float g[RH];
        if (d < D) {
          const float bd = md.bd[d], lam = md.lv[d], iv = expf(-lam);
          float xv[RH];
#pragma unroll
          for (int i = 0; i < RH; ++i) {
            const long long sr = srow[m * R + hf * RH + i];
            xv[i] = (sr >= 0 && cx.with_nll) ? cx.x[m][sr * D + d] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < RH; ++i) {
            const int n = hf * RH + i;
            const float loc = v[i] + bd;
            if (n < nr) {
              if (p == 0 && cx.out.rec_loc[m]) cx.out.rec_loc[m][(int64_t)(r0 + n) * D + d] = loc;
              const float diff = xv[i] - loc;
              const float w = diff * diff * iv;
              if (cx.with_nll) nll += 0.5f * w + 0.5f * lam + HALF_LOG_2PI;
              g[i] = cx.with_nll ? -diff * iv : 0.f;       // N x (d loss / d x_hat)
              gs += g[i]; ls += 0.5f * (1.f - w);
            } else g[i] = 0.f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < RH; ++i) g[i] = 0.f;
        }
        if (bwd) {
#pragma unroll
          for (int g8 = 0; g8 < RH / 8; ++g8) {
            uint4 hi, lo;
            split8(g + 8 * g8, hi, lo);
            const uint32_t rowg = hf * (RH / 8) + g8;
            const uint32_t off = (f >> 3) * SF + rowg * 128 + (f & 7) * 16;
            *reinterpret_cast<uint4*>(dxbuf + off) = hi;
            *reinterpret_cast<uint4*>(dxbuf + 256 * R + off) = lo;
            const uint32_t goff = (d >> 3) * SF + rowg * 128 + (d & 7) * 16;
            *reinterpret_cast<uint4*>(g_dx + goff) = hi;
            *reinterpret_cast<uint4*>(g_dx + tm.dxop_t / 2 + goff) = lo;
          }
          float* cp = colpart + (int64_t)hf * pl.ccols + tm.col0 + 384 + d;
          if (p == 0) { cp[0] = gs; cp[tm.MtD * 128] = ls; }
          else { cp[0] += gs; cp[tm.MtD * 128] += ls; }
        }
        if (bwd || mt + 1 < tm.MtD) publish_b(bars);
        else { tc_fence_before(); bar_compute(); }
        TCP(9);
      }
      if (cx.with_nll) block_add(sh_red, (p == 0 ? MOPOE_S_NLL : MOPOE_S_NLL_UNI) + m, nll);
      if (bwd) {        // d z of this (modality, pass): 1/N applied here, fp32 from now on
        await_acc(bars, sy, c.gerr);
        TCP(10);
        const int z = 32 * q + lane;
        if (q < 2) {
          float v[RH];
          tmem_ld_cols<RH>(lane_base + 2 * R + hf * RH, v);
          tmem_ld_wait();
          if (z < ZD) {
#pragma unroll
            for (int i = 0; i < RH; ++i) sh.dzz[((m * pl.np + p) * R + hf * RH + i) * pl.zdm + z] = v[i] * invN;
          }
        }
        tc_fence_before();
        bar_compute();
        TCP(11);
      }
    }
  }
  if (bwd) {
    // ================= latent backward (d heads overwrite the heads in place) =================
    lat_backward(mv, cx, b, r0, nr, sh);
    bar_compute();
    TCP(12);
    for (int m = 0; m < M; ++m) {
      if (!(present >> m & 1)) continue;
      const ModView& md = mv.mod[m];
      const TcMod& tm = pl.mod[m];
      const int HC = md.HC;
      // d heads x N -> operand (128 feature slots) + HBM copy + column sums (d bias of the heads)
      unsigned char* g_de = pl.base + tm.deop + (int64_t)c.tile * (128 * R * 4);
      for (int i = t; i < 128 * (R / 8); i += 256) {
        const int j = i / (R / 8), g8 = i - j * (R / 8);
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int n = g8 * 8 + k;
          x[k] = (j < HC && n < nr) ? sh.e[(m * R + n) * pl.hcm + j] * fN : 0.f;
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = (j >> 3) * SF + g8 * 128 + (j & 7) * 16;
        *reinterpret_cast<uint4*>(deop + off) = hi;
        *reinterpret_cast<uint4*>(deop + 256 * R + off) = lo;
        *reinterpret_cast<uint4*>(g_de + off) = hi;
        *reinterpret_cast<uint4*>(g_de + 256 * R + off) = lo;
      }
      if (t < 128) {
        float s = 0.f;
        if (t < HC)
          for (int n = 0; n < nr; ++n) s += sh.e[(m * R + n) * pl.hcm + t];
        colpart[tm.col0 + 256 + t] = s * fN;
        colpart[pl.ccols + tm.col0 + 256 + t] = 0.f;
      }
      publish_b(bars);
      TCP(13);
      // ---- S4 epilogue: d pre-activation x N = (Wh^T d heads) * relu' -> HBM operand of dW1, column sums (d b1) ----
      await_acc(bars, sy, c.gerr);
      TCP(14);
      {
        const int mt = hf, j = 128 * mt + 32 * q + lane;
        float v[R];
        tmem_ld_cols<R>(lane_base + mt * R, v);
        tmem_ld_wait();
        const uint32_t mk = mask[m * 256 + j];
        unsigned char* g_da = pl.base + tm.daop + (int64_t)c.tile * (256 * R * 4);
        float s = 0.f;
#pragma unroll
        for (int g8 = 0; g8 < R / 8; ++g8) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = g8 * 8 + i;
            x[i] = (mk >> n & 1) ? v[n] : 0.f;
            s += x[i];
          }
          uint4 hi, lo;
          split8(x, hi, lo);
          const uint32_t off = (j >> 3) * SF + g8 * 128 + (j & 7) * 16;
          *reinterpret_cast<uint4*>(g_da + off) = hi;
          *reinterpret_cast<uint4*>(g_da + 512 * R + off) = lo;
        }
        colpart[tm.col0 + j] = s;
        colpart[pl.ccols + tm.col0 + j] = 0.f;
      }
      tc_fence_before();
      bar_compute();
      TCP(15);
    }
  }
  bar_compute();
  if (t < MOPOE_N_SCALARS && sh_red[t] != 0.f) atomicAdd(c.ws->acc + t, (double)sh_red[t]);
}

// -------------------------------------------------------------------------------------------
// P3: weight gradients (+ Adam)
// -------------------------------------------------------------------------------------------
struct P3Geom {
  int m, g, mt, nb;          // unit
  int nw, n0;                // N width of this unit, first B feature
  int rows_valid, cols_valid, ld;
  int64_t pbase;             // parameter index of element (0, 0) of the full matrix
  const unsigned char* a; int a_t, a_half, a_off;   // A blob: tile 0, bytes per (tile[,pass]), lo-plane offset, slice offset
  const unsigned char* bq; int b_t, b_half, b_off;  // B blob likewise
  int b_mn, np;
};

template <int R>
__device__ __forceinline__ P3Geom p3_geom(const ModelView& mv, const StepCtx& cx, const TcPlan& pl, const TcUnit& u) {
  P3Geom g;
  const ModView& md = mv.mod[u.m];
  const TcMod& t = pl.mod[u.m];
  constexpr int SF = R * 16;
  g.m = u.m; g.g = u.g; g.mt = u.mt; g.nb = u.nb; g.np = 1; g.b_mn = 0;
  if (u.g == 0) {            // dW1[j][d] = sum_n dA[n][j] x[n][d]
    const int nb0 = cdiv(t.Dk16, pl.nw[0]);
    const int nw0 = cdiv(cdiv(t.Dk16, nb0), 16) * 16;
    g.n0 = u.nb * nw0; g.nw = min(nw0, t.Dk16 - g.n0);
    g.rows_valid = MOPOE_HIDDEN; g.cols_valid = md.D; g.ld = md.D; g.pbase = cx.lay.enc_w1[u.m];
    g.a = pl.base + t.daop; g.a_t = 256 * R * 4; g.a_half = 512 * R; g.a_off = u.mt * 16 * SF;
    g.bq = pl.base + t.xop; g.b_t = t.xop_t; g.b_half = t.xop_t / 2; g.b_off = (g.n0 / 8) * SF; g.b_mn = 1;
  } else if (u.g == 1) {     // dWh[j][k] = sum_n de[n][j] h[n][k]
    g.nw = pl.nw[1]; g.n0 = u.nb * g.nw;
    g.rows_valid = md.HC; g.cols_valid = MOPOE_HIDDEN; g.ld = MOPOE_HIDDEN; g.pbase = cx.lay.enc_wh[u.m];
    g.a = pl.base + t.deop; g.a_t = 128 * R * 4; g.a_half = 256 * R; g.a_off = 0;
    g.bq = pl.base + t.hop; g.b_t = 256 * R * 4; g.b_half = 512 * R; g.b_off = (g.n0 / 8) * SF;
  } else {                   // dWd[d][z] = sum_{p,n} dx[p][n][d] zz[p][n][z]
    g.nw = t.ZDk16; g.n0 = 0; g.np = pl.np;
    g.rows_valid = md.D; g.cols_valid = md.ZD; g.ld = md.ZD; g.pbase = cx.lay.dec_w[u.m];
    g.a = pl.base + t.dxop; g.a_t = t.dxop_t; g.a_half = t.dxop_t / 2; g.a_off = u.mt * 16 * SF;
    g.bq = pl.base + t.zzop; g.b_t = 64 * R * 4; g.b_half = 128 * R; g.b_off = 0;
  }
  return g;
}

// one (unit, row-range split): roles as in P2.  tiles [t_lo, t_hi) x passes are the K chunks
template <int R>
__device__ void p3_item(const ModelView& mv, const StepCtx& cx, const mopoe_batch_desc& b, const TcPlan& pl, unsigned char* sm,
                        Bars* bars, int* gerr, Ring& rg, Sync& sy, uint32_t tmem, int ui, int split, int nt) {
  const P3Geom g = p3_geom<R>(mv, cx, pl, pl.unit[ui]);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  unsigned char* ring = sm + pl.s_ring;
  // ring of this item: as many slots as fit (the slot holds one row tile of both operands); every item starts at
  // slot 0, the parity of each slot's uses is carried in the bits of rg.phase across items of different slot size
  const uint32_t a_bytes = 128 * R * 2, b_bytes = g.nw * R * 2;
  const int slot_bytes = 2 * (a_bytes + b_bytes);
  const int nslot3 = min(MAX_SLOTS, (pl.nslot * CHUNK) / slot_bytes);
  rg.slot = 0;
  const int t_lo = (int)((int64_t)split * nt / pl.ksplit), t_hi = (int)((int64_t)(split + 1) * nt / pl.ksplit);
  const int nchunk = (t_hi - t_lo) * g.np;
  constexpr uint32_t SF = R * 16;
  if (warp == 8) {
    if (lane == 0)
      for (int ci = 0; ci < nchunk; ++ci) {
        const int64_t tp = (int64_t)t_lo * g.np + ci;            // (tile, pass) index
        TCW(34, tc_wait(&bars->ring3_empty[rg.slot], ((rg.phase >> rg.slot) & 1) ^ 1, bars, gerr));
        mbar_expect_tx(&bars->ring3_full[rg.slot], 2 * a_bytes + 2 * b_bytes);
        unsigned char* dst = ring + rg.slot * slot_bytes;
        const unsigned char* as = g.a + tp * g.a_t + g.a_off;
        const unsigned char* bs = g.bq + tp * g.b_t + g.b_off;
        bulk_g2s(dst, as, a_bytes, &bars->ring3_full[rg.slot]);
        bulk_g2s(dst + a_bytes, as + g.a_half, a_bytes, &bars->ring3_full[rg.slot]);
        bulk_g2s(dst + 2 * a_bytes, bs, b_bytes, &bars->ring3_full[rg.slot]);
        bulk_g2s(dst + 2 * a_bytes + b_bytes, bs + g.b_half, b_bytes, &bars->ring3_full[rg.slot]);
        rg.phase ^= 1u << rg.slot;
        if (++rg.slot == nslot3) rg.slot = 0;
      }
  } else if (warp == 9) {
    if (lane == 0) {
      const uint32_t id = idesc(g.nw, 0, g.b_mn);
      uint32_t accum = 0;
      // the compute warps must have read the previous item's accumulator out of TMEM
      if (sy.nf > 0) { tc_wait(&bars->acc_free, (sy.nf - 1) & 1, bars, gerr); tc_fence_after(); }
      ++sy.nf;
      for (int ci = 0; ci < nchunk; ++ci) {
        TCW(35, tc_wait(&bars->ring3_full[rg.slot], (rg.phase >> rg.slot) & 1, bars, gerr));
        tc_fence_after();
        const uint32_t a = smem_u32(ring + rg.slot * slot_bytes);
        const uint32_t bb = a + 2 * a_bytes;
        for (int ks = 0; ks < R / 16; ++ks) {
          // K = rows: K groups of 8 rows are 128 bytes apart, MN groups of 8 features SF apart (both layouts)
          const uint64_t ah = smem_desc(a + ks * 256, 128, SF), al = smem_desc(a + a_bytes + ks * 256, 128, SF);
          const uint64_t bh = smem_desc(bb + ks * 256, 128, SF), bl = smem_desc(bb + b_bytes + ks * 256, 128, SF);
          mma_f16(tmem, ah, bh, id, accum); accum = 1;
          mma_f16(tmem, ah, bl, id, 1);
          mma_f16(tmem, al, bh, id, 1);
        }
        mma_commit(&bars->ring3_empty[rg.slot]);
        rg.phase ^= 1u << rg.slot;
        if (++rg.slot == nslot3) rg.slot = 0;
      }
      if (nchunk > 0) { mma_commit(&bars->acc_done); ++sy.na; }
    }
  } else {
    const int q = warp & 3, hf = warp >> 2;
    const float invN = 1.f / (float)b.n_rows;
    float bc1 = 1.f, bc2s = 1.f;
    if (cx.mode == 2) {
      const float tt = (float)(cx.adam_t[g.m] + 1);
      bc1 = 1.f - powf(cx.b1, tt);
      bc2s = sqrtf(1.f - powf(cx.b2, tt));
    }
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
#ifdef TC_PROF
    float* prof = reinterpret_cast<float*>(pl.base + pl.err) + 64;
#endif
    TCP_DECL;
    if (nchunk > 0) await_acc(bars, sy, gerr);
    TCP(36);
    float* part = reinterpret_cast<float*>(pl.base + pl.p3part) + ((int64_t)ui * pl.ksplit) * (128 * 256);
    int* cnt = reinterpret_cast<int*>(pl.base + pl.p3cnt) + ui;
    float* stile = reinterpret_cast<float*>(sm + pl.s_e);          // [128][CW + 1] staging tile (the P2 buffers are idle)
    __shared__ int s_last;
    bool reduce = false;
    if (pl.ksplit > 1) {
      // partial tile, column major ([col][lane row]: coalesced), then the last CTA to arrive sums in split order
      float* mine = part + (int64_t)split * (128 * 256);
      const int cw = g.nw / 2, c0 = hf * cw;
      for (int cb = 0; cb < cw; cb += 8) {
        float v[8];
        if (nchunk > 0) { tmem_ld8(lane_base + c0 + cb, v); tmem_ld_wait(); }
        else { for (int k = 0; k < 8; ++k) v[k] = 0.f; }
        for (int k = 0; k < 8; ++k) mine[(c0 + cb + k) * 128 + 32 * q + lane] = v[k];
      }
      tc_fence_before();
      __threadfence();
      bar_compute();
      if (t == 0) {
        mbar_arrive(&bars->acc_free);
        const int prev = atomicAdd(cnt, 1);
        s_last = prev == pl.ksplit - 1;
        if (s_last) *cnt = 0;
      }
      bar_compute();
      reduce = s_last != 0;
      if (reduce) __threadfence();
    }
    if (pl.ksplit == 1 || reduce) {
      // 32 (or 16) output columns at a time: accumulator (or the sum of the partials) -> staging tile, then the
      // parameter update with consecutive threads on consecutive columns of a row and all loads in flight
      const int CW = (g.nw & 31) ? 16 : 32;
      for (int cbase = 0; cbase < g.nw; cbase += CW) {
        const int hw = CW / 2, cc0 = cbase + hf * hw;
        if (!reduce) {
          float v[16];
          if (CW == 32) tmem_ld16(lane_base + cc0, v); else tmem_ld8(lane_base + cc0, v);
          tmem_ld_wait();
          for (int k = 0; k < hw; ++k) stile[(32 * q + lane) * (CW + 1) + hf * hw + k] = v[k];
        } else {
          for (int k = 0; k < hw; ++k) {
            float sum = 0.f;
            for (int sp = 0; sp < pl.ksplit; ++sp) sum += __ldcg(part + (int64_t)sp * (128 * 256) + (cc0 + k) * 128 + 32 * q + lane);
            stile[(32 * q + lane) * (CW + 1) + hf * hw + k] = sum;
          }
        }
        bar_compute();
        // 4 elements per thread and pass (all their loads in flight), a compact loop instead of one long
        // unrolled sequence: the update code runs once per column block and would otherwise be fetched cold
        const int per = (128 * CW) / 256;                     // 16 or 8 elements per thread
#pragma unroll 1
        for (int k0 = 0; k0 < per; k0 += 4) {
          float gv[4], pm[4], pv[4], pp[4];
          int64_t ix[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = t + 256 * (k0 + k), il = e / CW, jj = e - il * CW;
            const int i = 128 * g.mt + il, j = g.n0 + cbase + jj;
            ix[k] = (i < g.rows_valid && j < g.cols_valid) ? g.pbase + (int64_t)i * g.ld + j : -1;
            gv[k] = stile[il * (CW + 1) + jj] * invN;
          }
          if (cx.mode == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (ix[k] >= 0) cx.grads[ix[k]] = gv[k];
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (ix[k] >= 0) { pm[k] = cx.adam_m[ix[k]]; pv[k] = cx.adam_v[ix[k]]; pp[k] = cx.params[ix[k]]; }
#pragma unroll
            for (int k = 0; k < 4; ++k) if (ix[k] >= 0) {
              const float m_ = cx.b1 * pm[k] + (1.f - cx.b1) * gv[k];
              const float v_ = cx.b2 * pv[k] + (1.f - cx.b2) * gv[k] * gv[k];
              cx.adam_m[ix[k]] = m_;
              cx.adam_v[ix[k]] = v_;
              cx.params[ix[k]] = pp[k] - (cx.lr / bc1) * (m_ / (sqrtf(v_) / bc2s + cx.adam_eps));
            }
          }
        }
        bar_compute();
      }
    }
    if (pl.ksplit == 1) {
      tc_fence_before();
      bar_compute();
      if (t == 0) mbar_arrive(&bars->acc_free);
    }
    TCP(37);
  }
}

// column sums (bias / output log-variance gradients): thread per column, fixed-order sum over the tile partials
__device__ void p3_columns(const ModelView& mv, const StepCtx& cx, const mopoe_batch_desc& b, const TcPlan& pl, int nt, int item) {
  const int t = threadIdx.x;
  if (t >= 256) return;
  const int col = item * 256 + t;
  if (col >= pl.ccols) return;
  int m = 0;
  for (int mm = 0; mm < mv.M; ++mm) if (col >= pl.mod[mm].col0) m = mm;
  if (!(b.present_mask >> m & 1)) return;
  const ModView& md = mv.mod[m];
  const TcMod& tm = pl.mod[m];
  const int c = col - tm.col0;
  int64_t idx; bool ok; bool is_lv = false;
  if (c < 256) { idx = cx.lay.enc_b1[m] + c; ok = true; }
  else if (c < 384) { idx = cx.lay.enc_bh[m] + (c - 256); ok = c - 256 < md.HC; }
  else if (c < 384 + tm.MtD * 128) { idx = cx.lay.dec_b[m] + (c - 384); ok = c - 384 < md.D; }
  else { idx = cx.lay.dec_lv[m] + (c - 384 - tm.MtD * 128); ok = c - 384 - tm.MtD * 128 < md.D; is_lv = true; }
  if (!ok) return;
  const float* cp = reinterpret_cast<const float*>(pl.base + pl.colpart) + col;
  float s = 0.f, tot = 0.f;
  for (int r = 0; r < 2 * nt; ++r) {
    s += __ldcg(cp + (int64_t)r * pl.ccols);
    if ((r & 63) == 63) { tot += s; s = 0.f; }
  }
  tot += s;
  float bc1 = 1.f, bc2s = 1.f;
  if (cx.mode == 2) {
    const float tt = (float)(cx.adam_t[m] + 1);
    bc1 = 1.f - powf(cx.b1, tt);
    bc2s = sqrtf(1.f - powf(cx.b2, tt));
  }
  const float gval = tot / (float)b.n_rows;
  if (!is_lv || mv.learn_scale) apply_grad(cx, idx, gval, bc1, bc2s);
  else if (cx.mode == 1) cx.grads[idx] = 0.f;
}

// -------------------------------------------------------------------------------------------
// the persistent kernel
// -------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(THREADS, 1) train_tc_kernel(ModelView mv, StepCtx cx, const mopoe_batch_desc* batches,
                                                              int n_steps, float* scalars, Workspace ws, TcPlan pl) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ mopoe_batch_desc sb;
  Bars* bars = reinterpret_cast<Bars*>(sm + pl.s_bar);
  int* gerr = reinterpret_cast<int*>(pl.base + pl.err);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int tmem_cols = 256;   // P2 uses 3 R columns, a P3 output tile up to 256
  if (t == 0) {
    for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(&bars->ring_full[s], 1); mbar_init(&bars->ring_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->xfull[s], 1); mbar_init(&bars->xempty[s], 1); }
    for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(&bars->ring3_full[s], 1); mbar_init(&bars->ring3_empty[s], 1); }
    mbar_init(&bars->b_ready, 1); mbar_init(&bars->acc_done, 1); mbar_init(&bars->acc_free, 1);
    bars->dead = 0;
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(&bars->tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;
  Ring rg = {0, 0};
  Sync sy = {0, 0, 0, 0};
  Ring rg3 = {0, 0};
  unsigned int target = 0;
  for (int step = 0; step < n_steps; ++step) {
    __syncthreads();
    if (t == 0) sb = batches[step];
    __syncthreads();
    const mopoe_batch_desc& b = sb;
    if (blockIdx.x == 0 && t < MOPOE_N_SCALARS) ws.acc[t] = 0.0;
#ifdef TC_PROF
    float* prof = reinterpret_cast<float*>(pl.base + pl.err) + 64;
#endif
    TCP_DECL;
    tc_prep(mv, pl, reinterpret_cast<PrepBlob*>(sm + pl.s_e));
    TCP(20);
    grid_barrier(ws.bar, target);
    fence_async_all();
    TCP(21);
    const int nt = (b.n_rows + R - 1) / R;
    const bool bwd = cx.mode != 0;
    // ---- P1 + P2 ----
    for (int tile = blockIdx.x; tile < nt; tile += gridDim.x) {
      TileCtx c;
      c.mv = &mv; c.cx = &cx; c.b = &b; c.ws = &ws; c.pl = &pl; c.sm = sm; c.bars = bars; c.gerr = gerr;
      c.tile = tile; c.r0 = tile * R; c.nr = min(R, b.n_rows - tile * R);
      c.eps_base = (int64_t)step * cx.eps_step_stride; c.bwd = bwd;
      if (warp == 8) { if (lane == 0) tile_loader<R>(c, rg); }
      else if (warp == 9) { if (lane == 0) tile_mma<R>(c, rg, sy, tmem); }
      else tile_compute<R>(c, sy, tmem);
    }
    fence_async_all();
    TCP(22);
    grid_barrier(ws.bar, target);
    fence_async_all();
    TCP(23);
    if (blockIdx.x == 0 && t == 0) finalize_scalars(mv, cx, b, ws.acc, scalars + (int64_t)step * MOPOE_N_SCALARS);
    if (bwd) {
      // ---- P3 ----
      int n_active = 0;
      for (int ui = 0; ui < pl.n_units; ++ui) n_active += (b.present_mask >> pl.unit[ui].m) & 1;
      const int n_items = n_active * pl.ksplit;
      const int n_col_items = (pl.ccols + 255) / 256;
      // dynamic queue (units are sorted by cost on the host, heaviest first): the CTA takes its first item by
      // index, every further one from a global counter; all three roles of the CTA follow the same item
      int* queue = reinterpret_cast<int*>(pl.base + pl.p3cnt) + MAX_UNITS;
      __shared__ int s_item;
      int item = blockIdx.x;
      while (item < n_items + n_col_items) {
        if (item < n_items) {
          const int au = item / pl.ksplit, split = item - au * pl.ksplit;
          int ui = 0, seen = 0;
          for (int k = 0; k < pl.n_units; ++k) {
            if ((b.present_mask >> pl.unit[k].m) & 1) { if (seen == au) { ui = k; break; } ++seen; }
          }
          p3_item<R>(mv, cx, b, pl, sm, bars, gerr, rg3, sy, tmem, ui, split, nt);
        } else {
          p3_columns(mv, cx, b, pl, nt, item - n_items);
        }
        __syncthreads();
        if (t == 0) s_item = gridDim.x + atomicAdd(queue, 1);
        __syncthreads();
        item = s_item;
      }
      TCP(24);
      grid_barrier(ws.bar, target);
      TCP(25);
      if (blockIdx.x == 0 && t == 0) *(reinterpret_cast<int*>(pl.base + pl.p3cnt) + MAX_UNITS) = 0;   // next read two barriers away
      if (cx.mode == 2 && blockIdx.x == 0 && t < mv.M && (b.present_mask >> t & 1)) cx.adam_t[t] += 1;
    }
  }
  __syncthreads();
  if (bars->dead && t == 0) scalars[0] = __int_as_float(0x7fc00000);
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace tc
