// Host-side plumbing of the C-ABI: errors, descriptor validation, parameter layout, subset table.
#include <stdarg.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "mopoe_common.cuh"

namespace mopoe {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return MOPOE_ECUDA;
}

int num_sms() {   // of the CURRENT device (cached per device ordinal)
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (sms[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = n;
  }
  return sms[dev];
}

int check_desc(const mopoe_model_desc* d) {
  if (!d) { set_error("desc is NULL"); return MOPOE_EINVAL; }
  if (d->n_mods < 1 || d->n_mods > MOPOE_MAX_MODS) {
    set_error("n_mods=%d not in 1..%d", d->n_mods, MOPOE_MAX_MODS); return MOPOE_EINVAL; }
  if (d->hidden != MOPOE_HIDDEN) { set_error("hidden=%d unsupported (reference hard-codes 256)", d->hidden); return MOPOE_EINVAL; }
  if (d->n_hidden_enc < 0 || d->n_hidden_enc > MOPOE_MAX_LAYERS) { set_error("num_hidden_layer_encoder=%d not in 0..%d", d->n_hidden_enc, MOPOE_MAX_LAYERS); return MOPOE_EINVAL; }
  if (d->n_hidden_dec < 0 || d->n_hidden_dec > MOPOE_MAX_LAYERS) { set_error("num_hidden_layer_decoder=%d not in 0..%d", d->n_hidden_dec, MOPOE_MAX_LAYERS); return MOPOE_EINVAL; }
  if (d->method < MOPOE_METHOD_POE || d->method > MOPOE_METHOD_JSD) {
    set_error("method=%d unsupported (poe, moe, joint_elbo, jsd)", d->method); return MOPOE_EINVAL; }
  if (d->likelihood < 0 || d->likelihood > 1) { set_error("likelihood=%d unsupported (0 normal, 1 laplace)", d->likelihood); return MOPOE_EINVAL; }
  if (d->scale_mode < 0 || d->scale_mode > 1) { set_error("scale_mode=%d invalid", d->scale_mode); return MOPOE_EINVAL; }
  if (d->latent_dim < 1 || d->latent_dim > 32) { set_error("latent_dim=%d not in 1..32", d->latent_dim); return MOPOE_EINVAL; }
  for (int m = 0; m < d->n_mods; ++m) {
    if (d->dims[m] < 1 || d->dims[m] > 8192) { set_error("dims[%d]=%d not in 1..8192", m, d->dims[m]); return MOPOE_EINVAL; }
    if (d->style_dims[m] < 0 || d->style_dims[m] > 32) { set_error("style_dims[%d]=%d not in 0..32", m, d->style_dims[m]); return MOPOE_EINVAL; }
    if (d->name_rank[m] < 0 || d->name_rank[m] >= d->n_mods) { set_error("name_rank[%d]=%d invalid", m, d->name_rank[m]); return MOPOE_EINVAL; }
  }
  return MOPOE_OK;
}

static int64_t align32(int64_t v) { return (v + 31) & ~(int64_t)31; }

void build_subsets(const mopoe_model_desc* d, SubsetTable* t) {
  // BaseExperiment.set_subsets (BaseExperiment.py:58-79): itertools.combinations over the modality
  // list for sizes 1..M; the members of a subset are fused in the order of their sorted NAMES.
  memset(t, 0, sizeof(*t));
  int M = d->n_mods, n = 0;
  for (int size = 1; size <= M; ++size) {
    std::vector<int> c(size);
    for (int i = 0; i < size; ++i) c[i] = i;
    while (true) {
      int mask = 0;
      std::vector<int> mem(c.begin(), c.end());
      for (int v : c) mask |= 1 << v;
      std::sort(mem.begin(), mem.end(), [&](int a, int b) { return d->name_rank[a] < d->name_rank[b]; });
      t->mask[n] = mask;
      t->n_members[n] = size;
      for (int i = 0; i < size; ++i) t->members[n][i] = mem[i];
      ++n;
      int i = size - 1;
      while (i >= 0 && c[i] == M - size + i) --i;
      if (i < 0) break;
      ++c[i];
      for (int j = i + 1; j < size; ++j) c[j] = c[j - 1] + 1;
    }
  }
  t->n_subsets = n;
}

void build_view(const mopoe_model_desc* d, const mopoe_param_layout* lay, float* base, ModelView* v) {
  memset(v, 0, sizeof(*v));
  v->M = d->n_mods;
  v->L = d->latent_dim;
  v->method = d->method;
  v->learn_scale = d->learn_output_scale;
  v->beta = d->beta;
  v->beta_style = d->beta_style;
  v->beta_content = d->beta_content;
  int off = d->latent_dim, poff = (d->latent_dim + 3) & ~3;
  for (int m = 0; m < d->n_mods; ++m) {
    ModView& mv = v->mod[m];
    mv.D = d->dims[m];
    mv.S = d->style_dims[m];
    mv.HC = 2 * d->latent_dim + 2 * mv.S;
    mv.ZD = mv.S + d->latent_dim;
    mv.eps_off = off;
    off += mv.S;
    mv.peps_off = poff;
    poff += (mv.S + 3) & ~3;
    // (weight pointers are used by the fused kernels only: default architecture, where every offset exists)
    mv.w1 = base + lay->enc_w1[m];
    mv.b1 = base + lay->enc_b1[m];
    mv.wh = base + lay->enc_wh[m];
    mv.bh = base + lay->enc_bh[m];
    mv.wd = base + lay->dec_w[m];
    mv.bd = base + lay->dec_b[m];
    mv.lv = base + lay->dec_lv[m];
  }
  v->E = off;
  v->EP = poff;
  build_subsets(d, &v->sub);
}

__global__ void philox_fill_kernel(uint64_t seed, uint64_t stream, int64_t start, int64_t n, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = philox_normal1(seed, stream, (uint64_t)(start + i));
}

}  // namespace mopoe

using namespace mopoe;

extern "C" {

const char* mopoe_last_error(void) { return g_err; }
int mopoe_version(void) { return 100; }

int mopoe_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int mopoe_param_layout_of(const mopoe_model_desc* d, mopoe_param_layout* out) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!out) { set_error("out is NULL"); return MOPOE_EINVAL; }
  memset(out, 0, sizeof(*out));
  int64_t off = 0;
  const int L = d->latent_dim, H = MOPOE_HIDDEN, He = d->n_hidden_enc, Hd = d->n_hidden_dec;
  auto take = [&](int64_t n) { const int64_t o = off; off = align32(off + n); return o; };
  for (int m = 0; m < d->n_mods; ++m) {
    const int D = d->dims[m], S = d->style_dims[m];
    out->enc_w1[m] = He >= 1 ? take((int64_t)H * D) : -1;
    out->enc_b1[m] = He >= 1 ? take(H) : -1;
    for (int l = 1; l < MOPOE_MAX_LAYERS; ++l) {
      out->enc_wx[m][l - 1] = l < He ? take((int64_t)H * H) : -1;
      out->enc_bx[m][l - 1] = l < He ? take(H) : -1;
    }
    out->enc_wh[m] = take((int64_t)(2 * L + 2 * S) * (He >= 1 ? H : D));
    out->enc_bh[m] = take(2 * L + 2 * S);
  }
  for (int m = 0; m < d->n_mods; ++m) {
    const int D = d->dims[m], S = d->style_dims[m];
    const int in_o = Hd >= 1 ? H : S + L;
    out->dec_lv[m] = d->scale_mode == 0 ? take(D) : -1;
    for (int l = 0; l < MOPOE_MAX_LAYERS; ++l) {
      out->dec_hw[m][l] = l < Hd ? take((int64_t)H * (l == 0 ? S + L : H)) : -1;
      out->dec_hb[m][l] = l < Hd ? take(H) : -1;
    }
    out->dec_w[m] = take((int64_t)D * in_o);
    out->dec_b[m] = take(D);
    out->dec_lvw[m] = d->scale_mode == 1 ? take((int64_t)D * in_o) : -1;
    out->dec_lvb[m] = d->scale_mode == 1 ? take(D) : -1;
  }
  out->total = off;
  return MOPOE_OK;
}

int mopoe_philox_normal(uint64_t seed, uint64_t stream_id, int64_t start, int64_t n, float* out, void* stream) {
  if (mopoe_device_count() == 0) { set_error("no CUDA device"); return MOPOE_ENODEV; }
  if (n <= 0) return MOPOE_OK;
  int64_t blocks = (n + 255) / 256;
  philox_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seed, stream_id, start, n, out);
  MOPOE_CUDA(cudaGetLastError());
  return MOPOE_OK;
}

}  // extern "C"
