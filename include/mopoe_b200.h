/*
 * mopoe_b200.h -- C-ABI of the B200-native MoPoE-VAE hot path.
 *
 * Drop-in boundary for the one data-parallel hot path of
 * neurospin-projects/2022_cambroise_interpret_multivae: the joint-ELBO step of the MoPoE-VAE and the
 * Digital Avatars Analysis (DAA) sweep that re-runs its forward pass.  The reference has no FFI of
 * its own (it is pure Python / eager PyTorch); the seams this library replaces are the Python
 * functions cited on each entry point (paths relative to <reference>/experiments).  The host-side
 * mirror of those functions (same names / arguments / return structure) lives in the Python package
 * `2022_cambroise_interpret_multivae_b200` and binds this header through ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in
 *     `_host`; all buffers are caller-owned and pre-allocated, the library never allocates on the
 *     data path (scratch comes in through `workspace`, sized by mopoe_workspace_bytes).
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising; 0 = success, negative = MOPOE_E* (message via mopoe_last_error()).
 *   - float tensors are fp32 row-major contiguous; statistics tables are fp64.
 *   - noise: `eps` tensors are INJECTED standard-normal draws laid out as documented per call.
 *     A NULL eps pointer selects the built-in counter-based generator
 *     eps[i] = philox_normal(seed, stream_id, i)  (i = flat index into the same layout), so results
 *     do not depend on how a sweep is sharded over GPUs (oracle/philox.py restates it).  The two
 *     LATENT noise tensors of the DAA sweep (eps_base, eps_av: E columns per row) are addressed by
 *     row instead: row r is drawn as whole Philox blocks, counter r * (EP / 4) + b, laid out
 *     [content | style_0 | style_1 ...] with every section padded to a multiple of 4 normals
 *     (EP = padded width; oracle/philox.py: philox_rows), so the thread that owns a row draws whole
 *     blocks and nothing straddles rows.
 *   - no CPU fallback: with no CUDA device every compute entry point returns MOPOE_ENODEV.
 */
#ifndef MOPOE_B200_H
#define MOPOE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOPOE_MAX_MODS 4
#define MOPOE_MAX_SUBSETS 15 /* 2^4 - 1 */
#define MOPOE_HIDDEN 256     /* networks.py:15,50 hard-code the hidden width */
#define MOPOE_N_SCALARS 64

enum { MOPOE_METHOD_POE = 0, MOPOE_METHOD_MOE = 1, MOPOE_METHOD_JOINT_ELBO = 2,
       MOPOE_METHOD_JSD = 3 /* BaseMMVae.py:51-54: MoE subset posteriors, unimodal experts + the prior N(0, I) in the
                             * mixture (:217-223), divergence to the dynamic prior (:81-93, mm_div.py:23-35,69-89) */ };

enum {
  MOPOE_OK = 0,
  MOPOE_EINVAL = -1,   /* bad argument / unsupported configuration (message says which) */
  MOPOE_ENODEV = -2,   /* no CUDA device */
  MOPOE_ECUDA = -3,    /* CUDA runtime error (message carries cudaGetErrorString) */
  MOPOE_ENOSPC = -4,   /* workspace too small */
  MOPOE_EDEVICE = -5   /* a kernel flagged a device-side protocol error (bounded tcgen05 / mbarrier wait timed out) */
};

/* Model description: the subset of the reference `flags` namespace (workflow.py:98-149) that the
 * hot path reads.  Unsupported values are rejected with MOPOE_EINVAL, never silently ignored. */
typedef struct mopoe_model_desc {
  int32_t n_mods;                       /* len(flags.input_dim), 1..4 */
  int32_t dims[MOPOE_MAX_MODS];         /* flags.input_dim */
  int32_t style_dims[MOPOE_MAX_MODS];   /* flags.style_dim, zeros when not factorised (workflow.py:148) */
  int32_t latent_dim;                   /* flags.class_dim, 1..32 */
  int32_t hidden;                       /* must be 256 */
  int32_t n_hidden_enc;                 /* flags.num_hidden_layer_encoder, 0..4 (1 = the fused kernels, else the layered path) */
  int32_t n_hidden_dec;                 /* flags.num_hidden_layer_decoder, 0..4 (0 = the fused kernels, else the layered path) */
  int32_t method;                       /* MOPOE_METHOD_* (flags.modality_poe/moe/jsd/joint_elbo) */
  int32_t likelihood;                   /* 0 = normal, 1 = laplace (modality.py:18-30; layered path); others rejected */
  int32_t scale_mode;                   /* 0 = per-feature logvar Parameter (networks.py:61-64), 1 = learn_output_sample_scale:
                                           logvar = Linear(decoder features) per sample (networks.py:58-59,73-74; layered path) */
  int32_t learn_output_scale;           /* flags.learn_output_scale */
  int32_t name_rank[MOPOE_MAX_MODS];    /* rank of modality m's NAME in sorted order: fusion order
                                           inside a subset (BaseExperiment.py:72-77) */
  float beta, beta_style, beta_content; /* flags.beta, beta_style, beta_content */
} mopoe_model_desc;

/* Offsets (in floats) of every parameter block inside the flat parameter buffer.  Each block keeps
 * torch's own [out, in] row-major layout so that nn.Parameters can be views of the buffer and the
 * reference's state-dict keys/shapes are preserved:
 *   enc_w1  = encoders.<m>.shared_encoder.0.weight (256, D)      enc_b1 = ....bias (256)
 *   enc_wh  = rows [class_mu (L); class_logvar (L); style_mu (S); style_logvar (S)] x 256
 *   enc_bh  = the matching biases (2L + 2S)
 *   dec_w   = decoders.<m>.out_mu.weight (D, S + L)              dec_b = ....bias (D)
 *   dec_lv  = decoders.<m>.logvar (1, D)
 * Adam moments and gradient buffers use the same layout. */
#define MOPOE_MAX_LAYERS 4
typedef struct mopoe_param_layout {
  int64_t enc_w1[MOPOE_MAX_MODS], enc_b1[MOPOE_MAX_MODS];
  int64_t enc_wh[MOPOE_MAX_MODS], enc_bh[MOPOE_MAX_MODS];
  int64_t dec_w[MOPOE_MAX_MODS], dec_b[MOPOE_MAX_MODS], dec_lv[MOPOE_MAX_MODS];
  int64_t total; /* floats */
  /* non-default architectures (networks.py:16-20,51-59), -1 where a block does not exist:
   *   enc_wx[m][l-1] / enc_bx = encoders.<m>.shared_encoder.<3l>.weight (256, 256) / bias, hidden layers l = 1 ..
   *     (with num_hidden_layer_encoder = 0 enc_w1 / enc_b1 are -1 and the heads enc_wh are (2L+2S, D))
   *   dec_hw[m][l] / dec_hb    = decoders.<m>.shared_decoder.<3l>.weight (256, S+L | 256) / bias
   *     (out_mu dec_w is then (D, 256))
   *   dec_lvw[m] / dec_lvb[m]  = decoders.<m>.logvar.weight (D, in) / bias when learn_output_sample_scale
   *     (dec_lv is -1: the per-feature Parameter does not exist) */
  int64_t enc_wx[MOPOE_MAX_MODS][MOPOE_MAX_LAYERS - 1], enc_bx[MOPOE_MAX_MODS][MOPOE_MAX_LAYERS - 1];
  int64_t dec_hw[MOPOE_MAX_MODS][MOPOE_MAX_LAYERS], dec_hb[MOPOE_MAX_MODS][MOPOE_MAX_LAYERS];
  int64_t dec_lvw[MOPOE_MAX_MODS], dec_lvb[MOPOE_MAX_MODS];
} mopoe_param_layout;

/* One batch of the path.  A batch is homogeneous in its set of present modalities
 * (MissingModalitySampler, dataset.py:295-354); absent modalities are neither encoded nor decoded
 * (BaseMMVae.py:156,167-178).  `joint_bounds`/`moe_bounds` are the row boundaries of
 * utils.mixture_component_selection (utils/utils.py:63-85), computed ON THE HOST with the
 * reference's own fp32 expression int(floor(N * w_k)) so they are bit-faithful. */
typedef struct mopoe_batch_desc {
  int32_t n_rows;
  int32_t present_mask;                     /* bit m set <=> modality m is a key of input_batch */
  int32_t n_mix;                            /* K = number of subsets entering the mixture */
  int32_t joint_bounds[MOPOE_MAX_SUBSETS + 1]; /* K + 1 boundaries of the joint selection */
  int32_t moe_bounds[MOPOE_MAX_MODS + 1][MOPOE_MAX_MODS + 1]; /* [k][0..k]: selection inside a
                                               k-member subset (method = moe only) */
  int64_t row_offset;                       /* first entry of this batch in `row_index` */
  int32_t owner_div, owner_mod;             /* 0, 0: row n selects its mixture component by its own index (the reference's
                                               batches).  owner_mod = P > 0: the rows are P-row batches of the reference laid
                                               out side by side -- row n behaves like row (n / owner_div) % P of a P-row batch
                                               (the bounds then span P rows).  One launch can so carry the M base passes or the
                                               n_samples x n_scores perturbed forwards of a DAA validation (daa.py:
                                               daa_sweep_layered).  Noise and outputs stay indexed by n. */
} mopoe_batch_desc;

/* Outputs of one forward pass == the `results` dict of BaseMMVae.forward (BaseMMVae.py:137-165).
 * Any pointer may be NULL (that output is skipped).  Subset arrays are indexed by the subset's
 * position in BaseExperiment.set_subsets order without the empty set; rows of unavailable subsets
 * are left untouched. */
typedef struct mopoe_forward_out {
  float* enc_heads[MOPOE_MAX_MODS]; /* (N, 2L+2S_m): [class_mu | class_logvar | style_mu | style_logvar] */
  float* subset_mu;                 /* (n_subsets, N, L)   latents['subsets'][key][0] */
  float* subset_logvar;             /* (n_subsets, N, L) */
  float* joint_mu;                  /* (N, L)              latents['joint'][0] */
  float* joint_logvar;              /* (N, L) */
  float* z;                         /* (N, L)              class_embeddings */
  float* z_style[MOPOE_MAX_MODS];   /* (N, S_m) */
  float* rec_loc[MOPOE_MAX_MODS];   /* (N, D_m)            results['rec'][m].loc  (scale = exp(.5*logvar)) */
  float* scalars;                   /* (MOPOE_N_SCALARS)   see mopoe_scalar_index */
  float* rec_logvar[MOPOE_MAX_MODS];/* (N, D_m)            per-sample output log-variance (scale_mode 1 only) */
} mopoe_forward_out;

/* Index of each entry of a `scalars` row (all are means over rows, exactly the numbers
 * basic_routine_epoch returns and TBLogger.add_basic_logs logs, TBLogger.py:84-91). */
enum mopoe_scalar_index {
  MOPOE_S_TOTAL_LOSS = 0,   /* total_loss                       run_epochs.py:103,128 */
  MOPOE_S_JOINT_DIV = 1,    /* results['joint_divergence']      BaseMMVae.py:147-149 */
  MOPOE_S_NLL = 2,          /* +m: log_probs[m]                 run_epochs.py:27-38 */
  MOPOE_S_NLL_UNI = 6,      /* +m: unimodal-pass -log p (poe)   run_epochs.py:117-119 */
  MOPOE_S_KLD_SUBSET = 10,  /* +s: klds[subset s]               run_epochs.py:41-48 */
  MOPOE_S_KLD_STYLE = 25,   /* +m: klds_style[m]                run_epochs.py:51-59 */
  MOPOE_S_MEAN_HEAD = 29,   /* +4m+{0,1,2,3}: mean class_mu, class_logvar, style_mu, style_logvar */
  MOPOE_S_N_ROWS = 45,
  MOPOE_S_PRESENT = 46,
  MOPOE_S_JSD_DIV = 47      /* +k: results['individual_divs'][k] of method jsd (k-th present modality, then the prior):
                             * KL(component k || dynamic prior) / N                 mm_div.py:69-89 */
};

const char* mopoe_last_error(void);
int mopoe_version(void);
/* Number of CUDA devices visible (0 => every compute call returns MOPOE_ENODEV). */
int mopoe_device_count(void);

/* Validate `desc` and fill the parameter layout. */
int mopoe_param_layout_of(const mopoe_model_desc* desc, mopoe_param_layout* out);

/* Bytes of scratch the model calls need for batches of up to `max_rows` rows
 * (`n_pass` = 1, or 1 + n_mods when method = poe trains with its unimodal passes). */
int64_t mopoe_workspace_bytes(const mopoe_model_desc* desc, int64_t max_rows);

/* BaseMMVae.forward(input_batch, sample_latents, use_expert)           BaseMMVae.py:137-165
 * (encode :167-178, inference :181-239, poe mm_div.py:13-20, mixture_component_selection
 * utils.py:63-85, reparameterize :37-40, Decoder.forward networks.py:66-77).
 *   x[m]        (N, D_m) rows of modality m, NULL/ignored when absent from present_mask
 *   eps         (N, E) injected noise, E = L + sum_m S_m: cols [0,L) joint content, then the style
 *               block of each modality in order; NULL => philox(seed, MOPOE_STREAM_FORWARD)
 *   use_expert  subset index whose posterior becomes the joint (BaseMMVae.py:230-231), -1 = none
 *   with_nll    also fill the NLL scalars (needs x as target; run_epochs.calc_log_probs) */
int mopoe_forward(const mopoe_model_desc* desc, const float* params, const mopoe_batch_desc* batch_host,
                  const float* const* x_host_ptrs, const float* eps, uint64_t seed, int sample_latents,
                  int use_expert, int with_nll, const mopoe_forward_out* out_host, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* run_epochs.basic_routine_epoch + loss.backward() [+ Adam.step()]   run_epochs.py:73-135,180-182
 * executed for `n_steps` consecutive batches inside ONE cooperative persistent kernel launch.
 *   mode 0: losses only (run_epochs.test under no_grad, :187-219)
 *   mode 1: losses + gradients of total_loss w.r.t. every parameter into `grads` (flat layout;
 *           n_steps must be 1) -- backs the torch.autograd.Function of the Python mirror
 *   mode 2: losses + gradients + in-kernel Adam update of `params`, `adam_m`, `adam_v`
 *           (torch.optim.Adam as configured by experiment.py:268-271: betas (b1,b2), eps, no weight
 *           decay; parameters of absent modalities are skipped entirely like torch skips grad=None)
 *   data[m]       (n_data_rows_m, D_m) resident dataset block of modality m
 *   row_index     int32 rows of each step's batch (batches[s].row_offset .. +n_rows), one list per
 *                 modality (row_index[m]); NULL => rows 0..n_rows-1
 *   eps           (n_steps, n_pass, max_rows, E) injected noise or NULL => philox(seed, TRAIN)
 *   adam_t        int32 (n_mods) per-modality step counters (device), updated in mode 2
 *   scalars       (n_steps, MOPOE_N_SCALARS) fp32
 *   batches       device array of n_steps mopoe_batch_desc
 *   out_host      optional forward outputs of the (single) step, n_steps must be 1: lets
 *                 basic_routine_epoch return its `results` dict from the same launch */
int mopoe_train_steps(const mopoe_model_desc* desc, float* params, float* adam_m, float* adam_v,
                      int32_t* adam_t, float* grads, const float* const* data_host_ptrs,
                      const int32_t* const* row_index_host_ptrs, const mopoe_batch_desc* batches,
                      int32_t n_steps, int64_t max_rows, const float* eps, uint64_t seed, int mode,
                      float lr, float b1, float b2, float adam_eps, float* scalars,
                      const mopoe_forward_out* out_host, void* workspace, int64_t workspace_bytes,
                      void* stream);
/* implementation of the last mopoe_train_steps call: 1 = tensor-core kernel (tcgen05 / TMEM / TMA bulk copies,
 * csrc/mopoe_train_tc.cuh; the default whenever the configuration fits its tiling), 0 = CUDA-core kernel.
 * MOPOE_TRAIN_IMPL=tc|ffma in the environment forces one (the tests cross-check both). */
int mopoe_train_last_impl(void);

/* ---- Digital Avatars Analysis  (workflow.daa_exp, workflow.py:361-537) ----------------------- */

typedef struct mopoe_daa_desc {
  int32_t n_val;        /* validations handled by THIS call (a shard of params.n_validation) */
  int32_t val_begin;    /* global index of the first one (keys the philox streams) */
  int32_t n_val_total;  /* params.n_validation (sizes the philox index space) */
  int32_t n_subjects;   /* rows per validation batch (params.n_subjects) */
  int32_t n_samples;    /* params.n_samples */
  int32_t n_base;       /* params.M stochastic reconstructions */
  int32_t src_mod;      /* perturbed modality ("clinical") */
  int32_t dst_mod;      /* read-out modality ("rois") */
  int32_t sample_latents;
  int32_t reg_method;   /* 0 = hierarchical (stat_utils.py:66-75), 1 = fixed (:62-63) */
  int32_t base_mode;    /* how the mean over the M stochastic reconstructions (workflow.py:388-398) gets its noise:
                         * 0 = average M drawn noise rows (draw-for-draw what M reference forwards consume; required with
                         *     injected eps_base), 1 = draw the MEAN row directly, eps_mean ~ N(0, 1/M) (one Philox row per
                         *     subject scaled by 1/sqrt(M)): the default decoders are affine in z, so the mean of the M decodes
                         *     is the decode of mu + sd * eps_mean -- same distribution, 1/M of the draws; only with the
                         *     in-kernel generator (eps_base == NULL).  Equivalent to base_mode 0 with n_base = 1 and
                         *     eps_base = that row / sqrt(M) (tests/test_gpu_parity.py). */
  int32_t unit_begin;   /* shard units: unit u = v_local * n_scores + score, the regression of a (validation, */
  int32_t unit_end;     /* score) pair needs nothing outside it.  [unit_begin, unit_end) of the n_val * n_scores units of this
                         * call are OWNED: their rows of coefs / pvalues / betas and their avatars are produced; the other units
                         * of the first and last validation belong to another rank -- their table rows are NOT written (avatar
                         * tiles shared with an owned series are, with the values the owner computes: the noise is keyed
                         * globally).  0, 0 = every unit.  Base passes and scores cover whole validations either way. */
  int32_t score_mode;   /* 0 = sampling_strategy "likelihood" (workflow.py:401-405): scores = loc_hat + scale_hat * noise;
                         * 1 = the caller provides the artificial score VALUES in `eps_score` (n_val, n_samples, N, C), e.g.
                         *     sampling_strategy "linear" (workflow.py:337-346: a linspace between population quantiles) */
} mopoe_daa_desc;

/* Bytes of scratch mopoe_daa_sweep needs. */
int64_t mopoe_daa_workspace_bytes(const mopoe_model_desc* desc, const mopoe_daa_desc* daa);

/* The whole DAA sweep for `n_val` validations (all modalities present):
 *   workflow.py:388-400  M stochastic reconstructions -> mean src loc/scale, mean dst loc
 *   workflow.py:401-405  scores ~ Normal(loc_hat, scale_hat)          (sampling = "likelihood")
 *   workflow.py:406-419  one forward per (sample, score) with ONE src column overwritten -> dst loc
 *   workflow.py:466-505  make_regression per (validation, score, roi)  (stat_utils.py:55-79)
 * inputs
 *   x[m]        (n_val, n_subjects, D_m) the drawn test batches (every modality)
 *   eps_base    (n_val, n_base, N, E), eps_score (n_val, n_samples, N, C), eps_av
 *               (n_val, n_samples, C, N, E); each NULL => philox(seed, MOPOE_STREAM_DAA_*)
 *               (eps_base / eps_av by block-padded rows, eps_score by flat index, see "noise" above)
 * outputs (each may be NULL except coefs/pvalues)
 *   avatars         fp32 (n_val, N, C, n_samples, R)  rois_digital_avatars.npy (workflow.py:280-288)
 *   sampled_scores  fp32 (n_val, N, n_samples, C)     sampled_scores.npy       (workflow.py:435)
 *   reconstructions fp32 (n_val, N, R)                rois_reconstructions.npy (workflow.py:437)
 *   betas           fp64 (n_val, C, N, R)             per-subject slopes -> all_coefs.npy (:499-505)
 *   coefs, pvalues  fp64 (n_val, C, R)                coefs.npy / pvalues.npy  (workflow.py:510-511)
 * C = dims[src_mod], R = dims[dst_mod], N = n_subjects. */
int mopoe_daa_sweep(const mopoe_model_desc* desc, const float* params, const mopoe_daa_desc* daa,
                    const mopoe_batch_desc* batch_host, const float* const* x_host_ptrs,
                    const float* eps_base, const float* eps_score, const float* eps_av, uint64_t seed,
                    float* avatars, float* sampled_scores, float* reconstructions, double* betas,
                    double* coefs, double* pvalues, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* stat_utils.make_regression(method="hierarchical"|"fixed") on an avatar tensor that is already in
 * device memory (same layouts as above) -- the statistics stage alone (workflow.py:466-505). */
int mopoe_daa_regression(int32_t n_val, int32_t n_subjects, int32_t n_scores, int32_t n_samples,
                         int32_t n_rois, int32_t reg_method, const float* avatars,
                         const float* sampled_scores, const float* reconstructions, double* betas,
                         double* coefs, double* pvalues, void* stream);

/* ---- multi-GPU: the association tables over NVLink peer memory (the only exchange step) ------
 * Each rank computes the (n_val_local, C, R) slice of coefs / pvalues of ITS validations.  Instead of an NCCL
 * all_gather behind the sweep, mopoe_daa_exchange_tables stores the slice into every rank's FULL table through
 * peer-mapped pointers and releases a per-source sequence flag (system scope); a one-warp kernel then acquires
 * the flags of all sources.  Two small launches on `stream`, capturable in a CUDA graph (the sequence number lives
 * in the buffer).  The buffers are symmetric memory the host allocates and maps on every rank (Python:
 * torch.distributed._symmetric_memory, see daa.TableExchange); they must be zero-filled before the first call.
 * After the call, slot (number of calls so far) & 1 of the local buffer holds the full tables:
 *   local_base + 256 + slot * 2 * elems_total * 8 : coefs (elems_total doubles), then pvalues. */
#define MOPOE_MAX_PEERS 8
typedef struct mopoe_table_exchange {
  int32_t world, rank;
  int32_t root;          /* -1: every rank receives every slice (all-gather); >= 0: only `root` receives (gather -- what
                          * daa_exp needs: 1/world of the bytes and no rank but the root ever waits; senders stay at most two
                          * calls ahead of the root, which acknowledges every completed step) */
  int32_t reserved;
  int64_t elems_local;   /* n_val_local * C * R */
  int64_t elem_offset;   /* val_begin * C * R */
  int64_t elems_total;   /* n_val_total * C * R */
  void* peer_base[MOPOE_MAX_PEERS];   /* exchange buffer of every rank as mapped in THIS process; [rank] = own */
} mopoe_table_exchange;
int64_t mopoe_table_exchange_bytes(int64_t elems_total);
int mopoe_daa_exchange_tables(const mopoe_table_exchange* ex, const double* coefs_local, const double* pvalues_local,
                              void* stream);

/* Measurement hook (bench.py): when enabled, mopoe_daa_sweep brackets its dominant kernel
 * (daa_avatar_kernel) with CUDA events on the caller's stream; mopoe_daa_last_kernel_ms waits for
 * the last bracket and returns its duration. */
int mopoe_profile_enable(int on);
/* which avatar kernel the last mopoe_daa_sweep used: 2 = warp-specialised tcgen05 pipeline (the
 * production kernel: hierarchical regression, sampled latents), 1 = phase-serial tcgen05 kernel (also
 * "fixed" regression / mean latents), 0 = CUDA cores (shapes outside the tcgen05 tilings);
 * MOPOE_DAA_IMPL=pipe|umma|ffma in the environment forces one (the tests cross-check all three) */
int mopoe_daa_last_impl(void);
/* Synchronises `stream` and reports whether the last sweep run on `workspace` hit a device-side protocol error
 * (a bounded tcgen05 / mbarrier wait that timed out: the kernels then poison coefs / pvalues with NaN instead of
 * hanging the GPU).  MOPOE_OK, or MOPOE_EDEVICE with mopoe_last_error() set.  Hosts call it before they trust or
 * write the tables (daa.py: check_status; workflow.daa_exp refuses to write results otherwise). */
int mopoe_daa_status(const mopoe_model_desc* desc, const mopoe_daa_desc* daa, void* workspace, void* stream);
/* per-role cycle counters (max over CTAs, 32 slots) of the last tcgen05 avatar kernel run on
 * `workspace`; filled by profiling builds of the library only (csrc/Makefile EXTRA=-DPK_PROF) */
int mopoe_daa_read_phases(const mopoe_model_desc* desc, const mopoe_daa_desc* daa, void* workspace, int64_t* out32_host);
int mopoe_daa_last_kernel_ms(float* ms_out);

/* Fill `out[0..n)` with philox_normal(seed, stream_id, start + i): the production noise generator,
 * exposed so hosts/tests can materialise exactly what the kernels draw. */
int mopoe_philox_normal(uint64_t seed, uint64_t stream_id, int64_t start, int64_t n, float* out,
                        void* stream);

/* ---- representational similarity analysis (experiments/workflow.py:656-789 rsa_exp) ----
 * mopoe_rsa_cmat: stat_utils.py:25-33 data2cmat (categorical = 0: Euclidean distances of the n rows of `data` (n, d),
 *   fp64, summed in column order like scipy's pdist) or stat_utils.py:46-53 vec2cmat (d = 1; categorical = 1: the
 *   0 / 1 "differs" matrix).  cmat (n, n) fp64. */
int mopoe_rsa_cmat(int32_t n, int32_t d, const float* data, int32_t categorical, double* cmat, void* stream);
/* mopoe_rsa_kendall: stat_utils.py:81-95 fit_rsa for ONE matrix against n_ref reference matrices (n_ref, n, n): the
 *   upper triangles (stat_utils.py:36-43 cmat2triu, P = n (n - 1) / 2 entries) are compared entry pair by entry pair.
 *   counts (n_ref, 7) int64, exact:  [0] sum_{i != j} sign(x_i - x_j) sign(y_i - y_j) = 2 (concordant - discordant),
 *   [1..3] sum_i c_i, sum_i c_i (c_i - 1), sum_i c_i (2 c_i + 7) with c_i = #{j != i: x_j == x_i} (i.e. the sums over
 *   tie groups of t (t - 1), t (t - 1)(t - 2), t (t - 1)(2 t + 5) that scipy.stats.kendalltau uses), [4..6] the same
 *   for y.  tau-b and the asymptotic p-value follow on the host (rsa.py: kendall_from_counts). */
int64_t mopoe_rsa_kendall_workspace_bytes(int32_t n, int32_t n_ref);
int mopoe_rsa_kendall(int32_t n, int32_t n_ref, const double* cmat, const double* ref_cmats, int64_t* counts,
                      void* workspace, int64_t workspace_bytes, void* stream);

enum {
  MOPOE_STREAM_DAA_BASE = 1, MOPOE_STREAM_DAA_SCORE = 2, MOPOE_STREAM_DAA_AVATAR = 3,
  MOPOE_STREAM_TRAIN = 4, MOPOE_STREAM_FORWARD = 5
};

#ifdef __cplusplus
}
#endif
#endif /* MOPOE_B200_H */
