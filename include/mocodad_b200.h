/*
 * mocodad_b200.h -- C ABI of the B200-native MoCoDAD reverse-diffusion scoring path.
 *
 * The reference (aleflabo/MoCoDAD) is pure Python/PyTorch and has no plugin / FFI surface
 * (SURVEY.md section 8b): its hot path is the body of MoCoDAD.forward, models/mocodad.py:129-184.
 * This header is therefore the boundary a maintainer binds *beneath* that method (ctypes stub in
 * INTEGRATION.md; mocodad_b200/_lib.py is that stub).  Every entry point below names the
 * reference code it replaces.  Conventions:
 *
 *   - plain C types, raw pointers and sizes; no torch / C++ types cross the boundary;
 *   - every function returns an int status: MCD_OK (0) or a negative mcd_status; the message
 *     for the calling thread's last failure is mcd_last_error();
 *   - pointers prefixed d_ are DEVICE pointers on the model's device, h_ are HOST pointers;
 *   - device entry points are asynchronous on `stream` (a cudaStream_t passed as void*); they
 *     never allocate: scratch comes from the caller (`d_ws`, sized by mcd_workspace_bytes);
 *   - a model handle is immutable after mcd_model_finalize and may be shared by threads that
 *     each use their own stream + workspace (exceptions: mcd_score_windows_host and the
 *     profiling calls, which use state cached on the handle).
 *
 * Tensor layouts are the reference's: skeleton windows are float32 [B, C=2, T, V] contiguous.
 * "Virtual batch" = n_generated_samples x B windows, sample-major (index g*B + b), which is the
 * order the reference visits them in (mocodad.py:160-180).
 */
#ifndef MOCODAD_B200_H
#define MOCODAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCD_ABI_VERSION 5

#if defined(__GNUC__)
#define MCD_API __attribute__((visibility("default")))
#else
#define MCD_API
#endif

typedef enum mcd_status {
  MCD_OK = 0,
  MCD_ERR_INVALID_ARG = -1,      /* NULL pointer, negative size, unknown tensor name ...        */
  MCD_ERR_UNSUPPORTED = -2,      /* shape / strategy outside the compiled kernel set            */
  MCD_ERR_NOT_FINALIZED = -3,    /* compute call before mcd_model_finalize                      */
  MCD_ERR_MISSING_TENSOR = -4,   /* finalize with state_dict entries never set                  */
  MCD_ERR_CUDA = -5,             /* CUDA runtime failure (message holds cudaGetErrorString)     */
  MCD_ERR_WORKSPACE = -6         /* workspace too small for even one tile                       */
} mcd_status;

/* Loss used for the per-window score: models/mocodad.py:24,66 (`losses` table). */
typedef enum mcd_loss_fn { MCD_LOSS_SMOOTH_L1 = 0, MCD_LOSS_L1 = 1, MCD_LOSS_MSE = 2 } mcd_loss_fn;

/* Hyper-parameters the reference reads from the YAML Namespace (models/mocodad.py:46-81) that
 * shape the path.  The U-Net channel plan (16,32,32,64,64,128,64 / 64,32,32,2) and the joint
 * pyramid 17/12/10 are hard-wired exactly as in models/stsae/stsae_unet.py:11,14,229-230. */
typedef struct mcd_config {
  int32_t n_coords;        /* num_coords (2)                                                    */
  int32_t n_joints;        /* 17 (the reference's pyramid only accepts 17)       */
  int32_t n_frames;        /* seg_len                                                           */
  int32_t n_frames_cond;   /* len(conditioning_indices); 0 => 'no_condition'                    */
  int32_t cond_first;      /* 1: conditioning frames are [0, n_cond); 0: the last n_cond frames */
  int32_t embedding_dim;   /* embedding_dim == latent_dim (16)                                  */
  int32_t cond_h_dim;      /* h_dim (32)                                                        */
  int32_t cond_channels[3];/* channels ([32,16,32])                                             */
  int32_t noise_steps;     /* noise_steps N: the loop runs i = N-1 .. 1                         */
  int32_t loss_fn;         /* mcd_loss_fn                                                       */
  int32_t device;          /* CUDA device ordinal                                               */
  /* Latent variant (models/mocodad_latent.py, YAML keys latent_embedding_dim / hidden_sizes; selected by the presence of
   * `diffusion_on_latent`, eval_MoCoDAD.py:24).  latent_dim == 0: diffusion on the poses (the fields below are ignored). */
  int32_t latent_dim;      /* latent_embedding_dim (64); 0 = pose-space model                    */
  int32_t n_hidden;        /* len(hidden_sizes), 1..8                                            */
  int32_t hidden[8];       /* hidden_sizes ([64,128,128,64]); the last must equal latent_dim     */
} mcd_config;

typedef struct mcd_model mcd_model;

/* ---- library ------------------------------------------------------------------------------ */
MCD_API int mcd_abi_version(void);
MCD_API const char* mcd_last_error(void);
/* 1 when this build carries kernels for (n_frames_denoised T, n_frames_cond). */
MCD_API int mcd_shape_supported(int32_t T, int32_t T_cond);

/* ---- a1: noise schedule, utils/diffusion_utils.py:8-44 + models/mocodad.py:799-808 ---------
 * HOST function (no GPU needed).  Writes beta, alpha, alpha_hat (each [noise_steps] float32),
 * bit-identical to the reference's fp32 tensors.  Any output pointer may be NULL. */
MCD_API int mcd_schedule(int32_t noise_steps, float* h_beta, float* h_alpha, float* h_alpha_hat);
/* a4: STSE_Unet.pos_encoding, models/stsae/stsae_unet.py:161-179, for integer step t. HOST. */
MCD_API int mcd_pos_encoding(int32_t t, int32_t channels, float* h_out);
/* a9: the three per-step DDPM coefficients (mocodad.py:172-178) in the reference's fp32
 * arithmetic: c1 = 1/sqrt(alpha_t), c2 = (1-alpha_t)/sqrt(1-alpha_hat_t), c3 = sqrt(beta_t). HOST. */
MCD_API int mcd_ddpm_coefficients(int32_t noise_steps, int32_t t, float* c1, float* c2, float* c3);

/* ---- model lifetime: replaces MoCoDAD.build_model + load_state_dict (mocodad.py:90-126) ---- */
MCD_API int mcd_model_create(const mcd_config* cfg, mcd_model** out);
/* Hand over one state_dict entry by its REFERENCE name, e.g. "model.st_gcnnsd1.0.gcn.A" or
 * "condition_encoder.encoder.model_layers.2.tcn.1.running_var" (SURVEY.md section 8b lists all 335).
 * `h_data` is host float32, copied.  Entries the path does not use (decoder.*, rev_btlnk.*,
 * num_batches_tracked) are accepted and ignored so a whole checkpoint can be streamed in. */
MCD_API int mcd_model_set_tensor(mcd_model* m, const char* name, const float* h_data, int64_t numel);
/* Fold eval-mode BatchNorm into the 1x1 convolutions, lay weights out for the kernels, upload.
 * May be called again after further mcd_model_set_tensor calls (re-packs). */
MCD_API int mcd_model_finalize(mcd_model* m);
MCD_API void mcd_model_destroy(mcd_model* m);

/* Scratch bytes for the compute calls below when they process `n_virtual` windows in one pass
 * (mcd_reverse_diffusion accepts less and tiles the virtual batch; it needs at least
 * mcd_workspace_bytes(m, 1) + 4*(B*embedding_dim + G*B) bytes). */
MCD_API size_t mcd_workspace_bytes(const mcd_model* m, int64_t n_virtual);

/* ---- a3: MoCoDAD._encode_condition -> STSE.encode, stsae.py:59-92 --------------------------
 * d_data [B,C,n_frames,V] (the whole window; conditioning frames are selected in-kernel, which
 * replaces _select_frames, mocodad.py:743-748).  d_cond_emb [B, embedding_dim]. */
MCD_API int mcd_cond_encode(const mcd_model* m, const float* d_data, int64_t B, float* d_cond_emb,
                    void* d_ws, size_t ws_bytes, void* stream);

/* ---- a5..a8: one STSAE_Unet.forward, stsae_unet.py:406-438 ---------------------------------
 * d_x [n,2,T,V]; step t (uniform over the batch, mocodad.py:166); d_cond_emb [cond_B, E] with
 * window w using row (w % cond_B) (NULL when n_frames_cond == 0).  d_eps [n,2,T,V]. */
MCD_API int mcd_unet_forward(const mcd_model* m, const float* d_x, int64_t n, int32_t t,
                     const float* d_cond_emb, int64_t cond_B, float* d_eps,
                     void* d_ws, size_t ws_bytes, void* stream);

/* Debug/parity taps: run the denoiser like mcd_unet_forward and copy the activation produced by
 * layer `layer_name` ("st_gcnnsd1.0", "down1", ...) to d_out in the REFERENCE layout [n,C,T,V'].
 * (For "up3"/"up2" the tap is the CNN_layer output before the skip add, as a forward hook on
 * the reference module sees it.) */
MCD_API int mcd_unet_tap(const mcd_model* m, const float* d_x, int64_t n, int32_t t,
                 const float* d_cond_emb, int64_t cond_B, const char* layer_name, float* d_out,
                 void* d_ws, size_t ws_bytes, void* stream);

/* ---- a9/a10: DDPM update, mocodad.py:172-178 ------------------------------------------------
 * x <- 1/sqrt(alpha_t) * (x - (1-alpha_t)/sqrt(1-alpha_hat_t) * eps) + sqrt(beta_t) * z, in place,
 * for n windows [n,2,T,V].  z = d_noise (same shape) when non-NULL; otherwise z ~ N(0,1) from
 * counter-based Philox4x32-10 keyed by (seed; first_window + w, sample, noise_slot, element) --
 * independent of batching and rank count.  At t == 1 no noise is added (mocodad.py:176). */
MCD_API int mcd_ddpm_step(const mcd_model* m, float* d_x, const float* d_eps, const float* d_noise,
                  int64_t n, int32_t t, uint64_t seed, int64_t first_window, int32_t sample,
                  int32_t noise_slot, void* stream);
/* x_T ~ N(0,1) from the same Philox stream (noise_slot 0), mocodad.py:162. */
MCD_API int mcd_randn_windows(const mcd_model* m, float* d_x, int64_t n, uint64_t seed,
                      int64_t first_window, int32_t sample, int32_t noise_slot, void* stream);

/* ---- f1 (first "next" row): dataset items = affine transforms of base windows ------
 * mcd_pose_transform_matrix: rows 0-1 (6 floats, row-major) of ae_trans_list[index], index 0..4
 * (utils/dataset_utils.py:255-270, 308-314) -- identity, flip, rot 90, rot 90 + flip, rot 45.
 * mcd_expand_transforms: PoseDataset.__getitem__ (utils/dataset.py:67-76) + apply_pose_transform
 * (utils/dataset_utils.py:273-290) on the device: item idx = first_item + i is transform idx / N of base
 * window idx % N.  d_base [N,2,n_frames,V] (x, y), h_mats [num_transform][6], d_out [n_items,2,n_frames,V].
 * The base windows cross PCIe once instead of num_transform times. */
MCD_API int mcd_pose_transform_matrix(int32_t index, float* h_mat6);
MCD_API int mcd_expand_transforms(const mcd_model* m, const float* d_base, int64_t N, const float* h_mats,
                          int32_t num_transform, int64_t first_item, int64_t n_items, float* d_out, void* stream);

/* ---- f1, second slice: trajectory frame rows -> dataset items, without a host-side window tensor ----------
 * mcd_normalize_frames: Trajectory._from_image_to_centre_bounding_box (utils/data.py:165-187) over
 * compute_bounding_box (utils/data.py:11-44) for F frame rows [F,34] = (x1,y1,...,x17,y17) in image
 * coordinates; (0,0) marks a missing joint.  d_out [F,34] may alias d_rows.  Bit-identical to the
 * reference's float32 numpy arithmetic (numpy >= 2 promotion rules, as pinned by the oracle).
 * With h_center / h_scale [34] (center_ / scale_ of the fitted RobustScaler; both NULL = off) the rows
 * are also scaled (utils/data.py:345-354: zeros are missing and stay zero) -- the scaler acts per
 * column, so scaling a row once equals scaling every window that contains it.
 * mcd_build_items: dataset item idx = first_item + i is transform idx / N of window idx % N
 * (utils/dataset.py:67-76); window w is rows d_win_start[w] + k * row_step, k < n_frames, of the
 * frame array (utils/preprocessing.py:55-86), laid out [2,n_frames,17] (utils/dataset.py:241-256).
 * h_center / h_scale: scale here (rows are unscaled bounding-box-centre coordinates) or NULL, NULL
 * (rows were scaled by mcd_normalize_frames).  h_mats [num_transform][6] as for
 * mcd_expand_transforms (NULL = the identity, num_transform must be 1: the base windows).
 * d_win_start [N] int64 on the device; the caller guarantees every window lies inside [0,F).
 * d_out [n_items,2,n_frames,17].
 * The three ingest calls (mcd_expand_transforms, mcd_normalize_frames, mcd_build_items) read only the handle's device and
 * n_frames: they work on a handle that has no weights yet (before mcd_model_finalize), e.g. a data loader's own handle. */
MCD_API int mcd_normalize_frames(const mcd_model* m, const float* d_rows, int64_t F, float vid_w, float vid_h,
                         const double* h_center, const double* h_scale, float* d_out, void* stream);
MCD_API int mcd_build_items(const mcd_model* m, const float* d_rows, int64_t F, const int64_t* d_win_start, int64_t N,
                    int32_t row_step, const double* h_center, const double* h_scale, const float* h_mats,
                    int32_t num_transform, int64_t first_item, int64_t n_items, float* d_out, void* stream);

/* ---- a11: per-window loss + aggregation, mocodad.py:454-520 --------------------------------
 * d_x0 [G*B,2,T,V] sample-major; d_data [B,2,n_frames,V] (the corrupt frames are the target).
 * d_losses [G,B] (may be NULL when G == 1) receives every sample's loss; d_best / d_worst [B]
 * (may be NULL) the running min / max over samples ('best' / 'worst' strategies). */
MCD_API int mcd_window_loss(const mcd_model* m, const float* d_x0, const float* d_data, int64_t B,
                    int32_t G, float* d_losses, float* d_best, float* d_worst, void* stream);

/* ---- f2: score assembly, first stage -- MoCoDAD.post_processing, models/mocodad.py:386-391 over
 * compute_var_matrix (utils/eval_utils.py:27-34) + nanmax: every window spreads its loss over the seg_len frames it covers and
 * a frame of a (transformation, scene, clip, person) row keeps the MAXIMUM over the windows that contain it; frames never covered
 * stay 0.  d_loss [N] float32; d_frames [N, seg_len] int64, 1-based frame numbers (0 wraps to the row's last frame, like the
 * reference's numpy index -1); d_row [N] int64 = the window's output row (< 0: skip the window); d_row_len [rows] int32 = frame
 * count of the row's clip; d_out [rows, stride] float32, zero-filled by the call.  Bit-identical to the reference's values
 * (a maximum has no rounding).  Needs only the handle's device (no weights).  The rest of the tail (padding around absences,
 * mean + log-range mix over persons, shift + Gaussian filter, roc_auc_score) stays on the host: mocodad_b200/postproc.py. */
MCD_API int mcd_frame_scores(const mcd_model* m, const float* d_loss, const int64_t* d_frames, const int64_t* d_row,
                     const int32_t* d_row_len, int64_t N, int32_t seg_len, int64_t rows, int64_t stride, float* d_out,
                     void* stream);

/* ---- the hot loop: the body of MoCoDAD.forward, mocodad.py:129-184 --------------------------
 * For each of G samples: x_T, then for t = N-1..1 denoise + DDPM update; then per-window losses.
 *   d_data     [B,2,n_frames,V] float32
 *   d_noise    NULL (Philox, seeded by `seed`, windows numbered from `first_window`) or the
 *              pre-drawn tensor [G, N-1, B,2,T,V]: slot 0 = x_T, slot k = z after the k-th call
 *   d_losses   [G,B] or NULL;  d_best [B] or NULL;  d_worst [B] or NULL
 *   d_x0       [G,B,2,T,V] or NULL: the generated samples (needed by the *_pose strategies)
 * The virtual batch is processed in tiles that fit `ws_bytes`. */
MCD_API int mcd_reverse_diffusion(const mcd_model* m, const float* d_data, int64_t B, int32_t G,
                          const float* d_noise, uint64_t seed, int64_t first_window,
                          float* d_losses, float* d_best, float* d_worst, float* d_x0,
                          void* d_ws, size_t ws_bytes, void* stream);

/* ---- f4: the latent variant, MoCoDADlatent.forward at stage 'diffusion' (models/mocodad_latent.py:69-132) --------------
 * (handles created with cfg.latent_dim > 0; such a handle carries the down half of the denoiser -- STSE_Unet,
 *  models/stsae/stsae_unet.py:182-246 -- plus the MLP Denoiser, models/common/components.py:203-291, and serves only
 *  mcd_cond_encode and the three calls below.)
 * mcd_latent_encode: conditioning embedding (optional output d_cond_emb [B,E]) and latent code d_code [B,latent] of the
 *   corrupt frames: STSE_Unet.forward at the constant step t = -1 (mocodad_latent.py:91-100).
 * mcd_latent_denoise: one Denoiser.forward call (components.py:264-291) on d_x [n,latent] at step t; vector v uses
 *   conditioning row v % cond_B.  d_eps [n,latent].  Parity tap.
 * mcd_latent_reverse_diffusion: the whole forward: encode, then per sample x_T and noise_steps-1 denoiser calls + DDPM updates
 *   on vectors, loss against the latent code, aggregation.
 *     d_noise  NULL (Philox) or [G, N-1, B, latent]: slot 0 = x_T, slot k = z after the k-th call
 *     d_losses [G,B] or NULL; d_best / d_worst [B] or NULL; d_x0 [G,B,latent] or NULL; d_code [B,latent] or NULL (output) */
MCD_API int mcd_latent_encode(const mcd_model* m, const float* d_data, int64_t B, float* d_cond_emb, float* d_code,
                      void* d_ws, size_t ws_bytes, void* stream);
MCD_API int mcd_latent_denoise(const mcd_model* m, const float* d_x, int64_t n, int32_t t, const float* d_cond_emb,
                       int64_t cond_B, float* d_eps, void* stream);
MCD_API int mcd_latent_reverse_diffusion(const mcd_model* m, const float* d_data, int64_t B, int32_t G, const float* d_noise,
                                 uint64_t seed, int64_t first_window, float* d_losses, float* d_best, float* d_worst,
                                 float* d_x0, float* d_code, void* d_ws, size_t ws_bytes, void* stream);

/* Same call for HOST buffers: H2D of h_data, the loop, D2H of the [B] 'best' scores, one
 * stream sync.  Allocates (and caches on the handle) its own device scratch and stream.  This is
 * the end-to-end entry bench.py times as `e2e`.  Not re-entrant on one handle. */
MCD_API int mcd_score_windows_host(mcd_model* m, const float* h_data, int64_t B, int32_t G, uint64_t seed,
                           int64_t first_window, float* h_best);

/* ---- measurement support -------------------------------------------------------------------- */
/* Number of kernel launches issued through this handle so far (bench.py's gpu_launches). */
MCD_API int64_t mcd_launch_count(const mcd_model* m);
/* Debug: arm a device-side timeline for the next launch of denoiser block `slot` (0..10): CTA 0 appends up to `cap`
 * records of 4 x int64 (warp role, pair index, event id, clock64) to d_records.  Tensor-core block kernels only. */
MCD_API int mcd_debug_trace_next(mcd_model* m, int slot, long long* d_records, int cap);
/* Per-kernel device timing: when enabled every launch through this handle is bracketed by CUDA
 * events on its stream.  mcd_profile_read synchronises the device, adds the elapsed times into
 * per-kernel totals and returns them: for kernel slot i (0 <= i < mcd_profile_slots()),
 * ms[i] = total milliseconds, launches[i] = count, windows[i] = windows processed.  Totals are
 * cleared by mcd_profile_enable(m, 1). */
MCD_API int mcd_profile_slots(void);
MCD_API const char* mcd_profile_slot_name(int slot);
MCD_API int mcd_profile_enable(mcd_model* m, int on);
MCD_API int mcd_profile_read(mcd_model* m, double* ms, int64_t* launches, int64_t* windows);
/* Algorithmic cost of one window through kernel slot i (DESIGN.md "roofline"): bytes that must
 * cross HBM (activations in + out, fp32; weights excluded) and fp32 FLOPs (2 per FMA). */
MCD_API int mcd_profile_slot_cost(const mcd_model* m, int slot, double* bytes_per_window,
                          double* flops_per_window);
/* fp32 FMA throughput of the device (TFLOP/s) from a register-resident FMA loop; the fp32
 * roofline denominator bench.py reports next to the HBM one (MEASURED_PEAKS.json has none). */
MCD_API int mcd_probe_fp32_tflops(int32_t device, double* tflops);
/* The two instruction forms separately: scalar FFMA and packed FFMA2 (mcd_probe_fp32_tflops = max). */
MCD_API int mcd_probe_fp32_detail(int32_t device, double* ffma_tflops, double* ffma2_tflops);

#ifdef __cplusplus
}
#endif
#endif /* MOCODAD_B200_H */
