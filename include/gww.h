/* gww.h -- C ABI of the B200-native GW-Whisper sliding-window inference path.
 *
 * Drop-in boundary: the reference has no FFI; its boundary is a set of Python
 * signatures.  Every entry point below replaces one stage of that Python path and is what a
 * reference-side binding (ctypes / torch C++ shim) calls.  All pointers marked
 * "device" are CUDA device pointers on the current device; `stream` is a cudaStream_t passed as
 * void*.  No entry point allocates device memory on the hot path: callers query
 * gww_workspace_bytes() and pass a workspace (e.g. from the torch caching allocator).
 *
 * Return value: 0 on success, non-zero error code otherwise; gww_last_error() returns a
 * thread-local human-readable message.  There is NO CPU fallback: on a machine without an sm_100
 * device every compute entry point fails with GWW_ERR_NO_DEVICE.
 *
 * Ordering of det-windows: a batch of B windows with D detectors is laid out [B, D, ...]
 * (window-major, detector-minor), so pooled representations [B*D, d] reinterpret as the
 * reference's torch.cat(reps, dim=1) [B, D*d] (MLGWSC-1/inference.py:391).
 */
#ifndef GWW_H_
#define GWW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWW_OK 0
#define GWW_ERR_INVALID 1
#define GWW_ERR_CUDA 2
#define GWW_ERR_NO_DEVICE 3
#define GWW_ERR_WORKSPACE 4

#define GWW_T_SAMPLES 2048   /* samples per window (1 s @ 2048 Hz)                               */
#define GWW_N_MELS 80        /* feature rows, HF feature_extraction_whisper.py:70                 */
#define GWW_N_FRAMES 3000    /* feature columns (30 s / 10 ms), checked at modeling_whisper.py:613 */
#define GWW_N_CTX 1500       /* encoder tokens after the stride-2 conv                            */
#define GWW_MAX_HEAD_LAYERS 6

typedef struct gww_model gww_model_t;

/* Whisper encoder geometry (HF configuration_whisper.py): tiny 384/4/6/1536, base 512/6/8/2048,
 * small 768/12/12/3072.  head_dim must be 64. */
typedef struct {
  int d_model;
  int n_layers;
  int n_heads;
  int ffn_dim;
} gww_encoder_config_t;

/* PEFT-0.12 DoRA adapter of one Linear (reference: Signal_vs_Noise/results/.../adapter_config.json,
 * r=8, lora_alpha=32, use_dora=true).  All NULL => projection not adapted.  Merged at model creation:
 *   W' = diag(m / ||W0 + s B A||_row) (W0 + s B A),  s = alpha / r        */
typedef struct {
  const float* lora_A;     /* host [r, d_in]  */
  const float* lora_B;     /* host [d_out, r] */
  const float* magnitude;  /* host [d_out]    */
  int r;
  float scale;             /* alpha / r */
} gww_dora_t;

/* One WhisperEncoderLayer (HF modeling_whisper.py:361-415).  Host pointers, fp32, nn.Linear layout
 * [out, in].  k_proj has no bias (:279). */
typedef struct {
  const float *ln1_g, *ln1_b;
  const float *q_w, *q_b, *k_w, *v_w, *v_b, *o_w, *o_b;
  const float *ln2_g, *ln2_b;
  const float *fc1_w, *fc1_b, *fc2_w, *fc2_b;
  gww_dora_t dora_q, dora_k, dora_v, dora_o;
} gww_layer_weights_t;

typedef struct {
  const float* conv1_w;   /* host [d, 80, 3]  (modeling_whisper.py:558) */
  const float* conv1_b;   /* [d] */
  const float* conv2_w;   /* host [d, d, 3], stride 2 (:559) */
  const float* conv2_b;   /* [d] */
  const float* pos_emb;   /* [1500, d] embed_positions (:561) */
  const float *ln_post_g, *ln_post_b; /* final layer_norm (:643) */
  const gww_layer_weights_t* layers;  /* [n_layers] */
} gww_encoder_weights_t;

/* Pooled MLP classifier: Linear+ReLU ... Linear (+Softmax).  Mirrors the nn.Sequential heads at
 * MLGWSC-1/inference.py:371-382, Signal_vs_Noise/src/model.py:9-20,35-47,
 * Glitch_classification/src/model.py:10-21 (Dropout is inert in eval). */
typedef struct {
  int n_layers;
  int dims[GWW_MAX_HEAD_LAYERS + 1]; /* dims[0] = d_model * n_detectors */
  const float* w[GWW_MAX_HEAD_LAYERS]; /* host [dims[i+1], dims[i]] */
  const float* b[GWW_MAX_HEAD_LAYERS]; /* host [dims[i+1]] */
  int softmax;                         /* 1 = keep the trailing nn.Softmax(dim=1) */
} gww_head_weights_t;

/* ---- library / device ------------------------------------------------------------------------ */
const char* gww_last_error(void);
const char* gww_version(void);
/* "f16" or "bf16": the 16-bit tensor-core operand type this library was built with (activations and weights fed
 * to tcgen05.mma; accumulation is fp32 either way).  libgww_b200.so is the fp16 build (Whisper is an fp16
 * model; 8x less operand rounding than bf16 against the fp32 reference), libgww_b200_bf16.so the bf16 one.
 * The building-block entry points below that take "16-bit" device buffers use this type. */
const char* gww_operand_dtype(void);
/* 0 if the current device is sm_100 and the kernels can run, else GWW_ERR_NO_DEVICE. */
int gww_device_ok(void);

/* ---- model handle ---------------------------------------------------------------------------- */
/* Uploads weights, merges DoRA, casts matrices to the 16-bit operand type (fp32 merge), folds the
 * head_dim^-0.5 query scale into W_q/b_q.  Immutable afterwards; one in-flight call per workspace. */
int gww_model_create(const gww_encoder_config_t* cfg, const gww_encoder_weights_t* w,
                     gww_model_t** out);
int gww_model_set_head(gww_model_t* m, const gww_head_weights_t* head);
void gww_model_destroy(gww_model_t* m);
/* LayerNorm folding (the LayerNorms are evaluated inside the neighbouring GEMMs on the raw residual stream) loses
 * precision when residual rows have a large common mode.  The library watches max |mean|/std over the rows it
 * normalises and switches a model to the stand-alone LayerNorm kernel when it exceeds 4 (fp16 operands; 1.5 for
 * bf16): the first chunk is checked synchronously and recomputed if needed, later ones asynchronously.  This call
 * (synchronising) reports the maximum seen and whether folding is still active.  GWW_LN_FOLD=0 disables folding. */
int gww_model_ln_fold_state(const gww_model_t* m, float* max_ratio, int* fold_active);
/* Device bytes needed by forward calls that process up to `chunk` det-windows at a time. */
size_t gww_workspace_bytes(const gww_model_t* m, int chunk);

/* ---- front end A (replaces scipy.signal.resample + WhisperFeatureExtractor) -------------------- */
/* strain: device f32 [n, 2048]; feats: device f32 [n, 80, 3000] (reference layout/dtype). */
int gww_logmel_frontend(const float* strain, long n, float* feats, void* stream);

/* The two halves of front end A on their own, for callers that keep the reference's on-disk format (the
 * reference stores the RESAMPLED 16 kHz audio, Signal_vs_Noise/utils/preprocess.py:44-51,94-98, and feeds it to
 * WhisperFeatureExtractor per item, Signal_vs_Noise/src/dataset.py:20-24):
 *   gww_resample_16k     == scipy.signal.resample(x, 16000) stored as f32: strain device f32 [n, 2048] ->
 *                           audio device f32 [n, 16000] (16-byte aligned)
 *   gww_logmel_from_16k  == WhisperFeatureExtractor(audio, sampling_rate=16000).input_features:
 *                           audio device f32 [n, 16000] -> feats device f32 [n, 80, 3000] */
int gww_resample_16k(const float* strain, long n, float* audio, void* stream);
int gww_logmel_from_16k(const float* audio, long n, float* feats, void* stream);

/* ---- encoder (replaces HF WhisperEncoder.forward with merged DoRA) ----------------------------- */
/* feats: device f32 [n, 80, 3000].  Exactly one of the outputs may be NULL:
 *   last_hidden: device f32 [n, 1500, d]  (== encoder(feats).last_hidden_state)
 *   pooled:      device f32 [n, d]        (== last_hidden_state[:, -1, :] if use_last_token
 *                                             else .mean(dim=1); inference.py:390) */
int gww_encoder_forward(const gww_model_t* m, const float* feats, long n, float* last_hidden,
                        float* pooled, int use_last_token, void* workspace, size_t workspace_bytes,
                        int chunk, void* stream);

/* ---- head (replaces the nn.Sequential classifier) ---------------------------------------------- */
/* reps: device f32 [B, dims[0]]; out: device f32 [B, C]. */
int gww_head_forward(const gww_model_t* m, const float* reps, long B, float* out, void* stream);

/* ---- fused window path: strain -> log-mel -> encoder -> pooled -> head ------------------------- */
/* strain: device f32 [B, D, 2048]; out: device f32 [B, C].  Equivalent to
 * two_channel_ligo_binary_classifier.forward / one_channel_... on the log-mel features of each
 * detector (Signal_vs_Noise/src/model.py:22-29,49-52).  pooled_out (optional, may be NULL):
 * device f32 [B*D, d]. */
int gww_forward_windows_logmel(const gww_model_t* m, const float* strain, long B, int D, float* out,
                               float* pooled_out, void* workspace, size_t workspace_bytes,
                               int chunk, void* stream);

/* ---- sliding-window search over a resident strain segment -------------------------------------- */
/* strain: device f32 [D, n_samples] (whitened).  Window k covers samples [k*hop, k*hop+2048)
 * (SegmentSlicer, MLGWSC-1/inference.py:198-199,254-262).  scores: device f32 [n_windows]
 * (= out[:, 0], inference.py:481).  Triggers (score > thr, strictly; :484) are appended in window
 * order to trig_idx/trig_score (device, capacity entries); *trig_count (device int, must be
 * zeroed by the caller) receives the running count. n_windows = 1 + (n_samples-2048)/hop. */
int gww_stream_search_logmel(const gww_model_t* m, const float* strain, int D, long n_samples,
                             int hop, long first_window, long n_windows, float thr, float* scores,
                             long* trig_idx, float* trig_score, int* trig_count, int capacity,
                             void* workspace, size_t workspace_bytes, int chunk, void* stream);

/* Ordered threshold compaction on its own (K12): out device f32 [n, C], column 0 is the score. */
int gww_threshold_compact(const float* out, int C, long n, float thr, long idx_base, long* trig_idx,
                          float* trig_score, int* trig_count, int capacity, void* stream);

/* ---- front end B: Q-transform + Q-Adapter (MLGWSC-1) --------------------------------------------- */
/* Replaces QTransformAdapter (MLGWSC-1/inference.py:303-351) including ml4gw.transforms.QScan
 * (:316-321, :345).  The handle holds the (q, f) tiling plan, bisquare windows and, once set, the
 * adapter CNN weights.  Only duration * sample_rate == 2048 is supported (1 s @ 2048 Hz); the spectrogram sides
 * are multiples of 64 in [64, 512] ([512,512] at inference.py:310, [128,128] at train.py:104). */
typedef struct gww_qfront gww_qfront_t;

typedef struct {
  const float *conv1_w, *conv1_b;   /* host [16,1,3,3], [16]   freq_adapter.0 */
  const float *conv2_w, *conv2_b;   /* host [32,16,3,3], [32]  freq_adapter.3 */
  const float *conv3_w, *conv3_b;   /* host [64,32,3,3], [64]  freq_adapter.6 */
  const float *conv4_w, *conv4_b;   /* host [1,64,1,1], [1]    freq_adapter.8 */
  float scale, bias;                /* QTransformAdapter.scale / .bias (inference.py:334-335) */
  int n_detectors;                  /* <= 8 */
  const float *film_gamma, *film_beta; /* host [n_detectors] (:336-337) */
  int c1, c2, c3;                   /* CNN widths; 0 = the inference defaults 16 / 32 / 64.  MLGWSC-1/train.py:118-123
                                     * trains a 32 / 64 / 128 adapter on a 128x128 Q-spectrogram (:104): conv weights
                                     * are then [c1,1,3,3], [c2,c1,3,3], [c3,c2,3,3], [1,c3,1,1] */
} gww_qadapter_weights_t;

int gww_qfront_create(double duration, double sample_rate, double qmin, double qmax, double mismatch,
                      int spec_f, int spec_t, gww_qfront_t** out);
int gww_qfront_set_adapter(gww_qfront_t* qf, const gww_qadapter_weights_t* w);
void gww_qfront_destroy(gww_qfront_t* qf);
/* Tiling plan, for tests: counts, then per-row arrays (caller allocates n_rows entries, any may be
 * NULL).  Rows are in plane-major, ascending-frequency order. */
int gww_qfront_info(const gww_qfront_t* qf, int* n_planes, int* n_rows, int* n_tiles);
int gww_qfront_plan(const gww_qfront_t* qf, double* q_of_plane, int* plane_of_row, float* freq,
                    int* ntiles, int* windowsize, int* tile_offset);
/* Device bytes needed by gww_qscan / gww_qadapter / gww_forward_windows_qscan on n windows. */
size_t gww_qfront_workspace_bytes(const gww_qfront_t* qf, long n);
/* QScan.forward on ONE call of n windows (the plane choice is coupled across the n windows, as in
 * ml4gw).  strain: device f32, window w at strain + w*win_stride (win_stride >= 2048 or a hop for
 * overlapping windows).  spec: device f32 [n, spec_f, spec_t].  tiles (optional, may be NULL): device
 * f32 [n, n_tiles] normalised tile energies of all planes.  plane_idx (optional): device int. */
int gww_qscan(const gww_qfront_t* qf, const float* strain, long n, long win_stride, float* spec,
              float* tiles, int* plane_idx, void* workspace, size_t workspace_bytes, void* stream);
/* freq_adapter + final_pool + scale/bias + FiLM[det_idx] on spec [n, spec_f, spec_t].
 * feats_f32 (optional): device f32 [n, 80, 3000]. */
int gww_qadapter(const gww_qfront_t* qf, const float* spec, long n, int det_idx, float* feats_f32,
                 void* workspace, size_t workspace_bytes, void* stream);
/* GWWhisperClassifier.forward (inference.py:384-392): strain device f32 [B, D, 2048] -> out [B, C].
 * The B windows form one QScan call per detector (pass the reference's 256-window batches).
 * workspace: gww_workspace_bytes(m, B*D); q_workspace: gww_qfront_workspace_bytes(qf, B). */
int gww_forward_windows_qscan(const gww_model_t* m, const gww_qfront_t* qf, const float* strain, long B,
                              int D, int use_last_token, float* out, void* workspace,
                              size_t workspace_bytes, void* q_workspace, size_t q_workspace_bytes,
                              void* stream);
/* Sliding-window MLGWSC-1 search over a resident whitened segment strain [D, n_samples]: batches of
 * `batch` consecutive windows (256 in the reference, inference.py:465), score = out[:, 0]. */
int gww_stream_search_qscan(const gww_model_t* m, const gww_qfront_t* qf, const float* strain, int D,
                            long n_samples, int hop, long first_window, long n_windows, int batch,
                            float thr, float* scores, long* trig_idx, float* trig_score,
                            int* trig_count, int capacity, void* workspace, size_t workspace_bytes,
                            void* q_workspace, size_t q_workspace_bytes, void* stream);

/* ---- whitening (replaces `whiten`, MLGWSC-1/inference.py:56-137, psd=None branch; pycbc 2.4.0 semantics) ----- */
/* strain: device f64 [n] (one detector, n even).  Welch-median PSD over segments of seg_len samples every
 * seg_stride (TimeSeries.psd(segment_duration): seg_len = round(0.5 s * fs) = 1024, stride seg_len/2), interpolated
 * to the segment's resolution, inverse-spectrum truncation to max_filter_len taps (int(0.25 s * fs) = 512) with a
 * Hann window (trunc_hann) above low_frequency_cutoff (<= 0: none), strain filtered with psd_trunc^-1/2 and -- if
 * remove_corrupted -- cropped by max_filter_len/2 samples at both ends.
 * Outputs (device; white and white_f32 optional, at least one): white f64 / white_f32 f32 [n - max_filter_len]
 * (or [n]); psd_out (optional) f64 [seg_len/2 + 1] = the un-interpolated Welch PSD (`return_psd`).
 * fir_half: half-length of the time-domain filter that applies psd_trunc^-1/2 (0 = default 8192, see whiten.cuh). */
size_t gww_whiten_workspace_bytes(long n, int seg_len, int seg_stride, int max_filter_len, int fir_half);
int gww_whiten(const double* strain, long n, double delta_t, int seg_len, int seg_stride, int max_filter_len,
               double low_frequency_cutoff, int trunc_hann, int remove_corrupted, int fir_half, double* white,
               float* white_f32, double* psd_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- building blocks exported for parity tests and profiling ----------------------------------- */
/* C[M,N] = epilogue(A[M,K] * W[N,K]^T); A, W device 16-bit (gww_operand_dtype()) row-major; epilogue ids as in
 * gemm_tc.cuh (0 bias->16-bit, 1 bias+gelu->16-bit, 2 bias+resid->f32, 3 bias+gelu+pos->f32).  (The name keeps
 * its round-1 spelling; "bf16" here means "the 16-bit operand type of the build".) */
int gww_gemm_bf16(const void* A, const void* W, void* C, const float* bias, const float* resid,
                  const float* pos, long M, int N, int K, int epilogue, int block_n, void* stream);
/* qkv device 16-bit [n, T, 3d] -> out device 16-bit [n, T, d] */
int gww_attention(const void* qkv, void* out, long n, int T, int d_model, void* stream);
/* x device f32 [rows, d] -> out 16-bit operand type (out_bf16=1) or f32 [rows, d] */
int gww_layernorm(const float* x, void* out, const float* gamma, const float* beta, long rows, int d,
                  int out_bf16, void* stream);
/* The reference consumes only last_hidden_state[:, -1, :]; by default the final encoder layer is
 * therefore evaluated for the last token only (all tokens' K/V, one query row; exact up to summation
 * order) whenever gww_encoder_forward is asked for the last-token `pooled` output
 * without `last_hidden`.  enable=0 forces the full 1500-token final layer.  Returns the old value. */
int gww_set_last_layer_pruning(int enable);
/* number of kernel launches issued by this library in this process (bench.py gpu_launches) */
long gww_launch_count(void);
/* Optional per-kernel-class timing: between begin/end every launch is bracketed by CUDA events on
 * its own stream; end() waits for them and returns total milliseconds and launch counts per class
 * (arrays of gww_profile_num_kinds() entries).  Used by bench.py for the live roofline numbers. */
int gww_profile_num_kinds(void);
const char* gww_profile_kind_name(int kind);
int gww_profile_begin(void);
int gww_profile_end(double* ms_by_kind, long* count_by_kind);

#ifdef __cplusplus
}
#endif
#endif /* GWW_H_ */
