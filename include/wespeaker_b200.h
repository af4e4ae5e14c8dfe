/* wespeaker_b200 — C ABI of the B200-native speaker-embedding extraction + PLDA scoring engine.
 *
 * The reference (wenet-e2e/wespeaker) has no FFI layer for this path; the seams this library sits behind are
 * Python call signatures and one C++ virtual.  Each entry point below cites the reference
 * interface it replaces.  Conventions: plain pointers and sizes, no torch types; every function returns 0 on
 * success and non-zero on failure with a message in ws_last_error() (thread-local); no exceptions cross the ABI.
 * "dev" pointers are CUDA device pointers owned by the caller (e.g. torch tensors' data_ptr()); `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Handles are bound to one device and are not
 * internally thread-safe (same contract as the reference's single-threaded callers, extract.py:109-139).
 */
#ifndef WESPEAKER_B200_H_
#define WESPEAKER_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ws_engine ws_engine;
typedef struct ws_plda ws_plda;

int ws_version(void);
const char* ws_last_error(void);

/* ---- model engine: replaces `get_speaker_model(name)(**model_args)` + `load_checkpoint` + `model(features)`
 *      (wespeaker/models/speaker_model.py:31-62, wespeaker/utils/checkpoint.py:20-85, wespeaker/bin/extract.py:68-79,133)
 *      and is the B200 back-end for `SpeakerModel::ExtractEmbedding` (runtime/core/speaker/speaker_model.h:25-32).
 * model_name: ECAPA_TDNN_c512 | ECAPA_TDNN_GLOB_c512 | ECAPA_TDNN_c1024 | ECAPA_TDNN_GLOB_c1024 | ResNet18 |
 *             ResNet34 | CAMPPlus; SURVEY section 8(f) rank 4: ResNet50 | ResNet101 | ResNet152 | ResNet221 | ResNet293 |
 *             XVEC | Res2Net34_Base | Res2Net34_Large | ERes2Net34_Base | ERes2Net34_Large | ERes2Net34_aug.
 * precision:  "fp32" (exact IEEE fp32 FFMA path, parity <= 1e-4), "tf32x3" (tcgen05 with 3xTF32 error compensation,
 *             fp32-grade accuracy), "tf32", "bf16", "fp16" (tcgen05 tensor cores). */
int ws_engine_create(const char* model_name, const char* precision, int feat_dim, int embed_dim, int device,
                     ws_engine** out);
/* Plan-check engine: needs NO device and never computes.  set_option / set_tensor / finalize / plan_op_name behave as on a
 * real engine - the launch plan is built with placeholder addresses and every tensor map is checked against the
 * cuTensorMapEncodeTiled rules - so a checkpoint (`load_checkpoint`, wespeaker/utils/checkpoint.py:20-85) can be validated
 * against the plan builder (missing keys, shapes, kernel envelopes) on a host without a GPU.  Every compute entry point
 * returns an error on such an engine. */
int ws_engine_create_plan_check(const char* model_name, const char* precision, int feat_dim, int embed_dim, ws_engine** out);
/* options: "two_emb_layer", "emb_bn" (model_args), "cuda_graph" (default 1), "force_simt" (debug cross-check),
 * "tc_version" (1: one-tile-per-CTA tcgen05 kernel, 2: persistent / TMA-store kernel, 3 (default): + cta_group::2 CTA pairs
 * on the large layers),
 * "res2_fused" (default 1: ECAPA Res2 chains run as one fused persistent kernel per stage for 16-bit precisions) */
int ws_engine_set_option(ws_engine* e, const char* key, long long value);
/* one reference state_dict entry (fp32 host data, reference key names). */
int ws_engine_set_tensor(ws_engine* e, const char* key, const float* host_data, const long long* shape, int ndim);
/* packs / folds / converts weights once; errors if a required key is missing (strict for the forward path). */
int ws_engine_finalize(ws_engine* e);
int ws_engine_embed_dim(const ws_engine* e);
/* feats_dev: fp32 (B,T,feat_dim) already mean-normalised by the caller, as `model(features)` receives them
 * (extract.py:125-133); embs_dev: fp32 (B,embed_dim). */
int ws_engine_forward(ws_engine* e, const float* feats_dev, int B, int T, float* embs_dev, void* stream);
/* Length-masked batches (utterances of different lengths padded to a common T in ONE batch; the reference has no masking and
 * runs such sets at batch 1, extract_vox.sh:31): the result of every utterance equals what it produces alone - rows behind
 * its end stay zero wherever a convolution reads across time (the zero padding its own forward sees), SE / ASTP / TSTP /
 * CAM++ context statistics use its own frame count, CAM++'s ceil-mode segment pooling its own segment count
 * (campplus.py:117-135, pooling_layers.py:78-85,119-144).  n_frames_dev / n_samples_dev: int32[B] on the device.
 * ECAPA, ResNet and CAM++ (16-bit precisions for CAM++); XVEC is refused. */
int ws_engine_forward_masked(ws_engine* e, const float* feats_dev, const int* n_frames_dev, int B, int T, float* embs_dev,
                             void* stream);
int ws_engine_extract_wav_masked(ws_engine* e, const void* wav_dev, int wav_is_i16, long long wav_ld, const int* n_samples_dev,
                                 int max_samples, int B, const char* window_type, float* embs_dev, void* stream);
/* The same from a RAGGED buffer: utterance b is the n_samples_dev[b] samples starting at wav_dev + offsets_dev[b] samples
 * (int64[B] on the device) - concatenated PCM as a reader produces it, no padded copy.  max_samples >= every n_samples[b]
 * (it fixes the plan's T); frames behind an utterance's end are neither read nor computed.  Every utterance must hold at
 * least one 25 ms frame (400 samples; the reference's kaldi fbank fails below that too).  _async: see below. */
int ws_engine_extract_wav_ragged(ws_engine* e, const void* wav_dev, int wav_is_i16, const long long* offsets_dev,
                                 const int* n_samples_dev, int max_samples, int B, const char* window_type, float* embs_dev,
                                 void* stream);
int ws_engine_extract_wav_ragged_async(ws_engine* e, const void* wav_dev, int wav_is_i16, const long long* offsets_dev,
                                       const int* n_samples_dev, int max_samples, int B, const char* window_type,
                                       float* embs_dev, void* stream);

/* Variable-length jobs: one plan (and CUDA graph) per distinct (B, T); plans own disjoint buffers and are spread over a
 * few internal streams, so the buckets of such a job overlap on the GPU when they are enqueued with the *_async variants
 * (same arguments and semantics as ws_engine_forward / ws_engine_extract_wav, but `stream` is NOT made to wait for the
 * result) and closed with ONE ws_engine_join(e, stream).  Inputs are ordered behind `stream` as usual; outputs and inputs
 * must stay alive and untouched until the join.  The reference runs such sets at batch 1 (extract_vox.sh:31). */
int ws_engine_forward_async(ws_engine* e, const float* feats_dev, int B, int T, float* embs_dev, void* stream);
int ws_engine_extract_wav_async(ws_engine* e, const void* wav_dev, int wav_is_i16, long long wav_ld, int nsamples, int B,
                                const char* window_type, float* embs_dev, void* stream);
int ws_engine_join(ws_engine* e, void* stream);

/* same with HOST buffers (pinned or pageable): H2D + forward + D2H + stream sync; mirrors
 * `features.to(device)` ... `embeds.cpu()` (extract.py:116,135). */
int ws_engine_forward_host(ws_engine* e, const float* feats_host, int B, int T, float* embs_host);
/* waveform -> embedding on device: fbank (80 bins, 25/10 ms, dither 0) + CMN + forward.  Replaces
 * compute_fbank + apply_cmvn + model() (processor.py:496-526, dataset_utils.py:19-26, cli/speaker.py:130-167).
 * wav: (B, wav_ld) samples in int16 range — float32 (`wav * (1<<15)`) or int16 PCM; feats_out_dev optional. */
int ws_engine_extract_wav(ws_engine* e, const void* wav_dev, int wav_is_i16, long long wav_ld, int nsamples, int B,
                          const char* window_type, float* embs_dev, float* feats_out_dev, void* stream);
int ws_engine_extract_wav_host(ws_engine* e, const void* wav_host, int wav_is_i16, int nsamples, int B,
                               const char* window_type, float* embs_host);
/* Pipelined variant of ws_engine_extract_wav_host: submit() enqueues H2D (copy stream) + fbank + CMN + forward + D2H
 * for `slot` (0..3) and returns; collect() blocks until that slot's embs_host is filled.  Cycling through the slots
 * overlaps the H2D copies of the next batches with the kernels of batch i and rides out host scheduling jitter — the role DataLoader workers / prefetch_factor play
 * in extract.py:99-103.  Host buffers should be pinned and must stay valid until collect(). */
int ws_engine_submit_wav_host(ws_engine* e, int slot, const void* wav_host, int wav_is_i16, int nsamples, int B,
                              const char* window_type, float* embs_host);
int ws_engine_collect(ws_engine* e, int slot);
/* tuning aid: per-op device time (ms) of the (B,T) plan, measured with CUDA events in sequence context; returns #ops */
int ws_engine_profile_ops(ws_engine* e, int B, int T, int iters, float* ms_out, int max_ops);
const char* ws_engine_plan_op_name(ws_engine* e, int B, int T, int i, double* flops_out);   /* label + FLOPs of op i */
/* plan-check engines only: write the (B,T) launch plan as data (JSON op descriptions + placeholder allocation table + fp32
 * weight sources) so that a test can re-evaluate the plan's arithmetic on the host (tests/plan_interp.py) */
int ws_engine_plan_trace(ws_engine* e, int B, int T, int masked, const char* path);   /* masked: the length-masked plan of (B,T) */
/* number of this library's kernels launched by the most recent forward/extract call */
long long ws_engine_last_launches(const ws_engine* e);
void ws_engine_destroy(ws_engine* e);

/* ---- fbank + CMN: replaces torchaudio.compliance.kaldi.fbank as called at processor.py:518-525 /
 *      cli/speaker.py:92-99 and Fbank::Compute (runtime/core/frontend/fbank.h:138-198). */
int ws_fbank_num_frames(int nsamples);
/* Sinc resampling of (B, n_in) waveforms (int16 or float32) to new_freq: torchaudio.transforms.Resample(orig_freq, new_freq)
 * with its defaults, as the reference applies it before fbank (dataset/processor.py:242-262, cli/speaker.py:157-159).
 * out_dev: fp32 (B, out_ld), the first ws_resample_out_len(n_in, orig, new) = ceil(new * n_in / orig) samples of each row. */
int ws_resample_out_len(int n_in, int orig_freq, int new_freq);
int ws_resample(const void* wav_dev, int wav_is_i16, long long wav_ld, int n_in, int B, int orig_freq, int new_freq,
                float* out_dev, long long out_ld, void* stream);
int ws_fbank(const void* wav_dev, int wav_is_i16, long long wav_ld, int nsamples, int B, const char* window_type,
             int apply_cmn, float* feats_dev, void* stream);

/* ---- generic fused conv operator (channels-last), the building block of the engine; exported for parity tests
 *      against torch conv1d/conv2d.  Replaces Conv1dReluBn (ecapa_tdnn.py:85-106), BasicBlock convs
 *      (resnet.py:35-69), TDNNLayer (campplus.py:55-83). */
typedef struct {
    const void* x;            /* [B][F][T][x_ld] activations, dtype below, channels [0,Cin) used */
    int B, F, T, Cin;
    long long x_ld;
    const void* w;            /* [Cout][kf*kt*Cin], tap-major (tap = jf*kt + jt), dtype below */
    int Cout, kf, kt, dil_f, dil_t, pad_f, pad_t, stride_f, stride_t;
    const float* bias;        /* added before act1 (or NULL) */
    int act1;                 /* 0 none 1 relu 2 tanh 3 sigmoid 4 hardtanh(0,20) 5 silu */
    const float* scale;       /* per-channel affine after act1 (or NULL) */
    const float* shift;
    const void* res;          /* residual added after the affine (or NULL), [positions][res_ld] */
    long long res_ld;
    int act2;
    void* out;                /* [B][Fo][To][out_ld] */
    long long out_ld;
    int dtype;                /* 0 fp32 (tf32 MMA when use_tc), 1 bf16, 2 fp16 */
    int use_tc;               /* 0: fp32 FFMA kernel, 1: tcgen05 v1, 2: persistent tcgen05 v2, 3: v2 with cta_group::2 pairs */
    /* 3xTF32 (dtype 0, use_tc 2): low parts v - tf32_trunc(v) of x and w (inputs) and of out (written); all NULL = off */
    const void* x_lo;
    const void* w_lo;
    void* out_lo;
    /* fused SE squeeze (dense 1x1 conv, 16-bit dtype, use_tc >= 2, T >= 128; NULL = off): fp32 [2*ceil(B*T/64)][Cout];
     * row 2u+s = per-channel sum of the stored outputs over the positions of 64-position unit u that belong to
     * utterance floor(64u/T)+s.  Replaces the mean pass of SE_Connect (ecapa_tdnn.py:120-121). */
    float* colsum;
} ws_conv_desc;
int ws_conv(const ws_conv_desc* d, void* stream);

/* ---- two-covariance PLDA scorer: replaces TwoCovPLDA.transform_embedding / log_likelihood_ratio / the trial
 *      loop of eval_sv (wespeaker/utils/plda/two_cov_plda.py:156-184,186-256), fp64 like the reference. */
int ws_plda_create(int dim, const double* mu, const double* transform, const double* psi, const double* offset,
                   int normalize_length, int device, ws_plda** out);
/* y = transform_embedding(pre(x - mean_vec)); pre = sqrt(D)-length-norm (norm_embeddings, plda_utils.py:46-58) iff
 * pre_norm != 0 — eval_sv (:225-241) pre-normalises exactly when normalize_length is set.
 * x_dev fp32 (N,D); mean_vec_host fp64 (D) or NULL; y_dev fp64 (N,D). */
int ws_plda_transform(ws_plda* p, const float* x_dev, long long N, const double* mean_vec_host, int pre_norm,
                      double* y_dev, void* stream);
/* same with fp64 rows that are already mean-subtracted (the per-speaker session means of eval_sv :218-233 stay fp64) */
int ws_plda_transform64(ws_plda* p, const double* x64_dev, long long N, int pre_norm, double* y_dev, void* stream);
/* all-pairs LLR: out[i*out_ld + j] = log_likelihood_ratio(enroll_t[i], test_t[j], n_i);  n_i = counts_dev[i] or const_n */
int ws_plda_score_matrix(ws_plda* p, const double* enroll_t_dev, const int* counts_dev, int const_n, long long N,
                         const double* test_t_dev, long long M, void* out_dev, int out_is_f64, long long out_ld,
                         void* stream);
/* listed trials only (eval_sv :248-256): out[k] = LLR(enroll_t[ei[k]], test_t[ti[k]], n_{ei[k]}) */
int ws_plda_score_trials(ws_plda* p, const double* enroll_t_dev, const int* counts_dev, int const_n, long long N,
                         const double* test_t_dev, long long M, const long long* ei_dev, const long long* ti_dev,
                         long long ntrials, double* out_dev, void* stream);
void ws_plda_destroy(ws_plda* p);

/* ---- cosine scoring + S-norm / AS-norm: replaces trials_cosine_score (wespeaker/bin/score.py:38-72),
 *      get_mean_std (wespeaker/bin/score_norm.py:26-37) and the per-trial normalisation (score_norm.py:102-107).
 *      All pointers are device pointers; fp32 embeddings in, fp64 arithmetic. */
/* unit[r] = (x[r] - mean_vec) / |x[r] - mean_vec| (fp64, N x D); norms[r] = |x[r] - mean_vec| (the "mag" columns of
 * score_norm.py:109-110); mean_vec_dev (D fp64) and norms_dev may be NULL. */
int ws_score_unit_rows(const float* x_dev, long long N, int D, const double* mean_vec_dev, double* unit_dev,
                       double* norms_dev, void* stream);
/* out[k] = <unit[enroll_idx[k]], unit[test_idx[k]]>: cosine_similarity of the listed trials (score.py:62-63) */
int ws_score_cosine_trials(const double* unit_dev, const long long* enroll_idx_dev, const long long* test_idx_dev,
                           long long ntrials, int D, double* out_dev, void* stream);
/* mean / population std of the top_n largest cosine scores of each embedding against the M cohort rows
 * (get_mean_std; top_n >= M is S-norm).  work_dev: fp32 scratch of work_rows x M score tiles. */
int ws_score_cohort_stats(const double* unit_emb_dev, long long N, const double* unit_cohort_dev, long long M, int D,
                          int top_n, float* work_dev, long long work_rows, double* mean_dev, double* std_dev,
                          void* stream);
/* out[k] = 0.5 * ((s[k] - emean[ei[k]]) / estd[ei[k]] + (s[k] - tmean[ti[k]]) / tstd[ti[k]]) */
int ws_score_asnorm(const double* scores_dev, const long long* enroll_idx_dev, const long long* test_idx_dev,
                    long long ntrials, const double* enroll_mean_dev, const double* enroll_std_dev,
                    const double* test_mean_dev, const double* test_std_dev, double* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WESPEAKER_B200_H_ */
