/* hirest_b200.h — C ABI of libhirest_b200.so (B200 / sm_100a only).
 *
 * The reference (j-min/HiREST) has no FFI layer: its hot path is reached through Python methods.  This
 * header is the boundary a maintainer binds underneath those methods (ctypes stub in INTEGRATION.md):
 *
 *   EVA_CLIP.encode_image   EVA_clip/eva_model.py:317-318 -> vit_model.py:326-351   => hb_vit_encode
 *   EVA_CLIP.encode_text    EVA_clip/eva_model.py:320-321 -> :232-250               => hb_text_encode
 *   mean-pool + L2 norm     inference_video_retrieval.py:210-212, 283-285, 323-326  => hb_pool_normalize
 *   T @ V.T scoring         inference_video_retrieval.py:334                        => hb_similarity
 *   nn.Linear / F.linear    (any site listed in SURVEY.md §8(a))                    => hb_linear
 *
 * Conventions: every function returns 0 on success or a negative code (hb_strerror); nothing throws across
 * the ABI.  All data pointers are DEVICE pointers owned by the caller (e.g. torch tensors' data_ptr());
 * `stream` is a cudaStream_t passed as void*.  Handles own only repacked bf16 weights + workspace.  No hidden
 * synchronisation: results are ready when `stream` reaches the call's last kernel.  A handle is not
 * thread-safe; use one handle per stream / rank (one process per GPU, as run.py:846-856 does).
 *
 * Scope of state: hb_init binds the PROCESS to one device (the reference's model: one process per GPU); handles are independent
 * of each other.  There are no tuning switches in this header: the kernel-variant knobs used for A/B measurements live in
 * hirest_b200_debug.h (hb_debug_set / HB_DEBUG_* environment variables) and are not part of the drop-in boundary.
 */
#ifndef HIREST_B200_H_
#define HIREST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HB_API __attribute__((visibility("default")))
#else
#define HB_API
#endif

#define HB_OK 0
#define HB_ERR_INVALID (-22)   /* bad argument / unsupported shape */
#define HB_ERR_NOMEM (-12)     /* cudaMalloc failed */
#define HB_ERR_CUDA (-5)       /* CUDA runtime / driver error, see hb_last_error() */
#define HB_ERR_NODEVICE (-19)  /* no sm_100 device */

/* ---- library ------------------------------------------------------------------------------- */
/* Checks that `device` is compute capability 10.x, resolves the driver entry points and binds the PROCESS to that device (a second
 * call with another device fails: one process per GPU).  The calling thread's current device is left as it was; callers make
 * `device` current around the other calls (torch.cuda.device / cudaSetDevice).  Must be called before anything else. */
HB_API int hb_init(int device);
HB_API const char* hb_last_error(void);
HB_API const char* hb_strerror(int code);
/* Number of kernels launched by this library since process start (bench.py's gpu_launches). */
HB_API int64_t hb_launch_count(void);
/* Per-launch timing for bench.py's roofline: while enabled every kernel launch of this library is bracketed
 * by CUDA events on its own stream.  hb_profile_stop synchronises the device and sums per category:
 * 0 GEMM bf16-out (qkv), 1 GEMM GELU (fc1), 2 GEMM fp32-out (patch/proj/fc2/head/similarity), 3 ViT attention,
 * 4 LayerNorm, 5 other.  flops = algorithmic 2*M*N*K (attention: 4*B*H*257*257*88). */
#define HB_PROF_CATS 6
typedef struct {
  double ms[HB_PROF_CATS];
  double flops[HB_PROF_CATS];
  int64_t launches[HB_PROF_CATS];
} HbProfileSummary;
HB_API int hb_profile_start(void);
HB_API int hb_profile_stop(HbProfileSummary* out);

/* ---- EVA ViT frame encoder (EVA_clip/vit_model.py) -------------------------------------------- */
typedef struct {
  int image_size;   /* 224 */
  int patch_size;   /* 14  */
  int width;        /* 1408 */
  int layers;       /* 40  */
  int heads;        /* 16 (head_dim must be 88) */
  int mlp_hidden;   /* 6144 = int(width * 4.3637) */
  int embed_dim;    /* 1024 */
  float ln_eps;     /* 1e-6 (eva_model.py:304) */
} HbVitConfig;

/* fp32 device pointers in the reference state_dict layout.  Per-layer members are
 * HOST arrays of `layers` device pointers.  Borrowed only during hb_vit_create (weights are repacked). */
typedef struct {
  const float* cls_token;      /* [1,1,D] */
  const float* pos_embed;      /* [1,T,D] */
  const float* patch_w;        /* patch_embed.proj.weight [D,3,P,P] */
  const float* patch_b;        /* [D] */
  const float* const* norm1_w; const float* const* norm1_b;
  const float* const* q_bias;  const float* const* v_bias;   /* attn.q_bias / attn.v_bias [D] */
  const float* const* qkv_w;   /* attn.qkv.weight [3D, D] */
  const float* const* proj_w;  const float* const* proj_b;
  const float* const* norm2_w; const float* const* norm2_b;
  const float* const* fc1_w;   const float* const* fc1_b;    /* [F,D], [F] */
  const float* const* fc2_w;   const float* const* fc2_b;    /* [D,F], [D] */
  const float* norm_w; const float* norm_b;                  /* final norm */
  const float* head_w; const float* head_b;                  /* [E,D], [E] */
} HbVitWeights;

typedef struct HbVit HbVit;
HB_API int hb_vit_create(const HbVitConfig* cfg, const HbVitWeights* w, int max_batch, void* stream, HbVit** out);
/* frames: fp32 [B,3,S,S] NCHW; out: fp32 [B, embed_dim] (un-normalised, like encode_image). */
HB_API int hb_vit_encode(HbVit* m, const float* frames, int64_t B, float* out, void* stream);
/* Same from raw uint8 frames [B,3,S,S] (0..255): the ToTensor + Normalize(mean, std) steps of the reference's CPU
 * preprocessing (EVA_clip/eva_clip.py:144-153) are folded into the patch gather; mean / std are HOST arrays of 3 floats.
 * 4x fewer host->device bytes than fp32 frames. */
HB_API int hb_vit_encode_u8(HbVit* m, const uint8_t* frames, int64_t B, const float* mean, const float* stdv, float* out, void* stream);
/* Debug / parity taps: copy the fp32 residual stream [B*T, D] after `layer` blocks (0 = after patch embed)
 * of the most recent chunk into `dst`.  Must be requested before encode via hb_vit_set_tap(layer, dst). */
HB_API int hb_vit_set_tap(HbVit* m, int layer, float* dst);
HB_API void hb_vit_destroy(HbVit* m);

/* ---- EVA-CLIP text tower (EVA_clip/eva_model.py:177-250) -------------------------------------- */
typedef struct {
  int context_length; /* 77 */
  int vocab_size;     /* 49408 */
  int width;          /* 768 */
  int heads;          /* 12 (head_dim must be 64) */
  int layers;         /* 12 */
  int embed_dim;      /* 1024 */
  float ln_eps;       /* 1e-5 */
  int precise;        /* != 0: fp32-accurate path (3-term split-bf16 GEMMs, fp32 LayerNorm / attention / residual): encode_text
                         within ~1e-5 of the fp32 reference, ~3x the (tiny) GEMM work.  0: plain bf16 GEMMs + bf16 attention. */
} HbTextConfig;

typedef struct {
  const float* token_embedding;       /* [V, W] */
  const float* positional_embedding;  /* [C, W] */
  const float* const* ln1_w; const float* const* ln1_b;
  const float* const* in_proj_w;      /* attn.in_proj_weight [3W, W] */
  const float* const* in_proj_b;      /* [3W] */
  const float* const* out_proj_w; const float* const* out_proj_b;
  const float* const* ln2_w; const float* const* ln2_b;
  const float* const* fc_w;  const float* const* fc_b;      /* mlp.c_fc   [4W, W] */
  const float* const* cproj_w; const float* const* cproj_b; /* mlp.c_proj [W, 4W] */
  const float* ln_final_w; const float* ln_final_b;
  const float* text_projection;       /* [W, E] */
} HbTextWeights;

typedef struct HbText HbText;
HB_API int hb_text_create(const HbTextConfig* cfg, const HbTextWeights* w, int max_batch, void* stream, HbText** out);
/* ids: int64 [Q, context_length]; out: fp32 [Q, embed_dim] (un-normalised, like encode_text). */
HB_API int hb_text_encode(HbText* m, const int64_t* ids, int64_t Q, float* out, void* stream);
HB_API void hb_text_destroy(HbText* m);

/* ---- MomentModel shared encoder + heads (modeling.py:155-224, 272-474, 529-554) ----------------------- */
typedef struct {
  int embed_dim;  /* 512  (modeling.py:26) */
  int hidden;     /* 768  (visual_config.json) */
  int heads;      /* 12, head_dim must be 64 */
  int ffn;        /* 3072 */
  int layers;     /* 2    (args.py:53) */
  int asr_dim;    /* 384; 0 = ASR-free model (modeling.py:28-35): asr_* weights and the asr argument may be NULL */
  int clip_dim;   /* 1024 */
  int max_pos;    /* rows of visual.embeddings.position_embeddings (2048) */
} HbMomentConfig;

/* fp32 device pointers, reference state_dict tensors. head_w/head_b are the start / end / segment
 * predictors stacked to [3, hidden] / [3]. Per-layer members: host arrays of `layers` device pointers. */
typedef struct {
  const float* asr_ln_w; const float* asr_ln_b; const float* asr_w; const float* asr_b;
  const float* temp_w1; const float* temp_b1; const float* temp_w2; const float* temp_b2;
  const float* mask_embed; const float* boundary_embed;
  const float* head_w; const float* head_b;
  const float* vis_norm_w; const float* vis_norm_b;
  const float* clip_g_map_w; const float* clip_g_map_b; const float* clip_g_map_text_w; const float* clip_g_map_text_b;
  const float* emb_w; const float* emb_b; const float* pos_emb; const float* emb_ln_w; const float* emb_ln_b;
  const float* const* q_w; const float* const* q_b; const float* const* k_w; const float* const* k_b;
  const float* const* v_w; const float* const* v_b; const float* const* ao_w; const float* const* ao_b;
  const float* const* ao_ln_w; const float* const* ao_ln_b; const float* const* i_w; const float* const* i_b;
  const float* const* o_w; const float* const* o_b; const float* const* o_ln_w; const float* const* o_ln_b;
} HbMomentWeights;

typedef struct HbMoment HbMoment;
HB_API int hb_moment_create(const HbMomentConfig* cfg, const HbMomentWeights* w, int64_t max_rows, int max_batch, void* stream,
                            HbMoment** out);
#define HB_MOMENT_REUSE_BASE 1 /* video/text/asr/time terms unchanged since the previous call (segmentation iterations) */
/* foward_moment_shared + the three 768->1 heads.  video fp32 [B,T,clip_dim]; text_feat fp32 [B,clip_dim] (encode_text
 * output); asr fp32 [B,T,asr_dim]; masks int64 [B,T] (boundary_mask may be NULL, as in moment retrieval).
 * out_feats fp32 [B,T,hidden] (may be NULL); out_logits fp32 [B,T,3] = (start, end, segment).  All GEMMs run as 3-term
 * split-bf16 tcgen05 GEMMs (fp32-accurate) and attention in fp32, so that the integer outputs of the MR / MS decoders
 * match the fp32 reference.  T <= max_pos (2048, the reference's position-embedding cap, modeling.py:110). */
HB_API int hb_moment_forward(HbMoment* m, const float* video, const float* text_feat, const float* asr,
                             const int64_t* video_mask, const int64_t* moment_mask, const int64_t* boundary_mask, int B, int T,
                             int flags, float* out_feats, float* out_logits, void* stream);
HB_API void hb_moment_destroy(HbMoment* m);
/* MR decode (modeling.py:294-298): pred int64 [B,2] = argmax of start / end logits with -1e10 on padded frames. */
HB_API int hb_moment_mr_decode(const float* logits, const int64_t* video_mask, int64_t* pred, int B, int T, void* stream);
/* One MS iteration (modeling.py:394-433) on the device; masks updated in place; steps int32 [B,max_steps,2],
 * nsteps int32 [B]; probs_out fp32 [B,T] optional. */
HB_API int hb_moment_ms_step(const float* logits, int64_t* moment_mask, int64_t* boundary_mask, int32_t* steps, int32_t* nsteps,
                             int max_steps, int B, int T, double threshold, float* probs_out, void* stream);
/* trim_feats (modeling.py:529-554): x fp32 [B,T,C], mask int64 [B,T] -> out fp32 [B,F,C]. */
HB_API int hb_trim_feats(const float* x, const int64_t* mask, float* out, int B, int T, int C, int F, void* stream);

/* ---- caption decoder + beam search (module_decoder.py:372-406, beam.py:70-123, train.py:547-599, modeling.py:575-613) -- */
typedef struct {
  int hidden;     /* 768 */
  int heads;      /* 12 */
  int ffn;        /* 3072 */
  int layers;     /* 2 (args.py:54) */
  int vocab;      /* 30522 */
  int max_pos;    /* 512 rows of decoder position embeddings */
  int max_words;  /* 48 (args.py:52) */
  int bos;        /* [CLS] = 101 */
  int eos;        /* [SEP] = 102 */
} HbDecoderConfig;

/* fp32 device pointers (clip4cap_model.decoder.* of the reference state_dict).  Per-layer members: host arrays of device pointers. */
typedef struct {
  const float* word_emb; const float* pos_emb; const float* emb_ln_w; const float* emb_ln_b;
  const float* const* sq_w; const float* const* sq_b; const float* const* sk_w; const float* const* sk_b;
  const float* const* sv_w; const float* const* sv_b; const float* const* so_w; const float* const* so_b;
  const float* const* so_ln_w; const float* const* so_ln_b;
  const float* const* eq_w; const float* const* eq_b; const float* const* ek_w; const float* const* ek_b;
  const float* const* ev_w; const float* const* ev_b; const float* const* eo_w; const float* const* eo_b;
  const float* const* eo_ln_w; const float* const* eo_ln_b;
  const float* const* i_w; const float* const* i_b; const float* const* o_w; const float* const* o_b;
  const float* const* o_ln_w; const float* const* o_ln_b;
  const float* cls_dense_w; const float* cls_dense_b; const float* cls_ln_w; const float* cls_ln_b; const float* cls_bias;
} HbDecoderWeights;

typedef struct HbDecoder HbDecoder;
HB_API int hb_decoder_create(const HbDecoderConfig* cfg, const HbDecoderWeights* w, int max_inst, int max_beam, int max_enc_len,
                             void* stream, HbDecoder** out);
/* Start a beam search: enc fp32 [n_inst, enc_len, hidden] (shared-encoder output of the trimmed clip).  Projects the
 * cross-attention keys / values once (the reference re-projects them every step) and resets the beams. */
HB_API int hb_decoder_begin(HbDecoder* d, const float* enc, int n_inst, int enc_len, int beam, void* stream);
/* One decode step for all instances: KV-cached self-attention, cross-attention with the reference's -10000-on-every-key mask,
 * last-position classifier, log_softmax, Beam.advance, cache re-order.  Finished instances are frozen. No host sync. */
HB_API int hb_decoder_step(HbDecoder* d, void* stream);
/* Copy out: prev_k / ys int32 [max_words, n_inst, beam], nsteps int32 [n_inst], done int32 [n_inst], scores fp32 [n_inst, beam]. */
HB_API int hb_decoder_read(HbDecoder* d, int32_t* prev_k, int32_t* ys, int32_t* nsteps, int32_t* done, float* scores, void* stream);
HB_API void hb_decoder_destroy(HbDecoder* d);

/* ---- retrieval scoring ------------------------------------------------------------------------ */
/* out[v,:] = l2norm(mean_f emb[v,f,:]); emb fp32 [V,F,E]; out fp32 [V,E].  F = 1 gives plain L2 normalise. */
HB_API int hb_pool_normalize(const float* emb, int64_t V, int F, int E, float* out, void* stream);
/* Cached-feature path (inference_video_retrieval.py:298-327): feats fp32 [sum_T, E] = the per-video feature tensors packed
 * back to back, offsets int64 [V+1] (device).  Per video: rows np.linspace(0, T-1, n_sub).astype(int) (all rows if
 * n_sub <= 0) -> mean -> L2 normalise.  out fp32 [V,E].  Videos must have T >= 1. */
HB_API int hb_subsample_pool_normalize(const float* feats, const int64_t* offsets, int64_t V, int n_sub, int E, float* out, void* stream);
/* Same with the packed features stored as bf16 (half the bytes of the store; the values are the reference's fp32 features rounded
 * to bf16, so embeddings agree to ~2e-3 relative instead of bit for bit). */
HB_API int hb_subsample_pool_normalize_bf16(const void* feats, const int64_t* offsets, int64_t V, int n_sub, int E, float* out, void* stream);
/* Dataset-side frame resampling of cached features for the MomentModel feed (hirest_dataset.py:333-356, 383-403): feats fp32
 * [sum_T, C] packed per video, offsets int64 [V+1] -> out fp32 [V, n_out, C].  T > n_out: rows np.linspace(0, T-1, n_out).astype(int);
 * T <= n_out: repeat-pad (row k fills slots [(k*n_out)//T, ((k+1)*n_out)//T)); T == 0: zeros. */
HB_API int hb_resample_rows(const float* feats, const int64_t* offsets, int64_t V, int n_out, int C, float* out, void* stream);
/* ASR feature warping (hirest_dataset.py:370-381): asr fp32 [sum_S, C] = one feature row per subtitle sentence, packed per video
 * (sub_offsets int64 [V+1]); starts / ends int32 [sum_S] = the sentences' start / end seconds; frame_offsets int64 [V+1] = rows of
 * the output per video (the video lengths); row_video int32 [rows] = video of every output row.  out fp32 [rows, C]: row t of a
 * video = the feature of the LAST sentence with start <= t < end, zeros if none. */
HB_API int hb_asr_warp(const float* asr, const int64_t* sub_offsets, const int32_t* starts, const int32_t* ends, const int64_t* frame_offsets,
                       const int32_t* row_video, int64_t rows, int C, float* out, void* stream);
/* scores[q,v] = <text[q,:], video[v,:]>; text fp32 [Q,E], video fp32 [V,E], scores fp32 [Q, ld_scores].
 * exact != 0: one bf16 GEMM over 3-way split operands (hi+mid+lo = the fp32 value exactly), K = 6E: fp32-accurate
 * scores, so top-k matches the reference's fp32 matmul up to fp32 rounding ties; exact == 0: single plain bf16 GEMM. */
HB_API int hb_similarity(const float* text, int64_t Q, const float* video, int64_t V, int E, float* scores,
                  int64_t ld_scores, int exact, void* stream);

/* ---- frame preprocessing --------------------------------------------------- */
/* Resize(S, BICUBIC) + CenterCrop(S) of decoded RGB frames, bit-identical to the reference's CPU transform
 * (torchvision on PIL images: EVA_clip/eva_clip.py:144-147; callers inference_video_retrieval.py:43-49,
 * extract_features.py:48-50).  src: uint8 [B,H,W,3] (device, all frames of one call share H x W);
 * dst: uint8 [B,3,S,S] (device), ready for hb_vit_encode_u8 which folds ToTensor + Normalize.  S <= 256.
 * The fixed-point weight tables (Pillow's, computed on the host in double) are cached per (device, H, W, S). */
HB_API int hb_resize_crop_u8(const uint8_t* src, int64_t B, int H, int W, int S, uint8_t* dst, void* stream);
/* Geometry of the above: resized size (torchvision Resize(int)) and crop offsets (CenterCrop). */
HB_API int hb_resize_geometry(int H, int W, int S, int* new_h, int* new_w, int* top, int* left);

/* Host-only (no device call): the integer tables hb_resize_crop_u8 uses, for verification against Pillow.
 * info[0..5] = taps per column, taps per row, first source column, bytes per source row staged, rows per CTA, smem bytes;
 * tables (capacity cap ints) = col bounds [S,2] (first tap relative to info[2], taps) | col weights [S,info[0]] |
 * row bounds [S,2] | row weights [S,info[1]].  Returns the number of ints (written only if cap is large enough) or < 0. */
HB_API int64_t hb_resize_tables(int H, int W, int S, int info[6], int* tables, int64_t cap);

/* ---- generic fused linear (tcgen05 GEMM) ------------------------------------------------------ */
#define HB_EPI_BF16 0       /* out bf16 = x W^T + b                      */
#define HB_EPI_GELU_BF16 1  /* out bf16 = gelu_erf(x W^T + b)            */
#define HB_EPI_F32 2        /* out f32  = x W^T + b (+ resid)            */
/* x: bf16 [M, ldx]; w: bf16 [N, ldw] (K contiguous, nn.Linear layout); bias fp32 [N] or NULL;
 * resid fp32 [M, ldo] or NULL (HB_EPI_F32 only, may alias out).  K % 8 == 0, N % 16 == 0, 16-byte aligned. */
HB_API int hb_linear(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, const float* resid,
              void* out, int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue, void* stream);

/* LayerNorm over the last dim: x fp32 [rows, D] -> y (bf16 if out_bf16 else fp32) [rows, D]. */
HB_API int hb_layernorm(const float* x, const float* w, const float* b, float eps, int64_t rows, int D, void* y, int out_bf16,
                 void* stream);

/* ViT attention (vit_model.py:127-147): qkv bf16 [B*257, 3*H*88] (q pre-scaled) -> out bf16 [B*257, H*88]. */
HB_API int hb_vit_attention(const void* qkv, void* out, int64_t B, int H, void* stream);

/* head_dim-64 attention (text tower / MomentModel / decoder), see hb_attn.cuh SmallAttnParams.
 * q [B,Tq,*], k/v [B,Tk,*] bf16 with row strides ld* and batch strides bs* (elements); mask_mode 0 none,
 * 1 causal (-inf), 2 additive constant (+ soft causal -10000 if causal_soft). */
HB_API int hb_small_attention(const void* q, const void* k, const void* v, void* out, int B, int H, int Tq, int Tk, int ldq,
                       int ldk, int ldv, int ldo, int64_t bsq, int64_t bsk, int64_t bsv, int64_t bso, float scale,
                       int mask_mode, float mask_const, int causal_soft, void* stream);

/* fp32 in / fp32 out variant (MomentModel encoder, precise text tower, caption decoder): same arguments with fp32 pointers.
 * use_tensor_cores != 0: split-bf16 tcgen05 UMMAs for q.k^T and p.v with fp32 softmax (fp32-accurate, hb_attn_tc.cu);
 * 0: CUDA-core kernel.  Any Tq / Tk. */
HB_API int hb_small_attention_f32(const float* q, const float* k, const float* v, float* out, int B, int H, int Tq, int Tk, int ldq,
                                  int ldk, int ldv, int ldo, int64_t bsq, int64_t bsk, int64_t bsv, int64_t bso, float scale,
                                  int mask_mode, float mask_const, int causal_soft, int use_tensor_cores, void* stream);

/* ---- host-only batch tokenisers (no device work; usable before hb_init) ---------------------------------------------------
 * ASCII fast path of the two tokenisers on the path; a text with a byte >= 0x7F, a control character other than tab / newline /
 * carriage return (or, for BPE, an '&': HTML entities) is flagged and left to the caller's Unicode implementation
 * (hirest_b200/wordpiece.py, tokenizer.py), which gives the same ids for ASCII input.
 *
 * WordPiece = clip4caption/modules/tokenization.py BertTokenizer (basic tokenizer + greedy longest-match pieces, never_split =
 * the five special tokens) as used by hirest_dataset.py:119-121 and clip4cap_get_text (:533-580).
 * tokens: n_tokens NUL-terminated UTF-8 strings back to back (tokens_bytes in total), ids[i] = id of token i (a repeated token keeps
 * its last id, like the reference's dict).  [UNK], [CLS], [SEP] must be present. */
typedef struct HbWordPiece HbWordPiece;
HB_API int hb_wordpiece_create(const char* tokens, int64_t tokens_bytes, const int64_t* ids, int n_tokens, int do_lower_case,
                               HbWordPiece** out);
HB_API void hb_wordpiece_destroy(HbWordPiece* w);
/* clip4cap_get_text for n captions: input_ids = [CLS] w1 .., target_ids = w1 .. [SEP], mask, each int64 [n, max_words] (pieces
 * truncated to max_words - 1).  fallback[i] = 1: row i was NOT written (non-ASCII caption), the caller fills it. */
HB_API int hb_wordpiece_encode_captions(const HbWordPiece* w, const char* const* captions, int n, int max_words, int64_t* input_ids,
                                        int64_t* target_ids, int64_t* mask, uint8_t* fallback);
/* CLIP byte-pair encoding = EVA_clip/simple_tokenizer.py + clip.tokenize (hirest_dataset.py:528, inference_video_retrieval.py:203-206).
 * merges: the merge table's lines 1 .. 48894 ("left right", UTF-8), newline separated, in rank order. */
typedef struct HbBpe HbBpe;
HB_API int hb_bpe_create(const char* merges, int64_t merges_bytes, HbBpe** out);
HB_API void hb_bpe_destroy(HbBpe* b);
/* out int64 [n, context_length] = [SOT] ids [EOT] 0 0 ..; status[i]: 0 written, 1 not written (caller's Unicode path),
 * 2 longer than context_length and truncate == 0 (clip.tokenize raises).  Not thread-safe per handle (memoises words). */
HB_API int hb_bpe_tokenize(HbBpe* b, const char* const* texts, int n, int context_length, int truncate, int64_t* out, uint8_t* status);

#ifdef __cplusplus
}
#endif
#endif /* HIREST_B200_H_ */
