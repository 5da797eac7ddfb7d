// hb_moment.cuh — row kernels of the MomentModel path (implementation in hb_moment.cu).
// Reference: modeling.py:155-224 (shared encoder + heads), :272-310 (MR decode), :353-474 (MS decode), :529-554 (trim).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace hb {

// fp32 rows [R, K] -> bf16 [R, 3K] = [lo | hi | hi] (activation side of the ~fp32-accurate split GEMM; the weight side is
// [hi | lo | hi], so A'.W'^T = lo.hi + hi.lo + hi.hi, smallest terms first).  gelu != 0 applies erf-GELU first.
int split3_act_launch(const float* x, __nv_bfloat16* out, long long rows, int K, int gelu, cudaStream_t s);
// Finish of a split-K GEMM: y = [LayerNorm]([gelu](sum_s part[s*split_stride + .] + bias) + resid) for rows x N (N <= 1536,
// N % 4 == 0); y (fp32) and / or op (the [lo | hi | hi] operand of the next split GEMM, [rows, 3N]) may be null.
int splitk_finish_launch(const float* part, long long split_stride, int splits, const float* bias, const float* resid,
                         const float* ln_w, const float* ln_b, float eps, int gelu, float* y, __nv_bfloat16* op, int rows, int N,
                         cudaStream_t s);
// weight repack for the split GEMM: fp32 [N, K] -> bf16 [N, 3K] = [hi | lo | hi]
int split3_weight_launch(const float* w, __nv_bfloat16* out, int N, int K, cudaStream_t s);

// h1[b,t,:] = tanh(w1 * g(b,t) + b1), g = the reference's time grid (modeling.py:178-193): (linspace(0,1,n_b)[t]-0.5)*2 for
// t < n_b (torch.linspace's fp32 two-sided formula), 0 for padding.  Output fp32 [B*T, E].
int time_tanh_launch(const long long* video_mask, const float* w1, const float* b1, float* out, int B, int T, int E,
                     cudaStream_t s);

// base[b,t,:] = TF_LN(vlin[b,t,:]; lnw, lnb, 1e-12) * that[b,:] + asr_lin[b,t,:] + temporal[b,t,:]   (modeling.py:161-196)
int moment_base_launch(const float* vlin, const float* lnw, const float* lnb, const float* that, const float* asr_lin,
                       const float* temporal, float* base, int B, int T, int E, cudaStream_t s);
// f[b,t,:] = base + boundary_embed[bm[b,t]] (if bm) + mask_embed[mm[b,t]]                              (modeling.py:171-199)
int moment_embed_launch(const float* base, const float* bemb, const float* memb, const long long* bm, const long long* mm,
                        float* f, long long rows, int E, cudaStream_t s);
// logits[r, j] = feats[r,:] . w[j,:] + b[j], j < 3 (start, end, segment), fp32 (modeling.py:218-219,319)
int moment_heads_launch(const float* feats, const float* w3, const float* b3, float* logits, long long rows, int Hd,
                        cudaStream_t s);
// MR decode (modeling.py:294-298): logits[vmask==0] = -1e10; argmax over T for start / end -> pred int64 [B,2]
int mr_argmax_launch(const float* logits, const long long* vmask, long long* pred, int B, int T, cudaStream_t s);
// One MS iteration (modeling.py:394-433) per sample on the device: mask to -FLT_MAX outside the moment, softmax over T,
// argmax, region growing at `threshold` (ratios in double, like the reference's Python floats), then
// moment_mask[l..r] = 0, boundary_mask[l] = boundary_mask[r] = 1, steps[b, nsteps[b]++] = (l, r).
int ms_step_launch(const float* logits, long long* moment_mask, long long* boundary_mask, int* steps, int* nsteps,
                   int max_steps, int B, int T, double threshold, float* probs_out, cudaStream_t s);
// trim_feats (modeling.py:529-554): frames with mask==1, truncated to F or repeat-padded -> out fp32 [B, F, C]
int trim_feats_launch(const float* x, const long long* mask, float* out, int B, int T, int C, int F, cudaStream_t s);


// ---- caption decoder (clip4caption/modules/module_decoder.py:294-406, beam.py:70-123, train.py:547-599) ----------------
// x[r,:] = TF_LN(word_emb[tok[r]] + pos_emb[pos]; w, b, 1e-12)                       (module_decoder.py:309-320)
int dec_embed_launch(const long long* tok, const float* word_emb, const float* pos_emb, const float* lnw, const float* lnb, int pos,
                     float* x, __nv_bfloat16* op /* optional: [lo|hi|hi] split operand of x, [R, 3 Hd] */, int R, int Hd, cudaStream_t s);
// append this step's self-attention key / value (columns [Hd,2Hd) / [2Hd,3Hd) of qkv) at position `pos` of the caches; with
// row_idx (int [R, Tmax], may be null) also row_idx[r, pos] = r
int dec_cache_append_launch(const float* qkv, float* kc, float* vc, int* row_idx, int pos, int R, int Tmax, int Hd, cudaStream_t s);
// beams re-order as an index table: idx_new[r, 0..len) = idx_old[(r / beam) * beam + prev_k[r], 0..len)
int dec_index_advance_launch(const int* idx_old, int* idx_new, const int* prev_k, int len, int R, int beam, int Tmax, cudaStream_t s);
// beams re-order: dst[r, 0..len) = src[(r / beam) * beam + prev_k[r], 0..len) for both caches of one layer
int dec_cache_reorder_launch(const float* ksrc, const float* vsrc, float* kdst, float* vdst, const int* prev_k, int len, int R,
                             int beam, int Tmax, int Hd, cudaStream_t s);
int gelu_f32_launch(float* x, long long n, cudaStream_t s);
// Beam.advance for every still-active instance (one CTA each): log_softmax over V per beam row, + beam scores (row 0 only
// at the first step), flat top-`beam` (sorted), prev_k = id / V, y = id % V; instance done when the best beam emits `eos`.
// step is 0-based.  Records prev_k / ys of this step at [step, inst, :]; for finished instances prev_k is the identity.
// cand_v / cand_i: device scratch of n_inst * beam * beam floats / ints (per-row candidates between the two kernels).
int beam_advance_launch(const float* logits, int ldl, int V, float* scores, int* done, int* nsteps, int* prev_k_rec, int* ys_rec,
                        long long* tok, int step, int n_inst, int beam, int eos, float* cand_v, int* cand_i, cudaStream_t s);

}  // namespace hb
