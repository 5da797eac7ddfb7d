// hb_attn.cuh — attention kernels (implementation in hb_attn.cu / hb_attn_small.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace hb {

// EVA ViT-g/14 attention (EVA_clip/vit_model.py:127-147): 257 tokens, head_dim 88.
// qkv: bf16 [B*257, 3*H*88] with feature index = which*H*88 + head*88 + d (vit_model.py:127), q pre-scaled.
// out: bf16 [B*257, H*88]  (== (attn @ v).transpose(1,2).reshape(B,N,C), vit_model.py:147).
struct AttnParams {
  const __nv_bfloat16* qkv = nullptr;
  __nv_bfloat16* out = nullptr;
  int B = 0;
  int H = 0;
  long long* timing = nullptr;  // debug (HB_ATTN_TIMING builds): 16 clock64 stamps per CTA
  int prefetch_ahead = 0;       // v2: L2-prefetch the operands of block (blockIdx + prefetch_ahead); 0 = off
  int dots_late = 0;            // v3: bit t set = tile t computes the next item's extra-token dot products AFTER its output phase
};
int vit_attn_launch(const AttnParams& p, cudaStream_t stream);   // v1: one CTA per (frame, head), P through smem
int vit_attn2_launch(const AttnParams& p, cudaStream_t stream);  // v2: one CTA per 128-query tile, 2 CTAs/SM, P in TMEM
int vit_attn3_launch(const AttnParams& p, int num_sms, cudaStream_t stream);  // v3: persistent CTA per SM, pipelined over (frame, head)

// Generic small-sequence attention on CUDA cores (text tower, MomentModel encoder, caption decoder):
//   q: bf16 [B, Tq, ldq] (+ head*64), k/v: bf16 [B, Tk, ldk], head_dim 64, scores = q.k * scale
//   mask_mode 0: none; 1: causal (key > query masked with -inf, EVA_clip/eva_model.py:224-230);
//   2: additive constant `mask_const` on every logit, then causal -10000 if causal_soft != 0
//      (clip4caption/modules/module_visual.py:406-414, module_decoder.py:385-396).
struct SmallAttnParams {
  const __nv_bfloat16* q = nullptr;
  const __nv_bfloat16* k = nullptr;
  const __nv_bfloat16* v = nullptr;
  __nv_bfloat16* out = nullptr;  // [B, Tq, ldo] (+ head*64)
  int B = 0, H = 0, Tq = 0, Tk = 0;
  int ldq = 0, ldk = 0, ldv = 0, ldo = 0;  // row strides in elements
  long long bsq = 0, bsk = 0, bsv = 0, bso = 0;  // batch strides in elements
  float scale = 1.f;
  int mask_mode = 0;
  float mask_const = 0.f;
  int causal_soft = 0;
};
int small_attn_launch(const SmallAttnParams& p, cudaStream_t stream);

// fp32 in / fp32 out variant (keys streamed in tiles of 256 with an online softmax: any length), same mask semantics;
// used by the MomentModel / caption decoder fp32-accurate path.
struct SmallAttnF32Params {
  const float* q = nullptr;
  const float* k = nullptr;
  const float* v = nullptr;
  float* out = nullptr;
  int B = 0, H = 0, Tq = 0, Tk = 0;
  int ldq = 0, ldk = 0, ldv = 0, ldo = 0;
  long long bsq = 0, bsk = 0, bsv = 0, bso = 0;
  float scale = 1.f;
  int mask_mode = 0;
  float mask_const = 0.f;
  int causal_soft = 0;
  int kv_div = 1;  // K/V batch index = b / kv_div (beams of one instance share the encoder keys / values)
  // Decode kernel only (Tq == 1, Tk <= 64): key t of batch row b lives in K / V batch kv_row_idx[b * ld_idx + t] instead of b — the
  // caption decoder's beam re-ordering as an index table instead of copying the KV cache (kv_div must be 1)
  const int* kv_row_idx = nullptr;
  int ld_idx = 0;
  // Decode kernel only: also write the output as the [lo | hi | hi] split-bf16 operand of the following linear,
  // op_out [B, 3 * ld_op] with ld_op = H * 64 (what split3_act_kernel would produce from `out`)
  __nv_bfloat16* op_out = nullptr;
  int ld_op = 0;
};
int small_attn_f32_launch(const SmallAttnF32Params& p, cudaStream_t stream);
// Same contract on the tensor cores (hb_attn_tc.cu: split-bf16 UMMAs for q.k^T and p.v, fp32 softmax; fp32-accurate).  Needs a
// device workspace of small_attn_tc_workspace() bytes (bf16 split copies of q / k / v); kv_div must be 1.  Returns -8 if the
// workspace is too small.
size_t small_attn_tc_workspace(int B, int H, int Tq, int Tk);
int small_attn_tc_launch(const SmallAttnF32Params& p, void* workspace, size_t workspace_bytes, cudaStream_t stream);
void small_attn_tc_set_version(int v);   // 2 (default): two CTAs per SM, [hi | lo] operand images; 1: one CTA per SM, double-buffered

}  // namespace hb
