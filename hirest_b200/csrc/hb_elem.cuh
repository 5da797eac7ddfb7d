// hb_elem.cuh — HBM-bound row kernels (implementation in hb_elem.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace hb {

struct LayerNormParams {
  const float* x = nullptr;      // fp32 rows, leading dim ldx
  long long ldx = 0;
  const int* row_idx = nullptr;  // optional gather: output row r reads input row row_idx[r]
  void* y = nullptr;             // bf16 or fp32 rows, leading dim ldy
  long long ldy = 0;
  const float* w = nullptr;
  const float* b = nullptr;
  float eps = 1e-5f;
  int rows = 0;
  int D = 0;
};
int layernorm_launch(const LayerNormParams& p, bool out_bf16, cudaStream_t s);
int row_stats_launch(const float* x, __nv_bfloat16* xb, float* stats, long long rows, int D, int slots, cudaStream_t s);
int im2col_patch_launch(const float* img, __nv_bfloat16* out, int B, int S, int P, int ldo, cudaStream_t s);
int im2col_patch_u8_launch(const uint8_t* img, __nv_bfloat16* out, int B, int S, int P, int ldo, const float mean[3],
                           const float stdv[3], cudaStream_t s);
int cls_row_launch(float* x, const float* cls, const float* pos, int B, int T, int D, cudaStream_t s);
int text_embed_launch(const long long* ids, const float* tok, const float* pos, float* x, int* eot_row, int Q, int C, int W,
                      int V, cudaStream_t s);
int pool_normalize_launch(const float* emb, void* out, long long V, int F, int E, bool normalize, bool out_bf16,
                          cudaStream_t s);
int subsample_pool_normalize_launch(const float* feats, const long long* offsets, float* out, long long V, int n_sub, int E,
                                    cudaStream_t s);
int subsample_pool_normalize_bf16_launch(const __nv_bfloat16* feats, const long long* offsets, float* out, long long V, int n_sub, int E,
                                         cudaStream_t s);
// dataset-side frame resampling / ASR warping of packed cached features (hirest_dataset.py:328-405), see hb_elem.cu
int resample_rows_launch(const float* feats, const long long* offsets, float* out, long long V, int n_out, int C, cudaStream_t s);
int asr_warp_launch(const float* asr, const long long* sub_offsets, const int* starts, const int* ends, const long long* frame_offsets,
                    const int* row_video, float* out, long long rows, int C, cudaStream_t s);
int split_bf16_launch(const float* x, __nv_bfloat16* out, long long rows, int E, int mode, cudaStream_t s);
int f32_to_bf16_launch(const float* x, __nv_bfloat16* y, long long n, cudaStream_t s);

}  // namespace hb
