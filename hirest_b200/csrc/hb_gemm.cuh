// hb_gemm.cuh — host-side interface of the tcgen05 GEMM (implementation in hb_gemm.cu).
//
//   out[M,N] = epilogue( A[M,K] (bf16, K contiguous) x W[N,K]^T (bf16, K contiguous) )
//
// This is the dense contraction behind every nn.Linear / Conv2d-as-GEMM on the reference hot path
// (EVA_clip/vit_model.py:57-61,124-126,148,198,350; EVA_clip/eva_model.py:132,143-150,249;
//  clip4caption/modules/module_visual.py:118-233; module_decoder.py:160-292).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace hb {

enum GemmEpilogue : int {
  EPI_BF16 = 0,       // out bf16 = (acc + bias) * (col < qcols ? qscale : 1)
  EPI_GELU_BF16 = 1,  // out bf16 = gelu_erf(acc + bias)
  EPI_F32 = 2,        // out f32  = acc + bias [+ resid[row_out, col]]
  EPI_F32_ROWADD = 3, // out f32  = acc + bias + rowadd[row % remap_in, col], rows remapped (patch embed + pos embed)
};

struct GemmParams {
  int M = 0, N = 0, K = 0;        // logical sizes; K is the padded leading K of both operands' maps
  const float* bias = nullptr;    // [N] or null
  void* out = nullptr;            // bf16 or f32, row-major, leading dim ldo (elements)
  int ldo = 0;
  const float* resid = nullptr;   // EPI_F32: fp32 residual, indexed like out (may alias out)
  const float* rowadd = nullptr;  // EPI_F32_ROWADD: [remap_in, N] fp32 added by (row % remap_in)
  int remap_in = 0;               // if > 0: out_row = (row / remap_in) * remap_out + row % remap_in + remap_off
  int remap_out = 0;
  int remap_off = 0;
  float qscale = 1.0f;            // EPI_BF16 only
  int qcols = 0;
};

// Resolve cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency).
// Returns 0 on success.
int tmap_init();

// 2D bf16 tensor map: global [rows, cols] with leading dimension ld (elements, ld*2 % 16 == 0),
// box = [box_rows, 64 cols], SWIZZLE_128B, zero fill out of bounds.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

// 3D bf16 tensor map (dims/strides innermost first; strides in bytes for dims 1,2), SWIZZLE_128B, zero OOB fill.
int make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, const uint64_t dims[3], const uint64_t strides_bytes[2],
                      const uint32_t box[3]);

// Rows of W one CTA loads per stage for cta-group size cg (box_rows for the W map).
constexpr uint32_t gemm_w_box_rows(int cg) { return 256u / static_cast<uint32_t>(cg); }
constexpr uint32_t gemm_a_box_rows() { return 128u; }

// Launch. tmA must have box_rows = 128, tmW box_rows = gemm_w_box_rows(cg). cg in {1,2}.
// Returns cudaError_t as int.
int gemm_launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p, int epi, int cg, int num_sms,
                cudaStream_t stream);

}  // namespace hb
