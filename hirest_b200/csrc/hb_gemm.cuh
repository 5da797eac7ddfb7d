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
  // LayerNorm folded into the GEMMs (no LayerNorm kernel, no normalised copy of the residual stream in HBM):
  //   LN(x) W^T = rstd * (x (gamma*W)^T) - rstd*mean * c1 + c2,   c1[n] = sum_k gamma_k W[n,k],  c2[n] = sum_k beta_k W[n,k] + b[n]
  EPI_F32_STATS = 4,  // EPI_F32 + writes a bf16 copy of the new residual row and accumulates its (sum, sum of squares)
  EPI_BF16_LN = 5,    // out bf16 = (rstd*acc - rstd*mean*c1 + c2) * (col < qcols ? qscale : 1)   (A operand = raw bf16 residual)
  EPI_GELU_BF16_LN = 6,  // out bf16 = gelu_erf(rstd*acc - rstd*mean*c1 + c2)
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
  float qscale = 1.0f;            // EPI_BF16 / EPI_BF16_LN only
  int qcols = 0;
  // LayerNorm fold
  void* xb_out = nullptr;         // EPI_F32_STATS: bf16 copy of out, leading dim ld_xb
  int ld_xb = 0;
  float* stats_out = nullptr;     // EPI_F32_STATS: [M, ln_slots, 2] partial (sum, sumsq), one slot per (N tile, column half)
  const float* stats_in = nullptr;  // *_LN kinds: [M, ln_slots, 2] row statistics of the A operand's fp32 source
  const float* c1 = nullptr;      // *_LN kinds: [N]; `bias` carries c2
  float ln_eps = 1e-6f;
  int ln_dim = 0;                 // number of features the statistics were taken over
  int ln_slots = 1;               // partial-sum slots per row (producer writes slot 2 * n_tile + column half: >= 2 * ceil(N / 256))
  int prefetch_chunks = 0;        // EPI_F32*: residual lines are L2-prefetched this many 32-column chunks ahead (0 = off)
  int a_hint = 0, w_hint = 2;     // L2 eviction priority of the TMA operand loads: 0 normal, 1 evict-first, 2 evict-last
  int* sched = nullptr;           // optional dynamic tile scheduler: 2 zero-initialised device ints owned by the caller (one
                                  // pair per stream; the kernel re-zeroes them).  nullptr = static round-robin schedule.
  int n_tiles = 0;                // > minimum: use this many (balanced) N tiles, e.g. a multiple of the worker count; 0 = fewest
  int m_fastest = 0;              // tile order: consecutive work items share the N tile (W operand) instead of the M block (A operand):
                                  // for a few M blocks over a huge N (vocabulary projection at M = 768: W was read once per M block)
  int k_splits = 1;               // EPI_F32 only, bias / resid null: slice ks of K (k-blocks [ks*kbs, ks*kbs+kbs), kbs = ceil(k_blocks /
  long long split_stride = 0;     // k_splits)) writes its raw partial sums to out + ks * split_stride (elements); caller reduces
  int balanced_n = 1;             // N tiling: equal-cost tiles (see NTiling in hb_gemm.cu) instead of 256-wide tiles + narrow tail
};

// Resolve cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency).
// Returns 0 on success.
int tmap_init();

// 2D bf16 tensor map: global [rows, cols] with leading dimension ld (elements, ld*2 % 16 == 0),
// box = [box_rows, 64 cols], SWIZZLE_128B, zero fill out of bounds.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

// 3D bf16 tensor map (dims/strides innermost first; strides in bytes for dims 1,2), zero OOB fill;
// swizzle_bytes 128 (default) / 64 / 0 = SWIZZLE_128B / SWIZZLE_64B / SWIZZLE_NONE.
int make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, const uint64_t dims[3], const uint64_t strides_bytes[2],
                      const uint32_t box[3], int swizzle_bytes = 128);

// Rows of W one CTA loads per stage for cta-group size cg (box_rows for the W map).
constexpr uint32_t gemm_w_box_rows(int cg) { return 256u / static_cast<uint32_t>(cg); }
constexpr uint32_t gemm_a_box_rows() { return 128u; }

// Host-side view of the column tiling the kernel uses (for tests): fills n0 / width of up to `cap` tiles, returns the tile count.
int gemm_n_tiling(int N, int cg, int balanced, int* n0_out, int* width_out, int cap);

// Process-wide switch between balanced N tiles (default) and 256-wide tiles + narrow tail (A/B measurements).
void gemm_set_balanced_tiles(int on);
void gemm_set_resid_prefetch_chunks(int k);
void gemm_set_l2_hints(int a_hint, int w_hint);   // -1 keeps the default (A normal, W evict-last)

// Launch. tmA must have box_rows = 128, tmW box_rows = gemm_w_box_rows(cg). cg in {1,2}.
// Returns cudaError_t as int.
int gemm_launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p, int epi, int cg, int num_sms,
                cudaStream_t stream);

}  // namespace hb
