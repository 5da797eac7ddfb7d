/* hirest_b200_debug.h — kernel-variant switches for A/B measurements and cross-check tests.  NOT part of the drop-in boundary
 * (include/hirest_b200.h): nothing a caller of the reference's API needs, process-wide, and free to change between builds.
 * Every key can also be given as an environment variable HB_DEBUG_<KEY in upper case>, read once by hb_init.
 */
#ifndef HIREST_B200_DEBUG_H_
#define HIREST_B200_DEBUG_H_

#include "hirest_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* key                          values                default   meaning
 * "gemm_cta_group"             1 | 2                 2         tcgen05 cta_group of the GEMMs (2 = CTA pairs, 256 x 256 tiles)
 * "attention_version"          1 | 2 | 3             3         ViT attention kernel: 3 persistent pipelined CTA per SM (hb_attn3.cu),
 *                                                              2 one CTA per 128-query tile (hb_attn2.cu), 1 one CTA per (frame, head)
 * "small_attention_tc"         0 | 1 | 2             2         fp32 attention of the small sequence models on tensor cores (hb_attn_tc.cu;
 *                                                              2 = two CTAs per SM, 1 = the one-CTA-per-SM kernel, bit-identical results)
 *                                                              instead of CUDA cores (hb_attn_small.cu, 0); same results to ~1e-6
 * "resize_version"             1 | 2 | 3             1         hb_resize_crop_u8: 1 = byte loads + IMAD, 2 = planar word loads + dp4a on
 *                                                              byte-plane weights, 3 = planar word loads + byte extraction + IMAD
 *                                                              (both measured 15-40 % slower); bit-identical output
 * "decoder_graphs"             0 | 1                 1         caption decoder: replay each decode step as a CUDA graph from the second
 *                                                              beam search of a (n_inst, beam, enc_len) shape on; same results
 * "decoder_split_k"            0 | k-blocks          6         caption decoder: the hidden-width linears of a decode step run as split-K GEMMs
 *                                                              with this many 64-element k-blocks per slice (slice count depends on K only)
 *                                                              + one finish kernel (bias, residual, LayerNorm, next split operand); 0 = off
 * "profile_layer"              -1 | layer            -1        cudaProfilerStart / Stop around this ViT layer of every encode_image chunk
 *                                                              (ncu --profile-from-start off captures exactly its 5 kernels)
 * "decoder_kv_index"           0 | 1                 1         caption decoder: beam re-order through an index table read by the self-attention
 *                                                              (max_words <= 64) instead of copying every layer's KV prefix; same results
 * "ln_split_fuse"              0 | 1                 1         small models: LayerNorm + split-operand conversion in one kernel instead of a
 *                                                              LayerNorm kernel and a split kernel (different reduction order: ~1e-7)
 * "attention_dots_late"        0..3 (tile bit mask)  0         v3: tile t computes the next item's extra-token dot products after its output
 *                                                              phase instead of during its P.V MMAs (measured: 0.756 / 0.739 / 0.792 / 0.756 ms
 *                                                              per layer for masks 0 / 1 / 2 / 3 — the pipeline re-balances, within noise)
 * "attention_prefetch"         0 | 1                 0         v2 only: L2-prefetch the operands of the CTA one wave ahead (measured slower)
 * "ln_fold"                    0 | 1                 1         ViT handles created afterwards fold the block LayerNorms into the QKV / fc1
 *                                                              GEMM epilogues (1) or run separate LayerNorm kernels (0)
 * "gemm_balanced_tiles"        0 | 1                 1         equal-cost N tiles (1408 = 2 x 256 + 4 x 224) vs 256-wide tiles + narrow tail;
 *                                                              regroups the LayerNorm-fold partial sums (fp32 summation order)
 * "gemm_dynamic_schedule"      0 | 1                 1         ViT GEMM tiles from an atomic counter (in sequence order) vs static round-robin;
 *                                                              bit-identical results
 * "gemm_resid_prefetch_chunks" 0..3                  0         fp32-residual epilogues L2-prefetch their residual k chunks ahead (no gain)
 * Returns HB_OK or HB_ERR_INVALID (unknown key / value). */
HB_API int hb_debug_set(const char* key, int value);
/* Host-only: the column tiling the GEMM kernel uses for an N-wide output (first column and width of up to `cap` tiles);
 * returns the number of tiles or < 0.  For verification. */
HB_API int hb_gemm_n_tiling(int N, int cta_group, int balanced, int* n0, int* width, int cap);

#ifdef __cplusplus
}
#endif
#endif /* HIREST_B200_DEBUG_H_ */
