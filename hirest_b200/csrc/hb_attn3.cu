// hb_attn3.cu — ViT attention, third generation: ONE persistent CTA per SM, software-pipelined over (frame, head) items.
//
// Same math as hb_attn2.cu / hb_attn.cu (EVA_clip/vit_model.py:127-147: softmax(q k^T) v over 257 tokens, head_dim 88, q pre-scaled
// by the QKV GEMM epilogue).  v2 (one CTA per 128-query tile, two CTAs per SM) moved exactly its algorithmic bytes but ran at
// 28 % of HBM peak: every CTA started with an exposed DRAM-latency wait for its 96 KiB of operands, paid TMEM alloc / barrier
// init / dealloc, read K and V twice per (frame, head), had 10 warps per SM to hide latency with, and the CTA that also owned the
// extra query row (token 256) did ~2x the CUDA-core work of its sibling.  Here:
//   * a CTA loops over items (frame, head); a producer warp TMA-loads Q/K of item i+1 as soon as the S = Q.K^T MMAs of item i
//     have retired and V of item i+1 as soon as the P.V MMAs have; K / V are read once per item, not once per query tile;
//   * 16 softmax warps: both 128-query tiles of an item are in flight (TMEM 2 x 256 columns), TWO threads per query row
//     (keys 0..127 / 128..255); the MMA warp polls one S -> P.V sequence per tile, and the tiles take turns in the MUFU-bound
//     exp pass, so they settle half a period apart: one tile's exps overlap the other's P.V / output / dot products;
//   * each thread computes its NEXT item's CUDA-core dot product while the tensor core runs its tile's P.V;
//   * the row sum comes from the tensor core: column 88 of V (zero padding of the 96-wide operand) is set to 1.0, so
//     O[:, 88] = sum_j bf16(P[:, j]) — exactly the normalisation of the bf16 P the tensor core multiplies, and no FADD chain;
//   * the extra query row's P.V runs on the tensor core as O_x^T = V^T p_x (A = V as an MN-major operand, M = d, N = 16, into 16
//     TMEM columns that are free once the tile's softmax has consumed them), per tile over that tile's 128 keys relative to
//     the tile's own maximum, merged flash-attention style at the end; its 256 scores (half-1 threads: k_row . q_x) and the
//     extra KEY's score of every row (half-0 threads: q_row . k_x) are CUDA-core dot products;
//   * P is packed bf16 in TMEM (TS-form UMMA), V is an MN-major B operand (no transposes), as in v2;
//   * the output leaves through a dense bf16 staging tile per 128-query tile and ONE TMA store (cp.async.bulk.tensor, box 88 x 128):
//     thread-per-row 16-byte global stores touched 32 lines per instruction and backed the LSU pipe up for ~5k cycles per item.
// Measured (1024 frames x 16 heads, profiles/r02_attention_v3.txt): 1.02 ms (v2) -> 0.73 ms per layer = 4.0 TB/s of algorithmic
// traffic (61 % of the measured HBM peak).  What bounds it now: K / V are single-buffered (shared memory is full) and have two
// consumers half a period apart, so a load window is ~half an item period while one SM's share of HBM bandwidth needs ~2/3 of
// it; MUFU (4096 cycles per item) and the tensor pipe (~4300) are the next floors.  L2-prefetching the next item's boxes
// (-DHB_A3_L2_PREFETCH) slows the TMEM loads of the max pass 4x and costs 20 %; lock-step tiles (-DHB_A3_NO_TURNS) and dot
// products at the end of the item (AttnParams::dots_late, debug key "attention_dots_late") are within 2 % of the default.
// TMEM per tile (256 columns): S fp32 [0,256); P (bf16x2) keys 0..127 -> [0,64) (ascending, behind the reader), keys 128..255 ->
// [192,256) (that thread walks its S columns in DESCENDING order, so it too only overwrites columns it has consumed);
// O d 0..95 -> [64,160) (column 152 = row sum); extra-query partial output [160,176).
// Shared memory: Q 2 x 24 KiB | K 48 KiB (columns 0..63 as SWIZZLE_128B slabs, columns 64..95 as SWIZZLE_64B slabs: the 96-wide
// K dimension of S without 64 dead bytes per row) | V 64 KiB (SWIZZLE_128B, MN-major operand) | p_x operand 8 KiB |
// output staging 2 x 22 KiB | vectors.
#include "hb_attn.cuh"
#include "hb_gemm.cuh"
#include "hb_ptx.cuh"

namespace hb {
namespace {

constexpr int T_TOK = 257;
constexpr int TQ = 256;
constexpr int DH = 88;
constexpr int A3_SOFTMAX_THREADS = 512;
constexpr int A3_THREADS = A3_SOFTMAX_THREADS + 64;   // + warp 16: TMA producer, warp 17: MMA issue (+ TMEM alloc)
constexpr uint32_t Q_TILE = 24576, Q_S1 = 16384;     // per 128-query tile: slab 0 (16 KiB, SW128) | slab 1 (8 KiB, SW64)
constexpr uint32_t K_S1 = 32768;                      // K: slab 0 (256 rows x 128 B) | slab 1 (256 rows x 64 B)
constexpr uint32_t ST_TILE = 128 * DH * 2;            // dense bf16 output staging tile [128][88]
constexpr uint32_t Q_OFF = 0, K_OFF = 2 * Q_TILE, V_OFF = K_OFF + 49152, PX_OFF = V_OFF + 65536, ST_OFF = PX_OFF + 8192,
                   MISC_OFF = ST_OFF + 2 * ST_TILE;
constexpr uint32_t XT_FLOATS = 2 * 3 * 96, XE_FLOATS = 2 * 256, XS_FLOATS = 2 * 256, XM_FLOATS = 2 * 2 * 128, RED_FLOATS = 32;
constexpr uint32_t OX_FLOATS = 2 * 128;   // [2 item parities][96 partial outputs of the extra query from tile 0 | m_0 | sum_0]
constexpr uint32_t MISC_BYTES = (XT_FLOATS + XE_FLOATS + XS_FLOATS + XM_FLOATS + RED_FLOATS + OX_FLOATS) * 4;
constexpr uint32_t A3_SMEM = MISC_OFF + MISC_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;
constexpr float LOG2E = 1.4426950408889634f;

// Debug builds (-DHB_ATTN_TIMING, tools/attn3_timing.cu): clock64 stamps per phase, [cta][item][role 0..3][16];
// role 0 / 1 = lane 0 of the first softmax warp of tile 0 / 1, role 2 = MMA warp, role 3 = producer warp.
#ifdef HB_ATTN_TIMING
#define A3_STAMP(role, k)                                                                                                   \
  do {                                                                                                                      \
    if (p.timing != nullptr && lane == 0 && it < 64)                                                                        \
      p.timing[((static_cast<size_t>(blockIdx.x) * 64 + it) * 4 + (role)) * 16 + (k)] = clock64();                          \
  } while (0)
#else
#define A3_STAMP(role, k) do { } while (0)
#endif

enum Bar : int { B_QK_FULL = 0, B_V_FULL, B_QK_FREE, B_V_FREE, B_OX, B_S0, B_S1, B_P0, B_P1, B_O0, B_O1, B_TF0, B_TF1, B_PX0, B_PX1, B_COUNT };

__device__ __forceinline__ uint64_t desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]   (TS form: the A operand is read from tensor memory)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Partial dot of one smem row of Q / K (88 bf16: columns 0..63 in a SWIZZLE_128B slab, row at shared address `row0`; columns
// 64..87 in a SWIZZLE_64B slab, row at `row1`) with an fp32 vector in smem: 16-byte chunks [CH0, CH1) of the row's 11.
// (ld.shared through 32-bit shared addresses: C++ loads through the re-aligned dynamic-smem pointer compile to generic LD.E.128
// with 64-bit address arithmetic and spilled pointers.)
template <int CH0, int CH1>
__device__ __forceinline__ float dot_row_part(uint32_t row0, uint32_t row1, int row, const float* vec) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int ch = CH0; ch < CH1; ++ch) {
    const uint32_t addr = (ch < 8) ? row0 + (static_cast<uint32_t>(ch ^ (row & 7)) << 4)
                                   : row1 + (static_cast<uint32_t>((ch - 8) ^ ((row >> 1) & 3)) << 4);
    const uint4 x = lds16(addr);
    const float4 k0 = *reinterpret_cast<const float4*>(vec + ch * 8), k1 = *reinterpret_cast<const float4*>(vec + ch * 8 + 4);
    a0 = fmaf(bf_lo(x.x), k0.x, a0); a1 = fmaf(bf_hi(x.x), k0.y, a1); a2 = fmaf(bf_lo(x.y), k0.z, a2); a3 = fmaf(bf_hi(x.y), k0.w, a3);
    a0 = fmaf(bf_lo(x.z), k1.x, a0); a1 = fmaf(bf_hi(x.z), k1.y, a1); a2 = fmaf(bf_lo(x.w), k1.z, a2); a3 = fmaf(bf_hi(x.w), k1.w, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(A3_THREADS, 1) vit_attn3_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQK1,
                                                                  const __grid_constant__ CUtensorMap tmOut, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  float* xt = reinterpret_cast<float*>(smem + MISC_OFF);   // [2 buffers][3: k, v, q of token 256][96]
  float* xe = xt + XT_FLOATS;                              // [2 halves][256 keys]  partial extra-query scores
  float* xs = xe + XE_FLOATS;                              // [2 halves][256 rows]  partial extra-key scores
  float* xm = xs + XS_FLOATS;                              // [2 tiles][2 halves][128] partial row maxima
  float* red = xm + XM_FLOATS;                             // per tile [16]: 8 warp maxima of e, 4 warp sums of p_x
  float* oxs = red + RED_FLOATS;                           // [2][128] tile 0's share of the extra query row (see OX_FLOATS)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MISC_OFF + MISC_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT + 1);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int n_items = p.B * p.H;
  const int ldq = 3 * p.H * DH;
  const int ldo = p.H * DH;

  if (warp == 17) {
    if (lane == 0) {
      mbar_init(bars + B_QK_FULL, 1);
      mbar_init(bars + B_V_FULL, 1);
      mbar_init(bars + B_QK_FREE, 2 + 16);   // one tcgen05.commit per S tile + every softmax warp (CUDA-core reads of Q / K rows)
      mbar_init(bars + B_V_FREE, 2);         // one tcgen05.commit per tile's P.V group
      mbar_init(bars + B_OX, 3);             // tile 0's warps 0..2 parked their share of the extra query row
      for (int t = 0; t < 2; ++t) {
        mbar_init(bars + B_PX0 + t, 4);      // the tile's half-1 warps wrote p_x of their 128 keys
        mbar_init(bars + B_S0 + t, 1);
        mbar_init(bars + B_P0 + t, 8);       // the tile's P is in TMEM (all 8 warps)
        mbar_init(bars + B_O0 + t, 1);
        mbar_init(bars + B_TF0 + t, 8);
      }
      fence_mbar_init();
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmQK1);
      tma_prefetch_desc(&tmOut);
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ================================================================== producer: TMA loads + the extra token's vectors
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int b = item / p.H, h = item - b * p.H;
      const int row0 = b * T_TOK;
      A3_STAMP(3, 0);
      // token 256's key / value / query (3 x 88 bf16, plain loads issued before the wait) -> fp32 vectors in smem
      const __nv_bfloat16* xrow = p.qkv + static_cast<size_t>(row0 + TQ) * ldq + h * DH;
      float xv[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int idx = lane + 32 * i;            // 0..287: which = idx / 96 -> 0 k, 1 v, 2 q ; d = idx % 96
        const int which = idx / 96, d = idx - which * 96;
        const int src = (which == 0 ? 1 : (which == 1 ? 2 : 0)) * p.H * DH + d;
        xv[i] = (d < DH) ? __bfloat162float(xrow[src]) : 0.f;
      }
      if (it > 0) mbar_wait(bars + B_QK_FREE, (it - 1) & 1);
      A3_STAMP(3, 1);
      float* xb = xt + (it & 1) * 288;
#pragma unroll
      for (int i = 0; i < 9; ++i) xb[lane + 32 * i] = xv[i];
      __syncwarp();
      if (elect_one()) {
        mbar_arrive_expect_tx(bars + B_QK_FULL, 4u * 16384u + 4u * 8192u);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tma_load_3d(smem + Q_OFF + half * Q_TILE, &tmQKV, bars + B_QK_FULL, 0, h, row0 + half * 128);
          tma_load_3d(smem + Q_OFF + half * Q_TILE + Q_S1, &tmQK1, bars + B_QK_FULL, 64, h, row0 + half * 128);
          tma_load_3d(smem + K_OFF + half * 16384, &tmQKV, bars + B_QK_FULL, 0, p.H + h, row0 + half * 128);
          tma_load_3d(smem + K_OFF + K_S1 + half * 8192, &tmQK1, bars + B_QK_FULL, 64, p.H + h, row0 + half * 128);
        }
      }
      __syncwarp();
      A3_STAMP(3, 2);
      if (it > 0) mbar_wait(bars + B_V_FREE, (it - 1) & 1);
      A3_STAMP(3, 3);
      if (elect_one()) {
        mbar_arrive_expect_tx(bars + B_V_FULL, 4u * 16384u);
#pragma unroll
        for (int slab = 0; slab < 2; ++slab)
#pragma unroll
          for (int half = 0; half < 2; ++half)
            tma_load_3d(smem + V_OFF + slab * 32768 + half * 16384, &tmQKV, bars + B_V_FULL, slab * 64, 2 * p.H + h, row0 + half * 128);
      }
      __syncwarp();
#ifdef HB_A3_L2_PREFETCH   // measured: harmful (0.73 -> 0.88 ms per layer; the TMEM loads of the max pass slow down 4x)
      // Pull the NEXT item's boxes into L2 now.  At the HBM roofline one SM's share of the bandwidth (6.5 TB/s / 148 = 44 GB/s)
      // needs the whole item period to deliver an item's 160 KiB, but the shared-memory buffers are only free for part of it
      // (Q / K from "S retired" to the next item's dot products, V from "P.V retired" to the next P.V): with the lines already
      // in L2 the TMA loads complete in L2 time and DRAM streams all the time.
      const int nitem = item + static_cast<int>(gridDim.x);
      if (nitem < n_items && elect_one()) {
        const int nb = nitem / p.H, nh = nitem - nb * p.H, nrow0 = nb * T_TOK;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tma_prefetch_3d(&tmQKV, 0, nh, nrow0 + half * 128);
          tma_prefetch_3d(&tmQK1, 64, nh, nrow0 + half * 128);
          tma_prefetch_3d(&tmQKV, 0, p.H + nh, nrow0 + half * 128);
          tma_prefetch_3d(&tmQK1, 64, p.H + nh, nrow0 + half * 128);
          tma_prefetch_3d(&tmQKV, 0, 2 * p.H + nh, nrow0 + half * 128);
          tma_prefetch_3d(&tmQKV, 64, 2 * p.H + nh, nrow0 + half * 128);
        }
      }
      __syncwarp();
#endif
    }
  } else if (warp == 17) {
    // ================================================================== MMA issue
    // The whole warp walks this sequence and one elected lane issues (warp-uniform descriptors, see hb_gemm.cu).
    const uint32_t sb = __shfl_sync(0xffffffffu, sbase, 0);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_s = umma_idesc_bf16(128, 256);
    const uint32_t idesc_o = umma_idesc_bf16(128, 96) | (1u << 16);    // B (= V) MN-major
    const uint32_t idesc_x = umma_idesc_bf16(128, 16) | (1u << 15);    // A (= V) MN-major, B = p_x K-major
    // Two independent sequences (one per 128-query tile): S_t(it) -> P.V_t(it) -> S_t(it+1) ...  The warp polls both, so a
    // tile never waits for the other tile's softmax (the tiles run out of phase: while one is in its MUFU-bound exp pass the
    // other one is storing its output / computing the extra-token dot products).
    const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    int it_t[2] = {0, 0}, stage[2] = {0, 0};
    int ones_it = -1;   // V's ones column is in place for this item
    const uint32_t bar0 = smem_u32(bars);
    const uint64_t t_start = globaltimer_ns();
    while (it_t[0] < my_items || it_t[1] < my_items) {
      bool progressed = false;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int it = it_t[t];
        if (it >= my_items) continue;
        const uint32_t ph = it & 1;
        if (stage[t] == 0) {
          if (!mbar_test_wait(bar0 + 8 * B_QK_FULL, ph)) continue;
          if (it > 0 && !mbar_test_wait(bar0 + 8 * (B_TF0 + t), (it - 1) & 1)) continue;   // the tile's TMEM columns are drained
          if (t == 0) { A3_STAMP(2, 0); A3_STAMP(2, 1); }
          A3_STAMP(2, 2 + 2 * t);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {   // K dimension 96 = 4 steps in the SWIZZLE_128B slabs + 2 in the SWIZZLE_64B slabs
              const uint32_t qa = sb + Q_OFF + static_cast<uint32_t>(t) * Q_TILE;
              const uint64_t ad = k < 4 ? umma_desc_sw128(qa + k * 32) : umma_desc_sw64(qa + Q_S1 + (k - 4) * 32);
              const uint64_t bd = k < 4 ? umma_desc_sw128(sb + K_OFF + k * 32) : umma_desc_sw64(sb + K_OFF + K_S1 + (k - 4) * 32);
              umma_bf16<1>(tb + t * 256, ad, bd, idesc_s, k > 0 ? 1u : 0u);
            }
            umma_commit<1>(bars + B_S0 + t);
            umma_commit<1>(bars + B_QK_FREE);
          }
          __syncwarp();
          A3_STAMP(2, 3 + 2 * t);
          stage[t] = 1;
          progressed = true;
        } else {
          // (P.V cannot start before the whole exp pass has finished: its accumulator lives in S columns [64,160), which the two
          // halves consume last.)
          if (!mbar_test_wait(bar0 + 8 * B_V_FULL, ph)) continue;
          if (!mbar_test_wait(bar0 + 8 * (B_P0 + t), ph)) continue;
          if (!mbar_test_wait(bar0 + 8 * (B_PX0 + t), ph)) continue;
          A3_STAMP(2, 7 + 2 * t);
          if (ones_it != it) {
            // ones column: V[key, 88] = 1.0 (bf16 0x3F80) in the zero padding of slab 1 -> O[:, 88] = row sum of P
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t key = static_cast<uint32_t>(lane + 32 * i);
              const uint32_t a = sb + V_OFF + 32768u + (key >> 3) * 1024u + (key & 7u) * 128u + ((3u ^ (key & 7u)) << 4);
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(static_cast<uint16_t>(0x3F80)) : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            ones_it = it;
          }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t tt = tb + t * 256;
#pragma unroll
            for (int k = 0; k < 16; ++k) {   // 16 keys per step: 8 packed TMEM columns of P, 16 smem rows (2048 B) of V
              const uint32_t pa = tt + static_cast<uint32_t>(k < 8 ? k * 8 : 192 + (k - 8) * 8);
              umma_bf16_ts(tt + 64, pa, desc_sw128_mn(sb + V_OFF + static_cast<uint32_t>(k) * 2048u, 32768u), idesc_o, k > 0 ? 1u : 0u);
            }
            // extra query, this tile's 128 keys: O_x^T [d, 0] = sum_key V[key, d] p_x[key]  (p_x relative to the tile's own maximum)
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
              const int k = t * 8 + k8;
              const uint64_t ad = desc_sw128_mn(sb + V_OFF + static_cast<uint32_t>(k) * 2048u, 32768u);
              const uint64_t bd = umma_desc_sw128(sb + PX_OFF + static_cast<uint32_t>(k >> 2) * 2048u + (k & 3) * 32);
              umma_bf16<1>(tt + 160, ad, bd, idesc_x, k8 > 0 ? 1u : 0u);
            }
            umma_commit<1>(bars + B_O0 + t);
            umma_commit<1>(bars + B_V_FREE);
          }
          __syncwarp();
          A3_STAMP(2, 8 + 2 * t);
          progressed = true;
          stage[t] = 0;
          it_t[t] = it + 1;
        }
      }
      if (!progressed) {
        __nanosleep(40);
        // a protocol bug must trap instead of hanging the GPU (same bound as mbar_wait, per launch)
        if (globaltimer_ns() - t_start > 20ull * HB_WAIT_TIMEOUT_NS) __trap();
      }
    }
  } else {
    // ================================================================== softmax / output: 2 threads per query row
    const int tile = warp >> 3;            // 0: rows 0..127, 1: rows 128..255
    const int hf = (warp >> 2) & 1;        // 0: keys 0..127, 1: keys 128..255
    const int wq = warp & 3;               // TMEM lane quarter
    const int r = wq * 32 + lane;          // row inside the tile
    const int row = tile * 128 + r;        // query row of this thread; also the KEY a half-1 thread scores for the extra query
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(tile * 256);
    const uint32_t t_s = t_row + static_cast<uint32_t>(hf * 128);
    // CUDA-core dot product of this thread: half 0 -> extra KEY's score of its query row (q_row . k_x, needed by both halves
    // after the first pass); half 1 -> extra QUERY's score of key `row` (k_row . q_x)
    const uint32_t drow0 = hf == 0 ? sbase + Q_OFF + static_cast<uint32_t>(tile * Q_TILE + (r >> 3) * 1024 + (r & 7) * 128)
                                   : sbase + K_OFF + static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128);
    const uint32_t drow1 = hf == 0 ? sbase + Q_OFF + static_cast<uint32_t>(tile * Q_TILE + Q_S1 + (r >> 3) * 512 + (r & 7) * 64)
                                   : sbase + K_OFF + K_S1 + static_cast<uint32_t>((row >> 3) * 512 + (row & 7) * 64);
    const int drow = hf == 0 ? r : row;
    const uint32_t stage_row = sbase + ST_OFF + static_cast<uint32_t>(tile) * ST_TILE + static_cast<uint32_t>(r) * (DH * 2);
    float* redt = red + tile * 16;
    bool store_pending = false;   // this thread issued a TMA store whose smem source has not been waited for yet
    const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
#ifdef HB_ATTN_TIMING
    const int trole = (warp == 0) ? 0 : (warp == 8 ? 1 : -1);
#define A3_SSTAMP(k) do { if (trole >= 0) A3_STAMP(trole, k); } while (0)
#else
#define A3_SSTAMP(k) do { } while (0)
#endif

    // dot(it): this thread's CUDA-core dot product for item `it` (+ the extra token's own score q_x . k_x, per warp)
    float dotv = 0.f, e_self_next = 0.f;
    auto dots = [&](int it) {
      mbar_wait(bars + B_QK_FULL, it & 1);
      const float* kx = xt + (it & 1) * 288;
      const float* qx = kx + 192;
      dotv = dot_row_part<0, 11>(drow0, drow1, drow, hf == 0 ? kx : qx);
      float es = qx[lane] * kx[lane] + qx[lane + 32] * kx[lane + 32] + (lane < DH - 64 ? qx[lane + 64] * kx[lane + 64] : 0.f);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
      e_self_next = es;
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + B_QK_FREE);   // this warp is done with the Q / K rows in shared memory
    };
    if (my_items > 0) dots(0);

    for (int it = 0; it < my_items; ++it) {
      const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const uint32_t ph = it & 1;
      const int b = item / p.H, h = item - b * p.H;
      __nv_bfloat16* og = p.out + static_cast<size_t>(b) * T_TOK * ldo + h * DH;
      const float* vx = xt + (it & 1) * 288 + 96;
      const float e_self = e_self_next;
      A3_SSTAMP(0);

      // ---- extra query, half-1 threads: softmax numerators over THIS tile's 128 keys relative to the tile's own maximum m_t
      // (the two tiles' partial results are merged at the end, flash-attention style, so the tiles never wait for each other)
      float m_t = 0.f;
      if (hf == 1) {
        const float e = dotv;
        float wm = e;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        if (lane == 0) redt[wq] = wm;
        named_bar(3 + tile, 128);
        m_t = fmaxf(fmaxf(redt[0], redt[1]), fmaxf(redt[2], redt[3]));
        // p_x rounded to bf16: the tensor core multiplies exactly these values, so the normaliser sums the same ones.
        // B operand of the O_x MMA: row 0 of a [16 x 256] K-major SWIZZLE_128B tile (row 0 is not permuted by the swizzle)
        const __nv_bfloat16 pkb = __float2bfloat16(ex2((e - m_t) * LOG2E));
        *reinterpret_cast<__nv_bfloat16*>(smem + PX_OFF + (row >> 6) * 2048 + (row & 63) * 2) = pkb;
        float ws = __bfloat162float(pkb);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
        if (lane == 0) redt[8 + wq] = ws;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + B_PX0 + tile);
      } else {
        xs[row] = dotv;   // extra key's score of this row, for both halves (read after the pair barrier below)
      }
      A3_SSTAMP(1);

      // ---- pass 1: row maximum over this thread's 128 keys
      mbar_wait(bars + B_S0 + tile, ph);
      A3_SSTAMP(2);
      tc_fence_after();
      float m = -INFINITY;
      {
        uint32_t v[32];
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld_32x32(t_s + c * 32, v);
          tmem_ld_wait();
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[j])); m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[j + 2])); m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
          }
          m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        }
      }
      xm[(tile * 2 + hf) * 128 + r] = m;
      // the previous item's TMA store has finished reading the staging tile before anyone of the pair writes it again
      if (store_pending) { tma_store_wait_read(); store_pending = false; }
      A3_SSTAMP(3);
      named_bar(1 + tile, 256);
      A3_SSTAMP(4);
      const float s_x = xs[row];
      m = fmaxf(fmaxf(m, xm[(tile * 2 + (hf ^ 1)) * 128 + r]), s_x);
      const float m2 = m * LOG2E;

      // ---- pass 2: P = exp(S - m) as packed bf16 back into TMEM, over S columns this thread has already consumed:
      // half 0 walks S[0,128) upwards and writes P chunk c at [16c, 16c+16); half 1 walks S[128,256) DOWNWARDS and writes P chunk
      // c at [192+16c, 192+16c+16) -- so [64,192) is free for O and the extra-query partial when the P.V MMAs start.
      // The pass is MUFU-bound (65536 ex2 per item at 16 per clock per SM = 4096 cycles), so the two tiles take turns: tile 1
      // starts its pass when tile 0 has finished, tile 0's next pass follows tile 1's; the tiles settle half a period apart.
#ifndef HB_A3_NO_TURNS
      if (tile == 1) mbar_wait(bars + B_P0, ph);
      else if (it > 0) mbar_wait(bars + B_P1, (it - 1) & 1);
#endif
      {
        uint32_t v[32];
        tmem_ld_32x32(t_s + (hf ? 3 : 0) * 32, v);
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c = hf ? 3 - cc : cc;
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            pk[j] = pack_bf16x2(ex2(fmaf(__uint_as_float(v[2 * j]), LOG2E, -m2)), ex2(fmaf(__uint_as_float(v[2 * j + 1]), LOG2E, -m2)));
          tmem_st_32x16(t_row + static_cast<uint32_t>(hf * 192 + c * 16), pk);
          if (cc < 3) tmem_ld_32x32(t_s + (hf ? 2 - cc : cc + 1) * 32, v);   // next chunk's S while this chunk's P store drains
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + B_P0 + tile);
      const float p_x = ex2(fmaf(s_x, LOG2E, -m2));
      A3_SSTAMP(5);

      // ---- while the tensor core runs this tile's P.V: the NEXT item's dot products (its Q / K landed long ago)
      const bool dots_late = (p.dots_late >> tile) & 1;   // warp-uniform
      if (!dots_late && it + 1 < my_items) dots(it + 1);
      A3_SSTAMP(6);

      // ---- output: (O + the extra key's rank-1 term) / row sum -> dense bf16 staging tile -> one TMA store per tile
      mbar_wait(bars + B_O0 + tile, ph);
      A3_SSTAMP(7);
      tc_fence_after();
      uint32_t ox_raw = 0;
      {
        uint32_t v[32], w[16];
        float inv, pxi;
        auto pack8 = [&](const uint32_t* src, int d0) {
          uint4 q;
          q.x = pack_bf16x2(fmaf(__uint_as_float(src[0]), inv, pxi * vx[d0 + 0]), fmaf(__uint_as_float(src[1]), inv, pxi * vx[d0 + 1]));
          q.y = pack_bf16x2(fmaf(__uint_as_float(src[2]), inv, pxi * vx[d0 + 2]), fmaf(__uint_as_float(src[3]), inv, pxi * vx[d0 + 3]));
          q.z = pack_bf16x2(fmaf(__uint_as_float(src[4]), inv, pxi * vx[d0 + 4]), fmaf(__uint_as_float(src[5]), inv, pxi * vx[d0 + 5]));
          q.w = pack_bf16x2(fmaf(__uint_as_float(src[6]), inv, pxi * vx[d0 + 6]), fmaf(__uint_as_float(src[7]), inv, pxi * vx[d0 + 7]));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + static_cast<uint32_t>(d0) * 2u), "r"(q.x), "r"(q.y),
                       "r"(q.z), "r"(q.w)
                       : "memory");
        };
        if (hf == 0) {   // d 0..47: columns [64,112); row sum: column 152
          tmem_ld_32x32(t_row + 64, v);
          tmem_ld_32x16(t_row + 96, w);
          const uint32_t su = tmem_ld_32x1(t_row + 152);
          tmem_ld_wait();
          inv = 1.0f / (__uint_as_float(su) + p_x);
          pxi = p_x * inv;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) pack8(v + jj * 8, jj * 8);
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) pack8(w + jj * 8, 32 + jj * 8);
        } else {         // d 48..63: columns [112,128); d 64..87: columns [128,152); row sum: column 152 = v[24]
          tmem_ld_32x16(t_row + 112, w);
          tmem_ld_32x32(t_row + 128, v);
          if (wq < 3) ox_raw = tmem_ld_32x1(t_row + 160);   // extra-query partial, lanes = d
          tmem_ld_wait();
          inv = 1.0f / (__uint_as_float(v[24]) + p_x);
          pxi = p_x * inv;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) pack8(w + jj * 8, 48 + jj * 8);
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) pack8(v + jj * 8, 64 + jj * 8);
        }
      }
      // (read before the TF arrive below: once both tiles have drained TMEM the producer may refill this xt buffer two items on)
      const float vx_r = (hf == 1 && r < DH) ? vx[r] : 0.f;
      // the tile's TMEM columns are drained: S of the next item may start while the store is being staged
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + B_TF0 + tile);
      fence_proxy_async_smem();
      named_bar(1 + tile, 256);
      if ((warp & 7) == 0 && elect_one()) {
        tma_store_3d(&tmOut, sbase + ST_OFF + static_cast<uint32_t>(tile) * ST_TILE, 0, h, b * T_TOK + tile * 128);
        tma_store_commit();
        store_pending = true;
      }
      A3_SSTAMP(8);
      // ---- extra query row (token 256): O_x^T partials in TMEM (lanes = d, column 160 of each tile), over each tile's 128 keys
      // relative to that tile's maximum.  Tile 0 parks (partial, m_0, sum_0) in smem; tile 1 merges both with the extra key's own
      // term and stores the row.  Half-1 warps with wq < 3: lanes cover d = 0..95.
      if (hf == 1 && wq < 3) {
        const float sum_t = (redt[8] + redt[9]) + (redt[10] + redt[11]);
        float* park = oxs + (it & 1) * 128;
        if (tile == 0) {
          park[r] = __uint_as_float(ox_raw);
          if (r == 0) { park[96] = m_t; park[97] = sum_t; }
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + B_OX);
        } else {
          mbar_wait(bars + B_OX, ph);
          const float m_0 = park[96], sum_0 = park[97];
          const float M = fmaxf(fmaxf(m_0, m_t), e_self);
          const float f0 = ex2((m_0 - M) * LOG2E), f1 = ex2((m_t - M) * LOG2E), fs = ex2((e_self - M) * LOG2E);
          const float tot = fmaf(sum_0, f0, fmaf(sum_t, f1, fs));
          if (r < DH) {
            const float o = fmaf(park[r], f0, fmaf(__uint_as_float(ox_raw), f1, fs * vx_r));
            og[static_cast<size_t>(TQ) * ldo + r] = __float2bfloat16(o / tot);
          }
        }
      }
      if (dots_late && it + 1 < my_items) dots(it + 1);
    }
    if (store_pending) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 17) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace

int vit_attn3_launch(const AttnParams& p, int num_sms, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0 || num_sms <= 0) return -3;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vit_attn3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  CUtensorMap tm, tm1, tmo;
  const uint64_t dims[3] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(3 * p.H), static_cast<uint64_t>(p.B) * T_TOK};
  const uint64_t strides[2] = {static_cast<uint64_t>(DH) * 2, static_cast<uint64_t>(3 * p.H * DH) * 2};
  const uint32_t box[3] = {64, 1, 128};
  if (int r = make_tmap_bf16_3d(&tm, p.qkv, dims, strides, box)) return r;
  const uint32_t box1[3] = {32, 1, 128};   // columns 64..95 of Q / K (88..95 zero-filled) as SWIZZLE_64B slabs
  if (int r = make_tmap_bf16_3d(&tm1, p.qkv, dims, strides, box1, 64)) return r;
  const uint64_t odims[3] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(p.H), static_cast<uint64_t>(p.B) * T_TOK};
  const uint64_t ostrides[2] = {static_cast<uint64_t>(DH) * 2, static_cast<uint64_t>(p.H * DH) * 2};
  const uint32_t obox[3] = {DH, 1, 128};   // dense [128][88] staging tile -> out[b*257 + tile*128 + r, h*88 + d]
  if (int r = make_tmap_bf16_3d(&tmo, p.out, odims, ostrides, obox, 0)) return r;
  const int items = p.B * p.H;
  vit_attn3_kernel<<<items < num_sms ? items : num_sms, A3_THREADS, A3_SMEM, stream>>>(tm, tm1, tmo, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hb
