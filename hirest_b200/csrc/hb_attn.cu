// hb_attn.cu — fused multi-head attention for the EVA ViT-g/14 block on sm_100a (tcgen05 + TMEM).
//
// Replaces EVA_clip/vit_model.py:127-147: reshape/permute of qkv, q*scale, q@k^T, softmax, @v,
// transpose/reshape — without materialising the [B,16,257,257] score tensor.
//
// One CTA per (frame, head).  257 tokens = 2 x 128 query rows + 1: tokens 0..255 run on the tensor
// cores (S = Q K^T as two 128x256 UMMA tiles in TMEM, O = P V as two 128x96 tiles), token 256 is the
// "extra" token: its key/value column is folded into the softmax / output by the softmax threads
// (rank-1 update), and its query row is computed by one CUDA-core warp.  head_dim 88 is padded to 96
// (zero chunk) for the K=16 UMMA steps.
//
//   warps 0-3 : loaders (cp.async), then softmax + output for query tile 0 (one thread per row)
//   warps 4-7 : loaders (cp.async), then softmax + output for query tile 1
//   warp  8   : TMEM allocation + single-thread MMA issue
// Query row 256 is shared out over the same 256 threads: thread t scores key t from smem before P overwrites K,
// a block reduction gives its softmax, and its P.V product is accumulated while the tensor core runs P.V of the
// two big tiles.
//
// Shared memory (SWIZZLE_128B slabs, 128 B per row, 8-row groups 1024 B apart):
//   Q region 64 KiB: tile g, slab s (d 0..63 | d 64..95)           -> later overwritten by P tile 1
//   K region 64 KiB: slab s, 256 key rows                           -> later overwritten by P tile 0
//   V region 64 KiB: slab s, 256 key rows (MN-major B operand of P·V)
// q arrives pre-scaled by head_dim^-0.5 (QKV GEMM epilogue).
#include "hb_attn.cuh"
#include "hb_gemm.cuh"
#include "hb_ptx.cuh"

namespace hb {
namespace {

constexpr int T_TOK = 257;
constexpr int TQ = 256;       // tokens handled on tensor cores (queries and keys)
constexpr int DH = 88;
constexpr int NCHUNK = 11;    // 16-byte chunks per head row
constexpr int ATT_THREADS = 288;
constexpr uint32_t Q_OFF = 0, K_OFF = 65536, V_OFF = 131072, MISC_OFF = 196608;
constexpr uint32_t ATT_SMEM = MISC_OFF + 5632 + 128 + 1024;
constexpr float LOG2E = 1.4426950408889634f;

// MN-major SWIZZLE_128B descriptor (B operand = V[key][d], d contiguous): 64-element (128 B) rows along N,
// 8-key groups 1024 B apart (SBO), 64-wide d slabs `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void sts16(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

#ifdef HB_ATTN_TIMING
#define TSTAMP(i) do { if (threadIdx.x == 0 && p.timing) p.timing[blockIdx.x * 16 + (i)] = clock64(); } while (0)
#else
#define TSTAMP(i) do { } while (0)
#endif

__global__ void __launch_bounds__(ATT_THREADS, 1) vit_attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  float* kx = reinterpret_cast<float*>(smem + MISC_OFF);        // [96] key of token 256
  float* vx = kx + 96;                                          // [96] value of token 256
  float* qx = vx + 96;                                          // [96] query of token 256 (pre-scaled)
  float* red = qx + 96;                                         // [32] block-reduction scratch
  float* px = red + 32;                                         // [256] softmax numerators of query 256
  float* accx = px + 256;                                       // [8][96] per-warp partial outputs of query 256
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MISC_OFF + 5632);
  uint64_t* bar_load = bars;       // count 1 + 12 TMA boxes of 16 KiB (complete_tx)
  uint64_t* bar_s = bars + 1;      // [2] count 1 (commit)
  uint64_t* bar_p = bars + 3;      // [2] count 128
  uint64_t* bar_o = bars + 5;      // [2] count 1 (commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  TSTAMP(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int ldq = 3 * p.H * DH;
  const __nv_bfloat16* qg = p.qkv + static_cast<size_t>(b) * T_TOK * ldq + h * DH;
  const __nv_bfloat16* kg = qg + p.H * DH;
  const __nv_bfloat16* vg = kg + p.H * DH;
  __nv_bfloat16* og = p.out + static_cast<size_t>(b) * T_TOK * (p.H * DH) + h * DH;
  const int ldo = p.H * DH;

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(bar_load, 1);
      tma_prefetch_desc(&tmQKV);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_s[i], 1);
        mbar_init(&bar_p[i], 128);
        mbar_init(&bar_o[i], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  TSTAMP(1);

  if (warp < 8) {
    // ------------------------------------------------------------------ load phase
    const int tid = threadIdx.x;
    // Q/K/V tiles arrive by TMA (issued by warp 8); meanwhile fetch token 256's q/k/v rows (extra key / extra query).
    float xa = 0.f, xb = 0.f;
    if (tid < 96) {
      if (tid < DH) {
        xa = __bfloat162float(kg[static_cast<size_t>(TQ) * ldq + tid]);
        xb = __bfloat162float(qg[static_cast<size_t>(TQ) * ldq + tid]);
      }
    } else if (tid >= 128 && tid < 224) {
      if (tid - 128 < DH) xa = __bfloat162float(vg[static_cast<size_t>(TQ) * ldq + (tid - 128)]);
    }
    if (tid < 96) { kx[tid] = xa; qx[tid] = xb; }
    else if (tid >= 128 && tid < 224) vx[tid - 128] = xa;
    TSTAMP(2);
    mbar_wait(bar_load, 0);
    // all 256 threads must see kx/qx/vx before the extra-key / extra-query dot products
    asm volatile("bar.sync 1, 256;" ::: "memory");

    TSTAMP(3);
    const int g = warp >> 2;           // query tile
    const int wq = warp & 3;           // TMEM lane quarter
    const int r = wq * 32 + lane;      // row within the tile
    // score against the extra key (token 256): s_x = q_row . k_x  (q is pre-scaled)
    float s_x;
    {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const uint32_t qrow = sbase + Q_OFF + static_cast<uint32_t>(g * 2) * 16384u + static_cast<uint32_t>(r >> 3) * 1024u +
                            static_cast<uint32_t>(r & 7) * 128u;
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        const int slab = ch >> 3, cs = ch & 7;
        const uint4 v = lds16(qrow + static_cast<uint32_t>(slab) * 16384u + (static_cast<uint32_t>(cs ^ (r & 7)) << 4));
        const float4 k0 = *reinterpret_cast<const float4*>(kx + ch * 8), k1 = *reinterpret_cast<const float4*>(kx + ch * 8 + 4);
        a0 = fmaf(bf_lo(v.x), k0.x, a0); a1 = fmaf(bf_hi(v.x), k0.y, a1); a2 = fmaf(bf_lo(v.y), k0.z, a2); a3 = fmaf(bf_hi(v.y), k0.w, a3);
        a0 = fmaf(bf_lo(v.z), k1.x, a0); a1 = fmaf(bf_hi(v.z), k1.y, a1); a2 = fmaf(bf_lo(v.w), k1.z, a2); a3 = fmaf(bf_hi(v.w), k1.w, a3);
      }
      s_x = (a0 + a1) + (a2 + a3);
    }
    // Query row 256, step A: thread tid scores key tid (K row tid from smem), block-wide softmax statistics.
    float e_t;
    {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const uint32_t krow = sbase + K_OFF + static_cast<uint32_t>(tid >> 3) * 1024u + static_cast<uint32_t>(tid & 7) * 128u;
#pragma unroll
      for (int ch = 0; ch < NCHUNK; ++ch) {
        const int slab = ch >> 3, cs = ch & 7;
        const uint4 v = lds16(krow + static_cast<uint32_t>(slab) * 32768u + (static_cast<uint32_t>(cs ^ (tid & 7)) << 4));
        const float4 q0 = *reinterpret_cast<const float4*>(qx + ch * 8), q1 = *reinterpret_cast<const float4*>(qx + ch * 8 + 4);
        a0 = fmaf(bf_lo(v.x), q0.x, a0); a1 = fmaf(bf_hi(v.x), q0.y, a1); a2 = fmaf(bf_lo(v.y), q0.z, a2); a3 = fmaf(bf_hi(v.y), q0.w, a3);
        a0 = fmaf(bf_lo(v.z), q1.x, a0); a1 = fmaf(bf_hi(v.z), q1.y, a1); a2 = fmaf(bf_lo(v.w), q1.z, a2); a3 = fmaf(bf_hi(v.w), q1.w, a3);
      }
      e_t = (a0 + a1) + (a2 + a3);
    }
    float e_self;  // query 256 . key 256: lanes over d, warp reduction (every warp computes it redundantly)
    {
      float a = qx[lane] * kx[lane] + qx[lane + 32] * kx[lane + 32] + qx[lane + 64] * kx[lane + 64];  // d >= 88 are zeros
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      e_self = a;
    }
    {
      float wm = e_t;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
      if (lane == 0) red[warp] = wm;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float mx = e_self;
#pragma unroll
    for (int w = 0; w < 8; ++w) mx = fmaxf(mx, red[w]);
    const float px_t = exp2f((e_t - mx) * LOG2E);
    const float px_self = exp2f((e_self - mx) * LOG2E);
    px[tid] = px_t;
    {
      float ws = px_t;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
      if (lane == 0) red[8 + warp] = ws;
    }
    TSTAMP(4);
    // Both S tiles must be complete (K and Q smem dead) and every thread done reading Q/K before P overwrites them.
    mbar_wait(&bar_s[0], 0);
    mbar_wait(&bar_s[1], 0);
    tc_fence_after();
    asm volatile("bar.sync 1, 256;" ::: "memory");

    TSTAMP(5);
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(g * 256);
    float m = s_x;
    {
      uint32_t v[32];
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x32(t_s + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
      }
    }
    TSTAMP(6);
    const float m2 = m * LOG2E;
    float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
    const uint32_t prow = sbase + (g == 0 ? K_OFF : Q_OFF) + static_cast<uint32_t>(r >> 3) * 1024u +
                          static_cast<uint32_t>(r & 7) * 128u;
    {
      uint32_t v[32];
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x32(t_s + c * 32, v);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float ex;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(__uint_as_float(v[j]), LOG2E, -m2)));
          e[j] = ex;
        }
        // (a polynomial exp2 on the FMA pipe for every other element was tried and was slower: this pass is
        //  issue/latency-bound with two warps per scheduler, not MUFU-bound)
#pragma unroll
        for (int j = 0; j < 32; j += 4) { sum0 += e[j]; sum1 += e[j + 1]; sum2 += e[j + 2]; sum3 += e[j + 3]; }
        const uint32_t slab_addr = prow + static_cast<uint32_t>(c >> 1) * 16384u;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 q;
          q.x = pack_bf16x2(e[8 * jj + 0], e[8 * jj + 1]);
          q.y = pack_bf16x2(e[8 * jj + 2], e[8 * jj + 3]);
          q.z = pack_bf16x2(e[8 * jj + 4], e[8 * jj + 5]);
          q.w = pack_bf16x2(e[8 * jj + 6], e[8 * jj + 7]);
          const int cs = (c & 1) * 4 + jj;
          sts16(slab_addr + (static_cast<uint32_t>(cs ^ (r & 7)) << 4), q);
        }
      }
    }
    const float p_x = exp2f(s_x * LOG2E - m2);
    const float sum = (sum0 + sum1) + (sum2 + sum3) + p_x;
    const float inv = 1.0f / sum;
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(&bar_p[g]);
    TSTAMP(7);

    // Query row 256, step B (overlaps the tensor-core P.V): warp w accumulates keys [32w, 32w+32), lanes over d.
    {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      const bool has2 = lane < DH - 64;
      const uint32_t voff0 = static_cast<uint32_t>(lane >> 3), vin = static_cast<uint32_t>(lane & 7) * 2u;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const int key = warp * 32 + j;
        const float pj = px[key];
        const uint32_t vrow = sbase + V_OFF + static_cast<uint32_t>(key >> 3) * 1024u + static_cast<uint32_t>(key & 7) * 128u;
        const uint32_t sw = static_cast<uint32_t>(key & 7);
        uint16_t u0, u1, u2 = 0;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u0) : "r"(vrow + (((voff0) ^ sw) << 4) + vin));
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u1) : "r"(vrow + (((voff0 + 4u) ^ sw) << 4) + vin));
        if (has2) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u2) : "r"(vrow + 32768u + (((voff0) ^ sw) << 4) + vin));
        a0 += pj * __uint_as_float(static_cast<uint32_t>(u0) << 16);
        a1 += pj * __uint_as_float(static_cast<uint32_t>(u1) << 16);
        a2 += pj * __uint_as_float(static_cast<uint32_t>(u2) << 16);
      }
      accx[warp * 96 + lane] = a0;
      accx[warp * 96 + 32 + lane] = a1;
      accx[warp * 96 + 64 + lane] = a2;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid < DH) {
      float tot = px_self;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += red[8 + w];
      float o = px_self * vx[tid];
#pragma unroll
      for (int w = 0; w < 8; ++w) o += accx[w * 96 + tid];
      og[static_cast<size_t>(TQ) * ldo + tid] = __float2bfloat16(o / tot);
    }

    TSTAMP(8);
    // ------------------------------------------------------------------ output
    mbar_wait(&bar_o[g], 0);
    TSTAMP(9);
    tc_fence_after();
    // O rows (88 bf16 = 176 B) are staged in this tile's dead P region and written out as contiguous 16-byte chunks
    // (thread-per-row global stores cost 32 L1 wavefronts per instruction).
    const uint32_t ostage = sbase + (g == 0 ? K_OFF : Q_OFF);
    {
      uint32_t v[32];
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        tmem_ld_32x32(t_s + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int d0 = c * 32 + jj * 8;
          if (d0 < DH) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (__uint_as_float(v[jj * 8 + j]) + p_x * vx[d0 + j]) * inv;
            uint4 q;
            q.x = pack_bf16x2(o[0], o[1]);
            q.y = pack_bf16x2(o[2], o[3]);
            q.z = pack_bf16x2(o[4], o[5]);
            q.w = pack_bf16x2(o[6], o[7]);
            sts16(ostage + static_cast<uint32_t>(r) * 176u + static_cast<uint32_t>(d0) * 2u, q);
          }
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory");
    {
      const int tl = tid & 127;
      __nv_bfloat16* otile = og + static_cast<size_t>(g * 128) * ldo;
#pragma unroll
      for (int i = 0; i < NCHUNK; ++i) {
        const int c = tl + 128 * i;           // chunk index within the tile: row = c / 11, 16-byte chunk = c % 11
        const int row = c / NCHUNK, ch = c - row * NCHUNK;
        const uint4 q = lds16(ostage + static_cast<uint32_t>(c) * 16u);
        *reinterpret_cast<uint4*>(otile + static_cast<size_t>(row) * ldo + ch * 8) = q;
      }
    }
    tc_fence_before();
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issue
    if (lane == 0) {
      // TMA: 12 boxes of [128 rows x 64 d] (d >= 88 zero-filled by the tensor map's bounds) straight into the
      // SWIZZLE_128B slabs.  Tensor map dims: (d = 88, which*H + head, row).
      mbar_arrive_expect_tx(bar_load, 12u * 16384u);
      const int row0 = b * T_TOK;
#pragma unroll
      for (int w = 0; w < 3; ++w) {      // 0 = Q, 1 = K, 2 = V (same [tile/half][slab] order for all three)
        uint8_t* region = smem + (w == 0 ? Q_OFF : (w == 1 ? K_OFF : V_OFF));
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int slab = 0; slab < 2; ++slab) {
            // Q: tile-major [tile][slab] x 16 KiB; K/V: [slab] x 32 KiB with the two 128-row halves back to back
            const uint32_t off = (w == 0) ? static_cast<uint32_t>(half * 2 + slab) * 16384u
                                          : static_cast<uint32_t>(slab) * 32768u + static_cast<uint32_t>(half) * 16384u;
            tma_load_3d(region + off, &tmQKV, bar_load, slab * 64, w * p.H + h, row0 + half * 128);
          }
        }
      }
      mbar_wait(bar_load, 0);
      tc_fence_after();
      const uint32_t idesc_s = umma_idesc_bf16(128, 256);
      for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int slab = k >> 2, kk = k & 3;
          const uint64_t ad = umma_desc_sw128(sbase + Q_OFF + static_cast<uint32_t>(g * 2 + slab) * 16384u + kk * 32);
          const uint64_t bd = umma_desc_sw128(sbase + K_OFF + static_cast<uint32_t>(slab) * 32768u + kk * 32);
          umma_bf16<1>(tmem_base + g * 256, ad, bd, idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit<1>(&bar_s[g]);
      }
      const uint32_t idesc_o = umma_idesc_bf16(128, 96) | (1u << 16);  // B operand MN-major
      for (int g = 0; g < 2; ++g) {
        mbar_wait(&bar_p[g], 0);
        tc_fence_after();
        const uint32_t pbase = sbase + (g == 0 ? K_OFF : Q_OFF);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint64_t ad = umma_desc_sw128(pbase + static_cast<uint32_t>(k >> 2) * 16384u + (k & 3) * 32);
          const uint64_t bd = umma_desc_sw128_mn(sbase + V_OFF + static_cast<uint32_t>(k) * 2048u, 32768u);
          umma_bf16<1>(tmem_base + g * 256, ad, bd, idesc_o, k > 0 ? 1u : 0u);
        }
        umma_commit<1>(&bar_o[g]);
      }
    }
    __syncwarp();
  }

  TSTAMP(10);
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc<1>(tmem_base, 512);
  TSTAMP(11);
}

}  // namespace

int vit_attn_launch(const AttnParams& p, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0) return -3;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vit_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  CUtensorMap tm;
  const uint64_t dims[3] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(3 * p.H), static_cast<uint64_t>(p.B) * T_TOK};
  const uint64_t strides[2] = {static_cast<uint64_t>(DH) * 2, static_cast<uint64_t>(3 * p.H * DH) * 2};
  const uint32_t box[3] = {64, 1, 128};
  if (int r = make_tmap_bf16_3d(&tm, p.qkv, dims, strides, box)) return r;
  vit_attn_kernel<<<p.B * p.H, ATT_THREADS, ATT_SMEM, stream>>>(tm, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hb
