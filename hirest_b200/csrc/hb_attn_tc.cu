// hb_attn_tc.cu — fp32-accurate head_dim-64 attention on the tensor cores (tcgen05), for the small sequence models whose INTEGER
// outputs must match the fp32 reference: MomentModel temporal encoder (clip4caption/modules/module_visual.py:154-180, T ~ 300..2048
// frames), precise EVA-CLIP text tower (EVA_clip/eva_model.py:132,146; 77 tokens, causal).
//
// Same semantics as small_attn_f32_kernel (hb_attn_small.cu: fp32 q / k / v in, fp32 out, mask modes incl. the reference's
// "-10000 on every key" fp32-add quirk), which spent 13 SM-cycles per (query, key) pair on CUDA cores and was 80 % of the
// moment-segmentation time.  Here both contractions run as split-bf16 UMMAs with fp32 accumulation in TMEM:
//   S = q k^T   : q' = [lo | hi | hi], k' = [hi | lo | hi] (K = 3 x 64), so q'.k' = lo.hi + hi.lo + hi.hi — every product of two bf16
//                 is exact in fp32; only the lo.lo term (2^-16 relative) is dropped, as in the split GEMMs of hb_api.cu;
//   O = p v     : p = hi + lo written back into the S columns it came from as packed bf16 (TS-form UMMA, A operand from TMEM),
//                 v = hi + lo as MN-major B operands: p_lo.v_hi + p_hi.v_lo + p_hi.v_hi;
//   softmax in fp32 (ex2.approx on exactly formed differences), online over 128-key tiles (running max / sum per query row; O rescaled in TMEM when the max moves).
// A prologue kernel splits q / k / v once per call into bf16 workspace buffers laid out [row, head, {192 | 192 | 64 | 64}], so the
// main kernel's operands arrive by TMA in SWIZZLE_128B slabs.  One CTA per (batch, head, 128-query tile); warps 0-7 softmax (two
// threads per query row, 64 keys of the tile each: with one thread per row the exp pass of a single warp per scheduler was
// the critical path, 7 us per key tile), warp 8 MMA issue, warp 9 TMA producer; K / V double-buffered, S double-buffered in TMEM.
#include "hb_attn.cuh"
#include "hb_gemm.cuh"
#include "hb_ptx.cuh"

#include <cmath>

namespace hb {
namespace {

constexpr int DH = 64;
constexpr int TC_THREADS = 320;   // warps 0-7 softmax (2 threads per query row), warp 8 MMA issue, warp 9 TMA producer
constexpr uint32_t SLAB = 16384;                       // 128 rows x 128 B
constexpr uint32_t Q_OFF = 0;                          // 3 slabs
constexpr uint32_t KV_OFF = 3 * SLAB;                  // per stage: K 3 slabs | V_hi | V_lo
constexpr uint32_t STAGE = 5 * SLAB;
constexpr uint32_t XM_OFF = KV_OFF + 2 * STAGE;          // [2 halves][128] floats: row maxima / row sums exchanged between the halves
constexpr uint32_t BAR_OFF = XM_OFF + 1024;
constexpr uint32_t TC_SMEM = BAR_OFF + 256 + 1024;
constexpr float LOG2E = 1.4426950408889634f;

enum { B_Q = 0, B_KV_FULL0, B_KV_FULL1, B_KV_FREE0, B_KV_FREE1, B_S0, B_S1, B_P0, B_P1, B_O, B_N };

__device__ __forceinline__ uint64_t desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
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
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- prologue: fp32 q / k / v -> bf16 split operands -----------------------------------------------------------------------
//   Qs [B*Tq, H, 192] = [lo | hi | hi]   Ks [B*Tk, H, 192] = [hi | lo | hi]   Vh / Vl [B*Tk, H, 64] = hi / lo
//   PAIR (kernel v2): Qs / Ks [rows, H, 128] = [hi | lo] — the three product terms pick their slabs through the UMMA descriptors
template <bool PAIR>
__global__ void __launch_bounds__(256) attn_split_kernel(const SmallAttnF32Params p, __nv_bfloat16* __restrict__ qs, __nv_bfloat16* __restrict__ ks,
                                                         __nv_bfloat16* __restrict__ vh, __nv_bfloat16* __restrict__ vl) {
  constexpr int W = PAIR ? 128 : 192;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // one thread per (row, head, 4 dims)
  const long long nq = static_cast<long long>(p.B) * p.Tq * p.H * (DH / 4);
  const long long nk = static_cast<long long>(p.B) * p.Tk * p.H * (DH / 4);
  auto split4 = [](float4 x, uint2& hi, uint2& lo) {
    const __nv_bfloat16 h0 = __float2bfloat16(x.x), h1 = __float2bfloat16(x.y), h2 = __float2bfloat16(x.z), h3 = __float2bfloat16(x.w);
    hi.x = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
    hi.y = pack_bf16x2(__bfloat162float(h2), __bfloat162float(h3));
    lo.x = pack_bf16x2(x.x - __bfloat162float(h0), x.y - __bfloat162float(h1));
    lo.y = pack_bf16x2(x.z - __bfloat162float(h2), x.w - __bfloat162float(h3));
  };
  if (t < nq) {
    const int d4 = static_cast<int>(t % (DH / 4));
    const int h = static_cast<int>((t / (DH / 4)) % p.H);
    const long long row = t / (static_cast<long long>(DH / 4) * p.H);
    const long long b = row / p.Tq, i = row - b * p.Tq;
    const float4 x = *reinterpret_cast<const float4*>(p.q + b * p.bsq + i * p.ldq + h * DH + d4 * 4);
    uint2 hi, lo;
    split4(x, hi, lo);
    uint2* dst = reinterpret_cast<uint2*>(qs + (row * p.H + h) * W + d4 * 4);
    if constexpr (PAIR) {
      dst[0] = hi;
      dst[16] = lo;    // + 64 bf16
    } else {
      dst[0] = lo;
      dst[16] = hi;
      dst[32] = hi;    // + 128 bf16
    }
  }
  if (t < nk) {
    const int d4 = static_cast<int>(t % (DH / 4));
    const int h = static_cast<int>((t / (DH / 4)) % p.H);
    const long long row = t / (static_cast<long long>(DH / 4) * p.H);
    const long long b = row / p.Tk, i = row - b * p.Tk;
    const float4 xk = *reinterpret_cast<const float4*>(p.k + b * p.bsk + i * p.ldk + h * DH + d4 * 4);
    const float4 xv = *reinterpret_cast<const float4*>(p.v + b * p.bsv + i * p.ldv + h * DH + d4 * 4);
    uint2 hi, lo;
    split4(xk, hi, lo);
    uint2* dk = reinterpret_cast<uint2*>(ks + (row * p.H + h) * W + d4 * 4);
    dk[0] = hi;
    dk[16] = lo;
    if constexpr (!PAIR) dk[32] = hi;
    split4(xv, hi, lo);
    *reinterpret_cast<uint2*>(vh + (row * p.H + h) * DH + d4 * 4) = hi;
    *reinterpret_cast<uint2*>(vl + (row * p.H + h) * DH + d4 * 4) = lo;
  }
}

// ---- main kernel ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) small_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                      const __grid_constant__ CUtensorMap tmVh,
                                                                      const __grid_constant__ CUtensorMap tmVl, const SmallAttnF32Params p,
                                                                      int q_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_N + 1);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  const int qt = blockIdx.x % q_tiles;
  const int bh = blockIdx.x / q_tiles;
  const int b = bh / p.H, h = bh - b * p.H;
  const int q0 = qt * 128;
  int n_kt = (p.Tk + 127) / 128;
  if (p.mask_mode == 1) n_kt = min(n_kt, (min(p.Tq, q0 + 128) + 127) / 128);   // hard causal: no key beyond the tile's last query

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(bars + B_Q, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(bars + B_KV_FULL0 + s, 1);
        mbar_init(bars + B_KV_FREE0 + s, 1);
        mbar_init(bars + B_S0 + s, 1);
        mbar_init(bars + B_P0 + s, 8);
      }
      mbar_init(bars + B_O, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(bars + B_Q, 3 * SLAB);
      for (int s = 0; s < 3; ++s) tma_load_3d(smem + Q_OFF + s * SLAB, &tmQ, bars + B_Q, s * 64, h, b * p.Tq + q0);
    }
    __syncwarp();
    for (int j = 0; j < n_kt; ++j) {
      const int st = j & 1;
      if (j >= 2) mbar_wait(bars + B_KV_FREE0 + st, ((j >> 1) - 1) & 1);
      if (elect_one()) {
        uint8_t* base = smem + KV_OFF + st * STAGE;
        const int krow = (b / p.kv_div) * p.Tk + j * 128;
        mbar_arrive_expect_tx(bars + B_KV_FULL0 + st, 5 * SLAB);
        for (int s = 0; s < 3; ++s) tma_load_3d(base + s * SLAB, &tmK, bars + B_KV_FULL0 + st, s * 64, h, krow);
        tma_load_3d(base + 3 * SLAB, &tmVh, bars + B_KV_FULL0 + st, 0, h, krow);
        tma_load_3d(base + 4 * SLAB, &tmVl, bars + B_KV_FULL0 + st, 0, h, krow);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issue (whole warp walks, one elected lane issues)
    const uint32_t sb = __shfl_sync(0xffffffffu, sbase, 0);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_s = umma_idesc_bf16(128, 128);
    const uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) MN-major
    auto issue_s = [&](int j) {
      const int st = j & 1;
      mbar_wait(bars + B_KV_FULL0 + st, (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t kb = sb + KV_OFF + st * STAGE;
#pragma unroll
        for (int ks = 0; ks < 12; ++ks) {
          const uint64_t ad = umma_desc_sw128(sb + Q_OFF + (ks >> 2) * SLAB + (ks & 3) * 32);
          const uint64_t bd = umma_desc_sw128(kb + (ks >> 2) * SLAB + (ks & 3) * 32);
          umma_bf16<1>(tb + st * 128, ad, bd, idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit<1>(bars + B_S0 + st);
      }
      __syncwarp();
    };
    mbar_wait(bars + B_Q, 0);
    issue_s(0);
    for (int j = 0; j < n_kt; ++j) {
      const int st = j & 1;
      if (j + 1 < n_kt) issue_s(j + 1);   // runs behind P.V of tile j-1 on the tensor pipe (in issue order), ahead of tile j's softmax
      mbar_wait(bars + B_P0 + st, (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t pb = tb + st * 128;
        const uint32_t vhb = sb + KV_OFF + st * STAGE + 3 * SLAB, vlb = vhb + SLAB;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 keys per step; P chunk c = ks / 2: hi at [32c, 32c+16), lo at [32c+16, 32c+32)
          const uint32_t a_hi = pb + (ks >> 1) * 32 + (ks & 1) * 8, a_lo = a_hi + 16;
          const uint64_t b_hi = desc_sw128_mn(vhb + ks * 2048, SLAB), b_lo = desc_sw128_mn(vlb + ks * 2048, SLAB);
          umma_bf16_ts(tb + 256, a_lo, b_hi, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          umma_bf16_ts(tb + 256, a_hi, b_lo, idesc_o, 1u);
          umma_bf16_ts(tb + 256, a_hi, b_hi, idesc_o, 1u);
        }
        umma_commit<1>(bars + B_O);
        umma_commit<1>(bars + B_KV_FREE0 + st);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax: two threads per query row
    const int hf = warp >> 2;           // keys [64 hf, 64 hf + 64) of every tile; O columns [32 hf, 32 hf + 32)
    const int r = (warp & 3) * 32 + lane;
    const int i = q0 + r;               // query index inside the sequence
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float* xm = reinterpret_cast<float*>(smem + XM_OFF);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_kt; ++j) {
      const int st = j & 1;
      const uint32_t t_s = t_row + st * 128 + hf * 64;
      const int key0 = j * 128 + hf * 64;
      mbar_wait(bars + B_S0 + st, (j >> 1) & 1);
      tc_fence_after();
      // logits of this thread's 64 keys; masked keys -> -inf.  exp(x - m) is evaluated as exp2((x - m) * log2 e): the difference is
      // taken FIRST, in natural units (exact for the -10000-shifted logits, which are multiples of 2^-10), then scaled
      auto logit2 = [&](uint32_t raw, int key) -> float {
        float s = __uint_as_float(raw) * p.scale;
        if (p.mask_mode == 2) {
          s = s + p.mask_const;                               // fp32 add: quantises the logit exactly as the reference's mask add
          if (p.causal_soft && key > i) s = s + (-10000.0f);
        }
        if (key >= p.Tk || (p.mask_mode == 1 && key > i)) s = -INFINITY;
        return s;
      };
      float tmax = -INFINITY;
      {
        uint32_t v[32];
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32(t_s + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) tmax = fmaxf(tmax, logit2(v[k], key0 + c * 32 + k));
        }
      }
      xm[hf * 128 + r] = tmax;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tmax = fmaxf(tmax, xm[(hf ^ 1) * 128 + r]);
      // (Measured and dropped: "lazy" rescaling — keep the running maximum until a tile exceeds it by > 8, so that warps whose rows
      // did not move skip this O round trip and the wait for the previous P.V: no change at 64 x 300 tokens, 23.9 vs 23.7 ms.)
      const float m_new = fmaxf(m_run, tmax);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;        // a row with no visible key yet: every p is exp2(-inf) = 0
      float alpha;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"((m_run - m_use) * LOG2E));   // 0 while m_run = -inf
      if (j > 0) {
        // O holds sum_j' p v relative to m_run: bring it to m_new before this tile's P.V accumulates on top
        mbar_wait(bars + B_O, (j - 1) & 1);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32(t_row + 256 + hf * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
        tmem_st_32x32(t_row + 256 + hf * 32, o);
      }
      float tsum = 0.f;
      {
        uint32_t v[32];
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32(t_s + c * 32, v);
          tmem_ld_wait();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float p0, p1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"((logit2(v[2 * k], key0 + c * 32 + 2 * k) - m_use) * LOG2E));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"((logit2(v[2 * k + 1], key0 + c * 32 + 2 * k + 1) - m_use) * LOG2E));
            tsum += p0 + p1;
            const __nv_bfloat16 h0 = __float2bfloat16(p0), h1 = __float2bfloat16(p1);
            hi[k] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
            lo[k] = pack_bf16x2(p0 - __bfloat162float(h0), p1 - __bfloat162float(h1));
          }
          tmem_st_32x16(t_s + c * 32, hi);        // over S columns this thread has just consumed
          tmem_st_32x16(t_s + c * 32 + 16, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + B_P0 + st);
      l_run = l_run * alpha + tsum;
      m_run = m_new;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // both halves have read xm before the next tile overwrites it
    }
    // ---- output: this thread's 32 columns of O / (row sum over both halves)
    xm[hf * 128 + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.0f / (l_run + xm[(hf ^ 1) * 128 + r]);
    mbar_wait(bars + B_O, (n_kt - 1) & 1);
    tc_fence_after();
    float* og = p.out + static_cast<long long>(b) * p.bso + static_cast<long long>(i) * p.ldo + h * DH + hf * 32;
    uint32_t o[32];
    tmem_ld_32x32(t_row + 256 + hf * 32, o);
    tmem_ld_wait();
    if (i < p.Tq) {
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        *reinterpret_cast<float4*>(og + k) = make_float4(__uint_as_float(o[k]) * inv, __uint_as_float(o[k + 1]) * inv,
                                                         __uint_as_float(o[k + 2]) * inv, __uint_as_float(o[k + 3]) * inv);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc<1>(tmem_base, 512);
}

// ---- main kernel, v2 (default): two CTAs per SM -------------------------------------------------------------------------------
// v1 keeps one CTA per SM (208 KiB of smem, all 512 TMEM columns), so a CTA's prologue (barrier init, TMEM allocation, first TMA
// round trip), its serial S -> softmax -> P.V chain over only ceil(Tk / 128) key tiles and its epilogue are all exposed: 15.6 us
// per CTA at 64 x 300 tokens for ~5 us of tensor + MUFU work.  v2 shrinks a CTA to 97 KiB and 256 TMEM columns so that two are
// resident and fill each other's bubbles:
//   * Q / K are staged as [hi | lo] (2 slabs instead of the 3-slab [lo|hi|hi] / [hi|lo|hi] images): the three product terms are
//     formed by pointing the A / B descriptors at (q_lo, k_hi), (q_hi, k_lo), (q_hi, k_hi) — same MMAs, same accumulation order,
//     bit-identical S; a third less operand traffic;
//   * K and V are single-buffered with separate full / free barriers: K(j+1) is fetched as soon as S(j) has been computed (during
//     tile j's softmax), V(j+1) as soon as P.V(j) has; S is single-buffered in TMEM (S(j+1) is issued behind P.V(j) on the in-order
//     tensor pipe, so it cannot overwrite P(j) early).
constexpr uint32_t Q2_OFF = 0;                 // 2 slabs: hi | lo
constexpr uint32_t K2_OFF = 2 * SLAB;          // 2 slabs: hi | lo
constexpr uint32_t V2_OFF = 4 * SLAB;          // V_hi | V_lo
constexpr uint32_t XM2_OFF = 6 * SLAB;
constexpr uint32_t BAR2_OFF = XM2_OFF + 1024;
constexpr uint32_t TC2_SMEM = BAR2_OFF + 256 + 1024;
enum { C_Q = 0, C_K_FULL, C_K_FREE, C_V_FULL, C_V_FREE, C_S, C_P, C_O, C_N };

__global__ void __launch_bounds__(TC_THREADS, 2) small_attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                       const __grid_constant__ CUtensorMap tmVh,
                                                                       const __grid_constant__ CUtensorMap tmVl, const SmallAttnF32Params p,
                                                                       int q_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR2_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C_N + 1);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  const int qt = blockIdx.x % q_tiles;
  const int bh = blockIdx.x / q_tiles;
  const int b = bh / p.H, h = bh - b * p.H;
  const int q0 = qt * 128;
  int n_kt = (p.Tk + 127) / 128;
  if (p.mask_mode == 1) n_kt = min(n_kt, (min(p.Tq, q0 + 128) + 127) / 128);   // hard causal: no key beyond the tile's last query

  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < C_N; ++i) mbar_init(bars + i, i == C_P ? 8 : 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 256);   // S / P [0, 128), O [128, 192): two CTAs share the SM's 512 columns
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(bars + C_Q, 2 * SLAB);
      for (int s = 0; s < 2; ++s) tma_load_3d(smem + Q2_OFF + s * SLAB, &tmQ, bars + C_Q, s * 64, h, b * p.Tq + q0);
    }
    __syncwarp();
    for (int j = 0; j < n_kt; ++j) {
      const int krow = (b / p.kv_div) * p.Tk + j * 128;
      if (j >= 1) mbar_wait(bars + C_K_FREE, (j - 1) & 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bars + C_K_FULL, 2 * SLAB);
        for (int s = 0; s < 2; ++s) tma_load_3d(smem + K2_OFF + s * SLAB, &tmK, bars + C_K_FULL, s * 64, h, krow);
      }
      __syncwarp();
      if (j >= 1) mbar_wait(bars + C_V_FREE, (j - 1) & 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bars + C_V_FULL, 2 * SLAB);
        tma_load_3d(smem + V2_OFF, &tmVh, bars + C_V_FULL, 0, h, krow);
        tma_load_3d(smem + V2_OFF + SLAB, &tmVl, bars + C_V_FULL, 0, h, krow);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issue (whole warp walks, one elected lane issues)
    const uint32_t sb = __shfl_sync(0xffffffffu, sbase, 0);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_s = umma_idesc_bf16(128, 128);
    const uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);   // B (= V) MN-major
    auto issue_s = [&](int j) {
      mbar_wait(bars + C_K_FULL, j & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 12; ++ks) {   // terms in v1's order: q_lo.k_hi, q_hi.k_lo, q_hi.k_hi (slab 0 = hi, slab 1 = lo)
          const int term = ks >> 2;
          const uint32_t q_slab = (term == 0) ? 1u : 0u, k_slab = (term == 1) ? 1u : 0u;
          const uint64_t ad = umma_desc_sw128(sb + Q2_OFF + q_slab * SLAB + (ks & 3) * 32);
          const uint64_t bd = umma_desc_sw128(sb + K2_OFF + k_slab * SLAB + (ks & 3) * 32);
          umma_bf16<1>(tb, ad, bd, idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit<1>(bars + C_S);
        umma_commit<1>(bars + C_K_FREE);
      }
      __syncwarp();
    };
    mbar_wait(bars + C_Q, 0);
    issue_s(0);
    for (int j = 0; j < n_kt; ++j) {
      mbar_wait(bars + C_P, j & 1);
      mbar_wait(bars + C_V_FULL, j & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vhb = sb + V2_OFF, vlb = vhb + SLAB;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 keys per step; P chunk c = ks / 2: hi at [32c, 32c+16), lo at [32c+16, 32c+32)
          const uint32_t a_hi = tb + (ks >> 1) * 32 + (ks & 1) * 8, a_lo = a_hi + 16;
          const uint64_t b_hi = desc_sw128_mn(vhb + ks * 2048, SLAB), b_lo = desc_sw128_mn(vlb + ks * 2048, SLAB);
          umma_bf16_ts(tb + 128, a_lo, b_hi, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          umma_bf16_ts(tb + 128, a_hi, b_lo, idesc_o, 1u);
          umma_bf16_ts(tb + 128, a_hi, b_hi, idesc_o, 1u);
        }
        umma_commit<1>(bars + C_O);
        umma_commit<1>(bars + C_V_FREE);
      }
      __syncwarp();
      if (j + 1 < n_kt) issue_s(j + 1);   // behind P.V(j) on the in-order tensor pipe: P(j) is consumed before S(j+1) lands on it
    }
  } else {
    // ------------------------------------------------------------------ softmax: two threads per query row (as v1)
    const int hf = warp >> 2;           // keys [64 hf, 64 hf + 64) of every tile; O columns [32 hf, 32 hf + 32)
    const int r = (warp & 3) * 32 + lane;
    const int i = q0 + r;               // query index inside the sequence
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t t_s = t_row + hf * 64, t_o = t_row + 128 + hf * 32;
    float* xm = reinterpret_cast<float*>(smem + XM2_OFF);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_kt; ++j) {
      const int key0 = j * 128 + hf * 64;
      mbar_wait(bars + C_S, j & 1);
      tc_fence_after();
      auto logit2 = [&](uint32_t raw, int key) -> float {   // see v1
        float s = __uint_as_float(raw) * p.scale;
        if (p.mask_mode == 2) {
          s = s + p.mask_const;
          if (p.causal_soft && key > i) s = s + (-10000.0f);
        }
        if (key >= p.Tk || (p.mask_mode == 1 && key > i)) s = -INFINITY;
        return s;
      };
      float tmax = -INFINITY;
      {
        uint32_t v[32];
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32(t_s + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) tmax = fmaxf(tmax, logit2(v[k], key0 + c * 32 + k));
        }
      }
      xm[hf * 128 + r] = tmax;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tmax = fmaxf(tmax, xm[(hf ^ 1) * 128 + r]);
      const float m_new = fmaxf(m_run, tmax);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      float alpha;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"((m_run - m_use) * LOG2E));
      if (j > 0) {
        mbar_wait(bars + C_O, (j - 1) & 1);
        tc_fence_after();
        uint32_t o[32];
        tmem_ld_32x32(t_o, o);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
        tmem_st_32x32(t_o, o);
      }
      float tsum = 0.f;
      {
        uint32_t v[32];
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          tmem_ld_32x32(t_s + c * 32, v);
          tmem_ld_wait();
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float p0, p1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"((logit2(v[2 * k], key0 + c * 32 + 2 * k) - m_use) * LOG2E));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"((logit2(v[2 * k + 1], key0 + c * 32 + 2 * k + 1) - m_use) * LOG2E));
            tsum += p0 + p1;
            const __nv_bfloat16 h0 = __float2bfloat16(p0), h1 = __float2bfloat16(p1);
            hi[k] = pack_bf16x2(__bfloat162float(h0), __bfloat162float(h1));
            lo[k] = pack_bf16x2(p0 - __bfloat162float(h0), p1 - __bfloat162float(h1));
          }
          tmem_st_32x16(t_s + c * 32, hi);
          tmem_st_32x16(t_s + c * 32 + 16, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + C_P);
      l_run = l_run * alpha + tsum;
      m_run = m_new;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // both halves have read xm before the next tile overwrites it
    }
    xm[hf * 128 + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.0f / (l_run + xm[(hf ^ 1) * 128 + r]);
    mbar_wait(bars + C_O, (n_kt - 1) & 1);
    tc_fence_after();
    float* og = p.out + static_cast<long long>(b) * p.bso + static_cast<long long>(i) * p.ldo + h * DH + hf * 32;
    uint32_t o[32];
    tmem_ld_32x32(t_o, o);
    tmem_ld_wait();
    if (i < p.Tq) {
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        *reinterpret_cast<float4*>(og + k) = make_float4(__uint_as_float(o[k]) * inv, __uint_as_float(o[k + 1]) * inv,
                                                         __uint_as_float(o[k + 2]) * inv, __uint_as_float(o[k + 3]) * inv);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc<1>(tmem_base, 256);
}

}  // namespace

size_t small_attn_tc_workspace(int B, int H, int Tq, int Tk) {
  const size_t rq = static_cast<size_t>(B) * Tq * H, rk = static_cast<size_t>(B) * Tk * H;
  return (rq * 192 + rk * 192 + rk * 64 * 2) * 2 + 4 * 256;
}

namespace { int g_tc_version = 2; }
void small_attn_tc_set_version(int v) { g_tc_version = (v == 1) ? 1 : 2; }

int small_attn_tc_launch(const SmallAttnF32Params& p, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0 || p.Tq <= 0 || p.Tk <= 0) return -3;
  if (p.kv_div != 1) return -7;   // shared K / V across beams: decode steps (Tq = 1) stay on the CUDA-core kernel
  if ((p.ldq | p.ldk | p.ldv | p.ldo) % 4 || (p.bsq | p.bsk | p.bsv | p.bso) % 4) return -7;
  if (workspace == nullptr || workspace_bytes < small_attn_tc_workspace(p.B, p.H, p.Tq, p.Tk)) return -8;
  const bool v2 = g_tc_version == 2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(small_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(small_attn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const size_t rq = static_cast<size_t>(p.B) * p.Tq, rk = static_cast<size_t>(p.B) * p.Tk;
  const size_t W = v2 ? 128 : 192;   // bf16 per (row, head) of the Q / K images (the workspace is sized for the wider one)
  auto align256 = [](size_t x) { return (x + 255) / 256 * 256; };
  uint8_t* w = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(w);
  w += align256(rq * p.H * W * 2);
  __nv_bfloat16* ks = reinterpret_cast<__nv_bfloat16*>(w);
  w += align256(rk * p.H * W * 2);
  __nv_bfloat16* vh = reinterpret_cast<__nv_bfloat16*>(w);
  w += align256(rk * p.H * 64 * 2);
  __nv_bfloat16* vl = reinterpret_cast<__nv_bfloat16*>(w);
  const long long nthreads = static_cast<long long>(rq > rk ? rq : rk) * p.H * (DH / 4);
  if (v2) attn_split_kernel<true><<<static_cast<unsigned>((nthreads + 255) / 256), 256, 0, stream>>>(p, qs, ks, vh, vl);
  else attn_split_kernel<false><<<static_cast<unsigned>((nthreads + 255) / 256), 256, 0, stream>>>(p, qs, ks, vh, vl);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  CUtensorMap tq, tk, tvh, tvl;
  const uint32_t box[3] = {64, 1, 128};
  {
    const uint64_t dq[3] = {W, static_cast<uint64_t>(p.H), rq}, dk[3] = {W, static_cast<uint64_t>(p.H), rk};
    const uint64_t sw[2] = {W * 2, static_cast<uint64_t>(p.H) * W * 2};
    const uint64_t dv[3] = {64, static_cast<uint64_t>(p.H), rk};
    const uint64_t s64[2] = {64 * 2, static_cast<uint64_t>(p.H) * 64 * 2};
    if (int r = make_tmap_bf16_3d(&tq, qs, dq, sw, box)) return r;
    if (int r = make_tmap_bf16_3d(&tk, ks, dk, sw, box)) return r;
    if (int r = make_tmap_bf16_3d(&tvh, vh, dv, s64, box)) return r;
    if (int r = make_tmap_bf16_3d(&tvl, vl, dv, s64, box)) return r;
  }
  const int q_tiles = (p.Tq + 127) / 128;
  if (v2) small_attn_tc2_kernel<<<p.B * p.H * q_tiles, TC_THREADS, TC2_SMEM, stream>>>(tq, tk, tvh, tvl, p, q_tiles);
  else small_attn_tc_kernel<<<p.B * p.H * q_tiles, TC_THREADS, TC_SMEM, stream>>>(tq, tk, tvh, tvl, p, q_tiles);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hb
