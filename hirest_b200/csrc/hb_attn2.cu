// hb_attn2.cu — ViT attention, second generation: one CTA per (frame, head, 128-query tile), TWO CTAs resident per SM.
//
// Same math as hb_attn.cu (EVA_clip/vit_model.py:127-147) with a smaller footprint so that one CTA's TMA latency hides
// behind the other CTA's softmax (profiles/r01_attention_phases.txt: 6.5k of 25.5k cycles per CTA were exposed load wait):
//   * P never touches shared memory: the softmax threads write it back into TMEM as packed bf16 (tcgen05.st) over the S
//     columns they have already consumed, and P.V runs as a TS-form UMMA (A operand from TMEM);
//   * K is dead once S = Q.K^T has been computed, so V is TMA-loaded into the same 64 KiB region while the softmax runs;
//   * TMEM: 256 columns per CTA — S fp32 [0,256), P bf16x2 [0,128) in place, O fp32 [128,224).
// Shared memory: Q tile 32 KiB + K/V region 64 KiB + 4 KiB scratch = ~101 KiB  ->  2 CTAs / SM.
// Token 256 (257 = 2*128 + 1): its key/value is a rank-1 update in every row's softmax/output; its query row is computed by
// the g == 0 CTA of each (frame, head) on CUDA cores from the smem-resident K and V.
#include "hb_attn.cuh"
#include "hb_gemm.cuh"
#include "hb_ptx.cuh"

namespace hb {
namespace {

constexpr int T_TOK = 257;
constexpr int TQ = 256;
constexpr int DH = 88;
constexpr int NCHUNK = 11;
constexpr int A2_THREADS = 160;  // warps 0-3: softmax / output (one thread per query row); warp 4: TMEM alloc, TMA, MMA issue
constexpr uint32_t Q_OFF = 0, KV_OFF = 32768, MISC_OFF = 98304;
constexpr uint32_t A2_SMEM = MISC_OFF + 4096 + 128 + 1024;
constexpr float LOG2E = 1.4426950408889634f;

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

__device__ __forceinline__ void sts16(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// dot of one SWIZZLE_128B smem row (88 bf16 in two slabs, `slab_stride` bytes apart) with an fp32 vector in smem
__device__ __forceinline__ float dot_row(uint32_t row_addr, int row, uint32_t slab_stride, const float* vec) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int ch = 0; ch < NCHUNK; ++ch) {
    const int slab = ch >> 3, cs = ch & 7;
    const uint4 v = lds16(row_addr + static_cast<uint32_t>(slab) * slab_stride + (static_cast<uint32_t>(cs ^ (row & 7)) << 4));
    const float4 k0 = *reinterpret_cast<const float4*>(vec + ch * 8), k1 = *reinterpret_cast<const float4*>(vec + ch * 8 + 4);
    a0 = fmaf(bf_lo(v.x), k0.x, a0); a1 = fmaf(bf_hi(v.x), k0.y, a1); a2 = fmaf(bf_lo(v.y), k0.z, a2); a3 = fmaf(bf_hi(v.y), k0.w, a3);
    a0 = fmaf(bf_lo(v.z), k1.x, a0); a1 = fmaf(bf_hi(v.z), k1.y, a1); a2 = fmaf(bf_lo(v.w), k1.z, a2); a3 = fmaf(bf_hi(v.w), k1.w, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(A2_THREADS, 2) vit_attn2_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  float* kx = reinterpret_cast<float*>(smem + MISC_OFF);  // [96] key of token 256
  float* vx = kx + 96;                                    // [96] value of token 256
  float* qx = vx + 96;                                    // [96] query of token 256
  float* red = qx + 96;                                   // [16]
  float* px = red + 16;                                   // [256] softmax numerators of query 256
  float* accx = px + 256;                                 // [4][96]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MISC_OFF + 4096);
  uint64_t* bar_qk = bars;         // TMA: Q tile + K
  uint64_t* bar_v = bars + 1;      // TMA: V (into the K region)
  uint64_t* bar_s = bars + 2;      // S = Q.K^T committed
  uint64_t* bar_kfree = bars + 3;  // 128 softmax threads are done reading K / Q rows from smem
  uint64_t* bar_p = bars + 4;      // 128 threads wrote P into TMEM
  uint64_t* bar_o = bars + 5;      // O = P.V committed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int g = blockIdx.x & 1;
  const int bh = blockIdx.x >> 1;
  const int b = bh / p.H, h = bh - b * p.H;
  const int ldq = 3 * p.H * DH;
  const __nv_bfloat16* qg = p.qkv + static_cast<size_t>(b) * T_TOK * ldq + h * DH;
  const __nv_bfloat16* kg = qg + p.H * DH;
  const __nv_bfloat16* vg = kg + p.H * DH;
  __nv_bfloat16* og = p.out + static_cast<size_t>(b) * T_TOK * (p.H * DH) + h * DH;
  const int ldo = p.H * DH;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_kfree, 128);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      fence_mbar_init();
      tma_prefetch_desc(&tmQKV);
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA + MMA issue
    // The whole warp walks this sequence and one elected lane issues; addresses routed through shfl(…, 0) are known to be
    // warp-uniform, so descriptors live in uniform registers and every tcgen05.mma issues directly (with `if (lane == 0)`
    // each one was wrapped in an ELECT + R2UR.BROADCAST waterfall of ~25 instructions on the CTA's critical path).
    const uint32_t sb = __shfl_sync(0xffffffffu, sbase, 0);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int row0 = b * T_TOK;
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_qk, 6u * 16384u);
#pragma unroll
      for (int slab = 0; slab < 2; ++slab) {
        tma_load_3d(smem + Q_OFF + slab * 16384, &tmQKV, bar_qk, slab * 64, h, row0 + g * 128);
#pragma unroll
        for (int half = 0; half < 2; ++half)
          tma_load_3d(smem + KV_OFF + slab * 32768 + half * 16384, &tmQKV, bar_qk, slab * 64, p.H + h, row0 + half * 128);
      }
    }
    __syncwarp();
    // Optional (off by default): while this CTA's own Q / K boxes are in flight, pull the boxes of the CTA that will follow it
    // on this SM (one wave = 2 CTAs x #SMs later in block order) into L2.  A CTA lives ~10 us and starts with a DRAM-latency
    // wait on 96 KiB of operands (15 % of the kernel's stall samples sit on that barrier) — but measured, the extra TMA
    // traffic costs more than the shorter wait saves: 1.05 -> 1.14 ms per layer.
    if (p.prefetch_ahead > 0) {
      const int nb = static_cast<int>(blockIdx.x) + p.prefetch_ahead;
      if (nb < static_cast<int>(gridDim.x) && elect_one()) {
        const int g2 = nb & 1, bh2 = nb >> 1, b2 = bh2 / p.H, h2 = bh2 - b2 * p.H, r2 = b2 * T_TOK;
#pragma unroll
        for (int slab = 0; slab < 2; ++slab) {
          tma_prefetch_3d(&tmQKV, slab * 64, h2, r2 + g2 * 128);
          if (g2 == 0) {   // K and V are shared by the two query tiles of a (frame, head): the g = 0 CTA's successor fetches them
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              tma_prefetch_3d(&tmQKV, slab * 64, p.H + h2, r2 + half * 128);
              tma_prefetch_3d(&tmQKV, slab * 64, 2 * p.H + h2, r2 + half * 128);
            }
          }
        }
      }
      __syncwarp();
    }
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 256);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int slab = k >> 2, kk = k & 3;
        const uint64_t ad = umma_desc_sw128(sb + Q_OFF + static_cast<uint32_t>(slab) * 16384u + kk * 32);
        const uint64_t bd = umma_desc_sw128(sb + KV_OFF + static_cast<uint32_t>(slab) * 32768u + kk * 32);
        umma_bf16<1>(tb, ad, bd, idesc_s, k > 0 ? 1u : 0u);
      }
      umma_commit<1>(bar_s);
    }
    __syncwarp();
    // K is dead once the MMAs have retired and the extra-token dot products have read their rows: V takes its place
    mbar_wait(bar_s, 0);
    mbar_wait(bar_kfree, 0);
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_v, 4u * 16384u);
#pragma unroll
      for (int slab = 0; slab < 2; ++slab)
#pragma unroll
        for (int half = 0; half < 2; ++half)
          tma_load_3d(smem + KV_OFF + slab * 32768 + half * 16384, &tmQKV, bar_v, slab * 64, 2 * p.H + h, row0 + half * 128);
    }
    __syncwarp();
    mbar_wait(bar_p, 0);
    mbar_wait(bar_v, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc_o = umma_idesc_bf16(128, 96) | (1u << 16);  // B (= V) is MN-major
#pragma unroll
      for (int k = 0; k < 16; ++k) {   // 16 keys per step: 8 packed TMEM columns of P, 16 smem rows (2048 B) of V
        const uint64_t bd = desc_sw128_mn(sb + KV_OFF + static_cast<uint32_t>(k) * 2048u, 32768u);
        umma_bf16_ts(tb + 128, tb + static_cast<uint32_t>(k * 8), bd, idesc_o, k > 0 ? 1u : 0u);
      }
      umma_commit<1>(bar_o);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax / output: thread = query row
    const int tid = threadIdx.x;   // 0..127
    const int r = tid;
    const int wq = warp;           // TMEM lane quarter
    {
      float xa = 0.f, xb = 0.f, xc = 0.f;
      if (tid < DH) {
        xa = __bfloat162float(kg[static_cast<size_t>(TQ) * ldq + tid]);
        xc = __bfloat162float(vg[static_cast<size_t>(TQ) * ldq + tid]);
        if (g == 0) xb = __bfloat162float(qg[static_cast<size_t>(TQ) * ldq + tid]);
      }
      if (tid < 96) { kx[tid] = xa; vx[tid] = xc; qx[tid] = xb; }
    }
    mbar_wait(bar_qk, 0);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // extra key (token 256) against this thread's query row; q is pre-scaled
    const uint32_t qrow = sbase + Q_OFF + static_cast<uint32_t>(r >> 3) * 1024u + static_cast<uint32_t>(r & 7) * 128u;
    const float s_x = dot_row(qrow, r, 16384u, kx);
    float e0 = 0.f, e1 = 0.f, e_self = 0.f;
    if (g == 0) {   // extra query (token 256): this thread scores keys tid and tid + 128
      const uint32_t krow0 = sbase + KV_OFF + static_cast<uint32_t>(tid >> 3) * 1024u + static_cast<uint32_t>(tid & 7) * 128u;
      e0 = dot_row(krow0, tid, 32768u, qx);
      e1 = dot_row(krow0 + 16384u, tid, 32768u, qx);   // row tid + 128: 16 groups of 8 rows further
      float a = qx[lane] * kx[lane] + qx[lane + 32] * kx[lane + 32] + qx[lane + 64] * kx[lane + 64];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      e_self = a;
    }
    mbar_arrive(bar_kfree);
    float px_self = 0.f;
    if (g == 0) {
      float wm = fmaxf(e0, e1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
      if (lane == 0) red[warp] = wm;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float mx = fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), e_self);
      const float p0 = exp2f((e0 - mx) * LOG2E), p1 = exp2f((e1 - mx) * LOG2E);
      px_self = exp2f((e_self - mx) * LOG2E);
      px[tid] = p0;
      px[tid + 128] = p1;
      float ws = p0 + p1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
      if (lane == 0) red[8 + warp] = ws;
    }
    mbar_wait(bar_s, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    float m = s_x;
    {
      uint32_t v[32];
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
      }
    }
    const float m2 = m * LOG2E;
    float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
    {
      uint32_t v[32];
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x32(t_row + c * 32, v);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float ex;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(__uint_as_float(v[j]), LOG2E, -m2)));
          e[j] = ex;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) { sum0 += e[j]; sum1 += e[j + 1]; sum2 += e[j + 2]; sum3 += e[j + 3]; }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(e[2 * j], e[2 * j + 1]);
        // P (bf16 pairs) overwrites S columns that this thread has already consumed: [16c, 16c+16) is inside [0, 32c+32)
        tmem_st_32x16(t_row + c * 16, pk);
      }
    }
    const float p_x = exp2f(s_x * LOG2E - m2);
    const float inv = 1.0f / ((sum0 + sum1) + (sum2 + sum3) + p_x);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar_p);

    if (g == 0) {
      // extra query: P.V from smem while the tensor core runs the tile's P.V; warp w takes keys [64w, 64w+64), lanes over d
      asm volatile("bar.sync 1, 128;" ::: "memory");   // px[] of all 256 keys visible
      mbar_wait(bar_v, 0);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      const bool has2 = lane < DH - 64;
      const uint32_t voff0 = static_cast<uint32_t>(lane >> 3), vin = static_cast<uint32_t>(lane & 7) * 2u;
#pragma unroll 8
      for (int j = 0; j < 64; ++j) {
        const int key = warp * 64 + j;
        const float pj = px[key];
        const uint32_t vrow = sbase + KV_OFF + static_cast<uint32_t>(key >> 3) * 1024u + static_cast<uint32_t>(key & 7) * 128u;
        const uint32_t sw = static_cast<uint32_t>(key & 7);
        uint16_t u0, u1, u2 = 0;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u0) : "r"(vrow + (((voff0) ^ sw) << 4) + vin));
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u1) : "r"(vrow + (((voff0 + 4u) ^ sw) << 4) + vin));
        if (has2) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u2) : "r"(vrow + 32768u + (((voff0) ^ sw) << 4) + vin));
        a0 = fmaf(pj, __uint_as_float(static_cast<uint32_t>(u0) << 16), a0);
        a1 = fmaf(pj, __uint_as_float(static_cast<uint32_t>(u1) << 16), a1);
        a2 = fmaf(pj, __uint_as_float(static_cast<uint32_t>(u2) << 16), a2);
      }
      accx[warp * 96 + lane] = a0;
      accx[warp * 96 + 32 + lane] = a1;
      accx[warp * 96 + 64 + lane] = a2;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid < DH) {
        const float tot = px_self + ((red[8] + red[9]) + (red[10] + red[11]));
        const float o = px_self * vx[tid] + ((accx[tid] + accx[96 + tid]) + (accx[192 + tid] + accx[288 + tid]));
        og[static_cast<size_t>(TQ) * ldo + tid] = __float2bfloat16(o / tot);
      }
    }

    // ------------------------------------------------------------------ output
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const uint32_t ostage = sbase + Q_OFF;   // Q is dead (S retired, every thread is past its q-row dot product)
    {
      uint32_t v[32];
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        tmem_ld_32x32(t_row + 128 + c * 32, v);
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
    asm volatile("bar.sync 1, 128;" ::: "memory");
    {
      __nv_bfloat16* otile = og + static_cast<size_t>(g * 128) * ldo;
#pragma unroll
      for (int i = 0; i < NCHUNK; ++i) {
        const int c = tid + 128 * i;
        const int row = c / NCHUNK, ch = c - row * NCHUNK;
        const uint4 q = lds16(ostage + static_cast<uint32_t>(c) * 16u);
        *reinterpret_cast<uint4*>(otile + static_cast<size_t>(row) * ldo + ch * 8) = q;
      }
    }
  }

  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc<1>(tmem_base, 256);
}

}  // namespace

int vit_attn2_launch(const AttnParams& p, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0) return -3;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vit_attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  CUtensorMap tm;
  const uint64_t dims[3] = {static_cast<uint64_t>(DH), static_cast<uint64_t>(3 * p.H), static_cast<uint64_t>(p.B) * T_TOK};
  const uint64_t strides[2] = {static_cast<uint64_t>(DH) * 2, static_cast<uint64_t>(3 * p.H * DH) * 2};
  const uint32_t box[3] = {64, 1, 128};
  if (int r = make_tmap_bf16_3d(&tm, p.qkv, dims, strides, box)) return r;
  vit_attn2_kernel<<<p.B * p.H * 2, A2_THREADS, A2_SMEM, stream>>>(tm, p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hb
