// hb_attn_small.cu — head_dim-64 attention for the small sequence models on the hot path:
//   * EVA-CLIP text tower, causal, 77 tokens            (EVA_clip/eva_model.py:132,146,224-230)
//   * MomentModel temporal encoder, T ~ 300 frames       (clip4caption/modules/module_visual.py:154-180)
//   * caption decoder self / cross attention             (clip4caption/modules/module_decoder.py:220-247)
// One CTA per (batch, head); K and V of that head live in shared memory, one warp per query row,
// lanes over keys for q.k and over d for p.v.  These models are launch/latency bound, so
// this kernel is CUDA-core work; the 97 % of FLOPs that matter go through hb_gemm.cu / hb_attn.cu.
#include "hb_attn.cuh"

#include <cmath>

namespace hb {
namespace {

constexpr int DH = 64;
constexpr int K_STRIDE = 144;   // bytes per K row in smem (128 + 16 pad -> conflict-free 16-byte reads)
constexpr int V_STRIDE = 128;   // bytes per V row
constexpr int MAX_TK = 768;
constexpr int MAXJ = MAX_TK / 32;
constexpr int SA_THREADS = 256;

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

__global__ void __launch_bounds__(SA_THREADS) small_attn_kernel(const SmallAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* sk = smem;
  uint8_t* sv = smem + static_cast<size_t>(p.Tk) * K_STRIDE;
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* qg = p.q + b * p.bsq + h * DH;
  const __nv_bfloat16* kg = p.k + b * p.bsk + h * DH;
  const __nv_bfloat16* vg = p.v + b * p.bsv + h * DH;
  __nv_bfloat16* og = p.out + b * p.bso + h * DH;

  // stage K and V (8 x 16-byte chunks per row)
  for (int c = threadIdx.x; c < p.Tk * 8; c += SA_THREADS) {
    const int row = c >> 3, ch = c & 7;
    const uint4 kv = __ldg(reinterpret_cast<const uint4*>(kg + static_cast<size_t>(row) * p.ldk + ch * 8));
    const uint4 vv = __ldg(reinterpret_cast<const uint4*>(vg + static_cast<size_t>(row) * p.ldv + ch * 8));
    *reinterpret_cast<uint4*>(sk + row * K_STRIDE + ch * 16) = kv;
    *reinterpret_cast<uint4*>(sv + row * V_STRIDE + ch * 16) = vv;
  }
  __syncthreads();

  for (int i = warp; i < p.Tq; i += SA_THREADS / 32) {
    float q[DH];
    {
      const uint4* q4 = reinterpret_cast<const uint4*>(qg + static_cast<size_t>(i) * p.ldq);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 v = __ldg(q4 + ch);
        q[ch * 8 + 0] = bf_lo(v.x); q[ch * 8 + 1] = bf_hi(v.x);
        q[ch * 8 + 2] = bf_lo(v.y); q[ch * 8 + 3] = bf_hi(v.y);
        q[ch * 8 + 4] = bf_lo(v.z); q[ch * 8 + 5] = bf_hi(v.z);
        q[ch * 8 + 6] = bf_lo(v.w); q[ch * 8 + 7] = bf_hi(v.w);
      }
    }
    const int tk_eff = (p.mask_mode == 1) ? min(p.Tk, i + 1) : p.Tk;  // hard causal: keys <= query only
    float s[MAXJ];
    float m = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      s[jj] = -INFINITY;
      if (jj * 32 < tk_eff) {
        const int key = jj * 32 + lane;
        if (key < tk_eff) {
          const uint8_t* kr = sk + key * K_STRIDE;
          float acc = 0.f;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 v = *reinterpret_cast<const uint4*>(kr + ch * 16);
            acc += q[ch * 8 + 0] * bf_lo(v.x) + q[ch * 8 + 1] * bf_hi(v.x) + q[ch * 8 + 2] * bf_lo(v.y) +
                   q[ch * 8 + 3] * bf_hi(v.y) + q[ch * 8 + 4] * bf_lo(v.z) + q[ch * 8 + 5] * bf_hi(v.z) +
                   q[ch * 8 + 6] * bf_lo(v.w) + q[ch * 8 + 7] * bf_hi(v.w);
          }
          acc *= p.scale;
          if (p.mask_mode == 2) {
            acc = acc + p.mask_const;                       // fp32 add, as the reference does (quantises the logit)
            if (p.causal_soft && key > i) acc = acc + (-10000.0f);
          }
          s[jj] = acc;
          m = fmaxf(m, acc);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      if (jj * 32 < tk_eff) {
        s[jj] = __expf(s[jj] - m);  // exp(-inf) = 0 for keys past tk_eff
        sum += s[jj];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) {
      if (jj * 32 < tk_eff) {
        const int nk = min(32, tk_eff - jj * 32);
        for (int j = 0; j < nk; ++j) {
          const float pj = __shfl_sync(0xffffffffu, s[jj], j);
          const uint32_t vv = *reinterpret_cast<const uint32_t*>(sv + (jj * 32 + j) * V_STRIDE + lane * 4);
          o0 += pj * bf_lo(vv);
          o1 += pj * bf_hi(vv);
        }
      }
    }
    __nv_bfloat162 r = __floats2bfloat162_rn(o0 * inv, o1 * inv);
    *reinterpret_cast<__nv_bfloat162*>(og + static_cast<size_t>(i) * p.ldo + lane * 2) = r;
  }
}

// ---- fp32 variant (MomentModel / caption decoder "precise" path) ----------------------------------------------------------
// q/k/v/out fp32.  One CTA per (batch, head, block of 64 queries); keys are streamed through shared memory in tiles of 256
// with an online softmax (running max / sum / output per query in registers), so any sequence length works.
constexpr int KF_STRIDE = 272;  // 256 B + 16 B pad: conflict-free 16-byte reads with lanes on consecutive keys
constexpr int VF_STRIDE = 256;
constexpr int KT_F32 = 256;     // keys per tile
constexpr int QB_F32 = 64;      // queries per CTA (8 per warp)

__global__ void __launch_bounds__(SA_THREADS) small_attn_f32_kernel(const SmallAttnF32Params p, int q_blocks) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* sk = smem;
  uint8_t* sv = smem + static_cast<size_t>(KT_F32) * KF_STRIDE;
  const int qb = blockIdx.x % q_blocks;
  const int bh = blockIdx.x / q_blocks;
  const int b = bh / p.H, h = bh - b * p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* qg = p.q + b * p.bsq + h * DH;
  const float* kg = p.k + (b / p.kv_div) * p.bsk + h * DH;
  const float* vg = p.v + (b / p.kv_div) * p.bsv + h * DH;
  float* og = p.out + b * p.bso + h * DH;
  const int q0 = qb * QB_F32 + warp * 8;  // this warp's queries: q0 .. q0+7
  float m[8], sum[8], o0[8], o1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -INFINITY; sum[i] = 0.f; o0[i] = 0.f; o1[i] = 0.f; }
  // keys needed by this CTA (hard-causal: nothing beyond its last query)
  const int q_last = min(p.Tq, qb * QB_F32 + QB_F32) - 1;
  const int tk_cta = (p.mask_mode == 1) ? min(p.Tk, q_last + 1) : p.Tk;
  for (int t0 = 0; t0 < tk_cta; t0 += KT_F32) {
    const int tn = min(KT_F32, tk_cta - t0);
    __syncthreads();  // previous tile fully consumed
    for (int c = threadIdx.x; c < tn * 16; c += SA_THREADS) {
      const int row = c >> 4, ch = c & 15;
      *reinterpret_cast<float4*>(sk + row * KF_STRIDE + ch * 16) =
          __ldg(reinterpret_cast<const float4*>(kg + static_cast<size_t>(t0 + row) * p.ldk + ch * 4));
      *reinterpret_cast<float4*>(sv + row * VF_STRIDE + ch * 16) =
          __ldg(reinterpret_cast<const float4*>(vg + static_cast<size_t>(t0 + row) * p.ldv + ch * 4));
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < 8; ++qi) {
      const int i = q0 + qi;
      if (i >= p.Tq) continue;  // warp-uniform
      const int tk_eff = (p.mask_mode == 1) ? min(tn, i + 1 - t0) : tn;  // keys of this tile visible to query i
      if (tk_eff <= 0) continue;
      float q[DH];
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(qg + static_cast<size_t>(i) * p.ldq) + ch);
        q[ch * 4] = v.x; q[ch * 4 + 1] = v.y; q[ch * 4 + 2] = v.z; q[ch * 4 + 3] = v.w;
      }
      float s[KT_F32 / 32];
      float tm = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < KT_F32 / 32; ++jj) {
        s[jj] = -INFINITY;
        const int key = jj * 32 + lane;
        if (key < tk_eff) {
          const uint8_t* kr = sk + key * KF_STRIDE;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int ch = 0; ch < 16; ++ch) {
            const float4 v = *reinterpret_cast<const float4*>(kr + ch * 16);
            a0 = fmaf(q[ch * 4], v.x, a0); a1 = fmaf(q[ch * 4 + 1], v.y, a1);
            a2 = fmaf(q[ch * 4 + 2], v.z, a2); a3 = fmaf(q[ch * 4 + 3], v.w, a3);
          }
          float acc = ((a0 + a1) + (a2 + a3)) * p.scale;
          if (p.mask_mode == 2) {
            acc = acc + p.mask_const;   // fp32 add: quantises the logit exactly as the reference's mask add does
            if (p.causal_soft && (t0 + key) > i) acc = acc + (-10000.0f);
          }
          s[jj] = acc;
          tm = fmaxf(tm, acc);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, o));
      const float m_new = fmaxf(m[qi], tm);
      const float corr = expf(m[qi] - m_new);  // 0 on the first tile (m = -inf)
      float ts = 0.f;
#pragma unroll
      for (int jj = 0; jj < KT_F32 / 32; ++jj) {
        s[jj] = expf(s[jj] - m_new);  // exp(-inf) = 0 for masked / absent keys
        ts += s[jj];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ts += __shfl_xor_sync(0xffffffffu, ts, o);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int jj = 0; jj < KT_F32 / 32; ++jj) {
        if (jj * 32 < tk_eff) {
          const int nk = min(32, tk_eff - jj * 32);
          for (int j = 0; j < nk; ++j) {
            const float pj = __shfl_sync(0xffffffffu, s[jj], j);
            const float2 vv = *reinterpret_cast<const float2*>(sv + (jj * 32 + j) * VF_STRIDE + lane * 8);
            a0 = fmaf(pj, vv.x, a0);
            a1 = fmaf(pj, vv.y, a1);
          }
        }
      }
      m[qi] = m_new;
      sum[qi] = sum[qi] * corr + ts;
      o0[qi] = o0[qi] * corr + a0;
      o1[qi] = o1[qi] * corr + a1;
    }
  }
#pragma unroll
  for (int qi = 0; qi < 8; ++qi) {
    const int i = q0 + qi;
    if (i < p.Tq) {
      const float inv = 1.0f / sum[qi];
      *reinterpret_cast<float2*>(og + static_cast<size_t>(i) * p.ldo + lane * 2) = make_float2(o0[qi] * inv, o1[qi] * inv);
    }
  }
}

// Decode step (Tq == 1): one WARP per (batch row, head) — the caption decoder's self-attention over the KV cache (<= 48 keys)
// and its cross-attention over the 20 clip frames (module_decoder.py:220-247).  Lanes take keys for q.k (fp32, q broadcast from
// registers), the softmax is a warp reduction, lanes take dimension pairs for p.v.  The tiled kernel above staged a 256-key tile
// with two __syncthreads per CTA for ONE query: 74 us per call, 24 % of step captioning.
__global__ void __launch_bounds__(256) decode_attn_f32_kernel(const SmallAttnF32Params p) {
  // Tk <= 64: every lane owns keys `lane` and `lane + 32`.  Operation order (4-way split dot product, max, exp, lane-major sum,
  // key-ordered p.v) is that of small_attn_f32_kernel for a single tile, so both kernels give bit-identical results.
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= p.B * p.H) return;
  const int b = w / p.H, h = w - b * p.H;
  const float* qg = p.q + b * p.bsq + h * DH;
  // K / V batch of this lane's two keys (and, by shuffle, of every key in the p.v loop): b / kv_div, or the beam index table
  long long kb[2];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const int key = jj * 32 + lane;
    kb[jj] = (p.kv_row_idx != nullptr) ? static_cast<long long>(p.kv_row_idx[static_cast<long long>(b) * p.ld_idx + (key < p.Tk ? key : 0)])
                                       : static_cast<long long>(b / p.kv_div);
  }
  const float2 q2 = *reinterpret_cast<const float2*>(qg + lane * 2);   // lane holds q[2 lane], q[2 lane + 1]
  float s[2];
  float m = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const int key = jj * 32 + lane;
    s[jj] = -INFINITY;
    const float* kr = p.k + kb[jj] * p.bsk + h * DH + static_cast<size_t>(key < p.Tk ? key : 0) * p.ldk;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int ch = 0; ch < 16; ++ch) {
      const float4 kv = __ldg(reinterpret_cast<const float4*>(kr + ch * 4));
      const float qa = __shfl_sync(0xffffffffu, q2.x, ch * 2), qb = __shfl_sync(0xffffffffu, q2.y, ch * 2);
      const float qc = __shfl_sync(0xffffffffu, q2.x, ch * 2 + 1), qd = __shfl_sync(0xffffffffu, q2.y, ch * 2 + 1);
      a0 = fmaf(qa, kv.x, a0); a1 = fmaf(qb, kv.y, a1); a2 = fmaf(qc, kv.z, a2); a3 = fmaf(qd, kv.w, a3);
    }
    if (key < p.Tk) {
      float acc = ((a0 + a1) + (a2 + a3)) * p.scale;
      if (p.mask_mode == 2) acc = acc + p.mask_const;   // fp32 add, as the reference's mask add (the single query sees every key)
      s[jj] = acc;
      m = fmaxf(m, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float ts = 0.f;
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    s[jj] = expf(s[jj] - m);   // exp(-inf) = 0 for absent keys
    ts += s[jj];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ts += __shfl_xor_sync(0xffffffffu, ts, o);
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {
    const int nk = min(32, p.Tk - jj * 32);
    // eight V rows in flight per step (one load at a time left the warp waiting on L2 for every key: 14 us per call); the
    // accumulation order over keys is unchanged
    for (int j0 = 0; j0 < nk; j0 += 8) {
      float2 vv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        vv[u] = make_float2(0.f, 0.f);
        if (j0 + u < nk) {   // warp-uniform
          const long long vb = __shfl_sync(0xffffffffu, kb[jj], j0 + u);
          vv[u] = __ldg(reinterpret_cast<const float2*>(p.v + vb * p.bsv + h * DH + static_cast<size_t>(jj * 32 + j0 + u) * p.ldv + lane * 2));
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (j0 + u < nk) {   // warp-uniform
          const float pw = __shfl_sync(0xffffffffu, s[jj], j0 + u);
          o0 = fmaf(pw, vv[u].x, o0);
          o1 = fmaf(pw, vv[u].y, o1);
        }
      }
    }
  }
  const float inv = 1.0f / ts;
  const float y0 = o0 * inv, y1 = o1 * inv;
  *reinterpret_cast<float2*>(p.out + b * p.bso + h * DH + lane * 2) = make_float2(y0, y1);
  if (p.op_out != nullptr) {
    const __nv_bfloat16 h0 = __float2bfloat16(y0), h1 = __float2bfloat16(y1);
    const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(y0 - __bfloat162float(h0), y1 - __bfloat162float(h1));
    __nv_bfloat16* o = p.op_out + static_cast<long long>(b) * 3 * p.ld_op + h * DH + lane * 2;
    *reinterpret_cast<__nv_bfloat162*>(o) = ll;
    *reinterpret_cast<__nv_bfloat162*>(o + p.ld_op) = hh;
    *reinterpret_cast<__nv_bfloat162*>(o + 2 * p.ld_op) = hh;
  }
}

}  // namespace

int small_attn_f32_launch(const SmallAttnF32Params& p, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0 || p.Tq <= 0 || p.Tk <= 0) return -3;
  if ((p.kv_row_idx != nullptr || p.op_out != nullptr) && !(p.Tq == 1 && p.Tk <= 64 && p.mask_mode != 1 && !p.causal_soft && (p.kv_div == 1 || p.kv_row_idx == nullptr))) return -3;
  if (p.Tq == 1 && p.Tk <= 64 && p.mask_mode != 1 && !p.causal_soft) {
    const long long warps = static_cast<long long>(p.B) * p.H;
    decode_attn_f32_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, stream>>>(p);
    return static_cast<int>(cudaGetLastError());
  }
  const size_t smem = static_cast<size_t>(KT_F32) * (KF_STRIDE + VF_STRIDE);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(small_attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const int q_blocks = (p.Tq + QB_F32 - 1) / QB_F32;
  small_attn_f32_kernel<<<p.B * p.H * q_blocks, SA_THREADS, smem, stream>>>(p, q_blocks);
  return static_cast<int>(cudaGetLastError());
}

int small_attn_launch(const SmallAttnParams& p, cudaStream_t stream) {
  if (p.B <= 0 || p.H <= 0 || p.Tq <= 0 || p.Tk <= 0) return -3;
  if (p.Tk > MAX_TK) return -6;
  const size_t smem = static_cast<size_t>(p.Tk) * (K_STRIDE + V_STRIDE);
  static size_t attr_set = 0;
  if (smem > 48 * 1024 && smem > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(small_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MAX_TK * (K_STRIDE + V_STRIDE));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = MAX_TK * (K_STRIDE + V_STRIDE);
  }
  small_attn_kernel<<<p.B * p.H, SA_THREADS, smem, stream>>>(p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hb
