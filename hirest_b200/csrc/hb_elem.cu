// hb_elem.cu — HBM-bound row kernels around the GEMMs: LayerNorm, patch gather (im2col), cls/pos rows,
// token embedding, frame mean-pool + L2 normalise, split-bf16 packing for the exact similarity GEMM.
// All use 128-bit loads/stores and one warp per row.
#include "hb_elem.cuh"

#include <cuda_bf16.h>
#include <cstdint>

namespace hb {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: y = (x - mean) / sqrt(var_biased + eps) * w + b      (nn.LayerNorm, vit_model.py:159-165,
// eva_model.py:19-25,304; identical formula to the "TF-style" LayerNorm of until_module.py:40-53)
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 12;  // float4 per lane cached in registers -> D <= 1536

template <bool OUT_BF16>
__global__ void __launch_bounds__(256) layernorm_kernel(const LayerNormParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.rows) return;
  long long in_row = warp;
  if (p.row_idx != nullptr) in_row = p.row_idx[warp];
  const float4* x4 = reinterpret_cast<const float4*>(p.x + in_row * p.ldx);
  const int nv = p.D >> 2;
  float4 v[LN_MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      v[i] = x4[idx];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(p.D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(p.D) + p.eps);
  const float4* w4 = reinterpret_cast<const float4*>(p.w);
  const float4* b4 = reinterpret_cast<const float4*>(p.b);
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float4 w = __ldg(w4 + idx), b = __ldg(b4 + idx);
      const float y0 = (v[i].x - mean) * rstd * w.x + b.x;
      const float y1 = (v[i].y - mean) * rstd * w.y + b.y;
      const float y2 = (v[i].z - mean) * rstd * w.z + b.z;
      const float y3 = (v[i].w - mean) * rstd * w.w + b.w;
      if constexpr (OUT_BF16) {
        uint2 o;
        o.x = pack2(y0, y1);
        o.y = pack2(y2, y3);
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.y) + static_cast<long long>(warp) * p.ldy)[idx] = o;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + static_cast<long long>(warp) * p.ldy)[idx] =
            make_float4(y0, y1, y2, y3);
      }
    }
  }
}

// Row statistics + bf16 copy for the LayerNorm-folded GEMMs (entry of the ViT block stack): stats[r, 0] = (sum, sum of squares),
// stats[r, 1..slots) = 0 (the GEMM epilogues fill one slot per 128 columns), xb[r,:] = bf16(x[r,:]).  One warp per row.
__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                                                        float* __restrict__ stats, long long rows, int D, int slots) {
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* x4 = reinterpret_cast<const float4*>(x + warp * D);
  uint2* o2 = reinterpret_cast<uint2*>(xb + warp * D);
  float s = 0.f, ss = 0.f;
  for (int i = lane; i < (D >> 2); i += 32) {
    const float4 v = x4[i];
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    uint2 o;
    o.x = pack2(v.x, v.y);
    o.y = pack2(v.z, v.w);
    o2[i] = o;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  float2* st = reinterpret_cast<float2*>(stats) + warp * slots;
  if (lane < slots) st[lane] = lane == 0 ? make_float2(s, ss) : make_float2(0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------
// Patch gather for Conv2d(3, D, k=14, s=14) as a GEMM (vit_model.py:198,205): row = b*256 + ph*16 + pw,
// column = c*196 + kh*14 + kw (the conv weight's own (c,kh,kw) flattening), padded to ldo columns with zeros.
// One thread per (row, c, kh): 14 contiguous input floats -> 14 bf16.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_patch_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                           int B, int S, int P, int ldo) {
  const int G = S / P;  // patches per side
  const long long total = static_cast<long long>(B) * G * G * 3 * P;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int pw = static_cast<int>(t % G);
  long long r = t / G;
  const int kh = static_cast<int>(r % P); r /= P;
  const int c = static_cast<int>(r % 3); r /= 3;
  const int ph = static_cast<int>(r % G);
  const int b = static_cast<int>(r / G);
  const float* src = img + ((static_cast<long long>(b) * 3 + c) * S + (ph * P + kh)) * S + pw * P;
  __nv_bfloat16* dst = out + (static_cast<long long>(b) * G * G + ph * G + pw) * ldo + (c * P + kh) * P;
  for (int i = 0; i < P; i += 2) {
    const float2 v = *reinterpret_cast<const float2*>(src + i);
    *reinterpret_cast<uint32_t*>(dst + i) = pack2(v.x, v.y);
  }
  if (c == 2 && kh == P - 1) {
    for (int i = 3 * P * P; i < ldo; i += 2) *reinterpret_cast<uint32_t*>(dst + (i - (c * P + kh) * P)) = 0u;
  }
}

// Same gather from raw uint8 frames [B,3,S,S] with the CPU preprocessing folded in: ToTensor (x / 255) and
// Normalize ((x - mean[c]) / std[c]) of EVA_clip/eva_clip.py:144-153, computed in fp32 exactly as torchvision does.
__global__ void __launch_bounds__(256) im2col_patch_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                                              int B, int S, int P, int ldo, float m0, float m1, float m2,
                                                              float s0, float s1, float s2) {
  const int G = S / P;
  const long long total = static_cast<long long>(B) * G * G * 3 * P;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int pw = static_cast<int>(t % G);
  long long r = t / G;
  const int kh = static_cast<int>(r % P); r /= P;
  const int c = static_cast<int>(r % 3); r /= 3;
  const int ph = static_cast<int>(r % G);
  const int b = static_cast<int>(r / G);
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
  const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
  const uint8_t* src = img + ((static_cast<long long>(b) * 3 + c) * S + (ph * P + kh)) * S + pw * P;
  __nv_bfloat16* dst = out + (static_cast<long long>(b) * G * G + ph * G + pw) * ldo + (c * P + kh) * P;
  for (int i = 0; i < P; i += 2) {
    const uint16_t two = *reinterpret_cast<const uint16_t*>(src + i);
    const float a = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(two & 0xFF), 255.0f), mean), sd);
    const float bq = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(two >> 8), 255.0f), mean), sd);
    *reinterpret_cast<uint32_t*>(dst + i) = pack2(a, bq);
  }
  if (c == 2 && kh == P - 1) {
    for (int i = 3 * P * P; i < ldo; i += 2) *reinterpret_cast<uint32_t*>(dst + (i - (c * P + kh) * P)) = 0u;
  }
}

// x[b*T + 0, :] = cls + pos[0]   (vit_model.py:330-333)
__global__ void cls_row_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int B,
                               int T, int D) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(B) * D) return;
  const int d = static_cast<int>(t % D);
  const long long b = t / D;
  x[b * T * D + d] = cls[d] + pos[d];
}

// x[q*C + t, :] = tok_emb[ids[q,t]] + pos[t]   (eva_model.py:233-235);  eot_row[q] = q*C + argmax_t ids[q,t] (:243)
__global__ void __launch_bounds__(256) text_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ tok,
                                                         const float* __restrict__ pos, float* __restrict__ x,
                                                         int* __restrict__ eot_row, int Q, int C, int W, int V) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= Q * C) return;
  const int q = warp / C, t = warp - q * C;
  long long id = ids[warp];
  if (id < 0) id = 0;
  if (id >= V) id = V - 1;
  const float4* e4 = reinterpret_cast<const float4*>(tok + id * W);
  const float4* p4 = reinterpret_cast<const float4*>(pos + static_cast<long long>(t) * W);
  float4* o4 = reinterpret_cast<float4*>(x + static_cast<long long>(warp) * W);
  for (int i = lane; i < (W >> 2); i += 32) {
    const float4 a = __ldg(e4 + i), b = __ldg(p4 + i);
    o4[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
  if (t == 0 && lane == 0) {
    long long best = ids[static_cast<long long>(q) * C];
    int arg = 0;
    for (int j = 1; j < C; ++j) {
      const long long v = ids[static_cast<long long>(q) * C + j];
      if (v > best) { best = v; arg = j; }  // first maximum, like torch.argmax
    }
    eot_row[q] = q * C + arg;
  }
}

// out[v, :] = normalize(mean_f emb[v, f, :])   (inference_video_retrieval.py:283-285 / 323-326; F = 1: :210-212)
// One CTA (256 threads) per video; E <= 4096.
template <bool OUT_BF16>
__global__ void __launch_bounds__(256) pool_normalize_kernel(const float* __restrict__ emb, void* __restrict__ out, int F,
                                                             int E, int do_normalize) {
  __shared__ float red[8];
  const long long v = blockIdx.x;
  const float* base = emb + v * F * E;
  float acc[16];
  float ss = 0.f;
  const float invF = 1.0f / static_cast<float>(F);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int e = threadIdx.x + i * 256;
    acc[i] = 0.f;
    if (e < E) {
      float s = 0.f;
      for (int f = 0; f < F; ++f) s += base[static_cast<long long>(f) * E + e];
      acc[i] = s * invF;
      ss += acc[i] * acc[i];
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv = do_normalize ? 1.0f / sqrtf(tot) : 1.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int e = threadIdx.x + i * 256;
    if (e < E) {
      if constexpr (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(out)[v * E + e] = __float2bfloat16(acc[i] * inv);
      else reinterpret_cast<float*>(out)[v * E + e] = acc[i] * inv;
    }
  }
}

// Ragged variant for the cached-feature path (inference_video_retrieval.py:298-327): video v owns rows
// [offsets[v], offsets[v+1]) of one packed [sum T, E] feature blob.  If n_sub > 0 the rows are first subsampled exactly like
// `np.linspace(0, T - 1, n_sub).astype(int)` (:313: float64 `i * step`, last sample forced to T - 1, truncation), then
// mean-pooled in fp32 and L2-normalised (:321-326).  One CTA per video; the gather indices are computed on the fly.
template <typename TIn>
__global__ void __launch_bounds__(256) subsample_pool_normalize_kernel(const TIn* __restrict__ feats, const long long* __restrict__ offsets,
                                                                       float* __restrict__ out, int n_sub, int E) {
  __shared__ float red[8];
  const long long v = blockIdx.x;
  const long long r0 = offsets[v];
  const int T = static_cast<int>(offsets[v + 1] - r0);
  const int F = (n_sub > 0) ? n_sub : T;
  const double step = (n_sub > 1) ? static_cast<double>(T - 1) / static_cast<double>(n_sub - 1) : 0.0;
  float acc[16];
  float ss = 0.f;
  const float invF = 1.0f / static_cast<float>(F);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int f = 0; f < F; ++f) {
    int row = f;
    if (n_sub > 0) row = (f == n_sub - 1 && n_sub > 1) ? (T - 1) : static_cast<int>(static_cast<double>(f) * step);
    const TIn* base = feats + (r0 + row) * E;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int e = threadIdx.x + i * 256;
      if (e < E) acc[i] += static_cast<float>(base[e]);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    acc[i] = (T > 0) ? acc[i] * invF : 0.f;
    ss += acc[i] * acc[i];
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv = 1.0f / sqrtf(tot);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int e = threadIdx.x + i * 256;
    if (e < E) out[v * E + e] = acc[i] * inv;
  }
}

// Dataset-side frame resampling of cached features (hirest_dataset.py:333-356 for the video features, :383-403 for the warped
// ASR features): per video, T rows -> n_out rows.  T > n_out: rows np.linspace(0, T-1, n_out).astype(int) (float64 i*step, last
// sample forced to T-1, truncation); T <= n_out: repeat-pad, source row k fills output slots [(k*n_out)//T, ((k+1)*n_out)//T),
// i.e. output slot j reads row ceil((j+1)*T / n_out) - 1.  One warp per output row, 16-byte copies.
__global__ void __launch_bounds__(256) resample_rows_kernel(const float* __restrict__ feats, const long long* __restrict__ offsets,
                                                            float* __restrict__ out, long long V, int n_out, int C) {
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= V * n_out) return;
  const long long v = w / n_out;
  const int j = static_cast<int>(w - v * n_out);
  const long long r0 = offsets[v];
  const int T = static_cast<int>(offsets[v + 1] - r0);
  float4* dst = reinterpret_cast<float4*>(out + w * C);
  if (T <= 0) {
    for (int i = lane; i < C / 4; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  int row;
  if (T > n_out) {
    const double step = (n_out > 1) ? static_cast<double>(T - 1) / static_cast<double>(n_out - 1) : 0.0;
    row = (j == n_out - 1 && n_out > 1) ? (T - 1) : static_cast<int>(static_cast<double>(j) * step);
  } else {
    row = static_cast<int>((static_cast<long long>(j + 1) * T + n_out - 1) / n_out) - 1;
  }
  const float4* src = reinterpret_cast<const float4*>(feats + (r0 + row) * C);
  for (int i = lane; i < C / 4; i += 32) dst[i] = __ldg(src + i);
}

// ASR feature warping (hirest_dataset.py:370-381): per video a [len, C] zero tensor in which sentence i's feature row fills the
// seconds [start_i, end_i) (Python slice semantics: clamped to [0, len]; later sentences overwrite earlier ones).  Output rows
// are packed like the video features (frame_offsets); one warp per output row scans its video's sentences for the LAST hit.
__global__ void __launch_bounds__(256) asr_warp_kernel(const float* __restrict__ asr, const long long* __restrict__ sub_offsets,
                                                       const int* __restrict__ starts, const int* __restrict__ ends,
                                                       const long long* __restrict__ frame_offsets, const int* __restrict__ row_video,
                                                       float* __restrict__ out, long long rows, int C) {
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows) return;
  const int v = row_video[w];
  const int t = static_cast<int>(w - frame_offsets[v]);
  const int len = static_cast<int>(frame_offsets[v + 1] - frame_offsets[v]);
  const long long s0 = sub_offsets[v], s1 = sub_offsets[v + 1];
  int hit = -1;
  for (long long i = s0 + lane; i < s1; i += 32) {
    int a = starts[i], b = ends[i];
    if (a < 0) a = max(0, a + len);   // Python slice: negative indices count from the end
    if (b < 0) b = max(0, b + len);
    if (a <= t && t < b) hit = static_cast<int>(i - s0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hit = max(hit, __shfl_xor_sync(0xffffffffu, hit, o));
  float4* dst = reinterpret_cast<float4*>(out + w * C);
  if (hit < 0) {
    for (int i = lane; i < C / 4; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const float4* src = reinterpret_cast<const float4*>(asr + (s0 + hit) * C);
    for (int i = lane; i < C / 4; i += 32) dst[i] = __ldg(src + i);
  }
}

// Split fp32 rows into three bf16 parts (hi + mid + lo = the fp32 value exactly: 3 x 8 significant bits) and lay
// them out so that ONE bf16 GEMM over K = 6E reproduces the fp32 dot product to ~2^-23 relative:
//   mode 0 (text side):  [hi | lo | mid | mid | hi  | hi]
//   mode 1 (video side): [lo | hi | mid | hi  | mid | hi]
//   A'.B'^T = hi.lo + lo.hi + mid.mid + mid.hi + hi.mid + hi.hi      (dropped terms are <= 2^-24 relative)
// Smallest terms first: the tensor core's fp32 accumulation truncates, so the error of each add scales with the
// running sum; accumulating the 2^-16 / 2^-8 corrections before the dominant hi.hi segment keeps it ~1e-7.
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                         long long rows, int E, int mode) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= rows * E) return;
  const long long r = t / E;
  const int e = static_cast<int>(t - r * E);
  const float v = x[t];
  const __nv_bfloat16 hi = __float2bfloat16(v);
  const float r1 = v - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16(r1);
  const __nv_bfloat16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
  __nv_bfloat16* o = out + r * 6 * E;
  if (mode == 0) {
    o[e] = hi; o[E + e] = lo; o[2 * E + e] = mid; o[3 * E + e] = mid; o[4 * E + e] = hi; o[5 * E + e] = hi;
  } else {
    o[e] = lo; o[E + e] = hi; o[2 * E + e] = mid; o[3 * E + e] = hi; o[4 * E + e] = mid; o[5 * E + e] = hi;
  }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          long long n4) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[t];
  uint2 o;
  o.x = pack2(v.x, v.y);
  o.y = pack2(v.z, v.w);
  reinterpret_cast<uint2*>(y)[t] = o;
}

inline unsigned blocks_for(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }

}  // namespace

int layernorm_launch(const LayerNormParams& p, bool out_bf16, cudaStream_t s) {
  if (p.rows <= 0) return 0;
  if (p.D % 4 != 0 || p.D > LN_MAXV * 128) return -7;
  const unsigned blocks = blocks_for(p.rows, 8);
  if (out_bf16) layernorm_kernel<true><<<blocks, 256, 0, s>>>(p);
  else layernorm_kernel<false><<<blocks, 256, 0, s>>>(p);
  return static_cast<int>(cudaGetLastError());
}

int row_stats_launch(const float* x, __nv_bfloat16* xb, float* stats, long long rows, int D, int slots, cudaStream_t s) {
  if (rows <= 0) return 0;
  if (D % 4 != 0 || slots < 1 || slots > 32) return -7;
  row_stats_kernel<<<blocks_for(rows, 8), 256, 0, s>>>(x, xb, stats, rows, D, slots);
  return static_cast<int>(cudaGetLastError());
}

int im2col_patch_launch(const float* img, __nv_bfloat16* out, int B, int S, int P, int ldo, cudaStream_t s) {
  if (S % P != 0 || P % 2 != 0 || ldo < 3 * P * P || ldo % 2 != 0) return -7;
  const long long total = static_cast<long long>(B) * (S / P) * (S / P) * 3 * P;
  im2col_patch_kernel<<<blocks_for(total, 256), 256, 0, s>>>(img, out, B, S, P, ldo);
  return static_cast<int>(cudaGetLastError());
}

int im2col_patch_u8_launch(const uint8_t* img, __nv_bfloat16* out, int B, int S, int P, int ldo, const float mean[3],
                           const float stdv[3], cudaStream_t s) {
  if (S % P != 0 || P % 2 != 0 || S % 2 != 0 || ldo < 3 * P * P || ldo % 2 != 0) return -7;
  const long long total = static_cast<long long>(B) * (S / P) * (S / P) * 3 * P;
  im2col_patch_u8_kernel<<<blocks_for(total, 256), 256, 0, s>>>(img, out, B, S, P, ldo, mean[0], mean[1], mean[2], stdv[0], stdv[1],
                                                                stdv[2]);
  return static_cast<int>(cudaGetLastError());
}

int cls_row_launch(float* x, const float* cls, const float* pos, int B, int T, int D, cudaStream_t s) {
  cls_row_kernel<<<blocks_for(static_cast<long long>(B) * D, 256), 256, 0, s>>>(x, cls, pos, B, T, D);
  return static_cast<int>(cudaGetLastError());
}

int text_embed_launch(const long long* ids, const float* tok, const float* pos, float* x, int* eot_row, int Q, int C, int W,
                      int V, cudaStream_t s) {
  if (W % 4 != 0) return -7;
  text_embed_kernel<<<blocks_for(static_cast<long long>(Q) * C, 8), 256, 0, s>>>(ids, tok, pos, x, eot_row, Q, C, W, V);
  return static_cast<int>(cudaGetLastError());
}

int pool_normalize_launch(const float* emb, void* out, long long V, int F, int E, bool normalize, bool out_bf16,
                          cudaStream_t s) {
  if (V <= 0) return 0;
  if (E > 4096 || F <= 0) return -7;
  if (out_bf16) pool_normalize_kernel<true><<<static_cast<unsigned>(V), 256, 0, s>>>(emb, out, F, E, normalize ? 1 : 0);
  else pool_normalize_kernel<false><<<static_cast<unsigned>(V), 256, 0, s>>>(emb, out, F, E, normalize ? 1 : 0);
  return static_cast<int>(cudaGetLastError());
}

int subsample_pool_normalize_launch(const float* feats, const long long* offsets, float* out, long long V, int n_sub, int E,
                                    cudaStream_t s) {
  if (V <= 0) return 0;
  if (E > 4096) return -7;
  subsample_pool_normalize_kernel<float><<<static_cast<unsigned>(V), 256, 0, s>>>(feats, offsets, out, n_sub, E);
  return static_cast<int>(cudaGetLastError());
}

int subsample_pool_normalize_bf16_launch(const __nv_bfloat16* feats, const long long* offsets, float* out, long long V, int n_sub, int E,
                                         cudaStream_t s) {
  if (V <= 0) return 0;
  if (E > 4096) return -7;
  subsample_pool_normalize_kernel<__nv_bfloat16><<<static_cast<unsigned>(V), 256, 0, s>>>(feats, offsets, out, n_sub, E);
  return static_cast<int>(cudaGetLastError());
}

int resample_rows_launch(const float* feats, const long long* offsets, float* out, long long V, int n_out, int C, cudaStream_t s) {
  if (V <= 0 || n_out <= 0) return 0;
  if (C % 4 != 0) return -7;
  resample_rows_kernel<<<blocks_for(V * n_out * 32, 256), 256, 0, s>>>(feats, offsets, out, V, n_out, C);
  return static_cast<int>(cudaGetLastError());
}

int asr_warp_launch(const float* asr, const long long* sub_offsets, const int* starts, const int* ends, const long long* frame_offsets,
                    const int* row_video, float* out, long long rows, int C, cudaStream_t s) {
  if (rows <= 0) return 0;
  if (C % 4 != 0) return -7;
  asr_warp_kernel<<<blocks_for(rows * 32, 256), 256, 0, s>>>(asr, sub_offsets, starts, ends, frame_offsets, row_video, out, rows, C);
  return static_cast<int>(cudaGetLastError());
}

int split_bf16_launch(const float* x, __nv_bfloat16* out, long long rows, int E, int mode, cudaStream_t s) {
  if (rows <= 0) return 0;
  split_bf16_kernel<<<blocks_for(rows * E, 256), 256, 0, s>>>(x, out, rows, E, mode);
  return static_cast<int>(cudaGetLastError());
}

int f32_to_bf16_launch(const float* x, __nv_bfloat16* y, long long n, cudaStream_t s) {
  if (n % 4 != 0) return -7;
  if (n == 0) return 0;
  f32_to_bf16_kernel<<<blocks_for(n / 4, 256), 256, 0, s>>>(x, y, n / 4);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hb
