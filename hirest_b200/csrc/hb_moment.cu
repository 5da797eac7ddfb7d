// hb_moment.cu — row kernels of the MomentModel path (see hb_moment.cuh for the reference lines each one restates).
// These kernels are launch/latency-bound at the reference's sizes (B*T ~ 19k rows): one warp per row, 128-bit accesses.
#include "hb_moment.cuh"

#include <cfloat>

namespace hb {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float gelu_exact(float x) { return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f)); }

// 8 elements per thread: two float4 loads, three 16-byte stores (K % 8 == 0; the scalar form spent 27 % of the segmentation time
// in 2-byte stores)
__global__ void __launch_bounds__(256) split3_act_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                         long long rows, int K, int gelu) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int k8 = K >> 3;
  if (t >= rows * k8) return;
  const long long r = t / k8;
  const int k = static_cast<int>(t - r * k8) * 8;
  const float4 a = *reinterpret_cast<const float4*>(x + r * K + k), b = *reinterpret_cast<const float4*>(x + r * K + k + 4);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v0 = v[2 * i], v1 = v[2 * i + 1];
    if (gelu) { v0 = gelu_exact(v0); v1 = gelu_exact(v1); }
    const __nv_bfloat16 h0 = __float2bfloat16(v0), h1 = __float2bfloat16(v1);
    const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
    hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
    lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  __nv_bfloat16* o = out + r * 3 * K + k;
  const uint4 H = make_uint4(hi[0], hi[1], hi[2], hi[3]), L = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  *reinterpret_cast<uint4*>(o) = L;
  *reinterpret_cast<uint4*>(o + K) = H;
  *reinterpret_cast<uint4*>(o + 2 * K) = H;
}

// Finishes a split-K GEMM (GemmParams::k_splits): y = [LayerNorm]([gelu](sum_s part[s] + bias) + resid), slices added in index
// order (deterministic, and a row's result does not depend on how many rows share the launch).  Writes y as fp32 and / or as the
// [lo | hi | hi] split operand of the NEXT GEMM, so the decoder's dense -> (+residual) -> LayerNorm -> next dense chain
// (module_decoder.py:160-292) is GEMM, finish, GEMM with no separate bias / LayerNorm / split kernels;
// one CTA per row, one float4 of the row per thread (N <= 1536); two-pass LayerNorm with rsqrtf like layernorm_kernel.
constexpr int FIN_MAX_THREADS = 384;   // one float4 of the row per thread -> N <= 1536
__device__ __forceinline__ float fin_block_sum(float v, float* red, int warp, int lane, int nwarps) {
  v = warp_sum(v);
  __syncthreads();                  // `red` may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nwarps; ++w) t += red[w];   // same order in every thread
  return t;
}
__global__ void __launch_bounds__(FIN_MAX_THREADS) splitk_finish_kernel(const float* __restrict__ part, long long split_stride, int splits,
                                                                        const float* __restrict__ bias, const float* __restrict__ resid,
                                                                        const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                                        float eps, int gelu, float* __restrict__ y,
                                                                        __nv_bfloat16* __restrict__ op, int N) {
  __shared__ float red[FIN_MAX_THREADS / 32];
  const long long row = blockIdx.x;
  const int idx = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
  const bool live = idx < (N >> 2);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const float* p0 = part + row * N + idx * 4;
    a = *reinterpret_cast<const float4*>(p0);
    int sl = 1;
    for (; sl + 4 <= splits; sl += 4) {   // four independent loads in flight, added in slice order
      const float4 t0 = *reinterpret_cast<const float4*>(p0 + (sl + 0) * split_stride);
      const float4 t1 = *reinterpret_cast<const float4*>(p0 + (sl + 1) * split_stride);
      const float4 t2 = *reinterpret_cast<const float4*>(p0 + (sl + 2) * split_stride);
      const float4 t3 = *reinterpret_cast<const float4*>(p0 + (sl + 3) * split_stride);
      a.x = (((a.x + t0.x) + t1.x) + t2.x) + t3.x; a.y = (((a.y + t0.y) + t1.y) + t2.y) + t3.y;
      a.z = (((a.z + t0.z) + t1.z) + t2.z) + t3.z; a.w = (((a.w + t0.w) + t1.w) + t2.w) + t3.w;
    }
    for (; sl < splits; ++sl) {
      const float4 t = *reinterpret_cast<const float4*>(p0 + sl * split_stride);
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    if (bias != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + idx);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (gelu) { a.x = gelu_exact(a.x); a.y = gelu_exact(a.y); a.z = gelu_exact(a.z); a.w = gelu_exact(a.w); }
    if (resid != nullptr) {
      const float4 r = *reinterpret_cast<const float4*>(resid + row * N + idx * 4);
      a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
    }
  }
  if (ln_w != nullptr) {   // two-pass LayerNorm over the row (block-uniform branch)
    const float mean = fin_block_sum(live ? (a.x + a.y) + (a.z + a.w) : 0.f, red, warp, lane, nwarps) / static_cast<float>(N);
    const float d0 = a.x - mean, d1 = a.y - mean, d2 = a.z - mean, d3 = a.w - mean;
    const float var = fin_block_sum(live ? (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3) : 0.f, red, warp, lane, nwarps) / static_cast<float>(N);
    const float rstd = rsqrtf(var + eps);
    if (live) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(ln_w) + idx), b = __ldg(reinterpret_cast<const float4*>(ln_b) + idx);
      a.x = d0 * rstd * w.x + b.x; a.y = d1 * rstd * w.y + b.y; a.z = d2 * rstd * w.z + b.z; a.w = d3 * rstd * w.w + b.w;
    }
  }
  if (!live) return;
  if (y != nullptr) *reinterpret_cast<float4*>(y + row * N + idx * 4) = a;
  if (op != nullptr) {
    const float f[4] = {a.x, a.y, a.z, a.w};
    uint32_t hi[2], lo[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const __nv_bfloat16 h0 = __float2bfloat16(f[2 * j]), h1 = __float2bfloat16(f[2 * j + 1]);
      const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * j] - __bfloat162float(h0), f[2 * j + 1] - __bfloat162float(h1));
      hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    __nv_bfloat16* o = op + row * 3 * N + idx * 4;
    *reinterpret_cast<uint2*>(o) = make_uint2(lo[0], lo[1]);
    *reinterpret_cast<uint2*>(o + N) = make_uint2(hi[0], hi[1]);
    *reinterpret_cast<uint2*>(o + 2 * N) = make_uint2(hi[0], hi[1]);
  }
}

__global__ void __launch_bounds__(256) split3_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int N,
                                                            int K) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(N) * K) return;
  const long long n = t / K;
  const int k = static_cast<int>(t - n * K);
  const float v = w[t];
  const __nv_bfloat16 hi = __float2bfloat16(v);
  const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
  __nv_bfloat16* o = out + n * 3 * K;
  o[k] = hi;
  o[K + k] = lo;
  o[2 * K + k] = hi;
}

__global__ void __launch_bounds__(256) time_tanh_kernel(const long long* __restrict__ vmask, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, float* __restrict__ out, int B, int T, int E) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * T) return;
  const int b = warp / T, t = warp - b * T;
  int n = 0;
  for (int i = lane; i < T; i += 32) n += (vmask[static_cast<long long>(b) * T + i] != 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  float g = 0.f;
  if (t < n) {
    float lin = 0.f;
    if (n > 1) {  // torch.linspace(0, 1, n) in fp32: two-sided formula around the midpoint
      const float step = __fdiv_rn(1.0f, static_cast<float>(n - 1));
      lin = (t < n / 2) ? __fmul_rn(step, static_cast<float>(t)) : __fsub_rn(1.0f, __fmul_rn(step, static_cast<float>(n - t - 1)));
    }
    g = __fmul_rn(__fsub_rn(lin, 0.5f), 2.0f);
  }
  float* o = out + static_cast<long long>(warp) * E;
  for (int e = lane; e < E; e += 32) o[e] = tanhf(fmaf(w1[e], g, b1[e]));
}

// E = 512 fixed by the reference (modeling.py:26); generic in E % 128 == 0, E <= 1024 (<= 8 float4 per lane).
__global__ void __launch_bounds__(256) moment_base_kernel(const float* __restrict__ vlin, const float* __restrict__ lnw,
                                                          const float* __restrict__ lnb, const float* __restrict__ that,
                                                          const float* __restrict__ asr_lin, const float* __restrict__ temporal,
                                                          float* __restrict__ base, int B, int T, int E) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * T) return;
  const int b = warp / T;
  const int nv = E >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(vlin + static_cast<long long>(warp) * E);
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) { v[i] = x4[idx]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
  }
  const float mean = warp_sum(s) / static_cast<float>(E);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
      ss += a * a + c * c + d * d + e * e;
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(ss) / static_cast<float>(E) + 1e-12f);
  const float4* w4 = reinterpret_cast<const float4*>(lnw);
  const float4* b4 = reinterpret_cast<const float4*>(lnb);
  const float4* t4 = reinterpret_cast<const float4*>(that + static_cast<long long>(b) * E);
  const float4* a4 = reinterpret_cast<const float4*>(asr_lin + static_cast<long long>(warp) * E);
  const float4* p4 = reinterpret_cast<const float4*>(temporal + static_cast<long long>(warp) * E);
  float4* o4 = reinterpret_cast<float4*>(base + static_cast<long long>(warp) * E);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float4 w = w4[idx], bb = b4[idx], th = t4[idx], p = p4[idx];
      const float4 a = (asr_lin != nullptr) ? a4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);   // ASR-free model: modeling.py:167 skipped
      float4 o;
      o.x = (w.x * ((v[i].x - mean) * rstd) + bb.x) * th.x + a.x + p.x;
      o.y = (w.y * ((v[i].y - mean) * rstd) + bb.y) * th.y + a.y + p.y;
      o.z = (w.z * ((v[i].z - mean) * rstd) + bb.z) * th.z + a.z + p.z;
      o.w = (w.w * ((v[i].w - mean) * rstd) + bb.w) * th.w + a.w + p.w;
      o4[idx] = o;
    }
  }
}

__global__ void __launch_bounds__(256) moment_embed_kernel(const float* __restrict__ base, const float* __restrict__ bemb,
                                                           const float* __restrict__ memb, const long long* __restrict__ bm,
                                                           const long long* __restrict__ mm, float* __restrict__ f, long long rows,
                                                           int E) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int nv = E >> 2;
  if (t >= rows * nv) return;
  const long long r = t / nv;
  const int i = static_cast<int>(t - r * nv);
  float4 v = reinterpret_cast<const float4*>(base)[t];
  if (bm != nullptr) {
    const float4 e = reinterpret_cast<const float4*>(bemb + (bm[r] != 0 ? E : 0))[i];
    v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
  }
  const float4 e = reinterpret_cast<const float4*>(memb + (mm[r] != 0 ? E : 0))[i];
  v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
  reinterpret_cast<float4*>(f)[t] = v;
}

__global__ void __launch_bounds__(256) moment_heads_kernel(const float* __restrict__ feats, const float* __restrict__ w3,
                                                           const float* __restrict__ b3, float* __restrict__ logits, long long rows,
                                                           int Hd) {
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* x = feats + warp * Hd;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int k = lane; k < Hd; k += 32) {
    const float v = x[k];
    a0 = fmaf(v, w3[k], a0);
    a1 = fmaf(v, w3[Hd + k], a1);
    a2 = fmaf(v, w3[2 * Hd + k], a2);
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  if (lane == 0) {
    logits[warp * 3 + 0] = a0 + b3[0];
    logits[warp * 3 + 1] = a1 + b3[1];
    logits[warp * 3 + 2] = a2 + b3[2];
  }
}

// one warp per (sample, head): first-max argmax over T with the -1e10 fill on padded frames
__global__ void __launch_bounds__(64) mr_argmax_kernel(const float* __restrict__ logits, const long long* __restrict__ vmask,
                                                       long long* __restrict__ pred, int B, int T) {
  const int b = blockIdx.x, which = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int t = lane; t < T; t += 32) {
    float v = logits[(static_cast<long long>(b) * T + t) * 3 + which];
    if (vmask[static_cast<long long>(b) * T + t] == 0) v = -1e10f;
    if (v > best) { best = v; arg = t; }  // ascending t per lane: keeps the first maximum
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (lane == 0) pred[b * 2 + which] = arg;
}

__global__ void __launch_bounds__(256) ms_step_kernel(const float* __restrict__ logits, long long* __restrict__ moment_mask,
                                                      long long* __restrict__ boundary_mask, int* __restrict__ steps,
                                                      int* __restrict__ nsteps, int max_steps, int T, double threshold,
                                                      float* __restrict__ probs_out) {
  extern __shared__ float sp[];  // [T] probabilities
  __shared__ float red[8];
  __shared__ int redi[8];
  __shared__ int bounds[2];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long* mm = moment_mask + static_cast<long long>(b) * T;
  long long* bm = boundary_mask + static_cast<long long>(b) * T;
  // masked logits and their maximum
  float best = -INFINITY;
  for (int t = tid; t < T; t += 256) {
    float v = logits[(static_cast<long long>(b) * T + t) * 3 + 2];
    if (mm[t] == 0) v = -FLT_MAX;
    sp[t] = v;
    best = fmaxf(best, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0) red[warp] = best;
  __syncthreads();
  best = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) best = fmaxf(best, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int t = tid; t < T; t += 256) {
    const float e = expf(sp[t] - best);
    sp[t] = e;
    s += e;
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  __syncthreads();
  // probabilities and their FIRST maximum (torch.argmax on the softmax output, modeling.py:397)
  float pbest = -1.f;
  int arg = 0x7fffffff;
  for (int t = tid; t < T; t += 256) {
    const float pr = sp[t] / tot;
    sp[t] = pr;
    if (probs_out != nullptr) probs_out[static_cast<long long>(b) * T + t] = pr;
    if (pr > pbest) { pbest = pr; arg = t; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, pbest, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > pbest || (ob == pbest && oa < arg)) { pbest = ob; arg = oa; }
  }
  if (lane == 0) { red[warp] = pbest; redi[warp] = arg; }
  __syncthreads();
  pbest = red[0]; arg = redi[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) {
    if (red[w] > pbest || (red[w] == pbest && redi[w] < arg)) { pbest = red[w]; arg = redi[w]; }
  }
  if (tid == 0) {
    const double mx = static_cast<double>(sp[arg]);
    int l = arg, r = arg;
    bool accept = !(mx < 0.00001);
    if (accept) {
      while (static_cast<double>(sp[l]) / mx > threshold) { if (l == 0) break; --l; }
      while (static_cast<double>(sp[r]) / mx > threshold) { if (r == T - 1) break; ++r; }
      if (l == 0 || r == 0) accept = false;
    }
    bounds[0] = accept ? l : -1;
    bounds[1] = r;
    if (accept) {
      const int k = nsteps[b];
      if (k < max_steps) { steps[(b * max_steps + k) * 2] = l; steps[(b * max_steps + k) * 2 + 1] = r; nsteps[b] = k + 1; }
    }
  }
  __syncthreads();
  const int l = bounds[0], r = bounds[1];
  if (l >= 0) {
    for (int t = l + tid; t <= r; t += 256) mm[t] = 0;
    if (tid == 0) { bm[l] = 1; bm[r] = 1; }
  }
}

__global__ void __launch_bounds__(256) trim_feats_kernel(const float* __restrict__ x, const long long* __restrict__ mask,
                                                         float* __restrict__ out, int T, int C, int F) {
  __shared__ int src[64];  // source frame of each output slot (F <= 64)
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int t = 0; t < T; ++t) n += (mask[static_cast<long long>(b) * T + t] == 1);
    // k-th selected frame -> slots [(k*F)/n, ((k+1)*F)/n) when n <= F, else the first F selected frames
    int k = 0;
    for (int j = 0; j < F; ++j) src[j] = -1;
    for (int t = 0; t < T; ++t) {
      if (mask[static_cast<long long>(b) * T + t] != 1) continue;
      if (n > F) { if (k < F) src[k] = t; }
      else { for (int j = (k * F) / n; j < ((k + 1) * F) / n; ++j) src[j] = t; }
      ++k;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < F * C; i += 256) {
    const int j = i / C, c = i - j * C;
    const int t = src[j];
    out[(static_cast<long long>(b) * F + j) * C + c] = (t >= 0) ? x[(static_cast<long long>(b) * T + t) * C + c] : 0.f;
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// caption decoder
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_embed_kernel(const long long* __restrict__ tok, const float* __restrict__ word_emb,
                                                        const float* __restrict__ pos_emb, const float* __restrict__ lnw,
                                                        const float* __restrict__ lnb, int pos, float* __restrict__ x,
                                                        __nv_bfloat16* __restrict__ op, int R, int Hd) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float* e = word_emb + tok[warp] * Hd;
  const float* pe = pos_emb + static_cast<long long>(pos) * Hd;
  float v[32];  // Hd <= 1024
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = lane + i * 32;
    v[i] = (k < Hd) ? e[k] + pe[k] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / static_cast<float>(Hd);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = lane + i * 32;
    if (k < Hd) { const float d = v[i] - mean; ss += d * d; }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(ss) / static_cast<float>(Hd) + 1e-12f);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = lane + i * 32;
    if (k < Hd) {
      const float y = lnw[k] * ((v[i] - mean) * rstd) + lnb[k];
      x[static_cast<long long>(warp) * Hd + k] = y;
      if (op != nullptr) {   // the [lo | hi | hi] operand of the first layer's QKV GEMM (what split3_act_kernel would write)
        const __nv_bfloat16 hi = __float2bfloat16(y), lo = __float2bfloat16(y - __bfloat162float(hi));
        __nv_bfloat16* o = op + static_cast<long long>(warp) * 3 * Hd + k;
        o[0] = lo; o[Hd] = hi; o[2 * Hd] = hi;
      }
    }
  }
}

__global__ void __launch_bounds__(256) dec_cache_append_kernel(const float* __restrict__ qkv, float* __restrict__ kc,
                                                               float* __restrict__ vc, int* __restrict__ row_idx, int pos, int R,
                                                               int Tmax, int Hd) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(R) * Hd) return;
  const long long r = t / Hd;
  const int k = static_cast<int>(t - r * Hd);
  kc[(r * Tmax + pos) * Hd + k] = qkv[r * 3 * Hd + Hd + k];
  vc[(r * Tmax + pos) * Hd + k] = qkv[r * 3 * Hd + 2 * Hd + k];
  if (row_idx != nullptr && k == 0) row_idx[r * Tmax + pos] = static_cast<int>(r);   // this beam's newest key lives in its own cache row
}

// Beam re-ordering without moving the KV cache: the history of new beam j is that of old beam prev_k[j], so its index row is a copy
// of that beam's (idx[row][t] = cache row that holds key t of the row's hypothesis).  One table serves every decoder layer.
__global__ void __launch_bounds__(256) dec_index_advance_kernel(const int* __restrict__ idx_old, int* __restrict__ idx_new,
                                                                const int* __restrict__ prev_k, int len, int R, int beam, int Tmax) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * len) return;
  const int r = t / len, k = t - r * len;
  const int src_r = (r / beam) * beam + prev_k[r];
  idx_new[r * Tmax + k] = idx_old[src_r * Tmax + k];
}

__global__ void __launch_bounds__(256) dec_cache_reorder_kernel(const float* __restrict__ ksrc, const float* __restrict__ vsrc,
                                                                float* __restrict__ kdst, float* __restrict__ vdst,
                                                                const int* __restrict__ prev_k, int len, int R, int beam, int Tmax,
                                                                int Hd) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_row = static_cast<long long>(len) * (Hd / 4);
  if (t >= static_cast<long long>(R) * per_row) return;
  const long long r = t / per_row;
  const long long off = t - r * per_row;
  const long long src_r = (r / beam) * beam + prev_k[r];
  reinterpret_cast<float4*>(kdst + r * Tmax * Hd)[off] = reinterpret_cast<const float4*>(ksrc + src_r * Tmax * Hd)[off];
  reinterpret_cast<float4*>(vdst + r * Tmax * Hd)[off] = reinterpret_cast<const float4*>(vsrc + src_r * Tmax * Hd)[off];
}

__global__ void __launch_bounds__(256) gelu_f32_kernel(float* __restrict__ x, long long n) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n) x[t] = gelu_exact(x[t]);
}

constexpr int BEAM_MAX = 8;

// Beam.advance in two kernels.  (a) one CTA per (instance, beam) row of the logits: log-sum-exp over the vocabulary, then the row's top-`beam` candidates of val = (logit - lse) + beam score under the order "value descending,
// lower vocabulary index first" (thread-local sorted lists -> warp merge by shuffles -> block merge).  (b) one thread per
// instance merges its rows' candidates (lower flat index = row * V + word first on ties), records back-pointers / words / scores
// and the done flag.  (The one-CTA-per-instance form read every row three times with 64 CTAs and finished with a serial scan by
// one thread: 0.35 ms per decode step, 28 % of step captioning.)
constexpr int BR_THREADS = 512;

template <int BEAM>
__device__ __forceinline__ void lane_insert(float (&bv)[BEAM], int (&bi)[BEAM], float val, int idx) {
  if (val > bv[BEAM - 1]) {   // strict: on equal values the earlier (lower-index) candidate stays
    bv[BEAM - 1] = val; bi[BEAM - 1] = idx;
#pragma unroll
    for (int pos = BEAM - 1; pos > 0; --pos) {
      if (bv[pos] > bv[pos - 1]) {
        const float tv = bv[pos]; bv[pos] = bv[pos - 1]; bv[pos - 1] = tv;
        const int ti = bi[pos]; bi[pos] = bi[pos - 1]; bi[pos - 1] = ti;
      }
    }
  }
}

__device__ __forceinline__ bool cand_better(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }

// Warp-wide merge of per-lane sorted candidate lists: returns (in every lane) the warp's top-`beam` in order.
__device__ __forceinline__ void warp_topk(float* bv, int* bi, int beam, float* out_v, int* out_i) {
  const int lane = threadIdx.x & 31;
  int head = 0;
  for (int sel = 0; sel < beam; ++sel) {
    float v = (head < beam) ? bv[head] : -INFINITY;
    int i = (head < beam) ? bi[head] : 0x7fffffff;
    int who = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
      const int w2 = __shfl_xor_sync(0xffffffffu, who, o);
      if (cand_better(v2, i2, v, i) || (v2 == v && i2 == i && w2 < who)) { v = v2; i = i2; who = w2; }
    }
    out_v[sel] = v; out_i[sel] = i;
    if (lane == who) ++head;
  }
}

__global__ void __launch_bounds__(BR_THREADS) beam_rows_generic_kernel(const float* __restrict__ logits, int ldl, int V, const float* __restrict__ scores,
                                                               const int* __restrict__ done, float* __restrict__ cand_v, int* __restrict__ cand_i,
                                                               int step, int beam) {
  __shared__ float red_m[BR_THREADS / 32], red_s[BR_THREADS / 32];
  __shared__ float wv[(BR_THREADS / 32) * BEAM_MAX];
  __shared__ int wi[(BR_THREADS / 32) * BEAM_MAX];
  const int row = blockIdx.x, inst = row / beam, k = row - inst * beam;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (done[inst] || (step == 0 && k > 0)) return;   // Beam.advance uses word_prob[0] only before any back-pointer exists
  const float* lr = logits + static_cast<long long>(row) * ldl;
  // pass 1: exact maximum, then sum of exp(x - max) -> log-sum-exp (the row is 120 KB: the second and third reads hit L1 / L2)
  float m = -INFINITY;
  for (int j = tid; j < V; j += BR_THREADS) m = fmaxf(m, lr[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red_m[warp] = m;
  __syncthreads();
  float M = red_m[0];
#pragma unroll
  for (int w = 1; w < BR_THREADS / 32; ++w) M = fmaxf(M, red_m[w]);
  float ssum = 0.f;
  for (int j = tid; j < V; j += BR_THREADS) ssum += expf(lr[j] - M);
  ssum = warp_sum(ssum);
  if (lane == 0) red_s[warp] = ssum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < BR_THREADS / 32; ++w) tot += red_s[w];
  const float lse = M + logf(tot);
  const float add = (step == 0) ? 0.f : scores[row];
  // pass 2: top-`beam` of val = log_softmax + beam score (beam.py:76), flat index = k * V + word
  float bv[BEAM_MAX];
  int bi[BEAM_MAX];
#pragma unroll
  for (int i = 0; i < BEAM_MAX; ++i) { bv[i] = -INFINITY; bi[i] = 0x7fffffff; }
  for (int j = tid; j < V; j += BR_THREADS) {
    const float val = (lr[j] - lse) + add;
    if (val > bv[beam - 1]) {
      int pos = beam - 1;
      bv[pos] = val; bi[pos] = k * V + j;
      while (pos > 0 && bv[pos] > bv[pos - 1]) {
        const float tv = bv[pos]; bv[pos] = bv[pos - 1]; bv[pos - 1] = tv;
        const int ti = bi[pos]; bi[pos] = bi[pos - 1]; bi[pos - 1] = ti;
        --pos;
      }
    }
  }
  float ov[BEAM_MAX];
  int oi[BEAM_MAX];
  warp_topk(bv, bi, beam, ov, oi);
  if (lane == 0)
    for (int i = 0; i < beam; ++i) { wv[warp * BEAM_MAX + i] = ov[i]; wi[warp * BEAM_MAX + i] = oi[i]; }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < BEAM_MAX; ++i) { bv[i] = -INFINITY; bi[i] = 0x7fffffff; }
    if (lane < BR_THREADS / 32)
      for (int i = 0; i < beam; ++i) { bv[i] = wv[lane * BEAM_MAX + i]; bi[i] = wi[lane * BEAM_MAX + i]; }
    warp_topk(bv, bi, beam, ov, oi);
    if (lane == 0)
      for (int i = 0; i < beam; ++i) { cand_v[row * beam + i] = ov[i]; cand_i[row * beam + i] = oi[i]; }
  }
}

// Same result rules as beam_rows_generic_kernel, with the row held in registers: ONE pass over the logits (float4 loads, up to
// BR_NV per thread in flight) instead of three scalar passes — the generic kernel was 47 us per decode step at 192 rows x 30522.
constexpr int BRR_THREADS = 1024;   // 32 warps: ~60 registers per thread with 8 float4 of logits each (512 x 16 float4 needed 97 registers:
constexpr int BR_NV = 8;            // one CTA per SM either way, but twice the threads share a row); V <= 4 * BR_NV * BRR_THREADS = 32768
template <int BEAM>         // compile-time beam width: the running top-BEAM list stays in registers
__global__ void __launch_bounds__(BRR_THREADS) beam_rows_kernel(const float* __restrict__ logits, int ldl, int V, const float* __restrict__ scores,
                                                               const int* __restrict__ done, float* __restrict__ cand_v, int* __restrict__ cand_i,
                                                               int step) {
  constexpr int beam = BEAM;
  __shared__ float red_m[BRR_THREADS / 32], red_s[BRR_THREADS / 32];
  __shared__ float wv[(BRR_THREADS / 32) * BEAM_MAX];
  __shared__ int wi[(BRR_THREADS / 32) * BEAM_MAX];
  const int row = blockIdx.x, inst = row / beam, k = row - inst * beam;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (done[inst] || (step == 0 && k > 0)) return;   // Beam.advance uses word_prob[0] only before any back-pointer exists
  const float* lr = logits + static_cast<long long>(row) * ldl;
  float4 x[BR_NV];
#pragma unroll
  for (int i = 0; i < BR_NV; ++i) {
    const int j = (i * BRR_THREADS + tid) * 4;
    x[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (j < V) {   // ldl >= round_up(V, 4): the load stays inside the row; columns >= V are padding
      const float4 t = *reinterpret_cast<const float4*>(lr + j);
      x[i].x = t.x;
      if (j + 1 < V) x[i].y = t.y;
      if (j + 2 < V) x[i].z = t.z;
      if (j + 3 < V) x[i].w = t.w;
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < BR_NV; ++i) m = fmaxf(fmaxf(fmaxf(m, x[i].x), fmaxf(x[i].y, x[i].z)), x[i].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red_m[warp] = m;
  __syncthreads();
  float M = red_m[0];
#pragma unroll
  for (int w = 1; w < BRR_THREADS / 32; ++w) M = fmaxf(M, red_m[w]);
  float ssum = 0.f;
#pragma unroll
  for (int i = 0; i < BR_NV; ++i) ssum += (expf(x[i].x - M) + expf(x[i].y - M)) + (expf(x[i].z - M) + expf(x[i].w - M));   // exp(-inf) = 0
  ssum = warp_sum(ssum);
  if (lane == 0) red_s[warp] = ssum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < BRR_THREADS / 32; ++w) tot += red_s[w];
  const float lse = M + logf(tot);
  const float add = (step == 0) ? 0.f : scores[row];
  // top-`beam` of val = log_softmax + beam score (beam.py:76), flat index = k * V + word; a thread meets its words in index order
  float tv[BEAM];
  int ti[BEAM];
#pragma unroll
  for (int i = 0; i < BEAM; ++i) { tv[i] = -INFINITY; ti[i] = 0x7fffffff; }
#pragma unroll
  for (int i = 0; i < BR_NV; ++i) {
    const int j = (i * BRR_THREADS + tid) * 4;
    const float xs[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float val = (xs[c] - lse) + add;   // -inf for padding: never better than anything
      if (val > tv[BEAM - 1]) {                // insertion by bubbling up (strict >: the earlier word keeps its place on ties)
        tv[BEAM - 1] = val; ti[BEAM - 1] = k * V + j + c;
#pragma unroll
        for (int pp = BEAM - 1; pp > 0; --pp) {
          if (tv[pp] > tv[pp - 1]) {
            const float fv = tv[pp]; tv[pp] = tv[pp - 1]; tv[pp - 1] = fv;
            const int fi = ti[pp]; ti[pp] = ti[pp - 1]; ti[pp - 1] = fi;
          }
        }
      }
    }
  }
  float bv[BEAM_MAX];
  int bi[BEAM_MAX];
#pragma unroll
  for (int i = 0; i < BEAM_MAX; ++i) { bv[i] = i < BEAM ? tv[i < BEAM ? i : 0] : -INFINITY; bi[i] = i < BEAM ? ti[i < BEAM ? i : 0] : 0x7fffffff; }
  float ov[BEAM_MAX];
  int oi[BEAM_MAX];
  warp_topk(bv, bi, beam, ov, oi);
  if (lane == 0)
    for (int i = 0; i < beam; ++i) { wv[warp * BEAM_MAX + i] = ov[i]; wi[warp * BEAM_MAX + i] = oi[i]; }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < BEAM_MAX; ++i) { bv[i] = -INFINITY; bi[i] = 0x7fffffff; }
    if (lane < BRR_THREADS / 32)
      for (int i = 0; i < beam; ++i) { bv[i] = wv[lane * BEAM_MAX + i]; bi[i] = wi[lane * BEAM_MAX + i]; }
    warp_topk(bv, bi, beam, ov, oi);
    if (lane == 0)
      for (int i = 0; i < beam; ++i) { cand_v[row * beam + i] = ov[i]; cand_i[row * beam + i] = oi[i]; }
  }
}

__global__ void __launch_bounds__(128) beam_merge_kernel(const float* __restrict__ cand_v, const int* __restrict__ cand_i, int V,
                                                         float* __restrict__ scores, int* __restrict__ done, int* __restrict__ nsteps,
                                                         int* __restrict__ prev_k_rec, int* __restrict__ ys_rec, long long* __restrict__ tok,
                                                         int step, int n_inst, int beam, int eos) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= n_inst) return;
  int* pk_out = prev_k_rec + (static_cast<long long>(step) * n_inst + inst) * beam;
  int* ys_out = ys_rec + (static_cast<long long>(step) * n_inst + inst) * beam;
  if (done[inst]) {  // finished: frozen; identity back-pointers keep the cache re-order a no-op
    for (int i = 0; i < beam; ++i) { pk_out[i] = i; ys_out[i] = 0; }
    return;
  }
  const int rows = (step == 0) ? 1 : beam;
  int head[BEAM_MAX];
  for (int k = 0; k < BEAM_MAX; ++k) head[k] = 0;
  const float* cv = cand_v + static_cast<long long>(inst) * beam * beam;
  const int* ci = cand_i + static_cast<long long>(inst) * beam * beam;
  for (int sel = 0; sel < beam; ++sel) {
    int bk = -1;
    for (int k = 0; k < rows; ++k) {
      if (head[k] >= beam) continue;
      if (bk < 0 || cand_better(cv[k * beam + head[k]], ci[k * beam + head[k]], cv[bk * beam + head[bk]], ci[bk * beam + head[bk]])) bk = k;
    }
    const float val = cv[bk * beam + head[bk]];
    const int id = ci[bk * beam + head[bk]];
    ++head[bk];
    const int pk = id / V, y = id - pk * V;
    scores[inst * beam + sel] = val;
    pk_out[sel] = pk;
    ys_out[sel] = y;
    tok[inst * beam + sel] = y;
    if (sel == 0 && y == eos) done[inst] = 1;
  }
  nsteps[inst] = step + 1;
}

inline unsigned nblocks(long long n, int per) { return static_cast<unsigned>((n + per - 1) / per); }

}  // namespace

int split3_act_launch(const float* x, __nv_bfloat16* out, long long rows, int K, int gelu, cudaStream_t s) {
  if (rows <= 0) return 0;
  if (K % 8 != 0) return -7;
  split3_act_kernel<<<nblocks(rows * (K / 8), 256), 256, 0, s>>>(x, out, rows, K, gelu);
  return static_cast<int>(cudaGetLastError());
}
int splitk_finish_launch(const float* part, long long split_stride, int splits, const float* bias, const float* resid,
                         const float* ln_w, const float* ln_b, float eps, int gelu, float* y, __nv_bfloat16* op, int rows, int N,
                         cudaStream_t s) {
  if (rows <= 0) return 0;
  if (N % 4 != 0 || N > FIN_MAX_THREADS * 4 || splits < 1) return -7;
  const int threads = ((N / 4 + 31) / 32) * 32;
  splitk_finish_kernel<<<static_cast<unsigned>(rows), threads, 0, s>>>(part, split_stride, splits, bias, resid, ln_w, ln_b, eps, gelu, y, op, N);
  return static_cast<int>(cudaGetLastError());
}
int split3_weight_launch(const float* w, __nv_bfloat16* out, int N, int K, cudaStream_t s) {
  split3_weight_kernel<<<nblocks(static_cast<long long>(N) * K, 256), 256, 0, s>>>(w, out, N, K);
  return static_cast<int>(cudaGetLastError());
}
int time_tanh_launch(const long long* video_mask, const float* w1, const float* b1, float* out, int B, int T, int E,
                     cudaStream_t s) {
  time_tanh_kernel<<<nblocks(static_cast<long long>(B) * T, 8), 256, 0, s>>>(video_mask, w1, b1, out, B, T, E);
  return static_cast<int>(cudaGetLastError());
}
int moment_base_launch(const float* vlin, const float* lnw, const float* lnb, const float* that, const float* asr_lin,
                       const float* temporal, float* base, int B, int T, int E, cudaStream_t s) {
  if (E % 4 != 0 || E > 1024) return -7;
  moment_base_kernel<<<nblocks(static_cast<long long>(B) * T, 8), 256, 0, s>>>(vlin, lnw, lnb, that, asr_lin, temporal, base, B, T, E);
  return static_cast<int>(cudaGetLastError());
}
int moment_embed_launch(const float* base, const float* bemb, const float* memb, const long long* bm, const long long* mm,
                        float* f, long long rows, int E, cudaStream_t s) {
  if (E % 4 != 0) return -7;
  moment_embed_kernel<<<nblocks(rows * (E / 4), 256), 256, 0, s>>>(base, bemb, memb, bm, mm, f, rows, E);
  return static_cast<int>(cudaGetLastError());
}
int moment_heads_launch(const float* feats, const float* w3, const float* b3, float* logits, long long rows, int Hd,
                        cudaStream_t s) {
  moment_heads_kernel<<<nblocks(rows, 8), 256, 0, s>>>(feats, w3, b3, logits, rows, Hd);
  return static_cast<int>(cudaGetLastError());
}
int mr_argmax_launch(const float* logits, const long long* vmask, long long* pred, int B, int T, cudaStream_t s) {
  mr_argmax_kernel<<<B, 64, 0, s>>>(logits, vmask, pred, B, T);
  return static_cast<int>(cudaGetLastError());
}
int ms_step_launch(const float* logits, long long* moment_mask, long long* boundary_mask, int* steps, int* nsteps, int max_steps,
                   int B, int T, double threshold, float* probs_out, cudaStream_t s) {
  if (T > 8192) return -7;
  ms_step_kernel<<<B, 256, static_cast<size_t>(T) * 4, s>>>(logits, moment_mask, boundary_mask, steps, nsteps, max_steps, T,
                                                             threshold, probs_out);
  return static_cast<int>(cudaGetLastError());
}
int trim_feats_launch(const float* x, const long long* mask, float* out, int B, int T, int C, int F, cudaStream_t s) {
  if (F > 64) return -7;
  trim_feats_kernel<<<B, 256, 0, s>>>(x, mask, out, T, C, F);
  return static_cast<int>(cudaGetLastError());
}

int dec_embed_launch(const long long* tok, const float* word_emb, const float* pos_emb, const float* lnw, const float* lnb, int pos,
                     float* x, __nv_bfloat16* op, int R, int Hd, cudaStream_t s) {
  if (Hd > 1024) return -7;
  dec_embed_kernel<<<nblocks(R, 8), 256, 0, s>>>(tok, word_emb, pos_emb, lnw, lnb, pos, x, op, R, Hd);
  return static_cast<int>(cudaGetLastError());
}
int dec_cache_append_launch(const float* qkv, float* kc, float* vc, int* row_idx, int pos, int R, int Tmax, int Hd, cudaStream_t s) {
  dec_cache_append_kernel<<<nblocks(static_cast<long long>(R) * Hd, 256), 256, 0, s>>>(qkv, kc, vc, row_idx, pos, R, Tmax, Hd);
  return static_cast<int>(cudaGetLastError());
}
int dec_index_advance_launch(const int* idx_old, int* idx_new, const int* prev_k, int len, int R, int beam, int Tmax, cudaStream_t s) {
  if (len <= 0) return 0;
  dec_index_advance_kernel<<<nblocks(static_cast<long long>(R) * len, 256), 256, 0, s>>>(idx_old, idx_new, prev_k, len, R, beam, Tmax);
  return static_cast<int>(cudaGetLastError());
}
int dec_cache_reorder_launch(const float* ksrc, const float* vsrc, float* kdst, float* vdst, const int* prev_k, int len, int R,
                             int beam, int Tmax, int Hd, cudaStream_t s) {
  if (len <= 0) return 0;
  dec_cache_reorder_kernel<<<nblocks(static_cast<long long>(R) * len * (Hd / 4), 256), 256, 0, s>>>(ksrc, vsrc, kdst, vdst, prev_k, len,
                                                                                                    R, beam, Tmax, Hd);
  return static_cast<int>(cudaGetLastError());
}
int gelu_f32_launch(float* x, long long n, cudaStream_t s) {
  gelu_f32_kernel<<<nblocks(n, 256), 256, 0, s>>>(x, n);
  return static_cast<int>(cudaGetLastError());
}
int beam_advance_launch(const float* logits, int ldl, int V, float* scores, int* done, int* nsteps, int* prev_k_rec, int* ys_rec,
                        long long* tok, int step, int n_inst, int beam, int eos, float* cand_v, int* cand_i, cudaStream_t s) {
  if (beam < 1 || beam > BEAM_MAX || cand_v == nullptr || cand_i == nullptr) return -7;
  if (V <= 4 * BR_NV * BRR_THREADS && ldl % 4 == 0 && ldl >= (V + 3) / 4 * 4 && (reinterpret_cast<uintptr_t>(logits) & 15u) == 0)
    switch (beam) {
#define HB_BR(B) case B: beam_rows_kernel<B><<<n_inst * beam, BRR_THREADS, 0, s>>>(logits, ldl, V, scores, done, cand_v, cand_i, step); break;
      HB_BR(1) HB_BR(2) HB_BR(3) HB_BR(4) HB_BR(5) HB_BR(6) HB_BR(7) HB_BR(8)
#undef HB_BR
    }
  else
    beam_rows_generic_kernel<<<n_inst * beam, BR_THREADS, 0, s>>>(logits, ldl, V, scores, done, cand_v, cand_i, step, beam);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  beam_merge_kernel<<<(n_inst + 127) / 128, 128, 0, s>>>(cand_v, cand_i, V, scores, done, nsteps, prev_k_rec, ys_rec, tok, step, n_inst, beam,
                                                        eos);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace hb
