// hb_preproc.cu — frame preprocessing on the GPU: Resize(S, BICUBIC) + CenterCrop(S) of decoded
// uint8 RGB frames, bit-identical to the reference's CPU pipeline
//   torchvision Resize/CenterCrop on PIL images  (EVA_clip/eva_clip.py:144-147, used at inference_video_retrieval.py:43-49
//   and extract_features.py:48-50)
// whose arithmetic is Pillow's 8-bit ImagingResample: separable, horizontal pass first into an 8-bit intermediate, bicubic
// (a = -0.5) weights evaluated in double, normalised, converted to 22-bit fixed point, accumulated in int32 from 1 << 21 and
// shifted back with saturation.  The weights are computed on the host in double exactly as Pillow does (integer tables, cached
// per source size); the kernel does only integer work.  ToTensor + Normalize stay folded into the patch gather
// (hb_vit_encode_u8), so a decoded frame goes H2D once as bytes and is never materialised in fp32.
//
// One fused kernel: a CTA owns TY output rows of one frame.  It streams the source rows those output rows need through
// shared memory (coalesced 4-byte loads of only the column span the crop needs), resamples each horizontally into an
// 8-bit smem intermediate, then resamples vertically out of smem and writes the planar [3,S,S] crop.  The intermediate
// image of the two-pass reference never touches HBM.  HBM-bound: algorithmic bytes = needed source region + output.
#include "hb_preproc.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace hb {

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;
constexpr int G = 4;  // source rows resampled per smem stage

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for the full-image box; only outputs [first, first + count) are kept.
void axis_tables(int in_size, int out_size, int first, int count, std::vector<int>& bounds, std::vector<int>& coeffs, int& ksize) {
  if (in_size == out_size) {  // Pillow skips the pass; the identity table gives the same bytes
    ksize = 1;
    bounds.resize(static_cast<size_t>(count) * 2);
    coeffs.resize(count);
    for (int i = 0; i < count; ++i) { bounds[2 * i] = first + i; bounds[2 * i + 1] = 1; coeffs[i] = 1 << PRECISION_BITS; }
    return;
  }
  double scale, filterscale;
  filterscale = scale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  bounds.assign(static_cast<size_t>(count) * 2, 0);
  coeffs.assign(static_cast<size_t>(count) * ksize, 0);
  std::vector<double> w(ksize);
  for (int i = 0; i < count; ++i) {
    const int xx = first + i;
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      double k = w[x];
      if (ww != 0.0) k /= ww;
      coeffs[static_cast<size_t>(i) * ksize + x] =
          (k < 0) ? static_cast<int>(-0.5 + k * (1 << PRECISION_BITS)) : static_cast<int>(0.5 + k * (1 << PRECISION_BITS));
    }
    bounds[2 * i] = xmin;
    bounds[2 * i + 1] = xmax;
  }
}

__device__ __forceinline__ int clip8(int v) {
  v >>= PRECISION_BITS;
  return min(max(v, 0), 255);
}

struct KParams {
  const uint8_t* src;  // [B, H, W, 3]
  uint8_t* dst;        // [B, 3, S, S]
  const int* hb;       // [S, 2]  (first source column - x0, taps)
  const int* hk;       // [S, kh]
  const int* vb;       // [S, 2]  (first source row, taps)
  const int* vk;       // [S, kv]
  int H, W, S, kh, kv, x0, span_bytes, row_pitch /*smem bytes per staged source row*/, ty, max_rows, tmp_pitch;
};

__global__ void __launch_bounds__(256) resize_crop_kernel(const KParams p) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint8_t* stage = sm;                           // [G][row_pitch]
  uint8_t* tmp = sm + G * p.row_pitch;           // [max_rows][tmp_pitch]  horizontally resampled rows (x-major, channel-minor)
  const int b = blockIdx.y;
  const int y0 = blockIdx.x * p.ty;
  const int ny = min(p.ty, p.S - y0);
  const int r0 = p.vb[2 * y0];
  const int r1 = p.vb[2 * (y0 + ny - 1)] + p.vb[2 * (y0 + ny - 1) + 1];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const uint8_t* frame = p.src + static_cast<size_t>(b) * p.H * p.W * 3;

  // this thread's output column
  const int x = tid;
  int hx0 = 0, hn = 0;
  const int* hkx = nullptr;
  if (x < p.S) { hx0 = p.hb[2 * x] * 3; hn = p.hb[2 * x + 1]; hkx = p.hk + static_cast<size_t>(x) * p.kh; }

  for (int r = r0; r < r1; r += G) {
    const int nr = min(G, r1 - r);
    // ---- stage nr source rows (only the byte span the crop needs); words inside the span go as 4-byte loads ----
    for (int j = 0; j < nr; ++j) {
      const uint8_t* g = frame + (static_cast<size_t>(r + j) * p.W + p.x0) * 3;
      const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(g) & 3u);
      uint8_t* s = stage + j * p.row_pitch + mis;  // s[i] = g[i]; s + head is 4-byte aligned
      const int len = p.span_bytes;
      int head = mis ? 4 - mis : 0;
      if (head > len) head = len;
      const int nw = (len - head) >> 2;
      if (tid < head) s[tid] = g[tid];
      const uint32_t* gw = reinterpret_cast<const uint32_t*>(g + head);
      uint32_t* sw = reinterpret_cast<uint32_t*>(s + head);
      for (int i = tid; i < nw; i += nthr) sw[i] = __ldg(gw + i);
      const int done = head + 4 * nw;
      if (tid < len - done) s[done + tid] = g[done + tid];
    }
    __syncthreads();
    // ---- horizontal pass: thread x produces 3 channels of nr rows; each weight is loaded once for 3 * nr products ----
    if (x < p.S) {
      int acc[G][3];
#pragma unroll
      for (int j = 0; j < G; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 1 << (PRECISION_BITS - 1);
      const uint8_t* srow[G];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const uint8_t* g = frame + (static_cast<size_t>(r + min(j, nr - 1)) * p.W + p.x0) * 3;
        srow[j] = stage + min(j, nr - 1) * p.row_pitch + static_cast<int>(reinterpret_cast<uintptr_t>(g) & 3u) + hx0;
      }
      for (int k = 0; k < hn; ++k) {
        const int w = __ldg(hkx + k);
#pragma unroll
        for (int j = 0; j < G; ++j) {
          acc[j][0] += static_cast<int>(srow[j][3 * k + 0]) * w;
          acc[j][1] += static_cast<int>(srow[j][3 * k + 1]) * w;
          acc[j][2] += static_cast<int>(srow[j][3 * k + 2]) * w;
        }
      }
#pragma unroll
      for (int j = 0; j < G; ++j) {
        if (j < nr) {
          uint8_t* t = tmp + static_cast<size_t>(r - r0 + j) * p.tmp_pitch + 3 * x;
          t[0] = static_cast<uint8_t>(clip8(acc[j][0]));
          t[1] = static_cast<uint8_t>(clip8(acc[j][1]));
          t[2] = static_cast<uint8_t>(clip8(acc[j][2]));
        }
      }
    }
    __syncthreads();
  }
  // ---- vertical pass out of smem; planar output, coalesced along x ----
  if (x < p.S) {
    for (int yy = 0; yy < ny; ++yy) {
      const int y = y0 + yy;
      const int v0 = p.vb[2 * y] - r0, vn = p.vb[2 * y + 1];
      const int* vky = p.vk + static_cast<size_t>(y) * p.kv;
      int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
      const uint8_t* t = tmp + static_cast<size_t>(v0) * p.tmp_pitch + 3 * x;
      for (int k = 0; k < vn; ++k) {
        const int w = __ldg(vky + k);
        a0 += static_cast<int>(t[0]) * w;
        a1 += static_cast<int>(t[1]) * w;
        a2 += static_cast<int>(t[2]) * w;
        t += p.tmp_pitch;
      }
      uint8_t* d = p.dst + (static_cast<size_t>(b) * 3 * p.S + y) * p.S + x;
      const size_t plane = static_cast<size_t>(p.S) * p.S;
      d[0] = static_cast<uint8_t>(clip8(a0));
      d[plane] = static_cast<uint8_t>(clip8(a1));
      d[2 * plane] = static_cast<uint8_t>(clip8(a2));
    }
  }
}


// ---- kernels v2 / v3 (cross-checks, not the default): word loads + dp4a / + byte extraction -----------------------------------------------------------------------------------
// v1 spends one LDS.U8 and one IMAD per (pixel, tap, channel) and is LSU-bound at 7 % of the HBM roofline (360p).  v2 keeps the
// arithmetic exact and moves four taps per instruction:
//   * staged source rows are de-interleaved into R / G / B planes in shared memory, so the taps of one output column are
//     consecutive BYTES of one plane row: ceil((misalignment + taps) / 4) aligned 4-byte words;
//   * a 22-bit weight w is split into byte planes w = b0 + 256 b1 + 65536 b2 (b0, b1 unsigned, b2 signed), laid out on the host
//     in the same word alignment (zero weights on the slack bytes), so  sum p w = dp4a(p, b0) + (dp4a(p, b1) << 8) + (dp4a(p, b2)
//     << 16)  in wrap-around int32 — the value Pillow's int32 accumulator holds;
//   * the horizontal pass writes its 8-bit results TRANSPOSED, tmp[channel][column][source row], four source rows per word (groups
//     of 4 rows aligned to absolute row numbers), so the vertical pass reads its taps as words too; the row weights are
//     pre-aligned the same way (the misalignment of a row's first tap is the same for every column).
// Rows / pixels past the image or the crop's span only ever meet zero weights.
struct KParams2 {
  const uint8_t* src;
  uint8_t* dst;
  const int* hw0;          // [S]
  const uint32_t* hwt;     // [S][3][nwh]
  const int* vw0;          // [S]
  const uint32_t* vwt;     // [S][3][nwv]
  const int* hwi;          // [S][4 nwh]  the same taps as plain 22-bit weights, one per byte of the words (kernel v3), 16-byte aligned rows
  const int* vwi;          // [S][4 nwv]
  int H, W, S, x0, span_bytes, row_pitch, ng4, pw, tw, nwh, nwv, ty;
};

__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {   // a unsigned bytes, b signed bytes
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// DP4A = false is kernel v3: the same layout and word loads, but each byte of a word is extracted (ALU pipe) and multiplied by its
// plain 22-bit weight with an IMAD — dp4a issues at a fraction of the IMAD rate on this part (v2 measured dp4a-bound).
__device__ __forceinline__ int mac4(uint32_t px, const int4& w, int acc) {
  acc += static_cast<int>(px & 0xffu) * w.x;
  acc += static_cast<int>((px >> 8) & 0xffu) * w.y;
  acc += static_cast<int>((px >> 16) & 0xffu) * w.z;
  acc += static_cast<int>(px >> 24) * w.w;
  return acc;
}

template <bool DP4A>
__global__ void __launch_bounds__(256) resize_crop2_kernel(const KParams2 p) {
  extern __shared__ __align__(16) uint8_t sm[];
  uint8_t* stage = sm;                                                          // [G][row_pitch]  interleaved source rows
  uint32_t* planes = reinterpret_cast<uint32_t*>(sm + G * p.row_pitch);        // [3][G][pw]      R / G / B planes of those rows
  uint32_t* tmp = planes + 3 * G * p.pw;                                        // [3][S][tw]      resampled columns, 4 source rows per word
  const int b = blockIdx.y;
  const int y0 = blockIdx.x * p.ty;
  const int ny = min(p.ty, p.S - y0);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const uint8_t* frame = p.src + static_cast<size_t>(b) * p.H * p.W * 3;
  const int r0a = p.vw0[y0];                                                    // first source row of the window (multiple of 4)
  int r_end = 0;                                                                // one past the last row any output row of the tile reads
  for (int yy = 0; yy < ny; ++yy) r_end = max(r_end, p.vw0[y0 + yy] + 4 * p.nwv);
  r_end = min(r_end, (p.H + 3) & ~3);
  const int n_groups = (r_end - r0a) >> 2;

  const int x = tid;
  const bool active = x < p.S;
  int hw0 = 0;
  const uint32_t* hwx = p.hwt;
  if (active) { hw0 = p.hw0[x]; hwx = p.hwt + static_cast<size_t>(x) * 3 * p.nwh; }

  for (int g = 0; g < n_groups; ++g) {
    const int r = r0a + 4 * g;
    // ---- stage 4 source rows (only the byte span the crop needs), aligned 4-byte loads
    for (int j = 0; j < G; ++j) {
      const int row = min(r + j, p.H - 1);
      const uint8_t* gp = frame + (static_cast<size_t>(row) * p.W + p.x0) * 3;
      const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(gp) & 3u);
      uint8_t* s = stage + j * p.row_pitch + mis;  // s[i] = gp[i]; s + head is 4-byte aligned
      const int len = p.span_bytes;
      int head = mis ? 4 - mis : 0;
      if (head > len) head = len;
      const int nw = (len - head) >> 2;
      if (tid < head) s[tid] = gp[tid];
      const uint32_t* gw = reinterpret_cast<const uint32_t*>(gp + head);
      uint32_t* sw = reinterpret_cast<uint32_t*>(s + head);
      for (int i = tid; i < nw; i += nthr) sw[i] = __ldg(gw + i);
      const int done = head + 4 * nw;
      if (tid < len - done) s[done + tid] = gp[done + tid];
    }
    __syncthreads();
    // ---- de-interleave: 4 pixels (12 bytes at byte offset mis + 12 t) -> one word of each plane
    for (int j = 0; j < G; ++j) {
      const int row = min(r + j, p.H - 1);
      const uint32_t sh = (static_cast<uint32_t>(reinterpret_cast<uintptr_t>(frame + (static_cast<size_t>(row) * p.W + p.x0) * 3)) & 3u) * 8u;
      const uint32_t* srow = reinterpret_cast<const uint32_t*>(stage + j * p.row_pitch);
      for (int t = tid; t < p.ng4; t += nthr) {
        const uint32_t w0 = srow[3 * t], w1 = srow[3 * t + 1], w2 = srow[3 * t + 2], w3 = srow[3 * t + 3];
        const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh), a2 = __funnelshift_r(w2, w3, sh);
        // a0 = R0 G0 B0 R1 | a1 = G1 B1 R2 G2 | a2 = B2 R3 G3 B3   (byte 0 first)
        planes[(0 * G + j) * p.pw + t] = __byte_perm(__byte_perm(a0, a1, 0x0630), a2, 0x5210);
        planes[(1 * G + j) * p.pw + t] = __byte_perm(__byte_perm(a0, a1, 0x0741), a2, 0x6210);
        planes[(2 * G + j) * p.pw + t] = __byte_perm(__byte_perm(a0, a1, 0x0052), a2, 0x7410);
      }
    }
    __syncthreads();
    // ---- horizontal pass: thread x, 3 channels x 4 rows, four taps per dp4a
    if (active) {
      if constexpr (DP4A) {
        int a0[G][3], a1[G][3], a2[G][3];
#pragma unroll
        for (int j = 0; j < G; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c) { a0[j][c] = 0; a1[j][c] = 0; a2[j][c] = 0; }
        for (int n = 0; n < p.nwh; ++n) {
          const uint32_t b0 = __ldg(hwx + n), b1 = __ldg(hwx + p.nwh + n), b2 = __ldg(hwx + 2 * p.nwh + n);
#pragma unroll
          for (int j = 0; j < G; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const uint32_t px = planes[(c * G + j) * p.pw + hw0 + n];
              a0[j][c] = dp4a_uu(px, b0, a0[j][c]);
              a1[j][c] = dp4a_uu(px, b1, a1[j][c]);
              a2[j][c] = dp4a_us(px, b2, a2[j][c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t word = 0;
#pragma unroll
          for (int j = 0; j < G; ++j) {
            const int v = (1 << (PRECISION_BITS - 1)) + a0[j][c] + (a1[j][c] << 8) + (a2[j][c] << 16);
            word |= static_cast<uint32_t>(clip8(v)) << (8 * j);
          }
          tmp[(static_cast<size_t>(c) * p.S + x) * p.tw + g] = word;
        }
      } else {
        int acc[G][3];
#pragma unroll
        for (int j = 0; j < G; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[j][c] = 1 << (PRECISION_BITS - 1);
        const int4* wx = reinterpret_cast<const int4*>(p.hwi + static_cast<size_t>(x) * 4 * p.nwh);
        for (int n = 0; n < p.nwh; ++n) {
          const int4 w = __ldg(wx + n);
#pragma unroll
          for (int j = 0; j < G; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[j][c] = mac4(planes[(c * G + j) * p.pw + hw0 + n], w, acc[j][c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t word = 0;
#pragma unroll
          for (int j = 0; j < G; ++j) word |= static_cast<uint32_t>(clip8(acc[j][c])) << (8 * j);
          tmp[(static_cast<size_t>(c) * p.S + x) * p.tw + g] = word;
        }
      }
    }
    // (the next group's staging only writes `stage`; its de-interleave starts behind the barrier that follows the staging)
  }
  __syncthreads();
  // ---- vertical pass out of the transposed window; planar output, coalesced along x
  if (active) {
    const size_t plane = static_cast<size_t>(p.S) * p.S;
    for (int yy = 0; yy < ny; ++yy) {
      const int y = y0 + yy;
      const int wi = (p.vw0[y] - r0a) >> 2;
      const uint32_t* vwy = p.vwt + static_cast<size_t>(y) * 3 * p.nwv;
      uint8_t* d = p.dst + (static_cast<size_t>(b) * 3 * p.S + y) * p.S + x;
      if constexpr (DP4A) {
        int a0[3] = {0, 0, 0}, a1[3] = {0, 0, 0}, a2[3] = {0, 0, 0};
        for (int n = 0; n < p.nwv; ++n) {
          const uint32_t b0 = __ldg(vwy + n), b1 = __ldg(vwy + p.nwv + n), b2 = __ldg(vwy + 2 * p.nwv + n);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const uint32_t px = tmp[(static_cast<size_t>(c) * p.S + x) * p.tw + wi + n];
            a0[c] = dp4a_uu(px, b0, a0[c]);
            a1[c] = dp4a_uu(px, b1, a1[c]);
            a2[c] = dp4a_us(px, b2, a2[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
          d[c * plane] = static_cast<uint8_t>(clip8((1 << (PRECISION_BITS - 1)) + a0[c] + (a1[c] << 8) + (a2[c] << 16)));
      } else {
        int acc[3] = {1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1), 1 << (PRECISION_BITS - 1)};
        const int4* wy = reinterpret_cast<const int4*>(p.vwi + static_cast<size_t>(y) * 4 * p.nwv);
        for (int n = 0; n < p.nwv; ++n) {
          const int4 w = __ldg(wy + n);
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[c] = mac4(tmp[(static_cast<size_t>(c) * p.S + x) * p.tw + wi + n], w, acc[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) d[c * plane] = static_cast<uint8_t>(clip8(acc[c]));
      }
    }
  }
}

// v2 tables: taps regrouped into aligned words, weights as byte planes.  false when a weight does not fit 3 byte planes.
bool build_v2_axis(const std::vector<int>& bounds, const std::vector<int>& coeffs, int ksize, int count, bool align_absolute,
                   int* nw_out, std::vector<int>& first_out, std::vector<uint32_t>& wt) {
  int nw = 1;
  for (int i = 0; i < count; ++i) nw = std::max(nw, ((bounds[2 * i] & 3) + bounds[2 * i + 1] + 3) / 4);
  first_out.resize(count);
  wt.assign(static_cast<size_t>(count) * 3 * nw, 0u);
  for (int i = 0; i < count; ++i) {
    const int first = bounds[2 * i], taps = bounds[2 * i + 1], mis = first & 3;
    first_out[i] = align_absolute ? (first & ~3) : (first >> 2);
    for (int k = 0; k < taps; ++k) {
      const int w = coeffs[static_cast<size_t>(i) * ksize + k];
      const int b2 = w >> 16;   // arithmetic shift: w = (w & 0xffff) + 65536 * b2
      if (b2 < -128 || b2 > 127) return false;
      const uint32_t planes3[3] = {static_cast<uint32_t>(w) & 0xffu, (static_cast<uint32_t>(w) >> 8) & 0xffu, static_cast<uint32_t>(b2) & 0xffu};
      const int byte = mis + k;
      for (int pl = 0; pl < 3; ++pl) wt[(static_cast<size_t>(i) * 3 + pl) * nw + byte / 4] |= planes3[pl] << (8 * (byte & 3));
    }
  }
  *nw_out = nw;
  return true;
}

void build_v3_axis(const std::vector<int>& bounds, const std::vector<int>& coeffs, int ksize, int count, int nw, std::vector<int>& wi) {
  wi.assign(static_cast<size_t>(count) * 4 * nw, 0);
  for (int i = 0; i < count; ++i) {
    const int mis = bounds[2 * i] & 3;
    for (int k = 0; k < bounds[2 * i + 1]; ++k) wi[static_cast<size_t>(i) * 4 * nw + mis + k] = coeffs[static_cast<size_t>(i) * ksize + k];
  }
}

void build_v2(ResizePlanHost* plan) {
  const int S = plan->S;
  plan->v2 = false;
  if (!build_v2_axis(plan->hb, plan->hk, plan->kh, S, false, &plan->nwh, plan->hw0, plan->hwt)) return;
  if (!build_v2_axis(plan->vb, plan->vk, plan->kv, S, true, &plan->nwv, plan->vw0, plan->vwt)) return;
  build_v3_axis(plan->hb, plan->hk, plan->kh, S, plan->nwh, plan->hwi);
  build_v3_axis(plan->vb, plan->vk, plan->kv, S, plan->nwv, plan->vwi);
  const int span_px = plan->span_bytes / 3;
  plan->ng4 = (span_px + 3) / 4;
  plan->pw = plan->ng4 + plan->nwh;                                  // reads run up to nwh words past a column's first word
  plan->row_pitch2 = (plan->ng4 * 12 + 8 + 15) / 16 * 16;            // the de-interleave reads words 3t .. 3t+3 behind a <= 3-byte offset
  for (int ty : {16, 8, 4, 2, 1}) {
    int groups = 0;
    for (int y0 = 0; y0 < S; y0 += ty) {
      int r_end = 0;
      for (int y = y0; y < std::min(S, y0 + ty); ++y) r_end = std::max(r_end, plan->vw0[y] + 4 * plan->nwv);
      groups = std::max(groups, (r_end - plan->vw0[y0]) / 4);
    }
    const int tw = groups | 1;                                       // odd word stride between columns: conflict-free
    const long long smem = static_cast<long long>(G) * plan->row_pitch2 + 3LL * G * plan->pw * 4 + 3LL * S * tw * 4;
    if (smem <= 72 * 1024 || (ty == 1 && smem <= 200 * 1024)) {
      plan->ty2 = ty; plan->tw = tw; plan->smem2 = static_cast<int>(smem); plan->v2 = true;
      return;
    }
  }
}

int g_resize_version = 1;   // v2 is bit-identical but measured 20 % slower: kept as a cross-check

}  // namespace

void resize_set_version(int v) { g_resize_version = (v >= 1 && v <= 3) ? v : 1; }

void resized_output_size(int H, int W, int S, int* nh, int* nw) {
  // torchvision _compute_resized_output_size(size=[S]): shorter edge -> S, longer -> int(S * long / short)
  const int shortv = (W <= H) ? W : H, longv = (W <= H) ? H : W;
  const int new_long = static_cast<int>(static_cast<double>(S) * longv / shortv);
  if (W <= H) { *nw = S; *nh = new_long; } else { *nh = S; *nw = new_long; }
}

static int round_half_even(double v) { return static_cast<int>(std::nearbyint(v)); }  // Python round() (default FE_TONEAREST)

int resize_plan_build(ResizePlanHost* plan, int H, int W, int S) {
  if (H <= 0 || W <= 0 || S <= 0 || S > 256) return -3;
  plan->H = H; plan->W = W; plan->S = S;
  resized_output_size(H, W, S, &plan->nh, &plan->nw);
  if (plan->nh < S || plan->nw < S) return -3;
  plan->top = round_half_even((plan->nh - S) / 2.0);   // torchvision center_crop
  plan->left = round_half_even((plan->nw - S) / 2.0);
  axis_tables(W, plan->nw, plan->left, S, plan->hb, plan->hk, plan->kh);
  axis_tables(H, plan->nh, plan->top, S, plan->vb, plan->vk, plan->kv);
  // column span the crop needs; horizontal bounds become relative to it
  plan->x0 = plan->hb[0];
  int x1 = 0;
  for (int i = 0; i < S; ++i) x1 = std::max(x1, plan->hb[2 * i] + plan->hb[2 * i + 1]);
  for (int i = 0; i < S; ++i) plan->hb[2 * i] -= plan->x0;
  plan->span_bytes = (x1 - plan->x0) * 3;
  plan->row_pitch = (plan->span_bytes + 3 + 15) / 16 * 16;
  plan->tmp_pitch = (3 * S + 15) / 16 * 16;
  // rows per CTA: the largest tile whose source-row window fits in shared memory
  const int smem_cap = 200 * 1024;
  plan->ty = 0;
  for (int ty : {8, 4, 2, 1}) {
    int max_rows = 0;
    for (int y0 = 0; y0 < S; y0 += ty) {
      const int y1 = std::min(S, y0 + ty) - 1;
      max_rows = std::max(max_rows, plan->vb[2 * y1] + plan->vb[2 * y1 + 1] - plan->vb[2 * y0]);
    }
    const long long smem = static_cast<long long>(G) * plan->row_pitch + static_cast<long long>(max_rows) * plan->tmp_pitch;
    if (smem <= smem_cap) { plan->ty = ty; plan->max_rows = max_rows; plan->smem_bytes = static_cast<int>(smem); break; }
  }
  if (plan->ty == 0) return -6;  // source too large for the shared-memory window
  build_v2(plan);
  return 0;
}

static size_t v3_offset(const ResizePlanHost& plan);

int resize_crop_launch(const ResizePlanHost& plan, const int* d_tables, const uint8_t* src, uint8_t* dst, long long B, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(resize_crop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(resize_crop2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(resize_crop2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  if (plan.v2 && g_resize_version >= 2) {
    const int S = plan.S;
    KParams2 q;
    const int* t2 = d_tables + resize_plan_table_ints(plan);   // the v2 tables follow the v1 tables
    q.hw0 = t2;
    q.hwt = reinterpret_cast<const uint32_t*>(q.hw0 + S);
    q.vw0 = reinterpret_cast<const int*>(q.hwt + static_cast<size_t>(S) * 3 * plan.nwh);
    q.vwt = reinterpret_cast<const uint32_t*>(q.vw0 + S);
    q.hwi = d_tables + v3_offset(plan);
    q.vwi = q.hwi + static_cast<size_t>(S) * 4 * plan.nwh;
    q.H = plan.H; q.W = plan.W; q.S = S; q.x0 = plan.x0; q.span_bytes = plan.span_bytes; q.row_pitch = plan.row_pitch2;
    q.ng4 = plan.ng4; q.pw = plan.pw; q.tw = plan.tw; q.nwh = plan.nwh; q.nwv = plan.nwv; q.ty = plan.ty2;
    const int threads = (S + 31) / 32 * 32;
    for (long long b0 = 0; b0 < B; b0 += 65535) {
      const int nb = static_cast<int>(std::min<long long>(65535, B - b0));
      q.src = src + static_cast<size_t>(b0) * plan.H * plan.W * 3;
      q.dst = dst + static_cast<size_t>(b0) * 3 * S * S;
      dim3 grid(static_cast<unsigned>((S + plan.ty2 - 1) / plan.ty2), static_cast<unsigned>(nb));
      if (g_resize_version == 2) resize_crop2_kernel<true><<<grid, threads, plan.smem2, s>>>(q);
      else resize_crop2_kernel<false><<<grid, threads, plan.smem2, s>>>(q);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return static_cast<int>(e);
    }
    return 0;
  }
  KParams p;
  p.src = src; p.dst = dst;
  const int S = plan.S;
  p.hb = d_tables;
  p.hk = p.hb + 2 * S;
  p.vb = p.hk + static_cast<size_t>(S) * plan.kh;
  p.vk = p.vb + 2 * S;
  p.H = plan.H; p.W = plan.W; p.S = S; p.kh = plan.kh; p.kv = plan.kv; p.x0 = plan.x0; p.span_bytes = plan.span_bytes;
  p.row_pitch = plan.row_pitch; p.ty = plan.ty; p.max_rows = plan.max_rows; p.tmp_pitch = plan.tmp_pitch;
  const int threads = (S + 31) / 32 * 32;
  for (long long b0 = 0; b0 < B; b0 += 65535) {
    const int nb = static_cast<int>(std::min<long long>(65535, B - b0));
    p.src = src + static_cast<size_t>(b0) * plan.H * plan.W * 3;
    p.dst = dst + static_cast<size_t>(b0) * 3 * S * S;
    dim3 grid(static_cast<unsigned>((S + plan.ty - 1) / plan.ty), static_cast<unsigned>(nb));
    resize_crop_kernel<<<grid, threads, plan.smem_bytes, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  return 0;
}

size_t resize_plan_table_ints(const ResizePlanHost& plan) {
  return static_cast<size_t>(plan.S) * (4 + plan.kh + plan.kv);
}

// ints from the start of the whole table (v1 | v2 | pad | v3) to the v3 weights: a multiple of 4, so that the int4 loads are aligned
static size_t v3_offset(const ResizePlanHost& plan) {
  const size_t end_v2 = resize_plan_table_ints(plan) + static_cast<size_t>(plan.S) * (2 + 3 * plan.nwh + 3 * plan.nwv);
  return (end_v2 + 3) / 4 * 4;
}

size_t resize_plan_table_ints_v2(const ResizePlanHost& plan) {
  if (!plan.v2) return 0;
  return v3_offset(plan) + static_cast<size_t>(plan.S) * 4 * (plan.nwh + plan.nwv) - resize_plan_table_ints(plan);
}

void resize_plan_pack_v2(const ResizePlanHost& plan, int* out) {
  if (!plan.v2) return;
  const int S = plan.S;
  std::memcpy(out, plan.hw0.data(), sizeof(int) * S);
  out += S;
  std::memcpy(out, plan.hwt.data(), sizeof(uint32_t) * plan.hwt.size());
  out += plan.hwt.size();
  std::memcpy(out, plan.vw0.data(), sizeof(int) * S);
  out += S;
  std::memcpy(out, plan.vwt.data(), sizeof(uint32_t) * plan.vwt.size());
  out += plan.vwt.size();
  out += v3_offset(plan) - (resize_plan_table_ints(plan) + static_cast<size_t>(S) * (2 + 3 * plan.nwh + 3 * plan.nwv));   // alignment pad
  std::memcpy(out, plan.hwi.data(), sizeof(int) * plan.hwi.size());
  out += plan.hwi.size();
  std::memcpy(out, plan.vwi.data(), sizeof(int) * plan.vwi.size());
}

void resize_plan_pack(const ResizePlanHost& plan, int* out) {
  const int S = plan.S;
  std::memcpy(out, plan.hb.data(), sizeof(int) * 2 * S);
  out += 2 * S;
  std::memcpy(out, plan.hk.data(), sizeof(int) * static_cast<size_t>(S) * plan.kh);
  out += static_cast<size_t>(S) * plan.kh;
  std::memcpy(out, plan.vb.data(), sizeof(int) * 2 * S);
  out += 2 * S;
  std::memcpy(out, plan.vk.data(), sizeof(int) * static_cast<size_t>(S) * plan.kv);
}

}  // namespace hb
