// Host-side check of the resize kernels v2 / v3 tables and arithmetic (no GPU needed): emulates resize_crop2_kernel's data flow on the
// CPU — aligned-word staging, the byte_perm de-interleave, byte-plane weights + dp4a, the transposed 4-rows-per-word window —
// from the plan's v2 tables, and compares every output byte with the plain two-pass evaluation of the v1 tables (Pillow's
// arithmetic).   nvcc -std=c++17 -o tools/_build/preproc_host_check tools/preproc_host_check.cu build/hb_preproc.o
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../hirest_b200/csrc/hb_preproc.cuh"

static uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t v = (static_cast<uint64_t>(y) << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= static_cast<uint32_t>((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
static uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  const uint64_t v = (static_cast<uint64_t>(hi) << 32) | lo;
  return static_cast<uint32_t>(v >> (sh & 31));
}
static int dp4a_uu(uint32_t a, uint32_t b, int c) {
  uint32_t acc = static_cast<uint32_t>(c);
  for (int i = 0; i < 4; ++i) acc += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
  return static_cast<int>(acc);
}
static int dp4a_us(uint32_t a, uint32_t b, int c) {
  uint32_t acc = static_cast<uint32_t>(c);
  for (int i = 0; i < 4; ++i) acc += static_cast<uint32_t>(static_cast<int>((a >> (8 * i)) & 0xff) * static_cast<int>(static_cast<int8_t>((b >> (8 * i)) & 0xff)));
  return static_cast<int>(acc);
}
static int clip8(int v) { v >>= 22; return std::min(std::max(v, 0), 255); }

static int check(int H, int W, int S) {
  hb::ResizePlanHost p;
  if (int r = hb::resize_plan_build(&p, H, W, S)) { std::printf("%dx%d: plan failed %d\n", H, W, r); return 1; }
  if (!p.v2) { std::printf("%dx%d: no v2 form\n", H, W); return 1; }
  std::vector<uint8_t> src(static_cast<size_t>(H) * W * 3 + 16);
  for (auto& b : src) b = static_cast<uint8_t>(std::rand());
  // reference: v1 tables, two passes
  std::vector<uint8_t> ref(static_cast<size_t>(3) * S * S), got(ref.size(), 0);
  {
    const int r0 = p.vb[0], r1 = p.vb[2 * (S - 1)] + p.vb[2 * (S - 1) + 1];
    std::vector<uint8_t> tmp(static_cast<size_t>(r1 - r0) * S * 3);
    for (int r = r0; r < r1; ++r)
      for (int x = 0; x < S; ++x)
        for (int c = 0; c < 3; ++c) {
          int acc = 1 << 21;
          for (int k = 0; k < p.hb[2 * x + 1]; ++k) acc += src[(static_cast<size_t>(r) * W + p.x0 + p.hb[2 * x] + k) * 3 + c] * p.hk[static_cast<size_t>(x) * p.kh + k];
          tmp[(static_cast<size_t>(r - r0) * S + x) * 3 + c] = static_cast<uint8_t>(clip8(acc));
        }
    for (int y = 0; y < S; ++y)
      for (int x = 0; x < S; ++x)
        for (int c = 0; c < 3; ++c) {
          int acc = 1 << 21;
          for (int k = 0; k < p.vb[2 * y + 1]; ++k) acc += tmp[(static_cast<size_t>(p.vb[2 * y] + k - r0) * S + x) * 3 + c] * p.vk[static_cast<size_t>(y) * p.kv + k];
          ref[(static_cast<size_t>(c) * S + y) * S + x] = static_cast<uint8_t>(clip8(acc));
        }
  }
  // emulation of resize_crop2_kernel, one "CTA" per row tile
  const int G = 4;
  for (int y0 = 0; y0 < S; y0 += p.ty2) {
    const int ny = std::min(p.ty2, S - y0);
    std::vector<uint8_t> stage(static_cast<size_t>(G) * p.row_pitch2, 0xAB);
    std::vector<uint32_t> planes(static_cast<size_t>(3) * G * p.pw, 0xDEADBEEF), tmp(static_cast<size_t>(3) * S * p.tw, 0xDEADBEEF);
    const int r0a = p.vw0[y0];
    int r_end = 0;
    for (int yy = 0; yy < ny; ++yy) r_end = std::max(r_end, p.vw0[y0 + yy] + 4 * p.nwv);
    r_end = std::min(r_end, (H + 3) & ~3);
    const int n_groups = (r_end - r0a) >> 2;
    if (n_groups > p.tw) { std::printf("%dx%d: window %d groups > tw %d\n", H, W, n_groups, p.tw); return 1; }
    for (int g = 0; g < n_groups; ++g) {
      const int r = r0a + 4 * g;
      for (int j = 0; j < G; ++j) {
        const int row = std::min(r + j, H - 1);
        const size_t off = (static_cast<size_t>(row) * W + p.x0) * 3;
        const int mis = static_cast<int>(off & 3);   // the frame base is 4-byte aligned here (device: address & 3)
        for (int i = 0; i < p.span_bytes; ++i) stage[static_cast<size_t>(j) * p.row_pitch2 + mis + i] = src[off + i];
        const uint32_t* srow = reinterpret_cast<const uint32_t*>(stage.data() + static_cast<size_t>(j) * p.row_pitch2);
        for (int t = 0; t < p.ng4; ++t) {
          if ((3 * t + 3) * 4 + 4 > p.row_pitch2) { std::printf("stage overrun\n"); return 1; }
          const uint32_t w0 = srow[3 * t], w1 = srow[3 * t + 1], w2 = srow[3 * t + 2], w3 = srow[3 * t + 3];
          const uint32_t sh = mis * 8;
          const uint32_t a0 = funnel_r(w0, w1, sh), a1 = funnel_r(w1, w2, sh), a2 = funnel_r(w2, w3, sh);
          planes[(0 * G + j) * p.pw + t] = byte_perm(byte_perm(a0, a1, 0x0630), a2, 0x5210);
          planes[(1 * G + j) * p.pw + t] = byte_perm(byte_perm(a0, a1, 0x0741), a2, 0x6210);
          planes[(2 * G + j) * p.pw + t] = byte_perm(byte_perm(a0, a1, 0x0052), a2, 0x7410);
        }
      }
      for (int x = 0; x < S; ++x)
        for (int c = 0; c < 3; ++c) {
          uint32_t word = 0;
          for (int j = 0; j < G; ++j) {
            int a0 = 0, a1 = 0, a2 = 0;
            for (int n = 0; n < p.nwh; ++n) {
              if (p.hw0[x] + n >= p.pw) { std::printf("plane overrun\n"); return 1; }
              const uint32_t px = planes[(c * G + j) * p.pw + p.hw0[x] + n];
              a0 = dp4a_uu(px, p.hwt[(static_cast<size_t>(x) * 3 + 0) * p.nwh + n], a0);
              a1 = dp4a_uu(px, p.hwt[(static_cast<size_t>(x) * 3 + 1) * p.nwh + n], a1);
              a2 = dp4a_us(px, p.hwt[(static_cast<size_t>(x) * 3 + 2) * p.nwh + n], a2);
            }
            const int v = static_cast<int>(static_cast<uint32_t>(1 << 21) + static_cast<uint32_t>(a0) + (static_cast<uint32_t>(a1) << 8) + (static_cast<uint32_t>(a2) << 16));
            uint32_t v3 = 1u << 21;   // kernel v3: the same words, plain weights per byte
            for (int n = 0; n < p.nwh; ++n) {
              const uint32_t px = planes[(c * G + j) * p.pw + p.hw0[x] + n];
              for (int i = 0; i < 4; ++i) v3 += ((px >> (8 * i)) & 0xff) * static_cast<uint32_t>(p.hwi[(static_cast<size_t>(x) * p.nwh + n) * 4 + i]);
            }
            if (static_cast<int>(v3) != v) { std::printf("%dx%d: v3 horizontal sum differs at x=%d\n", H, W, x); return 1; }
            word |= static_cast<uint32_t>(clip8(v)) << (8 * j);
          }
          tmp[(static_cast<size_t>(c) * S + x) * p.tw + g] = word;
        }
    }
    for (int yy = 0; yy < ny; ++yy) {
      const int y = y0 + yy, wi = (p.vw0[y] - r0a) >> 2;
      for (int x = 0; x < S; ++x)
        for (int c = 0; c < 3; ++c) {
          int a0 = 0, a1 = 0, a2 = 0;
          for (int n = 0; n < p.nwv; ++n) {
            if (wi + n >= p.tw) { std::printf("window overrun\n"); return 1; }
            const uint32_t px = tmp[(static_cast<size_t>(c) * S + x) * p.tw + wi + n];
            a0 = dp4a_uu(px, p.vwt[(static_cast<size_t>(y) * 3 + 0) * p.nwv + n], a0);
            a1 = dp4a_uu(px, p.vwt[(static_cast<size_t>(y) * 3 + 1) * p.nwv + n], a1);
            a2 = dp4a_us(px, p.vwt[(static_cast<size_t>(y) * 3 + 2) * p.nwv + n], a2);
          }
          const int v = static_cast<int>(static_cast<uint32_t>(1 << 21) + static_cast<uint32_t>(a0) + (static_cast<uint32_t>(a1) << 8) + (static_cast<uint32_t>(a2) << 16));
          uint32_t v3 = 1u << 21;
          for (int n = 0; n < p.nwv; ++n) {
            const uint32_t px = tmp[(static_cast<size_t>(c) * S + x) * p.tw + wi + n];
            for (int i = 0; i < 4; ++i) v3 += ((px >> (8 * i)) & 0xff) * static_cast<uint32_t>(p.vwi[(static_cast<size_t>(y) * p.nwv + n) * 4 + i]);
          }
          if (static_cast<int>(v3) != v) { std::printf("%dx%d: v3 vertical sum differs at y=%d\n", H, W, y); return 1; }
          got[(static_cast<size_t>(c) * S + y) * S + x] = static_cast<uint8_t>(clip8(v));
        }
    }
  }
  size_t bad = 0;
  for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != got[i];
  std::printf("%4dx%-4d -> %d: taps %d/%d, words %d/%d, rows per tile %d, window %d words, smem %d B: %s (%zu differing bytes)\n", H, W, S, p.kh, p.kv,
              p.nwh, p.nwv, p.ty2, p.tw, p.smem2, bad ? "MISMATCH" : "identical", bad);
  return bad ? 1 : 0;
}

int main() {
  int rc = 0;
  const int sizes[][2] = {{360, 640}, {720, 1280}, {224, 224}, {225, 300}, {341, 256}, {1080, 1920}, {480, 854}, {240, 426}, {256, 256}, {299, 299}, {2160, 3840}};
  for (auto& s : sizes) rc |= check(s[0], s[1], 224);
  rc |= check(360, 640, 112);
  rc |= check(97, 131, 64);
  return rc;
}
