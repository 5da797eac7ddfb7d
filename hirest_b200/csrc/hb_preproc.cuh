// hb_preproc.cuh — GPU frame preprocessing: Resize(S, BICUBIC) + CenterCrop(S) on uint8 RGB frames, bit-identical to
// torchvision-on-PIL (implementation and reference citations in hb_preproc.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace hb {

// Host-side plan for one source size: torchvision's output size / crop offsets and Pillow's fixed-point weight tables
// restricted to the S x S crop.
struct ResizePlanHost {
  int H = 0, W = 0, S = 0;
  int nh = 0, nw = 0, top = 0, left = 0;  // resized size and crop offsets
  int kh = 0, kv = 0;                     // taps per output column / row (table row strides)
  int x0 = 0, span_bytes = 0;             // first source column the crop needs; bytes per source row it needs
  int row_pitch = 0, tmp_pitch = 0, ty = 0, max_rows = 0, smem_bytes = 0;
  std::vector<int> hb, hk, vb, vk;        // bounds [S,2] = (first tap, taps), weights [S,k] (22-bit fixed point)
  // kernel v2 (word loads + dp4a): taps regrouped into 4-byte words aligned to the source, weights as three byte planes
  bool v2 = false;
  int nwh = 0, nwv = 0;                   // words of taps per output column / row (incl. the alignment slack)
  int ng4 = 0, pw = 0, tw = 0;            // 4-pixel groups per staged row; words per planar row; words per (channel, column) of the window
  int ty2 = 0, row_pitch2 = 0, smem2 = 0;
  std::vector<int> hw0, vw0;              // [S] first word of a column's taps in the planar row / first source row (multiple of 4) of a row's taps
  std::vector<uint32_t> hwt, vwt;         // [S][3][nwh] / [S][3][nwv]: weight bytes b0 | b1 | b2 (w = b0 + 256 b1 + 65536 b2, b2 signed)
  std::vector<int> hwi, vwi;              // [S][4 nwh] / [S][4 nwv]: the same taps as plain weights, one per byte of the words (kernel v3)
};

void resized_output_size(int H, int W, int S, int* nh, int* nw);
int resize_plan_build(ResizePlanHost* plan, int H, int W, int S);   // 0, -3 invalid size, -6 source too large
size_t resize_plan_table_ints(const ResizePlanHost& plan);
void resize_plan_pack(const ResizePlanHost& plan, int* out);        // hb | hk | vb | vk, as the kernel expects them
size_t resize_plan_table_ints_v2(const ResizePlanHost& plan);       // 0 when the plan has no v2 form
void resize_plan_pack_v2(const ResizePlanHost& plan, int* out);     // hw0 | hwt | vw0 | vwt | pad | hwi | vwi; stored right behind the v1 tables
void resize_set_version(int v);                                     // 1: byte loads + IMAD; 2: word loads + dp4a; 3: word loads + byte extraction + IMAD
// src [B,H,W,3] uint8 (device) -> dst [B,3,S,S] uint8 (device). d_tables = packed tables on the device.
int resize_crop_launch(const ResizePlanHost& plan, const int* d_tables, const uint8_t* src, uint8_t* dst, long long B, cudaStream_t s);

}  // namespace hb
