// attn_timing.cu — per-phase clock64 stamps of the ViT attention kernel (debug build with -DHB_ATTN_TIMING).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../hirest_b200/csrc/hb_attn.cuh"
int main() {
  const int B = 64, H = 16, D = H * 88;
  size_t n = (size_t)B * 257 * 3 * D;
  std::vector<__nv_bfloat16> h(n);
  unsigned s = 1;
  for (auto& v : h) { s = s * 1664525u + 1013904223u; v = __float2bfloat16((((s >> 8) & 0xFFFF) / 65536.0f - 0.5f)); }
  __nv_bfloat16 *qkv, *out; long long* tim;
  cudaMalloc(&qkv, n * 2); cudaMalloc(&out, (size_t)B * 257 * D * 2); cudaMalloc(&tim, (size_t)B * H * 16 * 8);
  cudaMemcpy(qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
  hb::AttnParams p; p.qkv = qkv; p.out = out; p.B = B; p.H = H; p.timing = tim;
  for (int i = 0; i < 3; ++i) hb::vit_attn_launch(p, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> t((size_t)B * H * 16);
  cudaMemcpy(t.data(), tim, t.size() * 8, cudaMemcpyDeviceToHost);
  double acc[12] = {0};
  int cnt = 0;
  for (int c = 148; c < B * H; ++c) {  // skip the first wave
    for (int i = 1; i < 12; ++i) acc[i] += double(t[c * 16 + i] - t[c * 16 + i - 1]);
    ++cnt;
  }
  const char* names[13] = {"", "setup(alloc,bar)", "issue loads", "cp.async wait", "bar+kx", "s_x,e_t,stats", "wait S + bar", "rowmax pass", "exp pass + P write", "stepB (extra q PV)", "wait O", "output store", "final sync+dealloc"};
  double tot = 0;
  for (int i = 1; i < 12; ++i) { printf("%-22s %8.0f cycles\n", names[i], acc[i] / cnt); tot += acc[i] / cnt; }
  printf("total %8.0f cycles per CTA (warp 0 thread 0 view)\n", tot);
  return 0;
}
