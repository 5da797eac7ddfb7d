// attn3_timing.cu — per-phase clock64 stamps of hb::vit_attn3_kernel (build with -DHB_ATTN_TIMING, tools/build_tools.sh).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hirest_b200/csrc/hb_attn.cuh"
#include "../hirest_b200/csrc/hb_gemm.cuh"
int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 256, H = 16, D = H * 88;
  if (hb::tmap_init() != 0) { printf("tmap_init failed\n"); return 1; }
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  size_t n = (size_t)B * 257 * 3 * D;
  std::vector<__nv_bfloat16> h(n);
  unsigned s = 1;
  for (auto& v : h) { s = s * 1664525u + 1013904223u; v = __float2bfloat16((((s >> 8) & 0xFFFF) / 65536.0f - 0.5f)); }
  __nv_bfloat16 *qkv, *out; long long* tim;
  const size_t tn = (size_t)sms * 64 * 4 * 16;
  cudaMalloc(&qkv, n * 2); cudaMalloc(&out, (size_t)B * 257 * D * 2); cudaMalloc(&tim, tn * 8);
  cudaMemset(tim, 0, tn * 8);
  cudaMemcpy(qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
  hb::AttnParams p; p.qkv = qkv; p.out = out; p.B = B; p.H = H; p.timing = tim;
  for (int i = 0; i < 3; ++i) hb::vit_attn3_launch(p, sms, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> t(tn);
  cudaMemcpy(t.data(), tim, tn * 8, cudaMemcpyDeviceToHost);
  const int items_per_cta = (B * H) / sms;
  const char* sn[12] = {"start (prev: extra row)", "e softmax (half 1) / s_x", "wait S", "pass 1 (max)", "pair barrier (+store wait)",
                        "turn wait + pass 2 (exp, P)", "next item's dots (wait QK_FULL)", "wait O", "output + stage + TMA store", "", "", ""};
  const char* mn[11] = {"start", "wait QK_FULL", "wait TF0", "issue S0", "wait TF1", "issue S1", "wait V_FULL + ones", "wait P0/PX", "issue PV0+x", "wait P1", "issue PV1"};
  const char* pn[4] = {"start", "x loads + wait QK_FREE", "issue QK", "wait V_FREE"};
  for (int role = 0; role < 4; ++role) {
    const int ns = role < 2 ? 9 : (role == 2 ? 11 : 4);
    std::vector<double> acc(16, 0.0);
    double period = 0; int cnt = 0, pc = 0;
    for (int c = 0; c < sms; ++c)
      for (int it = 2; it < items_per_cta && it < 64; ++it) {   // skip the pipeline fill
        const long long* x = &t[(((size_t)c * 64 + it) * 4 + role) * 16];
        const long long* xp = &t[(((size_t)c * 64 + it - 1) * 4 + role) * 16];
        if (x[0] == 0 || xp[0] == 0) continue;
        for (int k = 1; k < ns; ++k) acc[k] += double(x[k] - x[k - 1]);
        period += double(x[0] - xp[0]); ++pc; ++cnt;
      }
    printf("---- role %d (%s), %d samples, period %.0f cycles / item\n", role, role == 0 ? "softmax tile 0" : role == 1 ? "softmax tile 1" : role == 2 ? "MMA warp" : "producer", cnt, pc ? period / pc : 0.0);
    for (int k = 1; k < ns; ++k) printf("  %-34s %8.0f\n", role < 2 ? sn[k] : (role == 2 ? mn[k] : pn[k]), cnt ? acc[k] / cnt : 0.0);
  }
  return 0;
}
