// gemm_selftest.cu — standalone correctness + timing probe for the tcgen05 GEMM (no torch needed).
// Build: see tools/build_tools.sh.  Run on a B200: ./tools/_build/gemm_selftest [quick]
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>

#include "../hirest_b200/csrc/hb_gemm.cuh"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

static uint32_t g_seed = 12345u;
static int* g_sched = nullptr;   // dynamic tile scheduler counters (argv[8] = 1)
static float frand() {
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

struct Case {
  const char* name;
  int M, N, K, epi, cg;
  bool resid, qscale;
};

static int run_case(const Case& c, int num_sms, bool verify, int iters) {
  const int M = c.M, N = c.N, K = c.K;
  const size_t na = (size_t)M * K, nw = (size_t)N * K;
  std::vector<__nv_bfloat16> hA(na), hW(nw);
  std::vector<float> hbias(N), hres;
  for (size_t i = 0; i < na; ++i) hA[i] = __float2bfloat16(frand());
  for (size_t i = 0; i < nw; ++i) hW[i] = __float2bfloat16(frand() * 0.1f);
  for (int i = 0; i < N; ++i) hbias[i] = frand();
  __nv_bfloat16 *dA, *dW;
  float* dbias;
  void* dout;
  float* dres = nullptr;
  const size_t out_bytes = (size_t)M * N * (c.epi == hb::EPI_F32 ? 4 : 2);
  CK(cudaMalloc(&dA, na * 2));
  CK(cudaMalloc(&dW, nw * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMalloc(&dout, out_bytes));
  CK(cudaMemcpy(dA, hA.data(), na * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), nw * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, out_bytes));
  if (c.resid) {
    hres.resize((size_t)M * N);
    for (auto& v : hres) v = frand();
    CK(cudaMalloc(&dres, (size_t)M * N * 4));
    CK(cudaMemcpy(dres, hres.data(), (size_t)M * N * 4, cudaMemcpyHostToDevice));
  }
  CUtensorMap tmA, tmW;
  int r = hb::make_tmap_bf16(&tmA, dA, M, K, K, hb::gemm_a_box_rows());
  if (r) { printf("tmap A failed %d\n", r); return 1; }
  r = hb::make_tmap_bf16(&tmW, dW, N, K, K, hb::gemm_w_box_rows(c.cg));
  if (r) { printf("tmap W failed %d\n", r); return 1; }
  hb::GemmParams p;
  p.M = M; p.N = N; p.K = K; p.bias = dbias; p.out = dout; p.ldo = N;
  p.resid = dres;
  p.sched = g_sched;
  if (c.qscale) { p.qscale = 0.25f; p.qcols = N / 3; }
  r = hb::gemm_launch(tmA, tmW, p, c.epi, c.cg, num_sms, 0);
  if (r) { printf("%s: launch failed %d\n", c.name, r); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", c.name, cudaGetErrorString(e)); return 2; }

  double max_err = 0, max_ref = 0;
  long bad = 0, checked = 0;
  if (verify) {
    std::vector<uint8_t> hout(out_bytes);
    CK(cudaMemcpy(hout.data(), dout, out_bytes, cudaMemcpyDeviceToHost));
    std::vector<int> rows;
    for (int i = 0; i < M; i += 61) rows.push_back(i);
    for (int i = std::max(0, M - 3); i < M; ++i) rows.push_back(i);
    for (int i : {127, 128, 129, 255, 256, 257}) if (i < M) rows.push_back(i);
    for (int row : rows) {
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        const __nv_bfloat16* a = &hA[(size_t)row * K];
        const __nv_bfloat16* w = &hW[(size_t)n * K];
        for (int k = 0; k < K; ++k) acc += (double)__bfloat162float(a[k]) * (double)__bfloat162float(w[k]);
        acc += hbias[n];
        double ref, got;
        if (c.epi == hb::EPI_F32) {
          if (c.resid) acc += hres[(size_t)row * N + n];
          ref = acc;
          got = reinterpret_cast<float*>(hout.data())[(size_t)row * N + n];
        } else {
          if (c.epi == hb::EPI_GELU_BF16) acc = 0.5 * acc * (1.0 + erf(acc * 0.7071067811865476));
          if (c.qscale && n < N / 3) acc *= 0.25;
          ref = acc;
          got = __bfloat162float(reinterpret_cast<__nv_bfloat16*>(hout.data())[(size_t)row * N + n]);
        }
        const double err = fabs(got - ref);
        const double tol = (c.epi == hb::EPI_F32 ? 2e-3 : 1e-2) * (1.0 + fabs(ref));
        if (!(err <= tol)) {
          if (bad < 5) printf("  mismatch row %d col %d got %f ref %f\n", row, n, got, ref);
          ++bad;
        }
        if (err > max_err) max_err = err;
        if (fabs(ref) > max_ref) max_ref = fabs(ref);
        ++checked;
      }
    }
  }
  float ms = 0;
  if (iters > 0) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) hb::gemm_launch(tmA, tmW, p, c.epi, c.cg, num_sms, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) hb::gemm_launch(tmA, tmW, p, c.epi, c.cg, num_sms, 0);
    cudaEventRecord(e1);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: timing run failed: %s\n", c.name, cudaGetErrorString(e)); return 2; }
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
  }
  const double tflops = ms > 0 ? 2.0 * M * N * (double)K / (ms * 1e-3) / 1e12 : 0;
  printf("{\"case\":\"%s\",\"M\":%d,\"N\":%d,\"K\":%d,\"epi\":%d,\"cg\":%d,\"checked\":%ld,\"bad\":%ld,\"max_err\":%.4g,"
         "\"max_ref\":%.4g,\"ms\":%.4f,\"tflops\":%.1f}\n",
         c.name, M, N, K, c.epi, c.cg, checked, bad, max_err, max_ref, ms, tflops);
  fflush(stdout);
  cudaFree(dA); cudaFree(dW); cudaFree(dbias); cudaFree(dout);
  if (dres) cudaFree(dres);
  return bad ? 3 : 0;
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && std::string(argv[1]) == "quick";
  const int only_cg = argc > 2 ? atoi(argv[2]) : 0;  // 0 = both
  const std::string only_case = argc > 3 ? argv[3] : "";  // run just this big case (for ncu)
  if (argc > 5) hb::gemm_set_l2_hints(atoi(argv[4]), atoi(argv[5]));  // L2 eviction hints of the A / W loads (0 normal, 1 first, 2 last)
  if (argc > 6) hb::gemm_set_balanced_tiles(atoi(argv[6]));
  if (argc > 8 && atoi(argv[8]) != 0) {
    if (cudaMalloc(&g_sched, 8) != cudaSuccess || cudaMemset(g_sched, 0, 8) != cudaSuccess) { printf("sched alloc failed\n"); return 1; }
  }
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  if (hb::tmap_init() != 0) { printf("tmap_init failed\n"); return 1; }
  const int sms = prop.multiProcessorCount;
  int fails = 0;
  // correctness (small, odd M to exercise guards; N with a 128-column tail tile; partial K block)
  Case small[] = {
      {"c1_bf16_small", 300, 384, 256, hb::EPI_BF16, 1, false, true},
      {"c1_bf16_tail", 549, 1408, 1408, hb::EPI_BF16, 1, false, false},
      {"c1_gelu", 549, 512, 592, hb::EPI_GELU_BF16, 1, false, false},
      {"c1_f32_resid", 549, 1408, 704, hb::EPI_F32, 1, true, false},
      {"c1_n48", 200, 48, 128, hb::EPI_F32, 1, false, false},
      {"c2_bf16_small", 300, 384, 256, hb::EPI_BF16, 2, false, true},
      {"c2_bf16_tail", 549, 1408, 1408, hb::EPI_BF16, 2, false, false},
      {"c2_gelu", 549, 512, 592, hb::EPI_GELU_BF16, 2, false, false},
      {"c2_f32_resid", 549, 1408, 704, hb::EPI_F32, 2, true, false},
      {"c2_n96", 200, 96, 128, hb::EPI_F32, 2, false, false},
  };
  for (const Case& c : small) {
    if (only_cg && c.cg != only_cg) continue;
    if (!only_case.empty()) continue;
    int r = run_case(c, sms, true, 0);
    if (r == 2) { printf("sticky CUDA error, aborting remaining cases\n"); return 2; }
    fails += (r != 0);
  }
  if (!quick) {
    const int M = 257 * (argc > 7 ? atoi(argv[7]) : 512);  // 512 frames by default
    Case big[] = {
        {"qkv_cg1", M, 4224, 1408, hb::EPI_BF16, 1, false, true},
        {"qkv_cg2", M, 4224, 1408, hb::EPI_BF16, 2, false, true},
        {"proj_cg1", M, 1408, 1408, hb::EPI_F32, 1, true, false},
        {"proj_cg2", M, 1408, 1408, hb::EPI_F32, 2, true, false},
        {"fc1_cg1", M, 6144, 1408, hb::EPI_GELU_BF16, 1, false, false},
        {"fc1_cg2", M, 6144, 1408, hb::EPI_GELU_BF16, 2, false, false},
        {"fc2_cg1", M, 1408, 6144, hb::EPI_F32, 1, true, false},
        {"fc2_cg2", M, 1408, 6144, hb::EPI_F32, 2, true, false},
    };
    for (const Case& c : big) {
      if (only_cg && c.cg != only_cg) continue;
      if (!only_case.empty() && only_case != c.name) continue;
      int r = run_case(c, sms, false, 5);
      if (r == 2) { printf("sticky CUDA error, aborting\n"); return 2; }
    }
  }
  printf("selftest done, failing cases: %d\n", fails);
  return fails ? 1 : 0;
}
