// hb_api.cu — the C ABI (include/hirest_b200.h): model handles (weights repacked to bf16 once, workspace,
// TMA tensor maps) and the per-call kernel sequences for encode_image / encode_text / retrieval scoring.
#include "../../include/hirest_b200.h"
#include "../../include/hirest_b200_debug.h"

#include <cuda_bf16.h>
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>

#include <array>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <tuple>
#include <vector>

#include "hb_attn.cuh"
#include "hb_elem.cuh"
#include "hb_gemm.cuh"
#include "hb_moment.cuh"
#include "hb_preproc.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
int g_cg = 2;
int g_attn_version = 3;
int g_small_attn_tc = 1;
int g_dec_split_kbs = 6;    // caption decoder: k-blocks (64 bf16) per split-K slice of the hidden-width linears; 0 = no split-K
int g_dec_kv_index = 1;     // caption decoder: beam re-order through an index table instead of copying the KV caches (max_words <= 64)
int g_ln_split_fuse = 1;    // small models: LayerNorm + split-operand conversion in one kernel (ln_split)
int g_decoder_graphs = 1;   // replay the caption decoder's steps as CUDA graphs (from the second search of a shape on)
int g_profile_layer = -1;   // debug: cudaProfilerStart/Stop around this ViT layer (ncu --profile-from-start off)   // fp32 small-sequence attention on tensor cores (hb_attn_tc.cu) instead of CUDA cores
int g_attn_prefetch = 0;   // attention v2: L2-prefetch the operands of the CTA one wave ahead (measured: 1.05 -> 1.14 ms, off)
int g_dyn_sched = 1; // ViT GEMMs take their tiles from an atomic counter (in sequence order) instead of a static round-robin
int g_ln_fold = 1;   // fold the ViT block LayerNorms into the QKV / fc1 GEMM epilogues (no LayerNorm kernel)
int g_num_sms = 148;
bool g_inited = false;
int g_device = -1;

int g_attn_dots_late = 0;   // attention v3: per-tile bit mask, see AttnParams::dots_late
int vit_attn_dispatch(const hb::AttnParams& ap_in, cudaStream_t s) {
  if (g_attn_version == 1) return hb::vit_attn_launch(ap_in, s);
  if (g_attn_version == 2) return hb::vit_attn2_launch(ap_in, s);
  hb::AttnParams ap = ap_in;
  ap.dots_late = g_attn_dots_late;
  return hb::vit_attn3_launch(ap, g_num_sms, s);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define HB_CUDA(x)                                                                                   \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) return fail(HB_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define HB_LAUNCH(x)                                                                                 \
  do {                                                                                               \
    int r_ = (x);                                                                                    \
    if (r_ != 0) {                                                                                   \
      if (r_ > 0) return fail(HB_ERR_CUDA, "%s: %s (%s:%d)", #x, cudaGetErrorString((cudaError_t)r_), __FILE__, __LINE__); \
      return fail(HB_ERR_INVALID, "%s: invalid argument code %d (%s:%d)", #x, r_, __FILE__, __LINE__); \
    }                                                                                                \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                              \
  } while (0)

// ---- optional per-launch timing (bench.py roofline): CUDA events around every launch, on its own stream ----
enum ProfCat { CAT_GEMM_BF16 = 0, CAT_GEMM_GELU = 1, CAT_GEMM_F32 = 2, CAT_VIT_ATTN = 3, CAT_LAYERNORM = 4, CAT_OTHER = 5, CAT_COUNT = 6 };
struct ProfRec { cudaEvent_t e0, e1; int cat; double flops; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t prof_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {
  bool on; ProfRec rec; cudaStream_t s;
  ProfScope(int cat, double flops, cudaStream_t s_) : on(g_prof_on), s(s_) {
    if (on) { rec.cat = cat; rec.flops = flops; rec.e0 = prof_event(); rec.e1 = prof_event(); cudaEventRecord(rec.e0, s); }
  }
  ~ProfScope() { if (on) { cudaEventRecord(rec.e1, s); g_prof.push_back(rec); } }
};
#define HB_LAUNCH_P(cat, flops, stream, x) do { ProfScope ps_((cat), (flops), (stream)); HB_LAUNCH(x); } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t n) {
    if (p) { cudaFree(p); p = nullptr; }
    bytes = n;
    if (n == 0) return 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) { p = nullptr; return fail(HB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e)); }
    return 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// fp32 [N,K] (optionally transposed source [K,N]) -> bf16 [N,Kpad], zero padded.
__global__ void repack_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int K, int Kpad,
                                     int transposed, const float* __restrict__ gamma) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(N) * Kpad) return;
  const int n = static_cast<int>(t / Kpad), k = static_cast<int>(t % Kpad);
  float v = 0.f;
  if (k < K) v = transposed ? src[static_cast<long long>(k) * N + n] : src[static_cast<long long>(n) * K + k];
  if (gamma != nullptr && k < K) v *= gamma[k];   // LayerNorm fold: W' = W * diag(gamma)
  dst[t] = __float2bfloat16(v);
}

// LayerNorm-fold vectors of one Linear (one warp per output feature n):
//   c1[n] = sum_k W'[n,k]  with the bf16-ROUNDED folded weight (exactly what the tensor core multiplies, so a constant row
//           still normalises to beta W^T), c2[n] = sum_k beta_k W[n,k] + bias[n]
__global__ void ln_fold_vectors_kernel(const __nv_bfloat16* __restrict__ wq, int Kpad, const float* __restrict__ w,
                                       const float* __restrict__ beta, const float* __restrict__ bias, float* __restrict__ c1,
                                       float* __restrict__ c2, int N, int K) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  float a = 0.f, b = 0.f;
  for (int k = lane; k < K; k += 32) {
    a += __bfloat162float(wq[static_cast<long long>(n) * Kpad + k]);
    b = fmaf(beta[k], w[static_cast<long long>(n) * K + k], b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    c1[n] = a;
    c2[n] = b + (bias != nullptr ? bias[n] : 0.f);
  }
}

// A Linear layer held by a handle: bf16 weight [N,Kpad] + fp32 bias + TMA map.
struct Linear {
  DevBuf w, b;
  CUtensorMap tm;
  int N = 0, K = 0, Kpad = 0, cg = 2;
  bool has_bias = false;
  // bias may be assembled from up to three pieces (q_bias | zeros | v_bias), each of length N/3.
  DevBuf c1;  // LayerNorm fold (see GemmEpilogue): c1 here, c2 replaces the bias
  int init(const float* w_src, int N_, int K_, const float* bias, bool transposed, cudaStream_t s, const float* bias_q = nullptr,
           const float* bias_v = nullptr, const float* ln_gamma = nullptr, const float* ln_beta = nullptr) {
    N = N_; K = K_; cg = (N_ % 32 == 0) ? g_cg : 1;  // CTA pairs need N % 32 == 0
    Kpad = (K + 7) / 8 * 8;
    if (int r = w.alloc(static_cast<size_t>(N) * Kpad * 2)) return r;
    const long long total = static_cast<long long>(N) * Kpad;
    repack_weight_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w_src, w.as<__nv_bfloat16>(), N, K, Kpad,
                                                                                     transposed ? 1 : 0, ln_gamma);
    HB_CUDA(cudaGetLastError());
    if (bias != nullptr || bias_q != nullptr) {
      has_bias = true;
      if (int r = b.alloc(static_cast<size_t>(N) * 4)) return r;
      if (bias != nullptr) {
        HB_CUDA(cudaMemcpyAsync(b.p, bias, static_cast<size_t>(N) * 4, cudaMemcpyDeviceToDevice, s));
      } else {
        const int D = N / 3;
        HB_CUDA(cudaMemsetAsync(b.p, 0, static_cast<size_t>(N) * 4, s));
        HB_CUDA(cudaMemcpyAsync(b.p, bias_q, static_cast<size_t>(D) * 4, cudaMemcpyDeviceToDevice, s));
        HB_CUDA(cudaMemcpyAsync(b.as<float>() + 2 * D, bias_v, static_cast<size_t>(D) * 4, cudaMemcpyDeviceToDevice, s));
      }
    }
    if (ln_gamma != nullptr) {   // replace the bias by c2 and build c1 (needs the plain bias first, done above)
      if (transposed) return fail(HB_ERR_INVALID, "LayerNorm fold on a transposed weight is not supported");
      if (int r = c1.alloc(static_cast<size_t>(N) * 4)) return r;
      DevBuf c2;
      if (int r = c2.alloc(static_cast<size_t>(N) * 4)) return r;
      ln_fold_vectors_kernel<<<static_cast<unsigned>((N + 7) / 8), 256, 0, s>>>(w.as<__nv_bfloat16>(), Kpad, w_src, ln_beta,
                                                                                has_bias ? b.as<float>() : nullptr, c1.as<float>(),
                                                                                c2.as<float>(), N, K);
      HB_CUDA(cudaGetLastError());
      if (!has_bias) { if (int r = b.alloc(static_cast<size_t>(N) * 4)) return r; has_bias = true; }
      HB_CUDA(cudaMemcpyAsync(b.p, c2.p, static_cast<size_t>(N) * 4, cudaMemcpyDeviceToDevice, s));
      HB_CUDA(cudaStreamSynchronize(s));  // c2 is a temporary
    }
    int r = hb::make_tmap_bf16(&tm, w.p, N, Kpad, Kpad, hb::gemm_w_box_rows(cg));
    if (r) return fail(HB_ERR_CUDA, "cuTensorMapEncodeTiled(weight %dx%d) failed: %d", N, Kpad, r);
    return 0;
  }
  const float* bias() const { return has_bias ? b.as<float>() : nullptr; }
};

// bf16 activation buffer usable as a GEMM A operand.
struct Act {
  DevBuf buf;
  CUtensorMap tm;
  long long rows = 0;
  int cols = 0;
  int init(long long rows_, int cols_) {
    rows = rows_; cols = cols_;
    if (int r = buf.alloc(static_cast<size_t>(rows) * cols * 2)) return r;
    int r = hb::make_tmap_bf16(&tm, buf.p, rows, cols, cols, hb::gemm_a_box_rows());
    if (r) return fail(HB_ERR_CUDA, "cuTensorMapEncodeTiled(act %lldx%d) failed: %d", rows, cols, r);
    return 0;
  }
  __nv_bfloat16* ptr() const { return buf.as<__nv_bfloat16>(); }
};

struct F32Vec {
  DevBuf d;
  int init(const float* src, size_t n, cudaStream_t s) {
    if (int r = d.alloc(n * 4)) return r;
    HB_CUDA(cudaMemcpyAsync(d.p, src, n * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  const float* ptr() const { return d.as<float>(); }
};

struct LnFold {
  const float* stats_in = nullptr;  // consumer side
  float eps = 1e-6f;
  int dim = 0;
  void* xb_out = nullptr;           // producer side
  int ld_xb = 0;
  float* stats_out = nullptr;
  int slots = 1;                    // partial-sum slots per row = 2 * ceil(D / 256)
};

int run_gemm(const CUtensorMap& tmA, const Linear& L, long long M, void* out, int ldo, int epi, cudaStream_t s,
             const float* resid = nullptr, float qscale = 1.f, int qcols = 0, const float* rowadd = nullptr, int remap_in = 0,
             int remap_out = 0, int remap_off = 0, const LnFold* lf = nullptr, int* sched = nullptr) {
  hb::GemmParams p;
  p.sched = g_dyn_sched ? sched : nullptr;
  if (lf != nullptr) {
    p.stats_in = lf->stats_in; p.ln_eps = lf->eps; p.ln_dim = lf->dim; p.c1 = L.c1.as<float>();
    p.xb_out = lf->xb_out; p.ld_xb = lf->ld_xb; p.stats_out = lf->stats_out; p.ln_slots = lf->slots;
  }
  p.M = static_cast<int>(M); p.N = L.N; p.K = L.Kpad;
  p.bias = L.bias(); p.out = out; p.ldo = ldo; p.resid = resid;
  p.qscale = qscale; p.qcols = qcols;
  p.rowadd = rowadd; p.remap_in = remap_in; p.remap_out = remap_out; p.remap_off = remap_off;
  HB_LAUNCH_P((epi == hb::EPI_BF16 || epi == hb::EPI_BF16_LN) ? CAT_GEMM_BF16
                  : ((epi == hb::EPI_GELU_BF16 || epi == hb::EPI_GELU_BF16_LN) ? CAT_GEMM_GELU : CAT_GEMM_F32),
              2.0 * static_cast<double>(M) * L.N * L.K, s, hb::gemm_launch(tmA, L.tm, p, epi, L.cg, g_num_sms, s));
  return 0;
}

// src fp32 [K,N] -> dst fp32 [N,K]
__global__ void transpose_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int K) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<long long>(N) * K) return;
  const int n = static_cast<int>(t / K), k = static_cast<int>(t % K);
  dst[t] = src[static_cast<long long>(k) * N + n];
}

// nn.Linear as a 3-term split-bf16 GEMM: weight [N, 3K] = [hi | lo | hi], activations [R, 3K] = [lo | hi | hi].
struct SplitLinear {
  DevBuf w, b;
  CUtensorMap tm;
  int N = 0, K = 0, cg = 2;
  bool has_bias = false;
  // w_f32: device [N,K], or [K,N] when transposed (x @ P == x (P^T)^T, eva_model.py:249)
  int init_from(const float* w_f32, const float* bias, int N_, int K_, cudaStream_t s, bool transposed = false) {
    N = N_; K = K_; cg = (N_ % 32 == 0) ? g_cg : 1;
    if (int r = w.alloc(static_cast<size_t>(N) * 3 * K * 2)) return r;
    DevBuf wt;
    if (transposed) {
      if (int r = wt.alloc(static_cast<size_t>(N) * K * 4)) return r;
      const long long total = static_cast<long long>(N) * K;
      transpose_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(w_f32, wt.as<float>(), N, K);
      HB_CUDA(cudaGetLastError());
      w_f32 = wt.as<float>();
    }
    if (int r = hb::split3_weight_launch(w_f32, w.as<__nv_bfloat16>(), N, K, s)) return fail(HB_ERR_CUDA, "split3_weight launch failed: %d", r);
    if (transposed) HB_CUDA(cudaStreamSynchronize(s));  // wt is a temporary
    if (bias) {
      has_bias = true;
      if (int r = b.alloc(static_cast<size_t>(N) * 4)) return r;
      HB_CUDA(cudaMemcpyAsync(b.p, bias, static_cast<size_t>(N) * 4, cudaMemcpyDeviceToDevice, s));
    }
    if (hb::make_tmap_bf16(&tm, w.p, N, 3 * K, 3 * K, hb::gemm_w_box_rows(cg))) return fail(HB_ERR_CUDA, "tensor map (split weight) failed");
    return 0;
  }
};


// fp32 activations [rows, K] -> split operand in `op` -> 3-term split-bf16 GEMM with L -> out fp32 [rows, N]
// (act == nullptr: `op` already holds the split operand, written by ln_split below)
int split_gemm_op(__nv_bfloat16* op, const float* act, long long rows, int K, int gelu, const CUtensorMap& tmA, const SplitLinear& L,
                  float* out, int epi, cudaStream_t s, const float* resid = nullptr, const float* rowadd = nullptr, int remap = 0) {
  if (act != nullptr) HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::split3_act_launch(act, op, rows, K, gelu, s));
  hb::GemmParams p;
  p.M = static_cast<int>(rows); p.N = L.N; p.K = 3 * K;
  p.bias = L.has_bias ? L.b.as<float>() : nullptr; p.out = out; p.ldo = L.N; p.resid = resid;
  p.rowadd = rowadd; p.remap_in = remap; p.remap_out = remap; p.remap_off = 0;
  HB_LAUNCH_P(CAT_GEMM_F32, 2.0 * rows * L.N * 3.0 * K, s, hb::gemm_launch(tmA, L.tm, p, epi, L.cg, g_num_sms, s));
  return 0;
}

int ln_f32(const float* x, float* y, const F32Vec& w, const F32Vec& b, float eps, long long rows, int D, cudaStream_t s,
           const int* row_idx = nullptr) {
  hb::LayerNormParams ln;
  ln.x = x; ln.ldx = D; ln.y = y; ln.ldy = D; ln.w = w.ptr(); ln.b = b.ptr(); ln.eps = eps; ln.rows = static_cast<int>(rows); ln.D = D;
  ln.row_idx = row_idx;
  HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s, hb::layernorm_launch(ln, false, s));
  return 0;
}

// LayerNorm whose output feeds a split GEMM: one kernel writes the normalised row as fp32 (y, may be null when only the GEMM
// reads it) and as the [lo | hi | hi] operand in `op` (splitk_finish_kernel with a single slice), instead of a LayerNorm
// kernel + a split kernel re-reading its output.  Falls back to the two kernels when the fusion is off or D > 1536.
bool ln_split_ok(int D) { return g_ln_split_fuse && D % 4 == 0 && D <= 1536; }
int ln_split(const float* x, float* y, __nv_bfloat16* op, const F32Vec& w, const F32Vec& b, float eps, long long rows, int D, cudaStream_t s) {
  HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s,
              hb::splitk_finish_launch(x, 0, 1, nullptr, nullptr, w.ptr(), b.ptr(), eps, 0, y, op, static_cast<int>(rows), D, s));
  return 0;
}

// fp32 attention of the small sequence models: tensor-core kernel when a whole query tile is worth it and the handle's workspace
// fits, CUDA-core kernel otherwise (decode steps with one query, shared K / V across beams).
int small_attn_f32_auto(const hb::SmallAttnF32Params& ap, const DevBuf& ws, cudaStream_t s) {
  if (g_small_attn_tc && ap.Tq >= 16 && ap.kv_div == 1 && ws.p != nullptr && ws.bytes >= hb::small_attn_tc_workspace(ap.B, ap.H, ap.Tq, ap.Tk)) {
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::small_attn_tc_launch(ap, ws.p, ws.bytes, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);   // split prologue + main kernel
    return 0;
  }
  HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::small_attn_f32_launch(ap, s));
  return 0;
}

}  // namespace

// =================================================================================================
// ViT
// =================================================================================================
struct HbVit {
  HbVitConfig cfg;
  int T = 0, Kpatch = 0, max_batch = 0;
  F32Vec cls, pos;
  Linear patch;
  struct Layer {
    F32Vec n1w, n1b, n2w, n2b;
    Linear qkv, proj, fc1, fc2;
  };
  std::vector<std::unique_ptr<Layer>> layers;
  F32Vec nw, nb;
  Linear head;
  Act col, h, hid, clsn, xb;
  DevBuf x, qkv, cls_idx, stats1, stats2, sched;
  bool ln_fold = false;
  int tap_layer = -1;
  float* tap_dst = nullptr;
};

extern "C" {

int hb_init(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return fail(HB_ERR_NODEVICE, "no CUDA device visible");
  if (device < 0 || device >= n) return fail(HB_ERR_INVALID, "device %d out of range (%d devices)", device, n);
  if (g_inited && device != g_device)
    return fail(HB_ERR_INVALID, "already initialised for device %d: one process per GPU (per-device kernel attributes and the SM count are "
                                "process-wide); start another process for device %d", g_device, device);
  int prev_device = -1;
  HB_CUDA(cudaGetDevice(&prev_device));
  HB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  HB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(HB_ERR_NODEVICE, "device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
  if (hb::tmap_init() != 0) return fail(HB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  if (!g_inited) {   // A/B switches from the environment (hirest_b200_debug.h), once
    static const char* keys[] = {"gemm_cta_group", "attention_version", "small_attention_tc", "resize_version", "decoder_graphs", "decoder_split_k", "decoder_kv_index", "ln_split_fuse", "attention_dots_late", "profile_layer", "attention_prefetch", "ln_fold", "gemm_balanced_tiles",
                                 "gemm_dynamic_schedule", "gemm_resid_prefetch_chunks"};
    for (const char* key : keys) {
      std::string env = std::string("HB_DEBUG_") + key;
      for (auto& ch : env) ch = static_cast<char>(toupper(static_cast<unsigned char>(ch)));
      if (const char* v = getenv(env.c_str())) {
        if (int r = hb_debug_set(key, atoi(v))) return r;
      }
    }
  }
  g_inited = true;
  g_device = device;
  if (prev_device >= 0 && prev_device != device) HB_CUDA(cudaSetDevice(prev_device));   // leave the caller's current device alone
  return HB_OK;
}

const char* hb_last_error(void) { return g_err; }

const char* hb_strerror(int code) {
  switch (code) {
    case HB_OK: return "ok";
    case HB_ERR_INVALID: return "invalid argument or unsupported shape";
    case HB_ERR_NOMEM: return "out of device memory";
    case HB_ERR_CUDA: return "CUDA error";
    case HB_ERR_NODEVICE: return "no sm_100 device";
    default: return "unknown error";
  }
}

int64_t hb_launch_count(void) { return g_launches.load(); }

int hb_gemm_n_tiling(int N, int cta_group, int balanced, int* n0, int* width, int cap) {
  const int r = hb::gemm_n_tiling(N, cta_group, balanced, n0, width, cap);
  if (r < 0) return fail(HB_ERR_INVALID, "need N > 0 and cta_group 1 or 2");
  return r;
}

int hb_debug_set(const char* key, int value) {
  if (!key) return fail(HB_ERR_INVALID, "null key");
  const std::string k(key);
  if (k == "gemm_cta_group") {
    if (value != 1 && value != 2) return fail(HB_ERR_INVALID, "cta group must be 1 or 2");
    g_cg = value;
  } else if (k == "attention_version") {
    if (value < 1 || value > 3) return fail(HB_ERR_INVALID, "attention version must be 1, 2 or 3");
    g_attn_version = value;
  } else if (k == "resize_version") {
    if (value < 1 || value > 3) return fail(HB_ERR_INVALID, "resize_version must be 1, 2 or 3");
    hb::resize_set_version(value);
  } else if (k == "small_attention_tc") {
    if (value < 0 || value > 2) return fail(HB_ERR_INVALID, "small_attention_tc must be 0, 1 or 2");
    g_small_attn_tc = value ? 1 : 0;
    if (value) hb::small_attn_tc_set_version(value);
  } else if (k == "decoder_graphs") {
    g_decoder_graphs = value ? 1 : 0;
  } else if (k == "decoder_split_k") {
    if (value < 0) return fail(HB_ERR_INVALID, "decoder_split_k must be >= 0");
    g_dec_split_kbs = value;
  } else if (k == "profile_layer") {
    g_profile_layer = value;
  } else if (k == "decoder_kv_index") {
    g_dec_kv_index = value ? 1 : 0;
  } else if (k == "ln_split_fuse") {
    g_ln_split_fuse = value ? 1 : 0;
  } else if (k == "attention_dots_late") {
    if (value < 0 || value > 3) return fail(HB_ERR_INVALID, "attention_dots_late is a 2-bit tile mask");
    g_attn_dots_late = value;
  } else if (k == "attention_prefetch") {
    g_attn_prefetch = value ? 1 : 0;
  } else if (k == "ln_fold") {
    g_ln_fold = value ? 1 : 0;
  } else if (k == "gemm_balanced_tiles") {
    hb::gemm_set_balanced_tiles(value);
  } else if (k == "gemm_dynamic_schedule") {
    g_dyn_sched = value ? 1 : 0;
  } else if (k == "gemm_resid_prefetch_chunks") {
    hb::gemm_set_resid_prefetch_chunks(value);
  } else {
    return fail(HB_ERR_INVALID, "unknown debug key '%s'", key);
  }
  return HB_OK;
}

int hb_profile_start(void) {
  for (auto& r : g_prof) { g_event_pool.push_back(r.e0); g_event_pool.push_back(r.e1); }
  g_prof.clear();
  g_prof_on = true;
  return HB_OK;
}

int hb_profile_stop(HbProfileSummary* out) {
  g_prof_on = false;
  if (!out) return fail(HB_ERR_INVALID, "null argument");
  HB_CUDA(cudaDeviceSynchronize());
  std::memset(out, 0, sizeof(*out));
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && r.cat >= 0 && r.cat < HB_PROF_CATS) {
      out->ms[r.cat] += ms;
      out->flops[r.cat] += r.flops;
      out->launches[r.cat] += 1;
    }
    g_event_pool.push_back(r.e0);
    g_event_pool.push_back(r.e1);
  }
  g_prof.clear();
  return HB_OK;
}

int hb_vit_create(const HbVitConfig* cfg, const HbVitWeights* w, int max_batch, void* stream, HbVit** out) {
  if (!g_inited) return fail(HB_ERR_INVALID, "hb_init() not called");
  if (!cfg || !w || !out || max_batch <= 0) return fail(HB_ERR_INVALID, "null argument");
  const int D = cfg->width, H = cfg->heads, F = cfg->mlp_hidden, E = cfg->embed_dim, P = cfg->patch_size, S = cfg->image_size;
  if (D != H * 88) return fail(HB_ERR_INVALID, "head_dim must be 88 (width %d, heads %d)", D, H);
  if (S != 224 || P != 14) return fail(HB_ERR_INVALID, "only 224x224 / patch 14 (257 tokens) is supported");  // vit_model.py:203-204
  if (D % 16 || F % 16 || E % 16) return fail(HB_ERR_INVALID, "width/mlp_hidden/embed_dim must be multiples of 16");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::unique_ptr<HbVit> m(new (std::nothrow) HbVit);
  if (!m) return fail(HB_ERR_NOMEM, "host allocation failed");
  m->cfg = *cfg;
  m->T = (S / P) * (S / P) + 1;
  m->Kpatch = 3 * P * P;
  m->max_batch = max_batch;
  m->ln_fold = (g_ln_fold != 0);
  int r;
  if ((r = m->cls.init(w->cls_token, D, s))) return r;
  if ((r = m->pos.init(w->pos_embed, static_cast<size_t>(m->T) * D, s))) return r;
  if ((r = m->patch.init(w->patch_w, D, m->Kpatch, w->patch_b, false, s))) return r;
  for (int i = 0; i < cfg->layers; ++i) {
    std::unique_ptr<HbVit::Layer> L(new HbVit::Layer);
    if ((r = L->n1w.init(w->norm1_w[i], D, s))) return r;
    if ((r = L->n1b.init(w->norm1_b[i], D, s))) return r;
    if ((r = L->n2w.init(w->norm2_w[i], D, s))) return r;
    if ((r = L->n2b.init(w->norm2_b[i], D, s))) return r;
    // qkv bias = cat(q_bias, zeros, v_bias), built once instead of every forward (vit_model.py:124)
    if ((r = L->qkv.init(w->qkv_w[i], 3 * D, D, nullptr, false, s, w->q_bias[i], w->v_bias[i], m->ln_fold ? w->norm1_w[i] : nullptr,
                         m->ln_fold ? w->norm1_b[i] : nullptr)))
      return r;
    if ((r = L->proj.init(w->proj_w[i], D, D, w->proj_b[i], false, s))) return r;
    if ((r = L->fc1.init(w->fc1_w[i], F, D, w->fc1_b[i], false, s, nullptr, nullptr, m->ln_fold ? w->norm2_w[i] : nullptr,
                         m->ln_fold ? w->norm2_b[i] : nullptr)))
      return r;
    if ((r = L->fc2.init(w->fc2_w[i], D, F, w->fc2_b[i], false, s))) return r;
    m->layers.push_back(std::move(L));
  }
  if ((r = m->nw.init(w->norm_w, D, s))) return r;
  if ((r = m->nb.init(w->norm_b, D, s))) return r;
  if ((r = m->head.init(w->head_w, E, D, w->head_b, false, s))) return r;
  const long long rows = static_cast<long long>(max_batch) * m->T;
  if ((r = m->col.init(static_cast<long long>(max_batch) * (m->T - 1), m->patch.Kpad))) return r;
  if ((r = m->h.init(rows, D))) return r;
  if ((r = m->hid.init(rows, F))) return r;
  if ((r = m->clsn.init(max_batch, D))) return r;
  if (m->ln_fold) {
    if ((r = m->xb.init(rows, D))) return r;
    const size_t slots = static_cast<size_t>(2 * ((D + 255) / 256));   // one per (N tile, column half) of the producing GEMM
    if ((r = m->stats1.alloc(static_cast<size_t>(rows) * slots * 8))) return r;
    if ((r = m->stats2.alloc(static_cast<size_t>(rows) * slots * 8))) return r;
    // a slot whose column half is empty (last tile narrower than 128) is never written: it must read as zero
    HB_CUDA(cudaMemsetAsync(m->stats1.p, 0, m->stats1.bytes, s));
    HB_CUDA(cudaMemsetAsync(m->stats2.p, 0, m->stats2.bytes, s));
  }
  if ((r = m->x.alloc(static_cast<size_t>(rows) * D * 4))) return r;
  if ((r = m->qkv.alloc(static_cast<size_t>(rows) * 3 * D * 2))) return r;
  if ((r = m->cls_idx.alloc(static_cast<size_t>(max_batch) * 4))) return r;
  if ((r = m->sched.alloc(2 * sizeof(int)))) return r;   // dynamic tile scheduler counters (re-zeroed by every GEMM that uses them)
  HB_CUDA(cudaMemsetAsync(m->sched.p, 0, m->sched.bytes, s));
  {
    std::vector<int> idx(max_batch);
    for (int i = 0; i < max_batch; ++i) idx[i] = i * m->T;
    HB_CUDA(cudaMemcpyAsync(m->cls_idx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, s));
    HB_CUDA(cudaStreamSynchronize(s));  // idx is a host temporary; also surfaces repack errors at create time
  }
  *out = m.release();
  return HB_OK;
}

int hb_vit_set_tap(HbVit* m, int layer, float* dst) {
  if (!m) return fail(HB_ERR_INVALID, "null handle");
  m->tap_layer = dst ? layer : -1;
  m->tap_dst = dst;
  return HB_OK;
}

static int vit_encode_chunk(HbVit* m, const float* frames, const uint8_t* frames_u8, const float* mean, const float* stdv, int B,
                            float* out, cudaStream_t s) {
  const HbVitConfig& c = m->cfg;
  const int D = c.width, T = m->T, F = c.mlp_hidden, E = c.embed_dim;
  const long long M = static_cast<long long>(B) * T;
  float* x = m->x.as<float>();
  __nv_bfloat16* qkv = m->qkv.as<__nv_bfloat16>();
  int r;
  // patch embed: gather -> GEMM (+bias +pos, rows remapped past the cls slot); cls row separately
  if (frames_u8 != nullptr)
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::im2col_patch_u8_launch(frames_u8, m->col.ptr(), B, c.image_size, c.patch_size, m->patch.Kpad, mean, stdv, s));
  else
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::im2col_patch_launch(frames, m->col.ptr(), B, c.image_size, c.patch_size, m->patch.Kpad, s));
  if ((r = run_gemm(m->col.tm, m->patch, static_cast<long long>(B) * (T - 1), x, D, hb::EPI_F32_ROWADD, s, nullptr, 1.f, 0,
                    m->pos.ptr() + D, T - 1, T, 1)))
    return r;
  HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::cls_row_launch(x, m->cls.ptr(), m->pos.ptr(), B, T, D, s));
  if (m->tap_layer == 0 && m->tap_dst)
    HB_CUDA(cudaMemcpyAsync(m->tap_dst, x, static_cast<size_t>(M) * D * 4, cudaMemcpyDeviceToDevice, s));
  const float qscale = 1.0f / sqrtf(88.0f);
  int* sched = m->sched.as<int>();
  if (m->ln_fold) {
    // LayerNorm folded into the GEMMs: the residual stream travels as fp32 x + a bf16 copy xb + per-row (sum, sumsq)
    // partials, one slot per 128 columns, each written by exactly one warp (no atomics: bit-reproducible, batch-independent).
    // QKV / fc1 read xb and apply rstd / mean in their epilogue, proj / fc2 refresh xb and the statistics in theirs.
    const int slots = 2 * ((D + 255) / 256);
    HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s, hb::row_stats_launch(x, m->xb.ptr(), m->stats1.as<float>(), M, D, slots, s));
    LnFold lf1, lf2, lp1, lp2;
    lf1.slots = lf2.slots = lp1.slots = lp2.slots = slots;
    lf1.stats_in = m->stats1.as<float>(); lf1.eps = c.ln_eps; lf1.dim = D;
    lf2.stats_in = m->stats2.as<float>(); lf2.eps = c.ln_eps; lf2.dim = D;
    lp2.xb_out = m->xb.ptr(); lp2.ld_xb = D; lp2.stats_out = m->stats2.as<float>();   // proj  -> statistics for LN2
    lp1.xb_out = m->xb.ptr(); lp1.ld_xb = D; lp1.stats_out = m->stats1.as<float>();   // fc2   -> statistics for the next LN1
    for (int i = 0; i < c.layers; ++i) {
      HbVit::Layer& L = *m->layers[i];
      if (i == g_profile_layer) { cudaStreamSynchronize(s); cudaProfilerStart(); }
      if ((r = run_gemm(m->xb.tm, L.qkv, M, qkv, 3 * D, hb::EPI_BF16_LN, s, nullptr, qscale, D, nullptr, 0, 0, 0, &lf1, sched))) return r;
      hb::AttnParams ap;
      ap.qkv = qkv; ap.out = m->h.ptr(); ap.B = B; ap.H = c.heads; ap.prefetch_ahead = g_attn_prefetch ? 2 * g_num_sms : 0;
      HB_LAUNCH_P(CAT_VIT_ATTN, 4.0 * B * c.heads * 257.0 * 257.0 * 88.0, s, vit_attn_dispatch(ap, s));
      if ((r = run_gemm(m->h.tm, L.proj, M, x, D, hb::EPI_F32_STATS, s, x, 1.f, 0, nullptr, 0, 0, 0, &lp2, sched))) return r;
      if ((r = run_gemm(m->xb.tm, L.fc1, M, m->hid.ptr(), F, hb::EPI_GELU_BF16_LN, s, nullptr, 1.f, 0, nullptr, 0, 0, 0, &lf2, sched))) return r;
      if ((r = run_gemm(m->hid.tm, L.fc2, M, x, D, hb::EPI_F32_STATS, s, x, 1.f, 0, nullptr, 0, 0, 0, &lp1, sched))) return r;
      if (i == g_profile_layer) { cudaStreamSynchronize(s); cudaProfilerStop(); }
      if (m->tap_layer == i + 1 && m->tap_dst)
        HB_CUDA(cudaMemcpyAsync(m->tap_dst, x, static_cast<size_t>(M) * D * 4, cudaMemcpyDeviceToDevice, s));
    }
  } else
  for (int i = 0; i < c.layers; ++i) {
    HbVit::Layer& L = *m->layers[i];
    hb::LayerNormParams ln;
    ln.x = x; ln.ldx = D; ln.y = m->h.ptr(); ln.ldy = D; ln.w = L.n1w.ptr(); ln.b = L.n1b.ptr();
    ln.eps = c.ln_eps; ln.rows = static_cast<int>(M); ln.D = D;
    HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s, hb::layernorm_launch(ln, true, s));
    if ((r = run_gemm(m->h.tm, L.qkv, M, qkv, 3 * D, hb::EPI_BF16, s, nullptr, qscale, D, nullptr, 0, 0, 0, nullptr, sched))) return r;
    hb::AttnParams ap;
    ap.qkv = qkv; ap.out = m->h.ptr(); ap.B = B; ap.H = c.heads; ap.prefetch_ahead = g_attn_prefetch ? 2 * g_num_sms : 0;
    HB_LAUNCH_P(CAT_VIT_ATTN, 4.0 * B * c.heads * 257.0 * 257.0 * 88.0, s, vit_attn_dispatch(ap, s));
    if ((r = run_gemm(m->h.tm, L.proj, M, x, D, hb::EPI_F32, s, x, 1.f, 0, nullptr, 0, 0, 0, nullptr, sched))) return r;
    ln.w = L.n2w.ptr(); ln.b = L.n2b.ptr();
    HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s, hb::layernorm_launch(ln, true, s));
    if ((r = run_gemm(m->h.tm, L.fc1, M, m->hid.ptr(), F, hb::EPI_GELU_BF16, s, nullptr, 1.f, 0, nullptr, 0, 0, 0, nullptr, sched))) return r;
    if ((r = run_gemm(m->hid.tm, L.fc2, M, x, D, hb::EPI_F32, s, x, 1.f, 0, nullptr, 0, 0, 0, nullptr, sched))) return r;
    if (m->tap_layer == i + 1 && m->tap_dst)
      HB_CUDA(cudaMemcpyAsync(m->tap_dst, x, static_cast<size_t>(M) * D * 4, cudaMemcpyDeviceToDevice, s));
  }
  // final norm on the cls rows only (LayerNorm is row-wise; rows 1..256 are never read, vit_model.py:340-346)
  hb::LayerNormParams ln;
  ln.x = x; ln.ldx = D; ln.row_idx = m->cls_idx.as<int>(); ln.y = m->clsn.ptr(); ln.ldy = D;
  ln.w = m->nw.ptr(); ln.b = m->nb.ptr(); ln.eps = c.ln_eps; ln.rows = B; ln.D = D;
  HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s, hb::layernorm_launch(ln, true, s));
  if ((r = run_gemm(m->clsn.tm, m->head, B, out, E, hb::EPI_F32, s))) return r;
  return HB_OK;
}

int hb_vit_encode(HbVit* m, const float* frames, int64_t B, float* out, void* stream) {
  if (B == 0 && m) return HB_OK;  // empty batch: nothing to do (pointers of empty tensors may be null)
  if (!m || !frames || !out) return fail(HB_ERR_INVALID, "null argument");
  if (B < 0) return fail(HB_ERR_INVALID, "negative batch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t frame_elems = static_cast<size_t>(3) * m->cfg.image_size * m->cfg.image_size;
  for (int64_t b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = static_cast<int>(std::min<int64_t>(m->max_batch, B - b0));
    int r = vit_encode_chunk(m, frames + b0 * frame_elems, nullptr, nullptr, nullptr, nb, out + b0 * m->cfg.embed_dim, s);
    if (r) return r;
  }
  return HB_OK;
}

int hb_vit_encode_u8(HbVit* m, const uint8_t* frames, int64_t B, const float* mean, const float* stdv, float* out, void* stream) {
  if (B == 0 && m) return HB_OK;
  if (!m || !frames || !out || !mean || !stdv) return fail(HB_ERR_INVALID, "null argument");
  if (B < 0) return fail(HB_ERR_INVALID, "negative batch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t frame_elems = static_cast<size_t>(3) * m->cfg.image_size * m->cfg.image_size;
  for (int64_t b0 = 0; b0 < B; b0 += m->max_batch) {
    const int nb = static_cast<int>(std::min<int64_t>(m->max_batch, B - b0));
    int r = vit_encode_chunk(m, nullptr, frames + b0 * frame_elems, mean, stdv, nb, out + b0 * m->cfg.embed_dim, s);
    if (r) return r;
  }
  return HB_OK;
}

void hb_vit_destroy(HbVit* m) { delete m; }

}  // extern "C"

// =================================================================================================
// Text tower
// =================================================================================================
struct HbText {
  HbTextConfig cfg;
  int max_batch = 0;
  F32Vec tok, pos;
  struct Layer {
    F32Vec l1w, l1b, l2w, l2b;
    Linear qkv, out, fc, cproj;
  };
  std::vector<std::unique_ptr<Layer>> layers;
  F32Vec lfw, lfb;
  Linear proj;
  Act h, hid, eot;
  DevBuf x, qkv, eot_row;
  // precise mode (cfg.precise != 0, the default of the Python surface): every Linear as a 3-term split-bf16 GEMM, LayerNorm /
  // attention / residual stream in fp32 -- encode_text within ~1e-5 of the fp32 reference, so that the integer decisions the
  // MomentModel derives from the text feature (argmax, region growing, beam top-k) match the reference's.
  bool precise = false;
  int chunk = 0;   // queries per pass
  struct PLayer { SplitLinear qkv, out, fc, cproj; };
  std::vector<std::unique_ptr<PLayer>> players;
  SplitLinear proj_s;
  DevBuf op, opE, lnb, qkvf, att, mid, eotf, attn_ws;
  CUtensorMap tm_w, tm_4w, tm_eot;
};

extern "C" {

int hb_text_create(const HbTextConfig* cfg, const HbTextWeights* w, int max_batch, void* stream, HbText** out) {
  if (!g_inited) return fail(HB_ERR_INVALID, "hb_init() not called");
  if (!cfg || !w || !out || max_batch <= 0) return fail(HB_ERR_INVALID, "null argument");
  const int W = cfg->width, C = cfg->context_length, E = cfg->embed_dim;
  if (W != cfg->heads * 64) return fail(HB_ERR_INVALID, "text head_dim must be 64 (width %d, heads %d)", W, cfg->heads);
  if (W % 16 || E % 16) return fail(HB_ERR_INVALID, "width/embed_dim must be multiples of 16");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::unique_ptr<HbText> m(new (std::nothrow) HbText);
  if (!m) return fail(HB_ERR_NOMEM, "host allocation failed");
  m->cfg = *cfg;
  m->max_batch = max_batch;
  m->precise = (cfg->precise != 0);
  m->chunk = m->precise ? std::min(max_batch, 128) : max_batch;
  int r;
  if ((r = m->tok.init(w->token_embedding, static_cast<size_t>(cfg->vocab_size) * W, s))) return r;
  if ((r = m->pos.init(w->positional_embedding, static_cast<size_t>(C) * W, s))) return r;
  for (int i = 0; i < cfg->layers; ++i) {
    std::unique_ptr<HbText::Layer> L(new HbText::Layer);
    if ((r = L->l1w.init(w->ln1_w[i], W, s))) return r;
    if ((r = L->l1b.init(w->ln1_b[i], W, s))) return r;
    if ((r = L->l2w.init(w->ln2_w[i], W, s))) return r;
    if ((r = L->l2b.init(w->ln2_b[i], W, s))) return r;
    if (m->precise) {
      std::unique_ptr<HbText::PLayer> P(new HbText::PLayer);
      if ((r = P->qkv.init_from(w->in_proj_w[i], w->in_proj_b[i], 3 * W, W, s))) return r;
      if ((r = P->out.init_from(w->out_proj_w[i], w->out_proj_b[i], W, W, s))) return r;
      if ((r = P->fc.init_from(w->fc_w[i], w->fc_b[i], 4 * W, W, s))) return r;
      if ((r = P->cproj.init_from(w->cproj_w[i], w->cproj_b[i], W, 4 * W, s))) return r;
      m->players.push_back(std::move(P));
    } else {
      if ((r = L->qkv.init(w->in_proj_w[i], 3 * W, W, w->in_proj_b[i], false, s))) return r;
      if ((r = L->out.init(w->out_proj_w[i], W, W, w->out_proj_b[i], false, s))) return r;
      if ((r = L->fc.init(w->fc_w[i], 4 * W, W, w->fc_b[i], false, s))) return r;
      if ((r = L->cproj.init(w->cproj_w[i], W, 4 * W, w->cproj_b[i], false, s))) return r;
    }
    m->layers.push_back(std::move(L));
  }
  if ((r = m->lfw.init(w->ln_final_w, W, s))) return r;
  if ((r = m->lfb.init(w->ln_final_b, W, s))) return r;
  const long long rows = static_cast<long long>(m->chunk) * C;
  if ((r = m->x.alloc(static_cast<size_t>(rows) * W * 4))) return r;
  if ((r = m->eot_row.alloc(static_cast<size_t>(m->chunk) * 4))) return r;
  if (m->precise) {
    if ((r = m->proj_s.init_from(w->text_projection, nullptr, E, W, s, /*transposed=*/true))) return r;  // x @ P == x (P^T)^T
    if ((r = m->op.alloc(static_cast<size_t>(rows) * 3 * 4 * W * 2))) return r;
    if ((r = m->opE.alloc(static_cast<size_t>(m->chunk) * 3 * W * 2))) return r;
    if ((r = m->lnb.alloc(static_cast<size_t>(rows) * W * 4))) return r;
    if ((r = m->qkvf.alloc(static_cast<size_t>(rows) * 3 * W * 4))) return r;
    if ((r = m->att.alloc(static_cast<size_t>(rows) * W * 4))) return r;
    if ((r = m->mid.alloc(static_cast<size_t>(rows) * 4 * W * 4))) return r;
    if ((r = m->eotf.alloc(static_cast<size_t>(m->chunk) * W * 4))) return r;
    if ((r = m->attn_ws.alloc(hb::small_attn_tc_workspace(m->chunk, cfg->heads, C, C)))) return r;
    if (hb::make_tmap_bf16(&m->tm_w, m->op.p, rows, 3 * W, 3 * W, hb::gemm_a_box_rows()) ||
        hb::make_tmap_bf16(&m->tm_4w, m->op.p, rows, 12 * W, 12 * W, hb::gemm_a_box_rows()) ||
        hb::make_tmap_bf16(&m->tm_eot, m->opE.p, m->chunk, 3 * W, 3 * W, hb::gemm_a_box_rows()))
      return fail(HB_ERR_CUDA, "tensor map (text operands) failed");
  } else {
    if ((r = m->proj.init(w->text_projection, E, W, nullptr, /*transposed=*/true, s))) return r;  // x @ P == x P'^T, P' = P^T
    if ((r = m->h.init(rows, W))) return r;
    if ((r = m->hid.init(rows, 4 * W))) return r;
    if ((r = m->eot.init(m->chunk, W))) return r;
    if ((r = m->qkv.alloc(static_cast<size_t>(rows) * 3 * W * 2))) return r;
  }
  HB_CUDA(cudaStreamSynchronize(s));
  *out = m.release();
  return HB_OK;
}

static int text_encode_chunk(HbText* m, const int64_t* ids, int Q, float* out, cudaStream_t s) {
  const HbTextConfig& c = m->cfg;
  const int W = c.width, C = c.context_length, E = c.embed_dim;
  const long long M = static_cast<long long>(Q) * C;
  float* x = m->x.as<float>();
  __nv_bfloat16* qkv = m->qkv.as<__nv_bfloat16>();
  int r;
  HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::text_embed_launch(reinterpret_cast<const long long*>(ids), m->tok.ptr(), m->pos.ptr(), x,
                                  m->eot_row.as<int>(), Q, C, W, c.vocab_size, s));
  if (m->precise) {
    __nv_bfloat16* op = m->op.as<__nv_bfloat16>();
    float *lnb = m->lnb.as<float>(), *qf = m->qkvf.as<float>(), *att = m->att.as<float>(), *mid = m->mid.as<float>();
    for (int i = 0; i < c.layers; ++i) {
      HbText::Layer& L = *m->layers[i];
      HbText::PLayer& P = *m->players[i];
      const bool fuse = ln_split_ok(W);
      if (fuse) { if ((r = ln_split(x, nullptr, op, L.l1w, L.l1b, c.ln_eps, M, W, s))) return r; }
      else if ((r = ln_f32(x, lnb, L.l1w, L.l1b, c.ln_eps, M, W, s))) return r;
      if ((r = split_gemm_op(op, fuse ? nullptr : lnb, M, W, 0, m->tm_w, P.qkv, qf, hb::EPI_F32, s))) return r;
      hb::SmallAttnF32Params ap;
      ap.q = qf; ap.k = qf + W; ap.v = qf + 2 * W; ap.out = att;
      ap.B = Q; ap.H = c.heads; ap.Tq = C; ap.Tk = C;
      ap.ldq = ap.ldk = ap.ldv = 3 * W; ap.ldo = W;
      ap.bsq = ap.bsk = ap.bsv = static_cast<long long>(C) * 3 * W; ap.bso = static_cast<long long>(C) * W;
      ap.scale = 0.125f; ap.mask_mode = 1;   // causal, eva_model.py:224-230
      if ((r = small_attn_f32_auto(ap, m->attn_ws, s))) return r;
      if ((r = split_gemm_op(op, att, M, W, 0, m->tm_w, P.out, x, hb::EPI_F32, s, x))) return r;
      if (fuse) { if ((r = ln_split(x, nullptr, op, L.l2w, L.l2b, c.ln_eps, M, W, s))) return r; }
      else if ((r = ln_f32(x, lnb, L.l2w, L.l2b, c.ln_eps, M, W, s))) return r;
      if ((r = split_gemm_op(op, fuse ? nullptr : lnb, M, W, 0, m->tm_w, P.fc, mid, hb::EPI_F32, s))) return r;
      if ((r = split_gemm_op(op, mid, M, 4 * W, /*gelu=*/1, m->tm_4w, P.cproj, x, hb::EPI_F32, s, x))) return r;
    }
    if ((r = ln_f32(x, m->eotf.as<float>(), m->lfw, m->lfb, c.ln_eps, Q, W, s, m->eot_row.as<int>()))) return r;
    if ((r = split_gemm_op(m->opE.as<__nv_bfloat16>(), m->eotf.as<float>(), Q, W, 0, m->tm_eot, m->proj_s, out, hb::EPI_F32, s))) return r;
    return HB_OK;
  }
  for (int i = 0; i < c.layers; ++i) {
    HbText::Layer& L = *m->layers[i];
    hb::LayerNormParams ln;
    ln.x = x; ln.ldx = W; ln.y = m->h.ptr(); ln.ldy = W; ln.w = L.l1w.ptr(); ln.b = L.l1b.ptr();
    ln.eps = c.ln_eps; ln.rows = static_cast<int>(M); ln.D = W;
    HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s, hb::layernorm_launch(ln, true, s));
    if ((r = run_gemm(m->h.tm, L.qkv, M, qkv, 3 * W, hb::EPI_BF16, s))) return r;
    hb::SmallAttnParams ap;
    ap.q = qkv; ap.k = qkv + W; ap.v = qkv + 2 * W; ap.out = m->h.ptr();
    ap.B = Q; ap.H = c.heads; ap.Tq = C; ap.Tk = C;
    ap.ldq = ap.ldk = ap.ldv = 3 * W; ap.ldo = W;
    ap.bsq = ap.bsk = ap.bsv = static_cast<long long>(C) * 3 * W; ap.bso = static_cast<long long>(C) * W;
    ap.scale = 0.125f;   // head_dim^-0.5, nn.MultiheadAttention
    ap.mask_mode = 1;    // causal, eva_model.py:224-230
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::small_attn_launch(ap, s));
    if ((r = run_gemm(m->h.tm, L.out, M, x, W, hb::EPI_F32, s, x))) return r;
    ln.w = L.l2w.ptr(); ln.b = L.l2b.ptr();
    HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s, hb::layernorm_launch(ln, true, s));
    if ((r = run_gemm(m->h.tm, L.fc, M, m->hid.ptr(), 4 * W, hb::EPI_GELU_BF16, s))) return r;
    if ((r = run_gemm(m->hid.tm, L.cproj, M, x, W, hb::EPI_F32, s, x))) return r;
  }
  // ln_final on the EOT rows only (row-wise op; eva_model.py:239-243), then @ text_projection (:249)
  hb::LayerNormParams ln;
  ln.x = x; ln.ldx = W; ln.row_idx = m->eot_row.as<int>(); ln.y = m->eot.ptr(); ln.ldy = W;
  ln.w = m->lfw.ptr(); ln.b = m->lfb.ptr(); ln.eps = c.ln_eps; ln.rows = Q; ln.D = W;
  HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s, hb::layernorm_launch(ln, true, s));
  if ((r = run_gemm(m->eot.tm, m->proj, Q, out, E, hb::EPI_F32, s))) return r;
  return HB_OK;
}

int hb_text_encode(HbText* m, const int64_t* ids, int64_t Q, float* out, void* stream) {
  if (Q == 0 && m) return HB_OK;
  if (!m || !ids || !out) return fail(HB_ERR_INVALID, "null argument");
  if (Q < 0) return fail(HB_ERR_INVALID, "negative batch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int64_t q0 = 0; q0 < Q; q0 += m->chunk) {
    const int nq = static_cast<int>(std::min<int64_t>(m->chunk, Q - q0));
    int r = text_encode_chunk(m, ids + q0 * m->cfg.context_length, nq, out + q0 * m->cfg.embed_dim, s);
    if (r) return r;
  }
  return HB_OK;
}

void hb_text_destroy(HbText* m) { delete m; }

// =================================================================================================
// retrieval scoring + generic ops
// =================================================================================================
int hb_pool_normalize(const float* emb, int64_t V, int F, int E, float* out, void* stream) {
  if (!emb || !out) return fail(HB_ERR_INVALID, "null argument");
  if (V == 0) return HB_OK;
  HB_LAUNCH(hb::pool_normalize_launch(emb, out, V, F, E, true, false, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

int hb_subsample_pool_normalize(const float* feats, const int64_t* offsets, int64_t V, int n_sub, int E, float* out, void* stream) {
  if (V == 0) return HB_OK;
  if (!feats || !offsets || !out) return fail(HB_ERR_INVALID, "null argument");
  if (V < 0 || E <= 0 || E > 4096) return fail(HB_ERR_INVALID, "need V >= 0 and 1 <= E <= 4096");
  HB_LAUNCH(hb::subsample_pool_normalize_launch(feats, reinterpret_cast<const long long*>(offsets), out, V, n_sub, E,
                                                static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

int hb_subsample_pool_normalize_bf16(const void* feats, const int64_t* offsets, int64_t V, int n_sub, int E, float* out, void* stream) {
  if (V == 0) return HB_OK;
  if (!feats || !offsets || !out) return fail(HB_ERR_INVALID, "null argument");
  if (V < 0 || E <= 0 || E > 4096) return fail(HB_ERR_INVALID, "need V >= 0 and 1 <= E <= 4096");
  HB_LAUNCH(hb::subsample_pool_normalize_bf16_launch(static_cast<const __nv_bfloat16*>(feats), reinterpret_cast<const long long*>(offsets), out,
                                                     V, n_sub, E, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

int hb_resample_rows(const float* feats, const int64_t* offsets, int64_t V, int n_out, int C, float* out, void* stream) {
  if (V == 0) return HB_OK;
  if (!feats || !offsets || !out) return fail(HB_ERR_INVALID, "null argument");
  if (V < 0 || n_out <= 0 || C <= 0 || C % 4) return fail(HB_ERR_INVALID, "need V >= 0, n_out > 0 and C %% 4 == 0");
  HB_LAUNCH(hb::resample_rows_launch(feats, reinterpret_cast<const long long*>(offsets), out, V, n_out, C, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

int hb_asr_warp(const float* asr, const int64_t* sub_offsets, const int32_t* starts, const int32_t* ends, const int64_t* frame_offsets,
                const int32_t* row_video, int64_t rows, int C, float* out, void* stream) {
  if (rows == 0) return HB_OK;
  if (!sub_offsets || !frame_offsets || !row_video || !out) return fail(HB_ERR_INVALID, "null argument");
  if (rows < 0 || C <= 0 || C % 4) return fail(HB_ERR_INVALID, "need rows >= 0 and C %% 4 == 0");
  HB_LAUNCH(hb::asr_warp_launch(asr, reinterpret_cast<const long long*>(sub_offsets), starts, ends, reinterpret_cast<const long long*>(frame_offsets),
                                row_video, out, rows, C, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

// ---- frame preprocessing: plans (host tables + device copy) cached per (device, H, W, S) ----
namespace {
struct ResizeEntry { hb::ResizePlanHost plan; DevBuf tables; };
std::mutex g_resize_mu;
std::map<std::tuple<int, int, int, int>, std::unique_ptr<ResizeEntry>> g_resize_plans;
}  // namespace

int hb_resize_geometry(int H, int W, int S, int* new_h, int* new_w, int* top, int* left) {
  if (!new_h || !new_w || !top || !left) return fail(HB_ERR_INVALID, "null argument");
  hb::ResizePlanHost plan;
  const int r = hb::resize_plan_build(&plan, H, W, S);
  if (r == -3) return fail(HB_ERR_INVALID, "invalid frame size %dx%d -> %d (need H, W >= 1 and 1 <= S <= 256)", H, W, S);
  *new_h = plan.nh; *new_w = plan.nw; *top = plan.top; *left = plan.left;
  return HB_OK;
}

int64_t hb_resize_tables(int H, int W, int S, int info[6], int* tables, int64_t cap) {
  if (!info) return fail(HB_ERR_INVALID, "null argument");
  hb::ResizePlanHost plan;
  const int r = hb::resize_plan_build(&plan, H, W, S);
  if (r) return fail(HB_ERR_INVALID, "resize plan for %dx%d -> %d failed (%d)", H, W, S, r);
  info[0] = plan.kh; info[1] = plan.kv; info[2] = plan.x0; info[3] = plan.span_bytes; info[4] = plan.ty; info[5] = plan.smem_bytes;
  const int64_t n = static_cast<int64_t>(hb::resize_plan_table_ints(plan));
  if (tables && cap >= n) hb::resize_plan_pack(plan, tables);
  return n;
}

int hb_resize_crop_u8(const uint8_t* src, int64_t B, int H, int W, int S, uint8_t* dst, void* stream) {
  if (!g_inited) return fail(HB_ERR_INVALID, "hb_init() not called");
  if (B == 0) return HB_OK;
  if (!src || !dst) return fail(HB_ERR_INVALID, "null argument");
  if (B < 0) return fail(HB_ERR_INVALID, "negative batch");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int dev = 0;
  HB_CUDA(cudaGetDevice(&dev));
  ResizeEntry* e = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_resize_mu);
    auto key = std::make_tuple(dev, H, W, S);
    auto it = g_resize_plans.find(key);
    if (it == g_resize_plans.end()) {
      std::unique_ptr<ResizeEntry> ne(new ResizeEntry);
      const int r = hb::resize_plan_build(&ne->plan, H, W, S);
      if (r == -3) return fail(HB_ERR_INVALID, "invalid frame size %dx%d -> %d (need H, W >= 1 and 1 <= S <= 256)", H, W, S);
      if (r == -6) return fail(HB_ERR_INVALID, "frame %dx%d is too large for the shared-memory row window of the resize kernel", H, W);
      if (r) return fail(HB_ERR_INVALID, "resize plan failed (%d)", r);
      const size_t n1 = hb::resize_plan_table_ints(ne->plan);
      std::vector<int> packed(n1 + hb::resize_plan_table_ints_v2(ne->plan));
      hb::resize_plan_pack(ne->plan, packed.data());
      hb::resize_plan_pack_v2(ne->plan, packed.data() + n1);
      if (int a = ne->tables.alloc(packed.size() * sizeof(int))) return a;
      HB_CUDA(cudaMemcpy(ne->tables.p, packed.data(), packed.size() * sizeof(int), cudaMemcpyHostToDevice));  // first use of this size only
      it = g_resize_plans.emplace(key, std::move(ne)).first;
    }
    e = it->second.get();
  }
  HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::resize_crop_launch(e->plan, e->tables.as<int>(), src, dst, B, s));
  return HB_OK;
}

int hb_similarity(const float* text, int64_t Q, const float* video, int64_t V, int E, float* scores, int64_t ld_scores,
                  int exact, void* stream) {
  if (!g_inited) return fail(HB_ERR_INVALID, "hb_init() not called");
  if (!text || !video || !scores) return fail(HB_ERR_INVALID, "null argument");
  if (Q == 0 || V == 0) return HB_OK;
  if (E % 8 != 0 || V % 16 != 0 || ld_scores % 4 != 0 || ld_scores < V)
    return fail(HB_ERR_INVALID, "hb_similarity needs E %% 8 == 0, V %% 16 == 0 (pad the gallery), ld_scores %% 4 == 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Kx = exact ? 6 * E : E;
  // scratch is allocated on the stream (cudaMallocAsync): no hidden sync, freed in stream order
  __nv_bfloat16 *tb = nullptr, *vb = nullptr;
  HB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tb), static_cast<size_t>(Q) * Kx * 2, s));
  HB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&vb), static_cast<size_t>(V) * Kx * 2, s));
  int rc = HB_OK;
  do {
    if (exact) {
      int r1 = hb::split_bf16_launch(text, tb, Q, E, 0, s), r2 = hb::split_bf16_launch(video, vb, V, E, 1, s);
      if (r1 || r2) { rc = fail(HB_ERR_CUDA, "split_bf16 launch failed"); break; }
    } else {
      int r1 = hb::f32_to_bf16_launch(text, tb, Q * E, s), r2 = hb::f32_to_bf16_launch(video, vb, V * E, s);
      if (r1 || r2) { rc = fail(HB_ERR_CUDA, "f32_to_bf16 launch failed"); break; }
    }
    g_launches.fetch_add(2);
    CUtensorMap tmT, tmV;
    const int cg = (V % 32 == 0) ? g_cg : 1;
    if (hb::make_tmap_bf16(&tmT, tb, Q, Kx, Kx, hb::gemm_a_box_rows()) ||
        hb::make_tmap_bf16(&tmV, vb, V, Kx, Kx, hb::gemm_w_box_rows(cg))) {
      rc = fail(HB_ERR_CUDA, "cuTensorMapEncodeTiled failed for similarity operands");
      break;
    }
    hb::GemmParams p;
    p.M = static_cast<int>(Q); p.N = static_cast<int>(V); p.K = Kx; p.out = scores; p.ldo = static_cast<int>(ld_scores);
    int r = hb::gemm_launch(tmT, tmV, p, hb::EPI_F32, cg, g_num_sms, s);
    if (r) { rc = fail(HB_ERR_CUDA, "similarity GEMM launch failed: %d", r); break; }
    g_launches.fetch_add(1);
  } while (0);
  cudaFreeAsync(tb, s);
  cudaFreeAsync(vb, s);
  return rc;
}

int hb_linear(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, const float* resid, void* out,
              int64_t ldo, int64_t M, int64_t N, int64_t K, int epilogue, void* stream) {
  if (!g_inited) return fail(HB_ERR_INVALID, "hb_init() not called");
  if (!x || !w || !out) return fail(HB_ERR_INVALID, "null argument");
  if (M == 0) return HB_OK;
  if (K % 8 || N % 16 || ldx % 8 || ldw % 8) return fail(HB_ERR_INVALID, "hb_linear needs K %% 8 == 0, N %% 16 == 0, ld %% 8 == 0");
  if (epilogue < 0 || epilogue > 2) return fail(HB_ERR_INVALID, "bad epilogue %d", epilogue);
  if (resid && epilogue != HB_EPI_F32) return fail(HB_ERR_INVALID, "residual needs HB_EPI_F32");
  CUtensorMap tmA, tmW;
  const int cg = (N % 32 == 0) ? g_cg : 1;
  if (hb::make_tmap_bf16(&tmA, x, M, K, ldx, hb::gemm_a_box_rows()) || hb::make_tmap_bf16(&tmW, w, N, K, ldw, hb::gemm_w_box_rows(cg)))
    return fail(HB_ERR_INVALID, "cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned)");
  hb::GemmParams p;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.bias = bias; p.out = out; p.ldo = static_cast<int>(ldo); p.resid = resid;
  HB_LAUNCH(hb::gemm_launch(tmA, tmW, p, epilogue, cg, g_num_sms, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

int hb_layernorm(const float* x, const float* w, const float* b, float eps, int64_t rows, int D, void* y, int out_bf16,
                 void* stream) {
  if (!x || !w || !b || !y) return fail(HB_ERR_INVALID, "null argument");
  hb::LayerNormParams ln;
  ln.x = x; ln.ldx = D; ln.y = y; ln.ldy = D; ln.w = w; ln.b = b; ln.eps = eps; ln.rows = static_cast<int>(rows); ln.D = D;
  HB_LAUNCH(hb::layernorm_launch(ln, out_bf16 != 0, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

int hb_vit_attention(const void* qkv, void* out, int64_t B, int H, void* stream) {
  if (!qkv || !out) return fail(HB_ERR_INVALID, "null argument");
  if (B == 0) return HB_OK;
  hb::AttnParams ap;
  ap.qkv = static_cast<const __nv_bfloat16*>(qkv); ap.out = static_cast<__nv_bfloat16*>(out);
  ap.B = static_cast<int>(B); ap.H = H;
  HB_LAUNCH(vit_attn_dispatch(ap, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

int hb_small_attention_f32(const float* q, const float* k, const float* v, float* out, int B, int H, int Tq, int Tk, int ldq, int ldk, int ldv,
                           int ldo, int64_t bsq, int64_t bsk, int64_t bsv, int64_t bso, float scale, int mask_mode, float mask_const,
                           int causal_soft, int use_tensor_cores, void* stream) {
  if (!g_inited) return fail(HB_ERR_INVALID, "hb_init() not called");
  if (!q || !k || !v || !out) return fail(HB_ERR_INVALID, "null argument");
  hb::SmallAttnF32Params ap;
  ap.q = q; ap.k = k; ap.v = v; ap.out = out; ap.B = B; ap.H = H; ap.Tq = Tq; ap.Tk = Tk;
  ap.ldq = ldq; ap.ldk = ldk; ap.ldv = ldv; ap.ldo = ldo; ap.bsq = bsq; ap.bsk = bsk; ap.bsv = bsv; ap.bso = bso;
  ap.scale = scale; ap.mask_mode = mask_mode; ap.mask_const = mask_const; ap.causal_soft = causal_soft;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!use_tensor_cores) {
    HB_LAUNCH(hb::small_attn_f32_launch(ap, s));
    return HB_OK;
  }
  const size_t bytes = hb::small_attn_tc_workspace(B, H, Tq, Tk);
  void* ws = nullptr;
  HB_CUDA(cudaMallocAsync(&ws, bytes, s));
  const int r = hb::small_attn_tc_launch(ap, ws, bytes, s);
  cudaFreeAsync(ws, s);
  if (r != 0) return fail(r > 0 ? HB_ERR_CUDA : HB_ERR_INVALID, "small_attn_tc_launch failed: %d", r);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return HB_OK;
}

int hb_small_attention(const void* q, const void* k, const void* v, void* out, int B, int H, int Tq, int Tk, int ldq, int ldk,
                       int ldv, int ldo, int64_t bsq, int64_t bsk, int64_t bsv, int64_t bso, float scale, int mask_mode,
                       float mask_const, int causal_soft, void* stream) {
  if (!q || !k || !v || !out) return fail(HB_ERR_INVALID, "null argument");
  hb::SmallAttnParams ap;
  ap.q = static_cast<const __nv_bfloat16*>(q); ap.k = static_cast<const __nv_bfloat16*>(k);
  ap.v = static_cast<const __nv_bfloat16*>(v); ap.out = static_cast<__nv_bfloat16*>(out);
  ap.B = B; ap.H = H; ap.Tq = Tq; ap.Tk = Tk; ap.ldq = ldq; ap.ldk = ldk; ap.ldv = ldv; ap.ldo = ldo;
  ap.bsq = bsq; ap.bsk = bsk; ap.bsv = bsv; ap.bso = bso; ap.scale = scale; ap.mask_mode = mask_mode;
  ap.mask_const = mask_const; ap.causal_soft = causal_soft;
  HB_LAUNCH(hb::small_attn_launch(ap, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

}  // extern "C"

// =================================================================================================
// MomentModel shared encoder + heads
// =================================================================================================
struct HbMoment {
  HbMomentConfig cfg;
  long long max_rows = 0;
  int max_batch = 0;
  F32Vec asr_ln_w, asr_ln_b, temp_w1, temp_b1, memb, bemb, head_w, head_b, vn_w, vn_b, pos, emb_ln_w, emb_ln_b;
  SplitLinear asr, temp2, gmap, gmap_text, emb;
  struct Layer {
    SplitLinear qkv, ao, inter, out;
    F32Vec ao_ln_w, ao_ln_b, o_ln_w, o_ln_b;
  };
  std::vector<std::unique_ptr<Layer>> layers;
  DevBuf op, opT;                 // bf16 split operands [max_rows, 3*ffn], [max_batch, 3*clip_dim]
  DevBuf vlin, asr_ln, asr_lin, tanh_in, temporal, base, f, tlin, that, e_lin, x, qkv, att, t1, h, mid, attn_ws;
  // A-operand maps over `op` for each K in use
  CUtensorMap tm_clip, tm_asr, tm_e, tm_hd, tm_ffn, tm_text;
};

namespace {

int split_gemm(HbMoment* m, const float* act, long long rows, int K, int gelu, const CUtensorMap& tmA, const SplitLinear& L,
               float* out, int epi, cudaStream_t s, const float* resid = nullptr, const float* rowadd = nullptr, int remap = 0,
               __nv_bfloat16* opbuf = nullptr) {
  return split_gemm_op(opbuf ? opbuf : m->op.as<__nv_bfloat16>(), act, rows, K, gelu, tmA, L, out, epi, s, resid, rowadd, remap);
}

}  // namespace

extern "C" {

int hb_moment_create(const HbMomentConfig* cfg, const HbMomentWeights* w, int64_t max_rows, int max_batch, void* stream,
                     HbMoment** out) {
  if (!g_inited) return fail(HB_ERR_INVALID, "hb_init() not called");
  if (!cfg || !w || !out || max_rows <= 0 || max_batch <= 0) return fail(HB_ERR_INVALID, "null argument");
  const int E = cfg->embed_dim, Hd = cfg->hidden, Ff = cfg->ffn, A = cfg->asr_dim, Cd = cfg->clip_dim;
  if (Hd != cfg->heads * 64) return fail(HB_ERR_INVALID, "head_dim must be 64");
  if (E % 32 || Hd % 32 || Ff % 32 || A < 0 || A % 8 || Cd % 8 || E > 1024) return fail(HB_ERR_INVALID, "unsupported MomentModel dims");
  const bool use_asr = A > 0;   // modeling.py:28-35: asr_dim <= 0 builds no asr_enc_layer
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::unique_ptr<HbMoment> m(new (std::nothrow) HbMoment);
  if (!m) return fail(HB_ERR_NOMEM, "host allocation failed");
  m->cfg = *cfg; m->max_rows = max_rows; m->max_batch = max_batch;
  int r;
#define INITV(dst, src, n) if ((r = m->dst.init(w->src, static_cast<size_t>(n), s))) return r
  if (use_asr) { INITV(asr_ln_w, asr_ln_w, A); INITV(asr_ln_b, asr_ln_b, A); }
  INITV(temp_w1, temp_w1, E); INITV(temp_b1, temp_b1, E);
  INITV(memb, mask_embed, 2 * E); INITV(bemb, boundary_embed, 2 * E); INITV(head_w, head_w, 3 * Hd); INITV(head_b, head_b, 3);
  INITV(vn_w, vis_norm_w, E); INITV(vn_b, vis_norm_b, E); INITV(pos, pos_emb, static_cast<size_t>(cfg->max_pos) * Hd);
  INITV(emb_ln_w, emb_ln_w, Hd); INITV(emb_ln_b, emb_ln_b, Hd);
#undef INITV
  if (use_asr && (r = m->asr.init_from(w->asr_w, w->asr_b, E, A, s))) return r;
  if ((r = m->temp2.init_from(w->temp_w2, w->temp_b2, E, E, s))) return r;
  if ((r = m->gmap.init_from(w->clip_g_map_w, w->clip_g_map_b, E, Cd, s))) return r;
  if ((r = m->gmap_text.init_from(w->clip_g_map_text_w, w->clip_g_map_text_b, E, Cd, s))) return r;
  if ((r = m->emb.init_from(w->emb_w, w->emb_b, Hd, E, s))) return r;
  DevBuf tmpw, tmpb;
  if ((r = tmpw.alloc(static_cast<size_t>(3) * Hd * Hd * 4))) return r;
  if ((r = tmpb.alloc(static_cast<size_t>(3) * Hd * 4))) return r;
  for (int i = 0; i < cfg->layers; ++i) {
    std::unique_ptr<HbMoment::Layer> L(new HbMoment::Layer);
    const float* ws[3] = {w->q_w[i], w->k_w[i], w->v_w[i]};
    const float* bs[3] = {w->q_b[i], w->k_b[i], w->v_b[i]};
    for (int j = 0; j < 3; ++j) {  // separate query / key / value Linears fused into one [3*Hd, Hd] GEMM
      HB_CUDA(cudaMemcpyAsync(tmpw.as<float>() + static_cast<size_t>(j) * Hd * Hd, ws[j], static_cast<size_t>(Hd) * Hd * 4, cudaMemcpyDeviceToDevice, s));
      HB_CUDA(cudaMemcpyAsync(tmpb.as<float>() + static_cast<size_t>(j) * Hd, bs[j], static_cast<size_t>(Hd) * 4, cudaMemcpyDeviceToDevice, s));
    }
    if ((r = L->qkv.init_from(tmpw.as<float>(), tmpb.as<float>(), 3 * Hd, Hd, s))) return r;
    if ((r = L->ao.init_from(w->ao_w[i], w->ao_b[i], Hd, Hd, s))) return r;
    if ((r = L->inter.init_from(w->i_w[i], w->i_b[i], Ff, Hd, s))) return r;
    if ((r = L->out.init_from(w->o_w[i], w->o_b[i], Hd, Ff, s))) return r;
    if ((r = L->ao_ln_w.init(w->ao_ln_w[i], Hd, s))) return r;
    if ((r = L->ao_ln_b.init(w->ao_ln_b[i], Hd, s))) return r;
    if ((r = L->o_ln_w.init(w->o_ln_w[i], Hd, s))) return r;
    if ((r = L->o_ln_b.init(w->o_ln_b[i], Hd, s))) return r;
    m->layers.push_back(std::move(L));
  }
  const size_t R = static_cast<size_t>(max_rows);
  const int Kmax = std::max(std::max(Ff, Cd), std::max(Hd, E));
  if ((r = m->op.alloc(R * 3 * Kmax * 2))) return r;
  if ((r = m->opT.alloc(static_cast<size_t>(max_batch) * 3 * Cd * 2))) return r;
#define ALLOCF(buf, cols) if ((r = m->buf.alloc(R * static_cast<size_t>(cols) * 4))) return r
  ALLOCF(vlin, E); if (use_asr) { ALLOCF(asr_ln, A); ALLOCF(asr_lin, E); } ALLOCF(tanh_in, E); ALLOCF(temporal, E); ALLOCF(base, E); ALLOCF(f, E);
  ALLOCF(e_lin, Hd); ALLOCF(x, Hd); ALLOCF(qkv, 3 * Hd); ALLOCF(att, Hd); ALLOCF(t1, Hd); ALLOCF(h, Hd); ALLOCF(mid, Ff);
#undef ALLOCF
  // tensor-core attention workspace: bf16 split copies of q / k / v for max_rows tokens (any B x T with B * T <= max_rows)
  if ((r = m->attn_ws.alloc(static_cast<size_t>(max_rows) * cfg->heads * (192 + 192 + 128) * 2 + 4096))) return r;
  if ((r = m->tlin.alloc(static_cast<size_t>(max_batch) * E * 4))) return r;
  if ((r = m->that.alloc(static_cast<size_t>(max_batch) * E * 4))) return r;
  auto amap = [&](CUtensorMap* tm, void* ptr, long long rows, int K) {
    return hb::make_tmap_bf16(tm, ptr, rows, 3 * K, 3 * K, hb::gemm_a_box_rows());
  };
  if (amap(&m->tm_clip, m->op.p, max_rows, Cd) || (use_asr && amap(&m->tm_asr, m->op.p, max_rows, A)) || amap(&m->tm_e, m->op.p, max_rows, E) ||
      amap(&m->tm_hd, m->op.p, max_rows, Hd) || amap(&m->tm_ffn, m->op.p, max_rows, Ff) || amap(&m->tm_text, m->opT.p, max_batch, Cd))
    return fail(HB_ERR_CUDA, "tensor map (moment operands) failed");
  HB_CUDA(cudaStreamSynchronize(s));  // tmpw/tmpb are released on return
  *out = m.release();
  return HB_OK;
}

void hb_moment_destroy(HbMoment* m) { delete m; }

int hb_moment_forward(HbMoment* m, const float* video, const float* text_feat, const float* asr, const int64_t* video_mask,
                      const int64_t* moment_mask, const int64_t* boundary_mask, int B, int T, int flags, float* out_feats,
                      float* out_logits, void* stream) {
  if (!m || !moment_mask || !out_logits) return fail(HB_ERR_INVALID, "null argument");
  if (B <= 0 || T <= 0) return fail(HB_ERR_INVALID, "empty batch");
  const long long R = static_cast<long long>(B) * T;
  const HbMomentConfig& c = m->cfg;
  if (R > m->max_rows || B > m->max_batch) return fail(HB_ERR_INVALID, "batch of %d x %d frames exceeds the handle's capacity", B, T);
  if (T > c.max_pos) return fail(HB_ERR_INVALID, "T = %d exceeds max_position_embeddings %d", T, c.max_pos);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int E = c.embed_dim, Hd = c.hidden, Ff = c.ffn, A = c.asr_dim, Cd = c.clip_dim;
  int r;
  if (!(flags & HB_MOMENT_REUSE_BASE)) {
    if (!video || !text_feat || (A > 0 && !asr) || !video_mask) return fail(HB_ERR_INVALID, "null argument");
    // clip_g_map (modeling.py:158), clip_g_map_text + L2 norm (:159,163)
    if ((r = split_gemm(m, video, R, Cd, 0, m->tm_clip, m->gmap, m->vlin.as<float>(), hb::EPI_F32, s))) return r;
    if ((r = split_gemm(m, text_feat, B, Cd, 0, m->tm_text, m->gmap_text, m->tlin.as<float>(), hb::EPI_F32, s, nullptr, nullptr, 0,
                        m->opT.as<__nv_bfloat16>())))
      return r;
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::pool_normalize_launch(m->tlin.as<float>(), m->that.p, B, 1, E, true, false, s));
    // asr_enc_layer = LayerNorm(1e-5) -> Linear (:167-169)
    if (A > 0) {
      if ((r = ln_f32(asr, m->asr_ln.as<float>(), m->asr_ln_w, m->asr_ln_b, 1e-5f, R, A, s))) return r;
      if ((r = split_gemm(m, m->asr_ln.as<float>(), R, A, 0, m->tm_asr, m->asr, m->asr_lin.as<float>(), hb::EPI_F32, s))) return r;
    }
    // temporal_embed = Linear(1,E) -> tanh -> Linear(E,E) on the per-sample time grid (:178-196)
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::time_tanh_launch(reinterpret_cast<const long long*>(video_mask), m->temp_w1.ptr(),
                                                        m->temp_b1.ptr(), m->tanh_in.as<float>(), B, T, E, s));
    if ((r = split_gemm(m, m->tanh_in.as<float>(), R, E, 0, m->tm_e, m->temp2, m->temporal.as<float>(), hb::EPI_F32, s))) return r;
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::moment_base_launch(m->vlin.as<float>(), m->vn_w.ptr(), m->vn_b.ptr(), m->that.as<float>(),
                                                          A > 0 ? m->asr_lin.as<float>() : nullptr, m->temporal.as<float>(), m->base.as<float>(), B, T, E, s));
  }
  HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::moment_embed_launch(m->base.as<float>(), m->bemb.ptr(), m->memb.ptr(),
                                                         reinterpret_cast<const long long*>(boundary_mask),
                                                         reinterpret_cast<const long long*>(moment_mask), m->f.as<float>(), R, E, s));
  // VisualEmbeddings: Linear(E,Hd) + position embedding + LN (module_visual.py:118-130)
  if ((r = split_gemm(m, m->f.as<float>(), R, E, 0, m->tm_e, m->emb, m->e_lin.as<float>(), hb::EPI_F32_ROWADD, s, nullptr, m->pos.ptr(), T)))
    return r;
  float* x = m->x.as<float>();
  const bool fuse = ln_split_ok(Hd);   // every LayerNorm below feeds the next linear: it also writes that GEMM's split operand
  __nv_bfloat16* opb = m->op.as<__nv_bfloat16>();
  if (fuse) { if ((r = ln_split(m->e_lin.as<float>(), x, opb, m->emb_ln_w, m->emb_ln_b, 1e-12f, R, Hd, s))) return r; }
  else if ((r = ln_f32(m->e_lin.as<float>(), x, m->emb_ln_w, m->emb_ln_b, 1e-12f, R, Hd, s))) return r;
  for (size_t li = 0; li < m->layers.size(); ++li) {
    HbMoment::Layer& L = *m->layers[li];
    const bool last = (li + 1 == m->layers.size());
    if ((r = split_gemm(m, fuse ? nullptr : x, R, Hd, 0, m->tm_hd, L.qkv, m->qkv.as<float>(), hb::EPI_F32, s))) return r;
    hb::SmallAttnF32Params ap;
    ap.q = m->qkv.as<float>(); ap.k = ap.q + Hd; ap.v = ap.q + 2 * Hd; ap.out = m->att.as<float>();
    ap.B = B; ap.H = c.heads; ap.Tq = T; ap.Tk = T;
    ap.ldq = ap.ldk = ap.ldv = 3 * Hd; ap.ldo = Hd;
    ap.bsq = ap.bsk = ap.bsv = static_cast<long long>(T) * 3 * Hd; ap.bso = static_cast<long long>(T) * Hd;
    ap.scale = 0.125f; ap.mask_mode = 2; ap.mask_const = -10000.0f;  // all-zeros mask quirk, modeling.py:208
    if ((r = small_attn_f32_auto(ap, m->attn_ws, s))) return r;
    if ((r = split_gemm(m, m->att.as<float>(), R, Hd, 0, m->tm_hd, L.ao, m->t1.as<float>(), hb::EPI_F32, s, x))) return r;
    if (fuse) { if ((r = ln_split(m->t1.as<float>(), m->h.as<float>(), opb, L.ao_ln_w, L.ao_ln_b, 1e-12f, R, Hd, s))) return r; }
    else if ((r = ln_f32(m->t1.as<float>(), m->h.as<float>(), L.ao_ln_w, L.ao_ln_b, 1e-12f, R, Hd, s))) return r;
    if ((r = split_gemm(m, fuse ? nullptr : m->h.as<float>(), R, Hd, 0, m->tm_hd, L.inter, m->mid.as<float>(), hb::EPI_F32, s))) return r;
    if ((r = split_gemm(m, m->mid.as<float>(), R, Ff, /*gelu=*/1, m->tm_ffn, L.out, m->t1.as<float>(), hb::EPI_F32, s, m->h.as<float>()))) return r;
    if (fuse) { if ((r = ln_split(m->t1.as<float>(), x, last ? nullptr : opb, L.o_ln_w, L.o_ln_b, 1e-12f, R, Hd, s))) return r; }
    else if ((r = ln_f32(m->t1.as<float>(), x, L.o_ln_w, L.o_ln_b, 1e-12f, R, Hd, s))) return r;
  }
  HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::moment_heads_launch(x, m->head_w.ptr(), m->head_b.ptr(), out_logits, R, Hd, s));
  if (out_feats) HB_CUDA(cudaMemcpyAsync(out_feats, x, static_cast<size_t>(R) * Hd * 4, cudaMemcpyDeviceToDevice, s));
  return HB_OK;
}

int hb_moment_mr_decode(const float* logits, const int64_t* video_mask, int64_t* pred, int B, int T, void* stream) {
  if (!logits || !video_mask || !pred) return fail(HB_ERR_INVALID, "null argument");
  HB_LAUNCH(hb::mr_argmax_launch(logits, reinterpret_cast<const long long*>(video_mask), reinterpret_cast<long long*>(pred), B, T,
                                 static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

int hb_moment_ms_step(const float* logits, int64_t* moment_mask, int64_t* boundary_mask, int32_t* steps, int32_t* nsteps, int max_steps,
                      int B, int T, double threshold, float* probs_out, void* stream) {
  if (!logits || !moment_mask || !boundary_mask || !steps || !nsteps) return fail(HB_ERR_INVALID, "null argument");
  HB_LAUNCH(hb::ms_step_launch(logits, reinterpret_cast<long long*>(moment_mask), reinterpret_cast<long long*>(boundary_mask), steps,
                               nsteps, max_steps, B, T, threshold, probs_out, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

int hb_trim_feats(const float* x, const int64_t* mask, float* out, int B, int T, int C, int F, void* stream) {
  if (!x || !mask || !out) return fail(HB_ERR_INVALID, "null argument");
  HB_LAUNCH(hb::trim_feats_launch(x, reinterpret_cast<const long long*>(mask), out, B, T, C, F, static_cast<cudaStream_t>(stream)));
  return HB_OK;
}

}  // extern "C"

// =================================================================================================
// Caption decoder + beam search
// =================================================================================================
struct HbDecoder {
  HbDecoderConfig cfg;
  int max_inst = 0, max_beam = 0, max_enc = 0, Vpad = 0;
  int n_inst = 0, beam = 0, enc_len = 0, step = 0;
  F32Vec word_emb, pos_emb, emb_ln_w, emb_ln_b, cls_ln_w, cls_ln_b;
  SplitLinear cls_dense, cls_vocab;
  struct Layer {
    SplitLinear sqkv, so, eq, ekv, eo, inter, out;
    F32Vec so_ln_w, so_ln_b, eo_ln_w, eo_ln_b, o_ln_w, o_ln_b;
    DevBuf kc[2], vc[2];   // self-attention caches [R, max_words, Hd], ping-pong for the beam re-order
    DevBuf ekv_buf;        // cross keys | values [n_inst * enc_len, 2 * Hd]
  };
  std::vector<std::unique_ptr<Layer>> layers;
  int cur = 0;             // which ping-pong cache holds the live data
  DevBuf op, x, qkv, att, t1, s1, qc, c1, mid, th, logits;
  DevBuf tok, scores, done, nsteps, prev_k, ys, cand_v, cand_i;
  static constexpr int MAX_SPLITS = 16;
  DevBuf skws;             // split-K partial sums [MAX_SPLITS, max rows, hidden] fp32
  DevBuf kv_idx[2];        // beam re-ordering as an index table int32 [R, max_words] (ping-pong), shared by all layers
  CUtensorMap tm_hd, tm_ffn, tm_enc;
  // CUDA graphs of the decode steps, one set per (n_inst, beam, enc_len) shape: a job's full batches and its ragged last batch
  // alternate, so the sets of the MAX_SHAPES most recently used shapes are kept
  struct GraphSet {
    std::vector<cudaGraphExec_t> exec;     // per step index
    std::vector<int64_t> launches;         // kernels per step graph (hb_launch_count stays the number of kernels run)
    int searches = 0;                      // searches begun with this shape before the current one
    uint64_t last_use = 0;
    void destroy() { for (auto& e : exec) if (e) { cudaGraphExecDestroy(e); e = nullptr; } }
  };
  static constexpr size_t MAX_SHAPES = 8;
  std::map<std::array<int, 3>, GraphSet> graph_sets;
  GraphSet* gs = nullptr;                  // the current search's set
  uint64_t use_clock = 0;
  cudaStream_t cap_stream = nullptr;     // capture happens here: the caller's stream may be the legacy default stream, which cannot capture
  ~HbDecoder() {
    for (auto& kv : graph_sets) kv.second.destroy();
    if (cap_stream) cudaStreamDestroy(cap_stream);
  }
};

extern "C" {

int hb_decoder_create(const HbDecoderConfig* cfg, const HbDecoderWeights* w, int max_inst, int max_beam, int max_enc_len, void* stream,
                      HbDecoder** out) {
  if (!g_inited) return fail(HB_ERR_INVALID, "hb_init() not called");
  if (!cfg || !w || !out || max_inst <= 0 || max_beam <= 0 || max_beam > 8 || max_enc_len <= 0) return fail(HB_ERR_INVALID, "bad argument");
  const int Hd = cfg->hidden, Ff = cfg->ffn, V = cfg->vocab;
  if (Hd != cfg->heads * 64 || Hd % 32 || Ff % 32 || cfg->max_words > cfg->max_pos)
    return fail(HB_ERR_INVALID, "unsupported decoder dims");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::unique_ptr<HbDecoder> d(new (std::nothrow) HbDecoder);
  if (!d) return fail(HB_ERR_NOMEM, "host allocation failed");
  d->cfg = *cfg; d->max_inst = max_inst; d->max_beam = max_beam; d->max_enc = max_enc_len;
  d->Vpad = (V + 31) / 32 * 32;
  const size_t R = static_cast<size_t>(max_inst) * max_beam;
  int r;
  if ((r = d->word_emb.init(w->word_emb, static_cast<size_t>(V) * Hd, s))) return r;
  if ((r = d->pos_emb.init(w->pos_emb, static_cast<size_t>(cfg->max_pos) * Hd, s))) return r;
  if ((r = d->emb_ln_w.init(w->emb_ln_w, Hd, s))) return r;
  if ((r = d->emb_ln_b.init(w->emb_ln_b, Hd, s))) return r;
  if ((r = d->cls_ln_w.init(w->cls_ln_w, Hd, s))) return r;
  if ((r = d->cls_ln_b.init(w->cls_ln_b, Hd, s))) return r;
  if ((r = d->cls_dense.init_from(w->cls_dense_w, w->cls_dense_b, Hd, Hd, s))) return r;
  {  // tied classifier: word embedding [V, Hd] padded with zero rows to Vpad; bias padded likewise
    DevBuf tw, tb;
    if ((r = tw.alloc(static_cast<size_t>(d->Vpad) * Hd * 4))) return r;
    if ((r = tb.alloc(static_cast<size_t>(d->Vpad) * 4))) return r;
    HB_CUDA(cudaMemsetAsync(tw.p, 0, tw.bytes, s));
    HB_CUDA(cudaMemsetAsync(tb.p, 0, tb.bytes, s));
    HB_CUDA(cudaMemcpyAsync(tw.p, w->word_emb, static_cast<size_t>(V) * Hd * 4, cudaMemcpyDeviceToDevice, s));
    HB_CUDA(cudaMemcpyAsync(tb.p, w->cls_bias, static_cast<size_t>(V) * 4, cudaMemcpyDeviceToDevice, s));
    if ((r = d->cls_vocab.init_from(tw.as<float>(), tb.as<float>(), d->Vpad, Hd, s))) return r;
    HB_CUDA(cudaStreamSynchronize(s));
  }
  DevBuf tmpw, tmpb;
  if ((r = tmpw.alloc(static_cast<size_t>(3) * Hd * Hd * 4))) return r;
  if ((r = tmpb.alloc(static_cast<size_t>(3) * Hd * 4))) return r;
  auto fuse = [&](const float* const* ws, const float* const* bs, int n) -> int {
    for (int j = 0; j < n; ++j) {
      HB_CUDA(cudaMemcpyAsync(tmpw.as<float>() + static_cast<size_t>(j) * Hd * Hd, ws[j], static_cast<size_t>(Hd) * Hd * 4, cudaMemcpyDeviceToDevice, s));
      HB_CUDA(cudaMemcpyAsync(tmpb.as<float>() + static_cast<size_t>(j) * Hd, bs[j], static_cast<size_t>(Hd) * 4, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
  };
  for (int i = 0; i < cfg->layers; ++i) {
    std::unique_ptr<HbDecoder::Layer> L(new HbDecoder::Layer);
    const float* w3[3] = {w->sq_w[i], w->sk_w[i], w->sv_w[i]};
    const float* b3[3] = {w->sq_b[i], w->sk_b[i], w->sv_b[i]};
    if ((r = fuse(w3, b3, 3))) return r;
    if ((r = L->sqkv.init_from(tmpw.as<float>(), tmpb.as<float>(), 3 * Hd, Hd, s))) return r;
    const float* w2[2] = {w->ek_w[i], w->ev_w[i]};
    const float* b2[2] = {w->ek_b[i], w->ev_b[i]};
    if ((r = fuse(w2, b2, 2))) return r;
    if ((r = L->ekv.init_from(tmpw.as<float>(), tmpb.as<float>(), 2 * Hd, Hd, s))) return r;
    if ((r = L->so.init_from(w->so_w[i], w->so_b[i], Hd, Hd, s))) return r;
    if ((r = L->eq.init_from(w->eq_w[i], w->eq_b[i], Hd, Hd, s))) return r;
    if ((r = L->eo.init_from(w->eo_w[i], w->eo_b[i], Hd, Hd, s))) return r;
    if ((r = L->inter.init_from(w->i_w[i], w->i_b[i], Ff, Hd, s))) return r;
    if ((r = L->out.init_from(w->o_w[i], w->o_b[i], Hd, Ff, s))) return r;
    if ((r = L->so_ln_w.init(w->so_ln_w[i], Hd, s))) return r;
    if ((r = L->so_ln_b.init(w->so_ln_b[i], Hd, s))) return r;
    if ((r = L->eo_ln_w.init(w->eo_ln_w[i], Hd, s))) return r;
    if ((r = L->eo_ln_b.init(w->eo_ln_b[i], Hd, s))) return r;
    if ((r = L->o_ln_w.init(w->o_ln_w[i], Hd, s))) return r;
    if ((r = L->o_ln_b.init(w->o_ln_b[i], Hd, s))) return r;
    for (int j = 0; j < 2; ++j) {
      if ((r = L->kc[j].alloc(R * cfg->max_words * Hd * 4))) return r;
      if ((r = L->vc[j].alloc(R * cfg->max_words * Hd * 4))) return r;
    }
    if ((r = L->ekv_buf.alloc(static_cast<size_t>(max_inst) * max_enc_len * 2 * Hd * 4))) return r;
    d->layers.push_back(std::move(L));
  }
  const size_t Rop = std::max(R, static_cast<size_t>(max_inst) * max_enc_len);
  if ((r = d->op.alloc(Rop * 3 * Ff * 2))) return r;
#define ALLOCD(buf, cols) if ((r = d->buf.alloc(R * static_cast<size_t>(cols) * 4))) return r
  ALLOCD(x, Hd); ALLOCD(qkv, 3 * Hd); ALLOCD(att, Hd); ALLOCD(t1, Hd); ALLOCD(s1, Hd); ALLOCD(qc, Hd); ALLOCD(c1, Hd); ALLOCD(mid, Ff);
  ALLOCD(th, Hd); ALLOCD(logits, d->Vpad);
#undef ALLOCD
  if ((r = d->tok.alloc(R * 8))) return r;
  if ((r = d->scores.alloc(R * 4))) return r;
  if ((r = d->cand_v.alloc(R * max_beam * 4))) return r;
  if ((r = d->cand_i.alloc(R * max_beam * 4))) return r;
  if ((r = d->skws.alloc(static_cast<size_t>(HbDecoder::MAX_SPLITS) * R * Hd * 4))) return r;
  for (int i = 0; i < 2; ++i)
    if ((r = d->kv_idx[i].alloc(R * static_cast<size_t>(cfg->max_words) * 4))) return r;
  if ((r = d->done.alloc(static_cast<size_t>(max_inst) * 4))) return r;
  if ((r = d->nsteps.alloc(static_cast<size_t>(max_inst) * 4))) return r;
  if ((r = d->prev_k.alloc(static_cast<size_t>(cfg->max_words) * R * 4))) return r;
  if ((r = d->ys.alloc(static_cast<size_t>(cfg->max_words) * R * 4))) return r;
  if (hb::make_tmap_bf16(&d->tm_hd, d->op.p, R, 3 * Hd, 3 * Hd, hb::gemm_a_box_rows()) ||
      hb::make_tmap_bf16(&d->tm_ffn, d->op.p, R, 3 * Ff, 3 * Ff, hb::gemm_a_box_rows()) ||
      hb::make_tmap_bf16(&d->tm_enc, d->op.p, static_cast<uint64_t>(max_inst) * max_enc_len, 3 * Hd, 3 * Hd, hb::gemm_a_box_rows()))
    return fail(HB_ERR_CUDA, "tensor map (decoder operands) failed");
  HB_CUDA(cudaStreamSynchronize(s));
  *out = d.release();
  return HB_OK;
}

void hb_decoder_destroy(HbDecoder* d) { delete d; }

}  // extern "C"

namespace {

// act [rows, K] fp32 -> split operand -> GEMM with SplitLinear -> out fp32 [rows, N] (+ fp32 residual)
// GEMM of a decoder linear whose split operand already sits in d->op
int dec_gemm_raw(HbDecoder* d, long long rows, int K, const CUtensorMap& tmA, const SplitLinear& L, float* out, cudaStream_t s,
                 const float* resid = nullptr) {
  (void)d;
  hb::GemmParams p;
  p.M = static_cast<int>(rows); p.N = L.N; p.K = 3 * K;
  {  // wide outputs (the vocabulary projection): whole rounds of N tiles over the workers; depends on N only
    const int workers = g_num_sms / L.cg, min_tiles = (L.N + 255) / 256;
    if (min_tiles > workers) { p.n_tiles = (min_tiles + workers - 1) / workers * workers; p.m_fastest = 1; }
  }
  p.bias = L.has_bias ? L.b.as<float>() : nullptr; p.out = out; p.ldo = L.N; p.resid = resid;
  HB_LAUNCH_P(CAT_GEMM_F32, 2.0 * rows * L.N * 3.0 * K, s, hb::gemm_launch(tmA, L.tm, p, hb::EPI_F32, L.cg, g_num_sms, s));
  return 0;
}

int dec_gemm(HbDecoder* d, const float* act, long long rows, int K, int gelu, const CUtensorMap& tmA, const SplitLinear& L, float* out,
             cudaStream_t s, const float* resid = nullptr) {
  HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::split3_act_launch(act, d->op.as<__nv_bfloat16>(), rows, K, gelu, s));
  return dec_gemm_raw(d, rows, K, tmA, L, out, s, resid);
}

// Split-K form for the decode steps' narrow linears (N = hidden: 3 column tiles, so 3 CTA pairs streamed the whole K = 2304 /
// 9216 operand at one SM's L2 bandwidth — 20 / 51 us per GEMM at R = 192 rows).  The slice count depends on K only, never on
// the number of rows, so a beam's arithmetic is the same whatever else shares the batch.  Operand in d->op; partial sums go
// to d->skws and are reduced by dec_finish.
int dec_split_count(int K) {
  if (g_dec_split_kbs <= 0) return 1;
  const int k_blocks = (3 * K + 63) / 64;
  int S = k_blocks / g_dec_split_kbs;
  S = S < 1 ? 1 : (S > HbDecoder::MAX_SPLITS ? HbDecoder::MAX_SPLITS : S);
  const int kbs = (k_blocks + S - 1) / S;
  return (k_blocks + kbs - 1) / kbs;   // every slice non-empty
}
int dec_gemm_split(HbDecoder* d, long long rows, int K, const CUtensorMap& tmA, const SplitLinear& L, int S, cudaStream_t s) {
  hb::GemmParams p;
  p.M = static_cast<int>(rows); p.N = L.N; p.K = 3 * K;
  p.out = d->skws.p; p.ldo = L.N; p.k_splits = S; p.split_stride = rows * L.N;
  HB_LAUNCH_P(CAT_GEMM_F32, 2.0 * rows * L.N * 3.0 * K, s, hb::gemm_launch(tmA, L.tm, p, hb::EPI_F32, L.cg, g_num_sms, s));
  return 0;
}
int dec_finish(HbDecoder* d, long long rows, const SplitLinear& L, int S, const float* resid, int gelu, const F32Vec* lnw, const F32Vec* lnb,
               float* y, bool emit_op, cudaStream_t s) {
  HB_LAUNCH_P(CAT_LAYERNORM, 0.0, s,
              hb::splitk_finish_launch(d->skws.as<float>(), rows * L.N, S, L.has_bias ? L.b.as<float>() : nullptr, resid,
                                       lnw ? lnw->ptr() : nullptr, lnb ? lnb->ptr() : nullptr, 1e-12f, gelu, y,
                                       emit_op ? d->op.as<__nv_bfloat16>() : nullptr, static_cast<int>(rows), L.N, s));
  return 0;
}

}  // namespace

extern "C" {

int hb_decoder_begin(HbDecoder* d, const float* enc, int n_inst, int enc_len, int beam, void* stream) {
  if (!d || !enc) return fail(HB_ERR_INVALID, "null argument");
  if (n_inst <= 0 || n_inst > d->max_inst || beam <= 0 || beam > d->max_beam || enc_len <= 0 || enc_len > d->max_enc)
    return fail(HB_ERR_INVALID, "decoder batch (%d instances, beam %d, %d frames) exceeds the handle's capacity", n_inst, beam, enc_len);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    const std::array<int, 3> key{n_inst, beam, enc_len};
    auto it = d->graph_sets.find(key);
    if (it == d->graph_sets.end()) {
      if (d->graph_sets.size() >= HbDecoder::MAX_SHAPES) {   // evict the least recently used shape
        auto lru = d->graph_sets.begin();
        for (auto j = d->graph_sets.begin(); j != d->graph_sets.end(); ++j)
          if (j->second.last_use < lru->second.last_use) lru = j;
        lru->second.destroy();
        d->graph_sets.erase(lru);
      }
      it = d->graph_sets.emplace(key, HbDecoder::GraphSet()).first;
    } else {
      it->second.searches += 1;
    }
    it->second.last_use = ++d->use_clock;
    d->gs = &it->second;   // std::map nodes are stable
  }
  d->n_inst = n_inst; d->beam = beam; d->enc_len = enc_len; d->step = 0; d->cur = 0;
  const int R = n_inst * beam, Hd = d->cfg.hidden;
  int r;
  for (auto& L : d->layers)  // cross-attention keys / values: once per search instead of once per step
    if ((r = dec_gemm(d, enc, static_cast<long long>(n_inst) * enc_len, Hd, 0, d->tm_enc, L->ekv, L->ekv_buf.as<float>(), s))) return r;
  std::vector<long long> tok(R, d->cfg.bos);
  HB_CUDA(cudaMemcpyAsync(d->tok.p, tok.data(), tok.size() * 8, cudaMemcpyHostToDevice, s));
  HB_CUDA(cudaMemsetAsync(d->scores.p, 0, static_cast<size_t>(R) * 4, s));
  HB_CUDA(cudaMemsetAsync(d->done.p, 0, static_cast<size_t>(n_inst) * 4, s));
  HB_CUDA(cudaMemsetAsync(d->nsteps.p, 0, static_cast<size_t>(n_inst) * 4, s));
  HB_CUDA(cudaMemsetAsync(d->prev_k.p, 0, d->prev_k.bytes, s));
  HB_CUDA(cudaMemsetAsync(d->ys.p, 0, d->ys.bytes, s));
  HB_CUDA(cudaStreamSynchronize(s));  // `tok` is a host temporary
  return HB_OK;
}

}  // extern "C"

// The ~50 launches of one decode step (everything below reads / writes handle-owned buffers only, so a step is a pure function
// of (step index, n_inst, beam, enc_len) and can be replayed as a CUDA graph).
static int decoder_step_launches(HbDecoder* d, cudaStream_t s) {
  const HbDecoderConfig& c = d->cfg;
  const int R = d->n_inst * d->beam, Hd = c.hidden, Ff = c.ffn, pos = d->step, Tmax = c.max_words;
  float* x = d->x.as<float>();
  int r;
  // sk: the hidden-width linears run as split-K GEMM + finish (bias, residual, LayerNorm and the next GEMM's split operand in
  // one kernel); `op_ready`: d->op already holds the split operand of the next linear's input
  const bool sk = g_dec_split_kbs > 0 && Hd % 4 == 0 && Hd <= 1536;
  const bool fuse_op = sk && Tmax <= 64 && d->enc_len <= 64;   // producers of a linear's input also write its split operand
  HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::dec_embed_launch(d->tok.as<long long>(), d->word_emb.ptr(), d->pos_emb.ptr(), d->emb_ln_w.ptr(),
                                                      d->emb_ln_b.ptr(), pos, x, fuse_op ? d->op.as<__nv_bfloat16>() : nullptr, R, Hd, s));
  // kvi: the beam re-order is an index table the self-attention reads through (one tiny kernel per step) instead of a copy of
  // every layer's K / V prefix (2 x R x (pos + 1) x hidden floats per layer per step)
  const bool kvi = g_dec_kv_index && Tmax <= 64;
  const int S_hd = dec_split_count(Hd), S_ff = dec_split_count(Ff);
  bool op_ready = fuse_op;
  for (auto& Lp : d->layers) {
    HbDecoder::Layer& L = *Lp;
    float* kc = L.kc[kvi ? 0 : d->cur].as<float>();
    float* vc = L.vc[kvi ? 0 : d->cur].as<float>();
    if (op_ready) { if ((r = dec_gemm_raw(d, R, Hd, d->tm_hd, L.sqkv, d->qkv.as<float>(), s))) return r; }
    else if ((r = dec_gemm(d, x, R, Hd, 0, d->tm_hd, L.sqkv, d->qkv.as<float>(), s))) return r;
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::dec_cache_append_launch(d->qkv.as<float>(), kc, vc, kvi ? d->kv_idx[d->cur].as<int>() : nullptr, pos, R,
                                                               Tmax, Hd, s));
    hb::SmallAttnF32Params ap;  // one query (the new token) over the cached prefix: every cached key is <= the query position
    ap.q = d->qkv.as<float>(); ap.k = kc; ap.v = vc; ap.out = d->att.as<float>();
    ap.B = R; ap.H = c.heads; ap.Tq = 1; ap.Tk = pos + 1;
    ap.ldq = 3 * Hd; ap.ldk = Hd; ap.ldv = Hd; ap.ldo = Hd;
    ap.bsq = 3 * Hd; ap.bsk = ap.bsv = static_cast<long long>(Tmax) * Hd; ap.bso = Hd;
    ap.scale = 0.125f; ap.mask_mode = 0;
    if (kvi) { ap.kv_row_idx = d->kv_idx[d->cur].as<int>(); ap.ld_idx = Tmax; }
    if (fuse_op) { ap.op_out = d->op.as<__nv_bfloat16>(); ap.ld_op = Hd; }
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::small_attn_f32_launch(ap, s));
    // self-attention output: s1 = LN(dense(att) + x); cross-attention query: qc = dense(s1)
    if (sk) {
      if (!fuse_op) HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::split3_act_launch(d->att.as<float>(), d->op.as<__nv_bfloat16>(), R, Hd, 0, s));
      if ((r = dec_gemm_split(d, R, Hd, d->tm_hd, L.so, S_hd, s))) return r;
      if ((r = dec_finish(d, R, L.so, S_hd, x, 0, &L.so_ln_w, &L.so_ln_b, d->s1.as<float>(), true, s))) return r;
      if ((r = dec_gemm_split(d, R, Hd, d->tm_hd, L.eq, S_hd, s))) return r;
      if ((r = dec_finish(d, R, L.eq, S_hd, nullptr, 0, nullptr, nullptr, d->qc.as<float>(), false, s))) return r;
    } else {
      if ((r = dec_gemm(d, d->att.as<float>(), R, Hd, 0, d->tm_hd, L.so, d->t1.as<float>(), s, x))) return r;
      if ((r = ln_f32(d->t1.as<float>(), d->s1.as<float>(), L.so_ln_w, L.so_ln_b, 1e-12f, R, Hd, s))) return r;
      if ((r = dec_gemm(d, d->s1.as<float>(), R, Hd, 0, d->tm_hd, L.eq, d->qc.as<float>(), s))) return r;
    }
    hb::SmallAttnF32Params cp;  // cross attention; the all-zeros video mask puts -10000 on every key (modeling.py:591)
    cp.q = d->qc.as<float>(); cp.k = L.ekv_buf.as<float>(); cp.v = L.ekv_buf.as<float>() + Hd; cp.out = d->att.as<float>();
    cp.B = R; cp.H = c.heads; cp.Tq = 1; cp.Tk = d->enc_len;
    cp.ldq = Hd; cp.ldk = cp.ldv = 2 * Hd; cp.ldo = Hd;
    cp.bsq = Hd; cp.bsk = cp.bsv = static_cast<long long>(d->enc_len) * 2 * Hd; cp.bso = Hd;
    cp.scale = 0.125f; cp.mask_mode = 2; cp.mask_const = -10000.0f; cp.kv_div = d->beam;
    if (fuse_op) { cp.op_out = d->op.as<__nv_bfloat16>(); cp.ld_op = Hd; }
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::small_attn_f32_launch(cp, s));
    // c1 = LN(dense(att) + s1); feed-forward: x = LN(dense(gelu(dense(c1))) + c1)
    if (sk) {
      if (!fuse_op) HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::split3_act_launch(d->att.as<float>(), d->op.as<__nv_bfloat16>(), R, Hd, 0, s));
      if ((r = dec_gemm_split(d, R, Hd, d->tm_hd, L.eo, S_hd, s))) return r;
      if ((r = dec_finish(d, R, L.eo, S_hd, d->s1.as<float>(), 0, &L.eo_ln_w, &L.eo_ln_b, d->c1.as<float>(), true, s))) return r;
      if ((r = dec_gemm_raw(d, R, Hd, d->tm_hd, L.inter, d->mid.as<float>(), s))) return r;
      HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::split3_act_launch(d->mid.as<float>(), d->op.as<__nv_bfloat16>(), R, Ff, 1, s));
      if ((r = dec_gemm_split(d, R, Ff, d->tm_ffn, L.out, S_ff, s))) return r;
      if ((r = dec_finish(d, R, L.out, S_ff, d->c1.as<float>(), 0, &L.o_ln_w, &L.o_ln_b, x, true, s))) return r;
      op_ready = true;
    } else {
      if ((r = dec_gemm(d, d->att.as<float>(), R, Hd, 0, d->tm_hd, L.eo, d->t1.as<float>(), s, d->s1.as<float>()))) return r;
      if ((r = ln_f32(d->t1.as<float>(), d->c1.as<float>(), L.eo_ln_w, L.eo_ln_b, 1e-12f, R, Hd, s))) return r;
      if ((r = dec_gemm(d, d->c1.as<float>(), R, Hd, 0, d->tm_hd, L.inter, d->mid.as<float>(), s))) return r;
      if ((r = dec_gemm(d, d->mid.as<float>(), R, Ff, 1, d->tm_ffn, L.out, d->t1.as<float>(), s, d->c1.as<float>()))) return r;
      if ((r = ln_f32(d->t1.as<float>(), x, L.o_ln_w, L.o_ln_b, 1e-12f, R, Hd, s))) return r;
    }
  }
  // classifier on the last position only: transform (dense -> GELU -> LN) then the tied vocabulary projection
  if (sk) {
    if (!op_ready) HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::split3_act_launch(x, d->op.as<__nv_bfloat16>(), R, Hd, 0, s));
    if ((r = dec_gemm_split(d, R, Hd, d->tm_hd, d->cls_dense, S_hd, s))) return r;
    if ((r = dec_finish(d, R, d->cls_dense, S_hd, nullptr, 1, &d->cls_ln_w, &d->cls_ln_b, nullptr, true, s))) return r;
    if ((r = dec_gemm_raw(d, R, Hd, d->tm_hd, d->cls_vocab, d->logits.as<float>(), s))) return r;
  } else {
    if ((r = dec_gemm(d, x, R, Hd, 0, d->tm_hd, d->cls_dense, d->th.as<float>(), s))) return r;
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::gelu_f32_launch(d->th.as<float>(), static_cast<long long>(R) * Hd, s));
    if ((r = ln_f32(d->th.as<float>(), d->t1.as<float>(), d->cls_ln_w, d->cls_ln_b, 1e-12f, R, Hd, s))) return r;
    if ((r = dec_gemm(d, d->t1.as<float>(), R, Hd, 0, d->tm_hd, d->cls_vocab, d->logits.as<float>(), s))) return r;
  }
  int* pk = d->prev_k.as<int>() + static_cast<size_t>(pos) * R;
  HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::beam_advance_launch(d->logits.as<float>(), d->Vpad, c.vocab, d->scores.as<float>(), d->done.as<int>(),
                                                         d->nsteps.as<int>(), d->prev_k.as<int>(), d->ys.as<int>(), d->tok.as<long long>(),
                                                         pos, d->n_inst, d->beam, c.eos, d->cand_v.as<float>(), d->cand_i.as<int>(), s));
  g_launches.fetch_add(1, std::memory_order_relaxed);   // beam_advance is two kernels
  // beams re-order: new beam j continues old beam prev_k[j]
  if (kvi) {
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::dec_index_advance_launch(d->kv_idx[d->cur].as<int>(), d->kv_idx[d->cur ^ 1].as<int>(), pk, pos + 1, R,
                                                                d->beam, Tmax, s));
    return HB_OK;
  }
  for (auto& Lp : d->layers) {
    HbDecoder::Layer& L = *Lp;
    HB_LAUNCH_P(CAT_OTHER, 0.0, s, hb::dec_cache_reorder_launch(L.kc[d->cur].as<float>(), L.vc[d->cur].as<float>(), L.kc[d->cur ^ 1].as<float>(),
                                                                L.vc[d->cur ^ 1].as<float>(), pk, pos + 1, R, d->beam, Tmax, Hd, s));
  }
  return HB_OK;
}

extern "C" {

int hb_decoder_step(HbDecoder* d, void* stream) {
  if (!d || d->n_inst <= 0) return fail(HB_ERR_INVALID, "hb_decoder_begin() not called");
  if (d->step >= d->cfg.max_words) return fail(HB_ERR_INVALID, "max_words steps already taken");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int r = HB_OK;
  // CUDA graphs: the first search with a given (n_inst, beam, enc_len) runs eagerly (it also warms every lazily set kernel
  // attribute); from the second search on each step index is captured once and then replayed with one cudaGraphLaunch instead of
  // ~50 kernel launches (launch-bound: a step is ~0.85 ms of 15-us kernels).  Per-launch profiling bypasses the graphs.
  const bool use_graph = g_decoder_graphs && !g_prof_on && d->gs != nullptr && d->gs->searches >= 1;
  if (use_graph) {
    HbDecoder::GraphSet& gs = *d->gs;
    if (gs.exec.size() != static_cast<size_t>(d->cfg.max_words)) { gs.exec.assign(d->cfg.max_words, nullptr); gs.launches.assign(d->cfg.max_words, 0); }
    cudaGraphExec_t& exec = gs.exec[d->step];
    if (exec == nullptr) {
      cudaGraph_t g = nullptr;
      const int64_t n0 = g_launches.load();
      // recording only: nothing runs on cap_stream, the instantiated graph is launched on the caller's stream below
      if (!d->cap_stream) HB_CUDA(cudaStreamCreateWithFlags(&d->cap_stream, cudaStreamNonBlocking));
      HB_CUDA(cudaStreamBeginCapture(d->cap_stream, cudaStreamCaptureModeThreadLocal));
      r = decoder_step_launches(d, d->cap_stream);
      gs.launches[d->step] = g_launches.exchange(n0) - n0;   // kernels recorded, not yet run: counted at every replay below
      cudaError_t e = cudaStreamEndCapture(d->cap_stream, &g);
      if (r != HB_OK || e != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        if (r == HB_OK) return fail(HB_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
        return r;
      }
      e = cudaGraphInstantiate(&exec, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) { exec = nullptr; return fail(HB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    }
    HB_CUDA(cudaGraphLaunch(exec, s));
    g_launches.fetch_add(gs.launches[d->step], std::memory_order_relaxed);
  } else {
    r = decoder_step_launches(d, s);
    if (r != HB_OK) return r;
  }
  d->cur ^= 1;
  d->step += 1;
  return HB_OK;
}

int hb_decoder_read(HbDecoder* d, int32_t* prev_k, int32_t* ys, int32_t* nsteps, int32_t* done, float* scores, void* stream) {
  if (!d || d->n_inst <= 0) return fail(HB_ERR_INVALID, "hb_decoder_begin() not called");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t R = static_cast<size_t>(d->n_inst) * d->beam;
  if (prev_k) HB_CUDA(cudaMemcpyAsync(prev_k, d->prev_k.p, static_cast<size_t>(d->cfg.max_words) * R * 4, cudaMemcpyDefault, s));
  if (ys) HB_CUDA(cudaMemcpyAsync(ys, d->ys.p, static_cast<size_t>(d->cfg.max_words) * R * 4, cudaMemcpyDefault, s));
  if (nsteps) HB_CUDA(cudaMemcpyAsync(nsteps, d->nsteps.p, static_cast<size_t>(d->n_inst) * 4, cudaMemcpyDefault, s));
  if (done) HB_CUDA(cudaMemcpyAsync(done, d->done.p, static_cast<size_t>(d->n_inst) * 4, cudaMemcpyDefault, s));
  if (scores) HB_CUDA(cudaMemcpyAsync(scores, d->scores.p, R * 4, cudaMemcpyDefault, s));
  return HB_OK;
}

}  // extern "C"
