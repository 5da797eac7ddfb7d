// hb_gemm.cu — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[M,N] = epilogue( A[M,K] x W[N,K]^T ),  A and W bf16 with K contiguous, fp32 accumulate in TMEM.
//
// Structure (one CTA per SM, or one CTA pair per TPC when CG == 2):
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> SWIZZLE_128B smem ring, mbarrier complete_tx); in the leader CTA
//                 also the dynamic tile scheduler (atomic tile counter -> st.async ring to both CTAs of the pair)
//   warp 1      : MMA issuer     (whole warp walks the loop, one elected lane issues tcgen05.mma kind::f16,
//                 128xN (CG=1) or 256xN (CG=2) tiles)
//   warp 2      : TMEM allocator (512 columns = two 256-column accumulator buffers)
//   warps 4..11 : epilogue       (tcgen05.ld -> bias / LayerNorm fold / GELU / residual -> per-warp smem staging ->
//                 line-coalesced st.global: four full 128-byte lines per store instruction)
// The accumulator is double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Reference call sites replaced: every F.linear on the hot path, e.g. EVA_clip/vit_model.py:57-61
// (fc1 -> GELU -> fc2), :124-126 (qkv with cat(q_bias, 0, v_bias)), :148 (proj), :198 (patch conv as
// GEMM), :350 (head); EVA_clip/eva_model.py:132,143-150,249 (text tower).
#include "hb_gemm.cuh"
#include "hb_ptx.cuh"

#include <cstdio>
#include <mutex>

namespace hb {

namespace {

constexpr int BM = 128;   // rows per CTA
constexpr int BN = 256;   // max columns per tile (one UMMA N)
constexpr int BK = 64;    // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;  // 384
constexpr int TMEM_COLS = 512;

template <int CG>
struct Cfg {
  static constexpr int W_ROWS = BN / CG;              // rows of W this CTA stages per k-block
  static constexpr int A_BYTES = BM * BK * 2;         // 16 KiB
  static constexpr int W_BYTES = W_ROWS * BK * 2;     // 32 / 16 KiB
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int STAGES = (CG == 1) ? 4 : 6;    // 192 KiB either way
  static constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;  // per epilogue warp: one 32x32 fp32 transpose buffer
  static constexpr int BIAS_BYTES = 2 * BN * 4;        // the tile's bias slice and its LayerNorm-fold c1 slice
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + NUM_EPI_WARPS * EPI_STAGE_BYTES + BIAS_BYTES;
};

// Static persistent tile schedule shared by the three roles of a worker: tiles are visited n-fastest, worker w takes
// tiles w, w + W, w + 2W, ...  (Tried and dropped: merging the narrow N-tail tiles of consecutive M-blocks into equal-cost
// work units to keep workers in lockstep — fc2's DRAM reads went UP, 3.4 -> 6.5 GB at M = 131584, and it ran 9 % slower;
// a worker count coprime to the number of N tiles, so every worker meets the tail tile equally often — no effect.)
struct TileIter {
  int t, stride, total, n_tiles;
  __device__ TileIter(int worker, int num_workers, int m_tiles, int n_tiles_) {
    t = worker; stride = num_workers; n_tiles = n_tiles_; total = m_tiles * n_tiles_;
  }
  __device__ bool next(int& m, int& n) {
    if (t >= total) return false;
    m = t / n_tiles; n = t - m * n_tiles;
    t += stride;
    return true;
  }
};

// Column tiling of N.  Plain: tiles of 256 with a narrow tail (1408 = 5 x 256 + 128).  Balanced (default): the same NUMBER of
// tiles, widths differing by at most one 32-column unit (1408 = 2 x 256 + 4 x 224, 4224 = 13 x 256 + 4 x 224): every tile costs
// about the same, so the workers that share an M-block's A operand through L2 stay in step (with a half-width tail, and a
// worker count that is even, the odd workers got all the tails, ran ~17 % faster and pulled the A k-blocks through DRAM twice).
struct NTiling {
  int n_tiles, base, rem, unit;
  // want > 0 (GemmParams::n_tiles): that many tiles instead of the minimum, e.g. a multiple of the worker count so that a
  // single-M-block GEMM over a huge N (the tied vocabulary projection, 120 tiles on 74 workers) has no half-empty last round
  __host__ __device__ NTiling(int N, int cg, int balanced, int want = 0) {
    n_tiles = (N + BN - 1) / BN;
    unit = 16 * cg;
    if (want > n_tiles && balanced && N % unit == 0 && want <= N / unit) n_tiles = want;
    if (balanced && N % unit == 0) {
      base = (N / n_tiles) / unit * unit;
      rem = (N - base * n_tiles) / unit;
    } else {
      base = BN; rem = 0; unit = 0;
    }
  }
  __host__ __device__ int n0(int n) const { return n * base + (n < rem ? n : rem) * unit; }
  __host__ __device__ int width(int n, int N) const {
    const int w = base + (n < rem ? unit : 0);
    const int left = N - n0(n);
    return w < left ? w : left;
  }
};

// ---- tile hand-out ----------------------------------------------------------------------------------------------------
// Static: worker w takes tiles w, w + W, w + 2W, ... (TileIter).  Dynamic (GemmParams::sched): tile 0..W-1 are taken the same
// way, every later tile id comes from an atomic counter, fetched one tile ahead by the leader CTA's producer warp and sent
// to both CTAs of the pair with st.async (value + mbarrier complete_tx in one operation) through a 4-slot ring.  Tiles are
// then STARTED in sequence order whatever delays individual workers pick up, so the N-tile workers that share an M-block's
// A operand through L2 always run within a few microseconds of each other (static schedules drift apart by up to a tile
// over the ~80 rounds of a launch and fc2 read its 3.2 GB A operand through DRAM twice).
constexpr int SCHED_RING = 4;

struct TileFeed {   // consumer view of either schedule; every lane of the warp calls next()
  bool dyn;
  TileIter st;
  uint64_t* full; uint64_t* empty; const int* tile;
  int it, total, n_tiles;
  bool arm, remote;
  __device__ TileFeed(bool dyn_, int worker, int num_workers, int m_tiles, int n_tiles_, uint64_t* f, uint64_t* e, const int* t,
                      bool arm_, bool remote_)
      : dyn(dyn_), st(worker, num_workers, m_tiles, n_tiles_), full(f), empty(e), tile(t), it(0), total(m_tiles * n_tiles_),
        n_tiles(n_tiles_), arm(arm_), remote(remote_) {}
  __device__ bool next(int& m, int& n) {
    if (!dyn) return st.next(m, n);
    const int slot = it % SCHED_RING;
    mbar_wait(&full[slot], static_cast<uint32_t>(it / SCHED_RING) & 1u);
    const int t = __shfl_sync(0xffffffffu, *reinterpret_cast<const volatile int*>(tile + slot), 0);
    __syncwarp();
    if (elect_one()) {
      if (arm) mbar_arrive_expect_tx(&full[slot], 4);          // arm the slot's next use (its current phase is complete)
      if (remote) mbar_arrive_cluster(&empty[slot], 0); else mbar_arrive(&empty[slot]);
    }
    __syncwarp();
    ++it;
    if (t >= total) return false;
    m = t / n_tiles; n = t - m * n_tiles;
    return true;
  }
};

// bf16-output epilogues: r = 32 consecutive fp32 accumulator columns [col0, col0+32) of this thread's row -> four 16-byte
// granules of packed bf16 (bias / LayerNorm fold / GELU / q-scale applied).
template <int EPI>
__device__ __forceinline__ void epilogue_pack_chunk(const GemmParams& p, const uint32_t (&r)[32], int col0,
                                                    const float* sb /*smem bias of this chunk*/, const float* sc1, float ln_a,
                                                    float ln_b, uint4 (&q)[4]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  {  // bias comes from the per-tile smem copy (zeros when the layer has none): broadcast LDS, no global-load latency here
    const float4* b4 = reinterpret_cast<const float4*>(sb);
    const float4* c4 = reinterpret_cast<const float4*>(sc1);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 b = b4[g];
      if constexpr (EPI == EPI_BF16_LN || EPI == EPI_GELU_BF16_LN) {
        // LayerNorm fold: ln_a = rstd, ln_b = -rstd * mean of this row
        const float4 c = c4[g];
        v[4 * g + 0] = fmaf(v[4 * g + 0], ln_a, fmaf(ln_b, c.x, b.x));
        v[4 * g + 1] = fmaf(v[4 * g + 1], ln_a, fmaf(ln_b, c.y, b.y));
        v[4 * g + 2] = fmaf(v[4 * g + 2], ln_a, fmaf(ln_b, c.z, b.z));
        v[4 * g + 3] = fmaf(v[4 * g + 3], ln_a, fmaf(ln_b, c.w, b.w));
      } else {
        v[4 * g + 0] += b.x;
        v[4 * g + 1] += b.y;
        v[4 * g + 2] += b.z;
        v[4 * g + 3] += b.w;
      }
    }
  }
  if constexpr (EPI == EPI_GELU_BF16 || EPI == EPI_GELU_BF16_LN) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else {
    if (col0 < p.qcols) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (col0 + j < p.qcols) ? v[j] * p.qscale : v[j];
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    q[g].x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]);
    q[g].y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
    q[g].z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
    q[g].w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
  }
}

template <int CG, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  using C = Cfg<CG>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;  // no static smem in this kernel: the dynamic window starts 1024-byte aligned (checked below)
  if ((smem_u32(smem) & 1023u) != 0u) __trap();

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full_bar = bars + 2 * C::STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint64_t* sched_full = tmem_empty_bar + 3;       // [SCHED_RING]  dynamic tile scheduler (p.sched != nullptr)
  uint64_t* sched_empty = sched_full + SCHED_RING; // [SCHED_RING]  (leader CTA's copy is the one that is used)
  int* sched_tile = reinterpret_cast<int*>(sched_empty + SCHED_RING);  // [SCHED_RING]
  float* sbias = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256 + NUM_EPI_WARPS * C::EPI_STAGE_BYTES);  // [2][BN]

  // warp-uniform values are routed through shfl(…, 0) so the compiler KNOWS they are uniform: the producer / MMA loops then
  // live in uniform registers and tcgen05 / TMA instructions issue directly.  With `if (lane == 0)` around those loops every
  // tcgen05.mma was wrapped in an ELECT + 7x R2UR.BROADCAST "waterfall" (136 instructions per k-block): the single issuing
  // thread, not the tensor pipe, was the bottleneck (ncu: MMA warp never waits on a barrier, tensor pipe 60-72 % active).
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);

  const int tile_m = BM * CG;
  const int m_tiles = (p.M + tile_m - 1) / tile_m;
  const NTiling nt(p.N, CG, p.balanced_n, p.n_tiles);
  const int n_tiles = nt.n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;
  // split-K (EPI_F32 only, GemmParams::k_splits): the split index rides on the N-tile index of the schedule — work item
  // (m, n') covers N tile n' % n_tiles and k-blocks [ks * kbs, ks * kbs + kbs) with ks = n' / n_tiles, and stores its raw
  // partial sums to out + ks * split_stride.  The host guarantees that every slice is non-empty.
  const int k_splits = p.k_splits > 1 ? p.k_splits : 1;
  const int kbs = (k_blocks + k_splits - 1) / k_splits;
  const int n_items = n_tiles * k_splits;
  const int worker = blockIdx.x / CG;
  const int num_workers = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], NUM_EPI_WARPS * CG);
    }
    for (int i = 0; i < SCHED_RING; ++i) {
      mbar_init(&sched_full[i], 1);
      mbar_init(&sched_empty[i], (NUM_EPI_WARPS + 1) * CG);   // per CTA: 8 epilogue warps + (MMA warp | the peer's producer warp)
    }
    fence_mbar_init();
    for (int i = 0; i < SCHED_RING; ++i) mbar_arrive_expect_tx(&sched_full[i], 4);   // first use of every slot is armed here
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, TMEM_COLS);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t smem_base = __shfl_sync(0xffffffffu, smem_u32(smem), 0);

  const bool dyn = (p.sched != nullptr);
  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t hint_a = p.a_hint == 1 ? kEvictFirst : (p.a_hint == 2 ? kEvictLast : kEvictNormal);
    const uint64_t hint_w = p.w_hint == 1 ? kEvictFirst : (p.w_hint == 2 ? kEvictLast : kEvictNormal);
    (void)hint_a; (void)hint_w;
    auto load_tile = [&](int m_blk, int n_item) {
      if (p.m_fastest) {   // same flat index, M-block fastest (see GemmParams::m_fastest)
        const int t = m_blk * n_items + n_item;
        n_item = t / m_tiles; m_blk = t - n_item * m_tiles;
      }
      const int ks = n_item / n_tiles, n_blk = n_item - ks * n_tiles;
      const int kb0 = ks * kbs, kb1 = min(k_blocks, kb0 + kbs);
      const int n0 = nt.n0(n_blk);
      const int n_eff = nt.width(n_blk, p.N);
      const int row_a = m_blk * tile_m + static_cast<int>(cta_rank) * BM;
      const int row_w = n0 + static_cast<int>(cta_rank) * (n_eff / CG);
      // (Tried and dropped: TMA L2-prefetch of the A boxes 8 k-blocks ahead — no measurable change; the ring is not latency-starved.)
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sw = sa + C::A_BYTES;
          if constexpr (CG == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, row_a);
            tma_load_2d(sw, &tmW, &full_bar[stage], kb * BK, row_w);
          } else {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
            // default hints: A normal (evict-first was 5 % slower: its k-blocks are shared by the N-tile workers; evict-last on
            // fc2's A: no change), W evict-last (weights are re-read by every M-block)
            tma_load_2d_pair_hint(sa, &tmA, &full_bar[stage], kb * BK, row_a, hint_a);
            tma_load_2d_pair_hint(sw, &tmW, &full_bar[stage], kb * BK, row_w, hint_w);
          }
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
    };
    if (dyn && leader) {
      // ---- scheduler: this warp hands tile ids to every consumer warp of the pair, one tile ahead of its own loads ----
      const int total = m_tiles * n_items;
      auto publish = [&](int it, int t) {
        const int slot = it % SCHED_RING;
        if (it >= SCHED_RING) mbar_wait(&sched_empty[slot], static_cast<uint32_t>(it / SCHED_RING - 1) & 1u);
        if (elect_one()) {
#pragma unroll
          for (uint32_t c = 0; c < static_cast<uint32_t>(CG); ++c)
            st_async_u32(mapa_u32(smem_u32(sched_tile + slot), c), static_cast<uint32_t>(t), mapa_u32(smem_u32(&sched_full[slot]), c));
        }
        __syncwarp();
      };
      int it = 0;
      int t = min(worker, total);
      publish(0, t);
      while (t < total) {
        int v = 0;
        if (lane == 0) v = num_workers + atomicAdd(p.sched, 1);
        const int t_next = min(__shfl_sync(0xffffffffu, v, 0), total);   // `total` is the end marker
        publish(it + 1, t_next);
        const int m_blk = t / n_items;
        load_tile(m_blk, t - m_blk * n_items);
        t = t_next;
        ++it;
      }
      if (lane == 0) {   // the last pair to finish re-zeroes the counters for the next launch on this stream
        if (atomicAdd(p.sched + 1, 1) == num_workers - 1) { p.sched[0] = 0; p.sched[1] = 0; __threadfence(); }
      }
      __syncwarp();
    } else {
      TileFeed feed(dyn, worker, num_workers, m_tiles, n_items, sched_full, sched_empty, sched_tile, /*arm=*/true, /*remote=*/true);
      int m_blk, n_blk;
      while (feed.next(m_blk, n_blk)) load_tile(m_blk, n_blk);
    }
  } else if (warp == 1) {
   if (leader) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    TileFeed it(dyn, worker, num_workers, m_tiles, n_items, sched_full, sched_empty, sched_tile, /*arm=*/true, /*remote=*/false);
    int m_blk, n_blk;
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B (umma_desc_sw128)
    while (it.next(m_blk, n_blk)) {
      if (p.m_fastest) {
        const int t = m_blk * n_items + n_blk;
        n_blk = t / m_tiles; m_blk = t - n_blk * m_tiles;
      }
      const int ks = n_blk / n_tiles;
      n_blk -= ks * n_tiles;
      const int n_kb = min(k_blocks, ks * kbs + kbs) - ks * kbs;
      const int n_eff = nt.width(n_blk, p.N);
      const uint32_t idesc = umma_idesc_bf16(BM * CG, static_cast<uint32_t>(n_eff));
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        // descriptor low words: (address >> 4) | LBO(1) << 16; advancing K by 16 bf16 = 32 bytes adds 2
        const uint32_t a_lo = (((smem_base + static_cast<uint32_t>(stage) * C::STAGE_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t w_lo = a_lo + (C::A_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = (static_cast<uint64_t>(DESC_HI) << 32) | (a_lo + 2u * k);
            const uint64_t wdesc = (static_cast<uint64_t>(DESC_HI) << 32) | (w_lo + 2u * k);
            umma_bf16<CG>(d_tmem, adesc, wdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit<CG>(&empty_bar[stage]);  // smem slot free once these MMAs retire
          if (kb == n_kb - 1) umma_commit<CG>(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
   }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;     // column half of the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    TileFeed it(dyn, worker, num_workers, m_tiles, n_items, sched_full, sched_empty, sched_tile, /*arm=*/false, /*remote=*/!leader);
    int m_blk, n_blk;
    while (it.next(m_blk, n_blk)) {
      if (p.m_fastest) {
        const int t = m_blk * n_items + n_blk;
        n_blk = t / m_tiles; m_blk = t - n_blk * m_tiles;
      }
      const int ks = n_blk / n_tiles;
      n_blk -= ks * n_tiles;
      const int n0 = nt.n0(n_blk);
      const int n_eff = nt.width(n_blk, p.N);
      const int row_in = m_blk * tile_m + static_cast<int>(cta_rank) * BM + q * 32 + lane;
      long long row_out = row_in;
      if (p.remap_in > 0) row_out = static_cast<long long>(row_in / p.remap_in) * p.remap_out + (row_in % p.remap_in) + p.remap_off;
      const int c_begin = half * (BN / 2);
      const int c_end = min(n_eff, (half + 1) * (BN / 2));
      // Stage this tile's bias slice in smem before the accumulator is ready (one element per epilogue thread): the
      // per-chunk global bias loads were consumed immediately and left the epilogue warps waiting on L2 latency
      // (fc1: 64 % tensor-pipe active with `stall_long_sb` on the bias FADDs, profiles/r01_one_layer_ncu_full.txt).
      float* sb_tile = sbias;
      float* sc1_tile = sbias + BN;
      {
        const int et = (warp - 4) * 32 + lane;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // every epilogue warp is done with the previous tile's slices
        sb_tile[et] = (p.bias != nullptr && et < n_eff) ? __ldg(p.bias + n0 + et) : 0.f;
        if constexpr (EPI == EPI_BF16_LN || EPI == EPI_GELU_BF16_LN) sc1_tile[et] = (et < n_eff) ? __ldg(p.c1 + n0 + et) : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if constexpr (EPI == EPI_F32 || EPI == EPI_F32_ROWADD || EPI == EPI_F32_STATS) {
        // fp32 output (+ fp32 residual / row-add): the accumulator chunk is transposed through a per-warp swizzled smem
        // buffer so that every global access is a full 128-byte row segment (thread-per-row access costs 32 L1
        // wavefronts per instruction and made this epilogue the bottleneck of the K=1408 proj GEMM).  Residual values
        // are prefetched one chunk ahead, in the coalesced mapping: instruction i covers rows 4i + lane/8, 16-byte
        // column chunk lane%8.
        const int rr_base = lane >> 3, cc = lane & 7;
        const uint32_t stage = smem_u32(smem + C::STAGES * C::STAGE_BYTES + 256) + static_cast<uint32_t>(warp - 4) * C::EPI_STAGE_BYTES;
        const int row_base = m_blk * tile_m + static_cast<int>(cta_rank) * BM + q * 32;
        float* out_f32 = reinterpret_cast<float*>(p.out) + static_cast<long long>(ks) * p.split_stride;
        uint32_t okmask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (row_base + 4 * i + rr_base < p.M) okmask |= (1u << i);
        }
        // element offset of output row i (and the row-add row), recomputed on use to keep registers for the prefetch
        auto row_off = [&](int i, int& radd) -> long long {
          const int ri = row_base + 4 * i + rr_base;
          radd = 0;
          if constexpr (EPI == EPI_F32_ROWADD) {  // row remapping exists only for the patch-embed kind
            if (p.remap_in > 0) {
              radd = ri % p.remap_in;
              return (static_cast<long long>(ri / p.remap_in) * p.remap_out + radd + p.remap_off) * p.ldo;
            }
          }
          return static_cast<long long>(ri) * p.ldo;
        };
        constexpr bool ROWADD = (EPI == EPI_F32_ROWADD);
        const bool has_add = ROWADD ? (p.rowadd != nullptr) : (p.resid != nullptr);
        // Optional (prefetch_chunks > 0, default off): each lane pulls its row's 128-byte residual line of the chunk
        // `prefetch_chunks` ahead into L2, so the one-chunk-ahead register prefetch only has to cover L2 latency.  Measured:
        // proj 1.062 -> 1.034 ms alone, nothing inside the power-capped step, fc2 +0.4 GB of DRAM reads (some lines are
        // evicted before use).  The earlier variant prefetched the NEXT TILE's residual, 15-50 us ahead — longer than a line
        // survives in L2 under this kernel's 3-5 TB/s of DRAM traffic: evicted and read twice (proj 3.0 vs 2.25 GB).
        const int pf_row = row_base + lane;
        auto l2_prefetch = [&](int c) {   // tile-relative first column of the chunk
          if (!ROWADD && p.resid != nullptr && p.prefetch_chunks > 0 && c < c_end && pf_row < p.M)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + static_cast<long long>(pf_row) * p.ldo + n0 + c));
        };
        auto load_add = [&](int col0, int n_valid, float4 (&res)[8]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (((okmask >> i) & 1u) && cc * 4 < n_valid) {
              // NOTE: one independent load per slot and nothing consuming it here — a dependent (even predicated-off)
              // instruction would wait on the load's scoreboard and serialise the eight prefetches.
              int radd;
              const long long off = row_off(i, radd);
              if constexpr (ROWADD) res[i] = __ldg(reinterpret_cast<const float4*>(p.rowadd + static_cast<long long>(radd) * p.N + col0 + cc * 4));
              else res[i] = __ldcs(reinterpret_cast<const float4*>(p.resid + off + col0 + cc * 4));
            }
          }
        };
        float psum[8], psq[8];   // EPI_F32_STATS: per-row partial (sum, sum of squares) over this warp's column span
#pragma unroll
        for (int i = 0; i < 8; ++i) { psum[i] = 0.f; psq[i] = 0.f; }
        auto do_chunk = [&](int c, const float4 (&res)[8]) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + c), r);
          tmem_ld_wait();
          const uint32_t my_row = stage + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + (static_cast<uint32_t>(g ^ (lane & 7)) << 4)),
                         "r"(r[4 * g]), "r"(r[4 * g + 1]), "r"(r[4 * g + 2]), "r"(r[4 * g + 3]) : "memory");
          }
          __syncwarp();
          const int col0 = n0 + c, n_valid = n_eff - c;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cc * 4 < n_valid) b4 = *reinterpret_cast<const float4*>(sb_tile + c + cc * 4);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + rr_base;
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(stage + static_cast<uint32_t>(rr) * 128u + (static_cast<uint32_t>(cc ^ (rr & 7)) << 4)));
            v.x += b4.x + res[i].x; v.y += b4.y + res[i].y; v.z += b4.z + res[i].z; v.w += b4.w + res[i].w;
            if (((okmask >> i) & 1u) && cc * 4 < n_valid) {
              int radd;
              __stcs(reinterpret_cast<float4*>(out_f32 + row_off(i, radd) + col0 + cc * 4), v);
              if constexpr (EPI == EPI_F32_STATS) {
                psum[i] += (v.x + v.y) + (v.z + v.w);
                psq[i] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                uint2 pk;
                pk.x = pack_bf16x2(v.x, v.y);
                pk.y = pack_bf16x2(v.z, v.w);
                const long long rowi = row_base + 4 * i + rr_base;
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.xb_out) + rowi * p.ld_xb + col0 + cc * 4) = pk;
              }
            }
          }
          __syncwarp();
        };
        float4 res_a[8], res_b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) res_a[i] = res_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#ifdef HB_EXP_INTERLEAVE_HALVES
        // EXPERIMENT (compile-time -DHB_EXP_INTERLEAVE_HALVES, not built by default; measured: results identical, proj 1079 ->
        // 1114 TFLOP/s alone, nothing inside the power-capped step): chunk k of this warp sits at
        // tile column 64 k + 32 half instead of 128 half + 32 k, so the two warps of a lane quarter touch ADJACENT 128-byte
        // lines of a row at about the same time (better DRAM page locality for the HBM-bound proj epilogue).  Statistics slots
        // stay one per (tile, half): any fixed partition of the columns works.
        auto colk = [&](int k) { return 64 * k + 32 * half; };
        if (has_add && colk(0) < n_eff) load_add(n0 + colk(0), n_eff - colk(0), res_a);
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int k = 0; colk(k) < n_eff; k += 2) {
          const int ca = colk(k), cb = colk(k + 1), cn = colk(k + 2);
          const bool second = cb < n_eff;
          if (has_add && second) load_add(n0 + cb, n_eff - cb, res_b);
          do_chunk(ca, res_a);
          if (second) {
            if (has_add && cn < n_eff) load_add(n0 + cn, n_eff - cn, res_a);
            do_chunk(cb, res_b);
          }
        }
#else
        if (has_add && c_begin < c_end) load_add(n0 + c_begin, n_eff - c_begin, res_a);
        for (int k = 1; k <= p.prefetch_chunks; ++k) l2_prefetch(c_begin + 32 * k);
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 64) {
          const bool second = (c + 32 < c_end);
          if (has_add && second) load_add(n0 + c + 32, n_eff - c - 32, res_b);
          l2_prefetch(c + 32 * (p.prefetch_chunks + 1));
          do_chunk(c, res_a);
          if (second) {
            if (has_add && c + 64 < c_end) load_add(n0 + c + 64, n_eff - c - 64, res_a);
            l2_prefetch(c + 32 * (p.prefetch_chunks + 2));
            do_chunk(c + 32, res_b);
          }
        }
#endif
        if constexpr (EPI == EPI_F32_STATS) {
          // the 8 lanes that share a row (same lane >> 3) combine their partials in a fixed order and one lane writes them to
          // this warp's slot (one per 128-column span): no atomics, so the statistics — and everything downstream — are
          // bit-reproducible and independent of which other rows share the launch.
          const int slot = 2 * n_blk + half;   // one slot per (N tile, column half): ln_slots >= 2 * n_tiles
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float a = psum[i], b = psq[i];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, o);
              b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (cc == 0 && ((okmask >> i) & 1u)) {   // an empty column half writes zeros: every slot is defined after every launch
              const long long rowi = row_base + 4 * i + rr_base;
              *reinterpret_cast<float2*>(p.stats_out + 2 * (rowi * p.ln_slots + slot)) = make_float2(a, b);
            }
          }
        }
      } else {
        const bool row_ok = row_in < p.M;
        float ln_a = 1.f, ln_b = 0.f;
        if constexpr (EPI == EPI_BF16_LN || EPI == EPI_GELU_BF16_LN) {
          if (row_ok) {   // row statistics of the A operand's fp32 source (accumulated by the producing GEMM's epilogue)
            float2 st = make_float2(0.f, 0.f);
            const float2* sp = reinterpret_cast<const float2*>(p.stats_in) + static_cast<long long>(row_in) * p.ln_slots;
            for (int sl = 0; sl < p.ln_slots; ++sl) { const float2 t = sp[sl]; st.x += t.x; st.y += t.y; }
            const float inv_d = 1.0f / static_cast<float>(p.ln_dim);
            const float mean = st.x * inv_d;
            const float var = fmaxf(st.y * inv_d - mean * mean, 0.f);
            ln_a = rsqrtf(var + p.ln_eps);
            ln_b = -ln_a * mean;
          }
        }
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        // 64 columns (= 128 bytes of bf16 per row) at a time go through this warp's 4 KiB staging buffer (32 rows x 128 B,
        // 16-byte granules XOR-swizzled by row) so that every global store instruction writes four full 128-byte lines.
        // Thread-per-row 16-byte stores touched 32 lines per instruction and left every L2 sector half filled (ncu: 16 of
        // 32 bytes per written sector, 139 M write requests per QKV launch).
        (void)row_ok;
        const uint32_t stage = smem_u32(smem + C::STAGES * C::STAGE_BYTES + 256) + static_cast<uint32_t>(warp - 4) * C::EPI_STAGE_BYTES;
        const int row_base = m_blk * tile_m + static_cast<int>(cta_rank) * BM + q * 32;
        const int rr_base = lane >> 3, gg = lane & 7;
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 64) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cc0 = c + 32 * h;
            if (cc0 < c_end) {
              uint32_t r[32];
              tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + cc0), r);
              tmem_ld_wait();
              uint4 pk[4];
              epilogue_pack_chunk<EPI>(p, r, n0 + cc0, sb_tile + cc0, sc1_tile + cc0, ln_a, ln_b, pk);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + static_cast<uint32_t>(lane) * 128u +
                                                                             (static_cast<uint32_t>((4 * h + g) ^ (lane & 7)) << 4)),
                             "r"(pk[g].x), "r"(pk[g].y), "r"(pk[g].z), "r"(pk[g].w) : "memory");
              }
            }
          }
          __syncwarp();
          const int col = c + gg * 8;                 // tile-relative first column of this lane's granule
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + rr_base;
            const int ri = row_base + rr;
            if (ri < p.M && col < c_end) {
              uint4 v;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                           : "r"(stage + static_cast<uint32_t>(rr) * 128u + (static_cast<uint32_t>(gg ^ (rr & 7)) << 4)));
              long long ro = ri;
              if (p.remap_in > 0) ro = static_cast<long long>(ri / p.remap_in) * p.remap_out + (ri % p.remap_in) + p.remap_off;
              // written once, far larger than L2: streaming store, do not displace the weights
              __stcs(reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + ro * p.ldo + n0 + col), v);
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(&tmem_empty_bar[acc]);
        else mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  // ===================== teardown =====================
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, TMEM_COLS);
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;
std::once_flag g_encode_once;
int g_encode_status = -1;

template <int CG, int EPI>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p, int num_sms, cudaStream_t stream) {
  using C = Cfg<CG>;
  static bool attr_set = false;
  auto kern = gemm_kernel<CG, EPI>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const int tile_m = BM * CG;
  const int m_tiles = (p.M + tile_m - 1) / tile_m;
  const int n_tiles = NTiling(p.N, CG, p.balanced_n, p.n_tiles).n_tiles;
  if ((EPI == EPI_F32_STATS) && p.ln_slots < 2 * n_tiles) return -8;
  if (p.k_splits > 1) {
    const int k_blocks = (p.K + BK - 1) / BK, kbs = (k_blocks + p.k_splits - 1) / p.k_splits;
    // raw partial sums only (the caller reduces the slices), every slice non-empty
    if (EPI != EPI_F32 || p.bias || p.resid || p.remap_in > 0 || p.split_stride <= 0 || (p.k_splits - 1) * kbs >= k_blocks) return -9;
  }
  const int total = m_tiles * n_tiles * (p.k_splits > 1 ? p.k_splits : 1);
  int workers = num_sms / CG;
  if (workers > total) workers = total;
  if (workers < 1) workers = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(workers * CG));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return static_cast<int>(cudaLaunchKernelEx(&cfg, kern, tmA, tmW, p));
}

}  // namespace

int tmap_init() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr) {
      g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
      g_encode_status = 0;
    } else {
      g_encode_status = -1;
    }
  });
  return g_encode_status;
}

int make_tmap_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  if (tmap_init() != 0) return -1;
  if ((ld * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(ptr) % 16) != 0) return -2;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(1000 + static_cast<int>(r));
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, const uint64_t dims[3], const uint64_t strides_bytes[2],
                      const uint32_t box[3], int swizzle_bytes) {
  if (tmap_init() != 0) return -1;
  if ((reinterpret_cast<uintptr_t>(ptr) % 16) != 0 || strides_bytes[0] % 16 != 0 || strides_bytes[1] % 16 != 0) return -2;
  cuuint64_t d[3] = {dims[0], dims[1], dims[2]};
  cuuint64_t st[2] = {strides_bytes[0], strides_bytes[1]};
  cuuint32_t bx[3] = {box[0], box[1], box[2]};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  if (swizzle_bytes != 128 && swizzle_bytes != 64 && swizzle_bytes != 0) return -3;
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), d, st, bx, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(1000 + static_cast<int>(r));
}

int gemm_n_tiling(int N, int cg, int balanced, int* n0_out, int* width_out, int cap) {
  if (N <= 0 || (cg != 1 && cg != 2)) return -1;
  const NTiling nt(N, cg, balanced);
  if (n0_out && width_out) {
    for (int n = 0; n < nt.n_tiles && n < cap; ++n) { n0_out[n] = nt.n0(n); width_out[n] = nt.width(n, N); }
  }
  return nt.n_tiles;
}

namespace { int g_balanced_n = 1; int g_prefetch_chunks = 0; int g_a_hint = -1; int g_w_hint = -1; }
void gemm_set_l2_hints(int a_hint, int w_hint) { g_a_hint = a_hint; g_w_hint = w_hint; }
void gemm_set_balanced_tiles(int on) { g_balanced_n = on ? 1 : 0; }
void gemm_set_resid_prefetch_chunks(int k) { g_prefetch_chunks = k < 0 ? 0 : (k > 3 ? 3 : k); }

int gemm_launch(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p_in, int epi, int cg, int num_sms,
                cudaStream_t stream) {
  GemmParams p = p_in;
  p.balanced_n = g_balanced_n;
  p.prefetch_chunks = g_prefetch_chunks;
  if (g_a_hint >= 0) p.a_hint = g_a_hint;
  if (g_w_hint >= 0) p.w_hint = g_w_hint;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return -3;
  if (p.N % 16 != 0) return -4;           // UMMA N granularity at M = 128/256 (and 16-byte stores)
  if (cg == 2 && p.N % 32 != 0) return -4;
  if (cg == 1) {
    switch (epi) {
      case EPI_BF16: return launch_impl<1, EPI_BF16>(tmA, tmW, p, num_sms, stream);
      case EPI_GELU_BF16: return launch_impl<1, EPI_GELU_BF16>(tmA, tmW, p, num_sms, stream);
      case EPI_F32: return launch_impl<1, EPI_F32>(tmA, tmW, p, num_sms, stream);
      case EPI_F32_ROWADD: return launch_impl<1, EPI_F32_ROWADD>(tmA, tmW, p, num_sms, stream);
      case EPI_F32_STATS: return launch_impl<1, EPI_F32_STATS>(tmA, tmW, p, num_sms, stream);
      case EPI_BF16_LN: return launch_impl<1, EPI_BF16_LN>(tmA, tmW, p, num_sms, stream);
      case EPI_GELU_BF16_LN: return launch_impl<1, EPI_GELU_BF16_LN>(tmA, tmW, p, num_sms, stream);
    }
  } else if (cg == 2) {
    switch (epi) {
      case EPI_BF16: return launch_impl<2, EPI_BF16>(tmA, tmW, p, num_sms, stream);
      case EPI_GELU_BF16: return launch_impl<2, EPI_GELU_BF16>(tmA, tmW, p, num_sms, stream);
      case EPI_F32: return launch_impl<2, EPI_F32>(tmA, tmW, p, num_sms, stream);
      case EPI_F32_ROWADD: return launch_impl<2, EPI_F32_ROWADD>(tmA, tmW, p, num_sms, stream);
      case EPI_F32_STATS: return launch_impl<2, EPI_F32_STATS>(tmA, tmW, p, num_sms, stream);
      case EPI_BF16_LN: return launch_impl<2, EPI_BF16_LN>(tmA, tmW, p, num_sms, stream);
      case EPI_GELU_BF16_LN: return launch_impl<2, EPI_GELU_BF16_LN>(tmA, tmW, p, num_sms, stream);
    }
  }
  return -5;
}

}  // namespace hb
