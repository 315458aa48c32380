// grad_bwd.cu -- backward of the symmetric InfoNCE head by tile recompute.
//
//   dX = scale * G Y,   G_ij = (exp(S_ij - lse_x_i) + exp(S_ij - lse_y_j) - 2 delta_ij) / (2 N)
//   dscale = sum_ij G_ij <X_i, Y_j>
// (closed form of autograd through VisionLanguageModule.py:459 and :550-552; called once with
// (X, Y) = (I, T) for dI and once with (T, I) for dT -- G is symmetric under the swap.)
//
// A 128-row block of dX in fp32 is 128 x 512 x 4 B = the whole 256 KB tensor memory of one SM, so the
// S tile and the accumulator cannot share an SM at d = 512. The kernel therefore runs on CTA PAIRS
// (cluster of 2, one CTA per SM):
//   rank 0 "producer":  S tile = X Y_t^T with X resident in TMEM (TS form), 8 softmax warps turn
//                       it into the fp16 tile G*2^13 (one ex2 per logit on the fast path, two on
//                       the guarded fallback), staged in local smem (K-major, 128B swizzle) and
//                       pushed to the peer with one cp.async.bulk shared::cta -> shared::cluster
//                       (18 B/cycle measured).
//   rank 1 "consumer":  dX block [128 x d] fp32 stays in TMEM for a whole segment of the sweep;
//                       acc += G_tile (A, smem, K-major) * Y_t (B, smem, MN-major), N = 256 per
//                       instruction; slot release back to the producer by a multicast
//                       tcgen05.commit. The epilogue scales and stores full 128-byte lines:
//                       final rows (to dX, or straight into the owning rank's NVLink peer window
//                       for the fused reduce-scatter) when the segment covers its row block, else
//                       the pair's partial slot, summed later in fixed order.
// Work is cut "stream-K" style: every SM pair sweeps the same number of tiles (see WorkRange).
// Neither S nor G ever leaves the SM pair.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include "common.cuh"
#include "grad_sched.cuh"
#include "../../include/vlpclip.h"

namespace vlp {

constexpr int SMX_GROUPS = 2;                      // column groups of the S tile: 4 softmax warps each (4 groups measured 3 % slower)
constexpr int SMX_COLS = 128 / SMX_GROUPS;         // logits per thread per tile
constexpr int SMX_WARPS = 4 * SMX_GROUPS;
constexpr int BWD_THREADS = 64 + 32 * SMX_WARPS;   // TMA warp + MMA warp + softmax warps
constexpr int P_KB_PER_STAGE = 2;           // producer ring stage: 2 boxes of [128 q x 64 k] fp16
constexpr int P_BOX_BYTES = 16384;
constexpr int P_STAGE_BYTES = P_KB_PER_STAGE * P_BOX_BYTES;
constexpr int P_STAGES = 4;
constexpr int C_STAGES = 4;                 // consumer ring: [64 q x 256 d] fp16 = 32 KB
constexpr int C_STAGE_BYTES = 32768;
constexpr int RING_BYTES = 131072;          // both rings occupy the first 128 KB
constexpr int G_SLOT_BYTES = 32768;         // [128 rows x 128 q] fp16
constexpr int G_SLOTS = 2;
constexpr int BAR_BYTES = 1024;               // barrier block
constexpr int EPI_STAGE_BYTES = 4 * 4096;    // consumer epilogue: one 32 x 32 fp32 tile per warp
constexpr uint32_t BWD_TMEM_X = 0;          // producer: X block (fp16 packed), d/2 columns
// producer S buffers (128 columns each) at the top of TMEM: two while d <= 512, one for d <= 768
constexpr float G_SCALE = 8192.f;           // 2^13: keeps softmax tails out of fp16 subnormals

constexpr int MAX_OWNERS = 8;   // one NVSwitch box
struct RowScatter {
  void* base[MAX_OWNERS];
  int rows_per_owner;   // 0: disabled
};

struct GradParams {
  const __half* x;
  int ldx;
  const float* xmax;    // [n_row_blocks*128] raw row max of X rows (0 padded)
  const float* xlg;     // [n_row_blocks*128] log2(sum) - k*max of X rows (+inf padded)
  const float* ymax;    // [total_tiles*128]  same for the rows of Y
  const float* ylg;
  const float* xr;      // [n_row_blocks*128] 2^(lse2_x - C)  (fast path; 0 for padded rows)
  const float* yc;      // [total_tiles*128]  2^(C - lse2_y)  (fast path; 0 for padded rows)
  const int* fast_flag; // 1: global LSE spread small enough for the single-ex2 path
  const float* xq;      // [n_rows] 1 - P_row(positive)   (unpadded, read only on the diagonal)
  const float* yq;      // [n_cols] 1 - P_col(positive)
  float w_row, w_col;
  int n_rows, n_cols, d;
  int kblocks, total_tiles, n_row_blocks;
  int db0, ndb;         // consumer: first 64-column block / number of blocks of dX in this pass
  int diag_shift;
  const float* scale_ptr;  // device scalar s
  float out_scale;      // 1 / (2 n_global) / 2^13   (multiplied by s in the epilogue)
  void* dx;             // [n_rows, d] final output (fp32 or bf16): row blocks swept by ONE cluster
  int dx_bf16;          // final output dtype
  const float* out_mul; // optional device scalar multiplied into the output (upstream gradient)
  float* part;          // [n_clusters][2][128, d] fp32 partial blocks of row blocks that are split
                        // between clusters (slot 0: head segment, slot 1: tail segment)
  float* ds_part;       // [n_clusters * SMX_WARPS] or nullptr
  long long* wait_prof; // [n_clusters][16] blocked-cycle counters (VLP_PROFILE_WAITS builds only)
  RowScatter scatter;   // optional: final rows go to per-owner buffers (fused reduce-scatter)
};

// ---- work partition ------------------------------------------------------------------------
// The (row block, column tile) grid is flattened row-block-major and cut into n_clusters equal
// contiguous ranges ("stream-K"): every cluster sweeps the same number of tiles (+-1) whatever the
// shape.  A range crosses row-block boundaries, so a cluster works through SEGMENTS
// (rb, [t0, t1)); a segment covering its whole row block writes the final rows, the (at most two)
// partial ones per cluster go to that cluster's partial slots and dx_reduce_kernel sums the
// pieces of each split row block in cluster order (fixed order => bit-reproducible).
struct Segment {
  int rb, t0, t1;
  int slot;   // -1: whole row block; 0 / 1: partial (head / tail segment of the cluster's range)
};
__host__ __device__ inline long long range_begin(int cluster, int n_clusters, long long total) {
  return total * cluster / n_clusters;
}
struct WorkRange {
  long long start, end, cur;
  int tiles;
  __host__ __device__ WorkRange(int cluster, int n_clusters, int n_row_blocks, int total_tiles) {
    const long long total = (long long)n_row_blocks * total_tiles;
    start = range_begin(cluster, n_clusters, total);
    end = range_begin(cluster + 1, n_clusters, total);
    cur = start;
    tiles = total_tiles;
  }
  __host__ __device__ bool next(Segment& s) {
    if (cur >= end) return false;
    s.rb = (int)(cur / tiles);
    const long long rb0 = (long long)s.rb * tiles;
    const long long seg_end = end < rb0 + tiles ? end : rb0 + tiles;
    s.t0 = (int)(cur - rb0);
    s.t1 = (int)(seg_end - rb0);
    s.slot = (s.t0 == 0 && s.t1 == tiles) ? -1 : (cur == start ? 0 : 1);
    cur = seg_end;
    return true;
  }
};
// cluster whose range holds flattened tile x
__host__ __device__ inline int range_owner(long long x, int n_clusters, long long total) {
  int c = (int)((x * n_clusters) / total);
  if (c >= n_clusters) c = n_clusters - 1;
  while (c + 1 < n_clusters && range_begin(c + 1, n_clusters, total) <= x) ++c;
  while (c > 0 && range_begin(c, n_clusters, total) > x) --c;
  return c;
}

// pieces of a split row block, as dx_reduce_kernel sums them
struct SplitBlock {
  int c_first, c_last, first_slot;
};
__host__ __device__ inline bool split_block(int rb, int tiles, int n_clusters, long long total,
                                            SplitBlock& sb) {
  const long long x0 = (long long)rb * tiles, x1 = x0 + tiles - 1;
  sb.c_first = range_owner(x0, n_clusters, total);
  sb.c_last = range_owner(x1, n_clusters, total);
  sb.first_slot = range_begin(sb.c_first, n_clusters, total) == x0 ? 0 : 1;
  return sb.c_first != sb.c_last;
}
__host__ __device__ inline bool split_source(const SplitBlock& sb, int c, int n_clusters,
                                             long long total, int& slot) {
  if (range_begin(c, n_clusters, total) == range_begin(c + 1, n_clusters, total)) return false;
  slot = c == sb.c_first ? sb.first_slot : 0;
  return true;
}

// destination of final row `row`: plain [n_rows, d] or, for the fused reduce-scatter, the buffer of
// the rank that owns the row (peer memory over NVLink), at the row's index inside that shard
__device__ __forceinline__ size_t scatter_row(const RowScatter& sc, int row, uint8_t*& base,
                                              void* dx) {
  if (sc.rows_per_owner <= 0) {
    base = reinterpret_cast<uint8_t*>(dx);
    return (size_t)row;
  }
  const int o = row / sc.rows_per_owner;
  base = reinterpret_cast<uint8_t*>(sc.base[o]);
  return (size_t)(row - o * sc.rows_per_owner);
}

// -DVLP_PROFILE_WAITS: cycles each role spends blocked on each barrier (dev tool, tools/wait_profile.py)
#ifdef VLP_PROFILE_WAITS
#define VLP_WAIT(idx, stmt)                  \
  do {                                       \
    const long long t0__ = clock64();        \
    stmt;                                    \
    wait_cyc[idx] += clock64() - t0__;       \
  } while (0)
#else
#define VLP_WAIT(idx, stmt) stmt
#endif

struct BwdBarriers {
  uint64_t full[P_STAGES];
  uint64_t empty[P_STAGES];
  uint64_t s_full[2];
  uint64_t s_empty[2];
  uint64_t x_ready;
  uint64_t x_free;
  uint64_t g_full[G_SLOTS];   // lives in the consumer, armed remotely by the producer
  uint64_t g_empty[G_SLOTS];  // lives in the producer, arrived by the consumer's commit
  uint64_t acc_full;
  uint64_t acc_free;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// P_row = 2^(k (s - xmax_i) - xlg_i),  P_col = 2^(k (s - ymax_j) - ylg_j): subtracting the raw
// maxima first keeps probabilities near 1 (dominant positive pair) accurate to ~1e-7.
template <bool kDiag>
__device__ __forceinline__ void softmax_tile(const uint32_t (&v)[SMX_COLS],
                                             const float4* __restrict__ ymax4,
                                             const float4* __restrict__ ylg4, float xmax, float xlg,
                                             float scale_log2, float diag_val, int diag_j,
                                             uint32_t (&out)[SMX_COLS / 2], float& ds_acc) {
#pragma unroll
  for (int q = 0; q < SMX_COLS / 4; ++q) {
    const float4 ym = __ldg(ymax4 + q);
    const float4 yl = __ldg(ylg4 + q);
    const float ymv[4] = {ym.x, ym.y, ym.z, ym.w};
    const float ylv[4] = {yl.x, yl.y, yl.z, yl.w};
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = q * 4 + e;
      const float s = __uint_as_float(v[j]);
      const float a = ex2_approx(fmaf(s - xmax, scale_log2, -xlg));
      const float b = ex2_approx(fmaf(s - ymv[e], scale_log2, -ylv[e]));
      float gg = a + b;
      if (kDiag) gg = (j == diag_j) ? diag_val : gg;  // = -(w_r (1-P_row) + w_c (1-P_col))
      ds_acc = fmaf(gg, s, ds_acc);
      g[e] = gg * G_SCALE;
    }
    out[q * 2 + 0] = pack_f16x2(g[0], g[1]);
    out[q * 2 + 1] = pack_f16x2(g[2], g[3]);
  }
}

// Single-ex2 variant: with a = w_row P_row 2^13 = 2^(k (s - xmax_i) - xlg_i + 13) the column
// term is a * 2^(lse2_x_i - lse2_y_j) = a * xr_i * yc_j (xr, yc precomputed around the global
// midpoint C of all LSEs).  Valid while every |lse2_x_i - lse2_y_j| stays far from the fp32
// exponent range; the prep kernel checks the global spread and sets fast_flag accordingly.
// Halves the MUFU work, which bounds this kernel (16 ex2 / clk / SM).
template <bool kDiag>
__device__ __forceinline__ void softmax_tile_fast(const uint32_t (&v)[SMX_COLS],
                                                  const float4* __restrict__ yc4, float xmax,
                                                  float xlg13, float xr, float scale_log2,
                                                  float diag_val_scaled, int diag_j,
                                                  uint32_t (&out)[SMX_COLS / 2], float& ds_acc) {
#pragma unroll
  for (int q = 0; q < SMX_COLS / 4; ++q) {
    const float4 yc = __ldg(yc4 + q);
    const float ycv[4] = {yc.x, yc.y, yc.z, yc.w};
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = q * 4 + e;
      const float s = __uint_as_float(v[j]);
      const float a = ex2_approx(fmaf(s - xmax, scale_log2, -xlg13));
      float gg = fmaf(a, xr * ycv[e], a);
      if (kDiag) gg = (j == diag_j) ? diag_val_scaled : gg;
      ds_acc = fmaf(gg, s, ds_acc);
      g[e] = gg;
    }
    out[q * 2 + 0] = pack_f16x2(g[0], g[1]);
    out[q * 2 + 1] = pack_f16x2(g[2], g[3]);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BWD_THREADS, 1)
grad_pair_kernel(const __grid_constant__ CUtensorMap map_y_k,   // box {64 k, 128 q}
                 const __grid_constant__ CUtensorMap map_y_mn,  // box {64 d, 64 q}
                 const GradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  BwdBarriers* bars =
      reinterpret_cast<BwdBarriers*>(smem + RING_BYTES + G_SLOTS * G_SLOT_BYTES);
  const uint32_t ring = smem_u32(smem);
  const uint32_t gslots = ring + RING_BYTES;
  const uint32_t stage = gslots + G_SLOTS * G_SLOT_BYTES + BAR_BYTES;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
#ifdef VLP_PROFILE_WAITS
  long long wait_cyc[16] = {0};
  const long long kernel_t0 = clock64();
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < P_STAGES; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->s_empty[i]), SMX_WARPS);
      mbar_init(smem_u32(&bars->g_full[i]), 1);
      mbar_init(smem_u32(&bars->g_empty[i]), 1);
    }
    mbar_init(smem_u32(&bars->x_ready), 8);
    mbar_init(smem_u32(&bars->x_free), 1);
    mbar_init(smem_u32(&bars->acc_full), 1);
    mbar_init(smem_u32(&bars->acc_free), 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(smem_u32(&bars->tmem_base), 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_y_k);
    tma_prefetch_desc(&map_y_mn);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const float scale_dev = __ldg(p.scale_ptr);
  const float scale_log2 = scale_dev * kLog2e;
  const uint32_t nbuf = p.kblocks <= 8 ? 2u : 1u;
  const uint32_t tmem_s_col = 512u - nbuf * 128u;

  if (rank == 0) {
    // =====================================================================================
    // producer CTA: S tiles + softmax -> G tiles pushed to the peer
    // =====================================================================================
    if (warp == 0) {
      uint32_t it = 0;
      WorkRange work(cluster_id, n_clusters, p.n_row_blocks, p.total_tiles);
      Segment sg;
      while (work.next(sg)) {
        const int t0 = sg.t0, t1 = sg.t1;
        for (int t = t0; t < t1; ++t)
          for (int kb = 0; kb < p.kblocks; kb += P_KB_PER_STAGE, ++it) {
            const uint32_t st = it % P_STAGES, ph = (it / P_STAGES) & 1;
            const int nkb = min(P_KB_PER_STAGE, p.kblocks - kb);
            VLP_WAIT(0, mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1));
            if (elect_one()) {
              mbar_expect_tx(smem_u32(&bars->full[st]), nkb * P_BOX_BYTES);
              for (int q = 0; q < nkb; ++q)
                tma_load_2d(ring + st * P_STAGE_BYTES + q * P_BOX_BYTES, &map_y_k,
                            smem_u32(&bars->full[st]), (kb + q) * 64, t * 128);
            }
            __syncwarp();
          }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc(UMMA_F16, UMMA_F16, MAJOR_K, MAJOR_K, 128, 128);
      uint32_t it = 0, tile_ctr = 0, item_ctr = 0;
      WorkRange work(cluster_id, n_clusters, p.n_row_blocks, p.total_tiles);
      Segment sg;
      for (; work.next(sg); ++item_ctr) {
        const int t0 = sg.t0, t1 = sg.t1;
        VLP_WAIT(1, mbar_wait(smem_u32(&bars->x_ready), item_ctr & 1));
        tc_fence_after();
        for (int t = t0; t < t1; ++t, ++tile_ctr) {
          const uint32_t buf = nbuf == 2 ? (tile_ctr & 1) : 0;
          const uint32_t use = nbuf == 2 ? (tile_ctr >> 1) : tile_ctr;
          VLP_WAIT(2, mbar_wait(smem_u32(&bars->s_empty[buf]), (use & 1) ^ 1));
          tc_fence_after();
          const uint32_t d_tmem = tmem + tmem_s_col + buf * 128;
          for (int kb = 0; kb < p.kblocks; kb += P_KB_PER_STAGE, ++it) {
            const uint32_t st = it % P_STAGES, ph = (it / P_STAGES) & 1;
            const int nkb = min(P_KB_PER_STAGE, p.kblocks - kb);
            VLP_WAIT(3, mbar_wait(smem_u32(&bars->full[st]), ph));
            tc_fence_after();
            if (elect_one()) {
              for (int q = 0; q < nkb; ++q) {
                const uint32_t sb = ring + st * P_STAGE_BYTES + q * P_BOX_BYTES;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_ts<1>(d_tmem, tmem + BWD_TMEM_X + (kb + q) * 32 + ks * 8,
                             make_sdesc_sw128(sb + ks * 32, 0, 1024), idesc, (kb | q | ks) != 0);
              }
              umma_commit<1>(smem_u32(&bars->empty[st]));
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit<1>(smem_u32(&bars->s_full[buf]));
          __syncwarp();
        }
        if (elect_one()) umma_commit<1>(smem_u32(&bars->x_free));
        __syncwarp();
      }
      if (item_ctr > 0) mbar_wait(smem_u32(&bars->x_free), (item_ctr - 1) & 1);
    } else {
      // ---- softmax warps ----
      const uint32_t quarter = warp & 3;
      const uint32_t grp = (warp - 2) >> 2;          // column group of the S tile
      const uint32_t row_in_blk = quarter * 32 + lane;
      const uint32_t lane_addr = (quarter * 32u) << 16;
      const int dp = p.kblocks * 64;
      const uint32_t sw = row_in_blk & 7;
      uint32_t tile_ctr = 0, item_ctr = 0;
      double ds_total = 0.0;   // lane 0: this warp's share of dscale over the cluster's whole range
      WorkRange work(cluster_id, n_clusters, p.n_row_blocks, p.total_tiles);
      Segment sg;
      for (; work.next(sg); ++item_ctr) {
        const int rb = sg.rb;
        const int t0 = sg.t0, t1 = sg.t1;
        const int row = rb * 128 + row_in_blk;
        const bool row_ok = row < p.n_rows;
        if (item_ctr > 0) {
          VLP_WAIT(4, mbar_wait(smem_u32(&bars->x_free), (item_ctr - 1) & 1));
          tc_fence_after();
        }
        if (grp < 2) {   // the first 8 warps stage the X block (two K halves) into TMEM
          const int k_begin = grp * (dp / 2);
          const uint4* src =
              reinterpret_cast<const uint4*>(p.x + (size_t)(row_ok ? row : 0) * p.ldx);
          for (int c0 = 0; c0 < dp / 4; c0 += 16) {
            uint32_t v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int k = k_begin + c0 * 2 + q * 8;
              uint4 w = make_uint4(0, 0, 0, 0);
              if (row_ok && k < p.d) w = __ldg(src + (k >> 3));
              v[q * 4 + 0] = w.x;
              v[q * 4 + 1] = w.y;
              v[q * 4 + 2] = w.z;
              v[q * 4 + 3] = w.w;
            }
            tmem_st_x16(tmem + lane_addr + BWD_TMEM_X + k_begin / 2 + c0, v);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->x_ready));
        }
        const float xmax = p.xmax[row];  // padded to n_row_blocks*128
        const float xlg = p.xlg[row];
        const float xr = p.xr[row];
        const bool fast = *p.fast_flag != 0;
        const int dcol = row_ok ? row - p.diag_shift : -1000000000;
        float ds_acc = 0.f;

        for (int t = t0; t < t1; ++t, ++tile_ctr) {
          const uint32_t buf = nbuf == 2 ? (tile_ctr & 1) : 0;
          const uint32_t use = nbuf == 2 ? (tile_ctr >> 1) : tile_ctr;
          VLP_WAIT(5, mbar_wait(smem_u32(&bars->s_full[buf]), use & 1));
          tc_fence_after();
          uint32_t v[SMX_COLS];
          {
            const uint32_t a = tmem + lane_addr + tmem_s_col + buf * 128 + grp * SMX_COLS;
#pragma unroll
            for (int h = 0; h < SMX_COLS / 32; ++h)
              tmem_ld_x32(a + 32 * h, *reinterpret_cast<uint32_t(*)[32]>(&v[32 * h]));
            tmem_ld_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->s_empty[buf]));

          const int col0 = t * 128 + grp * SMX_COLS;
          const float4* ymax4 = reinterpret_cast<const float4*>(p.ymax + col0);
          const float4* ylg4 = reinterpret_cast<const float4*>(p.ylg + col0);
          uint32_t out[SMX_COLS / 2];
          const int diag_j = dcol - col0;
          const bool has_diag = diag_j >= 0 && diag_j < SMX_COLS;
          const bool any_diag = __any_sync(0xffffffffu, has_diag);
          float diag_val = 0.f;
          if (has_diag) diag_val = -(p.w_row * p.xq[row] + p.w_col * p.yq[dcol]);
          if (fast) {
            const float4* yc4 = reinterpret_cast<const float4*>(p.yc + col0);
            float acc = 0.f;   // carries the 2^13 tile scale
            if (any_diag)
              softmax_tile_fast<true>(v, yc4, xmax, xlg - 13.f, xr, scale_log2, diag_val * G_SCALE,
                                      diag_j, out, acc);
            else
              softmax_tile_fast<false>(v, yc4, xmax, xlg - 13.f, xr, scale_log2, 0.f, diag_j, out,
                                       acc);
            ds_acc = fmaf(acc, 1.0f / G_SCALE, ds_acc);
          } else if (any_diag) {
            softmax_tile<true>(v, ymax4, ylg4, xmax, xlg, scale_log2, diag_val, diag_j, out,
                               ds_acc);
          } else {
            softmax_tile<false>(v, ymax4, ylg4, xmax, xlg, scale_log2, 0.f, diag_j, out, ds_acc);
          }

          // stage the fp16 G tile (K-major, 128B swizzle) and push it to the consumer CTA
          const uint32_t slot = tile_ctr & 1;
          if (tile_ctr >= 2)
            VLP_WAIT(6, mbar_wait_cluster(smem_u32(&bars->g_empty[slot]), ((tile_ctr >> 1) - 1) & 1));
          // (the tile is two [128 rows x 64 logits] K-major blocks of 16 KB; 16-byte chunks of a
          // row are XOR-swizzled with the row index)
          const uint32_t dst = gslots + slot * G_SLOT_BYTES + ((grp * SMX_COLS) >> 6) * 16384 +
                               row_in_blk * 128;
          const uint32_t cb = ((grp * SMX_COLS) & 63) >> 3;
#pragma unroll
          for (int c = 0; c < SMX_COLS / 8; ++c) {
            const uint32_t a = dst + (((cb + c) ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(out[c * 4 + 0]),
                         "r"(out[c * 4 + 1]), "r"(out[c * 4 + 2]), "r"(out[c * 4 + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          VLP_WAIT(7, bar_sync(1, 32 * SMX_WARPS));
          if (warp == 2 && lane == 0) {
            const uint32_t rbar = mapa_shared(smem_u32(&bars->g_full[slot]), 1);
            const uint32_t rdst = mapa_shared(gslots + slot * G_SLOT_BYTES, 1);
            asm volatile(
                "mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(
                    rbar),
                "r"(G_SLOT_BYTES)
                : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], "
                "%2, [%3];" ::"r"(rdst),
                "r"(gslots + slot * G_SLOT_BYTES), "r"(G_SLOT_BYTES), "r"(rbar)
                : "memory");
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, o);
        ds_total += (double)ds_acc;
      }
      if (p.ds_part != nullptr && lane == 0)
        p.ds_part[(size_t)cluster_id * SMX_WARPS + (warp - 2)] = (float)ds_total;
      // drain: the consumer must have released every slot we pushed before we may exit
      for (uint32_t back = 0; back < 2 && back < tile_ctr; ++back) {
        const uint32_t tc = tile_ctr - 1 - back;
        mbar_wait_cluster(smem_u32(&bars->g_empty[tc & 1]), (tc >> 1) & 1);
      }
    }
  } else {
    // =====================================================================================
    // consumer CTA: dX block accumulates in TMEM over the whole column sweep
    // =====================================================================================
    const int n_nc = (p.ndb + 3) / 4;  // 256-wide accumulator chunks of this pass
    if (warp == 0) {
      uint32_t it = 0;
      WorkRange work(cluster_id, n_clusters, p.n_row_blocks, p.total_tiles);
      Segment sg;
      while (work.next(sg)) {
        const int t0 = sg.t0, t1 = sg.t1;
        for (int t = t0; t < t1; ++t)
          for (int nc = 0; nc < n_nc; ++nc) {
            const int nb = min(4, p.ndb - nc * 4);
            for (int kh = 0; kh < 2; ++kh, ++it) {
              const uint32_t st = it % C_STAGES, ph = (it / C_STAGES) & 1;
              VLP_WAIT(8, mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1));
              if (elect_one()) {
                mbar_expect_tx(smem_u32(&bars->full[st]), nb * 8192);
                for (int b = 0; b < nb; ++b)
                  tma_load_2d(ring + st * C_STAGE_BYTES + b * 8192, &map_y_mn,
                              smem_u32(&bars->full[st]), (p.db0 + nc * 4 + b) * 64,
                              t * 128 + kh * 64);
              }
              __syncwarp();
            }
          }
      }
    } else if (warp == 1) {
      uint32_t it = 0, tile_ctr = 0, item_ctr = 0;
      WorkRange work(cluster_id, n_clusters, p.n_row_blocks, p.total_tiles);
      Segment sg;
      for (; work.next(sg); ++item_ctr) {
        const int t0 = sg.t0, t1 = sg.t1;
        if (item_ctr > 0) {
          VLP_WAIT(9, mbar_wait(smem_u32(&bars->acc_free), (item_ctr - 1) & 1));
          tc_fence_after();
        }
        for (int t = t0; t < t1; ++t, ++tile_ctr) {
          const uint32_t slot = tile_ctr & 1;
          VLP_WAIT(10, mbar_wait_cluster(smem_u32(&bars->g_full[slot]), (tile_ctr >> 1) & 1));
          tc_fence_after();
          const uint32_t ga = gslots + slot * G_SLOT_BYTES;
          for (int nc = 0; nc < n_nc; ++nc) {
            const int nb = min(4, p.ndb - nc * 4);
            const uint32_t idesc = make_idesc(UMMA_F16, UMMA_F16, MAJOR_K, MAJOR_MN, 128, nb * 64);
            for (int kh = 0; kh < 2; ++kh, ++it) {
              const uint32_t st = it % C_STAGES, ph = (it / C_STAGES) & 1;
              VLP_WAIT(11, mbar_wait(smem_u32(&bars->full[st]), ph));
              tc_fence_after();
              if (elect_one()) {
                const uint32_t sb = ring + st * C_STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint64_t ad = make_sdesc_sw128(ga + kh * 16384 + i * 32, 0, 1024);
                  const uint64_t bd = make_sdesc_sw128(sb + i * 2048, 8192, 1024);
                  umma_ss<1>(tmem + nc * 256, ad, bd, idesc, !(t == t0 && kh == 0 && i == 0));
                }
                umma_commit<1>(smem_u32(&bars->empty[st]));
              }
              __syncwarp();
            }
          }
          // release the G slot in the producer CTA (rank 0)
          if (elect_one()) umma_commit_mcast<1>(smem_u32(&bars->g_empty[slot]), 0x1);
          __syncwarp();
        }
        if (elect_one()) umma_commit<1>(smem_u32(&bars->acc_full));
        __syncwarp();
      }
    } else if (warp < 6) {
      // ---- epilogue: TMEM accumulator -> global ----
      const uint32_t quarter = warp & 3;
      const uint32_t lane_addr = (quarter * 32u) << 16;
      uint32_t item_ctr = 0;
      WorkRange work(cluster_id, n_clusters, p.n_row_blocks, p.total_tiles);
      Segment sg;
      for (; work.next(sg); ++item_ctr) {
        mbar_wait(smem_u32(&bars->acc_full), item_ctr & 1);
        tc_fence_after();
        const bool final_out = sg.slot < 0;
        const float mulv =
            scale_dev * ((final_out && p.out_mul) ? p.out_scale * __ldg(p.out_mul) : p.out_scale);
        // Each lane owns one accumulator row (TMEM lane), but rows are 2 KB apart in memory: the
        // 32 x 32 fp32 chunk goes through a swizzled smem tile so that every store instruction
        // writes four full 128-byte lines (matters most for the NVLink peer stores).
        const int sub = lane >> 3, ch = lane & 7;
        uint8_t* orow8[8];   // destination row of (i * 4 + sub), i = 0..7
        bool ok8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + sub;
          const int grow = sg.rb * 128 + (int)quarter * 32 + r;
          ok8[i] = grow < p.n_rows;
          if (final_out) {
            uint8_t* base;
            const size_t rr = scatter_row(p.scatter, ok8[i] ? grow : 0, base, p.dx);
            orow8[i] = base + rr * p.d * (p.dx_bf16 ? 2 : 4);
          } else {
            orow8[i] = reinterpret_cast<uint8_t*>(
                p.part + ((size_t)(cluster_id * 2 + sg.slot) * 128 + quarter * 32 + r) * p.d);
          }
        }
        const bool as_bf16 = final_out && p.dx_bf16;
        const uint32_t stg = stage + (warp - 2) * 4096;
        const int cbase = p.db0 * 64;   // first dX column of this pass
        for (int cc = 0; cc < p.ndb * 64; cc += 32) {
          uint32_t v[32];
          tmem_ld_x32(tmem + lane_addr + cc, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t a = stg + lane * 128 + ((c ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a),
                         "f"(__uint_as_float(v[c * 4 + 0]) * mulv),
                         "f"(__uint_as_float(v[c * 4 + 1]) * mulv),
                         "f"(__uint_as_float(v[c * 4 + 2]) * mulv),
                         "f"(__uint_as_float(v[c * 4 + 3]) * mulv)
                         : "memory");
          }
          __syncwarp();
          const int col = cbase + cc + ch * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + sub;
            float4 o;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                         : "r"(stg + r * 128 + ((ch ^ (r & 7)) << 4))
                         : "memory");
            if (ok8[i] && col < p.d) {
              if (as_bf16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y);
                __nv_bfloat162 hi = __floats2bfloat162_rn(o.z, o.w);
                uint2 w;
                w.x = *reinterpret_cast<uint32_t*>(&lo);
                w.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(orow8[i] + (size_t)col * 2) = w;
              } else {
                *reinterpret_cast<float4*>(orow8[i] + (size_t)col * 4) = o;
              }
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->acc_free));
      }
    }
  }

#ifdef VLP_PROFILE_WAITS
  if (p.wait_prof != nullptr && lane == 0) {
    long long* o = p.wait_prof + (size_t)cluster_id * 16;
    // one representative warp per role writes its counters (indices are disjoint between roles)
    const bool rep = (rank == 0) ? (warp <= 2) : (warp <= 2);
    if (rep)
      for (int i = 0; i < 12; ++i)
        if (wait_cyc[i] != 0) o[i] = wait_cyc[i];
    if (warp == 0) o[12 + rank] = clock64() - kernel_t0;
  }
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

// Global range of the log2-domain LSEs (rows of X and of Y, direction weights folded in), stage 1:
// RANGE_BLOCKS blocks each write their (min, max) to part[2 * block].
constexpr int RANGE_BLOCKS = 32;
__global__ void lse_range_kernel(const float* __restrict__ xmax, const float* __restrict__ xlg,
                                 int nx, float log2wx, const float* __restrict__ ymax,
                                 const float* __restrict__ ylg, int ny, float log2wy,
                                 const float* __restrict__ scale_ptr, float* __restrict__ part) {
  __shared__ float smin[256], smax[256];
  const float scale_log2 = __ldg(scale_ptr) * kLog2e;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx + ny; i += gridDim.x * blockDim.x) {
    const float e = i < nx ? fmaf(scale_log2, xmax[i], xlg[i] - log2wx)
                           : fmaf(scale_log2, ymax[i - nx], ylg[i - nx] - log2wy);
    if (e == e && fabsf(e) != INFINITY) {   // a zero direction weight gives +inf: ignore
      lo = fminf(lo, e);
      hi = fmaxf(hi, e);
    }
  }
  smin[threadIdx.x] = lo;
  smax[threadIdx.x] = hi;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + s]);
      smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = smin[0];
    part[2 * blockIdx.x + 1] = smax[0];
  }
}

// stage 2 (inlined into the consumers of the range): midpoint C and the single-ex2 flag
__device__ __forceinline__ void lse_range_finish(const float* __restrict__ part, int force_slow,
                                                 float& c_mid, int& fast) {
  float l = INFINITY, h = -INFINITY;
  for (int b = 0; b < RANGE_BLOCKS; ++b) {
    l = fminf(l, part[2 * b]);
    h = fmaxf(h, part[2 * b + 1]);
  }
  // (a zero direction weight makes one factor of the factorisation vanish: use two ex2)
  const bool ok = !force_slow && (h >= l) && (h - l) < 60.f;
  c_mid = ok ? 0.5f * (l + h) : 0.f;
  fast = ok ? 1 : 0;
}

// pad the per-row statistics to whole tiles: max -> 0, lg2l -> +inf (probability 0); a direction
// weight w >= 0 is folded in as lg2l - log2(w) (w * 2^e = 2^(e + log2 w)).  fac = 2^(sign*(e - C))
// with e = k*max + lg2l is the fast-path row (sign +1) / column (sign -1) factor, 0 when padded.
__global__ void stats_pad_kernel(const float* __restrict__ mx, const float* __restrict__ lg, int n,
                                 int n_pad, float log2w, const float* __restrict__ scale_ptr,
                                 float sign,
                                 const float* __restrict__ range_part, int force_slow,
                                 int* __restrict__ fast_flag, float* __restrict__ mx_out,
                                 float* __restrict__ lg_out, float* __restrict__ fac_out) {
  __shared__ float c_sh;
  const float scale_log2 = __ldg(scale_ptr) * kLog2e;
  if (threadIdx.x == 0) {
    float c;
    int fast;
    lse_range_finish(range_part, force_slow, c, fast);
    c_sh = c;
    if (blockIdx.x == 0) *fast_flag = fast;
  }
  __syncthreads();
  const float c_mid = c_sh;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) {
    const float m = i < n ? mx[i] : 0.f;
    const float l = i < n ? lg[i] - log2w : INFINITY;
    mx_out[i] = m;
    lg_out[i] = l;
    const float e = fmaf(scale_log2, m, l);
    fac_out[i] = (i < n && fabsf(e) != INFINITY) ? exp2f(sign * (e - c_mid)) : 0.f;
  }
}

// Row blocks whose column sweep was split between clusters: dx rows = mul * (sum of the partial
// blocks in cluster order).  One block row of the grid per row block; unsplit ones return at once.
constexpr int RED_SPLIT = 32;
__global__ void dx_reduce_kernel(const float* __restrict__ part, int n_clusters, int n_row_blocks,
                                 int tiles, int n_rows, int d, const float* __restrict__ out_mul,
                                 int out_bf16, void* __restrict__ dx, const RowScatter scatter) {
  const int rb = blockIdx.x;
  const long long total = (long long)n_row_blocks * tiles;
  SplitBlock sb;
  if (!split_block(rb, tiles, n_clusters, total, sb)) return;
  const int rows = min(128, n_rows - rb * 128);
  const int d4 = d >> 2;
  const float m = out_mul ? __ldg(out_mul) : 1.f;
  const size_t blk4 = (size_t)128 * d4;
  const float4* part4 = reinterpret_cast<const float4*>(part);
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < rows * d4; i += gridDim.y * blockDim.x) {
    const int r = i / d4, c4 = i - r * d4;
    const size_t off = (size_t)r * d4 + c4;
    float4 a = part4[(size_t)(sb.c_first * 2 + sb.first_slot) * blk4 + off];
    for (int c = sb.c_first + 1; c <= sb.c_last; ++c) {
      int slot;
      if (!split_source(sb, c, n_clusters, total, slot)) continue;
      const float4 b = part4[(size_t)(c * 2 + slot) * blk4 + off];
      a.x += b.x;
      a.y += b.y;
      a.z += b.z;
      a.w += b.w;
    }
    uint8_t* base;
    const size_t orow = scatter_row(scatter, rb * 128 + r, base, dx);
    if (out_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(a.x * m, a.y * m);
      __nv_bfloat162 hi = __floats2bfloat162_rn(a.z * m, a.w * m);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(base)[orow * d4 + c4] = o;
    } else {
      reinterpret_cast<float4*>(base)[orow * d4 + c4] = make_float4(a.x * m, a.y * m, a.z * m, a.w * m);
    }
  }
}

__global__ void ds_reduce_kernel(const float* __restrict__ part, int n, float mul,
                                 float* __restrict__ out) {
  __shared__ double sh[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += (double)part[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(sh[0] * (double)mul);
}

#include "grad_both.cuh"

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

static int n_pairs_of_device() {
  const int n = usable_sms();
  return n > 1 ? n / 2 : 74;   // no device (host-side size queries): assume a B200
}

// optional CUDA-event bracket around the grad_pair_kernel launches of the next vlpclip_grad calls
// (bench.py's roofline line: the dominant kernel alone, on the stream it runs on)
struct KernelTimer {
  bool enabled = false;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool pending = false;
};
static KernelTimer& kernel_timer() {
  static KernelTimer t;
  return t;
}

static long long*& wait_prof_buffer() {
  static long long* p = nullptr;
  return p;
}

static int plan_clusters(int n_row_blocks, int total_tiles) {
  const long long total = (long long)n_row_blocks * total_tiles;
  const int n_pairs = n_pairs_of_device();
  return total < n_pairs ? (int)total : n_pairs;
}

static size_t grad_ws_bytes(int n_rows, int n_cols, int d) {
  const size_t nrb = (n_rows + 127) / 128, nt = (n_cols + 127) / 128;
  const size_t max_pairs = 74;   // sized for a whole B200 whatever the current SM limit
  const size_t partials = align256(max_pairs * 2 * 128 * (size_t)d * 4);
  return 3 * align256(nrb * 128 * 4) + 3 * align256(nt * 128 * 4) + align256(max_pairs * SMX_WARPS * 4) +
         partials + 1024;
}


// ---- single-recompute backward: roles and workspace layout -----------------------------------
// n_sms SMs -> np producers, np dI consumers and nq = n_sms - 2 np >= np dT consumers (one column per
// dT consumer at a time).  VLP_B200_GB_NP overrides the split (fewer producers = more dT consumers).
static bool gb_roles(int n_sms, int* np_out, int* nq_out) {
  if (n_sms < 3) return false;
  int np = n_sms / 3;
  const char* e = getenv("VLP_B200_GB_NP");
  if (e && atoi(e) > 0 && atoi(e) < np) np = atoi(e);
  *np_out = np;
  *nq_out = n_sms - 2 * np;
  return true;
}

struct GbLayout {
  Sched s;
  int n_sms, np, nq;
  size_t off_ds, off_parts, off_dy_part, off_ring, off_flags, flag_bytes, off_ids, total;
};
static bool gb_layout(int n_rows, int n_cols, int d, GbLayout& L) {
  const int nrb = (n_rows + 127) / 128, nt = (n_cols + 127) / 128;
  L.n_sms = usable_sms();
  if (L.n_sms <= 0) L.n_sms = 148;   // no device (host-side size queries): assume a B200
  if (!gb_roles(L.n_sms, &L.np, &L.nq)) return false;
  L.s = make_sched(nrb, nt, L.np, L.nq);
  size_t o = 3 * align256((size_t)nrb * 128 * 4) + 3 * align256((size_t)nt * 128 * 4) + 512;
  L.off_ds = o;
  o += align256((size_t)L.np * GB_SMX_WARPS * 4);
  L.off_parts = o;
  o += align256((size_t)L.s.n_parts * 128 * (size_t)d * 4);
  L.off_dy_part = o;
  o += align256((size_t)nt * 128 * (size_t)d * 4);
  L.off_ring = o;
  o += (size_t)L.np * GB_RING_DEPTH * G_SLOT_BYTES;
  L.off_flags = o;
  L.flag_bytes = (3 * (size_t)L.np * GB_RING_DEPTH + (size_t)nt) * GB_FLAG_STRIDE * 4;
  o += align256(L.flag_bytes);
  L.off_ids = o;                                   // column caption ids padded to whole tiles (masked loss)
  o += align256((size_t)nt * 128 * 4);
  L.total = o + 1024;
  return true;
}

}  // namespace vlp

using namespace vlp;

extern "C" {

size_t vlpclip_grad_workspace_bytes(int n_rows, int n_cols, int d) {
  if (n_rows <= 0 || n_cols <= 0 || d <= 0) return 0;
  return grad_ws_bytes(n_rows, n_cols, d);
}

static int grad_impl(const void* x, int ldx, const void* y, int ldy, const float* x_max,
                     const float* x_lg2l, const float* x_q, const float* y_max,
                     const float* y_lg2l, const float* y_q, int n_rows, int n_cols, int d,
                     const float* scale, int diag_shift, int n_global, float w_row, float w_col,
                     const float* out_mul, int dx_bf16, void* dx, float* dscale, void* workspace,
                     size_t workspace_bytes, void* stream_, const RowScatter* scatter) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_rows <= 0 || n_cols <= 0) return fail(-1, "grad: empty problem (%d x %d)", n_rows, n_cols);
  if (!x || !y || !scale || !x_max || !x_lg2l || !x_q || !y_max || !y_lg2l || !y_q ||
      (!dx && !scatter) || !workspace)
    return fail(-1, "grad: null pointer");
  if (d <= 0 || d % 8 != 0 || d > 768)
    return fail(-1, "grad: embedding dim %d unsupported (need a multiple of 8, <= 768)", d);
  if (ldx % 8 != 0 || ldy % 8 != 0) return fail(-1, "grad: row strides must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(dx) & 15) != 0)
    return fail(-1, "grad: X and dX must be 16-byte aligned");
  if (n_global <= 0) return fail(-1, "grad: bad n_global");
  if (!(w_row >= 0.f) || !(w_col >= 0.f) || !(w_row + w_col > 0.f))
    return fail(-1, "grad: direction weights must be >= 0 and not both zero");
  int rc = check_device_sm100();
  if (rc) return rc;
  if (workspace_bytes < grad_ws_bytes(n_rows, n_cols, d))
    return fail(-1, "grad: workspace too small (%zu < %zu)", workspace_bytes,
                grad_ws_bytes(n_rows, n_cols, d));

  GradParams p = {};
  p.x = (const __half*)x;
  p.ldx = ldx;
  p.n_rows = n_rows;
  p.n_cols = n_cols;
  p.d = d;
  p.kblocks = (d + 63) / 64;
  p.total_tiles = (n_cols + 127) / 128;
  p.n_row_blocks = (n_rows + 127) / 128;
  const int clusters = plan_clusters(p.n_row_blocks, p.total_tiles);
  if (clusters > 74) return fail(-1, "grad: %d SM pairs exceed the workspace layout", clusters);
  p.diag_shift = diag_shift;
  p.w_row = w_row;
  p.w_col = w_col;
  p.xq = x_q;
  p.yq = y_q;
  p.scale_ptr = scale;
  p.out_scale = 1.0f / (2.0f * (float)n_global) / G_SCALE;

  uint8_t* ws = (uint8_t*)workspace;
  const int npx = p.n_row_blocks * 128, npy = p.total_tiles * 128;
  float* xmax = (float*)ws;
  ws += align256((size_t)npx * 4);
  float* xlg = (float*)ws;
  ws += align256((size_t)npx * 4);
  float* ymax = (float*)ws;
  ws += align256((size_t)npy * 4);
  float* ylg = (float*)ws;
  ws += align256((size_t)npy * 4);
  float* xr = (float*)ws;
  ws += align256((size_t)npx * 4);
  float* yc = (float*)ws;
  ws += align256((size_t)npy * 4);
  float* range_part = (float*)ws;                 // RANGE_BLOCKS (min, max) pairs
  int* fast_flag = (int*)(ws + 256);
  ws += 512;
  float* ds_part = (float*)ws;
  ws += align256((size_t)74 * SMX_WARPS * 4);
  float* dx_part = (float*)ws;
  p.dx = dx;
  p.part = dx_part;
  p.dx_bf16 = dx_bf16;
  p.out_mul = out_mul;
  if (scatter) p.scatter = *scatter;
  p.wait_prof = wait_prof_buffer();
  p.xmax = xmax;
  p.xlg = xlg;
  p.ymax = ymax;
  p.ylg = ylg;
  p.xr = xr;
  p.yc = yc;
  p.fast_flag = fast_flag;
  p.ds_part = dscale ? ds_part : nullptr;
  const float l2wr = log2f(w_row), l2wc = log2f(w_col);
  const int force_slow = (w_row == 0.f || w_col == 0.f) ? 1 : 0;
  lse_range_kernel<<<RANGE_BLOCKS, 256, 0, stream>>>(x_max, x_lg2l, n_rows, l2wr, y_max, y_lg2l,
                                                     n_cols, l2wc, scale, range_part);
  VLP_COUNT_LAUNCH(1);
  stats_pad_kernel<<<(npx + 255) / 256, 256, 0, stream>>>(x_max, x_lg2l, n_rows, npx, l2wr,
                                                          scale, 1.f, range_part, force_slow,
                                                          fast_flag, xmax, xlg, xr);
  VLP_COUNT_LAUNCH(1);
  stats_pad_kernel<<<(npy + 255) / 256, 256, 0, stream>>>(y_max, y_lg2l, n_cols, npy, l2wc,
                                                          scale, -1.f, range_part, force_slow,
                                                          fast_flag, ymax, ylg, yc);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());

  CUtensorMap map_k, map_mn;
  rc = make_tmap_sw128(&map_k, y, 2, (uint64_t)d, (uint64_t)n_cols, (uint64_t)ldy, 128);
  if (rc) return rc;
  rc = make_tmap_sw128(&map_mn, y, 2, (uint64_t)d, (uint64_t)n_cols, (uint64_t)ldy, 64);
  if (rc) return rc;

  static_assert(sizeof(BwdBarriers) <= BAR_BYTES, "barrier block");
  const size_t smem = RING_BYTES + G_SLOTS * G_SLOT_BYTES + BAR_BYTES + EPI_STAGE_BYTES + 1024;
  VLP_CUDA_OK(set_smem_attr_once((const void*)grad_pair_kernel, (int)smem, 0));
  // the dX block of a pass must fit the consumer's 512 TMEM columns: d <= 512 in one pass,
  // 512 < d <= 768 in two passes of half the 64-column blocks each (S is recomputed per pass)
  const int n_pass = p.kblocks > 8 ? 2 : 1;
  const int per_pass = (p.kblocks + n_pass - 1) / n_pass;
  float* ds_keep = p.ds_part;
  KernelTimer& kt = kernel_timer();
  if (kt.enabled) VLP_CUDA_OK(cudaEventRecord(kt.e0, stream));
  for (int pass = 0; pass < n_pass; ++pass) {
    p.db0 = pass * per_pass;
    p.ndb = (p.kblocks - p.db0) < per_pass ? (p.kblocks - p.db0) : per_pass;
    p.ds_part = pass == 0 ? ds_keep : nullptr;
    grad_pair_kernel<<<clusters * 2, BWD_THREADS, smem, stream>>>(map_k, map_mn, p);
    VLP_COUNT_LAUNCH(1);
  }
  if (kt.enabled) {
    VLP_CUDA_OK(cudaEventRecord(kt.e1, stream));
    kt.pending = true;
  }
  {
    // (a no-op for row blocks swept by a single cluster)
    dx_reduce_kernel<<<dim3(p.n_row_blocks, RED_SPLIT), 256, 0, stream>>>(
        dx_part, clusters, p.n_row_blocks, p.total_tiles, n_rows, d, out_mul, dx_bf16, dx, p.scatter);
    VLP_COUNT_LAUNCH(1);
    VLP_CUDA_OK(cudaGetLastError());
  }
  if (dscale) {
    ds_reduce_kernel<<<1, 256, 0, stream>>>(ds_part, clusters * SMX_WARPS, 1.0f / (2.0f * (float)n_global),
                                            dscale);
  VLP_COUNT_LAUNCH(1);
    VLP_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// dev builds (-DVLP_PROFILE_WAITS): device buffer of 74 x 16 int64 the next grad kernels fill
int vlpclip_dev_set_wait_profile(void* buf) {
  wait_prof_buffer() = (long long*)buf;
  return 0;
}

int vlpclip_time_grad_kernel(int enable) {
  KernelTimer& kt = kernel_timer();
  if (enable && !kt.e0) {
    VLP_CUDA_OK(cudaEventCreate(&kt.e0));
    VLP_CUDA_OK(cudaEventCreate(&kt.e1));
  }
  kt.enabled = enable != 0;
  kt.pending = false;
  return 0;
}

// duration of the grad_pair_kernel launch(es) of the most recent vlpclip_grad call, in ms
// (synchronises with that call); negative when nothing was recorded
float vlpclip_last_grad_kernel_ms(void) {
  KernelTimer& kt = kernel_timer();
  if (!kt.pending) return -1.f;
  if (cudaEventSynchronize(kt.e1) != cudaSuccess) return -1.f;
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, kt.e0, kt.e1) != cudaSuccess) return -1.f;
  return ms;
}

// host-side view of the work partition (tests): seg rows = {cluster, row block, t0, t1, slot},
// red rows = {row block, cluster, slot} in summation order
int vlpclip_grad_plan(int n_row_blocks, int tiles, int n_clusters, int* seg, int max_seg, int* n_seg,
                      int* red, int max_red, int* n_red) {
  if (n_row_blocks <= 0 || tiles <= 0 || n_clusters <= 0 || !seg || !red || !n_seg || !n_red)
    return fail(-1, "grad_plan: bad arguments");
  const long long total = (long long)n_row_blocks * tiles;
  int ns = 0, nr = 0;
  for (int c = 0; c < n_clusters; ++c) {
    WorkRange w(c, n_clusters, n_row_blocks, tiles);
    Segment sg;
    while (w.next(sg)) {
      if (ns >= max_seg) return fail(-1, "grad_plan: segment buffer too small");
      int* o = seg + 5 * ns++;
      o[0] = c; o[1] = sg.rb; o[2] = sg.t0; o[3] = sg.t1; o[4] = sg.slot;
    }
  }
  for (int rb = 0; rb < n_row_blocks; ++rb) {
    SplitBlock sb;
    if (!split_block(rb, tiles, n_clusters, total, sb)) continue;
    for (int c = sb.c_first; c <= sb.c_last; ++c) {
      int slot;
      if (!split_source(sb, c, n_clusters, total, slot)) continue;
      if (nr >= max_red) return fail(-1, "grad_plan: reduce buffer too small");
      int* o = red + 3 * nr++;
      o[0] = rb; o[1] = c; o[2] = slot;
    }
  }
  *n_seg = ns;
  *n_red = nr;
  return 0;
}

int vlpclip_grad(const void* x, int ldx, const void* y, int ldy, const float* x_max,
                 const float* x_lg2l, const float* x_q, const float* y_max, const float* y_lg2l,
                 const float* y_q, int n_rows, int n_cols, int d, const float* scale,
                 int diag_shift, int n_global, float w_row, float w_col, const float* out_mul,
                 int dx_bf16, void* dx, float* dscale, void* workspace, size_t workspace_bytes,
                 void* stream) {
  return grad_impl(x, ldx, y, ldy, x_max, x_lg2l, x_q, y_max, y_lg2l, y_q, n_rows, n_cols, d, scale,
                   diag_shift, n_global, w_row, w_col, out_mul, dx_bf16, dx, dscale, workspace,
                   workspace_bytes, stream, nullptr);
}

int vlpclip_grad_scatter(const void* x, int ldx, const void* y, int ldy, const float* x_max,
                         const float* x_lg2l, const float* x_q, const float* y_max,
                         const float* y_lg2l, const float* y_q, int n_rows, int n_cols, int d,
                         const float* scale, int diag_shift, int n_global, float w_row, float w_col,
                         void* const* owner_rows, int n_owners, int rows_per_owner, float* dscale,
                         void* workspace, size_t workspace_bytes, void* stream) {
  if (!owner_rows || n_owners <= 0 || n_owners > MAX_OWNERS)
    return fail(-1, "grad_scatter: need 1..%d owner buffers (got %d)", MAX_OWNERS, n_owners);
  if (rows_per_owner <= 0 || (long long)rows_per_owner * n_owners < n_rows)
    return fail(-1, "grad_scatter: %d owners x %d rows do not cover %d rows", n_owners,
                rows_per_owner, n_rows);
  RowScatter sc = {};
  for (int i = 0; i < n_owners; ++i) {
    if (!owner_rows[i] || (reinterpret_cast<uintptr_t>(owner_rows[i]) & 15) != 0)
      return fail(-1, "grad_scatter: owner buffer %d is null or not 16-byte aligned", i);
    sc.base[i] = owner_rows[i];
  }
  sc.rows_per_owner = rows_per_owner;
  return grad_impl(x, ldx, y, ldy, x_max, x_lg2l, x_q, y_max, y_lg2l, y_q, n_rows, n_cols, d, scale,
                   diag_shift, n_global, w_row, w_col, nullptr, 0, nullptr, dscale, workspace,
                   workspace_bytes, stream, &sc);
}


// ---- single-recompute backward: dI, dT and dscale from ONE sweep over the logit tiles ----------
size_t vlpclip_grad_both_workspace_bytes(int n_rows, int n_cols, int d) {
  if (n_rows <= 0 || n_cols <= 0 || d <= 0) return 0;
  GbLayout L;
  if (!gb_layout(n_rows, n_cols, d, L)) return 0;
  return L.total;
}

static int grad_both_impl(const void* x, int ldx, const void* y, int ldy, const float* x_max,
                          const float* x_lg2l, const float* x_q, const float* y_max,
                          const float* y_lg2l, const float* y_q, int n_rows, int n_cols, int d,
                          const float* scale, int diag_shift, int n_global, float w_row, float w_col,
                          const float* out_mul, int dx_bf16, void* dx, int dy_bf16, void* dy,
                          void* const* dy_owner_rows, int n_owners, int rows_per_owner, float* dscale,
                          const int* row_ids, const int* col_ids, void* workspace,
                          size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if ((row_ids == nullptr) != (col_ids == nullptr))
    return fail(-1, "grad_both: row and column caption ids must be given together");
  if (n_rows <= 0 || n_cols <= 0) return fail(-1, "grad_both: empty problem (%d x %d)", n_rows, n_cols);
  if (!x || !y || !scale || !x_max || !x_lg2l || !x_q || !y_max || !y_lg2l || !y_q || !dx ||
      (!dy && !dy_owner_rows) || !workspace)
    return fail(-1, "grad_both: null pointer");
  if (d <= 0 || d % 8 != 0 || d > 768)
    return fail(-1, "grad_both: embedding dim %d unsupported (need a multiple of 8, <= 768)", d);
  if (ldx % 8 != 0 || ldy % 8 != 0) return fail(-1, "grad_both: row strides must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(y) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(dx) & 15) != 0 || (reinterpret_cast<uintptr_t>(dy) & 15) != 0)
    return fail(-1, "grad_both: X, Y, dX and dY must be 16-byte aligned");
  if (n_global <= 0) return fail(-1, "grad_both: bad n_global");
  if (!(w_row >= 0.f) || !(w_col >= 0.f) || !(w_row + w_col > 0.f))
    return fail(-1, "grad_both: direction weights must be >= 0 and not both zero");
  RowScatter dy_sc = {};
  if (dy_owner_rows) {
    if (n_owners <= 0 || n_owners > MAX_OWNERS)
      return fail(-1, "grad_both: need 1..%d owner buffers (got %d)", MAX_OWNERS, n_owners);
    if (rows_per_owner <= 0 || (long long)rows_per_owner * n_owners < n_cols)
      return fail(-1, "grad_both: %d owners x %d rows do not cover %d rows", n_owners,
                  rows_per_owner, n_cols);
    for (int i = 0; i < n_owners; ++i) {
      if (!dy_owner_rows[i] || (reinterpret_cast<uintptr_t>(dy_owner_rows[i]) & 15) != 0)
        return fail(-1, "grad_both: owner buffer %d is null or not 16-byte aligned", i);
      dy_sc.base[i] = dy_owner_rows[i];
    }
    dy_sc.rows_per_owner = rows_per_owner;
  }
  int rc = check_device_sm100();
  if (rc) return rc;
  GbLayout L;
  if (!gb_layout(n_rows, n_cols, d, L)) return fail(-1, "grad_both: needs at least 3 SMs");
  if (workspace_bytes < L.total)
    return fail(-1, "grad_both: workspace too small (%zu < %zu)", workspace_bytes, L.total);

  GradBothParams P = {};
  GradParams& p = P.g;
  p.x = (const __half*)x;
  p.ldx = ldx;
  p.n_rows = n_rows;
  p.n_cols = n_cols;
  p.d = d;
  p.kblocks = (d + 63) / 64;
  p.total_tiles = (n_cols + 127) / 128;
  p.n_row_blocks = (n_rows + 127) / 128;
  p.diag_shift = diag_shift;
  p.w_row = w_row;
  p.w_col = w_col;
  p.xq = x_q;
  p.yq = y_q;
  p.scale_ptr = scale;
  p.out_scale = 1.0f / (2.0f * (float)n_global) / G_SCALE;

  uint8_t* ws = (uint8_t*)workspace;
  const int npx = p.n_row_blocks * 128, npy = p.total_tiles * 128;
  float* xmax = (float*)ws;
  ws += align256((size_t)npx * 4);
  float* xlg = (float*)ws;
  ws += align256((size_t)npx * 4);
  float* ymax = (float*)ws;
  ws += align256((size_t)npy * 4);
  float* ylg = (float*)ws;
  ws += align256((size_t)npy * 4);
  float* xr = (float*)ws;
  ws += align256((size_t)npx * 4);
  float* yc = (float*)ws;
  ws += align256((size_t)npy * 4);
  float* range_part = (float*)ws;
  int* fast_flag = (int*)(ws + 256);
  uint8_t* base = (uint8_t*)workspace;
  float* ds_part = (float*)(base + L.off_ds);
  p.dx = dx;
  p.part = (float*)(base + L.off_parts);
  p.dx_bf16 = dx_bf16;
  p.out_mul = out_mul;
  p.xmax = xmax;
  p.xlg = xlg;
  p.ymax = ymax;
  p.ylg = ylg;
  p.xr = xr;
  p.yc = yc;
  p.fast_flag = fast_flag;
  p.ds_part = dscale ? ds_part : nullptr;
  P.dy = dy;
  P.dy_bf16 = dy_owner_rows ? 0 : dy_bf16;
  P.dy_mul = dy_owner_rows ? nullptr : out_mul;
  P.dy_scatter = dy_sc;
  P.dy_part = (float*)(base + L.off_dy_part);
  P.gring = base + L.off_ring;
  int* flags = (int*)(base + L.off_flags);
  P.ready = flags;
  P.done_i = flags + (size_t)L.np * GB_RING_DEPTH * GB_FLAG_STRIDE;
  P.done_t = P.done_i + (size_t)L.np * GB_RING_DEPTH * GB_FLAG_STRIDE;
  P.col_turn = P.done_t + (size_t)L.np * GB_RING_DEPTH * GB_FLAG_STRIDE;
  P.np = L.np;
  P.s = L.s;
  P.dy_direct = (dy != nullptr && !dy_owner_rows && !dy_bf16) ? 1 : 0;
  if (col_ids) {   // pad the column ids to whole tiles with a value no caption id (>= 0) can take
    int* yid_pad = (int*)(base + L.off_ids);
    VLP_CUDA_OK(cudaMemsetAsync(yid_pad, 0xFE, (size_t)npy * 4, stream));
    VLP_CUDA_OK(cudaMemcpyAsync(yid_pad, col_ids, (size_t)n_cols * 4, cudaMemcpyDeviceToDevice, stream));
    P.xid = row_ids;
    P.yid = yid_pad;
  }
  p.wait_prof = wait_prof_buffer();

  const float l2wr = log2f(w_row), l2wc = log2f(w_col);
  const int force_slow = (w_row == 0.f || w_col == 0.f) ? 1 : 0;
  lse_range_kernel<<<RANGE_BLOCKS, 256, 0, stream>>>(x_max, x_lg2l, n_rows, l2wr, y_max, y_lg2l,
                                                     n_cols, l2wc, scale, range_part);
  VLP_COUNT_LAUNCH(1);
  stats_pad_kernel<<<(npx + 255) / 256, 256, 0, stream>>>(x_max, x_lg2l, n_rows, npx, l2wr, scale,
                                                          1.f, range_part, force_slow, fast_flag,
                                                          xmax, xlg, xr);
  VLP_COUNT_LAUNCH(1);
  stats_pad_kernel<<<(npy + 255) / 256, 256, 0, stream>>>(y_max, y_lg2l, n_cols, npy, l2wc, scale,
                                                          -1.f, range_part, force_slow, fast_flag,
                                                          ymax, ylg, yc);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());

  CUtensorMap map_k, map_mn, map_xmn;
  rc = make_tmap_sw128(&map_k, y, 2, (uint64_t)d, (uint64_t)n_cols, (uint64_t)ldy, 128);
  if (rc) return rc;
  rc = make_tmap_sw128(&map_mn, y, 2, (uint64_t)d, (uint64_t)n_cols, (uint64_t)ldy, 64);
  if (rc) return rc;
  rc = make_tmap_sw128(&map_xmn, x, 2, (uint64_t)d, (uint64_t)n_rows, (uint64_t)ldx, 64);
  if (rc) return rc;
  // the dT sums leave through TMA tile stores / reduce-adds: 32 x 32 fp32 boxes of dy (direct) or dy_part
  CUtensorMap map_dy;
  rc = make_tmap_sw128(&map_dy, P.dy_direct ? dy : (void*)P.dy_part, 4, (uint64_t)d, (uint64_t)n_cols,
                       (uint64_t)d, 32, true);
  if (rc) return rc;

  VLP_CUDA_OK(set_smem_attr_once((const void*)grad_both_kernel, GB_SMEM, 1));
  const int n_pass = p.kblocks > 8 ? 2 : 1;
  const int per_pass = (p.kblocks + n_pass - 1) / n_pass;
  float* ds_keep = p.ds_part;
  KernelTimer& kt = kernel_timer();
  if (kt.enabled) VLP_CUDA_OK(cudaEventRecord(kt.e0, stream));
  for (int pass = 0; pass < n_pass; ++pass) {
    p.db0 = pass * per_pass;
    p.ndb = (p.kblocks - p.db0) < per_pass ? (p.kblocks - p.db0) : per_pass;
    p.ds_part = pass == 0 ? ds_keep : nullptr;
    VLP_CUDA_OK(cudaMemsetAsync(flags, 0, L.flag_bytes, stream));
    {
      // cooperative launch: the CTAs spin-wait on one another, so the grid must be co-resident as
      // a whole (a plain launch could interleave with another persistent grid and deadlock)
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(2 * L.np + L.nq);
      lc.blockDim = dim3(GB_THREADS);
      lc.dynamicSmemBytes = GB_SMEM;
      lc.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeCooperative;
      at[0].val.cooperative = 1;
      lc.attrs = at;
      lc.numAttrs = 1;
      VLP_CUDA_OK(cudaLaunchKernelEx(&lc, grad_both_kernel, map_k, map_mn, map_xmn, map_dy, P));
    }
    VLP_COUNT_LAUNCH(1);
  }
  if (kt.enabled) {
    VLP_CUDA_OK(cudaEventRecord(kt.e1, stream));
    kt.pending = true;
  }
  VLP_CUDA_OK(cudaGetLastError());
  for (int pi = 0; pi < L.s.n_ph; ++pi) {
    const SchedPhase& ph = L.s.ph[pi];
    if (ph.n_seg <= 1) continue;
    dx_seg_reduce_kernel<<<dim3(ph.n_rows, RED_SPLIT), 256, 0, stream>>>(
        p.part, ph, n_rows, d, out_mul, dx_bf16, dx, p.scatter);
    VLP_COUNT_LAUNCH(1);
    VLP_CUDA_OK(cudaGetLastError());
  }
  if (dscale) {
    ds_reduce_kernel<<<1, 256, 0, stream>>>(ds_part, L.np * GB_SMX_WARPS,
                                            1.0f / (2.0f * (float)n_global), dscale);
    VLP_COUNT_LAUNCH(1);
    VLP_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

int vlpclip_grad_both(const void* x, int ldx, const void* y, int ldy, const float* x_max,
                      const float* x_lg2l, const float* x_q, const float* y_max,
                      const float* y_lg2l, const float* y_q, int n_rows, int n_cols, int d,
                      const float* scale, int diag_shift, int n_global, float w_row, float w_col,
                      const float* out_mul, int dx_bf16, void* dx, int dy_bf16, void* dy,
                      void* const* dy_owner_rows, int n_owners, int rows_per_owner, float* dscale,
                      void* workspace, size_t workspace_bytes, void* stream) {
  return grad_both_impl(x, ldx, y, ldy, x_max, x_lg2l, x_q, y_max, y_lg2l, y_q, n_rows, n_cols, d, scale,
                        diag_shift, n_global, w_row, w_col, out_mul, dx_bf16, dx, dy_bf16, dy, dy_owner_rows,
                        n_owners, rows_per_owner, dscale, nullptr, nullptr, workspace, workspace_bytes, stream);
}

int vlpclip_grad_both_masked(const void* x, int ldx, const void* y, int ldy, const float* x_max,
                             const float* x_lg2l, const float* x_q, const float* y_max,
                             const float* y_lg2l, const float* y_q, int n_rows, int n_cols, int d,
                             const float* scale, int diag_shift, int n_global, float w_row, float w_col,
                             const float* out_mul, int dx_bf16, void* dx, int dy_bf16, void* dy,
                             void* const* dy_owner_rows, int n_owners, int rows_per_owner, float* dscale,
                             const int* row_ids, const int* col_ids, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (!row_ids || !col_ids) return fail(-1, "grad_both_masked: null caption ids");
  return grad_both_impl(x, ldx, y, ldy, x_max, x_lg2l, x_q, y_max, y_lg2l, y_q, n_rows, n_cols, d, scale,
                        diag_shift, n_global, w_row, w_col, out_mul, dx_bf16, dx, dy_bf16, dy, dy_owner_rows,
                        n_owners, rows_per_owner, dscale, row_ids, col_ids, workspace, workspace_bytes, stream);
}

// host-only: the schedule of the single-recompute backward for np producer slots / nq dT consumers.
// info = {phases, steps, dI partial slots, runs, waves, np, nq, 0}; prod rows = {slot, step, row
// block, column tile, dI partial slot or -1}; cons rows = {consumer, slot, step, row block, column
// tile, rank of the piece in its column, pieces of the column}, in the order each consumer works.
int vlpclip_grad_both_plan(int n_row_blocks, int n_col_tiles, int np, int nq, int* info, int* prod,
                           int max_prod, int* n_prod, int* cons, int max_cons, int* n_cons) {
  if (n_row_blocks <= 0 || n_col_tiles <= 0 || !info || !prod || !cons || !n_prod || !n_cons)
    return fail(-1, "grad_both_plan: bad arguments");
  if (np <= 0 || nq <= 0) {
    const int n_sms = usable_sms();
    if (!gb_roles(n_sms > 0 ? n_sms : 148, &np, &nq)) return fail(-1, "grad_both_plan: too few SMs");
  }
  if (nq < np) return fail(-1, "grad_both_plan: need nq >= np (got %d < %d)", nq, np);
  const Sched s = make_sched(n_row_blocks, n_col_tiles, np, nq);
  info[0] = s.n_ph; info[1] = s.t_total; info[2] = s.n_parts; info[3] = s.n_runs;
  info[4] = s.n_waves; info[5] = s.np; info[6] = s.nq; info[7] = 0;
  int npd = 0, ncs = 0;
  for (int a = 0; a < s.np; ++a)
    for (int pi = 0; pi < s.n_ph; ++pi) {
      const SchedPhase& ph = s.ph[pi];
      for (int w = 0; w < ph.n_waves; ++w) {
        VRow vr;
        if (!sched_vrow(s, ph, w, a, vr)) continue;
        for (int u = 0; u < ph.cs; ++u) {
          const int col = sched_col(ph, vr, a, u);
          if (col < 0) continue;
          if (npd >= max_prod) return fail(-1, "grad_both_plan: producer buffer too small");
          int* o = prod + 5 * npd++;
          o[0] = a; o[1] = ph.t0 + w * ph.cs + u; o[2] = vr.rb; o[3] = col; o[4] = vr.part;
        }
      }
    }
  for (int q = 0; q < s.nq; ++q) {
    PieceIter it(s, q);
    Piece pc;
    while (it.next(pc)) {
      int rank, total;
      sched_piece_rank(s, pc.col, pc.gw, pc.wrapped, rank, total);
      for (int a = pc.a_hi; a >= pc.a_lo; --a) {
        if (ncs >= max_cons) return fail(-1, "grad_both_plan: consumer buffer too small");
        int* o = cons + 7 * ncs++;
        o[0] = q; o[1] = a; o[2] = pc.t_hi + (pc.a_hi - a); o[3] = pc.rb_hi - (pc.a_hi - a);
        o[4] = pc.col; o[5] = rank; o[6] = total;
      }
    }
  }
  *n_prod = npd;
  *n_cons = ncs;
  return 0;
}

// out = mul * (slot 0 + slot 1 + ... ) in slot order: the local half of the fused reduce-scatter
// (every peer has stored its partial rows into its slot of this rank's window)
__global__ void slot_sum_kernel(const float4* __restrict__ slots, size_t slot_stride4, int n_slots,
                                size_t n4, const float* __restrict__ out_mul, int out_bf16,
                                void* __restrict__ out) {
  const float m = out_mul ? __ldg(out_mul) : 1.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = slots[i];
    for (int s = 1; s < n_slots; ++s) {
      const float4 b = slots[(size_t)s * slot_stride4 + i];
      a.x += b.x;
      a.y += b.y;
      a.z += b.z;
      a.w += b.w;
    }
    if (out_bf16) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(a.x * m, a.y * m);
      __nv_bfloat162 hi = __floats2bfloat162_rn(a.z * m, a.w * m);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(out)[i] = o;
    } else {
      reinterpret_cast<float4*>(out)[i] = make_float4(a.x * m, a.y * m, a.z * m, a.w * m);
    }
  }
}

int vlpclip_slot_sum(const float* slots, int n_slots, size_t slot_elems, const float* out_mul,
                     int out_bf16, void* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!slots || !out || n_slots <= 0 || slot_elems == 0 || slot_elems % 4 != 0)
    return fail(-1, "slot_sum: bad arguments (n_slots %d, slot_elems %zu)", n_slots, slot_elems);
  if ((reinterpret_cast<uintptr_t>(slots) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0)
    return fail(-1, "slot_sum: buffers must be 16-byte aligned");
  const size_t n4 = slot_elems / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  slot_sum_kernel<<<blocks, 256, 0, stream>>>((const float4*)slots, n4, n_slots, n4, out_mul,
                                              out_bf16, out);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- peer windows (CUDA IPC): the only device memory this library allocates itself ----------
int vlpclip_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
  if (!dev_ptr || !handle64 || bytes == 0) return fail(-1, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  VLP_CUDA_OK(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(-2, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  VLP_CUDA_OK(cudaMemset(p, 0, bytes));
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}

int vlpclip_peer_open(const unsigned char* handle64, void** dev_ptr) {
  if (!dev_ptr || !handle64) return fail(-1, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  VLP_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int vlpclip_peer_close(void* dev_ptr) {
  if (dev_ptr) VLP_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

int vlpclip_peer_free(void* dev_ptr) {
  if (dev_ptr) VLP_CUDA_OK(cudaFree(dev_ptr));
  return 0;
}

}  // extern "C"
