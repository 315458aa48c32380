// lse_fwd.cu -- forward statistics of the symmetric InfoNCE head.
//
// Computes, for every row i of X, the online (max, sum-exp) pair of  S_ij = scale * <X_i, Y_j>
// over all rows j of Y, plus the positive-pair logit S_ii, WITHOUT materialising S
// (reference: VisionLanguageModule.py:459 `logits = (I @ T.T) * logit_scale` and the
// log-softmax half of F.cross_entropy at :550 / :551).
//
// Mapping to the SM (one persistent CTA per SM, 320 threads):
//   warp 0      TMA producer: streams Y as [128 rows x 64 k] bf16 boxes (SW128) through an
//               8-deep smem ring (16 KB per stage).
//   warp 1      tcgen05 issuer: S[128 x 128] (fp32, TMEM, double buffered) = X * Y_tile^T with the
//               A operand (X row block, bf16 packed) RESIDENT IN TMEM (TS form: 74 cycles per
//               K=16 step instead of 107 for the SS form -- see profiles/r01_probe_notes.md).
//   warps 2..9  softmax: thread = (row, 64-column half); tcgen05.ld the S row, online max/sum in
//               the log2 domain, one ex2 per logit.
// Work item = (row block, chunk of column tiles); partial (m, l) pairs per (chunk, half) go to
// a workspace and are merged by lse_merge_kernel in fixed order (bit reproducible).
#include <cuda_bf16.h>
#include "common.cuh"
#include "../../include/vlpclip.h"

namespace vlp {

constexpr int FWD_KB_PER_STAGE = 2;          // k-blocks (128 rows x 64 bf16 = 16 KB boxes) per stage
constexpr int FWD_BOX_BYTES = 128 * 128;
constexpr int FWD_STAGE_BYTES = FWD_KB_PER_STAGE * FWD_BOX_BYTES;
constexpr int FWD_STAGES = 6;
constexpr int FWD_SMW = 16;                  // softmax warps: 4 TMEM lane quarters x 4 column groups
constexpr int FWD_CG = FWD_SMW / 4;          // column groups per tile
constexpr int FWD_CPT = 128 / FWD_CG;        // columns per thread (32)
constexpr int FWD_THREADS = 64 + FWD_SMW * 32;
// warps 0..15: softmax (warp & 3 = TMEM lane quarter); the TMA and MMA warps get the highest warp ids:
// the issue arbiter of an SM sub-partition prefers the highest id, and their issue latency is critical
constexpr int FWD_TMA_WARP = FWD_SMW;
constexpr int FWD_MMA_WARP = FWD_SMW + 1;
constexpr uint32_t TMEM_X_COL = 0;      // bf16 X block: d/2 columns (256 at d = 512, 384 at d = 768)
// S buffers of 128 columns sit at the top of TMEM: two (double buffered) while d <= 512, a single
// one for 512 < d <= 768 (MMA and softmax of consecutive tiles then serialise)

// what the softmax warps do with an S tile
enum : int {
  MODE_ROWS = 0,   // row (max, sum-exp) statistics
  MODE_COLS = 1,   // + fused column statistics
  MODE_RANK = 2,   // retrieval: how many columns beat the positive pair of each row (recall@k)
  MODE_TOPK = 3    // retrieval: the RK best columns of each row, (value desc, index asc) (precision@k)
};
constexpr int RK = 16;   // entries per row of the streaming top-k (k_for_precision_at_k goes up to 15, + self)

struct LseParams {
  int operand_f16;   // 0: X, Y are bf16; 1: fp16 (the backward's operand copies)
  const __nv_bfloat16* x;
  int ldx;
  int n_rows, n_cols, d;
  int kblocks;          // ceil(d / 64)
  int total_tiles;      // ceil(n_cols / 128)
  int tiles_per_chunk;
  int n_chunks;
  int n_row_blocks;
  int diag_shift;       // delta_ij = 1 iff i == j + diag_shift
  const float* scale_ptr;  // device scalar s = clamp(exp(logit_scale), max=100)
  float* part_m;        // [n_chunks * FWD_CG][n_rows] raw (unscaled) running max of <X_i, Y_j>
  float* part_l;
  float* diag;          // [n_rows]: raw <X_i, Y_{i - diag_shift}>
  // fused column statistics (kCols): per (row block, column) partial over the block's 128 rows,
  // sum_i exp2(k c_ij - ref) with a log2-domain reference `ref`; the positive pair is left out
  float* col_ref;       // [n_row_blocks][total_tiles * 128]
  float* col_l;         // [n_row_blocks][total_tiles * 128]
  // duplicate-caption mask (optional): a logit whose row and column carry the same caption id is
  // excluded from both soft-maxes unless it is the positive pair itself (reference _get_mask, :506-530)
  const int* xid;       // [n_rows] caption id of the rows, or nullptr
  const int* yid;       // [n_cols] caption id of the columns
  // retrieval modes: partial results per (chunk, column group), merged by rank_merge / topk_merge
  int* part_cnt;        // MODE_RANK [n_chunks * FWD_CG][n_rows]
  float* part_val;      // MODE_TOPK [n_chunks * FWD_CG][n_rows][RK]
  int* part_idx;        // MODE_TOPK [n_chunks * FWD_CG][n_rows][RK]
};

struct FwdBarriers {
  uint64_t full[FWD_STAGES];
  uint64_t empty[FWD_STAGES];
  uint64_t s_full[2];
  uint64_t s_empty[2];
  uint64_t x_ready;
  uint64_t x_free;
  uint32_t tmem_base;
  uint32_t pad_;
  float col_s[2][FWD_SMW][FWD_CPT];   // [tile parity][softmax warp][its column]: warp partial sums
  float col_r[2][FWD_SMW];            // their log2-domain references
  float diag_s[128];                  // MODE_RANK: S_ii of the block's rows (from the diagonal tile)
};

constexpr float COL_HEADROOM = 100.f;  // partial sums carry 2^100: 226 log2 units of range below a
                                       // warp's largest row maximum before a term can underflow

// (value desc, index asc): the order of a stable descending sort
__device__ __forceinline__ bool rk_better(float s, int c, float v, int i) {
  return s > v || (s == v && c < i);
}

template <int kMode>
__global__ void __launch_bounds__(FWD_THREADS, 1)
lse_partial_kernel(const __grid_constant__ CUtensorMap map_y, const LseParams p) {
  constexpr bool kCols = kMode == MODE_COLS;
  // MODE_RANK: every work item first sweeps the DIAGONAL tile of its row block (tile index = row
  // block: the positive pair of row i is column i) to learn S_ii from the same MMAs that produce
  // the values it is compared with -- duplicate captions then tie bit-exactly
  constexpr int kLead = kMode == MODE_RANK ? 1 : 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  FwdBarriers* bars = reinterpret_cast<FwdBarriers*>(smem + FWD_STAGES * FWD_STAGE_BYTES);
  const uint32_t ring = smem_u32(smem);
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < FWD_STAGES; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->s_empty[i]), FWD_SMW);  // one arrive per softmax warp
    }
    mbar_init(smem_u32(&bars->x_ready), FWD_SMW);
    mbar_init(smem_u32(&bars->x_free), 1);
    fence_mbar_init();
  }
  if (warp == FWD_MMA_WARP) tmem_alloc<1>(smem_u32(&bars->tmem_base), 512);
  if (warp == FWD_TMA_WARP && lane == 0) tma_prefetch_desc(&map_y);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  const int n_items = p.n_row_blocks * p.n_chunks;
  // k = s * log2(e), device-side scalar (the retrieval modes rank raw cosines: no temperature)
  const float scale_log2 = kMode >= MODE_RANK ? 0.f : __ldg(p.scale_ptr) * kLog2e;
  const uint32_t nbuf = p.kblocks <= 8 ? 2u : 1u;
  const uint32_t tmem_s_col = 512u - nbuf * 128u;

  if (warp == FWD_TMA_WARP) {
    // ================= TMA producer (whole warp converged, one elected lane issues) ==========
    uint32_t it = 0;  // running stage counter
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / p.n_row_blocks;
      const int t0 = chunk * p.tiles_per_chunk;
      const int t1 = min(p.total_tiles, t0 + p.tiles_per_chunk);
      const int rb_item = item % p.n_row_blocks;
      for (int q = 0; q < kLead + t1 - t0; ++q) {
        const int t = q < kLead ? rb_item : t0 + q - kLead;
        for (int kb = 0; kb < p.kblocks; kb += FWD_KB_PER_STAGE, ++it) {
          const uint32_t st = it % FWD_STAGES;
          const uint32_t ph = (it / FWD_STAGES) & 1;
          const int nkb = min(FWD_KB_PER_STAGE, p.kblocks - kb);
          mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(smem_u32(&bars->full[st]), nkb * FWD_BOX_BYTES);
            for (int q = 0; q < nkb; ++q)
              tma_load_2d(ring + st * FWD_STAGE_BYTES + q * FWD_BOX_BYTES, &map_y,
                          smem_u32(&bars->full[st]), (kb + q) * 64, t * 128);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == FWD_MMA_WARP) {
    // ================= MMA issuer (whole warp converged, one elected lane issues) ===========
    const uint32_t fmt = p.operand_f16 ? UMMA_F16 : UMMA_BF16;
    const uint32_t idesc = make_idesc(fmt, fmt, MAJOR_K, MAJOR_K, 128, 128);
    // lean issue path (see sm100_ptx.cuh): running stage / parity counters, descriptor words advanced by adds
    static_assert(FWD_KB_PER_STAGE == 2 && FWD_BOX_BYTES == 16384, "umma_ts_stage layout");
    const uint32_t b_hi = sdesc_hi_sw128(1024);
    const uint32_t b_lo0 = sdesc_lo_sw128(ring, 0);
    const uint32_t full0 = smem_u32(&bars->full[0]), empty0 = smem_u32(&bars->empty[0]);
    // ONE elected thread runs the whole issue loop (no per-stage elect / reconvergence)
    if (elect_one()) {
      uint32_t st = 0, par = 0, tile_ctr = 0, item_ctr = 0, peek = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_ctr) {
        const int chunk = item / p.n_row_blocks;
        const int t0 = chunk * p.tiles_per_chunk;
        const int t1 = min(p.total_tiles, t0 + p.tiles_per_chunk);
        mbar_wait(smem_u32(&bars->x_ready), item_ctr & 1);
        tc_fence_after();
        for (int q = 0; q < kLead + t1 - t0; ++q, ++tile_ctr) {
          const uint32_t buf = nbuf == 2 ? (tile_ctr & 1) : 0;
          const uint32_t use = nbuf == 2 ? (tile_ctr >> 1) : tile_ctr;
          mbar_wait(smem_u32(&bars->s_empty[buf]), (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem + tmem_s_col + buf * 128;
          for (int kb = 0; kb < p.kblocks; kb += FWD_KB_PER_STAGE) {
            if (!peek) mbar_wait(full0 + st * 8, par);
            const uint32_t nst = st + 1 == FWD_STAGES ? 0u : st + 1;
            const uint32_t npar = nst == 0 ? par ^ 1u : par;
            const uint32_t lo = b_lo0 + st * (FWD_STAGE_BYTES >> 4);
            const uint32_t at = tmem + TMEM_X_COL + kb * 32;
            if (p.kblocks - kb >= 2)
              peek = umma_ts_stage_peek<8>(d_tmem, at, lo, b_hi, idesc, kb != 0, empty0 + st * 8,
                                           full0 + nst * 8, npar);
            else
              peek = umma_ts_stage_peek<4>(d_tmem, at, lo, b_hi, idesc, kb != 0, empty0 + st * 8,
                                           full0 + nst * 8, npar);
            st = nst;
            par = npar;
          }
          umma_commit<1>(smem_u32(&bars->s_full[buf]));
        }
        umma_commit<1>(smem_u32(&bars->x_free));
      }
      // do not exit with an arrive still in flight
      if (item_ctr > 0) mbar_wait(smem_u32(&bars->x_free), (item_ctr - 1) & 1);
    }
    __syncwarp();
  } else {
    // ================= softmax warps =================
    // thread = (row, column group): quarter = TMEM lane quarter of the warp, cg = 32-column group
    const uint32_t quarter = warp & 3;
    const uint32_t wslot = warp;                 // 0 .. FWD_SMW-1
    const uint32_t cg = wslot >> 2;
    const uint32_t row_in_blk = quarter * 32 + lane;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    const int dp = p.kblocks * 64;               // padded K
    uint32_t tile_ctr = 0, item_ctr = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_ctr) {
      const int chunk = item / p.n_row_blocks;
      const int rb = item % p.n_row_blocks;
      const int t0 = chunk * p.tiles_per_chunk;
      const int t1 = min(p.total_tiles, t0 + p.tiles_per_chunk);
      const int row = rb * 128 + row_in_blk;
      const bool row_ok = row < p.n_rows;

      // ---- stage the X row block into TMEM (bf16 pairs packed per 32-bit column) ----
      if (item_ctr > 0) {
        mbar_wait(smem_u32(&bars->x_free), (item_ctr - 1) & 1);
        tc_fence_after();
      }
      {
        const int k_begin = cg * (dp / FWD_CG);    // this thread's share of the row: dp/4 elements
        const uint4* src = reinterpret_cast<const uint4*>(p.x + (size_t)(row_ok ? row : 0) * p.ldx);
        for (int c0 = 0; c0 < dp / (2 * FWD_CG); c0 += 8) {
          uint32_t v[8];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int k = k_begin + c0 * 2 + q * 8;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (row_ok && k < p.d) w = __ldg(src + (k >> 3));
            v[q * 4 + 0] = w.x;
            v[q * 4 + 1] = w.y;
            v[q * 4 + 2] = w.z;
            v[q * 4 + 3] = w.w;
          }
          tmem_st_x8(tmem + lane_addr + TMEM_X_COL + k_begin / 2 + c0, v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->x_ready));
      }

      // column holding this row's positive pair (none for padded rows)
      const int dcol = row_ok ? row - p.diag_shift : -1000000000;
      if (kMode >= MODE_RANK) {
        // ---- retrieval epilogues: ranks by the raw cosine, no exponentials ----
        float diag = 0.f;
        int cnt = 0;
        float tv[RK];
        int ti[RK];
        if (kMode == MODE_TOPK) {
#pragma unroll
          for (int e = 0; e < RK; ++e) {
            tv[e] = -INFINITY;
            ti[e] = 0x7fffffff;
          }
        }
        for (int q = 0; q < kLead + t1 - t0; ++q, ++tile_ctr) {
          const int t = q < kLead ? rb : t0 + q - kLead;
          const uint32_t buf = nbuf == 2 ? (tile_ctr & 1) : 0;
          const uint32_t use = nbuf == 2 ? (tile_ctr >> 1) : tile_ctr;
          mbar_wait(smem_u32(&bars->s_full[buf]), use & 1);
          tc_fence_after();
          uint32_t v[FWD_CPT];
          tmem_ld_x32(tmem + lane_addr + tmem_s_col + buf * 128 + cg * FWD_CPT, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->s_empty[buf]));
          const int col0 = t * 128 + cg * FWD_CPT;
          if (kMode == MODE_RANK && q < kLead) {   // diagonal tile: publish S_ii to the row's 4 threads
#pragma unroll
            for (int j = 0; j < FWD_CPT; ++j)
              if (col0 + j == dcol) bars->diag_s[row_in_blk] = __uint_as_float(v[j]);
            bar_sync(2, FWD_SMW * 32);
            diag = bars->diag_s[row_in_blk];
            continue;
          }
#pragma unroll
          for (int j = 0; j < FWD_CPT; ++j) {
            const int c = col0 + j;
            const float sv = c < p.n_cols ? __uint_as_float(v[j]) : -INFINITY;   // TMA zero-fill past n_cols
            if (kMode == MODE_RANK) {
              cnt += rk_better(sv, c, diag, dcol) ? 1 : 0;
            } else if (rk_better(sv, c, tv[RK - 1], ti[RK - 1])) {
              tv[RK - 1] = sv;
              ti[RK - 1] = c;
#pragma unroll
              for (int e = RK - 1; e > 0; --e) {   // one bubble pass: the list was sorted
                if (rk_better(tv[e], ti[e], tv[e - 1], ti[e - 1])) {
                  const float fv = tv[e];
                  tv[e] = tv[e - 1];
                  tv[e - 1] = fv;
                  const int fi = ti[e];
                  ti[e] = ti[e - 1];
                  ti[e - 1] = fi;
                }
              }
            }
          }
        }
        if (row_ok) {
          const size_t o = (size_t)(chunk * FWD_CG + cg) * p.n_rows + row;
          if (kMode == MODE_RANK) {
            p.part_cnt[o] = cnt;
          } else {
#pragma unroll
            for (int e = 0; e < RK; ++e) {
              p.part_val[o * RK + e] = tv[e];
              p.part_idx[o * RK + e] = ti[e];
            }
          }
        }
        if (kMode == MODE_RANK) bar_sync(2, FWD_SMW * 32);   // diag_s is rewritten by the next item
        continue;
      }
      float m_run = -INFINITY, mraw_run = -INFINITY, l_run = 0.f;
      const int my_id = (p.xid != nullptr && row_ok) ? __ldg(p.xid + row) : -1;
      for (int t = t0; t < t1; ++t, ++tile_ctr) {
        const uint32_t buf = nbuf == 2 ? (tile_ctr & 1) : 0;
        const uint32_t use = nbuf == 2 ? (tile_ctr >> 1) : tile_ctr;
        mbar_wait(smem_u32(&bars->s_full[buf]), use & 1);
        tc_fence_after();
        uint32_t v[FWD_CPT];
        tmem_ld_x32(tmem + lane_addr + tmem_s_col + buf * 128 + cg * FWD_CPT, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->s_empty[buf]));

        const int col0 = t * 128 + cg * FWD_CPT;
        // columns past n_cols were zero-filled by TMA: mask them out of the statistics
        if (col0 + FWD_CPT > p.n_cols) {
#pragma unroll
          for (int j = 0; j < FWD_CPT; ++j)
            if (col0 + j >= p.n_cols) v[j] = __float_as_uint(-INFINITY);
        }
        if (p.yid != nullptr) {   // duplicate captions: not negatives of this row (nor of that column)
          if (col0 + FWD_CPT <= p.n_cols) {
            const int4* y4 = reinterpret_cast<const int4*>(p.yid + col0);
#pragma unroll
            for (int q = 0; q < FWD_CPT / 4; ++q) {
              const int4 w = __ldg(y4 + q);
              const int c = col0 + q * 4;
              if (w.x == my_id && c + 0 != dcol) v[q * 4 + 0] = __float_as_uint(-INFINITY);
              if (w.y == my_id && c + 1 != dcol) v[q * 4 + 1] = __float_as_uint(-INFINITY);
              if (w.z == my_id && c + 2 != dcol) v[q * 4 + 2] = __float_as_uint(-INFINITY);
              if (w.w == my_id && c + 3 != dcol) v[q * 4 + 3] = __float_as_uint(-INFINITY);
            }
          } else {
#pragma unroll
            for (int j = 0; j < FWD_CPT; ++j)
              if (col0 + j < p.n_cols && __ldg(p.yid + col0 + j) == my_id && col0 + j != dcol)
                v[j] = __float_as_uint(-INFINITY);
          }
        }
        float mx = __uint_as_float(v[0]);
#pragma unroll
        for (int j = 1; j < FWD_CPT; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        // positive pair: remember its raw logit, keep it in the running max, but leave it out of
        // the sum (the merge step adds it back through log1p -> no cancellation for tiny losses)
        if (__any_sync(0xffffffffu, dcol >= col0 && dcol < col0 + FWD_CPT)) {
          float dv = 0.f;
#pragma unroll
          for (int j = 0; j < FWD_CPT; ++j)
            if (col0 + j == dcol) {
              dv = __uint_as_float(v[j]);
              v[j] = __float_as_uint(-INFINITY);
            }
          if (row_ok && dcol >= col0 && dcol < col0 + FWD_CPT) p.diag[row] = dv;
        }
        const float mraw_new = fmaxf(mraw_run, mx);
        const float m_new = mraw_new * scale_log2;  // log2-domain reference of this row
        if (mraw_new != -INFINITY) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int j = 0; j < FWD_CPT; j += 2) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(v[j]), scale_log2, -m_new));
            const float e1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), scale_log2, -m_new));
            s0 += e0;
            s1 += e1;
            if (kCols) {
              v[j] = __float_as_uint(e0);
              v[j + 1] = __float_as_uint(e1);
            }
          }
          l_run = l_run * ex2_approx(m_run - m_new) + (s0 + s1);
          m_run = m_new;
          mraw_run = mraw_new;
        } else if (kCols) {
#pragma unroll
          for (int j = 0; j < FWD_CPT; ++j) v[j] = 0u;
        }

        if (kCols) {
          // ---- column sums of this 32-row x 32-column slab: sum_i e_ij * 2^(m_i - R + 100) ----
          float R = (row_ok && mraw_new != -INFINITY) ? m_new : -INFINITY;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) R = fmaxf(R, __shfl_xor_sync(0xffffffffu, R, o));
          const float f = (row_ok && mraw_new != -INFINITY)
                              ? ex2_approx(m_new - R + COL_HEADROOM) : 0.f;
          float c[FWD_CPT];
#pragma unroll
          for (int j = 0; j < FWD_CPT; ++j) c[j] = __uint_as_float(v[j]) * f;
          // butterfly: after the xor-16/8/4/2/1 steps lane L holds column L of the group
#pragma unroll
          for (int w = FWD_CPT / 2, o = 16; o > 0; w >>= 1, o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < w; ++j) {
              const float send = up ? c[j] : c[j + w];
              const float keep = up ? c[j + w] : c[j];
              c[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          const uint32_t par = tile_ctr & 1;
          bars->col_s[par][wslot][lane] = c[0];
          if (lane == 0) bars->col_r[par][wslot] = R - COL_HEADROOM;
          bar_sync(2, FWD_SMW * 32);
          const uint32_t st = wslot * 32 + lane;   // the first 128 softmax threads own one column
          if (st < 128) {
            const uint32_t g = st >> 5, cc = st & 31;   // warps of column group g: slots g*4 .. g*4+3
            float r4[4], l4[4], M = -INFINITY;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              r4[q] = bars->col_r[par][g * 4 + q];
              l4[q] = bars->col_s[par][g * 4 + q][cc];
              M = fmaxf(M, r4[q]);
            }
            float l = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (r4[q] != -INFINITY) l += l4[q] * ex2_approx(r4[q] - M);
            const size_t o = (size_t)rb * ((size_t)p.total_tiles * 128) + (size_t)t * 128 + st;
            p.col_ref[o] = M;
            p.col_l[o] = l;
          }
        }
      }
      if (row_ok) {
        const size_t o = (size_t)(chunk * FWD_CG + cg) * p.n_rows + row;
        p.part_m[o] = mraw_run;
        p.part_l[o] = l_run;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FWD_MMA_WARP) tmem_dealloc<1>(tmem, 512);
}

// Merge [nparts][n] partial (max, l) pairs in fixed order.
// A partial means  sum_{j != positive} exp(scale * c_j) = l * 2^fl(k * max)  with
// k = scale * log2(e), fl() the fp32 product (exactly what the streaming kernel subtracted inside
// ex2) and max taken over all columns INCLUDING the positive pair.  With t = 2^(k (diag - max)):
//   out_lg2l = log2(l + t) - k*max   (fl() residual removed with an exact fma)
//   out_q    = l / (l + t) = 1 - P(positive)
//   out_loss = ln(sum_j exp(S_ij)) - S_ii = ln2 * (log2(l + t) + k (max - diag)), via log1p when the
//              positive pair is the row maximum
// diag == nullptr: no positive pair in these columns (t = 0).
__global__ void lse_merge_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l,
                                 const float* __restrict__ diag, int nparts, int n,
                                 const float* __restrict__ scale_ptr, float* __restrict__ lse,
                                 float* __restrict__ out_max, float* __restrict__ out_l,
                                 float* __restrict__ out_lg2l, float* __restrict__ out_q,
                                 float* __restrict__ out_loss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float scale_log2 = __ldg(scale_ptr) * kLog2e;
  float m = -INFINITY;
  for (int q = 0; q < nparts; ++q) m = fmaxf(m, part_m[(size_t)q * n + i]);
  const float m2 = m * scale_log2;
  float l = 0.f;
  if (m != -INFINITY) {
    for (int q = 0; q < nparts; ++q) {
      const float mq = part_m[(size_t)q * n + i];
      if (mq != -INFINITY) l += part_l[(size_t)q * n + i] * exp2f(mq * scale_log2 - m2);
    }
  }
  float t = 0.f, gap = 0.f;  // gap = k (max - diag) >= 0
  if (diag) {
    gap = scale_log2 * (m - diag[i]);
    t = exp2f(-gap);
  }
  const float tot = l + t;
  const float lg = (t == 1.0f) ? log1pf(l) * kLog2e : log2f(tot);
  // ln(sum_j exp(S_ij - S_ii)) = log1p(l * 2^gap): accurate for tiny losses even when `max` is only
  // an upper reference (fused column statistics); falls back when 2^gap would overflow
  const float loss2 = (diag && gap < 64.f) ? log1pf(l * exp2f(gap)) * kLog2e : lg + gap;
  if (out_max) out_max[i] = m;
  if (out_l) out_l[i] = l;
  if (out_lg2l) out_lg2l[i] = lg - fmaf(scale_log2, m, -m2);
  if (out_q) out_q[i] = l / tot;
  if (out_loss) out_loss[i] = loss2 * kLn2;
  if (lse) lse[i] = (m2 + lg) * kLn2;
}

// Fused forward: merge the per-row-block column partials (log2-domain reference, l) of one column
// in fixed order and convert to the (raw max, l) convention of lse_merge_kernel.
constexpr int CM_SLICES = 16;   // row-block slices per block
__device__ __forceinline__ void col_merge_step(float& M, float& L, float rr, float ll) {
  if (rr != -INFINITY) {
    const float Mn = fmaxf(M, rr);
    L = L * exp2f(M - Mn) + ll * exp2f(rr - Mn);   // exp2f(-inf) = 0 on the first term
    M = Mn;
  }
}
__global__ void __launch_bounds__(32 * CM_SLICES)
col_merge_kernel(const float* __restrict__ col_ref, const float* __restrict__ col_l,
                 int n_row_blocks, size_t stride, int n_cols,
                 const float* __restrict__ scale_ptr, float* __restrict__ out_max,
                 float* __restrict__ out_l) {
  // block = 128 columns (4 per lane, 16-byte loads) x CM_SLICES row-block slices; the slices are
  // combined through smem in fixed order.  `stride` and the column count are multiples of 128
  // in the workspace layout, so the vector loads never leave the partial arrays.
  __shared__ float sm[CM_SLICES][128], sl[CM_SLICES][128];
  const int j0 = blockIdx.x * 128 + threadIdx.x * 4;
  const int slice = threadIdx.y;
  const float scale_log2 = __ldg(scale_ptr) * kLog2e;
  float M[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, L[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int r = slice; r < n_row_blocks; r += CM_SLICES) {
    const float4 rr = *reinterpret_cast<const float4*>(col_ref + (size_t)r * stride + j0);
    const float4 ll = *reinterpret_cast<const float4*>(col_l + (size_t)r * stride + j0);
    col_merge_step(M[0], L[0], rr.x, ll.x);
    col_merge_step(M[1], L[1], rr.y, ll.y);
    col_merge_step(M[2], L[2], rr.z, ll.z);
    col_merge_step(M[3], L[3], rr.w, ll.w);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    sm[slice][threadIdx.x * 4 + e] = M[e];
    sl[slice][threadIdx.x * 4 + e] = L[e];
  }
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;
  const int j = blockIdx.x * 128 + t;
  if (t < 128 && j < n_cols) {
    float Mt = -INFINITY;
#pragma unroll
    for (int q = 0; q < CM_SLICES; ++q) Mt = fmaxf(Mt, sm[q][t]);
    float Lt = 0.f;
    if (Mt != -INFINITY) {
#pragma unroll
      for (int q = 0; q < CM_SLICES; ++q)
        if (sm[q][t] != -INFINITY) Lt += sl[q][t] * exp2f(sm[q][t] - Mt);
    }
    const float mx = Mt / scale_log2;                 // surrogate "max" in raw cosine units
    out_max[j] = mx;
    out_l[j] = (Mt != -INFINITY) ? Lt * exp2f(Mt - mx * scale_log2) : 0.f;   // relative to fl(k * mx)
  }
}

// retrieval: rank[i] = number of columns ranked before the positive pair of row i
__global__ void rank_merge_kernel(const int* __restrict__ part_cnt, int nparts, int n, int* __restrict__ rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = 0;
  for (int q = 0; q < nparts; ++q) c += part_cnt[(size_t)q * n + i];
  rank[i] = c;
}

// retrieval: merge the partial top-RK lists of a row (each sorted) into its k best columns
__global__ void topk_merge_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx,
                                  int nparts, int n, int k, int* __restrict__ out_idx,
                                  float* __restrict__ out_val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float tv[RK];
  int ti[RK];
#pragma unroll
  for (int e = 0; e < RK; ++e) {
    tv[e] = -INFINITY;
    ti[e] = 0x7fffffff;
  }
  for (int q = 0; q < nparts; ++q) {
    const float* pv = part_val + ((size_t)q * n + i) * RK;
    const int* pi = part_idx + ((size_t)q * n + i) * RK;
    for (int c = 0; c < RK; ++c) {
      const float sv = pv[c];
      const int sc = pi[c];
      if (!rk_better(sv, sc, tv[RK - 1], ti[RK - 1])) break;   // the rest of this list ranks lower still
      tv[RK - 1] = sv;
      ti[RK - 1] = sc;
#pragma unroll
      for (int e = RK - 1; e > 0; --e) {
        if (rk_better(tv[e], ti[e], tv[e - 1], ti[e - 1])) {
          const float fv = tv[e];
          tv[e] = tv[e - 1];
          tv[e - 1] = fv;
          const int fi = ti[e];
          ti[e] = ti[e - 1];
          ti[e - 1] = fi;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < RK; ++e)
    if (e < k) {
      out_idx[(size_t)i * k + e] = ti[e] == 0x7fffffff ? -1 : ti[e];
      if (out_val) out_val[(size_t)i * k + e] = tv[e];
    }
}

// out2[0] = sum(row_loss), out2[1] = sum(col_loss); single block, fixed order => reproducible
__global__ void loss_reduce_kernel(const float* __restrict__ row_loss,
                                   const float* __restrict__ col_loss, int n,
                                   float* __restrict__ out2) {
  __shared__ double sh[2][1024];
  double a = 0.0, b = 0.0;
  // 8 independent loads in flight per thread; the per-thread order is still fixed
  int i = threadIdx.x;
  for (; i + 7 * (int)blockDim.x < n; i += 8 * blockDim.x) {
    float ra[8], rb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      ra[u] = row_loss ? row_loss[i + u * blockDim.x] : 0.f;
      rb[u] = col_loss ? col_loss[i + u * blockDim.x] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a += (double)ra[u];
      b += (double)rb[u];
    }
  }
  for (; i < n; i += blockDim.x) {
    if (row_loss) a += (double)row_loss[i];
    if (col_loss) b += (double)col_loss[i];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out2[0] = (float)sh[0][0];
    out2[1] = (float)sh[1][0];
  }
}

// s = min(exp(l), 100) and ds/dl (= exp(l) while unclamped, else 0), evaluated in double like the
// reference's float64 logit_scale parameter (VisionLanguageModule.py:111, :456-457)
__global__ void scale_prep_kernel(const void* __restrict__ logit_scale, int is_f64,
                                  float* __restrict__ scale, float* __restrict__ dscale_dls) {
  const double l = is_f64 ? *reinterpret_cast<const double*>(logit_scale)
                          : (double)*reinterpret_cast<const float*>(logit_scale);
  const double e = exp(l);
  scale[0] = (float)(e > 100.0 ? 100.0 : e);
  dscale_dls[0] = e <= 100.0 ? (float)e : 0.f;   // NaN propagates through both
  if (e != e) scale[0] = (float)e;
}

// (image_loss, text_loss) = sums / N, loss = (image_loss + text_loss) / 2   (reference :550-552)
__global__ void loss_finish_kernel(const float* __restrict__ sums2, float inv_n,
                                   float* __restrict__ out3) {
  const float il = sums2[0] * inv_n, tl = sums2[1] * inv_n;
  out3[0] = (il + tl) * 0.5f;
  out3[1] = il;
  out3[2] = tl;
}

__global__ void cast_bf16_to_f16_kernel(const __nv_bfloat16* __restrict__ src,
                                        __half* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = __float2half_rn(__bfloat162float(src[i]));
}

// fp32 embeddings -> both operand copies in one pass: bf16 (forward sweep) and the fp16 image of the
// bf16-ROUNDED value (backward GEMMs: the recompute must see exactly the forward's operands)
__global__ void cast_f32_operands_kernel(const float4* __restrict__ src, uint2* __restrict__ dst_bf16,
                                         uint2* __restrict__ dst_f16, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = src[i];
    const __nv_bfloat162 b0 = __floats2bfloat162_rn(v.x, v.y), b1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 ob;
    ob.x = *reinterpret_cast<const uint32_t*>(&b0);
    ob.y = *reinterpret_cast<const uint32_t*>(&b1);
    dst_bf16[i] = ob;
    if (dst_f16) {
      const __half2 h0 = __floats2half2_rn(__low2float(b0), __high2float(b0));
      const __half2 h1 = __floats2half2_rn(__low2float(b1), __high2float(b1));
      uint2 oh;
      oh.x = *reinterpret_cast<const uint32_t*>(&h0);
      oh.y = *reinterpret_cast<const uint32_t*>(&h1);
      dst_f16[i] = oh;
    }
  }
}

// ---- chunking policy: items = row_blocks * chunks should fill whole waves of SMs ----
static void pick_chunks(int n_row_blocks, int total_tiles, int n_sm, int* n_chunks,
                        int* tiles_per_chunk) {
  int best_c = 1;
  double best_eff = -1.0;
  const int max_c = total_tiles < 64 ? total_tiles : 64;
  for (int c = 1; c <= max_c; ++c) {
    const int tpc = (total_tiles + c - 1) / c;
    const int cc = (total_tiles + tpc - 1) / tpc;  // non-empty chunks
    if (cc != c) continue;
    const long items = (long)n_row_blocks * cc;
    const long waves = (items + n_sm - 1) / n_sm;
    // per-item fixed cost (X staging) ~ 1.5 tiles worth
    const double eff = (double)items * tpc / ((double)waves * n_sm * (tpc + 1.5));
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best_c = cc;
    }
  }
  *n_chunks = best_c;
  *tiles_per_chunk = (total_tiles + best_c - 1) / best_c;
  *n_chunks = (total_tiles + *tiles_per_chunk - 1) / *tiles_per_chunk;
}

static size_t lse_ws_bytes(int n_rows, int n_cols) {
  const int total_tiles = (n_cols + 127) / 128;
  const int max_chunks = total_tiles < 64 ? total_tiles : 64;
  return (size_t)2 * (size_t)(max_chunks * FWD_CG) * (size_t)n_rows * sizeof(float);
}

static size_t lse_fused_ws_bytes(int n_rows, int n_cols) {
  const size_t nrb = (n_rows + 127) / 128, nt = (n_cols + 127) / 128;
  return lse_ws_bytes(n_rows, n_cols) + 2 * nrb * nt * 128 * sizeof(float);
}

}  // namespace vlp

using namespace vlp;

extern "C" {

int vlpclip_version(void) { return VLPCLIP_VERSION; }
const char* vlpclip_last_error(void) { return err_buf(); }
int vlpclip_sm_count(void) { return sm_count(); }
unsigned long long vlpclip_launch_count(void) { return launch_counter(); }
int vlpclip_set_sm_limit(int n_sms) {
  sm_limit() = n_sms > 0 ? n_sms : 0;
  return usable_sms();
}

int vlpclip_cast_bf16_to_f16(const void* src, void* dst, size_t n, void* stream) {
  if (n == 0) return 0;
  if (!src || !dst) return fail(-1, "cast: null pointer");
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cast_bf16_to_f16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)src, (__half*)dst, n);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

int vlpclip_cast_f32_operands(const float* src, void* dst_bf16, void* dst_f16, size_t n, void* stream) {
  if (n == 0) return 0;
  if (!src || !dst_bf16) return fail(-1, "cast_f32_operands: null pointer");
  if (n % 4 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(dst_bf16) & 7) != 0 || (reinterpret_cast<uintptr_t>(dst_f16) & 7) != 0)
    return fail(-1, "cast_f32_operands: need n %% 4 == 0 and 16-byte aligned source (8-byte aligned outputs)");
  const size_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cast_f32_operands_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)src, (uint2*)dst_bf16,
                                                                     (uint2*)dst_f16, n4);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

size_t vlpclip_lse_workspace_bytes(int n_rows, int n_cols, int d) {
  (void)d;
  if (n_rows <= 0 || n_cols <= 0) return 0;
  return lse_ws_bytes(n_rows, n_cols);
}

static int lse_fwd_impl(const void* x, int ldx, const void* y, int ldy, int n_rows, int n_cols,
                        int d, const float* scale, int diag_shift, float* row_max, float* row_l,
                        float* diag,
                        float* col_max, float* col_l, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream, bool operand_f16 = false, const int* row_ids = nullptr,
                        const int* col_ids = nullptr) {
  const bool fused = col_max != nullptr;
  if (n_rows <= 0 || n_cols <= 0) return fail(-1, "lse_fwd: empty problem (%d x %d)", n_rows, n_cols);
  if (!x || !y || !scale || !row_max || !row_l || !diag || !workspace || (fused && !col_l))
    return fail(-1, "lse_fwd: null pointer");
  if (d <= 0 || d % 8 != 0 || d > 768)
    return fail(-1, "lse_fwd: embedding dim %d unsupported (need a multiple of 8, <= 768)", d);
  if (ldx % 8 != 0 || ldy % 8 != 0)
    return fail(-1, "lse_fwd: row strides must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0)
    return fail(-1, "lse_fwd: X must be 16-byte aligned");
  int rc = check_device_sm100();
  if (rc) return rc;
  const size_t need = fused ? lse_fused_ws_bytes(n_rows, n_cols) : lse_ws_bytes(n_rows, n_cols);
  if (workspace_bytes < need)
    return fail(-1, "lse_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);

  LseParams p = {};
  p.operand_f16 = operand_f16 ? 1 : 0;
  p.x = (const __nv_bfloat16*)x;
  p.ldx = ldx;
  p.n_rows = n_rows;
  p.n_cols = n_cols;
  p.d = d;
  p.kblocks = (d + 63) / 64;
  p.total_tiles = (n_cols + 127) / 128;
  p.n_row_blocks = (n_rows + 127) / 128;
  const int nsm = usable_sms();
  pick_chunks(p.n_row_blocks, p.total_tiles, nsm, &p.n_chunks, &p.tiles_per_chunk);
  p.diag_shift = diag_shift;
  p.scale_ptr = scale;
  if ((row_ids == nullptr) != (col_ids == nullptr))
    return fail(-1, "lse_fwd: row and column caption ids must be given together");
  if (col_ids && (reinterpret_cast<uintptr_t>(col_ids) & 15) != 0)
    return fail(-1, "lse_fwd: column caption ids must be 16-byte aligned");
  p.xid = row_ids;
  p.yid = col_ids;
  const int nparts = p.n_chunks * FWD_CG;
  p.part_m = (float*)workspace;
  p.part_l = p.part_m + (size_t)nparts * n_rows;
  p.diag = diag;
  const size_t col_stride = (size_t)p.total_tiles * 128;
  p.col_ref = nullptr;
  p.col_l = nullptr;
  if (fused) {
    p.col_ref = (float*)((uint8_t*)workspace + lse_ws_bytes(n_rows, n_cols));
    p.col_l = p.col_ref + (size_t)p.n_row_blocks * col_stride;
  }

  CUtensorMap map_y;
  rc = make_tmap_sw128(&map_y, y, 2, (uint64_t)d, (uint64_t)n_cols, (uint64_t)ldy, 128);
  if (rc) return rc;

  const size_t smem = FWD_STAGES * FWD_STAGE_BYTES + sizeof(FwdBarriers) + 1024;
  VLP_CUDA_OK(set_smem_attr_once((const void*)lse_partial_kernel<MODE_ROWS>, (int)smem, 2));
  VLP_CUDA_OK(set_smem_attr_once((const void*)lse_partial_kernel<MODE_COLS>, (int)smem, 3));
  const int n_items = p.n_row_blocks * p.n_chunks;
  const int grid = n_items < nsm ? n_items : nsm;
  if (fused)
    lse_partial_kernel<MODE_COLS><<<grid, FWD_THREADS, smem, stream>>>(map_y, p);
  else
    lse_partial_kernel<MODE_ROWS><<<grid, FWD_THREADS, smem, stream>>>(map_y, p);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  lse_merge_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(p.part_m, p.part_l, nullptr, nparts,
                                                             n_rows, scale, nullptr, row_max,
                                                             row_l, nullptr, nullptr, nullptr);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  if (fused) {
    col_merge_kernel<<<(n_cols + 127) / 128, dim3(32, CM_SLICES), 0, stream>>>(
        p.col_ref, p.col_l, p.n_row_blocks, col_stride, n_cols, scale, col_max, col_l);
    VLP_COUNT_LAUNCH(1);
    VLP_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// ---- retrieval metrics without the M x M matrix (VisionLanguageModule.py:364-439) ----------------
static size_t retrieval_ws_bytes(int n_rows, int n_cols, bool topk) {
  const int n_row_blocks = (n_rows + 127) / 128, total_tiles = (n_cols + 127) / 128;
  int nsm = usable_sms();
  if (nsm <= 0) nsm = 148;
  int n_chunks, tpc;
  pick_chunks(n_row_blocks, total_tiles, nsm, &n_chunks, &tpc);
  const size_t nparts = (size_t)n_chunks * FWD_CG;
  return nparts * (size_t)n_rows * (topk ? (size_t)RK * 8 : 4) + 256;
}

size_t vlpclip_retrieval_workspace_bytes(int n_rows, int n_cols, int d, int k) {
  (void)d;
  if (n_rows <= 0 || n_cols <= 0) return 0;
  return retrieval_ws_bytes(n_rows, n_cols, k > 0);
}

static int retrieval_impl(const void* x, int ldx, const void* y, int ldy, int n_rows, int n_cols, int d,
                          int k, int* out_rank, int* out_idx, float* out_val, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream) {
  const bool topk = k > 0;
  if (n_rows <= 0 || n_cols <= 0) return fail(-1, "retrieval: empty problem (%d x %d)", n_rows, n_cols);
  if (!x || !y || !workspace || (topk ? !out_idx : !out_rank)) return fail(-1, "retrieval: null pointer");
  if (topk && (k > RK || k > n_cols))
    return fail(-1, "retrieval: k = %d unsupported (need 1 <= k <= min(%d, n_cols))", k, RK);
  if (!topk && n_cols < n_rows)
    return fail(-1, "retrieval: row i is paired with column i: need n_cols >= n_rows (%d < %d)", n_cols, n_rows);
  if (d <= 0 || d % 8 != 0 || d > 768)
    return fail(-1, "retrieval: embedding dim %d unsupported (need a multiple of 8, <= 768)", d);
  if (ldx % 8 != 0 || ldy % 8 != 0) return fail(-1, "retrieval: row strides must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return fail(-1, "retrieval: X must be 16-byte aligned");
  int rc = check_device_sm100();
  if (rc) return rc;
  if (workspace_bytes < retrieval_ws_bytes(n_rows, n_cols, topk))
    return fail(-1, "retrieval: workspace too small (%zu < %zu)", workspace_bytes,
                retrieval_ws_bytes(n_rows, n_cols, topk));
  LseParams p = {};
  p.operand_f16 = 0;
  p.x = (const __nv_bfloat16*)x;
  p.ldx = ldx;
  p.n_rows = n_rows;
  p.n_cols = n_cols;
  p.d = d;
  p.kblocks = (d + 63) / 64;
  p.total_tiles = (n_cols + 127) / 128;
  p.n_row_blocks = (n_rows + 127) / 128;
  const int nsm = usable_sms();
  pick_chunks(p.n_row_blocks, p.total_tiles, nsm, &p.n_chunks, &p.tiles_per_chunk);
  p.diag_shift = 0;
  p.scale_ptr = nullptr;
  const int nparts = p.n_chunks * FWD_CG;
  if (topk) {
    p.part_val = (float*)workspace;
    p.part_idx = (int*)(p.part_val + (size_t)nparts * n_rows * RK);
  } else {
    p.part_cnt = (int*)workspace;
  }
  CUtensorMap map_y;
  rc = make_tmap_sw128(&map_y, y, 2, (uint64_t)d, (uint64_t)n_cols, (uint64_t)ldy, 128);
  if (rc) return rc;
  const size_t smem = FWD_STAGES * FWD_STAGE_BYTES + sizeof(FwdBarriers) + 1024;
  VLP_CUDA_OK(set_smem_attr_once((const void*)lse_partial_kernel<MODE_RANK>, (int)smem, 5));
  VLP_CUDA_OK(set_smem_attr_once((const void*)lse_partial_kernel<MODE_TOPK>, (int)smem, 6));
  const int n_items = p.n_row_blocks * p.n_chunks;
  const int grid = n_items < nsm ? n_items : nsm;
  if (topk) {
    lse_partial_kernel<MODE_TOPK><<<grid, FWD_THREADS, smem, stream>>>(map_y, p);
    VLP_COUNT_LAUNCH(1);
    topk_merge_kernel<<<(n_rows + 127) / 128, 128, 0, stream>>>(p.part_val, p.part_idx, nparts, n_rows, k,
                                                                out_idx, out_val);
  } else {
    lse_partial_kernel<MODE_RANK><<<grid, FWD_THREADS, smem, stream>>>(map_y, p);
    VLP_COUNT_LAUNCH(1);
    rank_merge_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(p.part_cnt, nparts, n_rows, out_rank);
  }
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

int vlpclip_retrieval_ranks(const void* q_bf16, int ldq, const void* k_bf16, int ldk, int n_rows,
                            int n_cols, int d, int* rank, void* workspace, size_t workspace_bytes,
                            void* stream) {
  return retrieval_impl(q_bf16, ldq, k_bf16, ldk, n_rows, n_cols, d, 0, rank, nullptr, nullptr, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

int vlpclip_retrieval_topk(const void* q_bf16, int ldq, const void* k_bf16, int ldk, int n_rows,
                           int n_cols, int d, int k, int* idx, float* val, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (k <= 0) return fail(-1, "retrieval_topk: k must be positive");
  return retrieval_impl(q_bf16, ldq, k_bf16, ldk, n_rows, n_cols, d, k, nullptr, idx, val, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

int vlpclip_lse_fwd(const void* x, int ldx, const void* y, int ldy, int n_rows, int n_cols, int d,
                    const float* scale, int diag_shift, float* row_max, float* row_l, float* diag,
                    void* workspace, size_t workspace_bytes, void* stream_) {
  return lse_fwd_impl(x, ldx, y, ldy, n_rows, n_cols, d, scale, diag_shift, row_max, row_l, diag,
                      nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream_);
}

size_t vlpclip_lse_fused_workspace_bytes(int n_rows, int n_cols, int d) {
  (void)d;
  if (n_rows <= 0 || n_cols <= 0) return 0;
  return lse_fused_ws_bytes(n_rows, n_cols);
}

int vlpclip_lse_fwd_fused(const void* x, int ldx, const void* y, int ldy, int n_rows, int n_cols,
                          int d, const float* scale, int diag_shift, float* row_max, float* row_l,
                          float* diag, float* col_max, float* col_l, void* workspace,
                          size_t workspace_bytes, void* stream_) {
  if (!col_max || !col_l) return fail(-1, "lse_fwd_fused: null column outputs");
  return lse_fwd_impl(x, ldx, y, ldy, n_rows, n_cols, d, scale, diag_shift, row_max, row_l, diag,
                      col_max, col_l, workspace, workspace_bytes, (cudaStream_t)stream_);
}

int vlpclip_lse_fwd_fused_masked(const void* x, int ldx, const void* y, int ldy, int n_rows, int n_cols,
                                 int d, int operand_f16, const float* scale, int diag_shift,
                                 const int* row_ids, const int* col_ids, float* row_max, float* row_l,
                                 float* diag, float* col_max, float* col_l, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  if (!col_max || !col_l) return fail(-1, "lse_fwd_fused_masked: null column outputs");
  if (!row_ids || !col_ids) return fail(-1, "lse_fwd_fused_masked: null caption ids");
  return lse_fwd_impl(x, ldx, y, ldy, n_rows, n_cols, d, scale, diag_shift, row_max, row_l, diag,
                      col_max, col_l, workspace, workspace_bytes, (cudaStream_t)stream_,
                      operand_f16 != 0, row_ids, col_ids);
}

int vlpclip_lse_fwd_f16(const void* x, int ldx, const void* y, int ldy, int n_rows, int n_cols,
                        int d, const float* scale, int diag_shift, float* row_max, float* row_l,
                        float* diag, void* workspace, size_t workspace_bytes, void* stream_) {
  return lse_fwd_impl(x, ldx, y, ldy, n_rows, n_cols, d, scale, diag_shift, row_max, row_l, diag,
                      nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream_, true);
}

int vlpclip_lse_fwd_fused_f16(const void* x, int ldx, const void* y, int ldy, int n_rows, int n_cols,
                              int d, const float* scale, int diag_shift, float* row_max,
                              float* row_l, float* diag, float* col_max, float* col_l,
                              void* workspace, size_t workspace_bytes, void* stream_) {
  if (!col_max || !col_l) return fail(-1, "lse_fwd_fused: null column outputs");
  return lse_fwd_impl(x, ldx, y, ldy, n_rows, n_cols, d, scale, diag_shift, row_max, row_l, diag,
                      col_max, col_l, workspace, workspace_bytes, (cudaStream_t)stream_, true);
}

// bf16 -> fp16 cast of a local shard, stored into n_dst destinations at once: with dsts[r] pointing
// into rank r's gather window (NVLink peer memory) at this rank's row offset, the cast IS the
// all-gather -- every store instruction writes full 128-byte lines.
constexpr int PUSH_MAX_DST = 8;
struct PushDst {
  void* p[PUSH_MAX_DST];
};
__global__ void cast_push_f16_kernel(const uint4* __restrict__ src, size_t n8, const PushDst dst,
                                     int n_dst) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 v = src[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
      const __half2 h = __floats2half2_rn(__bfloat162float(b.x), __bfloat162float(b.y));
      o[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    const uint4 out = make_uint4(o[0], o[1], o[2], o[3]);
    for (int r = 0; r < n_dst; ++r) reinterpret_cast<uint4*>(dst.p[r])[i] = out;
  }
}

int vlpclip_cast_push_f16(const void* src_bf16, size_t n_elems, void* const* dsts, int n_dst,
                          void* stream) {
  if (!src_bf16 || !dsts || n_dst <= 0 || n_dst > PUSH_MAX_DST || n_elems == 0 || n_elems % 8 != 0)
    return fail(-1, "cast_push: bad arguments (n_elems %zu, n_dst %d)", n_elems, n_dst);
  if ((reinterpret_cast<uintptr_t>(src_bf16) & 15) != 0)
    return fail(-1, "cast_push: source must be 16-byte aligned");
  PushDst d = {};
  for (int r = 0; r < n_dst; ++r) {
    if (!dsts[r] || (reinterpret_cast<uintptr_t>(dsts[r]) & 15) != 0)
      return fail(-1, "cast_push: destination %d is null or not 16-byte aligned", r);
    d.p[r] = dsts[r];
  }
  const size_t n8 = n_elems / 8;
  int blocks = (int)((n8 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cast_push_f16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src_bf16, n8, d, n_dst);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

int vlpclip_lse_merge(const float* part_max, const float* part_l, const float* diag, int nparts,
                      int n, const float* scale, float* lse, float* out_max, float* out_l,
                      float* out_lg2l,
                      float* out_q, float* out_loss, void* stream) {
  if (n <= 0 || nparts <= 0) return fail(-1, "lse_merge: empty input");
  if (!part_max || !part_l || !scale) return fail(-1, "lse_merge: null pointer");
  if (out_loss && !diag) return fail(-1, "lse_merge: per-row loss needs the positive-pair logits");
  lse_merge_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      part_max, part_l, diag, nparts, n, scale, lse, out_max, out_l, out_lg2l, out_q,
      out_loss);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

int vlpclip_loss_reduce(const float* row_loss, const float* col_loss, int n, float* out2,
                        void* stream) {
  if (n <= 0) return fail(-1, "loss_reduce: empty input");
  if (!out2) return fail(-1, "loss_reduce: null pointer");
  loss_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_loss, col_loss, n, out2);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

int vlpclip_scale_prep(const void* logit_scale, int is_f64, float* scale, float* dscale_dls,
                       void* stream) {
  if (!logit_scale || !scale || !dscale_dls) return fail(-1, "scale_prep: null pointer");
  scale_prep_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(logit_scale, is_f64, scale, dscale_dls);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

int vlpclip_loss_finish(const float* sums2, int n_global, float* out3, void* stream) {
  if (!sums2 || !out3 || n_global <= 0) return fail(-1, "loss_finish: bad arguments");
  loss_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums2, 1.0f / (float)n_global, out3);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
