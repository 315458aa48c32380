// grad_bwd_quad.cuh -- staged variant of the backward (-DVLP_BWD_QUAD, included by grad_bwd.cu):
// the producer / consumer pipeline of grad_pair_kernel on clusters of FOUR CTAs with
// cta_group::2 MMAs, so that every SM stages only half of each Y tile.
//
//   rank 0 = P0 (producer, leader)   rank 1 = P1 (producer)     row blocks 2u / 2u + 1 of unit u
//   rank 2 = C0 (consumer, leader)   rank 3 = C1 (consumer)
//
//   producers: S[256 x 128] = [X_2u; X_2u+1] Y_t^T as ONE M = 256 MMA per k-step issued by P0 (TS
//              form, each CTA's X block in its own TMEM); each producer stages its 64 of the tile's
//              128 columns ([64 q x 64 k] boxes); the S rows of a CTA's row block land in its own
//              TMEM, its softmax warps form the fp16 G tile exactly as in the pair kernel and push
//              it to "its" consumer (rank + 2).
//   consumers: acc[256 x d] += [G_2u; G_2u+1] Y_t as M = 256, N = 256 MMAs issued by C0: each consumer
//              stages 128 of the 256 accumulator columns of an instruction ([128 q x 64 d] boxes);
//              C1's idle MMA warp forwards "my G tile arrived" to C0.
//   Barriers the issuing warp waits on live in the pair's leader and collect both CTAs; barriers it
//   signals are multicast tcgen05.commit arrivals (mask = the CTAs that wait on them).
// Per SM and tile this streams 64 KB of Y instead of 128 KB (the Y streams are
// bytes-in-flight bound).  Work is cut stream-K style over (row-block pair, column tile).
// Restrictions of the experiment build: d a multiple of 128, d <= 512.
#pragma once

namespace vlp {

constexpr int QP_KB_PER_STAGE = 4;                       // producer stage: 4 boxes [64 q x 64 k] = 32 KB
constexpr int QP_BOX_BYTES = 8192;
constexpr int QP_STAGE_BYTES = QP_KB_PER_STAGE * QP_BOX_BYTES;
constexpr int QP_STAGES = P_RING_BYTES / QP_STAGE_BYTES;
constexpr int QC_BOX_BYTES = 16384;                      // consumer box [128 q x 64 d]
constexpr int QC_STAGE_BYTES = 2 * QC_BOX_BYTES;         // its 128 accumulator columns of one MMA group
constexpr int QC_STAGES = C_RING_BYTES / QC_STAGE_BYTES;
constexpr int Q_RING_BARS = QP_STAGES > QC_STAGES ? QP_STAGES : QC_STAGES;

struct QuadBarriers {
  uint64_t full[Q_RING_BARS];    // leader of the pair: TMA bytes of both CTAs
  uint64_t empty[Q_RING_BARS];   // every CTA: multicast commit of its pair's leader
  uint64_t s_full[2];            // producers: multicast commit of P0
  uint64_t s_empty[2];           // P0: softmax warps of P0 and P1
  uint64_t x_ready;              // P0: staging warps of P0 and P1
  uint64_t x_free;               // producers: multicast commit of P0
  uint64_t g_full[G_SLOTS];      // consumers: armed remotely by "their" producer's push
  uint64_t g_pair[G_SLOTS];      // C0: C1 reports that its G tile has arrived
  uint64_t g_empty[G_SLOTS];     // producers: multicast commit of C0
  uint64_t acc_full;             // consumers: multicast commit of C0
  uint64_t acc_free;             // C0: epilogue warps of C0 and C1
  uint32_t tmem_base;
};
static_assert(sizeof(QuadBarriers) <= BAR_BYTES, "barrier block");

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(BWD_THREADS, 1)
grad_quad_kernel(const __grid_constant__ CUtensorMap map_p,   // box {64 k, 64 q}
                 const __grid_constant__ CUtensorMap map_c,   // box {64 d, 128 q}
                 const GradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  QuadBarriers* bars = reinterpret_cast<QuadBarriers*>(smem + G_SLOTS * G_SLOT_BYTES);
  const uint32_t gslots = smem_u32(smem);
  const uint32_t ring = gslots + G_SLOTS * G_SLOT_BYTES + BAR_BYTES;
  const uint32_t stage = ring + C_RING_BYTES;   // consumer only
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pr = rank & 1;                  // position inside the MMA pair
  const bool is_prod = rank < 2;
  const bool leader = pr == 0;
  const uint32_t leader_rank = rank & 2;         // 0 for producers, 2 for consumers
  const uint16_t pair_mask = is_prod ? 0x3 : 0xC;
  const int cluster_id = blockIdx.x >> 2;
  const int n_clusters = gridDim.x >> 2;
  const int n_units = (p.n_row_blocks + 1) >> 1;
#ifdef VLP_PROFILE_WAITS
  long long wait_cyc[16] = {0};
  const long long kernel_t0 = clock64();
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < Q_RING_BARS; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->s_empty[i]), 2 * SMX_WARPS);
    }
    for (int i = 0; i < G_SLOTS; ++i) {
      mbar_init(smem_u32(&bars->g_full[i]), 1);
      mbar_init(smem_u32(&bars->g_pair[i]), 1);
      mbar_init(smem_u32(&bars->g_empty[i]), 1);
    }
    mbar_init(smem_u32(&bars->x_ready), 2 * 8);
    mbar_init(smem_u32(&bars->x_free), 1);
    mbar_init(smem_u32(&bars->acc_full), 1);
    mbar_init(smem_u32(&bars->acc_free), 2 * EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2>(smem_u32(&bars->tmem_base), 512);   // both CTAs of each MMA pair
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_p);
    tma_prefetch_desc(&map_c);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const float scale_dev = __ldg(p.scale_ptr);
  const float scale_log2 = scale_dev * kLog2e;
  const uint32_t tmem_s_col = 512u - 2u * 128u;   // two S buffers (d <= 512)

  // arrive on a barrier the pair's issuing warp waits on (it lives in the pair's leader)
  auto arrive_leader = [&](uint64_t* bar) {
    if (leader)
      mbar_arrive(smem_u32(bar));
    else
      mbar_arrive_cluster(mapa_shared(smem_u32(bar), leader_rank));
  };

  if (is_prod) {
    // =====================================================================================
    // producer pair: S tiles (one M = 256 MMA for both row blocks) + softmax -> G tiles
    // =====================================================================================
    if (warp == 0) {
      uint32_t it = 0;
      WorkRange work(cluster_id, n_clusters, n_units, p.total_tiles);
      Segment sg;
      while (work.next(sg)) {
        for (int t = sg.t0; t < sg.t1; ++t)
          for (int kb = 0; kb < p.kblocks; kb += QP_KB_PER_STAGE, ++it) {
            const uint32_t st = it % QP_STAGES, ph = (it / QP_STAGES) & 1;
            const int nkb = min(QP_KB_PER_STAGE, p.kblocks - kb);
            VLP_WAIT(0, mbar_wait_cluster(smem_u32(&bars->empty[st]), ph ^ 1));
            if (elect_one()) {
              if (leader) mbar_expect_tx(smem_u32(&bars->full[st]), 2 * nkb * QP_BOX_BYTES);
              for (int q = 0; q < nkb; ++q)
                tma_load_2d_pair(ring + st * QP_STAGE_BYTES + q * QP_BOX_BYTES, &map_p,
                                 smem_u32(&bars->full[st]), (kb + q) * 64, t * 128 + pr * 64);
            }
            __syncwarp();
          }
      }
      // tail: the multicast commits that free the last stages must have landed in this CTA
      const uint32_t last = it < (uint32_t)QP_STAGES ? it : (uint32_t)QP_STAGES;
      for (uint32_t k = 0; k < last; ++k) {
        const uint32_t j = it - 1 - k;
        mbar_wait_cluster(smem_u32(&bars->empty[j % QP_STAGES]), (j / QP_STAGES) & 1);
      }
    } else if (warp == 1) {
      if (leader) {
        const uint32_t idesc = make_idesc(UMMA_F16, UMMA_F16, MAJOR_K, MAJOR_K, 256, 128);
        uint32_t it = 0, tile_ctr = 0, item_ctr = 0;
        WorkRange work(cluster_id, n_clusters, n_units, p.total_tiles);
        Segment sg;
        for (; work.next(sg); ++item_ctr) {
          VLP_WAIT(1, mbar_wait_cluster(smem_u32(&bars->x_ready), item_ctr & 1));
          tc_fence_after();
          for (int t = sg.t0; t < sg.t1; ++t, ++tile_ctr) {
            const uint32_t buf = tile_ctr & 1, use = tile_ctr >> 1;
            VLP_WAIT(2, mbar_wait_cluster(smem_u32(&bars->s_empty[buf]), (use & 1) ^ 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem + tmem_s_col + buf * 128;
            for (int kb = 0; kb < p.kblocks; kb += QP_KB_PER_STAGE, ++it) {
              const uint32_t st = it % QP_STAGES, ph = (it / QP_STAGES) & 1;
              const int nkb = min(QP_KB_PER_STAGE, p.kblocks - kb);
              VLP_WAIT(3, mbar_wait(smem_u32(&bars->full[st]), ph));
              tc_fence_after();
              if (elect_one()) {
                for (int q = 0; q < nkb; ++q) {
                  const uint32_t sb = ring + st * QP_STAGE_BYTES + q * QP_BOX_BYTES;
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_ts<2>(d_tmem, tmem + BWD_TMEM_X + (kb + q) * 32 + ks * 8,
                               make_sdesc_sw128(sb + ks * 32, 0, 1024), idesc, (kb | q | ks) != 0);
                }
                umma_commit_mcast<2>(smem_u32(&bars->empty[st]), 0x3);
              }
              __syncwarp();
            }
            if (elect_one()) umma_commit_mcast<2>(smem_u32(&bars->s_full[buf]), 0x3);
            __syncwarp();
          }
          if (elect_one()) umma_commit_mcast<2>(smem_u32(&bars->x_free), 0x3);
          __syncwarp();
        }
      }
    } else {
      // ---- softmax warps (both producers; same arithmetic as grad_pair_kernel) ----
      const uint32_t quarter = warp & 3;
      const uint32_t grp = (warp - 2) >> 2;          // 64-column group of the S tile
      const uint32_t row_in_blk = quarter * 32 + lane;
      const uint32_t lane_addr = (quarter * 32u) << 16;
      const int dp = p.kblocks * 64;
      const uint32_t sw = row_in_blk & 7;
      const uint32_t peer = rank + 2;                // the consumer this producer feeds
      uint32_t tile_ctr = 0, item_ctr = 0;
      double ds_total = 0.0;
      WorkRange work(cluster_id, n_clusters, n_units, p.total_tiles);
      Segment sg;
      for (; work.next(sg); ++item_ctr) {
        const int rb = sg.rb * 2 + (int)pr;          // may be one past the last row block
        const int row = rb * 128 + row_in_blk;
        const bool row_ok = row < p.n_rows;
        if (item_ctr > 0) {
          VLP_WAIT(4, mbar_wait_cluster(smem_u32(&bars->x_free), (item_ctr - 1) & 1));
          tc_fence_after();
        }
        {   // all 8 warps stage the X block (two K halves) into this CTA's TMEM
          const int k_begin = grp * (dp / 2);
          const uint4* src =
              reinterpret_cast<const uint4*>(p.x + (size_t)(row_ok ? row : 0) * p.ldx);
          if (kXUnroll && dp == 512) {
            // all 32 loads of the thread's half row in flight at once: one L2 round trip instead of
            // eight dependent ones while the tensor pipe waits for the new X block
            uint4 w[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const int k = k_begin + q * 8;
              w[q] = (row_ok && k < p.d) ? __ldg(src + (k >> 3)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              uint32_t v[16];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[q * 4 + 0] = w[g * 4 + q].x;
                v[q * 4 + 1] = w[g * 4 + q].y;
                v[q * 4 + 2] = w[g * 4 + q].z;
                v[q * 4 + 3] = w[g * 4 + q].w;
              }
              tmem_st_x16(tmem + lane_addr + BWD_TMEM_X + k_begin / 2 + g * 16, v);
            }
          } else
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
          if (lane == 0) arrive_leader(&bars->x_ready);
        }
        const float xmax = p.xmax[row];  // statistics are padded to whole row-block PAIRS
        const float xlg = p.xlg[row];
        const float xr = p.xr[row];
        const bool fast = *p.fast_flag != 0;
        const int dcol = row_ok ? row - p.diag_shift : -1000000000;
        float ds_acc = 0.f;

        for (int t = sg.t0; t < sg.t1; ++t, ++tile_ctr) {
          const uint32_t buf = tile_ctr & 1, use = tile_ctr >> 1;
          VLP_WAIT(5, mbar_wait_cluster(smem_u32(&bars->s_full[buf]), use & 1));
          tc_fence_after();
          const uint32_t slot = tile_ctr % G_SLOTS, slot_use = tile_ctr / G_SLOTS;
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
          if (lane == 0) arrive_leader(&bars->s_empty[buf]);

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
            float acc = 0.f;
            if (any_diag)
              softmax_tile_fast<true>(v, yc4, xmax, xlg - 13.f, xr, scale_log2, diag_val * G_SCALE,
                                      diag_j, out, acc);
            else
              softmax_tile_fast<false>(v, yc4, xmax, xlg - 13.f, xr, scale_log2, 0.f, diag_j, out,
                                       acc);
            ds_acc = fmaf(acc, 1.0f / G_SCALE, ds_acc);
          } else if (any_diag) {
            softmax_tile<true>(v, ymax4, ylg4, xmax, xlg, scale_log2, diag_val, diag_j, out, ds_acc);
          } else {
            softmax_tile<false>(v, ymax4, ylg4, xmax, xlg, scale_log2, 0.f, diag_j, out, ds_acc);
          }

          if (slot_use > 0)
            VLP_WAIT(6, mbar_wait_cluster(smem_u32(&bars->g_empty[slot]), (slot_use - 1) & 1));
          const uint32_t dst = gslots + slot * G_SLOT_BYTES + grp * 16384 + row_in_blk * 128;
#pragma unroll
          for (int c = 0; c < SMX_COLS / 8; ++c) {
            const uint32_t a = dst + ((c ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(out[c * 4 + 0]),
                         "r"(out[c * 4 + 1]), "r"(out[c * 4 + 2]), "r"(out[c * 4 + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          VLP_WAIT(7, bar_sync(1, 32 * SMX_WARPS));
          if (warp == 2 && lane < (uint32_t)PUSH_SPLIT) {
            constexpr uint32_t kChunk = G_SLOT_BYTES / PUSH_SPLIT;
            const uint32_t rbar = mapa_shared(smem_u32(&bars->g_full[slot]), peer);
            const uint32_t src = gslots + slot * G_SLOT_BYTES + lane * kChunk;
            const uint32_t rdst = mapa_shared(src, peer);
            if (lane == 0)
              asm volatile(
                  "mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(
                      rbar),
                  "r"(G_SLOT_BYTES)
                  : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], "
                "%2, [%3];" ::"r"(rdst),
                "r"(src), "r"(kChunk), "r"(rbar)
                : "memory");
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, o);
        ds_total += (double)ds_acc;
      }
      if (p.ds_part != nullptr && lane == 0)
        p.ds_part[((size_t)cluster_id * 2 + pr) * SMX_WARPS + (warp - 2)] = (float)ds_total;
      // drain: the consumers must have released every slot we pushed, and the last multicast
      // commit of P0 (x_free) must have landed, before this CTA may exit
      for (uint32_t back = 0; back < (uint32_t)G_SLOTS && back < tile_ctr; ++back) {
        const uint32_t tc = tile_ctr - 1 - back;
        mbar_wait_cluster(smem_u32(&bars->g_empty[tc % G_SLOTS]), (tc / G_SLOTS) & 1);
      }
      if (item_ctr > 0) mbar_wait_cluster(smem_u32(&bars->x_free), (item_ctr - 1) & 1);
    }
  } else {
    // =====================================================================================
    // consumer pair: dX blocks of both row blocks accumulate in TMEM (M = 256 MMAs issued by C0)
    // =====================================================================================
    const int n_nc = (p.ndb + 3) / 4;  // 256-wide accumulator chunks of this pass
    if (warp == 0) {
      uint32_t it = 0;
      WorkRange work(cluster_id, n_clusters, n_units, p.total_tiles);
      Segment sg;
      while (work.next(sg)) {
        for (int t = sg.t0; t < sg.t1; ++t)
          for (int nc = 0; nc < n_nc; ++nc, ++it) {
            const int nb = min(4, p.ndb - nc * 4) >> 1;   // 64-column blocks staged by THIS consumer
            const uint32_t st = it % QC_STAGES, ph = (it / QC_STAGES) & 1;
            VLP_WAIT(8, mbar_wait_cluster(smem_u32(&bars->empty[st]), ph ^ 1));
            if (elect_one()) {
              if (leader) mbar_expect_tx(smem_u32(&bars->full[st]), 2 * nb * QC_BOX_BYTES);
              for (int b = 0; b < nb; ++b)
                tma_load_2d_pair(ring + st * QC_STAGE_BYTES + b * QC_BOX_BYTES, &map_c,
                                 smem_u32(&bars->full[st]),
                                 (p.db0 + nc * 4 + (int)pr * nb + b) * 64, t * 128);
            }
            __syncwarp();
          }
      }
      const uint32_t last = it < (uint32_t)QC_STAGES ? it : (uint32_t)QC_STAGES;
      for (uint32_t k = 0; k < last; ++k) {
        const uint32_t j = it - 1 - k;
        mbar_wait_cluster(smem_u32(&bars->empty[j % QC_STAGES]), (j / QC_STAGES) & 1);
      }
    } else if (warp == 1) {
      uint32_t it = 0, tile_ctr = 0, item_ctr = 0;
      WorkRange work(cluster_id, n_clusters, n_units, p.total_tiles);
      Segment sg;
      if (leader) {
        for (; work.next(sg); ++item_ctr) {
          if (item_ctr > 0) {
            VLP_WAIT(9, mbar_wait_cluster(smem_u32(&bars->acc_free), (item_ctr - 1) & 1));
            tc_fence_after();
          }
          for (int t = sg.t0; t < sg.t1; ++t, ++tile_ctr) {
            const uint32_t slot = tile_ctr % G_SLOTS, gph = (tile_ctr / G_SLOTS) & 1;
            VLP_WAIT(10, mbar_wait_cluster(smem_u32(&bars->g_full[slot]), gph));
            VLP_WAIT(10, mbar_wait_cluster(smem_u32(&bars->g_pair[slot]), gph));
            tc_fence_after();
            const uint32_t ga = gslots + slot * G_SLOT_BYTES;
            for (int nc = 0; nc < n_nc; ++nc, ++it) {
              const int nbt = min(4, p.ndb - nc * 4);   // 64-column blocks of this MMA group (even)
              const uint32_t idesc = make_idesc(UMMA_F16, UMMA_F16, MAJOR_K, MAJOR_MN, 256, nbt * 64);
              const uint32_t st = it % QC_STAGES, ph = (it / QC_STAGES) & 1;
              VLP_WAIT(11, mbar_wait(smem_u32(&bars->full[st]), ph));
              tc_fence_after();
              if (elect_one()) {
                const uint32_t sb = ring + st * QC_STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < 8; ++i) {   // K = 128 logit columns of the tile, 16 per step
                  const uint64_t ad = make_sdesc_sw128(ga + (i >> 2) * 16384 + (i & 3) * 32, 0, 1024);
                  const uint64_t bd = make_sdesc_sw128(sb + i * 2048, QC_BOX_BYTES, 1024);
                  umma_ss<2>(tmem + nc * 256, ad, bd, idesc, !(t == sg.t0 && i == 0));
                }
                umma_commit_mcast<2>(smem_u32(&bars->empty[st]), 0xC);
              }
              __syncwarp();
            }
            // release the G slot in both producers
            if (elect_one()) umma_commit_mcast<2>(smem_u32(&bars->g_empty[slot]), 0x3);
            __syncwarp();
          }
          if (elect_one()) umma_commit_mcast<2>(smem_u32(&bars->acc_full), 0xC);
          __syncwarp();
        }
      } else {
        // C1: tell C0 when the G tile of THIS row block has arrived (bulk copy of P1)
        const uint32_t remote = mapa_shared(smem_u32(&bars->g_pair[0]), 2);
        while (work.next(sg))
          for (int t = sg.t0; t < sg.t1; ++t, ++tile_ctr) {
            const uint32_t slot = tile_ctr % G_SLOTS;
            mbar_wait_cluster(smem_u32(&bars->g_full[slot]), (tile_ctr / G_SLOTS) & 1);
            if (lane == 0) mbar_arrive_cluster(remote + slot * 8);
            __syncwarp();
          }
      }
    } else if (warp < 2 + EPI_WARPS) {
      // ---- epilogue (both consumers): own TMEM accumulator -> global ----
      const uint32_t quarter = warp & 3;
      const uint32_t lane_addr = (quarter * 32u) << 16;
      uint32_t item_ctr = 0;
      WorkRange work(cluster_id, n_clusters, n_units, p.total_tiles);
      Segment sg;
      for (; work.next(sg); ++item_ctr) {
        VLP_WAIT(12, mbar_wait_cluster(smem_u32(&bars->acc_full), item_ctr & 1));
        tc_fence_after();
        const int rb = sg.rb * 2 + (int)pr;
        const bool final_out = sg.slot < 0;
        const float mulv =
            scale_dev * ((final_out && p.out_mul) ? p.out_scale * __ldg(p.out_mul) : p.out_scale);
        const int sub = lane >> 3, ch = lane & 7;
        uint8_t* orow8[8];
        bool ok8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + sub;
          const int grow = rb * 128 + (int)quarter * 32 + r;
          ok8[i] = grow < p.n_rows;
          if (final_out) {
            uint8_t* base;
            const size_t rr = scatter_row(p.scatter, ok8[i] ? grow : 0, base, p.dx);
            orow8[i] = base + rr * p.d * (p.dx_bf16 ? 2 : 4);
          } else {
            orow8[i] = reinterpret_cast<uint8_t*>(
                p.part + ((size_t)((cluster_id * 2 + sg.slot) * 2 + (int)pr) * 128 + quarter * 32 + r) * p.d);
          }
        }
        const bool as_bf16 = final_out && p.dx_bf16;
        const uint32_t stg = stage + (warp - 2) * 4096;
        const int cbase = p.db0 * 64;
        const int cc_per = p.ndb * 64 / EPI_PARTS;          // columns of this warp's share
        const int cc0 = (int)((warp - 2) >> 2) * cc_per;
        for (int cc = cc0; cc < cc0 + cc_per; cc += 32) {
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
        if (lane == 0) arrive_leader(&bars->acc_free);
      }
    }
  }

#ifdef VLP_PROFILE_WAITS
  if (p.wait_prof != nullptr && lane == 0 && leader && warp <= 2) {
    long long* o = p.wait_prof + (size_t)cluster_id * 16;   // P0 and C0 report (disjoint indices)
    for (int i = 0; i < 12; ++i)
      if (wait_cyc[i] != 0) o[i] = wait_cyc[i];
    if (warp == 0) o[12 + (is_prod ? 0 : 1)] = clock64() - kernel_t0;
  }
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc<2>(tmem, 512);
}

}  // namespace vlp
