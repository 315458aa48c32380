// prologue.cu -- projection + L2-normalise of both embedding streams and its backward.
//
// Reference: VisionLanguageModule.py:448-449 (`features @ projection`, x @ W convention) and
// :452-453 (`F.normalize`, eps 1e-12).  The GEMMs are small (0.8 % of the head's flops at
// N = 32k) and HBM-bound (K <= 512: 128 MB of operand + result traffic per 17 GFLOP), so ONE
// tcgen05 kernel serves all of them (gemm2_tf32_kernel: fp32 operands, kind::tf32 (10-bit mantissa),
// fp32 accumulation), reading every operand in place -- a "transposed" operand is the other
// major-ness of the UMMA descriptor, never a copy.  The forward is GEMM + one row pass that
// normalises and emits the fp32 embedding (returned to the caller), its bf16 copy (operand of the
// loss kernels) and its fp16 copy (operand of the backward GEMMs).  (Round 1's variant with the
// normalisation fused into a 128 x 512 single-tile epilogue measured 103 us per stream against
// 41 + 29 us for this pair at 32768 x 512: its epilogue did not overlap any MMA.)
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "../../include/vlpclip.h"

namespace vlp {

// -------------------------------------------------------------------------------------------------
// gemm2_tf32_kernel -- the plain GEMMs of the head (projection, d features, dW):
//     C[M, N] = A * B,   A given as [M][K] (K-major) or [K][M] (MN-major), B as [N][K] or [K][N]
// so that no operand is ever transposed in memory (round 2: the two transposes of dW = feat^T du
// cost more than the GEMM).  Persistent CTAs walk over (128 x 256 tile, K split) items; the two
// 256-column halves of TMEM are two accumulators, so the epilogue of one item (thread == row,
// TMEM -> global) overlaps the MMAs of the next; 4-stage ring of {A 16 KB, B 32 KB} per 32-deep
// K slab.  MN-major slabs land as 4 KB boxes of [32 k rows][32 elements = 128 B] in the 128B swizzle
// with 32-byte atoms (TMA SWIZZLE_128B_ATOM_32B = descriptor layout type 1, the only layout the tensor
// cores take for MN-major 32-bit operands): one kind::tf32 MMA (K = 8) reads 8 rows (1 KB) of every
// box; boxes are LBO = 4 KB apart, 4-row groups SBO = 512 B.
// -------------------------------------------------------------------------------------------------
constexpr int G2_THREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int G2_STAGES = 4;
constexpr int G2_A_BYTES = 128 * 128;
constexpr int G2_B_BYTES = 256 * 128;
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_NT = 256;        // columns per item

struct Gemm2Params {
  int m, n, k;
  int k_per_split;      // multiple of 32
  int n_splits;
  int m_tiles, n_tiles;
  float* c;             // [m, n] row-major, or slab `split` of [n_splits][m, n] partials
  int ldc;
  size_t split_stride;  // elements between partial slabs (0: no split)
};

struct G2Barriers {
  uint64_t full[G2_STAGES];
  uint64_t empty[G2_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

template <bool kAMn, bool kBMn>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2_tf32_kernel(const __grid_constant__ CUtensorMap map_a,   // K-major: box {32 k, 128 rows}; MN-major: box {32 m, 32 k}
                  const __grid_constant__ CUtensorMap map_b,   // K-major: box {32 k, 256 rows}; MN-major: box {32 n, 32 k}
                  const Gemm2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  G2Barriers* bars = reinterpret_cast<G2Barriers*>(smem + G2_STAGES * G2_STAGE_BYTES);
  const uint32_t ring = smem_u32(smem);
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < G2_STAGES; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->acc_full[i]), 1);
      mbar_init(smem_u32(&bars->acc_empty[i]), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(smem_u32(&bars->tmem_base), 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  // items: the n tiles of one m tile are neighbours (they share the A slabs in L2); K splits outermost
  const int n_items = p.m_tiles * p.n_tiles * p.n_splits;

  if (warp == 0) {
    // ================= TMA producer: one elected thread =================
    if (elect_one()) {
      uint32_t st = 0, par = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int split = item / (p.m_tiles * p.n_tiles);
        const int mn = item - split * (p.m_tiles * p.n_tiles);
        const int m0 = (mn / p.n_tiles) * 128, n0 = (mn % p.n_tiles) * G2_NT;
        const int k0 = split * p.k_per_split;
        const int k1 = min(p.k, k0 + p.k_per_split);
        const int n_cols = min(G2_NT, p.n - n0);
        for (int kk = k0; kk < k1; kk += 32) {
          mbar_wait(smem_u32(&bars->empty[st]), par ^ 1);
          const uint32_t sa = ring + st * G2_STAGE_BYTES, sb = sa + G2_A_BYTES;
          const uint32_t fb = smem_u32(&bars->full[st]);
          const int b_boxes = kBMn ? (n_cols + 31) / 32 : 1;
          mbar_expect_tx(fb, G2_A_BYTES + (kBMn ? b_boxes * 4096 : G2_B_BYTES));
          if (kAMn) {
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_load_2d(sa + c * 4096, &map_a, fb, m0 + c * 32, kk);
          } else {
            tma_load_2d(sa, &map_a, fb, kk, m0);
          }
          if (kBMn) {
            for (int c = 0; c < b_boxes; ++c) tma_load_2d(sb + c * 4096, &map_b, fb, n0 + c * 32, kk);
          } else {
            tma_load_2d(sb, &map_b, fb, kk, n0);
          }
          if (++st == G2_STAGES) {
            st = 0;
            par ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer: one elected thread =================
    if (elect_one()) {
      const uint32_t a_hi = kAMn ? sdesc_hi_sw128_base32(512) : sdesc_hi_sw128(1024);
      const uint32_t b_hi = kBMn ? sdesc_hi_sw128_base32(512) : sdesc_hi_sw128(1024);
      uint32_t st = 0, par = 0, it_ctr = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it_ctr) {
        const int split = item / (p.m_tiles * p.n_tiles);
        const int mn = item - split * (p.m_tiles * p.n_tiles);
        const int n0 = (mn % p.n_tiles) * G2_NT;
        const int k0 = split * p.k_per_split;
        const int k1 = min(p.k, k0 + p.k_per_split);
        const int n_pad = (min(G2_NT, p.n - n0) + 15) & ~15;
        const uint32_t idesc = make_idesc(UMMA_TF32, UMMA_TF32, kAMn ? MAJOR_MN : MAJOR_K,
                                          kBMn ? MAJOR_MN : MAJOR_K, 128, n_pad);
        const uint32_t acc = it_ctr & 1, use = it_ctr >> 1;
        mbar_wait(smem_u32(&bars->acc_empty[acc]), (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + acc * 256;
        for (int kk = k0; kk < k1; kk += 32) {
          mbar_wait(smem_u32(&bars->full[st]), par);
          const uint32_t sa = ring + st * G2_STAGE_BYTES, sb = sa + G2_A_BYTES;
          // K-major: +32 B per K = 8 inside the 128-byte rows; MN-major: +1 KB (8 k rows), boxes LBO apart
          const uint32_t a_lo = sdesc_lo_sw128(sa, kAMn ? 4096 : 0), b_lo = sdesc_lo_sw128(sb, kBMn ? 4096 : 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss_tf32_w(d_tmem, a_lo + ks * (kAMn ? 64 : 2), a_hi, b_lo + ks * (kBMn ? 64 : 2), b_hi, idesc,
                           (kk != k0 || ks != 0) ? 1u : 0u);
          umma_commit<1>(smem_u32(&bars->empty[st]));
          if (++st == G2_STAGES) {
            st = 0;
            par ^= 1;
          }
        }
        umma_commit<1>(smem_u32(&bars->acc_full[acc]));
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: thread == output row =================
    const uint32_t quarter = warp & 3;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    uint32_t it_ctr = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it_ctr) {
      const int split = item / (p.m_tiles * p.n_tiles);
      const int mn = item - split * (p.m_tiles * p.n_tiles);
      const int m0 = (mn / p.n_tiles) * 128, n0 = (mn % p.n_tiles) * G2_NT;
      const int n_cols = min(G2_NT, p.n - n0);
      const int n_pad = (n_cols + 15) & ~15;
      const uint32_t acc = it_ctr & 1, use = it_ctr >> 1;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.m;
      float* orow = p.c + (size_t)split * p.split_stride + (row_ok ? (size_t)row : 0) * p.ldc + n0;
      const bool vec = ((reinterpret_cast<uintptr_t>(orow) & 15) == 0);
      mbar_wait(smem_u32(&bars->acc_full[acc]), use & 1);
      tc_fence_after();
      for (int c = 0; c < n_pad; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem + lane_addr + acc * 256 + c, v);   // warp-collective: never under a lane predicate
        tmem_ld_wait();
        if (row_ok) {
          if (vec && c + 16 <= n_cols) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(orow + c + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                              __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c + j < n_cols) orow[c + j] = __uint_as_float(v[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[acc]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

// du = (dE - E * <E, dE>) * inv_norm, one warp per row
__global__ void normalize_bwd_kernel(const float* __restrict__ emb, const float* __restrict__ d_emb,
                                     const float* __restrict__ inv_norm, int n, int d,
                                     float* __restrict__ du) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* e = emb + (size_t)row * d;
  const float* g = d_emb + (size_t)row * d;
  float dot = 0.f;
  for (int c = lane; c < d; c += 32) dot = fmaf(e[c], g[c], dot);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  const float inv = inv_norm[row];
  float* o = du + (size_t)row * d;
  for (int c = lane; c < d; c += 32) o[c] = (g[c] - e[c] * dot) * inv;
}

// emb = u / max(||u||, 1e-12) in place (u = emb_f32 on entry), plus the bf16 / fp16 operand copies
__global__ void normalize_fwd_kernel(float* __restrict__ emb, __nv_bfloat16* __restrict__ emb_bf16,
                                     __half* __restrict__ emb_f16, float* __restrict__ inv_norm,
                                     int n, int d) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float4* u4 = reinterpret_cast<float4*>(emb + (size_t)row * d);   // d % 8 == 0
  const int d4 = d >> 2;
  float ss = 0.f;
  for (int c = lane; c < d4; c += 32) {
    const float4 x = u4[c];
    ss = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, ss))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0) inv_norm[row] = inv;
  uint2* ob = reinterpret_cast<uint2*>(emb_bf16 + (size_t)row * d);
  uint2* oh = reinterpret_cast<uint2*>(emb_f16 + (size_t)row * d);
  for (int c = lane; c < d4; c += 32) {
    float4 x = u4[c];
    x.x *= inv;
    x.y *= inv;
    x.z *= inv;
    x.w *= inv;
    u4[c] = x;
    const __nv_bfloat162 b0 = __floats2bfloat162_rn(x.x, x.y), b1 = __floats2bfloat162_rn(x.z, x.w);
    // the fp16 copy is the fp16 image of the bf16-ROUNDED value, so that the backward recompute sees
    // exactly the operands of the forward
    const __half2 h0 = __floats2half2_rn(__low2float(b0), __high2float(b0));
    const __half2 h1 = __floats2half2_rn(__low2float(b1), __high2float(b1));
    uint2 pb, ph;
    pb.x = *reinterpret_cast<const uint32_t*>(&b0);
    pb.y = *reinterpret_cast<const uint32_t*>(&b1);
    ph.x = *reinterpret_cast<const uint32_t*>(&h0);
    ph.y = *reinterpret_cast<const uint32_t*>(&h1);
    ob[c] = pb;
    oh[c] = ph;
  }
}

static size_t pg_align(size_t x) { return (x + 255) & ~size_t(255); }

// c[i] = sum over the split-K partial slabs in split order (fixed order => bit-reproducible)
__global__ void splitk_reduce_kernel(const float4* __restrict__ part, int n_splits, size_t stride4,
                                     size_t n4, float4* __restrict__ c) {
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
    float4 a = part[i];
    for (int s = 1; s < n_splits; ++s) {
      const float4 b = part[(size_t)s * stride4 + i];
      a.x += b.x;
      a.y += b.y;
      a.z += b.z;
      a.w += b.w;
    }
    c[i] = a;
  }
}

// tiling of gemm2: K per split and number of splits (split-K only when the output has few tiles)
static void gemm2_plan(int m, int n, int k, int* kps_out, int* splits_out) {
  const int tiles = ((m + 127) / 128) * ((n + G2_NT - 1) / G2_NT);
  int splits = 1;
  int nsm = sm_count();
  if (nsm <= 0) nsm = 148;
  if (tiles < nsm / 2 && k >= 1024) {
    splits = nsm / tiles;
    const int max_splits = k / 512;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int kps = (k + splits - 1) / splits;
  kps = (kps + 31) & ~31;
  *kps_out = kps;
  *splits_out = (k + kps - 1) / kps;
}

// C[m,n] = A * B.  a_mn: A is given as [k][m] (row stride lda), else [m][k]; b_mn: B is [k][n], else [n][k].
// `slabs`: workspace for the split-K partials (gemm2_plan says whether it is needed).
static int launch_gemm2(const float* a, int lda, bool a_mn, const float* b, int ldb, bool b_mn, int m, int n,
                        int k, float* c, int ldc, float* slabs, cudaStream_t stream) {
  if (lda % 4 != 0 || ldb % 4 != 0)
    return fail(-1, "gemm: operand row strides (%d, %d) must be multiples of 4 floats (16 bytes)", lda, ldb);
  CUtensorMap map_a, map_b;
  int rc = a_mn ? make_tmap_sw128(&map_a, a, 4, (uint64_t)m, (uint64_t)k, (uint64_t)lda, 32, false, true)
                : make_tmap_sw128(&map_a, a, 4, (uint64_t)k, (uint64_t)m, (uint64_t)lda, 128);
  if (rc) return rc;
  rc = b_mn ? make_tmap_sw128(&map_b, b, 4, (uint64_t)n, (uint64_t)k, (uint64_t)ldb, 32, false, true)
            : make_tmap_sw128(&map_b, b, 4, (uint64_t)k, (uint64_t)n, (uint64_t)ldb, 256);
  if (rc) return rc;
  Gemm2Params p = {};
  p.m = m;
  p.n = n;
  p.k = k;
  gemm2_plan(m, n, k, &p.k_per_split, &p.n_splits);
  if (!slabs) {   // no workspace for partials (the forward projection: K = feature width): one split
    p.k_per_split = (k + 31) & ~31;
    p.n_splits = 1;
  }
  p.m_tiles = (m + 127) / 128;
  p.n_tiles = (n + G2_NT - 1) / G2_NT;
  p.ldc = ldc;
  if (p.n_splits > 1) {
    if (ldc != n) return fail(-1, "gemm: split-K needs a dense C");
    p.c = slabs;
    p.split_stride = (size_t)m * n;
  } else {
    p.c = c;
    p.split_stride = 0;
  }
  const size_t smem = G2_STAGES * G2_STAGE_BYTES + sizeof(G2Barriers) + 1024;
  const void* fn = a_mn ? (b_mn ? (const void*)gemm2_tf32_kernel<true, true> : (const void*)gemm2_tf32_kernel<true, false>)
                        : (b_mn ? (const void*)gemm2_tf32_kernel<false, true> : (const void*)gemm2_tf32_kernel<false, false>);
  VLP_CUDA_OK(set_smem_attr_once(fn, (int)smem, 8 + (a_mn ? 2 : 0) + (b_mn ? 1 : 0)));
  int nsm = sm_count();
  if (nsm <= 0) nsm = 148;
  const int n_items = p.m_tiles * p.n_tiles * p.n_splits;
  const int grid = n_items < nsm ? n_items : nsm;
  if (a_mn) {
    if (b_mn) gemm2_tf32_kernel<true, true><<<grid, G2_THREADS, smem, stream>>>(map_a, map_b, p);
    else gemm2_tf32_kernel<true, false><<<grid, G2_THREADS, smem, stream>>>(map_a, map_b, p);
  } else {
    if (b_mn) gemm2_tf32_kernel<false, true><<<grid, G2_THREADS, smem, stream>>>(map_a, map_b, p);
    else gemm2_tf32_kernel<false, false><<<grid, G2_THREADS, smem, stream>>>(map_a, map_b, p);
  }
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  if (p.n_splits > 1) {
    const size_t mn = (size_t)m * n;
    if (mn % 4 != 0 || (reinterpret_cast<uintptr_t>(c) & 15) != 0)
      return fail(-1, "gemm: split-K needs m*n %% 4 == 0 and a 16-byte aligned C");
    int blocks = (int)((mn / 4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>((const float4*)slabs, p.n_splits, mn / 4, mn / 4, (float4*)c);
    VLP_COUNT_LAUNCH(1);
    VLP_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

}  // namespace vlp

using namespace vlp;

extern "C" {

size_t vlpclip_project_workspace_bytes(int n, int f, int d) {
  (void)n;
  if (f <= 0 || d <= 0) return 0;
  return 256;   // (kept in the ABI: the operands are read in place, nothing is staged any more)
}

int vlpclip_project_normalize_fwd(const float* feat, const float* w, int n, int f, int d,
                                  float* emb_f32, void* emb_bf16, void* emb_f16, float* inv_norm,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n <= 0 || f <= 0 || d <= 0) return fail(-1, "project: empty problem");
  if (!feat || !w || !emb_f32 || !emb_bf16 || !emb_f16 || !inv_norm || !workspace)
    return fail(-1, "project: null pointer");
  if (d % 8 != 0 || d > 768)
    return fail(-1, "project: embedding dim %d unsupported (multiple of 8, <= 768)", d);
  if (f % 4 != 0) return fail(-1, "project: feature dim %d must be a multiple of 4", f);
  int rc = check_device_sm100();
  if (rc) return rc;
  if (workspace_bytes < vlpclip_project_workspace_bytes(n, f, d))
    return fail(-1, "project: workspace too small");
  // u = feat W straight from the row-major operands (W [f][d] is the MN-major B operand), then one
  // row-normalise pass that also emits the bf16 / fp16 operand copies
  (void)workspace;
  int rc2 = launch_gemm2(feat, f, false, w, d, true, n, d, f, emb_f32, d, nullptr, stream);
  if (rc2) return rc2;
  normalize_fwd_kernel<<<(n + 7) / 8, 256, 0, stream>>>(emb_f32, (__nv_bfloat16*)emb_bf16,
                                                        (__half*)emb_f16, inv_norm, n, d);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

int vlpclip_normalize_bwd(const float* emb_f32, const float* d_emb, const float* inv_norm, int n,
                          int d, float* du, void* stream) {
  if (n <= 0 || d <= 0) return fail(-1, "normalize_bwd: empty problem");
  if (!emb_f32 || !d_emb || !inv_norm || !du) return fail(-1, "normalize_bwd: null pointer");
  normalize_bwd_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(emb_f32, d_emb, inv_norm, n,
                                                                      d, du);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

size_t vlpclip_gemm_workspace_bytes(int m, int n, int k) {
  if (m <= 0 || n <= 0 || k <= 0) return 0;
  int kps, splits;
  gemm2_plan(m, n, k, &kps, &splits);
  return 256 + (splits > 1 ? pg_align((size_t)splits * m * n * 4) : 0);   // split-K partial slabs only
}

// C[m,n] = op(A) op(B): A is [m][k] (trans_a = 0) or [k][m] (trans_a = 1), B is [k][n] (trans_b = 0) or
// [n][k] (trans_b = 1), all row-major fp32.  Neither operand is copied: a "transposed" operand is
// simply the other major-ness of the UMMA descriptor.  The contiguous extent of each operand must be
// a multiple of 4 floats (TMA row stride); the batch size never is one (it is M of d features and K of
// dW, and the reference's sampler yields remainder batches of arbitrary size).
int vlpclip_gemm_tf32(const float* a, const float* b, float* c, int m, int n, int k, int trans_a,
                      int trans_b, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (m <= 0 || n <= 0 || k <= 0) return fail(-1, "gemm: empty problem");
  if (!a || !b || !c || !workspace) return fail(-1, "gemm: null pointer");
  int rc = check_device_sm100();
  if (rc) return rc;
  if (workspace_bytes < vlpclip_gemm_workspace_bytes(m, n, k))
    return fail(-1, "gemm: workspace too small");
  float* slabs = (float*)(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  return launch_gemm2(a, trans_a ? m : k, trans_a != 0, b, trans_b ? k : n, trans_b == 0, m, n, k, c, n,
                      slabs, stream);
}

}  // extern "C"
