// prologue.cu -- projection + L2-normalise of both embedding streams and its backward.
//
// Reference: VisionLanguageModule.py:448-449 (`features @ projection`, x @ W convention) and
// :452-453 (`F.normalize`, eps 1e-12).  The GEMMs are small (0.8 % of the head's flops at
// N = 32k) so one simple tcgen05 kernel serves all of them:
//     C[M, N] (+)= A[M, K] * B[N, K]^T     fp32 operands, kind::tf32 (10-bit mantissa), fp32 acc
// with both operands K-major (operands that are not K-major in memory are transposed into the
// workspace first).  One CTA owns a [128 x <=512] output tile whose accumulator fills the 512
// TMEM columns; the forward projection fuses the row L2-norm into its epilogue (thread == row, so
// the norm is a thread-local reduction over the TMEM row) and emits the fp32 embedding (returned
// to the caller), its bf16 copy (operand of the loss kernels) and its fp16 copy (operand of the
// backward GEMMs).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "../../include/vlpclip.h"

namespace vlp {

constexpr int PG_THREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int PG_STAGES = 2;
constexpr int PG_A_BYTES = 128 * 128;       // 128 rows x 32 fp32
constexpr int PG_B_BYTES = 512 * 128;       // up to 512 rows x 32 fp32
constexpr int PG_STAGE_BYTES = PG_A_BYTES + PG_B_BYTES;

struct GemmParams {
  int m, n, k;          // problem
  int n_tile;           // columns per CTA (<= 512, multiple of 16)
  int k_per_split;      // multiple of 32
  int n_splits;
  float* c;             // [m, n] row-major (split-K: slab `blockIdx.z` of [n_splits][m, n] partials)
  int ldc;
  size_t split_stride;  // elements between the partial slabs of split-K (0: no split)
  // fused normalise epilogue (n_tile covers all of n, no split-K)
  int normalize;
  float* emb_f32;
  __nv_bfloat16* emb_bf16;
  __half* emb_f16;
  float* inv_norm;
};

struct PgBarriers {
  uint64_t full[PG_STAGES];
  uint64_t empty[PG_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(PG_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a,   // box {32 k, 128 rows}
                 const __grid_constant__ CUtensorMap map_b,   // box {32 k, 256 rows}
                 const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  PgBarriers* bars = reinterpret_cast<PgBarriers*>(smem + PG_STAGES * PG_STAGE_BYTES);
  const uint32_t ring = smem_u32(smem);
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

  const int m0 = blockIdx.x * 128;
  const int n0 = blockIdx.y * p.n_tile;
  const int k0 = blockIdx.z * p.k_per_split;
  const int k1 = min(p.k, k0 + p.k_per_split);
  const int kiters = (k1 - k0 + 31) / 32;
  const int n_cols = min(p.n_tile, p.n - n0);          // valid columns of this tile
  const int n_pad = (n_cols + 15) & ~15;                // UMMA N granularity
  const int n_hi = n_pad > 256 ? n_pad - 256 : 0;       // second instruction's N
  const int n_lo = n_pad > 256 ? 256 : n_pad;

  if (threadIdx.x == 0) {
    for (int i = 0; i < PG_STAGES; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), 1);
    }
    mbar_init(smem_u32(&bars->acc_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(smem_u32(&bars->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // (TMA and MMA are issued by converged warps under elect.sync: a lane-predicated region makes
  //  ptxas wrap every UTMALDG / UTCHMMA in a uniformisation loop)
  if (warp == 0) {
    for (int it = 0; it < kiters; ++it) {
      const uint32_t st = it % PG_STAGES, ph = (it / PG_STAGES) & 1;
      mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1);
      if (elect_one()) {
        const uint32_t sa = ring + st * PG_STAGE_BYTES, sb = sa + PG_A_BYTES;
        const uint32_t bytes = PG_A_BYTES + (n_pad > 256 ? 2 : 1) * 256 * 128;
        mbar_expect_tx(smem_u32(&bars->full[st]), bytes);
        tma_load_2d(sa, &map_a, smem_u32(&bars->full[st]), k0 + it * 32, m0);
        tma_load_2d(sb, &map_b, smem_u32(&bars->full[st]), k0 + it * 32, n0);
        if (n_pad > 256)
          tma_load_2d(sb + 256 * 128, &map_b, smem_u32(&bars->full[st]), k0 + it * 32, n0 + 256);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc_lo = make_idesc(UMMA_TF32, UMMA_TF32, MAJOR_K, MAJOR_K, 128, n_lo);
    const uint32_t idesc_hi = make_idesc(UMMA_TF32, UMMA_TF32, MAJOR_K, MAJOR_K, 128, n_hi);
    for (int it = 0; it < kiters; ++it) {
      const uint32_t st = it % PG_STAGES, ph = (it / PG_STAGES) & 1;
      mbar_wait(smem_u32(&bars->full[st]), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = ring + st * PG_STAGE_BYTES, sb = sa + PG_A_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // 4 x (K = 8 fp32 = 32 B)
          const uint64_t ad = make_sdesc_sw128(sa + ks * 32, 0, 1024);
          umma_ss_tf32(tmem, ad, make_sdesc_sw128(sb + ks * 32, 0, 1024), idesc_lo, (it | ks) != 0);
          if (n_hi > 0)
            umma_ss_tf32(tmem + 256, ad, make_sdesc_sw128(sb + 256 * 128 + ks * 32, 0, 1024),
                         idesc_hi, (it | ks) != 0);
        }
        umma_commit<1>(smem_u32(&bars->empty[st]));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit<1>(smem_u32(&bars->acc_full));
    __syncwarp();
  } else {
    // ---- epilogue: thread == output row ----
    const uint32_t quarter = warp & 3;
    const int row = m0 + quarter * 32 + lane;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    mbar_wait(smem_u32(&bars->acc_full), 0);
    tc_fence_after();
    if (p.normalize) {
      float ss = 0.f;
      for (int c = 0; c < n_pad; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem + lane_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float x = (c + j < n_cols) ? __uint_as_float(v[j]) : 0.f;
          ss = fmaf(x, x, ss);
        }
      }
      const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);   // F.normalize eps
      const bool row_ok = row < p.m;
      const size_t rsafe = row_ok ? (size_t)row : 0;
      if (row_ok) p.inv_norm[row] = inv;
      {
        float* o32 = p.emb_f32 + rsafe * p.n;
        __nv_bfloat16* ob = p.emb_bf16 + rsafe * p.n;
        __half* oh = p.emb_f16 + rsafe * p.n;
        for (int c = 0; c < n_pad; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(tmem + lane_addr + c, v);   // warp-collective: never under a lane predicate
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            if (row_ok && c + j < n_cols) {   // n % 8 == 0
              float e[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) e[q] = __uint_as_float(v[j + q]) * inv;
              *reinterpret_cast<float4*>(o32 + c + j) = make_float4(e[0], e[1], e[2], e[3]);
              *reinterpret_cast<float4*>(o32 + c + j + 4) = make_float4(e[4], e[5], e[6], e[7]);
              __nv_bfloat162 b0 = __floats2bfloat162_rn(e[0], e[1]);
              __nv_bfloat162 b1 = __floats2bfloat162_rn(e[2], e[3]);
              __nv_bfloat162 b2 = __floats2bfloat162_rn(e[4], e[5]);
              __nv_bfloat162 b3 = __floats2bfloat162_rn(e[6], e[7]);
              uint4 pb;
              pb.x = *reinterpret_cast<uint32_t*>(&b0);
              pb.y = *reinterpret_cast<uint32_t*>(&b1);
              pb.z = *reinterpret_cast<uint32_t*>(&b2);
              pb.w = *reinterpret_cast<uint32_t*>(&b3);
              *reinterpret_cast<uint4*>(ob + c + j) = pb;
              // the fp16 copy is the fp16 image of the bf16-ROUNDED value, so that the backward
              // recompute sees exactly the operands of the forward
              __half2 h0 = __floats2half2_rn(__low2float(b0), __high2float(b0));
              __half2 h1 = __floats2half2_rn(__low2float(b1), __high2float(b1));
              __half2 h2 = __floats2half2_rn(__low2float(b2), __high2float(b2));
              __half2 h3 = __floats2half2_rn(__low2float(b3), __high2float(b3));
              uint4 ph;
              ph.x = *reinterpret_cast<uint32_t*>(&h0);
              ph.y = *reinterpret_cast<uint32_t*>(&h1);
              ph.z = *reinterpret_cast<uint32_t*>(&h2);
              ph.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(oh + c + j) = ph;
            }
          }
        }
      }
    } else {
      const bool row_ok = row < p.m;
      // split-K partials go to their own slab and are summed in split order afterwards
      // (splitk_reduce_kernel): no atomics, so dW is bit-reproducible
      float* orow = p.c + (size_t)blockIdx.z * p.split_stride + (row_ok ? (size_t)row : 0) * p.ldc + n0;
      for (int c = 0; c < n_pad; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(tmem + lane_addr + c, v);   // warp-collective: never under a lane predicate
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (row_ok && c + j < n_cols) orow[c + j] = __uint_as_float(v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

// out[c][r] = in[r][c]; rows of `out` are `ldo` >= rows floats apart (the pad is never read: the
// tensor maps give the true K extent and TMA zero-fills beyond it)
__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int rows,
                                     int cols, int ldo) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * ldo + r] = tile[threadIdx.x][i];
  }
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

// emb = u / max(||u||, 1e-12) in place (u = emb_f32 on entry), plus the bf16 / fp16 operand copies;
// used when the embedding dim exceeds the 512 accumulator columns of the fused epilogue
__global__ void normalize_fwd_kernel(float* __restrict__ emb, __nv_bfloat16* __restrict__ emb_bf16,
                                     __half* __restrict__ emb_f16, float* __restrict__ inv_norm,
                                     int n, int d) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float* u = emb + (size_t)row * d;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) ss = fmaf(u[c], u[c], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0) inv_norm[row] = inv;
  for (int c = lane; c < d; c += 32) {
    const float e = u[c] * inv;
    const __nv_bfloat16 b = __float2bfloat16_rn(e);
    u[c] = e;
    emb_bf16[(size_t)row * d + c] = b;
    emb_f16[(size_t)row * d + c] = __float2half_rn(__bfloat162float(b));
  }
}

static size_t pg_align(size_t x) { return (x + 255) & ~size_t(255); }

// C[m,n] = A[m,k] * B[n,k]^T with both operands K-major fp32 (row strides lda / ldb floats, multiples of 4)
static int launch_gemm_kmajor(const float* a, int lda, const float* b, int ldb, int m, int n, int k,
                              GemmParams p, cudaStream_t stream) {
  if (lda % 4 != 0 || ldb % 4 != 0)
    return fail(-1, "gemm: operand row strides (%d, %d) must be multiples of 4 floats (16 bytes)", lda, ldb);
  CUtensorMap map_a, map_b;
  int rc = make_tmap_sw128(&map_a, a, 4, (uint64_t)k, (uint64_t)m, (uint64_t)lda, 128);
  if (rc) return rc;
  rc = make_tmap_sw128(&map_b, b, 4, (uint64_t)k, (uint64_t)n, (uint64_t)ldb, 256);
  if (rc) return rc;
  p.m = m;
  p.n = n;
  p.k = k;
  const size_t smem = PG_STAGES * PG_STAGE_BYTES + sizeof(PgBarriers) + 1024;
  VLP_CUDA_OK(set_smem_attr_once((const void*)gemm_tf32_kernel, (int)smem, 4));
  dim3 grid((m + 127) / 128, (n + p.n_tile - 1) / p.n_tile, p.n_splits);
  gemm_tf32_kernel<<<grid, PG_THREADS, smem, stream>>>(map_a, map_b, p);
  VLP_COUNT_LAUNCH(1);
  VLP_CUDA_OK(cudaGetLastError());
  return 0;
}

static void launch_transpose(const float* in, float* out, int rows, int cols, int ldo, cudaStream_t stream) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_f32_kernel<<<grid, block, 0, stream>>>(in, out, rows, cols, ldo);
  VLP_COUNT_LAUNCH(1);
}

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

// tiling of the generic GEMM: columns per CTA, K per split, number of splits
static void gemm_plan(int m, int n, int k, int* n_tile, int* kps_out, int* splits_out) {
  *n_tile = n >= 512 ? 512 : ((n + 15) & ~15);
  const int tiles = ((m + 127) / 128) * ((n + *n_tile - 1) / *n_tile);
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

}  // namespace vlp

using namespace vlp;

extern "C" {

size_t vlpclip_project_workspace_bytes(int n, int f, int d) {
  (void)n;
  if (f <= 0 || d <= 0) return 0;
  return pg_align((size_t)f * d * sizeof(float));
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
  float* wt = (float*)workspace;                 // W^T [d][f]: K-major B operand
  launch_transpose(w, wt, f, d, f, stream);
  VLP_CUDA_OK(cudaGetLastError());
  GemmParams p = {};
  if (d > 512) {
    // row does not fit one accumulator: plain GEMM into emb_f32, then a row-normalise pass
    p.n_tile = 512;
    p.k_per_split = (f + 31) & ~31;
    p.n_splits = 1;
    p.c = emb_f32;
    p.ldc = d;
    int rc2 = launch_gemm_kmajor(feat, f, wt, f, n, d, f, p, stream);
    if (rc2) return rc2;
    normalize_fwd_kernel<<<(n + 7) / 8, 256, 0, stream>>>(emb_f32, (__nv_bfloat16*)emb_bf16,
                                                          (__half*)emb_f16, inv_norm, n, d);
    VLP_COUNT_LAUNCH(1);
    VLP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  p.n_tile = (d + 15) & ~15;
  p.k_per_split = (f + 31) & ~31;
  p.n_splits = 1;
  p.normalize = 1;
  p.emb_f32 = emb_f32;
  p.emb_bf16 = (__nv_bfloat16*)emb_bf16;
  p.emb_f16 = (__half*)emb_f16;
  p.inv_norm = inv_norm;
  return launch_gemm_kmajor(feat, f, wt, f, n, d, f, p, stream);
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
  const size_t kp = (size_t)((k + 3) & ~3);
  int n_tile, kps, splits;
  gemm_plan(m, n, k, &n_tile, &kps, &splits);
  const size_t slabs = splits > 1 ? pg_align((size_t)splits * m * n * 4) : 0;
  return pg_align((size_t)m * kp * 4) + pg_align((size_t)n * kp * 4) + slabs;
}

int vlpclip_gemm_tf32(const float* a, const float* b, float* c, int m, int n, int k, int trans_a,
                      int trans_b, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (m <= 0 || n <= 0 || k <= 0) return fail(-1, "gemm: empty problem");
  if (!a || !b || !c || !workspace) return fail(-1, "gemm: null pointer");
  int rc = check_device_sm100();
  if (rc) return rc;
  if (workspace_bytes < vlpclip_gemm_workspace_bytes(m, n, k))
    return fail(-1, "gemm: workspace too small");
  // operands that are not K-major in memory are transposed into the workspace with their K
  // extent padded to a multiple of 4 floats: any K works there (dW = feat^T du has K = batch size,
  // and the reference's sampler yields remainder batches of arbitrary size)
  const int kp = (k + 3) & ~3;
  const float* a_k = a;
  const float* b_k = b;
  int lda = k, ldb = k;
  float* ws_a = (float*)workspace;
  float* ws_b = (float*)((uint8_t*)workspace + pg_align((size_t)m * kp * 4));
  float* ws_slab = (float*)((uint8_t*)ws_b + pg_align((size_t)n * kp * 4));
  if (trans_a) {  // given [k][m] -> need [m][k]
    launch_transpose(a, ws_a, k, m, kp, stream);
    a_k = ws_a;
    lda = kp;
  }
  if (!trans_b) {  // given [k][n] -> need [n][k]
    launch_transpose(b, ws_b, k, n, kp, stream);
    b_k = ws_b;
    ldb = kp;
  }
  VLP_CUDA_OK(cudaGetLastError());
  if ((lda % 4) != 0 || (ldb % 4) != 0)
    return fail(-1, "gemm: K (%d) must be a multiple of 4 for an operand that is already K-major", k);
  GemmParams p = {};
  int kps, splits;
  gemm_plan(m, n, k, &p.n_tile, &kps, &splits);
  p.k_per_split = kps;
  p.n_splits = splits;
  p.ldc = n;
  if (splits > 1) {
    p.c = ws_slab;
    p.split_stride = (size_t)m * n;
  } else {
    p.c = c;
    p.split_stride = 0;
  }
  rc = launch_gemm_kmajor(a_k, lda, b_k, ldb, m, n, k, p, stream);
  if (rc) return rc;
  if (splits > 1) {
    const size_t mn = (size_t)m * n;
    if (mn % 4 != 0 || (reinterpret_cast<uintptr_t>(c) & 15) != 0)
      return fail(-1, "gemm: split-K needs m*n %% 4 == 0 and a 16-byte aligned C");
    int blocks = (int)((mn / 4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>((const float4*)ws_slab, splits, mn / 4, mn / 4,
                                                     (float4*)c);
    VLP_COUNT_LAUNCH(1);
    VLP_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
