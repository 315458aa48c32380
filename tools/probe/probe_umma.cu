// probe_umma.cu -- hardware fact-finding for the fused contrastive head (test infrastructure,
// not product code). Checks, against a CPU reference on exact small-integer data:
//   * TMA SW128 loads feeding tcgen05.mma with K-major / MN-major A and B operands
//   * cta_group::1 (M=128) and cta_group::2 (M=256 and M=128) accumulator layouts in TMEM
//   * mixed fp16(A) x bf16(B) under kind::f16
//   * A operand sourced from TMEM (TS form) written with tcgen05.st
// and measures
//   * tcgen05.mma issue rate per shape (all SMs busy, power-limited clocks)
//   * L2 -> smem TMA bandwidth per SM with every SM streaming
//   * the MMA rate under background TMA traffic into the same SM's shared memory ("b" tests:
//     does the 128 B/cycle of smem co-limit the SS/TS forms, and by how much does cta_group::2 help)
// Output: one line per test on stdout (PASS/FAIL + numbers).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "../../vision-language-pretraining-for-bone-tumor-detection_b200/csrc/sm100_ptx.cuh"

using namespace vlp;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static void init_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) {
    printf("no cuTensorMapEncodeTiled\n");
    exit(2);
  }
  g_encode = (EncodeTiledFn)fn;
}

// 16-bit row-major [outer][inner] matrix, box {64, box_outer}, 128B swizzle
static CUtensorMap make_map_16b(void* gptr, uint64_t inner, uint64_t outer, uint32_t box_outer) {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, gptr, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed %d\n", (int)r);
    exit(2);
  }
  return m;
}

// ---------------------------------------------------------------------------
// correctness kernel: one tile, D = A * B^T, K = KT
// ---------------------------------------------------------------------------
struct TileCfg {
  int cta_group;   // 1 or 2
  int M;           // UMMA M (total across the pair)
  int N;           // UMMA N (total across the pair)
  int KT;          // total K (multiple of 64)
  int a_major;     // 0 = K, 1 = MN
  int b_major;
  int a_fp16;      // A operand is fp16 (B always bf16)
  int a_tmem;      // A from TMEM (TS); requires a_major == K
  int pair_tma;    // use the cta_group::2 TMA form + single leader barrier
  int a_manual;    // A tile written to smem by threads (manual 128B swizzle) instead of TMA
};

template <int CG>
__global__ void __launch_bounds__(128, 1)
tile_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
            TileCfg cfg, const uint16_t* __restrict__ a_gmem, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  const int Ma = cfg.M / CG;  // A rows held by this CTA
  const int Nb = cfg.N / CG;  // B rows held by this CTA
  const int KT = cfg.KT;

  const uint32_t a_bytes = Ma * KT * 2;
  const uint32_t b_bytes = Nb * KT * 2;
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + ((a_bytes + 1023) & ~1023u);
  const uint32_t bar_load = smem_u32(&bars[0]);
  const uint32_t bar_mma = smem_u32(&bars[1]);

  if (threadIdx.x == 0) {
    uint32_t load_arrivals = 1;
    mbar_init(bar_load, load_arrivals);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<CG>(smem_u32(&tmem_base_s), 512);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  // TMEM column plan: accumulator at col 0 (<=256 cols), A-in-TMEM at col 256
  const uint32_t tmem_acc = tmem;
  const uint32_t tmem_a = tmem + 256;

  // ---- loads ----
  if (threadIdx.x == 0) {
    const bool pair = (CG == 2) && cfg.pair_tma;
    uint32_t tx = ((cfg.a_tmem || cfg.a_manual) ? 0 : a_bytes) + b_bytes;
    if (pair) {
      if (rank == 0) mbar_expect_tx(bar_load, tx * 2);
    } else {
      mbar_expect_tx(bar_load, tx);
    }
    auto ld = [&](uint32_t dst, const CUtensorMap* mp, int c0, int c1) {
      if (pair)
        tma_load_2d_pair(dst, mp, bar_load, c0, c1);
      else
        tma_load_2d(dst, mp, bar_load, c0, c1);
    };
    if (!cfg.a_tmem && !cfg.a_manual) {
      if (cfg.a_major == MAJOR_K) {
        for (int kb = 0; kb < KT / 64; ++kb)
          ld(sA + kb * Ma * 128, &mapA, kb * 64, rank * Ma);
      } else {
        for (int g = 0; g < Ma / 64; ++g) ld(sA + g * KT * 128, &mapA, rank * Ma + g * 64, 0);
      }
    }
    if (cfg.b_major == MAJOR_K) {
      for (int kb = 0; kb < KT / 64; ++kb) ld(sB + kb * Nb * 128, &mapB, kb * 64, rank * Nb);
    } else {
      for (int g = 0; g < Nb / 64; ++g) ld(sB + g * KT * 128, &mapB, rank * Nb + g * 64, 0);
    }
  }

  if (threadIdx.x == 0 && !((CG == 2) && cfg.pair_tma)) {
    mbar_wait(bar_load, 0);  // own loads landed (non-pair form)
  }
  // ---- A written by threads with the 128B swizzle pattern (generic proxy) ----
  if (cfg.a_manual && !cfg.a_tmem) {
    // a_gmem is row-major [M][KT]; rows of this CTA: rank*Ma .. +Ma
    for (int idx = threadIdx.x; idx < Ma * KT; idx += blockDim.x) {
      int m = idx / KT, k = idx % KT;
      uint16_t v = a_gmem[(size_t)(rank * Ma + m) * KT + k];
      uint32_t off;
      if (cfg.a_major == MAJOR_K) {
        off = (k / 64) * Ma * 128 + m * 128 + ((((k % 64) / 8) ^ (m & 7)) * 16) + (k % 8) * 2;
      } else {
        off = (m / 64) * KT * 128 + k * 128 + ((((m % 64) / 8) ^ (k & 7)) * 16) + (m % 8) * 2;
      }
      *reinterpret_cast<uint16_t*>(smem + off) = v;
    }
    fence_proxy_async_smem();
  }
  // ---- A into TMEM (TS form): thread = lane = row, K packed 2 per column ----
  if (cfg.a_tmem) {
    const int row = rank * Ma + threadIdx.x;  // Ma == 128 here
    const uint32_t taddr = tmem_a + ((warp * 32u) << 16);
    for (int c0 = 0; c0 < KT / 2; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t lo = a_gmem[(size_t)row * KT + 2 * (c0 + j)];
        uint32_t hi = a_gmem[(size_t)row * KT + 2 * (c0 + j) + 1];
        v[j] = lo | (hi << 16);
      }
      tmem_st_x8(taddr + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();

  // ---- MMA ----
  if (warp == 0 && (CG == 1 || rank == 0)) {
    const bool pair = (CG == 2) && cfg.pair_tma;
    if (CG == 2 && !pair) {
      // both CTAs waited on their own barrier below before the cluster sync
    }
    mbar_wait(bar_load, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = make_idesc(cfg.a_fp16 ? UMMA_F16 : UMMA_BF16, cfg.a_fp16 ? UMMA_F16 : UMMA_BF16, cfg.a_major,
                                        cfg.b_major, cfg.M, cfg.N);
      for (int ks = 0; ks < KT / 16; ++ks) {
        uint64_t ad = 0, bd;
        if (!cfg.a_tmem) {
          if (cfg.a_major == MAJOR_K)
            ad = make_sdesc_sw128(sA + (ks / 4) * Ma * 128 + (ks % 4) * 32, 0, 1024);
          else
            ad = make_sdesc_sw128(sA + ks * 2048, KT * 128, 1024);
        }
        if (cfg.b_major == MAJOR_K)
          bd = make_sdesc_sw128(sB + (ks / 4) * Nb * 128 + (ks % 4) * 32, 0, 1024);
        else
          bd = make_sdesc_sw128(sB + ks * 2048, KT * 128, 1024);
        if (cfg.a_tmem)
          umma_ts<CG>(tmem_acc, tmem_a + ks * 8, bd, idesc, ks > 0);
        else
          umma_ss<CG>(tmem_acc, ad, bd, idesc, ks > 0);
      }
      if (CG == 2)
        umma_commit_mcast<CG>(bar_mma, 0x3);
      else
        umma_commit<CG>(bar_mma);
    }
    __syncwarp();
  }
  // non-pair TMA in cta_group::2: the non-leader must make sure its own loads landed before the
  // leader issues; we do that conservatively with an extra cluster barrier ahead of the MMA when
  // pair_tma == 0 (handled on the host by launching that variant with `pre-wait`, see below).

  mbar_wait_cluster(bar_mma, 0);
  tc_fence_after();

  // ---- dump TMEM: out[rank][lane][col] for col < ncols ----
  const int ncols = (CG == 2 && cfg.M == 128) ? cfg.N / 2 : cfg.N;
  const uint32_t lane_base = warp * 32u;
  float* o = out + ((size_t)rank * 128 + threadIdx.x) * 256;
  for (int c = 0; c < ncols; c += 8) {
    uint32_t v[8];
    tmem_ld_x8(tmem_acc + (lane_base << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) o[c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) tmem_dealloc<CG>(tmem, 512);
}

// helper kernel variant: in cta_group::2 without pair TMA each CTA must wait for its own loads
// before the cluster barrier that precedes the MMA. We fold that in by having thread 0 of the
// non-leader wait on its barrier before the barrier; implemented via a tiny wrapper flag.
// (kept simple: pair_tma=0 path is only used with CG==1 in the test list below.)

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

static bool run_tile_test(const char* name, TileCfg cfg) {
  const int M = cfg.M, N = cfg.N, KT = cfg.KT;
  std::vector<float> A((size_t)M * KT), B((size_t)N * KT);
  srand(1234 + M * 7 + N * 3 + cfg.a_major * 11 + cfg.b_major * 13);
  for (auto& v : A) v = (float)((rand() % 7) - 3);
  for (auto& v : B) v = (float)((rand() % 7) - 3);
  // device storage in the operand's major-ness
  std::vector<uint16_t> hA((size_t)M * KT), hB((size_t)N * KT), hArow((size_t)M * KT);
  auto enc_a = [&](float f) -> uint16_t {
    if (cfg.a_fp16) {
      __half h = __float2half(f);
      return *reinterpret_cast<uint16_t*>(&h);
    }
    __nv_bfloat16 h = __float2bfloat16(f);
    return *reinterpret_cast<uint16_t*>(&h);
  };
  auto enc_b = [&](float f) -> uint16_t {
    if (cfg.a_fp16) {
      __half h = __float2half(f);
      return *reinterpret_cast<uint16_t*>(&h);
    }
    __nv_bfloat16 h = __float2bfloat16(f);
    return *reinterpret_cast<uint16_t*>(&h);
  };
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < KT; ++k) {
      uint16_t e = enc_a(A[(size_t)m * KT + k]);
      hArow[(size_t)m * KT + k] = e;
      if (cfg.a_major == MAJOR_K)
        hA[(size_t)m * KT + k] = e;
      else
        hA[(size_t)k * M + m] = e;
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < KT; ++k) {
      uint16_t e = enc_b(B[(size_t)n * KT + k]);
      if (cfg.b_major == MAJOR_K)
        hB[(size_t)n * KT + k] = e;
      else
        hB[(size_t)k * N + n] = e;
    }
  uint16_t *dA, *dB, *dArow;
  float* dOut;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dArow, hArow.size() * 2));
  CK(cudaMalloc(&dOut, 2 * 128 * 256 * sizeof(float)));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dArow, hArow.data(), hArow.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dOut, 0xFF, 2 * 128 * 256 * sizeof(float)));

  const int Ma = M / cfg.cta_group, Nb = N / cfg.cta_group;
  CUtensorMap mapA, mapB;
  if (cfg.a_major == MAJOR_K)
    mapA = make_map_16b(dA, KT, M, Ma);
  else
    mapA = make_map_16b(dA, M, KT, KT);
  if (cfg.b_major == MAJOR_K)
    mapB = make_map_16b(dB, KT, N, Nb);
  else
    mapB = make_map_16b(dB, N, KT, KT);

  size_t smem = (size_t)Ma * KT * 2 + (size_t)Nb * KT * 2 + 4096;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(cfg.cta_group);
  lc.blockDim = dim3(128);
  lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cfg.cta_group;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  cudaError_t le;
  if (cfg.cta_group == 1) {
    CK(cudaFuncSetAttribute(tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    le = cudaLaunchKernelEx(&lc, tile_kernel<1>, mapA, mapB, cfg, (const uint16_t*)dArow, dOut);
  } else {
    CK(cudaFuncSetAttribute(tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    le = cudaLaunchKernelEx(&lc, tile_kernel<2>, mapA, mapB, cfg, (const uint16_t*)dArow, dOut);
  }
  if (le != cudaSuccess) {
    printf("TILE %-44s LAUNCH-FAIL %s\n", name, cudaGetErrorString(le));
    cudaGetLastError();
    return false;
  }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) {
    printf("TILE %-44s RUNTIME-FAIL %s\n", name, cudaGetErrorString(se));
    exit(3);  // context is likely dead
  }
  std::vector<float> out(2 * 128 * 256);
  CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));

  // reference
  std::vector<float> D((size_t)M * N);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0;
      for (int k = 0; k < KT; ++k) acc += A[(size_t)m * KT + k] * B[(size_t)n * KT + k];
      D[(size_t)m * N + n] = acc;
    }
  long bad = 0;
  double maxerr = 0;
  for (int r = 0; r < cfg.cta_group; ++r)
    for (int l = 0; l < 128; ++l) {
      const int ncols = (cfg.cta_group == 2 && M == 128) ? N / 2 : N;
      for (int c = 0; c < ncols; ++c) {
        int m, n;
        if (cfg.cta_group == 1) {
          m = l;
          n = c;
        } else if (M == 256) {
          m = r * 128 + l;
          n = c;
        } else {
          m = r * 64 + (l % 64);
          n = (l / 64) * (N / 2) + c;
        }
        float got = out[((size_t)r * 128 + l) * 256 + c];
        float want = D[(size_t)m * N + n];
        double e = fabs((double)got - (double)want);
        if (!(e <= 1e-3)) ++bad;
        if (e > maxerr || e != e) maxerr = e;
      }
    }
  printf("TILE %-44s %s bad=%ld maxerr=%g\n", name, bad == 0 ? "PASS" : "FAIL", bad, maxerr);
  if (bad) {
    std::string fn = std::string("gpurun_out/probe_dump_") + name + ".bin";
    FILE* f = fopen(fn.c_str(), "wb");
    if (f) {
      int hdr[4] = {M, N, KT, cfg.cta_group};
      fwrite(hdr, 4, 4, f);
      fwrite(out.data(), 4, out.size(), f);
      fwrite(D.data(), 4, D.size(), f);
      fclose(f);
    }
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dArow);
  cudaFree(dOut);
  return bad == 0;
}

// ---------------------------------------------------------------------------
// MMA issue-rate benchmark: every CTA (pair) issues `iters` x 8 k-steps on resident operands
// ---------------------------------------------------------------------------
struct RateCfg {
  int cta_group, M, N, a_tmem, a_major, b_major, iters;
};

template <int CG>
__global__ void __launch_bounds__(128, 1)
rate_kernel(RateCfg cfg, long long* __restrict__ cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[1];
  __shared__ uint32_t tmem_base_s;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  const int Ma = cfg.M / CG, Nb = cfg.N / CG;
  const int KT = 128;  // operands cover 8 k-steps; we loop over them
  // fill smem with small pseudo-random bf16 (0x3c00..0x3fff -> ~0.008..1.99)
  const uint32_t total = (Ma + Nb) * KT * 2 + 2048;
  for (uint32_t i = threadIdx.x; i < total / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 97u);
    uint32_t lo = 0x3c00u | (h & 0x3ffu), hi = 0x3c00u | ((h >> 10) & 0x3ffu);
    reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
  }
  fence_proxy_async_smem();
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + ((Ma * KT * 2 + 1023) & ~1023u);
  const uint32_t bar = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<CG>(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (cfg.a_tmem) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0x3c003c00u + threadIdx.x + j;
    for (int c = 0; c < 64; c += 8) tmem_st_x8(tmem + 256 + ((warp * 32u) << 16) + c, v);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0 && (CG == 1 || rank == 0)) {
    if (elect_one()) {
      const uint32_t idesc =
          make_idesc(UMMA_BF16, UMMA_BF16, cfg.a_major, cfg.b_major, cfg.M, cfg.N);
      const uint32_t acc_cols = (CG == 2 && cfg.M == 128) ? cfg.N / 2 : cfg.N;
      const long long t0 = clock64();
      for (int it = 0; it < cfg.iters; ++it) {
        // two accumulators alternate when they fit in columns [0,256)
        const uint32_t acc = tmem + ((acc_cols <= 128 && (it & 1)) ? acc_cols : 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          uint64_t ad, bd;
          if (cfg.a_major == MAJOR_K)
            ad = make_sdesc_sw128(sA + (ks / 4) * Ma * 128 + (ks % 4) * 32, 0, 1024);
          else
            ad = make_sdesc_sw128(sA + ks * 2048, KT * 128, 1024);
          if (cfg.b_major == MAJOR_K)
            bd = make_sdesc_sw128(sB + (ks / 4) * Nb * 128 + (ks % 4) * 32, 0, 1024);
          else
            bd = make_sdesc_sw128(sB + ks * 2048, KT * 128, 1024);
          if (cfg.a_tmem)
            umma_ts<CG>(acc, tmem + 256 + ks * 8, bd, idesc, 1);
          else
            umma_ss<CG>(acc, ad, bd, idesc, 1);
        }
      }
      if (CG == 2)
        umma_commit_mcast<CG>(bar, 0x3);
      else
        umma_commit<CG>(bar);
      mbar_wait_cluster(bar, 0);
      const long long t1 = clock64();
      cycles_out[blockIdx.x / CG] = t1 - t0;
    }
    __syncwarp();
  }
  mbar_wait_cluster(bar, 0);
  tc_fence_after();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) tmem_dealloc<CG>(tmem, 512);
}

static void run_rate(const char* name, RateCfg cfg, int nsm) {
  const int Ma = cfg.M / cfg.cta_group, Nb = cfg.N / cfg.cta_group;
  size_t smem = (size_t)(Ma + Nb) * 128 * 2 + 4096;
  int grid = (nsm / cfg.cta_group) * cfg.cta_group;
  long long* dcyc;
  CK(cudaMalloc(&dcyc, grid * sizeof(long long)));
  CK(cudaMemset(dcyc, 0, grid * sizeof(long long)));
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(128);
  lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cfg.cta_group;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    cudaError_t le;
    if (cfg.cta_group == 1) {
      CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      le = cudaLaunchKernelEx(&lc, rate_kernel<1>, cfg, dcyc);
    } else {
      CK(cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      le = cudaLaunchKernelEx(&lc, rate_kernel<2>, cfg, dcyc);
    }
    if (le != cudaSuccess) {
      printf("RATE %-40s LAUNCH-FAIL %s\n", name, cudaGetErrorString(le));
      cudaGetLastError();
      return;
    }
    CK(cudaEventRecord(e1));
    cudaError_t se = cudaEventSynchronize(e1);
    if (se != cudaSuccess) {
      printf("RATE %-40s RUNTIME-FAIL %s\n", name, cudaGetErrorString(se));
      exit(3);
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  int units = grid / cfg.cta_group;
  for (int i = 0; i < units; ++i) avg += (double)cyc[i];
  avg /= units;
  double n_instr = (double)cfg.iters * 8;
  double macs = (double)cfg.M * cfg.N * 16 * n_instr;  // per CTA group
  double flops_total = 2.0 * macs * units;
  printf("RATE %-40s cyc/instr=%.1f  MAC/cyc/SM=%.0f  kernel=%.3f ms  %.1f TFLOP/s (grid %d)\n", name,
         avg / n_instr, macs / avg / cfg.cta_group, best, flops_total / (best * 1e-3) / 1e12, grid);
  cudaFree(dcyc);
}

// ---------------------------------------------------------------------------
// MMA rate UNDER background shared-memory traffic: the same issue loop as rate_kernel while warp 2
// streams TMA boxes (16 KB each, L2-resident source) into a scratch ring of the same SM, throttled
// to one box per `bg_interval` cycles (0 = as fast as completions allow).  Tests the round-1
// hypothesis that the backward's consumer SM (SS N=256: 70 B/cycle of operand reads + 47 B/cycle of
// TMA writes + 12 B/cycle of peer pushes) is limited by the 128 B/cycle of shared memory, and how
// much a cta_group::2 pair (each SM stages/reads half of B) relieves it.
// ---------------------------------------------------------------------------
template <int CG>
__global__ void __launch_bounds__(128, 1)
rate_bg_kernel(RateCfg cfg, const __grid_constant__ CUtensorMap map, int rows_total,
               int bg_interval, long long* __restrict__ cycles_out,
               long long* __restrict__ bg_boxes_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  constexpr int BG_STAGES = 4;
  __shared__ __align__(8) uint64_t bars[1 + BG_STAGES];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done_flag;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  const int Ma = cfg.M / CG, Nb = cfg.N / CG;
  const int KT = 128;
  const uint32_t op_bytes = ((Ma * KT * 2 + 1023) & ~1023u) + ((Nb * KT * 2 + 1023) & ~1023u);
  for (uint32_t i = threadIdx.x; i < op_bytes / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 97u);
    uint32_t lo = 0x3c00u | (h & 0x3ffu), hi = 0x3c00u | ((h >> 10) & 0x3ffu);
    reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
  }
  fence_proxy_async_smem();
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + ((Ma * KT * 2 + 1023) & ~1023u);
  const uint32_t ring = sA + op_bytes;
  const uint32_t bar = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    for (int i = 0; i < BG_STAGES; ++i) mbar_init(smem_u32(&bars[1 + i]), 1);
    done_flag = 0;
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<CG>(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (cfg.a_tmem) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0x3c003c00u + threadIdx.x + j;
    for (int c = 0; c < 64; c += 8) tmem_st_x8(tmem + 256 + ((warp * 32u) << 16) + c, v);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 2) {
    // background stream (every CTA of the group runs its own)
    if (elect_one()) {
      const int nblk_rows = rows_total / 128;
      long long boxes = 0;
      long long next = clock64();
      int issued = 0, waited = 0;
      auto issue = [&](int i) {
        const int st = i % BG_STAGES;
        const uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 7919u;
        mbar_expect_tx(smem_u32(&bars[1 + st]), 16384);
        tma_load_2d(ring + st * 16384, &map, smem_u32(&bars[1 + st]), ((h >> 4) % 8) * 64,
                    ((h >> 8) % nblk_rows) * 128);
      };
      for (; issued < BG_STAGES; ++issued) issue(issued);
      while (!done_flag) {
        const int st = waited % BG_STAGES;
        mbar_wait(smem_u32(&bars[1 + st]), (waited / BG_STAGES) & 1);
        ++waited;
        ++boxes;
        if (bg_interval > 0) {
          next += bg_interval;
          while (clock64() < next && !done_flag) {
          }
        }
        issue(issued);
        ++issued;
      }
      for (; waited < issued; ++waited)   // drain: no TMA may be in flight at exit
        mbar_wait(smem_u32(&bars[1 + waited % BG_STAGES]), (waited / BG_STAGES) & 1);
      bg_boxes_out[blockIdx.x] = boxes;
    }
    __syncwarp();
  }
  if (warp == 0 && (CG == 1 || rank == 0)) {
    if (elect_one()) {
      const uint32_t idesc =
          make_idesc(UMMA_BF16, UMMA_BF16, cfg.a_major, cfg.b_major, cfg.M, cfg.N);
      const uint32_t acc_cols = (CG == 2 && cfg.M == 128) ? cfg.N / 2 : cfg.N;
      const long long t0 = clock64();
      for (int it = 0; it < cfg.iters; ++it) {
        const uint32_t acc = tmem + ((acc_cols <= 128 && (it & 1)) ? acc_cols : 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          uint64_t ad, bd;
          if (cfg.a_major == MAJOR_K)
            ad = make_sdesc_sw128(sA + (ks / 4) * Ma * 128 + (ks % 4) * 32, 0, 1024);
          else
            ad = make_sdesc_sw128(sA + ks * 2048, KT * 128, 1024);
          if (cfg.b_major == MAJOR_K)
            bd = make_sdesc_sw128(sB + (ks / 4) * Nb * 128 + (ks % 4) * 32, 0, 1024);
          else
            bd = make_sdesc_sw128(sB + ks * 2048, KT * 128, 1024);
          if (cfg.a_tmem)
            umma_ts<CG>(acc, tmem + 256 + ks * 8, bd, idesc, 1);
          else
            umma_ss<CG>(acc, ad, bd, idesc, 1);
        }
      }
      if (CG == 2)
        umma_commit_mcast<CG>(bar, 0x3);
      else
        umma_commit<CG>(bar);
      mbar_wait_cluster(bar, 0);
      const long long t1 = clock64();
      cycles_out[blockIdx.x / CG] = t1 - t0;
    }
    __syncwarp();
  }
  mbar_wait_cluster(bar, 0);   // every CTA of the group sees the commit
  if (threadIdx.x == 0) done_flag = 1;
  tc_fence_after();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) tmem_dealloc<CG>(tmem, 512);
}

static void run_rate_bg(const char* name, RateCfg cfg, int bg_interval, int nsm) {
  const int Ma = cfg.M / cfg.cta_group, Nb = cfg.N / cfg.cta_group;
  const size_t smem = (size_t)(Ma + Nb) * 128 * 2 + 4 * 16384 + 6144;
  const int grid = (nsm / cfg.cta_group) * cfg.cta_group;
  const int rows_total = 65536;   // 64 MB source: L2 resident
  uint16_t* d;
  CK(cudaMalloc(&d, (size_t)rows_total * 512 * 2));
  CK(cudaMemset(d, 0x11, (size_t)rows_total * 512 * 2));
  CUtensorMap map = make_map_16b(d, 512, rows_total, 128);
  long long *dcyc, *dbox;
  CK(cudaMalloc(&dcyc, grid * sizeof(long long)));
  CK(cudaMalloc(&dbox, grid * sizeof(long long)));
  CK(cudaMemset(dcyc, 0, grid * sizeof(long long)));
  CK(cudaMemset(dbox, 0, grid * sizeof(long long)));
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid);
  lc.blockDim = dim3(128);
  lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cfg.cta_group;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t le;
    if (cfg.cta_group == 1) {
      CK(cudaFuncSetAttribute(rate_bg_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      le = cudaLaunchKernelEx(&lc, rate_bg_kernel<1>, cfg, map, rows_total, bg_interval, dcyc, dbox);
    } else {
      CK(cudaFuncSetAttribute(rate_bg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      le = cudaLaunchKernelEx(&lc, rate_bg_kernel<2>, cfg, map, rows_total, bg_interval, dcyc, dbox);
    }
    if (le != cudaSuccess) {
      printf("RATEBG %-36s LAUNCH-FAIL %s\n", name, cudaGetErrorString(le));
      cudaGetLastError();
      return;
    }
    cudaError_t se = cudaDeviceSynchronize();
    if (se != cudaSuccess) {
      printf("RATEBG %-36s RUNTIME-FAIL %s\n", name, cudaGetErrorString(se));
      exit(3);
    }
  }
  std::vector<long long> cyc(grid), box(grid);
  CK(cudaMemcpy(cyc.data(), dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(box.data(), dbox, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0, boxes = 0;
  const int units = grid / cfg.cta_group;
  for (int i = 0; i < units; ++i) avg += (double)cyc[i];
  for (int i = 0; i < grid; ++i) boxes += (double)box[i];
  avg /= units;
  boxes /= grid;
  const double n_instr = (double)cfg.iters * 8;
  const double macs = (double)cfg.M * cfg.N * 16 * n_instr;
  printf("RATEBG %-36s interval=%4d  cyc/instr=%.1f  MAC/cyc/SM=%.0f  background=%.1f B/cyc/SM\n",
         name, bg_interval, avg / n_instr, macs / avg / cfg.cta_group, boxes * 16384.0 / avg);
  cudaFree(d);
  cudaFree(dcyc);
  cudaFree(dbox);
}

// ---------------------------------------------------------------------------
// L2 -> smem TMA streaming bandwidth
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
l2bw_kernel(const __grid_constant__ CUtensorMap map, int rows_total, int iters, int same_tiles,
            long long* __restrict__ cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  constexpr int STAGES = 8;
  __shared__ __align__(8) uint64_t bars[STAGES];
  const uint32_t s0 = smem_u32(smem);
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nblk_rows = rows_total / 128;
    long long t0 = clock64();
    // each "load" = one 128-row x 64-col bf16 box = 16 KB
    int issued = 0, waited = 0;
    uint32_t seed = same_tiles ? 0u : blockIdx.x * 7919u;
    auto issue = [&](int i) {
      int st = i % STAGES;
      uint32_t h = (uint32_t)i * 2654435761u + seed;
      int rb = (h >> 8) % nblk_rows;
      int kb = (h >> 4) % 8;
      mbar_expect_tx(smem_u32(&bars[st]), 16384);
      tma_load_2d(s0 + st * 16384, &map, smem_u32(&bars[st]), kb * 64, rb * 128);
    };
    for (; issued < STAGES && issued < iters; ++issued) issue(issued);
    for (; waited < iters; ++waited) {
      int st = waited % STAGES;
      mbar_wait(smem_u32(&bars[st]), (waited / STAGES) & 1);
      if (issued < iters) {
        issue(issued);
        ++issued;
      }
    }
    long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
}

static void run_l2bw(const char* name, int rows_total, int same_tiles, int nsm) {
  uint16_t* d;
  size_t bytes = (size_t)rows_total * 512 * 2;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(d, 0x11, bytes));
  CUtensorMap map = make_map_16b(d, 512, rows_total, 128);
  long long* dcyc;
  CK(cudaMalloc(&dcyc, nsm * sizeof(long long)));
  size_t smem = 8 * 16384 + 2048;
  CK(cudaFuncSetAttribute(l2bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 4096;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    l2bw_kernel<<<nsm, 128, smem>>>(map, rows_total, iters, same_tiles, dcyc);
    CK(cudaEventRecord(e1));
    cudaError_t se = cudaEventSynchronize(e1);
    if (se != cudaSuccess) {
      printf("L2BW %-30s RUNTIME-FAIL %s\n", name, cudaGetErrorString(se));
      exit(3);
    }
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<long long> cyc(nsm);
  CK(cudaMemcpy(cyc.data(), dcyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (auto c : cyc) avg += (double)c;
  avg /= nsm;
  double bytes_per_sm = (double)iters * 16384;
  printf("L2BW %-30s B/cyc/SM=%.1f  aggregate=%.2f TB/s  (%.3f ms, footprint %.0f MB)\n", name,
         bytes_per_sm / avg, bytes_per_sm * nsm / (best * 1e-3) / 1e12, best, bytes / 1e6);
  cudaFree(d);
  cudaFree(dcyc);
}


// ---------------------------------------------------------------------------
// DSMEM hand-off: CTA0 writes a 128x128 fp16 K-major/SW128 tile into CTA1's smem with
// st.shared::cluster, signals a remote mbarrier; CTA1 feeds it to tcgen05.mma as the A operand
// and frees the slot with a multicast commit. Checks the accumulated result and the rate.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(288, 1)
dsmem_kernel(const __grid_constant__ CUtensorMap mapB, const uint16_t* __restrict__ g_rowmajor,
             int iters, int mode, float* __restrict__ out, long long* __restrict__ cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  constexpr int SLOTS = 2;
  __shared__ __align__(8) uint64_t bar_full[SLOTS], bar_empty[SLOTS], bar_b, bar_done;
  __shared__ uint32_t tmem_base_s;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const uint32_t sG = smem_u32(smem);              // SLOTS x 32 KB
  const uint32_t sB = sG + SLOTS * 32768;          // 128 x 128 bf16/fp16 K-major: 32 KB
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) {
      mbar_init(smem_u32(&bar_full[i]), mode == 0 ? 8 : 1);   // 8 producer warps / 1 bulk copy
      mbar_init(smem_u32(&bar_empty[i]), 1);  // one commit
    }
    mbar_init(smem_u32(&bar_b), 1);
    mbar_init(smem_u32(&bar_done), 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<1>(smem_u32(&tmem_base_s), 128);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (rank == 0) {
    // ---------------- producer ----------------
    if (warp < 8) {
      const int row = threadIdx.x & 127, half = threadIdx.x >> 7;  // 64 k-elements each
      uint4 v[8];
      const uint4* src = reinterpret_cast<const uint4*>(g_rowmajor + (size_t)row * 128 + half * 64);
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = src[c];
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const int slot = it % SLOTS;
        if (it >= SLOTS) mbar_wait_cluster(smem_u32(&bar_empty[slot]), ((it / SLOTS) - 1) & 1);
        const uint32_t base_local = sG + slot * 32768 + half * 16384 + row * 128;
        if (mode == 0) {
          const uint32_t base_remote = mapa_shared(base_local, 1);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t a = base_remote + ((c ^ (row & 7)) << 4);
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v[c].x),
                         "r"(v[c].y), "r"(v[c].z), "r"(v[c].w)
                         : "memory");
          }
          asm volatile("fence.proxy.async.shared::cluster;" ::: "memory");
          __syncwarp();
          if (lane_id() == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&bar_full[slot]), 1));
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t a = base_local + ((c ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v[c].x),
                         "r"(v[c].y), "r"(v[c].z), "r"(v[c].w)
                         : "memory");
          }
          fence_proxy_async_smem();
          bar_sync(1, 256);
          if (threadIdx.x == 0) {
            const uint32_t rbar = mapa_shared(smem_u32(&bar_full[slot]), 1);
            const uint32_t rdst = mapa_shared(sG + slot * 32768, 1);
            asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;"
                         ::"r"(rbar), "r"(32768) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(rdst), "r"(sG + slot * 32768), "r"(32768), "r"(rbar) : "memory");
          }
        }
      }
      // wait until the consumer drained everything
      for (int s2 = 0; s2 < SLOTS; ++s2) {
        int last = ((iters - 1 - s2) / SLOTS);  // use-count index of the last use of that slot
        int slot = (iters - 1 - s2) % SLOTS;
        if (iters - 1 - s2 >= 0) mbar_wait_cluster(smem_u32(&bar_empty[slot]), last & 1);
      }
      if (threadIdx.x == 0) cycles_out[0] = clock64() - t0;
    }
  } else {
    // ---------------- consumer ----------------
    if (warp == 8) {
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bar_b), 32768);
        tma_load_2d(sB, &mapB, smem_u32(&bar_b), 0, 0);
        tma_load_2d(sB + 16384, &mapB, smem_u32(&bar_b), 64, 0);
        mbar_wait(smem_u32(&bar_b), 0);
        const uint32_t idesc = make_idesc(UMMA_F16, UMMA_F16, MAJOR_K, MAJOR_K, 128, 128);
        for (int it = 0; it < iters; ++it) {
          const int slot = it % SLOTS;
          mbar_wait_cluster(smem_u32(&bar_full[slot]), (it / SLOTS) & 1);
          tc_fence_after();
          for (int ks = 0; ks < 8; ++ks) {
            uint64_t ad = make_sdesc_sw128(sG + slot * 32768 + (ks / 4) * 16384 + (ks % 4) * 32, 0, 1024);
            uint64_t bd = make_sdesc_sw128(sB + (ks / 4) * 16384 + (ks % 4) * 32, 0, 1024);
            umma_ss<1>(tmem, ad, bd, idesc, (it | ks) != 0);
          }
          // free the slot in the producer CTA (rank 0)
          umma_commit_mcast<1>(smem_u32(&bar_empty[slot]), 0x1);
        }
        umma_commit<1>(smem_u32(&bar_done));
      }
      __syncwarp();
    }
    if (warp < 4) {
      mbar_wait(smem_u32(&bar_done), 0);
      tc_fence_after();
      for (int c = 0; c < 128; c += 8) {
        uint32_t v[8];
        tmem_ld_x8(tmem + ((warp * 32u) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out[threadIdx.x * 128 + c + j] = __uint_as_float(v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 8) tmem_dealloc<1>(tmem, 128);
}

static void run_dsmem(int iters, int mode) {
  const int M = 128, N = 128, K = 128;
  std::vector<float> G((size_t)M * K), B((size_t)N * K);
  srand(77);
  for (auto& v : G) v = (float)((rand() % 5) - 2);
  for (auto& v : B) v = (float)((rand() % 5) - 2);
  std::vector<uint16_t> hG(G.size()), hB(B.size());
  for (size_t i = 0; i < G.size(); ++i) { __half h = __float2half(G[i]); hG[i] = *reinterpret_cast<uint16_t*>(&h); }
  for (size_t i = 0; i < B.size(); ++i) { __half h = __float2half(B[i]); hB[i] = *reinterpret_cast<uint16_t*>(&h); }
  uint16_t *dG, *dB; float* dOut; long long* dCyc;
  CK(cudaMalloc(&dG, hG.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dOut, M * N * 4)); CK(cudaMalloc(&dCyc, 8));
  CK(cudaMemcpy(dG, hG.data(), hG.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap mapB = make_map_16b(dB, K, N, 128);
  size_t smem = 3 * 32768 + 2048;
  CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(2); lc.blockDim = dim3(288); lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&lc, dsmem_kernel, mapB, (const uint16_t*)dG, iters, mode, dOut, dCyc);
  if (le != cudaSuccess) { printf("DSMEM LAUNCH-FAIL %s\n", cudaGetErrorString(le)); return; }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("DSMEM RUNTIME-FAIL %s\n", cudaGetErrorString(se)); exit(3); }
  std::vector<float> out(M * N); long long cyc;
  CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dCyc, 8, cudaMemcpyDeviceToHost));
  long bad = 0; double maxerr = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    float acc = 0; for (int k = 0; k < K; ++k) acc += G[m * K + k] * B[n * K + k];
    double e = fabs((double)out[m * N + n] - (double)acc * iters);
    if (e > 1e-2 * iters) ++bad; if (e > maxerr) maxerr = e;
  }
  printf("DSMEM handoff mode=%d iters=%d %s bad=%ld maxerr=%g cyc/tile=%.0f  => %.1f B/cyc pushed\n", mode, iters,
         bad == 0 ? "PASS" : "FAIL", bad, maxerr, (double)cyc / iters, 32768.0 * iters / (double)cyc);
  cudaFree(dG); cudaFree(dB); cudaFree(dOut); cudaFree(dCyc);
}

int main(int argc, char** argv) {
  init_encode();
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d smem/block optin=%zu\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  const int nsm = prop.multiProcessorCount;
  bool do_tiles = true, do_rate = true, do_l2 = true, do_ds = true, do_bg = true;
  if (argc > 1) {
    do_bg = strstr(argv[1], "b") != nullptr;
    do_tiles = strstr(argv[1], "t") != nullptr;
    do_rate = strstr(argv[1], "r") != nullptr;
    do_l2 = strstr(argv[1], "l") != nullptr;
    do_ds = strstr(argv[1], "d") != nullptr;
  }
  if (do_tiles) {
    //                                   cg   M    N   KT  aM bM f16 aT pair manual
    run_tile_test("cg1_M128_N128_KK",   {1, 128, 128, 128, 0, 0, 0, 0, 0, 0});
    run_tile_test("cg1_M128_N256_KK",   {1, 128, 256, 128, 0, 0, 0, 0, 0, 0});
    run_tile_test("cg1_M128_N128_K_MN", {1, 128, 128, 128, 0, 1, 0, 0, 0, 0});
    run_tile_test("cg1_M128_N256_K_MN", {1, 128, 256, 128, 0, 1, 0, 0, 0, 0});
    run_tile_test("cg1_M128_N128_MN_MN", {1, 128, 128, 128, 1, 1, 0, 0, 0, 0});
    run_tile_test("cg1_M128_N128_MN_K", {1, 128, 128, 128, 1, 0, 0, 0, 0, 0});
    run_tile_test("cg1_M128_N128_KK_f16A", {1, 128, 128, 128, 0, 0, 1, 0, 0, 0});
    run_tile_test("cg1_M128_N128_K_MN_f16A", {1, 128, 128, 128, 0, 1, 1, 0, 0, 0});
    run_tile_test("cg1_M128_N128_TS_K", {1, 128, 128, 128, 0, 0, 0, 1, 0, 0});
    run_tile_test("cg1_M128_N256_TS_MN_f16A", {1, 128, 256, 128, 0, 1, 1, 1, 0, 0});
    run_tile_test("cg1_M128_N128_manualA_K", {1, 128, 128, 128, 0, 0, 0, 0, 0, 1});
    run_tile_test("cg1_M128_N128_manualA_MN", {1, 128, 128, 128, 1, 0, 0, 0, 0, 1});
    run_tile_test("cg1_M128_N256_manualA_MN_f16_BMN", {1, 128, 256, 128, 1, 1, 1, 0, 0, 1});
    run_tile_test("cg2_M256_N128_KK_nopair", {2, 256, 128, 128, 0, 0, 0, 0, 0, 0});
    run_tile_test("cg2_M256_N128_KK",   {2, 256, 128, 128, 0, 0, 0, 0, 1, 0});
    run_tile_test("cg2_M256_N256_KK",   {2, 256, 256, 128, 0, 0, 0, 0, 1, 0});
    run_tile_test("cg2_M256_N256_K_MN", {2, 256, 256, 128, 0, 1, 0, 0, 1, 0});
    run_tile_test("cg2_M256_N256_MN_MN", {2, 256, 256, 128, 1, 1, 0, 0, 1, 0});
    run_tile_test("cg2_M128_N128_KK",   {2, 128, 128, 128, 0, 0, 0, 0, 1, 0});
    run_tile_test("cg2_M128_N256_KK",   {2, 128, 256, 128, 0, 0, 0, 0, 1, 0});
    run_tile_test("cg2_M128_N256_K_MN", {2, 128, 256, 128, 0, 1, 0, 0, 1, 0});
    run_tile_test("cg2_M256_N256_TS_MN", {2, 256, 256, 128, 0, 1, 0, 1, 1, 0});
  }
  if (do_rate) {
    const int it = 2000;
    //                                      cg   M    N  aT aM bM iters
    run_rate("cg1_M128_N128_SS_KK",        {1, 128, 128, 0, 0, 0, it}, nsm);
    run_rate("cg1_M128_N256_SS_KK",        {1, 128, 256, 0, 0, 0, it}, nsm);
    run_rate("cg1_M128_N256_SS_K_MN",      {1, 128, 256, 0, 0, 1, it}, nsm);
    run_rate("cg1_M128_N128_SS_MN_MN",     {1, 128, 128, 0, 1, 1, it}, nsm);
    run_rate("cg1_M128_N128_TS_K",         {1, 128, 128, 1, 0, 0, it}, nsm);
    run_rate("cg1_M128_N256_TS_MN",        {1, 128, 256, 1, 0, 1, it}, nsm);
    run_rate("cg2_M256_N128_SS_KK",        {2, 256, 128, 0, 0, 0, it}, nsm);
    run_rate("cg2_M256_N256_SS_KK",        {2, 256, 256, 0, 0, 0, it}, nsm);
    run_rate("cg2_M256_N256_SS_K_MN",      {2, 256, 256, 0, 0, 1, it}, nsm);
    run_rate("cg2_M128_N256_SS_KK",        {2, 128, 256, 0, 0, 0, it}, nsm);
    run_rate("cg2_M128_N256_SS_K_MN",      {2, 128, 256, 0, 0, 1, it}, nsm);
    run_rate("cg2_M256_N256_TS_MN",        {2, 256, 256, 1, 0, 1, it}, nsm);
  }
  if (do_bg) {
    const int it = 1000;
    // interval 350 = one 16 KB box per 350 cycles = 47 B/cycle (the consumer's Y stream), 700 = half
    for (int iv : {0, 350, 700, 100000000}) {
      run_rate_bg("cg1_M128_N256_SS_K_MN", {1, 128, 256, 0, 0, 1, it}, iv, nsm);
      run_rate_bg("cg2_M256_N256_SS_K_MN", {2, 256, 256, 0, 0, 1, it}, iv, nsm);
      run_rate_bg("cg1_M128_N128_TS_K", {1, 128, 128, 1, 0, 0, it}, iv, nsm);
      run_rate_bg("cg2_M256_N128_TS_K", {2, 256, 128, 1, 0, 0, it}, iv, nsm);
    }
  }
  if (do_l2) {
    run_l2bw("rows8192_distinct", 8192, 0, nsm);     // 8 MB footprint
    run_l2bw("rows65536_distinct", 65536, 0, nsm);   // 64 MB footprint
    run_l2bw("rows65536_same", 65536, 1, nsm);       // every SM walks the same tile sequence
    run_l2bw("rows524288_distinct", 524288, 0, nsm); // 512 MB footprint: HBM
  }
  if (do_ds) {
    run_dsmem(7, 0);
    run_dsmem(2000, 0);
    run_dsmem(1, 1);
    run_dsmem(7, 1);
    run_dsmem(2000, 1);
  }
  printf("PROBE DONE\n");
  return 0;
}
