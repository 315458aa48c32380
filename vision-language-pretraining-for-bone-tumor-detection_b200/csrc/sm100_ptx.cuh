// sm100_ptx.cuh -- thin inline-PTX wrappers for the Blackwell (sm_100a) primitives
// the fused contrastive head is built from: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / st / fences), clusters.
//
// Everything here is a one-instruction wrapper; no policy. Bit layouts of the
// shared-memory ("matrix") descriptor and the instruction descriptor follow the
// sm_100 UMMA definition (start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 |
// layout <<61; idesc: c_fmt[4,6) a_fmt[7,10) b_fmt[10,13) a_major 15 b_major 16
// N>>3 [17,23) M>>4 [24,29)).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace vlp {

#define VLP_DEVICE __device__ __forceinline__

// ----------------------------------------------------------------------------
// address helpers
// ----------------------------------------------------------------------------
VLP_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

VLP_DEVICE uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

VLP_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

VLP_DEVICE uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

// one lane of a fully-active warp is elected; returns 1 in that lane
VLP_DEVICE uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xFFFFFFFF;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred));
  return pred;
}

// map a local shared::cta address to the same offset in CTA `rank` of the cluster
VLP_DEVICE uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// ----------------------------------------------------------------------------
// cluster barrier
// ----------------------------------------------------------------------------
VLP_DEVICE void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
VLP_DEVICE void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
VLP_DEVICE void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
VLP_DEVICE void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
VLP_DEVICE void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
VLP_DEVICE void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier that lives in another CTA of the cluster (addr from mapa_shared)
VLP_DEVICE void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar)
               : "memory");
}
VLP_DEVICE void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
VLP_DEVICE uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// non-blocking test (try_wait may suspend the thread): used to peek at the NEXT pipeline stage
// before issuing the current stage's MMAs, so that its ~100-cycle latency overlaps the issue
VLP_DEVICE uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
VLP_DEVICE void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// cluster-scope acquire variant (needed when the arrival came from a peer CTA / multicast commit)
VLP_DEVICE void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

// generic-proxy writes to smem -> visible to the async proxy (TMA / UMMA operand reads)
VLP_DEVICE void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// named barrier among a subset of warps
VLP_DEVICE void bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
VLP_DEVICE void bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------
VLP_DEVICE void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// 2D tiled load, completes `bytes` on the CTA-local mbarrier
VLP_DEVICE void tma_load_2d(uint32_t dst_smem, const void* tmap, uint32_t bar, int32_t c0,
                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// 2D tiled load, multicast to the CTAs in `mask` (same smem offset + same barrier offset in each)
VLP_DEVICE void tma_load_2d_mcast(uint32_t dst_smem, const void* tmap, uint32_t bar, int32_t c0,
                                  int32_t c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "h"(mask), "r"(c0),
        "r"(c1)
      : "memory");
}

// 2D tiled load issued inside a CTA pair: data lands in THIS CTA's smem, the
// transaction bytes are counted on the barrier of the pair's leader (even) CTA.
VLP_DEVICE void tma_load_2d_pair(uint32_t dst_smem, const void* tmap, uint32_t bar, int32_t c0,
                                 int32_t c1) {
  uint32_t leader_bar = bar & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}

// 2D tiled store smem -> global (bulk async group)
VLP_DEVICE void tma_store_2d(const void* tmap, uint32_t src_smem, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tmap)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
VLP_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
VLP_DEVICE void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
VLP_DEVICE void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// 1D bulk copies (no tensor map): byte images move unchanged, so a 128B-swizzled smem tile can
// travel smem -> global -> smem of another SM and still match its UMMA descriptor.
VLP_DEVICE void bulk_store_1d(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
VLP_DEVICE void bulk_load_1d(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          dst_smem),
      "l"(src_gmem), "r"(bytes), "r"(bar)
      : "memory");
}
// orders generic-proxy and async-proxy (TMA / bulk copy) accesses of this thread, all state spaces
VLP_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ----------------------------------------------------------------------------
// global-memory flags between persistent CTAs (gpu scope)
// ----------------------------------------------------------------------------
VLP_DEVICE int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
VLP_DEVICE void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
VLP_DEVICE void st_relaxed_gpu(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
VLP_DEVICE float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// ----------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ----------------------------------------------------------------------------
template <int CTA_GROUP>
VLP_DEVICE void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (CTA_GROUP == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CTA_GROUP>
VLP_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CTA_GROUP == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
  } else {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
  }
}

VLP_DEVICE void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
VLP_DEVICE void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ----------------------------------------------------------------------------
// tcgen05: descriptors
// ----------------------------------------------------------------------------
enum : uint32_t { UMMA_F16 = 0, UMMA_BF16 = 1, UMMA_TF32 = 2 };
enum : uint32_t { MAJOR_K = 0, MAJOR_MN = 1 };

// kind::f16 instruction descriptor, fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc(uint32_t a_fmt, uint32_t b_fmt, uint32_t a_major,
                                                  uint32_t b_major, uint32_t M, uint32_t N) {
  return (1u << 4)             // D format = F32
         | (a_fmt << 7)        // A format
         | (b_fmt << 10)       // B format
         | (a_major << 15)     // A major-ness
         | (b_major << 16)     // B major-ness
         | ((N >> 3) << 17)    // N
         | ((M >> 4) << 24);   // M
}

// 128B-swizzled shared-memory operand descriptor.
//  K-major : rows of 128 B (64 x 16-bit along K), 8-row atoms; SBO = byte stride between
//            8-row groups (1024 when rows are dense), LBO unused.
//  MN-major: rows of 128 B (64 x 16-bit along M/N), 8 k-rows per atom; SBO = byte stride
//            between 8-k groups, LBO = byte stride between 64-element groups along M/N.
VLP_DEVICE uint64_t make_sdesc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}

// The two 32-bit words of a descriptor.  The high word (SBO, version, swizzle) is loop invariant and an
// smem offset only moves the low word's 14-bit address field ((addr & 0x3FFFF) >> 4 never carries out
// of it for a valid address), so an issue loop needs ONE 32-bit add per MMA instead of rebuilding the
// 64-bit descriptor (the MMA warp's issue path is a dependent chain of uniform-datapath instructions:
// its length, not the tensor pipe, bounded the S-tile kernels -- profiles/r02_issue_path.txt).
VLP_DEVICE uint32_t sdesc_lo_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
VLP_DEVICE uint32_t sdesc_hi_sw128(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}
// layout type 1 = 128B swizzle with a 32-byte base (MN-major tf32 operands; atom = 4 k rows of 128 B,
// SBO = byte stride between 4-row groups)
VLP_DEVICE uint32_t sdesc_hi_sw128_base32(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (1u << 29);
}

// ----------------------------------------------------------------------------
// tcgen05: MMA + commit
// ----------------------------------------------------------------------------
template <int CTA_GROUP>
VLP_DEVICE void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                        uint32_t accumulate) {
  if constexpr (CTA_GROUP == 1) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// kind::tf32: fp32 operands in smem (K = 8 per instruction), fp32 accumulate
VLP_DEVICE void umma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// word-form variant (see umma_ss_w)
VLP_DEVICE void umma_ss_tf32_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 ad, bd;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 ad, {%1, %2};\n"
      "mov.b64 bd, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %5, p;\n"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A operand from tensor memory (K-major only), B from shared memory
template <int CTA_GROUP>
VLP_DEVICE void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                        uint32_t accumulate) {
  if constexpr (CTA_GROUP == 1) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// word-form variants (cta_group::1): descriptors passed as {lo, hi} 32-bit words, accumulate flag as
// a compile-time-foldable integer
VLP_DEVICE void umma_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 bd;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 bd, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
VLP_DEVICE void umma_ss_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 ad, bd;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "mov.b64 ad, {%1, %2};\n"
      "mov.b64 bd, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One ring stage of an S tile, TS form: NM MMAs (K = 16 each) over NM / 4 boxes of [128 rows x 64 k]
// (16 KB apart); A advances 8 TMEM columns per MMA.  Fully unrolled: one add per operand per MMA.
template <int NM>
VLP_DEVICE void umma_ts_stage(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                              uint32_t idesc, uint32_t first_accumulate) {
#pragma unroll
  for (int j = 0; j < NM; ++j)
    umma_ts_w(d_tmem, a_tmem + j * 8, b_lo + (j >> 2) * (16384 >> 4) + (j & 3) * 2, b_hi, idesc,
              j == 0 ? first_accumulate : 1u);
}

// all previously issued MMAs of this thread arrive (once) on `bar` when they complete

// One ring stage in ONE asm block: peek at the NEXT stage's "full" barrier (mbarrier.test_wait, ~100
// cycles of latency), issue this stage's MMAs, commit them to this stage's "empty" barrier, and only
// then read the peek's answer -- so the barrier latency hides under the issue.  Returns 1 when the
// next stage had landed.  Meant for a single elected thread that runs the whole issue loop.
template <int NM>
VLP_DEVICE uint32_t umma_ts_stage_peek(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                       uint32_t idesc, uint32_t first_accumulate, uint32_t commit_bar,
                                       uint32_t peek_bar, uint32_t peek_parity);
template <>
VLP_DEVICE uint32_t umma_ts_stage_peek<8>(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t first_accumulate, uint32_t commit_bar,
                                             uint32_t peek_bar, uint32_t peek_parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p, pt, q;\n"
      ".reg .b64 bd;\n"
      ".reg .b32 ta, lo;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%8], %9;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "setp.ne.b32 pt, %10, 0;\n"
      "mov.b64 bd, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], bd, %5, p;\n"
      "add.u32 ta, %2, 8;\n"
      "add.u32 lo, %3, 2;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "add.u32 ta, %2, 16;\n"
      "add.u32 lo, %3, 4;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "add.u32 ta, %2, 24;\n"
      "add.u32 lo, %3, 6;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "add.u32 ta, %2, 32;\n"
      "add.u32 lo, %3, 1024;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "add.u32 ta, %2, 40;\n"
      "add.u32 lo, %3, 1026;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "add.u32 ta, %2, 48;\n"
      "add.u32 lo, %3, 1028;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "add.u32 ta, %2, 56;\n"
      "add.u32 lo, %3, 1030;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n"
      "selp.u32 %0, 1, 0, q;\n"
      "}\n"
      : "=r"(ok)
      : "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(first_accumulate), "r"(commit_bar),
        "r"(peek_bar), "r"(peek_parity), "r"(1u)
      : "memory");
  return ok;
}
template <>
VLP_DEVICE uint32_t umma_ts_stage_peek<4>(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t first_accumulate, uint32_t commit_bar,
                                             uint32_t peek_bar, uint32_t peek_parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p, pt, q;\n"
      ".reg .b64 bd;\n"
      ".reg .b32 ta, lo;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%8], %9;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "setp.ne.b32 pt, %10, 0;\n"
      "mov.b64 bd, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], bd, %5, p;\n"
      "add.u32 ta, %2, 8;\n"
      "add.u32 lo, %3, 2;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "add.u32 ta, %2, 16;\n"
      "add.u32 lo, %3, 4;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "add.u32 ta, %2, 24;\n"
      "add.u32 lo, %3, 6;\n"
      "mov.b64 bd, {lo, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], [ta], bd, %5, pt;\n"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n"
      "selp.u32 %0, 1, 0, q;\n"
      "}\n"
      : "=r"(ok)
      : "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(first_accumulate), "r"(commit_bar),
        "r"(peek_bar), "r"(peek_parity), "r"(1u)
      : "memory");
  return ok;
}
// ... SS form, 4 MMAs of one consumer stage: A low word advances by a_step per MMA, B by 128 (2 KB)
VLP_DEVICE uint32_t umma_ss_stage4_peek(uint32_t d_tmem, uint32_t a_lo, uint32_t a_step, uint32_t b_lo,
                                        uint32_t d_hi, uint32_t idesc, uint32_t first_accumulate,
                                        uint32_t commit_bar, uint32_t peek_bar, uint32_t peek_parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p, pt, q;\n"
      ".reg .b64 ad, bd;\n"
      ".reg .b32 la, lb;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 q, [%9], %10;\n"
      "setp.ne.b32 p, %7, 0;\n"
      "setp.ne.b32 pt, %11, 0;\n"
      "mov.b64 ad, {%2, %5};\n"
      "mov.b64 bd, {%4, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], ad, bd, %6, p;\n"
      "mad.lo.u32 la, %3, 1, %2;\n"
      "add.u32 lb, %4, 128;\n"
      "mov.b64 ad, {la, %5};\n"
      "mov.b64 bd, {lb, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], ad, bd, %6, pt;\n"
      "mad.lo.u32 la, %3, 2, %2;\n"
      "add.u32 lb, %4, 256;\n"
      "mov.b64 ad, {la, %5};\n"
      "mov.b64 bd, {lb, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], ad, bd, %6, pt;\n"
      "mad.lo.u32 la, %3, 3, %2;\n"
      "add.u32 lb, %4, 384;\n"
      "mov.b64 ad, {la, %5};\n"
      "mov.b64 bd, {lb, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%1], ad, bd, %6, pt;\n"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n"
      "selp.u32 %0, 1, 0, q;\n"
      "}\n"
      : "=r"(ok)
      : "r"(d_tmem), "r"(a_lo), "r"(a_step), "r"(b_lo), "r"(d_hi), "r"(idesc), "r"(first_accumulate),
        "r"(commit_bar), "r"(peek_bar), "r"(peek_parity), "r"(1u)
      : "memory");
  return ok;
}

template <int CTA_GROUP>
VLP_DEVICE void umma_commit(uint32_t bar) {
  if constexpr (CTA_GROUP == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     bar)
                 : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     bar)
                 : "memory");
  }
}
// ... and on the same barrier offset in every CTA of `mask`
template <int CTA_GROUP>
VLP_DEVICE void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  if constexpr (CTA_GROUP == 1) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
        " [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
  } else {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
        " [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
  }
}

// ----------------------------------------------------------------------------
// tcgen05: TMEM <-> registers (32 lanes x 32-bit, N consecutive columns per thread)
// ----------------------------------------------------------------------------
VLP_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
VLP_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

VLP_DEVICE void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
}

VLP_DEVICE void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

VLP_DEVICE void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

VLP_DEVICE void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7])
      : "memory");
}

VLP_DEVICE void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15])
      : "memory");
}

// ----------------------------------------------------------------------------
// misc math
// ----------------------------------------------------------------------------
VLP_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
VLP_DEVICE float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace vlp
