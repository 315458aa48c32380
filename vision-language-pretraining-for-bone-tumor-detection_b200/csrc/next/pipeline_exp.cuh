// pipeline_exp.cuh -- shared by the staged ("next") versions of lse_fwd.cu / grad_bwd.cu:
// blocked-cycle accounting of the -DVLP_PROFILE_WAITS dev builds.
//
// csrc/next/ holds the NEXT versions of the two loss kernels: same algorithms as csrc/*.cu plus
// experiment switches (timing mocks, ping-pong softmax, ring geometry, cta_group::2 forward).  They
// are compiled only by tools/pipeline_experiments.py into tools/variants/*.so; the shipped library
// is built from csrc/*.cu, which stay byte-for-byte what was verified on the GPU.  A variant that
// wins on the B200 (parity + time) is promoted by copying it over the shipped file.
#pragma once
#include "../common.cuh"

namespace vlp {

// -DVLP_PROFILE_WAITS: cycles a role spends blocked on a barrier are added to wait_cyc[idx];
// otherwise VLP_WAIT(idx, stmt) is just stmt.
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
// device buffer the profiled kernels write to: [74 SM pairs][16] int64 (backward) followed by
// [148 SMs][8] int64 (forward); set with vlpclip_dev_set_wait_profile, null = off
constexpr int WAIT_PROF_BWD_WORDS = 74 * 16;
constexpr int WAIT_PROF_FWD_WORDS = 148 * 8;
inline long long*& wait_prof_buffer() {
  static long long* p = nullptr;
  return p;
}

// -DVLP_WAIT_WATCHDOG (set for every experiment build): a barrier wait that lasts longer than 2^31
// cycles (~1 s; a whole kernel takes milliseconds) reports itself and traps, so that a protocol bug
// in a staged variant ends the process with an error instead of hanging the GPU.
#ifdef VLP_WAIT_WATCHDOG
VLP_DEVICE void wait_watchdog_fire(uint32_t bar, uint32_t parity, int cluster_scope) {
  printf("VLP WATCHDOG: block %d thread %d (cluster rank %u) stuck on mbarrier smem+0x%x parity %u%s\n",
         (int)blockIdx.x, (int)threadIdx.x, cluster_ctarank(), bar & 0xFFFFFFu, parity,
         cluster_scope ? " (cluster-scope wait)" : "");
  __trap();
}
VLP_DEVICE void mbar_wait_wd(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > (1ll << 31)) wait_watchdog_fire(bar, parity, 0);
}
VLP_DEVICE void mbar_wait_cluster_wd(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
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
    if (!ok && clock64() - t0 > (1ll << 31)) wait_watchdog_fire(bar, parity, 1);
  }
}
#define mbar_wait mbar_wait_wd
#define mbar_wait_cluster mbar_wait_cluster_wd
#endif

}  // namespace vlp
