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

}  // namespace vlp
