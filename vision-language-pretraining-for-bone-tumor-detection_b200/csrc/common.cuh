// common.cuh -- host-side helpers shared by the kernels of the fused contrastive head:
// thread-local error string, TMA tensor-map construction, launch helpers.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"

namespace vlp {

// ---- error reporting across the C ABI -------------------------------------------------
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
#define VLP_CUDA_OK(expr)                                                             \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess)                                                           \
      return ::vlp::fail(-2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                         __FILE__, __LINE__);                                         \
  } while (0)

// ---- TMA descriptors --------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        p)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Row-major [outer][inner] matrix of `elem_bytes`-wide elements, row stride `ld` elements.
// Box = {128 B worth of inner elements, box_outer rows}, 128-byte swizzle, OOB reads give 0.
// atom32: the 128B swizzle moves 32-byte chunks (XOR with row mod 4) instead of 16-byte ones -- the only
// shared-memory layout the tensor cores accept for MN-major 32-bit (tf32) operands.
inline int make_tmap_sw128(CUtensorMap* out, const void* gptr, int elem_bytes, uint64_t inner,
                           uint64_t outer, uint64_t ld, uint32_t box_outer, bool as_float32 = false,
                           bool atom32 = false) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(-3, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  if ((reinterpret_cast<uintptr_t>(gptr) & 15) != 0)
    return fail(-4, "tensor base address must be 16-byte aligned");
  if ((ld * elem_bytes) % 16 != 0)
    return fail(-4, "row stride must be a multiple of 16 bytes (got %llu elements of %d B)",
                (unsigned long long)ld, elem_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), box_outer};
  cuuint32_t estr[2] = {1, 1};
  // (loads and stores only move bytes; the element type matters for TMA reduce-add)
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                           : as_float32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                           : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  CUresult r = enc(out, dt, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-5, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// per-device caches (a process may drive several GPUs)
constexpr int kMaxDevices = 64;
inline int sm_count() {
  static int n[kMaxDevices] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  if (!n[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    n[dev] = v;
  }
  return n[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per
// (kernel slot, device).  `slot` is a small per-kernel index chosen by the caller.
inline cudaError_t set_smem_attr_once(const void* fn, int bytes, int slot = 0) {
  static unsigned char done[16][kMaxDevices] = {{0}};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (slot < 0 || slot >= 16 || dev < 0 || dev >= kMaxDevices)
    return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (done[slot][dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[slot][dev] = 1;
  return e;
}

// optional cap on the SMs the persistent kernels occupy (0 = all): leaves room for NCCL kernels
// that overlap with the backward in the sharded variant
inline int& sm_limit() {
  static int n = 0;
  return n;
}
inline int usable_sms() {
  const int n = sm_count();
  const int lim = sm_limit();
  return (lim > 0 && lim < n) ? lim : n;
}

inline int check_device_sm100() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(-6, "no CUDA device");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return fail(-6, "device compute capability %d.%d is not sm_100 (B200); no fallback path",
                major, minor);
  return 0;
}

// number of kernels this library has launched (bench.py reports it as gpu_launches)
inline unsigned long long& launch_counter() {
  static unsigned long long n = 0;
  return n;
}
#define VLP_COUNT_LAUNCH(k) (::vlp::launch_counter() += (k))

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

}  // namespace vlp
