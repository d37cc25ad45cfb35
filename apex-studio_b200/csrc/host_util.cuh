// Host-side helpers shared by the C-ABI entry points: error codes, the driver entry point for
// cuTensorMapEncodeTiled (resolved at run time so the library links against cudart only and still
// loads on a machine without a driver), and bf16 tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/apex_b200.h"

namespace b200 {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// bf16 tensor map with 128-byte swizzle. dims/strides are innermost-first; strides in ELEMENTS for
// dims 1..rank-1 (dim 0 is contiguous). box is innermost-first. Returns B200_OK or an error code.
inline int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_elems, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return B200_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return B200_ERR_ALIGN;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_elems[i] * 2;  // bytes
      if (gstr[i - 1] % 16 != 0) return B200_ERR_ALIGN;
    }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[apex_b200] cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return B200_ERR_TMAP;
  }
  return B200_OK;
}

constexpr int kMaxDevices = 64;

// Index of the current device, or -1.  One-time per-DEVICE setup (the opt-in to large dynamic shared memory is a per-device
// function attribute; the SM count may differ) is keyed by it; std::atomic keeps concurrent first calls (ctypes releases the
// GIL) benign -- both threads would do the same idempotent work.
inline int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

inline int num_sms() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  if (dev < 0) return 148;
  int v = n[dev].load(std::memory_order_relaxed);
  if (v) return v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  n[dev].store(v, std::memory_order_relaxed);
  return v;
}

// Runs `setup()` (returns true on success) once per device; false if it failed or there is no current device.
template <typename F>
inline bool once_per_device(std::atomic<bool> (&done)[kMaxDevices], F&& setup) {
  const int dev = current_device();
  if (dev < 0) return false;
  if (done[dev].load(std::memory_order_acquire)) return true;
  if (!setup()) return false;
  done[dev].store(true, std::memory_order_release);
  return true;
}

#define B200_CHECK_LAUNCH()                                                     \
  do {                                                                          \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) {                                                    \
      fprintf(stderr, "[apex_b200] %s:%d launch failed: %s\n", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return B200_ERR_LAUNCH;                                                   \
    }                                                                           \
  } while (0)

}  // namespace b200
