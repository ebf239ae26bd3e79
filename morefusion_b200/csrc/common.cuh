// Shared device/host helpers for libmorefusion_sm100a.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/morefusion_b200.h"

#define MF_CUDA_TRY(expr)                      \
  do {                                         \
    cudaError_t _e = (expr);                   \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

#define MF_LAUNCH_CHECK()                      \
  do {                                         \
    cudaError_t _e = cudaPeekAtLastError();    \
    if (_e != cudaSuccess) return (int)_e;     \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: remember
// what has been set per device, not per process, so a process that drives several GPUs works.
// Races between host threads are benign (the attribute call is idempotent).
// Every kernel of a launch sequence asks for the SAME shared-memory carveout (all shared): an SM
// has one L1/shared split at a time, and a kernel that prefers a different split than the one
// the SM is configured for can only start once the SM has drained - which serialises kernels
// that are meant to overlap on different streams / graph branches.
#define MF_ENSURE_DYN_SMEM(func, bytes)                                                        \
  do {                                                                                         \
    static int _mf_smem_set[64];                                                               \
    int _mf_dev = 0;                                                                           \
    MF_CUDA_TRY(cudaGetDevice(&_mf_dev));                                                      \
    if (_mf_dev < 0 || _mf_dev >= 64 || _mf_smem_set[_mf_dev] < (int)(bytes)) {                \
      MF_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                       (int)(bytes)));                                         \
      MF_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout,   \
                                       cudaSharedmemCarveoutMaxShared));                       \
      if (_mf_dev >= 0 && _mf_dev < 64) _mf_smem_set[_mf_dev] = (int)(bytes);                  \
    }                                                                                          \
  } while (0)

// same carveout request for kernels without dynamic shared memory
#define MF_PREFER_MAX_SMEM(func)                                                               \
  do {                                                                                         \
    static bool _mf_car_set[64];                                                               \
    int _mf_dev = 0;                                                                           \
    MF_CUDA_TRY(cudaGetDevice(&_mf_dev));                                                      \
    if (_mf_dev < 0 || _mf_dev >= 64 || !_mf_car_set[_mf_dev]) {                               \
      MF_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout,   \
                                       cudaSharedmemCarveoutMaxShared));                       \
      if (_mf_dev >= 0 && _mf_dev < 64) _mf_car_set[_mf_dev] = true;                           \
    }                                                                                          \
  } while (0)

namespace mf {

static inline unsigned div_up(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// (int)round((p - origin) / pitch): the reference's voxel index
// (average_voxelization_3d.py:84-86).  Explicit _rn intrinsics: never
// contracted, bit-identical to the oracle's float32 NumPy arithmetic.
__device__ __forceinline__ int voxel_coord(float p, float origin, float pitch) {
  return static_cast<int>(roundf(__fdiv_rn(__fsub_rn(p, origin), pitch)));
}
__device__ __forceinline__ float voxel_coord_f(float p, float origin, float pitch) {
  return __fdiv_rn(__fsub_rn(p, origin), pitch);
}

__device__ __forceinline__ float sq3(float a, float b, float c) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace mf
