// TEST INFRASTRUCTURE: host spellings of the CUDA constructs morefusion_b200/csrc/mapping.cu uses,
// so that g++ can run its kernels serially (tests/test_mapping_emu.py).  One thread at a time, warps
// of one lane; compiled with -ffp-contract=off (the .cu is built with -fmad=false).
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) alignas(n)

struct ulonglong2 { unsigned long long x, y; };

struct emu_dim3 { unsigned x, y, z; };
static emu_dim3 blockIdx, threadIdx, blockDim, gridDim;

using std::isnan;
using std::min;

template <class T> static inline T atomicCAS(T* p, T cmp, T val) { T o = *p; if (o == cmp) *p = val; return o; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
// warps of one lane: the lane is active, alone, and its own leader
static inline unsigned __ballot_sync(unsigned, bool pred) { return pred ? 1u : 0u; }
template <class T> static inline unsigned __match_any_sync(unsigned, T) { return 1u; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline float __uint_as_float(unsigned u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
