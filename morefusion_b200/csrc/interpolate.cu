// interpolate_voxel_grid (trilinear gather / scatter) for sm_100a.
//
// Replaces the reference's CuPy kernels K7/K8:
//   morefusion/functions/geometry/interpolate_voxel_grid.py:6-59 (weights),
//   :170-212 (forward), :224-266 (backward).
// Semantics kept: (int) truncation of the coordinate, weight order
// w000,w100,w010,w001,w110,w011,w101,w111, corners outside the grid skipped
// without renormalisation, accumulation in corner order j=0..7.  Uses the
// correct (Y*Z, Z, 1) strides (the reference forward's (X*Y, Y, 1) is equal
// for cubic grids only).
#include "common.cuh"

namespace mf {

struct Tri {
  float w[8];
  int ix[8], iy[8], iz[8];
};

__device__ __forceinline__ void trilinear(float x, float y, float z, Tri& t) {
  int x0 = static_cast<int>(x), y0 = static_cast<int>(y), z0 = static_cast<int>(z);
  float lx = __fsub_rn(x, (float)x0), ly = __fsub_rn(y, (float)y0), lz = __fsub_rn(z, (float)z0);
  float hx = __fsub_rn(1.f, lx), hy = __fsub_rn(1.f, ly), hz = __fsub_rn(1.f, lz);
  t.w[0] = __fmul_rn(__fmul_rn(hx, hy), hz);
  t.w[1] = __fmul_rn(__fmul_rn(lx, hy), hz);
  t.w[2] = __fmul_rn(__fmul_rn(hx, ly), hz);
  t.w[3] = __fmul_rn(__fmul_rn(hx, hy), lz);
  t.w[4] = __fmul_rn(__fmul_rn(lx, ly), hz);
  t.w[5] = __fmul_rn(__fmul_rn(hx, ly), lz);
  t.w[6] = __fmul_rn(__fmul_rn(lx, hy), lz);
  t.w[7] = __fmul_rn(__fmul_rn(lx, ly), lz);
  const int dx[8] = {0, 1, 0, 0, 1, 0, 1, 1};
  const int dy[8] = {0, 0, 1, 0, 1, 1, 0, 1};
  const int dz[8] = {0, 0, 0, 1, 0, 1, 1, 1};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    t.ix[j] = x0 + dx[j];
    t.iy[j] = y0 + dy[j];
    t.iz[j] = z0 + dz[j];
  }
}

// one thread per (point, channel); channel fastest so that the [P,C] output is coalesced
// and, for channels_last grids, so are the 8 corner reads.
template <bool kChannelsLast>
__global__ void k_interp_fwd(const float* __restrict__ vox, const float* __restrict__ points,
                             const int* __restrict__ bi, long long P, int B, int C, int X, int Y,
                             int Z, float* __restrict__ values) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * C) return;
  long long n = i / C;
  int c = (int)(i - n * C);
  int b = bi[n];
  Tri t;
  trilinear(points[3 * n], points[3 * n + 1], points[3 * n + 2], t);
  const long long V = (long long)X * Y * Z;
  float acc = 0.f;
  if (b >= 0 && b < B) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (t.ix[j] >= 0 && t.ix[j] < X && t.iy[j] >= 0 && t.iy[j] < Y && t.iz[j] >= 0 &&
          t.iz[j] < Z) {
        long long flat = ((long long)t.ix[j] * Y + t.iy[j]) * Z + t.iz[j];
        float v = kChannelsLast ? __ldg(vox + ((long long)b * V + flat) * C + c)
                                : __ldg(vox + ((long long)b * C + c) * V + flat);
        acc = __fadd_rn(acc, __fmul_rn(t.w[j], v));
      }
    }
  }
  values[i] = acc;
}

template <bool kChannelsLast>
__global__ void k_interp_bwd(const float* __restrict__ gvalues, const float* __restrict__ points,
                             const int* __restrict__ bi, long long P, int B, int C, int X, int Y,
                             int Z, float* __restrict__ gvox) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * C) return;
  long long n = i / C;
  int c = (int)(i - n * C);
  int b = bi[n];
  if (b < 0 || b >= B) return;
  Tri t;
  trilinear(points[3 * n], points[3 * n + 1], points[3 * n + 2], t);
  const long long V = (long long)X * Y * Z;
  float g = gvalues[i];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (t.ix[j] >= 0 && t.ix[j] < X && t.iy[j] >= 0 && t.iy[j] < Y && t.iz[j] >= 0 &&
        t.iz[j] < Z) {
      long long flat = ((long long)t.ix[j] * Y + t.iy[j]) * Z + t.iz[j];
      float* dst = kChannelsLast ? gvox + ((long long)b * V + flat) * C + c
                                 : gvox + ((long long)b * C + c) * V + flat;
      atomicAdd(dst, __fmul_rn(t.w[j], g));
    }
  }
}

// ---- plane-staged forward for the reference layout [B,C,X,Y,Z] -----------------------------
// The channel-major grid makes the 8 corner reads of one (point, channel) pair land in 8
// different 128-byte lines per channel, so the simple kernel above moves ~8x the algorithmic
// bytes through L2.  Here a CTA owns (batch b, CG consecutive channels): it streams those CG
// grid planes into shared memory with coalesced 16-byte loads (every grid element is read from
// HBM exactly once), then each thread takes one point of batch b and gathers its 8 corners x CG
// channels from shared memory.  Same arithmetic and accumulation order as k_interp_fwd.
template <int CG>
__global__ void __launch_bounds__(256)
k_interp_fwd_planes(const float* __restrict__ vox, const float* __restrict__ points,
                    const int* __restrict__ bi, long long P, int B, int C, int X, int Y, int Z,
                    float* __restrict__ values) {
  extern __shared__ __align__(16) float planes[];          // [CG][V]
  const int V = X * Y * Z;
  const int b = blockIdx.y, c0 = blockIdx.x * CG;
  const int cg = min(CG, C - c0);
  const float* src = vox + ((long long)b * C + c0) * V;
  if ((V & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(planes);
    for (int e = threadIdx.x; e < cg * (V >> 2); e += 256) d4[e] = __ldg(s4 + e);
  } else {
    for (int e = threadIdx.x; e < cg * V; e += 256) planes[e] = __ldg(src + e);
  }
  __syncthreads();
  // points of batch b: the batch indices of 8 strided points are loaded together (independent
  // loads in flight) before any of them is processed
  for (long long base = 0; base < P; base += 256 * 8) {
    int bb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      long long n = base + u * 256 + threadIdx.x;
      bb[u] = (n < P) ? __ldg(bi + n) : -1;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (bb[u] != b) continue;
      const long long n = base + u * 256 + threadIdx.x;
      Tri t;
      trilinear(points[3 * n], points[3 * n + 1], points[3 * n + 2], t);
      float acc[CG];
#pragma unroll
      for (int k = 0; k < CG; ++k) acc[k] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (t.ix[j] >= 0 && t.ix[j] < X && t.iy[j] >= 0 && t.iy[j] < Y && t.iz[j] >= 0 &&
            t.iz[j] < Z) {
          const int flat = (t.ix[j] * Y + t.iy[j]) * Z + t.iz[j];
#pragma unroll
          for (int k = 0; k < CG; ++k)
            if (k < cg) acc[k] = __fadd_rn(acc[k], __fmul_rn(t.w[j], planes[k * V + flat]));
        }
      }
      float* dst = values + n * C + c0;
      if (CG >= 4 && cg == CG && ((C & 3) == 0)) {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int k = 0; k < CG / 4; ++k)
          d4[k] = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < CG; ++k)
          if (k < cg) dst[k] = acc[k];
      }
    }
  }
}

// ---- plane-staged backward for the reference layout [B,C,X,Y,Z] ----------------------------
// Mirror image of the forward above: a CTA owns (batch b, CG consecutive channel planes) in shared
// memory, zeroes them, lets every point of batch b add its 8 weighted corner contributions with
// SHARED-memory atomics, then streams the planes out with coalesced 16-byte stores.  Every element
// of the gradient grid is written exactly once: no memset of the 33 MB output, no global atomics
// (the simple kernel issues 8 P C of them into a channel-major grid: 16-32 M scattered float
// atomics at the model shapes).
template <int CG>
__global__ void __launch_bounds__(256)
k_interp_bwd_planes(const float* __restrict__ gvalues, const float* __restrict__ points,
                    const int* __restrict__ bi, long long P, int B, int C, int X, int Y, int Z,
                    float* __restrict__ gvox) {
  extern __shared__ __align__(16) float planes[];          // [CG][V]
  const int V = X * Y * Z;
  const int b = blockIdx.y, c0 = blockIdx.x * CG;
  const int cg = min(CG, C - c0);
  const bool vec = (V & 3) == 0;
  if (vec) {
    float4* d4 = reinterpret_cast<float4*>(planes);
    for (int e = threadIdx.x; e < cg * (V >> 2); e += 256) d4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int e = threadIdx.x; e < cg * V; e += 256) planes[e] = 0.f;
  }
  __syncthreads();
  for (long long base = 0; base < P; base += 256 * 8) {
    int bb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      long long n = base + u * 256 + threadIdx.x;
      bb[u] = (n < P) ? __ldg(bi + n) : -1;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (bb[u] != b) continue;
      const long long n = base + u * 256 + threadIdx.x;
      Tri t;
      trilinear(points[3 * n], points[3 * n + 1], points[3 * n + 2], t);
      float g[CG];
      const float* src = gvalues + n * C + c0;
      if (CG >= 4 && cg == CG && ((C & 3) == 0)) {
#pragma unroll
        for (int k = 0; k < CG / 4; ++k) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + k);
          g[4 * k] = v.x; g[4 * k + 1] = v.y; g[4 * k + 2] = v.z; g[4 * k + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < CG; ++k) g[k] = (k < cg) ? __ldg(src + k) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (t.ix[j] >= 0 && t.ix[j] < X && t.iy[j] >= 0 && t.iy[j] < Y && t.iz[j] >= 0 &&
            t.iz[j] < Z) {
          const int flat = (t.ix[j] * Y + t.iy[j]) * Z + t.iz[j];
#pragma unroll
          for (int k = 0; k < CG; ++k)
            if (k < cg) atomicAdd(planes + k * V + flat, __fmul_rn(t.w[j], g[k]));
        }
      }
    }
  }
  __syncthreads();
  float* dst = gvox + ((long long)b * C + c0) * V;          // the cg planes are contiguous
  if (vec) {
    const float4* s4 = reinterpret_cast<const float4*>(planes);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int e = threadIdx.x; e < cg * (V >> 2); e += 256) __stcs(d4 + e, s4[e]);
  } else {
    for (int e = threadIdx.x; e < cg * V; e += 256) dst[e] = planes[e];
  }
}

// channel planes per CTA: ~64 KB of shared memory (3 CTAs per SM) and enough CTAs to fill 148 SMs
static int planes_per_cta(long long V) {
  if (V * 4 > 160 * 1024) return 0;          // a plane does not fit: simple kernels
  if (V > 8192) return 1;
  if (V > 4096) return 2;
  if (V > 2048) return 4;
  if (V > 1024) return 8;
  return 16;
}

}  // namespace mf

using namespace mf;

#define MF_INTERP_PLANES(KERNEL, CGV, ...)                                              \
  do {                                                                                  \
    MF_ENSURE_DYN_SMEM(KERNEL<CGV>, 160 * 1024);                                        \
    dim3 grid((C + CGV - 1) / CGV, B);                                                  \
    KERNEL<CGV><<<grid, 256, (size_t)V * 4 * CGV, stream>>>(__VA_ARGS__);               \
  } while (0)

extern "C" int mf_interpolate_voxel_grid_fwd(const float* voxelized, const float* points,
                                             const int32_t* batch_indices, int64_t P, int B, int C,
                                             int X, int Y, int Z, int channels_last, float* values,
                                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0 || B <= 0 || C <= 0 || X <= 0 || Y <= 0 || Z <= 0) return MF_E_BADARG;
  if (P == 0) return MF_OK;
  if (!voxelized || !points || !batch_indices || !values) return MF_E_BADARG;
  const long long V = (long long)X * Y * Z;
  if (channels_last) {
    k_interp_fwd<true><<<div_up(P * C, 256), 256, 0, stream>>>(voxelized, points, batch_indices, P,
                                                               B, C, X, Y, Z, values);
  } else if (planes_per_cta(V) > 0 && P <= (1LL << 22) && B <= 65535) {
    // plane-staged path: CG channel planes per CTA in shared memory
    switch (planes_per_cta(V)) {
      case 1: MF_INTERP_PLANES(k_interp_fwd_planes, 1, voxelized, points, batch_indices, P, B, C, X, Y, Z, values); break;
      case 2: MF_INTERP_PLANES(k_interp_fwd_planes, 2, voxelized, points, batch_indices, P, B, C, X, Y, Z, values); break;
      case 4: MF_INTERP_PLANES(k_interp_fwd_planes, 4, voxelized, points, batch_indices, P, B, C, X, Y, Z, values); break;
      case 8: MF_INTERP_PLANES(k_interp_fwd_planes, 8, voxelized, points, batch_indices, P, B, C, X, Y, Z, values); break;
      default: MF_INTERP_PLANES(k_interp_fwd_planes, 16, voxelized, points, batch_indices, P, B, C, X, Y, Z, values); break;
    }
  } else {
    k_interp_fwd<false><<<div_up(P * C, 256), 256, 0, stream>>>(voxelized, points, batch_indices,
                                                                P, B, C, X, Y, Z, values);
  }
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_interpolate_voxel_grid_bwd(const float* gvalues, const float* points,
                                             const int32_t* batch_indices, int64_t P, int B, int C,
                                             int X, int Y, int Z, int channels_last,
                                             float* gvoxelized, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0 || B <= 0 || C <= 0 || X <= 0 || Y <= 0 || Z <= 0) return MF_E_BADARG;
  if (!gvoxelized) return MF_E_BADARG;
  if (P > 0 && (!gvalues || !points || !batch_indices)) return MF_E_BADARG;
  const long long V = (long long)X * Y * Z;
  if (!channels_last && P > 0 && planes_per_cta(V) > 0 && P <= (1LL << 22) && B <= 65535) {
    // plane-staged path: every output element written once, no memset, no global atomics
    switch (planes_per_cta(V)) {
      case 1: MF_INTERP_PLANES(k_interp_bwd_planes, 1, gvalues, points, batch_indices, P, B, C, X, Y, Z, gvoxelized); break;
      case 2: MF_INTERP_PLANES(k_interp_bwd_planes, 2, gvalues, points, batch_indices, P, B, C, X, Y, Z, gvoxelized); break;
      case 4: MF_INTERP_PLANES(k_interp_bwd_planes, 4, gvalues, points, batch_indices, P, B, C, X, Y, Z, gvoxelized); break;
      case 8: MF_INTERP_PLANES(k_interp_bwd_planes, 8, gvalues, points, batch_indices, P, B, C, X, Y, Z, gvoxelized); break;
      default: MF_INTERP_PLANES(k_interp_bwd_planes, 16, gvalues, points, batch_indices, P, B, C, X, Y, Z, gvoxelized); break;
    }
    MF_LAUNCH_CHECK();
    return MF_OK;
  }
  // fallback (channels-last grids, planes that do not fit shared memory): zero + global atomics
  MF_CUDA_TRY(cudaMemsetAsync(gvoxelized, 0, (size_t)B * C * X * Y * Z * 4, stream));
  if (P == 0) return MF_OK;
  if (channels_last)
    k_interp_bwd<true><<<div_up(P * C, 256), 256, 0, stream>>>(gvalues, points, batch_indices, P, B,
                                                               C, X, Y, Z, gvoxelized);
  else
    k_interp_bwd<false><<<div_up(P * C, 256), 256, 0, stream>>>(gvalues, points, batch_indices, P,
                                                                B, C, X, Y, Z, gvoxelized);
  MF_LAUNCH_CHECK();
  return MF_OK;
}
