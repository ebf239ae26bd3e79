// truncated_distance_function, pseudo_occupancy_voxelization, occupancy_grid_3d (sm_100a).
//
// Replaces the reference's CuPy kernels K5/K6 and the cupy/chainer op chains around them:
//   morefusion/functions/geometry/truncated_distance_function.py:21-101 (fwd),
//   :103-166 (bwd), :181-213 (pseudo occupancy weights)
//   morefusion/functions/geometry/occupancy_grid_3d.py:31-85
//
// The reference's winner selection (float atomicMin followed by a non-atomic
// atomicExch, :72-78) is racy; here distance and point id are packed into one
// 64-bit key and reduced with a single atomicMin, i.e. winner = lexicographic
// min (distance, point id) -- deterministic and equal to the oracle's tie-break.
#include "common.cuh"

namespace mf {

struct GridGeom {
  float ox, oy, oz, pitch;
  int X, Y, Z;
};

__host__ __device__ inline int tdf_ksize(float pitch, float truncation) {
  int k = (int)ceilf(truncation / pitch);   // truncated_distance_function.py:36-38
  if ((k & 1) == 0) k += 1;
  return k;
}

constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

// one thread per (point, neighbour offset)
__global__ void k_tdf_scatter(const float* __restrict__ points, long long P, GridGeom g,
                              float trunc, int ks, unsigned long long* __restrict__ keys) {
  const int K = ks * ks * ks;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * K) return;
  long long p = i / K;
  int k = (int)(i - p * K);
  int half = ks / 2;
  int dz = k % ks - half, dy = (k / ks) % ks - half, dx = k / (ks * ks) - half;
  float fx = voxel_coord_f(points[3 * p], g.ox, g.pitch);
  float fy = voxel_coord_f(points[3 * p + 1], g.oy, g.pitch);
  float fz = voxel_coord_f(points[3 * p + 2], g.oz, g.pitch);
  // int ix = round(ix_f) + kernel[3k]  (float add, then conversion; :60-62)
  int ix = static_cast<int>(__fadd_rn(roundf(fx), (float)dx));
  int iy = static_cast<int>(__fadd_rn(roundf(fy), (float)dy));
  int iz = static_cast<int>(__fadd_rn(roundf(fz), (float)dz));
  if (ix < 0 || ix >= g.X || iy < 0 || iy >= g.Y || iz < 0 || iz >= g.Z) return;
  float ddx = __fsub_rn(fx, (float)ix), ddy = __fsub_rn(fy, (float)iy), ddz = __fsub_rn(fz, (float)iz);
  float dist = __fmul_rn(g.pitch, __fsqrt_rn(sq3(ddx, ddy, ddz)));
  if (dist < trunc) {
    unsigned long long key =
        ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)(unsigned int)p;
    atomicMin(&keys[((long long)ix * g.Y + iy) * g.Z + iz], key);
  }
}

__global__ void k_tdf_finalize(const unsigned long long* __restrict__ keys, long long V,
                               float trunc, float* __restrict__ tdf, int* __restrict__ indices) {
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  unsigned long long key = keys[v];
  bool hit = key != kEmptyKey;
  tdf[v] = hit ? __uint_as_float((unsigned int)(key >> 32)) : trunc;
  indices[v] = hit ? (int)(unsigned int)(key & 0xffffffffu) : -1;
}

// one thread per point: gather the gradient from the voxels this point won, in ascending
// voxel order (deterministic; the reference scatters with atomicAdd, :141-145)
__global__ void k_tdf_bwd(const float* __restrict__ gtdf, const float* __restrict__ points,
                          const int* __restrict__ indices, long long P, GridGeom g, int ks,
                          float* __restrict__ gpoints) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float fx = voxel_coord_f(points[3 * p], g.ox, g.pitch);
  float fy = voxel_coord_f(points[3 * p + 1], g.oy, g.pitch);
  float fz = voxel_coord_f(points[3 * p + 2], g.oz, g.pitch);
  float rx = roundf(fx), ry = roundf(fy), rz = roundf(fz);
  int half = ks / 2;
  float ax = 0.f, ay = 0.f, az = 0.f;
  for (int dx = -half; dx <= half; ++dx) {
    int ix = static_cast<int>(__fadd_rn(rx, (float)dx));
    if (ix < 0 || ix >= g.X) continue;
    for (int dy = -half; dy <= half; ++dy) {
      int iy = static_cast<int>(__fadd_rn(ry, (float)dy));
      if (iy < 0 || iy >= g.Y) continue;
      for (int dz = -half; dz <= half; ++dz) {
        int iz = static_cast<int>(__fadd_rn(rz, (float)dz));
        if (iz < 0 || iz >= g.Z) continue;
        long long flat = ((long long)ix * g.Y + iy) * g.Z + iz;
        if (indices[flat] != (int)p) continue;
        float ddx = __fsub_rn(fx, (float)ix), ddy = __fsub_rn(fy, (float)iy),
              ddz = __fsub_rn(fz, (float)iz);
        float n = __fsqrt_rn(sq3(ddx, ddy, ddz));
        if (n > 0.f) {
          float gg = gtdf[flat];
          ax = __fadd_rn(ax, __fmul_rn(__fdiv_rn(ddx, n), gg));
          ay = __fadd_rn(ay, __fmul_rn(__fdiv_rn(ddy, n), gg));
          az = __fadd_rn(az, __fmul_rn(__fdiv_rn(ddz, n), gg));
        }
      }
    }
  }
  gpoints[3 * p] = ax;
  gpoints[3 * p + 1] = ay;
  gpoints[3 * p + 2] = az;
}

// ---- pseudo occupancy (truncated_distance_function.py:181-213)
// stage 1: decode winners, raw inside-weight, grid-wide max
__global__ void k_po_stage1(const unsigned long long* __restrict__ keys,
                            const float* __restrict__ sdf, long long V, float trunc,
                            float sdf_offset, float* __restrict__ tdf, int* __restrict__ indices,
                            float* __restrict__ w_inside, float* __restrict__ w_surface,
                            unsigned int* __restrict__ wmax_bits) {
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float w = 0.f;
  if (v < V) {
    unsigned long long key = keys[v];
    bool hit = key != kEmptyKey;
    int idx = hit ? (int)(unsigned int)(key & 0xffffffffu) : -1;
    tdf[v] = hit ? __uint_as_float((unsigned int)(key >> 32)) : trunc;
    indices[v] = idx;
    float wr = __fadd_rn(hit ? sdf[idx] : -1.f, sdf_offset);
    bool mask = wr < 0.f;
    w = mask ? 0.f : wr;
    w_inside[v] = w;
    w_surface[v] = mask ? 1.f : 0.f;   // mask bit, consumed by stage 2
  }
  float m = warp_max(w);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float mm = (threadIdx.x < (blockDim.x >> 5)) ? s[threadIdx.x] : 0.f;
    mm = warp_max(mm);
    if (threadIdx.x == 0 && mm > 0.f) atomicMax(wmax_bits, __float_as_uint(mm));
  }
}

__global__ void k_po_stage2(long long V, float trunc, const float* __restrict__ tdf,
                            const unsigned int* __restrict__ wmax_bits,
                            float* __restrict__ w_inside, float* __restrict__ w_surface,
                            float* __restrict__ grid, float* __restrict__ surface,
                            float* __restrict__ inside) {
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float wmax = __uint_as_float(*wmax_bits);
  float wi = __fdiv_rn(w_inside[v], wmax);           // 0/0 -> NaN by design (:204)
  bool mask = w_surface[v] != 0.f;
  float ws = mask ? wi : __fsub_rn(1.f, wi);
  float gr = __fsub_rn(1.f, __fdiv_rn(tdf[v], trunc));
  w_inside[v] = wi;
  w_surface[v] = ws;
  grid[v] = gr;
  surface[v] = __fmul_rn(gr, ws);
  inside[v] = __fmul_rn(gr, wi);
}

// ---- occupancy_grid_3d (occupancy_grid_3d.py:31-85): matrix-free min over points
constexpr int kOccTile = 256;

__global__ void k_occ_fwd(const float* __restrict__ points, long long P, GridGeom g,
                          float threshold, float* __restrict__ grid, float* __restrict__ dmin_out) {
  __shared__ float sq[kOccTile * 3];
  long long V = (long long)g.X * g.Y * g.Z;
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int iz = (int)(v % g.Z), iy = (int)((v / g.Z) % g.Y), ix = (int)(v / ((long long)g.Y * g.Z));
  // min_p sqrt(s_p) == sqrt(min_p s_p) bit-for-bit (correctly rounded sqrt is monotone), so the
  // IEEE square root is taken once per voxel instead of once per (voxel, point)
  float smin = __int_as_float(0x7f800000);
  for (long long base = 0; base < P; base += kOccTile) {
    int n = (int)min((long long)kOccTile, P - base);
    for (int e = threadIdx.x; e < n * 3; e += blockDim.x) {
      int a = e % 3;
      float o = a == 0 ? g.ox : (a == 1 ? g.oy : g.oz);
      sq[e] = voxel_coord_f(points[base * 3 + e], o, g.pitch);
    }
    __syncthreads();
    if (v < V)
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        float d0 = __fsub_rn((float)ix, sq[3 * j]), d1 = __fsub_rn((float)iy, sq[3 * j + 1]),
              d2 = __fsub_rn((float)iz, sq[3 * j + 2]);
        smin = fminf(smin, sq3(d0, d1, d2));
      }
    __syncthreads();
  }
  if (v < V) {
    const float dmin = __fsqrt_rn(smin);
    dmin_out[v] = dmin;
    float m = fmaxf(__fsub_rn(threshold, dmin), 0.f);
    grid[v] = fminf(m, 1.f);
  }
}

// F.minimum(m,1): grad where m<=1; relu: where threshold-d>0; F.min: to ALL tied minima;
// sqrt: d0/d; OccupancyGrid3D.backward: -1/pitch (:56-74)
__global__ void k_occ_bwd(const float* __restrict__ ggrid, const float* __restrict__ dmin_in,
                          const float* __restrict__ points, long long P, GridGeom g,
                          float threshold, float* __restrict__ gpoints) {
  __shared__ float sq[kOccTile * 3];
  long long V = (long long)g.X * g.Y * g.Z;
  long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int iz = (int)(v % g.Z), iy = (int)((v / g.Z) % g.Y), ix = (int)(v / ((long long)g.Y * g.Z));
  float dmin = 0.f, gd = 0.f, s_hi = 0.f;
  bool active = false;
  if (v < V) {
    dmin = dmin_in[v];
    s_hi = dmin * dmin * 1.000001f + 1e-30f;
    float r = __fsub_rn(threshold, dmin);
    active = (r > 0.f) && (r <= 1.f);
    gd = -ggrid[v];
  }
  for (long long base = 0; base < P; base += kOccTile) {
    int n = (int)min((long long)kOccTile, P - base);
    for (int e = threadIdx.x; e < n * 3; e += blockDim.x) {
      int a = e % 3;
      float o = a == 0 ? g.ox : (a == 1 ? g.oy : g.oz);
      sq[e] = voxel_coord_f(points[base * 3 + e], o, g.pitch);
    }
    __syncthreads();
    if (active)
      for (int j = 0; j < n; ++j) {
        float d0 = __fsub_rn((float)ix, sq[3 * j]), d1 = __fsub_rn((float)iy, sq[3 * j + 1]),
              d2 = __fsub_rn((float)iz, sq[3 * j + 2]);
        const float s2 = sq3(d0, d1, d2);
        if (s2 > s_hi) continue;                   // cannot round to dmin: skip the IEEE sqrt
        float d = __fsqrt_rn(s2);
        if (d == dmin) {
          float inv = __fdiv_rn(gd, d);
          float* gp = gpoints + (base + j) * 3;
          atomicAdd(gp + 0, -__fdiv_rn(__fmul_rn(inv, d0), g.pitch));
          atomicAdd(gp + 1, -__fdiv_rn(__fmul_rn(inv, d1), g.pitch));
          atomicAdd(gp + 2, -__fdiv_rn(__fmul_rn(inv, d2), g.pitch));
        }
      }
    __syncthreads();
  }
}

}  // namespace mf

using namespace mf;

extern "C" size_t mf_truncated_distance_function_workspace_bytes(int X, int Y, int Z) {
  return (size_t)X * Y * Z * 8 + 256;
}
extern "C" size_t mf_pseudo_occupancy_voxelization_workspace_bytes(int X, int Y, int Z) {
  return (size_t)X * Y * Z * 8 + 256;
}

static int tdf_scatter(const float* points, int64_t P, GridGeom g, float truncation,
                       unsigned long long* keys, cudaStream_t stream) {
  long long V = (long long)g.X * g.Y * g.Z;
  MF_CUDA_TRY(cudaMemsetAsync(keys, 0xFF, (size_t)V * 8, stream));
  int ks = tdf_ksize(g.pitch, truncation);
  long long K = (long long)ks * ks * ks;
  if (P * K >= (1LL << 40)) return MF_E_TOOLARGE;
  if (P > 0) {
    k_tdf_scatter<<<div_up(P * K, 256), 256, 0, stream>>>(points, P, g, truncation, ks, keys);
    MF_LAUNCH_CHECK();
  }
  return MF_OK;
}

extern "C" int mf_truncated_distance_function_fwd(
    const float* points, int64_t P, float pitch, float ox, float oy, float oz, int X, int Y, int Z,
    float truncation, float* tdf, int32_t* indices, void* workspace, size_t workspace_bytes,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0 || X <= 0 || Y <= 0 || Z <= 0 || !(pitch > 0.f)) return MF_E_BADARG;
  if (!tdf || !indices || !workspace || (P > 0 && !points)) return MF_E_BADARG;
  if (P >= (1LL << 31)) return MF_E_TOOLARGE;
  long long V = (long long)X * Y * Z;
  if (workspace_bytes < (size_t)V * 8) return MF_E_WORKSPACE;
  GridGeom g{ox, oy, oz, pitch, X, Y, Z};
  unsigned long long* keys = (unsigned long long*)workspace;
  int rc = tdf_scatter(points, P, g, truncation, keys, stream);
  if (rc) return rc;
  k_tdf_finalize<<<div_up(V, 256), 256, 0, stream>>>(keys, V, truncation, tdf, indices);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_truncated_distance_function_bwd(
    const float* gtdf, const float* points, const int32_t* indices, int64_t P, float pitch,
    float ox, float oy, float oz, int X, int Y, int Z, float truncation, float* gpoints,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0 || X <= 0 || Y <= 0 || Z <= 0 || !(pitch > 0.f)) return MF_E_BADARG;
  if (P == 0) return MF_OK;
  if (!gtdf || !points || !indices || !gpoints) return MF_E_BADARG;
  GridGeom g{ox, oy, oz, pitch, X, Y, Z};
  int ks = tdf_ksize(pitch, truncation);
  k_tdf_bwd<<<div_up(P, 128), 128, 0, stream>>>(gtdf, points, indices, P, g, ks, gpoints);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_pseudo_occupancy_voxelization_fwd(
    const float* points, const float* sdf, int64_t P, float pitch, float ox, float oy, float oz,
    int X, int Y, int Z, float threshold, float sdf_offset, float* grid, float* grid_surface,
    float* grid_inside, float* tdf, int32_t* indices, float* w_surface, float* w_inside,
    void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0 || X <= 0 || Y <= 0 || Z <= 0 || !(pitch > 0.f)) return MF_E_BADARG;
  if (!grid || !grid_surface || !grid_inside || !tdf || !indices || !w_surface || !w_inside ||
      !workspace || (P > 0 && (!points || !sdf)))
    return MF_E_BADARG;
  if (P >= (1LL << 31)) return MF_E_TOOLARGE;
  long long V = (long long)X * Y * Z;
  if (workspace_bytes < (size_t)V * 8 + 256) return MF_E_WORKSPACE;
  GridGeom g{ox, oy, oz, pitch, X, Y, Z};
  float trunc = threshold * pitch;   // :184
  unsigned long long* keys = (unsigned long long*)workspace;
  unsigned int* wmax = (unsigned int*)((char*)workspace + (size_t)V * 8);
  MF_CUDA_TRY(cudaMemsetAsync(wmax, 0, 4, stream));
  int rc = tdf_scatter(points, P, g, trunc, keys, stream);
  if (rc) return rc;
  k_po_stage1<<<div_up(V, 256), 256, 0, stream>>>(keys, sdf, V, trunc, sdf_offset, tdf, indices,
                                                  w_inside, w_surface, wmax);
  MF_LAUNCH_CHECK();
  k_po_stage2<<<div_up(V, 256), 256, 0, stream>>>(V, trunc, tdf, wmax, w_inside, w_surface, grid,
                                                  grid_surface, grid_inside);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_occupancy_grid_3d_fwd(const float* points, int64_t P, float pitch, float ox,
                                        float oy, float oz, int X, int Y, int Z, float threshold,
                                        float* grid, float* dmin, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0 || X <= 0 || Y <= 0 || Z <= 0) return MF_E_BADARG;
  if (!grid || !dmin || (P > 0 && !points)) return MF_E_BADARG;
  GridGeom g{ox, oy, oz, pitch, X, Y, Z};
  long long V = (long long)X * Y * Z;
  k_occ_fwd<<<div_up(V, 256), 256, 0, stream>>>(points, P, g, threshold, grid, dmin);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_occupancy_grid_3d_bwd(const float* ggrid, const float* dmin, const float* points,
                                        int64_t P, float pitch, float ox, float oy, float oz,
                                        int X, int Y, int Z, float threshold, float* gpoints,
                                        void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (P < 0 || X <= 0 || Y <= 0 || Z <= 0) return MF_E_BADARG;
  if (P == 0) return MF_OK;
  if (!ggrid || !dmin || !points || !gpoints) return MF_E_BADARG;
  GridGeom g{ox, oy, oz, pitch, X, Y, Z};
  long long V = (long long)X * Y * Z;
  MF_CUDA_TRY(cudaMemsetAsync(gpoints, 0, (size_t)P * 12, stream));
  k_occ_bwd<<<div_up(V, 256), 256, 0, stream>>>(ggrid, dmin, points, P, g, threshold, gpoints);
  MF_LAUNCH_CHECK();
  return MF_OK;
}
