// average_voxelization_3d / max_voxelization_3d for sm_100a.
//
// Replaces the reference's CuPy kernels K1-K4:
//   morefusion/functions/geometry/average_voxelization_3d.py:57-115, :163-218
//   morefusion/functions/geometry/max_voxelization_3d.py:75-138, :153-183
//
// average_voxelization_3d forward is a single-pass *output-tile* design:
// a CTA owns (batch b, channel chunk, VT consecutive voxels), gathers the points
// that fall into its voxel range into a shared-memory tile (ordered stream
// compaction -> per-voxel sums in ascending point order, no atomics, bit-exact
// against the oracle), divides by the counts and writes every output element
// exactly once with coalesced stores.  HBM traffic = output bytes + a 4-byte key
// per point per tile scan; no memset pass, no read-modify-write on HBM.
#include "common.cuh"

namespace mf {

constexpr int kThreads = 256;
constexpr int kListCap = 1024;           // keys scanned per round = 4 per thread
constexpr int kHdrSortedBit = 2;         // hdr[0] bit: set = batch_indices non-decreasing

struct VoxGeom {
  float ox, oy, oz, pitch;
  int X, Y, Z, B;
};

// ---------------------------------------------------------------- prepass
// key[n] = b*V + flat voxel index, or -1 if out of bounds.  Also: NaN flag,
// sortedness, per-batch [seg_start, seg_end) when sorted.
__global__ void k_vox_keys(const float* __restrict__ points, const int* __restrict__ bi,
                           long long N, VoxGeom g, int* __restrict__ keys,
                           int* __restrict__ hdr, int* __restrict__ seg_start,
                           int* __restrict__ seg_end, int* __restrict__ flags) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float x = points[3 * n], y = points[3 * n + 1], z = points[3 * n + 2];
  int b = bi[n];
  int f = 0;
  if (isnan(x) || isnan(y) || isnan(z)) f |= MF_FLAG_NAN_POINTS;
  int ix = voxel_coord(x, g.ox, g.pitch);
  int iy = voxel_coord(y, g.oy, g.pitch);
  int iz = voxel_coord(z, g.oz, g.pitch);
  bool okb = (b >= 0) && (b < g.B);
  if (!okb) f |= MF_FLAG_BAD_BATCH_INDEX;
  bool ok = okb && ix >= 0 && ix < g.X && iy >= 0 && iy < g.Y && iz >= 0 && iz < g.Z;
  keys[n] = ok ? (b * (g.X * g.Y * g.Z) + (ix * g.Y + iy) * g.Z + iz) : -1;
  int prev = (n > 0) ? bi[n - 1] : b;
  if (prev > b) {
    atomicAnd(&hdr[0], ~kHdrSortedBit);
    f |= MF_FLAG_UNSORTED_BATCH;
  }
  if (okb) {
    if (n == 0 || prev != b) seg_start[b] = (int)n;
    if (n == N - 1 || bi[n + 1] != b) seg_end[b] = (int)(n + 1);
  }
  if (f && flags) atomicOr(flags, f);
}

// ---------------------------------------------------------------- forward tiles
struct AvgParams {
  const float* values;
  const int* keys;
  const int* hdr;
  const int* seg_start;
  const int* seg_end;
  long long N;
  int C, B, V;
  int VT;       // voxels per tile
  int CC;       // channels per chunk
  int CCp;      // padded (odd) row length of the smem tile
  int G;        // thread groups per CTA (each owns voxels v % G == g)
  float* matrix;
  int* counts;
};

__global__ void __launch_bounds__(kThreads)
k_avg_tiles(AvgParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);          // [VT][CCp]
  int* cnt = reinterpret_cast<int*>(tile + (size_t)p.VT * p.CCp);  // [VT]
  int* list_n = cnt + p.VT;                                   // [kListCap]
  int* list_v = list_n + kListCap;                            // [kListCap]
  __shared__ int s_warp[kThreads / 32];
  __shared__ int s_total;
  __shared__ int s_any;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z, chunk = blockIdx.y;
  const int vbase = blockIdx.x * p.VT;
  const int vt = min(p.VT, p.V - vbase);
  const int c0 = chunk * p.CC;
  const int cc = min(p.CC, p.C - c0);

  for (int e = tid; e < p.VT * p.CCp; e += kThreads) tile[e] = 0.f;
  for (int e = tid; e < p.VT; e += kThreads) cnt[e] = 0;
  if (tid == 0) s_any = 0;

  long long lo = 0, hi = p.N;
  if (p.hdr[0] & kHdrSortedBit) {
    int s = p.seg_start[b];
    if (s < 0) { lo = hi = 0; } else { lo = s; hi = p.seg_end[b]; }
  }
  const int klo = b * p.V + vbase, khi = klo + vt;
  const int g = tid / p.CC, c = tid - g * p.CC;   // accumulate role
  const bool acc_thread = (g < p.G) && (c < cc);
  __syncthreads();

  for (long long base = lo; base < hi; base += kListCap) {
    // ---- ordered compaction of the keys in [base, base+kListCap)
    long long i0 = base + 4LL * tid;
    int k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = (i0 + j < hi) ? __ldg(p.keys + i0 + j) : -1;
    int m = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) m += (k[j] >= klo && k[j] < khi) ? 1 : 0;
    int incl = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) woff += (w < warp) ? s_warp[w] : 0;
    if (tid == kThreads - 1) s_total = woff + incl;
    int pos = woff + incl - m;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (k[j] >= klo && k[j] < khi) {
        list_n[pos] = (int)(i0 + j - 0);  // point id (fits int: N < 2^31 checked on host)
        list_v[pos] = k[j] - klo;
        ++pos;
      }
    __syncthreads();
    const int L = s_total;
    if (L > 0) {
      if (tid == 0) s_any = 1;
      if (acc_thread) {
        const float* vsrc = p.values + c0 + c;
        for (int l = 0; l < L; ++l) {
          int v = list_v[l];
          if ((v % p.G) == g) {
            int n = list_n[l];
            float val = __ldg(vsrc + (long long)n * p.C);
            tile[v * p.CCp + c] = __fadd_rn(tile[v * p.CCp + c], val);
            if (c == 0) cnt[v] += 1;
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- write the tile: every output element exactly once, coalesced along v
  const bool any = s_any != 0;
  float* out = p.matrix + ((long long)b * p.C + c0) * p.V + vbase;
  if (!any) {
    const bool vec = ((p.V & 3) == 0) && ((vt & 3) == 0) && ((vbase & 3) == 0);
    if (vec) {
      const int vt4 = vt >> 2;
      for (int e = tid; e < cc * vt4; e += kThreads) {
        int ch = e / vt4, v4 = e - ch * vt4;
        __stcs(reinterpret_cast<float4*>(out + (long long)ch * p.V) + v4,
               make_float4(0.f, 0.f, 0.f, 0.f));
      }
    } else {
      for (int e = tid; e < cc * vt; e += kThreads) {
        int ch = e / vt, v = e - ch * vt;
        out[(long long)ch * p.V + v] = 0.f;
      }
    }
  } else {
    for (int e = tid; e < cc * vt; e += kThreads) {
      int ch = e / vt, v = e - ch * vt;
      int n = cnt[v];
      float s = tile[v * p.CCp + ch];
      __stcs(out + (long long)ch * p.V + v, n > 0 ? __fdiv_rn(s, (float)n) : 0.f);
    }
  }
  if (chunk == 0) {
    int* oc = p.counts + (long long)b * p.V + vbase;
    for (int v = tid; v < vt; v += kThreads) oc[v] = any ? cnt[v] : 0;
  }
}

// ---------------------------------------------------------------- backward
// gvalues[n,c] = gmatrix[b,c,idx] / counts[b,idx]   (average_voxelization_3d.py:170-206)
__global__ void k_avg_bwd(const float* __restrict__ gmatrix, const int* __restrict__ counts,
                          const float* __restrict__ points, const int* __restrict__ bi,
                          long long N, int C, VoxGeom g, float* __restrict__ gvalues) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  long long n = i / C;
  int c = (int)(i - n * C);
  int b = bi[n];
  int ix = voxel_coord(points[3 * n], g.ox, g.pitch);
  int iy = voxel_coord(points[3 * n + 1], g.oy, g.pitch);
  int iz = voxel_coord(points[3 * n + 2], g.oz, g.pitch);
  float out = 0.f;
  if (b >= 0 && b < g.B && ix >= 0 && ix < g.X && iy >= 0 && iy < g.Y && iz >= 0 && iz < g.Z) {
    long long V = (long long)g.X * g.Y * g.Z;
    long long flat = ((long long)ix * g.Y + iy) * g.Z + iz;
    out = __fdiv_rn(gmatrix[((long long)b * C + c) * V + flat], (float)counts[b * V + flat]);
  }
  gvalues[i] = out;
}

// ---------------------------------------------------------------- max voxelization
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// key = (orderable(intensity) << 32) | ~n : atomicMax -> max intensity, lowest n on ties
__global__ void k_max_scatter(const float* __restrict__ points, const int* __restrict__ bi,
                              const float* __restrict__ inten, long long N, VoxGeom g,
                              unsigned long long* __restrict__ keys, int* __restrict__ flags) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float x = points[3 * n], y = points[3 * n + 1], z = points[3 * n + 2];
  int b = bi[n];
  int f = 0;
  if (isnan(x) || isnan(y) || isnan(z)) f |= MF_FLAG_NAN_POINTS;
  int ix = voxel_coord(x, g.ox, g.pitch);
  int iy = voxel_coord(y, g.oy, g.pitch);
  int iz = voxel_coord(z, g.oz, g.pitch);
  bool okb = (b >= 0) && (b < g.B);
  if (!okb) f |= MF_FLAG_BAD_BATCH_INDEX;
  if (okb && ix >= 0 && ix < g.X && iy >= 0 && iy < g.Y && iz >= 0 && iz < g.Z) {
    long long V = (long long)g.X * g.Y * g.Z;
    long long flat = ((long long)ix * g.Y + iy) * g.Z + iz;
    unsigned long long key =
        ((unsigned long long)f2ord(inten[n]) << 32) | (unsigned long long)(~(unsigned int)n);
    atomicMax(&keys[b * V + flat], key);
  }
  if (f && flags) atomicOr(flags, f);
}

__global__ void k_max_gather(const float* __restrict__ values,
                             const unsigned long long* __restrict__ keys, int C, long long V,
                             long long BV, float* __restrict__ matrix, int* __restrict__ indices) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over B*C*V
  if (i >= BV * C) return;
  long long v = i % V;
  long long bc = i / V;
  int c = (int)(bc % C);
  long long b = bc / C;
  unsigned long long key = keys[b * V + v];
  int n = key ? (int)(~(unsigned int)(key & 0xffffffffu)) : -1;
  matrix[i] = (n >= 0) ? values[(long long)n * C + c] : 0.f;
  if (c == 0) indices[b * V + v] = n;
}

// gvalues[n,c] = gmatrix[b,c,v] for the (single) voxel n wins (max_voxelization_3d.py:158-179)
__global__ void k_max_bwd(const float* __restrict__ gmatrix, const int* __restrict__ indices,
                          int C, long long V, long long BV, float* __restrict__ gvalues) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BV * C) return;
  long long v = i % V;
  long long bc = i / V;
  int c = (int)(bc % C);
  long long b = bc / C;
  int n = indices[b * V + v];
  if (n >= 0) gvalues[(long long)n * C + c] = gmatrix[i];  // a point lies in exactly one voxel
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace mf

using namespace mf;

extern "C" size_t mf_average_voxelization_3d_workspace_bytes(int64_t n_points) {
  // hdr[4] + seg_start/end for up to 65536 batches are carved from a fixed 512 KiB
  // header so the size depends on n_points only.
  return (size_t)(512 * 1024) + align_up((size_t)(n_points > 0 ? n_points : 1) * 4, 256);
}

extern "C" int mf_average_voxelization_3d_fwd(
    const float* values, const float* points, const int32_t* batch_indices, int64_t N,
    int C, int B, float ox, float oy, float oz, float pitch, int X, int Y, int Z,
    float* matrix, int32_t* counts, void* workspace, size_t workspace_bytes, int32_t* flags,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N < 0 || C <= 0 || B <= 0 || X <= 0 || Y <= 0 || Z <= 0) return MF_E_BADARG;
  if (!matrix || !counts || !workspace) return MF_E_BADARG;
  if (N > 0 && (!values || !points || !batch_indices)) return MF_E_BADARG;
  long long V = (long long)X * Y * Z;
  if (V * B >= (1LL << 31) || N >= (1LL << 31)) return MF_E_TOOLARGE;
  if (B > 65535) return MF_E_TOOLARGE;
  if (workspace_bytes < mf_average_voxelization_3d_workspace_bytes(N)) return MF_E_WORKSPACE;

  int* hdr = (int*)workspace;
  int* seg_start = hdr + 4;
  int* seg_end = seg_start + 65536;
  int* keys = (int*)((char*)workspace + 512 * 1024);
  // hdr = all ones (sorted bit set), seg_* = -1 (empty)
  MF_CUDA_TRY(cudaMemsetAsync(hdr, 0xFF, (4 + (size_t)B) * sizeof(int), stream));
  MF_CUDA_TRY(cudaMemsetAsync(seg_end, 0xFF, (size_t)B * sizeof(int), stream));
  VoxGeom g{ox, oy, oz, pitch, X, Y, Z, B};
  if (N > 0) {
    k_vox_keys<<<div_up(N, 256), 256, 0, stream>>>(points, batch_indices, N, g, keys, hdr,
                                                   seg_start, seg_end, flags);
    MF_LAUNCH_CHECK();
  }
  AvgParams p;
  p.values = values; p.keys = keys; p.hdr = hdr; p.seg_start = seg_start; p.seg_end = seg_end;
  p.N = N; p.C = C; p.B = B; p.V = (int)V;
  int nChunks = (C + 63) / 64;
  p.CC = (C + nChunks - 1) / nChunks;
  nChunks = (C + p.CC - 1) / p.CC;
  p.CCp = p.CC | 1;
  p.G = kThreads / p.CC;
  if (p.G < 1) p.G = 1;
  if (p.G > 32) p.G = 32;
  p.VT = 256;
  if (V < 256) p.VT = (int)V;
  p.matrix = matrix; p.counts = counts;
  size_t smem = (size_t)p.VT * p.CCp * 4 + (size_t)p.VT * 4 + (size_t)kListCap * 8;
  static bool attr_set = false;
  if (!attr_set) {
    MF_CUDA_TRY(cudaFuncSetAttribute(k_avg_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     100 * 1024));
    attr_set = true;
  }
  dim3 grid(div_up(V, p.VT), nChunks, B);
  k_avg_tiles<<<grid, kThreads, smem, stream>>>(p);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_average_voxelization_3d_bwd(
    const float* gmatrix, const int32_t* counts, const float* points,
    const int32_t* batch_indices, int64_t N, int C, int B, float ox, float oy, float oz,
    float pitch, int X, int Y, int Z, float* gvalues, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N < 0 || C <= 0 || B <= 0 || X <= 0 || Y <= 0 || Z <= 0) return MF_E_BADARG;
  if (N == 0) return MF_OK;
  if (!gmatrix || !counts || !points || !batch_indices || !gvalues) return MF_E_BADARG;
  VoxGeom g{ox, oy, oz, pitch, X, Y, Z, B};
  k_avg_bwd<<<div_up(N * C, 256), 256, 0, stream>>>(gmatrix, counts, points, batch_indices, N, C,
                                                    g, gvalues);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" size_t mf_max_voxelization_3d_workspace_bytes(int B, int X, int Y, int Z) {
  return (size_t)B * X * Y * Z * 8;
}

extern "C" int mf_max_voxelization_3d_fwd(
    const float* values, const float* points, const int32_t* batch_indices,
    const float* intensities, int64_t N, int C, int B, float ox, float oy, float oz, float pitch,
    int X, int Y, int Z, float* matrix, int32_t* indices, void* workspace,
    size_t workspace_bytes, int32_t* flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N < 0 || C <= 0 || B <= 0 || X <= 0 || Y <= 0 || Z <= 0) return MF_E_BADARG;
  if (!matrix || !indices || !workspace) return MF_E_BADARG;
  if (N > 0 && (!values || !points || !batch_indices || !intensities)) return MF_E_BADARG;
  if (N >= (1LL << 31)) return MF_E_TOOLARGE;
  long long V = (long long)X * Y * Z, BV = V * B;
  if (workspace_bytes < (size_t)BV * 8) return MF_E_WORKSPACE;
  unsigned long long* keys = (unsigned long long*)workspace;
  MF_CUDA_TRY(cudaMemsetAsync(keys, 0, (size_t)BV * 8, stream));
  VoxGeom g{ox, oy, oz, pitch, X, Y, Z, B};
  if (N > 0) {
    k_max_scatter<<<div_up(N, 256), 256, 0, stream>>>(points, batch_indices, intensities, N, g,
                                                      keys, flags);
    MF_LAUNCH_CHECK();
  }
  k_max_gather<<<div_up(BV * C, 256), 256, 0, stream>>>(values, keys, C, V, BV, matrix, indices);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_max_voxelization_3d_bwd(const float* gmatrix, const int32_t* indices, int64_t N,
                                          int C, int B, int X, int Y, int Z, float* gvalues,
                                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N < 0 || C <= 0 || B <= 0 || X <= 0 || Y <= 0 || Z <= 0) return MF_E_BADARG;
  if (N == 0) return MF_OK;
  if (!gmatrix || !indices || !gvalues) return MF_E_BADARG;
  long long V = (long long)X * Y * Z, BV = V * B;
  MF_CUDA_TRY(cudaMemsetAsync(gvalues, 0, (size_t)N * C * 4, stream));
  k_max_bwd<<<div_up(BV * C, 256), 256, 0, stream>>>(gmatrix, indices, C, V, BV, gvalues);
  MF_LAUNCH_CHECK();
  return MF_OK;
}
