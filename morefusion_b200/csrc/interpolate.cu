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
#include "tc_ptx.cuh"

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
    // the coordinates of every matching point of the chunk are requested before the first one is
    // used: one memory round trip per chunk instead of one per point
    float px[8], py[8], pz[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      px[u] = py[u] = pz[u] = 0.f;
      if (bb[u] == b) {
        const long long n = base + u * 256 + threadIdx.x;
        px[u] = __ldg(points + 3 * n); py[u] = __ldg(points + 3 * n + 1); pz[u] = __ldg(points + 3 * n + 2);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (bb[u] != b) continue;
      const long long n = base + u * 256 + threadIdx.x;
      Tri t;
      trilinear(px[u], py[u], pz[u], t);
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
    // coordinates and (for narrow channel groups) the gradient rows of every matching point of
    // the chunk are requested before the first one is used
    constexpr bool kPrefetchG = CG <= 4;
    float px[8], py[8], pz[8];
    float gpre[kPrefetchG ? 8 : 1][CG];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      px[u] = py[u] = pz[u] = 0.f;
      if (bb[u] == b) {
        const long long n = base + u * 256 + threadIdx.x;
        px[u] = __ldg(points + 3 * n); py[u] = __ldg(points + 3 * n + 1); pz[u] = __ldg(points + 3 * n + 2);
        if (kPrefetchG) {
#pragma unroll
          for (int k = 0; k < CG; ++k) gpre[kPrefetchG ? u : 0][k] = (k < cg) ? __ldg(gvalues + n * C + c0 + k) : 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (bb[u] != b) continue;
      const long long n = base + u * 256 + threadIdx.x;
      Tri t;
      trilinear(px[u], py[u], pz[u], t);
      float g[CG];
      const float* src = gvalues + n * C + c0;
      if (kPrefetchG) {
#pragma unroll
        for (int k = 0; k < CG; ++k) g[k] = gpre[kPrefetchG ? u : 0][k];
      } else if (CG >= 4 && cg == CG && ((C & 3) == 0)) {
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


// ---- streaming (persistent) plane kernels for the reference layout [B,C,X,Y,Z] ---------------
// The plane-staged kernels above are latency-bound: a CTA does load -> scan -> gather once and
// exits, 512 short dependent chains in 1.15 waves (ncu: 36 / 71 us, warps 33 % active, DRAM 33 MB).
// Here ONE CTA per SM walks a contiguous range of work items (batch b, CG channel planes),
// batch-major, with
//   * the planes of item k+1 arriving by cp.async.bulk (UBLKCP, mbarrier complete_tx) into the
//     second shared-memory buffer while item k is being gathered (forward), or the planes of item
//     k leaving by cp.async.bulk shared -> global while item k+1 accumulates (backward);
//   * the points of batch b found ONCE per batch (all batch-index loads of the scan in flight
//     together) and kept in shared memory (index + coordinates) for every item of that batch.
// Arithmetic per (point, channel) is unchanged (corner order j = 0..7, explicit _rn ops): the
// forward stays bit-identical.
constexpr int ST_THREADS = 512;
constexpr int ST_CHUNK = 2048;             // batch indices scanned per step (all loads in flight)
constexpr int ST_MAXP = 2 * ST_CHUNK;      // cached points per pass (index + xyz = 16 B each): a batch
                                           // of <= ST_CHUNK points is always cached whole

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

struct StreamArgs {
  const float* grid_in;    // fwd: voxelized            bwd: unused
  float* grid_out;         // fwd: unused               bwd: gvoxelized
  const float* rows_in;    // fwd: unused               bwd: gvalues [P,C]
  float* rows_out;         // fwd: values [P,C]         bwd: unused
  const float* points;
  const int* bi;
  long long P;
  int B, C, X, Y, Z;
  int n_groups, n_items;
};

struct PointCache {
  int n[ST_MAXP];
  float x[ST_MAXP], y[ST_MAXP], z[ST_MAXP];
  int count;
  int next;                // scan position after this pass
};

// points of batch b with index >= from, up to `cap` (<= ST_MAXP) of them, into the cache (any order)
template <int NT>
__device__ __forceinline__ void cache_points(PointCache& pc, const StreamArgs& a, int b, long long from,
                                             int cap = ST_MAXP) {
  constexpr int PER = ST_CHUNK / NT;
  __syncthreads();
  if (threadIdx.x == 0) { pc.count = 0; pc.next = (int)a.P; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // chunks of ST_CHUNK indices; a chunk is only taken if it is certain to fit
  for (long long base = from; base < a.P; base += ST_CHUNK) {
    // every thread reads the count, THEN the block votes (barrier): nobody can be past the vote
    // and already appending while another thread still reads, and the decision is uniform
    const int cnt = pc.count;
    if (__syncthreads_or(cnt + ST_CHUNK > cap && cnt > 0)) {
      if (threadIdx.x == 0) pc.next = (int)base;
      break;
    }
    int bb[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const long long n = base + u * NT + threadIdx.x;
      bb[u] = (n < a.P) ? __ldg(a.bi + n) : -1;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const bool m = bb[u] == b;
      const unsigned bal = __ballot_sync(0xffffffffu, m);
      if (bal) {
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&pc.count, __popc(bal));
        pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(bal & ((1u << lane) - 1u));
        if (m && pos < ST_MAXP) {
          const long long n = base + u * NT + threadIdx.x;
          pc.n[pos] = (int)n;
          pc.x[pos] = __ldg(a.points + 3 * n);
          pc.y[pos] = __ldg(a.points + 3 * n + 1);
          pc.z[pos] = __ldg(a.points + 3 * n + 2);
        }
      }
    }
    __syncthreads();                                   // pc.count is final for this chunk
  }
  __syncthreads();
}

// ---- cell-owner backward -----------------------------------------------------------------
// Scattering 8 P C products needs atomics, and shared-memory float atomics are CAS loops on this
// architecture (ATOMS.CAST.SPIN in the SASS; measured 115 / 290 us with the planes in shared
// memory, worse than the global-atomic kernel).  A per-voxel gather over cell lists has no atomics
// but wastes the warp (measured 151 / 186 us at 5.6 active threads per instruction: most lists are
// empty or one point long).  Here, per batch, the points are binned by their base cell (counting
// sort in shared memory, each cell's list in ascending point index) and the NON-EMPTY cells are
// listed; an item's planes live in shared memory and are filled in 8 phases, one per corner: in
// phase j a thread owns one non-empty cell and adds its points' w_j g products to the voxel
// (cell + offset_j) -- distinct cells hit distinct voxels within a phase, so plain read-modify-
// write is race-free; block barriers separate the phases.  No atomics, deterministic order
// (corner j = 0..7, then point index), every output element written exactly once, no memset.
constexpr int GB_THREADS = 512;
constexpr int GB_MAXP = 2048;            // cached points per pass
constexpr int GB_CHUNK = 1024;           // batch indices scanned per step: <= 1024 points per batch
                                         // are always cached whole
constexpr int GB_MAX_CELLS = 8192;       // (X+1)(Y+1)(Z+1) base cells that can touch the grid

struct GbCache {
  int n[GB_MAXP];
  float x[GB_MAXP], y[GB_MAXP], z[GB_MAXP];
  int perm[GB_MAXP];                     // cache slots in (cell, point index) order
  int tmp[GB_MAXP];                      // cache slots in cell order before the per-cell ranking
  int nz[GB_MAXP];                       // non-empty cells
  int count, next, n_nz;
};

template <int CG>
__global__ void __launch_bounds__(GB_THREADS, 1)
k_interp_bwd_cells(StreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int X = a.X, Y = a.Y, Z = a.Z, V = X * Y * Z;
  const int NC = (X + 1) * (Y + 1) * (Z + 1);
  const int VS = V + 4;                                           // plane stride: the CG channel threads of
                                                                  // a cell hit different banks
  float* planes = reinterpret_cast<float*>(smem_raw);             // [CG][VS]
  float* rows = planes + (size_t)CG * VS;                         // [GB_MAXP][CG] staged gradient rows
  GbCache& pc = *reinterpret_cast<GbCache*>(rows + (size_t)GB_MAXP * CG);
  int* cell_start = reinterpret_cast<int*>(&pc + 1);              // [NC + 1]
  int* cursor = cell_start + NC + 1;                              // [NC]
  __shared__ int warp_tot[GB_THREADS / 32], warp_nz[GB_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (a.n_items + gridDim.x - 1) / gridDim.x;
  const int it0 = blockIdx.x * per, it1 = min(a.n_items, it0 + per);
  if (it0 >= it1) return;
  int cached_b = -1;
  bool cache_whole = false;
  for (int item = it0; item < it1; ++item) {
    const int b = item / a.n_groups;
    const int c0 = (item - b * a.n_groups) * CG;
    const int cg = min(CG, a.C - c0);
    for (int e = tid; e < cg * (VS >> 2); e += GB_THREADS)
      reinterpret_cast<float4*>(planes)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    long long from = 0;
    for (;;) {
      if (!(cached_b == b && cache_whole)) {
        // ---- points of batch b from index `from` on (all batch-index loads of a step in flight)
        __syncthreads();
        if (tid == 0) { pc.count = 0; pc.next = (int)a.P; pc.n_nz = 0; }
        __syncthreads();
        for (long long base = from; base < a.P; base += GB_CHUNK) {
          // read, then vote with a barrier: uniform decision, no thread appends while another reads
          const int cnt = pc.count;
          if (__syncthreads_or(cnt + GB_CHUNK > GB_MAXP)) {
            if (tid == 0) pc.next = (int)base;
            break;
          }
          constexpr int PER = GB_CHUNK / GB_THREADS;
          int bb[PER];
#pragma unroll
          for (int u = 0; u < PER; ++u) {
            const long long n = base + u * GB_THREADS + tid;
            bb[u] = (n < a.P) ? __ldg(a.bi + n) : -1;
          }
#pragma unroll
          for (int u = 0; u < PER; ++u) {
            const bool m = bb[u] == b;
            const unsigned bal = __ballot_sync(0xffffffffu, m);
            if (bal) {
              int pos = 0;
              if (lane == 0) pos = atomicAdd(&pc.count, __popc(bal));
              pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(bal & ((1u << lane) - 1u));
              if (m && pos < GB_MAXP) {
                const long long n = base + u * GB_THREADS + tid;
                pc.n[pos] = (int)n;
                pc.x[pos] = __ldg(a.points + 3 * n);
                pc.y[pos] = __ldg(a.points + 3 * n + 1);
                pc.z[pos] = __ldg(a.points + 3 * n + 2);
              }
            }
          }
          __syncthreads();
        }
        __syncthreads();
        cached_b = b;
        cache_whole = (from == 0 && pc.next >= (int)a.P);
        // ---- bin the cached points by base cell
        const int count = pc.count;
        for (int c = tid; c < NC; c += GB_THREADS) cursor[c] = 0;
        __syncthreads();
        for (int e = tid; e < count; e += GB_THREADS) {
          const int x0 = (int)pc.x[e], y0 = (int)pc.y[e], z0 = (int)pc.z[e];
          if (x0 >= -1 && x0 < X && y0 >= -1 && y0 < Y && z0 >= -1 && z0 < Z)
            atomicAdd(cursor + ((x0 + 1) * (Y + 1) + (y0 + 1)) * (Z + 1) + (z0 + 1), 1);
        }
        __syncthreads();
        {   // exclusive scans over the NC cells (contiguous slice per thread, warp scan, warp
            // totals): of the counts -> cell_start, and of the "non-empty" flags -> the list of
            // non-empty cells in ascending cell order (a fixed task order = a fixed sum order)
          const int chunk = (NC + GB_THREADS - 1) / GB_THREADS;
          const int cb = tid * chunk, ce = min(NC, cb + chunk);
          int sum = 0, nzs = 0;
          for (int c = cb; c < ce; ++c) { sum += cursor[c]; nzs += cursor[c] > 0; }
          int incl = sum, incl_nz = nzs;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            const int v2 = __shfl_up_sync(0xffffffffu, incl_nz, d);
            if (lane >= d) { incl += v; incl_nz += v2; }
          }
          if (lane == 31) { warp_tot[warp] = incl; warp_nz[warp] = incl_nz; }
          __syncthreads();
          int off = incl - sum, off_nz = incl_nz - nzs;
          for (int w = 0; w < warp; ++w) { off += warp_tot[w]; off_nz += warp_nz[w]; }
          for (int c = cb; c < ce; ++c) {
            const int n_c = cursor[c];
            cell_start[c] = off;
            cursor[c] = off;
            off += n_c;
            if (n_c > 0) pc.nz[off_nz++] = c;
          }
          if (tid == GB_THREADS - 1) { cell_start[NC] = off; pc.n_nz = off_nz; }
          __syncthreads();
        }
        for (int e = tid; e < count; e += GB_THREADS) {
          const int x0 = (int)pc.x[e], y0 = (int)pc.y[e], z0 = (int)pc.z[e];
          if (x0 >= -1 && x0 < X && y0 >= -1 && y0 < Y && z0 >= -1 && z0 < Z)
            pc.tmp[atomicAdd(cursor + ((x0 + 1) * (Y + 1) + (y0 + 1)) * (Z + 1) + (z0 + 1), 1)] = e;
        }
        __syncthreads();
        // each cell's list in ascending point index (the sum order is then fixed): every point
        // ranks itself among its cell's points -- parallel over points, independent loads (an
        // insertion sort per cell took 150 us on clustered points: 100-point cells, dependent
        // shared-memory loads)
        for (int e = tid; e < count; e += GB_THREADS) {
          const int x0 = (int)pc.x[e], y0 = (int)pc.y[e], z0 = (int)pc.z[e];
          if (!(x0 >= -1 && x0 < X && y0 >= -1 && y0 < Y && z0 >= -1 && z0 < Z)) continue;
          const int c = ((x0 + 1) * (Y + 1) + (y0 + 1)) * (Z + 1) + (z0 + 1);
          const int s0 = cell_start[c], s1 = cell_start[c + 1], key = pc.n[e];
          int rank = 0;
          for (int i = s0; i < s1; ++i) rank += pc.n[pc.tmp[i]] < key;
          pc.perm[s0 + rank] = e;
        }
      }
      __syncthreads();
      // ---- stage the gradient rows of the cached points for this item's channels
      const int count = pc.count;
      for (int i = tid; i < count * CG; i += GB_THREADS) {
        const int e = i / CG, kk = i - e * CG;
        rows[i] = (kk < cg) ? __ldg(a.rows_in + (long long)pc.n[e] * a.C + c0 + kk) : 0.f;
      }
      __syncthreads();
      // ---- tasks = (non-empty cell, channel), channel fastest: a thread walks its cell's points
      // once, accumulating all 8 corner sums (corner order of the reference: w000 w100 w010 w001
      // w110 w011 w101 w111), then the 8 sums go to the planes in 8 phases: within a phase distinct
      // cells write distinct voxels, so the read-modify-write needs no atomic
      const int n_tasks = pc.n_nz * cg;
      for (int t0 = 0; t0 < n_tasks; t0 += GB_THREADS) {
        const int task = t0 + tid;
        const bool on = task < n_tasks;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        int x0 = 0, y0 = 0, z0 = 0, kk = 0;
        if (on) {
          const int ci = task / cg;
          kk = task - ci * cg;
          const int c = pc.nz[ci];
          z0 = c % (Z + 1) - 1; y0 = (c / (Z + 1)) % (Y + 1) - 1; x0 = c / ((Z + 1) * (Y + 1)) - 1;
          const int s0 = cell_start[c], s1 = cell_start[c + 1];
          for (int i = s0; i < s1; ++i) {
            const int e = pc.perm[i];
            const float lx = __fsub_rn(pc.x[e], (float)x0), ly = __fsub_rn(pc.y[e], (float)y0),
                        lz = __fsub_rn(pc.z[e], (float)z0);
            const float hx = __fsub_rn(1.f, lx), hy = __fsub_rn(1.f, ly), hz = __fsub_rn(1.f, lz);
            const float hh = __fmul_rn(hx, hy), lh = __fmul_rn(lx, hy), hl = __fmul_rn(hx, ly),
                        ll = __fmul_rn(lx, ly);
            const float g = rows[e * CG + kk];
            acc[0] = __fadd_rn(acc[0], __fmul_rn(__fmul_rn(hh, hz), g));
            acc[1] = __fadd_rn(acc[1], __fmul_rn(__fmul_rn(lh, hz), g));
            acc[2] = __fadd_rn(acc[2], __fmul_rn(__fmul_rn(hl, hz), g));
            acc[3] = __fadd_rn(acc[3], __fmul_rn(__fmul_rn(hh, lz), g));
            acc[4] = __fadd_rn(acc[4], __fmul_rn(__fmul_rn(ll, hz), g));
            acc[5] = __fadd_rn(acc[5], __fmul_rn(__fmul_rn(hl, lz), g));
            acc[6] = __fadd_rn(acc[6], __fmul_rn(__fmul_rn(lh, lz), g));
            acc[7] = __fadd_rn(acc[7], __fmul_rn(__fmul_rn(ll, lz), g));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int dx = (0xD2 >> j) & 1, dy = (0xB4 >> j) & 1, dz = (0xE8 >> j) & 1;
          const int ix = x0 + dx, iy = y0 + dy, iz = z0 + dz;
          if (on && ix >= 0 && ix < X && iy >= 0 && iy < Y && iz >= 0 && iz < Z) {
            float* dst = planes + (size_t)kk * VS + (ix * Y + iy) * Z + iz;
            *dst = __fadd_rn(*dst, acc[j]);
          }
          __syncthreads();
        }
      }
      if (pc.next >= (int)a.P) break;
      from = pc.next;
      cached_b = -1;
    }
    // ---- the item's planes are contiguous in the output: coalesced 16-byte stores
    float4* d4 = reinterpret_cast<float4*>(a.grid_out + ((long long)b * a.C + c0) * V);
    const int v4 = V >> 2;
    for (int e = tid; e < cg * v4; e += GB_THREADS) {
      const int kk = e / v4, r = e - kk * v4;
      __stcs(d4 + e, reinterpret_cast<const float4*>(planes + (size_t)kk * VS)[r]);
    }
    __syncthreads();
  }
}

template <int CG, bool kBackward>
__global__ void __launch_bounds__(ST_THREADS, 1)
k_interp_stream(StreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int V = a.X * a.Y * a.Z;
  float* buf0 = reinterpret_cast<float*>(smem_raw);
  float* buf1 = buf0 + (size_t)CG * V;
  PointCache& pc = *reinterpret_cast<PointCache*>(buf1 + (size_t)CG * V);
  uint64_t* full = reinterpret_cast<uint64_t*>(&pc + 1);          // [2]
  const int tid = threadIdx.x;
  // contiguous range of items for this CTA (batch-major: neighbours share the point cache)
  const int per = (a.n_items + gridDim.x - 1) / gridDim.x;
  const int it0 = blockIdx.x * per, it1 = min(a.n_items, it0 + per);
  if (it0 >= it1) return;
  if (!kBackward && tid == 0) {
    mbar_init(full, 1);
    mbar_init(full + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto item_planes = [&](int item, int* b, int* c0, int* cg) {
    *b = item / a.n_groups;
    *c0 = (item - *b * a.n_groups) * CG;
    *cg = min(CG, a.C - *c0);
  };
  if (!kBackward && tid == 0) {
    int b, c0, cg;
    item_planes(it0, &b, &c0, &cg);
    mbar_expect_tx(full, (uint32_t)(cg * V * 4));
    bulk_g2s(buf0, a.grid_in + ((long long)b * a.C + c0) * V, (uint32_t)(cg * V * 4), full);
  }
  int cached_b = -1;
  bool cache_whole = false;
  for (int item = it0; item < it1; ++item) {
    const int k = item - it0;
    float* buf = (k & 1) ? buf1 : buf0;
    int b, c0, cg;
    item_planes(item, &b, &c0, &cg);
    if (kBackward) {
      // the bulk store that last read this buffer (item k-2) must have finished reading it
      if (tid == 0) bulk_wait_read<1>();
      __syncthreads();
      float4* d4 = reinterpret_cast<float4*>(buf);
      for (int e = tid; e < cg * (V >> 2); e += ST_THREADS) d4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncthreads();
    } else {
      // request the next item's planes into the other buffer (its readers finished at the
      // barrier that ended item k-1)
      if (tid == 0 && item + 1 < it1) {
        int nb, nc0, ncg;
        item_planes(item + 1, &nb, &nc0, &ncg);
        float* nbuf = (k & 1) ? buf0 : buf1;
        mbar_expect_tx(full + ((k + 1) & 1), (uint32_t)(ncg * V * 4));
        bulk_g2s(nbuf, a.grid_in + ((long long)nb * a.C + nc0) * V, (uint32_t)(ncg * V * 4),
                 full + ((k + 1) & 1));
      }
    }
    bool waited = false;
    long long from = 0;
    for (;;) {                                   // passes over the batch's points (one, normally)
      if (!(cached_b == b && cache_whole)) {
        cache_points<ST_THREADS>(pc, a, b, from);
        cached_b = b;
        cache_whole = (from == 0 && pc.next >= (int)a.P);
      }
      if (!kBackward && !waited) {
        mbar_wait(full + (k & 1), (uint32_t)((k >> 1) & 1), nullptr, 0);
        waited = true;
      }
      const int count = pc.count;
      for (int e0 = tid; e0 < count; e0 += ST_THREADS * 4) {
        // up to 4 cached points per thread with their row loads in flight together (backward)
        float g[kBackward ? 4 : 1][CG];
        if (kBackward) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int e = e0 + q * ST_THREADS;
            if (e < count) {
              const float* src = a.rows_in + (long long)pc.n[e] * a.C + c0;
#pragma unroll
              for (int kk = 0; kk < CG; ++kk) g[kBackward ? q : 0][kk] = (kk < cg) ? __ldg(src + kk) : 0.f;
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = e0 + q * ST_THREADS;
          if (e >= count) continue;
          Tri t;
          trilinear(pc.x[e], pc.y[e], pc.z[e], t);
          if (kBackward) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (t.ix[j] >= 0 && t.ix[j] < a.X && t.iy[j] >= 0 && t.iy[j] < a.Y && t.iz[j] >= 0 &&
                  t.iz[j] < a.Z) {
                const int flat = (t.ix[j] * a.Y + t.iy[j]) * a.Z + t.iz[j];
#pragma unroll
                for (int kk = 0; kk < CG; ++kk)
                  if (kk < cg) atomicAdd(buf + kk * V + flat, __fmul_rn(t.w[j], g[kBackward ? q : 0][kk]));
              }
            }
          } else {
            float acc[CG];
#pragma unroll
            for (int kk = 0; kk < CG; ++kk) acc[kk] = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (t.ix[j] >= 0 && t.ix[j] < a.X && t.iy[j] >= 0 && t.iy[j] < a.Y && t.iz[j] >= 0 &&
                  t.iz[j] < a.Z) {
                const int flat = (t.ix[j] * a.Y + t.iy[j]) * a.Z + t.iz[j];
#pragma unroll
                for (int kk = 0; kk < CG; ++kk)
                  if (kk < cg) acc[kk] = __fadd_rn(acc[kk], __fmul_rn(t.w[j], buf[kk * V + flat]));
              }
            }
            float* dst = a.rows_out + (long long)pc.n[e] * a.C + c0;
            if (CG >= 4 && cg == CG && ((a.C & 3) == 0)) {
#pragma unroll
              for (int kk = 0; kk < CG / 4; ++kk)
                reinterpret_cast<float4*>(dst)[kk] =
                    make_float4(acc[4 * kk], acc[4 * kk + 1], acc[4 * kk + 2], acc[4 * kk + 3]);
            } else {
#pragma unroll
              for (int kk = 0; kk < CG; ++kk)
                if (kk < cg) dst[kk] = acc[kk];
            }
          }
        }
      }
      if (pc.next >= (int)a.P) break;
      from = pc.next;
      cached_b = -1;                             // partial cache: rebuild on the next pass / item
    }
    if (kBackward) {
      fence_async_smem();                        // generic-proxy writes -> visible to the bulk copy
      __syncthreads();
      if (tid == 0)
        bulk_s2g(a.grid_out + ((long long)b * a.C + c0) * V, buf, (uint32_t)(cg * V * 4));
    } else {
      __syncthreads();                           // everybody is done reading buf
    }
  }
  if (kBackward && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// channel planes per item of the streaming kernels: buffers of <= 64 KB and, if possible, at
// least three items per SM
static int stream_planes(long long V, int C, int B, int n_sm) {
  if ((V & 3) != 0 || V * 4 > 64 * 1024) return 0;
  int cg = 16;
  while (cg > 1 && (cg * V * 4 > 64 * 1024 || (long long)((C + cg - 1) / cg) * B * 2 < 3LL * n_sm)) cg >>= 1;
  return cg;
}

template <int CG, bool kBackward>
static int launch_stream(const StreamArgs& a0, int n_sm, cudaStream_t stream) {
  StreamArgs a = a0;
  a.n_groups = (a.C + CG - 1) / CG;
  a.n_items = a.n_groups * a.B;
  const size_t V = (size_t)a.X * a.Y * a.Z;
  const size_t smem = 2 * CG * V * 4 + sizeof(PointCache) + 2 * sizeof(uint64_t) + 128;
  MF_ENSURE_DYN_SMEM((k_interp_stream<CG, kBackward>), smem);
  const int grid = a.n_items < n_sm ? a.n_items : n_sm;
  k_interp_stream<CG, kBackward><<<grid, ST_THREADS, smem, stream>>>(a);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

static size_t bwd_gather_smem(int X, int Y, int Z, int cg) {
  const size_t NC = (size_t)(X + 1) * (Y + 1) * (Z + 1), V = (size_t)X * Y * Z;
  return (size_t)cg * (V + 4) * 4 + (size_t)GB_MAXP * cg * 4 + sizeof(GbCache) + (2 * NC + 1) * 4 + 128;
}

template <int CG>
static int launch_bwd_gather(const StreamArgs& a0, int n_sm, cudaStream_t stream) {
  StreamArgs a = a0;
  a.n_groups = (a.C + CG - 1) / CG;
  a.n_items = a.n_groups * a.B;
  const size_t smem = bwd_gather_smem(a.X, a.Y, a.Z, CG);
  MF_ENSURE_DYN_SMEM((k_interp_bwd_cells<CG>), smem);
  const int grid = a.n_items < n_sm ? a.n_items : n_sm;
  k_interp_bwd_cells<CG><<<grid, GB_THREADS, smem, stream>>>(a);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

static int dispatch_bwd_gather(int cg, const StreamArgs& a, int n_sm, cudaStream_t stream) {
  if (cg > 8) cg = 8;
  while (cg > 1 && bwd_gather_smem(a.X, a.Y, a.Z, cg) > 220 * 1024) cg >>= 1;
  switch (cg) {
    case 1: return launch_bwd_gather<1>(a, n_sm, stream);
    case 2: return launch_bwd_gather<2>(a, n_sm, stream);
    case 4: return launch_bwd_gather<4>(a, n_sm, stream);
    default: return launch_bwd_gather<8>(a, n_sm, stream);
  }
}

template <bool kBackward>
static int dispatch_stream(int cg, const StreamArgs& a, int n_sm, cudaStream_t stream) {
  switch (cg) {
    case 1: return launch_stream<1, kBackward>(a, n_sm, stream);
    case 2: return launch_stream<2, kBackward>(a, n_sm, stream);
    case 4: return launch_stream<4, kBackward>(a, n_sm, stream);
    case 8: return launch_stream<8, kBackward>(a, n_sm, stream);
    default: return launch_stream<16, kBackward>(a, n_sm, stream);
  }
}

// channel planes per CTA: ~32 KB of shared memory, i.e. 7 CTAs (56 warps) per SM and, at the model
// shapes, all CTAs of the launch resident in one wave (measured: 64 KB / 3 CTAs per SM ran 1.15
// waves of latency-bound CTAs at 33 % warp occupancy)
static int planes_per_cta(long long V) {
  if (V * 4 > 160 * 1024) return 0;          // a plane does not fit: simple kernels
  if (V > 4096) return 1;
  if (V > 2048) return 2;
  if (V > 1024) return 4;
  if (V > 512) return 8;
  return 16;
}

}  // namespace mf

using namespace mf;

static int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  return n;
}
static int stream_cg(long long V, int C, int B) { return stream_planes(V, C, B, sm_count()); }

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
  } else if (P < (1LL << 31) && stream_cg(V, C, B) > 0) {
    // persistent streaming path: bulk-copied planes double-buffered in shared memory
    StreamArgs a;
    a.grid_in = voxelized; a.grid_out = nullptr; a.rows_in = nullptr; a.rows_out = values;
    a.points = points; a.bi = batch_indices; a.P = P; a.B = B; a.C = C; a.X = X; a.Y = Y; a.Z = Z;
    return dispatch_stream<false>(stream_cg(V, C, B), a, sm_count(), stream);
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
  if (!channels_last && P > 0 && P < (1LL << 31) && stream_cg(V, C, B) > 0) {
    StreamArgs a;
    a.grid_in = nullptr; a.grid_out = gvoxelized; a.rows_in = gvalues; a.rows_out = nullptr;
    a.points = points; a.bi = batch_indices; a.P = P; a.B = B; a.C = C; a.X = X; a.Y = Y; a.Z = Z;
    if ((long long)(X + 1) * (Y + 1) * (Z + 1) <= GB_MAX_CELLS &&
        bwd_gather_smem(X, Y, Z, 1) <= 220 * 1024)
      return dispatch_bwd_gather(stream_cg(V, C, B), a, sm_count(), stream);
    return dispatch_stream<true>(stream_cg(V, C, B), a, sm_count(), stream);
  }
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
