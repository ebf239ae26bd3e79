// average_voxelization_3d / max_voxelization_3d for sm_100a.
//
// Replaces the reference's CuPy kernels K1-K4:
//   morefusion/functions/geometry/average_voxelization_3d.py:57-115, :163-218
//   morefusion/functions/geometry/max_voxelization_3d.py:75-138, :153-183
//
// average_voxelization_3d forward:
//   k_avg_keys_fill  prepass (keys, tile counts, batch segments, mode) fused into a dense,
//                    look-up-free zero fill of matrix/counts (runs at the HBM write roofline)
//   k_avg_leaders    fast mode (sorted batch indices): one warp per point; the lowest-index point
//                    of a voxel sums the voxel's points in ascending order and overwrites it
//   k_avg_tiles      general mode / ragged shapes: a CTA owns (batch, channel chunk, 256 voxels),
//                    ordered stream compaction into a shared-memory tile, coalesced tile write
// No atomics on the output, per-voxel sums in ascending point order -> bit-exact vs the oracle.
#include "common.cuh"

namespace mf {

constexpr int kThreads = 256;
constexpr int kListCap = 1024;           // keys scanned per round = 4 per thread
constexpr int kStage = 64;               // list entries whose values are staged in smem at once
constexpr int kHdrSortedBit = 2;         // hdr[0] bit: set = batch_indices non-decreasing

struct VoxGeom {
  float ox, oy, oz, pitch;
  int X, Y, Z, B;
};

constexpr int kLeaderMaxSeg = 4096;      // longest batch segment the per-point leader path scans
constexpr int kLeaderSmemKeys = 8192;    // keys a leader CTA can stage in shared memory

__device__ __forceinline__ void vox_keys_point(
    const float* __restrict__ points, const int* __restrict__ bi, long long N, VoxGeom g,
    int* __restrict__ keys, int* __restrict__ hdr, int* __restrict__ seg_start,
    int* __restrict__ seg_end, int* __restrict__ flags, int* __restrict__ tile_count, int VT,
    int* __restrict__ flags2, int* __restrict__ tile_list, int* __restrict__ tile_list_n,
    unsigned int* __restrict__ occ_bits, long long n);

// ---------------------------------------------------------------- prepass
// key[n] = b*V + flat voxel index, or -1 if out of bounds.  Also: NaN flag,
// sortedness, per-batch [seg_start, seg_end) when sorted.
__global__ void k_vox_keys(const float* __restrict__ points, const int* __restrict__ bi,
                           long long N, VoxGeom g, int* __restrict__ keys,
                           int* __restrict__ hdr, int* __restrict__ seg_start,
                           int* __restrict__ seg_end, int* __restrict__ flags,
                           int* __restrict__ tile_count, int VT, int* __restrict__ flags2,
                           int* __restrict__ tile_list, int* __restrict__ tile_list_n,
                           unsigned int* __restrict__ occ_bits, int* __restrict__ done,
                           int* __restrict__ mode) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) vox_keys_point(points, bi, N, g, keys, hdr, seg_start, seg_end, flags, tile_count, VT,
                            flags2, tile_list, tile_list_n, occ_bits, n);
  // ---- the last block to finish decides the forward mode: "fast" (per-point leader scatter)
  // needs sorted batch indices and short segments; anything else takes the general tile kernel
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return;
  __threadfence();
  int maxlen = 0;
  for (int b = threadIdx.x; b < g.B; b += 32) {
    int s0 = __ldcg(seg_start + b), s1 = __ldcg(seg_end + b);
    if (s0 >= 0) maxlen = max(maxlen, s1 - s0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  if (threadIdx.x == 0)
    *mode = ((__ldcg(hdr) & kHdrSortedBit) && maxlen <= kLeaderMaxSeg) ? 1 : 0;
}

__device__ __forceinline__ void vox_keys_point(
    const float* __restrict__ points, const int* __restrict__ bi, long long N, VoxGeom g,
    int* __restrict__ keys, int* __restrict__ hdr, int* __restrict__ seg_start,
    int* __restrict__ seg_end, int* __restrict__ flags, int* __restrict__ tile_count, int VT,
    int* __restrict__ flags2, int* __restrict__ tile_list, int* __restrict__ tile_list_n,
    unsigned int* __restrict__ occ_bits, long long n) {
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
  const int V = g.X * g.Y * g.Z;
  const int flat = (ix * g.Y + iy) * g.Z + iz;
  keys[n] = ok ? (b * V + flat) : -1;
  if (ok) {
    const int tile = b * ((V + VT - 1) / VT) + flat / VT;
    // the first point to hit a tile appends it to the list of occupied tiles
    if (atomicAdd(&tile_count[tile], 1) == 0) tile_list[atomicAdd(tile_list_n, 1)] = tile;
    const int key = b * V + flat;
    if (occ_bits) atomicOr(&occ_bits[key >> 5], 1u << (key & 31));
  }
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
  if (f && flags2) atomicOr(flags2, f);
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
  const int* tile_count;   // [B][tiles_per_batch] points per voxel tile (from the prepass)
  const int* tile_list;    // occupied tiles (unordered) and their number
  const int* tile_list_n;
  const int* mode;         // 1 = fast (leader scatter), 0 = general (tile kernel)
  int n_lead_ctas, zero_groups;
  int tiles_per_batch, n_chunks, n_items;
};

// Prepass + dense zero fill in ONE launch.  CTAs [0, n_key_ctas) run the key / segment /
// tile-count prepass (k_vox_keys' body; the last of them picks the forward mode), every other CTA
// streams 16 KiB of zeros into matrix / counts with no look-ups at all, so the bulk of the
// operator -- writing the dense, mostly-zero output -- runs at plain-fill speed while the
// latency-bound prepass hides inside the first wave.  The occupied voxels are overwritten
// afterwards by the leader scatter (fast mode) or the tile kernel (general mode); that is
// N*(C+1) extra element writes (3 % of the output at the model shape).
struct KeysArgs {
  const float* points;
  const int* bi;
  VoxGeom g;
  int* keys;
  int* hdr;
  int* seg_start;
  int* seg_end;
  int* flags;
  int* tile_count;
  int* flags2;
  int* tile_list;
  int* tile_list_n;
  int* done;
  int* mode;
  int n_key_ctas;
};

__global__ void __launch_bounds__(256)
k_avg_keys_fill(AvgParams p, KeysArgs ka) {
  if ((int)blockIdx.x < ka.n_key_ctas) {
    long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < p.N)
      vox_keys_point(ka.points, ka.bi, p.N, ka.g, ka.keys, ka.hdr, ka.seg_start, ka.seg_end,
                     ka.flags, ka.tile_count, p.VT, ka.flags2, ka.tile_list, ka.tile_list_n,
                     nullptr, n);
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ka.done, 1) == ka.n_key_ctas - 1);
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    int maxlen = 0;
    for (int b = threadIdx.x; b < ka.g.B; b += 32) {
      int s0 = __ldcg(ka.seg_start + b), s1 = __ldcg(ka.seg_end + b);
      if (s0 >= 0) maxlen = max(maxlen, s1 - s0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    if (threadIdx.x == 0)
      *ka.mode = ((__ldcg(ka.hdr) & kHdrSortedBit) && maxlen <= kLeaderMaxSeg) ? 1 : 0;
    return;
  }
  // fill role: the output is walked linearly, plane by plane, 16 KiB per CTA
  const int fid = blockIdx.x - ka.n_key_ctas;
  const int plane = fid / p.zero_groups, seg = fid - plane * p.zero_groups;
  const int b = plane / (p.C + 1), c = plane - b * (p.C + 1);
  float* base = (c < p.C) ? p.matrix + ((long long)b * p.C + c) * p.V
                          : reinterpret_cast<float*>(p.counts) + (long long)b * p.V;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4* dst = reinterpret_cast<float4*>(base) + seg * 1024 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) __stcs(dst + k * 256, z4);
}

// Kernel A: per-point leader scatter (fast mode; exits otherwise).  One warp per point; the
// lowest-index point of a voxel sums the voxel's points in ascending order (lanes over
// channels), divides and writes C values + the count over the zeros of k_avg_keys_fill.
__global__ void __launch_bounds__(256)
k_avg_leaders(AvgParams p) {
  const int fast = __ldg(p.mode);
  __shared__ int skeys[kLeaderSmemKeys];
  const int lead_id = blockIdx.x;
  if (!fast) return;
  // The CTA's 8 points lie in at most a few batch segments; their keys are staged in shared
  // memory once and every warp scans them from there.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_first = (long long)lead_id * 8;
  const long long n_last = min(n_first + 7, p.N - 1);
  int k_first = -1, k_last = -1;                            // first / last valid key of the CTA
  for (long long j = n_first; j <= n_last; ++j) {
    int kk = __ldg(p.keys + j);
    if (kk >= 0) { if (k_first < 0) k_first = kk; k_last = kk; }
  }
  if (k_first < 0) return;                                  // uniform: all 8 points out of bounds
  const int s_lo = __ldg(p.seg_start + k_first / p.V), s_hi = __ldg(p.seg_end + k_last / p.V);
  const bool staged = (s_hi - s_lo) <= kLeaderSmemKeys;
  if (staged)
    for (int e = threadIdx.x; e < s_hi - s_lo; e += 256) skeys[e] = __ldg(p.keys + s_lo + e);
  __syncthreads();
  const long long n = n_first + warp;
  if (n >= p.N) return;
  const int key = staged ? skeys[n - s_lo] : __ldg(p.keys + n);
  if (key < 0) return;
  const int b = key / p.V;
  const int lo = __ldg(p.seg_start + b), hi = __ldg(p.seg_end + b);
  const int nn = (int)n;
  bool earlier = false;
  for (int j0 = lo; j0 < nn && !earlier; j0 += 32) {
    int j = j0 + lane;
    bool m = (j < nn) && ((staged ? skeys[j - s_lo] : __ldg(p.keys + j)) == key);
    earlier = __any_sync(0xffffffffu, m);
  }
  if (earlier) return;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  int count = 0;
  for (int j0 = nn - (nn - lo) % 32; j0 < hi; j0 += 32) {   // aligned chunks: j ascends
    int j = j0 + lane;
    bool m = (j >= nn) && (j < hi) && ((staged ? skeys[j - s_lo] : __ldg(p.keys + j)) == key);
    unsigned mask = __ballot_sync(0xffffffffu, m);
    while (mask) {
      int l = __ffs(mask) - 1;
      mask &= mask - 1;
      const float* src = p.values + (long long)(j0 + l) * p.C;
      for (int k = 0; k * 32 < p.C && k < 8; ++k) {
        int c = lane + 32 * k;
        if (c < p.C) acc[k] = __fadd_rn(acc[k], __ldg(src + c));
      }
      ++count;
    }
  }
  const int flat = key - b * p.V;
  const float cf = (float)count;
  float* dst = p.matrix + (long long)b * p.C * p.V + flat;
  for (int k = 0; k * 32 < p.C && k < 8; ++k) {
    int c = lane + 32 * k;
    if (c < p.C) dst[(long long)c * p.V] = __fdiv_rn(acc[k], cf);
  }
  if (lane == 0) p.counts[(long long)b * p.V + flat] = count;
}

// Kernel B: one CTA per (occupied tile, channel chunk): ordered compaction of the point keys in
// its voxel range, per-voxel sums in ascending point order in shared memory, divide, write.
__global__ void __launch_bounds__(kThreads)
k_avg_tiles(AvgParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);          // [VT][CCp]
  int* cnt = reinterpret_cast<int*>(tile + (size_t)p.VT * p.CCp);  // [VT]
  int* list_n = cnt + p.VT;                                   // [kListCap]
  int* list_v = list_n + kListCap;                            // [kListCap]
  float* stage = reinterpret_cast<float*>(list_v + kListCap); // [kStage][CCp]
  __shared__ int s_warp[kThreads / 32];
  __shared__ int s_total;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.y;
  if (p.mode && __ldg(p.mode)) return;                       // fast mode already wrote everything
  // grid-stride over the occupied tiles (sparse mode: the count is only known on the device,
  // so the grid is a fixed few CTAs per SM) or over all tiles (ragged shapes)
  const int n_work = p.tile_list ? __ldg(p.tile_list_n) : p.n_items;
  for (int item = blockIdx.x; item < n_work; item += gridDim.x) {
  const int tb = p.tile_list ? __ldg(p.tile_list + item) : item;
  const int ti = tb % p.tiles_per_batch, b = tb / p.tiles_per_batch;
  const int vbase = ti * p.VT;
  const int vt = min(p.VT, p.V - vbase);
  const int c0 = chunk * p.CC;
  const int cc = min(p.CC, p.C - c0);
  float* out = p.matrix + ((long long)b * p.C + c0) * p.V + vbase;

  for (int e = tid; e < p.VT * p.CCp; e += kThreads) tile[e] = 0.f;
  for (int e = tid; e < p.VT; e += kThreads) cnt[e] = 0;
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
        list_n[pos] = (int)(i0 + j);  // point id (N < 2^31 checked on host)
        list_v[pos] = k[j] - klo;
        ++pos;
      }
    __syncthreads();
    const int L = s_total;
    // values of the listed points are staged through shared memory in batches of kStage rows
    // (independent, coalesced loads), then added per voxel in ascending point order
    for (int l0 = 0; l0 < L; l0 += kStage) {
      const int nl = min(kStage, L - l0);
      for (int e = tid; e < nl * cc; e += kThreads) {
        int l = e / cc, ch = e - l * cc;
        stage[l * p.CCp + ch] = __ldg(p.values + (long long)list_n[l0 + l] * p.C + c0 + ch);
      }
      __syncthreads();
      if (acc_thread) {
        for (int l = 0; l < nl; ++l) {
          int v = list_v[l0 + l];
          if ((v % p.G) == g) {
            tile[v * p.CCp + c] = __fadd_rn(tile[v * p.CCp + c], stage[l * p.CCp + c]);
            if (c == 0) cnt[v] += 1;
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- write the tile: every output element exactly once, coalesced along v.
  // thread <-> voxel (no integer division in the loop; the count is read once)
  for (int v = tid; v < vt; v += kThreads) {
    const int n = cnt[v];
    const float fn = (float)n;
    const float* row = tile + v * p.CCp;
    float* o = out + v;
    if (n > 0) {
      for (int ch = 0; ch < cc; ++ch) __stcs(o + (long long)ch * p.V, __fdiv_rn(row[ch], fn));
    } else {
#pragma unroll 8
      for (int ch = 0; ch < cc; ++ch) __stcs(o + (long long)ch * p.V, 0.f);
    }
  }
  if (chunk == 0) {
    int* oc = p.counts + (long long)b * p.V + vbase;
    for (int v = tid; v < vt; v += kThreads) oc[v] = cnt[v];
  }
  __syncthreads();                                           // shared tile is reused by the next item
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

// workspace layout (T = voxel tiles of this call, W = B*V/32 bitmap words):
//   region A, memset 0xFF : hdr[4] | seg_start[B] | seg_end[B]          (reserved: 4 + 2*65536 ints)
//   region B, memset 0x00 : flags | tile_list_n | done | mode | tile_count[T] | occ_bits[W]
//                                                           (reserved: 4 + 2^20 + 2^21 ints)
//   tile_list[2^20] | keys[N]
constexpr size_t kWsA = (4 + 2 * 65536) * sizeof(int);
constexpr size_t kMaxTiles = 1 << 20;
constexpr size_t kMaxBitWords = 1 << 21;                     // B*V <= 2^26 voxels
constexpr size_t kWsListOff = kWsA + (4 + kMaxTiles + kMaxBitWords) * sizeof(int);
constexpr size_t kWsKeysOff = kWsListOff + kMaxTiles * sizeof(int);

extern "C" size_t mf_average_voxelization_3d_flags_offset(void) { return kWsA; }

extern "C" size_t mf_average_voxelization_3d_workspace_bytes(int64_t n_points) {
  return kWsKeysOff + align_up((size_t)(n_points > 0 ? n_points : 1) * 4, 256);
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

  AvgParams p;
  p.VT = 256;
  if (V < 256) p.VT = (int)V;
  p.tiles_per_batch = (int)((V + p.VT - 1) / p.VT);
  if ((long long)p.tiles_per_batch * B > (1 << 20)) return MF_E_TOOLARGE;

  int* hdr = (int*)workspace;
  int* seg_start = hdr + 4;
  int* seg_end = seg_start + B;
  int* regB = (int*)((char*)workspace + kWsA);
  int* flags2 = regB;            // internal copy of the flag word (workspace-resident)
  int* tile_list_n = regB + 1;
  int* done = regB + 2;
  int* mode = regB + 3;
  int* tile_count = regB + 4;
  const long long n_tiles_ws = (long long)p.tiles_per_batch * B;
  int* tile_list = (int*)((char*)workspace + kWsListOff);
  int* keys = (int*)((char*)workspace + kWsKeysOff);
  // hdr = all ones (sorted bit set), seg_* = -1 (empty); flags/counter/tile counts = 0
  MF_CUDA_TRY(cudaMemsetAsync(hdr, 0xFF, (4 + 2 * (size_t)B) * sizeof(int), stream));
  MF_CUDA_TRY(cudaMemsetAsync(regB, 0, (4 + (size_t)n_tiles_ws) * sizeof(int), stream));
  VoxGeom g{ox, oy, oz, pitch, X, Y, Z, B};
  p.values = values; p.keys = keys; p.hdr = hdr; p.seg_start = seg_start; p.seg_end = seg_end;
  p.N = N; p.C = C; p.B = B; p.V = (int)V;
  int nChunks = (C + 63) / 64;
  p.CC = (C + nChunks - 1) / nChunks;
  nChunks = (C + p.CC - 1) / p.CC;
  p.CCp = p.CC | 1;
  p.G = kThreads / p.CC;
  if (p.G < 1) p.G = 1;
  if (p.G > 32) p.G = 32;
  p.matrix = matrix; p.counts = counts;
  p.tile_count = tile_count; p.tile_list = tile_list; p.tile_list_n = tile_list_n;
  p.n_chunks = nChunks;
  const long long n_tiles = (long long)p.tiles_per_batch * B;
  p.n_items = (int)n_tiles;
  size_t smem = (size_t)p.VT * p.CCp * 4 + (size_t)p.VT * 4 + (size_t)kListCap * 8 +
                (size_t)kStage * p.CCp * 4;
  static bool attr_set = false;
  if (!attr_set) {
    MF_CUDA_TRY(cudaFuncSetAttribute(k_avg_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     100 * 1024));
    attr_set = true;
  }
  p.mode = mode;
  p.zero_groups = 0;
  p.n_lead_ctas = 0;
  const bool sparse_ok = (V * B <= (1LL << 26)) && (p.VT == 256) && (V % 4096 == 0) && C <= 256 &&
                         N > 0;
  if (sparse_ok) {
    // 1. prepass + dense zero fill in one launch; 2. per-point leader scatter (fast mode);
    // 3. occupied tiles through shared memory (general mode).  2 and 3 are mutually exclusive
    // (device-side mode word) and overwrite only occupied voxels / tiles.
    p.zero_groups = (int)(V / 4096);                           // 16 KiB segments per plane
    const long long n_fill = (long long)B * (C + 1) * p.zero_groups;
    const long long n_key = div_up(N, 256);
    const long long n_lead = (N * 32 + 255) / 256;
    if (n_fill + n_key >= (1LL << 31) || n_lead >= (1LL << 31)) return MF_E_TOOLARGE;
    p.n_lead_ctas = (int)n_lead;
    KeysArgs ka{points, batch_indices, g, keys, hdr, seg_start, seg_end, flags, tile_count,
                flags2, tile_list, tile_list_n, done, mode, (int)n_key};
    k_avg_keys_fill<<<(unsigned)(n_key + n_fill), 256, 0, stream>>>(p, ka);
    MF_LAUNCH_CHECK();
    k_avg_leaders<<<(unsigned)n_lead, 256, 0, stream>>>(p);
    MF_LAUNCH_CHECK();
    long long occ_max = n_tiles < N ? n_tiles : N;             // at most one new tile per point
    if (occ_max > 592) occ_max = 592;                          // 4 x 148: grid-stride inside
    if (occ_max > 0) {
      dim3 grid((unsigned)occ_max, nChunks, 1);
      k_avg_tiles<<<grid, kThreads, smem, stream>>>(p);
      MF_LAUNCH_CHECK();
    }
  } else {
    if (N > 0) {
      k_vox_keys<<<div_up(N, 256), 256, 0, stream>>>(points, batch_indices, N, g, keys, hdr,
                                                     seg_start, seg_end, flags, tile_count, p.VT,
                                                     flags2, tile_list, tile_list_n, nullptr, done,
                                                     mode);
      MF_LAUNCH_CHECK();
    }
    // ragged shapes: every tile through the shared-memory path
    p.tile_list = nullptr;
    p.mode = nullptr;
    dim3 grid((unsigned)n_tiles, nChunks, 1);
    k_avg_tiles<<<grid, kThreads, smem, stream>>>(p);
    MF_LAUNCH_CHECK();
  }
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
