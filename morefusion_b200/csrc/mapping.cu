// Occupancy-grid producer: the map behind
//   morefusion/contrib/multi_instance_octree_mapping.py:7-94   (OcTree per instance, get_target_grids)
// as a device-resident hashed log-odds voxel map.  The reference delegates to OctoMap (octomap-python,
// absent from the reference tree); the semantics kept here are OctoMap 1.9's, as restated and cited
// in oracle/octomap.py: 16-bit keys per axis, float32 log-odds with clamping, the Amanatides & Woo
// key ray of OcTreeBaseImpl::computeRayKeys in the library's float/double mix, "every cell updated
// at most once per scan, end points win over free cells" (OccupancyOcTreeBase::computeUpdate).
//
// B200 design.  A pruned octree is, for search(depth=0), a flat key -> log-odds map: ONE open-
// addressing hash table in HBM holds every instance's cells (key = instance:16 | kx:16 | ky:16 |
// kz:16), so a scan is two launches and the three 32^3 grids of all targets are one launch:
//   k_map_scan_rays   one thread per pixel: end-point ("hit") update + the DDA, run twice - once
//                     to count the ray's cells, once to write their keys into a slice of the key
//                     buffer obtained with one warp-aggregated atomicAdd.  The DDA is sequential
//                     double arithmetic per ray (tMax accumulates rounding exactly as the library's
//                     does), so the memory work is split off:
//   k_map_free_apply  one thread per (ray, cell) key: probe / insert, test-before-atomic on the
//                     cell's scan stamp; the first toucher of a cell in this scan applies the miss
//                     update.  Millions of independent probes keep the L2 / HBM pipes full, which a
//                     thread-per-ray walk with dependent probes cannot.
//   k_map_query_grids one thread per (target, voxel): looks the voxel centre up in every instance,
//                     in insertion order, and writes grid_target / grid_nontarget / grid_empty with
//                     the reference's overwrite rules (multi_instance_octree_mapping.py:85-92).
// A cell is updated by exactly one thread per scan, so results do not depend on thread order:
// bit-identical to the serial oracle.  -fmad=false: the double expressions must not contract.
#include <math.h>

#ifdef MF_HOST_EMU
// tests/emu/mapping_emu.cpp runs the kernels of this file serially on the host (one-lane "warps")
// against oracle/octomap.py in the CPU test suite; it supplies the CUDA spellings used below.
#include "mapping_emu_stubs.h"
#else
#include "common.cuh"
#endif

namespace mf {

typedef unsigned long long u64;
static constexpr u64 MAP_EMPTY = ~0ull;
static constexpr int TREE_MAX_VAL = 32768;
#ifdef MF_HOST_EMU
static constexpr int WARP_LAST = 0;
#else
static constexpr int WARP_LAST = 31;
#endif

struct MapTable {
  u64* keys;        // [cap]   MAP_EMPTY = free slot
  float* lo;        // [cap]   log-odds, 0 for a slot that has never been updated
  unsigned* stamp;  // [cap]   2*scan (missed in scan) / 2*scan+1 (hit in scan), monotonic
  unsigned* cnt;    // [cap]   pending updateNodes() multiplicity (0 between calls)
  unsigned mask;    // cap - 1 (cap is a power of two)
  int* counters;    // [0] live entries  [1] table overflow (sticky)
                    // [2 + 2 (scan & 1)] ray keys of the scan, [3 + 2 (scan & 1)] key buffer too small:
                    // the scan's second launch re-zeroes the other parity's pair, so there is no
                    // memset node and the host can still read a scan's pair until the next scan ends
};

__device__ __forceinline__ unsigned map_hash(u64 k) {
  k *= 0x9E3779B97F4A7C15ull;
  k ^= k >> 29;
  k *= 0xBF58476D1CE4E5B9ull;
  return (unsigned)(k >> 32);
}

// slot of `key`, inserting it when absent; -1 when the table is full (overflow flag raised)
__device__ __forceinline__ int map_find_or_insert(const MapTable& t, u64 key) {
  unsigned s = map_hash(key) & t.mask;
  for (unsigned n = 0; n <= t.mask; ++n) {
    u64 k = t.keys[s];
    if (k == key) return (int)s;
    if (k == MAP_EMPTY) {
      const u64 old = atomicCAS(t.keys + s, MAP_EMPTY, key);
      if (old == MAP_EMPTY) { atomicAdd(t.counters, 1); return (int)s; }
      if (old == key) return (int)s;
    }
    s = (s + 1) & t.mask;
  }
  t.counters[1] = 1;
  return -1;
}

__device__ __forceinline__ int map_find(const MapTable& t, u64 key) {
  unsigned s = map_hash(key) & t.mask;
  for (unsigned n = 0; n <= t.mask; ++n) {
    const u64 k = __ldg(t.keys + s);
    if (k == key) return (int)s;
    if (k == MAP_EMPTY) return -1;
    s = (s + 1) & t.mask;
  }
  return -1;
}

__device__ __forceinline__ float map_clamped_add(float v, float upd, float lo_min, float lo_max) {
  v = __fadd_rn(v, upd);                      // OccupancyOcTreeBase::updateNodeLogOdds
  if (v < lo_min) v = lo_min;
  if (v > lo_max) v = lo_max;
  return v;
}

// OcTreeBaseImpl::coordToKeyChecked for one axis: floor(coord / resolution) + 32768, -1 if outside
__device__ __forceinline__ int coord_to_key(double c, double res_factor) {
  const double f = floor(res_factor * c) + (double)TREE_MAX_VAL;     // exact: |floor| < 2^52
  return (f >= 0.0 && f < (double)(2 * TREE_MAX_VAL)) ? (int)f : -1;  // NaN / inf fail both tests
}

__device__ __forceinline__ u64 pack_key(int inst, int kx, int ky, int kz) {
  return ((u64)(unsigned)inst << 48) | ((u64)(unsigned)kx << 32) | ((u64)(unsigned)ky << 16) |
         (u64)(unsigned)kz;
}

// The key ray of computeRayKeys (oracle/octomap.py::OcTree.computeRayKeys has the line-by-line
// commentary): first cell included, end cell excluded.  Returns the number of cells; when `out`
// is non-null also writes their packed keys.
__device__ int ray_keys(const float o[3], const float e[3], const int ko[3], const int ke[3],
                        double res, int inst, u64* out) {
  if (ko[0] == ke[0] && ko[1] == ke[1] && ko[2] == ke[2]) return 0;
  int n = 0;
  if (out) out[n] = pack_key(inst, ko[0], ko[1], ko[2]);
  ++n;
  float d[3] = {__fsub_rn(e[0], o[0]), __fsub_rn(e[1], o[1]), __fsub_rn(e[2], o[2])};
  const float nsq = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])),
                              __fmul_rn(d[2], d[2]));
  const float length = (float)sqrt((double)nsq);
  int step[3], cur[3];
  double tmax[3], tdelta[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d[i] = __fdiv_rn(d[i], length);
    cur[i] = ko[i];
    step[i] = d[i] > 0.f ? 1 : (d[i] < 0.f ? -1 : 0);
    if (step[i] != 0) {
      double border = ((double)(cur[i] - TREE_MAX_VAL) + 0.5) * res;
      border += (double)(float)((double)step[i] * res * 0.5);
      tmax[i] = (border - (double)o[i]) / (double)d[i];
      tdelta[i] = res / fabs((double)d[i]);
    } else {
      tmax[i] = 1.7976931348623157e308;
      tdelta[i] = 1.7976931348623157e308;
    }
  }
  const double flen = (double)length;
  for (;;) {
    int dim;
    if (tmax[0] < tmax[1]) dim = (tmax[0] < tmax[2]) ? 0 : 2;
    else dim = (tmax[1] < tmax[2]) ? 1 : 2;
    // no dynamic indexing of the register arrays
    if (dim == 0) { cur[0] += step[0]; tmax[0] += tdelta[0]; }
    else if (dim == 1) { cur[1] += step[1]; tmax[1] += tdelta[1]; }
    else { cur[2] += step[2]; tmax[2] += tdelta[2]; }
    if (cur[0] == ke[0] && cur[1] == ke[1] && cur[2] == ke[2]) break;
    if (fmin(fmin(tmax[0], tmax[1]), tmax[2]) > flen) break;
    const int c = dim == 0 ? cur[0] : (dim == 1 ? cur[1] : cur[2]);
    if (c < 0 || c >= 2 * TREE_MAX_VAL) break;
    if (out) out[n] = pack_key(inst, cur[0], cur[1], cur[2]);
    ++n;
  }
  return n;
}

struct ScanParams {
  const float* points;        // [n,3]
  const unsigned char* mask;  // [n] or null
  int n;
  float ox, oy, oz;           // sensor origin (point3d: floats)
  double res, res_factor;
  int inst;
  unsigned scan;
  float hit, miss, lo_min, lo_max;
  u64* ray_buf;
  int ray_cap;
};

__global__ void __launch_bounds__(256)
k_map_scan_rays(ScanParams p, MapTable t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool valid = i < p.n;
  float e[3] = {0.f, 0.f, 0.f};
  if (valid) {
    e[0] = p.points[3 * (size_t)i];
    e[1] = p.points[3 * (size_t)i + 1];
    e[2] = p.points[3 * (size_t)i + 2];
    valid = !(isnan(e[0]) || isnan(e[1]) || isnan(e[2])) && (!p.mask || p.mask[i]);
  }
  const float o[3] = {p.ox, p.oy, p.oz};
  int ko[3], ke[3];
  bool ok_o = true, ok_e = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    ko[a] = coord_to_key((double)o[a], p.res_factor);
    ke[a] = coord_to_key((double)e[a], p.res_factor);
    ok_o &= ko[a] >= 0;
    ok_e &= ke[a] >= 0;
  }
  // end point: hit update by the first thread that reaches the cell in this scan
  if (valid && ok_e) {
    const int s = map_find_or_insert(t, pack_key(p.inst, ke[0], ke[1], ke[2]));
    if (s >= 0) {
      const unsigned want = 2u * p.scan + 1u;
      if (atomicMax(t.stamp + s, want) < want)
        t.lo[s] = map_clamped_add(t.lo[s], p.hit, p.lo_min, p.lo_max);
    }
  }
  // free cells: count, take a slice of the key buffer (one atomic per warp), write the keys
  const bool ray = valid && ok_o && ok_e;
  const int n = ray ? ray_keys(o, e, ko, ke, p.res, p.inst, nullptr) : 0;
  int incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, WARP_LAST);
  int base = 0;
  int* ctr = t.counters + 2 + 2 * (int)(p.scan & 1u);
  if (lane == WARP_LAST && total > 0) base = atomicAdd(ctr, total);
  base = __shfl_sync(0xffffffffu, base, WARP_LAST);
  if (n > 0) {
    const long long off = (long long)base + incl - n;
    if (off + n <= (long long)p.ray_cap) ray_keys(o, e, ko, ke, p.res, p.inst, p.ray_buf + off);
    else ctr[1] = 1;
  }
}

__global__ void __launch_bounds__(256)
k_map_free_apply(const u64* __restrict__ ray_buf, int ray_cap, unsigned scan, float miss,
                 float lo_min, float lo_max, MapTable t) {
  const int* ctr = t.counters + 2 + 2 * (int)(scan & 1u);
  if (blockIdx.x == 0 && threadIdx.x == 0) {    // next scan's pair (nobody reads it during this scan)
    t.counters[2 + 2 * (int)(~scan & 1u)] = 0;
    t.counters[3 + 2 * (int)(~scan & 1u)] = 0;
  }
  if (ctr[1]) return;                           // scan did not fit the key buffer: leave it unapplied
  const int total = min(ctr[0], ray_cap);
  const unsigned want = 2u * scan;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int s = map_find_or_insert(t, ray_buf[i]);
    if (s < 0) continue;
    if (*(volatile unsigned*)(t.stamp + s) >= want) continue;      // already missed or hit in this scan
    if (atomicMax(t.stamp + s, want) < want)
      t.lo[s] = map_clamped_add(t.lo[s], miss, lo_min, lo_max);
  }
}

// octomap-python updateNodes(points, occupied): one updateNode per row (double coordinates, no
// de-duplication): count the rows per cell, then the thread that collects a cell's count applies
// the update that many times (the update sequence of a cell is the same value k times, so the
// result depends on k only).
__global__ void __launch_bounds__(256)
k_map_update_count(const double* __restrict__ pts, int m, double res_factor, int inst, MapTable t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int kx = coord_to_key(pts[3 * (size_t)i], res_factor);
  const int ky = coord_to_key(pts[3 * (size_t)i + 1], res_factor);
  const int kz = coord_to_key(pts[3 * (size_t)i + 2], res_factor);
  if (kx < 0 || ky < 0 || kz < 0) return;       // NaN lands here too: floor(NaN) -> INT_MIN
  const int s = map_find_or_insert(t, pack_key(inst, kx, ky, kz));
  if (s >= 0) atomicAdd(t.cnt + s, 1u);
}

__global__ void __launch_bounds__(256)
k_map_update_apply(const double* __restrict__ pts, int m, double res_factor, int inst, float upd,
                   float lo_min, float lo_max, MapTable t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int kx = coord_to_key(pts[3 * (size_t)i], res_factor);
  const int ky = coord_to_key(pts[3 * (size_t)i + 1], res_factor);
  const int kz = coord_to_key(pts[3 * (size_t)i + 2], res_factor);
  if (kx < 0 || ky < 0 || kz < 0) return;
  const int s = map_find(t, pack_key(inst, kx, ky, kz));
  if (s < 0) return;
  unsigned k = atomicExch(t.cnt + s, 0u);
  if (!k) return;
  float v = t.lo[s];
  for (; k; --k) v = map_clamped_add(v, upd, lo_min, lo_max);
  t.lo[s] = v;
}

// get_target_grids for T targets at once (multi_instance_octree_mapping.py:35-94)
__global__ void __launch_bounds__(256)
k_map_query_grids(const int* __restrict__ target_index, const double* __restrict__ pitch,
                  const double* __restrict__ origin, int T, int X, int Y, int Z,
                  const double* __restrict__ res_factor, int n_inst, MapTable t,
                  float* __restrict__ g_target, float* __restrict__ g_nontarget,
                  float* __restrict__ g_empty) {
  const long long V = (long long)X * Y * Z;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= V * T) return;
  const int tg = (int)(gid / V);
  const int v = (int)(gid - (long long)tg * V);
  const int k = v % Z, j = (v / Z) % Y, i = v / (Z * Y);
  const double pt = pitch[tg];
  // trimesh matrix_to_points: indices * pitch + origin (float64)
  const double cx = (double)i * pt + origin[3 * tg];
  const double cy = (double)j * pt + origin[3 * tg + 1];
  const double cz = (double)k * pt + origin[3 * tg + 2];
  const int tin = target_index[tg];
  float gt = 0.f, gn = 0.f, ge = 0.f;
  for (int q = 0; q < n_inst; ++q) {
    const double rf = res_factor[q];
    const int kx = coord_to_key(cx, rf), ky = coord_to_key(cy, rf), kz = coord_to_key(cz, rf);
    if (kx < 0 || ky < 0 || kz < 0) continue;
    const int s = map_find(t, pack_key(q, kx, ky, kz));
    if (s < 0) continue;
    const double occ = 1.0 - (1.0 / (1.0 + exp((double)t.lo[s])));   // octomap::probability
    if (occ >= 0.5) {
      if (q == tin) gt = (float)occ; else gn = (float)occ;
    } else if (occ >= 0.0) {
      ge = (float)(1.0 - occ);
    }
  }
  g_target[gid] = gt;
  g_nontarget[gid] = gn;
  g_empty[gid] = ge;
}

__global__ void __launch_bounds__(256)
k_map_rehash(const u64* __restrict__ keys, const float* __restrict__ lo,
             const unsigned* __restrict__ stamp, unsigned old_cap, MapTable t) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < old_cap; i += gridDim.x * blockDim.x) {
    const u64 k = keys[i];
    if (k == MAP_EMPTY) continue;
    const int s = map_find_or_insert(t, k);
    if (s >= 0) { t.lo[s] = lo[i]; t.stamp[s] = stamp[i]; }
  }
}

}  // namespace mf

#ifndef MF_HOST_EMU
using namespace mf;

static bool table_ok(const void* keys, const void* lo, const void* stamp, const void* cnt,
                     long long cap, const void* counters) {
  return keys && lo && stamp && cnt && counters && cap >= 64 && cap <= (1LL << 31) &&
         (cap & (cap - 1)) == 0;
}

static MapTable make_table(void* keys, void* lo, void* stamp, void* cnt, long long cap, void* counters) {
  MapTable t;
  t.keys = (u64*)keys; t.lo = (float*)lo; t.stamp = (unsigned*)stamp; t.cnt = (unsigned*)cnt;
  t.mask = (unsigned)(cap - 1); t.counters = (int*)counters;
  return t;
}

extern "C" int mf_map_integrate(const float* points, const uint8_t* mask, int64_t n, float ox,
                                float oy, float oz, double resolution, int instance,
                                uint32_t scan, float hit, float miss, float lo_min, float lo_max,
                                void* keys, void* lo, void* stamp, void* cnt, int64_t capacity,
                                int32_t* counters, void* ray_buf, int64_t ray_capacity,
                                void* stream_) {
  if (!points || n <= 0 || n >= (1LL << 31) || !(resolution > 0.0) || instance < 0 ||
      instance >= 0xFFFF || scan == 0 || scan >= (1u << 31) || !ray_buf || ray_capacity <= 0 ||
      ray_capacity >= (1LL << 31) || !table_ok(keys, lo, stamp, cnt, capacity, counters))
    return MF_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  MapTable t = make_table(keys, lo, stamp, cnt, capacity, counters);
  ScanParams p;
  p.points = points; p.mask = mask; p.n = (int)n; p.ox = ox; p.oy = oy; p.oz = oz;
  p.res = resolution; p.res_factor = 1.0 / resolution; p.inst = instance; p.scan = scan;
  p.hit = hit; p.miss = miss; p.lo_min = lo_min; p.lo_max = lo_max;
  p.ray_buf = (u64*)ray_buf; p.ray_cap = (int)ray_capacity;
  k_map_scan_rays<<<div_up(n, 256), 256, 0, st>>>(p, t);
  MF_LAUNCH_CHECK();
  int dev = 0, n_sm = 0;
  MF_CUDA_TRY(cudaGetDevice(&dev));
  MF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  k_map_free_apply<<<n_sm * 8, 256, 0, st>>>((const u64*)ray_buf, (int)ray_capacity, scan, miss,
                                              lo_min, lo_max, t);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_map_update_points(const double* points, int64_t m, double resolution,
                                    int instance, float update, float lo_min, float lo_max,
                                    void* keys, void* lo, void* stamp, void* cnt,
                                    int64_t capacity, int32_t* counters, void* stream_) {
  if (!points || m <= 0 || m >= (1LL << 31) || !(resolution > 0.0) || instance < 0 ||
      instance >= 0xFFFF || !table_ok(keys, lo, stamp, cnt, capacity, counters))
    return MF_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  MapTable t = make_table(keys, lo, stamp, cnt, capacity, counters);
  k_map_update_count<<<div_up(m, 256), 256, 0, st>>>(points, (int)m, 1.0 / resolution, instance, t);
  MF_LAUNCH_CHECK();
  k_map_update_apply<<<div_up(m, 256), 256, 0, st>>>(points, (int)m, 1.0 / resolution, instance,
                                                     update, lo_min, lo_max, t);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_map_query_grids(const int32_t* target_index, const double* pitch,
                                  const double* origin, int T, int X, int Y, int Z,
                                  const double* res_factor, int n_instances, void* keys, void* lo,
                                  void* stamp, void* cnt, int64_t capacity, int32_t* counters,
                                  float* grid_target, float* grid_nontarget, float* grid_empty,
                                  void* stream_) {
  if (!target_index || !pitch || !origin || T <= 0 || X <= 0 || Y <= 0 || Z <= 0 || !res_factor ||
      n_instances < 0 || !grid_target || !grid_nontarget || !grid_empty ||
      (long long)T * X * Y * Z >= (1LL << 40) || !table_ok(keys, lo, stamp, cnt, capacity, counters))
    return MF_E_BADARG;
  MapTable t = make_table(keys, lo, stamp, cnt, capacity, counters);
  const long long total = (long long)T * X * Y * Z;
  k_map_query_grids<<<div_up(total, 256), 256, 0, (cudaStream_t)stream_>>>(
      target_index, pitch, origin, T, X, Y, Z, res_factor, n_instances, t, grid_target,
      grid_nontarget, grid_empty);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_map_rehash(const void* old_keys, const void* old_lo, const void* old_stamp,
                             int64_t old_capacity, void* keys, void* lo, void* stamp, void* cnt,
                             int64_t capacity, int32_t* counters, void* stream_) {
  if (!old_keys || !old_lo || !old_stamp || old_capacity <= 0 || old_capacity > (1LL << 31) ||
      capacity < old_capacity || !table_ok(keys, lo, stamp, cnt, capacity, counters))
    return MF_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  MapTable t = make_table(keys, lo, stamp, cnt, capacity, counters);
  MF_CUDA_TRY(cudaMemsetAsync(counters, 0, 2 * sizeof(int32_t), st));   // entries recounted by the inserts
  int dev = 0, n_sm = 0;
  MF_CUDA_TRY(cudaGetDevice(&dev));
  MF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  k_map_rehash<<<n_sm * 8, 256, 0, st>>>((const u64*)old_keys, (const float*)old_lo,
                                          (const unsigned*)old_stamp, (unsigned)old_capacity, t);
  MF_LAUNCH_CHECK();
  return MF_OK;
}
#endif  // MF_HOST_EMU
