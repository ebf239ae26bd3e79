// Occupancy-grid producer: the map behind
//   morefusion/contrib/multi_instance_octree_mapping.py:7-94   (OcTree per instance, get_target_grids)
// as a device-resident hashed log-odds voxel map.  The reference delegates to OctoMap (octomap-python,
// absent from the reference tree); the semantics kept here are OctoMap 1.9's, as restated and cited
// in oracle/octomap.py: 16-bit keys per axis, float32 log-odds with clamping, the Amanatides & Woo
// key ray of OcTreeBaseImpl::computeRayKeys in the library's float/double mix, "every cell updated
// at most once per scan, end points win over free cells" (OccupancyOcTreeBase::computeUpdate).
//
// B200 design.  A pruned octree is, for search(depth=0), a flat key -> log-odds map: ONE open-
// addressing hash table in HBM holds every instance's cells as 16-byte entries
// {key = instance:16 | kx:16 | ky:16 | kz:16, float log-odds, u32 scan stamp}: a probe is one
// 16-byte L2 access that brings the key, the value and the stamp together.  A scan is two launches,
// the three 32^3 grids of all targets are one launch:
//   k_map_scan_hits   one thread per pixel: end-point ("hit") update.
//   k_map_scan_free   one thread per pixel walks its ray (the DDA is sequential double arithmetic:
//                     tMax accumulates rounding exactly as the library's does).  The 32 lanes of a
//                     warp are 32 neighbouring pixels; at a 1 cm cell and <1.7 mm between rays they
//                     sit in the same one to five cells at every step, so the lanes of a step are
//                     grouped by key (__match_any_sync) and only one lane per distinct cell probes
//                     the table; a cell already stamped by this scan costs that one read
//                     (test-before-atomic), the first toucher applies the miss update.  No key
//                     buffer: the first version wrote every (ray, cell) key to HBM and applied them
//                     in a second kernel (profiles/r02_mapping_history.md: 39 M keys, 310 MB written
//                     and read back, 130 visits per distinct cell).
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

struct __align__(16) MapCell {
  u64 key;          // MAP_EMPTY = free slot
  float lo;         // log-odds, 0 for a slot that has never been updated
  unsigned stamp;   // 2*scan (missed in scan) / 2*scan+1 (hit in scan), monotonic
};

struct MapTable {
  MapCell* cells;   // [cap]
  unsigned* cnt;    // [cap]   pending updateNodes() multiplicity (0 between calls)
  unsigned mask;    // cap - 1 (cap is a power of two)
  int* counters;    // [0] live entries  [1] table full (sticky; updates were dropped)
};

__device__ __forceinline__ unsigned map_hash(u64 k) {
  k *= 0x9E3779B97F4A7C15ull;
  k ^= k >> 29;
  k *= 0xBF58476D1CE4E5B9ull;
  return (unsigned)(k >> 32);
}

// one 16-byte L2 access (other SMs update cells during a scan: never through the incoherent L1)
__device__ __forceinline__ MapCell map_load(const MapCell* c) {
  const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(c));
  MapCell r;
  r.key = v.x;
  r.lo = __uint_as_float((unsigned)(v.y & 0xffffffffull));
  r.stamp = (unsigned)(v.y >> 32);
  return r;
}

// slot of `key`, inserting it when absent; -1 when the table is full (overflow flag raised).
// *seen = the cell as loaded (stamp 0 for a cell this call inserted or lost the race for: a stale
// stamp only costs a redundant atomicMax, never an update).
// Probe sequences are bounded (MAP_MAX_PROBES): a table that is (nearly) full reports overflow
// instead of walking millions of slots per thread.
static constexpr unsigned MAP_MAX_PROBES = 4096;
__device__ __forceinline__ int map_find_or_insert(const MapTable& t, u64 key, MapCell* seen) {
  unsigned s = map_hash(key) & t.mask;
  const unsigned limit = t.mask < MAP_MAX_PROBES ? t.mask : MAP_MAX_PROBES;
  for (unsigned n = 0; n <= limit; ++n) {
    const MapCell c = map_load(t.cells + s);
    if (c.key == key) { *seen = c; return (int)s; }
    if (c.key == MAP_EMPTY) {
      const u64 old = atomicCAS(&t.cells[s].key, MAP_EMPTY, key);
      if (old == MAP_EMPTY || old == key) {
        if (old == MAP_EMPTY) atomicAdd(t.counters, 1);
        seen->key = key; seen->lo = 0.f; seen->stamp = 0u;
        return (int)s;
      }
    }
    s = (s + 1) & t.mask;
  }
  t.counters[1] = 1;
  return -1;
}

__device__ __forceinline__ int map_find(const MapTable& t, u64 key, MapCell* seen) {
  unsigned s = map_hash(key) & t.mask;
  const unsigned limit = t.mask < MAP_MAX_PROBES ? t.mask : MAP_MAX_PROBES;
  for (unsigned n = 0; n <= limit; ++n) {
    const MapCell c = map_load(t.cells + s);
    if (c.key == key) { *seen = c; return (int)s; }
    if (c.key == MAP_EMPTY) return -1;
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

// the once-per-scan update of a cell: the thread whose atomicMax raises the stamp to `want` owns it
__device__ __forceinline__ void map_scan_update(const MapTable& t, int s, const MapCell& seen,
                                                unsigned want, float upd, float lo_min, float lo_max) {
  if (seen.stamp >= want) return;                                  // already done in this scan
  if (atomicMax(&t.cells[s].stamp, want) < want) {
    // nobody else writes lo in this launch; the value is from an earlier launch (or 0)
    const float v = __ldcg(&t.cells[s].lo);
    t.cells[s].lo = map_clamped_add(v, upd, lo_min, lo_max);
  }
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

struct ScanParams {
  const float* points;        // [n,3]
  const unsigned char* mask;  // [n] or null          (single-instance scan)
  const int* labels;          // [n] or null          (labelled frame: every instance in one scan)
  const int* lut;             // [lut_n] label - lut_lo -> dense instance index, -1 = not mapped
  const double* inst_res;     // [n_inst] resolution of each instance (labelled frame)
  int lut_lo, lut_n;
  int n;
  float ox, oy, oz;           // sensor origin (point3d: floats)
  double res;                 // single-instance scan
  int inst;
  unsigned scan;
  float hit, miss, lo_min, lo_max;
};

// a pixel's end point, its instance / resolution and the keys of both ends; false for masked,
// unmapped and NaN pixels
__device__ __forceinline__ bool scan_pixel(const ScanParams& p, int i, float e[3], int ko[3],
                                           int ke[3], bool* ok_o, bool* ok_e, int* inst,
                                           double* res) {
  bool valid = i < p.n;
  e[0] = e[1] = e[2] = 0.f;
  *inst = p.inst;
  *res = p.res;
  if (valid) {
    e[0] = p.points[3 * (size_t)i];
    e[1] = p.points[3 * (size_t)i + 1];
    e[2] = p.points[3 * (size_t)i + 2];
    valid = !(isnan(e[0]) || isnan(e[1]) || isnan(e[2])) && (!p.mask || p.mask[i]);
    if (valid && p.labels) {
      const int l = p.labels[i] - p.lut_lo;
      const int q = (l >= 0 && l < p.lut_n) ? p.lut[l] : -1;
      valid = q >= 0;
      if (valid) { *inst = q; *res = p.inst_res[q]; }
    }
  }
  const double res_factor = 1.0 / *res;
  const float o[3] = {p.ox, p.oy, p.oz};
  *ok_o = *ok_e = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    ko[a] = coord_to_key((double)o[a], res_factor);
    ke[a] = coord_to_key((double)e[a], res_factor);
    *ok_o &= ko[a] >= 0;
    *ok_e &= ke[a] >= 0;
  }
  return valid;
}

__global__ void __launch_bounds__(256)
k_map_scan_hits(ScanParams p, MapTable t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float e[3];
  int ko[3], ke[3], inst;
  double res;
  bool ok_o, ok_e;
  if (!scan_pixel(p, i, e, ko, ke, &ok_o, &ok_e, &inst, &res) || !ok_e) return;
  MapCell seen;
  const int s = map_find_or_insert(t, pack_key(inst, ke[0], ke[1], ke[2]), &seen);
  if (s >= 0) map_scan_update(t, s, seen, 2u * p.scan + 1u, p.hit, p.lo_min, p.lo_max);
}

// The key ray of computeRayKeys (oracle/octomap.py::OcTree.computeRayKeys has the line-by-line
// commentary): first cell included, end cell excluded; every cell of the ray that this scan has not
// hit gets the miss update.  Runs after k_map_scan_hits of the same scan (stream order).
__global__ void __launch_bounds__(256, 5)
k_map_scan_free(ScanParams p, MapTable t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  float e[3];
  int ko[3], ke[3], inst;
  double res;
  bool ok_o, ok_e;
  const bool valid = scan_pixel(p, i, e, ko, ke, &ok_o, &ok_e, &inst, &res);
  bool act = valid && ok_o && ok_e && !(ko[0] == ke[0] && ko[1] == ke[1] && ko[2] == ke[2]);
  const float o[3] = {p.ox, p.oy, p.oz};
  int step0 = 0, step1 = 0, step2 = 0, cur0 = ko[0], cur1 = ko[1], cur2 = ko[2];
  double tmax0 = 0.0, tmax1 = 0.0, tmax2 = 0.0, td0 = 0.0, td1 = 0.0, td2 = 0.0, flen = 0.0;
  if (act) {
    float d[3] = {__fsub_rn(e[0], o[0]), __fsub_rn(e[1], o[1]), __fsub_rn(e[2], o[2])};
    const float nsq = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])),
                                __fmul_rn(d[2], d[2]));
    const float length = (float)sqrt((double)nsq);
    flen = (double)length;
    int step[3];
    double tmax[3], tdelta[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      d[a] = __fdiv_rn(d[a], length);
      step[a] = d[a] > 0.f ? 1 : (d[a] < 0.f ? -1 : 0);
      if (step[a] != 0) {
        double border = ((double)(ko[a] - TREE_MAX_VAL) + 0.5) * res;
        border += (double)(float)((double)step[a] * res * 0.5);
        tmax[a] = (border - (double)o[a]) / (double)d[a];
        tdelta[a] = res / fabs((double)d[a]);
      } else {
        tmax[a] = 1.7976931348623157e308;
        tdelta[a] = 1.7976931348623157e308;
      }
    }
    step0 = step[0]; step1 = step[1]; step2 = step[2];
    tmax0 = tmax[0]; tmax1 = tmax[1]; tmax2 = tmax[2];
    td0 = tdelta[0]; td1 = tdelta[1]; td2 = tdelta[2];
  }
  const unsigned want = 2u * p.scan;
  u64 key = pack_key(inst, cur0, cur1, cur2);
  for (;;) {
    const unsigned bal = __ballot_sync(0xffffffffu, act);
    if (!bal) break;
    if (act) {
      // lanes of this step that stand in the same cell: the lowest one does the table work.  Its
      // 16-byte probe is issued here and looked at after the DDA advance below (ALU work that
      // does not depend on it), which hides most of the L2 latency.
      const unsigned peers = __match_any_sync(bal, key);
      const bool lead = (__ffs(peers) - 1) == lane;
      const unsigned s0 = map_hash(key) & t.mask;
      MapCell c;
      c.key = 0; c.lo = 0.f; c.stamp = 0u;
      if (lead) c = map_load(t.cells + s0);
      const u64 kcur = key;
      // advance to the next cell.  The library's selection
      //   if (tMax0 < tMax1) dim = tMax0 < tMax2 ? 0 : 2;  else dim = tMax1 < tMax2 ? 1 : 2;
      // as predicates (no divergence); "min(tMax) > length" <=> every tMax > length
      const bool lt01 = tmax0 < tmax1, lt02 = tmax0 < tmax2, lt12 = tmax1 < tmax2;
      const bool a0 = lt01 && lt02, a1 = !lt01 && lt12;
      const bool a2 = !(a0 || a1);
      const double n0 = tmax0 + td0, n1 = tmax1 + td1, n2 = tmax2 + td2;
      tmax0 = a0 ? n0 : tmax0;
      tmax1 = a1 ? n1 : tmax1;
      tmax2 = a2 ? n2 : tmax2;
      cur0 += a0 ? step0 : 0;
      cur1 += a1 ? step1 : 0;
      cur2 += a2 ? step2 : 0;
      const int cc = a0 ? cur0 : (a1 ? cur1 : cur2);
      const bool at_end = cur0 == ke[0] && cur1 == ke[1] && cur2 == ke[2];
      const bool over = tmax0 > flen && tmax1 > flen && tmax2 > flen;
      act = !(at_end || over || cc < 0 || cc >= 2 * TREE_MAX_VAL);
      key = pack_key(inst, cur0, cur1, cur2);
      // the probe: the common case is "first slot holds the key and is already stamped"
      if (lead) {
        if (c.key == kcur) {
          map_scan_update(t, (int)s0, c, want, p.miss, p.lo_min, p.lo_max);
        } else {
          MapCell seen;
          const int s = map_find_or_insert(t, kcur, &seen);
          if (s >= 0) map_scan_update(t, s, seen, want, p.miss, p.lo_min, p.lo_max);
        }
      }
    }
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
  if (kx < 0 || ky < 0 || kz < 0) return;       // NaN rows land here too
  MapCell seen;
  const int s = map_find_or_insert(t, pack_key(inst, kx, ky, kz), &seen);
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
  MapCell seen;
  const int s = map_find(t, pack_key(inst, kx, ky, kz), &seen);
  if (s < 0) return;
  unsigned k = atomicExch(t.cnt + s, 0u);
  if (!k) return;
  float v = seen.lo;                            // written by an earlier launch only
  for (; k; --k) v = map_clamped_add(v, upd, lo_min, lo_max);
  t.cells[s].lo = v;
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
    MapCell c;
    if (map_find(t, pack_key(q, kx, ky, kz), &c) < 0) continue;
    const double occ = 1.0 - (1.0 / (1.0 + exp((double)c.lo)));      // octomap::probability
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
k_map_rehash(const MapCell* __restrict__ old, unsigned old_cap, MapTable t) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < old_cap; i += gridDim.x * blockDim.x) {
    const MapCell c = map_load(old + i);
    if (c.key == MAP_EMPTY) continue;
    MapCell seen;
    const int s = map_find_or_insert(t, c.key, &seen);
    if (s >= 0) { t.cells[s].lo = c.lo; t.cells[s].stamp = c.stamp; }
  }
}

}  // namespace mf

#ifndef MF_HOST_EMU
using namespace mf;

static bool table_ok(const void* cells, const void* cnt, long long cap, const void* counters) {
  return cells && cnt && counters && cap >= 64 && cap <= (1LL << 31) && (cap & (cap - 1)) == 0 &&
         ((uintptr_t)cells & 15) == 0;
}

static MapTable make_table(void* cells, void* cnt, long long cap, void* counters) {
  MapTable t;
  t.cells = (MapCell*)cells; t.cnt = (unsigned*)cnt; t.mask = (unsigned)(cap - 1);
  t.counters = (int*)counters;
  return t;
}

extern "C" int mf_map_integrate(const float* points, const uint8_t* mask, int64_t n, float ox,
                                float oy, float oz, double resolution, int instance,
                                uint32_t scan, float hit, float miss, float lo_min, float lo_max,
                                void* cells, void* cnt, int64_t capacity, int32_t* counters,
                                void* stream_) {
  if (!points || n <= 0 || n >= (1LL << 31) || !(resolution > 0.0) || instance < 0 ||
      instance >= 0xFFFF || scan == 0 || scan >= (1u << 31) ||
      !table_ok(cells, cnt, capacity, counters))
    return MF_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  MapTable t = make_table(cells, cnt, capacity, counters);
  ScanParams p;
  p.points = points; p.mask = mask; p.labels = nullptr; p.lut = nullptr; p.inst_res = nullptr;
  p.lut_lo = 0; p.lut_n = 0; p.n = (int)n; p.ox = ox; p.oy = oy; p.oz = oz;
  p.res = resolution; p.inst = instance; p.scan = scan;
  p.hit = hit; p.miss = miss; p.lo_min = lo_min; p.lo_max = lo_max;
  k_map_scan_hits<<<div_up(n, 256), 256, 0, st>>>(p, t);
  MF_LAUNCH_CHECK();
  k_map_scan_free<<<div_up(n, 256), 256, 0, st>>>(p, t);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_map_integrate_labelled(const float* points, const int32_t* labels, int64_t n,
                                         float ox, float oy, float oz, const int32_t* lut,
                                         int lut_lo, int lut_n, const double* inst_resolution,
                                         uint32_t scan, float hit, float miss, float lo_min,
                                         float lo_max, void* cells, void* cnt, int64_t capacity,
                                         int32_t* counters, void* stream_) {
  if (!points || !labels || !lut || lut_n <= 0 || !inst_resolution || n <= 0 || n >= (1LL << 31) ||
      scan == 0 || scan >= (1u << 31) || !table_ok(cells, cnt, capacity, counters))
    return MF_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  MapTable t = make_table(cells, cnt, capacity, counters);
  ScanParams p;
  p.points = points; p.mask = nullptr; p.labels = labels; p.lut = lut; p.inst_res = inst_resolution;
  p.lut_lo = lut_lo; p.lut_n = lut_n; p.n = (int)n; p.ox = ox; p.oy = oy; p.oz = oz;
  p.res = 1.0; p.inst = 0; p.scan = scan;
  p.hit = hit; p.miss = miss; p.lo_min = lo_min; p.lo_max = lo_max;
  k_map_scan_hits<<<div_up(n, 256), 256, 0, st>>>(p, t);
  MF_LAUNCH_CHECK();
  k_map_scan_free<<<div_up(n, 256), 256, 0, st>>>(p, t);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_map_update_points(const double* points, int64_t m, double resolution,
                                    int instance, float update, float lo_min, float lo_max,
                                    void* cells, void* cnt, int64_t capacity, int32_t* counters,
                                    void* stream_) {
  if (!points || m <= 0 || m >= (1LL << 31) || !(resolution > 0.0) || instance < 0 ||
      instance >= 0xFFFF || !table_ok(cells, cnt, capacity, counters))
    return MF_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  MapTable t = make_table(cells, cnt, capacity, counters);
  k_map_update_count<<<div_up(m, 256), 256, 0, st>>>(points, (int)m, 1.0 / resolution, instance, t);
  MF_LAUNCH_CHECK();
  k_map_update_apply<<<div_up(m, 256), 256, 0, st>>>(points, (int)m, 1.0 / resolution, instance,
                                                     update, lo_min, lo_max, t);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_map_query_grids(const int32_t* target_index, const double* pitch,
                                  const double* origin, int T, int X, int Y, int Z,
                                  const double* res_factor, int n_instances, void* cells,
                                  void* cnt, int64_t capacity, int32_t* counters,
                                  float* grid_target, float* grid_nontarget, float* grid_empty,
                                  void* stream_) {
  if (!target_index || !pitch || !origin || T <= 0 || X <= 0 || Y <= 0 || Z <= 0 || !res_factor ||
      n_instances < 0 || !grid_target || !grid_nontarget || !grid_empty ||
      (long long)T * X * Y * Z >= (1LL << 40) || !table_ok(cells, cnt, capacity, counters))
    return MF_E_BADARG;
  MapTable t = make_table(cells, cnt, capacity, counters);
  const long long total = (long long)T * X * Y * Z;
  k_map_query_grids<<<div_up(total, 256), 256, 0, (cudaStream_t)stream_>>>(
      target_index, pitch, origin, T, X, Y, Z, res_factor, n_instances, t, grid_target,
      grid_nontarget, grid_empty);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

// growth (rare, never on the per-frame path): `counters` restarts from the re-inserted cells
extern "C" int mf_map_rehash(const void* old_cells, int64_t old_capacity, void* cells, void* cnt,
                             int64_t capacity, int32_t* counters, void* stream_) {
  if (!old_cells || old_capacity <= 0 || old_capacity > (1LL << 31) || capacity < old_capacity ||
      !table_ok(cells, cnt, capacity, counters))
    return MF_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream_;
  MapTable t = make_table(cells, cnt, capacity, counters);
  MF_CUDA_TRY(cudaMemsetAsync(counters, 0, 2 * sizeof(int32_t), st));
  int dev = 0, n_sm = 0;
  MF_CUDA_TRY(cudaGetDevice(&dev));
  MF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  k_map_rehash<<<n_sm * 8, 256, 0, st>>>((const MapCell*)old_cells, (unsigned)old_capacity, t);
  MF_LAUNCH_CHECK();
  return MF_OK;
}
#endif  // MF_HOST_EMU
