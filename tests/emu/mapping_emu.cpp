// TEST INFRASTRUCTURE: serial host execution of the kernels in morefusion_b200/csrc/mapping.cu
// (built by tests/test_mapping_emu.py with g++ -DMF_HOST_EMU).  Same argument lists as the
// C ABI entry points mf_map_* (include/morefusion_b200.h), minus the stream.
#define MF_HOST_EMU 1
#include "../../morefusion_b200/csrc/mapping.cu"

using namespace mf;

template <class F> static void run(unsigned grid, unsigned block, F f) {
  gridDim = {grid, 1, 1};
  blockDim = {block, 1, 1};
  for (unsigned b = 0; b < grid; ++b)
    for (unsigned t = 0; t < block; ++t) {
      blockIdx = {b, 0, 0};
      threadIdx = {t, 0, 0};
      f();
    }
}

static MapTable tab(void* keys, void* lo, void* stamp, void* cnt, long long cap, void* counters) {
  MapTable t;
  t.keys = (u64*)keys; t.lo = (float*)lo; t.stamp = (unsigned*)stamp; t.cnt = (unsigned*)cnt;
  t.mask = (unsigned)(cap - 1); t.counters = (int*)counters;
  return t;
}

extern "C" int emu_map_integrate(const float* points, const uint8_t* mask, int64_t n, float ox,
                                 float oy, float oz, double resolution, int instance,
                                 uint32_t scan, float hit, float miss, float lo_min, float lo_max,
                                 void* keys, void* lo, void* stamp, void* cnt, int64_t capacity,
                                 int32_t* counters, void* ray_buf, int64_t ray_capacity) {
  MapTable t = tab(keys, lo, stamp, cnt, capacity, counters);
  ScanParams p;
  p.points = points; p.mask = mask; p.n = (int)n; p.ox = ox; p.oy = oy; p.oz = oz;
  p.res = resolution; p.res_factor = 1.0 / resolution; p.inst = instance; p.scan = scan;
  p.hit = hit; p.miss = miss; p.lo_min = lo_min; p.lo_max = lo_max;
  p.ray_buf = (u64*)ray_buf; p.ray_cap = (int)ray_capacity;
  run((unsigned)n, 1, [&] { k_map_scan_rays(p, t); });          // blockDim 1: lane 0 only
  run(7, 3, [&] { k_map_free_apply((const u64*)ray_buf, (int)ray_capacity, scan, miss, lo_min, lo_max, t); });
  return 0;
}

extern "C" int emu_map_update_points(const double* points, int64_t m, double resolution,
                                     int instance, float update, float lo_min, float lo_max,
                                     void* keys, void* lo, void* stamp, void* cnt,
                                     int64_t capacity, int32_t* counters) {
  MapTable t = tab(keys, lo, stamp, cnt, capacity, counters);
  run((unsigned)m, 1, [&] { k_map_update_count(points, (int)m, 1.0 / resolution, instance, t); });
  run((unsigned)m, 1, [&] { k_map_update_apply(points, (int)m, 1.0 / resolution, instance, update, lo_min, lo_max, t); });
  return 0;
}

extern "C" int emu_map_query_grids(const int32_t* target_index, const double* pitch,
                                   const double* origin, int T, int X, int Y, int Z,
                                   const double* res_factor, int n_instances, void* keys, void* lo,
                                   void* stamp, void* cnt, int64_t capacity, int32_t* counters,
                                   float* g_target, float* g_nontarget, float* g_empty) {
  MapTable t = tab(keys, lo, stamp, cnt, capacity, counters);
  const long long total = (long long)T * X * Y * Z;
  run((unsigned)((total + 63) / 64), 64, [&] {
    k_map_query_grids(target_index, pitch, origin, T, X, Y, Z, res_factor, n_instances, t,
                      g_target, g_nontarget, g_empty);
  });
  return 0;
}

extern "C" int emu_map_rehash(const void* old_keys, const void* old_lo, const void* old_stamp,
                              int64_t old_capacity, void* keys, void* lo, void* stamp, void* cnt,
                              int64_t capacity, int32_t* counters) {
  MapTable t = tab(keys, lo, stamp, cnt, capacity, counters);
  counters[0] = 0; counters[1] = 0;
  run(5, 4, [&] { k_map_rehash((const u64*)old_keys, (const float*)old_lo, (const unsigned*)old_stamp,
                               (unsigned)old_capacity, t); });
  return 0;
}
