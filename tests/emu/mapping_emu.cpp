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

static MapTable tab(void* cells, void* cnt, long long cap, void* counters) {
  MapTable t;
  t.cells = (MapCell*)cells; t.cnt = (unsigned*)cnt; t.mask = (unsigned)(cap - 1);
  t.counters = (int*)counters;
  return t;
}

extern "C" int emu_map_integrate(const float* points, const uint8_t* mask, int64_t n, float ox,
                                 float oy, float oz, double resolution, int instance,
                                 uint32_t scan, float hit, float miss, float lo_min, float lo_max,
                                 void* cells, void* cnt, int64_t capacity, int32_t* counters) {
  MapTable t = tab(cells, cnt, capacity, counters);
  ScanParams p;
  p.points = points; p.mask = mask; p.labels = nullptr; p.lut = nullptr; p.inst_res = nullptr;
  p.lut_lo = 0; p.lut_n = 0; p.n = (int)n; p.ox = ox; p.oy = oy; p.oz = oz;
  p.res = resolution; p.inst = instance; p.scan = scan;
  p.hit = hit; p.miss = miss; p.lo_min = lo_min; p.lo_max = lo_max;
  run((unsigned)n, 1, [&] { k_map_scan_hits(p, t); });          // blockDim 1: lane 0 only
  run((unsigned)n, 1, [&] { k_map_scan_free(p, t); });
  return 0;
}

extern "C" int emu_map_integrate_labelled(const float* points, const int32_t* labels, int64_t n,
                                          float ox, float oy, float oz, const int32_t* lut,
                                          int lut_lo, int lut_n, const double* inst_resolution,
                                          uint32_t scan, float hit, float miss, float lo_min,
                                          float lo_max, void* cells, void* cnt, int64_t capacity,
                                          int32_t* counters) {
  MapTable t = tab(cells, cnt, capacity, counters);
  ScanParams p;
  p.points = points; p.mask = nullptr; p.labels = labels; p.lut = lut; p.inst_res = inst_resolution;
  p.lut_lo = lut_lo; p.lut_n = lut_n; p.n = (int)n; p.ox = ox; p.oy = oy; p.oz = oz;
  p.res = 1.0; p.inst = 0; p.scan = scan;
  p.hit = hit; p.miss = miss; p.lo_min = lo_min; p.lo_max = lo_max;
  run((unsigned)n, 1, [&] { k_map_scan_hits(p, t); });
  run((unsigned)n, 1, [&] { k_map_scan_free(p, t); });
  return 0;
}

extern "C" int emu_map_update_points(const double* points, int64_t m, double resolution,
                                     int instance, float update, float lo_min, float lo_max,
                                     void* cells, void* cnt, int64_t capacity, int32_t* counters) {
  MapTable t = tab(cells, cnt, capacity, counters);
  run((unsigned)m, 1, [&] { k_map_update_count(points, (int)m, 1.0 / resolution, instance, t); });
  run((unsigned)m, 1, [&] { k_map_update_apply(points, (int)m, 1.0 / resolution, instance, update, lo_min, lo_max, t); });
  return 0;
}

extern "C" int emu_map_query_grids(const int32_t* target_index, const double* pitch,
                                   const double* origin, int T, int X, int Y, int Z,
                                   const double* res_factor, int n_instances, void* cells,
                                   void* cnt, int64_t capacity, int32_t* counters,
                                   float* g_target, float* g_nontarget, float* g_empty) {
  MapTable t = tab(cells, cnt, capacity, counters);
  const long long total = (long long)T * X * Y * Z;
  run((unsigned)((total + 63) / 64), 64, [&] {
    k_map_query_grids(target_index, pitch, origin, T, X, Y, Z, res_factor, n_instances, t,
                      g_target, g_nontarget, g_empty);
  });
  return 0;
}

extern "C" int emu_map_rehash(const void* old_cells, int64_t old_capacity, void* cells, void* cnt,
                              int64_t capacity, int32_t* counters) {
  MapTable t = tab(cells, cnt, capacity, counters);
  counters[0] = 0; counters[1] = 0;
  run(5, 4, [&] { k_map_rehash((const MapCell*)old_cells, (unsigned)old_capacity, t); });
  return 0;
}
