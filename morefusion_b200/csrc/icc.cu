// Fused IterativeCollisionCheck (ICC) refinement for sm_100a.
//
// Replaces, per iteration, everything the reference launches from
//   morefusion/contrib/iterative_collision_check_link.py:31-99   (forward graph)
//   + chainer's reverse pass through it
//   + chainer.optimizers.Adam.update
//     (examples/ycb_video/pose_refinement/check_iterative_collision_check_link.py:48-50,73-79)
// i.e. ~2N TDF launches, ~40N small cupy/chainer kernels and >=2N host syncs, by ONE persistent
// cooperative kernel that runs all iterations.  A scene is owned by a group of G CTAs that meet
// at a group barrier between phases; independent scenes (different groups) never synchronise,
// so a batch of scenes fills the GPU.
//
// Phases of one iteration (work item = (target grid i, source object j, 256-point chunk)):
//   P1 scatter   x = R_j p + t_j, voxel frame of grid i, 64-bit atomicMin of (distance, point id)
//                into keys[i][self|other]                                (K5, tdf.py:51-79)
//   P2 weights   per voxel: winner's sdf -> grid-wide max of relu(sdf + offset)   (:198-204)
//   P3 loss      per voxel: grid / surface / inside, F.maximum with the other-object grid,
//                partial sums of reward/penalty; stores the 4 coefficients backward needs
//   P4 backward  totals -> d loss / d tdf; per point gather over the voxels it won (K6, :119-145),
//                chain through x = R p + t into per-item (gt, gR) partials (fixed-order trees)
//   P5 update    per object: reduce partials, quaternion_matrix backward
//                (quaternion_matrix.py:41-49), Chainer-form Adam; reset keys for the next iteration
// Reductions: the four loss sums and the 12 per-object (gR | gt) sums are accumulated in FP64 from
// exactly-representable terms (fp32 values / products of two fp32 values), then rounded to fp32
// once.  The result is therefore independent of the summation order (up to 1e-16 relative), which
// is what closed-loop parity needs: Adam divides every gradient component by its own sqrt(v), so
// fp32 summation noise on a nearly-cancelling component turns into O(alpha) differences in the
// pose after a few dozen iterations.  The oracle (oracle/icc.py) sums the same terms in fp64.
#include <cooperative_groups.h>

#include "common.cuh"

namespace mf {

constexpr int kIccThreads = 256;
constexpr int kIccMaxObj = 32;
constexpr int CHUNK_PTS = 256;          // points per work chunk (CHUNK in the host binding)
constexpr unsigned long long kKeyEmpty = 0xFFFFFFFFFFFFFFFFull;

struct IccParams {
  int S, D;
  float threshold, sdf_offset;
  const int* scene_obj_off;    // [S+1]
  const int* obj_pt_off;       // [Ntot+1]
  const int* scene_chunk_off;  // [S+1]
  const int* chunk_obj;        // [Ctot] global object id of each 256-point chunk
  const int* chunk_start;      // [Ctot] global index of the chunk's first point
  const int* scene_slot_off;   // [S+1] prefix of N_s * C_s
  const int* obj_chunk_off;    // [Ntot+1] first chunk of every object (chunks of an object are contiguous)
  const float* points;         // [Ptot,3] CAD frame
  const float* sdf;            // [Ptot]
  const float* pitch;          // [Ntot]
  const float* origin;         // [Ntot,3]
  const float* grid_target;    // [Ntot,V]
  const float* gne;            // [Ntot,V]
  float* q;                    // [Ntot,4]
  float* t;                    // [Ntot,3]
  float* adam;                 // [4][Ntot,7]... laid out m_q[Ntot,4] m_t[Ntot,3] v_q[Ntot,4] v_t[Ntot,3]
  unsigned long long* keys;    // [Ntot][2][V]
  float4* coefs;               // [Ntot][V]
  unsigned int* maxbits;       // [Ntot][2]
  double* partials;            // [S][G][4]
  double* slots;               // [S][G][kIccMaxObj][12]
  unsigned int* barrier;       // [S][1 + kBarSub]
  float* loss;                 // [S][n_iter]
  float* grads;                // [Ntot,7] gradient of the last iteration (gq | gt)
  int Ntot;
  int G, n_iter, update;
  float one_minus_beta1, one_minus_beta2, eps, eta;
  unsigned long long* phase_ns;  // optional [S][n_iter][8] globaltimer stamps at phase boundaries (profiling)
};

struct IccAlpha {
  float aq[128];
  float at[128];
};

// ------------------------------------------------------------------ helpers
// Two-level barrier over the G CTAs of a scene.  Same-address atomics from different SMs
// serialise in L2 (~15-30 cycles each), so 592 arrivals on one word cost ~10 us; here a CTA
// arrives on one of kBarSub sub-counters and only the last arriver of a sub-group touches the
// master word that everybody polls.  All counters are monotonic (never reset inside the kernel).
constexpr int kBarSub = 24;
__device__ __forceinline__ void group_barrier(unsigned int* bar /* [1 + kBarSub] */, int G, int cta,
                                              unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nsub = G < kBarSub ? G : kBarSub;
    const int sub = cta % nsub;
    const unsigned int sub_size = (unsigned int)((G - sub + nsub - 1) / nsub);
    __threadfence();
    unsigned int old = atomicAdd(bar + 1 + sub, 1u);
    if ((old + 1u) % sub_size == 0u) atomicAdd(bar, 1u);
    epoch += (unsigned)nsub;
    long long t0 = clock64();
    while (true) {
      unsigned int v = *((volatile unsigned int*)bar);
      if ((int)(v - epoch) >= 0) break;
      if (clock64() - t0 > 6000000000LL) __trap();   // never hang the GPU
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV values per thread; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* [8*NV] */) {
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum_d(v[k]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[warp * NV + k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int w = 0; w < kIccThreads / 32; ++w) s += smem[w * NV + k];
      v[k] = s;
    }
  }
}

__device__ __forceinline__ void quat_to_rot9(const float q[4], float R[9]) {
  float n = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q[0], q[0]), __fmul_rn(q[1], q[1])),
                                __fmul_rn(q[2], q[2])), __fmul_rn(q[3], q[3]));
  float s = __fsqrt_rn(__fdiv_rn(2.f, n));
  float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
  float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
  float xw = x * w, yw = y * w, zw = z * w;
  R[0] = 1.f - yy - zz; R[1] = xy - zw;        R[2] = xz + yw;
  R[3] = xy + zw;       R[4] = 1.f - xx - zz;  R[5] = yz - xw;
  R[6] = xz - yw;       R[7] = yz + xw;        R[8] = 1.f - xx - yy;
}

__device__ __forceinline__ void quat_grad(const float q[4], const float gR[9], float gq[4]) {
  float n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  float s = sqrtf(2.f / n);
  float qs[4] = {q[0] * s, q[1] * s, q[2] * s, q[3] * s};
  float g10 = -gR[5] + gR[7], g11 = -gR[4] - gR[8], g12 = gR[1] + gR[3], g13 = gR[2] + gR[6];
  float g20 = gR[2] - gR[6], g22 = -gR[0] - gR[8], g23 = gR[5] + gR[7];
  float g30 = -gR[1] + gR[3], g33 = -gR[0] - gR[4];
  float gqs[4];
  gqs[0] = g10 * qs[1] + g20 * qs[2] + g30 * qs[3];
  gqs[1] = (g10 * qs[0] + g11 * qs[1] + g12 * qs[2] + g13 * qs[3]) + g11 * qs[1];
  gqs[2] = (g20 * qs[0] + g22 * qs[2] + g23 * qs[3]) + (g12 * qs[1] + g22 * qs[2]);
  gqs[3] = (g30 * qs[0] + g33 * qs[3]) + (g13 * qs[1] + g23 * qs[2] + g33 * qs[3]);
  float gs = gqs[0] * q[0] + gqs[1] * q[1] + gqs[2] * q[2] + gqs[3] * q[3];
  float gn = gs * (-s / (2.f * n));
#pragma unroll
  for (int k = 0; k < 4; ++k) gq[k] = gqs[k] * s + gn * 2.f * q[k];
}

__device__ __forceinline__ int ksize_of(float pitch, float trunc) {
  int k = (int)ceilf(trunc / pitch);
  if ((k & 1) == 0) k += 1;
  return k;
}

struct SceneCtx {
  int o0, N, c0, C, slot0;
};

// x = R p + t  ->  f = (x - origin_i) / pitch_i   (same expression in P1 and P4: bit-identical)
__device__ __forceinline__ void point_in_grid(const float* R, const float* tt, float px, float py,
                                              float pz, float ox, float oy, float oz, float pitch,
                                              float& fx, float& fy, float& fz) {
  float x = R[0] * px + R[1] * py + R[2] * pz + tt[0];
  float y = R[3] * px + R[4] * py + R[5] * pz + tt[1];
  float z = R[6] * px + R[7] * py + R[8] * pz + tt[2];
  fx = (x - ox) / pitch;
  fy = (y - oy) / pitch;
  fz = (z - oz) / pitch;
}

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define ICC_STAMP(k)                                                                      \
  if (p.phase_ns && cta == 0 && tid == 0)                                                 \
    p.phase_ns[((size_t)scene * p.n_iter + it) * 8 + (k)] = gtime_ns();

// ------------------------------------------------------------------ the persistent kernel
__global__ void __launch_bounds__(kIccThreads, 4)
k_icc_run(IccParams p, IccAlpha alpha) {
  __shared__ float sR[kIccMaxObj][9];
  __shared__ float sT[kIccMaxObj][3];
  __shared__ double sred[8 * 12];
  __shared__ float stot[4];
  // per-scene constants staged once (they are read by every work item of every iteration)
  __shared__ float sPitch[kIccMaxObj];
  __shared__ float sOrigin[kIccMaxObj][3];
  __shared__ int sPtEnd[kIccMaxObj];
  // per-warp running sums of (gR | gt) per source object: no block barrier per work item
  __shared__ double sAcc[kIccThreads / 32][kIccMaxObj][12];

  const int tid = threadIdx.x;
  const int scene = blockIdx.x / p.G, cta = blockIdx.x % p.G;
  SceneCtx sc;
  sc.o0 = p.scene_obj_off[scene];
  sc.N = p.scene_obj_off[scene + 1] - sc.o0;
  sc.c0 = p.scene_chunk_off[scene];
  sc.C = p.scene_chunk_off[scene + 1] - sc.c0;
  sc.slot0 = p.scene_slot_off[scene];
  const int D = p.D, V = D * D * D;
  unsigned int* bar = p.barrier + (size_t)scene * (1 + kBarSub);
  unsigned int epoch = 0;
  if (tid < sc.N) {
    int o = sc.o0 + tid;
    sPitch[tid] = p.pitch[o];
    sOrigin[tid][0] = p.origin[3 * o]; sOrigin[tid][1] = p.origin[3 * o + 1]; sOrigin[tid][2] = p.origin[3 * o + 2];
    sPtEnd[tid] = p.obj_pt_off[o + 1];
  }
  __syncthreads();
  const int warp_id = tid >> 5, lane_id = tid & 31;

  for (int it = 0; it < p.n_iter; ++it) {
    ICC_STAMP(0)
    // ---- P0: rotations of this scene's objects (q, t may have been updated by other CTAs)
    if (tid < sc.N) {
      int o = sc.o0 + tid;
      float q[4] = {__ldcg(p.q + 4 * o), __ldcg(p.q + 4 * o + 1), __ldcg(p.q + 4 * o + 2),
                    __ldcg(p.q + 4 * o + 3)};
      float R[9];
      quat_to_rot9(q, R);
#pragma unroll
      for (int k = 0; k < 9; ++k) sR[tid][k] = R[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) sT[tid][k] = __ldcg(p.t + 3 * o + k);
    }
    __syncthreads();

    // ---- P1: scatter.  Work unit = (target grid i, 32 consecutive points of one 256-point
    // chunk): the source object j, its pose and the grid geometry are warp-uniform (registers /
    // broadcast reads), no per-lane index arithmetic.  Units are interleaved over the warps of
    // all CTAs of the scene.  Per point the 3 x ks per-axis voxel indices and squared offsets
    // are formed once; a candidate voxel then costs two adds, the square root, one multiply and
    // the 64-bit min.
    {
      const int n_units = sc.N * sc.C * (CHUNK_PTS / 32);
      for (int u = cta * (kIccThreads / 32) + warp_id; u < n_units; u += p.G * (kIccThreads / 32)) {
        const int sub = u & (CHUNK_PTS / 32 - 1);
        const int ic = u / (CHUNK_PTS / 32);
        const int il = ic / sc.C, c = ic - il * sc.C;
        const int gj = __ldg(p.chunk_obj + sc.c0 + c), jl = gj - sc.o0;
        const int pt = __ldg(p.chunk_start + sc.c0 + c) + sub * 32 + lane_id;
        if (pt >= sPtEnd[jl]) continue;
        const int gi = sc.o0 + il;
        const float pitch = sPitch[il];
        const float trunc = p.threshold * pitch;
        const int ks = ksize_of(pitch, trunc), half = ks / 2;
        float fx, fy, fz;
        point_in_grid(sR[jl], sT[jl], p.points[3 * pt], p.points[3 * pt + 1], p.points[3 * pt + 2],
                      sOrigin[il][0], sOrigin[il][1], sOrigin[il][2], pitch, fx, fy, fz);
        const float rx = roundf(fx), ry = roundf(fy), rz = roundf(fz);
        if (!(rx >= (float)(-half) && rx <= (float)(D - 1 + half) && ry >= (float)(-half) &&
              ry <= (float)(D - 1 + half) && rz >= (float)(-half) && rz <= (float)(D - 1 + half)))
          continue;
        unsigned long long* keys = p.keys + ((size_t)gi * 2 + (gi == gj ? 0 : 1)) * V;
        if (half == 1) {
          // the common case (voxel_threshold = 2): 3 offsets per axis, fully unrolled
          int vx[3], vy[3], vz[3];
          float sx[3], sy[3], sz[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int ix = (int)(rx + (float)(k - 1)), iy = (int)(ry + (float)(k - 1)),
                      iz = (int)(rz + (float)(k - 1));
            const float dx = fx - (float)ix, dy = fy - (float)iy, dz = fz - (float)iz;
            vx[k] = (ix >= 0 && ix < D) ? ix * D * D : -1;
            vy[k] = (iy >= 0 && iy < D) ? iy * D : -1;
            vz[k] = (iz >= 0 && iz < D) ? iz : -1;
            sx[k] = dx * dx; sy[k] = dy * dy; sz[k] = dz * dz;
          }
          // x-slab of the point's own voxel first (the most likely winners), then the two
          // neighbouring slabs.  Per slab the 9 current keys are loaded together (independent L2
          // reads, one round trip) and a candidate only issues its 64-bit atomic if it beats the
          // value it saw: the key of a voxel only ever decreases, so a stale read can cost a
          // redundant atomic but never lose a winner.  L2 serves ~50 G 64-bit atomics/s; this
          // cuts the candidates per voxel from ~27 to a handful.
#pragma unroll
          for (int ai = 0; ai < 3; ++ai) {
            const int a = ai == 0 ? 1 : (ai == 1 ? 0 : 2);
            if (vx[a] < 0) continue;
            unsigned long long cur[9];
#pragma unroll
            for (int b2 = 0; b2 < 3; ++b2)
#pragma unroll
              for (int c2 = 0; c2 < 3; ++c2)
                cur[b2 * 3 + c2] = (vy[b2] >= 0 && vz[c2] >= 0)
                                       ? __ldcg(keys + vx[a] + vy[b2] + vz[c2]) : 0ull;
#pragma unroll
            for (int b2 = 0; b2 < 3; ++b2) {
              const float sxy = sx[a] + sy[b2];
#pragma unroll
              for (int c2 = 0; c2 < 3; ++c2) {
                const float dist = pitch * sqrtf(sxy + sz[c2]);
                const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) |
                                               (unsigned long long)(unsigned int)pt;
                if (dist < trunc && key < cur[b2 * 3 + c2])
                  atomicMin(keys + vx[a] + vy[b2] + vz[c2], key);
              }
            }
          }
        } else {
          for (int dx = -half; dx <= half; ++dx) {
            const int ix = (int)(rx + (float)dx);
            if (ix < 0 || ix >= D) continue;
            const float ddx = fx - (float)ix;
            for (int dy = -half; dy <= half; ++dy) {
              const int iy = (int)(ry + (float)dy);
              if (iy < 0 || iy >= D) continue;
              const float ddy = fy - (float)iy;
              for (int dz = -half; dz <= half; ++dz) {
                const int iz = (int)(rz + (float)dz);
                if (iz < 0 || iz >= D) continue;
                const float ddz = fz - (float)iz;
                const float dist = pitch * sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
                if (dist < trunc) {
                  const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) |
                                                 (unsigned long long)(unsigned int)pt;
                  atomicMin(keys + (ix * D + iy) * D + iz, key);
                }
              }
            }
          }
        }
      }
    }
    group_barrier(bar, p.G, cta, epoch);

    ICC_STAMP(1)
    // ---- P2: winners' weights -> per-(grid, kind) max
    {
      const int n_gk = sc.N * 2;
      const int blocks_per = (V + kIccThreads - 1) / kIccThreads;
      for (int w = cta; w < n_gk * blocks_per; w += p.G) {
        int gk = w / blocks_per, v = (w - gk * blocks_per) * kIccThreads + tid;
        int gi = sc.o0 + (gk >> 1), kind = gk & 1;
        float wv = 0.f;
        if (v < V && !(kind == 1 && sc.N == 1)) {
          unsigned long long key = __ldcg(p.keys + ((size_t)gi * 2 + kind) * V + v);
          if (key != kKeyEmpty) {
            float wr = p.sdf[(unsigned int)(key & 0xffffffffu)] + (kind == 0 ? p.sdf_offset : 0.f);
            wv = wr < 0.f ? 0.f : wr;
          }
        }
        float m = warp_max(wv);
        if ((tid & 31) == 0 && m > 0.f) atomicMax(p.maxbits + gi * 2 + kind, __float_as_uint(m));
      }
    }
    group_barrier(bar, p.G, cta, epoch);

    ICC_STAMP(2)
    // ---- P3: grids, loss partial sums, backward coefficients
    {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};   // rew_num, rew_den, pen_num, pen_den (fp64 sums of fp32 terms)
      const int blocks_per = (V + kIccThreads - 1) / kIccThreads;
      for (int w = cta; w < sc.N * blocks_per; w += p.G) {
        int il = w / blocks_per, v = (w - il * blocks_per) * kIccThreads + tid;
        if (v >= V) continue;
        int gi = sc.o0 + il;
        float pitch = p.pitch[gi], trunc = p.threshold * pitch;
        // self
        unsigned long long key = __ldcg(p.keys + ((size_t)gi * 2) * V + v);
        bool hit = key != kKeyEmpty;
        float tdf = hit ? __uint_as_float((unsigned int)(key >> 32)) : trunc;
        float g = 1.f - tdf / trunc;
        float wr = (hit ? p.sdf[(unsigned int)(key & 0xffffffffu)] : -1.f) + p.sdf_offset;
        bool mask = wr < 0.f;
        float wmax = __uint_as_float(__ldcg(p.maxbits + gi * 2));
        float wi = (mask ? 0.f : wr) / wmax;                 // 0/0 -> NaN as in the reference (:204)
        float ws = mask ? wi : 1.f - wi;
        float surface = g * ws, inside = g * wi;
        float gt = p.grid_target[(size_t)gi * V + v];
        float gne = p.gne[(size_t)gi * V + v];
        float b1 = 0.f;
        if (sc.N > 1) {
          float omax = __uint_as_float(__ldcg(p.maxbits + gi * 2 + 1));
          if (omax > 0.f) {                                   // else grid_other is NaN: skipped (:82)
            unsigned long long ko = __ldcg(p.keys + ((size_t)gi * 2 + 1) * V + v);
            bool ho = ko != kKeyEmpty;
            float tdfo = ho ? __uint_as_float((unsigned int)(ko >> 32)) : trunc;
            float go = 1.f - tdfo / trunc;
            float wro = ho ? p.sdf[(unsigned int)(ko & 0xffffffffu)] : -1.f;
            float wio = (wro < 0.f ? 0.f : wro) / omax;
            float oin = go * wio;
            if (!(gne >= oin)) {                               // F.maximum: first arg wins ties
              gne = oin;
              b1 = wio * inside;
            }
          }
        }
        p.coefs[(size_t)gi * V + v] = make_float4(ws * gt, wi * gne, wi, b1);
        acc[0] += (double)(surface * gt);
        acc[1] += (double)gt;
        acc[2] += (double)(inside * gne);
        acc[3] += (double)inside;
      }
      block_sum<4>(acc, sred);
      if (tid == 0) {
        double* dst = p.partials + ((size_t)scene * p.G + cta) * 4;
        dst[0] = acc[0]; dst[1] = acc[1]; dst[2] = acc[2]; dst[3] = acc[3];
      }
    }
    group_barrier(bar, p.G, cta, epoch);

    ICC_STAMP(3)
    // ---- P4: totals, then gather-backward
    if (tid < 32) {
      double a[4] = {0.0, 0.0, 0.0, 0.0};
      for (int c = tid; c < p.G; c += 32) {
        const double* src = p.partials + ((size_t)scene * p.G + c) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] += __ldcg(src + k);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] = warp_sum_d(a[k]);
      if (tid == 0) {
        // the four sums are rounded to fp32 once; everything after is fp32 like the reference
#pragma unroll
        for (int k = 0; k < 4; ++k) stot[k] = (float)a[k];
        if (cta == 0)
          p.loss[(size_t)scene * p.n_iter + it] = stot[2] / stot[3] - stot[0] / stot[1];
      }
    }
    __syncthreads();
    const float c_rw = 1.f / stot[1];
    const float c_in0 = 1.f / stot[3];
    const float c_in1 = stot[2] / (stot[3] * stot[3]);

    for (int e = tid; e < (kIccThreads / 32) * kIccMaxObj * 12; e += kIccThreads)
      (&sAcc[0][0][0])[e] = 0.0;
    __syncthreads();
    // Voxel-centric gather (the structure of the reference's backward kernel, tdf.py:119-145):
    // every voxel that has a winner contributes  unit(f - v) * d loss/d tdf[v]  to its winning
    // point; the dependent chain per voxel is only key -> (coefficients, point) -> arithmetic.
    {
      const int n_gk = sc.N * 2;
      const int blocks_per = (V + kIccThreads - 1) / kIccThreads;
      for (int w = cta; w < n_gk * blocks_per; w += p.G) {
        const int gk = w / blocks_per, v = (w - gk * blocks_per) * kIccThreads + tid;
        const int il = gk >> 1, kind = gk & 1, gi = sc.o0 + il;
        double g12[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) g12[k] = 0.0;
        bool any = false;
        int jl = 1 << 30;
        if (v < V && !(kind == 1 && sc.N == 1)) {
          const unsigned long long key = __ldcg(p.keys + ((size_t)gi * 2 + kind) * V + v);
          if (key != kKeyEmpty) {
            const float4 cf = __ldcg(p.coefs + (size_t)gi * V + v);
            const float dg = kind == 0 ? (cf.y * c_in0 - cf.z * c_in1 - cf.x * c_rw) : cf.w * c_in0;
            if (dg != 0.f) {
              const int pt = (int)(unsigned int)(key & 0xffffffffu);
              jl = 0;
              while (jl + 1 < sc.N && pt >= sPtEnd[jl]) ++jl;
              const float pitch = sPitch[il];
              const float trunc = p.threshold * pitch;
              const float px = p.points[3 * pt], py = p.points[3 * pt + 1], pz = p.points[3 * pt + 2];
              float fx, fy, fz;
              point_in_grid(sR[jl], sT[jl], px, py, pz, sOrigin[il][0], sOrigin[il][1],
                            sOrigin[il][2], pitch, fx, fy, fz);
              const int iz = v % D, iy = (v / D) % D, ix = v / (D * D);
              const float ddx = fx - (float)ix, ddy = fy - (float)iy, ddz = fz - (float)iz;
              const float n = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
              if (n > 0.f) {
                const float gtdf = -dg / trunc;
                const float ax = ddx / n * gtdf, ay = ddy / n * gtdf, az = ddz / n * gtdf;
                // x = R p + t :  gt += gx ; gR += gx (x) p
                // exact fp64 products of fp32 factors
                const double dax = ax, day = ay, daz = az, dpx = px, dpy = py, dpz = pz;
                g12[0] = dax * dpx; g12[1] = dax * dpy; g12[2] = dax * dpz; g12[3] = dax;
                g12[4] = day * dpx; g12[5] = day * dpy; g12[6] = day * dpz; g12[7] = day;
                g12[8] = daz * dpx; g12[9] = daz * dpy; g12[10] = daz * dpz; g12[11] = daz;
                any = true;
              }
            }
          }
        }
        // winners in one warp's 32 voxels may belong to different source objects: objects are
        // taken in order of first occurrence, each with a fixed-order masked tree
        unsigned todo = __ballot_sync(0xffffffffu, any);
        while (todo) {
          const int leader = __ffs(todo) - 1;
          const int jcur = __shfl_sync(0xffffffffu, jl, leader);
          const bool mine = any && (jl == jcur);
          double sums[12];
#pragma unroll
          for (int k = 0; k < 12; ++k) sums[k] = warp_sum_d(mine ? g12[k] : 0.0);
          if (lane_id == 0)
#pragma unroll
            for (int k = 0; k < 12; ++k) sAcc[warp_id][jcur][k] += sums[k];
          todo &= ~__ballot_sync(0xffffffffu, mine);
        }
      }
    }
    __syncthreads();
    // one slot per (CTA, object): warps summed in fixed order
    for (int e = tid; e < sc.N * 12; e += kIccThreads) {
      int jl = e / 12, k = e - jl * 12;
      double s = 0.0;
#pragma unroll
      for (int wv = 0; wv < kIccThreads / 32; ++wv) s += sAcc[wv][jl][k];
      p.slots[(((size_t)scene * p.G + cta) * kIccMaxObj + jl) * 12 + k] = s;
    }
    group_barrier(bar, p.G, cta, epoch);

    ICC_STAMP(4)
    // ---- P5: per-object reduction + quaternion backward + Chainer Adam; reset keys
    for (int jl = cta; jl < sc.N; jl += p.G) {
      {
        int gj = sc.o0 + jl;
        // all threads: slots of the G CTAs (thread-strided), then a fixed-order block tree
        double a[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k] = 0.0;
        for (int c = tid; c < p.G; c += kIccThreads) {
          const double* slot = p.slots + (((size_t)scene * p.G + c) * kIccMaxObj + jl) * 12;
#pragma unroll
          for (int k = 0; k < 12; ++k) a[k] += __ldcg(slot + k);
        }
        block_sum<12>(a, sred);
        if (tid == 0) {
          float q[4] = {p.q[4 * gj], p.q[4 * gj + 1], p.q[4 * gj + 2], p.q[4 * gj + 3]};
          float gR[9] = {(float)a[0], (float)a[1], (float)a[2], (float)a[4], (float)a[5],
                         (float)a[6], (float)a[8], (float)a[9], (float)a[10]};
          float g7[7];
          quat_grad(q, gR, g7);
          g7[4] = (float)a[3]; g7[5] = (float)a[7]; g7[6] = (float)a[11];
#pragma unroll
          for (int k = 0; k < 7; ++k) p.grads[(size_t)gj * 7 + k] = g7[k];
          if (p.update) {
            float* mq = p.adam;
            float* mt = mq + (size_t)p.Ntot * 4;
            float* vq = mt + (size_t)p.Ntot * 3;
            float* vt = vq + (size_t)p.Ntot * 4;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
              bool isq = k < 4;
              float* pm = isq ? mq + 4 * gj + k : mt + 3 * gj + (k - 4);
              float* pv = isq ? vq + 4 * gj + k : vt + 3 * gj + (k - 4);
              float* pp = isq ? p.q + 4 * gj + k : p.t + 3 * gj + (k - 4);
              float g = g7[k];
              float m = *pm, v = *pv;
              m = m + p.one_minus_beta1 * (g - m);
              v = v + p.one_minus_beta2 * (g * g - v);
              float a_t = isq ? alpha.aq[it] : alpha.at[it];
              float step = a_t * m / (sqrtf(v) + p.eps);
              *pp = *pp - p.eta * step;
              *pm = m;
              *pv = v;
            }
          }
        }
      }
    }
    ICC_STAMP(5)
    if (it + 1 < p.n_iter) {
      // reset for the next iteration (dense: 16 B per voxel-kind, L2 resident)
      const size_t nk = (size_t)sc.N * 2 * V;
      unsigned long long* k0 = p.keys + (size_t)sc.o0 * 2 * V;
      for (size_t e = (size_t)cta * kIccThreads + tid; e < nk; e += (size_t)p.G * kIccThreads)
        k0[e] = kKeyEmpty;
      for (int e = cta * kIccThreads + tid; e < sc.N * 2; e += p.G * kIccThreads)
        p.maxbits[sc.o0 * 2 + e] = 0u;
      group_barrier(bar, p.G, cta, epoch);
    }
  }
}

__global__ void k_icc_init(unsigned long long* keys, size_t n_keys, unsigned int* maxbits, int n_max,
                           unsigned int* barrier, int S) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_keys) keys[i] = kKeyEmpty;
  if (i < (size_t)n_max) maxbits[i] = 0u;
  if (i < (size_t)S * (1 + kBarSub)) barrier[i] = 0u;
}

static size_t al(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace mf

using namespace mf;

struct IccLayout {
  size_t keys, coefs, maxbits, partials, slots, barrier, total;
};

static IccLayout icc_layout(int Ntot, int D, int S, int G, int n_slots) {
  size_t V = (size_t)D * D * D;
  IccLayout L;
  size_t o = 0;
  L.keys = o; o += al((size_t)Ntot * 2 * V * 8);
  L.coefs = o; o += al((size_t)Ntot * V * 16);
  L.maxbits = o; o += al((size_t)Ntot * 2 * 4);
  L.partials = o; o += al((size_t)S * G * 32);
  L.slots = o; o += al((size_t)S * G * kIccMaxObj * 96);   // one fp64 (gR|gt) slot per (CTA, object)
  (void)n_slots;
  L.barrier = o; o += al((size_t)S * (1 + kBarSub) * 4);
  L.total = o;
  return L;
}

extern "C" int mf_icc_max_group_size(int n_scenes) {
  int dev = 0, sms = 0, per = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_icc_run, kIccThreads, 0) != cudaSuccess)
    return -1;
  if (n_scenes <= 0) return -1;
  return (sms * per) / n_scenes;
}

extern "C" size_t mf_icc_workspace_bytes(int n_objects_total, int voxel_dim, int n_scenes,
                                         int group_size, int n_slots) {
  return icc_layout(n_objects_total, voxel_dim, n_scenes, group_size, n_slots).total;
}

extern "C" int mf_icc_run_profiled(
    int, int, int, float, float, const int32_t*, const int32_t*, const int32_t*, const int32_t*,
    const int32_t*, const int32_t*, const int32_t*, int, const float*, const float*, const float*,
    const float*, const float*, const float*, float*, float*, float*, int, int, const float*,
    const float*, double, double, double, double, float*, float*, int, void*, size_t,
    unsigned long long*, void*);

extern "C" int mf_icc_run(
    int n_scenes, int n_objects_total, int voxel_dim, float voxel_threshold, float sdf_offset,
    const int32_t* scene_obj_off, const int32_t* obj_pt_off, const int32_t* scene_chunk_off,
    const int32_t* chunk_obj, const int32_t* chunk_start, const int32_t* scene_slot_off,
    const int32_t* obj_chunk_off, int n_slots,
    const float* points, const float* sdf, const float* pitch, const float* origin,
    const float* grid_target, const float* grid_nontarget_empty,
    float* quaternion, float* translation, float* adam_state,
    int n_iter, int update, const float* alpha_q_host, const float* alpha_t_host,
    double beta1, double beta2, double eps, double eta,
    float* loss_history, float* grads, int group_size,
    void* workspace, size_t workspace_bytes, void* stream_) {
  return mf_icc_run_profiled(n_scenes, n_objects_total, voxel_dim, voxel_threshold, sdf_offset,
                             scene_obj_off, obj_pt_off, scene_chunk_off, chunk_obj, chunk_start,
                             scene_slot_off, obj_chunk_off, n_slots, points, sdf, pitch, origin,
                             grid_target, grid_nontarget_empty, quaternion, translation, adam_state,
                             n_iter, update, alpha_q_host, alpha_t_host, beta1, beta2, eps, eta,
                             loss_history, grads, group_size, workspace, workspace_bytes, nullptr,
                             stream_);
}

extern "C" int mf_icc_run_profiled(
    int n_scenes, int n_objects_total, int voxel_dim, float voxel_threshold, float sdf_offset,
    const int32_t* scene_obj_off, const int32_t* obj_pt_off, const int32_t* scene_chunk_off,
    const int32_t* chunk_obj, const int32_t* chunk_start, const int32_t* scene_slot_off,
    const int32_t* obj_chunk_off, int n_slots,
    const float* points, const float* sdf, const float* pitch, const float* origin,
    const float* grid_target, const float* grid_nontarget_empty,
    float* quaternion, float* translation, float* adam_state,
    int n_iter, int update, const float* alpha_q_host, const float* alpha_t_host,
    double beta1, double beta2, double eps, double eta,
    float* loss_history, float* grads, int group_size,
    void* workspace, size_t workspace_bytes, unsigned long long* phase_ns, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_scenes <= 0 || n_objects_total <= 0 || voxel_dim <= 0 || n_iter <= 0 || n_iter > 128)
    return MF_E_BADARG;
  if (!scene_obj_off || !obj_pt_off || !scene_chunk_off || !chunk_obj || !chunk_start ||
      !scene_slot_off || !obj_chunk_off || !points || !sdf || !pitch || !origin || !grid_target ||
      !grid_nontarget_empty || !quaternion || !translation || !loss_history || !grads || !workspace)
    return MF_E_BADARG;
  if (update && (!adam_state || !alpha_q_host || !alpha_t_host)) return MF_E_BADARG;
  int maxG = mf_icc_max_group_size(n_scenes);
  if (maxG < 1) return MF_E_UNSUPPORTED;     // more scenes than co-resident CTAs: split the batch
  int G = group_size > 0 ? group_size : maxG;
  if (G > maxG) G = maxG;
  IccLayout L = icc_layout(n_objects_total, voxel_dim, n_scenes, G, n_slots);
  if (workspace_bytes < L.total) return MF_E_WORKSPACE;
  char* ws = (char*)workspace;
  IccParams p;
  p.S = n_scenes; p.D = voxel_dim; p.threshold = voxel_threshold; p.sdf_offset = sdf_offset;
  p.scene_obj_off = scene_obj_off; p.obj_pt_off = obj_pt_off; p.scene_chunk_off = scene_chunk_off;
  p.chunk_obj = chunk_obj; p.chunk_start = chunk_start; p.scene_slot_off = scene_slot_off;
  p.obj_chunk_off = obj_chunk_off;
  p.points = points; p.sdf = sdf; p.pitch = pitch; p.origin = origin;
  p.grid_target = grid_target; p.gne = grid_nontarget_empty;
  p.q = quaternion; p.t = translation; p.adam = adam_state;
  p.keys = (unsigned long long*)(ws + L.keys);
  p.coefs = (float4*)(ws + L.coefs);
  p.maxbits = (unsigned int*)(ws + L.maxbits);
  p.partials = (double*)(ws + L.partials);
  p.slots = (double*)(ws + L.slots);
  p.barrier = (unsigned int*)(ws + L.barrier);
  p.loss = loss_history; p.grads = grads;
  p.Ntot = n_objects_total; p.G = G; p.n_iter = n_iter; p.update = update;
  // chainer: `m += (1 - hp.beta1) * (grad - m)` with python-float hyperparameters -> the
  // difference is taken in double and only then rounded to the array dtype
  p.one_minus_beta1 = (float)(1.0 - beta1);
  p.one_minus_beta2 = (float)(1.0 - beta2);
  p.eps = (float)eps; p.eta = (float)eta;
  p.phase_ns = phase_ns;
  IccAlpha alpha;
  for (int i = 0; i < 128; ++i) {
    alpha.aq[i] = (update && i < n_iter) ? alpha_q_host[i] : 0.f;
    alpha.at[i] = (update && i < n_iter) ? alpha_t_host[i] : 0.f;
  }
  size_t V = (size_t)voxel_dim * voxel_dim * voxel_dim;
  size_t n_keys = (size_t)n_objects_total * 2 * V;
  k_icc_init<<<div_up((long long)n_keys, 256), 256, 0, stream>>>(p.keys, n_keys, p.maxbits,
                                                                 n_objects_total * 2, p.barrier,
                                                                 n_scenes);
  MF_LAUNCH_CHECK();
  void* args[] = {&p, &alpha};
  MF_CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_icc_run, dim3(n_scenes * G), dim3(kIccThreads),
                                          args, 0, stream));
  return MF_OK;
}
