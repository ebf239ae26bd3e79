// average_distance (ADD / ADD-S training loss) with a matrix-free nearest-neighbour search (sm_100a).
//
// Replaces
//   morefusion/functions/loss/average_distance.py:40-85  (transform, optional NN re-indexing,
//       sqrt / mean through chainer generic ops)
//   morefusion/geometry/knn/nn.py:17-48 + cuComputeDistanceGlobal.cu:20-86 (dense R x Q squared
//       distance matrix -- 1 GB per symmetric object per step at R=500, Q=1000*500 -- then argmin)
// One CTA per predicted transform m: the P reference points T_true p live in shared memory, thread p
// transforms its query, scans the references (first minimum of the fp32 squared distance, like
// cupy.argmin over the reference kernel's matrix), and the block takes a fixed-order mean.
#include "common.cuh"

namespace mf {

constexpr int kAdThreads = 256;

__device__ __forceinline__ void tf3(const float* T, float x, float y, float z, float& ox, float& oy,
                                    float& oz) {
  ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z)), T[3]);
  oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[4], x), __fmul_rn(T[5], y)), __fmul_rn(T[6], z)), T[7]);
  oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[8], x), __fmul_rn(T[9], y)), __fmul_rn(T[10], z)), T[11]);
}

template <int NV>
__device__ __forceinline__ void ad_block_sum(float (&v)[NV], float* smem) {
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[warp * NV + k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float s = 0.f;
      for (int w = 0; w < kAdThreads / 32; ++w) s += smem[w * NV + k];
      v[k] = s;
    }
  }
}

// out[m] = mean_p || a_{idx(m,p)} - T_pred[m] p ||,  a_r = T_true p_r,  idx = p or NN
__global__ void __launch_bounds__(kAdThreads)
k_avg_dist_fwd(const float* __restrict__ points, int P, const float* __restrict__ Ttrue,
               const float* __restrict__ Tpred, int symmetric, float* __restrict__ out,
               int* __restrict__ idx_out) {
  extern __shared__ float sref[];              // [P][3] reference points in the true pose
  __shared__ float sred[kAdThreads / 32];
  const int m = blockIdx.x;
  for (int r = threadIdx.x; r < P; r += kAdThreads) {
    float ax, ay, az;
    tf3(Ttrue, points[3 * r], points[3 * r + 1], points[3 * r + 2], ax, ay, az);
    sref[3 * r] = ax; sref[3 * r + 1] = ay; sref[3 * r + 2] = az;
  }
  __syncthreads();
  const float* T = Tpred + 16 * (long long)m;
  float acc[1] = {0.f};
  // gridDim.y > 1 (few poses, many points: the metrics): the query points are split over
  // blockIdx.y and out receives one partial mean per part, out[m * gridDim.y + y]
  const int per = (P + gridDim.y - 1) / gridDim.y;
  const int p_lo = blockIdx.y * per, p_hi = min(P, p_lo + per);
  for (int p = p_lo + threadIdx.x; p < p_hi; p += kAdThreads) {
    float bx, by, bz;
    tf3(T, points[3 * p], points[3 * p + 1], points[3 * p + 2], bx, by, bz);
    int best = p;
    if (symmetric) {
      float bestd = __int_as_float(0x7f800000);
      best = 0;
      for (int r = 0; r < P; ++r) {
        float dx = __fsub_rn(sref[3 * r], bx), dy = __fsub_rn(sref[3 * r + 1], by),
              dz = __fsub_rn(sref[3 * r + 2], bz);
        float d2 = sq3(dx, dy, dz);
        if (d2 < bestd) { bestd = d2; best = r; }
      }
    }
    float dx = __fsub_rn(sref[3 * best], bx), dy = __fsub_rn(sref[3 * best + 1], by),
          dz = __fsub_rn(sref[3 * best + 2], bz);
    acc[0] += __fsqrt_rn(sq3(dx, dy, dz));
    if (idx_out) idx_out[(long long)m * P + p] = best;
  }
  ad_block_sum<1>(acc, sred);
  if (threadIdx.x == 0) out[(long long)m * gridDim.y + blockIdx.y] = acc[0] / (float)P;
}

// gTpred[m][:3,:] = sum_p g_b (x) [p,1];   part_true[m][:3,:] = sum_p (-g_b) (x) [p_idx,1]
//   g_b = d out[m] / d b = (b - a) / ||a - b|| / P * gout[m]
__global__ void __launch_bounds__(kAdThreads)
k_avg_dist_bwd(const float* __restrict__ gout, const float* __restrict__ points, int P,
               const float* __restrict__ Ttrue, const float* __restrict__ Tpred,
               const int* __restrict__ idx, float* __restrict__ gTpred,
               float* __restrict__ part_true) {
  __shared__ float sred[(kAdThreads / 32) * 24];
  const int m = blockIdx.x;
  const float* T = Tpred + 16 * (long long)m;
  const float scale = gout[m] / (float)P;
  float acc[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) acc[k] = 0.f;
  for (int p = threadIdx.x; p < P; p += kAdThreads) {
    float px = points[3 * p], py = points[3 * p + 1], pz = points[3 * p + 2];
    int r = idx ? idx[(long long)m * P + p] : p;
    float qx = points[3 * r], qy = points[3 * r + 1], qz = points[3 * r + 2];
    float ax, ay, az, bx, by, bz;
    tf3(Ttrue, qx, qy, qz, ax, ay, az);
    tf3(T, px, py, pz, bx, by, bz);
    float dx = bx - ax, dy = by - ay, dz = bz - az;
    float d = sqrtf(dx * dx + dy * dy + dz * dz);
    float g[3] = {dx / d * scale, dy / d * scale, dz / d * scale};   // 0/0 = NaN as in F.sqrt's backward
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      acc[4 * i + 0] += g[i] * px; acc[4 * i + 1] += g[i] * py;
      acc[4 * i + 2] += g[i] * pz; acc[4 * i + 3] += g[i];
      acc[12 + 4 * i + 0] -= g[i] * qx; acc[12 + 4 * i + 1] -= g[i] * qy;
      acc[12 + 4 * i + 2] -= g[i] * qz; acc[12 + 4 * i + 3] -= g[i];
    }
  }
  ad_block_sum<24>(acc, sred);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 12; ++k) gTpred[16 * (long long)m + k] = acc[k];
#pragma unroll
    for (int k = 12; k < 16; ++k) gTpred[16 * (long long)m + k] = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) part_true[12 * (long long)m + k] = acc[12 + k];
  }
}

// fixed-order reduction of the per-m partials of gTtrue
__global__ void __launch_bounds__(kAdThreads)
k_avg_dist_bwd_true(const float* __restrict__ part_true, int M, float* __restrict__ gTtrue) {
  __shared__ float sred[(kAdThreads / 32) * 12];
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int m = threadIdx.x; m < M; m += kAdThreads)
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] += part_true[12 * (long long)m + k];
  ad_block_sum<12>(acc, sred);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 12; ++k) gTtrue[k] = acc[k];
#pragma unroll
    for (int k = 12; k < 16; ++k) gTtrue[k] = 0.f;
  }
}

}  // namespace mf

using namespace mf;

extern "C" int mf_average_distance_fwd(const float* points, int n_points, const float* transform_true,
                                       const float* transforms_pred, int n_pred, int symmetric,
                                       float* out, int32_t* nn_indices, void* stream_) {
  if (n_points <= 0 || n_pred < 0) return MF_E_BADARG;
  if (n_pred == 0) return MF_OK;
  if (!points || !transform_true || !transforms_pred || !out) return MF_E_BADARG;
  size_t smem = (size_t)n_points * 12;
  if (smem > 200 * 1024) return MF_E_UNSUPPORTED;
  MF_ENSURE_DYN_SMEM(k_avg_dist_fwd, 200 * 1024);
  k_avg_dist_fwd<<<n_pred, kAdThreads, smem, (cudaStream_t)stream_>>>(
      points, n_points, transform_true, transforms_pred, symmetric, out, nn_indices);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_average_distance_fwd_parts(const float* points, int n_points,
                                             const float* transform_true,
                                             const float* transforms_pred, int n_pred,
                                             int symmetric, int n_parts, float* out_parts,
                                             int32_t* nn_indices, void* stream_) {
  if (n_points <= 0 || n_pred <= 0 || n_parts <= 0 || n_parts > 65535) return MF_E_BADARG;
  if (!points || !transform_true || !transforms_pred || !out_parts) return MF_E_BADARG;
  size_t smem = (size_t)n_points * 12;
  if (smem > 200 * 1024) return MF_E_UNSUPPORTED;
  MF_ENSURE_DYN_SMEM(k_avg_dist_fwd, 200 * 1024);
  dim3 grid((unsigned)n_pred, (unsigned)n_parts);
  k_avg_dist_fwd<<<grid, kAdThreads, smem, (cudaStream_t)stream_>>>(
      points, n_points, transform_true, transforms_pred, symmetric, out_parts, nn_indices);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_average_distance_bwd(const float* gout, const float* points, int n_points,
                                       const float* transform_true, const float* transforms_pred,
                                       int n_pred, const int32_t* nn_indices, float* g_transforms_pred,
                                       float* g_transform_true, float* workspace /*[n_pred,12]*/,
                                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_points <= 0 || n_pred < 0) return MF_E_BADARG;
  if (n_pred == 0) return MF_OK;
  if (!gout || !points || !transform_true || !transforms_pred || !g_transforms_pred ||
      !g_transform_true || !workspace)
    return MF_E_BADARG;
  k_avg_dist_bwd<<<n_pred, kAdThreads, 0, stream>>>(gout, points, n_points, transform_true,
                                                    transforms_pred, nn_indices, g_transforms_pred,
                                                    workspace);
  MF_LAUNCH_CHECK();
  k_avg_dist_bwd_true<<<1, kAdThreads, 0, stream>>>(workspace, n_pred, g_transform_true);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

// ---- the training loss calls average_distance once per object of the batch (model.py:416-431):
// the same kernels over B objects from ONE call (B small launches back to back, no per-object
// trip through the host binding).  symmetric: HOST array of B flags.  Layouts: points
// [B,n_points,3], transform_true [B,4,4], transforms_pred [B,n_pred,4,4], out [B,n_pred],
// nn_indices [B,n_pred,n_points] (rows of non-symmetric objects unused).
extern "C" int mf_average_distance_fwd_batched(const float* points, int n_points,
                                               const float* transform_true,
                                               const float* transforms_pred, int n_pred, int B,
                                               const int32_t* symmetric_host, float* out,
                                               int32_t* nn_indices, void* stream_) {
  if (B <= 0 || !symmetric_host) return MF_E_BADARG;
  for (int b = 0; b < B; ++b) {
    int32_t* idx = nn_indices ? nn_indices + (size_t)b * n_pred * n_points : nullptr;
    if (symmetric_host[b] && !idx) return MF_E_BADARG;
    int rc = mf_average_distance_fwd(points + (size_t)b * n_points * 3, n_points,
                                     transform_true + (size_t)b * 16,
                                     transforms_pred + (size_t)b * n_pred * 16, n_pred,
                                     symmetric_host[b] ? 1 : 0, out + (size_t)b * n_pred,
                                     symmetric_host[b] ? idx : nullptr, stream_);
    if (rc) return rc;
  }
  return MF_OK;
}

extern "C" int mf_average_distance_bwd_batched(const float* gout, const float* points, int n_points,
                                               const float* transform_true,
                                               const float* transforms_pred, int n_pred, int B,
                                               const int32_t* symmetric_host,
                                               const int32_t* nn_indices, float* g_transforms_pred,
                                               float* g_transform_true,
                                               float* workspace /*[B,n_pred,12]*/, void* stream_) {
  if (B <= 0 || !symmetric_host) return MF_E_BADARG;
  for (int b = 0; b < B; ++b) {
    const int32_t* idx = (symmetric_host[b] && nn_indices)
                             ? nn_indices + (size_t)b * n_pred * n_points : nullptr;
    if (symmetric_host[b] && !idx) return MF_E_BADARG;
    int rc = mf_average_distance_bwd(gout + (size_t)b * n_pred, points + (size_t)b * n_points * 3,
                                     n_points, transform_true + (size_t)b * 16,
                                     transforms_pred + (size_t)b * n_pred * 16, n_pred, idx,
                                     g_transforms_pred + (size_t)b * n_pred * 16,
                                     g_transform_true + (size_t)b * 16,
                                     workspace + (size_t)b * n_pred * 12, stream_);
    if (rc) return rc;
  }
  return MF_OK;
}

