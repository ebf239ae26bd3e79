// Per-frame front end of the pose pipeline: the NumPy helpers that run on the
// host in the reference before every network call,
//   morefusion/geometry/pointcloud_from_depth.py:4-26   depth image -> camera-frame point image
//   morefusion/geometry/masks_to_bboxes.py:4-38         instance masks -> bounding boxes
// as single CUDA launches on device-resident images (no host round trip per frame).
#include "common.cuh"

namespace mf {

// pcd[r, c] = (z (c - cx) / fx, z (r - cy) / fy, z), NaN where depth is NaN; "euclidean" depth is
// the ray length: the point is rescaled by z / |p|  (pointcloud_from_depth.py:19-25)
__global__ void k_pointcloud_from_depth(const float* __restrict__ depth, int H, int W, float fx,
                                        float fy, float cx, float cy, int euclidean,
                                        float* __restrict__ pcd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int r = i / W, c = i - r * W;
  const float z = depth[i];
  float x = __fdiv_rn(__fmul_rn(z, __fsub_rn((float)c, cx)), fx);
  float y = __fdiv_rn(__fmul_rn(z, __fsub_rn((float)r, cy)), fy);
  float zz = z;
  if (euclidean) {
    const float n = sqrtf(x * x + y * y + z * z);
    const float s = z / n;
    x *= s; y *= s; zz *= s;
  }
  pcd[3 * i] = x;                 // NaN depth propagates to all three coordinates
  pcd[3 * i + 1] = y;
  pcd[3 * i + 2] = zz;
}

// one CTA per mask: (y1, x1, y2, x2) with exclusive upper bounds, zeros for an empty mask
__global__ void __launch_bounds__(256)
k_masks_to_bboxes(const unsigned char* __restrict__ masks, int H, int W, int* __restrict__ bboxes) {
  __shared__ int s[4];
  if (threadIdx.x == 0) { s[0] = 0x7fffffff; s[1] = 0x7fffffff; s[2] = -1; s[3] = -1; }
  __syncthreads();
  const unsigned char* m = masks + (size_t)blockIdx.x * H * W;
  int y1 = 0x7fffffff, x1 = 0x7fffffff, y2 = -1, x2 = -1;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x)
    if (m[i]) {
      const int r = i / W, c = i - r * W;
      y1 = min(y1, r); x1 = min(x1, c); y2 = max(y2, r); x2 = max(x2, c);
    }
  if (y2 >= 0) {
    atomicMin(&s[0], y1); atomicMin(&s[1], x1); atomicMax(&s[2], y2); atomicMax(&s[3], x2);
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const bool any = s[2] >= 0;
    bboxes[blockIdx.x * 4 + threadIdx.x] = any ? (threadIdx.x < 2 ? s[threadIdx.x] : s[threadIdx.x] + 1) : 0;
  }
}

}  // namespace mf

using namespace mf;

extern "C" int mf_pointcloud_from_depth(const float* depth, int H, int W, float fx, float fy,
                                        float cx, float cy, int euclidean, float* pcd, void* stream_) {
  if (!depth || !pcd || H <= 0 || W <= 0 || (long long)H * W >= (1LL << 31)) return MF_E_BADARG;
  k_pointcloud_from_depth<<<div_up((long long)H * W, 256), 256, 0, (cudaStream_t)stream_>>>(
      depth, H, W, fx, fy, cx, cy, euclidean, pcd);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_masks_to_bboxes(const uint8_t* masks, int N, int H, int W, int32_t* bboxes,
                                  void* stream_) {
  if (!masks || !bboxes || N <= 0 || H <= 0 || W <= 0) return MF_E_BADARG;
  k_masks_to_bboxes<<<N, 256, 0, (cudaStream_t)stream_>>>(masks, H, W, bboxes);
  MF_LAUNCH_CHECK();
  return MF_OK;
}
