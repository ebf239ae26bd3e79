// Backward-pass kernels of the singleview_3d pose model around the tcgen05 GEMMs (gemm_train.cu,
// conv3d_tc.cu): everything chainer's reverse pass does for
//   morefusion/contrib/singleview_3d/models/model.py:93-141 (_extract), :239-273 (heads, pose)
// that is not a matrix product.  Activations are the channels-last bf16 buffers of the forward
// (cnn.cu); gradients that feed a GEMM are bf16, accumulators and parameter gradients fp32.
//
//   k_head4_bwd      pose epilogue (normalise / voxel->camera frame / sigmoid) + last head layer:
//                    only the rows of each object's class carry gradient (model.py:266-269)
//   k_relu_mask      dZ = dY * (Y > 0)                              (F.relu backward)
//   k_colsum         bias gradients
//   k_interp_bwd     trilinear scatter into a channels-last fp32 grid (interpolate_voxel_grid.py
//                    :230-262, 16-byte vector atomics)
//   k_mask_pack      dY = bf16(dgrid * (act > 0)) into the zero-padded layout the conv dgrad /
//                    wgrad TMA boxes read, + the conv bias gradient
//   k_vox_bwd        average_voxelization_3d backward (average_voxelization_3d.py:170-206) from
//                    the parity-ordered conv3 input gradient, + the direct feat2 gradient
//   k_point_mlp_bwd  conv1/conv2 rgb+pcd backward (weights, biases, d values)
//   k_adam           fused 1/world unscale + Chainer-form Adam on a flat fp32 buffer
#include <cuda_bf16.h>

#include "common.cuh"
#include "cnn.cuh"

namespace mf {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float bf(const bf16 x) { return __bfloat162float(x); }

// ------------------------------------------------------------------ head layer 4 + pose epilogue
// g_rot [NP,4] (w.r.t. the NORMALISED quaternion), g_trans [NP,3] (camera frame), g_conf [NP]
// (after the sigmoid); raw8 [NP,8] = the 4+3+1 selected outputs of conv4_{rot,trans,conf} saved
// by the forward.  d raw:  rot: r = o / (|o| + eps):  d o = g / (n + eps) - o (o.g) / (n (n+eps)^2)
//                        trans: d = g * pitch;  conf: d = g * s (1 - s).
constexpr int kH4bPts = 16;
__global__ void __launch_bounds__(256)
k_head4_bwd(const float* __restrict__ g_rot, const float* __restrict__ g_trans,
            const float* __restrict__ g_conf, const float* __restrict__ raw8,
            const bf16* __restrict__ hd3, int ld,
            const bf16* __restrict__ w_rot, const bf16* __restrict__ w_trans,
            const bf16* __restrict__ w_conf, const int* __restrict__ class_id,
            const float* __restrict__ pitch, int B, int P, int nfg,
            bf16* __restrict__ dhd3,
            float* __restrict__ dw_rot, float* __restrict__ db_rot,
            float* __restrict__ dw_trans, float* __restrict__ db_trans,
            float* __restrict__ dw_conf, float* __restrict__ db_conf) {
  __shared__ float w[8][128];
  __shared__ float d8[kH4bPts][8];
  __shared__ float hs[kH4bPts][3][128 + 1];
  const int b = blockIdx.y;
  int fg = class_id[b] - 1;
  if (fg < 0) fg += nfg;
  fg = min(max(fg, 0), nfg - 1);
  const int tid = threadIdx.x;
  for (int e = tid; e < 8 * 128; e += 256) {
    const int r = e >> 7, k = e & 127;
    const bf16* src = r < 4 ? w_rot + (size_t)(fg * 4 + r) * 128
                    : r < 7 ? w_trans + (size_t)(fg * 3 + r - 4) * 128
                            : w_conf + (size_t)fg * 128;
    w[r][k] = bf(src[k]);
  }
  const int p0 = blockIdx.x * kH4bPts;
  // ---- d raw of this CTA's points
  if (tid < kH4bPts) {
    const int p = p0 + tid;
    float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (p < P) {
      const long long n = (long long)b * P + p;
      const float* o = raw8 + n * 8;
      const float nr = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
      const float ne = nr + 1e-5f;
      const float* g = g_rot + n * 4;
      const float og = o[0] * g[0] + o[1] * g[1] + o[2] * g[2] + o[3] * g[3];
      const float k2 = nr > 0.f ? og / (nr * ne * ne) : 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) d[k] = g[k] / ne - o[k] * k2;
      const float pt = pitch[b];
#pragma unroll
      for (int k = 0; k < 3; ++k) d[4 + k] = g_trans[n * 3 + k] * pt;
      const float s = 1.f / (1.f + expf(-o[7]));
      d[7] = g_conf[n] * s * (1.f - s);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) d8[tid][k] = d[k];
  }
  // ---- this CTA's hd3 rows (three 128-wide segments) as fp32
  for (int e = tid; e < kH4bPts * 384; e += 256) {
    const int pt = e / 384, c = e - pt * 384;
    const int p = p0 + pt;
    hs[pt][c >> 7][c & 127] = p < P ? bf(hd3[((long long)b * P + p) * ld + c]) : 0.f;
  }
  __syncthreads();
  // ---- d hd3 = W^T d raw, masked by hd3 > 0 (ReLU of layer 3)
  for (int e = tid; e < kH4bPts * 384; e += 256) {
    const int pt = e / 384, c = e - pt * 384, seg = c >> 7, k = c & 127;
    const int p = p0 + pt;
    if (p >= P) continue;
    float a = 0.f;
    if (seg == 0) a = d8[pt][0] * w[0][k] + d8[pt][1] * w[1][k] + d8[pt][2] * w[2][k] + d8[pt][3] * w[3][k];
    else if (seg == 1) a = d8[pt][4] * w[4][k] + d8[pt][5] * w[5][k] + d8[pt][6] * w[6][k];
    else a = d8[pt][7] * w[7][k];
    dhd3[((long long)b * P + p) * ld + c] = __float2bfloat16(hs[pt][seg][k] > 0.f ? a : 0.f);
  }
  // ---- d W4[row r][k] += sum_points d raw[r] * hd3[seg(r)][k];  d b4[r] += sum d raw[r]
  for (int e = tid; e < 8 * 128; e += 256) {
    const int r = e >> 7, k = e & 127, seg = r < 4 ? 0 : r < 7 ? 1 : 2;
    float a = 0.f;
    for (int pt = 0; pt < kH4bPts; ++pt) a += d8[pt][r] * hs[pt][seg][k];
    float* dst = r < 4 ? dw_rot + (size_t)(fg * 4 + r) * 128
               : r < 7 ? dw_trans + (size_t)(fg * 3 + r - 4) * 128
                       : dw_conf + (size_t)fg * 128;
    atomicAdd(dst + k, a);
  }
  if (tid < 8) {
    float a = 0.f;
    for (int pt = 0; pt < kH4bPts; ++pt) a += d8[pt][tid];
    float* dst = tid < 4 ? db_rot + fg * 4 + tid : tid < 7 ? db_trans + fg * 3 + tid - 4 : db_conf + fg;
    atomicAdd(dst, a);
  }
}

// ------------------------------------------------------------------ elementwise / reductions
// x[m, n] = act[m, n] > 0 ? x[m, n] : 0     (8 bf16 per thread)
__global__ void k_relu_mask(bf16* __restrict__ x, long long ldx, const bf16* __restrict__ act,
                            long long lda, long long M, int N8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N8) return;
  const long long m = i / N8;
  const int c = (int)(i - m * N8) * 8;
  uint4 xv = *reinterpret_cast<const uint4*>(x + m * ldx + c);
  const uint4 av = *reinterpret_cast<const uint4*>(act + m * lda + c);
  bf16* xe = reinterpret_cast<bf16*>(&xv);
  const bf16* ae = reinterpret_cast<const bf16*>(&av);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (!(bf(ae[k]) > 0.f)) xe[k] = __float2bfloat16(0.f);
  *reinterpret_cast<uint4*>(x + m * ldx + c) = xv;
}

// out[n] += sum_m x[m, n]; CTA = 64 columns x 256 rows
__global__ void __launch_bounds__(256)
k_colsum(const bf16* __restrict__ x, long long ldx, long long M, int N, float* __restrict__ out) {
  __shared__ float part[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63), rg = threadIdx.x >> 6;
  const long long m0 = (long long)blockIdx.y * 256;
  float a = 0.f;
  if (c < N)
    for (long long m = m0 + rg; m < min(m0 + 256, M); m += 4) a += bf(x[m * ldx + c]);
  part[rg][threadIdx.x & 63] = a;
  __syncthreads();
  if (threadIdx.x < 64 && c < N)
    atomicAdd(out + c, part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] +
                           part[3][threadIdx.x]);
}

// ------------------------------------------------------------------ trilinear scatter
// dgrid[b, corner, c] += w_corner * g[n, c]  (channels-last fp32 grid, 4 channels per thread)
__global__ void k_interp_bwd(const bf16* __restrict__ g, long long ldg, int col_off,
                             const float* __restrict__ points /*[B,3,P]*/, int B, int P, int C,
                             int D, float divisor, float* __restrict__ dgrid) {
  const int C4 = C >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * P * C4) return;
  const long long n = i / C4;
  const int c = (int)(i - n * C4) * 4;
  const long long b = n / P, pp = n % P;
  const float x = __fdiv_rn(points[(b * 3 + 0) * P + pp], divisor);
  const float y = __fdiv_rn(points[(b * 3 + 1) * P + pp], divisor);
  const float z = __fdiv_rn(points[(b * 3 + 2) * P + pp], divisor);
  const int x0 = static_cast<int>(x), y0 = static_cast<int>(y), z0 = static_cast<int>(z);
  const float lx = x - (float)x0, ly = y - (float)y0, lz = z - (float)z0;
  const float hx = 1.f - lx, hy = 1.f - ly, hz = 1.f - lz;
  const float w[8] = {hx * hy * hz, lx * hy * hz, hx * ly * hz, hx * hy * lz,
                      lx * ly * hz, hx * ly * lz, lx * hy * lz, lx * ly * lz};
  const int dx[8] = {0, 1, 0, 0, 1, 0, 1, 1};
  const int dy[8] = {0, 0, 1, 0, 1, 1, 0, 1};
  const int dz[8] = {0, 0, 0, 1, 0, 1, 1, 1};
  const uint2 gv = *reinterpret_cast<const uint2*>(g + n * ldg + col_off + c);
  const bf16* ge = reinterpret_cast<const bf16*>(&gv);
  const float g0 = bf(ge[0]), g1 = bf(ge[1]), g2 = bf(ge[2]), g3 = bf(ge[3]);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ix = x0 + dx[j], iy = y0 + dy[j], iz = z0 + dz[j];
    if (ix < 0 || ix >= D || iy < 0 || iy >= D || iz < 0 || iz >= D) continue;
    float4* dst = reinterpret_cast<float4*>(dgrid + ((((b * D + ix) * D + iy) * D + iz) * (long long)C + c));
    atomicAdd(dst, make_float4(w[j] * g0, w[j] * g1, w[j] * g2, w[j] * g3));
  }
}

// ------------------------------------------------------------------ ReLU mask + pad + bias grad
// out[b, 1+x, 1+y, 1+z, c] = act(b,x,y,z,c) > 0 ? bf16(dgrid[b,x,y,z,c]) : 0;  dbias[c] += sum.
// act is the forward activation: row-major [B,D^3,C] (act_s2d = 0) or the s2d input buffer of the
// next convolution (act_s2d = 1, cnn.cu layout).  One thread = 8 channels of one voxel.
__global__ void __launch_bounds__(256)
k_mask_pack(const float* __restrict__ dgrid, const bf16* __restrict__ act, int act_s2d, int B,
            int D, int C, bf16* __restrict__ out, float* __restrict__ dbias) {
  const int C8 = C >> 3;
  // thread -> (voxel slot, channel block): consecutive threads walk the channel blocks of a
  // voxel, a CTA covers 256 / C8 voxels per step; channel block fixed per thread (bias partials)
  const int cb = threadIdx.x % C8, vs = threadIdx.x / C8, vper = blockDim.x / C8;
  const long long NV = (long long)B * D * D * D;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int J = D / 2 + 1, Dp = D + 2;
  for (long long v = (long long)blockIdx.x * vper + vs; v < NV; v += (long long)gridDim.x * vper) {
    if (vs >= vper) break;
    const int z = (int)(v % D), y = (int)((v / D) % D), x = (int)((v / ((long long)D * D)) % D);
    const long long b = v / ((long long)D * D * D);
    long long aoff;
    if (act_s2d) {
      const int pd = x + 1, ph = y + 1, pw = z + 1;
      const int r = ((pd & 1) << 2) | ((ph & 1) << 1) | (pw & 1);
      aoff = ((((b * J + (pd >> 1)) * J + (ph >> 1)) * J + (pw >> 1)) * 8LL + r) * C + cb * 8;
    } else {
      aoff = v * C + cb * 8;
    }
    const uint4 av = *reinterpret_cast<const uint4*>(act + aoff);
    const bf16* ae = reinterpret_cast<const bf16*>(&av);
    const float4 d0 = *reinterpret_cast<const float4*>(dgrid + v * C + cb * 8);
    const float4 d1 = *reinterpret_cast<const float4*>(dgrid + v * C + cb * 8 + 4);
    const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    uint4 ov;
    bf16* oe = reinterpret_cast<bf16*>(&ov);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t = bf(ae[k]) > 0.f ? d[k] : 0.f;
      oe[k] = __float2bfloat16(t);
      acc[k] += bf(oe[k]);
    }
    *reinterpret_cast<uint4*>(out + ((((b * Dp + x + 1) * Dp + y + 1) * Dp + z + 1) * (long long)C + cb * 8)) = ov;
  }
  // bias gradient: the CTA's threads that share a channel block are summed in shared memory
  // first (one atomic per CTA and channel instead of one per thread)
  __shared__ float sb[256 * 8];
#pragma unroll
  for (int k = 0; k < 8; ++k) sb[threadIdx.x * 8 + k] = (vs < vper) ? acc[k] : 0.f;
  __syncthreads();
  if (dbias && threadIdx.x < C8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a = 0.f;
      for (int j = 0; j < vper; ++j) a += sb[(j * C8 + threadIdx.x) * 8 + k];
      atomicAdd(dbias + threadIdx.x * 8 + k, a);
    }
  }
}

// ------------------------------------------------------------------ voxelisation backward
// dfeat2[n, c] = g_direct[n, 72 + c] + dX3(voxel of n, c) / count(voxel)      c < C = 144
// dX3 is the conv3 input gradient in parity order: [8][B * (D/2)^3][ldp] bf16, row
// (b, o) of parity r = input voxel 2 o + 1 - r.  One warp per point; the count is found by
// scanning the object's keys (P = 1000: 32 steps).
__global__ void __launch_bounds__(256)
k_vox_bwd(const bf16* __restrict__ dx3, long long ldp, const int* __restrict__ keys, int B, int P,
          int C, int D, const bf16* __restrict__ g_direct, long long ldg, int col_off,
          float* __restrict__ dfeat2) {
  const int n = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (n >= B * P) return;
  const int raw = keys[n];
  const int key = raw < 0 ? -1 : (raw & MF_S2D_KEY_MASK);
  const int b = n / P, lo = b * P, hi = lo + P;
  int count = raw >= 0 ? 1 : 0;
  if (raw >= 0 && (raw & MF_S2D_DUP_BIT)) {       // shared voxel: count its points
    count = 0;
    for (int j0 = lo; j0 < hi; j0 += 32) {
      const int j = j0 + lane;
      const bool m = j < hi && keys[j] == raw;
      count += __popc(__ballot_sync(0xffffffffu, m));
    }
  }
  const int V = D * D * D, Do = D / 2;
  long long row = 0;
  if (key >= 0) {
    const int flat = key - b * V;
    const int z = flat % D, y = (flat / D) % D, x = flat / (D * D);
    const int rd = 1 - (x & 1), rh = 1 - (y & 1), rw = 1 - (z & 1);     // x = 2 o + 1 - r
    const int od = (x - 1 + rd) >> 1, oh = (y - 1 + rh) >> 1, ow = (z - 1 + rw) >> 1;
    const int r = (rd << 2) | (rh << 1) | rw;
    row = (long long)r * B * Do * Do * Do + (((long long)b * Do + od) * Do + oh) * Do + ow;
  }
  const float inv = count > 0 ? 1.f / (float)count : 0.f;
  for (int c = lane; c < C; c += 32) {
    float v = bf(g_direct[(long long)n * ldg + col_off + c]);
    if (key >= 0) v += bf(dx3[row * ldp + c]) * inv;
    dfeat2[(long long)n * C + c] = v;
  }
}

// ------------------------------------------------------------------ per-point MLP backward
// Forward (k_point_mlp): x = [values(32) | center - point (3)], h1 = relu(W1 x + b1) (64 | 8),
// h2 = relu(W2 h1 + b2) (128 | 16).  Inputs: g2in [N,144] fp32 = total gradient of h2 (feat2),
// dfeat (bf16) columns 0..71 = direct gradient of h1 (feat1 goes into the concat feature too),
// feat (bf16) columns 0..71 = h1, feat2 fp32 = h2.
// 64 points per CTA; weight gradients are reduced over the CTA's points in registers, then one
// atomicAdd per entry.
constexpr int kMbPts = 64;
constexpr int kMbSmemFloats = 128 * 64 + 64 * 32 + 16 * 8 + 8 * 3 + kMbPts * (145 + 73 + 73 + 36);
__global__ void __launch_bounds__(256)
k_point_mlp_bwd(const float* __restrict__ values, const float* __restrict__ points,
                const bf16* __restrict__ feat, long long ldf, const bf16* __restrict__ dfeat,
                long long ldd, const float* __restrict__ feat2, const float* __restrict__ g2in,
                const float* __restrict__ w1r /*[64,32]*/, const float* __restrict__ w1p /*[8,3]*/,
                const float* __restrict__ w2r /*[128,64]*/, const float* __restrict__ w2p /*[16,8]*/,
                int B, int P, float center,
                float* __restrict__ dw1r, float* __restrict__ db1r, float* __restrict__ dw1p,
                float* __restrict__ db1p, float* __restrict__ dw2r, float* __restrict__ db2r,
                float* __restrict__ dw2p, float* __restrict__ db2p,
                float* __restrict__ dvalues /*[B,32,P] or null*/) {
  extern __shared__ float sm[];
  float* sw2r = sm;                         // [128][64]
  float* sw1r = sw2r + 128 * 64;            // [64][32]
  float* sw2p = sw1r + 64 * 32;             // [16][8]
  float* sw1p = sw2p + 16 * 8;              // [8][3]
  float* g2 = sw1p + 8 * 3;                 // [kMbPts][145]  (masked)
  float* g1 = g2 + kMbPts * 145;            // [kMbPts][73]   (masked)
  float* h1 = g1 + kMbPts * 73;             // [kMbPts][73]
  float* xin = h1 + kMbPts * 73;            // [kMbPts][36]
  const int tid = threadIdx.x;
  for (int e = tid; e < 128 * 64; e += 256) sw2r[e] = w2r[e];
  for (int e = tid; e < 64 * 32; e += 256) sw1r[e] = w1r[e];
  for (int e = tid; e < 16 * 8; e += 256) sw2p[e] = w2p[e];
  for (int e = tid; e < 8 * 3; e += 256) sw1p[e] = w1p[e];
  const long long n0 = (long long)blockIdx.x * kMbPts, NP = (long long)B * P;
  for (int e = tid; e < kMbPts * 35; e += 256) {
    const int c = e / kMbPts, pt = e % kMbPts;
    const long long n = n0 + pt;
    float v = 0.f;
    if (n < NP) {
      const long long b = n / P, p = n % P;
      v = (c < 32) ? values[(b * 32 + c) * P + p] : (center - points[(b * 3 + (c - 32)) * P + p]);
    }
    xin[pt * 36 + c] = v;
  }
  for (int e = tid; e < kMbPts * 144; e += 256) {
    const int pt = e / 144, c = e - pt * 144;
    const long long n = n0 + pt;
    float v = 0.f;
    if (n < NP && feat2[n * 144 + c] > 0.f) v = g2in[n * 144 + c];
    g2[pt * 145 + c] = v;
  }
  for (int e = tid; e < kMbPts * 72; e += 256) {
    const int pt = e / 72, c = e - pt * 72;
    const long long n = n0 + pt;
    h1[pt * 73 + c] = n < NP ? bf(feat[n * ldf + c]) : 0.f;
  }
  __syncthreads();
  // g1 = (direct + W2^T g2) masked by h1 > 0
  for (int e = tid; e < kMbPts * 72; e += 256) {
    const int pt = e / 72, c = e - pt * 72;
    const long long n = n0 + pt;
    float a = 0.f;
    if (n < NP && h1[pt * 73 + c] > 0.f) {
      a = bf(dfeat[n * ldd + c]);
      if (c < 64) {
        for (int o = 0; o < 128; ++o) a += sw2r[o * 64 + c] * g2[pt * 145 + o];
      } else {
        for (int o = 0; o < 16; ++o) a += sw2p[o * 8 + (c - 64)] * g2[pt * 145 + 128 + o];
      }
    }
    g1[pt * 73 + c] = a;
  }
  __syncthreads();
  // ---- weight / bias gradients: sums over the CTA's points
  for (int e = tid; e < 128 * 64; e += 256) {
    const int o = e >> 6, k = e & 63;
    float a = 0.f;
    for (int pt = 0; pt < kMbPts; ++pt) a += g2[pt * 145 + o] * h1[pt * 73 + k];
    atomicAdd(dw2r + e, a);
  }
  for (int e = tid; e < 16 * 8; e += 256) {
    const int o = e >> 3, k = e & 7;
    float a = 0.f;
    for (int pt = 0; pt < kMbPts; ++pt) a += g2[pt * 145 + 128 + o] * h1[pt * 73 + 64 + k];
    atomicAdd(dw2p + e, a);
  }
  for (int e = tid; e < 64 * 32; e += 256) {
    const int o = e >> 5, k = e & 31;
    float a = 0.f;
    for (int pt = 0; pt < kMbPts; ++pt) a += g1[pt * 73 + o] * xin[pt * 36 + k];
    atomicAdd(dw1r + e, a);
  }
  for (int e = tid; e < 8 * 3; e += 256) {
    const int o = e / 3, k = e - o * 3;
    float a = 0.f;
    for (int pt = 0; pt < kMbPts; ++pt) a += g1[pt * 73 + 64 + o] * xin[pt * 36 + 32 + k];
    atomicAdd(dw1p + e, a);
  }
  for (int e = tid; e < 144 + 72; e += 256) {
    float a = 0.f;
    if (e < 144) {
      for (int pt = 0; pt < kMbPts; ++pt) a += g2[pt * 145 + e];
      atomicAdd(e < 128 ? db2r + e : db2p + (e - 128), a);
    } else {
      const int c = e - 144;
      for (int pt = 0; pt < kMbPts; ++pt) a += g1[pt * 73 + c];
      atomicAdd(c < 64 ? db1r + c : db1p + (c - 64), a);
    }
  }
  // ---- d values = W1r^T g1_rgb   (gradient into the 2-D feature extractor)
  if (dvalues)
    for (int e = tid; e < kMbPts * 32; e += 256) {
      const int k = e / kMbPts, pt = e % kMbPts;
      const long long n = n0 + pt;
      if (n >= NP) continue;
      float a = 0.f;
      for (int o = 0; o < 64; ++o) a += sw1r[o * 32 + k] * g1[pt * 73 + o];
      const long long b = n / P, p = n % P;
      dvalues[(b * 32 + k) * P + p] = a;
    }
}

// ------------------------------------------------------------------ fused unscale + Chainer Adam
// g <- g * gscale (1 / world size after the sum all-reduce);  m += (1-b1)(g-m);
// v += (1-b2)(g^2-v);  p -= eta * alpha_t * m / (sqrt(v) + eps)        (chainer optimizers/adam.py)
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, long long n, float alpha_t, float omb1, float omb2,
                       float eps, float eta, float gscale) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 4 <= n) {
    float4 pv = *reinterpret_cast<float4*>(p + i), mv = *reinterpret_cast<float4*>(m + i),
           vv = *reinterpret_cast<float4*>(v + i);
    const float4 gv = *reinterpret_cast<const float4*>(g + i);
    float* pe = &pv.x; float* me = &mv.x; float* ve = &vv.x;
    const float* ge = &gv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gg = ge[k] * gscale;
      me[k] += omb1 * (gg - me[k]);
      ve[k] += omb2 * (gg * gg - ve[k]);
      pe[k] -= eta * (alpha_t * me[k] / (sqrtf(ve[k]) + eps));
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
  } else {
    for (long long j = i; j < n; ++j) {
      const float gg = g[j] * gscale;
      m[j] += omb1 * (gg - m[j]);
      v[j] += omb2 * (gg * gg - v[j]);
      p[j] -= eta * (alpha_t * m[j] / (sqrtf(v[j]) + eps));
    }
  }
}

}  // namespace mf

using namespace mf;

extern "C" int mf_train_head4_bwd(const float* g_rot, const float* g_trans, const float* g_conf,
                                  const float* raw8, const void* hd3, int ld, const void* w_rot,
                                  const void* w_trans, const void* w_conf, const int32_t* class_id,
                                  const float* pitch, int B, int P, int nfg, void* dhd3,
                                  float* dw_rot, float* db_rot, float* dw_trans, float* db_trans,
                                  float* dw_conf, float* db_conf, void* stream_) {
  if (B <= 0 || P <= 0 || nfg <= 0 || ld < 384 || B > 65535) return MF_E_BADARG;
  if (!g_rot || !g_trans || !g_conf || !raw8 || !hd3 || !w_rot || !w_trans || !w_conf ||
      !class_id || !pitch || !dhd3 || !dw_rot || !db_rot || !dw_trans || !db_trans || !dw_conf ||
      !db_conf)
    return MF_E_BADARG;
  dim3 grid((unsigned)div_up(P, kH4bPts), (unsigned)B);
  k_head4_bwd<<<grid, 256, 0, (cudaStream_t)stream_>>>(
      g_rot, g_trans, g_conf, raw8, (const bf16*)hd3, ld, (const bf16*)w_rot, (const bf16*)w_trans,
      (const bf16*)w_conf, class_id, pitch, B, P, nfg, (bf16*)dhd3, dw_rot, db_rot, dw_trans,
      db_trans, dw_conf, db_conf);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_train_relu_mask(void* x, long long ldx, const void* act, long long lda,
                                  long long M, int N, void* stream_) {
  if (!x || !act || M <= 0 || N <= 0 || (N & 7) || (ldx & 7) || (lda & 7)) return MF_E_BADARG;
  k_relu_mask<<<div_up(M * (N / 8), 256), 256, 0, (cudaStream_t)stream_>>>(
      (bf16*)x, ldx, (const bf16*)act, lda, M, N / 8);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_train_colsum(const void* x, long long ldx, long long M, int N, float* out,
                               void* stream_) {
  if (!x || !out || M <= 0 || N <= 0) return MF_E_BADARG;
  dim3 grid((unsigned)div_up(N, 64), (unsigned)div_up(M, 256));
  k_colsum<<<grid, 256, 0, (cudaStream_t)stream_>>>((const bf16*)x, ldx, M, N, out);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_train_interp_bwd(const void* g, long long ldg, int col_off, const float* points,
                                   int B, int P, int C, int D, float divisor, float* dgrid,
                                   void* stream_) {
  if (!g || !points || !dgrid || B <= 0 || P <= 0 || C <= 0 || (C & 3) || (col_off & 3) ||
      (ldg & 3) || D <= 0)
    return MF_E_BADARG;
  k_interp_bwd<<<div_up((long long)B * P * (C / 4), 256), 256, 0, (cudaStream_t)stream_>>>(
      (const bf16*)g, ldg, col_off, points, B, P, C, D, divisor, dgrid);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_train_mask_pack(const float* dgrid, const void* act, int act_s2d, int B, int D,
                                  int C, void* out, float* dbias, void* stream_) {
  if (!dgrid || !act || !out || B <= 0 || D <= 0 || (D & 1) || C <= 0 || (C & 7) || C / 8 > 256)
    return MF_E_BADARG;
  k_mask_pack<<<592, 256, 0, (cudaStream_t)stream_>>>(dgrid, (const bf16*)act, act_s2d, B, D, C,
                                                       (bf16*)out, dbias);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_train_vox_bwd(const void* dx3, long long ldp, const int32_t* keys, int B, int P,
                                int C, int D, const void* g_direct, long long ldg, int col_off,
                                float* dfeat2, void* stream_) {
  if (!dx3 || !keys || !g_direct || !dfeat2 || B <= 0 || P <= 0 || C <= 0 || D <= 0 || (D & 1))
    return MF_E_BADARG;
  k_vox_bwd<<<div_up((long long)B * P * 32, 256), 256, 0, (cudaStream_t)stream_>>>(
      (const bf16*)dx3, ldp, keys, B, P, C, D, (const bf16*)g_direct, ldg, col_off, dfeat2);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_train_point_mlp_bwd(const float* values, const float* points, const void* feat,
                                      long long ldf, const void* dfeat, long long ldd,
                                      const float* feat2, const float* g2, const float* w1r,
                                      const float* w1p, const float* w2r, const float* w2p, int B,
                                      int P, float center, float* dw1r, float* db1r, float* dw1p,
                                      float* db1p, float* dw2r, float* db2r, float* dw2p,
                                      float* db2p, float* dvalues, void* stream_) {
  if (B <= 0 || P <= 0 || !values || !points || !feat || !dfeat || !feat2 || !g2 || !w1r || !w1p ||
      !w2r || !w2p || !dw1r || !db1r || !dw1p || !db1p || !dw2r || !db2r || !dw2p || !db2p)
    return MF_E_BADARG;
  MF_ENSURE_DYN_SMEM(k_point_mlp_bwd, kMbSmemFloats * 4);
  k_point_mlp_bwd<<<div_up((long long)B * P, kMbPts), 256, kMbSmemFloats * 4, (cudaStream_t)stream_>>>(
      values, points, (const bf16*)feat, ldf, (const bf16*)dfeat, ldd, feat2, g2, w1r, w1p, w2r,
      w2p, B, P, center, dw1r, db1r, dw1p, db1p, dw2r, db2r, dw2p, db2p, dvalues);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_train_adam(float* params, const float* grads, float* m, float* v, int64_t n,
                             float alpha_t, double beta1, double beta2, double eps, double eta,
                             float grad_scale, void* stream_) {
  if (!params || !grads || !m || !v || n <= 0) return MF_E_BADARG;
  if (((uintptr_t)params & 15) || ((uintptr_t)grads & 15) || ((uintptr_t)m & 15) || ((uintptr_t)v & 15))
    return MF_E_UNSUPPORTED;
  k_adam<<<div_up((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream_>>>(
      params, grads, m, v, n, alpha_t, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
      (float)eta, grad_scale);
  MF_LAUNCH_CHECK();
  return MF_OK;
}
