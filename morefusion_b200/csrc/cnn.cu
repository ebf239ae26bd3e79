// 3D-CNN section of the singleview_3d pose model (sm_100a): SIMT stages.
//
// Replaces the chainer/cuDNN layers of
//   morefusion/contrib/singleview_3d/models/model.py:93-141 (_extract), :239-273 (heads, pose)
// with channels-last bf16 activations and fp32 accumulation:
//   * per-point 1x1 "Conv1D" stacks (model.py:62-66,97-108)           -> k_point_mlp
//   * conv1_occ / conv2_occ thin 3^3 stencils (model.py:69-70,114-125) -> k_occ_conv1/2
//   * space-to-depth packing of the 32^3 x160 grid for the k4 s2 convs -> k_pack_s2d
//   * GEMM (implicit-GEMM conv3/conv4 over the s2d layout, and the head Conv1Ds);
//     this file holds the SIMT version, conv3d_tc.cu the tcgen05 one      -> k_gemm_simt
//   * trilinear gather from channels-last bf16 grids (model.py:130-139) -> k_interp_cl
//   * class selection / normalise / voxel->camera frame (model.py:256-273) -> k_pose
//
// Space-to-depth: a k=4, stride-2, pad-1 convolution over x equals a k=2, stride-1
// convolution over y[j, r] = xpad[2j + r] (r = parity bits per axis, xpad = x padded by 1):
//   out[o] = sum_{a in {0,1}^3} sum_{r, ci} W[2a+r, ci] * y[o + a, (r, ci)]
// so the A operand of the implicit GEMM for one K block is a dense box of y.
#include <cuda_bf16.h>

#include "common.cuh"
#include "cnn.cuh"

namespace mf {

using bf16 = __nv_bfloat16;

// ------------------------------------------------------------------ per-point MLP
// 32 points per CTA (250 CTAs at 8 x 1000 points, 4 resident per SM so that one CTA's loads hide
// under another's FMAs); weights staged transposed in shared memory ([k][oc]) by cp.async - the
// 32 KB second-layer block lands while layer 1 computes - so that lanes (consecutive oc) read
// consecutive words and the activations are warp broadcasts.
constexpr int kMlpPts = 32;
constexpr int kMlpSmemFloats = 32 * 64 + 3 * 8 + 64 * 128 + 8 * 16 + 72 + 144 + kMlpPts * 36 + kMlpPts * 73;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

struct PointMlpArgs {
  const float* values;              // [B,32,P]
  const float* points;              // [B,3,P] voxel frame
  const float *w1r, *b1r;           // [32,64] (k-major)
  const float *w1p, *b1p;           // [3,8]
  const float *w2r, *b2r;           // [64,128]
  const float *w2p, *b2p;           // [8,16]
  int B, P;
  float center;
  bf16* feat;
  int ldf;
  float* feat2;
  float* feat1;                     // [N,72] fp32 or null
};

__device__ __forceinline__ void point_mlp_body(const PointMlpArgs& a, int cta, float* sm) {
  const float* __restrict__ values = a.values;
  const float* __restrict__ points = a.points;
  const float* __restrict__ w1r = a.w1r; const float* __restrict__ b1r = a.b1r;
  const float* __restrict__ w1p = a.w1p; const float* __restrict__ b1p = a.b1p;
  const float* __restrict__ w2r = a.w2r; const float* __restrict__ b2r = a.b2r;
  const float* __restrict__ w2p = a.w2p; const float* __restrict__ b2p = a.b2p;
  const int B = a.B, P = a.P, ldf = a.ldf;
  const float center = a.center;
  bf16* __restrict__ feat = a.feat;
  float* __restrict__ feat2 = a.feat2;
  float* __restrict__ feat1 = a.feat1;
  float* t1r = sm;                 // [32][64]
  float* t1p = t1r + 32 * 64;      // [3][8]
  float* t2r = t1p + 3 * 8;        // [64][128]  (16-byte aligned: 2048 + 24 floats before it)
  float* t2p = t2r + 64 * 128;     // [8][16]
  float* bb1 = t2p + 8 * 16;       // [72]
  float* bb2 = bb1 + 72;           // [144]
  float* xin = bb2 + 144;          // [kMlpPts][36]
  float* h1 = xin + kMlpPts * 36;  // [kMlpPts][73]
  const int tid = threadIdx.x;
  // weights arrive pre-transposed ([k][oc], done once when the model packs its weights)
  for (int e = tid; e < 64 * 32 / 4; e += 256) cp_async16(t1r + e * 4, w1r + e * 4);
  cp_async_commit();
  for (int e = tid; e < 128 * 64 / 4; e += 256) cp_async16(t2r + e * 4, w2r + e * 4);
  cp_async_commit();
  const int n0 = cta * kMlpPts;
  const int NP = B * P;
  {
    // inputs: pt fastest (coalesced along P); a CTA's points span at most two objects
    const int pt = tid & (kMlpPts - 1);
    const int n = n0 + pt;
    const int b = n / P, p = n - b * P;
    const bool ok = n < NP;
    const float* vsrc = values + ((long long)b * 32) * P + p;
    const float* psrc = points + ((long long)b * 3) * P + p;
    float v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int c = (tid >> 5) + 8 * i;          // 8 warps x 5 = 40 >= 35 channels
      v[i] = 0.f;
      if (ok && c < 35) v[i] = c < 32 ? __ldg(vsrc + (long long)c * P) : center - __ldg(psrc + (long long)(c - 32) * P);
    }
    for (int e = tid; e < 8 * 3; e += 256) t1p[e] = w1p[e];
    for (int e = tid; e < 16 * 8; e += 256) t2p[e] = w2p[e];
    for (int e = tid; e < 72; e += 256) bb1[e] = e < 64 ? b1r[e] : b1p[e - 64];
    for (int e = tid; e < 144; e += 256) bb2[e] = e < 128 ? b2r[e] : b2p[e - 128];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int c = (tid >> 5) + 8 * i;
      if (c < 35) xin[pt * 36 + c] = v[i];
    }
  }
  cp_async_wait<1>();              // layer-1 weights landed (layer 2 still in flight)
  __syncthreads();
  // stage 1: 32 points x 72 channels; thread = 4 points x 1 channel
  //   rgb: 64 oc x 32 pts -> (8 pt-groups of 4) x 64 oc = 512 work items, 2 per thread
  for (int e = tid; e < (kMlpPts / 4) * 64; e += 256) {
    int oc = e % 64, pg = e / 64;
    float a0 = bb1[oc], a1 = a0, a2 = a0, a3 = a0;
    const float* x0 = xin + (pg * 4) * 36;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      float wv = t1r[k * 64 + oc];
      a0 = fmaf(wv, x0[k], a0);
      a1 = fmaf(wv, x0[36 + k], a1);
      a2 = fmaf(wv, x0[72 + k], a2);
      a3 = fmaf(wv, x0[108 + k], a3);
    }
    h1[(pg * 4 + 0) * 73 + oc] = fmaxf(a0, 0.f);
    h1[(pg * 4 + 1) * 73 + oc] = fmaxf(a1, 0.f);
    h1[(pg * 4 + 2) * 73 + oc] = fmaxf(a2, 0.f);
    h1[(pg * 4 + 3) * 73 + oc] = fmaxf(a3, 0.f);
  }
  for (int e = tid; e < kMlpPts * 8; e += 256) {
    int pt = e / 8, o = e % 8;
    float acc = bb1[64 + o];
#pragma unroll
    for (int k = 0; k < 3; ++k) acc = fmaf(t1p[k * 8 + o], xin[pt * 36 + 32 + k], acc);
    h1[pt * 73 + 64 + o] = fmaxf(acc, 0.f);
  }
  cp_async_wait<0>();
  __syncthreads();
  for (int e = tid; e < kMlpPts * 36; e += 256) {      // 72 channels as bf16 pairs
    int pt = e / 36, oc = (e % 36) * 2;
    int n = n0 + pt;
    if (n < NP) {
      const float u0 = h1[pt * 73 + oc], u1 = h1[pt * 73 + oc + 1];
      *reinterpret_cast<__nv_bfloat162*>(feat + (long long)n * ldf + oc) = __floats2bfloat162_rn(u0, u1);
      if (feat1) *reinterpret_cast<float2*>(feat1 + (long long)n * 72 + oc) = make_float2(u0, u1);
    }
  }
  // stage 2 rgb: 128 oc x 32 pts; thread = 4 points x 4 consecutive oc (one LDS.128 of weights
  // feeds 16 FMAs): 32 oc-quads x 8 point-groups = 256 items, 1 per thread
  for (int e = tid; e < 32 * (kMlpPts / 4); e += 256) {
    int oq = e % 32, pg = e / 32;
    float4 bq = *reinterpret_cast<const float4*>(bb2 + oq * 4);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = bq.x; acc[i][1] = bq.y; acc[i][2] = bq.z; acc[i][3] = bq.w; }
    const float* hp = h1 + (pg * 4) * 73;
#pragma unroll 4
    for (int k = 0; k < 64; ++k) {
      float4 wv = *reinterpret_cast<const float4*>(t2r + k * 128 + oq * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x = hp[i * 73 + k];
        acc[i][0] = fmaf(wv.x, x, acc[i][0]);
        acc[i][1] = fmaf(wv.y, x, acc[i][1]);
        acc[i][2] = fmaf(wv.z, x, acc[i][2]);
        acc[i][3] = fmaf(wv.w, x, acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long n = n0 + pg * 4 + i;
      if (n >= NP) continue;
      float v0 = fmaxf(acc[i][0], 0.f), v1 = fmaxf(acc[i][1], 0.f), v2 = fmaxf(acc[i][2], 0.f),
            v3 = fmaxf(acc[i][3], 0.f);
      *reinterpret_cast<float4*>(feat2 + n * 144 + oq * 4) = make_float4(v0, v1, v2, v3);
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1b = __floats2bfloat162_rn(v2, v3);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1b);
      *reinterpret_cast<uint2*>(feat + n * ldf + 72 + oq * 4) = u;
    }
  }
  for (int e = tid; e < kMlpPts * 16; e += 256) {
    int pt = e / 16, o = e % 16;
    float acc = bb2[128 + o];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(t2p[k * 16 + o], h1[pt * 73 + 64 + k], acc);
    acc = fmaxf(acc, 0.f);
    long long n = n0 + pt;
    if (n < NP) {
      feat[n * ldf + 72 + 128 + o] = __float2bfloat16(acc);
      feat2[n * 144 + 128 + o] = acc;
    }
  }
}

__global__ void __launch_bounds__(256) k_point_mlp(PointMlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  point_mlp_body(a, blockIdx.x, sm);
}

// ------------------------------------------------------------------ occupancy stencils
// conv1_occ: Conv3D(1->8, k3, s1, p1) + ReLU; channels-last fp32 output [B,V,8]
__global__ void k_occ_conv1(const float* __restrict__ gne, const float* __restrict__ w,
                            const float* __restrict__ bias, int B, int D,
                            float* __restrict__ h1) {
  __shared__ float sw[8 * 27 + 8];
  for (int e = threadIdx.x; e < 8 * 27; e += blockDim.x) sw[e] = w[e];
  if (threadIdx.x < 8) sw[8 * 27 + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  long long V = (long long)D * D * D;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * V) return;
  long long b = i / V, v = i % V;
  int z = (int)(v % D), y = (int)((v / D) % D), x = (int)(v / ((long long)D * D));
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = sw[8 * 27 + c];
  for (int kd = 0; kd < 3; ++kd) {
    int xx = x + kd - 1;
    if (xx < 0 || xx >= D) continue;
    for (int kh = 0; kh < 3; ++kh) {
      int yy = y + kh - 1;
      if (yy < 0 || yy >= D) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int zz = z + kw - 1;
        if (zz < 0 || zz >= D) continue;
        float in = gne[b * V + ((long long)xx * D + yy) * D + zz];
        int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(sw[c * 27 + tap], in, acc[c]);
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(h1 + i * 8);
  o[0] = make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
  o[1] = make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
}

// conv2_occ: Conv3D(8->16, k3, s1, p2, dilate 2) + ReLU.
// One thread = 4 consecutive z voxels x 16 output channels (64 accumulators): every LDS.128 of
// weights (4 output channels, warp broadcast) feeds 16 FMAs; inputs are 16-byte channel vectors.
// Output: fp32 [B,V,16] (h2) or, if X != null, bf16 straight into channels [c_off, c_off+16) of
// the s2d conv3 input.
__global__ void __launch_bounds__(64)
k_occ_conv2(const float* __restrict__ h1, const float* __restrict__ w,
            const float* __restrict__ bias, int B, int D, float* __restrict__ h2,
            bf16* __restrict__ X, int Ct, int c_off) {
  __shared__ __align__(16) float sw[27 * 8 * 16 + 16];   // [tap][ci][co]
  for (int e = threadIdx.x; e < 16 * 8 * 27; e += blockDim.x) {
    int co = e / (8 * 27), r = e % (8 * 27), ci = r / 27, tap = r % 27;   // OIDHW
    sw[(tap * 8 + ci) * 16 + co] = w[e];
  }
  if (threadIdx.x < 16) sw[27 * 8 * 16 + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int Dz4 = D / 4;
  long long V = (long long)D * D * D;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * D * D * Dz4) return;
  int z0 = (int)(i % Dz4) * 4;
  long long t = i / Dz4;
  int y = (int)(t % D), x = (int)((t / D) % D);
  long long b = t / ((long long)D * D);
  float acc[4][16];
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[v][c] = sw[27 * 8 * 16 + c];
  for (int kd = 0; kd < 3; ++kd) {
    int xx = x + 2 * (kd - 1);
    if (xx < 0 || xx >= D) continue;
    for (int kh = 0; kh < 3; ++kh) {
      int yy = y + 2 * (kh - 1);
      if (yy < 0 || yy >= D) continue;
      const float* row = h1 + (b * V + ((long long)xx * D + yy) * D) * 8;
      // z positions z0-2 .. z0+5 (8 of them) cover the 3 dilated taps of the 4 voxels
      float in[8][8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        int zz = z0 - 2 + q;
        if (zz >= 0 && zz < D) {
          const float4* src = reinterpret_cast<const float4*>(row + (long long)zz * 8);
          float4 lo = __ldg(src), hi = __ldg(src + 1);
          in[q][0] = lo.x; in[q][1] = lo.y; in[q][2] = lo.z; in[q][3] = lo.w;
          in[q][4] = hi.x; in[q][5] = hi.y; in[q][6] = hi.z; in[q][7] = hi.w;
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) in[q][c] = 0.f;
        }
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const float4* wv = reinterpret_cast<const float4*>(sw + (tap * 8 + ci) * 16);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            float4 ww = wv[q4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              float a = in[v + 2 * kw][ci];          // voxel z0+v, tap offset 2*(kw-1)
              acc[v][4 * q4 + 0] = fmaf(ww.x, a, acc[v][4 * q4 + 0]);
              acc[v][4 * q4 + 1] = fmaf(ww.y, a, acc[v][4 * q4 + 1]);
              acc[v][4 * q4 + 2] = fmaf(ww.z, a, acc[v][4 * q4 + 2]);
              acc[v][4 * q4 + 3] = fmaf(ww.w, a, acc[v][4 * q4 + 3]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    int z = z0 + v;
    if (z >= D) break;
    if (X) {
      const int J = D / 2 + 1;
      int pd = x + 1, ph = y + 1, pw = z + 1;
      int r = ((pd & 1) << 2) | ((ph & 1) << 1) | (pw & 1);
      long long dst = ((((long long)b * J + (pd >> 1)) * J + (ph >> 1)) * J + (pw >> 1)) * (8LL * Ct) +
                      (long long)r * Ct + c_off;
      uint4 o[2];
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        h[c] = __floats2bfloat162_rn(fmaxf(acc[v][2 * c], 0.f), fmaxf(acc[v][2 * c + 1], 0.f));
      uint4* d4 = reinterpret_cast<uint4*>(X + dst);
      d4[0] = o[0];
      d4[1] = o[1];
    } else {
      float4* o = reinterpret_cast<float4*>(h2 + (b * V + ((long long)x * D + y) * D + z) * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        o[q] = make_float4(fmaxf(acc[v][4 * q], 0.f), fmaxf(acc[v][4 * q + 1], 0.f),
                           fmaxf(acc[v][4 * q + 2], 0.f), fmaxf(acc[v][4 * q + 3], 0.f));
    }
  }
}

// ------------------------------------------------------------------ conv2_occ on tensor cores
// FP32 FFMA with three register operands issues at half rate on this SM, so even a perfectly
// tiled SIMT stencil needs ~55 us for the 0.9 GMAC of conv2_occ.  Warp-level mma.sync
// (m16n8k16, bf16 in, fp32 accumulate) fits this thin layer (N = 16 output channels, K = 27
// taps x 8 channels): tcgen05's 128-row tiles + TMA boxes cannot express the dilated gather.
//   A (16 voxels x 16 k): gathered straight from a shared-memory tile of conv1's bf16 output
//   B (16 k x 8 co): all 14 k-steps x 2 n-tiles live in 56 registers per lane for the whole CTA
// One CTA = one x-slab (32 x 32 voxels) of one object; smem tile = slabs x-2, x, x+2 with a
// 2-voxel zero halo in y and z.
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4],
                                               const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int kOccT = 36;                       // 32 + 2*2 halo
constexpr int kOccTileBytes = 3 * kOccT * kOccT * 16;
constexpr int kOccFragBytes = 14 * 2 * 2 * 32 * 4;     // conv2_occ weights in mma fragment order
constexpr int kOccConv2Bytes = kOccTileBytes + kOccFragBytes;

// conv2_occ's weights: read once per CTA, coalesced, and scattered as bf16 into mma fragment order
// [s][nt][h][lane] in shared memory (each lane then fetches its 56 words conflict-free; 112
// strided scalar loads per thread straight from global were 20 % of the kernel).  Call early: the
// loads are independent of everything else the CTA stages.
__device__ __forceinline__ void occ2_stage_weights(unsigned char* occ_smem, const float* __restrict__ w) {
  const int tid = threadIdx.x;
  bf16* wf = reinterpret_cast<bf16*>(occ_smem + kOccTileBytes);
  float wv[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) {                 // 16*8*27 = 3456 = 13.5 x 256: loads first
    const int e = tid + 256 * i;
    wv[i] = e < 16 * 8 * 27 ? __ldg(w + e) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 14; ++i) {
    const int e = tid + 256 * i;
    if (e < 16 * 8 * 27) {
      const int tap = e % 27, ci = (e / 27) & 7, co = e / (27 * 8);
      const int idx = (((tap >> 1) * 2 + (co >> 3)) * 2 + (tap & 1)) * 32 + ((co & 7) * 4 + (ci >> 1));
      wf[idx * 2 + (ci & 1)] = __float2bfloat16(wv[i]);
    }
  }
  if (tid < 64) reinterpret_cast<uint32_t*>(wf)[((13 * 2 + (tid >> 5)) * 2 + 1) * 32 + (tid & 31)] = 0u;
}

// conv2_occ of one x-slab from the staged tile + weights (call with shared memory NOT yet
// synchronised: the barrier is the first thing it does)
__device__ __forceinline__ void occ2_mma_slab(const unsigned char* occ_smem, const float* __restrict__ /*w*/,
                                              const float* __restrict__ bias, int b, int x,
                                              bf16* __restrict__ X, int Ct, int c_off) {
  constexpr int D = 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // ---- B fragments: k = tap_in_step*8 + ci, step s covers taps 2s, 2s+1 (tap 27 = zero pad),
  // staged by occ2_stage_weights
  const bf16* wf = reinterpret_cast<const bf16*>(occ_smem + kOccTileBytes);
  __syncthreads();
  uint32_t bf[14][2][2];
#pragma unroll
  for (int s = 0; s < 14; ++s)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        bf[s][nt][h] = reinterpret_cast<const uint32_t*>(wf)[((s * 2 + nt) * 2 + h) * 32 + lane];
  float bias_r[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    bias_r[nt][0] = bias[nt * 8 + 2 * t];
    bias_r[nt][1] = bias[nt * 8 + 2 * t + 1];
  }
  // 64 groups of 16 consecutive-z voxels (y = grp/2, z0 = (grp&1)*16); 8 per warp.  One
  // ldmatrix.x4 per k-step fetches the whole A fragment: matrices 0/1 = voxels z0+[0,8) / z0+[8,16)
  // at tap 2s, matrices 2/3 = the same voxels at tap 2s+1; a matrix row is one voxel's 8 channels
  // (16 B), which is exactly the (row g, k 2t..2t+1) ownership of the m16n8k16 A fragment.
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(occ_smem);
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhi = lane >> 4;
  for (int grp = warp; grp < 64; grp += 8) {
    const int y = grp >> 1, z0 = (grp & 1) * 16;
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      acc[nt][0] = bias_r[nt][0]; acc[nt][1] = bias_r[nt][1];
      acc[nt][2] = bias_r[nt][0]; acc[nt][3] = bias_r[nt][1];
    }
    const uint32_t row_s = tile_s + (uint32_t)((y * kOccT + z0 + lrow) * 16);
#pragma unroll
    for (int s = 0; s < 14; ++s) {
      // tile coords of a tap: slab kd, y + 2*(kh-1) + 2, z + 2*(kw-1) + 2.  Tap 27 does not
      // exist (its weights are zero): its lanes re-read tap 26, any finite value will do.
      const int ta = 2 * s, tb = 2 * s + 1 < 27 ? 2 * s + 1 : 26;
      const int offa = (((ta / 9) * kOccT + 2 * ((ta / 3) % 3)) * kOccT + 2 * (ta % 3)) * 16;
      const int offb = (((tb / 9) * kOccT + 2 * ((tb / 3) % 3)) * kOccT + 2 * (tb % 3)) * 16;
      uint32_t af[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(af[0]), "=r"(af[1]), "=r"(af[2]), "=r"(af[3])
                   : "r"(row_s + (uint32_t)(lhi ? offb : offa)));
      mma_bf16_16816(acc[0], af, bf[s][0]);
      mma_bf16_16816(acc[1], af, bf[s][1]);
    }
    // ---- epilogue: ReLU, bf16, into channels [c_off, c_off+16) of the s2d conv3 input
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int z = z0 + g + 8 * r;
      const int J = D / 2 + 1;
      int pd = x + 1, ph = y + 1, pw = z + 1;
      int rr = ((pd & 1) << 2) | ((ph & 1) << 1) | (pw & 1);
      long long dst = ((((long long)b * J + (pd >> 1)) * J + (ph >> 1)) * J + (pw >> 1)) * (8LL * Ct) +
                      (long long)rr * Ct + c_off;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        __nv_bfloat162 o = __floats2bfloat162_rn(fmaxf(acc[nt][2 * r], 0.f), fmaxf(acc[nt][2 * r + 1], 0.f));
        *reinterpret_cast<__nv_bfloat162*>(X + dst + nt * 8 + 2 * t) = o;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_occ_conv2_mma(const bf16* __restrict__ h1 /*[B,V,8] bf16*/, const float* __restrict__ w /*OIDHW*/,
                const float* __restrict__ bias, int B, bf16* __restrict__ X, int Ct, int c_off) {
  constexpr int D = 32;
  extern __shared__ __align__(16) unsigned char occ_smem[];
  uint4* tile = reinterpret_cast<uint4*>(occ_smem);          // [3][36][36] x 16 B
  const int x = blockIdx.x % D, b = blockIdx.x / D;
  // ---- stage the three input slabs (zero outside the grid)
  for (int e = threadIdx.x; e < 3 * kOccT * kOccT; e += 256) {
    int zz = e % kOccT - 2, yy = (e / kOccT) % kOccT - 2, s = e / (kOccT * kOccT);
    int xx = x + 2 * (s - 1);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (xx >= 0 && xx < D && yy >= 0 && yy < D && zz >= 0 && zz < D)
      v = __ldg(reinterpret_cast<const uint4*>(h1) + ((long long)b * D * D * D + (xx * D + yy) * D + zz));
    tile[e] = v;
  }
  occ2_stage_weights(occ_smem, w);
  occ2_mma_slab(occ_smem, w, bias, b, x, X, Ct, c_off);
}

// conv1_occ + conv2_occ in one kernel: the CTA of slab x evaluates conv1_occ (1 -> 8, k3, p1, ReLU,
// bf16) for the three slabs x-2, x, x+2 that the dilated conv2_occ reads, straight from the
// occupancy grid (7 input slabs, 7 KB as bytes) into the shared-memory tile - conv1's output never
// goes through global memory and there is no second launch.  Each conv1 slab is evaluated by the
// three CTAs that consume it (3 x 0.06 GFMA in total, ~4 us of FFMA per SM) in exchange for the
// 12.6 MB round trip and the launch boundary.  Tap order and arithmetic = k_occ_conv1_bf16
// (out-of-grid taps contribute fmaf(w, 0, acc) = acc), so the result is bit-identical.
constexpr int kOccG = 34;                        // rows per input slab: 32 + 2*1 halo
constexpr int kOccGS = 40;                       // row stride: col -1 at element 3, col 0 at 4 (aligned)
template <typename TIn>
constexpr int occ_fused_bytes() {
  return kOccConv2Bytes + (27 * 8 + 8) * 4 + 7 * kOccG * kOccGS * (int)sizeof(TIn);
}

// the 6 input values (columns 4zq-1 .. 4zq+4) of one staged row as floats
__device__ __forceinline__ void occ_row6(const float* row, int zq, float (&in)[6]) {
#pragma unroll
  for (int q = 0; q < 6; ++q) in[q] = row[3 + 4 * zq + q];
}
__device__ __forceinline__ void occ_row6(const uint8_t* row, int zq, float (&in)[6]) {
  // bytes -> float without I2F: 0x4B0000bb is 8388608 + bb exactly
  const uint32_t* w = reinterpret_cast<const uint32_t*>(row) + zq;
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
  in[0] = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7443)) - 8388608.f;
  in[1] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7440)) - 8388608.f;
  in[2] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7441)) - 8388608.f;
  in[3] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7442)) - 8388608.f;
  in[4] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7443)) - 8388608.f;
  in[5] = __uint_as_float(__byte_perm(w2, 0x4B000000u, 0x7440)) - 8388608.f;
}

template <typename TIn>
__global__ void __launch_bounds__(256)
k_occ_fused(const TIn* __restrict__ gne /*[B,32,32,32]*/, const float* __restrict__ w1,
            const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
            int B, bf16* __restrict__ X, int Ct, int c_off) {
  constexpr int D = 32;
  extern __shared__ __align__(16) unsigned char occ_smem[];
  uint4* tile = reinterpret_cast<uint4*>(occ_smem);                        // [3][36][36] x 16 B
  float* sw = reinterpret_cast<float*>(occ_smem + kOccConv2Bytes);         // [27][8] + bias[8]
  TIn* gs = reinterpret_cast<TIn*>(sw + 27 * 8 + 8);                       // [7][34][40]
  const int x = blockIdx.x % D, b = blockIdx.x / D;
  const int tid = threadIdx.x;
  occ2_stage_weights(occ_smem, w2);
  // ---- stage the 7 input slabs x-3 .. x+3 (zero outside the grid), 4-byte units
  constexpr int kPer = 4 / (int)sizeof(TIn);           // elements per 4-byte unit
  constexpr int kUnits = kOccGS / kPer;                // units per staged row
  constexpr int kStageIters = (7 * kOccG * kUnits + 255) / 256;
  for (int i0 = 0; i0 < kStageIters; i0 += 10) {       // loads of 10 units in flight, then stores
    uint32_t v[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const int e = tid + 256 * (i0 + i);
      const int u = e % kUnits, yy = (e / kUnits) % kOccG - 1, xx = x - 3 + e / (kUnits * kOccG);
      const int z0 = u * kPer - 4;                     // first grid column of this unit
      v[i] = 0u;
      if (e < 7 * kOccG * kUnits && xx >= 0 && xx < D && yy >= 0 && yy < D && z0 >= 0 && z0 < D)
        v[i] = __ldg(reinterpret_cast<const uint32_t*>(gne + (long long)b * D * D * D + (xx * D + yy) * D + z0));
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const int e = tid + 256 * (i0 + i);
      if (e < 7 * kOccG * kUnits) reinterpret_cast<uint32_t*>(gs)[e] = v[i];
    }
  }
  for (int e = tid; e < 27 * 8; e += 256) sw[e] = w1[(e & 7) * 27 + (e >> 3)];   // [tap][c]
  if (tid < 8) sw[27 * 8 + tid] = b1[tid];
  // zero padding of conv2_occ: the tile's 2-voxel halo (4 rows x 36 + 32 rows x 4 columns per
  // slab) and, at the grid boundary, the whole slab
  for (int e = tid; e < 3 * 272; e += 256) {
    const int s3 = e / 272, r = e % 272;
    int row, col;
    if (r < 144) { row = r / 36; row = row < 2 ? row : row + 32; col = r % 36; }
    else { row = 2 + (r - 144) / 4; col = (r - 144) % 4; col = col < 2 ? col : col + 32; }
    tile[(s3 * kOccT + row) * kOccT + col] = make_uint4(0, 0, 0, 0);
  }
  if (x < 2 || x >= D - 2) {
    const int s3 = x < 2 ? 0 : 2;                      // slab x-2 < 0 or x+2 >= D
    for (int e = tid; e < D * D; e += 256)
      tile[(s3 * kOccT + 2 + e / D) * kOccT + 2 + e % D] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  // ---- conv1: item = (slab s, row yy, 4 consecutive z): 3 x 32 x 8 = 768 items, 3 per thread
  for (int it = tid; it < 3 * D * 8; it += 256) {
    const int zq = it & 7, yy = (it >> 3) & 31, s = it >> 8;
    const int xx = x + 2 * (s - 1);
    if (xx < 0 || xx >= D) continue;                 // slab outside the grid: conv2's zero padding
    float acc[4][8];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[v][c] = sw[27 * 8 + c];
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        // input slab (xx + kd - 1) - (x - 3) = 2s + kd; staged row (yy + kh - 1) + 1
        float in[6];
        occ_row6(gs + ((2 * s + kd) * kOccG + (yy + kh)) * kOccGS, zq, in);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 wa = *reinterpret_cast<const float4*>(sw + ((kd * 3 + kh) * 3 + kw) * 8);
          const float4 wb = *reinterpret_cast<const float4*>(sw + ((kd * 3 + kh) * 3 + kw) * 8 + 4);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float a = in[v + kw];
            acc[v][0] = fmaf(wa.x, a, acc[v][0]); acc[v][1] = fmaf(wa.y, a, acc[v][1]);
            acc[v][2] = fmaf(wa.z, a, acc[v][2]); acc[v][3] = fmaf(wa.w, a, acc[v][3]);
            acc[v][4] = fmaf(wb.x, a, acc[v][4]); acc[v][5] = fmaf(wb.y, a, acc[v][5]);
            acc[v][6] = fmaf(wb.z, a, acc[v][6]); acc[v][7] = fmaf(wb.w, a, acc[v][7]);
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      uint4 o;
      __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaxf(acc[v][0], 0.f), fmaxf(acc[v][1], 0.f));
      __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaxf(acc[v][2], 0.f), fmaxf(acc[v][3], 0.f));
      __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaxf(acc[v][4], 0.f), fmaxf(acc[v][5], 0.f));
      __nv_bfloat162 p3 = __floats2bfloat162_rn(fmaxf(acc[v][6], 0.f), fmaxf(acc[v][7], 0.f));
      o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
      o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
      tile[(s * kOccT + (yy + 2)) * kOccT + (4 * zq + v + 2)] = o;
    }
  }
  occ2_mma_slab(occ_smem, w2, b2, b, x, X, Ct, c_off);
}

// conv1_occ variant writing bf16 channels-last (input of k_occ_conv2_mma)
template <typename TIn>
__global__ void k_occ_conv1_bf16(const TIn* __restrict__ gne, const float* __restrict__ w,
                                 const float* __restrict__ bias, int B, int D,
                                 bf16* __restrict__ h1) {
  __shared__ float sw[8 * 27 + 8];
  for (int e = threadIdx.x; e < 8 * 27; e += blockDim.x) sw[e] = w[e];
  if (threadIdx.x < 8) sw[8 * 27 + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  long long V = (long long)D * D * D;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * V) return;
  long long b = i / V, v = i % V;
  int z = (int)(v % D), y = (int)((v / D) % D), x = (int)(v / ((long long)D * D));
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = sw[8 * 27 + c];
  for (int kd = 0; kd < 3; ++kd) {
    int xx = x + kd - 1;
    if (xx < 0 || xx >= D) continue;
    for (int kh = 0; kh < 3; ++kh) {
      int yy = y + kh - 1;
      if (yy < 0 || yy >= D) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int zz = z + kw - 1;
        if (zz < 0 || zz >= D) continue;
        float in = (float)gne[b * V + ((long long)xx * D + yy) * D + zz];
        int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(sw[c * 27 + tap], in, acc[c]);
      }
    }
  }
  uint4 o;
  __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f));
  __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
  __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f));
  __nv_bfloat162 p3 = __floats2bfloat162_rn(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
  o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
  o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
  reinterpret_cast<uint4*>(h1)[i] = o;
}

// ------------------------------------------------------------------ fused voxelise -> s2d
// Model path of average_voxelization_3d (model.py:143-164: origin 0, pitch 1, 32^3): the 144
// averaged feature channels go straight into the bf16 s2d conv3 input.  The grid is >= 97 %
// empty, so instead of rewriting 100 MB per call the kernel zeroes only the voxels the PREVIOUS
// call wrote (keys kept in `prev_keys`) and then writes the new ones.  Same arithmetic as
// k_avg_tiles: per-voxel fp32 sums in ascending point order, IEEE divide, then bf16 rounding.
__device__ __forceinline__ long long s2d_voxel_offset(int b, int flat, int D, int Ct) {
  const int J = D / 2 + 1;
  int iz = flat % D, iy = (flat / D) % D, ix = flat / (D * D);
  int pd = ix + 1, ph = iy + 1, pw = iz + 1;
  int r = ((pd & 1) << 2) | ((ph & 1) << 1) | (pw & 1);
  return ((((long long)b * J + (pd >> 1)) * J + (ph >> 1)) * J + (pw >> 1)) * (8LL * Ct) +
         (long long)r * Ct;
}

// one warp per point of the previous call: zero its voxel's C channels
__global__ void k_s2d_clear(const int* __restrict__ prev_keys, int N, int C, int D, int Ct,
                            bf16* __restrict__ X) {
  int n = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  int key = prev_keys[n];
  if (key < 0) return;
  key &= MF_S2D_KEY_MASK;
  const int V = D * D * D;
  bf16* dst = X + s2d_voxel_offset(key / V, key % V, D, Ct);
  for (int c = lane * 2; c < C; c += 64)
    *reinterpret_cast<uint32_t*>(dst + c) = 0u;
}

// One CTA per object: voxel keys + the "shares its voxel" bit.  Two direct-addressed bit tables in
// shared memory (D^3 / 8 bytes each: 4 KB at D = 32): `once` = some point fell into the voxel,
// `twice` = a second one did.  When they do not fit (D > 96) every point is conservatively
// marked as shared.
__device__ __forceinline__ int s2d_flat_key(const float* __restrict__ points, int b, int p, int P,
                                            int D) {
  const float fx = points[((long long)b * 3 + 0) * P + p], fy = points[((long long)b * 3 + 1) * P + p],
              fz = points[((long long)b * 3 + 2) * P + p];
  int ix = static_cast<int>(roundf(fx));   // (p - 0) / 1.0
  int iy = static_cast<int>(roundf(fy));
  int iz = static_cast<int>(roundf(fz));
  // NaN points are dropped (the public operator raises "points include nan";
  // Model.predict masks them out before this point, model.py:178)
  bool ok = !(isnan(fx) || isnan(fy) || isnan(fz)) && ix >= 0 && ix < D && iy >= 0 && iy < D &&
            iz >= 0 && iz < D;
  return ok ? (ix * D + iy) * D + iz : -1;
}

__global__ void __launch_bounds__(1024)
k_s2d_keys(const float* __restrict__ points /*[B,3,P]*/, int B, int P, int D,
           int* __restrict__ keys, int use_table) {
  extern __shared__ unsigned int bits[];     // once[W] | twice[W], W = ceil(D^3 / 32) words
  const int b = blockIdx.x, V = D * D * D, W = (V + 31) / 32;
  // the first 1024 points stay in registers across the barrier (P = 1000 in the model)
  const int flat0 = threadIdx.x < P ? s2d_flat_key(points, b, threadIdx.x, P, D) : -1;
  if (use_table) {
    for (int e = threadIdx.x; e < 2 * W; e += 1024) bits[e] = 0u;
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += 1024) {
      const int flat = p < 1024 ? flat0 : s2d_flat_key(points, b, p, P, D);
      if (flat >= 0) {
        const unsigned int m = 1u << (flat & 31);
        if (atomicOr(&bits[flat >> 5], m) & m) atomicOr(&bits[W + (flat >> 5)], m);
      }
    }
    __syncthreads();
  }
  for (int p = threadIdx.x; p < P; p += 1024) {
    const int flat = p < 1024 ? flat0 : s2d_flat_key(points, b, p, P, D);
    int key = -1;
    if (flat >= 0) {
      const bool dup = !use_table || ((bits[W + (flat >> 5)] >> (flat & 31)) & 1u);
      key = (b * V + flat) | (dup ? MF_S2D_DUP_BIT : 0);
    }
    keys[(long long)b * P + p] = key;
  }
}

// one warp per point, 8 consecutive points per CTA.  A point alone in its voxel (no DUP bit; the
// overwhelmingly common case: 1000 points in 32^3 voxels) converts its own feature row.  For shared
// voxels the lowest-index point ("leader") sums the voxel's points in ascending order, lanes over
// channels; only CTAs that contain such a point stage their object's keys in shared memory for
// the two scans (leader test over earlier points, member search over later ones).
constexpr int kScatWarps = 8;
__global__ void __launch_bounds__(256)
k_s2d_scatter(const float* __restrict__ feat2 /*[N,C]*/, const int* __restrict__ keys, int B,
              int P, int C, int D, int Ct, bf16* __restrict__ X) {
  extern __shared__ int skeys[];               // keys[klo, khi): the objects this CTA touches
  const int N = B * P;
  const int nf = blockIdx.x * kScatWarps;
  const int n = nf + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int raw = n < N ? keys[n] : -1;
  const bool dup = raw >= 0 && (raw & MF_S2D_DUP_BIT);
  const int V = D * D * D;
  const int b = n / P;
  if (raw >= 0 && !dup) {
    // same arithmetic as the general path with one member: (0 + v) / 1
    const float* src = feat2 + (long long)n * C;
    bf16* dst = X + s2d_voxel_offset(b, (raw & MF_S2D_KEY_MASK) - b * V, D, Ct);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int c = lane + 32 * k;
      v[k] = c < C ? __ldg(src + c) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int c = lane + 32 * k;
      if (c < C) dst[c] = __float2bfloat16(__fdiv_rn(__fadd_rn(0.f, v[k]), 1.f));
    }
  }
  if (!__syncthreads_or(dup)) return;
  const int nl = min(nf + kScatWarps, N) - 1;
  const int klo = (nf / P) * P, khi = min((nl / P + 1) * P, N);
  for (int e = threadIdx.x; e < khi - klo; e += 256) skeys[e] = keys[klo + e];
  __syncthreads();
  if (!dup) return;
  const int lo = b * P;
  const int* sk = skeys + (lo - klo);          // this object's keys, indexed by j - lo
  const int m = n - lo;
  // leader test: any earlier point of this object with the same key?
  bool earlier = false;
  for (int j0 = 0; j0 < m; j0 += 128) {
    bool any = false;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int j = j0 + 32 * u + lane;
      any |= (j < m) && (sk[j] == raw);
    }
    if (__any_sync(0xffffffffu, any)) { earlier = true; break; }
  }
  if (earlier) return;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  int count = 0;
  for (int j0 = m - m % 32; j0 < P; j0 += 32) {     // chunks aligned so that j ascends
    int j = j0 + lane;
    bool hit = (j >= m) && (j < P) && (sk[j] == raw);
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    while (mask) {
      int l = __ffs(mask) - 1;
      mask &= mask - 1;
      const float* src = feat2 + (long long)(lo + j0 + l) * C;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int c = lane + 32 * k;
        if (c < C) acc[k] = __fadd_rn(acc[k], __ldg(src + c));
      }
      ++count;
    }
  }
  bf16* dst = X + s2d_voxel_offset(b, (raw & MF_S2D_KEY_MASK) - b * V, D, Ct);
  const float cf = (float)count;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = lane + 32 * k;
    if (c < C) dst[c] = __float2bfloat16(__fdiv_rn(acc[k], cf));
  }
}

// ---- sorted variant (P <= 4096): real objects put ~2/3 of their 1000 points into shared voxels,
// so instead of searching per voxel, every point's position in the (voxel, point index) order is
// computed up front by comparison counting - rank(p) = #{q : comp[q] < comp[p]} over the object's
// keys in shared memory, two threads per point, no barriers, 64 points per CTA so that the work
// spreads over B * P / 64 CTAs - while the point MLP runs on another branch.  A voxel's points
// are then consecutive in `order`, already in ascending point order, and the scatter does no
// search at all.  comp = voxel * n2 + point index (n2 = pow2 >= P) fits 31 bits (host-checked).
constexpr int kRankPts = 64;
// blockDim.x = 64 * tpp threads (tpp = 2, 4, ... threads per point, each counting a slice of the
// object's keys).  With X != null the CTA first zeroes the C channels of the voxels that ITS 64
// points occupied in the previous call (their old keys are still in keys[]: each CTA overwrites
// only its own slice, after this), which makes the separate sparse-clear launch unnecessary.
__device__ __forceinline__ void s2d_rank_body(int cta, const float* __restrict__ points, int B, int P,
                                              int D, int n2, int chunks, int* __restrict__ keys,
                                              int* __restrict__ order, bf16* __restrict__ X, int C,
                                              int Ct, int* comp /* shared, [(P+7)&~7] */) {
  const int b = cta / chunks, p0 = (cta % chunks) * kRankPts, V = D * D * D;
  const int P8 = (P + 7) & ~7;
  const int nthr = blockDim.x, tpp = nthr / kRankPts;
  if (X) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = nthr >> 5;
    for (int i = warp; i < kRankPts && p0 + i < P; i += nwarp) {
      int key = keys[(long long)b * P + p0 + i];
      if (key < 0) continue;
      key &= MF_S2D_KEY_MASK;
      bf16* dst = X + s2d_voxel_offset(key / V, key % V, D, Ct);
      for (int c = lane * 2; c < C; c += 64) *reinterpret_cast<uint32_t*>(dst + c) = 0u;
    }
  }
  for (int p = threadIdx.x; p < P8; p += nthr) {
    int c = 0x7FFFFFFF;
    if (p < P) {
      const int flat = s2d_flat_key(points, b, p, P, D);
      c = (flat < 0 ? V : flat) * n2 + p;             // dropped points sort behind every voxel
    }
    comp[p] = c;
  }
  __syncthreads();        // also orders the old-key reads above before the new-key writes below
  const int p = p0 + threadIdx.x / tpp, part = threadIdx.x % tpp;
  const int mine = p < P ? comp[p] : 0x7FFFFFFF;
  const int sh = 31 - __clz(n2);                       // comp >> sh = voxel
  const int vox = mine >> sh;
  int rank = 0, same = 0;
  // this thread's slice of the keys, 4 per shared-memory load
  const int4* c4 = reinterpret_cast<const int4*>(comp);
  const int q4 = P8 / 4, per = (q4 + tpp - 1) / tpp;
  for (int q = part * per; q < min(q4, (part + 1) * per); ++q) {
    const int4 c = c4[q];
    rank += (c.x < mine) + (c.y < mine) + (c.z < mine) + (c.w < mine);
    same += ((c.x >> sh) == vox) + ((c.y >> sh) == vox) + ((c.z >> sh) == vox) + ((c.w >> sh) == vox);
  }
  for (int o = 1; o < tpp; o <<= 1) {
    rank += __shfl_xor_sync(0xffffffffu, rank, o);
    same += __shfl_xor_sync(0xffffffffu, same, o);
  }
  if (part == 0 && p < P) {
    order[(long long)b * P + rank] = p;
    keys[(long long)b * P + p] = vox >= V ? -1 : (b * V + vox) | (same > 1 ? MF_S2D_DUP_BIT : 0);
  }
}

__global__ void __launch_bounds__(2 * kRankPts)
k_s2d_rank(const float* __restrict__ points /*[B,3,P]*/, int B, int P, int D, int n2, int chunks,
           int* __restrict__ keys, int* __restrict__ order) {
  extern __shared__ __align__(16) int comp[];         // [P rounded up to 8], padding = INT_MAX
  s2d_rank_body(blockIdx.x, points, B, P, D, n2, chunks, keys, order, nullptr, 0, 0, comp);
}

// point MLP + (sparse clear + rank) in ONE launch: CTAs [0, rank_ctas) do the voxel bookkeeping,
// the rest the MLP.  Independent graph branches start in an order the runtime picks; as CTAs of
// one grid the bookkeeping is dispatched first and runs beside the MLP by construction.
__global__ void __launch_bounds__(256)
k_point_mlp_vox(PointMlpArgs a, int rank_ctas, int D, int n2, int chunks, int* __restrict__ keys,
                int* __restrict__ order, bf16* __restrict__ X, int C, int Ct) {
  extern __shared__ __align__(16) float sm[];
  if ((int)blockIdx.x < rank_ctas)
    s2d_rank_body(blockIdx.x, a.points, a.B, a.P, D, n2, chunks, keys, order, X, C, Ct,
                  reinterpret_cast<int*>(sm));
  else
    point_mlp_body(a, blockIdx.x - rank_ctas, sm);
}

// one warp per SORTED position: the first position of a voxel's run sums the run in order
__global__ void __launch_bounds__(256)
k_s2d_scatter_sorted(const float* __restrict__ feat2 /*[N,C]*/, const int* __restrict__ keys,
                     const int* __restrict__ order, int B, int P, int C, int D, int Ct,
                     bf16* __restrict__ X) {
  const int gpos = blockIdx.x * kScatWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gpos >= B * P) return;
  const int b = gpos / P, i = gpos - b * P, lo = b * P;
  // lanes look at positions i-1+lane: lane 0 = the predecessor, lane 1 = this position, ...
  const int pj = i - 1 + lane;
  int oj = -1, kj = -2;
  if (pj >= 0 && pj < P) {
    oj = __ldg(order + lo + pj);
    kj = __ldg(keys + lo + oj);
  }
  const int raw = __shfl_sync(0xffffffffu, kj, 1);
  if (raw < 0 || __shfl_sync(0xffffffffu, kj, 0) == raw) return;     // dropped / not a run start
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  int count = 0;
  int base = i;                                   // position held by lane 1
  unsigned run = __ballot_sync(0xffffffffu, kj == raw) >> 1;          // bit m = position base+m
  for (;;) {
    const int len = run == 0x7FFFFFFFu ? 31 : __ffs(~run) - 1;        // leading run of ones
    for (int m = 0; m < len; ++m) {
      const float* src = feat2 + (long long)(lo + __shfl_sync(0xffffffffu, oj, m + 1)) * C;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        int c = lane + 32 * k;
        if (c < C) acc[k] = __fadd_rn(acc[k], __ldg(src + c));
      }
      ++count;
    }
    if (len < 31) break;
    base += 31;                                   // the run may continue: next 31 positions
    const int q = base - 1 + lane;
    oj = -1; kj = -2;
    if (q >= 0 && q < P) {
      oj = __ldg(order + lo + q);
      kj = __ldg(keys + lo + oj);
    }
    run = __ballot_sync(0xffffffffu, kj == raw) >> 1;
  }
  const int V = D * D * D;
  bf16* dst = X + s2d_voxel_offset(b, (raw & MF_S2D_KEY_MASK) - b * V, D, Ct);
  const float cf = (float)count;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = lane + 32 * k;
    if (c < C) dst[c] = __float2bfloat16(__fdiv_rn(acc[k], cf));
  }
}

// ------------------------------------------------------------------ s2d packing
// X[b][jd][jh][jw][r*Ct + c] = xpad[b][c][2jd+rd][2jh+rh][2jw+rw]  (bf16), Ct = C + Cocc.
// One CTA per (b, id, ih) row of D voxels: coalesced reads along iw, transpose in smem,
// coalesced channel-contiguous writes.  Border entries (xpad == 0) are never written: the
// buffer is zeroed once at allocation.
__global__ void __launch_bounds__(256)
k_pack_s2d(const float* __restrict__ vox,   // [B,C,D,D,D]
           const float* __restrict__ hocc,  // [B,V,Cocc] or null
           int B, int C, int Cocc, int D, bf16* __restrict__ X) {
  extern __shared__ float tile[];            // [Ct][D+1]
  const int Ct = C + Cocc, J = D / 2 + 1;
  int ih = blockIdx.x % D, id = (blockIdx.x / D) % D, b = blockIdx.x / (D * D);
  const long long V = (long long)D * D * D;
  const long long rowv = ((long long)id * D + ih) * D;
  for (int e = threadIdx.x; e < C * D; e += blockDim.x) {
    int c = e / D, iw = e % D;
    tile[c * (D + 1) + iw] = vox[((long long)b * C + c) * V + rowv + iw];
  }
  for (int e = threadIdx.x; e < Cocc * D; e += blockDim.x) {
    int iw = e / Cocc, c = e % Cocc;
    tile[(C + c) * (D + 1) + iw] = hocc[((long long)b * V + rowv + iw) * Cocc + c];
  }
  __syncthreads();
  int pd = id + 1, ph = ih + 1;
  int jd = pd >> 1, rd = pd & 1, jh = ph >> 1, rh = ph & 1;
  for (int e = threadIdx.x; e < D * Ct; e += blockDim.x) {
    int iw = e / Ct, c = e % Ct;
    int pw = iw + 1, jw = pw >> 1, rw = pw & 1;
    int r = (rd << 2) | (rh << 1) | rw;
    long long dst = ((((long long)b * J + jd) * J + jh) * J + jw) * (8LL * Ct) + (long long)r * Ct + c;
    X[dst] = __float2bfloat16(tile[c * (D + 1) + iw]);
  }
}

// ------------------------------------------------------------------ SIMT GEMM
// out[m, n] = act( sum_k A(m,k) * W[n,k] + bias[n] ),  bf16 operands, fp32 accumulate.
constexpr int BM = 64, BN = 64, BK = 32;

__device__ __forceinline__ long long a_offset(const GemmParams& p, int m, int k) {
  if (p.mode == GEMM_LINEAR) return (long long)m * p.lda + k;
  int Do = p.Do, Js = Do + 1;
  int ow = m % Do, oh = (m / Do) % Do, od = (m / (Do * Do)) % Do, b = m / (Do * Do * Do);
  int a = k / p.Ci8, kc = k - a * p.Ci8;
  int ad = (a >> 2) & 1, ah = (a >> 1) & 1, aw = a & 1;
  return ((((long long)b * Js + od + ad) * Js + oh + ah) * Js + ow + aw) * p.Ci8 + kc;
}

__device__ __forceinline__ void store_out(const GemmParams& p, int m, int n, float v) {
  if (p.relu) v = fmaxf(v, 0.f);
  if (p.out_mode == OUT_F32) {
    reinterpret_cast<float*>(p.out)[(long long)m * p.ldo + p.col_off + n] = v;
  } else if (p.out_mode == OUT_BF16) {
    reinterpret_cast<bf16*>(p.out)[(long long)m * p.ldo + p.col_off + n] = __float2bfloat16(v);
  } else {  // OUT_S2D_BF16: row m = (b,od,oh,ow) of a Do^3 grid -> s2d-padded layout of the next conv
    int Do = p.Do, J = Do / 2 + 1;
    int ow = m % Do, oh = (m / Do) % Do, od = (m / (Do * Do)) % Do, b = m / (Do * Do * Do);
    int pd = od + 1, ph = oh + 1, pw = ow + 1;
    int r = ((pd & 1) << 2) | ((ph & 1) << 1) | (pw & 1);
    long long dst = ((((long long)b * J + (pd >> 1)) * J + (ph >> 1)) * J + (pw >> 1)) * (8LL * p.N) +
                    (long long)r * p.N + n;
    reinterpret_cast<bf16*>(p.out)[dst] = __float2bfloat16(v);
  }
}

struct SimtGroup {
  GemmParams p[3];
};

__global__ void __launch_bounds__(256)
k_gemm_simt(const __grid_constant__ SimtGroup grp) {
  const GemmParams& p = grp.p[blockIdx.z];
  if ((int)(blockIdx.x * BN) >= p.N || (int)(blockIdx.y * BM) >= p.M) return;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % 16, ty = tid / 16;       // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid / 4, lk = (tid % 4) * 8;   // loader: row 0..63, k offset 0,8,16,24
  const bf16* A = reinterpret_cast<const bf16*>(p.A);
  const bf16* W = reinterpret_cast<const bf16*>(p.W);
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    {
      int m = m0 + lr, k = k0 + lk;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (m < p.M && k < p.K) v = __ldg(reinterpret_cast<const uint4*>(A + a_offset(p, m, k)));
      const bf16* h = reinterpret_cast<const bf16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) As[lk + j][lr] = __bfloat162float(h[j]);
    }
    {
      int n = n0 + lr, k = k0 + lk;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (n < p.N && k < p.K) v = __ldg(reinterpret_cast<const uint4*>(W + (long long)n * p.ldw + k));
      const bf16* h = reinterpret_cast<const bf16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) Ws[lk + j][lr] = __bfloat162float(h[j]);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      store_out(p, m, n, acc[i][j] + (p.bias ? p.bias[n] : 0.f));
    }
  }
}

// ------------------------------------------------------------------ trilinear gather (bf16, channels-last)
// same arithmetic as interpolate.cu (reference interpolate_voxel_grid.py:6-59,176-208) on the
// model's internal layouts; `scale` = 1/2 or 1/4 applied as points / (1/scale) like model.py:131,137.
template <bool kS2D>
__global__ void k_interp_cl(const bf16* __restrict__ grid, const float* __restrict__ points,  // [B,3,P]
                            int B, int P, int C, int D, float divisor, bf16* __restrict__ feat,
                            int ldf, int col_off) {
  // one thread per (point, 8 consecutive channels): 16-byte corner loads
  const int C8 = C >> 3;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long NP = (long long)B * P;
  if (i >= NP * C8) return;
  long long n = i / C8;
  int c = (int)(i - n * C8) * 8;
  long long b = n / P, pp = n % P;
  float x = __fdiv_rn(points[(b * 3 + 0) * P + pp], divisor);
  float y = __fdiv_rn(points[(b * 3 + 1) * P + pp], divisor);
  float z = __fdiv_rn(points[(b * 3 + 2) * P + pp], divisor);
  int x0 = static_cast<int>(x), y0 = static_cast<int>(y), z0 = static_cast<int>(z);
  float lx = x - (float)x0, ly = y - (float)y0, lz = z - (float)z0;
  float hx = 1.f - lx, hy = 1.f - ly, hz = 1.f - lz;
  const float w[8] = {hx * hy * hz, lx * hy * hz, hx * ly * hz, hx * hy * lz,
                      lx * ly * hz, hx * ly * lz, lx * hy * lz, lx * ly * lz};
  const int dx[8] = {0, 1, 0, 0, 1, 0, 1, 1};
  const int dy[8] = {0, 0, 1, 0, 1, 1, 0, 1};
  const int dz[8] = {0, 0, 0, 1, 0, 1, 1, 1};
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const int J = D / 2 + 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int ix = x0 + dx[j], iy = y0 + dy[j], iz = z0 + dz[j];
    if (ix < 0 || ix >= D || iy < 0 || iy >= D || iz < 0 || iz >= D) continue;
    long long src;
    if (kS2D) {
      int pd = ix + 1, ph = iy + 1, pw = iz + 1;
      int r = ((pd & 1) << 2) | ((ph & 1) << 1) | (pw & 1);
      src = ((((long long)b * J + (pd >> 1)) * J + (ph >> 1)) * J + (pw >> 1)) * (8LL * C) +
            (long long)r * C + c;
    } else {
      src = (((b * D + ix) * D + iy) * D + iz) * (long long)C + c;
    }
    uint4 v = __ldg(reinterpret_cast<const uint4*>(grid + src));
    const bf16* h = reinterpret_cast<const bf16*>(&v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(w[j], __bfloat162float(h[k]), acc[k]);
  }
  uint4 o;
  __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[0], acc[1]), p1 = __floats2bfloat162_rn(acc[2], acc[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[4], acc[5]), p3 = __floats2bfloat162_rn(acc[6], acc[7]);
  o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
  o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
  *reinterpret_cast<uint4*>(feat + n * ldf + col_off + c) = o;
}


// ---- small row-major grids (conv4 level, 8^3 x 512): shared-memory staged gather -------------
// k_interp_cl re-reads every corner row from L2 (8 corners x P points x C channels: 65 MB for a
// 4 MB grid).  Here a CTA owns (object b, 16 consecutive channels): it stages that 16-channel
// slab of the object's grid (D^3 x 32 B) in shared memory once, then 2 threads per point gather
// the 8 corners x 8 channels each from shared memory (256 points per pass, so the dependent
// point-load -> gather chain is walked only P/256 times).  Same arithmetic, same corner order.
constexpr int kInterpCH = 16;
constexpr int kInterpThreads = 512;
constexpr int kInterpTPP = kInterpCH / 8;            // threads per point
__global__ void __launch_bounds__(kInterpThreads)
k_interp_cl_staged(const bf16* __restrict__ grid, const float* __restrict__ points,  // [B,3,P]
                   int B, int P, int C, int D, float divisor, bf16* __restrict__ feat, int ldf,
                   int col_off) {
  extern __shared__ __align__(16) unsigned char slab_raw[];
  uint4* slab = reinterpret_cast<uint4*>(slab_raw);          // [D^3][kInterpTPP] x 16 B
  const int b = blockIdx.y, c0 = blockIdx.x * kInterpCH;
  const int V = D * D * D;
  const bf16* src = grid + (long long)b * V * C + c0;
  for (int e = threadIdx.x; e < V * kInterpTPP; e += kInterpThreads)
    slab[e] = __ldg(reinterpret_cast<const uint4*>(src + (long long)(e / kInterpTPP) * C) +
                    (e % kInterpTPP));
  __syncthreads();
  const int t = threadIdx.x % kInterpTPP;
  for (int pp = threadIdx.x / kInterpTPP; pp < P; pp += kInterpThreads / kInterpTPP) {
    float x = __fdiv_rn(points[((long long)b * 3 + 0) * P + pp], divisor);
    float y = __fdiv_rn(points[((long long)b * 3 + 1) * P + pp], divisor);
    float z = __fdiv_rn(points[((long long)b * 3 + 2) * P + pp], divisor);
    int x0 = static_cast<int>(x), y0 = static_cast<int>(y), z0 = static_cast<int>(z);
    float lx = x - (float)x0, ly = y - (float)y0, lz = z - (float)z0;
    float hx = 1.f - lx, hy = 1.f - ly, hz = 1.f - lz;
    const float w[8] = {hx * hy * hz, lx * hy * hz, hx * ly * hz, hx * hy * lz,
                        lx * ly * hz, hx * ly * lz, lx * hy * lz, lx * ly * lz};
    const int dx[8] = {0, 1, 0, 0, 1, 0, 1, 1};
    const int dy[8] = {0, 0, 1, 0, 1, 1, 0, 1};
    const int dz[8] = {0, 0, 0, 1, 0, 1, 1, 1};
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int ix = x0 + dx[j], iy = y0 + dy[j], iz = z0 + dz[j];
      if (ix < 0 || ix >= D || iy < 0 || iy >= D || iz < 0 || iz >= D) continue;
      const uint4 v = slab[((ix * D + iy) * D + iz) * kInterpTPP + t];
      const bf16* h = reinterpret_cast<const bf16*>(&v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(w[j], __bfloat162float(h[k]), acc[k]);
    }
    uint4 o;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[0], acc[1]), p1 = __floats2bfloat162_rn(acc[2], acc[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[4], acc[5]), p3 = __floats2bfloat162_rn(acc[6], acc[7]);
    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(feat + ((long long)b * P + pp) * ldf + col_off + c0 + t * 8) = o;
  }
}

// ------------------------------------------------------------------ pose epilogue (model.py:256-273)
__global__ void k_pose(const float* __restrict__ out_rot,    // [B*P, nfg*4]
                       const float* __restrict__ out_trans,  // [B*P, nfg*3]
                       const float* __restrict__ out_conf,   // [B*P, nfg]
                       const float* __restrict__ points,     // [B,3,P] voxel frame
                       const int* __restrict__ class_id, const float* __restrict__ pitch,
                       const float* __restrict__ origin, int B, int P, int nfg,
                       float* __restrict__ rot, float* __restrict__ trans,
                       float* __restrict__ conf) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= (long long)B * P) return;
  long long b = n / P, p = n % P;
  int fg = class_id[b] - 1;
  if (fg < 0) fg += nfg;                  // python-style wrap of the reference's fancy index
  fg = min(max(fg, 0), nfg - 1);
  const float* q = out_rot + n * (nfg * 4) + fg * 4;
  float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + 1e-5f;  // F.normalize
#pragma unroll
  for (int k = 0; k < 4; ++k) rot[n * 4 + k] = q[k] / nrm;
  float pt = pitch[b];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float cam = points[(b * 3 + k) * P + p] * pt + origin[b * 3 + k];
    trans[n * 3 + k] = cam + out_trans[n * (nfg * 3) + fg * 3 + k] * pt;
  }
  float cf = out_conf[n * nfg + fg];
  conf[n] = 1.f / (1.f + expf(-cf));
}


// ---- last head layer fused with the pose epilogue ------------------------------------------
// The reference evaluates conv4_{rot,trans,conf} for all n_fg classes and then keeps the rows of
// the object's class (model.py:249-262).  Only those 4 + 3 + 1 rows are computed here.  A CTA
// handles 32 points of ONE object and stages the 8 weight rows of its class in shared memory
// (fp32, rows padded to 132 floats so the 8 rows fall into different banks); 8 consecutive lanes
// own one point, lane r computing output row r as a K=128 dot product (bf16 operands, fp32
// accumulate, k ascending).  The pose epilogue of k_pose follows after an 8-lane exchange.
constexpr int kH4Pts = 32;
__global__ void __launch_bounds__(256)
k_head4_pose(const bf16* __restrict__ hd3, int ld,              // [B*P, ld]: rot|trans|conf x 128
             const bf16* __restrict__ w_rot, const float* __restrict__ b_rot,      // [nfg*4,128]
             const bf16* __restrict__ w_trans, const float* __restrict__ b_trans,  // [nfg*3,128]
             const bf16* __restrict__ w_conf, const float* __restrict__ b_conf,    // [nfg,128]
             const float* __restrict__ points, const int* __restrict__ class_id,
             const float* __restrict__ pitch, const float* __restrict__ origin, int B, int P,
             int nfg, float* __restrict__ rot, float* __restrict__ trans,
             float* __restrict__ conf, float* __restrict__ raw8) {
  __shared__ __align__(16) float w[8][132];
  __shared__ float bias[8];
  const int b = blockIdx.y;
  const int r = threadIdx.x & 7;
  const int p = blockIdx.x * kH4Pts + (threadIdx.x >> 3);
  const bool ok = p < P;
  const long long n = (long long)b * P + (ok ? p : 0);
  const int seg = r < 4 ? 0 : r < 7 ? 1 : 2;               // rot / trans / conf feature block
  const uint4* hv = reinterpret_cast<const uint4*>(hd3 + n * ld + seg * 128);
  uint4 hu[16];     // the whole 256-byte row segment in flight, under the weight staging below
#pragma unroll
  for (int v = 0; v < 16; ++v) hu[v] = __ldg(hv + v);
  int fg = class_id[b] - 1;
  if (fg < 0) fg += nfg;                  // python-style wrap of the reference's fancy index
  fg = min(max(fg, 0), nfg - 1);
  for (int e = threadIdx.x; e < 8 * 128; e += 256) {
    int rr = e >> 7, k = e & 127;
    const bf16* src = rr < 4 ? w_rot + (size_t)(fg * 4 + rr) * 128
                    : rr < 7 ? w_trans + (size_t)(fg * 3 + rr - 4) * 128
                             : w_conf + (size_t)fg * 128;
    w[rr][k] = __bfloat162float(src[k]);
  }
  if (threadIdx.x < 8) {
    int rr = threadIdx.x;
    bias[rr] = rr < 4 ? b_rot[fg * 4 + rr] : rr < 7 ? b_trans[fg * 3 + rr - 4] : b_conf[fg];
  }
  __syncthreads();
  float acc = 0.f;
#pragma unroll
  for (int v = 0; v < 16; ++v) {
    const uint4 u = hu[v];
    const bf16* hb = reinterpret_cast<const bf16*>(&u);
    const float4 w0 = *reinterpret_cast<const float4*>(&w[r][v * 8]);
    const float4 w1 = *reinterpret_cast<const float4*>(&w[r][v * 8 + 4]);
    acc = fmaf(w0.x, __bfloat162float(hb[0]), acc);
    acc = fmaf(w0.y, __bfloat162float(hb[1]), acc);
    acc = fmaf(w0.z, __bfloat162float(hb[2]), acc);
    acc = fmaf(w0.w, __bfloat162float(hb[3]), acc);
    acc = fmaf(w1.x, __bfloat162float(hb[4]), acc);
    acc = fmaf(w1.y, __bfloat162float(hb[5]), acc);
    acc = fmaf(w1.z, __bfloat162float(hb[6]), acc);
    acc = fmaf(w1.w, __bfloat162float(hb[7]), acc);
  }
  acc += bias[r];
  // gather the point's 8 outputs to every lane of its 8-lane group
  const int base = (threadIdx.x & 31) & ~7;
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = __shfl_sync(0xffffffffu, acc, base + j);
  if (!ok) return;
  if (raw8) raw8[n * 8 + r] = acc;              // training: the 8 selected pre-activation outputs
  if (r == 0) {
    float nrm = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]) + 1e-5f;  // F.normalize
    *reinterpret_cast<float4*>(rot + n * 4) =
        make_float4(o[0] / nrm, o[1] / nrm, o[2] / nrm, o[3] / nrm);
  } else if (r >= 4 && r < 7) {
    const int k = r - 4;
    const float pt = pitch[b];
    float cam = points[((long long)b * 3 + k) * P + p] * pt + origin[b * 3 + k];
    trans[n * 3 + k] = cam + o[r] * pt;
  } else if (r == 7) {
    conf[n] = 1.f / (1.f + expf(-o[7]));
  }
}

}  // namespace mf

using namespace mf;

static int point_mlp(const float* values, const float* points, const float* w1r, const float* b1r,
                     const float* w1p, const float* b1p, const float* w2r, const float* b2r,
                     const float* w2p, const float* b2p, int B, int P, float center, void* feat,
                     int ldf, float* feat2, float* feat1, void* stream_);

extern "C" int mf_cnn_point_mlp(const float* values, const float* points, const float* w1r,
                                const float* b1r, const float* w1p, const float* b1p,
                                const float* w2r, const float* b2r, const float* w2p,
                                const float* b2p, int B, int P, float center, void* feat, int ldf,
                                float* feat2, void* stream_) {
  return point_mlp(values, points, w1r, b1r, w1p, b1p, w2r, b2r, w2p, b2p, B, P, center, feat, ldf,
                   feat2, nullptr, stream_);
}

extern "C" int mf_cnn_point_mlp_f32(const float* values, const float* points, const float* w1r,
                                    const float* b1r, const float* w1p, const float* b1p,
                                    const float* w2r, const float* b2r, const float* w2p,
                                    const float* b2p, int B, int P, float center, void* feat,
                                    int ldf, float* feat2, float* feat1, void* stream_) {
  if (!feat1) return MF_E_BADARG;
  return point_mlp(values, points, w1r, b1r, w1p, b1p, w2r, b2r, w2p, b2p, B, P, center, feat, ldf,
                   feat2, feat1, stream_);
}

static int point_mlp(const float* values, const float* points, const float* w1r, const float* b1r,
                     const float* w1p, const float* b1p, const float* w2r, const float* b2r,
                     const float* w2p, const float* b2p, int B, int P, float center, void* feat,
                     int ldf, float* feat2, float* feat1, void* stream_) {
  if (B <= 0 || P <= 0 || ldf < 216) return MF_E_BADARG;
  if (!values || !points || !w1r || !b1r || !w1p || !b1p || !w2r || !b2r || !w2p || !b2p ||
      !feat || !feat2)
    return MF_E_BADARG;
  long long NP = (long long)B * P;
  MF_ENSURE_DYN_SMEM(k_point_mlp, kMlpSmemFloats * 4);
  PointMlpArgs a{values, points, w1r, b1r, w1p, b1p, w2r, b2r, w2p, b2p, B, P, center, (bf16*)feat,
                 ldf, feat2, feat1};
  k_point_mlp<<<div_up(NP, kMlpPts), 256, kMlpSmemFloats * 4, (cudaStream_t)stream_>>>(a);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_cnn_point_mlp_voxkeys(const float* values, const float* points, const float* w1r,
                                        const float* b1r, const float* w1p, const float* b1p,
                                        const float* w2r, const float* b2r, const float* w2p,
                                        const float* b2p, int B, int P, float center, void* feat,
                                        int ldf, float* feat2, int C, int D, int Ct,
                                        int32_t* prev_keys, void* X, void* stream_) {
  if (B <= 0 || P <= 0 || ldf < 216 || (ldf & 1)) return MF_E_BADARG;
  if (!values || !points || !w1r || !b1r || !w1p || !b1p || !w2r || !b2r || !w2p || !b2p ||
      !feat || !feat2 || !prev_keys || !X)
    return MF_E_BADARG;
  if (C <= 0 || C > 256 || (C & 1) || D <= 0 || (D & 1) || Ct < C) return MF_E_BADARG;
  if ((long long)B * D * D * D >= (1LL << 30)) return MF_E_TOOLARGE;
  int n2 = 32;
  while (n2 < P) n2 <<= 1;
  if (P > 4096 || ((long long)D * D * D + 1) * n2 >= (1LL << 31)) return MF_E_UNSUPPORTED;
  const long long NP = (long long)B * P;
  const int chunks = div_up(P, kRankPts), rank_ctas = B * chunks;
  MF_ENSURE_DYN_SMEM(k_point_mlp_vox, kMlpSmemFloats * 4);
  PointMlpArgs a{values, points, w1r, b1r, w1p, b1p, w2r, b2r, w2p, b2p, B, P, center, (bf16*)feat,
                 ldf, feat2, nullptr};
  k_point_mlp_vox<<<rank_ctas + div_up(NP, kMlpPts), 256, kMlpSmemFloats * 4, (cudaStream_t)stream_>>>(
      a, rank_ctas, D, n2, chunks, prev_keys, prev_keys + NP, (bf16*)X, C, Ct);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_cnn_occ_convs(const float* gne, const float* w1, const float* b1,
                                const float* w2, const float* b2, int B, int D, float* h1,
                                float* h2, void* X, int Ct, int c_off, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (B <= 0 || D <= 0 || (D & 3) || !gne || !w1 || !b1 || !w2 || !b2 || !h1) return MF_E_BADARG;
  if (!h2 && !X) return MF_E_BADARG;
  if (X && ((Ct & 7) || (c_off & 7) || c_off + 16 > Ct)) return MF_E_BADARG;
  long long BV = (long long)B * D * D * D;
  k_occ_conv1<<<div_up(BV, 128), 128, 0, stream>>>(gne, w1, b1, B, D, h1);
  MF_LAUNCH_CHECK();
  k_occ_conv2<<<div_up(BV / 4, 64), 64, 0, stream>>>(h1, w2, b2, B, D, h2, (bf16*)X, Ct, c_off);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

template <typename TIn>
static int occ_convs_tc(const TIn* gne, const float* w1, const float* b1, const float* w2,
                        const float* b2, int B, int D, void* h1_bf16, void* X, int Ct, int c_off,
                        cudaStream_t stream) {
  if (B <= 0 || !gne || !w1 || !b1 || !w2 || !b2 || !h1_bf16 || !X) return MF_E_BADARG;
  if (D != 32) return MF_E_UNSUPPORTED;
  if ((Ct & 7) || (c_off & 7) || c_off + 16 > Ct) return MF_E_BADARG;
  long long BV = (long long)B * D * D * D;
  k_occ_conv1_bf16<TIn><<<div_up(BV, 128), 128, 0, stream>>>(gne, w1, b1, B, D, (bf16*)h1_bf16);
  MF_LAUNCH_CHECK();
  MF_ENSURE_DYN_SMEM(k_occ_conv2_mma, kOccConv2Bytes);
  k_occ_conv2_mma<<<(unsigned)(B * D), 256, kOccConv2Bytes, stream>>>(
      (const bf16*)h1_bf16, w2, b2, B, (bf16*)X, Ct, c_off);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

template <typename TIn>
static int occ_fused(const TIn* gne, const float* w1, const float* b1, const float* w2,
                     const float* b2, int B, int D, void* X, int Ct, int c_off,
                     cudaStream_t stream) {
  if (B <= 0 || !gne || !w1 || !b1 || !w2 || !b2 || !X) return MF_E_BADARG;
  if (D != 32) return MF_E_UNSUPPORTED;
  if ((Ct & 7) || (c_off & 7) || c_off + 16 > Ct) return MF_E_BADARG;
  MF_ENSURE_DYN_SMEM(k_occ_fused<TIn>, occ_fused_bytes<TIn>());
  k_occ_fused<TIn><<<(unsigned)(B * D), 256, occ_fused_bytes<TIn>(), stream>>>(gne, w1, b1, w2, b2, B,
                                                                      (bf16*)X, Ct, c_off);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_cnn_occ_fused(const float* gne, const float* w1, const float* b1,
                                const float* w2, const float* b2, int B, int D, void* X, int Ct,
                                int c_off, void* stream_) {
  return occ_fused<float>(gne, w1, b1, w2, b2, B, D, X, Ct, c_off, (cudaStream_t)stream_);
}

extern "C" int mf_cnn_occ_fused_u8(const uint8_t* gne, const float* w1, const float* b1,
                                   const float* w2, const float* b2, int B, int D, void* X,
                                   int Ct, int c_off, void* stream_) {
  return occ_fused<uint8_t>(gne, w1, b1, w2, b2, B, D, X, Ct, c_off, (cudaStream_t)stream_);
}

extern "C" int mf_cnn_occ_convs_tc(const float* gne, const float* w1, const float* b1,
                                   const float* w2, const float* b2, int B, int D,
                                   void* h1_bf16, void* X, int Ct, int c_off, void* stream_) {
  return occ_convs_tc<float>(gne, w1, b1, w2, b2, B, D, h1_bf16, X, Ct, c_off,
                             (cudaStream_t)stream_);
}

extern "C" int mf_cnn_occ_convs_tc_u8(const uint8_t* gne, const float* w1, const float* b1,
                                      const float* w2, const float* b2, int B, int D,
                                      void* h1_bf16, void* X, int Ct, int c_off, void* stream_) {
  return occ_convs_tc<uint8_t>(gne, w1, b1, w2, b2, B, D, h1_bf16, X, Ct, c_off,
                               (cudaStream_t)stream_);
}

static int voxelize_s2d(const float* feat2, const float* points, int B, int P, int C, int D,
                        int Ct, int32_t* prev_keys, void* X, int phases, cudaStream_t stream) {
  if (B <= 0 || P <= 0 || C <= 0 || C > 256 || (C & 1) || D <= 0 || (D & 1) || Ct < C)
    return MF_E_BADARG;
  if (!points || !prev_keys || !X || ((phases & 2) && !feat2)) return MF_E_BADARG;
  if ((long long)B * D * D * D >= (1LL << 31)) return MF_E_TOOLARGE;
  const int N = B * P;
  // P <= 4096: rank-based path (prev_keys holds 2N ints: keys | sorted order); else search-based
  int n2 = 32;
  while (n2 < P) n2 <<= 1;
  const bool sorted = P <= 4096 && ((long long)D * D * D + 1) * n2 < (1LL << 31);
  int32_t* order = prev_keys + N;
  if (phases & 1) {
    MF_PREFER_MAX_SMEM(k_s2d_clear);
    k_s2d_clear<<<div_up((long long)N * 32, 256), 256, 0, stream>>>(prev_keys, N, C, D, Ct, (bf16*)X);
    MF_LAUNCH_CHECK();
    if (sorted) {
      const int chunks = div_up(P, kRankPts);
      MF_PREFER_MAX_SMEM(k_s2d_rank);
      k_s2d_rank<<<B * chunks, 2 * kRankPts, (size_t)((P + 7) & ~7) * 4, stream>>>(
          points, B, P, D, n2, chunks, prev_keys, order);
    } else {
      const size_t tab = (size_t)(((long long)D * D * D + 31) / 32) * 8;     // two bit tables
      const int use_table = tab <= 200 * 1024;
      MF_ENSURE_DYN_SMEM(k_s2d_keys, 200 * 1024);
      k_s2d_keys<<<B, 1024, use_table ? tab : 0, stream>>>(points, B, P, D, prev_keys, use_table);
    }
    MF_LAUNCH_CHECK();
  }
  if (phases & 2) {
    if (sorted) {
      MF_PREFER_MAX_SMEM(k_s2d_scatter_sorted);
      k_s2d_scatter_sorted<<<div_up(N, kScatWarps), 256, 0, stream>>>(feat2, prev_keys, order, B, P, C,
                                                                     D, Ct, (bf16*)X);
    } else {
      // shared keys: the (at most 8 / P + 2) objects a CTA's 8 consecutive points belong to
      const long long span = ((long long)(kScatWarps - 1) / P + 2) * P;
      if (span * 4 > 160 * 1024) return MF_E_TOOLARGE;
      MF_ENSURE_DYN_SMEM(k_s2d_scatter, 160 * 1024);
      k_s2d_scatter<<<div_up(N, kScatWarps), 256, (size_t)span * 4, stream>>>(feat2, prev_keys, B, P, C,
                                                                             D, Ct, (bf16*)X);
    }
    MF_LAUNCH_CHECK();
  }
  return MF_OK;
}

extern "C" int mf_cnn_voxelize_s2d(const float* feat2, const float* points, int B, int P, int C,
                                   int D, int Ct, int32_t* prev_keys, void* X, void* stream_) {
  return voxelize_s2d(feat2, points, B, P, C, D, Ct, prev_keys, X, 3, (cudaStream_t)stream_);
}

extern "C" int mf_cnn_voxelize_s2d_phase(const float* feat2, const float* points, int B, int P,
                                         int C, int D, int Ct, int32_t* prev_keys, void* X,
                                         int phases, void* stream_) {
  if (phases < 1 || phases > 3) return MF_E_BADARG;
  return voxelize_s2d(feat2, points, B, P, C, D, Ct, prev_keys, X, phases, (cudaStream_t)stream_);
}

extern "C" int mf_cnn_pack_s2d(const float* vox, const float* hocc, int B, int C, int Cocc, int D,
                               void* X, void* stream_) {
  if (B <= 0 || C <= 0 || Cocc < 0 || D <= 0 || (D & 1) || !vox || !X) return MF_E_BADARG;
  if (Cocc > 0 && !hocc) return MF_E_BADARG;
  size_t smem = (size_t)(C + Cocc) * (D + 1) * 4;
  MF_ENSURE_DYN_SMEM(k_pack_s2d, 96 * 1024);
  if (smem > 96 * 1024) return MF_E_UNSUPPORTED;
  k_pack_s2d<<<(unsigned)(B * D * D), 256, smem, (cudaStream_t)stream_>>>(vox, hocc, B, C, Cocc,
                                                                        D, (bf16*)X);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_gemm_bf16_simt_grouped(const GemmParams* hp, int n_groups, void* stream_) {
  if (!hp || n_groups < 1 || n_groups > 3) return MF_E_BADARG;
  SimtGroup grp;
  int maxM = 0, maxN = 0;
  for (int g = 0; g < n_groups; ++g) {
    const GemmParams& p = hp[g];
    if (p.M <= 0 || p.N <= 0 || p.K <= 0 || !p.A || !p.W || !p.out) return MF_E_BADARG;
    if (p.K % 8 != 0 || p.ldw % 8 != 0) return MF_E_UNSUPPORTED;
    if (p.mode == GEMM_LINEAR && p.lda % 8 != 0) return MF_E_UNSUPPORTED;
    if (p.mode == GEMM_CONV_S2D && (p.Ci8 % 8 != 0 || p.K != 8 * p.Ci8)) return MF_E_UNSUPPORTED;
    grp.p[g] = p;
    maxM = p.M > maxM ? p.M : maxM;
    maxN = p.N > maxN ? p.N : maxN;
  }
  for (int g = n_groups; g < 3; ++g) grp.p[g] = hp[0];
  dim3 grid(div_up(maxN, BN), div_up(maxM, BM), n_groups);
  k_gemm_simt<<<grid, 256, 0, (cudaStream_t)stream_>>>(grp);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_gemm_bf16_simt(const GemmParams* hp, void* stream_) {
  return mf_gemm_bf16_simt_grouped(hp, 1, stream_);
}

extern "C" int mf_cnn_interp_cl(const void* grid, int s2d, const float* points, int B, int P,
                                int C, int D, float divisor, void* feat, int ldf, int col_off,
                                void* stream_) {
  if (B <= 0 || P <= 0 || C <= 0 || D <= 0 || !grid || !points || !feat) return MF_E_BADARG;
  if ((C & 7) || (ldf & 7) || (col_off & 7)) return MF_E_UNSUPPORTED;
  long long tot = (long long)B * P * (C / 8);
  const size_t slab_bytes = (size_t)D * D * D * kInterpCH * 2;
  if (!s2d && C % kInterpCH == 0 && slab_bytes <= 48 * 1024 && B <= 65535) {
    // small grid: stage a 16-channel slab per CTA, read the grid from L2 once
    dim3 g((unsigned)(C / kInterpCH), (unsigned)B);
    MF_PREFER_MAX_SMEM(k_interp_cl_staged);
    k_interp_cl_staged<<<g, kInterpThreads, slab_bytes, (cudaStream_t)stream_>>>(
        (const bf16*)grid, points, B, P, C, D, divisor, (bf16*)feat, ldf, col_off);
    MF_LAUNCH_CHECK();
    return MF_OK;
  }
  if (s2d) {
    MF_PREFER_MAX_SMEM(k_interp_cl<true>);
    k_interp_cl<true><<<div_up(tot, 256), 256, 0, (cudaStream_t)stream_>>>(
        (const bf16*)grid, points, B, P, C, D, divisor, (bf16*)feat, ldf, col_off);
  } else {
    MF_PREFER_MAX_SMEM(k_interp_cl<false>);
    k_interp_cl<false><<<div_up(tot, 256), 256, 0, (cudaStream_t)stream_>>>(
        (const bf16*)grid, points, B, P, C, D, divisor, (bf16*)feat, ldf, col_off);
  }
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_cnn_pose(const float* out_rot, const float* out_trans, const float* out_conf,
                           const float* points, const int32_t* class_id, const float* pitch,
                           const float* origin, int B, int P, int nfg, float* rot, float* trans,
                           float* conf, void* stream_) {
  if (B <= 0 || P <= 0 || nfg <= 0) return MF_E_BADARG;
  if (!out_rot || !out_trans || !out_conf || !points || !class_id || !pitch || !origin || !rot ||
      !trans || !conf)
    return MF_E_BADARG;
  MF_PREFER_MAX_SMEM(k_pose);
  k_pose<<<div_up((long long)B * P, 128), 128, 0, (cudaStream_t)stream_>>>(
      out_rot, out_trans, out_conf, points, class_id, pitch, origin, B, P, nfg, rot, trans, conf);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

static int head4_pose(const void* hd3, int ld, const void* w_rot, const float* b_rot,
                      const void* w_trans, const float* b_trans, const void* w_conf,
                      const float* b_conf, const float* points, const int32_t* class_id,
                      const float* pitch, const float* origin, int B, int P, int nfg, float* rot,
                      float* trans, float* conf, float* raw8, void* stream_);

extern "C" int mf_cnn_head4_pose(const void* hd3, int ld, const void* w_rot, const float* b_rot,
                                 const void* w_trans, const float* b_trans, const void* w_conf,
                                 const float* b_conf, const float* points,
                                 const int32_t* class_id, const float* pitch, const float* origin,
                                 int B, int P, int nfg, float* rot, float* trans, float* conf,
                                 void* stream_) {
  return head4_pose(hd3, ld, w_rot, b_rot, w_trans, b_trans, w_conf, b_conf, points, class_id,
                    pitch, origin, B, P, nfg, rot, trans, conf, nullptr, stream_);
}

extern "C" int mf_cnn_head4_pose_train(const void* hd3, int ld, const void* w_rot,
                                       const float* b_rot, const void* w_trans,
                                       const float* b_trans, const void* w_conf,
                                       const float* b_conf, const float* points,
                                       const int32_t* class_id, const float* pitch,
                                       const float* origin, int B, int P, int nfg, float* rot,
                                       float* trans, float* conf, float* raw8, void* stream_) {
  if (!raw8) return MF_E_BADARG;
  return head4_pose(hd3, ld, w_rot, b_rot, w_trans, b_trans, w_conf, b_conf, points, class_id,
                    pitch, origin, B, P, nfg, rot, trans, conf, raw8, stream_);
}

static int head4_pose(const void* hd3, int ld, const void* w_rot, const float* b_rot,
                      const void* w_trans, const float* b_trans, const void* w_conf,
                      const float* b_conf, const float* points, const int32_t* class_id,
                      const float* pitch, const float* origin, int B, int P, int nfg, float* rot,
                      float* trans, float* conf, float* raw8, void* stream_) {
  if (B <= 0 || P <= 0 || nfg <= 0 || ld < 384 || (ld & 7) || B > 65535) return MF_E_BADARG;
  if (!hd3 || !w_rot || !b_rot || !w_trans || !b_trans || !w_conf || !b_conf || !points ||
      !class_id || !pitch || !origin || !rot || !trans || !conf)
    return MF_E_BADARG;
  dim3 grid((unsigned)div_up(P, kH4Pts), (unsigned)B);
  MF_PREFER_MAX_SMEM(k_head4_pose);
  k_head4_pose<<<grid, 256, 0, (cudaStream_t)stream_>>>(
      (const bf16*)hd3, ld, (const bf16*)w_rot, b_rot, (const bf16*)w_trans, b_trans,
      (const bf16*)w_conf, b_conf, points, class_id, pitch, origin, B, P, nfg, rot, trans, conf,
      raw8);
  MF_LAUNCH_CHECK();
  return MF_OK;
}
