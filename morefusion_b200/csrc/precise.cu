// fp32-class ("bf16x3") mode of the 3-D section of the pose model: the parity mode.
//
// The reference runs model.py:93-141,:239-273 in fp32 (cuDNN).  The throughput path stores bf16
// activations / weights (8 mantissa bits): poses agree with the fp32 oracle to ~1e-2 only.  Here
// every GEMM operand is split x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits
// together) and every product A.W is evaluated on the same tcgen05 kernels as
//     A_hi.W_hi + A_lo.W_hi + A_hi.W_lo          (fp32 accumulation; lo.lo ~ 2^-18 dropped)
// as three GEMMs into fp32 slices; the kernels below add the slices, apply bias / ReLU in fp32 and
// split the result again for the next layer.  Gathers and the pose epilogue read hi + lo.
#include <cuda_bf16.h>

#include "cnn.cuh"
#include "common.cuh"

namespace mf {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void split2(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16(v);
  lo = __float2bfloat16(v - __bfloat162float(hi));
}

// fp32 rows -> hi / lo bf16 rows (per-point features)
__global__ void k_px_split(const float* __restrict__ src, long long lds, long long M, int N,
                           bf16* __restrict__ hi, bf16* __restrict__ lo, long long ldd, int col_off) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const long long m = i / N;
  const int n = (int)(i - m * N);
  bf16 h, l;
  split2(src[m * lds + n], h, l);
  hi[m * ldd + col_off + n] = h;
  lo[m * ldd + col_off + n] = l;
}

// out = relu?(ws0 + ws1 + ws2 + bias) -> hi / lo, row-major or scattered into the s2d layout of
// the next convolution (cnn.cu).  4 columns per thread.
__global__ void k_px_combine(const float* __restrict__ ws, int n_slices, long long M, int N,
                             const float* __restrict__ bias, int relu, int out_mode, int Do,
                             bf16* __restrict__ hi, bf16* __restrict__ lo, long long ldo, int col_off) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long MN = M * N;
  if (i * 4 >= MN) return;
  const long long m = (i * 4) / N;
  const int n = (int)((i * 4) - m * N);
  float4 v = bias ? __ldg(reinterpret_cast<const float4*>(bias + n)) : make_float4(0, 0, 0, 0);
  // smallest terms first: (A_hi W_lo + A_lo W_hi) + A_hi W_hi
  float4 s = make_float4(0, 0, 0, 0);
  for (int k = n_slices - 1; k >= 1; --k) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(ws + k * MN) + i);
    s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
  }
  const float4 a0 = __ldcs(reinterpret_cast<const float4*>(ws) + i);
  v.x += s.x + a0.x; v.y += s.y + a0.y; v.z += s.z + a0.z; v.w += s.w + a0.w;
  if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  long long off;
  if (out_mode == OUT_S2D_BF16) {
    const int J = Do / 2 + 1;
    const int ow = (int)(m % Do), oh = (int)((m / Do) % Do), od = (int)((m / (Do * Do)) % Do);
    const long long b = m / ((long long)Do * Do * Do);
    const int pd = od + 1, ph = oh + 1, pw = ow + 1;
    const int r = ((pd & 1) << 2) | ((ph & 1) << 1) | (pw & 1);
    off = ((((b * J + (pd >> 1)) * J + (ph >> 1)) * J + (pw >> 1)) * 8LL + r) * N + n;
  } else {
    off = m * ldo + col_off + n;
  }
  bf16 h[4], l[4];
  split2(v.x, h[0], l[0]); split2(v.y, h[1], l[1]); split2(v.z, h[2], l[2]); split2(v.w, h[3], l[3]);
  *reinterpret_cast<uint2*>(hi + off) = *reinterpret_cast<uint2*>(h);
  *reinterpret_cast<uint2*>(lo + off) = *reinterpret_cast<uint2*>(l);
}

// dense s2d packing with hi / lo outputs (cnn.cu k_pack_s2d, fp32 in)
__global__ void __launch_bounds__(256)
k_px_pack_s2d(const float* __restrict__ vox, const float* __restrict__ hocc, int B, int C, int Cocc,
              int D, bf16* __restrict__ Xh, bf16* __restrict__ Xl) {
  extern __shared__ float tile[];            // [Ct][D+1]
  const int Ct = C + Cocc, J = D / 2 + 1;
  const int ih = blockIdx.x % D, id = (blockIdx.x / D) % D, b = blockIdx.x / (D * D);
  const long long V = (long long)D * D * D;
  const long long rowv = ((long long)id * D + ih) * D;
  for (int e = threadIdx.x; e < C * D; e += blockDim.x) {
    const int c = e / D, iw = e % D;
    tile[c * (D + 1) + iw] = vox[((long long)b * C + c) * V + rowv + iw];
  }
  for (int e = threadIdx.x; e < Cocc * D; e += blockDim.x) {
    const int iw = e / Cocc, c = e % Cocc;
    tile[(C + c) * (D + 1) + iw] = hocc[((long long)b * V + rowv + iw) * Cocc + c];
  }
  __syncthreads();
  const int pd = id + 1, ph = ih + 1;
  const int jd = pd >> 1, rd = pd & 1, jh = ph >> 1, rh = ph & 1;
  for (int e = threadIdx.x; e < D * Ct; e += blockDim.x) {
    const int iw = e / Ct, c = e % Ct;
    const int pw = iw + 1, jw = pw >> 1, rw = pw & 1;
    const int r = (rd << 2) | (rh << 1) | rw;
    const long long dst = ((((long long)b * J + jd) * J + jh) * J + jw) * (8LL * Ct) + (long long)r * Ct + c;
    bf16 h, l;
    split2(tile[c * (D + 1) + iw], h, l);
    Xh[dst] = h;
    Xl[dst] = l;
  }
}

// trilinear gather from a hi / lo channels-last grid (cnn.cu k_interp_cl arithmetic on hi + lo)
template <bool kS2D>
__global__ void k_px_interp(const bf16* __restrict__ gh, const bf16* __restrict__ gl,
                            const float* __restrict__ points, int B, int P, int C, int D,
                            float divisor, bf16* __restrict__ fh, bf16* __restrict__ fl, int ldf,
                            int col_off) {
  const int C8 = C >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long NP = (long long)B * P;
  if (i >= NP * C8) return;
  const long long n = i / C8;
  const int c = (int)(i - n * C8) * 8;
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
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const int J = D / 2 + 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ix = x0 + dx[j], iy = y0 + dy[j], iz = z0 + dz[j];
    if (ix < 0 || ix >= D || iy < 0 || iy >= D || iz < 0 || iz >= D) continue;
    long long src;
    if (kS2D) {
      const int pd = ix + 1, ph = iy + 1, pw = iz + 1;
      const int r = ((pd & 1) << 2) | ((ph & 1) << 1) | (pw & 1);
      src = ((((long long)b * J + (pd >> 1)) * J + (ph >> 1)) * J + (pw >> 1)) * (8LL * C) +
            (long long)r * C + c;
    } else {
      src = (((b * D + ix) * D + iy) * D + iz) * (long long)C + c;
    }
    const uint4 vh = __ldg(reinterpret_cast<const uint4*>(gh + src));
    const uint4 vl = __ldg(reinterpret_cast<const uint4*>(gl + src));
    const bf16* eh = reinterpret_cast<const bf16*>(&vh);
    const bf16* el = reinterpret_cast<const bf16*>(&vl);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      acc[k] = fmaf(w[j], __bfloat162float(eh[k]) + __bfloat162float(el[k]), acc[k]);
  }
  bf16 h[8], l[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) split2(acc[k], h[k], l[k]);
  *reinterpret_cast<uint4*>(fh + n * ldf + col_off + c) = *reinterpret_cast<uint4*>(h);
  *reinterpret_cast<uint4*>(fl + n * ldf + col_off + c) = *reinterpret_cast<uint4*>(l);
}

// last head layer + class selection + pose epilogue (cnn.cu k_head4_pose) with fp32 weights and
// hi + lo activations; one 8-lane group per point, lane r = output row r
__global__ void __launch_bounds__(256)
k_px_head4_pose(const bf16* __restrict__ h3h, const bf16* __restrict__ h3l, int ld,
                const float* __restrict__ w_rot, const float* __restrict__ b_rot,
                const float* __restrict__ w_trans, const float* __restrict__ b_trans,
                const float* __restrict__ w_conf, const float* __restrict__ b_conf,
                const float* __restrict__ points, const int* __restrict__ class_id,
                const float* __restrict__ pitch, const float* __restrict__ origin, int B, int P,
                int nfg, float* __restrict__ rot, float* __restrict__ trans,
                float* __restrict__ conf) {
  __shared__ __align__(16) float w[8][132];
  __shared__ float bias[8];
  const int b = blockIdx.y;
  int fg = class_id[b] - 1;
  if (fg < 0) fg += nfg;
  fg = min(max(fg, 0), nfg - 1);
  for (int e = threadIdx.x; e < 8 * 128; e += 256) {
    const int r = e >> 7, k = e & 127;
    const float* src = r < 4 ? w_rot + (size_t)(fg * 4 + r) * 128
                     : r < 7 ? w_trans + (size_t)(fg * 3 + r - 4) * 128
                             : w_conf + (size_t)fg * 128;
    w[r][k] = src[k];
  }
  if (threadIdx.x < 8) {
    const int r = threadIdx.x;
    bias[r] = r < 4 ? b_rot[fg * 4 + r] : r < 7 ? b_trans[fg * 3 + r - 4] : b_conf[fg];
  }
  __syncthreads();
  const int r = threadIdx.x & 7;
  const int p = blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool ok = p < P;
  const long long n = (long long)b * P + (ok ? p : 0);
  const int seg = r < 4 ? 0 : r < 7 ? 1 : 2;
  const bf16* hh = h3h + n * ld + seg * 128;
  const bf16* hl = h3l + n * ld + seg * 128;
  float acc = 0.f;
  for (int k = 0; k < 128; ++k)
    acc = fmaf(w[r][k], __bfloat162float(hh[k]) + __bfloat162float(hl[k]), acc);
  acc += bias[r];
  const int base = (threadIdx.x & 31) & ~7;
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = __shfl_sync(0xffffffffu, acc, base + j);
  if (!ok) return;
  if (r == 0) {
    const float nrm = sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]) + 1e-5f;
    *reinterpret_cast<float4*>(rot + n * 4) = make_float4(o[0] / nrm, o[1] / nrm, o[2] / nrm, o[3] / nrm);
  } else if (r >= 4 && r < 7) {
    const int k = r - 4;
    const float pt = pitch[b];
    const float cam = points[((long long)b * 3 + k) * P + p] * pt + origin[b * 3 + k];
    trans[n * 3 + k] = cam + o[r] * pt;
  } else if (r == 7) {
    conf[n] = 1.f / (1.f + expf(-o[7]));
  }
}

}  // namespace mf

using namespace mf;

extern "C" int mf_px_split(const float* src, long long lds, long long M, int N, void* hi, void* lo,
                           long long ldd, int col_off, void* stream_) {
  if (!src || !hi || !lo || M <= 0 || N <= 0) return MF_E_BADARG;
  k_px_split<<<div_up(M * N, 256), 256, 0, (cudaStream_t)stream_>>>(src, lds, M, N, (bf16*)hi,
                                                                    (bf16*)lo, ldd, col_off);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_px_combine(const float* ws, int n_slices, long long M, int N, const float* bias,
                             int relu, int out_mode, int Do, void* hi, void* lo, long long ldo,
                             int col_off, void* stream_) {
  if (!ws || !hi || !lo || n_slices < 1 || M <= 0 || N <= 0 || (N & 3) || (col_off & 3) || (ldo & 3))
    return MF_E_BADARG;
  if (out_mode != OUT_BF16 && out_mode != OUT_S2D_BF16) return MF_E_BADARG;
  k_px_combine<<<div_up(M * N / 4, 256), 256, 0, (cudaStream_t)stream_>>>(
      ws, n_slices, M, N, bias, relu, out_mode, Do, (bf16*)hi, (bf16*)lo, ldo, col_off);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_px_pack_s2d(const float* vox, const float* hocc, int B, int C, int Cocc, int D,
                              void* Xh, void* Xl, void* stream_) {
  if (B <= 0 || C <= 0 || Cocc < 0 || D <= 0 || (D & 1) || !vox || !Xh || !Xl) return MF_E_BADARG;
  if (Cocc > 0 && !hocc) return MF_E_BADARG;
  const size_t smem = (size_t)(C + Cocc) * (D + 1) * 4;
  if (smem > 96 * 1024) return MF_E_UNSUPPORTED;
  MF_ENSURE_DYN_SMEM(k_px_pack_s2d, 96 * 1024);
  k_px_pack_s2d<<<(unsigned)(B * D * D), 256, smem, (cudaStream_t)stream_>>>(vox, hocc, B, C, Cocc,
                                                                           D, (bf16*)Xh, (bf16*)Xl);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_px_interp(const void* gh, const void* gl, int s2d, const float* points, int B,
                            int P, int C, int D, float divisor, void* fh, void* fl, int ldf,
                            int col_off, void* stream_) {
  if (B <= 0 || P <= 0 || C <= 0 || D <= 0 || !gh || !gl || !points || !fh || !fl) return MF_E_BADARG;
  if ((C & 7) || (ldf & 7) || (col_off & 7)) return MF_E_UNSUPPORTED;
  const long long tot = (long long)B * P * (C / 8);
  if (s2d)
    k_px_interp<true><<<div_up(tot, 256), 256, 0, (cudaStream_t)stream_>>>(
        (const bf16*)gh, (const bf16*)gl, points, B, P, C, D, divisor, (bf16*)fh, (bf16*)fl, ldf, col_off);
  else
    k_px_interp<false><<<div_up(tot, 256), 256, 0, (cudaStream_t)stream_>>>(
        (const bf16*)gh, (const bf16*)gl, points, B, P, C, D, divisor, (bf16*)fh, (bf16*)fl, ldf, col_off);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_px_head4_pose(const void* h3h, const void* h3l, int ld, const float* w_rot,
                                const float* b_rot, const float* w_trans, const float* b_trans,
                                const float* w_conf, const float* b_conf, const float* points,
                                const int32_t* class_id, const float* pitch, const float* origin,
                                int B, int P, int nfg, float* rot, float* trans, float* conf,
                                void* stream_) {
  if (B <= 0 || P <= 0 || nfg <= 0 || ld < 384 || B > 65535) return MF_E_BADARG;
  if (!h3h || !h3l || !w_rot || !b_rot || !w_trans || !b_trans || !w_conf || !b_conf || !points ||
      !class_id || !pitch || !origin || !rot || !trans || !conf)
    return MF_E_BADARG;
  dim3 grid((unsigned)div_up(P, 32), (unsigned)B);
  k_px_head4_pose<<<grid, 256, 0, (cudaStream_t)stream_>>>(
      (const bf16*)h3h, (const bf16*)h3l, ld, w_rot, b_rot, w_trans, b_trans, w_conf, b_conf, points,
      class_id, pitch, origin, B, P, nfg, rot, trans, conf);
  MF_LAUNCH_CHECK();
  return MF_OK;
}
