// Training-side tcgen05 GEMMs of the singleview_3d pose model (sm_100a): the reductions over
// rows / voxels that the backward pass of
//   morefusion/contrib/singleview_3d/models/model.py:93-141 (conv3, conv4), :239-254 (heads)
// needs and that chainer obtains from cuDNN's backward-data / backward-filter algorithms.
//
//   TN     dW[n, k]        = sum_m dZ[m, n] * X[m, k]                 (Conv1D weight gradients)
//   WGRAD  dW[co, (a, c)]  = sum_{b,o} dYp[b, o + 1, co] * Xs2d[b, o + a, c]
//                                                     (k4 s2 Conv3D weight gradient, s2d form)
//   DGRAD  dX_r[(b,o), ci] = sum_{a, co} dYp[b, o + 2 - r - a, co] * Wd_r[ci, (a, co)]
//                            for each of the 8 input-voxel parities r (k4 s2 Conv3D input
//                            gradient: per parity a k2 s1 correlation over the padded dY)
//
// Same persistent warp-specialised structure as conv3d_tc.cu (TMA producer / MMA issuer / TMEM
// double buffer / 8 epilogue warps).  What is new:
//   * MN-major operands.  In TN and WGRAD the contraction index (m, or the output voxel) is the
//     SLOW dimension of both operands as they sit in HBM.  A TMA box {64 elements along the row,
//     64 rows} with SWIZZLE_128B lands in shared memory as the canonical MN-major SW128 atom of
//     tcgen05 (cute/atom/mma_traits_sm100.hpp: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte
//     units): 128-byte rows = 64 MN elements, 8 rows per 1024-byte swizzle group (SBO), the
//     next 64 MN elements in the next box (LBO = 8192 B).  Instruction-descriptor bits 15 / 16
//     select MN-major A / B (cute/arch/mma_sm100_desc.hpp).  No transposed copies anywhere.
//   * the 5-D voxel boxes of the convolutions as either operand, with per-tap coordinate offsets
//     that depend on the tap (wgrad) or on tap and parity (dgrad).
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mf {

constexpr int TG_BLOCK_M = 128;
constexpr int TG_BLOCK_K = 64;
constexpr int TG_A_BYTES = TG_BLOCK_M * TG_BLOCK_K * 2;   // 16 KiB
constexpr int TG_THREADS = 384;
constexpr int TG_MAX_GROUPS = 8;

enum { TG_TN = 0, TG_WGRAD = 1, TG_DGRAD = 2 };
enum { TE_F32 = 0, TE_PARITY_ADD_F32 = 1, TE_BF16 = 2 };

struct TgArgs {
  CUtensorMap tmA[TG_MAX_GROUPS];
  CUtensorMap tmB[TG_MAX_GROUPS];
  int mode, epi;
  int M, N;                       // logical output rows / columns per group
  int kb_total, kb_per_split, splitk;
  int m_tiles, n_tiles, groups, n_units;
  int Do;                         // conv output edge (8 / 16)
  int kb_per_batch;               // WGRAD: K blocks (64 voxels) per batch element
  int kb_per_a;                   // DGRAD: K blocks per tap (= Co / 64)
  int C8;                         // WGRAD: channels per voxel of the s2d input (N = 8 * C8)
  int atomic;                     // TE_F32: accumulate with atomics (split-K / accumulate)
  float* out_f32;
  bf16* out_bf16;
  long long ldo;
  long long out_group_stride;     // elements between the outputs of consecutive groups
  int* err;
};

// MN-major SWIZZLE_128B operand descriptor: LBO = 8192 B between 64-element MN atoms (one TMA
// box each), SBO = 1024 B between 8-row K groups, version 1, layout type 2.
__device__ __forceinline__ uint64_t make_sw128_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tg_unit(const TgArgs& a, int u, int& g, int& mt, int& nt, int& split) {
  split = u % a.splitk;
  int r = u / a.splitk;
  nt = r % a.n_tiles;
  r /= a.n_tiles;
  mt = r % a.m_tiles;
  g = r / a.m_tiles;
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(TG_THREADS, 1)
k_gemm_train(const __grid_constant__ TgArgs args) {
  constexpr int B_BYTES = BLOCK_N * TG_BLOCK_K * 2;
  constexpr int STAGE_BYTES = TG_A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool mn_major = args.mode != TG_DGRAD;

  if (threadIdx.x == 0) {
    for (int g = 0; g < args.groups; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.tmA[g])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.tmB[g])) : "memory");
    }
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 8);
    mbar_init(&tmem_empty_bar[1], 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)(2 * BLOCK_N))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer
    int kc = 0;
    for (int u = blockIdx.x; u < args.n_units; u += gridDim.x) {
      int g, mt, nt, split;
      tg_unit(args, u, g, mt, nt, split);
      const int m0 = mt * TG_BLOCK_M, n0 = nt * BLOCK_N;
      const int kb0 = split * args.kb_per_split;
      const int kb1 = min(kb0 + args.kb_per_split, args.kb_total);
      const int Do = args.Do;
      int cb = 0, cw = 0, ch = 0, cd = 0;
      if (args.mode == TG_DGRAD) {
        cw = m0 % Do;
        ch = (m0 / Do) % Do;
        cd = (m0 / (Do * Do)) % Do;
        cb = m0 / (Do * Do * Do);
      }
      for (int kb = kb0; kb < kb1; ++kb, ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, args.err, 1, 64);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
        unsigned char* sb = sa + TG_A_BYTES;
        if (args.mode == TG_TN) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_2d(sa + j * 8192, &args.tmA[g], &full_bar[s], m0 + 64 * j, kb * 64);
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_2d(sb + j * 8192, &args.tmB[g], &full_bar[s], n0 + 64 * j, kb * 64);
        } else if (args.mode == TG_WGRAD) {
          // K block = 64 consecutive output voxels of one batch element: 16 x 4 (Do = 16) or
          // 8 x 8 (Do = 8) voxels of one d-slice
          const int b = kb / args.kb_per_batch, r = kb - b * args.kb_per_batch;
          const int d0 = (Do == 16) ? (r >> 2) : r, h0 = (Do == 16) ? ((r & 3) * 4) : 0;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_5d(sa + j * 8192, &args.tmA[g], &full_bar[s], m0 + 64 * j, 1, 1 + h0, 1 + d0, b);
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) {
            const int n = n0 + 64 * j;
            const int a = n / args.C8, c = n - a * args.C8;
            tma_load_5d(sb + j * 8192, &args.tmB[g], &full_bar[s], c, (a & 1), ((a >> 1) & 1) + h0,
                        ((a >> 2) & 1) + d0, b);
          }
        } else {
          const int a = kb / args.kb_per_a, c = (kb - a * args.kb_per_a) * TG_BLOCK_K;
          const int rw = g & 1, rh = (g >> 1) & 1, rd = (g >> 2) & 1;
          tma_load_5d(sa, &args.tmA[g], &full_bar[s], c, cw + 2 - rw - (a & 1),
                      ch + 2 - rh - ((a >> 1) & 1), cd + 2 - rd - ((a >> 2) & 1), cb);
          tma_load_2d(sb, &args.tmB[g], &full_bar[s], kb * TG_BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) |
                           ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(TG_BLOCK_M >> 4) << 24);
    int kc = 0, it = 0;
    for (int u = blockIdx.x; u < args.n_units; u += gridDim.x, ++it) {
      int g, mt, nt, split;
      tg_unit(args, u, g, mt, nt, split);
      const int kb0 = split * args.kb_per_split;
      const int kb1 = min(kb0 + args.kb_per_split, args.kb_total);
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], aph ^ 1, args.err, 4, 32);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = kb0; kb < kb1; ++kb, ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(&full_bar[s], ph, args.err, 2, 20);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        if (mn_major) {
          const uint64_t adesc = make_sw128_desc_mn(sa);
          const uint64_t bdesc = make_sw128_desc_mn(sa + TG_A_BYTES);
#pragma unroll
          for (int k = 0; k < TG_BLOCK_K / 16; ++k)      // 16 K rows = 2048 B = 128 (>>4) per step
            umma_bf16(tmem_d, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
        } else {
          const uint64_t adesc = make_sw128_desc(sa);
          const uint64_t bdesc = make_sw128_desc(sa + TG_A_BYTES);
#pragma unroll
          for (int k = 0; k < TG_BLOCK_K / 16; ++k)
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tmem_full_bar[acc]);
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp w reads TMEM lane quarter w % 4 and the column half (w - 4) / 4
    const int q = warp & 3, hsel = (warp - 4) >> 2;
    constexpr int EPI_COLS = BLOCK_N / 2;
    int it = 0;
    for (int u = blockIdx.x; u < args.n_units; u += gridDim.x, ++it) {
      int g, mt, nt, split;
      tg_unit(args, u, g, mt, nt, split);
      const int m0 = mt * TG_BLOCK_M, n0 = nt * BLOCK_N + hsel * EPI_COLS;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], aph, args.err, 3, 128);
      tcgen05_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < args.M;
      const uint32_t tacc = tmem_base + (uint32_t)(acc * BLOCK_N + hsel * EPI_COLS) +
                            ((uint32_t)(q * 32) << 16);
      long long roff = 0;
      if (row_ok) {
        if (args.epi == TE_PARITY_ADD_F32) {
          // row m = (b, od, oh, ow) of parity g -> input voxel 2 o + 1 - r of the (2 Do)^3 grid
          const int Do = args.Do, Di = 2 * Do;
          const int ow = m % Do, oh = (m / Do) % Do, od = (m / (Do * Do)) % Do, b = m / (Do * Do * Do);
          const int x = 2 * od + 1 - ((g >> 2) & 1), y = 2 * oh + 1 - ((g >> 1) & 1),
                    z = 2 * ow + 1 - (g & 1);
          roff = ((((long long)b * Di + x) * Di + y) * Di + z) * args.ldo;
        } else {
          roff = (long long)g * args.out_group_stride + (long long)m * args.ldo;
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tacc + (uint32_t)c0, r);
        const int n = n0 + c0;
        if (!row_ok || n >= args.N) continue;
        const bool full = n + 32 <= args.N;
        if (args.epi == TE_BF16) {
          bf16* dst = args.out_bf16 + roff + n;
          if (full) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 uu;
              uu.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
              uu.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
              uu.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
              uu.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
              reinterpret_cast<uint4*>(dst)[j] = uu;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < args.N) dst[j] = __float2bfloat16(__uint_as_float(r[j]));
          }
          continue;
        }
        float* dst = args.out_f32 + roff + n;
        if (args.epi == TE_PARITY_ADD_F32) {
          // every element belongs to exactly one (parity, tile): plain read-modify-write
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = reinterpret_cast<float4*>(dst)[j];
            o.x += __uint_as_float(r[4 * j + 0]);
            o.y += __uint_as_float(r[4 * j + 1]);
            o.z += __uint_as_float(r[4 * j + 2]);
            o.w += __uint_as_float(r[4 * j + 3]);
            reinterpret_cast<float4*>(dst)[j] = o;
          }
        } else if (full && !(args.ldo & 3)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 o = make_float4(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1]),
                                         __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            if (args.atomic) atomicAdd(reinterpret_cast<float4*>(dst) + j, o);
            else reinterpret_cast<float4*>(dst)[j] = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n + j < args.N) {
              if (args.atomic) atomicAdd(dst + j, __uint_as_float(r[j]));
              else dst[j] = __uint_as_float(r[j]);
            }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc]))
                     : "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(2 * BLOCK_N))
                 : "memory");
  }
}

// ------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled tg_get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

static int tg_encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  PFN_encodeTiled fn = tg_get_encode();
  if (!fn) return MF_E_UNSUPPORTED;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MF_OK : MF_E_BADARG;
}

// 5-D map over a channels-last voxel tensor [B, J, J, J, C] with a box of `nvox` consecutive
// output voxels (w fastest) x 64 channels
static int tg_encode_vox(CUtensorMap* tm, const void* base, int Bn, int J, int C, int Do, int nvox) {
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)J, (cuuint64_t)J, (cuuint64_t)J, (cuuint64_t)Bn};
  cuuint64_t str[4] = {(cuuint64_t)C * 2, (cuuint64_t)J * C * 2, (cuuint64_t)J * J * C * 2,
                       (cuuint64_t)J * J * J * C * 2};
  const int rows = nvox / Do;                      // h (and d) extent of the box
  cuuint32_t box[5] = {64, (cuuint32_t)Do, (cuuint32_t)(rows > Do ? Do : rows),
                       (cuuint32_t)(rows > Do ? rows / Do : 1), 1};
  return tg_encode(tm, base, 5, dims, str, box);
}

template <int BLOCK_N, int STAGES>
static int tg_launch(const TgArgs& args, cudaStream_t stream) {
  constexpr int smem = STAGES * (TG_A_BYTES + BLOCK_N * TG_BLOCK_K * 2) + 1024;
  MF_ENSURE_DYN_SMEM((k_gemm_train<BLOCK_N, STAGES>), smem);
  int n_sm = 148, dev = 0;
  MF_CUDA_TRY(cudaGetDevice(&dev));
  MF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const int grid = args.n_units < n_sm ? args.n_units : n_sm;
  k_gemm_train<BLOCK_N, STAGES><<<grid, TG_THREADS, smem, stream>>>(args);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

static int tg_finish(TgArgs& a, int BN, int want_units, cudaStream_t stream) {
  const bool overwrite = (a.epi == TE_F32) && !a.atomic;
  a.m_tiles = (a.M + TG_BLOCK_M - 1) / TG_BLOCK_M;
  a.n_tiles = (a.N + BN - 1) / BN;
  const int tiles = a.m_tiles * a.n_tiles * a.groups;
  int splitk = 1;
  if (want_units > 0 && a.epi == TE_F32 && tiles < want_units) {
    splitk = want_units / tiles;
    if (splitk > a.kb_total / 8) splitk = a.kb_total / 8;    // >= 8 K blocks per unit
    if (splitk < 1) splitk = 1;
  }
  a.kb_per_split = (a.kb_total + splitk - 1) / splitk;
  a.splitk = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  if (a.splitk > 1) {
    a.atomic = 1;
    if (overwrite)        // split-K accumulates with atomics: "overwrite" needs a zeroed output
      for (int g = 0; g < a.groups; ++g)
        MF_CUDA_TRY(cudaMemset2DAsync(a.out_f32 + (long long)g * a.out_group_stride,
                                      (size_t)a.ldo * 4, 0, (size_t)a.N * 4, (size_t)a.M, stream));
  }
  a.n_units = tiles * a.splitk;
  a.err = nullptr;
  return BN == 256 ? tg_launch<256, 4>(a, stream) : tg_launch<128, 6>(a, stream);
}

}  // namespace mf

using namespace mf;

extern "C" int mf_train_gemm_tn(const void* dZ, long long ldz, const void* X, long long ldx,
                                int m_rows, int n_out, int k_in, float* out, long long ldo,
                                int n_groups, long long dz_group_cols, long long x_group_cols,
                                long long out_group_stride, int accumulate, void* stream_) {
  if (!dZ || !X || !out || m_rows <= 0 || n_out <= 0 || k_in <= 0 || n_groups < 1 ||
      n_groups > TG_MAX_GROUPS)
    return MF_E_BADARG;
  if ((ldz & 7) || (ldx & 7) || (dz_group_cols & 7) || (x_group_cols & 7)) return MF_E_UNSUPPORTED;
  if (((uintptr_t)dZ & 15) || ((uintptr_t)X & 15) || ((uintptr_t)out & 15)) return MF_E_UNSUPPORTED;
  TgArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TG_TN; a.epi = TE_F32; a.atomic = accumulate ? 1 : 0;
  a.M = n_out; a.N = k_in; a.groups = n_groups;
  a.kb_total = (m_rows + TG_BLOCK_K - 1) / TG_BLOCK_K;
  a.out_f32 = out; a.ldo = ldo; a.out_group_stride = out_group_stride;
  const int BN = k_in > 128 ? 256 : 128;
  for (int g = 0; g < n_groups; ++g) {
    const bf16* pa = reinterpret_cast<const bf16*>(dZ) + g * dz_group_cols;
    const bf16* pb = reinterpret_cast<const bf16*>(X) + g * x_group_cols;
    cuuint64_t da[2] = {(cuuint64_t)n_out, (cuuint64_t)m_rows}, sa[1] = {(cuuint64_t)ldz * 2};
    cuuint64_t db[2] = {(cuuint64_t)k_in, (cuuint64_t)m_rows}, sb[1] = {(cuuint64_t)ldx * 2};
    cuuint32_t box[2] = {64, 64};
    int rc = tg_encode(&a.tmA[g], pa, 2, da, sa, box);
    if (rc) return rc;
    rc = tg_encode(&a.tmB[g], pb, 2, db, sb, box);
    if (rc) return rc;
  }
  return tg_finish(a, BN, 148, (cudaStream_t)stream_);
}

extern "C" int mf_train_conv_wgrad(const void* dYp, const void* Xs2d, int batch, int Do, int Co,
                                   int C8, float* out, int accumulate, void* stream_) {
  if (!dYp || !Xs2d || !out || batch <= 0) return MF_E_BADARG;
  if (!(Do == 8 || Do == 16) || Co % 128 != 0 || C8 % 64 != 0 || (8 * C8) % 256 != 0)
    return MF_E_UNSUPPORTED;
  TgArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TG_WGRAD; a.epi = TE_F32; a.atomic = accumulate ? 1 : 0;
  a.M = Co; a.N = 8 * C8; a.groups = 1; a.Do = Do; a.C8 = C8;
  a.kb_per_batch = Do * Do * Do / 64;
  a.kb_total = batch * a.kb_per_batch;
  a.out_f32 = out; a.ldo = 8LL * C8; a.out_group_stride = 0;
  int rc = tg_encode_vox(&a.tmA[0], dYp, batch, Do + 2, Co, Do, 64);
  if (rc) return rc;
  rc = tg_encode_vox(&a.tmB[0], Xs2d, batch, Do + 1, C8, Do, 64);
  if (rc) return rc;
  return tg_finish(a, 256, 4 * 148, (cudaStream_t)stream_);
}

extern "C" int mf_train_conv_dgrad(const void* dYp, const void* Wd, int batch, int Do, int Co,
                                   int Ci, int epilogue, void* out, long long ldo,
                                   long long out_group_stride, void* stream_) {
  if (!dYp || !Wd || !out || batch <= 0) return MF_E_BADARG;
  if (!(Do == 8 || Do == 16) || Co % 64 != 0 || Ci % 8 != 0 || Ci > 256) return MF_E_UNSUPPORTED;
  if (epilogue != TE_PARITY_ADD_F32 && epilogue != TE_BF16) return MF_E_BADARG;
  if (epilogue == TE_PARITY_ADD_F32 && (Ci % 32 != 0 || (ldo & 3))) return MF_E_UNSUPPORTED;
  if (epilogue == TE_BF16 && (ldo & 7)) return MF_E_UNSUPPORTED;
  TgArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TG_DGRAD; a.epi = epilogue;
  a.M = batch * Do * Do * Do; a.N = Ci; a.groups = 8; a.Do = Do;
  a.kb_per_a = Co / 64;
  a.kb_total = 8 * a.kb_per_a;
  a.ldo = ldo; a.out_group_stride = out_group_stride;
  if (epilogue == TE_BF16) a.out_bf16 = (bf16*)out; else a.out_f32 = (float*)out;
  const int BN = Ci > 128 ? 256 : 128;
  for (int g = 0; g < 8; ++g) {
    int rc = tg_encode_vox(&a.tmA[g], dYp, batch, Do + 2, Co, Do, 128);
    if (rc) return rc;
    const bf16* w = reinterpret_cast<const bf16*>(Wd) + (long long)g * Ci * 8 * Co;
    cuuint64_t dw[2] = {(cuuint64_t)(8 * Co), (cuuint64_t)Ci}, sw[1] = {(cuuint64_t)(8 * Co) * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)BN};
    rc = tg_encode(&a.tmB[g], w, 2, dw, sw, box);
    if (rc) return rc;
  }
  return tg_finish(a, BN, 0, (cudaStream_t)stream_);
}
