// tcgen05 / TMEM / TMA GEMM for sm_100a: the implicit-GEMM k4-s2 Conv3Ds (conv3, conv4) and the
// Conv1D heads of the singleview_3d pose model.
//
// Replaces the cuDNN ConvolutionND calls behind
//   morefusion/contrib/singleview_3d/models/model.py:73-74 (conv3, conv4), :77-91 (heads).
//
// out[m, n] = act( sum_k A(m,k) * W[n,k] + bias[n] )     bf16 operands, fp32 accumulation in TMEM.
//
// CTA tile 128 x BLOCK_N x 64.  Warp roles (one CTA per SM, 256 threads):
//   warp 0  (1 lane)  TMA producer: per K block one box of A and one of W into a
//                     SWIZZLE_128B smem stage, mbarrier expect_tx / complete_tx
//   warp 1  (1 lane)  MMA issuer: 4 x tcgen05.mma (M128 x N x K16) per stage, tcgen05.commit
//                     releases the stage; a final commit signals the epilogue
//   warp 2            TMEM allocator (BLOCK_N fp32 columns)
//   warps 4-7         epilogue: tcgen05.ld (32 lanes x 32 columns per warp), bias, ReLU, bf16
//                     pack, 16-byte stores (row-major, or scattered into the space-to-depth
//                     layout the next convolution's TMA boxes read)
// For the convolutions the A operand needs no im2col: with the space-to-depth input layout
// (cnn.cu) the 128 output voxels x 64 channels of one K block are ONE 5-D TMA box.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "cnn.cuh"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace mf {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;   // 16 KiB

struct TcExtra {
  int kb_per_a;      // conv: K blocks per kernel-offset a (= Ci8 / 64)
  int kb_total;      // K blocks overall
  int kb_per_split;  // K blocks per split-K slice
  int splitk;
  float* ws;         // fp32 [splitk][M][N] partial sums when splitk > 1 (plain stores, no atomics)
  int* err;          // device error word (pipeline time-out), may be null
};

constexpr int TC_MAX_GROUPS = 3;
// up to 3 same-shape GEMMs in one launch (the three pose heads): grid.z = group * splitk + split
struct TcArgs {
  CUtensorMap tmA[TC_MAX_GROUPS];
  CUtensorMap tmW[TC_MAX_GROUPS];
  GemmParams p[TC_MAX_GROUPS];
  TcExtra e;
};

__device__ __forceinline__ long long out_row_offset(const GemmParams& p, int m) {
  if (p.out_mode != OUT_S2D_BF16) return (long long)m * p.ldo + p.col_off;
  int Do = p.Do, J = Do / 2 + 1;
  int ow = m % Do, oh = (m / Do) % Do, od = (m / (Do * Do)) % Do, b = m / (Do * Do * Do);
  int pd = od + 1, ph = oh + 1, pw = ow + 1;
  int r = ((pd & 1) << 2) | ((ph & 1) << 1) | (pw & 1);
  return ((((long long)b * J + (pd >> 1)) * J + (ph >> 1)) * J + (pw >> 1)) * (8LL * p.N) +
         (long long)r * p.N;
}

// ------------------------------------------------------------------ kernel
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(256, 1)
k_gemm_tc(const __grid_constant__ TcArgs args) {
  const int group = blockIdx.z / args.e.splitk, split = blockIdx.z - group * args.e.splitk;
  const CUtensorMap& tmA = args.tmA[group];
  const CUtensorMap& tmW = args.tmW[group];
  const GemmParams& p = args.p[group];
  const TcExtra& e = args.e;
  constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B operands need 1024-byte alignment
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N, m0 = blockIdx.y * TC_BLOCK_M;
  const int kb0 = split * e.kb_per_split;
  const int kb1 = min(kb0 + e.kb_per_split, e.kb_total);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)BLOCK_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer
    int cb = 0, cw = 0, ch = 0, cd = 0;
    if (p.mode == GEMM_CONV_S2D) {
      int Do = p.Do;
      cw = m0 % Do;                       // 0 for Do in {8,16}: tiles start at ow = 0
      ch = (m0 / Do) % Do;
      cd = (m0 / (Do * Do)) % Do;
      cb = m0 / (Do * Do * Do);
    }
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1, e.err, 1);
      mbar_expect_tx(&full_bar[s], STAGE_BYTES);
      unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
      unsigned char* sb = sa + TC_A_BYTES;
      const int kb = kb0 + i;
      if (p.mode == GEMM_CONV_S2D) {
        int a = kb / e.kb_per_a, c = (kb - a * e.kb_per_a) * TC_BLOCK_K;
        tma_load_5d(sa, &tmA, &full_bar[s], c, cw + (a & 1), ch + ((a >> 1) & 1),
                    cd + ((a >> 2) & 1), cb);
      } else {
        tma_load_2d(sa, &tmA, &full_bar[s], kb * TC_BLOCK_K, m0);
      }
      tma_load_2d(sb, &tmW, &full_bar[s], kb * TC_BLOCK_K, n0);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1,
    // b=BF16 [10,13)=1, K-major both (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                               ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(TC_BLOCK_M >> 4) << 24);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full_bar[s], ph, e.err, 2);
      tcgen05_fence_after();
      const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
      const uint64_t adesc = make_sw128_desc(sa);
      const uint64_t bdesc = make_sw128_desc(sa + TC_A_BYTES);
#pragma unroll
      for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
        // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in addr>>4 units
        umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                  (i > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs retire
    }
    umma_commit(&tmem_full_bar);    // accumulator complete
  } else if (warp >= 4) {
    // ===== epilogue
    mbar_wait(&tmem_full_bar, 0, e.err, 3);
    tcgen05_fence_after();
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    const long long roff = row_ok ? out_row_offset(p, m) : 0;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      const int n = n0 + c0;
      if (!row_ok || n >= p.N) continue;
      if (p.N & 31) {
        // ragged N (small-N heads, fp32 row-major output only; enforced on the host): rows are
        // not 16-byte aligned, scalar stores
        float* dst = reinterpret_cast<float*>(p.out) + roff + n;
        const int nj = min(32, p.N - n);
#pragma unroll
        for (int j = 0; j < 32; ++j) {      // static indices: keeps r[] in registers
          if (j < nj) {
            float x = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + n + j) : 0.f);
            dst[j] = p.relu ? fmaxf(x, 0.f) : x;
          }
        }
        continue;
      }
      if (e.splitk > 1) {
        float4* dst = reinterpret_cast<float4*>(e.ws + ((long long)split * p.M + m) * p.N + n);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                               __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        continue;
      }
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + n + j) : 0.f);
        v[j] = p.relu ? fmaxf(x, 0.f) : x;
      }
      if (p.out_mode == OUT_F32) {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + roff + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + roff + n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0);
          u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2);
          u.w = *reinterpret_cast<uint32_t*>(&h3);
          dst[j] = u;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)BLOCK_N)
                 : "memory");
  }
}

// ------------------------------------------------------------------ persistent variant
// One CTA per SM loops over work units (group, m-tile, n-tile, k-split).  The shared-memory
// stage ring runs continuously across units and the accumulator is double-buffered in TMEM
// (2 x BLOCK_N columns), so the epilogue of unit i (tcgen05.ld, bias/ReLU, stores) overlaps the
// TMA + MMA main loop of unit i+1 and the per-CTA set-up (barrier init, TMEM allocation,
// descriptor prefetch, pipeline fill) is paid once per SM instead of once per tile -- which is
// what the short-K head GEMMs (2-16 K blocks per tile) are dominated by in the one-shot kernel.
struct TcSched {
  int m_tiles, n_tiles, groups, splitk, n_units;
  long long* stamps;      // profiling hook (mf_gemm_bf16_tc_ex): clock64 per role per unit, CTA 0
  // stream-K (one group, splitk == 1): the tiles x K-blocks iteration space is cut into gridDim.x
  // equal contiguous ranges, so every SM does the same number of K blocks whatever the tile
  // count (256 tiles on 148 SMs would otherwise leave 13.5 % of the machine idle in the second
  // wave).  A CTA whose range starts inside a tile stores that segment's raw fp32 accumulator in
  // its workspace slot and raises its flags; the CTA that holds the tile's first K block owns
  // the tile and adds the slots of the CTAs after it, in CTA order, in its epilogue.
  int streamk;
  int* sk_flags;          // [gridDim.x][8] (one word per epilogue warp), zero between launches
  float* sk_ws;           // [gridDim.x][128 x BLOCK_N] fp32
};
#define TC_STAMP(slot)                                                        \
  do {                                                                        \
    if (sc.stamps && blockIdx.x == 0 && it < 16) sc.stamps[it * 8 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ void unit_decode(const TcSched& sc, int u, int& g, int& mt, int& nt,
                                            int& split) {
  split = u % sc.splitk;
  int r = u / sc.splitk;
  nt = r % sc.n_tiles;
  r /= sc.n_tiles;
  mt = r % sc.m_tiles;
  g = r / sc.m_tiles;
}

// One piece of work of a persistent CTA: K blocks [kb0, kb1) of unit u.
struct TcSeg {
  int u, kb0, kb1;
  int partial;            // stream-K: store the raw accumulator to this CTA's slot (not the output)
  int peer0, n_peers;     // stream-K owner: add the slots of CTAs peer0 .. peer0 + n_peers - 1
};

struct TcSegIter {
  long long pos, hi, W;   // stream-K: position in the tiles x K-blocks space
  int u;                  // otherwise: next unit
  __device__ __forceinline__ static long long lo_of(long long W, int c) {
    return W * c / (int)gridDim.x;
  }
  __device__ __forceinline__ TcSegIter(const TcSched& sc, int kb_total) {
    W = (long long)sc.n_units * kb_total;
    pos = sc.streamk ? lo_of(W, blockIdx.x) : 0;
    hi = sc.streamk ? lo_of(W, blockIdx.x + 1) : 0;
    u = blockIdx.x;
  }
  __device__ __forceinline__ bool next(const TcSched& sc, const TcExtra& e, TcSeg& s) {
    s.partial = 0; s.peer0 = 0; s.n_peers = 0;
    if (sc.streamk) {
      if (pos >= hi) return false;
      const int t = (int)(pos / e.kb_total);
      const int k0 = (int)(pos - (long long)t * e.kb_total);
      const long long room = hi - pos;
      const int k1 = room < (long long)(e.kb_total - k0) ? k0 + (int)room : e.kb_total;
      s.u = t; s.kb0 = k0; s.kb1 = k1;
      s.partial = k0 != 0;
      if (k0 == 0 && k1 < e.kb_total) {
        const long long tile_end = (long long)(t + 1) * e.kb_total;
        s.peer0 = blockIdx.x + 1;
        int c = s.peer0;
        while (c < (int)gridDim.x && lo_of(W, c) < tile_end) ++c;
        s.n_peers = c - s.peer0;
      }
      pos += k1 - k0;
      return true;
    }
    if (u >= sc.n_units) return false;
    const int split = u % sc.splitk;
    s.u = u;
    s.kb0 = split * e.kb_per_split;
    s.kb1 = min(s.kb0 + e.kb_per_split, e.kb_total);
    u += gridDim.x;
    return true;
  }
};

constexpr int TC_P_THREADS = 384;        // warps 0-3: TMA / MMA / TMEM alloc / idle, 4-11: epilogue
constexpr int TC_STG_BYTES = 32 * 128;   // per-epilogue-warp staging tile

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(TC_P_THREADS, 1)
k_gemm_tc_persistent(const __grid_constant__ TcArgs args, const TcSched sc) {
  constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* out_stage = smem + (size_t)STAGES * STAGE_BYTES;   // 8 x TC_STG_BYTES
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcExtra& e = args.e;

  if (threadIdx.x == 0) {
    for (int g = 0; g < sc.groups; ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.tmA[g])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.tmW[g])) : "memory");
    }
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 8);       // one arrival per epilogue warp
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
    int kc = 0;                                        // K blocks issued so far (ring position)
    int it = 0;
    TcSegIter iter(sc, e.kb_total);
    TcSeg seg;
    for (; iter.next(sc, e, seg); ++it) {
      int g, mt, nt, split;
      unit_decode(sc, seg.u, g, mt, nt, split);
      const GemmParams& p = args.p[g];
      TC_STAMP(0);
      const int m0 = mt * TC_BLOCK_M, n0 = nt * BLOCK_N;
      const int kb0 = seg.kb0, kb1 = seg.kb1;
      int cb = 0, cw = 0, ch = 0, cd = 0;
      if (p.mode == GEMM_CONV_S2D) {
        int Do = p.Do;
        cw = m0 % Do;
        ch = (m0 / Do) % Do;
        cd = (m0 / (Do * Do)) % Do;
        cb = m0 / (Do * Do * Do);
      }
      for (int kb = kb0; kb < kb1; ++kb, ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, e.err, 1, 64);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
        unsigned char* sb = sa + TC_A_BYTES;
        if (p.mode == GEMM_CONV_S2D) {
          int a = kb / e.kb_per_a, c = (kb - a * e.kb_per_a) * TC_BLOCK_K;
          tma_load_5d(sa, &args.tmA[g], &full_bar[s], c, cw + (a & 1), ch + ((a >> 1) & 1),
                      cd + ((a >> 2) & 1), cb);
        } else {
          tma_load_2d(sa, &args.tmA[g], &full_bar[s], kb * TC_BLOCK_K, m0);
        }
        tma_load_2d(sb, &args.tmW[g], &full_bar[s], kb * TC_BLOCK_K, n0);
      }
      TC_STAMP(1);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                               ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(TC_BLOCK_M >> 4) << 24);
    int kc = 0, it = 0;
    TcSegIter iter(sc, e.kb_total);
    TcSeg seg;
    for (; iter.next(sc, e, seg); ++it) {
      const int kb0 = seg.kb0, kb1 = seg.kb1;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], aph ^ 1, e.err, 4, 32);    // epilogue has drained this buffer
      tcgen05_fence_after();
      TC_STAMP(2);
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = kb0; kb < kb1; ++kb, ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(&full_bar[s], ph, e.err, 2, 20);
        tcgen05_fence_after();
        if (kb == kb0) TC_STAMP(3);
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t adesc = make_sw128_desc(sa);
        const uint64_t bdesc = make_sw128_desc(sa + TC_A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BLOCK_K / 16; ++k)
          umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                    (kb > kb0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tmem_full_bar[acc]);
      TC_STAMP(4);
    }
  } else if (warp >= 4) {
    // ===== epilogue: 8 warps.  Warp w reads TMEM lane quarter w%4 (hardware rule) and the
    // column half (w-4)/4 of the tile, so a tile is drained by two warps per lane quarter.
    const int q = warp & 3, hsel = (warp - 4) >> 2;
    constexpr int EPI_COLS = BLOCK_N / 2, NCH = EPI_COLS / 32;
    const uint32_t stg_s = smem_u32(out_stage + (size_t)(warp - 4) * TC_STG_BYTES);  // 32 rows x 128 B
    int it = 0;
    TcSegIter iter(sc, e.kb_total);
    TcSeg seg;
    // stream-K slots: this warp's 32 rows x EPI_COLS columns, stored chunk-major / lane-minor so
    // that a warp-wide float4 access is 512 contiguous bytes
    const int wslot = warp - 4;
    constexpr long long SLOT_FLOATS = (long long)TC_BLOCK_M * BLOCK_N;
    for (; iter.next(sc, e, seg); ++it) {
      int g, mt, nt, split;
      unit_decode(sc, seg.u, g, mt, nt, split);
      const GemmParams& p = args.p[g];
      const int m0 = mt * TC_BLOCK_M, n0 = nt * BLOCK_N + hsel * EPI_COLS;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], aph, e.err, 3, 128);
      tcgen05_fence_after();
      if (threadIdx.x == 128) TC_STAMP(5);
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      const long long roff = row_ok ? out_row_offset(p, m) : 0;
      const uint32_t tacc = tmem_base + (uint32_t)(acc * BLOCK_N + hsel * EPI_COLS) +
                            ((uint32_t)(q * 32) << 16);
      const bool coalesced_bf16 = !(p.N & 31) && e.splitk == 1 && p.out_mode != OUT_F32;
      if (seg.partial) {
        // stream-K: this CTA holds a later K range of a tile another CTA owns
        float4* slot = reinterpret_cast<float4*>(sc.sk_ws + (long long)blockIdx.x * SLOT_FLOATS +
                                                 (long long)wslot * 32 * EPI_COLS);
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tacc + (uint32_t)(c * 32), r);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            slot[(c * 8 + j) * 32 + lane] =
                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
        __threadfence();
        __syncwarp();
        if (lane == 0)
          asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(sc.sk_flags + blockIdx.x * 8 + wslot),
                       "r"(1) : "memory");
      } else if (coalesced_bf16) {
        if (seg.n_peers > 0) {
          // wait for the same-numbered epilogue warp of every peer (they ran these K blocks at the
          // START of their range, this CTA reaches the tile at the END of its own: normally no wait)
          if (lane == 0) {
            for (int pi = 0; pi < seg.n_peers; ++pi) {
              const int* f = sc.sk_flags + (seg.peer0 + pi) * 8 + wslot;
              int v = 0;
              for (long long spin = 0; spin < (1LL << 24); ++spin) {
                asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if (v) break;
                __nanosleep(64);
              }
              if (!v && e.err) atomicExch(e.err, 7);
            }
          }
          __syncwarp();
        }
        // bf16 outputs.  Chunks of 32 columns: the tcgen05.ld of chunk c+1 is in flight while
        // chunk c gets bias/ReLU/convert.  Two chunks (64 columns = 128 B per row) are staged in
        // this warp's XOR-swizzled shared-memory tile and written out with row-contiguous
        // 16-byte stores: one instruction covers 4 rows x 128 B (whole lines) instead of 32
        // rows x 16 B.
        const unsigned okmask = __ballot_sync(0xffffffffu, row_ok);
        // p = args.p[g] is an indexed constant-bank access (long latency).  Pin the fields the
        // hot loop needs in registers; the empty asm makes them opaque so the compiler cannot
        // rematerialise them with another indexed LDC in front of every store.
        int relu_i = p.relu, Nn = p.N;
        const float* bias = p.bias;
        bf16* outp = reinterpret_cast<bf16*>(p.out);
        asm volatile("" : "+r"(relu_i), "+r"(Nn), "+l"(bias), "+l"(outp));
        const bool relu = relu_i != 0;
        long long rofs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rofs[i] = __shfl_sync(0xffffffffu, roff, i * 4 + (lane >> 3));
        uint32_t ra[32], rb[32];
        auto process = [&](uint32_t (&r)[32], int c) {
          const int n = n0 + c * 32;
          if (n < Nn) {
            for (int pi = 0; pi < seg.n_peers; ++pi) {      // stream-K: later K ranges of this tile
              const float4* ps = reinterpret_cast<const float4*>(
                  sc.sk_ws + (long long)(seg.peer0 + pi) * SLOT_FLOATS + (long long)wslot * 32 * EPI_COLS);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 a = __ldcg(ps + (c * 8 + j) * 32 + lane);
                r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + a.x);
                r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + a.y);
                r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + a.z);
                r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + a.w);
              }
            }
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 bq = bias ? __ldg(reinterpret_cast<const float4*>(bias + n) + j)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bq.x;
              v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bq.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bq.z;
              v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bq.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 uu;
              if (relu) {       // ReLU folded into the conversion (cvt.rn.relu)
                uu.x = pack_bf16x2_relu(v[8 * j + 0], v[8 * j + 1]);
                uu.y = pack_bf16x2_relu(v[8 * j + 2], v[8 * j + 3]);
                uu.z = pack_bf16x2_relu(v[8 * j + 4], v[8 * j + 5]);
                uu.w = pack_bf16x2_relu(v[8 * j + 6], v[8 * j + 7]);
              } else {
                uu.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                uu.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                uu.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                uu.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              }
              const int piece = ((c & 1) * 4 + j) ^ (lane & 7);
              sts_v4(stg_s + lane * 128 + piece * 16, uu);
            }
          }
          if (c & 1) {
            // flush the 64 staged columns [n0 + (c-1)*32, +64)
            __syncwarp();
            const int ncol = n0 + (c - 1) * 32 + (lane & 7) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + (lane >> 3);
              const uint4 uu = lds_v4(stg_s + row * 128 + (((lane & 7) ^ (row & 7)) * 16));
              if (((okmask >> row) & 1u) && ncol < Nn)
                *reinterpret_cast<uint4*>(outp + rofs[i] + ncol) = uu;
            }
            __syncwarp();
          }
        };
        tmem_ld_32x32_issue(tacc, ra);
        tmem_ld_wait(ra);
#pragma unroll
        for (int c = 0; c < NCH; c += 2) {
          tmem_ld_32x32_issue(tacc + (uint32_t)((c + 1) * 32), rb);
          process(ra, c);
          tmem_ld_wait(rb);
          if (c + 2 < NCH) tmem_ld_32x32_issue(tacc + (uint32_t)((c + 2) * 32), ra);
          process(rb, c + 1);
          if (c + 2 < NCH) tmem_ld_wait(ra);
        }
        if (seg.n_peers > 0) {                  // slots consumed: re-arm the flags for the next launch
          __syncwarp();
          if (lane == 0)
            for (int pi = 0; pi < seg.n_peers; ++pi) sc.sk_flags[(seg.peer0 + pi) * 8 + wslot] = 0;
        }
      } else {
        int pN = p.N, pM = p.M, prelu = p.relu, pmode = p.out_mode;
        const float* pbias = p.bias;
        void* pout = p.out;
        asm volatile("" : "+r"(pN), "+r"(pM), "+r"(prelu), "+r"(pmode), "+l"(pbias), "+l"(pout));
#pragma unroll 1
        for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tacc + (uint32_t)c0, r);
          const int n = n0 + c0;
          if (!row_ok || n >= pN) continue;
          if (pN & 31) {
            // ragged N (small-N heads, fp32 row-major output only; enforced on the host)
            float* dst = reinterpret_cast<float*>(pout) + roff + n;
            const int nj = min(32, pN - n);
#pragma unroll
            for (int j = 0; j < 32; ++j) {      // static indices: keeps r[] in registers
              if (j < nj) {
                float x = __uint_as_float(r[j]) + (pbias ? __ldg(pbias + n + j) : 0.f);
                dst[j] = prelu ? fmaxf(x, 0.f) : x;
              }
            }
            continue;
          }
          if (e.splitk > 1) {
            float4* dst = reinterpret_cast<float4*>(e.ws + ((long long)split * pM + m) * pN + n);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            continue;
          }
          // fp32 row-major output with bias / ReLU
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(pout) + roff + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 bq = pbias ? __ldg(reinterpret_cast<const float4*>(pbias + n) + j)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o = make_float4(__uint_as_float(r[4 * j + 0]) + bq.x,
                                   __uint_as_float(r[4 * j + 1]) + bq.y,
                                   __uint_as_float(r[4 * j + 2]) + bq.z,
                                   __uint_as_float(r[4 * j + 3]) + bq.w);
            if (prelu) {
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
            dst[j] = o;
          }
        }
      }
      if (threadIdx.x == 128) TC_STAMP(6);
      // this warp is done reading the accumulator buffer: hand it back to the MMA issuer
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc]))
                     : "memory");
      if (threadIdx.x == 128) TC_STAMP(7);
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

// ------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) compute one 256 x 256 tile: each CTA stages its
// own 128 rows of A and HALF of the B tile (128 of the 256 weight rows) per K block - 32 KB per
// CTA per K block instead of 48 KB - and the leader CTA's single thread issues one
// tcgen05.mma.cta_group::2 (M = 256) that reads both CTAs' shared memory and writes each CTA's
// 128 accumulator rows into that CTA's own TMEM.  The single-CTA kernel is bound by operand
// delivery (4 stages x 48 KB in flight against ~1.7 us of TMA latency, measured 0.46-0.58 us per
// K block against 0.37 us of tensor work); here the same shared memory holds 6 stages and every
// stage carries 1.5x the tensor work.
//   full_bar   : leader only; its producer arms 2 x 32 KB, BOTH CTAs' TMA loads credit it
//   empty_bar  : per CTA; released by the leader's multicast commit (both CTAs' producers wake)
//   tmem_full  : per CTA; multicast commit
//   tmem_empty : leader only; 16 arrivals = 8 epilogue warps of each CTA (the peer arrives remotely)
// Scheduling: units are 256-row tiles dealt to CLUSTERS; stream-K as in the single-CTA kernel with
// cluster ranges, slots and flags per CTA (each CTA parks / adds its own 128-row half).
struct TcPairIter {
  // stream-K hybrid: the first dp_tiles = floor(tiles / clusters) x clusters tiles are dealt whole
  // (one per cluster and wave: no partial sums, and the two K halves of a tile are not fetched
  // 100 us apart), only the remaining tiles x K-blocks space is cut into equal ranges
  long long pos, hi, W;
  int u, ncl, cl, dp_tiles;
  __device__ __forceinline__ long long lo_of(int c) const { return W * c / ncl; }
  __device__ __forceinline__ TcPairIter(const TcSched& sc, int kb_total) {
    ncl = (int)gridDim.x >> 1;
    cl = (int)blockIdx.x >> 1;
    dp_tiles = sc.streamk ? (sc.n_units / ncl) * ncl : sc.n_units;
    W = (long long)(sc.n_units - dp_tiles) * kb_total;
    pos = lo_of(cl);
    hi = lo_of(cl + 1);
    u = cl;
  }
  __device__ __forceinline__ bool next(const TcSched& sc, const TcExtra& e, TcSeg& s) {
    s.partial = 0; s.peer0 = 0; s.n_peers = 0;
    if (u < dp_tiles) {
      s.u = u; s.kb0 = 0; s.kb1 = e.kb_total;
      u += ncl;
      return true;
    }
    if (pos >= hi) return false;
    const int t = (int)(pos / e.kb_total);
    const int k0 = (int)(pos - (long long)t * e.kb_total);
    const long long room = hi - pos;
    const int k1 = room < (long long)(e.kb_total - k0) ? k0 + (int)room : e.kb_total;
    s.u = dp_tiles + t; s.kb0 = k0; s.kb1 = k1;
    s.partial = k0 != 0;
    if (k0 == 0 && k1 < e.kb_total) {
      const long long tile_end = (long long)(t + 1) * e.kb_total;
      s.peer0 = cl + 1;
      int c = s.peer0;
      while (c < ncl && lo_of(c) < tile_end) ++c;
      s.n_peers = c - s.peer0;
    }
    pos += k1 - k0;
    return true;
  }
};

template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_P_THREADS, 1)
k_gemm_tc_pair(const __grid_constant__ TcArgs args, const TcSched sc) {
  constexpr int BLOCK_N = 256;                        // N of the pair tile; each CTA stages 128 rows
  constexpr int B_BYTES = (BLOCK_N / 2) * TC_BLOCK_K * 2;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;   // 32 KB
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* out_stage = smem + (size_t)STAGES * STAGE_BYTES;   // 8 x TC_STG_BYTES
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcExtra& e = args.e;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.tmA[0])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.tmW[0])) : "memory");
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar[0], 1);
    mbar_init(&tmem_full_bar[1], 1);
    mbar_init(&tmem_empty_bar[0], 16);      // 8 epilogue warps of each CTA of the pair
    mbar_init(&tmem_empty_bar[1], 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)(2 * BLOCK_N))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer's barriers exist before anything signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const GemmParams& p = args.p[0];

  if (warp == 0 && lane == 0) {
    // ===== TMA producer (both CTAs: own A rows, own half of the B tile)
    int kc = 0;
    TcPairIter iter(sc, e.kb_total);
    TcSeg seg;
    // linear A is streamed (each byte feeds the n_tiles of one row tile): let it leave L2 first;
    // the weights every cluster re-reads and the stream-K slots must survive the pass
    const uint64_t pol_a = l2_policy_evict_first(), pol_w = l2_policy_evict_last();
    while (iter.next(sc, e, seg)) {
      int g, mt, nt, split;
      unit_decode(sc, seg.u, g, mt, nt, split);
      const int m0 = mt * (2 * TC_BLOCK_M) + (int)rank * TC_BLOCK_M;
      const int n0 = nt * BLOCK_N + (int)rank * (BLOCK_N / 2);
      int cb = 0, cw = 0, ch = 0, cd = 0;
      if (p.mode == GEMM_CONV_S2D) {
        int Do = p.Do;
        cw = m0 % Do;
        ch = (m0 / Do) % Do;
        cd = (m0 / (Do * Do)) % Do;
        cb = m0 / (Do * Do * Do);
      }
      for (int kb = seg.kb0; kb < seg.kb1; ++kb, ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, e.err, 1, 64);
        if (leader) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
        unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
        if (p.mode == GEMM_CONV_S2D) {
          int a = kb / e.kb_per_a, c = (kb - a * e.kb_per_a) * TC_BLOCK_K;
          // conv: every s2d input voxel is read by 8 kernel offsets (of different tiles): default policy
          tma_load_5d_pair(sa, &args.tmA[0], &full_bar[s], c, cw + (a & 1), ch + ((a >> 1) & 1),
                           cd + ((a >> 2) & 1), cb);
        } else {
          tma_load_2d_pair_hint(sa, &args.tmA[0], &full_bar[s], kb * TC_BLOCK_K, m0, pol_a);
        }
        tma_load_2d_pair_hint(sa + TC_A_BYTES, &args.tmW[0], &full_bar[s], kb * TC_BLOCK_K, n0, pol_w);
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===== MMA issuer (leader CTA only): M = 256 over the pair, N = 256
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                               ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)((2 * TC_BLOCK_M) >> 4) << 24);
    int kc = 0, it = 0;
    TcPairIter iter(sc, e.kb_total);
    TcSeg seg;
    for (; iter.next(sc, e, seg); ++it) {
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], aph ^ 1, e.err, 4, 32);    // both CTAs drained this buffer
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = seg.kb0; kb < seg.kb1; ++kb, ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(&full_bar[s], ph, e.err, 2, 20);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t adesc = make_sw128_desc(sa);
        const uint64_t bdesc = make_sw128_desc(sa + TC_A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BLOCK_K / 16; ++k)
          umma_bf16_pair(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                         (kb > seg.kb0 || k > 0) ? 1u : 0u);
        umma_commit_pair(&empty_bar[s]);
      }
      umma_commit_pair(&tmem_full_bar[acc]);
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, own 128 accumulator rows): as in k_gemm_tc_persistent
    const int q = warp & 3, hsel = (warp - 4) >> 2;
    constexpr int EPI_COLS = BLOCK_N / 2, NCH = EPI_COLS / 32;
    const uint32_t stg_s = smem_u32(out_stage + (size_t)(warp - 4) * TC_STG_BYTES);
    const int wslot = warp - 4;
    constexpr long long SLOT_FLOATS = (long long)TC_BLOCK_M * BLOCK_N;
    int it = 0;
    TcPairIter iter(sc, e.kb_total);
    TcSeg seg;
    for (; iter.next(sc, e, seg); ++it) {
      int g, mt, nt, split;
      unit_decode(sc, seg.u, g, mt, nt, split);
      const int m0 = mt * (2 * TC_BLOCK_M) + (int)rank * TC_BLOCK_M;
      const int n0 = nt * BLOCK_N + hsel * EPI_COLS;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], aph, e.err, 3, 128);
      tcgen05_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      const long long roff = row_ok ? out_row_offset(p, m) : 0;
      const uint32_t tacc = tmem_base + (uint32_t)(acc * BLOCK_N + hsel * EPI_COLS) +
                            ((uint32_t)(q * 32) << 16);
      if (seg.partial) {
        float4* slot = reinterpret_cast<float4*>(sc.sk_ws + (long long)blockIdx.x * SLOT_FLOATS +
                                                 (long long)wslot * 32 * EPI_COLS);
        const uint64_t pol_keep = l2_policy_evict_last();     // read back at the end of the kernel
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tacc + (uint32_t)(c * 32), r);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_global_v4_hint(slot + (c * 8 + j) * 32 + lane,
                              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                          __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])),
                              pol_keep);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0)
          asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(sc.sk_flags + blockIdx.x * 8 + wslot),
                       "r"(1) : "memory");
      } else {
        // peers of this CTA: the same-rank CTA of the following clusters
        if (seg.n_peers > 0) {
          if (lane == 0) {
            for (int pi = 0; pi < seg.n_peers; ++pi) {
              const int* f = sc.sk_flags + (2 * (seg.peer0 + pi) + (int)rank) * 8 + wslot;
              int v = 0;
              for (long long spin = 0; spin < (1LL << 24); ++spin) {
                asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if (v) break;
                __nanosleep(64);
              }
              if (!v && e.err) atomicExch(e.err, 7);
            }
          }
          __syncwarp();
        }
        const unsigned okmask = __ballot_sync(0xffffffffu, row_ok);
        int relu_i = p.relu, Nn = p.N;
        const float* bias = p.bias;
        bf16* outp = reinterpret_cast<bf16*>(p.out);
        asm volatile("" : "+r"(relu_i), "+r"(Nn), "+l"(bias), "+l"(outp));
        const bool relu = relu_i != 0;
        long long rofs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rofs[i] = __shfl_sync(0xffffffffu, roff, i * 4 + (lane >> 3));
        uint32_t ra[32], rb[32];
        auto process = [&](uint32_t (&r)[32], int c) {
          const int n = n0 + c * 32;
          if (n < Nn) {
            for (int pi = 0; pi < seg.n_peers; ++pi) {
              const float4* ps = reinterpret_cast<const float4*>(
                  sc.sk_ws + (long long)(2 * (seg.peer0 + pi) + (int)rank) * SLOT_FLOATS +
                  (long long)wslot * 32 * EPI_COLS);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 a = __ldcg(ps + (c * 8 + j) * 32 + lane);
                r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + a.x);
                r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + a.y);
                r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + a.z);
                r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + a.w);
              }
            }
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 bq = bias ? __ldg(reinterpret_cast<const float4*>(bias + n) + j)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bq.x;
              v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bq.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bq.z;
              v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bq.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 uu;
              if (relu) {
                uu.x = pack_bf16x2_relu(v[8 * j + 0], v[8 * j + 1]);
                uu.y = pack_bf16x2_relu(v[8 * j + 2], v[8 * j + 3]);
                uu.z = pack_bf16x2_relu(v[8 * j + 4], v[8 * j + 5]);
                uu.w = pack_bf16x2_relu(v[8 * j + 6], v[8 * j + 7]);
              } else {
                uu.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                uu.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                uu.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                uu.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              }
              const int piece = ((c & 1) * 4 + j) ^ (lane & 7);
              sts_v4(stg_s + lane * 128 + piece * 16, uu);
            }
          }
          if (c & 1) {
            __syncwarp();
            const int ncol = n0 + (c - 1) * 32 + (lane & 7) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = i * 4 + (lane >> 3);
              const uint4 uu = lds_v4(stg_s + row * 128 + (((lane & 7) ^ (row & 7)) * 16));
              if (((okmask >> row) & 1u) && ncol < Nn)
                *reinterpret_cast<uint4*>(outp + rofs[i] + ncol) = uu;
            }
            __syncwarp();
          }
        };
        tmem_ld_32x32_issue(tacc, ra);
        tmem_ld_wait(ra);
#pragma unroll
        for (int c = 0; c < NCH; c += 2) {
          tmem_ld_32x32_issue(tacc + (uint32_t)((c + 1) * 32), rb);
          process(ra, c);
          tmem_ld_wait(rb);
          if (c + 2 < NCH) tmem_ld_32x32_issue(tacc + (uint32_t)((c + 2) * 32), ra);
          process(rb, c + 1);
          if (c + 2 < NCH) tmem_ld_wait(ra);
        }
        if (seg.n_peers > 0) {
          __syncwarp();
          if (lane == 0)
            for (int pi = 0; pi < seg.n_peers; ++pi)
              sc.sk_flags[(2 * (seg.peer0 + pi) + (int)rank) * 8 + wslot] = 0;
        }
      }
      // this warp is done reading the accumulator buffer: tell the leader's MMA issuer
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                       // both CTAs are done with the pair's TMEM / barriers
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(2 * BLOCK_N))
                 : "memory");
  }
}

// ------------------------------------------------------------------ the pose heads as ONE launch
// Layers 1-3 of the three heads (model.py:239-254: head1 = the three first layers side by side,
// 3 x conv2_*, 3 x conv3_*) are 7 GEMMs whose only coupling is per 128-row tile: a layer-2 tile
// needs the layer-1 tiles of its rows, a layer-3 tile the layer-2 tile of its head and rows.  As
// separate launches each layer ends in a partially filled wave plus a launch gap (36.5 + 16.1 +
// 10.0 us for 33 us of tensor work).  Here the 504 + 189 + 189 tiles are ONE unit list of a
// persistent kernel: units are dealt round-robin in list order, a unit's TMA producer waits for
// a global arrival counter of its input tile(s) before the first load, the epilogue warps bump
// the counter of their output tile after their stores.  A unit only ever waits for units earlier
// in the list and every CTA walks the list in order, so there is no cycle.
// Counters are never reset: every launch reads an epoch word at start (thresholds = (epoch + 1) x
// arrivals per launch) and the last CTA to finish bumps it, so CUDA-graph replays need no memset.
constexpr int HD_LAYERS = 7;
struct HeadsArgs {
  CUtensorMap tmA[HD_LAYERS];
  CUtensorMap tmW[HD_LAYERS];
  GemmParams p[HD_LAYERS];
  int bn[HD_LAYERS];            // N tile: 256 or 128
  int kb[HD_LAYERS];            // K blocks
  int n_tiles[HD_LAYERS];
  int unit_begin[HD_LAYERS + 1];
  int dep_base[HD_LAYERS];      // counter (per m tile) to wait on, -1 = none
  int dep_target[HD_LAYERS];    // its arrivals per launch
  int sig_base[HD_LAYERS];      // counter (per m tile) to bump, -1 = none
  int m_tiles, n_units;
  int* sync;                    // [0] epoch, [1] finished CTAs, [16 ..] counters
  int* err;
};

__device__ __forceinline__ void hd_decode(const HeadsArgs& a, int u, int& q, int& mt, int& nt) {
  q = 0;
#pragma unroll
  for (int i = 1; i < HD_LAYERS; ++i) q += (u >= a.unit_begin[i]) ? 1 : 0;
  const int r = u - a.unit_begin[q];
  nt = r % a.n_tiles[q];
  mt = r / a.n_tiles[q];
}

template <int STAGES>
__global__ void __launch_bounds__(TC_P_THREADS, 1)
k_heads_tc(const __grid_constant__ HeadsArgs args) {
  constexpr int MAX_N = 256;
  constexpr int STAGE_BYTES = TC_A_BYTES + MAX_N * TC_BLOCK_K * 2;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* out_stage = smem + (size_t)STAGES * STAGE_BYTES;   // 8 x TC_STG_BYTES
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ int epoch_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    int ep;
    asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(ep) : "l"(args.sync) : "memory");
    epoch_slot = ep;
#pragma unroll
    for (int q = 0; q < HD_LAYERS; ++q) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.tmA[q])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&args.tmW[q])) : "memory");
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
                 "r"((uint32_t)(2 * MAX_N))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const unsigned epoch1 = (unsigned)epoch_slot + 1u;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer
    int kc = 0;
    for (int u = blockIdx.x; u < args.n_units; u += gridDim.x) {
      int q, mt, nt;
      hd_decode(args, u, q, mt, nt);
      if (args.dep_base[q] >= 0) {
        // input rows of this tile are written by earlier units (other SMs, generic-proxy stores)
        const int* c = args.sync + 16 + args.dep_base[q] + mt;
        const unsigned target = epoch1 * (unsigned)args.dep_target[q];
        long long t0 = clock64();
        for (;;) {
          int v;
          asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
          if ((int)((unsigned)v - target) >= 0) break;
          __nanosleep(100);
          if (clock64() - t0 > 4000000000LL) {
            if (args.err) atomicExch(args.err, 9);
            __threadfence_system();
            __trap();
          }
        }
        // order the acquired (generic-proxy) view before this thread's async-proxy (TMA) reads
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      const int m0 = mt * TC_BLOCK_M, n0 = nt * args.bn[q];
      const uint32_t bytes = (uint32_t)(TC_A_BYTES + args.bn[q] * TC_BLOCK_K * 2);
      const int kbn = args.kb[q];
      for (int kb = 0; kb < kbn; ++kb, ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, args.err, 1, 64);
        mbar_expect_tx(&full_bar[s], bytes);
        unsigned char* sa = smem + (size_t)s * STAGE_BYTES;
        tma_load_2d(sa, &args.tmA[q], &full_bar[s], kb * TC_BLOCK_K, m0);
        tma_load_2d(sa + TC_A_BYTES, &args.tmW[q], &full_bar[s], kb * TC_BLOCK_K, n0);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer
    int kc = 0, it = 0;
    for (int u = blockIdx.x; u < args.n_units; u += gridDim.x, ++it) {
      int q, mt, nt;
      hd_decode(args, u, q, mt, nt);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                             ((uint32_t)(args.bn[q] >> 3) << 17) | ((uint32_t)(TC_BLOCK_M >> 4) << 24);
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], aph ^ 1, args.err, 4, 32);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * MAX_N);
      const int kbn = args.kb[q];
      for (int kb = 0; kb < kbn; ++kb, ++kc) {
        const int s = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1;
        mbar_wait(&full_bar[s], ph, args.err, 2, 20);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
        const uint64_t adesc = make_sw128_desc(sa);
        const uint64_t bdesc = make_sw128_desc(sa + TC_A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BLOCK_K / 16; ++k)
          umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&tmem_full_bar[acc]);
    }
  } else if (warp >= 4) {
    // ===== epilogue: 8 warps; warp w reads TMEM lane quarter w%4 and column half (w-4)/4
    const int q4 = warp & 3, hsel = (warp - 4) >> 2;
    const uint32_t stg_s = smem_u32(out_stage + (size_t)(warp - 4) * TC_STG_BYTES);
    int it = 0;
    for (int u = blockIdx.x; u < args.n_units; u += gridDim.x, ++it) {
      int q, mt, nt;
      hd_decode(args, u, q, mt, nt);
      const GemmParams& p = args.p[q];
      const int epi_cols = args.bn[q] >> 1, nch = epi_cols >> 5;
      const int m0 = mt * TC_BLOCK_M, n0 = nt * args.bn[q] + hsel * epi_cols;
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], aph, args.err, 3, 128);
      tcgen05_fence_after();
      const int m = m0 + q4 * 32 + lane;
      const bool row_ok = m < p.M;
      const long long roff = row_ok ? (long long)m * p.ldo + p.col_off : 0;
      const uint32_t tacc = tmem_base + (uint32_t)(acc * MAX_N + hsel * epi_cols) +
                            ((uint32_t)(q4 * 32) << 16);
      const unsigned okmask = __ballot_sync(0xffffffffu, row_ok);
      int relu_i = p.relu, Nn = p.N;
      const float* bias = p.bias;
      bf16* outp = reinterpret_cast<bf16*>(p.out);
      asm volatile("" : "+r"(relu_i), "+r"(Nn), "+l"(bias), "+l"(outp));
      const bool relu = relu_i != 0;
      long long rofs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rofs[i] = __shfl_sync(0xffffffffu, roff, i * 4 + (lane >> 3));
      uint32_t ra[32], rb[32];
      auto process = [&](uint32_t (&r)[32], int c) {
        const int n = n0 + c * 32;
        if (n < Nn) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 bq = bias ? __ldg(reinterpret_cast<const float4*>(bias + n) + j)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bq.x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bq.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bq.z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bq.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 uu;
            if (relu) {
              uu.x = pack_bf16x2_relu(v[8 * j + 0], v[8 * j + 1]);
              uu.y = pack_bf16x2_relu(v[8 * j + 2], v[8 * j + 3]);
              uu.z = pack_bf16x2_relu(v[8 * j + 4], v[8 * j + 5]);
              uu.w = pack_bf16x2_relu(v[8 * j + 6], v[8 * j + 7]);
            } else {
              uu.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              uu.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              uu.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              uu.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            }
            const int piece = ((c & 1) * 4 + j) ^ (lane & 7);
            sts_v4(stg_s + lane * 128 + piece * 16, uu);
          }
        }
        if (c & 1) {
          __syncwarp();
          const int ncol = n0 + (c - 1) * 32 + (lane & 7) * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + (lane >> 3);
            const uint4 uu = lds_v4(stg_s + row * 128 + (((lane & 7) ^ (row & 7)) * 16));
            if (((okmask >> row) & 1u) && ncol < Nn)
              *reinterpret_cast<uint4*>(outp + rofs[i] + ncol) = uu;
          }
          __syncwarp();
        }
      };
      tmem_ld_32x32_issue(tacc, ra);
      tmem_ld_wait(ra);
#pragma unroll 1
      for (int c = 0; c < nch; c += 2) {
        tmem_ld_32x32_issue(tacc + (uint32_t)((c + 1) * 32), rb);
        process(ra, c);
        tmem_ld_wait(rb);
        if (c + 2 < nch) tmem_ld_32x32_issue(tacc + (uint32_t)((c + 2) * 32), ra);
        process(rb, c + 1);
        if (c + 2 < nch) tmem_ld_wait(ra);
      }
      // accumulator drained: hand the buffer back to the MMA issuer
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc]))
                     : "memory");
      // this warp's part of the output tile is stored: publish it to the consumers of the tile
      if (args.sig_base[q] >= 0) {
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(args.sync + 16 + args.sig_base[q] + mt, 1);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(2 * MAX_N))
                 : "memory");
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const int d = atomicAdd(args.sync + 1, 1);
    if (d == (int)gridDim.x - 1) {          // last CTA: next launch gets the next epoch
      args.sync[1] = 0;
      __threadfence();
      atomicAdd(args.sync, 1);
    }
  }
}

// split-K epilogue: sum the slices, bias + activation + layout; 4 columns per thread
__global__ void k_splitk_finish(const float* __restrict__ ws, GemmParams p, int splitk) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long MN = (long long)p.M * p.N;
  if (i * 4 >= MN) return;
  int m = (int)((i * 4) / p.N), n = (int)((i * 4) % p.N);   // N % 32 == 0
  float4 v = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0, 0, 0, 0);
  for (int s = 0; s < splitk; ++s) {
    float4 a = __ldcs(reinterpret_cast<const float4*>(ws + s * MN) + i);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  long long off = out_row_offset(p, m) + n;
  if (p.out_mode == OUT_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = v;
  } else {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + off) = u;
  }
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
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

static int encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims,
                  const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  PFN_encodeTiled fn = get_encode();
  if (!fn) return MF_E_UNSUPPORTED;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MF_OK : MF_E_BADARG;
}

template <int BLOCK_N, int STAGES>
static int launch_persistent(const TcArgs& args, const TcSched& sc, cudaStream_t stream) {
  constexpr int smem = STAGES * (TC_A_BYTES + BLOCK_N * TC_BLOCK_K * 2) + 1024 + 8 * TC_STG_BYTES;
  MF_ENSURE_DYN_SMEM((k_gemm_tc_persistent<BLOCK_N, STAGES>), smem);
  int n_sm = 148, dev = 0;
  MF_CUDA_TRY(cudaGetDevice(&dev));
  MF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  int grid = (sc.n_units < n_sm && !sc.streamk) ? sc.n_units : n_sm;
  k_gemm_tc_persistent<BLOCK_N, STAGES><<<grid, TC_P_THREADS, smem, stream>>>(args, sc);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

template <int STAGES>
static int launch_pair(const TcArgs& args, const TcSched& sc, int n_sm, cudaStream_t stream) {
  constexpr int smem = STAGES * (TC_A_BYTES + 128 * TC_BLOCK_K * 2) + 1024 + 8 * TC_STG_BYTES;
  MF_ENSURE_DYN_SMEM((k_gemm_tc_pair<STAGES>), smem);
  int grid = n_sm & ~1;
  if (!sc.streamk && 2 * sc.n_units < grid) grid = 2 * sc.n_units;
  k_gemm_tc_pair<STAGES><<<grid, TC_P_THREADS, smem, stream>>>(args, sc);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

// how many CTA pairs of k_gemm_tc_pair the device can hold at once (cached per device)
static int pair_clusters(int dev) {
  static int cached[64];
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
  constexpr int smem = 6 * (TC_A_BYTES + 128 * TC_BLOCK_K * 2) + 1024 + 8 * TC_STG_BYTES;
  if (cudaFuncSetAttribute(k_gemm_tc_pair<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
      cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2, 1, 1);
  cfg.blockDim = dim3(TC_P_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_gemm_tc_pair<6>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  if (dev >= 0 && dev < 64) cached[dev] = n > 0 ? n : -1;
  return n;
}

template <int BLOCK_N, int STAGES>
static int launch(const TcArgs& args, dim3 grid, cudaStream_t stream) {
  constexpr int smem = STAGES * (TC_A_BYTES + BLOCK_N * TC_BLOCK_K * 2) + 1024;
  MF_ENSURE_DYN_SMEM((k_gemm_tc<BLOCK_N, STAGES>), smem);
  k_gemm_tc<BLOCK_N, STAGES><<<grid, 256, smem, stream>>>(args);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

static int check_shape(const GemmParams& p) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || !p.A || !p.W || !p.out) return MF_E_BADARG;
  // shapes this kernel family covers; everything else goes to the SIMT kernel
  if (p.K < 64 || p.K % 8 != 0 || p.ldw % 8 != 0) return MF_E_UNSUPPORTED;
  if (p.N % 32 != 0 || p.N < 128) {
    // ragged / small N: only fp32 row-major outputs, a single N tile, no split-K
    if (p.out_mode != OUT_F32 || p.N > 128 || p.N < 8) return MF_E_UNSUPPORTED;
  }
  if (((uintptr_t)p.A & 15) || ((uintptr_t)p.W & 15) || ((uintptr_t)p.out & 15))
    return MF_E_UNSUPPORTED;
  if (p.N % 32 == 0 && p.out_mode != OUT_S2D_BF16 && ((p.ldo % 8) || (p.col_off % 8)))
    return MF_E_UNSUPPORTED;
  return MF_OK;
}

static int make_maps(const GemmParams& p, int BN, CUtensorMap* tmA, CUtensorMap* tmW, int* kb_per_a) {
  int rc;
  *kb_per_a = 0;
  if (p.mode == GEMM_CONV_S2D) {
    const int Do = p.Do, Js = Do + 1;
    if (!(Do == 8 || Do == 16) || p.Ci8 % 64 != 0 || p.K != 8 * p.Ci8) return MF_E_UNSUPPORTED;
    if (p.M % (Do * Do * Do) != 0) return MF_E_UNSUPPORTED;
    const int Bn = p.M / (Do * Do * Do);
    *kb_per_a = p.Ci8 / 64;
    cuuint64_t dims[5] = {(cuuint64_t)p.Ci8, (cuuint64_t)Js, (cuuint64_t)Js, (cuuint64_t)Js,
                          (cuuint64_t)Bn};
    cuuint64_t str[4] = {(cuuint64_t)p.Ci8 * 2, (cuuint64_t)Js * p.Ci8 * 2,
                         (cuuint64_t)Js * Js * p.Ci8 * 2, (cuuint64_t)Js * Js * Js * p.Ci8 * 2};
    // 128 consecutive output voxels (w fastest): 16x8x1 for Do=16, 8x8x2 for Do=8
    cuuint32_t box[5] = {64, (cuuint32_t)Do, 8, (cuuint32_t)(Do == 16 ? 1 : 2), 1};
    rc = encode(tmA, p.A, 5, dims, str, box);
  } else {
    if (p.lda % 8 != 0) return MF_E_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
    cuuint64_t str[1] = {(cuuint64_t)p.lda * 2};
    cuuint32_t box[2] = {64, 128};
    rc = encode(tmA, p.A, 2, dims, str, box);
  }
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.N};
  cuuint64_t str[1] = {(cuuint64_t)p.ldw * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BN};
  return encode(tmW, p.W, 2, dims, str, box);
}

}  // namespace mf

using namespace mf;

extern "C" size_t mf_gemm_bf16_tc_workspace_bytes(int M, int N) {
  size_t splitk = (size_t)M * N * 4 * 4 + 256;              // up to 4 split-K slices
  size_t streamk = (size_t)256 * TC_BLOCK_M * 256 * 4;      // one 128 x 256 fp32 slot per SM (<= 256)
  return splitk > streamk ? splitk : streamk;
}

extern "C" int mf_gemm_bf16_tc_ex(const GemmParams* hp, int n_groups, void* workspace,
                                  size_t workspace_bytes, int32_t* sync, long long* stamps,
                                  int one_shot, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!hp || n_groups < 1 || n_groups > TC_MAX_GROUPS) return MF_E_BADARG;
  TcArgs args;
  memset(&args, 0, sizeof(args));
  const GemmParams& p0 = hp[0];
  for (int g = 0; g < n_groups; ++g) {
    int rc = check_shape(hp[g]);
    if (rc) return rc;
    const bool small_n = p0.N <= 128 && hp[g].N <= 128 && hp[g].out_mode == OUT_F32 &&
                         p0.out_mode == OUT_F32;
    if (hp[g].M != p0.M || (hp[g].N != p0.N && !small_n) || hp[g].K != p0.K ||
        hp[g].mode != p0.mode || hp[g].Do != p0.Do || hp[g].Ci8 != p0.Ci8)
      return MF_E_UNSUPPORTED;
    args.p[g] = hp[g];
  }
  const int BN = (p0.N >= 256) ? 256 : 128;
  args.e.kb_total = (p0.K + TC_BLOCK_K - 1) / TC_BLOCK_K;
  for (int g = 0; g < n_groups; ++g) {
    int rc = make_maps(hp[g], BN, &args.tmA[g], &args.tmW[g], &args.e.kb_per_a);
    if (rc) return rc;
  }
  int maxN = p0.N;
  for (int g = 1; g < n_groups; ++g) maxN = hp[g].N > maxN ? hp[g].N : maxN;
  const int m_tiles = (p0.M + TC_BLOCK_M - 1) / TC_BLOCK_M, n_tiles = (maxN + BN - 1) / BN;
  int splitk = 1;
  const int tiles = m_tiles * n_tiles * n_groups;
  int n_sm = 148, dev = 0;
  MF_CUDA_TRY(cudaGetDevice(&dev));
  MF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  // ---- CTA-pair kernel (cta_group::2): one group, 256-wide N tiles, bf16 output
  if (one_shot == 0 && n_groups == 1 && BN == 256 && p0.N % 32 == 0 && p0.out_mode != OUT_F32) {
    const int ncl = pair_clusters(dev);
    if (ncl > 0 && 2 * ncl >= (n_sm & ~1)) {
      int rc = make_maps(p0, 128, &args.tmA[0], &args.tmW[0], &args.e.kb_per_a);   // B box: 128 rows
      if (rc) return rc;
      const int m_tiles2 = (p0.M + 2 * TC_BLOCK_M - 1) / (2 * TC_BLOCK_M);
      const int tiles2 = m_tiles2 * n_tiles, clusters = (n_sm & ~1) / 2;
      bool sk = false;
      if (sync && args.e.kb_total >= 32 && n_sm <= MF_GEMM_TC_SYNC_INTS / 8) {
        const int waves = (tiles2 + clusters - 1) / clusters;
        const double busy = (double)tiles2 / ((double)waves * clusters);
        const size_t need = (size_t)n_sm * TC_BLOCK_M * 256 * 4;
        if (busy < 0.95 && workspace && workspace_bytes >= need) sk = true;
      }
      args.e.kb_per_split = args.e.kb_total;
      args.e.splitk = 1;
      args.e.err = nullptr;
      TcSched sc{m_tiles2, n_tiles, 1, 1, tiles2, stamps, sk ? 1 : 0, sync, (float*)workspace};
      return launch_pair<6>(args, sc, n_sm, stream);
    }
  }
  // stream-K when the caller provides the flag words and whole-tile scheduling would idle > 5 %
  // of the SMs (or there are fewer tiles than SMs) on a long-K problem with bf16 output
  bool streamk = false;
  // (short-K problems are epilogue-bound: the extra accumulator round trip costs more than the
  // idle tail it removes - measured on the head GEMMs: 36.5 -> 40.4, 16.1 -> 21.1, 10.0 -> 13.5 us)
  bool sk_ok = sync && one_shot != 1 && args.e.kb_total >= 32 && n_sm <= MF_GEMM_TC_SYNC_INTS / 8;
  for (int g = 0; g < n_groups; ++g)
    sk_ok = sk_ok && hp[g].N % 32 == 0 && hp[g].out_mode != OUT_F32;
  if (sk_ok) {
    const int waves = (tiles + n_sm - 1) / n_sm;
    const double busy = (double)tiles / ((double)waves * n_sm);
    const size_t need = (size_t)n_sm * TC_BLOCK_M * BN * 4;
    if (busy < 0.95 && workspace && workspace_bytes >= need) streamk = true;
  }
  if (!streamk && n_groups == 1 && tiles <= 74 && args.e.kb_total >= 64 && p0.N % 32 == 0) {
    splitk = 148 / tiles;
    if (splitk > 4) splitk = 4;
    if (splitk < 1) splitk = 1;
  }
  args.e.kb_per_split = (args.e.kb_total + splitk - 1) / splitk;
  splitk = (args.e.kb_total + args.e.kb_per_split - 1) / args.e.kb_per_split;
  args.e.splitk = splitk;
  args.e.err = nullptr;
  if (splitk > 1) {
    if (!workspace || workspace_bytes < (size_t)splitk * p0.M * p0.N * 4) return MF_E_WORKSPACE;
    args.e.ws = (float*)workspace;
  }
  dim3 grid(n_tiles, m_tiles, splitk * n_groups);
  int rc;
  if (one_shot != 1) {
    TcSched sc{m_tiles, n_tiles, n_groups, splitk, m_tiles * n_tiles * n_groups * splitk,
               stamps, streamk ? 1 : 0, sync, (float*)workspace};
    if (BN == 256) rc = launch_persistent<256, 4>(args, sc, stream);
    else rc = launch_persistent<128, 6>(args, sc, stream);
  } else if (BN == 256) rc = launch<256, 4>(args, grid, stream);
  else rc = launch<128, 6>(args, grid, stream);
  if (rc) return rc;
  if (splitk > 1) {
    MF_PREFER_MAX_SMEM(k_splitk_finish);
    k_splitk_finish<<<div_up((long long)p0.M * p0.N / 4, 256), 256, 0, stream>>>(args.e.ws, p0, splitk);
    MF_LAUNCH_CHECK();
  }
  return MF_OK;
}

extern "C" int mf_gemm_bf16_tc_grouped(const GemmParams* hp, int n_groups, void* workspace,
                                       size_t workspace_bytes, void* stream_) {
  return mf_gemm_bf16_tc_ex(hp, n_groups, workspace, workspace_bytes, nullptr, nullptr, 0, stream_);
}

extern "C" int mf_cnn_heads_tc(const GemmParams* layers, int n_layers, int32_t* sync, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!layers || n_layers != HD_LAYERS || !sync) return MF_E_BADARG;
  HeadsArgs a;
  memset(&a, 0, sizeof(a));
  const int M = layers[0].M;
  const int m_tiles = (M + TC_BLOCK_M - 1) / TC_BLOCK_M;
  if (16 + 4 * m_tiles > MF_HEADS_SYNC_INTS) return MF_E_TOOLARGE;
  int units = 0;
  for (int q = 0; q < HD_LAYERS; ++q) {
    const GemmParams& p = layers[q];
    int rc = check_shape(p);
    if (rc) return rc;
    if (p.M != M || p.mode != GEMM_LINEAR || p.out_mode != OUT_BF16 || p.N % 32 != 0 || p.N < 128)
      return MF_E_UNSUPPORTED;
    if (q >= 2 && q <= 3 && (p.N != layers[1].N || p.K != layers[1].K)) return MF_E_UNSUPPORTED;
    if (q >= 5 && (p.N != layers[4].N || p.K != layers[4].K)) return MF_E_UNSUPPORTED;
    a.p[q] = p;
    a.bn[q] = p.N >= 256 ? 256 : 128;
    a.kb[q] = (p.K + TC_BLOCK_K - 1) / TC_BLOCK_K;
    a.n_tiles[q] = (p.N + a.bn[q] - 1) / a.bn[q];
    a.unit_begin[q] = units;
    units += m_tiles * a.n_tiles[q];
    int dummy = 0;
    rc = make_maps(p, a.bn[q], &a.tmA[q], &a.tmW[q], &dummy);
    if (rc) return rc;
  }
  a.unit_begin[HD_LAYERS] = units;
  // counters: [0, m_tiles) layer-1 rows; [m_tiles * (1 + h), ...) layer-2 rows of head h
  a.dep_base[0] = -1;
  a.sig_base[0] = 0;
  for (int h = 0; h < 3; ++h) {
    a.dep_base[1 + h] = 0;
    a.dep_target[1 + h] = a.n_tiles[0] * 8;            // 8 epilogue warps per tile
    a.sig_base[1 + h] = m_tiles * (1 + h);
    a.dep_base[4 + h] = m_tiles * (1 + h);
    a.dep_target[4 + h] = a.n_tiles[1 + h] * 8;
    a.sig_base[4 + h] = -1;
  }
  a.m_tiles = m_tiles;
  a.n_units = units;
  a.sync = sync;
  a.err = nullptr;
  constexpr int STAGES = 4;
  constexpr int smem = STAGES * (TC_A_BYTES + 256 * TC_BLOCK_K * 2) + 1024 + 8 * TC_STG_BYTES;
  MF_ENSURE_DYN_SMEM((k_heads_tc<STAGES>), smem);
  int n_sm = 148, dev = 0;
  MF_CUDA_TRY(cudaGetDevice(&dev));
  MF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  // every CTA must be resident (units wait on units of other CTAs): one CTA per SM at most
  const int grid = units < n_sm ? units : n_sm;
  k_heads_tc<STAGES><<<grid, TC_P_THREADS, smem, stream>>>(a);
  MF_LAUNCH_CHECK();
  return MF_OK;
}

extern "C" int mf_gemm_bf16_tc(const GemmParams* hp, void* workspace, size_t workspace_bytes,
                               void* stream_) {
  return mf_gemm_bf16_tc_grouped(hp, 1, workspace, workspace_bytes, stream_);
}
