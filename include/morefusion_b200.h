/*
 * morefusion_b200.h -- C ABI of libmorefusion_sm100a.so
 *
 * Drop-in boundary for the volumetric-pose hot path of wkentaro/morefusion
 * (SURVEY.md section 8).  The reference has no native library: each entry point
 * below replaces one chainer.Function / Link method (cited per function, paths
 * relative to the reference checkout).  A binding only needs ctypes/cffi: plain
 * pointers, sizes and scalars -- no torch, chainer or cupy types.
 *
 * Conventions
 *   - every pointer named in/out is a DEVICE pointer unless marked "host";
 *     tensors are dense, C-contiguous, float32 / int32 as in the reference.
 *   - grids are [X,Y,Z] (z fastest), batched [B,C,X,Y,Z]; flat voxel index
 *     ix*Y*Z + iy*Z + iz (average_voxelization_3d.py:92-96).
 *   - quaternions are (w,x,y,z); transforms are row-major 4x4.
 *   - the caller owns all memory including `workspace`; kernels never allocate,
 *     never retain pointers past the call, never synchronise the host, and are
 *     safe under CUDA-graph capture.  Work is enqueued on `stream`
 *     (a cudaStream_t passed as void*).
 *   - return value: 0 = ok; <0 = invalid argument (MF_E_*); >0 = cudaError_t.
 *   - `flags` (device int32, may be NULL) receives OR-ed MF_FLAG_* bits so the
 *     host binding can raise the reference's ValueError lazily.
 */
#ifndef MOREFUSION_B200_H_
#define MOREFUSION_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MF_OK 0
#define MF_E_BADARG (-1)      /* null pointer / negative size */
#define MF_E_TOOLARGE (-2)    /* B*X*Y*Z (or P*K) does not fit in int32 keys */
#define MF_E_WORKSPACE (-3)   /* workspace smaller than *_workspace_bytes() */
#define MF_E_UNSUPPORTED (-4) /* shape outside what the kernel family covers */

#define MF_FLAG_NAN_POINTS 1      /* "points include nan" (average_voxelization_3d.py:13-14,47-48) */
#define MF_FLAG_UNSORTED_BATCH 2  /* batch_indices not non-decreasing (informational) */
#define MF_FLAG_BAD_BATCH_INDEX 4 /* batch index outside [0,B): point dropped */

int mf_abi_version(void);
/* number of SMs / device name the library sees (diagnostics for the binding) */
int mf_device_sm_count(int device);

/* ------------------------------------------------------------------------
 * a1  average_voxelization_3d
 *     replaces AverageVoxelization3D.forward_gpu / backward_gpu
 *     (morefusion/functions/geometry/average_voxelization_3d.py:43-118, :147-220)
 * values [N,C] points [N,3] batch_indices [N] -> matrix [B,C,X,Y,Z], counts [B,X,Y,Z]
 * Two launches (keys; gather + fill): every output voxel is written exactly once (no
 * memset, no atomics on HBM, nothing in `workspace` needs initialising); per-voxel sums
 * are taken in ascending point order (deterministic).
 * ------------------------------------------------------------------------ */
size_t mf_average_voxelization_3d_workspace_bytes(int64_t n_points);
/* byte offset inside `workspace` of an int32 copy of the MF_FLAG_* word of the last call (so a
 * binding needs no separate flags allocation; `flags` below may then be NULL) */
size_t mf_average_voxelization_3d_flags_offset(void);
int mf_average_voxelization_3d_fwd(
    const float* values, const float* points, const int32_t* batch_indices,
    int64_t n_points, int channels, int batch_size,
    float origin_x, float origin_y, float origin_z, float pitch,
    int X, int Y, int Z,
    float* matrix, int32_t* counts,
    void* workspace, size_t workspace_bytes, int32_t* flags, void* stream);
/* diagnostics: writes B*planes*V zero floats with one of five store patterns (see
 * csrc/voxelization.cu, scripts/fill_patterns.py) */
int mf_debug_fill_probe(float* out, int batch_size, int planes, int64_t V, int mode, void* stream);
int mf_average_voxelization_3d_bwd(
    const float* gmatrix, const int32_t* counts,
    const float* points, const int32_t* batch_indices,
    int64_t n_points, int channels, int batch_size,
    float origin_x, float origin_y, float origin_z, float pitch,
    int X, int Y, int Z,
    float* gvalues, void* stream);

/* ------------------------------------------------------------------------
 * a6  max_voxelization_3d
 *     replaces MaxVoxelization3D.forward_gpu / backward_gpu
 *     (morefusion/functions/geometry/max_voxelization_3d.py:57-143, :145-185)
 * winner per voxel = max intensity, lowest point id on exact ties.
 * ------------------------------------------------------------------------ */
size_t mf_max_voxelization_3d_workspace_bytes(int batch_size, int X, int Y, int Z);
int mf_max_voxelization_3d_fwd(
    const float* values, const float* points, const int32_t* batch_indices,
    const float* intensities,
    int64_t n_points, int channels, int batch_size,
    float origin_x, float origin_y, float origin_z, float pitch,
    int X, int Y, int Z,
    float* matrix, int32_t* indices,
    void* workspace, size_t workspace_bytes, int32_t* flags, void* stream);
int mf_max_voxelization_3d_bwd(
    const float* gmatrix, const int32_t* indices,
    int64_t n_points, int channels, int batch_size, int X, int Y, int Z,
    float* gvalues, void* stream);

/* ------------------------------------------------------------------------
 * a5  interpolate_voxel_grid
 *     replaces InterpolateVoxelGrid.forward_gpu / backward_gpu
 *     (morefusion/functions/geometry/interpolate_voxel_grid.py:159-214, :216-268)
 * voxelized [B,C,X,Y,Z], points [P,3] (voxel units), batch_indices [P] -> values [P,C]
 * `channels_last` != 0 reads/writes the grid as [B,X,Y,Z,C] (internal layout of
 * the 3D-CNN path; same arithmetic).
 * ------------------------------------------------------------------------ */
int mf_interpolate_voxel_grid_fwd(
    const float* voxelized, const float* points, const int32_t* batch_indices,
    int64_t n_points, int batch_size, int channels, int X, int Y, int Z,
    int channels_last, float* values, void* stream);
int mf_interpolate_voxel_grid_bwd(
    const float* gvalues, const float* points, const int32_t* batch_indices,
    int64_t n_points, int batch_size, int channels, int X, int Y, int Z,
    int channels_last, float* gvoxelized /* fully written */, void* stream);

/* ------------------------------------------------------------------------
 * a3  truncated_distance_function
 *     replaces TruncatedDistanceFunction.forward_gpu / backward_gpu
 *     (morefusion/functions/geometry/truncated_distance_function.py:21-101, :103-166)
 * points [P,3] -> tdf [X,Y,Z] (init = truncation), indices [X,Y,Z] (winning
 * point id, -1 if none; already divided by K as at :177).
 * Winner = lexicographic min of (distance, point id): deterministic.
 * ------------------------------------------------------------------------ */
size_t mf_truncated_distance_function_workspace_bytes(int X, int Y, int Z);
int mf_truncated_distance_function_fwd(
    const float* points, int64_t n_points,
    float pitch, float origin_x, float origin_y, float origin_z,
    int X, int Y, int Z, float truncation,
    float* tdf, int32_t* indices,
    void* workspace, size_t workspace_bytes, void* stream);
int mf_truncated_distance_function_bwd(
    const float* gtdf, const float* points, const int32_t* indices, int64_t n_points,
    float pitch, float origin_x, float origin_y, float origin_z,
    int X, int Y, int Z, float truncation,
    float* gpoints, void* stream);

/* ------------------------------------------------------------------------
 * a4  pseudo_occupancy_voxelization
 *     replaces the function at truncated_distance_function.py:181-213
 * Outputs grid/surface/inside [X,Y,Z]; also tdf, indices, w_surface, w_inside
 * (kept for backward: weights carry no gradient, :196-213).
 * ------------------------------------------------------------------------ */
size_t mf_pseudo_occupancy_voxelization_workspace_bytes(int X, int Y, int Z);
int mf_pseudo_occupancy_voxelization_fwd(
    const float* points, const float* sdf, int64_t n_points,
    float pitch, float origin_x, float origin_y, float origin_z,
    int X, int Y, int Z, float threshold, float sdf_offset,
    float* grid, float* grid_surface, float* grid_inside,
    float* tdf, int32_t* indices, float* w_surface, float* w_inside,
    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------
 * a2  occupancy_grid_3d
 *     replaces OccupancyGrid3D + the chainer op chain at
 *     morefusion/functions/geometry/occupancy_grid_3d.py:31-85
 * matrix-free: never materialises the [X,Y,Z,P] distance tensors.
 * ------------------------------------------------------------------------ */
int mf_occupancy_grid_3d_fwd(
    const float* points, int64_t n_points,
    float pitch, float origin_x, float origin_y, float origin_z,
    int X, int Y, int Z, float threshold,
    float* grid, float* dmin /* [X,Y,Z], kept for backward */, void* stream);
int mf_occupancy_grid_3d_bwd(
    const float* ggrid, const float* dmin, const float* points, int64_t n_points,
    float pitch, float origin_x, float origin_y, float origin_z,
    int X, int Y, int Z, float threshold,
    float* gpoints /* zero-filled by the call */, void* stream);

/* ------------------------------------------------------------------------
 * a7  rigid-transform operators
 *     quaternion_matrix.py:6-78, compose_transform.py:5-48,
 *     translation_matrix.py:5-39, transformation_matrix.py:5-18,
 *     transform_points.py:6-30
 * ------------------------------------------------------------------------ */
int mf_quaternion_matrix_fwd(const float* q /*[N,4]*/, int64_t n, float* R /*[N,4,4]*/, void* stream);
int mf_quaternion_matrix_bwd(const float* gR /*[N,4,4]*/, const float* q, int64_t n,
                             float* gq /*[N,4]*/, void* stream);
int mf_compose_transform_fwd(const float* R /*[N,3,3]*/, const float* t /*[N,3] or NULL*/,
                             int64_t n, float* T /*[N,4,4]*/, void* stream);
int mf_transform_points_fwd(const float* points /*[P,3]*/, int64_t n_points,
                            const float* T /*[M,4,4]*/, int64_t m,
                            float* out /*[M,P,3]*/, void* stream);
int mf_transform_points_bwd(const float* gout /*[M,P,3]*/, const float* points, int64_t n_points,
                            const float* T, int64_t m,
                            float* gpoints /*[P,3] or NULL*/, float* gT /*[M,4,4] or NULL*/,
                            void* stream);


/* ------------------------------------------------------------------------
 * a10 / a11  3D-CNN section of the singleview_3d pose model
 *     replaces the chainer/cuDNN layers of
 *     morefusion/contrib/singleview_3d/models/model.py:62-91 (layer defs),
 *     :93-141 (_extract), :239-273 (heads + pose assembly)
 * Activations are channels-last bf16, accumulation fp32.  The k4/s2/p1 Conv3Ds run as
 * implicit GEMMs over a space-to-depth ("s2d") layout of the zero-padded input:
 *   X[b][jd][jh][jw][r*C + c] = xpad[b][c][2jd+rd][2jh+rh][2jw+rw],  r = rd<<2|rh<<1|rw,
 * and weights Wg[co][((ad*2+ah)*2+aw)*8C + r*C + ci] = W[co][ci][2ad+rd][2ah+rh][2aw+rw].
 * ------------------------------------------------------------------------ */
#define GEMM_LINEAR 0    /* A[m][k] row-major, leading dimension lda */
#define GEMM_CONV_S2D 1  /* A gathered from an s2d grid: m=(b,od,oh,ow), k=(a, r*C+ci) */
#define OUT_BF16 0       /* out[m*ldo + col_off + n], bf16 */
#define OUT_F32 1        /* out[m*ldo + col_off + n], fp32 */
#define OUT_S2D_BF16 2   /* row m=(b,od,oh,ow) scattered into the s2d layout of the next conv */

typedef struct GemmParams {
  const void* A;      /* bf16 */
  const void* W;      /* bf16 [N][ldw], K-major */
  const float* bias;  /* [N] or NULL */
  void* out;
  int M, N, K;
  int mode;           /* GEMM_* */
  long long lda;      /* GEMM_LINEAR */
  long long ldw;
  int Do;             /* GEMM_CONV_S2D / OUT_S2D_BF16: output grid edge (16 for conv3, 8 for conv4) */
  int Ci8;            /* GEMM_CONV_S2D: 8 * input channels */
  int relu;
  int out_mode;       /* OUT_* */
  long long ldo;
  int col_off;
} GemmParams;

/* weights are passed TRANSPOSED ([in][out], k-major): w1_rgb [32,64], w1_pcd [3,8], ... */
int mf_cnn_point_mlp(const float* values /*[B,32,P]*/, const float* points /*[B,3,P] voxel frame*/,
                     const float* w1_rgb, const float* b1_rgb, const float* w1_pcd,
                     const float* b1_pcd, const float* w2_rgb, const float* b2_rgb,
                     const float* w2_pcd, const float* b2_pcd, int B, int P, float center,
                     void* feat_bf16 /*[B*P, ldf], cols 0..215*/, int ldf,
                     float* feat2 /*[B*P,144]*/, void* stream);
/* conv2's output goes to h2 (fp32 [B,V,16]) or, when X != NULL, as bf16 into channels
 * [c_off, c_off+16) of the s2d conv3 input X (row length 8*Ct). */
int mf_cnn_occ_convs(const float* grid_nontarget_empty /*[B,D,D,D]*/, const float* w1,
                     const float* b1, const float* w2, const float* b2, int B, int D,
                     float* h1 /*[B,V,8] scratch*/, float* h2 /*[B,V,16] or NULL*/,
                     void* X /*bf16 s2d or NULL*/, int Ct, int c_off, void* stream);
/* tensor-core variant (D == 32): conv1_occ -> bf16 [B,V,8] scratch, conv2_occ as warp-level
 * mma.sync m16n8k16 (bf16 operands, fp32 accumulate) writing bf16 into X. */
int mf_cnn_occ_convs_tc(const float* grid_nontarget_empty, const float* w1, const float* b1,
                        const float* w2, const float* b2, int B, int D, void* h1_bf16 /*[B,V,8]*/,
                        void* X, int Ct, int c_off, void* stream);
/* same, reading the occupancy grid as bytes (0/1; uint8 or bool storage): saves the cast the
 * reference does inside the model (model.py:115-116) */
int mf_cnn_occ_convs_tc_u8(const uint8_t* grid_nontarget_empty, const float* w1, const float* b1,
                           const float* w2, const float* b2, int B, int D,
                           void* h1_bf16 /*[B,V,8]*/, void* X, int Ct, int c_off, void* stream);
/* conv1_occ + conv2_occ (model.py:69-70,114-125) in ONE kernel (D == 32): each x-slab CTA evaluates
 * conv1_occ for the three slabs its dilated conv2_occ reads, in shared memory, then conv2_occ on
 * mma.sync; no [B,V,8] intermediate in global memory.  Same bits as mf_cnn_occ_convs_tc. */
int mf_cnn_occ_fused(const float* grid_nontarget_empty, const float* w1, const float* b1,
                     const float* w2, const float* b2, int B, int D, void* X, int Ct, int c_off,
                     void* stream);
int mf_cnn_occ_fused_u8(const uint8_t* grid_nontarget_empty, const float* w1, const float* b1,
                        const float* w2, const float* b2, int B, int D, void* X, int Ct, int c_off,
                        void* stream);
/* average_voxelization_3d of model.py:143-164 (origin 0, pitch 1, D^3) fused with the s2d/bf16
 * packing: writes channels [0,C) of X.  prev_keys [2*B*P] int32 (in/out, initialise to -1): the
 * first B*P entries hold the voxel key of every point of the previous call (b*D^3 + voxel, -1 =
 * dropped, bit 30 = shares its voxel) so that only those voxels are re-zeroed; the second B*P
 * entries are scratch (each object's points sorted by voxel, then index). */
int mf_cnn_voxelize_s2d(const float* feat2 /*[B*P,C]*/, const float* points /*[B,3,P]*/, int B,
                        int P, int C, int D, int Ct, int32_t* prev_keys, void* X, void* stream);
/* mf_cnn_point_mlp + phase 1 of mf_cnn_voxelize_s2d_phase (sparse clear of the previous call's
 * voxels + this call's keys / sorted order) in ONE launch: the bookkeeping CTAs are dispatched
 * first and run beside the MLP CTAs.  P <= 4096 (else MF_E_UNSUPPORTED: use the two calls).
 * Follow with mf_cnn_voxelize_s2d_phase(..., phases = 2). */
int mf_cnn_point_mlp_voxkeys(const float* values, const float* points, const float* w1_rgb,
                             const float* b1_rgb, const float* w1_pcd, const float* b1_pcd,
                             const float* w2_rgb, const float* b2_rgb, const float* w2_pcd,
                             const float* b2_pcd, int B, int P, float center, void* feat, int ldf,
                             float* feat2, int C, int D, int Ct, int32_t* prev_keys, void* X,
                             void* stream);
/* the same in two halves so that a caller can overlap the first with the point MLP that produces
 * feat2: phases bit 0 = sparse clear + new keys (feat2 may be NULL), bit 1 = ordered scatter */
int mf_cnn_voxelize_s2d_phase(const float* feat2, const float* points, int B, int P, int C, int D,
                              int Ct, int32_t* prev_keys, void* X, int phases, void* stream);
int mf_cnn_pack_s2d(const float* vox /*[B,C,D,D,D]*/, const float* hocc /*[B,V,Cocc] or NULL*/,
                    int B, int C, int Cocc, int D, void* X /*bf16 s2d, borders pre-zeroed*/,
                    void* stream);
int mf_gemm_bf16_simt(const GemmParams* p, void* stream);
/* up to 3 GEMMs in one launch (the three pose heads); tc variant needs identical M,N,K,mode */
int mf_gemm_bf16_simt_grouped(const GemmParams* p, int n_groups, void* stream);
int mf_gemm_bf16_tc_grouped(const GemmParams* p, int n_groups, void* workspace,
                            size_t workspace_bytes, void* stream);
/* tcgen05/TMEM/TMA kernel; returns MF_E_UNSUPPORTED for shapes it does not cover (N % 32, N < 128,
 * K < 64, unaligned views) -- the caller then uses mf_gemm_bf16_simt.  `workspace` holds fp32
 * [M][N] partial sums when the launch splits K (few output tiles, long K). */
size_t mf_gemm_bf16_tc_workspace_bytes(int M, int N);
int mf_gemm_bf16_tc(const GemmParams* p, void* workspace, size_t workspace_bytes, void* stream);
/* Extended entry.
 *  sync   : NULL, or MF_GEMM_TC_SYNC_INTS int32 words owned by the caller, zeroed ONCE before the
 *           first call (the kernel leaves them zero).  With it, bf16-output GEMMs (grouped or
 *           not) whose tile count would idle > 5 % of the SMs run "stream-K": the
 *           tiles x K-blocks space is cut into one equal range per SM, a CTA that starts inside
 *           a tile parks its fp32 accumulator in `workspace` (one 128 x 256 slot per SM) and the
 *           tile's first CTA adds the slots in CTA order in its epilogue: every SM does the same
 *           number of K blocks and there is no separate reduction launch.  Two launches that
 *           may run concurrently must not share `sync` / `workspace`.
 *  stamps : NULL, or a device buffer of 16*8 int64 receiving clock64 stamps of CTA 0's roles per
 *           work unit (0/1 TMA producer begin/end, 2/3/4 MMA issuer buffer-free /
 *           first-operands / committed, 5/6/7 epilogue accumulator-ready / drained / stored).
 *  one_shot : 0 = best kernel for the shape (the CTA-pair cta_group::2 kernel for single-group
 *           bf16-output problems with N >= 256, else the single-CTA persistent kernel);
 *           1 = the one-tile-per-CTA kernel, 2 = the single-CTA persistent kernel (baselines). */
#define MF_GEMM_TC_SYNC_INTS 2048
int mf_gemm_bf16_tc_ex(const GemmParams* p, int n_groups, void* workspace, size_t workspace_bytes,
                       int32_t* sync, long long* stamps, int one_shot, void* stream);
/* Layers 1-3 of the three pose heads (model.py:239-254) as ONE persistent launch.  layers[0] =
 * the three first layers side by side (N = 3 x 640), layers[1..3] = conv2_{rot,trans,conf}
 * reading their column block of layer 0's output, layers[4..6] = conv3_* reading layers[1..3]'s
 * outputs; all GEMM_LINEAR, OUT_BF16, same M.  Tiles of later layers start as soon as the
 * 128-row tiles they read are stored (global arrival counters), so the three layers share one
 * wave structure and there are no launch gaps.  sync: MF_HEADS_SYNC_INTS int32 words owned by the
 * caller, zeroed ONCE (never reset afterwards: an epoch word makes the counters monotonic). */
#define MF_HEADS_SYNC_INTS 4096
int mf_cnn_heads_tc(const GemmParams* layers, int n_layers /* 7 */, int32_t* sync, void* stream);
int mf_cnn_interp_cl(const void* grid_bf16, int s2d, const float* points /*[B,3,P]*/, int B,
                     int P, int C, int D, float divisor, void* feat_bf16, int ldf, int col_off,
                     void* stream);
int mf_cnn_pose(const float* out_rot, const float* out_trans, const float* out_conf,
                const float* points, const int32_t* class_id, const float* pitch,
                const float* origin, int B, int P, int n_fg_class, float* rot /*[B,P,4]*/,
                float* trans /*[B,P,3]*/, float* conf /*[B,P]*/, void* stream);
/* last head layer (conv4_{rot,trans,conf}, model.py:249-254) fused with the pose epilogue: only
 * the 4+3+1 output rows of each object's class are evaluated.  hd3 = bf16 [B*P, ld] holding the
 * rot | trans | conf 128-channel blocks; weights bf16 [rows,128] row-major, biases fp32.  class_id
 * is 1-based and must lie in [1, nfg] (as in the reference's indexing, model.py:256-262). */
int mf_cnn_head4_pose(const void* hd3, int ld, const void* w_rot, const float* b_rot,
                      const void* w_trans, const float* b_trans, const void* w_conf,
                      const float* b_conf, const float* points, const int32_t* class_id,
                      const float* pitch, const float* origin, int B, int P, int nfg, float* rot,
                      float* trans, float* conf, void* stream);

/* ------------------------------------------------------------------------
 * a8 / a9  IterativeCollisionCheckLink + its Adam loop, fused
 *     replaces IterativeCollisionCheckLink.forward
 *       (morefusion/contrib/iterative_collision_check_link.py:31-99), chainer's reverse pass
 *       through it and chainer.optimizers.Adam.update as driven by
 *       examples/ycb_video/pose_refinement/check_iterative_collision_check_link.py:44-79
 * One persistent cooperative kernel runs all `n_iter` (<= 128 per call) iterations for a BATCH of
 * independent scenes; scene s owns objects [scene_obj_off[s], scene_obj_off[s+1]) (<= 32 per
 * scene) and `group_size` CTAs.  Points/sdf are concatenated over all objects (obj_pt_off);
 * work tables: 256-point chunks (chunk_obj / chunk_start, grouped by scene via scene_chunk_off),
 * scene_slot_off = prefix sums of N_s * C_s.  All tables are DEVICE int32 arrays.
 *   update == 0 : one forward/backward, parameters untouched (loss_history, grads written)
 *   update == 1 : Chainer-form Adam, m += (1-b1)(g-m); v += (1-b2)(g*g-v);
 *                 p -= eta * alpha_t[it] * m / (sqrt(v) + eps); alpha_t given per iteration by the
 *                 caller (HOST float arrays) for quaternion and translation separately.
 * adam_state: device [m_q(N,4) | m_t(N,3) | v_q(N,4) | v_t(N,3)].  grads: device [N,7] = gq | gt
 * of the last iteration.  loss_history: device [n_scenes, n_iter].
 * ------------------------------------------------------------------------ */
int mf_icc_max_group_size(int n_scenes);
size_t mf_icc_workspace_bytes(int n_objects_total, int voxel_dim, int n_scenes, int group_size,
                              int n_slots);
int mf_icc_run(int n_scenes, int n_objects_total, int voxel_dim, float voxel_threshold,
               float sdf_offset, const int32_t* scene_obj_off, const int32_t* obj_pt_off,
               const int32_t* scene_chunk_off, const int32_t* chunk_obj,
               const int32_t* chunk_start, const int32_t* scene_slot_off,
               const int32_t* obj_chunk_off /*[N+1] first chunk of each object*/, int n_slots,
               const float* points, const float* sdf, const float* pitch, const float* origin,
               const float* grid_target, const float* grid_nontarget_empty, float* quaternion,
               float* translation, float* adam_state, int n_iter, int update,
               const float* alpha_q_host, const float* alpha_t_host, double beta1, double beta2,
               double eps, double eta, float* loss_history, float* grads, int group_size,
               void* workspace, size_t workspace_bytes, void* stream);
/* same, plus phase_ns: device uint64 [n_scenes][n_iter][8] receiving %globaltimer stamps at the
 * phase boundaries of every iteration (profiling aid; NULL = off) */
int mf_icc_run_profiled(int n_scenes, int n_objects_total, int voxel_dim, float voxel_threshold,
               float sdf_offset, const int32_t* scene_obj_off, const int32_t* obj_pt_off,
               const int32_t* scene_chunk_off, const int32_t* chunk_obj,
               const int32_t* chunk_start, const int32_t* scene_slot_off,
               const int32_t* obj_chunk_off, int n_slots,
               const float* points, const float* sdf, const float* pitch, const float* origin,
               const float* grid_target, const float* grid_nontarget_empty, float* quaternion,
               float* translation, float* adam_state, int n_iter, int update,
               const float* alpha_q_host, const float* alpha_t_host, double beta1, double beta2,
               double eps, double eta, float* loss_history, float* grads, int group_size,
               void* workspace, size_t workspace_bytes, unsigned long long* phase_ns, void* stream);

/* ------------------------------------------------------------------------
 * a10 / a11 fp32-class parity mode ("bf16x3"): every GEMM operand is split hi + lo (bf16 each)
 * and A.W is evaluated as A_hi.W_hi + A_lo.W_hi + A_hi.W_lo on the tcgen05 GEMMs into fp32
 * slices (mf_gemm_bf16_tc with OUT_F32); these helpers split / combine / gather in fp32
 * (csrc/precise.cu).  Same reference lines as the bf16 path (model.py:93-141, :239-273).
 * ------------------------------------------------------------------------ */
int mf_cnn_point_mlp_f32(const float* values, const float* points, const float* w1_rgb,
                         const float* b1_rgb, const float* w1_pcd, const float* b1_pcd,
                         const float* w2_rgb, const float* b2_rgb, const float* w2_pcd,
                         const float* b2_pcd, int B, int P, float center, void* feat_bf16, int ldf,
                         float* feat2 /*[B*P,144]*/, float* feat1 /*[B*P,72]*/, void* stream);
int mf_px_split(const float* src, long long lds, long long M, int N, void* hi, void* lo,
                long long ldd, int col_off, void* stream);
int mf_px_combine(const float* ws /*[n_slices][M][N]*/, int n_slices, long long M, int N,
                  const float* bias, int relu, int out_mode /*OUT_BF16 | OUT_S2D_BF16*/, int Do,
                  void* hi, void* lo, long long ldo, int col_off, void* stream);
int mf_px_pack_s2d(const float* vox, const float* hocc, int B, int C, int Cocc, int D, void* X_hi,
                   void* X_lo, void* stream);
int mf_px_interp(const void* grid_hi, const void* grid_lo, int s2d, const float* points, int B,
                 int P, int C, int D, float divisor, void* feat_hi, void* feat_lo, int ldf,
                 int col_off, void* stream);
int mf_px_head4_pose(const void* hd3_hi, const void* hd3_lo, int ld, const float* w_rot,
                     const float* b_rot, const float* w_trans, const float* b_trans,
                     const float* w_conf, const float* b_conf, const float* points,
                     const int32_t* class_id, const float* pitch, const float* origin, int B, int P,
                     int nfg, float* rot, float* trans, float* conf, void* stream);

/* ------------------------------------------------------------------------
 * a10 / a11 backward: training step of the 3-D section of singleview_3d.Model
 *     replaces chainer's reverse pass through morefusion/contrib/singleview_3d/models/model.py
 *     :93-141 (_extract) and :239-273 (heads, pose) -- cuDNN backward-data / backward-filter,
 *     interpolate_voxel_grid.py:216-268, average_voxelization_3d.py:147-220 -- and the
 *     ChainerMN-side optimizer step (examples/ycb_video/singleview_3d/train.py:342-344).
 * Activations are the forward's channels-last bf16 buffers; gradients that feed a GEMM are bf16,
 * accumulators and parameter gradients fp32.  All outputs named d* that are described as
 * "accumulated" must be zeroed by the caller once per step.
 * ------------------------------------------------------------------------ */
/* forward variant of mf_cnn_head4_pose that also saves the 8 selected pre-activation outputs */
int mf_cnn_head4_pose_train(const void* hd3, int ld, const void* w_rot, const float* b_rot,
                            const void* w_trans, const float* b_trans, const void* w_conf,
                            const float* b_conf, const float* points, const int32_t* class_id,
                            const float* pitch, const float* origin, int B, int P, int nfg,
                            float* rot, float* trans, float* conf, float* raw8 /*[B*P,8]*/,
                            void* stream);
/* tcgen05, both operands MN-major: out[g][n, k] (+)= sum_m dZ[m, g*dz_group_cols + n] *
 * X[m, g*x_group_cols + k]   (Conv1D weight gradients; bf16 operands, fp32 output) */
int mf_train_gemm_tn(const void* dZ, long long ldz, const void* X, long long ldx, int m_rows,
                     int n_out, int k_in, float* out, long long ldo, int n_groups,
                     long long dz_group_cols, long long x_group_cols, long long out_group_stride,
                     int accumulate, void* stream);
/* k4 s2 p1 Conv3D weight gradient in the space-to-depth form of the forward: dYp bf16
 * [B, Do+2, Do+2, Do+2, Co] (zero border), Xs2d bf16 [B, Do+1, Do+1, Do+1, C8]; out fp32
 * [Co][8 * C8] in the packed K order of the forward weights (tap-major) */
int mf_train_conv_wgrad(const void* dYp, const void* Xs2d, int batch, int Do, int Co, int C8,
                        float* out, int accumulate, void* stream);
/* k4 s2 p1 Conv3D input gradient: 8 parity GEMMs over the padded dY; Wd bf16 [8][Ci][8*Co]
 * (per parity r: Wd[r][ci][a*Co+co] = W[co][ci][2a+r]).  epilogue 1: added into a channels-last
 * fp32 grid [B,(2Do)^3,ldo]; epilogue 2: bf16 rows [8][B*Do^3][ldo] in parity order */
int mf_train_conv_dgrad(const void* dYp, const void* Wd, int batch, int Do, int Co, int Ci,
                        int epilogue, void* out, long long ldo, long long out_group_stride,
                        void* stream);
int mf_train_head4_bwd(const float* g_rot, const float* g_trans, const float* g_conf,
                       const float* raw8, const void* hd3, int ld, const void* w_rot,
                       const void* w_trans, const void* w_conf, const int32_t* class_id,
                       const float* pitch, int B, int P, int nfg, void* dhd3 /*bf16, masked*/,
                       float* dw_rot, float* db_rot, float* dw_trans, float* db_trans,
                       float* dw_conf, float* db_conf /*accumulated*/, void* stream);
int mf_train_relu_mask(void* x /*bf16 in/out*/, long long ldx, const void* act, long long lda,
                       long long M, int N, void* stream);
int mf_train_colsum(const void* x /*bf16*/, long long ldx, long long M, int N,
                    float* out /*accumulated*/, void* stream);
int mf_train_interp_bwd(const void* g /*bf16 [B*P, ldg]*/, long long ldg, int col_off,
                        const float* points /*[B,3,P]*/, int B, int P, int C, int D,
                        float divisor, float* dgrid /*[B,D^3,C] accumulated*/, void* stream);
int mf_train_mask_pack(const float* dgrid, const void* act, int act_s2d, int B, int D, int C,
                       void* out /*bf16 [B,(D+2)^3,C], interior written*/,
                       float* dbias /*accumulated, may be NULL*/, void* stream);
int mf_train_vox_bwd(const void* dx3 /*bf16 parity rows*/, long long ldp, const int32_t* keys,
                     int B, int P, int C, int D, const void* g_direct, long long ldg, int col_off,
                     float* dfeat2 /*[B*P, C]*/, void* stream);
int mf_train_point_mlp_bwd(const float* values, const float* points, const void* feat,
                           long long ldf, const void* dfeat, long long ldd, const float* feat2,
                           const float* g2, const float* w1r, const float* w1p, const float* w2r,
                           const float* w2p, int B, int P, float center, float* dw1r,
                           float* db1r, float* dw1p, float* db1p, float* dw2r, float* db2r,
                           float* dw2p, float* db2p /*accumulated*/,
                           float* dvalues /*[B,32,P], may be NULL*/, void* stream);
/* fused gradient unscale (x grad_scale, e.g. 1/world after a sum all-reduce) + chainer Adam on
 * flat fp32 buffers */
int mf_train_adam(float* params, const float* grads, float* m, float* v, int64_t n,
                  float alpha_t, double beta1, double beta2, double eps, double eta,
                  float grad_scale, void* stream);

/* ------------------------------------------------------------------------
 * f2  per-frame front end: replaces the host NumPy helpers
 *     morefusion/geometry/pointcloud_from_depth.py:4-26 and masks_to_bboxes.py:4-38
 * depth [H,W] f32 metres (NaN = invalid) -> pcd [H,W,3]; masks [N,H,W] u8 -> bboxes [N,4] int32
 * (y1, x1, y2, x2), upper bounds exclusive, zeros for an empty mask.
 * ------------------------------------------------------------------------ */
int mf_pointcloud_from_depth(const float* depth, int H, int W, float fx, float fy, float cx,
                             float cy, int euclidean, float* pcd, void* stream);
int mf_masks_to_bboxes(const uint8_t* masks, int N, int H, int W, int32_t* bboxes, void* stream);

/* ------------------------------------------------------------------------
 * f1  2-D extractor tail at the sampled pixels only: replaces the dense
 *     up3 (resize x2 + Conv2D 64->64 3x3 + PReLU), Conv2D 64->32 1x1 and log_softmax of
 *     morefusion/models/dense_fusion/pspnet.py:64-82 followed by the 1000-pixel gather of
 *     contrib/singleview_3d/models/model.py:222.
 * up2_nhwc [B,Hs,Ws,64] f32 channels-last (output of up2), pix [B,P] int64 row-major pixel index
 * into the 2Hs x 2Ws image, w3t [576,64] = W3[co][ci][dy][dx] transposed to [(ci,dy,dx)][co],
 * b3 [64], prelu_slope [1] (device), w1t [64,32] = W1 transposed, b1 [32] -> out [B,32,P] f32.
 * ------------------------------------------------------------------------ */
int mf_psp_tail_sampled(const float* up2_nhwc, const int64_t* pix, int B, int P, int Hs, int Ws,
                        const float* w3t, const float* b3, const float* prelu_slope,
                        const float* w1t, const float* b1, float* out, void* stream);

/* ------------------------------------------------------------------------
 * f3  occupancy-grid producer: replaces the OctoMap trees behind
 *     morefusion/contrib/multi_instance_octree_mapping.py:16-34 (initialize / integrate / update)
 *     and :35-94 (get_target_grids).  One caller-owned open-addressing hash table holds the cells
 *     of every instance: cells [capacity] 16-byte entries {u64 key, f32 log-odds, u32 scan stamp}
 *     (16-byte aligned; initialised ONCE to key = all ones, rest zero), cnt u32 [capacity] and
 *     counters int32 [8] (zero at allocation; the kernels keep them consistent afterwards).
 *     capacity is a power of two.  counters[0] = live cells, counters[1] = table full (sticky;
 *     updates were dropped).  `instance` is the caller's dense index (order of initialize()).
 * mf_map_integrate: octree.insertPointCloud(points[mask], origin) (:24): points [n,3] f32 (NaN rows
 *     and rows with mask == 0 are skipped; mask may be NULL), sensor origin as floats, OctoMap's
 *     sensor model as float32 log-odds; scan = 1, 2, ... increasing per table.  Two launches, no
 *     scratch memory.
 * mf_map_integrate_labelled: the same for every instance of a labelled frame in ONE scan (what
 *     datasets/rgbd_pose_estimation/base.py:30-50 and the ROS OctomapServer::insertScan,
 *     ros/.../OctomapServer.cpp:286-395, do instance by instance): labels [n] int32; lut [lut_n]
 *     int32 maps label - lut_lo to the dense instance index (-1 = not mapped: pixel skipped);
 *     inst_resolution [n_instances] f64 -- device arrays.  Equal to one mf_map_integrate per
 *     instance with mask = (labels == id), the instances' cells being disjoint.
 * mf_map_update_points: octree.updateNodes(points, True) (:30): points [m,3] f64, update = the
 *     log-odds increment applied once per row.
 * mf_map_query_grids: get_target_grids for T targets in one launch: target_index [T] int32 (dense
 *     instance index of each target), pitch [T] f64, origin [T,3] f64, res_factor [n_instances] f64
 *     (= 1 / resolution) -- device arrays; outputs [T,X,Y,Z] f32, every element written.
 * mf_map_rehash: move every cell of the old table into a larger, freshly initialised one.
 * ------------------------------------------------------------------------ */
int mf_map_integrate(const float* points, const uint8_t* mask, int64_t n, float ox, float oy,
                     float oz, double resolution, int instance, uint32_t scan, float hit,
                     float miss, float lo_min, float lo_max, void* cells, void* cnt,
                     int64_t capacity, int32_t* counters, void* stream);
int mf_map_integrate_labelled(const float* points, const int32_t* labels, int64_t n, float ox,
                              float oy, float oz, const int32_t* lut, int lut_lo, int lut_n,
                              const double* inst_resolution, uint32_t scan, float hit, float miss,
                              float lo_min, float lo_max, void* cells, void* cnt, int64_t capacity,
                              int32_t* counters, void* stream);
int mf_map_update_points(const double* points, int64_t m, double resolution, int instance,
                         float update, float lo_min, float lo_max, void* cells, void* cnt,
                         int64_t capacity, int32_t* counters, void* stream);
int mf_map_query_grids(const int32_t* target_index, const double* pitch, const double* origin,
                       int T, int X, int Y, int Z, const double* res_factor, int n_instances,
                       void* cells, void* cnt, int64_t capacity, int32_t* counters,
                       float* grid_target, float* grid_nontarget, float* grid_empty, void* stream);
int mf_map_rehash(const void* old_cells, int64_t old_capacity, void* cells, void* cnt,
                  int64_t capacity, int32_t* counters, void* stream);

/* ------------------------------------------------------------------------
 * a12  average_distance (ADD / ADD-S training loss) + nearest neighbour
 *     replaces morefusion/functions/loss/average_distance.py:40-85 and, for symmetric=1,
 *     morefusion/geometry/knn/nn.py:17-48 + cuComputeDistanceGlobal.cu:20-86 (matrix-free here)
 * points [P,3], transform_true [4,4], transforms_pred [M,4,4] -> out [M];
 * nn_indices [M,P] int32 (may be NULL when symmetric == 0) is kept for the backward pass.
 * ------------------------------------------------------------------------ */
int mf_average_distance_fwd(const float* points, int n_points, const float* transform_true,
                            const float* transforms_pred, int n_pred, int symmetric, float* out,
                            int32_t* nn_indices, void* stream);
/* same forward with the query points split over n_parts CTAs per pose (few poses, many points:
 * the evaluation metrics); out_parts [n_pred, n_parts] holds partial means, their sum is the
 * distance */
int mf_average_distance_fwd_parts(const float* points, int n_points, const float* transform_true,
                                  const float* transforms_pred, int n_pred, int symmetric,
                                  int n_parts, float* out_parts, int32_t* nn_indices, void* stream);
int mf_average_distance_bwd(const float* gout, const float* points, int n_points,
                            const float* transform_true, const float* transforms_pred, int n_pred,
                            const int32_t* nn_indices /*NULL: identity*/,
                            float* g_transforms_pred /*[M,4,4]*/, float* g_transform_true /*[4,4]*/,
                            float* workspace /*[n_pred,12] floats*/, void* stream);
/* the same over the B objects of a training batch from one call (model.py:416-431 calls
 * average_distance per object): points [B,n_points,3], transform_true [B,4,4], transforms_pred
 * [B,n_pred,4,4], out [B,n_pred], nn_indices [B,n_pred,n_points]; symmetric = HOST array [B] */
int mf_average_distance_fwd_batched(const float* points, int n_points, const float* transform_true,
                                    const float* transforms_pred, int n_pred, int B,
                                    const int32_t* symmetric_host, float* out, int32_t* nn_indices,
                                    void* stream);
int mf_average_distance_bwd_batched(const float* gout, const float* points, int n_points,
                                    const float* transform_true, const float* transforms_pred,
                                    int n_pred, int B, const int32_t* symmetric_host,
                                    const int32_t* nn_indices, float* g_transforms_pred,
                                    float* g_transform_true, float* workspace /*[B,n_pred,12]*/,
                                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOREFUSION_B200_H_ */
