/*
 * msr3d_b200 — C ABI of the B200-native MSR3D hot path (libmsr3d_b200.so).
 *
 * Plain pointers and sizes only; all pointers are DEVICE pointers unless a
 * parameter is documented as host.  Every entry point is asynchronous on the
 * given stream, allocates nothing the caller can see, never calls exit(), and
 * returns an int status (MSR3D_OK == 0).  msr3d_last_error() returns a
 * thread-local human-readable string for the last non-zero status.
 *
 * Reference interfaces replaced (paths relative to the MSR3D reference tree):
 *   group A  modules/third_party/pointnet2/_ext_src/src/bindings.cpp:6-19
 *            (pybind module `pointnet2._ext`, nine functions) and the host
 *            wrappers in sampling.cpp:15-87, ball_query.cpp:8-32,
 *            group_points.cpp:12-62, interpolate.cpp:14-99.
 *   group B  modules/third_party/pointnet2/pointnet2_modules.py:34-75
 *            (_PointnetSAModuleBase.forward) + pytorch_utils.py:11-36,157-188
 *            (SharedMLP = conv1x1 + eval-BN + ReLU) + modules/layers/pointnet.py:55-63.
 *   group C  modules/layers/transformers.py:200-252,314-329,
 *            modules/utils.py:60-137, model/ose3d_situation.py:31-59,368-435.
 *   group D  trainer/leo_trainer.py:189-195 + optim/build.py:7-17 (clip + AdamW).
 *
 * Contract differences from the reference, on purpose:
 *   - the caller owns every buffer (the reference allocates with torch::zeros);
 *     entry points that need zero-initialised outputs do the memset themselves;
 *   - errors come back as status codes (cuda_utils.h:30-39 calls exit(-1));
 *   - furthest_point_sampling needs no `temp` scratch (kept in registers);
 *   - gather_points_grad / group_points_grad are deterministic (atomic-free segmented reductions; the
 *     reference kernels use atomicAdd) whenever one cloud's index table fits in shared memory.
 */
#ifndef MSR3D_B200_H
#define MSR3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudaStream_t without pulling cuda_runtime.h into C callers */
typedef void *msr3d_stream_t;

enum {
  MSR3D_OK = 0,
  MSR3D_ERR_INVALID_ARG = 1,
  MSR3D_ERR_UNSUPPORTED = 2,
  MSR3D_ERR_CUDA = 3
};

const char *msr3d_last_error(void);
/* ABI version of this header; bumped on any signature change. */
int msr3d_abi_version(void);
/* Number of kernel launches issued through this library by the calling process. */
unsigned long long msr3d_launch_count(void);
/* Programmatic dependent launch (default on; env MSR3D_PDL=0 disables): kernels are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and wait with griddepcontrol.wait before their
 * first global access, so consecutive small kernels overlap launch latency and set-up code. */
void msr3d_set_pdl(int enabled);
int msr3d_get_pdl(void);
/* Tuning hook for tools/gemm_sweep.py: overrides one launch heuristic (0 = GEMM N-tile 64/128,
 * 1 = GEMM split-K factor, 2/3 = warps per CTA of the attention forward/backward kernels,
 * 4 = row blocks per (head, scene) of the attention forward); value 0 restores the heuristic. */
int msr3d_set_tunable(int id, int value);

/* ------------------------------------------------------------------ group A
 * The nine pointnet2._ext operators.  Layouts and semantics are exactly the
 * reference's; int32 outputs are bit-identical to the reference CUDA kernels
 * built with nvcc's default -fmad=true.                                     */

/* sampling.cpp:66-87 / sampling_gpu.cu:69-173.  xyz (B,N,3) -> idx (B,m). */
int msr3d_furthest_point_sampling_f32(const float *xyz, int32_t *idx, int B, int N,
                                      int m, msr3d_stream_t stream);
/* sampling.cpp:15-38 / sampling_gpu.cu:8-20.  points (B,C,N), idx (B,M) -> out (B,C,M). */
int msr3d_gather_points_f32(const float *points, const int32_t *idx, float *out, int B,
                            int C, int N, int M, msr3d_stream_t stream);
/* sampling.cpp:40-65 / sampling_gpu.cu:34-47.  grad_out (B,C,M) -> grad_points (B,C,N), zeroed here. */
int msr3d_gather_points_grad_f32(const float *grad_out, const int32_t *idx,
                                 float *grad_points, int B, int C, int N, int M,
                                 msr3d_stream_t stream);
/* ball_query.cpp:8-32 / ball_query_gpu.cu:9-44.  new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample). */
int msr3d_ball_query_f32(const float *new_xyz, const float *xyz, int32_t *idx, int B,
                         int N, int M, float radius, int nsample, msr3d_stream_t stream);
/* group_points.cpp:12-36 / group_points_gpu.cu:8-28.  points (B,C,N), idx (B,P,S) -> out (B,C,P,S). */
int msr3d_group_points_f32(const float *points, const int32_t *idx, float *out, int B,
                           int C, int N, int npoint, int nsample, msr3d_stream_t stream);
/* group_points.cpp:38-62 / group_points_gpu.cu:43-64.  grad_out (B,C,P,S) -> grad_points (B,C,N), zeroed here. */
int msr3d_group_points_grad_f32(const float *grad_out, const int32_t *idx,
                                float *grad_points, int B, int C, int N, int npoint,
                                int nsample, msr3d_stream_t stream);
/* interpolate.cpp:14-40 / interpolate_gpu.cu:9-59.  unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3), idx (B,n,3). */
int msr3d_three_nn_f32(const float *unknown, const float *known, float *dist2,
                       int32_t *idx, int B, int n, int m, msr3d_stream_t stream);
/* interpolate.cpp:42-69 / interpolate_gpu.cu:72-101.  points (B,c,m), idx/weight (B,n,3) -> out (B,c,n). */
int msr3d_three_interpolate_f32(const float *points, const int32_t *idx,
                                const float *weight, float *out, int B, int c, int m,
                                int n, msr3d_stream_t stream);
/* interpolate.cpp:71-99 / interpolate_gpu.cu:116-143.  grad_out (B,c,n) -> grad_points (B,c,m), zeroed here. */
int msr3d_three_interpolate_grad_f32(const float *grad_out, const int32_t *idx,
                                     const float *weight, float *grad_points, int B,
                                     int c, int n, int m, msr3d_stream_t stream);

/* ------------------------------------------------------------------ group B
 * PointNet++ set abstraction on the tensor cores (tcgen05, bf16 hi/lo split
 * operands, f32 accumulation in TMEM).  Replaces _PointnetSAModuleBase.forward
 * (pointnet2_modules.py:34-75), QueryAndGroup (pointnet2_utils.py:314-373) and
 * SharedMLP (pytorch_utils.py:11-36) with BatchNorm in eval mode folded into
 * the packed weights (scale) and the bias (shift).                           */

/* Bytes of the packed image of an (N x K) weight for tiles of BN rows x BK k. */
long long msr3d_packed_weight_bytes(int N, int K, int BN, int BK);
/* Packs W (element (n,k) at W[n*stride_n + k*stride_k]) into the shared-memory
 * image tcgen05.mma consumes: blocks [n_tile][k_block] of (BN x BK) bf16 hi tile
 * followed by the lo tile.  row_scale (N, nullable) multiplies row n first
 * (BatchNorm fold); col_perm (Kp int32, nullable) maps packed column -> source
 * column, -1 = zero.  nn.Linear / 1x1-conv weights are (N=out, K=in).         */
int msr3d_pack_weight_f32(const float *W, int N, int K, long long stride_n, long long stride_k,
                          const float *row_scale, const int32_t *col_perm, int BN, int BK,
                          void *packed, msr3d_stream_t stream);
/* out[M,N] (+)= act(A[M,K] * W^T + bias), W packed with BN=128, BK=64.
 * A element (m,k) at A[m*stride_am + k*stride_ak].  act: 0 none, 1 ReLU, 2 GELU(erf).
 * pool > 1: max over groups of `pool` consecutive rows (power of two <= 32), out is
 * (M/pool, N).  split: 3 = fp32-grade split-bf16, 1 = single bf16 pass.
 * accumulate: out += result (act 0, no pool).  Replaces cuDNN conv1x1 / cuBLAS
 * calls of pytorch_utils.py:157-188 and modules/layers/pointnet.py:53,63.    */
int msr3d_linear_tc_f32(const float *A, long long stride_am, long long stride_ak, int M, int K,
                        const void *packed_w, const float *bias, float *out, long long ldo, int N,
                        int act, int pool, int split, int accumulate, int tile_n, msr3d_stream_t stream);
/* Same contraction (out = act(A W^T + b), optional max over `pool` consecutive rows; packed weights with 128-wide
 * blocks, split-bf16 hh + hl + lh) for CHAINS of dense layers: the activation operand may be given PRE-SPLIT as two
 * bf16 tensors (A_hi, A_lo: (M, K), row pitch lda elements; A = NULL) and the result may be written as bf16 hi / lo
 * (out_hi, out_lo: (M, N), row pitch ldo elements; out = NULL) for the next layer — the fp32 -> bf16 split is then
 * done once, by the producer's epilogue, instead of by every CTA that reads the tile (SA3 of the frozen encoder:
 * pointnet2_modules.py:34-75 GroupAll MLP 259 -> 256 -> 512 -> 768). */
int msr3d_linear_bf16x3_f32(const float *A, const void *A_hi, const void *A_lo, long long lda, int M, int K,
                            const void *packed_w, const float *bias, float *out, void *out_hi, void *out_lo,
                            long long ldo, int N, int act, int pool, msr3d_stream_t stream);
/* General tensor-core GEMM with fp32 (strided) operands on both sides:
 *   out[M,N] (+)= epilogue( act( A[M,K] * B[N,K]^T + bias ) ) + residual
 * Used by the spatial transformer's forward and backward (weights and their transposes are
 * read in place through strides; nothing is pre-packed).  Element (m,k) of A at
 * A[m*stride_am + k*stride_ak]; element (n,k) of B at B[n*stride_bn + k*stride_bk].
 *   a_xform / b_xform: must be 0 (operand transforms are retired: MSR3D_ERR_UNSUPPORTED otherwise)
 *   epilogue: 0 plain; 1 retired;
 *             2 (FFN forward) out = dropout(gelu(x)), aux_out = gelu'(x) * mask(epi_site);  3 x *= aux[m,n]
 *   accumulate: 0 store, 1 out += (single CTA per element), 2 atomicAdd (required for ksplit > 1)
 *   ksplit: number of CTAs along K (split-K)
 *   split: 1 = one bf16 MMA per K slice, 3 = two-way split operands (hh+hm+mh, ~2^-17),
 *          6 = three-way split (adds hl+lh+mm, fp32-grade; used for weight gradients) */
typedef struct msr3d_gemm_desc {
  const float *A; long long stride_am, stride_ak;
  const float *B; long long stride_bn, stride_bk;
  const float *bias;       /* [N] or NULL */
  const float *residual;   /* [M, ld_residual] or NULL */
  long long ld_residual;
  const float *aux;        /* [M, ld_aux] for epilogue 3 */
  long long ld_aux;
  float *aux_out;          /* [M, ld_aux] second output of epilogue 2 */
  float *out; long long ld_out;
  int M, N, K;
  int act, split, accumulate, epilogue, ksplit;
  int a_xform, b_xform;
  unsigned a_site, b_site, epi_site;
  float dropout_p;
  unsigned long long seed;
  const unsigned long long *seed_ptr; /* optional device counter added to seed (graph replays) */
  /* ABI 2: scattered / typed output (plain store only).  row_map[m] = output row of GEMM row m
   * (negative: row dropped), NULL = identity; out_dtype 0 f32, 1 f16, 2 bf16 (`out` then points
   * to that type and ld_out counts its elements). */
  const int32_t *row_map;
  int out_dtype;
  int tile_n;   /* N tile: 0 = library heuristic, 64 or 128 (few large CTAs keep SMs free for a concurrent chain) */
} msr3d_gemm_desc;
int msr3d_gemm_tc_f32(const msr3d_gemm_desc *desc, msr3d_stream_t stream);
/* sizeof(msr3d_gemm_desc) in the loaded library (a binding's struct mirror must agree). */
int msr3d_sizeof_gemm_desc(void);
/* Same contraction, TMA-fed: both operands are fetched by tensor-map TMA (cp.async.bulk.tensor) straight
 * from the row-major fp32 tensors and multiplied with tcgen05.mma kind::tf32 as big/small pairs ("3xTF32":
 * As*Bb + Ab*Bs + Ab*Bb, fp32-grade, ~2^-22 per product).  Each operand must have one unit stride
 * (stride_ak == 1: stored [m][k]; stride_am == 1: stored [k][m] — transposes need no copy), a 16 B aligned
 * base and a row pitch that is a multiple of 4 floats.  Subset of the descriptor: act 0; epilogue 0, 2, 3;
 * accumulate 0 or 2; no row_map / typed output; `split` is ignored.  b_static != 0 promises that no kernel
 * of the current step writes B (weights): its tiles are then requested before the programmatic-dependent-
 * launch wait.  msr3d_gemm_tf32x3_supported returns 1 when a descriptor qualifies (callers fall back to
 * msr3d_gemm_tc_f32 otherwise, e.g. the K = 63 Fourier projection).
 * Replaces the cuBLAS calls behind nn.Linear in modules/layers/transformers.py:200-252,314-329. */
/* Development hook: device buffer of 16 int64 SM-clock slots per CTA that the TMA-fed kernels fill at their
 * phase boundaries (tools/trace_kernels.py); NULL (default) turns tracing off. */
int msr3d_set_trace_buffer(long long *buf);
int msr3d_gemm_tf32x3_supported(const msr3d_gemm_desc *desc);
int msr3d_gemm_tf32x3_f32(const msr3d_gemm_desc *desc, int b_static, msr3d_stream_t stream);
/* Fused feed-forward pair of one encoder layer (transformers.py:326-327), d_model = 256, one CTA per
 * (128-token tile, `slice`-wide slice of the F hidden units; slice = 32 or 64, F % slice == 0):
 *   forward :  h = dropout(gelu(y2 W1^T + b1));  f_accum += h W2^T (+ b2, added once)
 *              also stores h and gp = gelu'(u) * mask (M x F each) for the backward pass
 *   backward:  du = (df W2) * gp;  dy2_accum += du W1;  also stores du (M x F)
 * The hidden tile stays in shared memory between the two tensor-core GEMMs (3xTF32, TMA-fed, see above);
 * the per-slice partial sums of the second GEMM are reduced with red.global.add, so f_accum must be zeroed
 * (dy2_accum: pre-filled with the residual-path gradient) by the caller.  W1 (F x 256), W2 (256 x F)
 * row-major and 16 B aligned; dropout mask = Philox(seed (+ *seed_ptr), site, m * F + f). */
int msr3d_ffn_pair_fwd_f32(const float *y2, const float *W1, const float *b1, const float *W2, const float *b2,
                           int M, int F, unsigned long long seed, const unsigned long long *seed_ptr, float p,
                           unsigned site, float *h, float *gp, float *f_accum, int slice, msr3d_stream_t stream);
int msr3d_ffn_pair_bwd_f32(const float *df, const float *W1, const float *W2, const float *gp, int M, int F,
                           float *du, float *dy2_accum, int slice, msr3d_stream_t stream);
/* Pre-split variants of the pair: activations and weights as big / small tf32 images (no conversion in the kernel). */
int msr3d_ffn_pair_fwd_ps_f32(const float *y2_big, const float *y2_small, const float *W1_big, const float *W1_small,
                              const float *b1, const float *W2_big, const float *W2_small, const float *b2, int M, int F,
                              unsigned long long seed, const unsigned long long *seed_ptr, float p, unsigned site,
                              float *h, float *gp, float *f_accum, int slice, msr3d_stream_t stream);
int msr3d_ffn_pair_bwd_ps_f32(const float *df_big, const float *df_small, const float *W1_big, const float *W1_small,
                              const float *W2_big, const float *W2_small, const float *gp, int M, int F, float *du,
                              float *dy2_accum, int slice, msr3d_stream_t stream);
/* ------------------------------------------------------------------ group E
 * Scene-token splice into the LLM input (the step right after the path):
 * MSR3D.build_embeds, model/msr3d/msr3d.py:277-287 —
 *     scene_embeds = llm_proj(obj_tokens).to(inputs_embeds.dtype)
 *     inputs_embeds[where(input_ids == scene_sp_token)] = scene_embeds.reshape(-1, D)
 *     attention_mask (-> bool) [same positions] = obj_masks
 * The projection is msr3d_gemm_tc_f32 with row_map / out_dtype (rows land directly in the LLM input,
 * fp16 / bf16 conversion in the epilogue); the two entry points below are the index side and the
 * gather that starts the backward. */
/* rows[k] = flat position b*T + t of the k-th placeholder in row-major order (torch.where order),
 * k < max_rows; entries past the number found are -1; *count = number found (may exceed max_rows).
 * mask_out (B*T bytes, bool) = (attention_mask != 0) with obj_masks[k] written at rows[k];
 * attention_mask (int64) and obj_masks (bool bytes) may be NULL (mask_out untouched). */
int msr3d_scene_rows_i64(const int64_t *input_ids, int B, int T, long long sp_token,
                         const int64_t *attention_mask, const uint8_t *obj_masks, int32_t *rows,
                         int max_rows, int32_t *count, uint8_t *mask_out, msr3d_stream_t stream);
/* dY[m, 0:D] = float(d_embeds[rows[m], 0:D]) (zeros where rows[m] < 0): the slice of the LLM's input
 * gradient that belongs to the scene tokens.  dtype 0 f32, 1 f16, 2 bf16; ld in elements. */
int msr3d_splice_gather_grad(const void *d_embeds, int dtype, long long ld, const int32_t *rows, int M,
                             int D, float *dY, msr3d_stream_t stream);
/* ------------------------------------------------------------------ group F
 * Per-object point-cloud preprocessing on the device (the CPU work in front of the path):
 * MSR3DDataset.preprocess_pcd, data/datasets/msr3d.py:181-215, + the padding of dataset_wrapper.py:151-158.
 * points (total, 6) = xyz + rgb of all objects back to back; offsets (n_obj + 1) int32; sample_idx (n_obj, P)
 * indices INSIDE each object (np.random.choice stays on the host); rot = 9 floats, row-major 3x3
 * (build_rotate_mat) or NULL.  Per object: xyz = xyz @ rot^T; obj_locs = [mean, max - min] over ALL its points;
 * the P sampled points are centred on their own mean and divided by their largest norm (1 if < 1e-6).
 * Slots n_obj .. o_max-1 are padding: points of ones, zero location, mask 0.
 * Outputs obj_fts (o_max, P, 6), obj_locs (o_max, 6), obj_masks (o_max) bytes. */
int msr3d_preprocess_objects_f32(const float *points, const int32_t *offsets, const int32_t *sample_idx,
                                 const float *rot, int n_obj, int o_max, int P, float *obj_fts,
                                 float *obj_locs, uint8_t *obj_masks, msr3d_stream_t stream);
/* FPS + centre gather + ball query of one SA level (pointnet2_modules.py:54-58,
 * pointnet2_utils.py:331).  xyz: (B,N,ld) records whose first 3 floats are the
 * coordinates.  Outputs fps_idx (B,m) i32, new_xyz (B,m,3), ball_idx (B,m,nsample) i32. */
int msr3d_sa_group_f32(const float *xyz, int ld, int B, int N, int m, float radius, int nsample,
                       int32_t *fps_idx, float *new_xyz, int32_t *ball_idx, msr3d_stream_t stream);
/* ------------------------------------------------------------------ group B2
 * Unfrozen-encoder training path (modules/vision/pcd_pointnet_encoder.py:58-62,
 * pytorch_utils.py:39-58 BatchNorm2d in training mode, pointnet2_utils.py:176-224 GroupingOperation.backward ->
 * _ext_src/src/group_points_gpu.cu:43-75).  Everything is row-major [row][channel], row = (cloud, centre, sample):
 * the operand layout of msr3d_gemm_tc_f32, which runs the 1x1 convolutions (Z = X W^T), their data gradients
 * (dX = dZ W) and weight gradients (dW = dZ^T X) on the tensor cores.  csrc/sa_train.cu.
 *   sa_group_rows      X (rows, ldx >= 3 + C) = [xyz[idx] - new_xyz | feat[idx] | 0 pad]; idx == NULL: GroupAll (rows =
 *                      points, raw xyz); ldx rounded up to a multiple of 4 makes X a TMA operand
 *   bn_train_fwd       Y = relu(BN(Z)); batch_stats 1: statistics of this batch (deterministic two-stage reduction),
 *                      running statistics updated in place (momentum, unbiased variance); 0: running statistics.
 *                      stats (2, C) = mean, rstd (kept for the backward); partial: scratch of
 *                      2 * C * msr3d_sa_train_stat_chunks(rows) floats
 *   bn_train_bwd       from dY (grad wrt Y): dgamma / dbeta ACCUMULATED, dZ (may alias dY); bstat (2, C) scratch
 *   maxpool_rows(_bwd) max over the S consecutive rows of a group (first index wins ties), arg (groups, C) int32
 *   sa_group_rows_bwd  dfeat (B, N, C) = segmented sum of dX[:, col0 : col0 + C] over idx (B, pairs) — one CTA per
 *                      cloud, fixed order, no atomics */
int msr3d_sa_train_stat_chunks(long long rows);
int msr3d_sa_group_rows_f32(const float *xyz, int ld_xyz, const float *feat, int ld_feat, int C, const float *new_xyz,
                            const int32_t *idx, int N, int groups_per_cloud, int nsample, long long rows, float *X,
                            int ldx, msr3d_stream_t stream);
int msr3d_bn_train_fwd_f32(const float *Z, long long rows, int C, const float *gamma, const float *beta,
                           float *running_mean, float *running_var, int batch_stats, float eps, float momentum,
                           float *partial, float *stats, float *Y, msr3d_stream_t stream);
int msr3d_bn_train_bwd_f32(const float *dY, const float *Y, const float *Z, long long rows, int C, const float *gamma,
                           const float *stats, int batch_stats, float *partial, float *bstat, float *dgamma,
                           float *dbeta, float *dZ, msr3d_stream_t stream);
int msr3d_maxpool_rows_f32(const float *Y, long long groups, int S, int C, float *out, long long ld_out, int32_t *arg,
                           msr3d_stream_t stream);
int msr3d_maxpool_rows_bwd_f32(const float *dOut, long long ld_dout, const int32_t *arg, long long groups, int S,
                               int C, float *dY, msr3d_stream_t stream);
int msr3d_sa_group_rows_bwd_f32(const float *dX, int ldx, int col0, const int32_t *idx, int B, int pairs, int N, int C,
                                float *dfeat, msr3d_stream_t stream);
/* Writes xyz (rows,3) into columns [col0,col0+3) of a (rows, ld_dst) matrix: the xyz
 * channels GroupAll concatenates (pointnet2_utils.py:407), in the operand of SA3. */
int msr3d_scatter_xyz_cols_f32(const float *xyz, float *dst, long long ld_dst, int col0, int rows,
                               msr3d_stream_t stream);
/* Gather + 3 x (conv1x1+BN+ReLU) + max-pool over nsample as one kernel.
 * feat (B,Nsrc,ld_feat) with nfeat channels used, xyz (B,Nsrc,ld_xyz), new_xyz (B,npoint,3),
 * ball_idx (B,npoint,32).  Layer-1 weights are packed with K order [feat | dxyz | 0]
 * (BN = min(c,128), BK = padded K of that layer).  out (B,npoint,ld_out), point-major.
 * Fused instances: (nfeat+3 <= 16, 64,64,128) and (nfeat+3 in (128,144], 128,128,256);
 * anything else returns MSR3D_ERR_UNSUPPORTED (callers use the layer-wise path). */
int msr3d_sa_mlp_fused_f32(const float *feat, int ld_feat, int nfeat, const float *xyz, int ld_xyz,
                           const float *new_xyz, const int32_t *ball_idx, int B, int Nsrc,
                           int npoint, int nsample, const void *w1, const void *w2, const void *w3,
                           const float *b1, const float *b2, const float *b3, int c1, int c2, int c3,
                           float *out, long long ld_out, int split, msr3d_stream_t stream);

/* ------------------------------------------------------------------ group C
 * Situated spatial-attention transformer, everything that is not a GEMM
 * (modules/layers/transformers.py:200-252,314-329; modules/utils.py:60-137;
 * model/ose3d_situation.py:31-59,399-410).  Hidden size 256, head dim 32, L <= 128.
 * Dropout masks are Philox(seed, site, element offset) and are regenerated, never stored. */

/* calc_pairwise_locs('center', dim 5) + agent-frame Fourier features.  loc (B,L,6);
 * anchor (B,3), quat (B,4 xyzw) used when use_agent_frame; pw (B,L,L,5), ff (B*L,64), sz (B*L,3). */
int msr3d_spatial_geometry_f32(const float *loc, const float *anchor, const float *quat, int B, int L,
                               int dist_norm, int use_agent_frame, float *pw, float *ff, float *sz,
                               msr3d_stream_t stream);
/* qpos = LN_l(h1) + LN_s(Ws sz + bs); qpos0 = qpos + extra_a + extra_b (ose3d_situation.py:399-404,359-363). */
int msr3d_qpos_fwd_f32(const float *h1, const float *sz, int M, const float *Ws, const float *bs,
                       const float *gamma_l, const float *beta_l, const float *gamma_s,
                       const float *beta_s, const float *extra_a, const float *extra_b, float *xhat_l,
                       float *rstd_l, float *xhat_s, float *rstd_s, float *qpos, float *qpos0,
                       msr3d_stream_t stream);
/* Same; qpos0 is additionally ADDED (red.add) into x0_accum (M, 256), the zero-filled layer-0 input that the
 * object-projection GEMM reduces its split-K partial sums into (the two kernels run side by side). */
int msr3d_qpos_fwd_acc_f32(const float *h1, const float *sz, int M, const float *Ws, const float *bs,
                           const float *gamma_l, const float *beta_l, const float *gamma_s, const float *beta_s,
                           const float *extra_a, const float *extra_b, float *xhat_l, float *rstd_l, float *xhat_s,
                           float *rstd_s, float *qpos, float *qpos0, float *x0_accum, msr3d_stream_t stream);
/* dq, dq1, dq2: per-layer gradients wrt qpos, summed on load (dq1/dq2 nullable). */
int msr3d_qpos_bwd_f32(const float *dq, const float *dq1, const float *dq2, const float *sz, int M,
                       const float *gamma_l,
                       const float *gamma_s, const float *xhat_l, const float *rstd_l,
                       const float *xhat_s, const float *rstd_s, float *dh1, float *dgamma_l,
                       float *dbeta_l, float *dgamma_s, float *dbeta_s, float *dWs, float *dbs,
                       float *dbl, msr3d_stream_t stream);
/* Attention core with `cond` fusion (transformers.py:205-248).  qkvc (B*L, ldq) rows
 * [q | k | v | cond(H x 6)]; key_pad (B,L) bytes, 1 = padded key (nullable).
 * Outputs P, LOC (B,H,L,L) and ctx (B*L, ldc). */
int msr3d_attn_fwd_f32(const float *qkvc, int ldq, const float *pw, const unsigned char *key_pad, int B,
                       int L, int H, float *P, float *LOC, float *ctx, int ldc, msr3d_stream_t stream);
int msr3d_attn_bwd_f32(const float *dctx, int ldc, const float *qkvc, int ldq, const float *pw,
                       const float *P, const float *LOC, int B, int L, int H, float *dqkvc,
                       msr3d_stream_t stream);
/* The other fusions of MultiHeadAttentionSpatial (modules/layers/transformers.py:207-244).
 * fusion: 0 cond (= the two calls above), 1 mul, 2 bias (and ctx), 3 add.  The location term is always
 * z[l,t] = sw[0] + sw[1:6] . pw[l,t,:] with sw read from the last H*6 columns of the qkvc row of token l —
 * written by the QKV+cond GEMM for `cond`, by msr3d_spatial_weights_fwd_f32 otherwise.  `add` saves
 * softmax(S) in P and softmax(z) in LOC (the fused attention is their mean); `bias` / `ctx` leave LOC untouched. */
int msr3d_attn_fusion_fwd_f32(const float *qkvc, int ldq, const float *pw, const unsigned char *key_pad, int B,
                              int L, int H, int fusion, float *P, float *LOC, float *ctx, int ldc,
                              msr3d_stream_t stream);
int msr3d_attn_fusion_bwd_f32(const float *dctx, int ldc, const float *qkvc, int ldq, const float *pw,
                              const float *P, const float *LOC, int B, int L, int H, int fusion, float *dqkvc,
                              msr3d_stream_t stream);
/* pairwise_loc_fc (transformers.py:189-192) as spatial weights in the qkvc rows.  ctx_mode 0 (mul / bias / add):
 * W (spatial_n_head, 5), b (spatial_n_head), multihead = spatial_n_head == H.  ctx_mode 1 (ctx): W (H*32, 5),
 * b (H*32); sw = q_h . [b_h | W_h] / sqrt(32) (reads the q columns of the same row).  The backward reads
 * d(sw) from dqkvc, ACCUMULATES dW / db (deterministic) and, for ctx, adds the q gradient into dqkvc. */
int msr3d_spatial_weights_fwd_f32(const float *W, const float *b, int ctx_mode, int multihead, float *qkvc, int ldq,
                                  int M, int H, msr3d_stream_t stream);
int msr3d_spatial_weights_bwd_f32(const float *W, const float *b, int ctx_mode, int multihead, const float *qkvc,
                                  float *dqkvc, int ldq, int M, int H, float *dW, float *db,
                                  msr3d_stream_t stream);
/* Situation modes with an agent token in front of the objects: as_object, as_object_add_loc
 * (model/ose3d_situation.py:334-363, 385-396; csrc/modes.cu).  Tokens per scene L + 1, row 0 = the agent.
 *   anchor_inputs:  all_loc (B, L+1, 6) = [(anchor_locs, anchor_size); obj_locs], key_pad (B, L+1) = [0; !obj_masks]
 *   loc_layer_*:    qpos = LN(Linear_6(all_loc)) (loc_layers[0]); forward also stores qpos (+ extra_a + extra_b, both
 *                   nullable (256): object_orientation_feat / type_embedding[0] when no agent row exists) as the first
 *                   term of x0
 *   anchor_assemble:    x0[object row] += proj + object_orientation_feat + type_embedding[0];
 *                       x0[agent row]  += anchor_feat + orientation_encoder(fourier(anchor_orientation)) + type_embedding[1]
 *   anchor_disassemble: dproj (B*L, 256) = object rows of dx0; ACCUMULATES d anchor_feat, d type_embedding[1],
 *                       d orientation_encoder.{weight (256, 84), bias} from the agent rows (fixed order). */
int msr3d_anchor_inputs(const float *obj_locs, const unsigned char *obj_masks, const float *anchor_locs,
                        const float *anchor_size, int B, int L, float *all_loc, unsigned char *key_pad,
                        msr3d_stream_t stream);
int msr3d_loc_layer_fwd_f32(const float *all_loc, int M, const float *W6, const float *b6, const float *gamma,
                            const float *beta, float *xhat, float *rstd, float *qpos, float *x0,
                            const float *extra_a, const float *extra_b, msr3d_stream_t stream);
int msr3d_loc_layer_bwd_f32(const float *dq, const float *dq1, const float *dq2, const float *all_loc, int M,
                            const float *gamma, const float *xhat, const float *rstd, float *dgamma, float *dbeta,
                            float *dW6, float *db6, msr3d_stream_t stream);
int msr3d_anchor_assemble_f32(float *x0, const float *proj, const float *ori_feat, const float *type_w,
                              const float *anchor_feat, const float *W_ori, const float *b_ori,
                              const float *anchor_ori, int B, int L, msr3d_stream_t stream);
int msr3d_anchor_disassemble_f32(const float *dx0, const float *anchor_ori, int B, int L, float *dproj,
                                 float *d_anchor_feat, float *d_type1, float *dW_ori, float *db_ori,
                                 msr3d_stream_t stream);
/* as_embedding (ose3d_situation.py:385-393): s_b = loc_embedding_encoder(fourier(anchor_locs[b])) +
 * orientation_encoder(fourier(anchor_orientation[b])) added to the query position of every object of scene b.
 * Forward: qpos (B*L, 256) += s_b (plain) and x0 += s_b (red.add: the projection GEMM reduces into x0 concurrently);
 * saves xhat_sit (B, 256) / rstd_sit (B) of the LayerNorm; qpos = x0 = NULL with s_out (B, 256): only the vector
 * itself (as_cross_attention).  Backward: from the per-layer qpos gradients (dq1 / dq2
 * nullable) ACCUMULATES the gradients of loc_embedding_encoder.{0.weight (256, 63), 0.bias, 1.weight, 1.bias} and
 * orientation_encoder.{weight (256, 84), bias}. */
int msr3d_situation_embed_fwd_f32(const float *anchor_locs, const float *anchor_ori, int B, int L, const float *Wl,
                                  const float *bl, const float *gamma_l, const float *beta_l, const float *W_ori,
                                  const float *b_ori, float *xhat_sit, float *rstd_sit, float *qpos, float *x0,
                                  float *s_out, msr3d_stream_t stream);
int msr3d_situation_embed_bwd_f32(const float *dq, const float *dq1, const float *dq2, const float *anchor_locs,
                                  const float *anchor_ori, int B, int L, const float *gamma_l, const float *xhat_sit,
                                  const float *rstd_sit, float *dWl, float *dbl, float *dgamma_l, float *dbeta_l,
                                  float *dW_ori, float *db_ori, msr3d_stream_t stream);
/* as_cross_attention (ose3d_situation.py:411-418; CrossAttentionLayer, transformers.py:22-73, prenorm): every key /
 * value of a scene is the same situation vector s_b, so the attention output of head h is scale * v_h with
 * v = W_v s_b + b_v and scale[b, h, n] = mean over the keys of the attention-dropout mask (csrc/modes.cu explains why
 * W_q, W_k and norm1 receive exactly zero gradient).  Per layer:
 *   cross_vec_fwd          v (B, 256), P (B, 8, 256) = per-head slices of W_o v
 *   cross_residual_ln_fwd  xa = x + drop_2(b_o + sum_h scale P_h);  t3 = LN_3(xa)   (scale (M, 8), xhat3, rstd3 saved)
 *   [msr3d_ffn_pair_fwd_f32 on t3]
 *   cross_ffn_residual_fwd out = xa + drop_3(f)
 *   cross_ffn_residual_bwd df = drop_3 mask * d out;  d linear2.bias ACCUMULATED
 *   [msr3d_ffn_pair_bwd_f32, weight-gradient GEMMs]
 *   cross_ln_bwd           dxa = d out + LN_3 backward(dt3);  g_attn = drop_2 mask * dxa;  d norm3, d b_o ACCUMULATED
 *   cross_vec_bwd          d W_o, d W_v, d b_v ACCUMULATED; ds (B, 256) += W_v^T dv  (the situation vector's gradient,
 *                          summed over the layers and fed to msr3d_situation_embed_bwd_f32 with L = 1) */
int msr3d_cross_vec_fwd_f32(const float *s, const float *Wv, const float *bv, const float *Wo, int B, float *v, float *P,
                            msr3d_stream_t stream);
int msr3d_cross_residual_ln_fwd_f32(const float *x, const float *P, const float *bo, int M, int L, const float *gamma3,
                                    const float *beta3, unsigned long long seed, const unsigned long long *seed_ptr,
                                    float p, unsigned site_attn, unsigned site_2, float *scale, float *xa, float *xhat3,
                                    float *rstd3, float *t3, msr3d_stream_t stream);
int msr3d_cross_ffn_residual_fwd_f32(const float *xa, const float *f, int M, unsigned long long seed,
                                     const unsigned long long *seed_ptr, float p, unsigned site_3, float *out,
                                     msr3d_stream_t stream);
int msr3d_cross_ffn_residual_bwd_f32(const float *dx, int M, unsigned long long seed, const unsigned long long *seed_ptr,
                                     float p, unsigned site_3, float *df, float *db2, msr3d_stream_t stream);
int msr3d_cross_ln_bwd_f32(const float *dt3, const float *dx, int M, const float *gamma3, unsigned long long seed,
                           const unsigned long long *seed_ptr, float p, unsigned site_2, const float *xhat3,
                           const float *rstd3, float *dxa, float *g_attn, float *dgamma3, float *dbeta3, float *dbo,
                           msr3d_stream_t stream);
int msr3d_cross_vec_bwd_f32(const float *g_attn, const float *scale, int B, int L, const float *v, const float *s,
                            const float *Wv, const float *Wo, float *dWo, float *dWv, float *dbv, float *ds,
                            msr3d_stream_t stream);
/* Same contract as msr3d_attn_bwd_f32 on the tensor cores (csrc/attn_block.cu: dP, dq, dk, dv as tcgen05 kind::tf32
 * 3xTF32 products, one CTA per (head, scene)); requires L <= 64, 8 heads, ldq = 816. */
int msr3d_attn_bwd_tc_f32(const float *dctx, int ldc, const float *qkvc, int ldq, const float *pw, const float *P,
                          const float *LOC, int B, int L, int H, float *dqkvc, msr3d_stream_t stream);
/* big = x rounded to tf32 (low 13 mantissa bits zero), small = x - big: the two operand images of the "3xTF32"
 * tensor-core products.  Run once per step over the flat parameter buffer so that no GEMM CTA splits a weight
 * tile again.  n % 4 == 0, 16 B aligned buffers. */
int msr3d_split_tf32_f32(const float *src, float *big, float *small, long long n, msr3d_stream_t stream);
/* The attention block of one encoder layer in ONE launch (cluster of 8 CTAs per scene, CTA = head):
 * qkvc = x Wqkvc^T + b; P, LOC, ctx as msr3d_attn_fwd_f32; o = ctx Wo^T + bo;
 * y2 = LN_1(x + drop_1(LN_a(drop_fc(o) + x)))   — transformers.py:205-251, 324-325.
 * Weights are passed as the big / small images written by msr3d_split_tf32_f32 (Wqkvc = [w_qs; w_ks; w_vs;
 * lang_cond_fc] (816, 256), Wo (256, 256)).  Outputs: everything the backward pass reads (qkvc (B*L, 816),
 * P / LOC (B, 8, L, L), ctx, xhat_a, xhat_1, y2 (B*L, 256), rstd_a, rstd_1 (B*L)).  Requires
 * msr3d_attn_block_supported(L, heads, hidden) (L <= 64, 8 heads, hidden 256). */
int msr3d_attn_block_supported(int L, int H, int hidden);
int msr3d_attn_block_fwd_f32(const float *x, const float *wqkvc_big, const float *wqkvc_small, const float *bqkvc,
                             const float *wo_big, const float *wo_small, const float *bo, const float *pw,
                             const unsigned char *key_pad, int B, int L, const float *gamma_a, const float *beta_a,
                             const float *gamma_1, const float *beta_1, unsigned long long seed,
                             const unsigned long long *seed_ptr, float p, unsigned site_fc, unsigned site_1,
                             float *qkvc, float *P, float *LOC, float *ctx, float *xhat_a, float *rstd_a,
                             float *xhat_1, float *rstd_1, float *y2, msr3d_stream_t stream);
/* Same; additionally y2's big / small tf32 images (B*L, 256 each) for msr3d_ffn_pair_fwd_ps_f32. */
int msr3d_attn_block_fwd_ps_f32(const float *x, const float *wqkvc_big, const float *wqkvc_small, const float *bqkvc,
                                const float *wo_big, const float *wo_small, const float *bo, const float *pw,
                                const unsigned char *key_pad, int B, int L, const float *gamma_a, const float *beta_a,
                                const float *gamma_1, const float *beta_1, unsigned long long seed,
                                const unsigned long long *seed_ptr, float p, unsigned site_fc, unsigned site_1,
                                float *qkvc, float *P, float *LOC, float *ctx, float *xhat_a, float *rstd_a,
                                float *xhat_1, float *rstd_1, float *y2, float *y2_big, float *y2_small,
                                msr3d_stream_t stream);
/* y2 = LN_1(x + drop_1(LN_a(drop_fc(o) + x)))  — transformers.py:250-251 then :324-325. */
int msr3d_attn_tail_fwd_f32(const float *o, const float *x, int M, const float *gamma_a,
                            const float *beta_a, const float *gamma_1, const float *beta_1,
                            unsigned long long seed, const unsigned long long *seed_ptr, float p,
                            unsigned site_fc, unsigned site_1, float *xhat_a, float *rstd_a, float *xhat_1, float *rstd_1, float *y2,
                            msr3d_stream_t stream);
int msr3d_attn_tail_bwd_f32(const float *dy2, int M, const float *gamma_a, const float *gamma_1,
                            unsigned long long seed, const unsigned long long *seed_ptr, float p,
                            unsigned site_fc, unsigned site_1, const float *xhat_a, const float *rstd_a, const float *xhat_1,
                            const float *rstd_1, float *dx, float *d_o, float *dgamma_a, float *dbeta_a,
                            float *dgamma_1, float *dbeta_1, float *dbias_o, msr3d_stream_t stream);
/* y3 = LN_2(y2 + drop_2(f)); x_next = y3 + add  — transformers.py:326-328, ose3d_situation.py:409-410. */
int msr3d_ffn_tail_fwd_f32(const float *f, const float *y2, int M, const float *gamma_2,
                           const float *beta_2, unsigned long long seed,
                           const unsigned long long *seed_ptr, float p, unsigned site_2,
                           const float *add, float *xhat_2, float *rstd_2, float *y3, float *x_next,
                           msr3d_stream_t stream);
int msr3d_ffn_tail_bwd_f32(const float *dy3, int M, const float *gamma_2, unsigned long long seed,
                           const unsigned long long *seed_ptr, float p, unsigned site_2, const float *xhat_2, const float *rstd_2,
                           float *dy2, float *df, float *dgamma_2, float *dbeta_2, float *dbias_2,
                           msr3d_stream_t stream);
/* Same; additionally df's big / small tf32 images (M, 256 each) for msr3d_ffn_pair_bwd_ps_f32. */
int msr3d_ffn_tail_bwd_ps_f32(const float *dy3, int M, const float *gamma_2, unsigned long long seed,
                              const unsigned long long *seed_ptr, float p, unsigned site_2, const float *xhat_2,
                              const float *rstd_2, float *dy2, float *df, float *dgamma_2, float *dbeta_2, float *dbias_2,
                              float *df_big, float *df_small, msr3d_stream_t stream);
/* Test hook: the multiplicative dropout mask (0 or 1/(1-p)) applied at `site` to offsets [0,n). */
int msr3d_dropout_mask_f32(unsigned long long seed, const unsigned long long *seed_ptr, float p,
                           unsigned site, long long n, float *out, msr3d_stream_t stream);
/* dst_i[n] += sum_m src[m, n] for up to three destinations (dst1, dst2 nullable). */
int msr3d_colsum_add_f32(const float *src, long long ld, int M, int N, float *dst0, float *dst1,
                         float *dst2, msr3d_stream_t stream);

/* ------------------------------------------------------------------ group D
 * Global-norm clip + AdamW on one flat buffer (trainer/leo_trainer.py:192-195,
 * optim/build.py:7-17).  state: MSR3D_OPT_STATE_FLOATS device floats {sumsq, step, lr, beta1, beta2,
 * eps, weight_decay, max_norm, then scratch for the deterministic norm reduction}; ticket: one zeroed device uint32; seed_counter (nullable): device
 * uint64 incremented once per step (the dropout seed offset); extra_sumsq (nullable, ABI 3): one device float
 * added to the squared gradient norm before the clip coefficient is formed — the squared norm (after the same
 * grad_scale) of the trainable parameters the caller optimises elsewhere (llm_proj, LoRA ...), because the
 * reference clips over ALL model parameters at once (accelerator.clip_grad_norm_(self.model.parameters())).  Elements [decay_begin, decay_end)
 * decay (the flat layout is [late biases | late weights | early weights | early biases], msr3d_b200/dist.py). */
#define MSR3D_OPT_STATE_FLOATS (8 + 4 * 148)
int msr3d_adamw_clip_step_f32(float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                              long long n, long long decay_begin, long long decay_end, float grad_scale,
                              float *state, unsigned int *ticket, unsigned long long *seed_counter,
                              const float *extra_sumsq, msr3d_stream_t stream);
int msr3d_adamw_clip_step_zg_f32(float *params, float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                                 long long decay_begin, long long decay_end, float grad_scale, float *state,
                                 unsigned int *ticket, unsigned long long *seed_counter, const float *extra_sumsq,
                                 int zero_grads, msr3d_stream_t stream);
/* msr3d_adamw_clip_step_f32 that also writes the tf32 big / small operand images (msr3d_split_tf32_f32) of every
 * updated parameter: the next step's 3xTF32 kernels find their weight operands split. */
int msr3d_adamw_clip_step_split_f32(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                                    long long decay_begin, long long decay_end, float grad_scale, float *state,
                                    unsigned int *ticket, unsigned long long *seed_counter, const float *extra_sumsq,
                                    float *big, float *small, msr3d_stream_t stream);
/* Clip + AdamW when the squared gradient norm is already in state[0] (msr3d_grad_norm_finalize_f32): the update
 * kernel alone; big / small (nullable) as in msr3d_adamw_clip_step_split_f32. */
int msr3d_adamw_apply_f32(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                          long long decay_begin, long long decay_end, float grad_scale, float *state, float *big,
                          float *small, const int *gate, msr3d_stream_t stream);
/* Gradient accumulation inside a replayable step (configs/msr3d.yaml:33 gradient_accumulation_steps,
 * trainer/leo_trainer.py:66,211): state = int32 {micro, k, clear_now, sync_now} on the device.  msr3d_accum_tick with
 * phase 0 (start of a micro-step) sets clear_now = (micro == 0) and sync_now = (micro == k - 1); with phase 1 (end)
 * it advances micro modulo k and increments *seed_counter (fresh dropout masks per micro-step).  Kernels that take
 * `gate` (nullable: always run) return at once when *gate == 0: the gradient clear is gated by &state[2]; the
 * exchange, the norm and the update by &state[3]. */
int msr3d_accum_tick(int *state, unsigned long long *seed_counter, int phase, msr3d_stream_t stream);
int msr3d_zero_gated_f32(float *p, long long n, const int *gate, msr3d_stream_t stream);
int msr3d_adamw_clip_step_gated_f32(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                                    long long decay_begin, long long decay_end, float grad_scale, float *state,
                                    unsigned int *ticket, const float *extra_sumsq, float *big, float *small,
                                    const int *gate, msr3d_stream_t stream);
/* Data-parallel gradient exchange over NVLink peer memory (csrc/allreduce.cu; replaces the NCCL all-reduce behind DDP,
 * trainer/leo_trainer.py:50-52,135).  buf[q] / flags[q] / partials[q]: rank q's copy of three SYMMETRIC allocations as
 * mapped into this process (torch.distributed._symmetric_memory): the flat gradient buffer, 2 * MSR3D_AR_MAX_BLOCKS *
 * MSR3D_AR_MAX_WORLD * 2 zero-initialised uint32 flags, 2 * MSR3D_AR_MAX_WORLD * MSR3D_AR_MAX_BLOCKS floats.  epoch:
 * local zero-initialised uint32[2 * MSR3D_AR_MAX_BLOCKS].  channel 0 / 1: two buckets may be in flight at once.
 * msr3d_allreduce_sum_f32 sums elements [begin, end) of the gradient buffer over the ranks IN PLACE on every rank
 * (fixed summation order: bit-identical replicas) and leaves the squared norm of each rank's slice in `partials`;
 * every rank must launch it with the same range, channel and nblocks.  msr3d_grad_norm_finalize_f32 adds the partials
 * of both channels (nblocks1 = 0: one bucket) * grad_scale^2 (+ *extra_sumsq) into state[0], increments state[1]
 * (step) and *seed_counter: what the norm pass of msr3d_adamw_clip_step_f32 does. */
#define MSR3D_AR_MAX_WORLD 8
#define MSR3D_AR_MAX_BLOCKS 64
typedef struct msr3d_allreduce_desc {
  float *buf[MSR3D_AR_MAX_WORLD];
  unsigned *flags[MSR3D_AR_MAX_WORLD];
  float *partials[MSR3D_AR_MAX_WORLD];
  unsigned *epoch;
  const int *gate;          /* nullable; *gate == 0: the launch does nothing (gradient accumulation) */
  int rank, world, channel;
} msr3d_allreduce_desc;
int msr3d_allreduce_sum_f32(const msr3d_allreduce_desc *d, long long begin, long long end, int nblocks,
                            msr3d_stream_t stream);
int msr3d_grad_norm_finalize_f32(const float *partials, int world, int nblocks0, int nblocks1, float grad_scale,
                                 const float *extra_sumsq, float *state, unsigned long long *seed_counter,
                                 const int *gate, msr3d_stream_t stream);
/* Single-GPU counterpart of the norm partials msr3d_allreduce_sum_f32 pushes along: squared norm of g[begin, end)
 * (float4-aligned range) as `nblocks` block partials of rank 0, channel `channel`, in the layout
 * msr3d_grad_norm_finalize_f32(world = 1) sums.  The early gradient bucket can be done on a side stream under the
 * backward of layer 0. */
int msr3d_sumsq_partials_f32(const float *g, long long begin, long long end, int nblocks, float *partials, int channel,
                             msr3d_stream_t stream);
/* Input staging of a pipelined step in one launch (csrc/stage.cu): rotate k: a[i] <- b[i]; b[i] <- c[i] (c nullable)
 * over rot_n[k] floats; copy k: dst[i] <- src[i]; masks (bytes): the same rotation and key_pad[i] <- !a[i]. */
#define MSR3D_STAGE_MAX 8
typedef struct msr3d_stage_desc {
  float *rot_a[MSR3D_STAGE_MAX];
  float *rot_b[MSR3D_STAGE_MAX];
  const float *rot_c[MSR3D_STAGE_MAX];
  long long rot_n[MSR3D_STAGE_MAX];
  float *cp_dst[MSR3D_STAGE_MAX];
  const float *cp_src[MSR3D_STAGE_MAX];
  long long cp_n[MSR3D_STAGE_MAX];
  unsigned char *mask_a, *mask_b;
  const unsigned char *mask_c;
  unsigned char *key_pad;
  long long n_mask;
  int n_rotate, n_copy;
} msr3d_stage_desc;
int msr3d_stage_step(const msr3d_stage_desc *d, msr3d_stream_t stream);
/* out[0] = <a, b> over n floats (one CTA, fixed summation order): the scalar the training loop reads
 * back (trainer/leo_trainer.py:196-201 logs the loss every step). */
int msr3d_dot_f32(const float *a, const float *b, long long n, float *out, msr3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MSR3D_B200_H */
