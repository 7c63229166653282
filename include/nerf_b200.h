/*
 * nerf_b200.h — C ABI of libnerfb200.so, the B200 (sm_100a) ray-marching engine.
 *
 * Drop-in boundary for the render hot path of Enigmatisms/NeRF.  The reference has no FFI
 * layer: its "operator API" is a set of Python call signatures.  Each entry
 * point below names the reference function (file:line under /root/reference) it replaces; the
 * Python package `nerf_b200` binds them with ctypes and re-exposes the reference signatures.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - return 0 on success, <0 on error; nb2_last_error() returns a thread-local message.
 *   - every data pointer is a DEVICE pointer on the handle's device unless the name ends in
 *     `_host`; tensors are fp32, row-major, contiguous; indices are int64 (torch.long).
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and the call returns
 *     without synchronising.  Outputs are pre-allocated by the caller; the library never
 *     allocates or frees caller-visible memory (scratch comes from the caller-provided
 *     workspace, sized by nb2_render_workspace_bytes()).
 *   - one handle per device per process; a handle owns only its packed-weight buffers and per-device launch state.
 *     Every entry point switches to the handle's device for the duration of the call and restores the caller's
 *     current device, so handles of several GPUs can be driven from one thread.
 *   - `jitter` / `u` pointers may be NULL: the library then draws the uniforms on the device
 *     with Philox4x32-10 keyed by (seed, ray_offset + ray index, sample index), so results do
 *     not depend on how rays are sharded across launches or GPUs.
 */
#ifndef NERF_B200_H_
#define NERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB2_VERSION 200 /* 0.2.0 */

typedef struct nb2_handle nb2_handle;

/* Network kinds (layer tables are fixed by the reference architectures).  A handle holds up to 64 packed networks
 * ("slots"); slot ids 0 and 1 always exist and have kinds NB2_NET_PROPOSAL and NB2_NET_NERF, further slots come from
 * nb2_net_create (several module instances -- train / eval / EMA copies -- each own their packed image).  Every
 * `net_id` argument below is a slot id. */
enum {
  NB2_NET_PROPOSAL = 0, /* nerf/addtional.py:53-72   ProposalNetwork(10, 256): 63-256-256-256-256-1 */
  NB2_NET_NERF = 1      /* nerf/mip_model.py:14-38   MipNeRF(10, 4, 256): 8x256 trunk + heads       */
};

/* Arithmetic of the MLP contraction. */
enum {
  NB2_PREC_FP32 = 0,   /* fp32 FFMA on CUDA cores (strict mode; reference arithmetic)                 */
  NB2_PREC_FP16X3 = 1, /* tcgen05, operands split hi+lo fp16 (22-bit), 3 MMAs, fp32 accumulate:
                          products exact to ~2^-22 -> fp32-faithful                                   */
  NB2_PREC_BF16 = 2,   /* tcgen05, single bf16 pass, fp32 accumulate                                  */
  NB2_PREC_FP16 = 3,   /* tcgen05, single fp16 pass (the reference's own AMP dtype), fp32 accumulate  */
  NB2_PREC_BF16X3 = 4, /* tcgen05, hi+lo bf16 split (16-bit), 3 MMAs                                  */
  /* nb2_render_rays only: the proposal network (17 % of the FLOPs, but it places the fine samples, and the resampling
     step amplifies its density errors) in the split precision, the NeRF network in a single pass */
  NB2_PREC_FP16_MIXED = 5, /* proposal NB2_PREC_FP16X3, NeRF NB2_PREC_FP16 */
  NB2_PREC_BF16_MIXED = 6  /* proposal NB2_PREC_BF16X3, NeRF NB2_PREC_BF16 */
};

/* Flags for nb2_composite / nb2_render_rays. */
enum {
  NB2_WHITE_BKG = 1,      /* rgb += 1 - sum(w)                    nerf/nerf_base.py:103-105 */
  NB2_DENSITY_SOFTPLUS = 2, /* proposal density softplus'd (train.py:169) instead of raw     */
  NB2_PROPOSAL_IPE = 4    /* nb2_render_rays: the proposal network reads the INTEGRATED positional encoding of the conical
                             frustum [z_s, z_{s+1}) of every coarse sample (z_P = z_{P-1} + (far - near) / (P - 1)) and the
                             frustum's mean as its position: ipe_feature -> ProposalNetwork.forward(mu, encoded_pt)
                             (nerf/mip_methods.py:47-58, nerf/addtional.py:88-91), evaluated in the kernel's producer.  The
                             reference never connects the two; depths, weights and resampling are unchanged.  Keeps the
                             reference's batch-global ||d|| (mip_methods.py:31) over the rays of the call. */
};

/* Error codes. */
enum {
  NB2_OK = 0,
  NB2_ERR_INVALID = -1, /* bad argument                 */
  NB2_ERR_CUDA = -2,    /* CUDA runtime error           */
  NB2_ERR_STATE = -3,   /* e.g. weights not packed      */
  NB2_ERR_UNSUPPORTED = -4
};

const char* nb2_last_error(void);
int nb2_version(void);

/* ---- lifetime ------------------------------------------------------------------------- */
int nb2_create(nb2_handle** out, int device);
int nb2_destroy(nb2_handle* h);

/*
 * Pack one network's fp32 parameters into the layouts the kernels consume (fp32 K-major
 * transposes for NB2_PREC_FP32; pre-swizzled 128x64 bf16 hi/lo operand tiles for tcgen05).
 * `W_ptrs[i]` / `b_ptrs[i]` are HOST arrays of DEVICE pointers to nn.Linear weight (out,in)
 * and bias (out), in reference state_dict order:
 *   NB2_NET_PROPOSAL (5): layers.{0,2,4,6,8}                      nerf/addtional.py:67-71
 *   NB2_NET_NERF    (11): lin_block1.{0,2,4,6}, lin_block2.{0,2,4}, bottle_neck.0,
 *                         opacity_head.0, rgb_layer.{0,2}         nerf/mip_model.py:19-37
 * Must be re-run after every parameter update (the handle keeps a version stamp).  Setup call:
 * allocates the handle's packed buffers on first use and synchronises `stream` once.
 */
int nb2_pack_weights(nb2_handle* h, int net_id, const float* const* W_ptrs_host,
                     const float* const* b_ptrs_host, int n_layers, int pos_levels,
                     int dir_levels, int hidden, void* stream);
int nb2_weights_version(nb2_handle* h, int net_id);
/* Allocate / release an additional packed-network slot of `kind`; *net_id_out >= 2. */
int nb2_net_create(nb2_handle* h, int kind, int* net_id_out);
int nb2_net_destroy(nb2_handle* h, int net_id);

/* ---- a1: pixel grid -> camera rays           nerf/procedures.py:43-51,64 ----------------
 * pose: 12 floats (3x4 row-major, device).  rays_out: (n_rays, 6) = [origin, R*(cx/fx, cy/fy, -1)],
 * raster order; ray i of the launch is pixel (pix_offset + i) of the H x W image. */
int nb2_generate_rays(nb2_handle* h, const float* pose, int H, int W, float focal_x,
                      float focal_y, int64_t pix_offset, int64_t n_rays, float* rays_out,
                      void* stream);

/* ---- a3+a4: stratified coarse depths and points  nerf/procedures.py:52,59,65-66 ---------
 * z[r,s] = base_z[s] + jitter[r,s] * resolution ; pts = o + z*d.  pts_out may be NULL. */
int nb2_sample_coarse(nb2_handle* h, const float* rays, const float* base_z,
                      const float* jitter, float resolution, uint64_t seed,
                      int64_t ray_offset, int64_t n_rays, int n_samples, float* z_out,
                      float* pts_out, void* stream);

/* ---- a5: sinusoidal positional encoding       nerf/nerf_helper.py:38-48 -----------------
 * x: (n, 3) -> out: (n, 6*levels), per level [sin(2^l x)(3), cos(2^l x)(3)].  `dims` = 3. */
int nb2_posenc(nb2_handle* h, const float* x, int64_t n, int dims, int levels, float* out,
               void* stream);

/* ---- a14: integrated positional encoding      nerf/mip_methods.py:15-58 -----------------
 * zvals (R, C+1), rays (R, 6) -> feat (R, C, 6L), mu (R, C, 3), mu_t (R, C).
 * Keeps the reference's batch-global ||d||_F (mip_methods.py:31); `scratch` >= 8 bytes. */
int nb2_ipe(nb2_handle* h, const float* zvals, const float* rays, int64_t n_rays, int n_cones,
            int levels, float radius, float* feat_out, float* mu_out, float* mu_t_out,
            void* scratch, void* stream);

/* ---- a7 (+ weights half of a12): density -> alpha-compositing weights --------------------
 * nerf/addtional.py:99-107, nerf/nerf_base.py:79-86.  dirs may be NULL (no ||d|| scaling).
 * act: 0 = relu, 1 = softplus, 2 = identity. */
int nb2_weights_from_sigma(nb2_handle* h, const float* sigma, const float* z,
                           const float* dirs, int dir_stride, int64_t n_rays, int n_samples,
                           int act, float* weights_out, void* stream);

/* ---- a8: 2-tap max + 2-tap blur               nerf/mip_methods.py:61-66 ----------------- */
int nb2_max_blur(nb2_handle* h, const float* weights, int64_t n_rays, int n_samples,
                 float alpha, float* out, void* stream);

/* ---- a9: inverse-CDF sampling                 nerf/utils.py:34-44,108-133 ----------------
 * nb2_sample_pdf == sample_pdf: bins (R, B), weights (R, B-1), u (R, N) or NULL ->
 *   samples (R, N) fp32, below (R, N) int64, above (R, N) int64 (unsorted).
 * nb2_inverse_sample == inverseSample(weights (R,P), z (R,P), N, sort): bins = mid(z),
 *   weights[1:-1]; if sort != 0 samples ascend and `below` is gathered by the sort permutation. */
int nb2_sample_pdf(nb2_handle* h, const float* bins, const float* weights, const float* u,
                   uint64_t seed, int64_t ray_offset, int64_t n_rays, int n_bins,
                   int n_draw, float* samples_out, int64_t* below_out, int64_t* above_out,
                   void* stream);
int nb2_inverse_sample(nb2_handle* h, const float* weights, const float* z, const float* u,
                       uint64_t seed, int64_t ray_offset, int64_t n_rays, int n_samples,
                       int n_draw, int sort, float* samples_out, int64_t* below_out,
                       void* stream);
/* Search stage only: identical (cdf, u) -> indices bit-exact (SURVEY.md §7 hard part 1a). */
int nb2_search_cdf(nb2_handle* h, const float* cdf, const float* u, int64_t n_rays,
                   int n_cdf, int n_draw, int64_t* inds_out, void* stream);

/* ---- a7+a8+a9+drop-last fused (render path)   nerf/procedures.py:68-70,76 ----------------
 * sigma (R,P) raw proposal density, z (R,P), rays (R,6) -> z_fine (R, n_draw-1) ascending;
 * below_out (R, n_draw-1) int64 or NULL: bin index of every kept sample (costs the stable rank sort). */
int nb2_resample(nb2_handle* h, const float* sigma, const float* z, const float* rays,
                 const float* u, uint64_t seed, int64_t ray_offset, int64_t n_rays,
                 int n_samples, int n_draw, float blur_alpha, int flags, float* z_fine_out,
                 int64_t* below_out, void* stream);

/* ---- a10: depths -> sample points             nerf/nerf_base.py:52-56 -------------------
 * rays (R,6), z (R,P) -> (R,P,6) = [o + z*d, d]. */
int nb2_length2pts(nb2_handle* h, const float* rays, const float* z, int64_t n_rays,
                   int n_samples, float* pts_out, void* stream);

/* ---- a13: coarse/fine merge (Ref-NeRF branch) nerf/nerf_base.py:58-73 -------------------
 * cat(f_z (R,F), c_z (R,C)) -> sort -> drop last -> z (R,F+C-1), pts (R,F+C-1,6). */
int nb2_coarse_fine_merge(nb2_handle* h, const float* rays, const float* c_z, const float* f_z,
                          int64_t n_rays, int n_coarse, int n_fine, float* z_out,
                          float* pts_out, void* stream);
/* The same with the index bookkeeping of the training path (nerf_base.py:62-65,70-71): f_inds (R,F) int64 ->
 * all_inds (R, F+C) = gather(cat(f_inds, arange(C)), sort order), sort_inds (R, F+C-1) = the sort permutation, last dropped. */
int nb2_coarse_fine_merge_inds(nb2_handle* h, const float* rays, const float* c_z, const float* f_z, const int64_t* f_inds,
                               int64_t n_rays, int n_coarse, int n_fine, float* z_out, float* pts_out, int64_t* all_inds_out,
                               int64_t* sort_inds_out, void* stream);

/* ---- f1 (training-side callers of the path, forward only) -----------------------------------
 * validSampler  nerf/utils.py:72-94: ray r takes pixel indices[r] (NULL -> device Philox) of the flattened image:
 *   rgbs (N,3), coords (N,2) int64 = (col - W/2, H/2 - row); cam_tf 12 floats; base_z (P) = linspace(near, far-res, P);
 *   -> pts (R,P,3), lengths (R,P), rgb (R,3), rays (R,6).
 * getBounds     nerf/addtional.py:14-18: weights (R,P), inds (R,K) int64 -> (R,K-1) summed-area-table differences. */
int nb2_valid_sampler(nb2_handle* h, const float* rgbs, const int64_t* coords, const float* cam_tf,
                      const int64_t* indices, const float* base_z, const float* jitter, float focal_x, float focal_y,
                      float resolution, uint64_t seed, int64_t ray_offset, int64_t n_pixels, int64_t n_rays,
                      int n_samples, float* pts_out, float* lengths_out, float* rgb_out, float* rays_out, void* stream);
int nb2_get_bounds(nb2_handle* h, const float* weights, const int64_t* inds, int64_t n_rays, int n_samples,
                   int n_inds, float* out, void* stream);

/* ---- a6 / a11: MLP forward ---------------------------------------------------------------
 * NB2_NET_PROPOSAL: pts (R*P, 3) -> out (R*P) raw density.  ProposalNetwork.forward
 *                   nerf/addtional.py:88-96.  `pts_stride` = 3.
 * NB2_NET_NERF:     pts (R*P, 6) = [xyz, dir] -> out (R*P, 4) = [sigmoid rgb, raw sigma].
 *                   MipNeRF.forward nerf/mip_model.py:41-60.  `pts_stride` = 6. */
int nb2_mlp_forward(nb2_handle* h, int net_id, int precision, const float* pts,
                    int pts_stride, int64_t n_points, float* out, void* stream);

/* ProposalNetwork.forward(pts, encoded_pt): `encoded` (n, 6 * pos_levels) replaces the sinusoidal columns of the
 * network input (the hook through which integrated positional encoding enters; nerf/addtional.py:88-91). */
int nb2_mlp_forward_encoded(nb2_handle* h, int net_id, int precision, const float* pts, int pts_stride,
                            const float* encoded, int64_t n_points, float* out, void* stream);

/* ---- a12: alpha compositing                   nerf/nerf_base.py:90-113 ------------------
 * rgbo (R,P,4), z (R,P), dirs (R, dir_stride>=3) -> rgb (R,3), weights (R,P) or NULL,
 * depth (R) or NULL = (sum w*z*||d|| - near) / (far - near), acc (R) or NULL. */
int nb2_composite(nb2_handle* h, const float* rgbo, const float* z, const float* dirs,
                  int dir_stride, int64_t n_rays, int n_samples, int flags, float near_t,
                  float far_t, float* rgb_out, float* weights_out, float* depth_out,
                  float* acc_out, void* stream);
/* The same with one auxiliary per-sample channel: aux (R,P) -> aux_out (R) = sum_i w_i aux_i (the normal image of the
 * Ref-NeRF branch, nerf/nerf_base.py:110-112: aux = normal . cam_dir; the caller applies (x + 1) / 2). */
int nb2_composite_aux(nb2_handle* h, const float* rgbo, const float* z, const float* dirs, int dir_stride, int64_t n_rays,
                      int n_samples, int flags, float near_t, float far_t, const float* aux, float* rgb_out, float* weights_out,
                      float* depth_out, float* acc_out, float* aux_out, void* stream);

/* ---- the fused path: render_image's per-ray work   nerf/procedures.py:64-85 --------------
 * rays (R,6) -> rgb (R,3), depth (R) or NULL, acc (R) or NULL.  Three launches: fused
 * sample+encode+proposal-MLP, resample, fused encode+NeRF-MLP+composite.
 * Optional debug outputs (NULL to skip): z_coarse (R,Pc), sigma_prop (R,Pc), z_fine (R,Pf) and below_fine (R,Pf)
 * int64 = the cdf bin index of every kept fine sample (inverseSample's second result, nerf/utils.py:41-43). */
typedef struct nb2_render_params {
  int n_coarse;       /* 64   RENDER_COARSE_PNUM, nerf/procedures.py:22 */
  int n_fine;         /* 128  sample_num; n_fine+1 are drawn, the largest dropped */
  float near_t, far_t;
  float resolution;   /* (far-near)/sample_num, nerf/procedures.py:59 */
  float blur_alpha;   /* 0.01, nerf/procedures.py:69 */
  int flags;          /* NB2_WHITE_BKG | NB2_DENSITY_SOFTPLUS */
  int precision;      /* NB2_PREC_* */
  uint64_t seed;      /* Philox key when jitter/u are NULL */
  int64_t ray_offset; /* global index of rays[0] (shard invariance) */
  int prop_net_id;    /* packed-network slots to render with; 0 / 0 (a zeroed struct) selects the default slots 0 / 1 */
  int nerf_net_id;
  /* Multi-GPU gather fused into the compositing epilogue: when n_peers > 0 the rgb row of ray r is also
   * stored at peer_rgb[q] + 3 * (ray_offset + r) for q < n_peers <= 8 -- device pointers to the full (n_total_rays, 3)
   * image buffers of the other GPUs, mapped with nb2_ipc_open; the stores travel over NVLink and no gather kernel or
   * collective follows (callers fence with a barrier before reading the image). */
  int n_peers;
  float ipe_radius;   /* NB2_PROPOSAL_IPE: pixel-footprint radius r of coneParameters (mip_methods.py:15-23) */
  float* peer_rgb[8];
} nb2_render_params;

int64_t nb2_render_workspace_bytes(int64_t n_rays, const nb2_render_params* p);
int nb2_render_rays(nb2_handle* h, const nb2_render_params* p, const float* rays,
                    const float* base_z, const float* jitter, const float* u, int64_t n_rays,
                    float* rgb_out, float* depth_out, float* acc_out, float* z_coarse_out,
                    float* sigma_prop_out, float* z_fine_out, int64_t* below_fine_out, void* workspace,
                    int64_t workspace_bytes, void* stream);

/* ---- the layer-wise engine: training step and Ref-NeRF (8f-3) ----------------------------------------
 * One generic tensor-core GEMM on bf16 operands with fp32 accumulation:
 *     D[M, N] = epilogue( sum over segments s of  A_s[M, K_s] * B_s[N, K_s]^T )
 * Operands stay in their natural row-major layouts; `mn_major` = 0: element (i, k) at ptr[i * ld + k] (K contiguous);
 * `mn_major` = 1: element (i, k) at ptr[k * ld + i] (the M / N index contiguous: a transposed view, free on tcgen05).
 *   nn.Linear forward   Y  = X W^T     A = X (0), B = W (0)            reference: every F.linear of nerf/mip_model.py:53-59,
 *   its dgrad           dX = dY W      A = dY (0), B = W (1)                      nerf/addtional.py:92-96, nerf/ref_model.py:76-105
 *   its wgrad           dW = dY^T X    A = dY (1), B = X (1), splits > 1          (autograd's addmm backward in train.py:206)
 *   its bias gradient   db = dY^T 1    the same with B = a column of ones
 * Segments concatenate along K: the passes of the split precision (x = hi + lo: lo*hi, hi*lo, hi*hi) and torch.cat inputs.
 * Epilogue: + bias[n]; act 0 none | 1 relu | 2 sigmoid; * (mask[m][n] > 0) (relu backward on a saved activation);
 * outputs fp32 (ld_f32) and / or bf16 hi (+ lo residual) (ld_16).  splits > 1: split-K over M-tile x N-tile x split work
 * items, partial sums written to out_f32 + split * split_stride (reduce with nb2_reduce_splits: deterministic, no atomics).
 * All ld multiples of 8 elements, pointers 16-byte aligned, contiguous dimensions zero-padded to multiples of 8. */
#define NB2_GEMM_MAX_SEG 8
typedef struct nb2_gemm_operand {
  const void* ptr; /* bf16 */
  int64_t ld;
  int mn_major;
  int reserved;
} nb2_gemm_operand;
typedef struct nb2_gemm_desc {
  int64_t M;
  int N;
  int n_seg;
  struct {
    nb2_gemm_operand a, b;
    int64_t K;
  } seg[NB2_GEMM_MAX_SEG];
  const float* bias;
  int act;
  int ld_mask;
  const void* mask; /* bf16 (M, ld_mask) */
  float* out_f32;
  void* out_hi;     /* bf16 */
  void* out_lo;     /* bf16 */
  int64_t ld_f32, ld_16;
  int splits;
  int reserved;
  int64_t split_stride; /* floats */
  /* optional, weight-gradient shape only (A MN-major, split-precision triples, one work item per CTA): the row sums of A
   * over K, sum_k A[m][k] = the bias gradient db = dY^T 1 (autograd's sum over the batch for nn.Linear's bias), from one
   * extra N = 16 MMA per k-step against a constant tile of ones: a_rowsum_out[split * a_rowsum_stride + m], fp32 partial
   * sums per split (reduce like out_f32).  NULL: not computed. */
  float* a_rowsum_out;
  int64_t a_rowsum_stride; /* floats */
} nb2_gemm_desc;
int nb2_gemm_bf16(nb2_handle* h, const nb2_gemm_desc* d, void* stream);
/* fp32 (rows, cols; row stride ld_src) -> bf16 hi (+ lo residual or NULL), row stride ld_dst (multiple of 8, pad zeroed);
 * col_perm (device, cols ints, or NULL): destination column of every source column. */
int nb2_to_bf16(nb2_handle* h, const float* src, int64_t rows, int cols, int64_t ld_src, const int* col_perm, void* hi, void* lo,
                int ld_dst, void* stream);
/* The same for an array of matrices in one launch per 32 descriptors (the weight images of a network after an optimizer step). */
typedef struct nb2_to_bf16_desc {
  const float* src;
  int64_t rows, ld_src;
  int cols, ld_dst;
  const int* col_perm;
  void* hi;
  void* lo;
} nb2_to_bf16_desc;
int nb2_to_bf16_batch(nb2_handle* h, const nb2_to_bf16_desc* descs, int n, void* stream);
/* out[m][c] (=|+=) sum_s ws[s * split_stride + m * ld_ws + perm(c)]: second stage of the split-K wgrad. */
int nb2_reduce_splits(nb2_handle* h, const float* ws, int splits, int64_t split_stride, int rows, int cols, int ld_ws,
                      const int* col_perm, float* out, int ld_out, int accumulate, void* stream);
/* Launch plans: the same two calls over arrays of descriptors, in order, on one stream.  The training step records its
 * GEMMs once per (network, batch size) and replays them with one host call per group (at the reference's 1024-ray batch,
 * train.py:236, the per-launch host cost is what bounds the step).  Stops at, and returns, the first error.
 * nb2_reduce_splits_batch runs up to 32 reductions per launch (they must write disjoint outputs; accumulate = 0). */
typedef struct nb2_reduce_desc {
  const float* ws;
  int splits, rows, cols, ld_ws, ld_out, accumulate;
  int64_t split_stride; /* floats */
  const int* col_perm;
  float* out;
} nb2_reduce_desc;
int nb2_gemm_bf16_batch(nb2_handle* h, const nb2_gemm_desc* descs, int n, void* stream);
int nb2_reduce_splits_batch(nb2_handle* h, const nb2_reduce_desc* descs, int n, void* stream);

/* ---- training step, HBM-bound pieces: encodings as GEMM operands and the backward of the ray ops --------
 * In the reference these are autograd's derivatives of the torch ops the functions are written in (train.py:206).
 * nb2_encode_bf16: x (n, x_stride) fp32, point / direction in columns [x_col0, x_col0+3) (normalize != 0: d / ||d|| first,
 *   nerf/mip_model.py:44-45) -> rows [x, sin(2^l x), cos(2^l x) ...] (nerf/nerf_helper.py:38-48) as bf16 hi (+ lo) with
 *   row stride ld, zero-padded to `width` columns.
 * nb2_weights_from_sigma_backward: g_weights (R,P) -> d_sigma (R,P)          nerf/addtional.py:99-107, nerf_base.py:79-86
 * nb2_composite_backward: g_rgb (R,3), g_weights (R,P) or NULL -> d_rgbo (R,P,4)             nerf/nerf_base.py:90-113
 * nb2_max_blur_backward: weights (R,P), g_out (R,P) -> d_weights (R,P)                        nerf/mip_methods.py:61-66
 * nb2_get_bounds_backward: inds (R,K) int64, g_out (R,K-1) -> d_weights (R,P)                 nerf/addtional.py:14-18
 * nb2_nerf_head_backward: out (n,4) = MipNeRF.forward's result, g_out (n,4) -> d_z (n,8) bf16 hi/lo = g_rgb * rgb (1 - rgb)
 *   (rgb_layer.2 pre-activation gradient, columns 0..2) and d_s (n,8) bf16 hi/lo = g_sigma (column 0).  nerf/mip_model.py:57-60 */
int nb2_encode_bf16(nb2_handle* h, const float* x, int x_stride, int x_col0, int64_t n, int levels, int normalize, void* hi, void* lo,
                    int64_t ld, int width, void* stream);
int nb2_weights_from_sigma_backward(nb2_handle* h, const float* sigma, const float* z, const float* dirs, int dir_stride,
                                    int64_t n_rays, int n_samples, int act, const float* g_weights, float* d_sigma, void* stream);
int nb2_composite_backward(nb2_handle* h, const float* rgbo, const float* z, const float* dirs, int dir_stride, int64_t n_rays,
                           int n_samples, int flags, const float* g_rgb, const float* g_weights, float* d_rgbo, void* stream);
int nb2_max_blur_backward(nb2_handle* h, const float* weights, const float* g_out, int64_t n_rays, int n_samples, float* d_weights,
                          void* stream);
int nb2_get_bounds_backward(nb2_handle* h, const int64_t* inds, const float* g_out, int64_t n_rays, int n_samples, int n_inds,
                            float* d_weights, void* stream);
int nb2_nerf_head_backward(nb2_handle* h, const float* out, const float* g_out, int64_t n, void* dz_hi, void* dz_lo, void* ds_hi,
                           void* ds_lo, void* stream);

/* ---- Ref-NeRF forward glue (the GEMMs run on nb2_gemm_bf16, the directional encoding on nb2_ide) -----------
 * heads (n, ld_heads) fp32 = [normal(3), diffuse(3), tint(3), rho(1), density(1)]: norm_col_tint_head | rho_tau_head outputs.
 * nb2_ref_geometry: normal = -n / (||n|| + 1e-7), reflect = d - 2 (d . normal) normal, roughness = softplus(rho - 1),
 *   nv_dot = normal . d                                                                     nerf/ref_model.py:83-94
 * nb2_ref_dir_inputs: [ide (ide_width) | nv_dot | 0 ...] as bf16 hi / lo columns of the directional MLP input  nerf/ref_model.py:96
 * nb2_ref_color: rgb = [linear_to_srgb](spec * sigmoid(tint) + sigmoid(diffuse [- ln 3])), out (n,4) = [rgb, density]
 *   (density -> softplus(density + 0.5) when shift_softplus, nerf/procedures.py:74); ndot_out = normal . cam_dir (nerf_base.py:111)
 *                                                                                           nerf/ref_model.py:102-109 */
int nb2_ref_geometry(nb2_handle* h, const float* heads, int ld_heads, const float* dirs, int dir_stride, int64_t n, float* normal_out,
                     float* reflect_out, float* rough_out, float* nv_out, void* stream);
int nb2_ref_dir_inputs(nb2_handle* h, const float* ide, int ide_width, const float* nv_dot, int64_t n, void* hi, void* lo, int64_t ld,
                       int width, void* stream);
int nb2_ref_color(nb2_handle* h, const float* spec, const float* heads, int ld_heads, int use_srgb, int shift_softplus, const float* normal,
                  const float* cam_dir, int64_t n, float* out, float* ndot_out, void* stream);
/* out[i] = a[i] . b for a (n,3), b (3): `normal @ cam_dir` of NeRF.render's normal image (nerf/nerf_base.py:111). */
int nb2_dot3(nb2_handle* h, const float* a, const float* b, int64_t n, float* out, void* stream);

/* ---- Ref-NeRF training glue: backward of the three steps above (autograd's derivatives of nerf/ref_model.py:81-105 in the
 * reference's train.py:164-199 with is_ref_model; RefNeRF.get_grad, ref_model.py:118-124, is the d_x output below).
 * nb2_ref_color_backward: spec (n,3) = the sigmoid-ed specular head, g_out (n,4) = gradient of [rgb, density] ->
 *   ds (n,8) bf16 hi / lo: columns 0..2 = gradient of the specular head's pre-activation (GEMM operand), rest zero;
 *   d_heads (n, ld_dheads >= 11) fp32: columns 3..5 diffuse, 6..8 tint, 10 density, everything else zeroed.
 * nb2_ref_geometry_backward: d_in (n, ld_in) fp32 = gradient of the directional MLP's input, IDE columns at
 *   [ide_col0, ide_col0 + 2 n_pairs), nv_dot next; g_normal (n,3) = gradient of the returned normal or NULL; mat / ml as
 *   nb2_ide -> d_heads columns 0..2 (raw normal) and 9 (rho), through reflect / IDE / softplus / the normalisation.
 * nb2_encode_backward: d_enc (n, ld) fp32 = gradient of the rows nb2_encode_bf16 writes (normalize = 0) -> d_x (n,3). */
int nb2_ref_color_backward(nb2_handle* h, const float* spec, const float* heads, int ld_heads, int use_srgb, const float* g_out,
                           int64_t n, void* ds_hi, void* ds_lo, float* d_heads, int ld_dheads, void* stream);
int nb2_ref_geometry_backward(nb2_handle* h, const float* heads, int ld_heads, const float* dirs, int dir_stride, int64_t n,
                              const float* d_in, int64_t ld_in, int ide_col0, const float* g_normal, const float* mat, const int* ml,
                              int n_pairs, int n_pow, float* d_heads, int ld_dheads, void* stream);
int nb2_encode_backward(nb2_handle* h, const float* x, int x_stride, int x_col0, int64_t n, int levels, const float* d_enc, int64_t ld,
                        float* d_x, void* stream);

/* ---- peer memory for the fused gather (one process per GPU, CUDA IPC over NVLink / NVSwitch) -----------------
 * nb2_ipc_alloc: cudaMalloc `bytes` on the handle's device and export a 64-byte IPC handle for it.
 * nb2_ipc_open : map another process's allocation into this process (peer access is enabled on demand).
 * nb2_ipc_close / nb2_ipc_free: unmap / release. */
int nb2_ipc_alloc(nb2_handle* h, int64_t bytes, void** dev_ptr_out, void* ipc_handle64_out);
int nb2_ipc_open(nb2_handle* h, const void* ipc_handle64, void** dev_ptr_out);
int nb2_ipc_close(nb2_handle* h, void* dev_ptr);
int nb2_ipc_free(nb2_handle* h, void* dev_ptr);

/* Kernel-launch counter (all launches made through this handle since creation). */
int64_t nb2_launch_count(nb2_handle* h);

/* ---- next rows (Ref-NeRF forward helpers) ------------------------------------------------------------
 * Integrated directional encoding            nerf/ref_func.py:78-108 (closure built by generate_ide_fn, :51-76)
 * xyz: (n, 3) directions, kappa_inv: (n) roughness; mat: (n_pow, n_pairs) spherical-harmonic coefficient matrix and
 * ml: (2, n_pairs) int32 rows m then l, both built once on the host exactly as ref_func.py:38-76 does;
 * out: (n, 2 n_pairs) = [real parts, imaginary parts].  n_pairs <= 36, n_pow <= 17 (deg_view <= 5). */
int nb2_ide(nb2_handle* h, const float* xyz, const float* kappa_inv, int64_t n, const float* mat, const int* ml,
            int n_pairs, int n_pow, float* out, void* stream);

/* linear_to_srgb                              nerf/nerf_helper.py:50-56: elementwise over n floats. */
int nb2_linear_to_srgb(nb2_handle* h, const float* linear, int64_t n, float* out, void* stream);

/* Optional per-kernel timing of nb2_render_rays: four cudaEvent_t (created by the caller with timing
 * enabled) recorded on the render stream before launch 1 and after launches 1, 2, 3.  NULL disables. */
int nb2_set_profile_events(nb2_handle* h, void* const* events4);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H_ */
