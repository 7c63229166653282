// nb2_common.cuh — shared declarations for libnerfb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <unordered_map>

#include "../../include/nerf_b200.h"
#include "../../include/nerf_b200_debug.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "libnerfb200 is written for sm_100a only: compile with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace nb2 {

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define NB2_CHECK_ARG(cond, ...)                      \
  do {                                                \
    if (!(cond)) {                                    \
      nb2::set_error(__VA_ARGS__);                    \
      return NB2_ERR_INVALID;                         \
    }                                                 \
  } while (0)
#define NB2_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      nb2::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,   \
                     __LINE__);                                                          \
      return NB2_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)
#define NB2_LAUNCH_CHECK(h)                                                              \
  do {                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess) {                                                             \
      nb2::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),         \
                     __FILE__, __LINE__);                                                \
      return NB2_ERR_CUDA;                                                               \
    }                                                                                    \
    (h)->launches++;                                                                     \
  } while (0)

// ------------------------------------------------------------------------------------------
// network description shared by the pack step and the MLP kernels
// ------------------------------------------------------------------------------------------
constexpr int kHidden = 256;       // trunk width (reference default --nerf_net_width / --prop_net_width)
constexpr int kEncCols = 64;       // 3 + 6*10 = 63 encoded position columns, padded to 64
constexpr int kDirCols = 32;       // 3 + 6*4  = 27 encoded direction columns, padded to 32
constexpr int kRgbHidden = 128;    // rgb_layer.0 width (nerf/mip_model.py:35)
constexpr int kMaxPosLevels = 10;
constexpr int kMaxDirLevels = 4;

// Tensor-core operand tile: 128 rows x 64 bf16 columns, K-major, 128-byte swizzle (16 KB).
constexpr int kTileRows = 128;
constexpr int kTileCols = 64;
constexpr int kTileBytes = kTileRows * kTileCols * 2;  // 16384

// A-operand chunk ids inside one slot's activation buffer.
constexpr int kChunkH0 = 0;  // hidden activation columns [0,64) ... kChunkH0+3 = [192,256)
constexpr int kChunkE = 4;   // encoded position (later re-used for the encoded direction)
constexpr int kChunksPerSlot = 5;

enum EpiKind {
  EPI_RELU = 0,        // h = relu(acc + b)                       -> H
  EPI_RELU_SIGMA = 1,  // h = relu(acc + b); sigma = w_s.h + b_s  -> H (+ dir encoding into E)
  EPI_LINEAR = 2,      // h = acc + b                             -> H          (bottle_neck)
  EPI_RGB = 3,         // t = relu(acc + b); rgb = sigmoid(W1 t + b1)           (rgb_layer)
  EPI_SIGMA_OUT = 4    // h = relu(acc + b); out = w_s.h + b_s                  (proposal tail)
};

constexpr int kMaxTcLayers = 9;
struct TcLayer {
  int kc;        // number of 64-wide K chunks
  int nc;        // number of 128-wide N chunks
  int a_src[5];  // A chunk id for each K chunk
  int ks0[5];    // first 16-wide k-step of each K chunk (3 for a bias-only chunk, else 0)
  int epi;       // EpiKind
  int bias_off;  // offset (floats) into the fp32 bias array (CUDA-core path)
  int chunk0;    // first weight chunk of this layer in the packed stream
};
struct TcNet {
  int n_layers;
  int n_chunks;
  TcLayer layer[kMaxTcLayers];
};

// fp32 (CUDA-core) layer table: out = act(sum_seg X_seg . Wt_seg + b)
constexpr int kMaxSimtLayers = 9;
struct SimtLayer {
  int k0, k1;      // K of segment 0 (read from buffer src0) and segment 1 (src1); k1 = 0 if unused
  int src0, src1;  // 0 = E (encoded position), 1 = H, 2 = D (encoded direction)
  int n;           // output width (256 or 128)
  int epi;         // EpiKind
  int wt_off;      // offset (floats) of the transposed weight [(k0+k1) x n]
  int bias_off;
};
struct SimtNet {
  int n_layers;
  SimtLayer layer[kMaxSimtLayers];
};

struct PackedNet {
  bool in_use = false;   // slot allocated (slots 0 and 1 always are)
  int kind = 0;          // NB2_NET_PROPOSAL | NB2_NET_NERF
  bool packed = false;
  int version = 0;
  int pos_levels = 0, dir_levels = 0;
  TcNet tc;
  SimtNet simt;
  __nv_bfloat16* d_wchunks = nullptr;  // [n_chunks][4 (bf16 hi, lo, fp16 hi, lo)][8192] pre-swizzled tiles
  float* d_bias = nullptr;             // all MMA-layer biases, concatenated
  float* d_head = nullptr;             // sigma head: w[256], b ; rgb head: W1[3][128], b1[3]
  float* d_wt32 = nullptr;             // fp32 transposed weights for the CUDA-core path
};
constexpr int kHeadSigmaW = 0;      // 256 floats
constexpr int kHeadSigmaB = 256;    // 1 float
constexpr int kHeadRgbW = 260;      // 3 x 128 floats
constexpr int kHeadRgbB = 260 + 384;  // 3 floats
constexpr int kHeadFloats = 652;

}  // namespace nb2

namespace nb2 {
constexpr int kMaxNets = 64;   // packed-network slots per handle (0 / 1: the default proposal / NeRF slots)
constexpr int kMaxPeers = 8;   // GPUs of one node whose image buffers the fused render can write directly
// Per-kernel launch state.  cudaFuncSetAttribute and the co-resident cluster count are per DEVICE, so they are cached
// in the handle (one handle per device), never in function-local statics.
struct KernelCache {
  bool attr_set = false;
  int max_clusters = 0;
};
}  // namespace nb2

struct nb2_handle {
  int device = 0;
  int sm_count = 0;
  int64_t launches = 0;
  int tc_debug = 0;              // NB2_TC_DEBUG, read once in nb2_create (timing experiments: 1 = no weight waits, 2 = no MMAs)
  cudaEvent_t prof[4] = {nullptr, nullptr, nullptr, nullptr};
  long long* tc_prof = nullptr;  // debug: per-CTA role cycle counters of the last tensor-kernel launch
  nb2::PackedNet net[nb2::kMaxNets];
  std::unordered_map<const void*, nb2::KernelCache> kcache;
};

namespace nb2 {

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Philox4x32-10 (Salmon et al. 2011).  counter = (c0,c1,c2,c3), key = (k0,k1).
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c[0];
    uint64_t p1 = (uint64_t)M1 * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}
// Uniform in [0,1) with 24 random bits.  stream: 0 = coarse jitter, 1 = inverse-CDF u.
__host__ __device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t ray, uint32_t sample,
                                                         uint32_t stream) {
  uint32_t c[4] = {(uint32_t)ray, (uint32_t)(ray >> 32), sample >> 2, stream};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (float)(c[sample & 3] >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float softplus_f(float x) {
  // torch.nn.functional.softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float apply_density_act(float x, int act) {
  return act == 0 ? fmaxf(x, 0.f) : (act == 1 ? softplus_f(x) : x);
}

// ---- branch-free sin / cos for the fused encoders ------------------------------------------------------------------
// Three-constant Cody-Waite reduction by pi/2 and the Cephes single-precision minimax polynomials on [-pi/4, pi/4]:
// max |error| 9.4e-8 against fp64 for |x| <= 1e5 (checked on the host over 2e7 arguments; the largest argument of the fused
// encoders is 2^9 * |position| ~ 5e3), i.e. the accuracy and the range of sincosf's own fast path.  Unlike sincosf there is no large-argument branch, so the 30 (position)
// + 12 (direction) evaluations of a row are straight-line code the scheduler can interleave: the library call made the
// encoding ~350 cycles per evaluation with two warps per scheduler (profiles/r01_roles_*).
__device__ __forceinline__ void enc_sincos(float x, float& sn, float& cs) {
  const int q = __float2int_rn(x * 0.636619772f);
  const float j = __int2float_rn(q);
  float t = fmaf(-j, 1.57079601e+00f, x);
  t = fmaf(-j, 3.13916473e-07f, t);
  t = fmaf(-j, 5.39030253e-15f, t);
  const float s = t * t;
  float ps = fmaf(s, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, s, -1.6666654611e-1f);
  ps = fmaf(ps * s, t, t);
  float pc = fmaf(s, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, s, 4.166664568298827e-2f);
  pc = fmaf(pc * s, s, fmaf(s, -0.5f, 1.0f));
  const float rs = (q & 1) ? pc : ps, rc = (q & 1) ? ps : pc;
  sn = (q & 2) ? -rs : rs;
  cs = ((q + 1) & 2) ? -rc : rc;
}

// Same with the library's large-argument path behind a (normally uniform) branch, for the standalone encoders whose
// inputs are arbitrary.
__device__ __forceinline__ void sincos_any(float x, float& sn, float& cs) {
  if (fabsf(x) <= 1.0e5f) enc_sincos(x, sn, cs);
  else sincosf(x, &sn, &cs);
}

// Inclusive product scan over a warp.
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------
// per-handle launch helpers (nb2_api.cu)
// ------------------------------------------------------------------------------------------
// Every entry point runs on the handle's device whatever the caller's current device is, and restores it on exit.
struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};
#define NB2_ENTER(h)                               \
  NB2_CHECK_ARG((h) != nullptr, "null handle");    \
  nb2::DeviceGuard nb2_device_guard_((h)->device)
// opt-in dynamic shared memory above 48 KB: once per kernel per handle (= per device)
int kernel_set_smem(nb2_handle* h, const void* fn, int bytes);
// co-resident clusters of `fn` under `cfg` on the handle's device (cached)
int kernel_max_clusters(nb2_handle* h, const void* fn, const cudaLaunchConfig_t* cfg, int* out);
inline bool net_ok(const nb2_handle* h, int id) { return id >= 0 && id < kMaxNets && h->net[id].in_use; }

// launchers implemented in the other translation units --------------------------------------
struct MlpIo {
  // input selection
  int in_mode;            // 0 = explicit points, 1 = rays + z, 2 = rays + stratified sampling
  const float* pts;       // in_mode 0: (n_rows, pts_stride)
  const float* enc;       // in_mode 0, optional: (n_rows, 6 * pos_levels) externally encoded position features
                          // that replace the sin/cos columns (ProposalNetwork.forward(pts, encoded_pt), addtional.py:88-91)
  int pts_stride;
  const float* rays;      // in_mode 1/2: (n_rays, 6)
  const float* z;         // in_mode 1:   (n_rays, P)
  const float* base_z;    // in_mode 2:   (P)
  const float* jitter;    // in_mode 2:   (n_rays, P) or NULL -> Philox
  float resolution;
  uint64_t seed;
  int64_t ray_offset;
  // in_mode 2, optional (NB2_PROPOSAL_IPE): the network input is the integrated positional encoding of the conical frustum
  // [z_s, z_{s+1}) (z_P = z_{P-1} + ipe_last_step) instead of the point encoding at z_s        nerf/mip_methods.py:15-58
  int ipe;
  float ipe_radius, ipe_last_step;
  const double* ipe_sumsq;  // device: sum over the launch's rays of ||d||^2 (the reference's batch-global norm, mip_methods.py:31)
  int P;                  // samples per ray (in_mode 1/2)
  int64_t n_rows;         // total MLP rows (= n_rays * P)
  // output selection
  int out_mode;           // 0 = sigma (n_rows), 1 = rgbo (n_rows,4), 2 = composite per ray
  float* out;             // out_mode 0/1
  float* z_out;           // in_mode 2: sampled depths (n_rays, P) or NULL
  float* rgb_out;         // out_mode 2: (n_rays, 3)
  float* depth_out;       // out_mode 2: (n_rays) or NULL
  float* acc_out;         // out_mode 2: (n_rays) or NULL
  // out_mode 2, optional: the composited rgb row of ray r is ALSO stored at peer_rgb[q][(peer_row0 + r) * 3 ..] for every
  // q < n_peers -- the image buffers of the other GPUs of the node, mapped through CUDA IPC (stores travel over NVLink)
  float* peer_rgb[kMaxPeers];
  int n_peers;
  int64_t peer_row0;
  int flags;
  float near_t, far_t;
};
int launch_mlp_simt(nb2_handle* h, int net_id, const MlpIo& io, cudaStream_t st);
int launch_mlp_tc(nb2_handle* h, int net_id, int precision, const MlpIo& io, cudaStream_t st);
int launch_ipe_sumsq(nb2_handle* h, const float* rays, int64_t n_rays, double* out, cudaStream_t st);
int pack_network(nb2_handle* h, int net_id, const float* const* W, const float* const* b,
                 int n_layers, int pos_levels, int dir_levels, cudaStream_t st);

}  // namespace nb2
