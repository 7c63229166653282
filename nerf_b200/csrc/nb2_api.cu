// nb2_api.cu — handle lifetime, error reporting, weight packing, MLP forward and the fused
// render path of the C ABI declared in include/nerf_b200.h.
#include <algorithm>
#include <stdarg.h>

#include "nb2_common.cuh"

namespace nb2 {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int selftest_umma(nb2_handle* h, const void* A, const void* B, void* Bswz_scratch, float* D, cudaStream_t st);
int umma_bench(nb2_handle* h, const void* A, const void* B, float* D, long long* cycles, int mode, int iters, int flags,
               const void* gsrc, cudaStream_t st);
int selftest_umma_ts(nb2_handle* h, const void* A, const void* B, void* Bswz_scratch, float* D, cudaStream_t st);
int microbench(nb2_handle* h, int kind, int a0, int a1, int a2, int a3, int a4, const void* src, int n_chunks,
               long long* out, int* grid_out, cudaStream_t st);
}  // namespace nb2
using namespace nb2;

extern "C" const char* nb2_last_error(void) { return g_err; }
extern "C" int nb2_version(void) { return NB2_VERSION; }

extern "C" int nb2_create(nb2_handle** out, int device) {
  NB2_CHECK_ARG(out != nullptr, "nb2_create: null output pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("nb2_create: no CUDA device available (%s); this library has no CPU path",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return NB2_ERR_CUDA;
  }
  NB2_CHECK_ARG(device >= 0 && device < count, "nb2_create: device %d out of range [0,%d)", device, count);
  cudaDeviceProp prop;
  NB2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("nb2_create: device %d is sm_%d%d; libnerfb200 contains sm_100a code only", device, prop.major, prop.minor);
    return NB2_ERR_UNSUPPORTED;
  }
  NB2_CUDA(cudaSetDevice(device));
  nb2_handle* h = new nb2_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  *out = h;
  return NB2_OK;
}

extern "C" int nb2_destroy(nb2_handle* h) {
  if (!h) return NB2_OK;
  for (int i = 0; i < 2; ++i) {
    PackedNet& pn = h->net[i];
    if (pn.d_wchunks) cudaFree(pn.d_wchunks);
    if (pn.d_bias) cudaFree(pn.d_bias);
    if (pn.d_head) cudaFree(pn.d_head);
    if (pn.d_wt32) cudaFree(pn.d_wt32);
  }
  delete h;
  return NB2_OK;
}

extern "C" int64_t nb2_launch_count(nb2_handle* h) { return h ? h->launches : -1; }

// Debug: enable (n_ctas_max > 0) / read back the per-CTA role cycle counters of the LAST tensor-kernel launch.
// 16 counters per CTA: [0] streamer wait-empty, [1] streamer total, [2] ring entries, [3] MMA wait-A, [4] MMA wait-W,
// [5] MMA total, [6] group0 encode, [7] group0 wait-acc, [8] group0 hidden epilogues, [9] group0 last epilogue,
// [10] group0 total, [11] iterations, [12] layers.  Synchronises the device.
extern "C" int nb2_debug_tc_profile(nb2_handle* h, long long* out_host, int n_ctas) {
  NB2_CHECK_ARG(h != nullptr, "null handle");
  if (!h->tc_prof) {
    NB2_CUDA(cudaMalloc(&h->tc_prof, 256 * 16 * sizeof(long long)));
    NB2_CUDA(cudaMemset(h->tc_prof, 0, 256 * 16 * sizeof(long long)));
    return NB2_OK;
  }
  NB2_CHECK_ARG(out_host && n_ctas > 0 && n_ctas <= 256, "debug_tc_profile: bad arguments");
  NB2_CUDA(cudaDeviceSynchronize());
  NB2_CUDA(cudaMemcpy(out_host, h->tc_prof, (size_t)n_ctas * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return NB2_OK;
}

extern "C" int nb2_set_profile_events(nb2_handle* h, void* const* events4) {
  NB2_CHECK_ARG(h != nullptr, "null handle");
  for (int i = 0; i < 4; ++i) h->prof[i] = events4 ? (cudaEvent_t)events4[i] : nullptr;
  return NB2_OK;
}

extern "C" int nb2_pack_weights(nb2_handle* h, int net_id, const float* const* W, const float* const* b, int n_layers,
                                int pos_levels, int dir_levels, int hidden, void* stream) {
  NB2_CHECK_ARG(h != nullptr, "null handle");
  NB2_CHECK_ARG(net_id == NB2_NET_PROPOSAL || net_id == NB2_NET_NERF, "pack_weights: unknown network id %d", net_id);
  NB2_CHECK_ARG(W && b, "pack_weights: null pointer table");
  for (int i = 0; i < n_layers; ++i) NB2_CHECK_ARG(W[i] && b[i], "pack_weights: layer %d has a null weight or bias", i);
  if (hidden != kHidden) {
    set_error("pack_weights: hidden width %d is not supported (the kernels are built for %d)", hidden, kHidden);
    return NB2_ERR_UNSUPPORTED;
  }
  NB2_CHECK_ARG(pos_levels >= 1 && pos_levels <= kMaxPosLevels, "pack_weights: position levels must be in [1,%d]", kMaxPosLevels);
  NB2_CHECK_ARG(net_id == NB2_NET_PROPOSAL || (dir_levels >= 1 && dir_levels <= kMaxDirLevels),
                "pack_weights: direction levels must be in [1,%d]", kMaxDirLevels);
  return pack_network(h, net_id, W, b, n_layers, pos_levels, dir_levels, (cudaStream_t)stream);
}

extern "C" int nb2_weights_version(nb2_handle* h, int net_id) {
  if (!h || net_id < 0 || net_id > 1) return -1;
  return h->net[net_id].version;
}

static int mlp_dispatch(nb2_handle* h, int net_id, int precision, const MlpIo& io, cudaStream_t st) {
  if (precision == NB2_PREC_FP32) return launch_mlp_simt(h, net_id, io, st);
  if (precision == NB2_PREC_BF16 || precision == NB2_PREC_BF16X3 || precision == NB2_PREC_FP16 || precision == NB2_PREC_FP16X3)
    return launch_mlp_tc(h, net_id, precision, io, st);
  set_error("unknown precision %d", precision);
  return NB2_ERR_INVALID;
}

extern "C" int nb2_mlp_forward(nb2_handle* h, int net_id, int precision, const float* pts, int pts_stride,
                               int64_t n_points, float* out, void* stream) {
  NB2_CHECK_ARG(h != nullptr, "null handle");
  NB2_CHECK_ARG(net_id == NB2_NET_PROPOSAL || net_id == NB2_NET_NERF, "mlp_forward: unknown network id %d", net_id);
  if (n_points == 0) return NB2_OK;
  NB2_CHECK_ARG(pts && out && n_points > 0, "mlp_forward: bad arguments");
  NB2_CHECK_ARG(pts_stride >= (net_id == NB2_NET_NERF ? 6 : 3), "mlp_forward: pts_stride %d too small for network %d", pts_stride, net_id);
  MlpIo io;
  memset(&io, 0, sizeof(io));
  io.in_mode = 0;
  io.pts = pts;
  io.pts_stride = pts_stride;
  io.P = 1;
  io.n_rows = n_points;
  io.out_mode = (net_id == NB2_NET_NERF) ? 1 : 0;
  io.out = out;
  return mlp_dispatch(h, net_id, precision, io, (cudaStream_t)stream);
}

extern "C" int nb2_mlp_forward_encoded(nb2_handle* h, int net_id, int precision, const float* pts, int pts_stride,
                                       const float* encoded, int64_t n_points, float* out, void* stream) {
  NB2_CHECK_ARG(h != nullptr, "null handle");
  NB2_CHECK_ARG(net_id == NB2_NET_PROPOSAL, "mlp_forward_encoded: only the proposal network takes encoded_pt (nerf/addtional.py:88-91)");
  if (n_points == 0) return NB2_OK;
  NB2_CHECK_ARG(pts && encoded && out && n_points > 0 && pts_stride >= 3, "mlp_forward_encoded: bad arguments");
  if (!h->net[net_id].packed) {
    set_error("mlp_forward_encoded: weights of network %d have not been packed", net_id);
    return NB2_ERR_STATE;
  }
  MlpIo io;
  memset(&io, 0, sizeof(io));
  io.in_mode = 0;
  io.pts = pts;
  io.pts_stride = pts_stride;
  io.enc = encoded;
  io.P = h->net[net_id].pos_levels;   // row stride of `encoded` is 6 * pos_levels
  io.n_rows = n_points;
  io.out_mode = 0;
  io.out = out;
  return mlp_dispatch(h, net_id, precision, io, (cudaStream_t)stream);
}

static inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int64_t nb2_render_workspace_bytes(int64_t n_rays, const nb2_render_params* p) {
  if (!p || n_rays < 0) return -1;
  int64_t b = 0;
  b += align256(n_rays * p->n_coarse * 4);  // z_coarse
  b += align256(n_rays * p->n_coarse * 4);  // sigma_prop
  b += align256(n_rays * p->n_fine * 4);    // z_fine
  const bool fused = (p->precision != NB2_PREC_FP32) && (p->n_fine == 32 || p->n_fine == 64 || p->n_fine == 128);
  if (!fused) b += align256(n_rays * p->n_fine * 16);  // rgb-sigma per sample
  return b;
}

extern "C" int nb2_render_rays(nb2_handle* h, const nb2_render_params* p, const float* rays, const float* base_z,
                               const float* jitter, const float* u, int64_t n_rays, float* rgb_out, float* depth_out,
                               float* acc_out, float* z_coarse_out, float* sigma_prop_out, float* z_fine_out,
                               void* workspace, int64_t workspace_bytes, void* stream) {
  NB2_CHECK_ARG(h != nullptr, "null handle");
  NB2_CHECK_ARG(p != nullptr, "render_rays: null parameter block");
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(rays && base_z && rgb_out, "render_rays: null pointer");
  NB2_CHECK_ARG(p->n_coarse >= 3 && p->n_coarse <= 256, "render_rays: n_coarse must be in [3,256]");
  NB2_CHECK_ARG(p->n_fine >= 1 && p->n_fine + 1 <= 264, "render_rays: n_fine must be in [1,263]");
  NB2_CHECK_ARG(p->far_t != p->near_t, "render_rays: near == far");
  const int64_t need = nb2_render_workspace_bytes(n_rays, p);
  NB2_CHECK_ARG(workspace && workspace_bytes >= need, "render_rays: workspace too small (%lld < %lld bytes)",
                (long long)workspace_bytes, (long long)need);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* z_coarse = (float*)ws; ws += align256(n_rays * p->n_coarse * 4);
  float* sigma_prop = (float*)ws; ws += align256(n_rays * p->n_coarse * 4);
  float* z_fine = (float*)ws; ws += align256(n_rays * p->n_fine * 4);
  if (z_coarse_out) z_coarse = z_coarse_out;
  if (sigma_prop_out) sigma_prop = sigma_prop_out;
  if (z_fine_out) z_fine = z_fine_out;
  const bool fused = (p->precision != NB2_PREC_FP32) && (p->n_fine == 32 || p->n_fine == 64 || p->n_fine == 128);

  // 1. stratified sampling + encoding + proposal MLP           nerf/procedures.py:65-67
  MlpIo io;
  memset(&io, 0, sizeof(io));
  io.in_mode = 2;
  io.rays = rays;
  io.base_z = base_z;
  io.jitter = jitter;
  io.resolution = p->resolution;
  io.seed = p->seed;
  io.ray_offset = p->ray_offset;
  io.P = p->n_coarse;
  io.n_rows = n_rays * p->n_coarse;
  io.out_mode = 0;
  io.out = sigma_prop;
  io.z_out = z_coarse;
  if (h->prof[0]) NB2_CUDA(cudaEventRecord(h->prof[0], st));
  int rc = mlp_dispatch(h, NB2_NET_PROPOSAL, p->precision, io, st);
  if (rc != NB2_OK) return rc;
  if (h->prof[1]) NB2_CUDA(cudaEventRecord(h->prof[1], st));

  // 2. density -> weights -> max-blur -> inverse CDF -> sort -> drop last      :68-70,76
  rc = nb2_resample(h, sigma_prop, z_coarse, rays, u, p->seed, p->ray_offset, n_rays, p->n_coarse, p->n_fine + 1,
                    p->blur_alpha, p->flags, z_fine, stream);
  if (rc != NB2_OK) return rc;
  if (h->prof[2]) NB2_CUDA(cudaEventRecord(h->prof[2], st));

  // 3. encoding + NeRF MLP + alpha compositing                                    :77-85
  memset(&io, 0, sizeof(io));
  io.in_mode = 1;
  io.rays = rays;
  io.z = z_fine;
  io.P = p->n_fine;
  io.n_rows = n_rays * p->n_fine;
  io.flags = p->flags;
  io.near_t = p->near_t;
  io.far_t = p->far_t;
  if (fused) {
    io.out_mode = 2;
    io.rgb_out = rgb_out;
    io.depth_out = depth_out;
    io.acc_out = acc_out;
    rc = mlp_dispatch(h, NB2_NET_NERF, p->precision, io, st);
    if (rc == NB2_OK && h->prof[3]) NB2_CUDA(cudaEventRecord(h->prof[3], st));
    return rc;
  }
  float* rgbo = (float*)ws;
  io.out_mode = 1;
  io.out = rgbo;
  rc = mlp_dispatch(h, NB2_NET_NERF, p->precision, io, st);
  if (rc != NB2_OK) return rc;
  rc = nb2_composite(h, rgbo, z_fine, rays + 3, 6, n_rays, p->n_fine, p->flags, p->near_t, p->far_t, rgb_out, nullptr,
                     depth_out, acc_out, stream);
  if (rc == NB2_OK && h->prof[3]) NB2_CUDA(cudaEventRecord(h->prof[3], st));
  return rc;
}

extern "C" int nb2_selftest_umma(nb2_handle* h, const void* A_bf16, const void* B_bf16, void* scratch_16k, float* D_out,
                                 void* stream) {
  NB2_CHECK_ARG(h && A_bf16 && B_bf16 && scratch_16k && D_out, "selftest_umma: null pointer");
  return selftest_umma(h, A_bf16, B_bf16, scratch_16k, D_out, (cudaStream_t)stream);
}

extern "C" int nb2_debug_umma_bench(nb2_handle* h, const void* A_bf16, const void* B_bf16, float* D_out, long long* cycles_out,
                                    int mode, int iters, int flags, const void* gsrc_1mb, void* stream) {
  NB2_CHECK_ARG(h && A_bf16 && B_bf16 && D_out && cycles_out && gsrc_1mb && mode >= 0 && mode <= 2 && iters > 0, "debug_umma_bench: bad arguments");
  return umma_bench(h, A_bf16, B_bf16, D_out, cycles_out, mode, iters, flags, gsrc_1mb, (cudaStream_t)stream);
}

extern "C" int nb2_debug_microbench(nb2_handle* h, int kind, int a0, int a1, int a2, int a3, int a4, const void* src, int n_chunks,
                                    long long* out_dev, int* grid_out, void* stream) {
  NB2_CHECK_ARG(h && out_dev && grid_out, "debug_microbench: null pointer");
  return microbench(h, kind, a0, a1, a2, a3, a4, src, n_chunks, out_dev, grid_out, (cudaStream_t)stream);
}

extern "C" int nb2_selftest_umma_ts(nb2_handle* h, const void* A_bf16, const void* B_bf16, void* scratch_16k, float* D_out,
                                    void* stream) {
  NB2_CHECK_ARG(h && A_bf16 && B_bf16 && scratch_16k && D_out, "selftest_umma_ts: null pointer");
  return selftest_umma_ts(h, A_bf16, B_bf16, scratch_16k, D_out, (cudaStream_t)stream);
}
