// nb2_api.cu — handle lifetime, error reporting, weight packing, MLP forward and the fused
// render path of the C ABI declared in include/nerf_b200.h.
#include <algorithm>
#include <stdarg.h>
#include <stdlib.h>

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
  nb2_handle* h = new nb2_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  // the one runtime knob of the tensor kernels is read here, once (timing experiments only: results are garbage)
  const char* dbg = getenv("NB2_TC_DEBUG");
  h->tc_debug = dbg ? atoi(dbg) : 0;
  h->net[NB2_NET_PROPOSAL].in_use = true;
  h->net[NB2_NET_PROPOSAL].kind = NB2_NET_PROPOSAL;
  h->net[NB2_NET_NERF].in_use = true;
  h->net[NB2_NET_NERF].kind = NB2_NET_NERF;
  *out = h;
  return NB2_OK;
}

static void free_net(PackedNet& pn) {
  if (pn.d_wchunks) cudaFree(pn.d_wchunks);
  if (pn.d_bias) cudaFree(pn.d_bias);
  if (pn.d_head) cudaFree(pn.d_head);
  if (pn.d_wt32) cudaFree(pn.d_wt32);
  pn = PackedNet();
}

extern "C" int nb2_destroy(nb2_handle* h) {
  if (!h) return NB2_OK;
  {
    DeviceGuard g(h->device);
    for (int i = 0; i < kMaxNets; ++i) free_net(h->net[i]);
    if (h->tc_prof) cudaFree(h->tc_prof);
  }
  delete h;
  return NB2_OK;
}

extern "C" int nb2_net_create(nb2_handle* h, int kind, int* net_id_out) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(net_id_out != nullptr, "net_create: null output pointer");
  NB2_CHECK_ARG(kind == NB2_NET_PROPOSAL || kind == NB2_NET_NERF, "net_create: unknown network kind %d", kind);
  for (int i = 2; i < kMaxNets; ++i)
    if (!h->net[i].in_use) {
      h->net[i].in_use = true;
      h->net[i].kind = kind;
      *net_id_out = i;
      return NB2_OK;
    }
  set_error("net_create: all %d packed-network slots of this handle are in use", kMaxNets);
  return NB2_ERR_STATE;
}

extern "C" int nb2_net_destroy(nb2_handle* h, int net_id) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(net_id >= 2 && net_ok(h, net_id), "net_destroy: %d is not a slot created by nb2_net_create", net_id);
  // the slot's buffers may still be read by kernels in flight on any stream of this device
  NB2_CUDA(cudaDeviceSynchronize());
  free_net(h->net[net_id]);
  return NB2_OK;
}

namespace nb2 {
int kernel_set_smem(nb2_handle* h, const void* fn, int bytes) {
  KernelCache& kc = h->kcache[fn];
  if (!kc.attr_set) {
    NB2_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    kc.attr_set = true;
  }
  return NB2_OK;
}
int kernel_max_clusters(nb2_handle* h, const void* fn, const cudaLaunchConfig_t* cfg, int* out) {
  KernelCache& kc = h->kcache[fn];
  if (!kc.max_clusters) {
    int n = 0;
    NB2_CUDA(cudaOccupancyMaxActiveClusters(&n, fn, cfg));
    kc.max_clusters = n > 0 ? n : 1;
  }
  *out = kc.max_clusters;
  return NB2_OK;
}
}  // namespace nb2

extern "C" int64_t nb2_launch_count(nb2_handle* h) { return h ? h->launches : -1; }

// Debug: enable (n_ctas_max > 0) / read back the per-CTA role cycle counters of the LAST tensor-kernel launch.
// 16 counters per CTA: [0] streamer wait-empty, [1] streamer total, [2] ring entries, [3] MMA wait-A, [4] MMA wait-W,
// [5] MMA total, [6] group0 encode, [7] group0 wait-acc, [8] group0 hidden epilogues, [9] group0 last epilogue,
// [10] group0 total, [11] iterations, [12] layers.  Synchronises the device.
extern "C" int nb2_debug_tc_profile(nb2_handle* h, long long* out_host, int n_ctas) {
  NB2_ENTER(h);
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
  NB2_ENTER(h);
  for (int i = 0; i < 4; ++i) h->prof[i] = events4 ? (cudaEvent_t)events4[i] : nullptr;
  return NB2_OK;
}

extern "C" int nb2_pack_weights(nb2_handle* h, int net_id, const float* const* W, const float* const* b, int n_layers,
                                int pos_levels, int dir_levels, int hidden, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(net_ok(h, net_id), "pack_weights: unknown network slot %d", net_id);
  NB2_CHECK_ARG(W && b, "pack_weights: null pointer table");
  for (int i = 0; i < n_layers; ++i) NB2_CHECK_ARG(W[i] && b[i], "pack_weights: layer %d has a null weight or bias", i);
  if (hidden != kHidden) {
    set_error("pack_weights: hidden width %d is not supported (the kernels are built for %d)", hidden, kHidden);
    return NB2_ERR_UNSUPPORTED;
  }
  NB2_CHECK_ARG(pos_levels >= 1 && pos_levels <= kMaxPosLevels, "pack_weights: position levels must be in [1,%d]", kMaxPosLevels);
  NB2_CHECK_ARG(h->net[net_id].kind == NB2_NET_PROPOSAL || (dir_levels >= 1 && dir_levels <= kMaxDirLevels),
                "pack_weights: direction levels must be in [1,%d]", kMaxDirLevels);
  return pack_network(h, net_id, W, b, n_layers, pos_levels, dir_levels, (cudaStream_t)stream);
}

extern "C" int nb2_weights_version(nb2_handle* h, int net_id) {
  if (!h || !net_ok(h, net_id)) return -1;
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
  NB2_ENTER(h);
  NB2_CHECK_ARG(net_ok(h, net_id), "mlp_forward: unknown network slot %d", net_id);
  if (n_points == 0) return NB2_OK;
  NB2_CHECK_ARG(pts && out && n_points > 0, "mlp_forward: bad arguments");
  const bool is_nerf = h->net[net_id].kind == NB2_NET_NERF;
  NB2_CHECK_ARG(pts_stride >= (is_nerf ? 6 : 3), "mlp_forward: pts_stride %d too small for network %d", pts_stride, net_id);
  MlpIo io;
  memset(&io, 0, sizeof(io));
  io.in_mode = 0;
  io.pts = pts;
  io.pts_stride = pts_stride;
  io.P = 1;
  io.n_rows = n_points;
  io.out_mode = is_nerf ? 1 : 0;
  io.out = out;
  return mlp_dispatch(h, net_id, precision, io, (cudaStream_t)stream);
}

extern "C" int nb2_mlp_forward_encoded(nb2_handle* h, int net_id, int precision, const float* pts, int pts_stride,
                                       const float* encoded, int64_t n_points, float* out, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(net_ok(h, net_id) && h->net[net_id].kind == NB2_NET_PROPOSAL,
                "mlp_forward_encoded: only a proposal network takes encoded_pt (nerf/addtional.py:88-91)");
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
  b += 256;                                 // scalars (batch-global direction norm of the IPE mode)
  const bool fused = (p->precision != NB2_PREC_FP32) && (p->n_fine == 32 || p->n_fine == 64 || p->n_fine == 128);
  if (!fused) b += align256(n_rays * p->n_fine * 16);  // rgb-sigma per sample
  return b;
}

extern "C" int nb2_render_rays(nb2_handle* h, const nb2_render_params* p, const float* rays, const float* base_z,
                               const float* jitter, const float* u, int64_t n_rays, float* rgb_out, float* depth_out,
                               float* acc_out, float* z_coarse_out, float* sigma_prop_out, float* z_fine_out,
                               int64_t* below_fine_out, void* workspace, int64_t workspace_bytes, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(p != nullptr, "render_rays: null parameter block");
  const int prop_id = p->prop_net_id, nerf_id = p->nerf_net_id == 0 ? (int)NB2_NET_NERF : p->nerf_net_id;
  NB2_CHECK_ARG(net_ok(h, prop_id) && h->net[prop_id].kind == NB2_NET_PROPOSAL, "render_rays: slot %d is not a proposal network", prop_id);
  NB2_CHECK_ARG(net_ok(h, nerf_id) && h->net[nerf_id].kind == NB2_NET_NERF, "render_rays: slot %d is not a NeRF network", nerf_id);
  NB2_CHECK_ARG(p->n_peers >= 0 && p->n_peers <= kMaxPeers, "render_rays: n_peers must be in [0,%d]", kMaxPeers);
  for (int q = 0; q < p->n_peers; ++q) NB2_CHECK_ARG(p->peer_rgb[q] != nullptr, "render_rays: peer_rgb[%d] is null", q);
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
  double* scalars = (double*)ws; ws += 256;
  if (z_coarse_out) z_coarse = z_coarse_out;
  if (sigma_prop_out) sigma_prop = sigma_prop_out;
  if (z_fine_out) z_fine = z_fine_out;
  const bool fused = (p->precision != NB2_PREC_FP32) && (p->n_fine == 32 || p->n_fine == 64 || p->n_fine == 128);
  const int prec_prop = p->precision == NB2_PREC_FP16_MIXED ? (int)NB2_PREC_FP16X3 : (p->precision == NB2_PREC_BF16_MIXED ? (int)NB2_PREC_BF16X3 : p->precision);
  const int prec_fine = p->precision == NB2_PREC_FP16_MIXED ? (int)NB2_PREC_FP16 : (p->precision == NB2_PREC_BF16_MIXED ? (int)NB2_PREC_BF16 : p->precision);

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
  if (p->flags & NB2_PROPOSAL_IPE) {
    // BASELINE configs[2]: the proposal network reads the integrated positional encoding of the conical frustums between
    // consecutive coarse depths (nerf/mip_methods.py:47-58 -> ProposalNetwork.forward(mu, encoded_pt), addtional.py:88-91)
    NB2_CHECK_ARG(p->ipe_radius > 0.f, "render_rays: NB2_PROPOSAL_IPE needs ipe_radius > 0");
    int rc0 = launch_ipe_sumsq(h, rays, n_rays, scalars, st);
    if (rc0 != NB2_OK) return rc0;
    io.ipe = 1;
    io.ipe_radius = p->ipe_radius;
    io.ipe_last_step = (p->far_t - p->near_t) / (float)(p->n_coarse - 1);
    io.ipe_sumsq = scalars;
  }
  if (h->prof[0]) NB2_CUDA(cudaEventRecord(h->prof[0], st));
  int rc = mlp_dispatch(h, prop_id, prec_prop, io, st);
  if (rc != NB2_OK) return rc;
  if (h->prof[1]) NB2_CUDA(cudaEventRecord(h->prof[1], st));

  // 2. density -> weights -> max-blur -> inverse CDF -> sort -> drop last      :68-70,76
  rc = nb2_resample(h, sigma_prop, z_coarse, rays, u, p->seed, p->ray_offset, n_rays, p->n_coarse, p->n_fine + 1,
                    p->blur_alpha, p->flags, z_fine, below_fine_out, stream);
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
    io.n_peers = p->n_peers;
    io.peer_row0 = p->ray_offset;
    for (int q = 0; q < p->n_peers; ++q) io.peer_rgb[q] = p->peer_rgb[q];
    rc = mlp_dispatch(h, nerf_id, prec_fine, io, st);
    if (rc == NB2_OK && h->prof[3]) NB2_CUDA(cudaEventRecord(h->prof[3], st));
    return rc;
  }
  float* rgbo = (float*)ws;
  io.out_mode = 1;
  io.out = rgbo;
  rc = mlp_dispatch(h, nerf_id, prec_fine, io, st);
  if (rc != NB2_OK) return rc;
  rc = nb2_composite(h, rgbo, z_fine, rays + 3, 6, n_rays, p->n_fine, p->flags, p->near_t, p->far_t, rgb_out, nullptr,
                     depth_out, acc_out, stream);
  if (rc == NB2_OK && p->n_peers > 0) {
    // unfused compositing (fp32 CUDA-core mode, odd sample counts): the peers get this shard's rows by copy
    for (int q = 0; q < p->n_peers; ++q)
      NB2_CUDA(cudaMemcpyAsync(p->peer_rgb[q] + 3 * p->ray_offset, rgb_out, (size_t)n_rays * 3 * sizeof(float), cudaMemcpyDefault, st));
  }
  if (rc == NB2_OK && h->prof[3]) NB2_CUDA(cudaEventRecord(h->prof[3], st));
  return rc;
}

extern "C" int nb2_selftest_umma(nb2_handle* h, const void* A_bf16, const void* B_bf16, void* scratch_16k, float* D_out,
                                 void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(A_bf16 && B_bf16 && scratch_16k && D_out, "selftest_umma: null pointer");
  return selftest_umma(h, A_bf16, B_bf16, scratch_16k, D_out, (cudaStream_t)stream);
}

extern "C" int nb2_debug_umma_bench(nb2_handle* h, const void* A_bf16, const void* B_bf16, float* D_out, long long* cycles_out,
                                    int mode, int iters, int flags, const void* gsrc_1mb, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(A_bf16 && B_bf16 && D_out && cycles_out && gsrc_1mb && mode >= 0 && mode <= 2 && iters > 0, "debug_umma_bench: bad arguments");
  return umma_bench(h, A_bf16, B_bf16, D_out, cycles_out, mode, iters, flags, gsrc_1mb, (cudaStream_t)stream);
}

extern "C" int nb2_debug_microbench(nb2_handle* h, int kind, int a0, int a1, int a2, int a3, int a4, const void* src, int n_chunks,
                                    long long* out_dev, int* grid_out, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(out_dev && grid_out, "debug_microbench: null pointer");
  return microbench(h, kind, a0, a1, a2, a3, a4, src, n_chunks, out_dev, grid_out, (cudaStream_t)stream);
}

extern "C" int nb2_selftest_umma_ts(nb2_handle* h, const void* A_bf16, const void* B_bf16, void* scratch_16k, float* D_out,
                                    void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(A_bf16 && B_bf16 && scratch_16k && D_out, "selftest_umma_ts: null pointer");
  return selftest_umma_ts(h, A_bf16, B_bf16, scratch_16k, D_out, (cudaStream_t)stream);
}

// ---- peer memory for the fused multi-GPU gather ----------------------------------------------------------------------
extern "C" int nb2_ipc_alloc(nb2_handle* h, int64_t bytes, void** dev_ptr_out, void* ipc_handle64_out) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(bytes > 0 && dev_ptr_out && ipc_handle64_out, "ipc_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  void* p = nullptr;
  NB2_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t mh;
  cudaError_t e = cudaIpcGetMemHandle(&mh, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return NB2_ERR_CUDA;
  }
  memcpy(ipc_handle64_out, &mh, sizeof(mh));
  *dev_ptr_out = p;
  return NB2_OK;
}
extern "C" int nb2_ipc_open(nb2_handle* h, const void* ipc_handle64, void** dev_ptr_out) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(ipc_handle64 && dev_ptr_out, "ipc_open: null pointer");
  cudaIpcMemHandle_t mh;
  memcpy(&mh, ipc_handle64, sizeof(mh));
  NB2_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, mh, cudaIpcMemLazyEnablePeerAccess));
  return NB2_OK;
}
extern "C" int nb2_ipc_close(nb2_handle* h, void* dev_ptr) {
  NB2_ENTER(h);
  if (dev_ptr) NB2_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return NB2_OK;
}
extern "C" int nb2_ipc_free(nb2_handle* h, void* dev_ptr) {
  NB2_ENTER(h);
  if (dev_ptr) NB2_CUDA(cudaFree(dev_ptr));
  return NB2_OK;
}
