// nb2_train.cu — the HBM-bound pieces of the training step around the layer-wise GEMM engine:
//   * encode_bf16_kernel: sample points -> [x, sin(2^l x), cos(2^l x)] rows written as bf16 hi / lo into a (column view
//     of a) GEMM input buffer                                   (nerf/nerf_helper.py:38-48, nerf/mip_model.py:42-52)
//   * backward of ProposalNetwork.get_weights / NeRF.getNormedWeight  (nerf/addtional.py:99-107, nerf_base.py:79-86)
//   * backward of NeRF.render                                         (nerf/nerf_base.py:90-113)
//   * backward of maxBlurFilter                                       (nerf/mip_methods.py:61-66)
//   * backward of getBounds                                           (nerf/addtional.py:14-18)
//   * head glue of MipNeRF.forward's backward (sigmoid', packing d rgb / d sigma as GEMM operands)
// In the reference all of these are autograd's derivative formulas for the torch ops the functions are written in; the
// formulas are restated per kernel.  One warp per ray for the scan-shaped ones, as in nb2_ops.cu.
#include "nb2_common.cuh"
#include "nb2_tc_ptx.cuh"

namespace nb2 {
using namespace ptx;

static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }
constexpr int kTrMaxSamples = 256;
constexpr int kTrWarps = 4;

__device__ __forceinline__ float dir_norm3(const float* d) {
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
}

// ---- encoding rows as GEMM operands -----------------------------------------------------------------------------------
// x: (n, x_stride) fp32, columns [x_col0, x_col0 + 3) are the point (or the direction, normalised first when `normalize`);
// row r of the output view gets [x(3), sin(2^0 x)(3), cos(2^0 x)(3), ...] in columns [0, 3 + 6 L) and zeros up to `width`.
__global__ void encode_bf16_kernel(const float* __restrict__ x, int x_stride, int x_col0, int64_t n, int levels, int normalize,
                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld, int width) {
  const int64_t total = n * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    float v = 0.f;
    if (c < 3 + 6 * levels) {
      const float* p = x + r * x_stride + x_col0;
      float px = __ldg(p), py = __ldg(p + 1), pz = __ldg(p + 2);
      if (normalize) {
        const float d[3] = {px, py, pz};
        const float nrm = dir_norm3(d);
        px = __fdiv_rn(px, nrm); py = __fdiv_rn(py, nrm); pz = __fdiv_rn(pz, nrm);
      }
      const int k = c < 3 ? c : (c - 3) % 3;
      const float xv = k == 0 ? px : (k == 1 ? py : pz);
      if (c < 3) {
        v = xv;
      } else {
        const int l = (c - 3) / 6, w = (c - 3) % 6;
        float s, co;
        sincos_any(xv * exp2f((float)l), s, co);
        v = w < 3 ? s : co;
      }
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[r * ld + c] = h;
    if (lo) lo[r * ld + c] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// Tiled form (width, ld multiples of 8; 16-byte aligned outputs): one thread per (point, component) runs down the levels
// with the exact doubling v <- 2 v and ONE sin/cos range test, a block's 128 rows are assembled in shared memory (hi and lo)
// and leave as 16-byte vectors.  The element-per-thread kernel above evaluated a full sin/cos pair per output column (half of
// it unused), paid 64-bit index divisions per element and stored 2-byte scalars: 250 us per 1M points against 41 us of bytes.
constexpr int kEncPts = 128;
__global__ void __launch_bounds__(3 * kEncPts)
encode3_bf16_kernel(const float* __restrict__ x, int x_stride, int x_col0, int64_t n, int levels, int normalize,
                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld, int width) {
  extern __shared__ __align__(16) unsigned char enc_smem[];
  const int pitch = width + 8;                     // 16-byte aligned rows that do not all start in the same bank
  __nv_bfloat16* t_hi = reinterpret_cast<__nv_bfloat16*>(enc_smem);              // [kEncPts][pitch]
  __nv_bfloat16* t_lo = t_hi + kEncPts * pitch;
  const int64_t p0 = (int64_t)blockIdx.x * kEncPts;
  const int npts = (int)min((int64_t)kEncPts, n - p0);
  const int t = threadIdx.x;
  const int p = t / 3, k = t - 3 * p;
  if (p < npts) {
    const float* src = x + (p0 + p) * x_stride + x_col0;
    float v = __ldg(src + k);
    if (normalize) {
      const float d[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
      v = __fdiv_rn(v, dir_norm3(d));
    }
    __nv_bfloat16* rh = t_hi + p * pitch;
    __nv_bfloat16* rl = t_lo + p * pitch;
    auto put = [&](int c, float val) {
      const __nv_bfloat16 h = __float2bfloat16_rn(val);
      rh[c] = h;
      rl[c] = __float2bfloat16_rn(val - __bfloat162float(h));
    };
    put(k, v);
    const bool small = fabsf(v) * exp2f((float)(levels > 0 ? levels - 1 : 0)) <= 1.0e5f;
    for (int l = 0; l < levels; ++l) {
      float sn, cs;
      if (small) enc_sincos(v, sn, cs); else sincos_any(v, sn, cs);
      put(3 + 6 * l + k, sn);
      put(3 + 6 * l + 3 + k, cs);
      v = v * 2.f;
    }
    if (k == 0)
      for (int c = 3 + 6 * levels; c < width; ++c) { rh[c] = __float2bfloat16_rn(0.f); rl[c] = __float2bfloat16_rn(0.f); }
  }
  __syncthreads();
  const int cpr = width >> 3;                       // 16-byte chunks per row
  for (int q = t; q < npts * cpr; q += 3 * kEncPts) {
    const int row = q / cpr, ch = q - row * cpr;
    const int64_t off = (p0 + row) * ld + ch * 8;
    *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<const uint4*>(t_hi + row * pitch + ch * 8);
    if (lo) *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<const uint4*>(t_lo + row * pitch + ch * 8);
  }
}

// ---- shared forward recomputation: m_i, T_i, w_i of one ray ------------------------------------------------------------
struct RaySmem {
  float depth[kTrMaxSamples];
  float m[kTrMaxSamples];      // exp(-act(sigma) delta)
  float T[kTrMaxSamples];      // exclusive product of (m + 1e-10)
  float G[kTrMaxSamples];      // upstream gradient G_i = dL/dw_i
  float gw[kTrMaxSamples];     // G_i w_i (suffix-summed)
  float dact[kTrMaxSamples];   // act'(sigma) * delta
};

__device__ __forceinline__ float act_grad(float x, int act) {
  return act == 0 ? (x > 0.f ? 1.f : 0.f) : (act == 1 ? 1.f / (1.f + expf(-x)) : 1.f);   // relu | softplus | identity
}

// fills s.m, s.T, s.dact from densities (read through `sig(i)`) and s.depth
template <class SigFn>
__device__ __forceinline__ void ray_forward(RaySmem& s, SigFn sig, int P, int act, int lane) {
  float carry = 1.f;
  for (int base = 0; base < P; base += 32) {
    const int i = base + lane;
    float m = 1.f;
    if (i < P) {
      const float delta = (i + 1 < P) ? __fsub_rn(s.depth[i + 1], s.depth[i]) : 1e10f;
      const float x = sig(i);
      m = expf(-apply_density_act(x, act) * delta);
      s.m[i] = m;
      s.dact[i] = act_grad(x, act) * delta;
    }
    const float f = (i < P) ? (m + 1e-10f) : 1.f;
    const float inc = warp_scan_mul(f, lane);
    float exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 1.f;
    if (i < P) s.T[i] = carry * exc;
    carry *= __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
}

// d sigma_k = (-G_k T_k + S_k / (m_k + 1e-10)) * (-delta_k m_k act'(sigma_k)),  S_k = sum_{i > k} G_i w_i,  w_i = (1 - m_i) T_i
// (G_i = dL/dw_i in s.G).  out(i, S_i) is called once per sample.
template <class OutFn>
__device__ __forceinline__ void ray_backward(RaySmem& s, int P, int lane, OutFn out) {
  for (int i = lane; i < P; i += 32) s.gw[i] = s.G[i] * ((1.f - s.m[i]) * s.T[i]);   // G_i w_i
  __syncwarp();
  float carry = 0.f;                                                                // sum over samples beyond this block
  for (int base = ((P - 1) / 32) * 32; base >= 0; base -= 32) {
    const int i = base + lane;
    const float v = (i < P) ? s.gw[i] : 0.f;
    float suf = v;                                                                   // inclusive suffix sum inside the block
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += t;
    }
    const float S = carry + suf - v;                                                 // exclusive
    if (i < P) out(i, S);
    carry += __shfl_sync(0xffffffffu, suf, 0);
  }
}

// backward of get_weights / getNormedWeight: g_w (R,P) -> d_sigma (R,P)
__global__ void __launch_bounds__(32 * kTrWarps)
weights_backward_kernel(const float* __restrict__ sigma, const float* __restrict__ z, const float* __restrict__ dirs, int dir_stride,
                        int64_t n_rays, int P, int act, const float* __restrict__ g_w, float* __restrict__ d_sigma) {
  __shared__ RaySmem sm[kTrWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kTrWarps + warp;
  if (r >= n_rays) return;
  RaySmem& s = sm[warp];
  const float nrm = dirs ? dir_norm3(dirs + r * dir_stride) : 1.f;
  for (int i = lane; i < P; i += 32) {
    const float zz = z[r * P + i];
    s.depth[i] = dirs ? __fmul_rn(zz, nrm) : zz;
    s.G[i] = g_w[r * P + i];
  }
  __syncwarp();
  const float* sg = sigma + r * P;
  ray_forward(s, [&](int i) { return sg[i]; }, P, act, lane);
  ray_backward(s, P, lane, [&](int i, float S) {
    d_sigma[r * P + i] = (-s.G[i] * s.T[i] + S / (s.m[i] + 1e-10f)) * (-s.m[i] * s.dact[i]);
  });
}

// backward of NeRF.render: g_rgb (R,3), g_w (R,P) or NULL -> d_rgbo (R,P,4)
//   rgb = sum_i w_i c_i (+ 1 - sum_i w_i with a white background); weights = w
__global__ void __launch_bounds__(32 * kTrWarps)
composite_backward_kernel(const float* __restrict__ rgbo, const float* __restrict__ z, const float* __restrict__ dirs, int dir_stride,
                          int64_t n_rays, int P, int flags, const float* __restrict__ g_rgb, const float* __restrict__ g_w,
                          float* __restrict__ d_rgbo) {
  __shared__ RaySmem sm[kTrWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kTrWarps + warp;
  if (r >= n_rays) return;
  RaySmem& s = sm[warp];
  const float nrm = dir_norm3(dirs + r * dir_stride);
  const float4* c4 = reinterpret_cast<const float4*>(rgbo) + r * P;
  const float gr = g_rgb[r * 3 + 0], gg = g_rgb[r * 3 + 1], gb = g_rgb[r * 3 + 2];
  const float bkg = (flags & NB2_WHITE_BKG) ? (gr + gg + gb) : 0.f;
  for (int i = lane; i < P; i += 32) {
    s.depth[i] = __fmul_rn(z[r * P + i], nrm);
    const float4 c = c4[i];
    s.G[i] = (g_w ? g_w[r * P + i] : 0.f) + gr * c.x + gg * c.y + gb * c.z - bkg;
  }
  __syncwarp();
  ray_forward(s, [&](int i) { return c4[i].w; }, P, 0, lane);
  float4* o4 = reinterpret_cast<float4*>(d_rgbo) + r * P;
  ray_backward(s, P, lane, [&](int i, float S) {
    const float w = (1.f - s.m[i]) * s.T[i];
    const float ds = (-s.G[i] * s.T[i] + S / (s.m[i] + 1e-10f)) * (-s.m[i] * s.dact[i]);
    o4[i] = make_float4(gr * w, gg * w, gb * w, ds);
  });
}

// backward of maxBlurFilter: out_i = 0.5 (front_i + rear_i) + alpha, front = [w_0, mx], rear = [mx, w_{P-1}], mx_j = max(w_j, w_{j+1})
// torch.maximum's derivative: the larger argument takes the gradient, equal arguments share it.
__global__ void max_blur_backward_kernel(const float* __restrict__ w, const float* __restrict__ g, int64_t n_rays, int P, float* __restrict__ dw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * P) return;
  const int64_t r = i / P;
  const int k = (int)(i - r * P);
  const float* wr = w + r * P;
  const float* gr = g + r * P;
  float acc = 0.f;
  if (k == 0) acc += 0.5f * gr[0];
  if (k == P - 1) acc += 0.5f * gr[P - 1];
  if (k + 1 < P) {                       // mx_k = max(w_k, w_{k+1}) feeds front_{k+1} and rear_k
    const float gm = 0.5f * (gr[k + 1] + gr[k]);
    acc += wr[k] > wr[k + 1] ? gm : (wr[k] == wr[k + 1] ? 0.5f * gm : 0.f);
  }
  if (k >= 1) {                          // mx_{k-1} = max(w_{k-1}, w_k) feeds front_k and rear_{k-1}
    const float gm = 0.5f * (gr[k] + gr[k - 1]);
    acc += wr[k] > wr[k - 1] ? gm : (wr[k] == wr[k - 1] ? 0.5f * gm : 0.f);
  }
  dw[i] = acc;
}

// backward of getBounds: out_j = sat[e_j] - sat[s_j], sat[k] = sum_{i < k} w_i  ->  dw_i = sum_j g_j [s_j <= i < e_j]
__global__ void __launch_bounds__(32 * kTrWarps)
get_bounds_backward_kernel(const int64_t* __restrict__ inds, const float* __restrict__ g, int64_t n_rays, int P, int K, float* __restrict__ dw) {
  __shared__ int sh_s[kTrWarps][kTrMaxSamples + 8], sh_e[kTrWarps][kTrMaxSamples + 8];
  __shared__ float sh_g[kTrWarps][kTrMaxSamples + 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kTrWarps + warp;
  if (r >= n_rays) return;
  for (int j = lane; j < K - 1; j += 32) {
    int64_t st = inds[r * K + j], en = inds[r * K + j + 1] + 1;
    st = st < 0 ? 0 : (st > P ? P : st);
    en = en < 0 ? 0 : (en > P ? P : en);
    sh_s[warp][j] = (int)st;
    sh_e[warp][j] = (int)en;
    sh_g[warp][j] = g[r * (K - 1) + j];
  }
  __syncwarp();
  for (int i = lane; i < P; i += 32) {
    float acc = 0.f;
    for (int j = 0; j < K - 1; ++j) {
      // sat[e] - sat[s] with e < s (indices not ascending) is minus the span [e, s)
      const int a = sh_s[warp][j], b = sh_e[warp][j];
      if (a <= i && i < b) acc += sh_g[warp][j];
      else if (b <= i && i < a) acc -= sh_g[warp][j];
    }
    dw[r * P + i] = acc;
  }
}

// MipNeRF head glue for the backward pass: out (n,4) = [sigmoid rgb, sigma], g_out (n,4) ->
//   d_z (n, 8) bf16 hi/lo: columns 0..2 = g_rgb * rgb (1 - rgb) (the rgb_layer.2 pre-activation gradient), rest 0
//   d_s (n, 8) bf16 hi/lo: column 0 = g_sigma, rest 0
__global__ void nerf_head_backward_kernel(const float* __restrict__ out, const float* __restrict__ g_out, int64_t n,
                                          __nv_bfloat16* __restrict__ dz_hi, __nv_bfloat16* __restrict__ dz_lo,
                                          __nv_bfloat16* __restrict__ ds_hi, __nv_bfloat16* __restrict__ ds_lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 o = reinterpret_cast<const float4*>(out)[i];
  const float4 g = reinterpret_cast<const float4*>(g_out)[i];
  const float v[4] = {g.x * o.x * (1.f - o.x), g.y * o.y * (1.f - o.y), g.z * o.z * (1.f - o.z), g.w};
  uint32_t zh[4] = {0, 0, 0, 0}, zl[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0}, sl[4] = {0, 0, 0, 0};
  zh[0] = pack_bf16x2(v[0], v[1]);
  zl[0] = pack_bf16x2(v[0] - bf16_lo_to_f32(zh[0]), v[1] - bf16_hi_to_f32(zh[0]));
  zh[1] = pack_bf16x2(v[2], 0.f);
  zl[1] = pack_bf16x2(v[2] - bf16_lo_to_f32(zh[1]), 0.f);
  sh[0] = pack_bf16x2(v[3], 0.f);
  sl[0] = pack_bf16x2(v[3] - bf16_lo_to_f32(sh[0]), 0.f);
  reinterpret_cast<uint4*>(dz_hi)[i] = make_uint4(zh[0], zh[1], zh[2], zh[3]);
  reinterpret_cast<uint4*>(ds_hi)[i] = make_uint4(sh[0], sh[1], sh[2], sh[3]);
  if (dz_lo) reinterpret_cast<uint4*>(dz_lo)[i] = make_uint4(zl[0], zl[1], zl[2], zl[3]);
  if (ds_lo) reinterpret_cast<uint4*>(ds_lo)[i] = make_uint4(sl[0], sl[1], sl[2], sl[3]);
}

}  // namespace nb2

using namespace nb2;

extern "C" int nb2_encode_bf16(nb2_handle* h, const float* x, int x_stride, int x_col0, int64_t n, int levels, int normalize, void* hi,
                               void* lo, int64_t ld, int width, void* stream) {
  NB2_ENTER(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(x && hi && n > 0 && x_stride >= x_col0 + 3 && levels >= 0 && levels <= 16 && width >= 3 + 6 * levels && ld >= width,
                "encode_bf16: bad arguments");
  const size_t tile_bytes = (size_t)kEncPts * (width + 8) * 2 * sizeof(__nv_bfloat16);
  if ((width & 7) == 0 && (ld & 7) == 0 && ((uintptr_t)hi & 15) == 0 && (!lo || ((uintptr_t)lo & 15) == 0) && tile_bytes <= 48 * 1024) {
    encode3_bf16_kernel<<<grid_for(n, kEncPts), 3 * kEncPts, tile_bytes, (cudaStream_t)stream>>>(x, x_stride, x_col0, n, levels, normalize,
                                                                                              (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld, width);
  } else {
    const int blocks = (int)std::min<int64_t>(grid_for(n * width, 256), (int64_t)h->sm_count * 16);
    encode_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, x_stride, x_col0, n, levels, normalize, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld, width);
  }
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_weights_from_sigma_backward(nb2_handle* h, const float* sigma, const float* z, const float* dirs, int dir_stride,
                                               int64_t n_rays, int n_samples, int act, const float* g_weights, float* d_sigma, void* stream) {
  NB2_ENTER(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(sigma && z && g_weights && d_sigma && n_samples >= 1 && n_samples <= kTrMaxSamples, "weights_from_sigma_backward: n_samples must be in [1,%d]", kTrMaxSamples);
  NB2_CHECK_ARG((!dirs || dir_stride >= 3) && act >= 0 && act <= 2, "weights_from_sigma_backward: bad arguments");
  weights_backward_kernel<<<grid_for(n_rays, kTrWarps), 32 * kTrWarps, 0, (cudaStream_t)stream>>>(sigma, z, dirs, dir_stride, n_rays, n_samples, act,
                                                                                               g_weights, d_sigma);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_composite_backward(nb2_handle* h, const float* rgbo, const float* z, const float* dirs, int dir_stride, int64_t n_rays,
                                      int n_samples, int flags, const float* g_rgb, const float* g_weights, float* d_rgbo, void* stream) {
  NB2_ENTER(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(rgbo && z && dirs && g_rgb && d_rgbo && dir_stride >= 3, "composite_backward: null pointer");
  NB2_CHECK_ARG(n_samples >= 1 && n_samples <= kTrMaxSamples, "composite_backward: n_samples must be in [1,%d]", kTrMaxSamples);
  composite_backward_kernel<<<grid_for(n_rays, kTrWarps), 32 * kTrWarps, 0, (cudaStream_t)stream>>>(rgbo, z, dirs, dir_stride, n_rays, n_samples, flags,
                                                                                                 g_rgb, g_weights, d_rgbo);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_max_blur_backward(nb2_handle* h, const float* weights, const float* g_out, int64_t n_rays, int n_samples, float* d_weights,
                                     void* stream) {
  NB2_ENTER(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(weights && g_out && d_weights && n_samples >= 1, "max_blur_backward: bad arguments");
  max_blur_backward_kernel<<<grid_for(n_rays * n_samples, 256), 256, 0, (cudaStream_t)stream>>>(weights, g_out, n_rays, n_samples, d_weights);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_get_bounds_backward(nb2_handle* h, const int64_t* inds, const float* g_out, int64_t n_rays, int n_samples, int n_inds,
                                       float* d_weights, void* stream) {
  NB2_ENTER(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(inds && g_out && d_weights, "get_bounds_backward: null pointer");
  NB2_CHECK_ARG(n_samples >= 1 && n_samples <= kTrMaxSamples && n_inds >= 2 && n_inds <= kTrMaxSamples + 8, "get_bounds_backward: sizes out of range");
  get_bounds_backward_kernel<<<grid_for(n_rays, kTrWarps), 32 * kTrWarps, 0, (cudaStream_t)stream>>>(inds, g_out, n_rays, n_samples, n_inds, d_weights);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_nerf_head_backward(nb2_handle* h, const float* out, const float* g_out, int64_t n, void* dz_hi, void* dz_lo, void* ds_hi,
                                      void* ds_lo, void* stream) {
  NB2_ENTER(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(out && g_out && dz_hi && ds_hi && n > 0, "nerf_head_backward: null pointer");
  nerf_head_backward_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(out, g_out, n, (__nv_bfloat16*)dz_hi, (__nv_bfloat16*)dz_lo,
                                                                              (__nv_bfloat16*)ds_hi, (__nv_bfloat16*)ds_lo);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}
