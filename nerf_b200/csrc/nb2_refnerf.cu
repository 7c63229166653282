// nb2_refnerf.cu — the elementwise glue of RefNeRF.forward between its GEMMs (nerf/ref_model.py:78-118):
// head outputs -> normal / reflected direction / roughness, directional-MLP input columns, final colour composition.
// The MLP layers themselves run on the layer-wise tcgen05 GEMM (nb2_gemm.cu), the integrated directional encoding on
// ide_kernel (nb2_ops.cu).
#include "nb2_common.cuh"
#include "nb2_tc_ptx.cuh"

namespace nb2 {
using namespace ptx;

static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// heads (n, ld_h) = [normal(3), diffuse(3), tint(3), roughness_raw(1), density(1), ...]  (norm_col_tint_head | rho_tau_head)
//   roughness = softplus(rho - 1)                                   ref_model.py:83
//   normal    = -n / (||n|| + 1e-7)                                 ref_model.py:88
//   reflect   = d - 2 (d . normal) normal                           ref_model.py:91
//   nv_dot    = normal . d                                          ref_model.py:94
__global__ void ref_geometry_kernel(const float* __restrict__ heads, int ld_h, const float* __restrict__ dirs, int dir_stride, int64_t n,
                                    float* __restrict__ normal_out, float* __restrict__ reflect_out, float* __restrict__ rough_out,
                                    float* __restrict__ nv_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* hrow = heads + i * ld_h;
  const float nx = hrow[0], ny = hrow[1], nz = hrow[2];
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
  const float inv = __fadd_rn(nrm, 1e-7f);
  const float ax = __fdiv_rn(-nx, inv), ay = __fdiv_rn(-ny, inv), az = __fdiv_rn(-nz, inv);
  const float dx = dirs[i * dir_stride], dy = dirs[i * dir_stride + 1], dz = dirs[i * dir_stride + 2];
  const float dot = __fadd_rn(__fadd_rn(__fmul_rn(dx, ax), __fmul_rn(dy, ay)), __fmul_rn(dz, az));
  normal_out[i * 3] = ax; normal_out[i * 3 + 1] = ay; normal_out[i * 3 + 2] = az;
  const float two_dot = __fmul_rn(2.f, dot);
  reflect_out[i * 3] = __fsub_rn(dx, __fmul_rn(two_dot, ax));
  reflect_out[i * 3 + 1] = __fsub_rn(dy, __fmul_rn(two_dot, ay));
  reflect_out[i * 3 + 2] = __fsub_rn(dz, __fmul_rn(two_dot, az));
  rough_out[i] = softplus_f(__fsub_rn(hrow[9], 1.f));
  nv_out[i] = dot;
}

// columns [ide(w) | nv_dot | zero pad] of the directional MLP's input as bf16 hi / lo (ref_model.py:96: cat(bottleneck, ide, nv_dot))
__global__ void ref_dir_inputs_kernel(const float* __restrict__ ide, int w, const float* __restrict__ nv, int64_t n,
                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld, int width) {
  const int64_t total = n * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / width;
    const int c = (int)(i - r * width);
    const float v = c < w ? ide[r * w + c] : (c == w ? nv[r] : 0.f);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[r * ld + c] = h;
    if (lo) lo[r * ld + c] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
// vector form (width, ld multiples of 8; 16-byte aligned outputs): one thread per eight consecutive columns of a row -- one
// index division per eight values instead of one per value, 16-byte stores instead of 2-byte ones
__global__ void ref_dir_inputs_vec8_kernel(const float* __restrict__ ide, int w, const float* __restrict__ nv, int64_t n,
                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld, int width) {
  const int cpr = width >> 3;
  const int64_t total = n * cpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cpr;
    const int c0 = (int)(i - r * cpr) * 8;
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = c0 + 2 * j + e;
        v[e] = c < w ? __ldg(ide + r * w + c) : (c == w ? __ldg(nv + r) : 0.f);
      }
      ph[j] = ptx::pack_bf16x2(v[0], v[1]);
      pl[j] = ptx::residual16x2<false>(v[0], v[1], ph[j]);
    }
    *reinterpret_cast<uint4*>(hi + r * ld + c0) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (lo) *reinterpret_cast<uint4*>(lo + r * ld + c0) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
  }
}

// specular = spec * sigmoid(tint); diffuse = sigmoid(diffuse [- ln 3]); rgb = [linear_to_srgb](specular + diffuse)   ref_model.py:102-108
// out (n,4) = [rgb, density] (density passed through softplus(. + 0.5) when `shift_softplus`: nerf/procedures.py:74);
// ndot_out (n) = normal . cam_dir (nerf/nerf_base.py:111) when cam_dir is given.
__global__ void ref_color_kernel(const float* __restrict__ spec, const float* __restrict__ heads, int ld_h, int use_srgb, int shift_softplus,
                                 const float* __restrict__ normal, const float* __restrict__ cam_dir, int64_t n, float* __restrict__ out,
                                 float* __restrict__ ndot_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* hrow = heads + i * ld_h;
  const float eps = 1.1920928955078125e-07f;
  float rgb[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float tint = 1.f / (1.f + expf(-hrow[6 + k]));
    const float s = __fmul_rn(spec[i * 3 + k], tint);
    if (use_srgb) {
      const float dif = 1.f / (1.f + expf(-(hrow[3 + k] - 1.0986122886681098f)));
      const float v = __fadd_rn(s, dif);
      const float s0 = __fmul_rn(12.92f, v);
      const float s1 = __fdiv_rn(__fsub_rn(__fmul_rn(211.f, powf(fmaxf(eps, v), 0.41666666f)), 11.f), 200.f);
      rgb[k] = (v <= 0.0031308f) ? s0 : s1;
    } else {
      const float dif = 1.f / (1.f + expf(-hrow[3 + k]));
      rgb[k] = __fadd_rn(s, dif);
    }
  }
  float dens = hrow[10];
  if (shift_softplus) dens = softplus_f(__fadd_rn(dens, 0.5f));
  reinterpret_cast<float4*>(out)[i] = make_float4(rgb[0], rgb[1], rgb[2], dens);
  if (ndot_out != nullptr && cam_dir != nullptr)
    ndot_out[i] = __fadd_rn(__fadd_rn(__fmul_rn(normal[i * 3], cam_dir[0]), __fmul_rn(normal[i * 3 + 1], cam_dir[1])), __fmul_rn(normal[i * 3 + 2], cam_dir[2]));
}

// out[i] = a[i] . b   (normal @ cam_dir, nerf/nerf_base.py:111)
__global__ void dot3_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __fadd_rn(__fadd_rn(__fmul_rn(a[i * 3], b[0]), __fmul_rn(a[i * 3 + 1], b[1])), __fmul_rn(a[i * 3 + 2], b[2]));
}

// ---- backward of the glue (Ref-NeRF training, train.py:164-199 with is_ref_model) -------------------------------------------
// In the reference these are autograd's derivatives of the torch expressions of ref_model.py:81-105; restated per kernel.

// colour composition.  Forward: lin_k = spec_k sigmoid(tint_k) + sigmoid(diffuse_k [- ln 3]); rgb_k = [srgb](lin_k); density passes.
// spec is the ALREADY sigmoid-ed specular head (the GEMM epilogue applied it), so d(pre-activation) = d spec * spec (1 - spec).
//   ds (n,8) bf16 hi/lo: columns 0..2 = gradient of the specular head's pre-activation (wgrad / dgrad operand), 3..7 zero
//   d_heads (n, ld_dh) fp32: columns 3..5 diffuse, 6..8 tint, 10 density; every other column < ld_dh zero
//   (columns 0..2 and 9 are written afterwards by ref_geometry_backward_kernel)
__global__ void ref_color_backward_kernel(const float* __restrict__ spec, const float* __restrict__ heads, int ld_h, int use_srgb,
                                          const float* __restrict__ g_out, int64_t n, __nv_bfloat16* __restrict__ ds_hi,
                                          __nv_bfloat16* __restrict__ ds_lo, float* __restrict__ d_heads, int ld_dh) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* hrow = heads + i * ld_h;
  const float eps = 1.1920928955078125e-07f;
  const float4 g = reinterpret_cast<const float4*>(g_out)[i];
  const float gk[3] = {g.x, g.y, g.z};
  float dsl[3], ddif[3], dtint[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float tint = 1.f / (1.f + expf(-hrow[6 + k]));
    const float sp = spec[i * 3 + k];
    float dif, gv = gk[k];
    if (use_srgb) {
      dif = 1.f / (1.f + expf(-(hrow[3 + k] - 1.0986122886681098f)));
      const float v = __fadd_rn(__fmul_rn(sp, tint), dif);
      // d/dv of (v <= 0.0031308 ? 12.92 v : (211 max(eps, v)^(5/12) - 11) / 200)
      const float dv = (v <= 0.0031308f) ? 12.92f : (v > eps ? 1.055f * 0.41666666f * powf(v, 0.41666666f - 1.f) : 0.f);
      gv *= dv;
    } else {
      dif = 1.f / (1.f + expf(-hrow[3 + k]));
    }
    dsl[k] = gv * tint * sp * (1.f - sp);
    dtint[k] = gv * sp * tint * (1.f - tint);
    ddif[k] = gv * dif * (1.f - dif);
  }
  float* drow = d_heads + i * ld_dh;
  for (int c = 0; c < ld_dh; ++c) drow[c] = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) { drow[3 + k] = ddif[k]; drow[6 + k] = dtint[k]; }
  drow[10] = g.w;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float v = c < 3 ? dsl[c] : 0.f;
    const __nv_bfloat16 hb = __float2bfloat16_rn(v);
    ds_hi[i * 8 + c] = hb;
    if (ds_lo) ds_lo[i * 8 + c] = __float2bfloat16_rn(v - __bfloat162float(hb));
  }
}

// geometry + integrated directional encoding.  Forward (ref_geometry_kernel, ide_kernel): h = heads[0:3], s = |h| + 1e-7,
// nrm = -h / s, dot = d . nrm, r = d - 2 dot nrm, kinv = softplus(rho - 1), ide_i = (x + iy)^m P_i(z) exp(-l(l+1)/2 kinv) at r.
// Upstream: d_in row = gradient of the directional MLP's input, IDE columns at [ide_col0, ide_col0 + 2 n_pairs), nv_dot next;
// g_normal = gradient of the returned normal (or NULL).  Writes d_heads columns 0..2 (raw normal) and 9 (rho).
constexpr int kRgMaxPairs = 36;
constexpr int kRgMaxPow = 17;
__global__ void __launch_bounds__(128) ref_geometry_backward_kernel(const float* __restrict__ heads, int ld_h, const float* __restrict__ dirs,
                                                                    int dir_stride, int64_t n, const float* __restrict__ d_in, int64_t ld_in,
                                                                    int ide_col0, const float* __restrict__ g_normal,
                                                                    const float* __restrict__ mat, const int* __restrict__ ml, int n_pairs,
                                                                    int n_pow, float* __restrict__ d_heads, int ld_dh) {
  __shared__ float sh_mat[kRgMaxPow * kRgMaxPairs];
  __shared__ int sh_ml[2 * kRgMaxPairs];
  for (int t = threadIdx.x; t < n_pow * n_pairs; t += blockDim.x) sh_mat[t] = __ldg(mat + t);
  for (int t = threadIdx.x; t < 2 * n_pairs; t += blockDim.x) sh_ml[t] = __ldg(ml + t);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* hrow = heads + i * ld_h;
  const float hx = hrow[0], hy = hrow[1], hz = hrow[2];
  const float hn = sqrtf(hx * hx + hy * hy + hz * hz);
  const float s = hn + 1e-7f;
  const float ax = -hx / s, ay = -hy / s, az = -hz / s;
  const float dx = dirs[i * dir_stride], dy = dirs[i * dir_stride + 1], dz = dirs[i * dir_stride + 2];
  const float dot = dx * ax + dy * ay + dz * az;
  const float x = dx - 2.f * dot * ax, y = dy - 2.f * dot * ay, z = dz - 2.f * dot * az;   // reflected direction
  const float rho1 = hrow[9] - 1.f;
  const float kinv = softplus_f(rho1);
  float zp[kRgMaxPow], cr[kRgMaxPow], ci[kRgMaxPow];
  zp[0] = 1.f; cr[0] = 1.f; ci[0] = 0.f;
#pragma unroll
  for (int k = 1; k < kRgMaxPow; ++k) {
    zp[k] = zp[k - 1] * z;
    cr[k] = cr[k - 1] * x - ci[k - 1] * y;
    ci[k] = cr[k - 1] * y + ci[k - 1] * x;
  }
  const float* grow = d_in + i * ld_in + ide_col0;
  float gx = 0.f, gy = 0.f, gz = 0.f, gk = 0.f;
  for (int p = 0; p < n_pairs; ++p) {
    const int m = sh_ml[p], l = sh_ml[n_pairs + p];
    float acc = 0.f, dacc = 0.f;
#pragma unroll
    for (int k = 0; k < kRgMaxPow; ++k)
      if (k < n_pow) {
        const float c = sh_mat[k * n_pairs + p];
        acc = fmaf(zp[k], c, acc);
        if (k >= 1) dacc = fmaf((float)k * zp[k - 1], c, dacc);
      }
    const float sig = 0.5f * (float)(l * (l + 1));
    const float att = expf(-sig * kinv);
    float pr = 0.f, pi = 0.f, qr = 0.f, qi = 0.f;     // (x + iy)^m and (x + iy)^(m - 1)
#pragma unroll
    for (int k = 0; k < kRgMaxPow; ++k) {
      if (k == m) { pr = cr[k]; pi = ci[k]; }
      if (k == m - 1) { qr = cr[k]; qi = ci[k]; }
    }
    const float g_re = grow[p], g_im = grow[n_pairs + p];
    const float base = acc * att;
    gz += (g_re * pr + g_im * pi) * dacc * att;
    gx += (float)m * (g_re * qr + g_im * qi) * base;
    gy += (float)m * (g_im * qr - g_re * qi) * base;
    gk -= sig * (g_re * pr + g_im * pi) * base;
  }
  const float g_nv = grow[2 * n_pairs];
  // d nrm: returned normal, nv_dot = nrm . d, r = d - 2 (d . nrm) nrm
  const float gr_n = gx * ax + gy * ay + gz * az;
  float nx = g_nv * dx - 2.f * (gr_n * dx + dot * gx);
  float ny = g_nv * dy - 2.f * (gr_n * dy + dot * gy);
  float nz = g_nv * dz - 2.f * (gr_n * dz + dot * gz);
  if (g_normal != nullptr) { nx += g_normal[i * 3]; ny += g_normal[i * 3 + 1]; nz += g_normal[i * 3 + 2]; }
  // nrm = -h / (|h| + eps):  d h_b = -d nrm_b / s + (d nrm . h) h_b / (|h| s^2)
  const float nh = nx * hx + ny * hy + nz * hz;
  const float coef = hn > 0.f ? nh / (hn * s * s) : 0.f;
  float* drow = d_heads + i * ld_dh;
  drow[0] = -nx / s + coef * hx;
  drow[1] = -ny / s + coef * hy;
  drow[2] = -nz / s + coef * hz;
  drow[9] = gk / (1.f + expf(-rho1));     // d softplus
}

// positional encoding.  Forward row = [x, sin(2^l x), cos(2^l x) ...] (3-wide terms); d x_k = d row_k + sum_l 2^l (cos(2^l x_k)
// d sin_lk - sin(2^l x_k) d cos_lk).  d_enc (n, ld) fp32, d_x (n, 3).
__global__ void encode_backward_kernel(const float* __restrict__ x, int x_stride, int x_col0, int64_t n, int levels,
                                       const float* __restrict__ d_enc, int64_t ld, float* __restrict__ d_x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  const int64_t r = i / 3;
  const int k = (int)(i - r * 3);
  const float xv = __ldg(x + r * x_stride + x_col0 + k);
  const float* g = d_enc + r * ld;
  float acc = g[k];
  for (int l = 0; l < levels; ++l) {
    const float f = exp2f((float)l);
    float sn, cs;
    sincos_any(xv * f, sn, cs);
    acc += f * (cs * g[3 + 6 * l + k] - sn * g[3 + 6 * l + 3 + k]);
  }
  d_x[i] = acc;
}

}  // namespace nb2

using namespace nb2;

extern "C" int nb2_ref_color_backward(nb2_handle* h, const float* spec, const float* heads, int ld_heads, int use_srgb, const float* g_out,
                                      int64_t n, void* ds_hi, void* ds_lo, float* d_heads, int ld_dheads, void* stream) {
  NB2_ENTER(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(spec && heads && g_out && ds_hi && d_heads && ld_heads >= 11 && ld_dheads >= 11, "ref_color_backward: bad arguments");
  ref_color_backward_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(spec, heads, ld_heads, use_srgb, g_out, n, (__nv_bfloat16*)ds_hi,
                                                                                (__nv_bfloat16*)ds_lo, d_heads, ld_dheads);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_ref_geometry_backward(nb2_handle* h, const float* heads, int ld_heads, const float* dirs, int dir_stride, int64_t n,
                                         const float* d_in, int64_t ld_in, int ide_col0, const float* g_normal, const float* mat, const int* ml,
                                         int n_pairs, int n_pow, float* d_heads, int ld_dheads, void* stream) {
  NB2_ENTER(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(heads && dirs && d_in && mat && ml && d_heads && ld_heads >= 11 && ld_dheads >= 11 && dir_stride >= 3, "ref_geometry_backward: bad arguments");
  NB2_CHECK_ARG(n_pairs >= 1 && n_pairs <= kRgMaxPairs && n_pow >= 1 && n_pow <= kRgMaxPow && ide_col0 >= 0 && ld_in >= ide_col0 + 2 * n_pairs + 1,
                "ref_geometry_backward: at most %d (m, l) pairs, degree %d; d_in rows must hold the IDE and nv_dot columns", kRgMaxPairs, kRgMaxPow - 1);
  ref_geometry_backward_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(heads, ld_heads, dirs, dir_stride, n, d_in, ld_in, ide_col0, g_normal,
                                                                                   mat, ml, n_pairs, n_pow, d_heads, ld_dheads);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_encode_backward(nb2_handle* h, const float* x, int x_stride, int x_col0, int64_t n, int levels, const float* d_enc, int64_t ld,
                                   float* d_x, void* stream) {
  NB2_ENTER(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(x && d_enc && d_x && x_stride >= x_col0 + 3 && levels >= 0 && levels <= 16 && ld >= 3 + 6 * levels, "encode_backward: bad arguments");
  encode_backward_kernel<<<grid_for(n * 3, 256), 256, 0, (cudaStream_t)stream>>>(x, x_stride, x_col0, n, levels, d_enc, ld, d_x);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_dot3(nb2_handle* h, const float* a, const float* b, int64_t n, float* out, void* stream) {
  NB2_ENTER(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(a && b && out, "dot3: null pointer");
  dot3_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_ref_geometry(nb2_handle* h, const float* heads, int ld_heads, const float* dirs, int dir_stride, int64_t n, float* normal_out,
                                float* reflect_out, float* rough_out, float* nv_out, void* stream) {
  NB2_ENTER(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(heads && dirs && normal_out && reflect_out && rough_out && nv_out && ld_heads >= 11 && dir_stride >= 3, "ref_geometry: bad arguments");
  ref_geometry_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(heads, ld_heads, dirs, dir_stride, n, normal_out, reflect_out, rough_out, nv_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_ref_dir_inputs(nb2_handle* h, const float* ide, int ide_width, const float* nv_dot, int64_t n, void* hi, void* lo, int64_t ld,
                                  int width, void* stream) {
  NB2_ENTER(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(ide && nv_dot && hi && ide_width >= 1 && width >= ide_width + 1 && ld >= width, "ref_dir_inputs: bad arguments");
  if ((width & 7) == 0 && (ld & 7) == 0 && ((uintptr_t)hi & 15) == 0 && (!lo || ((uintptr_t)lo & 15) == 0)) {
    const int blocks = (int)std::min<int64_t>(grid_for(n * (width >> 3), 256), (int64_t)h->sm_count * 16);
    ref_dir_inputs_vec8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ide, ide_width, nv_dot, n, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld, width);
  } else {
    const int blocks = (int)std::min<int64_t>(grid_for(n * width, 256), (int64_t)h->sm_count * 16);
    ref_dir_inputs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ide, ide_width, nv_dot, n, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld, width);
  }
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_ref_color(nb2_handle* h, const float* spec, const float* heads, int ld_heads, int use_srgb, int shift_softplus, const float* normal,
                             const float* cam_dir, int64_t n, float* out, float* ndot_out, void* stream) {
  NB2_ENTER(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(spec && heads && out && ld_heads >= 11, "ref_color: bad arguments");
  NB2_CHECK_ARG(!ndot_out || (normal && cam_dir), "ref_color: ndot_out needs normal and cam_dir");
  ref_color_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(spec, heads, ld_heads, use_srgb, shift_softplus, normal, cam_dir, n, out, ndot_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}
