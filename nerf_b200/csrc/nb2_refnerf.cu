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

}  // namespace nb2

using namespace nb2;

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
  const int blocks = (int)std::min<int64_t>(grid_for(n * width, 256), (int64_t)h->sm_count * 16);
  ref_dir_inputs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ide, ide_width, nv_dot, n, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld, width);
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
