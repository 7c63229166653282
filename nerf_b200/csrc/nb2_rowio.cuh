// nb2_rowio.cuh — how an MLP row (one sample on one ray) gets its inputs.  Shared by the
// CUDA-core and the tcgen05 MLP kernels so both see bit-identical sample points (the tensor kernels
// evaluate the encoding with their own branch-free sin/cos, nb2_common.cuh:enc_sincos, within 1e-7).
#pragma once
#include "nb2_common.cuh"

namespace nb2 {

struct RowIn {
  float p[3];   // sample position
  float d[3];   // ray direction (un-normalised), zero if the network takes none
  float z;      // depth along the ray (0 for explicit points)
  int64_t ray;  // ray index (row / P), or the row itself for explicit points
  int s;        // sample index on the ray
  const float* enc;  // externally encoded position features of this row, or nullptr
  float cov[3];      // ipe: diagonal covariance of the conical frustum (p then holds its mean)
  bool ipe;
  bool valid;
};

__device__ __forceinline__ RowIn load_row(const MlpIo& io, int64_t row) {
  RowIn r;
  r.valid = row < io.n_rows;
  r.z = 0.f;
  r.ray = row;
  r.s = 0;
  r.enc = nullptr;
  r.ipe = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.p[k] = 0.f; r.d[k] = 0.f; r.cov[k] = 0.f; }
  if (!r.valid) return r;
  if (io.in_mode == 0) {
    const float* src = io.pts + row * io.pts_stride;
    if (io.enc) r.enc = io.enc + row * (int64_t)(6 * io.P);   // io.P carries pos_levels in this mode (see nb2_api.cu)
#pragma unroll
    for (int k = 0; k < 3; ++k) r.p[k] = __ldg(src + k);
    if (io.pts_stride >= 6) {
#pragma unroll
      for (int k = 0; k < 3; ++k) r.d[k] = __ldg(src + 3 + k);
    }
    return r;
  }
  r.ray = row / io.P;
  r.s = (int)(row - r.ray * io.P);
  const float* ray = io.rays + r.ray * 6;
  if (io.in_mode == 1) {
    r.z = __ldg(io.z + row);
  } else {
    // z = linspace[s] + U[0,1) * resolution       /root/reference/nerf/procedures.py:65
    float j = io.jitter ? __ldg(io.jitter + row)
                        : philox_uniform(io.seed, (uint64_t)(io.ray_offset + r.ray), (uint32_t)r.s, 0u);
    r.z = __fadd_rn(__ldg(io.base_z + r.s), __fmul_rn(j, io.resolution));
    if (io.z_out) io.z_out[row] = r.z;
  }
  if (io.in_mode == 2 && io.ipe) {
    // conical frustum [z_s, z_{s+1}): coneParameters / coneMeanCov (mip_methods.py:15-33) in the arithmetic of ipe_kernel
    float z1;
    if (r.s + 1 < io.P) {
      const float j1 = io.jitter ? __ldg(io.jitter + row + 1)
                                 : philox_uniform(io.seed, (uint64_t)(io.ray_offset + r.ray), (uint32_t)(r.s + 1), 0u);
      z1 = __fadd_rn(__ldg(io.base_z + r.s + 1), __fmul_rn(j1, io.resolution));
    } else {
      z1 = __fadd_rn(r.z, io.ipe_last_step);
    }
    const float z0 = r.z;
    const float mid = (z1 + z0) / 2.f, hw = (z1 - z0) / 2.f, diff = hw * hw;
    const float tmp1 = 3.f * mid * mid + diff;
    const float mu_t = mid + 2.f * mid * diff / tmp1;
    const float sigma_t2 = diff / 3.f - 4.f * (diff * diff) * (12.f * mid * mid - diff) / 15.f / (tmp1 * tmp1);
    const float sigma_r2 = (io.ipe_radius * io.ipe_radius) * (0.25f * mid * mid + 5.f / 12.f * diff - 4.f * diff * diff / (15.f * tmp1));
    const float gnorm = (float)sqrt(*io.ipe_sumsq);
    r.ipe = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      r.d[k] = __ldg(ray + 3 + k);
      r.p[k] = __ldg(ray + k) + mu_t * r.d[k];
      const float dd = r.d[k] * r.d[k];
      r.cov[k] = sigma_t2 * dd + sigma_r2 * (1.f - dd / gnorm);
    }
    return r;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r.d[k] = __ldg(ray + 3 + k);
    r.p[k] = __fadd_rn(__ldg(ray + k), __fmul_rn(r.d[k], r.z));  // o + z*d, nerf_base.py:54
  }
  return r;
}

// rotation = d / ||d||                             /root/reference/nerf/mip_model.py:44-45
__device__ __forceinline__ void normalize_dir(const float d[3], float out[3]) {
  float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
  for (int k = 0; k < 3; ++k) out[k] = __fdiv_rn(d[k], n);
}

// Column c of the encoded vector [x(3), sin(2^0 x)(3), cos(2^0 x)(3), sin(2^1 x)(3), ...]
// (cat_origin = True; nerf_helper.py:38-48 + mip_model.py:50-52).  Returns 0 past 3 + 6*levels.
__device__ __forceinline__ float enc_column(const float x[3], int c, int levels, const float* ext = nullptr, const float* cov = nullptr) {
  if (c < 3) return x[c];
  int j = c - 3;
  int l = j / 6, w = j % 6;
  if (l >= levels) return 0.f;
  if (ext) return __ldg(ext + j);
  const float sc = exp2f((float)l);
  float a = x[w % 3] * sc;
  const float v = (w < 3) ? sinf(a) : cosf(a);
  // integrated positional encoding: the level's features are attenuated by exp(-0.5 * 4^l * var)   mip_methods.py:36-58
  return cov ? v * expf(-0.5f * (sc * sc * cov[w % 3])) : v;
}

}  // namespace nb2
