// nb2_ops.cu — the HBM-bound stages of the ray-marching path: ray generation, stratified
// sampling, sinusoidal / integrated positional encoding, density->weights, max-blur,
// inverse-CDF resampling (search + sort), coarse/fine merge, alpha compositing.
//
// All kernels are coalesced streaming kernels: one thread per output element for the
// elementwise stages, one warp per ray for the scan / search / sort stages (row data staged in
// shared memory, transmittance via a warp exclusive product scan).
#include "nb2_common.cuh"

namespace nb2 {

static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// ------------------------------------------------------------------------------------------
// a1  ray generation                                   /root/reference/nerf/procedures.py:43-51
//   coords = (col - W/2 + 0.5, H/2 - row + 0.5) / focal ; d_i = sum_j coords_j * R_ij, c_2 = -1
// ------------------------------------------------------------------------------------------
__global__ void generate_rays_kernel(const float* __restrict__ pose, int H, int W, float fx,
                                     float fy, int64_t pix_offset, int64_t n,
                                     float* __restrict__ rays) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t pix = pix_offset + i;
  int row = (int)(pix / W), col = (int)(pix % W);
  // the reference builds integer-valued float grids, subtracts the half extents, adds 0.5
  float cx = __fadd_rn(__fsub_rn((float)col, (float)W / 2.f), 0.5f);
  float cy = __fadd_rn(__fsub_rn((float)H / 2.f, (float)row), 0.5f);
  cx = __fdiv_rn(cx, fx);
  cy = __fdiv_rn(cy, fy);
  float* o = rays + i * 6;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float r0 = __ldg(pose + k * 4 + 0), r1 = __ldg(pose + k * 4 + 1), r2 = __ldg(pose + k * 4 + 2);
    o[k] = __ldg(pose + k * 4 + 3);
    // torch.sum over the last (size-3) axis: ((c0*r0 + c1*r1) + c2*r2), products rounded first
    o[3 + k] = __fadd_rn(__fadd_rn(__fmul_rn(cx, r0), __fmul_rn(cy, r1)), __fmul_rn(-1.f, r2));
  }
}

// ------------------------------------------------------------------------------------------
// a3+a4  stratified coarse depths and points        /root/reference/nerf/procedures.py:65-66
// ------------------------------------------------------------------------------------------
__global__ void sample_coarse_kernel(const float* __restrict__ rays, const float* __restrict__ base_z,
                                     const float* __restrict__ jitter, float resolution,
                                     uint64_t seed, int64_t ray_offset, int64_t n_rays, int P,
                                     float* __restrict__ z_out, float* __restrict__ pts_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * P) return;
  int64_t r = i / P;
  int s = (int)(i % P);
  float j = jitter ? jitter[i] : philox_uniform(seed, (uint64_t)(ray_offset + r), (uint32_t)s, 0u);
  float z = __fadd_rn(__ldg(base_z + s), __fmul_rn(j, resolution));
  z_out[i] = z;
  if (pts_out) {
    const float* ray = rays + r * 6;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      pts_out[i * 3 + k] = __fadd_rn(__ldg(ray + k), __fmul_rn(z, __ldg(ray + 3 + k)));
  }
}

// Vectorised form (P % 4 == 0, 16-byte aligned pointers): one thread = four consecutive samples of a ray, so every global
// access is a 16-byte vector and a warp covers whole 128-byte lines (the scalar form stores pts with a 12-byte stride).
__global__ void __launch_bounds__(256)
sample_coarse_vec4_kernel(const float* __restrict__ rays, const float* __restrict__ base_z,
                          const float* __restrict__ jitter, float resolution, uint64_t seed, int64_t ray_offset,
                          int64_t n_rays, int P, float* __restrict__ z_out, float* __restrict__ pts_out) {
  __shared__ float4 slab[8][3 * 32];          // a warp's 32 x 48 point bytes, written back as three coalesced 512-byte rows
  const int P4 = P >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total = n_rays * P4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i_warp = i - lane;
  if (i_warp >= total) return;                 // whole warps only: the slab hand-over below is warp-synchronous
  if (i < total) {
    const int64_t r = i / P4;
    const int s0 = (int)(i - r * P4) * 4;
    float j[4];
    if (jitter) {
      const float4 jv = __ldg(reinterpret_cast<const float4*>(jitter) + i);
      j[0] = jv.x; j[1] = jv.y; j[2] = jv.z; j[3] = jv.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) j[k] = philox_uniform(seed, (uint64_t)(ray_offset + r), (uint32_t)(s0 + k), 0u);
    }
    const float4 bz = __ldg(reinterpret_cast<const float4*>(base_z + s0));
    const float b[4] = {bz.x, bz.y, bz.z, bz.w};
    float z[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) z[k] = __fadd_rn(b[k], __fmul_rn(j[k], resolution));
    reinterpret_cast<float4*>(z_out)[i] = make_float4(z[0], z[1], z[2], z[3]);
    if (pts_out) {
      const float* ray = rays + r * 6;
      const float o[3] = {__ldg(ray), __ldg(ray + 1), __ldg(ray + 2)}, d[3] = {__ldg(ray + 3), __ldg(ray + 4), __ldg(ray + 5)};
      float v[12];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) v[3 * k + c] = __fadd_rn(o[c], __fmul_rn(z[k], d[c]));
      slab[warp][lane * 3 + 0] = make_float4(v[0], v[1], v[2], v[3]);
      slab[warp][lane * 3 + 1] = make_float4(v[4], v[5], v[6], v[7]);
      slab[warp][lane * 3 + 2] = make_float4(v[8], v[9], v[10], v[11]);
    }
  }
  if (pts_out) {
    __syncwarp();
    const int n_valid = (int)min((int64_t)32, total - i_warp) * 3;      // float4s this warp owns
    float4* dst = reinterpret_cast<float4*>(pts_out) + i_warp * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int q = k * 32 + lane;
      if (q < n_valid) dst[q] = slab[warp][q];
    }
  }
}

// ------------------------------------------------------------------------------------------
// a5  sinusoidal positional encoding                /root/reference/nerf/nerf_helper.py:38-48
//   out[p, 6l + c] = sin(2^l x[p,c]), out[p, 6l + 3 + c] = cos(2^l x[p,c])
// One block = 128 consecutive points; (point, level, component) triples are spread over the
// threads, the 128 x 6L tile is assembled in shared memory and written back with float4 stores.
// ------------------------------------------------------------------------------------------
constexpr int kPeBlockPts = 128;
__global__ void __launch_bounds__(256) posenc_kernel(const float* __restrict__ x, int64_t n, int dims,
                                                     int levels, float* __restrict__ out) {
  extern __shared__ float pe_tile[];  // [kPeBlockPts][2*dims*levels]
  const int width = 2 * dims * levels;
  const int64_t p0 = (int64_t)blockIdx.x * kPeBlockPts;
  const int npts = (int)min((int64_t)kPeBlockPts, n - p0);
  const int per_pt = dims * levels;
  for (int t = threadIdx.x; t < npts * per_pt; t += blockDim.x) {
    int p = t / per_pt, j = t % per_pt;
    int l = j / dims, c = j % dims;
    float v = __ldg(x + (p0 + p) * dims + c) * exp2f((float)l);  // exact power-of-two scale
    float s, co;
    sincos_any(v, s, co);
    pe_tile[p * width + 2 * dims * l + c] = s;
    pe_tile[p * width + 2 * dims * l + dims + c] = co;
  }
  __syncthreads();
  float* dst = out + p0 * width;
  const int total = npts * width;
  if ((total & 3) == 0 && ((((uintptr_t)dst) & 15) == 0)) {
    for (int t = threadIdx.x; t < total / 4; t += blockDim.x)
      reinterpret_cast<float4*>(dst)[t] = reinterpret_cast<const float4*>(pe_tile)[t];
  } else {
    for (int t = threadIdx.x; t < total; t += blockDim.x) dst[t] = pe_tile[t];
  }
}

// dims == 3 (positions, directions): one thread per (point, component) runs down the levels with the exact doubling
// v <- 2 v (= x * 2^l bit for bit), so a block's input is 384 consecutive floats and there is no per-element index
// arithmetic; the generic kernel above spends ~4x the sin/cos cost on runtime divisions (0.29 of the copy bandwidth).
constexpr int kPe3Pts = 128;
__global__ void __launch_bounds__(3 * kPe3Pts) posenc3_kernel(const float* __restrict__ x, int64_t n, int levels, float* __restrict__ out) {
  extern __shared__ float pe_tile[];  // [kPe3Pts][6 * levels]
  const int width = 6 * levels;
  const int64_t p0 = (int64_t)blockIdx.x * kPe3Pts;
  const int npts = (int)min((int64_t)kPe3Pts, n - p0);
  const int t = threadIdx.x;
  if (t < 3 * npts) {
    const int p = t / 3, c = t - 3 * p;
    float v = __ldg(x + p0 * 3 + t);
    float* row = pe_tile + p * width + c;
    // one range test per thread instead of one per level: the branch-free reduction covers |2^(L-1) x| <= 1e5
    if (fabsf(v) * exp2f((float)(levels - 1)) <= 1.0e5f) {
      for (int l = 0; l < levels; ++l) {
        float sn, cs;
        enc_sincos(v, sn, cs);
        row[6 * l] = sn;
        row[6 * l + 3] = cs;
        v = __fmul_rn(v, 2.f);
      }
    } else {
      for (int l = 0; l < levels; ++l) {
        float sn, cs;
        sincos_any(v, sn, cs);
        row[6 * l] = sn;
        row[6 * l + 3] = cs;
        v = __fmul_rn(v, 2.f);
      }
    }
  }
  __syncthreads();
  float* dst = out + p0 * width;
  const int total = npts * width;
  if ((total & 3) == 0 && ((((uintptr_t)dst) & 15) == 0)) {
    for (int q = t; q < total / 4; q += 3 * kPe3Pts) reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(pe_tile)[q];
  } else {
    for (int q = t; q < total; q += 3 * kPe3Pts) dst[q] = pe_tile[q];
  }
}

// ------------------------------------------------------------------------------------------
// a14  integrated positional encoding              /root/reference/nerf/mip_methods.py:15-58
// ------------------------------------------------------------------------------------------
__global__ void ipe_sumsq_kernel(const float* __restrict__ rays, int64_t n_rays, double* __restrict__ acc) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rays; i += (int64_t)gridDim.x * blockDim.x) {
    float a = rays[i * 6 + 3], b = rays[i * 6 + 4], c = rays[i * 6 + 5];
    s += (double)a * a + (double)b * b + (double)c * c;
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}

// One thread per cone; a block's 128 x 6L feature tile is assembled in shared memory (row pitch 6L + 1: conflict-free for
// the per-thread writes) and written back with coalesced stores -- the direct form (60 scalar stores per thread at a
// 240-byte thread stride) ran at 0.10 of the copy bandwidth.
constexpr int kIpeBlock = 128;
__global__ void __launch_bounds__(kIpeBlock)
ipe_kernel(const float* __restrict__ zvals, const float* __restrict__ rays, int64_t n_rays, int C, int L, float radius,
           const double* __restrict__ sumsq, float* __restrict__ feat, float* __restrict__ mu_out, float* __restrict__ mu_t_out) {
  extern __shared__ float ipe_tile[];          // [kIpeBlock][pitch]
  // even L: rows packed (pitch = 6 L, a multiple of four floats) so the tile leaves as 16-byte vectors -- the per-thread
  // writes are then 4-way bank-conflicted (row stride 60 words), which costs far less than the scalar row copy did (ncu:
  // 1.4 k of the 3.6 k warp instructions per 32 cones); odd L keeps the conflict-free pitch 6 L + 1 and scalar rows
  const int width = 6 * L, pitch = (L & 1) ? width + 1 : width;
  const int64_t i0 = (int64_t)blockIdx.x * kIpeBlock;
  const int64_t i = i0 + threadIdx.x;
  const int64_t total = n_rays * C;
  if (i < total) {
    const int64_t r = i / C;
    const int c = (int)(i % C);
    const float gnorm = (float)sqrt(*sumsq);  // batch-global ||d||_F  (mip_methods.py:31)
    float z0 = zvals[r * (C + 1) + c], z1 = zvals[r * (C + 1) + c + 1];
    // coneParameters (mip_methods.py:15-23)
    float mid = (z1 + z0) / 2.f;
    float hw = (z1 - z0) / 2.f;
    float diff = hw * hw;
    float tmp1 = 3.f * mid * mid + diff;
    float mu_t = mid + 2.f * mid * diff / tmp1;
    float sigma_t2 = diff / 3.f - 4.f * (diff * diff) * (12.f * mid * mid - diff) / 15.f / (tmp1 * tmp1);
    float sigma_r2 = (radius * radius) * (0.25f * mid * mid + 5.f / 12.f * diff - 4.f * diff * diff / (15.f * tmp1));
    if (mu_t_out) mu_t_out[i] = mu_t;
    const float* ray = rays + r * 6;
    float* f = ipe_tile + threadIdx.x * pitch;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float o = __ldg(ray + k), d = __ldg(ray + 3 + k);
      float mu = o + mu_t * d;                       // coneMeanCov (mip_methods.py:27-33)
      float dd = d * d;
      float diag = sigma_t2 * dd + sigma_r2 * (1.f - dd / gnorm);
      if (mu_out) mu_out[i * 3 + k] = mu;
      float scale = 1.f;
      const bool small = fabsf(mu) * exp2f((float)(L - 1)) <= 1.0e5f && diag >= 0.f;   // one range test per component, not per level
      if (small) {
        for (int l = 0; l < L; ++l) {
          const float damp = __expf(-0.5f * (scale * scale * diag));
          float s, co;
          enc_sincos(scale * mu, s, co);
          f[6 * l + k] = s * damp;
          f[6 * l + 3 + k] = co * damp;
          scale *= 2.f;
        }
        continue;
      }
      for (int l = 0; l < L; ++l) {                   // multFreq + ipe_feature (mip_methods.py:36-58)
        // ex2.approx (2 ulp) behind one rounded multiply: relative error <= 2e-7 (1 + |x|), absolute <= 3e-7 on a factor in (0, 1]
        // (a negative variance -- the reference's batch-global norm allows it -- takes the library path)
        const float ex = -0.5f * (scale * scale * diag);
        float damp = ex <= 0.f ? __expf(ex) : expf(ex);
        float s, co;
        sincos_any(scale * mu, s, co);
        f[6 * l + k] = s * damp;
        f[6 * l + 3 + k] = co * damp;
        scale *= 2.f;
      }
    }
  }
  __syncthreads();
  const int ncones = (int)min((int64_t)kIpeBlock, total - i0);
  // one warp per cone row (6 L contiguous floats; the rows of a block are contiguous too): no index division per element
  float* dst = feat + i0 * width;
  if (pitch == width && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    const int n4 = ncones * width / 4;
    for (int q = threadIdx.x; q < n4; q += kIpeBlock) reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(ipe_tile)[q];
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cn = warp; cn < ncones; cn += kIpeBlock / 32) {
    const float* src = ipe_tile + cn * pitch;
    float* row = dst + (size_t)cn * width;
    for (int j = lane; j < width; j += 32) row[j] = src[j];
  }
}

// ------------------------------------------------------------------------------------------
// shared warp-per-ray machinery
// ------------------------------------------------------------------------------------------
constexpr int kMaxSamples = 256;  // per-ray sample count supported by the warp kernels
constexpr int kWarpsPerBlock = 8;

// ---- register form: a group of `lpr` lanes (a power of two) owns one ray and every lane 4 * S4 CONSECUTIVE samples --------
// One 16-byte load per four samples, the products of a lane's own samples taken serially, ONE log2(lpr)-step scan across the
// group instead of one 5-step scan per 32 samples, no shared memory.  At 64 samples a warp carries two rays, at 32 four.
// (The warp-per-ray kernels above stay as the fallback for sample counts that are not a multiple of four / unaligned rows.)
template <int S4>
__device__ __forceinline__ void ray_weights_regs(const float (&dep)[4 * S4], const float (&sig)[4 * S4], int s0, int P, int act,
                                                 int lpr, int sub, float (&w)[4 * S4]) {
  constexpr int S = 4 * S4;
  const float nxt = __shfl_down_sync(0xffffffffu, dep[0], 1, lpr);   // the next lane's first depth (unused by a ray's last sample)
  float alpha[S], pre[S];                                            // pre[j] = product of this lane's factors 0..j
#pragma unroll
  for (int j = 0; j < S; ++j) {
    const int i = s0 + j;
    float m = 1.f, f = 1.f;
    alpha[j] = 0.f;
    if (i < P) {
      const float dn = (j + 1 < S) ? dep[(j + 1) % S] : nxt;
      const float delta = (i + 1 < P) ? __fsub_rn(dn, dep[j]) : 1e10f;
      m = expf(-apply_density_act(sig[j], act) * delta);
      alpha[j] = 1.f - m;
      f = m + 1e-10f;
    }
    pre[j] = (j == 0) ? f : pre[(j + S - 1) % S] * f;
  }
  float inc = pre[S - 1];
  for (int d = 1; d < lpr; d <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, inc, d, lpr);
    if (sub >= d) inc *= o;
  }
  float exc = __shfl_up_sync(0xffffffffu, inc, 1, lpr);
  if (sub == 0) exc = 1.f;
#pragma unroll
  for (int j = 0; j < S; ++j) w[j] = alpha[j] * (j == 0 ? exc : exc * pre[(j + S - 1) % S]);
}

// weights w_i = (1 - m_i) * prod_{j<i} (m_j + 1e-10), m_i = exp(-act(sigma_i) * delta_i),
// delta_i = depth_{i+1} - depth_i (last = 1e10), depth = z * ||d||   (nerf_base.py:79-86)
// sh_z holds the (already scaled) depths, sh_s the densities; weights are written to sh_w.
__device__ __forceinline__ void ray_weights_warp(const float* sh_depth, const float* sh_sigma,
                                                 float* sh_w, int P, int act, int lane) {
  if ((P & 3) == 0) {
    // the register form's association order (serial inside groups of 4 * S4 consecutive samples, one scan across the groups),
    // so that the fused resample kernel, the standalone kernels and their fallbacks produce the same bits
    const int s4 = P <= 128 ? 1 : 2;
    const int need = (P + 4 * s4 - 1) / (4 * s4);
    int lpr = 1;
    while (lpr < need) lpr <<= 1;
    const int sub = lane & (lpr - 1), s0 = sub * 4 * s4;      // lanes >= lpr repeat the work of lane - lpr and write nothing
    if (s4 == 1) {
      float dep[4], sig[4], w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { dep[j] = s0 + j < P ? sh_depth[s0 + j] : 0.f; sig[j] = s0 + j < P ? sh_sigma[s0 + j] : 0.f; }
      ray_weights_regs<1>(dep, sig, s0, P, act, lpr, sub, w);
      __syncwarp();                                             // sh_w may alias an input row
      if (lane < lpr && s0 < P) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sh_w[s0 + j] = w[j];
      }
    } else {
      float dep[8], sig[8], w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { dep[j] = s0 + j < P ? sh_depth[s0 + j] : 0.f; sig[j] = s0 + j < P ? sh_sigma[s0 + j] : 0.f; }
      ray_weights_regs<2>(dep, sig, s0, P, act, lpr, sub, w);
      __syncwarp();
      if (lane < lpr) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (s0 + j < P) sh_w[s0 + j] = w[j];
      }
    }
    __syncwarp();
    return;
  }
  float carry = 1.f;
  for (int base = 0; base < P; base += 32) {
    int i = base + lane;
    float m = 1.f, alpha = 0.f;
    if (i < P) {
      float delta = (i + 1 < P) ? __fsub_rn(sh_depth[i + 1], sh_depth[i]) : 1e10f;
      m = expf(-apply_density_act(sh_sigma[i], act) * delta);
      alpha = 1.f - m;
    }
    float f = (i < P) ? (m + 1e-10f) : 1.f;
    float inc = warp_scan_mul(f, lane);
    float exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 1.f;
    if (i < P) sh_w[i] = alpha * (carry * exc);
    carry *= __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
}

__device__ __forceinline__ float dir_norm(const float* d) {
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
}

// a7  ProposalNetwork.get_weights / NeRF.getNormedWeight
// One warp per ray, warps loop over rays (persistent grid) and keep the NEXT ray's rows in flight in registers while the
// current ray is scanned out of shared memory: a warp-per-ray kernel is latency-bound otherwise (two 256-byte rows per
// warp in flight = 0.30 of the copy bandwidth at 64 samples).
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
weights_kernel(const float* __restrict__ sigma, const float* __restrict__ z, const float* __restrict__ dirs,
               int dir_stride, int64_t n_rays, int P, int act, float* __restrict__ w_out) {
  __shared__ float sh[kWarpsPerBlock][3][kMaxSamples];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * kWarpsPerBlock;
  int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  constexpr int kPer = kMaxSamples / 32;
  float zr[kPer], sr[kPer];
  float nrm = 1.f;
  auto fetch = [&](int64_t ray) {
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = lane + 32 * q;
      if (i < P) { zr[q] = __ldg(z + ray * P + i); sr[q] = __ldg(sigma + ray * P + i); }
    }
    nrm = dirs ? dir_norm(dirs + ray * dir_stride) : 1.f;
  };
  if (r < n_rays) fetch(r);
  for (; r < n_rays; r += stride) {
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = lane + 32 * q;
      if (i < P) { sh[warp][0][i] = dirs ? __fmul_rn(zr[q], nrm) : zr[q]; sh[warp][1][i] = sr[q]; }
    }
    if (r + stride < n_rays) fetch(r + stride);          // in flight during the scan below
    __syncwarp();
    ray_weights_warp(sh[warp][0], sh[warp][1], sh[warp][2], P, act, lane);
    for (int i = lane; i < P; i += 32) w_out[r * P + i] = sh[warp][2][i];
    __syncwarp();
  }
}

// lanes per ray for the register kernels: the power of two >= ceil(P / (4 * S4)); S4 = 1 up to 128 samples, 2 up to 256
static inline void regs_geometry(int P, int& s4, int& lpr) {
  s4 = P <= 128 ? 1 : 2;
  const int need = (P + 4 * s4 - 1) / (4 * s4);
  lpr = 1;
  while (lpr < need) lpr <<= 1;
}
static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

template <int S4>
__global__ void __launch_bounds__(256)
weights_regs_kernel(const float* __restrict__ sigma, const float* __restrict__ z, const float* __restrict__ dirs, int dir_stride,
                    int64_t n_rays, int P, int act, int lpr, float* __restrict__ w_out) {
  constexpr int S = 4 * S4;
  const int lane = threadIdx.x & 31, sub = lane & (lpr - 1);
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t r = warp_id * (32 / lpr) + lane / lpr;
  const int s0 = sub * S;
  const bool rv = r < n_rays;
  float dep[S], sig[S], w[S];
#pragma unroll
  for (int g = 0; g < S4; ++g) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (rv && s0 + 4 * g < P) {
      a = __ldg(reinterpret_cast<const float4*>(z + r * P + s0 + 4 * g));
      b = __ldg(reinterpret_cast<const float4*>(sigma + r * P + s0 + 4 * g));
    }
    dep[4 * g] = a.x; dep[4 * g + 1] = a.y; dep[4 * g + 2] = a.z; dep[4 * g + 3] = a.w;
    sig[4 * g] = b.x; sig[4 * g + 1] = b.y; sig[4 * g + 2] = b.z; sig[4 * g + 3] = b.w;
  }
  if (dirs && rv) {
    const float nrm = dir_norm(dirs + r * dir_stride);
#pragma unroll
    for (int j = 0; j < S; ++j) dep[j] = __fmul_rn(dep[j], nrm);
  }
  ray_weights_regs<S4>(dep, sig, s0, P, act, lpr, sub, w);
#pragma unroll
  for (int g = 0; g < S4; ++g)
    if (rv && s0 + 4 * g < P)
      *reinterpret_cast<float4*>(w_out + r * P + s0 + 4 * g) = make_float4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
}

// a8  maxBlurFilter                                  /root/reference/nerf/mip_methods.py:61-66
__device__ __forceinline__ float max_blur_at(const float* w, int i, int P, float alpha) {
  float front = (i == 0) ? w[0] : fmaxf(w[i - 1], w[i]);
  float rear = (i == P - 1) ? w[P - 1] : fmaxf(w[i], w[i + 1]);
  return __fadd_rn(__fmul_rn(0.5f, __fadd_rn(front, rear)), alpha);
}
__global__ void max_blur_kernel(const float* __restrict__ w, int64_t n_rays, int P, float alpha,
                                float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * P) return;
  int64_t r = i / P;
  int s = (int)(i % P);
  out[i] = max_blur_at(w + r * P, s, P, alpha);
}

// vectorised: four samples per thread (P % 4 == 0): one float4 load (+ two neighbours through L1), one float4 store
__global__ void max_blur_vec4_kernel(const float* __restrict__ w, int64_t n_rays, int P, float alpha, float* __restrict__ out) {
  const int P4 = P >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * P4) return;
  const int64_t r = i / P4;
  const int s0 = (int)(i - r * P4) * 4;
  const float* wr = w + r * P;
  const float4 c = __ldg(reinterpret_cast<const float4*>(wr + s0));
  const float v[6] = {s0 > 0 ? __ldg(wr + s0 - 1) : c.x, c.x, c.y, c.z, c.w, s0 + 4 < P ? __ldg(wr + s0 + 4) : c.w};
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float front = (s0 + k == 0) ? v[1] : fmaxf(v[k], v[k + 1]);
    const float rear = (s0 + k == P - 1) ? v[k + 1] : fmaxf(v[k + 1], v[k + 2]);
    o[k] = __fadd_rn(__fmul_rn(0.5f, __fadd_rn(front, rear)), alpha);
  }
  reinterpret_cast<float4*>(out)[i] = make_float4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------
// a9  inverse-CDF sampling                       /root/reference/nerf/utils.py:34-44,108-133
//
// Reduction order (documented; see DESIGN.md "staged exactness"):
//   total = fp32( sum_i (double)(w_i + 1e-5) )         (warp tree in fp64, rounded once)
//   pdf_i = fp32(w_i + 1e-5) / total                   (IEEE fp32 divide)
//   cdf_k = fp32( sum_{i<k} (double)pdf_i )            (fp64 running sum, rounded per element —
//            what torch.cumsum produces on CPU for fp32 input)
// ------------------------------------------------------------------------------------------
// Builds cdf[0..B-1] in shared memory from B-1 weights.  sh_cdf may alias nothing else.
__device__ __forceinline__ void build_cdf_warp(const float* sh_wgt /*B-1*/, float* sh_cdf /*B*/, int B,
                                               int lane) {
  const int nw = B - 1;
  double part = 0.0;
  for (int i = lane; i < nw; i += 32) part += (double)__fadd_rn(sh_wgt[i], 1e-5f);
  const float total = (float)warp_sum_d(part);
  double carry = 0.0;
  if (lane == 0) sh_cdf[0] = 0.f;
  for (int base = 0; base < nw; base += 32) {
    int i = base + lane;
    double v = (i < nw) ? (double)__fdiv_rn(__fadd_rn(sh_wgt[i], 1e-5f), total) : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (i < nw) sh_cdf[i + 1] = (float)(carry + v);
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
  __syncwarp();
}

// torch.searchsorted(cdf, u, right=True): first index with cdf[idx] > u, in [0, B].
__device__ __forceinline__ int upper_bound(const float* cdf, int B, float u) {
  int lo = 0, hi = B;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One draw: returns the sample, writes below/above.
__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int B, float u,
                                            int* below_o, int* above_o) {
  int idx = upper_bound(cdf, B, u);
  int below = max(idx - 1, 0);
  int above = min(idx, B - 1);
  float cb = cdf[below], ca = cdf[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < 1e-5f) denom = 1.f;
  float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  float bb = bins[below], ba = bins[above];
  *below_o = below;
  *above_o = above;
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

// Rank sort of n keys held in shared memory (ascending, ties broken by original index so the
// result is the stable order).  Each lane owns keys lane, lane+32, ...; writes rank[] in place
// of a scatter: out_key[rank] = key, out_pay[rank] = payload.
__device__ __forceinline__ void rank_sort_warp(const float* sh_key, const int* sh_pay, int n,
                                               float* sh_key_out, int* sh_pay_out, int lane) {
  for (int i = lane; i < n; i += 32) {
    float k = sh_key[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      float o = sh_key[j];
      rank += (o < k) || (o == k && j < i);
    }
    sh_key_out[rank] = k;
    if (sh_pay) sh_pay_out[rank] = sh_pay[i];
  }
  __syncwarp();
}

constexpr int kMaxDraw = 264;  // >= 129 (Mip), 193 merged (Ref)

// Sorting the draws of one ray when only the values are needed (the fused resample): order them by their uniform u
// first -- the inverse CDF is monotone, so that IS the sorted order except for 1-ulp effects at bin edges, and u is
// uniform whatever the density looks like, so a counting sort over 128 buckets is O(n) with ~1 key per bucket -- then
// repair with odd-even transposition passes until one pass swaps nothing (normally the first).  The result is the
// sorted value sequence, i.e. exactly what rank_sort_warp / torch.sort produce; only the cost differs
// (n^2 / 32 = 520 comparisons per lane before).
constexpr int kSortBuckets = 128;
__device__ __forceinline__ int u_bucket(float u) {
  int b = (int)(u * (float)kSortBuckets);
  return b < 0 ? 0 : (b >= kSortBuckets ? kSortBuckets - 1 : b);
}
__device__ __forceinline__ void sort_by_u_warp(const float* key, const float* su, int n, int* cnt, int* off, float* tmp_u,
                                               float* tmp_z, float* out, int lane) {
  for (int b = lane; b < kSortBuckets; b += 32) cnt[b] = 0;
  __syncwarp();
  for (int i = lane; i < n; i += 32) atomicAdd(&cnt[u_bucket(su[i])], 1);
  __syncwarp();
  int carry = 0;
#pragma unroll
  for (int base = 0; base < kSortBuckets; base += 32) {
    const int v = cnt[base + lane];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    off[base + lane] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
  for (int b = lane; b < kSortBuckets; b += 32) cnt[b] = 0;
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const float uu = su[i];
    const int b = u_bucket(uu);
    const int p = off[b] + atomicAdd(&cnt[b], 1);
    tmp_u[p] = uu;
    tmp_z[p] = key[i];
  }
  __syncwarp();
  for (int p = lane; p < n; p += 32) {
    const float uu = tmp_u[p];
    const int b = u_bucket(uu);
    const int lo = off[b], hi = lo + cnt[b];
    int r = 0;
    for (int q = lo; q < hi; ++q) {
      const float uq = tmp_u[q];
      r += (uq < uu) || (uq == uu && q < p);
    }
    out[lo + r] = tmp_z[p];
  }
  __syncwarp();
  for (;;) {
    bool swapped = false;
#pragma unroll
    for (int phase = 0; phase < 2; ++phase) {
      for (int a = phase + 2 * lane; a + 1 < n; a += 64) {
        const float x = out[a], y = out[a + 1];
        if (x > y) { out[a] = y; out[a + 1] = x; swapped = true; }
      }
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, swapped)) break;
  }
}

// sample_pdf: bins (R,B), weights (R,B-1), u (R,N)
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, const float* __restrict__ u,
                  uint64_t seed, int64_t ray_offset, int64_t n_rays, int B, int N,
                  float* __restrict__ samples, int64_t* __restrict__ below_o, int64_t* __restrict__ above_o) {
  __shared__ float sh_bins[kWarpsPerBlock][kMaxSamples];
  __shared__ float sh_wgt[kWarpsPerBlock][kMaxSamples];
  __shared__ float sh_cdf[kWarpsPerBlock][kMaxSamples];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= n_rays) return;
  for (int i = lane; i < B; i += 32) sh_bins[warp][i] = bins[r * B + i];
  for (int i = lane; i < B - 1; i += 32) sh_wgt[warp][i] = weights[r * (B - 1) + i];
  __syncwarp();
  build_cdf_warp(sh_wgt[warp], sh_cdf[warp], B, lane);
  for (int i = lane; i < N; i += 32) {
    float uu = u ? u[r * N + i] : philox_uniform(seed, (uint64_t)(ray_offset + r), (uint32_t)i, 1u);
    int b, a;
    samples[r * N + i] = invert_cdf(sh_cdf[warp], sh_bins[warp], B, uu, &b, &a);
    below_o[r * N + i] = b;
    if (above_o) above_o[r * N + i] = a;
  }
}

__global__ void search_cdf_kernel(const float* __restrict__ cdf, const float* __restrict__ u, int64_t n_rays,
                                  int B, int N, int64_t* __restrict__ inds) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * N) return;
  int64_t r = i / N;
  inds[i] = upper_bound(cdf + r * B, B, u[i]);
}

// The per-ray resampling core shared by inverse_sample (weights given) and resample (density
// given).  mode 0: sh_w already holds the P proposal weights to sample from (already blurred).
//          mode 1: sh_w holds raw density; do get_weights + maxBlur first.
// CAP: capacity of the per-sample rows (and, + 8, of the per-draw rows).  The 64-sample / 129-draw render path takes
// CAP = 136 (4.9 KB per warp instead of 9.3 KB: 44 instead of 24 resident warps per SM for a latency-bound kernel).
template <int CAP>
struct ResampleSmemT {
  float z[CAP];      // coarse depths
  float a[CAP];      // density / weights scratch
  float w[CAP];      // weights
  float bins[CAP];   // mid points
  float cdf[CAP];
  float key[CAP + 8];
  float key_sorted[CAP + 8];
  int pay[CAP + 8];
  int pay_sorted[CAP + 8];
};
static_assert(kMaxSamples + 8 == kMaxDraw, "draw rows are sample rows + 8");

// The draws of one lane, kDrawIL at a time: their binary searches advance in lockstep so that the dependent shared-memory
// loads of one search hide under the others' (ncu: the one-draw-at-a-time loop held 27 % of the kernel's stall samples).
// Same comparisons as upper_bound / invert_cdf, hence the same indices and values.
constexpr int kDrawIL = 5;
constexpr int kResampleWarps = 4;
constexpr int kSmallCap = 136;     // >= 129 draws / 128 buckets of the u-ordered sort
__device__ __forceinline__ void draw_samples_warp(const float* cdf, const float* bins, int B, const float* __restrict__ u_row, uint64_t seed,
                                                  uint64_t ray_id, int N, float* key, int* pay, float* u_keep, int lane) {
  const int steps = 32 - __clz(B);                       // |[0, B]| = B + 1 candidates
  for (int base = 0; base < N; base += 32 * kDrawIL) {
    float uu[kDrawIL];
    int lo[kDrawIL], hi[kDrawIL];
#pragma unroll
    for (int q = 0; q < kDrawIL; ++q) {
      const int i = base + lane + 32 * q;
      uu[q] = 0.f;
      if (i < N) uu[q] = u_row ? u_row[i] : philox_uniform(seed, ray_id, (uint32_t)i, 1u);
      lo[q] = 0;
      hi[q] = i < N ? B : 0;
    }
    for (int st = 0; st < steps; ++st) {
#pragma unroll
      for (int q = 0; q < kDrawIL; ++q) {
        const int mid = (lo[q] + hi[q]) >> 1;
        const float c = cdf[mid < B ? mid : B - 1];
        if (lo[q] < hi[q]) {
          if (c <= uu[q]) lo[q] = mid + 1; else hi[q] = mid;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kDrawIL; ++q) {
      const int i = base + lane + 32 * q;
      if (i < N) {
        const int below = max(lo[q] - 1, 0), above = min(lo[q], B - 1);
        const float cb = cdf[below], ca = cdf[above];
        float denom = __fsub_rn(ca, cb);
        if (denom < 1e-5f) denom = 1.f;
        const float t = __fdiv_rn(__fsub_rn(uu[q], cb), denom);
        const float bb = bins[below], ba = bins[above];
        key[i] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
        pay[i] = below;
        if (u_keep) u_keep[i] = uu[q];
      }
    }
  }
}

template <int CAP>
__global__ void __launch_bounds__(32 * kResampleWarps)
resample_kernel(int mode, const float* __restrict__ win /*weights or sigma (R,P)*/, const float* __restrict__ z,
                const float* __restrict__ rays, const float* __restrict__ u, uint64_t seed, int64_t ray_offset,
                int64_t n_rays, int P, int N, int sort, int n_keep, float blur_alpha, int act,
                float* __restrict__ samples_out, int64_t* __restrict__ below_out) {
  __shared__ ResampleSmemT<CAP> sm[kResampleWarps];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * kResampleWarps + warp;
  if (r >= n_rays) return;
  ResampleSmemT<CAP>& s = sm[warp];
  for (int i = lane; i < P; i += 32) {
    s.z[i] = z[r * P + i];
    s.a[i] = win[r * P + i];
  }
  __syncwarp();
  if (mode == 1) {
    // ProposalNetwork.get_weights(density, z, dirs)     (addtional.py:99-107)
    float nrm = dir_norm(rays + r * 6 + 3);
    for (int i = lane; i < P; i += 32) s.bins[i] = __fmul_rn(s.z[i], nrm);  // scaled depths (temp)
    __syncwarp();
    ray_weights_warp(s.bins, s.a, s.cdf, P, act, lane);                      // raw weights -> cdf (temp)
    for (int i = lane; i < P; i += 32) s.w[i] = max_blur_at(s.cdf, i, P, blur_alpha);
    __syncwarp();
  } else {
    for (int i = lane; i < P; i += 32) s.w[i] = s.a[i];
    __syncwarp();
  }
  // inverseSample (utils.py:34-44): bins = mid(z) (P-1), weights[1:-1] (P-2)
  const int B = P - 1;
  for (int i = lane; i < B; i += 32) s.bins[i] = __fmul_rn(0.5f, __fadd_rn(s.z[i + 1], s.z[i]));
  __syncwarp();
  build_cdf_warp(s.w + 1, s.cdf, B, lane);
  const bool values_only = sort && below_out == nullptr && N <= CAP;   // the fused resample path
  // (the density / weight scratch s.a is dead by now: it keeps the uniforms for the u-ordered sort)
  draw_samples_warp(s.cdf, s.bins, B, u ? u + r * N : nullptr, seed, (uint64_t)(ray_offset + r), N, s.key, s.pay, values_only ? s.a : nullptr, lane);
  __syncwarp();
  const float* kk = s.key;
  const int* pp = s.pay;
  if (values_only) {
    // w, bins, cdf, key_sorted and z are dead after the draws: counters, offsets, grouped u, grouped z, sorted output
    sort_by_u_warp(s.key, s.a, N, reinterpret_cast<int*>(s.w), reinterpret_cast<int*>(s.bins), s.cdf, s.key_sorted, s.z, lane);
    kk = s.z;
  } else if (sort) {
    rank_sort_warp(s.key, s.pay, N, s.key_sorted, s.pay_sorted, lane);
    kk = s.key_sorted;
    pp = s.pay_sorted;
  }
  for (int i = lane; i < n_keep; i += 32) {
    samples_out[r * n_keep + i] = kk[i];
    if (below_out) below_out[r * n_keep + i] = pp[i];
  }
}

// a10  NeRF.length2pts                               /root/reference/nerf/nerf_base.py:52-56
__global__ void length2pts_kernel(const float* __restrict__ rays, const float* __restrict__ z, int64_t n_rays,
                                  int P, float* __restrict__ pts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * P) return;
  int64_t r = i / P;
  const float* ray = rays + r * 6;
  float zz = z[i];
  float* o = pts + i * 6;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float d = __ldg(ray + 3 + k);
    o[k] = __fadd_rn(__ldg(ray + k), __fmul_rn(d, zz));
    o[3 + k] = d;
  }
}

// vectorised: four samples per thread (P % 4 == 0): a float4 of depths in, six float4 (4 x [point, direction]) out.  A warp's
// 32 x 96 output bytes are contiguous in global memory, so they are transposed through a 3 KB shared-memory slab and written
// as six fully coalesced 512-byte rows (direct stores at the 96-byte thread stride half-fill every 32-byte sector per
// instruction: 0.64 of the copy bandwidth).  (A flat one-float4-per-thread form measured SLOWER, 292 vs 137 us: 64-bit index
// divisions per output element.)
__global__ void __launch_bounds__(256) length2pts_vec4_kernel(const float* __restrict__ rays, const float* __restrict__ z, int64_t n_rays,
                                                              int P, float* __restrict__ pts) {
  __shared__ float4 slab[8][6 * 32];
  const int P4 = P >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t total = n_rays * P4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i_warp = i - lane;
  if (i_warp >= total) return;                 // whole warps only: the slab hand-over below is warp-synchronous
  if (i < total) {
    const int64_t r = i / P4;
    const float* ray = rays + r * 6;
    const float o[3] = {__ldg(ray), __ldg(ray + 1), __ldg(ray + 2)}, d[3] = {__ldg(ray + 3), __ldg(ray + 4), __ldg(ray + 5)};
    const float4 zv = __ldg(reinterpret_cast<const float4*>(z) + i);
    const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
    float v[24];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        v[6 * k + c] = __fadd_rn(o[c], __fmul_rn(d[c], zz[k]));
        v[6 * k + 3 + c] = d[c];
      }
#pragma unroll
    for (int k = 0; k < 6; ++k) slab[warp][lane * 6 + k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
  __syncwarp();
  const int n_valid = (int)min((int64_t)32, total - i_warp) * 6;      // float4s this warp owns
  float4* dst = reinterpret_cast<float4*>(pts) + i_warp * 6;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int q = k * 32 + lane;
    if (q < n_valid) dst[q] = slab[warp][q];
  }
}

// a13  NeRF.coarseFineMerge                          /root/reference/nerf/nerf_base.py:58-73
// Optional index bookkeeping (training path, :62-65,70-71): sort_inds = the stable sort permutation of cat(f_z, c_z) (last
// dropped), all_inds = gather(cat(f_inds, arange(C)), permutation) (full length).
__global__ void __launch_bounds__(32 * kResampleWarps)
merge_kernel(const float* __restrict__ rays, const float* __restrict__ cz, const float* __restrict__ fz,
             const int64_t* __restrict__ f_inds, int64_t n_rays, int C, int F, float* __restrict__ z_out, float* __restrict__ pts_out,
             int64_t* __restrict__ all_inds_out, int64_t* __restrict__ sort_inds_out) {
  __shared__ float key[kResampleWarps][2 * kMaxDraw];
  __shared__ float srt[kResampleWarps][2 * kMaxDraw];
  __shared__ int pay[kResampleWarps][2 * kMaxDraw];
  __shared__ int pay_srt[kResampleWarps][2 * kMaxDraw];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * kResampleWarps + warp;
  if (r >= n_rays) return;
  const int n = C + F;
  const bool want_inds = (all_inds_out != nullptr) || (sort_inds_out != nullptr);
  for (int i = lane; i < F; i += 32) key[warp][i] = fz[r * F + i];      // cat((f_zvals, c_zvals))
  for (int i = lane; i < C; i += 32) key[warp][F + i] = cz[r * C + i];
  if (want_inds)
    for (int i = lane; i < n; i += 32) pay[warp][i] = i;
  __syncwarp();
  rank_sort_warp(key[warp], want_inds ? pay[warp] : nullptr, n, srt[warp], pay_srt[warp], lane);
  const float* ray = rays + r * 6;
  for (int i = lane; i < n - 1; i += 32) {
    float zz = srt[warp][i];
    z_out[r * (n - 1) + i] = zz;
    if (sort_inds_out) sort_inds_out[r * (n - 1) + i] = pay_srt[warp][i];
    if (pts_out) {
      float* o = pts_out + (r * (n - 1) + i) * 6;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float d = __ldg(ray + 3 + k);
        o[k] = __fadd_rn(__ldg(ray + k), __fmul_rn(d, zz));
        o[3 + k] = d;
      }
    }
  }
  if (all_inds_out)
    for (int i = lane; i < n; i += 32) {
      const int src = pay_srt[warp][i];
      all_inds_out[r * n + i] = src < F ? f_inds[r * F + src] : (int64_t)(src - F);
    }
}

// ------------------------------------------------------------------------------------------
// a12  NeRF.render                                  /root/reference/nerf/nerf_base.py:90-113
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
composite_kernel(const float* __restrict__ rgbo, const float* __restrict__ z, const float* __restrict__ dirs,
                 int dir_stride, int64_t n_rays, int P, int flags, float near_t, float far_t,
                 float* __restrict__ rgb_out, float* __restrict__ w_out, float* __restrict__ depth_out,
                 float* __restrict__ acc_out, const float* __restrict__ aux, float* __restrict__ aux_out) {
  __shared__ float sh[kWarpsPerBlock][3][kMaxSamples];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= n_rays) return;
  float nrm = dir_norm(dirs + r * dir_stride);
  const float4* c4 = reinterpret_cast<const float4*>(rgbo) + r * P;
  for (int i = lane; i < P; i += 32) {
    sh[warp][0][i] = __fmul_rn(z[r * P + i], nrm);
    sh[warp][1][i] = c4[i].w;
  }
  __syncwarp();
  ray_weights_warp(sh[warp][0], sh[warp][1], sh[warp][2], P, 0, lane);
  float cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f, dep = 0.f, ax = 0.f;
  for (int i = lane; i < P; i += 32) {
    float w = sh[warp][2][i];
    float4 c = c4[i];
    if (aux) ax = fmaf(w, aux[r * P + i], ax);
    cr = fmaf(w, c.x, cr);
    cg = fmaf(w, c.y, cg);
    cb = fmaf(w, c.z, cb);
    acc += w;
    dep = fmaf(w, sh[warp][0][i], dep);
    if (w_out) w_out[r * P + i] = w;
  }
  cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); acc = warp_sum(acc); dep = warp_sum(dep);
  if (aux) ax = warp_sum(ax);
  if (lane == 0) {
    if (aux_out) aux_out[r] = ax;
    if (flags & NB2_WHITE_BKG) {
      float bg = 1.f - acc;
      cr += bg; cg += bg; cb += bg;
    }
    rgb_out[r * 3 + 0] = cr;
    rgb_out[r * 3 + 1] = cg;
    rgb_out[r * 3 + 2] = cb;
    if (depth_out) depth_out[r] = (dep - near_t) / (far_t - near_t);
    if (acc_out) acc_out[r] = acc;
  }
}

// register form (ray_weights_regs): the lane's 4 * S4 [r,g,b,sigma] samples are 16 * S4 * 4 contiguous bytes, read once
template <int S4>
__global__ void __launch_bounds__(256)
composite_regs_kernel(const float* __restrict__ rgbo, const float* __restrict__ z, const float* __restrict__ dirs, int dir_stride,
                      int64_t n_rays, int P, int flags, float near_t, float far_t, int lpr, float* __restrict__ rgb_out,
                      float* __restrict__ w_out, float* __restrict__ depth_out, float* __restrict__ acc_out,
                      const float* __restrict__ aux, float* __restrict__ aux_out) {
  constexpr int S = 4 * S4;
  const int lane = threadIdx.x & 31, sub = lane & (lpr - 1);
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t r = warp_id * (32 / lpr) + lane / lpr;
  const int s0 = sub * S;
  const bool rv = r < n_rays;
  float dep[S], sig[S], w[S], cr_[S], cg_[S], cb_[S], ax_[S];
  const float4* c4 = reinterpret_cast<const float4*>(rgbo) + r * P;
#pragma unroll
  for (int g = 0; g < S4; ++g) {
    const bool v = rv && s0 + 4 * g < P;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), x = a;
    if (v) a = __ldg(reinterpret_cast<const float4*>(z + r * P + s0 + 4 * g));
    if (v && aux) x = __ldg(reinterpret_cast<const float4*>(aux + r * P + s0 + 4 * g));
    dep[4 * g] = a.x; dep[4 * g + 1] = a.y; dep[4 * g + 2] = a.z; dep[4 * g + 3] = a.w;
    ax_[4 * g] = x.x; ax_[4 * g + 1] = x.y; ax_[4 * g + 2] = x.z; ax_[4 * g + 3] = x.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v) c = __ldg(c4 + s0 + 4 * g + k);
      cr_[4 * g + k] = c.x; cg_[4 * g + k] = c.y; cb_[4 * g + k] = c.z; sig[4 * g + k] = c.w;
    }
  }
  if (rv) {
    const float nrm = dir_norm(dirs + r * dir_stride);
#pragma unroll
    for (int j = 0; j < S; ++j) dep[j] = __fmul_rn(dep[j], nrm);
  }
  ray_weights_regs<S4>(dep, sig, s0, P, 0, lpr, sub, w);
  float cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f, dp = 0.f, ax = 0.f;
#pragma unroll
  for (int j = 0; j < S; ++j) {
    cr = fmaf(w[j], cr_[j], cr);
    cg = fmaf(w[j], cg_[j], cg);
    cb = fmaf(w[j], cb_[j], cb);
    acc += w[j];
    dp = fmaf(w[j], dep[j], dp);
    ax = fmaf(w[j], ax_[j], ax);
  }
  if (w_out) {
#pragma unroll
    for (int g = 0; g < S4; ++g)
      if (rv && s0 + 4 * g < P)
        *reinterpret_cast<float4*>(w_out + r * P + s0 + 4 * g) = make_float4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
  }
  for (int d = lpr >> 1; d > 0; d >>= 1) {
    cr += __shfl_xor_sync(0xffffffffu, cr, d);
    cg += __shfl_xor_sync(0xffffffffu, cg, d);
    cb += __shfl_xor_sync(0xffffffffu, cb, d);
    acc += __shfl_xor_sync(0xffffffffu, acc, d);
    dp += __shfl_xor_sync(0xffffffffu, dp, d);
    if (aux) ax += __shfl_xor_sync(0xffffffffu, ax, d);
  }
  if (rv && sub == 0) {
    if (aux_out) aux_out[r] = ax;
    if (flags & NB2_WHITE_BKG) {
      const float bg = 1.f - acc;
      cr += bg; cg += bg; cb += bg;
    }
    rgb_out[r * 3 + 0] = cr;
    rgb_out[r * 3 + 1] = cg;
    rgb_out[r * 3 + 2] = cb;
    if (depth_out) depth_out[r] = (dp - near_t) / (far_t - near_t);
    if (acc_out) acc_out[r] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// f1  validSampler (training-side ray sampler)          /root/reference/nerf/utils.py:72-94
//   ray r picks pixel indices[r]; dir = R * ((cx + .5)/fx, (cy + .5)/fy, -1); z = base[s] + U * res; pts = o + dir * z
// ------------------------------------------------------------------------------------------
__global__ void valid_sampler_kernel(const float* __restrict__ rgbs, const int64_t* __restrict__ coords,
                                     const float* __restrict__ cam_tf, const int64_t* __restrict__ indices,
                                     const float* __restrict__ base_z, const float* __restrict__ jitter, float fx, float fy,
                                     float resolution, uint64_t seed, int64_t ray_offset, int64_t n_pixels, int64_t n_rays, int P,
                                     float* __restrict__ pts_out, float* __restrict__ len_out, float* __restrict__ rgb_out,
                                     float* __restrict__ rays_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * P) return;
  int64_t r = i / P;
  int s = (int)(i % P);
  int64_t pix;
  if (indices) {
    pix = indices[r];
  } else {   // device RNG: stream 2, one draw per ray
    uint32_t c[4] = {(uint32_t)(ray_offset + r), (uint32_t)((uint64_t)(ray_offset + r) >> 32), 0u, 2u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    pix = (int64_t)((((uint64_t)c[0] << 32) | c[1]) % (uint64_t)n_pixels);
  }
  float cx = __fdiv_rn(__fadd_rn((float)coords[pix * 2 + 0], 0.5f), fx);
  float cy = __fdiv_rn(__fadd_rn((float)coords[pix * 2 + 1], 0.5f), fy);
  float d[3], o[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float r0 = __ldg(cam_tf + k * 4 + 0), r1 = __ldg(cam_tf + k * 4 + 1), r2 = __ldg(cam_tf + k * 4 + 2);
    o[k] = __ldg(cam_tf + k * 4 + 3);
    d[k] = __fadd_rn(__fadd_rn(__fmul_rn(cx, r0), __fmul_rn(cy, r1)), __fmul_rn(-1.f, r2));
  }
  float j = jitter ? jitter[i] : philox_uniform(seed, (uint64_t)(ray_offset + r), (uint32_t)s, 0u);
  float z = __fadd_rn(__ldg(base_z + s), __fmul_rn(j, resolution));
  len_out[i] = z;
#pragma unroll
  for (int k = 0; k < 3; ++k) pts_out[i * 3 + k] = __fadd_rn(o[k], __fmul_rn(d[k], z));
  if (s == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rgb_out[r * 3 + k] = rgbs[pix * 3 + k];
      rays_out[r * 6 + k] = o[k];
      rays_out[r * 6 + 3 + k] = d[k];
    }
  }
}

// ------------------------------------------------------------------------------------------
// f1  getBounds (proposal-loss bounds)                 /root/reference/nerf/addtional.py:14-18
//   sat = [0, cumsum(w)]; out[r, j] = sat[inds[r, j+1] + 1] - sat[inds[r, j]]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
get_bounds_kernel(const float* __restrict__ w, const int64_t* __restrict__ inds, int64_t n_rays, int P, int K,
                  float* __restrict__ out) {
  __shared__ float sat[kWarpsPerBlock][kMaxSamples + 1];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= n_rays) return;
  double carry = 0.0;
  if (lane == 0) sat[warp][0] = 0.f;
  for (int base = 0; base < P; base += 32) {      // torch.cumsum on fp32 == fp64 running sum rounded per element
    int i = base + lane;
    double v = (i < P) ? (double)w[r * P + i] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (i < P) sat[warp][i + 1] = (float)(carry + v);
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
  __syncwarp();
  for (int j = lane; j < K - 1; j += 32) {
    int64_t st = inds[r * K + j], en = inds[r * K + j + 1] + 1;
    st = st < 0 ? 0 : (st > P ? P : st);
    en = en < 0 ? 0 : (en > P ? P : en);
    out[r * (K - 1) + j] = __fsub_rn(sat[warp][en], sat[warp][st]);
  }
}

// ------------------------------------------------------------------------------------------
// next row 8f-3 (Ref-NeRF, forward): integrated directional encoding   /root/reference/nerf/ref_func.py:78-108
//   out[p, i]           = Re((x + iy)^m_i) * (sum_k z^k mat[k, i]) * exp(-l_i (l_i + 1) / 2 * kappa_inv[p])
//   out[p, n_pairs + i] = Im(...)                       for the (m_i, l_i) pairs of ref_func.py:38-49
// One thread per direction; a block's 128 x 2 n_pairs tile is assembled in shared memory and written back coalesced.
// ------------------------------------------------------------------------------------------
constexpr int kIdeBlockPts = 128;
constexpr int kIdeMaxPairs = 36;   // deg_view = 5
constexpr int kIdeMaxPow = 17;     // l_max + 1
__global__ void __launch_bounds__(kIdeBlockPts) ide_kernel(const float* __restrict__ xyz, const float* __restrict__ kappa_inv,
                                                           int64_t n, int n_pairs, int n_pow, const float* __restrict__ mat,
                                                           const int* __restrict__ ml, float* __restrict__ out) {
  extern __shared__ float ide_smem[];
  float* sh_mat = ide_smem;                        // [n_pow][n_pairs]
  int* sh_ml = reinterpret_cast<int*>(sh_mat + kIdeMaxPow * kIdeMaxPairs);   // [2][n_pairs]: m then l
  float* tile = reinterpret_cast<float*>(sh_ml + 2 * kIdeMaxPairs);         // [kIdeBlockPts][2 n_pairs]
  for (int t = threadIdx.x; t < n_pow * n_pairs; t += blockDim.x) sh_mat[t] = __ldg(mat + t);
  for (int t = threadIdx.x; t < 2 * n_pairs; t += blockDim.x) sh_ml[t] = __ldg(ml + t);
  __syncthreads();
  const int width = 2 * n_pairs;
  const int64_t p0 = (int64_t)blockIdx.x * kIdeBlockPts;
  const int npts = (int)min((int64_t)kIdeBlockPts, n - p0);
  if ((int)threadIdx.x < npts) {
    const int64_t p = p0 + threadIdx.x;
    const float x = __ldg(xyz + 3 * p), y = __ldg(xyz + 3 * p + 1), z = __ldg(xyz + 3 * p + 2);
    const float kinv = __ldg(kappa_inv + p);
    float zp[kIdeMaxPow];          // z^k
    float cr[kIdeMaxPow], ci[kIdeMaxPow];   // (x + iy)^m
    zp[0] = 1.f; cr[0] = 1.f; ci[0] = 0.f;
#pragma unroll
    for (int k = 1; k < kIdeMaxPow; ++k) {
      zp[k] = zp[k - 1] * z;
      cr[k] = cr[k - 1] * x - ci[k - 1] * y;
      ci[k] = cr[k - 1] * y + ci[k - 1] * x;
    }
    float* row = tile + threadIdx.x * width;
    for (int i = 0; i < n_pairs; ++i) {
      const int m = sh_ml[i], l = sh_ml[n_pairs + i];
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kIdeMaxPow; ++k)
        if (k < n_pow) acc = fmaf(zp[k], sh_mat[k * n_pairs + i], acc);
      const float att = expf(-(0.5f * (float)(l * (l + 1))) * kinv);
      float pr = 0.f, pi = 0.f;
#pragma unroll
      for (int k = 0; k < kIdeMaxPow; ++k)
        if (k == m) { pr = cr[k]; pi = ci[k]; }
      row[i] = pr * acc * att;
      row[n_pairs + i] = pi * acc * att;
    }
  }
  __syncthreads();
  float* dst = out + p0 * width;
  const int total = npts * width;
  if ((total & 3) == 0 && ((((uintptr_t)dst) & 15) == 0)) {
    for (int t = threadIdx.x; t < total / 4; t += blockDim.x)
      reinterpret_cast<float4*>(dst)[t] = reinterpret_cast<const float4*>(tile)[t];
  } else {
    for (int t = threadIdx.x; t < total; t += blockDim.x) dst[t] = tile[t];
  }
}

// linear_to_srgb                                                        /root/reference/nerf/nerf_helper.py:50-56
__global__ void linear_to_srgb_kernel(const float* __restrict__ lin, int64_t n, float* __restrict__ out) {
  const float eps = 1.1920928955078125e-07f;   // torch.finfo(torch.float32).eps
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(lin + i);
    const float s0 = __fmul_rn(12.92f, v);
    const float s1 = __fdiv_rn(__fsub_rn(__fmul_rn(211.f, powf(fmaxf(eps, v), 0.41666666f)), 11.f), 200.f);
    out[i] = (v <= 0.0031308f) ? s0 : s1;
  }
}

int launch_ipe_sumsq(nb2_handle* h, const float* rays, int64_t n_rays, double* out, cudaStream_t st) {
  NB2_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
  int g = (int)std::min<int64_t>(grid_for(n_rays, 256), 1184);
  ipe_sumsq_kernel<<<g, 256, 0, st>>>(rays, n_rays, out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

}  // namespace nb2

// ==========================================================================================
// C ABI wrappers for the stages above
// ==========================================================================================
using namespace nb2;

#define NB2_H(h) NB2_ENTER(h)

extern "C" int nb2_generate_rays(nb2_handle* h, const float* pose, int H, int W, float fx, float fy,
                                 int64_t pix_offset, int64_t n, float* rays_out, void* stream) {
  NB2_H(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(pose && rays_out && H > 0 && W > 0 && fx != 0.f && fy != 0.f, "generate_rays: bad arguments");
  NB2_CHECK_ARG(pix_offset >= 0 && n >= 0 && pix_offset + n <= (int64_t)H * W, "generate_rays: pixel range outside image");
  if (n == 0) return NB2_OK;
  generate_rays_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(pose, H, W, fx, fy, pix_offset, n, rays_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_sample_coarse(nb2_handle* h, const float* rays, const float* base_z, const float* jitter,
                                 float resolution, uint64_t seed, int64_t ray_offset, int64_t n_rays,
                                 int n_samples, float* z_out, float* pts_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(base_z && z_out && n_samples > 0 && n_rays >= 0, "sample_coarse: bad arguments");
  NB2_CHECK_ARG(!pts_out || rays, "sample_coarse: pts_out requires rays");
  if (n_rays == 0) return NB2_OK;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if ((n_samples & 3) == 0 && al16(base_z) && al16(jitter) && al16(z_out) && al16(pts_out)) {
    sample_coarse_vec4_kernel<<<grid_for(n_rays * (n_samples >> 2), 256), 256, 0, (cudaStream_t)stream>>>(
        rays, base_z, jitter, resolution, seed, ray_offset, n_rays, n_samples, z_out, pts_out);
  } else {
    sample_coarse_kernel<<<grid_for(n_rays * n_samples, 256), 256, 0, (cudaStream_t)stream>>>(
        rays, base_z, jitter, resolution, seed, ray_offset, n_rays, n_samples, z_out, pts_out);
  }
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_posenc(nb2_handle* h, const float* x, int64_t n, int dims, int levels, float* out, void* stream) {
  NB2_H(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(x && out && n >= 0 && dims >= 1 && dims <= 4 && levels >= 1 && levels <= 16, "posenc: bad arguments");
  if (n == 0) return NB2_OK;
  size_t smem = (size_t)kPeBlockPts * 2 * dims * levels * sizeof(float);
  {
    const int rc = kernel_set_smem(h, (const void*)posenc_kernel, 128 * 1024);
    if (rc != NB2_OK) return rc;
  }
  if (dims == 3) {
    const int rc = kernel_set_smem(h, (const void*)posenc3_kernel, 128 * 1024);
    if (rc != NB2_OK) return rc;
    posenc3_kernel<<<grid_for(n, kPe3Pts), 3 * kPe3Pts, (size_t)kPe3Pts * 6 * levels * sizeof(float), (cudaStream_t)stream>>>(x, n, levels, out);
  } else {
    posenc_kernel<<<grid_for(n, kPeBlockPts), 256, smem, (cudaStream_t)stream>>>(x, n, dims, levels, out);
  }
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_ipe(nb2_handle* h, const float* zvals, const float* rays, int64_t n_rays, int n_cones,
                       int levels, float radius, float* feat_out, float* mu_out, float* mu_t_out,
                       void* scratch, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(zvals && rays && feat_out && scratch && n_cones > 0 && levels >= 1 && levels <= 16, "ipe: bad arguments");
  if (n_rays == 0) return NB2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  NB2_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  int g = (int)std::min<int64_t>(grid_for(n_rays, 256), 1184);
  ipe_sumsq_kernel<<<g, 256, 0, st>>>(rays, n_rays, (double*)scratch);
  NB2_LAUNCH_CHECK(h);
  {
    const int smem = kIpeBlock * (6 * levels + 1) * (int)sizeof(float);
    const int rc = kernel_set_smem(h, (const void*)ipe_kernel, 64 * 1024);
    if (rc != NB2_OK) return rc;
    ipe_kernel<<<grid_for(n_rays * n_cones, kIpeBlock), kIpeBlock, smem, st>>>(zvals, rays, n_rays, n_cones, levels, radius,
                                                                              (const double*)scratch, feat_out, mu_out, mu_t_out);
  }
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_weights_from_sigma(nb2_handle* h, const float* sigma, const float* z, const float* dirs,
                                      int dir_stride, int64_t n_rays, int n_samples, int act, float* weights_out,
                                      void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(sigma && z && weights_out && n_samples >= 1 && n_samples <= kMaxSamples, "weights_from_sigma: n_samples must be in [1,%d]", kMaxSamples);
  NB2_CHECK_ARG(!dirs || dir_stride >= 3, "weights_from_sigma: dir_stride < 3");
  NB2_CHECK_ARG(act >= 0 && act <= 2, "weights_from_sigma: unknown activation %d", act);
  if (n_rays == 0) return NB2_OK;
  if ((n_samples & 3) == 0 && aligned16(sigma) && aligned16(z) && aligned16(weights_out)) {
    int s4, lpr;
    regs_geometry(n_samples, s4, lpr);
    const int grid = grid_for(grid_for(n_rays, 32 / lpr), 8);
    if (s4 == 1)
      weights_regs_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(sigma, z, dirs, dir_stride, n_rays, n_samples, act, lpr, weights_out);
    else
      weights_regs_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(sigma, z, dirs, dir_stride, n_rays, n_samples, act, lpr, weights_out);
  } else {
    const int64_t blocks = grid_for(n_rays, kWarpsPerBlock);
    const int grid = (int)std::min<int64_t>(blocks, (int64_t)h->sm_count * 8);     // 64 resident warps per SM, each looping over rays
    weights_kernel<<<grid, 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(sigma, z, dirs, dir_stride, n_rays, n_samples, act, weights_out);
  }
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_max_blur(nb2_handle* h, const float* weights, int64_t n_rays, int n_samples, float alpha,
                            float* out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(weights && out && n_samples >= 1, "max_blur: bad arguments");
  if (n_rays == 0) return NB2_OK;
  if ((n_samples & 3) == 0 && n_samples >= 8 && (reinterpret_cast<uintptr_t>(weights) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    max_blur_vec4_kernel<<<grid_for(n_rays * (n_samples >> 2), 256), 256, 0, (cudaStream_t)stream>>>(weights, n_rays, n_samples, alpha, out);
  else
    max_blur_kernel<<<grid_for(n_rays * n_samples, 256), 256, 0, (cudaStream_t)stream>>>(weights, n_rays, n_samples, alpha, out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_sample_pdf(nb2_handle* h, const float* bins, const float* weights, const float* u, uint64_t seed,
                              int64_t ray_offset, int64_t n_rays, int n_bins, int n_draw, float* samples_out,
                              int64_t* below_out, int64_t* above_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(bins && weights && samples_out && below_out, "sample_pdf: null pointer");
  NB2_CHECK_ARG(n_bins >= 2 && n_bins <= kMaxSamples, "sample_pdf: n_bins must be in [2,%d]", kMaxSamples);
  NB2_CHECK_ARG(n_draw >= 1, "sample_pdf: n_draw < 1");
  if (n_rays == 0) return NB2_OK;
  sample_pdf_kernel<<<grid_for(n_rays, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(
      bins, weights, u, seed, ray_offset, n_rays, n_bins, n_draw, samples_out, below_out, above_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_search_cdf(nb2_handle* h, const float* cdf, const float* u, int64_t n_rays, int n_cdf, int n_draw,
                              int64_t* inds_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(cdf && u && inds_out && n_cdf >= 1 && n_draw >= 1, "search_cdf: bad arguments");
  if (n_rays == 0) return NB2_OK;
  search_cdf_kernel<<<grid_for(n_rays * n_draw, 256), 256, 0, (cudaStream_t)stream>>>(cdf, u, n_rays, n_cdf, n_draw, inds_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_inverse_sample(nb2_handle* h, const float* weights, const float* z, const float* u, uint64_t seed,
                                  int64_t ray_offset, int64_t n_rays, int n_samples, int n_draw, int sort,
                                  float* samples_out, int64_t* below_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(weights && z && samples_out, "inverse_sample: null pointer");
  NB2_CHECK_ARG(n_samples >= 3 && n_samples <= kMaxSamples, "inverse_sample: n_samples must be in [3,%d]", kMaxSamples);
  NB2_CHECK_ARG(n_draw >= 1 && n_draw <= kMaxDraw, "inverse_sample: n_draw must be in [1,%d]", kMaxDraw);
  if (n_rays == 0) return NB2_OK;
  if (n_samples <= kSmallCap && n_draw <= kSmallCap)
    resample_kernel<kSmallCap><<<grid_for(n_rays, kResampleWarps), 32 * kResampleWarps, 0, (cudaStream_t)stream>>>(
        0, weights, z, nullptr, u, seed, ray_offset, n_rays, n_samples, n_draw, sort, n_draw, 0.f, 0, samples_out, below_out);
  else
    resample_kernel<kMaxSamples><<<grid_for(n_rays, kResampleWarps), 32 * kResampleWarps, 0, (cudaStream_t)stream>>>(
        0, weights, z, nullptr, u, seed, ray_offset, n_rays, n_samples, n_draw, sort, n_draw, 0.f, 0, samples_out, below_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_resample(nb2_handle* h, const float* sigma, const float* z, const float* rays, const float* u,
                            uint64_t seed, int64_t ray_offset, int64_t n_rays, int n_samples, int n_draw,
                            float blur_alpha, int flags, float* z_fine_out, int64_t* below_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(sigma && z && rays && z_fine_out, "resample: null pointer");
  NB2_CHECK_ARG(n_samples >= 3 && n_samples <= kMaxSamples, "resample: n_samples must be in [3,%d]", kMaxSamples);
  NB2_CHECK_ARG(n_draw >= 2 && n_draw <= kMaxDraw, "resample: n_draw must be in [2,%d]", kMaxDraw);
  if (n_rays == 0) return NB2_OK;
  int act = (flags & NB2_DENSITY_SOFTPLUS) ? 1 : 0;
  // softplus'd density is then passed through get_weights' own relu (a no-op on positives)
  if (n_samples <= kSmallCap && n_draw <= kSmallCap)
    resample_kernel<kSmallCap><<<grid_for(n_rays, kResampleWarps), 32 * kResampleWarps, 0, (cudaStream_t)stream>>>(
        1, sigma, z, rays, u, seed, ray_offset, n_rays, n_samples, n_draw, 1, n_draw - 1, blur_alpha, act, z_fine_out, below_out);
  else
    resample_kernel<kMaxSamples><<<grid_for(n_rays, kResampleWarps), 32 * kResampleWarps, 0, (cudaStream_t)stream>>>(
        1, sigma, z, rays, u, seed, ray_offset, n_rays, n_samples, n_draw, 1, n_draw - 1, blur_alpha, act, z_fine_out, below_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_length2pts(nb2_handle* h, const float* rays, const float* z, int64_t n_rays, int n_samples,
                              float* pts_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(rays && z && pts_out && n_samples >= 1, "length2pts: bad arguments");
  if (n_rays == 0) return NB2_OK;
  if ((n_samples & 3) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(pts_out) & 15) == 0)
    length2pts_vec4_kernel<<<grid_for(n_rays * (n_samples >> 2), 256), 256, 0, (cudaStream_t)stream>>>(rays, z, n_rays, n_samples, pts_out);
  else
    length2pts_kernel<<<grid_for(n_rays * n_samples, 256), 256, 0, (cudaStream_t)stream>>>(rays, z, n_rays, n_samples, pts_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_coarse_fine_merge(nb2_handle* h, const float* rays, const float* c_z, const float* f_z,
                                     int64_t n_rays, int n_coarse, int n_fine, float* z_out, float* pts_out,
                                     void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(rays && c_z && f_z && z_out, "coarse_fine_merge: null pointer");
  NB2_CHECK_ARG(n_coarse >= 1 && n_fine >= 1 && n_coarse + n_fine <= 2 * kMaxDraw, "coarse_fine_merge: too many samples");
  if (n_rays == 0) return NB2_OK;
  merge_kernel<<<grid_for(n_rays, kResampleWarps), 32 * kResampleWarps, 0, (cudaStream_t)stream>>>(
      rays, c_z, f_z, nullptr, n_rays, n_coarse, n_fine, z_out, pts_out, nullptr, nullptr);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_coarse_fine_merge_inds(nb2_handle* h, const float* rays, const float* c_z, const float* f_z, const int64_t* f_inds,
                                          int64_t n_rays, int n_coarse, int n_fine, float* z_out, float* pts_out, int64_t* all_inds_out,
                                          int64_t* sort_inds_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(rays && c_z && f_z && z_out, "coarse_fine_merge_inds: null pointer");
  NB2_CHECK_ARG(!all_inds_out || f_inds, "coarse_fine_merge_inds: all_inds_out needs f_inds");
  NB2_CHECK_ARG(n_coarse >= 1 && n_fine >= 1 && n_coarse + n_fine <= 2 * kMaxDraw, "coarse_fine_merge_inds: too many samples");
  merge_kernel<<<grid_for(n_rays, kResampleWarps), 32 * kResampleWarps, 0, (cudaStream_t)stream>>>(
      rays, c_z, f_z, f_inds, n_rays, n_coarse, n_fine, z_out, pts_out, all_inds_out, sort_inds_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

static int launch_composite(nb2_handle* h, const float* rgbo, const float* z, const float* dirs, int dir_stride, int64_t n_rays, int n_samples,
                            int flags, float near_t, float far_t, float* rgb_out, float* weights_out, float* depth_out, float* acc_out,
                            const float* aux, float* aux_out, cudaStream_t st) {
  NB2_CHECK_ARG(aligned16(rgbo), "composite: rgbo must be 16-byte aligned ([r,g,b,sigma] samples are read as float4)");
  if ((n_samples & 3) == 0 && aligned16(z) && (!weights_out || aligned16(weights_out)) && (!aux || aligned16(aux))) {
    int s4, lpr;
    regs_geometry(n_samples, s4, lpr);
    const int grid = grid_for(grid_for(n_rays, 32 / lpr), 8);
    if (s4 == 1)
      composite_regs_kernel<1><<<grid, 256, 0, st>>>(rgbo, z, dirs, dir_stride, n_rays, n_samples, flags, near_t, far_t, lpr, rgb_out, weights_out,
                                                     depth_out, acc_out, aux, aux_out);
    else
      composite_regs_kernel<2><<<grid, 256, 0, st>>>(rgbo, z, dirs, dir_stride, n_rays, n_samples, flags, near_t, far_t, lpr, rgb_out, weights_out,
                                                     depth_out, acc_out, aux, aux_out);
  } else {
    composite_kernel<<<grid_for(n_rays, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, st>>>(
        rgbo, z, dirs, dir_stride, n_rays, n_samples, flags, near_t, far_t, rgb_out, weights_out, depth_out, acc_out, aux, aux_out);
  }
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_composite(nb2_handle* h, const float* rgbo, const float* z, const float* dirs, int dir_stride,
                             int64_t n_rays, int n_samples, int flags, float near_t, float far_t, float* rgb_out,
                             float* weights_out, float* depth_out, float* acc_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(rgbo && z && dirs && rgb_out, "composite: null pointer");
  NB2_CHECK_ARG(dir_stride >= 3, "composite: dir_stride < 3");
  NB2_CHECK_ARG(n_samples >= 1 && n_samples <= kMaxSamples, "composite: n_samples must be in [1,%d]", kMaxSamples);
  if (n_rays == 0) return NB2_OK;
  return launch_composite(h, rgbo, z, dirs, dir_stride, n_rays, n_samples, flags, near_t, far_t, rgb_out, weights_out, depth_out, acc_out,
                          nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int nb2_composite_aux(nb2_handle* h, const float* rgbo, const float* z, const float* dirs, int dir_stride, int64_t n_rays,
                                 int n_samples, int flags, float near_t, float far_t, const float* aux, float* rgb_out, float* weights_out,
                                 float* depth_out, float* acc_out, float* aux_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(rgbo && z && dirs && rgb_out && aux && aux_out, "composite_aux: null pointer");
  NB2_CHECK_ARG(dir_stride >= 3, "composite_aux: dir_stride < 3");
  NB2_CHECK_ARG(n_samples >= 1 && n_samples <= kMaxSamples, "composite_aux: n_samples must be in [1,%d]", kMaxSamples);
  return launch_composite(h, rgbo, z, dirs, dir_stride, n_rays, n_samples, flags, near_t, far_t, rgb_out, weights_out, depth_out, acc_out,
                          aux, aux_out, (cudaStream_t)stream);
}

extern "C" int nb2_valid_sampler(nb2_handle* h, const float* rgbs, const int64_t* coords, const float* cam_tf,
                                 const int64_t* indices, const float* base_z, const float* jitter, float focal_x, float focal_y,
                                 float resolution, uint64_t seed, int64_t ray_offset, int64_t n_pixels, int64_t n_rays,
                                 int n_samples, float* pts_out, float* lengths_out, float* rgb_out, float* rays_out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(rgbs && coords && cam_tf && base_z && pts_out && lengths_out && rgb_out && rays_out, "valid_sampler: null pointer");
  NB2_CHECK_ARG(n_pixels > 0 && n_samples > 0 && focal_x != 0.f && focal_y != 0.f, "valid_sampler: bad arguments");
  valid_sampler_kernel<<<grid_for(n_rays * n_samples, 256), 256, 0, (cudaStream_t)stream>>>(
      rgbs, coords, cam_tf, indices, base_z, jitter, focal_x, focal_y, resolution, seed, ray_offset, n_pixels, n_rays, n_samples,
      pts_out, lengths_out, rgb_out, rays_out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_get_bounds(nb2_handle* h, const float* weights, const int64_t* inds, int64_t n_rays, int n_samples,
                              int n_inds, float* out, void* stream) {
  NB2_H(h);
  if (n_rays == 0) return NB2_OK;
  NB2_CHECK_ARG(weights && inds && out, "get_bounds: null pointer");
  NB2_CHECK_ARG(n_samples >= 1 && n_samples <= kMaxSamples && n_inds >= 2, "get_bounds: n_samples must be in [1,%d], n_inds >= 2", kMaxSamples);
  get_bounds_kernel<<<grid_for(n_rays, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(weights, inds, n_rays,
                                                                                                     n_samples, n_inds, out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_ide(nb2_handle* h, const float* xyz, const float* kappa_inv, int64_t n, const float* mat, const int* ml,
                       int n_pairs, int n_pow, float* out, void* stream) {
  NB2_H(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(xyz && kappa_inv && mat && ml && out && n > 0, "ide: bad arguments");
  NB2_CHECK_ARG(n_pairs >= 1 && n_pairs <= kIdeMaxPairs && n_pow >= 1 && n_pow <= kIdeMaxPow,
                "ide: at most %d (m, l) pairs and degree %d (deg_view <= 5, ref_func.py:67-68)", kIdeMaxPairs, kIdeMaxPow - 1);
  const size_t smem = (size_t)(kIdeMaxPow * kIdeMaxPairs + 2 * kIdeMaxPairs + kIdeBlockPts * 2 * n_pairs) * sizeof(float);
  {
    const int rc = kernel_set_smem(h, (const void*)ide_kernel, 64 * 1024);
    if (rc != NB2_OK) return rc;
  }
  ide_kernel<<<grid_for(n, kIdeBlockPts), kIdeBlockPts, smem, (cudaStream_t)stream>>>(xyz, kappa_inv, n, n_pairs, n_pow, mat, ml, out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_linear_to_srgb(nb2_handle* h, const float* linear, int64_t n, float* out, void* stream) {
  NB2_H(h);
  if (n == 0) return NB2_OK;
  NB2_CHECK_ARG(linear && out && n > 0, "linear_to_srgb: bad arguments");
  const int blocks = (int)std::min<int64_t>(grid_for(n, 256), 148 * 16);
  linear_to_srgb_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(linear, n, out);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}
