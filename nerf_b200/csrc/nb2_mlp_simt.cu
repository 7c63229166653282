// nb2_mlp_simt.cu — NB2_PREC_FP32: the proposal / NeRF MLP evaluated with fp32 FFMA on the CUDA
// cores.  This is the strict-arithmetic mode (same number format as the reference's fp32
// nn.Linear path) and the on-device yardstick the tensor-core kernel is checked against.
//
// One CTA (256 threads) owns a tile of 64 rows.  Activations live transposed in shared memory
// (X^T[k][row], so the 8 rows of a warp are one broadcast 2 x LDS.128), weights are read
// K-major from the packed fp32 transposes (one coalesced 512 B line per warp per k).  Each
// thread keeps an 8-row x 8-column accumulator block.
#include "nb2_common.cuh"
#include "nb2_rowio.cuh"

namespace nb2 {

constexpr int kSimtRows = 64;
constexpr int kSimtThreads = 256;

struct SimtSmem {
  float E[kEncCols][kSimtRows];       // encoded position, transposed
  float D[kDirCols][kSimtRows];       // encoded direction, transposed
  float H[2][kHidden][kSimtRows];     // ping-pong hidden activations, transposed
  float sigma[kSimtRows];
};

template <int NG>  // NG column groups of 4 per lane: 2 -> N = 256, 1 -> N = 128
__device__ __forceinline__ void simt_accumulate(float (&acc)[8][4 * NG], const float* __restrict__ xt /*[k][64]*/,
                                                const float* __restrict__ wt /*[k][n]*/, int K, int n, int r0, int lane) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float4 a0 = *reinterpret_cast<const float4*>(xt + k * kSimtRows + r0);
    float4 a1 = *reinterpret_cast<const float4*>(xt + k * kSimtRows + r0 + 4);
    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float w[4 * NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      float4 wv = __ldg(reinterpret_cast<const float4*>(wt + (size_t)k * n + g * 128 + lane * 4));
      w[4 * g + 0] = wv.x; w[4 * g + 1] = wv.y; w[4 * g + 2] = wv.z; w[4 * g + 3] = wv.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4 * NG; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
}

__device__ __forceinline__ const float* simt_src(const SimtSmem& sm, int src, int cur) {
  return src == 0 ? &sm.E[0][0] : (src == 2 ? &sm.D[0][0] : &sm.H[cur][0][0]);
}

__global__ void __launch_bounds__(kSimtThreads, 1)
mlp_simt_kernel(SimtNet net, MlpIo io, const float* __restrict__ wt32, const float* __restrict__ bias,
                const float* __restrict__ head, int pos_levels, int dir_levels, int has_dir) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SimtSmem& sm = *reinterpret_cast<SimtSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r0 = warp * 8;
  const int64_t n_tiles = (io.n_rows + kSimtRows - 1) / kSimtRows;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row_base = tile * kSimtRows;
    // ---- encode: thread (row, quarter) fills a quarter of the encoded columns ------------
    {
      const int row = tid & (kSimtRows - 1), part = tid / kSimtRows;  // part in [0,4)
      RowIn in = load_row(io, row_base + row);
      for (int c = part; c < kEncCols; c += 4) sm.E[c][row] = in.valid ? enc_column(in.p, c, pos_levels, in.enc, in.ipe ? in.cov : nullptr) : 0.f;
      if (has_dir) {
        float rot[3] = {0.f, 0.f, 0.f};
        if (in.valid) normalize_dir(in.d, rot);
        for (int c = part; c < kDirCols; c += 4) sm.D[c][row] = in.valid ? enc_column(rot, c, dir_levels) : 0.f;
      }
    }
    __syncthreads();

    int cur = 0;  // H buffer that holds the current layer's input
    for (int l = 0; l < net.n_layers; ++l) {
      const SimtLayer L = net.layer[l];
      const float* wt = wt32 + L.wt_off;
      const float* b = bias + L.bias_off;
      const int nxt = cur ^ 1;
      if (L.n == kHidden) {
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        simt_accumulate<2>(acc, simt_src(sm, L.src0, cur), wt, L.k0, L.n, r0, lane);
        if (L.k1 > 0) simt_accumulate<2>(acc, simt_src(sm, L.src1, cur), wt + (size_t)L.k0 * L.n, L.k1, L.n, r0, lane);
        float sig[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) sig[i] = 0.f;
        const bool want_sigma = (L.epi == EPI_RELU_SIGMA || L.epi == EPI_SIGMA_OUT);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = (j >> 2) * 128 + lane * 4 + (j & 3);
          const float bj = __ldg(b + col);
          const float ws = want_sigma ? __ldg(head + kHeadSigmaW + col) : 0.f;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float t = acc[i][j] + bj;
            if (L.epi != EPI_LINEAR) t = fmaxf(t, 0.f);
            v[i] = t;
            sig[i] = fmaf(t, ws, sig[i]);
          }
          float* dst = &sm.H[nxt][col][r0];
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (want_sigma) {
          const float bs = __ldg(head + kHeadSigmaB);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float s = warp_sum(sig[i]) + bs;
            if (lane == 0) {
              if (L.epi == EPI_SIGMA_OUT) {
                int64_t row = row_base + r0 + i;
                if (row < io.n_rows) io.out[row] = s;
              } else {
                sm.sigma[r0 + i] = s;
              }
            }
          }
        }
      } else {  // rgb_layer: N = 128, then the 128 -> 3 sigmoid head
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        simt_accumulate<1>(acc, simt_src(sm, L.src0, cur), wt, L.k0, L.n, r0, lane);
        if (L.k1 > 0) simt_accumulate<1>(acc, simt_src(sm, L.src1, cur), wt + (size_t)L.k0 * L.n, L.k1, L.n, r0, lane);
        float rgb[8][3];
#pragma unroll
        for (int i = 0; i < 8; ++i) rgb[i][0] = rgb[i][1] = rgb[i][2] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = lane * 4 + j;
          const float bj = __ldg(b + col);
          const float w0 = __ldg(head + kHeadRgbW + col), w1 = __ldg(head + kHeadRgbW + 128 + col),
                      w2 = __ldg(head + kHeadRgbW + 256 + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float t = fmaxf(acc[i][j] + bj, 0.f);
            rgb[i][0] = fmaf(t, w0, rgb[i][0]);
            rgb[i][1] = fmaf(t, w1, rgb[i][1]);
            rgb[i][2] = fmaf(t, w2, rgb[i][2]);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float c0 = warp_sum(rgb[i][0]) + __ldg(head + kHeadRgbB + 0);
          float c1 = warp_sum(rgb[i][1]) + __ldg(head + kHeadRgbB + 1);
          float c2 = warp_sum(rgb[i][2]) + __ldg(head + kHeadRgbB + 2);
          int64_t row = row_base + r0 + i;
          if (lane == 0 && row < io.n_rows) {
            float4 o = make_float4(1.f / (1.f + expf(-c0)), 1.f / (1.f + expf(-c1)), 1.f / (1.f + expf(-c2)),
                                   sm.sigma[r0 + i]);
            reinterpret_cast<float4*>(io.out)[row] = o;
          }
        }
      }
      cur = nxt;
      __syncthreads();
    }
  }
}

int launch_mlp_simt(nb2_handle* h, int net_id, const MlpIo& io, cudaStream_t st) {
  PackedNet& pn = h->net[net_id];
  if (!pn.packed) {
    set_error("mlp_forward: weights of network %d have not been packed (call nb2_pack_weights)", net_id);
    return NB2_ERR_STATE;
  }
  if (io.out_mode == 2) {
    set_error("mlp_forward: fused compositing is not available in NB2_PREC_FP32");
    return NB2_ERR_UNSUPPORTED;
  }
  if (io.n_rows == 0) return NB2_OK;
  {
    const int rc = kernel_set_smem(h, (const void*)mlp_simt_kernel, (int)sizeof(SimtSmem));
    if (rc != NB2_OK) return rc;
  }
  int64_t n_tiles = (io.n_rows + kSimtRows - 1) / kSimtRows;
  int grid = (int)std::min<int64_t>(n_tiles, (int64_t)h->sm_count);
  mlp_simt_kernel<<<grid, kSimtThreads, sizeof(SimtSmem), st>>>(pn.simt, io, pn.d_wt32, pn.d_bias, pn.d_head,
                                                                pn.pos_levels, pn.dir_levels, pn.kind == NB2_NET_NERF);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

}  // namespace nb2
