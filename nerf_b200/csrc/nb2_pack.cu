// nb2_pack.cu — nn.Linear parameters (fp32, (out,in) row-major; reference state_dict order)
// -> the operand images the MLP kernels stream:
//   * tcgen05 path: 128(n) x 64(k) bf16 tiles, K-major, 128-byte swizzle, hi and lo halves
//     (w = hi + lo), stored in exactly the order the MMA issuer consumes them, so the weight
//     streamer is a sequence of contiguous 16 KB cp.async.bulk copies;
//   * CUDA-core path: fp32 transposes Wt[k][n] (K-major) so a warp reads one coalesced line per k;
//   * biases and the tiny heads (sigma 256->1, rgb 128->3) stay fp32.
// The concatenations of the reference (skip: cat(enc, h) mip_model.py:55; rgb: cat(bottleneck,
// enc_dir) mip_model.py:59) become K-chunk lists, so no activation is ever concatenated.
#include <algorithm>
#include <vector>

#include <cuda_fp16.h>

#include "nb2_common.cuh"
#include "nb2_tc_ptx.cuh"

namespace nb2 {

struct PackChunkDesc {
  const float* W;     // may be null for a bias-only chunk
  const float* bias;  // if non-null, tile column 63 carries bias[n0 + n] (the A operand's column 63 is 1)
  int ld;             // in_features of the source matrix
  int n0;             // first output row of this tile
  int col_off;        // first source column
  int valid_cols;     // columns past this are zero padding
};
struct PackSimtDesc {
  const float* W;
  int ld, n;
  int k0, off0, valid0;
  int k1, off1, valid1;
  int wt_off;
};

__global__ void pack_chunks_kernel(const PackChunkDesc* __restrict__ descs, __nv_bfloat16* __restrict__ out) {
  const PackChunkDesc d = descs[blockIdx.x];
  // per chunk: [bf16 hi][bf16 lo][fp16 hi][fp16 lo], 16 KB each
  __nv_bfloat16* bhi = out + ((size_t)blockIdx.x * 4 + 0) * (kTileBytes / 2);
  __nv_bfloat16* blo = out + ((size_t)blockIdx.x * 4 + 1) * (kTileBytes / 2);
  __half* hhi = reinterpret_cast<__half*>(out + ((size_t)blockIdx.x * 4 + 2) * (kTileBytes / 2));
  __half* hlo = reinterpret_cast<__half*>(out + ((size_t)blockIdx.x * 4 + 3) * (kTileBytes / 2));
  for (int i = threadIdx.x; i < kTileRows * kTileCols; i += blockDim.x) {
    const int n = i / kTileCols, k = i % kTileCols;
    float w = (d.W != nullptr && k < d.valid_cols) ? d.W[(size_t)(d.n0 + n) * d.ld + d.col_off + k] : 0.f;
    if (d.bias != nullptr && k == kTileCols - 1) w = d.bias[d.n0 + n];
    const uint32_t o = ptx::swz128_offset(n, k) / 2;
    const __nv_bfloat16 bh = __float2bfloat16_rn(w);
    bhi[o] = bh;
    blo[o] = __float2bfloat16_rn(w - __bfloat162float(bh));
    const __half hh = __float2half_rn(w);
    hhi[o] = hh;
    hlo[o] = __float2half_rn(w - __half2float(hh));
  }
}

__global__ void pack_simt_kernel(const PackSimtDesc* __restrict__ descs, float* __restrict__ wt32) {
  const PackSimtDesc d = descs[blockIdx.y];
  const int total = (d.k0 + d.k1) * d.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / d.n, n = i % d.n;
    float w = 0.f;
    if (k < d.k0) {
      if (k < d.valid0) w = d.W[(size_t)n * d.ld + d.off0 + k];
    } else {
      const int kk = k - d.k0;
      if (kk < d.valid1) w = d.W[(size_t)n * d.ld + d.off1 + kk];
    }
    wt32[d.wt_off + i] = w;
  }
}

// One MMA layer = nc N-chunks x kc K-chunks of weight tiles.  The bias is folded into the K-chunk that reads the
// encoding tile (column 63); layers without one get an extra bias-only chunk whose MMA runs a single k-step.
static void add_tc_layer(TcNet& net, std::vector<PackChunkDesc>& chunks, const float* W, const float* bias, int ld,
                         int kc, int nc, const int* a_src, const int* col_off, const int* valid, int epi, int bias_off) {
  TcLayer& L = net.layer[net.n_layers++];
  int e_chunk = -1;
  for (int k = 0; k < kc; ++k) {
    L.a_src[k] = a_src[k];
    L.ks0[k] = 0;
    if (a_src[k] == kChunkE) e_chunk = k;
  }
  const bool extra = (e_chunk < 0);
  if (extra) {
    L.a_src[kc] = kChunkE;
    L.ks0[kc] = 3;
  }
  L.kc = kc + (extra ? 1 : 0);
  L.nc = nc;
  for (int k = L.kc; k < 5; ++k) { L.a_src[k] = 0; L.ks0[k] = 0; }
  L.epi = epi;
  L.bias_off = bias_off;
  L.chunk0 = (int)chunks.size();
  for (int n = 0; n < nc; ++n) {
    for (int k = 0; k < kc; ++k)
      chunks.push_back({W, (k == e_chunk) ? bias : nullptr, ld, n * kTileRows, col_off[k], valid[k]});
    if (extra) chunks.push_back({nullptr, bias, ld, n * kTileRows, 0, 0});
  }
  net.n_chunks = (int)chunks.size();
}

static void add_simt_layer(SimtNet& net, std::vector<PackSimtDesc>& descs, int& wt_off, const float* W, int ld, int n,
                           int k0, int src0, int off0, int valid0, int k1, int src1, int off1, int valid1, int epi,
                           int bias_off) {
  SimtLayer& L = net.layer[net.n_layers++];
  L.k0 = k0; L.k1 = k1; L.src0 = src0; L.src1 = src1; L.n = n; L.epi = epi; L.wt_off = wt_off; L.bias_off = bias_off;
  descs.push_back({W, ld, n, k0, off0, valid0, k1, off1, valid1, wt_off});
  wt_off += (k0 + k1) * n;
}

int pack_network(nb2_handle* h, int net_id, const float* const* W, const float* const* b, int n_layers,
                 int pos_levels, int dir_levels, cudaStream_t st) {
  PackedNet& pn = h->net[net_id];
  const int enc = 3 + 6 * pos_levels;  // cat_origin = True
  const int dir = 3 + 6 * dir_levels;
  const int H4[4] = {kChunkH0, kChunkH0 + 1, kChunkH0 + 2, kChunkH0 + 3};
  const int off4[4] = {0, 64, 128, 192};
  const int v4[4] = {64, 64, 64, 64};

  TcNet tc;
  memset(&tc, 0, sizeof(tc));
  SimtNet simt;
  memset(&simt, 0, sizeof(simt));
  std::vector<PackChunkDesc> chunks;
  std::vector<PackSimtDesc> sdescs;
  int wt_off = 0;
  struct Copy { float* dst; const float* src; int n; };
  std::vector<Copy> copies;
  const int n_bias_floats = 2048 + kRgbHidden;

  // first use: allocate this network's device buffers (worst-case sizes, fixed by the architecture)
  if (!pn.d_wchunks) {
    const size_t max_chunks = 80;
    const size_t max_wt = (size_t)64 * 256 + 7 * 256 * 256 + 64 * 256 + 288 * 128;
    NB2_CUDA(cudaMalloc(&pn.d_wchunks, max_chunks * 4 * kTileBytes));
    NB2_CUDA(cudaMalloc(&pn.d_bias, n_bias_floats * sizeof(float)));
    NB2_CUDA(cudaMalloc(&pn.d_head, kHeadFloats * sizeof(float)));
    NB2_CUDA(cudaMalloc(&pn.d_wt32, max_wt * sizeof(float)));
  }
  NB2_CUDA(cudaMemsetAsync(pn.d_head, 0, kHeadFloats * sizeof(float), st));
  NB2_CUDA(cudaMemsetAsync(pn.d_bias, 0, n_bias_floats * sizeof(float), st));

  // a failed (re-)pack must not leave a half-written image marked valid
  pn.packed = false;
  if (pn.kind == NB2_NET_PROPOSAL) {
    NB2_CHECK_ARG(n_layers == 5, "pack_weights: the proposal network has 5 linear layers, got %d", n_layers);
    const int eE[1] = {kChunkE}, eo[1] = {0}, ev[1] = {enc};
    add_tc_layer(tc, chunks, W[0], b[0], enc, 1, 2, eE, eo, ev, EPI_RELU, 0);
    add_simt_layer(simt, sdescs, wt_off, W[0], enc, kHidden, kEncCols, 0, 0, enc, 0, 0, 0, 0, EPI_RELU, 0);
    copies.push_back({pn.d_bias + 0, b[0], kHidden});
    for (int l = 1; l <= 3; ++l) {
      const int epi = (l == 3) ? EPI_SIGMA_OUT : EPI_RELU;
      add_tc_layer(tc, chunks, W[l], b[l], kHidden, 4, 2, H4, off4, v4, epi, l * kHidden);
      add_simt_layer(simt, sdescs, wt_off, W[l], kHidden, kHidden, kHidden, 1, 0, kHidden, 0, 0, 0, 0, epi, l * kHidden);
      copies.push_back({pn.d_bias + l * kHidden, b[l], kHidden});
    }
    copies.push_back({pn.d_head + kHeadSigmaW, W[4], kHidden});
    copies.push_back({pn.d_head + kHeadSigmaB, b[4], 1});
  } else {
    NB2_CHECK_ARG(n_layers == 11, "pack_weights: the NeRF network has 11 linear layers, got %d", n_layers);
    // lin_block1                                         mip_model.py:19-23
    const int eE[1] = {kChunkE}, eo[1] = {0}, ev[1] = {enc};
    add_tc_layer(tc, chunks, W[0], b[0], enc, 1, 2, eE, eo, ev, EPI_RELU, 0);
    add_simt_layer(simt, sdescs, wt_off, W[0], enc, kHidden, kEncCols, 0, 0, enc, 0, 0, 0, 0, EPI_RELU, 0);
    for (int l = 1; l <= 3; ++l) {
      add_tc_layer(tc, chunks, W[l], b[l], kHidden, 4, 2, H4, off4, v4, EPI_RELU, l * kHidden);
      add_simt_layer(simt, sdescs, wt_off, W[l], kHidden, kHidden, kHidden, 1, 0, kHidden, 0, 0, 0, 0, EPI_RELU, l * kHidden);
    }
    // lin_block2.0 on cat(enc, h)                        mip_model.py:24-27,55
    {
      const int src[5] = {kChunkE, kChunkH0, kChunkH0 + 1, kChunkH0 + 2, kChunkH0 + 3};
      const int off[5] = {0, enc, enc + 64, enc + 128, enc + 192};
      const int val[5] = {enc, 64, 64, 64, 64};
      add_tc_layer(tc, chunks, W[4], b[4], enc + kHidden, 5, 2, src, off, val, EPI_RELU, 4 * kHidden);
      add_simt_layer(simt, sdescs, wt_off, W[4], enc + kHidden, kHidden, kEncCols, 0, 0, enc, kHidden, 1, enc, kHidden,
                     EPI_RELU, 4 * kHidden);
    }
    add_tc_layer(tc, chunks, W[5], b[5], kHidden, 4, 2, H4, off4, v4, EPI_RELU, 5 * kHidden);
    add_simt_layer(simt, sdescs, wt_off, W[5], kHidden, kHidden, kHidden, 1, 0, kHidden, 0, 0, 0, 0, EPI_RELU, 5 * kHidden);
    // lin_block2.4 feeds both opacity_head and bottle_neck   mip_model.py:56-58
    add_tc_layer(tc, chunks, W[6], b[6], kHidden, 4, 2, H4, off4, v4, EPI_RELU_SIGMA, 6 * kHidden);
    add_simt_layer(simt, sdescs, wt_off, W[6], kHidden, kHidden, kHidden, 1, 0, kHidden, 0, 0, 0, 0, EPI_RELU_SIGMA, 6 * kHidden);
    add_tc_layer(tc, chunks, W[7], b[7], kHidden, 4, 2, H4, off4, v4, EPI_LINEAR, 7 * kHidden);
    add_simt_layer(simt, sdescs, wt_off, W[7], kHidden, kHidden, kHidden, 1, 0, kHidden, 0, 0, 0, 0, EPI_LINEAR, 7 * kHidden);
    // rgb_layer.0 on cat(bottleneck, enc_dir)            mip_model.py:34-37,59
    {
      const int src[5] = {kChunkH0, kChunkH0 + 1, kChunkH0 + 2, kChunkH0 + 3, kChunkE};
      const int off[5] = {0, 64, 128, 192, 256};
      const int val[5] = {64, 64, 64, 64, dir};
      add_tc_layer(tc, chunks, W[9], b[9], kHidden + dir, 5, 1, src, off, val, EPI_RGB, 8 * kHidden);
      add_simt_layer(simt, sdescs, wt_off, W[9], kHidden + dir, kRgbHidden, kHidden, 1, 0, kHidden, kDirCols, 2, kHidden,
                     dir, EPI_RGB, 8 * kHidden);
    }
    const int bsrc[9] = {0, 1, 2, 3, 4, 5, 6, 7, 9};
    for (int l = 0; l < 9; ++l) copies.push_back({pn.d_bias + l * kHidden, b[bsrc[l]], l == 8 ? kRgbHidden : kHidden});
    copies.push_back({pn.d_head + kHeadSigmaW, W[8], kHidden});
    copies.push_back({pn.d_head + kHeadSigmaB, b[8], 1});
    copies.push_back({pn.d_head + kHeadRgbW, W[10], 3 * kRgbHidden});
    copies.push_back({pn.d_head + kHeadRgbB, b[10], 3});
  }

  // descriptor tables travel through a small device scratch that lives until the kernels ran
  void* d_desc = nullptr;
  const size_t cbytes = chunks.size() * sizeof(PackChunkDesc), sbytes = sdescs.size() * sizeof(PackSimtDesc);
  NB2_CUDA(cudaMallocAsync(&d_desc, cbytes + sbytes, st));
  // every exit below releases the scratch (stream-ordered, so kernels already enqueued still see it)
  auto run = [&]() -> int {
    NB2_CUDA(cudaMemcpyAsync(d_desc, chunks.data(), cbytes, cudaMemcpyHostToDevice, st));
    NB2_CUDA(cudaMemcpyAsync((char*)d_desc + cbytes, sdescs.data(), sbytes, cudaMemcpyHostToDevice, st));
    // pageable sources are staged before cudaMemcpyAsync returns; make that explicit
    NB2_CUDA(cudaStreamSynchronize(st));
    pack_chunks_kernel<<<(int)chunks.size(), 256, 0, st>>>((const PackChunkDesc*)d_desc, pn.d_wchunks);
    NB2_LAUNCH_CHECK(h);
    pack_simt_kernel<<<dim3(64, (int)sdescs.size()), 256, 0, st>>>((const PackSimtDesc*)((char*)d_desc + cbytes), pn.d_wt32);
    NB2_LAUNCH_CHECK(h);
    for (const Copy& c : copies)
      NB2_CUDA(cudaMemcpyAsync(c.dst, c.src, c.n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return NB2_OK;
  };
  const int rc = run();
  cudaFreeAsync(d_desc, st);
  if (rc != NB2_OK) return rc;

  pn.tc = tc;
  pn.simt = simt;
  pn.pos_levels = pos_levels;
  pn.dir_levels = dir_levels;
  pn.packed = true;
  pn.version++;
  return NB2_OK;
}

}  // namespace nb2
