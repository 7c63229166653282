// nb2_mlp_tc.cu — the per-sample MLP (proposal 4x256 / NeRF 8x256 + heads) as ONE persistent,
// warp-specialised tcgen05 kernel.  NB2_PREC_BF16 / NB2_PREC_FP16 (single pass) and NB2_PREC_FP16X3 / NB2_PREC_BF16X3 (split).
//
//   warp 0      weight streamer: cp.async.bulk (TMA unit) of pre-swizzled 128x64 bf16 weight
//               tiles from L2 into a 4-stage shared-memory ring, mbarrier full/empty handshake
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=128, K=16) with both
//               operands in shared memory and the fp32 accumulator in TMEM
//   warp 2      TMEM allocator
//   warps 4..   one 128-thread "slot group" per resident 128-row tile: it produces the tile's
//               first A operand (sample point -> sinusoidal encoding -> swizzled bf16 rows),
//               and after every layer drains the accumulator from TMEM (tcgen05.ld), applies
//               bias / ReLU, re-quantises to bf16 (hi [+ lo]) and writes the next layer's A
//               operand in place.  The last epilogue evaluates the 128->3 / 256->1 heads on the
//               fp32 values and either stores rgb-sigma or alpha-composites the ray.
//
// single pass (bf16 | fp16): two slots ping-pong, so one tile's epilogue overlaps the other tile's MMAs.
// split (fp16x3 | bf16x3)   : every operand is split x = hi + lo (both 16-bit); each product is evaluated
//                  as hi*hi + lo*hi + hi*lo with fp32 accumulation (fp16: 22-bit operands, ~2^-22 relative
//                  per product, i.e. fp32-faithful; bf16: ~2^-16), one slot (the hi/lo activation pair
//                  fills the shared memory of two slots).
//
// Shared memory (bytes):  activations NSLOTS * (SPLIT ? 2 : 1) * 5 * 16 KB = 160 KB,
//                         weight ring 4 * 16 KB = 64 KB, barriers + scratch < 1 KB.
// TMEM: 512 columns; slot s owns columns [256 s, 256 s + 256) (128 lanes x 256 fp32).  Split mode has one
//       slot and uses columns [256, 512) as a second accumulator for the hi*lo + lo*hi cross terms, so the
//       main accumulator sees a third of the (truncating) tensor-core additions; the two are summed in fp32
//       by the epilogue.
// Biases ride on the tensor core: encoding column 63 is the constant 1 and the matching weight column is
// the bias (layers without an encoding K-chunk get a one-k-step "bias chunk").
#include <stdlib.h>

#include "nb2_common.cuh"
#include "nb2_rowio.cuh"
#include "nb2_tc_ptx.cuh"

namespace nb2 {
using namespace ptx;

constexpr int kStages = 4;
constexpr int kRolesThreads = 128;  // warps 0..3

struct TcParams {
  TcNet net;
  MlpIo io;
  const __nv_bfloat16* wchunks;
  const float* bias;
  const float* head;
  int pos_levels, dir_levels, has_dir;
  int cluster;   // CTAs per cluster sharing every weight tile through multicast bulk copies (1, 2 or 4)
  int64_t n_tiles;
};

struct TcMisc {
  uint64_t w_full[kStages];
  uint64_t w_empty[kStages];
  uint64_t a_ready[2];
  uint64_t acc_full[2];
  uint32_t tmem_base;
  uint32_t pad;
  float scratch[2][4][8];  // per slot, per warp: cross-warp scan / reduction staging
};

template <int NSLOTS, bool SPLIT>
struct TcLayout {
  static constexpr int kActTiles = kChunksPerSlot * (SPLIT ? 2 : 1);
  static constexpr int kSlotBytes = kActTiles * kTileBytes;
  static constexpr int kActBytes = NSLOTS * kSlotBytes;
  static constexpr int kRingBytes = kStages * kTileBytes;
  static constexpr int kMiscBytes = 1024;
  static constexpr int kTotal = kActBytes + kRingBytes + kMiscBytes + 1024 /* alignment slack */;
  static_assert(sizeof(TcMisc) <= kMiscBytes, "misc region too small");
  static_assert(kTotal <= 232448, "exceeds 227 KB of shared memory");
};

// ---- writing one row of an A-operand tile ---------------------------------------------------
// v[0..8) are 8 consecutive columns starting at column `col` (multiple of 8) of row `row`.
template <bool SPLIT, bool F16>
__device__ __forceinline__ void store_a8(uint32_t tile_hi, uint32_t tile_lo, int row, int col, const float (&v)[8]) {
  const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)col >> 3) ^ ((uint32_t)row & 7u)) << 4);
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = pack16x2<F16>(v[2 * i], v[2 * i + 1]);
  st_shared_v4(tile_hi + off, h[0], h[1], h[2], h[3]);
  if (SPLIT) {
    uint32_t l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) l[i] = residual16x2<F16>(v[2 * i], v[2 * i + 1], h[i]);
    st_shared_v4(tile_lo + off, l[0], l[1], l[2], l[3]);
  }
}

// Encoded position / direction row -> E tile.  NCOLS = 64 (position) or 32 (direction).
// Column 63 of the position row is the constant 1: the matching weight column carries the layer bias,
// so biases are added by the tensor core (see nb2_pack.cu) and never touch the epilogue.
template <bool SPLIT, bool F16, int NCOLS, int MAXLEV>
__device__ __forceinline__ void write_enc_row(uint32_t tile_hi, uint32_t tile_lo, int row, const float x[3],
                                              int levels, bool valid) {
  float v[NCOLS];
#pragma unroll
  for (int c = 0; c < NCOLS; ++c) v[c] = 0.f;
  if (valid) {
    v[0] = x[0]; v[1] = x[1]; v[2] = x[2];
#pragma unroll
    for (int l = 0; l < MAXLEV; ++l) {
      if (l < levels) {
        const float sc = (float)(1 << l);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float s, c;
          sincosf(x[k] * sc, &s, &c);
          v[3 + 6 * l + k] = s;
          v[3 + 6 * l + 3 + k] = c;
        }
      }
    }
  }
  if (NCOLS == 64) v[NCOLS - 1] = 1.f;
#pragma unroll
  for (int g = 0; g < NCOLS / 8; ++g) {
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = v[8 * g + i];
    store_a8<SPLIT, F16>(tile_hi, tile_lo, row, 8 * g, w);
  }
}

// ---- hidden-layer epilogue: TMEM accumulator -> act -> 16-bit A operand of the next layer (in place) ----
// EPI: EPI_RELU / EPI_LINEAR / EPI_RELU_SIGMA / EPI_SIGMA_OUT.  Returns the density-head dot product
// (without its bias) for the *_SIGMA kinds.  In SPLIT mode the cross terms live in a second accumulator
// (columns +256) and are added here in fp32.
template <int EPI, bool SPLIT, bool F16>
__device__ __forceinline__ float epilogue_hidden(uint32_t acc, uint32_t slot_base, uint32_t lo_off, int row,
                                                 const float* __restrict__ head) {
  constexpr bool kRelu = (EPI != EPI_LINEAR);
  constexpr bool kSigma = (EPI == EPI_RELU_SIGMA || EPI == EPI_SIGMA_OUT);
  constexpr bool kStore = (EPI != EPI_SIGMA_OUT);
  float sg = 0.f;
#pragma unroll 1
  for (int cb = 0; cb < kHidden / 32; ++cb) {
    uint32_t r[32];
    tmem_ld32(acc + cb * 32, r);
    if (SPLIT) {
      uint32_t c[32];
      tmem_ld32(acc + 256 + cb * 32, c);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(c[j]));
    } else {
      tmem_ld_wait();
    }
    const uint32_t h_hi = slot_base + (uint32_t)(cb >> 1) * kTileBytes;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = cb * 32 + g * 8;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
      if (kRelu && (SPLIT || kSigma)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (kSigma) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(head + kHeadSigmaW + col));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(head + kHeadSigmaW + col + 4));
        sg = fmaf(v[0], w0.x, sg); sg = fmaf(v[1], w0.y, sg); sg = fmaf(v[2], w0.z, sg); sg = fmaf(v[3], w0.w, sg);
        sg = fmaf(v[4], w1.x, sg); sg = fmaf(v[5], w1.y, sg); sg = fmaf(v[6], w1.z, sg); sg = fmaf(v[7], w1.w, sg);
      }
      if (kStore) {
        if (SPLIT || kSigma || !kRelu) {
          store_a8<SPLIT, F16>(h_hi, h_hi + lo_off, row, col & 63, v);
        } else {
          // single pass: relu fused into the fp32 -> 16-bit conversion
          const uint32_t off = (uint32_t)row * 128u + (((((uint32_t)col & 63u) >> 3) ^ ((uint32_t)row & 7u)) << 4);
          uint32_t h[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = pack16x2_relu<F16>(v[2 * i], v[2 * i + 1]);
          st_shared_v4(h_hi + off, h[0], h[1], h[2], h[3]);
        }
      }
    }
  }
  return sg;
}

// ---- the kernel ----------------------------------------------------------------------------------
// LOCKSTEP (two-slot mode only): both resident tiles consume every weight tile back to back, halving the
// L2 -> SM weight traffic per FLOP at the price of not overlapping one tile's epilogue with the other's MMAs.
template <int NSLOTS, bool SPLIT, bool F16, bool LOCKSTEP>
__global__ void __launch_bounds__(kRolesThreads + 128 * NSLOTS, 1) mlp_tc_kernel(const __grid_constant__ TcParams p) {
  using LT = TcLayout<NSLOTS, SPLIT>;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t act_base = smem_base;
  const uint32_t ring_base = smem_base + LT::kActBytes;
  TcMisc* misc = reinterpret_cast<TcMisc*>(smem_al + LT::kActBytes + LT::kRingBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcNet& net = p.net;
  const int64_t tiles_per_iter = (int64_t)gridDim.x * NSLOTS;
  const int64_t n_iters = (p.n_tiles + tiles_per_iter - 1) / tiles_per_iter;   // identical for every CTA: the
  // CTAs of a cluster walk the same (iteration, layer, slot) schedule; tiles past n_tiles run on zero rows
  const uint32_t cl_size = (uint32_t)p.cluster;
  const uint32_t cl_rank = cl_size > 1 ? cluster_ctarank() : 0u;
  const uint16_t cl_mask = (uint16_t)((1u << cl_size) - 1u);

  // ---- one-time setup ---------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&misc->w_full[i]), 1);
      mbar_init(smem_u32(&misc->w_empty[i]), cl_size);   // one arrive per consumer CTA of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&misc->a_ready[s]), 128);
      mbar_init(smem_u32(&misc->acc_full[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&misc->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (cl_size > 1) cluster_sync_all();   // every CTA's barriers exist before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = misc->tmem_base;

  if (warp == 0) {
    // =========================== weight streamer ==================================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, q = 0;
      for (int64_t it = 0; it < n_iters; ++it) {
        for (int l = 0; l < net.n_layers; ++l) {
          const int n_chunks = net.layer[l].kc * net.layer[l].nc;
          const int chunk0 = net.layer[l].chunk0;
          for (int s = 0; s < (LOCKSTEP ? 1 : NSLOTS); ++s) {
            for (int c = 0; c < n_chunks; ++c) {
#pragma unroll
              for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
                mbar_wait(smem_u32(&misc->w_empty[stage]), phase ^ 1u);   // all consumers of the cluster released it
                const uint32_t full = smem_u32(&misc->w_full[stage]);
                mbar_arrive_expect_tx(full, kTileBytes);
                const __nv_bfloat16* src = p.wchunks + ((size_t)(chunk0 + c) * 4 + (F16 ? 2 : 0) + part) * (kTileBytes / 2);
                if (cl_size == 1) bulk_g2s(ring_base + stage * kTileBytes, src, kTileBytes, full);
                else if (q % cl_size == cl_rank) bulk_g2s_mcast(ring_base + stage * kTileBytes, src, kTileBytes, full, cl_mask);
                ++q;
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =======================================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_16(128, 128, F16);
      uint32_t stage = 0, phase = 0;
      uint32_t pa[2] = {0u, 0u};
      auto release = [&](uint32_t bar) {
        if (cl_size == 1) umma_commit(bar); else umma_commit_mcast(bar, cl_mask);
      };
      for (int64_t it = 0; it < n_iters; ++it) {
        for (int l = 0; l < net.n_layers; ++l) {
          const TcLayer& L = net.layer[l];
          if (LOCKSTEP) {
            // both tiles' operands are ready -> every weight tile is used twice
            for (int s = 0; s < NSLOTS; ++s) {
              mbar_wait(smem_u32(&misc->a_ready[s]), pa[s]);
              pa[s] ^= 1u;
            }
            tc_fence_after();
            for (int n = 0; n < L.nc; ++n) {
              for (int k = 0; k < L.kc; ++k) {
                const int ks0 = L.ks0[k];
                mbar_wait(smem_u32(&misc->w_full[stage]), phase);
                tc_fence_after();
                const uint32_t w_hi = ring_base + stage * kTileBytes;
                for (int s = 0; s < NSLOTS; ++s) {
                  const uint32_t a_hi = act_base + s * LT::kSlotBytes + (uint32_t)L.a_src[k] * kTileBytes;
                  const uint32_t d_main = tmem_base + (uint32_t)(s * 256) + n * 128;
                  for (int ks = ks0; ks < 4; ++ks)
                    umma_bf16_ss(d_main, umma_smem_desc(a_hi + ks * 32), umma_smem_desc(w_hi + ks * 32), idesc,
                                 (uint32_t)((k | ks) != 0));
                }
                release(smem_u32(&misc->w_empty[stage]));
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
              }
            }
            for (int s = 0; s < NSLOTS; ++s) umma_commit(smem_u32(&misc->acc_full[s]));
            continue;
          }
          for (int s = 0; s < NSLOTS; ++s) {
            mbar_wait(smem_u32(&misc->a_ready[s]), pa[s]);
            pa[s] ^= 1u;
            tc_fence_after();
            const uint32_t slot_base = act_base + s * LT::kSlotBytes;
            const uint32_t acc = tmem_base + (uint32_t)(s * 256);
            for (int n = 0; n < L.nc; ++n) {
              const uint32_t d_main = acc + n * 128;
              const uint32_t d_corr = d_main + 256;   // SPLIT only: cross terms hi*lo + lo*hi
              for (int k = 0; k < L.kc; ++k) {
                const uint32_t a_hi = slot_base + (uint32_t)L.a_src[k] * kTileBytes;
                const uint32_t a_lo = a_hi + kChunksPerSlot * kTileBytes;
                const int ks0 = L.ks0[k];
                mbar_wait(smem_u32(&misc->w_full[stage]), phase);
                tc_fence_after();
                const uint32_t w_hi = ring_base + stage * kTileBytes;
                for (int ks = ks0; ks < 4; ++ks)
                  umma_bf16_ss(d_main, umma_smem_desc(a_hi + ks * 32), umma_smem_desc(w_hi + ks * 32), idesc,
                               (uint32_t)((k | ks) != 0));
                if (SPLIT) {
                  for (int ks = ks0; ks < 4; ++ks)
                    umma_bf16_ss(d_corr, umma_smem_desc(a_lo + ks * 32), umma_smem_desc(w_hi + ks * 32), idesc,
                                 (uint32_t)((k | ks) != 0));
                }
                release(smem_u32(&misc->w_empty[stage]));
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
                if (SPLIT) {
                  mbar_wait(smem_u32(&misc->w_full[stage]), phase);
                  tc_fence_after();
                  const uint32_t w_lo = ring_base + stage * kTileBytes;
                  for (int ks = ks0; ks < 4; ++ks)
                    umma_bf16_ss(d_corr, umma_smem_desc(a_hi + ks * 32), umma_smem_desc(w_lo + ks * 32), idesc, 1u);
                  release(smem_u32(&misc->w_empty[stage]));
                  if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
              }
            }
            umma_commit(smem_u32(&misc->acc_full[s]));
          }
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== slot group: producer + epilogue ====================================
    const int s = (warp - 4) >> 2;
    const int wq = warp & 3;            // TMEM lane quadrant this warp may access
    const int row = wq * 32 + lane;     // row of the tile == TMEM lane
    const uint32_t slot_base = act_base + s * LT::kSlotBytes;
    const uint32_t lo_off = kChunksPerSlot * kTileBytes;
    const uint32_t e_hi = slot_base + kChunkE * kTileBytes, e_lo = e_hi + lo_off;
    const uint32_t acc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(s * 256);
    const uint32_t a_ready = smem_u32(&misc->a_ready[s]);
    const uint32_t acc_full = smem_u32(&misc->acc_full[s]);
    float* scratch = &misc->scratch[s][0][0];
    uint32_t pacc = 0;

    for (int64_t it = 0; it < n_iters; ++it) {
      const int64_t tile = (it * gridDim.x + blockIdx.x) * NSLOTS + s;   // may lie past n_tiles: rows invalid
      const int64_t grow = tile * kTileRows + row;
      const RowIn in = load_row(p.io, grow);
      write_enc_row<SPLIT, F16, kEncCols, kMaxPosLevels>(e_hi, e_lo, row, in.p, p.pos_levels, in.valid);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_ready);

      float sigma = 0.f;
      for (int l = 0; l < net.n_layers; ++l) {
        const int epi = net.layer[l].epi;
        mbar_wait(acc_full, pacc);
        pacc ^= 1u;
        __syncwarp();
        tc_fence_after();

        if (epi == EPI_RGB) {
          // ---- rgb_layer: t = relu(acc) (128 wide, bias folded), rgb = sigmoid(W1 t + b1) --------
          float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll 1
          for (int cb = 0; cb < kRgbHidden / 32; ++cb) {
            uint32_t r[32];
            tmem_ld32(acc + cb * 32, r);
            if (SPLIT) {
              uint32_t c[32];
              tmem_ld32(acc + 256 + cb * 32, c);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(c[j]));
            } else {
              tmem_ld_wait();
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int col = cb * 32 + g * 4;
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.head + kHeadRgbW + col));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.head + kHeadRgbW + 128 + col));
              const float4 w2 = __ldg(reinterpret_cast<const float4*>(p.head + kHeadRgbW + 256 + col));
              const float t0 = fmaxf(__uint_as_float(r[g * 4 + 0]), 0.f), t1 = fmaxf(__uint_as_float(r[g * 4 + 1]), 0.f);
              const float t2 = fmaxf(__uint_as_float(r[g * 4 + 2]), 0.f), t3 = fmaxf(__uint_as_float(r[g * 4 + 3]), 0.f);
              c0 = fmaf(t0, w0.x, c0); c0 = fmaf(t1, w0.y, c0); c0 = fmaf(t2, w0.z, c0); c0 = fmaf(t3, w0.w, c0);
              c1 = fmaf(t0, w1.x, c1); c1 = fmaf(t1, w1.y, c1); c1 = fmaf(t2, w1.z, c1); c1 = fmaf(t3, w1.w, c1);
              c2 = fmaf(t0, w2.x, c2); c2 = fmaf(t1, w2.y, c2); c2 = fmaf(t2, w2.z, c2); c2 = fmaf(t3, w2.w, c2);
            }
          }
          c0 = 1.f / (1.f + expf(-(c0 + __ldg(p.head + kHeadRgbB + 0))));
          c1 = 1.f / (1.f + expf(-(c1 + __ldg(p.head + kHeadRgbB + 1))));
          c2 = 1.f / (1.f + expf(-(c2 + __ldg(p.head + kHeadRgbB + 2))));
          if (p.io.out_mode == 1) {
            if (in.valid) reinterpret_cast<float4*>(p.io.out)[grow] = make_float4(c0, c1, c2, sigma);
          } else {
            // ---- alpha compositing over the rows of each ray (nerf_base.py:79-113) -------------
            const int P = p.io.P;                 // 32, 64 or 128: rays cover whole warps
            const int wpr = P >> 5;               // warps per ray
            const int wseg = wq % wpr;            // this warp's position inside its ray
            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(in.d[0], in.d[0]), __fmul_rn(in.d[1], in.d[1])),
                                              __fmul_rn(in.d[2], in.d[2])));
            const float depth = __fmul_rn(in.z, nrm);
            float next = __shfl_down_sync(0xffffffffu, depth, 1);
            if (lane == 0) scratch[wq * 8 + 0] = depth;
            named_bar_sync(1 + s, 128);
            if (lane == 31 && wq < 3) next = scratch[(wq + 1) * 8 + 0];
            const bool last = (in.s == P - 1);
            const float delta = last ? 1e10f : __fsub_rn(next, depth);
            const float m = in.valid ? expf(-fmaxf(sigma, 0.f) * delta) : 1.f;
            const float alpha = 1.f - m;
            const float inc = warp_scan_mul(m + 1e-10f, lane);
            float exc = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) exc = 1.f;
            if (lane == 31) scratch[wq * 8 + 1] = inc;
            named_bar_sync(1 + s, 128);
            float carry = 1.f;
            for (int w = wq - wseg; w < wq; ++w) carry *= scratch[w * 8 + 1];
            const float wgt = in.valid ? alpha * (carry * exc) : 0.f;
            float sr = warp_sum(wgt * c0), sg = warp_sum(wgt * c1), sb = warp_sum(wgt * c2);
            float sa = warp_sum(wgt), sd = warp_sum(wgt * depth);
            if (lane == 0) {
              scratch[wq * 8 + 2] = sr; scratch[wq * 8 + 3] = sg; scratch[wq * 8 + 4] = sb;
              scratch[wq * 8 + 5] = sa; scratch[wq * 8 + 6] = sd;
            }
            named_bar_sync(1 + s, 128);
            if (lane == 0 && wseg == 0 && in.valid) {
              for (int w = wq + 1; w < wq + wpr; ++w) {
                sr += scratch[w * 8 + 2]; sg += scratch[w * 8 + 3]; sb += scratch[w * 8 + 4];
                sa += scratch[w * 8 + 5]; sd += scratch[w * 8 + 6];
              }
              if (p.io.flags & NB2_WHITE_BKG) {
                const float bg = 1.f - sa;
                sr += bg; sg += bg; sb += bg;
              }
              p.io.rgb_out[in.ray * 3 + 0] = sr;
              p.io.rgb_out[in.ray * 3 + 1] = sg;
              p.io.rgb_out[in.ray * 3 + 2] = sb;
              if (p.io.depth_out) p.io.depth_out[in.ray] = (sd - p.io.near_t) / (p.io.far_t - p.io.near_t);
              if (p.io.acc_out) p.io.acc_out[in.ray] = sa;
            }
            named_bar_sync(1 + s, 128);  // scratch is reused by the next tile
          }
        } else if (epi == EPI_RELU) {
          epilogue_hidden<EPI_RELU, SPLIT, F16>(acc, slot_base, lo_off, row, p.head);
        } else if (epi == EPI_LINEAR) {
          epilogue_hidden<EPI_LINEAR, SPLIT, F16>(acc, slot_base, lo_off, row, p.head);
        } else if (epi == EPI_RELU_SIGMA) {
          sigma = epilogue_hidden<EPI_RELU_SIGMA, SPLIT, F16>(acc, slot_base, lo_off, row, p.head) + __ldg(p.head + kHeadSigmaB);
          if (p.has_dir) {
            // the encoded position is dead after the skip layer: re-use its tile for the direction
            float rot[3] = {0.f, 0.f, 0.f};
            if (in.valid) normalize_dir(in.d, rot);
            write_enc_row<SPLIT, F16, kDirCols, kMaxDirLevels>(e_hi, e_lo, row, rot, p.dir_levels, in.valid);
          }
        } else {  // EPI_SIGMA_OUT
          sigma = epilogue_hidden<EPI_SIGMA_OUT, SPLIT, F16>(acc, slot_base, lo_off, row, p.head) + __ldg(p.head + kHeadSigmaB);
          if (in.valid) p.io.out[grow] = sigma;
        }
        if (l + 1 < net.n_layers) {
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(a_ready);
        }
      }
      tc_fence_before();  // accumulator reads of the last layer precede the next tile's a_ready arrive
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (cl_size > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side ----------------------------------------------------------------------------------------
// Variant knobs (defaults chosen from measurements; overridable for experiments):
//   NB2_TC_CLUSTER = 1 | 2 | 4      CTAs sharing each weight tile via multicast
//   NB2_TC_LOCKSTEP = 0 | 1         two-slot modes: both tiles consume each weight tile
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <int NSLOTS, bool SPLIT, bool F16, bool LOCKSTEP>
static int launch_tc_impl(nb2_handle* h, TcParams& prm, cudaStream_t st) {
  using LT = TcLayout<NSLOTS, SPLIT>;
  auto kern = mlp_tc_kernel<NSLOTS, SPLIT, F16, LOCKSTEP>;
  static bool attr_set = false;
  if (!attr_set) {
    NB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LT::kTotal));
    attr_set = true;
  }
  const int cluster = prm.cluster;
  int64_t ctas = (prm.n_tiles + NSLOTS - 1) / NSLOTS;
  ctas = (ctas + cluster - 1) / cluster * cluster;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kRolesThreads + 128 * NSLOTS);
  cfg.dynamicSmemBytes = LT::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_ctas = h->sm_count / cluster * cluster;
  if (cluster > 1) {
    // co-resident clusters are limited by GPC boundaries: ask the driver
    static int cached[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!cached[cluster]) {
      cfg.gridDim = dim3(max_ctas);
      int n = 0;
      NB2_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      cached[cluster] = n > 0 ? n : 1;
    }
    max_ctas = cached[cluster] * cluster;
  }
  cfg.gridDim = dim3((unsigned)std::min<int64_t>(ctas, (int64_t)max_ctas));
  NB2_CUDA(cudaLaunchKernelEx(&cfg, kern, prm));
  h->launches++;
  return NB2_OK;
}

int launch_mlp_tc(nb2_handle* h, int net_id, int precision, const MlpIo& io, cudaStream_t st) {
  PackedNet& pn = h->net[net_id];
  if (!pn.packed) {
    set_error("mlp_forward: weights of network %d have not been packed (call nb2_pack_weights)", net_id);
    return NB2_ERR_STATE;
  }
  if (io.out_mode == 2 && !(io.in_mode != 0 && (io.P == 32 || io.P == 64 || io.P == 128))) {
    set_error("mlp_forward: fused compositing needs rays + depths with 32, 64 or 128 samples per ray (got %d)", io.P);
    return NB2_ERR_UNSUPPORTED;
  }
  if (io.n_rows == 0) return NB2_OK;
  TcParams prm;
  prm.net = pn.tc;
  prm.io = io;
  prm.wchunks = pn.d_wchunks;
  prm.bias = pn.d_bias;
  prm.head = pn.d_head;
  prm.pos_levels = pn.pos_levels;
  prm.dir_levels = pn.dir_levels;
  prm.has_dir = (net_id == NB2_NET_NERF);
  prm.n_tiles = (io.n_rows + kTileRows - 1) / kTileRows;
  const bool split = (precision == NB2_PREC_BF16X3 || precision == NB2_PREC_FP16X3);
  int cluster = env_int("NB2_TC_CLUSTER", split ? 2 : 2);
  if (cluster != 1 && cluster != 2 && cluster != 4) {
    set_error("NB2_TC_CLUSTER must be 1, 2 or 4 (got %d)", cluster);
    return NB2_ERR_INVALID;
  }
  if (prm.n_tiles < 2 * cluster) cluster = 1;
  prm.cluster = cluster;
  const bool lockstep = env_int("NB2_TC_LOCKSTEP", 1) != 0;
  if (precision == NB2_PREC_BF16) return lockstep ? launch_tc_impl<2, false, false, true>(h, prm, st) : launch_tc_impl<2, false, false, false>(h, prm, st);
  if (precision == NB2_PREC_FP16) return lockstep ? launch_tc_impl<2, false, true, true>(h, prm, st) : launch_tc_impl<2, false, true, false>(h, prm, st);
  if (precision == NB2_PREC_BF16X3) return launch_tc_impl<1, true, false, false>(h, prm, st);
  if (precision == NB2_PREC_FP16X3) return launch_tc_impl<1, true, true, false>(h, prm, st);
  set_error("mlp_forward: unknown tensor-core precision %d", precision);
  return NB2_ERR_INVALID;
}

// ======================================================================================================
// Self-test: D (128x128 fp32) = A (128x64 bf16) * B^T (128x64 bf16) through exactly the operand layout,
// descriptors, bulk copy, commit and TMEM read-out used above.  A is written with generic stores (as
// the epilogue does), B arrives through cp.async.bulk from a pre-swizzled global image (as weights do).
// ======================================================================================================
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bswz, float* __restrict__ D) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* al = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t a_tile = base, b_tile = base + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(al + 2 * kTileBytes);  // [0] = B landed, [1] = MMA done
  uint32_t* tptr = reinterpret_cast<uint32_t*>(al + 2 * kTileBytes + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tptr), 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;

  for (int g = 0; g < 8; ++g) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(A[row * 64 + g * 8 + i]);
    store_a8<false, false>(a_tile, a_tile, row, g * 8, v);
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bars[0]), kTileBytes);
    bulk_g2s(b_tile, Bswz, kTileBytes, smem_u32(&bars[0]));
    mbar_wait(smem_u32(&bars[0]), 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_16(128, 128, false);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      umma_bf16_ss(tmem, umma_smem_desc(a_tile + ks * 32), umma_smem_desc(b_tile + ks * 32), idesc, (uint32_t)(ks != 0));
    umma_commit(smem_u32(&bars[1]));
  }
  mbar_wait(smem_u32(&bars[1]), 0);
  __syncwarp();
  tc_fence_after();
  for (int cb = 0; cb < 4; ++cb) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[row * 128 + cb * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// Swizzle a row-major 128x64 bf16 matrix into the tile image (used by the self-test and by tests).
__global__ void swizzle_tile_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kTileRows * kTileCols) return;
  int r = i / kTileCols, c = i % kTileCols;
  dst[swz128_offset(r, c) / 2] = src[i];
}

int selftest_umma(nb2_handle* h, const void* A, const void* B, void* Bswz_scratch, float* D, cudaStream_t st) {
  swizzle_tile_kernel<<<(kTileRows * kTileCols + 255) / 256, 256, 0, st>>>((const __nv_bfloat16*)B,
                                                                          (__nv_bfloat16*)Bswz_scratch);
  NB2_LAUNCH_CHECK(h);
  const int smem = 2 * kTileBytes + 1024 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    NB2_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  umma_selftest_kernel<<<1, 128, smem, st>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)Bswz_scratch, D);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

}  // namespace nb2
