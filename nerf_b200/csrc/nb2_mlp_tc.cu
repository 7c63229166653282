// nb2_mlp_tc.cu — the per-sample MLP (proposal 4x256 / NeRF 8x256 + heads) as ONE persistent, warp-specialised tcgen05
// kernel with both operands in shared memory, plus the precision dispatch (launch_mlp_tc) for every tensor-core kernel.
//
//   mlp_tc2_kernel   CTA pair (cta_group::2, M256 x N256 x K16): the single-pass precisions (NB2_PREC_BF16 /
//                    NB2_PREC_FP16), two resident tiles in ping-pong.  The split precisions run nb2_mlp_tc4.cu.
//
//   warp 0      weight streamer: cp.async.bulk (TMA unit) of pre-swizzled 128x64 16-bit weight tiles from L2 into a
//               4-stage shared-memory ring, mbarrier full/empty handshake (pair kernel: each CTA streams its half of
//               every tile, the peer relays "landed" onto the leader's stage-full barrier with a relaxed remote arrive)
//   warp 1      MMA issuer (pair kernel: leader CTA only, the whole warp runs the loop and elects one lane per tcgen05
//               instruction); fp32 accumulators in TMEM, completion through tcgen05.commit (multicast to both CTAs)
//   warp 2      TMEM allocator
//   warps 4..   one 128-thread "slot group" per resident 128-row tile (split modes: both groups share the one tile by
//               column halves): it produces the tile's first A operand (sample point -> sinusoidal encoding -> swizzled
//               16-bit rows), and after every layer drains the accumulator from TMEM (tcgen05.ld), applies ReLU,
//               re-quantises to 16 bit (hi [+ lo]) and writes the next layer's A operand in place.  The last epilogue
//               evaluates the 128->3 / 256->1 heads on the fp32 values and either stores rgb-sigma or alpha-composites the
//               ray.  The next tile's encoding and the direction encoding are computed while the group would otherwise
//               wait for an accumulator.
//
// single pass (bf16 | fp16): two tiles; ping-pong (default in the pair kernel): the issuer alternates tiles layer by layer, so
//                  one tile's epilogue runs under the other tile's MMAs; lockstep (NB2_TC_LOCKSTEP=1): both tiles consume every
//                  weight stage back to back.
// split (fp16x3 | bf16x3)   : every operand is split x = hi + lo (both 16-bit); each product is evaluated
//                  as hi*hi + lo*hi + hi*lo with fp32 accumulation (fp16: 22-bit operands, ~2^-22 relative
//                  per product, i.e. fp32-faithful; bf16: ~2^-16), one tile (the hi/lo activation pair
//                  fills the shared memory of two tiles).
//
// Shared memory (bytes):  activations NSLOTS * (SPLIT ? 2 : 1) * 5 * 16 KB = 160 KB,
//                         weight ring 4 * 16 KB = 64 KB, barriers + scratch < 1 KB.
// TMEM: 512 columns; tile s owns columns [256 s, 256 s + 256) (128 lanes x 256 fp32).  Split mode has one
//       tile and uses columns [256, 512) as a second accumulator for the hi*lo + lo*hi cross terms, so the
//       main accumulator sees a third of the (truncating) tensor-core additions; the two are summed in fp32
//       by the epilogue.
// Biases ride on the tensor core: encoding column 63 is the constant 1 and the matching weight column is
// the bias (layers without an encoding K-chunk get a one-k-step "bias chunk").
#include <stdlib.h>

#include "nb2_tc_device.cuh"

namespace nb2 {
using namespace ptx;

// ---- slot group: per-tile producer (sample -> encoding) and per-layer epilogue ------------------------------------
// Shared by the single-CTA kernel and the CTA-pair kernel (PAIR: the operand-ready barrier lives in the pair's leader).
template <int NSLOTS, bool SPLIT, bool F16, bool PAIR>
__device__ __forceinline__ void slot_group_run(const TcParams& p, TcMisc* misc, uint32_t act_base, uint32_t tmem_base,
                                               int64_t n_iters, int warp, int lane, uint32_t cl_rank) {
  using LT = TcLayout<NSLOTS, SPLIT>;
  const TcNet& net = p.net;
  // one arrival per warp: every lane has fenced its own shared-memory writes (fence.proxy.async) / TMEM reads, __syncwarp
  // orders them before the elected lane's arrive.  The peer CTA's arrive is relaxed: the consumer of the data is the
  // tensor core of this SM pair, not the waiting thread, and the release / acquire pair at cluster scope costs several
  // hundred cycles per hand-over (see mbar_arrive_remote_relaxed)
  auto arrive_a = [&](uint32_t bar) {
    __syncwarp();
    if (lane == 0) {
      if (PAIR && cl_rank != 0) mbar_arrive_remote_relaxed(bar, 0); else mbar_arrive(bar);
    }
  };
  // =========================== slot group: producer + epilogue ====================================
  constexpr int EW = GroupsPerSlot<NSLOTS>::value;   // warpgroups sharing one tile (column halves)
  const int wg = (warp - 4) >> 2;
  const int s = (EW == 2) ? 0 : wg;
  const int half = (EW == 2) ? wg : 0;
  const int cb0 = half * (8 / EW), cb1 = cb0 + 8 / EW;   // this thread's 32-column blocks of a 256-wide layer
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
  long long t_pe = 0, t_wacc = 0, t_epi = 0, t_last = 0, t0e = NB2_CLK();

  // The encoding of tile it+1 is computed in the windows where this group would otherwise spin on the accumulator
  // barrier (while layers 1 and 2 of tile it run on the tensor core), and stored (tile "begun") inside tile it's last
  // epilogue, as soon as its accumulator has been read.  Measured before this change (profiles/r01_roles_pair2wg.txt):
  // 7.8 k (single pass) / 4.4 k (split) exposed cycles per iteration for the encoding, plus the direction encoding.
  auto tile_of = [&](int64_t it) { return (it * gridDim.x + blockIdx.x) * NSLOTS + s; };   // may lie past n_tiles
  EncRegs<SPLIT, F16, 0, 4> enc_a;   // encoding column groups 0-3: warpgroup 0 of a shared tile, or phase 0 of a whole-tile group
  EncRegs<SPLIT, F16, 4, 4> enc_b;   // groups 4-7:                 warpgroup 1,                  or phase 1
  auto enc_make = [&](const RowIn& r, int phase) {
    if (EW == 2 ? (half == 0) : (phase == 0)) enc_compute(enc_a, r.p, p.pos_levels, r.valid, r.enc, r.ipe ? r.cov : nullptr);
    else enc_compute(enc_b, r.p, p.pos_levels, r.valid, r.enc, r.ipe ? r.cov : nullptr);
  };
  auto begin_tile = [&]() {
    if (EW == 1 || half == 0) enc_store(enc_a, e_hi, e_lo, row);
    if (EW == 1 || half == 1) enc_store(enc_b, e_hi, e_lo, row);
    fence_proxy_async_smem();
    tc_fence_before();
    arrive_a(a_ready);
  };
  RowIn in = load_row(p.io, tile_of(0) * kTileRows + row);
  enc_make(in, 0);
  if (EW == 1) enc_make(in, 1);
  begin_tile();

  for (int64_t it = 0; it < n_iters; ++it) {
    const int64_t grow = tile_of(it) * kTileRows + row;
    const bool has_next = (it + 1 < n_iters);
    RowIn in_next = in;
    if (has_next) in_next = load_row(p.io, tile_of(it + 1) * kTileRows + row);   // consumed after layer 0's epilogue

    float sigma = 0.f;
    for (int l = 0; l < net.n_layers; ++l) {
      const int epi = net.layer[l].epi;
      const long long cw = NB2_CLK();
      mbar_wait(acc_full, pacc);
      pacc ^= 1u;
      __syncwarp();
      tc_fence_after();
      const long long ce = NB2_CLK();
      t_wacc += ce - cw;

      if (epi == EPI_RGB) {
        // ---- rgb_layer: t = relu(acc) (128 wide, bias folded), rgb = sigmoid(W1 t + b1) --------
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        bool continue_flag = true;   // group 1 of a two-group tile stops after handing over its partial sums
#pragma unroll 1
        for (int cb = half * (4 / EW); cb < half * (4 / EW) + 4 / EW; ++cb) {
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
        if (EW == 2) {
          // combine the two column halves: group 1 parks its partial sums (and its half of the density dot
          // product) in the slot's first activation tile, which is dead once this layer's MMAs completed
          const uint32_t xaddr = slot_base + (uint32_t)row * 16u;
          if (half == 1) st_shared_v4(xaddr, __float_as_uint(sigma), __float_as_uint(c0), __float_as_uint(c1), __float_as_uint(c2));
          named_bar_sync(3, 256);
          if (half == 1) continue_flag = false;
          else {
            uint32_t x0, x1, x2, x3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(xaddr));
            sigma += __uint_as_float(x0); c0 += __uint_as_float(x1); c1 += __uint_as_float(x2); c2 += __uint_as_float(x3);
          }
        }
        if (has_next) begin_tile();   // accumulator and activation tiles of this tile are no longer needed
        if (continue_flag) {
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
            // fused gather: the same row goes to every peer GPU's image buffer over NVLink (nb2_render_params.peer_rgb)
            for (int q = 0; q < p.io.n_peers; ++q) {
              float* o = p.io.peer_rgb[q] + (p.io.peer_row0 + in.ray) * 3;
              o[0] = sr; o[1] = sg; o[2] = sb;
            }
            if (p.io.depth_out) p.io.depth_out[in.ray] = (sd - p.io.near_t) / (p.io.far_t - p.io.near_t);
            if (p.io.acc_out) p.io.acc_out[in.ray] = sa;
          }
          named_bar_sync(1 + s, 128);  // scratch is reused by the next tile
        }
        }  // continue_flag
      } else if (epi == EPI_RELU) {
        epilogue_hidden<EPI_RELU, SPLIT, F16>(acc, slot_base, lo_off, row, p.head, cb0, cb1);
      } else if (epi == EPI_LINEAR) {
        epilogue_hidden<EPI_LINEAR, SPLIT, F16>(acc, slot_base, lo_off, row, p.head, cb0, cb1);
      } else if (epi == EPI_RELU_SIGMA) {
        // (with two groups per tile each keeps the dot product over its own columns; the bias is added once)
        sigma = epilogue_hidden<EPI_RELU_SIGMA, SPLIT, F16>(acc, slot_base, lo_off, row, p.head, cb0, cb1) +
                (half == 0 ? __ldg(p.head + kHeadSigmaB) : 0.f);
      } else {  // EPI_SIGMA_OUT
        sigma = epilogue_hidden<EPI_SIGMA_OUT, SPLIT, F16>(acc, slot_base, lo_off, row, p.head, cb0, cb1) +
                (half == 0 ? __ldg(p.head + kHeadSigmaB) : 0.f);
        if (EW == 2) {
          const uint32_t xaddr = slot_base + (uint32_t)row * 4u;
          if (half == 1) asm volatile("st.shared.b32 [%0], %1;" ::"r"(xaddr), "r"(__float_as_uint(sigma)));
          named_bar_sync(3, 256);
          if (half == 0) {
            uint32_t x0;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x0) : "r"(xaddr));
            sigma += __uint_as_float(x0);
          }
        }
        if (has_next) begin_tile();
        if (in.valid && half == 0) p.io.out[grow] = sigma;
      }
      if (l + 1 < net.n_layers) {
        fence_proxy_async_smem();
        tc_fence_before();
        arrive_a(a_ready);
        t_epi += NB2_CLK() - ce;
        // ---- work hidden behind the next layer's MMAs -------------------------------------------------------------
        const long long cpe = NB2_CLK();
        if (has_next && l == 0) enc_make(in_next, 0);
        if (has_next && l == 1 && EW == 1) enc_make(in_next, 1);
        if (l == p.dir_layer && half == 0) {
          // the encoded position is dead after the skip layer: its tile now takes the encoded direction (columns 0-31;
          // the bias k-step of the running layer reads columns 48-63 of the same rows, other 16-byte units)
          float rot[3] = {0.f, 0.f, 0.f};
          if (in.valid) normalize_dir(in.d, rot);
          write_enc_row<SPLIT, F16, kDirCols, kMaxDirLevels>(e_hi, e_lo, row, rot, p.dir_levels, in.valid);
          fence_proxy_async_smem();
        }
        t_pe += NB2_CLK() - cpe;
      } else {
        t_last += NB2_CLK() - ce;
      }
    }
    in = in_next;
  }
  if (NB2_PROF_ON && threadIdx.x == kRolesThreads) {
    long long* o = p.prof + blockIdx.x * 16;
    o[6] = t_pe; o[7] = t_wacc; o[8] = t_epi; o[9] = t_last; o[10] = NB2_CLK() - t0e; o[11] = n_iters; o[12] = net.n_layers;
  }
}

// ======================================================================================================
// CTA-pair kernel (cta_group::2).  Two CTAs of a cluster run ONE M=256 x N=256 x K=16 tcgen05.mma per k-step:
// each CTA contributes its own 128 activation rows (A, from its shared memory) and HALF of the weight tile (B rows
// [128 r, 128 r + 128) of the N=256 output columns), and receives its 128 rows of the fp32 accumulator in its own TMEM.
// Versus the single-CTA kernel this halves (a) the MMA instructions issued per FLOP (the issuing thread is the scarce
// resource: ~100 cycles of dependent issue latency per instruction), (b) the weight bytes each SM pulls from L2 and
// writes to shared memory, so the 4 x 16 KB ring now covers 4 full K-chunks.
//   leader CTA (rank 0): warp 1 issues all MMAs; waits for both CTAs' operands (a_ready counts 2 x 128 arrivals per slot)
//                        and for both halves of each weight tile (one w_full barrier: own bulk copy + the peer's relaxed relay arrive);
//   both CTAs          : warp 0 streams this CTA's half tiles; slot groups as in the single-CTA kernel; completion
//                        (tcgen05.commit .multicast::cluster) is signalled into both CTAs.
// Two-slot (single pass) mode runs the slots in lockstep: every weight half-tile feeds 2 x 4 MMAs = 1024 tensor cycles.
// ======================================================================================================
template <int NSLOTS, bool SPLIT, bool F16, bool LOCKSTEP>
__global__ void __launch_bounds__(kTcThreads, 1) mlp_tc2_kernel(const __grid_constant__ TcParams p) {
  constexpr int kPasses = LOCKSTEP ? 1 : NSLOTS;   // how many times a layer's weight stream is consumed per iteration
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
  const int64_t n_iters = (p.n_tiles + tiles_per_iter - 1) / tiles_per_iter;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      // leader: a stage is full when its own bulk copy has landed AND the peer has relayed that its half has landed
      mbar_init(smem_u32(&misc->w_full[i]), rank == 0 ? 2 : 1);
      mbar_init(smem_u32(&misc->w_empty[i]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&misc->a_ready[s]), 8 * GroupsPerSlot<NSLOTS>::value);   // every slot-group warp of both CTAs
      mbar_init(smem_u32(&misc->acc_full[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc2(smem_u32(&misc->tmem_base), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = misc->tmem_base;

  if (warp == 0) {
    // =========================== weight streamer: this CTA's half of every tile =====================
    reg_dealloc<kRoleRegs>();
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < n_iters; ++it) {
        for (int l = 0; l < net.n_layers; ++l) {
          const TcLayer& L = net.layer[l];
          // nc == 2: this CTA owns N-chunk `rank` (16 KB per K-chunk); nc == 1: rows [64 rank, 64 rank + 64) (8 KB)
          const int c_first = L.chunk0 + (L.nc == 2 ? (int)rank * L.kc : 0);
          const uint32_t bytes = L.nc == 2 ? kTileBytes : kTileBytes / 2;
          const size_t sub = L.nc == 2 ? 0 : (size_t)rank * (kTileBytes / 4);   // in 16-bit elements
          for (int pass = 0; pass < kPasses; ++pass)
          for (int k = 0; k < L.kc; ++k) {
#pragma unroll
            for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
              mbar_wait(smem_u32(&misc->w_empty[stage]), phase ^ 1u);
              const uint32_t full = smem_u32(&misc->w_full[stage]);
              mbar_arrive_expect_tx(full, bytes);
              bulk_g2s(ring_base + stage * kTileBytes,
                       p.wchunks + ((size_t)(c_first + k) * 4 + (F16 ? 2 : 0) + part) * (kTileBytes / 2) + sub, bytes, full);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    reg_dealloc<kRoleRegs>();
    if (lane == 0 && rank != 0) {
      // =========================== peer: relay "my half has landed" to the leader =====================
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < n_iters; ++it)
        for (int l = 0; l < net.n_layers; ++l) {
          const int n_entries = net.layer[l].kc * (SPLIT ? 2 : 1) * kPasses;
          for (int e = 0; e < n_entries; ++e) {
            mbar_wait(smem_u32(&misc->w_full[stage]), phase);
            mbar_arrive_remote_relaxed(smem_u32(&misc->w_full[stage]), 0);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
    } else if (rank == 0) {
      // =========================== leader: MMA issuer for the pair ====================================
      // (the whole warp runs the loop and elects one lane per tcgen05 instruction, see nb2_tc_ptx.cuh)
      uint32_t stage = 0, phase = 0;
      uint32_t pa[2] = {0u, 0u};
      long long t_wa = 0, t_ww = 0, t0m = NB2_CLK();
      const uint32_t ring_lo = umma_desc_lo(ring_base);
      for (int64_t it = 0; it < n_iters; ++it) {
        for (int l = 0; l < net.n_layers; ++l) {
          const TcLayer& L = net.layer[l];
          const uint32_t idesc = umma_idesc_16(256, L.nc * 128, F16);
          for (int pass = 0; pass < kPasses; ++pass) {
            // LOCKSTEP: one pass, both slots share every weight half-tile.  Otherwise one pass per slot (ping-pong):
            // the other slot's epilogue runs while this slot's MMAs execute.
            const int s_lo = LOCKSTEP ? 0 : pass, s_hi = LOCKSTEP ? NSLOTS : pass + 1;
            { const long long c0 = NB2_CLK();
            for (int s = s_lo; s < s_hi; ++s) {
              mbar_wait(smem_u32(&misc->a_ready[s]), pa[s]);
              pa[s] ^= 1u;
            }
            t_wa += NB2_CLK() - c0; }
            tc_fence_after();
            for (int k = 0; k < L.kc; ++k) {
              const int ks0 = L.ks0[k];
              const uint32_t a_lo0 = umma_desc_lo(act_base + (uint32_t)L.a_src[k] * kTileBytes);   // slot 0, hi part
              { const long long c0 = NB2_CLK();
              mbar_wait(smem_u32(&misc->w_full[stage]), phase);
              t_ww += NB2_CLK() - c0; }
              tc_fence_after();
              const uint32_t w_lo0 = ring_lo + stage * (kTileBytes >> 4);
              for (int s = s_lo; s < s_hi; ++s) {
                const uint32_t a_s = a_lo0 + s * (LT::kSlotBytes >> 4);
                const uint32_t d_main = tmem_base + (uint32_t)(s * 256);
                for (int ks = ks0; ks < 4; ++ks)
                  umma2_f16_ss_elect(d_main, umma_desc_from_lo(a_s + 2 * ks), umma_desc_from_lo(w_lo0 + 2 * ks), idesc,
                                (uint32_t)((k | ks) != 0));
                if (SPLIT) {
                  const uint32_t a_l = a_s + (kChunksPerSlot * kTileBytes >> 4);
                  for (int ks = ks0; ks < 4; ++ks)
                    umma2_f16_ss_elect(d_main + 256, umma_desc_from_lo(a_l + 2 * ks), umma_desc_from_lo(w_lo0 + 2 * ks), idesc,
                                  (uint32_t)((k | ks) != 0));
                }
              }
              umma2_commit_mcast_elect(smem_u32(&misc->w_empty[stage]), 3);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
              if (SPLIT) {
                { const long long c0 = NB2_CLK();
                mbar_wait(smem_u32(&misc->w_full[stage]), phase);
                t_ww += NB2_CLK() - c0; }
                tc_fence_after();
                const uint32_t wl = ring_lo + stage * (kTileBytes >> 4);
                for (int ks = ks0; ks < 4; ++ks)
                  umma2_f16_ss_elect(tmem_base + 256, umma_desc_from_lo(a_lo0 + 2 * ks), umma_desc_from_lo(wl + 2 * ks), idesc, 1u);
                umma2_commit_mcast_elect(smem_u32(&misc->w_empty[stage]), 3);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
              }
            }
            for (int s = s_lo; s < s_hi; ++s) umma2_commit_mcast_elect(smem_u32(&misc->acc_full[s]), 3);
          }
        }
      }
      if (NB2_PROF_ON && lane == 0) { p.prof[blockIdx.x * 16 + 3] = t_wa; p.prof[blockIdx.x * 16 + 4] = t_ww; p.prof[blockIdx.x * 16 + 5] = NB2_CLK() - t0m; }
    }
  } else if (warp >= 4) {
    reg_alloc<kGroupRegs>();
    slot_group_run<NSLOTS, SPLIT, F16, true>(p, misc, act_base, tmem_base, n_iters, warp, lane, rank);
  } else {
    reg_dealloc<kRoleRegs>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---- host side ----------------------------------------------------------------------------------------
template <bool F16>
static int launch_tc2_impl(nb2_handle* h, TcParams& prm, cudaStream_t st) {
  using LT = TcLayout<2, false>;
  auto kern = mlp_tc2_kernel<2, false, F16, false>;
  int rc = kernel_set_smem(h, (const void*)kern, LT::kTotal);
  if (rc != NB2_OK) return rc;
  int64_t ctas = (prm.n_tiles + 1) / 2;   // two resident tiles per CTA
  ctas = (ctas + 1) / 2 * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = LT::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(h->sm_count / 2 * 2);
  int max_clusters = 0;
  rc = kernel_max_clusters(h, (const void*)kern, &cfg, &max_clusters);
  if (rc != NB2_OK) return rc;
  cfg.gridDim = dim3((unsigned)std::min<int64_t>(ctas, (int64_t)max_clusters * 2));
  prm.cluster = 2;
  NB2_CUDA(cudaLaunchKernelEx(&cfg, kern, prm));
  h->launches++;
  return NB2_OK;
}

// Precision dispatch of the tensor-core kernels: split precisions -> TMEM-operand kernel (nb2_mlp_tc4.cu), single pass ->
// the ping-pong pair kernel above.  (The single-CTA, lockstep, N-half and two-accumulator variants of round 1 measured
// slower and were retired; DESIGN.md section 9 keeps their numbers.)
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
  prm.has_dir = (pn.kind == NB2_NET_NERF);
  prm.debug = h->tc_debug;
  prm.dir_layer = -1;   // the direction encoding is written after the epilogue of the layer before the density layer
  for (int l = 0; l < pn.tc.n_layers; ++l)
    if (prm.has_dir && pn.tc.layer[l].epi == EPI_RELU_SIGMA) prm.dir_layer = l - 1;
  prm.n_tiles = (io.n_rows + kTileRows - 1) / kTileRows;
  prm.prof = h->tc_prof;
  prm.cluster = 2;
  if (precision == NB2_PREC_BF16X3 || precision == NB2_PREC_FP16X3) return launch_mlp_tc4(h, prm, precision, st);
  if (precision == NB2_PREC_BF16) return launch_tc2_impl<false>(h, prm, st);
  if (precision == NB2_PREC_FP16) return launch_tc2_impl<true>(h, prm, st);
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

// ======================================================================================================
// 2-CTA self-test + MMA micro-benchmark.  mode 0: cta_group::1 M128 N128; mode 1: cta_group::1 M128 N256;
// mode 2: cta_group::2 M256 N256 (cluster of 2).  A: (256 x 64) bf16 row-major, rows 128r.. belong to CTA r.
// B: (256 x 64) bf16 row-major (n-major).  D_out: (256 x 256) fp32.  cycles_out[cta] = cycles per MMA instruction
// over `iters` back-to-back K=64 sweeps (4 MMAs each).
// ======================================================================================================
// (a kernel that contains cta_group::2 instructions can only be launched as a cluster of 2: one instantiation per mode)
// `flags` (stressors, cta_group::1 modes): 1 = alternate between two accumulators every 4 MMAs; 2 = walk A over 8 and B over
// 4 distinct tiles; 4 = a second warp streams 16 KB bulk copies into a 4-stage ring concurrently; 8 = two warps issue
// 16-byte shared stores concurrently; 16 = commit + wait after every 4 MMAs (per-chunk barrier round trip).
template <int mode>
__global__ void __launch_bounds__(256, 1)
umma_bench_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
                  long long* __restrict__ cycles_out, int iters, int flags, const __nv_bfloat16* __restrict__ gsrc) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* al = smem_dyn + (base - smem_u32(smem_dyn));
  // layout: A tiles 8 x 16 KB | B tiles 4 x 16 KB (mode 1/2 use 32 KB of it) | 1 KB misc
  const uint32_t a_tile = base, b_tile = base + 8 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(al + 12 * kTileBytes);   // [0] done, [1..4] ring full, [5] chunk
  uint32_t* tptr = reinterpret_cast<uint32_t*>(al + 12 * kTileBytes + 64);
  volatile int* stop = reinterpret_cast<volatile int*>(al + 12 * kTileBytes + 72);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x & 127;
  constexpr bool pair = (mode == 2);
  const uint32_t rank = pair ? cluster_ctarank() : 0u;
  const int b_rows = (mode == 0) ? 128 : (pair ? 128 : 256);                // B rows held by THIS CTA
  const int N = (mode == 0) ? 128 : 256;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&bars[i]), 1);
    *stop = 0;
    mbar_fence_init();
  }
  if (warp == 0) {
    if (pair) { tmem_alloc2(smem_u32(tptr), 512); tmem_relinquish2(); }
    else      { tmem_alloc(smem_u32(tptr), 512); tmem_relinquish(); }
  }
  const int a_row0 = pair ? 128 * (int)rank : 0;
  if (threadIdx.x < 128) {
    for (int t = 0; t < 8; ++t)
      for (int g = 0; g < 8; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(A[(a_row0 + row) * 64 + g * 8 + i]);
        store_a8<false, false>(a_tile + t * kTileBytes, 0, row, g * 8, v);
      }
    const int b_row0 = pair ? 128 * (int)rank : 0;
    for (int t = 0; t < (mode == 0 ? 4 : 1); ++t)
      for (int r = row; r < b_rows; r += 128)
        for (int g = 0; g < 8; ++g) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(B[(b_row0 + r) * 64 + g * 8 + i]);
          store_a8<false, false>(b_tile + t * kTileBytes, 0, r, g * 8, v);
        }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tptr;

  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_16(pair ? 256 : 128, N, false);
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (((flags & 1) && (it & 1)) ? 256u : 0u);
      const uint32_t at = a_tile + ((flags & 2) ? (uint32_t)(it & 7) * kTileBytes : 0u);
      const uint32_t bt = b_tile + (((flags & 2) && mode == 0) ? (uint32_t)(it & 3) * kTileBytes : 0u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t acc = (uint32_t)((it > ((flags & 1) ? 1 : 0)) || ks != 0);
        if (pair) umma2_bf16_ss(d, umma_smem_desc(at + ks * 32), umma_smem_desc(bt + ks * 32), idesc, acc);
        else      umma_bf16_ss(d, umma_smem_desc(at + ks * 32), umma_smem_desc(bt + ks * 32), idesc, acc);
      }
      if (flags & 16) {
        umma_commit(smem_u32(&bars[5]));
        mbar_wait(smem_u32(&bars[5]), ph);
        ph ^= 1u;
        tc_fence_after();
      }
    }
    if (pair) umma2_commit_mcast(smem_u32(&bars[0]), 3); else umma_commit(smem_u32(&bars[0]));
    mbar_wait(smem_u32(&bars[0]), 0);
    cycles_out[blockIdx.x] = (clock64() - t0) / (4ll * iters);
    *stop = 1;
  } else if (warp == 4 && lane == 0 && (flags & 4) && !pair) {
    // concurrent weight-style traffic: 16 KB bulk copies into a 4-stage ring (placed after the B tiles is not possible in
    // this layout, so the ring aliases A tiles 4..7, unused unless flag 2)
    uint32_t st = 0, ph = 0, n = 0;
    while (!*stop && n < 100000) {
      const uint32_t full = smem_u32(&bars[1 + st]);
      mbar_arrive_expect_tx(full, kTileBytes);
      bulk_g2s(a_tile + (4 + st) * kTileBytes, gsrc + (size_t)(n & 63) * (kTileBytes / 2), kTileBytes, full);
      if (st == 3) {   // wait for the whole batch of 4 before re-arming (keeps <= 4 in flight)
        for (int i = 0; i < 4; ++i) mbar_wait(smem_u32(&bars[1 + i]), ph);
        ph ^= 1u;
      }
      st = (st + 1) & 3;
      ++n;
    }
    if (st != 0) for (uint32_t i = 0; i < st; ++i) mbar_wait(smem_u32(&bars[1 + i]), ph);
  } else if ((warp == 5 || warp == 6) && (flags & 8) && !pair) {
    // concurrent epilogue-style traffic: 16-byte shared stores, conflict-free pattern, into A tiles 4..7 region end
    uint32_t n = 0;
    const uint32_t dst = a_tile + 7 * kTileBytes + (uint32_t)(warp - 5) * 8192u;
    while (!*stop && n < 4000000) {
      st_shared_v4(dst + ((n & 15) * 512u) + lane * 16u, n, n, n, n);
      ++n;
    }
  }
  mbar_wait(smem_u32(&bars[0]), 0);
  __syncwarp();
  tc_fence_after();
  if (blockIdx.x < (pair ? 2 : 1) && threadIdx.x < 128) {
    const float scale = 1.f / (float)(((flags & 1) ? (iters + 1) / 2 : iters));
    for (int cb = 0; cb < N / 32; ++cb) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) D[(size_t)(a_row0 + row) * 256 + cb * 32 + j] = __uint_as_float(r[j]) * scale;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (pair) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

int umma_bench(nb2_handle* h, const void* A, const void* B, float* D, long long* cycles, int mode, int iters, int flags,
               const void* gsrc, cudaStream_t st) {
  const int smem = 12 * kTileBytes + 1024 + 1024;
  void (*kern)(const __nv_bfloat16*, const __nv_bfloat16*, float*, long long*, int, int, const __nv_bfloat16*) =
      mode == 0 ? umma_bench_kernel<0> : (mode == 1 ? umma_bench_kernel<1> : umma_bench_kernel<2>);
  {
    const int rc = kernel_set_smem(h, (const void*)kern, smem);
    if (rc != NB2_OK) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(256);
  cfg.gridDim = dim3(h->sm_count / 2 * 2);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (mode == 2) ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NB2_CUDA(cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, cycles, iters, flags,
                              (const __nv_bfloat16*)gsrc));
  h->launches++;
  return NB2_OK;
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
  {
    const int rc = kernel_set_smem(h, (const void*)umma_selftest_kernel, smem);
    if (rc != NB2_OK) return rc;
  }
  umma_selftest_kernel<<<1, 128, smem, st>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)Bswz_scratch, D);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

}  // namespace nb2
