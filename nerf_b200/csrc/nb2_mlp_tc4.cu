// nb2_mlp_tc4.cu — split-precision (fp16x3 / bf16x3) MLP kernel with the hidden activations in TENSOR MEMORY.
//
// What the layer-serial pair kernel (mlp_tc2_kernel, split mode) spends per 128-row tile and iteration on B200
// (profiles/r01_roles_pair2wg.txt): 53.5 k cycles issuing MMAs, 18 k waiting for weight stages (the hi + lo activation
// tiles fill 160 KB of shared memory, which leaves a 64 KB ring against a ~3 k-cycle refill latency: 21 B/cycle of
// supply for 21 B/cycle of demand), 26 k in hidden-layer epilogues (two accumulators to drain, 128 KB of st.shared per
// layer and a fence.proxy.async before every hand-over).  This kernel removes the causes instead of tuning them:
//   * the next layer's A operand (hi and lo halves, K = 256) is written by the epilogue with tcgen05.st into TMEM columns
//     [256, 512) and consumed by tcgen05.mma in its A-from-TMEM form; only the 64-column encoding tile stays in shared
//     memory.  No st.shared / fence.proxy.async on the per-layer path, no A-operand reads competing for shared memory;
//   * shared memory is now a 176 KB weight ring (11 x 16 KB): the refill latency is covered with room to spare;
//   * ONE fp32 accumulator (columns [0, 256)).  The accuracy argument of the two-accumulator scheme (DESIGN.md section 5:
//     the tensor core's fp32 accumulation truncates, so the large hi x hi partial sums must see as few additions as
//     possible) is kept by ORDER: every layer first accumulates all cross terms (lo x Wh, hi x Wl: small), then all
//     hi x Wh terms.  The large-magnitude additions are the same 16 (+ bias) as before, and the final fp32 add of the
//     two accumulators disappears.  Price: Wh is streamed twice per layer (48 KB instead of 32 KB per K chunk; the L2
//     stream sustains 60 B/cycle/SM, 32 are needed).
// Everything else (CTA pair, M256 x N256 x K16 instructions, roles, weight image, last-layer epilogues, encoders) is
// mlp_tc2_kernel's.  Single-pass modes keep mlp_tc2_kernel: two resident tiles need all 512 TMEM columns for accumulators.
#include <stdlib.h>

#include "nb2_tc_device.cuh"

namespace nb2 {
using namespace ptx;

constexpr int kStages4 = 11;
constexpr int kAHiCol = 256;   // TMEM columns of the A operand: hi halves [256, 384), lo halves [384, 512); 2 x 16-bit per column
constexpr int kALoCol = 384;

struct Tc4Misc {
  uint64_t w_full[kStages4];
  uint64_t w_empty[kStages4];
  uint64_t a_ready;            // leader only: the next tile's encoding is in shared memory and the accumulator has been read (16 warp arrivals)
  uint64_t acc_full;
  uint64_t acc_free;           // leader only: every warp of both CTAs has the accumulator in registers (16 warp arrivals)
  uint64_t a_chunk[4];         // leader only: K chunk c (64 columns, hi + lo) of the next A operand is in TMEM in both CTAs (8 warp arrivals)
  uint32_t tmem_base;
  uint32_t pad;
  float scratch[4][8];
};
struct Tc4Layout {
  static constexpr int kEncBytes = 2 * kTileBytes;              // encoding tile, hi + lo
  static constexpr int kRingBytes = kStages4 * kTileBytes;
  static constexpr int kParkBytes = 3 * kTileRows * 16;         // warpgroups 1..3 -> warpgroup 0 partial sums (last epilogue)
  static constexpr int kMiscBytes = 1024;
  static constexpr int kTotal = kEncBytes + kRingBytes + kParkBytes + kMiscBytes + 1024 /* alignment slack */;
  static_assert(sizeof(Tc4Misc) <= kMiscBytes, "misc region too small");
  static_assert(kTotal <= 232448, "exceeds 227 KB of shared memory");
};

// D (256 x N, fp32, TMEM) (+)= A (TMEM: this CTA's 128 rows x 16 K-elements, two 16-bit values per column) * B^T (shared memory)
__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait4() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- hidden-layer epilogue: accumulator -> activation -> hi / lo halves of the next A operand in TMEM -----------------
// NG warpgroups share the tile: this thread owns row (TMEM lane) `row` and output columns [256 / NG * g, 256 / NG * (g + 1)).
// Returns its part of the density dot product.  NG = 2: two 32-column blocks per tcgen05.wait::ld (216 registers per thread);
// NG = 4: one block per wait (112 registers per thread, four warps per scheduler hide the round trips).
template <int EPI, bool F16, int NG>
__device__ __forceinline__ float epilogue_hidden4(uint32_t acc_lane, int g, const float* __restrict__ head) {
  constexpr bool kRelu = (EPI != EPI_LINEAR);
  constexpr bool kSigma = (EPI == EPI_RELU_SIGMA || EPI == EPI_SIGMA_OUT);
  constexpr bool kStore = (EPI != EPI_SIGMA_OUT);
  constexpr int kBlocks = 8 / NG;          // 32-column blocks owned by this thread
  constexpr int kBatch = (NG == 2) ? 2 : 1;
  float sg = 0.f;
#pragma unroll 1
  for (int b = 0; b < kBlocks; b += kBatch) {
    uint32_t r0[32], r1[kBatch == 2 ? 32 : 1];
    const int c0 = (256 / NG) * g + 32 * b;
    tmem_ld32(acc_lane + c0, r0);
    if constexpr (kBatch == 2) tmem_ld32(acc_lane + c0 + 32, r1);
    tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int col = c0 + 32 * u;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a, bq;
        if constexpr (kBatch == 2) {
          a = __uint_as_float(u ? r1[2 * i] : r0[2 * i]);
          bq = __uint_as_float(u ? r1[2 * i + 1] : r0[2 * i + 1]);
        } else {
          a = __uint_as_float(r0[2 * i]);
          bq = __uint_as_float(r0[2 * i + 1]);
        }
        if (kRelu) { a = fmaxf(a, 0.f); bq = fmaxf(bq, 0.f); }
        if (kSigma) {
          const float2 w = __ldg(reinterpret_cast<const float2*>(head + kHeadSigmaW + col + 2 * i));
          sg = fmaf(a, w.x, sg);
          sg = fmaf(bq, w.y, sg);
        }
        if (kStore) {
          hi[i] = pack16x2<F16>(a, bq);
          lo[i] = residual16x2<F16>(a, bq, hi[i]);
        }
      }
      if (kStore) {
        // output column n is K element n of the next layer: TMEM column n / 2 of the A region
        tmem_st16(acc_lane + kAHiCol + (col >> 1), hi);
        tmem_st16(acc_lane + kALoCol + (col >> 1), lo);
      }
    }
  }
  return sg;
}

// The pipelined form for the two-warpgroup kernel.  The accumulator and the hi / lo A operand fill all 512 TMEM columns, so
// the next layer cannot own a second accumulator; what CAN overlap is the conversion: the A operand of the finished layer is
// dead, so each 64-column K chunk is converted and published on its own barrier (`arrive_chunk`), and the accumulator is
// released (`arrive_free`) as soon as its last columns are in registers.  The issuer starts the next layer's first K chunk
// while this warpgroup still converts its second one (round 1 / early round 2: one hand-over after all 256 columns).
template <int EPI, bool F16, class ArriveFree, class ArriveChunk>
__device__ __forceinline__ float epilogue_hidden5(uint32_t acc_lane, int g, const float* __restrict__ head, ArriveFree&& arrive_free,
                                                  ArriveChunk&& arrive_chunk, bool late) {
  constexpr bool kRelu = (EPI != EPI_LINEAR);
  constexpr bool kSigma = (EPI == EPI_RELU_SIGMA);
  float sg = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r0[32], r1[32];
    const int c0 = 128 * g + 64 * half;
    tmem_ld32(acc_lane + c0, r0);
    tmem_ld32(acc_lane + c0 + 32, r1);
    tmem_ld_wait();
    if (half == 1 && !late) {
      tc_fence_before();
      arrive_free();
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int col = c0 + 32 * u;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a = __uint_as_float(u ? r1[2 * i] : r0[2 * i]);
        float bq = __uint_as_float(u ? r1[2 * i + 1] : r0[2 * i + 1]);
        if (kRelu) { a = fmaxf(a, 0.f); bq = fmaxf(bq, 0.f); }
        if (kSigma) {
          const float2 w = __ldg(reinterpret_cast<const float2*>(head + kHeadSigmaW + col + 2 * i));
          sg = fmaf(a, w.x, sg);
          sg = fmaf(bq, w.y, sg);
        }
        hi[i] = pack16x2<F16>(a, bq);
        lo[i] = residual16x2<F16>(a, bq, hi[i]);
      }
      tmem_st16(acc_lane + kAHiCol + (col >> 1), hi);
      tmem_st16(acc_lane + kALoCol + (col >> 1), lo);
    }
    tmem_st_wait4();      // this K chunk of the A operand is in TMEM ...
    tc_fence_before();    // ... ordered before the arrive that lets the issuer read it
    if (!late) arrive_chunk(2 * g + half);
  }
  if (late) {             // NB2_TC_DEBUG & 4 (A/B timing): one hand-over after all columns, as before the chunked scheme
    arrive_free();
    arrive_chunk(2 * g);
    arrive_chunk(2 * g + 1);
  }
  return sg;
}

template <bool F16, int NG, int G>
__device__ __forceinline__ void slot_group_run4(const TcParams& p, Tc4Misc* misc, uint32_t enc_base, uint32_t park_base,
                                                uint32_t tmem_base, int64_t n_iters, int warp, int lane, uint32_t rank) {
  const TcNet& net = p.net;
  constexpr int g = G;                // which 256 / NG columns of every layer's output this warpgroup owns
  const int wq = warp & 3;            // TMEM lane quadrant this warp may access
  const int row = wq * 32 + lane;     // row of the tile == TMEM lane
  const uint32_t e_hi = enc_base, e_lo = enc_base + kTileBytes;
  const uint32_t acc = tmem_base + ((uint32_t)(wq * 32) << 16);
  const uint32_t a_ready = smem_u32(&misc->a_ready);
  const uint32_t acc_full = smem_u32(&misc->acc_full);
  float* scratch = &misc->scratch[0][0];
  uint32_t pacc = 0;
  long long t_pe = 0, t_wacc = 0, t_epi = 0, t_last = 0, t_head = 0, t_begin = 0, t0e = NB2_CLK();

  auto arrive_a = [&]() {
    __syncwarp();
    if (lane == 0) {
      if (rank != 0) mbar_arrive_remote_relaxed(a_ready, 0); else mbar_arrive(a_ready);
    }
  };
  auto arrive_on = [&](uint32_t bar) {
    __syncwarp();
    if (lane == 0) {
      if (rank != 0) mbar_arrive_remote_relaxed(bar, 0); else mbar_arrive(bar);
    }
  };
  const uint32_t acc_free = smem_u32(&misc->acc_free);
  const uint32_t a_chunk0 = smem_u32(&misc->a_chunk[0]);
  const bool late = (p.debug & 4) != 0;
  auto arrive_free = [&]() { arrive_on(acc_free); };
  auto arrive_chunk = [&](int c) { arrive_on(a_chunk0 + 8u * (uint32_t)c); };
  auto tile_of = [&](int64_t it) { return it * gridDim.x + blockIdx.x; };   // may lie past n_tiles
  EncRegs<true, F16, (8 / NG) * G, 8 / NG> enc;   // this warpgroup's column groups (8 columns each) of the next tile's encoding
  auto begin_tile = [&]() {
    enc_store(enc, e_hi, e_lo, row);
    fence_proxy_async_smem();
    tc_fence_before();
    arrive_a();
  };
  // ---- alpha compositing over the rows of each ray (nerf_base.py:79-113), warpgroup 0 -------------------------------------
  // Deferred: the rgb-layer epilogue only collects a row's colour sums and density; the sigmoids, the transmittance scan and
  // the stores run one layer-window later, under the NEXT tile's layer-2 MMAs.  Done in place (as it was), this tail made the
  // group ~4 k cycles late for the next tile's layer-0 epilogue -- layer 0 is 12 MMAs -- and the issuer waited 6.8 k cycles
  // before layer 1 against ~2.1 k before the other layers (per-layer wait counters of the PROFILE build).
  struct Pend { float sigma, c0, c1, c2, depth; int s; int64_t ray, grow; bool valid; };
  Pend pend = {0.f, 0.f, 0.f, 0.f, 0.f, 0, 0, 0, false};
  bool has_pend = false;
  auto composite_tail = [&](const Pend& q) {
    const float c0 = 1.f / (1.f + expf(-(q.c0 + __ldg(p.head + kHeadRgbB + 0))));
    const float c1 = 1.f / (1.f + expf(-(q.c1 + __ldg(p.head + kHeadRgbB + 1))));
    const float c2 = 1.f / (1.f + expf(-(q.c2 + __ldg(p.head + kHeadRgbB + 2))));
    const int P = p.io.P;                 // 32, 64 or 128: rays cover whole warps
    const int wpr = P >> 5;               // warps per ray
    const int wseg = wq % wpr;            // this warp's position inside its ray
    const float depth = q.depth;
    float next = __shfl_down_sync(0xffffffffu, depth, 1);
    if (lane == 0) scratch[wq * 8 + 0] = depth;
    named_bar_sync(1, 128);
    if (lane == 31 && wq < 3) next = scratch[(wq + 1) * 8 + 0];
    const bool last = (q.s == P - 1);
    const float delta = last ? 1e10f : __fsub_rn(next, depth);
    const float m = q.valid ? expf(-fmaxf(q.sigma, 0.f) * delta) : 1.f;
    const float alpha = 1.f - m;
    const float inc = warp_scan_mul(m + 1e-10f, lane);
    float exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 1.f;
    if (lane == 31) scratch[wq * 8 + 1] = inc;
    named_bar_sync(1, 128);
    float carry = 1.f;
    for (int w = wq - wseg; w < wq; ++w) carry *= scratch[w * 8 + 1];
    const float wgt = q.valid ? alpha * (carry * exc) : 0.f;
    float sr = warp_sum(wgt * c0), sgn = warp_sum(wgt * c1), sb = warp_sum(wgt * c2);
    float sa = warp_sum(wgt), sd = warp_sum(wgt * depth);
    if (lane == 0) {
      scratch[wq * 8 + 2] = sr; scratch[wq * 8 + 3] = sgn; scratch[wq * 8 + 4] = sb;
      scratch[wq * 8 + 5] = sa; scratch[wq * 8 + 6] = sd;
    }
    named_bar_sync(1, 128);
    if (lane == 0 && wseg == 0 && q.valid) {
      for (int w = wq + 1; w < wq + wpr; ++w) {
        sr += scratch[w * 8 + 2]; sgn += scratch[w * 8 + 3]; sb += scratch[w * 8 + 4];
        sa += scratch[w * 8 + 5]; sd += scratch[w * 8 + 6];
      }
      if (p.io.flags & NB2_WHITE_BKG) {
        const float bg = 1.f - sa;
        sr += bg; sgn += bg; sb += bg;
      }
      p.io.rgb_out[q.ray * 3 + 0] = sr;
      p.io.rgb_out[q.ray * 3 + 1] = sgn;
      p.io.rgb_out[q.ray * 3 + 2] = sb;
      // fused gather: the same row goes to every peer GPU's image buffer over NVLink (nb2_render_params.peer_rgb)
      for (int k = 0; k < p.io.n_peers; ++k) {
        float* o = p.io.peer_rgb[k] + (p.io.peer_row0 + q.ray) * 3;
        o[0] = sr; o[1] = sgn; o[2] = sb;
      }
      if (p.io.depth_out) p.io.depth_out[q.ray] = (sd - p.io.near_t) / (p.io.far_t - p.io.near_t);
      if (p.io.acc_out) p.io.acc_out[q.ray] = sa;
    }
    named_bar_sync(1, 128);  // scratch is reused by the next tile
  };

  // every warpgroup: meet, then warpgroup 0 adds the parked partial sums of the others and finishes the row
  auto finish_rgb = [&](const Pend& q0) {
    named_bar_sync(3, 128 * NG);
    if (g != 0) return;
    Pend q = q0;
    const uint32_t xaddr = park_base + (uint32_t)row * 16u;
#pragma unroll
    for (int k = 0; k < NG - 1; ++k) {
      uint32_t x0, x1, x2, x3;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                   : "r"(xaddr + (uint32_t)k * (kTileRows * 16)));
      q.sigma += __uint_as_float(x0); q.c0 += __uint_as_float(x1); q.c1 += __uint_as_float(x2); q.c2 += __uint_as_float(x3);
    }
    if (p.io.out_mode == 1) {
      const float c0 = 1.f / (1.f + expf(-(q.c0 + __ldg(p.head + kHeadRgbB + 0))));
      const float c1 = 1.f / (1.f + expf(-(q.c1 + __ldg(p.head + kHeadRgbB + 1))));
      const float c2 = 1.f / (1.f + expf(-(q.c2 + __ldg(p.head + kHeadRgbB + 2))));
      if (q.valid) reinterpret_cast<float4*>(p.io.out)[q.grow] = make_float4(c0, c1, c2, q.sigma);
    } else {
      composite_tail(q);
    }
  };

  RowIn in = load_row(p.io, tile_of(0) * kTileRows + row);
  enc_compute(enc, in.p, p.pos_levels, in.valid, in.enc, in.ipe ? in.cov : nullptr);
  begin_tile();

  for (int64_t it = 0; it < n_iters; ++it) {
    const int64_t grow = tile_of(it) * kTileRows + row;
    const bool has_next = (it + 1 < n_iters);
    RowIn in_next = in;
    if (has_next) in_next = load_row(p.io, tile_of(it + 1) * kTileRows + row);

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
#pragma unroll 1
        for (int cb = (4 / NG) * g; cb < (4 / NG) * (g + 1); ++cb) {
          uint32_t r[32];
          tmem_ld32(acc + cb * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int gq = 0; gq < 8; ++gq) {
            const int col = cb * 32 + gq * 4;
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.head + kHeadRgbW + col));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.head + kHeadRgbW + 128 + col));
            const float4 w2 = __ldg(reinterpret_cast<const float4*>(p.head + kHeadRgbW + 256 + col));
            const float t0 = fmaxf(__uint_as_float(r[gq * 4 + 0]), 0.f), t1 = fmaxf(__uint_as_float(r[gq * 4 + 1]), 0.f);
            const float t2 = fmaxf(__uint_as_float(r[gq * 4 + 2]), 0.f), t3 = fmaxf(__uint_as_float(r[gq * 4 + 3]), 0.f);
            c0 = fmaf(t0, w0.x, c0); c0 = fmaf(t1, w0.y, c0); c0 = fmaf(t2, w0.z, c0); c0 = fmaf(t3, w0.w, c0);
            c1 = fmaf(t0, w1.x, c1); c1 = fmaf(t1, w1.y, c1); c1 = fmaf(t2, w1.z, c1); c1 = fmaf(t3, w1.w, c1);
            c2 = fmaf(t0, w2.x, c2); c2 = fmaf(t1, w2.y, c2); c2 = fmaf(t2, w2.z, c2); c2 = fmaf(t3, w2.w, c2);
          }
        }
        // the accumulator has been read and the encoding tile is free: hand the next tile to the issuer before the rest
        // of this epilogue (collecting the heads, compositing), which runs one layer-window later (finish_rgb)
        const long long ch = NB2_CLK();
        t_head += ch - ce;
        if (has_next) begin_tile();
        t_begin += NB2_CLK() - ch;
        // combine the column groups: warpgroups 1.. park their partial sums (and their part of the density dot product);
        // warpgroup 0 collects them in finish_rgb -- one layer-window later when another tile follows
        const uint32_t xaddr = park_base + (uint32_t)row * 16u;
        if (g != 0) st_shared_v4(xaddr + (uint32_t)(g - 1) * (kTileRows * 16), __float_as_uint(sigma), __float_as_uint(c0),
                                 __float_as_uint(c1), __float_as_uint(c2));
        const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(in.d[0], in.d[0]), __fmul_rn(in.d[1], in.d[1])), __fmul_rn(in.d[2], in.d[2])));
        pend = Pend{sigma, c0, c1, c2, __fmul_rn(in.z, nrm), in.s, in.ray, grow, in.valid};
        has_pend = true;
        if (!has_next) {                  // the CTA's last tile: nothing left to hide under
          finish_rgb(pend);
          has_pend = false;
        }
        t_last += NB2_CLK() - ce;
        continue;
      }
      if (epi == EPI_SIGMA_OUT) {
        // ---- proposal tail: out = w_s . relu(acc) + b_s ------------------------------------------------------------------
        sigma = epilogue_hidden4<EPI_SIGMA_OUT, F16, NG>(acc, g, p.head) + (g == 0 ? __ldg(p.head + kHeadSigmaB) : 0.f);
        if (has_next) begin_tile();
        const uint32_t xaddr = park_base + (uint32_t)row * 16u;
        if (g != 0) asm volatile("st.shared.b32 [%0], %1;" ::"r"(xaddr + (uint32_t)(g - 1) * (kTileRows * 16)), "r"(__float_as_uint(sigma)));
        named_bar_sync(3, 128 * NG);
        if (g == 0) {
#pragma unroll
          for (int q = 0; q < NG - 1; ++q) {
            uint32_t x0;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x0) : "r"(xaddr + (uint32_t)q * (kTileRows * 16)));
            sigma += __uint_as_float(x0);
          }
        }
        if (in.valid && g == 0) p.io.out[grow] = sigma;
        t_last += NB2_CLK() - ce;
        continue;
      }
      static_assert(NG == 2, "the chunked hand-over is written for two warpgroups per tile");
      if (epi == EPI_RELU) {
        epilogue_hidden5<EPI_RELU, F16>(acc, g, p.head, arrive_free, arrive_chunk, late);
      } else if (epi == EPI_LINEAR) {
        epilogue_hidden5<EPI_LINEAR, F16>(acc, g, p.head, arrive_free, arrive_chunk, late);
      } else {   // EPI_RELU_SIGMA (each warpgroup keeps the dot product over its own columns; the bias is added once)
        sigma = epilogue_hidden5<EPI_RELU_SIGMA, F16>(acc, g, p.head, arrive_free, arrive_chunk, late) + (g == 0 ? __ldg(p.head + kHeadSigmaB) : 0.f);
      }
      t_epi += NB2_CLK() - ce;

      // ---- work hidden behind the next layer's MMAs ------------------------------------------------------------------
      const long long cpe = NB2_CLK();
      if (has_next && l == 0) enc_compute(enc, in_next.p, p.pos_levels, in_next.valid, in_next.enc, in_next.ipe ? in_next.cov : nullptr);
      if (l == 1 && has_pend) {           // the previous tile's heads and compositing, under this tile's layer-2 MMAs
        finish_rgb(pend);
        has_pend = false;
      }
      if (l == p.dir_layer && g == 0) {
        // the encoded position is dead after the skip layer: its tile now takes the encoded direction (columns 0-31;
        // the bias k-step of the running layer reads columns 48-63 of the same rows, other 16-byte units)
        float rot[3] = {0.f, 0.f, 0.f};
        if (in.valid) normalize_dir(in.d, rot);
        write_enc_row<true, F16, kDirCols, kMaxDirLevels>(e_hi, e_lo, row, rot, p.dir_levels, in.valid);
        fence_proxy_async_smem();
      }
      t_pe += NB2_CLK() - cpe;
    }
    in = in_next;
  }
  if (NB2_PROF_ON && threadIdx.x == kRolesThreads) {
    long long* o = p.prof + blockIdx.x * 16;
    o[6] = t_pe; o[7] = t_wacc; o[8] = t_epi; o[9] = t_last; o[10] = NB2_CLK() - t0e; o[11] = n_iters; o[12] = net.n_layers;
    if ((blockIdx.x & 1) == 0) { o[13] = t_head; o[14] = t_begin; }     // (the odd rows' free slots carry the issuer's per-layer waits)
  }
}

// setmaxnreg moves registers inside the CTA's own allocation: what the slot groups gain must equal what warps 0-3 give up.
// NG = 2: 384 threads launched at 168: roles 72 (-96 x 128), groups 216 (+48 x 256).
// NG = 4: 640 threads launched at 96: roles 64 (-32 x 128), groups 104 (+8 x 512).
template <int NG> struct Tc4Regs { static constexpr int role = (NG == 2) ? kRoleRegs : 64, group = (NG == 2) ? kGroupRegs : 104; };
template <bool F16, int NG>
__global__ void __launch_bounds__(128 + 128 * NG, 1) mlp_tc4_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t enc_base = smem_base;
  const uint32_t ring_base = smem_base + Tc4Layout::kEncBytes;
  const uint32_t park_base = ring_base + Tc4Layout::kRingBytes;
  Tc4Misc* misc = reinterpret_cast<Tc4Misc*>(smem_al + Tc4Layout::kEncBytes + Tc4Layout::kRingBytes + Tc4Layout::kParkBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcNet& net = p.net;
  const int64_t n_iters = (p.n_tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages4; ++i) {
      // leader: a stage is full when its own bulk copy has landed AND the peer has relayed that its half has landed
      mbar_init(smem_u32(&misc->w_full[i]), rank == 0 ? 2 : 1);
      mbar_init(smem_u32(&misc->w_empty[i]), 1);
    }
    mbar_init(smem_u32(&misc->a_ready), 8 * NG);   // every slot-group warp of both CTAs
    mbar_init(smem_u32(&misc->acc_full), 1);
    mbar_init(smem_u32(&misc->acc_free), 8 * NG);
    for (int c = 0; c < 4; ++c) mbar_init(smem_u32(&misc->a_chunk[c]), 2 * (8 / NG));   // the four warps that own the chunk, both CTAs
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
    // =========================== weight streamer: this CTA's half of every tile ==========================
    // per layer: K chunks in table order, each as [Wh, Wl] (cross-term sweep), then every Wh again (main sweep)
    reg_dealloc<Tc4Regs<NG>::role>();
    if (lane == 0 && !(p.debug & 1)) {
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < n_iters; ++it) {
        for (int l = 0; l < net.n_layers; ++l) {
          const TcLayer& L = net.layer[l];
          // nc == 2: this CTA owns N-chunk `rank` (16 KB per K-chunk); nc == 1: rows [64 rank, 64 rank + 64) (8 KB)
          const int c_first = L.chunk0 + (L.nc == 2 ? (int)rank * L.kc : 0);
          const uint32_t bytes = L.nc == 2 ? kTileBytes : kTileBytes / 2;
          const size_t sub = L.nc == 2 ? 0 : (size_t)rank * (kTileBytes / 4);   // in 16-bit elements
          for (int e = 0; e < 3 * L.kc; ++e) {
            const int k = e < 2 * L.kc ? (e >> 1) : e - 2 * L.kc;
            const int part = e < 2 * L.kc ? (e & 1) : 0;
            // a bias-only chunk (ks0 != 0: its A column is the constant 1, whose lo half is 0) has no lo x Wh term: its Wh
            // tile is not streamed in the cross-term sweep (the issuer and the relay skip the same entry)
            if (e < 2 * L.kc && part == 0 && L.ks0[k] != 0) continue;
            mbar_wait(smem_u32(&misc->w_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&misc->w_full[stage]);
            mbar_arrive_expect_tx(full, bytes);
            bulk_g2s(ring_base + stage * kTileBytes,
                     p.wchunks + ((size_t)(c_first + k) * 4 + (F16 ? 2 : 0) + part) * (kTileBytes / 2) + sub, bytes, full);
            if (++stage == kStages4) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    reg_dealloc<Tc4Regs<NG>::role>();
    if (lane == 0 && rank != 0 && !(p.debug & 1)) {
      // =========================== peer: relay "my half has landed" to the leader =====================
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < n_iters; ++it)
        for (int l = 0; l < net.n_layers; ++l) {
          int n_entries = 3 * net.layer[l].kc;
          for (int k = 0; k < net.layer[l].kc; ++k) n_entries -= (net.layer[l].ks0[k] != 0);   // bias-only chunks: no Wh in sweep 1
          for (int e = 0; e < n_entries; ++e) {
            mbar_wait(smem_u32(&misc->w_full[stage]), phase);
            mbar_arrive_remote_relaxed(smem_u32(&misc->w_full[stage]), 0);
            if (++stage == kStages4) { stage = 0; phase ^= 1u; }
          }
        }
    } else if (rank == 0) {
      // =========================== leader: MMA issuer for the pair (whole warp, one elected lane per instruction) ======
      uint32_t stage = 0, phase = 0, pa = 0, pf = 0, pc = 0;
      long long t_wa = 0, t_ww = 0, t0m = NB2_CLK();
      long long t_wl[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // PROFILE build: the issuer's operand waits by layer
      const uint32_t ring_lo = umma_desc_lo(ring_base);
      const uint32_t enc_lo = umma_desc_lo(enc_base);    // hi half of the encoding tile; the lo half is one tile further
      const bool no_weights = (p.debug & 1) != 0, no_mma = (p.debug & 2) != 0;
      auto wait_stage = [&]() {
        if (no_weights) return;
        const long long c0 = NB2_CLK();
        mbar_wait(smem_u32(&misc->w_full[stage]), phase);
        t_ww += NB2_CLK() - c0;
        tc_fence_after();
      };
      auto release_stage = [&]() {
        if (!no_weights) umma2_commit_mcast_elect(smem_u32(&misc->w_empty[stage]), 3);
        if (++stage == kStages4) { stage = 0; phase ^= 1u; }
      };
      // one K chunk (up to four k-steps) of A (hi or lo half) against the weight stage that just landed
      auto issue_chunk = [&](int a_src, int ks0, bool lo_half, uint32_t idesc, bool first) {
        if (no_mma) return;
        const uint32_t w0 = ring_lo + stage * (kTileBytes >> 4);
        if (a_src == kChunkE) {
          const uint32_t a0 = enc_lo + (lo_half ? (uint32_t)(kTileBytes >> 4) : 0u);
          for (int ks = ks0; ks < 4; ++ks)
            umma2_f16_ss_elect(tmem_base, umma_desc_from_lo(a0 + 2 * ks), umma_desc_from_lo(w0 + 2 * ks), idesc,
                          (uint32_t)(!(first && ks == ks0)));
        } else {
          const uint32_t a0 = tmem_base + (uint32_t)(lo_half ? kALoCol : kAHiCol) + (uint32_t)(32 * a_src);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma2_f16_ts_elect(tmem_base, a0 + 8 * ks, umma_desc_from_lo(w0 + 2 * ks), idesc, (uint32_t)(!(first && ks == 0)));
        }
      };
      for (int64_t it = 0; it < n_iters; ++it) {
        for (int l = 0; l < net.n_layers; ++l) {
          const TcLayer& L = net.layer[l];
          const uint32_t idesc = umma_idesc_16(256, L.nc * 128, F16);
          // layer 0 reads the encoding tile (a_ready: written, and the previous tile's accumulator read); every other layer
          // may overwrite the accumulator once it is in registers and reads K chunk c of its A operand once c is published
          { const long long c0 = NB2_CLK();
          if (l == 0) {
            mbar_wait(smem_u32(&misc->a_ready), pa);
            pa ^= 1u;
          } else {
            mbar_wait(smem_u32(&misc->acc_free), pf);
            pf ^= 1u;
          }
          t_wa += NB2_CLK() - c0; if (NB2_PROF_ON) t_wl[l < 9 ? l : 8] += NB2_CLK() - c0; }
          tc_fence_after();
          uint32_t pending = l == 0 ? 0u : 0xFu;       // chunk barriers of this layer not consumed yet
          auto wait_chunk = [&](int c) {
            if (!(pending >> c & 1u)) return;
            const long long c0 = NB2_CLK();
            mbar_wait(smem_u32(&misc->a_chunk[c]), pc);
            t_wa += NB2_CLK() - c0;
            if (NB2_PROF_ON) t_wl[l < 9 ? l : 8] += NB2_CLK() - c0;
            tc_fence_after();
            pending &= ~(1u << c);
          };
          // sweep 1: the small cross terms lo x Wh and hi x Wl
          for (int k = 0; k < L.kc; ++k) {
            if (L.a_src[k] != kChunkE) wait_chunk(L.a_src[k]);
            // (bias-only chunk: its A column is the constant 1, whose lo half is 0 -- no lo x Wh term, no Wh stage)
            if (L.ks0[k] == 0) {
              wait_stage();
              issue_chunk(L.a_src[k], 0, true, idesc, k == 0);
              release_stage();
            }
            wait_stage();
            issue_chunk(L.a_src[k], L.ks0[k], false, idesc, false);
            release_stage();
          }
          for (int c = 0; c < 4; ++c) wait_chunk(c);   // (a layer that does not read every chunk still consumes each phase)
          if (l != 0) pc ^= 1u;
          // sweep 2: hi x Wh on top
          for (int k = 0; k < L.kc; ++k) {
            wait_stage();
            issue_chunk(L.a_src[k], L.ks0[k], false, idesc, false);
            release_stage();
          }
          umma2_commit_mcast_elect(smem_u32(&misc->acc_full), 3);
        }
      }
      if (NB2_PROF_ON && lane == 0) { p.prof[blockIdx.x * 16 + 3] = t_wa; p.prof[blockIdx.x * 16 + 4] = t_ww; p.prof[blockIdx.x * 16 + 5] = NB2_CLK() - t0m; }
      if (NB2_PROF_ON && lane == 0) {
        // the per-layer split goes to the free slots of the PEER's row (0-5, 13-15: it has no streamer / issuer counters)
        const int slot[9] = {0, 1, 2, 3, 4, 5, 13, 14, 15};
        for (int l = 0; l < 9; ++l) p.prof[(blockIdx.x + 1) * 16 + slot[l]] = t_wl[l];
      }
    }
  } else if (warp >= 4) {
    reg_alloc<Tc4Regs<NG>::group>();
    const int wgi = (warp - 4) >> 2;
    if (wgi == 0) slot_group_run4<F16, NG, 0>(p, misc, enc_base, park_base, tmem_base, n_iters, warp, lane, rank);
    else if (wgi == 1) slot_group_run4<F16, NG, 1>(p, misc, enc_base, park_base, tmem_base, n_iters, warp, lane, rank);
    else if (NG == 4 && wgi == 2) slot_group_run4<F16, NG, (NG == 4 ? 2 : 0)>(p, misc, enc_base, park_base, tmem_base, n_iters, warp, lane, rank);
    else if (NG == 4) slot_group_run4<F16, NG, (NG == 4 ? 3 : 0)>(p, misc, enc_base, park_base, tmem_base, n_iters, warp, lane, rank);
  } else {
    reg_dealloc<Tc4Regs<NG>::role>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---- self-test of the A-from-TMEM operand convention --------------------------------------------------------------------
// D (128 x 128 fp32) = A (128 x 64 bf16) * B^T (128 x 64 bf16) with A written to TMEM by tcgen05.st exactly as the epilogue
// above writes it (row = lane, two consecutive K elements per 32-bit column, low half first) and consumed by the
// A-from-TMEM form of tcgen05.mma (cta_group::1 here); B arrives pre-swizzled through cp.async.bulk like the weights.
__global__ void __launch_bounds__(128, 1)
umma_ts_selftest_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bswz, float* __restrict__ D) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* al = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t b_tile = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(al + kTileBytes);  // [0] = B landed, [1] = MMA done
  uint32_t* tptr = reinterpret_cast<uint32_t*>(al + kTileBytes + 64);
  const int warp = threadIdx.x >> 5, row = threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tptr), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
      r[i] = pack_bf16x2(__bfloat162float(A[row * 64 + 32 * h + 2 * i]), __bfloat162float(A[row * 64 + 32 * h + 2 * i + 1]));
    tmem_st16(lane_addr + 128 + 16 * h, r);
  }
  tmem_st_wait4();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    mbar_arrive_expect_tx(smem_u32(&bars[0]), kTileBytes);
    bulk_g2s(b_tile, Bswz, kTileBytes, smem_u32(&bars[0]));
    mbar_wait(smem_u32(&bars[0]), 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_16(128, 128, false);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t acc_flag = (uint32_t)(ks != 0);
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
          ::"r"(tmem), "r"(tmem + 128 + 8 * ks), "l"(umma_smem_desc(b_tile + ks * 32)), "r"(idesc), "r"(acc_flag)
          : "memory");
    }
    umma_commit(smem_u32(&bars[1]));
  }
  mbar_wait(smem_u32(&bars[1]), 0);
  __syncwarp();
  tc_fence_after();
  for (int cb = 0; cb < 4; ++cb) {
    uint32_t r[32];
    tmem_ld32(lane_addr + cb * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[row * 128 + cb * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

__global__ void swizzle_tile_kernel4(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kTileRows * kTileCols) return;
  dst[swz128_offset(i / kTileCols, i % kTileCols) / 2] = src[i];
}

int selftest_umma_ts(nb2_handle* h, const void* A, const void* B, void* Bswz_scratch, float* D, cudaStream_t st) {
  swizzle_tile_kernel4<<<(kTileRows * kTileCols + 255) / 256, 256, 0, st>>>((const __nv_bfloat16*)B, (__nv_bfloat16*)Bswz_scratch);
  NB2_LAUNCH_CHECK(h);
  const int smem = kTileBytes + 1024 + 1024;
  {
    const int rc = kernel_set_smem(h, (const void*)umma_ts_selftest_kernel, smem);
    if (rc != NB2_OK) return rc;
  }
  umma_ts_selftest_kernel<<<1, 128, smem, st>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)Bswz_scratch, D);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

// ---- host side --------------------------------------------------------------------------------------------------------
template <bool F16>
static int launch_tc4_impl(nb2_handle* h, TcParams& prm, cudaStream_t st) {
  auto kern = mlp_tc4_kernel<F16, 2>;
  int rc = kernel_set_smem(h, (const void*)kern, Tc4Layout::kTotal);
  if (rc != NB2_OK) return rc;
  int64_t ctas = (prm.n_tiles + 1) / 2 * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(128 + 128 * 2);
  cfg.dynamicSmemBytes = Tc4Layout::kTotal;
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

int launch_mlp_tc4(nb2_handle* h, const TcParams& base, int precision, cudaStream_t st) {
  TcParams prm = base;
  for (int l = 0; l < prm.net.n_layers; ++l) {
    const TcLayer& L = prm.net.layer[l];
    for (int k = 0; k < L.kc; ++k)
      if (L.a_src[k] != kChunkE && L.ks0[k] != 0) {
        set_error("mlp_forward: layer %d reads a partial hidden chunk, which the TMEM-operand kernel does not support", l);
        return NB2_ERR_UNSUPPORTED;
      }
  }
  // (two epilogue warpgroups per tile; the four-warpgroup variant of round 1 spilled at 104 registers and was retired)
  if (precision == NB2_PREC_FP16X3) return launch_tc4_impl<true>(h, prm, st);
  if (precision == NB2_PREC_BF16X3) return launch_tc4_impl<false>(h, prm, st);
  set_error("mlp_forward: the TMEM-operand kernel runs the split precisions only (got %d)", precision);
  return NB2_ERR_INVALID;
}

}  // namespace nb2
