// nb2_mlp_tc3.cu — the per-sample MLP as a CTA-pair tcgen05 kernel whose layers are pipelined by OUTPUT HALVES.
//
// STATUS: correct (runs the whole GPU suite with NB2_TC_NHALF=1) but SLOWER than the defaults, kept opt-in for the record:
// an N = 128 tcgen05.mma with both operands in shared memory takes ~100 cycles instead of 64 (the shared-memory operand port
// delivers ~64 B/cycle, which M256 x N256 x K16 exactly saturates), so halving N costs more than the hidden epilogue returns.
//
// Measured on B200 (profiles/r01_roles_pair2wg.txt, r01_microbench_tmem_l2.txt): in the layer-serial pair kernel
// (nb2_mlp_tc.cu) the MMA issuer spends 36-50 % of every iteration waiting for the epilogue, although neither TMEM
// reads (~900 B/cycle/SM) nor the L2 weight stream (60 B/cycle/SM) are near a limit: the time is the per-layer
// round trip  commit -> tcgen05.ld -> convert -> st.shared -> fence -> arrive -> first MMA of the next layer.
// This kernel takes most of that round trip off the critical path:
//   * every 256-wide layer is issued as two N = 128 halves (M = 256 across the pair, K = 16), each with its own
//     accumulator columns and its own completion barrier;
//   * while half 1 accumulates, the slot groups drain half 0 from TMEM, apply the activation and keep the
//     16-bit result in REGISTERS (the activation tiles are still being read by half 1's MMAs);
//   * when half 1 completes they store those registers (K chunks 0-1 of the next layer's A operand), signal "chunks
//     0-1 ready", and only then drain and store half 1 (K chunks 2-3), signalling "rest ready";
//   * the issuer starts the next layer's half 0 on K chunks 0-1 as soon as they are ready, and continues with the
//     other chunks when the rest is.  Exposed per layer: one store + fence + arrive, instead of the whole epilogue.
//   * the next tile's encoding and the direction encoding are computed in the windows where a slot group would
//     otherwise spin on an accumulator barrier, and are stored when their tile is free.
// Biases ride on the tensor core as in mlp_tc2_kernel (a first version added them in the epilogue to save the bias
// chunks' weight bytes: 128 dependent L1 loads per thread and layer made the epilogue 3x slower).
//   * TWO threads issue MMAs (leader CTA, warps 1 and 3): an N = 128 instruction occupies the tensor pipe for 64 cycles,
//     but one thread sustains only one tcgen05.mma per ~110 cycles.  Single pass: one issuer per resident tile;
//     split: one issuer per accumulator (main / correction), so no accumulator is ever written from two threads.
// Roles, operand layouts, weight image, precisions and the last-layer epilogues are those of mlp_tc2_kernel.
#include <stdlib.h>

#include "nb2_tc_device.cuh"

namespace nb2 {
using namespace ptx;

constexpr int kStages3 = 8;                     // weight ring: 8 x 8 KB (this CTA's 64 rows of a 128(n) x 64(k) tile)
constexpr int kStageBytes3 = kTileBytes / 2;

// Per-layer issue plan (host-built, lives in the kernel parameters): K chunks in issue order — first those reading
// H chunks 0-1 (available early), then the rest (encoding chunk, bias-only chunk).
struct Tc3Layer {
  unsigned char n, n_early, pad0, pad1;
  unsigned char k[6];     // K chunk (index into TcLayer::a_src / the layer's weight tiles) per issue position
  unsigned char ks0[6];   // first 16-wide k-step of that chunk (3 for the bias-only chunk)
};
struct Tc3Params {
  TcParams base;
  Tc3Layer plan[kMaxTcLayers];
};

struct Tc3Misc {
  uint64_t w_full[kStages3];
  uint64_t w_empty[kStages3];
  uint64_t w_peer[kStages3];    // leader only: the peer CTA's half of the stage has landed
  uint64_t a_ready[2];          // leader only: [0] H chunks 0-1 of all resident tiles written, [1] everything else
  uint64_t acc_full[2];         // N-half h of all resident tiles accumulated
  uint32_t tmem_base;
  uint32_t pad;
  float scratch[2][4][8];
};
static_assert(sizeof(Tc3Misc) <= 1024, "misc region too small");

// ---- one 32-column block of an N-half: fp32 accumulator (+ correction) -> (+ bias) -> activation -> packed 16-bit ------
template <int EPI, bool SPLIT, bool F16>
__device__ __forceinline__ void block_pack(uint32_t (&m)[32], const uint32_t (&c)[32], int col,
                                           const float* __restrict__ head, float& sg, uint32_t (&hi)[16], uint32_t (&lo)[16]) {
  constexpr bool kRelu = (EPI != EPI_LINEAR);
  constexpr bool kSigma = (EPI == EPI_RELU_SIGMA || EPI == EPI_SIGMA_OUT);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = __uint_as_float(m[4 * q + i]);
      if (SPLIT) v[i] += __uint_as_float(c[4 * q + i]);
    }
    if (kRelu && (SPLIT || kSigma)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (kSigma) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(head + kHeadSigmaW + col + 4 * q));
      sg = fmaf(v[0], w.x, sg); sg = fmaf(v[1], w.y, sg); sg = fmaf(v[2], w.z, sg); sg = fmaf(v[3], w.w, sg);
    }
    if (EPI != EPI_SIGMA_OUT) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        uint32_t h;
        if (!SPLIT && kRelu && !kSigma) h = pack16x2_relu<F16>(a, b); else h = pack16x2<F16>(a, b);
        hi[2 * q + i] = h;
        if (SPLIT) lo[2 * q + i] = residual16x2<F16>(a, b, h);
      }
    }
  }
}

// 32 packed columns (16 + 16 registers) -> rows of the A-operand chunk that holds absolute column `col`
template <bool SPLIT>
__device__ __forceinline__ void store_block(uint32_t slot_base, uint32_t lo_off, int row, int col, const uint32_t (&hi)[16],
                                            const uint32_t (&lo)[16]) {
  const uint32_t tile = slot_base + (uint32_t)(col >> 6) * kTileBytes;
  const uint32_t g0 = ((uint32_t)col & 63u) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t off = (uint32_t)row * 128u + (((g0 + q) ^ ((uint32_t)row & 7u)) << 4);
    st_shared_v4(tile + off, hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
    if (SPLIT) st_shared_v4(tile + lo_off + off, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
  }
}

// W columns of one N-half held in registers between "half 0 drained" and "half 1 complete"
template <bool SPLIT, int W>
struct HeldHalf {
  uint32_t hi[W / 32][16];
  uint32_t lo[SPLIT ? W / 32 : 1][16];
};

// Drain W columns starting at absolute column c0 and keep the packed result (half 0: off the critical path).
template <int EPI, bool SPLIT, bool F16, int W>
__device__ __forceinline__ void drain_hold(uint32_t acc, int c0, const float* head, float& sg, HeldHalf<SPLIT, W>& held) {
  constexpr int NB = W / 32;
  if (SPLIT) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      uint32_t m[32], c[32];
      tmem_ld32(acc + c0 + 32 * b, m);
      tmem_ld32(acc + 256 + c0 + 32 * b, c);
      tmem_ld_wait();
      block_pack<EPI, SPLIT, F16>(m, c, c0 + 32 * b, head, sg, held.hi[b], held.lo[b]);
    }
  } else {
#pragma unroll
    for (int b = 0; b < NB; b += 2) {
      uint32_t m0[32], m1[32];
      tmem_ld32(acc + c0 + 32 * b, m0);
      tmem_ld32(acc + c0 + 32 * b + 32, m1);
      tmem_ld_wait();
      block_pack<EPI, SPLIT, F16>(m0, m0, c0 + 32 * b, head, sg, held.hi[b], held.lo[0]);
      block_pack<EPI, SPLIT, F16>(m1, m1, c0 + 32 * b + 32, head, sg, held.hi[b + 1], held.lo[0]);
    }
  }
}

// ---- slot group -------------------------------------------------------------------------------------------------------
// G: (two warpgroups per tile only) which 64 columns of each N-half this warpgroup owns; compile-time so that each
// instantiation keeps exactly one set of encoding registers and constant column offsets.
template <int NSLOTS, bool SPLIT, bool F16, int G>
__device__ __forceinline__ void slot_group_run3(const Tc3Params& q, Tc3Misc* misc, uint32_t act_base, uint32_t tmem_base,
                                                int64_t n_iters, int warp, int lane, uint32_t rank) {
  using LT = TcLayout<NSLOTS, SPLIT>;
  const TcParams& p = q.base;
  const TcNet& net = p.net;
  constexpr int EW = 2 / NSLOTS;        // warpgroups sharing one tile
  constexpr int W = 128 / EW;           // columns of each N-half owned by one thread
  const int wg = (warp - 4) >> 2;
  const int s = (EW == 2) ? 0 : wg;     // resident tile
  constexpr int g = G;                  // column group inside each N-half
  const int wq = warp & 3;              // TMEM lane quadrant
  const int row = wq * 32 + lane;
  const uint32_t slot_base = act_base + s * LT::kSlotBytes;
  const uint32_t lo_off = kChunksPerSlot * kTileBytes;
  const uint32_t e_hi = slot_base + kChunkE * kTileBytes, e_lo = e_hi + lo_off;
  const uint32_t acc = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(s * 256);
  const uint32_t ready0 = smem_u32(&misc->a_ready[0]), ready1 = smem_u32(&misc->a_ready[1]);
  const uint32_t full0 = smem_u32(&misc->acc_full[0]), full1 = smem_u32(&misc->acc_full[1]);
  float* scratch = &misc->scratch[s][0][0];
  uint32_t ph0 = 0, ph1 = 0;
  long long t_wacc0 = 0, t_wacc1 = 0, t_h0 = 0, t_st0 = 0, t_h1 = 0, t_enc = 0, t_last = 0, t0e = NB2_CLK();

  // one arrival per warp (every lane fenced its own shared-memory writes / TMEM reads before the __syncwarp)
  auto arrive = [&](uint32_t bar) {
    __syncwarp();
    if (lane == 0) {
      if (rank != 0) mbar_arrive_remote_relaxed(bar, 0); else mbar_arrive(bar);
    }
  };
  auto publish = [&](uint32_t bar) {   // shared-memory writes -> visible to the tensor core, then signal
    fence_proxy_async_smem();
    tc_fence_before();
    arrive(bar);
  };
  auto wait_acc = [&](uint32_t bar, uint32_t& ph, long long& t) {
    const long long c0 = NB2_CLK();
    mbar_wait(bar, ph);
    ph ^= 1u;
    __syncwarp();
    tc_fence_after();
    t += NB2_CLK() - c0;
  };

  auto tile_of = [&](int64_t it) { return (it * gridDim.x + blockIdx.x) * NSLOTS + s; };
  // EW == 1: this thread produces both halves of its row's encoding (two phases: column groups 0-3, then 4-7);
  // EW == 2: warpgroup g produces groups 4g..4g+3 in phase 0 (enc_b is unused and costs no registers)
  EncRegs<SPLIT, F16, 4 * G, 4> enc_a;
  EncRegs<SPLIT, F16, 4, 4> enc_b;
  auto enc_phase = [&](const RowIn& r, int phase) {
    if (EW == 2 || phase == 0) enc_compute(enc_a, r.p, p.pos_levels, r.valid, r.enc);
    else enc_compute(enc_b, r.p, p.pos_levels, r.valid, r.enc);
  };
  auto begin_tile = [&]() {
    enc_store(enc_a, e_hi, e_lo, row);
    if (EW == 1) enc_store(enc_b, e_hi, e_lo, row);
    fence_proxy_async_smem();
    tc_fence_before();
    arrive(ready0);
    arrive(ready1);
  };

  RowIn in = load_row(p.io, tile_of(0) * kTileRows + row);
  enc_phase(in, 0);
  if (EW == 1) enc_phase(in, 1);
  begin_tile();

  for (int64_t it = 0; it < n_iters; ++it) {
    const int64_t grow = tile_of(it) * kTileRows + row;
    const bool has_next = (it + 1 < n_iters);
    RowIn in_next = in;
    if (has_next) in_next = load_row(p.io, tile_of(it + 1) * kTileRows + row);

    float sigma = 0.f;
    for (int l = 0; l < net.n_layers; ++l) {
      const int epi = net.layer[l].epi;

      if (epi == EPI_RGB) {
        // ---- rgb_layer (N = 128: half 0 only): t = relu(acc), rgb = sigmoid(W1 t + b1), then alpha compositing ----------
        wait_acc(full0, ph0, t_wacc0);
        const long long ce = NB2_CLK();
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        bool continue_flag = true;
#pragma unroll 1
        for (int cb = g * (4 / EW); cb < g * (4 / EW) + 4 / EW; ++cb) {
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
        if (EW == 2) {
          // group 1 parks its partial sums in the tile's first activation chunk (dead: this layer's MMAs completed)
          const uint32_t xaddr = slot_base + (uint32_t)row * 16u;
          if (g == 1) st_shared_v4(xaddr, __float_as_uint(sigma), __float_as_uint(c0), __float_as_uint(c1), __float_as_uint(c2));
          named_bar_sync(3, 256);
          if (g == 1) continue_flag = false;
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
            float sr = warp_sum(wgt * c0), sgn = warp_sum(wgt * c1), sb = warp_sum(wgt * c2);
            float sa = warp_sum(wgt), sd = warp_sum(wgt * depth);
            if (lane == 0) {
              scratch[wq * 8 + 2] = sr; scratch[wq * 8 + 3] = sgn; scratch[wq * 8 + 4] = sb;
              scratch[wq * 8 + 5] = sa; scratch[wq * 8 + 6] = sd;
            }
            named_bar_sync(1 + s, 128);
            if (lane == 0 && wseg == 0 && in.valid) {
              for (int w = wq + 1; w < wq + wpr; ++w) {
                sr += scratch[w * 8 + 2]; sgn += scratch[w * 8 + 3]; sb += scratch[w * 8 + 4];
                sa += scratch[w * 8 + 5]; sd += scratch[w * 8 + 6];
              }
              if (p.io.flags & NB2_WHITE_BKG) {
                const float bg = 1.f - sa;
                sr += bg; sgn += bg; sb += bg;
              }
              p.io.rgb_out[in.ray * 3 + 0] = sr;
              p.io.rgb_out[in.ray * 3 + 1] = sgn;
              p.io.rgb_out[in.ray * 3 + 2] = sb;
              if (p.io.depth_out) p.io.depth_out[in.ray] = (sd - p.io.near_t) / (p.io.far_t - p.io.near_t);
              if (p.io.acc_out) p.io.acc_out[in.ray] = sa;
            }
            named_bar_sync(1 + s, 128);  // scratch is reused by the next tile
          }
        }
        t_last += NB2_CLK() - ce;
        continue;
      }

      // ---- 256-wide layers: half 0 is drained into registers while half 1 accumulates ---------------------------------
      const int c_h0 = W * g, c_h1 = 128 + W * g;   // this thread's columns of each half
      HeldHalf<SPLIT, W> held;
      wait_acc(full0, ph0, t_wacc0);
      const long long ch0 = NB2_CLK();
      if (epi == EPI_RELU) drain_hold<EPI_RELU, SPLIT, F16, W>(acc, c_h0, p.head, sigma, held);
      else if (epi == EPI_LINEAR) drain_hold<EPI_LINEAR, SPLIT, F16, W>(acc, c_h0, p.head, sigma, held);
      else if (epi == EPI_RELU_SIGMA) drain_hold<EPI_RELU_SIGMA, SPLIT, F16, W>(acc, c_h0, p.head, sigma, held);
      else drain_hold<EPI_SIGMA_OUT, SPLIT, F16, W>(acc, c_h0, p.head, sigma, held);

      t_h0 += NB2_CLK() - ch0;
      wait_acc(full1, ph1, t_wacc1);
      const long long cs0 = NB2_CLK();
      if (epi != EPI_SIGMA_OUT) {
        // all MMAs of this layer are complete: the activation tiles may be rewritten.  Half 0 first (K chunks 0-1).
#pragma unroll
        for (int b = 0; b < W / 32; ++b) store_block<SPLIT>(slot_base, lo_off, row, c_h0 + 32 * b, held.hi[b], held.lo[SPLIT ? b : 0]);
        publish(ready0);
        t_st0 += NB2_CLK() - cs0;
        // half 1 -> K chunks 2-3
#pragma unroll
        for (int b = 0; b < W / 32; b += (SPLIT ? 1 : 2)) {
          uint32_t hi0[16], lo0[16];
          if (SPLIT) {
            uint32_t m[32], c[32];
            tmem_ld32(acc + c_h1 + 32 * b, m);
            tmem_ld32(acc + 256 + c_h1 + 32 * b, c);
            tmem_ld_wait();
            if (epi == EPI_RELU) block_pack<EPI_RELU, SPLIT, F16>(m, c, c_h1 + 32 * b, p.head, sigma, hi0, lo0);
            else if (epi == EPI_LINEAR) block_pack<EPI_LINEAR, SPLIT, F16>(m, c, c_h1 + 32 * b, p.head, sigma, hi0, lo0);
            else block_pack<EPI_RELU_SIGMA, SPLIT, F16>(m, c, c_h1 + 32 * b, p.head, sigma, hi0, lo0);
            store_block<SPLIT>(slot_base, lo_off, row, c_h1 + 32 * b, hi0, lo0);
          } else {
            uint32_t m0[32], m1[32], hi1[16];
            tmem_ld32(acc + c_h1 + 32 * b, m0);
            tmem_ld32(acc + c_h1 + 32 * b + 32, m1);
            tmem_ld_wait();
            if (epi == EPI_RELU) {
              block_pack<EPI_RELU, SPLIT, F16>(m0, m0, c_h1 + 32 * b, p.head, sigma, hi0, lo0);
              block_pack<EPI_RELU, SPLIT, F16>(m1, m1, c_h1 + 32 * b + 32, p.head, sigma, hi1, lo0);
            } else if (epi == EPI_LINEAR) {
              block_pack<EPI_LINEAR, SPLIT, F16>(m0, m0, c_h1 + 32 * b, p.head, sigma, hi0, lo0);
              block_pack<EPI_LINEAR, SPLIT, F16>(m1, m1, c_h1 + 32 * b + 32, p.head, sigma, hi1, lo0);
            } else {
              block_pack<EPI_RELU_SIGMA, SPLIT, F16>(m0, m0, c_h1 + 32 * b, p.head, sigma, hi0, lo0);
              block_pack<EPI_RELU_SIGMA, SPLIT, F16>(m1, m1, c_h1 + 32 * b + 32, p.head, sigma, hi1, lo0);
            }
            store_block<SPLIT>(slot_base, lo_off, row, c_h1 + 32 * b, hi0, lo0);
            store_block<SPLIT>(slot_base, lo_off, row, c_h1 + 32 * b + 32, hi1, lo0);
          }
        }
        publish(ready1);
        if (epi == EPI_RELU_SIGMA && g == 0) sigma += __ldg(p.head + kHeadSigmaB);
      } else {
        // ---- proposal tail: out = w_s . relu(acc + b) + b_s; nothing is stored for a next layer ------------------------
#pragma unroll
        for (int b = 0; b < W / 32; ++b) {
          uint32_t m[32], c[32], hi0[16], lo0[16];
          tmem_ld32(acc + c_h1 + 32 * b, m);
          if (SPLIT) tmem_ld32(acc + 256 + c_h1 + 32 * b, c);
          tmem_ld_wait();
          block_pack<EPI_SIGMA_OUT, SPLIT, F16>(m, SPLIT ? c : m, c_h1 + 32 * b, p.head, sigma, hi0, lo0);
        }
        if (g == 0) sigma += __ldg(p.head + kHeadSigmaB);
        if (EW == 2) {
          const uint32_t xaddr = slot_base + (uint32_t)row * 4u;
          if (g == 1) asm volatile("st.shared.b32 [%0], %1;" ::"r"(xaddr), "r"(__float_as_uint(sigma)));
          named_bar_sync(3, 256);
          if (g == 0) {
            uint32_t x0;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x0) : "r"(xaddr));
            sigma += __uint_as_float(x0);
          }
        }
        if (has_next) begin_tile();
        if (in.valid && g == 0) p.io.out[grow] = sigma;
      }

      // ---- work that fits into the wait for the next layer's half 0 -----------------------------------------------------
      const long long cw = NB2_CLK();
      t_h1 += cw - cs0;
      if (has_next && l == 0) enc_phase(in_next, 0);
      if (has_next && l == 1 && EW == 1) enc_phase(in_next, 1);
      if (l == p.dir_layer && g == 0) {
        // the encoded position is dead after the skip layer: its tile now takes the encoded direction (columns 0-31)
        float rot[3] = {0.f, 0.f, 0.f};
        if (in.valid) normalize_dir(in.d, rot);
        write_enc_row<SPLIT, F16, kDirCols, kMaxDirLevels>(e_hi, e_lo, row, rot, p.dir_levels, in.valid);
        fence_proxy_async_smem();
      }
      t_enc += NB2_CLK() - cw;
    }
    in = in_next;
  }
  if (NB2_PROF_ON && threadIdx.x == kRolesThreads) {
    long long* o = p.prof + blockIdx.x * 16;
    o[6] = t_enc; o[7] = t_wacc0; o[8] = t_h0; o[9] = t_last; o[10] = NB2_CLK() - t0e; o[11] = n_iters; o[12] = net.n_layers;
    o[13] = t_wacc1; o[14] = t_st0; o[15] = t_h1;   // t_h1 includes t_st0 (everything after "half 1 complete")
  }
}

// ---- MMA issuer (leader CTA; one elected thread per issuing warp) ----------------------------------------------------------
// ISSUER is a template parameter so that every descriptor stays in uniform registers (a value derived from the warp
// index is not provably uniform and would cost three R2UR transfers per instruction).
// Single pass: issuer i owns resident tile i.  Split: issuer 0 owns the main accumulator (hi x Wh), issuer 1 the
// correction accumulator (lo x Wh, hi x Wl).  Both walk every ring stage and both commit every stage and every N-half
// (those barriers count two arrivals).
template <int NSLOTS, bool SPLIT, bool F16, int ISSUER>
__device__ __forceinline__ void mma_issuer3(const Tc3Params& q, Tc3Misc* misc, uint32_t act_base, uint32_t ring_base,
                                            uint32_t tmem_base, int64_t n_iters) {
  using LT = TcLayout<NSLOTS, SPLIT>;
  const TcParams& p = q.base;
  const TcNet& net = p.net;
  constexpr int issuer = ISSUER;
  uint32_t stage = 0, phase = 0, pr0 = 0, pr1 = 0;
  long long t_wa = 0, t_ww = 0, t0m = NB2_CLK();
  const uint32_t ring_lo = umma_desc_lo(ring_base);
  const uint32_t idesc = umma_idesc_16(256, 128, F16);
  constexpr uint32_t kLoPart = (uint32_t)(kChunksPerSlot * kTileBytes) >> 4;
  // single pass: this issuer's tile; split: the only tile
  const uint32_t act_lo = umma_desc_lo(act_base + (SPLIT ? 0u : (uint32_t)issuer * (uint32_t)LT::kSlotBytes));
  const uint32_t d_base = tmem_base + (SPLIT ? (uint32_t)(256 * issuer) : (uint32_t)(256 * issuer));
  auto wait_stage = [&]() {
    const long long c0 = NB2_CLK();
    mbar_wait(smem_u32(&misc->w_full[stage]), phase);
    mbar_wait(smem_u32(&misc->w_peer[stage]), phase);
    t_ww += NB2_CLK() - c0;
    tc_fence_after();
  };
  auto release_stage = [&]() {
    umma2_commit_mcast(smem_u32(&misc->w_empty[stage]), 3);
    if (++stage == kStages3) { stage = 0; phase ^= 1u; }
  };
  auto wait_ready = [&](int which, uint32_t& par) {
    const long long c0 = NB2_CLK();
    mbar_wait(smem_u32(&misc->a_ready[which]), par);
    par ^= 1u;
    t_wa += NB2_CLK() - c0;
    tc_fence_after();
  };
  for (int64_t it = 0; it < n_iters; ++it) {
    for (int l = 0; l < net.n_layers; ++l) {
      const TcLayer& L = net.layer[l];
      const Tc3Layer& P = q.plan[l];
      for (int h = 0; h < L.nc; ++h) {
        const uint32_t d = d_base + (uint32_t)(128 * h);
        for (int j = 0; j < P.n; ++j) {
          if (h == 0 && j == 0) wait_ready(0, pr0);
          if (h == 0 && j == P.n_early) wait_ready(1, pr1);
          const int ks0 = P.ks0[j];
          const uint32_t a_hi = act_lo + (uint32_t)L.a_src[P.k[j]] * (kTileBytes >> 4);
          wait_stage();
          const uint32_t w0 = ring_lo + stage * (kStageBytes3 >> 4);
          // single pass: hi x W;   split issuer 0: hi x Wh -> main;   split issuer 1: lo x Wh -> correction
          const uint32_t a_op = (SPLIT && issuer == 1) ? a_hi + kLoPart : a_hi;
          if (ks0 == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma2_bf16_ss(d, umma_desc_from_lo(a_op + 2 * ks), umma_desc_from_lo(w0 + 2 * ks), idesc, (uint32_t)((j | ks) != 0));
          } else {   // bias-only chunk: the one k-step that holds the constant-1 column
            umma2_bf16_ss(d, umma_desc_from_lo(a_op + 6), umma_desc_from_lo(w0 + 6), idesc, (uint32_t)(j != 0));
          }
          release_stage();
          if (SPLIT) {
            wait_stage();
            if (issuer == 1) {
              const uint32_t wl = ring_lo + stage * (kStageBytes3 >> 4);
              if (ks0 == 0) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma2_bf16_ss(d, umma_desc_from_lo(a_hi + 2 * ks), umma_desc_from_lo(wl + 2 * ks), idesc, 1u);
              } else {
                umma2_bf16_ss(d, umma_desc_from_lo(a_hi + 6), umma_desc_from_lo(wl + 6), idesc, 1u);
              }
            }
            release_stage();
          }
        }
        if (h == 0 && P.n_early >= P.n) wait_ready(1, pr1);   // (no late chunks: still consume the second barrier)
        umma2_commit_mcast(smem_u32(&misc->acc_full[h]), 3);
      }
    }
  }
  if (NB2_PROF_ON && issuer == 0) { p.prof[blockIdx.x * 16 + 3] = t_wa; p.prof[blockIdx.x * 16 + 4] = t_ww; p.prof[blockIdx.x * 16 + 5] = NB2_CLK() - t0m; }
}

template <int NSLOTS, bool SPLIT, bool F16>
__global__ void __launch_bounds__(kTcThreads, 1) mlp_tc3_kernel(const __grid_constant__ Tc3Params q) {
  using LT = TcLayout<NSLOTS, SPLIT>;
  static_assert(kStages3 * kStageBytes3 == LT::kRingBytes, "ring size");
  const TcParams& p = q.base;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t act_base = smem_base;
  const uint32_t ring_base = smem_base + LT::kActBytes;
  Tc3Misc* misc = reinterpret_cast<Tc3Misc*>(smem_al + LT::kActBytes + LT::kRingBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcNet& net = p.net;
  const int64_t tiles_per_iter = (int64_t)gridDim.x * NSLOTS;
  const int64_t n_iters = (p.n_tiles + tiles_per_iter - 1) / tiles_per_iter;
  const uint32_t rank = cluster_ctarank();
  constexpr int kParts = SPLIT ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages3; ++i) {
      mbar_init(smem_u32(&misc->w_full[i]), 1);
      mbar_init(smem_u32(&misc->w_peer[i]), 1);
      mbar_init(smem_u32(&misc->w_empty[i]), 2);    // both MMA issuers release every stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&misc->a_ready[i]), 16);   // every slot-group warp of both CTAs
      mbar_init(smem_u32(&misc->acc_full[i]), 2);   // both MMA issuers commit every N-half
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
    // =========================== weight streamer: this CTA's 64 rows of every 128(n) x 64(k) tile ======================
    reg_dealloc<kRoleRegs>();
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < n_iters; ++it) {
        for (int l = 0; l < net.n_layers; ++l) {
          const TcLayer& L = net.layer[l];
          const Tc3Layer& P = q.plan[l];
          for (int h = 0; h < L.nc; ++h) {
            for (int j = 0; j < P.n; ++j) {
              const size_t chunk = (size_t)(L.chunk0 + h * L.kc + P.k[j]);
#pragma unroll
              for (int part = 0; part < kParts; ++part) {
                mbar_wait(smem_u32(&misc->w_empty[stage]), phase ^ 1u);
                const uint32_t full = smem_u32(&misc->w_full[stage]);
                mbar_arrive_expect_tx(full, kStageBytes3);
                bulk_g2s(ring_base + stage * kStageBytes3,
                         p.wchunks + (chunk * 4 + (F16 ? 2 : 0) + part) * (kTileBytes / 2) + (size_t)rank * (kStageBytes3 / 2),
                         kStageBytes3, full);
                if (++stage == kStages3) { stage = 0; phase ^= 1u; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    reg_dealloc<kRoleRegs>();
    if (lane == 0 && rank != 0 && warp == 1) {
      // =========================== peer: relay "my half has landed" to the leader =====================
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < n_iters; ++it)
        for (int l = 0; l < net.n_layers; ++l) {
          const int n_entries = net.layer[l].nc * q.plan[l].n * kParts;
          for (int e = 0; e < n_entries; ++e) {
            mbar_wait(smem_u32(&misc->w_full[stage]), phase);
            mbar_arrive_remote_relaxed(smem_u32(&misc->w_peer[stage]), 0);
            if (++stage == kStages3) { stage = 0; phase ^= 1u; }
          }
        }
    } else if (lane == 0 && rank == 0) {
      // =========================== leader: two MMA issuers for the pair =================================
      if (warp == 1) mma_issuer3<NSLOTS, SPLIT, F16, 0>(q, misc, act_base, ring_base, tmem_base, n_iters);
      else mma_issuer3<NSLOTS, SPLIT, F16, 1>(q, misc, act_base, ring_base, tmem_base, n_iters);
    }
  } else if (warp >= 4) {
    reg_alloc<kGroupRegs>();
    if (NSLOTS == 2 || warp < 8) slot_group_run3<NSLOTS, SPLIT, F16, 0>(q, misc, act_base, tmem_base, n_iters, warp, lane, rank);
    else slot_group_run3<NSLOTS, SPLIT, F16, (NSLOTS == 2 ? 0 : 1)>(q, misc, act_base, tmem_base, n_iters, warp, lane, rank);
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

// ---- host side --------------------------------------------------------------------------------------------------------
template <int NSLOTS, bool SPLIT, bool F16>
static int launch_tc3_impl(nb2_handle* h, Tc3Params& prm, cudaStream_t st) {
  using LT = TcLayout<NSLOTS, SPLIT>;
  auto kern = mlp_tc3_kernel<NSLOTS, SPLIT, F16>;
  static bool attr_set = false;
  if (!attr_set) {
    NB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LT::kTotal));
    attr_set = true;
  }
  int64_t ctas = (prm.base.n_tiles + NSLOTS - 1) / NSLOTS;
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
  static int max_clusters = 0;
  if (!max_clusters) {
    cfg.gridDim = dim3(h->sm_count / 2 * 2);
    int n = 0;
    NB2_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    max_clusters = n > 0 ? n : 1;
  }
  cfg.gridDim = dim3((unsigned)std::min<int64_t>(ctas, (int64_t)max_clusters * 2));
  prm.base.cluster = 2;
  NB2_CUDA(cudaLaunchKernelEx(&cfg, kern, prm));
  h->launches++;
  return NB2_OK;
}

int launch_mlp_tc3(nb2_handle* h, const TcParams& base, int precision, cudaStream_t st) {
  Tc3Params prm;
  memset(&prm, 0, sizeof(prm));
  prm.base = base;
  const TcNet& net = base.net;
  for (int l = 0; l < net.n_layers; ++l) {
    const TcLayer& L = net.layer[l];
    Tc3Layer& P = prm.plan[l];
    int n = 0;
    for (int k = 0; k < L.kc; ++k)
      if (L.a_src[k] < 2) { P.k[n] = (unsigned char)k; P.ks0[n++] = (unsigned char)L.ks0[k]; }
    P.n_early = (unsigned char)n;
    for (int k = 0; k < L.kc; ++k)
      if (L.a_src[k] >= 2) { P.k[n] = (unsigned char)k; P.ks0[n++] = (unsigned char)L.ks0[k]; }
    P.n = (unsigned char)n;
    if (L.nc != 2 && L.epi != EPI_RGB) {
      set_error("mlp_forward: layer %d has an unsupported shape for the N-half kernel", l);
      return NB2_ERR_UNSUPPORTED;
    }
  }
  if (precision == NB2_PREC_BF16) return launch_tc3_impl<2, false, false>(h, prm, st);
  if (precision == NB2_PREC_FP16) return launch_tc3_impl<2, false, true>(h, prm, st);
  if (precision == NB2_PREC_BF16X3) return launch_tc3_impl<1, true, false>(h, prm, st);
  if (precision == NB2_PREC_FP16X3) return launch_tc3_impl<1, true, true>(h, prm, st);
  set_error("mlp_forward: unknown tensor-core precision %d", precision);
  return NB2_ERR_INVALID;
}

}  // namespace nb2
