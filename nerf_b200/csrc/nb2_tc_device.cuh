// nb2_tc_device.cuh — device-side pieces shared by the tcgen05 MLP kernels (nb2_mlp_tc.cu: single-pass CTA-pair kernel;
// nb2_mlp_tc4.cu: split-precision CTA-pair kernel with the activations in tensor memory): kernel parameters,
// shared-memory layout, A-operand row writers (encoding, 16-bit hi/lo split) and the hidden-layer epilogue blocks.
#pragma once
#include "nb2_common.cuh"
#include "nb2_rowio.cuh"
#include "nb2_tc_ptx.cuh"

namespace nb2 {
using namespace ptx;

constexpr int kStages = 4;
// two-slot kernels (384 threads, 168 registers at launch): warps 0-3 give registers back, the slot groups take them
constexpr int kRoleRegs = 72;    // 128 x 72 + 256 x 216 = 64512 <= 65536
constexpr int kGroupRegs = 216;
constexpr int kRolesThreads = 128;  // warps 0..3
// Every tensor kernel runs 128 role threads + two 128-thread warpgroups.  Single-pass modes: one warpgroup per
// resident tile (two slots).  Split modes (one slot): both warpgroups work on the same tile, each on half of the
// columns (a TMEM lane quadrant is reachable from any warp with the same warp % 4).
constexpr int kTcThreads = 384;
template <int NSLOTS> struct GroupsPerSlot { static constexpr int value = 2 / NSLOTS; };

struct TcParams {
  TcNet net;
  MlpIo io;
  const __nv_bfloat16* wchunks;
  const float* bias;
  const float* head;
  int pos_levels, dir_levels, has_dir;
  int debug;       // timing experiments only (results are garbage): 1 = no weight streaming / waits, 2 = no MMAs (handshake only)
  int dir_layer;   // layer after whose epilogue the direction encoding replaces the position encoding (-1: none)
  int cluster;   // CTAs per cluster sharing every weight tile through multicast bulk copies (1, 2 or 4)
  int64_t n_tiles;
  long long* prof;   // optional (debug): 16 cycle counters per CTA, see nb2_debug_tc_profile
};
// Role cycle counters are compiled in only with -DNB2_TC_PROFILE=1 (make PROFILE=1): they cost ~12 registers.
#ifndef NB2_TC_PROFILE
#define NB2_TC_PROFILE 0
#endif
#if NB2_TC_PROFILE
#define NB2_CLK() (p.prof ? clock64() : 0ll)
#define NB2_PROF_ON (p.prof != nullptr)
#else
#define NB2_CLK() 0ll
#define NB2_PROF_ON false
#endif

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
                                              int levels, bool valid, int g_begin = 0, int g_end = NCOLS / 8) {
  // only column groups [g_begin, g_end) (8 columns each) are produced by this thread; a frequency level is
  // evaluated when any of its six columns [3 + 6l, 9 + 6l) falls inside that range
  float v[NCOLS];
#pragma unroll
  for (int c = 0; c < NCOLS; ++c) v[c] = 0.f;
  if (valid) {
    v[0] = x[0]; v[1] = x[1]; v[2] = x[2];
#pragma unroll
    for (int l = 0; l < MAXLEV; ++l) {
      if (l < levels && 3 + 6 * l < 8 * g_end && 9 + 6 * l > 8 * g_begin) {
        const float sc = (float)(1 << l);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float s, c;
          enc_sincos(x[k] * sc, s, c);
          v[3 + 6 * l + k] = s;
          v[3 + 6 * l + 3 + k] = c;
        }
      }
    }
  }
  if (NCOLS == 64) v[NCOLS - 1] = 1.f;
#pragma unroll
  for (int g = 0; g < NCOLS / 8; ++g) {
    if (g >= g_begin && g < g_end) {
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = v[8 * g + i];
      store_a8<SPLIT, F16>(tile_hi, tile_lo, row, 8 * g, w);
    }
  }
}

// The same encoding, split in two steps so the arithmetic of the NEXT tile can run while the current tile's first
// MMAs execute: enc_compute fills packed 16-bit registers for column groups [G0, G0 + NG) of the 64-column position
// tile, enc_store writes them (16 bytes per group and row) once the tile buffer is free.
template <bool SPLIT, bool F16, int G0, int NG>
struct EncRegs {
  uint32_t hi[NG * 4];
  uint32_t lo[SPLIT ? NG * 4 : 1];
};
template <bool SPLIT, bool F16, int G0, int NG>
__device__ __forceinline__ void enc_compute(EncRegs<SPLIT, F16, G0, NG>& e, const float x[3], int levels, bool valid,
                                            const float* ext, const float* cov = nullptr) {
  float v[kEncCols];
#pragma unroll
  for (int c = 0; c < kEncCols; ++c) v[c] = 0.f;
  if (valid && ext != nullptr) {
    // externally encoded features (e.g. integrated positional encoding) take the place of the sin/cos columns
    v[0] = x[0]; v[1] = x[1]; v[2] = x[2];
#pragma unroll
    for (int c = 3; c < kEncCols - 1; ++c)
      if (c - 3 < 6 * levels && c >= 8 * G0 && c < 8 * (G0 + NG)) v[c] = __ldg(ext + c - 3);
  } else if (valid) {
    v[0] = x[0]; v[1] = x[1]; v[2] = x[2];
#pragma unroll
    for (int l = 0; l < kMaxPosLevels; ++l) {
      if (l < levels && 3 + 6 * l < 8 * (G0 + NG) && 9 + 6 * l > 8 * G0) {
        const float sc = (float)(1 << l);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float sn, cs;
          enc_sincos(x[k] * sc, sn, cs);
          if (cov != nullptr) {
            // integrated positional encoding (mip_methods.py:36-58): attenuation exp(-0.5 * 4^l * var_k) of the level
            const float damp = expf(-0.5f * (sc * sc * cov[k]));
            sn *= damp;
            cs *= damp;
          }
          v[3 + 6 * l + k] = sn;
          v[3 + 6 * l + 3 + k] = cs;
        }
      }
    }
  }
  v[kEncCols - 1] = 1.f;   // bias column
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = v[8 * (G0 + g) + 2 * i], b = v[8 * (G0 + g) + 2 * i + 1];
      e.hi[4 * g + i] = pack16x2<F16>(a, b);
      if (SPLIT) e.lo[4 * g + i] = residual16x2<F16>(a, b, e.hi[4 * g + i]);
    }
}
template <bool SPLIT, bool F16, int G0, int NG>
__device__ __forceinline__ void enc_store(const EncRegs<SPLIT, F16, G0, NG>& e, uint32_t tile_hi, uint32_t tile_lo, int row) {
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)(G0 + g)) ^ ((uint32_t)row & 7u)) << 4);
    st_shared_v4(tile_hi + off, e.hi[4 * g], e.hi[4 * g + 1], e.hi[4 * g + 2], e.hi[4 * g + 3]);
    if (SPLIT) st_shared_v4(tile_lo + off, e.lo[4 * g], e.lo[4 * g + 1], e.lo[4 * g + 2], e.lo[4 * g + 3]);
  }
}

// One 32-column block of the hidden epilogue: values (already summed with the correction accumulator in SPLIT mode)
// -> activation -> 16-bit A operand rows (+ running density-head dot product).
template <int EPI, bool SPLIT, bool F16>
__device__ __forceinline__ void epilogue_block(const uint32_t (&r)[32], int cb, uint32_t slot_base, uint32_t lo_off, int row,
                                               const float* __restrict__ head, float& sg) {
  constexpr bool kRelu = (EPI != EPI_LINEAR);
  constexpr bool kSigma = (EPI == EPI_RELU_SIGMA || EPI == EPI_SIGMA_OUT);
  constexpr bool kStore = (EPI != EPI_SIGMA_OUT);
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

// ---- hidden-layer epilogue: TMEM accumulator -> act -> 16-bit A operand of the next layer (in place) ----
// EPI: EPI_RELU / EPI_LINEAR / EPI_RELU_SIGMA / EPI_SIGMA_OUT.  Returns the density-head dot product
// (without its bias) for the *_SIGMA kinds.  In SPLIT mode the cross terms live in a second accumulator
// (columns +256) and are added here in fp32.  (A software-pipelined variant that kept the next block's
// tcgen05.ld in flight during the conversion measured 4-15 % SLOWER on B200 and was dropped.)
template <int EPI, bool SPLIT, bool F16>
__device__ __forceinline__ float epilogue_hidden(uint32_t acc, uint32_t slot_base, uint32_t lo_off, int row,
                                                 const float* __restrict__ head, int cb_begin, int cb_end) {
  float sg = 0.f;
  if (SPLIT) {
    // two blocks (main + correction each) per tcgen05.wait::ld round trip
#pragma unroll 1
    for (int cb = cb_begin; cb < cb_end; cb += 2) {
      uint32_t m0[32], c0[32], m1[32], c1[32];
      tmem_ld32(acc + cb * 32, m0);
      tmem_ld32(acc + 256 + cb * 32, c0);
      tmem_ld32(acc + (cb + 1) * 32, m1);
      tmem_ld32(acc + 256 + (cb + 1) * 32, c1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) m0[j] = __float_as_uint(__uint_as_float(m0[j]) + __uint_as_float(c0[j]));
      epilogue_block<EPI, SPLIT, F16>(m0, cb, slot_base, lo_off, row, head, sg);
#pragma unroll
      for (int j = 0; j < 32; ++j) m1[j] = __float_as_uint(__uint_as_float(m1[j]) + __uint_as_float(c1[j]));
      epilogue_block<EPI, SPLIT, F16>(m1, cb + 1, slot_base, lo_off, row, head, sg);
    }
  } else {
    // four blocks per tcgen05.wait::ld round trip: the load latency (~250 cycles) is paid twice per layer, not 8 times
#pragma unroll 1
    for (int cb = cb_begin; cb < cb_end; cb += 4) {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld32(acc + cb * 32, r0);
      tmem_ld32(acc + (cb + 1) * 32, r1);
      tmem_ld32(acc + (cb + 2) * 32, r2);
      tmem_ld32(acc + (cb + 3) * 32, r3);
      tmem_ld_wait();
      epilogue_block<EPI, SPLIT, F16>(r0, cb, slot_base, lo_off, row, head, sg);
      epilogue_block<EPI, SPLIT, F16>(r1, cb + 1, slot_base, lo_off, row, head, sg);
      epilogue_block<EPI, SPLIT, F16>(r2, cb + 2, slot_base, lo_off, row, head, sg);
      epilogue_block<EPI, SPLIT, F16>(r3, cb + 3, slot_base, lo_off, row, head, sg);
    }
  }
  return sg;
}


// split-precision CTA-pair kernel with the hidden activations in tensor memory (nb2_mlp_tc4.cu)
int launch_mlp_tc4(nb2_handle* h, const TcParams& base, int precision, cudaStream_t st);

}  // namespace nb2
