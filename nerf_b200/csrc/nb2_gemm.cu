// nb2_gemm.cu — the layer-wise engine: one generic tcgen05 GEMM behind nb2_gemm_bf16.
//
//   D[M, N] = epilogue( sum over segments s of  A_s[M, K_s] * B_s[N, K_s]^T ),   bf16 operands, fp32 accumulation in TMEM
//
// It is what the TRAINING step and the Ref-NeRF forward (8f-3) are built from: every nn.Linear forward,
// its dgrad and its wgrad are instances of it, with the matrices left in their natural row-major layouts:
//   forward   Y  = X  W^T         A = X  [rows][in]   K-major      B = W  [out][in]  K-major
//   dgrad     dX = dY W           A = dY [rows][out]  K-major      B = W  [out][in]  MN-major (its N = `in` is contiguous)
//   wgrad     dW = dY^T X         A = dY [rows][out]  MN-major     B = X  [rows][in] MN-major, split-K over CTAs
// "MN-major" is the tcgen05 name for an operand whose non-contracted dimension is contiguous; the tensor core transposes
// it for free through the matrix descriptor (a_major / b_major bits of the instruction descriptor), so no transposed
// copies of activations or weights exist anywhere.  Segments concatenate along K: the three passes of the split
// precision (lo x hi, hi x lo, hi x hi; small terms first) and the reference's torch.cat inputs
// (skip connection, bottleneck + encoded direction) are just more segments accumulating into the same tile.
//
// Kernel (persistent, one CTA per SM, cta_group::1, 128 x BN output tile, K chunks of 64):
//   warps 0-3, 6-9   epilogue, one group of four warps per accumulator buffer (two tiles drain at once while a third
//               accumulates... the MMA warp waits for a buffer's group): TMEM accumulator -> bias / act /
//               relu-mask -> bf16 hi (+ lo residual) and / or fp32 rows, transposed through a swizzled staging tile so
//               that every global store / mask load instruction covers whole 128-byte rows
//   warp 4      producer: one elected thread issues TMA tensor copies (cp.async.bulk.tensor.2d, SWIZZLE_128B tensor maps
//               built on the host per operand) of [rows][64 elements] boxes of the row-major global matrices into the
//               shared-memory layout both operand kinds share; out-of-bounds rows / columns are zero-filled by the TMA
//               unit; completion is counted in bytes on the stage's mbarrier.  (The first version used 128 threads of
//               16-byte cp.async copies and measured 15 B/cycle/SM; the TMA unit sustains ~60.)
//   warp 5      TMEM allocation + MMA issue (tcgen05.mma.kind::f16, M128 x BN x K16), tcgen05.commit frees stages
// The kernel is HBM-bound for the network's 256-wide layers (a 128 x 256 tile reads 64 KB and writes 64-128 KB for
// 2048 tensor cycles); its roofline is the measured copy bandwidth, DESIGN.md section 12.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "nb2_common.cuh"
#include "nb2_tc_ptx.cuh"

namespace nb2 {
using namespace ptx;

constexpr int kGStages = 4;
constexpr int kGABytes = 128 * 128;         // A operand stage: 128 rows x 64 bf16 (K-major) or 2 blocks of 64 x 64 (MN-major)
constexpr int kGBBytes = 256 * 128;         // B operand stage: up to 256 rows (K-major) or 4 blocks of 64 x 64 (MN-major)
constexpr int kGStageBytes = kGABytes + kGBBytes;
constexpr int kGThreads = 320;            // warps 0-3 / 6-9: epilogue groups 0 / 1; warp 4: TMA producer; warp 5: MMA issuer
constexpr int kGStagingBytes = 8 * 4096;   // epilogue: per warp one 32 rows x 128 B staging tile (mask, hi, lo / fp32 rows in turn)
constexpr int kGSmem = kGStages * kGStageBytes + kGStagingBytes + 1024 /* barriers */ + 1024 /* alignment slack */;
static_assert(kGSmem <= 232448, "exceeds 227 KB of shared memory");

struct GemmBars {
  uint64_t full[kGStages];
  uint64_t empty[kGStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

struct GemmKParams {
  CUtensorMap amap[NB2_GEMM_MAX_SEG];   // per segment: A as [rows][cols] row-major, box 64 columns x (128 | 64) rows
  CUtensorMap bmap[NB2_GEMM_MAX_SEG];   // per segment: B, box 64 columns x (bn | 64) rows
  nb2_gemm_desc d;
  int bn;                 // N tile (multiple of 32, <= 256)
  int m_tiles, n_tiles;
  int64_t k_per_split;    // rows of K per split (multiple of 64); splits == 1: unused
  int fused3;             // segments come in split-precision triples (A_lo B_hi, A_hi B_lo, A_hi B_hi): tiles shared between passes
  int tma_out;            // out_hi / out_lo leave through the TMA unit (omap), one 32 x 64 box per warp and pass
  int reuse3;             // forward / dgrad split-precision triple with K = 256: pass order (hi x lo), (lo x hi), (hi x hi); the
                          // last pass reuses the B tiles of the second one in place and reloads only the A slots
  int rowsum;             // d.a_rowsum_out: one extra N = 16 MMA per k-step of the A_lo and A_hi tiles against a tile of ones
  CUtensorMap omap[2];    // out_hi, out_lo as [M][N] bf16, box 64 columns x 32 rows
  int dbg;                // NB2_TC_DEBUG (timing ablations only): 16 no global stores, 32 no bias loads
};

// TMA: one [box rows][64 columns] box of a 2-D tensor map -> shared memory (128-byte swizzle), bytes counted on `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int col, int row, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(col), "r"(row), "r"(bar)
               : "memory");
}

// TMA store: one [32 rows][64 columns] box of the staging tile (128-byte swizzle) -> global; rows / columns beyond the
// tensor's extent are clipped by the TMA unit.  Bulk-group completion: wait_group.read = the shared-memory source may be reused.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col), "r"(row), "r"(src) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// D (128 x N, fp32, TMEM) (+)= A (smem) * B^T (smem), single CTA
__device__ __forceinline__ void umma1_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma1_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar)
      : "memory");
}
// shared-memory matrix descriptor, 128-byte swizzle.  K-major: LBO unused, SBO = 1024 (8 rows x 128 B).
// MN-major: atoms of 64 MN-elements x 8 k-rows; SBO = 1024 (next 8 k-rows), LBO = 8192 (next 64 MN-elements: one
// [64 k-rows][128 B] block per 64 MN-elements).  Field layout: cute::UMMA::SmemDescriptor.
__device__ __forceinline__ uint64_t gemm_smem_desc(uint32_t saddr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(mn_major ? (8192u >> 4) : 1u) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t gemm_idesc(int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

// EPI = epilogue warps: 8 (two groups of four, one per accumulator buffer; every shape) or 16 (four groups: two per buffer,
// each draining one half of the tile's columns in 32-column passes; the bf16-output shapes of the hot layers, whose epilogue
// is latency-bound per warp -- two epilogue warps per scheduler ran at IPC 0.33).
template <int EPI>
__global__ void __launch_bounds__(32 * (EPI + 2), 1) gemm_bf16_kernel(const __grid_constant__ GemmKParams p) {
  constexpr int kPW = EPI == 8 ? 4 : 16, kMW = EPI == 8 ? 5 : 17;     // producer / MMA warp (EPI == 8 keeps warps 0-3, 6-9 as epilogue)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the next kernel of the stream may be scheduled as SMs free up
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  GemmBars* bars = reinterpret_cast<GemmBars*>(smem_al + kGStages * kGStageBytes + kGStagingBytes);
  const uint32_t stg_base = smem_base + kGStages * kGStageBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const nb2_gemm_desc& d = p.d;
  const int splits = d.splits > 1 ? d.splits : 1;
  const int64_t n_items = (int64_t)p.m_tiles * p.n_tiles * splits;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kGStages; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);      // the producer's arrive.expect_tx; the TMA unit completes the bytes
      mbar_init(smem_u32(&bars->empty[i]), 1);     // tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->acc_full[i]), 1);
      mbar_init(smem_u32(&bars->acc_empty[i]), EPI / 2); // one arrival per epilogue warp of the buffer
    }
    mbar_fence_init();
  }
  if (warp == kMW) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  if (p.rowsum) {
    // bias gradient riding on the weight gradient: an 8 KB MN-major B block [64 k-rows][64 columns] of bf16 ones (constant, so
    // the 128-byte swizzle is irrelevant) in the first two staging tiles -- the host guarantees ONE work item per CTA in
    // this mode, and its epilogue (the staging tiles' only other user) starts after the last MMA has completed
    for (int i = threadIdx.x; i < 512; i += 32 * (EPI + 2)) st_shared_v4(stg_base + 16u * i, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // Programmatic dependent launch: consecutive GEMMs of a launch plan are launched with the stream-serialization attribute,
  // so this CTA may have started (barriers, tensor memory, the prologue above) while the previous kernel of the stream was
  // still draining its last tiles on other SMs; nothing before this point touches global memory.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // the K range of one work item inside segment s (split-K shares the range across the segments)
  auto k_range = [&](int s, int split, int64_t& k0, int64_t& k1) {
    k0 = 0;
    k1 = d.seg[s].K;
    if (splits > 1) {
      k0 = (int64_t)split * p.k_per_split;
      k1 = k0 + p.k_per_split < d.seg[s].K ? k0 + p.k_per_split : d.seg[s].K;
      if (k0 > k1) k0 = k1;
    }
  };

  if (warp == kPW) {
    // =========================================== producer (one thread drives the TMA unit) ===========================
    if (lane == 0) {
      uint32_t issued = 0;
      for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int m_tile = (int)(it % p.m_tiles);
        const int n_tile = (int)((it / p.m_tiles) % p.n_tiles);
        const int split = (int)(it / ((int64_t)p.m_tiles * p.n_tiles));
        const int m0 = m_tile * 128, n0 = n_tile * p.bn;
        // one ring stage: the A tile of segment sa and / or the B tile of segment sb (-1: that half of the stage stays unused)
        auto load_stage = [&](int sa, int sb, int64_t k) {
          const int sm = sa >= 0 ? sa : sb;
          const bool a_mn = d.seg[sm].a.mn_major != 0, b_mn = d.seg[sm].b.mn_major != 0;
          const int b_boxes = b_mn ? (p.bn + 63) / 64 : 1;
          const uint32_t bytes = (sa >= 0 ? (uint32_t)kGABytes : 0u) + (sb >= 0 ? (uint32_t)(b_mn ? b_boxes * 8192 : p.bn * 128) : 0u);
          const uint32_t stage = issued % kGStages;
          if (issued >= kGStages) mbar_wait(smem_u32(&bars->empty[stage]), ((issued / kGStages) - 1) & 1u);
          const uint32_t a_dst = smem_base + stage * kGStageBytes, b_dst = a_dst + kGABytes;
          const uint32_t full = smem_u32(&bars->full[stage]);
          mbar_arrive_expect_tx(full, bytes);
          if (sa >= 0) {
            if (!a_mn) {
              tma_load_2d(a_dst, &p.amap[sa], (int)k, m0, full);                      // 128 rows (M) x 64 columns (K)
            } else {
              tma_load_2d(a_dst, &p.amap[sa], m0, (int)k, full);                      // 64 rows (K) x 64 columns (M), twice
              tma_load_2d(a_dst + 8192, &p.amap[sa], m0 + 64, (int)k, full);
            }
          }
          if (sb >= 0) {
            if (!b_mn) {
              tma_load_2d(b_dst, &p.bmap[sb], (int)k, n0, full);                      // bn rows (N) x 64 columns (K)
            } else {
              for (int j = 0; j < b_boxes; ++j) tma_load_2d(b_dst + j * 8192, &p.bmap[sb], n0 + 64 * j, (int)k, full);
            }
          }
          ++issued;
        };
        if (p.fused3) {
          // split-precision triples: per K chunk the four distinct tiles are loaded ONCE -- (A_lo, B_hi), then A_hi, then B_lo
          // (96 KB instead of 144 KB for the three passes; the L2 -> shared stream is what bounds this kernel)
          for (int t = 0; t < d.n_seg; t += 3) {
            int64_t k0, k1;
            k_range(t, split, k0, k1);
            for (int64_t k = k0; k < k1; k += 64) {
              load_stage(t, t, k);
              load_stage(t + 2, -1, k);
              load_stage(-1, t + 1, k);
            }
          }
        } else if (p.reuse3) {
          // both cross terms first, but (A_hi, B_lo) before (A_lo, B_hi): the four (A_lo, B_hi) stages
          // are then exactly the ring's four stages, and the closing A_hi x B_hi pass finds its B tiles already in place --
          // it reloads only the 16 KB A slots (448 KB instead of 576 KB of tiles per 128 x 256 output tile)
          for (int64_t k = 0; k < d.seg[1].K; k += 64) load_stage(1, 1, k);
          for (int64_t k = 0; k < d.seg[0].K; k += 64) load_stage(0, 0, k);
          for (int64_t k = 0; k < d.seg[2].K; k += 64) load_stage(2, -1, k);
        } else {
          for (int s = 0; s < d.n_seg; ++s) {
            int64_t k0, k1;
            k_range(s, split, k0, k1);
            for (int64_t k = k0; k < k1; k += 64) load_stage(s, s, k);
          }
        }
      }
    }
  } else if (warp == kMW) {
    // =========================================== MMA issuer (whole warp, one elected lane per instruction) ===========
    uint32_t consumed = 0, item = 0;
    for (int64_t it = blockIdx.x; it < n_items; it += gridDim.x, ++item) {
      const int split = (int)(it / ((int64_t)p.m_tiles * p.n_tiles));
      const uint32_t buf = item & 1u;
      if (item >= 2) mbar_wait(smem_u32(&bars->acc_empty[buf]), ((item >> 1) - 1) & 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * 256;
      bool first = true;
      auto wait_full = [&]() {
        const uint32_t stage = consumed % kGStages;
        mbar_wait(smem_u32(&bars->full[stage]), (consumed / kGStages) & 1u);
        tc_fence_after();
        ++consumed;
        return stage;
      };
      // the k-steps of one K chunk: A tile of ring stage sa against the B tile of ring stage sb
      auto mma_chunk = [&](uint32_t sa, uint32_t sb, int ksteps, uint32_t idesc, bool a_mn, bool b_mn) {
        const uint32_t a_s = smem_base + sa * kGStageBytes, b_s = smem_base + sb * kGStageBytes + kGABytes;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t ad = gemm_smem_desc(a_s + (a_mn ? ks * 2048 : ks * 32), a_mn);
          const uint64_t bd = gemm_smem_desc(b_s + (b_mn ? ks * 2048 : ks * 32), b_mn);
          umma1_elect(tmem_d, ad, bd, idesc, first ? 0u : 1u);
          first = false;
        }
      };
      // row sums of A (the bias gradient): D[128, 16] at TMEM columns [256, 272) of the single item of this CTA; only the
      // N tile 0 items compute them (the other N tiles of the same M tile would repeat the same sums)
      const bool want_rowsum = p.rowsum && ((it / p.m_tiles) % p.n_tiles) == 0;
      bool first_ones = true;
      auto ones_chunk = [&](uint32_t sa, int ksteps) {
        const uint32_t a_s = smem_base + sa * kGStageBytes;
        const uint32_t idesc1 = gemm_idesc(16, true, true);
        for (int ks = 0; ks < ksteps; ++ks) {
          umma1_elect(tmem_base + 256, gemm_smem_desc(a_s + ks * 2048, true), gemm_smem_desc(stg_base + ks * 2048, true), idesc1, first_ones ? 0u : 1u);
          first_ones = false;
        }
      };
      const int seg_step = p.fused3 ? 3 : 1;
      for (int si = 0; si < d.n_seg; si += seg_step) {
        const int s = p.reuse3 ? (si == 0 ? 1 : (si == 1 ? 0 : 2)) : si;     // reuse3: (hi x lo), (lo x hi), (hi x hi)
        const bool a_mn = d.seg[s].a.mn_major != 0, b_mn = d.seg[s].b.mn_major != 0;
        const uint32_t idesc = gemm_idesc(p.bn, a_mn, b_mn);
        int64_t k0, k1;
        k_range(s, split, k0, k1);
        for (int64_t k = k0; k < k1; k += 64) {
          const int ksteps = (int)((k1 - k + 15) / 16 < 4 ? (k1 - k + 15) / 16 : 4);
          if (!p.fused3) {
            const uint32_t st = wait_full();
            mma_chunk(st, st, ksteps, idesc, a_mn, b_mn);
            umma1_commit_elect(smem_u32(&bars->empty[st]));
          } else {
            const uint32_t st0 = wait_full();                 // (A_lo, B_hi)
            mma_chunk(st0, st0, ksteps, idesc, a_mn, b_mn);   // lo x hi
            if (want_rowsum) ones_chunk(st0, ksteps);         // lo x 1
            const uint32_t st1 = wait_full();                 // A_hi
            mma_chunk(st1, st0, ksteps, idesc, a_mn, b_mn);   // hi x hi
            if (want_rowsum) ones_chunk(st1, ksteps);         // hi x 1
            umma1_commit_elect(smem_u32(&bars->empty[st0]));
            const uint32_t st2 = wait_full();                 // B_lo
            mma_chunk(st1, st2, ksteps, idesc, a_mn, b_mn);   // hi x lo
            umma1_commit_elect(smem_u32(&bars->empty[st1]));
            umma1_commit_elect(smem_u32(&bars->empty[st2]));
          }
        }
      }
      umma1_commit_elect(smem_u32(&bars->acc_full[buf]));
    }
  } else if constexpr (EPI == 16) {
    // =========================================== epilogue, 16 warps ===================================================
    // Shapes the host selects it for: bf16 hi (+ lo) outputs through the TMA unit, optional aligned bias, activation, optional
    // bf16 relu mask; no fp32 output, no split-K.  Warp w: TMEM lane quadrant w & 3, group w >> 2 = (column half) * 2 + buffer.
    // A pass is 32 columns: TMEM -> bias / act -> bf16 hi + lo words (mask ANDed in) -> a 32 x 32 staging tile (64-byte
    // rows, 64-byte swizzle) -> one bulk tensor store; hi and lo take the tile in turn.
    const int group = warp >> 2, quad = warp & 3;
    const uint32_t buf = (uint32_t)(group & 1), half = (uint32_t)(group >> 1);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256;
    const uint32_t stg = stg_base + (uint32_t)warp * 2048u;
    auto sw64 = [](int r, int ch) { return (uint32_t)r * 64u + (uint32_t)((ch ^ ((r >> 1) & 3)) << 4); };
    const __nv_bfloat16* mask = reinterpret_cast<const __nv_bfloat16*>(d.mask);
    const int half_cols = p.bn >> 1;
    uint32_t item = buf;
    bool stg_busy = false;
    auto stg_acquire = [&]() {
      if (stg_busy) {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        stg_busy = false;
      }
    };
    for (int64_t it = blockIdx.x + (int64_t)buf * gridDim.x; it < n_items; it += 2 * (int64_t)gridDim.x, item += 2) {
      const int m_tile = (int)(it % p.m_tiles);
      const int n_tile = (int)((it / p.m_tiles) % p.n_tiles);
      const int64_t row0 = (int64_t)m_tile * 128 + quad * 32;
      const int64_t n0 = (int64_t)n_tile * p.bn + half * half_cols;
      mbar_wait(smem_u32(&bars->acc_full[buf]), (item >> 1) & 1u);
      __syncwarp();
      tc_fence_after();
      for (int cb = 0; cb < half_cols / 32; ++cb) {
        const int64_t c0 = n0 + cb * 32;
        if (c0 >= d.N) break;                                   // (N is a multiple of 32 here: whole passes only)
        if (mask != nullptr) {
          stg_acquire();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int q = lane + 32 * i, r = q >> 2, ch = q & 3;
            uint4 m = make_uint4(0u, 0u, 0u, 0u);
            if (row0 + r < d.M) m = __ldg(reinterpret_cast<const uint4*>(mask + (row0 + r) * d.ld_mask + c0 + 8 * ch));
            st_shared_v4(stg + sw64(r, ch), m.x, m.y, m.z, m.w);
          }
          __syncwarp();
        }
        uint32_t ra[32];
        tmem_ld32(lane_addr + half * half_cols + cb * 32, ra);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]);
        if (d.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(d.bias + c0);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 bq = __ldg(b4 + k);
            v[4 * k] += bq.x; v[4 * k + 1] += bq.y; v[4 * k + 2] += bq.z; v[4 * k + 3] += bq.w;
          }
        }
        if (d.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (d.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          hi[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          lo[j] = residual16x2<false>(v[2 * j], v[2 * j + 1], hi[j]);
        }
        if (mask != nullptr) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t m[4];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(m[0]), "=r"(m[1]), "=r"(m[2]), "=r"(m[3]) : "r"(stg + sw64(lane, k)));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t sel = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&m[e]), __float2bfloat162_rn(0.f));
              hi[4 * k + e] &= sel;
              lo[4 * k + e] &= sel;
            }
          }
        }
        for (int part = 0; part < (d.out_lo != nullptr ? 2 : 1); ++part) {
          stg_acquire();                                        // (own-row mask reads above are done: the tile may be overwritten)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (part == 0) st_shared_v4(stg + sw64(lane, k), hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
            else st_shared_v4(stg + sw64(lane, k), lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tma_store_2d(&p.omap[part], stg, (int)c0, (int)row0);
          stg_busy = true;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[buf]));
    }
    if (lane == 0) tma_store_wait_all();
  } else {
    // =========================================== epilogue ============================================================
    // two groups of four warps: group g drains accumulator buffer g, i.e. this CTA's items g, g + 2, ... -- two tiles are
    // in the epilogue at once and each scheduler has two epilogue warps to hide the TMEM / L1 round trips of the other
    const int ew = warp < 4 ? warp : warp - 2;          // 0..7
    const int group = ew >> 2;
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may read
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t item = (uint32_t)group;
    bool stg_busy = false;                              // a TMA store may still be reading this warp's staging tile
    auto stg_acquire = [&]() {
      if (stg_busy) {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        stg_busy = false;
      }
    };
    for (int64_t it = blockIdx.x + (int64_t)group * gridDim.x; it < n_items; it += 2 * (int64_t)gridDim.x, item += 2) {
      const int m_tile = (int)(it % p.m_tiles);
      const int n_tile = (int)((it / p.m_tiles) % p.n_tiles);
      const int split = (int)(it / ((int64_t)p.m_tiles * p.n_tiles));
      const uint32_t buf = (uint32_t)group;
      const int64_t row = (int64_t)m_tile * 128 + quad * 32 + lane;
      const int64_t n0 = (int64_t)n_tile * p.bn;
      // a work item whose K range is empty issued no MMA: its accumulator is undefined -> contributes zeros
      bool empty_k = true;
      for (int s = 0; s < d.n_seg; ++s) {
        int64_t k0, k1;
        k_range(s, split, k0, k1);
        if (k1 > k0) empty_k = false;
      }
      mbar_wait(smem_u32(&bars->acc_full[buf]), (item >> 1) & 1u);
      __syncwarp();
      tc_fence_after();
      float* out32 = d.out_f32 ? d.out_f32 + (int64_t)split * d.split_stride : nullptr;
      // Global accesses go through a warp-private shared-memory staging tile (32 rows x 128 B, 128-byte swizzle) so that
      // consecutive lanes touch consecutive 16-byte units of a row: one thread owns one accumulator ROW (TMEM lane), and
      // row-per-lane stores / mask loads cost one 128-byte line request per lane per instruction (measured: the first
      // version of this kernel spent 30 us per 128-row tile in them).  Unaligned / odd shapes take the scalar path.
      const bool vec16 = d.out_hi && (d.ld_16 & 7) == 0 && (d.N & 7) == 0 && ((reinterpret_cast<uintptr_t>(d.out_hi) & 15) == 0) &&
                         (!d.out_lo || (reinterpret_cast<uintptr_t>(d.out_lo) & 15) == 0);
      const bool vecm = d.mask && (d.ld_mask & 7) == 0 && (d.N & 7) == 0 && ((reinterpret_cast<uintptr_t>(d.mask) & 15) == 0);
      const bool vec32 = out32 && !d.out_hi && (d.ld_f32 & 3) == 0 && (d.N & 3) == 0 && ((reinterpret_cast<uintptr_t>(out32) & 15) == 0);
      const int64_t row0 = (int64_t)m_tile * 128 + quad * 32;
      const uint32_t stg = stg_base + (uint32_t)ew * 4096u;
      auto sw = [](int r, int ch) { return (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4); };
      const __nv_bfloat16* mask = reinterpret_cast<const __nv_bfloat16*>(d.mask);
      __nv_bfloat16* out_hi = reinterpret_cast<__nv_bfloat16*>(d.out_hi);
      __nv_bfloat16* out_lo = reinterpret_cast<__nv_bfloat16*>(d.out_lo);
      for (int cb = 0; cb < p.bn / 32; cb += 2) {
        const int nb = (cb + 1 < p.bn / 32) ? 2 : 1;          // 32-column blocks in this pass (64 columns = one 128-byte bf16 row)
        const int sh = nb == 2 ? 3 : 2;                       // 16-byte units per bf16 row of the pass: 8 or 4
        const int64_t c0 = n0 + cb * 32;
        uint32_t keep[2] = {0u, 0u};                          // relu mask of this thread's row: bit j of keep[hb] = column 32 hb + j passes
        const bool mask_packed = vecm && vec16 && out32 == nullptr;   // the dgrad shape: mask applied to the packed hi / lo words
        if (vecm) {
          stg_acquire();
          for (int q = lane; q < (32 << sh); q += 32) {
            const int r = q >> sh, ch = q & ((1 << sh) - 1);
            const int64_t grow = row0 + r, gcol = c0 + 8 * ch;
            uint4 m = make_uint4(0u, 0u, 0u, 0u);
            if (grow < d.M && gcol < d.N) m = __ldg(reinterpret_cast<const uint4*>(mask + grow * d.ld_mask + gcol));
            st_shared_v4(stg + sw(r, ch), m.x, m.y, m.z, m.w);
          }
          __syncwarp();
          // dgrad shape (mask_packed): every lane re-reads its OWN row of the staged mask right before it packs that half
          // (below); the other shapes expand the mask into one bit per column here
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (!mask_packed && k < (1 << sh)) {
              uint32_t m[4];
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(m[0]), "=r"(m[1]), "=r"(m[2]), "=r"(m[3]) : "r"(stg + sw(lane, k)));
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                keep[k >> 2] |= (bf16_lo_to_f32(m[e]) > 0.f ? 1u : 0u) << (8 * (k & 3) + 2 * e);
                keep[k >> 2] |= (bf16_hi_to_f32(m[e]) > 0.f ? 1u : 0u) << (8 * (k & 3) + 2 * e + 1);
              }
            }
          }
          __syncwarp();
        }
        uint32_t ra[32], rb[32];
        uint32_t lo_keep[2][16];                              // lo residuals wait in registers while the hi rows use the staging tile
        // (an ablation branch that filled ra / rb with constants used to sit here: the compiler hoisted its 64 register
        //  fills above the branch, i.e. into every pass -- ncu source page, 8 % of the epilogue's instructions)
        tmem_ld32(lane_addr + buf * 256 + cb * 32, ra);
        if (nb == 2) tmem_ld32(lane_addr + buf * 256 + cb * 32 + 32, rb);
        tmem_ld_wait();
        if (empty_k) {                                          // warp-uniform and rare: a branch, not 64 selects per pass
#pragma unroll
          for (int j = 0; j < 32; ++j) { ra[j] = 0u; rb[j] = 0u; }
        }
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          if (hb >= nb) break;
          const int64_t cc = c0 + 32 * hb;
          const int nv = (int)(d.N - cc < 32 ? (d.N - cc > 0 ? d.N - cc : 0) : 32);
          // (straight-line code: per-element branches around the bias loads serialised 64 dependent L1 round trips per
          //  pass and made the epilogue 10x longer than the main loop)
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(hb ? rb[j] : ra[j]);
          if (d.bias != nullptr && !(p.dbg & 32)) {
            if (nv == 32 && (reinterpret_cast<uintptr_t>(d.bias) & 15) == 0) {
              const float4* b4 = reinterpret_cast<const float4*>(d.bias + cc);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float4 bq = __ldg(b4 + k);
                v[4 * k] += bq.x; v[4 * k + 1] += bq.y; v[4 * k + 2] += bq.z; v[4 * k + 3] += bq.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __ldg(d.bias + (j < nv ? cc + j : 0)) * (j < nv ? 1.f : 0.f);
            }
          }
          if (d.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (d.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
          }
          if (mask_packed) {
            // applied to the packed words below
          } else if (vecm) {
            const uint32_t kb = keep[hb];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((kb >> j) & 1u) ? v[j] : 0.f;
          } else if (mask != nullptr && row < d.M) {
            const __nv_bfloat16* mrow = mask + row * d.ld_mask + cc;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nv && !(__bfloat162float(mrow[j]) > 0.f)) v[j] = 0.f;
          }
          if (vec32) {
            stg_acquire();
#pragma unroll
            for (int k = 0; k < 8; ++k)
              st_shared_v4(stg + sw(lane, k), __float_as_uint(v[4 * k]), __float_as_uint(v[4 * k + 1]), __float_as_uint(v[4 * k + 2]),
                           __float_as_uint(v[4 * k + 3]));
            __syncwarp();
            for (int q = lane; q < 256; q += 32) {
              const int r = q >> 3, ch = q & 7;
              const int64_t grow = row0 + r, gcol = cc + 4 * ch;
              if (grow < d.M && gcol < d.N) {
                uint4 x;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(stg + sw(r, ch)));
                *reinterpret_cast<uint4*>(out32 + grow * d.ld_f32 + gcol) = x;
              }
            }
            __syncwarp();
          } else if (out32 != nullptr && row < d.M) {
            float* o = out32 + row * d.ld_f32 + cc;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nv) o[j] = v[j];
          }
          if (out_hi != nullptr) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              hi[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
              lo[j] = residual16x2<false>(v[2 * j], v[2 * j + 1], hi[j]);   // v - float(hi): one FHFMA.BF16 per element, exact
            }
            if (mask_packed) {
              // one packed compare per column pair (HSET2.BF16 with a mask result) instead of ~6 instructions per element;
              // chunks 4 hb .. 4 hb + 3 of this lane's row still hold the mask: the hi rows of this half are stored after
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint32_t m[4];
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(m[0]), "=r"(m[1]), "=r"(m[2]), "=r"(m[3]) : "r"(stg + sw(lane, 4 * hb + k)));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t sel = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&m[e]), __float2bfloat162_rn(0.f));
                  hi[4 * k + e] &= sel;
                  lo[4 * k + e] &= sel;
                }
              }
            }
            if (vec16) {
              stg_acquire();
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                st_shared_v4(stg + sw(lane, 4 * hb + k), hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
                lo_keep[hb][4 * k] = lo[4 * k]; lo_keep[hb][4 * k + 1] = lo[4 * k + 1]; lo_keep[hb][4 * k + 2] = lo[4 * k + 2]; lo_keep[hb][4 * k + 3] = lo[4 * k + 3];
              }
            } else if (row < d.M) {
              __nv_bfloat16* oh = out_hi + row * d.ld_16 + cc;
              __nv_bfloat16* ol = out_lo ? out_lo + row * d.ld_16 + cc : nullptr;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) {
                  const uint32_t hw = (j & 1) ? (hi[j >> 1] >> 16) : (hi[j >> 1] & 0xFFFFu);
                  reinterpret_cast<unsigned short*>(oh)[j] = (unsigned short)hw;
                  if (ol != nullptr) {
                    const uint32_t lw = (j & 1) ? (lo[j >> 1] >> 16) : (lo[j >> 1] & 0xFFFFu);
                    reinterpret_cast<unsigned short*>(ol)[j] = (unsigned short)lw;
                  }
                }
            }
          }
        }
        if (vec16 && p.tma_out && nb == 2) {
          // the staged 32 x 64 tile leaves as ONE bulk tensor store per part (the copy loops below were ~35 % of the
          // epilogue's instructions); the lo part reuses the tile as soon as the TMA unit has read the hi part
          for (int part = 0; part < (out_lo != nullptr ? 2 : 1); ++part) {
            if (part) {
              stg_acquire();
#pragma unroll
              for (int hb = 0; hb < 2; ++hb) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  st_shared_v4(stg + sw(lane, 4 * hb + k), lo_keep[hb][4 * k], lo_keep[hb][4 * k + 1], lo_keep[hb][4 * k + 2], lo_keep[hb][4 * k + 3]);
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(p.dbg & 16)) tma_store_2d(&p.omap[part], stg, (int)c0, (int)row0);
            stg_busy = true;
          }
        } else if (vec16) {
          stg_acquire();
          for (int part = 0; part < (out_lo != nullptr ? 2 : 1); ++part) {
            __nv_bfloat16* dst = part ? out_lo : out_hi;
            if (part) {
#pragma unroll
              for (int hb = 0; hb < 2; ++hb)
                if (hb < nb) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    st_shared_v4(stg + sw(lane, 4 * hb + k), lo_keep[hb][4 * k], lo_keep[hb][4 * k + 1], lo_keep[hb][4 * k + 2], lo_keep[hb][4 * k + 3]);
                }
            }
            __syncwarp();
            for (int q = lane; q < (32 << sh); q += 32) {
              const int r = q >> sh, ch = q & ((1 << sh) - 1);
              const int64_t grow = row0 + r, gcol = c0 + 8 * ch;
              if (grow < d.M && gcol < d.N) {
                uint4 x;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(stg + sw(r, ch)));
                if (!(p.dbg & 16)) *reinterpret_cast<uint4*>(dst + grow * d.ld_16 + gcol) = x;
              }
            }
            __syncwarp();
          }
        }
      }
      if (p.rowsum && n_tile == 0) {
        uint32_t rs[32];
        tmem_ld32(lane_addr + 256, rs);                   // 16 identical columns (+ 16 unused): column 0 is the sum
        tmem_ld_wait();
        if (row < d.M) d.a_rowsum_out[(int64_t)split * d.a_rowsum_stride + row] = empty_k ? 0.f : __uint_as_float(rs[0]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[buf]));
    }
    if (lane == 0) tma_store_wait_all();                // the staging tile must outlive its last store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- small companions of the layer-wise engine ------------------------------------------------------------------------
// fp32 (rows x cols, row stride ld_src) -> bf16 hi (+ lo residual) with row stride ld_dst >= cols (multiple of 8); the
// pad columns [cols, ld_dst) are zeroed.  col_perm (optional, cols entries): destination column of source column c --
// how a weight matrix that multiplies a torch.cat input is stored in the column order of the engine's buffers.
__global__ void to_bf16_kernel(const float* __restrict__ src, int64_t rows, int cols, int64_t ld_src, const int* __restrict__ col_perm,
                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ld_dst) {
  const int64_t n = rows * ld_dst;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_dst;
    const int c = (int)(i - r * ld_dst);
    if (c >= cols) {
      if (col_perm == nullptr) {     // with a permutation the pad columns are whatever the permutation leaves unassigned
        hi[i] = __float2bfloat16_rn(0.f);
        if (lo) lo[i] = __float2bfloat16_rn(0.f);
      }
      continue;
    }
    const float v = src[r * ld_src + c];
    const int64_t o = col_perm ? r * ld_dst + col_perm[c] : i;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[o] = h;
    if (lo) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// every conversion of a launch in one grid: blockIdx.y selects the descriptor (a training step converts ~16-35 weight matrices
// after each optimizer step; one launch each cost more than the conversions themselves)
constexpr int kToBf16Batch = 32;
struct ToBf16Batch {
  nb2_to_bf16_desc d[kToBf16Batch];
};
__global__ void to_bf16_batch_kernel(const __grid_constant__ ToBf16Batch b) {
  const nb2_to_bf16_desc& d = b.d[blockIdx.y];
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(d.hi);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(d.lo);
  const int64_t n = d.rows * d.ld_dst;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d.ld_dst;
    const int c = (int)(i - r * d.ld_dst);
    if (c >= d.cols) {
      if (d.col_perm == nullptr) {
        hi[i] = __float2bfloat16_rn(0.f);
        if (lo) lo[i] = __float2bfloat16_rn(0.f);
      }
      continue;
    }
    const float v = d.src[r * d.ld_src + c];
    const int64_t o = d.col_perm ? r * d.ld_dst + d.col_perm[c] : i;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[o] = h;
    if (lo) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// out[m][c] (fp32, row stride ld_out, c < cols) = sum over splits of ws[s][m][perm(c)] (row stride ld_ws): the
// deterministic second stage of the split-K wgrad (no atomics: gradients are bit-reproducible run to run).
// Eight independent loads in flight per thread (one dependent load per split made the kernel latency-bound: 36 us for a
// 256 x 256 gradient over 74 splits); the additions stay in split order, so the result does not depend on the batching.
__device__ __forceinline__ void reduce_splits_body(const float* __restrict__ ws, int splits, int64_t split_stride, int rows, int cols, int ld_ws,
                                                   const int* __restrict__ col_perm, float* __restrict__ out, int ld_out, int accumulate,
                                                   int64_t first, int64_t stride) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = first; i < n; i += stride) {
    const int m = (int)(i / cols), c = (int)(i - (int64_t)m * cols);
    const int wc = col_perm ? col_perm[c] : c;
    const float* src = ws + (int64_t)m * ld_ws + wc;
    float acc = 0.f;
    int s = 0;
    for (; s + 8 <= splits; s += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (int64_t)(s + j) * split_stride);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j];
    }
    for (; s < splits; ++s) acc += __ldg(src + (int64_t)s * split_stride);
    float* o = out + (int64_t)m * ld_out + c;
    *o = accumulate ? *o + acc : acc;
  }
}

__global__ void reduce_splits_kernel(const float* __restrict__ ws, int splits, int64_t split_stride, int rows, int cols, int ld_ws,
                                     const int* __restrict__ col_perm, float* __restrict__ out, int ld_out, int accumulate) {
  reduce_splits_body(ws, splits, split_stride, rows, cols, ld_ws, col_perm, out, ld_out, accumulate,
                     (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// every reduction of a launch plan in one launch: blockIdx.y selects the descriptor
constexpr int kReduceBatch = 32;
struct ReduceBatch {
  nb2_reduce_desc d[kReduceBatch];
};
__global__ void reduce_splits_batch_kernel(const __grid_constant__ ReduceBatch b) {
  const nb2_reduce_desc& d = b.d[blockIdx.y];
  reduce_splits_body(d.ws, d.splits, d.split_stride, d.rows, d.cols, d.ld_ws, d.col_perm, d.out, d.ld_out, d.accumulate,
                     (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

}  // namespace nb2

using namespace nb2;

// ---- tensor maps -----------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime (no link against libcuda) and its
// results are cached per (pointer, shape, box): a training step re-uses the same few dozen buffers.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct MapKey {
  const void* ptr;
  int64_t rows, cols, ld;
  int box_rows;          // + 1000 * box columns when they are not 64 (the 32 x 32, 64-byte-swizzle boxes of the 16-warp epilogue)
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1000003u ^ std::hash<int64_t>()(k.rows);
    h = h * 1000003u ^ std::hash<int64_t>()(k.cols);
    h = h * 1000003u ^ std::hash<int64_t>()(k.ld);
    return h * 1000003u ^ (size_t)k.box_rows;
  }
};
static std::mutex g_map_mutex;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
static EncodeTiledFn g_encode = nullptr;

// [rows][cols] bf16, row stride ld elements; boxes of box_rows x 64 columns, 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols = 64) {
  std::lock_guard<std::mutex> lock(g_map_mutex);
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("gemm: cuTensorMapEncodeTiled is not available from this driver (%s)", e == cudaSuccess ? "symbol not found" : cudaGetErrorString(e));
      return NB2_ERR_CUDA;
    }
    g_encode = (EncodeTiledFn)fn;
  }
  const MapKey key{ptr, rows, cols, ld, box_rows + (box_cols == 64 ? 0 : 1000 * box_cols)};
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) {
    *out = it->second;
    return NB2_OK;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm: cuTensorMapEncodeTiled failed (%d) for a %lld x %lld matrix with row stride %lld, box %d x 64", (int)r, (long long)rows,
              (long long)cols, (long long)ld, box_rows);
    return NB2_ERR_CUDA;
  }
  if (g_map_cache.size() > 8192) g_map_cache.clear();
  g_map_cache.emplace(key, *out);
  return NB2_OK;
}

extern "C" int nb2_gemm_bf16(nb2_handle* h, const nb2_gemm_desc* d, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(d != nullptr, "gemm: null descriptor");
  if (d->M == 0 || d->N == 0) return NB2_OK;
  NB2_CHECK_ARG(d->M > 0 && d->N > 0 && d->n_seg >= 1 && d->n_seg <= NB2_GEMM_MAX_SEG, "gemm: bad shape / segment count");
  NB2_CHECK_ARG(d->out_f32 || d->out_hi, "gemm: no output");
  NB2_CHECK_ARG(!d->out_lo || d->out_hi, "gemm: out_lo needs out_hi");
  const int splits = d->splits > 1 ? d->splits : 1;
  for (int s = 0; s < d->n_seg; ++s) {
    const nb2_gemm_operand& A = d->seg[s].a;
    const nb2_gemm_operand& B = d->seg[s].b;
    NB2_CHECK_ARG(A.ptr && B.ptr && d->seg[s].K > 0, "gemm: segment %d has a null operand or K <= 0", s);
    NB2_CHECK_ARG((A.ld & 7) == 0 && (B.ld & 7) == 0, "gemm: leading dimensions must be multiples of 8 elements (16-byte rows)");
    NB2_CHECK_ARG(((uintptr_t)A.ptr & 15) == 0 && ((uintptr_t)B.ptr & 15) == 0, "gemm: operands must be 16-byte aligned");
    // 16-byte units must be wholly inside or outside the valid extent of the contiguous dimension
    NB2_CHECK_ARG(A.mn_major ? (d->M & 7) == 0 || A.ld >= ((d->M + 7) & ~7LL) : (d->seg[s].K & 7) == 0 || A.ld >= ((d->seg[s].K + 7) & ~7LL),
                  "gemm: segment %d: A rows must be padded to a multiple of 8 elements", s);
    NB2_CHECK_ARG(B.mn_major ? (d->N & 7) == 0 || B.ld >= ((d->N + 7) & ~7) : (d->seg[s].K & 7) == 0 || B.ld >= ((d->seg[s].K + 7) & ~7LL),
                  "gemm: segment %d: B rows must be padded to a multiple of 8 elements", s);
    if (splits > 1) NB2_CHECK_ARG(d->seg[s].K == d->seg[0].K, "gemm: split-K needs the same K in every segment");
  }
  NB2_CHECK_ARG(splits == 1 || (d->out_f32 && !d->out_hi && !d->bias && d->act == 0 && !d->mask),
                "gemm: split-K writes plain fp32 partial sums (no bias / activation / mask / 16-bit output)");
  GemmKParams p;
  p.d = *d;
  const int n_pad = (d->N + 31) / 32 * 32;
  p.bn = n_pad <= 256 ? n_pad : 256;
  // wide outputs: balance the N tiles (e.g. N = 320 -> 2 x 160) so no tile is mostly padding
  p.n_tiles = (n_pad + 255) / 256;
  if (p.n_tiles > 1) p.bn = ((n_pad / 32 + p.n_tiles - 1) / p.n_tiles) * 32;
  p.m_tiles = (int)((d->M + 127) / 128);
  p.k_per_split = 0;
  if (splits > 1) p.k_per_split = ((d->seg[0].K + splits - 1) / splits + 63) / 64 * 64;
  // split-precision triples (A_lo B_hi, A_hi B_lo, A_hi B_hi over the same K): the kernel loads the four tiles of a K chunk once
  // Only for the weight-gradient shape (A = dY^T, MN-major): there the kernel is bound by the L2 -> shared stream.  The forward
  // and dgrad shapes are bound by their epilogue (measured: no gain), and keep the more accurate order of additions -- all
  // cross terms of the layer first, then hi x hi -- which sharing tiles per K chunk would interleave.
  p.fused3 = (d->n_seg % 3 == 0 && d->seg[0].a.mn_major) ? 1 : 0;
  for (int t = 0; p.fused3 && t < d->n_seg; t += 3) {
    const nb2_gemm_operand &a0 = d->seg[t].a, &a1 = d->seg[t + 1].a, &a2 = d->seg[t + 2].a;
    const nb2_gemm_operand &b0 = d->seg[t].b, &b1 = d->seg[t + 1].b, &b2 = d->seg[t + 2].b;
    const bool same = d->seg[t].K == d->seg[t + 1].K && d->seg[t].K == d->seg[t + 2].K && a1.ptr == a2.ptr && a1.ld == a2.ld && b0.ptr == b2.ptr &&
                      b0.ld == b2.ld && a0.ld == a1.ld && b0.ld == b1.ld && a0.mn_major == a1.mn_major && a1.mn_major == a2.mn_major &&
                      b0.mn_major == b1.mn_major && b1.mn_major == b2.mn_major;
    if (!same) p.fused3 = 0;
  }
  if (h->tc_debug & 8) p.fused3 = 0;      // NB2_TC_DEBUG & 8 (A/B timing): every pass loads its own tiles
  p.reuse3 = 0;
  if (!p.fused3 && d->n_seg == 3 && splits == 1 && !(h->tc_debug & 8)) {
    const nb2_gemm_operand &a1 = d->seg[1].a, &a2 = d->seg[2].a, &b0 = d->seg[0].b, &b2 = d->seg[2].b;
    p.reuse3 = d->seg[0].K == 64 * kGStages && d->seg[1].K == d->seg[0].K && d->seg[2].K == d->seg[0].K && a1.ptr == a2.ptr && a1.ld == a2.ld &&
               a1.mn_major == a2.mn_major && b0.ptr == b2.ptr && b0.ld == b2.ld && b0.mn_major == b2.mn_major &&
               d->seg[0].a.mn_major == a1.mn_major && d->seg[1].b.mn_major == b0.mn_major;
  }
  p.rowsum = 0;
  if (d->a_rowsum_out) {
    const int64_t n_items = (int64_t)p.m_tiles * p.n_tiles * splits;
    NB2_CHECK_ARG(p.fused3 && d->n_seg == 3 && n_items <= h->sm_count && d->a_rowsum_stride >= d->M,
                  "gemm: a_rowsum_out needs the weight-gradient shape (MN-major split-precision triple), one work item per SM "
                  "(%lld items, %d SMs) and a_rowsum_stride >= M", (long long)n_items, h->sm_count);
    p.rowsum = 1;
  }
  p.dbg = h->tc_debug;
  int rc = kernel_set_smem(h, (const void*)gemm_bf16_kernel<8>, kGSmem);
  if (rc != NB2_OK) return rc;
  for (int s = 0; s < d->n_seg; ++s) {
    const nb2_gemm_operand& A = d->seg[s].a;
    const nb2_gemm_operand& B = d->seg[s].b;
    const int64_t K = d->seg[s].K, K8 = (K + 7) & ~7LL, M8 = (d->M + 7) & ~7LL, N8 = ((int64_t)d->N + 7) & ~7LL;
    // the contiguous extent is declared up to its zero padding (a multiple of 8 elements: 16-byte rows for the TMA unit)
    rc = A.mn_major ? make_map(&p.amap[s], A.ptr, K, M8, A.ld, 64) : make_map(&p.amap[s], A.ptr, d->M, K8, A.ld, 128);
    if (rc != NB2_OK) return rc;
    rc = B.mn_major ? make_map(&p.bmap[s], B.ptr, K, N8, B.ld, 64) : make_map(&p.bmap[s], B.ptr, d->N, K8, B.ld, p.bn);
    if (rc != NB2_OK) return rc;
  }
  p.tma_out = 0;
  const bool tma_ok = d->out_hi && (d->ld_16 & 7) == 0 && (d->N & 7) == 0 && ((uintptr_t)d->out_hi & 15) == 0 &&
                      (!d->out_lo || ((uintptr_t)d->out_lo & 15) == 0) && !(h->tc_debug & 64);   // NB2_TC_DEBUG & 64 (A/B timing): copy loops
  // the 16-warp epilogue: bf16 outputs only, whole 32-column passes, two column halves per tile, vector-aligned bias / mask
  const bool epi16 = tma_ok && !d->out_f32 && splits == 1 && (d->N & 31) == 0 && (p.bn & 63) == 0 && (!d->bias || ((uintptr_t)d->bias & 15) == 0) &&
                     (!d->mask || ((d->ld_mask & 7) == 0 && ((uintptr_t)d->mask & 15) == 0)) && d->M >= 8192 && !(h->tc_debug & 256);   // NB2_TC_DEBUG & 256 (A/B timing): the 8-warp epilogue
  if (tma_ok) {
    const int bc = epi16 ? 32 : 64;
    rc = make_map(&p.omap[0], d->out_hi, d->M, d->N, d->ld_16, 32, bc);
    if (rc != NB2_OK) return rc;
    if (d->out_lo) {
      rc = make_map(&p.omap[1], d->out_lo, d->M, d->N, d->ld_16, 32, bc);
      if (rc != NB2_OK) return rc;
    }
    p.tma_out = 1;
  }
  const int64_t items = (int64_t)p.m_tiles * p.n_tiles * splits;
  const int grid = (int)std::min<int64_t>(items, h->sm_count);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.dynamicSmemBytes = kGSmem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (h->tc_debug & 512) ? 0 : 1;    // NB2_TC_DEBUG & 512 (A/B timing): plain launches
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (epi16) {
    rc = kernel_set_smem(h, (const void*)gemm_bf16_kernel<16>, kGSmem);
    if (rc != NB2_OK) return rc;
    cfg.blockDim = dim3(32 * 18);
    NB2_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<16>, p));
  } else {
    cfg.blockDim = dim3(kGThreads);
    NB2_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<8>, p));
  }
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_to_bf16(nb2_handle* h, const float* src, int64_t rows, int cols, int64_t ld_src, const int* col_perm, void* hi, void* lo,
                           int ld_dst, void* stream) {
  NB2_ENTER(h);
  if (rows == 0) return NB2_OK;
  NB2_CHECK_ARG(src && hi && rows > 0 && cols > 0 && ld_dst >= cols && (ld_dst & 7) == 0 && ld_src >= cols, "to_bf16: bad arguments");
  const int64_t n = rows * ld_dst;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 16);
  to_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, ld_src, col_perm, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_dst);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

extern "C" int nb2_to_bf16_batch(nb2_handle* h, const nb2_to_bf16_desc* d, int n, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(n >= 0 && (d != nullptr || n == 0), "to_bf16_batch: bad arguments");
  for (int i0 = 0; i0 < n; i0 += kToBf16Batch) {
    const int cnt = std::min(kToBf16Batch, n - i0);
    ToBf16Batch b;
    int64_t n_max = 0;
    for (int i = 0; i < cnt; ++i) {
      const nb2_to_bf16_desc& t = d[i0 + i];
      NB2_CHECK_ARG(t.src && t.hi && t.rows > 0 && t.cols > 0 && t.ld_dst >= t.cols && (t.ld_dst & 7) == 0 && t.ld_src >= t.cols,
                    "to_bf16_batch: bad descriptor %d", i0 + i);
      b.d[i] = t;
      n_max = std::max<int64_t>(n_max, t.rows * t.ld_dst);
    }
    const int blocks = (int)std::min<int64_t>((n_max + 255) / 256, (int64_t)h->sm_count * 2);
    to_bf16_batch_kernel<<<dim3(blocks, cnt), 256, 0, (cudaStream_t)stream>>>(b);
    NB2_LAUNCH_CHECK(h);
  }
  return NB2_OK;
}

extern "C" int nb2_reduce_splits(nb2_handle* h, const float* ws, int splits, int64_t split_stride, int rows, int cols, int ld_ws,
                                 const int* col_perm, float* out, int ld_out, int accumulate, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(ws && out && splits >= 1 && rows > 0 && cols > 0 && ld_ws > 0 && ld_out >= cols, "reduce_splits: bad arguments");
  const int64_t n = (int64_t)rows * cols;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8);
  reduce_splits_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ws, splits, split_stride, rows, cols, ld_ws, col_perm, out, ld_out, accumulate);
  NB2_LAUNCH_CHECK(h);
  return NB2_OK;
}

// ---- batched entry points: one host call for a recorded sequence (the training step's launch plans) ----------------------
extern "C" int nb2_gemm_bf16_batch(nb2_handle* h, const nb2_gemm_desc* d, int n, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(n >= 0 && (d != nullptr || n == 0), "gemm_batch: bad arguments");
  for (int i = 0; i < n; ++i) {
    const int rc = nb2_gemm_bf16(h, d + i, stream);
    if (rc != NB2_OK) return rc;
  }
  return NB2_OK;
}

extern "C" int nb2_reduce_splits_batch(nb2_handle* h, const nb2_reduce_desc* d, int n, void* stream) {
  NB2_ENTER(h);
  NB2_CHECK_ARG(n >= 0 && (d != nullptr || n == 0), "reduce_splits_batch: bad arguments");
  for (int i0 = 0; i0 < n; i0 += kReduceBatch) {
    const int cnt = std::min(kReduceBatch, n - i0);
    ReduceBatch b;
    int64_t n_max = 0;
    for (int i = 0; i < cnt; ++i) {
      const nb2_reduce_desc& r = d[i0 + i];
      NB2_CHECK_ARG(r.ws && r.out && r.splits >= 1 && r.rows > 0 && r.cols > 0 && r.ld_ws > 0 && r.ld_out >= r.cols,
                    "reduce_splits_batch: bad descriptor %d", i0 + i);
      // accumulating descriptors read what an earlier descriptor of the same launch may write: keep those in order
      NB2_CHECK_ARG(!r.accumulate, "reduce_splits_batch: descriptor %d accumulates; use nb2_reduce_splits for ordered accumulation", i0 + i);
      b.d[i] = r;
      n_max = std::max<int64_t>(n_max, (int64_t)r.rows * r.cols);
    }
    const int blocks = (int)std::min<int64_t>((n_max + 255) / 256, (int64_t)h->sm_count * 8);
    reduce_splits_batch_kernel<<<dim3(blocks, cnt), 256, 0, (cudaStream_t)stream>>>(b);
    NB2_LAUNCH_CHECK(h);
  }
  return NB2_OK;
}
