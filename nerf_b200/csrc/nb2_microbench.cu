// nb2_microbench.cu — hardware micro-benchmarks that size the MLP kernel's two non-tensor bottlenecks on B200:
//   (1) TMEM <-> register bandwidth (tcgen05.ld / tcgen05.st, 32x32b.x32) with 4 / 8 / 16 warps per CTA: the floor
//       of every per-layer epilogue (a 128 x 256 fp32 accumulator is 128 KB of TMEM reads);
//   (2) L2 -> shared-memory bulk-copy throughput when every CTA streams the SAME 16 KB weight tiles in the same order
//       (what the weight streamer does), with ring depth, cluster multicast and address skew as parameters.
// Debug entry point only (nb2_debug_microbench); nothing on the render path calls into this file.
#include "nb2_common.cuh"
#include "nb2_tc_ptx.cuh"

namespace nb2 {
using namespace ptx;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- (1) TMEM bandwidth -----------------------------------------------------------------------------------------
// `nwarps` (4, 8 or 16) warps sweep all 512 columns of the 128 lanes `iters` times; warp w owns lane quadrant w & 3 and
// the column range [(w >> 2) * 512 / parts, ...).  BATCH loads (stores) are in flight per tcgen05.wait.
// out[3 * cta + 0] = cycles, [1] = bytes moved by the CTA, [2] = checksum (defeats dead-code elimination).
template <int NW, int BATCH, bool STORE>
__global__ void __launch_bounds__(NW * 32, 1) tmem_bw_kernel(long long* __restrict__ out, int iters) {
  constexpr int nwarps = NW;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tptr), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  const int parts = nwarps >> 2;
  const int cols = 512 / parts;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * cols);
    uint32_t r[BATCH][32];
#pragma unroll
    for (int b = 0; b < BATCH; ++b)
#pragma unroll
      for (int j = 0; j < 32; ++j) r[b][j] = threadIdx.x * 33u + j + b;
    // initialise the columns so loads return defined data
    for (int c = 0; c < cols; c += 32) tmem_st32(base + c, r[0]);
    tmem_st_wait();
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
      for (int c = 0; c < cols; c += 32 * BATCH) {
        if (STORE) {
#pragma unroll
          for (int b = 0; b < BATCH; ++b) tmem_st32(base + c + 32 * b, r[b]);
          tmem_st_wait();
#pragma unroll
          for (int b = 0; b < BATCH; ++b) r[b][7] += 1u;
        } else {
#pragma unroll
          for (int b = 0; b < BATCH; ++b) tmem_ld32(base + c + 32 * b, r[b]);
          tmem_ld_wait();
#pragma unroll
          for (int b = 0; b < BATCH; ++b) acc ^= r[b][0] ^ r[b][13] ^ r[b][31];
        }
      }
    }
    t1 = clock64();
    if (STORE) acc = r[0][0];
  }
  // longest warp time of the CTA
  __shared__ long long tmax;
  if (threadIdx.x == 0) tmax = 0;
  __syncthreads();
  if (warp < nwarps && (threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&tmax, (unsigned long long)(t1 - t0));
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    out[3 * blockIdx.x + 0] = tmax;
    out[3 * blockIdx.x + 1] = (long long)iters * 128 * 512 * 4;
  }
  if (acc == 0xdeadbeefu) out[3 * blockIdx.x + 2] = acc;
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---- (2) L2 -> shared bulk-copy stream -----------------------------------------------------------------------------
// One producer thread keeps `stages` 16 KB copies in flight, one consumer thread retires them in order.  Every CTA reads
// chunk (i * step + (pairlike ? blockIdx.x & 1 : 0) + blockIdx.x * skew) % n_chunks of `src` (n_chunks x 16 KB), i.e. with
// skew = 0 all CTAs walk the same addresses at the same time, exactly like the weight streamer.  CL > 1: the CTAs of a cluster
// each fetch 1/CL of every tile and multicast it to the whole cluster.
constexpr int kMbMaxStages = 12;
template <int CL>
__global__ void __launch_bounds__(64, 1)
l2_stream_kernel(const unsigned char* __restrict__ src, int n_chunks, int n_loads, int stages, int skew, int pairlike,
                 long long* __restrict__ out) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* al = smem_dyn + (base - smem_u32(smem_dyn));
  uint64_t* full = reinterpret_cast<uint64_t*>(al + (size_t)stages * kTileBytes);
  uint64_t* empty = full + kMbMaxStages;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), CL);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  const int first = (pairlike ? (int)(blockIdx.x & 1) : 0) + (int)(blockIdx.x / CL) * skew;
  const int step = pairlike ? 2 : 1;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < n_loads; ++i) {
      mbar_wait_cluster(smem_u32(&empty[st]), ph ^ 1u);
      const uint32_t fb = smem_u32(&full[st]);
      mbar_arrive_expect_tx(fb, kTileBytes);
      const unsigned char* s = src + (size_t)((first + i * step) % n_chunks) * kTileBytes;
      if (CL == 1) {
        bulk_g2s(base + st * kTileBytes, s, kTileBytes, fb);
      } else {
        const uint32_t part = kTileBytes / CL;
        bulk_g2s_mcast(base + st * kTileBytes + rank * part, s + rank * part, part, fb, (uint16_t)((1u << CL) - 1u));
      }
      if (++st == (uint32_t)stages) { st = 0; ph ^= 1u; }
    }
  } else if (threadIdx.x == 32) {
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < n_loads; ++i) {
      mbar_wait(smem_u32(&full[st]), ph);
      if (CL == 1) mbar_arrive(smem_u32(&empty[st]));
      else
        for (uint32_t c = 0; c < (uint32_t)CL; ++c) mbar_arrive_remote(smem_u32(&empty[st]), c);
      if (++st == (uint32_t)stages) { st = 0; ph ^= 1u; }
    }
    out[3 * blockIdx.x + 0] = clock64() - t0;
    out[3 * blockIdx.x + 1] = (long long)n_loads * kTileBytes;
    out[3 * blockIdx.x + 2] = al[(size_t)(n_loads % stages) * kTileBytes];
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
}

template <int CL>
static int launch_l2_stream(nb2_handle* h, const void* src, int n_chunks, int n_loads, int stages, int skew, int pairlike,
                            long long* out, int* grid_out, cudaStream_t st) {
  auto kern = l2_stream_kernel<CL>;
  const int smem = stages * kTileBytes + 2048;
  NB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int grid = h->sm_count / CL * CL;
  if (CL > 1) {
    cfg.gridDim = dim3(grid);
    int n = 0;
    NB2_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    grid = std::min(grid, std::max(n, 1) * CL);
  }
  cfg.gridDim = dim3(grid);
  *grid_out = grid;
  NB2_CUDA(cudaLaunchKernelEx(&cfg, kern, (const unsigned char*)src, n_chunks, n_loads, stages, skew, pairlike, out));
  h->launches++;
  return NB2_OK;
}

// kind 0: TMEM load, kind 1: TMEM store  (a0 = warps 4|8|16, a1 = batch 1|2|4, a2 = iterations)
// kind 2: L2 stream                       (a0 = cluster size 1|2|4|8, a1 = ring stages <= 12, a2 = loads per CTA,
//                                          a3 = skew, a4 = pairlike; src = n_chunks x 16 KB device buffer)
// out: 3 x long long per CTA (cycles, bytes, checksum); returns the grid size through *grid_out.
int microbench(nb2_handle* h, int kind, int a0, int a1, int a2, int a3, int a4, const void* src, int n_chunks,
               long long* out, int* grid_out, cudaStream_t st) {
  if (kind == 0 || kind == 1) {
    NB2_CHECK_ARG((a0 == 4 || a0 == 8 || a0 == 16) && (a1 == 1 || a1 == 2 || (a1 == 4 && a0 < 16)) && a2 > 0,
                  "microbench: bad TMEM arguments (warps 4|8|16, batch 1|2|4; 16 warps x batch 4 exceeds the register file)");
    *grid_out = h->sm_count;
    const dim3 g(h->sm_count);
#define NB2_TMEM_CASE(W, B)                                                                     \
  if (a0 == W && a1 == B) {                                                                     \
    if (kind == 0) tmem_bw_kernel<W, B, false><<<g, dim3(W * 32), 0, st>>>(out, a2);            \
    else           tmem_bw_kernel<W, B, true><<<g, dim3(W * 32), 0, st>>>(out, a2);             \
  }
    NB2_TMEM_CASE(4, 1) NB2_TMEM_CASE(4, 2) NB2_TMEM_CASE(4, 4) NB2_TMEM_CASE(8, 1) NB2_TMEM_CASE(8, 2) NB2_TMEM_CASE(8, 4)
    NB2_TMEM_CASE(16, 1) NB2_TMEM_CASE(16, 2)
#undef NB2_TMEM_CASE
    NB2_LAUNCH_CHECK(h);
    return NB2_OK;
  }
  if (kind == 2) {
    NB2_CHECK_ARG(src && n_chunks > 0 && a1 >= 1 && a1 <= kMbMaxStages && a2 > 0, "microbench: bad L2-stream arguments");
    switch (a0) {
      case 1: return launch_l2_stream<1>(h, src, n_chunks, a2, a1, a3, a4, out, grid_out, st);
      case 2: return launch_l2_stream<2>(h, src, n_chunks, a2, a1, a3, a4, out, grid_out, st);
      case 4: return launch_l2_stream<4>(h, src, n_chunks, a2, a1, a3, a4, out, grid_out, st);
      case 8: return launch_l2_stream<8>(h, src, n_chunks, a2, a1, a3, a4, out, grid_out, st);
    }
  }
  set_error("microbench: unknown kind %d / cluster %d", kind, a0);
  return NB2_ERR_INVALID;
}

}  // namespace nb2
