/*
 * nerf_b200_debug.h — debug / self-test / micro-benchmark entry points of libnerfb200.so.
 *
 * Not part of the drop-in boundary (include/nerf_b200.h): these exist to validate the tcgen05 building blocks on the
 * device (descriptor conventions, operand layouts) and to measure the hardware facts the kernel design rests on
 * (DESIGN.md section 4).  Same conventions as the main header (extern "C", device pointers, caller's stream).
 */
#ifndef NERF_B200_DEBUG_H_
#define NERF_B200_DEBUG_H_

#include "nerf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Debug instrumentation of the tensor-core MLP kernel.  First call (any arguments) enables it; later calls
 * synchronise the device and copy 16 cycle counters per CTA of the LAST launch into out_host (n_ctas <= 256):
 * weight-streamer / MMA-issuer / slot-group wait and work times (see nb2_api.cu). */
int nb2_debug_tc_profile(nb2_handle* h, long long* out_host, int n_ctas);

/* Debug: MMA micro-benchmark / 2-CTA convention check.  mode 0: cta_group::1 M128 N128, 1: cta_group::1 M128 N256,
 * 2: cta_group::2 M256 N256.  A, B: (256 x 64) bf16 row-major; D_out: (256 x 256) fp32 (rows/cols the mode covers);
 * cycles_out[cta]: cycles per MMA instruction on every SM (all SMs run the loop concurrently).
 * flags: stressors running beside the MMA loop (see nb2_mlp_tc.cu); gsrc_1mb: 1 MB of device memory to stream from. */
int nb2_debug_umma_bench(nb2_handle* h, const void* A_bf16, const void* B_bf16, float* D_out, long long* cycles_out,
                         int mode, int iters, int flags, const void* gsrc_1mb, void* stream);

/* Debug: hardware micro-benchmarks behind the MLP kernel's design numbers (nb2_microbench.cu).
 * kind 0 / 1: TMEM -> register load / register -> TMEM store bandwidth (a0 = warps 4|8|16, a1 = loads in flight per
 * wait 1|2|4, a2 = sweeps over the 128 x 512 x 4 B TMEM);  kind 2: L2 -> shared bulk-copy stream of 16 KB tiles that
 * every CTA reads in the same order (a0 = multicast cluster size 1|2|4|8, a1 = ring stages <= 12, a2 = loads per CTA,
 * a3 = per-cluster address skew in tiles, a4 = 1: even/odd CTAs read alternate tiles like a CTA pair; src = n_chunks x 16 KB).
 * out_dev: 3 x int64 per CTA (cycles, bytes, checksum) in device memory; *grid_out = CTAs launched. */
int nb2_debug_microbench(nb2_handle* h, int kind, int a0, int a1, int a2, int a3, int a4, const void* src, int n_chunks,
                         long long* out_dev, int* grid_out, void* stream);

/* Device-side self-test of the tcgen05 building blocks: D (128x128 fp32) = A (128x64 bf16,
 * row-major) * B^T (128x64 bf16, row-major), computed by ONE UMMA sequence through the same
 * operand swizzle, descriptors, bulk copy, commit and TMEM read-out the MLP kernel uses.
 * scratch_16k: 16 KB of device scratch. */
int nb2_selftest_umma(nb2_handle* h, const void* A_bf16, const void* B_bf16, void* scratch_16k,
                      float* D_out, void* stream);

/* Same product with A written to tensor memory by tcgen05.st and consumed by the A-from-TMEM form of tcgen05.mma
 * (the operand convention of the split-precision kernel in nb2_mlp_tc4.cu). */
int nb2_selftest_umma_ts(nb2_handle* h, const void* A_bf16, const void* B_bf16, void* scratch_16k,
                         float* D_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_DEBUG_H_ */
