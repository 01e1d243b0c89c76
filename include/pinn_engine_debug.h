/* Measurement, profiling and test hooks of libpinn_engine.so.  NOT part of the drop-in boundary
 * (include/pinn_engine.h): nothing here replaces a reference interface; bench.py, tools/ and tests/ use it. */
#ifndef PINN_ENGINE_DEBUG_H
#define PINN_ENGINE_DEBUG_H
#include "pinn_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

/* fp32 FMA-pipe microbenchmark (roofline denominator of the SIMT path):
 * returns achieved TFLOP/s; variant 0 = scalar FFMA, 1 = packed fma.rn.f32x2 */
int pinn_fma_peak(int device, int variant, double* tflops_out);
/* phase clocks (CTA 0) of the last launch of the experimental tcgen05 kernel family (PINN_B200_KERNEL=umma) */
int pinn_engine_umma_clocks(pinn_engine_t* h, long long* out8);
/* tcgen05 probe (measurement helper): D = A * B^T on one CTA, tf32 inputs / fp32 TMEM accumulator.  A, B0, B1
 * are RAW shared-memory images (the host lays the operands out), cfg = {M, N, k-steps, A MN-major, B MN-major,
 * products (2 = second one with B1 into lanes +16, M = 64), repetitions, words of A, words of B,
 * A: LBO, SBO, k-step advance in bytes, B: the same}; dumps TMEM as out[128 lanes][512 columns]. */
int pinn_umma_probe(int device, const float* A, const float* B0, const float* B1, const int* cfg, float* out,
                    double* cycles, int* status);

/* roofline helper: average device time (ms) of the collocation kernel and of the
 * boundary kernel launched alone, CUDA events on the engine stream, an L2 flush of
 * flush_bytes between launches. */
int pinn_engine_time_kernels(pinn_engine_t* h, int32_t reps, int64_t flush_bytes, double* col_ms, double* bc_ms);

/* phase profile of the collocation kernel (tensor-core kernel only): clock64 totals of CTA 0 for
 * {fwd GEMM, activation fwd, output+residual, activation bwd, smem restage, wgrad, dgrad, rest} */
int pinn_engine_phase_profile(pinn_engine_t* h, int64_t* out8);

/* L-BFGS test hooks: keep the first `cap` trial parameter vectors of the next pinn_engine_lbfgs calls (cap = 0
 * switches the trace off), read them back, and the number of host synchronisations of the last call. */
int pinn_engine_lbfgs_trace(pinn_engine_t* h, int32_t cap);
int32_t pinn_engine_lbfgs_trace_rows(pinn_engine_t* h);
int pinn_engine_lbfgs_trace_get(pinn_engine_t* h, float* out_host, int32_t rows);
int32_t pinn_engine_lbfgs_host_syncs(pinn_engine_t* h);

/* L-BFGS search direction in isolation (test hook): d = -H g from host arrays S, Y [m][n] (m <= 10), rho [m], the live
 * pair count cnt and the ring head (newest pair at slot head - 1), through the engine's vector-free two-loop recursion
 * (Gram matrix -> coefficient recursion -> combination; csrc/lbfgs_dev.cu).  No engine handle needed. */
int pinn_lbfgs_direction_test(int device, int32_t n, int32_t m, int32_t cnt, int32_t head, const float* g, const float* S,
                              const float* Y, const double* rho, float* d_out);

#ifdef __cplusplus
}
#endif
#endif
