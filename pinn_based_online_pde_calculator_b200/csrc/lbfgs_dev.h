// Device-side pieces of the L-BFGS loop (software.py:499-514): the kernels a CUDA-graph WHILE node replays once per
// objective evaluation.  See lbfgs_ctl.h for the shared controller and engine.cu for the loop drivers.
#pragma once
#include <cuda_runtime.h>

#include "lbfgs_ctl.h"

// xt = x + a_next * d (the trial point of the evaluation about to be enqueued); optional trace of the trial points
cudaError_t lb_begin_eval(int P, const float* x, const float* d, const LbfgsCtl* ctl, float* xt, float* trace, cudaStream_t st);
// consume the evaluation (loss_info row just appended to the ring, g.d in scal[0]): line-search step
cudaError_t lb_post_eval(LbfgsCtl* ctl, const double* ring, const int* ring_pos, int n_info, const double* scal, cudaStream_t st);
// if the line search accepted its point (or this was the initial evaluation): push (s, y), x <- xt, g <- g_new
cudaError_t lb_push(int P, LbfgsCtl* ctl, float* x, float* g, const float* xt, const float* gt, float* Sh, float* Yh, double* rho,
                    double* scal2, cudaStream_t st);
// iteration bookkeeping, new direction (vector-free two-loop recursion), start of the next line search, WHILE condition
cudaError_t lb_direction(int P, LbfgsCtl* ctl, const float* g, const float* Sh, const float* Yh, const double* rho, float* d,
                         double* scratch, const double* scal2, unsigned long long cond_handle, int set_cond, cudaStream_t st);
// d = -H g (Nocedal & Wright alg. 7.4) in its vector-free form: Gram matrix of [S | Y | g] in one multi-block pass, the
// recursion on coefficient vectors, one multi-block combination.  ctl == nullptr: explicit (m <= 10, cnt, head).
cudaError_t lb_two_loop(int P, const LbfgsCtl* ctl, int m, int cnt, int head, const float* g, const float* Sh, const float* Yh,
                        const double* rho, float* d, double* scratch, cudaStream_t st);
// doubles of scratch lb_two_loop / lb_direction need for P parameters
size_t lb_scratch_doubles(int P);
