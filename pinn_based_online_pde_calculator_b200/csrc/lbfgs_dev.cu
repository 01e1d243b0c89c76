// Device-resident L-BFGS loop: kernels replayed by a CUDA-graph WHILE node, one trip per objective evaluation.
// Compiled with -fmad=false: the line-search arithmetic (lbfgs_ctl.h) must round exactly like the host compiler's
// code so that the host-driven and the device-resident loops produce bit-identical iterates.
#include "lbfgs_dev.h"

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* sh) {  // all threads get the sum; fixed order
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;
}
__device__ __forceinline__ double block_max(double v, double* sh) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double m = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, sh[w]);
  return m;
}

__global__ void k_lb_begin(int n, const float* __restrict__ x, const float* __restrict__ d, const LbfgsCtl* __restrict__ ctl,
                           float* __restrict__ xt, float* __restrict__ trace) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = ctl->a_next;
  const float v = (float)fma(a, (double)d[i], (double)x[i]);
  xt[i] = v;
  if (trace && ctl->total_evals < ctl->trace_cap) trace[(size_t)ctl->total_evals * n + i] = v;
}

__global__ void k_lb_post(LbfgsCtl* __restrict__ ctl, const double* __restrict__ ring, const int* __restrict__ ring_pos, int n_info,
                          const double* __restrict__ scal) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  LbfgsCtl s = *ctl;
  const double* row = ring + (size_t)((*ring_pos - 1) % s.ring_cap) * n_info;
  const LsPhi res = ls_result(s, s.a_next, row[0], scal[0]);
  s.total_evals += 1;
  s.rows += 1;
  s.do_push = 0;
  if (s.init_eval) {
    s.fcur = res.f;
    s.do_push = 1;  // "push" in its initial form: g <- g(x0), |g|_inf
  } else {
    s.evals += 1;
    const int r = ls_resume(s, res);
    if (r == LS_FOUND) s.do_push = 1;
    else if (r == LS_FAIL) s.failed = 1;
  }
  *ctl = s;
}

// (s, y) = (xt - x, gt - g) into the history slot `head` when s.y > 0 (a rejected pair must not overwrite the
// oldest live pair); the iterate always moves.  scal2 = {s.y, |g_new|_inf}.  Initial form: g <- gt only.
__global__ void __launch_bounds__(1024) k_lb_push(int n, LbfgsCtl* __restrict__ ctl, float* __restrict__ x, float* __restrict__ g, const float* __restrict__ xt,
                          const float* __restrict__ gt, float* __restrict__ Sh, float* __restrict__ Yh, double* __restrict__ rho,
                          double* __restrict__ scal2) {
  __shared__ double sh[32];
  __shared__ int accept;
  if (!ctl->do_push) return;
  const bool init = ctl->init_eval != 0;
  const int slot = ctl->head;
  float* s = Sh + (size_t)slot * n;
  float* y = Yh + (size_t)slot * n;
  double sy = 0.0, mx = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float si = xt[i] - x[i], yi = gt[i] - g[i];
    sy = fma((double)si, (double)yi, sy);  // explicit fma: what -fmad=true made of the round-1 kernel
    mx = fmax(mx, fabs((double)gt[i]));
  }
  sy = block_sum(sy, sh);
  mx = block_max(mx, sh);
  if (threadIdx.x == 0) {
    scal2[0] = sy;
    scal2[1] = mx;
    accept = (!init && sy > 0.0 && ls_finite(sy)) ? 1 : 0;
    if (accept) rho[slot] = 1.0 / sy;
  }
  __syncthreads();
  const bool acc = accept != 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (acc) { s[i] = xt[i] - x[i]; y[i] = gt[i] - g[i]; }
    x[i] = xt[i]; g[i] = gt[i];
  }
}

// ---------------------------------------------------------------- search direction d = -H g, "vector-free" form
// The two-loop recursion (Nocedal & Wright alg. 7.4) only ever needs inner products between the 2m+1 vectors
// b = [s_0 .. s_{m-1} | y_0 .. y_{m-1} | g] and produces a linear combination of them.  So: ONE multi-block pass computes
// the Gram matrix G = b^T b (fp64, fixed-order partial sums), one thread runs the recursion on coefficient vectors, and
// ONE multi-block pass forms d = -sum_b coef_b b_b.  (The single-block vector form -- 2*cnt dependent dot / axpy sweeps
// with a block-wide reduction each -- took 131 us per direction at the reference's problem size, more than the
// objective evaluation itself.)
constexpr int LB_MAXM = 10, LB_NB = 2 * LB_MAXM + 1, LB_NPAIR = LB_NB * (LB_NB + 1) / 2;   // 21 vectors, 231 pairs
constexpr int LB_GT = 256, LB_TILE = 64, LB_LD = LB_TILE + 1;                              // gram block: threads, staged elements, padded stride

__device__ __forceinline__ const float* lb_vec(int b, int m, int n, const float* g, const float* Sh, const float* Yh) {
  return b < m ? Sh + (size_t)b * n : (b < 2 * m ? Yh + (size_t)(b - m) * n : g);
}
// (cnt, head, m): from the controller when ctl != nullptr (and nothing to do unless it asked for a direction)
__device__ __forceinline__ bool lb_dir_args(const LbfgsCtl* ctl, int& m, int& cnt, int& head) {
  if (ctl) {
    if (!ctl->need_dir) return false;
    m = ctl->m; cnt = ctl->cnt; head = ctl->head;
  }
  return true;
}

// partial Gram matrix of the slice [lo, hi) of the parameter vector by the first LB_GT threads of the block: part[pair],
// pair (a <= b) in row-major upper-triangular order over the 2m+1 vectors (stale history slots included: finite, their
// coefficients are zero).  All threads of the block must call it.
__device__ void lb_gram_slice(int n, int lo, int hi, int m, const float* __restrict__ g, const float* __restrict__ Sh,
                              const float* __restrict__ Yh, double* __restrict__ part, float* tile, unsigned char* pa, unsigned char* pb) {
  const int nb = 2 * m + 1, npair = nb * (nb + 1) / 2;
  const bool worker = threadIdx.x < LB_GT;
  if (worker)
    for (int p = threadIdx.x; p < npair; p += LB_GT) {   // pair index -> (a, b)
      int a = 0, r = p;
      while (r >= nb - a) { r -= nb - a; ++a; }
      pa[p] = (unsigned char)a; pb[p] = (unsigned char)(a + r);
    }
  double acc[(LB_NPAIR + LB_GT - 1) / LB_GT];
#pragma unroll
  for (int k = 0; k < (LB_NPAIR + LB_GT - 1) / LB_GT; ++k) acc[k] = 0.0;
  for (int base = lo; base < hi; base += LB_TILE) {
    __syncthreads();
    if (worker)
      for (int idx = threadIdx.x; idx < nb * LB_TILE; idx += LB_GT) {
        const int b = idx / LB_TILE, e = idx % LB_TILE;
        tile[b * LB_LD + e] = (base + e < hi) ? lb_vec(b, m, n, g, Sh, Yh)[base + e] : 0.f;
      }
    __syncthreads();
    if (worker) {
#pragma unroll
      for (int k = 0; k < (LB_NPAIR + LB_GT - 1) / LB_GT; ++k) {
        const int p = threadIdx.x + k * LB_GT;
        if (p < npair) {
          const float* va = tile + pa[p] * LB_LD;
          const float* vb = tile + pb[p] * LB_LD;
          double t = acc[k];
#pragma unroll 8
          for (int e = 0; e < LB_TILE; ++e) t = fma((double)va[e], (double)vb[e], t);
          acc[k] = t;
        }
      }
    }
  }
  if (worker) {
#pragma unroll
    for (int k = 0; k < (LB_NPAIR + LB_GT - 1) / LB_GT; ++k) {
      const int p = threadIdx.x + k * LB_GT;
      if (p < npair) part[p] = acc[k];
    }
  }
}

__global__ void __launch_bounds__(LB_GT) k_lb_gram(int n, const LbfgsCtl* __restrict__ ctl, int m, int cnt, int head, const float* __restrict__ g,
                                                   const float* __restrict__ Sh, const float* __restrict__ Yh, double* __restrict__ part) {
  __shared__ float tile[LB_NB * LB_LD];
  __shared__ unsigned char pa[LB_NPAIR], pb[LB_NPAIR];
  if (!lb_dir_args(ctl, m, cnt, head)) return;
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(n, lo + per);
  lb_gram_slice(n, lo, hi, m, g, Sh, Yh, part + (size_t)blockIdx.x * LB_NPAIR, tile, pa, pb);
}

// Gram matrix = fixed-order sum of the partials; two-loop recursion on coefficient vectors; coef[b] of d = -sum coef_b b_b.
// All threads of the block must call it; the first LB_GT threads work.
__device__ void lb_coef_block(int m, int cnt, int head, int nblocks, const double* __restrict__ part, const double* __restrict__ rho,
                              double* __restrict__ coef, double (*G)[LB_NB]) {
  const int nb = 2 * m + 1, npair = nb * (nb + 1) / 2;
  if (threadIdx.x < LB_GT)
    for (int p = threadIdx.x; p < npair; p += LB_GT) {
      int a = 0, r = p;
      while (r >= nb - a) { r -= nb - a; ++a; }
      double t = 0.0;
      for (int k = 0; k < nblocks; ++k) t += part[(size_t)k * LB_NPAIR + p];
      G[a][a + r] = t;
      G[a + r][a] = t;
    }
  __syncthreads();
  if (threadIdx.x < 32) {
    // warp 0: lane b holds the coefficient of vector b (nb <= 32); a row product is one multiply per lane and a
    // fixed-order butterfly sum.  Dead history slots have coefficient 0 and a finite Gram row: they add exact zeros.
    const int lane = threadIdx.x;
    const bool live = lane < nb;
    double c = (lane == 2 * m) ? 1.0 : 0.0;            // q = g
    double alpha = 0.0;                                // lane sl keeps alpha of slot sl
    for (int j = 0; j < cnt; ++j) {                    // newest -> oldest
      const int sl = ((head - 1 - j) % m + m) % m;
      const double a = rho[sl] * warp_sum(live ? c * G[sl][lane] : 0.0);
      if (lane == sl) alpha = a;
      if (lane == m + sl) c -= a;
    }
    if (cnt > 0) {
      const int sl = ((head - 1) % m + m) % m;
      c *= G[sl][m + sl] / G[m + sl][m + sl];
    }
    for (int j = cnt - 1; j >= 0; --j) {               // oldest -> newest
      const int sl = ((head - 1 - j) % m + m) % m;
      const double beta = rho[sl] * warp_sum(live ? c * G[m + sl][lane] : 0.0);
      if (lane == sl) c += alpha - beta;
    }
    if (live) coef[lane] = c;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LB_GT) k_lb_coef(const LbfgsCtl* __restrict__ ctl, int m, int cnt, int head, int nblocks,
                                                   const double* __restrict__ part, const double* __restrict__ rho, double* __restrict__ coef) {
  __shared__ double G[LB_NB][LB_NB];
  if (!lb_dir_args(ctl, m, cnt, head)) return;
  lb_coef_block(m, cnt, head, nblocks, part, rho, coef, G);
}

// d_i = -(sum over the live vectors, fixed order) coef_b * b_b[i]
__device__ __forceinline__ float lb_combine_one(int i, int n, int m, int cnt, int head, const float* __restrict__ g, const float* __restrict__ Sh,
                                                const float* __restrict__ Yh, const double* __restrict__ coef) {
  double t = coef[2 * m] * (double)g[i];
  for (int j = 0; j < cnt; ++j) {
    const int sl = ((head - 1 - j) % m + m) % m;
    t = fma(coef[sl], (double)Sh[(size_t)sl * n + i], t);
    t = fma(coef[m + sl], (double)Yh[(size_t)sl * n + i], t);
  }
  return (float)(-t);
}
__global__ void k_lb_combine(int n, const LbfgsCtl* __restrict__ ctl, int m, int cnt, int head, const float* __restrict__ g,
                             const float* __restrict__ Sh, const float* __restrict__ Yh, const double* __restrict__ coef, float* __restrict__ d) {
  if (!lb_dir_args(ctl, m, cnt, head)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = lb_combine_one(i, n, m, cnt, head, g, Sh, Yh, coef);
}
// small parameter vectors (one Gram block): the three steps in ONE single-block launch, same arithmetic
__global__ void __launch_bounds__(1024) k_lb_two_loop_small(int n, int m, int cnt, int head, const float* __restrict__ g, const float* __restrict__ Sh,
                                                            const float* __restrict__ Yh, const double* __restrict__ rho, double* __restrict__ scratch,
                                                            float* __restrict__ d) {
  __shared__ float tile[LB_NB * LB_LD];
  __shared__ unsigned char pa[LB_NPAIR], pb[LB_NPAIR];
  __shared__ double G[LB_NB][LB_NB];
  lb_gram_slice(n, 0, n, m, g, Sh, Yh, scratch, tile, pa, pb);
  __threadfence_block();
  __syncthreads();
  lb_coef_block(m, cnt, head, 1, scratch, rho, scratch + LB_NPAIR, G);
  __threadfence_block();
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = lb_combine_one(i, n, m, cnt, head, g, Sh, Yh, scratch + LB_NPAIR);
}

// iteration bookkeeping after the push: does the loop need a new search direction?
__global__ void k_lb_pre(LbfgsCtl* __restrict__ ctl, const double* __restrict__ scal2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  LbfgsCtl s = *ctl;
  if (s.do_push) {
    if (s.init_eval) {
      s.ginf = scal2[1];
      s.converged = s.ginf <= s.tol;
      s.init_eval = 0;
    } else {
      lb_after_push(s, scal2[0], scal2[1]);
    }
  }
  s.need_dir = (s.do_push && !s.converged && !s.failed && s.iter < s.max_iter) ? 1 : 0;
  s.do_push = 0;
  *ctl = s;
}

// with the new direction in d: g.d, start of the next line search; WHILE condition
__global__ void __launch_bounds__(1024) k_lb_direction(int n, LbfgsCtl* __restrict__ ctl, const float* __restrict__ g, const float* __restrict__ d,
                               cudaGraphConditionalHandle handle, int set_cond) {
  __shared__ double sh[32];
  if (ctl->need_dir) {
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) t = fma((double)g[i], (double)d[i], t);
    t = block_sum(t, sh);
    if (threadIdx.x == 0) {
      LbfgsCtl s = *ctl;
      if (!(t < 0.0) || !ls_finite(t)) {
        s.failed = 1;
      } else {
        ls_begin(s, s.fcur, t);
        ls_resume(s, LsPhi{0.0, 0.0, 0.0});  // first request: step 1
      }
      s.need_dir = 0;
      *ctl = s;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    LbfgsCtl s = *ctl;
    // leave the loop when the optimisation ended or the host has to drain the loss_info ring / trace
    const bool done = s.converged || s.failed || s.iter >= s.max_iter;
    s.stop = (done || s.rows >= s.ring_cap) ? 1 : 0;
    *ctl = s;
    if (set_cond) cudaGraphSetConditional(handle, s.stop ? 0u : 1u);
  }
}

// small parameter vectors: bookkeeping, direction and line-search start in ONE single-block launch (same arithmetic as
// k_lb_pre -> k_lb_gram -> k_lb_coef -> k_lb_combine -> k_lb_direction)
__global__ void __launch_bounds__(1024) k_lb_direction_small(int n, LbfgsCtl* __restrict__ ctl, const float* __restrict__ g, const float* __restrict__ Sh,
                                     const float* __restrict__ Yh, const double* __restrict__ rho, float* __restrict__ d,
                                     double* __restrict__ scratch, const double* __restrict__ scal2, cudaGraphConditionalHandle handle,
                                     int set_cond) {
  __shared__ float tile[LB_NB * LB_LD];
  __shared__ unsigned char pa[LB_NPAIR], pb[LB_NPAIR];
  __shared__ double G[LB_NB][LB_NB];
  __shared__ double sh[32];
  __shared__ int need_dir, m, cnt, head;
  if (threadIdx.x == 0) {
    LbfgsCtl s = *ctl;
    if (s.do_push) {
      if (s.init_eval) {
        s.ginf = scal2[1];
        s.converged = s.ginf <= s.tol;
        s.init_eval = 0;
      } else {
        lb_after_push(s, scal2[0], scal2[1]);
      }
    }
    s.need_dir = (s.do_push && !s.converged && !s.failed && s.iter < s.max_iter) ? 1 : 0;
    s.do_push = 0;
    need_dir = s.need_dir; m = s.m; cnt = s.cnt; head = s.head;
    *ctl = s;
  }
  __syncthreads();
  if (need_dir) {
    lb_gram_slice(n, 0, n, m, g, Sh, Yh, scratch, tile, pa, pb);
    __threadfence_block();
    __syncthreads();
    lb_coef_block(m, cnt, head, 1, scratch, rho, scratch + LB_NPAIR, G);
    __threadfence_block();
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = lb_combine_one(i, n, m, cnt, head, g, Sh, Yh, scratch + LB_NPAIR);
    __threadfence_block();
    __syncthreads();
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) t = fma((double)g[i], (double)d[i], t);
    t = block_sum(t, sh);
    if (threadIdx.x == 0) {
      LbfgsCtl s = *ctl;
      if (!(t < 0.0) || !ls_finite(t)) {
        s.failed = 1;
      } else {
        ls_begin(s, s.fcur, t);
        ls_resume(s, LsPhi{0.0, 0.0, 0.0});  // first request: step 1
      }
      s.need_dir = 0;
      *ctl = s;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    LbfgsCtl s = *ctl;
    const bool done = s.converged || s.failed || s.iter >= s.max_iter;
    s.stop = (done || s.rows >= s.ring_cap) ? 1 : 0;
    *ctl = s;
    if (set_cond) cudaGraphSetConditional(handle, s.stop ? 0u : 1u);
  }
}

}  // namespace

cudaError_t lb_begin_eval(int P, const float* x, const float* d, const LbfgsCtl* ctl, float* xt, float* trace, cudaStream_t st) {
  k_lb_begin<<<(P + 255) / 256, 256, 0, st>>>(P, x, d, ctl, xt, trace);
  return cudaGetLastError();
}
cudaError_t lb_post_eval(LbfgsCtl* ctl, const double* ring, const int* ring_pos, int n_info, const double* scal, cudaStream_t st) {
  k_lb_post<<<1, 32, 0, st>>>(ctl, ring, ring_pos, n_info, scal);
  return cudaGetLastError();
}
cudaError_t lb_push(int P, LbfgsCtl* ctl, float* x, float* g, const float* xt, const float* gt, float* Sh, float* Yh, double* rho,
                    double* scal2, cudaStream_t st) {
  k_lb_push<<<1, 1024, 0, st>>>(P, ctl, x, g, xt, gt, Sh, Yh, rho, scal2);
  return cudaGetLastError();
}
int lb_gram_blocks(int P) { const int b = (P + 1023) / 1024; return b < 1 ? 1 : (b > 64 ? 64 : b); }
size_t lb_scratch_doubles(int P) { return (size_t)lb_gram_blocks(P) * LB_NPAIR + LB_NB; }

// d = -H g from the history (cnt pairs, newest at slot head-1): Gram matrix -> coefficients -> combination.
// ctl != nullptr: (m, cnt, head) come from the controller and nothing happens unless it asked for a direction.
cudaError_t lb_two_loop(int P, const LbfgsCtl* ctl, int m, int cnt, int head, const float* g, const float* Sh, const float* Yh,
                        const double* rho, float* d, double* scratch, cudaStream_t st) {
  if (!ctl && m > LB_MAXM) return cudaErrorInvalidValue;
  const int gb = lb_gram_blocks(P);
  if (gb == 1 && !ctl) {   // small parameter vector: one single-block launch
    k_lb_two_loop_small<<<1, 1024, 0, st>>>(P, m, cnt, head, g, Sh, Yh, rho, scratch, d);
    return cudaGetLastError();
  }
  double* part = scratch;
  double* coef = scratch + (size_t)gb * LB_NPAIR;
  k_lb_gram<<<gb, LB_GT, 0, st>>>(P, ctl, m, cnt, head, g, Sh, Yh, part);
  k_lb_coef<<<1, LB_GT, 0, st>>>(ctl, m, cnt, head, gb, part, rho, coef);
  k_lb_combine<<<(P + 255) / 256, 256, 0, st>>>(P, ctl, m, cnt, head, g, Sh, Yh, coef, d);
  return cudaGetLastError();
}

cudaError_t lb_direction(int P, LbfgsCtl* ctl, const float* g, const float* Sh, const float* Yh, const double* rho, float* d,
                         double* scratch, const double* scal2, unsigned long long cond_handle, int set_cond, cudaStream_t st) {
  if (lb_gram_blocks(P) == 1) {   // small parameter vector: bookkeeping + direction + line-search start in one launch
    k_lb_direction_small<<<1, 1024, 0, st>>>(P, ctl, g, Sh, Yh, rho, d, scratch, scal2, (cudaGraphConditionalHandle)cond_handle, set_cond);
    return cudaGetLastError();
  }
  k_lb_pre<<<1, 32, 0, st>>>(ctl, scal2);
  cudaError_t e = lb_two_loop(P, ctl, 0, 0, 0, g, Sh, Yh, rho, d, scratch, st);
  if (e != cudaSuccess) return e;
  k_lb_direction<<<1, 1024, 0, st>>>(P, ctl, g, d, (cudaGraphConditionalHandle)cond_handle, set_cond);
  return cudaGetLastError();
}
