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

// two-loop recursion (Nocedal & Wright alg. 7.4), single block; same arithmetic as k_lbfgs_direction
__device__ void two_loop(int n, int m, int cnt, int head, const float* __restrict__ g, const float* __restrict__ Sh,
                         const float* __restrict__ Yh, const double* __restrict__ rho, float* __restrict__ d, double* __restrict__ alpha,
                         double* sh) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = g[i];
  __syncthreads();
  for (int j = 0; j < cnt; ++j) {
    const int slot = ((head - 1 - j) % m + m) % m;
    const float* s = Sh + (size_t)slot * n;
    const float* y = Yh + (size_t)slot * n;
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) t = fma((double)s[i], (double)d[i], t);
    t = block_sum(t, sh);
    const double a = rho[slot] * t;
    if (threadIdx.x == 0) alpha[slot] = a;
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = (float)fma(-a, (double)y[i], (double)d[i]);
    __syncthreads();
  }
  if (cnt > 0) {
    const int slot = ((head - 1) % m + m) % m;
    const float* s = Sh + (size_t)slot * n;
    const float* y = Yh + (size_t)slot * n;
    double sy = 0.0, yy = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      sy = fma((double)s[i], (double)y[i], sy);
      yy = fma((double)y[i], (double)y[i], yy);
    }
    sy = block_sum(sy, sh);
    yy = block_sum(yy, sh);
    const double gamma = sy / yy;
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = (float)(gamma * (double)d[i]);
    __syncthreads();
  }
  for (int j = cnt - 1; j >= 0; --j) {
    const int slot = ((head - 1 - j) % m + m) % m;
    const float* s = Sh + (size_t)slot * n;
    const float* y = Yh + (size_t)slot * n;
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) t = fma((double)y[i], (double)d[i], t);
    t = block_sum(t, sh);
    const double b = rho[slot] * t;
    const double a = alpha[slot];
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = (float)fma(a - b, (double)s[i], (double)d[i]);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = -d[i];
  __syncthreads();
}

__global__ void __launch_bounds__(1024) k_lb_direction(int n, LbfgsCtl* __restrict__ ctl, const float* __restrict__ g, const float* __restrict__ Sh,
                               const float* __restrict__ Yh, const double* __restrict__ rho, float* __restrict__ d,
                               double* __restrict__ alpha, const double* __restrict__ scal2, cudaGraphConditionalHandle handle,
                               int set_cond) {
  __shared__ double sh[32];
  __shared__ int need_dir, cnt, head;
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
    need_dir = s.need_dir; cnt = s.cnt; head = s.head;
    *ctl = s;
  }
  __syncthreads();
  if (need_dir) {
    two_loop(n, ctl->m, cnt, head, g, Sh, Yh, rho, d, alpha, sh);
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
cudaError_t lb_direction(int P, LbfgsCtl* ctl, const float* g, const float* Sh, const float* Yh, const double* rho, float* d,
                         double* alpha, const double* scal2, unsigned long long cond_handle, int set_cond, cudaStream_t st) {
  k_lb_direction<<<1, 1024, 0, st>>>(P, ctl, g, Sh, Yh, rho, d, alpha, scal2, (cudaGraphConditionalHandle)cond_handle, set_cond);
  return cudaGetLastError();
}
