/* C-ABI of the B200 PINN residual-loss + gradient engine (libpinn_engine.so).
 *
 * Drop-in boundary for the hot path of Cc1-Yy/PINN-based-online-PDE-calculator.
 * Every entry point cites the reference interface (pinn_app/software.py, "sw:")
 * it replaces.  Plain pointers and sizes only; no torch / Python types.
 *
 * Conventions: functions return 0 on success, non-zero on failure with a
 * thread-local message in pinn_last_error().  All kernels are enqueued on the
 * stream set by pinn_engine_set_stream (default: an engine-owned stream); the
 * only host synchronisation is where a host scalar / host buffer is returned.
 * The engine never falls back to the CPU: without a CUDA device create() fails.
 */
#ifndef PINN_ENGINE_H
#define PINN_ENGINE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pinn_engine pinn_engine_t;

/* Problem + network description.  Replaces the closures built by
 * sol_pred_create (sw:207-218), mNN_pred_create (sw:221-234), gov_eqn
 * (sw:283-297) and loss_create (sw:310-383). */
typedef struct pinn_spec {
  int32_t d_in;        /* number of inputs: 1..3 (reference: 2)                        */
  int32_t feat_mode;   /* 0: affine 2(z-lb)/(ub-lb)-1 per input; 1: reference polar map
                          [2(r-lb0)/(ub0-lb0)-1, cos t, sin t]  (sw:172-175)            */
  int32_t n_hidden;    /* hidden layers  (reference kwarg network_size["width"] !)     */
  int32_t width;       /* units/layer    (reference kwarg network_size["depth"] !)     */
  int32_t act_first;   /* 0 tanh, 1 sin  (sw:170,178)                                  */
  int32_t act_hidden;  /* 0 tanh (sw:180-181), 1 sin (extension)                       */
  float scl, epsil;    /* first-layer scale (sw:178), output scale (sw:215)            */
  float lb[3], ub[3];  /* domain bounds (sw:705-707)                                   */
  int32_t n1, n2, mix; /* jet channels of the collocation term: u, first derivatives
                          wrt inputs 0..n1-1, pure second derivatives wrt inputs
                          0..n2-1, and (mix==1) the 0-1 mixed derivative;
                          mix==2 (n2==0): ONE combined second-order channel
                          L = sum_i lap_beta_i d^2/dz_i^2 (Laplacian-type operators)   */
  int32_t n_ops;       /* residual bytecode (see csrc/pinn_common.h, PinnOp)           */
  const int32_t* ops;
  int32_t n_consts;
  const float* consts;
  int32_t n_aux_col;   /* aux columns the residual program reads per collocation point:
                          n_aux_user caller-supplied columns followed by hoisted columns   */
  int32_t n_bc;        /* boundary/initial-condition groups = data loss terms (sw:334) */
  int32_t n_aux_user;  /* caller-supplied aux columns (source terms, data)             */
  int32_t n_aux_ops;   /* "aux program": evaluated ONCE per point when points are set;  */
  const int32_t* aux_ops; /* fills the hoisted columns (jet-free sub-expressions such as
                          source terms / variable coefficients) from coords and user aux */
  float lap_beta[3];   /* mix==2: constant coefficient of d_ii ...                      */
  int32_t lap_aux[3];  /* ... or (>= 0) the aux column holding the per-point coefficient */
} pinn_spec_t;

typedef struct pinn_lbfgs_result {
  int32_t iterations;      /* L-BFGS iterations performed                       */
  int32_t evaluations;     /* objective evaluations (sw:511 num_objective_evaluations) */
  int32_t converged;       /* ||g||_inf <= tol                                  */
  int32_t failed;          /* line search failed                                */
  double final_loss;
} pinn_lbfgs_result_t;

/* per-evaluation callback of L-BFGS: loss_info row (sw:485-488) */
typedef void (*pinn_eval_cb)(const double* loss_info, int32_t n_info, void* user);

const char* pinn_last_error(void);
int pinn_device_count(void);

/* sol_pred_create + loss_create (sw:207, 310).  Fails when no CUDA device. */
int pinn_engine_create(const pinn_spec_t* spec, int device, pinn_engine_t** out);
void pinn_engine_destroy(pinn_engine_t* h);
int pinn_engine_set_stream(pinn_engine_t* h, void* cuda_stream);
/* wait for everything enqueued on the engine stream (needed by callers that consume device outputs on
 * another stream; no entry point synchronises implicitly unless it returns host data) */
int pinn_engine_sync(pinn_engine_t* h);

/* P = number of parameters in jax.flatten_util.ravel_pytree order (sw:466,502) */
int64_t pinn_engine_num_params(pinn_engine_t* h);
/* length of loss_info = 3 + n_bc + 1 (sw:377-378) */
int32_t pinn_engine_num_loss_info(pinn_engine_t* h);
/* points per CTA tile / grid size of the collocation kernel (for roofline maths) */
int32_t pinn_engine_tile_points(pinn_engine_t* h);
/* kernels one loss/gradient evaluation enqueues (an Adam step adds one) -- the bench's gpu_launches claim */
int32_t pinn_engine_launches_per_eval(pinn_engine_t* h);
/* kernels one Adam step of pinn_engine_adam_steps enqueues (evaluation kernels + the fused reduce / loss_info / Adam /
 * re-pack tail; PINN_B200_FUSED_TAIL=0 restores the separate kernels) */
int32_t pinn_engine_launches_per_adam_step(pinn_engine_t* h);
/* kernel family chosen at create: 0 = fp32 SIMT (packed FFMA2), 1 = split-precision mma.sync tensor-core kernel,
 * 2 = experimental tcgen05 family C (opt-in), 3 = tcgen05 family D (bf16x3 split, accumulators in Tensor Memory).
 * Selection: environment PINN_B200_KERNEL = simt | mma | tc | umma | auto (auto: family D for padded widths 128 and
 * 256 with at least three jet channels, the mma.sync kernel for padded width 64, the fp32 kernel for 32). */
int32_t pinn_engine_kernel_kind(pinn_engine_t* h);

/* params pytree <-> flat fp32 vector (sw:142-154 layout, sw:466 order) */
int pinn_engine_set_params(pinn_engine_t* h, const float* flat, int on_device);
int pinn_engine_get_params(pinn_engine_t* h, float* flat_out, int on_device);

/* the `data` dict of sw:572: x_col [n_col,d_in]; per BC group x_bd[i] [n_bd[i],d_in],
 * u_bd[i] [n_bd[i]] (per-point targets, sw:557); optional aux_col [n_col,n_aux_col]
 * and base_col [n_col,K] / base_bd[i] [n_bd[i]] (frozen stage-1 jets, sw:228-232).
 * Host pointers are copied; device pointers (on_device=1) to x_col/aux_col/base_col
 * are BORROWED until the next call; BC arrays are always copied (they are small).
 * Each rank passes ITS shard. */
int pinn_engine_set_points(pinn_engine_t* h, const float* x_col, int64_t n_col, const float* aux_col,
                           const float* base_col, int32_t n_bc, const float* const* x_bd,
                           const float* const* u_bd, const float* const* base_bd, const int64_t* n_bd,
                           int on_device);
/* Pipelined refresh of the point set with HOST buffers of the SAME shapes as the current set (no user
 * aux/base columns): prefetch copies them to staging memory on a separate copy stream, overlapping the
 * step in flight; commit swaps them in on the engine stream.  Neither blocks the host.  The data-side
 * analogue of re-sampling inside the training loop (software.py:708-716) without stalling the step. */
int pinn_engine_prefetch_points(pinn_engine_t* h, const float* x_col, int64_t n_col, int32_t n_bc, const float* const* x_bd,
                                const float* const* u_bd, const int64_t* n_bd);
int pinn_engine_commit_points(pinn_engine_t* h);

/* multi-GPU: the GLOBAL point counts the means are taken over (default: local counts) */
int pinn_engine_set_global_counts(pinn_engine_t* h, int64_t n_col_global, const int64_t* n_bd_global);
/* loss_fun.lw[0] and loss_fun.ref (sw:381-382, 739) */
int pinn_engine_set_loss(pinn_engine_t* h, double lw_eqn, double lref);

/* grad(lossf, has_aux=True)(params, data) (sw:390, 479): writes the flat gradient of
 * loss/lref (device pointer, may be NULL) and the UN-normalised loss_info (host
 * pointer, may be NULL; forces a stream sync when given).  params_dev NULL => the
 * engine's current parameters.  Includes the allreduce when a communicator is set. */
int pinn_engine_loss_grad(pinn_engine_t* h, const float* params_dev, float* grad_out_dev,
                          double* loss_info_host);

/* adam_minimizer (sw:387-393), optax.adam defaults, on the engine's parameters.
 * adam_init == opt.init (sw:400).  adam_steps runs n_steps fused steps as one CUDA
 * graph replayed n_steps times; loss_info rows (per step, sw:425) are returned in
 * loss_rows_host [n_steps][n_info] if non-NULL. lr may change between calls while
 * m, v, count are kept (sw:439-440). */
int pinn_engine_adam_init(pinn_engine_t* h);
int pinn_engine_adam_steps(pinn_engine_t* h, int32_t n_steps, double lr, double* loss_rows_host);
/* The loss_info rows of the LAST pinn_engine_adam_steps call (n_rows <= min(its n_steps, 4096)), with the host
 * synchronisation that call skipped when it was given loss_rows_host == NULL: the host can prepare the next
 * collocation set (sw:416-422) while the steps run and collect the rows for logging (sw:418-419, 425) afterwards. */
int pinn_engine_adam_rows(pinn_engine_t* h, int32_t n_rows, double* loss_rows_host);

/* f_u (sw:213) and gov_eqn (sw:283) without parameter gradient (sw:608-616, 766-770).
 * Outputs may be NULL. jets_out: [n,K]. */
int pinn_engine_eval(pinn_engine_t* h, const float* z, int64_t n, const float* aux, const float* base,
                     float* u_out, float* f_out, float* jets_out, int on_device);

/* lbfgs_optimizer (sw:499-514): tfp.optimizer.lbfgs_minimize restated (m=10,
 * Hager-Zhang line search).  value_unnormalised=1 reproduces the reference's
 * (un-normalised value, normalised gradient) pairing (sw:479-490). */
int pinn_engine_lbfgs(pinn_engine_t* h, int32_t max_iter, double tol, int32_t value_unnormalised,
                      pinn_eval_cb cb, void* user, pinn_lbfgs_result_t* out);
/* The loop runs ON THE DEVICE (one CUDA-graph WHILE node: trial point, evaluation, Hager-Zhang state machine,
 * history push, two-loop recursion, loop condition); the host synchronises once per batch and then delivers the
 * loss_info rows (one per evaluation, in order) to `cb`.  PINN_B200_LBFGS = device | host | legacy selects the
 * device-resident loop (default on one GPU), the same kernels enqueued trip by trip with a synchronisation per
 * evaluation (default with a communicator), or the round-1 host line search; all three give bit-identical iterates. */

/* NCCL data parallelism: one fused allreduce [grad | loss partial sums] per evaluation.
 * pinn_nccl_unique_id fills 128 bytes; every rank then calls pinn_engine_init_nccl. */
int pinn_nccl_unique_id(uint8_t id_out[128]);
int pinn_engine_init_nccl(pinn_engine_t* h, const uint8_t id[128], int32_t rank, int32_t world);

/* Device-side samplers.  pinn_sample_lhs: Latin-hypercube points, the
 * device counterpart of pyDOE.lhs (sw:553,562): out[i][col0+j] = lo_j + (perm_j(i)+U)/n*(hi_j-lo_j),
 * row stride ld.  pinn_sample_cdf2d: colloc2D_set (sw:87-136): cum_host = [0, cumsum(F cells)]
 * over ncy x ncx cells (row-major), lower-left grid corner (x0,y0), cell size (dx,dy). */
int pinn_sample_lhs(int device, void* stream, uint32_t seed, int64_t n, int32_t d, const float* lo,
                    const float* hi, float* out_dev, int32_t ld, int32_t col0);
int pinn_sample_cdf2d(int device, void* stream, uint32_t seed, int64_t n, const double* cum_host, int32_t ncy,
                      int32_t ncx, float x0, float y0, float dx, float dy, float* out_dev, int32_t ld);

/* timing helper: device time (ms) of the last adam_steps / loss_grad call measured
 * with CUDA events on the engine stream */
double pinn_engine_last_ms(pinn_engine_t* h);

#ifdef __cplusplus
}
#endif
#endif
