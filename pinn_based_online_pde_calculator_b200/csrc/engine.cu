// C-ABI implementation of include/pinn_engine.h: per-call engine handle, device
// buffers, kernel dispatch, CUDA-graph Adam loop, on-device L-BFGS, NCCL hook.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/pinn_engine.h"
#include "../../include/pinn_engine_debug.h"
#include "aux_kernels.cuh"
#include "sampler_kernels.cuh"
#include "jet_launch.h"
#include "jet_umma.h"
#include "jet_tc.h"
#include "lbfgs_dev.h"
#include "pinn_common.h"

static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

extern "C" const char* pinn_last_error(void) { return g_err.c_str(); }
extern "C" int pinn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// ---------------------------------------------------------------- NCCL via dlopen
struct Id128 { char b[128]; };  // ncclUniqueId (passed by value)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.lib) return 0;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail("cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) return fail("libnccl: missing symbols");
  g_nccl.lib = lib;
  return 0;
}

// ---------------------------------------------------------------- engine
struct PointSet {
  float* coords = nullptr;  // device
  float* aux = nullptr;
  float* base = nullptr;
  bool own_coords = false, own_aux = false, own_base = false;
  size_t cap_coords = 0, cap_aux = 0, cap_base = 0;  // owned capacities (floats)
};

struct pinn_engine {
  int device = 0, num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // staged (prefetched) copy of the next point set: filled on copy_stream while the engine stream computes
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_staged = nullptr, ev_consumed = nullptr;
  float *stg_col = nullptr, *stg_bc = nullptr, *stg_ubc = nullptr;
  size_t stg_col_cap = 0, stg_bc_cap = 0;
  bool staged = false;
  // the boundary kernel runs on a forked stream next to the collocation kernel (own accumulator / stash rows)
  cudaStream_t bc_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool fork_bc = false;  // decided in set_points: the collocation grid leaves CTA slots free
  long long umma_clk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool use_umma = false;         // PINN_B200_KERNEL=umma and the configuration is supported: collocation term on tcgen05
  void* d_wimg = nullptr;        // bf16x3 weight-image stream of the tcgen05 production family (kind 3)
  float* d_uimg = nullptr;       // pre-split operand images of the experimental tcgen05 family C
  long long* d_uclk = nullptr;  // phase clocks of the last tcgen05 launch (experimental family)
  pinn_spec_t spec{};
  std::vector<int32_t> ops, aux_ops;
  std::vector<float> consts;
  PinnNet net{};
  FlatMap fmap{};
  int n_info = 0, n_slots = 0;
  const JetKernelInfo* kcol = nullptr;
  const JetKernelInfo* kbc = nullptr;
  PinnProgram prog_col{}, prog_bc{}, prog_aux{};

  // device state
  float *d_params = nullptr, *d_fused = nullptr, *d_m = nullptr, *d_v = nullptr, *d_wpack = nullptr;
  float *d_stash = nullptr, *d_gacc = nullptr, *d_seg_scale = nullptr, *d_lr = nullptr, *d_adam_c = nullptr;
  double *d_loss_part = nullptr, *d_ring = nullptr;
  int *d_ring_pos = nullptr, *d_adam_count = nullptr;
  LossMeta* d_meta = nullptr;
  int ring_cap = 4096;
  size_t stash_floats = 0, gacc_floats = 0;
  int grid_max = 0, grid_max_col = 0, grid_max_bc = 0;

  // points
  PointSet col, bc;
  int64_t n_col = 0;
  std::vector<int64_t> n_bd;
  int64_t n_bd_total = 0;
  int64_t n_col_global = 0;
  std::vector<int64_t> n_bd_global;
  double lw = 1.0, lref = 1.0;
  bool points_set = false;

  // launches
  PinnLaunch Lcol{}, Lbc{};
  int grid_col = 0, grid_bc = 0;

  // graph
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_valid = false;
  bool fused_tail = true;        // Adam step = evaluation kernels + ONE tail kernel (PINN_B200_FUSED_TAIL=0: separate kernels)
  unsigned* d_ticket = nullptr;  // last-block ticket of the fused tail
  int tc_share = 1;              // tcgen05 family, padded width 256: CTAs per (token-ordered) gradient-accumulator row
  unsigned* d_tokens = nullptr;  // [rows][PINN_TOKENS] flush-order tokens of the shared rows
  double cur_lr = -1.0;

  // nccl
  void* comm = nullptr;
  int world = 1, rank = 0;

  // timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;

  // reusable scratch of pinn_engine_eval (grown on demand, freed with the handle)
  std::vector<std::pair<void*, size_t>> eval_bufs;
  size_t l2_carve = 0;  // this engine's share of the device-wide persisting-L2 carve-out

  // device-resident L-BFGS loop: controller state, second scalar pair, optional trace of the trial points, and the
  // instantiated WHILE graph (rebuilt when the point set / communicator / stream changes)
  LbfgsCtl* d_ctl = nullptr;
  double* d_scal2 = nullptr;
  float* d_trace = nullptr;
  int trace_cap = 0, trace_rows = 0;
  cudaGraphExec_t lb_exec = nullptr;
  bool lb_valid = false;
  int lbfgs_syncs = 0;  // host synchronisations of the last pinn_engine_lbfgs call

  // lbfgs buffers
  float *d_x = nullptr, *d_g = nullptr, *d_d = nullptr, *d_xt = nullptr, *d_S = nullptr, *d_Y = nullptr;
  double *d_rho = nullptr, *d_alpha = nullptr, *d_scal = nullptr;
};

static void apply_l2_policy(pinn_engine* h);
// replicas of the tcgen05 family's weight-image stream (all CTAs walk the same sequence at the same time: replicas
// spread the reads over more L2 slices); PINN_TC_COPIES overrides the default for experiments
static int tc_image_copies() {
  static const int n = [] {
    const char* e = getenv("PINN_TC_COPIES");
    const int v = e ? atoi(e) : PINN_TC_IMAGE_COPIES;
    return v < 1 ? 1 : (v > 64 ? 64 : v);
  }();
  return n;
}
static void l2_release(pinn_engine* h);

static int pad_width(int w) {
  if (w <= 32) return 32;
  if (w <= 64) return 64;
  if (w <= 128) return 128;
  if (w <= 256) return 256;
  return -1;
}

static int build_layout(pinn_engine* h, int ldw) {
  const pinn_spec_t& s = h->spec;
  PinnNet& n = h->net;
  const int WP = pad_width(s.width);
  if (WP < 0) return fail("width %d > 256 not supported", s.width);
  if (s.n_hidden < 1 || s.n_hidden > PINN_MAX_LAYERS) return fail("n_hidden must be in 1..%d", PINN_MAX_LAYERS);
  if (s.d_in < 1 || s.d_in > 3) return fail("d_in must be 1..3");
  if (s.feat_mode == PINN_FEAT_POLAR && s.d_in != 2) return fail("polar feature map needs d_in == 2");
  n.d_in = s.d_in;
  n.feat_mode = s.feat_mode;
  n.n_feat = (s.feat_mode == PINN_FEAT_POLAR) ? 3 : s.d_in;
  n.n_hidden = s.n_hidden;
  n.width = s.width;
  n.wp = WP;
  n.act_first = s.act_first;
  n.act_hidden = s.act_hidden;
  n.scl = s.scl;
  n.epsil = s.epsil;
  for (int i = 0; i < 3; ++i) { n.lap_beta[i] = s.lap_beta[i]; n.lap_aux[i] = (s.mix == 2) ? s.lap_aux[i] : -1; }
  for (int i = 0; i < 3; ++i) {
    const double lb = s.lb[i], ub = s.ub[i];
    if (i < s.d_in && ub != lb) {
      n.fa[i] = (float)(2.0 / (ub - lb));
      n.fb[i] = (float)(-2.0 * lb / (ub - lb) - 1.0);
    } else {
      n.fa[i] = 0.f;
      n.fb[i] = 0.f;
    }
  }
  int o = 0;
  n.off_w0 = o; o += 4 * WP;
  n.off_b0 = o; o += WP;
  n.off_b[0] = n.off_b0;
  for (int l = 1; l < s.n_hidden; ++l) {
    n.off_w[l] = o; o += WP * ldw;
    n.off_b[l] = o; o += WP;
  }
  n.off_wl = o; o += WP;
  n.off_bl = o; o += 4;
  n.pg = o;
  for (int l = 1; l < s.n_hidden; ++l) {
    n.off_wt[l] = o; o += WP * ldw;
  }
  n.pw = o;

  FlatMap& M = h->fmap;
  M.n_layers = s.n_hidden + 1;
  M.wp = ldw;  // row stride of the transposed copy
  int f = 0;
  for (int l = 0; l <= s.n_hidden; ++l) {
    M.in_dim[l] = (l == 0) ? n.n_feat : s.width;
    M.out_dim[l] = (l == s.n_hidden) ? 1 : s.width;
    M.f_w[l] = f;
    f += M.in_dim[l] * M.out_dim[l] + M.out_dim[l];
    if (l == 0) { M.p_w[l] = n.off_w0; M.p_b[l] = n.off_b0; M.ld[l] = WP; M.p_wt[l] = -1; }
    else if (l == s.n_hidden) { M.p_w[l] = n.off_wl; M.p_b[l] = n.off_bl; M.ld[l] = 1; M.p_wt[l] = -1; }
    else { M.p_w[l] = n.off_w[l]; M.p_b[l] = n.off_b[l]; M.ld[l] = ldw; M.p_wt[l] = n.off_wt[l]; }
  }
  M.n_params = f;
  return 0;
}

extern "C" void pinn_engine_destroy(pinn_engine_t* h);

// everything after the handle exists; on failure the caller destroys the partially built handle
static int create_impl(pinn_engine* h, const pinn_spec_t* spec, int device) {
  h->device = device;
  h->spec = *spec;
  h->ops.assign(spec->ops, spec->ops + spec->n_ops);
  h->consts.assign(spec->consts, spec->consts + spec->n_consts);
  h->spec.ops = h->ops.data();
  h->spec.consts = h->consts.data();
  if (spec->n_aux_ops < 0 || spec->n_aux_ops > PINN_MAX_OPS) { return fail("aux program too long"); }
  if (spec->n_aux_user < 0 || spec->n_aux_user > spec->n_aux_col) { return fail("n_aux_user out of range"); }
  if (spec->n_aux_ops == 0 && spec->n_aux_user != spec->n_aux_col) { return fail("hoisted columns need an aux program"); }
  if (spec->n_aux_ops > 0) h->aux_ops.assign(spec->aux_ops, spec->aux_ops + spec->n_aux_ops);
  h->spec.aux_ops = h->aux_ops.data();
  {
    // kernel family: PINN_B200_KERNEL = simt | mma | tc | auto.  auto: the tcgen05 bf16x3 kernel (kind 3) for the
    // collocation term when it is instantiated for this width and jet structure (padded widths 128 / 256, at
    // least two hidden layers), else the 3xTF32 mma.sync kernel, else the fp32 SIMT kernel.  The boundary
    // term (value-only jets, ~1 % of the points) stays on the mma.sync / SIMT kernel.
    const int wp = pad_width(spec->width);
    if (wp < 0) { return fail("width %d > 256 not supported", spec->width); }
    const char* env = getenv("PINN_B200_KERNEL");
    const std::string want = env ? env : "auto";
    const JetKernelInfo *c1 = pinn_find_kernel(wp, spec->n1, spec->n2, spec->mix, 1), *b1 = pinn_find_kernel(wp, 0, 0, 0, 1);
    const JetKernelInfo *c0 = pinn_find_kernel(wp, spec->n1, spec->n2, spec->mix, 0), *b0 = pinn_find_kernel(wp, 0, 0, 0, 0);
    const JetKernelInfo* c3 = (spec->n_hidden >= 2) ? pinn_find_kernel(wp, spec->n1, spec->n2, spec->mix, 3) : nullptr;
    if (want == "tc") { h->kcol = c3; h->kbc = b1; }
    else if (want == "mma" || want == "umma") { h->kcol = c1; h->kbc = b1; }
    else if (want == "simt") { h->kcol = c0; h->kbc = b0; }
    else if (c3 && b1) { h->kcol = c3; h->kbc = b1; }
    else if (c1 && b1) { h->kcol = c1; h->kbc = b1; }
    else { h->kcol = c0; h->kbc = b0; }
    if (want == "umma" && (!h->kcol || !h->kbc))
      return fail("PINN_B200_KERNEL=umma: the experimental tcgen05 family C supports padded width 64 with jets (value, 2 first, combined second order), 2..4 hidden layers");
    if (!h->kcol || !h->kbc) {
      return fail("no %s kernel instantiation for WP=%d jets (n1=%d,n2=%d,mix=%d)", want.c_str(), wp, spec->n1, spec->n2, spec->mix);
    }
  }
  if (build_layout(h, h->kcol->ldw)) { return 1; }
  int occ_col = 1, occ_bc = 1;
  cudaError_t e = h->kcol->prepare(&occ_col);
  if (e == cudaSuccess) e = h->kbc->prepare(&occ_bc);
  if (e != cudaSuccess) { return fail("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
  // (cudaDeviceGetAttribute, not cudaGetDeviceProperties: the latter takes tens of milliseconds per call, and the
  // reference's call site creates two engines per training run)
  CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
  h->grid_max_col = h->num_sms * occ_col;
  if (const char* g = getenv("PINN_TC_GRID")) {  // experiment knob: cap the collocation grid of the tcgen05 family
    if (h->kcol->kind == 3 && atoi(g) > 0) h->grid_max_col = std::min(h->grid_max_col, atoi(g));
  }
  h->grid_max_bc = h->num_sms * occ_bc;
  // tcgen05 family: one row of (stash, gradient accumulators) per SM only -- the boundary term (~1 % of the
  // points) gets one CTA per SM too, so that the whole per-CTA scratch (C4: 148 x 0.85 MB) stays a window the
  // 126 MB L2 can hold instead of being diluted over rows the collocation kernel never touches
  if (h->kcol->kind == 3) h->grid_max_bc = h->num_sms;
  h->grid_max = std::max(h->grid_max_col, h->grid_max_bc);
  {
    const char* kenv = getenv("PINN_B200_KERNEL");
    if (kenv && !strcmp(kenv, "umma")) {
      if (!jet_umma_supported(h->net, h->kcol->k, spec->n1, spec->n2, spec->mix)) {
        return fail("PINN_B200_KERNEL=umma: the tcgen05 family supports padded width 64 with jets (value, 2 first, combined second order), 2..4 hidden layers");
      }
      h->use_umma = true;
    }
  }
  h->n_slots = spec->n_bc + 1;
  h->n_info = 3 + h->n_slots;
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));

  // programs
  h->prog_col.n_ops = spec->n_ops;
  memcpy(h->prog_col.ops, spec->ops, sizeof(int32_t) * spec->n_ops);
  memcpy(h->prog_col.consts, spec->consts, sizeof(float) * spec->n_consts);
  h->prog_aux.n_ops = spec->n_aux_ops;
  if (spec->n_aux_ops > 0) memcpy(h->prog_aux.ops, spec->aux_ops, sizeof(int32_t) * spec->n_aux_ops);
  memcpy(h->prog_aux.consts, spec->consts, sizeof(float) * spec->n_consts);
  h->prog_bc.n_ops = 3;  // u - aux0   (software.py:344)
  h->prog_bc.ops[0] = OP_JET | (0 << 8);
  h->prog_bc.ops[1] = OP_AUX | (0 << 8);
  h->prog_bc.ops[2] = OP_SUB;

  const int P = h->fmap.n_params;
  const size_t nf = (size_t)P + 2 * h->n_slots;
  CK(cudaMalloc(&h->d_params, sizeof(float) * P));
  CK(cudaMalloc(&h->d_fused, sizeof(float) * nf));
  CK(cudaMalloc(&h->d_m, sizeof(float) * P));
  CK(cudaMalloc(&h->d_v, sizeof(float) * P));
  CK(cudaMalloc(&h->d_wpack, sizeof(float) * h->net.pw));
  CK(cudaMemset(h->d_wpack, 0, sizeof(float) * h->net.pw));
  CK(cudaMemset(h->d_params, 0, sizeof(float) * P));
  CK(cudaMemset(h->d_m, 0, sizeof(float) * P));
  CK(cudaMemset(h->d_v, 0, sizeof(float) * P));
  h->stash_floats = ((size_t)h->grid_max * spec->n_hidden *
                     std::max(h->kcol->stash_floats_per_layer, h->kbc->stash_floats_per_layer) + 31) / 32 * 32;
  // stash and gradient accumulators share one allocation so that one persisting L2 window covers both
  h->gacc_floats = ((size_t)h->grid_max * h->net.pg + 31) / 32 * 32;
  CK(cudaMalloc(&h->d_stash, sizeof(float) * (h->stash_floats + h->gacc_floats)));
  h->d_gacc = h->d_stash + h->stash_floats;
  CK(cudaMalloc(&h->d_loss_part, sizeof(double) * (size_t)h->grid_max * h->n_slots));
  CK(cudaStreamCreateWithFlags(&h->bc_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CK(cudaMalloc(&h->d_seg_scale, sizeof(float) * PINN_MAX_SEG));
  CK(cudaMalloc(&h->d_lr, sizeof(float)));
  CK(cudaMalloc(&h->d_adam_c, sizeof(float) * 2));
  CK(cudaMalloc(&h->d_ring, sizeof(double) * (size_t)h->ring_cap * h->n_info));
  CK(cudaMalloc(&h->d_ring_pos, sizeof(int)));
  CK(cudaMalloc(&h->d_adam_count, sizeof(int)));
  CK(cudaMalloc(&h->d_meta, sizeof(LossMeta)));
  CK(cudaMemset(h->d_ring_pos, 0, sizeof(int)));
  CK(cudaMemset(h->d_adam_count, 0, sizeof(int)));
  CK(cudaMalloc(&h->d_ticket, sizeof(unsigned)));
  CK(cudaMemset(h->d_ticket, 0, sizeof(unsigned)));
  if (const char* ft = getenv("PINN_B200_FUSED_TAIL")) h->fused_tail = atoi(ft) != 0;
  if (h->kcol->kind == 3) CK(cudaMalloc(&h->d_wimg, jet_tc_image_bytes(h->net) * tc_image_copies()));
  if (h->kcol->kind == 3) {
    // Padded width 256: 148 private accumulator rows of 1.3 MB (194 MB) cannot stay in the L2 and every weight-gradient
    // flush became a DRAM read-modify-write; four CTAs per row (48 MB, inside the persisting carve-out) with a
    // token-ordered, bit-reproducible flush (jet_tc_kernel.cuh).  PINN_TC_SHARE overrides (1 = private rows).
    h->tc_share = (h->net.wp == 256) ? 4 : 1;
    if (const char* e = getenv("PINN_TC_SHARE")) h->tc_share = (h->net.wp == 256) ? std::max(1, std::min(8, atoi(e))) : 1;
    int coop = 0;   // the members of a row wait for each other: only with a cooperative launch (all CTAs resident)
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess || !coop) { cudaGetLastError(); h->tc_share = 1; }
    if (h->tc_share > 1) {
      const size_t rows = (size_t)(h->grid_max_col + h->tc_share - 1) / h->tc_share;
      CK(cudaMalloc(&h->d_tokens, sizeof(unsigned) * rows * PINN_TOKENS));
      CK(cudaMemset(h->d_tokens, 0, sizeof(unsigned) * rows * PINN_TOKENS));
    }
  }
  if (h->use_umma) {
    CK(cudaMalloc(&h->d_uimg, sizeof(float) * jet_umma_image_floats(h->net)));
    CK(cudaMalloc(&h->d_uclk, sizeof(long long) * 8));
    CK(cudaMemset(h->d_uclk, 0, sizeof(long long) * 8));
  }
  if (!getenv("PINN_B200_NO_L2_WINDOW")) apply_l2_policy(h);
  return 0;
}

extern "C" int pinn_engine_create(const pinn_spec_t* spec, int device, pinn_engine_t** out) {
  if (!spec || !out) return fail("null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: the B200 engine has no CPU fallback");
  if (device < 0 || device >= ndev) return fail("device %d out of range (%d devices)", device, ndev);
  CK(cudaSetDevice(device));
  if (spec->n_ops <= 0 || spec->n_ops > PINN_MAX_OPS) return fail("residual program length %d not in 1..%d", spec->n_ops, PINN_MAX_OPS);
  if (spec->n_consts < 0 || spec->n_consts > PINN_MAX_CONSTS) return fail("too many constants");
  if (spec->n_bc < 0 || spec->n_bc > PINN_MAX_SEG - 1) return fail("n_bc must be 0..%d", PINN_MAX_SEG - 1);
  pinn_engine* h = new pinn_engine();
  if (create_impl(h, spec, device)) {
    const std::string err = g_err;  // destroy may touch the error slot
    pinn_engine_destroy(h);
    g_err = err;
    return 1;
  }
  *out = h;
  return 0;
}

static void free_set(PointSet& s) {
  if (s.own_coords && s.coords) cudaFree(s.coords);
  if (s.own_aux && s.aux) cudaFree(s.aux);
  if (s.own_base && s.base) cudaFree(s.base);
  s = PointSet();
}

extern "C" void pinn_engine_destroy(pinn_engine_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  if (h->lb_exec) cudaGraphExecDestroy(h->lb_exec);
  if (h->d_ctl) cudaFree(h->d_ctl);
  if (h->d_scal2) cudaFree(h->d_scal2);
  if (h->d_trace) cudaFree(h->d_trace);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  void* bufs[] = {h->d_params, h->d_fused, h->d_m, h->d_v, h->d_wpack, h->d_stash, h->d_seg_scale,
                  h->d_lr, h->d_adam_c, h->d_loss_part, h->d_ring, h->d_ring_pos, h->d_adam_count, h->d_meta,
                  h->d_x, h->d_g, h->d_d, h->d_xt, h->d_S, h->d_Y, h->d_rho, h->d_alpha, h->d_scal, h->d_uimg, h->d_uclk, h->d_wimg, h->d_ticket, h->d_tokens};
  for (void* b : bufs)
    if (b) cudaFree(b);
  for (auto& b : h->eval_bufs)
    if (b.first) cudaFree(b.first);
  l2_release(h);
  free_set(h->col);
  free_set(h->bc);
  if (h->bc_stream) cudaStreamDestroy(h->bc_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->stg_col) cudaFree(h->stg_col);
  if (h->stg_bc) cudaFree(h->stg_bc);
  if (h->stg_ubc) cudaFree(h->stg_ubc);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->ev_staged) cudaEventDestroy(h->ev_staged);
  if (h->ev_consumed) cudaEventDestroy(h->ev_consumed);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// Keep the per-CTA stash (rewritten every tile, read back once) resident in L2: a persisting
// access-policy window on the engine stream, so dirty stash lines are overwritten in place
// instead of being evicted to HBM.
// cudaLimitPersistingL2CacheSize is DEVICE-global: several live engines (stage-2 model next to stage 1,
// concurrent sessions) must not shrink each other's carve-out, so the limit is the maximum over the
// live engines of the device.
#include <mutex>
static std::mutex g_l2_mu;
static std::vector<std::pair<int, pinn_engine*>> g_l2_live;  // (device, engine) with l2_carve > 0
static void l2_set_limit_locked(int device) {
  size_t want = 0;
  for (auto& e : g_l2_live)
    if (e.first == device) want = std::max(want, e.second->l2_carve);
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) cudaGetLastError();
}
static void l2_release(pinn_engine* h) {
  std::lock_guard<std::mutex> lk(g_l2_mu);
  bool had = false;
  for (size_t i = 0; i < g_l2_live.size(); ++i)
    if (g_l2_live[i].second == h) { g_l2_live.erase(g_l2_live.begin() + i); had = true; break; }
  if (had) l2_set_limit_locked(h->device);
}
static void apply_l2_policy(pinn_engine* h) {
  struct { int persistingL2CacheMaxSize = 0, accessPolicyMaxWindowSize = 0, l2CacheSize = 0; } prop;
  if (cudaDeviceGetAttribute(&prop.persistingL2CacheMaxSize, cudaDevAttrMaxPersistingL2CacheSize, h->device) != cudaSuccess ||
      cudaDeviceGetAttribute(&prop.accessPolicyMaxWindowSize, cudaDevAttrMaxAccessPolicyWindowSize, h->device) != cudaSuccess ||
      cudaDeviceGetAttribute(&prop.l2CacheSize, cudaDevAttrL2CacheSize, h->device) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
  // what the window covers: the whole per-CTA scratch [stash | gradient accumulators] (default), or only one of
  // the two regions (PINN_B200_L2_WINDOW = both | gacc | stash; experiment knob for scratch sets near the L2 size)
  // Default: both regions, except for the tcgen05 family, whose scratch (C4: 148 x 0.85 MB) exceeds the persisting
  // carve-out: there the gradient accumulators (read-modify-written every tile) persist and the stash competes for
  // the rest of the L2 (measured on C4: gacc 21.2 ms, stash 21.6, both 22.2, no window 22.1).
  const char* wsel = getenv("PINN_B200_L2_WINDOW");
  const bool tcfam = h->kcol && h->kcol->kind == 3;
  const bool only_gacc = wsel ? !strcmp(wsel, "gacc") : tcfam, only_stash = wsel && !strcmp(wsel, "stash");
  char* base = reinterpret_cast<char*>(only_gacc ? h->d_gacc : h->d_stash);
  size_t bytes = (only_gacc ? h->gacc_floats : only_stash ? h->stash_floats : h->stash_floats + h->gacc_floats) * sizeof(float);
  if (only_gacc && tcfam && h->tc_share > 1)   // shared rows: the collocation kernel only touches the first grid / share rows
    bytes = (size_t)((h->grid_max_col + h->tc_share - 1) / h->tc_share) * h->net.pg * sizeof(float);
  const size_t carve = std::min<size_t>(bytes, (size_t)prop.persistingL2CacheMaxSize);
  {
    std::lock_guard<std::mutex> lk(g_l2_mu);
    h->l2_carve = carve;
    bool present = false;
    for (auto& e : g_l2_live) present |= (e.second == h);
    if (!present) g_l2_live.emplace_back(h->device, h);
    l2_set_limit_locked(h->device);
  }
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof attr);
  attr.accessPolicyWindow.base_ptr = base;
  attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
  attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)attr.accessPolicyWindow.num_bytes);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  // lines of the window that do not get the persisting property stay NORMAL (not streaming / evict-first): for a
  // scratch just above the carve-out an evict-first remainder thrashes (measured on C4: 35.2 ms vs 31.2 ms)
  attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
  if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
  if (getenv("PINN_B200_DEBUG"))
    fprintf(stderr, "[pinn] L2 window: %.1f MB of %.1f MB (persisting max %.1f MB, window max %.1f MB, L2 %.1f MB)\n",
            carve / 1e6, bytes / 1e6, prop.persistingL2CacheMaxSize / 1e6, prop.accessPolicyMaxWindowSize / 1e6, prop.l2CacheSize / 1e6);
}

extern "C" int pinn_engine_set_stream(pinn_engine_t* h, void* s) {
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)s;
  h->own_stream = false;
  h->graph_valid = false; h->lb_valid = false;
  if (!getenv("PINN_B200_NO_L2_WINDOW")) apply_l2_policy(h);
  return 0;
}

extern "C" int pinn_engine_sync(pinn_engine_t* h) {
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int64_t pinn_engine_num_params(pinn_engine_t* h) { return h->fmap.n_params; }
extern "C" int32_t pinn_engine_num_loss_info(pinn_engine_t* h) { return h->n_info; }
extern "C" int32_t pinn_engine_tile_points(pinn_engine_t* h) { return h->kcol->tile_points; }
// kernels one evaluation enqueues (memsets and the NCCL allreduce not counted): pack, [weight images], [boundary
// kernel], collocation kernel, gradient reduce, loss reduce, loss_info; an Adam step adds k_adam
extern "C" int32_t pinn_engine_launches_per_eval(pinn_engine_t* h) {
  return 5 + (h->kcol->kind == 3 ? 1 : 0) + (h->use_umma ? 1 : 0) + ((h->points_set && h->Lbc.n_tiles > 0) ? 1 : 0);
}
// kernels one Adam step enqueues: with the fused tail [weight images], [boundary kernel], collocation kernel, tail
// (+ gradient reduce and loss reduce in front of an NCCL allreduce); otherwise one evaluation + k_adam
extern "C" int32_t pinn_engine_launches_per_adam_step(pinn_engine_t* h) {
  if (!h->fused_tail) return pinn_engine_launches_per_eval(h) + 1;
  return 2 + (h->comm ? 2 : 0) + (h->kcol->kind == 3 ? 1 : 0) + (h->use_umma ? 1 : 0) + ((h->points_set && h->Lbc.n_tiles > 0) ? 1 : 0);
}
extern "C" int32_t pinn_engine_kernel_kind(pinn_engine_t* h) { return h->use_umma ? 2 : h->kcol->kind; }

extern "C" int pinn_engine_set_params(pinn_engine_t* h, const float* flat, int on_device) {
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_params, flat, sizeof(float) * h->fmap.n_params,
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  if (!on_device) CK(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int pinn_engine_get_params(pinn_engine_t* h, float* out, int on_device) {
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(out, h->d_params, sizeof(float) * h->fmap.n_params,
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  if (!on_device) CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int ensure(float** p, size_t* cap, bool* own, size_t need) {
  if (*own && *p && *cap >= need) return 0;
  if (*own && *p) cudaFree(*p);
  *p = nullptr;
  const size_t c = std::max<size_t>(need, 16);
  CK(cudaMalloc(p, sizeof(float) * c));
  *cap = c;
  *own = true;
  return 0;
}

static int upload_meta(pinn_engine* h) {
  LossMeta M{};
  M.n_slots = h->n_slots;
  float sc[PINN_MAX_SEG] = {0};
  for (int i = 0; i < h->spec.n_bc; ++i) {
    const double N = (double)(h->n_bd_global.empty() ? h->n_bd[i] : h->n_bd_global[i]);
    M.count[i] = N > 0 ? N : 1.0;
    M.weight[i] = 1.0;
    sc[i] = (float)(2.0 / (M.count[i] * h->lref));
  }
  const int e = h->n_slots - 1;
  const double Nc = (double)(h->n_col_global > 0 ? h->n_col_global : h->n_col);
  M.count[e] = Nc > 0 ? Nc : 1.0;
  M.weight[e] = h->lw;
  sc[e] = (float)(2.0 * h->lw / (M.count[e] * h->lref));
  CK(cudaMemcpyAsync(h->d_meta, &M, sizeof M, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_seg_scale, sc, sizeof sc, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));  // host temporaries
  return 0;
}

static void fill_launch(pinn_engine* h, PinnLaunch& L, const JetKernelInfo* k, const PinnProgram& prog) {
  memset(&L, 0, sizeof L);
  L.net = h->net;
  L.wpack = h->d_wpack;
  L.seg_scale = h->d_seg_scale;
  L.stash = h->d_stash;
  L.gacc = h->d_gacc;
  L.loss_part = h->d_loss_part;
  L.n_slots = h->n_slots;
  L.prog = prog;
  L.wimg = h->d_wimg;
  L.wimg_copy_bytes = (long long)jet_tc_image_bytes(h->net);
  L.wimg_copies = tc_image_copies();
  L.ldw = h->kcol->ldw;
  L.gacc_share = (k == h->kcol) ? h->tc_share : 1;
  L.gacc_token = (k == h->kcol && h->tc_share > 1) ? h->d_tokens : nullptr;
}

// everything the fused kernels read besides the points: the padded fp32 pack and, for the tcgen05
// family, the bf16x3 weight-image stream
static int enqueue_pack(pinn_engine* h, const float* params_dev, cudaStream_t st, bool pack = true) {
  const int P = h->fmap.n_params;
  if (pack) {   // (the fused Adam tail re-packs the updated parameters itself)
    k_pack<<<(P + 255) / 256, 256, 0, st>>>(h->fmap, params_dev ? params_dev : h->d_params, h->d_wpack);
    CK(cudaGetLastError());
  }
  if (h->kcol->kind == 3) CK(jet_tc_build_images(h->d_wpack, h->net, h->kcol->ldw, h->d_wimg, tc_image_copies(), st));
  return 0;
}

extern "C" int pinn_engine_set_points(pinn_engine_t* h, const float* x_col, int64_t n_col, const float* aux_col,
                                      const float* base_col, int32_t n_bc, const float* const* x_bd,
                                      const float* const* u_bd, const float* const* base_bd,
                                      const int64_t* n_bd, int on_device) {
  CK(cudaSetDevice(h->device));
  if (n_bc != h->spec.n_bc) return fail("n_bc %d != spec.n_bc %d", n_bc, h->spec.n_bc);
  if (n_col <= 0) return fail("n_col must be > 0");
  const int d = h->spec.d_in, K = h->kcol->k;
  // the captured Adam graph bakes these six pointers by value: any change invalidates it
  const float* const before[6] = {h->col.coords, h->col.aux, h->col.base, h->bc.coords, h->bc.aux, h->bc.base};
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  // collocation set
  const int na = h->spec.n_aux_col, na_user = h->spec.n_aux_user;
  const bool has_aux_prog = h->spec.n_aux_ops > 0;
  if (na_user > 0 && !aux_col) return fail("aux_col required (n_aux_user=%d)", na_user);
  float* tmp_user = nullptr;  // device copy of host user aux when the aux program needs it
  if (on_device) {
    if (h->col.own_coords && h->col.coords) cudaFree(h->col.coords);
    h->col.coords = const_cast<float*>(x_col); h->col.own_coords = false; h->col.cap_coords = 0;
    if (!has_aux_prog) {
      if (h->col.own_aux && h->col.aux) cudaFree(h->col.aux);
      h->col.aux = const_cast<float*>(aux_col); h->col.own_aux = false; h->col.cap_aux = 0;
    }
    if (h->col.own_base && h->col.base) cudaFree(h->col.base);
    h->col.base = const_cast<float*>(base_col); h->col.own_base = false; h->col.cap_base = 0;
  } else {
    if (ensure(&h->col.coords, &h->col.cap_coords, &h->col.own_coords, (size_t)n_col * d)) return 1;
    CK(cudaMemcpyAsync(h->col.coords, x_col, sizeof(float) * n_col * d, kind, h->stream));
    if (!has_aux_prog) {
      if (aux_col) {
        if (ensure(&h->col.aux, &h->col.cap_aux, &h->col.own_aux, (size_t)n_col * na)) return 1;
        CK(cudaMemcpyAsync(h->col.aux, aux_col, sizeof(float) * n_col * na, kind, h->stream));
      } else if (!h->col.own_aux) h->col.aux = nullptr;
    }
    if (base_col) {
      if (ensure(&h->col.base, &h->col.cap_base, &h->col.own_base, (size_t)n_col * K)) return 1;
      CK(cudaMemcpyAsync(h->col.base, base_col, sizeof(float) * n_col * K, kind, h->stream));
    } else { if (h->col.own_base && h->col.base) cudaFree(h->col.base); h->col.base = nullptr; h->col.own_base = false; h->col.cap_base = 0; }
  }
  if (has_aux_prog) {
    // engine-owned combined aux buffer [n_col][na]: user columns copied, hoisted columns evaluated once
    if (ensure(&h->col.aux, &h->col.cap_aux, &h->col.own_aux, (size_t)n_col * na)) return 1;
    const float* user_dev = aux_col;
    if (na_user > 0 && !on_device) {
      CK(cudaMalloc(&tmp_user, sizeof(float) * n_col * na_user));
      CK(cudaMemcpyAsync(tmp_user, aux_col, sizeof(float) * n_col * na_user, cudaMemcpyHostToDevice, h->stream));
      user_dev = tmp_user;
    }
    k_eval_aux<<<(unsigned)((n_col + 255) / 256), 256, 0, h->stream>>>(h->prog_aux, h->col.coords, d, user_dev, na_user,
                                                                       h->col.aux, na, n_col);
    CK(cudaGetLastError());
    if (tmp_user) { CK(cudaStreamSynchronize(h->stream)); cudaFree(tmp_user); }
  }
  const bool shape_changed = (n_col != h->n_col);
  h->n_col = n_col;
  // boundary groups: concatenated copies
  int64_t tot = 0;
  std::vector<int64_t> nb(n_bc);
  for (int i = 0; i < n_bc; ++i) { nb[i] = n_bd[i]; tot += n_bd[i]; }
  bool bc_changed = (nb != h->n_bd);
  bool has_base_bd = false;
  for (int i = 0; i < n_bc; ++i) if (base_bd && base_bd[i]) has_base_bd = true;
  if (tot > 0) {
    if (ensure(&h->bc.coords, &h->bc.cap_coords, &h->bc.own_coords, (size_t)tot * d)) return 1;
    if (ensure(&h->bc.aux, &h->bc.cap_aux, &h->bc.own_aux, (size_t)tot)) return 1;
    if (has_base_bd) { if (ensure(&h->bc.base, &h->bc.cap_base, &h->bc.own_base, (size_t)tot)) return 1; }
    else if (h->bc.base) { cudaFree(h->bc.base); h->bc.base = nullptr; h->bc.own_base = false; h->bc.cap_base = 0; bc_changed = true; }
    int64_t o = 0;
    for (int i = 0; i < n_bc; ++i) {
      if (nb[i] == 0) continue;
      CK(cudaMemcpyAsync(h->bc.coords + o * d, x_bd[i], sizeof(float) * nb[i] * d, kind, h->stream));
      CK(cudaMemcpyAsync(h->bc.aux + o, u_bd[i], sizeof(float) * nb[i], kind, h->stream));
      if (has_base_bd) CK(cudaMemcpyAsync(h->bc.base + o, base_bd[i], sizeof(float) * nb[i], kind, h->stream));
      o += nb[i];
    }
  }
  h->n_bd = nb;
  h->n_bd_total = tot;

  // launch descriptors
  fill_launch(h, h->Lcol, h->kcol, h->prog_col);
  {
    PinnLaunch& L = h->Lcol;
    L.coords = h->col.coords; L.aux = h->col.aux; L.base = h->col.base; L.n_aux = h->spec.n_aux_col;
    L.n_seg = 1;
    const int tp = h->kcol->tile_points;
    L.n_tiles = (int)((n_col + tp - 1) / tp);
    L.seg_tile_end[0] = L.n_tiles; L.seg_pt_begin[0] = 0; L.seg_pt_end[0] = n_col; L.seg_slot[0] = h->n_slots - 1;
    h->grid_col = std::min(L.n_tiles, h->grid_max_col);
    if (h->use_umma) {  // tiles of 32 points, one CTA per SM
      L.n_tiles = (int)((n_col + 31) / 32);
      L.seg_tile_end[0] = L.n_tiles;
      h->grid_col = std::min(L.n_tiles, h->num_sms);
    }
  }
  fill_launch(h, h->Lbc, h->kbc, h->prog_bc);
  {
    PinnLaunch& L = h->Lbc;
    L.coords = h->bc.coords; L.aux = h->bc.aux; L.base = h->bc.base; L.n_aux = 1;
    const int tp = h->kbc->tile_points;
    int tiles = 0, ns = 0;
    int64_t o = 0;
    for (int i = 0; i < n_bc; ++i) {
      if (nb[i] > 0) {
        tiles += (int)((nb[i] + tp - 1) / tp);
        L.seg_tile_end[ns] = tiles; L.seg_pt_begin[ns] = o; L.seg_pt_end[ns] = o + nb[i]; L.seg_slot[ns] = i;
        ++ns;
      }
      o += nb[i];
    }
    L.n_seg = ns; L.n_tiles = tiles;
    h->grid_bc = std::min(tiles, h->grid_max_bc);
    // When the collocation kernel does not fill the machine (small problems, e.g. the reference's 5,200
    // points) the boundary kernel runs next to it on a forked stream and its CTAs own the rows after the
    // collocation CTAs'.  Otherwise the two persistent grids would only fight for the same CTA slots: they
    // run back to back and share rows (fewer rows to zero and reduce).
    h->fork_bc = tiles > 0 && h->grid_col + h->grid_bc <= h->grid_max;  // both grids fit next to each other (rows and CTA slots)
    // A FEW boundary tiles next to a collocation grid that fills every slot (the reference's stage 2: 10,400 + 400 points
    // = 326 + 13 tiles on 296 slots): the boundary kernel's single-tile latency (43 us) would be serial time in front of
    // a 190 us collocation kernel.  Give it its own CTAs instead -- the collocation grid shrinks by as many (it needs a
    // second round either way) -- and run the two side by side.  Not for the tcgen05 family (one 221 KB CTA per SM
    // leaves no room for a neighbour) and not for large boundary sets (measured slower on C2).
    if (!h->fork_bc && tiles > 0 && h->kcol->kind != 3 && h->grid_bc <= h->grid_max / 16 && h->grid_col > 2 * h->grid_bc &&
        !getenv("PINN_B200_NO_SMALL_FORK")) {
      h->grid_col = std::min(h->grid_col, h->grid_max - h->grid_bc);
      h->fork_bc = true;
    }
    if (h->fork_bc) {
      const size_t stash_row = (size_t)h->spec.n_hidden * std::max(h->kcol->stash_floats_per_layer, h->kbc->stash_floats_per_layer);
      L.stash = h->d_stash + (size_t)h->grid_col * stash_row;
      L.gacc = h->d_gacc + (size_t)h->grid_col * h->net.pg;
      L.loss_part = h->d_loss_part + (size_t)h->grid_col * h->n_slots;
    }
  }
  if (!on_device) CK(cudaStreamSynchronize(h->stream));  // host buffers may be reused by the caller
  h->points_set = true;
  const float* const after[6] = {h->col.coords, h->col.aux, h->col.base, h->bc.coords, h->bc.aux, h->bc.base};
  bool ptr_changed = false;
  for (int i = 0; i < 6; ++i) ptr_changed |= (before[i] != after[i]);
  if (shape_changed || bc_changed || on_device || ptr_changed) { h->graph_valid = false; h->lb_valid = false; }
  // global counts (multi-GPU means) survive a resample with unchanged local shapes; a shape change resets them
  if (shape_changed || bc_changed) {
    h->n_col_global = 0;
    h->n_bd_global.clear();
  }
  return upload_meta(h);
}

// Pipelined refresh of the point set (same shapes as the current one): prefetch copies the NEXT
// step's host buffers into staging memory on a separate copy stream -- it overlaps the step the
// engine stream is computing -- and commit swaps them in (device-to-device, then the hoisted aux
// program) on the engine stream.  Neither call blocks the host.
extern "C" int pinn_engine_prefetch_points(pinn_engine_t* h, const float* x_col, int64_t n_col, int32_t n_bc,
                                           const float* const* x_bd, const float* const* u_bd, const int64_t* n_bd) {
  CK(cudaSetDevice(h->device));
  if (!h->points_set) return fail("prefetch_points: call set_points once first (it fixes the shapes)");
  if (n_col != h->n_col || n_bc != h->spec.n_bc) return fail("prefetch_points: shapes differ from the current point set");
  for (int i = 0; i < n_bc; ++i)
    if (n_bd[i] != h->n_bd[i]) return fail("prefetch_points: boundary group %d has %lld points, current set has %lld", i,
                                           (long long)n_bd[i], (long long)h->n_bd[i]);
  if (h->spec.n_aux_user > 0 || h->col.base || h->bc.base || !h->col.own_coords)
    return fail("prefetch_points: only for engine-owned point sets without user aux/base columns (use set_points)");
  const int d = h->spec.d_in;
  if (!h->copy_stream) {
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_staged, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_consumed, cudaEventDisableTiming));
  }
  if (h->stg_col_cap < (size_t)n_col * d) {
    if (h->stg_col) cudaFree(h->stg_col);
    CK(cudaMalloc(&h->stg_col, sizeof(float) * n_col * d));
    h->stg_col_cap = (size_t)n_col * d;
  }
  if (h->n_bd_total > 0 && h->stg_bc_cap < (size_t)h->n_bd_total) {
    if (h->stg_bc) cudaFree(h->stg_bc);
    if (h->stg_ubc) cudaFree(h->stg_ubc);
    CK(cudaMalloc(&h->stg_bc, sizeof(float) * h->n_bd_total * d));
    CK(cudaMalloc(&h->stg_ubc, sizeof(float) * h->n_bd_total));
    h->stg_bc_cap = (size_t)h->n_bd_total;
  }
  if (h->staged) return fail("prefetch_points: the previous prefetch has not been committed");
  CK(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed, 0));  // staging free again (no-op before the first commit)
  CK(cudaMemcpyAsync(h->stg_col, x_col, sizeof(float) * n_col * d, cudaMemcpyHostToDevice, h->copy_stream));
  int64_t o = 0;
  for (int i = 0; i < n_bc; ++i) {
    if (n_bd[i] == 0) continue;
    CK(cudaMemcpyAsync(h->stg_bc + o * d, x_bd[i], sizeof(float) * n_bd[i] * d, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(h->stg_ubc + o, u_bd[i], sizeof(float) * n_bd[i], cudaMemcpyHostToDevice, h->copy_stream));
    o += n_bd[i];
  }
  CK(cudaEventRecord(h->ev_staged, h->copy_stream));
  h->staged = true;
  return 0;
}

extern "C" int pinn_engine_commit_points(pinn_engine_t* h) {
  CK(cudaSetDevice(h->device));
  if (!h->staged) return fail("commit_points: nothing was prefetched");
  const int d = h->spec.d_in;
  CK(cudaStreamWaitEvent(h->stream, h->ev_staged, 0));
  CK(cudaMemcpyAsync(h->col.coords, h->stg_col, sizeof(float) * h->n_col * d, cudaMemcpyDeviceToDevice, h->stream));
  if (h->n_bd_total > 0) {
    CK(cudaMemcpyAsync(h->bc.coords, h->stg_bc, sizeof(float) * h->n_bd_total * d, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->bc.aux, h->stg_ubc, sizeof(float) * h->n_bd_total, cudaMemcpyDeviceToDevice, h->stream));
  }
  CK(cudaEventRecord(h->ev_consumed, h->stream));
  if (h->spec.n_aux_ops > 0) {
    k_eval_aux<<<(unsigned)((h->n_col + 255) / 256), 256, 0, h->stream>>>(h->prog_aux, h->col.coords, d, nullptr, 0, h->col.aux,
                                                                          h->spec.n_aux_col, h->n_col);
    CK(cudaGetLastError());
  }
  h->staged = false;
  return 0;
}

extern "C" int pinn_engine_set_global_counts(pinn_engine_t* h, int64_t n_col_global, const int64_t* n_bd_global) {
  CK(cudaSetDevice(h->device));
  h->n_col_global = n_col_global;
  h->n_bd_global.assign(n_bd_global, n_bd_global + h->spec.n_bc);
  return upload_meta(h);
}

extern "C" int pinn_engine_set_loss(pinn_engine_t* h, double lw_eqn, double lref) {
  CK(cudaSetDevice(h->device));
  h->lw = lw_eqn;
  h->lref = lref;
  if (!h->points_set) return 0;
  return upload_meta(h);
}

// enqueue one evaluation: pack -> zero -> bc -> col -> reduce -> (allreduce) -> loss_info
// adam_tail: the evaluation is the first half of an Adam step whose second half is the fused tail kernel (gradient
// reduction + loss_info + Adam + re-pack in ONE launch; the caller packs once before the first step of a sequence)
static int enqueue_eval(pinn_engine* h, const float* params_dev, int tick, bool adam_tail = false) {
  if (!h->points_set) return fail("set_points has not been called");
  cudaStream_t st = h->stream;
  const int P = h->fmap.n_params;
  const bool has_bc = h->Lbc.n_tiles > 0, fork = has_bc && h->fork_bc;
  const int nb = fork ? h->grid_col + h->grid_bc : std::max(h->grid_col, h->grid_bc);
  if (enqueue_pack(h, params_dev, st, !adam_tail)) return 1;
  CK(cudaMemsetAsync(h->d_gacc, 0, sizeof(float) * (size_t)nb * h->net.pg, st));
  CK(cudaMemsetAsync(h->d_loss_part, 0, sizeof(double) * (size_t)nb * h->n_slots, st));
  if (fork) {
    CK(cudaEventRecord(h->ev_fork, st));
    CK(cudaStreamWaitEvent(h->bc_stream, h->ev_fork, 0));
    CK(h->kbc->launch(h->Lbc, true, h->grid_bc, h->bc_stream));
    CK(cudaEventRecord(h->ev_join, h->bc_stream));
  } else if (has_bc) {
    CK(h->kbc->launch(h->Lbc, true, h->grid_bc, st));
  }
  if (h->use_umma) {
    CK(jet_umma_build_images(h->d_wpack, h->net, h->kcol->ldw, h->d_uimg, st));
    CK(jet_umma_train_launch(h->Lcol, h->d_uimg, h->kcol->ldw, h->grid_col, st, h->d_uclk));
  } else {
    CK(h->kcol->launch(h->Lcol, true, h->grid_col, st));
  }
  if (fork) CK(cudaStreamWaitEvent(st, h->ev_join, 0));  // join
  if (adam_tail && !h->comm) {
    k_adam_tail<true><<<(P + 127) / 128, 128, 0, st>>>(h->fmap, h->d_gacc, nb, h->net.pg, h->d_fused, h->d_loss_part, h->n_slots,
                                                       h->d_meta, h->d_ring, h->d_ring_pos, h->ring_cap, h->d_adam_count, h->d_adam_c,
                                                       h->d_params, h->d_m, h->d_v, h->d_lr, h->d_wpack, h->d_ticket);
    CK(cudaGetLastError());
    return 0;
  }
  k_grad_reduce<<<(P + 127) / 128, 128, 0, st>>>(h->fmap, h->d_gacc, nb, h->net.pg, h->d_fused);
  CK(cudaGetLastError());
  k_loss_reduce<<<1, 32 * h->n_slots, 0, st>>>(h->d_loss_part, nb, h->n_slots, h->d_fused + P);
  CK(cudaGetLastError());
  if (h->comm) {
    const int rc = g_nccl.AllReduce(h->d_fused, h->d_fused, (size_t)P + 2 * h->n_slots, /*ncclFloat32*/ 7,
                                    /*ncclSum*/ 0, h->comm, st);
    if (rc != 0) return fail("ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
  }
  if (adam_tail) {   // after the allreduce: loss_info + Adam + re-pack in one launch
    k_adam_tail<false><<<(P + 127) / 128, 128, 0, st>>>(h->fmap, h->d_gacc, nb, h->net.pg, h->d_fused, h->d_loss_part, h->n_slots,
                                                        h->d_meta, h->d_ring, h->d_ring_pos, h->ring_cap, h->d_adam_count, h->d_adam_c,
                                                        h->d_params, h->d_m, h->d_v, h->d_lr, h->d_wpack, h->d_ticket);
    CK(cudaGetLastError());
    return 0;
  }
  k_loss_info<<<1, 32, 0, st>>>(h->d_meta, h->d_fused + P, h->d_ring, h->d_ring_pos, h->ring_cap, tick,
                                h->d_adam_count, h->d_adam_c);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int pinn_engine_loss_grad(pinn_engine_t* h, const float* params_dev, float* grad_out_dev,
                                     double* loss_info_host) {
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(h->d_ring_pos, 0, sizeof(int), h->stream));
  CK(cudaEventRecord(h->ev0, h->stream));
  if (enqueue_eval(h, params_dev, 0)) return 1;
  CK(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  if (grad_out_dev)
    CK(cudaMemcpyAsync(grad_out_dev, h->d_fused, sizeof(float) * h->fmap.n_params, cudaMemcpyDeviceToDevice, h->stream));
  if (loss_info_host) {
    CK(cudaMemcpyAsync(loss_info_host, h->d_ring, sizeof(double) * h->n_info, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

extern "C" int pinn_engine_adam_init(pinn_engine_t* h) {
  CK(cudaSetDevice(h->device));
  const int P = h->fmap.n_params;
  CK(cudaMemsetAsync(h->d_m, 0, sizeof(float) * P, h->stream));
  CK(cudaMemsetAsync(h->d_v, 0, sizeof(float) * P, h->stream));
  CK(cudaMemsetAsync(h->d_adam_count, 0, sizeof(int), h->stream));
  return 0;
}

static int enqueue_adam_step(pinn_engine* h) {
  if (h->fused_tail) return enqueue_eval(h, nullptr, 1, true);
  if (enqueue_eval(h, nullptr, 1)) return 1;
  const int P = h->fmap.n_params;
  k_adam<<<(P + 255) / 256, 256, 0, h->stream>>>(P, h->d_params, h->d_fused, h->d_m, h->d_v, h->d_lr, h->d_adam_c);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int pinn_engine_adam_steps(pinn_engine_t* h, int32_t n_steps, double lr, double* rows) {
  CK(cudaSetDevice(h->device));
  if (n_steps <= 0) return 0;
  if (lr != h->cur_lr) {
    k_set_f32<<<1, 1, 0, h->stream>>>(h->d_lr, (float)lr);
    CK(cudaGetLastError());
    h->cur_lr = lr;
  }
  if (!h->graph_valid) {
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_adam_step(h);
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return 1; }
    if (e != cudaSuccess) return fail("graph capture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail("graph instantiate: %s", cudaGetErrorString(e));
    h->graph_valid = true;
  }
  CK(cudaEventRecord(h->ev0, h->stream));
  if (h->fused_tail) {   // the steps re-pack after their update; the first one needs the pack of the current parameters
    const int P = h->fmap.n_params;
    k_pack<<<(P + 255) / 256, 256, 0, h->stream>>>(h->fmap, h->d_params, h->d_wpack);
    CK(cudaGetLastError());
  }
  int done = 0;
  while (done < n_steps) {
    const int chunk = std::min(n_steps - done, h->ring_cap);
    CK(cudaMemsetAsync(h->d_ring_pos, 0, sizeof(int), h->stream));
    for (int i = 0; i < chunk; ++i) CK(cudaGraphLaunch(h->graph_exec, h->stream));
    if (rows) {
      CK(cudaMemcpyAsync(rows + (size_t)done * h->n_info, h->d_ring, sizeof(double) * (size_t)chunk * h->n_info,
                         cudaMemcpyDeviceToHost, h->stream));
      if (done + chunk < n_steps) CK(cudaStreamSynchronize(h->stream));
    }
    done += chunk;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  h->timed = true;
  if (rows) CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int pinn_engine_adam_rows(pinn_engine_t* h, int32_t n_rows, double* rows) {
  CK(cudaSetDevice(h->device));
  if (!rows || n_rows <= 0 || n_rows > h->ring_cap) return fail("adam_rows: n_rows must be 1..%d", h->ring_cap);
  CK(cudaMemcpyAsync(rows, h->d_ring, sizeof(double) * (size_t)n_rows * h->n_info, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" double pinn_engine_last_ms(pinn_engine_t* h) {
  if (!h->timed) return -1.0;
  cudaSetDevice(h->device);
  if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}

extern "C" int pinn_engine_eval(pinn_engine_t* h, const float* z, int64_t n, const float* aux, const float* base,
                                float* u_out, float* f_out, float* jets_out, int on_device) {
  CK(cudaSetDevice(h->device));
  if (n <= 0) return 0;
  const int d = h->spec.d_in, K = h->kcol->k, na = h->spec.n_aux_col, na_user = h->spec.n_aux_user;
  const bool has_aux_prog = h->spec.n_aux_ops > 0;
  if (na_user > 0 && !aux) return fail("aux required");
  cudaStream_t st = h->stream;
  // scratch comes from the handle (grown on demand, reused by later calls: predictF / resampling / stage-2
  // set_data call this every 100-2000 steps); reuse is ordered by the engine stream
  size_t n_tmp = 0;
  auto dalloc = [&](size_t bytes) -> void* {
    if (n_tmp == h->eval_bufs.size()) h->eval_bufs.emplace_back(nullptr, 0);
    auto& b = h->eval_bufs[n_tmp++];
    if (b.second < bytes) {
      if (b.first) { cudaStreamSynchronize(st); cudaFree(b.first); b.first = nullptr; b.second = 0; }
      if (cudaMalloc(&b.first, bytes) != cudaSuccess) { b.first = nullptr; return nullptr; }
      b.second = bytes;
    }
    return b.first;
  };
  auto cleanup = [&]() {};
  const float *dz = z, *daux = aux, *dbase = base;
  float *du = u_out, *df = f_out, *dj = jets_out;
  if (!on_device) {
    float* t = (float*)dalloc(sizeof(float) * n * d);
    if (!t) { cleanup(); return fail("cudaMalloc failed"); }
    cudaMemcpyAsync(t, z, sizeof(float) * n * d, cudaMemcpyHostToDevice, st); dz = t;
    if (aux && na_user > 0) { t = (float*)dalloc(sizeof(float) * n * na_user); if (!t) { cleanup(); return fail("cudaMalloc failed"); }
      cudaMemcpyAsync(t, aux, sizeof(float) * n * na_user, cudaMemcpyHostToDevice, st); daux = t; }
    if (base) { t = (float*)dalloc(sizeof(float) * n * K); if (!t) { cleanup(); return fail("cudaMalloc failed"); }
      cudaMemcpyAsync(t, base, sizeof(float) * n * K, cudaMemcpyHostToDevice, st); dbase = t; }
    if (u_out) { du = (float*)dalloc(sizeof(float) * n); if (!du) { cleanup(); return fail("cudaMalloc failed"); } }
    if (f_out) { df = (float*)dalloc(sizeof(float) * n); if (!df) { cleanup(); return fail("cudaMalloc failed"); } }
    if (jets_out) { dj = (float*)dalloc(sizeof(float) * n * K); if (!dj) { cleanup(); return fail("cudaMalloc failed"); } }
  }
  if (has_aux_prog) {
    float* comb = (float*)dalloc(sizeof(float) * n * na);
    if (!comb) { cleanup(); return fail("cudaMalloc failed"); }
    k_eval_aux<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->prog_aux, dz, d, daux, na_user, comb, na, n);
    daux = comb;
  }
  if (enqueue_pack(h, nullptr, st)) { cleanup(); return 1; }
  PinnLaunch L;
  fill_launch(h, L, h->kcol, h->prog_col);
  L.coords = dz; L.aux = daux; L.base = dbase; L.n_aux = na;
  L.out_u = du; L.out_f = df; L.out_jets = dj;
  L.n_seg = 1;
  const int tp = h->kcol->tile_points;
  L.n_tiles = (int)((n + tp - 1) / tp);
  L.seg_tile_end[0] = L.n_tiles; L.seg_pt_begin[0] = 0; L.seg_pt_end[0] = n; L.seg_slot[0] = 0;
  cudaError_t e;
  if (h->use_umma) {
    // experimental tcgen05 family C
    if (!jet_umma_supported(h->net, K, h->spec.n1, h->spec.n2, h->spec.mix)) { cleanup(); return fail("PINN_B200_KERNEL=umma: configuration not supported by the tcgen05 family"); }
    float* images = (float*)dalloc(sizeof(float) * jet_umma_image_floats(h->net));
    long long* clk = (long long*)dalloc(sizeof(long long) * 8);
    if (!images || !clk) { cleanup(); return fail("cudaMalloc failed"); }
    cudaMemsetAsync(clk, 0, sizeof(long long) * 8, st);
    L.n_tiles = (int)((n + 31) / 32);
    L.seg_tile_end[0] = L.n_tiles;
    e = jet_umma_build_images(h->d_wpack, h->net, h->kcol->ldw, images, st);
    if (e == cudaSuccess) e = jet_umma_eval_launch(L, images, 2 * h->num_sms, st, clk);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->umma_clk, clk, sizeof(long long) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  } else {
    e = h->kcol->launch(L, false, std::min(L.n_tiles, h->grid_max_col), st);
  }
  if (e != cudaSuccess) { cleanup(); return fail("eval launch: %s", cudaGetErrorString(e)); }
  if (!on_device) {
    if (u_out) cudaMemcpyAsync(u_out, du, sizeof(float) * n, cudaMemcpyDeviceToHost, st);
    if (f_out) cudaMemcpyAsync(f_out, df, sizeof(float) * n, cudaMemcpyDeviceToHost, st);
    if (jets_out) cudaMemcpyAsync(jets_out, dj, sizeof(float) * n * K, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess) return fail("eval: %s", cudaGetErrorString(e));
  }
  return 0;
}

extern "C" int pinn_engine_umma_clocks(pinn_engine_t* h, long long* out8) {
  if (h->use_umma) {  // clocks of the last training launch
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out8, h->d_uclk, sizeof(long long) * 8, cudaMemcpyDeviceToHost));
    return 0;
  }
  for (int i = 0; i < 8; ++i) out8[i] = h->umma_clk[i];
  return 0;
}

// Time the two fused kernels alone (roofline numerator): each launch bracketed by CUDA
// events on the engine stream, an L2 flush (memset of flush_bytes) between launches.
extern "C" int pinn_engine_time_kernels(pinn_engine_t* h, int32_t reps, int64_t flush_bytes, double* col_ms,
                                        double* bc_ms) {
  CK(cudaSetDevice(h->device));
  if (!h->points_set) return fail("set_points has not been called");
  cudaStream_t st = h->stream;
  const int P = h->fmap.n_params;
  const int nb = h->fork_bc ? h->grid_col + h->grid_bc : std::max(h->grid_col, h->grid_bc);
  void* flush = nullptr;
  if (flush_bytes > 0) CK(cudaMalloc(&flush, (size_t)flush_bytes));
  if (enqueue_pack(h, nullptr, st)) return 1;
  double tc = 0.0, tb = 0.0;
  for (int r = 0; r < reps; ++r) {
    CK(cudaMemsetAsync(h->d_gacc, 0, sizeof(float) * (size_t)nb * h->net.pg, st));
    CK(cudaMemsetAsync(h->d_loss_part, 0, sizeof(double) * (size_t)nb * h->n_slots, st));
    if (flush) CK(cudaMemsetAsync(flush, r & 0xff, (size_t)flush_bytes, st));
    float ms = 0.f;
    if (h->Lbc.n_tiles > 0) {
      CK(cudaEventRecord(h->ev0, st));
      CK(h->kbc->launch(h->Lbc, true, h->grid_bc, st));
      CK(cudaEventRecord(h->ev1, st));
      CK(cudaEventSynchronize(h->ev1));
      CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
      tb += ms;
    }
    if (flush) CK(cudaMemsetAsync(flush, (r + 1) & 0xff, (size_t)flush_bytes, st));
    CK(cudaEventRecord(h->ev0, st));
    CK(h->kcol->launch(h->Lcol, true, h->grid_col, st));
    CK(cudaEventRecord(h->ev1, st));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    tc += ms;
  }
  if (flush) cudaFree(flush);
  h->timed = false;
  if (col_ms) *col_ms = tc / reps;
  if (bc_ms) *bc_ms = tb / reps;
  return 0;
}

// Phase profile of the collocation kernel: clock64 totals of CTA 0 / thread 0 per phase
// (0 fwd GEMM, 1 activation fwd, 2 output+residual, 3 activation bwd, 4 smem restage, 5 wgrad,
//  6 dgrad, 7 rest).  Only the tensor-core kernel is instrumented.
extern "C" int pinn_engine_phase_profile(pinn_engine_t* h, int64_t* out8) {
  CK(cudaSetDevice(h->device));
  if (!h->points_set) return fail("set_points has not been called");
  cudaStream_t st = h->stream;
  const int P = h->fmap.n_params;
  long long* d = nullptr;
  CK(cudaMalloc(&d, 16 * sizeof(long long)));
  CK(cudaMemsetAsync(d, 0, 16 * sizeof(long long), st));
  if (enqueue_pack(h, nullptr, st)) return 1;
  CK(cudaMemsetAsync(h->d_gacc, 0, sizeof(float) * (size_t)h->grid_col * h->net.pg, st));
  CK(cudaMemsetAsync(h->d_loss_part, 0, sizeof(double) * (size_t)h->grid_col * h->n_slots, st));
  PinnLaunch L = h->Lcol;
  L.phase_clk = d;
  if (const char* ex = getenv("PINN_TC_EXP")) L.exp_flags = atoi(ex);
  CK(h->kcol->launch(L, true, h->grid_col, st));
  long long host[16];
  CK(cudaMemcpyAsync(host, d, sizeof host, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(d);
  if (getenv("PINN_B200_DEBUG"))  // tcgen05 family: GEMM durations seen by the issuing lane (fwd, dgrad incl. queueing, wgrad)
    fprintf(stderr, "[pinn] GEMM clocks: fwd %lld dgrad %lld wgrad %lld\n", host[8], host[9], host[10]);
  for (int i = 0; i < 8; ++i) out8[i] = host[i];
  return 0;
}

// ---------------------------------------------------------------- L-BFGS (software.py:499-514)
namespace {
struct Phi { double a, f, d; };  // step, value, directional derivative

struct LineSearch {
  // Hager & Zhang (2006) "Algorithm 851: CG_DESCENT" line search, the algorithm
  // tfp.optimizer.linesearch.hager_zhang implements (delta=.1, sigma=.9, eps=1e-6,
  // gamma=.66, rho=5, bisection theta=.5, max 50 evaluations).
  pinn_engine* h;
  int value_unnorm;
  pinn_eval_cb cb;
  void* user;
  int evals = 0, max_evals = 50;
  double f_lim = 0, phi0 = 0, dphi0 = 0;
  bool error = false;
  std::vector<double> info;

  int eval(double a, Phi& out);
  bool wolfe(const Phi& p) const {
    const double delta = 0.1, sigma = 0.9;
    if (!(isfinite(p.f) && isfinite(p.d))) return false;
    const bool exact = (p.f <= phi0 + delta * p.a * dphi0) && (p.d >= sigma * dphi0);
    const bool approx = (p.f <= f_lim) && ((2 * delta - 1) * dphi0 >= p.d) && (p.d >= sigma * dphi0);
    return exact || approx;
  }
};

int LineSearch::eval(double a, Phi& out) {
  pinn_engine* e = h;
  const int P = e->fmap.n_params;
  cudaStream_t st = e->stream;
  k_axpy_out<<<(P + 255) / 256, 256, 0, st>>>(P, e->d_x, e->d_d, a, e->d_xt);
  CK(cudaMemsetAsync(e->d_ring_pos, 0, sizeof(int), st));
  if (enqueue_eval(e, e->d_xt, 0)) return 1;
  k_dot_inf<<<1, 1024, 0, st>>>(P, e->d_fused, e->d_d, e->d_scal);
  CK(cudaGetLastError());
  double sc[2];
  info.resize(e->n_info);
  CK(cudaMemcpyAsync(info.data(), e->d_ring, sizeof(double) * e->n_info, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(sc, e->d_scal, sizeof sc, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ++evals;
  ++e->lbfgs_syncs;
  if (cb) cb(info.data(), e->n_info, user);
  out.a = a;
  out.f = value_unnorm ? info[0] : info[0] / e->lref;
  out.d = sc[0];
  if (!isfinite(out.f)) { out.f = INFINITY; out.d = -1.0; }
  return 0;
}
}  // namespace

static int lbfgs_alloc(pinn_engine* h) {
  if (h->d_x) return 0;
  const int P = h->fmap.n_params, m = 10;
  CK(cudaMalloc(&h->d_x, sizeof(float) * P));
  CK(cudaMalloc(&h->d_g, sizeof(float) * P));
  CK(cudaMalloc(&h->d_d, sizeof(float) * P));
  CK(cudaMalloc(&h->d_xt, sizeof(float) * P));
  CK(cudaMalloc(&h->d_S, sizeof(float) * (size_t)P * m));
  CK(cudaMalloc(&h->d_Y, sizeof(float) * (size_t)P * m));
  CK(cudaMalloc(&h->d_rho, sizeof(double) * m));
  CK(cudaMalloc(&h->d_alpha, sizeof(double) * lb_scratch_doubles(P)));   // scratch of the vector-free two-loop recursion
  CK(cudaMemsetAsync(h->d_S, 0, sizeof(float) * (size_t)P * m, h->stream));   // (its Gram pass reads every history slot)
  CK(cudaMemsetAsync(h->d_Y, 0, sizeof(float) * (size_t)P * m, h->stream));
  CK(cudaMalloc(&h->d_scal, sizeof(double) * 4));
  return 0;
}

// Round-1 implementation: the line search runs on the HOST, one stream synchronisation per evaluation.  Kept as
// PINN_B200_LBFGS=legacy: the regression reference the new loops are compared with bit for bit.
static int lbfgs_legacy(pinn_engine_t* h, int32_t max_iter, double tol, int32_t value_unnorm,
                        pinn_eval_cb cb, void* user, pinn_lbfgs_result_t* out) {
  const int P = h->fmap.n_params, m = 10;
  cudaStream_t st = h->stream;
  pinn_lbfgs_result_t R{};
  LineSearch ls{h, value_unnorm, cb, user};
  h->lbfgs_syncs = 0;
  CK(cudaMemcpyAsync(h->d_x, h->d_params, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(h->d_d, 0, sizeof(float) * P, st));
  // initial evaluation at x0 (tfp evaluates value_and_gradients at the initial position)
  Phi p0;
  if (ls.eval(0.0, p0)) return 1;
  CK(cudaMemcpyAsync(h->d_g, h->d_fused, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  k_dot_inf<<<1, 1024, 0, st>>>(P, h->d_g, h->d_g, h->d_scal);
  double sc[2];
  CK(cudaMemcpyAsync(sc, h->d_scal, sizeof sc, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double fcur = p0.f, ginf = sc[1];
  int total_evals = ls.evals;
  int cnt = 0, head = 0;
  R.converged = ginf <= tol;
  while (!R.converged && !R.failed && R.iterations < max_iter) {
    CK(lb_two_loop(P, nullptr, m, cnt, head, h->d_g, h->d_S, h->d_Y, h->d_rho, h->d_d, h->d_alpha, st));
    k_dot_inf<<<1, 1024, 0, st>>>(P, h->d_g, h->d_d, h->d_scal);
    CK(cudaMemcpyAsync(sc, h->d_scal, sizeof sc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const double dphi0 = sc[0];
    if (!(dphi0 < 0.0) || !isfinite(dphi0)) { R.failed = 1; break; }
    // ---- Hager-Zhang line search
    ls.evals = 0;
    ls.phi0 = fcur;
    ls.dphi0 = dphi0;
    ls.f_lim = fcur + 1e-6 * fabs(fcur);
    Phi lo{0.0, fcur, dphi0}, hi{}, c{};
    bool found = false, ok = true;
    auto ev = [&](double a, Phi& p) -> bool {  // returns true when p satisfies the Wolfe test
      if (ls.eval(a, p)) { ok = false; ls.error = true; return false; }
      return ls.wolfe(p);
    };
    // U3 bisection on [A, B] with phi'(A) < 0, phi(A) <= f_lim, phi'(B) < 0, phi(B) > f_lim
    auto bisect = [&](Phi& A, Phi& B) -> bool {
      while (ok && ls.evals < ls.max_evals) {
        Phi d;
        if (ev(0.5 * (A.a + B.a), d)) { c = d; return true; }
        if (!ok) return false;
        if (d.d >= 0) { B = d; return false; }
        if (d.f <= ls.f_lim) A = d; else B = d;
        if (B.a - A.a <= 1e-16 * fmax(1.0, fabs(B.a))) break;
      }
      return false;
    };
    // update(a,b,c): returns true if a Wolfe point was found during a nested bisection
    auto update = [&](Phi& A, Phi& B, const Phi& p) -> bool {
      if (!(p.a > A.a && p.a < B.a)) return false;
      if (p.d >= 0) { B = p; return false; }
      if (p.f <= ls.f_lim) { A = p; return false; }
      Phi Bb = p;
      const bool f = bisect(A, Bb);
      B = Bb;
      return f;
    };
    // bracket, starting from step 1 (tfp initial_step_size = 1)
    {
      Phi prev = lo;
      double a = 1.0;
      bool bracketed = false;
      while (ok && ls.evals < ls.max_evals) {
        if (ev(a, c)) { found = true; break; }
        if (!ok) break;
        if (c.d >= 0) { lo = prev; hi = c; bracketed = true; break; }
        if (c.f > ls.f_lim) {
          lo = Phi{0.0, fcur, dphi0}; hi = c;
          if (bisect(lo, hi)) found = true;
          bracketed = true;
          break;
        }
        prev = c;
        a *= 5.0;
      }
      if (!found && !bracketed) ok = false;
    }
    auto secant = [](const Phi& A, const Phi& B) { return (A.a * B.d - B.a * A.d) / (B.d - A.d); };
    while (ok && !found && ls.evals < ls.max_evals) {
      const Phi a0 = lo, b0 = hi;
      // secant2
      Phi p;
      double cs = secant(lo, hi);
      if (!isfinite(cs) || !(cs > lo.a && cs < hi.a)) cs = 0.5 * (lo.a + hi.a);
      if (ev(cs, p)) { c = p; found = true; break; }
      if (!ok) break;
      if (update(lo, hi, p)) { found = true; break; }
      double c2 = NAN;
      if (p.a == hi.a) c2 = secant(b0, hi);
      else if (p.a == lo.a) c2 = secant(a0, lo);
      if (isfinite(c2) && c2 > lo.a && c2 < hi.a && ls.evals < ls.max_evals) {
        Phi p2;
        if (ev(c2, p2)) { c = p2; found = true; break; }
        if (!ok) break;
        if (update(lo, hi, p2)) { found = true; break; }
      }
      if (hi.a - lo.a > 0.66 * (b0.a - a0.a) && ls.evals < ls.max_evals) {
        Phi pm;
        if (ev(0.5 * (lo.a + hi.a), pm)) { c = pm; found = true; break; }
        if (!ok) break;
        if (update(lo, hi, pm)) { found = true; break; }
      }
      if (hi.a - lo.a <= 1e-16 * fmax(1.0, hi.a)) break;
    }
    total_evals += ls.evals;
    if (ls.error) return 1;  // a CUDA / NCCL error inside an evaluation is an error, not a line-search failure
    if (!found) { R.failed = 1; break; }
    // accept: the last evaluation was at c (xt, fused hold x_new, g_new)
    k_lbfgs_push<<<1, 1024, 0, st>>>(P, head, h->d_x, h->d_g, h->d_xt, h->d_fused, h->d_S, h->d_Y, h->d_rho, h->d_scal);
    CK(cudaMemcpyAsync(sc, h->d_scal, sizeof sc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (sc[0] > 0.0 && isfinite(sc[0])) { head = (head + 1) % m; cnt = std::min(cnt + 1, m); }
    ginf = sc[1];
    const double fprev = fcur;
    fcur = c.f;
    ++R.iterations;
    if (ginf <= tol) R.converged = 1;
    if (fcur == fprev && c.a == 0.0) R.converged = 1;  // x_tolerance = f_relative_tolerance = 0
  }
  CK(cudaMemcpyAsync(h->d_params, h->d_x, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  CK(cudaStreamSynchronize(st));
  R.evaluations = total_evals;
  R.final_loss = value_unnorm ? fcur : fcur * h->lref;
  if (out) *out = R;
  return 0;
}

// ---------------------------------------------------------------- L-BFGS, controller on the device
// One loop trip = one objective evaluation:  xt = x + a*d  ->  loss/gradient at xt (the same kernels as an Adam
// step)  ->  g.d  ->  k_lb_post (Hager-Zhang state machine, lbfgs_ctl.h)  ->  k_lb_push  ->  k_lb_direction (iteration
// bookkeeping, two-loop recursion, next line search, loop condition).  "device" mode replays the trip as the body of
// a CUDA-graph WHILE node: the host launches ONE graph and synchronises once per batch (the whole optimisation, or
// every ring_cap evaluations when loss_info rows have to be drained); "host" mode enqueues the same kernels trip
// by trip and synchronises after each one (per-evaluation callbacks, multi-GPU runs).
static int lb_enqueue_trip(pinn_engine* h, unsigned long long cond, int set_cond) {
  cudaStream_t st = h->stream;
  const int P = h->fmap.n_params;
  CK(lb_begin_eval(P, h->d_x, h->d_d, h->d_ctl, h->d_xt, h->d_trace, st));
  if (enqueue_eval(h, h->d_xt, 0)) return 1;
  k_dot_inf<<<1, 1024, 0, st>>>(P, h->d_fused, h->d_d, h->d_scal);
  CK(cudaGetLastError());
  CK(lb_post_eval(h->d_ctl, h->d_ring, h->d_ring_pos, h->n_info, h->d_scal, st));
  CK(lb_push(P, h->d_ctl, h->d_x, h->d_g, h->d_xt, h->d_fused, h->d_S, h->d_Y, h->d_rho, h->d_scal2, st));
  CK(lb_direction(P, h->d_ctl, h->d_g, h->d_S, h->d_Y, h->d_rho, h->d_d, h->d_alpha, h->d_scal2, cond, set_cond, st));
  return 0;
}

static int lb_build_graph(pinn_engine* h) {
  if (h->lb_exec) { cudaGraphExecDestroy(h->lb_exec); h->lb_exec = nullptr; }
  cudaGraph_t g = nullptr;
  CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle handle;
  cudaError_t e = cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail("cudaGraphConditionalHandleCreate: %s", cudaGetErrorString(e)); }
  cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
  memset(&np, 0, sizeof np);
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  e = cudaGraphAddNode(&node, g, nullptr, 0, &np);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail("cudaGraphAddNode(conditional): %s", cudaGetErrorString(e)); }
  cudaGraph_t body = np.conditional.phGraph_out[0];
  e = cudaStreamBeginCaptureToGraph(h->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail("cudaStreamBeginCaptureToGraph: %s", cudaGetErrorString(e)); }
  const int rc = lb_enqueue_trip(h, (unsigned long long)handle, 1);
  cudaGraph_t captured = nullptr;
  e = cudaStreamEndCapture(h->stream, &captured);
  if (rc) { cudaGraphDestroy(g); return 1; }
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail("L-BFGS loop capture: %s", cudaGetErrorString(e)); }
  e = cudaGraphInstantiate(&h->lb_exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { h->lb_exec = nullptr; return fail("L-BFGS loop instantiate: %s", cudaGetErrorString(e)); }
  h->lb_valid = true;
  return 0;
}

extern "C" int pinn_engine_lbfgs(pinn_engine_t* h, int32_t max_iter, double tol, int32_t value_unnorm,
                                 pinn_eval_cb cb, void* user, pinn_lbfgs_result_t* out) {
  CK(cudaSetDevice(h->device));
  if (lbfgs_alloc(h)) return 1;
  const char* env = getenv("PINN_B200_LBFGS");
  std::string mode = env ? env : "auto";
  if (mode == "legacy") return lbfgs_legacy(h, max_iter, tol, value_unnorm, cb, user, out);
  // auto: the device-resident loop on one GPU; with a communicator the collective is enqueued trip by trip
  if (mode == "auto") mode = h->comm ? "host" : "device";
  const int P = h->fmap.n_params;
  cudaStream_t st = h->stream;
  if (!h->d_ctl) {
    CK(cudaMalloc(&h->d_ctl, sizeof(LbfgsCtl)));
    CK(cudaMalloc(&h->d_scal2, sizeof(double) * 2));
  }
  LbfgsCtl c;
  memset(&c, 0, sizeof c);
  c.max_iter = max_iter; c.m = 10; c.value_unnorm = value_unnorm; c.ls_max_evals = 50; c.ring_cap = h->ring_cap;
  c.trace_cap = h->d_trace ? h->trace_cap : 0;
  c.tol = tol; c.lref = h->lref;
  c.init_eval = 1; c.a_next = 0.0;
  CK(cudaMemcpyAsync(h->d_ctl, &c, sizeof c, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->d_x, h->d_params, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(h->d_d, 0, sizeof(float) * P, st));
  CK(cudaMemsetAsync(h->d_g, 0, sizeof(float) * P, st));
  h->lbfgs_syncs = 0;
  h->trace_rows = 0;
  if (mode == "device" && !h->lb_valid) {
    CK(cudaStreamSynchronize(st));  // the copy of the host-side controller above must not be captured
    if (lb_build_graph(h)) {
      if (getenv("PINN_B200_DEBUG")) fprintf(stderr, "[pinn] device L-BFGS loop unavailable (%s); host loop\n", g_err.c_str());
      cudaGetLastError();
      mode = "host";
    }
  }
  std::vector<double> rows;
  bool finished = false;
  while (!finished) {
    CK(cudaMemsetAsync(h->d_ring_pos, 0, sizeof(int), st));
    if (mode == "device") {
      CK(cudaGraphLaunch(h->lb_exec, st));
    } else {
      if (lb_enqueue_trip(h, 0ull, 0)) return 1;
    }
    CK(cudaMemcpyAsync(&c, h->d_ctl, sizeof c, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ++h->lbfgs_syncs;
    const int n_new = mode == "device" ? c.rows : 1;
    if (cb && n_new > 0) {
      rows.resize((size_t)n_new * h->n_info);
      CK(cudaMemcpy(rows.data(), h->d_ring, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost));
      for (int r = 0; r < n_new; ++r) cb(rows.data() + (size_t)r * h->n_info, h->n_info, user);
    }
    finished = c.converged || c.failed || c.iter >= c.max_iter;
    if (!finished && mode == "device") {  // ring drained: continue the loop where it stopped
      c.rows = 0; c.stop = 0;
      CK(cudaMemcpyAsync(h->d_ctl, &c, sizeof c, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
    }
  }
  h->trace_rows = std::min(c.total_evals, c.trace_cap);
  CK(cudaMemcpyAsync(h->d_params, h->d_x, sizeof(float) * P, cudaMemcpyDeviceToDevice, st));
  CK(cudaStreamSynchronize(st));
  pinn_lbfgs_result_t R{};
  R.iterations = c.iter; R.evaluations = c.total_evals; R.converged = c.converged; R.failed = c.failed;
  R.final_loss = value_unnorm ? c.fcur : c.fcur * h->lref;
  if (out) *out = R;
  return 0;
}

/* trace of the L-BFGS trial points (test / debugging aid): keep the first `cap` evaluated parameter vectors */
extern "C" int pinn_engine_lbfgs_trace(pinn_engine_t* h, int32_t cap) {
  CK(cudaSetDevice(h->device));
  if (h->d_trace) { cudaFree(h->d_trace); h->d_trace = nullptr; }
  h->trace_cap = 0; h->trace_rows = 0;
  h->lb_valid = false;  // the trace pointer is baked into the loop graph
  if (cap > 0) {
    CK(cudaMalloc(&h->d_trace, sizeof(float) * (size_t)cap * h->fmap.n_params));
    h->trace_cap = cap;
  }
  return 0;
}
extern "C" int32_t pinn_engine_lbfgs_trace_rows(pinn_engine_t* h) { return h->trace_rows; }
extern "C" int pinn_engine_lbfgs_trace_get(pinn_engine_t* h, float* out_host, int32_t rows) {
  CK(cudaSetDevice(h->device));
  if (rows > h->trace_rows) return fail("trace holds %d rows", h->trace_rows);
  if (rows > 0) CK(cudaMemcpy(out_host, h->d_trace, sizeof(float) * (size_t)rows * h->fmap.n_params, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int pinn_lbfgs_direction_test(int device, int32_t n, int32_t m, int32_t cnt, int32_t head, const float* g, const float* S,
                                         const float* Y, const double* rho, float* d_out) {
  if (n <= 0 || m <= 0 || m > 10 || cnt < 0 || cnt > m) return fail("lbfgs_direction_test: bad sizes");
  CK(cudaSetDevice(device));
  float *dg = nullptr, *dS = nullptr, *dY = nullptr, *dd = nullptr;
  double *drho = nullptr, *dscr = nullptr;
  auto cleanup = [&]() { for (void* p : {(void*)dg, (void*)dS, (void*)dY, (void*)dd, (void*)drho, (void*)dscr}) if (p) cudaFree(p); };
  cudaError_t e = cudaMalloc(&dg, sizeof(float) * n);
  if (e == cudaSuccess) e = cudaMalloc(&dS, sizeof(float) * (size_t)n * m);
  if (e == cudaSuccess) e = cudaMalloc(&dY, sizeof(float) * (size_t)n * m);
  if (e == cudaSuccess) e = cudaMalloc(&dd, sizeof(float) * n);
  if (e == cudaSuccess) e = cudaMalloc(&drho, sizeof(double) * m);
  if (e == cudaSuccess) e = cudaMalloc(&dscr, sizeof(double) * lb_scratch_doubles(n));
  if (e == cudaSuccess) e = cudaMemcpy(dg, g, sizeof(float) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dS, S, sizeof(float) * (size_t)n * m, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dY, Y, sizeof(float) * (size_t)n * m, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(drho, rho, sizeof(double) * m, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = lb_two_loop(n, nullptr, m, cnt, head, dg, dS, dY, drho, dd, dscr, 0);
  if (e == cudaSuccess) e = cudaMemcpy(d_out, dd, sizeof(float) * n, cudaMemcpyDeviceToHost);
  cleanup();
  if (e != cudaSuccess) return fail("lbfgs_direction_test: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int32_t pinn_engine_lbfgs_host_syncs(pinn_engine_t* h) { return h->lbfgs_syncs; }

// ---------------------------------------------------------------- NCCL
extern "C" int pinn_nccl_unique_id(uint8_t id_out[128]) {
  if (nccl_load()) return 1;
  Id128 id;
  memset(&id, 0, sizeof id);
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail("ncclGetUniqueId failed (%d)", rc);
  memcpy(id_out, id.b, 128);
  return 0;
}
extern "C" int pinn_engine_init_nccl(pinn_engine_t* h, const uint8_t id[128], int32_t rank, int32_t world) {
  CK(cudaSetDevice(h->device));
  if (nccl_load()) return 1;
  Id128 uid;
  memcpy(uid.b, id, 128);
  void* comm = nullptr;
  const int rc = g_nccl.CommInitRank(&comm, world, uid, rank);
  if (rc != 0) return fail("ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
  h->comm = comm;
  h->rank = rank;
  h->world = world;
  h->graph_valid = false; h->lb_valid = false;
  return 0;
}

// ---------------------------------------------------------------- device samplers
extern "C" int pinn_sample_lhs(int device, void* stream, uint32_t seed, int64_t n, int32_t d, const float* lo,
                               const float* hi, float* out_dev, int32_t ld, int32_t col0) {
  CK(cudaSetDevice(device));
  if (n <= 0) return 0;
  if (d < 1 || d > 3) return fail("d must be 1..3");
  if (n >= (1ll << 32)) return fail("n too large");
  uint32_t bits = 1;
  while ((1ull << (2 * bits)) < (unsigned long long)n) ++bits;
  const float3 l3 = make_float3(lo[0], d > 1 ? lo[1] : 0.f, d > 2 ? lo[2] : 0.f);
  const float3 h3 = make_float3(hi[0], d > 1 ? hi[1] : 0.f, d > 2 ? hi[2] : 0.f);
  k_sample_lhs<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out_dev, n, d, ld, col0, l3, h3, seed, bits);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int pinn_sample_cdf2d(int device, void* stream, uint32_t seed, int64_t n, const double* cum_host,
                                 int32_t ncy, int32_t ncx, float x0, float y0, float dx, float dy, float* out_dev,
                                 int32_t ld) {
  CK(cudaSetDevice(device));
  if (n <= 0) return 0;
  const int ncell = ncy * ncx;
  double* d_cum = nullptr;
  CK(cudaMalloc(&d_cum, sizeof(double) * (ncell + 1)));
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(d_cum, cum_host, sizeof(double) * (ncell + 1), cudaMemcpyHostToDevice, st));
  k_sample_cdf2d<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out_dev, n, ld, d_cum, ncell, ncx, x0, y0, dx, dy, seed);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  cudaFree(d_cum);
  return 0;
}

// ---------------------------------------------------------------- FMA microbenchmark
extern "C" int pinn_fma_peak(int device, int variant, double* tflops_out) {
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  float* d = nullptr;
  CK(cudaMalloc(&d, 4));
  const int grid = prop.multiProcessorCount * 8, iters = 4096;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a));
    if (variant == 0) k_fma_peak<0><<<grid, 256>>>(d, iters, 0.5f);
    else if (variant == 1) k_fma_peak<1><<<grid, 256>>>(d, iters, 0.5f);
    else if (variant <= 3) k_fma_outer<<<prop.multiProcessorCount * (variant == 2 ? 1 : 2), 256>>>(d, iters * 2, 0.5f);
    else if (variant <= 5) k_fma2_outer<<<prop.multiProcessorCount * (variant == 4 ? 1 : 2), 256>>>(d, iters * 2, 0.5f);
    else if (variant <= 7) k_fma2_outer_b<<<prop.multiProcessorCount * (variant == 6 ? 1 : 2), 256>>>(d, iters * 2, 0.5f);
    else k_mma_tf32_probe<<<prop.multiProcessorCount * (variant == 8 ? 1 : 2), 256>>>(d, iters * 2, 0.5f);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double flops = (variant >= 8) ? 2.0 * 8 * 1024 * (double)(iters * 2) * 8.0 * prop.multiProcessorCount * (variant == 8 ? 1 : 2)
                       : (variant <= 1) ? 2.0 * 16 * 8 * (double)iters * 256.0 * grid
                                        : 2.0 * 80 * (double)(iters * 2) * 256.0 * prop.multiProcessorCount * ((variant == 2 || variant == 4 || variant == 6) ? 1 : 2);
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  *tflops_out = best;
  return 0;
}
