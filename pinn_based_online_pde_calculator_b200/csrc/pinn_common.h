// Shared host/device descriptors for the fused jet-MLP kernels.
// Reference semantics: pinn_app/software.py:158-184 (network), 283-297 (residual),
// 310-383 (loss).  See DESIGN.md for the HBM layout.
#pragma once
#include <stdint.h>

#define PINN_MAX_SEG 16      // segments (BC groups / terms) one launch may cover
#define PINN_MAX_OPS 192     // residual program length
#define PINN_MAX_CONSTS 48
#define PINN_VM_STACK 12
#define PINN_MAX_LAYERS 16   // hidden layers
#define PINN_TOKENS 32       // flush-order tokens per shared gradient-accumulator row (one per layer + one for the final fold)
#define PINN_NT 256          // threads per CTA of the fused kernel (2 CTAs per SM, <=128 regs)
#define PINN_TU 8            // units per thread

enum PinnAct { PINN_TANH = 0, PINN_SIN = 1 };
enum PinnFeat { PINN_FEAT_AFFINE = 0, PINN_FEAT_POLAR = 1 };

// residual bytecode: op = word & 0xff, arg = word >> 8 (signed)
enum PinnOp {
  OP_CONST = 0,  // push consts[arg]
  OP_COORD = 1,  // push raw input coordinate column arg
  OP_JET = 2,    // push network output channel arg (derivative seed e_arg)
  OP_AUX = 3,    // push per-point aux column arg
  OP_ADD = 4, OP_SUB = 5, OP_MUL = 6, OP_DIV = 7, OP_NEG = 8,
  OP_POWI = 9,   // integer power arg >= 0
  OP_POWF = 10,  // real power consts[arg]
  OP_SIN = 11, OP_COS = 12, OP_EXP = 13, OP_LOG = 14, OP_TANH = 15, OP_SQRT = 16,
  OP_STORE_AUX = 17  // aux program only: pop -> aux column arg
};

struct PinnProgram {
  int n_ops;
  int ops[PINN_MAX_OPS];
  float consts[PINN_MAX_CONSTS];
};

// Network description + packed-parameter layout.
// "gpack" layout (size pg floats), used for packed weights AND gradient
// accumulators:  W0[4][WP] | b0[WP] | for l=1..L-1: W_l[WP][WP] | b_l[WP] | wl[WP] | bl[4]
// The weight pack additionally carries transposed hidden matrices W_l^T at off_wt[l].
struct PinnNet {
  int d_in, n_feat, feat_mode, n_hidden, width, wp;
  int act_first, act_hidden;
  float scl, epsil;
  float fa[3], fb[3];  // feature = fa*z + fb (affine) ; polar uses fa[0], fb[0] for r
  int off_w0, off_b0;
  int off_w[PINN_MAX_LAYERS], off_wt[PINN_MAX_LAYERS], off_b[PINN_MAX_LAYERS];
  int off_wl, off_bl;
  // combined second-order channel (MIX == 2): L = sum_i lap_beta_i d_ii ; beta_i from aux column
  // lap_aux[i] (>= 0) or the constant lap_beta[i]
  float lap_beta[3];
  int lap_aux[3];
  int pg;      // gpack size (floats)
  int pw;      // weight pack size (floats) = pg + (L-1)*WP*WP
};

struct PinnLaunch {
  PinnNet net;
  const float* wpack;
  const float* coords;    // [n][d_in]
  const float* aux;       // [n][n_aux] or null
  const float* base;      // [n][K] frozen base jets added to the network output, or null
  int n_aux;
  int n_seg;
  int seg_tile_end[PINN_MAX_SEG];        // cumulative tile count
  long long seg_pt_begin[PINN_MAX_SEG];  // first point of the segment
  long long seg_pt_end[PINN_MAX_SEG];
  int seg_slot[PINN_MAX_SEG];            // loss-term slot of the segment
  const float* seg_scale;  // device [n_slots]: 2*w_t/(N_t*lref)  (backward seed scale)
  float* stash;            // [grid][L][TP*K*WP]
  float* gacc;             // [grid][pg] CTA-private gradient accumulators (+=)
  double* loss_part;       // [grid][n_slots] (+=) sum of f^2
  int n_slots;
  float* out_u;            // eval: [n]
  float* out_f;            // eval: [n]
  float* out_jets;         // eval: [n][K] or null
  int n_tiles;
  long long* phase_clk;    // optional [8] clock64 totals per phase (CTA 0, thread 0), else null
  const void* wimg;        // tcgen05 family: pre-split bf16 weight images (stream of 4 KB chunks), else null
  long long wimg_copy_bytes;  // bytes of one replica of the image stream
  int wimg_copies;         // replicas (CTA b reads replica b % copies: spreads the stream over the L2 slices)
  int ldw;                 // row stride of the hidden matrices in wpack / gacc
  int gacc_share;          // tcgen05 family: CTAs per gradient-accumulator row (1 = CTA-private rows)
  unsigned* gacc_token;    // tcgen05 family, gacc_share > 1: [rows][PINN_TOKENS] flush-order tokens (zero before every launch)
  int exp_flags;           // experiment switches of the profiling instantiation (PINN_TC_EXP; 0 in production)
  PinnProgram prog;
};
