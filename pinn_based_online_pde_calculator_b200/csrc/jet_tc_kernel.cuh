// tcgen05 kernel family ("family D", kind 3): the production kernel for padded widths 128 and 256.
//
// All three hidden-layer GEMMs of the fused jet-MLP step run on tcgen05.mma kind::f16 (bf16 inputs,
// fp32 accumulators in Tensor Memory), with every fp32 operand split into THREE bf16 planes
// (x = b0 + b1 + b2, 24 significant bits) and six MMAs per product (b0*b0, b0*b1, b1*b0, b1*b1,
// b0*b2, b2*b0: everything down to 2^-24), the same cost as 3xTF32 but with 16-bit operands, whose
// swizzled shared-memory tiles can be read K-major AND MN-major through two descriptors.  That is
// what makes ONE on-chip copy of the activations serve the forward GEMM, the data-gradient GEMM and
// the (transposed) weight-gradient GEMM (layouts measured with tools/umma_probe_bf16.py,
// profiles/r02_umma_bf16_probe.txt).
//
// Roles are swapped with respect to a textbook GEMM: the UNITS of a layer are the M rows (TMEM
// lanes), the tile's (point, jet channel) pairs are the N columns:
//     forward        D[out][n]  = sum_in  W[in][out] * Y[in][n]      A = weights (K-major image, streamed
//     data gradient  D[in][n]   = sum_out W[in][out] * G[out][n]         through a cp.async.bulk ring),
//                                                                     B = activation tile, MN-major view
//     weight grad.   DW[in][out] = sum_n  Y[in][n] * G[out][n]        A, B = the same tiles, K-major view
// An epilogue thread owns ONE unit (= its TMEM lane) and loops over points: all jet channels of a
// (point, unit) pair are columns of the same lane, so the activation jets (sigma', sigma'' from the value
// channel), the bias add, and the bias / first-layer / output-layer gradients need no cross-thread
// exchange at all; operands are written with 16-byte stores along n.
//
// Warp roles: 4*Q epilogue warps (warp % 4 = TMEM lane quadrant = 32 units, warp / 4 = block of 8
// points), one MMA-issue warp, one producer warp that streams the pre-split weight images.
// Accumulation is two-level: the five small products go to a second TMEM accumulator and are added
// to the b0*b0 accumulator in the epilogue with a round-to-nearest add (the tensor core truncates
// when it accumulates: profiles/r02_umma_bf16_probe.txt).
//
// Reference semantics: pinn_app/software.py:158-184 (network), 246-297 (derivatives, residual),
// 318-379 (loss) and grad(loss_fun) (390); math as in jet_mma_kernel.cuh / SURVEY.md 8(a) addendum.
#pragma once
#include "jet_kernel.cuh"
#include "umma_common.cuh"

#ifndef TC_DISCARD
#define TC_DISCARD 0   // 1: discard the stash lines of a layer from the L2 (discard.global.L2, SASS CCTL.E.RML2) once its backward
                       // pass has consumed them, so that dead dirty lines are never written back.  Measured: C4 19.8 ms with,
                       // 19.2 ms without; C5 49.6 / 49.7 ms -- the write-backs are not what the kernel waits for.  Off.
#endif
// B2MERGE (one-M-block kernels): the weight-gradient operand Y^l is produced at the TOP of backward iteration l, from the
// stash values B1(l) has just loaded anyway (one stash read per layer instead of two), while dgrad(l+1) runs; the tensor
// core sees the same order of GEMMs as before.
// Measured on C4 (1M points, same box): bit-identical results, 19.75 ms with, 19.67 ms without -- the stash loads that
// were hidden behind the wait for dgrad(l+1) are now needed at once (Y-part 6.0 M -> 8.2 M cycles per CTA), which costs
// what the second read cost.  Off by default.
#ifndef TC_B2MERGE
#define TC_B2MERGE 0
#endif
// Measured on C5 (512k points, same box): 41.4 -> 40.7 ms (waiting for the weight-gradient blocks 20.4 M -> 13.0 M cycles
// per CTA, but the data-gradient GEMM now ends later: 10.7 M -> 15.2 M) -- its chunks are bound by the weight-image stream
// (192 KB per layer and 16-point tile at ~16 B/cycle), not by their place in the issue order.
#ifndef TC_DGRAD_INTERLEAVE
#define TC_DGRAD_INTERLEAVE 1
#endif
// bulk copies per ring stage.  Measured: 3 copies of 4 KB change nothing (C4 19.47 vs 19.56 ms, C5 40.6 vs 40.7), 12 copies of
// 1 KB cost C5 9 %: the GEMMs do not wait for the copy engine but for shared-memory bandwidth (per k-step the MMAs read
// ~31 KB of operands and the ring writes 12 KB: ~340 cycles at 128 B/cycle against 412 cycles of MMA issue).
#ifndef TC_RING_SPLIT
#define TC_RING_SPLIT 1
#endif
#ifndef TC_EXP
#define TC_EXP 0   // timing experiments (wrong results): 1 no accumulator flush, 2 no B2 stash read, 4 no stash write, 8 no B1 stash read
#endif

template <int WP_, int N1_, int N2_, int MIX_>
struct TcCfg {
  static constexpr int WP = WP_, N1 = N1_, N2 = N2_, MIX = MIX_;
  static constexpr bool LAP = (MIX == 2);
  static constexpr int K = 1 + N1 + N2 + (MIX ? 1 : 0);
  static constexpr int MB = WP / 128;                        // M blocks (128 units each)
  static constexpr int NP = (WP == 128 && K <= 4) ? 32 : 16;  // points per tile
  static constexpr int Q = NP / 8;                           // epilogue warps per lane quadrant (8 points each)
  static constexpr int V = 4;                                // points an epilogue thread holds in registers at a time
  static constexpr int NROW = NP * K;                        // N of the forward / data-gradient MMAs
  static constexpr int SWB = (NROW % 64 == 0) ? 128 : 32;    // swizzle span (bytes per line) of the activation tiles
  static constexpr int NBLK = (NROW * 2 + SWB - 1) / SWB;    // n-blocks per plane
  // HALVES (C4-type tiles: one M block, 32 points x 4 channels): the tile is two 16-point halves (pass h of every
  // epilogue thread) with their own operand regions, accumulator columns and forward MMAs, so that the epilogue of one
  // half runs under the forward GEMM of the other.  Column order n = h*64 + n8*16 + c*4 + i (a thread's 16 values of a
  // pass are 16 consecutive TMEM columns: one tcgen05.ld.x16 per accumulator block).
  // MEASURED on C4 (1M points, same box): 21.3 ms with HALVES against 19.1 ms without.  The forward wait for the tensor
  // core drops from 7.4 M to 3.9 M cycles per CTA, but the weight image of a layer is streamed twice, the epilogue passes
  // get 30 % slower while they share shared memory and TMEM with running MMAs (act_fwd 8.8 M -> 11.4 M cycles), and the
  // three-block accumulator reads cost registers (spills 132 -> 320 B) that the backward pass pays for.  Off by default.
#ifndef TC_HALVES
#define TC_HALVES 0
#endif
  static constexpr bool HALVES = (TC_HALVES != 0) && (WP == 128) && (NP == 32) && (K == 4);
  static constexpr int HN = NROW / 2;                        // columns per half
  static constexpr int HP = WP * SWB;                        // HALVES: one (half, plane) block = WP lines x 128 B; R1 order [h][plane]
  // FWD3: forward GEMM in the four-MMA form with THREE accumulator blocks per M block (A = leading products only,
  // B + C = the five small ones; see gemm_fwd3) instead of six MMAs into two blocks -- the same two-level precision at
  // two thirds of the tensor-pipe time.  TC_FWD3 bit 0: one-M-block kernels, bit 1: two-M-block kernels.
  // Measured (same box): C4 19.07 -> 18.78 ms (forward wait 7.2 M -> 6.5 M cycles per CTA), C5 unchanged (50.6 ms: its
  // forward wait drops 20.5 M -> 17.3 M cycles, but the step is bound by the weight-gradient flush).
#ifndef TC_FWD3
#define TC_FWD3 3
#endif
  static constexpr bool FWD3 = !HALVES && (((TC_FWD3 & 1) && WP == 128) || ((TC_FWD3 & 2) && WP == 256)) && (2 * NROW <= 256) &&
                               ((WP / 128) * 3 * NROW <= 512);
  static constexpr int FWD_COLS = HALVES ? 6 * HN : FWD3 ? (WP / 128) * 3 * NROW : (WP / 128) * 2 * NROW;   // forward accumulators
  static constexpr int PLANE1 = NBLK * WP * SWB;             // region 1: all WP lines
  static constexpr int PLANE2 = NBLK * 128 * SWB;            // region 2: one block of 128 lines
#ifndef TC_YP
#define TC_YP 2      // (1: timing experiment only -- wrong weight gradient)
#endif
#ifndef TC_NSLOT
#define TC_NSLOT 4
#endif
  static constexpr int YP = TC_YP;                           // planes of the recomputed layer input (weight gradient only)
  // [y0 | y1] as ONE N = 2*NROW operand (four MMAs per k-step instead of six).  Measured on C4: no gain (the tensor
  // pipe is not the limiter) and three truncating adds per k-step in the leading accumulator instead of one
  // (u 2.3e-7 -> 7.9e-7, gradient 6.9e-7 -> 1.2e-6), so it stays off.
#ifndef TC_CONCAT_FWD
#define TC_CONCAT_FWD 0
#endif
  static constexpr bool CONCAT = (TC_CONCAT_FWD != 0) && (MB == 2) && (2 * NROW <= 256);
  // ... but in the DATA-GRADIENT GEMM of the two-block kernel (padded width 256: N = 80 per plane, the tensor pipe is
  // the limiter there and an MMA costs 103 cycles for any N <= 192) the concatenated form saves a third of the MMAs;
  // the gradient tolerates the two extra truncating adds per k-step (measured on C5: see DESIGN.md).
  static constexpr bool CONCAT_DGRAD = (MB == 2) && (2 * NROW <= 256);
  // two-level accumulation (second TMEM block for the small products) also in the data-gradient GEMM?
  // Measured on C4: one accumulator saves TMEM reads in B1 (5.9 -> 5.2 k cycles per layer) but not a microsecond of
  // the step (the backward pass is bound by the dgrad -> wgrad chain on the tensor pipe), and doubles the gradient
  // error (6.9e-7 -> 1.5e-6): stays on.
  static constexpr bool DGRAD_TWO_LEVEL = true;
  static constexpr int R1_BYTES = 3 * PLANE1, R2_BYTES = YP * PLANE2;
  static constexpr int NEPI = 4 * Q;
  static constexpr int NEPI_T = NEPI * 32;
  static constexpr int NT = NEPI_T + 64;                     // + MMA-issue warp + producer warp
  static constexpr int KS = WP / 16;                         // k-steps of the forward / data-gradient GEMMs
  static constexpr int KSW = NROW / 16;                      // k-steps of the weight-gradient GEMM
  static constexpr int PLANE_W = 4096;                       // one weight plane of a k-step: 128 rows x 32 B
  static constexpr int SLOT = 3 * PLANE_W;                   // ring stage = one k-step, planes b2 | b1 | b0
  static constexpr int NSLOT = TC_NSLOT;
  static constexpr int CHUNKS = MB * KS;                     // ring stages per forward / data-gradient GEMM
  static constexpr int PARTLD = NROW + 4;                    // row stride of the output-layer partial products
  // TMEM columns: per M block (big | small) accumulators, then two weight-gradient blocks (ping-pong)
  __host__ __device__ static constexpr int TC_D(int mb) { return mb * 2 * NROW; }
  static constexpr int TC_DW = MB * 2 * NROW;                // weight-gradient blocks of 128 columns: two (ping-pong) for one M block,
  static constexpr int NDW = (MB == 1) ? 2 : 1;              // one for two (the accumulators of both M blocks take 4*NROW columns)
  static constexpr int TC_USED = TC_DW + 128 * NDW;
  static constexpr int MISC_FLOATS = NP * 4 /*z*/ + 3 * NP /*beta*/ + NP * K * 4 /*hj*/ + NROW /*ubar*/ + 4 * NROW /*psum*/ +
                                     Q * WP /*gradient partials*/ + PINN_MAX_OPS + PINN_MAX_CONSTS;
  static constexpr size_t smem_bytes() { return (size_t)R1_BYTES + R2_BYTES + NSLOT * SLOT + MISC_FLOATS * 4 + 256; }
  // YSIDE (one M block): the forward pass also keeps the bf16 planes b0, b1 of every layer's OUTPUT tile -- a raw copy of
  // the swizzled operand region, stored by the bulk-copy engine (cp.async.bulk shared -> global) while the next GEMM runs
  // -- and the backward pass loads them straight back into the weight-gradient operand region R2, instead of
  // recomputing them from the fp32 stash in the epilogue warps (B2: 15 % of the C4 step).
  // MEASURED on C4 (1M points, same box): correct (bit-identical gradient), B2 4.2 M cycles per CTA cheaper, but the forward
  // GEMM wait grows by 4.5 M (the 64 KB bulk store out of the operand tile takes longer than the GEMM that reads the same
  // tile) and the flush by 2.7 M (it now shares the L2 path with the 64 KB reload): 21.8 ms vs 19.2 ms.  Off by default.
#ifndef TC_YSIDE
#define TC_YSIDE 0
#endif
  static constexpr bool YSIDE = (TC_YSIDE != 0) && (MB == 1) && !HALVES;
  static constexpr size_t STL_F = (size_t)(K + 1) * WP * NP;           // fp32 stash floats per layer and CTA (K jets + cos for the sin activation)
  static constexpr size_t YS_F = YSIDE ? (size_t)YP * PLANE2 / 4 : 0;  // side stash (floats) per layer and CTA
  static constexpr size_t STL = STL_F + YS_F;                          // per-layer stride of the CTA's scratch
  static constexpr bool OK = (NROW % 16 == 0) && (NROW <= 256) && (TC_USED <= 512) && (FWD_COLS <= 512) && (smem_bytes() <= 232448 - 1024) &&
                             ((size_t)WP * PARTLD * 4 <= (size_t)R1_BYTES) && (MB <= 2);
};

namespace tc {

// ---- bf16 split of a pair (x0 -> lower half, x1 -> upper half); every residual is exact
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
template <int NPL>
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t (&p)[3]) {
  p[0] = cvt_bf16x2(x0, x1);
  const float r0 = x0 - __uint_as_float(p[0] << 16), r1 = x1 - __uint_as_float(p[0] & 0xffff0000u);
  p[1] = cvt_bf16x2(r0, r1);
  if (NPL > 2) {
    const float s0 = r0 - __uint_as_float(p[1] << 16), s1 = r1 - __uint_as_float(p[1] & 0xffff0000u);
    p[2] = cvt_bf16x2(s0, s1);
  }
}

// byte offset of the 16-byte chunk (line, n8 = n / 8) inside a plane of `lines` lines
template <int SWB>
__device__ __forceinline__ uint32_t chunk_off(int line, int n8, int lines) {
  if (SWB == 128) return (uint32_t)((n8 >> 3) * (lines * 128) + line * 128 + (((n8 & 7) ^ (line & 7)) << 4));
  return (uint32_t)((n8 >> 1) * (lines * 32) + line * 32 + (((n8 & 1) ^ ((line >> 2) & 1)) << 4));
}
// byte offset of k-step ks (16 n) for the K-major view of a plane of `lines` lines
template <int SWB>
__device__ __forceinline__ uint32_t kmajor_koff(int ks, int lines) {
  if (SWB == 128) return (uint32_t)((ks >> 2) * (lines * 128) + (ks & 3) * 32);
  return (uint32_t)(ks * (lines * 32));
}
template <int SWB>
__device__ __forceinline__ constexpr uint32_t layout_type() { return SWB == 128 ? 2u : 6u; }

// split 4 values (half h of the 8-element chunk) and store them into NPL planes
template <int NPL>
__device__ __forceinline__ void store_split4(uint8_t* chunk, int plane_bytes, int h, const float (&v)[4]) {
  uint32_t a[3], b[3];
  split_pair<NPL>(v[0], v[1], a);
  split_pair<NPL>(v[2], v[3], b);
  uint8_t* d = chunk + 8 * h;
#pragma unroll
  for (int p = 0; p < NPL; ++p) *reinterpret_cast<uint2*>(d + p * plane_bytes) = make_uint2(a[p], b[p]);
}

// split 4 values and store them into NPL planes at the 8-byte slot `d` of plane 0
template <int NPL>
__device__ __forceinline__ void store_split4p(uint8_t* d, int plane_bytes, const float (&v)[4]) {
  uint32_t a[3], b[3];
  split_pair<NPL>(v[0], v[1], a);
  split_pair<NPL>(v[2], v[3], b);
#pragma unroll
  for (int p = 0; p < NPL; ++p) *reinterpret_cast<uint2*>(d + p * plane_bytes) = make_uint2(a[p], b[p]);
}

// inline sin/cos, fast path only: three-constant Cody-Waite reduction by pi/2 (quadrant from the magic-number
// rounding trick, no conversion instructions) + minimax polynomials on [-pi/4, pi/4] (about 1 ulp for
// |x| <= 3e4); straight-line code, so the V evaluations of a thread interleave
__device__ __forceinline__ void sincos_fast(float x, float& s, float& c) {
  const float kf = fmaf(x, 0.636619772f, 12582912.0f);   // 1.5 * 2^23: the integer lands in the low mantissa bits
  const int q = __float_as_int(kf);
  const float k = kf - 12582912.0f;
  float r = fmaf(k, -1.57079601e+00f, x);
  r = fmaf(k, -3.13916473e-07f, r);
  r = fmaf(k, -5.39030253e-15f, r);
  const float r2 = r * r;
  float sp = fmaf(r2, -1.95152959e-4f, 8.33216087e-3f);
  sp = fmaf(sp, r2, -1.66666546e-1f);
  sp = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.44331571e-5f, -1.38873163e-3f);
  cp = fmaf(cp, r2, 4.16666457e-2f);
  cp = fmaf(cp, r2, -0.5f);
  cp = fmaf(cp, r2, 1.0f);
  const float ss = (q & 1) ? cp : sp, cc = (q & 1) ? sp : cp;
  s = (q & 2) ? -ss : ss;
  c = ((q + 1) & 2) ? -cc : cc;
}

// y and the first two derivatives of the activation for V pre-activations.  The backward pass needs no
// pre-activation again: it restarts from the stashed y (tanh) or from the stashed (sin, cos) pair.
template <int V>
__device__ __forceinline__ void act_fwd(int act, const float (&a)[V], float (&y)[V], float (&d1)[V], float (&d2)[V]) {
  if (act == PINN_TANH) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float t = tanh_bf(a[i]);
      y[i] = t;
      d1[i] = fmaf(-t, t, 1.0f);
      d2[i] = -2.0f * t * d1[i];
    }
  } else {
    bool huge = false;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float sn, cs;
      sincos_fast(a[i], sn, cs);
      y[i] = sn; d1[i] = cs; d2[i] = -sn;
      huge |= fabsf(a[i]) > 3.0e4f;
    }
    if (huge) {  // rare: arguments beyond the range of the three-constant reduction take the library path
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float sn, cs;
        sincos_ni(a[i], &sn, &cs);
        y[i] = sn; d1[i] = cs; d2[i] = -sn;
      }
    }
  }
}
// first, second (and third) derivative of the activation from the stashed output y (for sin: d1 holds the stashed cos)
template <int V, bool D3>
__device__ __forceinline__ void act_bwd(int act, const float (&y)[V], float (&d1)[V], float (&d2)[V], float (&d3)[V]) {
  if (act == PINN_TANH) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float t = y[i];
      d1[i] = fmaf(-t, t, 1.0f);
      d2[i] = -2.0f * t * d1[i];
      if (D3) d3[i] = d1[i] * fmaf(6.0f * t, t, -2.0f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) { d2[i] = -y[i]; if (D3) d3[i] = -d1[i]; }
  }
}

// pre-activation jets a[1..K-1] (a[0] is overwritten with y) -> output jets, in place
template <class C, int V>
__device__ __forceinline__ void jets_outputs(float (&a)[C::K][V], const float (&y)[V], const float (&d1)[V], const float (&d2)[V],
                                             const float (&beta)[3][V]) {
#pragma unroll
  for (int i = 0; i < V; ++i) {
#pragma unroll
    for (int k = 0; k < C::N2; ++k) {
      const float Ai = a[1 + k][i];
      a[1 + C::N1 + k][i] = fmaf(d2[i] * Ai, Ai, d1[i] * a[1 + C::N1 + k][i]);
    }
    if (C::MIX == 1) a[C::K - 1][i] = fmaf(d2[i] * a[1][i], a[2][i], d1[i] * a[C::K - 1][i]);
    if (C::LAP) {
      float S = 0.f;
#pragma unroll
      for (int k = 0; k < C::N1; ++k) S = fmaf(beta[k][i] * a[1 + k][i], a[1 + k][i], S);
      a[C::K - 1][i] = fmaf(d2[i], S, d1[i] * a[C::K - 1][i]);
    }
#pragma unroll
    for (int k = 0; k < C::N1; ++k) a[1 + k][i] *= d1[i];
    a[0][i] = y[i];
  }
}

// adjoint of the activation jets: yb = adjoint of the layer outputs (in), st = stashed jets (st[0] = y, st[c>0] =
// pre-activation jets); on return yb holds the adjoint of the pre-activations
template <class C, int V>
__device__ __forceinline__ void jets_adjoint(float (&yb)[C::K][V], const float (&st)[C::K][V], const float (&d1)[V], const float (&d2)[V],
                                             const float (&d3)[V], const float (&beta)[3][V]) {
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float ab0 = d1[i] * yb[0][i];
    float ab1[C::N1 > 0 ? C::N1 : 1];
#pragma unroll
    for (int k = 0; k < C::N1; ++k) {
      const float v = yb[1 + k][i];
      ab1[k] = d1[i] * v;
      ab0 = fmaf(d2[i] * st[1 + k][i], v, ab0);
    }
#pragma unroll
    for (int k = 0; k < C::N2; ++k) {
      const float v = yb[1 + C::N1 + k][i];
      const float Ai = st[1 + k][i], Aii = st[1 + C::N1 + k][i];
      yb[1 + C::N1 + k][i] = d1[i] * v;
      ab1[k] = fmaf(2.0f * d2[i] * Ai, v, ab1[k]);
      ab0 = fmaf(fmaf(d3[i] * Ai, Ai, d2[i] * Aii), v, ab0);
    }
    if (C::MIX == 1) {
      const float v = yb[C::K - 1][i];
      const float A0 = st[1][i], A1 = st[2][i], A01 = st[C::K - 1][i];
      yb[C::K - 1][i] = d1[i] * v;
      ab1[0] = fmaf(d2[i] * A1, v, ab1[0]);
      ab1[C::N1 > 1 ? 1 : 0] = fmaf(d2[i] * A0, v, ab1[C::N1 > 1 ? 1 : 0]);
      ab0 = fmaf(fmaf(d3[i] * A0, A1, d2[i] * A01), v, ab0);
    }
    if (C::LAP) {
      const float v = yb[C::K - 1][i];
      float S = 0.f;
#pragma unroll
      for (int k = 0; k < C::N1; ++k) {
        const float bA = beta[k][i] * st[1 + k][i];
        S = fmaf(bA, st[1 + k][i], S);
        ab1[k] = fmaf(2.0f * d2[i] * bA, v, ab1[k]);
      }
      yb[C::K - 1][i] = d1[i] * v;
      ab0 = fmaf(fmaf(d3[i], S, d2[i] * st[C::K - 1][i]), v, ab0);
    }
#pragma unroll
    for (int k = 0; k < C::N1; ++k) yb[1 + k][i] = ab1[k];
    yb[0][i] = ab0;
  }
}

__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }

// bounded mbarrier wait: a protocol bug traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  if (!umma::mbar_wait(bar, parity, 1u << 24)) __trap();
}
// TMA bulk prefetch of a contiguous global range into L2 (no destination, no completion tracking)
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// drop a 128-byte line from the L2 WITHOUT writing it back (its contents become indeterminate): for scratch whose
// last reader is done -- the next tile overwrites it before anybody reads it again
__device__ __forceinline__ void discard_l2(const void* p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }
// bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(umma::smem_addr(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources read
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }        // writes done
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

}  // namespace tc

enum { TC_BAR_EPI = 1, TC_BAR_OP1 = 2, TC_BAR_OP2 = 3, TC_BAR_OP3 = 4, TC_BAR_OP1B = 5 };

// ---------------------------------------------------------------- the kernel
template <class C, bool TRAIN, bool PROF = false>
__global__ void __launch_bounds__(C::NT, 1) jet_tc_kernel(const __grid_constant__ PinnLaunch L) {
  static_assert(C::OK, "invalid tcgen05 kernel configuration");
  constexpr int K = C::K, WP = C::WP, NP = C::NP, Q = C::Q, NROW = C::NROW, SWB = C::SWB, V = C::V;
  constexpr int NSLOT = C::NSLOT, PB = NP / 8, NH = 8 / V;
  constexpr uint32_t LT = tc::layout_type<SWB>();
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  uint8_t* const smem = tc_smem;
  uint8_t* const R1 = smem;                                    // forward: layer input Y; backward: adjoints G (3 planes)
  uint8_t* const R2 = smem + C::R1_BYTES;                      // backward: recomputed layer input Y (2 planes, 128 lines)
  uint8_t* const ring = R2 + C::R2_BYTES;                      // weight-image ring
  float* const s_z = reinterpret_cast<float*>(ring + NSLOT * C::SLOT);  // [NP][4]: coordinates, valid flag
  float* const s_beta = s_z + NP * 4;                          // [3][NP]
  float* const s_hj = s_beta + 3 * NP;                         // [NP][K][4] feature jets
  float* const s_ubar = s_hj + NP * K * 4;                     // [NROW] epsil * adjoint of the network outputs
  float* const s_psum = s_ubar + NROW;                         // [4][NROW] output-layer partial sums
  float* const s_bg = s_psum + 4 * NROW;                       // [Q][WP] per-warp gradient partials (fixed-order fold)
  int* const s_ops = reinterpret_cast<int*>(s_bg + Q * WP);
  float* const s_consts = reinterpret_cast<float*>(s_ops + PINN_MAX_OPS);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(s_consts + PINN_MAX_CONSTS);
  uint64_t* const bar_full = bars;
  uint64_t* const bar_empty = bars + NSLOT;
  uint64_t* const bar_fd = bars + 2 * NSLOT;                   // [2]: forward GEMM of half h (HALVES); [0] otherwise and for dgrad
  uint64_t* const bar_w = bar_fd + 2;
  uint64_t* const bar_y = bar_w + 1;                           // YSIDE: side-stash tile landed in R2
  __shared__ uint32_t s_tmem;
  float* const part = reinterpret_cast<float*>(R1);            // [WP][PARTLD] output-layer partial products

  const PinnNet& net = L.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Lh = net.n_hidden, NG = Lh - 1;
  const int my_tiles = (L.n_tiles > (int)blockIdx.x) ? (L.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  for (int i = tid; i < L.prog.n_ops; i += C::NT) s_ops[i] = L.prog.ops[i];
  for (int i = tid; i < PINN_MAX_CONSTS; i += C::NT) s_consts[i] = L.prog.consts[i];
  if (warp == C::NEPI) umma::tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) { umma::mbar_init(&bar_full[i], 1); umma::mbar_init(&bar_empty[i], 1); }
    umma::mbar_init(&bar_fd[0], 1);
    umma::mbar_init(&bar_fd[1], 1);
    umma::mbar_init(bar_w, 1);
    umma::mbar_init(bar_y, 1);
    umma::fence_mbar_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = s_tmem;

  if (warp == C::NEPI + 1) {
    // ================================================================ producer: weight-image stream
    // One 12 KB stage = the three bf16 planes of one k-step.  Each CTA reads replica blockIdx % copies of the
    // image stream (every CTA walks the same sequence at about the same time).
    if (lane == 0) {
      const uint8_t* img = reinterpret_cast<const uint8_t*>(L.wimg) + (size_t)(blockIdx.x % L.wimg_copies) * L.wimg_copy_bytes;
      // HALVES: every forward GEMM is issued once per half, so its image is streamed twice in a row
      constexpr int REP = C::HALVES ? 2 : 1;
      const int fwd_len = REP * NG * C::CHUNKS;
      const int per_tile = fwd_len + (TRAIN ? NG * C::CHUNKS : 0);
      const long long total = (long long)my_tiles * per_tile;
      int pos = 0;   // position inside the tile's stream
      for (long long i = 0; i < total; ++i) {
        const int s = (int)(i % NSLOT);
        const uint32_t round = (uint32_t)(i / NSLOT);
        // image chunk of this position
        const int c = (pos < fwd_len) ? (pos / (REP * C::CHUNKS)) * C::CHUNKS + pos % C::CHUNKS : NG * C::CHUNKS + (pos - fwd_len);
        if (round > 0) tc::wait_bar(&bar_empty[s], (round - 1) & 1);
        mbar_expect_tx(&bar_full[s], C::SLOT);
#pragma unroll
        for (int part = 0; part < TC_RING_SPLIT; ++part)   // (several smaller bulk copies per stage: more copies in flight)
          bulk_g2s(ring + s * C::SLOT + part * (C::SLOT / TC_RING_SPLIT), img + (size_t)c * C::SLOT + part * (C::SLOT / TC_RING_SPLIT),
                   C::SLOT / TC_RING_SPLIT, &bar_full[s]);
        if (TRAIN && c >= NG * C::CHUNKS && c % C::CHUNKS == 0 && !(PROF && (L.exp_flags & 4))) {
          // first stage of dgrad(l): pull what the epilogue warps touch one layer later into L2 -- the stash of layer
          // l-1 (B2(l), B1(l-1)) and the accumulator block of layer l (flushed during iteration l-1).  The per-CTA
          // scratch of all SMs together is about the size of the L2, so a part of it lives in HBM between uses.
          const int l = (Lh - 1) - (c / C::CHUNKS - NG);
          if (l > 1) tc::prefetch_l2(L.stash + ((size_t)blockIdx.x * Lh + (l - 1)) * C::STL, (uint32_t)(C::STL_F * 4));  // layer 0 has no stash
          for (int blk = 0; blk < C::MB; ++blk)
            tc::prefetch_l2(L.gacc + (size_t)(blockIdx.x / (L.gacc_share > 1 ? L.gacc_share : 1)) * net.pg + net.off_w[l] + (size_t)blk * 128 * L.ldw, (uint32_t)(128 * L.ldw * 4));
        }
        if (++pos == per_tile) pos = 0;
      }
    }
  } else if (warp == C::NEPI) {
    // ================================================================ MMA issue warp
    long long ci = 0;
    const uint32_t id_wx = umma::idesc_bf16(128, NROW, 0, 1);
    const uint32_t id_wx2 = umma::idesc_bf16(128, (2 * NROW <= 256) ? 2 * NROW : NROW, 0, 1);
    const uint32_t id_wg = umma::idesc_bf16(128, 128, 0, 0);
    const uint32_t r1a = umma::smem_addr(R1), r2a = umma::smem_addr(R2), rga = umma::smem_addr(ring);
    // D = W x act(R1), weights from the ring.  Per k-step FOUR MMAs: the activation planes b0 and b1 lie back to
    // back in shared memory, so ONE descriptor with N = 2*NROW reads [y0 | y1] and W_p x [y0 | y1] (p = 2, 1, 0) yields
    // the columns [sum_p w_p*y0 | sum_p w_p*y1] at the N = 256 rate (128 cycles instead of 2 x 103); the sixth
    // product w0*y2 goes into the second block.  The epilogue adds the two blocks with a round-to-nearest add
    // (two-level accumulation; the extra w2*y1 term is 2^-24 of the leading one).
    // (chunk q of NQ: the k-steps [q, q+1) * MB*KS/NQ of the (M block, k-step) sequence -- the data-gradient GEMM of the
    // two-block kernel is issued in four chunks between the weight-gradient blocks; the commit comes with the last chunk)
    auto gemm_wx = [&](bool two_level, bool concat, bool store_in_flight = false, int q = 0, int NQ = 1) {
      const int per = C::MB * C::KS / NQ;
      for (int mb = 0; mb < C::MB; ++mb) {
        const uint32_t DB = tb + C::TC_D(mb), DS = two_level ? DB + NROW : DB;
        for (int ks = 0; ks < C::KS; ++ks) {
          if ((mb * C::KS + ks) / per != q) continue;
          // HALVES (data-gradient GEMM over the whole tile): plane p of half h sits at (3h + p) * HP, so the two n-blocks
          // of a plane are 3 * HP apart
          constexpr uint32_t PST = C::HALVES ? C::HP : C::PLANE1, LBO = C::HALVES ? 3 * C::HP : WP * SWB;
          const uint64_t b01 = umma::smem_desc(r1a + ks * 16 * SWB, LBO, 8 * SWB, LT);
          const uint64_t b2 = umma::smem_desc(r1a + 2 * PST + ks * 16 * SWB, LBO, 8 * SWB, LT);
          const uint32_t acc = ks > 0 ? 1u : 0u;
          const int s = (int)(ci % NSLOT);
          tc::wait_bar(&bar_full[s], (uint32_t)((ci / NSLOT) & 1));
          umma::fence_after_sync();
          const uint64_t w2 = umma::smem_desc(rga + s * C::SLOT, 16, 256, 6);
          const uint64_t w1 = umma::smem_desc(rga + s * C::SLOT + C::PLANE_W, 16, 256, 6);
          const uint64_t w0 = umma::smem_desc(rga + s * C::SLOT + 2 * C::PLANE_W, 16, 256, 6);
          if (concat) {
            umma::mma_bf16_ss(DB, w2, b01, id_wx2, acc);
            umma::mma_bf16_ss(DB, w1, b01, id_wx2, 1u);
            umma::mma_bf16_ss(DB, w0, b01, id_wx2, 1u);
            umma::mma_bf16_ss(DS, w0, b2, id_wx, 1u);
          } else {
            const uint64_t b1 = umma::smem_desc(r1a + PST + ks * 16 * SWB, LBO, 8 * SWB, LT);
            umma::mma_bf16_ss(DS, w2, b01, id_wx, acc);
            umma::mma_bf16_ss(DS, w1, b1, id_wx, 1u);
            umma::mma_bf16_ss(DS, w1, b01, id_wx, 1u);
            umma::mma_bf16_ss(DS, w0, b2, id_wx, 1u);
            umma::mma_bf16_ss(DS, w0, b1, id_wx, 1u);
            umma::mma_bf16_ss(DB, w0, b01, id_wx, two_level ? acc : 1u);
          }
          umma::commit(&bar_empty[s]);
          ++ci;
        }
      }
      // (YSIDE: the side-stash store reads the same operand tile; the epilogue may overwrite it once bar_fd completes)
      if (store_in_flight) tc::bulk_wait_read();
      if (q == NQ - 1) umma::commit(bar_fd);
    };
    // FWD3 forward GEMM: per M block the accumulator blocks A | B | C of NROW columns,
    //   C = w2*y0;  A, B = w0*[y0 | y1];  B, C += w1*[y0 | y1];  C += w0*y2        (epilogue: A + (B + C))
    auto gemm_fwd3 = [&](bool store_in_flight) {
      for (int mb = 0; mb < C::MB; ++mb) {
        const uint32_t DA = tb + mb * 3 * NROW;
        for (int ks = 0; ks < C::KS; ++ks) {
          const uint64_t b01 = umma::smem_desc(r1a + ks * 16 * SWB, WP * SWB, 8 * SWB, LT);
          const uint64_t b2 = umma::smem_desc(r1a + 2 * C::PLANE1 + ks * 16 * SWB, WP * SWB, 8 * SWB, LT);
          const uint32_t acc = ks > 0 ? 1u : 0u;
          const int s = (int)(ci % NSLOT);
          tc::wait_bar(&bar_full[s], (uint32_t)((ci / NSLOT) & 1));
          umma::fence_after_sync();
          const uint64_t w2 = umma::smem_desc(rga + s * C::SLOT, 16, 256, 6);
          const uint64_t w1 = umma::smem_desc(rga + s * C::SLOT + C::PLANE_W, 16, 256, 6);
          const uint64_t w0 = umma::smem_desc(rga + s * C::SLOT + 2 * C::PLANE_W, 16, 256, 6);
          umma::mma_bf16_ss(DA + 2 * NROW, w2, b01, id_wx, acc);
          umma::mma_bf16_ss(DA, w0, b01, id_wx2, acc);
          umma::mma_bf16_ss(DA + NROW, w1, b01, id_wx2, 1u);
          umma::mma_bf16_ss(DA + 2 * NROW, w0, b2, id_wx, 1u);
          umma::commit(&bar_empty[s]);
          ++ci;
        }
      }
      if (store_in_flight) tc::bulk_wait_read();
      umma::commit(bar_fd);
    };
    // HALVES forward GEMM of half h: FOUR MMAs per k-step into three accumulator blocks of HN columns,
    //   C    = w2*y0            (N = HN;   initialises C at k-step 0)
    //   A, B = w0*[y0 | y1]     (N = 2 HN: the planes b0, b1 of a half are adjacent n-blocks)
    //   B, C += w1*[y0 | y1]
    //   C   += w0*y2
    // A holds only the leading products, B + C the five small ones: the same two-level accumulation as the
    // six-MMA form (the epilogue adds A + (B + C) with round-to-nearest adds), at two thirds of the MMA count.
    const uint32_t id_h = umma::idesc_bf16(128, C::HN, 0, 1), id_2h = umma::idesc_bf16(128, 2 * C::HN, 0, 1);
    auto gemm_fwd_half = [&](int h) {
      const uint32_t DA = tb + h * 3 * C::HN;
      for (int ks = 0; ks < C::KS; ++ks) {
        const uint32_t lo = r1a + h * 3 * C::HP + ks * 16 * SWB;
        const uint64_t b01 = umma::smem_desc(lo, C::HP, 8 * SWB, LT);
        const uint64_t b2 = umma::smem_desc(lo + 2 * C::HP, C::HP, 8 * SWB, LT);
        const uint32_t acc = ks > 0 ? 1u : 0u;
        const int s = (int)(ci % NSLOT);
        tc::wait_bar(&bar_full[s], (uint32_t)((ci / NSLOT) & 1));
        umma::fence_after_sync();
        const uint64_t w2 = umma::smem_desc(rga + s * C::SLOT, 16, 256, 6);
        const uint64_t w1 = umma::smem_desc(rga + s * C::SLOT + C::PLANE_W, 16, 256, 6);
        const uint64_t w0 = umma::smem_desc(rga + s * C::SLOT + 2 * C::PLANE_W, 16, 256, 6);
        umma::mma_bf16_ss(DA + 2 * C::HN, w2, b01, id_h, acc);
        umma::mma_bf16_ss(DA, w0, b01, id_2h, acc);
        umma::mma_bf16_ss(DA + C::HN, w1, b01, id_2h, 1u);
        umma::mma_bf16_ss(DA + 2 * C::HN, w0, b2, id_h, 1u);
        umma::commit(&bar_empty[s]);
        ++ci;
      }
      umma::commit(&bar_fd[h]);
    };
    // DW[out][in] = sum_n G[out][n] * Y[in][n]: G in R1 (M rows = output units: the flush then writes the gradient
    // with coalesced accesses), Y block in R2 (two planes), both K-major views; five products (everything down
    // to 2^-16 of the leading term; the sum over the points averages the remaining rounding noise)
    auto gemm_wgrad = [&](int buf, int iblk) {   // iblk: block of 128 OUTPUT units (rows of G) -> TMEM lanes
      const uint32_t DW = tb + C::TC_DW + 128 * buf;
      for (int ks = 0; ks < C::KSW; ++ks) {
        uint64_t ag[3], by[2];
#pragma unroll
        for (int p = 0; p < 3; ++p)
          ag[p] = C::HALVES ? umma::smem_desc(r1a + ((ks >> 2) * 3 + p) * C::HP + (ks & 3) * 32, 16, 8 * SWB, LT)
                            : umma::smem_desc(r1a + p * C::PLANE1 + iblk * 128 * SWB + tc::kmajor_koff<SWB>(ks, WP), 16, 8 * SWB, LT);
#pragma unroll
        for (int p = 0; p < 2; ++p) by[p] = umma::smem_desc(r2a + (p % C::YP) * C::PLANE2 + tc::kmajor_koff<SWB>(ks, 128), 16, 8 * SWB, LT);
        umma::mma_bf16_ss(DW, ag[2], by[0], id_wg, ks > 0 ? 1u : 0u);
        umma::mma_bf16_ss(DW, ag[1], by[1], id_wg, 1u);
        umma::mma_bf16_ss(DW, ag[1], by[0], id_wg, 1u);
        umma::mma_bf16_ss(DW, ag[0], by[1], id_wg, 1u);
        umma::mma_bf16_ss(DW, ag[0], by[0], id_wg, 1u);
      }
      umma::commit(bar_w);
    };
    // PROF: the issuing lane also waits for each GEMM and accumulates its duration (slots 8 fwd, 9 dgrad, 10 wgrad)
    const bool mprof = PROF && (L.phase_clk != nullptr) && blockIdx.x == 0 && lane == 0;
    long long gclk[3] = {0, 0, 0};
    uint32_t mp_fd = 0, mp_w = 0, mp_y = 0;
    for (int it = 0; it < my_tiles; ++it) {
      for (int l = 1; l < Lh; ++l) {
        if (C::HALVES) {
          for (int h = 0; h < 2; ++h) {
            tc::named_sync(h ? TC_BAR_OP1B : TC_BAR_OP1, C::NEPI_T + 32);
            umma::fence_after_sync();
            if (lane == 0) gemm_fwd_half(h);
            __syncwarp();
          }
          if (PROF) mp_fd ^= 1;   // bar_fd[0] completed one more phase
        } else {
          tc::named_sync(TC_BAR_OP1, C::NEPI_T + 32);
          umma::fence_after_sync();
          const long long t0 = PROF ? clock64() : 0;
          const bool side = TRAIN && C::YSIDE;
          if (side && lane == 0)   // Y^(l-1), planes b0 | b1 -> side stash of layer l-1
            tc::bulk_s2g(L.stash + ((size_t)blockIdx.x * Lh + (l - 1)) * C::STL + C::STL_F, R1, (uint32_t)(C::YS_F * 4));
          if (lane == 0) {   // forward: two-level accumulation (loss / residual precision)
            if (C::FWD3) gemm_fwd3(side); else gemm_wx(true, C::CONCAT, side);
          }
          if (PROF && C::MB == 1 && lane == 0) { tc::wait_bar(bar_fd, mp_fd); mp_fd ^= 1; gclk[0] += clock64() - t0; }
          __syncwarp();
        }
      }
      if (TRAIN) {
        for (int l = Lh - 1; l >= 1; --l) {
          tc::named_sync(TC_BAR_OP1, C::NEPI_T + 32);
          umma::fence_after_sync();
          const long long t0 = PROF ? clock64() : 0;
          if (C::YSIDE && lane == 0) {
            // R2 is free (B1(l) waited for wgrad(l+1) before it wrote G): fetch Y^(l-1) from the side stash under dgrad(l)
            if (l == Lh - 1) tc::bulk_wait_all();   // the forward pass's last side-stash store has reached global memory
            mbar_expect_tx(bar_y, (uint32_t)(C::YS_F * 4));
            bulk_g2s(R2, L.stash + ((size_t)blockIdx.x * Lh + (l - 1)) * C::STL + C::STL_F, (uint32_t)(C::YS_F * 4), bar_y);
          }
          constexpr bool ILV = (TC_DGRAD_INTERLEAVE != 0) && (C::MB == 2) && (C::KS % 2 == 0);
          if (lane == 0 && !ILV) gemm_wx(C::DGRAD_TWO_LEVEL, C::CONCAT_DGRAD);
          __syncwarp();
          if (C::MB == 1) {
            if (C::YSIDE) {
              if (lane == 0) { tc::wait_bar(bar_y, mp_y); umma::fence_after_sync(); }
              mp_y ^= 1;
            } else {
              tc::named_sync(TC_BAR_OP2, C::NEPI_T + 32);
              umma::fence_after_sync();
            }
            const long long t1 = PROF ? clock64() : 0;
            if (lane == 0) gemm_wgrad(l & 1, 0);
            if (PROF && lane == 0) {
              tc::wait_bar(bar_fd, mp_fd); mp_fd ^= 1; gclk[1] += clock64() - t0;
              tc::wait_bar(bar_w, mp_w); mp_w ^= 1; gclk[2] += clock64() - t1;
            }
            __syncwarp();
          } else {
            // two blocks of input units (the Y block in R2 is recomputed per block) x two blocks of output units;
            // ONE weight-gradient block in TMEM: the epilogue warps drain it between the two output blocks
            // ILV: the data-gradient GEMM goes to the tensor core in four chunks, one BEHIND each weight-gradient block, so
            // that the pipe works on a chunk while the epilogue warps drain the block (and recompute the next Y block)
            // instead of idling through the drain: W(0,0) D.0 | W(1,0) D.1 | W(0,1) D.2 | W(1,1) D.3.
            for (int j = 0; j < C::MB; ++j) {
              tc::named_sync(TC_BAR_OP2, C::NEPI_T + 32);
              umma::fence_after_sync();
              if (lane == 0) { gemm_wgrad(0, 0); if (ILV) gemm_wx(C::DGRAD_TWO_LEVEL, C::CONCAT_DGRAD, false, 2 * j, 4); }
              __syncwarp();
              tc::named_sync(TC_BAR_OP3, C::NEPI_T + 32);
              umma::fence_after_sync();
              if (lane == 0) { gemm_wgrad(0, 1); if (ILV) gemm_wx(C::DGRAD_TWO_LEVEL, C::CONCAT_DGRAD, false, 2 * j + 1, 4); }
              __syncwarp();
            }
          }
        }
      }
    }
    if (mprof) { L.phase_clk[8] = gclk[0]; L.phase_clk[9] = gclk[1]; L.phase_clk[10] = gclk[2]; }
  } else {
    // ================================================================ epilogue warps
    const int quad = warp & 3, q = warp >> 2;
    const int u = 32 * quad + lane;                        // this thread's unit (= TMEM lane)
    const uint32_t tl = tb + ((uint32_t)(32 * quad) << 16);
    const int n8 = q;                                      // 8-point block of this warp
    uint32_t par_fd = 0, par_fd1 = 0, par_w = 0;   // par_fd: bar_fd[0], par_fd1: bar_fd[1] (HALVES forward)
    auto epi_sync = [&]() { tc::named_sync(TC_BAR_EPI, C::NEPI_T); };
    // 8-byte slots of this thread's four points of pass h (channel c) in plane 0 of the operand tiles; the plane stride
    // of R1 is R1_PST.  HALVES: column n = h*HN + n8*16 + c*4 + i  ->  16-byte chunk n8*2 + c/2 of half h, slot c & 1.
    constexpr int R1_PST = C::HALVES ? C::HP : C::PLANE1;
    auto r1_slot = [&](int line, int c, int h) -> uint8_t* {
      if (C::HALVES) return R1 + h * 3 * C::HP + tc::chunk_off<SWB>(line, n8 * 2 + (c >> 1), WP) + 8 * (c & 1);
      return R1 + tc::chunk_off<SWB>(line, c * PB + n8, WP) + 8 * h;
    };
    auto r2_slot = [&](int line, int c, int h) -> uint8_t* {
      if (C::HALVES) return R2 + tc::chunk_off<SWB>(line, h * 8 + n8 * 2 + (c >> 1), 128) + 8 * (c & 1);
      return R2 + tc::chunk_off<SWB>(line, c * PB + n8, 128) + 8 * h;
    };
    // Hand an operand tile to the MMA warp: writer-side generic -> async proxy fence (the conventional place), then a
    // non-blocking arrive.  fence.proxy.async compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: it also waits for this
    // thread's outstanding GLOBAL stores, which is why the stash stores of the last half are issued after it.
    // (A consumer-side fence -- legal under the PTX memory model -- was measured: same speed, 20.39 vs 20.34 ms.)
    auto operands_ready = [&](int bar) {
      umma::fence_async_smem();
      umma::fence_before_sync();
      tc::named_arrive(bar, C::NEPI_T + 32);
    };
    float* const stash = TRAIN ? (L.stash + (size_t)blockIdx.x * Lh * C::STL) : nullptr;
    // Gradient-accumulator rows shared by `share` consecutive CTAs (padded width 256: 148 private rows of 1.3 MB do not
    // fit the L2 and every flush became a DRAM read-modify-write).  The additions into a shared row stay in a FIXED order
    // -- tile round by tile round, member by member -- through one token per layer: a CTA adds the layer's weight- and
    // bias-gradient contributions only when the token says it is its turn, and passes the token on when all its
    // additions have been performed.  Members run one layer's flush apart, so nobody waits in steady state, and the
    // gradient stays bit-reproducible.  (The launch is cooperative: all members are resident.)
    const int share = L.gacc_share > 1 ? L.gacc_share : 1;
    const bool shared_rows = TRAIN && share > 1;
    const int grp = blockIdx.x / share, mem = blockIdx.x % share;
    const int kcount = ((int)gridDim.x - grp * share) < share ? ((int)gridDim.x - grp * share) : share;
    float* const gacc = TRAIN ? (L.gacc + (size_t)grp * net.pg) : nullptr;
    unsigned* const tok = shared_rows ? (L.gacc_token + (size_t)grp * PINN_TOKENS) : nullptr;
    auto gadd = [&](float* p, float v) { if (shared_rows) tc::red_add(p, v); else *p += v; };
    auto token_wait = [&](int e, int round) {
      if (!shared_rows) return;
      if (lane == 0) {
        const unsigned want = (unsigned)(round * kcount + mem);
        bool ok = false;
        for (int i = 0; i < (1 << 24); ++i) {
          if (tc::ld_acquire_gpu(tok + e) == want) { ok = true; break; }
          __nanosleep(64);
        }
        if (!ok) __trap();   // a protocol bug traps (launch failure) instead of hanging the GPU
      }
      __syncwarp();
    };
    // every epilogue thread has fenced its additions and passed an epi_sync before thread 0 calls this
    auto token_pass = [&](int e, int round) {
      if (shared_rows && tid == 0) tc::st_release_gpu(tok + e, (unsigned)(round * kcount + mem + 1));
    };
    int pending_tok = -1, pending_round = 0;   // layer whose token this CTA still holds (all epilogue threads agree)
    auto pass_pending = [&]() {
      if (pending_tok < 0) return;
      if (shared_rows) __threadfence();
      epi_sync();
      token_pass(pending_tok, pending_round);
      pending_tok = -1;
    };
    const int ldw = L.ldw;
    float wlacc[C::MB], w0acc[C::MB][3], blacc = 0.f;   // this thread's units: u, u + 128 (one per M block)
#pragma unroll
    for (int mb = 0; mb < C::MB; ++mb) { wlacc[mb] = 0.f; w0acc[mb][0] = w0acc[mb][1] = w0acc[mb][2] = 0.f; }
    double lcur = 0.0;
    const int slot = L.seg_slot[0];
    const long long n_end = L.seg_pt_end[0];
    // phase clocks (PROF instantiation only: the counters live in local memory, which this kernel's 225 KB of
    // shared memory leaves no L1 for -- in the production instantiation lap() is a no-op)
    const bool prof = PROF && (L.phase_clk != nullptr) && blockIdx.x == 0 && tid == 0;
    long long pclk[PROF ? 8 : 1];
    long long tmark = 0;
    if (PROF) {
#pragma unroll
      for (int i = 0; i < 8; ++i) pclk[i] = 0;
      tmark = prof ? clock64() : 0;
    }
    auto lap = [&](int ph) {
      if (PROF) {
        if (prof) { const long long now = clock64(); pclk[ph] += now - tmark; tmark = now; }
      }
    };
    // accumulators of the last forward / data-gradient GEMM (big + small), points 4h..4h+3 of the block
    auto load_acc = [&](float (&a)[K][V], int h, int mb) {
      if constexpr (C::HALVES) {   // the thread's 16 values of the pass are 16 consecutive columns
        float big[16], sm[16];
        umma::tmem_ld16(tl + C::TC_D(0) + h * C::HN + n8 * 16, big);
        umma::tmem_ld16(tl + C::TC_D(0) + NROW + h * C::HN + n8 * 16, sm);
        umma::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
          for (int i = 0; i < V; ++i) a[c][i] = big[c * V + i] + sm[c * V + i];
      } else {
        float sm[K][V];
#pragma unroll
        for (int c = 0; c < K; ++c) {
          umma::tmem_ld4(tl + C::TC_D(mb) + c * NP + 8 * n8 + V * h, a[c]);
          umma::tmem_ld4(tl + C::TC_D(mb) + NROW + c * NP + 8 * n8 + V * h, sm[c]);
        }
        umma::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
          for (int i = 0; i < V; ++i) a[c][i] += sm[c][i];
      }
    };
    // FWD3 forward accumulators: blocks A | B | C of NROW columns per M block; result A + (B + C)
    auto load_acc_fwd3 = [&](float (&a)[K][V], int h, int mb) {
      float sb[K][V], sc[K][V];
      const uint32_t base = tl + mb * 3 * NROW + 8 * n8 + V * h;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        umma::tmem_ld4(base + NROW + c * NP, sb[c]);
        umma::tmem_ld4(base + 2 * NROW + c * NP, sc[c]);
        umma::tmem_ld4(base + c * NP, a[c]);
      }
      umma::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < K; ++c)
#pragma unroll
        for (int i = 0; i < V; ++i) a[c][i] += sb[c][i] + sc[c][i];
    };
    // HALVES forward accumulators of half h: blocks A | B | C of HN columns; result A + (B + C)
    auto load_acc_fwd = [&](float (&a)[K][V], int h) {
      float A[16], B[16], Cc[16];
      const uint32_t base = tl + h * 3 * C::HN + n8 * 16;
      umma::tmem_ld16(base + C::HN, B);
      umma::tmem_ld16(base + 2 * C::HN, Cc);
      umma::tmem_ld16(base, A);
      umma::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < K; ++c)
#pragma unroll
        for (int i = 0; i < V; ++i) a[c][i] = A[c * V + i] + (B[c * V + i] + Cc[c * V + i]);
    };
    // data-gradient GEMM with ONE accumulator: half the TMEM read traffic of B1 (TMEM reads run at ~64 B/cycle/SM:
    // measured 1.1 k cycles per half-pass for the two-block read)
    auto load_acc1 = [&](float (&a)[K][V], int h, int mb) {
      static_assert(!C::HALVES || C::DGRAD_TWO_LEVEL, "HALVES uses the two-level data-gradient accumulators");
#pragma unroll
      for (int c = 0; c < K; ++c) umma::tmem_ld4(tl + C::TC_D(mb) + c * NP + 8 * n8 + V * h, a[c]);
      umma::tmem_ld_wait();
    };
    auto load_beta = [&](float (&beta)[3][V], int h) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (C::LAP && k < C::N1) tc::ld4(s_beta + k * NP + 8 * n8 + V * h, beta[k]);
        else {
#pragma unroll
          for (int i = 0; i < V; ++i) beta[k][i] = 0.f;
        }
      }
    };
    // Layer 0 has no stash: its pre-activation jets are three FMAs per value away from the tile's feature jets
    // (shared memory), so the backward pass recomputes them (and the activation) instead of a global round trip.
    // st[0] = y, st[c > 0] = pre-activation jets, cs = cos for the sin activation -- the layout of a stash read.
    auto layer0_stash = [&](float (&st)[K][V], float (&cs)[V], int h, int uu) {
      const float w00 = __ldg(L.wpack + net.off_w0 + uu), w01 = __ldg(L.wpack + net.off_w0 + WP + uu),
                  w02 = __ldg(L.wpack + net.off_w0 + 2 * WP + uu);
      const float bias = __ldg(L.wpack + net.off_b[0] + uu);
#pragma unroll
      for (int i = 0; i < V; ++i)
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const float4 hh = *reinterpret_cast<const float4*>(s_hj + ((8 * n8 + V * h + i) * K + c) * 4);
          st[c][i] = net.scl * fmaf(hh.x, w00, fmaf(hh.y, w01, hh.z * w02));
        }
      float a0[V], y[V], d2[V];
#pragma unroll
      for (int i = 0; i < V; ++i) a0[i] = st[0][i] + bias;
      tc::act_fwd<V>(net.act_first, a0, y, cs, d2);
#pragma unroll
      for (int i = 0; i < V; ++i) st[0][i] = y[i];
    };
    // stash slot of (layer, channel, half); channel K = cos of the sin activation
    auto stash_ptr = [&](int l, int c, int h, int mb) -> float* {
      return stash + (size_t)l * C::STL + ((size_t)((mb * (K + 1) + c) * Q + q) * 128 + u) * 8 + V * h;
    };
    // weight-gradient block (lane = output unit, columns = input units) -> CTA-private accumulator rows, coalesced
    auto flush_dw = [&](int l, int buf, int iblk, int jblk) {   // block (output units 128*iblk.., input units 128*jblk..)
      if ((PROF && (L.exp_flags & 1)) || (TC_EXP & 1)) return;   // experiment: no accumulator flush (timing only, wrong gradient)
      constexpr int NC = 128 / Q;  // input units (columns) per warp
      float* gcol = gacc + net.off_w[l] + (size_t)(128 * jblk + q * NC) * ldw + 128 * iblk + u;
      const uint32_t src = tl + C::TC_DW + 128 * buf + q * NC;
      // The block is ADDED to the CTA-private accumulator rows with red.global.add.f32: the L2 performs the
      // read-modify-write, nothing returns to the SM and the warps do not wait for a round trip (measured on C4:
      // the load + add + store version cost 17 % of the step).  Each address is only ever updated by this one
      // thread of this one CTA, tile after tile, so the order of the additions -- and with it every bit of the
      // gradient -- is fixed by program order.
#pragma unroll
      for (int b = 0; b < NC / 16; ++b) {
        float w[16];
        umma::tmem_ld16(src + 16 * b, w);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(gcol + (size_t)(16 * b + i) * ldw), "f"(w[i]) : "memory");
      }
      umma::fence_before_sync();
    };

#pragma unroll 1
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const long long pbegin = L.seg_pt_begin[0] + (long long)tile * NP;
      // ---------------- tile header: coordinates, betas, feature jets of the tile's points
      if (tid < NP) {
        const int pt = tid;
        const bool valid = pbegin + pt < n_end;
        const long long gp = valid ? pbegin + pt : pbegin;
        const float* zp = L.coords + gp * net.d_in;
        float z[3];
        z[0] = __ldg(zp);
        z[1] = (net.d_in > 1) ? __ldg(zp + 1) : 0.f;
        z[2] = (net.d_in > 2) ? __ldg(zp + 2) : 0.f;
        float beta[3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          beta[i] = (C::LAP && net.lap_aux[i] >= 0) ? __ldg(L.aux + gp * L.n_aux + net.lap_aux[i]) : net.lap_beta[i];
        float hj[K][3];
        feature_jets<C>(net, z, beta, hj);
        s_z[pt * 4 + 0] = z[0]; s_z[pt * 4 + 1] = z[1]; s_z[pt * 4 + 2] = z[2]; s_z[pt * 4 + 3] = valid ? 1.f : 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) s_beta[i * NP + pt] = beta[i];
#pragma unroll
        for (int c = 0; c < K; ++c)
          *reinterpret_cast<float4*>(s_hj + (pt * K + c) * 4) = make_float4(hj[c][0], hj[c][1], hj[c][2], 0.f);
      }
      epi_sync();

      // ---------------- forward
#pragma unroll 1
      for (int l = 0; l < Lh; ++l) {
        const int act = (l == 0) ? net.act_first : net.act_hidden;
        const bool last = (l == Lh - 1);
        if (l > 0 && (!C::HALVES || last)) {
          // (HALVES, last layer: the output-layer partial products alias R1, so both halves of the GEMM must be done)
          tc::wait_bar(&bar_fd[0], par_fd);
          par_fd ^= 1;
          if (C::HALVES) { tc::wait_bar(&bar_fd[1], par_fd1); par_fd1 ^= 1; }
          umma::fence_after_sync();
          lap(0);
        }
#pragma unroll 1
        for (int mb = 0; mb < C::MB; ++mb) {
        const int uu = u + 128 * mb;                       // unit of this pass (= line of the operand tiles)
        const float bias = __ldg(L.wpack + net.off_b[l] + uu);
        const float wlv = last ? __ldg(L.wpack + net.off_wl + uu) : 0.f;
        float w00 = 0.f, w01 = 0.f, w02 = 0.f;
        if (l == 0) {
          w00 = __ldg(L.wpack + net.off_w0 + uu); w01 = __ldg(L.wpack + net.off_w0 + WP + uu); w02 = __ldg(L.wpack + net.off_w0 + 2 * WP + uu);
        }
        float sv[K + 1][V];   // stash values of the half in flight
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
          float a[K][V];
          if (l == 0) {
#pragma unroll
            for (int i = 0; i < V; ++i)
#pragma unroll
              for (int c = 0; c < K; ++c) {
                const float4 hh = *reinterpret_cast<const float4*>(s_hj + ((8 * n8 + V * h + i) * K + c) * 4);
                a[c][i] = net.scl * fmaf(hh.x, w00, fmaf(hh.y, w01, hh.z * w02));
              }
          } else if (C::HALVES) {
            if (!last) {   // forward GEMM of this half (the other half's may still be running)
              if (h == 0) { tc::wait_bar(&bar_fd[0], par_fd); par_fd ^= 1; }
              else { tc::wait_bar(&bar_fd[1], par_fd1); par_fd1 ^= 1; }
              umma::fence_after_sync();
              lap(0);
            }
            load_acc_fwd(a, h);
          } else if (C::FWD3) {
            load_acc_fwd3(a, h, mb);
          } else {
            load_acc(a, h, mb);
          }
          float beta[3][V];
          load_beta(beta, h);
          {
            float y[V], d1[V], d2[V];
#pragma unroll
            for (int i = 0; i < V; ++i) a[0][i] += bias;
            tc::act_fwd<V>(act, a[0], y, d1, d2);
            if (TRAIN) {
              // stash: y, (cos), pre-activation jets.  The global stores of the LAST half are issued after the
              // operand hand-over below: fence.proxy.async is a MEMBAR that would wait for them to drain
#pragma unroll
              for (int i = 0; i < V; ++i) { sv[0][i] = y[i]; sv[K][i] = d1[i]; }
#pragma unroll
              for (int c = 1; c < K; ++c)
#pragma unroll
                for (int i = 0; i < V; ++i) sv[c][i] = a[c][i];
            }
            tc::jets_outputs<C, V>(a, y, d1, d2, beta);
          }
          if (!last) {
#pragma unroll
            for (int c = 0; c < K; ++c) tc::store_split4p<3>(r1_slot(uu, c, h), R1_PST, a[c]);
          } else {
#pragma unroll
            for (int c = 0; c < K; ++c) {
              float pv[V];
#pragma unroll
              for (int i = 0; i < V; ++i) pv[i] = a[c][i] * wlv;
              tc::st4(part + (size_t)uu * C::PARTLD + c * NP + 8 * n8 + V * h, pv);
            }
          }
          if (C::HALVES) {
            if (!last) { operands_ready(h ? TC_BAR_OP1B : TC_BAR_OP1); lap(1); }   // this half may go to the tensor core
          } else if (TRAIN) {
            if (mb == C::MB - 1 && h == NH - 1 && !last) operands_ready(TC_BAR_OP1);
          }
          if (TRAIN) {
            if (!(TC_EXP & 4) && l > 0) {
#pragma unroll
              for (int c = 0; c < K; ++c) tc::st4(stash_ptr(l, c, h, mb), sv[c]);
              if (act == PINN_SIN) tc::st4(stash_ptr(l, K, h, mb), sv[K]);
            }
          }
        }
        }  // mb
        if (!TRAIN && !last && !C::HALVES) operands_ready(TC_BAR_OP1);
        lap(1);
      }

      // ---------------- output layer (fold over units) + residual program
      epi_sync();
      for (int idx = tid; idx < 4 * NROW; idx += C::NEPI_T) {
        const int n = idx % NROW, qq = idx / NROW;
        const float* pr = part + (size_t)(qq * (WP / 4)) * C::PARTLD + n;
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < WP / 4; ++r) s += pr[(size_t)r * C::PARTLD];
        s_psum[qq * NROW + n] = s;
      }
      epi_sync();
      if (warp == 0) {
        const int pt = lane < NP ? lane : NP - 1;
        const bool valid = lane < NP && s_z[pt * 4 + 3] != 0.f;
        const long long gp = valid ? pbegin + pt : pbegin;
        float z[1][3] = {{s_z[pt * 4], s_z[pt * 4 + 1], s_z[pt * 4 + 2]}};
        float uo[K][1], f[1], df[K][1];
        const float* auxp[1] = {L.aux ? (L.aux + gp * L.n_aux) : nullptr};
        const float bl = __ldg(L.wpack + net.off_bl);
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const int n = c * NP + pt;
          float s = (s_psum[n] + s_psum[NROW + n]) + (s_psum[2 * NROW + n] + s_psum[3 * NROW + n]);
          s = net.epsil * (s + (c == 0 ? bl : 0.f));
          if (L.base) s += __ldg(L.base + gp * K + c);
          uo[c][0] = s;
        }
        vm_run<K, 1>(s_ops, L.prog.n_ops, s_consts, z, auxp, uo, f, df);
        if (TRAIN) {
          const float sc = valid ? __ldg(L.seg_scale + slot) : 0.f;
          if (lane < NP) {
#pragma unroll
            for (int c = 0; c < K; ++c) s_ubar[c * NP + pt] = net.epsil * sc * f[0] * df[c][0];
          }
          if (valid) {
            lcur += (double)f[0] * (double)f[0];
            blacc += net.epsil * sc * f[0] * df[0][0];
          }
        } else if (valid) {
          if (L.out_u) L.out_u[gp] = uo[0][0];
          if (L.out_f) L.out_f[gp] = f[0];
          if (L.out_jets) {
#pragma unroll
            for (int c = 0; c < K; ++c) L.out_jets[gp * K + c] = uo[c][0];
          }
        }
      }
      epi_sync();
      lap(2);

      // ---------------- backward.  Per layer l the epilogue warps run B1(l) (adjoint of the pre-activations -> G),
      // B2(l) (the layer's input jets again -> Y) and the flush of the PREVIOUS layer's weight-gradient block,
      // while the tensor core works on dgrad(l) and wgrad(l) (ping-pong weight-gradient blocks in TMEM).
      if (TRAIN && C::MB == 1) {
        const float wlv = __ldg(L.wpack + net.off_wl + u);
#pragma unroll 1
        for (int l = Lh - 1; l >= 0; --l) {
          const int act = (l == 0) ? net.act_first : net.act_hidden;
          // stash of BOTH halves in flight before the wait for dgrad(l+1): the L2 round trip hides behind it.
          // (Keeping the values B2(l+1) loaded in registers instead -- one L2 read per layer -- was measured
          // slower: the 40 extra live registers spill, and this kernel's shared-memory footprint leaves no L1.)
          float stA[NH][K][V], csA[NH][V];
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            if (TC_EXP & 8) {
#pragma unroll
              for (int c = 0; c < K; ++c)
#pragma unroll
                for (int i = 0; i < V; ++i) { stA[h][c][i] = 0.25f; csA[h][i] = 0.5f; }
            } else if (l == 0) {
              layer0_stash(stA[h], csA[h], h, u);
            } else {
#pragma unroll
              for (int c = 0; c < K; ++c) tc::ld4(stash_ptr(l, c, h, 0), stA[h][c]);
              if (act == PINN_SIN) tc::ld4(stash_ptr(l, K, h, 0), csA[h]);
            }
          }
          constexpr bool MERGE = (TC_B2MERGE != 0) && !C::YSIDE;
          if (MERGE && l < Lh - 1) {
            // ---- Y^l (the layer's output jets, two planes) -> R2 for wgrad(l+1).  R2 is free: B1(l+1) waited for wgrad(l+2).
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              float o[K][V], y[V], d1[V], d2[V], d3[V], beta[3][V];
#pragma unroll
              for (int i = 0; i < V; ++i) { y[i] = stA[h][0][i]; d1[i] = csA[h][i]; }
#pragma unroll
              for (int c = 0; c < K; ++c)
#pragma unroll
                for (int i = 0; i < V; ++i) o[c][i] = stA[h][c][i];
              load_beta(beta, h);
              tc::act_bwd<V, false>(act, y, d1, d2, d3);
              tc::jets_outputs<C, V>(o, y, d1, d2, beta);
#pragma unroll
              for (int c = 0; c < K; ++c) tc::store_split4p<C::YP>(r2_slot(u, c, h), C::PLANE2, o[c]);
            }
            operands_ready(TC_BAR_OP2);
            lap(4);
          }
          if (l < Lh - 1) {
            tc::wait_bar(bar_fd, par_fd);  // dgrad(l+1)
            par_fd ^= 1;
            umma::fence_after_sync();
            lap(7);
          }
          float gb = 0.f;
          // ---- B1(l)
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            float yb[K][V];
            float d1[V], d2[V], d3[V], beta[3][V];
            float (&st)[K][V] = stA[h];
            if (act == PINN_SIN) {
#pragma unroll
              for (int i = 0; i < V; ++i) d1[i] = csA[h][i];
            }
            load_beta(beta, h);
            if (l < Lh - 1) { if (C::DGRAD_TWO_LEVEL) load_acc(yb, h, 0); else load_acc1(yb, h, 0); }
            tc::act_bwd<V, true>(act, st[0], d1, d2, d3);
            if (l == Lh - 1) {
              // seeds: ybar[c] = (epsil * ubar_c) * wl[u]; the output-layer weight gradient needs the layer outputs
#pragma unroll
              for (int c = 0; c < K; ++c) {
                float ub[V];
                tc::ld4(s_ubar + c * NP + 8 * n8 + V * h, ub);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                  float o;
                  if (c == 0) o = st[0][i];
                  else if (c <= C::N1) o = d1[i] * st[c][i];
                  else if (c <= C::N1 + C::N2) {
                    const float Ai = st[c - C::N1][i];
                    o = fmaf(d2[i] * Ai, Ai, d1[i] * st[c][i]);
                  } else if (C::MIX == 1) o = fmaf(d2[i] * st[1][i], st[2][i], d1[i] * st[c][i]);
                  else {
                    float S = 0.f;
#pragma unroll
                    for (int k = 0; k < C::N1; ++k) S = fmaf(beta[k][i] * st[1 + k][i], st[1 + k][i], S);
                    o = fmaf(d2[i], S, d1[i] * st[c][i]);
                  }
                  wlacc[0] = fmaf(ub[i], o, wlacc[0]);
                  yb[c][i] = ub[i] * wlv;
                }
              }
            }
            tc::jets_adjoint<C, V>(yb, st, d1, d2, d3, beta);
#pragma unroll
            for (int i = 0; i < V; ++i) gb += yb[0][i];
            if (l > 0) {
              if (h == 0 && l < Lh - 1) {
                lap(3);
                tc::wait_bar(bar_w, par_w);  // wgrad(l+1) has read G^(l+1) (R1) and Y^l (R2)
                par_w ^= 1;
                umma::fence_after_sync();
                lap(5);
              }
#pragma unroll
              for (int c = 0; c < K; ++c) tc::store_split4p<3>(r1_slot(u, c, h), R1_PST, yb[c]);
            } else {
              if (MERGE && h == 0 && Lh > 1) {
                lap(3);
                tc::wait_bar(bar_w, par_w);  // wgrad(1): its block is flushed below
                par_w ^= 1;
                umma::fence_after_sync();
                lap(5);
              }
#pragma unroll
              for (int i = 0; i < V; ++i)
#pragma unroll
                for (int c = 0; c < K; ++c) {
                  const float4 hh = *reinterpret_cast<const float4*>(s_hj + ((8 * n8 + V * h + i) * K + c) * 4);
                  const float t = net.scl * yb[c][i];
                  w0acc[0][0] = fmaf(hh.x, t, w0acc[0][0]);
                  w0acc[0][1] = fmaf(hh.y, t, w0acc[0][1]);
                  w0acc[0][2] = fmaf(hh.z, t, w0acc[0][2]);
                }
            }
          }
          s_bg[q * WP + u] = gb;
          if (l > 0) operands_ready(TC_BAR_OP1);
          if (TC_DISCARD && l > 0 && lane < 8) {
            // the stash of layer l is dead (B2(l+1) and B1(l) were its readers): 8 lines per channel and warp
            const int nch = K + (act == PINN_SIN ? 1 : 0);
            for (int c = 0; c < nch; ++c) tc::discard_l2(stash + (size_t)l * C::STL + ((size_t)(c * Q + q) * 128 + 32 * quad) * 8 + lane * 32);
          }
          lap(3);
          if (MERGE) {
            // ---- F(l+1): the block of wgrad(l+1) (complete: B1(l) waited for it), under dgrad(l)
            if (l < Lh - 1) flush_dw(l + 1, (l + 1) & 1, 0, 0);
            lap(6);
          } else if (l > 0) {
            // ---- F(l+1): flush the previous layer's weight-gradient block (wgrad(l+1) completed: B1 waited for it)
            // first -- global traffic only, dgrad(l) has the shared-memory bandwidth to itself meanwhile
            if (l < Lh - 1) flush_dw(l + 1, (l + 1) & 1, 0, 0);
            lap(6);
            // ---- B2(l): the layer's input jets Y^(l-1) again (from the stash of layer l-1) -> R2 (two planes)
            // (YSIDE: the MMA warp loads them from the side stash instead.  Issuing B2's stash reads before the flush, to hide
            // their L2 round trip behind it, was measured slower: 40 more live registers across the flush spill, 22.4 vs 20.2 ms.)
            const int actp = (l - 1 == 0) ? net.act_first : net.act_hidden;
#pragma unroll 1
            for (int h = 0; h < (C::YSIDE ? 0 : NH); ++h) {
              float st[K][V], d1[V], d2[V], d3[V], beta[3][V];
              if ((PROF && (L.exp_flags & 2)) || (TC_EXP & 2)) {  // experiment: B2 without its stash read (timing only)
#pragma unroll
                for (int c = 0; c < K; ++c)
#pragma unroll
                  for (int i = 0; i < V; ++i) { st[c][i] = 0.25f; d1[i] = 0.5f; }
              } else if (l == 1) {
                layer0_stash(st, d1, h, u);
              } else {
#pragma unroll
                for (int c = 0; c < K; ++c) tc::ld4(stash_ptr(l - 1, c, h, 0), st[c]);
                if (actp == PINN_SIN) tc::ld4(stash_ptr(l - 1, K, h, 0), d1);
              }
              load_beta(beta, h);
              tc::act_bwd<V, false>(actp, st[0], d1, d2, d3);
              {
                float y[V];
#pragma unroll
                for (int i = 0; i < V; ++i) y[i] = st[0][i];
                tc::jets_outputs<C, V>(st, y, d1, d2, beta);
              }
#pragma unroll
              for (int c = 0; c < K; ++c) tc::store_split4p<C::YP>(r2_slot(u, c, h), C::PLANE2, st[c]);
            }
            if (!C::YSIDE) operands_ready(TC_BAR_OP2);
            lap(4);
          } else if (Lh > 1) {
            tc::wait_bar(bar_w, par_w);  // wgrad(1)
            par_w ^= 1;
            umma::fence_after_sync();
            lap(5);
            flush_dw(1, 1, 0, 0);
            lap(6);
          }
          // ---- bias gradient of layer l: fixed-order fold over the Q point blocks
          epi_sync();
          if (q == 0) {
            float g = 0.f;
#pragma unroll
            for (int qq = 0; qq < Q; ++qq) g += s_bg[qq * WP + u];
            gacc[net.off_b[l] + u] += g;
          }
          epi_sync();
        }
      }
      // ---------------- backward, two M blocks (padded width 256).  Each thread owns units u and u + 128.  TMEM holds the
      // accumulators of both blocks (4*NROW columns) and ONE weight-gradient block, so per layer: B1 (both units) ->
      // dgrad; then per block j of INPUT units: B2 (Y block j -> R2) -> wgrad(output block 0, j) -> flush ->
      // wgrad(output block 1, j) -> flush.
      if (TRAIN && C::MB == 2) {
#pragma unroll 1
        for (int l = Lh - 1; l >= 0; --l) {
          const int act = (l == 0) ? net.act_first : net.act_hidden;
          if (l < Lh - 1) {
            tc::wait_bar(bar_fd, par_fd);  // dgrad(l+1)
            par_fd ^= 1;
            umma::fence_after_sync();
            lap(7);
          }
          // ---- B1(l)
#pragma unroll
          for (int mb = 0; mb < C::MB; ++mb) {
            const int uu = u + 128 * mb;
            const float wlv = __ldg(L.wpack + net.off_wl + uu);
            float gb = 0.f;
#pragma unroll 1
            for (int h = 0; h < NH; ++h) {
              float st[K][V], yb[K][V], d1[V], d2[V], d3[V], beta[3][V];
              if (l == 0) {
                layer0_stash(st, d1, h, uu);
              } else {
#pragma unroll
                for (int c = 0; c < K; ++c) tc::ld4(stash_ptr(l, c, h, mb), st[c]);
                if (act == PINN_SIN) tc::ld4(stash_ptr(l, K, h, mb), d1);
              }
              load_beta(beta, h);
              if (l < Lh - 1) { if (C::DGRAD_TWO_LEVEL) load_acc(yb, h, mb); else load_acc1(yb, h, mb); }
              tc::act_bwd<V, true>(act, st[0], d1, d2, d3);
              if (l == Lh - 1) {
#pragma unroll
                for (int c = 0; c < K; ++c) {
                  float ub[V];
                  tc::ld4(s_ubar + c * NP + 8 * n8 + V * h, ub);
#pragma unroll
                  for (int i = 0; i < V; ++i) {
                    float o;
                    if (c == 0) o = st[0][i];
                    else if (c <= C::N1) o = d1[i] * st[c][i];
                    else if (c <= C::N1 + C::N2) {
                      const float Ai = st[c - C::N1][i];
                      o = fmaf(d2[i] * Ai, Ai, d1[i] * st[c][i]);
                    } else if (C::MIX == 1) o = fmaf(d2[i] * st[1][i], st[2][i], d1[i] * st[c][i]);
                    else {
                      float S = 0.f;
#pragma unroll
                      for (int k = 0; k < C::N1; ++k) S = fmaf(beta[k][i] * st[1 + k][i], st[1 + k][i], S);
                      o = fmaf(d2[i], S, d1[i] * st[c][i]);
                    }
                    wlacc[mb] = fmaf(ub[i], o, wlacc[mb]);
                    yb[c][i] = ub[i] * wlv;
                  }
                }
              }
              tc::jets_adjoint<C, V>(yb, st, d1, d2, d3, beta);
#pragma unroll
              for (int i = 0; i < V; ++i) gb += yb[0][i];
              if (l > 0) {
#pragma unroll
                for (int c = 0; c < K; ++c) tc::store_split4<3>(R1 + tc::chunk_off<SWB>(uu, c * PB + n8, WP), C::PLANE1, h, yb[c]);
              } else {
#pragma unroll
                for (int i = 0; i < V; ++i)
#pragma unroll
                  for (int c = 0; c < K; ++c) {
                    const float4 hh = *reinterpret_cast<const float4*>(s_hj + ((8 * n8 + V * h + i) * K + c) * 4);
                    const float t = net.scl * yb[c][i];
                    w0acc[mb][0] = fmaf(hh.x, t, w0acc[mb][0]);
                    w0acc[mb][1] = fmaf(hh.y, t, w0acc[mb][1]);
                    w0acc[mb][2] = fmaf(hh.z, t, w0acc[mb][2]);
                  }
              }
            }
            s_bg[q * WP + uu] = gb;
          }
          if (l > 0) operands_ready(TC_BAR_OP1);
          if (TC_DISCARD && l > 0 && lane < 8) {
            const int nch = K + (act == PINN_SIN ? 1 : 0);
            for (int mb = 0; mb < C::MB; ++mb)
              for (int c = 0; c < nch; ++c)
                tc::discard_l2(stash + (size_t)l * C::STL + ((size_t)((mb * (K + 1) + c) * Q + q) * 128 + 32 * quad) * 8 + lane * 32);
          }
          lap(3);
          if (l > 0) {
            const int actp = (l - 1 == 0) ? net.act_first : net.act_hidden;
#pragma unroll 1
            for (int j = 0; j < C::MB; ++j) {
              // ---- B2(l), input-unit block j: Y^(l-1) of unit u + 128 j -> line u of R2
#pragma unroll 1
              for (int h = 0; h < NH; ++h) {
                float st[K][V], d1[V], d2[V], d3[V], beta[3][V];
                if (l == 1) {
                  layer0_stash(st, d1, h, u + 128 * j);
                } else {
#pragma unroll
                  for (int c = 0; c < K; ++c) tc::ld4(stash_ptr(l - 1, c, h, j), st[c]);
                  if (actp == PINN_SIN) tc::ld4(stash_ptr(l - 1, K, h, j), d1);
                }
                load_beta(beta, h);
                tc::act_bwd<V, false>(actp, st[0], d1, d2, d3);
                {
                  float y[V];
#pragma unroll
                  for (int i = 0; i < V; ++i) y[i] = st[0][i];
                  tc::jets_outputs<C, V>(st, y, d1, d2, beta);
                }
#pragma unroll
                for (int c = 0; c < K; ++c) tc::store_split4<C::YP>(R2 + tc::chunk_off<SWB>(u, c * PB + n8, 128), C::PLANE2, h, st[c]);
              }
              operands_ready(TC_BAR_OP2);
              lap(4);
#pragma unroll 1
              for (int i = 0; i < C::MB; ++i) {
                tc::wait_bar(bar_w, par_w);  // wgrad(output block i, input block j)
                par_w ^= 1;
                umma::fence_after_sync();
                lap(5);
                if (i == 0 && j == 0) { pass_pending(); token_wait(l, it); }
                flush_dw(l, 0, i, j);
                if (i == 0) tc::named_arrive(TC_BAR_OP3, C::NEPI_T + 32);  // block drained: the next one may be issued
                lap(6);
              }
            }
          }
          // ---- bias gradients of layer l: fixed-order fold over the Q point blocks
          if (l == 0) pass_pending();
          epi_sync();
          if (q == 0) {
#pragma unroll
            for (int mb = 0; mb < C::MB; ++mb) {
              float g = 0.f;
#pragma unroll
              for (int qq = 0; qq < Q; ++qq) g += s_bg[qq * WP + u + 128 * mb];
              if (l == 0 && mb == 0) token_wait(0, it);   // (layer 0 has no weight-gradient flush that took the token)
              gadd(gacc + net.off_b[l] + u + 128 * mb, g);
            }
          }
          // The token of layer l is passed on when this CTA reaches the flush of layer l-1 (pass_pending): by then the
          // layer's 256 KB of `red` additions have drained and the fence in front of the hand-over costs nothing (an
          // immediate fence waited ~10 k cycles per layer).  Layer 0 (bias only) hands over at once.
          if (l == 0) {
            if (shared_rows) __threadfence();
            epi_sync();
            token_pass(0, it);
          } else {
            epi_sync();
            pending_tok = l; pending_round = it;
          }
        }
      }
    }

    // ---------------- per-CTA epilogue: fold the register accumulators in fixed order
    if (TRAIN) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int mb = 0; mb < C::MB; ++mb) s_bg[q * WP + u + 128 * mb] = (r == 0) ? wlacc[mb] : w0acc[mb][r - 1];
        epi_sync();
        if (q == 0) {
#pragma unroll
          for (int mb = 0; mb < C::MB; ++mb) {
            float g = 0.f;
#pragma unroll
            for (int qq = 0; qq < Q; ++qq) g += s_bg[qq * WP + u + 128 * mb];
            const int dst = ((r == 0) ? net.off_wl : net.off_w0 + (r - 1) * WP) + u + 128 * mb;
            if (r == 0 && mb == 0) token_wait(Lh, 0);
            gadd(gacc + dst, g);
          }
        }
        epi_sync();
      }
      if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lcur += __shfl_xor_sync(0xffffffffu, lcur, o);
          blacc += __shfl_xor_sync(0xffffffffu, blacc, o);
        }
        if (lane == 0) {
          L.loss_part[(size_t)blockIdx.x * L.n_slots + slot] += lcur;
          gadd(gacc + net.off_bl, blacc);   // (warp 0 has q == 0: it holds the token of the final fold)
        }
      }
      if (shared_rows) {
        __threadfence();
        epi_sync();
        token_pass(Lh, 0);
      }
    }
    if (PROF) {
      if (prof) {
#pragma unroll
        for (int i = 0; i < 8; ++i) L.phase_clk[i] = pclk[i];
      }
    }
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == C::NEPI) umma::tmem_free(tb, 512);
}
