// tcgen05 kernel family (experimental): the hidden-layer GEMMs run on tcgen05.mma kind::tf32 with the
// activation operand in Tensor Memory (TS form) and the pre-split weights as K-major shared-memory
// operands; 3xTF32 products (lo*hi + hi*lo + hi*hi), fp32 accumulate in TMEM.
//
// Tile = 32 points.  MMA row = 32*channel + point, so warp c of the CTA (TMEM lane quadrant c) owns jet
// channel c of the tile's 32 points and lane p owns point p: every dot product over units is serial in
// a thread (no shuffles), the coupling between channels (sigma', sigma'' from the value channel, the
// squared first derivatives for the second-order channel) goes through a small shared-memory exchange.
#include "jet_umma.h"

#include "jet_kernel.cuh"
#include "umma_common.cuh"

namespace {

struct UCfg {
  static constexpr int K = 4, N1 = 2, N2 = 0, MIX = 2;
  static constexpr bool LAP = true;
};
constexpr int UW = 64;    // padded width
constexpr int UK = 4;     // jet channels
constexpr int UTP = 32;   // points per tile
constexpr int UCH = 16;   // units per epilogue chunk
constexpr int UIMG = UW * UW;
constexpr int UMAXL = 3;  // hidden GEMM layers whose operand images stay resident in shared memory
// TMEM columns of a tile
// D holds two column blocks: [0,64) = (hi + lo) * W_hi, [64,128) = (hi + lo) * W_lo (summed by the epilogue)
constexpr uint32_t TC_D = 0, TC_AHI = 128, TC_ALO = 192, TC_COLS = 256;

__device__ __forceinline__ void split_rn(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = __fadd_rn(x, -hi);
}

// images per hidden GEMM layer l = 1..L-1 (2*UIMG floats each): [0] forward operand, rows n = out unit
// (hi plane rows 0..63, lo plane rows 64..127), K = in unit; [1] data-gradient operand, rows n = in unit
// (hi | lo), K = out unit.  Un-swizzled K-major with 128 rows: word (k/4)*512 + 4*row + k%4.
// One N = 128 MMA chain then yields A*[W_hi | W_lo] (two of the three split products) in one pass.
__global__ void k_umma_images(const float* __restrict__ wpack, PinnNet net, int ldw, float* __restrict__ img) {
  const int l = blockIdx.y + 1;
  float* base = img + (size_t)(l - 1) * 4 * UIMG;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < UIMG; idx += gridDim.x * blockDim.x) {
    const int in = idx / UW, out = idx % UW;
    float hi, lo;
    split_rn(wpack[net.off_w[l] + in * ldw + out], hi, lo);
    const int f = (in >> 2) * (2 * UW * 4) + out * 4 + (in & 3);
    const int g = (out >> 2) * (2 * UW * 4) + in * 4 + (out & 3);
    base[f] = hi;
    base[f + UW * 4] = lo;
    base[2 * UIMG + g] = hi;
    base[2 * UIMG + g + UW * 4] = lo;
  }
}

// activation jets of a chunk with the activation branch OUTSIDE the unrolled loops (straight-line code,
// the 16 units of a chunk overlap their MUFU / FMA latencies)
__device__ __forceinline__ void act_fwd16(int act, const float (&a)[UCH], float (&y)[UCH], float (&d1)[UCH], float (&d2)[UCH], float (&s0)[UCH]) {
  if (act == PINN_TANH) {
#pragma unroll
    for (int i = 0; i < UCH; ++i) {
      const float t = tanh_bf(a[i]);
      y[i] = t; s0[i] = t;
      d1[i] = fmaf(-t, t, 1.0f);
      d2[i] = -2.0f * t * d1[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < UCH; ++i) {
      float sn, cs;
      sincos_ni(a[i], &sn, &cs);
      y[i] = sn; d1[i] = cs; d2[i] = -sn; s0[i] = a[i];
    }
  }
}
__device__ __forceinline__ void act_bwd16(int act, const float (&s0)[UCH], float (&y)[UCH], float (&d1)[UCH], float (&d2)[UCH], float (&d3)[UCH]) {
  if (act == PINN_TANH) {
#pragma unroll
    for (int i = 0; i < UCH; ++i) {
      const float t = s0[i];
      y[i] = t;
      d1[i] = fmaf(-t, t, 1.0f);
      d2[i] = -2.0f * t * d1[i];
      d3[i] = d1[i] * fmaf(6.0f * t, t, -2.0f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < UCH; ++i) {
      float sn, cs;
      sincos_ni(s0[i], &sn, &cs);
      y[i] = sn; d1[i] = cs; d2[i] = -sn; d3[i] = -cs;
    }
  }
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = v[i]; b[i] = v[8 + i]; }
  umma::tmem_st8(taddr, a);
  umma::tmem_st8(taddr + 8, b);
}

// D[128 x 128] = A_lo * [W_hi | W_lo] + A_hi * [W_hi | W_lo]: all four split products in 16 MMAs (N = 128 costs
// the same as N = 64), the small lo terms first so that the truncating accumulator adds them exactly;
// A (hi at TC_AHI, lo at TC_ALO) in TMEM, the N-concatenated weight image in smem; ONE thread
__device__ __forceinline__ void issue_layer_gemm(uint32_t tb, const float* img128) {
  const uint64_t dB = umma::smem_desc(umma::smem_addr(img128), 2 * UW * 16, 128);
  constexpr uint64_t STEP = (2 * 2 * UW * 16) >> 4;  // k-step advance (two 16-byte k groups of 128 rows) in 16-byte units
  const uint32_t i128 = umma::idesc_tf32(128, 2 * UW, 0, 0);
#pragma unroll
  for (int j = 0; j < UW / 8; ++j) umma::mma_tf32_ts(tb + TC_D, tb + TC_ALO + 8 * j, dB + j * STEP, i128, j > 0 ? 1u : 0u);
#pragma unroll
  for (int j = 0; j < UW / 8; ++j) umma::mma_tf32_ts(tb + TC_D, tb + TC_AHI + 8 * j, dB + j * STEP, i128, 1u);
}

// a[i] = D[u0 + i] + D[64 + u0 + i]
__device__ __forceinline__ void load_d16(uint32_t tl, int u0, float (&a)[UCH]) {
  float b[UCH];
  umma::tmem_ld16(tl + TC_D + u0, a);
  umma::tmem_ld16(tl + TC_D + UW + u0, b);
  umma::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < UCH; ++i) a[i] += b[i];
}

__global__ void __launch_bounds__(128, 2) jet_umma_eval_kernel(PinnLaunch L, const float* __restrict__ img, long long* __restrict__ clk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const PinnNet& net = L.net;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Lh = net.n_hidden, NL = Lh - 1;
  float* bimg = reinterpret_cast<float*>(smem_raw);  // [NL][2][UIMG]
  float* x1 = bimg + (size_t)NL * 2 * UIMG;          // [UCH][32] sigma'
  float* x2 = x1 + UCH * 32;                         // [UCH][32] sigma''
  float* xq = x2 + UCH * 32;                         // [2][UCH][32] beta_i * A_i^2
  float* us = xq + 2 * UCH * 32;                     // [4][32] network output channels
  int* s_ops = reinterpret_cast<int*>(us + 4 * 32);
  float* s_consts = reinterpret_cast<float*>(s_ops + PINN_MAX_OPS);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;

  for (int i = tid; i < NL * 2 * UIMG / 4; i += 128) {
    const int l = i / (2 * UIMG / 4), r = i % (2 * UIMG / 4);
    reinterpret_cast<float4*>(bimg)[i] = __ldg(reinterpret_cast<const float4*>(img + (size_t)l * 4 * UIMG) + r);
  }
  for (int i = tid; i < L.prog.n_ops; i += 128) s_ops[i] = L.prog.ops[i];
  for (int i = tid; i < PINN_MAX_CONSTS; i += 128) s_consts[i] = L.prog.consts[i];
  umma::fence_async_smem();
  if (warp == 0) umma::tmem_alloc(&tbase, TC_COLS);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tbase;
  const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);  // this warp's lane quadrant
  uint32_t parity = 0;
  const uint32_t idesc = umma::idesc_tf32(128, UW, 0, 0);
  const long long n_end = L.seg_pt_end[0];
  const float* W0 = L.wpack + net.off_w0;
  const float* wl = L.wpack + net.off_wl;
  const bool prof = clk != nullptr && blockIdx.x == 0 && tid == 0;
  long long c_epi = 0, c_mma = 0, c_out = 0, t_mark = prof ? clock64() : 0;
  auto lap = [&](long long& acc) {
    if (prof) { const long long now = clock64(); acc += now - t_mark; t_mark = now; }
  };

#pragma unroll 1
  for (int tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x) {
    const long long p0 = L.seg_pt_begin[0] + (long long)tile * UTP;
    const bool valid = p0 + lane < n_end;
    const long long gp = valid ? p0 + lane : p0;
    float z[1][3];
    {
      const float* zp = L.coords + gp * net.d_in;
      z[0][0] = __ldg(zp);
      z[0][1] = (net.d_in > 1) ? __ldg(zp + 1) : 0.f;
      z[0][2] = (net.d_in > 2) ? __ldg(zp + 2) : 0.f;
    }
    float beta[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) beta[i] = (net.lap_aux[i] >= 0) ? __ldg(L.aux + gp * L.n_aux + net.lap_aux[i]) : net.lap_beta[i];
    float h0 = 0.f, h1 = 0.f, h2 = 0.f;  // feature jet of this warp's channel
    {
      float hj[UK][3];
      feature_jets<UCfg>(net, z[0], beta, hj);
#pragma unroll
      for (int c = 0; c < UK; ++c)
        if (c == warp) { h0 = hj[c][0]; h1 = hj[c][1]; h2 = hj[c][2]; }
    }
    const float bq = (warp == 1) ? beta[0] : beta[1];
    float uacc = 0.f;

#pragma unroll 1
    for (int l = 0; l < Lh; ++l) {
      const int act = (l == 0) ? net.act_first : net.act_hidden;
      const float* bias = L.wpack + net.off_b[l];
      const bool last = (l == Lh - 1);
#pragma unroll 1
      for (int ch = 0; ch < UW / UCH; ++ch) {
        const int u0 = ch * UCH;
        float a[UCH], y[UCH];
        if (l == 0) {
#pragma unroll
          for (int i = 0; i < UCH; ++i)
            a[i] = net.scl * fmaf(h0, __ldg(W0 + u0 + i), fmaf(h1, __ldg(W0 + UW + u0 + i), h2 * __ldg(W0 + 2 * UW + u0 + i)));
        } else {
          load_d16(tl, u0, a);
        }
        if (warp == 0) {
          float d1[UCH], d2[UCH], s0[UCH];
#pragma unroll
          for (int i = 0; i < UCH; ++i) a[i] += __ldg(bias + u0 + i);
          act_fwd16(act, a, y, d1, d2, s0);
#pragma unroll
          for (int i = 0; i < UCH; ++i) {
            x1[i * 32 + lane] = d1[i];
            x2[i * 32 + lane] = d2[i];
          }
        } else if (warp < 3) {
#pragma unroll
          for (int i = 0; i < UCH; ++i) xq[((warp - 1) * UCH + i) * 32 + lane] = bq * a[i] * a[i];
        }
        __syncthreads();
        if (warp == 1 || warp == 2) {
#pragma unroll
          for (int i = 0; i < UCH; ++i) y[i] = x1[i * 32 + lane] * a[i];
        } else if (warp == 3) {
#pragma unroll
          for (int i = 0; i < UCH; ++i)
            y[i] = fmaf(x2[i * 32 + lane], xq[i * 32 + lane] + xq[(UCH + i) * 32 + lane], x1[i * 32 + lane] * a[i]);
        }
        if (last) {
#pragma unroll
          for (int i = 0; i < UCH; ++i) uacc = fmaf(y[i], __ldg(wl + u0 + i), uacc);
        } else {
          float hi[UCH], lo[UCH];
#pragma unroll
          for (int i = 0; i < UCH; ++i) split_rn(y[i], hi[i], lo[i]);
          tmem_st16(tl + TC_AHI + u0, hi);
          tmem_st16(tl + TC_ALO + u0, lo);
        }
        __syncthreads();
      }
      lap(c_epi);
      if (!last) {
        umma::tmem_st_wait();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        if (tid == 0) {
          issue_layer_gemm(tb, bimg + (size_t)(l * 2) * UIMG);
          umma::commit(&bar);
        }
        umma::mbar_wait(&bar, parity);
        parity ^= 1;
        umma::fence_after_sync();
        lap(c_mma);
      }
    }

    // output layer (per-thread dot product accumulated above) + residual program
    float uc = net.epsil * (uacc + (warp == 0 ? __ldg(L.wpack + net.off_bl) : 0.f));
    if (L.base) uc += __ldg(L.base + gp * UK + warp);
    us[warp * 32 + lane] = uc;
    __syncthreads();
    if (warp == 0) {
      float u[UK][1], f[1], df[UK][1];
      const float* auxp[1] = {L.aux ? (L.aux + gp * L.n_aux) : nullptr};
#pragma unroll
      for (int c = 0; c < UK; ++c) u[c][0] = us[c * 32 + lane];
      vm_run<UK, 1>(s_ops, L.prog.n_ops, s_consts, z, auxp, u, f, df);
      if (valid) {
        if (L.out_u) L.out_u[gp] = u[0][0];
        if (L.out_f) L.out_f[gp] = f[0];
        if (L.out_jets) {
#pragma unroll
          for (int c = 0; c < UK; ++c) L.out_jets[gp * UK + c] = u[c][0];
        }
      }
    }
    __syncthreads();
    lap(c_out);
  }
  if (prof) { clk[0] = c_epi; clk[1] = c_mma; clk[2] = c_out; }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tb, TC_COLS);
}

// ================================================================ training kernel (loss + gradient)
// shared memory map (bytes)
constexpr int SM_WIMG = 0;                    // current weight images (hi, lo): 2 x 16 KB
constexpr int SM_STH = 32768;                 // wgrad A operand: layer-input jets H (hi plane, lo plane), 2 x 32 KB
constexpr int SM_STA = SM_STH + 65536;        // wgrad B operand: pre-activation adjoints (hi plane | lo plane), 2 x 32 KB
constexpr int SM_EXCH = SM_STA + 65536;       // exchange arrays [unit group][NEX][UCH][32]
constexpr int NQ = 2;                         // warps per jet channel (each handles 64 / NQ units)
constexpr int CPW = (UW / UCH) / NQ;          // 16-unit chunks per warp and pass
constexpr int NEX = 7;                        // exchange arrays: sigma', sigma'', q_x, q_y, p_x, p_y, ybar_L
constexpr int UNT = 128 * NQ;                 // threads per CTA
constexpr int SM_MISC = SM_EXCH + NQ * NEX * UCH * 32 * 4;
constexpr int SM_TRAIN_BYTES = SM_MISC + ((NQ + 1) * 4 * 32 + PINN_MAX_OPS + PINN_MAX_CONSTS) * 4 + 64;
constexpr uint32_t TC_DW = 256, TC_TRAIN_COLS = 512;
constexpr int PLANE = 32768;  // bytes of one staging plane: 2 groups of 32 units x 128 rows x 128 B

// byte offset of (row, unit) inside a staging plane: MN-major operand, 128B swizzle with 32B atomicity
// (layout type 1): 32 units per 128 B row, 4 rows per 512 B atom, 16-byte chunk index ^ 2*(row%4)
__device__ __forceinline__ uint32_t st_off(int row, int unit) {
  return (uint32_t)(((unit >> 5) << 14) + ((row >> 2) << 9) + ((row & 3) << 7) + (((((unit & 31) >> 2) ^ ((row & 3) << 1))) << 4) +
                    ((unit & 3) << 2));
}

__device__ __forceinline__ void stage16(uint8_t* plane_hi, uint8_t* plane_lo, int row, int u0, const float (&v)[UCH]) {
#pragma unroll
  for (int j = 0; j < UCH / 4; ++j) {
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_rn(v[4 * j + i], h[i], l[i]);
    const uint32_t o = st_off(row, u0 + 4 * j);
    *reinterpret_cast<float4*>(plane_hi + o) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(plane_lo + o) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// DW[128 x 128] = [H_hi | H_lo]^T * [A_hi | A_lo] over the K = 128 rows: ONE chain of 16 MMAs gives all split
// products (rows 0..63 = H_hi units, 64..127 = H_lo units; columns 0..63 = A_hi, 64..127 = A_lo);
// W-bar[i][j] = DW[i][j] + DW[i][64+j] + DW[64+i][j] (the lo*lo block is not used)
__device__ __forceinline__ void issue_wgrad(uint32_t tb, const uint8_t* smem) {
  const uint64_t dH = umma::smem_desc(umma::smem_addr(smem + SM_STH), 16384, 512, 1);
  const uint64_t dA = umma::smem_desc(umma::smem_addr(smem + SM_STA), 16384, 512, 1);
  const uint32_t idw = umma::idesc_tf32(128, 128, 1, 1);
  constexpr uint64_t STEP = 1024 >> 4;  // 8 rows = two 512 B atoms
#pragma unroll
  for (int j = 0; j < 16; ++j) umma::mma_tf32_ss(tb + TC_DW, dH + j * STEP, dA + j * STEP, idw, j > 0 ? 1u : 0u);
}

__device__ __forceinline__ void bulk_load_w(uint8_t* smem, const float* src, uint64_t* bar) {
  mbar_expect_tx(bar, 2 * UIMG * 4);
  bulk_g2s(smem + SM_WIMG, src, UIMG * 4, bar);
  bulk_g2s(smem + SM_WIMG + UIMG * 4, src + UIMG, UIMG * 4, bar);
}

__global__ void __launch_bounds__(UNT, 1) jet_umma_train_kernel(PinnLaunch L, const float* __restrict__ img, int ldw,
                                                                long long* __restrict__ clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const PinnNet& net = L.net;
  // 4 * NQ warps: warp = jet channel (= TMEM lane quadrant, wid % 4), hsel = which group of 64 / NQ units the warp handles
  const int tid = threadIdx.x, wid = tid >> 5, warp = wid & 3, hsel = wid >> 2, lane = tid & 31;
  const int row = 32 * warp + lane;  // MMA row = 32 * channel + point
  const int Lh = net.n_hidden;
  float* x1 = reinterpret_cast<float*>(smem + SM_EXCH) + hsel * NEX * UCH * 32;
  float* x2 = x1 + UCH * 32;
  float* qx = x2 + UCH * 32;
  float* qy = qx + UCH * 32;
  float* px = qy + UCH * 32;
  float* py = px + UCH * 32;
  float* yl = py + UCH * 32;
  float* us = reinterpret_cast<float*>(smem + SM_MISC);  // [NQ][4][32] partial dot products of the unit groups
  float* ub = us + NQ * 4 * 32;                          // [4][32]
  int* s_ops = reinterpret_cast<int*>(ub + 4 * 32);
  float* s_consts = reinterpret_cast<float*>(s_ops + PINN_MAX_OPS);
  uint8_t* const Hh = smem + SM_STH;
  uint8_t* const Hl = Hh + PLANE;
  uint8_t* const Ah = smem + SM_STA;
  uint8_t* const Al = Ah + PLANE;
  __shared__ uint64_t barD, barW, barL;
  __shared__ uint32_t tbase;

  for (int i = tid; i < L.prog.n_ops; i += UNT) s_ops[i] = L.prog.ops[i];
  for (int i = tid; i < PINN_MAX_CONSTS; i += UNT) s_consts[i] = L.prog.consts[i];
  if (wid == 0) umma::tmem_alloc(&tbase, TC_TRAIN_COLS);
  if (tid == 0) {
    umma::mbar_init(&barD, 1);
    umma::mbar_init(&barW, 1);
    umma::mbar_init(&barL, 1);
    umma::fence_mbar_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tbase;
  const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);
  uint32_t parD = 0, parW = 0, parL = 0;
  // the cross-channel exchange only couples the four warps of a unit half: a named barrier per half lets
  // the two halves drift apart and overlap each other's latencies
  auto half_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + hsel) : "memory"); };
  const uint32_t idesc = umma::idesc_tf32(128, UW, 0, 0);
  const float* W0 = L.wpack + net.off_w0;
  const float* wl = L.wpack + net.off_wl;
  float* gacc = L.gacc + (size_t)blockIdx.x * net.pg;
  float* stash = L.stash + (size_t)blockIdx.x * Lh * (UK * UW * UTP);
  const int slot = L.seg_slot[0];
  const long long n_end = L.seg_pt_end[0];
  double lsum = 0.0;
  float blsum = 0.f;
  const bool prof = clk != nullptr && blockIdx.x == 0 && tid == 0;
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_mark = prof ? clock64() : 0;
  auto lap = [&](int k) {
    if (prof) { const long long now = clock64(); pc[k] += now - t_mark; t_mark = now; }
  };
  // flush the weight-gradient tile of GEMM layer g from TMEM into the CTA's accumulators (RN adds):
  // W-bar[i][j] = DW[i][j] + DW[i][64+j] + DW[64+i][j].  Row r of DW sits in TMEM lane r: quadrants 0/1
  // hold the H_hi rows i, quadrants 2/3 the H_lo rows 64+i, which travel through shared memory
  // (xs[j][i], the exchange buffer is idle here); the two warps of a quadrant split the columns.
  float* const xs = reinterpret_cast<float*>(smem + SM_EXCH);  // [64 columns][64 rows]
  auto flush_dw = [&](int g) {
    constexpr int NC = 64 / NQ;      // columns per warp
    const int r = 32 * warp + lane;  // DW row of this lane
    const int c0 = NC * hsel;        // first column handled by this warp
    const bool rmw = warp < 2 && g != Lh && (g != 0 || r < 3);
    float* const d = ((g == 0) ? gacc + net.off_w0 + (r < 3 ? r : 0) * UW : gacc + net.off_w[g == Lh ? 1 : g] + (r & 63) * ldw) + c0;
    float4 acc4[NC / 4];
    if (rmw) {  // issued first: the L2 round trip overlaps the TMEM reads and the exchange
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) acc4[j] = *reinterpret_cast<const float4*>(d + 4 * j);
    }
    if (warp >= 2) {
      float a[NC];
#pragma unroll
      for (int q = 0; q < NC / 16; ++q) umma::tmem_ld16(tl + TC_DW + c0 + 16 * q, reinterpret_cast<float(&)[16]>(a[16 * q]));
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < NC; ++j) xs[(c0 + j) * 64 + (r - 64)] = a[j];
    }
    __syncthreads();
    if (warp < 2) {
      float a[NC], b[NC];
#pragma unroll
      for (int q = 0; q < NC / 16; ++q) {
        umma::tmem_ld16(tl + TC_DW + c0 + 16 * q, reinterpret_cast<float(&)[16]>(a[16 * q]));
        umma::tmem_ld16(tl + TC_DW + 64 + c0 + 16 * q, reinterpret_cast<float(&)[16]>(b[16 * q]));
      }
      umma::tmem_ld_wait();
      if (g == Lh) {
        if (hsel == 0) gacc[net.off_wl + r] += a[0] + b[0] + xs[r];
      } else if (rmw) {
#pragma unroll
        for (int j = 0; j < NC / 4; ++j) {
          acc4[j].x += a[4 * j] + b[4 * j] + xs[(c0 + 4 * j) * 64 + r];
          acc4[j].y += a[4 * j + 1] + b[4 * j + 1] + xs[(c0 + 4 * j + 1) * 64 + r];
          acc4[j].z += a[4 * j + 2] + b[4 * j + 2] + xs[(c0 + 4 * j + 2) * 64 + r];
          acc4[j].w += a[4 * j + 3] + b[4 * j + 3] + xs[(c0 + 4 * j + 3) * 64 + r];
          *reinterpret_cast<float4*>(d + 4 * j) = acc4[j];
        }
      }
    }
    __syncthreads();  // xs (= the exchange buffer) is reused by the next pass
  };

#pragma unroll 1
  for (int tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x) {
    const long long p0 = L.seg_pt_begin[0] + (long long)tile * UTP;
    const bool valid = p0 + lane < n_end;
    const long long gp = valid ? p0 + lane : p0;
    float z[1][3];
    {
      const float* zp = L.coords + gp * net.d_in;
      z[0][0] = __ldg(zp);
      z[0][1] = (net.d_in > 1) ? __ldg(zp + 1) : 0.f;
      z[0][2] = (net.d_in > 2) ? __ldg(zp + 2) : 0.f;
    }
    float beta[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) beta[i] = (net.lap_aux[i] >= 0) ? __ldg(L.aux + gp * L.n_aux + net.lap_aux[i]) : net.lap_beta[i];
    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
    {
      float hj[UK][3];
      feature_jets<UCfg>(net, z[0], beta, hj);
#pragma unroll
      for (int c = 0; c < UK; ++c)
        if (c == warp) { h0 = hj[c][0]; h1 = hj[c][1]; h2 = hj[c][2]; }
    }
    const float bq = (warp == 1) ? beta[0] : beta[1];
    float uacc = 0.f;
    if (tid == 0 && Lh > 1) bulk_load_w(smem, img, &barL);  // forward image of GEMM layer 1

    // ------------------------------------------------------------ forward
#pragma unroll 1
    for (int l = 0; l < Lh; ++l) {
      const int act = (l == 0) ? net.act_first : net.act_hidden;
      const float* bias = L.wpack + net.off_b[l];
      const bool last = (l == Lh - 1);
      float* st_l = stash + ((size_t)(l * UK + warp) * UW) * UTP + lane;
#pragma unroll 1
      for (int ch = CPW * hsel; ch < CPW * (hsel + 1); ++ch) {
        const int u0 = ch * UCH;
        float a[UCH], y[UCH];
        if (l == 0) {
#pragma unroll
          for (int i = 0; i < UCH; ++i)
            a[i] = net.scl * fmaf(h0, __ldg(W0 + u0 + i), fmaf(h1, __ldg(W0 + UW + u0 + i), h2 * __ldg(W0 + 2 * UW + u0 + i)));
        } else {
          load_d16(tl, u0, a);
        }
        if (warp == 0) {
          float d1[UCH], d2[UCH], s0[UCH];
#pragma unroll
          for (int i = 0; i < UCH; ++i) a[i] += __ldg(bias + u0 + i);
          act_fwd16(act, a, y, d1, d2, s0);
#pragma unroll
          for (int i = 0; i < UCH; ++i) {
            x1[i * 32 + lane] = d1[i];
            x2[i * 32 + lane] = d2[i];
            st_l[(u0 + i) * UTP] = s0[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < UCH; ++i) st_l[(u0 + i) * UTP] = a[i];
          if (warp < 3) {
            float* q = (warp == 1) ? qx : qy;
#pragma unroll
            for (int i = 0; i < UCH; ++i) q[i * 32 + lane] = bq * a[i] * a[i];
          }
        }
        half_sync();
        if (warp == 1 || warp == 2) {
#pragma unroll
          for (int i = 0; i < UCH; ++i) y[i] = x1[i * 32 + lane] * a[i];
        } else if (warp == 3) {
#pragma unroll
          for (int i = 0; i < UCH; ++i) y[i] = fmaf(x2[i * 32 + lane], qx[i * 32 + lane] + qy[i * 32 + lane], x1[i * 32 + lane] * a[i]);
        }
        if (last) {
#pragma unroll
          for (int i = 0; i < UCH; ++i) uacc = fmaf(y[i], __ldg(wl + u0 + i), uacc);
        } else {
          float hi[UCH], lo[UCH];
#pragma unroll
          for (int i = 0; i < UCH; ++i) split_rn(y[i], hi[i], lo[i]);
          tmem_st16(tl + TC_AHI + u0, hi);
          tmem_st16(tl + TC_ALO + u0, lo);
        }
        half_sync();
      }
      lap(0);
      if (!last) {
        umma::tmem_st_wait();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        if (tid == 0) {
          umma::mbar_wait(&barL, parL);
          issue_layer_gemm(tb, reinterpret_cast<const float*>(smem + SM_WIMG));
          umma::commit(&barD);
        }
        parL ^= 1;
        umma::mbar_wait(&barD, parD);
        parD ^= 1;
        umma::fence_after_sync();
        if (tid == 0) {
          // the image buffer is free again: next forward image, or the data-gradient image of the top hidden layer
          const float* nxt = (l + 2 < Lh) ? img + (size_t)(l + 1) * 4 * UIMG : img + (size_t)(Lh - 2) * 4 * UIMG + 2 * UIMG;
          bulk_load_w(smem, nxt, &barL);
        }
        lap(1);
      }
    }

    // ------------------------------------------------------------ output layer, residual, seeds
    {
      float uc = net.epsil * (uacc + ((warp == 0 && hsel == 0) ? __ldg(L.wpack + net.off_bl) : 0.f));
      if (L.base && hsel == 0) uc += __ldg(L.base + gp * UK + warp);
      us[(hsel * 4 + warp) * 32 + lane] = uc;
    }
    __syncthreads();
    if (wid == 0) {
      float u[UK][1], f[1], df[UK][1];
      const float* auxp[1] = {L.aux ? (L.aux + gp * L.n_aux) : nullptr};
#pragma unroll
      for (int c = 0; c < UK; ++c) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) t += us[(q * 4 + c) * 32 + lane];
        u[c][0] = t;
      }
      vm_run<UK, 1>(s_ops, L.prog.n_ops, s_consts, z, auxp, u, f, df);
      const float sc = valid ? __ldg(L.seg_scale + slot) : 0.f;
#pragma unroll
      for (int c = 0; c < UK; ++c) ub[c * 32 + lane] = sc * f[0] * df[c][0];
      if (valid) lsum += (double)f[0] * (double)f[0];
      blsum += net.epsil * sc * f[0] * df[0][0];
    }
    __syncthreads();
    const float e = net.epsil * ub[warp * 32 + lane];  // adjoint of this row's pre-epsilon network output
    lap(2);

    // ------------------------------------------------------------ backward
    // adjoint operand of the output layer: one output unit (unit 0 = e), the other units zero
    {
      float v[UCH];
#pragma unroll
      for (int i = 0; i < UCH; ++i) v[i] = 0.f;
#pragma unroll 1
      for (int ch = CPW * hsel; ch < CPW * (hsel + 1); ++ch) {
        v[0] = (ch == 0) ? e : 0.f;
        stage16(Ah, Al, row, ch * UCH, v);
      }
    }
#pragma unroll 1
    for (int l = Lh - 1; l >= -1; --l) {
      if (l >= 0) {
        // ---- FB(l): ONE pass over the stash of hidden layer l gives (a) its output jets again -> wgrad A
        // operand of layer l+1 and (b) the adjoint of its pre-activations -> TMEM (dgrad operand of layer l)
        const int act = (l == 0) ? net.act_first : net.act_hidden;
        const float* st_l = stash + ((size_t)(l * UK + warp) * UW) * UTP + lane;
        float s2[CPW][UCH];  // this warp's stash of its chunks: the L2 round trip overlaps the data-gradient wait
#pragma unroll
        for (int c2 = 0; c2 < CPW; ++c2)
#pragma unroll
          for (int i = 0; i < UCH; ++i) s2[c2][i] = st_l[((CPW * hsel + c2) * UCH + i) * UTP];
        if (l < Lh - 1) {
          umma::mbar_wait(&barD, parD);  // data gradient of layer l+1 (issued in the previous iteration)
          parD ^= 1;
          umma::fence_after_sync();
          if (tid == 0 && l >= 1) bulk_load_w(smem, img + (size_t)(l - 1) * 4 * UIMG + 2 * UIMG, &barL);  // dgrad image of layer l
        }
#pragma unroll 1
        for (int ch = CPW * hsel; ch < CPW * (hsel + 1); ++ch) {
          const int u0 = ch * UCH;
          float s[UCH], yb[UCH], y[UCH], ab[UCH], d3v[UCH];
#pragma unroll
          for (int i = 0; i < UCH; ++i) s[i] = (CPW == 1 || ch == CPW * hsel) ? s2[0][i] : s2[CPW - 1][i];
          if (l == Lh - 1) {
#pragma unroll
            for (int i = 0; i < UCH; ++i) yb[i] = e * __ldg(wl + u0 + i);
          } else {
            load_d16(tl, u0, yb);
          }
          float aL[UCH];  // value channel only: the second-order channel's pre-activations (a_L * ybar_L term)
          if (warp == 0) {
            float d1[UCH], d2[UCH];
            const float* st_3 = stash + ((size_t)(l * UK + 3) * UW) * UTP + lane;
#pragma unroll
            for (int i = 0; i < UCH; ++i) aL[i] = st_3[(u0 + i) * UTP];
            act_bwd16(act, s, y, d1, d2, d3v);
#pragma unroll
            for (int i = 0; i < UCH; ++i) {
              x1[i * 32 + lane] = d1[i];
              x2[i * 32 + lane] = d2[i];
            }
          } else if (warp < 3) {
            float* q = (warp == 1) ? qx : qy;
            float* p = (warp == 1) ? px : py;
#pragma unroll
            for (int i = 0; i < UCH; ++i) {
              q[i * 32 + lane] = bq * s[i] * s[i];
              p[i * 32 + lane] = s[i] * yb[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < UCH; ++i) yl[i * 32 + lane] = yb[i];
          }
          half_sync();
          if (warp == 0) {
#pragma unroll
            for (int i = 0; i < UCH; ++i) {
              const int o = i * 32 + lane;
              const float yL = yl[o];
              const float P = fmaf(aL[i], yL, px[o] + py[o]);
              const float R = (qx[o] + qy[o]) * yL;
              ab[i] = fmaf(d3v[i], R, fmaf(x2[o], P, x1[o] * yb[i]));
            }
          } else if (warp < 3) {
#pragma unroll
            for (int i = 0; i < UCH; ++i) {
              const int o = i * 32 + lane;
              y[i] = x1[o] * s[i];
              ab[i] = fmaf(2.0f * x2[o] * bq * s[i], yl[o], x1[o] * yb[i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < UCH; ++i) {
              const int o = i * 32 + lane;
              y[i] = fmaf(x2[o], qx[o] + qy[o], x1[o] * s[i]);
              ab[i] = x1[o] * yb[i];
            }
          }
          stage16(Hh, Hl, row, u0, y);
          {
            float hi[UCH], lo[UCH];
#pragma unroll
            for (int i = 0; i < UCH; ++i) split_rn(ab[i], hi[i], lo[i]);
            tmem_st16(tl + TC_AHI + u0, hi);
            tmem_st16(tl + TC_ALO + u0, lo);
          }
          half_sync();
        }
        umma::tmem_st_wait();
      } else {
        // l = -1: the input of GEMM layer 0 are the feature jets (units 0..2)
        float v[UCH];
#pragma unroll
        for (int i = 0; i < UCH; ++i) v[i] = 0.f;
#pragma unroll 1
        for (int ch = CPW * hsel; ch < CPW * (hsel + 1); ++ch) {
          if (ch == 0) { v[0] = net.scl * h0; v[1] = net.scl * h1; v[2] = net.scl * h2; }
          else { v[0] = v[1] = v[2] = 0.f; }
          stage16(Hh, Hl, row, ch * UCH, v);
        }
      }
      lap(3);
      // ---- weight gradient of GEMM layer l+1 (staged H x staged adjoints) and data gradient of layer l (TMEM adjoints)
      umma::fence_async_smem();
      umma::fence_before_sync();
      __syncthreads();
      umma::fence_after_sync();
      if (tid == 0) {
        issue_wgrad(tb, smem);
        umma::commit(&barW);
        if (l >= 1) {
          umma::mbar_wait(&barL, parL);
          issue_layer_gemm(tb, reinterpret_cast<const float*>(smem + SM_WIMG));
          umma::commit(&barD);
        }
      }
      if (l >= 1) parL ^= 1;
      // bias gradient of layer l+1 (value-channel rows 0..31 of the staged adjoints), except for the output layer
      if (l + 1 < Lh && tid < UW) {
        float sb = 0.f;
#pragma unroll 8
        for (int r = 0; r < UTP; ++r) {
          const uint32_t o = st_off(r, tid);
          sb += *reinterpret_cast<const float*>(Ah + o) + *reinterpret_cast<const float*>(Al + o);
        }
        gacc[net.off_b[l + 1] + tid] += sb;
      }
      umma::mbar_wait(&barW, parW);
      parW ^= 1;
      umma::fence_after_sync();
      lap(4);
      flush_dw(l + 1);
      lap(5);
      if (l < 0) break;
      // ---- C(l): the adjoints of layer l (already split, in TMEM) become the staged operand of wgrad(l)
      __syncthreads();  // the bias-gradient readers of the old operand are done
#pragma unroll 1
      for (int ch = CPW * hsel; ch < CPW * (hsel + 1); ++ch) {
        const int u0 = ch * UCH;
        float hi[UCH], lo[UCH];
        umma::tmem_ld16(tl + TC_AHI + u0, hi);
        umma::tmem_ld16(tl + TC_ALO + u0, lo);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < UCH / 4; ++j) {
          const uint32_t o = st_off(row, u0 + 4 * j);
          *reinterpret_cast<float4*>(Ah + o) = make_float4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          *reinterpret_cast<float4*>(Al + o) = make_float4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
      }
      lap(6);
    }
    lap(7);
  }

  // ------------------------------------------------------------ per-CTA scalars
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
      blsum += __shfl_xor_sync(0xffffffffu, blsum, o);
    }
    if (lane == 0) {
      L.loss_part[(size_t)blockIdx.x * L.n_slots + slot] += lsum;
      gacc[net.off_bl] += blsum;
    }
  }
  if (prof) {
#pragma unroll
    for (int i = 0; i < 8; ++i) clk[i] = pc[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (wid == 0) umma::tmem_free(tb, TC_TRAIN_COLS);
}

size_t eval_smem_bytes(const PinnNet& net) {
  return sizeof(float) * ((size_t)(net.n_hidden - 1) * 2 * UIMG + 4 * UCH * 32 + 4 * 32 + PINN_MAX_OPS + PINN_MAX_CONSTS) + 16;
}

}  // namespace

bool jet_umma_supported(const PinnNet& net, int k, int n1, int n2, int mix) {
  return net.wp == UW && k == UK && n1 == 2 && n2 == 0 && mix == 2 && net.n_hidden >= 2 && net.n_hidden - 1 <= UMAXL;
}

size_t jet_umma_image_floats(const PinnNet& net) { return (size_t)(net.n_hidden - 1) * 4 * UIMG; }

cudaError_t jet_umma_build_images(const float* wpack, const PinnNet& net, int ldw, float* images, cudaStream_t st) {
  k_umma_images<<<dim3(8, net.n_hidden - 1), 256, 0, st>>>(wpack, net, ldw, images);
  return cudaGetLastError();
}

cudaError_t jet_umma_train_launch(const PinnLaunch& L, const float* images, int ldw, int grid, cudaStream_t st, long long* clk) {
  cudaError_t e = cudaFuncSetAttribute(jet_umma_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TRAIN_BYTES);
  if (e != cudaSuccess) return e;
  jet_umma_train_kernel<<<grid, UNT, SM_TRAIN_BYTES, st>>>(L, images, ldw, clk);
  return cudaGetLastError();
}

cudaError_t jet_umma_eval_launch(const PinnLaunch& L, const float* images, int grid_max, cudaStream_t st, long long* clk) {
  const size_t smem = eval_smem_bytes(L.net);
  cudaError_t e = cudaFuncSetAttribute(jet_umma_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = L.n_tiles < grid_max ? L.n_tiles : grid_max;
  jet_umma_eval_kernel<<<grid, 128, smem, st>>>(L, images, clk);
  return cudaGetLastError();
}
