// Fused jet-MLP kernel (fp32 SIMT path): forward Taylor-mode propagation of
// (u, du, d2u) through the MLP, residual program, loss partial sums, and the
// hand-written backward to the parameter gradient -- one persistent CTA per SM,
// activations in shared memory, per-layer stash in a CTA-private (L2-resident)
// scratch, weights streamed by cp.async.bulk (TMA bulk copy) + mbarrier.
//
// Replaces the nested reverse-mode autograd of pinn_app/software.py:246-297
// (vgmat/vectgrad/gov_eqn) and grad(loss_fun) at software.py:318-379, 390.
// Math: SURVEY.md section 8(a) addendum; layout + roofline: DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pinn_common.h"

template <int WP_, int N1_, int N2_, int MIX_, int NT_ = PINN_NT>
struct JetCfg {
  static constexpr int WP = WP_, N1 = N1_, N2 = N2_, MIX = MIX_;
  static constexpr bool LAP = (MIX == 2);       // one combined second-order channel sum_i beta_i d_ii
  static constexpr int K = 1 + N1 + N2 + (MIX ? 1 : 0);   // jet channels
  static constexpr int NT = NT_, TU = PINN_TU;
  static constexpr int UT = WP / TU;            // threads across units
  static constexpr int ROWS = NT / UT;          // thread rows across points
  static constexpr int PT = (NT >= 256) ? ((K == 1) ? 4 : (K <= 2 ? 2 : 1))
                                        : ((K == 1) ? 8 : (K <= 4 ? 3 : 2));  // points per thread
  static constexpr int TP = ROWS * PT;          // points per tile
  static constexpr int SP = K * WP + 4;         // smem stride per point (== 4 mod 32)
  static constexpr int CHUNK_FLOATS = (K >= 6) ? 1024 : 2048;     // 4 KB / 8 KB weight chunks
  static constexpr int KC = (CHUNK_FLOATS / WP) < WP ? (CHUNK_FLOATS / WP) : WP;  // weight rows per chunk
  static constexpr int NCH = WP / KC;
  static constexpr int TWU = (NT >= 256) ? 4 : 8;  // wgrad thread tile: 8 (k) x TWU (u)
  static constexpr int WT8 = WP / 8;
  static constexpr int WTU = WP / TWU;
  static constexpr int TILES = WT8 * WTU;       // 8 x TWU wgrad tiles
  static constexpr int NG = TILES <= NT ? NT / TILES : 1;
  static constexpr int NPASS = TILES <= NT ? 1 : TILES / NT;
  static constexpr int BG = NT >= WP ? NT / WP : 1;   // bias-gradient point groups
  static constexpr int HS_FLOATS = TP * SP;
  static constexpr uint32_t CHUNK_BYTES = KC * WP * 4;
  static_assert(WP % 32 == 0, "padded width must be a multiple of 32");
  // wgrad cross-group scratch must fit in Hs, the final small-gradient scratch in Hs+Gs,
  // point groups must divide the tile
  static constexpr bool OK = (NG == 1 || NG * WP * WP <= HS_FLOATS) && (5 * ROWS * WP + ROWS <= 2 * HS_FLOATS) &&
                             (TP % NG == 0) && (TP % BG == 0) && (ROWS >= 1);
  static constexpr size_t smem_bytes(bool train) {
    return (size_t)(HS_FLOATS * (train ? 2 : 1) + 2 * KC * WP + (NT > WP ? NT : WP) + PINN_MAX_OPS + PINN_MAX_CONSTS) * 4 + 64;
  }
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(b))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}

// packed fp32x2 FMA (Blackwell FFMA2): d.xy = a.xy * b.xy + d.xy.  ptxas folds a {x,x}
// pack into the scalar-broadcast operand form, so the broadcast costs no instruction.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float x, float y) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ void ffma2(f32x2_t& d, f32x2_t a, f32x2_t b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// ---------------------------------------------------------------- activations
// Branch-free tanh, abs error ~1e-7: odd polynomial below 0.25, 1 - 2/(exp(2|x|)+1) above
// (ex2.approx + rcp.approx, the two MUFU ops), sign restored with copysign.
__device__ __forceinline__ float tanh_bf(float x) {
  const float ax = fabsf(x);
  const float x2 = ax * ax;
  float p = fmaf(x2, 2.1869488536155203e-2f, -5.3968253968253968e-2f);
  p = fmaf(x2, p, 1.3333333333333333e-1f);
  p = fmaf(x2, p, -3.3333333333333333e-1f);
  p = fmaf(x2 * ax, p, ax);
  const float e = exp2f(ax * 2.8853900817779268f);
  const float r = 1.0f - __fdividef(2.0f, e + 1.0f);
  return copysignf(ax < 0.25f ? p : r, x);
}

// accurate sincos kept out of line: its slow range-reduction path is large and would be
// replicated 16x per call site (instruction-cache footprint)
static __device__ __noinline__ void sincos_ni(float x, float* s, float* c) { sincosf(x, s, c); }

// forward: y = act(a0), d1 = act'(a0), d2 = act''(a0); s0 = value stashed for backward
__device__ __forceinline__ void act_fwd(int act, float a0, float& y, float& d1, float& d2, float& s0) {
  if (act == PINN_TANH) {
    y = tanh_bf(a0);
    d1 = fmaf(-y, y, 1.0f);
    d2 = -2.0f * y * d1;
    s0 = y;
  } else {
    float s, c;
    sincos_ni(a0, &s, &c);
    y = s; d1 = c; d2 = -s; s0 = a0;
  }
}
// from the stashed value: y, d1, d2, d3
__device__ __forceinline__ void act_bwd(int act, float s0, float& y, float& d1, float& d2, float& d3) {
  if (act == PINN_TANH) {
    y = s0;
    d1 = fmaf(-y, y, 1.0f);
    d2 = -2.0f * y * d1;
    d3 = d1 * fmaf(6.0f * y, y, -2.0f);
  } else {
    float s, c;
    sincos_ni(s0, &s, &c);
    y = s; d1 = c; d2 = -s; d3 = -c;
  }
}

// ---------------------------------------------------------------- residual VM
// Forward-mode dual numbers over the K network-output channels, evaluated for the PT
// points of a thread in ONE pass over the bytecode (ops/consts staged in shared memory).
template <int K, int PT>
__device__ __noinline__ void vm_run(const int* __restrict__ ops, int n_ops, const float* __restrict__ consts,
                                    const float (*z)[3], const float* const* aux, const float (*u)[PT],
                                    float* f, float (*df)[PT]) {
  float sv[PINN_VM_STACK][PT];
  float sd[PINN_VM_STACK][K][PT];
  int sp = 0;
  for (int i = 0; i < n_ops; ++i) {
    const int w = ops[i];
    const int op = w & 0xff, arg = w >> 8;
    if (op <= OP_AUX) {  // pushes
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        float v;
        if (op == OP_CONST) v = consts[arg];
        else if (op == OP_COORD) v = z[p][arg];
        else if (op == OP_JET) v = u[arg][p];
        else v = aux[p][arg];
        sv[sp][p] = v;
#pragma unroll
        for (int c = 0; c < K; ++c) sd[sp][c][p] = (op == OP_JET && c == arg) ? 1.f : 0.f;
      }
      ++sp;
    } else if (op <= OP_DIV) {  // binary
      --sp;
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        const float a = sv[sp - 1][p], b = sv[sp][p];
        if (op == OP_ADD) {
          sv[sp - 1][p] = a + b;
#pragma unroll
          for (int c = 0; c < K; ++c) sd[sp - 1][c][p] += sd[sp][c][p];
        } else if (op == OP_SUB) {
          sv[sp - 1][p] = a - b;
#pragma unroll
          for (int c = 0; c < K; ++c) sd[sp - 1][c][p] -= sd[sp][c][p];
        } else if (op == OP_MUL) {
          sv[sp - 1][p] = a * b;
#pragma unroll
          for (int c = 0; c < K; ++c) sd[sp - 1][c][p] = fmaf(sd[sp - 1][c][p], b, a * sd[sp][c][p]);
        } else {
          const float ib = 1.0f / b, q = a * ib;
          sv[sp - 1][p] = q;
#pragma unroll
          for (int c = 0; c < K; ++c) sd[sp - 1][c][p] = (sd[sp - 1][c][p] - q * sd[sp][c][p]) * ib;
        }
      }
    } else {  // unary: value v and derivative factor dv
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        const float a = sv[sp - 1][p];
        float v, dv;
        switch (op) {
          case OP_NEG: v = -a; dv = -1.f; break;
          case OP_POWI: { float pm1 = 1.f;
            for (int q = 1; q < arg; ++q) pm1 *= a;
            v = (arg == 0) ? 1.f : pm1 * a; dv = (arg == 0) ? 0.f : (float)arg * pm1; break; }
          case OP_POWF: { const float e = consts[arg]; v = powf(a, e); dv = e * powf(a, e - 1.0f); break; }
          case OP_SIN: { float s, co; sincosf(a, &s, &co); v = s; dv = co; break; }
          case OP_COS: { float s, co; sincosf(a, &s, &co); v = co; dv = -s; break; }
          case OP_EXP: v = expf(a); dv = v; break;
          case OP_LOG: v = logf(a); dv = 1.0f / a; break;
          case OP_TANH: v = tanhf(a); dv = 1.f - v * v; break;
          case OP_SQRT: v = sqrtf(a); dv = 0.5f / v; break;
          default: v = a; dv = 1.f; break;
        }
        sv[sp - 1][p] = v;
#pragma unroll
        for (int c = 0; c < K; ++c) sd[sp - 1][c][p] *= dv;
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PT; ++p) {
    f[p] = sv[0][p];
#pragma unroll
    for (int c = 0; c < K; ++c) df[c][p] = sd[0][c][p];
  }
}

// ---------------------------------------------------------------- building blocks
// acc[c][p][j] += sum_{k in chunk} S[pt(p)][c][kbase+k] * wc[k][unit(j)]
// S points at the thread-row's first point; wc at the chunk base (smem).
// acc[c][p][j] += sum_{k in chunk} S[pt(p)][c][kbase+k] * wc[k][unit(j)]  (packed FFMA2)
template <class C>
__device__ __forceinline__ void gemm_chunk(float (&acc)[C::K][C::PT][8], const float* __restrict__ S,
                                           const float* __restrict__ wc, int ua, int kbase) {
  f32x2_t acc2[C::K][C::PT][4];
#pragma unroll
  for (int c = 0; c < C::K; ++c)
#pragma unroll
    for (int p = 0; p < C::PT; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[c][p][j] = pack2(acc[c][p][2 * j], acc[c][p][2 * j + 1]);
#pragma unroll 1
  for (int kk = 0; kk < C::KC; kk += 4) {
    float4 a[C::K][C::PT];
#pragma unroll
    for (int c = 0; c < C::K; ++c)
#pragma unroll
      for (int p = 0; p < C::PT; ++p)
        a[c][p] = *reinterpret_cast<const float4*>(S + p * (C::ROWS * C::SP) + c * C::WP + kbase + kk);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b0 = *reinterpret_cast<const float4*>(wc + (kk + q) * C::WP + ua);
      const float4 b1 = *reinterpret_cast<const float4*>(wc + (kk + q) * C::WP + C::WP / 2 + ua);
      const f32x2_t bp0 = pack2(b0.x, b0.y), bp1 = pack2(b0.z, b0.w), bp2 = pack2(b1.x, b1.y), bp3 = pack2(b1.z, b1.w);
#pragma unroll
      for (int c = 0; c < C::K; ++c)
#pragma unroll
        for (int p = 0; p < C::PT; ++p) {
          const float av = (q == 0) ? a[c][p].x : (q == 1) ? a[c][p].y : (q == 2) ? a[c][p].z : a[c][p].w;
          const f32x2_t aa = pack2(av, av);
          ffma2(acc2[c][p][0], aa, bp0);
          ffma2(acc2[c][p][1], aa, bp1);
          ffma2(acc2[c][p][2], aa, bp2);
          ffma2(acc2[c][p][3], aa, bp3);
        }
    }
  }
#pragma unroll
  for (int c = 0; c < C::K; ++c)
#pragma unroll
    for (int p = 0; p < C::PT; ++p)
#pragma unroll
      for (int j = 0; j < 4; ++j) unpack2(acc2[c][p][j], acc[c][p][2 * j], acc[c][p][2 * j + 1]);
}

// store the thread's [K][PT][8] register tile to smem S (point-major, channel, unit)
template <class C>
__device__ __forceinline__ void store_tile(float* __restrict__ S, const float (&acc)[C::K][C::PT][8],
                                           int row, int ua) {
#pragma unroll
  for (int p = 0; p < C::PT; ++p)
#pragma unroll
    for (int c = 0; c < C::K; ++c) {
      float* d = S + (row + p * C::ROWS) * C::SP + c * C::WP;
      *reinterpret_cast<float4*>(d + ua) = make_float4(acc[c][p][0], acc[c][p][1], acc[c][p][2], acc[c][p][3]);
      *reinterpret_cast<float4*>(d + C::WP / 2 + ua) =
          make_float4(acc[c][p][4], acc[c][p][5], acc[c][p][6], acc[c][p][7]);
    }
}

// feature jets of the network input (software.py:172-175 for 'polar')
template <class C>
__device__ __forceinline__ void feature_jets(const PinnNet& net, const float (&z)[3], const float (&beta)[3],
                                             float (&hj)[C::K][3]) {
#pragma unroll
  for (int c = 0; c < C::K; ++c) hj[c][0] = hj[c][1] = hj[c][2] = 0.f;
  if (net.feat_mode == PINN_FEAT_POLAR) {
    float s, co;
    sincosf(z[1], &s, &co);
    hj[0][0] = fmaf(net.fa[0], z[0], net.fb[0]); hj[0][1] = co; hj[0][2] = s;
    if (C::N1 >= 1) hj[1][0] = net.fa[0];
    if (C::N1 >= 2) { hj[2][1] = -s; hj[2][2] = co; }
    if (C::N2 >= 2) { hj[1 + C::N1 + 1][1] = -co; hj[1 + C::N1 + 1][2] = -s; }
    if (C::LAP) { hj[C::K - 1][1] = -beta[1] * co; hj[C::K - 1][2] = -beta[1] * s; }
  } else {
#pragma unroll
    for (int f = 0; f < 3; ++f) hj[0][f] = (f < net.d_in) ? fmaf(net.fa[f], z[f], net.fb[f]) : 0.f;
#pragma unroll
    for (int i = 0; i < C::N1; ++i) hj[1 + i][i] = net.fa[i];
  }
}

// pre-activation jets A_c (bias NOT yet added) -> output jets Y_c, in place; optional stash
template <class C, bool TRAIN>
__device__ __forceinline__ void act_forward(float (&acc)[C::K][C::PT][8], const float* __restrict__ bias,
                                            int act, float* __restrict__ stash_l, int row, int ua,
                                            const float (&beta)[C::PT][3]) {
  const float4 ba = __ldg(reinterpret_cast<const float4*>(bias + ua));
  const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + C::WP / 2 + ua));
  const float b[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
  for (int p = 0; p < C::PT; ++p) {
    float y[8], d1[8], d2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s0;
      act_fwd(act, acc[0][p][j] + b[j], y[j], d1[j], d2[j], s0);
      acc[0][p][j] = s0;
    }
    if (TRAIN) {
#pragma unroll
      for (int c = 0; c < C::K; ++c) {
        float* d = stash_l + ((size_t)(row + p * C::ROWS) * C::K + c) * C::WP;
        *reinterpret_cast<float4*>(d + ua) = make_float4(acc[c][p][0], acc[c][p][1], acc[c][p][2], acc[c][p][3]);
        *reinterpret_cast<float4*>(d + C::WP / 2 + ua) =
            make_float4(acc[c][p][4], acc[c][p][5], acc[c][p][6], acc[c][p][7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int i = 0; i < C::N2; ++i) {
        const float Ai = acc[1 + i][p][j];
        acc[1 + C::N1 + i][p][j] = fmaf(d2[j] * Ai, Ai, d1[j] * acc[1 + C::N1 + i][p][j]);
      }
      if (C::MIX == 1)
        acc[C::K - 1][p][j] = fmaf(d2[j] * acc[1][p][j], acc[2][p][j], d1[j] * acc[C::K - 1][p][j]);
      if (C::LAP) {
        float S = 0.f;
#pragma unroll
        for (int i = 0; i < C::N1; ++i) S = fmaf(beta[p][i] * acc[1 + i][p][j], acc[1 + i][p][j], S);
        acc[C::K - 1][p][j] = fmaf(d2[j], S, d1[j] * acc[C::K - 1][p][j]);
      }
#pragma unroll
      for (int i = 0; i < C::N1; ++i) acc[1 + i][p][j] *= d1[j];
      acc[0][p][j] = y[j];
    }
  }
}

// adjoint of act_forward: acc holds Ybar_c on entry, Abar_c on exit (reads the stash)
template <class C>
__device__ __forceinline__ void act_backward(float (&acc)[C::K][C::PT][8], int act,
                                             const float* __restrict__ stash_l, int row, int ua,
                                             const float (&beta)[C::PT][3]) {
#pragma unroll
  for (int p = 0; p < C::PT; ++p) {
    float st[C::K][8];
#pragma unroll
    for (int c = 0; c < C::K; ++c) {
      const float* s = stash_l + ((size_t)(row + p * C::ROWS) * C::K + c) * C::WP;
      const float4 v0 = *reinterpret_cast<const float4*>(s + ua);
      const float4 v1 = *reinterpret_cast<const float4*>(s + C::WP / 2 + ua);
      st[c][0] = v0.x; st[c][1] = v0.y; st[c][2] = v0.z; st[c][3] = v0.w;
      st[c][4] = v1.x; st[c][5] = v1.y; st[c][6] = v1.z; st[c][7] = v1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y, d1, d2, d3;
      act_bwd(act, st[0][j], y, d1, d2, d3);
      float ab0 = d1 * acc[0][p][j];
      float ab1[C::N1 > 0 ? C::N1 : 1];
#pragma unroll
      for (int i = 0; i < C::N1; ++i) {
        const float yb = acc[1 + i][p][j];
        ab1[i] = d1 * yb;
        ab0 = fmaf(d2 * st[1 + i][j], yb, ab0);
      }
#pragma unroll
      for (int i = 0; i < C::N2; ++i) {
        const float yb = acc[1 + C::N1 + i][p][j];
        const float Ai = st[1 + i][j], Aii = st[1 + C::N1 + i][j];
        acc[1 + C::N1 + i][p][j] = d1 * yb;
        ab1[i] = fmaf(2.0f * d2 * Ai, yb, ab1[i]);
        ab0 = fmaf(fmaf(d3 * Ai, Ai, d2 * Aii), yb, ab0);
      }
      if (C::MIX == 1) {
        const float yb = acc[C::K - 1][p][j];
        const float A0 = st[1][j], A1 = st[2][j], A01 = st[C::K - 1][j];
        acc[C::K - 1][p][j] = d1 * yb;
        ab1[0] = fmaf(d2 * A1, yb, ab1[0]);
        ab1[C::N1 > 1 ? 1 : 0] = fmaf(d2 * A0, yb, ab1[C::N1 > 1 ? 1 : 0]);
        ab0 = fmaf(fmaf(d3 * A0, A1, d2 * A01), yb, ab0);
      }
      if (C::LAP) {
        const float yb = acc[C::K - 1][p][j];
        float S = 0.f;
#pragma unroll
        for (int i = 0; i < C::N1; ++i) {
          const float bA = beta[p][i] * st[1 + i][j];
          S = fmaf(bA, st[1 + i][j], S);
          ab1[i] = fmaf(2.0f * d2 * bA, yb, ab1[i]);
        }
        acc[C::K - 1][p][j] = d1 * yb;
        ab0 = fmaf(fmaf(d3, S, d2 * st[C::K - 1][j]), yb, ab0);
      }
#pragma unroll
      for (int i = 0; i < C::N1; ++i) acc[1 + i][p][j] = ab1[i];
      acc[0][p][j] = ab0;
    }
  }
}

// recompute a layer's OUTPUT jets Y_c from its stash and write them to smem S
template <class C>
__device__ __forceinline__ void recompute_outputs(float* __restrict__ S, int act,
                                                  const float* __restrict__ stash_l, int row, int ua,
                                                  const float (&beta)[C::PT][3]) {
#pragma unroll
  for (int p = 0; p < C::PT; ++p) {
    float st[C::K][8];
#pragma unroll
    for (int c = 0; c < C::K; ++c) {
      const float* s = stash_l + ((size_t)(row + p * C::ROWS) * C::K + c) * C::WP;
      const float4 v0 = *reinterpret_cast<const float4*>(s + ua);
      const float4 v1 = *reinterpret_cast<const float4*>(s + C::WP / 2 + ua);
      st[c][0] = v0.x; st[c][1] = v0.y; st[c][2] = v0.z; st[c][3] = v0.w;
      st[c][4] = v1.x; st[c][5] = v1.y; st[c][6] = v1.z; st[c][7] = v1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y, d1, d2, d3;
      act_bwd(act, st[0][j], y, d1, d2, d3);
#pragma unroll
      for (int i = 0; i < C::N2; ++i) {
        const float Ai = st[1 + i][j];
        st[1 + C::N1 + i][j] = fmaf(d2 * Ai, Ai, d1 * st[1 + C::N1 + i][j]);
      }
      if (C::MIX == 1) st[C::K - 1][j] = fmaf(d2 * st[1][j], st[2][j], d1 * st[C::K - 1][j]);
      if (C::LAP) {
        float Sq = 0.f;
#pragma unroll
        for (int i = 0; i < C::N1; ++i) Sq = fmaf(beta[p][i] * st[1 + i][j], st[1 + i][j], Sq);
        st[C::K - 1][j] = fmaf(d2, Sq, d1 * st[C::K - 1][j]);
      }
#pragma unroll
      for (int i = 0; i < C::N1; ++i) st[1 + i][j] *= d1;
      st[0][j] = y;
    }
#pragma unroll
    for (int c = 0; c < C::K; ++c) {
      float* d = S + (row + p * C::ROWS) * C::SP + c * C::WP;
      *reinterpret_cast<float4*>(d + ua) = make_float4(st[c][0], st[c][1], st[c][2], st[c][3]);
      *reinterpret_cast<float4*>(d + C::WP / 2 + ua) = make_float4(st[c][4], st[c][5], st[c][6], st[c][7]);
    }
  }
}

// unit/row index of wgrad-tile element e (0..7): split halves like the GEMM tile
template <class C>
__device__ __forceinline__ int tile_idx(int t, int e) {
  return (e < 4) ? (4 * t + e) : (C::WP / 2 + 4 * t + (e - 4));
}

// hidden-layer weight gradient for one layer: gW[k][u] += sum_{pt,c} H[pt][c][k] * G[pt][c][u]
// and gB[u] += sum_pt G[pt][0][u].  Hs doubles as cross-group scratch when NG > 1
// (callers guarantee Hs/Gs are complete on entry; bsc is a small smem scratch).
template <class C>
__device__ __forceinline__ void wgrad_layer(float* __restrict__ Hs, const float* __restrict__ Gs,
                                            float* __restrict__ bsc, float* __restrict__ gW,
                                            float* __restrict__ gB, int tid) {
  // bias gradient partials: BG point groups x WP units
  for (int idx = tid; idx < C::WP * C::BG; idx += C::NT) {
    const int u = idx % C::WP, g = idx / C::WP;
    constexpr int PB = C::TP / C::BG;
    const float* gp = Gs + (g * PB) * C::SP + u;
    float b0 = 0.f, b1 = 0.f;
#pragma unroll 4
    for (int pt = 0; pt + 1 < PB; pt += 2) { b0 += gp[pt * C::SP]; b1 += gp[(pt + 1) * C::SP]; }
    if (PB & 1) b0 += gp[(PB - 1) * C::SP];
    bsc[idx] = b0 + b1;
  }
#pragma unroll 1
  for (int pass = 0; pass < C::NPASS; ++pass) {
    const int tt = (C::NG > 1) ? (tid % C::TILES) : (tid + pass * C::NT);
    const int grp = (C::NG > 1) ? (tid / C::TILES) : 0;
    const int ki = tt / C::WTU, ui = tt % C::WTU;
    constexpr int PPG = C::TP / C::NG;  // points per group
    constexpr int TWU = C::TWU;
    float w[8][TWU];
    f32x2_t w2[8][TWU / 2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TWU / 2; ++j) w2[i][j] = pack2(0.f, 0.f);
    const float* hp = Hs + (grp * PPG) * C::SP + 4 * ki;
    const float* gp = Gs + (grp * PPG) * C::SP + 4 * ui;
#pragma unroll 1
    for (int pt = 0; pt < PPG; ++pt) {
#pragma unroll
      for (int c = 0; c < C::K; ++c) {
        const float4 h0 = *reinterpret_cast<const float4*>(hp + pt * C::SP + c * C::WP);
        const float4 h1 = *reinterpret_cast<const float4*>(hp + pt * C::SP + c * C::WP + C::WP / 2);
        const float4 g0 = *reinterpret_cast<const float4*>(gp + pt * C::SP + c * C::WP);
        float4 g1 = g0;
        if (TWU == 8) g1 = *reinterpret_cast<const float4*>(gp + pt * C::SP + c * C::WP + C::WP / 2);
        const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const f32x2_t gq[4] = {pack2(g0.x, g0.y), pack2(g0.z, g0.w), pack2(g1.x, g1.y), pack2(g1.z, g1.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const f32x2_t hh = pack2(h[i], h[i]);
#pragma unroll
          for (int j = 0; j < TWU / 2; ++j) ffma2(w2[i][j], hh, gq[j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TWU / 2; ++j) unpack2(w2[i][j], w[i][2 * j], w[i][2 * j + 1]);
    if (C::NG == 1) {
      // read-modify-write the CTA-private accumulator directly
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float* d = gW + tile_idx<C>(ki, i) * C::WP;
        float4* d0 = reinterpret_cast<float4*>(d + 4 * ui);
        float4 v0 = *d0;
        v0.x += w[i][0]; v0.y += w[i][1]; v0.z += w[i][2]; v0.w += w[i][3];
        *d0 = v0;
        if (TWU == 8) {
          float4* d1 = reinterpret_cast<float4*>(d + C::WP / 2 + 4 * ui);
          float4 v1 = *d1;
          v1.x += w[i][TWU - 4]; v1.y += w[i][TWU - 3]; v1.z += w[i][TWU - 2]; v1.w += w[i][TWU - 1];
          *d1 = v1;
        }
      }
    } else {
      __syncthreads();  // everyone finished reading Hs
      float* sc = Hs + grp * (C::WP * C::WP);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float* d = sc + tile_idx<C>(ki, i) * C::WP;
        *reinterpret_cast<float4*>(d + 4 * ui) = make_float4(w[i][0], w[i][1], w[i][2], w[i][3]);
        if (TWU == 8)
          *reinterpret_cast<float4*>(d + C::WP / 2 + 4 * ui) =
              make_float4(w[i][TWU - 4], w[i][TWU - 3], w[i][TWU - 2], w[i][TWU - 1]);
      }
      __syncthreads();
      constexpr int NV = C::WP * C::WP / 4;  // float4 outputs
      for (int v = tid; v < NV; v += C::NT) {
        float4 s = reinterpret_cast<const float4*>(Hs)[v];
#pragma unroll
        for (int g = 1; g < C::NG; ++g) {
          const float4 t = reinterpret_cast<const float4*>(Hs + g * (C::WP * C::WP))[v];
          s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        float4* d = reinterpret_cast<float4*>(gW) + v;
        float4 o = *d;
        o.x += s.x; o.y += s.y; o.z += s.z; o.w += s.w;
        *d = o;
      }
    }
  }
  if (C::NG == 1) __syncthreads();  // bsc complete (the NG > 1 path has synchronised already)
  for (int u = tid; u < C::WP; u += C::NT) {
    float b = bsc[u];
#pragma unroll
    for (int g = 1; g < C::BG; ++g) b += bsc[g * C::WP + u];
    gB[u] += b;
  }
}

// ---------------------------------------------------------------- the kernel
template <class C, bool TRAIN>
__global__ void __launch_bounds__(C::NT, 2) jet_mlp_kernel(const __grid_constant__ PinnLaunch L) {
  static_assert(C::OK, "invalid kernel configuration");
  constexpr int K = C::K, PT = C::PT, WP = C::WP, SP = C::SP, ROWS = C::ROWS, NCH = C::NCH, KC = C::KC;
  extern __shared__ __align__(128) float smem[];
  float* Hs = smem;
  float* Gs = Hs + C::HS_FLOATS;
  float* Wc = TRAIN ? (Gs + C::HS_FLOATS) : Gs;
  float* bsc = Wc + 2 * KC * WP;
  int* s_ops = reinterpret_cast<int*>(bsc + (C::NT > WP ? C::NT : WP));
  float* s_consts = reinterpret_cast<float*>(s_ops + PINN_MAX_OPS);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_consts + PINN_MAX_CONSTS);

  const PinnNet& net = L.net;
  const int tid = threadIdx.x;
  const int ut = tid % C::UT, row = tid / C::UT;
  const int ua = 4 * ut;
  const int Lh = net.n_hidden;
  const int nF = (Lh - 1) * NCH;
  const int S = TRAIN ? 2 * nF : nF;
  const int my_tiles = (L.n_tiles > (int)blockIdx.x) ? (L.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total = (long long)my_tiles * S;
  long long gpos = 0;

  auto issue = [&](long long g) {
    const int p = (int)(g % S);
    const float* src;
    if (p < nF) {
      const int l = 1 + p / NCH, ch = p % NCH;
      src = L.wpack + net.off_w[l] + ch * (KC * WP);
    } else {
      const int q = p - nF;
      const int l = (Lh - 1) - q / NCH, ch = q % NCH;
      src = L.wpack + net.off_wt[l] + ch * (KC * WP);
    }
    const int st = (int)(g & 1);
    mbar_expect_tx(&mbar[st], C::CHUNK_BYTES);
    bulk_g2s(Wc + st * (KC * WP), src, C::CHUNK_BYTES, &mbar[st]);
  };

  for (int i = tid; i < L.prog.n_ops; i += C::NT) s_ops[i] = L.prog.ops[i];
  for (int i = tid; i < PINN_MAX_CONSTS; i += C::NT) s_consts[i] = L.prog.consts[i];
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && total > 0) issue(0);

  float* stash = TRAIN ? (L.stash + (size_t)blockIdx.x * Lh * (C::TP * K * WP)) : nullptr;
  float* gacc = TRAIN ? (L.gacc + (size_t)blockIdx.x * net.pg) : nullptr;
  constexpr size_t STL = (size_t)C::TP * K * WP;  // stash floats per layer

  // persistent small-gradient accumulators (first layer, output layer)
  float w0acc[3][8], b0acc[8], wlacc[8], blacc = 0.f;
  double lcur = 0.0;   // loss partial of the current slot
  int cur_slot = -1;
  if (TRAIN) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { w0acc[0][j] = w0acc[1][j] = w0acc[2][j] = 0.f; b0acc[j] = 0.f; wlacc[j] = 0.f; }
  }
  // per-slot loss partials are folded through smem (fixed order => deterministic)
  auto flush_loss = [&](int slot) {
    double* dsc = reinterpret_cast<double*>(Gs);
    __syncthreads();
    if (ut == 0) dsc[row] = lcur;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int r = 0; r < ROWS; ++r) t += dsc[r];
      L.loss_part[(size_t)blockIdx.x * L.n_slots + slot] += t;
    }
    __syncthreads();
    lcur = 0.0;
  };

  // one consumed weight chunk: prefetch the next, wait for this one
  auto chunk_begin = [&]() -> const float* {
    if (tid == 0 && gpos + 1 < total) issue(gpos + 1);
    const int st = (int)(gpos & 1);
    mbar_wait(&mbar[st], (uint32_t)((gpos >> 1) & 1));
    return Wc + st * (KC * WP);
  };

#pragma unroll 1
  for (int it = 0; it < my_tiles; ++it) {
    const int tile = blockIdx.x + it * gridDim.x;
    int seg = 0;
    while (seg + 1 < L.n_seg && tile >= L.seg_tile_end[seg]) ++seg;
    const int tile0 = seg ? L.seg_tile_end[seg - 1] : 0;
    const long long pbegin = L.seg_pt_begin[seg] + (long long)(tile - tile0) * C::TP;
    const long long rem = L.seg_pt_end[seg] - pbegin;
    const int cnt = rem < C::TP ? (int)rem : C::TP;
    const int slot = L.seg_slot[seg];
    if (TRAIN && slot != cur_slot) {   // uniform across the CTA
      if (cur_slot >= 0) flush_loss(cur_slot);
      cur_slot = slot;
    }

    float z[PT][3];
    long long gp[PT];
    bool valid[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      const int lp = row + p * ROWS;
      valid[p] = lp < cnt;
      gp[p] = pbegin + (valid[p] ? lp : 0);
      const float* zp = L.coords + gp[p] * net.d_in;
      z[p][0] = __ldg(zp);
      z[p][1] = (net.d_in > 1) ? __ldg(zp + 1) : 0.f;
      z[p][2] = (net.d_in > 2) ? __ldg(zp + 2) : 0.f;
    }
    float beta[PT][3];
#pragma unroll
    for (int p = 0; p < PT; ++p)
#pragma unroll
      for (int i = 0; i < 3; ++i)
        beta[p][i] = (C::LAP && net.lap_aux[i] >= 0) ? __ldg(L.aux + gp[p] * L.n_aux + net.lap_aux[i]) : net.lap_beta[i];

    float acc[K][PT][8];
    // ---------------- forward through the hidden layers (layer 0 = feature layer)
#pragma unroll 1
    for (int l = 0; l < Lh; ++l) {
      if (l == 0) {
        // A_c = scl * hjet_c . W0  (software.py:178)
        float w0[3][8];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(L.wpack + net.off_w0 + f * WP + ua));
          const float4 b = __ldg(reinterpret_cast<const float4*>(L.wpack + net.off_w0 + f * WP + WP / 2 + ua));
          w0[f][0] = a.x; w0[f][1] = a.y; w0[f][2] = a.z; w0[f][3] = a.w;
          w0[f][4] = b.x; w0[f][5] = b.y; w0[f][6] = b.z; w0[f][7] = b.w;
        }
#pragma unroll
        for (int p = 0; p < PT; ++p) {
          float hj[K][3];
          feature_jets<C>(net, z[p], beta[p], hj);
#pragma unroll
          for (int c = 0; c < K; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              acc[c][p][j] = net.scl * fmaf(hj[c][0], w0[0][j], fmaf(hj[c][1], w0[1][j], hj[c][2] * w0[2][j]));
        }
      } else {
        __syncthreads();  // previous readers of Hs are done
        store_tile<C>(Hs, acc, row, ua);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
          for (int p = 0; p < PT; ++p)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[c][p][j] = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          const float* wc = chunk_begin();
          gemm_chunk<C>(acc, Hs + row * SP, wc, ua, ch * KC);
          __syncthreads();
          ++gpos;
        }
      }
      act_forward<C, TRAIN>(acc, L.wpack + net.off_b[l], l == 0 ? net.act_first : net.act_hidden,
                            stash + l * STL, row, ua, beta);
    }

    // ---------------- output layer (software.py:183, 215) + residual program
    float wl[8];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(L.wpack + net.off_wl + ua));
      const float4 b = __ldg(reinterpret_cast<const float4*>(L.wpack + net.off_wl + WP / 2 + ua));
      wl[0] = a.x; wl[1] = a.y; wl[2] = a.z; wl[3] = a.w; wl[4] = b.x; wl[5] = b.y; wl[6] = b.z; wl[7] = b.w;
    }
    const float bl = __ldg(L.wpack + net.off_bl);
    float ubar[K][PT];
    {
      float u[K][PT], f[PT], df[K][PT];
      const float* auxp[PT];
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        auxp[p] = L.aux ? (L.aux + gp[p] * L.n_aux) : nullptr;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) s = fmaf(acc[c][p][j], wl[j], s);
#pragma unroll
          for (int o = C::UT / 2; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          u[c][p] = net.epsil * (s + (c == 0 ? bl : 0.f));
          if (L.base) u[c][p] += __ldg(L.base + gp[p] * K + c);
        }
      }
      vm_run<K, PT>(s_ops, L.prog.n_ops, s_consts, z, auxp, u, f, df);
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        if (TRAIN) {
          const float sc = valid[p] ? __ldg(L.seg_scale + slot) : 0.f;
#pragma unroll
          for (int c = 0; c < K; ++c) ubar[c][p] = sc * f[p] * df[c][p];
          if (ut == 0 && valid[p]) lcur += (double)f[p] * (double)f[p];
        } else if (ut == 0 && valid[p]) {
          if (L.out_u) L.out_u[gp[p]] = u[0][p];
          if (L.out_f) L.out_f[gp[p]] = f[p];
          if (L.out_jets)
            for (int c = 0; c < K; ++c) L.out_jets[gp[p] * K + c] = u[c][p];
        }
      }
    }

    if (TRAIN) {
      // output-layer gradients and the adjoint of the last hidden layer's outputs
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        if (ut == 0) blacc += net.epsil * ubar[0][p];
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const float e = net.epsil * ubar[c][p];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            wlacc[j] = fmaf(e, acc[c][p][j], wlacc[j]);
            acc[c][p][j] = e * wl[j];
          }
        }
      }
      // ---------------- backward through the layers
#pragma unroll 1
      for (int l = Lh - 1; l >= 0; --l) {
        act_backward<C>(acc, l == 0 ? net.act_first : net.act_hidden, stash + l * STL, row, ua, beta);
        if (l == 0) break;
        __syncthreads();  // previous readers of Hs/Gs are done
        store_tile<C>(Gs, acc, row, ua);
        recompute_outputs<C>(Hs, (l - 1 == 0) ? net.act_first : net.act_hidden, stash + (l - 1) * STL, row, ua, beta);
        __syncthreads();
        wgrad_layer<C>(Hs, Gs, bsc, gacc + net.off_w[l], gacc + net.off_b[l], tid);
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
          for (int p = 0; p < PT; ++p)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[c][p][j] = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          const float* wc = chunk_begin();
          gemm_chunk<C>(acc, Gs + row * SP, wc, ua, ch * KC);
          __syncthreads();
          ++gpos;
        }
      }
      // ---------------- first layer gradients (acc = adjoint of the first pre-activations)
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        float hj[K][3];
        feature_jets<C>(net, z[p], beta[p], hj);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          b0acc[j] += acc[0][p][j];
#pragma unroll
          for (int c = 0; c < K; ++c) {
            const float a = net.scl * acc[c][p][j];
            w0acc[0][j] = fmaf(hj[c][0], a, w0acc[0][j]);
            w0acc[1][j] = fmaf(hj[c][1], a, w0acc[1][j]);
            w0acc[2][j] = fmaf(hj[c][2], a, w0acc[2][j]);
          }
        }
      }
    }
  }

  if (TRAIN) {
    if (cur_slot >= 0) flush_loss(cur_slot);
    // ---------------- fold the per-thread small accumulators (fixed order => deterministic)
    __syncthreads();
    float* sc = Hs;  // [5][ROWS][WP]
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const float* src = (q < 3) ? w0acc[q] : (q == 3 ? b0acc : wlacc);
      float* d = sc + (q * ROWS + row) * WP;
      *reinterpret_cast<float4*>(d + ua) = make_float4(src[0], src[1], src[2], src[3]);
      *reinterpret_cast<float4*>(d + WP / 2 + ua) = make_float4(src[4], src[5], src[6], src[7]);
    }
    float* sc2 = sc + 5 * ROWS * WP;  // [ROWS]
    if (ut == 0) sc2[row] = blacc;
    __syncthreads();
    for (int idx = tid; idx < 5 * WP; idx += C::NT) {
      const int q = idx / WP, u = idx % WP;
      float s = 0.f;
      for (int r = 0; r < ROWS; ++r) s += sc[(q * ROWS + r) * WP + u];
      const int dst = (q < 3) ? (net.off_w0 + q * WP + u) : (q == 3 ? net.off_b0 + u : net.off_wl + u);
      gacc[dst] += s;
    }
    if (tid == 0) {
      float s = 0.f;
      for (int r = 0; r < ROWS; ++r) s += sc2[r];
      gacc[net.off_bl] += s;
    }
  }
}
