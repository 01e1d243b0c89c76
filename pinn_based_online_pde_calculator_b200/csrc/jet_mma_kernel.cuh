// Tensor-core variant of the fused jet-MLP kernel: the three hidden-layer GEMMs per layer
// (forward A_c = Y_c W, data gradient Ybar = Abar W^T, weight gradient Wbar = Y^T Abar) run on
// warp-level mma.sync m16n8k8 TF32 (SASS HMMA.1688.F32.TF32) with 3xTF32 error compensation:
// every operand is split in REGISTERS after the shared-memory load (hi = top 19 bits,
// lo = x - hi, exact) and a*b is accumulated as hi*hi + lo*hi + hi*lo in fp32.
// Splitting in registers keeps ONE fp32 copy of the activations in shared memory, which is
// what lets forward + backward of a point tile stay on chip.
//
// Same math, stash, gradient accumulators and launch interface as jet_kernel.cuh; only the
// thread <-> element mapping follows the mma fragment layout:
//   lane = 4*g + t ;  thread owns points {16*mw + g, +8} and units {32*nw + 8*nt + 2t, +1}, nt=0..3
// Shared tiles are [pt][c][k ^ swz(pt)] with stride == 8 (mod 32) and swz = 4*((pt>>2)&1), which
// makes every fragment load of the three GEMMs bank-conflict free.
#pragma once
#include "jet_kernel.cuh"

#ifndef PINN_TMEM_COLS
#define PINN_TMEM_COLS 256   // TMEM columns per CTA of the (optional) Tensor-Memory stash
#endif

template <int WP_, int N1_, int N2_, int MIX_, int NT_>
struct MmaCfg {
  static constexpr int WP = WP_, N1 = N1_, N2 = N2_, MIX = MIX_;
  static constexpr bool LAP = (MIX == 2);
  static constexpr int K = 1 + N1 + N2 + (MIX ? 1 : 0);
  static constexpr int NT = NT_;
  static constexpr int PT = 2;                  // fragment rows g and g+8
  static constexpr int NWARP = NT / 32;
  static constexpr int NW = WP / 32;            // warps across units (32 units = 4 n-tiles each)
  static constexpr int MW = NWARP / NW;         // warps across points (16 points each)
  static constexpr int TP = MW * 16;            // points per tile
  static constexpr int ROWS = TP / 2;           // thread rows (mw, g)
  static constexpr int SP = K * WP + 8;         // smem point stride == 8 (mod 32)
  static constexpr int WPS = WP + 8;            // weight row stride (smem chunk and pack)
  // Experiment knobs (padded width 64).  Three CTAs per SM -- 8-row weight chunks (73 KB of shared memory) and a 168-register
  // cap, 12 instead of 8 warps per SM -- measured SLOWER on every width-64 workload: C2 6.00 -> 7.20 ms, C3 54.1 -> 65.2 ms,
  // R0 99.6 -> 148.5 us (the eight mbarrier round trips per weight matrix and the register cap cost more than the extra
  // warps hide).
#ifndef PINN_MMA_KC
#define PINN_MMA_KC 0       // weight rows per chunk (0 = the tuned table below)
#endif
#ifndef PINN_MMA_REGCAP
#define PINN_MMA_REGCAP 256 // registers per thread the occupancy target assumes
#endif
  static constexpr int KC = (PINN_MMA_KC > 0 && WP_ == 64) ? PINN_MMA_KC : (NT <= 64) ? 16 : (K >= 6) ? (WP <= 64 ? 16 : 8) : (WP <= 64 ? (K <= 4 ? 64 : 32) : (WP <= 128 && K <= 4) ? 32 : 16);  // W = 128, K <= 4: 32 keeps 2 CTAs/SM (101 KB); W = 256, K <= 5: 16 fits (199 KB)
  static constexpr int NCH = WP / KC;
  static constexpr uint32_t CHUNK_BYTES = KC * WPS * 4;
  static constexpr int SCR_HALF = ((5 * ROWS * WP + ROWS + 1) / 2 + 3) / 4 * 4;  // final-fold scratch / 2
  static constexpr int HS_FLOATS = (TP * SP > SCR_HALF) ? TP * SP : SCR_HALF;
  // weight-gradient work items: (32 k-rows) x (32 units)
  static constexpr int WITEMS = (WP / 32) * (WP / 32);
  static constexpr int WPASS = (WITEMS + NWARP - 1) / NWARP;
  static constexpr bool OK = (WP >= 64) && (NWARP % NW == 0) && (MW >= 1) && (WITEMS % NWARP == 0 || WITEMS < NWARP) &&
                             (5 * ROWS * WP + ROWS <= 2 * HS_FLOATS);
  static constexpr size_t smem_bytes(bool train) {
    return (size_t)(HS_FLOATS * (train ? 2 : 1) + 2 * KC * WPS + WP + PINN_MAX_OPS + PINN_MAX_CONSTS) * 4 + 64;
  }
  // TMEM stash: 4-warp CTAs only (one warp per 32-lane quadrant); 16*K columns per layer, 256 per CTA
  // Measured on B200 (round 1): merely executing tcgen05.alloc in this kernel -- even 32 columns that
  // are never touched -- slows the mma.sync (HMMA) GEMM phases by 1.4x (C2: 7.06 -> 10.0 ms/step), so
  // legacy warp-level MMA and explicit Tensor-Memory allocations do not mix; the TMEM stash is therefore
  // compiled out by default (-DPINN_TMEM_STASH=1 re-enables it for experiments).
#ifndef PINN_TMEM_STASH
#define PINN_TMEM_STASH 0
#endif
  static constexpr bool USE_TMEM = (PINN_TMEM_STASH != 0) && (NT == 128);
  static constexpr int TMEM_LAYERS = PINN_TMEM_COLS / (16 * K);
  static constexpr int MINB_S = (int)(232448 / (smem_bytes(true) + 1024));            // smem-limited CTAs per SM
  static constexpr int MINB_R = 65536 / (NT * (WP_ == 64 ? PINN_MMA_REGCAP : 256));      // at 255 registers per thread
  static constexpr int MINB_U = MINB_S < 1 ? 1 : (MINB_S < MINB_R ? MINB_S : MINB_R);
  static constexpr int MINB = (USE_TMEM && MINB_U > 2) ? 2 : MINB_U;   // resident CTAs per SM (2 x 256 TMEM columns)
};

// ---------------------------------------------------------------- tensor-core helpers
// Round-to-nearest split: hi = x rounded to the 11 significant bits of TF32 (add half an ulp to the
// bit pattern, clear the 13 low mantissa bits: two integer-pipe ops), lo = x - hi (exact, one fp32
// add).  Rounding (not truncating) keeps the split unbiased, so the residual errors of the 3xTF32
// product accumulate like a random walk instead of coherently.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(__fadd_rn(x, -__uint_as_float(hi)));
}
__device__ __forceinline__ void mma_tf32(float& d0, float& d1, float& d2, float& d3, const uint32_t (&a)[4],
                                         uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// bf16 MMA for the two small split terms (lo*hi, hi*lo): one m16n8k16 covers the 16 k of TWO tf32 k-steps at
// the cost of one tf32 MMA.  The k <-> register-slot assignment inside an MMA is arbitrary as long as A and B
// agree, so the fragments already loaded for the tf32 steps are reused: register 0/1 (rows g, g+8) pack
// (k = t, t+4) of the first step, register 2/3 the second step; B register 0/1 likewise.
__device__ __forceinline__ uint32_t pack_bf16(uint32_t x_lowk, uint32_t x_highk) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(x_highk)), "f"(__uint_as_float(x_lowk)));
  return r;
}
__device__ __forceinline__ void mma_bf16(float& d0, float& d1, float& d2, float& d3, const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// packed fp32x2 add (Blackwell FADD2): (a0, a1) += (b0, b1) in one instruction
__device__ __forceinline__ void fadd2(float& a0, float& a1, float b0, float b1) {
  f32x2_t A = pack2(a0, a1);
  const f32x2_t B = pack2(b0, b1);
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(A) : "l"(B));
  unpack2(A, a0, a1);
}
#ifndef PINN_BF16_SMALL
#define PINN_BF16_SMALL 1
#endif
// d += a*b with 3xTF32 compensation
__device__ __forceinline__ void mma3(float& d0, float& d1, float& d2, float& d3, const uint32_t (&ah)[4],
                                     const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(d0, d1, d2, d3, al, bh0, bh1);
  mma_tf32(d0, d1, d2, d3, ah, bl0, bl1);
  mma_tf32(d0, d1, d2, d3, ah, bh0, bh1);
}

// ---------------------------------------------------------------- TMEM scratchpad (tcgen05.st / tcgen05.ld)
// The backward stash of the first layers lives in Tensor Memory instead of the L2-resident global
// scratch: 256 columns per CTA (two CTAs per SM share the 512), lane = the thread's lane inside
// its warp's 32-lane quadrant, column = (layer, channel, point, unit slot).  SASS: STTM / LDTM.
__device__ __forceinline__ void tmem_alloc256(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"((uint32_t)PINN_TMEM_COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc256(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)PINN_TMEM_COLS) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// thread geometry of the fragment layout
struct MmaGeo {
  int lane, g, t, mw, nw, row;  // row = 8*mw + g (thread row over points)
  int pt0;                      // first owned point (second = pt0 + 8)
  int n0;                       // first unit of the warp (32*nw)
  int swz;                      // column swizzle of the owned points
  uint32_t tmem;                // TMEM address of this warp's lane quadrant, column 0 (0 = TMEM stash off)
};

// unit index of register slot j8 (0..7): n-tile j8/2, column 2t + j8%2
#define MMA_UNIT(geo, j8) ((geo).n0 + 8 * ((j8) >> 1) + 2 * (geo).t + ((j8) & 1))

// acc[c][p][2nt+e] += sum_k S[pt(p)][c][kbase+k] * wc[k][n0 + 8nt + 2t'...]  (fragment layout)
// acc[c][p][2nt+e] += sum_k S[pt(p)][c][kbase+k] * wc[k][n0 + 8nt + 2t + e]  (fragment layout)
template <class C, bool BF16_SMALL>
__device__ __forceinline__ void mma_gemm_chunk(float (&acc)[C::K][2][8], const float* __restrict__ S,
                                               const float* __restrict__ wc, int kbase, const MmaGeo& G) {
  // Two-level accumulation: the TF32 MMAs of KS k-steps accumulate into a partial sum that starts
  // from ZERO; the partial is then added to the running accumulator with fp32 round-to-nearest
  // adds.  The tensor core truncates when it adds into C; doing that on a small partial instead of
  // on the running sum removes the coherent bias that otherwise grows with the reduction length
  // (measured 1.6e-5 on loss terms at W=256, 3e-7 with this scheme).
  constexpr int KS = (C::KC >= 16) ? 2 : 1;
  constexpr int K_ = C::K;
  const int tA = G.t ^ G.swz, tB = tA ^ 4;
#if PINN_BF16_SMALL
  if (BF16_SMALL && KS == 2) {
    // two k-steps per pass: hi*hi on TF32 MMAs (one per k-step), lo*hi and hi*lo on ONE bf16 MMA each
#pragma unroll 1
    for (int kk = 0; kk < C::KC; kk += 16) {
      uint32_t bh[2][4][2], pbh[4][2], pbl[4][2];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float* bp = wc + (kk + 8 * ks + G.t) * C::WPS + G.n0 + 8 * nt + G.g;
          uint32_t l0, l1;
          split_tf32(bp[0], bh[ks][nt][0], l0);
          split_tf32(bp[4 * C::WPS], bh[ks][nt][1], l1);
          pbh[nt][ks] = pack_bf16(bh[ks][nt][0], bh[ks][nt][1]);
          pbl[nt][ks] = pack_bf16(l0, l1);
        }
#pragma unroll
      for (int c0 = 0; c0 < K_; c0 += 2) {
        float tq[2][4][4];
        uint32_t pah[2][4], pal[2][4];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) tq[cc][nt][0] = tq[cc][nt][1] = tq[cc][nt][2] = tq[cc][nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t ah[2][4];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            if (c0 + cc < K_) {
              const float* sp = S + G.pt0 * C::SP + (c0 + cc) * C::WP + kbase + kk + 8 * ks;
              uint32_t al[4];
              split_tf32(sp[tA], ah[cc][0], al[0]);
              split_tf32(sp[8 * C::SP + tA], ah[cc][1], al[1]);
              split_tf32(sp[tB], ah[cc][2], al[2]);
              split_tf32(sp[8 * C::SP + tB], ah[cc][3], al[3]);
              pah[cc][2 * ks] = pack_bf16(ah[cc][0], ah[cc][2]);      // row g:   (k = t, t+4)
              pah[cc][2 * ks + 1] = pack_bf16(ah[cc][1], ah[cc][3]);  // row g+8
              pal[cc][2 * ks] = pack_bf16(al[0], al[2]);
              pal[cc][2 * ks + 1] = pack_bf16(al[1], al[3]);
            }
          }
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            if (c0 + cc < K_) {
#pragma unroll
              for (int nt = 0; nt < 4; ++nt)
                mma_tf32(tq[cc][nt][0], tq[cc][nt][1], tq[cc][nt][2], tq[cc][nt][3], ah[cc], bh[ks][nt][0], bh[ks][nt][1]);
            }
          }
        }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            if (c0 + cc < K_) {
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) {
                if (pass == 0)
                  mma_bf16(tq[cc][nt][0], tq[cc][nt][1], tq[cc][nt][2], tq[cc][nt][3], pal[cc], pbh[nt][0], pbh[nt][1]);
                else
                  mma_bf16(tq[cc][nt][0], tq[cc][nt][1], tq[cc][nt][2], tq[cc][nt][3], pah[cc], pbl[nt][0], pbl[nt][1]);
              }
            }
          }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (c0 + cc < K_) {
            const int c = c0 + cc;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              fadd2(acc[c][0][2 * nt], acc[c][0][2 * nt + 1], tq[cc][nt][0], tq[cc][nt][1]);
              fadd2(acc[c][1][2 * nt], acc[c][1][2 * nt + 1], tq[cc][nt][2], tq[cc][nt][3]);
            }
          }
        }
      }
    }
    return;
  }
#endif
#pragma unroll 1
  for (int kk = 0; kk < C::KC; kk += 8 * KS) {
    uint32_t bh[KS][4][2], bl[KS][4][2];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* bp = wc + (kk + 8 * ks + G.t) * C::WPS + G.n0 + 8 * nt + G.g;
        split_tf32(bp[0], bh[ks][nt][0], bl[ks][nt][0]);
        split_tf32(bp[4 * C::WPS], bh[ks][nt][1], bl[ks][nt][1]);
      }
    // channels two at a time: 8 independent accumulator tiles per pass keep dependent mmas
    // eight instructions apart (HMMA latency), small terms first
#pragma unroll
    for (int c0 = 0; c0 < K_; c0 += 2) {
      float tq[2][4][4];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) tq[cc][nt][0] = tq[cc][nt][1] = tq[cc][nt][2] = tq[cc][nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (c0 + cc < K_) {
            const float* sp = S + G.pt0 * C::SP + (c0 + cc) * C::WP + kbase + kk + 8 * ks;
            split_tf32(sp[tA], ah[cc][0], al[cc][0]);
            split_tf32(sp[8 * C::SP + tA], ah[cc][1], al[cc][1]);
            split_tf32(sp[tB], ah[cc][2], al[cc][2]);
            split_tf32(sp[8 * C::SP + tB], ah[cc][3], al[cc][3]);
          }
        }
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            if (c0 + cc < K_) {
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) {
                if (pass == 0)
                  mma_tf32(tq[cc][nt][0], tq[cc][nt][1], tq[cc][nt][2], tq[cc][nt][3], al[cc], bh[ks][nt][0], bh[ks][nt][1]);
                else if (pass == 1)
                  mma_tf32(tq[cc][nt][0], tq[cc][nt][1], tq[cc][nt][2], tq[cc][nt][3], ah[cc], bl[ks][nt][0], bl[ks][nt][1]);
                else
                  mma_tf32(tq[cc][nt][0], tq[cc][nt][1], tq[cc][nt][2], tq[cc][nt][3], ah[cc], bh[ks][nt][0], bh[ks][nt][1]);
              }
            }
          }
      }
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        if (c0 + cc < K_) {
          const int c = c0 + cc;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            fadd2(acc[c][0][2 * nt], acc[c][0][2 * nt + 1], tq[cc][nt][0], tq[cc][nt][1]);
            fadd2(acc[c][1][2 * nt], acc[c][1][2 * nt + 1], tq[cc][nt][2], tq[cc][nt][3]);
          }
        }
      }
    }
  }
}

// register tile -> smem tile (float2 stores, swizzled columns)
template <class C>
__device__ __forceinline__ void mma_store_tile(float* __restrict__ S, const float (&acc)[C::K][2][8],
                                               const MmaGeo& G) {
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int c = 0; c < C::K; ++c) {
      float* d = S + (G.pt0 + 8 * p) * C::SP + c * C::WP;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        *reinterpret_cast<float2*>(d + ((G.n0 + 8 * nt + 2 * G.t) ^ G.swz)) =
            make_float2(acc[c][p][2 * nt], acc[c][p][2 * nt + 1]);
    }
}

// thread-private, warp-coalesced stash slot of (c, p, q): float4 holding j8 = 4q..4q+3
template <class C>
__device__ __forceinline__ float4* mma_stash_ptr(float* stash_l, int c, int p, int q, int tid) {
  return reinterpret_cast<float4*>(stash_l) + ((size_t)((c * 2 + p) * 2 + q) * C::NT + tid);
}

// activation jets, forward (same math as act_forward in jet_kernel.cuh)
template <class C, bool TRAIN>
__device__ __forceinline__ void mma_act_forward(float (&acc)[C::K][2][8], const float* __restrict__ bias, int act,
                                                float* __restrict__ stash_l, const MmaGeo& G, int tid,
                                                const float (&beta)[2][3], int tm_col) {
  float b[8];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(bias + G.n0 + 8 * nt + 2 * G.t));
    b[2 * nt] = v.x; b[2 * nt + 1] = v.y;
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float y[8], d1[8], d2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s0;
      act_fwd(act, acc[0][p][j] + b[j], y[j], d1[j], d2[j], s0);
      acc[0][p][j] = s0;
    }
    if (TRAIN) {
      if (tm_col >= 0) {  // layer stash in Tensor Memory
#pragma unroll
        for (int c = 0; c < C::K; ++c) tmem_st8(G.tmem + tm_col + (c * 2 + p) * 8, acc[c][p]);
      } else {
#pragma unroll
        for (int c = 0; c < C::K; ++c) {
          *mma_stash_ptr<C>(stash_l, c, p, 0, tid) = make_float4(acc[c][p][0], acc[c][p][1], acc[c][p][2], acc[c][p][3]);
          *mma_stash_ptr<C>(stash_l, c, p, 1, tid) = make_float4(acc[c][p][4], acc[c][p][5], acc[c][p][6], acc[c][p][7]);
        }
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
  if (TRAIN && tm_col >= 0) tmem_wait_st();
}

template <class C>
__device__ __forceinline__ void mma_load_stash(float (&st)[C::K][8], const float* __restrict__ stash_l, int p, int tid,
                                               uint32_t tmem, int tm_col) {
  if (tm_col >= 0) {
#pragma unroll
    for (int c = 0; c < C::K; ++c) tmem_ld8(tmem + tm_col + (c * 2 + p) * 8, st[c]);
    tmem_wait_ld();
    return;
  }
#pragma unroll
  for (int c = 0; c < C::K; ++c) {
    const float4 v0 = *mma_stash_ptr<C>(const_cast<float*>(stash_l), c, p, 0, tid);
    const float4 v1 = *mma_stash_ptr<C>(const_cast<float*>(stash_l), c, p, 1, tid);
    st[c][0] = v0.x; st[c][1] = v0.y; st[c][2] = v0.z; st[c][3] = v0.w;
    st[c][4] = v1.x; st[c][5] = v1.y; st[c][6] = v1.z; st[c][7] = v1.w;
  }
}

// adjoint of the activation jets (same math as act_backward)
template <class C>
__device__ __forceinline__ void mma_act_backward(float (&acc)[C::K][2][8], int act, const float* __restrict__ stash_l,
                                                 int tid, const float (&beta)[2][3], uint32_t tmem, int tm_col) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float st[C::K][8];
    mma_load_stash<C>(st, stash_l, p, tid, tmem, tm_col);
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

// recompute a layer's output jets from its stash into the smem tile S
template <class C>
__device__ __forceinline__ void mma_recompute_outputs(float* __restrict__ S, int act, const float* __restrict__ stash_l,
                                                      const MmaGeo& G, int tid, const float (&beta)[2][3], int tm_col) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float st[C::K][8];
    mma_load_stash<C>(st, stash_l, p, tid, G.tmem, tm_col);
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
      float* d = S + (G.pt0 + 8 * p) * C::SP + c * C::WP;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        *reinterpret_cast<float2*>(d + ((G.n0 + 8 * nt + 2 * G.t) ^ G.swz)) = make_float2(st[c][2 * nt], st[c][2 * nt + 1]);
    }
  }
}

// weight gradient of one hidden layer on tensor cores:
//   gW[k][u] += sum_{pt,c} H[pt][c][k] * G[pt][c][u],  gB[u] += sum_pt G[pt][0][u]
template <class C>
__device__ __forceinline__ void mma_wgrad_layer(const float* __restrict__ Hs, const float* __restrict__ Gs,
                                                float* __restrict__ bsc, float* __restrict__ gW,
                                                float* __restrict__ gB, int tid, const MmaGeo& G) {
  // bias gradient: thread u sums the value-channel adjoint over the tile's points
  for (int u = tid; u < C::WP; u += C::NT) {
    float b0 = 0.f, b1 = 0.f;
#pragma unroll 4
    for (int pt = 0; pt < C::TP; pt += 2) {
      b0 += Gs[pt * C::SP + (u ^ (((pt >> 2) & 1) << 2))];
      b1 += Gs[(pt + 1) * C::SP + (u ^ ((((pt + 1) >> 2) & 1) << 2))];
    }
    gB[u] += b0 + b1;
  }
  (void)bsc;
  const int warp = tid >> 5;
  // work item = (32 k-rows = 2 m-tiles) x (32 units = 4 n-tiles): 8 accumulator tiles per warp,
  // 8 A + 8 B fragment elements per reduction step
#pragma unroll 1
  for (int item = warp; item < C::WITEMS; item += C::NWARP) {
    const int mt = item / (C::WP / 32), ng = item % (C::WP / 32);
    const int k0 = 32 * mt, u0 = 32 * ng;
    float w[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) w[m][nt][0] = w[m][nt][1] = w[m][nt][2] = w[m][nt][3] = 0.f;
#pragma unroll 1
    for (int ps = 0; ps < C::TP / 8; ++ps) {
      const int pa = 8 * ps + G.t, pb = pa + 4;  // swizzle 0 for pa, 4 for pb
#pragma unroll
      for (int c0 = 0; c0 < C::K; c0 += 2) {
        float tq[2][4][4];  // two-level accumulation (see mma_gemm_chunk): flushed every two reduction steps
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) tq[m][nt][0] = tq[m][nt][1] = tq[m][nt][2] = tq[m][nt][3] = 0.f;
#if PINN_BF16_SMALL
        // the two channels of the pair are the two k-steps of ONE bf16 MMA for each small split term
        uint32_t pah[2][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}}, pal[2][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}};
        uint32_t pbh[4][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}}, pbl[4][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}};
#endif
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          if (c0 + cc < C::K) {
            const int c = c0 + cc;
            const float* ha = Hs + pa * C::SP + c * C::WP;
            const float* hb = Hs + pb * C::SP + c * C::WP;
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              const int kr = k0 + 16 * m + G.g;
              split_tf32(ha[kr], ah[m][0], al[m][0]);
              split_tf32(ha[kr + 8], ah[m][1], al[m][1]);
              split_tf32(hb[kr ^ 4], ah[m][2], al[m][2]);
              split_tf32(hb[(kr + 8) ^ 4], ah[m][3], al[m][3]);
            }
            const float* ga = Gs + pa * C::SP + c * C::WP + u0 + G.g;
            const float* gb = Gs + pb * C::SP + c * C::WP + ((u0 + G.g) ^ 4);
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              split_tf32(ga[8 * nt], bh[nt][0], bl[nt][0]);
              split_tf32(gb[8 * nt], bh[nt][1], bl[nt][1]);
            }
#if PINN_BF16_SMALL
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              pah[m][2 * cc] = pack_bf16(ah[m][0], ah[m][2]);
              pah[m][2 * cc + 1] = pack_bf16(ah[m][1], ah[m][3]);
              pal[m][2 * cc] = pack_bf16(al[m][0], al[m][2]);
              pal[m][2 * cc + 1] = pack_bf16(al[m][1], al[m][3]);
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              pbh[nt][cc] = pack_bf16(bh[nt][0], bh[nt][1]);
              pbl[nt][cc] = pack_bf16(bl[nt][0], bl[nt][1]);
            }
#else
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) mma_tf32(tq[m][nt][0], tq[m][nt][1], tq[m][nt][2], tq[m][nt][3], al[m], bh[nt][0], bh[nt][1]);
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) mma_tf32(tq[m][nt][0], tq[m][nt][1], tq[m][nt][2], tq[m][nt][3], ah[m], bl[nt][0], bl[nt][1]);
#endif
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
              for (int nt = 0; nt < 4; ++nt) mma_tf32(tq[m][nt][0], tq[m][nt][1], tq[m][nt][2], tq[m][nt][3], ah[m], bh[nt][0], bh[nt][1]);
          }
        }
#if PINN_BF16_SMALL
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_bf16(tq[m][nt][0], tq[m][nt][1], tq[m][nt][2], tq[m][nt][3], pal[m], pbh[nt][0], pbh[nt][1]);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_bf16(tq[m][nt][0], tq[m][nt][1], tq[m][nt][2], tq[m][nt][3], pah[m], pbl[nt][0], pbl[nt][1]);
#endif
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            fadd2(w[m][nt][0], w[m][nt][1], tq[m][nt][0], tq[m][nt][1]);
            fadd2(w[m][nt][2], w[m][nt][3], tq[m][nt][2], tq[m][nt][3]);
          }
      }
    }
    // read-modify-write the CTA-private accumulator (rows k0+16m+g, +8; cols u0+8nt+2t, +1)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float2* d0 = reinterpret_cast<float2*>(gW + (size_t)(k0 + 16 * m + G.g) * C::WPS + u0 + 8 * nt + 2 * G.t);
        float2* d1 = reinterpret_cast<float2*>(gW + (size_t)(k0 + 16 * m + G.g + 8) * C::WPS + u0 + 8 * nt + 2 * G.t);
        float2 v0 = *d0, v1 = *d1;
        v0.x += w[m][nt][0]; v0.y += w[m][nt][1];
        v1.x += w[m][nt][2]; v1.y += w[m][nt][3];
        *d0 = v0; *d1 = v1;
      }
  }
}

// ---------------------------------------------------------------- the kernel
template <class C, bool TRAIN>
__global__ void __launch_bounds__(C::NT, C::MINB) jet_mma_kernel(const __grid_constant__ PinnLaunch L) {
  static_assert(C::OK, "invalid mma kernel configuration");
  constexpr int K = C::K, WP = C::WP, NCH = C::NCH, KC = C::KC, WPS = C::WPS;
  extern __shared__ __align__(128) float smem[];
  float* Hs = smem;
  float* Gs = Hs + C::HS_FLOATS;
  float* Wc = TRAIN ? (Gs + C::HS_FLOATS) : Gs;
  float* bsc = Wc + 2 * KC * WPS;
  int* s_ops = reinterpret_cast<int*>(bsc + WP);
  float* s_consts = reinterpret_cast<float*>(s_ops + PINN_MAX_OPS);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_consts + PINN_MAX_CONSTS);

  const PinnNet& net = L.net;
  const int tid = threadIdx.x;
  MmaGeo G;
  G.lane = tid & 31; G.g = G.lane >> 2; G.t = G.lane & 3;
  const int warp = tid >> 5;
  G.nw = warp % C::NW; G.mw = warp / C::NW;
  G.row = 8 * G.mw + G.g;
  G.pt0 = 16 * G.mw + G.g;
  G.n0 = 32 * G.nw;
  G.swz = ((G.g >> 2) & 1) << 2;
  G.tmem = 0;
  const int Lh = net.n_hidden;
  const int nF = (Lh - 1) * NCH;
  const int S = TRAIN ? 2 * nF : nF;
  const int my_tiles = (L.n_tiles > (int)blockIdx.x) ? (L.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total = (long long)my_tiles * S;
  long long gpos = 0;

  auto issue = [&](long long gq) {
    const int p = (int)(gq % S);
    const float* src;
    if (p < nF) {
      const int l = 1 + p / NCH, ch = p % NCH;
      src = L.wpack + net.off_w[l] + ch * (KC * WPS);
    } else {
      const int q = p - nF;
      const int l = (Lh - 1) - q / NCH, ch = q % NCH;
      src = L.wpack + net.off_wt[l] + ch * (KC * WPS);
    }
    const int st = (int)(gq & 1);
    mbar_expect_tx(&mbar[st], C::CHUNK_BYTES);
    bulk_g2s(Wc + st * (KC * WPS), src, C::CHUNK_BYTES, &mbar[st]);
  };

  for (int i = tid; i < L.prog.n_ops; i += C::NT) s_ops[i] = L.prog.ops[i];
  for (int i = tid; i < PINN_MAX_CONSTS; i += C::NT) s_consts[i] = L.prog.consts[i];
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr bool TM = TRAIN && C::USE_TMEM;
  __shared__ uint32_t s_tmem_base;
  if (TM && warp == 0) tmem_alloc256(&s_tmem_base);
  if (TM) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (TM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    G.tmem = s_tmem_base + ((uint32_t)(warp & 3) << 21);  // lane field = 32 * (warp % 4), bits [31:16]
  }
  // stash column of layer l in TMEM, or -1 when the layer's stash stays in the global scratch
  auto tm_col = [&](int l) -> int { return (TM && l < C::TMEM_LAYERS) ? l * (16 * K) : -1; };
  if (tid == 0 && total > 0) issue(0);

  constexpr size_t STL = (size_t)C::TP * K * WP;  // stash floats per layer (= K*16*NT)
  float* stash = TRAIN ? (L.stash + (size_t)blockIdx.x * Lh * STL) : nullptr;
  float* gacc = TRAIN ? (L.gacc + (size_t)blockIdx.x * net.pg) : nullptr;

  float w0acc[3][8], b0acc[8], wlacc[8], blacc = 0.f;
  double lcur = 0.0;
  int cur_slot = -1;
  if (TRAIN) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { w0acc[0][j] = w0acc[1][j] = w0acc[2][j] = 0.f; b0acc[j] = 0.f; wlacc[j] = 0.f; }
  }
  auto flush_loss = [&](int slot) {
    double* dsc = reinterpret_cast<double*>(Gs);
    __syncthreads();
    if (G.t == 0 && G.nw == 0) { dsc[2 * G.row] = lcur; }
    __syncthreads();
    if (tid == 0) {
      double tsum = 0.0;
      for (int r = 0; r < C::ROWS; ++r) tsum += dsc[2 * r];
      L.loss_part[(size_t)blockIdx.x * L.n_slots + slot] += tsum;
    }
    __syncthreads();
    lcur = 0.0;
  };
  const bool prof = (L.phase_clk != nullptr) && blockIdx.x == 0 && tid == 0;
  long long pclk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tmark = prof ? clock64() : 0;
  auto lap = [&](int ph) {
    if (prof) { const long long now = clock64(); pclk[ph] += now - tmark; tmark = now; }
  };
  auto chunk_begin = [&]() -> const float* {
    if (tid == 0 && gpos + 1 < total) issue(gpos + 1);
    const int st = (int)(gpos & 1);
    mbar_wait(&mbar[st], (uint32_t)((gpos >> 1) & 1));
    return Wc + st * (KC * WPS);
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
    if (TRAIN && slot != cur_slot) {
      if (cur_slot >= 0) flush_loss(cur_slot);
      cur_slot = slot;
    }

    float z[2][3];
    long long gp[2];
    bool valid[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int lp = G.pt0 + 8 * p;
      valid[p] = lp < cnt;
      gp[p] = pbegin + (valid[p] ? lp : 0);
      const float* zp = L.coords + gp[p] * net.d_in;
      z[p][0] = __ldg(zp);
      z[p][1] = (net.d_in > 1) ? __ldg(zp + 1) : 0.f;
      z[p][2] = (net.d_in > 2) ? __ldg(zp + 2) : 0.f;
    }
    float beta[2][3];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int i = 0; i < 3; ++i)
        beta[p][i] = (C::LAP && net.lap_aux[i] >= 0) ? __ldg(L.aux + gp[p] * L.n_aux + net.lap_aux[i]) : net.lap_beta[i];

    float acc[K][2][8];
#pragma unroll 1
    for (int l = 0; l < Lh; ++l) {
      if (l == 0) {
        float w0[3][8];
#pragma unroll
        for (int f = 0; f < 3; ++f)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(L.wpack + net.off_w0 + f * WP + G.n0 + 8 * nt + 2 * G.t));
            w0[f][2 * nt] = v.x; w0[f][2 * nt + 1] = v.y;
          }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          float hj[K][3];
          feature_jets<C>(net, z[p], beta[p], hj);
#pragma unroll
          for (int c = 0; c < K; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              acc[c][p][j] = net.scl * fmaf(hj[c][0], w0[0][j], fmaf(hj[c][1], w0[1][j], hj[c][2] * w0[2][j]));
        }
      } else {
        __syncthreads();
        mma_store_tile<C>(Hs, acc, G);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[c][p][j] = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          const float* wc = chunk_begin();
          mma_gemm_chunk<C, false>(acc, Hs, wc, ch * KC, G);  // forward: full 3xTF32 (loss / residual precision)
          __syncthreads();
          ++gpos;
        }
      }
      lap(0);
      mma_act_forward<C, TRAIN>(acc, L.wpack + net.off_b[l], l == 0 ? net.act_first : net.act_hidden,
                                stash + l * STL, G, tid, beta, tm_col(l));
      lap(1);
    }

    // ---------------- output layer + residual program
    float wl[8];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(L.wpack + net.off_wl + G.n0 + 8 * nt + 2 * G.t));
      wl[2 * nt] = v.x; wl[2 * nt + 1] = v.y;
    }
    const float bl = __ldg(L.wpack + net.off_bl);
    float ubar[K][2];
    {
      // partial dots over the thread's 8 units -> reduce over the 4 t-lanes (shuffle) and the NW warps (smem)
      float u[K][2], f[2], df[K][2];
      const float* auxp[2];
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int c = 0; c < K; ++c) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) s = fmaf(acc[c][p][j], wl[j], s);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          u[c][p] = s;
        }
      if (C::NW > 1) {
        // cross-warp fold through smem: part[nw][pt][c]
        float* part = Hs;  // the last GEMM that read Hs has completed (barrier after its last chunk)
        __syncthreads();
        if (G.t == 0) {
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int c = 0; c < K; ++c) part[(G.nw * C::TP + G.pt0 + 8 * p) * K + c] = u[c][p];
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int c = 0; c < K; ++c) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < C::NW; ++w) s += part[(w * C::TP + G.pt0 + 8 * p) * K + c];
            u[c][p] = s;
          }
        __syncthreads();
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        auxp[p] = L.aux ? (L.aux + gp[p] * L.n_aux) : nullptr;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          u[c][p] = net.epsil * (u[c][p] + (c == 0 ? bl : 0.f));
          if (L.base) u[c][p] += __ldg(L.base + gp[p] * K + c);
        }
      }
      vm_run<K, 2>(s_ops, L.prog.n_ops, s_consts, z, auxp, u, f, df);
      const bool owner = (G.t == 0 && G.nw == 0);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        if (TRAIN) {
          const float sc = valid[p] ? __ldg(L.seg_scale + slot) : 0.f;
#pragma unroll
          for (int c = 0; c < K; ++c) ubar[c][p] = sc * f[p] * df[c][p];
          if (owner && valid[p]) lcur += (double)f[p] * (double)f[p];
        } else if (owner && valid[p]) {
          if (L.out_u) L.out_u[gp[p]] = u[0][p];
          if (L.out_f) L.out_f[gp[p]] = f[p];
          if (L.out_jets)
            for (int c = 0; c < K; ++c) L.out_jets[gp[p] * K + c] = u[c][p];
        }
      }
    }

    lap(2);
    if (TRAIN) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        if (G.t == 0 && G.nw == 0) blacc += net.epsil * ubar[0][p];
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
#pragma unroll 1
      for (int l = Lh - 1; l >= 0; --l) {
        mma_act_backward<C>(acc, l == 0 ? net.act_first : net.act_hidden, stash + l * STL, tid, beta, G.tmem, tm_col(l));
        lap(3);
        if (l == 0) break;
        __syncthreads();
        mma_store_tile<C>(Gs, acc, G);
        mma_recompute_outputs<C>(Hs, (l - 1 == 0) ? net.act_first : net.act_hidden, stash + (l - 1) * STL, G, tid, beta, tm_col(l - 1));
        __syncthreads();
        lap(4);
        mma_wgrad_layer<C>(Hs, Gs, bsc, gacc + net.off_w[l], gacc + net.off_b[l], tid, G);
        lap(5);
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[c][p][j] = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          const float* wc = chunk_begin();
          mma_gemm_chunk<C, true>(acc, Gs, wc, ch * KC, G);  // data gradient: small split terms on bf16 MMAs
          __syncthreads();
          ++gpos;
        }
        lap(6);
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) {
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

  lap(7);
  if (prof) {
#pragma unroll
    for (int i = 0; i < 8; ++i) L.phase_clk[i] = pclk[i];
  }
  if (TRAIN) {
    if (cur_slot >= 0) flush_loss(cur_slot);
    __syncthreads();
    float* sc = Hs;  // [5][ROWS][WP], folded over the ROWS thread rows in fixed order
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const float* src = (q < 3) ? w0acc[q] : (q == 3 ? b0acc : wlacc);
      float* d = sc + (q * C::ROWS + G.row) * WP;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        *reinterpret_cast<float2*>(d + G.n0 + 8 * nt + 2 * G.t) = make_float2(src[2 * nt], src[2 * nt + 1]);
    }
    float* sc2 = sc + 5 * C::ROWS * WP;
    if (G.t == 0 && G.nw == 0) sc2[G.row] = blacc;
    __syncthreads();
    for (int idx = tid; idx < 5 * WP; idx += C::NT) {
      const int q = idx / WP, u = idx % WP;
      float s = 0.f;
      for (int r = 0; r < C::ROWS; ++r) s += sc[(q * C::ROWS + r) * WP + u];
      const int dst = (q < 3) ? (net.off_w0 + q * WP + u) : (q == 3 ? net.off_b0 + u : net.off_wl + u);
      gacc[dst] += s;
    }
    if (tid == 0) {
      float s = 0.f;
      for (int r = 0; r < C::ROWS; ++r) s += sc2[r];
      gacc[net.off_bl] += s;
    }
  }
  if (TM) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc256(s_tmem_base);
  }
}
