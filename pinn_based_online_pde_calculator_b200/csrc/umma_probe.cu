// tcgen05 probe: runs D = A * B^T (tf32 inputs, fp32 accumulate in TMEM) on ONE CTA for operands given
// as RAW shared-memory images (the host lays them out) or, for A, as a row-major matrix loaded into
// TMEM (TS form), with explicit descriptor fields, and dumps the whole TMEM tile.  Used by
// tools/umma_probe.py to pin down, on the real chip, (1) the K-major / MN-major descriptor conventions
// and swizzle modes, (2) the TMEM lane layout of M=128 and of two interleaved M=64 tiles, (3) how inputs
// and the accumulator round, (4) the issue rate.  Measurement helper, not on the product path.
#include <stdio.h>

#include "../../include/pinn_engine.h"
#include "umma_common.cuh"

namespace {

struct ProbeArgs {
  const float* A;   // raw shared-memory image of operand A (a_words floats)
  const float* B0;  // raw image of operand B
  const float* B1;  // second B image (second product, lanes +16; M = 64 only)
  int a_words, b_words;
  int M, N, ksteps, a_mn, b_mn, nsets, reps, nd, a_lt, b_lt, a_tmem, bf16;
  uint32_t a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step, a_step2, b_step2;
  float* out;  // [128][512]
  long long* cycles;
  int* status;
};

__global__ void __launch_bounds__(128) umma_probe_kernel(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* As = reinterpret_cast<float*>(smem);
  float* Bs0 = As + p.a_words;
  float* Bs1 = Bs0 + p.b_words;
  for (int i = tid; i < p.a_words; i += 128) As[i] = p.A[i];
  for (int i = tid; i < p.b_words; i += 128) {
    Bs0[i] = p.B0[i];
    if (p.nsets > 1) Bs1[i] = p.B1[i];
  }
  umma::fence_async_smem();
  if (warp == 0) umma::tmem_alloc(&tbase, 512);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tbase;
  if (p.a_tmem) {
    // A given row-major [128][8*ksteps]: lane = row, columns 256.. of TMEM
    for (int col = 0; col < 8 * p.ksteps; col += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = p.A[(size_t)tid * 8 * p.ksteps + col + i];
      umma::tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + 256 + col, v);
    }
    umma::tmem_st_wait();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (tid == 0) {
    const uint32_t idesc = p.bf16 ? umma::idesc_bf16(p.M, p.N, p.a_mn, p.b_mn) : umma::idesc_tf32(p.M, p.N, p.a_mn, p.b_mn);
    // descriptors precomputed; the issue loop only bumps the 14-bit start-address field
    const uint64_t a0 = umma::smem_desc(umma::smem_addr(As), p.a_lbo, p.a_sbo, p.a_lt);
    const uint64_t b0[2] = {umma::smem_desc(umma::smem_addr(Bs0), p.b_lbo, p.b_sbo, p.b_lt),
                            umma::smem_desc(umma::smem_addr(Bs1), p.b_lbo, p.b_sbo, p.b_lt)};
    const uint32_t da = p.a_step >> 4, db = p.b_step >> 4;
    const int nd = p.nd < 1 ? 1 : p.nd;  // round-robin over nd accumulator regions (timing only)
    const long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep)
      for (int s = 0; s < p.nsets; ++s) {
        const uint32_t d = tb + ((uint32_t)(16 * s) << 16) + (uint32_t)((rep % nd) * p.N);
        if (p.bf16) {
          // k-step j sits at (j % 4) * step + (j / 4) * step2 when a second-level advance is given
          const uint32_t da2 = p.a_step2 >> 4, db2 = p.b_step2 >> 4;
          for (int j = 0; j < p.ksteps; ++j) {
            const uint32_t oa = da2 ? (j & 3) * da + (j >> 2) * da2 : j * da;
            const uint32_t ob = db2 ? (j & 3) * db + (j >> 2) * db2 : j * db;
            umma::mma_bf16_ss(d, a0 + (uint64_t)oa, b0[s] + (uint64_t)ob, idesc, (rep >= nd || j) ? 1u : 0u);
          }
        } else if (p.ksteps == 8 && !p.a_tmem) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma::mma_tf32_ss(d, a0 + (uint64_t)(j * da), b0[s] + (uint64_t)(j * db), idesc, (rep >= nd || j) ? 1u : 0u);
        } else if (p.a_tmem) {
          for (int j = 0; j < p.ksteps; ++j)
            umma::mma_tf32_ts(d, tb + 256 + 8 * j, b0[s] + (uint64_t)(j * db), idesc, (rep >= nd || j) ? 1u : 0u);
        } else {
          for (int j = 0; j < p.ksteps; ++j)
            umma::mma_tf32_ss(d, a0 + (uint64_t)(j * da), b0[s] + (uint64_t)(j * db), idesc, (rep >= nd || j) ? 1u : 0u);
        }
      }
    umma::commit(&bar);
    const bool ok = umma::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    p.cycles[0] = t1 - t0;
    p.status[0] = ok ? 1 : 0;
    p.status[1] = (int)tb;
  }
  __syncthreads();
  umma::fence_after_sync();
  for (int col = 0; col < p.N; col += 8) {
    float v[8];
    umma::tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + col, v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) p.out[(size_t)(warp * 32 + lane) * 512 + col + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tb, 512);
}

#define CKP(x)                                                                    \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      fprintf(stderr, "umma_probe: %s -> %s\n", #x, cudaGetErrorString(e_));     \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

}  // namespace

// cfg: [0] M, [1] N, [2] ksteps, [3] a_mn_major, [4] b_mn_major, [5] nsets, [6] reps, [7] a_words, [8] b_words,
//      [9..11] A: lbo, sbo, k-step advance (bytes), [12..14] B: lbo, sbo, k-step advance, [15] accumulator regions (timing),
//      [16], [17] swizzle layout type of A, B, [18] A operand from TMEM (A = row-major [128][8*ksteps]),
//      [19] 1 = kind::f16 with bf16 operands (images hold packed bf16 pairs, K = 16 per instruction),
//      [20], [21] second-level k-step advance of A, B in bytes (0 = none)
extern "C" int pinn_umma_probe(int device, const float* A, const float* B0, const float* B1, const int* cfg, float* out,
                               double* cycles, int* status) {
  CKP(cudaSetDevice(device));
  const int M = cfg[0], N = cfg[1], nsets = cfg[5], aw = cfg[7], bw = cfg[8];
  if ((M != 64 && M != 128) || N % 8 || N < 8 || N > 256 || cfg[2] <= 0 || nsets < 1 || nsets > 2 || (nsets == 2 && M != 64)) return 2;
  const size_t smem = (size_t)(aw + 2 * bw) * 4;
  if (smem > 200 * 1024 || aw <= 0 || bw <= 0) return 2;
  ProbeArgs p{};
  float *dA, *dB0, *dB1 = nullptr, *dout;
  long long* dcyc;
  int* dst;
  CKP(cudaMalloc(&dA, (size_t)aw * 4));
  CKP(cudaMalloc(&dB0, (size_t)bw * 4));
  CKP(cudaMalloc(&dB1, (size_t)bw * 4));
  CKP(cudaMalloc(&dout, 128 * 512 * 4));
  CKP(cudaMalloc(&dcyc, 8));
  CKP(cudaMalloc(&dst, 8));
  CKP(cudaMemcpy(dA, A, (size_t)aw * 4, cudaMemcpyHostToDevice));
  CKP(cudaMemcpy(dB0, B0, (size_t)bw * 4, cudaMemcpyHostToDevice));
  if (B1) CKP(cudaMemcpy(dB1, B1, (size_t)bw * 4, cudaMemcpyHostToDevice));
  CKP(cudaMemset(dout, 0xff, 128 * 512 * 4));
  CKP(cudaMemset(dst, 0, 8));
  p.A = dA; p.B0 = dB0; p.B1 = B1 ? dB1 : dB0;
  p.a_words = aw; p.b_words = bw;
  p.M = M; p.N = N; p.ksteps = cfg[2]; p.a_mn = cfg[3]; p.b_mn = cfg[4]; p.nsets = nsets; p.reps = cfg[6] < 1 ? 1 : cfg[6];
  p.nd = cfg[15]; p.a_lt = cfg[16]; p.b_lt = cfg[17]; p.a_tmem = cfg[18]; p.bf16 = cfg[19]; p.a_step2 = cfg[20]; p.b_step2 = cfg[21]; p.a_lbo = cfg[9]; p.a_sbo = cfg[10]; p.a_step = cfg[11]; p.b_lbo = cfg[12]; p.b_sbo = cfg[13]; p.b_step = cfg[14];
  p.out = dout; p.cycles = dcyc; p.status = dst;
  CKP(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem>>>(p);
  CKP(cudaGetLastError());
  CKP(cudaDeviceSynchronize());
  long long cyc = 0;
  CKP(cudaMemcpy(out, dout, 128 * 512 * 4, cudaMemcpyDeviceToHost));
  CKP(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
  CKP(cudaMemcpy(status, dst, 8, cudaMemcpyDeviceToHost));
  *cycles = (double)cyc;
  cudaFree(dA); cudaFree(dB0); cudaFree(dB1); cudaFree(dout); cudaFree(dcyc); cudaFree(dst);
  return 0;
}
