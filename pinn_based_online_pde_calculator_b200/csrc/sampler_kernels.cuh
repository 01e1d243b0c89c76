// Device-side collocation samplers: Latin-hypercube points
// (pyDOE.lhs, software.py:553,562) and inverse-CDF sampling from a cell distribution
// (colloc2D_set, software.py:87-136).  Counter-based RNG (Philox4x32-10) so that every point is
// generated independently: no sort, no scan on the device (the 110x110-cell cumulative sum is
// done on the host and uploaded).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

// keyed bijection on [0, n): 4-round Feistel network on the enclosing power-of-4 domain + cycle walking
__device__ __forceinline__ uint32_t feistel_perm(uint32_t i, uint32_t n, uint32_t half_bits, uint32_t key) {
  const uint32_t mask = (1u << half_bits) - 1u;
  uint32_t x = i;
  do {
    uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      uint32_t f = (r + key * (2u * round + 1u)) * 0x9E3779B1u;
      f ^= f >> 15; f *= 0x85EBCA77u; f ^= f >> 13;
      const uint32_t nl = r, nr = (l ^ f) & mask;
      l = nl; r = nr;
    }
    x = (l << half_bits) | r;
  } while (x >= n);
  return x;
}

// out[i][j] = lo_j + (perm_j(i) + U) / n * (hi_j - lo_j)
__global__ void k_sample_lhs(float* __restrict__ out, long long n, int d, int ld, int col0, float3 lo, float3 hi,
                             uint32_t seed, uint32_t half_bits) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0x4c485331u, 0u, seed, 0x5eed0001u, r);
  const float los[3] = {lo.x, lo.y, lo.z}, his[3] = {hi.x, hi.y, hi.z};
  for (int j = 0; j < d; ++j) {
    const uint32_t s = feistel_perm((uint32_t)i, (uint32_t)n, half_bits, seed * 0x9E3779B9u + 0x1234567u * (j + 1));
    const float t = ((float)s + u01(r[j])) / (float)n;
    out[i * ld + col0 + j] = los[j] + t * (his[j] - los[j]);
  }
}

// inverse-CDF sampling: cum[0..ncell] (cum[0]=0) over the (ny-1)x(nx-1) cells, row-major;
// point = lower-left corner of the drawn cell + uniform fraction of the cell size
__global__ void k_sample_cdf2d(float* __restrict__ out, long long n, int ld, const double* __restrict__ cum,
                               int ncell, int ncx, float x0, float y0, float dx, float dy, uint32_t seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0x43444632u, 0u, seed, 0x5eed0002u, r);
  const double c = ((double)r[0] * 4294967296.0 + (double)r[1] + 0.5) * (1.0 / 18446744073709551616.0) * cum[ncell];
  int lo = 0, hi = ncell;  // largest lo with cum[lo] <= c  (software.py:117-119: floor(interp(c, b, seq)))
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cum[mid] <= c) lo = mid; else hi = mid;
  }
  const int iy = lo / ncx, ix = lo % ncx;
  out[i * ld + 0] = x0 + (ix + u01(r[2])) * dx;
  out[i * ld + 1] = y0 + (iy + u01(r[3])) * dy;
}
