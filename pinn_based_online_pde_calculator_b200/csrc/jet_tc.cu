// Weight images of the tcgen05 family D: stream of 4 KB chunks in consumption order -- GEMM g = 0..L-2 is
// the forward product of layer g+1, then the data-gradient products of layers L-1..1; inside a GEMM:
// output block of 128 rows, k-step of 16, plane (b2, b1, b0: small products are issued first).  A chunk
// is a K-major image [128 rows m][16 k] with 32-byte swizzle:
//   byte(m, k) = m*32 + (((k/8) ^ ((m>>2)&1)) * 16) + (k%8)*2
// forward rows m = output unit, k = input unit; data gradient rows m = input unit, k = output unit.
#include "jet_tc.h"

#include <stdint.h>

namespace {
__global__ void k_tc_images(const float* __restrict__ wpack, PinnNet net, int ldw, uint16_t* __restrict__ img, size_t copy_elems) {
  const int WP = net.wp, MB = WP / 128, KS = WP / 16, NG = net.n_hidden - 1;
  const size_t per_gemm = (size_t)MB * KS * 3 * 2048;  // bf16 elements
  const int g = blockIdx.y;                            // 0 .. 2*NG-1
  const bool fwd = g < NG;
  const int l = fwd ? g + 1 : (net.n_hidden - 1) - (g - NG);
  const float* W = wpack + net.off_w[l];
  uint16_t* out = img + (size_t)blockIdx.z * copy_elems + (size_t)g * per_gemm;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < MB * KS * 2048; idx += gridDim.x * blockDim.x) {
    const int kk = idx & 15, m = (idx >> 4) & 127, ks = (idx >> 11) % KS, mb = (idx >> 11) / KS;
    const int row = 128 * mb + m, k = 16 * ks + kk;
    float r = fwd ? W[(size_t)k * ldw + row] : W[(size_t)row * ldw + k];
    uint16_t pl[3];  // three-way split, round to nearest at every level (residuals are exact)
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      uint32_t pk;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(0.f), "f"(r));
      pl[p] = (uint16_t)(pk & 0xffffu);
      r -= __uint_as_float((pk & 0xffffu) << 16);
    }
    const int off = m * 16 + (((kk >> 3) ^ ((m >> 2) & 1)) << 3) + (kk & 7);
    uint16_t* c = out + (size_t)((mb * KS + ks) * 3) * 2048;
    c[off] = pl[2];
    c[2048 + off] = pl[1];
    c[4096 + off] = pl[0];
  }
}
}  // namespace

size_t jet_tc_image_bytes(const PinnNet& net) {
  const int MB = net.wp / 128, KS = net.wp / 16, NG = net.n_hidden - 1;
  return (size_t)2 * NG * MB * KS * 3 * 4096;
}

cudaError_t jet_tc_build_images(const float* wpack, const PinnNet& net, int ldw, void* images, int copies, cudaStream_t st) {
  const int NG = net.n_hidden - 1;
  if (NG <= 0) return cudaSuccess;
  k_tc_images<<<dim3(8, 2 * NG, copies), 256, 0, st>>>(wpack, net, ldw, reinterpret_cast<uint16_t*>(images), jet_tc_image_bytes(net) / 2);
  return cudaGetLastError();
}
