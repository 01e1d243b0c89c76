// One translation unit per tensor-core kernel configuration:
//   nvcc ... -DJ_WP=64 -DJ_N1=2 -DJ_N2=2 -DJ_MIX=0 -c jet_mma_inst.cu -o mma_64_220.o
#include "jet_mma_kernel.cuh"
#include "jet_launch.h"

#ifndef J_WP
#error "compile with -DJ_WP= -DJ_N1= -DJ_N2= -DJ_MIX="
#endif

#ifndef J_NT
#define J_NT (J_WP <= 128 ? 128 : 256)
#endif
using Cfg = MmaCfg<J_WP, J_N1, J_N2, J_MIX, J_NT>;

static cudaError_t launch_impl(const PinnLaunch& L, bool train, int grid, cudaStream_t stream) {
  if (train)
    jet_mma_kernel<Cfg, true><<<grid, Cfg::NT, Cfg::smem_bytes(true), stream>>>(L);
  else
    jet_mma_kernel<Cfg, false><<<grid, Cfg::NT, Cfg::smem_bytes(false), stream>>>(L);
  return cudaGetLastError();
}

static cudaError_t prepare_impl(int* ctas_per_sm) {
  cudaError_t e = cudaFuncSetAttribute(jet_mma_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Cfg::smem_bytes(true));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(jet_mma_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)Cfg::smem_bytes(false));
  if (e != cudaSuccess) return e;
  int n = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, jet_mma_kernel<Cfg, true>, Cfg::NT, Cfg::smem_bytes(true));
  if (Cfg::USE_TMEM && n > 2) n = 2;  // two CTAs x 256 TMEM columns per SM
  if (ctas_per_sm) *ctas_per_sm = n < 1 ? 1 : n;
  return e;
}

#define CAT_(a, b, c, d) pinn_mma_info_##a##_##b##c##d
#define CAT(a, b, c, d) CAT_(a, b, c, d)

extern const JetKernelInfo CAT(J_WP, J_N1, J_N2, J_MIX) = {
    J_WP, J_N1, J_N2, J_MIX, Cfg::K, Cfg::TP, Cfg::smem_bytes(true), Cfg::smem_bytes(false),
    (size_t)Cfg::TP * Cfg::K * Cfg::WP, launch_impl, prepare_impl, /*kind=*/1, /*ldw=*/Cfg::WPS};
