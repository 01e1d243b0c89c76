// One translation unit per tcgen05 (family D) kernel configuration:
//   nvcc ... -DJ_WP=128 -DJ_N1=2 -DJ_N2=0 -DJ_MIX=2 -c jet_tc_inst.cu -o tc_128_202.o
#include "jet_tc_kernel.cuh"
#include "jet_launch.h"

#ifndef J_WP
#error "compile with -DJ_WP= -DJ_N1= -DJ_N2= -DJ_MIX="
#endif

using Cfg = TcCfg<J_WP, J_N1, J_N2, J_MIX>;

// shared gradient-accumulator rows (gacc_share > 1): the flush-order tokens start at zero, and the members of a row wait
// for each other, so the launch is cooperative (the grid starts only when all of its CTAs can be resident)
template <class Kern>
static cudaError_t launch_shared_rows(Kern kern, const PinnLaunch& L, int grid, cudaStream_t stream) {
  const int rows = (grid + L.gacc_share - 1) / L.gacc_share;
  cudaError_t e = cudaMemsetAsync(L.gacc_token, 0, sizeof(unsigned) * (size_t)rows * PINN_TOKENS, stream);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::NT);
  cfg.dynamicSmemBytes = Cfg::smem_bytes();
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, L);
}

static cudaError_t launch_impl(const PinnLaunch& L, bool train, int grid, cudaStream_t stream) {
  if (train && L.gacc_share > 1 && L.gacc_token)
    return L.phase_clk ? launch_shared_rows(jet_tc_kernel<Cfg, true, true>, L, grid, stream)
                       : launch_shared_rows(jet_tc_kernel<Cfg, true, false>, L, grid, stream);
  if (train && L.phase_clk)
    jet_tc_kernel<Cfg, true, true><<<grid, Cfg::NT, Cfg::smem_bytes(), stream>>>(L);   // phase-clock instantiation
  else if (train)
    jet_tc_kernel<Cfg, true><<<grid, Cfg::NT, Cfg::smem_bytes(), stream>>>(L);
  else
    jet_tc_kernel<Cfg, false><<<grid, Cfg::NT, Cfg::smem_bytes(), stream>>>(L);
  return cudaGetLastError();
}

static cudaError_t prepare_impl(int* ctas_per_sm) {
  cudaError_t e = cudaFuncSetAttribute(jet_tc_kernel<Cfg, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(jet_tc_kernel<Cfg, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(jet_tc_kernel<Cfg, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes());
  if (ctas_per_sm) *ctas_per_sm = 1;  // one CTA per SM: all 512 TMEM columns, ~225 KB of shared memory
  return e;
}

#define CAT_(a, b, c, d) pinn_tc_info_##a##_##b##c##d
#define CAT(a, b, c, d) CAT_(a, b, c, d)

extern const JetKernelInfo CAT(J_WP, J_N1, J_N2, J_MIX) = {
    J_WP, J_N1, J_N2, J_MIX, Cfg::K, Cfg::NP, Cfg::smem_bytes(), Cfg::smem_bytes(),
    Cfg::STL, launch_impl, prepare_impl, /*kind=*/3, /*ldw=*/J_WP + 8};
