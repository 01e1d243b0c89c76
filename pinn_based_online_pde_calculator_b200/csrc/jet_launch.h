// Launch interface between the engine and the per-configuration instantiations
// of the fused jet-MLP kernel (jet_inst.cu is compiled once per (WP, N1, N2, MIX)).
#pragma once
#include <cuda_runtime.h>
#include "pinn_common.h"

struct JetKernelInfo {
  int wp, n1, n2, mix;
  int k;             // channels
  int tile_points;   // points per CTA tile
  size_t smem_train, smem_eval;
  size_t stash_floats_per_layer;  // per CTA
  cudaError_t (*launch)(const PinnLaunch& L, bool train, int grid, cudaStream_t stream);
  cudaError_t (*prepare)(int* ctas_per_sm);  // opt in to large dynamic smem; resident CTAs per SM (train kernel)
  int kind;          // 0: fp32 SIMT (FFMA2) kernel, 1: 3xTF32 mma.sync tensor-core kernel, 3: tcgen05 bf16x3 kernel (family D)
  int ldw;           // row stride of the hidden-layer matrices in the weight/gradient pack
};

// generated list (jet_registry.cu)
const JetKernelInfo* pinn_find_kernel(int wp, int n1, int n2, int mix, int kind);
int pinn_kernel_count();
const JetKernelInfo* pinn_kernel_at(int i);
