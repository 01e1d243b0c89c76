// tcgen05 (UMMA) kernel family, experimental (PINN_B200_KERNEL=umma): padded width 64, jets
// (value, d/dx0, d/dx1, combined second order).  See DESIGN.md section 4.1 (round-2 plan) and
// umma_common.cuh for the measured descriptor conventions.
#pragma once
#include <cuda_runtime.h>

#include "pinn_common.h"

// true when the network / jet configuration can run on the UMMA family
bool jet_umma_supported(const PinnNet& net, int k, int n1, int n2, int mix);
// floats of the pre-split operand images for `net` (4 images per hidden GEMM layer)
size_t jet_umma_image_floats(const PinnNet& net);
// build the images (K-major B operands, hi/lo tf32 split) from the weight pack
cudaError_t jet_umma_build_images(const float* wpack, const PinnNet& net, int ldw, float* images, cudaStream_t st);
// evaluation (u, f, jets) of the points described by L
cudaError_t jet_umma_eval_launch(const PinnLaunch& L, const float* images, int grid_max, cudaStream_t st, long long* clk);
// loss + gradient of the collocation term (same accumulator / stash conventions as the other families:
// L.gacc[grid][pg], L.stash[grid][n_hidden][4*64*32], L.loss_part[grid][n_slots]); tiles of 32 points
cudaError_t jet_umma_train_launch(const PinnLaunch& L, const float* images, int ldw, int grid, cudaStream_t st, long long* clk);
