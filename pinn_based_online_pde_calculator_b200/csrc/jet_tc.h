// tcgen05 kernel family D (kind 3), shared host helpers: the pre-split bf16 weight images the kernels
// stream through their cp.async.bulk ring.  See jet_tc_kernel.cuh.
#pragma once
#include <cuda_runtime.h>

#include "pinn_common.h"

#define PINN_TC_IMAGE_COPIES 2  // replicas of the stream (each CTA reads replica blockIdx % copies)
// bytes of ONE replica of the image stream for `net` (forward + data-gradient image per hidden GEMM layer)
size_t jet_tc_image_bytes(const PinnNet& net);
// build the stream from the fp32 weight pack (row stride ldw): one launch per evaluation
cudaError_t jet_tc_build_images(const float* wpack, const PinnNet& net, int ldw, void* images, int copies, cudaStream_t st);
