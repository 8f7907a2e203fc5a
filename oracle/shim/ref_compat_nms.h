/* Compatibility shim for compiling the REFERENCE's own NMS CUDA op (mmdet/ops/nms/src/nms_cuda.cpp,
 * nms_kernel.cu), unmodified and where it lies, against torch 2.x.  Force-included by oracle/build.py.  It first
 * pulls in the torch headers the reference file includes (their include guards then make the file's own
 * includes no-ops), and only afterwards maps the removed THC names: THCState / lazyInitCUDA(), THCudaMalloc,
 * THCudaFree (-> the CUDA caching allocator), THCCeilDiv, THCudaCheck, AT_CHECK.  Test infrastructure only,
 * no reference code in here. */
#pragma once
#include <ATen/ATen.h>
#include <ATen/DeviceGuard.h>
#include <ATen/ceil_div.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDACachingAllocator.h>
#include <c10/cuda/CUDAException.h>
#include <torch/extension.h>

#include <iostream>
#include <vector>

struct THCState;
namespace at {
struct HvrRefShimContext {
  THCState* lazyInitCUDA() const {
    at::globalContext().lazyInitDevice(c10::DeviceType::CUDA);
    return nullptr;
  }
};
inline HvrRefShimContext hvrRefShimContext() { return {}; }
}  // namespace at
#define globalContext hvrRefShimContext
#define THCudaMalloc(state, n) c10::cuda::CUDACachingAllocator::raw_alloc(n)
#define THCudaFree(state, p) c10::cuda::CUDACachingAllocator::raw_delete(p)
#define THCCeilDiv(a, b) at::ceil_div((a), (b))
#ifndef THCudaCheck
#define THCudaCheck(x) C10_CUDA_CHECK(x)
#endif
#ifndef AT_CHECK
#define AT_CHECK TORCH_CHECK
#endif
