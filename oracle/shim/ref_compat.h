/* Compatibility shim for compiling the REFERENCE's own CUDA op sources (mmdet/ops/roi_align/src/*.cu, *.cpp),
 * unmodified and where they lie under /root/reference, against torch 2.x: names that torch removed since
 * mmdetection v1 are mapped to their successors.  Force-included by oracle/build.py (-include); test
 * infrastructure only, no reference code in here. */
#pragma once
#include <c10/cuda/CUDAException.h>
#include <c10/util/Exception.h>
#ifndef AT_CHECK
#define AT_CHECK TORCH_CHECK
#endif
#ifndef THCudaCheck
#define THCudaCheck(x) C10_CUDA_CHECK(x)
#endif
