// Build shim (TEST INFRASTRUCTURE ONLY, used by oracle/build_ref.py): the reference's
// third_party/PyTorchEMD/cuda/emd_kernel.cu:14 includes <THC/THC.h>, a header PyTorch removed in
// 1.11, for exactly two macros.  Supplying them lets the reference kernel compile UNMODIFIED from
// where it lies, so that it can serve as the ground truth for lion_emd_approx.
#pragma once
#include <cuda_runtime.h>
#include <c10/cuda/CUDAException.h>
#include <c10/util/Exception.h>
#define THCudaCheck(x) C10_CUDA_CHECK(x)
#ifndef CHECK_EQ
#define CHECK_EQ(a, b) TORCH_CHECK((a) == (b), "CHECK_EQ failed: " #a " == " #b)
#endif
