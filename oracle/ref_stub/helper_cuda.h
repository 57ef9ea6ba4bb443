// helper_cuda.h -- one-macro stand-in for the CUDA-samples header the reference
// includes (pqt/helper.hh:9), plus token-level spellings of three pre-Volta
// intrinsics the reference still uses, so that its .cu files compile UNMODIFIED for
// sm_100a.  TEST INFRASTRUCTURE (oracle/_ref cross-check).
#ifndef PQT_REF_STUB_HELPER_CUDA_H
#define PQT_REF_STUB_HELPER_CUDA_H
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
// the CUDA-samples header pulls these in (via helper_string.h); the reference relies on it
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
// everything below that mentions the legacy intrinsics by name must be parsed BEFORE the
// macros at the end of this file (include guards keep later re-includes inert)
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>
#include <thrust/sort.h>

#define checkCudaErrors(val)                                                            \
  do {                                                                                  \
    cudaError_t e__ = (val);                                                            \
    if (e__ != cudaSuccess) {                                                           \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

// pqt/PerturbationProTree.cu:6013, pqt/ProTree.cu:2665 ("syncthreads()" typo)
#define syncthreads __syncthreads
// pqt/PerturbationProTree.cu:3955,4257,4259
#define __any(x) __any_sync(0xffffffffu, (x))
// pqt/PerturbationProTree.cu:5185 (called inside a divergent loop: active lanes only)
#define __shfl_down(v, s) __shfl_down_sync(__activemask(), (v), (s))
#endif
