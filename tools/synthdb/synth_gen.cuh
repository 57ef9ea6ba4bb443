// synth_gen.cuh -- the synthetic SIFT-shaped generator of pqt_b200/synth.py as a CUDA kernel
// (bench / test plumbing, not part of the product library).  Integer arithmetic only, so any
// chunk of the database is regenerated bit-identically on any GPU:
//   h(x)      = murmur3 fmix32
//   cluster g = h(seed ^ h(i)) mod G
//   x_i[d]    = clip(mu_g[d] + ((b0+b1+b2+b3 - 510) * 83 >> 10), 0, 255), b = bytes of
//               h(h(seed + i) + d * 0x85EBCA77)
// mu (the cluster centres, uint8 [G][dim]) is a table computed by the caller.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pqts {

__host__ __device__ inline uint32_t fmix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

__device__ __forceinline__ int noise(uint32_t base, uint32_t d, int shift) {
  const uint32_t hh = fmix32(base + d * 0x85EBCA77u);
  const int s = (int)((hh & 0xFFu) + ((hh >> 8) & 0xFFu) + ((hh >> 16) & 0xFFu) + (hh >> 24));
  return ((s - 510) * 83) >> shift;
}

// one thread per 4 consecutive dimensions of one vector (dim % 4 == 0): 4-byte stores
__global__ void db_u8_kernel(uint8_t* out, uint64_t i0, uint32_t n, uint32_t dim, const uint8_t* mu,
                             uint32_t n_clusters, uint32_t seed) {
  const uint32_t q = dim >> 2;
  const size_t total = (size_t)n * q;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const uint32_t row = (uint32_t)(e / q), d4 = (uint32_t)(e - (size_t)row * q);
    const uint32_t id = (uint32_t)(i0 + row);
    const uint32_t g = fmix32(seed ^ fmix32(id)) % n_clusters;
    const uint32_t base = fmix32(seed + id);
    const uint32_t m = *reinterpret_cast<const uint32_t*>(mu + (size_t)g * dim + d4 * 4);
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int x = (int)((m >> (8 * j)) & 0xFFu) + noise(base, d4 * 4 + j, 10);
      x = x < 0 ? 0 : (x > 255 ? 255 : x);
      w |= (uint32_t)x << (8 * j);
    }
    *reinterpret_cast<uint32_t*>(out + (size_t)row * dim + d4 * 4) = w;
  }
}

inline cudaError_t db_u8(void* d_out, uint64_t i0, uint32_t n, uint32_t dim, const void* d_mu,
                         uint32_t n_clusters, uint32_t seed, cudaStream_t st) {
  if (dim % 4) return cudaErrorInvalidValue;
  db_u8_kernel<<<148 * 16, 256, 0, st>>>(static_cast<uint8_t*>(d_out), i0, n, dim,
                                          static_cast<const uint8_t*>(d_mu), n_clusters, seed);
  return cudaGetLastError();
}

}  // namespace pqts
