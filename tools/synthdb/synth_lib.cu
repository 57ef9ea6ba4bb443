// libpqt_synth.so -- C entry point of the synthetic generator for bench.py / tests (ctypes).
#include "synth_gen.cuh"

extern "C" int pqts_db_u8(void* d_out, uint64_t i0, uint32_t n, uint32_t dim, const void* d_mu,
                          uint32_t n_clusters, uint32_t seed, void* stream) {
  return (int)pqts::db_u8(d_out, i0, n, dim, d_mu, n_clusters, seed, static_cast<cudaStream_t>(stream));
}
