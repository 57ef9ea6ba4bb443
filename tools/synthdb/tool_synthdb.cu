// tool_synthdb -- builds the index files of a SYNTHETIC database without ever holding the raw
// vectors: chunks of the counter-based generator (synth_gen.cuh) are produced on the device and
// handed to the chunked builder of libpqt_b200.so, exactly the way test/test1B.cpp:783-871 walks
// SIFT1B in 10-M-vector chunks.  Writes the files tool_createdb writes
// (tool_createdb.cpp:81-84,116-138):
//   <pre>.prefix  <pre>.count  <pre>.dbIdx  <pre>_<lineparts>.lines      pre = <base>_<dim>_<p>_<c1>_<c2>
// The codebook <pre>.ppqt must exist (training is offline).  Bench plumbing: bench.py runs it to
// give the CPU arms and the loader path (pqt_set_db / pqt_set_lines) a 1-B index to read.
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../product-quantization-tree_b200/host/PerturbationProTree.hh"
#include "../../product-quantization-tree_b200/host/flags.hpp"
#include "synth_gen.cuh"

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

static void dump_raw(const std::string& name, const void* p, size_t bytes) {
  FILE* f = fopen(name.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot open " + name + " for writing");
  const char* c = static_cast<const char*>(p);
  for (size_t off = 0; off < bytes;) {
    size_t n = std::min<size_t>(bytes - off, (size_t)1 << 30);
    if (fwrite(c + off, 1, n, f) != n) {
      fclose(f);
      throw std::runtime_error("write error on " + name + " (out of space?)");
    }
    off += n;
  }
  fclose(f);
  std::cout << "written " << name << std::endl;
}

int main(int argc, char** argv) {
  Flags fl;
  fl.add("device", "0", "selected cuda device");
  fl.add("c1", "32", "number of clusters in first level");
  fl.add("c2", "32", "number of refinements in second level");
  fl.add("p", "4", "parts per vector");
  fl.add("dim", "128", "dimension of each vector");
  fl.add("lineparts", "32", "vectorparts for reranking informations");
  fl.add("chunksize", "10000000", "number of vectors per chunk (test/test1B.cpp:623)");
  fl.add("hashsize", "400000000", "maximal number of bins");
  fl.add("basename", "tmp", "prefix for generated data");
  fl.add("n", "1000000", "number of synthetic database vectors");
  fl.add("clusters", "4096", "cluster centres of the synthetic data");
  fl.add("seed", "20160627", "generator seed");
  fl.add("mu", "", "file with the uint8 [clusters][dim] centre table");
  fl.add("nolines", "0", "1: skip the .lines file (bins only)");
  try {
    if (!fl.parse(argc, argv, "Builds the index files of a synthetic database on the GPU, chunk by chunk"))
      return 0;
    const uint32_t dim = (uint32_t)fl.num("dim"), p = (uint32_t)fl.num("p");
    const uint32_t c1 = (uint32_t)fl.num("c1"), c2 = (uint32_t)fl.num("c2");
    const uint32_t LP = (uint32_t)fl.num("lineparts");
    const uint64_t N64 = (uint64_t)fl.num("n");
    if (N64 == 0 || N64 > 0xFFFFFFFFull) throw std::runtime_error("--n out of range");
    const uint32_t N = (uint32_t)N64;
    const uint32_t chunk = (uint32_t)std::min<uint64_t>((uint64_t)fl.num("chunksize"), N);
    const uint32_t ncl = (uint32_t)fl.num("clusters"), seed = (uint32_t)fl.num("seed");
    const std::string pre = fl.str("basename") + "_" + std::to_string(dim) + "_" + std::to_string(p) +
                            "_" + std::to_string(c1) + "_" + std::to_string(c2);
    auto t0 = std::chrono::steady_clock::now();
    auto secs = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };

    pqt::PerturbationProTree ppt(dim, p, p, (int)fl.num("device"));
    ppt.setHashSize((uint32_t)fl.num("hashsize"));
    ppt.readTreeFromFile(pre + ".ppqt");

    // centre table -> device
    std::vector<uint8_t> mu((size_t)ncl * dim);
    {
      std::ifstream f(fl.str("mu"), std::ios::in | std::ios::binary);
      if (!f.good()) throw std::runtime_error("cannot open --mu " + fl.str("mu"));
      f.read(reinterpret_cast<char*>(mu.data()), (std::streamsize)mu.size());
      if ((size_t)f.gcount() != mu.size()) throw std::runtime_error("short read on " + fl.str("mu"));
    }
    CK(cudaSetDevice((int)fl.num("device")));
    uint8_t *d_mu = nullptr, *d_x = nullptr;
    uint32_t* d_bin = nullptr;
    CK(cudaMalloc(&d_mu, mu.size()));
    CK(cudaMemcpy(d_mu, mu.data(), mu.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_x, (size_t)chunk * dim));
    CK(cudaMalloc(&d_bin, (size_t)N * 4));

    // pass 1: bins
    for (uint32_t i0 = 0; i0 < N; i0 += chunk) {
      const uint32_t n = std::min(chunk, N - i0);
      CK(pqts::db_u8(d_x, i0, n, dim, d_mu, ncl, seed, 0));
      CK(cudaDeviceSynchronize());
      ppt.assignBins(d_x, true, true, n, d_bin + i0, true);
    }
    ppt.setDBFromBins(d_bin, true, N);
    CK(cudaFree(d_bin));
    std::cout << "bins + inverted lists of " << N << " vectors after " << secs() << " s" << std::endl;
    {
      std::vector<pqt::uint> v = ppt.getBinPrefix();
      dump_raw(pre + ".prefix", v.data(), v.size() * 4);
      v = ppt.getBinCounts();
      dump_raw(pre + ".count", v.data(), v.size() * 4);
      v = ppt.getDBIdx();
      dump_raw(pre + ".dbIdx", v.data(), v.size() * 4);
    }
    // pass 2: line codes, appended to the .lines file in id order
    if (!fl.num("nolines")) {
      const std::string name = pre + "_" + std::to_string(LP) + ".lines";
      FILE* f = fopen(name.c_str(), "wb");
      if (!f) throw std::runtime_error("cannot open " + name + " for writing");
      float* lines = nullptr;
      CK(cudaMallocHost(&lines, (size_t)chunk * LP * 4));
      ppt.lineDistBegin(N, LP);
      for (uint32_t i0 = 0; i0 < N; i0 += chunk) {
        const uint32_t n = std::min(chunk, N - i0);
        CK(pqts::db_u8(d_x, i0, n, dim, d_mu, ncl, seed, 0));
        CK(cudaDeviceSynchronize());
        ppt.lineDistChunk(d_x, true, true, i0, n, lines);
        if (fwrite(lines, 4, (size_t)n * LP, f) != (size_t)n * LP) {
          fclose(f);
          throw std::runtime_error("write error on " + name + " (out of space?)");
        }
      }
      ppt.lineDistEnd();
      fclose(f);
      cudaFreeHost(lines);
      std::cout << "written " << name << std::endl;
    }
    cudaFree(d_x);
    cudaFree(d_mu);
    std::cout << "built synthetic DB of " << N << " vectors in " << secs() << " s" << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "tool_synthdb: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
