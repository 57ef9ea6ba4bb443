// tool_createdb.cpp -- builds a PQT database for a .umem dataset.
// Same flags and file names as the reference's tool (tool_createdb.cpp:26-36, 58-84),
// implementing what that tool intends (SURVEY.md sections 0 and 3.3): it writes the
// device-built prefix / count / dbIdx arrays (the reference dumps zero-initialised host
// buffers, tool_createdb.cpp:97-99) and trains on real data (the reference trains on an
// uninitialised buffer, pqt/PerturbationProTree.cu:285-290).
//   <base>_<dim>_<p>_<c1>_<c2>.ppqt / .prefix / .count / .dbIdx / _<lineparts>.lines
#include <sys/stat.h>

#include <chrono>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "PerturbationProTree.hh"
#include "filereader.hpp"
#include "flags.hpp"
#include "kmeans.hpp"

static bool file_exists(const std::string& n) {
  struct stat b;
  return stat(n.c_str(), &b) == 0;
}

template <typename T>
static void dump(const std::string& name, const std::vector<T>& v) {
  std::ofstream f(name, std::ios::out | std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
  if (!f.good()) throw std::runtime_error("write error on " + name);
  std::cout << "written " << name << std::endl;
}

int main(int argc, char** argv) {
  Flags fl;
  fl.add("device", "0", "selected cuda device");
  fl.add("c1", "4", "number of clusters in first level");
  fl.add("c2", "4", "number of refinements in second level");
  fl.add("p", "2", "parts per vector");
  fl.add("dim", "128", "expected dimension for each vector");
  fl.add("lineparts", "32", "vectorparts for reranking informations");
  fl.add("chunksize", "10000000", "number of vectors per chunk");
  fl.add("hashsize", "400000000", "maximal number of bins");
  fl.add("basename", "tmp", "prefix for generated data");
  fl.add("dataset", "base.umem", "path to vector dataset");
  fl.add("train", "20000", "vectors used for codebook training (reference literal)");
  fl.add("compact", "0", "1: also write <pre>_<lineparts>.pqtx, the resident index as one file (tool_query --compact 1 loads it)");
  try {
    if (!fl.parse(argc, argv,
                  "This tool builds a database for a given dataset of vectors\n"
                  "Usage:\n    tool_createdb --c1 4 --c2 4 --p 2 --basename \"tmp\" --dataset base.umem"))
      return 0;
    const uint32_t dim = (uint32_t)fl.num("dim"), p = (uint32_t)fl.num("p");
    const uint32_t c1 = (uint32_t)fl.num("c1"), c2 = (uint32_t)fl.num("c2");
    const uint32_t LP = (uint32_t)fl.num("lineparts");
    const std::string pre = fl.str("basename") + "_" + std::to_string(dim) + "_" + std::to_string(p) +
                            "_" + std::to_string(c1) + "_" + std::to_string(c2);

    FileReader<float> reader(fl.str("dataset"));
    if (reader.dim() != dim) throw std::runtime_error("dataset dimension differs from --dim");
    // The whole dataset goes in, chunk by chunk (the reference tool stops after the first
    // chunk; its 1-B driver test/test1B.cpp:783-871 loops over the chunks like this)
    const uint32_t N = reader.num();
    const uint32_t chunk = (uint32_t)std::max<long long>(1, std::min<long long>(fl.num("chunksize"), N));
    std::cout << N << " x " << dim << " vectors in " << fl.str("dataset") << ", chunks of " << chunk << std::endl;

    pqt::PerturbationProTree ppt(dim, p, p, (int)fl.num("device"));
    ppt.setHashSize((uint32_t)fl.num("hashsize"));
    const std::string codebook_file = pre + ".ppqt";
    if (file_exists(codebook_file)) {
      std::cout << "codebook exists, reading from " << codebook_file << std::endl;
      ppt.readTreeFromFile(codebook_file);
    } else {
      const size_t ntrain = std::min<size_t>((size_t)fl.num("train"), N);
      std::vector<float> cb1, cb2, data = reader.data(ntrain);
      pqt_train::train_tree(data.data(), ntrain, dim, p, c1, c2, cb1, cb2);
      ppt.setTree(c1, c2, cb1.data(), cb2.data());
      ppt.writeTreeToFile(codebook_file);
      std::cout << "written " << codebook_file << std::endl;
    }

    auto t0 = std::chrono::steady_clock::now();
    // pass 1: bins of every chunk, then the inverted lists over all of them
    std::vector<pqt::uint> binOf(N);
    for (uint32_t i0 = 0; i0 < N; i0 += chunk) {
      const uint32_t n = std::min(chunk, N - i0);
      std::vector<uint8_t> rows = pqt_io::readPayload<uint8_t>(fl.str("dataset"), dim, n, i0);
      ppt.assignBins(rows.data(), true, false, n, binOf.data() + i0, false);
    }
    ppt.setDBFromBins(binOf.data(), false, N);
    binOf.clear();
    binOf.shrink_to_fit();
    // pass 2: line codes chunk by chunk, appended to the .lines file in id order
    const std::string lines_name = pre + "_" + std::to_string(LP) + ".lines";
    {
      std::ofstream lf(lines_name, std::ios::out | std::ios::binary);
      std::vector<float> lines((size_t)chunk * LP);
      ppt.lineDistBegin(N, LP);
      for (uint32_t i0 = 0; i0 < N; i0 += chunk) {
        const uint32_t n = std::min(chunk, N - i0);
        std::vector<uint8_t> rows = pqt_io::readPayload<uint8_t>(fl.str("dataset"), dim, n, i0);
        ppt.lineDistChunk(rows.data(), true, false, i0, n, lines.data());
        lf.write(reinterpret_cast<const char*>(lines.data()), (std::streamsize)((size_t)n * LP * 4));
      }
      ppt.lineDistEnd();
      if (!lf.good()) throw std::runtime_error("write error on " + lines_name);
    }
    std::cout << "written " << lines_name << std::endl;
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "built DB of " << N << " vectors in " << s << " s" << std::endl;

    dump(pre + ".prefix", ppt.getBinPrefix());
    dump(pre + ".count", ppt.getBinCounts());
    dump(pre + ".dbIdx", ppt.getDBIdx());
    if (fl.num("compact") != 0) {
      const std::string cname = pre + "_" + std::to_string(LP) + ".pqtx";
      ppt.saveIndex(cname);
      std::cout << "written " << cname << std::endl;
    }
  } catch (const std::exception& e) {
    std::cerr << "tool_createdb: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
