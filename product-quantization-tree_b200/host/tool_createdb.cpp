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
    const uint32_t N = (uint32_t)std::min<uint64_t>((uint64_t)fl.num("chunksize"), reader.num());
    std::vector<float> data = reader.data(N);
    std::cout << "read " << N << " x " << dim << " vectors from " << fl.str("dataset") << std::endl;

    pqt::PerturbationProTree ppt(dim, p, p, (int)fl.num("device"));
    ppt.setHashSize((uint32_t)fl.num("hashsize"));
    const std::string codebook_file = pre + ".ppqt";
    if (file_exists(codebook_file)) {
      std::cout << "codebook exists, reading from " << codebook_file << std::endl;
      ppt.readTreeFromFile(codebook_file);
    } else {
      const size_t ntrain = std::min<size_t>((size_t)fl.num("train"), N);
      std::vector<float> cb1, cb2;
      pqt_train::train_tree(data.data(), ntrain, dim, p, c1, c2, cb1, cb2);
      ppt.setTree(c1, c2, cb1.data(), cb2.data());
      ppt.writeTreeToFile(codebook_file);
      std::cout << "written " << codebook_file << std::endl;
    }

    auto t0 = std::chrono::steady_clock::now();
    ppt.buildKBestDB(data.data(), N);
    ppt.lineDist(data.data(), N, LP);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "built DB of " << N << " vectors in " << s << " s" << std::endl;

    dump(pre + "_" + std::to_string(LP) + ".lines", ppt.getLine());
    dump(pre + ".prefix", ppt.getBinPrefix());
    dump(pre + ".count", ppt.getBinCounts());
    dump(pre + ".dbIdx", ppt.getDBIdx());
  } catch (const std::exception& e) {
    std::cerr << "tool_createdb: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
