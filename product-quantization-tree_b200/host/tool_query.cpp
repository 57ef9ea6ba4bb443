// tool_query.cpp -- queries a PQT database.  Same flags and file names as the reference's
// tool (tool_query.cpp:26-36, 77-108); fixes what the shipped tool gets wrong (SURVEY.md
// section 0): it loads <pre>_<lineparts>.lines and hands the codes to the index, and it
// advances the query pointer by the slab offset (the reference multiplies twice,
// tool_query.cpp:153-155).  Extra flags: --k, --queries, --groundtruth, --out.
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "PerturbationProTree.hh"
#include "filereader.hpp"
#include "flags.hpp"

template <typename T>
static std::vector<T> slurp(const std::string& name, size_t count) {
  std::ifstream f(name, std::ios::in | std::ios::binary);
  if (!f.good()) throw std::runtime_error("cannot open " + name);
  std::vector<T> v(count);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(count * sizeof(T)));
  if ((size_t)f.gcount() != count * sizeof(T)) throw std::runtime_error("short read on " + name);
  std::cout << "read " << name << std::endl;
  return v;
}

int main(int argc, char** argv) {
  Flags fl;
  fl.add("device", "0", "selected cuda device");
  fl.add("c1", "4", "number of clusters in first level");
  fl.add("c2", "4", "number of refinements in second level");
  fl.add("p", "2", "parts per vector");
  fl.add("dim", "128", "expected dimension for each vector");
  fl.add("lineparts", "32", "vectorparts for reranking informations");
  fl.add("chunksize", "100000", "number of vectors per chunk (kept for flag compatibility; the DB size comes from the .dbIdx file)");
  fl.add("hashsize", "400000000", "maximal number of bins");
  fl.add("basename", "tmp", "prefix for generated data");
  fl.add("dataset", "base.umem", "path to vector dataset");
  fl.add("queryset", "query.umem", "path to query vectors");
  fl.add("k", "4096", "neighbours returned per query (the reference hard-codes 4096)");
  fl.add("queries", "0", "number of queries (0 = all in the query set)");
  fl.add("groundtruth", "", "optional .imem with exact neighbours: prints recall@1");
  fl.add("out", "", "optional output prefix: writes <out>.idx.imem and <out>.dist.fmem");
  fl.add("compact", "0", "1: load <pre>_<lineparts>.pqtx (tool_createdb --compact 1) instead of .prefix/.count/.dbIdx/.lines");
  try {
    if (!fl.parse(argc, argv,
                  "This tool queries a database built by tool_createdb\n"
                  "Usage:\n    tool_query --c1 4 --c2 4 --p 2 --basename \"tmp\" --dataset base.umem "
                  "--queryset query.umem"))
      return 0;
    const uint32_t dim = (uint32_t)fl.num("dim"), p = (uint32_t)fl.num("p");
    const uint32_t c1 = (uint32_t)fl.num("c1"), c2 = (uint32_t)fl.num("c2");
    const uint32_t LP = (uint32_t)fl.num("lineparts"), k = (uint32_t)fl.num("k");
    const uint32_t hashsize = (uint32_t)fl.num("hashsize");
    const std::string pre = fl.str("basename") + "_" + std::to_string(dim) + "_" + std::to_string(p) +
                            "_" + std::to_string(c1) + "_" + std::to_string(c2);

    FileReader<float> DataReader(fl.str("dataset"));
    FileReader<float> QueryReader(fl.str("queryset"));
    if (QueryReader.dim() != dim) throw std::runtime_error("query dimension differs from --dim");
    uint32_t QN = (uint32_t)fl.num("queries");
    if (QN == 0 || QN > QueryReader.num()) QN = QueryReader.num();
    std::vector<float> query = QueryReader.data(QN);

    pqt::PerturbationProTree ppt(dim, p, p, (int)fl.num("device"));
    ppt.setHashSize(hashsize);
    const std::string codebook_file = pre + ".ppqt";
    struct stat sb;
    if (stat(codebook_file.c_str(), &sb) != 0) {
      std::cout << "you need to generate a codebook first. No codebook found in " << codebook_file
                << std::endl;
      return 1;
    }
    std::cout << "codebook exists, reading from " << codebook_file << std::endl;
    ppt.readTreeFromFile(codebook_file);

    if (fl.num("compact") != 0) {
      // the resident index as one file: no dense .prefix/.count arrays, no re-ordering of the codes
      ppt.loadIndex(pre + "_" + std::to_string(LP) + ".pqtx");
    } else {
      // The database size is what tool_createdb wrote, i.e. the length of .dbIdx (the reference
      // takes DataReader.num(); --chunksize is not the DB size).  Every file must agree with it.
      const std::string idx_file = pre + ".dbIdx", lines_file = pre + "_" + std::to_string(LP) + ".lines";
      if (stat(idx_file.c_str(), &sb) != 0) throw std::runtime_error("cannot stat " + idx_file);
      if (sb.st_size == 0 || sb.st_size % 4) throw std::runtime_error(idx_file + ": size is not a multiple of 4");
      if ((uint64_t)sb.st_size / 4 > 0xFFFFFFFFull) throw std::runtime_error(idx_file + ": too many vectors");
      const uint32_t base_num = (uint32_t)(sb.st_size / 4);
      if (base_num > DataReader.num())
        throw std::runtime_error("the database holds " + std::to_string(base_num) + " vectors, the dataset only " +
                                 std::to_string(DataReader.num()));
      if (stat(lines_file.c_str(), &sb) != 0) throw std::runtime_error("cannot stat " + lines_file);
      if ((uint64_t)sb.st_size != (uint64_t)base_num * LP * 4)
        throw std::runtime_error(lines_file + ": size does not match " + std::to_string(base_num) + " vectors x " +
                                 std::to_string(LP) + " line parts");
      std::vector<pqt::uint> binPrefix = slurp<pqt::uint>(pre + ".prefix", hashsize);
      std::vector<pqt::uint> binCounts = slurp<pqt::uint>(pre + ".count", hashsize);
      std::vector<pqt::uint> dbIdx = slurp<pqt::uint>(pre + ".dbIdx", base_num);
      std::vector<float> hLines = slurp<float>(lines_file, (size_t)base_num * LP);

      ppt.setDB(base_num, binPrefix.data(), binCounts.data(), dbIdx.data());
      ppt.setLines(hLines.data(), base_num, LP);
      binPrefix.clear();
      binPrefix.shrink_to_fit();
      binCounts.clear();
      binCounts.shrink_to_fit();
    }

    std::vector<pqt::uint> allIdx((size_t)QN * k);
    std::vector<float> allDist((size_t)QN * k);
    std::vector<pqt::uint> resIdx;
    std::vector<float> resDist;
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t idxA = 0; idxA < QN; idxA += 4096) {  // 4096-query slabs (tool_query.cpp:153)
      const uint32_t len = std::min<uint32_t>(4096, QN - idxA);
      ppt.queryKNN(resIdx, resDist, query.data() + (size_t)idxA * dim, len, k);
      std::copy(resIdx.begin(), resIdx.end(), allIdx.begin() + (size_t)idxA * k);
      std::copy(resDist.begin(), resDist.end(), allDist.begin() + (size_t)idxA * k);
    }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "queried " << QN << " vectors, k = " << k << ", in " << s << " s  ("
              << (QN / s) << " queries/s)" << std::endl;
    for (uint32_t r = 0; r < std::min<uint32_t>(QN, 5); r++)
      std::cout << "query " << r << ": best " << allIdx[(size_t)r * k] << " (dist "
                << allDist[(size_t)r * k] << ")" << std::endl;

    if (!fl.str("groundtruth").empty()) {
      FileReader<int> gt(fl.str("groundtruth"));
      std::vector<int> g = gt.data(std::min<uint32_t>(QN, gt.entries()), 0);
      const uint32_t gd = gt.dimension(), n = std::min<uint32_t>(QN, gt.entries());
      uint32_t hit = 0;
      for (uint32_t r = 0; r < n; r++) hit += (allIdx[(size_t)r * k] == (pqt::uint)g[(size_t)r * gd]);
      std::cout << "recall@1 = " << (double)hit / n << " over " << n << " queries" << std::endl;
    }
    if (!fl.str("out").empty()) {
      pqt_io::writeMem<pqt::uint>(fl.str("out") + ".idx.imem", allIdx.data(), QN, k);
      pqt_io::writeMem<float>(fl.str("out") + ".dist.fmem", allDist.data(), QN, k);
      std::cout << "written " << fl.str("out") << ".idx.imem / .dist.fmem" << std::endl;
    }
  } catch (const std::exception& e) {
    std::cerr << "tool_query: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
