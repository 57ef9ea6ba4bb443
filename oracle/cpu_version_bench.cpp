// cpu_version_bench -- times the reference's CPU twin (cpu_version/, header-only C++11 + Eigen)
// on BASELINE config 1: treequantizer<float,128,16,8,4,4,32> (D=128, C1=16, C2=8, P=4, W=4,
// LP=32 = the repo defaults of cpu_version/tools/query.cpp:10-15 with P raised to 4), following
// the reference's own three tools in one process:
//   tools/build_tree.cpp:27-40   generate() on the learn set
//   tools/build_db.cpp:24-40     notify(N) + insert() of every base vector
//   tools/query.cpp:21-85,128-138 query(20000, 500, q, cand) per query, recall@R, avg ms/query
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/pqt_oracle.h).  The reference's headers are
// included from /root/reference at build time (oracle/Makefile -> oracle/_ref/cpu_version_bench);
// Eigen comes from oracle/eigen_shim (Eigen itself is not in this image).  The object keeps its
// scratch in members (treequantizer.hpp:913-915), so --threads N runs N replicas of the index,
// one per thread, each answering its share of the queries.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "helper.hpp"
#include "timer.hpp"
#include "iterator/memiterator.hpp"
#include "iterator/iterator.hpp"
#include "quantizer/treequantizer.hpp"

const uint D = 128;
const uint P = 4;
const uint C1 = 16;
const uint C2 = 8;
const uint H1 = 4;
const uint RE = 32;
typedef float T;
typedef treequantizer<T, D, C1, C2, P, H1, RE> tree_t;

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s base.umem query.umem groundtruth.imem [ntrain] [nquery] [threads] [boundVectors] [boundBins]\n", argv[0]);
    return 2;
  }
  const std::string base = argv[1], query = argv[2], truth = argv[3];
  uint ntrain = argc > 4 ? (uint)std::atoi(argv[4]) : 300000;
  uint nquery = argc > 5 ? (uint)std::atoi(argv[5]) : 1000;
  uint threads = argc > 6 ? (uint)std::atoi(argv[6]) : 1;
  const uint boundVectors = argc > 7 ? (uint)std::atoi(argv[7]) : 20000;  // tools/query.cpp:42
  const uint boundBins = argc > 8 ? (uint)std::atoi(argv[8]) : 500;
  std::streambuf* cout_buf = std::cout.rdbuf();
  std::cout.rdbuf(std::cerr.rdbuf());  // the reference's headers chat on stdout; keep it for the JSON line

  memiterator<float, uint8_t> base_set;
  base_set.open(base);
  const uint N = base_set.num();
  if (base_set.dim() != D) { std::fprintf(stderr, "base dimension %u != %u\n", base_set.dim(), D); return 2; }
  float* base_data = base_set.all();
  ntrain = std::min(ntrain, N);

  memiterator<float, uint8_t> query_set;
  query_set.open(query);
  nquery = std::min(nquery, query_set.num());
  float* query_data = query_set.all();
  iterator<float, 128> iter_query;
  iter_query.insertBatch(query_data, nquery);

  memiterator<int, int> truth_set;
  truth_set.open(truth);
  const uint gd = truth_set.dim();
  int* truth_data = truth_set.all();

  threads = std::max(1u, threads);
  std::vector<tree_t*> Q(threads);
  double t0 = now_s();
  {
    // learn set = the first ntrain base vectors (test/testPPQT.cpp:285 trains on a 300 k prefix)
    iterator<float, 128> iter_learn;
    iter_learn.insertBatch(base_data, ntrain);
    Q[0] = new tree_t();
    Q[0]->generate(iter_learn);
  }
  const double train_s = now_s() - t0;
  t0 = now_s();
  Q[0]->notify(N);
  for (uint n = 0; n < N; ++n) {
    Eigen::Matrix<T, D, 1> curVec = Eigen::Map<Eigen::Matrix<T, D, 1>>(base_data + (size_t)n * D);
    Q[0]->insert(curVec);
  }
  const double insert_s = now_s() - t0;
  // replicas for the other threads: same codebooks and bins through the reference's own files
  if (threads > 1) {
    const std::string tf = base + ".cpuv.tree", bf = base + ".cpuv.bins";
    Q[0]->saveTree(tf);
    Q[0]->saveBins(bf);
    for (uint t = 1; t < threads; ++t) {
      Q[t] = new tree_t();
      Q[t]->loadTree(tf);
      Q[t]->loadBins(bf);
    }
    std::remove(tf.c_str());
    std::remove(bf.c_str());
  }

  // ---- query (tools/query.cpp:21-85): thread t answers queries t, t + threads, ...
  std::vector<uint> rank_of(nquery, 0xFFFFFFFFu);
  std::vector<double> cand_len(threads, 0.0);
  auto worker = [&](uint t) {
    for (uint i = t; i < nquery; i += threads) {
      const uint correctId = (uint)truth_data[(size_t)i * gd];
      std::vector<std::pair<uint, T>> vectorCandidates;
      Q[t]->query(boundVectors, boundBins, iter_query[i], vectorCandidates);
      cand_len[t] += vectorCandidates.size();
      for (uint s = 0, s_e = (uint)vectorCandidates.size(); s < s_e; ++s)
        if (vectorCandidates[s].first == correctId) {
          rank_of[i] = s;
          break;
        }
    }
  };
  // one untimed pass over a few queries (page faults, caches), then the timed pass
  for (uint i = 0; i < std::min(nquery, 16u); ++i) {
    std::vector<std::pair<uint, T>> tmp;
    Q[0]->query(boundVectors, boundBins, iter_query[i], tmp);
  }
  t0 = now_s();
  if (threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (uint t = 0; t < threads; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
  }
  const double query_s = now_s() - t0;
  uint r1 = 0, r10 = 0, r100 = 0;
  double cl = 0;
  for (uint i = 0; i < nquery; ++i) {
    r1 += rank_of[i] < 1;
    r10 += rank_of[i] < 10;
    r100 += rank_of[i] < 100;
  }
  for (double c : cand_len) cl += c;
  std::cout.rdbuf(cout_buf);
  std::printf("{\"impl\": \"cpu_version treequantizer<float,128,16,8,4,4,32>::query(%u,%u)\", \"n\": %u, "
              "\"queries\": %u, \"threads\": %u, \"queries_per_s\": %.3f, \"ms_per_query\": %.5f, "
              "\"recall_at_1\": %.4f, \"recall_at_10\": %.4f, \"recall_at_100\": %.4f, "
              "\"avg_candidates\": %.1f, \"train_s\": %.2f, \"insert_s\": %.2f, \"ntrain\": %u}\n",
              boundVectors, boundBins, N, nquery, threads, nquery / query_s, 1000.0 * query_s * threads / nquery,
              (double)r1 / nquery, (double)r10 / nquery, (double)r100 / nquery, cl / nquery, train_s,
              insert_s, ntrain);
  return 0;
}
