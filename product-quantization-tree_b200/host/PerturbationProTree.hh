// PerturbationProTree.hh -- C++ host mirror of pqt::PerturbationProTree
// (reference: pqt/PerturbationProTree.hh:28-235) over the C ABI of
// include/pqt_b200.h.  Same method names and argument meaning for the calls the
// reference's tools make (tool_query.cpp:92-155, tool_createdb.cpp:74-114), so a
// tool written against the reference class compiles against this header after
// swapping the include.  Differences, all deliberate:
//   * errors throw pqt::Error instead of exit()ing (utils/helper.hpp:7-15);
//   * queryKNN accepts a HOST pointer too (q_on_device = false);
//   * the line codes are loaded explicitly with setLines() -- the shipped
//     tool_query forgets them;
//   * getBinPrefix()/getBinCounts()/getDBIdx()/getLine() return host copies.
#ifndef PQT_B200_HOST_PERTURBATIONPROTREE_HH
#define PQT_B200_HOST_PERTURBATIONPROTREE_HH

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pqt_b200.h"

#ifndef HASH_SIZE
#define HASH_SIZE 400000000  // pqt/PerturbationProTree.hh:12
#endif

namespace pqt {

typedef unsigned int uint;

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// pqt/PerturbationProTree.hh:21-25
typedef struct {
  char p1;
  char p2;
  unsigned short lambda;
} lineDescr;

class PerturbationProTree {
 public:
  PerturbationProTree(uint _dim, uint _p, uint _p2, int _device = 0) : h_(nullptr) {
    int rc = pqt_create(_dim, _p, _p2, _device, &h_);
    if (rc != PQT_OK)
      throw Error(rc, "pqt_create failed: no CUDA device (there is no CPU fallback) or p2 != p");
  }
  ~PerturbationProTree() { pqt_destroy(h_); }
  PerturbationProTree(const PerturbationProTree&) = delete;
  PerturbationProTree& operator=(const PerturbationProTree&) = delete;

  void writeTreeToFile(const std::string& _name) { chk(pqt_write_tree(h_, _name.c_str())); }
  void readTreeFromFile(const std::string& _name) { chk(pqt_read_tree(h_, _name.c_str())); }

  // what createTree leaves behind: the two codebooks (training itself is offline)
  void setTree(uint _c1, uint _c2, const float* _cb1, const float* _cb2) {
    chk(pqt_set_tree(h_, _c1, _c2, _cb1, _cb2));
  }

  // tool flag --hashsize (tool_query.cpp:33); the reference compiles HASH_SIZE in
  void setHashSize(uint _hashSize) {
    pqt_params p;
    chk(pqt_get_params(h_, &p));
    p.hash_size = _hashSize;
    chk(pqt_set_params(h_, &p));
  }
  pqt_params params() const {
    pqt_params p;
    chk(pqt_get_params(h_, &p));
    return p;
  }
  void setParams(const pqt_params& p) { chk(pqt_set_params(h_, &p)); }

  /** upload a previously stored db. pointers should be host pointers */
  void setDB(uint _N, const uint* _prefix, const uint* _counts, const uint* _dbIdx) {
    chk(pqt_set_db(h_, _N, _prefix, _counts, _dbIdx));
  }
  /** host lineDescr[N][LP] as written to <pre>_<LP>.lines (tool_createdb.cpp:116-118) */
  void setLines(const float* _hLines, uint _N, uint _lineParts) {
    chk(pqt_set_lines(h_, reinterpret_cast<const uint32_t*>(_hLines), _N, _lineParts));
  }

  /** as buildDB but tries k1 best clusters on the first level; _A is a host or device pointer */
  void buildKBestDB(const float* _A, uint _N, bool _onDevice = false) {
    chk(pqt_build_kbest_db(h_, _A, _onDevice ? 1 : 0, _N));
  }
  void lineDist(const float* _DB, uint _N, uint _lineParts = 16, bool _onDevice = false) {
    chk(pqt_line_dist(h_, _DB, _onDevice ? 1 : 0, _N, _lineParts));
  }

  // ---- chunked build: the reference's 1-B driver walks the base set in 10-M-vector chunks
  // (test/test1B.cpp:783-871); rows are float or the uint8 payload of a .umem file
  void assignBins(const void* _X, bool _isU8, bool _onDevice, uint _n, uint* _binOut,
                  bool _outOnDevice) {
    chk(pqt_assign_bins(h_, _X, _isU8 ? PQT_X_U8 : PQT_X_F32, _onDevice ? 1 : 0, _n, _binOut,
                        _outOnDevice ? 1 : 0));
  }
  void setDBFromBins(const uint* _binOf, bool _onDevice, uint _N) {
    chk(pqt_set_db_from_bins(h_, _binOf, _onDevice ? 1 : 0, _N));
  }
  void lineDistBegin(uint _N, uint _lineParts) { chk(pqt_line_dist_begin(h_, _N, _lineParts)); }
  /** encodes vectors _id0 .. _id0+_n-1; _linesOut (host, may be null): the chunk's lineDescr rows */
  void lineDistChunk(const void* _X, bool _isU8, bool _onDevice, uint _id0, uint _n,
                     float* _linesOut = nullptr) {
    chk(pqt_line_dist_chunk(h_, _X, _isU8 ? PQT_X_U8 : PQT_X_F32, _onDevice ? 1 : 0, _id0, _n,
                            reinterpret_cast<uint32_t*>(_linesOut)));
  }
  void lineDistEnd() { chk(pqt_line_dist_end(h_)); }

  void queryKNN(std::vector<uint>& _resIdx, std::vector<float>& _resDist, const float* _Q,
                uint _QN, uint _nVec, bool _qOnDevice = false) {
    _resIdx.resize((size_t)_QN * _nVec);
    _resDist.resize((size_t)_QN * _nVec);
    chk(pqt_query_knn(h_, _Q, _qOnDevice ? 1 : 0, _QN, _nVec, _resIdx.data(), _resDist.data(), 0));
  }

  /** the 1-B variant (pqt/PerturbationProTree.hh:80); the codes are the resident ones */
  void queryBIGKNNRerank2(std::vector<uint>& _resIdx, std::vector<float>& _resDist, const float* _Q,
                          uint _QN, uint _nVec, bool _qOnDevice = false) {
    _resIdx.resize((size_t)_QN * _nVec);
    _resDist.resize((size_t)_QN * _nVec);
    chk(pqt_query_big_knn_rerank2(h_, _Q, _qOnDevice ? 1 : 0, _QN, _nVec, _resIdx.data(),
                                  _resDist.data(), 0));
  }

  uint getNPerturbations() const { return 1; }

  std::vector<uint> getBinPrefix() { return getDense(0); }
  std::vector<uint> getBinCounts() { return getDense(1); }
  std::vector<uint> getDBIdx() {
    uint32_t n = 0, lp = 0;
    chk(pqt_get_db_size(h_, &n, &lp));
    std::vector<uint> v(n);
    chk(pqt_get_db(h_, nullptr, nullptr, v.data()));
    return v;
  }
  std::vector<float> getLine() {
    uint32_t n = 0, lp = 0;
    chk(pqt_get_db_size(h_, &n, &lp));
    std::vector<float> v((size_t)n * lp);
    chk(pqt_get_lines(h_, reinterpret_cast<uint32_t*>(v.data())));
    return v;
  }

  // ---- compact index file (no counterpart in the reference): the resident directory, dbIdx and
  // bin-ordered line codes as they are; loading skips the dense .prefix/.count arrays and the
  // re-ordering of the codes
  void saveIndex(const std::string& _name) { chk(pqt_save_index(h_, _name.c_str())); }
  void loadIndex(const std::string& _name) { chk(pqt_load_index(h_, _name.c_str())); }

  pqt_index* handle() { return h_; }

 private:
  void chk(int rc) const {
    if (rc != PQT_OK) throw Error(rc, pqt_last_error(h_));
  }
  std::vector<uint> getDense(int which) {
    std::vector<uint> v(params().hash_size);
    chk(which == 0 ? pqt_get_db(h_, v.data(), nullptr, nullptr)
                   : pqt_get_db(h_, nullptr, v.data(), nullptr));
    return v;
  }
  pqt_index* h_;
};

}  // namespace pqt

#endif
