// ref_gpu_shim.cu -- TEST INFRASTRUCTURE.  C wrapper around the reference's OWN
// pqt::PerturbationProTree, compiled from /root/reference/pqt/*.cu where they lie
// (nothing copied) into oracle/_ref/libpqt_ref_gpu.so, so that on the GPU box the
// oracle and the product can be cross-checked against the real reference kernels
// (tests/test_ref_gpu.py).  A subclass exposes the protected per-stage methods.
#include <cstdio>
#include <cstring>
#include <vector>

#include "pqt/PerturbationProTree.hh"

namespace {
class RefTree : public pqt::PerturbationProTree {
 public:
  RefTree(uint dim, uint p) : pqt::PerturbationProTree(dim, p, p) {}
  using pqt::PerturbationProTree::computeCBL1L1Dist;
  using pqt::PerturbationProTree::getBins;
  using pqt::PerturbationProTree::getBIGBins2D;
  using pqt::PerturbationProTree::getKBestAssignment;
  using pqt::PerturbationProTree::getKBestAssignment2;
  using pqt::PerturbationProTree::getLineAssignment;
  uint c1() const { return d_nClusters; }
  uint c2() const { return d_nClusters2; }
  uint P() const { return d_p; }
  uint dim() const { return d_dim; }
  uint lineParts() const { return d_lineParts; }
  float* cb1() { return d_multiCodeBook; }
  float* cb2() { return d_multiCodeBook2; }
  float* cbDist() { return d_codeBookDistL1L2; }
  uint* distSeq() { return d_distSeq; }
  void setLinesFromHost(const float* lines, uint N, uint LP) {
    prepareEmptyLambda(N, LP);  // sets d_lineParts, allocates d_lineLambda
    cudaMemcpy(getLine(), lines, (size_t)N * LP * sizeof(float), cudaMemcpyHostToDevice);
    computeCBL1L1Dist(LP);
  }
  void prepSeq(uint k1) { prepareDistSequence(d_nClusters2 * k1, d_p); }
};
}  // namespace

extern "C" {

void* refgpu_create(unsigned dim, unsigned p) { return new RefTree(dim, p); }
void refgpu_destroy(void* h) { delete static_cast<RefTree*>(h); }

void refgpu_read_tree(void* h, const char* path) { static_cast<RefTree*>(h)->readTreeFromFile(path); }

void refgpu_set_db(void* h, unsigned N, const unsigned* prefix, const unsigned* counts,
                   const unsigned* dbidx) {
  static_cast<RefTree*>(h)->setDB(N, prefix, counts, dbidx);
}

void refgpu_set_lines(void* h, const float* lines, unsigned N, unsigned LP) {
  static_cast<RefTree*>(h)->setLinesFromHost(lines, N, LP);
}

// the reference's queryKNN (pqt/PerturbationProTree.cu:8179-8323), host in / host out
void refgpu_query_knn(void* h, const float* Q, unsigned QN, unsigned k, unsigned* idx, float* dist) {
  RefTree* t = static_cast<RefTree*>(h);
  float* dQ = nullptr;
  cudaMalloc(&dQ, (size_t)QN * t->dim() * sizeof(float));
  cudaMemcpy(dQ, Q, (size_t)QN * t->dim() * sizeof(float), cudaMemcpyHostToDevice);
  std::vector<unsigned> ri;
  std::vector<float> rd;
  t->queryKNN(ri, rd, dQ, QN, k);
  std::memcpy(idx, ri.data(), (size_t)QN * k * sizeof(unsigned));
  std::memcpy(dist, rd.data(), (size_t)QN * k * sizeof(float));
  cudaFree(dQ);
}

// Steps A-D through the reference's protected methods, outputs copied to the host
// (layouts of SURVEY.md App. B)
void refgpu_stages(void* h, const float* Q, unsigned QN, unsigned k1, unsigned maxBins,
                   unsigned* assign, float* lut, float* aval, unsigned* aidx, unsigned* bins,
                   unsigned* nbins, float* cbdist) {
  RefTree* t = static_cast<RefTree*>(h);
  const unsigned p = t->P(), c1 = t->c1(), c2 = t->c2(), LP = t->lineParts(), dim = t->dim();
  float *dQ, *dLut, *dAval;
  unsigned *dAssign, *dAidx, *dBins, *dNbins;
  cudaMalloc(&dQ, (size_t)QN * dim * 4);
  cudaMemcpy(dQ, Q, (size_t)QN * dim * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dAssign, (size_t)QN * k1 * p * 4);
  cudaMalloc(&dLut, (size_t)QN * LP * c1 * 4);
  cudaMalloc(&dAval, (size_t)QN * p * k1 * c2 * 4);
  cudaMalloc(&dAidx, (size_t)QN * p * k1 * c2 * 4);
  cudaMalloc(&dBins, (size_t)QN * maxBins * 4);
  cudaMalloc(&dNbins, (size_t)QN * 4);
  t->prepSeq(k1);
  t->getKBestAssignment(dAssign, t->cb1(), dQ, c1, QN, k1);
  t->getLineAssignment(dLut, dQ, QN);
  t->getKBestAssignment2(dAval, dAidx, t->cb2(), dQ, c2, QN, dAssign, c1, k1);
  t->getBins(dBins, dNbins, dAval, dAidx, QN, k1, 4096, maxBins);
  cudaDeviceSynchronize();
  cudaMemcpy(assign, dAssign, (size_t)QN * k1 * p * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(lut, dLut, (size_t)QN * LP * c1 * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(aval, dAval, (size_t)QN * p * k1 * c2 * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(aidx, dAidx, (size_t)QN * p * k1 * c2 * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(bins, dBins, (size_t)QN * maxBins * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(nbins, dNbins, (size_t)QN * 4, cudaMemcpyDeviceToHost);
  if (cbdist && t->cbDist())
    cudaMemcpy(cbdist, t->cbDist(), (size_t)c1 * c1 * LP * 4, cudaMemcpyDeviceToHost);
  cudaFree(dQ); cudaFree(dAssign); cudaFree(dLut); cudaFree(dAval); cudaFree(dAidx);
  cudaFree(dBins); cudaFree(dNbins);
}

// the reference's own index build: buildKBestDB + lineDist (LP hard-set to 16 there)
void refgpu_build(void* h, const float* X, unsigned N, unsigned hash_size, unsigned* prefix,
                  unsigned* counts, unsigned* dbidx, float* lines16) {
  RefTree* t = static_cast<RefTree*>(h);
  float* dX = nullptr;
  cudaMalloc(&dX, (size_t)N * t->dim() * 4);
  cudaMemcpy(dX, X, (size_t)N * t->dim() * 4, cudaMemcpyHostToDevice);
  t->buildKBestDB(dX, N);
  t->lineDist(dX, N);
  cudaDeviceSynchronize();
  cudaMemcpy(prefix, t->getBinPrefix(), (size_t)hash_size * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(counts, t->getBinCounts(), (size_t)hash_size * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(dbidx, t->getDBIdx(), (size_t)N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(lines16, t->getLine(), (size_t)N * 16 * 4, cudaMemcpyDeviceToHost);
  cudaFree(dX);
}

// the 1-B variant: prepare2DDistSequence(512) + the steps of queryBIGKNNRerank2
// (pqt/PerturbationProTree.cu:8596-8701) up to getBIGBins2D, bins copied out (first `cap`
// slots per query), then the public queryBIGKNNRerank2 itself with the resident line codes
void refgpu_big(void* h, const float* Q, unsigned QN, unsigned k, unsigned cap, unsigned* bins_out,
                unsigned* nbins_out, unsigned* idx, float* dist, unsigned* seq2d_out) {
  RefTree* t = static_cast<RefTree*>(h);
  const unsigned p = t->P(), c1 = t->c1(), c2 = t->c2(), LP = t->lineParts(), dim = t->dim();
  const unsigned k1 = 16, maxBins = 64 * 8192;
  t->prepare2DDistSequence(512);
  float *dQ, *dLut, *dAval;
  unsigned *dAssign, *dAidx, *dBins, *dNbins;
  cudaMalloc(&dQ, (size_t)QN * dim * 4);
  cudaMemcpy(dQ, Q, (size_t)QN * dim * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&dAssign, (size_t)QN * k1 * p * 4);
  cudaMalloc(&dLut, (size_t)QN * LP * c1 * 4);
  cudaMalloc(&dAval, (size_t)QN * p * k1 * c2 * 4);
  cudaMalloc(&dAidx, (size_t)QN * p * k1 * c2 * 4);
  cudaMalloc(&dBins, (size_t)QN * maxBins * 4);
  cudaMalloc(&dNbins, (size_t)QN * 4);
  t->getKBestAssignment(dAssign, t->cb1(), dQ, c1, QN, k1);
  t->getLineAssignment(dLut, dQ, QN);
  t->getKBestAssignment2(dAval, dAidx, t->cb2(), dQ, c2, QN, dAssign, c1, k1);
  t->getBIGBins2D(dBins, dNbins, dAval, dAidx, QN, k1, k, maxBins);
  cudaDeviceSynchronize();
  for (unsigned q = 0; q < QN; q++)
    cudaMemcpy(bins_out + (size_t)q * cap, dBins + (size_t)q * maxBins, (size_t)cap * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(nbins_out, dNbins, (size_t)QN * 4, cudaMemcpyDeviceToHost);
  if (seq2d_out) cudaMemcpy(seq2d_out, t->distSeq(), (size_t)10 * 65536 * 4, cudaMemcpyDeviceToHost);
  cudaFree(dAssign); cudaFree(dLut); cudaFree(dAval); cudaFree(dAidx); cudaFree(dBins); cudaFree(dNbins);
  std::vector<unsigned> ri;
  std::vector<float> rd;
  t->queryBIGKNNRerank2(ri, rd, dQ, QN, k, t->getLine());
  std::memcpy(idx, ri.data(), (size_t)QN * k * sizeof(unsigned));
  std::memcpy(dist, rd.data(), (size_t)QN * k * sizeof(float));
  cudaFree(dQ);
}

unsigned refgpu_hash_size(void) { return HASH_SIZE; }

}  // extern "C"
