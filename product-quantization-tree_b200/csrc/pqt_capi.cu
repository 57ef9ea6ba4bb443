// pqt_capi.cu -- the C ABI of include/pqt_b200.h over the sm_100a kernels.
//
// Host-side orchestration only: handle state, load-time preparation, launches.
// Reference citations are file:line into /root/reference.
#include "../../include/pqt_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <utility>
#include <vector>

#include "build_kernels.cuh"
#include "build_fast_kernels.cuh"
#include "common.cuh"
#include "index_kernels.cuh"
#include "query_kernels.cuh"
#include "rerank_kernels.cuh"
#include "big_kernels.cuh"

using namespace pqtb;

namespace {

// Owning device allocation (grow-only arena slot).  Move-only; frees on destruction, so the
// temporaries of a call are released on every early (error) return as well.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) {
    o.p = nullptr;
    o.bytes = 0;
  }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p;
      bytes = o.bytes;
      o.p = nullptr;
      o.bytes = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

}  // namespace

struct pqt_index {
  int device = 0;
  int num_sms = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // device->host result copies, overlapped with the next slab
  cudaStream_t aux_stream = nullptr;   // multi-GPU: the LUTs of the other ranks' queries, beside Steps A-E1
  cudaEvent_t aux_ev[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> slab_ev;
  mutable std::string err;
  pqt_params prm{};

  // tree (readTreeFromFile / createTree state)
  uint32_t dim = 0, p = 0, c1 = 0, c2 = 0, vl = 0;
  std::vector<float> h_cb1, h_cb2;
  DevBuf d_cb1, d_cb2;
  DevBuf d_cb1T, d_cb2T;  // cb1T[dim][c1], cb2T[p][c1][vl][c2] (coalesced reads in tables_warp_kernel)

  // traversal order (prepareDistSequence), cached per (m, p)
  DevBuf d_distseq;
  DevBuf d_seqnib;  // same codes with the rank of part j in nibble j
  DevBuf d_seq2d;  // prepare2DDistSequence(512): [10][65536]
  std::vector<uint32_t> h_seq2d;
  DevBuf s_topv, s_topi;  // [QN][p][64] best Step-C entries (1-B variant)
  DevBuf g_bigbins, g_bignbins, g_phases;
  uint32_t dbg_big_QN = 0, dbg_big_cap = 0;
  DevBuf d_seqsorted;  // per 4096-batch, sorted by the ranks of all parts but the last (bins3)
  DevBuf d_seqfine;    // the same per 1024-batch (bins3 on dense indexes: finer early stop)
  DevBuf d_seqmega, d_seqplain;  // the same per 16384 codes + plain nibble codes (bins4)
  std::vector<uint32_t> h_distseq;
  uint32_t seq_m = 0, seq_p = 0;

  // DB (setDB / buildKBestDB state)
  bool has_db = false;
  uint32_t N = 0;
  uint32_t rank = 0, world = 1;
  uint32_t pos_lo = 0, pos_hi = 0;
  uint32_t db_hash_size = 0;
  uint32_t n_nonempty = 0;
  DevBuf d_bitmap, d_rank_base, d_cprefix, d_dbidx;  // d_dbidx: full [N], bin order

  // line codes (lineDist / prepareEmptyLambda state)
  bool has_lines = false;
  uint32_t LP = 0, sl = 0;
  DevBuf d_codes;  // [pos_hi - pos_lo][LP], bin order
  DevBuf d_cbd, d_cbd_dup;
  // chunked line encoding in progress (pqt_line_dist_begin .. _end)
  bool line_build = false;
  DevBuf b_inv, b_stage, b_x;  // inv[N]: bin-order position of every id; per-chunk staging

  // per-batch scratch
  DevBuf s_q, s_lut, s_idx16, s_cand, s_nvec, s_val, s_idx, s_outd, s_outi;
  DevBuf s_rootpos, s_ridx, s_nroot;  // candidates without repeats (bins3_kernel -> scan / rank kernels)
  bool have_roots = false;            // the last bin walk filled them
  // debug
  bool debug = false;
  uint32_t dbg_QN = 0, dbg_k = 0, dbg_maxvec = 0;
  DevBuf g_assign, g_lut, g_aval, g_aidx, g_bins, g_nbins, g_sel;
  // multi-GPU exchange (fused scan + peer stores)
  DevBuf x_val;    // own [q_per_rank][max_vec] distances, written by the shards (peers)
  DevBuf x_inbox;  // [q_per_rank * world][max_vec] (local position, entry number): this shard's
                   // candidates of every query of the batch, written by the queries' owners
  DevBuf x_cnt;    // [q_per_rank * world] length of every inbox row
  uint32_t x_q_per_rank = 0, x_max_vec = 0, x_world = 0;
  uint32_t x_lut_QN = 0;  // queries whose LUT the last pqt_shard_dispatch call left in s_lut
  float* x_peer_val[8] = {nullptr};
  uint2* x_peer_inbox[8] = {nullptr};
  uint32_t* x_peer_cnt[8] = {nullptr};
  bool x_ipc_opened[8] = {false};
  uint32_t scratch_queries = 0;  // largest slab of the batch in flight (query_common)
  DevBuf d_sched;  // one uint32: work counter of the fused scan+rank kernel
  DevBuf s_ids;    // [QN][max_vec] ids of the candidates when the scan kernel gathers them (PQT_SCAN_IDS=1)
  DevBuf d_exact;  // two uint64: queries ranked by the exact-network fallback, queries with re-ordered ties

  // profiling
  bool split_ranked = false;  // the last run_scan_chain ran the split pipeline (scan kernel + ranking kernel)
  bool profile = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  pqt_stats stats{};
};

namespace {

int fail(const pqt_index* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

#define CU_TRY(h, call)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess)                                                              \
      return fail(h, e__ == cudaErrorMemoryAllocation ? PQT_ERR_NOMEM : PQT_ERR_CUDA,    \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,     \
                  __LINE__);                                                             \
  } while (0)

#define PQ_TRY(call)          \
  do {                        \
    int rc__ = (call);        \
    if (rc__ != PQT_OK) return rc__; \
  } while (0)

bool is_pow2(uint32_t x) { return x && !(x & (x - 1)); }

constexpr uint32_t kSlabQueries = 2048;  // queries per pipelined slab (host outputs)

int check_tree_shape(const pqt_index* h, uint32_t dim, uint32_t p, uint32_t c1, uint32_t c2) {
  if (!dim || !p || !c1 || !c2) return fail(h, PQT_ERR_INVALID, "zero-sized tree shape");
  if (dim % p) return fail(h, PQT_ERR_INVALID, "dim %u not divisible by p %u", dim, p);
  uint32_t vl = dim / p;
  if (!is_pow2(vl) || vl > 128)
    return fail(h, PQT_ERR_INVALID,
                "segment length dim/p = %u must be a power of two <= 128 (the reference's tree "
                "reduction, pqt/PerturbationProTree.cu:7151-7159, needs 2^n)",
                vl);
  if (p > 8) return fail(h, PQT_ERR_INVALID, "p = %u > 8 (16^p traversal codes overflow uint32)", p);
  if (c1 > 127) return fail(h, PQT_ERR_INVALID, "c1 = %u > 127 (lineDescr stores char p1, p2)", c1);
  if (pow2ceil(c1) > 1024) return fail(h, PQT_ERR_INVALID, "c1 too large");
  return PQT_OK;
}

// ---- prepareDistSequence (pqt/ProTree.cu:128-207) on the host, cached ---------------
int ensure_dist_seq(pqt_index* h, uint32_t max_cluster) {
  uint32_t m = max_cluster > 16 ? 16 : max_cluster;
  if (h->d_distseq.p && h->seq_m == m && h->seq_p == h->p) return PQT_OK;
  uint64_t n64 = 1;
  for (uint32_t j = 0; j < h->p; j++) n64 *= m;
  // the reference computes uint nVec = pow(m, p) (pqt/ProTree.cu:139): it wraps for p = 8 and the
  // host sort of m^p pairs is impractical long before that
  if (n64 > (1ull << 24))
    return fail(h, PQT_ERR_INVALID, "traversal table of %u^%u codes is too large (p <= 6 at k1*c2 >= 16)", m, h->p);
  uint32_t nvec = (uint32_t)n64;
  std::vector<std::pair<float, uint32_t>> d(nvec);
  std::vector<uint32_t> den(h->p);
  den[0] = 1;
  for (uint32_t j = 1; j < h->p; j++) den[j] = den[j - 1] * m;
  for (uint32_t i = 0; i < nvec; i++) {
    float dist = 0.f;
    for (uint32_t j = 0; j < h->p; j++) dist += std::sqrt((float)((i / den[j]) % m));
    d[i] = std::make_pair(dist, i);
  }
  std::sort(d.begin(), d.end());
  h->h_distseq.assign(kNumDistSeq, 0u);
  uint32_t keep = std::min<uint32_t>(nvec, kNumDistSeq);
  for (uint32_t i = 0; i < keep; i++) h->h_distseq[i] = d[i].second;
  CU_TRY(h, h->d_distseq.ensure(kNumDistSeq * sizeof(uint32_t)));
  CU_TRY(h, cudaMemcpyAsync(h->d_distseq.p, h->h_distseq.data(), kNumDistSeq * sizeof(uint32_t),
                            cudaMemcpyHostToDevice, h->stream));
  std::vector<uint32_t> nib(kNumDistSeq, 0u);
  for (uint32_t i = 0; i < kNumDistSeq; i++) {
    uint32_t code = h->h_distseq[i], v = 0;
    for (uint32_t j = 0; j < h->p; j++) v |= ((code / den[j]) % m) << (4 * j);
    // bins2_kernel layout: batch of kBins2Threads*kProbesPerThread probes stored [r][thread]
    const uint32_t batch = kBins2Threads * kProbesPerThread;
    const uint32_t b = i / batch, u = i % batch;
    nib[b * batch + (u % kProbesPerThread) * kBins2Threads + u / kProbesPerThread] = v;
  }
  if (h->p <= 4) {
    // bins3_kernel: inside each batch visit the probes sorted by (ranks of parts 0..p-2,
    // rank of the last part); entry = (rank inside the batch << 16) | nibble code
    const uint32_t last_shift = 4 * (h->p - 1);
    for (int pass = 0; pass < 2; pass++) {
      std::vector<uint32_t> sorted(kNumDistSeq, 0u);
      const uint32_t batch = pass == 0 ? (uint32_t)kBins3Batch : (uint32_t)kBins3FineBatch;
      std::vector<std::pair<uint32_t, uint32_t>> key(batch);
      for (uint32_t b = 0; b < kNumDistSeq / batch; b++) {
        for (uint32_t u = 0; u < batch; u++) {
          uint32_t code = h->h_distseq[b * batch + u], v = 0;
          for (uint32_t j = 0; j < h->p; j++) v |= ((code / den[j]) % m) << (4 * j);
          uint32_t prefix = v & ((1u << last_shift) - 1u), last = v >> last_shift;
          key[u] = std::make_pair((prefix << 4) | last, (u << 16) | v);
        }
        std::sort(key.begin(), key.end());
        // thread-major storage [r][thread]: lane-consecutive reads hit consecutive entries
        for (uint32_t e = 0; e < batch; e++) sorted[b * batch + e] = key[e].second;
      }
      DevBuf& dst = pass == 0 ? h->d_seqsorted : h->d_seqfine;
      CU_TRY(h, dst.ensure(kNumDistSeq * sizeof(uint32_t)));
      CU_TRY(h, cudaMemcpyAsync(dst.p, sorted.data(), kNumDistSeq * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, h->stream));
      CU_TRY(h, cudaStreamSynchronize(h->stream));
    }
    // bins4_kernel: the same order over mega-batches of 16384 codes, plus the nibble codes
    // in plain traversal order (for the few kept probes)
    std::vector<uint32_t> mega(kNumDistSeq, 0u), plain(kNumDistSeq, 0u);
    std::vector<std::pair<uint32_t, uint32_t>> mkey(kBins4Mega);
    for (uint32_t b = 0; b < kNumDistSeq / kBins4Mega; b++) {
      for (uint32_t u = 0; u < (uint32_t)kBins4Mega; u++) {
        uint32_t code = h->h_distseq[b * kBins4Mega + u], v = 0;
        for (uint32_t j = 0; j < h->p; j++) v |= ((code / den[j]) % m) << (4 * j);
        plain[b * kBins4Mega + u] = v;
        uint32_t prefix = v & ((1u << last_shift) - 1u), last = v >> last_shift;
        mkey[u] = std::make_pair((prefix << 4) | last, (u << 16) | v);
      }
      std::sort(mkey.begin(), mkey.end());
      for (uint32_t e = 0; e < (uint32_t)kBins4Mega; e++) mega[b * kBins4Mega + e] = mkey[e].second;
    }
    CU_TRY(h, h->d_seqmega.ensure(kNumDistSeq * sizeof(uint32_t)));
    CU_TRY(h, h->d_seqplain.ensure(kNumDistSeq * sizeof(uint32_t)));
    CU_TRY(h, cudaMemcpyAsync(h->d_seqmega.p, mega.data(), kNumDistSeq * sizeof(uint32_t),
                              cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(h->d_seqplain.p, plain.data(), kNumDistSeq * sizeof(uint32_t),
                              cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
  }
  CU_TRY(h, h->d_seqnib.ensure(kNumDistSeq * sizeof(uint32_t)));
  CU_TRY(h, cudaMemcpyAsync(h->d_seqnib.p, nib.data(), kNumDistSeq * sizeof(uint32_t),
                            cudaMemcpyHostToDevice, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->seq_m = m;
  h->seq_p = h->p;
  return PQT_OK;
}

// ---- prepare2DDistSequence(512) (pqt/ProTree.cu:50-126) on the host, cached ----------
int ensure_dist_seq_2d(pqt_index* h) {
  if (h->d_seq2d.p) return PQT_OK;
  const uint32_t mc = kBigDistCluster, nvec = mc * mc;
  const uint32_t copy = std::min<uint32_t>(nvec, kNumDistSeq);
  h->h_seq2d.assign((size_t)kNumAnisoDir * kNumDistSeq, 0u);
  std::vector<std::pair<float, uint32_t>> d(nvec);
  for (uint32_t slope = 0; slope < kNumAnisoDir; slope++) {
    const float s = (float)std::pow(0.9 * (double)PQTB_ANISO_BASE, (double)((int)slope - (int)(kNumAnisoDir / 2)));
    for (uint32_t i = 0; i < nvec; i++) {
      const float x = (float)(i % mc), y = (float)(i / mc);
      const float n = 0.8f;
      const float px = std::pow(x, n);
      volatile float py = s * std::pow(y, n);  // separate roundings, as compiled for the host
      d[i] = std::make_pair(px + py, i);
    }
    std::sort(d.begin(), d.end());
    for (uint32_t i = 0; i < copy; i++) h->h_seq2d[(size_t)slope * kNumDistSeq + i] = d[i].second;
  }
  CU_TRY(h, h->d_seq2d.ensure(h->h_seq2d.size() * sizeof(uint32_t)));
  CU_TRY(h, cudaMemcpyAsync(h->d_seq2d.p, h->h_seq2d.data(), h->h_seq2d.size() * sizeof(uint32_t),
                            cudaMemcpyHostToDevice, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return PQT_OK;
}

int upload_tree(pqt_index* h) {
  CU_TRY(h, h->d_cb1.ensure(h->h_cb1.size() * sizeof(float)));
  CU_TRY(h, h->d_cb2.ensure(h->h_cb2.size() * sizeof(float)));
  CU_TRY(h, cudaMemcpyAsync(h->d_cb1.p, h->h_cb1.data(), h->h_cb1.size() * sizeof(float),
                            cudaMemcpyHostToDevice, h->stream));
  CU_TRY(h, cudaMemcpyAsync(h->d_cb2.p, h->h_cb2.data(), h->h_cb2.size() * sizeof(float),
                            cudaMemcpyHostToDevice, h->stream));
  // transposed copies: lanes run over centroids, so centroid must be the fastest index
  {
    const uint32_t dim = h->dim, c1 = h->c1, c2 = h->c2, p = h->p, vl = h->vl;
    std::vector<float> t1((size_t)dim * c1), t2((size_t)p * c1 * vl * c2);
    for (uint32_t c = 0; c < c1; c++)
      for (uint32_t d = 0; d < dim; d++) t1[(size_t)d * c1 + c] = h->h_cb1[(size_t)c * dim + d];
    for (uint32_t part = 0; part < p; part++)
      for (uint32_t l1 = 0; l1 < c1; l1++)
        for (uint32_t l2 = 0; l2 < c2; l2++)
          for (uint32_t t = 0; t < vl; t++)
            t2[(((size_t)part * c1 + l1) * vl + t) * c2 + l2] =
                h->h_cb2[(((size_t)part * c1 + l1) * c2 + l2) * vl + t];
    CU_TRY(h, h->d_cb1T.ensure(t1.size() * sizeof(float)));
    CU_TRY(h, h->d_cb2T.ensure(t2.size() * sizeof(float)));
    CU_TRY(h, cudaMemcpyAsync(h->d_cb1T.p, t1.data(), t1.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(h->d_cb2T.p, t2.data(), t2.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
  }
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->has_lines = false;  // cbDist depends on cb1
  h->has_db = false;     // the bin directory was built for the previous tree's c1 / c2 / p
  return PQT_OK;
}

uint32_t candidate_width(const pqt_index* h, uint32_t k) {
  return h->prm.max_vec ? h->prm.max_vec : pow2ceil(k);
}

// ---- builds bitmap / rank_base / cprefix from dense device arrays ---------------------
int build_directory(pqt_index* h, const uint32_t* d_counts, const uint32_t* d_prefix,
                    uint32_t hash_size, uint32_t N) {
  const size_t nwords = ((size_t)hash_size + 31) >> 5;
  const size_t ngroups = (nwords + 7) >> 3;
  CU_TRY(h, h->d_bitmap.ensure((ngroups * 8) * sizeof(uint32_t)));
  CU_TRY(h, cudaMemsetAsync(h->d_bitmap.p, 0, ngroups * 8 * sizeof(uint32_t), h->stream));
  CU_TRY(h, h->d_rank_base.ensure((ngroups + 1) * sizeof(uint32_t)));
  DevBuf tmp;
  CU_TRY(h, tmp.ensure(scan_tmp_words(ngroups + 1) * sizeof(uint32_t)));
  const int blocks = h->num_sms * 8;
  bitmap_build_kernel<<<blocks, 256, 0, h->stream>>>(d_counts, hash_size, h->d_bitmap.as<uint32_t>());
  CU_TRY(h, cudaMemsetAsync(h->d_rank_base.p, 0, (ngroups + 1) * sizeof(uint32_t), h->stream));
  group_popc_kernel<<<blocks, 256, 0, h->stream>>>(h->d_bitmap.as<uint32_t>(), nwords, ngroups,
                                                   h->d_rank_base.as<uint32_t>());
  device_exscan_u32(h->d_rank_base.as<uint32_t>(), h->d_rank_base.as<uint32_t>(), ngroups + 1,
                    tmp.as<uint32_t>(), h->stream);
  uint32_t nne = 0;
  CU_TRY(h, cudaMemcpyAsync(&nne, h->d_rank_base.as<uint32_t>() + ngroups, sizeof(uint32_t),
                            cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->n_nonempty = nne;
  CU_TRY(h, h->d_cprefix.ensure(((size_t)nne + 2) * sizeof(uint32_t)));
  cprefix_fill_kernel<<<blocks, 256, 0, h->stream>>>(d_counts, d_prefix, hash_size,
                                                     h->d_bitmap.as<uint32_t>(),
                                                     h->d_rank_base.as<uint32_t>(),
                                                     h->d_cprefix.as<uint32_t>());
  CU_TRY(h, cudaMemcpyAsync(h->d_cprefix.as<uint32_t>() + nne, &N, sizeof(uint32_t),
                            cudaMemcpyHostToDevice, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaGetLastError());
  tmp.release();
  h->db_hash_size = hash_size;
  h->N = N;
  h->pos_lo = (uint32_t)((uint64_t)N * h->rank / h->world);
  h->pos_hi = (uint32_t)((uint64_t)N * (h->rank + 1) / h->world);
  h->has_db = true;
  h->has_lines = false;
  return PQT_OK;
}

int check_lp(const pqt_index* h, uint32_t LP) {
  if (!LP || h->dim % LP) return fail(h, PQT_ERR_INVALID, "dim %u not divisible by lineparts %u", h->dim, LP);
  uint32_t sl = h->dim / LP;
  if (!is_pow2(LP) || LP > 32)
    return fail(h, PQT_ERR_INVALID,
                "lineparts = %u must be a power of two <= 32 (warp-shuffle reduction, "
                "pqt/PerturbationProTree.cu:5183-5187)",
                LP);
  if (!is_pow2(sl) || sl > 128) return fail(h, PQT_ERR_INVALID, "dim/lineparts = %u must be 2^n <= 128", sl);
  return PQT_OK;
}

int compute_cbd(pqt_index* h, uint32_t LP) {
  const size_t n = (size_t)h->c1 * h->c1 * LP;
  CU_TRY(h, h->d_cbd.ensure(n * sizeof(float)));
  CU_TRY(h, h->d_cbd_dup.ensure((size_t)h->c1 * h->c1 * 32 * sizeof(float)));
  cb_dist_kernel<<<64, 256, 0, h->stream>>>(h->d_cb1.as<float>(), h->c1, h->dim, LP, h->dim / LP,
                                            h->d_cbd.as<float>(), h->d_cbd_dup.as<float>());
  CU_TRY(h, cudaGetLastError());
  return PQT_OK;
}

struct QueryPlan {
  uint32_t QN, k, max_vec;
  const float* dQ;
};

// Steps D + E1 for p <= 4: bins4_kernel (visiting order sorted over 16384 codes) unless the
// index is so dense that the walk would stop inside the first codes, where bins3_kernel
// (early stop every 4096 codes) does less work.  Same output either way.
int launch_bins_p4(pqt_index* h, const pqt_params& P, uint32_t max_vec, uint32_t nq,
                   const uint32_t* idx16, uint32_t* cand_pos, uint32_t* n_vec, uint32_t* dbg_bins,
                   uint32_t* dbg_nbins, bool want_roots = false) {
  h->have_roots = false;
  const uint32_t n_probes = P.max_trials * P.bin_threads;
  const double density = (double)h->n_nonempty / (double)std::max<uint32_t>(1u, h->db_hash_size);
  // the walk ends after max_bins kept bins or max_vec listed candidates, whichever comes first
  const double vec_per_probe = (double)h->N / (double)std::max<uint32_t>(1u, h->db_hash_size);
  const bool early_stop_likely = density * kBins3Batch * 2.0 >= (double)P.max_bins ||
                                 (!dbg_bins && vec_per_probe * kBins3Batch * 2.0 >= (double)max_vec);
  static const int force = getenv("PQT_BINS_KERNEL") ? atoi(getenv("PQT_BINS_KERNEL")) : 0;  // 3 / 4: A/B runs
  const bool use4 = force == 4 || (force != 3 && !early_stop_likely && n_probes > kBins3Batch);
  uint32_t grid = std::min<uint32_t>(nq, (uint32_t)h->num_sms * 6);
  if (use4) {
    Bins4Args a{};
    a.idx16 = idx16;
    a.seq_mega = h->d_seqmega.as<uint32_t>();
    a.seq_plain = h->d_seqplain.as<uint32_t>();
    a.dir.bitmap = h->d_bitmap.as<uint32_t>();
    a.dir.rank_base = h->d_rank_base.as<uint32_t>();
    a.dir.cprefix = h->d_cprefix.as<uint32_t>();
    a.hash = make_magicmod(h->db_hash_size);
    a.QN = nq; a.p = h->p; a.c1c2 = h->c1 * h->c2;
    a.n_probes = n_probes;
    a.max_bins = P.max_bins; a.max_vec_per_bin = P.max_vec_per_bin; a.max_vec = max_vec;
    a.cand_pos = cand_pos; a.n_vec = n_vec; a.dbg_bins = dbg_bins; a.dbg_nbins = dbg_nbins;
    const size_t smem = (size_t)(P.max_bins + kBins4Mega / 32 + 2 * 256 + 2 * kBins2Threads + 32) * 4;
    if (h->p <= 2) {
      if (smem > 48 * 1024)
        CU_TRY(h, cudaFuncSetAttribute(bins4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bins4_kernel<1><<<grid, kBins2Threads, smem, h->stream>>>(a);
    } else {
      if (smem > 48 * 1024)
        CU_TRY(h, cudaFuncSetAttribute(bins4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bins4_kernel<2><<<grid, kBins2Threads, smem, h->stream>>>(a);
    }
  } else {
    Bins3Args a{};
    a.idx16 = idx16;
    a.seq_sorted = h->d_seqsorted.as<uint32_t>();
    a.dir.bitmap = h->d_bitmap.as<uint32_t>();
    a.dir.rank_base = h->d_rank_base.as<uint32_t>();
    a.dir.cprefix = h->d_cprefix.as<uint32_t>();
    a.hash = make_magicmod(h->db_hash_size);
    a.QN = nq; a.p = h->p; a.c1c2 = h->c1 * h->c2;
    a.n_probes = n_probes;
    a.max_bins = P.max_bins; a.max_vec_per_bin = P.max_vec_per_bin; a.max_vec = max_vec;
    a.cand_pos = cand_pos; a.n_vec = n_vec; a.dbg_bins = dbg_bins; a.dbg_nbins = dbg_nbins;
    // dense index: a 1024-code batch already lists a good part of the max_vec candidates, so
    // the candidate-count stop is checked after every 1024 codes; otherwise 4096-code batches
    const bool fine = vec_per_probe * kBins3FineBatch * 8.0 >= (double)max_vec && !dbg_bins;
    if (fine) a.seq_sorted = h->d_seqfine.as<uint32_t>();
    if (want_roots && max_vec <= 65536) {
      CU_TRY(h, h->s_rootpos.ensure((size_t)nq * max_vec * 4));
      CU_TRY(h, h->s_ridx.ensure((size_t)nq * max_vec * 2));
      CU_TRY(h, h->s_nroot.ensure((size_t)nq * 4));
      a.root_pos = h->s_rootpos.as<uint32_t>();
      a.ridx = h->s_ridx.as<uint16_t>();
      a.n_root = h->s_nroot.as<uint32_t>();
      h->have_roots = true;
    }
    const size_t smem = bins3_smem_bytes(P.max_bins, fine ? kBins3FineProbes : kProbesPerThread, h->have_roots);
    // never launch more CTAs than are resident at once; they draw their queries from a counter
    grid = std::min<uint32_t>(grid, (uint32_t)h->num_sms * std::max<uint32_t>(1u, std::min<uint32_t>(6u, (uint32_t)((227 * 1024) / (smem + 1024)))));
    CU_TRY(h, h->d_sched.ensure(64));
    CU_TRY(h, cudaMemsetAsync(h->d_sched.as<uint32_t>() + 2, 0, 4, h->stream));
    a.next_query = h->d_sched.as<uint32_t>() + 2;
#define LAUNCH_BINS3(NP, PPTV)                                                                       \
  do {                                                                                               \
    if (smem > 48 * 1024)                                                                            \
      CU_TRY(h, cudaFuncSetAttribute(bins3_kernel<NP, PPTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    bins3_kernel<NP, PPTV><<<grid, kBins2Threads, smem, h->stream>>>(a);                             \
  } while (0)
    if (h->p <= 2) {
      if (fine) LAUNCH_BINS3(1, kBins3FineProbes); else LAUNCH_BINS3(1, kProbesPerThread);
    } else {
      if (fine) LAUNCH_BINS3(2, kBins3FineProbes); else LAUNCH_BINS3(2, kProbesPerThread);
    }
#undef LAUNCH_BINS3
  }
  CU_TRY(h, cudaGetLastError());
  h->stats.kernel_launches++;
  return PQT_OK;
}

// Steps A..E2 (distance part) for QN queries already on the device; fills
// val/idx [QN][max_vec].  Records profile events ev[0..3] when enabled.
// With fused_out_* set (single GPU) the scan, the ranking and the first-k emit run in one
// kernel and d_val/d_idx are not touched.
int run_scan_chain(pqt_index* h, const float* dQ, uint32_t QN, uint32_t k, float* d_val,
                   uint32_t* d_idx, float* fused_out_dist = nullptr,
                   uint32_t* fused_out_idx = nullptr, bool big = false) {
  const uint32_t max_vec = candidate_width(h, k);
  pqt_params P = h->prm;
  if (big) P.k1 = P.big_k1;  // queryBIGKNNRerank2 :8604
  PQ_TRY(ensure_dist_seq(h, h->c2 * P.k1));  // :8191
  const uint32_t m = h->seq_m;
  const uint32_t n = P.k1 * h->c2;
  if (m > n) return fail(h, PQT_ERR_INVALID, "k1*c2 < traversal width");

  // scratch is sized for the largest slab of the call, so that it never grows (= device-wide
  // synchronisation in cudaFree) between the slabs of a pipelined batch
  const uint32_t QNa = std::max(QN, h->scratch_queries);
  CU_TRY(h, h->s_lut.ensure((size_t)QNa * h->c1 * 32 * sizeof(float)));
  CU_TRY(h, h->s_idx16.ensure((size_t)QNa * h->p * 16 * sizeof(uint32_t)));
  CU_TRY(h, h->s_cand.ensure((size_t)QNa * max_vec * sizeof(uint32_t)));
  CU_TRY(h, h->s_nvec.ensure((size_t)QNa * sizeof(uint32_t)));
  if (h->debug) {
    CU_TRY(h, h->g_assign.ensure((size_t)QN * P.k1 * h->p * 4));
    CU_TRY(h, h->g_lut.ensure((size_t)QN * h->LP * h->c1 * 4));
    CU_TRY(h, h->g_aval.ensure((size_t)QN * h->p * n * 4));
    CU_TRY(h, h->g_aidx.ensure((size_t)QN * h->p * n * 4));
    CU_TRY(h, h->g_bins.ensure((size_t)QN * P.max_bins * 4));
    CU_TRY(h, h->g_nbins.ensure((size_t)QN * 4));
    CU_TRY(h, h->g_sel.ensure((size_t)QN * max_vec * 4));
    h->dbg_QN = QN;
    h->dbg_k = k;
    h->dbg_maxvec = max_vec;
  }

  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[0], h->stream));
  // ---- Steps A+B+C
  {
    TablesArgs a{};
    a.Q = dQ;
    a.cb1 = h->d_cb1.as<float>();
    a.cb2 = h->d_cb2.as<float>();
    a.QN = QN; a.dim = h->dim; a.p = h->p; a.c1 = h->c1; a.c2 = h->c2; a.LP = h->LP;
    a.k1 = P.k1; a.vl = h->vl; a.sl = h->sl;
    a.npA = pow2ceil(h->c1);
    a.npC = pow2ceil(n);
    a.m = m;
    a.lut_dup = h->s_lut.as<float>();
    a.idx16 = h->s_idx16.as<uint32_t>();
    if (big) {
      CU_TRY(h, h->s_topv.ensure((size_t)QNa * h->p * kBigKMax * 4));
      CU_TRY(h, h->s_topi.ensure((size_t)QNa * h->p * kBigKMax * 4));
      a.top_val = h->s_topv.as<float>();
      a.top_idx = h->s_topi.as<uint32_t>();
      a.top_n = kBigKMax;
    }
    if (h->debug) {
      a.dbg_assign = h->g_assign.as<uint32_t>();
      a.dbg_lut = h->g_lut.as<float>();
      a.dbg_aval = h->g_aval.as<float>();
      a.dbg_aidx = h->g_aidx.as<uint32_t>();
    }
    const bool warp_path = h->c1 <= 32 && P.k1 <= 32 && (h->LP % h->p) == 0 &&
                           (h->vl == 8 || h->vl == 16 || h->vl == 32) && a.npC >= 2 && (h->dim % 4) == 0;
    if (warp_path) {
      TablesWarpArgs w{};
      w.t = a;
      w.cb1T = h->d_cb1T.as<float>();
      w.cb2T = h->d_cb2T.as<float>();
      size_t smem = (size_t)(h->dim + h->c1 * 32 + 2 * kTablesWarps * a.npC) * 4;
      uint32_t grid = std::min<uint32_t>(QN, (uint32_t)h->num_sms * 12);
      if (h->vl == 32 && h->c1 == 32 && h->c2 == 32) {  // the SIFT-shaped configuration
        tables_warp_kernel<32, 32><<<grid, kTablesWarps * 32, smem, h->stream>>>(w);
      } else {
        switch (h->vl) {
          case 8: tables_warp_kernel<8, 0><<<grid, kTablesWarps * 32, smem, h->stream>>>(w); break;
          case 16: tables_warp_kernel<16, 0><<<grid, kTablesWarps * 32, smem, h->stream>>>(w); break;
          default: tables_warp_kernel<32, 0><<<grid, kTablesWarps * 32, smem, h->stream>>>(w); break;
        }
      }
    } else {
      const uint32_t npMax = std::max(a.npA, a.npC);
      size_t smem = (size_t)(h->dim + 2 * h->p * npMax + P.k1 * h->p) * 4;
      if (smem > 48 * 1024)
        CU_TRY(h, cudaFuncSetAttribute(tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      uint32_t grid = std::min<uint32_t>(QN, (uint32_t)h->num_sms * 16);
      tables_kernel<<<grid, 128, smem, h->stream>>>(a);
    }
    CU_TRY(h, cudaGetLastError());
    h->stats.kernel_launches++;
  }
  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[1], h->stream));
  // Split pipeline (lineparts 16 / 32): a pure streaming scan kernel writes the distances, the
  // ranking kernel (4 CTAs per SM) sorts and emits.  The scan is the HBM-bound part and runs
  // without any ranking state in its way; PQT_SCAN_MODE=fused selects the fused kernel instead
  // (A/B runs).
  const char* env_mode = getenv("PQT_SCAN_MODE");
  const bool want_split = env_mode ? (strcmp(env_mode, "split") == 0) : true;
  const bool will_split = fused_out_dist && want_split && (h->LP == 16 || h->LP == 32) && max_vec >= 256 &&
                          max_vec <= 4096 &&
                          stream_scan_smem_bytes(h->c1, h->LP, h->LP == 32) <= (size_t)227 * 1024;
  // Listing the first occurrences in the bin walk costs that kernel about what the scan kernel
  // saves on one GPU (profiles/r02c_ab_bins_dedupe_1b.log), so it is off here by default; the
  // multi-GPU path, where it also cuts the NVLink traffic by a third, turns it on.
  const bool want_roots = will_split && !h->debug && getenv("PQT_BINS_DEDUPE") && atoi(getenv("PQT_BINS_DEDUPE")) == 1;
  h->have_roots = false;
  // ---- Steps D+E1
  if (big) {
    PQ_TRY(ensure_dist_seq_2d(h));
    BinsBigArgs a{};
    a.top_val = h->s_topv.as<float>();
    a.top_idx = h->s_topi.as<uint32_t>();
    a.seq2d = h->d_seq2d.as<uint32_t>();
    a.dir.bitmap = h->d_bitmap.as<uint32_t>();
    a.dir.rank_base = h->d_rank_base.as<uint32_t>();
    a.dir.cprefix = h->d_cprefix.as<uint32_t>();
    a.hash = make_magicmod(h->db_hash_size);
    a.QN = QN; a.c1c2 = h->c1 * h->c2;
    a.k2 = k;
    a.max_trials = P.big_max_trials; a.max_bins = P.big_max_bins;
    a.max_vec = max_vec; a.max_vec_per_bin = max_vec;  // maxNVecPerBin = pow2ceil(k) (:6525)
    a.list_cap = std::max<uint32_t>(max_vec, 32);
    a.cand_pos = h->s_cand.as<uint32_t>();
    a.n_vec = h->s_nvec.as<uint32_t>();
    if (h->debug) {
      CU_TRY(h, h->g_bigbins.ensure((size_t)QN * a.list_cap * 4));
      CU_TRY(h, h->g_bignbins.ensure((size_t)QN * 4));
      a.dbg_bins = h->g_bigbins.as<uint32_t>();
      a.dbg_nbins = h->g_bignbins.as<uint32_t>();
      h->dbg_big_QN = QN;
      h->dbg_big_cap = a.list_cap;
    }
    size_t smem = (size_t)(4 * kBigInter + 2 * kBigThreads + 4 * kBigKMax + a.list_cap + 32 + 4) * 4 + 2 * kBigThreads;
    if (smem > 48 * 1024)
      CU_TRY(h, cudaFuncSetAttribute(bins_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t grid = std::min<uint32_t>(QN, (uint32_t)h->num_sms * 2);
    CU_TRY(h, h->d_sched.ensure(64));
    CU_TRY(h, cudaMemsetAsync(h->d_sched.as<uint32_t>() + 2, 0, 4, h->stream));
    a.next_query = h->d_sched.as<uint32_t>() + 2;
    bins_big_kernel<<<grid, kBigThreads, smem, h->stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    h->stats.kernel_launches++;
  } else if (h->p <= 4) {
    PQ_TRY(launch_bins_p4(h, P, max_vec, QN, h->s_idx16.as<uint32_t>(), h->s_cand.as<uint32_t>(),
                          h->s_nvec.as<uint32_t>(), h->debug ? h->g_bins.as<uint32_t>() : nullptr,
                          h->debug ? h->g_nbins.as<uint32_t>() : nullptr, want_roots));
  } else {
    Bins2Args a{};
    a.idx16 = h->s_idx16.as<uint32_t>();
    a.seq_nib = h->d_seqnib.as<uint32_t>();
    a.dir.bitmap = h->d_bitmap.as<uint32_t>();
    a.dir.rank_base = h->d_rank_base.as<uint32_t>();
    a.dir.cprefix = h->d_cprefix.as<uint32_t>();
    a.hash = make_magicmod(h->db_hash_size);
    a.QN = QN; a.p = h->p; a.c1c2 = h->c1 * h->c2;
    a.n_probes = P.max_trials * P.bin_threads;
    a.max_bins = P.max_bins; a.max_vec_per_bin = P.max_vec_per_bin; a.max_vec = max_vec;
    a.cand_pos = h->s_cand.as<uint32_t>();
    a.n_vec = h->s_nvec.as<uint32_t>();
    if (h->debug) {
      a.dbg_bins = h->g_bins.as<uint32_t>();
      a.dbg_nbins = h->g_nbins.as<uint32_t>();
    }
    size_t smem = (size_t)(P.max_bins + ((h->p + 1) / 2) * 256 + 32) * 4;
    if (smem > 48 * 1024)
      CU_TRY(h, cudaFuncSetAttribute(bins2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t grid = std::min<uint32_t>(QN, (uint32_t)h->num_sms * 8);
    bins2_kernel<<<grid, kBins2Threads, smem, h->stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    h->stats.kernel_launches++;
  }
  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[2], h->stream));
  // ---- Step E2: ADC scan
  {
    ScanArgs a{};
    a.codes = h->d_codes.as<uint32_t>();
    a.ids = h->d_dbidx.as<uint32_t>() + h->pos_lo;
    a.cand_pos = h->s_cand.as<uint32_t>();
    a.n_vec = h->s_nvec.as<uint32_t>();
    a.lut_dup = h->s_lut.as<float>();
    a.cbd_dup = h->d_cbd_dup.as<float>();
    a.QN = QN; a.c1 = h->c1; a.max_vec = max_vec;
    a.pos_lo = h->pos_lo; a.pos_hi = h->pos_hi;
    a.owns_pad = (h->rank == 0) ? 1u : 0u;
    a.sharded = h->world > 1 ? 1u : 0u;
    a.out_val = d_val;
    a.out_idx = d_idx;
    size_t smem = ((size_t)h->c1 * h->c1 * 32 + 2 * (size_t)h->c1 * 32) * 4 + 64;
    // fused scan + rank: 4 thread groups per CTA over the canonical c^2 table when that fits
    // (LP <= 16), else 2 groups over the replicated table
    const size_t kSmemMax = 227 * 1024;
    // cp.async.bulk moves multiples of 16 bytes: the canonical c^2 table (c1*c1*LP floats) must
    // be one, else the replicated layout (32-float rows) is used
    const bool four = h->LP <= 16 && ((h->c1 * h->c1 * h->LP) % 4u) == 0 &&
                      rerank_smem_bytes(h->c1, h->LP, max_vec, 4, false) <= kSmemMax;
    const size_t smem_fused = four ? rerank_smem_bytes(h->c1, h->LP, max_vec, 4, false)
                                   : rerank_smem_bytes(h->c1, h->LP, max_vec, 2, true);
    h->split_ranked = false;
    if (will_split) {
      CU_TRY(h, h->s_val.ensure((size_t)QNa * max_vec * 4));
      StreamScanArgs sa{};
      sa.codes = h->d_codes.as<uint32_t>();
      // repeated candidates are evaluated once: the bin walk left the first occurrences
      sa.cand_pos = h->have_roots ? h->s_rootpos.as<uint32_t>() : h->s_cand.as<uint32_t>();
      sa.n_vec = h->have_roots ? h->s_nroot.as<uint32_t>() : h->s_nvec.as<uint32_t>();
      sa.lut_dup = h->s_lut.as<float>();
      sa.cbd = h->LP == 32 ? h->d_cbd_dup.as<float>() : h->d_cbd.as<float>();
      sa.QN = QN; sa.c1 = h->c1; sa.max_vec = max_vec;
      sa.out_val = h->s_val.as<float>();
      // PQT_SCAN_IDS=1: the scan also gathers the id of every candidate (one more load in its row
      // pipeline) and the ranking kernel reads id rows instead of chasing position -> id per
      // result.  Measured at 1 B vectors: ranking 1.11 -> 1.04 ms, scan +0.09 ms: off by default
      // (profiles/r02l_ab_ids_by_scan.log).
      static const bool scan_ids = getenv("PQT_SCAN_IDS") && atoi(getenv("PQT_SCAN_IDS")) == 1;
      const bool ids_by_scan = scan_ids && !h->have_roots;
      if (ids_by_scan) {
        CU_TRY(h, h->s_ids.ensure((size_t)QNa * max_vec * 4));
        sa.ids = h->d_dbidx.as<uint32_t>() + h->pos_lo;
        sa.out_id = h->s_ids.as<uint32_t>();
      }
      const size_t ssmem = stream_scan_smem_bytes(h->c1, h->LP, h->LP == 32);
      const uint32_t sgrid = std::min<uint32_t>(QN, (uint32_t)h->num_sms);
#define LAUNCH_STREAM(LPV, CREPV)                                                                   \
  do {                                                                                              \
    CU_TRY(h, cudaFuncSetAttribute(adc_stream_kernel<LPV, CREPV, 512>,                       \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));        \
    adc_stream_kernel<LPV, CREPV, 512><<<sgrid, 512, ssmem, h->stream>>>(sa);                 \
  } while (0)
      if (h->LP == 32) LAUNCH_STREAM(32, true); else LAUNCH_STREAM(16, false);
#undef LAUNCH_STREAM
      h->stats.stream_scan_launches++;
      CU_TRY(h, cudaGetLastError());
      h->stats.kernel_launches++;
      h->stats.scan_launches++;
      if (h->profile) {
        CU_TRY(h, cudaEventRecord(h->ev[3], h->stream));
        CU_TRY(h, cudaEventRecord(h->ev[4], h->stream));
      }
      Rank2Args ra{};
      ra.val = h->s_val.as<float>();
      ra.idx = h->s_cand.as<uint32_t>();
      ra.ids = h->d_dbidx.as<uint32_t>() + h->pos_lo;
      ra.QN = QN; ra.max_vec = max_vec; ra.k = k;
      ra.out_dist = fused_out_dist; ra.out_idx = fused_out_idx;
      ra.exact_counter = h->d_exact.as<unsigned long long>();
      ra.tie_counter = h->d_exact.as<unsigned long long>() + 1;
      ra.n_vec = h->s_nvec.as<uint32_t>();
      ra.ridx = h->have_roots ? h->s_ridx.as<uint16_t>() : nullptr;
      ra.fast_rank = (P.rank_mode == 0) ? 1u : 0u;
      if (h->debug) {
        CU_TRY(h, h->g_phases.ensure((size_t)QN * 8 * 8));
        ra.phase_dbg = h->g_phases.as<unsigned long long>();
      }
      CU_TRY(h, h->d_sched.ensure(64));
      CU_TRY(h, cudaMemsetAsync(h->d_sched.as<uint32_t>() + 1, 0, 4, h->stream));
      ra.next_query = h->d_sched.as<uint32_t>() + 1;
      const size_t rsmem = rank2_smem_bytes(max_vec);
      if (rsmem > 48 * 1024)
        CU_TRY(h, cudaFuncSetAttribute(rank2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
      if (ids_by_scan) {
        ra.idx = h->s_ids.as<uint32_t>();
        ra.ids = nullptr;
        if (rsmem > 48 * 1024)
          CU_TRY(h, cudaFuncSetAttribute(rank2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        rank2_kernel<true><<<std::min<uint32_t>(QN, (uint32_t)h->num_sms * 4), kRank2Threads, rsmem, h->stream>>>(ra);
      } else {
        rank2_kernel<false><<<std::min<uint32_t>(QN, (uint32_t)h->num_sms * 4), kRank2Threads, rsmem, h->stream>>>(ra);
      }
      CU_TRY(h, cudaGetLastError());
      h->stats.kernel_launches++;
      if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[5], h->stream));
      h->split_ranked = true;
      if (h->debug && h->world == 1) {
        gather_select_idx_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(
            h->s_cand.as<uint32_t>(), h->s_nvec.as<uint32_t>(), h->d_dbidx.as<uint32_t>(), QN, max_vec,
            h->g_sel.as<uint32_t>());
        CU_TRY(h, cudaGetLastError());
      }
      return PQT_OK;
    }
    if (fused_out_dist && smem_fused <= kSmemMax) {
      RerankArgs g{};
      g.s = a;
      if (four) g.s.cbd_dup = h->d_cbd.as<float>();  // canonical [c1][c1][LP]
      g.k = k;
      g.out_dist = fused_out_dist;
      g.out_idx = fused_out_idx;
      g.exact_counter = h->d_exact.as<unsigned long long>();
      g.tie_counter = h->d_exact.as<unsigned long long>() + 1;
      g.fast_rank = (P.rank_mode == 0) ? 1u : 0u;
      const int env_dedupe = getenv("PQT_RERANK_DEDUPE") ? atoi(getenv("PQT_RERANK_DEDUPE")) : 1;  // A/B runs
      const int env_tpb = getenv("PQT_RERANK_TPB") ? atoi(getenv("PQT_RERANK_TPB")) : 512;
      g.dedupe = env_dedupe ? 1u : 0u;
      CU_TRY(h, h->d_sched.ensure(64));
      CU_TRY(h, cudaMemsetAsync(h->d_sched.p, 0, 4, h->stream));
      g.next_query = h->d_sched.as<uint32_t>();
      if (h->debug) {
        CU_TRY(h, h->g_phases.ensure((size_t)QN * 8 * 8));
        g.phase_dbg = h->g_phases.as<unsigned long long>();
      }
      const uint32_t ng = four ? 4u : 2u;
      uint32_t grid = std::min<uint32_t>((QN + ng - 1) / ng, (uint32_t)h->num_sms);
#define LAUNCH_RERANK(LPV, NGV, CREPV)                                                           \
  do {                                                                                           \
    CU_TRY(h, cudaFuncSetAttribute(rerank_kernel<LPV, NGV, CREPV>,                               \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fused)); \
    rerank_kernel<LPV, NGV, CREPV><<<grid, kScanThreads, smem_fused, h->stream>>>(g);            \
  } while (0)
      // lineparts = 32: two groups of 256 threads with 128 registers each (pipelined scan)
#define LAUNCH_RERANK_512()                                                                      \
  do {                                                                                           \
    CU_TRY(h, cudaFuncSetAttribute(rerank_kernel<32, 2, true, 512>,                              \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fused)); \
    rerank_kernel<32, 2, true, 512><<<grid, 512, smem_fused, h->stream>>>(g);                    \
  } while (0)
      if (four) {
        switch (h->LP) {
          case 1: LAUNCH_RERANK(1, 4, false); break;
          case 2: LAUNCH_RERANK(2, 4, false); break;
          case 4: LAUNCH_RERANK(4, 4, false); break;
          case 8: LAUNCH_RERANK(8, 4, false); break;
          default: LAUNCH_RERANK(16, 4, false); break;
        }
      } else {
        switch (h->LP) {
          case 1: LAUNCH_RERANK(1, 2, true); break;
          case 2: LAUNCH_RERANK(2, 2, true); break;
          case 4: LAUNCH_RERANK(4, 2, true); break;
          case 8: LAUNCH_RERANK(8, 2, true); break;
          case 16: LAUNCH_RERANK(16, 2, true); break;
          default:
            if (env_tpb == 512) LAUNCH_RERANK_512(); else LAUNCH_RERANK(32, 2, true);
            break;
        }
      }
#undef LAUNCH_RERANK_512
#undef LAUNCH_RERANK
    } else if (smem <= 220 * 1024) {
      uint32_t grid = std::min<uint32_t>(QN, (uint32_t)h->num_sms);
#define LAUNCH_SCAN(LPV)                                                                         \
  do {                                                                                           \
    CU_TRY(h, cudaFuncSetAttribute(adc_scan_kernel<LPV>,                                         \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    adc_scan_kernel<LPV><<<grid, kScanThreads, smem, h->stream>>>(a);                            \
  } while (0)
      switch (h->LP) {
        case 1: LAUNCH_SCAN(1); break;
        case 2: LAUNCH_SCAN(2); break;
        case 4: LAUNCH_SCAN(4); break;
        case 8: LAUNCH_SCAN(8); break;
        case 16: LAUNCH_SCAN(16); break;
        default: LAUNCH_SCAN(32); break;
      }
#undef LAUNCH_SCAN
    } else {
      // c1 > 32: tables do not fit in shared memory; canonical-layout fallback
      if (!h->debug) CU_TRY(h, h->g_lut.ensure((size_t)QN * h->LP * h->c1 * 4));
      return fail(h, PQT_ERR_INVALID, "c1 = %u > 32 is not supported by the ADC scan yet", h->c1);
    }
    CU_TRY(h, cudaGetLastError());
    h->stats.kernel_launches++;
    h->stats.scan_launches++;
  }
  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[3], h->stream));
  if (h->debug && h->world == 1) {
    gather_select_idx_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(
        h->s_cand.as<uint32_t>(), h->s_nvec.as<uint32_t>(), h->d_dbidx.as<uint32_t>(), QN, max_vec,
        h->g_sel.as<uint32_t>());
    CU_TRY(h, cudaGetLastError());
  }
  return PQT_OK;
}

int run_rank(pqt_index* h, const float* d_val, const uint32_t* d_idx, uint32_t QN, uint32_t max_vec,
             uint32_t k, float* d_out_dist, uint32_t* d_out_idx) {
  if (!is_pow2(max_vec) || max_vec > 4096)
    return fail(h, PQT_ERR_INVALID, "candidate width %u must be a power of two <= 4096", max_vec);
  if (k > max_vec) return fail(h, PQT_ERR_INVALID, "k %u > candidate width %u", k, max_vec);
  Rank2Args a{};
  a.val = d_val; a.idx = d_idx; a.QN = QN; a.max_vec = max_vec; a.k = k;
  a.out_dist = d_out_dist; a.out_idx = d_out_idx;
  a.exact_counter = h->d_exact.as<unsigned long long>();
  a.fast_rank = (h->prm.rank_mode == 0) ? 1u : 0u;
  size_t smem = rank2_smem_bytes(max_vec);
  if (smem > 48 * 1024)
    CU_TRY(h, cudaFuncSetAttribute(rank2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  uint32_t grid = std::min<uint32_t>(QN, (uint32_t)h->num_sms * 4);
  rank2_kernel<true><<<grid, kRank2Threads, smem, h->stream>>>(a);
  CU_TRY(h, cudaGetLastError());
  h->stats.kernel_launches++;
  return PQT_OK;
}

int check_query_state(const pqt_index* h, uint32_t QN, uint32_t k) {
  if (!h->c1) return fail(h, PQT_ERR_STATE, "no tree loaded (pqt_read_tree / pqt_set_tree)");
  if (!h->has_db) return fail(h, PQT_ERR_STATE, "no DB loaded (pqt_set_db / pqt_build_kbest_db)");
  if (!h->has_lines)
    return fail(h, PQT_ERR_STATE,
                "no line codes loaded (pqt_set_lines / pqt_line_dist); the reference would "
                "divide by d_lineParts = 0 here (pqt/PerturbationProTree.cu:5281)");
  if (!QN || !k) return fail(h, PQT_ERR_INVALID, "QN and k must be non-zero");
  const pqt_params& P = h->prm;
  if (P.k1 > h->c1) return fail(h, PQT_ERR_INVALID, "k1 %u > c1 %u", P.k1, h->c1);
  if ((uint64_t)P.max_trials * P.bin_threads > kNumDistSeq)
    return fail(h, PQT_ERR_INVALID, "max_trials * bin_threads exceeds the %u traversal codes", kNumDistSeq);
  if (P.bin_threads > 16u * kBinsThreads) return fail(h, PQT_ERR_INVALID, "bin_threads > %u", 16 * kBinsThreads);
  if (pow2ceil(P.k1 * h->c2) > 1024) return fail(h, PQT_ERR_INVALID, "k1*c2 > 1024");
  if (P.max_bins < 2 || P.max_bins > 8192) return fail(h, PQT_ERR_INVALID, "max_bins out of range");
  uint32_t mv = candidate_width(h, k);
  if (!is_pow2(mv) || mv > 4096 || mv < k)
    return fail(h, PQT_ERR_INVALID,
                "candidate width %u (pow2ceil(k) or params.max_vec) must be a power of two in "
                "[k, 4096]",
                mv);
  if (P.hash_size != h->db_hash_size) return fail(h, PQT_ERR_STATE, "hash_size changed after the DB was set");
  return PQT_OK;
}

void accumulate_profile(pqt_index* h, uint32_t QN, bool with_rank) {
  float t = 0;
  cudaEventElapsedTime(&t, h->ev[0], h->ev[1]);
  h->stats.ms_tables += t;
  cudaEventElapsedTime(&t, h->ev[1], h->ev[2]);
  h->stats.ms_bins += t;
  cudaEventElapsedTime(&t, h->ev[2], h->ev[3]);
  h->stats.ms_scan += t;
  if (with_rank) {
    cudaEventElapsedTime(&t, h->ev[4], h->ev[5]);
    h->stats.ms_sort += t;
    cudaEventElapsedTime(&t, h->ev[0], h->ev[5]);
  } else {
    cudaEventElapsedTime(&t, h->ev[0], h->ev[3]);
  }
  h->stats.ms_total += t;
  h->stats.calls++;
  h->stats.queries += QN;
  std::vector<uint32_t> nv(QN);
  cudaMemcpy(nv.data(), h->s_nvec.p, (size_t)QN * 4, cudaMemcpyDeviceToHost);
  uint64_t s = 0;
  for (uint32_t v : nv) s += v;
  h->stats.candidates += s;
}

}  // namespace

// =====================================================================================
extern "C" {

int pqt_abi_version(void) { return PQT_ABI_VERSION; }

void pqt_default_params(pqt_params* prm) {
  std::memset(prm, 0, sizeof(*prm));
  prm->k1 = 8;
  prm->max_bins = 4096;
  prm->max_trials = 16;
  prm->bin_threads = 1024;
  prm->max_vec_per_bin = 2800;
  prm->hash_size = 400000000u;
  prm->k1_build = 16;
  prm->max_vec = 0;
  prm->big_k1 = 16;
  prm->big_max_bins = 64 * 8192;
  prm->big_max_trials = 2560;
}

int pqt_create(uint32_t dim, uint32_t p, uint32_t p2, int device, pqt_index** out) {
  if (!out) return PQT_ERR_INVALID;
  *out = nullptr;
  if (p2 != p || !dim || !p) return PQT_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return PQT_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= ndev) return PQT_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return PQT_ERR_CUDA;
  pqt_index* h = new pqt_index();
  h->device = device;
  h->dim = dim;
  h->p = p;
  pqt_default_params(&h->prm);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return PQT_ERR_CUDA;
  }
  h->stream = h->own_stream;
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaStreamDestroy(h->own_stream);
    delete h;
    return PQT_ERR_CUDA;
  }
  for (auto& e : h->ev) cudaEventCreate(&e);
  cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking);
  for (auto& e : h->aux_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  if (h->d_exact.ensure(16) != cudaSuccess || cudaMemset(h->d_exact.p, 0, 16) != cudaSuccess) {
    pqt_destroy(h);
    return PQT_ERR_CUDA;
  }
  *out = h;
  return PQT_OK;
}

int pqt_destroy(pqt_index* h) {
  if (!h) return PQT_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (DevBuf* b : {&h->d_cb1, &h->d_cb2, &h->d_cb1T, &h->d_cb2T, &h->d_distseq, &h->d_seqnib, &h->d_seqsorted, &h->d_seqfine, &h->d_seqmega, &h->d_seqplain, &h->d_seq2d, &h->s_topv, &h->s_topi, &h->g_bigbins, &h->g_bignbins, &h->d_bitmap, &h->d_rank_base, &h->d_cprefix,
                    &h->d_dbidx, &h->d_codes, &h->d_cbd, &h->d_cbd_dup, &h->s_q, &h->s_lut, &h->s_idx16,
                    &h->s_cand, &h->s_nvec, &h->s_val, &h->s_idx, &h->s_outd, &h->s_outi, &h->g_assign,
                    &h->g_lut, &h->g_aval, &h->g_aidx, &h->g_bins, &h->g_nbins, &h->g_sel, &h->d_exact, &h->x_val, &h->x_inbox, &h->x_cnt})
    b->release();
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  for (uint32_t r = 0; r < 8; r++) {
    if (h->x_ipc_opened[r]) {
      cudaIpcCloseMemHandle(h->x_peer_val[r]);
      cudaIpcCloseMemHandle(h->x_peer_inbox[r]);
      cudaIpcCloseMemHandle(h->x_peer_cnt[r]);
    }
  }
  for (auto& e : h->slab_ev) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  for (auto& e : h->aux_ev)
    if (e) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return PQT_OK;
}

const char* pqt_last_error(const pqt_index* h) { return h ? h->err.c_str() : "null handle"; }

int pqt_set_params(pqt_index* h, const pqt_params* prm) {
  if (!h || !prm) return PQT_ERR_INVALID;
  if (!prm->k1 || !prm->max_bins || !prm->max_trials || !prm->bin_threads || !prm->hash_size ||
      !prm->k1_build)
    return fail(h, PQT_ERR_INVALID, "zero parameter");
  if (prm->max_vec && !is_pow2(prm->max_vec)) return fail(h, PQT_ERR_INVALID, "max_vec must be a power of two");
  if (prm->rank_mode > 1) return fail(h, PQT_ERR_INVALID, "rank_mode must be 0 or 1");
  h->prm = *prm;
  return PQT_OK;
}

int pqt_get_params(const pqt_index* h, pqt_params* prm) {
  if (!h || !prm) return PQT_ERR_INVALID;
  *prm = h->prm;
  return PQT_OK;
}

int pqt_set_stream(pqt_index* h, void* cuda_stream) {
  if (!h) return PQT_ERR_INVALID;
  cudaStreamSynchronize(h->stream);
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return PQT_OK;
}

// ---- tree ---------------------------------------------------------------------------
int pqt_set_tree(pqt_index* h, uint32_t c1, uint32_t c2, const float* cb1, const float* cb2) {
  if (!h || !cb1 || !cb2) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  PQ_TRY(check_tree_shape(h, h->dim, h->p, c1, c2));
  h->c1 = c1;
  h->c2 = c2;
  h->vl = h->dim / h->p;
  h->h_cb1.assign(cb1, cb1 + (size_t)c1 * h->dim);
  h->h_cb2.assign(cb2, cb2 + (size_t)c1 * c2 * h->dim);
  return upload_tree(h);
}

int pqt_read_tree(pqt_index* h, const char* path) {
  if (!h || !path) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  std::ifstream f(path, std::ios::in | std::ios::binary);
  if (!f.good()) return fail(h, PQT_ERR_IO, "cannot open %s", path);
  uint32_t dim = 0, p = 0, p2 = 0, c1 = 0, c2 = 0, ndb = 0;
  f >> dim >> p >> p2 >> c1 >> c2 >> ndb;
  if (!f.good()) return fail(h, PQT_ERR_IO, "%s: bad .ppqt header", path);
  f.ignore(1);
  if (ndb != 1) return fail(h, PQT_ERR_INVALID, "%s: nDBs = %u (only 1 is supported: perturbations were removed from the reference)", path, ndb);
  if (p2 != p) return fail(h, PQT_ERR_INVALID, "%s: p2 != p", path);
  PQ_TRY(check_tree_shape(h, dim, p, c1, c2));
  std::vector<float> cb1((size_t)c1 * dim), cb2((size_t)c1 * c2 * dim);
  f.read(reinterpret_cast<char*>(cb1.data()), cb1.size() * sizeof(float));
  f.read(reinterpret_cast<char*>(cb2.data()), cb2.size() * sizeof(float));
  if (!f.good() && !f.eof()) return fail(h, PQT_ERR_IO, "%s: read error", path);
  if ((size_t)f.gcount() != cb2.size() * sizeof(float))
    return fail(h, PQT_ERR_IO, "%s: truncated codebooks (the reference would silently read garbage)", path);
  h->dim = dim;  // the file overrides the constructor arguments (:120-125)
  h->p = p;
  h->c1 = c1;
  h->c2 = c2;
  h->vl = dim / p;
  h->h_cb1.swap(cb1);
  h->h_cb2.swap(cb2);
  return upload_tree(h);
}

int pqt_write_tree(pqt_index* h, const char* path) {
  if (!h || !path) return PQT_ERR_INVALID;
  if (!h->c1) return fail(h, PQT_ERR_STATE, "no tree to write");
  std::ofstream f(path, std::ios::out | std::ios::binary);
  if (!f.good()) return fail(h, PQT_ERR_IO, "cannot open %s for writing", path);
  f << h->dim << std::endl << h->p << std::endl << h->p << std::endl;
  f << h->c1 << std::endl << h->c2 << std::endl << 1 << std::endl;
  f.write(reinterpret_cast<const char*>(h->h_cb1.data()), h->h_cb1.size() * sizeof(float));
  f.write(reinterpret_cast<const char*>(h->h_cb2.data()), h->h_cb2.size() * sizeof(float));
  f.close();
  if (!f.good()) return fail(h, PQT_ERR_IO, "write error on %s", path);
  return PQT_OK;
}

int pqt_get_tree_shape(const pqt_index* h, uint32_t* dim, uint32_t* p, uint32_t* c1, uint32_t* c2) {
  if (!h) return PQT_ERR_INVALID;
  if (dim) *dim = h->dim;
  if (p) *p = h->p;
  if (c1) *c1 = h->c1;
  if (c2) *c2 = h->c2;
  return PQT_OK;
}

int pqt_get_tree(const pqt_index* h, float* cb1, float* cb2) {
  if (!h) return PQT_ERR_INVALID;
  if (!h->c1) return fail(h, PQT_ERR_STATE, "no tree loaded");
  if (cb1) std::memcpy(cb1, h->h_cb1.data(), h->h_cb1.size() * sizeof(float));
  if (cb2) std::memcpy(cb2, h->h_cb2.data(), h->h_cb2.size() * sizeof(float));
  return PQT_OK;
}

// ---- DB -----------------------------------------------------------------------------
int pqt_set_shard(pqt_index* h, uint32_t rank, uint32_t world) {
  if (!h || !world || rank >= world) return PQT_ERR_INVALID;
  if (h->has_db) {
    CU_TRY(h, cudaSetDevice(h->device));
    const uint32_t lo = (uint32_t)((uint64_t)h->N * rank / world);
    const uint32_t hi = (uint32_t)((uint64_t)h->N * (rank + 1) / world);
    if (h->has_lines) {
      // trim the resident codes to the new slice (it must lie inside the current one)
      if (lo < h->pos_lo || hi > h->pos_hi)
        return fail(h, PQT_ERR_STATE, "pqt_set_shard can only narrow the resident slice; reload the line codes");
      DevBuf nb;
      const size_t bytes = (size_t)(hi - lo) * h->LP * 4;
      CU_TRY(h, nb.ensure(std::max<size_t>(bytes, 16)));
      CU_TRY(h, cudaMemcpyAsync(nb.p, h->d_codes.as<uint32_t>() + (size_t)(lo - h->pos_lo) * h->LP, bytes,
                                cudaMemcpyDeviceToDevice, h->stream));
      CU_TRY(h, cudaStreamSynchronize(h->stream));
      h->d_codes = std::move(nb);
    }
    h->pos_lo = lo;
    h->pos_hi = hi;
  }
  h->rank = rank;
  h->world = world;
  return PQT_OK;
}

int pqt_set_db(pqt_index* h, uint32_t N, const uint32_t* prefix, const uint32_t* counts,
               const uint32_t* db_idx) {
  if (!h || !prefix || !counts || !db_idx || !N) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  const uint32_t hs = h->prm.hash_size;
  // the lists must stay inside the N ids (a .dbIdx shorter than what .prefix / .count describe
  // would make the kernels read and write out of bounds)
  if ((uint64_t)prefix[hs - 1] + counts[hs - 1] > N)
    return fail(h, PQT_ERR_INVALID, "prefix/counts cover %llu vectors, N is only %u (hash_size %u)",
                (unsigned long long)prefix[hs - 1] + counts[hs - 1], N, hs);
  DevBuf d_counts, d_prefix;
  CU_TRY(h, d_counts.ensure((size_t)hs * 4));
  CU_TRY(h, d_prefix.ensure((size_t)hs * 4));
  // chunked H2D like the reference (:1212-1222); host memory may be pageable
  const size_t chunk = 100000000;
  for (size_t off = 0; off < hs; off += chunk) {
    size_t n = std::min(chunk, (size_t)hs - off);
    CU_TRY(h, cudaMemcpyAsync(d_counts.as<uint32_t>() + off, counts + off, n * 4, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(h, cudaMemcpyAsync(d_prefix.as<uint32_t>() + off, prefix + off, n * 4, cudaMemcpyHostToDevice, h->stream));
  }
  CU_TRY(h, h->d_dbidx.ensure((size_t)N * 4));
  CU_TRY(h, cudaMemcpyAsync(h->d_dbidx.p, db_idx, (size_t)N * 4, cudaMemcpyHostToDevice, h->stream));
  {
    DevBuf flag;
    CU_TRY(h, flag.ensure(4));
    CU_TRY(h, cudaMemsetAsync(flag.p, 0, 4, h->stream));
    check_below_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_dbidx.as<uint32_t>(), N, N, flag.as<uint32_t>());
    check_lists_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(d_counts.as<uint32_t>(), d_prefix.as<uint32_t>(), hs, N,
                                                             flag.as<uint32_t>());
    uint32_t bad = 0;
    CU_TRY(h, cudaMemcpyAsync(&bad, flag.p, 4, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (bad) return fail(h, PQT_ERR_INVALID, "dbIdx holds ids >= N = %u, or a bin's list leaves the N ids", N);
  }
  return build_directory(h, d_counts.as<uint32_t>(), d_prefix.as<uint32_t>(), hs, N);
}

int pqt_set_lines(pqt_index* h, const uint32_t* lines, uint32_t N, uint32_t line_parts) {
  if (!h || !lines) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->c1) return fail(h, PQT_ERR_STATE, "pqt_set_lines needs the tree (cbDist is derived from cb1)");
  if (!h->has_db) return fail(h, PQT_ERR_STATE, "pqt_set_lines must follow pqt_set_db (codes are stored in bin order)");
  if (N != h->N) return fail(h, PQT_ERR_INVALID, "N = %u differs from the DB's %u", N, h->N);
  PQ_TRY(check_lp(h, line_parts));
  const uint32_t LP = line_parts;
  h->LP = LP;
  h->sl = h->dim / LP;
  PQ_TRY(compute_cbd(h, LP));
  const uint32_t n_local = h->pos_hi - h->pos_lo;
  CU_TRY(h, h->d_codes.ensure(std::max<size_t>((size_t)n_local * LP * 4, 16)));
  DevBuf inv, stage;
  CU_TRY(h, inv.ensure((size_t)N * 4));
  invert_perm_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_dbidx.as<uint32_t>(), N, inv.as<uint32_t>());
  const uint32_t chunk = 4u << 20;
  CU_TRY(h, stage.ensure((size_t)std::min(chunk, N) * LP * 4));
  DevBuf flag;
  CU_TRY(h, flag.ensure(4));
  CU_TRY(h, cudaMemsetAsync(flag.p, 0, 4, h->stream));
  for (uint32_t id0 = 0; id0 < N; id0 += chunk) {
    uint32_t n = std::min(chunk, N - id0);
    CU_TRY(h, cudaMemcpyAsync(stage.p, lines + (size_t)id0 * LP, (size_t)n * LP * 4, cudaMemcpyHostToDevice, h->stream));
    // p1 / p2 index the shared-memory tables of the scan: they must be centroid numbers
    check_codes_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(stage.as<uint32_t>(), (size_t)n * LP, h->c1, flag.as<uint32_t>());
    scatter_codes_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(
        stage.as<uint32_t>(), id0, n, LP, inv.as<uint32_t>(), h->pos_lo, h->pos_hi, h->d_codes.as<uint32_t>());
    CU_TRY(h, cudaStreamSynchronize(h->stream));  // staging buffer reuse
  }
  CU_TRY(h, cudaGetLastError());
  uint32_t bad = 0;
  CU_TRY(h, cudaMemcpy(&bad, flag.p, 4, cudaMemcpyDeviceToHost));
  if (bad) return fail(h, PQT_ERR_INVALID, "line codes hold centroid numbers >= c1 = %u", h->c1);
  h->has_lines = true;
  return PQT_OK;
}

// ---- compact index file: the resident layout as it is ----------------------------------
//   The reference's files keep two dense hash_size arrays (.prefix / .count,
// 3.2 GB at HASH_SIZE 4e8) and the line codes by vector id; loading them means rebuilding the
// directory and re-ordering 128 GB of codes.  This file holds what the handle holds: header,
// then bitmap | rank_base | cprefix | dbIdx (bin order) | codes (bin order, this handle's
// slice), each as u64 byte count + bytes.
namespace {
struct IndexFileHeader {
  char magic[8];  // "PQTB2IDX"
  uint32_t version, dim, p, c1, c2, N, hash_size, n_nonempty, LP, rank, world, pos_lo, pos_hi, reserved;
};
static_assert(sizeof(IndexFileHeader) == 64, "header layout");
constexpr size_t kFileChunk = (size_t)64 << 20;

struct HostStage {
  void* p = nullptr;
  ~HostStage() {
    if (p) cudaFreeHost(p);
  }
};
struct FileCloser {
  FILE* f = nullptr;
  ~FileCloser() {
    if (f) fclose(f);
  }
};

int write_section(const pqt_index* h, FILE* f, const void* dev, size_t bytes, void* stage) {
  const uint64_t n64 = bytes;
  if (fwrite(&n64, 8, 1, f) != 1) return fail(h, PQT_ERR_IO, "short write");
  for (size_t off = 0; off < bytes; off += kFileChunk) {
    const size_t n = std::min(kFileChunk, bytes - off);
    CU_TRY(h, cudaMemcpy(stage, static_cast<const char*>(dev) + off, n, cudaMemcpyDeviceToHost));
    if (fwrite(stage, 1, n, f) != n) return fail(h, PQT_ERR_IO, "short write");
  }
  return PQT_OK;
}

int read_section(pqt_index* h, FILE* f, DevBuf& dst, size_t want_bytes, size_t alloc_bytes, void* stage,
                 const char* what) {
  uint64_t n64 = 0;
  if (fread(&n64, 8, 1, f) != 1 || n64 != want_bytes)
    return fail(h, PQT_ERR_IO, "index file: section '%s' holds %llu bytes, expected %llu", what,
                (unsigned long long)n64, (unsigned long long)want_bytes);
  CU_TRY(h, dst.ensure(std::max<size_t>(alloc_bytes, 16)));
  for (size_t off = 0; off < want_bytes; off += kFileChunk) {
    const size_t n = std::min(kFileChunk, want_bytes - off);
    if (fread(stage, 1, n, f) != n) return fail(h, PQT_ERR_IO, "index file: section '%s' is truncated", what);
    CU_TRY(h, cudaMemcpy(static_cast<char*>(dst.p) + off, stage, n, cudaMemcpyHostToDevice));
  }
  return PQT_OK;
}
}  // namespace

int pqt_save_index(const pqt_index* h, const char* path) {
  if (!h || !path) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->has_db || !h->has_lines) return fail(h, PQT_ERR_STATE, "pqt_save_index needs the DB and the line codes");
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  FileCloser fc;
  fc.f = fopen(path, "wb");
  if (!fc.f) return fail(h, PQT_ERR_IO, "cannot open %s for writing", path);
  HostStage st;
  CU_TRY(h, cudaMallocHost(&st.p, kFileChunk));
  IndexFileHeader hd{};
  memcpy(hd.magic, "PQTB2IDX", 8);
  hd.version = 1;
  hd.dim = h->dim; hd.p = h->p; hd.c1 = h->c1; hd.c2 = h->c2;
  hd.N = h->N; hd.hash_size = h->db_hash_size; hd.n_nonempty = h->n_nonempty; hd.LP = h->LP;
  hd.rank = h->rank; hd.world = h->world; hd.pos_lo = h->pos_lo; hd.pos_hi = h->pos_hi;
  if (fwrite(&hd, sizeof(hd), 1, fc.f) != 1) return fail(h, PQT_ERR_IO, "short write");
  const size_t nwords = ((size_t)h->db_hash_size + 31) >> 5, ngroups = (nwords + 7) >> 3;
  PQ_TRY(write_section(h, fc.f, h->d_bitmap.p, ngroups * 8 * 4, st.p));
  PQ_TRY(write_section(h, fc.f, h->d_rank_base.p, (ngroups + 1) * 4, st.p));
  PQ_TRY(write_section(h, fc.f, h->d_cprefix.p, ((size_t)h->n_nonempty + 1) * 4, st.p));
  PQ_TRY(write_section(h, fc.f, h->d_dbidx.p, (size_t)h->N * 4, st.p));
  PQ_TRY(write_section(h, fc.f, h->d_codes.p, (size_t)(h->pos_hi - h->pos_lo) * h->LP * 4, st.p));
  if (fflush(fc.f) != 0) return fail(h, PQT_ERR_IO, "short write");
  return PQT_OK;
}

int pqt_load_index(pqt_index* h, const char* path) {
  if (!h || !path) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->c1) return fail(h, PQT_ERR_STATE, "pqt_load_index needs the tree (pqt_read_tree / pqt_set_tree)");
  FileCloser fc;
  fc.f = fopen(path, "rb");
  if (!fc.f) return fail(h, PQT_ERR_IO, "cannot open %s", path);
  IndexFileHeader hd{};
  if (fread(&hd, sizeof(hd), 1, fc.f) != 1 || memcmp(hd.magic, "PQTB2IDX", 8) != 0 || hd.version != 1)
    return fail(h, PQT_ERR_IO, "%s is not a pqt_b200 index file (version 1)", path);
  if (hd.dim != h->dim || hd.p != h->p || hd.c1 != h->c1 || hd.c2 != h->c2)
    return fail(h, PQT_ERR_INVALID, "index file was built for a %u-d tree with p=%u c1=%u c2=%u", hd.dim, hd.p, hd.c1, hd.c2);
  if (hd.hash_size != h->prm.hash_size)
    return fail(h, PQT_ERR_INVALID, "index file uses hash_size %u, the handle %u (pqt_set_params)", hd.hash_size, h->prm.hash_size);
  if (hd.rank != h->rank || hd.world != h->world)
    return fail(h, PQT_ERR_INVALID, "index file holds shard %u of %u, the handle is shard %u of %u (pqt_set_shard)",
                hd.rank, hd.world, h->rank, h->world);
  const uint32_t lo = (uint32_t)((uint64_t)hd.N * hd.rank / hd.world), hi = (uint32_t)((uint64_t)hd.N * (hd.rank + 1) / hd.world);
  if (!hd.hash_size || hd.pos_lo != lo || hd.pos_hi != hi || hd.n_nonempty > hd.hash_size || hd.n_nonempty > hd.N)
    return fail(h, PQT_ERR_IO, "index file header is inconsistent");
  PQ_TRY(check_lp(h, hd.LP));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->has_db = false;
  h->has_lines = false;
  h->line_build = false;  // a chunked build in progress is abandoned
  h->b_inv.release();
  HostStage st;
  CU_TRY(h, cudaMallocHost(&st.p, kFileChunk));
  const size_t nwords = ((size_t)hd.hash_size + 31) >> 5, ngroups = (nwords + 7) >> 3;
  PQ_TRY(read_section(h, fc.f, h->d_bitmap, ngroups * 8 * 4, ngroups * 8 * 4, st.p, "bitmap"));
  PQ_TRY(read_section(h, fc.f, h->d_rank_base, (ngroups + 1) * 4, (ngroups + 1) * 4, st.p, "rank_base"));
  PQ_TRY(read_section(h, fc.f, h->d_cprefix, ((size_t)hd.n_nonempty + 1) * 4, ((size_t)hd.n_nonempty + 2) * 4, st.p, "cprefix"));
  PQ_TRY(read_section(h, fc.f, h->d_dbidx, (size_t)hd.N * 4, (size_t)hd.N * 4, st.p, "dbIdx"));
  const size_t code_words = (size_t)(hi - lo) * hd.LP;
  PQ_TRY(read_section(h, fc.f, h->d_codes, code_words * 4, code_words * 4, st.p, "codes"));
  // what a file can get wrong and a kernel would trip over: ids, centroid numbers, the list end
  DevBuf flag;
  CU_TRY(h, flag.ensure(8));
  CU_TRY(h, cudaMemsetAsync(flag.p, 0, 8, h->stream));
  if (hd.N) check_below_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_dbidx.as<uint32_t>(), hd.N, hd.N, flag.as<uint32_t>());
  if (code_words) check_codes_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_codes.as<uint32_t>(), code_words, h->c1, flag.as<uint32_t>() + 1);
  check_below_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_cprefix.as<uint32_t>(), (size_t)hd.n_nonempty + 1, hd.N + 1, flag.as<uint32_t>());
  check_monotone_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_cprefix.as<uint32_t>(), (size_t)hd.n_nonempty + 1, flag.as<uint32_t>());
  CU_TRY(h, cudaGetLastError());
  uint32_t bad[2] = {0, 0}, last = 0;
  CU_TRY(h, cudaMemcpyAsync(bad, flag.p, 8, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaMemcpyAsync(&last, h->d_cprefix.as<uint32_t>() + hd.n_nonempty, 4, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  if (bad[0] || bad[1] || last != hd.N)
    return fail(h, PQT_ERR_INVALID, "index file holds ids >= N, centroid numbers >= c1 or bin lists that leave the N ids");
  h->N = hd.N;
  h->db_hash_size = hd.hash_size;
  h->n_nonempty = hd.n_nonempty;
  h->pos_lo = lo;
  h->pos_hi = hi;
  h->has_db = true;
  h->LP = hd.LP;
  h->sl = h->dim / hd.LP;
  PQ_TRY(compute_cbd(h, hd.LP));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->has_lines = true;
  return PQT_OK;
}

int pqt_get_db_size(const pqt_index* h, uint32_t* N, uint32_t* line_parts) {
  if (!h) return PQT_ERR_INVALID;
  if (N) *N = h->has_db ? h->N : 0;
  if (line_parts) *line_parts = h->has_lines ? h->LP : 0;
  return PQT_OK;
}

int pqt_get_db(const pqt_index* hc, uint32_t* prefix, uint32_t* counts, uint32_t* db_idx) {
  pqt_index* h = const_cast<pqt_index*>(hc);
  if (!h) return PQT_ERR_INVALID;
  if (!h->has_db) return fail(h, PQT_ERR_STATE, "no DB");
  CU_TRY(h, cudaSetDevice(h->device));
  if (db_idx) CU_TRY(h, cudaMemcpy(db_idx, h->d_dbidx.p, (size_t)h->N * 4, cudaMemcpyDeviceToHost));
  if (prefix || counts) {
    const uint32_t hs = h->db_hash_size;
    DevBuf dc, dp;
    CU_TRY(h, dc.ensure((size_t)hs * 4));
    CU_TRY(h, dp.ensure((size_t)hs * 4));
    BinDir d{h->d_bitmap.as<uint32_t>(), h->d_rank_base.as<uint32_t>(), h->d_cprefix.as<uint32_t>()};
    expand_directory_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(d, hs, dc.as<uint32_t>(), dp.as<uint32_t>());
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (counts) CU_TRY(h, cudaMemcpy(counts, dc.p, (size_t)hs * 4, cudaMemcpyDeviceToHost));
    if (prefix) CU_TRY(h, cudaMemcpy(prefix, dp.p, (size_t)hs * 4, cudaMemcpyDeviceToHost));
    dc.release();
    dp.release();
  }
  return PQT_OK;
}

int pqt_get_lines(const pqt_index* hc, uint32_t* lines) {
  pqt_index* h = const_cast<pqt_index*>(hc);
  if (!h || !lines) return PQT_ERR_INVALID;
  if (!h->has_lines) return fail(h, PQT_ERR_STATE, "no line codes");
  if (h->world != 1) return fail(h, PQT_ERR_STATE, "pqt_get_lines on a sharded handle");
  CU_TRY(h, cudaSetDevice(h->device));
  DevBuf out;
  CU_TRY(h, out.ensure((size_t)h->N * h->LP * 4));
  gather_codes_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_codes.as<uint32_t>(), h->d_dbidx.as<uint32_t>(),
                                                             h->N, h->LP, out.as<uint32_t>());
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaMemcpy(lines, out.p, (size_t)h->N * h->LP * 4, cudaMemcpyDeviceToHost));
  out.release();
  return PQT_OK;
}

}  // extern "C"

int pqt_get_codes_binorder(const pqt_index* hc, uint64_t pos0, uint64_t n, uint32_t* codes) {
  pqt_index* h = const_cast<pqt_index*>(hc);
  if (!h || !codes) return PQT_ERR_INVALID;
  if (!h->has_lines) return fail(h, PQT_ERR_STATE, "no line codes");
  if (pos0 + n > (uint64_t)(h->pos_hi - h->pos_lo)) return fail(h, PQT_ERR_INVALID, "rows exceed the resident slice");
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaMemcpy(codes, h->d_codes.as<uint32_t>() + (size_t)pos0 * h->LP, (size_t)n * h->LP * 4,
                       cudaMemcpyDeviceToHost));
  return PQT_OK;
}

// ---- build side -----------------------------------------------------------------------
namespace {


__global__ void u8_to_f32_kernel(const uint8_t* in, size_t n, float* out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (float)in[i];
}

size_t x_elem_bytes(int x_kind) { return x_kind == PQT_X_U8 ? 1 : 4; }

// rows of a chunk on the device: the caller's pointer, or an upload into h->b_x
int stage_rows(pqt_index* h, const void* X, int x_kind, int x_on_device, uint32_t n, const void** out) {
  if (x_kind != PQT_X_F32 && x_kind != PQT_X_U8) return fail(h, PQT_ERR_INVALID, "x_kind must be PQT_X_F32 or PQT_X_U8");
  if (x_on_device) {
    *out = X;
    return PQT_OK;
  }
  const size_t bytes = (size_t)n * h->dim * x_elem_bytes(x_kind);
  CU_TRY(h, h->b_x.ensure(bytes));
  CU_TRY(h, cudaMemcpyAsync(h->b_x.p, X, bytes, cudaMemcpyHostToDevice, h->stream));
  *out = h->b_x.p;
  return PQT_OK;
}

// float view of device rows for the generic (any-shape) kernels
int rows_as_f32(pqt_index* h, const void* dX, int x_kind, uint32_t n, DevBuf& tmp, const float** out) {
  if (x_kind == PQT_X_F32) {
    *out = static_cast<const float*>(dX);
    return PQT_OK;
  }
  const size_t ne = (size_t)n * h->dim;
  CU_TRY(h, tmp.ensure(ne * 4));
  u8_to_f32_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(static_cast<const uint8_t*>(dX), ne, tmp.as<float>());
  CU_TRY(h, cudaGetLastError());
  *out = tmp.as<float>();
  return PQT_OK;
}

uint32_t build_grid(const pqt_index* h) { return ((uint32_t)h->num_sms * 2u) & ~3u; }

// bins of n device rows -> d_bin (device)
int launch_assign_bins(pqt_index* h, const void* dX, int x_kind, uint32_t n, uint32_t* d_bin) {
  const pqt_params& P = h->prm;
  if (P.k1_build > h->c1) return fail(h, PQT_ERR_INVALID, "k1_build %u > c1 %u (the reference needs c1 >= 16)", P.k1_build, h->c1);
  const bool fast = h->c1 <= 32 && h->c2 <= 32 && (h->vl == 8 || h->vl == 16 || h->vl == 32) &&
                    (h->dim % 16) == 0 && ((uintptr_t)dX & 15u) == 0;
  if (fast) {
    AssignWarpArgs a{};
    a.X = dX; a.n = n; a.dim = h->dim; a.p = h->p; a.c1 = h->c1; a.c2 = h->c2; a.k1 = P.k1_build;
    a.cb1T = h->d_cb1T.as<float>(); a.cb2T = h->d_cb2T.as<float>();
    a.hash = make_fastmod(P.hash_size);
    a.bin_of = d_bin;
    const uint32_t grid = build_grid(h), thr = kBuildWarpsPerCta * 32;
    const bool cc32 = h->c1 == 32 && h->c2 == 32;
#define LAUNCH_ASSIGN(XT)                                                                         \
  do {                                                                                            \
    if (h->vl == 32 && cc32) assign_bins_warp_kernel<XT, 32, 32><<<grid, thr, 0, h->stream>>>(a); \
    else if (h->vl == 32) assign_bins_warp_kernel<XT, 32, 0><<<grid, thr, 0, h->stream>>>(a);     \
    else if (h->vl == 16) assign_bins_warp_kernel<XT, 16, 0><<<grid, thr, 0, h->stream>>>(a);     \
    else assign_bins_warp_kernel<XT, 8, 0><<<grid, thr, 0, h->stream>>>(a);                       \
  } while (0)
    if (x_kind == PQT_X_U8) LAUNCH_ASSIGN(uint8_t); else LAUNCH_ASSIGN(float);
#undef LAUNCH_ASSIGN
  } else {
    DevBuf tmp;
    const float* xf = nullptr;
    PQ_TRY(rows_as_f32(h, dX, x_kind, n, tmp, &xf));
    AssignBinsArgs a{};
    a.X = xf; a.cb1 = h->d_cb1.as<float>(); a.cb2 = h->d_cb2.as<float>();
    a.N = n; a.dim = h->dim; a.p = h->p; a.c1 = h->c1; a.c2 = h->c2; a.vl = h->vl;
    a.k1 = P.k1_build; a.npA = pow2ceil(h->c1);
    a.hash = make_fastmod(P.hash_size);
    a.bin_of = d_bin;
    a.counts = nullptr;
    size_t smem = (size_t)(h->dim + 2 * h->p * a.npA + a.k1 * h->p + 2 * h->p * 4) * 4;
    uint32_t grid = std::min<uint32_t>(n, (uint32_t)h->num_sms * 16);
    assign_bins_kernel<<<grid, 128, smem, h->stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(h->stream));  // tmp is freed on return
  }
  CU_TRY(h, cudaGetLastError());
  h->stats.kernel_launches++;
  return PQT_OK;
}

// counts / exclusive prefix / ids grouped by bin (ascending id) from the bins of all N vectors;
// installs the DB (same state as pqt_set_db)
int install_db_from_bins(pqt_index* h, const uint32_t* d_bin, uint32_t N) {
  const uint32_t hs = h->prm.hash_size;
  DevBuf d_counts, d_prefix, tmp;
  CU_TRY(h, d_counts.ensure((size_t)hs * 4));
  CU_TRY(h, d_prefix.ensure((size_t)hs * 4));
  CU_TRY(h, tmp.ensure(scan_tmp_words(hs) * 4));
  CU_TRY(h, cudaMemsetAsync(d_counts.p, 0, (size_t)hs * 4, h->stream));
  bin_histogram_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(d_bin, N, d_counts.as<uint32_t>());
  CU_TRY(h, cudaGetLastError());
  device_exscan_u32(d_counts.as<uint32_t>(), d_prefix.as<uint32_t>(), hs, tmp.as<uint32_t>(), h->stream);
  CU_TRY(h, h->d_dbidx.ensure((size_t)N * 4));
  // directory first: bin_slot_kernel consumes the histogram as its cursor
  PQ_TRY(build_directory(h, d_counts.as<uint32_t>(), d_prefix.as<uint32_t>(), hs, N));
  bin_slot_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(d_bin, N, d_counts.as<uint32_t>(),
                                                         d_prefix.as<uint32_t>(), h->d_dbidx.as<uint32_t>());
  sort_within_bins_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(
      BinDir{h->d_bitmap.as<uint32_t>(), h->d_rank_base.as<uint32_t>(), h->d_cprefix.as<uint32_t>()},
      h->n_nonempty, h->d_dbidx.as<uint32_t>());
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  CU_TRY(h, cudaGetLastError());
  return PQT_OK;
}

}  // namespace

extern "C" {

int pqt_assign_bins(pqt_index* h, const void* X, int x_kind, int x_on_device, uint32_t n,
                    uint32_t* bin_out, int out_on_device) {
  if (!h || !X || !bin_out || !n) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->c1) return fail(h, PQT_ERR_STATE, "no tree loaded");
  const void* dX = nullptr;
  PQ_TRY(stage_rows(h, X, x_kind, x_on_device, n, &dX));
  DevBuf d_bin;
  uint32_t* out = bin_out;
  if (!out_on_device) {
    CU_TRY(h, d_bin.ensure((size_t)n * 4));
    out = d_bin.as<uint32_t>();
  }
  PQ_TRY(launch_assign_bins(h, dX, x_kind, n, out));
  if (!out_on_device)
    CU_TRY(h, cudaMemcpyAsync(bin_out, out, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return PQT_OK;
}

int pqt_set_db_from_bins(pqt_index* h, const uint32_t* bin_of, int on_device, uint32_t N) {
  if (!h || !bin_of || !N) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->c1) return fail(h, PQT_ERR_STATE, "no tree loaded");
  if (h->line_build) return fail(h, PQT_ERR_STATE, "a chunked line encoding is in progress");
  DevBuf d_bin;
  const uint32_t* db = bin_of;
  if (!on_device) {
    CU_TRY(h, d_bin.ensure((size_t)N * 4));
    CU_TRY(h, cudaMemcpyAsync(d_bin.p, bin_of, (size_t)N * 4, cudaMemcpyHostToDevice, h->stream));
    db = d_bin.as<uint32_t>();
  }
  // every bin must be a hash slot (a foreign array would index the histogram out of bounds)
  DevBuf flag;
  CU_TRY(h, flag.ensure(4));
  CU_TRY(h, cudaMemsetAsync(flag.p, 0, 4, h->stream));
  check_below_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(db, N, h->prm.hash_size, flag.as<uint32_t>());
  uint32_t bad = 0;
  CU_TRY(h, cudaMemcpyAsync(&bad, flag.p, 4, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  if (bad) return fail(h, PQT_ERR_INVALID, "bin_of holds values >= hash_size %u", h->prm.hash_size);
  return install_db_from_bins(h, db, N);
}

int pqt_build_kbest_db(pqt_index* h, const float* X, int x_on_device, uint32_t N) {
  if (!h || !X || !N) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->c1) return fail(h, PQT_ERR_STATE, "no tree loaded");
  if (h->line_build) return fail(h, PQT_ERR_STATE, "a chunked line encoding is in progress");
  DevBuf d_bin;
  CU_TRY(h, d_bin.ensure((size_t)N * 4));
  // host data goes through in chunks (bounded staging), device data in one pass
  const uint32_t chunk = x_on_device ? N : std::min<uint32_t>(N, 4u << 20);
  for (uint32_t i0 = 0; i0 < N; i0 += chunk) {
    const uint32_t n = std::min(chunk, N - i0);
    const void* dX = nullptr;
    PQ_TRY(stage_rows(h, X + (size_t)i0 * h->dim, PQT_X_F32, x_on_device, n, &dX));
    PQ_TRY(launch_assign_bins(h, dX, PQT_X_F32, n, d_bin.as<uint32_t>() + i0));
    if (!x_on_device) CU_TRY(h, cudaStreamSynchronize(h->stream));  // staging reuse
  }
  int rc = install_db_from_bins(h, d_bin.as<uint32_t>(), N);
  h->b_x.release();
  return rc;
}

int pqt_line_dist_begin(pqt_index* h, uint32_t N, uint32_t line_parts) {
  if (!h) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->c1) return fail(h, PQT_ERR_STATE, "no tree loaded");
  if (!h->has_db) return fail(h, PQT_ERR_STATE, "line encoding must follow pqt_build_kbest_db / pqt_set_db / pqt_set_db_from_bins");
  if (N != h->N) return fail(h, PQT_ERR_INVALID, "N = %u differs from the DB's %u", N, h->N);
  PQ_TRY(check_lp(h, line_parts));
  if (!is_pow2(h->c1)) return fail(h, PQT_ERR_INVALID, "line encoding needs c1 = 2^n (tree over centroids, pqt/PerturbationProTree.cu:7633-7641)");
  if (h->c1 * line_parts > 1024) return fail(h, PQT_ERR_INVALID, "lineparts * c1 > 1024 (:7694-7697)");
  const uint32_t LP = line_parts;
  h->has_lines = false;
  h->LP = LP;
  h->sl = h->dim / LP;
  PQ_TRY(compute_cbd(h, LP));
  const uint32_t n_local = h->pos_hi - h->pos_lo;
  CU_TRY(h, h->d_codes.ensure(std::max<size_t>((size_t)n_local * LP * 4, 16)));
  CU_TRY(h, h->b_inv.ensure((size_t)N * 4));
  invert_perm_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->d_dbidx.as<uint32_t>(), N, h->b_inv.as<uint32_t>());
  CU_TRY(h, cudaGetLastError());
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->line_build = true;
  return PQT_OK;
}

int pqt_line_dist_chunk(pqt_index* h, const void* X, int x_kind, int x_on_device, uint32_t id0,
                        uint32_t n, uint32_t* lines_out) {
  if (!h || !X || !n) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->line_build) return fail(h, PQT_ERR_STATE, "pqt_line_dist_begin first");
  if ((uint64_t)id0 + n > h->N) return fail(h, PQT_ERR_INVALID, "chunk [%u, %u + %u) exceeds N = %u", id0, id0, n, h->N);
  if (lines_out && h->world != 1) return fail(h, PQT_ERR_STATE, "lines_out needs an unsharded handle (a shard encodes only its own vectors)");
  const uint32_t LP = h->LP;
  const void* dX = nullptr;
  PQ_TRY(stage_rows(h, X, x_kind, x_on_device, n, &dX));
  CU_TRY(h, h->b_stage.ensure((size_t)n * LP * 4));
  const bool sharded = h->world > 1;
  const bool fast = (h->c1 == 16 || h->c1 == 32) && (h->sl == 4 || h->sl == 8 || h->sl == 16) &&
                    ((uintptr_t)dX & 15u) == 0 && (h->dim % 16) == 0;
  if (fast) {
    LineWarpArgs a{};
    a.X = dX; a.n = n; a.dim = h->dim; a.LP = LP;
    a.cb1 = h->d_cb1.as<float>(); a.cbd = h->d_cbd.as<float>();
    a.inv = sharded ? h->b_inv.as<uint32_t>() : nullptr;
    a.id0 = id0; a.pos_lo = h->pos_lo; a.pos_hi = h->pos_hi;
    a.staging = h->b_stage.as<uint32_t>();
    const uint32_t grid = build_grid(h), thr = kBuildWarpsPerCta * 32;
#define LAUNCH_LINE(XT, SLV)                                                                      \
  do {                                                                                            \
    if (h->c1 == 32) line_encode_warp_kernel<XT, SLV, 32><<<grid, thr, 0, h->stream>>>(a);        \
    else line_encode_warp_kernel<XT, SLV, 16><<<grid, thr, 0, h->stream>>>(a);                    \
  } while (0)
#define LAUNCH_LINE_SL(XT)                                                                        \
  do {                                                                                            \
    if (h->sl == 4) LAUNCH_LINE(XT, 4);                                                           \
    else if (h->sl == 8) LAUNCH_LINE(XT, 8);                                                      \
    else LAUNCH_LINE(XT, 16);                                                                     \
  } while (0)
    if (x_kind == PQT_X_U8) LAUNCH_LINE_SL(uint8_t); else LAUNCH_LINE_SL(float);
#undef LAUNCH_LINE_SL
#undef LAUNCH_LINE
    CU_TRY(h, cudaGetLastError());
  } else {
    DevBuf tmp;
    const float* xf = nullptr;
    PQ_TRY(rows_as_f32(h, dX, x_kind, n, tmp, &xf));
    LineEncodeArgs a{};
    a.X = xf; a.cb1 = h->d_cb1.as<float>(); a.cbd = h->d_cbd.as<float>();
    a.ids = nullptr;
    a.N = n; a.dim = h->dim; a.c1 = h->c1; a.LP = LP; a.sl = h->sl;
    a.codes = h->b_stage.as<uint32_t>();
    size_t smem = (size_t)(h->dim + 3 * LP * h->c1) * 4;
    if (smem > 48 * 1024)
      CU_TRY(h, cudaFuncSetAttribute(line_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t grid = std::min<uint32_t>(n, (uint32_t)h->num_sms * 8);
    line_encode_kernel<<<grid, LP * h->c1, smem, h->stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(h->stream));  // tmp is freed on return
  }
  scatter_rows_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->b_stage.as<uint32_t>(), id0, n, LP,
                                                            h->b_inv.as<uint32_t>(), h->pos_lo, h->pos_hi,
                                                            h->d_codes.as<uint32_t>());
  CU_TRY(h, cudaGetLastError());
  h->stats.kernel_launches += 2;
  if (lines_out)
    CU_TRY(h, cudaMemcpyAsync(lines_out, h->b_stage.p, (size_t)n * LP * 4, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return PQT_OK;
}

int pqt_line_dist_end(pqt_index* h) {
  if (!h) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->line_build) return fail(h, PQT_ERR_STATE, "pqt_line_dist_begin first");
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->b_inv.release();
  h->b_stage.release();
  h->b_x.release();
  h->line_build = false;
  h->has_lines = true;
  return PQT_OK;
}

int pqt_line_dist(pqt_index* h, const float* X, int x_on_device, uint32_t N, uint32_t line_parts) {
  if (!h || !X) return PQT_ERR_INVALID;
  PQ_TRY(pqt_line_dist_begin(h, N, line_parts));
  const uint32_t chunk = std::min<uint32_t>(N, 4u << 20);
  int rc = PQT_OK;
  for (uint32_t i0 = 0; i0 < N && rc == PQT_OK; i0 += chunk)
    rc = pqt_line_dist_chunk(h, X + (size_t)i0 * h->dim, PQT_X_F32, x_on_device, i0, std::min(chunk, N - i0), nullptr);
  if (rc != PQT_OK) {
    h->b_inv.release();
    h->b_stage.release();
    h->b_x.release();
    h->line_build = false;
    return rc;
  }
  return pqt_line_dist_end(h);
}

// ---- query ----------------------------------------------------------------------------
int pqt_candidate_width(const pqt_index* h, uint32_t k, uint32_t* max_vec) {
  if (!h || !max_vec || !k) return PQT_ERR_INVALID;
  *max_vec = candidate_width(h, k);
  return PQT_OK;
}

static int query_common(pqt_index* h, const float* Q, int q_on_device, uint32_t QN, uint32_t k,
                        uint32_t* idx, float* dist, int out_on_device, bool big) {
  if (!h || !Q || !idx || !dist) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  PQ_TRY(check_query_state(h, QN, k));
  if (big) {
    if (h->p != 4) return fail(h, PQT_ERR_INVALID, "queryBIGKNNRerank2 merges parts (0,1) and (2,3): p must be 4 (pqt/PerturbationProTree.cu:2952,3061-3070)");
    if (h->prm.big_k1 > h->c1 || h->prm.big_k1 > 32) return fail(h, PQT_ERR_INVALID, "big_k1 %u > c1 %u", h->prm.big_k1, h->c1);
    if (h->prm.big_k1 * h->c2 < kBigKMax) return fail(h, PQT_ERR_INVALID, "big_k1*c2 < 64: the 2-D merge reads 64 sorted entries per part (:3729)");
    if (pow2ceil(h->prm.big_k1 * h->c2) > 1024) return fail(h, PQT_ERR_INVALID, "big_k1*c2 > 1024");
  }
  if (h->world != 1)
    return fail(h, PQT_ERR_STATE, "sharded handle: a shard answers queries together with the others through pqt_shard_dispatch / pqt_shard_scan_p2p / pqt_shard_rank");
  const uint32_t max_vec = candidate_width(h, k);
  const float* dQ = Q;
  if (!q_on_device) {
    CU_TRY(h, h->s_q.ensure((size_t)QN * h->dim * 4));
    CU_TRY(h, cudaMemcpyAsync(h->s_q.p, Q, (size_t)QN * h->dim * 4, cudaMemcpyHostToDevice, h->stream));
    dQ = h->s_q.as<float>();
  }
  float* d_out_dist = dist;
  uint32_t* d_out_idx = idx;
  if (!out_on_device) {
    CU_TRY(h, h->s_outd.ensure((size_t)QN * k * 4));
    CU_TRY(h, h->s_outi.ensure((size_t)QN * k * 4));
    d_out_dist = h->s_outd.as<float>();
    d_out_idx = h->s_outi.as<uint32_t>();
  }
  // the fused scan + rank kernel needs its tables and candidate arrays in shared memory
  // (same test as in run_scan_chain)
  const bool fused = std::min(h->LP <= 16 ? rerank_smem_bytes(h->c1, h->LP, max_vec, 4, false) : ~(size_t)0,
                              rerank_smem_bytes(h->c1, h->LP, max_vec, 2, true)) <= (size_t)227 * 1024;
  // Host outputs: queries go through in slabs so that the device->host copy of one slab
  // (on the copy stream) overlaps the kernels of the next.  Device outputs / debug
  // recording: one pass.
  // The copies are the bottleneck at large k (PCIe), so what counts is how soon the first one can
  // start: the slabs grow from kSlabQueries / 4 to kSlabQueries.
  std::vector<uint32_t> slab_lo{0u};
  if (out_on_device || h->debug) {
    slab_lo.push_back(QN);
  } else {
    uint32_t sz = QN > kSlabQueries ? kSlabQueries / 4 : kSlabQueries;
    while (slab_lo.back() < QN) {
      slab_lo.push_back(std::min<uint64_t>(QN, (uint64_t)slab_lo.back() + sz));
      sz = std::min<uint32_t>(kSlabQueries, sz * 2);
    }
  }
  const uint32_t nslabs = (uint32_t)slab_lo.size() - 1;
  h->scratch_queries = 0;
  for (uint32_t s = 0; s < nslabs; s++) h->scratch_queries = std::max(h->scratch_queries, slab_lo[s + 1] - slab_lo[s]);
  if (!out_on_device) {
    while (h->slab_ev.size() < nslabs) {
      cudaEvent_t e;
      CU_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->slab_ev.push_back(e);
    }
  }
  auto issue_copy = [&](uint32_t s) -> int {
    const uint32_t q0 = slab_lo[s], n = slab_lo[s + 1] - q0;
    CU_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->slab_ev[s], 0));
    CU_TRY(h, cudaMemcpyAsync(idx + (size_t)q0 * k, d_out_idx + (size_t)q0 * k, (size_t)n * k * 4,
                              cudaMemcpyDeviceToHost, h->copy_stream));
    CU_TRY(h, cudaMemcpyAsync(dist + (size_t)q0 * k, d_out_dist + (size_t)q0 * k, (size_t)n * k * 4,
                              cudaMemcpyDeviceToHost, h->copy_stream));
    return PQT_OK;
  };
  for (uint32_t s = 0; s < nslabs; s++) {
    const uint32_t q0 = slab_lo[s], n = slab_lo[s + 1] - q0;
    const float* q = dQ + (size_t)q0 * h->dim;
    float* od = d_out_dist + (size_t)q0 * k;
    uint32_t* oi = d_out_idx + (size_t)q0 * k;
    if (fused) {
      PQ_TRY(run_scan_chain(h, q, n, k, nullptr, nullptr, od, oi, big));
      if (h->profile && !h->split_ranked) {
        CU_TRY(h, cudaEventRecord(h->ev[4], h->stream));
        CU_TRY(h, cudaEventRecord(h->ev[5], h->stream));
      }
    } else {
      CU_TRY(h, h->s_val.ensure((size_t)h->scratch_queries * max_vec * 4));
      CU_TRY(h, h->s_idx.ensure((size_t)h->scratch_queries * max_vec * 4));
      PQ_TRY(run_scan_chain(h, q, n, k, h->s_val.as<float>(), h->s_idx.as<uint32_t>(), nullptr, nullptr, big));
      if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[4], h->stream));
      PQ_TRY(run_rank(h, h->s_val.as<float>(), h->s_idx.as<uint32_t>(), n, max_vec, k, od, oi));
      if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[5], h->stream));
    }
    if (!out_on_device) CU_TRY(h, cudaEventRecord(h->slab_ev[s], h->stream));
    if (h->profile) {
      CU_TRY(h, cudaStreamSynchronize(h->stream));
      accumulate_profile(h, n, true);
    }
    // the previous slab's copy is issued after this slab's kernels are in flight
    if (!out_on_device && s > 0) PQ_TRY(issue_copy(s - 1));
  }
  if (!out_on_device) {
    PQ_TRY(issue_copy(nslabs - 1));
    CU_TRY(h, cudaStreamSynchronize(h->copy_stream));
  }
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  return PQT_OK;
}

int pqt_query_knn(pqt_index* h, const float* Q, int q_on_device, uint32_t QN, uint32_t k,
                  uint32_t* idx, float* dist, int out_on_device) {
  return query_common(h, Q, q_on_device, QN, k, idx, dist, out_on_device, false);
}

int pqt_query_big_knn_rerank2(pqt_index* h, const float* Q, int q_on_device, uint32_t QN,
                              uint32_t k, uint32_t* idx, float* dist, int out_on_device) {
  return query_common(h, Q, q_on_device, QN, k, idx, dist, out_on_device, true);
}

// ---- multi-GPU: bin-range shards, dispatch + scan fused with the exchange over peer memory ---
int pqt_shard_exchange_alloc(pqt_index* h, uint32_t q_per_rank, uint32_t max_vec) {
  if (!h || !q_per_rank || !max_vec) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  const size_t qn = (size_t)q_per_rank * h->world;
  // cudaIpcGetMemHandle needs plain cudaMalloc allocations
  h->x_val.release();
  h->x_inbox.release();
  h->x_cnt.release();
  CU_TRY(h, h->x_val.ensure((size_t)q_per_rank * max_vec * 4));
  CU_TRY(h, h->x_inbox.ensure(qn * max_vec * 8));
  CU_TRY(h, h->x_cnt.ensure(qn * 4));
  CU_TRY(h, cudaMemset(h->x_cnt.p, 0, qn * 4));
  h->x_q_per_rank = q_per_rank;
  h->x_max_vec = max_vec;
  return PQT_OK;
}

int pqt_shard_exchange_handle(pqt_index* h, void* handle192) {
  if (!h || !handle192) return PQT_ERR_INVALID;
  if (!h->x_val.p) return fail(h, PQT_ERR_STATE, "pqt_shard_exchange_alloc first");
  CU_TRY(h, cudaSetDevice(h->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t hv, hi, hc;
  CU_TRY(h, cudaIpcGetMemHandle(&hv, h->x_val.p));
  CU_TRY(h, cudaIpcGetMemHandle(&hi, h->x_inbox.p));
  CU_TRY(h, cudaIpcGetMemHandle(&hc, h->x_cnt.p));
  std::memcpy(handle192, &hv, 64);
  std::memcpy(static_cast<char*>(handle192) + 64, &hi, 64);
  std::memcpy(static_cast<char*>(handle192) + 128, &hc, 64);
  return PQT_OK;
}

int pqt_shard_exchange_open(pqt_index* h, uint32_t world, const void* handles) {
  if (!h || !handles || world == 0 || world > 8) return PQT_ERR_INVALID;
  if (world != h->world) return fail(h, PQT_ERR_STATE, "world differs from pqt_set_shard");
  if (!h->x_val.p) return fail(h, PQT_ERR_STATE, "pqt_shard_exchange_alloc first");
  CU_TRY(h, cudaSetDevice(h->device));
  for (uint32_t r = 0; r < world; r++) {
    if (r == h->rank) {
      h->x_peer_val[r] = h->x_val.as<float>();
      h->x_peer_inbox[r] = h->x_inbox.as<uint2>();
      h->x_peer_cnt[r] = h->x_cnt.as<uint32_t>();
      continue;
    }
    cudaIpcMemHandle_t hd[3];
    std::memcpy(hd, static_cast<const char*>(handles) + (size_t)r * 192, 192);
    void* pp[3] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < 3; j++) CU_TRY(h, cudaIpcOpenMemHandle(&pp[j], hd[j], cudaIpcMemLazyEnablePeerAccess));
    h->x_peer_val[r] = static_cast<float*>(pp[0]);
    h->x_peer_inbox[r] = static_cast<uint2*>(pp[1]);
    h->x_peer_cnt[r] = static_cast<uint32_t*>(pp[2]);
    h->x_ipc_opened[r] = true;
  }
  h->x_world = world;
  return PQT_OK;
}

int pqt_shard_exchange_set_peers(pqt_index* h, uint32_t world, void* const* val_ptrs,
                                 void* const* inbox_ptrs, void* const* cnt_ptrs) {
  if (!h || !val_ptrs || !inbox_ptrs || !cnt_ptrs || world == 0 || world > 8) return PQT_ERR_INVALID;
  if (world != h->world) return fail(h, PQT_ERR_STATE, "world differs from pqt_set_shard");
  for (uint32_t r = 0; r < world; r++) {
    h->x_peer_val[r] = static_cast<float*>(val_ptrs[r]);
    h->x_peer_inbox[r] = static_cast<uint2*>(inbox_ptrs[r]);
    h->x_peer_cnt[r] = static_cast<uint32_t*>(cnt_ptrs[r]);
  }
  h->x_world = world;
  return PQT_OK;
}

int pqt_shard_exchange_ptrs(pqt_index* h, void** val_ptr, void** inbox_ptr, void** cnt_ptr) {
  if (!h || !val_ptr || !inbox_ptr || !cnt_ptr) return PQT_ERR_INVALID;
  if (!h->x_val.p) return fail(h, PQT_ERR_STATE, "pqt_shard_exchange_alloc first");
  *val_ptr = h->x_val.p;
  *inbox_ptr = h->x_inbox.p;
  *cnt_ptr = h->x_cnt.p;
  return PQT_OK;
}

// Steps A-E1 for the own queries, Step B for everybody else's, then the dispatch of the own
// queries' candidates to the shards that hold them
int pqt_shard_dispatch(pqt_index* h, const float* Q, int q_on_device, uint32_t QN, uint32_t k,
                       uint32_t q_lo, uint32_t q_hi) {
  if (!h || !Q || q_lo >= q_hi || q_hi > QN) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  PQ_TRY(check_query_state(h, QN, k));
  const pqt_params& P = h->prm;
  const uint32_t max_vec = candidate_width(h, k);
  const uint32_t n = P.k1 * h->c2, npC = pow2ceil(n);
  const bool warp_path = h->c1 <= 32 && P.k1 <= 32 && (h->LP % h->p) == 0 &&
                         (h->vl == 8 || h->vl == 16 || h->vl == 32) && (h->dim % 4) == 0;
  if (!warp_path || h->p > 4 || (h->LP != 16 && h->LP != 32))
    return fail(h, PQT_ERR_INVALID, "the multi-GPU path supports c1 <= 32, p <= 4, dim/p in {8,16,32}, lineparts 16 / 32");
  if (!h->x_world || h->x_world != h->world) return fail(h, PQT_ERR_STATE, "exchange buffers are not connected (pqt_shard_exchange_open / _set_peers)");
  if (max_vec != h->x_max_vec) return fail(h, PQT_ERR_INVALID, "candidate width %u differs from the exchange buffers' %u", max_vec, h->x_max_vec);
  if ((uint64_t)h->x_q_per_rank * h->world < QN || q_hi - q_lo > h->x_q_per_rank || q_lo != h->rank * h->x_q_per_rank)
    return fail(h, PQT_ERR_INVALID, "query slice [%u, %u) does not match rank %u of q_per_rank %u", q_lo, q_hi, h->rank, h->x_q_per_rank);
  const float* dQ = Q;
  if (!q_on_device) {
    CU_TRY(h, h->s_q.ensure((size_t)QN * h->dim * 4));
    CU_TRY(h, cudaMemcpyAsync(h->s_q.p, Q, (size_t)QN * h->dim * 4, cudaMemcpyHostToDevice, h->stream));
    dQ = h->s_q.as<float>();
  }
  PQ_TRY(ensure_dist_seq(h, h->c2 * P.k1));
  const uint32_t nq = q_hi - q_lo;
  CU_TRY(h, h->s_lut.ensure((size_t)QN * h->c1 * 32 * sizeof(float)));
  CU_TRY(h, h->s_idx16.ensure((size_t)nq * h->p * 16 * sizeof(uint32_t)));
  CU_TRY(h, h->s_cand.ensure((size_t)nq * max_vec * sizeof(uint32_t)));
  CU_TRY(h, h->s_nvec.ensure((size_t)nq * sizeof(uint32_t)));
  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[0], h->stream));
  {
    TablesWarpArgs w{};
    TablesArgs& a = w.t;
    a.Q = dQ + (size_t)q_lo * h->dim;
    a.cb1 = h->d_cb1.as<float>();
    a.cb2 = h->d_cb2.as<float>();
    a.QN = nq; a.dim = h->dim; a.p = h->p; a.c1 = h->c1; a.c2 = h->c2; a.LP = h->LP;
    a.k1 = P.k1; a.vl = h->vl; a.sl = h->sl;
    a.npA = pow2ceil(h->c1);
    a.npC = npC;
    a.m = h->seq_m;
    a.lut_dup = h->s_lut.as<float>() + (size_t)q_lo * h->c1 * 32;
    a.idx16 = h->s_idx16.as<uint32_t>();
    w.cb1T = h->d_cb1T.as<float>();
    w.cb2T = h->d_cb2T.as<float>();
    size_t smem = (size_t)(h->dim + h->c1 * 32 + 2 * kTablesWarps * a.npC) * 4;
    uint32_t grid = std::min<uint32_t>(nq, (uint32_t)h->num_sms * 12);
    if (h->vl == 32 && h->c1 == 32 && h->c2 == 32) {
      tables_warp_kernel<32, 32><<<grid, kTablesWarps * 32, smem, h->stream>>>(w);
    } else {
      switch (h->vl) {
        case 8: tables_warp_kernel<8, 0><<<grid, kTablesWarps * 32, smem, h->stream>>>(w); break;
        case 16: tables_warp_kernel<16, 0><<<grid, kTablesWarps * 32, smem, h->stream>>>(w); break;
        default: tables_warp_kernel<32, 0><<<grid, kTablesWarps * 32, smem, h->stream>>>(w); break;
      }
    }
    CU_TRY(h, cudaGetLastError());
    h->stats.kernel_launches++;
    // Step B for the queries owned by other ranks (every shard scans its own candidates of
    // every query; recomputing c1*LP short segment distances is cheaper than shipping 4 KB),
    // on a second stream beside Steps A-E1 and the dispatch of the own queries
    const size_t lsmem = (size_t)(h->dim + h->c1 * 32) * 4;
    CU_TRY(h, cudaEventRecord(h->aux_ev[0], h->stream));  // the queries are on the device
    CU_TRY(h, cudaStreamWaitEvent(h->aux_stream, h->aux_ev[0], 0));
    if (q_lo > 0) {
      lut_kernel<<<std::min<uint32_t>(q_lo, (uint32_t)h->num_sms * 8), 128, lsmem, h->aux_stream>>>(
          dQ, h->d_cb1T.as<float>(), 0, q_lo, h->dim, h->c1, h->LP, h->sl, h->s_lut.as<float>());
      h->stats.kernel_launches++;
    }
    if (q_hi < QN) {
      lut_kernel<<<std::min<uint32_t>(QN - q_hi, (uint32_t)h->num_sms * 8), 128, lsmem, h->aux_stream>>>(
          dQ, h->d_cb1T.as<float>(), q_hi, QN, h->dim, h->c1, h->LP, h->sl, h->s_lut.as<float>());
      h->stats.kernel_launches++;
    }
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaEventRecord(h->aux_ev[1], h->aux_stream));
  }
  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[1], h->stream));
  const bool want_roots = !(getenv("PQT_BINS_DEDUPE") && atoi(getenv("PQT_BINS_DEDUPE")) == 0);
  PQ_TRY(launch_bins_p4(h, P, max_vec, nq, h->s_idx16.as<uint32_t>(), h->s_cand.as<uint32_t>(),
                        h->s_nvec.as<uint32_t>(), nullptr, nullptr, want_roots));
  {
    DispatchArgs d{};
    d.list_pos = h->have_roots ? h->s_rootpos.as<uint32_t>() : h->s_cand.as<uint32_t>();
    d.n_list = h->have_roots ? h->s_nroot.as<uint32_t>() : h->s_nvec.as<uint32_t>();
    d.q_own = nq; d.q_first = q_lo; d.max_vec = max_vec; d.world = h->world;
    for (uint32_t r = 0; r <= h->world; r++) d.shard_lo[r] = (uint32_t)((uint64_t)h->N * r / h->world);
    for (uint32_t r = 0; r < h->world; r++) {
      d.peer_inbox[r] = h->x_peer_inbox[r];
      d.peer_cnt[r] = h->x_peer_cnt[r];
    }
    dispatch_kernel<<<std::min<uint32_t>(nq, (uint32_t)h->num_sms * 8), 256, 0, h->stream>>>(d);
    CU_TRY(h, cudaGetLastError());
    h->stats.kernel_launches++;
  }
  h->x_lut_QN = QN;
  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[2], h->stream));
  CU_TRY(h, cudaStreamWaitEvent(h->stream, h->aux_ev[1], 0));  // the LUTs are ready for the scan
  if (h->profile) {
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    float t = 0;
    cudaEventElapsedTime(&t, h->ev[0], h->ev[1]);
    h->stats.ms_tables += t;
    cudaEventElapsedTime(&t, h->ev[1], h->ev[2]);
    h->stats.ms_bins += t;
    h->stats.queries += nq;
    h->stats.calls++;
    std::vector<uint32_t> nv(nq);
    cudaMemcpy(nv.data(), h->s_nvec.p, (size_t)nq * 4, cudaMemcpyDeviceToHost);
    uint64_t sum = 0;
    for (uint32_t v : nv) sum += v;
    h->stats.candidates += sum;
  }
  return PQT_OK;
}

// ADC scan of this shard's inbox (its own candidates of ALL queries); every distance is stored
// into the distance array of the rank that owns the query.  Call after a cross-rank barrier
// that orders it behind every rank's pqt_shard_dispatch.
int pqt_shard_scan_p2p(pqt_index* h, uint32_t QN, uint32_t k) {
  if (!h) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  PQ_TRY(check_query_state(h, QN, k));
  const uint32_t max_vec = candidate_width(h, k);
  if (!h->x_world || h->x_world != h->world) return fail(h, PQT_ERR_STATE, "exchange buffers are not connected (pqt_shard_exchange_open / _set_peers)");
  if (max_vec != h->x_max_vec) return fail(h, PQT_ERR_INVALID, "candidate width %u differs from the exchange buffers' %u", max_vec, h->x_max_vec);
  if ((uint64_t)h->x_q_per_rank * h->world < QN) return fail(h, PQT_ERR_INVALID, "QN exceeds q_per_rank * world");
  if (h->x_lut_QN != QN) return fail(h, PQT_ERR_STATE, "pqt_shard_scan_p2p needs the LUTs of the same %u queries (pqt_shard_dispatch was called for %u)", QN, h->x_lut_QN);
  if (h->LP != 16 && h->LP != 32) return fail(h, PQT_ERR_INVALID, "the multi-GPU path is built for lineparts 16 and 32");
  StreamScanArgs sa{};
  sa.codes = h->d_codes.as<uint32_t>();
  sa.cand_pos = nullptr;
  sa.n_vec = h->x_cnt.as<uint32_t>();
  sa.lut_dup = h->s_lut.as<float>();
  sa.cbd = h->LP == 32 ? h->d_cbd_dup.as<float>() : h->d_cbd.as<float>();
  sa.QN = QN; sa.c1 = h->c1; sa.max_vec = max_vec;
  sa.inbox = h->x_inbox.as<uint2>();
  sa.q_per_rank = h->x_q_per_rank;
  for (uint32_t r = 0; r < h->world; r++) sa.peer_val[r] = h->x_peer_val[r];
  const size_t wsmem = inbox_scan_smem_bytes(h->c1, h->LP, h->LP == 32);
  if (wsmem > (size_t)227 * 1024) return fail(h, PQT_ERR_INVALID, "c1 = %u > 32 is not supported by the ADC scan yet", h->c1);
  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[2], h->stream));
  CU_TRY(h, h->d_sched.ensure(64));
  CU_TRY(h, cudaMemsetAsync(h->d_sched.p, 0, 4, h->stream));
  const uint32_t wgrid = std::min<uint32_t>((QN + kInboxWarps - 1) / kInboxWarps, (uint32_t)h->num_sms);
  if (h->LP == 32) {
    CU_TRY(h, cudaFuncSetAttribute(adc_inbox_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
    adc_inbox_kernel<32, true><<<wgrid, kInboxWarps * 32, wsmem, h->stream>>>(sa, h->d_sched.as<uint32_t>());
  } else {
    CU_TRY(h, cudaFuncSetAttribute(adc_inbox_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
    adc_inbox_kernel<16, false><<<wgrid, kInboxWarps * 32, wsmem, h->stream>>>(sa, h->d_sched.as<uint32_t>());
  }
  CU_TRY(h, cudaGetLastError());
  h->stats.kernel_launches++;
  h->stats.scan_launches++;
  h->stats.stream_scan_launches++;
  if (h->profile) {
    CU_TRY(h, cudaEventRecord(h->ev[3], h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    float t = 0;
    cudaEventElapsedTime(&t, h->ev[2], h->ev[3]);
    h->stats.ms_scan += t;
  }
  return PQT_OK;
}

// Ranking of the own queries (those of the last pqt_shard_dispatch).  Call after a cross-rank
// barrier that orders it behind every rank's pqt_shard_scan_p2p.
int pqt_shard_rank(pqt_index* h, uint32_t q_own, uint32_t k, uint32_t* idx, float* dist, int out_on_device) {
  if (!h || !idx || !dist || !q_own || !k) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  if (!h->x_val.p) return fail(h, PQT_ERR_STATE, "pqt_shard_exchange_alloc first");
  if (q_own > h->x_q_per_rank) return fail(h, PQT_ERR_INVALID, "q_own exceeds the exchange buffers");
  const uint32_t max_vec = h->x_max_vec;
  if (!is_pow2(max_vec) || max_vec > 4096 || k > max_vec) return fail(h, PQT_ERR_INVALID, "bad candidate width / k");
  float* d_out_dist = dist;
  uint32_t* d_out_idx = idx;
  if (!out_on_device) {
    CU_TRY(h, h->s_outd.ensure((size_t)q_own * k * 4));
    CU_TRY(h, h->s_outi.ensure((size_t)q_own * k * 4));
    d_out_dist = h->s_outd.as<float>();
    d_out_idx = h->s_outi.as<uint32_t>();
  }
  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[4], h->stream));
  size_t smem = rank2_smem_bytes(max_vec);
  if (smem > 48 * 1024)
    CU_TRY(h, cudaFuncSetAttribute(rank2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // host outputs: rank in slabs so that the device->host copy of one slab overlaps the
  // ranking of the next (same scheme as pqt_query_knn)
  const uint32_t slab = out_on_device ? q_own : std::min<uint32_t>(q_own, std::max<uint32_t>(256u, (q_own + 3) / 4));
  const uint32_t nslabs = (q_own + slab - 1) / slab;
  if (!out_on_device) {
    while (h->slab_ev.size() < nslabs) {
      cudaEvent_t e;
      CU_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->slab_ev.push_back(e);
    }
  }
  auto issue_copy = [&](uint32_t sidx) -> int {
    const uint32_t q0 = sidx * slab, n = std::min(slab, q_own - q0);
    CU_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->slab_ev[sidx], 0));
    CU_TRY(h, cudaMemcpyAsync(idx + (size_t)q0 * k, d_out_idx + (size_t)q0 * k, (size_t)n * k * 4,
                              cudaMemcpyDeviceToHost, h->copy_stream));
    CU_TRY(h, cudaMemcpyAsync(dist + (size_t)q0 * k, d_out_dist + (size_t)q0 * k, (size_t)n * k * 4,
                              cudaMemcpyDeviceToHost, h->copy_stream));
    return PQT_OK;
  };
  for (uint32_t sidx = 0; sidx < nslabs; sidx++) {
    const uint32_t q0 = sidx * slab, n = std::min(slab, q_own - q0);
    Rank2Args a{};
    a.val = h->x_val.as<float>() + (size_t)q0 * max_vec;
    a.idx = h->s_cand.as<uint32_t>() + (size_t)q0 * max_vec;  // global bin-order positions
    a.ids = h->d_dbidx.as<uint32_t>();                          // replicated: id of every position
    a.QN = n; a.max_vec = max_vec; a.k = k;
    a.out_dist = d_out_dist + (size_t)q0 * k;
    a.out_idx = d_out_idx + (size_t)q0 * k;
    a.exact_counter = h->d_exact.as<unsigned long long>();
    a.tie_counter = h->d_exact.as<unsigned long long>() + 1;
    a.fast_rank = (h->prm.rank_mode == 0) ? 1u : 0u;
      a.n_vec = h->s_nvec.as<uint32_t>() + q0;
    a.ridx = h->have_roots ? h->s_ridx.as<uint16_t>() + (size_t)q0 * max_vec : nullptr;
    CU_TRY(h, h->d_sched.ensure(64));
    CU_TRY(h, cudaMemsetAsync(h->d_sched.as<uint32_t>() + 1, 0, 4, h->stream));
    a.next_query = h->d_sched.as<uint32_t>() + 1;
    rank2_kernel<false><<<std::min<uint32_t>(n, (uint32_t)h->num_sms * 4), kRank2Threads, smem, h->stream>>>(a);
    CU_TRY(h, cudaGetLastError());
    h->stats.kernel_launches++;
    if (!out_on_device) {
      CU_TRY(h, cudaEventRecord(h->slab_ev[sidx], h->stream));
      if (sidx > 0) PQ_TRY(issue_copy(sidx - 1));
    }
  }
  if (h->profile) CU_TRY(h, cudaEventRecord(h->ev[5], h->stream));
  if (!out_on_device) {
    PQ_TRY(issue_copy(nslabs - 1));
    CU_TRY(h, cudaStreamSynchronize(h->copy_stream));
  }
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  if (h->profile) {
    float t = 0;
    cudaEventElapsedTime(&t, h->ev[4], h->ev[5]);
    h->stats.ms_sort += t;
  }
  return PQT_OK;
}

// ---- measurement / introspection --------------------------------------------------------
int pqt_profile_enable(pqt_index* h, int on) {
  if (!h) return PQT_ERR_INVALID;
  h->profile = on != 0;
  return PQT_OK;
}
int pqt_get_stats(const pqt_index* h, pqt_stats* st) {
  if (!h || !st) return PQT_ERR_INVALID;
  *st = h->stats;
  unsigned long long ex[2] = {0, 0};
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaMemcpy(ex, h->d_exact.p, 16, cudaMemcpyDeviceToHost);
  st->exact_rank_queries = ex[0];
  st->tie_resolved_queries = ex[1];
  return PQT_OK;
}
int pqt_reset_stats(pqt_index* h) {
  if (!h) return PQT_ERR_INVALID;
  std::memset(&h->stats, 0, sizeof(h->stats));
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaMemset(h->d_exact.p, 0, 16);
  return PQT_OK;
}
int pqt_debug_enable(pqt_index* h, int on) {
  if (!h) return PQT_ERR_INVALID;
  h->debug = on != 0;
  return PQT_OK;
}

int pqt_debug_stage(const pqt_index* hc, int stage, void* host_out, size_t bytes) {
  pqt_index* h = const_cast<pqt_index*>(hc);
  if (!h || !host_out) return PQT_ERR_INVALID;
  CU_TRY(h, cudaSetDevice(h->device));
  const void* src = nullptr;
  size_t need = 0;
  const uint32_t QN = h->dbg_QN, n = h->prm.k1 * h->c2;
  switch (stage) {
    case PQT_STAGE_ASSIGN: src = h->g_assign.p; need = (size_t)QN * h->prm.k1 * h->p * 4; break;
    case PQT_STAGE_LUT: src = h->g_lut.p; need = (size_t)QN * h->LP * h->c1 * 4; break;
    case PQT_STAGE_ASSIGN_VAL: src = h->g_aval.p; need = (size_t)QN * h->p * n * 4; break;
    case PQT_STAGE_ASSIGN_IDX: src = h->g_aidx.p; need = (size_t)QN * h->p * n * 4; break;
    case PQT_STAGE_BINS: src = h->g_bins.p; need = (size_t)QN * h->prm.max_bins * 4; break;
    case PQT_STAGE_NBINS: src = h->g_nbins.p; need = (size_t)QN * 4; break;
    case PQT_STAGE_SELECT_IDX: src = h->g_sel.p; need = (size_t)QN * h->dbg_maxvec * 4; break;
    case PQT_STAGE_NVEC: src = h->s_nvec.p; need = (size_t)QN * 4; break;
    case PQT_STAGE_CB_DIST:
      if (!h->has_lines) return fail(h, PQT_ERR_STATE, "no line codes");
      src = h->d_cbd.p; need = (size_t)h->c1 * h->c1 * h->LP * 4; break;
    case PQT_STAGE_DIST_SEQ:
      if (h->h_distseq.empty()) return fail(h, PQT_ERR_STATE, "no query has run yet");
      if (bytes != kNumDistSeq * 4) return fail(h, PQT_ERR_INVALID, "size mismatch");
      std::memcpy(host_out, h->h_distseq.data(), bytes);
      return PQT_OK;
    case PQT_STAGE_DIST_SEQ_2D:
      if (h->h_seq2d.empty()) return fail(h, PQT_ERR_STATE, "no BIG query has run yet");
      if (bytes != h->h_seq2d.size() * 4) return fail(h, PQT_ERR_INVALID, "size mismatch");
      std::memcpy(host_out, h->h_seq2d.data(), bytes);
      return PQT_OK;
    case PQT_STAGE_BIG_BINS: src = h->g_bigbins.p; need = (size_t)h->dbg_big_QN * h->dbg_big_cap * 4; break;
    case PQT_STAGE_BIG_NBINS: src = h->g_bignbins.p; need = (size_t)h->dbg_big_QN * 4; break;
    case PQT_STAGE_RERANK_PHASES:
      if (!h->debug || !QN || !h->g_phases.p) return fail(h, PQT_ERR_STATE, "debug recording is off or no query has run");
      src = h->g_phases.p; need = (size_t)QN * 64; break;
    default: return fail(h, PQT_ERR_INVALID, "unknown stage %d", stage);
  }
  if ((stage == PQT_STAGE_BIG_BINS || stage == PQT_STAGE_BIG_NBINS) && (!h->debug || !h->dbg_big_QN))
    return fail(h, PQT_ERR_STATE, "debug recording is off or no BIG query has run");
  if (stage <= PQT_STAGE_NVEC && (!h->debug || !QN)) return fail(h, PQT_ERR_STATE, "debug recording is off or no query has run");
  if (bytes != need) return fail(h, PQT_ERR_INVALID, "stage %d holds %zu bytes, caller passed %zu", stage, need, bytes);
  CU_TRY(h, cudaMemcpy(host_out, src, need, cudaMemcpyDeviceToHost));
  return PQT_OK;
}

}  // extern "C"
