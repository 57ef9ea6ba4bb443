// big_kernels.cuh -- bin selection of the 1-B variant queryBIGKNNRerank2
// (pqt/PerturbationProTree.cu:8596-8701): getBIGBins2D (:3702-3778) =
// selectBinKernel2D2Parts (:2914-3006) + selectBinKernel2DFinal (:3012-3188), followed by
// Step E1 with the per-bin cap of rerankBIGKBestVectors2 (:6525).  One CTA of 1024
// threads per query (the reference's round size).  Steps A-C and E2 are the kernels of the
// small-DB path (k1 = 16; the tables kernel also emits the 64 best Step-C entries per part).
//
// Reproduced on purpose (SURVEY.md App. B style quirks, all in the oracle too):
//   * the kept-bin counter adds the EXCLUSIVE scan value of the last thread (:3163), so a bin
//     kept by thread 1023 is overwritten by the next round;
//   * rounds read d_distSeq at slope*65536 + round*1024, i.e. past the slope's own 65536
//     codes; the walk stops where the reference would leave its allocation.
#pragma once
#include "common.cuh"
#include "query_kernels.cuh"
#include "rerank_kernels.cuh"

namespace pqtb {

constexpr uint32_t kNumAnisoDir = 10;      // pqt/ProTree.hh:12
#define PQTB_ANISO_BASE 1.2f               // pqt/ProTree.hh:13
constexpr uint32_t kBigKMax = 64;          // :3729
constexpr uint32_t kBigInter = 256;        // :3735
constexpr uint32_t kBigDistCluster = 512;  // prepare2DDistSequence(512), test/test1B.cpp:1215
constexpr int kBigThreads = 1024;

// computeSlopeIdx (:2839-2862), same expression so that the device math is the reference's
__device__ __forceinline__ uint32_t slope_index(const float* val0, const float* val1, uint32_t N) {
  uint32_t sampleIdx = sqrtf(2.f * N);
  float slope = (val1[sampleIdx] + val1[sampleIdx - 1] - 2 * val1[0]) /
                (val0[sampleIdx] + val0[sampleIdx - 1] - 2 * val0[0]);
  int si = roundf(logf(slope) / logf(PQTB_ANISO_BASE)) + (kNumAnisoDir / 2);
  si = (si >= (int)kNumAnisoDir) ? ((int)kNumAnisoDir - 1) : si;
  si = (si < 0) ? 0 : si;
  return (uint32_t)si;
}

struct BinsBigArgs {
  const float* top_val;     // [QN][p][64] first 64 entries of the sorted Step-C lists
  const uint32_t* top_idx;  // [QN][p][64]
  const uint32_t* seq2d;    // [10][65536] prepare2DDistSequence
  BinDir dir;
  MagicMod hash;
  uint32_t QN, c1c2;
  uint32_t k2;  // kVec of the query (:8673)
  uint32_t max_trials, max_bins, max_vec, max_vec_per_bin;
  uint32_t list_cap;
  uint32_t* cand_pos;   // [QN][max_vec]
  uint32_t* n_vec;      // [QN]
  uint32_t* dbg_bins;   // [QN][list_cap] or null
  uint32_t* dbg_nbins;  // [QN] or null
  uint32_t* next_query;  // optional work counter (zeroed before the launch): the number of rounds
                         // differs from query to query, so CTAs draw their queries
};

// dynamic smem: l_val[512] | l_idx[512] | r_dist[1024] | r_bin[1024] | in_val[128] | in_idx[128]
//               | list[list_cap] | warp_sums[32] | misc[4] | pay u16[1024]
// The rounds' 1024-wide and the merges' 256-wide sorts are the reference's network (ties between
// equal summed distances follow it), run by grp_sort_dispatch: same compare-exchanges in
// registers / shuffles, shared memory and a barrier only for the cross-warp distances.
__global__ void __launch_bounds__(kBigThreads) bins_big_kernel(BinsBigArgs a) {
  extern __shared__ float smem_f[];
  float* l_val = smem_f;
  uint32_t* l_idx = reinterpret_cast<uint32_t*>(l_val + 2 * kBigInter);
  float* r_dist = reinterpret_cast<float*>(l_idx + 2 * kBigInter);
  uint32_t* r_bin = reinterpret_cast<uint32_t*>(r_dist + kBigThreads);
  float* in_val = reinterpret_cast<float*>(r_bin + kBigThreads);
  uint32_t* in_idx = reinterpret_cast<uint32_t*>(in_val + 2 * kBigKMax);
  uint32_t* list = in_idx + 2 * kBigKMax;
  uint32_t* warp_sums = list + a.list_cap;
  uint32_t* misc = warp_sums + 32;
  uint16_t* pay = reinterpret_cast<uint16_t*>(misc + 4);
  const uint32_t tid = threadIdx.x;
  const Grp G{tid, (uint32_t)kBigThreads, 0};
  const uint32_t K = a.c1c2;
  const uint32_t factor = K * K;  // uint32 wrap (:3101)
  const size_t seq_total = (size_t)kNumAnisoDir * kNumDistSeq;

  __shared__ uint32_t s_next;
  uint32_t qi = blockIdx.x;
  if (a.next_query) {
    if (tid == 0) s_next = atomicAdd(a.next_query, 1u);
    __syncthreads();
    qi = s_next;
  }
  for (; qi < a.QN;) {
    __syncthreads();
    if (a.next_query && tid == 0) s_next = atomicAdd(a.next_query, 1u);  // read at the end of the iteration
    const uint32_t q_this = qi;
    for (uint32_t e = tid; e < a.list_cap; e += blockDim.x) list[e] = 0;  // memset of _bins (:3716)
    // ---- selectBinKernel2D2Parts: parts (0,1) and (2,3)
    for (uint32_t pi = 0; pi < 2; pi++) {
      __syncthreads();
      if (tid < 2 * kBigKMax) {
        const uint32_t part = 2 * pi + (tid >> 6), r = tid & 63;
        in_val[tid] = a.top_val[((size_t)qi * 4 + part) * kBigKMax + r];
        in_idx[tid] = a.top_idx[((size_t)qi * 4 + part) * kBigKMax + r];
      }
      __syncthreads();
      if (tid == 0) misc[0] = slope_index(in_val, in_val + kBigKMax, kBigInter);
      __syncthreads();
      if (tid < kBigInter) {
        const uint32_t s = __ldg(a.seq2d + (size_t)misc[0] * kNumDistSeq + tid);
        const uint32_t x = s % kBigDistCluster, y = s / kBigDistCluster;
        float d = 99999999999.f;
        uint32_t b = 0;
        if (x < kBigKMax && y < kBigKMax) {
          d = __fadd_rn(in_val[x], in_val[kBigKMax + y]);
          b = in_idx[x] * K + in_idx[kBigKMax + y];
        }
        l_val[pi * kBigInter + tid] = d;
        r_bin[tid] = b;
        pay[tid] = (uint16_t)tid;
      }
      __syncthreads();
      grp_sort_dispatch(G, l_val + pi * kBigInter, pay, kBigInter);  // bitonic3 over 256 (:2985)
      if (tid < kBigInter) l_idx[pi * kBigInter + tid] = r_bin[pay[tid]];
    }
    __syncthreads();
    // ---- selectBinKernel2DFinal
    if (tid == 0) misc[0] = slope_index(l_val, l_val + kBigInter, 1024);
    __syncthreads();
    const uint32_t slope = misc[0];
    uint32_t n_out = 0, n_elements = 0, n_iter = 0;
    while (n_elements < a.k2 && n_iter < a.max_trials && n_out < a.max_bins) {
      const size_t off = (size_t)slope * kNumDistSeq + (size_t)n_iter * kBigThreads;
      if (off + kBigThreads > seq_total) break;
      {
        const uint32_t s = __ldg(a.seq2d + off + tid);
        const uint32_t x = s % kBigDistCluster, y = s / kBigDistCluster;
        float d = 99999999999.f;
        uint32_t b = 0;
        if (x < kBigInter && y < kBigInter) {
          d = __fadd_rn(l_val[x], l_val[kBigInter + y]);
          b = l_idx[x] * factor + l_idx[kBigInter + y];
        }
        r_dist[tid] = d;
        r_bin[tid] = magicmod(b, a.hash);
        pay[tid] = (uint16_t)tid;
      }
      __syncthreads();
      grp_sort_dispatch(G, r_dist, pay, kBigThreads);  // bitonic3(dist, outIdx, blockDim.x) :3106
      const uint32_t mybin = r_bin[pay[tid]];
      uint32_t start, cnt;
      dir_lookup(a.dir, mybin, start, cnt);
      const uint32_t e = cnt < 2 ? cnt : 2;  // maxVecPB = 2 (:3114)
      uint32_t total;
      const uint32_t ex = block_exscan(e, warp_sums, total);  // ex = inclusive sum of thread tid-1
      uint32_t reg = e;
      if (tid > 0 && (ex + n_elements) >= a.k2) reg = 0;  // :3126-3129
      n_elements += total;
      uint32_t kept_total;
      const uint32_t kex = block_exscan(reg ? 1u : 0u, warp_sums, kept_total);
      if (reg) {
        const uint32_t pos = kex + n_out;  // exclusive scan: 0-based (:3146-3151)
        if (pos < a.max_bins && pos < a.list_cap) list[pos] = mybin;
      }
      if (tid == kBigThreads - 1) misc[1] = kex;  // nElem[blockDim.x - 1] of the exclusive scan
      __syncthreads();
      n_out += misc[1];
      n_iter++;
      __syncthreads();
    }
    __syncthreads();
    const uint32_t nb = n_out < a.max_bins ? n_out : a.max_bins;
    if (a.dbg_bins) {
      for (uint32_t e = tid; e < a.list_cap; e += blockDim.x)
        a.dbg_bins[(size_t)qi * a.list_cap + e] = list[e];
      if (tid == 0) a.dbg_nbins[qi] = nb;
    }
    // ---- Step E1 with maxNVecPerBin = pow2ceil(k) (:6525)
    uint32_t offset = 0;
    uint32_t* cand = a.cand_pos + (size_t)qi * a.max_vec;
    const uint32_t nbl = nb < a.list_cap ? nb : a.list_cap;
    for (uint32_t c0 = 0; c0 < nbl && offset < a.max_vec; c0 += blockDim.x) {
      uint32_t b = c0 + tid;
      uint32_t start = 0, nv = 0;
      if (b < nbl) {
        uint32_t cnt;
        dir_lookup(a.dir, list[b], start, cnt);
        nv = cnt < a.max_vec_per_bin ? cnt : a.max_vec_per_bin;
      }
      uint32_t total;
      uint32_t pos = offset + block_exscan(nv, warp_sums, total);
      if (pos + nv > a.max_vec) nv = (pos >= a.max_vec) ? 0 : (a.max_vec - pos);
      for (uint32_t v = 0; v < nv; v++) cand[pos + v] = start + v;
      offset += total;
    }
    if (tid == 0) a.n_vec[qi] = offset < a.max_vec ? offset : a.max_vec;
    qi = a.next_query ? s_next : q_this + gridDim.x;
  }
}

}  // namespace pqtb
