// rerank_kernels.cuh -- the per-query kernels of Steps D, E1 and E2:
//
//   bins3_kernel / bins4_kernel   Steps D+E1: nibble-packed traversal codes in static visiting
//                   orders, per-query pair tables for the Horner hash, multiply-shift modulo,
//                   bitmap + rank directory; bins3 stops as soon as the candidate list is full
//                   (dense indexes), bins4 walks 16 384 codes at a time (sparse indexes);
//                   bins2_kernel (query_kernels.cuh) serves p > 4
//   adc_stream_kernel   Step E2, distance part: the streaming ADC scan of the split pipeline
//                   (persistent CTA per SM, TMA-staged tables, double-buffered code rows)
//   rank2_kernel    Step E2, ranking + emit: composite-key sort, repair, tie resolver; one CTA
//                   per query, queries drawn from a work counter
//   rerank_kernel   Step E2 fused (scan -> shared memory -> ranking -> first k in one persistent
//                   kernel): short candidate lists and PQT_SCAN_MODE=fused
//   dispatch_kernel / adc_inbox_kernel   multi-GPU: candidates routed to the shard that holds
//                   them; inbox scan, one warp per query, distances stored to the owners
//
// Ranking.  The reference ranks with a bitonic network over max_vec slots
// (pqt/bitonicSort.cuh:16-78).  grp_sort_pairs executes the same compare-exchanges (same
// pairs, same direction rule, strict compares) in registers / shuffles / shared memory, so
// at full width with the reference's 1e7 padding it returns the reference's order including
// ties.  For sparse candidate lists the kernels first sort only pow2ceil(nVec) slots; that
// is the same result unless a distance is >= 1e7 or two different ids have bit-equal
// distances, which is checked, and then the full-width network runs on the restored
// candidate order.  Both paths therefore return exactly what the reference's network returns.
#pragma once
#include "common.cuh"
#include "query_kernels.cuh"
#include "fast_rank.cuh"
#include "tie_resolve.cuh"

namespace pqtb {

// ---- exact x / d, x % d by multiply-shift when a 32-bit magic exists ----------------
struct MagicMod {
  uint32_t d, magic, shift, use64;
  FastMod f64;
};
inline MagicMod make_magicmod(uint32_t d) {
  MagicMod m{};
  m.d = d;
  m.f64 = make_fastmod(d);
  m.use64 = 1;
  if (d <= 1) return m;
  // q = umulhi(x, magic) >> shift is exact for all 32-bit x iff
  // magic*d - 2^(32+shift) <= 2^shift  (Granlund-Montgomery)
  for (uint32_t s = 0; s < 32; s++) {
    unsigned __int128 p = (unsigned __int128)1 << (32 + s);
    unsigned __int128 mg = (p + d - 1) / d;
    if (mg >> 32) continue;
    unsigned __int128 e = mg * d - p;
    if (e <= ((unsigned __int128)1 << s)) {
      m.magic = (uint32_t)mg;
      m.shift = s;
      m.use64 = 0;
      break;
    }
  }
  return m;
}
__device__ __forceinline__ uint32_t magicmod(uint32_t x, const MagicMod& m) {
  if (m.use64) return fastmod(x, m.f64);
  uint32_t q = __umulhi(x, m.magic) >> m.shift;
  return x - q * m.d;
}

// ============================================================================
// Steps D + E1, v2.
// Step D's trial structure only decides where the walk stops, and a stop can only
// happen once max_bins bins are kept -- after which nothing is written any more.
// Net effect: walk the first max_trials*bin_threads traversal
// codes in order, keep the non-empty bins at 1-based slots < max_bins,
// nBins = min(#kept, max_bins).  The kernel walks them in batches of
// kBins2Threads*kProbesPerThread probes (consecutive per thread, so one block scan
// per batch orders them) and stops early once max_bins are kept.
// ============================================================================
struct Bins2Args {
  const uint32_t* idx16;    // [QN][p][16]
  const uint32_t* seq_nib;  // [65536] traversal codes, nibble j = rank in part j; stored per
                            // batch as [r][thread] so that thread-consecutive probes
                            // (t = b0 + thread*16 + r) are read with coalesced loads
  BinDir dir;
  MagicMod hash;
  uint32_t QN, p, c1c2;
  uint32_t n_probes;  // max_trials * bin_threads
  uint32_t max_bins, max_vec_per_bin, max_vec;
  uint32_t* cand_pos;
  uint32_t* n_vec;
  uint32_t* dbg_bins;
  uint32_t* dbg_nbins;
};

constexpr int kBins2Threads = 256;
constexpr int kProbesPerThread = 16;

// dynamic smem: list[max_bins] | pair tables [npairs][256] | warp_sums[32]
__global__ void __launch_bounds__(kBins2Threads) bins2_kernel(Bins2Args a) {
  extern __shared__ uint32_t smem_u[];
  uint32_t* list = smem_u;
  uint32_t* pairs = list + a.max_bins;
  const uint32_t npairs = (a.p + 1) >> 1;
  uint32_t* warp_sums = pairs + npairs * 256;
  const uint32_t K = a.c1c2;
  // multiplier that shifts a value past one pair / one single part in the Horner chain
  const uint32_t K2 = K * K;

  for (uint32_t qi = blockIdx.x; qi < a.QN; qi += gridDim.x) {
    __syncthreads();
    // pair tables: T[pr][r0 | r1 << 4] = idx[2pr][r0] * K + idx[2pr+1][r1]  (uint32 wrap)
    const uint32_t* gi = a.idx16 + (size_t)qi * a.p * 16;
    for (uint32_t e = threadIdx.x; e < npairs * 256; e += blockDim.x) {
      uint32_t pr = e >> 8, r0 = e & 15, r1 = (e >> 4) & 15;
      uint32_t j0 = 2 * pr, j1 = j0 + 1;
      uint32_t v = __ldg(gi + j0 * 16 + r0);
      if (j1 < a.p) v = v * K + __ldg(gi + j1 * 16 + r1);
      pairs[e] = v;
    }
    if (threadIdx.x == 0) list[0] = 0;  // slot 0 keeps the memset value (:3561)
    __syncthreads();

    uint32_t n_out = 0;
    const uint32_t batch = kBins2Threads * kProbesPerThread;
    for (uint32_t b0 = 0; b0 < a.n_probes && n_out < a.max_bins; b0 += batch) {
      const uint32_t t0 = b0 + threadIdx.x * kProbesPerThread;
      uint32_t bins[kProbesPerThread];
      uint32_t words[kProbesPerThread];
#pragma unroll
      for (int r = 0; r < kProbesPerThread; r++) {
        const uint32_t t = t0 + r;
        uint32_t bin = 0, word = 0;
        if (t < a.n_probes) {
          const uint32_t s = __ldg(a.seq_nib + b0 + r * kBins2Threads + threadIdx.x);
          uint32_t o = pairs[s & 0xFF];
          for (uint32_t pr = 1; pr < npairs; pr++) {
            const bool full = (2 * pr + 1) < a.p;
            o = o * (full ? K2 : K) + pairs[pr * 256 + ((s >> (8 * pr)) & 0xFF)];
          }
          bin = magicmod(o, a.hash);
          word = __ldg(a.dir.bitmap + (bin >> 5));
        }
        bins[r] = bin;
        words[r] = word;
      }
      uint32_t keep_mask = 0;
#pragma unroll
      for (int r = 0; r < kProbesPerThread; r++) {
        const bool keep = (t0 + r < a.n_probes) && ((words[r] >> (bins[r] & 31)) & 1u);
        keep_mask |= (keep ? 1u : 0u) << r;
      }
      uint32_t total;
      uint32_t pos = n_out + block_exscan(__popc(keep_mask), warp_sums, total);
#pragma unroll
      for (int r = 0; r < kProbesPerThread; r++) {
        if ((keep_mask >> r) & 1u) {
          pos++;  // inclusive-scan position: 1-based (:3504-3510)
          if (pos < a.max_bins) list[pos] = bins[r];
        }
      }
      n_out += total;
    }
    __syncthreads();
    const uint32_t nb = n_out < a.max_bins ? n_out : a.max_bins;
    if (a.dbg_bins) {
      for (uint32_t e = threadIdx.x; e < a.max_bins; e += blockDim.x)
        a.dbg_bins[(size_t)qi * a.max_bins + e] =
            (e == 0) ? 0u : ((e <= n_out && e < a.max_bins) ? list[e] : 0u);
      if (threadIdx.x == 0) a.dbg_nbins[qi] = nb;
    }

    // ---- Step E1 (:4339-4417)
    uint32_t offset = 0;
    uint32_t* cand = a.cand_pos + (size_t)qi * a.max_vec;
    for (uint32_t b0 = 0; b0 < nb && offset < a.max_vec; b0 += blockDim.x) {
      uint32_t b = b0 + threadIdx.x;
      uint32_t start = 0, nv = 0;
      if (b < nb) {
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
    if (threadIdx.x == 0) a.n_vec[qi] = offset < a.max_vec ? offset : a.max_vec;
  }
}

// ============================================================================
// Steps D + E1, v3 (p <= 4): same walk as bins2_kernel, but inside every batch of
// kBins2Threads * PPT traversal codes the probes are visited in a STATIC order sorted by the
// ranks of all parts but the last.  Probes that share those ranks differ only in the last
// Horner digit (< c1*c2), so their bins fall into one 128-byte line of the bitmap when
// c1*c2 <= 1024: neighbouring lanes then share L1 wavefronts instead of touching 32
// different lines per load (bins2: 77-92 % of the L1 tag bandwidth).  Kept probes are
// put back into traversal order through a per-batch bit array before the ordered
// compaction, so the bin list is identical.
//
// Step E1 runs after every batch over the bins that became final, and the walk stops as soon
// as max_vec candidates are listed: nothing behind that point can reach cand_pos / n_vec (the
// reference's E1 loop ends there, :4339-4417), so the remaining probes -- random sector reads
// of a bitmap that a 1-B index's code traffic keeps evicting from L2 -- are skipped.  On a
// dense index (1 B vectors: most probed bins are occupied) PPT = 4 visits ~2 k of the 8 k
// codes the bin-count stop (max_bins kept) would need.  The debug bin list disables the early
// stop (it wants every kept bin).
// ============================================================================
struct Bins3Args {
  const uint32_t* idx16;       // [QN][p][16]
  const uint32_t* seq_sorted;  // [65536 / batch][batch]: (rank inside the batch << 16) | nibble code
  BinDir dir;
  MagicMod hash;
  uint32_t QN, p, c1c2;
  uint32_t n_probes;
  uint32_t max_bins, max_vec_per_bin, max_vec;
  uint32_t* cand_pos;
  uint32_t* n_vec;
  uint32_t* dbg_bins;
  uint32_t* dbg_nbins;
  // optional (all three or none): the candidates without the repeats.  The same bin can be
  // listed several times (the uint32 Horner hash keeps only idx_0 mod 4 of the first part), so
  // on a dense index about a third of the candidates repeat a code row that is already in the
  // list.  root_pos[q][i], i < n_root[q]: positions of the first occurrences, dense, in list
  // order (what the scan kernel evaluates); ridx[q][a]: index into that array for candidate
  // slot a (what the ranking kernel reads the distance through).
  uint32_t* root_pos;  // [QN][max_vec]
  uint16_t* ridx;      // [QN][max_vec]
  uint32_t* n_root;    // [QN]
  uint32_t* next_query;  // optional work counter (zeroed before the launch): the walk of a query
                         // stops after one to several probe batches, so CTAs draw their queries
};

constexpr int kBins3Batch = kBins2Threads * kProbesPerThread;  // 4096
constexpr uint32_t kDedupSlots = 2048;  // hash slots (key = first position of the bin's run)
constexpr int kBins3FineProbes = 4;                            // probes per thread of the dense-index variant
constexpr int kBins3FineBatch = kBins2Threads * kBins3FineProbes;  // 1024

inline size_t bins3_smem_bytes(uint32_t max_bins, int ppt, bool dedupe) {
  const size_t batch = (size_t)kBins2Threads * ppt;
  return (max_bins + std::max<size_t>(batch, 4 * kBins2Threads) + 2 * 256 + batch / 32 + 32 +
          (dedupe ? (size_t)kDedupSlots + kDedupSlots / 2 : 0)) * 4;
}

// dynamic smem: list[max_bins] | binbuf[max(batch, 512)] | pairs[2][256] | bits[batch/32] | warp_sums[32]
template <int NPAIRS, int PPT>
__global__ void __launch_bounds__(kBins2Threads, 6) bins3_kernel(Bins3Args a) {
  constexpr uint32_t kBatch = kBins2Threads * PPT;
  constexpr uint32_t kBuf = kBatch > 4u * kBins2Threads ? kBatch : 4u * kBins2Threads;
  extern __shared__ uint32_t smem_u[];
  uint32_t* list = smem_u;
  uint32_t* binbuf = list + a.max_bins;
  uint32_t* pairs = binbuf + kBuf;
  uint32_t* bits = pairs + 2 * 256;
  uint32_t* warp_sums = bits + kBatch / 32;
  uint32_t* tkey = warp_sums + 32;    // [kDedupSlots] dedupe table (only when a.root_pos)
  uint16_t* tval = reinterpret_cast<uint16_t*>(tkey + kDedupSlots);  // root-array base of the key's bin
  const bool dedupe = a.root_pos != nullptr;
  const uint32_t K = a.c1c2;
  const uint32_t p = a.p;
  // multiplier that moves the first pair past the second pair (or single last part)
  const uint32_t mul1 = (p == 4) ? K * K : K;
  const uint32_t n_probes = a.n_probes, max_bins = a.max_bins;
  const MagicMod hash = a.hash;
  const uint32_t* __restrict__ bitmap = a.dir.bitmap;

  __shared__ uint32_t s_next;
  uint32_t qi = blockIdx.x;
  if (a.next_query) {
    if (threadIdx.x == 0) s_next = atomicAdd(a.next_query, 1u);
    __syncthreads();
    qi = s_next;
  }
  for (; qi < a.QN;) {
    __syncthreads();
    const uint32_t* gi = a.idx16 + (size_t)qi * p * 16;
    for (uint32_t e = threadIdx.x; e < NPAIRS * 256; e += blockDim.x) {
      uint32_t pr = e >> 8, r0 = e & 15, r1 = (e >> 4) & 15;
      uint32_t j0 = 2 * pr, j1 = j0 + 1;
      uint32_t v = __ldg(gi + j0 * 16 + r0);
      if (j1 < p) v = v * K + __ldg(gi + j1 * 16 + r1);
      pairs[e] = v;
    }
    if (threadIdx.x == 0) {
      list[0] = 0;  // slot 0 keeps the memset value (:3561)
      if (a.next_query) s_next = atomicAdd(a.next_query, 1u);  // read at the end of the iteration
    }
    const uint32_t q_this = qi;
    __syncthreads();

    uint32_t n_out = 0;
    uint32_t offset = 0;     // candidates listed so far (unclipped, as the reference's scan offset)
    uint32_t done_bins = 0;  // list entries Step E1 has consumed
    uint32_t* cand = a.cand_pos + (size_t)qi * a.max_vec;
    uint32_t* s_pos = binbuf;  // E1 staging: the probe buffer is free between batches
    uint32_t* s_start = binbuf + kBins2Threads;
    uint32_t* s_rb = binbuf + 2 * kBins2Threads;
    uint32_t n_root = 0;  // first-occurrence candidates listed so far
    uint32_t* rootp = dedupe ? a.root_pos + (size_t)qi * a.max_vec : nullptr;
    uint16_t* ridx = dedupe ? a.ridx + (size_t)qi * a.max_vec : nullptr;
    if (dedupe) {
      for (uint32_t e = threadIdx.x; e < kDedupSlots; e += blockDim.x) tkey[e] = 0xFFFFFFFFu;
    }
    // ---- Step E1 (:4339-4417) over list entries [done_bins, upto).  The candidate slots of a
    // chunk of bins are written by the whole CTA (slot -> bin by binary search over the chunk's
    // exclusive scan): coalesced stores, and a bin with hundreds of vectors does not serialise
    // one thread.  Repeated bins are recognised by the first position of their run (unique per
    // non-empty bin) in a small hash table; tval holds the root-array base of the key's bin.
    auto step_e1 = [&](uint32_t upto) {
      for (uint32_t c0 = done_bins; c0 < upto && offset < a.max_vec; c0 += blockDim.x) {
        uint32_t b = c0 + threadIdx.x;
        uint32_t start = 0, nv = 0;
        if (b < upto) {
          uint32_t cnt;
          dir_lookup(a.dir, list[b], start, cnt);
          nv = cnt < a.max_vec_per_bin ? cnt : a.max_vec_per_bin;
        }
        uint32_t total;
        const uint32_t pos = offset + block_exscan(nv, warp_sums, total);
        s_pos[threadIdx.x] = pos;
        s_start[threadIdx.x] = start;
        if (dedupe) {
          const uint32_t nvc = pos >= a.max_vec ? 0u : (pos + nv > a.max_vec ? a.max_vec - pos : nv);
          // find-or-insert: the thread whose compare-and-swap installs the key owns the first
          // occurrence (which of several repeats inside one chunk wins is immaterial: they hold
          // the same code rows).  A bin cut short by the max_vec budget never registers -- a
          // repeat could need more rows than it lists -- and a key without a free slot is
          // treated as a first occurrence.
          uint32_t ts = 0xFFFFFFFFu;
          bool first = true;
          if (nvc) {
            uint32_t hh = (start * 2654435761u) >> 21;  // 11 bits
#pragma unroll 1
            for (uint32_t probe = 0; probe < 4u; probe++, hh = (hh + 1u) & (kDedupSlots - 1u)) {
              uint32_t kx = tkey[hh];
              if (kx == 0xFFFFFFFFu && nvc == nv) kx = atomicCAS(&tkey[hh], 0xFFFFFFFFu, start);
              if (kx == 0xFFFFFFFFu) {  // installed (or, for a clipped bin, not present)
                if (nvc == nv) ts = hh;
                break;
              }
              if (kx == start) {
                ts = hh;
                first = false;
                break;
              }
            }
          }
          uint32_t rtotal;
          uint32_t rb = n_root + block_exscan((nvc && first) ? nvc : 0u, warp_sums, rtotal);
          if (nvc && first && ts != 0xFFFFFFFFu) tval[ts] = (uint16_t)rb;
          __syncthreads();
          if (nvc && !first) rb = tval[ts];
          s_rb[threadIdx.x] = rb | (first ? 0x80000000u : 0u);
          n_root += rtotal;
        }
        __syncthreads();
        const uint32_t hi = offset + total < a.max_vec ? offset + total : a.max_vec;
        for (uint32_t slot = offset + threadIdx.x; slot < hi; slot += blockDim.x) {
          // the bin that holds this slot = the last one whose first slot is <= slot
          uint32_t j = 0;
#pragma unroll
          for (uint32_t step = kBins2Threads >> 1; step > 0; step >>= 1)
            if (s_pos[j + step] <= slot) j += step;
          const uint32_t off = slot - s_pos[j];
          const uint32_t cpos = s_start[j] + off;
          cand[slot] = cpos;
          if (dedupe) {
            const uint32_t r = s_rb[j];
            const uint32_t ri = (r & 0x7FFFFFFFu) + off;
            ridx[slot] = (uint16_t)ri;
            if (r >> 31) rootp[ri] = cpos;
          }
        }
        offset += total;
        __syncthreads();
      }
      done_bins = upto;
    };

    for (uint32_t b0 = 0; b0 < n_probes && n_out < max_bins; b0 += kBatch) {
      if (threadIdx.x < kBatch / 32) bits[threadIdx.x] = 0;
      __syncthreads();
      // only the loaded bitmap words (and the bit position inside them, 5 bits each) stay in
      // registers while the probes are in flight; bins of the few kept probes are
      // recomputed -- keeps the kernel at 6 CTAs per SM
      uint32_t words[PPT];
      uint32_t sh[(PPT + 3) / 4];
#pragma unroll
      for (int r = 0; r < (PPT + 3) / 4; r++) sh[r] = 0;
#pragma unroll
      for (int r = 0; r < PPT; r++) {
        const uint32_t ent = __ldg(a.seq_sorted + b0 + r * kBins2Threads + threadIdx.x);
        uint32_t o = pairs[ent & 0xFF];
        if (NPAIRS == 2) o = o * mul1 + pairs[256 + ((ent >> 8) & 0xFF)];
        const uint32_t bin = magicmod(o, hash);
        sh[r >> 2] |= (bin & 31u) << (8 * (r & 3));
        words[r] = (b0 + (ent >> 16) < n_probes) ? __ldg(bitmap + (bin >> 5)) : 0u;
      }
#pragma unroll
      for (int r = 0; r < PPT; r++) {
        if ((words[r] >> ((sh[r >> 2] >> (8 * (r & 3))) & 31u)) & 1u) {
          const uint32_t ent = __ldg(a.seq_sorted + b0 + r * kBins2Threads + threadIdx.x);
          const uint32_t u = ent >> 16;
          uint32_t o = pairs[ent & 0xFF];
          if (NPAIRS == 2) o = o * mul1 + pairs[256 + ((ent >> 8) & 0xFF)];
          atomicOr(&bits[u >> 5], 1u << (u & 31));
          binbuf[u] = magicmod(o, hash);
        }
      }
      __syncthreads();
      // ordered compaction in traversal order: thread i owns 32 consecutive probes
      uint32_t w = (threadIdx.x < kBatch / 32) ? bits[threadIdx.x] : 0u;
      uint32_t total;
      uint32_t pos = n_out + block_exscan(__popc(w), warp_sums, total);
      while (w) {
        const uint32_t bit = __ffs(w) - 1;
        w &= w - 1;
        pos++;  // inclusive-scan position: 1-based (:3504-3510)
        if (pos < max_bins) list[pos] = binbuf[threadIdx.x * 32 + bit];
      }
      n_out += total;
      __syncthreads();
      if (!a.dbg_bins) {
        // entries below min(n_out, max_bins) are final whatever the rest of the walk keeps
        step_e1(n_out < max_bins ? n_out : max_bins);
        if (offset >= a.max_vec) break;
      }
    }
    __syncthreads();
    const uint32_t nb = n_out < max_bins ? n_out : max_bins;
    if (a.dbg_bins) {
      for (uint32_t e = threadIdx.x; e < max_bins; e += blockDim.x)
        a.dbg_bins[(size_t)qi * max_bins + e] =
            (e == 0) ? 0u : ((e <= n_out && e < max_bins) ? list[e] : 0u);
      if (threadIdx.x == 0) a.dbg_nbins[qi] = nb;
    }
    step_e1(nb);
    if (threadIdx.x == 0) {
      a.n_vec[qi] = offset < a.max_vec ? offset : a.max_vec;
      if (dedupe) a.n_root[qi] = n_root;
    }
    qi = a.next_query ? s_next : q_this + gridDim.x;
  }
}

// ============================================================================
// Steps D + E1, v4 (p <= 4): as bins3_kernel, but the static visiting order is sorted over
// 16384 traversal codes at once.  The 16 probes that share the ranks of all parts but the
// last hit one 128-byte line of the bitmap; the traversal order spreads them over its
// 4096-code batches, so bins3 touches that line in up to four batches (7429 line visits per
// query) where one visit suffices (2913).  The walk is bound by L2 sector traffic, so this is
// what counts; the price is that the early stop (max_bins kept) is only checked every
// 16384 codes -- same output, more probes on indexes dense enough to stop early.
// Kept probes are flagged in a bit array indexed by traversal position; their bins are
// recomputed from the plain traversal table when the list is compacted (few are kept).
// ============================================================================
struct Bins4Args {
  const uint32_t* idx16;      // [QN][p][16]
  const uint32_t* seq_mega;   // [4 mega-batches][16384]: (position inside the mega-batch << 16) | nibble code
  const uint32_t* seq_plain;  // [65536] nibble codes in traversal order
  BinDir dir;
  MagicMod hash;
  uint32_t QN, p, c1c2;
  uint32_t n_probes;
  uint32_t max_bins, max_vec_per_bin, max_vec;
  uint32_t* cand_pos;
  uint32_t* n_vec;
  uint32_t* dbg_bins;
  uint32_t* dbg_nbins;
};

constexpr int kBins4Mega = 4 * kBins3Batch;  // 16384

// dynamic smem: list[max_bins] | bits[512] | pairs[2][256] | stage[2][256] | warp_sums[32]
template <int NPAIRS>
__global__ void __launch_bounds__(kBins2Threads, 6) bins4_kernel(Bins4Args a) {
  extern __shared__ uint32_t smem_u[];
  uint32_t* list = smem_u;
  uint32_t* bits = list + a.max_bins;
  uint32_t* pairs = bits + kBins4Mega / 32;
  uint32_t* s_pos = pairs + 2 * 256;
  uint32_t* s_start = s_pos + kBins2Threads;
  uint32_t* warp_sums = s_start + kBins2Threads;
  const uint32_t K = a.c1c2;
  const uint32_t p = a.p;
  const uint32_t mul1 = (p == 4) ? K * K : K;
  const uint32_t n_probes = a.n_probes, max_bins = a.max_bins;
  const MagicMod hash = a.hash;
  const uint32_t* __restrict__ bitmap = a.dir.bitmap;

  for (uint32_t qi = blockIdx.x; qi < a.QN; qi += gridDim.x) {
    __syncthreads();
    const uint32_t* gi = a.idx16 + (size_t)qi * p * 16;
    for (uint32_t e = threadIdx.x; e < NPAIRS * 256; e += blockDim.x) {
      uint32_t pr = e >> 8, r0 = e & 15, r1 = (e >> 4) & 15;
      uint32_t j0 = 2 * pr, j1 = j0 + 1;
      uint32_t v = __ldg(gi + j0 * 16 + r0);
      if (j1 < p) v = v * K + __ldg(gi + j1 * 16 + r1);
      pairs[e] = v;
    }
    if (threadIdx.x == 0) list[0] = 0;  // slot 0 keeps the memset value (:3561)
    __syncthreads();

    uint32_t n_out = 0;
    for (uint32_t m0 = 0; m0 < n_probes && n_out < max_bins; m0 += kBins4Mega) {
      bits[threadIdx.x] = 0;
      bits[threadIdx.x + kBins2Threads] = 0;
      __syncthreads();
#pragma unroll 1
      for (uint32_t b0 = m0; b0 < m0 + kBins4Mega; b0 += kBins3Batch) {
        uint32_t words[kProbesPerThread];
        uint32_t sh[kProbesPerThread / 4];
#pragma unroll
        for (int r = 0; r < kProbesPerThread / 4; r++) sh[r] = 0;
#pragma unroll
        for (int r = 0; r < kProbesPerThread; r++) {
          const uint32_t ent = __ldg(a.seq_mega + b0 + r * kBins2Threads + threadIdx.x);
          uint32_t o = pairs[ent & 0xFF];
          if (NPAIRS == 2) o = o * mul1 + pairs[256 + ((ent >> 8) & 0xFF)];
          const uint32_t bin = magicmod(o, hash);
          sh[r >> 2] |= (bin & 31u) << (8 * (r & 3));
          words[r] = (m0 + (ent >> 16) < n_probes) ? __ldg(bitmap + (bin >> 5)) : 0u;
        }
#pragma unroll
        for (int r = 0; r < kProbesPerThread; r++) {
          if ((words[r] >> ((sh[r >> 2] >> (8 * (r & 3))) & 31u)) & 1u) {
            const uint32_t u = __ldg(a.seq_mega + b0 + r * kBins2Threads + threadIdx.x) >> 16;
            atomicOr(&bits[u >> 5], 1u << (u & 31));
          }
        }
      }
      __syncthreads();
      // ordered compaction in traversal order: thread i owns 64 consecutive positions
      uint32_t w0 = bits[2 * threadIdx.x], w1 = bits[2 * threadIdx.x + 1];
      uint32_t total;
      uint32_t pos = n_out + block_exscan(__popc(w0) + __popc(w1), warp_sums, total);
#pragma unroll
      for (int half = 0; half < 2; half++) {
        uint32_t w = half ? w1 : w0;
        while (w) {
          const uint32_t bit = __ffs(w) - 1;
          w &= w - 1;
          pos++;  // inclusive-scan position: 1-based (:3504-3510)
          if (pos < max_bins) {
            const uint32_t code = __ldg(a.seq_plain + m0 + threadIdx.x * 64 + half * 32 + bit);
            uint32_t o = pairs[code & 0xFF];
            if (NPAIRS == 2) o = o * mul1 + pairs[256 + ((code >> 8) & 0xFF)];
            list[pos] = magicmod(o, hash);
          }
        }
      }
      n_out += total;
      __syncthreads();
    }
    __syncthreads();
    const uint32_t nb = n_out < max_bins ? n_out : max_bins;
    if (a.dbg_bins) {
      for (uint32_t e = threadIdx.x; e < max_bins; e += blockDim.x)
        a.dbg_bins[(size_t)qi * max_bins + e] =
            (e == 0) ? 0u : ((e <= n_out && e < max_bins) ? list[e] : 0u);
      if (threadIdx.x == 0) a.dbg_nbins[qi] = nb;
    }

    // ---- Step E1 (:4339-4417), as in bins3_kernel
    uint32_t offset = 0;
    uint32_t* cand = a.cand_pos + (size_t)qi * a.max_vec;
    for (uint32_t c0 = 0; c0 < nb && offset < a.max_vec; c0 += blockDim.x) {
      uint32_t b = c0 + threadIdx.x;
      uint32_t start = 0, nv = 0;
      if (b < nb) {
        uint32_t cnt;
        dir_lookup(a.dir, list[b], start, cnt);
        nv = cnt < a.max_vec_per_bin ? cnt : a.max_vec_per_bin;
      }
      uint32_t total;
      const uint32_t pos = offset + block_exscan(nv, warp_sums, total);
      s_pos[threadIdx.x] = pos;
      s_start[threadIdx.x] = start;
      __syncthreads();
      const uint32_t hi = offset + total < a.max_vec ? offset + total : a.max_vec;
      for (uint32_t slot = offset + threadIdx.x; slot < hi; slot += blockDim.x) {
        uint32_t j = 0;
#pragma unroll
        for (uint32_t step = kBins2Threads >> 1; step > 0; step >>= 1)
          if (s_pos[j + step] <= slot) j += step;
        cand[slot] = s_start[j] + (slot - s_pos[j]);
      }
      offset += total;
      __syncthreads();
    }
    if (threadIdx.x == 0) a.n_vec[qi] = offset < a.max_vec ? offset : a.max_vec;
  }
}

// ============================================================================
// Thread group = a contiguous set of warps of a CTA that works on one query and
// synchronises on its own named barrier (bar.sync id, n).  rerank_kernel runs two
// groups of 512 threads per CTA on different queries so that the scan phase of one
// overlaps the ranking phase of the other; rank2_kernel uses one group = the CTA.
// ============================================================================
constexpr uint32_t kPayPad = 0xFFFFu;  // payload of a padded sort slot (candidate positions < 4096)

// the reference's network (pqt/bitonicSort.cuh:16-78) over n elements, run by a group
__device__ __forceinline__ void bitonic_smem_grp(const Grp& g, float* val, uint16_t* idx, uint32_t n) {
  const uint32_t half = n >> 1;
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = g.t; t < half; t += g.n) {
        const uint32_t i = 2 * t - (t & (j - 1)), b = i + j;
        const float va = val[i], vb = val[b];
        const uint32_t ia = idx[i], ib = idx[b];
        const bool sw = ((i & k) == 0) ? (va > vb) : (va < vb);
        val[i] = sw ? vb : va;
        val[b] = sw ? va : vb;
        idx[i] = (uint16_t)(sw ? ib : ia);
        idx[b] = (uint16_t)(sw ? ia : ib);
      }
      g.sync();
    }
  }
}

// ============================================================================
// Group-wide ascending sort of n2 (power of two, >= 32) (val, pay) pairs held in shared
// memory.  A thread works on blocks of E consecutive elements in registers:
// compare-exchange distance j < E runs in registers, E <= j < 32E on shuffles, j >= 32E
// through shared memory by the whole group.  When n2 > E * group size the group makes
// several passes (the register stages of different blocks are independent), which keeps
// the register footprint at E elements.  Any correct sort would do here (see the header
// comment); the bitonic schedule is used because it needs no data-dependent control.
// Every thread of the group must call.
// ============================================================================
template <int E>
__device__ __forceinline__ void grp_sort_pairs(const Grp& g, float* sv, uint16_t* sp, uint32_t n2) {
  const uint32_t t = g.t, lane = t & 31;
  const uint32_t per = E * g.n;  // elements one pass of the group covers
  const uint32_t k1 = n2 < 32u * E ? n2 : 32u * E;
  // ---- phase 1: every k whose largest distance stays inside a warp's 32E elements
  for (uint32_t b0 = 0; b0 < n2; b0 += per) {
    const uint32_t base = b0 + t * E;
    if (base < n2) {
      float v[E];
      uint32_t p[E];
#pragma unroll
      for (int r = 0; r < E; r++) {
        v[r] = sv[base + r];
        p[r] = sp[base + r];
      }
      for (uint32_t k = 2; k <= k1; k <<= 1) sort_reg_stages<E>(v, p, base, lane, k, k >> 1);
#pragma unroll
      for (int r = 0; r < E; r++) {
        sv[base + r] = v[r];
        sp[base + r] = (uint16_t)p[r];
      }
    }
  }
  g.sync();
  // ---- phase 2: larger k: cross-warp distances in shared memory, the rest in registers
  for (uint32_t k = 64u * E; k <= n2; k <<= 1) {
    for (uint32_t j = k >> 1; j >= 32u * E; j >>= 1) {
      for (uint32_t e = t; e < (n2 >> 1); e += g.n) {
        const uint32_t i = 2 * e - (e & (j - 1)), b = i + j;
        const float va = sv[i], vb = sv[b];
        const uint32_t pa = sp[i], pb = sp[b];
        const bool sw = ((i & k) == 0) ? (va > vb) : (va < vb);
        sv[i] = sw ? vb : va;
        sv[b] = sw ? va : vb;
        sp[i] = (uint16_t)(sw ? pb : pa);
        sp[b] = (uint16_t)(sw ? pa : pb);
      }
      g.sync();
    }
    for (uint32_t b0 = 0; b0 < n2; b0 += per) {
      const uint32_t base = b0 + t * E;
      if (base < n2) {
        float v[E];
        uint32_t p[E];
#pragma unroll
        for (int r = 0; r < E; r++) {
          v[r] = sv[base + r];
          p[r] = sp[base + r];
        }
        sort_reg_stages<E>(v, p, base, lane, k, 16u * E);
#pragma unroll
        for (int r = 0; r < E; r++) {
          sv[base + r] = v[r];
          sp[base + r] = (uint16_t)p[r];
        }
      }
    }
    g.sync();
  }
}

// E = 4 elements per thread once there is enough work for the group; wider inputs take
// several passes instead of more registers
__device__ __forceinline__ void grp_sort_dispatch(const Grp& g, float* sv, uint16_t* sp, uint32_t n2) {
  if (n2 <= g.n)
    grp_sort_pairs<1>(g, sv, sp, n2);
  else if (n2 <= 2 * g.n)
    grp_sort_pairs<2>(g, sv, sp, n2);
  else
    grp_sort_pairs<4>(g, sv, sp, n2);
}

// Ranks the candidates of one query held in shared memory and writes the first k.
//   s_val[a], s_id[a] for a < nv: ADC distance / vector id in candidate order
//   s_pay: scratch [max_vec]; s_flag: one word.  g.n >= 256 (max_vec <= 4096).
//
// grp_sort_pairs executes exactly the compare-exchanges of the reference's network
// (pqt/bitonicSort.cuh:16-78: pairs (i, i^j), ascending iff (i & k) == 0, strict compare), so
// run over all max_vec slots with the reference's padding (1e7) it returns the reference's
// order INCLUDING ties.  Sparse candidate lists are first sorted at pow2ceil(nVec) width
// (pads +inf); that shortcut is valid unless a distance is >= 1e7 or two different ids tie,
// in which case the candidate order is restored and the full-width network runs.
template <class IdOf>
__device__ __forceinline__ void rank_and_emit(const Grp& g, float* s_val, uint16_t* s_pay,
                                              IdOf id_of, uint32_t* s_flag, uint32_t nv,
                                              uint32_t max_vec, uint32_t k, float* out_dist,
                                              uint32_t* out_idx, unsigned long long* exact_counter) {
  const uint32_t t = g.t;
  uint32_t n2 = pow2ceil(nv < 32 ? 32 : nv);
  if (n2 > max_vec) n2 = max_vec;  // max_vec < 32: tiny widths
  const bool full = (n2 == max_vec);
  // callers arrive here right after reading *s_flag (the verdict of the fast path): nobody may
  // clear it before every thread of the group has read it
  g.sync();
  if (t == 0) *s_flag = 0;
  g.sync();
  // payload = candidate position
  uint32_t bad = 0;
  for (uint32_t e = t; e < n2; e += g.n) {
    if (e < nv) {
      s_pay[e] = (uint16_t)e;
      const float v = s_val[e];
      if (!(fabsf(v) < __int_as_float(0x7f800000))) bad |= 2u;  // NaN / inf: min/max compare-exchange
      if (!(v < kPadDist)) bad |= 1u;                          // pad order matters
    } else {
      s_val[e] = full ? kPadDist : __int_as_float(0x7f800000);
      s_pay[e] = kPayPad;
    }
  }
  if (bad) atomicOr(s_flag, bad);
  g.sync();
  const bool nonfinite = (*s_flag & 2u) != 0;
  if (n2 >= 32 && !nonfinite) grp_sort_dispatch(g, s_val, s_pay, n2);
  if (!full && !nonfinite) {
    // ties between copies of the same vector are harmless (the same bin can be listed more
    // than once: the uint32 Horner hash keeps only idx_0 mod 4 of the first part); ties
    // between different ids expose the network's order
    for (uint32_t e = t + 1; e < nv; e += g.n)
      if (s_val[e] == s_val[e - 1] && id_of(s_pay[e]) != id_of(s_pay[e - 1])) atomicOr(s_flag, 1u);
  }
  g.sync();
  const uint32_t flag = *s_flag;
  if (n2 < 32 || nonfinite || (!full && flag)) {
    if (t == 0 && exact_counter) atomicAdd(exact_counter, 1ull);
    // restore candidate order, pad to max_vec with 1e7, run the full-width network (:5331-5340)
    if (!nonfinite && n2 >= 32) {
      float rv[16];
      uint32_t rp[16];
      int cnt = 0;
      for (uint32_t e = t; e < n2 && cnt < 16; e += g.n, cnt++) {
        rv[cnt] = s_val[e];
        rp[cnt] = s_pay[e];
      }
      g.sync();
      for (int c = 0; c < cnt; c++)
        if (rp[c] != kPayPad) s_val[rp[c]] = rv[c];
      g.sync();
    }
    for (uint32_t e = t; e < max_vec; e += g.n) {
      if (e >= nv) s_val[e] = kPadDist;
      s_pay[e] = e < nv ? (uint16_t)e : kPayPad;
    }
    g.sync();
    if (nonfinite || max_vec < 32)
      bitonic_smem_grp(g, s_val, s_pay, max_vec);  // literal strict-compare network
    else
      grp_sort_dispatch(g, s_val, s_pay, max_vec);
    for (uint32_t e = t; e < k; e += g.n) {
      const uint32_t a = s_pay[e];
      out_dist[e] = s_val[e];
      out_idx[e] = (a == kPayPad) ? kPadIdx : id_of(a);
    }
  } else if (full) {
    for (uint32_t e = t; e < k; e += g.n) {
      const uint32_t a = s_pay[e];
      out_dist[e] = s_val[e];
      out_idx[e] = (a == kPayPad) ? kPadIdx : id_of(a);
    }
  } else {
    for (uint32_t e = t; e < k; e += g.n) {
      if (e < nv) {
        out_dist[e] = s_val[e];
        out_idx[e] = id_of(s_pay[e]);
      } else {
        out_dist[e] = kPadDist;
        out_idx[e] = kPadIdx;
      }
    }
  }
  g.sync();
}

// ============================================================================
// Fused Step E2: scan + rank + emit.  Same scan arithmetic and lane mapping as
// adc_scan_kernel (see there); single-GPU only (all candidates are local).
// One persistent CTA per SM = two groups of 512 threads, each walking its own queries:
// shared cbd table, per-group LUT double buffer (TMA bulk copies) and candidate arrays.
// ============================================================================
struct RerankArgs {
  ScanArgs s;
  uint32_t k;
  float* out_dist;    // [QN][k]
  uint32_t* out_idx;  // [QN][k]
  unsigned long long* exact_counter;  // queries ranked by the exact network (may be null)
  unsigned long long* tie_counter;    // queries whose ties were re-ordered by tie_resolve (may be null)
  uint32_t fast_rank;                 // 1: composite-key sort first (fast_rank.cuh); 0: network only
  uint32_t dedupe;                    // 1: evaluate repeated candidates once (see rerank_kernel)
  unsigned long long* phase_dbg;      // optional [QN][8] clock64 stamps per query (debug)
  uint32_t* next_query;               // work counter (zeroed before the launch): the thread groups
                                      // draw queries from it, so a slow query does not hold up a
                                      // statically assigned tail
};

constexpr int kRerankGroups = 2;  // rank2_kernel / default configuration
constexpr int kRerankGroupThreads = kScanThreads / kRerankGroups;  // 512

// NG thread groups per CTA (each on its own query): more groups hide more of the per-query
// latency chains (code loads -> scan -> sort -> emit) but need NG candidate arrays in shared
// memory.  CREP: the c^2 table has 32-float rows (LP values replicated 32/LP times, conflict
// free, c1*c1*128 bytes); otherwise its rows are the LP values only (the canonical cbDist
// layout, c1*c1*LP*4 bytes; the 32/LP candidates of a warp step may collide on a bank).
inline size_t rerank_smem_bytes(uint32_t c1, uint32_t LP, uint32_t max_vec, int NG, bool crep) {
  return ((size_t)c1 * c1 * (crep ? 32 : LP) + (size_t)NG * 2 * c1 * 32) * 4 +
         (size_t)NG * ((size_t)8 * max_vec + 512 + 4 + 32 + 16) + 64 + 16;
}

// TPB threads per CTA = NG groups of TPB / NG threads.  1024 threads leave 64 registers per
// thread; the lineparts = 32 configuration runs 512 (two groups of 256 = the 256 sorter threads
// of a 4096-wide ranking) so that a warp can hold the code rows of TWO warp steps: the loads of
// the next step are in flight while the current one is evaluated.
template <int LP, int NG, bool CREP, int TPB = kScanThreads>
__global__ void __launch_bounds__(TPB, 1) rerank_kernel(RerankArgs g) {
  const ScanArgs& a = g.s;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr uint32_t kRerankGroups = NG;
  constexpr uint32_t kRerankGroupThreads = TPB / NG;
  constexpr uint32_t CROW = CREP ? 32u : (uint32_t)LP;  // floats per row of the c^2 table
  const uint32_t lut_floats = a.c1 * 32;
  const uint32_t cbd_floats = a.c1 * a.c1 * CROW;
  const uint32_t grp = threadIdx.x / kRerankGroupThreads;
  Grp G{threadIdx.x - grp * kRerankGroupThreads, (uint32_t)kRerankGroupThreads, 1 + grp};

  float* s_cbd = reinterpret_cast<float*>(smem_raw);
  float* s_luts = s_cbd + cbd_floats;                                  // [groups][2][lut_floats]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_luts + kRerankGroups * 2 * lut_floats);  // [groups][2] + cbd
  uint32_t* s_flags = reinterpret_cast<uint32_t*>(bars + 2 * kRerankGroups + 2);
  uint32_t* s_red = s_flags + kRerankGroups;                           // [groups][8]: umin, umax, bad, next query, roots
  uint32_t* s_fixes = s_red + 8 * kRerankGroups;                       // [groups][128]: emit bitmap
  float* s_arr = reinterpret_cast<float*>(                               // [groups][2][max_vec], 16-byte aligned
      (reinterpret_cast<uintptr_t>(s_fixes + 128 * kRerankGroups) + 15u) & ~(uintptr_t)15u);
  float* s_lut0 = s_luts + grp * 2 * lut_floats;
  float* s_lut1 = s_lut0 + lut_floats;
  uint64_t* gbar = bars + 2 * grp;
  uint64_t* cbar = bars + 2 * kRerankGroups;
  uint32_t* s_flag = s_flags + grp;
  uint32_t* s_min = s_red + 8 * grp;
  uint32_t* s_max = s_min + 1;
  uint32_t* s_bad = s_min + 2;
  uint32_t* s_fix = s_fixes + 128 * grp;
  uint32_t* s_q = s_min + 3;    // next query of this group
  uint32_t* s_cnt = s_min + 4;  // first occurrences among the candidates of the current query
  // per group: val f32[max_vec] | scratch u32[max_vec] (composite sort words; the (u16) payload
  // array of the exact network aliases it).  Vector ids are not staged: they are read once,
  // when a result is emitted.
  float* s_val = s_arr + (size_t)grp * 2 * a.max_vec;
  uint32_t* s_cmp = reinterpret_cast<uint32_t*>(s_val + a.max_vec);
  uint16_t* s_pay = reinterpret_cast<uint16_t*>(s_cmp);

  const uint32_t lane = G.t & 31, warp = G.t >> 5, nwarps = G.n >> 5;
  const uint32_t lp = lane & (LP - 1);
  const uint32_t cbd_b = smem_u32(s_cbd) + (CREP ? lane : lp) * 4u;
  const uint32_t* __restrict__ codes_lp = a.codes + lp;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < 2 * kRerankGroups + 1; i++) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t cbd_bytes = cbd_floats * 4;
    mbar_expect_tx(cbar, cbd_bytes);
    for (uint32_t off = 0; off < cbd_bytes; off += 32768) {
      uint32_t n = cbd_bytes - off < 32768 ? cbd_bytes - off : 32768;
      tma_bulk_g2s(reinterpret_cast<unsigned char*>(s_cbd) + off,
                   reinterpret_cast<const unsigned char*>(a.cbd_dup) + off, n, cbar);
    }
  }
  // first query of this group
  if (G.t == 0) {
    const uint32_t q0 = atomicAdd(g.next_query, 1u);
    *s_q = q0;
    if (q0 < a.QN) {
      mbar_expect_tx(&gbar[0], lut_floats * 4);
      tma_bulk_g2s(s_lut0, a.lut_dup + (size_t)q0 * lut_floats, lut_floats * 4, &gbar[0]);
    }
  }
  mbar_wait(cbar, 0);
  G.sync();
  uint32_t qi = *s_q;

  uint32_t buf = 0, phase0 = 0, phase1 = 0;
  while (qi < a.QN) {
    G.sync();  // everyone has read *s_q
    if (G.t == 0) {
      // draw the next query and start the copy of its LUT into the other buffer
      const uint32_t qn = atomicAdd(g.next_query, 1u);
      *s_q = qn;
      if (qn < a.QN) {
        uint64_t* nb = &gbar[buf ^ 1];
        mbar_expect_tx(nb, lut_floats * 4);
        tma_bulk_g2s(buf ? s_lut0 : s_lut1, a.lut_dup + (size_t)qn * lut_floats, lut_floats * 4, nb);
      }
      *s_min = 0xFFFFFFFFu;
      *s_max = 0u;
      *s_bad = 0u;
      *s_flag = 0u;
    }
    if (G.t < 128u) s_fix[G.t] = 0u;  // repair bitmap of the ranking pass
    const float* s_lut = buf ? s_lut1 : s_lut0;
    unsigned long long* ph = g.phase_dbg ? g.phase_dbg + (size_t)qi * 8 : nullptr;
    if (ph && G.t == 0) ph[0] = clock64();
    mbar_wait(&gbar[buf], buf ? phase1 : phase0);
    if (buf)
      phase1 ^= 1;
    else
      phase0 ^= 1;

    const uint32_t nv = min(__ldg(a.n_vec + qi), a.max_vec);
    const uint32_t* cand = a.cand_pos + (size_t)qi * a.max_vec;
    if (ph && G.t == 0) ph[1] = clock64();
    G.sync();
    const uint32_t q_next = *s_q;
    const uint32_t lut_b = smem_u32(s_lut) + lane * 4u;
    uint32_t umin = 0xFFFFFFFFu, umax = 0u, bad = 0u;
    auto note = [&](float v) {
      const uint32_t u = sortable_key(v);
      umin = min(umin, u);
      umax = max(umax, u);
      // the fast path needs finite distances below the pad value
      if (!(v < kPadDist) || !(v > -__int_as_float(0x7f800000))) bad = 1u;
    };
    // Evaluates the candidates slot_of(0) .. slot_of(M-1), 32 per warp step, software pipelined:
    // the code rows of step k+1 are requested before step k is evaluated (two register sets),
    // and the positions are fetched two steps ahead.  Lanes past the end work on element 0 (a
    // valid row, result unused).
    auto run_scan = [&](uint32_t M, auto slot_of) {
      const uint32_t stride = nwarps * 32;
      auto fetch = [&](uint32_t e, uint32_t& ca, uint32_t& pos) {
        ca = slot_of(e < M ? e : 0u);
        pos = __ldg(cand + ca);
      };
      auto load = [&](uint32_t (&w)[LP], uint32_t pos) {
        // the id is read when the result is emitted: pull its sector into L2 now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ids + pos));
        adc_load_rows<LP, false>(w, pos, codes_lp, lp, nullptr);
      };
      uint32_t e0 = warp * 32 + lane;
      if (warp * 32 >= M) return;  // warp-uniform
      uint32_t wA[LP], wB[LP];
      uint32_t ca0, pos0, ca1, pos1, ca2, pos2;
      fetch(e0, ca0, pos0);
      fetch(e0 + stride, ca1, pos1);
      load(wA, pos0);
      for (uint32_t base = warp * 32; base < M; base += 2 * stride) {
        const bool has1 = base + stride < M, has2 = base + 2 * stride < M;
        if (has1) load(wB, pos1);
        fetch(e0 + 2 * stride, ca2, pos2);
        {
          const float v = adc_eval_rows<LP, CROW>(wA, lut_b, cbd_b, a.c1, lp);
          if (e0 < M) s_val[ca0] = v;
        }
        if (!has1) break;
        if (has2) load(wA, pos2);
        uint32_t ca3, pos3;
        fetch(e0 + 3 * stride, ca3, pos3);
        {
          const float v = adc_eval_rows<LP, CROW>(wB, lut_b, cbd_b, a.c1, lp);
          if (e0 + stride < M) s_val[ca1] = v;
        }
        e0 += 2 * stride;
        ca0 = ca2;
        pos0 = pos2;
        ca1 = ca3;
        pos1 = pos3;
      }
    };
    // The same bin can be listed several times (the uint32 Horner hash keeps only idx_0 mod 4
    // of the first part), so about a third of the candidates of a dense list repeat a code row
    // that is already in the list.  Repeats are found with a small hash table over the
    // positions (the sort scratch is free during the scan), only the first occurrences are
    // evaluated, and the repeats copy their distance -- the same bits the evaluation would
    // give.  Detection may miss repeats (two table slots per position, no chaining): they are
    // simply evaluated again.
    constexpr uint32_t CPT = 4096u / kRerankGroupThreads;  // candidates per thread (max_vec <= 4096)
    if (g.dedupe && a.max_vec >= 256u && nv >= 128u) {
      constexpr uint32_t kEmpty = 0xFFFFFFFFu;
      const uint32_t tbits = 31u - (uint32_t)__clz((int)a.max_vec);
      for (uint32_t e = G.t; e < a.max_vec; e += G.n) s_cmp[e] = kEmpty;
      if (G.t == 0) *s_cnt = 0u;
      G.sync();
      // every stage for a chunk of the thread's candidates at once: the loads of a stage overlap
      uint32_t src[CPT];
      constexpr uint32_t CH = CPT < 8u ? CPT : 8u;
#pragma unroll
      for (uint32_t c0 = 0; c0 < CPT; c0 += CH) {
        uint32_t pos[CH], chk[CH];  // chk: slot whose position has to be compared (key match), or kEmpty
        uint2 ent[CH];
#pragma unroll
        for (uint32_t i = 0; i < CH; i++) {
          const uint32_t ca = G.t + (c0 + i) * G.n;
          pos[i] = ca < nv ? __ldg(cand + ca) : 0u;
        }
#pragma unroll
        for (uint32_t i = 0; i < CH; i++) {
          const uint32_t h = ((pos[i] * 2654435761u) >> (32u - tbits)) & ~1u;  // an aligned pair of slots
          ent[i] = *reinterpret_cast<const uint2*>(s_cmp + h);
        }
#pragma unroll
        for (uint32_t i = 0; i < CH; i++) {
          const uint32_t ca = G.t + (c0 + i) * G.n;
          const uint32_t mix = pos[i] * 2654435761u;
          const uint32_t key = mix & 0xFFFF0000u;
          const uint32_t h = (mix >> (32u - tbits)) & ~1u;
          src[c0 + i] = ca;
          chk[i] = kEmpty;
          if (ca < nv) {
            uint32_t e0 = ent[i].x, e1 = ent[i].y;
            // register as a first occurrence in the first free slot of the pair
            if (e0 == kEmpty) e0 = atomicCAS(&s_cmp[h], kEmpty, key | ca);
            if (e0 != kEmpty) {
              if (((e0 ^ key) >> 16) == 0u) {
                chk[i] = e0 & 0xFFFFu;
              } else {
                if (e1 == kEmpty) e1 = atomicCAS(&s_cmp[h + 1u], kEmpty, key | ca);
                if (e1 != kEmpty && ((e1 ^ key) >> 16) == 0u) chk[i] = e1 & 0xFFFFu;
              }
            }
          }
        }
#pragma unroll
        for (uint32_t i = 0; i < CH; i++) {
          if (chk[i] != kEmpty && __ldg(cand + chk[i]) == pos[i]) src[c0 + i] = chk[i];
        }
      }
      G.sync();  // the table is dead: its words now hold src u16[max_vec] | roots u16[max_vec]
      uint16_t* s_src = reinterpret_cast<uint16_t*>(s_cmp);
      uint16_t* s_root = s_src + a.max_vec;
#pragma unroll
      for (uint32_t i = 0; i < CPT; i++) {
        const uint32_t ca = G.t + i * G.n;
        const bool real = ca < nv;
        const bool root = real && src[i] == ca;
        if (real) s_src[ca] = (uint16_t)src[i];
        const uint32_t mask = __ballot_sync(0xffffffffu, root);
        if (mask) {
          uint32_t off = 0;
          if (lane == 0) off = atomicAdd(s_cnt, __popc(mask));
          off = __shfl_sync(0xffffffffu, off, 0) + __popc(mask & ((1u << lane) - 1u));
          if (root) s_root[off] = (uint16_t)ca;
        }
      }
      G.sync();
      run_scan(*s_cnt, [&](uint32_t e) { return (uint32_t)s_root[e]; });
      G.sync();
#pragma unroll
      for (uint32_t i = 0; i < CPT; i++) {
        const uint32_t ca = G.t + i * G.n;
        if (ca < nv) {
          const uint32_t sc = s_src[ca];
          const float v = s_val[sc];
          if (sc != ca) s_val[ca] = v;
          note(v);
        }
      }
    } else {
      run_scan(nv, [&](uint32_t e) { return e; });
      G.sync();
      for (uint32_t ca = G.t; ca < nv; ca += G.n) note(s_val[ca]);
    }
    umin = __reduce_min_sync(0xffffffffu, umin);
    umax = __reduce_max_sync(0xffffffffu, umax);
    bad = __any_sync(0xffffffffu, bad) ? 1u : 0u;
    if (lane == 0 && nv > 0) {
      atomicMin(s_min, umin);
      atomicMax(s_max, umax);
      if (bad) atomicOr(s_bad, 1u);
    }
    G.sync();
    if (ph && G.t == 0) {
      ph[2] = clock64();
      ph[3] = ph[4] = ph[5] = 0;
    }
    float* od = g.out_dist + (size_t)qi * g.k;
    uint32_t* oi = g.out_idx + (size_t)qi * g.k;
    auto id_of = [&](uint32_t slot) { return __ldg(a.ids + __ldg(cand + slot)); };
    const uint32_t n2 = nv ? pow2ceil(nv) : 0u;
    bool done = false;
    if (g.fast_rank && n2 >= kFastMinN2 && *s_bad == 0u) {
      FastRankState st{*s_min, *s_max};
      uint32_t f = fast_rank_emit<false>(G, 1 + kRerankGroups + grp, s_val, s_cmp, s_fix, s_flag, nv, n2, g.k, st,
                                                od, oi, cand, a.ids, ph);
      if (ph && G.t == 0) ph[3] = clock64();
      // bit-equal distances of different vectors: re-order them the way the network does
      if (f == 1u && g.k >= nv && tie_resolve<false>(G, s_val, s_cmp, s_flag, nv, a.max_vec, od, oi, cand, a.ids,
                                                                  1u + 2u * kRerankGroups + grp, 1u)) {
        f = 0u;
        if (G.t == 0 && g.tie_counter) atomicAdd(g.tie_counter, 1ull);
      }
      done = (f == 0u);
      if (ph && G.t == 0) ph[5] = clock64();
    }
    if (!done)
      rank_and_emit(G, s_val, s_pay, id_of, s_flag, nv, a.max_vec, g.k, od, oi, g.exact_counter);
    if (ph && G.t == 0) {
      ph[6] = clock64();
      ph[7] = ((unsigned long long)blockIdx.x << 32) | (nv << 1) | (done ? 1u : 0u);
    }
    qi = q_next;
    buf ^= 1;
  }
}

// ============================================================================
// Step E2, distance part, as a pure streaming kernel (the split pipeline): no ranking state in
// shared memory, so every warp of the SM streams code rows all the time.  One persistent CTA
// per SM; all warps work on the same query (its LUT is double-buffered with TMA bulk copies)
// and every warp keeps the rows of its NEXT 32 candidates in flight while it evaluates the
// current ones (two register sets).  Writes the distance of list entry i of query q to
// out_val[q][i] (i < n_vec[q]); ids are not touched (the ranking kernel reads them when it
// emits).
// ============================================================================
struct StreamScanArgs {
  const uint32_t* codes;     // [n_local][LP] line codes in bin order (this shard's slice)
  const uint32_t* cand_pos;  // [QN][max_vec] positions to evaluate: the candidate list, or its
                             // first occurrences only (root_pos of the bin walk)
  const uint32_t* n_vec;     // [QN] how many of them
  const float* lut_dup;      // [QN][c1][32]
  const float* cbd;          // [c1*c1][CROW] (replicated rows when CREP)
  uint32_t QN, c1, max_vec;
  float* out_val;            // [QN][max_vec]
  const uint32_t* ids;       // optional: id of every bin-order position; with out_id the kernel also
  uint32_t* out_id;          // stores ids[pos] of every list entry ([QN][max_vec]): the gather rides
                             // in the scan's load pipeline and the ranking kernel reads id rows
  // adc_inbox_kernel (multi-GPU)
  const uint2* inbox;        // [QN][max_vec] (local position, entry number)
  uint32_t q_per_rank;       // queries [r*q_per_rank, (r+1)*q_per_rank) belong to rank r
  float* peer_val[8];        // [world] each [q_per_rank][max_vec], mapped peer (or local) memory
};

inline size_t stream_scan_smem_bytes(uint32_t c1, uint32_t LP, bool crep) {
  return ((size_t)c1 * c1 * (crep ? 32 : LP) + 2 * (size_t)c1 * 32) * 4 + 64;
}

template <int LP, bool CREP, int TPB>
__global__ void __launch_bounds__(TPB, 1) adc_stream_kernel(StreamScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr uint32_t CROW = CREP ? 32u : (uint32_t)LP;
  const uint32_t lut_floats = a.c1 * 32;
  const uint32_t cbd_floats = a.c1 * a.c1 * CROW;
  float* s_cbd = reinterpret_cast<float*>(smem_raw);
  float* s_lut0 = s_cbd + cbd_floats;
  float* s_lut1 = s_lut0 + lut_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lut1 + lut_floats);  // [0],[1]: lut, [2]: cbd
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = TPB >> 5;
  const uint32_t lp = lane & (LP - 1);
  const uint32_t cbd_b = smem_u32(s_cbd) + (CREP ? lane : lp) * 4u;
  const uint32_t* __restrict__ codes_lp = a.codes + lp;
  if (blockIdx.x >= a.QN) return;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t cbd_bytes = cbd_floats * 4;
    mbar_expect_tx(&bars[2], cbd_bytes);
    for (uint32_t off = 0; off < cbd_bytes; off += 32768) {
      uint32_t n = cbd_bytes - off < 32768 ? cbd_bytes - off : 32768;
      tma_bulk_g2s(reinterpret_cast<unsigned char*>(s_cbd) + off,
                   reinterpret_cast<const unsigned char*>(a.cbd) + off, n, &bars[2]);
    }
    mbar_expect_tx(&bars[0], lut_floats * 4);
    tma_bulk_g2s(s_lut0, a.lut_dup + (size_t)blockIdx.x * lut_floats, lut_floats * 4, &bars[0]);
  }
  mbar_wait(&bars[2], 0);

  const uint32_t stride = nwarps * 32;
  uint32_t buf = 0, phase0 = 0, phase1 = 0;
  for (uint32_t qi = blockIdx.x; qi < a.QN; qi += gridDim.x) {
    const uint32_t qn = qi + gridDim.x;
    if (threadIdx.x == 0 && qn < a.QN) {  // the other buffer's readers passed the barrier below
      uint64_t* nb = &bars[buf ^ 1];
      mbar_expect_tx(nb, lut_floats * 4);
      tma_bulk_g2s(buf ? s_lut0 : s_lut1, a.lut_dup + (size_t)qn * lut_floats, lut_floats * 4, nb);
    }
    const uint32_t nv = min(__ldg(a.n_vec + qi), a.max_vec);
    const uint32_t* cand = a.cand_pos + (size_t)qi * a.max_vec;
    float* out = a.out_val + (size_t)qi * a.max_vec;
    // list entry e: (position of the code row, slot of the result)
    auto fetch = [&](uint32_t e, uint32_t& pos, uint32_t& slot) {
      pos = 0u;
      slot = e;
      if (e < nv) pos = __ldg(cand + e);
    };
    // the first two entries of this warp are requested before the LUT wait
    uint32_t e0 = warp * 32 + lane;
    uint32_t pos0, pos1, slot0, slot1;
    fetch(e0, pos0, slot0);
    fetch(e0 + stride, pos1, slot1);
    const float* s_lut = buf ? s_lut1 : s_lut0;
    const uint32_t lut_b = smem_u32(s_lut) + lane * 4u;
    if (warp * 32 < nv) {  // warp-uniform
      uint32_t wA[LP], wB[LP];
      adc_load_rows<LP, false>(wA, pos0, codes_lp, lp, nullptr);
      uint32_t* oid = a.out_id ? a.out_id + (size_t)qi * a.max_vec : nullptr;
      uint32_t idA = 0u, idB = 0u;  // ids of the candidates whose rows are in wA / wB
      if (oid) idA = __ldg(a.ids + pos0);
      mbar_wait(&bars[buf], buf ? phase1 : phase0);
      for (uint32_t base = warp * 32; base < nv; base += 2 * stride) {
        const bool has1 = base + stride < nv, has2 = base + 2 * stride < nv;
        if (has1) {
          adc_load_rows<LP, false>(wB, pos1, codes_lp, lp, nullptr);
          if (oid) idB = __ldg(a.ids + pos1);
        }
        uint32_t pos2, slot2;
        fetch(e0 + 2 * stride, pos2, slot2);
        {
          const float v = adc_eval_rows<LP, CROW>(wA, lut_b, cbd_b, a.c1, lp);
          if (e0 < nv) {
            out[slot0] = v;
            if (oid) oid[slot0] = idA;
          }
        }
        if (!has1) break;
        if (has2) {
          adc_load_rows<LP, false>(wA, pos2, codes_lp, lp, nullptr);
          if (oid) idA = __ldg(a.ids + pos2);
        }
        uint32_t pos3, slot3;
        fetch(e0 + 3 * stride, pos3, slot3);
        {
          const float v = adc_eval_rows<LP, CROW>(wB, lut_b, cbd_b, a.c1, lp);
          if (e0 + stride < nv) {
            out[slot1] = v;
            if (oid) oid[slot1] = idB;
          }
        }
        e0 += 2 * stride;
        slot0 = slot2;
        pos1 = pos3;
        slot1 = slot3;
      }
    } else {
      mbar_wait(&bars[buf], buf ? phase1 : phase0);
    }
    if (buf)
      phase1 ^= 1;
    else
      phase0 ^= 1;
    __syncthreads();  // everyone is done with s_lut[buf] before it is refilled
    buf ^= 1;
  }
}

// ============================================================================
// Inbox scan, one WARP per query (multi-GPU, index sharded by bin range).  The lists are this
// shard's inbox -- for every query of the batch the candidates that live in THIS shard's slice,
// as (local position, candidate slot) pairs written by the query's owner (dispatch_kernel) --
// and the distance of an entry is stored straight into the distance array of the rank that
// owns the query: peer memory over NVLink, 4 bytes per candidate.  The scan and the return
// exchange are one kernel.  An inbox row of an N-shard index holds about max_vec / N entries:
// one or two steps of a 512-thread CTA, so a CTA-per-query loop spends its time in the
// per-query chain (LUT wait -> inbox read -> code rows -> look-ups -> store) instead of
// streaming (measured at 1 B vectors: 0.282 ms vs 0.236 ms per 10 k queries at N = 8, 0.631
// vs 0.604 ms at N = 2).  Here the 16 warps of the persistent CTA walk 16 different
// queries, each with its own 4 KB LUT buffer (own mbarrier, refilled by the warp's lane 0) and
// its own double-buffered code rows; queries are drawn from a global counter.  Shared memory:
// c^2 table + 16 LUTs (128 + 64 KB at c1 = 32, LP = 32).
// ============================================================================
constexpr int kInboxWarps = 16;

inline size_t inbox_scan_smem_bytes(uint32_t c1, uint32_t LP, bool crep) {
  return ((size_t)c1 * c1 * (crep ? 32 : LP) + (size_t)kInboxWarps * c1 * 32) * 4 + (kInboxWarps + 1) * 8 + 64;
}

template <int LP, bool CREP>
__global__ void __launch_bounds__(kInboxWarps * 32, 1) adc_inbox_kernel(StreamScanArgs a, uint32_t* next_query) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr uint32_t CROW = CREP ? 32u : (uint32_t)LP;
  const uint32_t lut_floats = a.c1 * 32;
  const uint32_t cbd_floats = a.c1 * a.c1 * CROW;
  float* s_cbd = reinterpret_cast<float*>(smem_raw);
  float* s_luts = s_cbd + cbd_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_luts + (size_t)kInboxWarps * lut_floats);  // [w]: LUT of warp w, [16]: cbd
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lp = lane & (LP - 1);
  const uint32_t cbd_b = smem_u32(s_cbd) + (CREP ? lane : lp) * 4u;
  const uint32_t* __restrict__ codes_lp = a.codes + lp;
  float* s_lut = s_luts + (size_t)warp * lut_floats;
  const uint32_t lut_b = smem_u32(s_lut) + lane * 4u;
  if (threadIdx.x == 0) {
    for (int i = 0; i <= kInboxWarps; i++) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t cbd_bytes = cbd_floats * 4;
    mbar_expect_tx(&bars[kInboxWarps], cbd_bytes);
    for (uint32_t off = 0; off < cbd_bytes; off += 32768) {
      uint32_t n = cbd_bytes - off < 32768 ? cbd_bytes - off : 32768;
      tma_bulk_g2s(reinterpret_cast<unsigned char*>(s_cbd) + off,
                   reinterpret_cast<const unsigned char*>(a.cbd) + off, n, &bars[kInboxWarps]);
    }
  }
  uint32_t phase = 0;
  bool cbd_ready = false;
  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(next_query, 1u);
    qi = __shfl_sync(0xffffffffu, qi, 0);
    if (qi >= a.QN) break;
    const uint32_t nv = min(__ldg(a.n_vec + qi), a.max_vec);
    if (nv == 0u) continue;  // nothing of this query lives in this shard
    if (lane == 0) {  // the warp's previous reads of its LUT ended with the __syncwarp below
      mbar_expect_tx(&bars[warp], lut_floats * 4);
      tma_bulk_g2s(s_lut, a.lut_dup + (size_t)qi * lut_floats, lut_floats * 4, &bars[warp]);
    }
    const uint2* inbox = a.inbox + (size_t)qi * a.max_vec;
    const uint32_t owner = qi / a.q_per_rank;
    float* out = a.peer_val[owner] + (size_t)(qi - owner * a.q_per_rank) * a.max_vec;
    auto fetch = [&](uint32_t e, uint32_t& pos, uint32_t& slot) {
      pos = 0u;
      slot = 0u;
      if (e < nv) {
        const uint2 en = __ldg(inbox + e);
        pos = en.x;
        slot = en.y;
      }
    };
    uint32_t e0 = lane;
    uint32_t pos0, pos1, slot0, slot1;
    fetch(e0, pos0, slot0);
    fetch(e0 + 32u, pos1, slot1);
    uint32_t wA[LP], wB[LP];
    adc_load_rows<LP, false>(wA, pos0, codes_lp, lp, nullptr);
    if (!cbd_ready) {
      mbar_wait(&bars[kInboxWarps], 0);
      cbd_ready = true;
    }
    mbar_wait(&bars[warp], phase);
    phase ^= 1u;
    for (uint32_t base = 0; base < nv; base += 64u) {
      const bool has1 = base + 32u < nv, has2 = base + 64u < nv;
      if (has1) adc_load_rows<LP, false>(wB, pos1, codes_lp, lp, nullptr);
      uint32_t pos2, slot2;
      fetch(e0 + 64u, pos2, slot2);
      {
        const float v = adc_eval_rows<LP, CROW>(wA, lut_b, cbd_b, a.c1, lp);
        if (e0 < nv) out[slot0] = v;
      }
      if (!has1) break;
      if (has2) adc_load_rows<LP, false>(wA, pos2, codes_lp, lp, nullptr);
      uint32_t pos3, slot3;
      fetch(e0 + 96u, pos3, slot3);
      {
        const float v = adc_eval_rows<LP, CROW>(wB, lut_b, cbd_b, a.c1, lp);
        if (e0 + 32u < nv) out[slot1] = v;
      }
      e0 += 64u;
      slot0 = slot2;
      pos1 = pos3;
      slot1 = slot3;
    }
    __syncwarp();  // every lane is done with the LUT before lane 0 refills it
  }
}

// ============================================================================
// Multi-GPU dispatch: the owner of a query sends every shard the candidates that live in that
// shard's slice of the bin-ordered list -- (position inside the slice, entry number) pairs
// appended to the query's row of the shard's inbox (peer memory over NVLink, 8 bytes per
// candidate), plus the row's length.  One CTA per own query.
// ============================================================================
struct DispatchArgs {
  const uint32_t* list_pos;  // [q_own][max_vec] global positions (candidates or first occurrences)
  const uint32_t* n_list;    // [q_own]
  uint32_t q_own, q_first;   // own queries are q_first .. q_first + q_own - 1 of the batch
  uint32_t max_vec, world;
  uint32_t shard_lo[9];
  uint2* peer_inbox[8];      // [world] each [QN][max_vec]
  uint32_t* peer_cnt[8];     // [world] each [QN]
};

__global__ void __launch_bounds__(256) dispatch_kernel(DispatchArgs a) {
  __shared__ uint32_t s_cnt[8];
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t ql = blockIdx.x; ql < a.q_own; ql += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = min(__ldg(a.n_list + ql), a.max_vec);
    const uint32_t* lp = a.list_pos + (size_t)ql * a.max_vec;
    const size_t row = (size_t)(a.q_first + ql) * a.max_vec;
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
      const uint32_t i = i0 + threadIdx.x;
      uint32_t pos = 0, d = 0xFFu;
      if (i < n) {
        pos = __ldg(lp + i);
        d = 0;
        for (uint32_t r = 1; r < a.world; r++) d += (pos >= a.shard_lo[r]) ? 1u : 0u;
      }
      for (uint32_t r = 0; r < a.world; r++) {
        const uint32_t mask = __ballot_sync(0xffffffffu, d == r);
        if (mask) {
          uint32_t off = 0;
          if (lane == 0) off = atomicAdd(&s_cnt[r], __popc(mask));
          off = __shfl_sync(0xffffffffu, off, 0) + __popc(mask & ((1u << lane) - 1u));
          if (d == r) a.peer_inbox[r][row + off] = make_uint2(pos - a.shard_lo[r], i);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < a.world) a.peer_cnt[threadIdx.x][a.q_first + ql] = s_cnt[threadIdx.x];
  }
}

// ranking only: candidates already in global memory (the split pipeline and the multi-GPU
// paths).  DIRECT: idx[q][slot] is the vector id itself (assembled from shards); otherwise
// idx = cand_pos (bin-order positions) and ids[pos] is the id, read when a result is emitted.
struct Rank2Args {
  const float* val;     // [QN][max_vec]
  const uint32_t* idx;  // [QN][max_vec]: ids (DIRECT) or bin-order positions
  const uint32_t* ids;  // !DIRECT: id of each bin-order position
  uint32_t QN, max_vec, k;
  float* out_dist;
  uint32_t* out_idx;
  unsigned long long* exact_counter;
  unsigned long long* tie_counter;
  const uint32_t* n_vec;  // optional [QN]: number of real candidates; slots beyond are padding
                          // and are NOT read (peer-store mode leaves them unwritten)
  const uint16_t* ridx;   // optional [QN][max_vec]: the distance of candidate slot a is
                          // val[q][ridx[q][a]] (the scan evaluated repeated candidates once)
  uint32_t fast_rank;     // 1: composite-key sort first; 0: the network only
  unsigned long long* phase_dbg;  // optional [QN][8] clock64 stamps (layout of rerank_kernel's;
                                  // slot 1 = end of the sort proper instead of the LUT wait)
  uint32_t* next_query;           // optional work counter (zeroed before the launch): CTAs draw their
                                  // queries from it.  Queries with tie groups take twice as long as
                                  // the others, so a static stride leaves the kernel waiting for its
                                  // unluckiest CTA (measured: slowest CTA 1.28 x the mean)
};

inline size_t rank2_smem_bytes(uint32_t max_vec) { return (size_t)max_vec * 8 + 512 + 64; }
constexpr int kRank2Threads = 256;  // one query per CTA, 4 CTAs per SM

// (5 CTAs per SM at 48 registers were measured: the spills cost more than the occupancy gives,
// 1.28 vs 1.03 ms per 10 k queries at 100 M vectors; an L2 prefetch pipeline in the scan kernel
// likewise: 1.40-1.44 vs 1.34 ms -- profiles/r02k_ab_rank_ctas_scan_prefetch_100m.log)
template <bool DIRECT>
__global__ void __launch_bounds__(kRank2Threads, 4) rank2_kernel(Rank2Args a) {
  extern __shared__ float smem_f[];
  float* s_val = smem_f;
  uint32_t* s_cmp = reinterpret_cast<uint32_t*>(s_val + a.max_vec);  // sort words / network payloads
  uint32_t* s_fix = s_cmp + a.max_vec;                               // [128] repair bitmap
  uint32_t* s_misc = s_fix + 128;                                    // flag, nv, umin, umax, bad
  uint32_t* s_flag = s_misc;
  uint32_t* s_nv = s_misc + 1;
  uint32_t* s_min = s_misc + 2;
  uint32_t* s_max = s_misc + 3;
  uint32_t* s_bad = s_misc + 4;
  uint16_t* s_pay = reinterpret_cast<uint16_t*>(s_cmp);
  const Grp G{threadIdx.x, blockDim.x, 0};
  const uint32_t lane = threadIdx.x & 31u;
  __shared__ uint32_t s_next;
  uint32_t qi = blockIdx.x;
  if (a.next_query) {
    if (threadIdx.x == 0) s_next = atomicAdd(a.next_query, 1u);
    __syncthreads();
    qi = s_next;
  }
  for (; qi < a.QN;) {
    __syncthreads();
    if (a.next_query && threadIdx.x == 0) s_next = atomicAdd(a.next_query, 1u);  // read after the next barrier
    const uint32_t q_after = qi + gridDim.x;
    unsigned long long* ph = a.phase_dbg ? a.phase_dbg + (size_t)qi * 8 : nullptr;
    if (threadIdx.x == 0) {
      *s_flag = 0;
      *s_nv = 0;
      *s_min = 0xFFFFFFFFu;
      *s_max = 0u;
      *s_bad = 0u;
      if (ph) {
        ph[0] = clock64();
        ph[1] = ph[3] = ph[4] = ph[5] = 0;
      }
    }
    if (threadIdx.x < 128u) s_fix[threadIdx.x] = 0u;
    __syncthreads();
    const float* val_row = a.val + (size_t)qi * a.max_vec;
    const uint32_t* idx_row = a.idx + (size_t)qi * a.max_vec;
    // real candidates form a prefix: given (n_vec) or found (slots >= nVec hold (1e7, PAD))
    const uint32_t limit = a.n_vec ? min(a.n_vec[qi], a.max_vec) : a.max_vec;
    uint32_t local = 0, umin = 0xFFFFFFFFu, umax = 0u, bad = 0u;
    const uint16_t* ridx_row = a.ridx ? a.ridx + (size_t)qi * a.max_vec : nullptr;
    // 8 slots of a thread at a time: their loads are all issued before the first one is used
    // (the row comes from L2 / HBM; a rolled loop would pay that latency once per slot), and
    // the id of every candidate is requested into L2 for the emit
    constexpr int kSlots = 8;
    for (uint32_t eb = threadIdx.x; eb < limit; eb += kSlots * kRank2Threads) {
      uint32_t src[kSlots], ix[kSlots];
      float vv[kSlots];
#pragma unroll
      for (int r = 0; r < kSlots; r++) {
        const uint32_t e = eb + r * kRank2Threads;
        src[r] = e;
        if (ridx_row && e < limit) src[r] = ridx_row[e];
      }
#pragma unroll
      for (int r = 0; r < kSlots; r++) {
        const uint32_t e = eb + r * kRank2Threads;
        vv[r] = 0.f;
        ix[r] = 0u;
        if (e < limit) {
          vv[r] = val_row[src[r]];
          if (!DIRECT || !a.n_vec) ix[r] = idx_row[e];
        }
      }
#pragma unroll
      for (int r = 0; r < kSlots; r++) {
        const uint32_t e = eb + r * kRank2Threads;
        if (e < limit) {
          const float v = vv[r];
          s_val[e] = v;
          const bool real = a.n_vec ? true : !(ix[r] == kPadIdx && v == kPadDist);
          if (real) {
            local = e + 1;
            const uint32_t u = sortable_key(v);
            umin = min(umin, u);
            umax = max(umax, u);
            if (!(v < kPadDist) || !(v > -__int_as_float(0x7f800000))) bad = 1u;
            if (!DIRECT) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.ids + ix[r]));
          }
        }
      }
    }
    local = __reduce_max_sync(0xffffffffu, local);
    umin = __reduce_min_sync(0xffffffffu, umin);
    umax = __reduce_max_sync(0xffffffffu, umax);
    bad = __any_sync(0xffffffffu, bad) ? 1u : 0u;
    if (lane == 0) {
      atomicMax(s_nv, local);
      atomicMin(s_min, umin);
      atomicMax(s_max, umax);
      if (bad) atomicOr(s_bad, 1u);
    }
    __syncthreads();
    const uint32_t nv = a.n_vec ? limit : *s_nv;
    if (ph && threadIdx.x == 0) ph[2] = clock64();
    float* od = a.out_dist + (size_t)qi * a.k;
    uint32_t* oi = a.out_idx + (size_t)qi * a.k;
    const uint32_t* cand = DIRECT ? nullptr : idx_row;
    const uint32_t* ids = DIRECT ? idx_row : a.ids;
    auto id_of = [&](uint32_t slot) { return DIRECT ? __ldg(idx_row + slot) : __ldg(a.ids + __ldg(idx_row + slot)); };
    const uint32_t n2 = nv ? pow2ceil(nv) : 0u;
    bool done = false;
    if (a.fast_rank && n2 >= kFastMinN2 && *s_bad == 0u) {
      FastRankState st{*s_min, *s_max};
      uint32_t f = fast_rank_emit<DIRECT>(G, 1, s_val, s_cmp, s_fix, s_flag, nv, n2, a.k, st, od, oi, cand, ids,
                                          ph);
      if (ph && threadIdx.x == 0) ph[3] = clock64();
      if (ph && threadIdx.x == 0) s_misc[8] = s_misc[9] = 0u;
      if (f == 1u && a.k >= nv &&
          tie_resolve<DIRECT>(G, s_val, s_cmp, s_flag, nv, a.max_vec, od, oi, cand, ids, 2u, 2u,
                              ph ? s_misc + 8 : nullptr)) {
        f = 0u;
        if (threadIdx.x == 0 && a.tie_counter) atomicAdd(a.tie_counter, 1ull);
      }
      done = (f == 0u);
      if (ph && threadIdx.x == 0) ph[5] = clock64();
    }
    if (!done)
      rank_and_emit(G, s_val, s_pay, id_of, s_flag, nv, a.max_vec, a.k, od, oi, a.exact_counter);
    if (ph && threadIdx.x == 0) {
      ph[6] = clock64();
      // meta: block | tie groups << 48 | (cycles of the tie search >> 4) << 14 | nv << 1 | fast
      ph[7] = ((unsigned long long)blockIdx.x << 32) | ((unsigned long long)(s_misc[9] & 15u) << 48) |
              ((unsigned long long)min(s_misc[8] >> 4, 0x3FFFFu) << 14) | (nv << 1) | (done ? 1u : 0u) | (1ull << 63);
    }
    qi = a.next_query ? s_next : q_after;  // s_next: written before at least one barrier of this iteration
  }
}

}  // namespace pqtb
