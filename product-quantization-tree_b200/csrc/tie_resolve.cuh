// tie_resolve.cuh -- the reference network's order inside groups of bit-equal distances,
// without running the network on the data.
//
// The reference ranks with a bitonic network whose compare-exchanges are strict
// (pqt/bitonicSort.cuh:16-44: swap iff val[i] > val[ixj], resp. <), so candidates with
// bit-equal distances are never swapped with each other and their final order depends on
// how the rest of the input pushes them around.  For one group of equal distances (value v)
// every comparator outcome that matters is decided by three classes: Lower (< v), Equal,
// Higher (> v, pads included) -- a monotone image of the input, and the network commutes with
// monotone maps.  The classes of all max_vec input positions are held as bit planes (one
// 32-bit word per 32 positions, four words per lane of ONE warp) and the whole network is run
// on the planes with bitwise operations in registers and shuffles: 32 comparators per
// instruction instead of one, no shared memory, no barrier.  A third
// plane carries which of the (two) different vectors an Equal position holds; it is moved by
// the same swaps.  When the network is done the Equal positions are the group's output slots
// and the label plane says which vector goes where.
//
// Handles groups with exactly two different vectors (any number of duplicates of each), up to
// kTieMaxGroups groups per query; anything else is left to the caller (the full network).
#pragma once
#include "common.cuh"
#include "fast_rank.cuh"

namespace pqtb {

constexpr uint32_t kTieMaxGroups = 8;
constexpr uint32_t kTieMaxLen = 64;

// One network substage at bit distance j < 32 inside the 32-position words.
//   M: positions with (pos & j) == 0; D: those of them that sit in a descending block
__device__ __forceinline__ void tie_substage_inword(uint32_t& L, uint32_t& H, uint32_t& B, uint32_t j,
                                                    uint32_t M, uint32_t D) {
  const uint32_t aL = L & M, bL = (L >> j) & M;
  const uint32_t aH = H & M, bH = (H >> j) & M;
  const uint32_t aB = B & M, bB = (B >> j) & M;
  const uint32_t gt = (aH & ~bH) | (~aL & bL);  // val[i] > val[ixj]   (classes L < E < H)
  const uint32_t lt = (bH & ~aH) | (~bL & aL);  // val[i] < val[ixj]
  const uint32_t sw = (gt & ~D) | (lt & D);
  const uint32_t dL = (aL ^ bL) & sw, dH = (aH ^ bH) & sw, dB = (aB ^ bB) & sw;
  L ^= dL | (dL << j);
  H ^= dH | (dH << j);
  B ^= dB | (dB << j);
}

// One network substage between this thread's word and the word of its partner (oL, oH, oB).
__device__ __forceinline__ void tie_substage_xword(uint32_t& L, uint32_t& H, uint32_t& B, uint32_t oL,
                                                   uint32_t oH, uint32_t oB, bool lo, bool desc) {
  const uint32_t aL = lo ? L : oL, bL = lo ? oL : L;
  const uint32_t aH = lo ? H : oH, bH = lo ? oH : H;
  const uint32_t gt = (aH & ~bH) | (~aL & bL);
  const uint32_t lt = (bH & ~aH) | (~bL & aL);
  const uint32_t sw = desc ? lt : gt;
  L ^= (L ^ oL) & sw;
  H ^= (H ^ oH) & sw;
  B ^= (B ^ oB) & sw;
}

// Re-orders the ids inside every group of bit-equal distances of one query's result the way
// the reference's network (width max_vec, input = candidate order + 1e7 pads) leaves them.
//   s_val[a]  : distance of candidate slot a (a < nv), candidate order
//   s_scr     : scratch, max_vec words
//   out_dist / out_idx: the query's first k results, sorted by distance (k >= nv: every real
//               candidate is in the output); ids inside equal-distance groups in any order
//   cand / ids: candidate slot -> bin-order position -> vector id
// Every thread of the group must call.  Returns 1 when all groups were resolved, 0 when the
// caller has to rank the query with the network itself.  g.n is a multiple of 128; the first
// nteams * 128 threads of the group (nteams * 128 <= g.n) simulate, one group of ties per team,
// and use the hardware barriers team_bar0 .. team_bar0 + nteams - 1.
// DIRECT: ids[slot] is the id (see fast_rank.cuh).
template <bool DIRECT>
__device__ __forceinline__ uint32_t tie_resolve(const Grp& g, const float* s_val, uint32_t* s_scr,
                                                uint32_t* s_flag, uint32_t nv, uint32_t max_vec,
                                                const float* out_dist, uint32_t* out_idx,
                                                const uint32_t* __restrict__ cand,
                                                const uint32_t* __restrict__ ids,
                                                uint32_t team_bar0, uint32_t nteams,
                                                uint32_t* dbg = nullptr) {
  // dbg (optional, 2 shared words): cycles spent finding the groups, number of groups
  const uint32_t t = g.t;
  const long long tdbg0 = dbg ? clock64() : 0ll;
  uint32_t* s_cnt = s_scr;           // [1]
  uint32_t* s_gr = s_scr + 8;        // [kTieMaxGroups] first slot
  uint32_t* s_gm = s_gr + 8;         // length
  uint32_t* s_ga = s_gm + 8;         // first vector's id
  uint32_t* s_gb = s_ga + 8;         // the other vector's id
  const uint32_t nW = max_vec >> 5;  // words per plane
  if (nW == 0u || nW > 128u || max_vec < 64u) return 0u;  // planes do not fit a warp / scratch (uniform)
  if (t == 0) *s_cnt = 0;
  g.sync();
  // ---- A. find the groups: slot e starts a tie when it repeats the distance of e-1 with a
  // different id; the first such slot of a group registers the group.  A thread looks at
  // kScan consecutive slots, all loads issued together.
  bool fail = false;
  constexpr uint32_t kScan = 8;
  for (uint32_t e0 = t * kScan; e0 < nv; e0 += g.n * kScan) {
    float d[kScan + 1];
    uint32_t id[kScan + 1];
#pragma unroll
    for (uint32_t i = 0; i <= kScan; i++) {
      const uint32_t e = e0 + i - 1u;  // e0 == 0: wraps, fails the range test
      const bool ok = e < nv;
      d[i] = ok ? out_dist[e] : 0.f;
      id[i] = ok ? out_idx[e] : 0u;
    }
    uint32_t hits = 0;
#pragma unroll
    for (uint32_t i = 1; i <= kScan; i++) {
      const uint32_t e = e0 + i - 1u;
      if (e > 0u && e < nv && d[i - 1] == d[i] && id[i - 1] != id[i]) hits |= 1u << (i - 1u);
    }
    while (hits) {
      const uint32_t e = e0 + (uint32_t)__ffs(hits) - 1u;
      hits &= hits - 1u;
      const float d1 = out_dist[e];
      uint32_t r = e - 1u;
      bool first = true;
      while (r > 0u && out_dist[r - 1] == d1) {
        if (out_idx[r - 1] != out_idx[r]) {
          first = false;
          break;
        }
        if (e - r > kTieMaxLen) break;
        r--;
      }
      if (!first) continue;
      if (e - r > kTieMaxLen) {
        fail = true;
        continue;
      }
      uint32_t m = e - r + 1u;
      const uint32_t idA = out_idx[r], idB = out_idx[e];
      while (r + m < nv && out_dist[r + m] == d1 && m <= kTieMaxLen) {
        const uint32_t x = out_idx[r + m];
        if (x != idA && x != idB) fail = true;  // a third vector
        m++;
      }
      if (m > kTieMaxLen) fail = true;
      const uint32_t slot = atomicAdd(s_cnt, 1u);
      if (slot < kTieMaxGroups) {
        s_gr[slot] = r;
        s_gm[slot] = m;
        s_ga[slot] = idA;
        s_gb[slot] = idB;
      } else {
        fail = true;
      }
    }
  }
  if (fail) atomicOr(s_flag, 8u);
  g.sync();
  const uint32_t cnt = *s_cnt;
  if (dbg && t == 0) {
    dbg[0] = (uint32_t)(clock64() - tdbg0);
    dbg[1] = cnt;
  }
  if (*s_flag & 8u) return 0u;
  if (cnt == 0u) return 1u;  // only duplicates of one vector shared a distance: nothing to do

  // ---- B. one TEAM of 128 threads per group: thread x of the team owns word x of the three
  // planes (max_vec <= 4096 positions = 128 words).  Sub-stages at bit distance < 32 are
  // bitwise inside the word, distances 32 .. 512 are shuffles inside the team's warps, the
  // three sub-stages at distance 1024 / 2048 go through shared memory.  Teams synchronise on
  // their own named barrier (team_bar0 + team); threads beyond the teams wait at the end.
  const uint32_t lane = t & 31u;
  const uint32_t team = t >> 7, x = t & 127u, tw = x >> 5;
  if (team < nteams) {
    uint32_t* s_tm = s_scr + 64u + team * 384u;  // planes in transit (only lists > 1024 get there)
    uint32_t* s_lvl = s_scr + 48u + team;        // level mask of the fast-forward
    const uint32_t tbar = team_bar0 + team;
    auto team_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(tbar), "r"(128) : "memory"); };
    for (uint32_t gi = team; gi < cnt; gi += nteams) {  // team-uniform
      const uint32_t r = s_gr[gi], m = s_gm[gi];
      const uint32_t idA = s_ga[gi], idB = s_gb[gi];
      const float v = out_dist[r];
      if (x == 0u) *s_lvl = 0u;
      // classes of the input positions (candidate order, pads = Higher): warp tw of the team
      // builds words 32*tw .. 32*tw+31 with ballots, lane i keeps word 32*tw + i
      uint32_t L = 0u, H = ~0u, B = 0u;
      for (uint32_t i = 0; i < 32u; i++) {
        const uint32_t w = (tw << 5) + i;
        if (w < nW) {  // warp-uniform
          const uint32_t a = (w << 5) + lane;
          const float val = a < nv ? s_val[a] : kPadDist;
          const bool isL = val < v, isH = val > v;
          bool isB = false;
          if (!isL && !isH) isB = ident_id<DIRECT>(ids, slot_ident<DIRECT>(cand, ids, a)) != idA;
          const uint32_t bl = __ballot_sync(0xffffffffu, isL);
          const uint32_t bh = __ballot_sync(0xffffffffu, isH);
          const uint32_t bb = __ballot_sync(0xffffffffu, isB);
          if (lane == i) {
            L = bl;
            H = bh;
            B = bb;
          }
        }
      }
      // ---- fast-forward.  After stage kk of the network every aligned block of kk positions
      // is sorted (ascending iff (pos & kk) == 0): Lower, Equal, Higher in a row.  While no
      // block holds two Equal positions that arrangement follows from the block's class counts
      // alone -- the one Equal element, if any, keeps its label -- so the planes after stage
      // kk0 = the largest such block size (up to the 1024 positions of one warp) are written
      // down directly and only the later stages are simulated.
      const uint32_t cLH = __popc(L) | (__popc(H) << 16);          // Lower | Higher << 16
      const uint32_t cEB = __popc(~(L | H)) | (__popc(B) << 16);   // Equal | label-B << 16 (B is a subset of Equal)
      {
        uint32_t viol = __any_sync(0xffffffffu, (cEB & 0xFFFEu) != 0u) ? 1u : 0u;  // bit l: blocks of 2^l words
        uint32_t sEB = cEB;
        for (uint32_t ld = 1u, bit = 2u; ld < 32u; ld <<= 1, bit <<= 1) {
          sEB += __shfl_xor_sync(0xffffffffu, sEB, ld);
          if (__any_sync(0xffffffffu, (sEB & 0xFFFEu) != 0u)) viol |= bit;
        }
        team_sync();  // the level mask was cleared
        if (lane == 0u && viol) atomicOr(s_lvl, viol);
        team_sync();
      }
      uint32_t kk_first = 2u;
      {
        const uint32_t viol = *s_lvl;
        uint32_t bw0 = 0u;  // words per block of the last stage that needs no simulation
        for (uint32_t bw = 1u, bit = 1u; bw <= 32u && bw <= nW && !(viol & bit); bw <<= 1, bit <<= 1) bw0 = bw;
        if (bw0 != 0u) {
          uint32_t sLH = cLH, sEB = cEB;
          for (uint32_t ld = 1u; ld < bw0; ld <<= 1) {
            sLH += __shfl_xor_sync(0xffffffffu, sLH, ld);
            sEB += __shfl_xor_sync(0xffffffffu, sEB, ld);
          }
          auto low = [](int n) { return n <= 0 ? 0u : n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); };
          const int o = (int)((x & (bw0 - 1u)) << 5);  // first position of the word inside its block
          const int nL = (int)(sLH & 0xFFFFu), nH = (int)(sLH >> 16);
          const int nE = (int)(sEB & 0xFFFFu);
          if ((x & bw0) == 0u) {  // ascending block: Lower, Equal, Higher
            L = low(nL - o);
            H = ~low(nL + nE - o);
          } else {                // descending block: Higher, Equal, Lower
            H = low(nH - o);
            L = ~low(nH + nE - o);
          }
          B = (sEB >> 16) != 0u ? ~(L | H) : 0u;
          kk_first = bw0 << 6;  // the first stage that still has to run: 2 * 32 * bw0
        }
      }
      // the network of pqt/bitonicSort.cuh:16-44 / :47-78: for k = 2..n, for j = k/2..1:
      // pairs (i, i^j), ascending iff (i & k) == 0
      for (uint32_t kk = kk_first; kk <= max_vec; kk <<= 1) {
        const uint32_t kw = kk >> 5;
        for (uint32_t j = kk >> 1; j > 0u; j >>= 1) {
          if (j < 32u) {
            const uint32_t M = j == 1u ? 0x55555555u : j == 2u ? 0x33333333u : j == 4u ? 0x0F0F0F0Fu
                               : j == 8u ? 0x00FF00FFu : 0x0000FFFFu;
            uint32_t D;
            if (kk < 32u) {
              const uint32_t Mk = kk == 2u ? 0x33333333u : kk == 4u ? 0x0F0F0F0Fu
                                  : kk == 8u ? 0x00FF00FFu : 0x0000FFFFu;
              D = ~Mk & M;  // positions with (pos & kk) != 0
            } else {
              D = (x & kw) ? M : 0u;
            }
            tie_substage_inword(L, H, B, j, M, D);
          } else {
            const uint32_t wd = j >> 5;
            uint32_t oL, oH, oB;
            if (wd < 32u) {
              oL = __shfl_xor_sync(0xffffffffu, L, wd);
              oH = __shfl_xor_sync(0xffffffffu, H, wd);
              oB = __shfl_xor_sync(0xffffffffu, B, wd);
            } else {
              s_tm[x] = L;
              s_tm[128u + x] = H;
              s_tm[256u + x] = B;
              team_sync();
              oL = s_tm[x ^ wd];
              oH = s_tm[128u + (x ^ wd)];
              oB = s_tm[256u + (x ^ wd)];
              team_sync();
            }
            tie_substage_xword(L, H, B, oL, oH, oB, (x & wd) == 0u, (x & kw) != 0u);
          }
        }
      }
      // ---- the Equal positions are now the group's output slots r .. r+m-1
      if (x < nW) {
        const uint32_t base = x << 5;
        uint32_t want = 0;
        if (base < r + m && base + 32u > r) {
          const uint32_t lo = r > base ? r - base : 0u;
          const uint32_t hi = min(32u, r + m - base);
          want = (hi >= 32u ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
        }
        const uint32_t eq = ~(L | H);
        if (eq != want) atomicOr(s_flag, 8u);  // cannot happen for a sorted result; be safe
        uint32_t todo = want;
        while (todo) {
          const uint32_t b = __ffs(todo) - 1u;
          todo &= todo - 1u;
          out_idx[base + b] = ((B >> b) & 1u) ? idB : idA;
        }
      }
      team_sync();  // everyone has read the level mask before the next group clears it
    }
  }
  g.sync();
  return (*s_flag & 8u) ? 0u : 1u;
}

}  // namespace pqtb
