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
// caller has to rank the query with the network itself.  g.n is a multiple of 128.
// DIRECT: ids[slot] is the id (see fast_rank.cuh).
template <bool DIRECT>
__device__ __forceinline__ uint32_t tie_resolve(const Grp& g, const float* s_val, uint32_t* s_scr,
                                                uint32_t* s_flag, uint32_t nv, uint32_t max_vec,
                                                const float* out_dist, uint32_t* out_idx,
                                                const uint32_t* __restrict__ cand,
                                                const uint32_t* __restrict__ ids) {
  const uint32_t t = g.t;
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
  if (*s_flag & 8u) return 0u;
  if (cnt == 0u) return 1u;  // only duplicates of one vector shared a distance: nothing to do

  // ---- B. one WARP per group: lane l owns words 4l .. 4l+3 of the three planes (max_vec <=
  // 4096 positions = 128 words), so the whole network runs in registers and shuffles -- no
  // shared memory, no barrier, and the warps without a group stay out of the issue slots.
  const uint32_t lane = t & 31u, warp = t >> 5, nwarps = g.n >> 5;
  for (uint32_t g0 = 0; g0 < cnt; g0 += nwarps) {
    const uint32_t gi = g0 + warp;
    if (gi >= cnt) break;  // warp-uniform
    const uint32_t r = s_gr[gi], m = s_gm[gi];
    const uint32_t idA = s_ga[gi], idB = s_gb[gi];
    const float v = out_dist[r];
    uint32_t L[4] = {0u, 0u, 0u, 0u}, H[4] = {~0u, ~0u, ~0u, ~0u}, B[4] = {0u, 0u, 0u, 0u};
    // classes of the input positions (candidate order, pads = Higher); word w goes to lane w / 4
    for (uint32_t w0 = 0; w0 < nW; w0 += 4u) {
#pragma unroll
      for (uint32_t i = 0; i < 4u; i++) {
        const uint32_t w = w0 + i;
        if (w < nW) {  // warp-uniform
          const uint32_t a = (w << 5) + lane;
          const float val = a < nv ? s_val[a] : kPadDist;
          const bool isL = val < v, isH = val > v;
          bool isB = false;
          if (!isL && !isH) isB = ident_id<DIRECT>(ids, slot_ident<DIRECT>(cand, ids, a)) != idA;
          const uint32_t bl = __ballot_sync(0xffffffffu, isL);
          const uint32_t bh = __ballot_sync(0xffffffffu, isH);
          const uint32_t bb = __ballot_sync(0xffffffffu, isB);
          if (lane == (w0 >> 2)) {
            L[i] = bl;
            H[i] = bh;
            B[i] = bb;
          }
        }
      }
    }
    // ---- fast-forward.  After stage kk of the network every aligned block of kk positions is
    // sorted (ascending iff (pos & kk) == 0): Lower, Equal, Higher in a row.  While no block
    // holds two Equal positions that arrangement follows from the block's class counts alone --
    // the one Equal element, if any, keeps its label -- so the planes after stage kk0 = the
    // largest such block size are written down directly and only the stages that bring two
    // Equal elements into one block are simulated (for a two-element group at random positions
    // 12 of the 78 sub-stages half of the time, 23 a quarter of the time, ...).
    uint32_t kk_first = 2u;
    {
      uint32_t cLH[4], cEB[4];  // popcounts: Lower | Higher << 16, Equal | label-B << 16
#pragma unroll
      for (uint32_t i = 0; i < 4u; i++) {
        const uint32_t E = ~(L[i] | H[i]);
        cLH[i] = __popc(L[i]) | (__popc(H[i]) << 16);
        cEB[i] = __popc(E) | (__popc(B[i] & E) << 16);
      }
      // block totals for growing block sizes bw (in words); stop at the first size with two
      // Equal positions in some block
      uint32_t bw0 = 0u;
      uint32_t tLH[4], tEB[4];
      if (!__any_sync(0xffffffffu, ((cEB[0] | cEB[1] | cEB[2] | cEB[3]) & 0xFFFEu) != 0u)) {
        bw0 = 1u;
#pragma unroll
        for (uint32_t i = 0; i < 4u; i++) {
          tLH[i] = cLH[i];
          tEB[i] = cEB[i];
        }
        const uint32_t p0 = cEB[0] + cEB[1], p1 = cEB[2] + cEB[3];
        if (nW >= 2u && !__any_sync(0xffffffffu, ((p0 | p1) & 0xFFFEu) != 0u)) {
          bw0 = 2u;
          tLH[0] = tLH[1] = cLH[0] + cLH[1];
          tLH[2] = tLH[3] = cLH[2] + cLH[3];
          tEB[0] = tEB[1] = p0;
          tEB[2] = tEB[3] = p1;
          uint32_t sLH = tLH[0] + tLH[2], sEB = p0 + p1;
          if (nW >= 4u && !__any_sync(0xffffffffu, (sEB & 0xFFFEu) != 0u)) {
            bw0 = 4u;
            for (uint32_t ld = 1u; ld < 32u; ld <<= 1) {
              const uint32_t oLH = __shfl_xor_sync(0xffffffffu, sLH, ld);
              const uint32_t oEB = __shfl_xor_sync(0xffffffffu, sEB, ld);
              if ((ld << 3) > nW || __any_sync(0xffffffffu, ((sEB + oEB) & 0xFFFEu) != 0u)) break;
              sLH += oLH;
              sEB += oEB;
              bw0 = ld << 3;
            }
#pragma unroll
            for (uint32_t i = 0; i < 4u; i++) {
              tLH[i] = sLH;
              tEB[i] = sEB;
            }
          }
        }
      }
      if (bw0 != 0u) {
        auto low = [](int n) { return n <= 0 ? 0u : n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); };
#pragma unroll
        for (uint32_t i = 0; i < 4u; i++) {
          const uint32_t x = (lane << 2) + i;
          const int o = (int)((x & (bw0 - 1u)) << 5);  // first position of the word inside its block
          const int nL = (int)(tLH[i] & 0xFFFFu), nH = (int)(tLH[i] >> 16);
          const int nE = (int)(tEB[i] & 0xFFFFu);
          const bool lab = (tEB[i] >> 16) != 0u;
          if ((x & bw0) == 0u) {  // ascending block: Lower, Equal, Higher
            L[i] = low(nL - o);
            H[i] = ~low(nL + nE - o);
          } else {                // descending block: Higher, Equal, Lower
            H[i] = low(nH - o);
            L[i] = ~low(nH + nE - o);
          }
          B[i] = lab ? ~(L[i] | H[i]) : 0u;
        }
        kk_first = bw0 << 6;  // the first stage that still has to run: 2 * 32 * bw0
      }
    }
    // the network of pqt/bitonicSort.cuh:16-44 / :47-78: for k = 2..n, for j = k/2..1:
    // pairs (i, i^j), ascending iff (i & k) == 0.  x = 4*lane + i is the word index.
    for (uint32_t kk = kk_first; kk <= max_vec; kk <<= 1) {
      for (uint32_t j = kk >> 1; j > 0u; j >>= 1) {
        if (j < 32u) {
          const uint32_t M = j == 1u ? 0x55555555u : j == 2u ? 0x33333333u : j == 4u ? 0x0F0F0F0Fu
                             : j == 8u ? 0x00FF00FFu : 0x0000FFFFu;
          if (kk < 32u) {
            const uint32_t Mk = kk == 2u ? 0x33333333u : kk == 4u ? 0x0F0F0F0Fu
                                : kk == 8u ? 0x00FF00FFu : 0x0000FFFFu;
            const uint32_t D = ~Mk & M;  // positions with (pos & kk) != 0
#pragma unroll
            for (uint32_t i = 0; i < 4u; i++) tie_substage_inword(L[i], H[i], B[i], j, M, D);
          } else {
            const uint32_t kw = kk >> 5;
#pragma unroll
            for (uint32_t i = 0; i < 4u; i++)
              tie_substage_inword(L[i], H[i], B[i], j, M, (((lane << 2) + i) & kw) ? M : 0u);
          }
        } else {
          const uint32_t wd = j >> 5, kw = kk >> 5;
          if (wd < 4u) {
            // partner word in the same lane: i ^ wd
#pragma unroll
            for (uint32_t i = 0; i < 4u; i++) {
              if ((i & wd) == 0u) {
                const uint32_t q = i ^ wd;  // wd is 1 or 2: compile-time after unrolling on both
                const bool desc = (((lane << 2) + i) & kw) != 0u;
                uint32_t L0 = L[i], H0 = H[i], B0 = B[i];
                uint32_t L1 = wd == 1u ? L[i ^ 1u] : L[i ^ 2u];
                uint32_t H1 = wd == 1u ? H[i ^ 1u] : H[i ^ 2u];
                uint32_t B1 = wd == 1u ? B[i ^ 1u] : B[i ^ 2u];
                const uint32_t gt = (H0 & ~H1) | (~L0 & L1);
                const uint32_t lt = (H1 & ~H0) | (~L1 & L0);
                const uint32_t sw = desc ? lt : gt;
                const uint32_t dL = (L0 ^ L1) & sw, dH = (H0 ^ H1) & sw, dB = (B0 ^ B1) & sw;
                L[i] = L0 ^ dL;
                H[i] = H0 ^ dH;
                B[i] = B0 ^ dB;
                if (wd == 1u) {
                  L[i ^ 1u] = L1 ^ dL;
                  H[i ^ 1u] = H1 ^ dH;
                  B[i ^ 1u] = B1 ^ dB;
                } else {
                  L[i ^ 2u] = L1 ^ dL;
                  H[i ^ 2u] = H1 ^ dH;
                  B[i ^ 2u] = B1 ^ dB;
                }
                (void)q;
              }
            }
          } else {
            const uint32_t ld = wd >> 2;  // lane distance
            const bool lo = (lane & ld) == 0u;
#pragma unroll
            for (uint32_t i = 0; i < 4u; i++) {
              const uint32_t oL = __shfl_xor_sync(0xffffffffu, L[i], ld);
              const uint32_t oH = __shfl_xor_sync(0xffffffffu, H[i], ld);
              const uint32_t oB = __shfl_xor_sync(0xffffffffu, B[i], ld);
              tie_substage_xword(L[i], H[i], B[i], oL, oH, oB, lo, (((lane << 2) + i) & kw) != 0u);
            }
          }
        }
      }
    }
    // ---- the Equal positions are now the group's output slots r .. r+m-1
#pragma unroll
    for (uint32_t i = 0; i < 4u; i++) {
      const uint32_t x = (lane << 2) + i;
      if (x < nW) {
        const uint32_t base = x << 5;
        uint32_t want = 0;
        if (base < r + m && base + 32u > r) {
          const uint32_t lo = r > base ? r - base : 0u;
          const uint32_t hi = min(32u, r + m - base);
          want = (hi >= 32u ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
        }
        const uint32_t eq = ~(L[i] | H[i]);
        if (eq != want) atomicOr(s_flag, 8u);  // cannot happen for a sorted result; be safe
        uint32_t todo = want;
        while (todo) {
          const uint32_t b = __ffs(todo) - 1u;
          todo &= todo - 1u;
          out_idx[base + b] = ((B[i] >> b) & 1u) ? idB : idA;
        }
      }
    }
  }
  g.sync();
  return (*s_flag & 8u) ? 0u : 1u;
}

}  // namespace pqtb
