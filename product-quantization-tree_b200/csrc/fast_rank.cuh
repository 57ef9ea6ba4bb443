// fast_rank.cuh -- ranking of one query's candidates without the reference's network.
//
// The reference ranks with a key/value bitonic network and strict float compares
// (pqt/bitonicSort.cuh:16-78).  Its result is "ascending by distance"; only the order
// inside a group of bit-equal distances depends on the network.  The fast path sorts
// 32-bit composite words  (order-preserving key in the high bits | candidate slot)  with
// unsigned min/max compare-exchanges (2 ALU instructions instead of 5, one shuffle
// instead of two, 4 bytes per element), then repairs the few neighbours whose truncated
// keys collide by comparing their full 32-bit keys while the results are emitted.
// Equal distances of the SAME vector (a bin listed twice) are interchangeable; equal
// distances of DIFFERENT vectors are reported to the caller, which resolves the group
// (tie_resolve.cuh) or runs the reference's network -- so the output is exactly the
// network's in every case.
#pragma once
#include "common.cuh"

namespace pqtb {

// order-preserving map float -> uint32 (-0 and +0 compare equal under the network's
// float compares, so they get the same key)
__device__ __forceinline__ uint32_t sortable_key(float v) {
  uint32_t b = __float_as_uint(v);
  b = (b == 0x80000000u) ? 0u : b;
  return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}

__device__ __forceinline__ void ce_u32(uint32_t& a, uint32_t& b) {
  const uint32_t lo = min(a, b), hi = max(a, b);
  a = lo;
  b = hi;
}

// ascending sort of the E registers of one thread (bitonic, flip formulation: every
// compare-exchange keeps the minimum at the lower index)
template <int E>
__device__ __forceinline__ void reg_sort_u32(uint32_t (&c)[E]) {
#pragma unroll
  for (int k = 2; k <= E; k <<= 1) {
#pragma unroll
    for (int r = 0; r < E; r++) {
      const int q = r ^ (k - 1);
      if (q > r) ce_u32(c[r], c[q]);
    }
#pragma unroll
    for (int j = k >> 2; j > 0; j >>= 1) {
#pragma unroll
      for (int r = 0; r < E; r++) {
        const int q = r ^ j;
        if (q > r) ce_u32(c[r], c[q]);
      }
    }
  }
}

// half-cleaners at distances E/2 .. 1 inside a thread
template <int E>
__device__ __forceinline__ void reg_clean_u32(uint32_t (&c)[E]) {
#pragma unroll
  for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
    for (int r = 0; r < E; r++) {
      const int q = r ^ j;
      if (q > r) ce_u32(c[r], c[q]);
    }
  }
}

// One cross-thread step: thread t exchanges with thread t ^ m (register r with register r,
// or with register E-1-r when REV) and keeps the minima when it holds the lower indices.
// m < 32: shuffles; otherwise through s_x ([E][T] words) with the sub-group barrier.
template <int E, bool REV>
__device__ __forceinline__ void xthread_step_u32(uint32_t (&c)[E], uint32_t t, uint32_t T,
                                                 uint32_t m, bool keep_min, uint32_t* s_x,
                                                 uint32_t bar_id) {
  uint32_t o[E];
  if (m < 32u) {
#pragma unroll
    for (int r = 0; r < E; r++) o[r] = __shfl_xor_sync(0xffffffffu, c[REV ? E - 1 - r : r], m);
  } else {
#pragma unroll
    for (int r = 0; r < E; r++) s_x[r * T + t] = c[r];
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(T) : "memory");
    const uint32_t tp = t ^ m;
#pragma unroll
    for (int r = 0; r < E; r++) o[r] = s_x[(REV ? E - 1 - r : r) * T + tp];
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(T) : "memory");
  }
#pragma unroll
  for (int r = 0; r < E; r++) c[r] = keep_min ? min(c[r], o[r]) : max(c[r], o[r]);
}

// Ascending sort of n2 = E*T distinct words by T threads (T a multiple of 32, threads
// t = 0..T-1 of consecutive warps; all of them must call).  Thread t ends with the sorted
// elements t*E .. t*E+E-1 in c[0..E-1].
template <int E>
__device__ __forceinline__ void block_sort_u32(uint32_t (&c)[E], uint32_t t, uint32_t T,
                                               uint32_t* s_x, uint32_t bar_id) {
  reg_sort_u32<E>(c);
  const uint32_t n2 = E * T;
  for (uint32_t k = 2u * E; k <= n2; k <<= 1) {
    const uint32_t kt = k / E;  // threads per merged block
    xthread_step_u32<E, true>(c, t, T, kt - 1u, (t & (kt >> 1)) == 0u, s_x, bar_id);
    for (uint32_t jt = kt >> 2; jt > 0u; jt >>= 1)
      xthread_step_u32<E, false>(c, t, T, jt, (t & jt) == 0u, s_x, bar_id);
    reg_clean_u32<E>(c);
  }
}

constexpr uint32_t kFastRunMax = 16;  // slots on either side of one that a run of colliding truncated
                                      // keys may reach and still be repaired in place
constexpr uint32_t kFastMinN2 = 128;  // shortest list the composite sort takes (32 threads x 4)

struct FastRankState {
  uint32_t umin, umax;  // sortable keys of the real candidates
};

// composite words of the nv real candidates, sorted ascending into s_cmp[0 .. n2)
template <int E>
__device__ __forceinline__ void fast_sort_composites(uint32_t t, uint32_t sub_bar, const float* s_val,
                                                     uint32_t* s_cmp, uint32_t nv, uint32_t n2,
                                                     uint32_t umin, uint32_t shift, uint32_t sb) {
  const uint32_t T = n2 / E;
  if (t < T) {
    uint32_t c[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
      const uint32_t e = r * T + t;  // any assignment of slots to threads will do
      c[r] = e < nv ? ((((sortable_key(s_val[e]) - umin) >> shift) << sb) | e) : 0xFFFFFFFFu;
    }
    block_sort_u32<E>(c, t, T, s_cmp, sub_bar);
#pragma unroll
    for (int r = 0; r < E; r++) s_cmp[t * E + r] = c[r];
  }
}

// Fast ranking of the nv real candidates of one query (all < 1e7, finite).
//   s_val[a]: distance of candidate slot a (candidate order, untouched)
//   s_cmp   : scratch, n2 words (n2 = pow2ceil(nv), kFastMinN2 <= n2 <= 4096)
//   s_fix   : scratch bitmap, one bit per result slot (max_vec / 32 words)
//   cand[a] : bin-order position of candidate slot a (identifies the vector), ids[pos] its id
// Writes the first k results (pads after nv).  Returns (to every thread of the group) a
// flag word: bit 0 = some bit-equal distances belong to different vectors (their order in
// the output is by candidate slot, not yet the network's), bit 1 = a run of colliding keys
// was too long to repair (output incomplete).  kEmitW = consecutive result slots per
// thread in the emit pass (a multiple of 4; g.n * kEmitW >= max_vec).
template <int kEmitW>
__device__ __forceinline__ uint32_t fast_rank_emit(const Grp& g, uint32_t sub_bar, const float* s_val,
                                                   uint32_t* s_cmp, uint32_t* s_fix, uint32_t* s_flag,
                                                   uint32_t nv, uint32_t n2, uint32_t k,
                                                   FastRankState st, float* out_dist,
                                                   uint32_t* out_idx, const uint32_t* __restrict__ cand,
                                                   const uint32_t* __restrict__ ids,
                                                   unsigned long long* ph) {
  const uint32_t t = g.t;
  const uint32_t range = st.umax - st.umin;
  // composite word = key << sb | slot: sb = log2(n2) slot bits, the other 32 - sb bits hold
  // the order-preserving key (shorter lists get longer keys, hence fewer collisions)
  const uint32_t sb = 31u - (uint32_t)__clz((int)n2);
  const uint32_t smask = n2 - 1u;
  const uint32_t bits = 32u - (uint32_t)__clz((int)range);
  const uint32_t shift = bits > 32u - sb ? bits - (32u - sb) : 0u;
  if (t == 0) *s_flag = 0;
  if (t < 128u) s_fix[t] = 0;  // one bit per result slot (<= 4096)
  // 16 elements per thread cost the fewest instructions (the kernel is issue-bound; other
  // thread groups of the CTA cover the latency); short lists take 4 so that a warp is filled
  if (n2 >= 512u)
    fast_sort_composites<16>(t, sub_bar, s_val, s_cmp, nv, n2, st.umin, shift, sb);
  else
    fast_sort_composites<4>(t, sub_bar, s_val, s_cmp, nv, n2, st.umin, shift, sb);
  g.sync();
  if (ph && t == 0) ph[3] = clock64();
  // ---- emit.  Sorted slot e holds composite s_cmp[e]; neighbours whose truncated keys collide
  // are in candidate-slot order and may have to be re-ordered by their full keys.  A thread
  // looks at kEmitW consecutive slots plus one neighbour on each side, with all loads of a
  // level issued together (slot -> distance and position, position -> id).
  // slots just past k can still move below k when a run of equal keys straddles k
  const uint32_t e_end = k < nv ? min(nv, k + kFastRunMax) : k;
  const uint32_t e0 = t * kEmitW;
  uint32_t flag = 0;
  uint32_t id[kEmitW + 2];
  float v[kEmitW + 2];
  const bool active = e0 < e_end;
  if (active) {
    uint32_t eq20 = 0;  // bit b: slots e0-1+b and e0+b are real and share their truncated key
    {
      uint32_t c[kEmitW + 2];
#pragma unroll
      for (int i = 0; i < kEmitW + 2; i++) {
        const uint32_t e = e0 + i - 1u;  // e0 == 0: wraps, fails the range test
        c[i] = e < nv ? s_cmp[e] : 0xFFFFFFFFu;
      }
#pragma unroll
      for (int b = 0; b <= kEmitW; b++) {
        const uint32_t er = e0 + b;
        if (er < nv && er > 0u && ((c[b] ^ c[b + 1]) >> sb) == 0u) eq20 |= 1u << b;
      }
      uint32_t ps[kEmitW + 2];
#pragma unroll
      for (int i = 0; i < kEmitW + 2; i++) {
        const uint32_t e = e0 + i - 1u;
        const uint32_t a = c[i] & smask;
        v[i] = e < nv ? s_val[a] : kPadDist;
        ps[i] = e < nv ? __ldg(cand + a) : 0u;
      }
#pragma unroll
      for (int i = 0; i < kEmitW + 2; i++) {
        const uint32_t e = e0 + i - 1u;
        id[i] = e < nv ? __ldg(ids + ps[i]) : kPadIdx;
      }
    }
    uint32_t bad = 0;
#pragma unroll
    for (int b = 0; b <= kEmitW; b++) {
      if ((eq20 >> b) & 1u) {
        if (sortable_key(v[b]) != sortable_key(v[b + 1]))
          bad |= 1u << b;
        else if (id[b] != id[b + 1])
          flag |= 1u;  // bit-equal distances of different vectors
      }
    }
    // mark every slot of a run that holds unequal full keys (rare)
    if (bad) flag |= 4u;
    while (bad) {
      const uint32_t b = __ffs(bad) - 1u;
      bad &= bad - 1u;
      const uint32_t e = e0 + b;  // right slot of the bad boundary
      const uint32_t ce = s_cmp[e];
      uint32_t rs = e - 1u, re = e + 1u;
      while (rs > 0u && e - rs <= kFastRunMax && ((ce ^ s_cmp[rs - 1]) >> sb) == 0u) rs--;
      while (re < nv && re - e <= kFastRunMax && ((ce ^ s_cmp[re]) >> sb) == 0u) re++;
      if (re - rs > kFastRunMax + 1u) flag |= 2u;  // every member must see the whole run in its window
      for (uint32_t j = rs; j < re; j++) atomicOr(&s_fix[j >> 5], 1u << (j & 31u));
    }
  }
  if (flag) atomicOr(s_flag, flag);
  g.sync();
  if (ph && t == 0) ph[4] = clock64();
  const uint32_t f1 = *s_flag;
  if (f1 & 2u) {  // a run too long to repair: the caller ranks this query with the network
    g.sync();
    return 2u;
  }
  uint32_t fix = 0;
  if (active) {
    if (f1 & 4u) fix = (s_fix[e0 >> 5] >> (e0 & 31u)) & ((1u << kEmitW) - 1u);
    const bool vec_ok = (((uintptr_t)(out_dist + e0) | (uintptr_t)(out_idx + e0)) & 15u) == 0u;
    if (fix == 0u && vec_ok && e0 + kEmitW <= k) {
      // (pads beyond nv already carry kPadDist / kPadIdx)
      float4* od = reinterpret_cast<float4*>(out_dist + e0);
      uint4* oi = reinterpret_cast<uint4*>(out_idx + e0);
#pragma unroll
      for (int q = 0; q < kEmitW / 4; q++) {
        od[q] = make_float4(v[4 * q + 1], v[4 * q + 2], v[4 * q + 3], v[4 * q + 4]);
        oi[q] = make_uint4(id[4 * q + 1], id[4 * q + 2], id[4 * q + 3], id[4 * q + 4]);
      }
    } else {
#pragma unroll
      for (int i = 1; i <= kEmitW; i++) {
        const uint32_t e = e0 + i - 1u;
        if (e < e_end && e < k && !((fix >> (i - 1)) & 1u)) {
          out_dist[e] = v[i];
          out_idx[e] = id[i];
        }
      }
      // slots of runs with unequal full keys: rank inside the run by (full key, slot).  The
      // slots are sorted by their truncated keys, so the run is exactly the neighbours with an
      // equal truncated key: a fixed window, no data-dependent walk (longer runs were flagged).
      uint32_t fx = fix;
      while (fx) {
        const uint32_t e = e0 + (uint32_t)__ffs(fx) - 1u;
        fx &= fx - 1u;
        if (e >= e_end) continue;
        const uint32_t ce = s_cmp[e];
        const uint32_t a = ce & smask;
        const float ve = s_val[a];
        const uint32_t u = sortable_key(ve);
        uint32_t before = 0, rank = 0;
#pragma unroll 8
        for (int d = -(int)kFastRunMax; d < (int)kFastRunMax; d++) {
          const uint32_t j = e + (uint32_t)(d < 0 ? d : d + 1);  // wraps below 0: fails j < nv
          const uint32_t cj = j < nv ? s_cmp[j] : ~ce;
          const bool in_run = ((cj ^ ce) >> sb) == 0u;
          const uint32_t uj = sortable_key(s_val[in_run ? (cj & smask) : a]);
          if (in_run) {
            if (d < 0) before++;
            if (uj < u || (uj == u && d < 0)) rank++;
          }
        }
        const uint32_t dst = e - before + rank;
        if (dst < k) {
          out_dist[dst] = ve;
          out_idx[dst] = __ldg(ids + __ldg(cand + a));
        }
      }
    }
  }
  if (f1 & 4u) {
    // re-ordered runs: bit-equal distances of different vectors are now adjacent in the output
    g.sync();
    if (fix) {
#pragma unroll
      for (int i = 0; i < kEmitW; i++) {
        const uint32_t e = e0 + i;
        if (((fix >> i) & 1u) && e > 0u && e < k && e < nv) {
          if (out_dist[e] == out_dist[e - 1] && out_idx[e] != out_idx[e - 1]) atomicOr(s_flag, 1u);
        }
      }
    }
  }
  g.sync();
  const uint32_t f = *s_flag & 3u;
  g.sync();
  return f;
}

}  // namespace pqtb
