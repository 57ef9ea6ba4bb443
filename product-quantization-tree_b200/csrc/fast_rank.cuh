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
// m < 32: shuffles; otherwise through s_x ([E][Ta] words) with the sub-group barrier.
// Threads >= Ta hold only pad words (0xFFFFFFFF) and do not take part.
template <int E, bool REV>
__device__ __forceinline__ void xthread_step_u32(uint32_t (&c)[E], uint32_t t, uint32_t Ta,
                                                 uint32_t m, bool keep_min, uint32_t* s_x,
                                                 uint32_t bar_id) {
  uint32_t o[E];
  if (m < 32u) {
#pragma unroll
    for (int r = 0; r < E; r++) o[r] = __shfl_xor_sync(0xffffffffu, c[REV ? E - 1 - r : r], m);
  } else {
#pragma unroll
    for (int r = 0; r < E; r++) s_x[r * Ta + t] = c[r];
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(Ta) : "memory");
    const uint32_t tp = t ^ m;
    if (tp < Ta) {
#pragma unroll
      for (int r = 0; r < E; r++) o[r] = s_x[(REV ? E - 1 - r : r) * Ta + tp];
    } else {
#pragma unroll
      for (int r = 0; r < E; r++) o[r] = 0xFFFFFFFFu;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(Ta) : "memory");
  }
  // min or max by a predicate: one predicated pair per element instead of min + max + select
  // (the hardware's VIMNMX takes the choice as a predicate operand; nvcc does not emit that
  // form for the ternary)
  const uint32_t km = keep_min ? 1u : 0u;
#pragma unroll
  for (int r = 0; r < E; r++) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p min.u32 %0, %0, %1;\n\t@!p max.u32 %0, %0, %1;\n\t}"
        : "+r"(c[r])
        : "r"(o[r]), "r"(km));
  }
}

// Ascending sort of n2 = E*T distinct words of which only the first E*Ta (Ta a multiple of
// 32, Ta <= T) can differ from the pad word 0xFFFFFFFF; run by threads t = 0..Ta-1 (all of
// them must call).  Pads are the maximum and the network only moves minima down, so the
// threads >= Ta would never change: they are left out.  Thread t ends with the sorted
// elements t*E .. t*E+E-1 in c[0..E-1].
template <int E>
__device__ __forceinline__ void block_sort_u32(uint32_t (&c)[E], uint32_t t, uint32_t T, uint32_t Ta,
                                               uint32_t* s_x, uint32_t bar_id) {
  reg_sort_u32<E>(c);
  for (uint32_t kt = 2u; kt <= T; kt <<= 1) {  // kt: threads per merged block
    xthread_step_u32<E, true>(c, t, Ta, kt - 1u, (t & (kt >> 1)) == 0u, s_x, bar_id);
    for (uint32_t jt = kt >> 2; jt > 0u; jt >>= 1)
      xthread_step_u32<E, false>(c, t, Ta, jt, (t & jt) == 0u, s_x, bar_id);
    reg_clean_u32<E>(c);
  }
}

// Where a candidate slot's vector identity and id come from.  Fused single-GPU kernel:
// slot -> bin-order position (cand) -> id (ids).  DIRECT (candidates assembled from shards):
// ids[slot] is the id itself and doubles as the identity.
template <bool DIRECT>
__device__ __forceinline__ uint32_t slot_ident(const uint32_t* __restrict__ cand,
                                               const uint32_t* __restrict__ ids, uint32_t a) {
  return DIRECT ? __ldg(ids + a) : __ldg(cand + a);
}
template <bool DIRECT>
__device__ __forceinline__ uint32_t ident_id(const uint32_t* __restrict__ ids, uint32_t ident) {
  return DIRECT ? ident : __ldg(ids + ident);
}

constexpr uint32_t kFastRunMax = 16;  // slots on either side of one that a run of colliding truncated
                                      // keys may reach and still be repaired in place
constexpr uint32_t kFastMinN2 = 128;  // shortest list the composite sort takes (32 threads x 4)

struct FastRankState {
  uint32_t umin, umax;  // sortable keys of the real candidates
};

// Sort + emit by the sorter threads, E sorted slots each, straight from their registers.
// Every slot is stored at its sorted position; slots of runs (equal truncated keys) that hold
// unequal full keys are marked in s_fix and re-ordered afterwards by the caller.
//   flag bits: 1 = bit-equal distances of different vectors seen, 2 = run too long to
//   repair, 4 = some slots are marked in s_fix
template <int E, bool DIRECT>
__device__ __forceinline__ uint32_t fast_sort_emit(uint32_t t, uint32_t gn, uint32_t sub_bar,
                                                   const float* s_val, uint32_t* s_cmp,
                                                   uint32_t* s_fix, uint32_t nv, uint32_t n2,
                                                   uint32_t k, uint32_t umin, uint32_t shift,
                                                   uint32_t sb, float* out_dist, uint32_t* out_idx,
                                                   const uint32_t* __restrict__ cand,
                                                   const uint32_t* __restrict__ ids,
                                                   unsigned long long* ph = nullptr) {
  constexpr int CH = E < 8 ? E : 8;  // slots handled together (loads of a level issued together)
  const uint32_t T = n2 / E;
  const uint32_t Ta = min(T, ((nv + E - 1u) / E + 31u) & ~31u);  // sorter threads
  const uint32_t smask = n2 - 1u;
  uint32_t flag = 0;
  if (t < Ta) {
    uint32_t c[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
      const uint32_t e = t * E + r;
      c[r] = e < nv ? ((((sortable_key(s_val[e]) - umin) >> shift) << sb) | e) : 0xFFFFFFFFu;
    }
    block_sort_u32<E>(c, t, T, Ta, s_cmp, sub_bar);
    if (ph && t == 0) ph[1] = clock64();  // ranking-only kernels: end of the sort proper
    // publish: neighbours' edge slots and the repair windows read s_cmp
#pragma unroll
    for (int r = 0; r < E; r++) s_cmp[t * E + r] = c[r];
    asm volatile("bar.sync %0, %1;" ::"r"(sub_bar), "r"(Ta) : "memory");
    // left neighbour of the thread's first slot
    uint32_t c_prev = t > 0u ? s_cmp[t * E - 1u] : 0xFFFFFFFFu;
    float v_prev = 0.f;
    uint32_t p_prev = 0u;
    if (t > 0u && t * E - 1u < nv) {
      const uint32_t a = c_prev & smask;
      v_prev = s_val[a];
      p_prev = slot_ident<DIRECT>(cand, ids, a);
    }
#pragma unroll
    for (int h = 0; h < E / CH; h++) {
      const uint32_t e0 = t * E + h * CH;
      float v[CH];
      uint32_t ps[CH], id[CH];
#pragma unroll
      for (int i = 0; i < CH; i++) {
        const uint32_t a = c[h * CH + i] & smask;
        const bool real = e0 + i < nv;
        v[i] = real ? s_val[a] : kPadDist;
        ps[i] = real ? slot_ident<DIRECT>(cand, ids, a) : 0u;
      }
#pragma unroll
      for (int i = 0; i < CH; i++) id[i] = (e0 + i < nv) ? ident_id<DIRECT>(ids, ps[i]) : kPadIdx;
      // boundaries between slot e0+i-1 and e0+i
      uint32_t bad = 0;
#pragma unroll
      for (int i = 0; i < CH; i++) {
        const uint32_t cl = i ? c[h * CH + i - 1] : c_prev;
        const float vl = i ? v[i - 1] : v_prev;
        const uint32_t pl = i ? ps[i - 1] : p_prev;
        const uint32_t er = e0 + i;
        if (er < nv && er > 0u && ((cl ^ c[h * CH + i]) >> sb) == 0u) {
          if (sortable_key(vl) != sortable_key(v[i]))
            bad |= 1u << i;
          else if (pl != ps[i])
            flag |= 1u;  // bit-equal distances of different vectors
        }
      }
      c_prev = c[h * CH + CH - 1];
      v_prev = v[CH - 1];
      p_prev = ps[CH - 1];
      // mark every slot of a run that holds unequal full keys (rare)
      if (bad) flag |= 4u;
      while (bad) {
        const uint32_t e = e0 + (uint32_t)__ffs(bad) - 1u;  // right slot of the bad boundary
        bad &= bad - 1u;
        const uint32_t ce = s_cmp[e];
        uint32_t rs = e - 1u, re = e + 1u;
        while (rs > 0u && e - rs <= kFastRunMax && ((ce ^ s_cmp[rs - 1]) >> sb) == 0u) rs--;
        while (re < nv && re - e <= kFastRunMax && ((ce ^ s_cmp[re]) >> sb) == 0u) re++;
        if (re - rs > kFastRunMax + 1u) flag |= 2u;  // every member must see the whole run in its window
        for (uint32_t j = rs; j < re; j++) atomicOr(&s_fix[j >> 5], 1u << (j & 31u));
      }
      // store at the sorted position (marked slots are overwritten by the repair pass)
      const bool vec_ok = (CH % 4 == 0) && e0 + CH <= k &&
                          (((uintptr_t)(out_dist + e0) | (uintptr_t)(out_idx + e0)) & 15u) == 0u;
      if (vec_ok) {
        float4* od = reinterpret_cast<float4*>(out_dist + e0);
        uint4* oi = reinterpret_cast<uint4*>(out_idx + e0);
#pragma unroll
        for (int q = 0; q < CH / 4; q++) {
          od[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          oi[q] = make_uint4(id[4 * q], id[4 * q + 1], id[4 * q + 2], id[4 * q + 3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CH; i++) {
          if (e0 + i < k) {
            out_dist[e0 + i] = v[i];
            out_idx[e0 + i] = id[i];
          }
        }
      }
    }
    // every thread is a sorter: the pads behind the sorters' slots are theirs as well
    if (Ta == gn) {
      for (uint32_t e = Ta * E + t; e < k; e += gn) {
        out_dist[e] = kPadDist;
        out_idx[e] = kPadIdx;
      }
    }
  } else {
    // the other threads of the group write the pads behind the sorters' slots
    const uint32_t first = Ta * E;
    const uint32_t nt = gn - Ta;
    for (uint32_t e = first + (t - Ta); e < k; e += nt) {
      out_dist[e] = kPadDist;
      out_idx[e] = kPadIdx;
    }
  }
  return flag;
}

// Fast ranking of the nv real candidates of one query (all < 1e7, finite).
//   s_val[a]: distance of candidate slot a (candidate order, untouched)
//   s_cmp   : scratch, n2 words (n2 = pow2ceil(nv), kFastMinN2 <= n2 <= 4096)
//   s_fix   : scratch bitmap, one bit per result slot (128 words)
//   cand[a] : bin-order position of candidate slot a (identifies the vector), ids[pos] its id
// Writes the first k results (pads after nv).  Returns (to every thread of the group) a
// flag word: bit 0 = some bit-equal distances may belong to different vectors (their order
// in the output is by candidate slot, not yet the network's; tie_resolve checks), bit 1 = a run of colliding keys
// was too long to repair (output incomplete).  g.n * 16 >= max_vec.
template <bool DIRECT>
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
  // (s_flag and s_fix were cleared by the caller before the barrier that ended the scan)
  // Elements per thread: 16 cost the fewest instructions, fewer shorten the dependent chain
  // of a query (the sorters are the only busy warps of their group in this phase), which is
  // what counts while the SM has issue slots to spare.
  uint32_t flag;
  if (n2 >= 4096u)
    flag = fast_sort_emit<16, DIRECT>(t, g.n, sub_bar, s_val, s_cmp, s_fix, nv, n2, k, st.umin, shift, sb,
                                      out_dist, out_idx, cand, ids, ph && ph[1] == 0 ? ph : nullptr);
  else if (n2 >= 2048u)
    flag = fast_sort_emit<8, DIRECT>(t, g.n, sub_bar, s_val, s_cmp, s_fix, nv, n2, k, st.umin, shift, sb,
                                     out_dist, out_idx, cand, ids);
  else
    flag = fast_sort_emit<4, DIRECT>(t, g.n, sub_bar, s_val, s_cmp, s_fix, nv, n2, k, st.umin, shift, sb,
                                     out_dist, out_idx, cand, ids);
  if (flag) atomicOr(s_flag, flag);
  g.sync();
  if (ph && t == 0) ph[4] = clock64();
  const uint32_t f1 = *s_flag;
  if ((f1 & 6u) != 4u) return f1 & 3u;  // nothing to repair, or not repairable here
  // ---- repair pass: slots of runs with unequal full keys get their rank inside the run by
  // (full key, slot).  The slots are sorted by their truncated keys, so the run is exactly
  // the neighbours with an equal truncated key: a fixed window, no data-dependent walk.
  // slots just past k can still move below k when a run straddles k
  const uint32_t e_end = k < nv ? min(nv, k + kFastRunMax) : min(k, nv);
  // The marked slots are compacted into a list (which takes the place of the bitmap) so that
  // every marked slot gets its own thread: thread w < 128 owns word w of the bitmap.
  uint32_t tie = 0;
  uint32_t word = 0;
  if (t < 128u) {
    word = s_fix[t];
    const uint32_t lo = t << 5;
    if (lo >= e_end) word = 0u;
    else if (lo + 32u > e_end) word &= (1u << (e_end - lo)) - 1u;
  }
  // exclusive prefix of the popcounts over threads 0..127 (4 warps; totals through s_flag+...)
  uint32_t cntw = __popc(word);
  uint32_t incl = cntw;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31u) >= (uint32_t)o) incl += n;
  }
  g.sync();  // every word has been read: the bitmap area becomes scratch
  uint32_t* s_tot = s_fix;                                  // [4] warp totals
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_fix + 4);  // [248] marked slots
  constexpr uint32_t kListCap = 248;
  if (t < 128u && (t & 31u) == 31u) s_tot[t >> 5] = incl;
  g.sync();
  uint32_t base = 0, total = 0;
  {
    const uint32_t t0 = s_tot[0], t1 = s_tot[1], t2 = s_tot[2], t3 = s_tot[3];
    total = t0 + t1 + t2 + t3;
    const uint32_t wq = t >> 5;
    base = (wq > 0u ? t0 : 0u) + (wq > 1u ? t1 : 0u) + (wq > 2u ? t2 : 0u);
  }
  const uint32_t my_first = base + incl - cntw;  // list index of this word's first marked slot
  for (uint32_t r0 = 0; r0 < total; r0 += kListCap) {
    if (r0) g.sync();  // the previous round's list has been consumed
    if (t < 128u) {
      uint32_t wv = word, li = my_first;
      while (wv) {
        const uint32_t bpos = __ffs(wv) - 1u;
        wv &= wv - 1u;
        if (li >= r0 && li < r0 + kListCap) s_list[li - r0] = (uint16_t)((t << 5) + bpos);
        li++;
      }
    }
    g.sync();
    const uint32_t nlist = min(kListCap, total - r0);
    for (uint32_t li = t; li < nlist; li += g.n) {
      const uint32_t e = s_list[li];
      const uint32_t ce = s_cmp[e];
      const uint32_t a = ce & smask;
      const uint32_t pe = slot_ident<DIRECT>(cand, ids, a);  // in flight during the walk
      const float ve = s_val[a];
      const uint32_t u = sortable_key(ve);
      // walk the run outwards from e (runs are short; longer ones were flagged)
      uint32_t before = 0, rank = 0;
      for (uint32_t d = 1; d <= kFastRunMax && d <= e; d++) {
        const uint32_t cj = s_cmp[e - d];
        if (((cj ^ ce) >> sb) != 0u) break;
        const uint32_t uj = sortable_key(s_val[cj & smask]);
        before++;
        if (uj <= u) rank++;
        if (uj == u) tie = 1u;  // a duplicate of this vector or another vector: tie_resolve looks
      }
      for (uint32_t d = 1; d <= kFastRunMax && e + d < nv; d++) {
        const uint32_t cj = s_cmp[e + d];
        if (((cj ^ ce) >> sb) != 0u) break;
        const uint32_t uj = sortable_key(s_val[cj & smask]);
        if (uj < u) rank++;
        if (uj == u) tie = 1u;
      }
      const uint32_t dst = e - before + rank;
      if (dst < k) {
        out_dist[dst] = ve;
        out_idx[dst] = ident_id<DIRECT>(ids, pe);
      }
    }
  }
  if (tie) atomicOr(s_flag, 1u);
  g.sync();
  return *s_flag & 3u;
}

}  // namespace pqtb
