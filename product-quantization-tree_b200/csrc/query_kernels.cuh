// query_kernels.cuh -- the queryKNN chain as sm_100a kernels.
//
//   tables_warp_kernel / tables_kernel   Steps A+B+C (getKBestAssignment,
//                       getLineAssignment, getKBestAssignment2); lut_kernel = Step B alone
//   bins2_kernel        Steps D+E1 for p > 4 (getBins / selectBinKernelFast2,
//                       getKVectorIDsKernelFast); bins3/bins4 live in rerank_kernels.cuh
//   adc_warp_step & co  the ADC arithmetic of Step E2 (rerankKernelFast), shared by
//                       every scan kernel; adc_scan_kernel = stand-alone scan for
//                       shapes whose fused/streaming variants do not fit
//
// Citations are file:line into /root/reference/pqt/PerturbationProTree.cu unless
// another file is named.  Semantics follow SURVEY.md App. B and are checked
// bit-for-bit against oracle/pqt_oracle.c.
#pragma once
#include "common.cuh"
#include "fast_rank.cuh"

namespace pqtb {

// ============================================================================
// in-shared-memory bitonic network, identical compare-exchange order to
// pqt/bitonicSort.cuh:16-78 (ascending iff (i & k) == 0, swap on strict > / <).
// Sorts `narr` independent arrays of n (power of two) elements laid out back to
// back.  All threads of the block must call.
// ============================================================================
__device__ __forceinline__ void bitonic_smem(float* val, uint32_t* idx, uint32_t n, uint32_t narr) {
  const uint32_t half = n >> 1;
  const uint32_t total = half * narr;
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
        uint32_t arr = e / half, t = e - arr * half;
        uint32_t i = 2 * t - (t & (j - 1));  // t = q*j + r -> i = q*2j + r, (i & j) == 0
        uint32_t a = arr * n + i, b = a + j;
        const float va = val[a], vb = val[b];
        const uint32_t ia = idx[a], ib = idx[b];
        const bool sw = ((i & k) == 0) ? (va > vb) : (va < vb);
        val[a] = sw ? vb : va;  // branch-free: divergent swaps cost more than the stores
        val[b] = sw ? va : vb;
        idx[a] = sw ? ib : ia;
        idx[b] = sw ? ia : ib;
      }
      __syncthreads();
    }
  }
}

// ============================================================================
// Steps A + B + C: one CTA per query.
// ============================================================================
struct TablesArgs {
  const float* Q;    // [QN][dim]
  const float* cb1;  // [c1][dim]
  const float* cb2;  // [p][c1][c2][vl]
  uint32_t QN, dim, p, c1, c2, LP, k1, vl, sl;
  uint32_t npA;  // pow2ceil(c1)
  uint32_t npC;  // pow2ceil(k1*c2)
  uint32_t m;    // min(c2*k1, 16): entries of the sorted Step-C list reachable by Step D
  float* lut_dup;   // [QN][c1][32]: index c*32 + j*LP + lp, replicated j < 32/LP (see adc_scan)
  uint32_t* idx16;  // [QN][p][16]
  // optional debug outputs (canonical reference layouts), may be null
  uint32_t* dbg_assign;  // [QN][k1][p]
  float* dbg_lut;        // [QN][LP][c1]
  float* dbg_aval;       // [QN][p][k1*c2]
  uint32_t* dbg_aidx;    // [QN][p][k1*c2]
  // optional: first top_n entries of every sorted Step-C list (1-B variant), may be null
  float* top_val;     // [QN][p][top_n]
  uint32_t* top_idx;  // [QN][p][top_n]
  uint32_t top_n;
};

// dynamic smem: q[dim] | val[p*npMax] | idx[p*npMax] | assign[k1*p]
__global__ void __launch_bounds__(128) tables_kernel(TablesArgs a) {
  extern __shared__ float smem_f[];
  const uint32_t npMax = a.npA > a.npC ? a.npA : a.npC;
  float* sq = smem_f;
  float* sval = sq + a.dim;
  uint32_t* sidx = reinterpret_cast<uint32_t*>(sval + a.p * npMax);
  uint32_t* sassign = sidx + a.p * npMax;

  for (uint32_t qi = blockIdx.x; qi < a.QN; qi += gridDim.x) {
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < a.dim; t += blockDim.x) sq[t] = a.Q[(size_t)qi * a.dim + t];
    // Step A sort buffers: pad to 1e7 (:7185)
    for (uint32_t e = threadIdx.x; e < a.p * a.npA; e += blockDim.x) {
      sval[e] = kPadSortA;
      sidx[e] = kPadIdx;
    }
    __syncthreads();

    // ---- Step A distances (:7146-7176): val[part][c] over vl dims
    for (uint32_t e = threadIdx.x; e < a.p * a.c1; e += blockDim.x) {
      uint32_t part = e / a.c1, c = e - part * a.c1;
      sval[part * a.npA + c] =
          seg_dist_dyn(sq + part * a.vl, a.cb1 + (size_t)c * a.dim + part * a.vl, a.vl);
      sidx[part * a.npA + c] = c;
    }
    // ---- Step B (:7764-7796): lut[lp][c] over dim/LP dims
    {
      const uint32_t R = 32 / a.LP;
      float* lut = a.lut_dup + (size_t)qi * a.c1 * 32;
      for (uint32_t e = threadIdx.x; e < a.LP * a.c1; e += blockDim.x) {
        uint32_t c = e / a.LP, lp = e - c * a.LP;
        float v = seg_dist_dyn(sq + lp * a.sl, a.cb1 + (size_t)c * a.dim + lp * a.sl, a.sl);
        for (uint32_t j = 0; j < R; j++) lut[c * 32 + j * a.LP + lp] = v;
        if (a.dbg_lut) a.dbg_lut[((size_t)qi * a.LP + lp) * a.c1 + c] = v;
      }
    }
    __syncthreads();
    bitonic_smem(sval, sidx, a.npA, a.p);  // :7194
    // assign[k][part] (:7203-7209)
    for (uint32_t e = threadIdx.x; e < a.k1 * a.p; e += blockDim.x) {
      uint32_t k = e / a.p, part = e - k * a.p;
      uint32_t v = sidx[part * a.npA + k];
      sassign[e] = v;
      if (a.dbg_assign) a.dbg_assign[(size_t)qi * a.k1 * a.p + e] = v;
    }
    __syncthreads();

    // ---- Step C (:1592-1621): k1 cells x c2 centroids per part, pad 1e9 (:1627)
    const uint32_t n = a.k1 * a.c2;
    for (uint32_t e = threadIdx.x; e < a.p * a.npC; e += blockDim.x) {
      uint32_t part = e / a.npC, i = e - part * a.npC;
      float v = kPadSortC;
      uint32_t id = kPadIdx;
      if (i < n) {
        uint32_t k = i / a.c2, l2 = i - k * a.c2;
        uint32_t l1 = sassign[k * a.p + part];
        const float* cb = a.cb2 + ((size_t)(part * a.c1 + l1) * a.c2 + l2) * a.vl;  // getCBIdx
        v = seg_dist_dyn(sq + part * a.vl, cb, a.vl);
        id = l2 + l1 * a.c2;
      }
      sval[e] = v;
      sidx[e] = id;
    }
    __syncthreads();
    bitonic_smem(sval, sidx, a.npC, a.p);  // :1639
    for (uint32_t e = threadIdx.x; e < a.p * 16; e += blockDim.x) {
      uint32_t part = e >> 4, r = e & 15;
      a.idx16[(size_t)qi * a.p * 16 + e] = (r < a.m) ? sidx[part * a.npC + r] : 0u;
    }
    if (a.dbg_aval) {
      for (uint32_t e = threadIdx.x; e < a.p * n; e += blockDim.x) {
        uint32_t part = e / n, i = e - part * n;
        a.dbg_aval[(size_t)qi * a.p * n + e] = sval[part * a.npC + i];
        a.dbg_aidx[(size_t)qi * a.p * n + e] = sidx[part * a.npC + i];
      }
    }
    if (a.top_val) {
      for (uint32_t e = threadIdx.x; e < a.p * a.top_n; e += blockDim.x) {
        uint32_t part = e / a.top_n, i = e - part * a.top_n;
        a.top_val[(size_t)qi * a.p * a.top_n + e] = sval[part * a.npC + i];
        a.top_idx[(size_t)qi * a.p * a.top_n + e] = sidx[part * a.npC + i];
      }
    }
  }
}

// ============================================================================
// Step B alone (lineAssignmentKernel :7739-7799) for a range of queries: in the multi-GPU
// path Steps A/C/D of a query run on one rank only, but every shard needs the query's LUT
// to scan its own candidates; recomputing it (c1*LP short segment distances) is cheaper
// than shipping it.  Same arithmetic and output layout as the Step-B part of the tables
// kernels.  blockDim = 128, one CTA per query, cb1T = cb1 transposed [dim][c1].
// ============================================================================
__global__ void __launch_bounds__(128) lut_kernel(const float* Q, const float* cb1T, uint32_t q_begin,
                                                  uint32_t q_end, uint32_t dim, uint32_t c1,
                                                  uint32_t LP, uint32_t sl, float* lut_dup) {
  extern __shared__ float smem_f[];
  float* sq = smem_f;          // [dim]
  float* s_lut = sq + dim;     // [c1*32]
  const uint32_t R = 32 / LP;
  for (uint32_t qi = q_begin + blockIdx.x; qi < q_end; qi += gridDim.x) {
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < dim; t += blockDim.x) sq[t] = Q[(size_t)qi * dim + t];
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < LP * c1; e += blockDim.x) {
      const uint32_t lp = e / c1, c = e - lp * c1;  // c fastest: coalesced cb1T reads
      float s[128];
      const float* cb = cb1T + (size_t)(lp * sl) * c1 + c;
      float v;
      switch (sl) {
#define PQTB_LUT_CASE(L)                                                         \
  case L: {                                                                      \
    _Pragma("unroll") for (int t = 0; t < L; t++) {                              \
      float d = __fsub_rn(sq[lp * L + t], __ldg(cb + (size_t)t * c1));           \
      s[t] = __fmul_rn(d, d);                                                    \
    }                                                                            \
    _Pragma("unroll") for (int stride = L / 2; stride > 0; stride >>= 1) {       \
      _Pragma("unroll") for (int j = 0; j < stride; j++) s[j] = __fadd_rn(s[j], s[j + stride]); \
    }                                                                            \
    v = s[0];                                                                    \
  } break;
        PQTB_LUT_CASE(1)
        PQTB_LUT_CASE(2)
        PQTB_LUT_CASE(4)
        PQTB_LUT_CASE(8)
        PQTB_LUT_CASE(16)
        PQTB_LUT_CASE(32)
        PQTB_LUT_CASE(64)
        default:
          PQTB_LUT_CASE(128)
#undef PQTB_LUT_CASE
      }
      for (uint32_t j = 0; j < R; j++) s_lut[c * 32 + j * LP + lp] = v;
    }
    __syncthreads();
    float4* dst = reinterpret_cast<float4*>(lut_dup + (size_t)qi * c1 * 32);
    const float4* src = reinterpret_cast<const float4*>(s_lut);
    for (uint32_t e = threadIdx.x; e < c1 * 8; e += blockDim.x) dst[e] = src[e];
  }
}

// ============================================================================
// Bin directory: occupancy bitmap + rank.  Replaces probing the dense
// binCounts[hash_size] (1.6 GB) by one 32-byte sector of a hash_size/8-byte
// bitmap (50 MB, L2 resident), and the dense binPrefix by a compact array
// indexed by the rank of the non-empty bin.
//   bitmap[w]     bit b = (binCounts[32w + b] != 0)
//   rank_base[g]  number of non-empty bins before bin 256 g   (group = one sector)
//   cprefix[r]    binPrefix of the r-th non-empty bin; cprefix[nNonEmpty] = N
// ============================================================================
struct BinDir {
  const uint32_t* bitmap;
  const uint32_t* rank_base;
  const uint32_t* cprefix;
};

__device__ __forceinline__ bool dir_occupied(const BinDir& d, uint32_t bin) {
  return (__ldg(d.bitmap + (bin >> 5)) >> (bin & 31)) & 1u;
}

// (start, count) of a bin in the bin-ordered arrays; count 0 if empty.  The 8 bitmap words of
// the bin's 256-bin sector and the sector's rank base are fetched together (one dependent level
// before the prefix look-up, no data-dependent load loop).
__device__ __forceinline__ void dir_lookup(const BinDir& d, uint32_t bin, uint32_t& start,
                                           uint32_t& count) {
  const uint32_t g = bin >> 8, wi = (bin >> 5) & 7u;
  const uint4* sec = reinterpret_cast<const uint4*>(d.bitmap + (g << 3));
  const uint4 lo = __ldg(sec), hi = __ldg(sec + 1);
  uint32_t r = __ldg(d.rank_base + g);
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  uint32_t word = 0;
#pragma unroll
  for (uint32_t i = 0; i < 8; i++) {
    r += (i < wi) ? __popc(w[i]) : 0u;
    word = (i == wi) ? w[i] : word;
  }
  if (!((word >> (bin & 31)) & 1u)) {
    start = 0;
    count = 0;
    return;
  }
  r += __popc(word & ((1u << (bin & 31)) - 1u));
  start = __ldg(d.cprefix + r);
  count = __ldg(d.cprefix + r + 1) - start;
}

// block-wide exclusive scan of one uint per thread (blockDim.x <= 1024, multiple
// of 32); returns the exclusive prefix and the block total.  `warp_sums` is 32
// words of shared memory.
__device__ __forceinline__ uint32_t block_exscan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  __syncthreads();  // protect warp_sums reuse
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  uint32_t ws = (lane < nw) ? warp_sums[lane] : 0;
  uint32_t winc = ws;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
    if (lane >= d) winc += t;
  }
  total = __shfl_sync(0xffffffffu, winc, 31);
  uint32_t wbase = __shfl_sync(0xffffffffu, winc - ws, warp);
  return wbase + inc - v;
}

constexpr int kBinsThreads = 256;

// ============================================================================
// Step E2, distance part: the ADC scan over line codes (:5277-5329).
//
// Lane mapping (LP lanes per candidate, as the reference): lane = g*LP + lp.  A
// warp step evaluates R = 32/LP candidates; segment lp of candidate g reads
//   a2 = lut[p1][lane], b2 = lut[p2][lane], c2 = cbd[p2*c1 + p1][lane]
// from tables whose 32-float rows hold the LP values replicated R times, so the
// 32 lanes of a warp always hit 32 distinct banks (conflict-free by layout).
// The LP partial distances are summed with the same pairwise tree as
// warpReduceSum (:5183-5187), as an xor butterfly: fp32 addition is commutative,
// so every lane ends with the bit pattern lane 0 of the group gets in the
// reference.  Candidate `a` of a 32-chunk is evaluated by lane group a / LP at
// step a % LP, so lane a captures its own result without a final shuffle.
//
// One persistent CTA per SM: cbd (c1*c1*32 floats) is staged once, the per-query
// LUT (c1*32 floats) is double-buffered with TMA bulk copies (cp.async.bulk +
// mbarrier).
// ============================================================================
// ADC distances of the 32 candidates of one warp step (Step E2, :5277-5329).  `pos` = this
// lane's candidate (code row; 0 for idle lanes: a valid row, result unused).  Lane = g*LP + lp
// evaluates line part lp of the LP candidates of its lane group g, one after the other, and
// the LP partial distances of all of them are then summed with the reference's pairwise tree
// (warpReduceSum, :5183-5187: v += shfl_down(v, st), st = LP/2 .. 1) in one transposed
// butterfly: at distance st a lane keeps the half of its partial sums whose candidate index
// has bit st equal to its own lane bit and hands the other half to its partner, so every step
// adds exactly the two operands the reference adds (fp32 addition is commutative) and lane lp
// ends with candidate lp -- i.e. every lane returns the distance of its own candidate.
//   codes_lp: code array + lp;  lut_b / cbd_b: shared-space byte address of this lane's column
//   of the LUT (rows of 32 floats) / of the c^2 table (rows of CROW floats)
// ROWS: the code rows of the candidates may live in different allocations (shards mapped
// from peer GPUs): `row` = this lane's candidate's code row and the 64-bit pointer is
// shuffled instead of the position.
// Split in two so that a caller can issue the loads of the next step before it evaluates the
// current one (software pipelining, rerank_kernel):
//   adc_load_rows   the LP code words of this lane (line part lp of the LP candidates of its lane
//                   group), one coalesced row read per candidate
//   adc_eval_rows   table look-ups, dist(), summation butterfly -> distance of the lane's candidate
template <int LP, bool ROWS>
__device__ __forceinline__ void adc_load_rows(uint32_t (&w)[LP], uint32_t pos,
                                              const uint32_t* __restrict__ codes_lp, uint32_t lp,
                                              const uint32_t* row) {
#pragma unroll
  for (int s = 0; s < LP; s++) {
    if (ROWS) {
      const unsigned long long rp =
          __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)row, s, LP);
      w[s] = __ldg(reinterpret_cast<const uint32_t*>((uintptr_t)rp) + lp);
    } else {
      // position of candidate s of this lane group (shuffle inside the LP-lane segment)
      const uint32_t cpos = __shfl_sync(0xffffffffu, pos, s, LP);
      w[s] = __ldg(codes_lp + (size_t)cpos * LP);
    }
  }
}

// partial distance of one line code (Step E2, :5295-5310)
template <uint32_t CROW>
__device__ __forceinline__ float adc_code_dist(uint32_t w, uint32_t lut_b, uint32_t cbd_b, uint32_t c1,
                                               uint32_t sel) {
  // lineDescr {p1, p2, lambda} (pqt/PerturbationProTree.hh:21-25); one multiply-add per
  // table address: column base + row * row bytes
  const uint32_t p1 = w & 0xFFu;
  const uint32_t p2 = __byte_perm(w, 0u, 0x4441u);
  uint32_t aa, ab, ac, pr;
  asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(aa) : "r"(p1), "r"(lut_b));
  asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(ab) : "r"(p2), "r"(lut_b));
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(pr) : "r"(p2), "r"(c1), "r"(p1));
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ac) : "r"(pr), "n"(CROW * 4u), "r"(cbd_b));
  // toFloat (pqt/triangle.cuh:14-18): u16 dropped into the mantissa of 2^23, one FFMA
  const float lam = __fmaf_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, sel)), 1.220703125e-4f, -1028.f);
  const float a2 = lds_f32(aa);
  const float b2 = lds_f32(ab);
  const float c2 = lds_f32(ac);
  return tri_dist(a2, b2, c2, lam);
}

template <int LP, uint32_t CROW>
__device__ __forceinline__ float adc_eval_rows(const uint32_t (&w)[LP], uint32_t lut_b, uint32_t cbd_b,
                                               uint32_t c1, uint32_t lp) {
  // byte selector of the lambda conversion, pinned in a register for the whole step (the
  // compiler would otherwise re-create it next to every use)
  uint32_t sel;
  asm volatile("mov.u32 %0, 0x7632;" : "=r"(sel));
  if constexpr (LP == 1) {
    return adc_code_dist<CROW>(w[0], lut_b, cbd_b, c1, sel);
  } else {
    // The first butterfly stage (distance LP/2) is applied as soon as the two partial distances
    // it pairs exist, so only LP/2 sums stay live (same operand pairs as the reference's tree).
    constexpr int HF = LP / 2;
    float d[HF];
    const bool up0 = (lp & (uint32_t)HF) != 0u;
#pragma unroll
    for (int s = 0; s < HF; s++) {
      const float d0 = adc_code_dist<CROW>(w[s], lut_b, cbd_b, c1, sel);
      const float d1 = adc_code_dist<CROW>(w[s + HF], lut_b, cbd_b, c1, sel);
      const float send = up0 ? d0 : d1;
      const float keep = up0 ? d1 : d0;
      d[s] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, HF));
    }
#pragma unroll
    for (int st = HF >> 1; st > 0; st >>= 1) {
      const bool up = (lp & (uint32_t)st) != 0u;
#pragma unroll
      for (int s = 0; s < st; s++) {
        const float send = up ? d[s] : d[s + st];
        const float keep = up ? d[s + st] : d[s];
        d[s] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, st));
      }
    }
    return d[0];
  }
}

template <int LP, uint32_t CROW, bool ROWS = false>
__device__ __forceinline__ float adc_warp_step(uint32_t pos, const uint32_t* __restrict__ codes_lp,
                                               uint32_t lut_b, uint32_t cbd_b, uint32_t c1, uint32_t lp,
                                               const uint32_t* row = nullptr) {
  uint32_t w[LP];
  adc_load_rows<LP, ROWS>(w, pos, codes_lp, lp, row);
  return adc_eval_rows<LP, CROW>(w, lut_b, cbd_b, c1, lp);
}

struct ScanArgs {
  const uint32_t* codes;     // [n_local][LP] line codes in bin order (this shard's slice)
  const uint32_t* ids;       // [n_local] vector id of each position
  const uint32_t* cand_pos;  // [QN][max_vec] global positions
  const uint32_t* n_vec;     // [QN]
  const float* lut_dup;      // [QN][c1][32]
  const float* cbd_dup;      // [c1*c1][32]
  uint32_t QN, c1, max_vec;
  uint32_t pos_lo, pos_hi;  // this shard's slice of positions
  uint32_t owns_pad;        // 1: write (1e7, PAD) into slots >= nVec; 0: (+inf, 0)
  uint32_t sharded;         // 1: slots of other shards get (+inf, 0)
  float* out_val;           // [QN][max_vec]
  uint32_t* out_idx;        // [QN][max_vec]
  // peer-store mode (multi-GPU, fused scan + all-to-all): results of this shard's candidates
  // are stored straight into the candidate arrays of the rank that owns the query
  // (peer memory over NVLink); nothing is written for other shards' candidates or pads.
  uint32_t p2p;             // 1: use peer_val / peer_idx
  uint32_t q_per_rank;      // queries [r*q_per_rank, (r+1)*q_per_rank) are ranked by rank r
  float* peer_val[8];       // [world] each [q_per_rank][max_vec], mapped peer (or local) memory
  uint32_t* peer_idx[8];
};

constexpr int kScanThreads = 1024;

template <int LP>
__global__ void __launch_bounds__(kScanThreads, 1) adc_scan_kernel(ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t lut_floats = a.c1 * 32;
  const uint32_t cbd_floats = a.c1 * a.c1 * 32;
  float* s_cbd = reinterpret_cast<float*>(smem_raw);
  float* s_lut0 = s_cbd + cbd_floats;
  float* s_lut1 = s_lut0 + lut_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lut1 + lut_floats);  // [0],[1]: lut, [2]: cbd

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t lp = lane & (LP - 1);

  if (blockIdx.x >= a.QN) return;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // stage cbd once (chunks of <= 32 KB) and the first query's LUT
    const uint32_t cbd_bytes = cbd_floats * 4;
    mbar_expect_tx(&bars[2], cbd_bytes);
    for (uint32_t off = 0; off < cbd_bytes; off += 32768) {
      uint32_t n = cbd_bytes - off < 32768 ? cbd_bytes - off : 32768;
      tma_bulk_g2s(reinterpret_cast<unsigned char*>(s_cbd) + off,
                   reinterpret_cast<const unsigned char*>(a.cbd_dup) + off, n, &bars[2]);
    }
    mbar_expect_tx(&bars[0], lut_floats * 4);
    tma_bulk_g2s(s_lut0, a.lut_dup + (size_t)blockIdx.x * lut_floats, lut_floats * 4, &bars[0]);
  }
  mbar_wait(&bars[2], 0);

  uint32_t buf = 0, phase0 = 0, phase1 = 0;
  for (uint32_t qi = blockIdx.x; qi < a.QN; qi += gridDim.x) {
    // prefetch the next query's LUT into the other buffer (its readers passed the
    // __syncthreads at the end of the previous iteration)
    const uint32_t qn = qi + gridDim.x;
    if (threadIdx.x == 0 && qn < a.QN) {
      uint64_t* nb = &bars[buf ^ 1];
      mbar_expect_tx(nb, lut_floats * 4);
      tma_bulk_g2s(buf ? s_lut0 : s_lut1, a.lut_dup + (size_t)qn * lut_floats, lut_floats * 4, nb);
    }
    const float* s_lut = buf ? s_lut1 : s_lut0;
    mbar_wait(&bars[buf], buf ? phase1 : phase0);
    if (buf)
      phase1 ^= 1;
    else
      phase0 ^= 1;

    const uint32_t nv = min(__ldg(a.n_vec + qi), a.max_vec);
    const uint32_t* cand = a.cand_pos + (size_t)qi * a.max_vec;
    float* oval;
    uint32_t* oidx;
    if (a.p2p) {
      const uint32_t owner = qi / a.q_per_rank, ql = qi - owner * a.q_per_rank;
      oval = a.peer_val[owner] + (size_t)ql * a.max_vec;
      oidx = a.peer_idx[owner] + (size_t)ql * a.max_vec;
    } else {
      oval = a.out_val + (size_t)qi * a.max_vec;
      oidx = a.out_idx + (size_t)qi * a.max_vec;
    }
    const uint32_t scan_end = a.p2p ? nv : a.max_vec;  // peer-store mode: real candidates only

    for (uint32_t base = warp * 32; base < scan_end; base += nwarps * 32) {
      const uint32_t ca = base + lane;
      const bool valid = ca < nv;
      uint32_t pos = 0;
      bool mine = false;
      if (valid) {
        pos = __ldg(cand + ca);
        mine = (pos >= a.pos_lo) && (pos < a.pos_hi);
      }
      const uint32_t lpos = pos - a.pos_lo;
      float myval = 0.f;
      uint32_t myid = 0;
      const uint32_t any_mine = __ballot_sync(0xffffffffu, mine);
      if (any_mine) {
        if (mine) myid = __ldg(a.ids + lpos);
        myval = adc_warp_step<LP, 32u>(mine ? lpos : 0u, a.codes + lp, smem_u32(s_lut) + lane * 4u,
                                       smem_u32(s_cbd) + lane * 4u, a.c1, lp);
      }
      float v;
      uint32_t id;
      if (mine) {
        v = myval;
        id = myid;
      } else if (valid) {  // another shard's candidate
        v = __int_as_float(0x7f800000);
        id = kNotMineIdx;
      } else if (a.owns_pad) {
        v = kPadDist;
        id = kPadIdx;
      } else {
        v = __int_as_float(0x7f800000);
        id = kNotMineIdx;
      }
      if (a.p2p) {
        if (mine) {
          oval[ca] = v;
          oidx[ca] = id;
        }
      } else if (ca < a.max_vec) {  // candidate widths below 32 leave the upper lanes idle
        oval[ca] = v;
        oidx[ca] = id;
      }
    }
    __syncthreads();  // everyone is done with s_lut[buf] before it is refilled
    buf ^= 1;
  }
}

// debug: select_idx[a] = ids[cand_pos[a]] for a < nVec else 0 (:6183 memset)
__global__ void gather_select_idx_kernel(const uint32_t* cand_pos, const uint32_t* n_vec,
                                         const uint32_t* ids, uint32_t QN, uint32_t max_vec,
                                         uint32_t* out) {
  size_t total = (size_t)QN * max_vec;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    uint32_t qi = (uint32_t)(e / max_vec), ca = (uint32_t)(e - (size_t)qi * max_vec);
    out[e] = (ca < n_vec[qi]) ? ids[cand_pos[e]] : 0u;
  }
}

// ============================================================================
// tables_warp_kernel: Steps A+B+C, warp-per-(query, part) version for the common
// shapes (c1 <= 32, LP % p == 0, vl in {8,16,32,64}).  Same arithmetic as
// tables_kernel; what changes is the mapping:
//   * lanes run over centroids and read TRANSPOSED codebooks (cb1T[dim][c1],
//     cb2T[p][c1][vl][c2]) so that every load is one coalesced 128-byte request;
//   * the squared differences of a part are formed once in registers and feed both
//     the Step-A tree (over vl) and the Step-B trees (over dim/LP);
//   * the Step-A network runs on shuffles, the Step-C network in a warp-private
//     shared-memory array (no block barriers).
// ============================================================================
struct TablesWarpArgs {
  TablesArgs t;
  const float* cb1T;  // [dim][c1]
  const float* cb2T;  // [p][c1][vl][c2]
};

constexpr int kTablesWarps = 4;

// warp-private bitonic network over n (power of two >= 2) elements in shared memory
__device__ __forceinline__ void bitonic_warp_smem(float* val, uint32_t* idx, uint32_t n,
                                                  uint32_t lane) {
  const uint32_t half = n >> 1;
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = lane; t < half; t += 32) {
        uint32_t i = 2 * t - (t & (j - 1));
        uint32_t b = i + j;
        const float va = val[i], vb = val[b];
        const uint32_t ia = idx[i], ib = idx[b];
        const bool sw = ((i & k) == 0) ? (va > vb) : (va < vb);
        val[i] = sw ? vb : va;
        val[b] = sw ? va : vb;
        idx[i] = sw ? ib : ia;
        idx[b] = sw ? ia : ib;
      }
      __syncwarp();
    }
  }
}

// whole network over n = 32 * E elements of a warp-private shared-memory array, run in
// registers and shuffles (E consecutive elements per lane)
template <int E>
__device__ __forceinline__ void warp_sort_regs(float* sv, uint32_t* si, uint32_t lane) {
  float v[E];
  uint32_t p[E];
  const uint32_t base = lane * E;
#pragma unroll
  for (int r = 0; r < E; r++) {
    v[r] = sv[base + r];
    p[r] = si[base + r];
  }
  for (uint32_t k = 2; k <= 32u * E; k <<= 1) sort_reg_stages<E>(v, p, base, lane, k, k >> 1);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < E; r++) {
    sv[base + r] = v[r];
    si[base + r] = p[r];
  }
  __syncwarp();
}

// CC != 0: c1 == c2 == CC known at compile time (codebook strides become immediate load
// offsets, the entry -> (cell, centroid) split a shift); CC == 0: any c1, c2.
template <int VL, int CC>
__global__ void __launch_bounds__(kTablesWarps * 32) tables_warp_kernel(TablesWarpArgs w) {
  const TablesArgs& a = w.t;
  const uint32_t c1 = CC ? (uint32_t)CC : a.c1;
  const uint32_t c2 = CC ? (uint32_t)CC : a.c2;
  extern __shared__ float smem_f[];
  float* sq = smem_f;                                   // [dim]
  float* s_lut = sq + a.dim;                            // [c1*32]
  float* s_sortv = s_lut + c1 * 32;                   // [warps][npC]
  uint32_t* s_sorti = reinterpret_cast<uint32_t*>(s_sortv + kTablesWarps * a.npC);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t segs = a.LP / a.p;  // line segments inside one part
  const uint32_t R = 32 / a.LP;
  const uint32_t n = a.k1 * c2;

  for (uint32_t qi = blockIdx.x; qi < a.QN; qi += gridDim.x) {
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < a.dim; t += blockDim.x) sq[t] = a.Q[(size_t)qi * a.dim + t];
    __syncthreads();

    for (uint32_t part = warp; part < a.p; part += kTablesWarps) {
      float qreg[VL];
#pragma unroll
      for (int t = 0; t < VL; t++) qreg[t] = sq[part * VL + t];

      // ---- Steps A + B: lane = L1 centroid
      float va = kPadSortA;
      uint32_t ia = kPadIdx;
      if (lane < c1) {
        float s[VL], tb[VL];
        const float* cb = w.cb1T + (size_t)(part * VL) * c1 + lane;
#pragma unroll
        for (int t = 0; t < VL; t++) {
          float d = __fsub_rn(qreg[t], __ldg(cb + (uint32_t)t * c1));
          s[t] = __fmul_rn(d, d);
          tb[t] = s[t];
        }
        // Step B trees over sl = VL / segs elements (:7764-7776)
#pragma unroll
        for (int stride = VL / 2; stride > 0; stride >>= 1) {
          if ((uint32_t)stride < a.sl) {
#pragma unroll
            for (int j = 0; j + stride < VL; j++)
              if (((uint32_t)j & (a.sl - 1)) < (uint32_t)stride) tb[j] = __fadd_rn(tb[j], tb[j + stride]);
          }
        }
#pragma unroll
        for (int j = 0; j < VL; j++) {
          if (((uint32_t)j & (a.sl - 1)) == 0) {
            const uint32_t lp = part * segs + (uint32_t)j / a.sl;
            for (uint32_t r = 0; r < R; r++) s_lut[lane * 32 + r * a.LP + lp] = tb[j];
            if (a.dbg_lut) a.dbg_lut[((size_t)qi * a.LP + lp) * c1 + lane] = tb[j];
          }
        }
        // Step A tree over vl (:7151-7159)
#pragma unroll
        for (int stride = VL / 2; stride > 0; stride >>= 1) {
#pragma unroll
          for (int j = 0; j < stride; j++) s[j] = __fadd_rn(s[j], s[j + stride]);
        }
        va = s[0];
        ia = lane;
      }
      // bitonic network over npA <= 32 lanes (pqt/bitonicSort.cuh:16-44)
      for (uint32_t k = 2; k <= a.npA; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          float ov = __shfl_xor_sync(0xffffffffu, va, j);
          uint32_t oi = __shfl_xor_sync(0xffffffffu, ia, j);
          const bool lower = (lane & j) == 0;
          const float lo = lower ? va : ov, hi = lower ? ov : va;
          const bool sw = ((lane & k) == 0) ? (lo > hi) : (lo < hi);
          if (sw && lane < a.npA) {
            va = ov;
            ia = oi;
          }
        }
      }
      if (a.dbg_assign && lane < a.k1) a.dbg_assign[(size_t)qi * a.k1 * a.p + lane * a.p + part] = ia;

      // ---- Step C: entry e = k*c2 + l2, lanes over e (coalesced over l2)
      float* sv = s_sortv + warp * a.npC;
      uint32_t* si = s_sorti + warp * a.npC;
      for (uint32_t e0 = 0; e0 < a.npC; e0 += 32) {
        const uint32_t e = e0 + lane;
        float v = kPadSortC;
        uint32_t id = kPadIdx;
        // all lanes take part in the shuffle; k is warp-uniform only when c2 >= 32
        const uint32_t k = e < n ? e / c2 : 0, l2 = e < n ? e - k * c2 : 0;
        const uint32_t l1 = __shfl_sync(0xffffffffu, ia, k);
        if (e < n) {
          const float* cb = w.cb2T + ((size_t)(part * c1 + l1) * VL) * c2 + l2;
          float s[VL];
#pragma unroll
          for (int t = 0; t < VL; t++) {
            float d = __fsub_rn(qreg[t], __ldg(cb + (uint32_t)t * c2));
            s[t] = __fmul_rn(d, d);
          }
#pragma unroll
          for (int stride = VL / 2; stride > 0; stride >>= 1) {
#pragma unroll
            for (int j = 0; j < stride; j++) s[j] = __fadd_rn(s[j], s[j + stride]);
          }
          v = s[0];
          id = l2 + l1 * c2;
        }
        if (e < a.npC) {
          sv[e] = v;
          si[e] = id;
        }
      }
      __syncwarp();
      // Only the 16 best entries are consumed unless a caller wants the sorted table (the 1-B
      // variant, debug stages).  They are found with the composite-key sort of fast_rank.cuh:
      // 24-bit order-preserving key | entry number, unsigned min/max compare-exchanges.  If
      // two of the first 17 keys collide (equal or nearly equal distances: the network's own
      // order would decide) the network below runs instead.
      bool top16_done = false;
      if (a.npC == 256 && !a.top_val && !a.dbg_aval) {
        uint32_t c[8];
        uint32_t umin = 0xFFFFFFFFu, umax = 0u;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const uint32_t e = lane * 8 + r;
          c[r] = __float_as_uint(sv[e]);  // distances are sums of squares: >= +0, bits ordered
          if (e < n) {
            umin = min(umin, c[r]);
            umax = max(umax, c[r]);
          }
        }
        umin = __reduce_min_sync(0xffffffffu, umin);
        umax = __reduce_max_sync(0xffffffffu, umax);
        const uint32_t bits = 32u - (uint32_t)__clz((int)(umax - umin));
        const uint32_t shift = bits > 24u ? bits - 24u : 0u;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const uint32_t e = lane * 8 + r;
          c[r] = e < n ? ((((c[r] - umin) >> shift) << 8) | e) : 0xFFFFFFFFu;
        }
        block_sort_u32<8>(c, lane, 32u, 32u, nullptr, 0u);  // lane t: sorted slots 8t .. 8t+7
        // slots 0..16 must have pairwise different keys
        uint32_t clash = 0;
#pragma unroll
        for (int r = 0; r < 7; r++) clash |= (((c[r] ^ c[r + 1]) >> 8) == 0u) ? 1u : 0u;
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, c[0], 1);
        clash |= (((c[7] ^ nxt) >> 8) == 0u) ? 1u : 0u;
        if (!(__ballot_sync(0xffffffffu, clash != 0u) & 3u)) {
          if (lane < 2) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
              const uint32_t slot = lane * 8 + r;
              a.idx16[((size_t)qi * a.p + part) * 16 + slot] = (slot < a.m) ? si[c[r] & 0xFFu] : 0u;
            }
          }
          top16_done = true;
        }
      }
      // The 1-B variant (k1 = 16 cells: 512 entries) consumes the 64 best entries per part:
      // same selection with 23-bit keys | 9-bit entry numbers, slots 0..64 pairwise different.
      if (a.npC == 512 && a.top_val && a.top_n == 64 && !a.dbg_aval) {
        uint32_t c[16];
        uint32_t umin = 0xFFFFFFFFu, umax = 0u;
#pragma unroll
        for (int r = 0; r < 16; r++) {
          const uint32_t e = lane * 16 + r;
          c[r] = __float_as_uint(sv[e]);
          if (e < n) {
            umin = min(umin, c[r]);
            umax = max(umax, c[r]);
          }
        }
        umin = __reduce_min_sync(0xffffffffu, umin);
        umax = __reduce_max_sync(0xffffffffu, umax);
        const uint32_t bits = 32u - (uint32_t)__clz((int)(umax - umin));
        const uint32_t shift = bits > 23u ? bits - 23u : 0u;
#pragma unroll
        for (int r = 0; r < 16; r++) {
          const uint32_t e = lane * 16 + r;
          c[r] = e < n ? ((((c[r] - umin) >> shift) << 9) | e) : 0xFFFFFFFFu;
        }
        block_sort_u32<16>(c, lane, 32u, 32u, nullptr, 0u);  // lane t: sorted slots 16t .. 16t+15
        uint32_t clash = 0;
#pragma unroll
        for (int r = 0; r < 15; r++) clash |= (((c[r] ^ c[r + 1]) >> 9) == 0u) ? 1u : 0u;
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, c[0], 1);
        clash |= (((c[15] ^ nxt) >> 9) == 0u) ? 1u : 0u;
        if (!(__ballot_sync(0xffffffffu, clash != 0u) & 15u)) {  // lanes 0..3: boundaries of slots 0..64
          if (lane < 4) {
#pragma unroll
            for (int r = 0; r < 16; r++) {
              const uint32_t slot = lane * 16 + r, e = c[r] & 0x1FFu;
              const bool real = c[r] != 0xFFFFFFFFu;
              a.top_val[((size_t)qi * a.p + part) * a.top_n + slot] = real ? sv[e] : kPadSortC;
              a.top_idx[((size_t)qi * a.p + part) * a.top_n + slot] = real ? si[e] : kPadIdx;
              if (slot < 16) a.idx16[((size_t)qi * a.p + part) * 16 + slot] = (slot < a.m && real) ? si[e] : 0u;
            }
          }
          top16_done = true;
        }
      }
      if (!top16_done) {
        switch (a.npC) {  // :1639 bitonic3(shm, shmIdx, _NP2)
          case 32: warp_sort_regs<1>(sv, si, lane); break;
          case 64: warp_sort_regs<2>(sv, si, lane); break;
          case 128: warp_sort_regs<4>(sv, si, lane); break;
          case 256: warp_sort_regs<8>(sv, si, lane); break;
          case 512: warp_sort_regs<16>(sv, si, lane); break;  // k1 = 16 cells (the 1-B variant)
          default: bitonic_warp_smem(sv, si, a.npC, lane); break;
        }
        if (lane < 16) a.idx16[((size_t)qi * a.p + part) * 16 + lane] = (lane < a.m) ? si[lane] : 0u;
      }
      if (a.top_val && !top16_done) {
        for (uint32_t e = lane; e < a.top_n; e += 32) {
          a.top_val[((size_t)qi * a.p + part) * a.top_n + e] = sv[e];
          a.top_idx[((size_t)qi * a.p + part) * a.top_n + e] = si[e];
        }
      }
      if (a.dbg_aval) {
        for (uint32_t e = lane; e < n; e += 32) {
          a.dbg_aval[((size_t)qi * a.p + part) * n + e] = sv[e];
          a.dbg_aidx[((size_t)qi * a.p + part) * n + e] = si[e];
        }
      }
      __syncwarp();
    }
    __syncthreads();
    // LUT rows out, coalesced
    float4* dst = reinterpret_cast<float4*>(a.lut_dup + (size_t)qi * c1 * 32);
    const float4* src = reinterpret_cast<const float4*>(s_lut);
    for (uint32_t e = threadIdx.x; e < c1 * 8; e += blockDim.x) dst[e] = src[e];
  }
}

}  // namespace pqtb
