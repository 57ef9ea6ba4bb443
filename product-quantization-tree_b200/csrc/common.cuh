// common.cuh -- shared device/host helpers for libpqt_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pqtb {

constexpr uint32_t kNumDistSeq = 65536;  // pqt/ProTree.hh:9 NUM_DISTSEQ
constexpr uint32_t kPadIdx = 0xFFFFFFFFu;
// id written by a shard into candidate slots it does not own: INT32_MIN, so that an
// element-wise signed MAX across shards keeps the owner's id (ids < 2^31) or PAD (-1)
constexpr uint32_t kNotMineIdx = 0x80000000u;
constexpr float kPadDist = 10000000.f;   // pqt/PerturbationProTree.cu:5333
constexpr float kPadSortA = 10000000.f;  // :7185
constexpr float kPadSortC = 1000000000.f;  // :1627

// pqt/helper.hh:27-37 ("log2": next power of two)
__host__ __device__ inline uint32_t pow2ceil(uint32_t x) {
  uint32_t y;
  for (y = 0; y < 32; y++)
    if (!((x - 1) >> y)) break;
  return 1u << y;
}

// ---- exact fp32 arithmetic of the reference kernels -------------------------
// Every operation is pinned with an explicit round-to-nearest intrinsic so that
// nvcc can neither fuse nor reorder: the results are bit-identical to the
// oracle (oracle/pqt_oracle.c), which follows the reference kernels' order.

// pqt/PerturbationProTree.cu:7146-7160: s[t] = sqr(q - c); pairwise tree
// s[j] += s[j + stride], stride = L/2 .. 1.
template <int L>
__device__ __forceinline__ float seg_dist(const float* __restrict__ q,
                                          const float* __restrict__ c) {
  float s[L];
#pragma unroll
  for (int t = 0; t < L; t++) {
    float d = __fsub_rn(q[t], c[t]);
    s[t] = __fmul_rn(d, d);
  }
#pragma unroll
  for (int stride = L / 2; stride > 0; stride >>= 1) {
#pragma unroll
    for (int j = 0; j < stride; j++) s[j] = __fadd_rn(s[j], s[j + stride]);
  }
  return s[0];
}

__device__ __forceinline__ float seg_dist_dyn(const float* __restrict__ q,
                                              const float* __restrict__ c, uint32_t len) {
  switch (len) {
    case 1: return seg_dist<1>(q, c);
    case 2: return seg_dist<2>(q, c);
    case 4: return seg_dist<4>(q, c);
    case 8: return seg_dist<8>(q, c);
    case 16: return seg_dist<16>(q, c);
    case 32: return seg_dist<32>(q, c);
    case 64: return seg_dist<64>(q, c);
    default: return seg_dist<128>(q, c);
  }
}

// pqt/triangle.cuh:14-18 toFloat: lambda = u16 * 2^-13 - 4 (exact).  The u16 is
// dropped into the mantissa of 2^23 (0x4B000000) so that one FFMA finishes the
// conversion: (2^23 + u) * 2^-13 - 1028 == u * 2^-13 - 4 exactly.
__device__ __forceinline__ float lambda_of(uint32_t w) {
  uint32_t bits = __byte_perm(w, 0x4B000000u, 0x7632);  // bytes: w.b2, w.b3, 0x00, 0x4B
  return __fmaf_rn(__uint_as_float(bits), 1.220703125e-4f, -1028.f);
}

// pqt/triangle.cuh:55-63 dist() in the form nvcc contracts it to in device code
// (see DESIGN.md "FMA pinning"): l2 = l*l; t = fma(c2, l2, b2); u = (a2-b2)-c2;
// d = fma(u, l, t).
__device__ __forceinline__ float tri_dist(float a2, float b2, float c2, float l) {
  float l2 = __fmul_rn(l, l);
  float t = __fmaf_rn(c2, l2, b2);
  float u = __fsub_rn(__fsub_rn(a2, b2), c2);
  return __fmaf_rn(u, l, t);
}

// pqt/triangle.cuh:102-110 project() with the d2 output, in the form the reference's
// device code takes: lambda = (-0.5 * ((a2 - b2) - c2)) / c2 (IEEE division) and
// d2 = fma(-c2, lambda*lambda, b2) (ptxas fuses the trailing mul + sub into one FFMA)
__device__ __forceinline__ float tri_project(float a2, float b2, float c2, float& d2) {
  float u = __fsub_rn(__fsub_rn(a2, b2), c2);
  float l = __fdiv_rn(__fmul_rn(-0.5f, u), c2);
  d2 = __fmaf_rn(-c2, __fmul_rn(l, l), b2);
  return l;
}

// pqt/triangle.cuh:6-12 toUShort (cvt.rzi semantics: NaN -> 0, truncation)
__device__ __forceinline__ uint32_t to_ushort(float f) {
  float ftrans = __fmul_rn(__fadd_rn(f, 4.f), 8192.f);
  float sel = (f >= 4.f) ? 65535.f : ((f < -4.f) ? 0.f : ftrans);
  return __float2uint_rz(sel) & 0xFFFFu;
}

// ---- exact x % d for a fixed 32-bit divisor (Lemire fastmod) -------------------
struct FastMod {
  uint64_t M;
  uint32_t d;
};
inline FastMod make_fastmod(uint32_t d) {
  FastMod f;
  f.d = d;
  f.M = d > 1 ? (0xFFFFFFFFFFFFFFFFull / d + 1) : 0;
  return f;
}
__device__ __forceinline__ uint32_t fastmod(uint32_t a, const FastMod& f) {
  if (f.d <= 1) return 0;
  uint64_t low = f.M * (uint64_t)a;
  return (uint32_t)__umul64hi(low, (uint64_t)f.d);
}

// ---- register / shuffle stages of the bitonic network ----------------------------------
// A thread holds E consecutive elements (global indices base .. base+E-1, base = E * thread
// index inside a 32-thread warp block).  Runs the compare-exchange distances jstart, jstart/2,
// ..., 1 of merge size k: E <= j on shuffles, j < E in registers.  Same pairs, direction rule
// and strict compares as pqt/bitonicSort.cuh:16-44.
template <int E>
__device__ __forceinline__ void sort_reg_stages(float (&v)[E], uint32_t (&p)[E], uint32_t base,
                                                uint32_t lane, uint32_t k, uint32_t jstart) {
  uint32_t j = jstart;
  // ascending block for all elements of this thread once k >= E (base & k ignores r)
  const bool desc_t = (base & k) != 0;
  // distances inside a warp: shuffles.  The lane holding the lower index of a pair keeps
  // the minimum when ascending; `nv != v` is exactly "the pair swaps" (ties keep their
  // own bits and payload, like the network's strict compare).
  for (; j >= (uint32_t)E; j >>= 1) {
    const uint32_t lj = j / E;  // lane distance
    const bool want_min = ((lane & lj) == 0) != desc_t;
#pragma unroll
    for (int r = 0; r < E; r++) {
      const float ov = __shfl_xor_sync(0xffffffffu, v[r], lj);
      const uint32_t op = __shfl_xor_sync(0xffffffffu, p[r], lj);
      const float nv = want_min ? fminf(v[r], ov) : fmaxf(v[r], ov);
      const bool sw = nv != v[r];
      v[r] = sw ? ov : v[r];
      p[r] = sw ? op : p[r];
    }
  }
  // distances inside a thread: registers
#pragma unroll
  for (int jj = E >> 1; jj > 0; jj >>= 1) {
    if ((uint32_t)jj <= j) {
#pragma unroll
      for (int r = 0; r < E; r++) {
        if ((r & jj) == 0) {
          const bool desc = ((base + r) & k) != 0;
          const float lo = v[r], hi = v[r + jj];
          const uint32_t plo = p[r], phi = p[r + jj];
          const bool sw = desc ? (lo < hi) : (lo > hi);
          v[r] = sw ? hi : lo;
          v[r + jj] = sw ? lo : hi;
          p[r] = sw ? phi : plo;
          p[r + jj] = sw ? plo : phi;
        }
      }
    }
  }
}

// ---- thread group: a contiguous set of warps of a CTA that works on one query and
// synchronises on its own named barrier (bar.sync id, n)
struct Grp {
  uint32_t t;    // thread index inside the group
  uint32_t n;    // threads in the group (multiple of 32)
  uint32_t bar;  // hardware barrier id (0 = the CTA-wide barrier of __syncthreads)
  __device__ __forceinline__ void sync() const {
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n) : "memory");
  }
};

// request the line of a global address into L2 (no register result, no stall)
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- mbarrier / TMA bulk copy (cp.async.bulk, SASS: UBLKCP) ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// shared-memory load through a 32-bit shared-space address (keeps table look-ups at one
// address instruction: base + (row << shift))
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier.  bytes % 16 == 0,
// both addresses 16-byte aligned.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace pqtb
