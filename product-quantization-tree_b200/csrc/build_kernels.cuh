// build_kernels.cuh -- index construction on the GPU (creates the query path's
// inputs): bin assignment of DB vectors, inverted lists, line encoding.
// Semantics: SURVEY.md App. B.2; citations are file:line into
// /root/reference/pqt/PerturbationProTree.cu.  Checked bit-for-bit against the
// oracle's builder (pqto_assign_bins / pqto_build_lists / pqto_line_encode).
#pragma once
#include "common.cuh"
#include "query_kernels.cuh"

namespace pqtb {

// ============================================================================
// buildKBestDB (:1231-1315): Step A with k1_build cells, then
// assignPerturbationBestBinKernel2 (:830-942): nearest L2 centroid among those
// cells (first strictly smallest in (k, l2) order), uint32 Horner, % HASH_SIZE.
// One CTA (128 threads) per DB vector; also accumulates the bin histogram
// (countBinsKernel :625-634).
// ============================================================================
struct AssignBinsArgs {
  const float* X;  // [N][dim]
  const float* cb1;
  const float* cb2;
  uint32_t N, dim, p, c1, c2, vl, k1, npA;
  FastMod hash;
  uint32_t* bin_of;  // [N]
  uint32_t* counts;  // [hash_size], pre-zeroed; null: no histogram
};

// dynamic smem: x[dim] | val[p*npA] | idx[p*npA] | assign[k1*p] | best[p]
__global__ void __launch_bounds__(128) assign_bins_kernel(AssignBinsArgs a) {
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sval = sx + a.dim;
  uint32_t* sidx = reinterpret_cast<uint32_t*>(sval + a.p * a.npA);
  uint32_t* sassign = sidx + a.p * a.npA;
  uint32_t* sbest = sassign + a.k1 * a.p;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

  for (uint32_t i = blockIdx.x; i < a.N; i += gridDim.x) {
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < a.dim; t += blockDim.x) sx[t] = a.X[(size_t)i * a.dim + t];
    for (uint32_t e = threadIdx.x; e < a.p * a.npA; e += blockDim.x) {
      sval[e] = kPadSortA;
      sidx[e] = kPadIdx;
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < a.p * a.c1; e += blockDim.x) {
      uint32_t part = e / a.c1, c = e - part * a.c1;
      sval[part * a.npA + c] =
          seg_dist_dyn(sx + part * a.vl, a.cb1 + (size_t)c * a.dim + part * a.vl, a.vl);
      sidx[part * a.npA + c] = c;
    }
    __syncthreads();
    bitonic_smem(sval, sidx, a.npA, a.p);
    for (uint32_t e = threadIdx.x; e < a.k1 * a.p; e += blockDim.x) {
      uint32_t k = e / a.p, part = e - k * a.p;
      sassign[e] = sidx[part * a.npA + k];
    }
    __syncthreads();

    // nearest L2 centroid over the k1 cells: min over (value, visiting order)
    const uint32_t n = a.k1 * a.c2;
    for (uint32_t part = warp; part < a.p; part += nwarps) {
      float bv = 0.f;
      uint32_t be = 0xFFFFFFFFu, bidx = 0;
      for (uint32_t e = lane; e < n; e += 32) {
        uint32_t k = e / a.c2, l2 = e - k * a.c2;
        uint32_t l1 = sassign[k * a.p + part];
        const float* cb = a.cb2 + ((size_t)(part * a.c1 + l1) * a.c2 + l2) * a.vl;
        float v = seg_dist_dyn(sx + part * a.vl, cb, a.vl);
        // "val > new || first" (:906-912): strictly smaller wins, earlier e on ties
        if (be == 0xFFFFFFFFu || v < bv) {
          bv = v;
          be = e;
          bidx = l2 + l1 * a.c2;
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, d);
        uint32_t oe = __shfl_xor_sync(0xffffffffu, be, d);
        uint32_t oi = __shfl_xor_sync(0xffffffffu, bidx, d);
        bool take = (oe != 0xFFFFFFFFu) && (be == 0xFFFFFFFFu || ov < bv || (ov == bv && oe < be));
        if (take) {
          bv = ov;
          be = oe;
          bidx = oi;
        }
      }
      if (lane == 0) sbest[part] = bidx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t o = sbest[0];
      for (uint32_t part = 1; part < a.p; part++) o = o * a.c1 * a.c2 + sbest[part];  // :929-931
      uint32_t bin = fastmod(o, a.hash);
      a.bin_of[i] = bin;
      if (a.counts) atomicAdd(a.counts + bin, 1u);
    }
  }
}

// sortIdxKernel (:715-727): slot = prefix[bin] + (remaining count - 1).  Uses the
// histogram itself as the cursor (it is consumed: all zeros afterwards).  The
// order inside a bin is made deterministic by sort_within_bins_kernel.
__global__ void bin_slot_kernel(const uint32_t* bin_of, uint32_t N, uint32_t* counts,
                                const uint32_t* prefix, uint32_t* db_idx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (size_t)gridDim.x * blockDim.x) {
    uint32_t bin = bin_of[i];
    uint32_t slot = prefix[bin] + (atomicSub(counts + bin, 1u) - 1u);
    db_idx[slot] = (uint32_t)i;
  }
}

// ascending id inside every bin (the reference's atomicInc order is run-dependent,
// SURVEY.md App. C).  One thread per non-empty bin: insertion sort for short
// lists, heap sort for long ones.
__global__ void sort_within_bins_kernel(BinDir d, uint32_t n_nonempty, uint32_t* db_idx) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_nonempty;
       r += (size_t)gridDim.x * blockDim.x) {
    uint32_t lo = d.cprefix[r], hi = d.cprefix[r + 1];
    uint32_t n = hi - lo;
    uint32_t* v = db_idx + lo;
    if (n < 2) continue;
    if (n <= 48) {
      for (uint32_t i = 1; i < n; i++) {
        uint32_t x = v[i];
        uint32_t j = i;
        while (j > 0 && v[j - 1] > x) {
          v[j] = v[j - 1];
          j--;
        }
        v[j] = x;
      }
    } else {
      // heap sort
      for (uint32_t start = n / 2; start-- > 0;) {
        uint32_t root = start;
        for (;;) {
          uint32_t child = 2 * root + 1;
          if (child >= n) break;
          if (child + 1 < n && v[child] < v[child + 1]) child++;
          if (v[root] >= v[child]) break;
          uint32_t t = v[root];
          v[root] = v[child];
          v[child] = t;
          root = child;
        }
      }
      for (uint32_t end = n - 1; end > 0; end--) {
        uint32_t t = v[0];
        v[0] = v[end];
        v[end] = t;
        uint32_t root = 0;
        for (;;) {
          uint32_t child = 2 * root + 1;
          if (child >= end) break;
          if (child + 1 < end && v[child] < v[child + 1]) child++;
          if (v[root] >= v[child]) break;
          uint32_t t2 = v[root];
          v[root] = v[child];
          v[child] = t2;
          root = child;
        }
      }
    }
  }
}

// dense binCounts / binPrefix back from the directory (pqt_get_db)
__global__ void expand_directory_kernel(BinDir d, uint32_t hash_size, uint32_t* counts,
                                        uint32_t* prefix) {
  for (size_t bin = (size_t)blockIdx.x * blockDim.x + threadIdx.x; bin < hash_size;
       bin += (size_t)gridDim.x * blockDim.x) {
    const size_t w = bin >> 5, g = bin >> 8;
    uint32_t word = d.bitmap[w];
    uint32_t r = d.rank_base[g];
    for (size_t ww = g << 3; ww < w; ww++) r += __popc(d.bitmap[ww]);
    r += __popc(word & ((1u << (bin & 31)) - 1u));
    uint32_t start = d.cprefix[r];
    bool occ = (word >> (bin & 31)) & 1u;
    prefix[bin] = start;
    counts[bin] = occ ? d.cprefix[r + 1] - start : 0u;
  }
}

// ============================================================================
// lineDist (:7663-7737) / lineClusterKernelFast (:7527-7661): the line encoder.
// One CTA of LP*c1 threads per DB vector; thread (lp, cIdx) walks minId over all
// centroids.  Vectors are visited in bin order and the code row is written at
// the vector's bin-order position (the layout the ADC scan streams).
// ============================================================================
struct LineEncodeArgs {
  const float* X;       // [N][dim], indexed by vector id
  const float* cb1;     // [c1][dim]
  const float* cbd;     // [c1][c1][LP] canonical
  const uint32_t* ids;  // [N] vector id at each bin-order position; null: row r is vector r
                        // (chunked build: codes = the chunk's staging buffer in id order)
  uint32_t N, dim, c1, LP, sl;
  uint32_t* codes;  // [N][LP] in bin order
};

// dynamic smem: x[dim] | val[LP*c1] | dist[LP*c1] | code[LP*c1]
__global__ void line_encode_kernel(LineEncodeArgs a) {
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sval = sx + a.dim;
  float* sdist = sval + a.LP * a.c1;
  uint32_t* scode = reinterpret_cast<uint32_t*>(sdist + a.LP * a.c1);
  const uint32_t t = threadIdx.x;
  const uint32_t lp = t / a.c1, c = t - lp * a.c1;

  for (uint32_t pos = blockIdx.x; pos < a.N; pos += gridDim.x) {
    const uint32_t id = a.ids ? a.ids[pos] : pos;
    __syncthreads();
    for (uint32_t e = t; e < a.dim; e += blockDim.x) sx[e] = a.X[(size_t)id * a.dim + e];
    __syncthreads();
    sval[t] = seg_dist_dyn(sx + lp * a.sl, a.cb1 + (size_t)c * a.dim + lp * a.sl, a.sl);
    __syncthreads();
    float best = 0.f;
    uint32_t code = 0;
    const float va = sval[t];
    for (uint32_t mn = 0; mn < a.c1; mn++) {
      float cc = __ldg(a.cbd + ((size_t)mn * a.c1 + c) * a.LP + lp);
      float d;
      float l = tri_project(va, sval[lp * a.c1 + mn], cc, d);
      if (c == mn) d = 999999999999.f;
      if (mn == 0 || d < best) {
        best = d;
        code = (c & 0xFFu) | ((mn & 0xFFu) << 8) | (to_ushort(l) << 16);
      }
    }
    sdist[t] = best;
    scode[t] = code;
    // tree over centroids keeping the lower index unless strictly larger (:7633-7641)
    for (uint32_t stride = a.c1 >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      if (c < stride) {
        if (sdist[t] > sdist[t + stride]) {
          sdist[t] = sdist[t + stride];
          scode[t] = scode[t + stride];
        }
      }
    }
    __syncthreads();
    if (t < a.LP) a.codes[(size_t)pos * a.LP + t] = scode[t * a.c1];
  }
}

}  // namespace pqtb
