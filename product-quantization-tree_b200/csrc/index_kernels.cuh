// index_kernels.cuh -- load-time kernels: bin directory (bitmap + rank + compact
// prefix), bin-ordered re-layout of the line codes, cbDist, and a device-wide
// exclusive scan.  None of these is on the per-query path.
#pragma once
#include "common.cuh"

namespace pqtb {

// ---- device-wide exclusive scan of uint32 (ProTree::scan, pqt/ProTree.cu:1250-1299)
constexpr int kScanBlock = 512;
constexpr int kScanPerThread = 4;
constexpr int kScanTile = kScanBlock * kScanPerThread;

__device__ __forceinline__ uint32_t block_exscan512(uint32_t v, uint32_t* warp_sums,
                                                    uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  uint32_t ws = (lane < (kScanBlock >> 5)) ? warp_sums[lane] : 0;
  uint32_t winc = ws;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
    if (lane >= d) winc += t;
  }
  total = __shfl_sync(0xffffffffu, winc, 31);
  uint32_t wbase = __shfl_sync(0xffffffffu, winc - ws, warp);
  __syncthreads();
  return wbase + inc - v;
}

__global__ void __launch_bounds__(kScanBlock) scan_tile_sums_kernel(const uint32_t* in, size_t n,
                                                                    uint32_t* tile_sums) {
  __shared__ uint32_t warp_sums[32];
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanPerThread;
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < kScanPerThread; r++)
    if (base + r < n) s += in[base + r];
  uint32_t total;
  block_exscan512(s, warp_sums, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const uint32_t* in, size_t n,
                                                                const uint32_t* tile_offsets,
                                                                uint32_t* out) {
  __shared__ uint32_t warp_sums[32];
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanPerThread;
  uint32_t v[kScanPerThread];
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < kScanPerThread; r++) {
    v[r] = (base + r < n) ? in[base + r] : 0;
    s += v[r];
  }
  uint32_t total;
  uint32_t ex = block_exscan512(s, warp_sums, total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int r = 0; r < kScanPerThread; r++) {
    if (base + r < n) out[base + r] = ex;
    ex += v[r];
  }
}

// single-CTA scan for the top level (n <= kScanTile)
__global__ void __launch_bounds__(kScanBlock) scan_small_kernel(const uint32_t* in, size_t n,
                                                                uint32_t* out) {
  __shared__ uint32_t warp_sums[32];
  size_t base = (size_t)threadIdx.x * kScanPerThread;
  uint32_t v[kScanPerThread];
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < kScanPerThread; r++) {
    v[r] = (base + r < n) ? in[base + r] : 0;
    s += v[r];
  }
  uint32_t total;
  uint32_t ex = block_exscan512(s, warp_sums, total);
#pragma unroll
  for (int r = 0; r < kScanPerThread; r++) {
    if (base + r < n) out[base + r] = ex;
    ex += v[r];
  }
}

// out may alias in.  tmp must hold scan_tmp_words(n) uint32.
inline size_t scan_tmp_words(size_t n) {
  size_t words = 0;
  while (n > (size_t)kScanTile) {
    n = (n + kScanTile - 1) / kScanTile;
    words += n;
  }
  return words + 1;
}
inline void device_exscan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp,
                              cudaStream_t st) {
  if (n == 0) return;
  if (n <= (size_t)kScanTile) {
    scan_small_kernel<<<1, kScanBlock, 0, st>>>(in, n, out);
    return;
  }
  size_t tiles = (n + kScanTile - 1) / kScanTile;
  scan_tile_sums_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(in, n, tmp);
  device_exscan_u32(tmp, tmp, tiles, tmp + tiles, st);
  scan_apply_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(in, n, tmp, out);
}

// ---- bin directory ---------------------------------------------------------------
// bitmap word w covers bins [32w, 32w+32); one thread per bin, one ballot per word
__global__ void bitmap_build_kernel(const uint32_t* counts, uint32_t hash_size, uint32_t* bitmap) {
  const size_t nwords = ((size_t)hash_size + 31) >> 5;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (size_t w = warp0; w < nwords; w += nwarps) {
    size_t bin = (w << 5) + lane;
    bool occ = (bin < hash_size) && (counts[bin] != 0);
    uint32_t word = __ballot_sync(0xffffffffu, occ);
    if (lane == 0) bitmap[w] = word;
  }
}

// group g = 8 words = 256 bins = one 32-byte sector
__global__ void group_popc_kernel(const uint32_t* bitmap, size_t nwords, size_t ngroups,
                                  uint32_t* group_cnt) {
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups;
       g += (size_t)gridDim.x * blockDim.x) {
    uint32_t c = 0;
    for (int r = 0; r < 8; r++) {
      size_t w = g * 8 + r;
      if (w < nwords) c += __popc(bitmap[w]);
    }
    group_cnt[g] = c;
  }
}

// cprefix[rank(bin)] = prefix[bin] for every non-empty bin
__global__ void cprefix_fill_kernel(const uint32_t* counts, const uint32_t* prefix,
                                    uint32_t hash_size, const uint32_t* bitmap,
                                    const uint32_t* rank_base, uint32_t* cprefix) {
  for (size_t bin = (size_t)blockIdx.x * blockDim.x + threadIdx.x; bin < hash_size;
       bin += (size_t)gridDim.x * blockDim.x) {
    if (counts[bin] == 0) continue;
    const size_t w = bin >> 5, g = bin >> 8;
    uint32_t r = rank_base[g];
    for (size_t ww = g << 3; ww < w; ww++) r += __popc(bitmap[ww]);
    r += __popc(bitmap[w] & ((1u << (bin & 31)) - 1u));
    cprefix[r] = prefix[bin];
  }
}

// ---- input validation (load path) -------------------------------------------------------
// flag = 1 if any v[i] >= bound
__global__ void check_below_kernel(const uint32_t* v, size_t n, uint32_t bound, uint32_t* flag) {
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    bad |= v[i] >= bound;
  if (bad) atomicOr(flag, 1u);
}
// flag = 1 unless v[0] <= v[1] <= ... <= v[n-1]
__global__ void check_monotone_kernel(const uint32_t* v, size_t n, uint32_t* flag) {
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += (size_t)gridDim.x * blockDim.x)
    bad |= v[i] > v[i + 1];
  if (bad) atomicOr(flag, 1u);
}
// flag = 1 if a non-empty bin's list [prefix, prefix + count) leaves the N ids
__global__ void check_lists_kernel(const uint32_t* counts, const uint32_t* prefix, size_t hs, uint32_t N,
                                   uint32_t* flag) {
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hs; i += (size_t)gridDim.x * blockDim.x)
    bad |= counts[i] != 0 && (uint64_t)prefix[i] + counts[i] > N;
  if (bad) atomicOr(flag, 1u);
}
// flag = 1 if a lineDescr holds p1 or p2 >= c1
__global__ void check_codes_kernel(const uint32_t* w, size_t n, uint32_t c1, uint32_t* flag) {
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    bad |= ((w[i] & 0xFFu) >= c1) || (((w[i] >> 8) & 0xFFu) >= c1);
  if (bad) atomicOr(flag, 1u);
}

// ---- bin-ordered re-layout of line codes -------------------------------------------
// inv[id] = position of vector id in dbIdx
__global__ void invert_perm_kernel(const uint32_t* db_idx, uint32_t N, uint32_t* inv) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (size_t)gridDim.x * blockDim.x)
    inv[db_idx[i]] = (uint32_t)i;
}

// codes[(inv[id] - pos_lo)][lp] = lines_chunk[(id - id0)][lp] for ids of this chunk whose
// position falls into [pos_lo, pos_hi)
__global__ void scatter_codes_kernel(const uint32_t* lines_chunk, uint32_t id0, uint32_t n_chunk,
                                     uint32_t LP, const uint32_t* inv, uint32_t pos_lo,
                                     uint32_t pos_hi, uint32_t* codes) {
  const size_t total = (size_t)n_chunk * LP;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    uint32_t i = (uint32_t)(e / LP), lp = (uint32_t)(e - (size_t)i * LP);
    uint32_t pos = inv[id0 + i];
    if (pos >= pos_lo && pos < pos_hi) codes[(size_t)(pos - pos_lo) * LP + lp] = lines_chunk[e];
  }
}

// inverse: lines[id][lp] = codes[pos][lp] (pqt_get_lines)
__global__ void gather_codes_kernel(const uint32_t* codes, const uint32_t* ids, uint32_t n_local,
                                    uint32_t LP, uint32_t* lines) {
  const size_t total = (size_t)n_local * LP;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    uint32_t i = (uint32_t)(e / LP), lp = (uint32_t)(e - (size_t)i * LP);
    lines[(size_t)ids[i] * LP + lp] = codes[e];
  }
}

// ---- cbDist (computeCBL1L1Dist :1902-1917 + calcDistKernel pqt/ProQuantization.cu:101-137)
// canonical cbd[(b*c1 + a)*LP + lp]; dup layout cbd_dup[(b*c1 + a)*32 + j*LP + lp], j < 32/LP
__global__ void cb_dist_kernel(const float* cb1, uint32_t c1, uint32_t dim, uint32_t LP,
                               uint32_t sl, float* cbd, float* cbd_dup) {
  const uint32_t total = c1 * c1 * LP;
  const uint32_t R = LP <= 32 ? 32 / LP : 0;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += gridDim.x * blockDim.x) {
    uint32_t lp = e % LP, ba = e / LP;
    uint32_t b = ba / c1, a = ba - b * c1;
    float v = seg_dist_dyn(cb1 + (size_t)b * dim + lp * sl, cb1 + (size_t)a * dim + lp * sl, sl);
    cbd[e] = v;
    if (cbd_dup)
      for (uint32_t j = 0; j < R; j++) cbd_dup[(size_t)ba * 32 + j * LP + lp] = v;
  }
}

}  // namespace pqtb
