// kmeans.hpp -- small deterministic host-side Lloyd k-means used by tool_createdb to
// obtain the two codebooks.  The reference trains with split-and-Lloyd on the GPU
// (createTree, pqt/ProTree.cu:457-510; pqt/ProQuantization.cu:1047-1169); codebooks are
// INPUTS to the query path (shared through the .ppqt file), so training only has to be
// deterministic, not bit-identical to the reference.  Offline step, not on the hot path.
#ifndef PQT_B200_HOST_KMEANS_HPP
#define PQT_B200_HOST_KMEANS_HPP

#include <cstdint>
#include <limits>
#include <vector>

namespace pqt_train {

// x: n rows of d floats (row stride ld).  Returns k centroids (k*d floats).
inline std::vector<float> lloyd(const float* x, size_t n, size_t d, size_t ld, size_t k, int iters) {
  std::vector<float> cent(k * d, 0.f);
  if (n == 0) return cent;
  for (size_t c = 0; c < k; c++) {  // evenly spaced seeds; duplicates get a tiny offset
    const float* src = x + ((c * n) / k) * ld;
    for (size_t j = 0; j < d; j++) cent[c * d + j] = src[j] + (n < k ? 0.01f * (float)(c + 1) : 0.f);
  }
  std::vector<uint32_t> assign(n);
  std::vector<double> sum(k * d);
  std::vector<size_t> cnt(k);
  for (int it = 0; it < iters; it++) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; i++) {
      float best = std::numeric_limits<float>::max();
      uint32_t bi = 0;
      for (size_t c = 0; c < k; c++) {
        float s = 0.f;
        for (size_t j = 0; j < d; j++) {
          float t = x[i * ld + j] - cent[c * d + j];
          s += t * t;
        }
        if (s < best) {
          best = s;
          bi = (uint32_t)c;
        }
      }
      assign[i] = bi;
    }
    std::fill(sum.begin(), sum.end(), 0.0);
    std::fill(cnt.begin(), cnt.end(), 0);
    for (size_t i = 0; i < n; i++) {
      cnt[assign[i]]++;
      for (size_t j = 0; j < d; j++) sum[assign[i] * d + j] += x[i * ld + j];
    }
    for (size_t c = 0; c < k; c++) {
      if (cnt[c]) {
        for (size_t j = 0; j < d; j++) cent[c * d + j] = (float)(sum[c * d + j] / (double)cnt[c]);
      } else {  // re-seed an empty cluster next to a populated one
        const float* src = x + ((c * 7919 + (size_t)it * 104729) % n) * ld;
        for (size_t j = 0; j < d; j++) cent[c * d + j] = src[j] + 0.01f * (float)(c + 1);
      }
    }
  }
  return cent;
}

// Two-level tree in the reference's layouts: cb1[c1][dim] (part j = columns j*vl..),
// cb2[p][c1][c2][vl] = absolute level-2 centroids of the vectors of each L1 cell.
inline void train_tree(const float* x, size_t n, uint32_t dim, uint32_t p, uint32_t c1, uint32_t c2,
                       std::vector<float>& cb1, std::vector<float>& cb2, int iters = 10) {
  const uint32_t vl = dim / p;
  cb1.assign((size_t)c1 * dim, 0.f);
  cb2.assign((size_t)p * c1 * c2 * vl, 0.f);
  std::vector<float> cell;
  for (uint32_t part = 0; part < p; part++) {
    std::vector<float> cent = lloyd(x + part * vl, n, vl, dim, c1, iters);
    for (uint32_t c = 0; c < c1; c++)
      for (uint32_t j = 0; j < vl; j++) cb1[(size_t)c * dim + part * vl + j] = cent[c * vl + j];
    std::vector<uint32_t> a(n);
    for (size_t i = 0; i < n; i++) {
      float best = std::numeric_limits<float>::max();
      for (uint32_t c = 0; c < c1; c++) {
        float s = 0.f;
        for (uint32_t j = 0; j < vl; j++) {
          float t = x[i * dim + part * vl + j] - cent[c * vl + j];
          s += t * t;
        }
        if (s < best) {
          best = s;
          a[i] = c;
        }
      }
    }
    for (uint32_t c = 0; c < c1; c++) {
      cell.clear();
      for (size_t i = 0; i < n; i++)
        if (a[i] == c) cell.insert(cell.end(), x + i * dim + part * vl, x + i * dim + (part + 1) * vl);
      size_t m = cell.size() / vl;
      if (m == 0) {
        cell.assign(cent.begin() + c * vl, cent.begin() + (c + 1) * vl);
        m = 1;
      }
      std::vector<float> c2c = lloyd(cell.data(), m, vl, vl, c2, iters);
      for (size_t e = 0; e < (size_t)c2 * vl; e++) cb2[((size_t)part * c1 + c) * c2 * vl + e] = c2c[e];
    }
  }
}

}  // namespace pqt_train
#endif
