/*
 * pqt_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see pqt_oracle.h).
 *
 * Restates, step by step, what the reference's CUDA kernels compute for
 * PerturbationProTree::queryKNN, including rounding order: every float
 * operation below is a single IEEE-754 binary32 operation in the order the
 * reference's kernels perform it.  Where nvcc (default --fmad=true) contracts
 * a multiply-add in the reference's source into an FMA, the contracted form is
 * written out with fmaf() -- verified against the PTX nvcc 12.9 emits for
 * pqt/triangle.cuh (see DESIGN.md "FMA pinning").
 *
 * Build with -ffp-contract=off so that gcc adds no contractions of its own.
 * All file:line citations are into /root/reference.
 */
#include "pqt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* helper.hh / triangle.cuh                                                  */
/* ------------------------------------------------------------------------- */

void pqto_default_params(pqto_params *prm, uint32_t dim, uint32_t p, uint32_t c1, uint32_t c2,
                         uint32_t line_parts) {
  prm->dim = dim;
  prm->p = p;
  prm->c1 = c1;
  prm->c2 = c2;
  prm->line_parts = line_parts;
  prm->k1 = 8;                 /* pqt/PerturbationProTree.cu:8187 */
  prm->max_bins = 4096;        /* :8218 */
  prm->max_trials = 16;        /* :3569 */
  prm->bin_threads = 1024;     /* :3556 */
  prm->max_vec_per_bin = 2800; /* :6208 */
  prm->hash_size = 400000000u; /* pqt/PerturbationProTree.hh:12 */
}

/* pqt/helper.hh:27-37: despite its name ("log2") this returns the next power of two */
uint32_t pqto_pow2ceil(uint32_t x) {
  uint32_t y;
  for (y = 0; y < 32; y++)
    if (!((x - 1) >> y)) break;
  return 1u << y;
}

/* pqt/triangle.cuh:6-12.  Device semantics of the float->ushort conversion
 * (cvt.rzi.u32.f32 then truncation to 16 bit; NaN -> 0). */
uint16_t pqto_to_ushort(float f) {
  float ftrans = (f + 4.f) * (65536.f / 8.f);
  float sel = (f >= 4.f) ? 65535.f : ((f < -4.f) ? 0.f : ftrans);
  uint32_t u;
  if (sel != sel) /* NaN */
    u = 0;
  else if (sel <= 0.f)
    u = 0;
  else if (sel >= 4294967040.f)
    u = 0xFFFFFFFFu;
  else
    u = (uint32_t)sel; /* truncation toward zero */
  return (uint16_t)(u & 0xFFFFu);
}

/* pqt/triangle.cuh:14-18 (8/65536 is a power of two: fused or not is identical) */
float pqto_to_float(uint16_t s) { return (float)s * (8.f / 65536.f) - 4.f; }

/* pqt/triangle.cuh:55-63, as nvcc contracts it in device code:
 *   l2 = l*l; t = fma(c2, l2, b2); u = (a2 - b2) - c2; d = fma(u, l, t)        */
float pqto_dist(float a2, float b2, float c2, float lambda) {
  float l2 = lambda * lambda;
  float t = fmaf(c2, l2, b2);
  float u = (a2 - b2) - c2;
  return fmaf(u, lambda, t);
}

/* same expression with every operation rounded separately (host compilation) */
float pqto_dist_host(float a2, float b2, float c2, float lambda) {
  volatile float l2 = lambda * lambda;
  volatile float t1 = l2 * c2;
  volatile float t2 = b2 + t1;
  volatile float u1 = a2 - b2;
  volatile float u2 = u1 - c2;
  volatile float t3 = lambda * u2;
  return t2 + t3;
}

/* pqt/triangle.cuh:80-82 */
float pqto_project(float a2, float b2, float c2) {
  volatile float u1 = a2 - b2;
  volatile float u2 = u1 - c2;
  volatile float n = -0.5f * u2;
  return n / c2;
}

/* pqt/triangle.cuh:102-110 in the form the device code takes: nvcc emits mul, mul, sub
 * for d2 (the volatile out-parameter blocks the front-end contraction) and ptxas then
 * fuses the last two into FFMA d2 = fma(-c2, l*l, b2) (verified in the SASS for sm_100a
 * and against the reference's lineClusterKernelFast on the GPU, tests/test_ref_gpu.py). */
float pqto_project_d(float a2, float b2, float c2, float *d2) {
  float lambda = pqto_project(a2, b2, c2);
  volatile float l2 = lambda * lambda;
  *d2 = fmaf(-c2, l2, b2);
  return lambda;
}

/* same with every operation rounded separately (host compilation of triangle.cuh) */
float pqto_project_d_host(float a2, float b2, float c2, float *d2) {
  float lambda = pqto_project(a2, b2, c2);
  volatile float l2 = lambda * lambda;
  volatile float t = c2 * l2;
  *d2 = b2 - t;
  return lambda;
}

/* ------------------------------------------------------------------------- */
/* bitonicSort.cuh                                                           */
/* ------------------------------------------------------------------------- */

/* pqt/bitonicSort.cuh:16-44 (bitonic3) and :47-78 (bitonicLarge): identical
 * compare-exchange network; within one (k, j) stage all exchanges are disjoint. */
void pqto_bitonic(float *val, uint32_t *idx, uint32_t n) {
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = 0; i < n; i++) {
        uint32_t ixj = i ^ j;
        if (ixj > i && ixj < n) {
          int do_swap;
          if ((i & k) == 0)
            do_swap = val[i] > val[ixj];
          else
            do_swap = val[i] < val[ixj];
          if (do_swap) {
            float tv = val[i];
            val[i] = val[ixj];
            val[ixj] = tv;
            uint32_t ti = idx[i];
            idx[i] = idx[ixj];
            idx[ixj] = ti;
          }
        }
      }
    }
  }
}

/* pqt/bitonicSort.cuh:112-163: per-32 Hillis-Steele scan, then a scan of the
 * per-warp totals, then the fix-up.  Integer adds are exact, so the result is
 * the ordinary prefix sum; the structure is kept to mirror the reference. */
void pqto_scan(uint32_t *v, uint32_t n, int inclusive) {
  uint32_t nw = (n + 31) / 32;
  uint32_t *tot = (uint32_t *)malloc(sizeof(uint32_t) * (nw ? nw : 1));
  for (uint32_t w = 0; w < nw; w++) {
    uint32_t lo = w * 32, hi = lo + 32 < n ? lo + 32 : n;
    uint32_t tmp[32];
    for (uint32_t d = 1; d < 32; d <<= 1) {
      for (uint32_t i = lo; i < hi; i++) tmp[i - lo] = (i - lo >= d) ? v[i - d] + v[i] : v[i];
      for (uint32_t i = lo; i < hi; i++) v[i] = tmp[i - lo];
    }
    tot[w] = v[hi - 1];
    if (!inclusive) {
      for (uint32_t i = hi - 1; i > lo; i--) v[i] = v[i - 1];
      v[lo] = 0;
    }
  }
  uint32_t run = 0;
  for (uint32_t w = 0; w < nw; w++) {
    uint32_t lo = w * 32, hi = lo + 32 < n ? lo + 32 : n;
    for (uint32_t i = lo; i < hi; i++) v[i] += run;
    run += tot[w];
  }
  free(tot);
}

/* ------------------------------------------------------------------------- */
/* segment distance                                                          */
/* ------------------------------------------------------------------------- */

/* pqt/PerturbationProTree.cu:7146-7160: s[t] = sqr(b - a) (sub and mul rounded
 * separately, the value passes through shared memory), then for
 * stride = len/2 .. 1: s[j] += s[j + stride], j < stride.  len = power of two. */
float pqto_seg_dist(const float *q, const float *c, uint32_t len) {
  float s[1024];
  for (uint32_t t = 0; t < len; t++) {
    float d = q[t] - c[t]; /* -ffp-contract=off: sub and mul stay separate */
    s[t] = d * d;
  }
  for (uint32_t stride = len >> 1; stride > 0; stride >>= 1)
    for (uint32_t j = 0; j < stride; j++) s[j] = s[j] + s[j + stride];
  return s[0];
}

/* ------------------------------------------------------------------------- */
/* a2: prepareDistSequence                                                   */
/* ------------------------------------------------------------------------- */

typedef struct {
  float d;
  uint32_t code;
} seq_pair;

static int seq_pair_cmp(const void *a, const void *b) {
  const seq_pair *x = (const seq_pair *)a, *y = (const seq_pair *)b;
  if (x->d < y->d) return -1;
  if (x->d > y->d) return 1;
  return (x->code > y->code) - (x->code < y->code);
}

/* pqt/ProTree.cu:128-207.  std::sort of pair<float,uint> orders lexicographically
 * and all codes are distinct, so the order is unique (no stability question). */
uint32_t pqto_dist_seq(uint32_t max_cluster, uint32_t p, uint32_t *seq, uint32_t *n_valid) {
  uint32_t m = max_cluster > 16 ? 16 : max_cluster;
  uint64_t n_vec64 = 1;
  for (uint32_t j = 0; j < p; j++) n_vec64 *= m;
  uint32_t n_vec = (uint32_t)n_vec64; /* "uint nVec = pow(...)" :139 */
  seq_pair *d = (seq_pair *)malloc(sizeof(seq_pair) * (n_vec ? n_vec : 1));
  uint32_t denom[32];
  denom[0] = 1;
  for (uint32_t j = 1; j < p; j++) denom[j] = denom[j - 1] * m;
  for (uint32_t i = 0; i < n_vec; i++) {
    float dist = 0.f;
    for (uint32_t j = 0; j < p; j++) {
      uint32_t v = (i / denom[j]) % m;
      dist = dist + sqrtf((float)v);
    }
    d[i].d = dist;
    d[i].code = i;
  }
  qsort(d, n_vec, sizeof(seq_pair), seq_pair_cmp);
  for (uint32_t i = 0; i < PQTO_NUM_DISTSEQ; i++) seq[i] = 0;
  uint32_t keep = n_vec < PQTO_NUM_DISTSEQ ? n_vec : PQTO_NUM_DISTSEQ;
  for (uint32_t i = 0; i < keep; i++) seq[i] = d[i].code;
  free(d);
  if (n_valid) *n_valid = keep;
  return m;
}

/* ------------------------------------------------------------------------- */
/* a9: cbDist                                                                */
/* ------------------------------------------------------------------------- */

/* computeCBL1L1Dist :1902-1917 calls calcDist(res, cb1, cb1, c1, c1, dim, LP):
 * res[(iter*c1 + a)*LP + lp] = segdist(cb1[iter], cb1[a]) with b = B[iter], a = A[a]
 * (pqt/ProQuantization.cu:113-132). */
void pqto_cb_dist(const pqto_params *prm, const float *cb1, float *cb_dist) {
  uint32_t LP = prm->line_parts, c1 = prm->c1, dim = prm->dim, sl = dim / LP;
  for (uint32_t b = 0; b < c1; b++)
    for (uint32_t a = 0; a < c1; a++)
      for (uint32_t lp = 0; lp < LP; lp++)
        cb_dist[(b * c1 + a) * LP + lp] =
            pqto_seg_dist(cb1 + (size_t)b * dim + lp * sl, cb1 + (size_t)a * dim + lp * sl, sl);
}

/* ------------------------------------------------------------------------- */
/* Steps A..E                                                                */
/* ------------------------------------------------------------------------- */

/* :7105-7212.  assign[k*p + part] = k-th best L1 centroid of that part. */
void pqto_step_a(const pqto_params *prm, const float *cb1, const float *q, uint32_t k1,
                 uint32_t *assign) {
  uint32_t p = prm->p, c1 = prm->c1, dim = prm->dim, vl = dim / p;
  uint32_t np2 = pqto_pow2ceil(c1);
  float val[1024];
  uint32_t idx[1024];
  for (uint32_t part = 0; part < p; part++) {
    for (uint32_t i = 0; i < np2; i++) {
      val[i] = 10000000.f; /* :7185 */
      idx[i] = PQTO_PAD_IDX;
    }
    for (uint32_t c = 0; c < c1; c++) {
      val[c] = pqto_seg_dist(q + part * vl, cb1 + (size_t)c * dim + part * vl, vl);
      idx[c] = c;
    }
    pqto_bitonic(val, idx, np2);
    for (uint32_t k = 0; k < k1; k++) assign[k * p + part] = idx[k];
  }
}

/* :7739-7799.  lut[lp*c1 + c] */
void pqto_step_b(const pqto_params *prm, const float *cb1, const float *q, float *lut) {
  uint32_t LP = prm->line_parts, c1 = prm->c1, dim = prm->dim, sl = dim / LP;
  for (uint32_t c = 0; c < c1; c++)
    for (uint32_t lp = 0; lp < LP; lp++)
      lut[lp * c1 + c] = pqto_seg_dist(q + lp * sl, cb1 + (size_t)c * dim + lp * sl, sl);
}

/* :1534-1664.  cb2 layout [p][c1][c2][vl] (getCBIdx, pqt/ProTree.hh:17-21). */
void pqto_step_c(const pqto_params *prm, const float *cb2, const float *q, uint32_t k1,
                 const uint32_t *assign, float *assign_val, uint32_t *assign_idx) {
  uint32_t p = prm->p, c1 = prm->c1, c2 = prm->c2, dim = prm->dim, vl = dim / p;
  uint32_t n = k1 * c2, np2 = pqto_pow2ceil(n);
  float *val = (float *)malloc(sizeof(float) * np2);
  uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * np2);
  for (uint32_t part = 0; part < p; part++) {
    for (uint32_t i = 0; i < np2; i++) {
      val[i] = 1000000000.f; /* :1627 */
      idx[i] = PQTO_PAD_IDX;
    }
    for (uint32_t k = 0; k < k1; k++) {
      uint32_t l1 = assign[k * p + part];
      const float *cb = cb2 + (size_t)(part * c1 + l1) * vl * c2;
      for (uint32_t l2 = 0; l2 < c2; l2++) {
        val[k * c2 + l2] = pqto_seg_dist(q + part * vl, cb + (size_t)l2 * vl, vl);
        idx[k * c2 + l2] = l2 + l1 * c2; /* c1scale = c2, :1684 */
      }
    }
    pqto_bitonic(val, idx, np2);
    for (uint32_t i = 0; i < n; i++) {
      assign_val[part * n + i] = val[i];
      assign_idx[part * n + i] = idx[i];
    }
  }
  free(val);
  free(idx);
}

/* :3374-3549 (+ the memsets in getBins :3561-3563). */
uint32_t pqto_step_d(const pqto_params *prm, uint32_t k1, const uint32_t *dist_seq, uint32_t m,
                     const uint32_t *assign_idx, const uint32_t *bin_counts, uint32_t *bins) {
  uint32_t p = prm->p, T = prm->bin_threads, n = k1 * prm->c2;
  uint32_t max_out = prm->max_bins;
  uint32_t denom[32];
  denom[0] = 1;
  for (uint32_t j = 1; j < p; j++) denom[j] = denom[j - 1] * m;
  memset(bins, 0, sizeof(uint32_t) * max_out);
  uint32_t n_out = 0, n_iter = 0;
  /* nElements is never updated in the reference (stays 0 < _k) */
  while (n_iter < prm->max_trials && n_out < max_out) {
    uint32_t kept = 0;
    for (uint32_t t = 0; t < T; t++) {
      uint32_t s = dist_seq[n_iter * T + t];
      uint32_t o = 0;
      for (uint32_t j = 0; j < p; j++) {
        uint32_t bp = (s / denom[j]) % m;
        o = o * prm->c1 * prm->c2 + assign_idx[j * n + bp]; /* uint32 wrap */
      }
      uint32_t bin = o % prm->hash_size;
      if (bin_counts[bin]) {
        kept++;                      /* inclusive scan value of this thread */
        uint32_t pos = kept + n_out; /* 1-based: slot 0 keeps the memset 0 */
        if (pos < max_out) bins[pos] = bin;
      }
    }
    n_out += kept;
    n_iter++;
  }
  return n_out > max_out ? max_out : n_out;
}

/* :4308-4419.  Net effect of the chunked scan: concatenate, in bin-list order,
 * the first min(count, max_vec_per_bin) ids of every listed bin, truncated at
 * max_vec. */
uint32_t pqto_step_e1(const pqto_params *prm, const uint32_t *bins, uint32_t n_bins,
                      const uint32_t *bin_prefix, const uint32_t *bin_counts,
                      const uint32_t *db_idx, uint32_t max_vec, uint32_t *select_idx) {
  uint32_t T = max_vec < 1024 ? max_vec : 1024; /* block size, :6165 */
  uint32_t nb = n_bins < prm->max_bins ? n_bins : prm->max_bins;
  uint32_t bin_iter = nb / T + 1;
  uint32_t offset = 0;
  memset(select_idx, 0, sizeof(uint32_t) * max_vec);
  for (uint32_t it = 0; it < bin_iter; it++) {
    uint32_t run = offset, last_pos = offset, last_nv = 0;
    for (uint32_t t = 0; t < T; t++) {
      uint32_t b = it * T + t;
      uint32_t nv = 0, cur = 0;
      if (b < nb) {
        cur = bins[b];
        nv = bin_counts[cur] < prm->max_vec_per_bin ? bin_counts[cur] : prm->max_vec_per_bin;
      }
      uint32_t pos = run; /* exclusive scan + offset */
      run += nv;
      if (pos + nv > max_vec) nv = (pos >= max_vec) ? 0 : (max_vec - pos);
      if (b < nb)
        for (uint32_t v = 0; v < nv; v++) select_idx[pos + v] = db_idx[bin_prefix[cur] + v];
      last_pos = pos;
      last_nv = nv;
    }
    offset = last_pos + last_nv;
  }
  return offset > max_vec ? max_vec : offset;
}

/* one candidate: :5295-5310 + warpReduceSum :5183-5187 */
float pqto_line_adc(const pqto_params *prm, const float *lut, const float *cb_dist,
                    const uint32_t *code) {
  uint32_t LP = prm->line_parts, c1 = prm->c1;
  float v[64];
  for (uint32_t lp = 0; lp < LP; lp++) {
    uint32_t w = code[lp];
    uint32_t l1 = (uint32_t)(int32_t)(int8_t)(w & 0xFFu);        /* char p1 */
    uint32_t l2 = (uint32_t)(int32_t)(int8_t)((w >> 8) & 0xFFu); /* char p2 */
    float lambda = pqto_to_float((uint16_t)(w >> 16));
    float cc = cb_dist[l2 * c1 * LP + l1 * LP + lp];
    v[lp] = pqto_dist(lut[lp * c1 + l1], lut[lp * c1 + l2], cc, lambda);
  }
  for (uint32_t s = LP >> 1; s > 0; s >>= 1)
    for (uint32_t i = 0; i < s; i++) v[i] = v[i] + v[i + s];
  return v[0];
}

/* :5189-5351 */
void pqto_step_e2(const pqto_params *prm, const float *lut, const float *cb_dist,
                  const uint32_t *lines, const uint32_t *select_idx, uint32_t n_vec,
                  uint32_t max_vec, uint32_t k, float *out_dist, uint32_t *out_idx) {
  uint32_t LP = prm->line_parts;
  float *val = (float *)malloc(sizeof(float) * max_vec);
  uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * max_vec);
  uint32_t nv = n_vec < max_vec ? n_vec : max_vec;
  for (uint32_t a = 0; a < max_vec; a++) {
    if (a < nv) {
      idx[a] = select_idx[a];
      val[a] = pqto_line_adc(prm, lut, cb_dist, lines + (size_t)idx[a] * LP);
    } else {
      val[a] = 10000000.f; /* :5333 */
      idx[a] = PQTO_PAD_IDX;
    }
  }
  pqto_bitonic(val, idx, max_vec);
  for (uint32_t i = 0; i < k; i++) {
    out_dist[i] = val[i];
    out_idx[i] = idx[i];
  }
  free(val);
  free(idx);
}

/* ------------------------------------------------------------------------- */
/* a1: queryKNN                                                              */
/* ------------------------------------------------------------------------- */

static int shapes_ok(const pqto_params *prm) {
  uint32_t vl, sl;
  if (!prm->p || !prm->line_parts || !prm->c1 || !prm->c2) return 0;
  if (prm->dim % prm->p || prm->dim % prm->line_parts) return 0;
  vl = prm->dim / prm->p;
  sl = prm->dim / prm->line_parts;
  if ((vl & (vl - 1)) || (sl & (sl - 1))) return 0; /* tree reduction needs 2^n */
  if (prm->line_parts & (prm->line_parts - 1)) return 0;
  if (prm->line_parts > 32) return 0; /* shuffle reduction inside one warp */
  if (prm->c1 > 127 || prm->c1 > 1024 || prm->dim > 1024 || prm->p > 8) return 0;
  return 1;
}

int pqto_query_knn(const pqto_params *prm, const float *cb1, const float *cb2,
                   const uint32_t *bin_prefix, const uint32_t *bin_counts,
                   const uint32_t *db_idx, const uint32_t *lines, const float *Q, uint32_t QN,
                   uint32_t k, float *out_dist, uint32_t *out_idx, pqto_stages *st,
                   int nthreads) {
  if (!shapes_ok(prm) || prm->k1 > prm->c1 || k == 0) return -1;
  uint32_t p = prm->p, c1 = prm->c1, c2 = prm->c2, LP = prm->line_parts, k1 = prm->k1;
  uint32_t n = k1 * c2;
  uint32_t max_vec = pqto_pow2ceil(k); /* rerankKBestVectors :6159 */
  if (prm->max_trials * prm->bin_threads > PQTO_NUM_DISTSEQ) return -1;

  uint32_t *seq = (uint32_t *)malloc(sizeof(uint32_t) * PQTO_NUM_DISTSEQ);
  uint32_t m = pqto_dist_seq(c2 * k1, p, seq, NULL); /* :8191 */
  if (m > n) {
    free(seq);
    return -1;
  }
  float *cb_dist = (float *)malloc(sizeof(float) * (size_t)c1 * c1 * LP);
  pqto_cb_dist(prm, cb1, cb_dist);

#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    uint32_t *assign = (uint32_t *)malloc(sizeof(uint32_t) * k1 * p);
    float *lut = (float *)malloc(sizeof(float) * LP * c1);
    float *aval = (float *)malloc(sizeof(float) * p * n);
    uint32_t *aidx = (uint32_t *)malloc(sizeof(uint32_t) * p * n);
    uint32_t *bins = (uint32_t *)malloc(sizeof(uint32_t) * prm->max_bins);
    uint32_t *sel = (uint32_t *)malloc(sizeof(uint32_t) * max_vec);
#pragma omp for schedule(dynamic, 4)
    for (int64_t qi = 0; qi < (int64_t)QN; qi++) {
      const float *q = Q + (size_t)qi * prm->dim;
      pqto_step_a(prm, cb1, q, k1, assign);
      pqto_step_b(prm, cb1, q, lut);
      pqto_step_c(prm, cb2, q, k1, assign, aval, aidx);
      uint32_t nb = pqto_step_d(prm, k1, seq, m, aidx, bin_counts, bins);
      uint32_t nv = pqto_step_e1(prm, bins, nb, bin_prefix, bin_counts, db_idx, max_vec, sel);
      pqto_step_e2(prm, lut, cb_dist, lines, sel, nv, max_vec, k, out_dist + (size_t)qi * k,
                   out_idx + (size_t)qi * k);
      if (st) {
        if (st->assign) memcpy(st->assign + (size_t)qi * k1 * p, assign, sizeof(uint32_t) * k1 * p);
        if (st->lut) memcpy(st->lut + (size_t)qi * LP * c1, lut, sizeof(float) * LP * c1);
        if (st->assign_val) memcpy(st->assign_val + (size_t)qi * p * n, aval, sizeof(float) * p * n);
        if (st->assign_idx)
          memcpy(st->assign_idx + (size_t)qi * p * n, aidx, sizeof(uint32_t) * p * n);
        if (st->bins)
          memcpy(st->bins + (size_t)qi * prm->max_bins, bins, sizeof(uint32_t) * prm->max_bins);
        if (st->n_bins) st->n_bins[qi] = nb;
        if (st->select_idx)
          memcpy(st->select_idx + (size_t)qi * max_vec, sel, sizeof(uint32_t) * max_vec);
        if (st->n_vec) st->n_vec[qi] = nv;
      }
    }
    free(assign);
    free(lut);
    free(aval);
    free(aidx);
    free(bins);
    free(sel);
  }
  free(cb_dist);
  free(seq);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* a11: queryBIGKNNRerank2                                                   */
/* ------------------------------------------------------------------------- */

/* pqt/ProTree.cu:50-126.  Host arithmetic of the reference: s = (float)pow(0.9 * 1.2f,
 * slope - 5) in double; dist = powf(x, 0.8f) + s * powf(y, 0.8f) in float; pairs sorted
 * lexicographically; the first min(maxCluster^2, 65536) codes of every slope are kept. */
void pqto_dist_seq_2d(uint32_t max_cluster, uint32_t *seq) {
  uint32_t n_vec = max_cluster * max_cluster;
  uint32_t copy = n_vec < PQTO_NUM_DISTSEQ ? n_vec : PQTO_NUM_DISTSEQ;
  seq_pair *d = (seq_pair *)malloc(sizeof(seq_pair) * n_vec);
  for (int slope = 0; slope < PQTO_NUM_ANISO_DIR; slope++) {
    float s = (float)pow(0.9 * (double)PQTO_ANISO_BASE, (double)(slope - (PQTO_NUM_ANISO_DIR / 2)));
    for (uint32_t i = 0; i < n_vec; i++) {
      float x = (float)(i % max_cluster);
      float y = (float)(i / max_cluster);
      float n = 0.8f;
      volatile float py = s * powf(y, n);
      d[i].d = powf(x, n) + py;
      d[i].code = i;
    }
    qsort(d, n_vec, sizeof(seq_pair), seq_pair_cmp);
    for (uint32_t i = 0; i < PQTO_NUM_DISTSEQ; i++) seq[(size_t)slope * PQTO_NUM_DISTSEQ + i] = 0;
    for (uint32_t i = 0; i < copy; i++) seq[(size_t)slope * PQTO_NUM_DISTSEQ + i] = d[i].code;
  }
  free(d);
}

/* :2839-2862 (device code: sqrtf, IEEE division, logf, roundf; float -> int conversion
 * saturates and maps NaN to 0) */
uint32_t pqto_slope_idx(const float *val0, const float *val1, uint32_t N, int *ambiguous) {
  uint32_t sample = (uint32_t)sqrtf(2.f * (float)N);
  float num = (val1[sample] + val1[sample - 1]) - 2.f * val1[0];
  float den = (val0[sample] + val0[sample - 1]) - 2.f * val0[0];
  float slope = num / den;
  float r = logf(slope) / logf(PQTO_ANISO_BASE);
  float f = roundf(r) + (float)(PQTO_NUM_ANISO_DIR / 2);
  int si;
  if (f != f)
    si = 0;
  else if (f >= 2147483648.f)
    si = 2147483647;
  else if (f <= -2147483648.f)
    si = -2147483647 - 1;
  else
    si = (int)f;
  if (si >= PQTO_NUM_ANISO_DIR) si = PQTO_NUM_ANISO_DIR - 1;
  if (si < 0) si = 0;
  if (ambiguous && r == r) {
    float frac = fabsf(r - floorf(r) - 0.5f);
    if (frac < 1e-3f && r > -6.f && r < 5.f) *ambiguous |= 1;
  }
  return (uint32_t)si;
}

uint32_t pqto_step_d_big(const pqto_params *prm, uint32_t k1, const uint32_t *seq2d,
                         const float *assign_val, const uint32_t *assign_idx,
                         const uint32_t *bin_counts, uint32_t k2, uint32_t *bins, int *ambiguous) {
  const uint32_t n = k1 * prm->c2, K = prm->c1 * prm->c2;
  const uint32_t kMax = PQTO_BIG_KMAX, nI = PQTO_BIG_NINTER, dc = PQTO_BIG_DISTCLUSTER;
  const uint32_t T = prm->bin_threads; /* 1024 */
  const uint32_t max_out = prm->max_bins;
  float *ival = (float *)malloc(sizeof(float) * 2 * nI);
  uint32_t *iidx = (uint32_t *)malloc(sizeof(uint32_t) * 2 * nI);
  float *dist = (float *)malloc(sizeof(float) * T);
  uint32_t *oidx = (uint32_t *)malloc(sizeof(uint32_t) * T);
  uint32_t *nel = (uint32_t *)malloc(sizeof(uint32_t) * T);
  memset(bins, 0, sizeof(uint32_t) * (size_t)max_out);

  /* ---- selectBinKernel2D2Parts: parts (0,1) -> list 0, parts (2,3) -> list 1 */
  for (uint32_t pi = 0; pi < prm->p / 2; pi++) {
    const float *v0 = assign_val + (size_t)(2 * pi) * n, *v1 = assign_val + (size_t)(2 * pi + 1) * n;
    const uint32_t *i0 = assign_idx + (size_t)(2 * pi) * n, *i1 = assign_idx + (size_t)(2 * pi + 1) * n;
    uint32_t si = pqto_slope_idx(v0, v1, nI, ambiguous);
    const uint32_t *seq = seq2d + (size_t)si * PQTO_NUM_DISTSEQ;
    float *dv = ival + pi * nI;
    uint32_t *di = iidx + pi * nI;
    for (uint32_t t = 0; t < nI; t++) {
      uint32_t s = seq[t], x = s % dc, y = s / dc;
      if (x < kMax && y < kMax) {
        dv[t] = v0[x] + v1[y];
        di[t] = i0[x] * K + i1[y];
      } else {
        dv[t] = 99999999999.f;
        di[t] = 0;
      }
    }
    if (pi < 2) pqto_bitonic(dv, di, nI);
  }
  /* ---- selectBinKernel2DFinal */
  uint32_t si = pqto_slope_idx(ival, ival + nI, 1024, ambiguous);
  const size_t seq_total = (size_t)PQTO_NUM_ANISO_DIR * PQTO_NUM_DISTSEQ;
  uint32_t n_out = 0, n_elements = 0, n_iter = 0;
  const uint32_t factor = K * K; /* uint32 wrap, :3101 */
  while (n_elements < k2 && n_iter < prm->max_trials && n_out < max_out) {
    size_t off = (size_t)si * PQTO_NUM_DISTSEQ + (size_t)n_iter * T;
    if (off + T > seq_total) { /* the reference reads past d_distSeq from here on */
      if (ambiguous) *ambiguous |= 2;
      break;
    }
    for (uint32_t t = 0; t < T; t++) {
      uint32_t s = seq2d[off + t], x = s % dc, y = s / dc;
      if (x < nI && y < nI) {
        dist[t] = ival[x] + ival[nI + y];
        oidx[t] = iidx[x] * factor + iidx[nI + y];
      } else {
        dist[t] = 99999999999.f;
        oidx[t] = 0;
      }
      oidx[t] = oidx[t] % prm->hash_size;
    }
    pqto_bitonic(dist, oidx, T);
    uint32_t run = 0, total = 0;
    for (uint32_t t = 0; t < T; t++) {
      uint32_t c = bin_counts[oidx[t]];
      total += c < 2 ? c : 2; /* maxVecPB = 2, :3114 */
    }
    uint32_t kept_before = 0, kept_excl_last = 0;
    for (uint32_t t = 0; t < T; t++) {
      uint32_t c = bin_counts[oidx[t]];
      uint32_t e = c < 2 ? c : 2;
      uint32_t reg = e;
      /* inclusive scan of the previous thread + nElements >= k -> dropped (:3126-3129) */
      if (t > 0 && (run + n_elements) >= k2) reg = 0;
      run += e;
      if (reg) {
        uint32_t pos = kept_before + n_out; /* exclusive scan: 0-based (:3146-3151) */
        if (pos < max_out) bins[pos] = oidx[t];
        kept_before++;
      }
      if (t == T - 2) kept_excl_last = kept_before;
    }
    if (T == 1) kept_excl_last = 0;
    n_elements += total;
    /* nOutBins += nElem[blockDim.x - 1] of the EXCLUSIVE scan: the last thread's own flag is
     * not counted (:3163) */
    n_out += kept_excl_last;
    n_iter++;
  }
  free(ival);
  free(iidx);
  free(dist);
  free(oidx);
  free(nel);
  return n_out > max_out ? max_out : n_out;
}

int pqto_query_big_knn_rerank2(const pqto_params *prm, const float *cb1, const float *cb2,
                               const uint32_t *bin_prefix, const uint32_t *bin_counts,
                               const uint32_t *db_idx, const uint32_t *lines, const float *Q,
                               uint32_t QN, uint32_t k, float *out_dist, uint32_t *out_idx,
                               uint32_t *n_bins_out, uint32_t *n_vec_out, int *ambiguous,
                               int nthreads) {
  if (!shapes_ok(prm) || prm->k1 > prm->c1 || k == 0 || prm->p != 4) return -1;
  if (prm->k1 * prm->c2 < PQTO_BIG_KMAX) return -1; /* reads 64 sorted entries per part */
  uint32_t p = prm->p, c1 = prm->c1, c2 = prm->c2, LP = prm->line_parts, k1 = prm->k1;
  uint32_t n = k1 * c2;
  uint32_t max_vec = pqto_pow2ceil(k);
  uint32_t *seq2d = (uint32_t *)malloc(sizeof(uint32_t) * PQTO_NUM_ANISO_DIR * PQTO_NUM_DISTSEQ);
  pqto_dist_seq_2d(PQTO_BIG_DISTCLUSTER, seq2d);
  float *cb_dist = (float *)malloc(sizeof(float) * (size_t)c1 * c1 * LP);
  pqto_cb_dist(prm, cb1, cb_dist);
  pqto_params e1 = *prm;
  e1.max_vec_per_bin = max_vec; /* :6525: maxNVecPerBin = nnn */
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    uint32_t *assign = (uint32_t *)malloc(sizeof(uint32_t) * k1 * p);
    float *lut = (float *)malloc(sizeof(float) * LP * c1);
    float *aval = (float *)malloc(sizeof(float) * p * n);
    uint32_t *aidx = (uint32_t *)malloc(sizeof(uint32_t) * p * n);
    uint32_t *bins = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)prm->max_bins);
    uint32_t *sel = (uint32_t *)malloc(sizeof(uint32_t) * max_vec);
#pragma omp for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < (int64_t)QN; qi++) {
      const float *q = Q + (size_t)qi * prm->dim;
      int amb = 0;
      pqto_step_a(prm, cb1, q, k1, assign);
      pqto_step_b(prm, cb1, q, lut);
      pqto_step_c(prm, cb2, q, k1, assign, aval, aidx);
      uint32_t nb = pqto_step_d_big(prm, k1, seq2d, aval, aidx, bin_counts, k, bins, &amb);
      uint32_t nv = pqto_step_e1(&e1, bins, nb, bin_prefix, bin_counts, db_idx, max_vec, sel);
      pqto_step_e2(prm, lut, cb_dist, lines, sel, nv, max_vec, k, out_dist + (size_t)qi * k,
                   out_idx + (size_t)qi * k);
      if (n_bins_out) n_bins_out[qi] = nb;
      if (n_vec_out) n_vec_out[qi] = nv;
      if (ambiguous) ambiguous[qi] = amb;
    }
    free(assign);
    free(lut);
    free(aval);
    free(aidx);
    free(bins);
    free(sel);
  }
  free(cb_dist);
  free(seq2d);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* build side                                                                */
/* ------------------------------------------------------------------------- */

/* buildKBestDB :1231-1315: Step A with k1 (=16), then
 * assignPerturbationBestBinKernel2 :830-942: over k = 0..k1-1, l2 = 0..c2-1 in
 * that order keep the strictly smaller distance (first entry unconditionally),
 * idx_p = l2 + l1_k*c2; bin = uint32 Horner over parts, % HASH_SIZE. */
void pqto_assign_bins(const pqto_params *prm, const float *cb1, const float *cb2, const float *X,
                      uint32_t N, uint32_t k1, uint32_t *bin_of, int nthreads) {
  uint32_t p = prm->p, c1 = prm->c1, c2 = prm->c2, dim = prm->dim, vl = dim / p;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    uint32_t *assign = (uint32_t *)malloc(sizeof(uint32_t) * k1 * p);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < (int64_t)N; i++) {
      const float *x = X + (size_t)i * dim;
      pqto_step_a(prm, cb1, x, k1, assign);
      uint32_t o = 0;
      for (uint32_t part = 0; part < p; part++) {
        float best = 0.f;
        uint32_t best_idx = 0;
        for (uint32_t k = 0; k < k1; k++) {
          uint32_t l1 = assign[k * p + part];
          const float *cb = cb2 + (size_t)(part * c1 + l1) * vl * c2;
          for (uint32_t l2 = 0; l2 < c2; l2++) {
            float v = pqto_seg_dist(x + part * vl, cb + (size_t)l2 * vl, vl);
            if ((best > v) || ((k + l2) == 0)) {
              best = v;
              best_idx = l2 + l1 * c2;
            }
          }
        }
        o = (part == 0) ? best_idx : (o * c1 * c2 + best_idx); /* :929-931, uint32 wrap */
      }
      bin_of[i] = o % prm->hash_size;
    }
    free(assign);
  }
}

void pqto_build_lists(const uint32_t *bin_of, uint32_t N, uint32_t hash_size, uint32_t *bin_counts,
                      uint32_t *bin_prefix, uint32_t *db_idx) {
  memset(bin_counts, 0, sizeof(uint32_t) * (size_t)hash_size);
  for (uint32_t i = 0; i < N; i++) bin_counts[bin_of[i]]++;
  uint32_t run = 0;
  for (uint32_t b = 0; b < hash_size; b++) {
    bin_prefix[b] = run;
    run += bin_counts[b];
  }
  /* ascending id inside a bin (the reference's atomicInc order is run-dependent) */
  uint32_t *cursor = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)hash_size);
  memcpy(cursor, bin_prefix, sizeof(uint32_t) * (size_t)hash_size);
  for (uint32_t i = 0; i < N; i++) db_idx[cursor[bin_of[i]]++] = i;
  free(cursor);
}

/* lineClusterKernelFast :7527-7661.  Thread (p, cIdx) walks minId = 0..c1-1 and
 * keeps the first strictly smallest d2 (minId == 0 unconditionally, cIdx == minId
 * forced to 999999999999.f); then a tree over cIdx (stride c1/2..1) keeps the lower
 * index unless the upper one is strictly smaller.  c1 must be a power of two for
 * that tree to cover every centroid. */
void pqto_line_encode(const pqto_params *prm, const float *cb1, const float *cb_dist,
                      const float *X, uint32_t N, uint32_t *lines, int nthreads) {
  uint32_t LP = prm->line_parts, c1 = prm->c1, dim = prm->dim, sl = dim / LP;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    float *val = (float *)malloc(sizeof(float) * c1);
    float *dbest = (float *)malloc(sizeof(float) * c1);
    uint32_t *code = (uint32_t *)malloc(sizeof(uint32_t) * c1);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < (int64_t)N; i++) {
      const float *x = X + (size_t)i * dim;
      for (uint32_t lp = 0; lp < LP; lp++) {
        for (uint32_t c = 0; c < c1; c++)
          val[c] = pqto_seg_dist(x + lp * sl, cb1 + (size_t)c * dim + lp * sl, sl);
        for (uint32_t c = 0; c < c1; c++) {
          for (uint32_t mn = 0; mn < c1; mn++) {
            float cc = cb_dist[mn * c1 * LP + c * LP + lp];
            float d;
            float l = pqto_project_d(val[c], val[mn], cc, &d);
            if (c == mn) d = 999999999999.f;
            if ((mn == 0) || (d < dbest[c])) {
              dbest[c] = d;
              code[c] = (c & 0xFFu) | ((mn & 0xFFu) << 8) | ((uint32_t)pqto_to_ushort(l) << 16);
            }
          }
        }
        for (uint32_t stride = c1 >> 1; stride > 0; stride >>= 1)
          for (uint32_t c = 0; c < stride; c++)
            if (dbest[c] > dbest[c + stride]) {
              dbest[c] = dbest[c + stride];
              code[c] = code[c + stride];
            }
        lines[(size_t)i * LP + lp] = code[0];
      }
    }
    free(val);
    free(dbest);
    free(code);
  }
}

/* ------------------------------------------------------------------------- */
/* brute force (double accumulation; ground truth for recall only)           */
/* ------------------------------------------------------------------------- */
void pqto_brute_force_1nn(const float *X, uint32_t N, const float *Q, uint32_t QN, uint32_t dim,
                          uint32_t *nn, int nthreads) {
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
  for (int64_t qi = 0; qi < (int64_t)QN; qi++) {
    const float *q = Q + (size_t)qi * dim;
    double best = 1e300;
    uint32_t bi = 0;
    for (uint32_t i = 0; i < N; i++) {
      const float *x = X + (size_t)i * dim;
      double s = 0;
      for (uint32_t t = 0; t < dim; t++) {
        double d = (double)q[t] - (double)x[t];
        s += d * d;
      }
      if (s < best) {
        best = s;
        bi = i;
      }
    }
    nn[qi] = bi;
  }
}
