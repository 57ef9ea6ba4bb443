/*
 * pqt_oracle.h -- CPU oracle for the Product-Quantization-Tree query path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * GPU `PerturbationProTree::queryKNN` chain (and of the build-side steps needed
 * to create its inputs).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call it, and only as the checker /
 * CPU baseline.  The product path (libpqt_b200.so) never links or calls this.
 *
 * Parity pins: the triangle arithmetic, the lambda quantiser, the bitonic
 * network and the scans are pinned to the known answers in the reference's
 * run.cu:9-115 and pqt/bitonicSort.cuh:213-252, and the triangle functions
 * are additionally checked bit-for-bit against the reference's own
 * pqt/triangle.cuh compiled for the host (oracle/_ref, see oracle/Makefile).
 * The reference holds no end-to-end golden vectors, so
 * end-to-end parity is pinned by fixtures recorded from this oracle
 * (tests/golden/) and, on the GPU box, by the reference's own .cu kernels
 * compiled unmodified into oracle/_ref/libpqt_ref.so (tests/test_ref_gpu.py).
 *
 * All citations are file:line into /root/reference.
 */
#ifndef PQT_ORACLE_H
#define PQT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQTO_NUM_DISTSEQ 65536u /* pqt/ProTree.hh:9 NUM_DISTSEQ */
#define PQTO_PAD_IDX 0xFFFFFFFFu /* id reported for padded result slots (the
                                    reference leaves stale shared memory there,
                                    pqt/PerturbationProTree.cu:5331-5334) */

typedef struct pqto_params {
  uint32_t dim;             /* vector dimension                                   */
  uint32_t p;               /* tree parts (d_p)                                   */
  uint32_t c1;              /* level-1 centroids per part (d_nClusters)           */
  uint32_t c2;              /* level-2 centroids per (part, L1 cell) (d_nClusters2) */
  uint32_t line_parts;      /* LP, d_lineParts                                    */
  uint32_t k1;              /* L1 cells expanded per part; 8 in queryKNN (:8187)  */
  uint32_t max_bins;        /* 4096 (:8218)                                       */
  uint32_t max_trials;      /* 16 (:3569)                                         */
  uint32_t bin_threads;     /* 1024 = block size of selectBinKernelFast2 (:3556)  */
  uint32_t max_vec_per_bin; /* 2800 (:6208)                                       */
  uint32_t hash_size;       /* HASH_SIZE, 400000000 (pqt/PerturbationProTree.hh:12) */
} pqto_params;

/* reference literals (queryKNN operating point) */
void pqto_default_params(pqto_params *prm, uint32_t dim, uint32_t p, uint32_t c1,
                         uint32_t c2, uint32_t line_parts);

/* ---- helper.hh / triangle.cuh ------------------------------------------------ */
uint32_t pqto_pow2ceil(uint32_t x);            /* pqt/helper.hh:27-37 ("log2")      */
uint16_t pqto_to_ushort(float f);              /* pqt/triangle.cuh:6-12             */
float pqto_to_float(uint16_t s);               /* pqt/triangle.cuh:14-18            */
float pqto_dist(float a2, float b2, float c2, float lambda); /* :55-63, device (FMA-contracted) form */
float pqto_dist_host(float a2, float b2, float c2, float lambda); /* :55-63, uncontracted host form  */
float pqto_project(float a2, float b2, float c2);            /* :80-82            */
float pqto_project_d(float a2, float b2, float c2, float *d2); /* :102-110, device (FFMA) form */
float pqto_project_d_host(float a2, float b2, float c2, float *d2); /* :102-110, host form */

/* ---- bitonicSort.cuh ---------------------------------------------------------- */
/* ascending key/value bitonic network over n = power of two (bitonic3 / bitonicLarge,
 * pqt/bitonicSort.cuh:16-78) */
void pqto_bitonic(float *val, uint32_t *idx, uint32_t n);
/* Hillis-Steele block scan (scan_block2 / scan_blockLarge, :112-211) */
void pqto_scan(uint32_t *v, uint32_t n, int inclusive);

/* ---- query chain -------------------------------------------- */
/* pairwise-tree squared L2 over len (power of two) elements, :7146-7160 */
float pqto_seg_dist(const float *q, const float *c, uint32_t len);

/* a2: ProTree::prepareDistSequence, pqt/ProTree.cu:128-207.
 * seq[PQTO_NUM_DISTSEQ]; returns m (d_distCluster); *n_valid = d_numDistSeq */
uint32_t pqto_dist_seq(uint32_t max_cluster, uint32_t p, uint32_t *seq, uint32_t *n_valid);

/* a9: cbDist[(b*c1+a)*LP+lp], computeCBL1L1Dist :1902-1917 + calcDistKernel
 * pqt/ProQuantization.cu:101-137 */
void pqto_cb_dist(const pqto_params *prm, const float *cb1, float *cb_dist);

/* a3 Step A: assign[k1][p] for one query, :7105-7212 */
void pqto_step_a(const pqto_params *prm, const float *cb1, const float *q, uint32_t k1,
                 uint32_t *assign);
/* a4 Step B: lut[LP][c1], :7739-7799 */
void pqto_step_b(const pqto_params *prm, const float *cb1, const float *q, float *lut);
/* a5 Step C: assign_val/assign_idx [p][k1*c2], :1534-1664 */
void pqto_step_c(const pqto_params *prm, const float *cb2, const float *q, uint32_t k1,
                 const uint32_t *assign, float *assign_val, uint32_t *assign_idx);
/* a6 Step D: bins[max_bins] (pre-zeroed by callee), returns nBins, :3374-3549 */
uint32_t pqto_step_d(const pqto_params *prm, uint32_t k1, const uint32_t *dist_seq, uint32_t m,
                     const uint32_t *assign_idx, const uint32_t *bin_counts, uint32_t *bins);
/* a7 Step E1: select_idx[max_vec] (pre-zeroed by callee), returns nVec, :4308-4419 */
uint32_t pqto_step_e1(const pqto_params *prm, const uint32_t *bins, uint32_t n_bins,
                      const uint32_t *bin_prefix, const uint32_t *bin_counts,
                      const uint32_t *db_idx, uint32_t max_vec, uint32_t *select_idx);
/* a8 Step E2: ADC over line codes + bitonic; writes first k of the sorted list, :5189-5351 */
void pqto_step_e2(const pqto_params *prm, const float *lut, const float *cb_dist,
                  const uint32_t *lines, const uint32_t *select_idx, uint32_t n_vec,
                  uint32_t max_vec, uint32_t k, float *out_dist, uint32_t *out_idx);
/* distance of one line-coded vector, same arithmetic as E2 (for tests) */
float pqto_line_adc(const pqto_params *prm, const float *lut, const float *cb_dist,
                    const uint32_t *code);

typedef struct pqto_stages { /* optional per-query intermediates, any pointer may be NULL */
  uint32_t *assign;      /* [QN][k1][p]        */
  float *lut;            /* [QN][LP][c1]       */
  float *assign_val;     /* [QN][p][k1*c2]     */
  uint32_t *assign_idx;  /* [QN][p][k1*c2]     */
  uint32_t *bins;        /* [QN][max_bins]     */
  uint32_t *n_bins;      /* [QN]               */
  uint32_t *select_idx;  /* [QN][max_vec]      */
  uint32_t *n_vec;       /* [QN]               */
} pqto_stages;

/* a1: queryKNN, :8179-8323.  Q is host float[QN][dim]; lines is uint32[N][LP]
 * (lineDescr packed little-endian, pqt/PerturbationProTree.hh:21-25), indexed by
 * original vector id.  Returns 0, or -1 on unsupported shapes.  nthreads<=0: all. */
int pqto_query_knn(const pqto_params *prm, const float *cb1, const float *cb2,
                   const uint32_t *bin_prefix, const uint32_t *bin_counts,
                   const uint32_t *db_idx, const uint32_t *lines, const float *Q, uint32_t QN,
                   uint32_t k, float *out_dist, uint32_t *out_idx, pqto_stages *stages,
                   int nthreads);

/* ---- a11: the 1-B variant queryBIGKNNRerank2 (:8596-8701) ------------------------------ */
#define PQTO_NUM_ANISO_DIR 10   /* pqt/ProTree.hh:12 */
#define PQTO_ANISO_BASE 1.2f    /* pqt/ProTree.hh:13 */
#define PQTO_BIG_KMAX 64        /* getBIGBins2D :3729 kMax           */
#define PQTO_BIG_NINTER 256     /* :3735 nIntermediateBin            */
#define PQTO_BIG_DISTCLUSTER 512 /* prepare2DDistSequence(512), test/test1B.cpp:1215 */

/* prepare2DDistSequence, pqt/ProTree.cu:50-126: seq[NUM_ANISO_DIR * NUM_DISTSEQ] */
void pqto_dist_seq_2d(uint32_t max_cluster, uint32_t *seq);
/* computeSlopeIdx :2839-2862.  *ambiguous is set when logf(slope)/logf(1.2) lies so close
 * to a rounding boundary that the device's logf may round the other way */
uint32_t pqto_slope_idx(const float *val0, const float *val1, uint32_t N, int *ambiguous);
/* getBIGBins2D :3702-3778 = selectBinKernel2D2Parts :2914-3006 + selectBinKernel2DFinal
 * :3012-3188.  assign_val/assign_idx [p][k1*c2] from Step C (k1 = 16).  bins[max_bins]
 * (zeroed by the callee); returns nBins.  k2 = the kVec of the query. */
uint32_t pqto_step_d_big(const pqto_params *prm, uint32_t k1, const uint32_t *seq2d,
                         const float *assign_val, const uint32_t *assign_idx,
                         const uint32_t *bin_counts, uint32_t k2, uint32_t *bins, int *ambiguous);
/* queryBIGKNNRerank2.  prm->k1 (16), prm->max_bins (524288), prm->max_trials (2560) are the
 * literals of :8604-8639 / :3725-3727; candidates per bin are capped at pow2ceil(k) (:6525).
 * ambiguous[QN] (may be NULL): bit 0 = a slope index sits on a logf rounding boundary,
 * bit 1 = the bin walk ran into the end of d_distSeq (the reference reads past its
 * allocation from there on: undefined, not comparable). */
int pqto_query_big_knn_rerank2(const pqto_params *prm, const float *cb1, const float *cb2,
                               const uint32_t *bin_prefix, const uint32_t *bin_counts,
                               const uint32_t *db_idx, const uint32_t *lines, const float *Q,
                               uint32_t QN, uint32_t k, float *out_dist, uint32_t *out_idx,
                               uint32_t *n_bins_out, uint32_t *n_vec_out, int *ambiguous,
                               int nthreads);

/* ---- build side (creates the query path's inputs) ---------- */
/* bin of each DB vector: buildKBestDB :1231-1315 + assignPerturbationBestBinKernel2 :830-942
 * (k1 = 16 there; passed explicitly) */
void pqto_assign_bins(const pqto_params *prm, const float *cb1, const float *cb2, const float *X,
                      uint32_t N, uint32_t k1, uint32_t *bin_of, int nthreads);
/* histogram, exclusive prefix, ids grouped by bin in ascending-id order
 * (countBins :625-661, ProTree::scan pqt/ProTree.cu:1250-1299, sortIdx :715-750) */
void pqto_build_lists(const uint32_t *bin_of, uint32_t N, uint32_t hash_size, uint32_t *bin_counts,
                      uint32_t *bin_prefix, uint32_t *db_idx);
/* line codes: lineDist :7663-7737 + lineClusterKernelFast :7527-7661 */
void pqto_line_encode(const pqto_params *prm, const float *cb1, const float *cb_dist,
                      const float *X, uint32_t N, uint32_t *lines, int nthreads);

/* ---- exact brute force (ground truth for recall) -------------------------------- */
void pqto_brute_force_1nn(const float *X, uint32_t N, const float *Q, uint32_t QN, uint32_t dim,
                          uint32_t *nn, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* PQT_ORACLE_H */
