/*
 * pqt_b200.h -- C ABI of the B200-native Product-Quantization-Tree query engine
 * (libpqt_b200.so).
 *
 * The reference has no FFI layer: its boundary is the C++ class
 * pqt::PerturbationProTree (pqt/PerturbationProTree.hh:28-235) used directly by
 * tool_query.cpp:92-155 and tool_createdb.cpp:74-114.  Every entry point below
 * names the reference method it replaces (file:line into /root/reference).  The
 * C++ mirror of that class over this ABI lives in
 * product-quantization-tree_b200/host/PerturbationProTree.hh; INTEGRATION.md
 * shows the binding a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every call returns PQT_OK (0) or a
 * negative pqt_status (the reference exit()s instead, utils/helper.hpp:7-15);
 * pqt_last_error() returns the message of the last failing call on that handle.
 * The caller owns every buffer it passes; the handle owns all device memory.
 * One handle = one device + one stream; calls on one handle must be serialised
 * by the caller (like the reference object, which is not re-entrant).
 * There is no CPU fallback: without a CUDA device pqt_create() fails with
 * PQT_ERR_CUDA.
 */
#ifndef PQT_B200_H
#define PQT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQT_ABI_VERSION 1

typedef enum pqt_status {
  PQT_OK = 0,
  PQT_ERR_INVALID = -1,     /* bad argument / unsupported shape            */
  PQT_ERR_STATE = -2,       /* call order (e.g. query before set_db)       */
  PQT_ERR_IO = -3,          /* file could not be read / written            */
  PQT_ERR_CUDA = -4,        /* CUDA runtime error (message has the detail) */
  PQT_ERR_NOMEM = -5        /* device or host allocation failed            */
} pqt_status;

/* id written to result slots beyond the number of candidates found.  The
 * reference leaves stale shared memory there (pqt/PerturbationProTree.cu:
 * 5331-5346); the distance of such a slot is 1e7f exactly as in the reference. */
#define PQT_PAD_IDX 0xFFFFFFFFu

typedef struct pqt_index pqt_index; /* opaque: pqt::PerturbationProTree */

/* The literals the reference hard-codes inside its methods;
 * defaults reproduce them. */
typedef struct pqt_params {
  uint32_t k1;              /* L1 cells expanded per part in queries; 8 (pqt/PerturbationProTree.cu:8187) */
  uint32_t max_bins;        /* 4096 (:8218)                                              */
  uint32_t max_trials;      /* 16 (:3569)                                                */
  uint32_t bin_threads;     /* 1024, probes per trial (:3556)                            */
  uint32_t max_vec_per_bin; /* 2800 (:6208)                                              */
  uint32_t hash_size;       /* HASH_SIZE 400000000 (pqt/PerturbationProTree.hh:12); the
                               tools' --hashsize flag (tool_query.cpp:33)                */
  uint32_t k1_build;        /* 16, L1 cells searched when binning DB vectors (:1237)     */
  uint32_t max_vec;         /* 0 = reference behaviour: candidates re-ranked per query =
                               pow2ceil(k) (:6159).  Non-zero (power of two >= k): fixed
                               candidate budget, results = first k of that ranking
                               (extension; lets k be small without shrinking the scan)  */
  /* literals of the 1-B variant queryBIGKNNRerank2 / getBIGBins2D */
  uint32_t big_k1;          /* 16 (:8604)                                                 */
  uint32_t big_max_bins;    /* 64 * 8192 (:8639)                                          */
  uint32_t big_max_trials;  /* 2560 rounds of 1024 merged bins (:3727)                    */
  uint32_t rank_mode;       /* 0 (default): composite-key sort, groups of bit-equal distances
                               of different vectors re-ordered as the reference's network
                               orders them (same output as 1, faster); 1: the reference's
                               bitonic network on every query (pqt/bitonicSort.cuh:16-78)  */
  uint32_t reserved[4];
} pqt_params;

/* cumulative device-side timings of the query kernels (CUDA events on the
 * handle's stream), filled only while profiling is enabled */
typedef struct pqt_stats {
  uint64_t calls;           /* pqt_query_knn calls measured                 */
  uint64_t queries;         /* queries processed                            */
  uint64_t candidates;      /* line-coded candidates scanned (sum of nVec)  */
  uint64_t kernel_launches; /* kernels launched by the library              */
  double ms_tables;         /* Steps A+B+C (distance-table build)           */
  double ms_bins;           /* Step D+E1 (bin enumeration, candidate list)  */
  double ms_scan;           /* Step E2 ADC scan over line codes             */
  double ms_sort;           /* exact bitonic ranking + top-k emit           */
  double ms_total;          /* first kernel start .. last kernel end        */
  uint64_t scan_launches;   /* launches of the ADC scan kernel              */
  uint64_t exact_rank_queries; /* queries whose ranking needed the exact bitonic network
                                  (tied or >= 1e7 distances); all others use the fast sort */
  uint64_t tie_resolved_queries; /* queries whose groups of bit-equal distances were put into
                                    the network's order by the bit-plane simulation */
  uint64_t stream_scan_launches; /* of scan_launches: launches of the streaming scan kernel
                                    (split pipeline: ms_scan = ADC scan alone, ms_sort = ranking) */
  uint64_t reserved[4];
} pqt_stats;

/* ---- lifetime ------------------------------------------------------------- */

/* PerturbationProTree(uint dim, uint p, uint p2), pqt/PerturbationProTree.hh:37
 * (+ cudaSetDevice, tool_query.cpp:74).  p2 must equal p. */
int pqt_create(uint32_t dim, uint32_t p, uint32_t p2, int device, pqt_index **out);
/* ~PerturbationProTree, pqt/PerturbationProTree.cu:37-58 */
int pqt_destroy(pqt_index *h);
const char *pqt_last_error(const pqt_index *h);
int pqt_abi_version(void);

void pqt_default_params(pqt_params *prm);
int pqt_set_params(pqt_index *h, const pqt_params *prm);
int pqt_get_params(const pqt_index *h, pqt_params *prm);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the
 * handle's own (non-blocking) stream.  The legacy default stream's handle is NULL as
 * well: to run on a default stream pass cudaStreamLegacy / cudaStreamPerThread.  Work the
 * caller enqueues on other streams (NCCL barriers of the multi-GPU path included) is
 * ordered against the library's only through this stream.  The reference uses the
 * default stream throughout. */
int pqt_set_stream(pqt_index *h, void *cuda_stream);

/* ---- index state (load path) ----------------------------------------------- */

/* readTreeFromFile(name), pqt/PerturbationProTree.cu:118-220 (.ppqt: ASCII
 * "dim p p2 c1 c2 nDBs", one skipped byte, float cb1[c1][dim],
 * float cb2[p][c1][c2][dim/p]); overrides dim/p from the file like the reference */
int pqt_read_tree(pqt_index *h, const char *path);
/* writeTreeToFile(name), :60-116 */
int pqt_write_tree(pqt_index *h, const char *path);
/* codebooks straight from host memory (what createTree leaves in
 * d_multiCodeBook / d_multiCodeBook2, :274-303) */
int pqt_set_tree(pqt_index *h, uint32_t c1, uint32_t c2, const float *cb1, const float *cb2);
int pqt_get_tree_shape(const pqt_index *h, uint32_t *dim, uint32_t *p, uint32_t *c1, uint32_t *c2);
int pqt_get_tree(const pqt_index *h, float *cb1, float *cb2);

/* setDB(N, prefix, counts, dbIdx), :1184-1229: HOST pointers, copied;
 * prefix/counts have params.hash_size entries, dbIdx has N. */
int pqt_set_db(pqt_index *h, uint32_t N, const uint32_t *prefix, const uint32_t *counts,
               const uint32_t *db_idx);
/* The line codes the reference keeps in d_lineLambda (lineDist :7663-7737 /
 * prepareEmptyLambda pqt/PerturbationProTree.hh:103 + upload, test/test1B.cpp:
 * 1181-1233): HOST lineDescr[N][LP] (4 bytes each, indexed by vector id).  Sets
 * d_lineParts = LP.  Must follow pqt_set_db (codes are re-laid in bin order). */
int pqt_set_lines(pqt_index *h, const uint32_t *lines, uint32_t N, uint32_t line_parts);

/* ---- query (hot path) -------------------------------------------------------- */

/* queryKNN(resIdx, resDist, Q, QN, k), :8179-8323.  Q: float[QN][dim], device
 * pointer if q_on_device (as in the reference) else host.  idx/dist: [QN][k],
 * device pointers if out_on_device else host.  Ascending by distance. */
int pqt_query_knn(pqt_index *h, const float *Q, int q_on_device, uint32_t QN, uint32_t k,
                  uint32_t *idx, float *dist, int out_on_device);

/* queryBIGKNNRerank2(resIdx, resDist, Q, QN, k, hLines), :8596-8701 -- the 1-B variant:
 * k1 = big_k1, bins from the 2-D anisotropic merge of parts (0,1) and (2,3) (getBIGBins2D
 * :3702-3778, needs p == 4), at most two vectors counted per bin while collecting, every
 * listed bin contributes up to pow2ceil(k) candidates (:6525), same ADC + ranking.  The
 * reference fetches the line codes from pinned host memory (hLines); here they are the
 * resident codes of pqt_set_lines / pqt_line_dist.  prepare2DDistSequence(512)
 * (pqt/ProTree.cu:50-126, test/test1B.cpp:1215) is done on first use. */
int pqt_query_big_knn_rerank2(pqt_index *h, const float *Q, int q_on_device, uint32_t QN,
                              uint32_t k, uint32_t *idx, float *dist, int out_on_device);

/* ---- build side (creates the query path's inputs) ----------------------------- */

/* buildKBestDB(A, N), :1231-1315 -- bins every vector (k1_build L1 cells), builds
 * counts / exclusive prefix / ids grouped by bin (ascending id inside a bin) on
 * the device and installs them as the DB (same state as pqt_set_db).
 * X: float[N][dim], device pointer if x_on_device. */
int pqt_build_kbest_db(pqt_index *h, const float *X, int x_on_device, uint32_t N);
/* lineDist(DB, N), :7663-7737 with an explicit LP (the reference hard-sets 16):
 * encodes and installs the line codes (same state as pqt_set_lines). */
int pqt_line_dist(pqt_index *h, const float *X, int x_on_device, uint32_t N, uint32_t line_parts);
/* ---- chunked build: the 1-B path (test/test1B.cpp:783-871) ------------------------------
 * The reference's 1-B driver walks the base set in chunks of 10 M vectors: buildKBestDB on a
 * chunk, append the chunk's bins behind the existing bin content with the ids offset by the
 * chunk start (test/test1B.cpp:816-852), and lineDist per chunk.  The merged lists are the
 * ids grouped by bin in ascending-id order, i.e. what one buildKBestDB over all N vectors
 * gives; the entry points below produce exactly that without ever holding more than one
 * chunk of raw vectors:
 *   pass 1   pqt_assign_bins(chunk) -> bins of the chunk (any chunking, any rank)
 *            pqt_set_db_from_bins(bin_of[N])       counts / prefix / dbIdx, installs the DB
 *   pass 2   pqt_line_dist_begin(N, LP); pqt_line_dist_chunk(chunk, id0) ...; pqt_line_dist_end
 * Rows are float (as the reference passes them) or the uint8 payload of a .umem file
 * (utils/filereader.hpp:40-47 widens it on the host; here the kernels do).
 * On a sharded handle (pqt_set_shard before pqt_set_db_from_bins) pass 2 encodes and keeps
 * only the vectors of the own bin-range slice, so 8 ranks build a 1-B index without any of
 * them holding the whole code array. */
#define PQT_X_F32 0
#define PQT_X_U8 1
/* bins of n vectors (assignPerturbationBestBinKernel2 :830-942 after Step A with k1_build
 * cells); bin_out: uint32[n], device pointer if out_on_device */
int pqt_assign_bins(pqt_index *h, const void *X, int x_kind, int x_on_device, uint32_t n,
                    uint32_t *bin_out, int out_on_device);
/* countBins + scan + sortIdx (:1263-1300) over the bins of all N vectors (ascending id inside
 * a bin); same state as pqt_set_db afterwards */
int pqt_set_db_from_bins(pqt_index *h, const uint32_t *bin_of, int on_device, uint32_t N);
int pqt_line_dist_begin(pqt_index *h, uint32_t N, uint32_t line_parts);
/* rows = vectors id0 .. id0+n-1.  lines_out (HOST, may be NULL, unsharded handles only):
 * lineDescr[n][LP] of the chunk in id order = the chunk's slice of the .lines file */
int pqt_line_dist_chunk(pqt_index *h, const void *X, int x_kind, int x_on_device, uint32_t id0,
                        uint32_t n, uint32_t *lines_out);
int pqt_line_dist_end(pqt_index *h);

/* getBinPrefix/getBinCounts/getDBIdx/getLine, pqt/PerturbationProTree.hh:97-101,
 * as host copies (any pointer may be NULL) */
int pqt_get_db(const pqt_index *h, uint32_t *prefix, uint32_t *counts, uint32_t *db_idx);
int pqt_get_lines(const pqt_index *h, uint32_t *lines);
/* the resident line codes as stored: rows pos0 .. pos0+n-1 of this handle's slice of the
 * BIN-ORDERED list (row r holds vector dbIdx[pos_lo + r]); HOST lineDescr[n][LP].  Unlike
 * pqt_get_lines it needs no second copy on the device (1-B indexes). */
int pqt_get_codes_binorder(const pqt_index *h, uint64_t pos0, uint64_t n, uint32_t *codes);
int pqt_get_db_size(const pqt_index *h, uint32_t *N, uint32_t *line_parts);

/* ---- compact index file (no reference counterpart) ---------------------------------
 * The reference's index files are two dense HASH_SIZE arrays (.prefix / .count, 3.2 GB at
 * 4e8 bins), .dbIdx and the line codes by vector id (tool_createdb.cpp:105-114,
 * test/test1B.cpp:1181-1233); loading them rebuilds the directory and re-orders the codes.
 * pqt_save_index writes what the handle holds -- bitmap, rank directory, compact prefix,
 * dbIdx and the line codes in bin order (of a sharded handle: its own slice) -- and
 * pqt_load_index restores it: header checks against the handle's tree, hash_size and shard,
 * then plain uploads (ids, centroid numbers and bin lists are range-checked on the device).
 * pqt_get_db / pqt_get_lines still expand the reference's arrays from a loaded index. */
int pqt_save_index(const pqt_index *h, const char *path);
int pqt_load_index(pqt_index *h, const char *path);

/* ---- multi-GPU: bin-range shards ----------------------------------------------
 * The reference is single-GPU; its 1-B mode keeps the line codes in pinned host memory
 * (test/test1B.cpp:1121-1192).  Here the bin-ordered code array is cut into `world` contiguous
 * slices of equal vector counts, one per GPU (one process per GPU); the bin directory, the ids
 * and the codebooks are replicated.  Rank r owns the queries [r*q_per_rank, (r+1)*q_per_rank)
 * of a batch (Steps A-E1 and the ranking) and its slice of the codes (the ADC scan of every
 * query's candidates that live there).  Per batch, on every rank:
 *   1. pqt_shard_dispatch   Steps A-E1 for the own queries; Step B (the LUT) for ALL queries
 *                           (cheaper than shipping 4 KB per query); then every shard is sent
 *                           the own queries' candidates that live in its slice: (position,
 *                           entry) pairs stored into the shard's inbox -- peer memory over
 *                           NVLink, 8 bytes per candidate, repeats of a code row sent once
 *   2. a cross-rank barrier on the stream (caller: any tiny NCCL collective)
 *   3. pqt_shard_scan_p2p   streaming ADC scan of the own inbox; every distance is stored
 *                           straight into the distance array of the rank that owns the query
 *                           (peer memory, 4 bytes per candidate): the scan and the all-to-all
 *                           of its results are one kernel
 *   4. a cross-rank barrier
 *   5. pqt_shard_rank       ranking + first-k emit of the own queries
 * The result is bit-identical to pqt_query_knn on the unsharded index (every candidate has
 * exactly one evaluator and one slot).  The exchange buffers are owned by the handle; peers map
 * them through CUDA IPC handles (or raw pointers inside one process). */

/* Keep only the line codes whose position in the bin-ordered list lies in this rank's slice
 * [rank*N/world, (rank+1)*N/world).  Either before the DB is installed (pqt_set_db /
 * pqt_set_db_from_bins: the slice is all that is uploaded / encoded) or after the index is
 * complete (the resident codes are trimmed in place). */
int pqt_set_shard(pqt_index *h, uint32_t rank, uint32_t world);
/* own buffers: distances [q_per_rank][max_vec], inbox [q_per_rank*world][max_vec] x 8 bytes,
 * inbox row lengths [q_per_rank*world]; call after pqt_set_shard */
int pqt_shard_exchange_alloc(pqt_index *h, uint32_t q_per_rank, uint32_t max_vec);
/* 192 bytes: cudaIpcMemHandle_t of the distance buffer, of the inbox, of the row lengths */
int pqt_shard_exchange_handle(pqt_index *h, void *handle192);
/* handles: world * 192 bytes, entry r from rank r (the own entry is ignored) */
int pqt_shard_exchange_open(pqt_index *h, uint32_t world, const void *handles);
/* same-process alternative (tests): raw device pointers of every rank's buffers */
int pqt_shard_exchange_set_peers(pqt_index *h, uint32_t world, void *const *val_ptrs,
                                 void *const *inbox_ptrs, void *const *cnt_ptrs);
int pqt_shard_exchange_ptrs(pqt_index *h, void **val_ptr, void **inbox_ptr, void **cnt_ptr);
/* Q: all QN queries of the batch (device or host); the own queries are [q_lo, q_hi) */
int pqt_shard_dispatch(pqt_index *h, const float *Q, int q_on_device, uint32_t QN, uint32_t k,
                       uint32_t q_lo, uint32_t q_hi);
int pqt_shard_scan_p2p(pqt_index *h, uint32_t QN, uint32_t k);
/* idx/dist: [q_own][k] results of the own queries, device or host */
int pqt_shard_rank(pqt_index *h, uint32_t q_own, uint32_t k, uint32_t *idx, float *dist,
                   int out_on_device);

/* pow2ceil(k) or params.max_vec: row length of the candidate arrays */
int pqt_candidate_width(const pqt_index *h, uint32_t k, uint32_t *max_vec);

/* ---- measurement / introspection ------------------------------------------------ */

int pqt_profile_enable(pqt_index *h, int on); /* per-kernel CUDA-event timing */
int pqt_get_stats(const pqt_index *h, pqt_stats *st);
int pqt_reset_stats(pqt_index *h);

/* Per-query intermediates of the last pqt_query_knn call,
 * recorded only while pqt_debug_enable(h, 1); copied to HOST buffers. */
typedef enum pqt_stage {
  PQT_STAGE_ASSIGN = 0,     /* uint32 [QN][k1][p]        Step A  */
  PQT_STAGE_LUT = 1,        /* float  [QN][LP][c1]       Step B  */
  PQT_STAGE_ASSIGN_VAL = 2, /* float  [QN][p][k1*c2]     Step C  */
  PQT_STAGE_ASSIGN_IDX = 3, /* uint32 [QN][p][k1*c2]     Step C  */
  PQT_STAGE_BINS = 4,       /* uint32 [QN][max_bins]     Step D  */
  PQT_STAGE_NBINS = 5,      /* uint32 [QN]               Step D  */
  PQT_STAGE_SELECT_IDX = 6, /* uint32 [QN][max_vec]      Step E1 */
  PQT_STAGE_NVEC = 7,       /* uint32 [QN]               Step E1 */
  PQT_STAGE_CB_DIST = 8,    /* float  [c1][c1][LP]       computeCBL1L1Dist :1902-1917 */
  PQT_STAGE_DIST_SEQ = 9,   /* uint32 [65536]            prepareDistSequence pqt/ProTree.cu:128-207 */
  PQT_STAGE_DIST_SEQ_2D = 10, /* uint32 [10][65536]      prepare2DDistSequence pqt/ProTree.cu:50-126 */
  PQT_STAGE_BIG_BINS = 11,  /* uint32 [QN][candidate width] bins listed by getBIGBins2D (last BIG query) */
  PQT_STAGE_BIG_NBINS = 12, /* uint32 [QN] */
  PQT_STAGE_RERANK_PHASES = 13 /* uint64 [QN][8] clock64 stamps of the fused scan+rank kernel per
                                  query: start, LUT ready, scan done, repair done, sort + emit done,
                                  ties done, query done, (SM << 32 | nVec << 1 | fast) */
} pqt_stage;
int pqt_debug_enable(pqt_index *h, int on);
int pqt_debug_stage(const pqt_index *h, int stage, void *host_out, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PQT_B200_H */
