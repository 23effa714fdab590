/*
 * mvdb_b200.h -- C ABI of the B200-native flat inner-product engine that
 * replaces the faiss-cpu boundary of cnmoro/MiniVectorDB.
 *
 * The reference reaches its hot path through exactly four faiss SWIG calls
 * (reference paths relative to /root/reference):
 *
 *   faiss.IndexFlatIP(d)        minivectordb/vector_database.py:43, 511
 *                               minivectordb/sharded_vector_database.py:80, 639
 *   faiss.normalize_L2(x)       vector_database.py:45, 475 ; sharded_...py:82, 604
 *   index.add(x)                vector_database.py:46, 512 ; sharded_...py:83, 640
 *   index.search(q, k)          vector_database.py:497, 514 ; sharded_...py:626, 642
 *
 * and owns the matrix on the host through np.vstack / np.delete
 * (vector_database.py:72, 107, 126).  Every entry point below names the
 * reference interface it stands in for.  All functions are plain C: opaque
 * handle, raw pointers and sizes, int return code (0 = ok, <0 = error, text
 * from mvdb_last_error()).  No torch / C++ types cross this boundary.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device and
 * fails with MVDB_ERR_CUDA when none is present.
 *
 * Thread-safety: any number of host threads may call mvdb_index_search*
 * concurrently on one index (each call runs on its own stream/workspace);
 * mutating calls (add / remove / compact / reset) are serialised internally
 * and never move rows under a running search (mvdb_index_compact waits for
 * running searches to drain).
 */
#ifndef MVDB_B200_H
#define MVDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVDB_ABI_VERSION 1

enum {
    MVDB_OK = 0,
    MVDB_ERR_ARG = -1,    /* bad argument (dimension mismatch, k <= 0, NULL, row out of range) */
    MVDB_ERR_CUDA = -2,   /* CUDA runtime / driver failure, or no device */
    MVDB_ERR_OOM = -3,    /* device or pinned-host allocation failed */
    MVDB_ERR_STATE = -4   /* operation not valid in the index's current state */
};

/* scan kernel selection (mvdb_index_set_option "scan_variant") */
enum {
    MVDB_SCAN_AUTO = 0,
    MVDB_SCAN_TMA = 1,    /* cp.async.bulk + mbarrier ring, warp-specialised producer */
    MVDB_SCAN_LDG = 2     /* direct 128-bit ld.global.nc loads */
};

typedef struct mvdb_index mvdb_index; /* opaque; one HBM-resident matrix on one device */

/* ---- library ---------------------------------------------------------- */

int mvdb_abi_version(void);
/* Thread-local text of the last error returned to this thread. */
const char* mvdb_last_error(void);
/* Number of visible CUDA devices (0 and MVDB_OK when there is none). */
int mvdb_device_count(int* count);

/* ---- index lifetime ---------------------------------------------------
 * Replaces faiss.IndexFlatIP(d) (vector_database.py:43).  `capacity_hint`
 * rows of address space are reserved up front; physical HBM is mapped in
 * chunks as rows arrive, so row addresses never move while the index grows
 * (0 = default reservation). */
int mvdb_index_create(int d, int device, uint64_t capacity_hint, mvdb_index** out);
/* Mask handles, columns and workspaces made from the index may outlive it: destroying the index releases
 * their device memory and leaves them inert (every use fails with MVDB_ERR_STATE / MVDB_ERR_ARG, their own
 * *_destroy stays legal) -- a garbage-collected binding need not order its finalisers.  An index that belongs
 * to a shard group is refused (MVDB_ERR_STATE): destroy the group first. */
int mvdb_index_destroy(mvdb_index* ix);
/* Drop every row (fresh IndexFlatIP, as _build_index does at vector_database.py:43). */
int mvdb_index_reset(mvdb_index* ix);

/* Tunables; unknown names -> MVDB_ERR_ARG.
 *   "scan_variant"  MVDB_SCAN_*          "fused_k_max"  largest k served by the fused select
 *   "grid_ctas"     CTAs of the scan kernel (0 = one per SM)
 *   "consumer_warps" consumer warps per CTA of the TMA scan (0 = auto)
 *   "large_k_fast"  (default 1) host-buffer searches with fused_k_max < k <= 8192: select the k best by two
 *                   12-bit histogram passes over the score images, one collect and one single-CTA sort (5
 *                   launches per query) instead of the 8-pass radix select (22 launches); a query whose k-th
 *                   score is shared by more than 16384 rows falls back to the radix select.  Same results.
 *   "host_path"     (default 3) bit set for single-query searches with HOST buffers (mvdb_index_search,
 *                   _search_with_mask; k <= fused_k_max): 1 = the kernels write the k results straight
 *                   into pinned host memory (no device-to-host copy), 2 = the query and a per-call filter
 *                   are pulled from pinned host memory by a small grid that the scan is launched behind as
 *                   a programmatic dependent (the scan's start-up and ring fill overlap the PCIe round
 *                   trip; no host-to-device copy).  0 = cudaMemcpyAsync both ways.  Results are identical
 *                   for every value.  With bit 2, a filter that the caller holds in page-locked memory
 *                   (cudaHostAlloc / cudaHostRegister, 16-byte aligned) is pulled from where it lies: no
 *                   staging copy on the host.
 *   "pdl"           (default 0) launch the scans of mvdb_index_search_device / _search_exchange
 *                   with programmatic stream serialization: searches enqueued back to back on
 *                   one stream overlap the serial tail of one (last-CTA merge, cross-GPU
 *                   exchange) with the scan of the next.  Results are unchanged.  Contract while
 *                   on: q_dev / mask_dev of a search must not be written by a KERNEL that
 *                   immediately precedes it on the same stream (copies and events are fine)
 *   "dyn_tiles"     percentage (0..100, default 15) of the tiles the TMA scan claims from a
 *                   global counter instead of the static round-robin split, to level the
 *                   finishing times of the SMs; results are identical for every value
 *   "batch_mode"    large query batches on the tensor cores: 0 off (always the
 *                   fp32 scan), 1 exact (bf16 tcgen05 GEMM selects a rigorous
 *                   candidate superset, survivors re-scored in fp32: same ids and
 *                   distances as the scan; default), 2 bf16 (scores of the bf16 GEMM), 3 tf32
 *                   (scores of a tcgen05 kind::tf32 GEMM over the fp32 matrix itself: no shadow
 *                   copy, 10-bit mantissa operands, half the tensor rate)
 *   "batch_min_nq"  floor on the nq routed to the batched path (default 2; above it a cost model
 *                   picks the cheaper of the fp32 scan passes and one bf16 shadow pass; k <= 128)
 *   "batch_cost_model" 1 (default) = let that cost model decide, 0 = every nq >= batch_min_nq
 *                   goes to the tensor cores (tests)
 *   "coalesce"      1 (default) = concurrent single-query host-buffer calls share one pass over
 *                   the matrix (leader/follower, no added latency when idle); "coalesce_max"
 *                   caps a shared pass (default 64 queries); "coalesce_leaders" 0 (default) =
 *                   one batch in flight for matrices >= 256 MB, two below; 1 / 2 force it;
 *                   "coalesce_wait_pct" (default 40): callers served by one shared pass come back together, so a
 *                   leader that can expect that company (the previous batch had several queries) waits for it
 *                   -- while callers keep arriving, and for at most this % of one pass over the matrix --
 *                   before it starts the next pass (0 = never wait; a lone caller never waits)
 *   "gemm_variant"  tile scheme of the tensor-core batch: 0 one CTA per 128x256 tile, 1 CTA pairs
 *                   (cta_group::2), 2 clusters of 2 sharing the row tile by TMA multicast
 *                   (default), 3 clusters of 4; "gemm_l2_hint" 0/1/2 L2 eviction hints (A/B)
 *   "survivor_tail" (default 1) single-query scans with 32 < k <= 128 keep no per-warp top-k lists: a shared
 *                   threshold (k-th largest of the per-warp best scores) and one global list of the keys that
 *                   pass it, sorted by the last CTA -- same results, a much shorter serial tail; 0 = the
 *                   per-warp selects + merge tree (also the automatic fallback when the list overflows)
 *   "scan_shadow"   (default 0) 1 = single-query searches (k <= 128, d <= 1024, >= 16384 rows) stream an int8
 *                   SHADOW of the matrix (d + 16 bytes per row instead of 4 d; built lazily, kept in step
 *                   with appends) to select a rigorous candidate superset, and re-score the survivors from
 *                   the fp32 rows with the scan's own summation order: ids AND distances are bit-identical
 *                   to the fp32 scan, at ~1/4 of the HBM traffic.  The reference's low-precision precedent
 *                   is its usearch/int8 variant (sharded_vector_database_usearch.py:621-627); this mode is
 *                   exact.  A query whose candidate list overflows is answered by the fp32 scan (an on-device
 *                   conditional launch, no host round trip).
 *   "l2_pin_mb"     experiment: keep the head of the matrix L2-resident across scans (default 0)
 *   test / profiling hooks (see the mvdb_debug_* functions): "trace", "gemm_prof", and
 *   "gemm_debug" (bit mask that switches parts of the GEMM kernel OFF -- results are garbage) */
int mvdb_index_set_option(mvdb_index* ix, const char* name, int64_t value);

/* ---- ingest -------------------------------------------------------------
 * Replaces index.add(x) (vector_database.py:46) and the np.vstack append
 * (vector_database.py:72, 107).  x: n x d float32, row-major, HOST memory.
 * normalize != 0 additionally applies faiss.normalize_L2 semantics to every
 * row on the device while it is written (vector_database.py:45): rows with
 * zero norm are stored unchanged.  *first_row receives the row number of
 * x[0]; rows are numbered densely in arrival order, as faiss does. */
int mvdb_index_add(mvdb_index* ix, const float* x, uint64_t n, int normalize, int64_t* first_row);
/* Same with x in DEVICE memory of the index's device. */
int mvdb_index_add_device(mvdb_index* ix, const float* x_dev, uint64_t n, int normalize,
                          int64_t* first_row);
/* Append n rows of the counter-based synthetic generator (bench / tests;
 * bit-identical to oracle/faiss_flat_ip.c:orc_synth_rows), generated and
 * normalised on the device.  dist: 0 bell, 1 uniform[0,1). */
int mvdb_index_add_synthetic(mvdb_index* ix, uint64_t seed, int64_t row0, uint64_t n, int dist,
                             int normalize, int64_t* first_row);

/* ---- delete -------------------------------------------------------------
 * Replaces np.delete (vector_database.py:126; sharded_...py:229-232).  Rows
 * are tombstoned (never returned again); numbering of the other rows is
 * unchanged until mvdb_index_compact.  Unknown / already deleted rows ->
 * MVDB_ERR_ARG and nothing is changed. */
int mvdb_index_remove_rows(mvdb_index* ix, const int64_t* rows, uint64_t n);
/* Squeeze tombstoned rows out, PRESERVING the order of live rows (so that
 * exact-tie order, which follows row numbers, matches a reference that
 * renumbers on every delete: vector_database.py:138-152).  After it, live
 * row i of the old numbering is row rank(i).  *ntotal_out = new row count. */
int mvdb_index_compact(mvdb_index* ix, int64_t* ntotal_out);

/* ---- introspection ------------------------------------------------------ */
int mvdb_index_dim(const mvdb_index* ix, int* d);
/* ntotal = rows incl. tombstones (faiss Index.ntotal); nlive = rows a search can return. */
int mvdb_index_ntotal(const mvdb_index* ix, int64_t* ntotal, int64_t* nlive);
/* Copy row `row` (as stored, i.e. normalised if it was added so) into out[d]
 * (host).  Stands in for self.embeddings[row] (vector_database.py:55). */
int mvdb_index_reconstruct(mvdb_index* ix, int64_t row, float* out);
/* Copy rows [row0, row0+n) into out[n*d] (host). */
int mvdb_index_reconstruct_n(mvdb_index* ix, int64_t row0, uint64_t n, float* out);
/* Device address / leading dimension (floats) of the resident matrix, and
 * of the live-row bitmask (bit r&31 of word r>>5 set = row r live). */
int mvdb_index_device_view(mvdb_index* ix, const float** matrix_dev, int64_t* ld,
                           const uint32_t** live_dev);

/* ---- search -------------------------------------------------------------
 * Replaces index.search(q, k) (vector_database.py:497) and, with `mask`, the
 * reference's gather-into-a-temporary-index branch (vector_database.py:
 * 508-523): instead of copying admissible rows, the scan applies the filter
 * as a bitmask in its epilogue.
 *
 *   q     nq x d float32 row-major (host)
 *   mask  NULL, or ceil(mask_rows/8) bytes (host): bit (r & 7) of byte (r >> 3)
 *         set = row r admissible (numpy.packbits(..., bitorder="little")).
 *         Rows >= mask_rows (appended after the caller built the mask) are
 *         not admissible.  mask_rows is ignored when mask is NULL.
 *   normalize_queries != 0: apply faiss.normalize_L2 to each query first
 *         (vector_database.py:475)
 *   D     nq x k float32 (host): inner products, best first
 *   I     nq x k int64  (host): row numbers; when fewer than k rows are
 *         admissible the tail is (-FLT_MAX, -1), exactly as faiss pads
 *         (consumed at vector_database.py:500).
 * Order: score descending; exact ties by ascending row number. */
int mvdb_index_search(mvdb_index* ix, const float* q, int64_t nq, int64_t k, const uint8_t* mask,
                      uint64_t mask_rows, int normalize_queries, float* D, int64_t* I);

/* Device-resident filters.  A filter that is used by many queries (the same
 * metadata_filter repeated, vector_database.py:482) can be uploaded ONCE and
 * referenced by handle; searches through a handle move no mask bytes, and
 * concurrent single-query calls that carry handles are coalesced into one
 * tensor-core batch with per-query filters.  Rows >= mask_rows (appended after
 * the handle was made) are not admissible through it.  mvdb_index_compact renumbers rows and
 * therefore invalidates every handle and column made before it. */
typedef struct mvdb_mask mvdb_mask;
int mvdb_index_mask_create(mvdb_index* ix, const uint8_t* mask, uint64_t mask_rows, mvdb_mask** out);
int mvdb_mask_destroy(mvdb_mask* m);
/* mvdb_index_search with the filter given as a handle (NULL = unfiltered). */
int mvdb_index_search_with_mask(mvdb_index* ix, const float* q, int64_t nq, int64_t k, const mvdb_mask* m,
                                int normalize_queries, float* D, int64_t* I);

/* Device-side filter evaluation (the producer of the mask, vector_database.py:354-386).
 * A numeric metadata column is kept in HBM row-aligned with the index (value + presence bit
 * per row); predicates ($gt $gte $lt $lte $ne, equality: op 2 3 4 5 1 0) and the AND / OR /
 * exclude combinators (how 0 / 1 / 2 = and-not) run as small kernels and yield a mask handle
 * directly -- no per-row Python, no mask upload.  Rows lacking the key never match, as in
 * the reference's inverted-index walk (vector_database.py:260).
 * Ordering guarantee: the filter kernels run on one internal stream per index, in call order; a
 * mask handle carries an event recorded behind its latest writer and every search that is given
 * the handle waits for that event on its own stream.  So "predicate -> combine -> search" needs no
 * synchronisation by the caller, from any thread.  A handle must not be combined into / destroyed
 * while another thread is searching with it. */
typedef struct mvdb_column mvdb_column;
int mvdb_column_create(mvdb_index* ix, mvdb_column** out);
int mvdb_column_destroy(mvdb_column* c);
/* rows [len, len+n): values[n] (double), present[n] (0/1 bytes) */
int mvdb_column_append(mvdb_column* c, const double* values, const uint8_t* present, uint64_t n);
int mvdb_mask_from_predicate(mvdb_index* ix, const mvdb_column* c, int op, double operand, mvdb_mask** out);
int mvdb_mask_create_filled(mvdb_index* ix, uint64_t rows, mvdb_mask** out);   /* all rows admissible */
int mvdb_mask_combine(mvdb_mask* dst, const mvdb_mask* src, int how);
int mvdb_mask_count(const mvdb_mask* m, uint64_t* count);                     /* admissible AND live rows */

/* Device-buffer flavour for callers that keep queries/results in HBM (the
 * sharded path and the bench's device-resident leg).  All pointers are device
 * pointers on the index's device; `stream` is a cudaStream_t (NULL = default
 * stream); the call only enqueues work.  mask_dev: ceil(mask_rows/32) uint32
 * words (bits past mask_rows clear) or NULL.  label_offset is added to every returned row number (global
 * numbering of a row shard).  `workspace` must come from
 * mvdb_index_workspace_create and must not be shared by concurrent calls. */
typedef struct mvdb_workspace mvdb_workspace;
int mvdb_index_workspace_create(mvdb_index* ix, mvdb_workspace** out);
int mvdb_index_workspace_destroy(mvdb_workspace* ws);
int mvdb_index_search_device(mvdb_index* ix, mvdb_workspace* ws, const float* q_dev, int64_t nq,
                             int64_t k, const uint32_t* mask_dev, uint64_t mask_rows,
                             int normalize_queries, int64_t label_offset, float* D_dev,
                             int64_t* I_dev, void* stream);

/* ---- helpers ------------------------------------------------------------ */
/* faiss.normalize_L2(x) on host data, in place (vector_database.py:475):
 * copies to `device`, normalises with the ingest kernel, copies back. */
int mvdb_normalize_L2(float* x, uint64_t n, int d, int device);

/* Merge `nparts` per-shard result lists (each nq x k, best first, labels
 * already global, padding (-FLT_MAX,-1)) into the global best-first top-k.
 * D_parts / I_parts are laid out [part][nq][k] in DEVICE memory -- the shape
 * an all-gather of per-GPU results produces.  Enqueues on `stream`. */
int mvdb_merge_topk_device(int device, const float* D_parts, const int64_t* I_parts, int nparts,
                           int64_t nq, int64_t k, float* D_out, int64_t* I_out, void* stream);

/* ---- fused cross-GPU exchange (row-sharded search, one process per GPU) ----
 * No reference counterpart (the reference's search is one in-memory index,
 * sharded_vector_database.py:79-84).  Each rank creates an exchange object,
 * publishes its 64-byte CUDA-IPC handle to the other ranks (any transport),
 * and connects.  mvdb_index_search_exchange then runs scan + exchange + merge
 * in ONE kernel launch per query group: the last CTA of every rank stores its
 * k best (score,row) keys into every peer's receive buffer over NVLink,
 * publishes a sequence number, waits for the peers' and merges.  Every rank
 * ends with the same global (D, I); labels are offsets[rank] + local row.
 * All ranks must issue the same sequence of calls (same nq, k).
 * Limits: k <= k_max <= 128, world <= 16. */
typedef struct mvdb_exchange mvdb_exchange;
int mvdb_exchange_create(int device, int rank, int world, int k_max, int nq_max, mvdb_exchange** out);
int mvdb_exchange_ipc_handle(mvdb_exchange* x, void* handle64);
/* handles: world x 64 bytes in rank order (own entry ignored); offsets: world
 * int64 global row numbers of each rank's row 0. */
int mvdb_exchange_connect(mvdb_exchange* x, const void* handles, const int64_t* offsets);
int mvdb_exchange_set_offsets(mvdb_exchange* x, const int64_t* offsets);
/* Same-process flavour of connect: xs[0..n-1] are the exchange objects of ranks 0..n-1, all created in
 * THIS process (one per device); peers are reached through cudaDeviceEnablePeerAccess, no IPC handles. */
int mvdb_exchange_connect_local(mvdb_exchange* const* xs, int n, const int64_t* offsets);
/* Options: "timeout_ms" (default 2000) -- how long a search's last CTA waits for a peer's top-k
 * before it gives up and raises the status flag. */
int mvdb_exchange_set_option(mvdb_exchange* x, const char* name, int64_t value);
/* *timed_out = 1 if a kernel gave up waiting for a peer (peer died or never launched); the results of
 * that search are undefined.  The flag lives in pinned host memory: valid after any synchronise that
 * covers the search, reading it costs no CUDA call. */
int mvdb_exchange_status(mvdb_exchange* x, int* timed_out);
int mvdb_exchange_destroy(mvdb_exchange* x);
int mvdb_index_search_exchange(mvdb_index* ix, mvdb_workspace* ws, mvdb_exchange* x, const float* q_dev,
                               int64_t nq, int64_t k, const uint32_t* mask_dev, uint64_t mask_rows,
                               int normalize_queries, float* D_dev, int64_t* I_dev, void* stream);

/* ---- shard group: one process, several GPUs ---------------------------------
 * The reference's ShardedVectorDatabase searches ONE in-memory index over all rows
 * (sharded_vector_database.py:79-84, 598-662).  A group is that index spread over the GPUs of one
 * box: shards[i] holds a part of the rows on its own device, and mvdb_group_search answers a query
 * over all of them -- one host thread stages the query (and each shard's filter) to every device,
 * launches every device's scan, the scans' last CTAs exchange the per-shard top-k over NVLink and
 * merge (the same fused kernels as mvdb_index_search_exchange), and the merged list is read back
 * from shard 0 in one transfer.  Returned labels are (shard << 40) | row-of-that-shard; exact score
 * ties are ordered by (shard, row).  masks[i] / host_masks[i] (either may be NULL, per shard or as
 * a whole) filter shard i as in mvdb_index_search_with_mask / mvdb_index_search.  k <= 128.
 * Searches on one group are serialised.  The shards stay usable on their own (add / remove). */
typedef struct mvdb_group mvdb_group;
int mvdb_group_create(mvdb_index* const* shards, int n, mvdb_group** out);
int mvdb_group_destroy(mvdb_group* g);
int mvdb_group_set_option(mvdb_group* g, const char* name, int64_t value);   /* forwarded to every exchange */
int mvdb_group_search(mvdb_group* g, const float* q, int64_t nq, int64_t k, const mvdb_mask* const* masks,
                      const uint8_t* const* host_masks, const uint64_t* host_mask_rows,
                      int normalize_queries, float* D, int64_t* I);

/* Test hook: the raw bf16 tensor-core scores of q[nq,d] against every stored row
 * (out[nq][ntotal], host).  Exists so that the tcgen05 GEMM can be validated
 * against an independent bf16 matmul; not used by the product path. */
int mvdb_debug_gemm_scores(mvdb_index* ix, const float* q, int64_t nq, float* out);

/* Test hook: with option "trace" = 1 the single-query scan kernel stamps %globaltimer at fixed
 * points (slots 0-6: CTA 0, slots 8-13: the last CTA); reads the 16 stamps of the latest launch. */
int mvdb_debug_read_trace(mvdb_index* ix, uint64_t* out16);
/* Test hook: with option "gemm_prof" = 1 the cluster GEMM kernels count, per CTA, the cycles their
 * producer / MMA / epilogue threads spend blocked on each barrier (8 counters per CTA, layout at
 * GemmParams::prof in csrc/gemm_tc.cuh); reads the counters of the most recent launch. */
int mvdb_debug_read_gemm_prof(mvdb_index* ix, uint64_t* out, int ctas);

/* Test hook: counters of the LAST int8 shadow search ("scan_shadow") run on this workspace through
 * mvdb_index_search_device: out4 = {candidates the int8 pass selected, candidates that survived the exact
 * re-scoring, bit 0: the int8 lists overflowed / bit 1: the survivor list of the fp32 survivor-tail scan overflowed
 * (the classic scan answered instead), length of the survivor list of the last survivor-tail scan}. */
int mvdb_debug_read_shadow_counters(mvdb_workspace* ws, uint32_t* out4);

/* Number of kernel launches issued by this library since load (bench.py's
 * "gpu_launches" claim is read from here). */
uint64_t mvdb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MVDB_B200_H */
