// scan.cuh -- the hot path: flat inner-product scan with fused top-k.
//
// Replaces faiss IndexFlatIP.search for small query batches
// (exhaustive_inner_product_seq + HeapBlockResultHandler; call sites
// ref minivectordb/vector_database.py:497, 514 and
// minivectordb/sharded_vector_database.py:626, 642) and, through the
// admissible-row bitmask applied in the epilogue, the reference's
// gather-into-a-temporary-index filtered branch (vector_database.py:508-523).
//
// Work decomposition (HBM-bound: N*ld*4 bytes streamed once per query group):
//   * the matrix is cut into tiles of 8 consecutive rows (= one byte of the
//     admissible/live bitmasks); tile t belongs to CTA t % gridDim.x;
//   * TMA variant: warp 0 is the producer -- one lane issues one
//     cp.async.bulk (global -> shared, L2 evict_first) per tile into a ring of
//     S stages guarded by full/empty mbarriers; consumer warp w owns every
//     ncw-th tile of the CTA, so no CTA-wide barrier exists in steady state;
//   * LDG variant: every warp is a consumer and loads its tile straight from
//     global memory with 128-bit non-allocating loads;
//   * a consumer lane owns float4 chunk `lane + 32 j` of each row, accumulates
//     8 rows x NQ queries, reduce8() folds the 8 partials across the warp in 9
//     shuffles, and quad leaders push (score,row) keys into the WarpSelect;
//   * epilogue: warps merge inside the CTA, the CTA writes its k best keys,
//     the LAST CTA to finish (ticket counter) merges all partial lists and
//     writes (D, I) -- one launch per search.
#pragma once
#include <cfloat>

#include "select.cuh"

namespace mvdb {

struct ScanParams {
    const float* x;        // matrix, row-major, leading dimension ld (floats, multiple of 4)
    const float* q;        // queries, dense [nq][d]
    const uint32_t* live;  // live-row bitmask or nullptr (no tombstones)
    const uint32_t* mask;  // admissible-row bitmask or nullptr (no filter)
    uint64_t* partials;    // [nq][gridDim.x][k] keys
    unsigned int* ticket;  // zero before launch; reset by the last CTA
    unsigned int* tile_ctr;  // dynamic tile scheduler (NULL = static round-robin); zero before launch, reset by the last CTA
    uint32_t pdl_early;      // trigger the dependent launch at kernel entry (only when one CTA fills an SM)
    uint32_t dep_inputs;     // q and mask are being written by the grid this launch programmatically depends on
                             // (the host path's staging pull): wait for it before the first read of either
    uint32_t static_iters;   // ... iterations of every CTA served by the static round-robin first
    uint32_t dyn_tile0;      // ... first tile of the dynamically claimed remainder (= static_iters * grid, multiple of kDynChunk)
    float* outD;           // [nq][k]
    int64_t* outI;         // [nq][k]
    uint32_t* all_ord;     // large-k mode: [nq][n] score images (0 = not admissible); else nullptr
    int64_t label_offset;
    uint32_t n;            // rows to scan (snapshot of ntotal)
    int d;                 // logical dimension
    int ld4;               // ld / 4
    int nq;                // queries in this launch (== NQ template for the multi kernel)
    int k;
    int cap;               // WarpSelect capacity (select_cap(k))
    int normalize_q;
    int stages;            // TMA ring depth
    uint32_t stage_bytes;  // 8 * ld * 4 rounded to 128
    uint32_t sel_off, q_off, stage_off;  // byte offsets into dynamic shared memory
    uint32_t merge_off, merge_bytes;     // scratch for the last-CTA merge (the idle TMA ring, or a tail region)
    unsigned long long* trace;   // debug: globaltimer stamps of one consumer warp (nullptr in production)
    uint32_t pin_tiles;          // tiles [0, pin_tiles) are loaded with L2 evict_last: the head of the matrix stays
                                 // L2-resident across queries, the rest streams through with evict_first
    const uint32_t* qmask[8];    // per-query admissible bitmasks (coalesced searches; nullptr = none); multi kernel only
    uint32_t qmask_bytes[8];     // bytes available behind qmask[i]; rows past them are not admissible
    int has_qmask;
    const unsigned int* run_if;  // not null: the whole launch is a no-op unless *run_if != 0 (fallback behind the int8 / survivor modes)
    // survivor mode (scan_q1_kernel<.., kSurv = true>, 32 < k <= 128): no per-warp selects and no merge tree --
    // a shared threshold (k-th largest of the per-warp best scores, see select.cuh) and ONE global list of the keys
    // that pass it, sorted by the last CTA
    uint64_t* surv;              // [surv_cap] keys
    struct SurvCtl* sctl;
    unsigned int* best;          // [nbest] ordered images of the per-warp (k <= 32: per-CTA) best scores, zero before launch
    uint32_t nbest, surv_cap;
    unsigned int* ovf_host;      // pinned host word raised when the list overflows (the caller re-runs the classic scan), or nullptr
    const struct XchgDev* xchg;  // fused cross-GPU exchange (nullptr = single GPU)
    uint64_t xchg_seq;           // sequence number of this launch (same on every rank, > 0)
};

// ---------------------------------------------------------------------------
// Fused cross-GPU exchange (sharded search).  Every rank owns a receive buffer
// that all peers can write over NVLink (peer-mapped through CUDA IPC):
//   recv [2 parities][world][nq_max * k_max] keys,  flags [2][world] sequence numbers.
// The last CTA of rank r stores its final per-query lists into slot r of EVERY
// rank's buffer (plain st.global on peer pointers), fences system-wide, then
// publishes the launch's sequence number in every rank's flag r.  It then waits
// until all world flags of its own buffer carry that number and merges the
// world x k candidates -- scan, exchange and merge in ONE launch, no NCCL call.
// Parity double-buffering suffices: a rank can finish query s only after every
// rank has sent s, so nobody can be two queries ahead of a reader.
// ---------------------------------------------------------------------------
constexpr int kMaxWorld = 16;
struct XchgDev {
    uint64_t* recv[kMaxWorld];    // recv[p]  = rank p's receive buffer (peer-mapped)
    uint64_t* flags[kMaxWorld];   // flags[p] = rank p's flag array (peer-mapped)
    int64_t offsets[kMaxWorld];   // global row number of rank p's row 0
    int world, rank, k_max, nq_max;
    unsigned int* status;         // pinned host memory (mapped): set to 1 if a wait timed out -- the host reads it
                                  // after every synchronise without a CUDA call
    uint64_t timeout_ns;          // how long the last CTA waits for a peer's flag before giving up
};

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// send query qi's final list (keys[0..cnt), best first) to slot `rank` of every rank
__device__ __forceinline__ void xchg_send(const XchgDev* x, uint64_t seq, int qi, const uint64_t* keys,
                                          int cnt, int k, int lane) {
    const size_t slot = (size_t(seq & 1) * x->world + x->rank) * size_t(x->nq_max) * x->k_max +
                        size_t(qi) * x->k_max;
    for (int i = lane; i < k; i += kWarp) {
        const uint64_t key = (i < cnt) ? keys[i] : kEmptyKey;
        for (int p = 0; p < x->world; p++) x->recv[p][slot + i] = key;
    }
    __threadfence_system();
}

// merge the world lists of query qi that peers deposited in MY buffer
__device__ __forceinline__ void xchg_merge(const XchgDev* x, uint64_t seq, int qi, uint64_t* buf, int cap,
                                           int k, float* D, int64_t* I, int lane) {
    const uint64_t* mine = x->recv[x->rank] + size_t(seq & 1) * x->world * size_t(x->nq_max) * x->k_max +
                           size_t(qi) * x->k_max;
    const size_t src_stride = size_t(x->nq_max) * x->k_max;
    const int total = x->world * k;
    if (k <= kExtractMaxK && total <= 256) {
        // small k: the world x k candidates sit in registers (<= 8 per lane) and the k best are pulled out
        // by warp-wide arg-max rounds -- no shared memory, no sort
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int pos = lane + 32 * u;
            uint64_t key = kEmptyKey;
            if (pos < total) {
                const uint64_t orig = __ldcg(mine + size_t(pos / k) * src_stride + (pos % k));
                if (orig != kEmptyKey) key = (orig & 0xFFFFFFFF00000000ull) | uint64_t(0xFFFFFFFFu - uint32_t(pos));
            }
            v[u] = key;
        }
        const uint64_t key = warp_extract_topk<8>(v, k, lane);
        if (lane < k) {
            if (key == kEmptyKey) {
                D[lane] = -FLT_MAX;
                I[lane] = -1;
            } else {
                const int pos = int(key_row(key));
                const int src = pos / k;
                const uint64_t orig = __ldcg(mine + size_t(src) * src_stride + (pos % k));
                D[lane] = key_score(orig);
                I[lane] = int64_t(key_row(orig)) + x->offsets[src];
            }
        }
        return;
    }
    WarpSelect f;
    f.init(buf, cap, k);
    // candidate position pos = src * k + i encodes the tie order (rank, then in-list order
    // = ascending global row for contiguous row shards)
    for (int base = 0; base < total; base += kWarp) {
        int pos = base + lane;
        uint64_t key = kEmptyKey;
        if (pos < total) {
            uint64_t orig = __ldcg(mine + size_t(pos / k) * src_stride + (pos % k));
            if (orig != kEmptyKey) key = (orig & 0xFFFFFFFF00000000ull) | uint64_t(0xFFFFFFFFu - uint32_t(pos));
        }
        f.push(pos < total, key, lane);
    }
    f.compact(lane);
    __syncwarp();
    for (int i = lane; i < k; i += kWarp) {
        uint64_t key = (i < f.cnt) ? f.buf[i] : kEmptyKey;
        if (key == kEmptyKey) {
            D[i] = -FLT_MAX;
            I[i] = -1;
        } else {
            int pos = int(key_row(key));
            int src = pos / k;
            uint64_t orig = __ldcg(mine + size_t(src) * src_stride + (pos % k));
            D[i] = key_score(orig);
            I[i] = int64_t(key_row(orig)) + x->offsets[src];
        }
    }
}

// flags: publish `seq` to every rank, then wait for every rank's `seq` in my own flags.
// Called by one warp.  Gives up after x->timeout_ns (a peer died or never launched) and raises
// x->status; the results of that search are then garbage and the host raises an error.
__device__ __forceinline__ void xchg_publish_and_wait(const XchgDev* x, uint64_t seq, int lane) {
    const int par = int(seq & 1);
    if (lane < x->world) st_release_sys(x->flags[lane] + par * x->world + x->rank, seq);
    if (lane < x->world) {
        const uint64_t* f = x->flags[x->rank] + par * x->world + lane;
        const uint64_t t0 = global_timer_ns();
        while (ld_acquire_sys(f) < seq) {
            __nanosleep(200);
            if (global_timer_ns() - t0 > x->timeout_ns) {
                *reinterpret_cast<volatile unsigned int*>(x->status) = 1u;
                __threadfence_system();
                break;
            }
        }
    }
    __syncwarp();
}

struct SurvCtl {                 // zero before the first launch; every launch leaves it clean
    unsigned int count, ticket, overflow, last_count;
};

// shared-memory header (first 1024 bytes)
struct SmemHeader {
    uint64_t full[16];
    uint64_t empty[16];
    int cnts[64];      // [warp][query] list lengths for the CTA merge
    int last_flag;
    uint32_t tile_of[16];  // dynamic tile scheduler: tile held by ring stage s (kNoTile = stop)
    uint32_t adm_of[16];   // ... and its admissible byte (mask & live)
    unsigned int surv_thr, surv_refreshes;   // survivor mode: CTA-wide threshold (ordered image) and how often it was recomputed
    unsigned int tail_ns;                    // survivor mode, last CTA: length of the survivor list (read once, by the thread that resets it)
};
static_assert(sizeof(SmemHeader) <= 1024, "header too large");

// ---------------------------------------------------------------------------
// epilogue shared by all scan kernels
// ---------------------------------------------------------------------------
__device__ __forceinline__ void write_results(const uint64_t* keys, int cnt, int k, float* D,
                                              int64_t* I, int64_t label_offset, int lane) {
    for (int i = lane; i < k; i += kWarp) {
        uint64_t key = (i < cnt) ? keys[i] : kEmptyKey;
        if (key == kEmptyKey) {
            D[i] = -FLT_MAX;  // faiss pads IP results with the lowest float and id -1
            I[i] = -1;
        } else {
            D[i] = key_score(key);
            I[i] = int64_t(key_row(key)) + label_offset;
        }
    }
}

// Merge the per-warp lists of query qi (already compacted, lengths in
// hdr->cnts) into warp `cw`'s list.  Returns the merged select state.
__device__ __forceinline__ WarpSelect merge_cta_lists(SmemHeader* hdr, uint64_t* selbuf, int nq,
                                                      int qi, int cw, int ncw, int cap, int k,
                                                      int lane) {
    WarpSelect m;
    m.init(selbuf + size_t(cw * nq + qi) * cap, cap, k);
    m.cnt = hdr->cnts[cw * nq + qi];
    m.thr = (m.cnt == k) ? m.buf[k - 1] : kEmptyKey;
    for (int w = 0; w < ncw; w++) {
        if (w == cw) continue;
        m.push_array(selbuf + size_t(w * nq + qi) * cap, hdr->cnts[w * nq + qi], lane);
    }
    m.compact(lane);
    return m;
}

// Called by all consumer warps once their tiles are done.  `sel_cnt[qi]` must
// already be compacted lists in selbuf[(cw*nq+qi)*cap ..].
// bar_id/bar_threads: named barrier covering exactly the consumer warps.
// debug timeline: consumer warp 0 lane 0 of CTA 0 stamps slots 0-7, of the LAST CTA slots 8-15
__device__ __forceinline__ void trace_stamp(const ScanParams& p, int slot, int cw, int lane) {
    if (p.trace && cw == 0 && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[slot] = t;
    }
}

// Programmatic dependent launch (opt-in, option "pdl"): when two searches are enqueued back to
// back on one stream, the second starts scanning as soon as every CTA of the first has finished
// ITS scan, so the first one's serial tail (CTA merges, last-CTA merge, cross-GPU exchange) runs
// under the second one's scan.  Everything a scan shares with its predecessor -- partial lists,
// ticket, result buffers, exchange slots -- is touched only AFTER pdl_wait(), which returns once
// the previous grid has completed and its writes are visible.  The trigger comes after the wait,
// so grid N+2 cannot start before grid N is complete: at most two grids are in flight, and the
// only state used before the wait (the dynamic tile counter) alternates between two slots.
// When a CTA needs more than half an SM's shared memory (pdl_early; d >= ~256) the trigger moves
// to kernel entry: a CTA of N+1 can then only be placed on an SM that a CTA of N has LEFT, N+2
// only where N+1 has left -- i.e. after N completed -- so the same two-grid bound holds and N+1's
// scan also overlaps N's straggling scans and the launch latency (100 k x 512: 56 -> 47.5 us).
// With smaller CTAs an entry trigger lets N+2 run next to N+1 while N still runs, and N+2 then
// claims tiles from the counter N is about to reset (caught by
// test_programmatic_dependent_launch_is_result_neutral); those launches trigger after the wait.
// Both instructions are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Small k (<= 16, select buffers of 64 keys): the merge tails run on warp_extract_topk instead of
// sort-and-cut.  `fast_tail` must give the same answer everywhere in a launch.
__device__ __forceinline__ bool fast_tail(const ScanParams& p, int ncw) {
    return p.k <= kExtractMaxK && p.cap == 64 && ncw <= 8 &&
           uint32_t(gridDim.x) * uint32_t(p.k) <= uint32_t(ncw) * 32u * 16u;
}

// the per-warp select buffers of query qi (unsorted, lengths in hdr->cnts) -> the CTA's k best, lane j = j-th
template <int M>
__device__ __forceinline__ uint64_t cta_extract(const SmemHeader* hdr, const uint64_t* selbuf, int nq, int qi, int ncw,
                                                int cap, int k, int lane) {
    uint64_t v[M];
#pragma unroll
    for (int u = 0; u < M; u++) {
        const int w = u >> 1, i = lane + 32 * (u & 1);
        v[u] = (w < ncw && i < hdr->cnts[w * nq + qi]) ? selbuf[size_t(w * nq + qi) * cap + i] : kEmptyKey;
    }
    return warp_extract_topk<M>(v, k, lane);
}

// `total` keys at src (global memory, other SMs wrote them) split over the CTA's consumer threads -> this
// warp's k best; all loads are issued before the first use (one L2 round trip)
template <int M>
__device__ __forceinline__ uint64_t spread_extract(const uint64_t* src, int total, int t, int nthr, int k, int lane) {
    uint64_t v[M];
#pragma unroll
    for (int u = 0; u < M; u++) {
        const int i = t + u * nthr;
        v[u] = (i < total) ? __ldcg(src + i) : kEmptyKey;
    }
    return warp_extract_topk<M>(v, k, lane);
}

__device__ __forceinline__ void finish_scan_fast(const ScanParams& p, SmemHeader* hdr, uint64_t* selbuf, int cw, int ncw,
                                                 int lane, int bar_id, int bar_threads) {
    const int nq = p.nq, k = p.k, cap = p.cap;
    const int G = gridDim.x;
    // ---- CTA merge: warp (qi % ncw) pulls the k best of query qi out of all warps' buffers ----
    for (int qi = cw; qi < nq; qi += ncw) {
        const uint64_t key = (ncw <= 4) ? cta_extract<8>(hdr, selbuf, nq, qi, ncw, cap, k, lane)
                                        : cta_extract<16>(hdr, selbuf, nq, qi, ncw, cap, k, lane);
        if (lane < k) p.partials[(size_t(qi) * G + blockIdx.x) * k + lane] = key;
    }
    if (blockIdx.x == 0) trace_stamp(p, 4, cw, lane);
    __threadfence();
    named_bar_sync(bar_id, bar_threads);
    if (blockIdx.x == 0) trace_stamp(p, 5, cw, lane);
    if (cw == 0 && lane == 0) {
        unsigned t = atomicAdd(p.ticket, 1u);
        hdr->last_flag = (t == unsigned(G - 1));
    }
    named_bar_sync(bar_id, bar_threads);
    if (blockIdx.x == 0) trace_stamp(p, 6, cw, lane);
    if (!hdr->last_flag) return;
    // ---- last CTA: G x k keys per query, every warp takes a slice, then one warp per query finishes ----
    trace_stamp(p, 8, cw, lane);
    __threadfence();
    const int total = G * k, t = cw * kWarp + lane;
    const int per = (total + bar_threads - 1) / bar_threads;
    for (int qi = 0; qi < nq; qi++) {
        const uint64_t* src = p.partials + size_t(qi) * G * k;
        const uint64_t key = per <= 4    ? spread_extract<4>(src, total, t, bar_threads, k, lane)
                             : per <= 8  ? spread_extract<8>(src, total, t, bar_threads, k, lane)
                                         : spread_extract<16>(src, total, t, bar_threads, k, lane);
        if (lane < k) selbuf[size_t(cw * nq + qi) * cap + lane] = key;   // the warp's own buffer: its candidates are spent
        if (qi == 0) trace_stamp(p, 10, cw, lane);
    }
    named_bar_sync(bar_id, bar_threads);
    trace_stamp(p, 11, cw, lane);
    for (int qi = cw; qi < nq; qi += ncw) {
        uint64_t v[4];   // ncw * k <= 8 * 16 keys
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = lane + 32 * u, w = i / k;
            v[u] = (w < ncw) ? selbuf[size_t(w * nq + qi) * cap + (i - w * k)] : kEmptyKey;
        }
        const uint64_t key = warp_extract_topk<4>(v, k, lane);
        uint64_t* fin = selbuf + size_t(cw * nq + qi) * cap + 32;   // upper half of the same buffer (k <= 16)
        __syncwarp();
        if (lane < k) fin[lane] = key;
        __syncwarp();
        if (qi == 0) trace_stamp(p, 12, cw, lane);
        if (p.xchg) xchg_send(p.xchg, p.xchg_seq, qi, fin, k, k, lane);
        else write_results(fin, k, k, p.outD + size_t(qi) * k, p.outI + size_t(qi) * k, p.label_offset, lane);
    }
    if (cw == 0 && lane == 0) {  // ready for the next launch on this workspace
        *p.ticket = 0u;
        if (p.tile_ctr) *p.tile_ctr = 0u;
    }
    trace_stamp(p, 13, cw, lane);
    if (!p.xchg) return;
    named_bar_sync(bar_id, bar_threads);
    if (cw == 0) xchg_publish_and_wait(p.xchg, p.xchg_seq, lane);
    named_bar_sync(bar_id, bar_threads);
    for (int qi = cw; qi < nq; qi += ncw)
        xchg_merge(p.xchg, p.xchg_seq, qi, selbuf + size_t(cw * nq + qi) * cap, cap, k, p.outD + size_t(qi) * k,
                   p.outI + size_t(qi) * k, lane);
}

__device__ __forceinline__ void finish_scan(const ScanParams& p, uint8_t* smem_base, SmemHeader* hdr,
                                            uint64_t* selbuf, int cw, int ncw, int lane, int bar_id,
                                            int bar_threads) {
    const int nq = p.nq, k = p.k, cap = p.cap;
    const int G = gridDim.x;
    named_bar_sync(bar_id, bar_threads);
    if (blockIdx.x == 0) trace_stamp(p, 3, cw, lane);
    pdl_wait();   // the previous search on this stream is complete: shared scratch and outputs are ours
    if (!p.pdl_early) pdl_launch_dependents();   // ... and once EVERY CTA is here, the next search may start scanning
    if (fast_tail(p, ncw)) {
        finish_scan_fast(p, hdr, selbuf, cw, ncw, lane, bar_id, bar_threads);
        return;
    }
    // ---- CTA merge: warp (qi % ncw) owns query qi -------------------------
    for (int qi = cw; qi < nq; qi += ncw) {
        WarpSelect m = merge_cta_lists(hdr, selbuf, nq, qi, cw, ncw, cap, k, lane);
        uint64_t* dst = p.partials + (size_t(qi) * G + blockIdx.x) * k;
        for (int i = lane; i < k; i += kWarp) dst[i] = (i < m.cnt) ? m.buf[i] : kEmptyKey;
    }
    if (blockIdx.x == 0) trace_stamp(p, 4, cw, lane);
    __threadfence();
    named_bar_sync(bar_id, bar_threads);
    if (blockIdx.x == 0) trace_stamp(p, 5, cw, lane);
    if (cw == 0 && lane == 0) {
        unsigned t = atomicAdd(p.ticket, 1u);
        hdr->last_flag = (t == unsigned(G - 1));
    }
    named_bar_sync(bar_id, bar_threads);
    if (blockIdx.x == 0) trace_stamp(p, 6, cw, lane);
    if (!hdr->last_flag) return;
    // ---- last CTA: merge the G partial lists of every query ---------------
    trace_stamp(p, 8, cw, lane);
    __threadfence();
    // Every CTA list is sorted best-first, so the lists are streamed COLUMN by
    // column (all heads, then all second entries, ...): the threshold rises
    // after the first few columns and a column in which no list beats it ends
    // the merge -- deeper entries are smaller still.  Warp cw takes lists
    // cw, cw+ncw, ...; one lane per list.
    // The G*k keys of a query are first pulled into shared memory with one
    // coalesced sweep (one L2 round trip instead of one per column).
    uint64_t* mbuf = reinterpret_cast<uint64_t*>(smem_base + p.merge_off);
    const bool staged = size_t(G) * k * 8 <= p.merge_bytes;
    for (int qi = 0; qi < nq; qi++) {
        WarpSelect f;
        f.init(selbuf + size_t(cw * nq + qi) * cap, cap, k);
        const uint64_t* src = p.partials + size_t(qi) * G * k;
        if (staged) {
            if (qi) named_bar_sync(bar_id, bar_threads);
            // loads first, stores after: issued one by one the compiler keeps every global load behind
            // the previous shared-memory store (possible alias) and the sweep costs a round trip per key
            for (int i0 = cw * kWarp + lane; i0 < G * k; i0 += 8 * bar_threads) {
                uint64_t tmp[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + u * bar_threads;
                    tmp[u] = (i < G * k) ? __ldcg(src + i) : kEmptyKey;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int i = i0 + u * bar_threads;
                    if (i < G * k) mbuf[i] = tmp[u];
                }
            }
            named_bar_sync(bar_id, bar_threads);
            src = mbuf;
            if (qi == 0) trace_stamp(p, 9, cw, lane);
        }
        for (int col = 0; col < k; col++) {
            bool any = false;
            for (int base = cw; base < G; base += ncw * kWarp) {
                int c = base + lane * ncw;
                uint64_t key = kEmptyKey;
                if (c < G) key = staged ? src[size_t(c) * k + col] : __ldcg(src + size_t(c) * k + col);
                any |= f.push(c < G, key, lane);
            }
            if (!any) break;
        }
        if (qi == 0) trace_stamp(p, 10, cw, lane);
        f.compact(lane);
        if (lane == 0) hdr->cnts[cw * nq + qi] = f.cnt;
    }
    named_bar_sync(bar_id, bar_threads);
    trace_stamp(p, 11, cw, lane);
    for (int qi = cw; qi < nq; qi += ncw) {
        WarpSelect m = merge_cta_lists(hdr, selbuf, nq, qi, cw, ncw, cap, k, lane);
        __syncwarp();
        if (qi == 0) trace_stamp(p, 12, cw, lane);
        if (p.xchg) xchg_send(p.xchg, p.xchg_seq, qi, m.buf, m.cnt, k, lane);
        else write_results(m.buf, m.cnt, k, p.outD + size_t(qi) * k, p.outI + size_t(qi) * k, p.label_offset, lane);
    }
    if (cw == 0 && lane == 0) {  // ready for the next launch on this workspace
        *p.ticket = 0u;
        if (p.tile_ctr) *p.tile_ctr = 0u;
    }
    trace_stamp(p, 13, cw, lane);
    if (!p.xchg) return;
    // ---- fused exchange: all lists are on their way to every rank -----------
    named_bar_sync(bar_id, bar_threads);
    if (cw == 0) xchg_publish_and_wait(p.xchg, p.xchg_seq, lane);
    named_bar_sync(bar_id, bar_threads);
    for (int qi = cw; qi < nq; qi += ncw)
        xchg_merge(p.xchg, p.xchg_seq, qi, selbuf + size_t(cw * nq + qi) * cap, cap, k, p.outD + size_t(qi) * k,
                   p.outI + size_t(qi) * k, lane);
}

// Stand-alone exchange for a rank whose shard is empty (it still has to take part).
__global__ void __launch_bounds__(128) xchg_empty_kernel(const XchgDev* x, uint64_t seq, int nq, int k, float* D,
                                                         int64_t* I) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int cap = select_cap(k);
    for (int qi = warp; qi < nq; qi += nw) xchg_send(x, seq, qi, nullptr, 0, k, lane);
    __syncthreads();
    if (warp == 0) xchg_publish_and_wait(x, seq, lane);
    __syncthreads();
    for (int qi = warp; qi < nq; qi += nw)
        xchg_merge(x, seq, qi, buf + size_t(warp) * cap, cap, k, D + size_t(qi) * k, I + size_t(qi) * k, lane);
}

// Load (and optionally L2-normalise) this lane's float4 chunks of one query.
template <int D4>
__device__ __forceinline__ void load_query_regs(const float* q, int d, int ld4, int normalize,
                                                int lane, float4 (&qr)[D4]) {
    float nr = 0.f;
    const bool vec = ((d & 3) == 0) && ((reinterpret_cast<uintptr_t>(q) & 15) == 0);
#pragma unroll
    for (int j = 0; j < D4; j++) {
        int c = lane + 32 * j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec) {
            if (4 * c < d) v = reinterpret_cast<const float4*>(q)[c];
        } else if (c < ld4) {
            int b = 4 * c;
            if (b + 0 < d) v.x = q[b + 0];
            if (b + 1 < d) v.y = q[b + 1];
            if (b + 2 < d) v.z = q[b + 2];
            if (b + 3 < d) v.w = q[b + 3];
        }
        qr[j] = v;
        nr = dot4(v, v, nr);
    }
    if (normalize) {
        nr = warp_allsum(nr);
        if (nr > 0.f) {
            float inv = renorm_scale(nr);
#pragma unroll
            for (int j = 0; j < D4; j++) {
                qr[j].x *= inv;
                qr[j].y *= inv;
                qr[j].z *= inv;
                qr[j].w *= inv;
            }
        }
    }
}

// dynamic scheduler: tiles are claimed kDynChunk at a time = one 32-row word of the bitmasks
constexpr uint32_t kDynChunk = 4;
constexpr uint32_t kNoTile = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t admissible_word(const ScanParams& p, uint32_t chunk) {
    uint32_t adm = 0xFFFFFFFFu;
    if (p.mask) adm &= p.mask[chunk];
    if (p.live) adm &= p.live[chunk];
    return adm;
}

__device__ __forceinline__ uint32_t admissible_byte(const ScanParams& p, uint32_t tile) {
    uint32_t adm = 0xFFu;
    if (p.mask) adm &= reinterpret_cast<const uint8_t*>(p.mask)[tile];
    if (p.live) adm &= reinterpret_cast<const uint8_t*>(p.live)[tile];
    return adm;
}

// ---------------------------------------------------------------------------
// Producer of the TMA ring (one thread).  Tiles come from two schedules:
//   * static: iteration `it` < n_static of CTA b is tile b + it*G, as a plain round-robin;
//   * dynamic (p.tile_ctr != NULL): the remaining tiles [p.dyn_tile0, T) are claimed from a
//     global counter in chunks of kDynChunk consecutive tiles.  A purely static split leaves
//     the slowest SM ~15% behind the fastest (far-die / channel contention) and that tail is
//     idle HBM; a purely dynamic one pays the claim latency before the first byte moves.
//     Claims run two ahead (the first two are issued before the static phase), so neither the
//     atomic's nor the mask word's latency is ever waited on.
// The producer publishes (tile, admissible byte) of a dynamic stage in the header before it
// arms the full barrier, and ends with one stop marker per consumer warp (warp w owns the
// stages s with s % ncw == w).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void scan_producer(const ScanParams& p, SmemHeader* hdr, uint8_t* smem, int ncw,
                                              uint32_t iters, uint32_t T) {
    const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    const uint32_t row_bytes = uint32_t(p.ld4) * 16u;
    const uint32_t G = gridDim.x;
    const int S = p.stages;
    const bool dyn = p.tile_ctr != nullptr;
    const uint32_t n_static = dyn ? p.static_iters : iters;
    const uint32_t nchunks = dyn ? (T - p.dyn_tile0 + kDynChunk - 1) / kDynChunk : 0u;
    const uint32_t word0 = p.dyn_tile0 / kDynChunk;
    uint32_t c0 = 0, c1 = 0;
    if (dyn) {  // both issued unconditionally: nothing here waits on the first one's result
        c0 = atomicAdd(p.tile_ctr, 1u);
        c1 = atomicAdd(p.tile_ctr, 1u);
    }
    int s = 0;
    uint32_t ph = 0;
    auto issue = [&](uint32_t tile) {
        const uint32_t row0 = tile * kRowsPerTile;
        const uint32_t bytes = min(uint32_t(kRowsPerTile), p.n - row0) * row_bytes;
        mbar_arrive_expect_tx(&hdr->full[s], bytes);
        bulk_g2s(smem + p.stage_off + size_t(s) * p.stage_bytes, p.x + size_t(row0) * size_t(p.ld4) * 4, bytes,
                 &hdr->full[s], tile < p.pin_tiles ? pol_keep : pol_stream);
        if (++s == S) {
            s = 0;
            ph ^= 1u;
        }
    };
    for (uint32_t it = 0; it < n_static; it++) {
        mbar_wait(&hdr->empty[s], ph ^ 1u);
        issue(blockIdx.x + it * G);
    }
    if (!dyn) return;
    if (p.dep_inputs) pdl_wait();   // the filter words may still be on their way (the static phase reads none)
    uint32_t a0 = (c0 < nchunks) ? admissible_word(p, word0 + c0) : 0u;
    while (c0 < nchunks) {
        // every claim's result is consumed before this loop ends, so none is in flight when
        // the last CTA resets the counter
        const uint32_t c2 = (c1 < nchunks) ? atomicAdd(p.tile_ctr, 1u) : nchunks;
        const uint32_t a1 = (c1 < nchunks) ? admissible_word(p, word0 + c1) : 0u;
#pragma unroll
        for (uint32_t t = 0; t < kDynChunk; t++) {
            const uint32_t tile = p.dyn_tile0 + c0 * kDynChunk + t;
            if (tile >= T) break;
            mbar_wait(&hdr->empty[s], ph ^ 1u);
            hdr->tile_of[s] = tile;
            hdr->adm_of[s] = (a0 >> (8 * t)) & 0xFFu;
            issue(tile);
        }
        c0 = c1;
        c1 = c2;
        a0 = a1;
    }
    asm volatile("" ::"r"(c1));  // the second claim has returned even when the loop never ran
    for (int i = 0; i < ncw; i++) {
        mbar_wait(&hdr->empty[s], ph ^ 1u);
        hdr->tile_of[s] = kNoTile;
        mbar_arrive(&hdr->full[s]);
        if (++s == S) {
            s = 0;
            ph ^= 1u;
        }
    }
}

// ---------------------------------------------------------------------------
// Best k of the `ns` unsorted keys in sk[0, ns) (shared memory), by ALL `nthr` consumer threads of the CTA:
// a 32-round block-wide bisection finds the k-th largest score image (every thread keeps the high words of
// its <= 32 keys in registers; one warp redux + one shared atomic + one barrier per round), the keys at or
// above it (k plus ties) are gathered behind the list and warp 0 sorts those few in registers.  Replaces a
// full bitonic sort of 1-4 k keys (14-30 us in the trace) by ~5 us.  On return warp 0 (cw == 0) holds the
// sorted list at the returned pointer, every other warp gets nullptr and may leave.  ns <= kSelectMax; keys beyond
// the first 32 per thread are counted from shared memory (only very long lists get there).
// Scratch: cnts4 (4 words of shared memory), sk[kSelectMax ...) for the gathered keys.
// ---------------------------------------------------------------------------
constexpr uint32_t kSelectMax = 8192;
__device__ __forceinline__ uint64_t* block_select_topk(uint64_t* sk, unsigned int ns, int k, int* cnts4, int tid, int nthr,
                                                       int cw, int lane, int* out_cnt) {
    uint32_t npad = 64;
    while (npad < ns) npad <<= 1;
    if (npad <= 256) {   // short list: one register sort
        if (cw != 0) return nullptr;
        for (uint32_t i = ns + lane; i < npad; i += kWarp) sk[i] = kEmptyKey;
        __syncwarp();
        if (npad == 64) warp_sort_buffer<2>(sk, int(ns), lane);
        else if (npad == 128) warp_sort_buffer<4>(sk, int(ns), lane);
        else warp_sort_buffer<8>(sk, int(ns), lane);
        __syncwarp();
        *out_cnt = int(min(ns, unsigned(k)));
        return sk;
    }
    uint32_t hi[32];
#pragma unroll
    for (int u = 0; u < 32; u++) {
        const uint32_t i = uint32_t(tid) + uint32_t(u) * uint32_t(nthr);
        hi[u] = (i < ns) ? uint32_t(sk[i] >> 32) : 0u;
    }
    volatile int* cnt = cnts4;   // [0..2]: rotating counters, [3]: gather cursor
    if (tid == 0) cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
    named_bar_sync(1, nthr);
    uint32_t t = 0;
    int above = 0;   // keys with image >= t
#pragma unroll 1
    for (int r = 0; r < 32; r++) {
        const uint32_t cand = t | (1u << (31 - r));
        int c = 0;
#pragma unroll
        for (int u = 0; u < 32; u++) c += (hi[u] >= cand) ? 1 : 0;
        for (uint32_t i = uint32_t(tid) + 32u * uint32_t(nthr); i < ns; i += uint32_t(nthr)) c += (uint32_t(sk[i] >> 32) >= cand) ? 1 : 0;
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane == 0 && c) atomicAdd(&cnts4[r % 3], c);
        named_bar_sync(1, nthr);
        const int total = cnt[r % 3];
        if (tid == 0) cnt[(r + 2) % 3] = 0;   // used next in round r + 2: everybody is past its last read (round r - 1)
        if (total >= k) {
            t = cand;
            above = total;
        }
    }
    if (t == 0u) above = int(ns);   // fewer than k keys: everything counts
    if (above > 256) {   // a crowd of exact score ties: sort everything (rare)
        for (uint32_t i = ns + tid; i < npad; i += nthr) sk[i] = kEmptyKey;
        named_bar_sync(1, nthr);
        bitonic_sort_desc(sk, int(npad), tid, nthr, [&] { named_bar_sync(1, nthr); });
        if (cw != 0) return nullptr;
        *out_cnt = int(min(ns, unsigned(k)));
        return sk;
    }
    uint64_t* out = sk + kSelectMax;
#pragma unroll
    for (int u = 0; u < 32; u++) {
        const uint32_t i = uint32_t(tid) + uint32_t(u) * uint32_t(nthr);
        if (i < ns && hi[u] >= t) out[atomicAdd(&cnts4[3], 1)] = sk[i];
    }
    for (uint32_t i = uint32_t(tid) + 32u * uint32_t(nthr); i < ns; i += uint32_t(nthr))
        if (uint32_t(sk[i] >> 32) >= t) out[atomicAdd(&cnts4[3], 1)] = sk[i];
    named_bar_sync(1, nthr);
    if (cw != 0) return nullptr;
    const int m = cnt[3];
    if (m <= 64) warp_sort_buffer<2>(out, m, lane);
    else if (m <= 128) warp_sort_buffer<4>(out, m, lane);
    else warp_sort_buffer<8>(out, m, lane);
    __syncwarp();
    *out_cnt = min(m, k);
    return out;
}

// ---------------------------------------------------------------------------
// Survivor mode tail: the last CTA to finish sorts the survivor list (in the idle ring) and writes (D, I) -- or
// sends / merges over NVLink exactly as finish_scan does.  Leaves best[], the counters and the tile counter clean.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void finish_survivors(const ScanParams& p, uint8_t* smem_base, SmemHeader* hdr, int cw, int ncw,
                                                 int lane) {
    const uint32_t G = gridDim.x;
    const int tid = cw * kWarp + lane, nthr = ncw * kWarp;
    __threadfence();
    named_bar_sync(1, nthr);
    if (tid == 0) {
        const bool last = atomicAdd(&p.sctl->ticket, 1u) == G - 1;
        hdr->last_flag = last;
        if (last) {
            // ONE thread reads the list length and resets the shared state (every other CTA is done with it); the
            // others take the length from shared memory after the barrier -- a thread that read the global counter
            // itself could see the reset value and disagree with its CTA about the list
            __threadfence();
            const unsigned int cnt_now = *reinterpret_cast<volatile unsigned int*>(&p.sctl->count);
            hdr->tail_ns = cnt_now;
            p.sctl->last_count = cnt_now;
            p.sctl->count = 0u;
            p.sctl->ticket = 0u;
            if (p.tile_ctr) *p.tile_ctr = 0u;
        }
    }
    named_bar_sync(1, nthr);
    if (blockIdx.x == 0) trace_stamp(p, 6, cw, lane);
    if (!hdr->last_flag) return;
    trace_stamp(p, 8, cw, lane);
    __threadfence();
    const unsigned int ns = hdr->tail_ns;
    for (uint32_t i = tid; i < p.nbest; i += nthr) p.best[i] = 0u;
    if (ns > p.surv_cap) {   // the classic scan answers instead (conditional launch behind this one, or the host re-runs it)
        if (tid == 0) {
            p.sctl->overflow = 1u;
            if (p.ovf_host) {
                *reinterpret_cast<volatile unsigned int*>(p.ovf_host) = 1u;
                __threadfence_system();
            }
        }
        return;
    }
    uint64_t* sk = reinterpret_cast<uint64_t*>(smem_base + p.merge_off);
    for (uint32_t i = tid; i < ns; i += nthr) sk[i] = __ldcg(p.surv + i);
    named_bar_sync(1, nthr);
    trace_stamp(p, 9, cw, lane);
    int cnt = 0;
    uint64_t* fin = block_select_topk(sk, ns, p.k, hdr->cnts, tid, nthr, cw, lane, &cnt);
    if (!fin) return;
    trace_stamp(p, 10, cw, lane);
    if (p.xchg) {
        xchg_send(p.xchg, p.xchg_seq, 0, fin, cnt, p.k, lane);
        xchg_publish_and_wait(p.xchg, p.xchg_seq, lane);
        xchg_merge(p.xchg, p.xchg_seq, 0, fin == sk ? sk + kSelectMax : sk, select_cap(p.k), p.k, p.outD, p.outI, lane);
    } else {
        write_results(fin, cnt, p.k, p.outD, p.outI, p.label_offset, lane);
    }
    trace_stamp(p, 13, cw, lane);
}

// ---------------------------------------------------------------------------
// Kernel A: one query, query chunks in registers, D4 = ceil(ld4 / 32) known at
// compile time.  kTma selects the producer/consumer ring or direct loads.
// ---------------------------------------------------------------------------
template <int D4, bool kTma, bool kSurv = false>
__global__ void __launch_bounds__(kTma ? 288 : 256, kTma ? 1 : 2) scan_q1_kernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
    uint64_t* selbuf = reinterpret_cast<uint64_t*>(smem + p.sel_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (blockDim.x >> 5) - (kTma ? 1 : 0);
    const int cw = warp - (kTma ? 1 : 0);
    const uint32_t G = gridDim.x;
    if (p.run_if && *p.run_if == 0u) return;    // conditional launch: nothing to redo
    if (p.pdl_early) pdl_launch_dependents();   // one CTA per SM: the next search may take over every SM we leave
    const uint32_t T = (p.n + kRowsPerTile - 1) / kRowsPerTile;
    const uint32_t iters = (T > blockIdx.x) ? (T - blockIdx.x + G - 1) / G : 0;
    const int S = p.stages;

    if (kTma) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < S; s++) {
                mbar_init(&hdr->full[s], 1);
                mbar_init(&hdr->empty[s], 1);
            }
            if (kSurv) {
                hdr->surv_thr = 0u;
                hdr->surv_refreshes = 0u;
                if (blockIdx.x == 0) p.sctl->overflow = 0u;   // the previous search's conditional fallback has completed
            }
            mbar_fence_init();
        }
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) scan_producer(p, hdr, smem, ncw, iters, T);
            return;
        }
    }

    if (blockIdx.x == 0) trace_stamp(p, 0, cw, lane);
    if (p.dep_inputs) pdl_wait();   // query and filter staged by the preceding pull grid; the ring is filling meanwhile
    float4 qr[D4];
    load_query_regs<D4>(p.q, p.d, p.ld4, p.normalize_q, lane, qr);
    if (blockIdx.x == 0) trace_stamp(p, 1, cw, lane);

    WarpSelect sel;
    sel.init(selbuf + size_t(cw) * p.cap, p.cap, p.k);
    const int my_row = tile_row_of_lane(lane);
    const bool leader = (lane & 3) == 0;
    const int ld4 = p.ld4;

    // survivor mode state: slot of p.best this warp reports to (k <= 32: one per CTA), its best score image so far,
    // the threshold, and the refresh schedule (after 1, 2, 4, ... 256 tiles, then every 256)
    // One slot per CTA: 148 words make the threshold cheap to recompute (per-warp slots were measured: the k = 100
    // search at 1 M x 384 took 278 us instead of 241), and the k-th largest of 148 CTA-bests is still about the
    // 1.5 k ... 2.5 k-th best row -- the list holds a few thousand keys at most (cap 8192).
    const bool slot_per_cta = true;
    const uint32_t gw = blockIdx.x;
    uint32_t my_best = 0u, published = 0u, thr = 0u, done_tiles = 0, next_refresh = 2, seen_refreshes = 0;

    const bool dyn = kTma && p.tile_ctr != nullptr;
    const uint32_t n_static = dyn ? p.static_iters : iters;
    for (uint32_t it = cw;; it += ncw) {
        uint32_t tile, adm;
        if (it < n_static) {
            tile = blockIdx.x + it * G;
            adm = admissible_byte(p, tile);  // issued before the data wait
            if (kTma) mbar_wait(&hdr->full[it % S], (it / S) & 1u);
        } else if (dyn) {
            const int s = it % S;
            mbar_wait(&hdr->full[s], (it / S) & 1u);
            tile = hdr->tile_of[s];
            if (tile == kNoTile) break;
            adm = hdr->adm_of[s];
        } else {
            break;
        }
        const uint32_t row0 = tile * kRowsPerTile;
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; r++) acc[r] = 0.f;
        if (kTma) {
            const int s = it % S;
            const float4* st = reinterpret_cast<const float4*>(smem + p.stage_off + size_t(s) * p.stage_bytes);
#pragma unroll
            for (int j = 0; j < D4; j++) {
                const int c = lane + 32 * j;
                if (c < ld4) {
#pragma unroll
                    for (int r = 0; r < 8; r++) acc[r] = dot4(st[r * ld4 + c], qr[j], acc[r]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->empty[s]);
        } else {
            const float4* xt = reinterpret_cast<const float4*>(p.x) + size_t(row0) * ld4;
            const uint32_t rows = min(uint32_t(kRowsPerTile), p.n - row0);
#pragma unroll
            for (int j = 0; j < D4; j++) {
                const int c = lane + 32 * j;
                if (c < ld4) {
                    float4 v[8];
#pragma unroll
                    for (int r = 0; r < 8; r++)
                        v[r] = (uint32_t(r) < rows) ? ldg_stream(xt + size_t(r) * ld4 + c)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int r = 0; r < 8; r++) acc[r] = dot4(v[r], qr[j], acc[r]);
                }
            }
        }
        const float score = reduce8(acc, lane);
        const uint32_t row = row0 + my_row;
        const bool ok = leader && row < p.n && ((adm >> my_row) & 1u);
        if (kSurv) {
            // publish this warp's best score, refresh the shared threshold on schedule, THEN test this tile's rows
            // (so that not even the first tile floods the list) and append the keys that pass
            const uint32_t o = (ok && score == score) ? max(score_to_ord(score), 1u) : 0u;
            my_best = max(my_best, o);
            const uint32_t wb = __reduce_max_sync(0xFFFFFFFFu, my_best);
            if (wb > published) {
                published = wb;
                if (lane == 0) {
                    if (slot_per_cta) atomicMax(p.best + gw, wb);
                    else *reinterpret_cast<volatile unsigned int*>(p.best + gw) = wb;
                }
            }
            my_best = wb;
            ++done_tiles;
            if (done_tiles == 1) {
                // the first threshold: warp 0 of the CTA computes it as soon as k CTAs have reported their first tile,
                // the other warps wait on shared memory (cheap) -- nobody tests a row against "no threshold"
                if (cw == 0) {
                    uint32_t t = i8_threshold(p.best, p.nbest, p.k, lane);
                    for (int spin = 0; t == 0u && spin < 40; spin++) {
                        __nanosleep(200);
                        t = i8_threshold(p.best, p.nbest, p.k, lane);
                    }
                    if (lane == 0) {
                        atomicMax(&hdr->surv_thr, max(t, 1u));
                        atomicAdd(&hdr->surv_refreshes, 1u);
                    }
                } else {
                    for (int spin = 0; *reinterpret_cast<volatile unsigned int*>(&hdr->surv_thr) == 0u && spin < 400; spin++) __nanosleep(100);
                }
                seen_refreshes = *reinterpret_cast<volatile unsigned int*>(&hdr->surv_refreshes);
            } else if (done_tiles >= next_refresh) {
                next_refresh = done_tiles < 256 ? done_tiles + max(1u, done_tiles / 2) : done_tiles + 256;
                if (*reinterpret_cast<volatile unsigned int*>(&hdr->surv_refreshes) == seen_refreshes) {
                    const uint32_t t = i8_threshold(p.best, p.nbest, p.k, lane);
                    if (lane == 0) {
                        atomicMax(&hdr->surv_thr, t);
                        atomicAdd(&hdr->surv_refreshes, 1u);
                    }
                    thr = max(thr, t);
                }
                seen_refreshes = *reinterpret_cast<volatile unsigned int*>(&hdr->surv_refreshes);
            }
            thr = max(thr, *reinterpret_cast<volatile unsigned int*>(&hdr->surv_thr));
            const bool pass = o != 0u && o >= thr;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, pass);
            if (m) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&p.sctl->count, unsigned(__popc(m)));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                const unsigned pos = base + unsigned(__popc(m & ((1u << lane) - 1u)));
                if (pass && pos < p.surv_cap) p.surv[pos] = make_key(score, row);
            }
        } else if (p.all_ord) {
            if (leader && row < p.n) {
                uint32_t o = (ok && score == score) ? score_to_ord(score) : 0u;
                p.all_ord[row] = o;
            }
        } else {
            sel.push(ok, make_key(score, row), lane);
        }
    }
    if (kSurv) {
        if (blockIdx.x == 0) trace_stamp(p, 2, cw, lane);
        finish_survivors(p, smem, hdr, cw, ncw, lane);
        return;
    }
    if (p.all_ord) return;
    if (blockIdx.x == 0) trace_stamp(p, 2, cw, lane);
    if (!fast_tail(p, ncw)) sel.compact(lane);   // small k: the CTA merge reads the raw buffers (no sort)
    if (lane == 0) hdr->cnts[cw] = sel.cnt;
    finish_scan(p, smem, hdr, selbuf, cw, ncw, lane, 1, ncw * 32);
}

// ---------------------------------------------------------------------------
// Kernel B: NQ queries per pass (NQ in {1,2,4,8}), queries staged in shared
// memory, arbitrary dimension (runtime chunk loop).  Used for small batches
// and for dimensions whose chunk count exceeds kernel A's templates.
// ---------------------------------------------------------------------------
template <int NQ, bool kTma>
__global__ void __launch_bounds__(kTma ? 288 : 256, (kTma || NQ >= 8) ? 1 : 2) scan_multi_kernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
    uint64_t* selbuf = reinterpret_cast<uint64_t*>(smem + p.sel_off);
    float4* qs = reinterpret_cast<float4*>(smem + p.q_off);  // [NQ][ld4]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (blockDim.x >> 5) - (kTma ? 1 : 0);
    const int cw = warp - (kTma ? 1 : 0);
    const uint32_t G = gridDim.x;
    if (p.pdl_early) pdl_launch_dependents();   // one CTA per SM: the next search may take over every SM we leave
    const uint32_t T = (p.n + kRowsPerTile - 1) / kRowsPerTile;
    const uint32_t iters = (T > blockIdx.x) ? (T - blockIdx.x + G - 1) / G : 0;
    const int S = p.stages;
    const int ld4 = p.ld4;

    if (kTma) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < S; s++) {
                mbar_init(&hdr->full[s], 1);
                mbar_init(&hdr->empty[s], 1);
            }
            mbar_fence_init();
        }
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) scan_producer(p, hdr, smem, ncw, iters, T);
            return;
        }
    }

    // stage (and normalise) the queries: warp cw handles queries cw, cw+ncw, ...
    for (int qi = cw; qi < NQ; qi += ncw) {
        const float* q = p.q + size_t(qi) * p.d;
        float nr = 0.f;
        for (int c = lane; c < ld4; c += kWarp) {
            int b = 4 * c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b + 0 < p.d) v.x = q[b + 0];
            if (b + 1 < p.d) v.y = q[b + 1];
            if (b + 2 < p.d) v.z = q[b + 2];
            if (b + 3 < p.d) v.w = q[b + 3];
            qs[qi * ld4 + c] = v;
            nr = dot4(v, v, nr);
        }
        if (p.normalize_q) {
            nr = warp_allsum(nr);
            if (nr > 0.f) {
                float inv = renorm_scale(nr);
                for (int c = lane; c < ld4; c += kWarp) {
                    float4 v = qs[qi * ld4 + c];
                    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                    qs[qi * ld4 + c] = v;
                }
            }
        }
    }
    named_bar_sync(1, ncw * 32);

    WarpSelect sel[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) sel[qi].init(selbuf + size_t(cw * NQ + qi) * p.cap, p.cap, p.k);
    const int my_row = tile_row_of_lane(lane);
    const bool leader = (lane & 3) == 0;

    const bool dyn = kTma && p.tile_ctr != nullptr;
    const uint32_t n_static = dyn ? p.static_iters : iters;
    for (uint32_t it = cw;; it += ncw) {
        uint32_t tile, adm;
        if (it < n_static) {
            tile = blockIdx.x + it * G;
            adm = admissible_byte(p, tile);
            if (kTma) mbar_wait(&hdr->full[it % S], (it / S) & 1u);
        } else if (dyn) {
            mbar_wait(&hdr->full[it % S], (it / S) & 1u);
            tile = hdr->tile_of[it % S];
            if (tile == kNoTile) break;
            adm = hdr->adm_of[it % S];
        } else {
            break;
        }
        const uint32_t row0 = tile * kRowsPerTile;
        float acc[NQ][8];
#pragma unroll
        for (int qi = 0; qi < NQ; qi++)
#pragma unroll
            for (int r = 0; r < 8; r++) acc[qi][r] = 0.f;

        const float4* st;
        uint32_t rows = 8;
        int s = 0;
        if (kTma) {
            s = it % S;
            st = reinterpret_cast<const float4*>(smem + p.stage_off + size_t(s) * p.stage_bytes);
        } else {
            st = reinterpret_cast<const float4*>(p.x) + size_t(row0) * ld4;
            rows = min(uint32_t(kRowsPerTile), p.n - row0);
        }
        for (int c = lane; c < ld4; c += kWarp) {
            float4 v[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                if (kTma) v[r] = st[r * ld4 + c];
                else v[r] = (uint32_t(r) < rows) ? ldg_stream(st + size_t(r) * ld4 + c)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int qi = 0; qi < NQ; qi++) {
                const float4 qv = qs[qi * ld4 + c];
#pragma unroll
                for (int r = 0; r < 8; r++) acc[qi][r] = dot4(v[r], qv, acc[qi][r]);
            }
        }
        if (kTma) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr->empty[s]);
        }
        const uint32_t row = row0 + my_row;
        const bool ok = leader && row < p.n && ((adm >> my_row) & 1u);
#pragma unroll
        for (int qi = 0; qi < NQ; qi++) {
            const float score = reduce8(acc[qi], lane);
            bool ok_q = ok;
            if (p.has_qmask && p.qmask[qi])   // coalesced single-query searches keep their own filters
                ok_q = ok && tile < p.qmask_bytes[qi] &&
                       ((reinterpret_cast<const uint8_t*>(p.qmask[qi])[tile] >> my_row) & 1u);
            if (p.all_ord) {
                if (leader && row < p.n)
                    p.all_ord[size_t(qi) * p.n + row] = (ok_q && score == score) ? score_to_ord(score) : 0u;
            } else {
                sel[qi].push(ok_q, make_key(score, row), lane);
            }
        }
    }
    if (p.all_ord) return;
    const bool fast = fast_tail(p, ncw);
#pragma unroll
    for (int qi = 0; qi < NQ; qi++) {
        if (!fast) sel[qi].compact(lane);
        if (lane == 0) hdr->cnts[cw * NQ + qi] = sel[qi].cnt;
    }
    finish_scan(p, smem, hdr, selbuf, cw, ncw, lane, 1, ncw * 32);
}

}  // namespace mvdb
