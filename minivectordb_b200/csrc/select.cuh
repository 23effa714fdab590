// select.cuh -- top-k selection building blocks.
//
// WarpSelect: a warp streams candidate keys through a threshold filter into a
// small shared-memory buffer; when the buffer fills it is bitonic-sorted
// (warp shuffles are not needed: the buffer lives in shared memory and the
// warp synchronises with __syncwarp) and cut back to the best k, which raises
// the threshold.  With keys = (score, ~row) the result is independent of the
// order in which candidates arrive, so the scan is deterministic for any grid.
//
// This replaces faiss's HeapBlockResultHandler / ReservoirBlockResultHandler
// (consumed at ref vector_database.py:497).
#pragma once
#include "device_utils.cuh"

namespace mvdb {

// In-place bitonic sort, DESCENDING, of n (power of two) keys in shared memory
// by `nthr` cooperating threads; `sync` separates the compare-exchange stages.
template <class Sync>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* a, int n, int tid, int nthr, Sync sync) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (n >> 1); t += nthr) {
                int i = 2 * t - (t & (stride - 1));
                int j = i + stride;
                bool desc = (i & size) == 0;
                uint64_t x = a[i], y = a[j];
                if ((x < y) == desc) {
                    a[i] = y;
                    a[j] = x;
                }
            }
            sync();
        }
    }
}

struct WarpSyncer {
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};

// ---------------------------------------------------------------------------
// Register bitonic sort, DESCENDING, of 32*R keys held R per lane: element e = r*32 + lane.
// Compare-exchange distances below 32 are one 64-bit shuffle, distances of 32 and more are
// exchanges between registers of the same lane.  Same network as bitonic_sort_desc, but a stage
// costs one shuffle latency instead of a shared-memory round trip plus __syncwarp (measured on
// the scan's merge tail: 2.3 us per 64-key sort in shared memory).
// ---------------------------------------------------------------------------
template <int R, int RS>
__device__ __forceinline__ void warp_sort_xchg_regs(uint64_t (&v)[R], int size) {
#pragma unroll
    for (int r = 0; r < R; r++) {
        if ((r & RS) == 0 && (r | RS) < R) {
            const bool desc = ((r << 5) & size) == 0;   // size >= 64 here: only the register index decides
            const uint64_t x = v[r], y = v[r | RS];
            const bool sw = (x < y) == desc;
            v[r] = sw ? y : x;
            v[r | RS] = sw ? x : y;
        }
    }
}
template <int R>
__device__ __forceinline__ void warp_sort_desc_regs(uint64_t (&v)[R], int lane) {
#pragma unroll 1
    for (int size = 2; size <= 32 * R; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int rs = stride >> 5;
                if (rs == 1) warp_sort_xchg_regs<R, 1>(v, size);
                else if (rs == 2) warp_sort_xchg_regs<R, 2>(v, size);
                else warp_sort_xchg_regs<R, 4>(v, size);
            } else {
                const bool lower = (lane & stride) == 0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const bool desc = (((r << 5) | lane) & size) == 0;
                    const uint64_t x = v[r];
                    const uint64_t y = __shfl_xor_sync(0xFFFFFFFFu, x, stride);
                    const bool keep_max = lower == desc;
                    v[r] = keep_max ? (x > y ? x : y) : (x < y ? x : y);
                }
            }
        }
    }
}
// Sort buf[0, 32*R) descending; entries at and beyond cnt are treated as empty.  One copy per R
// in the module (not inlined: compaction is off the scan's hot path and is called from many places).
template <int R>
__device__ __noinline__ void warp_sort_buffer(uint64_t* buf, int cnt, int lane) {
    uint64_t v[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = r * kWarp + lane;
        v[r] = (i < cnt) ? buf[i] : kEmptyKey;
    }
    warp_sort_desc_regs<R>(v, lane);
#pragma unroll
    for (int r = 0; r < R; r++) buf[r * kWarp + lane] = v[r];
}

struct WarpSelect {
    uint64_t* buf;   // shared memory, `cap` keys, private to this warp (and query)
    int cap;         // power of two, >= 2 * 32 and >= 2 * next_pow2(k)... see select_cap()
    int k;
    int cnt;         // warp-uniform
    uint64_t thr;    // warp-uniform: keys <= thr cannot enter the top-k any more

    __device__ __forceinline__ void init(uint64_t* b, int cap_, int k_) {
        buf = b;
        cap = cap_;
        k = k_;
        cnt = 0;
        thr = kEmptyKey;
    }

    // Sort the buffer, keep the best k, refresh the threshold.
    __device__ __forceinline__ void compact(int lane) {
        __syncwarp();
        if (cap == 64) {
            warp_sort_buffer<2>(buf, cnt, lane);
        } else if (cap == 128) {
            warp_sort_buffer<4>(buf, cnt, lane);
        } else if (cap == 256) {
            warp_sort_buffer<8>(buf, cnt, lane);
        } else {
            for (int i = cnt + lane; i < cap; i += kWarp) buf[i] = kEmptyKey;
            __syncwarp();
            bitonic_sort_desc(buf, cap, lane, kWarp, WarpSyncer());
        }
        __syncwarp();
        if (cnt > k) cnt = k;
        thr = (cnt == k) ? buf[k - 1] : kEmptyKey;
    }

    // Called by all 32 lanes (converged).  `key` is considered only where
    // `valid`; kEmptyKey never passes because thr >= kEmptyKey.
    // Returns true (warp-uniform) when at least one candidate was admitted.
    __device__ __forceinline__ bool push(bool valid, uint64_t key, int lane) {
        bool pass = valid && key > thr;
        unsigned m = __ballot_sync(0xFFFFFFFFu, pass);
        if (m == 0) return false;
        if (cnt + __popc(m) > cap) {
            // make room only when the newcomers do not fit (merging a few short lists never sorts
            // in the middle), then re-test them against the raised threshold
            compact(lane);
            pass = pass && key > thr;
            m = __ballot_sync(0xFFFFFFFFu, pass);
            if (m == 0) return false;
        }
        int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (pass) buf[pos] = key;
        cnt += __popc(m);
        return true;
    }

    // Stream `n` keys from memory (global or shared) through the filter.
    __device__ __forceinline__ void push_array(const uint64_t* src, int n, int lane) {
        for (int i = 0; i < n; i += kWarp) {
            int j = i + lane;
            uint64_t key = (j < n) ? src[j] : kEmptyKey;
            push(j < n, key, lane);
        }
    }
};

// ---------------------------------------------------------------------------
// Small-k selection by repeated warp-wide arg-max.  Each lane holds M keys in registers (any order,
// kEmptyKey = none; real keys are unique).  On return lane j (j < k) holds the j-th best of the 32*M
// keys (kEmptyKey once they run out).  One round = two 32-bit redux.sync (high word, then low word among
// the lanes that tie on the high word) and a short fix-up in the single lane that owned the winner --
// ~100 cycles, no shared memory, no sort.  For k ~ 10 this replaces "bitonic-sort 64..256 keys, then
// cut": the merge tails of a search are a serial chain, so their latency is what counts.
// ---------------------------------------------------------------------------
constexpr int kExtractMaxK = 16;
template <int M>
__device__ __forceinline__ uint64_t warp_extract_topk(uint64_t (&v)[M], int k, int lane) {
    uint64_t mine = kEmptyKey;
    uint64_t lmax = v[0];
#pragma unroll
    for (int i = 1; i < M; i++) lmax = v[i] > lmax ? v[i] : lmax;
    for (int j = 0; j < k; j++) {
        const uint32_t hi = uint32_t(lmax >> 32);
        const uint32_t mhi = __reduce_max_sync(0xFFFFFFFFu, hi);
        const uint32_t lo = (hi == mhi) ? uint32_t(lmax) : 0u;   // a real key's low word (~row) is never 0
        const uint32_t mlo = __reduce_max_sync(0xFFFFFFFFu, lo);
        const uint64_t best = (uint64_t(mhi) << 32) | mlo;
        if (best == kEmptyKey) break;   // warp-uniform: nothing left
        if (lane == j) mine = best;
        if (lmax == best) {             // exactly one lane: drop the winner, refresh the local maximum
            uint64_t nm = kEmptyKey;
#pragma unroll
            for (int i = 0; i < M; i++) {
                if (v[i] == best) v[i] = kEmptyKey;
                nm = v[i] > nm ? v[i] : nm;
            }
            lmax = nm;
        }
    }
    return mine;
}

// ---------------------------------------------------------------------------
// Shared threshold without a merge (int8 shadow scan, fp32 survivor scan): every warp (or CTA) publishes
// the ordered image of the best lower bound / score it has seen in one 32-bit word; the k-th largest of
// those words is a lower bound of the final k-th best score (k distinct rows reach it).
// ---------------------------------------------------------------------------
// k-th largest of the 32-bit words spread over the warp (v[i] = word lane + 32 i; 0 = empty), truncated to
// its top 24 bits (a slightly LOWER value: still a valid lower bound): bitwise bisection, 24 rounds of
// (compare, warp-wide count).  Returns 0 when fewer than k words are non-zero.
template <int M>
__device__ __forceinline__ uint32_t warp_kth_largest_u32(const uint32_t (&v)[M], int k) {
    uint32_t t = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 8; bit--) {
        const uint32_t cand = t | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < M; i++) c += (v[i] >= cand) ? 1 : 0;
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (c >= k) t = cand;
    }
    return t;
}

constexpr int kI8BestM = 40;   // up to 1280 consumer warps in a grid

template <int M>
__device__ __forceinline__ uint32_t i8_threshold_m(const unsigned int* best, uint32_t nbest, int k, int lane) {
    uint32_t v[M];
#pragma unroll
    for (int i = 0; i < M; i++) {
        const uint32_t idx = uint32_t(lane) + 32u * uint32_t(i);
        v[i] = (idx < nbest) ? __ldcg(best + idx) : 0u;
    }
    return warp_kth_largest_u32<M>(v, k);
}
__device__ __forceinline__ uint32_t i8_threshold(const unsigned int* best, uint32_t nbest, int k, int lane) {
    if (nbest <= 160u) return i8_threshold_m<5>(best, nbest, k, lane);
    return nbest <= 640u ? i8_threshold_m<20>(best, nbest, k, lane) : i8_threshold_m<kI8BestM>(best, nbest, k, lane);
}

// Buffer capacity for a given k: room for the kept k plus at least one full
// warp of fresh candidates, rounded to a power of two for the bitonic network.
__host__ __device__ __forceinline__ int select_cap(int k) {
    int c = 64;
    while (c < 2 * k) c <<= 1;
    return c;
}
__host__ __device__ __forceinline__ int next_pow2(int v) {
    int c = 1;
    while (c < v) c <<= 1;
    return c;
}

}  // namespace mvdb
