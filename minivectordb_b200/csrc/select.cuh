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
        for (int i = cnt + lane; i < cap; i += kWarp) buf[i] = kEmptyKey;
        __syncwarp();
        bitonic_sort_desc(buf, cap, lane, kWarp, WarpSyncer());
        if (cnt > k) cnt = k;
        thr = (cnt == k) ? buf[k - 1] : kEmptyKey;
    }

    // Called by all 32 lanes (converged).  `key` is considered only where
    // `valid`; kEmptyKey never passes because thr >= kEmptyKey.
    // Returns true (warp-uniform) when at least one candidate was admitted.
    __device__ __forceinline__ bool push(bool valid, uint64_t key, int lane) {
        bool pass = valid && key > thr;
        unsigned m = __ballot_sync(0xFFFFFFFFu, pass);
        if (m) {
            int pos = cnt + __popc(m & ((1u << lane) - 1u));
            if (pass) buf[pos] = key;
            cnt += __popc(m);
            if (cnt + kWarp > cap) compact(lane);
        }
        return m != 0;
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
