// aux_kernels.cuh -- ingest, tombstones, large-k selection, shard merge.
#pragma once
#include <cfloat>

#include "select.cuh"

namespace mvdb {

// ---------------------------------------------------------------------------
// Host-buffer searches: the query (and a per-call filter) are staged in pinned host memory and PULLED into device
// memory by this small grid instead of a copy-engine transfer.  The scan that follows is launched as a programmatic
// dependent (this grid triggers it at entry), so its start-up and ring fill overlap the PCIe round trip and only its
// first read of the query waits (ScanParams::dep_inputs); a copy-engine transfer costs ~8 us before the scan may even
// start.  Reads are volatile (system scope): the host rewrites the buffer between launches.
// (Measured and dropped: launching this grid BEFORE the host has copied the filter into the pinned buffer, with a
// sequence word the grid waits for.  It was slower -- 266 vs 247 us at 1 M x 384 -- and any allocating CUDA call of
// another host thread could stall behind the waiting grid while the waiting grid's own host thread queued behind
// that call.)
// ---------------------------------------------------------------------------
// Vectors [0, n_direct) come from `direct` (a filter the caller already holds in pinned memory: no staging copy on the
// host at all), the rest from the staging buffer `src` (same indexing as dst).
__global__ void __launch_bounds__(256) pull_stage_kernel(const uint4* src, const uint4* direct, uint32_t n_direct,
                                                         uint4* __restrict__ dst, uint32_t nvec) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        const uint4* from = (i < n_direct) ? direct + i : src + i;
        uint4 v;
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(from) : "memory");
        dst[i] = v;
    }
}

// ---------------------------------------------------------------------------
// K1 ingest: one warp per row; optional faiss.normalize_L2 semantics
// (ref vector_database.py:45 -- nr = sum x^2 in fp32, scale by
// (float)(1.0/sqrtf(nr)) when nr > 0); the row is written with zero padding
// up to the leading dimension.  kSynth generates the row instead of reading it.
// ---------------------------------------------------------------------------
// HBM-bound: n*d*4 bytes read + n*ld*4 bytes written.  Rows of up to 1024 floats are held in registers
// between the norm pass and the write pass (one read of the source, 128-bit loads and stores when the
// layout allows); wider rows take the two-pass loop (the second read comes from L1/L2).
template <bool kSynth>
__global__ void __launch_bounds__(256) append_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                          uint64_t n, int d, int64_t ld, int normalize,
                                                          uint64_t seed, uint64_t synth_row0, int dist,
                                                          int* max_norm2_bits) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    const bool vec = !kSynth && (d & 3) == 0 && d <= 1024 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const bool regs = kSynth && d <= 1024;
    float wmax = 0.f;   // largest squared norm this warp has stored (one atomic per warp at the end)
    for (uint64_t r = warp; r < n; r += nwarps) {
        float* out = dst + r * ld;
        const float* in = kSynth ? nullptr : src + r * uint64_t(d);
        float nr = 0.f;
        float4 v4[8];
        float v1[32];
        if (vec) {
            const int d4 = d >> 2;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int c = lane + 32 * j;
                v4[j] = (c < d4) ? __ldcs(reinterpret_cast<const float4*>(in) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // same association order as the scalar loop below would give per lane? No: the norm is only
            // used for the scale, and faiss's own order is build dependent (SIMD lanes); any fp32 order is legal
#pragma unroll
            for (int j = 0; j < 8; j++) nr = dot4(v4[j], v4[j], nr);
        } else if (regs) {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int c = lane + 32 * j;
                v1[j] = (c < d) ? synth_value(seed, synth_row0 + r, uint32_t(c), dist) : 0.f;
                nr = fmaf(v1[j], v1[j], nr);
            }
        } else {
            for (int c = lane; c < d; c += kWarp) {
                float v = kSynth ? synth_value(seed, synth_row0 + r, uint32_t(c), dist) : in[c];
                nr = fmaf(v, v, nr);
            }
        }
        float inv = 1.f;
        bool scale = false;
        nr = warp_allsum(nr);
        if (normalize && nr > 0.f) {
            inv = renorm_scale(nr);
            scale = true;
        }
        // largest squared norm of any stored row (error bound of the bf16 batched path)
        wmax = fmaxf(wmax, scale ? nr * inv * inv : nr);
        if (vec) {
            const int ld4 = int(ld >> 2);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int c = lane + 32 * j;
                if (c < ld4) {
                    float4 v = v4[j];   // chunks past d/4 are zero (padding)
                    if (scale) { v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv; }
                    reinterpret_cast<float4*>(out)[c] = v;
                }
            }
        } else if (regs) {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int c = lane + 32 * j;
                if (c < ld) out[c] = scale ? v1[j] * inv : v1[j];
            }
        } else {
            for (int c = lane; c < ld; c += kWarp) {
                float v = 0.f;
                if (c < d) {
                    v = kSynth ? synth_value(seed, synth_row0 + r, uint32_t(c), dist) : in[c];
                    if (scale) v *= inv;
                }
                out[c] = v;
            }
        }
    }
    // non-negative floats order like their bit patterns
    if (lane == 0 && wmax > 0.f) atomicMax(max_norm2_bits, __float_as_int(wmax));
}

// live bitmask maintenance ---------------------------------------------------
__global__ void set_live_range_kernel(uint32_t* live, uint64_t row0, uint64_t n) {
    // one thread per 32-row word touched by [row0, row0+n)
    uint64_t w0 = row0 >> 5, w1 = (row0 + n + 31) >> 5;
    uint64_t w = w0 + uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (w >= w1) return;
    uint64_t lo = max(row0, w << 5), hi = min(row0 + n, (w + 1) << 5);
    if (hi <= lo) return;
    uint32_t nb = uint32_t(hi - lo);
    uint32_t bits = (nb == 32 ? 0xFFFFFFFFu : ((1u << nb) - 1u)) << uint32_t(lo & 31);
    atomicOr(live + w, bits);
}
__global__ void clear_live_rows_kernel(uint32_t* live, const int64_t* rows, uint64_t n) {
    uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t r = uint64_t(rows[i]);
    atomicAnd(live + (r >> 5), ~(1u << uint32_t(r & 31)));
}

// ---------------------------------------------------------------------------
// Device-side filter evaluation: a numeric metadata column lives in
// HBM next to the matrix (value + presence bit per row); a predicate becomes a bitmask without
// touching the host.  Replaces the per-row Python loops of ref vector_database.py:157-352 for
// the operators $gt $gte $lt $lte $ne and equality on numbers.  One thread per 32-row word.
// ---------------------------------------------------------------------------
enum { kOpEq = 0, kOpNe = 1, kOpGt = 2, kOpGe = 3, kOpLt = 4, kOpLe = 5 };
__global__ void predicate_mask_kernel(const double* __restrict__ vals, const uint32_t* __restrict__ has, uint64_t n,
                                      int op, double x, uint32_t* __restrict__ out, uint32_t words) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    const uint32_t present = has[w];
    uint32_t bits = 0;
    for (int j = 0; j < 32; j++) {
        const uint64_t r = uint64_t(w) * 32 + j;
        if (r >= n) break;
        if (!((present >> j) & 1u)) continue;   // only rows that HAVE the key can match (inverted index, VDB:260)
        const double v = vals[r];
        bool ok;
        switch (op) {
            case kOpEq: ok = v == x; break;
            case kOpNe: ok = v != x; break;
            case kOpGt: ok = v > x; break;
            case kOpGe: ok = v >= x; break;
            case kOpLt: ok = v < x; break;
            default: ok = v <= x; break;
        }
        bits |= uint32_t(ok) << j;
    }
    out[w] = bits;
}
// dst = dst AND src (0) / dst OR src (1) / dst AND NOT src (2); words past src_words read as 0
__global__ void combine_mask_kernel(uint32_t* dst, uint32_t dst_words, const uint32_t* __restrict__ src, uint32_t src_words,
                                    int how) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= dst_words) return;
    const uint32_t s = w < src_words ? src[w] : 0u;
    const uint32_t d = dst[w];
    dst[w] = how == 0 ? (d & s) : how == 1 ? (d | s) : (d & ~s);
}
__global__ void fill_mask_kernel(uint32_t* dst, uint64_t rows, uint32_t words) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    const uint64_t lo = uint64_t(w) * 32;
    dst[w] = rows >= lo + 32 ? 0xFFFFFFFFu : (rows > lo ? ((1u << uint32_t(rows - lo)) - 1u) : 0u);
}
// admissible AND live rows
__global__ void count_mask_kernel(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ live, uint32_t words,
                                  unsigned long long* out) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int c = 0;
    if (w < words) c = __popc(mask[w] & (live ? live[w] : 0xFFFFFFFFu));
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// order-preserving compaction: scratch[i] = x[src_rows[i]] (whole padded rows)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ x, float4* __restrict__ out,
                                                          const uint32_t* __restrict__ src_rows, uint64_t m,
                                                          int ld4) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
    for (uint64_t i = warp; i < m; i += nwarps) {
        const float4* s = x + uint64_t(src_rows[i]) * ld4;
        float4* o = out + i * ld4;
        for (int c = lane; c < ld4; c += kWarp) o[c] = s[c];
    }
}

// ---------------------------------------------------------------------------
// K5 large-k selection over the score images written by the scan in
// all_ord mode.  Keys are unique 64-bit (ord << 32 | ~row), so an 8-round
// MSB-first radix select finds the k-th largest key exactly with no tie
// handling; survivors are appended unordered and then bitonic-sorted.
// ---------------------------------------------------------------------------
struct RadixState {
    uint64_t prefix;      // bits decided so far (high to low)
    uint64_t mask;        // which bits are decided
    uint64_t remaining;   // rank still to resolve inside the current prefix (1-based)
    unsigned int hist[256];
    unsigned int out_count;
};

__device__ __forceinline__ uint64_t ord_key(uint32_t ord, uint32_t row) {
    return ord ? ((uint64_t(ord) << 32) | uint64_t(0xFFFFFFFFu - row)) : kEmptyKey;
}

__global__ void radix_init_kernel(RadixState* st, uint64_t k) {
    if (threadIdx.x == 0) {
        st->prefix = 0;
        st->mask = 0;
        st->remaining = k;
        st->out_count = 0;
    }
    if (threadIdx.x < 256) st->hist[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(512) radix_hist_kernel(const uint32_t* __restrict__ ord, uint32_t n,
                                                         RadixState* st, int shift) {
    __shared__ unsigned int h[256];
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t prefix = st->prefix, mask = st->mask;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint64_t key = ord_key(ord[i], i);
        if ((key & mask) == prefix) atomicAdd(&h[(key >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256 && h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void radix_pick_kernel(RadixState* st, int shift) {
    // single thread: walk the digits from the largest down
    if (threadIdx.x == 0) {
        uint64_t rem = st->remaining;
        int b = 255;
        for (; b > 0; b--) {
            unsigned int c = st->hist[b];
            if (rem <= c) break;
            rem -= c;
        }
        // if even digit 0 does not hold the rank (fewer than k keys match), the
        // k-th key is the empty key: digit 0, rank clamps
        st->remaining = rem;
        st->prefix |= uint64_t(b) << shift;
        st->mask |= uint64_t(0xFF) << shift;
    }
    __syncthreads();
    if (threadIdx.x < 256) st->hist[threadIdx.x] = 0;
}

// append every non-empty key >= the selected threshold (st->prefix)
__global__ void __launch_bounds__(512) radix_collect_kernel(const uint32_t* __restrict__ ord, uint32_t n,
                                                            RadixState* st, uint64_t* out, uint32_t out_cap) {
    const uint64_t thr = st->prefix;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint64_t key = ord_key(ord[i], i);
        if (key != kEmptyKey && key >= thr) {
            unsigned int pos = atomicAdd(&st->out_count, 1u);
            if (pos < out_cap) out[pos] = key;
        }
    }
}

// fill keys[count .. npad) with the empty key (count read from the device)
__global__ void pad_keys_kernel(uint64_t* keys, const RadixState* st, uint32_t npad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt = min(st->out_count, npad);
    if (i >= cnt && i < npad) keys[i] = kEmptyKey;
}

struct BlockSyncer {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// single-CTA bitonic sort (descending) of npad <= 16384 keys through shared memory
__global__ void __launch_bounds__(1024) sort_keys_smem_kernel(uint64_t* keys, uint32_t npad) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* a = reinterpret_cast<uint64_t*>(smem);
    for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) a[i] = keys[i];
    __syncthreads();
    bitonic_sort_desc(a, int(npad), int(threadIdx.x), int(blockDim.x), BlockSyncer());
    for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) keys[i] = a[i];
}

// one compare-exchange stage of a global-memory bitonic sort (any npad)
__global__ void __launch_bounds__(256) sort_keys_global_step_kernel(uint64_t* a, uint64_t npad, uint64_t size,
                                                                    uint64_t stride) {
    uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= (npad >> 1)) return;
    uint64_t i = 2 * t - (t & (stride - 1));
    uint64_t j = i + stride;
    bool desc = (i & size) == 0;
    uint64_t x = a[i], y = a[j];
    if ((x < y) == desc) {
        a[i] = y;
        a[j] = x;
    }
}

__global__ void keys_to_results_kernel(const uint64_t* keys, uint32_t navail, int64_t k, float* D, int64_t* I,
                                       int64_t label_offset) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= k) return;
    uint64_t key = (i < int64_t(navail)) ? keys[i] : kEmptyKey;
    if (key == kEmptyKey) {
        D[i] = -FLT_MAX;
        I[i] = -1;
    } else {
        D[i] = key_score(key);
        I[i] = int64_t(key_row(key)) + label_offset;
    }
}

// ---------------------------------------------------------------------------
// Fast select for 128 < k <= 8192 (host-buffer callers): two 12-bit digits of the score image fix a 24-bit
// prefix T below which no row can be among the k best; every row at or above it (k plus the few that share the k-th
// row's 24-bit prefix) is collected and ONE CTA sorts that list and writes the results.  5 launches per query (scan,
// two histograms, collect, sort) instead of the radix select's 22 -- the latter stays as the fallback for the case
// this one cannot bound: more than kFselCap rows at or above T (thousands of exact ties around the k-th score); the
// sort kernel then raises a flag and the host re-runs the query.  Every CTA derives the digits itself from the global
// histograms (a 4096-bin suffix scan), so no single-thread "pick" launches sit between the passes.
// ---------------------------------------------------------------------------
constexpr uint32_t kFselBins = 4096;
constexpr uint32_t kFselCap = 16384;
struct FselState {
    unsigned int h0[kFselBins];   // rows per top-12-bit digit of the score image (empty rows, image 0, are not counted)
    unsigned int h1[kFselBins];   // rows per next-12-bit digit among the rows whose top digit is the selected one
    unsigned int out_count;
    unsigned int overflow;
};

// The digit that holds the rem-th largest row (counting from the top bin down) and the rank left inside it.  All
// threads of the block call this (blockDim.x == 512, 8 bins per thread); fewer than rem rows: digit 0, rank clamps.
__device__ __forceinline__ uint32_t fsel_find_digit(const unsigned int* hist, uint32_t rem, uint32_t* rem_out, unsigned int* sh) {
    const int t = threadIdx.x;   // thread t owns the bins 4095 - 8t ... 4088 - 8t
    unsigned int c[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        c[j] = __ldcg(hist + (kFselBins - 1 - (8 * t + j)));
        sum += c[j];
    }
    sh[t] = sum;
    __syncthreads();
    for (int o = 1; o < 512; o <<= 1) {   // inclusive scan over the threads
        const unsigned int v = (t >= o) ? sh[t - o] : 0u;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    const unsigned int incl = sh[t], excl = incl - sum, total = sh[511];
    __syncthreads();
    if (t == 0) {
        sh[0] = 0u;     // digit
        sh[1] = rem > total ? 0xFFFFFFFFu : 0u;   // rank (beyond every count: the next level picks digit 0 too), overwritten below when it exists
    }
    __syncthreads();
    if (excl < rem && rem <= incl) {
        unsigned int r = rem - excl;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (r <= c[j]) {
                sh[0] = kFselBins - 1 - (8 * t + j);
                sh[1] = r;
                break;
            }
            r -= c[j];
        }
    }
    __syncthreads();
    const uint32_t digit = sh[0];
    *rem_out = sh[1];
    __syncthreads();
    return digit;
}

// level 0: histogram of the top digit; level 1: of the second digit among the rows of the selected top digit
__global__ void __launch_bounds__(512) fsel_hist_kernel(const uint32_t* __restrict__ ord, uint32_t n, FselState* st, uint32_t k,
                                                        int level) {
    __shared__ unsigned int h[kFselBins];
    __shared__ unsigned int sh[512];
    for (uint32_t i = threadIdx.x; i < kFselBins; i += blockDim.x) h[i] = 0u;
    uint32_t d0 = 0, rem = 0;
    if (level == 1) d0 = fsel_find_digit(st->h0, k, &rem, sh);
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t o = ord[i];
        if (o == 0u) continue;
        if (level == 0) atomicAdd(&h[o >> 20], 1u);
        else if ((o >> 20) == d0) atomicAdd(&h[(o >> 8) & 0xFFFu], 1u);
    }
    __syncthreads();
    unsigned int* dst = level == 0 ? st->h0 : st->h1;
    for (uint32_t i = threadIdx.x; i < kFselBins; i += blockDim.x)
        if (h[i]) atomicAdd(&dst[i], h[i]);
}

// every row whose 24-bit prefix is at or above the selected one
__global__ void __launch_bounds__(512) fsel_collect_kernel(const uint32_t* __restrict__ ord, uint32_t n, FselState* st, uint32_t k,
                                                           uint64_t* out) {
    __shared__ unsigned int sh[512];
    uint32_t rem = 0, rem2 = 0;
    const uint32_t d0 = fsel_find_digit(st->h0, k, &rem, sh);
    const uint32_t d1 = fsel_find_digit(st->h1, rem, &rem2, sh);
    const uint32_t thr = (d0 << 12) | d1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t o = ord[i];
        if (o != 0u && (o >> 8) >= thr) {
            const unsigned int pos = atomicAdd(&st->out_count, 1u);
            if (pos < kFselCap) out[pos] = ord_key(o, i);
        }
    }
}

// one CTA: sort the collected keys, write (D, I), leave the state clean for the next query
__global__ void __launch_bounds__(1024) fsel_sort_results_kernel(const uint64_t* __restrict__ keys, FselState* st, int64_t k, float* D,
                                                                  int64_t* I, int64_t label_offset, unsigned int* ovf_host) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* a = reinterpret_cast<uint64_t*>(smem);
    const unsigned int cnt = st->out_count;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kFselBins; i += blockDim.x) {
        st->h0[i] = 0u;
        st->h1[i] = 0u;
    }
    if (threadIdx.x == 0) st->out_count = 0u;
    if (cnt > kFselCap) {   // too many rows share the k-th row's prefix: the radix select answers (the host re-runs the query)
        if (threadIdx.x == 0) {
            st->overflow = 1u;
            if (ovf_host) {
                *reinterpret_cast<volatile unsigned int*>(ovf_host) = 1u;
                __threadfence_system();
            }
        }
        return;
    }
    uint32_t npad = 64;
    while (npad < cnt) npad <<= 1;
    for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) a[i] = (i < cnt) ? keys[i] : kEmptyKey;
    __syncthreads();
    bitonic_sort_desc(a, int(npad), int(threadIdx.x), int(blockDim.x), BlockSyncer());
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = (i < int64_t(cnt)) ? a[i] : kEmptyKey;
        if (key == kEmptyKey) {
            D[i] = -FLT_MAX;
            I[i] = -1;
        } else {
            D[i] = key_score(key);
            I[i] = int64_t(key_row(key)) + label_offset;
        }
    }
}

// ---------------------------------------------------------------------------
// shard merge (sharded path): per query, merge nparts best-first lists.
// Key = (score image, ~position) with position = part * k + i, so equal scores
// keep shard order then in-shard order (= ascending global row when shards
// are contiguous row blocks).  One CTA per query; npad keys in shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_topk_kernel(const float* __restrict__ Dp, const int64_t* __restrict__ Ip,
                                                         int nparts, int64_t nq, int64_t k, uint32_t npad,
                                                         float* D, int64_t* I) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* a = reinterpret_cast<uint64_t*>(smem);
    const int64_t qi = blockIdx.x;
    const uint32_t total = uint32_t(nparts * k);
    for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) {
        uint64_t key = kEmptyKey;
        if (i < total) {
            uint32_t part = i / uint32_t(k), j = i % uint32_t(k);
            size_t off = (size_t(part) * nq + qi) * k + j;
            if (Ip[off] >= 0) key = make_key(Dp[off], i);
        }
        a[i] = key;
    }
    __syncthreads();
    bitonic_sort_desc(a, int(npad), int(threadIdx.x), int(blockDim.x), BlockSyncer());
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t key = a[i];
        if (key == kEmptyKey) {
            D[qi * k + i] = -FLT_MAX;
            I[qi * k + i] = -1;
        } else {
            uint32_t pos = key_row(key);
            uint32_t part = pos / uint32_t(k), j = pos % uint32_t(k);
            size_t off = (size_t(part) * nq + qi) * k + j;
            D[qi * k + i] = Dp[off];
            I[qi * k + i] = Ip[off];
        }
    }
}

}  // namespace mvdb
