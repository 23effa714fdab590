// device_utils.cuh -- PTX wrappers and key encoding shared by every kernel.
// sm_100a only (cp.async.bulk, mbarrier, createpolicy).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mvdb {

constexpr int kWarp = 32;
constexpr uint64_t kEmptyKey = 0ull;      // "no result": sorts below every real score
constexpr int kRowsPerTile = 8;           // rows one warp reduces per step (one mask byte)

// ---------------------------------------------------------------------------
// 64-bit selection key: high word = order-preserving image of the fp32 score,
// low word = ~row.  A larger key is a better result: higher score first, exact
// ties by ascending row (the order a sequential faiss scan with strict '>'
// keeps at the boundary; ref vector_database.py:497).  NaN scores are dropped.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t score_to_ord(float s) {
    s += 0.0f;  // -0.0 -> +0.0 so that equal floats have equal images
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_to_score(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
__device__ __forceinline__ uint64_t make_key(float s, uint32_t row) {
    if (s != s) return kEmptyKey;
    return (uint64_t(score_to_ord(s)) << 32) | uint64_t(0xFFFFFFFFu - row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t key) { return 0xFFFFFFFFu - uint32_t(key); }
__device__ __forceinline__ float key_score(uint64_t key) { return ord_to_score(uint32_t(key >> 32)); }

// ---------------------------------------------------------------------------
// shared-memory addresses, mbarrier, bulk async copy (TMA 1-D)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the initialised barriers visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// L2 eviction policy for data that is streamed once per scan
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// global -> shared bulk copy; completion is signalled on `bar` (complete_tx).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// streaming 128-bit global load that does not allocate in L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
    return acc;
}

// Sum 8 per-lane partials across the warp in 9 shuffles.  On return every
// lane holds the full sum of row  tile_row_of_lane(lane)  (all 4 lanes of a
// quad hold the same value).  The association order is fixed, so a row's
// score does not depend on which warp/CTA computed it.
__device__ __forceinline__ int tile_row_of_lane(int lane) {
    return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}
__device__ __forceinline__ float reduce8(const float (&a)[8], int lane) {
    const unsigned full = 0xFFFFFFFFu;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    float b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float mine = h16 ? a[i + 4] : a[i];
        float give = h16 ? a[i] : a[i + 4];
        b[i] = mine + __shfl_xor_sync(full, give, 16);
    }
    float c[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        float mine = h8 ? b[i + 2] : b[i];
        float give = h8 ? b[i] : b[i + 2];
        c[i] = mine + __shfl_xor_sync(full, give, 8);
    }
    float mine = h4 ? c[1] : c[0];
    float give = h4 ? c[0] : c[1];
    float s = mine + __shfl_xor_sync(full, give, 4);
    s += __shfl_xor_sync(full, s, 2);
    s += __shfl_xor_sync(full, s, 1);
    return s;
}
// all-lanes sum (query norm)
__device__ __forceinline__ float warp_allsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
// faiss fvec_renorm_L2's scale: (float)(1.0 / sqrtf(nr))
__device__ __forceinline__ float renorm_scale(float nr) {
    return (float)(1.0 / (double)sqrtf(nr));
}

// splitmix64 finaliser -- the counter-based synthetic generator
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ float synth_value(uint64_t seed, uint64_t row, uint32_t col,
                                                      int dist) {
    uint64_t z = mix64((row << 20) + seed * 0x9E3779B97F4A7C15ull + col);
    if (dist == 0) {
        int s = int(z & 0xFFFF) + int((z >> 16) & 0xFFFF) + int((z >> 32) & 0xFFFF) + int(z >> 48);
        return float(s - 131070) * (1.0f / 65536.0f);
    }
    return float(z >> 40) * (1.0f / 16777216.0f);
}

}  // namespace mvdb
